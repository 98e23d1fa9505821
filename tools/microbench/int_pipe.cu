// int_pipe.cu -- issue-rate microbenchmark of the integer instructions the SAD kernels live on
// ("Peak SIMD-int issue rate on sm_100 is not in MEASURED_PEAKS.json").
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe int_pipe.cu && ./int_pipe
// Every thread runs ILP independent dependency chains of one instruction; the result is printed as
// thread-instructions per clock per SM (128 = one warp instruction per cycle on each of 4 SMSPs).
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ILP = 8, ITERS = 4096;

template <int OP>
__global__ void k(unsigned* out, unsigned seed, long long* cycles) {
  unsigned r[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) r[i] = seed + threadIdx.x * 31 + i;
  unsigned b = seed * 7 + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int i = 0; i < ILP; i++) {
      if (OP == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(r[(i + 1) % ILP]));
      if (OP == 1) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));
      if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(seed));
      if (OP == 3) asm volatile("min.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b ^ (unsigned)it));
      if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(b), "r"(seed));
      if (OP == 5) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(b), "r"(seed));
      if (OP == 6) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(seed));
    }
  }
  long long t1 = clock64();
  unsigned s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s ^= r[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps_per_sm) {
  int sms;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int threads = warps_per_sm * 32;
  unsigned* out;
  long long* cyc;
  cudaMalloc(&out, sms * threads * 4);
  cudaMalloc(&cyc, sms * 8);
  k<OP><<<sms, threads>>>(out, 12345u, cyc);
  cudaDeviceSynchronize();
  k<OP><<<sms, threads>>>(out, 12345u, cyc);
  cudaDeviceSynchronize();
  long long h[1024];
  cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < sms; i++) avg += h[i];
  avg /= sms;
  double per_clk = (double)threads * ILP * ITERS / avg;
  printf("%-22s warps/SM %2d : %7.1f thread-instr/clk/SM  (%.2f warp-instr/clk/SM)\n", name, warps_per_sm, per_clk,
         per_clk / 32);
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  for (int w : {4, 16, 32}) {
    run<0>("VABSDIFF4.ACC", w);
    run<6>("VABSDIFF4 (no acc)", w);
    run<1>("IADD", w);
    run<2>("IMAD", w);
    run<3>("VIMNMX (min.u32)", w);
    run<4>("LOP3", w);
    run<5>("IDP4A", w);
  }
  return 0;
}
