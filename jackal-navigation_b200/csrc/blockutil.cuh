// blockutil.cuh -- CTA-wide helpers (scan, async bulk copy, mbarrier).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// In-place exclusive prefix sum of data[0..n) by the whole CTA; returns the total.
// `part` is shared scratch of blockDim.x + 1 ints.  All threads must call.
template <typename E>
__device__ inline int block_exclusive_scan(E* data, int n, int* part) {
  const int T = blockDim.x, t = threadIdx.x;
  const int chunk = (n + T - 1) / T;
  const int lo = min(t * chunk, n), hi = min(lo + chunk, n);
  int s = 0;
  for (int i = lo; i < hi; i++) s += data[i];
  // inclusive scan of the T partials: shuffle scan inside each warp, then over the warp totals
  const int lane = t & 31, w = t >> 5, nw = (T + 31) >> 5;
  int inc = s;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    int v = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += v;
  }
  if (lane == 31) part[w] = inc;          // part[0..nw) = warp totals
  __syncthreads();
  if (w == 0) {
    int wt = (lane < nw) ? part[lane] : 0;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      int v = __shfl_up_sync(0xffffffffu, wt, off);
      if (lane >= off) wt += v;
    }
    if (lane < nw) part[lane] = wt;       // inclusive scan of the warp totals (nw <= 32)
  }
  __syncthreads();
  const int total = part[nw - 1];
  int run = inc - s + (w ? part[w - 1] : 0);   // exclusive prefix of this thread's chunk
  __syncthreads();
  for (int i = lo; i < hi; i++) {
    int v = data[i];
    data[i] = (E)run;
    run += v;
  }
  __syncthreads();
  return total;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// ---- mbarrier + 1-D bulk async copy (TMA unit, UBLKCP in SASS) -------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  } while (!done);
}
