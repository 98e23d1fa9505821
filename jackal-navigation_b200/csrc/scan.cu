// scan.cu -- disparity -> 3-D -> robot frame -> 90-bin obstacle scan, fused.
//
// Replaces, from src/obstacle_avoidance/point_cloud.cpp:
//   cacheDisparityValues                   :104-147   (gate_cache_kernel)
//   generateDisparityMap's convertTo(CV_8U) :421-422   (round half to even, saturate)
//   publishObstacleScan(Mat& dmap)          :213-296   (scan_kernel: one pass over the
//       float disparity map does the u8 conversion, the gate test, Q*[u,v,d,1],
//       XR*p+XT, atan2/sqrt and the per-bin minimum)
//   publishPointCloud (-g)                  :314-387   (points kernels, column-major order)
//   publishObstacleScan(vector<Point3d>)    :149-211   (scan_points_kernel)
//
// Arithmetic is double precision in the reference's operation order (compiled
// with -fmad=false); matrix products are evaluated left to right like OpenCV's
// small-matrix gemm.  Per-bin minima and the angle/range extrema are reduced in
// shared memory per CTA and merged with 64-bit atomics on order-preserving
// keys; a bin index outside [0,89] is skipped.
// Roofline: HBM, 4 N bytes read per frame (+1 N if the u8 map is written).
#include "common.cuh"
#include "blockutil.cuh"
#include <math.h>

struct ScanConst {
  double Q[16], XR[9], XT[3];
  double tan_gp;          // tan(4 * 3.1415 / 180)
  int W, H, ox, oy;
  int q_simple;           // Q has stereoRectify's sparsity: only Q03, Q13, Q23, Q32 and the two unit entries
};

namespace {

constexpr double SCAN_INF = 1e9;
constexpr double GP_HEIGHT = 0.05, GP_DIST = 1.0;

// order-preserving map double -> u64 (works for negative values too)
__device__ __forceinline__ unsigned long long okey(double x) {
  unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double okey_inv(unsigned long long k) {
  unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)b);
}

__device__ __forceinline__ void reproject(const ScanConst& c, double x, double y, double d, double out[3]) {
  double pos[4];
  if (c.q_simple) {
    // Q = [[1,0,0,Q03],[0,1,0,Q13],[0,0,0,Q23],[0,0,Q32,0]] (point_cloud.cpp:543, zero-disparity
    // flag): the products with the exact 0 and 1 entries change nothing, so these are the general
    // expressions below bit for bit (the + 0.0 keeps pos[3] = +0 for d = 0 as ((0+0)+Q32*d)+0 gives)
    pos[0] = x + c.Q[3];
    pos[1] = y + c.Q[7];
    pos[2] = c.Q[11];
    pos[3] = c.Q[14] * d + 0.0;
  } else {
#pragma unroll
    for (int i = 0; i < 4; i++)
      pos[i] = ((c.Q[4 * i] * x + c.Q[4 * i + 1] * y) + c.Q[4 * i + 2] * d) + c.Q[4 * i + 3] * 1.0;
  }
  double X = pos[0] / pos[3], Y = pos[1] / pos[3], Z = pos[2] / pos[3];
#pragma unroll
  for (int i = 0; i < 3; i++) out[i] = ((c.XR[3 * i] * X + c.XR[3 * i + 1] * Y) + c.XR[3 * i + 2] * Z) + c.XT[i];
}

__device__ __forceinline__ bool above_ground(const ScanConst& c, double X, double Z) {
  if (X < GP_DIST) return !(Z < GP_HEIGHT);
  return !(Z < GP_HEIGHT + c.tan_gp * (X - GP_DIST));
}

__global__ void gate_cache_kernel(ScanConst c, uint8_t* __restrict__ gate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i >= c.W) return;
  int d;
  for (d = 3; d <= 255; d++) {
    double r[3];
    reproject(c, (double)(i + c.ox), (double)(j + c.oy), (double)d, r);
    if (r[2] < 0.) continue;
    if (!above_ground(c, r[0], r[2])) continue;
    break;
  }
  gate[2 * ((size_t)j * c.W + i)] = (uint8_t)d;   // 256 wraps to 0 like the reference's uchar store
  gate[2 * ((size_t)j * c.W + i) + 1] = 255;
}

// per-frame accumulator: 90 range keys, then angle min/max, range min/max keys, n_points
constexpr int ACC_WORDS = JN_SCAN_BINS + 5;

struct BlockAcc {
  unsigned long long bins[JN_SCAN_BINS];
  unsigned long long amin, amax, rmin, rmax;
  unsigned int npts;
};

__device__ __forceinline__ void acc_init(BlockAcc& a, int tid, int nthreads) {
  for (int k = tid; k < JN_SCAN_BINS; k += nthreads) a.bins[k] = ~0ull;
  if (tid == 0) { a.amin = ~0ull; a.amax = 0ull; a.rmin = ~0ull; a.rmax = 0ull; a.npts = 0u; }
}

// Thread-private accumulator: a thread walks down one image column, so its bin index
// changes slowly; the running (bin, min range) pair lives in registers and only a bin
// change touches shared memory.  Extrema stay in registers until the end.
struct LocalAcc {
  unsigned long long amin, amax, rmin, rmax, curkey;
  unsigned int npts;
  int curbin;
};
__device__ __forceinline__ void lacc_init(LocalAcc& l) {
  l.amin = ~0ull; l.amax = 0ull; l.rmin = ~0ull; l.rmax = 0ull; l.curkey = ~0ull; l.npts = 0u; l.curbin = -1;
}
__device__ __forceinline__ void lacc_add(LocalAcc& l, BlockAcc& a, double X, double Y) {
  double th = atan2(Y, X);
  double deg = th * 180. / 3.1415;
  // defined here: a NaN reprojection (d = 0 through a wrapped gate divides by W = 0, H8) is skipped
  if (th != th || X != X || Y != Y) return;
  // ranges are tracked as r^2 (same roundings as the reference's sqrt argument); sqrt is monotone,
  // so min/max of sqrt(r2) = sqrt of min/max r2 and the root is taken once per bin at the end
  const double r2 = Y * Y + X * X;
  unsigned long long kt = okey(th), kr = okey(r2);
  l.amin = min(l.amin, kt);
  l.amax = max(l.amax, kt);
  l.rmin = min(l.rmin, kr);
  l.rmax = max(l.rmax, kr);
  l.npts++;
  double kf = floor((double)JN_SCAN_BINS * (90. / 2. - deg) / 90.);
  if (kf >= 0 && kf < (double)JN_SCAN_BINS) {
    int k = (int)kf;
    if (k != l.curbin) {
      if (l.curbin >= 0) atomicMin(&a.bins[l.curbin], l.curkey);
      l.curbin = k;
      l.curkey = kr;
    } else {
      l.curkey = min(l.curkey, kr);
    }
  }
}
// warp-reduce the extrema, one shared atomic per warp and value
__device__ __forceinline__ void lacc_finish(LocalAcc& l, BlockAcc& a) {
  if (l.curbin >= 0) atomicMin(&a.bins[l.curbin], l.curkey);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    l.amin = min(l.amin, __shfl_xor_sync(0xffffffffu, l.amin, off));
    l.amax = max(l.amax, __shfl_xor_sync(0xffffffffu, l.amax, off));
    l.rmin = min(l.rmin, __shfl_xor_sync(0xffffffffu, l.rmin, off));
    l.rmax = max(l.rmax, __shfl_xor_sync(0xffffffffu, l.rmax, off));
    l.npts += __shfl_xor_sync(0xffffffffu, l.npts, off);
  }
  if ((threadIdx.x & 31) == 0 && l.npts) {
    atomicMin(&a.amin, l.amin);
    atomicMax(&a.amax, l.amax);
    atomicMin(&a.rmin, l.rmin);
    atomicMax(&a.rmax, l.rmax);
    atomicAdd(&a.npts, l.npts);
  }
}

__device__ __forceinline__ void acc_flush(const BlockAcc& a, unsigned long long* g, int tid, int nthreads) {
  for (int k = tid; k < JN_SCAN_BINS; k += nthreads)
    if (a.bins[k] != ~0ull) atomicMin(&g[k], a.bins[k]);
  if (tid == 0 && a.npts) {
    atomicMin(&g[JN_SCAN_BINS + 0], a.amin);
    atomicMax(&g[JN_SCAN_BINS + 1], a.amax);
    atomicMin(&g[JN_SCAN_BINS + 2], a.rmin);
    atomicMax(&g[JN_SCAN_BINS + 3], a.rmax);
    atomicAdd(&g[JN_SCAN_BINS + 4], (unsigned long long)a.npts);
  }
}

__global__ void acc_reset_kernel(unsigned long long* acc, int n_frames) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_frames * ACC_WORDS) return;
  int k = i % ACC_WORDS;
  unsigned long long v = ~0ull;
  if (k == JN_SCAN_BINS + 1 || k == JN_SCAN_BINS + 3 || k == JN_SCAN_BINS + 4) v = 0ull;
  acc[i] = v;
}

__device__ __forceinline__ int to_u8(float x) { return min(max(__float2int_rn(x), 0), 255); }

constexpr int SCAN_THREADS = 256;

// grid (ceil(W/256), rows-per-block tiles, frames)
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(ScanConst c, const float* __restrict__ D, const uint8_t* __restrict__ gate,
            unsigned long long* __restrict__ acc, uint8_t* __restrict__ dmap_u8, int rows_per_block) {
  __shared__ BlockAcc s;
  const int tid = threadIdx.x, frame = blockIdx.z;
  acc_init(s, tid, SCAN_THREADS);
  __syncthreads();
  const int i = blockIdx.x * SCAN_THREADS + tid;
  const size_t fp = (size_t)frame * c.W * c.H;
  LocalAcc l;
  lacc_init(l);
  if (i < c.W) {
    const int j0 = blockIdx.y * rows_per_block, j1 = min(j0 + rows_per_block, c.H);
    for (int j = j0; j < j1; j++) {
      const size_t a = (size_t)j * c.W + i;
      const int d = to_u8(D[fp + a]);
      if (dmap_u8) dmap_u8[fp + a] = (uint8_t)d;
      const int g0 = gate[2 * a], g1 = gate[2 * a + 1];
      if (d < g0 || d > g1) continue;
      double r[3];
      reproject(c, (double)(i + c.ox), (double)(j + c.oy), (double)d, r);
      lacc_add(l, s, r[0], r[1]);
    }
  }
  lacc_finish(l, s);
  __syncthreads();
  acc_flush(s, acc + (size_t)frame * ACC_WORDS, tid, SCAN_THREADS);
}

__global__ void scan_finalize_kernel(const unsigned long long* __restrict__ acc, double* __restrict__ ranges,
                                     jn_scan_meta* __restrict__ meta) {
  __shared__ int s_fin;
  const int frame = blockIdx.x, tid = threadIdx.x;
  const unsigned long long* a = acc + (size_t)frame * ACC_WORDS;
  if (tid == 0) s_fin = 0;
  __syncthreads();
  if (tid < JN_SCAN_BINS) {
    double r = (a[tid] == ~0ull) ? SCAN_INF : sqrt(okey_inv(a[tid]));
    ranges[(size_t)frame * JN_SCAN_BINS + tid] = r;
    if (r < SCAN_INF - 1) atomicAdd(&s_fin, 1);
  }
  __syncthreads();
  if (tid == 0) {
    jn_scan_meta m;
    unsigned long long n = a[JN_SCAN_BINS + 4];
    m.angle_min = n ? okey_inv(a[JN_SCAN_BINS + 0]) : 400.;
    m.angle_max = n ? okey_inv(a[JN_SCAN_BINS + 1]) : -400.;
    m.range_min = n ? sqrt(okey_inv(a[JN_SCAN_BINS + 2])) : SCAN_INF;
    m.range_max = n ? sqrt(okey_inv(a[JN_SCAN_BINS + 3])) : -500.;
    m.n_finite = s_fin;
    m.n_points = (int)n;
    meta[frame] = m;
  }
}

// ---- -g path: every pixel with d >= 2, emitted columns-outer like the reference ----
__global__ void points_count_kernel(ScanConst c, const float* __restrict__ D, int* __restrict__ colcount) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.W) return;
  int n = 0;
  for (int j = 0; j < c.H; j++) n += to_u8(D[(size_t)j * c.W + i]) >= 2;
  colcount[i] = n;
}
__global__ void points_scan_kernel(int* colcount, int W, int* total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int run = 0;
    for (int i = 0; i < W; i++) { int v = colcount[i]; colcount[i] = run; run += v; }
    *total = run;
  }
}
__global__ void points_write_kernel(ScanConst c, const float* __restrict__ D, const int* __restrict__ coloff,
                                    double* __restrict__ pts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.W) return;
  size_t k = coloff[i];
  for (int j = 0; j < c.H; j++) {
    int d = to_u8(D[(size_t)j * c.W + i]);
    if (d < 2) continue;
    double r[3];
    reproject(c, (double)(i + c.ox), (double)(j + c.oy), (double)d, r);
    pts[3 * k] = r[0]; pts[3 * k + 1] = r[1]; pts[3 * k + 2] = r[2];
    k++;
  }
}
// Point32 + rgb channel of every point, in the order of points_write_kernel (point_cloud.cpp:351-383)
__global__ void cloud_pack_kernel(ScanConst c, const float* __restrict__ D, const int* __restrict__ coloff,
                                  const double* __restrict__ pts, const uint8_t* __restrict__ img, int stride,
                                  int channels, float* __restrict__ xyz, float* __restrict__ rgb) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c.W) return;
  size_t k = coloff[i];
  const size_t img_bytes = (size_t)stride * c.H;
  for (int j = 0; j < c.H; j++) {
    if (to_u8(D[(size_t)j * c.W + i]) < 2) continue;
    xyz[3 * k] = (float)pts[3 * k];
    xyz[3 * k + 1] = (float)pts[3 * k + 1];
    xyz[3 * k + 2] = (float)pts[3 * k + 2];
    int ch[3];
#pragma unroll
    for (int b = 0; b < 3; b++) {
      const size_t a = (size_t)j * stride + 3 * (size_t)i + b;
      ch[b] = (channels == 3 || a < img_bytes) ? img[a] : 0;
    }
    rgb[k] = __int_as_float((ch[2] << 16) | (ch[1] << 8) | ch[0]);
    k++;
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_points_kernel(ScanConst c, const double* __restrict__ pts, const int* __restrict__ n_ptr,
                   unsigned long long* __restrict__ acc) {
  __shared__ BlockAcc s;
  const int tid = threadIdx.x;
  acc_init(s, tid, SCAN_THREADS);
  __syncthreads();
  const int n = *n_ptr;
  LocalAcc l;
  lacc_init(l);
  for (int k = blockIdx.x * SCAN_THREADS + tid; k < n; k += gridDim.x * SCAN_THREADS) {
    double X = pts[3 * (size_t)k], Y = pts[3 * (size_t)k + 1], Z = pts[3 * (size_t)k + 2];
    if (!above_ground(c, X, Z)) continue;
    lacc_add(l, s, X, Y);
  }
  lacc_finish(l, s);
  __syncthreads();
  acc_flush(s, acc, tid, SCAN_THREADS);
}

// ---- -g path, batched and device resident: grid.y = frame -------------------------------------
// 1. points per column (d >= 2), 2. exclusive scan over the columns of each frame, 3. one thread per
// column writes its points (Point32 + packed rgb, point_cloud.cpp:351-383) at its offset -- the
// reference's order, columns outer -- and feeds the same double-precision points to the scan
// accumulators with the ground gate of publishObstacleScan(vector<Point3d>) (point_cloud.cpp:166-172).
__global__ void pc_count_kernel(ScanConst c, const float* __restrict__ D, int* __restrict__ colcount) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x, frame = blockIdx.y;
  if (i >= c.W) return;
  const float* Df = D + (size_t)frame * c.W * c.H;
  int n = 0;
  for (int j = 0; j < c.H; j++) n += to_u8(Df[(size_t)j * c.W + i]) >= 2;
  colcount[(size_t)frame * c.W + i] = n;
}
__global__ void __launch_bounds__(256) pc_scan_kernel(int* __restrict__ colcount, int W, int32_t* __restrict__ totals) {
  __shared__ int part[256 + 1];
  int* col = colcount + (size_t)blockIdx.x * W;
  const int total = block_exclusive_scan(col, W, part);
  if (threadIdx.x == 0) totals[blockIdx.x] = total;
}
__global__ void __launch_bounds__(SCAN_THREADS)
pc_write_kernel(ScanConst c, const float* __restrict__ D, const int* __restrict__ coloff, const uint8_t* __restrict__ img,
                int stride, int channels, size_t img_frame_bytes, float* __restrict__ xyz, float* __restrict__ rgb,
                unsigned long long* __restrict__ acc) {
  __shared__ BlockAcc s;
  const int tid = threadIdx.x, frame = blockIdx.y;
  acc_init(s, tid, SCAN_THREADS);
  __syncthreads();
  const int i = blockIdx.x * SCAN_THREADS + tid;
  const size_t n = (size_t)c.W * c.H;
  LocalAcc l;
  lacc_init(l);
  if (i < c.W) {
    const float* Df = D + (size_t)frame * n;
    float* xf = xyz + (size_t)frame * n * 3;
    float* cf = rgb ? rgb + (size_t)frame * n : nullptr;
    const uint8_t* im = img ? img + (size_t)frame * img_frame_bytes : nullptr;
    const size_t img_bytes = (size_t)stride * c.H;
    size_t k = coloff[(size_t)frame * c.W + i];
    for (int j = 0; j < c.H; j++) {
      const int d = to_u8(Df[(size_t)j * c.W + i]);
      if (d < 2) continue;
      double r[3];
      reproject(c, (double)(i + c.ox), (double)(j + c.oy), (double)d, r);
      xf[3 * k] = (float)r[0]; xf[3 * k + 1] = (float)r[1]; xf[3 * k + 2] = (float)r[2];
      if (cf) {
        int ch[3] = {0, 0, 0};
        if (im) {
#pragma unroll
          for (int b = 0; b < 3; b++) {
            const size_t a = (size_t)j * stride + 3 * (size_t)i + b;
            ch[b] = (channels == 3 || a < img_bytes) ? im[a] : 0;
          }
        }
        cf[k] = __int_as_float((ch[2] << 16) | (ch[1] << 8) | ch[0]);
      }
      k++;
      if (above_ground(c, r[0], r[2])) lacc_add(l, s, r[0], r[1]);
    }
  }
  lacc_finish(l, s);
  __syncthreads();
  acc_flush(s, acc + (size_t)frame * ACC_WORDS, tid, SCAN_THREADS);
}

}  // namespace

// ---- host-side object ---------------------------------------------------------------------
struct jn_scan {
  ScanConst c;
  int device;
  uint8_t* gate;                 // W*H*2
  unsigned long long* acc;       // cap_frames * ACC_WORDS
  int cap_frames;
  // single-frame scratch
  float* dD; double* dRanges; jn_scan_meta* dMeta; uint8_t* dU8;
  int* dCol; int* dTotal; double* dPts;
  // -g path scratch, allocated on first use and kept (no cudaMalloc per frame)
  uint8_t* dImg; size_t img_bytes; float* dXyz;
  int* dColBatch; int col_frames;   // batched -g path: per-column counts / offsets of n frames
};

static int ensure_acc(jn_scan* s, int n) {
  if (n <= s->cap_frames) return JN_OK;
  if (s->acc) cudaFree(s->acc);
  s->acc = nullptr;
  JN_CUDA_CHECK(cudaMalloc(&s->acc, (size_t)n * ACC_WORDS * sizeof(unsigned long long)));
  s->cap_frames = n;
  return JN_OK;
}

extern "C" jn_scan* jn_scan_create(const jn_calib* cal, int width, int height, int ox, int oy, int device) {
  if (!cal || !cal->has_q || width <= 0 || height <= 0) { jn_set_error("jn_scan_create: bad arguments"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { jn_set_error("jn_scan_create: no CUDA device %d", device); return nullptr; }
  jn_scan* s = new jn_scan();
  memset(s, 0, sizeof(*s));
  s->device = device;
  for (int i = 0; i < 16; i++) s->c.Q[i] = cal->Q[i];
  for (int i = 0; i < 9; i++) s->c.XR[i] = cal->XR[i];
  for (int i = 0; i < 3; i++) s->c.XT[i] = cal->XT[i];
  s->c.tan_gp = tan(4. * 3.1415 / 180.);
  {
    const double* q = s->c.Q;
    s->c.q_simple = q[0] == 1 && q[1] == 0 && q[2] == 0 && q[4] == 0 && q[5] == 1 && q[6] == 0 && q[8] == 0 &&
                    q[9] == 0 && q[10] == 0 && q[12] == 0 && q[13] == 0 && q[15] == 0;
  }
  s->c.W = width; s->c.H = height; s->c.ox = ox; s->c.oy = oy;
  size_t n = (size_t)width * height;
  if (cudaMalloc(&s->gate, 2 * n) != cudaSuccess || cudaMalloc(&s->dD, n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&s->dRanges, JN_SCAN_BINS * sizeof(double)) != cudaSuccess ||
      cudaMalloc(&s->dMeta, sizeof(jn_scan_meta)) != cudaSuccess || cudaMalloc(&s->dU8, n) != cudaSuccess ||
      cudaMalloc(&s->dCol, width * sizeof(int)) != cudaSuccess || cudaMalloc(&s->dTotal, sizeof(int)) != cudaSuccess ||
      cudaMalloc(&s->dPts, n * 3 * sizeof(double)) != cudaSuccess) {
    jn_set_error("jn_scan_create: cudaMalloc failed");
    jn_scan_destroy(s);
    return nullptr;
  }
  gate_cache_kernel<<<dim3((width + 127) / 128, height), 128>>>(s->c, s->gate);
  g_jn_launches += 1;
  if (cudaDeviceSynchronize() != cudaSuccess) {
    jn_set_error("jn_scan_create: gate cache kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    jn_scan_destroy(s);
    return nullptr;
  }
  return s;
}

extern "C" void jn_scan_destroy(jn_scan* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  cudaFree(s->dImg); cudaFree(s->dXyz); cudaFree(s->dColBatch);
  cudaFree(s->gate); cudaFree(s->acc); cudaFree(s->dD); cudaFree(s->dRanges); cudaFree(s->dMeta);
  cudaFree(s->dU8); cudaFree(s->dCol); cudaFree(s->dTotal); cudaFree(s->dPts);
  delete s;
}

extern "C" int jn_scan_gate_cache(jn_scan* s, uint8_t* out) {
  if (!s || !out) return JN_ERR_ARG;
  JN_CUDA_CHECK(cudaSetDevice(s->device));
  JN_CUDA_CHECK(cudaMemcpy(out, s->gate, 2 * (size_t)s->c.W * s->c.H, cudaMemcpyDeviceToHost));
  return JN_OK;
}

extern "C" int jn_scan_from_disparity_batch(jn_scan* s, int n, const float* D, double* ranges, jn_scan_meta* meta,
                                            uint8_t* dmap_u8, void* stream) {
  if (!s || n <= 0 || !D || !ranges || !meta) return JN_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  JN_CUDA_CHECK(cudaSetDevice(s->device));
  int rc = ensure_acc(s, n);
  if (rc) return rc;
  acc_reset_kernel<<<(n * ACC_WORDS + 255) / 256, 256, 0, st>>>(s->acc, n);
  const int rows_per_block = 16;
  dim3 grid((s->c.W + SCAN_THREADS - 1) / SCAN_THREADS, (s->c.H + rows_per_block - 1) / rows_per_block, n);
  scan_kernel<<<grid, SCAN_THREADS, 0, st>>>(s->c, D, s->gate, s->acc, dmap_u8, rows_per_block);
  scan_finalize_kernel<<<n, 96, 0, st>>>(s->acc, ranges, meta);
  g_jn_launches += 3;
  JN_CUDA_CHECK(cudaGetLastError());
  return JN_OK;
}

extern "C" int jn_scan_from_disparity(jn_scan* s, const float* D, double ranges[JN_SCAN_BINS], jn_scan_meta* meta,
                                      uint8_t* dmap_u8) {
  if (!s || !D || !ranges || !meta) return JN_ERR_ARG;
  size_t n = (size_t)s->c.W * s->c.H;
  JN_CUDA_CHECK(cudaSetDevice(s->device));
  JN_CUDA_CHECK(cudaMemcpy(s->dD, D, n * sizeof(float), cudaMemcpyHostToDevice));
  int rc = jn_scan_from_disparity_batch(s, 1, s->dD, s->dRanges, s->dMeta, dmap_u8 ? s->dU8 : nullptr, 0);
  if (rc) return rc;
  JN_CUDA_CHECK(cudaMemcpy(ranges, s->dRanges, JN_SCAN_BINS * sizeof(double), cudaMemcpyDeviceToHost));
  JN_CUDA_CHECK(cudaMemcpy(meta, s->dMeta, sizeof(jn_scan_meta), cudaMemcpyDeviceToHost));
  if (dmap_u8) JN_CUDA_CHECK(cudaMemcpy(dmap_u8, s->dU8, n, cudaMemcpyDeviceToHost));
  return JN_OK;
}

extern "C" int jn_points_from_disparity(jn_scan* s, const float* D, double* points, int32_t* n_points,
                                        double ranges[JN_SCAN_BINS], jn_scan_meta* meta) {
  if (!s || !D || !points || !n_points || !ranges || !meta) return JN_ERR_ARG;
  size_t n = (size_t)s->c.W * s->c.H;
  JN_CUDA_CHECK(cudaSetDevice(s->device));
  int rc = ensure_acc(s, 1);
  if (rc) return rc;
  JN_CUDA_CHECK(cudaMemcpy(s->dD, D, n * sizeof(float), cudaMemcpyHostToDevice));
  points_count_kernel<<<(s->c.W + 127) / 128, 128>>>(s->c, s->dD, s->dCol);
  points_scan_kernel<<<1, 32>>>(s->dCol, s->c.W, s->dTotal);
  points_write_kernel<<<(s->c.W + 127) / 128, 128>>>(s->c, s->dD, s->dCol, s->dPts);
  acc_reset_kernel<<<1, 128>>>(s->acc, 1);
  scan_points_kernel<<<148, SCAN_THREADS>>>(s->c, s->dPts, s->dTotal, s->acc);
  scan_finalize_kernel<<<1, 96>>>(s->acc, s->dRanges, s->dMeta);
  g_jn_launches += 6;
  int total = 0;
  JN_CUDA_CHECK(cudaMemcpy(&total, s->dTotal, sizeof(int), cudaMemcpyDeviceToHost));
  *n_points = total;
  JN_CUDA_CHECK(cudaMemcpy(points, s->dPts, (size_t)total * 3 * sizeof(double), cudaMemcpyDeviceToHost));
  JN_CUDA_CHECK(cudaMemcpy(ranges, s->dRanges, JN_SCAN_BINS * sizeof(double), cudaMemcpyDeviceToHost));
  JN_CUDA_CHECK(cudaMemcpy(meta, s->dMeta, sizeof(jn_scan_meta), cudaMemcpyDeviceToHost));
  return JN_OK;
}

extern "C" int jn_pointcloud_from_disparity(jn_scan* s, const float* D, const uint8_t* image, int32_t image_stride,
                                            int32_t channels, float* xyz, float* rgb, int32_t* n_points,
                                            double ranges[JN_SCAN_BINS], jn_scan_meta* meta) {
  if (!s || !D || !image || !xyz || !rgb || !n_points || !ranges || !meta || (channels != 1 && channels != 3) ||
      image_stride < s->c.W * channels) {
    jn_set_error("jn_pointcloud_from_disparity: bad arguments");
    return JN_ERR_ARG;
  }
  const size_t n = (size_t)s->c.W * s->c.H, ibytes = (size_t)image_stride * s->c.H;
  JN_CUDA_CHECK(cudaSetDevice(s->device));
  int rc = ensure_acc(s, 1);
  if (rc) return rc;
  if (s->img_bytes < ibytes) {
    cudaFree(s->dImg);
    s->dImg = nullptr; s->img_bytes = 0;
    JN_CUDA_CHECK(cudaMalloc(&s->dImg, ibytes));
    s->img_bytes = ibytes;
  }
  if (!s->dXyz) JN_CUDA_CHECK(cudaMalloc(&s->dXyz, n * 4 * sizeof(float)));
  uint8_t* dImg = s->dImg;
  float* dXyz = s->dXyz;
  float* dRgb = dXyz + n * 3;
  cudaMemcpy(dImg, image, ibytes, cudaMemcpyHostToDevice);
  cudaMemcpy(s->dD, D, n * sizeof(float), cudaMemcpyHostToDevice);
  points_count_kernel<<<(s->c.W + 127) / 128, 128>>>(s->c, s->dD, s->dCol);
  points_scan_kernel<<<1, 32>>>(s->dCol, s->c.W, s->dTotal);
  points_write_kernel<<<(s->c.W + 127) / 128, 128>>>(s->c, s->dD, s->dCol, s->dPts);
  cloud_pack_kernel<<<(s->c.W + 127) / 128, 128>>>(s->c, s->dD, s->dCol, s->dPts, dImg, image_stride, channels, dXyz, dRgb);
  acc_reset_kernel<<<1, 128>>>(s->acc, 1);
  scan_points_kernel<<<148, SCAN_THREADS>>>(s->c, s->dPts, s->dTotal, s->acc);
  scan_finalize_kernel<<<1, 96>>>(s->acc, s->dRanges, s->dMeta);
  g_jn_launches += 7;
  int total = 0;
  cudaError_t e = cudaMemcpy(&total, s->dTotal, sizeof(int), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(xyz, dXyz, (size_t)total * 3 * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(rgb, dRgb, (size_t)total * sizeof(float), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(ranges, s->dRanges, JN_SCAN_BINS * sizeof(double), cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(meta, s->dMeta, sizeof(jn_scan_meta), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { jn_set_error("jn_pointcloud_from_disparity: %s", cudaGetErrorString(e)); return JN_ERR_CUDA; }
  *n_points = total;
  return JN_OK;
}

// -g path for n frames, everything in DEVICE memory, asynchronous on `stream`, no allocation after
// the first call of a given n: D n*W*H floats; image (optional) n frames of H rows, image_stride bytes
// each, 1 or 3 channels; xyz n*W*H*3 floats capacity (frame f starts at f*W*H*3); rgb (optional)
// n*W*H floats; counts n int32 (points of each frame); ranges n*90 doubles; meta n.
extern "C" int jn_pointcloud_batch(jn_scan* s, int n, const float* D, const uint8_t* image, int32_t image_stride,
                                   int32_t channels, float* xyz, float* rgb, int32_t* counts, double* ranges,
                                   jn_scan_meta* meta, void* stream) {
  if (!s || n <= 0 || !D || !xyz || !counts || !ranges || !meta || (image && channels != 1 && channels != 3) ||
      (image && image_stride < s->c.W * channels)) {
    jn_set_error("jn_pointcloud_batch: bad arguments");
    return JN_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  JN_CUDA_CHECK(cudaSetDevice(s->device));
  int rc = ensure_acc(s, n);
  if (rc) return rc;
  if (s->col_frames < n) {
    cudaFree(s->dColBatch);
    s->dColBatch = nullptr; s->col_frames = 0;
    JN_CUDA_CHECK(cudaMalloc(&s->dColBatch, (size_t)n * s->c.W * sizeof(int)));
    s->col_frames = n;
  }
  const int W = s->c.W;
  pc_count_kernel<<<dim3((W + 127) / 128, n), 128, 0, st>>>(s->c, D, s->dColBatch);
  pc_scan_kernel<<<n, 256, 0, st>>>(s->dColBatch, W, counts);
  acc_reset_kernel<<<(n * ACC_WORDS + 255) / 256, 256, 0, st>>>(s->acc, n);
  pc_write_kernel<<<dim3((W + SCAN_THREADS - 1) / SCAN_THREADS, n), SCAN_THREADS, 0, st>>>(
      s->c, D, s->dColBatch, image, image_stride, channels, (size_t)image_stride * s->c.H, xyz, rgb, s->acc);
  scan_finalize_kernel<<<n, 96, 0, st>>>(s->acc, ranges, meta);
  g_jn_launches += 5;
  JN_CUDA_CHECK(cudaGetLastError());
  return JN_OK;
}

extern "C" int jn_scan_compact(const double ranges[JN_SCAN_BINS], float* out) {
  int n = 0;
  for (int i = JN_SCAN_BINS - 1; i >= 0; i--)
    if (ranges[i] < SCAN_INF - 1) out[n++] = (float)ranges[i];
  return n;
}
