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
  int fast_ok;            // q_simple and the table division below verified against IEEE division at create time
  const double* tab;      // device, 4 x 256 doubles indexed by the u8 disparity d: b = Q32*d + 0, RN(1/b),
                          // XR[2]*(Q23/b), XR[5]*(Q23/b)
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

// per-frame accumulator: 90 range keys, then angle min/max, range min/max keys, n_points, and one word
// holding the running FLOAT extrema of the angle (fkey(min) in the low half, fkey(max) in the high half):
// CTAs of the fast scan kernel start from what earlier CTAs of the frame have seen, so only points near the
// frame's extrema are ever evaluated in double
constexpr int ACC_WORDS = JN_SCAN_BINS + 6;
constexpr int ACC_FEXT = JN_SCAN_BINS + 5;

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
// The running (bin, key) pairs of a warp: lanes are neighbouring columns, so they sit in the same one or
// two bins.  A 64-bit shared-memory atomicMin is a compare-and-swap loop, and 32 lanes hitting one
// address serialise it; instead the lanes of each distinct bin reduce their keys with two REDUX.MIN (high
// word, then low word among the lanes that hold the smallest high word) and one lane issues the atomic.
// Called by all 32 lanes (curbin < 0: nothing pending).
__device__ __forceinline__ void warp_bin_flush(BlockAcc& a, int bin, unsigned long long key) {
  const int lane = threadIdx.x & 31;
  unsigned pending = __ballot_sync(0xffffffffu, bin >= 0);
  while (pending) {
    const int leader = __ffs(pending) - 1;
    const int b = __shfl_sync(0xffffffffu, bin, leader);
    const bool mine = bin == b;
    const unsigned hi = mine ? (unsigned)(key >> 32) : 0xffffffffu;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned lo = (mine && hi == mhi) ? (unsigned)key : 0xffffffffu;
    const unsigned mlo = __reduce_min_sync(0xffffffffu, lo);
    if (lane == leader) atomicMin(&a.bins[b], ((unsigned long long)mhi << 32) | mlo);
    pending &= ~__ballot_sync(0xffffffffu, mine);
  }
}

// warp-reduce the extrema, one shared atomic per warp and value
__device__ __forceinline__ void lacc_finish(LocalAcc& l, BlockAcc& a) {
  warp_bin_flush(a, l.curbin, l.curkey);
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
  if (k == ACC_FEXT) v = 0x807fffff7f800000ull;     // fkey(-inf) : fkey(+inf)
  acc[i] = v;
}

__device__ __forceinline__ int to_u8(float x) { return min(max(__float2int_rn(x), 0), 255); }

constexpr int SCAN_THREADS = 256;

// ---- fast path of scan_kernel (stereoRectify's sparse Q) -------------------------------------------
// The u8 disparity takes 256 values, so the divisor W = Q32*d of the reprojection, its correctly rounded
// reciprocal and the d-only terms XR[.]*Z come from a 256-entry table.  With r = RN(1/b) the quotient
// a/b is q0 = RN(a*r), rem = a - q0*b (exact in one FMA), q = RN(q0 + rem*r): three operations instead
// of the ~25 of the IEEE division sequence, and correctly rounded (Markstein).  The theorem's
// precondition (q0 faithful) is not taken on trust: scan_verify_kernel compares the sequence with the
// real division for EVERY numerator the image can produce (W + H of them) x 255 divisors when the
// object is created and clears fast_ok on any difference.
//
// atan2 is only evaluated in double where its value can matter: the bin index comes from a float
// atan2f unless the angle lies within BIN_MARGIN degrees of a bin border (float error bound 7e-5
// degrees: 2 ulp of atan2f, input roundings, the degree conversion), and angle_min / angle_max only
// need the exact angle of the points whose float angle is within ANG_MARGIN (>= 2 x 6e-7 rad error) of
// the running float extremum -- the one clearly extreme point is kept as a deferred candidate and
// evaluated once per thread at the end.  Every skipped evaluation provably cannot change a bin index or
// an extremum, so the result is the all-double result bit for bit.
constexpr float BIN_MARGIN = 1e-3f, ANG_MARGIN = 4e-6f;

__device__ __forceinline__ double table_div(double a, double b, double r) {
  const double q0 = a * r;
  const double rem = fma(-q0, b, a);
  return fma(rem, r, q0);
}

__global__ void scan_tab_kernel(ScanConst c, double* __restrict__ tab) {
  const int d = threadIdx.x;
  const double b = c.Q[14] * (double)d + 0.0;
  const double Z = c.Q[11] / b;
  tab[d] = b;
  tab[256 + d] = 1.0 / b;
  tab[512 + d] = c.XR[2] * Z;
  tab[768 + d] = c.XR[5] * Z;
}

// ok[0] is cleared if the table division differs from the IEEE division for any numerator / divisor pair
__global__ void scan_verify_kernel(ScanConst c, const double* __restrict__ tab, int* __restrict__ ok) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  bool good = true;
  for (int axis = 0; axis < 2; axis++) {
    if (t >= (axis ? c.H : c.W)) continue;
    const double a = (double)(t + (axis ? c.oy : c.ox)) + (axis ? c.Q[7] : c.Q[3]);
    for (int d = 1; d < 256; d++) {
      const double b = tab[d], r = tab[256 + d];
      const double q = table_div(a, b, r), e = a / b;
      good = good && (__double_as_longlong(q) == __double_as_longlong(e)) && (b == b) && (b != 0.0) &&
             (fabs(r) < 1e300) && (fabs(e) < 1e30);
    }
  }
  if (!good) atomicExch(ok, 0);
}

// order-preserving float -> int key (for shared-memory atomicMin / atomicMax on floats)
__device__ __forceinline__ int fkey(float x) {
  const int b = __float_as_int(x);
  return b >= 0 ? b : (b ^ 0x7fffffff);
}

struct FastAcc {
  float fmin, fmax;        // running float extrema of the angle
  float camin, camax;      // float angle of the deferred candidates (+/-inf: none)
  double minX, minY, maxX, maxY;
};
__device__ __forceinline__ float fkey_inv(int k) { return __int_as_float(k >= 0 ? k : (k ^ 0x7fffffff)); }
// the running extrema start from the frame's (any float angle of a real point is a valid bound); no candidates yet
__device__ __forceinline__ void facc_init(FastAcc& f, int kmin, int kmax) {
  f.fmin = fkey_inv(kmin); f.fmax = fkey_inv(kmax);
  f.camin = __int_as_float(0x7f800000); f.camax = __int_as_float(0xff800000);
  f.minX = f.minY = f.maxX = f.maxY = 0.0;
}

__device__ __forceinline__ void lacc_bin(LocalAcc& l, BlockAcc& a, int k, unsigned long long kr) {
  if (k != l.curbin) {
    if (l.curbin >= 0) atomicMin(&a.bins[l.curbin], l.curkey);
    l.curbin = k;
    l.curkey = kr;
  } else {
    l.curkey = min(l.curkey, kr);
  }
}

// one point with finite coordinates (d >= 1 through the verified table)
__device__ __forceinline__ void lacc_add_fast(LocalAcc& l, FastAcc& f, BlockAcc& a, double X, double Y) {
  const double r2 = Y * Y + X * X;                       // >= +0, never NaN here
  const unsigned long long kr = (unsigned long long)__double_as_longlong(r2) | 0x8000000000000000ull;   // = okey(r2)
  l.rmin = min(l.rmin, kr);
  l.rmax = max(l.rmax, kr);
  l.npts++;
  const float Xf = (float)X, Yf = (float)Y;
  const float mag = fmaxf(fabsf(Xf), fabsf(Yf));
  const float af = atan2f(Yf, Xf);
  const float t = 45.0f - af * 57.29746936f;             // 180 / 3.1415
  const float fl = floorf(t), fr = t - fl;
  bool exact = !(mag > 1e-30f && mag < 1e30f);           // float image of the point degenerate: no float reasoning
  if (!exact) {
    if (fr < BIN_MARGIN || fr > 1.0f - BIN_MARGIN) exact = true;
    if (af < f.fmin - ANG_MARGIN) { f.fmin = af; f.camin = af; f.minX = X; f.minY = Y; }
    else if (af <= f.fmin + ANG_MARGIN) { exact = true; f.fmin = fminf(f.fmin, af); }
    if (af > f.fmax + ANG_MARGIN) { f.fmax = af; f.camax = af; f.maxX = X; f.maxY = Y; }
    else if (af >= f.fmax - ANG_MARGIN) { exact = true; f.fmax = fmaxf(f.fmax, af); }
  }
  int k;
  if (exact) {
    const double th = atan2(Y, X);
    const double deg = th * 180. / 3.1415;
    const unsigned long long kt = okey(th);
    l.amin = min(l.amin, kt);
    l.amax = max(l.amax, kt);
    const double kf = floor((double)JN_SCAN_BINS * (90. / 2. - deg) / 90.);
    k = (kf >= 0 && kf < (double)JN_SCAN_BINS) ? (int)kf : -1;
  } else {
    k = (fl >= 0.0f && fl < (float)JN_SCAN_BINS) ? (int)fl : -1;
  }
  if (k >= 0) lacc_bin(l, a, k, kr);
}

// the deferred candidates: only those within ANG_MARGIN of the CTA's float extremum can be the extremum
__device__ __forceinline__ void facc_finish(LocalAcc& l, FastAcc& f, int* s_fext, unsigned long long* g_fext) {
  float wmin = f.fmin, wmax = f.fmax;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    wmin = fminf(wmin, __shfl_xor_sync(0xffffffffu, wmin, off));
    wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, off));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&s_fext[0], fkey(wmin));
    atomicMax(&s_fext[1], fkey(wmax));
  }
  __syncthreads();
  const int kmin = s_fext[0], kmax = s_fext[1];
  if (threadIdx.x == 0) {
    atomicMin(reinterpret_cast<int*>(g_fext), kmin);
    atomicMax(reinterpret_cast<int*>(g_fext) + 1, kmax);
  }
  const float bmin = fkey_inv(kmin), bmax = fkey_inv(kmax);
  if (f.camin <= bmin + ANG_MARGIN && f.camin < __int_as_float(0x7f800000)) {   // +inf: no candidate
    const unsigned long long kt = okey(atan2(f.minY, f.minX));
    l.amin = min(l.amin, kt);
    l.amax = max(l.amax, kt);
  }
  if (f.camax >= bmax - ANG_MARGIN && f.camax > __int_as_float(0xff800000)) {
    const unsigned long long kt = okey(atan2(f.maxY, f.maxX));
    l.amin = min(l.amin, kt);
    l.amax = max(l.amax, kt);
  }
}


// grid (ceil(W/256), ceil(H/SCAN_ROWS), frames).  A thread walks SCAN_ROWS rows of one column; all of its
// disparities and gate pairs are loaded up front (SCAN_ROWS independent loads in flight per thread: the
// row loop then never waits for memory -- with a load at the top of every iteration the kernel was bound
// by memory latency, not by its arithmetic).  FAST: see above (c.fast_ok); a pixel with d = 0 (only
// reachable through a wrapped gate, H8) divides by zero and takes the general expressions.
constexpr int SCAN_ROWS = 32;

template <bool FAST>
__global__ void __launch_bounds__(SCAN_THREADS)
scan_kernel(ScanConst c, const float* __restrict__ D, const uint8_t* __restrict__ gate,
            unsigned long long* __restrict__ acc, uint8_t* __restrict__ dmap_u8) {
  __shared__ BlockAcc s;
  __shared__ double s_tab[FAST ? 1024 : 1];
  __shared__ int s_fext[2];
  const int tid = threadIdx.x, frame = blockIdx.z;
  const int i = blockIdx.x * SCAN_THREADS + tid;
  const size_t fp = (size_t)frame * c.W * c.H;
  const int j0 = blockIdx.y * SCAN_ROWS;
  __shared__ uint8_t s_dv[SCAN_ROWS][SCAN_THREADS];   // u8 disparities (convertTo(CV_8U)) of the tile
  __shared__ unsigned short s_gv[SCAN_ROWS][SCAN_THREADS];
  const int rows = min(SCAN_ROWS, c.H - j0);
  unsigned long long* facc = acc + (size_t)frame * ACC_WORDS;
  if (i < c.W) {
    const float* dp = D + fp + (size_t)j0 * c.W + i;
    const unsigned short* gp = reinterpret_cast<const unsigned short*>(gate) + (size_t)j0 * c.W + i;
    float dv[SCAN_ROWS];
    unsigned short gv[SCAN_ROWS];
#pragma unroll
    for (int r = 0; r < SCAN_ROWS; r++) {
      dv[r] = 0.f;
      gv[r] = 0;
      if (r < rows) { dv[r] = dp[r * c.W]; gv[r] = gp[r * c.W]; }
    }
#pragma unroll
    for (int r = 0; r < SCAN_ROWS; r++) { s_dv[r][tid] = (uint8_t)to_u8(dv[r]); s_gv[r][tid] = gv[r]; }
  }
  acc_init(s, tid, SCAN_THREADS);
  if (FAST) {
    for (int k = tid; k < 1024; k += SCAN_THREADS) s_tab[k] = c.tab[k];
    if (tid == 0) {
      const unsigned long long w = __ldcg(facc + ACC_FEXT);
      s_fext[0] = (int)(unsigned)w;
      s_fext[1] = (int)(unsigned)(w >> 32);
    }
  }
  __syncthreads();
  LocalAcc l;
  lacc_init(l);
  FastAcc f;
  facc_init(f, FAST ? s_fext[0] : 0x7f800000, FAST ? s_fext[1] : (int)0x807fffff);
  if (i < c.W) {
    const double px = (double)(i + c.ox) + c.Q[3];
    uint8_t* up = dmap_u8 ? dmap_u8 + fp + (size_t)j0 * c.W + i : nullptr;
#pragma unroll 1
    for (int r = 0; r < rows; r++) {
      const int j = j0 + r;
      const int d = s_dv[r][tid];
      if (up) up[r * c.W] = (uint8_t)d;
      const unsigned gq = s_gv[r][tid];
      const int g0 = gq & 0xff, g1 = gq >> 8;
      if (d < g0 || d > g1) continue;
      if (FAST && d != 0) {
        const double py = (double)(j + c.oy) + c.Q[7];
        const double b = s_tab[d], rb = s_tab[256 + d];
        const double X = table_div(px, b, rb), Y = table_div(py, b, rb);
        const double o0 = ((c.XR[0] * X + c.XR[1] * Y) + s_tab[512 + d]) + c.XT[0];
        const double o1 = ((c.XR[3] * X + c.XR[4] * Y) + s_tab[768 + d]) + c.XT[1];
        lacc_add_fast(l, f, s, o0, o1);
      } else {
        double rr[3];
        reproject(c, (double)(i + c.ox), (double)(j + c.oy), (double)d, rr);
        lacc_add(l, s, rr[0], rr[1]);
      }
    }
  }
  if (FAST) facc_finish(l, f, s_fext, facc + ACC_FEXT);
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
  double* dTab;                     // 4 x 256 table of the fast scan path (c.tab)
};

static int ensure_acc(jn_scan* s, int n) {
  if (n <= s->cap_frames) return JN_OK;
  if (s->acc) cudaFree(s->acc);
  s->acc = nullptr;
  JN_CUDA_CHECK(cudaMalloc(&s->acc, (size_t)n * ACC_WORDS * sizeof(unsigned long long)));
  s->cap_frames = n;
  return JN_OK;
}

static int g_scan_fast_override = -1;   // test hook (jn_debug_scan_fast): -1 default, 0 general expressions only
extern "C" void jn_debug_scan_fast(int enable) { g_scan_fast_override = enable; }
extern "C" int jn_scan_fast_path(const jn_scan* s) { return s ? s->c.fast_ok : 0; }

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
  // fast path of scan_kernel: reciprocal table, then the table division checked against the IEEE division
  // for every numerator of this image size (JN_SCAN_FAST=0 keeps the general expressions)
  s->c.fast_ok = 0;
  s->c.tab = nullptr;
  const char* fast_env = getenv("JN_SCAN_FAST");
  const bool fast_off = g_scan_fast_override >= 0 ? g_scan_fast_override == 0 : (fast_env && fast_env[0] == '0');
  if (s->c.q_simple && !fast_off) {
    int* dOk = s->dTotal;
    const int one = 1;
    int ok = 0;
    if (cudaMalloc(&s->dTab, 1024 * sizeof(double)) == cudaSuccess &&
        cudaMemcpy(dOk, &one, sizeof(int), cudaMemcpyHostToDevice) == cudaSuccess) {
      s->c.tab = s->dTab;
      scan_tab_kernel<<<1, 256>>>(s->c, s->dTab);
      const int m = width > height ? width : height;
      scan_verify_kernel<<<(m + 127) / 128, 128>>>(s->c, s->dTab, dOk);
      g_jn_launches += 2;
      if (cudaMemcpy(&ok, dOk, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess) s->c.fast_ok = ok;
    }
  }
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
  cudaFree(s->dImg); cudaFree(s->dXyz); cudaFree(s->dColBatch); cudaFree(s->dTab);
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

// n frames on `stream`, accumulators acc_frame0 .. acc_frame0 + n - 1 (two calls on different streams must
// use disjoint accumulator ranges; jn_scan_reserve sizes the array beforehand)
int jn_scan_batch_at(jn_scan* s, int acc_frame0, int n, const float* D, double* ranges, jn_scan_meta* meta,
                     uint8_t* dmap_u8, cudaStream_t st) {
  if (acc_frame0 + n > s->cap_frames) { jn_set_error("jn_scan_batch_at: accumulators not reserved"); return JN_ERR_ARG; }
  unsigned long long* acc = s->acc + (size_t)acc_frame0 * ACC_WORDS;
  acc_reset_kernel<<<(n * ACC_WORDS + 255) / 256, 256, 0, st>>>(acc, n);
  dim3 grid((s->c.W + SCAN_THREADS - 1) / SCAN_THREADS, (s->c.H + SCAN_ROWS - 1) / SCAN_ROWS, n);
  if (s->c.fast_ok)
    scan_kernel<true><<<grid, SCAN_THREADS, 0, st>>>(s->c, D, s->gate, acc, dmap_u8);
  else
    scan_kernel<false><<<grid, SCAN_THREADS, 0, st>>>(s->c, D, s->gate, acc, dmap_u8);
  scan_finalize_kernel<<<n, 96, 0, st>>>(acc, ranges, meta);
  g_jn_launches += 3;
  JN_CUDA_CHECK(cudaGetLastError());
  return JN_OK;
}
int jn_scan_reserve(jn_scan* s, int n) {
  JN_CUDA_CHECK(cudaSetDevice(s->device));
  if (n > s->cap_frames) JN_CUDA_CHECK(cudaDeviceSynchronize());   // the array is replaced: nothing may still use it
  return ensure_acc(s, n);
}

extern "C" int jn_scan_from_disparity_batch(jn_scan* s, int n, const float* D, double* ranges, jn_scan_meta* meta,
                                            uint8_t* dmap_u8, void* stream) {
  if (!s || n <= 0 || !D || !ranges || !meta) return JN_ERR_ARG;
  JN_CUDA_CHECK(cudaSetDevice(s->device));
  int rc = ensure_acc(s, n);
  if (rc) return rc;
  return jn_scan_batch_at(s, 0, n, D, ranges, meta, dmap_u8, (cudaStream_t)stream);
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
  if (!ranges || !out) return JN_ERR_ARG;
  int n = 0;
  for (int i = JN_SCAN_BINS - 1; i >= 0; i--)
    if (ranges[i] < SCAN_INF - 1) out[n++] = (float)ranges[i];
  return n;
}
