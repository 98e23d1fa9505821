// post.cu -- post-processing of the raw disparity maps (elas.cpp:909-1560):
// leftRightConsistencyCheck, removeSmallSegments, gapInterpolation,
// adaptiveMean, median.  The maps are g.Wd x g.Hd: full resolution, or half
// resolution with subsampling (then the L/R check halves the disparity, the
// segment-size and gap limits are rescaled and the adaptive mean has 4 taps).
//
// All five are HBM-bound streaming passes over H*W floats.
//   L/R check        one thread per pixel, out of place (the reference copies both maps first).
//   small segments   4-connected components, tiled: horizontal runs and the union-find of a
//                    256 x 16 tile in shared memory, tile components linked across tile borders
//                    by a global union-find, sizes summed at the roots.  At this stage every
//                    invalid pixel is exactly -10 and similarity is symmetric, so the
//                    components equal the reference's breadth-first segments.
//   gap interpolation rows: one CTA per row (previous/next valid index by scans, any
//                    gap width); columns: one thread per column walking down, lanes =
//                    adjacent columns so every step is a coalesced row access.
//   adaptive mean    the reference's 8-tap filter including its bit-mask "abs"
//                     and its exact summation order, so results are
//                    bit-identical; horizontal then vertical.
//   median           7-tap horizontal then vertical selection.
#include "common.cuh"
#include "blockutil.cuh"

namespace {

// ------------------------------------------------------------------ L/R check
__global__ void lr_kernel(Geo g, Workspace ws) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t fp = (size_t)frame * W * H, a = fp + (size_t)v * W + u;
  const float* D1 = ws.Draw[0];
  const float* D2 = ws.Draw[1];
  const float thr = (float)g.p.lr_threshold;
  float d1 = D1[a], d2 = D2[a];
  float o1 = -10.f, o2 = -10.f;
  // elas.cpp:937-943: at half resolution a disparity of d moves d/2 map pixels
  const bool sub = g.p.subsampling != 0;
  float uw1 = sub ? (float)u - d1 / 2 : (float)u - d1, uw2 = sub ? (float)u + d2 / 2 : (float)u + d2;
  if (d1 >= 0 && uw1 >= 0 && uw1 < (float)W) {
    o1 = (fabsf(D2[fp + (size_t)v * W + (int)uw1] - d1) > thr) ? -10.f : d1;
  }
  if (d2 >= 0 && uw2 >= 0 && uw2 < (float)W) {
    o2 = (fabsf(D1[fp + (size_t)v * W + (int)uw2] - d2) > thr) ? -10.f : d2;
  }
  ws.Dlr[0][a] = o1;
  ws.Dlr[1][a] = o2;
}

// Four pixels per thread (map width a multiple of 4): 128-bit loads and stores, four times the
// memory-level parallelism per thread.  Same arithmetic as lr_kernel.
constexpr int Q_THREADS = 128;   // threads per CTA of the quad kernels; a CTA covers 512 pixels of a row

// RIGHT = false: the right map is neither post-processed nor returned (ROBOTICS preset with D2 = NULL),
// so only the left result is computed and written.
template <bool RIGHT>
__global__ void __launch_bounds__(Q_THREADS) lr4_kernel(Geo g, Workspace ws) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int q = blockIdx.x * Q_THREADS + threadIdx.x, v = blockIdx.y;
  if (4 * q >= W) return;
  const size_t rb = (size_t)frame * W * H + (size_t)v * W;
  const float* __restrict__ D1 = ws.Draw[0] + rb;
  const float* __restrict__ D2 = ws.Draw[1] + rb;
  const float thr = (float)g.p.lr_threshold;
  const bool sub = g.p.subsampling != 0;
  const float4 a = reinterpret_cast<const float4*>(D1)[q], b = reinterpret_cast<const float4*>(D2)[q];
  const float d1[4] = {a.x, a.y, a.z, a.w}, d2[4] = {b.x, b.y, b.z, b.w};
  float o1[4], o2[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const int u = 4 * q + j;
    const float uw1 = sub ? (float)u - d1[j] / 2 : (float)u - d1[j], uw2 = sub ? (float)u + d2[j] / 2 : (float)u + d2[j];
    o1[j] = -10.f;
    o2[j] = -10.f;
    if (d1[j] >= 0 && uw1 >= 0 && uw1 < (float)W) o1[j] = (fabsf(D2[(int)uw1] - d1[j]) > thr) ? -10.f : d1[j];
    if (RIGHT && d2[j] >= 0 && uw2 >= 0 && uw2 < (float)W)
      o2[j] = (fabsf(D1[(int)uw2] - d2[j]) > thr) ? -10.f : d2[j];
  }
  reinterpret_cast<float4*>(ws.Dlr[0] + rb)[q] = make_float4(o1[0], o1[1], o1[2], o1[3]);
  if (RIGHT) reinterpret_cast<float4*>(ws.Dlr[1] + rb)[q] = make_float4(o2[0], o2[1], o2[2], o2[3]);
}

// ------------------------------------------------------------ small segments
constexpr int ROW_THREADS = 256;

__device__ __forceinline__ int uf_find(int* label, int x) {
  int p = __ldcg(label + x);   // L2 reads: other SMs update labels with atomics
  while (p != x) {
    int gp = __ldcg(label + p);
    if (gp != p) label[x] = gp;  // path halving; labels only ever decrease toward an ancestor
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ void uf_union(int* label, int a, int b) {
  while (true) {
    a = uf_find(label, a);
    b = uf_find(label, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }
    int old = atomicMin(&label[a], b);
    if (old == a) return;
    a = old;
  }
}

// label[i] is the root of i's tile component, and after seg_roots_kernel every tile root points straight
// at its global root: two hops at most.
__global__ void seg_apply_kernel(Geo g, Workspace ws, int side) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t fp = (size_t)frame * W * H, a = fp + (size_t)v * W + u;
  int l = ws.label[a];
  if (l < 0) return;
  const int root = ws.label[fp + l];
  if (ws.segsize[fp + root] < g.speckle_eff) ws.Dlr[side][a] = -10.f;
}

// Quad version of seg_apply_kernel (map width a multiple of 4).
__global__ void __launch_bounds__(Q_THREADS) seg_apply4_kernel(Geo g, Workspace ws, int side) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int q = blockIdx.x * Q_THREADS + threadIdx.x, v = blockIdx.y;
  if (4 * q >= W) return;
  const size_t fp = (size_t)frame * W * H;
  const size_t a = fp + (size_t)v * W + 4 * q;
  const int4 lq = *reinterpret_cast<const int4*>(ws.label + a);
  if (lq.x < 0 && lq.y < 0 && lq.z < 0 && lq.w < 0) return;
  const int l[4] = {lq.x, lq.y, lq.z, lq.w};
  bool kill[4], any = false, pk = false;
  int pl = -1;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    if (l[j] < 0) { kill[j] = false; continue; }
    if (l[j] != pl) {       // pixels of one run share the label: look the component up once
      pl = l[j];
      pk = ws.segsize[fp + ws.label[fp + pl]] < g.speckle_eff;
    }
    kill[j] = pk;
    any = any || pk;
  }
  if (!any) return;
  float* D = ws.Dlr[side] + a;
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (kill[j]) D[j] = -10.f;
}

// ---- tiled components: the union-find of a 256 x 16 tile runs in shared memory ----------------------
// seg_tile_kernel labels the 4-connected components of ONE tile: horizontal runs from one ballot per 32 pixels (a run
// crossing a chunk border is carried in warp-uniform registers), vertical links by a union-find on the run starts with shared-memory atomics (tens
// of cycles per hop instead of an L2 round trip), then every pixel's label is its tile component's
// root (the smallest pixel index of the component, as a global index) and segsize is the tile
// component's size at the root and 0 everywhere else.  seg_border_kernel then links tile components
// across tile borders (1/16 of the row pairs, 1/256 of the column pairs) with the global union-find,
// seg_roots_kernel adds the size of every tile root that is not a global root to its global root, and
// seg_apply*_kernel reads label -> tile root -> global root -> size.
// A vertical link is skipped when the left neighbours make the same link (both rows continue their runs
// to the left and the left pixels are vertically similar too: the link is implied by those three); the
// links that argument relies on are in-tile run links or vertical links further left in the same tile,
// never links that could be skipped for the mirrored reason: at a tile's left column and on vertical
// tile borders every link is made.
constexpr int CT_W = 256, CT_H = 16, CT_THREADS = 256;

// no path compression: the trees of a 16-row tile are shallow, and the only writes of the union phase stay
// the atomicMin of sm_union (labels only ever decrease toward an ancestor)
__device__ __forceinline__ int sm_find(volatile int* lab, int x) {
  int p = lab[x];
  while (p != x) {
    x = p;
    p = lab[x];
  }
  return x;
}
__device__ __forceinline__ void sm_union(int* lab, int a, int b) {
  while (true) {
    a = sm_find(lab, a);
    b = sm_find(lab, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&lab[a], b);
    if (old == a) return;
    a = old;
  }
}

__global__ void __launch_bounds__(CT_THREADS) seg_tile_kernel(Geo g, Workspace ws, int side) {
  __shared__ float sD[CT_H * CT_W];   // the tile; reused for the component sizes
  __shared__ int lab[CT_H * CT_W];
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * CT_W, y0 = blockIdx.y * CT_H;
  const size_t fp = (size_t)frame * W * H;
  const float* __restrict__ D = ws.Dlr[side] + fp;
  const float thr = g.p.speckle_sim_threshold;
  const unsigned le_mask = 0xFFFFFFFFu >> (31 - lane);   // lanes 0..lane
  // rows of the tile: a warp per row, 32 pixels per step; pixels outside the map are invalid.  All loads
  // of a warp's rows are issued before the (serial, shuffle-linked) run detection starts.
  constexpr int RPW = CT_H / (CT_THREADS / 32), CPR = CT_W / 32;   // rows per warp, chunks per row
  float dv[RPW][CPR];
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int y = y0 + warp + i * (CT_THREADS / 32);
#pragma unroll
    for (int k = 0; k < CPR; k++) {
      const int x = x0 + 32 * k + lane;
      dv[i][k] = (y < H && x < W) ? D[(size_t)y * W + x] : -10.f;
    }
  }
#pragma unroll
  for (int i = 0; i < RPW; i++) {
    const int r = warp + i * (CT_THREADS / 32);
    int cur = -1;
    float prev = -10.f;
#pragma unroll
    for (int k = 0; k < CPR; k++) {
      const int u0 = 32 * k, c = u0 + lane;
      const float d = dv[i][k];
      sD[r * CT_W + c] = d;
      float dl = __shfl_up_sync(0xffffffffu, d, 1);
      if (lane == 0) dl = prev;
      prev = __shfl_sync(0xffffffffu, d, 31);
      const bool valid = d >= 0;
      const bool start = valid && (c == 0 || !(dl >= 0) || fabsf(d - dl) > thr);
      const unsigned below = __ballot_sync(0xffffffffu, start) & le_mask;
      // a valid pixel without a start at or before it in this chunk continues the carried run
      const int st = valid ? (below ? u0 + 31 - __clz(below) : cur) : -1;
      lab[r * CT_W + c] = valid ? r * CT_W + st : -1;
      cur = __shfl_sync(0xffffffffu, st, 31);
    }
  }
  __syncthreads();
  // vertical links inside the tile (thread = column, rows top to bottom)
  for (int r = 0; r < CT_H - 1; r++) {
    const int idx = r * CT_W + tid;
    const float d = sD[idx], e = sD[idx + CT_W];
    if (d >= 0 && e >= 0 && fabsf(d - e) <= thr) {
      bool dup = false;
      if (tid > 0) {
        const float dl = sD[idx - 1], el = sD[idx + CT_W - 1];
        dup = dl >= 0 && el >= 0 && fabsf(d - dl) <= thr && fabsf(e - el) <= thr && fabsf(dl - el) <= thr;
      }
      if (!dup) sm_union(lab, lab[idx], lab[idx + CT_W]);
    }
  }
  __syncthreads();
  int* ssz = reinterpret_cast<int*>(sD);
  for (int i = tid; i < CT_H * CT_W; i += CT_THREADS) ssz[i] = 0;
  __syncthreads();
  // roots and sizes: one shared atomic per (run, 32-pixel chunk)
  int root[CT_H];
#pragma unroll
  for (int r = 0; r < CT_H; r++) {
    const int idx = r * CT_W + tid;
    const int l = lab[idx];
    int x = l;
    if (l >= 0) {
      int p = lab[x];
      while (p != x) { x = p; p = lab[x]; }
    }
    root[r] = x;
    const int lp = __shfl_up_sync(0xffffffffu, l, 1);
    const bool head = l >= 0 && (lane == 0 || l != lp);
    const unsigned hm = __ballot_sync(0xffffffffu, head), vm = __ballot_sync(0xffffffffu, l >= 0);
    if (head) {
      const unsigned end = (hm | ~vm) & ~le_mask;
      atomicAdd(&ssz[x], end ? __ffs(end) - 1 - lane : 32 - lane);
    }
  }
  __syncthreads();
  int* __restrict__ label = ws.label + fp;
  int* __restrict__ segsize = ws.segsize + fp;
  const int x = x0 + tid;
  if (x < W) {
#pragma unroll
    for (int r = 0; r < CT_H; r++) {
      const int y = y0 + r;
      if (y >= H) break;
      const int rt = root[r];
      const size_t a = (size_t)y * W + x;
      label[a] = rt >= 0 ? (y0 + (rt >> 8)) * W + x0 + (rt & (CT_W - 1)) : -1;
      segsize[a] = (rt == r * CT_W + tid) ? ssz[rt] : 0;
    }
  }
}

// links across tile borders: first the pixel pairs (v, v+1) with v + 1 a multiple of CT_H, then the
// pairs (u, u+1) with u + 1 a multiple of CT_W
__global__ void __launch_bounds__(256) seg_border_kernel(Geo g, Workspace ws, int side) {
  const int frame = blockIdx.y;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const size_t fp = (size_t)frame * W * H;
  const float* __restrict__ D = ws.Dlr[side] + fp;
  int* label = ws.label + fp;
  const float thr = g.p.speckle_sim_threshold;
  const int nhb = (H - 1) / CT_H, nvb = (W - 1) / CT_W;
  const int total = nhb * W + nvb * H;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    if (t < nhb * W) {
      const int k = t / W, u = t - k * W, v = CT_H * (k + 1) - 1;
      const int i = v * W + u;
      const float d = D[i], e = D[i + W];
      if (!(d >= 0 && e >= 0 && fabsf(d - e) <= thr)) continue;
      if (u % CT_W != 0) {
        const float dl = D[i - 1], el = D[i + W - 1];
        if (dl >= 0 && el >= 0 && fabsf(d - dl) <= thr && fabsf(e - el) <= thr && fabsf(dl - el) <= thr) continue;
      }
      uf_union(label, i, i + W);
    } else {
      const int t2 = t - nhb * W;
      const int k = t2 / H, v = t2 - k * H, u = CT_W * (k + 1) - 1;
      const int i = v * W + u;
      const float d = D[i], e = D[i + 1];
      if (d >= 0 && e >= 0 && fabsf(d - e) <= thr) uf_union(label, i, i + 1);
    }
  }
}

// every tile root (segsize > 0) that is not a global root adds its size to its global root and is left
// pointing straight at it
template <bool QUAD>
__global__ void __launch_bounds__(256) seg_roots_kernel(Geo g, Workspace ws) {
  const int frame = blockIdx.y;
  if (ws.info[frame].status != JN_OK) return;
  const int n = g.Wd * g.Hd;
  const size_t fp = (size_t)frame * n;
  int* label = ws.label + fp;
  int* segsize = ws.segsize + fp;
  constexpr int PER = QUAD ? 4 : 1;
  const int i0 = (blockIdx.x * 256 + threadIdx.x) * PER;
  if (i0 >= n) return;
  int sz[PER];
  if constexpr (QUAD) {
    const int4 q = *reinterpret_cast<const int4*>(segsize + i0);
    sz[0] = q.x; sz[1] = q.y; sz[2] = q.z; sz[3] = q.w;
  } else {
    sz[0] = segsize[i0];
  }
#pragma unroll
  for (int j = 0; j < PER; j++) {
    if (sz[j] <= 0) continue;
    const int i = i0 + j;
    // a component's root is its smallest pixel index and labels only decrease towards it, so
    // compressing with atomicMin can never replace a root by a larger ancestor
    int x = i, p = __ldcg(label + x);
    while (p != x) {
      const int gp = __ldcg(label + p);
      if (gp != p) atomicMin(label + x, gp);
      x = p;
      p = gp;
    }
    if (x == i) continue;               // a global root keeps its own size and collects the others'
    atomicMin(label + i, x);
    atomicAdd(segsize + x, sz[j]);      // i is not a global root, so nobody adds to segsize[i]: sz[j] is final
  }
}

// ------------------------------------------------------------ gap interpolation
__device__ __forceinline__ float ipol(float d1, float d2) {
  return (fabsf(d1 - d2) < 3.0f) ? (d1 + d2) / 2 : fminf(d1, d2);
}

__global__ void __launch_bounds__(ROW_THREADS) gap_rows_kernel(Geo g, Workspace ws, int side) {
  extern __shared__ float s_row[];           // W floats
  __shared__ int s_prev[ROW_THREADS], s_next[ROW_THREADS];
  const int frame = blockIdx.y, v = blockIdx.x;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd, tid = threadIdx.x, gap = g.gap_eff;
  float* D = ws.Dlr[side] + (size_t)frame * W * H + (size_t)v * W;
  for (int u = tid; u < W; u += ROW_THREADS) s_row[u] = D[u];
  __syncthreads();
  const int chunk = (W + ROW_THREADS - 1) / ROW_THREADS;
  const int lo = min(tid * chunk, W), hi = min(lo + chunk, W);
  int lastv = -1, firstv = W;
  for (int u = lo; u < hi; u++)
    if (s_row[u] >= 0) { lastv = u; if (firstv == W) firstv = u; }
  s_prev[tid] = lastv;
  s_next[tid] = firstv;
  __syncthreads();
  for (int off = 1; off < ROW_THREADS; off <<= 1) {
    int a = (tid >= off) ? s_prev[tid - off] : -1;
    int b = (tid + off < ROW_THREADS) ? s_next[tid + off] : W;
    __syncthreads();
    s_prev[tid] = max(s_prev[tid], a);
    s_next[tid] = min(s_next[tid], b);
    __syncthreads();
  }
  const int first_all = s_next[0], last_all = s_prev[ROW_THREADS - 1];
  int prev = (tid == 0) ? -1 : s_prev[tid - 1];
  for (int u = lo; u < hi; u++) {
    if (s_row[u] >= 0) { prev = u; continue; }
    // next valid at or after u: inside my chunk or from the suffix scan
    int next = W;
    for (int k = u + 1; k < hi; k++)
      if (s_row[k] >= 0) { next = k; break; }
    if (next == W && tid + 1 < ROW_THREADS) next = s_next[tid + 1];
    if (prev >= 0 && next < W) {
      if (next - prev - 1 <= gap) D[u] = ipol(s_row[prev], s_row[next]);
    } else if (g.p.add_corners) {
      if (prev < 0 && first_all < W && first_all - u <= gap) D[u] = s_row[first_all];
      if (next >= W && last_all >= 0 && u - last_all <= gap) D[u] = s_row[last_all];
    }
  }
}

// Gap widths up to SMALL_GAP: one thread per pixel, out of place.  An invalid pixel looks for
// the nearest valid pixel on both sides along the line; the run it sits in is filled iff both
// exist and the run is at most `gap` long (elas.cpp:1137-1155).  Valid pixels never change in a
// pass, so reading the unmodified input is exactly the reference's sequential result.
constexpr int SMALL_GAP = 8;

template <bool ROWS>
__global__ void gap_small_kernel(Geo g, Workspace ws, const float* __restrict__ in, float* __restrict__ out) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd, gap = g.gap_eff;
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t fp = (size_t)frame * W * H;
  const float* D = in + fp;
  const int i = v * W + u;
  float d = D[i];
  if (d < 0) {
    const int pos = ROWS ? u : v, len = ROWS ? W : H, stride = ROWS ? 1 : W;
    int l = 0, r = 0;
    float dl = -1.f, dr = -1.f;
    for (int k = 1; k <= gap && pos - k >= 0; k++) {
      float t = D[i - k * stride];
      if (t >= 0) { l = k; dl = t; break; }
    }
    if (l > 0) {
      for (int k = 1; k <= gap + 1 - l && pos + k < len; k++) {
        float t = D[i + k * stride];
        if (t >= 0) { r = k; dr = t; break; }
      }
      if (r > 0) d = ipol(dl, dr);   // run length l + r - 1 <= gap
    }
  }
  out[fp + i] = d;
}

// Quad version: pixels are valid far more often than not, and a quad without an invalid pixel is
// one 128-bit load and one 128-bit store.
template <bool ROWS>
__global__ void __launch_bounds__(Q_THREADS)
gap_small4_kernel(Geo g, Workspace ws, const float* __restrict__ in, float* __restrict__ out) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd, gap = g.gap_eff;
  const int q = blockIdx.x * Q_THREADS + threadIdx.x, v = blockIdx.y;
  if (4 * q >= W) return;
  const size_t fp = (size_t)frame * W * H;
  const float* D = in + fp;
  const int i0 = v * W + 4 * q;
  const float4 c = *reinterpret_cast<const float4*>(D + i0);
  float d[4] = {c.x, c.y, c.z, c.w};
  if (!(c.x >= 0 && c.y >= 0 && c.z >= 0 && c.w >= 0)) {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      if (d[j] >= 0) continue;
      const int i = i0 + j;
      const int pos = ROWS ? 4 * q + j : v, len = ROWS ? W : H, stride = ROWS ? 1 : W;
      int l = 0, r = 0;
      float dl = -1.f, dr = -1.f;
      for (int k = 1; k <= gap && pos - k >= 0; k++) {
        float t = D[i - k * stride];
        if (t >= 0) { l = k; dl = t; break; }
      }
      if (l > 0) {
        for (int k = 1; k <= gap + 1 - l && pos + k < len; k++) {
          float t = D[i + k * stride];
          if (t >= 0) { r = k; dr = t; break; }
        }
        if (r > 0) d[j] = ipol(dl, dr);   // run length l + r - 1 <= gap
      }
    }
  }
  *reinterpret_cast<float4*>(out + fp + i0) = make_float4(d[0], d[1], d[2], d[3]);
}

__global__ void gap_cols_kernel(Geo g, Workspace ws, int side) {
  const int frame = blockIdx.y;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd, gap = g.gap_eff;
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= W) return;
  float* D = ws.Dlr[side] + (size_t)frame * W * H + u;
  int count = 0;
  float dprev = 0.f;  // value of the last valid pixel
  bool have_prev = false;
  for (int v = 0; v < H; v++) {
    float d = D[(size_t)v * W];
    if (d >= 0) {
      if (count >= 1 && count <= gap && have_prev) {
        float f = ipol(dprev, d);
        for (int k = v - count; k < v; k++) D[(size_t)k * W] = f;
      }
      count = 0;
      dprev = d;
      have_prev = true;
    } else {
      count++;
    }
  }
  if (g.p.add_corners) {
    for (int v = 0; v < H; v++) {
      float d = D[(size_t)v * W];
      if (d >= 0) {
        for (int k = max(v - gap, 0); k < v; k++) D[(size_t)k * W] = d;
        break;
      }
    }
    for (int v = H - 1; v >= 0; v--) {
      float d = D[(size_t)v * W];
      if (d >= 0) {
        for (int k = v; k <= min(v + gap, H - 1); k++) D[(size_t)k * W] = d;
        break;
      }
    }
  }
}

// ------------------------------------------------------------ adaptive mean
__device__ __forceinline__ float buggy_abs(float x) { return __uint_as_float(__float_as_uint(x) & 0x4F000000u); }

// fs / ws, bit for bit.  The weights are 0, 2 or 4 (H2), so ws is an even number up to 32 and a power of
// two wherever the window has no outlier (32 on a smooth surface): the quotient is then an exact
// scaling, done with one multiply by the exactly representable reciprocal (exponent flipped) instead
// of the ~25-instruction IEEE division; other sums take the division.
__device__ __forceinline__ float div_weight_sum(float fs, float ws) {
  const unsigned wb = __float_as_uint(ws);
  if ((wb & 0x007FFFFFu) == 0u) return fs * __uint_as_float(0x7F000000u - wb);
  return fs / ws;
}

// One output of the 8-tap filter.  x[k] = sample at coordinate c-4+k; the reference keeps
// samples in a ring indexed by coordinate mod 8 and adds lanes (s, s+4) first, then
// ((l0+l1)+l2)+l3 (elas.cpp:1425-1432); `rot` = (c-4) & 3 restores that order.
// UNIFORM = rot is the same for the whole warp (vertical pass: a warp works on one row): a
// uniform switch.  Otherwise (horizontal pass: rot = lane & 3) the four pair sums are rotated
// into ring order with selects instead of a four-way divergent branch.
template <bool UNIFORM>
__device__ __forceinline__ bool mean8(const float x[8], float centre, int rot, float& out) {
  float w[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    if (k == 4) {   // the centre sample itself (both callers pass centre = x[4]): difference 0, weight 4
      w[k] = 4.0f;
      continue;
    }
    float t = 4.0f - buggy_abs(x[k] - centre);
    w[k] = fmaxf(0.0f, t);
  }
  // The weights are exactly 0, 2 or 4 (buggy_abs returns a power of two >= 8, 2, or something below 2^-96), so
  // every product x * w is exact and the pair sum x[k]*w[k] + x[k+4]*w[k+4] is one multiply and one FMA with
  // the reference's rounding (the file is compiled with -fmad=false; this FMA is explicit).
  float wq[4], fq[4];
#pragma unroll
  for (int k = 0; k < 4; k++) { wq[k] = w[k] + w[k + 4]; fq[k] = __fmaf_rn(x[k + 4], w[k + 4], x[k] * w[k]); }
  // lane s holds pair k with (rot + k) & 3 == s  ->  k = (s - rot) & 3
  float ws, fs;
  if (UNIFORM) {
    switch (rot) {
      case 0: ws = ((wq[0] + wq[1]) + wq[2]) + wq[3]; fs = ((fq[0] + fq[1]) + fq[2]) + fq[3]; break;
      case 1: ws = ((wq[3] + wq[0]) + wq[1]) + wq[2]; fs = ((fq[3] + fq[0]) + fq[1]) + fq[2]; break;
      case 2: ws = ((wq[2] + wq[3]) + wq[0]) + wq[1]; fs = ((fq[2] + fq[3]) + fq[0]) + fq[1]; break;
      default: ws = ((wq[1] + wq[2]) + wq[3]) + wq[0]; fs = ((fq[1] + fq[2]) + fq[3]) + fq[0]; break;
    }
  } else {
    const bool r1 = rot & 1, r2 = rot & 2;
    float a[4], b[4], c[4], d[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { a[j] = r1 ? wq[(j + 3) & 3] : wq[j]; b[j] = r1 ? fq[(j + 3) & 3] : fq[j]; }
#pragma unroll
    for (int j = 0; j < 4; j++) { c[j] = r2 ? a[(j + 2) & 3] : a[j]; d[j] = r2 ? b[(j + 2) & 3] : b[j]; }
    ws = ((c[0] + c[1]) + c[2]) + c[3];
    fs = ((d[0] + d[1]) + d[2]) + d[3];
  }
  if (ws > 0) {
    const float d = div_weight_sum(fs, ws);
    if (d >= 0) { out = d; return true; }
  }
  return false;
}

// Both passes in one kernel: a CTA produces a MT_W x MT_H tile of the output.  The input tile
// (4 / 3 columns and rows of halo) is staged in shared memory once, the horizontally filtered
// rows the vertical pass needs (MT_H + 7 of them) are computed into shared memory and never
// touch HBM: one read and one write of the map instead of three reads and two writes.
constexpr int MT_W = 64, MT_H = 32, MT_IN_W = MT_W + 7, MT_ROWS = MT_H + 7;

constexpr int MI_S = MT_IN_W + 1;    // row stride of the input tile (16-byte aligned rows)

// The two filter passes on a staged tile.  s_in: MT_ROWS x MI_S, rows y0-4 .. y0+MT_H+2, columns x0-4 ..
// x0+MT_W+2 (0 outside the map); s_tmp: MT_ROWS x MT_W scratch for the horizontally filtered rows (the
// reference's D_tmp).  Four neighbouring outputs per thread in both passes: their 8-tap windows overlap
// (11 samples instead of 32) and the ring rotation (c-4) & 3 of output s is s itself (tile origins are
// multiples of 4), a compile-time constant, so the summation order needs no selects.
__device__ __forceinline__ void mean_tile_passes(const float* __restrict__ s_in, float* __restrict__ s_tmp, int W, int H,
                                                 int x0, int y0, int tid, float* __restrict__ dst) {
  // D_tmp: filtered for rows 3..H-4 and centres 4..W-4, otherwise -10 (invalid input) or 0 (H1)
  for (int i = tid; i < MT_ROWS * (MT_W / 4); i += 256) {
    const int r = i / (MT_W / 4), c = 4 * (i - r * (MT_W / 4));
    const int y = y0 - 4 + r;
    const bool yok = W >= 8 && y >= 3 && y <= H - 4;
    // the reference first sets every negative sample to -10 (elas.cpp:1304-1309); after the L/R check
    // (always run before, elas.cpp:108-118) -10 is the only negative value there is
    float v[12];
    const float4 q0 = *reinterpret_cast<const float4*>(&s_in[r * MI_S + c]);
    const float4 q1 = *reinterpret_cast<const float4*>(&s_in[r * MI_S + c + 4]);
    const float4 q2 = *reinterpret_cast<const float4*>(&s_in[r * MI_S + c + 8]);
    v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w; v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
    v[8] = q2.x; v[9] = q2.y; v[10] = q2.z; v[11] = q2.w;
    float o[4];
#pragma unroll
    for (int s4 = 0; s4 < 4; s4++) {
      const int x = x0 + c + s4;
      const float d = v[s4 + 4];
      o[s4] = (d < 0) ? -10.f : 0.f;
      if (yok && x >= 4 && x <= W - 4) {
        float m;
        if (mean8<true>(v + s4, d, s4, m)) o[s4] = m;
      }
    }
    *reinterpret_cast<float4*>(&s_tmp[r * MT_W + c]) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  for (int i = tid; i < (MT_H / 4) * MT_W; i += 256) {
    const int rq = i / MT_W, c = i - rq * MT_W;
    const int x = x0 + c;
    if (x >= W) continue;
    const bool xok = H >= 8 && x >= 3 && x <= W - 4;
    float v[11];
#pragma unroll
    for (int k = 0; k < 11; k++) v[k] = s_tmp[(4 * rq + k) * MT_W + c];   // rows y-4 .. y+3 of the four outputs
#pragma unroll
    for (int s4 = 0; s4 < 4; s4++) {
      const int r = 4 * rq + s4, y = y0 + r;
      if (y >= H) continue;
      float o = s_in[(r + 4) * MI_S + c + 4];
      if (xok && y >= 4 && y <= H - 4) {
        float m;
        if (mean8<true>(v + s4, v[s4 + 4], s4, m)) o = m;
      }
      dst[(size_t)y * W + x] = o;
    }
  }
}

__global__ void __launch_bounds__(256)
mean_fused_kernel(Geo g, Workspace ws, const float* __restrict__ in, float* __restrict__ out,
                  size_t out_frame_stride) {
  __shared__ __align__(16) float s_in[MT_ROWS * MI_S];
  __shared__ __align__(16) float s_tmp[MT_ROWS * MT_W];
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd, tid = threadIdx.x;
  const int x0 = blockIdx.x * MT_W, y0 = blockIdx.y * MT_H;
  const float* src = in + (size_t)frame * W * H;
  // a warp per tile row, lanes along the row (71 columns = 3 passes): no index division; all 15 loads of a
  // thread are issued before the first one is stored
  {
    constexpr int NR = (MT_ROWS + 7) / 8, NC = (MT_IN_W + 31) / 32;
    float t[NR][NC];
#pragma unroll
    for (int i = 0; i < NR; i++) {
      const int r = (tid >> 5) + 8 * i, y = y0 - 4 + r;
      const bool yin = r < MT_ROWS && y >= 0 && y < H;
      const float* row = src + (size_t)(yin ? y : 0) * W;
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const int c = (tid & 31) + 32 * k, x = x0 - 4 + c;
        t[i][k] = (yin && c < MT_IN_W && x >= 0 && x < W) ? row[x] : 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NR; i++) {
      const int r = (tid >> 5) + 8 * i;
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const int c = (tid & 31) + 32 * k;
        if (r < MT_ROWS && c < MT_IN_W) s_in[r * MI_S + c] = t[i][k];
      }
    }
  }
  __syncthreads();
  mean_tile_passes(s_in, s_tmp, W, H, x0, y0, tid, out + (size_t)frame * out_frame_stride);
}

// Half-resolution branch (elas.cpp:1323-1391): 4 taps at coordinates c-2 .. c+1, centre c.  The
// reference's ring is indexed by coordinate mod 4 and summed ((l0+l1)+l2)+l3; rot = (c-2) & 3.
__device__ __forceinline__ bool mean4(const float x[4], float centre, int rot, float& out) {
  float w[4], f[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (k == 2) {   // the centre sample itself (both callers pass centre = x[2])
      w[k] = 4.0f;
      f[k] = x[k] * 4.0f;
      continue;
    }
    float t = 4.0f - buggy_abs(x[k] - centre);
    w[k] = fmaxf(0.0f, t);
    f[k] = x[k] * w[k];
  }
  float ws, fs;
  switch (rot) {
    case 0: ws = ((w[0] + w[1]) + w[2]) + w[3]; fs = ((f[0] + f[1]) + f[2]) + f[3]; break;
    case 1: ws = ((w[3] + w[0]) + w[1]) + w[2]; fs = ((f[3] + f[0]) + f[1]) + f[2]; break;
    case 2: ws = ((w[2] + w[3]) + w[0]) + w[1]; fs = ((f[2] + f[3]) + f[0]) + f[1]; break;
    default: ws = ((w[1] + w[2]) + w[3]) + w[0]; fs = ((f[1] + f[2]) + f[3]) + f[0]; break;
  }
  if (ws > 0) {
    const float d = div_weight_sum(fs, ws);
    if (d >= 0) { out = d; return true; }
  }
  return false;
}

// horizontal: rows 3..H-4, centres 2..W-2; elsewhere -10 (invalid input) or 0 (unwritten, H1)
__global__ void mean4_h_kernel(Geo g, Workspace ws, const float* __restrict__ in, float* __restrict__ tmp) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int c = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (c >= W) return;
  const float* row = in + (size_t)frame * W * H + (size_t)v * W;
  float d = row[c];
  float o = (d < 0) ? -10.f : 0.f;
  if (W >= 4 && v >= 3 && v <= H - 4 && c >= 2 && c <= W - 2) {
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = row[c - 2 + k];   // negatives are exactly -10 here, see mean_fused_kernel
    float m;
    if (mean4(x, x[2], (c - 2) & 3, m)) o = m;
  }
  tmp[(size_t)frame * W * H + (size_t)v * W + c] = o;
}

// vertical: columns 3..W-4, centres 2..H-2; every other pixel keeps `in`
__global__ void mean4_v_kernel(Geo g, Workspace ws, const float* __restrict__ in, const float* __restrict__ tmp,
                               float* __restrict__ out, size_t out_frame_stride) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int u = blockIdx.x * blockDim.x + threadIdx.x, c = blockIdx.y;
  if (u >= W) return;
  const size_t fp = (size_t)frame * W * H;
  float o = in[fp + (size_t)c * W + u];
  if (H >= 4 && u >= 3 && u <= W - 4 && c >= 2 && c <= H - 2) {
    const float* col = tmp + fp + u;
    float x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) x[k] = col[(size_t)(c - 2 + k) * W];
    float m;
    if (mean4(x, x[2], (c - 2) & 3, m)) o = m;
  }
  out[(size_t)frame * out_frame_stride + (size_t)c * W + u] = o;
}

// ------------------------------------------------------------ median
__device__ __forceinline__ float median7(float v[7]) {
  // insertion sort, as elas.cpp:1518-1527 (NaNs cannot occur)
#pragma unroll
  for (int j = 1; j < 7; j++) {
    float t = v[j];
    int i = j - 1;
    while (i >= 0 && v[i] > t) { v[i + 1] = v[i]; i--; }
    v[i + 1] = t;
  }
  return v[3];
}

// horizontal pass into a zero-initialised temporary (calloc, elas.cpp:1505)
__global__ void median_h_kernel(Geo g, Workspace ws, const float* __restrict__ in, float* __restrict__ tmp) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t a = (size_t)frame * W * H + (size_t)v * W + u;
  float o = 0.f;
  if (u >= 3 && u <= W - 4 && v >= 3 && v <= H - 4) {
    float d = in[a];
    o = d;
    if (d >= 0) {
      float x[7];
#pragma unroll
      for (int k = 0; k < 7; k++) x[k] = in[a - 3 + k];
      o = median7(x);
    }
  }
  tmp[a] = o;
}

__global__ void median_v_kernel(Geo g, Workspace ws, const float* __restrict__ in, const float* __restrict__ tmp,
                                float* __restrict__ out, size_t out_frame_stride) {
  const int frame = blockIdx.z;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.Wd, H = g.Hd;
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t a = (size_t)frame * W * H + (size_t)v * W + u;
  float o = in[a];
  if (u >= 3 && u <= W - 4 && v >= 3 && v <= H - 4 && o >= 0) {
    float x[7];
#pragma unroll
    for (int k = 0; k < 7; k++) x[k] = tmp[a + (ptrdiff_t)(k - 3) * W];
    o = median7(x);
  }
  out[(size_t)frame * out_frame_stride + (size_t)v * W + u] = o;
}

__global__ void copy_out_kernel(Geo g, Workspace ws, const float* __restrict__ in, float* __restrict__ out,
                                int32_t* __restrict__ status) {
  const int frame = blockIdx.y;
  const int st = ws.info[frame].status;
  const size_t n = (size_t)g.Wd * g.Hd;
  if (status && blockIdx.x == 0 && threadIdx.x == 0) status[frame] = st;
  if (st != JN_OK || out == nullptr) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[(size_t)frame * n + i] = in[(size_t)frame * n + i];
}

}  // namespace

// Step-wise entry points (the debug dump calls them one at a time).
void post_lr(const Geo& g, int B, Workspace& ws, cudaStream_t s, bool need_right) {
  const dim3 qg((g.Wd / 4 + Q_THREADS - 1) / Q_THREADS, g.Hd, B);
  if (g.Wd % 4 == 0 && need_right) lr4_kernel<true><<<qg, Q_THREADS, 0, s>>>(g, ws);
  else if (g.Wd % 4 == 0) lr4_kernel<false><<<qg, Q_THREADS, 0, s>>>(g, ws);
  else lr_kernel<<<dim3((g.Wd + 255) / 256, g.Hd, B), 256, 0, s>>>(g, ws);
  g_jn_launches += 1;
}

void post_segments(const Geo& g, int B, Workspace& ws, int side, cudaStream_t s) {
  const int n = g.Wd * g.Hd;
  seg_tile_kernel<<<dim3((g.Wd + CT_W - 1) / CT_W, (g.Hd + CT_H - 1) / CT_H, B), CT_THREADS, 0, s>>>(g, ws, side);
  const int nb = ((g.Hd - 1) / CT_H) * g.Wd + ((g.Wd - 1) / CT_W) * g.Hd;
  if (nb > 0) seg_border_kernel<<<dim3((nb + 255) / 256, B), 256, 0, s>>>(g, ws, side);
  if (n % 4 == 0) seg_roots_kernel<true><<<dim3((n / 4 + 255) / 256, B), 256, 0, s>>>(g, ws);
  else seg_roots_kernel<false><<<dim3((n + 255) / 256, B), 256, 0, s>>>(g, ws);
  if (g.Wd % 4 == 0) {
    dim3 qg((g.Wd / 4 + Q_THREADS - 1) / Q_THREADS, g.Hd, B);
    seg_apply4_kernel<<<qg, Q_THREADS, 0, s>>>(g, ws, side);
  } else {
    seg_apply_kernel<<<dim3((g.Wd + 255) / 256, g.Hd, B), 256, 0, s>>>(g, ws, side);
  }
  g_jn_launches += nb > 0 ? 4 : 3;
}

void post_gap(const Geo& g, int B, Workspace& ws, int side, cudaStream_t s) {
  if (g.gap_eff <= SMALL_GAP && !g.p.add_corners) {
    if (g.Wd % 4 == 0) {
      dim3 qg((g.Wd / 4 + Q_THREADS - 1) / Q_THREADS, g.Hd, B);
      gap_small4_kernel<true><<<qg, Q_THREADS, 0, s>>>(g, ws, ws.Dlr[side], ws.Dtmp[side]);
      gap_small4_kernel<false><<<qg, Q_THREADS, 0, s>>>(g, ws, ws.Dtmp[side], ws.Dlr[side]);
    } else {
      dim3 pg((g.Wd + 255) / 256, g.Hd, B);
      gap_small_kernel<true><<<pg, 256, 0, s>>>(g, ws, ws.Dlr[side], ws.Dtmp[side]);
      gap_small_kernel<false><<<pg, 256, 0, s>>>(g, ws, ws.Dtmp[side], ws.Dlr[side]);
    }
  } else {
    gap_rows_kernel<<<dim3(g.Hd, B), ROW_THREADS, g.Wd * sizeof(float), s>>>(g, ws, side);
    gap_cols_kernel<<<dim3((g.Wd + 63) / 64, B), 64, 0, s>>>(g, ws, side);
  }
  g_jn_launches += 2;
}

// in -> out (frame stride of out given in floats); tmp = scratch
void post_mean(const Geo& g, int B, Workspace& ws, const float* in, float* tmp, float* out, size_t ostride,
               cudaStream_t s) {
  if (g.p.subsampling) {
    dim3 pg((g.Wd + 255) / 256, g.Hd, B);
    mean4_h_kernel<<<pg, 256, 0, s>>>(g, ws, in, tmp);
    mean4_v_kernel<<<pg, 256, 0, s>>>(g, ws, in, tmp, out, ostride);
    g_jn_launches += 2;
    return;
  }
  (void)tmp;   // the horizontally filtered rows live in shared memory only
  dim3 grid((g.Wd + MT_W - 1) / MT_W, (g.Hd + MT_H - 1) / MT_H, B);
  mean_fused_kernel<<<grid, 256, 0, s>>>(g, ws, in, out, ostride);
  g_jn_launches += 1;
}

void post_median(const Geo& g, int B, Workspace& ws, const float* in, float* tmp, float* out, size_t ostride,
                 cudaStream_t s) {
  dim3 pg((g.Wd + 255) / 256, g.Hd, B);
  median_h_kernel<<<pg, 256, 0, s>>>(g, ws, in, tmp);
  median_v_kernel<<<pg, 256, 0, s>>>(g, ws, in, tmp, out, ostride);
  g_jn_launches += 2;
}

void post_copy(const Geo& g, int B, Workspace& ws, const float* in, float* out, int32_t* status, cudaStream_t s) {
  copy_out_kernel<<<dim3(148, B), 256, 0, s>>>(g, ws, in, out, status);
  g_jn_launches += 1;
}

// The whole chain for a batch.  D1out/D2out: user buffers (B * W*H floats); D2out may be NULL.
void launch_post(const Geo& g, int B, Workspace& ws, float* D1out, float* D2out, int32_t* status, cudaStream_t s) {
  const size_t n = (size_t)g.Wd * g.Hd;
  const int sides = g.p.postprocess_only_left ? 1 : 2;
  post_lr(g, B, ws, s, sides == 2 || D2out != nullptr);
  for (int side = 0; side < sides; side++) {
    post_segments(g, B, ws, side, s);
    post_gap(g, B, ws, side, s);
  }
  for (int side = 0; side < 2; side++) {
    float* out = side ? D2out : D1out;
    if (out == nullptr) continue;
    const float* cur = ws.Dlr[side];
    const bool filt = side < sides;
    const bool do_mean = filt && g.p.filter_adaptive_mean, do_med = filt && g.p.filter_median;
    if (do_mean && do_med) {
      post_mean(g, B, ws, cur, ws.Dtmp[side], ws.Dtmp2[side], n, s);
      post_median(g, B, ws, ws.Dtmp2[side], ws.Dtmp[side], out, n, s);
    } else if (do_mean) {
      post_mean(g, B, ws, cur, ws.Dtmp[side], out, n, s);
    } else if (do_med) {
      post_median(g, B, ws, cur, ws.Dtmp[side], out, n, s);
    } else {
      post_copy(g, B, ws, cur, out, nullptr, s);
    }
  }
  if (status) post_copy(g, B, ws, ws.Dlr[0], nullptr, status, s);
}
