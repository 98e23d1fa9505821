// dense.cu -- dense matching: computeDisparity / findMatch / updatePosteriorMinimum
// (elas.cpp:783-907, 683-780, 661-681).
//
// The reference walks the triangles, scan-converts each one and matches every
// covered pixel; a pixel covered by two triangles keeps the result of the later
// one.  Here that is two passes, both in image order:
//
//   raster_kernel  one warp per triangle replays the reference scan converter
//                  (same float operations, same truncations) and
//                  records, per pixel, the HIGHEST triangle index covering it
//                  (atomicMax = "the later triangle wins");
//   dense_kernel   one thread per pixel, a warp = a run of 32 pixels of one row:
//                  plane prior of the recorded triangle, candidate set = grid
//                  bits outside the plane range, then the plane range with the
//                  integer prior P, 16-byte SAD per candidate on the integer
//                  pipe (4 x VABSDIFF4.U8.ACC), strict '<' in the reference's
//                  evaluation order (H7).  Left descriptors are read once, fully
//                  coalesced (512 B per warp); the right descriptors of one
//                  candidate disparity form one contiguous 512-byte span per
//                  warp and neighbouring disparities overlap in L1.
//
// Roofline: HBM (72 N bytes per frame: 2 passes x (2 descriptor images 32 N +
// 4 N written)), second roofline the integer pipe (16 byte-absdiffs = 4
// instructions per candidate).  Compiled with -fmad=false: d_plane and the edge
// equations must round exactly as the reference's SSE code does.
#include "common.cuh"

namespace {

__device__ __forceinline__ int f2u_lo32(float x) { return (int)(unsigned)(unsigned long long)__float2ll_rz(x); }

__global__ void raster_kernel(Geo g, Workspace ws) {
  const int side = blockIdx.y, frame = blockIdx.z;
  const FrameInfo* info = ws.info + frame;
  if (info->status != JN_OK) return;
  const int nt = info->n_tri[side];
  const int W = g.W, H = g.H;
  const int4* sup = reinterpret_cast<const int4*>(ws.sup) + (size_t)frame * g.cap_s;
  const int* tri = ws.tri[side] + (size_t)frame * g.cap_t * 3;
  int* map = ws.trimap[side] + (size_t)frame * W * H;
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int i = warp; i < nt; i += nwarps) {
    float tu[3], tv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      int4 s = sup[tri[3 * i + k]];
      tu[k] = side ? (float)(s.x - s.z) : (float)s.x;
      tv[k] = (float)s.y;
    }
    // the reference's 3-element exchange sort on u (elas.cpp:847-854)
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int k = 0; k < j; k++)
        if (tu[k] > tu[j]) {
          float t = tu[j]; tu[j] = tu[k]; tu[k] = t;
          t = tv[j]; tv[j] = tv[k]; tv[k] = t;
        }
    const float Au = tu[0], Av = tv[0], Bu = tu[1], Bv = tv[1], Cu = tu[2], Cv = tv[2];
    float ABa = 0, ACa = 0, BCa = 0;
    if ((int)Au != (int)Bu) ABa = (Av - Bv) / (Au - Bu);
    if ((int)Au != (int)Cu) ACa = (Av - Cv) / (Au - Cu);
    if ((int)Bu != (int)Cu) BCa = (Bv - Cv) / (Bu - Cu);
    const float ABb = Av - ABa * Au, ACb = Av - ACa * Au, BCb = Bv - BCa * Bu;
    if ((int)Au != (int)Bu)
      for (int u = max((int)Au, 0) + lane; u < min((int)Bu, W); u += 32) {
        int v1 = f2u_lo32(ACa * (float)u + ACb), v2 = f2u_lo32(ABa * (float)u + ABb);
        int vlo = max(min(v1, v2), 0), vhi = min(max(v1, v2), H);
        for (int v = vlo; v < vhi; v++) atomicMax(&map[v * W + u], i);
      }
    if ((int)Bu != (int)Cu)
      for (int u = max((int)Bu, 0) + lane; u < min((int)Cu, W); u += 32) {
        int v1 = f2u_lo32(ACa * (float)u + ACb), v2 = f2u_lo32(BCa * (float)u + BCb);
        int vlo = max(min(v1, v2), 0), vhi = min(max(v1, v2), H);
        for (int v = vlo; v < vhi; v++) atomicMax(&map[v * W + u], i);
      }
  }
}

constexpr int DENSE_THREADS = 128;

__global__ void __launch_bounds__(DENSE_THREADS)
dense_kernel(Geo g, Workspace ws) {
  const int side = blockIdx.z & 1, frame = blockIdx.z >> 1;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.W, H = g.H;
  const int u = blockIdx.x * DENSE_THREADS + threadIdx.x, v = blockIdx.y;
  if (u >= W) return;
  const size_t fpix = (size_t)frame * W * H;
  const int t = ws.trimap[side][fpix + (size_t)v * W + u];
  float out = -10.f;
  if (t >= 0 && u >= 2 && u < W - 2) {
    const int vl = max(min(v, H - 3), 2);
    const uint4* A = reinterpret_cast<const uint4*>(ws.desc[side] + fpix * 16) + (size_t)vl * W;
    const uint4* Bd = reinterpret_cast<const uint4*>(ws.desc[side ^ 1] + fpix * 16) + (size_t)vl * W;
    const uint4 a = __ldg(A + u);
    if ((int)texture16(a) >= g.p.match_texture) {
      const float* pl = ws.planes[side] + ((size_t)frame * g.cap_t + t) * 6;
      float pa, pb, pc, pd;
      if (!side) { pa = pl[0]; pb = pl[1]; pc = pl[2]; pd = pl[3]; }
      else { pa = pl[3]; pb = pl[4]; pc = pl[5]; pd = pl[0]; }
      const int d_plane = (int)(pa * (float)u + pb * (float)v + pc);
      const int r = g.plane_radius;
      const int lo = max(d_plane - r, 0), hi = min(d_plane + r, g.p.disp_max);
      const bool valid = (double)fabsf(pa) < 0.7 && (double)fabsf(pd) < 0.7;
      const int gs = g.p.grid_size;
      const uint32_t* cell = ws.gridmask[side] +
                             ((size_t)frame * g.gw * g.gh + (size_t)(v / gs) * g.gw + (u / gs)) * g.gwords;
      const int dir = side ? 1 : -1;
      int min_val = 10000, min_d = -1;
      for (int w = 0; w < g.gwords; w++) {
        uint32_t bits = __ldg(cell + w);
        while (bits) {
          int b = __ffs(bits) - 1;
          bits &= bits - 1;
          int d = (w << 5) + b;
          if (d < lo || d > hi) {
            int uw = u + dir * d;
            if (uw >= 2 && uw < W - 2) {
              int val = (int)sad16(a, __ldg(Bd + uw), 0u);
              if (val < min_val) { min_val = val; min_d = d; }
            }
          }
        }
      }
      for (int d = lo; d <= hi; d++) {
        int uw = u + dir * d;
        if (uw >= 2 && uw < W - 2) {
          int val = (int)sad16(a, __ldg(Bd + uw), 0u) + (valid ? g.P[abs(d - d_plane)] : 0);
          if (val < min_val) { min_val = val; min_d = d; }
        }
      }
      out = (min_d >= 0) ? (float)min_d : -1.f;
    }
  }
  ws.Draw[side][fpix + (size_t)v * W + u] = out;
}

}  // namespace

void launch_raster(const Geo& g, int B, Workspace& ws, cudaStream_t s) {
  size_t mbytes = (size_t)B * g.W * g.H * sizeof(int32_t);
  cudaMemsetAsync(ws.trimap[0], 0xff, mbytes, s);   // -1 = no triangle
  cudaMemsetAsync(ws.trimap[1], 0xff, mbytes, s);
  raster_kernel<<<dim3(64, 2, B), 256, 0, s>>>(g, ws);
  g_jn_launches += 1;
}

void launch_dense_match(const Geo& g, int B, Workspace& ws, cudaStream_t s) {
  dense_kernel<<<dim3((g.W + DENSE_THREADS - 1) / DENSE_THREADS, g.H, 2 * B), DENSE_THREADS, 0, s>>>(g, ws);
  g_jn_launches += 1;
}

void launch_dense(const Geo& g, int B, Workspace& ws, cudaStream_t s) {
  launch_raster(g, B, ws, s);
  launch_dense_match(g, B, ws, s);
}
