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
//   dense_kernel   one thread per column walking 8 rows, a warp = a run of 32 pixels
//                  of one row: plane prior of the recorded triangle, candidate set =
//                  the cell's sorted candidate list (bit set for overflowing cells)
//                  outside the plane range, then the plane range with the
//                  integer prior P, 16-byte SAD per candidate on the integer
//                  pipe (4 x VABSDIFF4.U8.ACC), strict '<' in the reference's
//                  evaluation order (H7).  Left descriptors are read once, fully
//                  coalesced (512 B per warp); the right descriptors of one
//                  candidate disparity form one contiguous 512-byte span per
//                  warp and neighbouring disparities overlap in L1.
//
// Roofline: HBM (72 N bytes per frame: 2 passes x (2 descriptor images 32 N +
// 4 N written)), second roofline the integer pipe (16 byte-absdiffs = 4
// instructions per candidate).  Measured: DRAM 35 % busy, issue slots 81 % with
// all 64 warps per SM resident at 32 registers: bound by instruction issue and
// the dependent loads triangle id -> plane -> candidates of every row (variants
// with more registers per thread, prefetches or streaming loads were all slower,
// see DESIGN.md section 3).  Compiled with -fmad=false: d_plane and the edge
// equations must round exactly as the reference's SSE code does.
#include "common.cuh"

namespace {

__device__ __forceinline__ int f2u_lo32(float x) { return (int)(unsigned)(unsigned long long)__float2ll_rz(x); }

// One column span [vlo, vhi) of triangle i.  Scan-converted triangles of a planar triangulation only
// ever overlap in the first or last row of a span (neighbours evaluate the same edge function, so
// spans abut exactly; what overlaps there is comes from edges meeting at a vertex and float
// rounding): "the later triangle wins" (atomicMax) is needed for those two rows only, the rows in
// between belong to this triangle alone and take a plain store.  Checked on the CPU restatement
// (tests/test_oracle_pin.py::test_raster_overlaps_stay_on_span_borders): 0 of 61 M covered pixels
// over 3 000 random and 37 pipeline triangulations had a second cover strictly inside a span.
// The argument needs the float edge evaluation a*u + b to be off by less than one row; its error grows
// with the magnitude of the operands, so the plain stores are only used while the image is small
// enough for that (`exact`: W, H <= 2048, where |a*u|, |b| < 2^23 keep the rounding error of the sum
// below 1/4 row -- checked at 2048 x 2048 with steep hull edges by the same test); larger images
// take atomicMax on every row.
__device__ __forceinline__ void write_span(int* __restrict__ map, int W, int u, int vlo, int vhi, int i, bool exact) {
  if (vhi <= vlo) return;
  if (!exact) {
    for (int v = vlo; v < vhi; v++) atomicMax(&map[v * W + u], i);
    return;
  }
  atomicMax(&map[vlo * W + u], i);
  if (vhi - 1 > vlo) atomicMax(&map[(vhi - 1) * W + u], i);
  for (int v = vlo + 1; v < vhi - 1; v++) map[v * W + u] = i;
}

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
  const bool exact = W <= 2048 && H <= 2048;
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
        write_span(map, W, u, vlo, vhi, i, exact);
      }
    if ((int)Bu != (int)Cu)
      for (int u = max((int)Bu, 0) + lane; u < min((int)Cu, W); u += 32) {
        int v1 = f2u_lo32(ACa * (float)u + ACb), v2 = f2u_lo32(BCa * (float)u + BCb);
        int vlo = max(min(v1, v2), 0), vhi = min(max(v1, v2), H);
        write_span(map, W, u, vlo, vhi, i, exact);
      }
  }
}

// Tile shape: columns per CTA x rows per thread.  Swept on the B200 at the end of round 1
// (64..256 x 4..16): all within 3 %, 256 x 8 best (wider tiles re-read less of the searched rows).
#ifndef JN_DENSE_THREADS
#define JN_DENSE_THREADS 256
#endif
#ifndef JN_DENSE_ROWS
#define JN_DENSE_ROWS 8
#endif
constexpr int DENSE_THREADS = JN_DENSE_THREADS;
constexpr int DENSE_ROWS = JN_DENSE_ROWS;         // image rows per thread (amortises the per-thread setup)
constexpr unsigned KEY_NONE = 0xFFFFFFFFu;

// One candidate: 16-byte SAD (+ prior), folded into a packed key
//   (cost + 2048) << 13 | class << 12 | d      class 0 = grid candidate, 1 = plane range
// The reference keeps the FIRST minimum in its evaluation order (grid candidates outside the
// plane range ascending, then the plane range ascending; strict '<', H7) = the smallest key.
// Bf = descriptors of the searched image (frame base), rowoff = first pixel of the row:
// 32-bit index arithmetic, one IMAD.WIDE per address.
__device__ __forceinline__ void eval_candidate(unsigned& best, const uint4& a, const uint4* __restrict__ Bf,
                                               unsigned rowoff, int u, int dir, int d, unsigned addend,
                                               unsigned wm4) {
  const int uw = u + dir * d;
  if ((unsigned)(uw - 2) < wm4) {   // u_warp in [2, W-3] (a branch-free variant measured slower)
    const uint4 b = __ldg(Bf + (rowoff + (unsigned)uw));
    best = min(best, sad16(a, b, 0u) * 8192u + (addend + (unsigned)d));
  }
}

__device__ __forceinline__ void eval_word(unsigned& best, unsigned bits, int w, int lo, int hi, const uint4& a,
                                          const uint4* __restrict__ Bf, unsigned rowoff, int u, int dir,
                                          unsigned wm4) {
  if (bits == 0u) return;
  const int l = lo - 32 * w, h = hi - 32 * w;     // plane range relative to this word
  if (h >= 0 && l <= 31) bits &= ~((0xFFFFFFFFu << max(l, 0)) & (0xFFFFFFFFu >> (31 - min(h, 31))));
  while (bits) {
    const int b = __ffs(bits) - 1;
    bits &= bits - 1;
    eval_candidate(best, a, Bf, rowoff, u, dir, 32 * w + b, 2048u << 13, wm4);
  }
}

// R = plane radius known at compile time (2: ROBOTICS, 3: MIDDLEBURY), 0 = run-time radius.
// SUB = subsampling: only pixels with even u and v are matched (elas.cpp:877-896) and the result
// goes to (u/2, v/2) of the W/2 x H/2 map (elas.cpp:693); everything else is unchanged.
template <int R, bool SUB>
__global__ void __launch_bounds__(DENSE_THREADS)
dense_kernel(Geo g, Workspace ws) {
  const int side = blockIdx.z & 1, frame = blockIdx.z >> 1;
  if (ws.info[frame].status != JN_OK) return;
  const int W = g.W, H = g.H;
  const int um = blockIdx.x * DENSE_THREADS + threadIdx.x;     // map column
  if (um >= g.Wd) return;
  const int u = SUB ? 2 * um : um;
  // frame-level bases once per thread; everything below is 32-bit offsets from them
  const size_t fpix = (size_t)frame * W * H;
  const uint4* __restrict__ Af = reinterpret_cast<const uint4*>(ws.desc[side] + fpix * 16);
  const uint4* __restrict__ Bf = reinterpret_cast<const uint4*>(ws.desc[side ^ 1] + fpix * 16);
  const int* __restrict__ tmap = ws.trimap[side] + fpix;
  float* __restrict__ outp = ws.Draw[side] + (SUB ? (size_t)frame * g.Wd * g.Hd : fpix);
  const float* __restrict__ planes = ws.planes[side] + (size_t)frame * g.cap_t * 6;
  const uint32_t* __restrict__ masks = ws.gridmask[side] + (size_t)frame * g.gw * g.gh * g.gwords;
  const uint16_t* __restrict__ lists = ws.gridlist[side] + (size_t)frame * g.gw * g.gh * GRID_LIST;
  const unsigned gx = __umulhi((unsigned)u, g.gs_magic);
  const int dir = side ? 1 : -1;
  const unsigned wm4 = (unsigned)(W - 4);
  const bool u_ok = u >= 2 && u < W - 2;
  const int po = side ? 3 : 0;              // this image's plane, the other image's slope
  const int P0 = g.P[0], P1 = g.P[1], P2 = g.P[2], P3 = g.P[3];
  const int v0 = blockIdx.y * DENSE_ROWS;
#pragma unroll 1
  for (int vm = v0; vm < min(v0 + DENSE_ROWS, g.Hd); vm++) {
    const int v = SUB ? 2 * vm : vm;
    const unsigned pix = (unsigned)(v * W + u);
    const unsigned rowoff = (unsigned)(max(min(v, H - 3), 2) * W);
    // independent loads first: triangle id, own descriptor
    const int t = __ldg(tmap + pix);
    const uint4 a = __ldg(Af + (rowoff + (unsigned)u));
    float out = -10.f;
    if (t >= 0 && u_ok && (int)texture16(a) >= g.p.match_texture) {
      const float* pl = planes + (unsigned)t * 6u;
      const float pa = __ldg(pl + po), pb = __ldg(pl + po + 1), pc = __ldg(pl + po + 2), pd = __ldg(pl + 3 - po);
      const unsigned gy = __umulhi((unsigned)v, g.gs_magic);
      const int d_plane = (int)(pa * (float)u + pb * (float)v + pc);
      const int r = R ? R : g.plane_radius;
      const int lo = max(d_plane - r, 0), hi = min(d_plane + r, g.p.disp_max);
      // (double)|x| < 0.7  <=>  |x| <= 0.7f  (0.7f is the largest float below 0.7)
      const bool valid = fabsf(pa) <= 0.7f && fabsf(pd) <= 0.7f;
      unsigned best = KEY_NONE;
      // grid candidates outside the plane range: compact sorted list, 2 x 128-bit loads
      const unsigned ci = gy * (unsigned)g.gw + gx;
      const uint4* lst = reinterpret_cast<const uint4*>(lists + ci * GRID_LIST);
      const uint4 l0 = __ldg(lst);
      if (l0.x != 0xFFFEFFFEu) {
        const uint4 l1 = __ldg(lst + 1);
        const unsigned wv[8] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        // "d < lo || d > hi" as one unsigned compare; an empty plane range (hi < lo: the plane
        // extrapolates outside [0, disp_max]) excludes nothing
        const int lo2 = (hi < lo) ? 0x7fffffff : lo;
        const unsigned span = (hi < lo) ? 0u : (unsigned)(hi - lo);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (wv[k] == 0xFFFFFFFFu) break;          // sorted: only padding follows
          const int d0 = (int)(wv[k] & 0xFFFFu), d1 = (int)(wv[k] >> 16);
          if ((unsigned)(d0 - lo2) > span) eval_candidate(best, a, Bf, rowoff, u, dir, d0, 2048u << 13, wm4);
          // a 0xFFFF pad in the upper half fails the image-bounds test inside eval_candidate
          if ((unsigned)(d1 - lo2) > span) eval_candidate(best, a, Bf, rowoff, u, dir, d1, 2048u << 13, wm4);
        }
      } else {
        // overflowing cell: decode the bit set
        const uint4* cell = reinterpret_cast<const uint4*>(masks + ci * (unsigned)g.gwords);
        for (int q = 0; q < (g.gwords >> 2); q++) {
          const uint4 m = __ldg(cell + q);
          eval_word(best, m.x, 4 * q + 0, lo, hi, a, Bf, rowoff, u, dir, wm4);
          eval_word(best, m.y, 4 * q + 1, lo, hi, a, Bf, rowoff, u, dir, wm4);
          eval_word(best, m.z, 4 * q + 2, lo, hi, a, Bf, rowoff, u, dir, wm4);
          eval_word(best, m.w, 4 * q + 3, lo, hi, a, Bf, rowoff, u, dir, wm4);
        }
      }
      if (R) {
        // key addend per |k|: (2048 + prior) << 13 | class bit
        const unsigned a0 = ((unsigned)(2048 + (valid ? P0 : 0)) << 13) | (1u << 12);
        const unsigned a1 = ((unsigned)(2048 + (valid ? P1 : 0)) << 13) | (1u << 12);
        const unsigned a2 = ((unsigned)(2048 + (valid ? P2 : 0)) << 13) | (1u << 12);
        const unsigned a3 = ((unsigned)(2048 + (valid ? P3 : 0)) << 13) | (1u << 12);
        // The 2R+1 plane-range candidates are consecutive descriptors of the searched row: issue
        // all loads first (clamped addresses, memory-level parallelism), then score and discard
        // the ones outside [0, disp_max] or outside the image.
        uint4 bb[2 * R + 1];
#pragma unroll
        for (int k = -R; k <= R; k++) {
          const int uw = u + dir * (d_plane + k);
          bb[k + R] = __ldg(Bf + (rowoff + (unsigned)min(max(uw, 2), (int)wm4 + 1)));
        }
#pragma unroll
        for (int k = -R; k <= R; k++) {
          const int d = d_plane + k, uw = u + dir * d;
          const int ak = k < 0 ? -k : k;
          const unsigned add = ak == 0 ? a0 : (ak == 1 ? a1 : (ak == 2 ? a2 : a3));
          const unsigned key = sad16(a, bb[k + R], 0u) * 8192u + (add + (unsigned)d);
          const bool okc = (unsigned)d <= (unsigned)g.p.disp_max && (unsigned)(uw - 2) < wm4;
          best = min(best, okc ? key : KEY_NONE);
        }
      } else {
        for (int d = lo; d <= hi; d++)
          eval_candidate(best, a, Bf, rowoff, u, dir, d,
                         ((unsigned)(2048 + (valid ? g.P[abs(d - d_plane)] : 0)) << 13) | (1u << 12), wm4);
      }
      out = (best != KEY_NONE) ? (float)(best & 0xFFFu) : -1.f;
    }
    outp[SUB ? (unsigned)(vm * g.Wd + um) : pix] = out;
  }
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
  dim3 grid((g.Wd + DENSE_THREADS - 1) / DENSE_THREADS, (g.Hd + DENSE_ROWS - 1) / DENSE_ROWS, 2 * B);
  if (g.p.subsampling) {
    if (g.plane_radius == 2) dense_kernel<2, true><<<grid, DENSE_THREADS, 0, s>>>(g, ws);
    else if (g.plane_radius == 3) dense_kernel<3, true><<<grid, DENSE_THREADS, 0, s>>>(g, ws);
    else dense_kernel<0, true><<<grid, DENSE_THREADS, 0, s>>>(g, ws);
  } else {
    if (g.plane_radius == 2) dense_kernel<2, false><<<grid, DENSE_THREADS, 0, s>>>(g, ws);
    else if (g.plane_radius == 3) dense_kernel<3, false><<<grid, DENSE_THREADS, 0, s>>>(g, ws);
    else dense_kernel<0, false><<<grid, DENSE_THREADS, 0, s>>>(g, ws);
  }
  g_jn_launches += 1;
}

void launch_dense(const Geo& g, int B, Workspace& ws, cudaStream_t s) {
  launch_raster(g, B, ws, s);
  launch_dense_match(g, B, ws, s);
}
