// dense.cu -- dense matching: computeDisparity / findMatch / updatePosteriorMinimum
// (elas.cpp:783-907, 683-780, 661-681).
//
// The reference walks the triangles, scan-converts each one and matches every
// covered pixel; a pixel covered by two triangles keeps the result of the later
// one.  Here that is two passes, both in image order:
//
//   raster_kernel  sixteen lanes per triangle replay the reference scan converter
//                  (same float operations, same truncations).  Everything
//                  that is per triangle is finished here: each covered pixel gets one
//                  32-bit "plane map" entry
//                      (triangle index + 1) << (dbits + 1) | plane valid << dbits | d_plane + r + 1
//                  with d_plane = (int)(a*u + b*v + c) of this triangle at this pixel
//                  (elas.cpp:722; clamped to [-(r+1), disp_max + r + 1], which keeps every
//                  in-range candidate and the emptiness of the range).  atomicMax on the
//                  entry = "the later triangle wins", and the winner's prior travels with it.
//   dense_kernel   one thread per column walking DENSE_ROWS rows, a warp = a run of 32
//                  pixels of one row.  Per pixel: one map entry (no triangle -> plane ->
//                  d_plane chain of dependent loads), the own descriptor, then the
//                  candidates: the cell's disparity set minus the plane range, then the
//                  plane range with the integer prior P.  The set comes as its few non-zero
//                  32-bit words (grid_words_kernel), loaded once per cell row into registers;
//                  "minus the plane range" is a mask operation and the survivors are walked
//                  with find-leading-one.  16-byte SAD per candidate on the integer pipe
//                  (4 x VABSDIFF4.U8.ACC), folded into an order-preserving key so that the
//                  reference's "first minimum in evaluation order" (H7) is a plain min.
//                  A warp whose 32 columns cannot leave the image for any disparity and whose
//                  plane ranges are all inside [0, disp_max] takes a path without any
//                  per-candidate test (the 2R+1 plane-range descriptors are then consecutive:
//                  one address, immediate offsets); other warps take the checked path.
//                  Left descriptors are read once, fully coalesced (512 B per warp); the
//                  right descriptors of one candidate disparity form one contiguous
//                  512-byte span per warp and neighbouring disparities overlap in L1.
//
// Roofline: HBM (72 N bytes per frame: 2 passes x (2 descriptor images 32 N +
// 4 N written)), second roofline the integer pipe / issue slots (16 byte-absdiffs = 4
// instructions per candidate).  Compiled with -fmad=false: d_plane and the edge
// equations must round exactly as the reference's SSE code does.
#include "common.cuh"

namespace {

__device__ __forceinline__ int f2u_lo32(float x) { return (int)(unsigned)(unsigned long long)__float2ll_rz(x); }

// One column span [vlo, vhi) of triangle i, plane d = a*u + b*v + c.
// Scan-converted triangles of a planar triangulation only ever overlap in the first or last row of
// a span (neighbours evaluate the same edge function, so spans abut exactly; what overlaps there
// is comes from edges meeting at a vertex and float rounding): "the later triangle wins"
// (atomicMax) is needed for those two rows only, the rows in between belong to this triangle alone
// and take a plain store.  Checked on the CPU restatement
// (tests/test_oracle_pin.py::test_raster_overlaps_stay_on_span_borders): 0 of 61 M covered pixels
// over 3 000 random and 37 pipeline triangulations had a second cover strictly inside a span.
// The argument needs the float edge evaluation a*u + b to be off by less than one row; its error grows
// with the magnitude of the operands, so the plain stores are only used while the image is small
// enough for that (`exact`: W, H <= 2048, where |a*u|, |b| < 2^23 keep the rounding error of the sum
// below 1/4 row -- checked at 2048 x 2048 with steep hull edges by the same test); larger images
// take atomicMax on every row.
struct SpanPlane {
  float au, b, c;        // a * (float)u, b, c of this image's plane
  unsigned hi;           // (i + 1) << (dbits + 1) | valid << dbits
  int dmin, dmax, bias;  // clamp range of d_plane, bias
};
__device__ __forceinline__ unsigned span_entry(const SpanPlane& p, int v) {
  const int d = (int)(p.au + p.b * (float)v + p.c);            // elas.cpp:722, same operation order
  return p.hi | (unsigned)(min(max(d, p.dmin), p.dmax) + p.bias);
}
__device__ __forceinline__ void write_span(unsigned* __restrict__ map, int W, int u, int vlo, int vhi,
                                           const SpanPlane& p, bool exact) {
  if (vhi <= vlo) return;
  if (!exact) {
    for (int v = vlo; v < vhi; v++) atomicMax(&map[v * W + u], span_entry(p, v));
    return;
  }
  atomicMax(&map[vlo * W + u], span_entry(p, vlo));
  if (vhi - 1 > vlo) atomicMax(&map[(vhi - 1) * W + u], span_entry(p, vhi - 1));
  for (int v = vlo + 1; v < vhi - 1; v++) map[v * W + u] = span_entry(p, v);
}

#ifndef JN_RASTER_LANES
#define JN_RASTER_LANES 16
#endif
constexpr int RASTER_LANES = JN_RASTER_LANES;

__global__ void raster_kernel(Geo g, Workspace ws) {
  const int side = blockIdx.y, frame = blockIdx.z;
  FrameInfo* info = ws.info + frame;
  if (info->status != JN_OK) return;
  const int nt = info->n_tri[side];
  const int W = g.W, H = g.H;
  const int dbits = g.pm_dbits;
  if (nt + 1 >= (1 << (31 - dbits))) return;      // does not fit the map entry: rejected by the guard kernel
  const int4* sup = reinterpret_cast<const int4*>(ws.sup) + (size_t)frame * g.cap_s;
  const int* tri = ws.tri[side] + (size_t)frame * g.cap_t * 3;
  const float* planes = ws.planes[side] + (size_t)frame * g.cap_t * 6;
  unsigned* map = reinterpret_cast<unsigned*>(ws.trimap[side]) + (size_t)frame * W * H;
  // RASTER_LANES lanes per triangle: support points sit on a 5-px lattice, so most triangles are 5 to
  // 10 columns wide and a full warp per triangle would leave three quarters of its lanes idle
  const int lane = threadIdx.x & (RASTER_LANES - 1);
  const bool exact = W <= 2048 && H <= 2048;
  const int po = side ? 3 : 0;              // this image's plane, the other image's slope
  const int grp = (blockIdx.x * blockDim.x + threadIdx.x) / RASTER_LANES;
  const int ngrps = (gridDim.x * blockDim.x) / RASTER_LANES;
  for (int i = grp; i < nt; i += ngrps) {
    float tu[3], tv[3];
#pragma unroll
    for (int k = 0; k < 3; k++) {
      int4 s = sup[tri[3 * i + k]];
      tu[k] = side ? (float)(s.x - s.z) : (float)s.x;
      tv[k] = (float)s.y;
    }
    const float* pl = planes + 6 * i;
    const float pa = pl[po], pd = pl[3 - po];
    SpanPlane sp;
    sp.b = pl[po + 1];
    sp.c = pl[po + 2];
    // (double)|x| < 0.7  <=>  |x| <= 0.7f  (0.7f is the largest float below 0.7), elas.cpp:872
    const unsigned valid = (fabsf(pa) <= 0.7f && fabsf(pd) <= 0.7f) ? 1u : 0u;
    sp.hi = ((unsigned)(i + 1) << (dbits + 1)) | (valid << dbits);
    sp.bias = g.plane_radius + 1;
    sp.dmin = -sp.bias;
    sp.dmax = g.p.disp_max + sp.bias;
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
      for (int u = max((int)Au, 0) + lane; u < min((int)Bu, W); u += RASTER_LANES) {
        int v1 = f2u_lo32(ACa * (float)u + ACb), v2 = f2u_lo32(ABa * (float)u + ABb);
        int vlo = max(min(v1, v2), 0), vhi = min(max(v1, v2), H);
        sp.au = pa * (float)u;
        write_span(map, W, u, vlo, vhi, sp, exact);
      }
    if ((int)Bu != (int)Cu)
      for (int u = max((int)Bu, 0) + lane; u < min((int)Cu, W); u += RASTER_LANES) {
        int v1 = f2u_lo32(ACa * (float)u + ACb), v2 = f2u_lo32(BCa * (float)u + BCb);
        int vlo = max(min(v1, v2), 0), vhi = min(max(v1, v2), H);
        sp.au = pa * (float)u;
        write_span(map, W, u, vlo, vhi, sp, exact);
      }
  }
}

// A frame whose triangle count does not fit the index field of a map entry is rejected (status
// JN_ERR_UNSUPPORTED, outputs untouched) before the raster kernel runs.  At 1920x1200 the field
// holds 2^21 triangles against at most 184 336.
__global__ void raster_guard_kernel(Geo g, Workspace ws, int B) {
  const int frame = blockIdx.x * blockDim.x + threadIdx.x;
  if (frame >= B) return;
  FrameInfo* info = ws.info + frame;
  if (info->status != JN_OK) return;
  const int lim = 1 << (31 - g.pm_dbits);
  if (info->n_tri[0] + 1 >= lim || info->n_tri[1] + 1 >= lim) info->status = JN_ERR_UNSUPPORTED;
}

// Tile shape: columns per CTA x rows per thread.  The cell's candidate words are reloaded whenever the
// thread crosses into the next grid row (10 rows = half a grid row of both presets).  Swept on the B200:
// 64..512 columns x 10..40 rows are within 5 %, 128 x 10 is the best.
#ifndef JN_DENSE_THREADS
#define JN_DENSE_THREADS 128
#endif
#ifndef JN_DENSE_ROWS
#define JN_DENSE_ROWS 10
#endif
constexpr int DENSE_THREADS = JN_DENSE_THREADS;
constexpr int DENSE_ROWS = JN_DENSE_ROWS;         // image rows per thread (amortises the per-thread setup)
constexpr unsigned KEY_NONE = 0xFFFFFFFFu;
constexpr unsigned KEY_BIAS = 2048u << 13;        // keeps cost + prior non-negative
constexpr unsigned KEY_PLANE = 1u << 12;          // class bit: plane-range candidates come second

// Candidate keys:  (cost + 2048) << 13 | class << 12 | d      class 0 = grid candidate, 1 = plane range.
// The reference keeps the FIRST minimum in its evaluation order (grid candidates outside the
// plane range ascending, then the plane range ascending; strict '<', H7) = the smallest key.

__device__ __forceinline__ unsigned shl_clamp(unsigned x, unsigned s) {   // shifts >= 32 give 0
  unsigned r;
  asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(s));
  return r;
}
// bits of one set word (disparities wbase ..) that are outside [lo, hi]; an empty range removes nothing
__device__ __forceinline__ unsigned outside_range(unsigned bits, int wbase, int lo, int hi) {
  const unsigned ge_lo = shl_clamp(0xFFFFFFFFu, (unsigned)max(lo - wbase, 0));
  const unsigned gt_hi = shl_clamp(0xFFFFFFFFu, (unsigned)max(hi + 1 - wbase, 0));
  return bits & ~(ge_lo & ~gt_hi);
}

// base + idx * (+-16 bytes): one IMAD.WIDE
template <int DIR>
__device__ __forceinline__ const uint4* step16(const uint4* base, int idx) {
  const uint4* r;
  asm("mad.wide.s32 %0, %1, %2, %3;" : "=l"(r) : "r"(idx), "n"(16 * DIR), "l"(base));
  return r;
}
// index of the highest set bit (FLO)
__device__ __forceinline__ int top_bit(unsigned x) {
  int r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}

// checked evaluation of one candidate (image bounds): Bf = descriptors of the searched image
__device__ __forceinline__ void eval_candidate(unsigned& best, const uint4& a, const uint4* __restrict__ Bf,
                                               unsigned rowoff, int u, int dir, int d, unsigned addend,
                                               unsigned wm4) {
  const int uw = u + dir * d;
  if ((unsigned)(uw - 2) < wm4) {   // u_warp in [2, W-3]
    const uint4 b = __ldg(Bf + (rowoff + (unsigned)uw));
    best = min(best, sad16(a, b, 0u) * 8192u + (addend + (unsigned)d));
  }
}
__device__ __forceinline__ void eval_bits_checked(unsigned& best, unsigned bits, int wbase, const uint4& a,
                                                  const uint4* __restrict__ Bf, unsigned rowoff, int u, int dir,
                                                  unsigned wm4) {
  while (bits) {
    const int b = top_bit(bits);
    bits ^= 1u << b;
    eval_candidate(best, a, Bf, rowoff, u, dir, wbase + b, KEY_BIAS, wm4);
  }
}

// R = plane radius known at compile time (2: ROBOTICS, 3: MIDDLEBURY), 0 = run-time radius.
// SUB = subsampling: only pixels with even u and v are matched (elas.cpp:877-896) and the result
// goes to (u/2, v/2) of the W/2 x H/2 map (elas.cpp:693); everything else is unchanged.
// DIR = -1: left image searches the right one at u - d; +1: right image searches the left one at u + d
// (compile-time so that candidate addresses are one IMAD.WIDE with an immediate).
#ifndef JN_DENSE_MINB
#define JN_DENSE_MINB (2048 / JN_DENSE_THREADS)   // 32 registers: all 64 warps of an SM resident
#endif

// Frame-level base addresses of a CTA (frame and side are the same for all its threads), kept in
// SHARED memory and re-read where they are needed (LDS.64 + one IMAD.WIDE per address).  Held in
// registers they do not fit next to the candidate loop at 32 registers per thread, and ptxas then
// rebuilds every address from the kernel parameters and the CTA / thread ids at every use: ~50 of
// the ~270 instructions per pixel of this issue-bound kernel were such address arithmetic
// (2.75 -> 2.52 ms per 64 frames).  Measured on top of this and NOT kept (profiles/r02_dense_variants.txt;
// ptxas is at the edge of the 32-register budget here and answers small changes with spills): an outer
// loop over grid rows instead of the per-row "cell changed?" test (+3.5 %), retiring the four border
// columns before the row loop (+14 %), 20 rows per CTA (+0.5 %), 42 registers at 12 CTAs per SM (+17 %).
enum { DB_PMAP = 0, DB_A, DB_B, DB_OUT, DB_CELLS, DB_MASKS, DB_COUNT };
// sb = shared-space address of the table (a link-time constant: the load is LDS.64 [imm]); volatile: every
// use re-reads it instead of keeping twelve registers alive across the row loop
__device__ __forceinline__ unsigned long long dense_base(unsigned sb, int k) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(sb + 8u * (unsigned)k));
  return v;
}
__device__ __forceinline__ void store_global_f32(unsigned long long addr, float x) {
  asm volatile("st.global.f32 [%0], %1;" ::"l"(addr), "f"(x) : "memory");
}

template <int R, bool SUB, int DIR>
__device__ __forceinline__ void dense_body(const Geo& g, const Workspace& ws, const unsigned sb) {
  constexpr int side = DIR > 0 ? 1 : 0;
  const int W = g.W, H = g.H;
  const int um = blockIdx.x * DENSE_THREADS + threadIdx.x;     // map column
  const unsigned am = __ballot_sync(0xffffffffu, um < g.Wd);   // lanes that stay
  if (um >= g.Wd) return;
  int u = SUB ? 2 * um : um;
  asm volatile("" : "+r"(u));          // one register, not three instructions (S2R, S2R, IMAD) at every use
  const unsigned gx = __umulhi((unsigned)u, g.gs_magic);
  constexpr int dir = DIR;
  const int dmax = g.p.disp_max;
  const unsigned wm4 = (unsigned)(W - 4);
  const bool u_ok = u >= 2 && u < W - 2;
  const int r = R ? R : g.plane_radius;
  const int dbits = g.pm_dbits;
  const unsigned dmask = (1u << dbits) - 1u;
  // no candidate of this column can leave the image: u_warp = u + dir*d in [2, W-3] for all d in [0, disp_max]
  const bool col_free = side ? (u >= 2 && u + dmax <= W - 3) : (u - dmax >= 2 && u <= W - 3);
  const bool warp_free = R != 0 && __all_sync(am, col_free);
  // key addends of the plane range per |k|: (2048 + prior) << 13 | class bit, with and without the prior
  const unsigned addv0 = ((unsigned)(2048 + g.P[0]) << 13) | KEY_PLANE;
  const unsigned addv1 = ((unsigned)(2048 + g.P[1]) << 13) | KEY_PLANE;
  const unsigned addv2 = ((unsigned)(2048 + g.P[2]) << 13) | KEY_PLANE;
  const unsigned addv3 = ((unsigned)(2048 + g.P[3]) << 13) | KEY_PLANE;
  const unsigned addn = KEY_BIAS | KEY_PLANE;
  const int vm_end = min((int)(blockIdx.y + 1) * DENSE_ROWS, g.Hd);
  int vm = blockIdx.y * DENSE_ROWS;
  uint4 cw = make_uint4(0u, 0u, 0u, 0u), cm = make_uint4(0u, 0u, 0u, 0u);
  unsigned ci = 0u, gy_cur = 0xFFFFFFFFu;
  int nwords = 0;
#pragma unroll 1
  for (; vm < vm_end; vm++) {
      const int v = SUB ? 2 * vm : vm;
      const unsigned e = __ldg(reinterpret_cast<const unsigned*>(dense_base(sb, DB_PMAP)) + ((unsigned)(v * W) + (unsigned)u));
      // descriptor rows exist for v in [2, H-3] (the border rows read the clamped row: their result is discarded)
      const unsigned pix = (unsigned)(max(min(v, H - 3), 2) * W) + (unsigned)u;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(dense_base(sb, DB_A)) + pix);
      // rows inside one grid row share the cell: (re)load its candidate words when the grid row changes
      const unsigned gy = __umulhi((unsigned)v, g.gs_magic);
      if (gy != gy_cur) {
        gy_cur = gy;
        ci = gy * (unsigned)g.gw + gx;
        const uint4* cells = reinterpret_cast<const uint4*>(dense_base(sb, DB_CELLS));
        cw = __ldg(cells + 2u * ci);
        cm = __ldg(cells + 2u * ci + 1u);
        nwords = (int)cm.y;
      }
      const bool act = e != 0u && u_ok && (int)texture16(a) >= g.p.match_texture;
      const int d_plane = (int)(e & dmask) - (r + 1);
      const bool valid = (e >> dbits) & 1u;
      const bool inside = d_plane - r >= 0 && d_plane + r <= dmax;       // whole plane range in [0, disp_max]
      const bool quick = warp_free && __all_sync(am, inside || !act);
      float out = -10.f;
      if (act) {
        unsigned best = KEY_NONE;
        if (R != 0 && quick) {
          // ---- no per-candidate test anywhere: candidate d of this pixel is descriptor pix + dir * d
          const int lo = d_plane - R, hi = d_plane + R;
          unsigned bestg = KEY_NONE;
          if (nwords != GRID_OVERFLOW) {
#pragma unroll
            for (int j = 0; j < GRID_WORDS; j++) {
              if (j < nwords) {
                const int wbase = 32 * (int)((cm.x >> (8 * j)) & 0xFFu);
                unsigned bits =
                    outside_range(j == 0 ? cw.x : (j == 1 ? cw.y : (j == 2 ? cw.z : cw.w)), wbase, lo, hi);
                if (bits) {
                  const uint4* __restrict__ Bw =
                      reinterpret_cast<const uint4*>(dense_base(sb, DB_B)) + (pix + (unsigned)(dir * wbase));
                  unsigned bw = KEY_NONE;
                  do {                                   // lowest set bit first: the minimum does not depend on the order
                    const unsigned t = bits - 1u;
                    const int b = top_bit(bits ^ t);
                    bits &= t;
                    const uint4 c = __ldg(step16<DIR>(Bw, b));
                    bw = min(bw, sad16(a, c, 0u) * 8192u + (unsigned)b);
                  } while (bits);
                  bestg = min(bestg, bw + (unsigned)wbase);
                }
              }
            }
          } else {
            const uint32_t* cell = reinterpret_cast<const uint32_t*>(dense_base(sb, DB_MASKS)) + ci * (unsigned)g.gwords;
            const uint4* __restrict__ Brow = reinterpret_cast<const uint4*>(dense_base(sb, DB_B)) + pix;
            for (int w = 0; w < g.gwords; w++) {
              unsigned bits = outside_range(__ldg(cell + w), 32 * w, lo, hi);
              while (bits) {
                const int b = top_bit(bits);
                bits ^= 1u << b;
                const int d = 32 * w + b;
                const uint4 c = __ldg(step16<DIR>(Brow, d));
                bestg = min(bestg, sad16(a, c, 0u) * 8192u + (unsigned)d);
              }
            }
          }
          if (bestg != KEY_NONE) best = bestg + KEY_BIAS;
          // the 2R+1 plane-range descriptors are consecutive: one address, immediate offsets, all loads
          // issued before the first SAD (after the grid walk: 20 registers that must not be live during it)
          const uint4* __restrict__ Bp =
              reinterpret_cast<const uint4*>(dense_base(sb, DB_B)) + (pix + (unsigned)(dir * d_plane));
          uint4 bb[2 * R + 1];
#pragma unroll
          for (int q = -R; q <= R; q++) bb[q + R] = __ldg(Bp + q);        // disparity d_plane + dir * q
          const unsigned k0 = (valid ? addv0 : addn) + (unsigned)d_plane;
          const unsigned k1 = (valid ? addv1 : addn) + (unsigned)d_plane;
          const unsigned k2 = (valid ? addv2 : addn) + (unsigned)d_plane;
          const unsigned k3 = (valid ? addv3 : addn) + (unsigned)d_plane;
#pragma unroll
          for (int q = -R; q <= R; q++) {
            const int aq = q < 0 ? -q : q;
            const unsigned kq = (aq == 0 ? k0 : (aq == 1 ? k1 : (aq == 2 ? k2 : k3))) + (unsigned)(dir * q);
            best = min(best, sad16(a, bb[q + R], 0u) * 8192u + kq);
          }
        } else {
          // ---- checked path: image borders, plane ranges clipped by [0, disp_max], run-time radius
          const uint4* __restrict__ Bf = reinterpret_cast<const uint4*>(dense_base(sb, DB_B));
          const unsigned rowoff = pix - (unsigned)u;
          const int lo = max(d_plane - r, 0), hi = min(d_plane + r, dmax);
          if (nwords != GRID_OVERFLOW) {
#pragma unroll
            for (int j = 0; j < GRID_WORDS; j++) {
              if (j < nwords) {
                const int wbase = 32 * (int)((cm.x >> (8 * j)) & 0xFFu);
                const unsigned bits =
                    outside_range(j == 0 ? cw.x : (j == 1 ? cw.y : (j == 2 ? cw.z : cw.w)), wbase, lo, hi);
                eval_bits_checked(best, bits, wbase, a, Bf, rowoff, u, dir, wm4);
              }
            }
          } else {
            const uint32_t* cell = reinterpret_cast<const uint32_t*>(dense_base(sb, DB_MASKS)) + ci * (unsigned)g.gwords;
            for (int w = 0; w < g.gwords; w++)
              eval_bits_checked(best, outside_range(__ldg(cell + w), 32 * w, lo, hi), 32 * w, a, Bf, rowoff, u,
                                dir, wm4);
          }
          for (int d = lo; d <= hi; d++) {
            const int ad = abs(d - d_plane);
            const unsigned add = valid ? (((unsigned)(2048 + g.P[ad]) << 13) | KEY_PLANE) : addn;
            eval_candidate(best, a, Bf, rowoff, u, dir, d, add, wm4);
          }
        }
        out = (best != KEY_NONE) ? (float)(best & 0xFFFu) : -1.f;
      }
      store_global_f32(dense_base(sb, DB_OUT) + 4ull * ((unsigned)(vm * g.Wd) + (unsigned)(SUB ? um : u)), out);
  }
}

template <int R, bool SUB>
__global__ void __launch_bounds__(DENSE_THREADS, JN_DENSE_MINB)
dense_kernel(Geo g, Workspace ws) {
  __shared__ unsigned long long s_base[DB_COUNT];
  const int frame = blockIdx.z >> 1, side = blockIdx.z & 1;
  if (ws.info[frame].status != JN_OK) return;
  if (threadIdx.x == 0) {
    const size_t fpix = (size_t)frame * g.W * g.H;
    s_base[DB_PMAP] = (unsigned long long)(reinterpret_cast<const unsigned*>(ws.trimap[side]) + fpix);
    s_base[DB_A] = (unsigned long long)(ws.desc[side] + fpix * 16);
    s_base[DB_B] = (unsigned long long)(ws.desc[side ^ 1] + fpix * 16);
    s_base[DB_OUT] = (unsigned long long)(ws.Draw[side] + (SUB ? (size_t)frame * g.Wd * g.Hd : fpix));
    s_base[DB_CELLS] = (unsigned long long)(ws.gridlist[side] + (size_t)frame * g.gw * g.gh * GRID_LIST);
    s_base[DB_MASKS] = (unsigned long long)(ws.gridmask[side] + (size_t)frame * g.gw * g.gh * g.gwords);
  }
  __syncthreads();
  const unsigned sb = (unsigned)__cvta_generic_to_shared(s_base);
  if (side) dense_body<R, SUB, 1>(g, ws, sb);
  else dense_body<R, SUB, -1>(g, ws, sb);
}

}  // namespace

void launch_raster(const Geo& g, int B, Workspace& ws, cudaStream_t s) {
  size_t mbytes = (size_t)B * g.W * g.H * sizeof(int32_t);
  cudaMemsetAsync(ws.trimap[0], 0, mbytes, s);   // 0 = no triangle
  cudaMemsetAsync(ws.trimap[1], 0, mbytes, s);
  raster_guard_kernel<<<(B + 127) / 128, 128, 0, s>>>(g, ws, B);
  raster_kernel<<<dim3(64, 2, B), 256, 0, s>>>(g, ws);
  g_jn_launches += 2;
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
