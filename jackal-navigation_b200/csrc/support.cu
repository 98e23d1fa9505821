// support.cu -- support-point matching, filtering and compaction.
//
// Replaces computeMatchingDisparity (elas.cpp:269-373), computeSupportMatches
// (elas.cpp:375-443), removeInconsistentSupportPoints (elas.cpp:153-179) and
// removeRedundantSupportPoints (elas.cpp:181-235).
//
// support_match_kernel: a CTA matches the candidates of (part of) a candidate row.  The
// descriptor rows they touch (v-2 and v+2 of both images; only the columns those candidates
// can reach) are fetched with 1-D bulk async copies (TMA unit) into shared memory; at 1920 px
// two half-row CTAs share an SM.  A warp owns four neighbouring candidates, its lanes stride
// the POSITIONS of the searched rows: a lane loads the four searched descriptors of its
// position once (conflict-free 128-bit shared loads) and scores them against all four
// candidates, 16 VABSDIFF4.U8.ACC each; the four equally long ranges are tiled without masked
// lanes (wrapped tiling, see match_group).  The (best, second best) pair is order independent
// (second best = second smallest energy of the multiset, best = lowest disparity among the
// minima, H7), so it is kept as two packed keys per lane and reduced over the warp with two
// REDUX.MIN.  The kernel is bound by the integer ALU pipe (~60 % of the VABSDIFF4 issue ceiling).
//
// The in-place, scan-ordered inconsistency filter (H3) is computed exactly by
// a monotone frontier propagation: a point's final validity only depends on
// how many EARLIER similar neighbours were invalidated, so starting from the
// points whose original support count is too low and decrementing the counters
// of their later neighbours reaches the unique fixed point = the sequential
// result; it runs on byte arrays in shared memory.  The two redundancy passes carry
// dependencies along one axis only: the side that looks at unprocessed entries is evaluated for
// all points in parallel, the serial walk keeps the last five decided entries in registers.
#include "common.cuh"
#include "blockutil.cuh"

namespace {

#ifndef JN_MATCH_THREADS
#define JN_MATCH_THREADS 256
#endif
constexpr int MATCH_THREADS = JN_MATCH_THREADS;
constexpr int KG = 4;   // candidates matched together by one warp
constexpr int KEY_EMPTY = 0x7fffffff;

// (best, second best) as two packed keys (energy << 16 | disparity): the smallest key is the
// reference's best match (lowest energy, lowest disparity among ties, H7) and the energy of the
// second smallest key is its second-best energy.  Updating needs only min/max.  Inside the
// search the keys are kept SHIFTED by a per-candidate constant (energy * 65536 + dir * position
// instead of + disparity; disparity = dir * (position - u)), which preserves their order, is the
// same for all KG candidates of a lane and is undone once after the search; shifted keys can be
// slightly negative, hence signed.
struct Best2 {
  int k1, k2;
};
__device__ __forceinline__ void best2_update(Best2& b, int key) {
  const int hi = max(b.k1, key);
  b.k1 = min(b.k1, key);
  b.k2 = min(b.k2, hi);
}
// Warp-wide (smallest, second smallest) with two REDUX.MIN: keys of one candidate are distinct
// across lanes and iterations (one per position), so the second smallest is the minimum over the
// lanes of "my smallest, or my second smallest if mine is the global one".
__device__ __forceinline__ Best2 best2_merge_warp(Best2 b) {
  Best2 r;
  r.k1 = __reduce_min_sync(0xffffffffu, b.k1);
  r.k2 = __reduce_min_sync(0xffffffffu, (b.k1 == r.k1) ? b.k2 : b.k1);
  return r;
}

// Gates of computeMatchingDisparity that do not depend on the search (elas.cpp:283-307): window
// inside the image, texture of the centre descriptor, at least 11 disparities to look at.
template <int DIR>
__device__ __forceinline__ bool candidate_gate(const Geo& g, int u, const uint4* __restrict__ centre_row) {
  if (u < 5 || u > g.W - 6) return false;
  const int dmin = max(g.p.disp_min, 0);
  const int dmax = (DIR < 0) ? min(g.p.disp_max, u - 5) : min(g.p.disp_max, g.W - u - 5);
  if (dmax - dmin < 10) return false;
  return (int)texture16(__ldg(centre_row + u)) >= g.p.support_texture;
}

// KG candidates, executed by a full warp.  rowA_* = descriptor rows of the image the pixels
// live in, rowB_* = rows of the image searched; DIR = -1 (left pixels, search u-d) or +1
// (right pixels, search u+d); bit k of okmask = candidate k passed candidate_gate (the others
// only need a readable u).  The lanes stride POSITIONS of the searched rows: every lane loads
// the four searched descriptors of its position once and scores them against all KG candidates
// (disparity = distance to the candidate), so the shared-memory traffic per candidate drops by
// KG.  32-position chunks that lie inside every candidate's range take a straight-line path whose
// integer-pipe work is the 16 VABSDIFF4 and 3 min/max per candidate and position (the keys are
// formed with one IMAD each on the FMA pipe).  res[k] = disparity or -1.
template <int DIR>
__device__ __forceinline__ void match_group(const Geo& g, const int (&u)[KG], unsigned okmask, const uint4* rowA_t,
                                            const uint4* rowA_b, const uint4* rowB_t, const uint4* rowB_b,
                                            int lane, int safe_u, int (&res)[KG]) {
  const int W = g.W;
  const int dmin = max(g.p.disp_min, 0);
#pragma unroll
  for (int k = 0; k < KG; k++) res[k] = -1;
  if (okmask == 0u) return;
  uint4 a[KG][4];
  int p0[KG], p1[KG];     // position range of candidate k (empty if not ok)
  Best2 best[KG];
  int plo = 0x7fffffff, phi = -0x7fffffff;   // union of the ranges
  int ilo = -0x7fffffff, ihi = 0x7fffffff;   // intersection of the ranges
#pragma unroll
  for (int k = 0; k < KG; k++) {
    const bool ok = (okmask >> k) & 1u;
    const int uk = ok ? u[k] : safe_u;        // any staged column: the scores are discarded
    a[k][0] = rowA_t[uk - 2]; a[k][1] = rowA_t[uk + 2];
    a[k][2] = rowA_b[uk - 2]; a[k][3] = rowA_b[uk + 2];
    best[k].k1 = KEY_EMPTY;
    best[k].k2 = KEY_EMPTY;
    const int dmax = (DIR < 0) ? min(g.p.disp_max, uk - 5) : min(g.p.disp_max, W - uk - 5);
    p0[k] = ok ? ((DIR < 0) ? uk - dmax : uk + dmin) : 1;
    p1[k] = ok ? ((DIR < 0) ? uk - dmin : uk + dmax) : 0;
    if (ok) {
      plo = min(plo, p0[k]); phi = max(phi, p1[k]);
      ilo = max(ilo, p0[k]); ihi = min(ihi, p1[k]);
    }
  }
  const bool allok = okmask == (1u << KG) - 1u;
  // Chunks of 32 positions from plo.  Chunks [j_lo, j_hi] lie inside every candidate's range and
  // take the straight-line path; the few before and after them (range borders) are masked per
  // candidate and lane.
  const int j_end = (phi - plo) >> 5;
  int j_lo = 0, j_hi = -1;
  // Wrapped tiling (the common case: four live candidates away from the image border, so all ranges
  // have the same length Lr, a multiple of 32): candidate k scores position P of the shared grid if
  // P >= p0[k] and position P + Lr otherwise -- both lie in its range, and over P = plo .. plo+Lr-1
  // every position of every range is visited exactly once.  Only the chunks below the largest
  // range start need per-candidate loads; no lane is ever masked and Lr/32 chunks replace the
  // Lr/32 + 1 (two of them masked) of the union tiling.
  const int Lr = p1[0] - p0[0] + 1;
  bool wrap = allok && Lr >= 32 && (Lr & 31) == 0;
#pragma unroll
  for (int k = 1; k < KG; k++) wrap = wrap && (p1[k] - p0[k] + 1 == Lr);
  int jw = 0;
  if (wrap) {
    jw = min(Lr >> 5, (ilo - plo + 31) >> 5);
    j_lo = jw;
    j_hi = (Lr >> 5) - 1;
  } else if (allok && ihi - ilo >= 31) {
    j_lo = (ilo - plo + 31) >> 5;
    j_hi = (ihi - 31 - plo) >> 5;          // >= j_lo - 1; empty if the intersection holds no whole chunk
  }
#pragma unroll 1
  for (int part = 0; part < 2; part++) {
    const bool mid = j_hi >= j_lo;
    if (wrap) {
      if (part == 0) {
#pragma unroll 1
        for (int j = 0; j < jw; j++) {
          const int P = plo + 32 * j + lane;
#pragma unroll
          for (int k = 0; k < KG; k++) {
            const int pk = P + ((P < p0[k]) ? Lr : 0);
            const uint4 s0 = rowB_t[pk - 2], s1 = rowB_t[pk + 2], s2 = rowB_b[pk - 2], s3 = rowB_b[pk + 2];
            unsigned e = sad16(a[k][0], s0, 0u);
            e = sad16(a[k][1], s1, e);
            e = sad16(a[k][2], s2, e);
            e = sad16(a[k][3], s3, e);
            best2_update(best[k], (int)e * 65536 + DIR * pk);
          }
        }
      }
    } else {
    // border chunks: [0, j_lo) before the middle, (j_hi, j_end] after it (everything if there is no middle)
    const int jb = part ? (mid ? j_hi + 1 : 0) : 0, je = part ? j_end : (mid ? j_lo - 1 : -1);
#pragma unroll 1
    for (int j = jb; j <= je; j++) {
      const int base = plo + 32 * j;
      const int p = min(base + lane, phi);
      const uint4 s0 = rowB_t[p - 2], s1 = rowB_t[p + 2], s2 = rowB_b[p - 2], s3 = rowB_b[p + 2];
      const int shift = DIR * p;
      const bool lane_in = base + lane <= phi;
#pragma unroll
      for (int k = 0; k < KG; k++) {
        unsigned e = sad16(a[k][0], s0, 0u);
        e = sad16(a[k][1], s1, e);
        e = sad16(a[k][2], s2, e);
        e = sad16(a[k][3], s3, e);
        // p0 > p1 for a candidate that is not ok: never in range
        best2_update(best[k], (lane_in && p >= p0[k] && p <= p1[k]) ? (int)e * 65536 + shift : KEY_EMPTY);
      }
    }
    }
    if (part == 0 && mid) {
      // middle: every candidate live at every lane -> 4 loads, 64 SADs, 4 keys, 12 min/max per chunk
      const uint4* qt = rowB_t + (plo + 32 * j_lo + lane);
      const uint4* qb = rowB_b + (plo + 32 * j_lo + lane);
      int shift = DIR * (plo + 32 * j_lo + lane);
#pragma unroll 1
      for (int j = j_lo; j <= j_hi; j++) {
        const uint4 s0 = qt[-2], s1 = qt[2], s2 = qb[-2], s3 = qb[2];
#pragma unroll
        for (int k = 0; k < KG; k++) {
          unsigned e = sad16(a[k][0], s0, 0u);
          e = sad16(a[k][1], s1, e);
          e = sad16(a[k][2], s2, e);
          e = sad16(a[k][3], s3, e);
          best2_update(best[k], (int)e * 65536 + shift);
        }
        qt += 32;
        qb += 32;
        shift += DIR * 32;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < KG; k++) {
    if (!((okmask >> k) & 1u)) continue;
    const Best2 b = best2_merge_warp(best[k]);
    // undo the shift: key = e * 65536 + DIR * (p - u)
    const int t1 = b.k1 - DIR * u[k];
    const float e1 = (float)(t1 >> 16);
    const float e2 = (b.k2 == KEY_EMPTY) ? 65535.f : (float)((b.k2 - DIR * u[k]) >> 16);
    if (e1 < __fmul_rn(g.p.support_threshold, e2)) res[k] = t1 & 0xFFFF;
  }
}

// A CTA matches the candidates of PART of a candidate row: groups [g_lo, g_hi) of KG lattice neighbours.
// It stages only the columns those candidates can touch -- their own columns +-2 and disp_max columns to
// either side of the left-image rows (searched by the backward pass), disp_max columns to the left of the
// right-image rows -- so that at 1920 px two half-row CTAs (91 KB each, 256 threads x 128 registers) share
// an SM: while one waits for its bulk copy, runs its gates or compacts its list, the other one keeps the
// integer pipe busy.  `nsplit` = parts per row, chosen by the launcher as the smallest that lets two CTAs
// co-reside.  Row pointers are biased by the first staged column, so all indexing is by image column.
__global__ void __launch_bounds__(MATCH_THREADS, 2)
support_match_kernel(Geo g, const uint8_t* __restrict__ desc1, const uint8_t* __restrict__ desc2,
                     int16_t* __restrict__ dcan, int nsplit, int groups_per_part, int lcols, int rcols) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_nvalid;
  const int W = g.W, H = g.H, Wc = g.Wc;
  const int vc = blockIdx.x / nsplit, part = blockIdx.x - vc * nsplit, frame = blockIdx.y;
  const int step = g.p.candidate_stepsize;
  const int v = vc * step;
  int16_t* out = dcan + (size_t)frame * g.Wc * g.Hc + (size_t)vc * Wc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = MATCH_THREADS / 32;

  // this part's groups and candidates (candidate 0 is never matched; part 0 writes its 0)
  const int ngroups = (Wc - 1 + KG - 1) / KG;
  const int g_lo = part * groups_per_part, g_hi = min(g_lo + groups_per_part, ngroups);
  const int uc_lo = 1 + g_lo * KG, uc_hi = min(1 + g_hi * KG, Wc);       // matched candidates [uc_lo, uc_hi)
  const int uc_w0 = part == 0 ? 0 : uc_lo;                                // written candidates [uc_w0, uc_hi)
  if (uc_lo >= uc_hi && part != 0) return;

  // rows 0 / column 0 of the candidate image are never matched and stay 0 (calloc, H3)
  const bool row_ok = (vc >= 1) && v >= 5 && v <= H - 6;
  if (!row_ok || uc_lo >= uc_hi) {
    for (int uc = uc_w0 + tid; uc < uc_hi; uc += MATCH_THREADS) out[uc] = (vc == 0 || uc == 0) ? 0 : -1;
    return;
  }

  // staged column windows (inclusive), see above
  const int dm = g.p.disp_max;
  const int ua = uc_lo * step, ub = (uc_hi - 1) * step;
  const int cL0 = max(ua - dm - 2, 0), cL1 = min(ub + dm + 2, W - 1);
  const int cR0 = max(ua - dm - 2, 0), cR1 = min(ub + 2, W - 1);
  const uint32_t nL = (uint32_t)(cL1 - cL0 + 1) * 16u, nR = (uint32_t)(cR1 - cR0 + 1) * 16u;
  const size_t lbytes = (size_t)lcols * 16, rbytes = (size_t)rcols * 16;
  const size_t wcb = (size_t)((Wc * 4 + 15) & ~15);
  const uint4* L_t = reinterpret_cast<const uint4*>(smem) - cL0;
  const uint4* L_b = reinterpret_cast<const uint4*>(smem + lbytes) - cL0;
  const uint4* R_t = reinterpret_cast<const uint4*>(smem + 2 * lbytes) - cR0;
  const uint4* R_b = reinterpret_cast<const uint4*>(smem + 2 * lbytes + rbytes) - cR0;
  uint8_t* arrays = smem + 2 * lbytes + 2 * rbytes;
  int* fwd = reinterpret_cast<int*>(arrays);              // forward disparity per candidate
  int* list = reinterpret_cast<int*>(arrays + wcb);       // candidates that go to the cross check
  int* resv = reinterpret_cast<int*>(arrays + 2 * wcb);   // gate flags, then final disparity
  uint64_t* bar = reinterpret_cast<uint64_t*>(arrays + 3 * wcb);

  const size_t rowbytes = (size_t)W * 16;
  const uint8_t* d1 = desc1 + (size_t)frame * W * H * 16;
  const uint8_t* d2 = desc2 + (size_t)frame * W * H * 16;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(bar, 2u * nL + 2u * nR);
    bulk_g2s(smem, d1 + (size_t)(v - 2) * rowbytes + (size_t)cL0 * 16, nL, bar);
    bulk_g2s(smem + lbytes, d1 + (size_t)(v + 2) * rowbytes + (size_t)cL0 * 16, nL, bar);
    bulk_g2s(smem + 2 * lbytes, d2 + (size_t)(v - 2) * rowbytes + (size_t)cR0 * 16, nR, bar);
    bulk_g2s(smem + 2 * lbytes + rbytes, d2 + (size_t)(v + 2) * rowbytes + (size_t)cR0 * 16, nR, bar);
  }
  const uint4* c1 = reinterpret_cast<const uint4*>(d1 + (size_t)v * rowbytes);
  const uint4* c2 = reinterpret_cast<const uint4*>(d2 + (size_t)v * rowbytes);
  // while the rows are in flight: gates of the forward candidates (centre descriptors from L2)
  for (int uc = uc_w0 + tid; uc < uc_hi; uc += MATCH_THREADS) {
    resv[uc] = (uc >= 1 && candidate_gate<-1>(g, uc * step, c1)) ? 1 : 0;
    fwd[uc] = -1;
  }
  __syncthreads();
  mbar_wait(bar, 0);

  // forward: left pixels (u,v) -> right image, KG lattice neighbours per warp pass
  for (int gi = g_lo + warp; gi < g_hi; gi += nwarps) {
    int u[KG], r[KG];
    unsigned okmask = 0u;
#pragma unroll
    for (int k = 0; k < KG; k++) {
      const int uc = 1 + gi * KG + k;
      u[k] = uc * step;
      if (uc < Wc && resv[uc]) okmask |= 1u << k;
    }
    match_group<-1>(g, u, okmask, L_t, L_b, R_t, R_b, lane, ua, r);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < KG; k++) {
        const int uc = 1 + gi * KG + k;
        if (uc < Wc) fwd[uc] = r[k];
      }
    }
  }
  __syncthreads();
  // gates of the backward candidates (right pixel u-d), all threads; then the final-result array
  for (int uc = uc_w0 + tid; uc < uc_hi; uc += MATCH_THREADS) {
    const int d = fwd[uc];
    const bool go = d >= 0 && candidate_gate<+1>(g, uc * step - d, c2);
    if (!go) fwd[uc] = -1;
    resv[uc] = (uc == 0) ? 0 : -1;
  }
  __syncthreads();
  // compact the candidates that go on (order preserved: neighbours stay together)
  if (warp == 0) {
    int n = 0;
    for (int base = uc_lo; base < uc_hi; base += 32) {
      int uc = base + lane;
      bool f = uc < uc_hi && fwd[uc] >= 0;
      unsigned m = __ballot_sync(0xffffffffu, f);
      if (f) list[n + __popc(m & ((1u << lane) - 1))] = uc;
      n += __popc(m);
    }
    if (lane == 0) s_nvalid = n;
  }
  __syncthreads();
  // backward: right pixels (u-d,v) -> left image, then the cross check (elas.cpp:404-411)
  const int nvalid = s_nvalid;
  for (int gi = warp; gi * KG < nvalid; gi += nwarps) {
    int u[KG], r[KG], ucs[KG], df[KG];
    unsigned okmask = 0u;
#pragma unroll
    for (int k = 0; k < KG; k++) {
      const int i = gi * KG + k;
      ucs[k] = (i < nvalid) ? list[i] : -1;
      df[k] = (ucs[k] >= 0) ? fwd[ucs[k]] : 0;
      u[k] = (ucs[k] >= 0) ? ucs[k] * step - df[k] : ua;
      if (ucs[k] >= 0) okmask |= 1u << k;
    }
    match_group<+1>(g, u, okmask, R_t, R_b, L_t, L_b, lane, ua, r);
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < KG; k++)
        if (ucs[k] >= 0 && r[k] >= 0 && abs(df[k] - r[k]) <= g.p.lr_threshold) resv[ucs[k]] = df[k];
    }
  }
  __syncthreads();
  for (int uc = uc_w0 + tid; uc < uc_hi; uc += MATCH_THREADS) out[uc] = (int16_t)resv[uc];
}

// c0(p) = number of lattice points q in the (2r+1)^2 window (p included) that are
// valid and within incon_threshold of p, on the ORIGINAL candidate image.
// RC = window radius known at compile time (5: both presets; fully unrolled), 0 = run-time radius
template <int RC>
__global__ void __launch_bounds__(256) incon_count_kernel(Geo g, const int16_t* __restrict__ dcan,
                                                          int32_t* __restrict__ cnt) {
  // candidate tile + window halo in shared memory (window radius <= 16, checked in make_geo);
  // lattice points outside the image read as invalid
  __shared__ int16_t tile[(8 + 32) * (32 + 32)];
  const int Wc = g.Wc, Hc = g.Hc, r = RC ? RC : g.p.incon_window_size, thr = g.p.incon_threshold;
  const int frame = blockIdx.z;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int TWD = 32 + 2 * r, THT = 8 + 2 * r;
  const int u0 = blockIdx.x * 32 - r, v0 = blockIdx.y * 8 - r;
  const int16_t* dc = dcan + (size_t)frame * Wc * Hc;
  for (int i = tid; i < TWD * THT; i += 256) {
    const int ty = i / TWD, tx = i - ty * TWD;
    const int uu = u0 + tx, vv = v0 + ty;
    // invalid (-1) and out-of-lattice entries become a far value: "valid and within thr" is then ONE unsigned
    // comparison, (unsigned)(d2 - (d - thr)) <= 2 * thr
    const int16_t t = (uu >= 0 && uu < Wc && vv >= 0 && vv < Hc) ? dc[vv * Wc + uu] : (int16_t)-1;
    tile[i] = t >= 0 ? t : (int16_t)-30000;
  }
  __syncthreads();
  const int u = blockIdx.x * 32 + threadIdx.x, v = blockIdx.y * 8 + threadIdx.y;
  if (u >= Wc || v >= Hc) return;
  const int d = tile[(threadIdx.y + r) * TWD + threadIdx.x + r];
  int c = 0;
  if (d >= 0) {
    const int dlo = d - thr;
    const unsigned span = 2u * (unsigned)thr;
#pragma unroll
    for (int dv = 0; dv <= 2 * r; dv++) {
      const int16_t* row = tile + (threadIdx.y + dv) * TWD + threadIdx.x;
#pragma unroll
      for (int du = 0; du <= 2 * r; du++) c += ((unsigned)((int)row[du] - dlo) <= span) ? 1 : 0;
    }
  }
  cnt[(size_t)frame * Wc * Hc + v * Wc + u] = c;
}

constexpr int FILT_THREADS = 1024;

__global__ void __launch_bounds__(FILT_THREADS, 1)
support_filter_kernel(Geo g, const int16_t* __restrict__ dcan_all, int32_t* __restrict__ cnt_all,
                      int32_t* __restrict__ frontier_all, int16_t* __restrict__ incon_all,
                      int16_t* __restrict__ final_all, int32_t* __restrict__ sup_all, int32_t* __restrict__ px0,
                      int32_t* __restrict__ px1, int32_t* __restrict__ py_all, FrameInfo* __restrict__ info,
                      int use_smem) {
  __shared__ int s_n[2];
  __shared__ int s_part[FILT_THREADS + 1];
  __shared__ int s_col[2048];
  const int Wc = g.Wc, Hc = g.Hc, NP = Wc * Hc;
  const int frame = blockIdx.x, tid = threadIdx.x, T = FILT_THREADS;
  const int lane = tid & 31, warp = tid >> 5, nwarps = T / 32;
  const int16_t* dc = dcan_all + (size_t)frame * NP;
  int32_t* cnt = cnt_all + (size_t)frame * NP;
  int32_t* fr[2] = {frontier_all + (size_t)frame * 2 * NP, frontier_all + (size_t)frame * 2 * NP + NP};
  int16_t* st1 = incon_all + (size_t)frame * NP;
  int16_t* st2 = final_all + (size_t)frame * NP;
  const int r = g.p.incon_window_size, thr = g.p.incon_threshold, minsup = g.p.incon_min_support;
  // Working copy of the candidate image for the redundancy passes and the compaction: shared
  // memory when it fits (their accesses are serial, data dependent and strided), else global.
  extern __shared__ int16_t s_wk[];
  int16_t* wk = use_smem ? s_wk : st2;

  // ---- inconsistent points: frontier propagation -----------------------------
  if (tid == 0) { s_n[0] = 0; s_n[1] = 0; }
  __syncthreads();
  int cur = 0, rounds = 0;
  const int later = r + r * (2 * r + 1);  // cells after p in scan order (u outer, v inner)
  // Fast path: the whole propagation runs on shared memory.  Disparities (<= 255) and counters
  // (<= (2r+1)^2 <= 255) are bytes, NP each, in the region the redundancy passes use afterwards; a
  // counter is decremented with a 32-bit shared atomic on its word (a valid point's counter never
  // drops below 1 -- it counts the point itself and only EARLIER neighbours are ever taken away -- so
  // no borrow crosses into the next byte, and "counter >= 1" doubles as the validity flag).  A warp
  // fetches 32 frontier points with one coalesced load and walks them with shuffles.
  const bool fast = use_smem && g.p.disp_max <= 255 && (2 * r + 1) * (2 * r + 1) <= 255 && later <= 128;
  if (fast) {
    uint8_t* d8 = reinterpret_cast<uint8_t*>(s_wk);
    uint8_t* c8 = d8 + ((NP + 3) & ~3);
    unsigned* c32 = reinterpret_cast<unsigned*>(c8);
    for (int p = tid; p < NP; p += T) {
      const int d = dc[p], c = cnt[p];
      d8[p] = (uint8_t)d;
      c8[p] = (d >= 0) ? (uint8_t)c : (uint8_t)0;
      if (d >= 0 && c < minsup) fr[0][atomicAdd(&s_n[0], 1)] = p;
    }
    // this lane's cells of the "later" half window: offsets (du, dv), up to four per lane
    int cdu[4], cdv[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int c = lane + 32 * k;
      if (c < r) { cdu[k] = 0; cdv[k] = c + 1; }
      else { const int kk = c - r; cdu[k] = 1 + kk / (2 * r + 1); cdv[k] = kk % (2 * r + 1) - r; }
      if (c >= later) cdu[k] = 1 << 20;     // never inside the lattice
    }
    __syncthreads();
    while (true) {
      const int n = s_n[cur];
      if (n == 0) break;
      rounds++;
      for (int base = warp * 32; base < n; base += nwarps * 32) {
        const int mq = (base + lane < n) ? fr[cur][base + lane] : 0;
        const int mqd = d8[mq], mqu = mq % Wc, mqv = mq / Wc;
        const int m = min(32, n - base);
        for (int j = 0; j < m; j++) {
          const int qd = __shfl_sync(0xffffffffu, mqd, j), qu = __shfl_sync(0xffffffffu, mqu, j),
                    qv = __shfl_sync(0xffffffffu, mqv, j);
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const int pu = qu + cdu[k], pv = qv + cdv[k];
            if (pu < Wc && pv >= 0 && pv < Hc) {
              const int p = pv * Wc + pu;
              if (c8[p] >= 1 && abs((int)d8[p] - qd) <= thr) {
                const int sh = 8 * (p & 3);
                const unsigned old = (atomicSub(&c32[p >> 2], 1u << sh) >> sh) & 0xffu;
                if ((int)old == minsup) fr[cur ^ 1][atomicAdd(&s_n[cur ^ 1], 1)] = p;
              }
            }
          }
        }
      }
      __syncthreads();
      if (tid == 0) s_n[cur] = 0;
      cur ^= 1;
      __syncthreads();
    }
    for (int p = tid; p < NP; p += T) {
      const int d = dc[p];
      st1[p] = (d >= 0 && (int)c8[p] >= minsup) ? (int16_t)d : (int16_t)-1;
    }
    __syncthreads();                       // the byte arrays are dead: the region becomes wk
    for (int p = tid; p < NP; p += T) wk[p] = st1[p];
  } else {
  for (int p = tid; p < NP; p += T)
    if (dc[p] >= 0 && cnt[p] < minsup) fr[0][atomicAdd(&s_n[0], 1)] = p;
  __syncthreads();
  while (true) {
    int n = s_n[cur];
    if (n == 0) break;
    rounds++;
    for (int i = warp; i < n; i += nwarps) {
      int q = fr[cur][i];
      int qu = q % Wc, qv = q / Wc, qd = dc[q];
      for (int c = lane; c < later; c += 32) {
        int du, dv;
        if (c < r) { du = 0; dv = c + 1; }
        else { int k = c - r; du = 1 + k / (2 * r + 1); dv = k % (2 * r + 1) - r; }
        int pu = qu + du, pv = qv + dv;
        if (pu < Wc && pv >= 0 && pv < Hc) {
          int p = pv * Wc + pu;
          int pd = dc[p];
          if (pd >= 0 && abs(pd - qd) <= thr) {
            int old = atomicSub(&cnt[p], 1);
            if (old == minsup) fr[cur ^ 1][atomicAdd(&s_n[cur ^ 1], 1)] = p;
          }
        }
      }
    }
    __syncthreads();
    if (tid == 0) s_n[cur] = 0;
    cur ^= 1;
    __syncthreads();
  }
  for (int p = tid; p < NP; p += T) {
    int d = dc[p];
    int16_t o = (d >= 0 && cnt[p] >= minsup) ? (int16_t)d : (int16_t)-1;
    st1[p] = o;
    wk[p] = o;
  }
  }
  __syncthreads();

  // ---- redundant points (elas.cpp:181-235): a point goes if it has a similar point within md cells on
  // BOTH sides along the line; the pass runs in place, so the "before" side sees this pass's removals
  // and the "after" side does not.  The "after" test therefore only needs the state at the start of the
  // pass: it is evaluated for all points in parallel into a bit array; the
  // serial walk along each line then carries the last md decided entries in registers and touches
  // shared memory once per cell (plus the flag word) -- no data-dependent inner loops, no divergence.
  const int md = 5, rt = 1;  // redun_max_dist, redun_threshold (elas.cpp:421-422)
  // one "after" bit per lattice point, next to the working copy (shared memory) or in the frontier
  // buffer, which is free by now (global-memory mode); a warp's 32 points are one word: no atomics
  const int NP32 = (NP + 31) & ~31;
  unsigned* flagw = use_smem ? reinterpret_cast<unsigned*>(reinterpret_cast<uint8_t*>(s_wk) + (((size_t)2 * NP + 16 + 3) & ~(size_t)3))
                             : reinterpret_cast<unsigned*>(fr[0]);
  // vertical pass: lines = columns
  for (int p = tid; p < NP32; p += T) {
    const int e = p < NP ? wk[p] : -1;
    bool after = false;
    if (e >= 0) {
#pragma unroll
      for (int j = 1; j <= md; j++) {
        const int q = p + j * Wc;
        if (q < NP) {
          const int e2 = wk[q];
          after = after || (e2 >= 0 && abs(e - e2) <= rt);
        }
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, after);
    if (lane == 0) flagw[p >> 5] = m;
  }
  __syncthreads();
  for (int u = tid; u < Wc; u += T) {
    int r1 = -1, r2 = -1, r3 = -1, r4 = -1, r5 = -1;   // decided entries v-1 .. v-5
    for (int v = 0; v < Hc; v++) {
      const int p = v * Wc + u;
      const int d = wk[p];
      int cur = -1;
      if (d >= 0) {
        const bool before = (r1 >= 0 && abs(d - r1) <= rt) || (r2 >= 0 && abs(d - r2) <= rt) ||
                            (r3 >= 0 && abs(d - r3) <= rt) || (r4 >= 0 && abs(d - r4) <= rt) ||
                            (r5 >= 0 && abs(d - r5) <= rt);
        cur = d;
        if (before && ((flagw[p >> 5] >> (p & 31)) & 1u)) {
          cur = -1;
          wk[p] = (int16_t)-1;
        }
      }
      r5 = r4; r4 = r3; r3 = r2; r2 = r1; r1 = cur;
    }
  }
  __syncthreads();
  // horizontal pass: lines = rows
  for (int p = tid; p < NP32; p += T) {
    const int e = p < NP ? wk[p] : -1;
    bool after = false;
    if (e >= 0) {
      const int left_in_row = Wc - 1 - p % Wc;
#pragma unroll
      for (int j = 1; j <= md; j++)
        if (j <= left_in_row) {
          const int e2 = wk[p + j];
          after = after || (e2 >= 0 && abs(e - e2) <= rt);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, after);
    if (lane == 0) flagw[p >> 5] = m;
  }
  __syncthreads();
  for (int v = tid; v < Hc; v += T) {
    int r1 = -1, r2 = -1, r3 = -1, r4 = -1, r5 = -1;
    for (int u = 0; u < Wc; u++) {
      const int p = v * Wc + u;
      const int d = wk[p];
      int cur = -1;
      if (d >= 0) {
        const bool before = (r1 >= 0 && abs(d - r1) <= rt) || (r2 >= 0 && abs(d - r2) <= rt) ||
                            (r3 >= 0 && abs(d - r3) <= rt) || (r4 >= 0 && abs(d - r4) <= rt) ||
                            (r5 >= 0 && abs(d - r5) <= rt);
        cur = d;
        if (before && ((flagw[p >> 5] >> (p & 31)) & 1u)) {
          cur = -1;
          wk[p] = (int16_t)-1;
        }
      }
      r5 = r4; r4 = r3; r3 = r2; r2 = r1; r1 = cur;
    }
  }
  __syncthreads();

  // ---- compaction in u-major order, lattice row/column 0 excluded (elas.cpp:426-431)
  for (int u = tid; u < Wc; u += T) {
    int c = 0;
    if (u >= 1)
      for (int v = 1; v < Hc; v++) c += wk[v * Wc + u] >= 0;
    s_col[u] = c;
  }
  __syncthreads();
  int total = block_exclusive_scan(s_col, Wc, s_part);
  if (wk != st2)
    for (int p = tid; p < NP; p += T) st2[p] = wk[p];   // final candidate image, kept for the stage dump
  const int step = g.p.candidate_stepsize;
  int4* sup = reinterpret_cast<int4*>(sup_all) + (size_t)frame * g.cap_s;
  int32_t* x0 = px0 + (size_t)frame * g.cap_s;
  int32_t* x1 = px1 + (size_t)frame * g.cap_s;
  int32_t* y = py_all + (size_t)frame * g.cap_s;
  for (int u = 1 + tid; u < Wc; u += T) {
    int k = s_col[u];
    for (int v = 1; v < Hc; v++) {
      int d = wk[v * Wc + u];
      if (d >= 0) {
        sup[k] = make_int4(u * step, v * step, d, 0);
        x0[k] = u * step;
        x1[k] = u * step - d + JN_XBIAS;
        y[k] = v * step;
        k++;
      }
    }
  }
  if (g.p.add_corners) {
    // addCornerSupportPoints (elas.cpp:237-267): the four image corners take the disparity of the
    // nearest support point (first one among equal squared distances), plus two copies of the
    // right corners shifted by their disparity.  Nearest = min over (dist << 32 | index) keys.
    __shared__ unsigned long long s_best[4][FILT_THREADS / 32];
    __syncthreads();
    const int cu[4] = {0, 0, g.W - 1, g.W - 1}, cv[4] = {0, g.H - 1, 0, g.H - 1};
    unsigned long long bk[4] = {~0ull, ~0ull, ~0ull, ~0ull};
    for (int j = tid; j < total; j += T) {
      const int4 sj = sup[j];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int du = cu[c] - sj.x, dv = cv[c] - sj.y;
        const unsigned dist = (unsigned)(du * du + dv * dv);
        if (dist < 10000000u) bk[c] = min(bk[c], ((unsigned long long)dist << 32) | (unsigned)j);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; c++) {
      for (int off = 16; off > 0; off >>= 1) bk[c] = min(bk[c], __shfl_xor_sync(0xffffffffu, bk[c], off));
      if (lane == 0) s_best[c][warp] = bk[c];
    }
    __syncthreads();
    if (tid == 0) {
      int cd[4];
      for (int c = 0; c < 4; c++) {
        unsigned long long b = ~0ull;
        for (int w = 0; w < nwarps; w++) b = min(b, s_best[c][w]);
        cd[c] = (b == ~0ull) ? 0 : sup[(unsigned)(b & 0xFFFFFFFFull)].z;
      }
      const int bu[6] = {0, 0, g.W - 1, g.W - 1, g.W - 1 + cd[2], g.W - 1 + cd[3]};
      const int bv[6] = {0, g.H - 1, 0, g.H - 1, 0, g.H - 1};
      const int bd[6] = {cd[0], cd[1], cd[2], cd[3], cd[2], cd[3]};
      for (int k = 0; k < 6; k++) {
        sup[total + k] = make_int4(bu[k], bv[k], bd[k], 0);
        x0[total + k] = bu[k];
        x1[total + k] = bu[k] - bd[k] + JN_XBIAS;
        y[total + k] = bv[k];
      }
    }
    total += 6;
  }
  if (tid == 0) {
    info[frame].n_support = total;
    info[frame].status = (total < 3) ? JN_FEW_SUPPORT : JN_OK;
    info[frame].incon_rounds = rounds;
    info[frame].n_tri[0] = 0;
    info[frame].n_tri[1] = 0;
  }
}

}  // namespace

int launch_support(const Geo& g, int B, Workspace& ws, cudaStream_t s) {
  int rc = launch_support_match(g, B, ws, s);
  if (rc) return rc;
  return launch_support_filter(g, B, ws, s);
}

// Parts per candidate row and the shared memory of one part (see support_match_kernel): the smallest
// split that lets two CTAs share an SM (227 KB), else a single CTA per SM on whole rows.
static size_t match_smem(const Geo& g, int nsplit, int* groups_per_part, int* lcols, int* rcols) {
  const int step = g.p.candidate_stepsize, dm = g.p.disp_max;
  const int ngroups = (g.Wc - 1 + KG - 1) / KG;
  const int gpp = (ngroups + nsplit - 1) / nsplit;
  const int span = (gpp * KG - 1) * step;                 // first to last candidate column of a part
  const int lc = min(g.W, span + 2 * dm + 5), rc = min(g.W, span + dm + 5);
  *groups_per_part = gpp; *lcols = lc; *rcols = rc;
  return (size_t)2 * lc * 16 + (size_t)2 * rc * 16 + 3 * (size_t)((g.Wc * 4 + 15) & ~15) + 16;
}

int launch_support_match(const Geo& g, int B, Workspace& ws, cudaStream_t s) {
  int nsplit = 1, gpp = 0, lc = 0, rc = 0;
  size_t smem = 0;
  bool found = false;
  for (int ns = 1; ns <= 8 && !found; ns++) {
    smem = match_smem(g, ns, &gpp, &lc, &rc);
    if (2 * (smem + 1024) <= 227 * 1024) { nsplit = ns; found = true; }
  }
  if (!found) {
    nsplit = 1;
    smem = match_smem(g, 1, &gpp, &lc, &rc);
  }
  if (smem > 226 * 1024 || g.Wc > 2048) {   // 227 KB per CTA minus the static shared memory
    jn_set_error("image width %d too large for the shared-memory support matcher", g.W);
    return JN_ERR_UNSUPPORTED;
  }
  // per device (context), not per process: set on every launch, it is a host-side table write
  JN_CUDA_CHECK(cudaFuncSetAttribute(support_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     226 * 1024));
  JN_CUDA_CHECK(cudaFuncSetAttribute(support_match_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxShared));
  support_match_kernel<<<dim3(g.Hc * nsplit, B), MATCH_THREADS, smem, s>>>(g, ws.desc[0], ws.desc[1], ws.dcan, nsplit,
                                                                           gpp, lc, rc);
  g_jn_launches += 1;
  return JN_OK;
}

// Filtering + compaction of the candidate images in ws.dcan (also driven directly by the tests).
int launch_support_filter(const Geo& g, int B, Workspace& ws, cudaStream_t s) {
  dim3 cb(32, 8), cg((g.Wc + 31) / 32, (g.Hc + 7) / 8, B);
  if (g.p.incon_window_size == 5) incon_count_kernel<5><<<cg, cb, 0, s>>>(g, ws.dcan, ws.cnt);
  else incon_count_kernel<0><<<cg, cb, 0, s>>>(g, ws.dcan, ws.cnt);
  // + 16: the byte arrays of the frontier phase (two of NP bytes, the second one word aligned) end up to
  // three bytes past 2 * NP
  const size_t np = (size_t)g.Wc * g.Hc;
  const size_t wk_bytes = ((np * sizeof(int16_t) + 16 + 3) & ~(size_t)3) + ((np + 31) / 32) * 4;   // + the "after" bits
  const int use_smem = wk_bytes <= 212 * 1024;
  // per device (context), not per process: set on every launch, it is a host-side table write
  JN_CUDA_CHECK(cudaFuncSetAttribute(support_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     212 * 1024));
  support_filter_kernel<<<B, FILT_THREADS, use_smem ? wk_bytes : 0, s>>>(
      g, ws.dcan, ws.cnt, ws.frontier, ws.dcan_incon, ws.dcan_final, ws.sup, ws.px[0], ws.px[1], ws.py, ws.info,
      use_smem);
  g_jn_launches += 2;
  return JN_OK;
}
