// support.cu -- support-point matching, filtering and compaction.
//
// Replaces computeMatchingDisparity (elas.cpp:269-373), computeSupportMatches
// (elas.cpp:375-443), removeInconsistentSupportPoints (elas.cpp:153-179) and
// removeRedundantSupportPoints (elas.cpp:181-235).
//
// support_match_kernel: one CTA per candidate row.  The four descriptor rows a
// candidate row touches (v-2 and v+2 of both images, W*16 bytes each) are
// fetched once with 1-D bulk async copies (TMA unit) into shared memory; a warp
// owns one candidate, its lanes stride the disparities, every disparity costs
// four conflict-free 128-bit shared loads and 16 VABSDIFF4.U8.ACC.  The
// (best, second best) pair is order independent (second best = second smallest
// energy of the multiset, best = lowest disparity among the minima, H7), so it
// is reduced with warp shuffles.
//
// The in-place, scan-ordered inconsistency filter (H3) is computed exactly by
// a monotone frontier propagation: a point's final validity only depends on
// how many EARLIER similar neighbours were invalidated, so starting from the
// points whose original support count is too low and decrementing the counters
// of their later neighbours reaches the unique fixed point = the sequential
// result.  The two redundancy passes carry dependencies along one axis only.
#include "common.cuh"
#include "blockutil.cuh"

namespace {

constexpr int MATCH_THREADS = 512;
constexpr unsigned EMPTY_KEY = (32767u << 16) | 0xFFFFu;

struct Best {
  unsigned key;  // (energy << 16) | disparity of the best match
  unsigned e2;   // second smallest energy
};

__device__ __forceinline__ void best_update(Best& b, unsigned e, unsigned d) {
  unsigned k = (e << 16) | d;
  if (k < b.key) {
    b.e2 = b.key >> 16;
    b.key = k;
  } else if (e < b.e2) {
    b.e2 = e;
  }
}

__device__ __forceinline__ Best best_merge_warp(Best b) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    unsigned ok = __shfl_xor_sync(0xffffffffu, b.key, off);
    unsigned oe = __shfl_xor_sync(0xffffffffu, b.e2, off);
    if (ok < b.key) {
      b.e2 = min(b.key >> 16, oe);
      b.key = ok;
    } else {
      b.e2 = min(ok >> 16, b.e2);
    }
  }
  return b;
}

// One candidate, executed by a full warp.  rowA_* = descriptor rows of the image
// the pixel lives in, rowB_* = rows of the image searched; dir = -1 (left pixel,
// search u-d) or +1 (right pixel, search u+d).  Returns the disparity or -1.
__device__ __forceinline__ int match_candidate(const Geo& g, int u, int dir, const uint4* rowA_t,
                                               const uint4* rowA_b, const uint4* rowB_t, const uint4* rowB_b,
                                               const uint4* centre_row, int lane) {
  const int W = g.W;
  if (u < 5 || u > W - 6) return -1;
  uint4 c = __ldg(centre_row + u);
  if ((int)texture16(c) < g.p.support_texture) return -1;
  int dmin = max(g.p.disp_min, 0);
  int dmax = (dir < 0) ? min(g.p.disp_max, u - 5) : min(g.p.disp_max, W - u - 5);
  if (dmax - dmin < 10) return -1;
  const uint4 a0 = rowA_t[u - 2], a1 = rowA_t[u + 2], a2 = rowA_b[u - 2], a3 = rowA_b[u + 2];
  Best b;
  b.key = EMPTY_KEY;
  b.e2 = 32767u;
  for (int d = dmin + lane; d <= dmax; d += 32) {
    int uw = u + dir * d;
    unsigned e = sad16(a0, rowB_t[uw - 2], 0u);
    e = sad16(a1, rowB_t[uw + 2], e);
    e = sad16(a2, rowB_b[uw - 2], e);
    e = sad16(a3, rowB_b[uw + 2], e);
    best_update(b, e, (unsigned)d);
  }
  b = best_merge_warp(b);
  float e1 = (float)(b.key >> 16), e2 = (float)b.e2;
  if (e1 < __fmul_rn(g.p.support_threshold, e2)) return (int)(b.key & 0xFFFFu);
  return -1;
}

__global__ void __launch_bounds__(MATCH_THREADS, 1)
support_match_kernel(Geo g, const uint8_t* __restrict__ desc1, const uint8_t* __restrict__ desc2,
                     int16_t* __restrict__ dcan) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int W = g.W, H = g.H, Wc = g.Wc;
  const int vc = blockIdx.x, frame = blockIdx.y;
  const int step = g.p.candidate_stepsize;
  const int v = vc * step;
  int16_t* out = dcan + (size_t)frame * g.Wc * g.Hc + (size_t)vc * Wc;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = MATCH_THREADS / 32;

  // rows 0 / column 0 of the candidate image are never matched and stay 0 (calloc, H3)
  const bool row_ok = (vc >= 1) && v >= 5 && v <= H - 6;
  if (!row_ok) {
    for (int uc = tid; uc < Wc; uc += MATCH_THREADS) out[uc] = (vc == 0 || uc == 0) ? 0 : -1;
    return;
  }

  const size_t rowbytes = (size_t)W * 16;
  uint4* L_t = reinterpret_cast<uint4*>(smem);
  uint4* L_b = reinterpret_cast<uint4*>(smem + rowbytes);
  uint4* R_t = reinterpret_cast<uint4*>(smem + 2 * rowbytes);
  uint4* R_b = reinterpret_cast<uint4*>(smem + 3 * rowbytes);
  int* fwd = reinterpret_cast<int*>(smem + 4 * rowbytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 4 * rowbytes + (size_t)((Wc * 4 + 15) & ~15));

  const uint8_t* d1 = desc1 + (size_t)frame * W * H * 16;
  const uint8_t* d2 = desc2 + (size_t)frame * W * H * 16;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_arrive_expect_tx(bar, (uint32_t)(4 * rowbytes));
    bulk_g2s(L_t, d1 + (size_t)(v - 2) * rowbytes, (uint32_t)rowbytes, bar);
    bulk_g2s(L_b, d1 + (size_t)(v + 2) * rowbytes, (uint32_t)rowbytes, bar);
    bulk_g2s(R_t, d2 + (size_t)(v - 2) * rowbytes, (uint32_t)rowbytes, bar);
    bulk_g2s(R_b, d2 + (size_t)(v + 2) * rowbytes, (uint32_t)rowbytes, bar);
  }
  mbar_wait(bar, 0);

  const uint4* c1 = reinterpret_cast<const uint4*>(d1 + (size_t)v * rowbytes);
  const uint4* c2 = reinterpret_cast<const uint4*>(d2 + (size_t)v * rowbytes);

  // forward: left pixel (u,v) -> right image
  for (int uc = 1 + warp; uc < Wc; uc += nwarps) {
    int d = match_candidate(g, uc * step, -1, L_t, L_b, R_t, R_b, c1, lane);
    if (lane == 0) fwd[uc] = d;
  }
  __syncthreads();
  // backward: right pixel (u-d,v) -> left image, then the cross check (elas.cpp:404-411)
  for (int uc = 1 + warp; uc < Wc; uc += nwarps) {
    int d = fwd[uc];
    int res = -1;
    if (d >= 0) {
      int d2 = match_candidate(g, uc * step - d, +1, R_t, R_b, L_t, L_b, c2, lane);
      if (d2 >= 0 && abs(d - d2) <= g.p.lr_threshold) res = d;
    }
    if (lane == 0) out[uc] = (int16_t)res;
  }
  if (tid == 0) out[0] = 0;
}

// c0(p) = number of lattice points q in the (2r+1)^2 window (p included) that are
// valid and within incon_threshold of p, on the ORIGINAL candidate image.
__global__ void incon_count_kernel(Geo g, const int16_t* __restrict__ dcan, int32_t* __restrict__ cnt) {
  const int Wc = g.Wc, Hc = g.Hc, r = g.p.incon_window_size, thr = g.p.incon_threshold;
  const int frame = blockIdx.z;
  const int u = blockIdx.x * blockDim.x + threadIdx.x, v = blockIdx.y * blockDim.y + threadIdx.y;
  if (u >= Wc || v >= Hc) return;
  const int16_t* dc = dcan + (size_t)frame * Wc * Hc;
  int d = dc[v * Wc + u];
  int c = 0;
  if (d >= 0) {
    int v0 = max(v - r, 0), v1 = min(v + r, Hc - 1), u0 = max(u - r, 0), u1 = min(u + r, Wc - 1);
    for (int v2 = v0; v2 <= v1; v2++)
      for (int u2 = u0; u2 <= u1; u2++) {
        int d2 = dc[v2 * Wc + u2];
        c += (d2 >= 0 && abs(d - d2) <= thr) ? 1 : 0;
      }
  }
  cnt[(size_t)frame * Wc * Hc + v * Wc + u] = c;
}

constexpr int FILT_THREADS = 1024;

__global__ void __launch_bounds__(FILT_THREADS, 1)
support_filter_kernel(Geo g, const int16_t* __restrict__ dcan_all, int32_t* __restrict__ cnt_all,
                      int32_t* __restrict__ frontier_all, int16_t* __restrict__ incon_all,
                      int16_t* __restrict__ final_all, int32_t* __restrict__ sup_all, int32_t* __restrict__ px0,
                      int32_t* __restrict__ px1, int32_t* __restrict__ py_all, FrameInfo* __restrict__ info) {
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

  // ---- inconsistent points: frontier propagation -----------------------------
  if (tid == 0) { s_n[0] = 0; s_n[1] = 0; }
  __syncthreads();
  for (int p = tid; p < NP; p += T)
    if (dc[p] >= 0 && cnt[p] < minsup) fr[0][atomicAdd(&s_n[0], 1)] = p;
  __syncthreads();
  int cur = 0, rounds = 0;
  const int later = r + r * (2 * r + 1);  // cells after p in scan order (u outer, v inner)
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
    st2[p] = o;
  }
  __syncthreads();

  // ---- redundant points, vertical pass (dependencies along a column only) ------
  const int md = 5, rt = 1;  // redun_max_dist, redun_threshold (elas.cpp:421-422)
  for (int u = tid; u < Wc; u += T) {
    for (int v = 0; v < Hc; v++) {
      int d = st2[v * Wc + u];
      if (d < 0) continue;
      bool up = false, down = false;
      for (int j = 1; j <= md && v - j >= 0; j++) {
        int d2 = st2[(v - j) * Wc + u];
        if (d2 >= 0 && abs(d - d2) <= rt) { up = true; break; }
      }
      if (!up) continue;
      for (int j = 1; j <= md && v + j < Hc; j++) {
        int d2 = st2[(v + j) * Wc + u];
        if (d2 >= 0 && abs(d - d2) <= rt) { down = true; break; }
      }
      if (down) st2[v * Wc + u] = -1;
    }
  }
  __syncthreads();
  // ---- horizontal pass (dependencies along a row only) --------------------------
  for (int v = tid; v < Hc; v += T) {
    for (int u = 0; u < Wc; u++) {
      int d = st2[v * Wc + u];
      if (d < 0) continue;
      bool lft = false, rgt = false;
      for (int j = 1; j <= md && u - j >= 0; j++) {
        int d2 = st2[v * Wc + u - j];
        if (d2 >= 0 && abs(d - d2) <= rt) { lft = true; break; }
      }
      if (!lft) continue;
      for (int j = 1; j <= md && u + j < Wc; j++) {
        int d2 = st2[v * Wc + u + j];
        if (d2 >= 0 && abs(d - d2) <= rt) { rgt = true; break; }
      }
      if (rgt) st2[v * Wc + u] = -1;
    }
  }
  __syncthreads();

  // ---- compaction in u-major order, lattice row/column 0 excluded (elas.cpp:426-431)
  for (int u = tid; u < Wc; u += T) {
    int c = 0;
    if (u >= 1)
      for (int v = 1; v < Hc; v++) c += st2[v * Wc + u] >= 0;
    s_col[u] = c;
  }
  __syncthreads();
  int total = block_exclusive_scan(s_col, Wc, s_part);
  const int step = g.p.candidate_stepsize;
  int4* sup = reinterpret_cast<int4*>(sup_all) + (size_t)frame * g.cap_s;
  int32_t* x0 = px0 + (size_t)frame * g.cap_s;
  int32_t* x1 = px1 + (size_t)frame * g.cap_s;
  int32_t* y = py_all + (size_t)frame * g.cap_s;
  for (int u = 1 + tid; u < Wc; u += T) {
    int k = s_col[u];
    for (int v = 1; v < Hc; v++) {
      int d = st2[v * Wc + u];
      if (d >= 0) {
        sup[k] = make_int4(u * step, v * step, d, 0);
        x0[k] = u * step;
        x1[k] = u * step - d;
        y[k] = v * step;
        k++;
      }
    }
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
  size_t smem = (size_t)4 * g.W * 16 + ((g.Wc * 4 + 15) & ~15) + 16;
  if (smem > 227 * 1024 || g.Wc > 2048) {
    jn_set_error("image width %d too large for the shared-memory support matcher", g.W);
    return JN_ERR_UNSUPPORTED;
  }
  static bool attr_set = false;
  if (!attr_set) {
    JN_CUDA_CHECK(cudaFuncSetAttribute(support_match_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
    attr_set = true;
  }
  support_match_kernel<<<dim3(g.Hc, B), MATCH_THREADS, smem, s>>>(g, ws.desc[0], ws.desc[1], ws.dcan);
  dim3 cb(32, 8), cg((g.Wc + 31) / 32, (g.Hc + 7) / 8, B);
  incon_count_kernel<<<cg, cb, 0, s>>>(g, ws.dcan, ws.cnt);
  support_filter_kernel<<<B, FILT_THREADS, 0, s>>>(g, ws.dcan, ws.cnt, ws.frontier, ws.dcan_incon, ws.dcan_final,
                                                   ws.sup, ws.px[0], ws.px[1], ws.py, ws.info);
  g_jn_launches += 3;
  return JN_OK;
}
