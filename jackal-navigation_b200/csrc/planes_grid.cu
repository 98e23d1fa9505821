// planes_grid.cu -- disparity planes per triangle and the per-cell disparity sets.
//
// computeDisparityPlanes (elas.cpp:507-577) solves two 3x3 systems per triangle
// with Matrix::solve (matrix.cpp:414-502): Gauss-Jordan, FULL pivoting, double,
// result narrowed to float.  d_plane = (int)(a*u+b*v+c) truncates, and at the
// support points the exact value is an integer, so the arithmetic has to be
// reproduced operation by operation: this translation unit is compiled with
// -fmad=false (x86-64 SSE has no fused multiply-add) and keeps the reference's
// operation order, including the ">= big" pivot rule that picks the LAST maximum.
//
// createGrid (elas.cpp:579-659) becomes a bit set per 20x20-px cell
// (gwords x u32, bit d = disparity d is a candidate): support points mark
// d-1..d+1, then a flat 3x3 OR over cell indices gw+1 .. gw*gh-gw-2 -- flat, so
// it wraps across grid rows exactly as the reference's nine running pointers do
// (SURVEY H5).  Ascending bit order = the reference's sorted candidate list.
#include "common.cuh"

namespace {

__device__ bool solve3(double A[3][3], double B[3]) {
  int ipiv[3] = {0, 0, 0};
  for (int i = 0; i < 3; i++) {
    double big = 0.0;
    int irow = 0, icol = 0;
    for (int j = 0; j < 3; j++)
      if (ipiv[j] != 1)
        for (int k = 0; k < 3; k++)
          if (ipiv[k] == 0)
            if (fabs(A[j][k]) >= big) { big = fabs(A[j][k]); irow = j; icol = k; }
    ++ipiv[icol];
    if (irow != icol) {
      for (int l = 0; l < 3; l++) { double t = A[irow][l]; A[irow][l] = A[icol][l]; A[icol][l] = t; }
      double t = B[irow]; B[irow] = B[icol]; B[icol] = t;
    }
    if (fabs(A[icol][icol]) < 1e-20) return false;
    double pivinv = 1.0 / A[icol][icol];
    A[icol][icol] = 1.0;
    for (int l = 0; l < 3; l++) A[icol][l] *= pivinv;
    B[icol] *= pivinv;
    for (int ll = 0; ll < 3; ll++)
      if (ll != icol) {
        double dum = A[ll][icol];
        A[ll][icol] = 0.0;
        for (int l = 0; l < 3; l++) A[ll][l] -= A[icol][l] * dum;
        B[ll] -= B[icol] * dum;
      }
  }
  return true;
}

__global__ void planes_kernel(Geo g, Workspace ws) {
  const int side = blockIdx.y, frame = blockIdx.z;
  const FrameInfo* info = ws.info + frame;
  if (info->status != JN_OK) return;
  const int nt = info->n_tri[side];
  const int4* sup = reinterpret_cast<const int4*>(ws.sup) + (size_t)frame * g.cap_s;
  const int* tri = ws.tri[side] + (size_t)frame * g.cap_t * 3;
  float* planes = ws.planes[side] + (size_t)frame * g.cap_t * 6;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < nt; t += gridDim.x * blockDim.x) {
    int4 s[3];
    for (int k = 0; k < 3; k++) s[k] = sup[tri[3 * t + k]];
    for (int rs = 0; rs < 2; rs++) {
      double A[3][3], Bv[3];
      for (int k = 0; k < 3; k++) {
        A[k][0] = rs ? (double)(s[k].x - s[k].z) : (double)s[k].x;
        A[k][1] = (double)s[k].y;
        A[k][2] = 1.0;
        Bv[k] = (double)s[k].z;
      }
      float* o = planes + 6 * t + 3 * rs;
      if (solve3(A, Bv)) { o[0] = (float)Bv[0]; o[1] = (float)Bv[1]; o[2] = (float)Bv[2]; }
      else { o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; }
    }
  }
}

__global__ void grid_scatter_kernel(Geo g, Workspace ws) {
  const int side = blockIdx.y, frame = blockIdx.z;
  const FrameInfo* info = ws.info + frame;
  if (info->status != JN_OK) return;
  const int n = info->n_support;
  const int4* sup = reinterpret_cast<const int4*>(ws.sup) + (size_t)frame * g.cap_s;
  uint32_t* tmp = ws.gridtmp[side] + (size_t)frame * g.gw * g.gh * g.gwords;
  const int gs = g.p.grid_size;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int4 s = sup[i];
    int xc = side ? (s.x - s.z) : s.x;
    if (xc < 0) continue;                 // only possible with corner points
    int x = xc / gs, y = s.y / gs;
    if (x >= g.gw || y >= g.gh) continue;
    int lo = max(s.z - 1, 0), hi = min(s.z + 1, g.p.disp_max);
    uint32_t* cell = tmp + (size_t)(y * g.gw + x) * g.gwords;
    for (int d = lo; d <= hi; d++) atomicOr(cell + (d >> 5), 1u << (d & 31));
  }
}

__global__ void grid_dilate_kernel(Geo g, Workspace ws) {
  const int side = blockIdx.y, frame = blockIdx.z;
  const int cells = g.gw * g.gh, gwords = g.gwords, gw = g.gw;
  const uint32_t* tmp = ws.gridtmp[side] + (size_t)frame * cells * gwords;
  uint32_t* out = ws.gridmask[side] + (size_t)frame * cells * gwords;
  const bool ok = ws.info[frame].status == JN_OK;
  const long first = gw + 1, last = (long)cells - gw - 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells * gwords; i += gridDim.x * blockDim.x) {
    int c = i / gwords, w = i - c * gwords;
    uint32_t r = 0;
    if (ok && g.gh >= 3 && c >= first && c <= last) {
      const uint32_t* p = tmp + (size_t)c * gwords + w;
      const long o = (long)gw * gwords;
      r = p[-o - gwords] | p[-o] | p[-o + gwords] | p[-gwords] | p[0] | p[gwords] | p[o - gwords] | p[o] |
          p[o + gwords];
    }
    out[i] = r;
  }
}

// Compact form of a cell's bit set for the dense matcher: the (at most GRID_WORDS) non-zero 32-bit
// words of the set.  32 bytes per cell: uint4 {bits of word 0..3}, uint4 {word indices packed as
// bytes, count, 0, 0}.  A support point marks d-1..d+1 and the dilation ORs 9 cells, so the set is a
// few short runs: one or two words for a cell on one surface, more across a depth edge.  A cell with
// more non-zero words (or word indices that do not fit a byte) stores count = GRID_OVERFLOW and is
// decoded from the full bit set.
__global__ void grid_words_kernel(Geo g, Workspace ws) {
  const int side = blockIdx.y, frame = blockIdx.z;
  const int cells = g.gw * g.gh, gwords = g.gwords;
  const uint32_t* mask = ws.gridmask[side] + (size_t)frame * cells * gwords;
  uint4* out = reinterpret_cast<uint4*>(ws.gridlist[side] + (size_t)frame * cells * GRID_LIST);
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += gridDim.x * blockDim.x) {
    uint32_t bits[GRID_WORDS] = {0u, 0u, 0u, 0u};
    uint32_t wpack = 0;
    int n = 0;
    for (int w = 0; w < gwords; w++) {
      const uint32_t b = mask[(size_t)c * gwords + w];
      if (b) {
        if (n < GRID_WORDS) { bits[n] = b; wpack |= (uint32_t)(w & 0xFF) << (8 * n); }
        n++;
      }
    }
    if (n > g.grid_list_limit || gwords > 256) n = GRID_OVERFLOW;
    out[2 * c] = make_uint4(bits[0], bits[1], bits[2], bits[3]);
    out[2 * c + 1] = make_uint4(wpack, (uint32_t)n, 0u, 0u);
  }
}

}  // namespace

// test hook: cells with more than `limit` non-zero set words take the bit-set path of the dense matcher
static int g_list_limit = GRID_WORDS;
extern "C" void jn_debug_grid_list_limit(int limit) {
  g_list_limit = (limit < 0 || limit > GRID_WORDS) ? GRID_WORDS : limit;
}

void launch_planes_grid(const Geo& g_in, int B, Workspace& ws, cudaStream_t s) {
  Geo g = g_in;
  g.grid_list_limit = g_list_limit;
  size_t gbytes = (size_t)B * g.gw * g.gh * g.gwords * sizeof(uint32_t);
  cudaMemsetAsync(ws.gridtmp[0], 0, gbytes, s);
  cudaMemsetAsync(ws.gridtmp[1], 0, gbytes, s);
  planes_kernel<<<dim3(64, 2, B), 128, 0, s>>>(g, ws);
  grid_scatter_kernel<<<dim3(16, 2, B), 256, 0, s>>>(g, ws);
  int cw = g.gw * g.gh * g.gwords;
  grid_dilate_kernel<<<dim3((cw + 255) / 256, 2, B), 256, 0, s>>>(g, ws);
  grid_words_kernel<<<dim3((g.gw * g.gh + 127) / 128, 2, B), 128, 0, s>>>(g, ws);
  g_jn_launches += 4;
}
