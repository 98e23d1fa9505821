// common.cuh -- shared declarations of the sm_100a stereo-to-obstacle path.
//
// Data layout in HBM (one "frame slot" per frame of a batch; every kernel takes
// the batch as its outermost grid dimension so small stages still fill 148 SMs):
//   desc[2]   H*W*16 u8    16-byte descriptor per pixel, row-major, pixel-major
//                          (same addressing as Descriptor::I_desc, descriptor.cpp:84)
//   dcan      Hc*Wc  i16   support candidates on the 5-px lattice (elas.cpp:388)
//   sup_*     cap_s  i32   compacted support points (u,v,d), u-major order
//   tri tables per side    3 neighbour handles + 3 vertex ids per row, 2*n rows
//   tri_out / planes       compacted triangles (c1,c2,c3) + 6 plane floats
//   gridmask  [2][gh*gw][GW] u32  per-cell disparity bit sets (createGrid)
//   trimap    [2][H*W] u32  plane map: last triangle covering each pixel + its plane prior there
//                          ((tri + 1) << (pm_dbits + 1) | valid << pm_dbits | d_plane + radius + 1; 0 = none)
//   D*        Hd*Wd f32    disparity maps / post-processing ping-pong buffers (Hd x Wd = H x W, or
//                          H/2 x W/2 with subsampling; frame slots are Hd*Wd floats apart)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/jn_elas.h"

#define JN_CUDA_CHECK(x)                                                              \
  do {                                                                                \
    cudaError_t e_ = (x);                                                             \
    if (e_ != cudaSuccess) {                                                          \
      jn_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
      return JN_ERR_CUDA;                                                             \
    }                                                                                 \
  } while (0)

void jn_set_error(const char* fmt, ...);
#include <atomic>
extern std::atomic<long long> g_jn_launches;   // kernels launched by this library (bench.py gpu_launches)

// Right-image x coordinates (u - d) are stored with this bias so that they stay non-negative
// (corner support points reach u - d = -d); the Delaunay predicates are translation invariant.
constexpr int JN_XBIAS = 4096;
constexpr int GRID_LIST = 16;    // u16 units per cell of the compact candidate form (= 32 bytes)
constexpr int GRID_WORDS = 4;    // non-zero 32-bit words of a cell's disparity set kept in the compact form
constexpr int GRID_OVERFLOW = 255;   // word count marker: decode the full bit set instead

// Per-call geometry + parameters, passed to kernels by value.
struct Geo {
  int W, H, bpl;        // image size, input stride in bytes
  int Wd, Hd;           // disparity map size: W x H, or W/2 x H/2 with subsampling (elas.cpp:914-917)
  int speckle_eff;      // removeSmallSegments size limit at the map's resolution (elas.cpp:986-991)
  int gap_eff;          // gapInterpolation width at the map's resolution (elas.cpp:1106-1111)
  int Wc, Hc;           // candidate lattice (elas.cpp:386-387)
  int gw, gh;           // disparity grid (elas.cpp:90-91)
  int gwords;           // u32 words per grid cell bit set, ceil((disp_max+1)/32) rounded up to 4
  unsigned gs_magic;    // ceil(2^32 / grid_size): x / grid_size == __umulhi(x, gs_magic) for x < 2^16
  int cap_s;            // support point capacity per frame
  int cap_t;            // triangle-table rows per side (2*cap_s)
  int plane_radius;     // elas.cpp:806
  int P[8];             // prior table entries 0..plane_radius (elas.cpp:802-805)
  int grid_list_limit;  // cells with more non-zero set words than this use the bit-set path (<= GRID_WORDS)
  int pm_dbits;         // plane map entry: [tri index + 1 | plane valid | d_plane + plane_radius + 1 (pm_dbits bits)]
  int dl_sort_max, dl_smem_max;   // Delaunay: point-count limits of the shared-memory paths
  jn_elas_params p;
};

// Per-frame counters / status written on the device (no host sync on the hot path).
struct FrameInfo {
  int n_support;
  int n_tri[2];
  int status;           // JN_OK or JN_FEW_SUPPORT
  int incon_rounds;     // diagnostics
  int dmerge_depth;     // diagnostics: levels of the D&C tree
  int pad[2];
  long long dt[2][6];   // diagnostics: globaltimer at the Delaunay phase boundaries, per side
};

// All per-batch device buffers.  Index [f] = frame slot.
struct Workspace {
  int B;                       // frame slots allocated
  uint8_t* desc[2];            // B * H*W*16
  int16_t* dcan;               // B * Hc*Wc   (after matching)
  int16_t* dcan_incon;         // B * Hc*Wc   (after inconsistent filter)
  int16_t* dcan_final;         // B * Hc*Wc
  int32_t* cnt;                // B * Hc*Wc   support counters of the inconsistent filter
  int32_t* frontier;           // B * 2 * Hc*Wc
  int32_t* sup;                // B * cap_s * 4  (u,v,d,pad)
  int32_t* px[2];              // B * cap_s   x coordinate per side (left u, right u-d)
  int32_t* py;                 // B * cap_s
  int32_t* occ;                // B * occ_cells   occupancy grid for ranking (right side)
  int32_t* xlist[2];           // B * cap_s   per side
  int32_t* ylist[2];
  int32_t* tmpA[2];            // B * cap_s   scratch (ranks, flags, scans)
  int32_t* tmpB[2];
  int32_t* tmpC[2];
  int32_t* tmpD[2];
  int32_t* nb[2];              // B * cap_t*3
  int32_t* vx[2];              // B * cap_t*3
  int32_t* nodeL[2];           // B * node_cap  far-left handle per D&C node
  int32_t* nodeR[2];
  int32_t* tri[2];             // B * cap_t*3   compacted triangles
  float*   planes[2];          // B * cap_t*6
  uint32_t* gridtmp[2];        // B * gh*gw*gwords
  uint32_t* gridmask[2];
  uint16_t* gridlist[2];       // B * gh*gw*32 bytes  non-zero words of each cell's set (grid_words_kernel)
  int32_t* trimap[2];          // B * H*W
  float* Draw[2];              // B * H*W
  float* Dlr[2];
  float* Dtmp[2];              // post-processing scratch
  float* Dtmp2[2];
  int32_t* label;              // B * H*W  (CCL)
  int32_t* segsize;            // B * H*W
  FrameInfo* info;             // B
  size_t bytes;
};

// ---- small device helpers -------------------------------------------------

// 4-byte SAD with accumulate: one VABSDIFF4.U8.ACC on sm_100a.
__device__ __forceinline__ unsigned sad4(unsigned a, unsigned b, unsigned acc) {
  unsigned r;
  asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(acc));
  return r;
}
__device__ __forceinline__ unsigned sad16(const uint4& a, const uint4& b, unsigned acc) {
  acc = sad4(a.x, b.x, acc);
  acc = sad4(a.y, b.y, acc);
  acc = sad4(a.z, b.z, acc);
  acc = sad4(a.w, b.w, acc);
  return acc;
}
// sum_i |byte_i - 128|   (texture test, elas.cpp:301-305, 715-719)
__device__ __forceinline__ unsigned texture16(const uint4& a) {
  const uint4 c = make_uint4(0x80808080u, 0x80808080u, 0x80808080u, 0x80808080u);
  return sad16(a, c, 0u);
}

// stage kernels (one translation unit each)
void launch_descriptor(const Geo& g, int B, const uint8_t* I1, const uint8_t* I2, Workspace& ws, cudaStream_t s);
int  launch_support(const Geo& g, int B, Workspace& ws, cudaStream_t s);
int  launch_support_match(const Geo& g, int B, Workspace& ws, cudaStream_t s);
int  launch_support_filter(const Geo& g, int B, Workspace& ws, cudaStream_t s);
int  launch_delaunay(const Geo& g, int B, Workspace& ws, cudaStream_t s);
void launch_planes_grid(const Geo& g, int B, Workspace& ws, cudaStream_t s);
void launch_dense(const Geo& g, int B, Workspace& ws, cudaStream_t s);
void launch_raster(const Geo& g, int B, Workspace& ws, cudaStream_t s);
void launch_dense_match(const Geo& g, int B, Workspace& ws, cudaStream_t s);
void launch_post(const Geo& g, int B, Workspace& ws, float* D1out, float* D2out, int32_t* status, cudaStream_t s);
