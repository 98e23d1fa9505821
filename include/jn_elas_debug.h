/*
 * jn_elas_debug.h -- stage-dump entry point of the CUDA path (used by the
 * parity tests and profiling only; not part of the reference-facing API).
 *
 * jn_elas_stages runs the same device pipeline as jn_elas_process on one frame
 * and copies every intermediate back to the host so each kernel can be
 * compared with the oracle's dump of the matching reference stage
 * (elas.cpp:57-140).
 */
#ifndef JN_ELAS_DEBUG_H
#define JN_ELAS_DEBUG_H
#include "jn_elas.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Every pointer may be NULL.  Sizes as in the reference: W,H image; Wc,Hc
 * candidate lattice (elas.cpp:386-387); gw,gh grid (elas.cpp:90-91). */
typedef struct jn_stage_dump {
  uint8_t* desc1;        /* H*W*16 */
  uint8_t* desc2;
  int16_t* dcan_raw;     /* Hc*Wc */
  int16_t* dcan_incon;
  int16_t* dcan_final;
  int32_t* support;      /* (u,v,d) triples */
  int32_t  cap_support;
  int32_t  n_support;    /* out */
  int32_t* tri1;         /* (c1,c2,c3) */
  float*   planes1;      /* 6 floats per triangle */
  int32_t* tri2;
  float*   planes2;
  int32_t  cap_tri;
  int32_t  n_tri1;       /* out */
  int32_t  n_tri2;       /* out */
  int32_t* grid1;        /* gh*gw*(disp_max+2), reference list format */
  int32_t* grid2;
  float*   D1_raw;
  float*   D2_raw;
  float*   D1_lr;
  float*   D2_lr;
  float*   D1_seg;
  float*   D2_seg;
  float*   D1_gap;
  float*   D2_gap;
  float*   D1_mean;
  float*   D2_mean;
  float*   D1;
  float*   D2;
  int64_t  dense_evals;  /* unused by the CUDA path */
  int64_t  dense_pixels;
} jn_stage_dump;

int jn_elas_stages(jn_elas* e, const uint8_t* I1, const uint8_t* I2, const int32_t dims[3],
                   jn_stage_dump* out);

/* Single stages with injected inputs (randomised parity tests): the support filter + compaction on a
 * caller-supplied candidate image, the Delaunay kernel on caller-supplied integer points, and the
 * post-processing chain on caller-supplied raw disparity maps. */
int jn_debug_support_filter(jn_elas* e, const int16_t* dcan, const int32_t dims[3], int16_t* out_incon,
                            int16_t* out_final, int32_t* support, int32_t cap_support, int32_t* n_support,
                            int32_t* rounds);
int jn_debug_triangulate(jn_elas* e, const int32_t* xy, int n, const int32_t dims[3], int32_t* tri,
                         int32_t cap_tri, int32_t* n_tri);
int jn_debug_postprocess(jn_elas* e, const float* D1raw, const float* D2raw, const int32_t dims[3],
                         jn_stage_dump* out);

/* Per-stage device timing of a batch call, CUDA events on the launching stream.
 * Stage order: descriptor, support (match+filter), delaunay, planes+grid, raster,
 * dense match, post-processing. */
#define JN_PROFILE_STAGES 7
int jn_elas_profile(jn_elas* e, int enable);
int jn_elas_profile_read(jn_elas* e, float ms[JN_PROFILE_STAGES]);

/* Diagnostics: copies the device-side per-frame record (counts, status, phase timestamps). */
int jn_elas_frameinfo(jn_elas* e, int frame, void* out, int bytes);

/* Test hook: point-count limits of the Delaunay kernel's shared-memory paths (-1 = default);
 * 0,0 forces the occupancy-grid ranking and the global-memory triangle tables. */
void jn_debug_delaunay_limits(int sort_max, int smem_max);

/* Test hook: grid cells with more than `limit` candidates are decoded from the bit set by the
 * dense matcher instead of the compact list (-1 = default 16; 0 = always the bit set). */
void jn_debug_grid_list_limit(int limit);

/* 1 if this scan object uses the table-division / float-filtered fast path of scan_kernel (Q has
 * stereoRectify's sparsity and the table division was verified against the IEEE division for this
 * image size when the object was created), 0 if it evaluates the general expressions. */
int jn_scan_fast_path(const jn_scan* s);

/* Test hook: 0 = objects created from now on never take the fast path, 1 = they may, -1 = default
 * (environment variable JN_SCAN_FAST=0 disables it). */
void jn_debug_scan_fast(int enable);

/* Number of kernel launches issued by this library since load (bench.py's gpu_launches). */
long long jn_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif
