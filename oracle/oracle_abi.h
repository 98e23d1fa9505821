/*
 * oracle_abi.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Shared C interface of the two CPU checkers:
 *   oracle/_ref/libelas_ref.so   the UNMODIFIED reference sources
 *                                (/root/reference/src/elas, five .cpp files) compiled in
 *                                place by oracle/Makefile, plus ref_shim.cpp
 *   oracle/libelas_port.so       the plain-C restatement (elas_port.c,
 *                                delaunay_port.c, scan_port.c)
 * Both export the same symbols with prefix ref_ / port_ so the parity tests
 * can swap them.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load these libraries; the product
 * (jackal-navigation_b200/) never does.
 */
#ifndef ORACLE_ABI_H
#define ORACLE_ABI_H

#include <stdint.h>
#include "../include/jn_elas.h"   /* jn_elas_params only */

#ifdef __cplusplus
extern "C" {
#endif

/* Every pointer may be NULL (stage not wanted).  Sizes: W,H image; Wc,Hc
 * candidate lattice; gw,gh grid; cap_* capacities given by the caller. */
typedef struct oracle_stages {
  uint8_t* desc1;        /* H*W*16 */
  uint8_t* desc2;
  int16_t* dcan_raw;     /* Hc*Wc after matching + L/R, before filtering */
  int16_t* dcan_incon;   /* after removeInconsistentSupportPoints */
  int16_t* dcan_final;   /* after both removeRedundantSupportPoints passes */
  int32_t* support;      /* (u,v,d) triples */
  int32_t  cap_support;
  int32_t  n_support;    /* out */
  int32_t* tri1;         /* (c1,c2,c3) triples, left image triangulation */
  float*   planes1;      /* (t1a,t1b,t1c,t2a,t2b,t2c) per triangle */
  int32_t* tri2;         /* right image triangulation */
  float*   planes2;
  int32_t  cap_tri;
  int32_t  n_tri1;       /* out */
  int32_t  n_tri2;       /* out */
  int32_t* grid1;        /* gh*gw*(disp_max+2) */
  int32_t* grid2;
  float*   D1_raw;       /* after computeDisparity */
  float*   D2_raw;
  float*   D1_lr;        /* after leftRightConsistencyCheck */
  float*   D2_lr;
  float*   D1_seg;       /* after removeSmallSegments */
  float*   D2_seg;
  float*   D1_gap;       /* after gapInterpolation */
  float*   D2_gap;
  float*   D1_mean;      /* after adaptiveMean (copy of gap stage if filter off) */
  float*   D2_mean;
  float*   D1;           /* final */
  float*   D2;
  int64_t  dense_evals;  /* out: SAD evaluations in computeDisparity (port only; 0 from ref) */
  int64_t  dense_pixels; /* out: findMatch calls that reached the candidate loop (port only) */
} oracle_stages;

/* scan side (point_cloud.cpp restatement; exists only in the port) */
typedef struct oracle_scan_meta {
  double angle_min, angle_max, range_min, range_max;
  int32_t n_finite, n_points;
} oracle_scan_meta;

#ifdef __cplusplus
}
#endif
#endif
