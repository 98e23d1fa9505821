/*
 * elas_port.c -- TEST INFRASTRUCTURE ONLY.  Never linked into the product.
 *
 * Plain-C, scalar restatement of the reference's stereo pipeline
 * (libelas as vendored in sourishg/jackal-navigation, src/elas).  Every
 * function cites the reference lines it restates.  It is pinned against the
 * unmodified reference (oracle/_ref/libelas_ref.so) by tests/test_oracle_pin.py
 * and against the committed fixtures under tests/golden/.
 *
 * Conventions fixed here for behaviour the reference leaves undefined
 * (SURVEY.md section 7.4):
 *   H1  bytes/floats the reference reads without ever writing them
 *       (descriptor border, adaptive-mean D_tmp border) are 0.
 *   H6  float -> uint32 of the scan-converter rows is "convert to int64,
 *       keep the low 32 bits" (what x86-64 does).
 * Compile with -ffp-contract=off: the reference runs on SSE without FMA.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "oracle_abi.h"

#define IMIN(a, b) ((a) < (b) ? (a) : (b))
#define IMAX(a, b) ((a) > (b) ? (a) : (b))

/* -------------------------------------------------------------------- */
/* Descriptor  (descriptor.cpp:28-36, 42-114; filter.cpp:372-416,176-267) */

static inline uint8_t sat8(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

/* Sobel responses as the SSE code produces them for interior pixels:
 * column pass S = I(v-1)+2I(v)+I(v+1), T = I(v-1)-I(v+1)   (filter.cpp:372-405)
 * du = sat8(((S(u-1)-S(u+1))>>2)+128)                        (filter.cpp:227-267)
 * dv = sat8(((T(u-1)+2T(u)+T(u+1))>>2)+128)                  (filter.cpp:176-222) */
static void sobel_port(const uint8_t* I, int w, int h, int stride, uint8_t* du, uint8_t* dv) {
  memset(du, 0, (size_t)w * h);
  memset(dv, 0, (size_t)w * h);
  for (int v = 1; v < h - 1; v++) {
    const uint8_t* r0 = I + (size_t)(v - 1) * stride;
    const uint8_t* r1 = I + (size_t)v * stride;
    const uint8_t* r2 = I + (size_t)(v + 1) * stride;
    for (int u = 1; u < w - 1; u++) {
      int Sl = r0[u - 1] + 2 * r1[u - 1] + r2[u - 1];
      int Sr = r0[u + 1] + 2 * r1[u + 1] + r2[u + 1];
      int Tl = r0[u - 1] - r2[u - 1];
      int Tc = r0[u] - r2[u];
      int Tr = r0[u + 1] - r2[u + 1];
      du[(size_t)v * w + u] = sat8(((Sl - Sr) >> 2) + 128);
      dv[(size_t)v * w + u] = sat8(((Tl + 2 * Tc + Tr) >> 2) + 128);
    }
  }
}

/* 16 bytes per pixel for u in [3,w-4], v in [3,h-4]; everything else 0 (H1).
 * Sample pattern: descriptor.cpp:92-110. */
/* half != 0: only rows 4, 6, 8, ... are computed (descriptor.cpp:48-78), the others stay 0. */
static int descriptor_port(const uint8_t* I, int w, int h, int stride, uint8_t* desc, int half) {
  uint8_t* du = (uint8_t*)malloc((size_t)w * h);
  uint8_t* dv = (uint8_t*)malloc((size_t)w * h);
  sobel_port(I, w, h, stride, du, dv);
  memset(desc, 0, (size_t)16 * w * h);
  for (int v = half ? 4 : 3; v < h - 3; v += half ? 2 : 1) {
    const uint8_t* u0 = du + (size_t)(v - 2) * w;
    const uint8_t* u1 = du + (size_t)(v - 1) * w;
    const uint8_t* u2 = du + (size_t)v * w;
    const uint8_t* u3 = du + (size_t)(v + 1) * w;
    const uint8_t* u4 = du + (size_t)(v + 2) * w;
    const uint8_t* v1 = dv + (size_t)(v - 1) * w;
    const uint8_t* v2 = dv + (size_t)v * w;
    const uint8_t* v3 = dv + (size_t)(v + 1) * w;
    for (int u = 3; u < w - 3; u++) {
      uint8_t* o = desc + ((size_t)v * w + u) * 16;
      o[0] = u0[u];
      o[1] = u1[u - 2]; o[2] = u1[u]; o[3] = u1[u + 2];
      o[4] = u2[u - 1]; o[5] = u2[u]; o[6] = u2[u]; o[7] = u2[u + 1];
      o[8] = u3[u - 2]; o[9] = u3[u]; o[10] = u3[u + 2];
      o[11] = u4[u];
      o[12] = v1[u]; o[13] = v2[u - 1]; o[14] = v2[u + 1]; o[15] = v3[u];
    }
  }
  free(du);
  free(dv);
  return 0;
}

int port_descriptor(const uint8_t* I, int w, int h, int stride, uint8_t* desc) {
  return descriptor_port(I, w, h, stride, desc, 0);
}

static inline int sad16(const uint8_t* a, const uint8_t* b) {
  int s = 0;
  for (int i = 0; i < 16; i++) s += abs((int)a[i] - (int)b[i]);
  return s;
}

static inline int texture16(const uint8_t* a) {
  int s = 0;
  for (int i = 0; i < 16; i++) s += abs((int)a[i] - 128);
  return s;
}

/* -------------------------------------------------------------------- */
/* Support matching  (elas.cpp:269-373) */

static int match_support(const jn_elas_params* p, int w, int h, int u, int v,
                         const uint8_t* desc_l, const uint8_t* desc_r, int right_image) {
  const int us = 2, vs = 2, win = 3;
  if (!(u >= win + us && u <= w - win - 1 - us && v >= win + vs && v <= h - win - 1 - vs)) return -1;
  const uint8_t* A = right_image ? desc_r : desc_l; /* image the pixel lives in */
  const uint8_t* B = right_image ? desc_l : desc_r; /* image searched */
  if (texture16(A + ((size_t)v * w + u) * 16) < p->support_texture) return -1;
  int dmin = IMAX(p->disp_min, 0);
  int dmax = right_image ? IMIN(p->disp_max, w - u - win - us) : IMIN(p->disp_max, u - win - us);
  if (dmax - dmin < 10) return -1;
  const uint8_t* a0 = A + ((size_t)(v - vs) * w + (u - us)) * 16;
  const uint8_t* a1 = A + ((size_t)(v - vs) * w + (u + us)) * 16;
  const uint8_t* a2 = A + ((size_t)(v + vs) * w + (u - us)) * 16;
  const uint8_t* a3 = A + ((size_t)(v + vs) * w + (u + us)) * 16;
  int e1 = 32767, d1 = -1, e2 = 32767, d2 = -1;
  for (int d = dmin; d <= dmax; d++) {
    int uw = right_image ? u + d : u - d;
    int s = sad16(a0, B + ((size_t)(v - vs) * w + (uw - us)) * 16) +
            sad16(a1, B + ((size_t)(v - vs) * w + (uw + us)) * 16) +
            sad16(a2, B + ((size_t)(v + vs) * w + (uw - us)) * 16) +
            sad16(a3, B + ((size_t)(v + vs) * w + (uw + us)) * 16);
    if (s < e1) { e2 = e1; d2 = d1; e1 = s; d1 = d; }
    else if (s < e2) { e2 = s; d2 = d; }
  }
  if (d1 >= 0 && d2 >= 0 && (float)e1 < p->support_threshold * (float)e2) return d1;
  return -1;
}

/* removeInconsistentSupportPoints (elas.cpp:153-179): in place, u outer / v inner. */
static void filter_inconsistent(const jn_elas_params* p, int16_t* dc, int wc, int hc) {
  int r = p->incon_window_size;
  for (int u = 0; u < wc; u++)
    for (int v = 0; v < hc; v++) {
      int d = dc[v * wc + u];
      if (d < 0) continue;
      int support = 0;
      for (int u2 = u - r; u2 <= u + r; u2++)
        for (int v2 = v - r; v2 <= v + r; v2++)
          if (u2 >= 0 && v2 >= 0 && u2 < wc && v2 < hc) {
            int d2 = dc[v2 * wc + u2];
            if (d2 >= 0 && abs(d - d2) <= p->incon_threshold) support++;
          }
      if (support < p->incon_min_support) dc[v * wc + u] = -1;
    }
}

/* removeRedundantSupportPoints (elas.cpp:181-235): in place, u outer / v inner. */
static void filter_redundant(int16_t* dc, int wc, int hc, int maxdist, int thr, int vertical) {
  int du[2] = {0, 0}, dv[2] = {0, 0};
  if (vertical) { dv[0] = -1; dv[1] = 1; } else { du[0] = -1; du[1] = 1; }
  for (int u = 0; u < wc; u++)
    for (int v = 0; v < hc; v++) {
      int d = dc[v * wc + u];
      if (d < 0) continue;
      int redundant = 1;
      for (int i = 0; i < 2 && redundant; i++) {
        int u2 = u, v2 = v, support = 0;
        for (int j = 0; j < maxdist; j++) {
          u2 += du[i]; v2 += dv[i];
          if (u2 < 0 || v2 < 0 || u2 >= wc || v2 >= hc) break;
          int d2 = dc[v2 * wc + u2];
          if (d2 >= 0 && abs(d - d2) <= thr) { support = 1; break; }
        }
        if (!support) redundant = 0;
      }
      if (redundant) dc[v * wc + u] = -1;
    }
}

void port_filter_dcan(const jn_elas_params* p, int16_t* dcan, int wc, int hc, int16_t* after_incon) {
  filter_inconsistent(p, dcan, wc, hc);
  if (after_incon) memcpy(after_incon, dcan, sizeof(int16_t) * wc * hc);
  filter_redundant(dcan, wc, hc, 5, 1, 1);
  filter_redundant(dcan, wc, hc, 5, 1, 0);
}

/* addCornerSupportPoints (elas.cpp:237-267) */
static int add_corners(int32_t* s, int n, int w, int h) {
  int32_t b[6][3] = {{0, 0, 0}, {0, h - 1, 0}, {w - 1, 0, 0}, {w - 1, h - 1, 0}, {0, 0, 0}, {0, 0, 0}};
  for (int i = 0; i < 4; i++) {
    int best = 10000000;
    for (int j = 0; j < n; j++) {
      int du = b[i][0] - s[3 * j], dv = b[i][1] - s[3 * j + 1];
      int dist = du * du + dv * dv;
      if (dist < best) { best = dist; b[i][2] = s[3 * j + 2]; }
    }
  }
  b[4][0] = b[2][0] + b[2][2]; b[4][1] = b[2][1]; b[4][2] = b[2][2];
  b[5][0] = b[3][0] + b[3][2]; b[5][1] = b[3][1]; b[5][2] = b[3][2];
  for (int i = 0; i < 6; i++) { s[3 * n] = b[i][0]; s[3 * n + 1] = b[i][1]; s[3 * n + 2] = b[i][2]; n++; }
  return n;
}

/* -------------------------------------------------------------------- */
/* Planes  (elas.cpp:507-577, matrix.cpp:414-502: Gauss-Jordan, full pivoting, double) */

static int gauss_jordan3(double A[3][3], double B[3], double eps) {
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
    if (fabs(A[icol][icol]) < eps) return 0;
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
  return 1; /* the column un-scrambling of matrix.cpp:489-494 only touches A */
}

void port_planes(const jn_elas_params* p, const int32_t* s, int n_support, const int32_t* tri,
                 int n_tri, int right_image, float* planes) {
  (void)p; (void)n_support; (void)right_image;
  for (int i = 0; i < n_tri; i++) {
    const int32_t* c = tri + 3 * i;
    for (int side = 0; side < 2; side++) {
      double A[3][3], B[3];
      for (int k = 0; k < 3; k++) {
        int u = s[3 * c[k]], v = s[3 * c[k] + 1], d = s[3 * c[k] + 2];
        A[k][0] = side ? u - d : u;
        A[k][1] = v;
        A[k][2] = 1;
        B[k] = d;
      }
      float* o = planes + 6 * i + 3 * side;
      if (gauss_jordan3(A, B, 1e-20)) { o[0] = (float)B[0]; o[1] = (float)B[1]; o[2] = (float)B[2]; }
      else { o[0] = o[1] = o[2] = 0; }
    }
  }
}

/* -------------------------------------------------------------------- */
/* Grid  (elas.cpp:579-659), including the row-wrapping dilation (H5) */

void port_grid(const jn_elas_params* p, int w, int h, const int32_t* s, int n, int right_image,
               int32_t* grid) {
  int gs = p->grid_size, D = p->disp_max + 1;
  int gw = (int)ceilf((float)w / (float)gs), gh = (int)ceilf((float)h / (float)gs);
  size_t cells = (size_t)gw * gh;
  uint8_t* t1 = (uint8_t*)calloc(cells * D, 1);
  uint8_t* t2 = (uint8_t*)calloc(cells * D, 1);
  for (int i = 0; i < n; i++) {
    int xc = s[3 * i], yc = s[3 * i + 1], dc = s[3 * i + 2];
    int lo = IMAX(dc - 1, 0), hi = IMIN(dc + 1, p->disp_max);
    for (int d = lo; d <= hi; d++) {
      int x = right_image ? (int)floorf((float)(xc - dc) / (float)gs) : (int)floorf((float)(xc / gs));
      int y = (int)floorf((float)yc / (float)gs);
      if (x >= 0 && x < gw && y >= 0 && y < gh) t1[((size_t)y * gw + x) * D + d] = 1;
    }
  }
  /* flat 3x3 OR over cell indices gw+1 .. gw*gh-gw-2 (elas.cpp:617-632) */
  if (gh >= 3 && gw >= 1) {
    long first = gw + 1, last = (long)gw * gh - gw - 2;
    const long off[9] = {-gw - 1, -gw, -gw + 1, -1, 0, 1, gw - 1, gw, gw + 1};
    for (long c = first; c <= last; c++)
      for (int d = 0; d < D; d++) {
        uint8_t r = 0;
        for (int k = 0; k < 9; k++) r |= t1[(size_t)(c + off[k]) * D + d];
        t2[(size_t)c * D + d] = r;
      }
  }
  memset(grid, 0, sizeof(int32_t) * cells * (D + 1));
  for (size_t c = 0; c < cells; c++) {
    int32_t* g = grid + c * (D + 1);
    int k = 0;
    for (int d = 0; d < D; d++)
      if (t2[c * D + d]) g[++k] = d;
    g[0] = k;
  }
  free(t1);
  free(t2);
}

/* -------------------------------------------------------------------- */
/* Dense matching  (elas.cpp:661-907) */

typedef struct { int64_t evals, pixels; } dense_count;

static void prior_table(const jn_elas_params* p, int disp_num, int32_t* P, int* plane_radius) {
  float two_sigma_squared = 2 * p->sigma * p->sigma;
  for (int dd = 0; dd < disp_num; dd++)
    P[dd] = (int32_t)((-logf(p->gamma + expf(-dd * dd / two_sigma_squared)) + logf(p->gamma)) / p->beta);
  *plane_radius = (int)fmaxf(ceilf(p->sigma * p->sradius), 2.0f);
}

/* findMatch (elas.cpp:683-780) */
static void find_match(const jn_elas_params* p, int w, int h, int u, int v, float pa, float pb, float pc,
                       const int32_t* grid, int gw, int disp_num, const uint8_t* desc_l,
                       const uint8_t* desc_r, const int32_t* P, int plane_radius, int valid,
                       int right_image, float* D, dense_count* cnt) {
  const int win = 2;
  if (u < win || u >= w - win) return;
  int vl = IMAX(IMIN(v, h - 3), 2);
  const uint8_t* A = (right_image ? desc_r : desc_l) + (size_t)vl * w * 16;
  const uint8_t* B = (right_image ? desc_l : desc_r) + (size_t)vl * w * 16;
  const uint8_t* a = A + 16 * u;
  if (texture16(a) < p->match_texture) return;
  int d_plane = (int32_t)(pa * (float)u + pb * (float)v + pc);
  int lo = IMAX(d_plane - plane_radius, 0), hi = IMIN(d_plane + plane_radius, disp_num - 1);
  int gx = u / p->grid_size, gy = v / p->grid_size;
  const int32_t* g = grid + ((size_t)gy * gw + gx) * (disp_num + 1);
  int num = g[0];
  int min_val = 10000, min_d = -1;
  cnt->pixels++;
  for (int i = 0; i < num; i++) {
    int d = g[1 + i];
    if (d < lo || d > hi) {
      int uw = right_image ? u + d : u - d;
      if (uw < win || uw >= w - win) continue;
      int val = sad16(a, B + 16 * uw);
      cnt->evals++;
      if (val < min_val) { min_val = val; min_d = d; }
    }
  }
  for (int d = lo; d <= hi; d++) {
    int uw = right_image ? u + d : u - d;
    if (uw < win || uw >= w - win) continue;
    int val = sad16(a, B + 16 * uw) + (valid ? P[abs(d - d_plane)] : 0);
    cnt->evals++;
    if (val < min_val) { min_val = val; min_d = d; }
  }
  size_t d_addr = p->subsampling ? (size_t)(v / 2) * (w / 2) + (u / 2) : (size_t)v * w + u;
  D[d_addr] = (min_d >= 0) ? (float)min_d : -1.0f;
}

static inline int32_t f2u_lo32(float x) { return (int32_t)(uint32_t)(int64_t)x; } /* H6 */

/* computeDisparity (elas.cpp:783-907) */
static void dense_port(const jn_elas_params* p, int w, int h, const uint8_t* desc1, const uint8_t* desc2,
                       const int32_t* s, const int32_t* tri, const float* planes, int n_tri,
                       const int32_t* grid, int right_image, float* D, dense_count* cnt) {
  int disp_num = p->disp_max + 1;
  int gw = (int)ceilf((float)w / (float)p->grid_size);
  const int sub = p->subsampling;
  const size_t nd = sub ? (size_t)(w / 2) * (h / 2) : (size_t)w * h;
  for (size_t i = 0; i < nd; i++) D[i] = -10;
  int32_t* P = (int32_t*)malloc(sizeof(int32_t) * disp_num);
  int plane_radius;
  prior_table(p, disp_num, P, &plane_radius);
  for (int i = 0; i < n_tri; i++) {
    const float* pl = planes + 6 * i;
    float pa, pb, pc, pd;
    if (!right_image) { pa = pl[0]; pb = pl[1]; pc = pl[2]; pd = pl[3]; }
    else { pa = pl[3]; pb = pl[4]; pc = pl[5]; pd = pl[0]; }
    float tu[3], tv[3];
    for (int k = 0; k < 3; k++) {
      int c = tri[3 * i + k];
      tu[k] = right_image ? (float)(s[3 * c] - s[3 * c + 2]) : (float)s[3 * c];
      tv[k] = (float)s[3 * c + 1];
    }
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < j; k++)
        if (tu[k] > tu[j]) {
          float t = tu[j]; tu[j] = tu[k]; tu[k] = t;
          t = tv[j]; tv[j] = tv[k]; tv[k] = t;
        }
    float Au = tu[0], Av = tv[0], Bu = tu[1], Bv = tv[1], Cu = tu[2], Cv = tv[2];
    float ABa = 0, ACa = 0, BCa = 0;
    if ((int)Au != (int)Bu) ABa = (Av - Bv) / (Au - Bu);
    if ((int)Au != (int)Cu) ACa = (Av - Cv) / (Au - Cu);
    if ((int)Bu != (int)Cu) BCa = (Bv - Cv) / (Bu - Cu);
    float ABb = Av - ABa * Au, ACb = Av - ACa * Au, BCb = Bv - BCa * Bu;
    int valid = fabs(pa) < 0.7 && fabs(pd) < 0.7;
    if ((int)Au != (int)Bu)
      for (int u = IMAX((int)Au, 0); u < IMIN((int)Bu, w); u++) {
        if (sub && u % 2) continue;
        int v1 = f2u_lo32(ACa * (float)u + ACb), v2 = f2u_lo32(ABa * (float)u + ABb);
        for (int v = IMIN(v1, v2); v < IMAX(v1, v2); v++)
          if (!sub || v % 2 == 0)
          find_match(p, w, h, u, v, pa, pb, pc, grid, gw, disp_num, desc1, desc2, P, plane_radius, valid,
                     right_image, D, cnt);
      }
    if ((int)Bu != (int)Cu)
      for (int u = IMAX((int)Bu, 0); u < IMIN((int)Cu, w); u++) {
        if (sub && u % 2) continue;
        int v1 = f2u_lo32(ACa * (float)u + ACb), v2 = f2u_lo32(BCa * (float)u + BCb);
        for (int v = IMIN(v1, v2); v < IMAX(v1, v2); v++)
          if (!sub || v % 2 == 0)
          find_match(p, w, h, u, v, pa, pb, pc, grid, gw, disp_num, desc1, desc2, P, plane_radius, valid,
                     right_image, D, cnt);
      }
  }
  free(P);
}

void port_dense(const jn_elas_params* p, int w, int h, const uint8_t* desc1, const uint8_t* desc2,
                const int32_t* support, int n_support, const int32_t* tri, const float* planes,
                int n_tri, const int32_t* grid, int right_image, float* D) {
  (void)n_support;
  dense_count c = {0, 0};
  dense_port(p, w, h, desc1, desc2, support, tri, planes, n_tri, grid, right_image, D, &c);
}

/* Study helper (tests / design notes only): how do the scan-converted triangles of computeDisparity
 * (elas.cpp:843-903) overlap?  Counts pixels covered by more than one triangle, and among those the
 * ones where some covering triangle has the pixel strictly inside its column span (not in the span's
 * first or last row after clamping to the image).  out = {pixels covered, pixels covered more than
 * once, of those with an interior covering, maximum cover count}. */
void port_raster_overlap_study(int w, int h, const int32_t* s, const int32_t* tri, int n_tri, int right_image,
                               int64_t out[4]) {
  uint8_t* cnt = (uint8_t*)calloc((size_t)w * h, 1);
  uint8_t* inter = (uint8_t*)calloc((size_t)w * h, 1);
  for (int i = 0; i < n_tri; i++) {
    float tu[3], tv[3];
    for (int k = 0; k < 3; k++) {
      int c = tri[3 * i + k];
      tu[k] = right_image ? (float)(s[3 * c] - s[3 * c + 2]) : (float)s[3 * c];
      tv[k] = (float)s[3 * c + 1];
    }
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < j; k++)
        if (tu[k] > tu[j]) {
          float t = tu[j]; tu[j] = tu[k]; tu[k] = t;
          t = tv[j]; tv[j] = tv[k]; tv[k] = t;
        }
    float Au = tu[0], Av = tv[0], Bu = tu[1], Bv = tv[1], Cu = tu[2], Cv = tv[2];
    float ABa = 0, ACa = 0, BCa = 0;
    if ((int)Au != (int)Bu) ABa = (Av - Bv) / (Au - Bu);
    if ((int)Au != (int)Cu) ACa = (Av - Cv) / (Au - Cu);
    if ((int)Bu != (int)Cu) BCa = (Bv - Cv) / (Bu - Cu);
    float ABb = Av - ABa * Au, ACb = Av - ACa * Au, BCb = Bv - BCa * Bu;
    for (int half = 0; half < 2; half++) {
      if (half == 0 ? (int)Au == (int)Bu : (int)Bu == (int)Cu) continue;
      int u0 = IMAX((int)(half ? Bu : Au), 0), u1 = IMIN((int)(half ? Cu : Bu), w);
      for (int u = u0; u < u1; u++) {
        int v1 = f2u_lo32(ACa * (float)u + ACb);
        int v2 = f2u_lo32(half ? BCa * (float)u + BCb : ABa * (float)u + ABb);
        int vlo = IMAX(IMIN(v1, v2), 0), vhi = IMIN(IMAX(v1, v2), h);
        for (int v = vlo; v < vhi; v++) {
          size_t a = (size_t)v * w + u;
          if (cnt[a] < 255) cnt[a]++;
          if (v > vlo && v < vhi - 1) inter[a] = 1;
        }
      }
    }
  }
  out[0] = out[1] = out[2] = out[3] = 0;
  for (size_t a = 0; a < (size_t)w * h; a++) {
    if (cnt[a]) out[0]++;
    if (cnt[a] > 1) { out[1]++; if (inter[a]) out[2]++; }
    if (cnt[a] > out[3]) out[3] = cnt[a];
  }
  free(cnt);
  free(inter);
}

/* prior table as used by the dense stage, exported for the host-logic tests */
void port_prior(const jn_elas_params* p, int32_t* P_out, int32_t* plane_radius_out) {
  int r;
  prior_table(p, p->disp_max + 1, P_out, &r);
  *plane_radius_out = r;
}

/* -------------------------------------------------------------------- */
/* Post-processing  (elas.cpp:909-1560), subsampling = 0 only */

/* leftRightConsistencyCheck (elas.cpp:909-979) */
static void lr_check(const jn_elas_params* p, int w, int h, float* D1, float* D2) {
  const int sub = p->subsampling;   /* w,h are the map dimensions (already halved) */
  size_t n = (size_t)w * h;
  float* c1 = (float*)malloc(n * sizeof(float));
  float* c2 = (float*)malloc(n * sizeof(float));
  memcpy(c1, D1, n * sizeof(float));
  memcpy(c2, D2, n * sizeof(float));
  for (int u = 0; u < w; u++)
    for (int v = 0; v < h; v++) {
      size_t a = (size_t)v * w + u;
      float d1 = c1[a], d2 = c2[a];
      float uw1 = sub ? (float)u - d1 / 2 : (float)u - d1, uw2 = sub ? (float)u + d2 / 2 : (float)u + d2;
      if (d1 >= 0 && uw1 >= 0 && uw1 < w) {
        if (fabs(c2[(size_t)v * w + (int)uw1] - d1) > p->lr_threshold) D1[a] = -10;
      } else D1[a] = -10;
      if (d2 >= 0 && uw2 >= 0 && uw2 < w) {
        if (fabs(c1[(size_t)v * w + (int)uw2] - d2) > p->lr_threshold) D2[a] = -10;
      } else D2[a] = -10;
    }
  free(c1);
  free(c2);
}

/* removeSmallSegments (elas.cpp:981-1099): breadth-first flood fill, seeds u outer / v inner */
static void remove_small_segments(const jn_elas_params* p, int w, int h, float* D) {
  /* elas.cpp:986-991: at half resolution the size limit becomes (int)(sqrt(speckle_size)*2) */
  const int speckle_size = p->subsampling ? (int)(sqrtf((float)p->speckle_size) * 2) : p->speckle_size;
  size_t n = (size_t)w * h;
  uint8_t* done = (uint8_t*)calloc(n, 1);
  int32_t* list = (int32_t*)malloc(n * sizeof(int32_t));
  for (int u = 0; u < w; u++)
    for (int v = 0; v < h; v++) {
      size_t start = (size_t)v * w + u;
      if (done[start]) continue;
      int count = 1, curr = 0;
      list[0] = (int32_t)start;
      while (curr < count) {
        int a = list[curr];
        int cu = a % w, cv = a / w;
        const int nu[4] = {cu - 1, cu + 1, cu, cu}, nv[4] = {cv, cv, cv - 1, cv + 1};
        for (int i = 0; i < 4; i++)
          if (nu[i] >= 0 && nv[i] >= 0 && nu[i] < w && nv[i] < h) {
            int b = nv[i] * w + nu[i];
            if (!done[b] && D[b] >= 0 && fabs(D[a] - D[b]) <= p->speckle_sim_threshold) {
              list[count++] = b;
              done[b] = 1;
            }
          }
        curr++;
        done[a] = 1;
      }
      if (count < speckle_size)
        for (int i = 0; i < count; i++) D[list[i]] = -10;
    }
  free(done);
  free(list);
}

/* one line of gapInterpolation (elas.cpp:1122-1199 rows, 1203-1283 columns) */
static void gap_line(const jn_elas_params* p, float* D, int len, size_t stride) {
  int gap = p->subsampling ? p->ipol_gap_width / 2 + 1 : p->ipol_gap_width, count = 0;   /* elas.cpp:1106-1111 */
  for (int i = 0; i < len; i++) {
    if (D[i * stride] >= 0) {
      if (count >= 1 && count <= gap) {
        int first = i - count, last = i - 1;
        if (first > 0 && last < len - 1) {
          float d1 = D[(first - 1) * stride], d2 = D[(last + 1) * stride];
          float dip = (fabs(d1 - d2) < 3.0f) ? (d1 + d2) / 2 : (d1 < d2 ? d1 : d2);
          for (int k = first; k <= last; k++) D[k * stride] = dip;
        }
      }
      count = 0;
    } else count++;
  }
  if (p->add_corners) {
    for (int i = 0; i < len; i++)
      if (D[i * stride] >= 0) {
        for (int k = IMAX(i - gap, 0); k < i; k++) D[k * stride] = D[i * stride];
        break;
      }
    for (int i = len - 1; i >= 0; i--)
      if (D[i * stride] >= 0) {
        for (int k = i; k <= IMIN(i + gap, len - 1); k++) D[k * stride] = D[i * stride];
        break;
      }
  }
}

static void gap_interpolation(const jn_elas_params* p, int w, int h, float* D) {
  for (int v = 0; v < h; v++) gap_line(p, D + (size_t)v * w, w, 1);
  for (int u = 0; u < w; u++) gap_line(p, D + u, h, (size_t)w);
}

/* The reference's "absolute value" mask is _mm_set1_ps(0x7FFFFFFF): the integer is
 * CONVERTED to float (2^31, bits 0x4F000000), so the AND keeps only exponent bits
 * 0x9E and clears sign + mantissa (elas.cpp:1320, 1413). */
static inline float buggy_abs(float x) {
  uint32_t b;
  memcpy(&b, &x, 4);
  b &= 0x4F000000u;
  memcpy(&x, &b, 4);
  return x;
}

/* one output of the 8-tap filter; win[k] is the sample whose coordinate is k mod 8
 * (elas.cpp:1411-1436) */
static inline int mean8(const float win[8], float centre, float* out) {
  float wgt[8], fac[8];
  for (int k = 0; k < 8; k++) {
    float t = 4.0f - buggy_abs(win[k] - centre);
    wgt[k] = t > 0.0f ? t : 0.0f;       /* _mm_max_ps(0, t) */
    fac[k] = win[k] * wgt[k];
  }
  float w4[4], f4[4];
  for (int k = 0; k < 4; k++) { w4[k] = wgt[k] + wgt[k + 4]; f4[k] = fac[k] + fac[k + 4]; }
  float ws = w4[0] + w4[1] + w4[2] + w4[3];
  float fs = f4[0] + f4[1] + f4[2] + f4[3];
  if (ws > 0) {
    float d = fs / ws;
    if (d >= 0) { *out = d; return 1; }
  }
  return 0;
}

/* one output of the 4-tap filter of the half-resolution branch (elas.cpp:1339-1355) */
static inline int mean4(const float win[4], float centre, float* out) {
  float wgt[4], fac[4];
  for (int k = 0; k < 4; k++) {
    float t = 4.0f - buggy_abs(win[k] - centre);
    wgt[k] = t > 0.0f ? t : 0.0f;
    fac[k] = win[k] * wgt[k];
  }
  float ws = wgt[0] + wgt[1] + wgt[2] + wgt[3];
  float fs = fac[0] + fac[1] + fac[2] + fac[3];
  if (ws > 0) {
    float d = fs / ws;
    if (d >= 0) { *out = d; return 1; }
  }
  return 0;
}

/* adaptiveMean, half resolution branch (elas.cpp:1323-1391): window = last 4 samples, centre u-1 */
static void adaptive_mean_half(int w, int h, float* D) {
  size_t n = (size_t)w * h;
  float* cp = (float*)malloc(n * sizeof(float));
  float* tmp = (float*)calloc(n, sizeof(float)); /* H1: unwritten = 0 */
  memcpy(cp, D, n * sizeof(float));
  for (size_t i = 0; i < n; i++)
    if (D[i] < 0) { cp[i] = -10; tmp[i] = -10; }
  float win[4];
  if (w >= 4)
    for (int v = 3; v < h - 3; v++) {
      const float* row = cp + (size_t)v * w;
      for (int u = 0; u < 3; u++) win[u] = row[u];
      for (int u = 3; u < w; u++) {
        win[u % 4] = row[u];
        float o;
        if (mean4(win, row[u - 1], &o)) tmp[(size_t)v * w + (u - 1)] = o;
      }
    }
  if (h >= 4)
    for (int u = 3; u < w - 3; u++) {
      for (int v = 0; v < 3; v++) win[v] = tmp[(size_t)v * w + u];
      for (int v = 3; v < h; v++) {
        win[v % 4] = tmp[(size_t)v * w + u];
        float o;
        if (mean4(win, tmp[(size_t)(v - 1) * w + u], &o)) D[(size_t)(v - 1) * w + u] = o;
      }
    }
  free(cp);
  free(tmp);
}

/* adaptiveMean, full resolution branch (elas.cpp:1287-1320, 1394-1492) */
static void adaptive_mean(int w, int h, float* D) {
  size_t n = (size_t)w * h;
  float* cp = (float*)malloc(n * sizeof(float));
  float* tmp = (float*)calloc(n, sizeof(float)); /* H1: unwritten = 0 */
  memcpy(cp, D, n * sizeof(float));
  for (size_t i = 0; i < n; i++)
    if (D[i] < 0) { cp[i] = -10; tmp[i] = -10; }
  float win[8];
  if (w >= 8)
    for (int v = 3; v < h - 3; v++) {
      const float* row = cp + (size_t)v * w;
      for (int u = 0; u < 7; u++) win[u] = row[u];
      for (int u = 7; u < w; u++) {
        win[u % 8] = row[u];
        float o;
        if (mean8(win, row[u - 3], &o)) tmp[(size_t)v * w + (u - 3)] = o;
      }
    }
  if (h >= 8)
    for (int u = 3; u < w - 3; u++) {
      for (int v = 0; v < 7; v++) win[v] = tmp[(size_t)v * w + u];
      for (int v = 7; v < h; v++) {
        win[v % 8] = tmp[(size_t)v * w + u];
        float o;
        if (mean8(win, tmp[(size_t)(v - 3) * w + u], &o)) D[(size_t)(v - 3) * w + u] = o;
      }
    }
  free(cp);
  free(tmp);
}

static float median7(const float* x, size_t stride) {
  float v[7];
  for (int j = 0; j < 7; j++) { /* insertion sort as elas.cpp:1518-1527 */
    float t = x[j * stride];
    int i = j - 1;
    while (i >= 0 && v[i] > t) { v[i + 1] = v[i]; i--; }
    v[i + 1] = t;
  }
  return v[3];
}

/* median (elas.cpp:1494-1560) */
static void median_filter(int w, int h, float* D) {
  size_t n = (size_t)w * h;
  float* tmp = (float*)calloc(n, sizeof(float));
  for (int u = 3; u < w - 3; u++)
    for (int v = 3; v < h - 3; v++) {
      size_t a = (size_t)v * w + u;
      tmp[a] = (D[a] >= 0) ? median7(D + a - 3, 1) : D[a];
    }
  for (int u = 3; u < w - 3; u++)
    for (int v = 3; v < h - 3; v++) {
      size_t a = (size_t)v * w + u;
      if (D[a] >= 0) D[a] = median7(tmp + a - 3 * (size_t)w, (size_t)w);
    }
  free(tmp);
}

static void put(float* dst, const float* src, size_t n) {
  if (dst) memcpy(dst, src, n * sizeof(float));
}

/* elas.cpp:108-140 */
void port_postprocess(const jn_elas_params* p, int w, int h, float* D1, float* D2, oracle_stages* st) {
  if (p->subsampling) { w /= 2; h /= 2; }   /* the maps are (w/2) x (h/2), elas.cpp:914-917 */
  size_t n = (size_t)w * h;
  lr_check(p, w, h, D1, D2);
  if (st) { put(st->D1_lr, D1, n); put(st->D2_lr, D2, n); }
  remove_small_segments(p, w, h, D1);
  if (!p->postprocess_only_left) remove_small_segments(p, w, h, D2);
  if (st) { put(st->D1_seg, D1, n); put(st->D2_seg, D2, n); }
  gap_interpolation(p, w, h, D1);
  if (!p->postprocess_only_left) gap_interpolation(p, w, h, D2);
  if (st) { put(st->D1_gap, D1, n); put(st->D2_gap, D2, n); }
  if (p->filter_adaptive_mean) {
    if (p->subsampling) adaptive_mean_half(w, h, D1); else adaptive_mean(w, h, D1);
    if (!p->postprocess_only_left) { if (p->subsampling) adaptive_mean_half(w, h, D2); else adaptive_mean(w, h, D2); }
  }
  if (st) { put(st->D1_mean, D1, n); put(st->D2_mean, D2, n); }
  if (p->filter_median) {
    median_filter(w, h, D1);
    if (!p->postprocess_only_left) median_filter(w, h, D2);
  }
  if (st) { put(st->D1, D1, n); put(st->D2, D2, n); }
}

/* -------------------------------------------------------------------- */
/* Whole pipeline  (elas.cpp:32-151, 375-443) */

int port_triangulate(const float* xy, int n, int32_t* tri_out, int cap_tri); /* delaunay_port.c */

static int triangulate_support(const int32_t* s, int n, int right_image, int32_t* tri, int cap) {
  float* xy = (float*)malloc(sizeof(float) * 2 * n);
  for (int i = 0; i < n; i++) {
    xy[2 * i] = right_image ? (float)(s[3 * i] - s[3 * i + 2]) : (float)s[3 * i];
    xy[2 * i + 1] = (float)s[3 * i + 1];
  }
  int nt = port_triangulate(xy, n, tri, cap);
  free(xy);
  return nt;
}

int port_elas_stages(const jn_elas_params* p, const uint8_t* I1, const uint8_t* I2, const int32_t* dims,
                     oracle_stages* st) {
  int w = dims[0], h = dims[1], stride = dims[2];
  size_t n = (size_t)w * h;
  uint8_t* desc1 = (uint8_t*)malloc(16 * n);
  uint8_t* desc2 = (uint8_t*)malloc(16 * n);
  descriptor_port(I1, w, h, stride, desc1, p->subsampling);
  descriptor_port(I2, w, h, stride, desc2, p->subsampling);
  if (st->desc1) memcpy(st->desc1, desc1, 16 * n);
  if (st->desc2) memcpy(st->desc2, desc2, 16 * n);

  int step = p->candidate_stepsize;
  if (p->subsampling) step += step % 2;   /* only every second line has descriptors (elas.cpp:379-381) */
  int wc = (w + step - 1) / step, hc = (h + step - 1) / step;
  int16_t* dc = (int16_t*)calloc((size_t)wc * hc, sizeof(int16_t)); /* row 0 / col 0 stay 0 (H3) */
  for (int uc = 1; uc < wc; uc++)
    for (int vc = 1; vc < hc; vc++) {
      int u = uc * step, v = vc * step;
      dc[vc * wc + uc] = -1;
      int d = match_support(p, w, h, u, v, desc1, desc2, 0);
      if (d >= 0) {
        int d2 = match_support(p, w, h, u - d, v, desc1, desc2, 1);
        if (d2 >= 0 && abs(d - d2) <= p->lr_threshold) dc[vc * wc + uc] = (int16_t)d;
      }
    }
  if (st->dcan_raw) memcpy(st->dcan_raw, dc, sizeof(int16_t) * wc * hc);
  filter_inconsistent(p, dc, wc, hc);
  if (st->dcan_incon) memcpy(st->dcan_incon, dc, sizeof(int16_t) * wc * hc);
  filter_redundant(dc, wc, hc, 5, 1, 1);
  filter_redundant(dc, wc, hc, 5, 1, 0);
  if (st->dcan_final) memcpy(st->dcan_final, dc, sizeof(int16_t) * wc * hc);

  int32_t* s = (int32_t*)malloc(sizeof(int32_t) * 3 * ((size_t)wc * hc + 8));
  int ns = 0;
  for (int uc = 1; uc < wc; uc++)
    for (int vc = 1; vc < hc; vc++)
      if (dc[vc * wc + uc] >= 0) {
        s[3 * ns] = uc * step; s[3 * ns + 1] = vc * step; s[3 * ns + 2] = dc[vc * wc + uc];
        ns++;
      }
  if (p->add_corners) ns = add_corners(s, ns, w, h);
  free(dc);
  st->n_support = ns;
  if (st->support && ns <= st->cap_support) memcpy(st->support, s, sizeof(int32_t) * 3 * ns);

  int rc = 0;
  if (ns < 3) {
    rc = 1;
  } else {
    int cap = 2 * ns + 8;
    int32_t* tri1 = (int32_t*)malloc(sizeof(int32_t) * 3 * cap);
    int32_t* tri2 = (int32_t*)malloc(sizeof(int32_t) * 3 * cap);
    int nt1 = triangulate_support(s, ns, 0, tri1, cap);
    int nt2 = triangulate_support(s, ns, 1, tri2, cap);
    float* pl1 = (float*)malloc(sizeof(float) * 6 * cap);
    float* pl2 = (float*)malloc(sizeof(float) * 6 * cap);
    port_planes(p, s, ns, tri1, nt1, 0, pl1);
    port_planes(p, s, ns, tri2, nt2, 1, pl2);
    st->n_tri1 = nt1;
    st->n_tri2 = nt2;
    if (nt1 <= st->cap_tri) {
      if (st->tri1) memcpy(st->tri1, tri1, sizeof(int32_t) * 3 * nt1);
      if (st->planes1) memcpy(st->planes1, pl1, sizeof(float) * 6 * nt1);
    }
    if (nt2 <= st->cap_tri) {
      if (st->tri2) memcpy(st->tri2, tri2, sizeof(int32_t) * 3 * nt2);
      if (st->planes2) memcpy(st->planes2, pl2, sizeof(float) * 6 * nt2);
    }
    int gw = (int)ceilf((float)w / (float)p->grid_size), gh = (int)ceilf((float)h / (float)p->grid_size);
    size_t gn = (size_t)gw * gh * (p->disp_max + 2);
    int32_t* g1 = (int32_t*)malloc(gn * sizeof(int32_t));
    int32_t* g2 = (int32_t*)malloc(gn * sizeof(int32_t));
    port_grid(p, w, h, s, ns, 0, g1);
    port_grid(p, w, h, s, ns, 1, g2);
    if (st->grid1) memcpy(st->grid1, g1, gn * sizeof(int32_t));
    if (st->grid2) memcpy(st->grid2, g2, gn * sizeof(int32_t));
    float* D1 = (float*)malloc(n * sizeof(float));
    float* D2 = (float*)malloc(n * sizeof(float));
    dense_count cnt = {0, 0};
    dense_port(p, w, h, desc1, desc2, s, tri1, pl1, nt1, g1, 0, D1, &cnt);
    dense_port(p, w, h, desc1, desc2, s, tri2, pl2, nt2, g2, 1, D2, &cnt);
    st->dense_evals = cnt.evals;
    st->dense_pixels = cnt.pixels;
    const size_t nd = p->subsampling ? (size_t)(w / 2) * (h / 2) : n;
    put(st->D1_raw, D1, nd);
    put(st->D2_raw, D2, nd);
    port_postprocess(p, w, h, D1, D2, st);
    free(D1); free(D2); free(g1); free(g2); free(pl1); free(pl2); free(tri1); free(tri2);
  }
  free(s); free(desc1); free(desc2);
  return rc;
}

/* Elas::process semantics: on "<3 support points" D1/D2 stay untouched (elas.cpp:66-71). */
int port_elas_process(const jn_elas_params* p, const uint8_t* I1, const uint8_t* I2, float* D1, float* D2,
                      const int32_t* dims) {
  size_t n = (size_t)dims[0] * dims[1];
  oracle_stages st;
  memset(&st, 0, sizeof(st));
  float* a = (float*)malloc(n * sizeof(float));
  float* b = (float*)malloc(n * sizeof(float));
  st.D1 = a;
  st.D2 = b;
  int rc = port_elas_stages(p, I1, I2, dims, &st);
  if (p->subsampling) n = (size_t)(dims[0] / 2) * (dims[1] / 2);
  if (rc == 0) { memcpy(D1, a, n * sizeof(float)); memcpy(D2, b, n * sizeof(float)); }
  free(a);
  free(b);
  return rc;
}

/* single post-processing filters, exported for unit tests */
void port_adaptive_mean(int w, int h, float* D) { adaptive_mean(w, h, D); }
void port_median(int w, int h, float* D) { median_filter(w, h, D); }
void port_gap_interpolation(const jn_elas_params* p, int w, int h, float* D) { gap_interpolation(p, w, h, D); }
void port_remove_small_segments(const jn_elas_params* p, int w, int h, float* D) { remove_small_segments(p, w, h, D); }
