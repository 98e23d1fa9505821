/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin extern "C" access to the UNMODIFIED reference ELAS.  The reference
 * sources are compiled where they lie (/root/reference/src/elas, see
 * oracle/Makefile); nothing is copied into this repository.  elas.cpp is
 * pulled into this translation unit with `private` opened up so the parity
 * tests can dump every intermediate stage of Elas::process (elas.cpp:32-151)
 * and drive single stages with injected inputs.
 *
 * ref_elas_stages() replays the stage sequence of Elas::process /
 * computeSupportMatches by CALLING the reference's own member functions in
 * the reference's order; it contains no arithmetic of its own.
 */
#include <malloc.h>
#include <string.h>
#include <stdlib.h>
#include <vector>

#include <algorithm>
#include <math.h>
#include <pmmintrin.h>

#define private public
#include "elas.h"
#undef private
#include "descriptor.h"
#include "matrix.h"
#include "filter.h"

/* adaptiveMean (elas.cpp:1297-1298) reads rows of D_tmp it never wrote.  Every header the
 * reference pulls in has been seen above, so this only rewrites the malloc CALLS inside elas.cpp:
 * its scratch buffers start zeroed, which is what a fresh mmap'd block gives the stock build and what
 * the port and the CUDA path define those reads to be.  (ref_set_deterministic_heap below covers the
 * separately compiled descriptor.cpp the same way for blocks glibc serves from fresh pages.) */
#define malloc(n) calloc(1, (n))
#include "elas.cpp"          /* reference translation unit, in place */
#undef malloc

#include "oracle_abi.h"

using std::vector;

static Elas::parameters to_ref(const jn_elas_params* p) {
  Elas::parameters q;
  q.disp_min = p->disp_min;
  q.disp_max = p->disp_max;
  q.support_threshold = p->support_threshold;
  q.support_texture = p->support_texture;
  q.candidate_stepsize = p->candidate_stepsize;
  q.incon_window_size = p->incon_window_size;
  q.incon_threshold = p->incon_threshold;
  q.incon_min_support = p->incon_min_support;
  q.add_corners = p->add_corners != 0;
  q.grid_size = p->grid_size;
  q.beta = p->beta;
  q.gamma = p->gamma;
  q.sigma = p->sigma;
  q.sradius = p->sradius;
  q.match_texture = p->match_texture;
  q.lr_threshold = p->lr_threshold;
  q.speckle_sim_threshold = p->speckle_sim_threshold;
  q.speckle_size = p->speckle_size;
  q.ipol_gap_width = p->ipol_gap_width;
  q.filter_median = p->filter_median != 0;
  q.filter_adaptive_mean = p->filter_adaptive_mean != 0;
  q.postprocess_only_left = p->postprocess_only_left != 0;
  q.subsampling = p->subsampling != 0;
  return q;
}

static void put(float* dst, const float* src, size_t n) {
  if (dst) memcpy(dst, src, n * sizeof(float));
}

extern "C" {

/* the reference reads descriptor bytes / D_tmp floats it never
 * wrote.  With a low mmap threshold every large block is a fresh zero page,
 * which makes those reads deterministic (= 0). */
void ref_set_deterministic_heap(int on) {
  mallopt(M_MMAP_THRESHOLD, on ? 4096 : 128 * 1024);
}

void ref_params_default(jn_elas_params* p, int setting) {
  Elas::parameters q(setting == 0 ? Elas::ROBOTICS : Elas::MIDDLEBURY);
  p->disp_min = q.disp_min; p->disp_max = q.disp_max;
  p->support_threshold = q.support_threshold; p->support_texture = q.support_texture;
  p->candidate_stepsize = q.candidate_stepsize; p->incon_window_size = q.incon_window_size;
  p->incon_threshold = q.incon_threshold; p->incon_min_support = q.incon_min_support;
  p->add_corners = q.add_corners; p->grid_size = q.grid_size;
  p->beta = q.beta; p->gamma = q.gamma; p->sigma = q.sigma; p->sradius = q.sradius;
  p->match_texture = q.match_texture; p->lr_threshold = q.lr_threshold;
  p->speckle_sim_threshold = q.speckle_sim_threshold; p->speckle_size = q.speckle_size;
  p->ipol_gap_width = q.ipol_gap_width; p->filter_median = q.filter_median;
  p->filter_adaptive_mean = q.filter_adaptive_mean;
  p->postprocess_only_left = q.postprocess_only_left; p->subsampling = q.subsampling;
}

/* The reference call, untouched: Elas(param).process(...)  (point_cloud.cpp:416-419). */
int ref_elas_process(const jn_elas_params* p, const uint8_t* I1, const uint8_t* I2,
                     float* D1, float* D2, const int32_t* dims) {
  Elas elas(to_ref(p));
  elas.process((uint8_t*)I1, (uint8_t*)I2, D1, D2, dims);
  return 0;
}

/* Descriptor::Descriptor on an image given with stride bpl_in (elas.cpp:37-58). */
int ref_descriptor(const uint8_t* I, int w, int h, int bpl_in, uint8_t* desc_out) {
  int bpl = w + 15 - (w - 1) % 16;
  uint8_t* A = (uint8_t*)_mm_malloc((size_t)bpl * h, 16);
  memset(A, 0, (size_t)bpl * h);
  for (int v = 0; v < h; v++) memcpy(A + (size_t)v * bpl, I + (size_t)v * bpl_in, w);
  {
    Descriptor d(A, w, h, bpl, false);
    memcpy(desc_out, d.I_desc, (size_t)16 * w * h);
  }
  _mm_free(A);
  return 0;
}

/* triangulate("zQB") exactly as computeDelaunayTriangulation calls it (elas.cpp:445-505). */
int ref_triangulate(const float* xy, int n, int32_t* tri_out, int cap_tri) {
  struct triangulateio in, out;
  memset(&in, 0, sizeof(in));
  memset(&out, 0, sizeof(out));
  in.numberofpoints = n;
  in.pointlist = (float*)malloc(sizeof(float) * 2 * n);
  memcpy(in.pointlist, xy, sizeof(float) * 2 * n);
  char sw[] = "zQB";
  triangulate(sw, &in, &out, NULL);
  int nt = out.numberoftriangles;
  if (nt <= cap_tri) memcpy(tri_out, out.trianglelist, sizeof(int) * 3 * nt);
  free(in.pointlist);
  free(out.pointlist);
  free(out.trianglelist);
  return nt;
}

/* In-place support filtering on a caller-supplied candidate image
 * (removeInconsistentSupportPoints + 2x removeRedundantSupportPoints, elas.cpp:416-422). */
void ref_filter_dcan(const jn_elas_params* p, int16_t* dcan, int wc, int hc, int16_t* after_incon) {
  Elas elas(to_ref(p));
  elas.removeInconsistentSupportPoints(dcan, wc, hc);
  if (after_incon) memcpy(after_incon, dcan, sizeof(int16_t) * wc * hc);
  elas.removeRedundantSupportPoints(dcan, wc, hc, 5, 1, true);
  elas.removeRedundantSupportPoints(dcan, wc, hc, 5, 1, false);
}

static vector<Elas::support_pt> to_support(const int32_t* s, int n) {
  vector<Elas::support_pt> v;
  v.reserve(n);
  for (int i = 0; i < n; i++) v.push_back(Elas::support_pt(s[3 * i], s[3 * i + 1], s[3 * i + 2]));
  return v;
}

/* computeDisparityPlanes on injected support points + triangles (elas.cpp:507-577). */
void ref_planes(const jn_elas_params* p, const int32_t* support, int n_support,
                const int32_t* tri, int n_tri, int right_image, float* planes_out) {
  Elas elas(to_ref(p));
  vector<Elas::support_pt> sp = to_support(support, n_support);
  vector<Elas::triangle> t;
  for (int i = 0; i < n_tri; i++) t.push_back(Elas::triangle(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]));
  elas.computeDisparityPlanes(sp, t, right_image);
  for (int i = 0; i < n_tri; i++) {
    planes_out[6 * i + 0] = t[i].t1a; planes_out[6 * i + 1] = t[i].t1b; planes_out[6 * i + 2] = t[i].t1c;
    planes_out[6 * i + 3] = t[i].t2a; planes_out[6 * i + 4] = t[i].t2b; planes_out[6 * i + 5] = t[i].t2c;
  }
}

/* createGrid on injected support points (elas.cpp:579-659). */
void ref_grid(const jn_elas_params* p, int w, int h, const int32_t* support, int n_support,
              int right_image, int32_t* grid_out) {
  Elas elas(to_ref(p));
  elas.width = w; elas.height = h;
  int gw = (int)ceil((float)w / (float)p->grid_size);
  int gh = (int)ceil((float)h / (float)p->grid_size);
  int32_t gd[3] = {p->disp_max + 2, gw, gh};
  memset(grid_out, 0, sizeof(int32_t) * (size_t)(p->disp_max + 2) * gw * gh);
  elas.createGrid(to_support(support, n_support), grid_out, gd, right_image != 0);
}

/* computeDisparity with injected descriptors / support / triangles+planes / grid
 * (elas.cpp:783-907).  desc buffers must be 16-byte aligned. */
void ref_dense(const jn_elas_params* p, int w, int h, const uint8_t* desc1, const uint8_t* desc2,
               const int32_t* support, int n_support, const int32_t* tri, const float* planes,
               int n_tri, const int32_t* grid, int right_image, float* D) {
  Elas elas(to_ref(p));
  elas.width = w; elas.height = h; elas.bpl = w + 15 - (w - 1) % 16;
  vector<Elas::triangle> t;
  for (int i = 0; i < n_tri; i++) {
    Elas::triangle x(tri[3 * i], tri[3 * i + 1], tri[3 * i + 2]);
    x.t1a = planes[6 * i + 0]; x.t1b = planes[6 * i + 1]; x.t1c = planes[6 * i + 2];
    x.t2a = planes[6 * i + 3]; x.t2b = planes[6 * i + 4]; x.t2c = planes[6 * i + 5];
    t.push_back(x);
  }
  int gw = (int)ceil((float)w / (float)p->grid_size);
  int gh = (int)ceil((float)h / (float)p->grid_size);
  int32_t gd[3] = {p->disp_max + 2, gw, gh};
  elas.computeDisparity(to_support(support, n_support), t, (int32_t*)grid, gd,
                        (uint8_t*)desc1, (uint8_t*)desc2, right_image != 0, D);
}

/* Post-processing chain on injected raw disparity maps, in place
 * (elas.cpp:108-140).  Stage copies optional. */
void ref_postprocess(const jn_elas_params* p, int w, int h, float* D1, float* D2, oracle_stages* st) {
  Elas elas(to_ref(p));
  elas.width = w; elas.height = h; elas.bpl = w + 15 - (w - 1) % 16;
  /* w, h = IMAGE size; with subsampling the maps hold (w/2) x (h/2) floats (elas.cpp:914-917) */
  size_t n = p->subsampling ? (size_t)(w / 2) * (h / 2) : (size_t)w * h;
  elas.leftRightConsistencyCheck(D1, D2);
  if (st) { put(st->D1_lr, D1, n); put(st->D2_lr, D2, n); }
  elas.removeSmallSegments(D1);
  if (!elas.param.postprocess_only_left) elas.removeSmallSegments(D2);
  if (st) { put(st->D1_seg, D1, n); put(st->D2_seg, D2, n); }
  elas.gapInterpolation(D1);
  if (!elas.param.postprocess_only_left) elas.gapInterpolation(D2);
  if (st) { put(st->D1_gap, D1, n); put(st->D2_gap, D2, n); }
  if (elas.param.filter_adaptive_mean) {
    elas.adaptiveMean(D1);
    if (!elas.param.postprocess_only_left) elas.adaptiveMean(D2);
  }
  if (st) { put(st->D1_mean, D1, n); put(st->D2_mean, D2, n); }
  if (elas.param.filter_median) {
    elas.median(D1);
    if (!elas.param.postprocess_only_left) elas.median(D2);
  }
  if (st) { put(st->D1, D1, n); put(st->D2, D2, n); }
}

/* Full pipeline with every stage dumped.  Mirrors the call order of
 * Elas::process (elas.cpp:32-151) and computeSupportMatches (elas.cpp:375-443).
 * D1/D2 in `st` (if given) receive the final maps; returns 1 for the
 * "<3 support points" early return. */
int ref_elas_stages(const jn_elas_params* p, const uint8_t* I1_, const uint8_t* I2_,
                    const int32_t* dims, oracle_stages* st) {
  Elas elas(to_ref(p));
  const Elas::parameters& param = elas.param;
  int width = elas.width = dims[0];
  int height = elas.height = dims[1];
  int bpl = elas.bpl = width + 15 - (width - 1) % 16;
  size_t n = (size_t)width * height;

  elas.I1 = (uint8_t*)_mm_malloc((size_t)bpl * height, 16);
  elas.I2 = (uint8_t*)_mm_malloc((size_t)bpl * height, 16);
  memset(elas.I1, 0, (size_t)bpl * height);
  memset(elas.I2, 0, (size_t)bpl * height);
  for (int v = 0; v < height; v++) {
    memcpy(elas.I1 + (size_t)v * bpl, I1_ + (size_t)v * dims[2], width);
    memcpy(elas.I2 + (size_t)v * bpl, I2_ + (size_t)v * dims[2], width);
  }
  int rc = 0;
  {
    Descriptor desc1(elas.I1, width, height, bpl, param.subsampling);
    Descriptor desc2(elas.I2, width, height, bpl, param.subsampling);
    if (st->desc1) memcpy(st->desc1, desc1.I_desc, 16 * n);
    if (st->desc2) memcpy(st->desc2, desc2.I_desc, 16 * n);

    /* --- computeSupportMatches, replayed so D_can can be dumped --- */
    int step = param.candidate_stepsize;
    if (param.subsampling) step += step % 2;
    int wc = 0, hc = 0;
    for (int u = 0; u < width; u += step) wc++;
    for (int v = 0; v < height; v += step) hc++;
    int16_t* D_can = (int16_t*)calloc((size_t)wc * hc, sizeof(int16_t));
    for (int uc = 1; uc < wc; uc++) {
      int u = uc * step;
      for (int vc = 1; vc < hc; vc++) {
        int v = vc * step;
        D_can[vc * wc + uc] = -1;
        int16_t d = elas.computeMatchingDisparity(u, v, desc1.I_desc, desc2.I_desc, false);
        if (d >= 0) {
          int16_t d2 = elas.computeMatchingDisparity(u - d, v, desc1.I_desc, desc2.I_desc, true);
          if (d2 >= 0 && abs(d - d2) <= param.lr_threshold) D_can[vc * wc + uc] = d;
        }
      }
    }
    if (st->dcan_raw) memcpy(st->dcan_raw, D_can, sizeof(int16_t) * wc * hc);
    elas.removeInconsistentSupportPoints(D_can, wc, hc);
    if (st->dcan_incon) memcpy(st->dcan_incon, D_can, sizeof(int16_t) * wc * hc);
    elas.removeRedundantSupportPoints(D_can, wc, hc, 5, 1, true);
    elas.removeRedundantSupportPoints(D_can, wc, hc, 5, 1, false);
    if (st->dcan_final) memcpy(st->dcan_final, D_can, sizeof(int16_t) * wc * hc);
    vector<Elas::support_pt> p_support;
    for (int uc = 1; uc < wc; uc++)
      for (int vc = 1; vc < hc; vc++)
        if (D_can[vc * wc + uc] >= 0)
          p_support.push_back(Elas::support_pt(uc * step, vc * step, D_can[vc * wc + uc]));
    if (param.add_corners) elas.addCornerSupportPoints(p_support);
    free(D_can);

    st->n_support = (int32_t)p_support.size();
    if (st->support && st->n_support <= st->cap_support)
      for (size_t i = 0; i < p_support.size(); i++) {
        st->support[3 * i] = p_support[i].u;
        st->support[3 * i + 1] = p_support[i].v;
        st->support[3 * i + 2] = p_support[i].d;
      }

    if (p_support.size() < 3) {
      rc = 1;
    } else {
      vector<Elas::triangle> tri_1 = elas.computeDelaunayTriangulation(p_support, 0);
      vector<Elas::triangle> tri_2 = elas.computeDelaunayTriangulation(p_support, 1);
      elas.computeDisparityPlanes(p_support, tri_1, 0);
      elas.computeDisparityPlanes(p_support, tri_2, 1);
      st->n_tri1 = (int32_t)tri_1.size();
      st->n_tri2 = (int32_t)tri_2.size();
      for (int k = 0; k < 2; k++) {
        vector<Elas::triangle>& t = k ? tri_2 : tri_1;
        int32_t* to = k ? st->tri2 : st->tri1;
        float* po = k ? st->planes2 : st->planes1;
        if ((int)t.size() > st->cap_tri) continue;
        for (size_t i = 0; i < t.size(); i++) {
          if (to) { to[3 * i] = t[i].c1; to[3 * i + 1] = t[i].c2; to[3 * i + 2] = t[i].c3; }
          if (po) {
            po[6 * i] = t[i].t1a; po[6 * i + 1] = t[i].t1b; po[6 * i + 2] = t[i].t1c;
            po[6 * i + 3] = t[i].t2a; po[6 * i + 4] = t[i].t2b; po[6 * i + 5] = t[i].t2c;
          }
        }
      }
      int gw = (int)ceil((float)width / (float)param.grid_size);
      int gh = (int)ceil((float)height / (float)param.grid_size);
      int32_t gd[3] = {param.disp_max + 2, gw, gh};
      size_t gn = (size_t)(param.disp_max + 2) * gw * gh;
      int32_t* g1 = (int32_t*)calloc(gn, sizeof(int32_t));
      int32_t* g2 = (int32_t*)calloc(gn, sizeof(int32_t));
      elas.createGrid(p_support, g1, gd, 0);
      elas.createGrid(p_support, g2, gd, 1);
      if (st->grid1) memcpy(st->grid1, g1, gn * sizeof(int32_t));
      if (st->grid2) memcpy(st->grid2, g2, gn * sizeof(int32_t));

      float* D1 = (float*)calloc(n, sizeof(float));
      float* D2 = (float*)calloc(n, sizeof(float));
      elas.computeDisparity(p_support, tri_1, g1, gd, desc1.I_desc, desc2.I_desc, 0, D1);
      elas.computeDisparity(p_support, tri_2, g2, gd, desc1.I_desc, desc2.I_desc, 1, D2);
      put(st->D1_raw, D1, n);
      put(st->D2_raw, D2, n);
      ref_postprocess(p, width, height, D1, D2, st);
      free(D1); free(D2); free(g1); free(g2);
    }
  }
  _mm_free(elas.I1);
  _mm_free(elas.I2);
  return rc;
}

}  /* extern "C" */
