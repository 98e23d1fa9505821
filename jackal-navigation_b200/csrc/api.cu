// api.cu -- the C ABI (include/jn_elas.h) over the CUDA stages: parameter presets,
// workspace management, Elas::process sequencing (elas.cpp:32-151) for a batch of
// independent frames, the stage dump used by the parity tests, and the calibration
// YAML reader (point_cloud.cpp:530-538).
//
// There is no CPU implementation of any stage in this library.  If the CUDA device
// cannot be used every compute entry point fails with JN_ERR_CUDA.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <atomic>
#include "common.cuh"
#include "../../include/jn_elas_debug.h"

std::atomic<long long> g_jn_launches{0};
static thread_local char g_err[512] = "";

void jn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// step-wise post-processing (post.cu)
void post_lr(const Geo& g, int B, Workspace& ws, cudaStream_t s, bool need_right = true);
void post_segments(const Geo& g, int B, Workspace& ws, int side, cudaStream_t s);
void post_gap(const Geo& g, int B, Workspace& ws, int side, cudaStream_t s);
void post_mean(const Geo& g, int B, Workspace& ws, const float* in, float* tmp, float* out, size_t ostride, cudaStream_t s);
void post_median(const Geo& g, int B, Workspace& ws, const float* in, float* tmp, float* out, size_t ostride, cudaStream_t s);
void post_copy(const Geo& g, int B, Workspace& ws, const float* in, float* out, int32_t* status, cudaStream_t s);

int launch_support_match(const Geo& g, int B, Workspace& ws, cudaStream_t s);
constexpr int JN_MAX_PARTS = 4;
#ifndef JN_DEFAULT_STAGGER
#define JN_DEFAULT_STAGGER 1
#endif

// Device-side resources of one handle: workspace arenas, streams, events, single-frame staging.
// They outlive the handle: jn_elas_destroy parks them in a small per-process cache and the next
// jn_elas_create on the same device picks them up again, so the reference's call pattern -- a fresh
// `Elas elas(param)` for every frame (point_cloud.cpp:416-419) -- costs no cudaMalloc / cudaFree.
struct DevRes {
  int device;
  Geo g;            // geometry the workspace was built for
  Workspace ws;     // frame slots for a whole batch (part 0 in split mode)
  void* arena;      // one cudaMalloc
  size_t arena_bytes;
  Workspace wsx[JN_MAX_PARTS - 1];   // frame slots of parts 1.. (split mode)
  void* arenax[JN_MAX_PARTS - 1];
  size_t arenax_bytes[JN_MAX_PARTS - 1];
  cudaStream_t aux[JN_MAX_PARTS - 1];
  cudaEvent_t ev_fork, ev_join[JN_MAX_PARTS - 1], ev_stag[JN_MAX_PARTS - 1];
  cudaStream_t hi[JN_MAX_PARTS];                 // high-priority streams for the latency-bound stages
  cudaEvent_t ev_hi_in[JN_MAX_PARTS], ev_hi_out[JN_MAX_PARTS];
  // single-frame staging for the host-pointer entry point + its stream
  uint8_t* dI[2];
  float* dD[2];
  int32_t* dStatus;
  size_t stage_pixels, stage_bytes;
  cudaStream_t own;
  // host-batch pipeline (jn_stereo_scan_batch_host): double-buffered chunks
  uint8_t* cI[2][2];      // [buffer][image]
  float* cD[2];
  int32_t* cStatus[2];
  double* cRanges[2];
  jn_scan_meta* cMeta[2];
  uint8_t* cU8[2];
  size_t chunk_frames, chunk_img_bytes, chunk_pixels;
  unsigned long long submits;   // buffer parity of the next submission
  cudaStream_t s_in, s_out;
  cudaEvent_t ev_in[2], ev_done[2], ev_out[2], ev_copied[2];
  // rolling pipeline of the submit entry points: one persistent stream per lane (= sub-batch slot with its
  // own workspace); a lane's parts follow each other across submissions without any stream joining them
  cudaStream_t lane[JN_MAX_PARTS];
  cudaEvent_t ev_lstag[JN_MAX_PARTS];        // lane k's current part is past the stagger stage
  cudaEvent_t ev_ldone[2][JN_MAX_PARTS];     // lane k finished its part of the submission using buffer set b
  int lanes_busy;                            // a submit call queued work on the lanes since the last wait
  // optional per-stage timing with CUDA events on the launching stream
  cudaEvent_t ev[JN_PROFILE_STAGES + 1];
};

struct jn_elas {
  jn_elas_params p;
  int device;
  int parts;        // > 1: run a batch as `parts` sub-batches on as many streams so that the
                    // latency-bound kernels (Delaunay, support filter: a CTA or two per frame) of
                    // one part overlap the bandwidth-bound kernels of the others
  int profile;
  int stagger;      // >= 0: sub-batch k+1 starts when sub-batch k has finished this stage, so the parts
                    // are always in DIFFERENT stages and a latency-bound stage of one (Delaunay, support
                    // filter: one or two CTAs per frame) runs next to a throughput-bound stage of the other
  DevRes* r;
};

extern "C" const char* jn_last_error(void) { return g_err; }
extern "C" long long jn_launch_count(void) { return g_jn_launches.load(); }

extern "C" void jn_elas_params_default(jn_elas_params* p, int setting) {
  // Elas::parameters(setting), elas.h:87-144
  if (!p) return;
  p->disp_min = 0;
  p->disp_max = 255;
  p->support_texture = 10;
  p->candidate_stepsize = 5;
  p->incon_window_size = 5;
  p->incon_threshold = 5;
  p->incon_min_support = 5;
  p->grid_size = 20;
  p->beta = 0.02f;
  p->sigma = 1;
  p->lr_threshold = 2;
  p->speckle_sim_threshold = 1;
  p->speckle_size = 200;
  p->subsampling = 0;
  if (setting == JN_ROBOTICS) {
    p->support_threshold = 0.85f;
    p->add_corners = 0;
    p->gamma = 3;
    p->sradius = 2;
    p->match_texture = 1;
    p->ipol_gap_width = 3;
    p->filter_median = 0;
    p->filter_adaptive_mean = 1;
    p->postprocess_only_left = 1;
  } else {
    p->support_threshold = 0.95f;
    p->add_corners = 1;
    p->gamma = 5;
    p->sradius = 3;
    p->match_texture = 0;
    p->ipol_gap_width = 5000;
    p->filter_median = 1;
    p->filter_adaptive_mean = 0;
    p->postprocess_only_left = 0;
  }
}

// ---- device-resource cache ------------------------------------------------------------------
#include <mutex>
static std::mutex g_cache_mu;
static std::vector<DevRes*> g_cache;           // parked resources, most recently used last
constexpr size_t JN_CACHE_MAX = 4;

static void devres_free(DevRes* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  if (r->arena) cudaFree(r->arena);
  for (int k = 0; k < JN_MAX_PARTS - 1; k++) {
    if (r->arenax[k]) cudaFree(r->arenax[k]);
    if (r->aux[k]) { cudaStreamDestroy(r->aux[k]); cudaEventDestroy(r->ev_join[k]); cudaEventDestroy(r->ev_stag[k]); }
  }
  for (int k = 0; k < 2; k++) {
    cudaFree(r->dI[k]); cudaFree(r->dD[k]);
    cudaFree(r->cI[k][0]); cudaFree(r->cI[k][1]); cudaFree(r->cD[k]); cudaFree(r->cStatus[k]);
    cudaFree(r->cRanges[k]); cudaFree(r->cMeta[k]); cudaFree(r->cU8[k]);
    if (r->ev_in[k]) { cudaEventDestroy(r->ev_in[k]); cudaEventDestroy(r->ev_done[k]);
                       cudaEventDestroy(r->ev_out[k]); cudaEventDestroy(r->ev_copied[k]); }
  }
  cudaFree(r->dStatus);
  if (r->ev[0])
    for (int i = 0; i <= JN_PROFILE_STAGES; i++) cudaEventDestroy(r->ev[i]);
  if (r->ev_fork) cudaEventDestroy(r->ev_fork);
  for (int k = 0; k < JN_MAX_PARTS; k++)
    if (r->hi[k]) { cudaStreamDestroy(r->hi[k]); cudaEventDestroy(r->ev_hi_in[k]); cudaEventDestroy(r->ev_hi_out[k]); }
  for (int k = 0; k < JN_MAX_PARTS; k++)
    if (r->lane[k]) {
      cudaStreamDestroy(r->lane[k]); cudaEventDestroy(r->ev_lstag[k]);
      cudaEventDestroy(r->ev_ldone[0][k]); cudaEventDestroy(r->ev_ldone[1][k]);
    }
  if (r->own) cudaStreamDestroy(r->own);
  if (r->s_in) cudaStreamDestroy(r->s_in);
  if (r->s_out) cudaStreamDestroy(r->s_out);
  delete r;
}

static DevRes* devres_acquire(int device) {
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    for (size_t i = g_cache.size(); i-- > 0;)
      if (g_cache[i]->device == device) {
        DevRes* r = g_cache[i];
        g_cache.erase(g_cache.begin() + i);
        return r;
      }
  }
  DevRes* r = new DevRes();
  memset(r, 0, sizeof(*r));
  r->device = device;
  return r;
}

static void devres_release(DevRes* r) {
  DevRes* evict = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    g_cache.push_back(r);
    if (g_cache.size() > JN_CACHE_MAX) { evict = g_cache.front(); g_cache.erase(g_cache.begin()); }
  }
  devres_free(evict);
}

// Frees everything the cache holds (process shutdown, tests).
extern "C" void jn_cache_clear(void) {
  std::vector<DevRes*> all;
  {
    std::lock_guard<std::mutex> lk(g_cache_mu);
    all.swap(g_cache);
  }
  for (DevRes* r : all) devres_free(r);
}

extern "C" jn_elas* jn_elas_create(const jn_elas_params* p, int device) {
  if (!p) { jn_set_error("jn_elas_create: null parameters"); return nullptr; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
    jn_set_error("jn_elas_create: CUDA device %d not available (%d devices); this library has no CPU path", device, ndev);
    return nullptr;
  }
  jn_elas* e = new jn_elas();
  memset(e, 0, sizeof(*e));
  e->p = *p;
  e->device = device;
  const char* sp = getenv("JN_ELAS_SPLIT");   // number of sub-batches / streams, 1 = off
  e->parts = sp ? atoi(sp) : 2;
  if (e->parts < 1) e->parts = 1;
  if (e->parts > JN_MAX_PARTS) e->parts = JN_MAX_PARTS;
  const char* sg = getenv("JN_ELAS_STAGGER");   // stage index (0 descriptor .. 5 dense), -1 = parts in lockstep
  e->stagger = sg ? atoi(sg) : JN_DEFAULT_STAGGER;
  if (e->stagger >= JN_PROFILE_STAGES - 1) e->stagger = JN_PROFILE_STAGES - 2;
  e->r = devres_acquire(device);
  return e;
}

// The handle goes away, its device resources are parked for the next handle on this device
// (a fresh Elas per frame as in point_cloud.cpp:416-419 allocates nothing).  Work still queued on
// the caller's streams keeps using the buffers; like any reuse of a handle, the next user must be
// ordered after it (same stream, or a synchronisation in between).
extern "C" void jn_elas_destroy(jn_elas* e) {
  if (!e) return;
  if (e->r->lanes_busy) {   // submissions still rolling on the internal streams: nobody else can order after them
    cudaSetDevice(e->device);
    if (e->r->s_out) cudaStreamSynchronize(e->r->s_out);
    for (int k = 0; k < JN_MAX_PARTS; k++)
      if (e->r->lane[k]) cudaStreamSynchronize(e->r->lane[k]);
    e->r->lanes_busy = 0;
  }
  devres_release(e->r);
  delete e;
}

static int make_geo(const jn_elas_params& p, const int32_t dims[3], Geo* out) {
  Geo g;
  memset(&g, 0, sizeof(g));
  g.p = p;
  g.W = dims[0]; g.H = dims[1]; g.bpl = dims[2];
  if (g.W < 16 || g.H < 16 || g.bpl < g.W) { jn_set_error("bad dims %dx%d stride %d", g.W, g.H, g.bpl); return JN_ERR_ARG; }
  if (p.disp_max < 10 || p.disp_max > 4095 || p.disp_min > p.disp_max || p.candidate_stepsize < 1 ||
      p.grid_size < 1 || p.incon_window_size < 0 || p.incon_window_size > 16 || p.incon_threshold < 0 ||
      p.incon_threshold > 8191) {
    jn_set_error("parameter out of the supported range");
    return JN_ERR_ARG;
  }
  if (g.W >= 8192 || g.H >= 8192) { jn_set_error("image too large for the 64-bit exact predicates"); return JN_ERR_UNSUPPORTED; }
  // subsampling: only every second line has descriptors, an odd lattice step is bumped (elas.cpp:379-381)
  if (p.subsampling) g.p.candidate_stepsize += g.p.candidate_stepsize % 2;
  g.Wd = p.subsampling ? g.W / 2 : g.W;
  g.Hd = p.subsampling ? g.H / 2 : g.H;
  g.speckle_eff = p.subsampling ? (int)(sqrtf((float)p.speckle_size) * 2) : p.speckle_size;
  g.gap_eff = p.subsampling ? p.ipol_gap_width / 2 + 1 : p.ipol_gap_width;
  const int step = g.p.candidate_stepsize;
  g.Wc = (g.W + step - 1) / step;
  g.Hc = (g.H + step - 1) / step;
  g.gw = (int)ceilf((float)g.W / (float)p.grid_size);
  g.gh = (int)ceilf((float)g.H / (float)p.grid_size);
  g.gwords = ((p.disp_max + 1 + 127) / 128) * 4;
  g.gs_magic = (unsigned)((0x100000000ull + (unsigned)p.grid_size - 1) / (unsigned)p.grid_size);
  if (p.grid_size >= 32768) { jn_set_error("grid_size too large"); return JN_ERR_ARG; }
  g.cap_s = g.Wc * g.Hc + 8;
  g.cap_t = 2 * g.cap_s;
  // prior table and radius exactly as computeDisparity builds them (elas.cpp:802-806), float math
  float two_sigma_squared = 2 * p.sigma * p.sigma;
  g.plane_radius = (int)fmaxf(ceilf(p.sigma * p.sradius), 2.0f);
  if (g.plane_radius > 7) { jn_set_error("plane radius %d > 7 not supported", g.plane_radius); return JN_ERR_UNSUPPORTED; }
  for (int dd = 0; dd <= g.plane_radius; dd++)
    g.P[dd] = (int32_t)((-logf(p.gamma + expf(-dd * dd / two_sigma_squared)) + logf(p.gamma)) / p.beta);
  // bits of the d_plane field of a plane map entry: values -(r+1) .. disp_max + r + 1, biased by r + 1
  for (g.pm_dbits = 1; (1 << g.pm_dbits) < p.disp_max + 2 * (g.plane_radius + 1) + 1; g.pm_dbits++) {}
  for (int dd = 0; dd <= g.plane_radius; dd++)
    if (g.P[dd] < -2000 || g.P[dd] > 1900) { jn_set_error("prior table out of the packed-key range"); return JN_ERR_UNSUPPORTED; }
  *out = g;
  return JN_OK;
}

template <typename T>
static void carve(char*& cur, T*& ptr, size_t count) {
  ptr = reinterpret_cast<T*>(cur);
  size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
  cur += bytes;
}

static void layout(const Geo& g, int B, Workspace& ws, char* base) {
  char* cur = base;
  const size_t n = (size_t)g.W * g.H, np = (size_t)g.Wc * g.Hc, b = (size_t)B;
  const size_t gcw = (size_t)g.gw * g.gh * g.gwords;
  ws.B = B;
  for (int k = 0; k < 2; k++) carve(cur, ws.desc[k], b * n * 16);
  carve(cur, ws.dcan, b * np);
  carve(cur, ws.dcan_incon, b * np);
  carve(cur, ws.dcan_final, b * np);
  carve(cur, ws.cnt, b * np);
  carve(cur, ws.frontier, b * 2 * np);
  carve(cur, ws.sup, b * g.cap_s * 4);
  for (int k = 0; k < 2; k++) carve(cur, ws.px[k], b * g.cap_s);
  carve(cur, ws.py, b * g.cap_s);
  carve(cur, ws.occ, b * 2 * (size_t)g.W * g.Hc);
  for (int k = 0; k < 2; k++) {
    carve(cur, ws.xlist[k], b * g.cap_s);
    carve(cur, ws.ylist[k], b * g.cap_s);
    carve(cur, ws.tmpA[k], b * g.cap_s);
    carve(cur, ws.tmpB[k], b * g.cap_s);
    carve(cur, ws.tmpC[k], b * g.cap_s);
    carve(cur, ws.tmpD[k], b * g.cap_t);
    carve(cur, ws.nb[k], b * g.cap_t * 3);
    carve(cur, ws.vx[k], b * g.cap_t * 3);
    carve(cur, ws.nodeL[k], b * g.cap_t);
    carve(cur, ws.nodeR[k], b * g.cap_t);
    carve(cur, ws.tri[k], b * g.cap_t * 3);
    carve(cur, ws.planes[k], b * g.cap_t * 6);
    carve(cur, ws.gridtmp[k], b * gcw);
    carve(cur, ws.gridmask[k], b * gcw);
    carve(cur, ws.gridlist[k], b * (size_t)g.gw * g.gh * GRID_LIST);
    carve(cur, ws.trimap[k], b * n);
    carve(cur, ws.Draw[k], b * n);
    carve(cur, ws.Dlr[k], b * n);
    carve(cur, ws.Dtmp[k], b * n);
    carve(cur, ws.Dtmp2[k], b * n);
  }
  carve(cur, ws.label, b * n);
  carve(cur, ws.segsize, b * n);
  carve(cur, ws.info, b);
  ws.bytes = (size_t)(cur - base);
}

// Same buffer layout <=> same image size, lattice, grid and map resolution.
static bool same_layout(const Geo& a, const Geo& b) {
  return a.W == b.W && a.H == b.H && a.Wc == b.Wc && a.Hc == b.Hc && a.gw == b.gw && a.gh == b.gh &&
         a.gwords == b.gwords && a.Wd == b.Wd && a.Hd == b.Hd;
}

static int ensure_workspace(jn_elas* e, const int32_t dims[3], int B) {
  Geo g;
  int rc = make_geo(e->p, dims, &g);
  if (rc) return rc;
  JN_CUDA_CHECK(cudaSetDevice(e->device));
  DevRes* r = e->r;
  // sub-batches: parts 1.. take B / K frames each, part 0 the rest
  const int K = (e->parts < B) ? e->parts : B, Bx = (K > 1) ? B / K : 0;
  bool fits = r->arena && same_layout(r->g, g) && r->ws.B >= B;
  for (int k = 0; k + 1 < K; k++) fits = fits && r->wsx[k].B >= Bx;
  if (fits) {
    r->g = g;  // stride and parameters may differ between calls
    return JN_OK;
  }
  // (re)build: keep an arena that is large enough (zeroed again), else allocate
  Workspace probe;
  memset(&probe, 0, sizeof(probe));
  layout(g, B, probe, nullptr);
  if (r->arena_bytes < probe.bytes) {
    if (r->arena) cudaFree(r->arena);
    r->arena = nullptr; r->arena_bytes = 0;
    JN_CUDA_CHECK(cudaMalloc(&r->arena, probe.bytes));
    r->arena_bytes = probe.bytes;
  }
  JN_CUDA_CHECK(cudaMemset(r->arena, 0, probe.bytes));
  layout(g, B, r->ws, (char*)r->arena);
  for (int k = 0; k < JN_MAX_PARTS - 1; k++) memset(&r->wsx[k], 0, sizeof(r->wsx[k]));
  for (int k = 0; k + 1 < K; k++) {
    layout(g, Bx, probe, nullptr);
    if (r->arenax_bytes[k] < probe.bytes) {
      if (r->arenax[k]) cudaFree(r->arenax[k]);
      r->arenax[k] = nullptr; r->arenax_bytes[k] = 0;
      JN_CUDA_CHECK(cudaMalloc(&r->arenax[k], probe.bytes));
      r->arenax_bytes[k] = probe.bytes;
    }
    JN_CUDA_CHECK(cudaMemset(r->arenax[k], 0, probe.bytes));
    layout(g, Bx, r->wsx[k], (char*)r->arenax[k]);
    if (!r->aux[k]) {
      JN_CUDA_CHECK(cudaStreamCreateWithFlags(&r->aux[k], cudaStreamNonBlocking));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_join[k], cudaEventDisableTiming));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_stag[k], cudaEventDisableTiming));
    }
  }
  if (K > 1 && !r->ev_fork) JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_fork, cudaEventDisableTiming));
  r->g = g;
  return JN_OK;
}

// One stage of Elas::process (elas.cpp:57-140) for the B frames of one workspace.
static int run_stage(int stage, const Geo& g, int B, Workspace& ws, const uint8_t* I1, const uint8_t* I2, float* D1,
                     float* D2, int32_t* status, cudaStream_t s) {
  switch (stage) {
    case 0: launch_descriptor(g, B, I1, I2, ws, s); return JN_OK;
    case 1: return launch_support(g, B, ws, s);
    case 2: return launch_delaunay(g, B, ws, s);
    case 3: launch_planes_grid(g, B, ws, s); return JN_OK;
    case 4: launch_raster(g, B, ws, s); return JN_OK;
    case 5: launch_dense_match(g, B, ws, s); return JN_OK;
    default: launch_post(g, B, ws, D1, D2, status, s); return JN_OK;
  }
}

// Elas::process for B frames.  Single stream: everything on `s`.  Split mode: the batch is cut
// into K sub-batches; part 0 runs on `s` with e->r->ws, part k on auxiliary stream k with e->r->wsx[k-1],
// stage launches interleaved; the auxiliary streams fork from and join back into `s`.
static int run_pipeline(jn_elas* e, int B, const uint8_t* I1, const uint8_t* I2, float* D1, float* D2,
                        int32_t* status, cudaStream_t s) {
  const Geo& g = e->r->g;
  const bool prof = e->profile != 0;
  int K = (e->parts < B) ? e->parts : B;
  if (prof) K = 1;
  const int Bx = (K > 1) ? B / K : 0;
  for (int k = 0; k + 1 < K; k++)
    if (e->r->wsx[k].B < Bx) K = 1;
  const int B0 = B - (K - 1) * Bx;
  const size_t n = (size_t)g.Wd * g.Hd, ibytes = (size_t)g.bpl * g.H;
  if (K > 1) {
    JN_CUDA_CHECK(cudaEventRecord(e->r->ev_fork, s));
    for (int k = 0; k + 1 < K; k++) JN_CUDA_CHECK(cudaStreamWaitEvent(e->r->aux[k], e->r->ev_fork, 0));
  }
  int rc = JN_OK;
  if (K > 1 && e->stagger >= 0) {
    // staggered: part by part; part k+1 is released when part k is past stage `stagger`.  The
    // latency-bound stages (support filter, Delaunay, planes + grid: one or two CTAs per frame) go to a
    // high-priority stream of the part, so that their few CTAs are placed as soon as an SM frees up
    // instead of queueing behind the other part's large grids.
    int lo_p = 0, hi_p = 0;
    cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
    for (int k = 0; k < K; k++)
      if (!e->r->hi[k]) {
        JN_CUDA_CHECK(cudaStreamCreateWithPriority(&e->r->hi[k], cudaStreamNonBlocking, hi_p));
        JN_CUDA_CHECK(cudaEventCreateWithFlags(&e->r->ev_hi_in[k], cudaEventDisableTiming));
        JN_CUDA_CHECK(cudaEventCreateWithFlags(&e->r->ev_hi_out[k], cudaEventDisableTiming));
      }
    for (int k = 0; k < K && rc == JN_OK; k++) {
      cudaStream_t sk = k ? e->r->aux[k - 1] : s, hk = e->r->hi[k];
      Workspace& wk = k ? e->r->wsx[k - 1] : e->r->ws;
      const int Bk = k ? Bx : B0;
      const size_t f0 = k ? (size_t)B0 + (size_t)(k - 1) * Bx : 0;
      if (k > 0) JN_CUDA_CHECK(cudaStreamWaitEvent(sk, e->r->ev_stag[k - 1], 0));
      launch_descriptor(g, Bk, I1 + f0 * ibytes, I2 + f0 * ibytes, wk, sk);
      if (e->stagger == 0 && k + 1 < K) JN_CUDA_CHECK(cudaEventRecord(e->r->ev_stag[k], sk));
      rc = launch_support_match(g, Bk, wk, sk);
      if (rc) break;
      if (e->stagger == 1 && k + 1 < K) JN_CUDA_CHECK(cudaEventRecord(e->r->ev_stag[k], sk));
      JN_CUDA_CHECK(cudaEventRecord(e->r->ev_hi_in[k], sk));
      JN_CUDA_CHECK(cudaStreamWaitEvent(hk, e->r->ev_hi_in[k], 0));
      rc = launch_support_filter(g, Bk, wk, hk);
      if (rc) break;
      rc = launch_delaunay(g, Bk, wk, hk);
      if (rc) break;
      launch_planes_grid(g, Bk, wk, hk);
      JN_CUDA_CHECK(cudaEventRecord(e->r->ev_hi_out[k], hk));
      JN_CUDA_CHECK(cudaStreamWaitEvent(sk, e->r->ev_hi_out[k], 0));
      if (e->stagger >= 2 && e->stagger <= 3 && k + 1 < K) JN_CUDA_CHECK(cudaEventRecord(e->r->ev_stag[k], sk));
      launch_raster(g, Bk, wk, sk);
      if (e->stagger == 4 && k + 1 < K) JN_CUDA_CHECK(cudaEventRecord(e->r->ev_stag[k], sk));
      launch_dense_match(g, Bk, wk, sk);
      if (e->stagger >= 5 && k + 1 < K) JN_CUDA_CHECK(cudaEventRecord(e->r->ev_stag[k], sk));
      launch_post(g, Bk, wk, D1 + f0 * n, D2 ? D2 + f0 * n : nullptr, status ? status + f0 : nullptr, sk);
    }
  } else
  for (int stage = 0; stage < JN_PROFILE_STAGES && rc == JN_OK; stage++) {
    if (prof) cudaEventRecord(e->r->ev[stage], s);
    rc = run_stage(stage, g, B0, e->r->ws, I1, I2, D1, D2, status, s);
    for (int k = 0; k + 1 < K && rc == JN_OK; k++) {
      const size_t f0 = (size_t)B0 + (size_t)k * Bx;   // first frame of part k+1
      rc = run_stage(stage, g, Bx, e->r->wsx[k], I1 + f0 * ibytes, I2 + f0 * ibytes, D1 + f0 * n,
                     D2 ? D2 + f0 * n : nullptr, status ? status + f0 : nullptr, e->r->aux[k]);
    }
  }
  if (prof) cudaEventRecord(e->r->ev[JN_PROFILE_STAGES], s);
  // the auxiliary streams always join back into `s`, also after a failed stage launch
  for (int k = 0; k + 1 < K; k++) {
    JN_CUDA_CHECK(cudaEventRecord(e->r->ev_join[k], e->r->aux[k]));
    JN_CUDA_CHECK(cudaStreamWaitEvent(s, e->r->ev_join[k], 0));
  }
  if (rc) return rc;
  JN_CUDA_CHECK(cudaGetLastError());
  return JN_OK;
}

// Per-stage device times of the LAST profiled batch call (ms): descriptor, support,
// delaunay, planes+grid, raster, dense match, post-processing.  The caller synchronises.
extern "C" int jn_elas_profile(jn_elas* e, int enable) {
  if (!e) return JN_ERR_ARG;
  JN_CUDA_CHECK(cudaSetDevice(e->device));
  if (enable && !e->r->ev[0])
    for (int i = 0; i <= JN_PROFILE_STAGES; i++) JN_CUDA_CHECK(cudaEventCreate(&e->r->ev[i]));
  e->profile = enable;
  return JN_OK;
}
extern "C" int jn_elas_profile_read(jn_elas* e, float ms[JN_PROFILE_STAGES]) {
  if (!e || !e->r->ev[0]) return JN_ERR_ARG;
  JN_CUDA_CHECK(cudaEventSynchronize(e->r->ev[JN_PROFILE_STAGES]));
  for (int i = 0; i < JN_PROFILE_STAGES; i++) JN_CUDA_CHECK(cudaEventElapsedTime(&ms[i], e->r->ev[i], e->r->ev[i + 1]));
  return JN_OK;
}

extern "C" int jn_elas_process_batch(jn_elas* e, int n, const uint8_t* I1, const uint8_t* I2, float* D1, float* D2,
                                     int32_t* status, const int32_t dims[3], void* stream) {
  if (!e || n <= 0 || !I1 || !I2 || !D1 || !dims) { jn_set_error("jn_elas_process_batch: bad arguments"); return JN_ERR_ARG; }
  int rc = ensure_workspace(e, dims, n);
  if (rc) return rc;
  return run_pipeline(e, n, I1, I2, D1, D2, status, (cudaStream_t)stream);
}

static int ensure_staging(jn_elas* e, const int32_t dims[3]) {
  size_t npix = (size_t)dims[0] * dims[1], nbytes = (size_t)dims[2] * dims[1];
  if (e->r->stage_pixels >= npix && e->r->stage_bytes >= nbytes) return JN_OK;
  for (int k = 0; k < 2; k++) { cudaFree(e->r->dI[k]); cudaFree(e->r->dD[k]); e->r->dI[k] = nullptr; e->r->dD[k] = nullptr; }
  cudaFree(e->r->dStatus);
  e->r->dStatus = nullptr;
  for (int k = 0; k < 2; k++) {
    JN_CUDA_CHECK(cudaMalloc(&e->r->dI[k], nbytes));
    JN_CUDA_CHECK(cudaMalloc(&e->r->dD[k], npix * sizeof(float)));
  }
  JN_CUDA_CHECK(cudaMalloc(&e->r->dStatus, sizeof(int32_t)));
  e->r->stage_pixels = npix;
  e->r->stage_bytes = nbytes;
  return JN_OK;
}

// Host-batch path: the map slot of a frame that was not matched (<3 support points, rejected) holds
// whatever the staging buffer held before.  The reference's caller zeroes its maps right before the
// call (point_cloud.cpp:413-414) and finds them untouched = zero afterwards; give it exactly that.
__global__ void zero_unmatched_kernel(float* D, const int32_t* status, size_t pix) {
  if (status[blockIdx.y] == JN_OK) return;
  float4* d = reinterpret_cast<float4*>(D + (size_t)blockIdx.y * pix);
  const size_t n4 = pix / 4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    d[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (blockIdx.x == 0 && threadIdx.x < (pix & 3)) D[(size_t)blockIdx.y * pix + n4 * 4 + threadIdx.x] = 0.f;
}

static int ensure_own_stream(DevRes* r) {
  if (!r->own) JN_CUDA_CHECK(cudaStreamCreateWithFlags(&r->own, cudaStreamNonBlocking));
  return JN_OK;
}

// Elas::process: host pointers, synchronous.  Everything runs on the handle's own stream (not the
// legacy default stream).  D2 may be NULL: the right map is then neither post-processed further nor
// copied back (the reference's caller, point_cloud.cpp:419-421, never looks at it).  Pinned caller
// buffers (jn_host_alloc) make both copies DMA transfers; pageable ones go through the driver's staging.
extern "C" int jn_elas_process(jn_elas* e, const uint8_t* I1, const uint8_t* I2, float* D1, float* D2,
                               const int32_t dims[3]) {
  if (!e || !I1 || !I2 || !D1 || !dims) { jn_set_error("jn_elas_process: bad arguments"); return JN_ERR_ARG; }
  int rc = ensure_workspace(e, dims, 1);
  if (rc) return rc;
  rc = ensure_staging(e, dims);
  if (rc) return rc;
  DevRes* r = e->r;
  if ((rc = ensure_own_stream(r))) return rc;
  cudaStream_t s = r->own;
  const size_t npix = (size_t)r->g.Wd * r->g.Hd, nbytes = (size_t)dims[2] * dims[1];
  JN_CUDA_CHECK(cudaMemcpyAsync(r->dI[0], I1, nbytes, cudaMemcpyHostToDevice, s));
  JN_CUDA_CHECK(cudaMemcpyAsync(r->dI[1], I2, nbytes, cudaMemcpyHostToDevice, s));
  rc = run_pipeline(e, 1, r->dI[0], r->dI[1], r->dD[0], D2 ? r->dD[1] : nullptr, r->dStatus, s);
  if (rc) return rc;
  int32_t st = 0;
  JN_CUDA_CHECK(cudaMemcpyAsync(&st, r->dStatus, sizeof(st), cudaMemcpyDeviceToHost, s));
  JN_CUDA_CHECK(cudaStreamSynchronize(s));
  if (st < 0) jn_set_error("frame rejected on the device (status %d): parameter combination not supported", st);
  if (st == JN_OK) {  // "<3 support points": outputs untouched (elas.cpp:66-71)
    JN_CUDA_CHECK(cudaMemcpyAsync(D1, r->dD[0], npix * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (D2) JN_CUDA_CHECK(cudaMemcpyAsync(D2, r->dD[1], npix * sizeof(float), cudaMemcpyDeviceToHost, s));
    JN_CUDA_CHECK(cudaStreamSynchronize(s));
  }
  return st;
}

// Page-locked host memory for callers that do not link the CUDA runtime themselves.
extern "C" void* jn_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
    jn_set_error("jn_host_alloc(%zu): %s", bytes, cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}
extern "C" void jn_host_free(void* p) { if (p) cudaFreeHost(p); }

// ---- host-pointer batch: image pairs in, obstacle scans out -----------------------------------
// The per-frame sequence of point_cloud.cpp (generateDisparityMap :406-429 -> publishObstacleScan
// :213-296) for n independent frames with HOST buffers.  Copies and compute run on three streams
// with double-buffered device staging, so consecutive submissions overlap: H2D of call i+1, the
// kernels of call i and D2H of call i-1.  submit() returns as soon as the work is queued; the
// caller's buffers belong to the library until wait() returns.
static int ensure_chunks(jn_elas* e, const int32_t dims[3], int n) {
  DevRes* r = e->r;
  const size_t img = (size_t)dims[2] * dims[1], pix = (size_t)r->g.Wd * r->g.Hd;
  if (!r->s_in) {
    JN_CUDA_CHECK(cudaStreamCreateWithFlags(&r->s_in, cudaStreamNonBlocking));
    JN_CUDA_CHECK(cudaStreamCreateWithFlags(&r->s_out, cudaStreamNonBlocking));
    for (int k = 0; k < 2; k++) {
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_in[k], cudaEventDisableTiming));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_done[k], cudaEventDisableTiming));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_out[k], cudaEventDisableTiming));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_copied[k], cudaEventDisableTiming));
    }
  }
  if (r->chunk_frames >= (size_t)n && r->chunk_img_bytes == img && r->chunk_pixels == pix) return JN_OK;
  JN_CUDA_CHECK(cudaDeviceSynchronize());
  for (int k = 0; k < 2; k++) {
    cudaFree(r->cI[k][0]); cudaFree(r->cI[k][1]); cudaFree(r->cD[k]); cudaFree(r->cStatus[k]);
    cudaFree(r->cRanges[k]); cudaFree(r->cMeta[k]); cudaFree(r->cU8[k]);
    r->cI[k][0] = r->cI[k][1] = nullptr; r->cD[k] = nullptr; r->cStatus[k] = nullptr;
    r->cRanges[k] = nullptr; r->cMeta[k] = nullptr; r->cU8[k] = nullptr;
  }
  r->chunk_frames = 0;
  for (int k = 0; k < 2; k++) {
    JN_CUDA_CHECK(cudaMalloc(&r->cI[k][0], img * n));
    JN_CUDA_CHECK(cudaMalloc(&r->cI[k][1], img * n));
    JN_CUDA_CHECK(cudaMalloc(&r->cD[k], pix * n * sizeof(float)));
    JN_CUDA_CHECK(cudaMalloc(&r->cStatus[k], n * sizeof(int32_t)));
    JN_CUDA_CHECK(cudaMalloc(&r->cRanges[k], (size_t)n * JN_SCAN_BINS * sizeof(double)));
    JN_CUDA_CHECK(cudaMalloc(&r->cMeta[k], (size_t)n * sizeof(jn_scan_meta)));
    JN_CUDA_CHECK(cudaMalloc(&r->cU8[k], pix * n));
    // a never-used buffer is "free" and "read back": record its events once
    JN_CUDA_CHECK(cudaEventRecord(r->ev_done[k], r->own));
    JN_CUDA_CHECK(cudaEventRecord(r->ev_copied[k], r->own));
  }
  r->chunk_frames = n; r->chunk_img_bytes = img; r->chunk_pixels = pix;
  return JN_OK;
}

struct jn_scan;
int jn_scan_batch_at(jn_scan* s, int acc_frame0, int n, const float* D, double* ranges, jn_scan_meta* meta,
                     uint8_t* dmap_u8, cudaStream_t st);
int jn_scan_reserve(jn_scan* s, int n);

static int ensure_lanes(DevRes* r, int K) {
  int lo_p = 0, hi_p = 0;
  cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
  for (int k = 0; k < K; k++) {
    if (!r->lane[k]) {
      JN_CUDA_CHECK(cudaStreamCreateWithFlags(&r->lane[k], cudaStreamNonBlocking));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_lstag[k], cudaEventDisableTiming));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_ldone[0][k], cudaEventDisableTiming));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_ldone[1][k], cudaEventDisableTiming));
    }
    if (!r->hi[k]) {
      JN_CUDA_CHECK(cudaStreamCreateWithPriority(&r->hi[k], cudaStreamNonBlocking, hi_p));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_hi_in[k], cudaEventDisableTiming));
      JN_CUDA_CHECK(cudaEventCreateWithFlags(&r->ev_hi_out[k], cudaEventDisableTiming));
    }
  }
  return JN_OK;
}

// One submission of n frames (device buffers) as K parts on the K lanes.  Lane k waits for `ready`
// (inputs there, output buffers free), for its own previous part (stream order) and for the previous
// part in the rolling order -- lane k-1 of this submission, or the last lane of the previous one -- to be
// past support matching: at any time the parts in flight are in DIFFERENT stages, so the latency-bound
// stages of one (support filter, Delaunay, planes + grid: a CTA or two per frame, on a high-priority
// stream) run next to a throughput-bound stage of another, and -- unlike a fork/join on the caller's
// stream -- the tail of one submission overlaps the head of the next.  ELAS, the zero map of an
// unmatched frame and the scan of a part all run on its lane; ev_ldone[b][k] marks its end.
static int run_rolling(jn_elas* e, jn_scan* sc, int n, int b, const uint8_t* I1, const uint8_t* I2, float* D,
                       int32_t* status, double* ranges, jn_scan_meta* meta, uint8_t* u8, cudaEvent_t* ready,
                       int n_ready) {
  DevRes* r = e->r;
  const Geo& g = r->g;
  const int K = (e->parts < n) ? e->parts : n;
  const int Bx = (K > 1) ? n / K : 0, B0 = n - (K - 1) * Bx;
  const size_t pix = (size_t)g.Wd * g.Hd, ibytes = (size_t)g.bpl * g.H;
  int rc = ensure_lanes(r, K);
  if (rc) return rc;
  if ((rc = jn_scan_reserve(sc, n))) return rc;
  r->lanes_busy = 1;
  for (int k = 0; k < K; k++) {
    cudaStream_t L = r->lane[k], hk = r->hi[k];
    Workspace& wk = k ? r->wsx[k - 1] : r->ws;
    const int Bk = k ? Bx : B0;
    const size_t f0 = k ? (size_t)B0 + (size_t)(k - 1) * Bx : 0;
    for (int i = 0; i < n_ready; i++) JN_CUDA_CHECK(cudaStreamWaitEvent(L, ready[i], 0));
    if (K > 1) JN_CUDA_CHECK(cudaStreamWaitEvent(L, r->ev_lstag[(k + K - 1) % K], 0));   // never recorded yet: no-op
    launch_descriptor(g, Bk, I1 + f0 * ibytes, I2 + f0 * ibytes, wk, L);
    if ((rc = launch_support_match(g, Bk, wk, L))) return rc;
    JN_CUDA_CHECK(cudaEventRecord(r->ev_lstag[k], L));
    JN_CUDA_CHECK(cudaEventRecord(r->ev_hi_in[k], L));
    JN_CUDA_CHECK(cudaStreamWaitEvent(hk, r->ev_hi_in[k], 0));
    if ((rc = launch_support_filter(g, Bk, wk, hk))) return rc;
    if ((rc = launch_delaunay(g, Bk, wk, hk))) return rc;
    launch_planes_grid(g, Bk, wk, hk);
    JN_CUDA_CHECK(cudaEventRecord(r->ev_hi_out[k], hk));
    JN_CUDA_CHECK(cudaStreamWaitEvent(L, r->ev_hi_out[k], 0));
    launch_raster(g, Bk, wk, L);
    launch_dense_match(g, Bk, wk, L);
    launch_post(g, Bk, wk, D + f0 * pix, nullptr, status + f0, L);
    zero_unmatched_kernel<<<dim3(64, Bk), 256, 0, L>>>(D + f0 * pix, status + f0, pix);
    g_jn_launches += 1;
    if ((rc = jn_scan_batch_at(sc, (int)f0, Bk, D + f0 * pix, ranges + f0 * JN_SCAN_BINS, meta + f0,
                               u8 ? u8 + f0 * pix : nullptr, L)))
      return rc;
    JN_CUDA_CHECK(cudaEventRecord(r->ev_ldone[b][k], L));
  }
  JN_CUDA_CHECK(cudaGetLastError());
  return K;
}

extern "C" int jn_stereo_scan_submit(jn_elas* e, jn_scan* sc, int n, const uint8_t* I1, const uint8_t* I2,
                                     const int32_t dims[3], float* D1, int32_t* status, double* ranges,
                                     jn_scan_meta* meta, uint8_t* dmap_u8) {
  if (!e || !sc || n <= 0 || !I1 || !I2 || !dims || !ranges || !meta) {
    jn_set_error("jn_stereo_scan_submit: bad arguments");
    return JN_ERR_ARG;
  }
  if (e->p.subsampling) { jn_set_error("jn_stereo_scan_submit: the obstacle scan takes full-resolution maps"); return JN_ERR_UNSUPPORTED; }
  int rc = ensure_workspace(e, dims, n);
  if (rc) return rc;
  DevRes* r = e->r;
  if ((rc = ensure_own_stream(r))) return rc;
  if ((rc = ensure_chunks(e, dims, n))) return rc;
  const int k = (int)(r->submits++ & 1);
  const size_t img = (size_t)dims[2] * dims[1], pix = (size_t)r->g.Wd * r->g.Hd;
  // H2D of this call's frames as soon as buffer set k is free again (the lanes of call i-2 are done with it)
  JN_CUDA_CHECK(cudaStreamWaitEvent(r->s_in, r->ev_done[k], 0));
  JN_CUDA_CHECK(cudaMemcpyAsync(r->cI[k][0], I1, img * n, cudaMemcpyHostToDevice, r->s_in));
  JN_CUDA_CHECK(cudaMemcpyAsync(r->cI[k][1], I2, img * n, cudaMemcpyHostToDevice, r->s_in));
  JN_CUDA_CHECK(cudaEventRecord(r->ev_in[k], r->s_in));
  // compute: inputs arrived, outputs of call i-2 already read back
  cudaEvent_t ready[2] = {r->ev_in[k], r->ev_copied[k]};
  const int K = run_rolling(e, sc, n, k, r->cI[k][0], r->cI[k][1], r->cD[k], r->cStatus[k], r->cRanges[k], r->cMeta[k],
                            dmap_u8 ? r->cU8[k] : nullptr, ready, 2);
  if (K < 0) return K;
  // the staging buffers are free / the results complete when every lane is done with its part: the copy
  // stream collects the lanes, nothing else does
  for (int j = 0; j < K; j++) JN_CUDA_CHECK(cudaStreamWaitEvent(r->s_out, r->ev_ldone[k][j], 0));
  JN_CUDA_CHECK(cudaEventRecord(r->ev_done[k], r->s_out));
  JN_CUDA_CHECK(cudaMemcpyAsync(ranges, r->cRanges[k], (size_t)n * JN_SCAN_BINS * sizeof(double), cudaMemcpyDeviceToHost, r->s_out));
  JN_CUDA_CHECK(cudaMemcpyAsync(meta, r->cMeta[k], (size_t)n * sizeof(jn_scan_meta), cudaMemcpyDeviceToHost, r->s_out));
  if (status) JN_CUDA_CHECK(cudaMemcpyAsync(status, r->cStatus[k], n * sizeof(int32_t), cudaMemcpyDeviceToHost, r->s_out));
  if (dmap_u8) JN_CUDA_CHECK(cudaMemcpyAsync(dmap_u8, r->cU8[k], pix * n, cudaMemcpyDeviceToHost, r->s_out));
  if (D1) JN_CUDA_CHECK(cudaMemcpyAsync(D1, r->cD[k], pix * n * sizeof(float), cudaMemcpyDeviceToHost, r->s_out));
  JN_CUDA_CHECK(cudaEventRecord(r->ev_copied[k], r->s_out));
  return JN_OK;
}

// The same rolling pipeline for frames that already live in DEVICE memory: no copies, no caller stream.
// I1/I2 must be complete when the call is made and -- like every buffer passed here -- stay untouched until
// jn_stereo_scan_wait returns; submissions in flight need their own output buffers.  D1 (n*W*H floats),
// status (n), ranges (n*90), meta (n) are required, dmap_u8 (n*W*H) is optional.
extern "C" int jn_stereo_scan_submit_device(jn_elas* e, jn_scan* sc, int n, const uint8_t* I1, const uint8_t* I2,
                                            const int32_t dims[3], float* D1, int32_t* status, double* ranges,
                                            jn_scan_meta* meta, uint8_t* dmap_u8) {
  if (!e || !sc || n <= 0 || !I1 || !I2 || !dims || !D1 || !status || !ranges || !meta) {
    jn_set_error("jn_stereo_scan_submit_device: bad arguments");
    return JN_ERR_ARG;
  }
  if (e->p.subsampling) { jn_set_error("jn_stereo_scan_submit_device: the obstacle scan takes full-resolution maps"); return JN_ERR_UNSUPPORTED; }
  int rc = ensure_workspace(e, dims, n);
  if (rc) return rc;
  DevRes* r = e->r;
  if ((rc = ensure_own_stream(r))) return rc;
  const int k = (int)(r->submits++ & 1);
  const int K = run_rolling(e, sc, n, k, I1, I2, D1, status, ranges, meta, dmap_u8, nullptr, 0);
  return K < 0 ? K : JN_OK;
}

// Blocks until every submitted batch has landed in the caller's buffers.
extern "C" int jn_stereo_scan_wait(jn_elas* e) {
  if (!e) return JN_ERR_ARG;
  JN_CUDA_CHECK(cudaSetDevice(e->device));
  if (e->r->s_out) JN_CUDA_CHECK(cudaStreamSynchronize(e->r->s_out));
  for (int k = 0; k < JN_MAX_PARTS; k++)
    if (e->r->lane[k]) JN_CUDA_CHECK(cudaStreamSynchronize(e->r->lane[k]));
  if (e->r->own) JN_CUDA_CHECK(cudaStreamSynchronize(e->r->own));
  e->r->lanes_busy = 0;
  return JN_OK;
}

extern "C" int jn_stereo_scan_batch_host(jn_elas* e, jn_scan* sc, int n, const uint8_t* I1, const uint8_t* I2,
                                         const int32_t dims[3], float* D1, int32_t* status, double* ranges,
                                         jn_scan_meta* meta, uint8_t* dmap_u8) {
  int rc = jn_stereo_scan_submit(e, sc, n, I1, I2, dims, D1, status, ranges, meta, dmap_u8);
  if (rc) return rc;
  return jn_stereo_scan_wait(e);
}

// ------------------------------------------------------------------------------------------
// stage dump
template <typename T>
static int d2h(T* dst, const T* src, size_t count) {
  if (!dst) return JN_OK;
  JN_CUDA_CHECK(cudaMemcpy(dst, src, count * sizeof(T), cudaMemcpyDeviceToHost));
  return JN_OK;
}

static int dump_grid(const Geo& g, const uint32_t* dmask, int32_t* out) {
  if (!out) return JN_OK;
  size_t cells = (size_t)g.gw * g.gh;
  std::vector<uint32_t> m(cells * g.gwords);
  JN_CUDA_CHECK(cudaMemcpy(m.data(), dmask, m.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  const int stride = g.p.disp_max + 2;
  memset(out, 0, cells * stride * sizeof(int32_t));
  for (size_t c = 0; c < cells; c++) {
    int k = 0;
    for (int d = 0; d <= g.p.disp_max; d++)
      if (m[c * g.gwords + (d >> 5)] >> (d & 31) & 1u) out[c * stride + (++k)] = d;
    out[c * stride] = k;
  }
  return JN_OK;
}

// diagnostics: raw FrameInfo of frame slot `frame` after the last call (caller synchronises)
extern "C" int jn_elas_frameinfo(jn_elas* e, int frame, void* out, int bytes) {
  if (!e || !e->r->arena || frame < 0 || frame >= e->r->ws.B || bytes > (int)sizeof(FrameInfo)) return JN_ERR_ARG;
  JN_CUDA_CHECK(cudaMemcpy(out, e->r->ws.info + frame, bytes, cudaMemcpyDeviceToHost));
  return JN_OK;
}

// Runs the post-processing chain on ws.Draw of frame slot 0 step by step, copying every stage out.
static int dump_post(jn_elas* e, jn_stage_dump* o) {
  const Geo& g = e->r->g;
  Workspace& ws = e->r->ws;
  const size_t n = (size_t)g.Wd * g.Hd;
  cudaStream_t s = 0;
  int rc;
  const int sides = g.p.postprocess_only_left ? 1 : 2;
  post_lr(g, 1, ws, s);
  if ((rc = d2h(o->D1_lr, ws.Dlr[0], n))) return rc;
  if ((rc = d2h(o->D2_lr, ws.Dlr[1], n))) return rc;
  for (int k = 0; k < sides; k++) post_segments(g, 1, ws, k, s);
  if ((rc = d2h(o->D1_seg, ws.Dlr[0], n))) return rc;
  if ((rc = d2h(o->D2_seg, ws.Dlr[1], n))) return rc;
  for (int k = 0; k < sides; k++) post_gap(g, 1, ws, k, s);
  if ((rc = d2h(o->D1_gap, ws.Dlr[0], n))) return rc;
  if ((rc = d2h(o->D2_gap, ws.Dlr[1], n))) return rc;
  const float* cur[2] = {ws.Dlr[0], ws.Dlr[1]};
  if (g.p.filter_adaptive_mean)
    for (int k = 0; k < sides; k++) {
      post_mean(g, 1, ws, cur[k], ws.Dtmp[k], ws.Dtmp2[k], n, s);
      cur[k] = ws.Dtmp2[k];
    }
  if ((rc = d2h(o->D1_mean, cur[0], n))) return rc;
  if ((rc = d2h(o->D2_mean, cur[1], n))) return rc;
  if (g.p.filter_median)
    for (int k = 0; k < sides; k++) {
      post_median(g, 1, ws, cur[k], ws.Dtmp[k], e->r->dD[k], n, s);
      cur[k] = e->r->dD[k];
    }
  if ((rc = d2h(o->D1, cur[0], n))) return rc;
  if ((rc = d2h(o->D2, cur[1], n))) return rc;
  JN_CUDA_CHECK(cudaGetLastError());
  return JN_OK;
}


// ---- single stages with injected inputs (randomised parity tests) ---------------------------------
// Support filtering + compaction on a caller-supplied candidate image (Hc x Wc int16).
extern "C" int jn_debug_support_filter(jn_elas* e, const int16_t* dcan, const int32_t dims[3], int16_t* out_incon,
                                       int16_t* out_final, int32_t* support, int32_t cap_support, int32_t* n_support,
                                       int32_t* rounds) {
  if (!e || !dcan || !dims) return JN_ERR_ARG;
  int rc = ensure_workspace(e, dims, 1);
  if (rc) return rc;
  const Geo& g = e->r->g;
  Workspace& ws = e->r->ws;
  const size_t np = (size_t)g.Wc * g.Hc;
  JN_CUDA_CHECK(cudaMemcpy(ws.dcan, dcan, np * sizeof(int16_t), cudaMemcpyHostToDevice));
  rc = launch_support_filter(g, 1, ws, 0);
  if (rc) return rc;
  JN_CUDA_CHECK(cudaDeviceSynchronize());
  if ((rc = d2h(out_incon, ws.dcan_incon, np))) return rc;
  if ((rc = d2h(out_final, ws.dcan_final, np))) return rc;
  FrameInfo info;
  JN_CUDA_CHECK(cudaMemcpy(&info, ws.info, sizeof(info), cudaMemcpyDeviceToHost));
  if (n_support) *n_support = info.n_support;
  if (rounds) *rounds = info.incon_rounds;
  if (support && info.n_support <= cap_support) {
    std::vector<int32_t> s4((size_t)info.n_support * 4);
    JN_CUDA_CHECK(cudaMemcpy(s4.data(), ws.sup, s4.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < info.n_support; i++)
      for (int k = 0; k < 3; k++) support[3 * i + k] = s4[4 * (size_t)i + k];
  }
  return JN_OK;
}

// Delaunay triangulation of caller-supplied integer points (x,y pairs, 0 <= x,y < 8192).
extern "C" int jn_debug_triangulate(jn_elas* e, const int32_t* xy, int n, const int32_t dims[3], int32_t* tri,
                                    int32_t cap_tri, int32_t* n_tri) {
  if (!e || !xy || !dims || n < 0) return JN_ERR_ARG;
  int rc = ensure_workspace(e, dims, 1);
  if (rc) return rc;
  const Geo& g = e->r->g;
  Workspace& ws = e->r->ws;
  if (n > g.cap_s) return JN_ERR_ARG;
  std::vector<int32_t> x(n), y(n);
  for (int i = 0; i < n; i++) { x[i] = xy[2 * i]; y[i] = xy[2 * i + 1]; }
  JN_CUDA_CHECK(cudaMemcpy(ws.px[0], x.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
  JN_CUDA_CHECK(cudaMemcpy(ws.px[1], x.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
  JN_CUDA_CHECK(cudaMemcpy(ws.py, y.data(), n * sizeof(int32_t), cudaMemcpyHostToDevice));
  FrameInfo info;
  memset(&info, 0, sizeof(info));
  info.n_support = n;
  info.status = JN_OK;
  JN_CUDA_CHECK(cudaMemcpy(ws.info, &info, sizeof(info), cudaMemcpyHostToDevice));
  rc = launch_delaunay(g, 1, ws, 0);
  if (rc) return rc;
  JN_CUDA_CHECK(cudaDeviceSynchronize());
  JN_CUDA_CHECK(cudaMemcpy(&info, ws.info, sizeof(info), cudaMemcpyDeviceToHost));
  if (info.status != JN_OK) return info.status;
  *n_tri = info.n_tri[0];
  if (info.n_tri[0] <= cap_tri) return d2h(tri, ws.tri[0], (size_t)info.n_tri[0] * 3);
  return JN_OK;
}

// Post-processing chain (elas.cpp:108-140) on caller-supplied raw disparity maps.
extern "C" int jn_debug_postprocess(jn_elas* e, const float* D1raw, const float* D2raw, const int32_t dims[3],
                                    jn_stage_dump* o) {
  if (!e || !D1raw || !D2raw || !dims || !o) return JN_ERR_ARG;
  int rc = ensure_workspace(e, dims, 1);
  if (rc) return rc;
  rc = ensure_staging(e, dims);
  if (rc) return rc;
  const Geo& g = e->r->g;
  Workspace& ws = e->r->ws;
  const size_t n = (size_t)g.Wd * g.Hd;
  JN_CUDA_CHECK(cudaMemcpy(ws.Draw[0], D1raw, n * sizeof(float), cudaMemcpyHostToDevice));
  JN_CUDA_CHECK(cudaMemcpy(ws.Draw[1], D2raw, n * sizeof(float), cudaMemcpyHostToDevice));
  FrameInfo info;
  memset(&info, 0, sizeof(info));
  info.status = JN_OK;
  JN_CUDA_CHECK(cudaMemcpy(ws.info, &info, sizeof(info), cudaMemcpyHostToDevice));
  return dump_post(e, o);
}

extern "C" int jn_elas_stages(jn_elas* e, const uint8_t* I1, const uint8_t* I2, const int32_t dims[3],
                              jn_stage_dump* o) {
  if (!e || !I1 || !I2 || !dims || !o) return JN_ERR_ARG;
  int rc = ensure_workspace(e, dims, 1);
  if (rc) return rc;
  rc = ensure_staging(e, dims);
  if (rc) return rc;
  const Geo& g = e->r->g;
  Workspace& ws = e->r->ws;
  const size_t n = (size_t)g.W * g.H, np = (size_t)g.Wc * g.Hc, nbytes = (size_t)dims[2] * dims[1];
  JN_CUDA_CHECK(cudaMemcpy(e->r->dI[0], I1, nbytes, cudaMemcpyHostToDevice));
  JN_CUDA_CHECK(cudaMemcpy(e->r->dI[1], I2, nbytes, cudaMemcpyHostToDevice));
  cudaStream_t s = 0;
  launch_descriptor(g, 1, e->r->dI[0], e->r->dI[1], ws, s);
  rc = launch_support(g, 1, ws, s);
  if (rc) return rc;
  JN_CUDA_CHECK(cudaDeviceSynchronize());
  if ((rc = d2h(o->desc1, ws.desc[0], n * 16))) return rc;
  if ((rc = d2h(o->desc2, ws.desc[1], n * 16))) return rc;
  if ((rc = d2h(o->dcan_raw, ws.dcan, np))) return rc;
  if ((rc = d2h(o->dcan_incon, ws.dcan_incon, np))) return rc;
  if ((rc = d2h(o->dcan_final, ws.dcan_final, np))) return rc;
  FrameInfo info;
  JN_CUDA_CHECK(cudaMemcpy(&info, ws.info, sizeof(info), cudaMemcpyDeviceToHost));
  o->n_support = info.n_support;
  if (o->support && info.n_support <= o->cap_support) {
    std::vector<int32_t> s4((size_t)info.n_support * 4);
    JN_CUDA_CHECK(cudaMemcpy(s4.data(), ws.sup, s4.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
    for (int i = 0; i < info.n_support; i++)
      for (int k = 0; k < 3; k++) o->support[3 * i + k] = s4[4 * (size_t)i + k];
  }
  o->n_tri1 = o->n_tri2 = 0;
  if (info.status != JN_OK) return JN_FEW_SUPPORT;

  rc = launch_delaunay(g, 1, ws, s);
  if (rc) return rc;
  launch_planes_grid(g, 1, ws, s);
  JN_CUDA_CHECK(cudaDeviceSynchronize());
  JN_CUDA_CHECK(cudaMemcpy(&info, ws.info, sizeof(info), cudaMemcpyDeviceToHost));
  o->n_tri1 = info.n_tri[0];
  o->n_tri2 = info.n_tri[1];
  for (int k = 0; k < 2; k++) {
    int nt = info.n_tri[k];
    if (nt > o->cap_tri) continue;
    if ((rc = d2h(k ? o->tri2 : o->tri1, ws.tri[k], (size_t)nt * 3))) return rc;
    if ((rc = d2h(k ? o->planes2 : o->planes1, ws.planes[k], (size_t)nt * 6))) return rc;
  }
  if ((rc = dump_grid(g, ws.gridmask[0], o->grid1))) return rc;
  if ((rc = dump_grid(g, ws.gridmask[1], o->grid2))) return rc;

  launch_dense(g, 1, ws, s);
  JN_CUDA_CHECK(cudaDeviceSynchronize());
  if ((rc = d2h(o->D1_raw, ws.Draw[0], (size_t)g.Wd * g.Hd))) return rc;
  if ((rc = d2h(o->D2_raw, ws.Draw[1], (size_t)g.Wd * g.Hd))) return rc;

  return dump_post(e, o);
}

// ------------------------------------------------------------------------------------------
// Calibration YAML (OpenCV FileStorage subset: "name: !!opencv-matrix ... data: [ ... ]" and
// "name: [ ... ]"), point_cloud.cpp:530-538.
static bool yaml_numbers(const std::string& txt, const char* key, std::vector<double>& out) {
  out.clear();
  std::string k = std::string(key) + ":";
  size_t pos = 0;
  while (true) {
    pos = txt.find(k, pos);
    if (pos == std::string::npos) return false;
    if (pos == 0 || txt[pos - 1] == '\n') break;   // key must start a line
    pos += k.size();
  }
  size_t next_key = txt.size();
  // the block ends at the next top-level key (a line that starts with a letter)
  for (size_t i = txt.find('\n', pos); i != std::string::npos && i + 1 < txt.size(); i = txt.find('\n', i + 1))
    if (isalpha((unsigned char)txt[i + 1])) { next_key = i + 1; break; }
  std::string block = txt.substr(pos + k.size(), next_key - pos - k.size());
  size_t dpos = block.find("data:");
  size_t lb = block.find('[', dpos == std::string::npos ? 0 : dpos);
  size_t rb = block.find(']', lb == std::string::npos ? 0 : lb);
  if (lb == std::string::npos || rb == std::string::npos) return false;
  std::string nums = block.substr(lb + 1, rb - lb - 1);
  const char* c = nums.c_str();
  while (*c) {
    while (*c && (*c == ',' || isspace((unsigned char)*c))) c++;
    if (!*c) break;
    char* end = nullptr;
    double v = strtod(c, &end);
    if (end == c) return false;
    out.push_back(v);
    c = end;
  }
  return true;
}

extern "C" int jn_calib_load_yaml(const char* path, jn_calib* c) {
  if (!path || !c) return JN_ERR_ARG;
  FILE* f = fopen(path, "rb");
  if (!f) { jn_set_error("cannot open %s", path); return JN_ERR_IO; }
  std::string txt;
  char buf[4096];
  size_t r;
  while ((r = fread(buf, 1, sizeof(buf), f)) > 0) txt.append(buf, r);
  fclose(f);
  memset(c, 0, sizeof(*c));
  struct { const char* key; double* dst; size_t n; } items[] = {
      {"K1", c->K1, 9}, {"K2", c->K2, 9}, {"D1", c->D1, 5}, {"D2", c->D2, 5},
      {"R", c->R, 9},   {"T", c->T, 3},   {"XR", c->XR, 9}, {"XT", c->XT, 3}};
  std::vector<double> v;
  for (auto& it : items) {
    if (!yaml_numbers(txt, it.key, v) || v.size() != it.n) {
      jn_set_error("calibration file %s: key %s missing or wrong size", path, it.key);
      return JN_ERR_IO;
    }
    for (size_t i = 0; i < it.n; i++) it.dst[i] = v[i];
  }
  return JN_OK;
}

extern "C" void jn_calib_set_q(jn_calib* c, double cx, double cy, double f, double tx) {
  if (!c) return;
  memset(c->Q, 0, sizeof(c->Q));
  c->Q[0] = 1;  c->Q[3] = -cx;
  c->Q[5] = 1;  c->Q[7] = -cy;
  c->Q[11] = f;
  c->Q[14] = -1.0 / tx;
  c->has_q = 1;
}
