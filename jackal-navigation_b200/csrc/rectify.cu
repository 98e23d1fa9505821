// rectify.cu -- per-frame undistortion + rectification + ROI crop ahead of the stereo path.
//
// Replaces cv::remap(tmp, leftim, lmapx, lmapy, INTER_LINEAR) and the Rect crop that follows
// (point_cloud.cpp:440-442, 481-483).  OpenCV's remap for 8-bit images is fixed point
// (imgwarp.cpp: remap -> remapBilinear, initInterTab2D): the float map entry is rounded to 1/32
// pixel with cvRound (half to even), split into an int16 pixel coordinate and a 5+5-bit
// fraction, the four bilinear weights are 15-bit integers (32 - fy)(32 - fx) * 32 ... that sum to
// 32768, and the result is (sum + 2^14) >> 15.  The split is done once per camera at create
// time; the kernel is a pure gather: 8 bytes of table + 4 source bytes (cached) + 1 byte out per
// pixel.  Roofline: HBM, 9 bytes per output pixel + the source frame once.
#include <math.h>
#include <vector>
#include "common.cuh"

struct jn_rectify {
  int device;
  int map_w, map_h;
  uint2* table;   // per map pixel: x = (uint16)ix | (uint16)iy << 16, y = fy * 32 + fx
};

namespace {

__global__ void __launch_bounds__(256) remap_kernel(const uint2* __restrict__ table, int map_w, int rx, int ry, int rw,
                                                    int rh, const uint8_t* __restrict__ src, int sw, int sh, int sstride,
                                                    uint8_t* __restrict__ dst, int dstride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, frame = blockIdx.z;
  if (x >= rw) return;
  const uint2 e = __ldg(table + (size_t)(ry + y) * map_w + (rx + x));
  const int sx = (int)(short)(e.x & 0xFFFFu), sy = (int)(short)(e.x >> 16);
  const int fx = (int)(e.y & 31u), fy = (int)(e.y >> 5);
  // initInterTab2D: exact products; (0,0) saturates to 32767 and the fix-up gives the last weight 1
  int w0 = (32 - fy) * (32 - fx) * 32, w1 = (32 - fy) * fx * 32, w2 = fy * (32 - fx) * 32, w3 = fy * fx * 32;
  if ((fx | fy) == 0) { w0 = 32767; w3 = 1; }
  const uint8_t* s = src + (size_t)frame * sstride * sh;
  int v = 0;
  if ((unsigned)sx < (unsigned)max(sw - 1, 0) && (unsigned)sy < (unsigned)max(sh - 1, 0)) {
    const uint8_t* p = s + (size_t)sy * sstride + sx;
    v = __ldg(p) * w0 + __ldg(p + 1) * w1 + __ldg(p + sstride) * w2 + __ldg(p + sstride + 1) * w3;
    v = (v + (1 << 14)) >> 15;
  } else if (!(sx >= sw || sx + 1 < 0 || sy >= sh || sy + 1 < 0)) {
    const bool x0 = sx >= 0 && sx < sw, x1 = sx + 1 >= 0 && sx + 1 < sw;
    const bool y0 = sy >= 0 && sy < sh, y1 = sy + 1 >= 0 && sy + 1 < sh;
    const int p0 = (x0 && y0) ? __ldg(s + (size_t)sy * sstride + sx) : 0;
    const int p1 = (x1 && y0) ? __ldg(s + (size_t)sy * sstride + sx + 1) : 0;
    const int p2 = (x0 && y1) ? __ldg(s + (size_t)(sy + 1) * sstride + sx) : 0;
    const int p3 = (x1 && y1) ? __ldg(s + (size_t)(sy + 1) * sstride + sx + 1) : 0;
    v = (p0 * w0 + p1 * w1 + p2 * w2 + p3 * w3 + (1 << 14)) >> 15;
  }
  dst[(size_t)frame * dstride * rh + (size_t)y * dstride + x] = (uint8_t)min(max(v, 0), 255);
}

inline int sat_short(int v) { return v < -32768 ? -32768 : (v > 32767 ? 32767 : v); }

}  // namespace

extern "C" jn_rectify* jn_rectify_create(const float* mapx, const float* mapy, int32_t map_w, int32_t map_h,
                                         int32_t device) {
  if (!mapx || !mapy || map_w <= 0 || map_h <= 0) { jn_set_error("jn_rectify_create: bad arguments"); return nullptr; }
  if (cudaSetDevice(device) != cudaSuccess) { jn_set_error("jn_rectify_create: no CUDA device %d", device); return nullptr; }
  const size_t n = (size_t)map_w * map_h;
  std::vector<uint2> t(n);
  for (size_t i = 0; i < n; i++) {
    // cv::remap's float -> fixed conversion: cvRound (round half to even) of map * INTER_TAB_SIZE
    const int sx = (int)lrintf(mapx[i] * 32.0f), sy = (int)lrintf(mapy[i] * 32.0f);
    const unsigned ix = (unsigned)(sat_short(sx >> 5) & 0xFFFF), iy = (unsigned)(sat_short(sy >> 5) & 0xFFFF);
    t[i].x = ix | (iy << 16);
    t[i].y = (unsigned)((sy & 31) * 32 + (sx & 31));
  }
  jn_rectify* r = new jn_rectify;
  r->device = device; r->map_w = map_w; r->map_h = map_h; r->table = nullptr;
  if (cudaMalloc(&r->table, n * sizeof(uint2)) != cudaSuccess ||
      cudaMemcpy(r->table, t.data(), n * sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess) {
    jn_set_error("jn_rectify_create: device allocation or upload failed");
    cudaFree(r->table);
    delete r;
    return nullptr;
  }
  return r;
}

extern "C" void jn_rectify_destroy(jn_rectify* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  cudaFree(r->table);
  delete r;
}

extern "C" int jn_rectify_batch(jn_rectify* r, int32_t n, const uint8_t* src, int32_t src_w, int32_t src_h,
                                int32_t src_stride, const int32_t roi[4], uint8_t* dst, int32_t dst_stride, void* stream) {
  if (!r || n <= 0 || !src || !dst || src_w <= 0 || src_h <= 0 || src_stride < src_w) {
    jn_set_error("jn_rectify_batch: bad arguments");
    return JN_ERR_ARG;
  }
  const int rx = roi ? roi[0] : 0, ry = roi ? roi[1] : 0, rw = roi ? roi[2] : r->map_w, rh = roi ? roi[3] : r->map_h;
  if (rx < 0 || ry < 0 || rw <= 0 || rh <= 0 || rx + rw > r->map_w || ry + rh > r->map_h || dst_stride < rw) {
    jn_set_error("jn_rectify_batch: roi %d,%d %dx%d outside the %dx%d map or stride too small", rx, ry, rw, rh, r->map_w,
                 r->map_h);
    return JN_ERR_ARG;
  }
  JN_CUDA_CHECK(cudaSetDevice(r->device));
  remap_kernel<<<dim3((rw + 255) / 256, rh, n), 256, 0, (cudaStream_t)stream>>>(r->table, r->map_w, rx, ry, rw, rh, src,
                                                                                 src_w, src_h, src_stride, dst, dst_stride);
  g_jn_launches += 1;
  JN_CUDA_CHECK(cudaGetLastError());
  return JN_OK;
}
