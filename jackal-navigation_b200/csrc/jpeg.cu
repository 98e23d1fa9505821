// jpeg.cu -- compressed camera frames in: the decode step ahead of rectification.
//
// Replaces, for a batch of frames,
//     Mat tmp = cv::imdecode(Mat(msg->data), CV_LOAD_IMAGE_GRAYSCALE);          point_cloud.cpp:436, 478
// The reference receives sensor_msgs/CompressedImage (JPEG) and decodes on the CPU; here the
// bitstreams go to the GPU and nvJPEG (library code, like cuBLAS for a GEMM) decodes them straight
// into the device buffers the rectification / ELAS kernels read: output format Y = the luma plane,
// which is what a grayscale imdecode of a colour JPEG returns.  A host<->device copy of a frame then
// carries the compressed size instead of W*H bytes.  Decoders differ in IDCT rounding: the luma
// planes agree with OpenCV's (libjpeg-turbo) within +-2 grey levels, not bit for bit
// (tests/test_gpu_api.py::test_jpeg_decode_matches_imdecode).
#include <nvjpeg.h>
#include <vector>
#include "common.cuh"

struct jn_jpeg {
  int device;
  nvjpegHandle_t handle;
  nvjpegJpegState_t state;
  int batch;               // batch size the state is initialised for
};

extern "C" jn_jpeg* jn_jpeg_create(int device) {
  if (cudaSetDevice(device) != cudaSuccess) { jn_set_error("jn_jpeg_create: no CUDA device %d", device); return nullptr; }
  jn_jpeg* j = new jn_jpeg();
  j->device = device;
  j->batch = 0;
  if (nvjpegCreateSimple(&j->handle) != NVJPEG_STATUS_SUCCESS) {
    jn_set_error("jn_jpeg_create: nvjpegCreateSimple failed");
    delete j;
    return nullptr;
  }
  if (nvjpegJpegStateCreate(j->handle, &j->state) != NVJPEG_STATUS_SUCCESS) {
    jn_set_error("jn_jpeg_create: nvjpegJpegStateCreate failed");
    nvjpegDestroy(j->handle);
    delete j;
    return nullptr;
  }
  return j;
}

extern "C" void jn_jpeg_destroy(jn_jpeg* j) {
  if (!j) return;
  cudaSetDevice(j->device);
  nvjpegJpegStateDestroy(j->state);
  nvjpegDestroy(j->handle);
  delete j;
}

// Width / height of a JPEG without decoding it (host).
extern "C" int jn_jpeg_info(jn_jpeg* j, const uint8_t* data, size_t length, int32_t* width, int32_t* height) {
  if (!j || !data || !width || !height) return JN_ERR_ARG;
  int nc = 0, ws[NVJPEG_MAX_COMPONENT], hs[NVJPEG_MAX_COMPONENT];
  nvjpegChromaSubsampling_t ss;
  if (nvjpegGetImageInfo(j->handle, data, length, &nc, &ss, ws, hs) != NVJPEG_STATUS_SUCCESS) {
    jn_set_error("jn_jpeg_info: not a decodable JPEG");
    return JN_ERR_IO;
  }
  *width = ws[0];
  *height = hs[0];
  return JN_OK;
}

// n JPEG bitstreams (host pointers) -> n grayscale frames in DEVICE memory, frame i at
// dst + i * frame_stride, rows dst_stride bytes apart (>= width).  Asynchronous on `stream`; the
// bitstreams may be reused when the call returns (nvJPEG copies them).
extern "C" int jn_jpeg_decode_gray_batch(jn_jpeg* j, int n, const uint8_t* const* data, const size_t* lengths, uint8_t* dst,
                                         int32_t width, int32_t height, int32_t dst_stride, size_t frame_stride,
                                         void* stream) {
  if (!j || n <= 0 || !data || !lengths || !dst || dst_stride < width || frame_stride < (size_t)dst_stride * height) {
    jn_set_error("jn_jpeg_decode_gray_batch: bad arguments");
    return JN_ERR_ARG;
  }
  JN_CUDA_CHECK(cudaSetDevice(j->device));
  for (int i = 0; i < n; i++) {
    int32_t w = 0, h = 0;
    int rc = jn_jpeg_info(j, data[i], lengths[i], &w, &h);
    if (rc) return rc;
    if (w != width || h != height) {
      jn_set_error("jn_jpeg_decode_gray_batch: frame %d is %dx%d, expected %dx%d", i, w, h, width, height);
      return JN_ERR_ARG;
    }
  }
  if (j->batch != n) {
    if (nvjpegDecodeBatchedInitialize(j->handle, j->state, n, 1, NVJPEG_OUTPUT_Y) != NVJPEG_STATUS_SUCCESS) {
      jn_set_error("nvjpegDecodeBatchedInitialize failed");
      return JN_ERR_CUDA;
    }
    j->batch = n;
  }
  std::vector<nvjpegImage_t> out(n);
  for (int i = 0; i < n; i++) {
    for (int c = 0; c < NVJPEG_MAX_COMPONENT; c++) { out[i].channel[c] = nullptr; out[i].pitch[c] = 0; }
    out[i].channel[0] = dst + (size_t)i * frame_stride;
    out[i].pitch[0] = (size_t)dst_stride;
  }
  const nvjpegStatus_t st = nvjpegDecodeBatched(j->handle, j->state, data, lengths, out.data(), (cudaStream_t)stream);
  if (st != NVJPEG_STATUS_SUCCESS) {
    jn_set_error("nvjpegDecodeBatched failed (%d)", (int)st);
    j->batch = 0;
    return JN_ERR_CUDA;
  }
  return JN_OK;
}
