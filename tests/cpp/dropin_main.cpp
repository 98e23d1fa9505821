// Reference-style caller of the drop-in header: the body of run() is the code of
// generateDisparityMap (reference src/obstacle_avoidance/point_cloud.cpp:406-419) with cv::Mat
// replaced by plain buffers.  Built and run by tests/test_cpp_dropin.py.
//   dropin_main <width> <height> <disp_max> <left.u8> <right.u8> <out_left.f32> <out_right.f32>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "elas.h"

static bool slurp(const char* path, std::vector<uint8_t>& buf) {
  FILE* f = fopen(path, "rb");
  if (!f) return false;
  size_t n = fread(buf.data(), 1, buf.size(), f);
  fclose(f);
  return n == buf.size();
}

static int run(int w, int h, int disp_max, char** files) {
  std::vector<uint8_t> left((size_t)w * h), right((size_t)w * h);
  if (!slurp(files[0], left) || !slurp(files[1], right)) { fprintf(stderr, "cannot read inputs\n"); return 2; }
  const int32_t dims[3] = {w, h, w};  // bytes per line = width
  std::vector<float> leftdpf((size_t)w * h, 0.f), rightdpf((size_t)w * h, 0.f);

  Elas::parameters param;
  param.postprocess_only_left = true;
  param.disp_max = disp_max;
  Elas elas(param);
  elas.process(left.data(), right.data(), leftdpf.data(), rightdpf.data(), dims);

  FILE* f = fopen(files[2], "wb");
  fwrite(leftdpf.data(), sizeof(float), leftdpf.size(), f);
  fclose(f);
  f = fopen(files[3], "wb");
  fwrite(rightdpf.data(), sizeof(float), rightdpf.size(), f);
  fclose(f);
  return 0;
}

int main(int argc, char** argv) {
  if (argc != 8) {
    Elas::parameters r(Elas::ROBOTICS), m(Elas::MIDDLEBURY);
    printf("presets: ROBOTICS support_threshold=%.2f gamma=%.0f ipol_gap_width=%d | MIDDLEBURY support_threshold=%.2f "
           "gamma=%.0f ipol_gap_width=%d\n", r.support_threshold, r.gamma, r.ipol_gap_width, m.support_threshold,
           m.gamma, m.ipol_gap_width);
    return 0;
  }
  try {
    return run(atoi(argv[1]), atoi(argv[2]), atoi(argv[3]), argv + 4);
  } catch (const std::exception& e) {
    fprintf(stderr, "%s\n", e.what());
    return 3;
  }
}
