/*
 * elas.h -- drop-in replacement for the reference's src/elas/elas.h (class Elas).
 *
 * Same public interface as the reference header (elas.h:52-162):
 *
 *     Elas::parameters param;                 // same 23 fields, same presets
 *     param.postprocess_only_left = true;
 *     Elas elas(param);
 *     elas.process(I1, I2, D1, D2, dims);     // dims = {width, height, bytes_per_line}
 *
 * so a caller such as generateDisparityMap (src/obstacle_avoidance/point_cloud.cpp:406-429)
 * compiles unchanged.  The work is done by the sm_100a CUDA kernels in libjn_elas.so through
 * the C ABI of jn_elas.h; this header holds no algorithm.  Link with -ljn_elas.
 *
 * Behavioural notes (identical to the reference unless stated):
 *   - process() is synchronous and leaves D1/D2 untouched and prints
 *     "ERROR: Need at least 3 support points!" when fewer than 3 support points are found
 *     (elas.cpp:66-71).
 *   - The reference's process() cannot fail otherwise; here a CUDA or argument error throws
 *     std::runtime_error (there is deliberately no CPU fallback).
 *   - Unlike the reference, an Elas object keeps its device workspace between calls (the
 *     reference allocates and frees everything per call, elas.cpp:40-41,147-150); results do
 *     not depend on it.  `device` selects the GPU (default 0).
 */
#ifndef __ELAS_H__
#define __ELAS_H__

#include <stdint.h>
#include <iostream>
#include <stdexcept>
#include <string>

#include "jn_elas.h"

class Elas {
 public:
  enum setting { ROBOTICS, MIDDLEBURY };

  // parameter settings: field for field Elas::parameters (elas.h:59-145)
  struct parameters {
    int32_t disp_min;
    int32_t disp_max;
    float   support_threshold;
    int32_t support_texture;
    int32_t candidate_stepsize;
    int32_t incon_window_size;
    int32_t incon_threshold;
    int32_t incon_min_support;
    bool    add_corners;
    int32_t grid_size;
    float   beta;
    float   gamma;
    float   sigma;
    float   sradius;
    int32_t match_texture;
    int32_t lr_threshold;
    float   speckle_sim_threshold;
    int32_t speckle_size;
    int32_t ipol_gap_width;
    bool    filter_median;
    bool    filter_adaptive_mean;
    bool    postprocess_only_left;
    bool    subsampling;

    parameters(setting s = ROBOTICS) {
      jn_elas_params p;
      jn_elas_params_default(&p, s == ROBOTICS ? JN_ROBOTICS : JN_MIDDLEBURY);
      disp_min = p.disp_min; disp_max = p.disp_max;
      support_threshold = p.support_threshold; support_texture = p.support_texture;
      candidate_stepsize = p.candidate_stepsize; incon_window_size = p.incon_window_size;
      incon_threshold = p.incon_threshold; incon_min_support = p.incon_min_support;
      add_corners = p.add_corners != 0; grid_size = p.grid_size;
      beta = p.beta; gamma = p.gamma; sigma = p.sigma; sradius = p.sradius;
      match_texture = p.match_texture; lr_threshold = p.lr_threshold;
      speckle_sim_threshold = p.speckle_sim_threshold; speckle_size = p.speckle_size;
      ipol_gap_width = p.ipol_gap_width; filter_median = p.filter_median != 0;
      filter_adaptive_mean = p.filter_adaptive_mean != 0;
      postprocess_only_left = p.postprocess_only_left != 0; subsampling = p.subsampling != 0;
    }
  };

  // constructor, input: parameters
  Elas(parameters param, int device = 0) : param(param), device_(device), handle_(0) {}

  ~Elas() {
    if (handle_) jn_elas_destroy(handle_);
  }

  // matching function
  // inputs: pointers to left (I1) and right (I2) intensity image (uint8, input)
  //         pointers to left (D1) and right (D2) disparity image (float, output)
  //         dims[0] = width of I1 and I2
  //         dims[1] = height of I1 and I2
  //         dims[2] = bytes per line (often equal to width, but allowed to differ)
  //         note: D1 and D2 must be allocated before (bytes per line = width)
  void process(uint8_t* I1, uint8_t* I2, float* D1, float* D2, const int32_t* dims) {
    if (!handle_) {
      jn_elas_params p = to_c();
      handle_ = jn_elas_create(&p, device_);
      if (!handle_) throw std::runtime_error(std::string("Elas: ") + jn_last_error());
    }
    int rc = jn_elas_process(handle_, I1, I2, D1, D2, dims);
    if (rc == JN_FEW_SUPPORT) {
      std::cout << "ERROR: Need at least 3 support points!" << std::endl;
      return;
    }
    if (rc != JN_OK) throw std::runtime_error(std::string("Elas::process: ") + jn_last_error());
  }

 private:
  Elas(const Elas&);             // the device workspace is not copyable
  Elas& operator=(const Elas&);

  jn_elas_params to_c() const {
    jn_elas_params p;
    p.disp_min = param.disp_min; p.disp_max = param.disp_max;
    p.support_threshold = param.support_threshold; p.support_texture = param.support_texture;
    p.candidate_stepsize = param.candidate_stepsize; p.incon_window_size = param.incon_window_size;
    p.incon_threshold = param.incon_threshold; p.incon_min_support = param.incon_min_support;
    p.add_corners = param.add_corners; p.grid_size = param.grid_size;
    p.beta = param.beta; p.gamma = param.gamma; p.sigma = param.sigma; p.sradius = param.sradius;
    p.match_texture = param.match_texture; p.lr_threshold = param.lr_threshold;
    p.speckle_sim_threshold = param.speckle_sim_threshold; p.speckle_size = param.speckle_size;
    p.ipol_gap_width = param.ipol_gap_width; p.filter_median = param.filter_median;
    p.filter_adaptive_mean = param.filter_adaptive_mean;
    p.postprocess_only_left = param.postprocess_only_left; p.subsampling = param.subsampling;
    return p;
  }

  // parameter set
  parameters param;
  int device_;
  jn_elas* handle_;
};

#endif
