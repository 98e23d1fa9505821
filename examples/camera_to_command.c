/*
 * camera_to_command.c -- the reference's two nodes in one plain-C program, through the C ABI only.
 *
 *   point_cloud node:  calibration YAML -> stereoRectify -> Q            (point_cloud.cpp:530-544)
 *                      cacheDisparityValues                              (:104-147)
 *                      per frame: Elas::process + convertTo(CV_8U)       (:406-429)
 *                                 publishObstacleScan                    (:213-296)
 *   navigate node:     laserScanCallback, checkObstacle, chooseDirection (navigate.cpp:344-363, 101-197)
 *                      safeNavigate in obstacle-avoid mode -> cmd_vel    (:302-342, 229-255)
 *
 *   cc -std=c99 -I include examples/camera_to_command.c -L jackal-navigation_b200 -ljn_elas \
 *      -Wl,-rpath,$PWD/jackal-navigation_b200 -lm -o camera_to_command
 *   ./camera_to_command tests/golden/calib_c920.yml [frames]
 *
 * Frames are synthetic (a textured floor and a box that comes closer; rectified by construction, so the remap
 * step of jn_rectify_* is not needed here).  Only host-pointer, synchronous entry points are used: what a caller
 * without the CUDA runtime links.  Exit code 3: no usable GPU (the library has no CPU path).
 *
 * Expected output with the shipped calibration: what the CPU checkers give for the same frames (tools/example_expected.py:
 * the compiled reference ELAS, the scan restatement, the same host code for the command) -- a box left of the centre
 * approaches, the robot accelerates, brakes, and turns away from it on the spot:
 *   frame 0: 61 scan bins, nearest 1.61 m -> cmd_vel linear 0.025 m/s, angular +0.000 rad/s (driving)
 *   frame 1: 61 scan bins, nearest 0.84 m -> cmd_vel linear 0.000 m/s, angular -0.050 rad/s (turning right)
 *   frame 2: 61 scan bins, nearest 0.56 m -> cmd_vel linear 0.000 m/s, angular -0.100 rad/s (turning right)
 *   frame 3: 61 scan bins, nearest 0.42 m -> cmd_vel linear 0.000 m/s, angular -0.150 rad/s (turning right)
 *   frame 4: 61 scan bins, nearest 0.34 m -> cmd_vel linear 0.000 m/s, angular -0.200 rad/s (turning right)
 *   frame 5: 61 scan bins, nearest 0.28 m -> cmd_vel linear 0.000 m/s, angular -0.250 rad/s (turning right)
 *   frame 6: 61 scan bins, nearest 0.24 m -> cmd_vel linear 0.000 m/s, angular -0.300 rad/s (turning right)
 *   frame 7: 61 scan bins, nearest 0.21 m -> cmd_vel linear 0.000 m/s, angular -0.350 rad/s (turning right)
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jn_elas.h"

enum { W = 640, H = 360, DISP_MAX = 255 };

static unsigned lcg(unsigned* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

/* left = noise texture; right(u - d, v) = left(u, v) with d from a slanted floor and a box of disparity box_d */
static void make_pair(uint8_t* L, uint8_t* R, int box_d, unsigned seed) {
  unsigned s = seed;
  int u, v, pass;
  for (v = 0; v < H; v++)
    for (u = 0; u < W; u++) { L[v * W + u] = (uint8_t)(lcg(&s) & 255); R[v * W + u] = (uint8_t)(lcg(&s) & 255); }
  for (pass = 0; pass < 2; pass++)            /* far surface first, the box over it */
    for (v = 0; v < H; v++)
      for (u = 0; u < W; u++) {
        const int in_box = u > W / 8 && u < W / 2 && v > H / 4 && v < 3 * H / 4;   /* left of the centre */
        const int d = in_box ? box_d : 6 + 40 * v / H;
        if (in_box != pass) continue;
        if (u - d >= 0) R[v * W + u - d] = L[v * W + u];
      }
}

int main(int argc, char** argv) {
  const char* yml = argc > 1 ? argv[1] : "tests/golden/calib_c920.yml";
  const int frames = argc > 2 ? atoi(argv[2]) : 8;
  jn_calib cal;
  jn_elas_params par;
  jn_elas* elas;
  jn_scan* scan;
  jn_navigate* nav;
  uint8_t *L, *R;
  float* D1;
  int32_t dims[3];
  int f;

  if (jn_calib_load_yaml(yml, &cal) != JN_OK) { fprintf(stderr, "calibration: %s\n", jn_last_error()); return 2; }
  /* stereoRectify(K1, D1, K2, D2, Size(640, 360), R, T, ..., CALIB_ZERO_DISPARITY, 0, rawimsize) -> Q */
  if (jn_calib_stereo_rectify(&cal, 640, 360, W, H, 1, 0.0, NULL, NULL, NULL, NULL) != JN_OK) {
    fprintf(stderr, "stereoRectify: %s\n", jn_last_error());
    return 2;
  }
  printf("Q: cx %.2f cy %.2f f %.2f -1/Tx %.4f\n", -cal.Q[3], -cal.Q[7], cal.Q[11], cal.Q[14]);

  jn_elas_params_default(&par, JN_ROBOTICS);
  par.postprocess_only_left = 1;                                   /* point_cloud.cpp:417 */
  par.disp_max = DISP_MAX;
  elas = jn_elas_create(&par, 0);
  if (!elas) { fprintf(stderr, "%s\n", jn_last_error()); return 3; }
  scan = jn_scan_create(&cal, W, H, 0, 0, 0);                      /* cacheDisparityValues */
  if (!scan) { fprintf(stderr, "%s\n", jn_last_error()); jn_elas_destroy(elas); return 3; }
  nav = jn_navigate_create();

  L = (uint8_t*)malloc((size_t)W * H);
  R = (uint8_t*)malloc((size_t)W * H);
  D1 = (float*)malloc(sizeof(float) * W * H);
  dims[0] = W; dims[1] = H; dims[2] = W;
  for (f = 0; f < frames; f++) {
    double ranges[JN_SCAN_BINS], vel[2];
    jn_scan_meta meta;
    int rc;
    make_pair(L, R, 30 + 25 * f, 1000u + (unsigned)f);             /* the box approaches */
    memset(D1, 0, sizeof(float) * W * H);                          /* the caller zeroes its maps, :413-414 */
    rc = jn_elas_process(elas, L, R, D1, NULL, dims);
    if (rc < 0) { fprintf(stderr, "Elas::process: %s\n", jn_last_error()); return 1; }
    if (jn_scan_from_disparity(scan, D1, ranges, &meta, NULL) != JN_OK) { fprintf(stderr, "%s\n", jn_last_error()); return 1; }
    jn_navigate_set_scan_bins(nav, ranges, &meta);
    /* the X button of the joystick held, stick fully forward: obstacleAvoidMode (navigate.cpp:229-255) + the ramp */
    jn_navigate_command(nav, JN_NAV_OBSTACLE_AVOID, 0.0, 1.0, vel);
    printf("frame %d: %d scan bins, nearest %.2f m -> cmd_vel linear %.3f m/s, angular %+.3f rad/s (%s)\n", f,
           (int)meta.n_finite, meta.range_min, vel[0], vel[1],
           vel[1] > 0 ? "turning left" : vel[1] < 0 ? "turning right" : vel[0] > 0 ? "driving" : "stopped");
  }
  free(L); free(R); free(D1);
  jn_navigate_destroy(nav);
  jn_scan_destroy(scan);
  jn_elas_destroy(elas);
  return 0;
}
