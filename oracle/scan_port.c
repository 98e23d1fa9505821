/*
 * scan_port.c -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the per-pixel part of the reference's
 * src/obstacle_avoidance/point_cloud.cpp.  PINNED three ways:
 *  (1) against the reference's own code: point_cloud.cpp compiled where it lies with stand-in ROS / OpenCV
 *      headers (oracle/standins, oracle/pointcloud_ref_shim.cpp -> oracle/_ref/libpointcloud_ref.so);
 *      tests/test_reference_nodes_pin.py calls the node's cacheDisparityValues, generateDisparityMap,
 *      publishPointCloud and both publishObstacleScan overloads and compares the gate cache, the CV_8U map,
 *      the Point32 / rgb payload and the published LaserScan with the functions below, bit for bit.  That
 *      pins the statements (loop order, constants, casts, conditions); the stand-in cv::Mat supplies the
 *      arithmetic of OpenCV's calls, so
 *  (2) the OpenCV arithmetic is pinned against OpenCV 4.13 itself: cv::Mat products are evaluated as OpenCV's
 *      small-matrix gemm does, ((a0*b0 + a1*b1) + a2*b2) [+ a3*b3], then + C; tests/golden/scan_cv2.npz holds
 *      cv2.gemm results for Q*V and XR*p+XT (sparse and dense Q) and the saturate_cast<uchar> of
 *      convertTo(CV_8U); tests/test_oracle_pin.py checks reproject() and port_convert_u8 against them bit
 *      for bit (the reference's OpenCV is 2.4-era and not pinned in package.xml: not available here);
 *  (3) against tests/golden/scan_statements.npz, a statement-by-statement execution of the same functions
 *      with cv2 (generator committed).  atan2/sqrt/floor are libm.
 *
 * One behaviour is DEFINED here because the reference leaves it undefined (H8):
 * a bin index outside [0,89] (|theta| > 45 deg) is skipped; the reference writes
 * outside scan[90].
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "oracle_abi.h"

#define BINS 90
static const double SCAN_INF = 1e9;            /* const int INF = 1e9  (point_cloud.cpp:55) */
static const double GP_HEIGHT = 0.05;          /* :66 */
static const double GP_ANGLE = 4. * 3.1415 / 180.; /* :67 */
static const double GP_DIST = 1.0;             /* :68 */

/* pos = Q*[x,y,d,1]; p = pos.xyz/pos.w; r = XR*p + XT   (point_cloud.cpp:237-253) */
static void reproject(const double* Q, const double* XR, const double* XT, double x, double y, double d,
                      double out[3]) {
  double V[4] = {x, y, d, 1.0}, pos[4];
  for (int i = 0; i < 4; i++)
    pos[i] = ((Q[4 * i] * V[0] + Q[4 * i + 1] * V[1]) + Q[4 * i + 2] * V[2]) + Q[4 * i + 3] * V[3];
  double X = pos[0] / pos[3], Y = pos[1] / pos[3], Z = pos[2] / pos[3];
  for (int i = 0; i < 3; i++)
    out[i] = ((XR[3 * i] * X + XR[3 * i + 1] * Y) + XR[3 * i + 2] * Z) + XT[i];
}

static int above_ground(double X, double Z) {
  if (X < GP_DIST) return !(Z < GP_HEIGHT);
  return !(Z < GP_HEIGHT + tan(GP_ANGLE) * (X - GP_DIST));
}

/* cacheDisparityValues (point_cloud.cpp:104-147): gate[2*(j*W+i)] = smallest d in [3,255]
 * above the ground gate (256 wraps to 0), gate[..+1] = 255. */
void port_gate_cache(const double* Q, const double* XR, const double* XT, int W, int H, int ox, int oy,
                     uint8_t* gate) {
  for (int i = 0; i < W; i++)
    for (int j = 0; j < H; j++) {
      int d;
      for (d = 3; d <= 255; d++) {
        double r[3];
        reproject(Q, XR, XT, (double)(i + ox), (double)(j + oy), (double)d, r);
        if (r[2] < 0.) continue;
        if (!above_ground(r[0], r[2])) continue;
        break;
      }
      gate[2 * ((size_t)j * W + i)] = (uint8_t)d;
      gate[2 * ((size_t)j * W + i) + 1] = 255;
    }
}

/* leftdpf.convertTo(show, CV_8U, 1.)  (point_cloud.cpp:421-422):
 * saturate_cast<uchar>(cvRound(x)), round-half-even. */
void port_convert_u8(const float* D, size_t n, uint8_t* out) {
  for (size_t i = 0; i < n; i++) {
    long r = lrintf(D[i]);
    out[i] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
  }
}

static void scan_init(double* scan, oracle_scan_meta* m) {
  for (int i = 0; i < BINS; i++) scan[i] = SCAN_INF;
  m->angle_min = 400; m->angle_max = -400;
  m->range_min = SCAN_INF; m->range_max = -500;
  m->n_finite = 0; m->n_points = 0;
}

static void scan_add(double X, double Y, double* scan, oracle_scan_meta* m) {
  double th = atan2(Y, X);
  double deg = th * 180. / 3.1415;
  /* DEFINED here: a pixel whose reprojection is NaN (d = 0 passing a wrapped gate divides by
   * W = 0, H8) is skipped; the reference would index scan[] with an undefined integer. */
  if (th != th || X != X || Y != Y) return;
  if (th < m->angle_min) m->angle_min = th;
  if (th > m->angle_max) m->angle_max = th;
  double r = sqrt(Y * Y + X * X);
  if (r > m->range_max) m->range_max = r;
  if (r < m->range_min) m->range_min = r;
  m->n_points++;
  double kf = floor((double)BINS * (90. / 2. - deg) / 90.);
  if (!(kf >= 0 && kf < BINS)) return; /* H8: defined as skip */
  int k = (int)kf;
  if (r < scan[k]) scan[k] = r;
}

static void scan_finish(const double* scan, oracle_scan_meta* m) {
  for (int i = 0; i < BINS; i++)
    if (scan[i] < SCAN_INF - 1) m->n_finite++;
}

/* publishObstacleScan(Mat& dmap, seq)  (point_cloud.cpp:213-296) */
void port_scan_from_dmap(const double* Q, const double* XR, const double* XT, const uint8_t* gate,
                         const uint8_t* dmap, int W, int H, int ox, int oy, double* scan,
                         oracle_scan_meta* meta) {
  scan_init(scan, meta);
  for (int i = 0; i < W; i++)
    for (int j = 0; j < H; j++) {
      int d = dmap[(size_t)j * W + i];
      const uint8_t* g = gate + 2 * ((size_t)j * W + i);
      if (d < g[0] || d > g[1]) continue;
      double r[3];
      reproject(Q, XR, XT, (double)(i + ox), (double)(j + oy), (double)d, r);
      scan_add(r[0], r[1], scan, meta);
    }
  scan_finish(scan, meta);
}

/* publishPointCloud, -g path (point_cloud.cpp:314-387): every pixel with d >= 2,
 * columns outer; returns the point count. */
int port_points_from_dmap(const double* Q, const double* XR, const double* XT, const uint8_t* dmap, int W,
                          int H, int ox, int oy, double* pts) {
  int n = 0;
  for (int i = 0; i < W; i++)
    for (int j = 0; j < H; j++) {
      int d = dmap[(size_t)j * W + i];
      if (d < 2) continue;
      reproject(Q, XR, XT, (double)(i + ox), (double)(j + oy), (double)d, pts + 3 * (size_t)n);
      n++;
    }
  return n;
}

/* sensor_msgs/PointCloud payload of publishPointCloud (point_cloud.cpp:351-383): per point a
 * geometry_msgs/Point32 (the robot-frame doubles narrowed to float32) and one "rgb" channel value =
 * the bits of int32 (red << 16 | green << 8 | blue) with red/green/blue = leftim_res.at<Vec3b>(j,i)[2,1,0].
 * channels = 3: `img` is that BGR image.  channels = 1: the reference decodes GRAYSCALE frames
 * (point_cloud.cpp:436) and still indexes them as Vec3b, i.e. it reads bytes 3i, 3i+1, 3i+2 of row j
 * (running into the following rows); restated literally, bytes beyond the H x stride buffer read as 0. */
int port_pointcloud_pack(const double* Q, const double* XR, const double* XT, const uint8_t* dmap, int W, int H,
                         int ox, int oy, const uint8_t* img, int stride, int channels, float* xyz, float* rgb) {
  int n = 0;
  const size_t img_bytes = (size_t)stride * H;
  for (int i = 0; i < W; i++)
    for (int j = 0; j < H; j++) {
      int d = dmap[(size_t)j * W + i];
      if (d < 2) continue;
      double p[3];
      reproject(Q, XR, XT, (double)(i + ox), (double)(j + oy), (double)d, p);
      xyz[3 * (size_t)n] = (float)p[0];
      xyz[3 * (size_t)n + 1] = (float)p[1];
      xyz[3 * (size_t)n + 2] = (float)p[2];
      int32_t c[3];
      for (int k = 0; k < 3; k++) {
        size_t a = (size_t)j * stride + 3 * (size_t)i + k;
        c[k] = (channels == 3 || a < img_bytes) ? img[a] : 0;
      }
      int32_t v = (c[2] << 16) | (c[1] << 8) | c[0];
      memcpy(rgb + n, &v, 4);
      n++;
    }
  return n;
}

/* publishObstacleScan(vector<Point3d>, seq)  (point_cloud.cpp:149-211) */
void port_scan_from_points(const double* pts, int n, double* scan, oracle_scan_meta* meta) {
  scan_init(scan, meta);
  for (int i = 0; i < n; i++) {
    const double* p = pts + 3 * (size_t)i;
    if (!above_ground(p[0], p[2])) continue;
    scan_add(p[0], p[1], scan, meta);
  }
  scan_finish(scan, meta);
}

/* LaserScan.ranges as published: finite bins, k = 89..0  (point_cloud.cpp:278-282) */
int port_scan_compact(const double* scan, float* out) {
  int n = 0;
  for (int i = BINS - 1; i >= 0; i--)
    if (scan[i] < SCAN_INF - 1) out[n++] = (float)scan[i];
  return n;
}
