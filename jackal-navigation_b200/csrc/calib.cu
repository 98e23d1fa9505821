// calib.cu -- init-time host code of the calibration boundary: stereo rectification and the
// undistort/rectify maps, so that a caller WITHOUT OpenCV gets from the calibration YAML
// (K1, K2, D1, D2, R, T) to Q and to the remap tables.
//
// Replaces the two OpenCV calls of the reference's main()
//     stereoRectify(K1, D1, K2, D2, calib_im_size, R, Mat(T), R1, R2, P1, P2, Q,
//                   CV_CALIB_ZERO_DISPARITY, 0, rawimsize);                        point_cloud.cpp:543-544
//     initUndistortRectifyMap(K1, D1, R1, P1, rawimsize, CV_32F, lmapx, lmapy);      point_cloud.cpp:553-554
// OpenCV is a third-party dependency that is not under /root/reference (version unpinned, 2.4-era
// API); this is a restatement of the published algorithm as OpenCV 4.13 implements it (Bouguet's
// rectification: half rotation of both cameras, alignment of the baseline with the x axis, common
// focal length = mean of the two, principal points from the undistorted image corners, the
// alpha scaling from a 9x9 grid of undistorted points; plumb-bob distortion k1,k2,p1,p2,k3,
// 5 fixed-point iterations for the inverse).  Checked against cv2 4.13 on the shipped calibration
// and on random ones: Q, R1, R2, P1, P2 agree to 1e-9 relative, the maps to 1e-3 px
// (tests/test_host_logic.py).  Horizontal and vertical stereo, CALIB_ZERO_DISPARITY on or off.
#include <float.h>
#include <math.h>
#include <string.h>
#include "common.cuh"

namespace {

typedef double M3[3][3];

void matmul3(const M3 a, const M3 b, M3 c) {
  M3 t;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
  memcpy(c, t, sizeof(M3));
}
void transpose3(const M3 a, M3 c) {
  M3 t;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) t[i][j] = a[j][i];
  memcpy(c, t, sizeof(M3));
}
bool inverse3(const M3 a, M3 c) {
  const double d = a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
                   a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
  if (d == 0) return false;
  const double id = 1.0 / d;
  M3 t;
  t[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * id;
  t[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
  t[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
  t[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * id;
  t[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
  t[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
  t[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * id;
  t[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
  t[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
  memcpy(c, t, sizeof(M3));
  return true;
}

// Rodrigues, rotation vector -> matrix.
void rodrigues_v2m(const double r[3], M3 R) {
  const double th = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (th < DBL_EPSILON) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) R[i][j] = i == j;
    return;
  }
  const double c = cos(th), s = sin(th), c1 = 1. - c, it = 1. / th;
  const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
  const double rrt[3][3] = {{x * x, x * y, x * z}, {x * y, y * y, y * z}, {x * z, y * z, z * z}};
  const double rx[3][3] = {{0, -z, y}, {z, 0, -x}, {-y, x, 0}};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = c * (i == j) + c1 * rrt[i][j] + s * rx[i][j];
}

// Rodrigues, matrix -> rotation vector.  OpenCV first replaces R by the nearest rotation (U V^T of its
// SVD); the same matrix is the orthogonal polar factor, reached here by Newton's iteration.
void rodrigues_m2v(const M3 Rin, double r[3]) {
  M3 R;
  memcpy(R, Rin, sizeof(M3));
  for (int it = 0; it < 20; it++) {
    M3 inv, invT, next;
    if (!inverse3(R, inv)) break;
    transpose3(inv, invT);
    double diff = 0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        next[i][j] = 0.5 * (R[i][j] + invT[i][j]);
        diff = fmax(diff, fabs(next[i][j] - R[i][j]));
      }
    memcpy(R, next, sizeof(M3));
    if (diff < 1e-16) break;
  }
  double v[3] = {R[2][1] - R[1][2], R[0][2] - R[2][0], R[1][0] - R[0][1]};
  const double s = sqrt((v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) * 0.25);
  double c = (R[0][0] + R[1][1] + R[2][2] - 1) * 0.5;
  c = c > 1. ? 1. : (c < -1. ? -1. : c);
  const double th = acos(c);
  if (s < 1e-5) {
    if (c > 0) { r[0] = r[1] = r[2] = 0; return; }
    double t = (R[0][0] + 1) * 0.5;
    double x = sqrt(fmax(t, 0.));
    t = (R[1][1] + 1) * 0.5;
    double y = sqrt(fmax(t, 0.)) * (R[0][1] < 0 ? -1. : 1.);
    t = (R[2][2] + 1) * 0.5;
    double z = sqrt(fmax(t, 0.)) * (R[0][2] < 0 ? -1. : 1.);
    if (fabs(x) < fabs(y) && fabs(x) < fabs(z) && (R[1][2] > 0) != (y * z > 0)) z = -z;
    const double n = th / sqrt(x * x + y * y + z * z);
    r[0] = x * n; r[1] = y * n; r[2] = z * n;
    return;
  }
  const double vth = th / (2 * s);
  r[0] = v[0] * vth; r[1] = v[1] * vth; r[2] = v[2] * vth;
}

// undistortPoints: pixel -> normalised, distortion removed by 5 fixed-point iterations, then R and P.
void undistort_point(double u, double v, const double K[9], const double D[5], const M3 RR, double& ox, double& oy) {
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
  double x = (u - cx) / fx, y = (v - cy) / fy;
  const double x0 = x, y0 = y;
  for (int j = 0; j < 5; j++) {
    const double r2 = x * x + y * y;
    const double icdist = 1. / (1 + ((k3 * r2 + k2) * r2 + k1) * r2);
    const double dX = 2 * p1 * x * y + p2 * (r2 + 2 * x * x);
    const double dY = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y;
    x = (x0 - dX) * icdist;
    y = (y0 - dY) * icdist;
  }
  const double xx = RR[0][0] * x + RR[0][1] * y + RR[0][2], yy = RR[1][0] * x + RR[1][1] * y + RR[1][2];
  const double ww = 1. / (RR[2][0] * x + RR[2][1] * y + RR[2][2]);
  ox = xx * ww;
  oy = yy * ww;
}

struct RectD { double x, y, w, h; };

// Inner / outer rectangle of the undistorted, rectified image area from a 9x9 grid of source points.
void get_rectangles(const double K[9], const double D[5], const M3 R, const double P[12], int w, int h, RectD& inner,
                    RectD& outer) {
  const int N = 9;
  M3 A = {{P[0], P[1], P[2]}, {P[4], P[5], P[6]}, {P[8], P[9], P[10]}}, RR;
  matmul3(A, R, RR);
  double iX0 = -FLT_MAX, iX1 = FLT_MAX, iY0 = -FLT_MAX, iY1 = FLT_MAX;
  double oX0 = FLT_MAX, oX1 = -FLT_MAX, oY0 = FLT_MAX, oY1 = -FLT_MAX;
  for (int y = 0; y < N; y++)
    for (int x = 0; x < N; x++) {
      double px, py;
      undistort_point((double)x * (w - 1) / (N - 1), (double)y * (h - 1) / (N - 1), K, D, RR, px, py);
      oX0 = fmin(oX0, px); oX1 = fmax(oX1, px); oY0 = fmin(oY0, py); oY1 = fmax(oY1, py);
      if (x == 0) iX0 = fmax(iX0, px);
      if (x == N - 1) iX1 = fmin(iX1, px);
      if (y == 0) iY0 = fmax(iY0, py);
      if (y == N - 1) iY1 = fmin(iY1, py);
    }
  inner = {iX0, iY0, iX1 - iX0, iY1 - iY0};
  outer = {oX0, oY0, oX1 - oX0, oY1 - oY0};
}

double max4(double a, double b, double c, double d) { return fmax(fmax(fmax(a, b), c), d); }
double min4(double a, double b, double c, double d) { return fmin(fmin(fmin(a, b), c), d); }

}  // namespace

// cv::stereoRectify(K1, D1, K2, D2, Size(calib_w, calib_h), R, T, R1, R2, P1, P2, Q, flags, alpha,
// Size(new_w, new_h)).  zero_disparity = CALIB_ZERO_DISPARITY (what the reference passes); alpha < 0 =
// no scaling, the reference passes 0.  Outputs row-major: R1, R2 3x3, P1, P2 3x4, Q 4x4 (also stored
// in c->Q).  new_w/new_h = 0: same size as the calibration images.
extern "C" int jn_calib_stereo_rectify(jn_calib* c, int calib_w, int calib_h, int new_w, int new_h, int zero_disparity,
                                       double alpha, double R1o[9], double R2o[9], double P1o[12], double P2o[12]) {
  if (!c || calib_w <= 0 || calib_h <= 0) { jn_set_error("jn_calib_stereo_rectify: bad arguments"); return JN_ERR_ARG; }
  M3 R, r_r, wR, R1, R2, tmp;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R[i][j] = c->R[3 * i + j];
  double om[3];
  rodrigues_m2v(R, om);
  for (int i = 0; i < 3; i++) om[i] *= -0.5;          // half of the relative rotation, for both cameras
  rodrigues_v2m(om, r_r);
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = r_r[i][0] * c->T[0] + r_r[i][1] * c->T[1] + r_r[i][2] * c->T[2];
  const int idx = fabs(t[0]) > fabs(t[1]) ? 0 : 1;   // horizontal or vertical stereo
  const double cc = t[idx], nt = sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
  if (!(nt > 0.0)) { jn_set_error("jn_calib_stereo_rectify: zero baseline"); return JN_ERR_ARG; }
  double uu[3] = {0, 0, 0};
  uu[idx] = cc > 0 ? 1 : -1;
  // rotation that takes the baseline onto the image x (y) axis
  double ww[3] = {t[1] * uu[2] - t[2] * uu[1], t[2] * uu[0] - t[0] * uu[2], t[0] * uu[1] - t[1] * uu[0]};
  const double nw = sqrt(ww[0] * ww[0] + ww[1] * ww[1] + ww[2] * ww[2]);
  if (nw > 0.0) {
    const double f = acos(fabs(cc) / nt) / nw;
    for (int i = 0; i < 3; i++) ww[i] *= f;
  }
  rodrigues_v2m(ww, wR);
  transpose3(r_r, tmp);
  matmul3(wR, tmp, R1);
  matmul3(wR, r_r, R2);
  for (int i = 0; i < 3; i++) t[i] = R2[i][0] * c->T[0] + R2[i][1] * c->T[1] + R2[i][2] * c->T[2];

  if (new_w <= 0 || new_h <= 0) { new_w = calib_w; new_h = calib_h; }
  const int nx = calib_w, ny = calib_h;
  const double ratio_x = (double)new_w / calib_w / 2, ratio_y = (double)new_h / calib_h / 2;
  const double ratio = idx == 1 ? ratio_x : ratio_y;
  const int di = (idx ^ 1) * 4;                        // K[idx^1][idx^1]
  double fc_new = (c->K1[di] + c->K2[di]) * ratio;
  double ccn[2][2];
  const M3 I3 = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int k = 0; k < 2; k++) {
    const double* A = k == 0 ? c->K1 : c->K2;
    const double* Dk = k == 0 ? c->D1 : c->D2;
    // OpenCV projects the undistorted corners with projectPoints(R as a rotation VECTOR): the matrix
    // goes through Rodrigues and back
    double rv[3];
    M3 Rk;
    rodrigues_m2v(k == 0 ? R1 : R2, rv);
    rodrigues_v2m(rv, Rk);
    double ax = 0, ay = 0;
    for (int i = 0; i < 4; i++) {
      const int j = i < 2 ? 0 : 1;
      double px, py;
      undistort_point((float)((i % 2) * (nx - 1)), (float)(j * (ny - 1)), A, Dk, I3, px, py);
      px = (float)px; py = (float)py;                  // CV_32FC2 points in between
      const double X = Rk[0][0] * px + Rk[0][1] * py + Rk[0][2], Y = Rk[1][0] * px + Rk[1][1] * py + Rk[1][2];
      const double Z = Rk[2][0] * px + Rk[2][1] * py + Rk[2][2];
      const double z = Z ? 1. / Z : 1.;
      ax += (float)(X * z * fc_new);
      ay += (float)(Y * z * fc_new);
    }
    ccn[k][0] = (nx - 1) / 2. - ax / 4;
    ccn[k][1] = (ny - 1) / 2. - ay / 4;
  }
  if (zero_disparity) {
    ccn[0][0] = ccn[1][0] = (ccn[0][0] + ccn[1][0]) * 0.5;
    ccn[0][1] = ccn[1][1] = (ccn[0][1] + ccn[1][1]) * 0.5;
  } else if (idx == 0) {
    ccn[0][1] = ccn[1][1] = (ccn[0][1] + ccn[1][1]) * 0.5;
  } else {
    ccn[0][0] = ccn[1][0] = (ccn[0][0] + ccn[1][0]) * 0.5;
  }
  double P1[12] = {fc_new, 0, ccn[0][0], 0, 0, fc_new, ccn[0][1], 0, 0, 0, 1, 0};
  double P2[12] = {fc_new, 0, ccn[1][0], 0, 0, fc_new, ccn[1][1], 0, 0, 0, 1, 0};
  P2[4 * idx + 3] = t[idx] * fc_new;                   // baseline * focal length

  if (alpha > 1.) alpha = 1.;
  RectD in1, out1, in2, out2;
  get_rectangles(c->K1, c->D1, R1, P1, nx, ny, in1, out1);
  get_rectangles(c->K2, c->D2, R2, P2, nx, ny, in2, out2);
  const double cx1_0 = ccn[0][0], cy1_0 = ccn[0][1], cx2_0 = ccn[1][0], cy2_0 = ccn[1][1];
  const double cx1 = new_w * cx1_0 / nx, cy1 = new_h * cy1_0 / ny, cx2 = new_w * cx2_0 / nx, cy2 = new_h * cy2_0 / ny;
  double s = 1.;
  if (alpha >= 0) {
    double s0 = max4(cx1 / (cx1_0 - in1.x), cy1 / (cy1_0 - in1.y), (new_w - 1 - cx1) / (in1.x + in1.w - cx1_0),
                     (new_h - 1 - cy1) / (in1.y + in1.h - cy1_0));
    s0 = fmax(s0, max4(cx2 / (cx2_0 - in2.x), cy2 / (cy2_0 - in2.y), (new_w - 1 - cx2) / (in2.x + in2.w - cx2_0),
                       (new_h - 1 - cy2) / (in2.y + in2.h - cy2_0)));
    double s1 = min4(cx1 / (cx1_0 - out1.x), cy1 / (cy1_0 - out1.y), (new_w - 1 - cx1) / (out1.x + out1.w - cx1_0),
                     (new_h - 1 - cy1) / (out1.y + out1.h - cy1_0));
    s1 = fmin(s1, min4(cx2 / (cx2_0 - out2.x), cy2 / (cy2_0 - out2.y), (new_w - 1 - cx2) / (out2.x + out2.w - cx2_0),
                       (new_h - 1 - cy2) / (out2.y + out2.h - cy2_0)));
    s = s0 * (1 - alpha) + s1 * alpha;
  }
  fc_new *= s;
  P1[0] = P1[5] = fc_new; P1[2] = cx1; P1[6] = cy1;
  P2[0] = P2[5] = fc_new; P2[2] = cx2; P2[6] = cy2;
  P2[4 * idx + 3] = s * P2[4 * idx + 3];
  const double q[16] = {1, 0, 0, -cx1, 0, 1, 0, -cy1, 0, 0, 0, fc_new, 0, 0, -1. / t[idx],
                        (idx == 0 ? cx1 - cx2 : cy1 - cy2) / t[idx]};
  memcpy(c->Q, q, sizeof(q));
  c->has_q = 1;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      if (R1o) R1o[3 * i + j] = R1[i][j];
      if (R2o) R2o[3 * i + j] = R2[i][j];
    }
  if (P1o) memcpy(P1o, P1, sizeof(P1));
  if (P2o) memcpy(P2o, P2, sizeof(P2));
  return JN_OK;
}

// cv::initUndistortRectifyMap(K, D, R, P, Size(w, h), CV_32F, mapx, mapy): for every pixel of the
// rectified image the source position in the distorted camera image (feed the pair to jn_rectify_create).
extern "C" int jn_calib_init_undistort_rectify_map(const double K[9], const double D[5], const double R[9],
                                                   const double P[12], int w, int h, float* mapx, float* mapy) {
  if (!K || !D || !R || !P || w <= 0 || h <= 0 || !mapx || !mapy) return JN_ERR_ARG;
  M3 A = {{P[0], P[1], P[2]}, {P[4], P[5], P[6]}, {P[8], P[9], P[10]}};
  M3 Rm = {{R[0], R[1], R[2]}, {R[3], R[4], R[5]}, {R[6], R[7], R[8]}}, AR, iR;
  matmul3(A, Rm, AR);
  if (!inverse3(AR, iR)) { jn_set_error("jn_calib_init_undistort_rectify_map: singular P*R"); return JN_ERR_ARG; }
  const double fx = K[0], fy = K[4], u0 = K[2], v0 = K[5];
  const double k1 = D[0], k2 = D[1], p1 = D[2], p2 = D[3], k3 = D[4];
  for (int i = 0; i < h; i++) {
    double _x = i * iR[0][1] + iR[0][2], _y = i * iR[1][1] + iR[1][2], _w = i * iR[2][1] + iR[2][2];
    for (int j = 0; j < w; j++, _x += iR[0][0], _y += iR[1][0], _w += iR[2][0]) {
      const double ww = 1. / _w, x = _x * ww, y = _y * ww;
      const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
      const double kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2;
      const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
      const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
      mapx[(size_t)i * w + j] = (float)(fx * xd + u0);
      mapy[(size_t)i * w + j] = (float)(fy * yd + v0);
    }
  }
  return JN_OK;
}

// composeRotationCamToRobot / composeTranslationCamToRobot (point_cloud.cpp:76-102), what the node's -m mode
// builds XR and XT from when the extrinsics are tuned through dynamic_reconfigure (:305-311; defaults
// cfg/CamToRobotCalibParams.cfg:8-13).  The reference takes the six values as `float` and calls cos / sin on
// floats (the float overloads), stores the results in double matrices and multiplies Z * Y * X with OpenCV's
// small-matrix product; restated in that order.  Checked bit for bit against the compiled node
// (tests/test_reference_nodes_pin.py).
extern "C" int jn_calib_compose_cam_to_robot(jn_calib* c, double phi_x, double phi_y, double phi_z, double trans_x,
                                             double trans_y, double trans_z) {
  if (!c) return JN_ERR_ARG;
  const float x = (float)phi_x, y = (float)phi_y, z = (float)phi_z;
  M3 X = {{1, 0, 0}, {0, (double)cosf(x), (double)(-sinf(x))}, {0, (double)sinf(x), (double)cosf(x)}};
  M3 Y = {{(double)cosf(y), 0, (double)sinf(y)}, {0, 1, 0}, {(double)(-sinf(y)), 0, (double)cosf(y)}};
  M3 Z = {{(double)cosf(z), (double)(-sinf(z)), 0}, {(double)sinf(z), (double)cosf(z), 0}, {0, 0, 1}};
  M3 ZY, R;
  matmul3(Z, Y, ZY);
  matmul3(ZY, X, R);
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c->XR[3 * i + j] = R[i][j];
  c->XT[0] = (double)(float)trans_x;
  c->XT[1] = (double)(float)trans_y;
  c->XT[2] = (double)(float)trans_z;
  return JN_OK;
}
