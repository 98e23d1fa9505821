/*
 * pointcloud_ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * extern "C" access to the UNMODIFIED reference src/obstacle_avoidance/point_cloud.cpp, compiled where it lies
 * against the stand-in headers of oracle/standins (see standins.h for what they do and do not pin) and linked
 * with the reference's own ELAS objects (the same ones as libelas_ref.so).  The functions below set the node's
 * file-scope state the way its main() does (point_cloud.cpp:530-559), CALL the reference's functions
 *     cacheDisparityValues                       :104-147
 *     publishPointCloud -> publishObstacleScan(Mat&)              :298-304, 213-296     (default path)
 *     publishPointCloud -> publishObstacleScan(vector<Point3d>)   :298-404, 149-211     (-g path)
 *     generateDisparityMap                       :406-429
 * and hand back what they wrote or published (the messages a ros::Publisher received).  No arithmetic here.
 *
 * The reference indexes scan[k] with k = floor(90 * (45 - theta_deg) / 90) unchecked: a point outside
 * the +-45 degree field of view, or a NaN angle (disparity 0 on a pixel whose gate wrapped to 0), writes outside a
 * stack array.  The tests only pass inputs for which every k is inside [0, 89] (they check that first with the
 * restatement, which can count such points); nothing here guards it -- that would be modifying the reference.
 */
#include "standins.h"

#include <sstream>

#define main jn_reference_pointcloud_main   /* the node's main() is compiled, never run */
#include "point_cloud.cpp"                  /* reference translation unit, in place (-I$(REF)/src/obstacle_avoidance -I$(REF)/src/elas) */
#undef main

namespace {

struct Quiet {          /* the reference prints progress to std::cout */
  std::ostringstream os;
  std::streambuf* old;
  Quiet() : old(std::cout.rdbuf(os.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};

Mat from_doubles(const double* p, int r, int c) {
  Mat m(r, c, CV_64FC1);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) m.at<double>(i, j) = p[i * c + j];
  return m;
}

/* W x H view of a zero-filled buffer with three spare rows below: the reference reads a grayscale frame as Vec3b
 * (:371-383), i.e. up to 2 W bytes past the last row; those reads see zeros here, as jn_pointcloud_from_disparity
 * defines them. */
Mat padded_image(const uint8_t* src, int W, int H, int stride, int channels) {
  Mat big(H + 3, W, channels == 3 ? CV_8UC3 : CV_8UC1, Scalar(0, 0, 0));
  for (int r = 0; r < H; r++) std::memcpy(big.data + (size_t)r * big.step, src + (size_t)r * stride, (size_t)W * channels);
  return big(Rect(0, 0, W, H));
}

void copy_scan(const sensor_msgs::LaserScan& s, float* ranges, int32_t* n, float meta[4]) {
  *n = (int32_t)s.ranges.size();
  for (size_t i = 0; i < s.ranges.size(); i++) ranges[i] = s.ranges[i];
  meta[0] = s.angle_min; meta[1] = s.angle_max; meta[2] = s.range_min; meta[3] = s.range_max;
}

}  // namespace

extern "C" {

/* main()'s set-up after stereoRectify (:546-559): Q, XR, XT, V, pos, the crop geometry. */
void ref_pc_setup(const double* Q16, const double* XR9, const double* XT3, int W, int H, int ox, int oy) {
  Q = from_doubles(Q16, 4, 4);
  XR = from_doubles(XR9, 3, 3);
  XT = from_doubles(XT3, 3, 1);
  V = Mat(4, 1, CV_64FC1);
  pos = Mat(4, 1, CV_64FC1);
  crop_im_width = W;
  crop_im_height = H;
  crop_offset_x = ox;
  crop_offset_y = oy;
  logging = false;
  calib_robot_to_cam = 0;
  leftim_res = Mat(H, W, CV_8UC1, Scalar(0));
  jn_standin::captured() = jn_standin::Captured();
}

/* composeRotationCamToRobot / composeTranslationCamToRobot (:76-102) as publishPointCloud calls them in -m mode
 * (:305-311): from the dynamic_reconfigure values (doubles, converted to the functions' float parameters). */
void ref_pc_compose(double phi_x, double phi_y, double phi_z, double tx, double ty, double tz, double* XR9, double* XT3) {
  Mat r = composeRotationCamToRobot(phi_x, phi_y, phi_z);
  Mat t = composeTranslationCamToRobot(tx, ty, tz);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) XR9[3 * i + j] = r.at<double>(i, j);
    XT3[i] = t.at<double>(i, 0);
  }
}

/* cacheDisparityValues(); gate_out receives valid_disp (H x W x Vec2b). */
void ref_pc_cache_gate(uint8_t* gate_out) {
  Quiet q;
  cacheDisparityValues();
  for (int r = 0; r < valid_disp.rows; r++)
    std::memcpy(gate_out + (size_t)r * valid_disp.cols * 2, valid_disp.data + (size_t)r * valid_disp.step, (size_t)valid_disp.cols * 2);
}

void ref_pc_set_gate(const uint8_t* gate) {
  valid_disp = Mat(crop_im_height, crop_im_width, CV_8UC2, Scalar(255, 3));
  for (int r = 0; r < valid_disp.rows; r++)
    std::memcpy(valid_disp.data + (size_t)r * valid_disp.step, gate + (size_t)r * valid_disp.cols * 2, (size_t)valid_disp.cols * 2);
}

/* Default path: publishPointCloud(dmap, seq) with gen_pcl = 0 -> publishObstacleScan(Mat&, seq).
 * ranges: the LaserScan.ranges the node published (compacted, <= 90 floats), meta = angle_min, angle_max,
 * range_min, range_max as the message carries them (float32). */
void ref_pc_scan(const uint8_t* dmap_u8, float* ranges, int32_t* n, float meta[4]) {
  Quiet q;
  gen_pcl = 0;
  Mat dmap(crop_im_height, crop_im_width, CV_8UC1);
  std::memcpy(dmap.data, dmap_u8, (size_t)crop_im_width * crop_im_height);
  jn_standin::captured().scans.clear();
  publishPointCloud(dmap, seq);
  copy_scan(jn_standin::captured().scans.back(), ranges, n, meta);
}

/* -g path: publishPointCloud(dmap, seq) with gen_pcl = 1 -> PointCloud message + publishObstacleScan(points, seq).
 * xyz: 3 floats per point (geometry_msgs/Point32), rgb: the "rgb" channel values; both in publication order. */
void ref_pc_pointcloud(const uint8_t* dmap_u8, const uint8_t* image, int stride, int channels, float* xyz, float* rgb,
                       int32_t* n_points, float* ranges, int32_t* n, float meta[4]) {
  Quiet q;
  gen_pcl = 1;
  leftim_res = padded_image(image, crop_im_width, crop_im_height, stride, channels);
  Mat dmap(crop_im_height, crop_im_width, CV_8UC1);
  std::memcpy(dmap.data, dmap_u8, (size_t)crop_im_width * crop_im_height);
  jn_standin::captured().scans.clear();
  jn_standin::captured().clouds.clear();
  publishPointCloud(dmap, seq);
  const sensor_msgs::PointCloud& pc = jn_standin::captured().clouds.back();
  *n_points = (int32_t)pc.points.size();
  for (size_t i = 0; i < pc.points.size(); i++) {
    xyz[3 * i] = pc.points[i].x; xyz[3 * i + 1] = pc.points[i].y; xyz[3 * i + 2] = pc.points[i].z;
    rgb[i] = pc.channels[0].values[i];
  }
  copy_scan(jn_standin::captured().scans.back(), ranges, n, meta);
  gen_pcl = 0;
}

/* generateDisparityMap(leftim_res, rightim_res): the reference's Elas with its default parameters +
 * postprocess_only_left, then convertTo(CV_8U).  out_u8: W x H. */
void ref_pc_generate_disparity(const uint8_t* I1, const uint8_t* I2, uint8_t* out_u8) {
  Quiet q;
  const int W = crop_im_width, H = crop_im_height;
  Mat left(H, W, CV_8UC1), right(H, W, CV_8UC1);
  std::memcpy(left.data, I1, (size_t)W * H);
  std::memcpy(right.data, I2, (size_t)W * H);
  Mat show = generateDisparityMap(left, right);
  for (int r = 0; r < H; r++) std::memcpy(out_u8 + (size_t)r * W, show.data + (size_t)r * show.step, (size_t)W);
}

}  // extern "C"
