/*
 * jn_elas.h -- C ABI of the B200-native stereo-to-obstacle hot path.
 *
 * This is the drop-in boundary.  Every entry point
 * replaces one interface of the reference (paths relative to the reference
 * tree, sourishg/jackal-navigation):
 *
 *   jn_elas_params           <- Elas::parameters            src/elas/elas.h:59-145
 *   jn_elas_params_default   <- parameters(setting s)       src/elas/elas.h:87-144
 *   jn_elas_create/destroy   <- Elas(parameters) / ~Elas    src/elas/elas.h:148-151
 *   jn_elas_process          <- Elas::process               src/elas/elas.h:162, elas.cpp:32-151
 *   jn_elas_process_batch    <- the per-frame call site     src/obstacle_avoidance/point_cloud.cpp:416-419
 *                               (one call per left frame), batched over independent frames
 *   jn_stereo_scan_batch_host / _submit / _wait
 *                            <- imageCallbackLeft's compute  point_cloud.cpp:448 (generateDisparityMap)
 *                               + :465 (publishPointCloud -> publishObstacleScan), host buffers in and out
 *   jn_calib_load_yaml       <- cv::FileStorage reads       point_cloud.cpp:530-538
 *   jn_calib_set_q           <- stereoRectify -> Q          point_cloud.cpp:543-544
 *   jn_scan_cache_gate       <- cacheDisparityValues        point_cloud.cpp:104-147
 *   jn_scan_from_disparity   <- generateDisparityMap's convertTo(CV_8U) (point_cloud.cpp:421-422)
 *                               + publishObstacleScan(Mat&) (point_cloud.cpp:213-296)
 *   jn_points_from_disparity <- publishPointCloud (-g path) point_cloud.cpp:298-404
 *   jn_scan_from_points      <- publishObstacleScan(vector<Point3d>) point_cloud.cpp:149-211
 *
 * Plain pointers and sizes only.  All compute runs in hand-written CUDA
 * kernels for sm_100a; there is no CPU fallback: every compute entry point
 * returns JN_ERR_CUDA if no device is usable.
 */
#ifndef JN_ELAS_H
#define JN_ELAS_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same 23 fields, same order and meaning as Elas::parameters (elas.h:60-84);
 * the reference's `bool` members are int32 here (0/1). */
typedef struct jn_elas_params {
  int32_t disp_min;
  int32_t disp_max;
  float   support_threshold;
  int32_t support_texture;
  int32_t candidate_stepsize;
  int32_t incon_window_size;
  int32_t incon_threshold;
  int32_t incon_min_support;
  int32_t add_corners;
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
  int32_t filter_median;
  int32_t filter_adaptive_mean;
  int32_t postprocess_only_left;
  int32_t subsampling;
} jn_elas_params;

enum { JN_ROBOTICS = 0, JN_MIDDLEBURY = 1 };   /* Elas::setting, elas.h:56 */

/* Return codes.  The reference's process() is void; it reports "<3 support
 * points" by printing and returning with D1/D2 untouched (elas.cpp:66-71).
 * JN_FEW_SUPPORT is that case (outputs untouched, same message printed by the
 * C++ shim). */
enum {
  JN_OK            = 0,
  JN_FEW_SUPPORT   = 1,
  JN_ERR_ARG       = -1,
  JN_ERR_CUDA      = -2,
  JN_ERR_UNSUPPORTED = -3,  /* outside what the kernels cover, see "Supported envelope" below */
  JN_ERR_IO        = -4
};

/* Supported envelope.  The reference has none of these limits (it allocates per call and Triangle grows its
 * pools); inside the envelope the results are the reference's bit for bit, outside it a call fails loudly and
 * writes no output -- it never returns a different answer.
 *
 *   JN_ERR_ARG (rejected before any work is queued)
 *     width, height >= 16, bytes_per_line >= width
 *     10 <= disp_max <= 4095, disp_min <= disp_max, candidate_stepsize >= 1, 1 <= grid_size < 32768,
 *     0 <= incon_window_size <= 16, 0 <= incon_threshold <= 8191
 *   JN_ERR_UNSUPPORTED at the first process call of a size
 *     width, height < 8192              exact Delaunay predicates are evaluated in 64-bit integers
 *     ceil(sigma * sradius) <= 7        plane radius; the presets have 2 (ROBOTICS) and 3 (MIDDLEBURY)
 *     prior table entries (elas.cpp:802-806) within [-2000, 1900]   packed (energy, disparity) keys
 *     subsampling == 0 for the jn_stereo_scan_* entry points          the scan takes full-resolution maps
 *   JN_ERR_UNSUPPORTED per frame (status[i]; the frame's outputs are left untouched, the other frames of the
 *   batch are unaffected)
 *     add_corners together with more than 16 384 support points in one frame (the corner points are off the
 *       candidate lattice the large-set ordering is built on); without add_corners there is no limit below
 *       the lattice size.  The presets at 1920x1200 give 6-10 k support points.
 *     more triangles than the index field of a plane-map entry holds: 2^(31 - b) with b = bits of
 *       disp_max + 2 r + 3 (r = plane radius), i.e. millions at disp_max 255 against at most 2 * lattice size
 *       = 184 336 triangles at 1920x1200 -- reachable only above ~50 Mpixel at the presets' lattice step
 *     a point set with coincident points whose replay of Triangle's quicksort needs more than 4 096 pending
 *       partitions (seen on no input; a guard, not a known case) */

typedef struct jn_elas jn_elas;

void jn_elas_params_default(jn_elas_params* p, int setting);

/* device: CUDA ordinal.  The workspace is sized lazily at the first process call (and grown when a
 * larger batch arrives).  NULL on error (no usable device: there is no CPU path).
 *
 * Creating and destroying handles is cheap: jn_elas_destroy parks the device resources (workspace,
 * streams, staging) in a small per-process cache and the next jn_elas_create on the same device
 * takes them over, so the reference's pattern of a fresh `Elas` per frame (point_cloud.cpp:416-419)
 * does not allocate.  jn_cache_clear releases what the cache holds.
 *
 * Concurrency: a handle owns ONE workspace -- one call in flight at a time.  The asynchronous entry
 * points order their work on the stream they are given; do not issue calls on the same handle from
 * two host threads or on two unordered streams.  Use one handle per thread / per stream. */
jn_elas* jn_elas_create(const jn_elas_params* p, int device);
void     jn_elas_destroy(jn_elas* e);
void     jn_cache_clear(void);
const char* jn_last_error(void);

/* Page-locked host memory (cudaHostAlloc) for callers that do not link the CUDA runtime: image and
 * map buffers allocated here make every host<->device copy of this library a DMA transfer. */
void* jn_host_alloc(size_t bytes);
void  jn_host_free(void* p);

/* Host pointers, synchronous.  dims = {width, height, bytes_per_line}.
 * D1, D2: caller-allocated width*height floats, or (width/2)*(height/2) with
 * params.subsampling (elas.h:159-161).  D2 may be NULL (right map not wanted, as in
 * point_cloud.cpp:419-421).  Returns JN_OK, JN_FEW_SUPPORT (outputs untouched) or an error. */
int jn_elas_process(jn_elas* e, const uint8_t* I1, const uint8_t* I2,
                    float* D1, float* D2, const int32_t dims[3]);

/* Batched, device-resident.  I1/I2: n frames, frame-major, device pointers,
 * each frame height rows of bytes_per_line bytes.  D1/D2: device pointers,
 * n maps of width*height floats ((width/2)*(height/2) with subsampling), frame-major
 * (D2 may be NULL: right map not returned).
 * status: device pointer to n int32 (JN_OK / JN_FEW_SUPPORT per frame), may be
 * NULL.  Work is enqueued on `stream` (a cudaStream_t passed as void*); the
 * call does not synchronise.  A frame with <3 support points leaves its
 * D1/D2 slice untouched. */
int jn_elas_process_batch(jn_elas* e, int n, const uint8_t* I1, const uint8_t* I2,
                          float* D1, float* D2, int32_t* status,
                          const int32_t dims[3], void* stream);

/* ------------------------------------------------------------------ */
/* Calibration + obstacle scan (point_cloud.cpp)                       */

typedef struct jn_calib {
  double K1[9], K2[9];
  double D1[5], D2[5];
  double R[9];
  double T[3];
  double XR[9];
  double XT[3];
  double Q[16];       /* disparity-to-depth matrix; set by jn_calib_set_q */
  int32_t has_q;
} jn_calib;

/* Reads the OpenCV-YAML calibration file (keys K1,K2,D1,D2,R,T,XR,XT; T is a
 * bare 3-sequence, the rest !!opencv-matrix with dt: d). */
int jn_calib_load_yaml(const char* path, jn_calib* c);
/* Q = [[1,0,0,-cx],[0,1,0,-cy],[0,0,0,f],[0,0,-1/Tx,0]] (CALIB_ZERO_DISPARITY). */
void jn_calib_set_q(jn_calib* c, double cx, double cy, double f, double tx);

/* cv::stereoRectify as the reference calls it (point_cloud.cpp:543-544: flags = CALIB_ZERO_DISPARITY,
 * alpha = 0, newImageSize = rawimsize) without OpenCV: fills c->Q and returns R1, R2 (3x3), P1, P2 (3x4),
 * row-major (any output pointer may be NULL).  calib_w x calib_h = size of the calibration images
 * (640x360 in the reference), new_w x new_h = size of the rectified images (0 = same).
 * Agrees with cv2 4.13 to 1e-9 relative. */
int jn_calib_stereo_rectify(jn_calib* c, int calib_w, int calib_h, int new_w, int new_h, int zero_disparity,
                            double alpha, double R1[9], double R2[9], double P1[12], double P2[12]);
/* cv::initUndistortRectifyMap(K, D, R, P, Size(w, h), CV_32F, mapx, mapy) (point_cloud.cpp:553-554):
 * host arrays of w*h floats, the input of jn_rectify_create.  Within 1e-3 px of cv2 4.13. */
int jn_calib_init_undistort_rectify_map(const double K[9], const double D[5], const double R[9],
                                        const double P[12], int w, int h, float* mapx, float* mapy);

/* XR / XT from three Euler angles and a translation, as the node's -m mode composes them while the extrinsics are
 * tuned (composeRotationCamToRobot / composeTranslationCamToRobot, point_cloud.cpp:76-102, 305-311): XR = Z * Y * X,
 * the six values taken as float like the reference's signatures.  A jn_scan built afterwards uses them. */
int jn_calib_compose_cam_to_robot(jn_calib* c, double phi_x, double phi_y, double phi_z, double trans_x,
                                  double trans_y, double trans_z);

#define JN_SCAN_BINS 90

typedef struct jn_scan_meta {
  double angle_min, angle_max;   /* radians; 400 / -400 if no point */
  double range_min, range_max;   /* 1e9 / -500 if no point */
  int32_t n_finite;              /* number of finite bins */
  int32_t n_points;              /* pixels that passed the gate */
} jn_scan_meta;

typedef struct jn_scan jn_scan;

/* Builds the per-pixel ground-plane gate cache on the device
 * (cacheDisparityValues, point_cloud.cpp:104-147).
 * Concurrency: like a jn_elas handle, a jn_scan handle owns one set of accumulators (bins, extrema,
 * point counters) -- one call in flight at a time, one stream at a time; the gate cache itself is
 * read-only after creation.  Use one handle per thread / per stream. */
jn_scan* jn_scan_create(const jn_calib* c, int width, int height,
                        int crop_offset_x, int crop_offset_y, int device);
void     jn_scan_destroy(jn_scan* s);
/* Copies the W*H*2 gate cache (Vec2b {dmin,255}) to the host. */
int jn_scan_gate_cache(jn_scan* s, uint8_t* gate_out);

/* Device-resident batched: D float n*W*H (output of jn_elas_process_batch),
 * ranges: device n*90 doubles (1e9 = empty bin), meta: device n jn_scan_meta.
 * dmap_u8: optional device n*W*H u8 output of the convertTo(CV_8U) step. */
int jn_scan_from_disparity_batch(jn_scan* s, int n, const float* D,
                                 double* ranges, jn_scan_meta* meta,
                                 uint8_t* dmap_u8, void* stream);
/* Host-pointer, synchronous single frame. */
int jn_scan_from_disparity(jn_scan* s, const float* D, double ranges[JN_SCAN_BINS],
                           jn_scan_meta* meta, uint8_t* dmap_u8);

/* Host buffers in, obstacle scans out: n independent frames through ELAS and the scan in one call
 * (I1/I2: n frames of height rows x bytes_per_line; ranges: n*90 doubles; meta: n; optional per frame:
 * status (n int32), dmap_u8 (n*W*H, the convertTo(CV_8U) map), D1 (n*W*H floats)).  Host<->device
 * copies run on their own streams with double-buffered device staging.
 *   jn_stereo_scan_batch_host  synchronous.
 *   jn_stereo_scan_submit      queues the batch and returns; up to two submissions overlap (copy-in
 *                              of the next with the kernels of the current and copy-out of the
 *                              previous one).  All buffers belong to the library until
 *   jn_stereo_scan_wait        returns (every submitted batch has landed).
 * Pinned buffers (jn_host_alloc) are needed for the overlap; pageable ones work, serialised.
 * A frame with <3 support points reports JN_FEW_SUPPORT in status and returns an all-zero map and an
 * empty scan: what the reference's caller, which zeroes its maps before the call
 * (point_cloud.cpp:413-414), is left with. */
int jn_stereo_scan_batch_host(jn_elas* e, jn_scan* s, int n, const uint8_t* I1, const uint8_t* I2,
                              const int32_t dims[3], float* D1, int32_t* status, double* ranges,
                              jn_scan_meta* meta, uint8_t* dmap_u8);
int jn_stereo_scan_submit(jn_elas* e, jn_scan* s, int n, const uint8_t* I1, const uint8_t* I2,
                          const int32_t dims[3], float* D1, int32_t* status, double* ranges,
                          jn_scan_meta* meta, uint8_t* dmap_u8);
int jn_stereo_scan_wait(jn_elas* e);
/* The same pipeline for frames already in DEVICE memory (no copies, no caller stream): I1/I2 must be
 * complete when the call is made; D1 (n*W*H floats), status (n int32), ranges (n*90 doubles) and meta (n)
 * are device buffers too and required, dmap_u8 (n*W*H) is optional.  Submissions in flight need their own
 * output buffers; everything passed belongs to the library until jn_stereo_scan_wait returns.
 * Both submit calls cut a batch into sub-batches that roll through persistent internal streams one stage
 * apart (JN_ELAS_SPLIT sub-batches, default 2), across consecutive submissions as well: call
 * jn_stereo_scan_wait before using the handle through any other entry point. */
int jn_stereo_scan_submit_device(jn_elas* e, jn_scan* s, int n, const uint8_t* I1, const uint8_t* I2,
                                 const int32_t dims[3], float* D1, int32_t* status, double* ranges,
                                 jn_scan_meta* meta, uint8_t* dmap_u8);

/* -g path: every pixel with u8 disparity >= 2 -> robot-frame XYZ (double,
 * 3 per point, pixel order columns-outer like the reference) and the scan
 * from those points with the ground gate applied per point.
 * points: host W*H*3 doubles capacity; n_points out. */
int jn_points_from_disparity(jn_scan* s, const float* D, double* points,
                             int32_t* n_points, double ranges[JN_SCAN_BINS],
                             jn_scan_meta* meta);

/* sensor_msgs/PointCloud payload of the -g path (point_cloud.cpp:351-383), same point order as
 * jn_points_from_disparity: xyz = geometry_msgs/Point32 per point (float32 x,y,z), rgb = one
 * ChannelFloat32 value per point (bits of int32 red<<16 | green<<8 | blue taken from the left
 * image at the point's pixel).  image: host, `channels` = 3 (BGR, 3 bytes per pixel) or 1 (the
 * reference's grayscale frames, which it still indexes as Vec3b: bytes 3i..3i+2 of row j, 0
 * beyond the buffer); image_stride in bytes.  xyz: W*H*3 floats capacity, rgb: W*H floats.
 * Also returns the scan of those points like jn_points_from_disparity. */
int jn_pointcloud_from_disparity(jn_scan* s, const float* D, const uint8_t* image, int32_t image_stride,
                                 int32_t channels, float* xyz, float* rgb, int32_t* n_points,
                                 double ranges[JN_SCAN_BINS], jn_scan_meta* meta);

/* The -g path for n frames with everything in DEVICE memory (no host synchronisation, no allocation
 * after the first call): D n*W*H floats; image optional (n frames, H rows of image_stride bytes, 1 or 3
 * channels); xyz n*W*H*3 floats (frame f at f*W*H*3, its first counts[f] points valid, reference order);
 * rgb optional n*W*H floats; counts n int32; ranges n*90 doubles and meta n: the scan from those points
 * (publishObstacleScan(vector<Point3d>), point_cloud.cpp:149-211).  Asynchronous on `stream`. */
int jn_pointcloud_batch(jn_scan* s, int n, const float* D, const uint8_t* image, int32_t image_stride,
                        int32_t channels, float* xyz, float* rgb, int32_t* counts, double* ranges,
                        jn_scan_meta* meta, void* stream);

/* Compacted LaserScan.ranges as the reference publishes them: finite bins,
 * k = 89..0 (point_cloud.cpp:278-282).  Returns the count (JN_ERR_ARG for a NULL pointer). */
int jn_scan_compact(const double ranges[JN_SCAN_BINS], float* out);

/* ---- rectification ahead of the stereo path ------------
 * Replaces, per camera frame,
 *     cv::remap(tmp, leftim, lmapx, lmapy, cv::INTER_LINEAR);           point_cloud.cpp:440, 481
 *     leftim_res = leftim(Rect(crop_offset_x, crop_offset_y, w, h));    point_cloud.cpp:442, 483
 * for 8-bit single-channel frames and the default border (constant 0), bit for bit with
 * OpenCV's fixed-point bilinear remap (1/32-pixel coordinates, 15-bit weights). */
typedef struct jn_rectify jn_rectify;

/* mapx/mapy: the CV_32FC1 pair cv::initUndistortRectifyMap returns (point_cloud.cpp:553-554),
 * host pointers, map_w x map_h, row-major.  The tables are converted to OpenCV's fixed-point form
 * once and kept on `device`.  NULL on error (jn_last_error). */
jn_rectify* jn_rectify_create(const float* mapx, const float* mapy, int32_t map_w, int32_t map_h, int32_t device);
void jn_rectify_destroy(jn_rectify* r);

/* n frames, DEVICE pointers, asynchronous on `stream` (cudaStream_t).  src: n frames of src_h
 * rows, src_stride bytes per row.  roi = {x, y, w, h} in map coordinates (NULL = the whole map).
 * dst: n frames of roi h rows, dst_stride bytes per row -- rows laid out as Elas::process takes them. */
int jn_rectify_batch(jn_rectify* r, int32_t n, const uint8_t* src, int32_t src_w, int32_t src_h, int32_t src_stride,
                     const int32_t roi[4], uint8_t* dst, int32_t dst_stride, void* stream);

/* ---- compressed frames in ---------------------------------------------
 * cv::imdecode(Mat(msg->data), CV_LOAD_IMAGE_GRAYSCALE) (point_cloud.cpp:436, 478) for a batch of JPEG
 * bitstreams, decoded by nvJPEG into DEVICE buffers (frame i at dst + i*frame_stride, rows dst_stride
 * bytes apart) that jn_rectify_batch / jn_elas_process_batch read.  Luma plane = grayscale; within +-2
 * grey levels of OpenCV's decoder (IDCT rounding), not bit-exact.  Asynchronous on `stream`. */
typedef struct jn_jpeg jn_jpeg;
jn_jpeg* jn_jpeg_create(int device);
void jn_jpeg_destroy(jn_jpeg* j);
int jn_jpeg_info(jn_jpeg* j, const uint8_t* data, size_t length, int32_t* width, int32_t* height);
int jn_jpeg_decode_gray_batch(jn_jpeg* j, int n, const uint8_t* const* data, const size_t* lengths, uint8_t* dst,
                              int32_t width, int32_t height, int32_t dst_stride, size_t frame_stride, void* stream);

/* ---- the scan's consumer without ROS --------------------------------
 * laserScanCallback / checkObstacle / chooseDirection of the `navigate` node
 * (src/obstacle_avoidance/navigate.cpp:344-363, 101-153, 155-197): host code, O(90) per frame.
 *   jn_navigate_set_scan        LaserScan.ranges (compacted float32, as jn_scan_compact emits) + angle_min/max
 *   jn_navigate_set_scan_bins   the 90-bin output of jn_scan_from_disparity directly (angle_min/max rounded to
 *                               float32, as the LaserScan message between the two nodes carries them)
 *   jn_navigate_points          the laser points (x, y pairs) the node holds and publishes as Marker points
 *                               (visualizeLaserPoints, navigate.cpp:77-98); returns their number, copies at most
 *                               `capacity` of them
 *   jn_navigate_check_obstacle  returns isObstacle after the spatial filter, the 50 cm rule and the 20-frame
 *                               vote; report[4] = {points in the safe box, laser points, closest, confidence}
 *   jn_navigate_choose_direction  0 keep / 1 left / 2 right, with the reference's hysteresis on last_dir
 *                               (the caller stores its choice with jn_navigate_set_last_dir, as
 *                               obstacleAvoidMode does)
 *   jn_navigate_command         the velocity command of safeNavigate (navigate.cpp:302-342) for one of its modes:
 *                               runs the vote once, then the node's acceleration ramp on the velocities commanded
 *                               last; vel = {linear.x, angular.z} of the geometry_msgs/Twist it publishes.
 *                               side / front = the joystick axes (front only matters below 0.4 in
 *                               JN_NAV_OBSTACLE_AVOID); the waypoint mode ("navigation doesn't work yet", :317) is not offered */
enum {
  JN_NAV_STOP_IN_FRONT_MANUAL = 0,   /* R1 + R2: stopInFrontMode(side, front), navigate.cpp:208-217 */
  JN_NAV_OBSTACLE_AVOID       = 1,   /* X:       obstacleAvoidMode(front),      navigate.cpp:229-255 */
  JN_NAV_STOP_IN_FRONT        = 2    /* O:       stopInFrontMode(),             navigate.cpp:219-227 */
};
typedef struct jn_navigate jn_navigate;
jn_navigate* jn_navigate_create(void);
void jn_navigate_destroy(jn_navigate* n);
void jn_navigate_set_clearance(jn_navigate* n, double clear_front, double clear_side, int laser_pt_thresh);
void jn_navigate_set_last_dir(jn_navigate* n, int dir);
int  jn_navigate_last_dir(const jn_navigate* n);
int  jn_navigate_set_scan(jn_navigate* n, const float* ranges, int count, double angle_min, double angle_max);
int  jn_navigate_set_scan_bins(jn_navigate* n, const double ranges[JN_SCAN_BINS], const jn_scan_meta* meta);
int  jn_navigate_points(const jn_navigate* n, double* xy, int capacity);
int  jn_navigate_check_obstacle(jn_navigate* n, double report[4]);
int  jn_navigate_choose_direction(const jn_navigate* n);
void jn_navigate_set_max_forward_vel(jn_navigate* n, float max_forward_vel);     /* -f, navigate.cpp:32 */
int  jn_navigate_command(jn_navigate* n, int mode, double side, double front, double vel[2]);

#ifdef __cplusplus
}
#endif
#endif /* JN_ELAS_H */
