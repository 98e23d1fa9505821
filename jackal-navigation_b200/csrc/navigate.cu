// navigate.cu -- the consumer side of the obstacle scan without ROS: laserScanCallback,
// checkObstacle and chooseDirection of the reference's `navigate` node
// (src/obstacle_avoidance/navigate.cpp:344-363, 101-153, 155-197).
//
// O(90) scalar work per frame: host code, no kernel.  It closes the loop of a
// camera-to-command demo: the 90-bin scan of
// jn_scan_from_disparity goes in, "obstacle ahead / turn left / turn right" comes out.
// The arithmetic follows the reference statement by statement (float ranges widened to double,
// uniform angular spacing over the COMPACTED ranges, 20-frame vote, hysteresis on the last turn).
#include <math.h>
#include <deque>
#include <vector>
#include "common.cuh"

struct jn_navigate {
  std::vector<double> px, py;     // laserPoints (navigate.cpp:22)
  std::deque<int> commands;       // last 20 classifications (navigate.cpp:45)
  int last_dir;                   // navigate.cpp:46
  double clear_front, clear_side; // navigate.cpp:37-38
  int laser_pt_thresh;            // navigate.cpp:42
};

extern "C" jn_navigate* jn_navigate_create(void) {
  jn_navigate* n = new jn_navigate();
  n->last_dir = 0;
  n->clear_front = 0.24 + 0.8;
  n->clear_side = 0.3;
  n->laser_pt_thresh = 8;
  return n;
}
extern "C" void jn_navigate_destroy(jn_navigate* n) { delete n; }

extern "C" void jn_navigate_set_clearance(jn_navigate* n, double clear_front, double clear_side, int laser_pt_thresh) {
  if (!n) return;
  n->clear_front = clear_front; n->clear_side = clear_side; n->laser_pt_thresh = laser_pt_thresh;
}
extern "C" void jn_navigate_set_last_dir(jn_navigate* n, int dir) { if (n) n->last_dir = dir; }
extern "C" int jn_navigate_last_dir(const jn_navigate* n) { return n ? n->last_dir : 0; }

// laserScanCallback (navigate.cpp:344-363): ranges = LaserScan.ranges (float32, the compacted finite
// bins in the order jn_scan_compact emits them), angle_min/max from the scan.
extern "C" int jn_navigate_set_scan(jn_navigate* n, const float* ranges, int count, double angle_min, double angle_max) {
  if (!n || count < 0 || (count > 0 && !ranges)) return JN_ERR_ARG;
  const unsigned int numPoints = (unsigned int)count;
  n->px.resize(numPoints);
  n->py.resize(numPoints);
  for (int i = 0; i < count; i++) {
    const double angle = (double)i * (angle_max - angle_min) / (double)numPoints + angle_min;
    n->px[i] = ranges[i] * cos(angle);
    n->py[i] = ranges[i] * sin(angle);
  }
  return JN_OK;
}

// The 90-bin scan of jn_scan_from_disparity straight into the vote: compaction as the reference
// publishes it (point_cloud.cpp:278-282), then laserScanCallback.  Between the two nodes the scan travels as a
// sensor_msgs/LaserScan, whose angle_min / angle_max are float32 (point_cloud.cpp:271-272 assign doubles to them):
// the angles are rounded the same way here, so the laser points are the ones the navigate node computes.
extern "C" int jn_navigate_set_scan_bins(jn_navigate* n, const double ranges[JN_SCAN_BINS], const jn_scan_meta* meta) {
  if (!n || !ranges || !meta) return JN_ERR_ARG;
  float tmp[JN_SCAN_BINS];
  const int count = jn_scan_compact(ranges, tmp);
  return jn_navigate_set_scan(n, tmp, count, (double)(float)meta->angle_min, (double)(float)meta->angle_max);
}

// laserPoints (navigate.cpp:22, filled at :356-361): what visualizeLaserPoints (:77-98) publishes as Marker points.
extern "C" int jn_navigate_points(const jn_navigate* n, double* xy, int capacity) {
  if (!n || capacity < 0 || (capacity > 0 && !xy)) return JN_ERR_ARG;
  const int count = (int)n->px.size();
  for (int i = 0; i < count && i < capacity; i++) { xy[2 * i] = n->px[i]; xy[2 * i + 1] = n->py[i]; }
  return count;
}

// checkObstacle (navigate.cpp:101-153).  Returns isObstacle (0/1); report = {count in the safe box,
// number of laser points, closest distance, confidence of the 20-frame vote}.
extern "C" int jn_navigate_check_obstacle(jn_navigate* n, double report[4]) {
  if (!n) return JN_ERR_ARG;
  const int INF = 1000000000;
  int count = 0, isObstacle = 0;
  double closestObst = INF;
  for (size_t i = 0; i < n->px.size(); i++) {
    const double dist = sqrt(n->px[i] * n->px[i] + n->py[i] * n->py[i]);
    closestObst = fmin(closestObst, dist);
    if (n->px[i] > 0. && n->px[i] < n->clear_front && n->py[i] > -n->clear_side && n->py[i] < n->clear_side) count++;
  }
  if (count > n->laser_pt_thresh) isObstacle = 1;            // spatial filter
  if (closestObst < 0.5) isObstacle = 1;                     // anything closer than 50 cm
  if (n->commands.size() < 20) {
    n->commands.push_back(isObstacle);
  } else {
    n->commands.pop_front();
    n->commands.push_back(isObstacle);
  }
  int one = 0, zero = 0;
  for (int c : n->commands) {
    if (c == 1) one++;
    else zero++;
  }
  if (one > 2) isObstacle = 1;                               // temporal filter
  if (report) {
    report[0] = count;
    report[1] = (double)n->px.size();
    report[2] = closestObst;
    report[3] = (double)one / (double)(one + zero);
  }
  return isObstacle;
}

// chooseDirection (navigate.cpp:155-197): 0 = keep, 1 = left, 2 = right.  The reference's caller
// stores the choice in last_dir itself; jn_navigate_set_last_dir does that here.
extern "C" int jn_navigate_choose_direction(const jn_navigate* n) {
  if (!n) return JN_ERR_ARG;
  int left_count = 0, right_count = 0;
  for (size_t i = 0; i < n->px.size(); i++) {
    if (n->px[i] > 0. && n->px[i] < n->clear_front) {
      if (n->py[i] < 0) right_count++;
      else left_count++;
    }
  }
  if (left_count + right_count < 2) return 0;
  const double conf_left = 2. * (double)right_count / (double)(left_count + right_count);
  const double conf_right = 2. * (double)left_count / (double)(left_count + right_count);
  int dir = 0;
  if (conf_left > conf_right) {
    if (n->last_dir != 1) dir = (conf_left - conf_right > 0.5) ? 1 : n->last_dir;
    else dir = 1;
  } else {
    if (n->last_dir != 2) dir = (conf_right - conf_left > 0.5) ? 2 : n->last_dir;
    else dir = 2;
  }
  return dir;
}
