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
  double forward_vel, rot_vel;    // navigate.cpp:28: the velocities last commanded
  float max_forward_vel;          // navigate.cpp:32 (a float in the reference; -f)
};

extern "C" jn_navigate* jn_navigate_create(void) {
  jn_navigate* n = new jn_navigate();
  n->last_dir = 0;
  n->clear_front = 0.24 + 0.8;
  n->clear_side = 0.3;
  n->laser_pt_thresh = 8;
  n->forward_vel = 0.;
  n->rot_vel = 0.;
  n->max_forward_vel = 0.6f;
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

extern "C" void jn_navigate_set_max_forward_vel(jn_navigate* n, float v) { if (n) n->max_forward_vel = v; }

// safeNavigate (navigate.cpp:302-342) without the joystick message: the mode the buttons select, the two stick
// axes, and out comes the Twist the node publishes (vel[0] = linear.x, vel[1] = angular.z).
//   JN_NAV_STOP_IN_FRONT_MANUAL  R1 + R2: stopInFrontMode(side, front)  (:208-217)  drive by stick, no forward motion
//                                                                                   while there is an obstacle
//   JN_NAV_OBSTACLE_AVOID        X:       obstacleAvoidMode(front)      (:229-255)  turn away on the spot, else forward
//   JN_NAV_STOP_IN_FRONT         O:       stopInFrontMode()             (:219-227)  full speed ahead until an obstacle
// Every mode runs checkObstacle once (the 20-frame vote advances), obstacleAvoidMode also stores the chosen
// direction; the desired velocities then pass the node's acceleration ramp (:328-337: 0.025 up / 0.1 down per
// call forward, 0.05 per call in rotation) on the velocities commanded last.
extern "C" int jn_navigate_command(jn_navigate* n, int mode, double side, double front, double vel[2]) {
  if (!n || !vel) return JN_ERR_ARG;
  const double trans_accel = 0.025, trans_decel = 0.1, rot_accel = 0.05, max_rot_vel = 1.3;   // navigate.cpp:30-34
  double desired_forward_vel, desired_rot_vel;
  if (mode == JN_NAV_STOP_IN_FRONT_MANUAL) {
    desired_forward_vel = n->max_forward_vel * front;
    desired_rot_vel = max_rot_vel * side;
    if (jn_navigate_check_obstacle(n, nullptr) == 1) desired_forward_vel = fmin(desired_forward_vel, 0.);
  } else if (mode == JN_NAV_STOP_IN_FRONT) {
    desired_forward_vel = n->max_forward_vel * 1.0;
    desired_rot_vel = 0.0;
    if (jn_navigate_check_obstacle(n, nullptr) == 1) desired_forward_vel = fmin(desired_forward_vel, 0.);
  } else if (mode == JN_NAV_OBSTACLE_AVOID) {
    if (jn_navigate_check_obstacle(n, nullptr)) {
      const int dir = jn_navigate_choose_direction(n);
      n->last_dir = dir;
      if (dir == 1) desired_rot_vel = max_rot_vel * 0.4;                  // rotate left
      else if (dir == 2) desired_rot_vel = max_rot_vel * 0.4 * (-1);      // rotate right
      else desired_rot_vel = max_rot_vel * 0.0;                           // no good direction
      desired_forward_vel = n->max_forward_vel * 0.0;                     // stop while rotating
    } else {
      desired_forward_vel = n->max_forward_vel * fmax(0.4, front);
      desired_rot_vel = max_rot_vel * 0.0;
      n->last_dir = 0;
    }
  } else {
    return JN_ERR_ARG;
  }
  if (desired_forward_vel < n->forward_vel) n->forward_vel = fmax(desired_forward_vel, n->forward_vel - trans_decel);
  else n->forward_vel = fmin(desired_forward_vel, n->forward_vel + trans_accel);
  if (desired_rot_vel < n->rot_vel) n->rot_vel = fmax(desired_rot_vel, n->rot_vel - rot_accel);
  else n->rot_vel = fmin(desired_rot_vel, n->rot_vel + rot_accel);
  vel[0] = n->forward_vel;
  vel[1] = n->rot_vel;
  return JN_OK;
}
