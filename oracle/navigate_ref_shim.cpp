/*
 * navigate_ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * extern "C" access to the UNMODIFIED reference src/obstacle_avoidance/navigate.cpp, compiled where it lies
 * against the stand-in headers of oracle/standins (ROS message types as plain structs; the file uses nothing of
 * OpenCV but cv::Point2d).  The functions below CALL the reference's laserScanCallback (navigate.cpp:344-363),
 * checkObstacle (:101-153) and chooseDirection (:155-197) on the reference's own file-scope state; they contain
 * no arithmetic of their own.  checkObstacle reports count / closest / confidence only on stdout (:151), so the
 * line it prints is captured and returned verbatim.
 */
#include "standins.h"

#include <sstream>

#define main jn_reference_navigate_main     /* the node's main() is compiled, never run */
#include "navigate.cpp"                     /* reference translation unit, in place (-I$(REF)/src/obstacle_avoidance) */
#undef main

extern "C" {

/* The state a freshly started node has (navigate.cpp:21-45). */
void ref_nav_reset(void) {
  laserPoints.clear();
  laserScan.clear();
  laserAngles.clear();
  commands.clear();
  last_dir = 0;
  clear_front = 0.24 + 0.8;
  clear_side = 0.3;
  laser_pt_thresh = 8;
  forward_vel = 0.;
  rot_vel = 0.;
  max_forward_vel = 0.6;
}

void ref_nav_set_clearance(double front, double side, int thresh) {
  clear_front = front;
  clear_side = side;
  laser_pt_thresh = thresh;
}

void ref_nav_set_last_dir(int d) { last_dir = d; }
int ref_nav_last_dir(void) { return last_dir; }

/* sensor_msgs/LaserScan as point_cloud.cpp publishes it -> laserScanCallback */
void ref_nav_laser_scan(const float* ranges, int n, float angle_min, float angle_max) {
  std::shared_ptr<sensor_msgs::LaserScan> m(new sensor_msgs::LaserScan());
  m->ranges.assign(ranges, ranges + n);
  m->angle_min = angle_min;
  m->angle_max = angle_max;
  laserScanCallback(m);
}

int ref_nav_point_count(void) { return (int)laserPoints.size(); }
void ref_nav_points(double* xy) {
  for (size_t i = 0; i < laserPoints.size(); i++) { xy[2 * i] = laserPoints[i].x; xy[2 * i + 1] = laserPoints[i].y; }
}

/* returns isObstacle; line receives what checkObstacle printed ("count, points, Y|N, closest, conf") */
int ref_nav_check_obstacle(char* line, int cap) {
  std::ostringstream os;
  std::streambuf* old = std::cout.rdbuf(os.rdbuf());
  const int r = checkObstacle();
  std::cout.rdbuf(old);
  if (line && cap > 0) {
    std::string s = os.str();
    std::strncpy(line, s.c_str(), (size_t)cap - 1);
    line[cap - 1] = 0;
  }
  return r;
}

int ref_nav_choose_direction(void) { return chooseDirection(); }

/* obstacleAvoidMode (navigate.cpp:229-257): the caller of both, which stores the choice in last_dir */
int ref_nav_obstacle_avoid_mode(double front, double vel[2]) {
  std::ostringstream os;
  std::streambuf* old = std::cout.rdbuf(os.rdbuf());
  std::pair<double, double> v = obstacleAvoidMode(front);
  std::cout.rdbuf(old);
  vel[0] = v.first;
  vel[1] = v.second;
  return last_dir;
}

/* safeNavigate (navigate.cpp:302-342) on a sensor_msgs/Joy with the given buttons pressed (R2 = 9, R1 = 11,
 * triangle = 12, O = 13, X = 14) and stick axes; returns 1 and the published Twist's linear.x / angular.z, or 0 if
 * the node published nothing (no mode button). */
int ref_nav_safe_navigate(int r1, int r2, int x, int o, float side, float front, double vel[2]) {
  std::shared_ptr<sensor_msgs::Joy> m(new sensor_msgs::Joy());
  m->buttons.assign(17, 0);
  m->axes.assign(4, 0.f);
  m->buttons[9] = r2; m->buttons[11] = r1; m->buttons[14] = x; m->buttons[13] = o;
  m->axes[0] = side; m->axes[1] = front;
  jn_standin::captured().twists.clear();
  std::ostringstream os;
  std::streambuf* old = std::cout.rdbuf(os.rdbuf());
  safeNavigate(m);
  std::cout.rdbuf(old);
  if (jn_standin::captured().twists.empty()) return 0;
  vel[0] = jn_standin::captured().twists.back().linear.x;
  vel[1] = jn_standin::captured().twists.back().angular.z;
  return 1;
}

void ref_nav_set_max_forward_vel(float v) { max_forward_vel = v; }

}  // extern "C"
