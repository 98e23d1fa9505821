"""TEST INFRASTRUCTURE ONLY -- plain-Python restatement of the `navigate` node's scan consumer
(src/obstacle_avoidance/navigate.cpp), used by tests/ to check jn_navigate_*.  PINNED against the reference's own
code: navigate.cpp compiled where it lies with stand-in ROS headers (oracle/standins, oracle/navigate_ref_shim.cpp ->
oracle/_ref/libnavigate_ref.so); tests/test_reference_nodes_pin.py runs 1 500 scans through the node's
laserScanCallback / checkObstacle / chooseDirection and through this class: equal laser points (bit for bit), votes,
printed reports and directions.

  laser_scan_callback   navigate.cpp:344-363
  check_obstacle        navigate.cpp:101-153
  choose_direction      navigate.cpp:155-197
"""
import math
from collections import deque

INF = int(1e9)                      # navigate.cpp:48


class Navigate:
    def __init__(self):
        self.laser_points = []          # :22
        self.commands = deque()         # :45
        self.last_dir = 0               # :46
        self.clear_front = 0.24 + 0.8   # :37
        self.clear_side = 0.3           # :38
        self.laser_pt_thresh = 8        # :42

    def laser_scan_callback(self, ranges_f32, angle_min, angle_max):
        num_points = len(ranges_f32)                                        # :345
        self.laser_points = []
        for i in range(num_points):                                         # :356
            angle = float(i) * (angle_max - angle_min) / float(num_points) + angle_min
            r = float(ranges_f32[i])
            self.laser_points.append((r * math.cos(angle), r * math.sin(angle)))

    def check_obstacle(self):
        count = 0
        is_obstacle = 0
        closest = float(INF)
        for x, y in self.laser_points:                                      # :105
            dist = math.sqrt(x * x + y * y)
            closest = min(closest, dist)
            if 0. < x < self.clear_front and -self.clear_side < y < self.clear_side:
                count += 1
        if count > self.laser_pt_thresh:                                    # :115
            is_obstacle = 1
        if closest < 0.5:                                                   # :126
            is_obstacle = 1
        if len(self.commands) < 20:                                         # :130
            self.commands.append(is_obstacle)
        else:
            self.commands.popleft()
            self.commands.append(is_obstacle)
        one = sum(1 for c in self.commands if c == 1)
        zero = len(self.commands) - one
        if one > 2:                                                         # :146
            is_obstacle = 1
        conf = float(one) / float(one + zero)
        return is_obstacle, (count, len(self.laser_points), closest, conf)

    def choose_direction(self):
        left = right = 0
        for x, y in self.laser_points:                                      # :157
            if 0. < x < self.clear_front:
                if y < 0:
                    right += 1
                else:
                    left += 1
        if left + right < 2:                                                # :167
            return 0
        conf_left = 2. * float(right) / float(left + right)
        conf_right = 2. * float(left) / float(left + right)
        d = 0
        if conf_left > conf_right:                                          # :175
            if self.last_dir != 1:
                d = 1 if conf_left - conf_right > 0.5 else self.last_dir
            else:
                d = 1
        else:
            if self.last_dir != 2:
                d = 2 if conf_right - conf_left > 0.5 else self.last_dir
            else:
                d = 2
        return d
