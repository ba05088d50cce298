// oracle shim (test infrastructure only): stand-in for the header catkin generates from cfg/LidarFilters.cfg.
// Field names/defaults follow /root/reference/cfg/LidarFilters.cfg:10-84.
#pragma once
#include <string>
namespace urban_road_filter {
struct LidarFiltersConfig {
  std::string fixed_frame = "left_os1/os1_lidar";
  std::string topic_name = "/left_os1/os1_cloud_node/points";
  bool x_zero_method = true, z_zero_method = true, star_shaped_method = true, blind_spots = true;
  int xDirection = 0;
  double interval = 0.18, curb_height = 0.05;
  int curb_points = 5;
  double beamZone = 30;
  double min_x = 0, max_x = 30, min_y = -10, max_y = 10, min_z = -3, max_z = -1;
  double cylinder_deg_x = 150, cylinder_deg_z = 140, curb_slope_deg = 50;
  double kdev_param = 1.225, kdist_param = 2;
  bool starbeam_filter = false;
  int dmin_param = 10;
  bool simple_poly_allow = true;
  double poly_s_param = 0.7, poly_z_manual = -1.5;
  bool poly_z_avg_allow = true;
};
}
