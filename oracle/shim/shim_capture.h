// oracle shim (test infrastructure only): what the reference node published during one Detector::filtered() call.
#pragma once
#include <map>
#include <string>
#include <pcl/point_cloud.h>
#include "shim_msgs.h"
#include <sensor_msgs/PointCloud2.h>
namespace shim {
struct Capture {
  std::map<std::string, pcl::PointCloud<pcl::PointXYZI>> clouds;   // by topic: road, curb, roi, road_probably
  std::map<std::string, bool> cloud_seen;
  std::map<std::string, sensor_msgs::PointCloud2> clouds2;         // what ros/urf_node_cloud2.cpp publishes, by topic
  visualization_msgs::MarkerArray markers;
  bool markers_seen = false;
  void reset() { clouds.clear(); cloud_seen.clear(); clouds2.clear(); markers.markers.clear(); markers_seen = false; }
};
inline Capture& capture() { static Capture c; return c; }
template <class M> struct Sink;
template <> struct Sink<pcl::PointCloud<pcl::PointXYZI>> {
  static void put(const std::string& t, const pcl::PointCloud<pcl::PointXYZI>& c) { capture().clouds[t] = c; capture().cloud_seen[t] = true; }
};
template <> struct Sink<std::shared_ptr<pcl::PointCloud<pcl::PointXYZI>>> {
  static void put(const std::string& t, const std::shared_ptr<pcl::PointCloud<pcl::PointXYZI>>& c) { capture().clouds[t] = *c; capture().cloud_seen[t] = true; }
};
template <> struct Sink<sensor_msgs::PointCloud2> {
  static void put(const std::string& t, const sensor_msgs::PointCloud2& c) { capture().clouds2[t] = c; capture().cloud_seen[t] = true; }
};
template <> struct Sink<visualization_msgs::MarkerArray> {
  static void put(const std::string&, const visualization_msgs::MarkerArray& m) { capture().markers = m; capture().markers_seen = true; }
};
template <class M> void capture_publish(const std::string& topic, const M& msg) { Sink<M>::put(topic, msg); }
}  // namespace shim
