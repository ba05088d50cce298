// ros/urf_node.cpp — thin ROS1 glue that drops liburf_b200 into the existing urban_road_filter graph.
//
// Same node name, same subscribed topic, same five published topics and the same LidarFilters.cfg dynamic_reconfigure
// surface as the reference node (src/main.cpp:1-56, src/lidar_segmentation.cpp:51-65,601-621): only the body of the
// scan callback changes — it hands the cloud to urf_process() (include/urf.h) instead of running the CPU detectors, then
// rebuilds the reference's four clouds and its road_marker MarkerArray from the labels, the emission order and the marker
// vertices. ROS / PCL are not available in the build container: this file is compile- and run-checked against the shim
// headers of oracle/shim by tests/test_glue.py (on the GPU box its published output is compared with the fixtures the
// unmodified reference produced); in a catkin workspace it builds unchanged against the real headers.
#include <cstring>
#include <string>
#include <vector>

#include <ros/ros.h>
#include <visualization_msgs/Marker.h>
#include <visualization_msgs/MarkerArray.h>
#include <dynamic_reconfigure/server.h>
#include <urban_road_filter/LidarFiltersConfig.h>
#include <pcl_conversions/pcl_conversions.h>
#include <pcl/point_cloud.h>
#include <pcl_ros/point_cloud.h>
#include <pcl/point_types.h>

#include "urf.h"

namespace urf_glue {

urf_params g_params;          // what paramsCallback last received (the reference keeps them in params:: globals)
bool g_params_dirty = true;

// paramsCallback, src/main.cpp:4-34: same fields, same order; narrowing to float happens inside urf_set_params
void paramsCallback(urban_road_filter::LidarFiltersConfig& config, uint32_t /*level*/) {
  urf_params& p = g_params;
  std::snprintf(p.fixed_frame, sizeof(p.fixed_frame), "%s", config.fixed_frame.c_str());
  std::snprintf(p.topic_name, sizeof(p.topic_name), "%s", config.topic_name.c_str());
  p.x_zero_method = config.x_zero_method;
  p.z_zero_method = config.z_zero_method;
  p.star_shaped_method = config.star_shaped_method;
  p.blind_spots = config.blind_spots;
  p.xDirection = config.xDirection;
  p.interval = config.interval;
  p.curb_height = config.curb_height;
  p.curb_points = config.curb_points;
  p.beamZone = config.beamZone;
  p.cylinder_deg_x = config.cylinder_deg_x;
  p.cylinder_deg_z = config.cylinder_deg_z;
  p.curb_slope_deg = config.curb_slope_deg;
  p.min_x = config.min_x; p.max_x = config.max_x;
  p.min_y = config.min_y; p.max_y = config.max_y;
  p.min_z = config.min_z; p.max_z = config.max_z;
  p.kdev_param = config.kdev_param;
  p.kdist_param = config.kdist_param;
  p.starbeam_filter = config.starbeam_filter;
  p.dmin_param = config.dmin_param;
  p.simple_poly_allow = config.simple_poly_allow;
  p.poly_s_param = config.poly_s_param;
  p.poly_z_avg_allow = config.poly_z_avg_allow;
  p.poly_z_manual = config.poly_z_manual;
  g_params_dirty = true;
  ROS_INFO("Updated params %s", ros::this_node::getName().c_str());
}

class Detector {
 public:
  // device: CUDA device index; max_points: largest scan the sensor can produce; channels: ring count (reference constant 64)
  Detector(ros::NodeHandle* nh, int device = 0, int max_points = 1 << 20, int channels = 64) : channels_(channels) {
    const int rc = urf_create(&ctx_, device, max_points, 1);
    if (rc != URF_OK) ROS_INFO("urf_create failed: %s", urf_strerror(rc));
    sub_ = nh->subscribe(std::string(g_params.topic_name), 1, &Detector::filtered, this);        // lidar_segmentation.cpp:53
    pub_road_ = nh->advertise<pcl::PCLPointCloud2>("road", 1);                                     // :55-59
    pub_high_ = nh->advertise<pcl::PCLPointCloud2>("curb", 1);
    pub_box_ = nh->advertise<pcl::PCLPointCloud2>("roi", 1);
    pub_pobroad_ = nh->advertise<pcl::PCLPointCloud2>("road_probably", 1);
    pub_marker_ = nh->advertise<visualization_msgs::MarkerArray>("road_marker", 1);
    ROS_INFO("Ready");
  }
  ~Detector() { urf_destroy(ctx_); }

  void set_ghostcount(int g) { ghostcount_ = g; }
  int ghostcount() const { return ghostcount_; }

  // scan callback, replaces Detector::filtered (lidar_segmentation.cpp:95-622)
  void filtered(const pcl::PointCloud<pcl::PointXYZI>& cloud) {
    if (!ctx_) return;
    if (g_params_dirty) {
      g_params.channels = channels_;
      if (urf_set_params(ctx_, &g_params) != URF_OK) return;
      g_params_dirty = false;
    }
    const int n = (int)cloud.points.size();
    xyzi_.resize((size_t)4 * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) {                      // first 16 bytes of the 32-byte PointXYZI record, intensity in w
      const pcl::PointXYZI& p = cloud.points[i];
      xyzi_[4 * i] = p.x; xyzi_[4 * i + 1] = p.y; xyzi_[4 * i + 2] = p.z; xyzi_[4 * i + 3] = p.intensity;
    }
    label_.resize(n > 0 ? n : 1); order_.resize(n > 0 ? n : 1);
    ring_start_.resize(URF_MAX_CHANNELS + 1);
    urf_result res;
    std::memset(&res, 0, sizeof(res));
    res.label = label_.data(); res.order = order_.data(); res.ring_start = ring_start_.data();
    if (urf_process(ctx_, xyzi_.data(), n, &res) != URF_OK) return;
    if (res.status == URF_TOO_FEW_POINTS) return;      // lidar_segmentation.cpp:124-126: nothing is published

    pcl::PointCloud<pcl::PointXYZI> road, high, probably;
    auto box = boost::make_shared<pcl::PointCloud<pcl::PointXYZI>>();
    for (int i = 0; i < n; i++) if (label_[i] >= 0) box->push_back(cloud.points[i]);               // roi cloud, input order
    for (int k = 0; k < res.n_order; k++) {                                                        // :354-367
      const int i = order_[k];
      if (label_[i] == URF_LABEL_ROAD) road.push_back(cloud.points[i]);
      else if (label_[i] == URF_LABEL_CURB) high.push_back(cloud.points[i]);
    }
    if (res.n_rings > 10)                                                                          // :605-608
      for (int k = ring_start_[10]; k < ring_start_[11]; k++) probably.push_back(cloud.points[order_[k]]);

    // road_marker, lidar_segmentation.cpp:369-602
    if (res.n_vert > 2) {
      urf_strip strips[URF_MAX_VERTS * 2 + 64];
      double pts[3 * 4 * URF_MAX_VERTS];
      int npts = 0;
      const int ns = urf_build_markers(&g_params, res.vert, res.n_vert, &ghostcount_, strips, URF_MAX_VERTS * 2 + 64, pts,
                                       4 * URF_MAX_VERTS, &npts);
      if (ns >= 0) {
        visualization_msgs::MarkerArray ma;
        for (int s = 0; s < ns; s++) {
          visualization_msgs::Marker m;
          m.header.frame_id = g_params.fixed_frame;                                               // :424-427
          m.header.stamp = ros::Time();
          m.type = visualization_msgs::Marker::LINE_STRIP;
          m.action = strips[s].action == 2 ? visualization_msgs::Marker::DELETE : visualization_msgs::Marker::ADD;
          m.id = strips[s].id;
          m.pose.orientation.w = 1.0;                                                              // marker_init, :25-39
          m.scale.x = m.scale.y = m.scale.z = 0.5;
          m.color.r = strips[s].red ? 1.0f : 0.0f; m.color.g = strips[s].red ? 0.0f : 1.0f; m.color.b = 0.0f; m.color.a = 1.0f;
          m.lifetime = ros::Duration(0);
          for (int k = 0; k < strips[s].count; k++) {
            geometry_msgs::Point q;
            q.x = pts[3 * (strips[s].first + k)]; q.y = pts[3 * (strips[s].first + k) + 1]; q.z = pts[3 * (strips[s].first + k) + 2];
            m.points.push_back(q);
          }
          ma.markers.push_back(m);
        }
        pub_marker_.publish(ma);                                                                   // :601
      }
    }
    road.header = cloud.header; probably.header = cloud.header; high.header = cloud.header; box->header = cloud.header;   // :612-615
    pub_road_.publish(road);                                                                       // :618-621
    pub_high_.publish(high);
    pub_box_.publish(box);
    pub_pobroad_.publish(probably);
  }

 private:
  urf_ctx* ctx_ = nullptr;
  int channels_;
  int ghostcount_ = 0;                                 // lidar_segmentation.cpp:23
  std::vector<float> xyzi_;
  std::vector<int32_t> label_, order_, ring_start_;
  ros::Publisher pub_road_, pub_high_, pub_box_, pub_pobroad_, pub_marker_;
  ros::Subscriber sub_;
};

}  // namespace urf_glue

#ifndef URF_GLUE_NO_MAIN
int main(int argc, char** argv) {                        // src/main.cpp:37-56
  ros::init(argc, argv, "urban_road_filt");
  ROS_INFO("Initializing %s", ros::this_node::getName().c_str());
  urf_default_params(&urf_glue::g_params);
  dynamic_reconfigure::Server<urban_road_filter::LidarFiltersConfig> server;
  dynamic_reconfigure::Server<urban_road_filter::LidarFiltersConfig>::CallbackType f = &urf_glue::paramsCallback;
  server.setCallback(f);
  ros::NodeHandle nh;
  urf_glue::Detector detector(&nh);
  ros::spin();
  return 0;
}
#endif
