// tests/kat/glue_entry.cpp — TEST CODE. Builds ros/urf_node.cpp (the product's ROS glue) against the shim ROS/PCL headers
// of oracle/shim and exposes the same C entry as oracle/ref_entry.cpp, so tests/test_glue.py can feed a cloud through the
// glue node's scan callback and compare everything it publishes with what the unmodified reference published.
#define URF_GLUE_NO_MAIN
#include "../../oracle/shim/shim_capture.h"
#include "../../ros/urf_node.cpp"
#include "../../ros/urf_node_cloud2.cpp"

namespace { inline int32_t id_of(const pcl::PointXYZI& p) { int32_t v; std::memcpy(&v, &p.data[3], 4); return v; } }

extern "C" int urf_glue_run(const float* xyzi, int n, const urf_params* prm, int max_points, int ghost_in, int32_t* label,
                            int32_t* emit, int32_t* prob, int32_t* counts, urf_strip* strips, int max_strips,
                            double* strip_points, int max_strip_points) {
  static ros::NodeHandle nh;
  urf_glue::g_params = *prm;
  urf_glue::g_params_dirty = true;
  urf_glue::Detector det(&nh, 0, max_points, prm->channels);
  det.set_ghostcount(ghost_in);
  pcl::PointCloud<pcl::PointXYZI> cloud;
  cloud.points.resize(n);
  for (int i = 0; i < n; i++) {
    pcl::PointXYZI& p = cloud.points[i];
    p.x = xyzi[4 * i]; p.y = xyzi[4 * i + 1]; p.z = xyzi[4 * i + 2]; p.intensity = xyzi[4 * i + 3];
    int32_t id = i; std::memcpy(&p.data[3], &id, 4);
  }
  shim::capture().reset();
  det.filtered(cloud);
  shim::Capture& c = shim::capture();
  for (int i = 0; i < 8; i++) counts[i] = 0;
  for (int i = 0; i < n; i++) label[i] = URF_LABEL_OUTSIDE;
  const bool published = c.cloud_seen.count("roi") > 0;
  counts[0] = published;
  if (!published) return 0;
  const auto& roi = c.clouds["roi"].points; const auto& road = c.clouds["road"].points;
  const auto& curb = c.clouds["curb"].points; const auto& pr = c.clouds["road_probably"].points;
  counts[1] = (int)roi.size(); counts[2] = (int)road.size(); counts[3] = (int)curb.size(); counts[4] = (int)pr.size();
  for (const auto& p : roi) label[id_of(p)] = URF_LABEL_NONE;
  int k = 0;
  for (const auto& p : road) { label[id_of(p)] = URF_LABEL_ROAD; emit[k++] = id_of(p); }
  for (const auto& p : curb) { label[id_of(p)] = URF_LABEL_CURB; emit[k++] = id_of(p); }
  k = 0;
  for (const auto& p : pr) prob[k++] = id_of(p);
  counts[7] = c.markers_seen ? 1 : 0;
  int np = 0, ns = 0;
  for (const auto& m : c.markers.markers) {
    if (ns >= max_strips || np + (int)m.points.size() > max_strip_points) return -1;
    urf_strip& s = strips[ns++];
    s.id = m.id; s.action = m.action; s.red = (m.color.r == 1.0f) ? 1 : 0; s.first = np; s.count = (int)m.points.size();
    for (const auto& q : m.points) { strip_points[3 * np] = q.x; strip_points[3 * np + 1] = q.y; strip_points[3 * np + 2] = q.z; np++; }
  }
  counts[5] = ns; counts[6] = np;
  return det.ghostcount();
}

// The PointCloud2-in / PointCloud2-out node (ros/urf_node_cloud2.cpp): the cloud goes in as a sensor_msgs/PointCloud2 with
// `step`-byte records (x, y, z at 0 / 4 / 8, intensity at 16 when step >= 20; the intensity carries the point's index so
// that the published records can be traced back), the four published PointCloud2 messages are decoded again.
extern "C" int urf_glue_run_cloud2(const float* xyzi, int n, const urf_params* prm, int max_points, int ghost_in, int step, int32_t* label,
                                   int32_t* emit, int32_t* prob, int32_t* counts, urf_strip* strips, int max_strips,
                                   double* strip_points, int max_strip_points) {
  static ros::NodeHandle nh;
  urf_glue::g_params = *prm;
  urf_glue::g_params_dirty = true;
  urf_glue::DetectorCloud2 det(&nh, 0, max_points, prm->channels);
  det.set_ghostcount(ghost_in);
  sensor_msgs::PointCloud2 msg;
  msg.header.frame_id = "sensor"; msg.header.seq = 42;
  msg.height = 1; msg.width = (uint32_t)n; msg.point_step = (uint32_t)step; msg.row_step = (uint32_t)(step * n);
  const char* names[4] = {"x", "y", "z", "intensity"};
  const uint32_t offs[4] = {0, 4, 8, 16};
  for (int k = 0; k < 4; k++) { sensor_msgs::PointField f; f.name = names[k]; f.offset = offs[k]; f.datatype = sensor_msgs::PointField::FLOAT32; f.count = 1; msg.fields.push_back(f); }
  msg.data.assign((size_t)step * n, 0xA5);
  for (int i = 0; i < n; i++) {
    uint8_t* rec = msg.data.data() + (size_t)i * step;
    std::memcpy(rec, xyzi + 4 * i, 12);
    const float id = (float)i;                              // exact below 2^24
    std::memcpy(rec + 16, &id, 4);
  }
  shim::capture().reset();
  det.filtered(msg);
  shim::Capture& c = shim::capture();
  for (int i = 0; i < 8; i++) counts[i] = 0;
  for (int i = 0; i < n; i++) label[i] = URF_LABEL_OUTSIDE;
  const bool published = c.cloud_seen.count("roi") > 0;
  counts[0] = published;
  if (!published) return 0;
  auto ids = [&](const char* topic, std::vector<int32_t>* out) {
    const sensor_msgs::PointCloud2& m = c.clouds2[topic];
    if (m.point_step != 32 || m.fields.size() != 4 || m.fields[3].offset != 16 || m.header.frame_id != "sensor" || m.header.seq != 42 ||
        m.data.size() != (size_t)m.width * 32) return false;
    for (uint32_t k = 0; k < m.width; k++) {
      float rec[8]; std::memcpy(rec, m.data.data() + (size_t)k * 32, 32);
      const int i = (int)rec[4];
      if (i < 0 || i >= n || rec[0] != xyzi[4 * i] || rec[1] != xyzi[4 * i + 1] || rec[2] != xyzi[4 * i + 2] || rec[3] != 1.0f) return false;
      out->push_back(i);
    }
    return true;
  };
  std::vector<int32_t> roi, road, curb, pr;
  if (!ids("roi", &roi) || !ids("road", &road) || !ids("curb", &curb) || !ids("road_probably", &pr)) return -2;
  counts[1] = (int)roi.size(); counts[2] = (int)road.size(); counts[3] = (int)curb.size(); counts[4] = (int)pr.size();
  for (int i : roi) label[i] = URF_LABEL_NONE;
  int k = 0;
  for (int i : road) { label[i] = URF_LABEL_ROAD; emit[k++] = i; }
  for (int i : curb) { label[i] = URF_LABEL_CURB; emit[k++] = i; }
  k = 0;
  for (int i : pr) prob[k++] = i;
  counts[7] = c.markers_seen ? 1 : 0;
  int np = 0, ns = 0;
  for (const auto& m : c.markers.markers) {
    if (ns >= max_strips || np + (int)m.points.size() > max_strip_points) return -1;
    urf_strip& s = strips[ns++];
    s.id = m.id; s.action = m.action; s.red = (m.color.r == 1.0f) ? 1 : 0; s.first = np; s.count = (int)m.points.size();
    for (const auto& q : m.points) { strip_points[3 * np] = q.x; strip_points[3 * np + 1] = q.y; strip_points[3 * np + 2] = q.z; np++; }
  }
  counts[5] = ns; counts[6] = np;
  return det.ghostcount();
}
