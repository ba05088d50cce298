// tests/kat/glue_entry.cpp — TEST CODE. Builds ros/urf_node.cpp (the product's ROS glue) against the shim ROS/PCL headers
// of oracle/shim and exposes the same C entry as oracle/ref_entry.cpp, so tests/test_glue.py can feed a cloud through the
// glue node's scan callback and compare everything it publishes with what the unmodified reference published.
#define URF_GLUE_NO_MAIN
#include "../../oracle/shim/shim_capture.h"
#include "../../ros/urf_node.cpp"

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
