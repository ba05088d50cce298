// oracle/ref_entry.cpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// C entry points around the UNMODIFIED reference: /root/reference/src/{lidar_segmentation,star_shaped_search,
// x_zero_method,z_zero_method,blind_spots}.cpp are compiled in place against oracle/shim/ and linked with this file into
// oracle/_ref/liburf_ref.so. Nothing in the product path may load it; tests, smoke() and bench.py's cpu_baseline do.
//
// Observability: Detector::filtered() keeps array2D/array3D local, so per-point labels are recovered from
// the four published clouds. The harness stashes the input index in the padding float data[3] of each PointXYZI; the
// reference copies `p` whole (lidar_segmentation.cpp:147,238,358), so the index survives to the publishers.
#include <pthread.h>
#include <chrono>
#include <cstring>
#include <vector>

#include "urban_road_filter/data_structures.hpp"
#include "shim_capture.h"
#include "../include/urf.h"

extern int channels;     // lidar_segmentation.cpp:4 (non-const, external linkage)
extern int ghostcount;   // lidar_segmentation.cpp:23

namespace {

void apply_params(const urf_params* p) {
  // Same narrowing assignments as paramsCallback, src/main.cpp:5-32.
  params::fixedFrame = p->fixed_frame;
  params::topicName = p->topic_name;
  params::x_zero_method = p->x_zero_method != 0;
  params::z_zero_method = p->z_zero_method != 0;
  params::star_shaped_method = p->star_shaped_method != 0;
  params::blind_spots = p->blind_spots != 0;
  params::xDirection = p->xDirection;
  params::interval = p->interval;
  params::curbHeight = p->curb_height;
  params::curbPoints = p->curb_points;
  params::beamZone = p->beamZone;
  params::angleFilter1 = p->cylinder_deg_x;
  params::angleFilter2 = p->cylinder_deg_z;
  params::angleFilter3 = p->curb_slope_deg;
  params::min_X = p->min_x;
  params::max_X = p->max_x;
  params::min_Y = p->min_y;
  params::max_Y = p->max_y;
  params::min_Z = p->min_z;
  params::max_Z = p->max_z;
  params::kdev_param = p->kdev_param;
  params::kdist_param = p->kdist_param;
  params::starbeam_filter = p->starbeam_filter != 0;
  params::dmin_param = p->dmin_param;
  params::polysimp_allow = p->simple_poly_allow != 0;
  params::polysimp = p->poly_s_param;
  params::zavg_allow = p->poly_z_avg_allow != 0;
  params::polyz = p->poly_z_manual;
  channels = p->channels;
}

Detector* detector() {
  static ros::NodeHandle nh;
  static Detector* d = new Detector(&nh);   // runs beam_init() once, like the node
  return d;
}

inline int32_t id_of(const pcl::PointXYZI& p) { int32_t v; std::memcpy(&v, &p.data[3], 4); return v; }

struct Job {
  const float* xyzi; int n; const urf_params* prm; int repeat; double seconds;
};

void* run_job(void* arg) {
  Job* j = static_cast<Job*>(arg);
  pcl::PointCloud<pcl::PointXYZI> cloud;
  cloud.points.resize(j->n);
  for (int i = 0; i < j->n; i++) {
    pcl::PointXYZI& p = cloud.points[i];
    p.x = j->xyzi[4 * i + 0]; p.y = j->xyzi[4 * i + 1]; p.z = j->xyzi[4 * i + 2];
    p.intensity = j->xyzi[4 * i + 3];
    int32_t id = i; std::memcpy(&p.data[3], &id, 4);
  }
  cloud.width = j->n; cloud.height = 1;
  apply_params(j->prm);
  Detector* d = detector();
  auto t0 = std::chrono::steady_clock::now();
  for (int r = 0; r < j->repeat; r++) {
    shim::capture().reset();
    d->filtered(cloud);
  }
  j->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return nullptr;
}

// Detector::filtered puts channels*piece*48 B on the heap but `float markerPointsArray[piece][4]` on the stack
// (lidar_segmentation.cpp:298): run on a thread with a big stack.
double run_on_big_stack(Job& j) {
  pthread_attr_t attr;
  pthread_attr_init(&attr);
  pthread_attr_setstacksize(&attr, (size_t)1 << 30);
  pthread_t th;
  pthread_create(&th, &attr, run_job, &j);
  pthread_join(th, nullptr);
  pthread_attr_destroy(&attr);
  return j.seconds;
}

}  // namespace

extern "C" {

// Runs Detector::filtered() once on the n x 4 float cloud and fills:
//   label[n]      -1 outside ROI, else 0/1/2 as recovered from roi/road/curb clouds
//   emit[n]       input indices in the order road/curb points were emitted, merged (label 1 and 2 interleaved are not
//                 recoverable, so road ids come first, then curb ids): n_road, n_curb returned through counts[]
//   prob[n]       input indices of the road_probably cloud (ring 10, azimuth order); count in counts[]
//   counts[8]     0: published (0/1)  1: n_roi  2: n_road  3: n_curb  4: n_prob  5: n_markers  6: n_marker_points  7: markers published
//   strips[max_strips], strip_points[3*max_strip_points]: the MarkerArray (ADD strips then DELETE ghosts)
// Returns 0, or -1 if a buffer was too small.
int urf_ref_run(const float* xyzi, int n, const urf_params* prm, int32_t* label, int32_t* emit, int32_t* prob,
                int32_t* counts, urf_strip* strips, int max_strips, double* strip_points, int max_strip_points) {
  Job j{xyzi, n, prm, 1, 0.0};
  run_on_big_stack(j);
  shim::Capture& c = shim::capture();
  for (int i = 0; i < 8; i++) counts[i] = 0;
  for (int i = 0; i < n; i++) label[i] = URF_LABEL_OUTSIDE;
  bool published = c.cloud_seen.count("roi") > 0;
  counts[0] = published ? 1 : 0;
  if (!published) return 0;
  const auto& roi = c.clouds["roi"].points;
  const auto& road = c.clouds["road"].points;
  const auto& curb = c.clouds["curb"].points;
  const auto& pr = c.clouds["road_probably"].points;
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
  return 0;
}

// Times `repeat` back-to-back Detector::filtered() calls on one cloud (publishers capture into memory, i.e. are
// no-ops as far as ROS goes). Returns seconds for all repeats. Used by bench.py's cpu_baseline / --impl reference.
double urf_ref_time(const float* xyzi, int n, const urf_params* prm, int repeat) {
  Job j{xyzi, n, prm, repeat, 0.0};
  return run_on_big_stack(j);
}

void urf_ref_set_ghostcount(int g) { ghostcount = g; }
int urf_ref_get_ghostcount(void) { return ghostcount; }

}  // extern "C"
