// urf_markers.cpp — host routine for the reference's marker tail (lidar_segmentation.cpp:369-602): flag edge rules and
// smoothing, green/red line-strip splitting, running z average, optional Douglas-Peucker simplification and the ghost
// DELETE bookkeeping. At most 361 vertices per scan, so this stays on the CPU next to the ROS glue.
//
// Boost.Geometry is not part of the reference tree (unpinned system dependency). `simplify` is restated from its
// published algorithm (Douglas-Peucker, keep a point iff its squared point-to-segment distance is strictly greater than
// max_distance^2, computed in the coordinate type float) — PARITY UNPINNED for simple_poly_allow=1, see DESIGN.md.
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

#include "../../include/urf.h"

namespace {

struct XY { float x, y; };

float sq_dist_seg(const XY& p, const XY& a, const XY& b) {
  const float vx = b.x - a.x, vy = b.y - a.y, wx = p.x - a.x, wy = p.y - a.y;
  const float c1 = wx * vx + wy * vy;
  if (c1 <= 0.f) return wx * wx + wy * wy;
  const float c2 = vx * vx + vy * vy;
  if (c2 <= c1) { const float dx = p.x - b.x, dy = p.y - b.y; return dx * dx + dy * dy; }
  const float t = c1 / c2;
  const float dx = p.x - (a.x + t * vx), dy = p.y - (a.y + t * vy);
  return dx * dx + dy * dy;
}

// Douglas-Peucker without recursion: a work list of open spans [first, last] whose end points are kept; the farthest
// intermediate point of a span (strictly farther than every earlier one) is kept iff its squared distance to the span's
// chord exceeds max_distance^2, and splits the span in two. (oracle/shim/boost/geometry.hpp states the same published
// algorithm recursively, for the reference build; tests/test_markers.py checks both against the algorithm's properties —
// end points kept, output nested in max_distance, every dropped point within max_distance of the output — and each other.)
std::vector<XY> simplify(const std::vector<XY>& in, float max_distance) {
  const size_t n = in.size();
  if (n <= 2 || max_distance < 0.f) return in;
  const float max_sq = max_distance * max_distance;
  std::vector<char> keep(n, 0);
  keep[0] = keep[n - 1] = 1;
  std::vector<std::pair<size_t, size_t>> open;
  open.emplace_back(0, n - 1);
  while (!open.empty()) {
    const std::pair<size_t, size_t> span = open.back();
    open.pop_back();
    if (span.second <= span.first + 1) continue;
    float far = -1.f;
    size_t at = span.first;
    for (size_t i = span.first + 1; i < span.second; i++) {
      const float d = sq_dist_seg(in[i], in[span.first], in[span.second]);
      if (far < d) { far = d; at = i; }
    }
    if (max_sq < far) {
      keep[at] = 1;
      open.emplace_back(at, span.second);
      open.emplace_back(span.first, at);
    }
  }
  std::vector<XY> out;
  out.reserve(n);
  for (size_t i = 0; i < n; i++) if (keep[i]) out.push_back(in[i]);
  return out;
}

struct P3 { double x, y, z; };

}  // namespace

extern "C" int urf_build_markers(const urf_params* prm, const float (*vert)[4], int n_vert, int* ghostcount,
                                 urf_strip* strips, int max_strips, double* points_xyz, int max_points,
                                 int* n_points_out) {
  if (!prm || !ghostcount || !strips || !points_xyz || !n_points_out || n_vert < 0 || (n_vert > 0 && !vert))
    return URF_ERR_INVALID;
  *n_points_out = 0;
  const int cM = n_vert;
  if (!(cM > 2)) return 0;                                          // :371 — nothing is published, ghostcount untouched
  const bool polysimp_allow = prm->simple_poly_allow != 0, zavg_allow = prm->poly_z_avg_allow != 0;
  const float polysimp = (float)prm->poly_s_param, polyz = (float)prm->poly_z_manual;   // main.cpp:30,32 narrowing
  std::vector<float> f(cM);
  for (int i = 0; i < cM; i++) f[i] = vert[i][3];
  // :381-397 first/last point adopt the colour of their neighbour
  if (f[0] == 0 && f[1] == 1) f[0] = 1;
  if (f[cM - 1] == 0 && f[cM - 2] == 1) f[cM - 1] = 1;
  if (f[0] == 1 && f[1] == 0) f[0] = 0;
  if (f[cM - 1] == 1 && f[cM - 2] == 0) f[cM - 1] = 0;
  // :402-415 two in-place passes: lone green between reds -> red, lone red between greens -> green
  for (int i = 2; i <= cM - 3; i++) if (f[i] == 0 && f[i - 1] == 1 && f[i + 1] == 1) f[i] = 1;
  for (int i = 2; i <= cM - 3; i++) if (f[i] == 1 && f[i - 1] == 0 && f[i + 1] == 0) f[i] = 0;

  int ns = 0, np = 0;
  bool overflow = false;
  std::vector<P3> cur;        // line_strip.points
  std::vector<XY> line;       // boost linestring of the current strip
  int lineStripID = 0;
  int marker_id = 0;          // line_strip.id (Marker default 0)
  float zavg = 0.0f;
  auto close_strip = [&](int id, int red) {                          // :458-489 / :501-526 / :537-562
    marker_id = id;
    std::vector<P3> pts = cur;
    if (polysimp_allow) {
      pts.clear();
      for (const XY& q : simplify(line, polysimp)) pts.push_back(P3{(double)q.x, (double)q.y, (double)polyz});
    }
    if (ns >= max_strips || np + (int)pts.size() > max_points) { overflow = true; return; }
    strips[ns].id = id; strips[ns].action = 0; strips[ns].red = red; strips[ns].first = np; strips[ns].count = (int)pts.size();
    ns++;
    for (const P3& q : pts) { points_xyz[3 * np] = q.x; points_xyz[3 * np + 1] = q.y; points_xyz[3 * np + 2] = q.z; np++; }
    cur.clear();
    line.clear();
  };
  auto push = [&](const P3& q) { cur.push_back(q); line.push_back(XY{(float)q.x, (float)q.y}); };   // :444-445
  for (int i = 0; i < cM; i++) {
    const P3 pt{vert[i][0], vert[i][1], vert[i][2]};                 // :433-435
    zavg *= i; zavg += pt.z; zavg /= i + 1;                          // :436-438
    if (i == 0) push(pt);
    else if (f[i] == f[i - 1]) {                                     // :450
      push(pt);
      if (i == cM - 1) close_strip(lineStripID, f[i] == 0 ? 0 : 1);
    } else if (f[i] == 0) {                                          // :495 red -> green: the joining segment is still red
      push(pt);
      close_strip(lineStripID, 1);
      lineStripID++;
      push(pt);
    } else {                                                         // :534 green -> red
      close_strip(lineStripID, 0);
      lineStripID++;
      push(P3{vert[i - 1][0], vert[i - 1][1], vert[i - 1][2]});
      push(pt);
    }
  }
  if (overflow) return URF_ERR_CAPACITY;
  if (zavg_allow) for (int k = 0; k < np; k++) points_xyz[3 * k + 2] = zavg;   // :580-589
  // :592-598 delete markers of previous scans that no longer exist
  for (int del = lineStripID; del < *ghostcount; del++) {
    marker_id++;
    if (ns >= max_strips) return URF_ERR_CAPACITY;
    strips[ns].id = marker_id; strips[ns].action = 2; strips[ns].red = 0; strips[ns].first = np; strips[ns].count = 0;
    ns++;
  }
  *ghostcount = lineStripID;
  *n_points_out = np;
  return ns;
}
