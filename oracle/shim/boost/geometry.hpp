// oracle shim (test infrastructure only): the slice of Boost.Geometry the reference's marker tail uses
// (lidar_segmentation.cpp:445-562): point_xy, linestring, clear, get<>, simplify (Douglas-Peucker).
// Third-party dependency absent from /root/reference (Boost, unpinned). The published algorithm is restated:
// Douglas-Peucker keeps both end points, and recursively keeps the farthest intermediate point of a span iff its
// point-to-segment distance is strictly greater than max_distance (compared as squared distances, computed in the
// coordinate type, here float). PARITY UNPINNED for this function: no Boost build is available to check it.
#pragma once
#include <vector>
#include <cstddef>
namespace boost { namespace geometry {
namespace model {
namespace d2 {
template <class T> class point_xy {
 public:
  point_xy() : x_(0), y_(0) {}
  point_xy(T x, T y) : x_(x), y_(y) {}
  T x() const { return x_; }
  T y() const { return y_; }
 private:
  T x_, y_;
};
}  // namespace d2
template <class P> class linestring : public std::vector<P> {
 public:
  linestring& operator+=(const P& p) { this->push_back(p); return *this; }   // boost::assign's container +=
};
}  // namespace model
template <class G> void clear(G& g) { g.clear(); }
template <std::size_t I, class T> T get(const model::d2::point_xy<T>& p) { return I == 0 ? p.x() : p.y(); }

namespace detail {
template <class T> T sq_dist_point_segment(const model::d2::point_xy<T>& p, const model::d2::point_xy<T>& a,
                                           const model::d2::point_xy<T>& b) {
  T vx = b.x() - a.x(), vy = b.y() - a.y();
  T wx = p.x() - a.x(), wy = p.y() - a.y();
  T c1 = wx * vx + wy * vy;
  if (c1 <= T(0)) return wx * wx + wy * wy;
  T c2 = vx * vx + vy * vy;
  if (c2 <= c1) { T dx = p.x() - b.x(), dy = p.y() - b.y(); return dx * dx + dy * dy; }
  T t = c1 / c2;
  T qx = a.x() + t * vx, qy = a.y() + t * vy;
  T dx = p.x() - qx, dy = p.y() - qy;
  return dx * dx + dy * dy;
}
template <class T> void dp_consider(const std::vector<model::d2::point_xy<T>>& pts, std::vector<char>& keep,
                                    std::size_t first, std::size_t last, T max_sq) {
  if (last <= first + 1) return;
  T md = T(-1);
  std::size_t cand = first;
  for (std::size_t i = first + 1; i < last; i++) {
    T d = sq_dist_point_segment(pts[i], pts[first], pts[last]);
    if (d > md) { md = d; cand = i; }
  }
  if (max_sq < md) {
    keep[cand] = 1;
    dp_consider(pts, keep, first, cand, max_sq);
    dp_consider(pts, keep, cand, last, max_sq);
  }
}
}  // namespace detail

template <class T, class D>
void simplify(const model::linestring<model::d2::point_xy<T>>& in, model::linestring<model::d2::point_xy<T>>& out,
              D max_distance) {
  out.clear();
  if (in.size() <= 2 || max_distance < D(0)) { for (auto& p : in) out.push_back(p); return; }
  std::vector<char> keep(in.size(), 0);
  keep.front() = keep.back() = 1;
  T md = (T)max_distance;
  detail::dp_consider<T>(in, keep, 0, in.size() - 1, md * md);
  for (std::size_t i = 0; i < in.size(); i++) if (keep[i]) out.push_back(in[i]);
}
}}  // namespace boost::geometry
