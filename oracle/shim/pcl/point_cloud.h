// oracle shim (test infrastructure only): the slice of pcl::PointCloud<T> the reference touches.
#pragma once
#include <vector>
#include <memory>
#include <pcl/point_types.h>
namespace boost {
template <class T> using shared_ptr = std::shared_ptr<T>;
template <class T, class... A> std::shared_ptr<T> make_shared(A&&... a) { return std::make_shared<T>(std::forward<A>(a)...); }
}
namespace pcl {
template <class PointT> class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT>> Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  size_t size() const { return points.size(); }
  void clear() { points.clear(); width = height = 0; }
};
}
