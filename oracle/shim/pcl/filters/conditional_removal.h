// oracle shim (test infrastructure only): pcl::ConditionBase / pcl::ConditionalRemoval restated.
// Third-party dependency absent from /root/reference: PCL (unpinned; README.md:7-8 names ROS Kinetic/Melodic => PCL 1.7/1.8).
// Published behaviour restated: with keep_organized=false, filter() drops points whose x/y/z is not finite, keeps points
// for which the condition evaluates true, preserves input order, and tolerates output aliasing the input.
#pragma once
#include <cmath>
#include <pcl/point_cloud.h>
namespace pcl {
template <class PointT> class ConditionBase {
 public:
  typedef boost::shared_ptr<ConditionBase<PointT>> Ptr;
  virtual ~ConditionBase() {}
  virtual bool evaluate(const PointT& point) const = 0;
};
template <class PointT> class ConditionalRemoval {
 public:
  void setCondition(typename ConditionBase<PointT>::Ptr c) { cond_ = c; }
  void setInputCloud(const boost::shared_ptr<PointCloud<PointT>>& in) { in_ = in; }
  void filter(PointCloud<PointT>& out) {
    std::vector<PointT> kept;
    kept.reserve(in_->points.size());
    for (const PointT& p : in_->points) {
      if (!std::isfinite(p.x) || !std::isfinite(p.y) || !std::isfinite(p.z)) continue;
      if (cond_->evaluate(p)) kept.push_back(p);
    }
    PCLHeader h = in_->header;
    out.points.swap(kept);
    out.header = h;
    out.width = (uint32_t)out.points.size();
    out.height = 1;
    out.is_dense = true;
  }
 private:
  typename ConditionBase<PointT>::Ptr cond_;
  boost::shared_ptr<PointCloud<PointT>> in_;
};
}
