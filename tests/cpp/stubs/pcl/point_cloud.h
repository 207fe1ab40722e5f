// stub of <pcl/point_cloud.h>: pcl::PointCloud<PointT> as far as the reference's tests and the drop-in headers use it
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <class PointT>
class PointCloud {
 public:
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;  // boost::shared_ptr in PCL 1.8
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = 0; height = 0; }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  typename std::vector<PointT>::iterator begin() { return points.begin(); }
  typename std::vector<PointT>::iterator end() { return points.end(); }
};
}  // namespace pcl
