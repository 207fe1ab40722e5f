// stub of <pcl_conversions/pcl_conversions.h>: pcl::toROSMsg / pcl::fromROSMsg through the drop-in's cloud I/O
// (include/CloudIO_b200.hpp: the gather into PointXYZRGB rows runs on the GPU)
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <class Cloud>
inline void toROSMsg(const Cloud& cloud, sensor_msgs::PointCloud2& msg);
template <class Cloud>
inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, Cloud& cloud);
}  // namespace pcl

#include "CloudIO_b200.hpp"
namespace pcl {
template <class Cloud>
inline void toROSMsg(const Cloud& cloud, sensor_msgs::PointCloud2& msg) {
  gicpb_shim::toROSMsg(cloud, msg);
}
template <class Cloud>
inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, Cloud& cloud) {
  gicpb_shim::fromROSMsg(msg, cloud);
}
}  // namespace pcl
