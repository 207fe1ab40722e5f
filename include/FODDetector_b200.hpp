// FODDetector_b200.hpp - header-only C++ drop-ins for the two callers either side of the registration path
// (SURVEY section 8f rows 1 and 3), on top of the C ABI of libgicp_b200.so (include/gicp_b200.h):
//
//   class FODDetector            reference include/FODDetector.h:24-104, src/FODDetector.cpp:21-115 - same
//                                constructor, setters and clusterPossibleFODs / getFODIndices /
//                                fodIndicesToPointCloud, so src/LeicaStateMachine.cpp:200-205 and
//                                test/test_fod_detector.cpp compile unchanged (fodIndicesToROSMsg only with
//                                -DGICPB_WITH_ROS)
//   gicpb_shim::downsampleCloud  Filter::downsampleCloud, reference src/Filter.cpp:91-105 (pcl::VoxelGrid)
//
// With -DGICPB_WITH_PCL the cluster indices are std::vector<pcl::PointIndices> as in the reference; without PCL a
// layout-free stand-in with the same `indices` member is used.
#pragma once
#ifndef FOD_DETECTOR_B200_HPP_
#define FOD_DETECTOR_B200_HPP_

#include "GICPAlignment_b200.hpp"

#ifdef GICPB_WITH_PCL
#include <pcl/PointIndices.h>
#endif

namespace gicpb_shim {

#ifdef GICPB_WITH_PCL
typedef pcl::PointIndices PointIndicesT;
#else
struct PointIndices {
  std::vector<int> indices;
};
typedef PointIndices PointIndicesT;
#endif

// Filter::downsampleCloud (reference src/Filter.cpp:91-105): pcl::VoxelGrid with leaf (l, l, l), every field
// downsampled; centroids in ascending voxel index, height 1, is_dense true.
template <class CloudPtr>
inline void downsampleCloud(const CloudPtr& cloud, const CloudPtr& cloud_downsampled, double leaf_size,
                            Context* shared = nullptr) {
  log(kInfo, "Downsample cloud with leaf_size : %f", leaf_size);
  std::shared_ptr<Context> hold;
  shared = resolve(shared, hold);
  CloudT out;
  const int64_t n = (int64_t)cloud->points.size();
  if (n > 0) {
    out.points.resize((size_t)n);
    int64_t m = 0;
    shared->check(gicpb_voxel_grid(shared->get(), &cloud->points[0].x, n, (int64_t)sizeof(cloud->points[0]), 0, leaf_size,
                                   &out.points[0].x, &m),
                  "gicpb_voxel_grid");
    out.points.resize((size_t)m);
  }
  out.width = (uint32_t)out.points.size();
  out.height = 1;
  out.is_dense = true;
  *cloud_downsampled = out;
}

}  // namespace gicpb_shim

class FODDetector {
 public:
  typedef gicpb_shim::CloudT PointCloudRGB;
#ifdef GICPB_WITH_PCL
  typedef PointCloudRGB::Ptr CloudPtr;
#else
  typedef std::shared_ptr<PointCloudRGB> CloudPtr;
#endif
  typedef gicpb_shim::PointIndicesT PointIndices;

  // reference src/FODDetector.cpp:21-26
  FODDetector(CloudPtr cloud, double cluster_tolerance, double min_fod_points)
      : ctx_holder_(gicpb_shim::Context::shared()), ctx_(*ctx_holder_), cloud_(cloud) {
    setClusterTolerance(cluster_tolerance);
    setMinFODpoints(min_fod_points);
  }
  ~FODDetector() {}

  // :28-38 - a zero tolerance is invalid and falls back to 4e-3
  void setClusterTolerance(double tolerance) {
    if (tolerance == 0) {
      gicpb_shim::log(gicpb_shim::kWarn, "FODDetector: invalid tolerance value: %f", tolerance);
      cluster_tolerance_ = 4e-3;
    } else {
      cluster_tolerance_ = tolerance;
    }
    gicpb_shim::log(gicpb_shim::kInfo, "FODDetector: cluster tolerance set to: %f", cluster_tolerance_);
  }
  // :40-43
  void setMinFODpoints(double min_fod_points) { min_cluster_size_ = min_fod_points; }

  // :45-58 - pcl::EuclideanClusterExtraction::extract on the GPU: tolerance as set, min size as set (the int
  // conversion of setMinClusterSize), no upper size limit (setMaxClusterSize is commented out in the reference)
  void clusterPossibleFODs() {
    cluster_indices_.clear();
    const int64_t n = cloud_ ? (int64_t)cloud_->points.size() : 0;
    std::vector<int32_t> labels((size_t)n, -1);
    int64_t nc = 0;
    ctx_.check(gicpb_euclidean_clusters(ctx_.get(), n ? (const void*)&cloud_->points[0].x : nullptr, n,
                                        (int64_t)sizeof(gicpb_shim::CloudT().points[0]), 0, cluster_tolerance_,
                                        (int64_t)(int)min_cluster_size_, 0, labels.data(), &nc),
               "gicpb_euclidean_clusters");
    cluster_indices_.resize((size_t)nc);
    for (int64_t i = 0; i < n; ++i)
      if (labels[(size_t)i] >= 0) cluster_indices_[(size_t)labels[(size_t)i]].indices.push_back((int)i);
    gicpb_shim::log(gicpb_shim::kInfo, "cluster_indices_size: %zd", cluster_indices_.size());
  }

  // :112-115
  void getFODIndices(std::vector<PointIndices>& fod_indices) { fod_indices = cluster_indices_; }

  // :60-78 - one cloud per cluster appended to fod_cloud_array; returns their number
  int fodIndicesToPointCloud(std::vector<CloudPtr>& fod_cloud_array) {
    int n_fods = 0;
    for (size_t k = 0; k < cluster_indices_.size(); ++k) {
      CloudPtr cloud_cluster(new PointCloudRGB);
      for (size_t j = 0; j < cluster_indices_[k].indices.size(); ++j)
        cloud_cluster->points.push_back(cloud_->points[(size_t)cluster_indices_[k].indices[j]]);
      cloud_cluster->width = (uint32_t)cloud_cluster->points.size();
      cloud_cluster->height = 1;
      cloud_cluster->is_dense = true;
      gicpb_shim::log(gicpb_shim::kInfo, "cluster_size: %zd", cloud_cluster->points.size());
      fod_cloud_array.push_back(cloud_cluster);
      n_fods++;
    }
    return n_fods;
  }

#ifdef GICPB_WITH_ROS
  // :80-110
  int fodIndicesToROSMsg(std::vector<sensor_msgs::PointCloud2>& fod_msg_array) {
    std::vector<CloudPtr> fod_cloud_array;
    const int n_fods = fodIndicesToPointCloud(fod_cloud_array);
    sensor_msgs::PointCloud2 cluster_msg;
    for (const auto& cloud : fod_cloud_array) {
      pcl::toROSMsg(*cloud, cluster_msg);
      fod_msg_array.push_back(cluster_msg);
    }
    return n_fods;
  }
#endif

 private:
  std::shared_ptr<gicpb_shim::Context> ctx_holder_;  // the process-wide context (GICPAlignment_b200.hpp)
  gicpb_shim::Context& ctx_;
  CloudPtr cloud_;
  double cluster_tolerance_;
  double min_cluster_size_;
  std::vector<PointIndices> cluster_indices_;
};

#endif  // FOD_DETECTOR_B200_HPP_
