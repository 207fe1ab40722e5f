// stands in for the reference's include/Utils.h (which pulls ros/ros.h, pcl_ros and the viewer): the static helpers the
// reference's tests call.  rotateCloud builds its matrix and moves the points with the CPU oracle (fixture side);
// getNormals / computeCloudResolution go through the drop-in (include/GICPAlignment_b200.hpp).
#pragma once
#ifndef _UTILS_H
#define _UTILS_H

#include "ros/ros.h"
#include "ros/package.h"
#include "pcl_conversions/pcl_conversions.h"
#include <pcl_ros/point_cloud.h>

#include <Viewer.h>

#include "GICPAlignment_b200.hpp"

extern "C" {
void orc_rotation_rpy(double roll, double pitch, double yaw, float* T16);
void orc_transform(const float* T16, const float* in, int n, float* out);
}

class Utils {
  typedef pcl::PointCloud<pcl::PointXYZ> PointCloudXYZ;
  typedef pcl::PointCloud<pcl::PointXYZRGB> PointCloudRGB;
  typedef pcl::PointCloud<pcl::Normal> PointCloudNormal;

 private:
  Utils() {}
  ~Utils() {}

 public:
  // reference src/Utils.cpp:27-44
  static bool getNormals(PointCloudRGB::Ptr& cloud, double normal_radius, PointCloudNormal::Ptr& normals) {
    return gicpb_shim::getNormals(cloud, normal_radius, normals);
  }
  static bool isValidCloud(PointCloudXYZ::Ptr cloud) { return cloud->size() > 1; }
  static bool isValidCloud(PointCloudRGB::Ptr cloud) { return cloud->size() > 1; }
  static bool isValidTransform(Eigen::Matrix4f transform) { return gicpb_shim::isValidTransform(transform); }
  // reference src/Utils.cpp:145-174
  static double computeCloudResolution(PointCloudRGB::Ptr cloud) { return gicpb_shim::computeCloudResolution(cloud); }
  // reference src/Utils.cpp:215-232: T = [Rx(roll) Ry(pitch) Rz(yaw)] through float quaternions, pcl::transformPointCloud
  static void rotateCloud(PointCloudRGB::Ptr cloud_in, PointCloudRGB::Ptr cloud_out, double roll, double pitch, double yaw) {
    float T[16];
    orc_rotation_rpy(roll, pitch, yaw, T);
    const size_t n = cloud_in->points.size();
    std::vector<float> in(3 * n), out(3 * n);
    for (size_t i = 0; i < n; ++i) {
      in[3 * i] = cloud_in->points[i].x;
      in[3 * i + 1] = cloud_in->points[i].y;
      in[3 * i + 2] = cloud_in->points[i].z;
    }
    orc_transform(T, in.data(), (int)n, out.data());
    if (cloud_out.get() != cloud_in.get()) *cloud_out = *cloud_in;
    for (size_t i = 0; i < n; ++i) {
      cloud_out->points[i].x = out[3 * i];
      cloud_out->points[i].y = out[3 * i + 1];
      cloud_out->points[i].z = out[3 * i + 2];
    }
  }
};

#endif
