// stands in for the reference's include/CADToPointCloud.h + src/CADToPointCloud.cpp (VTK mesh loading and area-weighted
// sampling with libc rand()): reads a binary little-endian .ply triangle mesh and samples it with the CPU oracle's
// restatement of the reference sampler (oracle/gicp_oracle.cpp orc_sample_mesh), continuing the process-wide rand()
// stream from call to call as the reference does (it never seeds).
#pragma once
#ifndef _CAD_TO_POINTCLOUD_H
#define _CAD_TO_POINTCLOUD_H

#include <Utils.h>

#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

extern "C" void orc_sample_mesh(const float* verts, int nv, const int* faces, int nf, int n_samples, long skip_draws,
                                float* out_xyz);

class CADToPointCloud {
  typedef pcl::PointCloud<pcl::PointXYZ> PointCloudXYZ;
  typedef pcl::PointCloud<pcl::PointXYZRGB> PointCloudRGB;

 public:
  CADToPointCloud(const std::string& cad_file_path, int sample_points) : sample_points_(sample_points) {
    std::ifstream f(cad_file_path.c_str(), std::ios::binary);
    if (!f) throw std::runtime_error("cannot open " + cad_file_path);
    std::string line;
    int nv = 0, nf = 0;
    bool binary = false;
    while (std::getline(f, line)) {
      if (line.compare(0, 14, "element vertex") == 0) nv = std::atoi(line.c_str() + 15);
      if (line.compare(0, 12, "element face") == 0) nf = std::atoi(line.c_str() + 13);
      if (line.compare(0, 27, "format binary_little_endian") == 0) binary = true;
      if (line.compare(0, 10, "end_header") == 0) break;
    }
    if (!binary || nv <= 0 || nf <= 0) throw std::runtime_error("unsupported .ply: " + cad_file_path);
    verts_.resize(3 * (size_t)nv);
    f.read(reinterpret_cast<char*>(verts_.data()), (std::streamsize)(verts_.size() * sizeof(float)));
    for (int i = 0; i < nf; ++i) {
      unsigned char k = 0;
      f.read(reinterpret_cast<char*>(&k), 1);
      int idx[8] = {0};
      if (k < 3 || k > 8) throw std::runtime_error("unsupported face in " + cad_file_path);
      f.read(reinterpret_cast<char*>(idx), 4 * k);
      for (int t = 1; t + 1 < k; ++t) {  // fan
        faces_.push_back(idx[0]);
        faces_.push_back(idx[t]);
        faces_.push_back(idx[t + 1]);
      }
    }
    if (!f) throw std::runtime_error("truncated .ply: " + cad_file_path);
  }
  ~CADToPointCloud() {}

  void convertCloud(PointCloudRGB::Ptr cloud) {
    static long draws_so_far = 0;  // three rand() draws per sample, one stream per process
    std::vector<float> xyz(3 * (size_t)sample_points_);
    orc_sample_mesh(verts_.data(), (int)(verts_.size() / 3), faces_.data(), (int)(faces_.size() / 3), sample_points_,
                    draws_so_far, xyz.data());
    draws_so_far += 3L * sample_points_;
    cloud->points.resize((size_t)sample_points_);
    for (int i = 0; i < sample_points_; ++i) {
      pcl::PointXYZRGB p;
      p.x = xyz[3 * i];
      p.y = xyz[3 * i + 1];
      p.z = xyz[3 * i + 2];
      p.r = p.g = p.b = 255;
      cloud->points[(size_t)i] = p;
    }
    cloud->width = (uint32_t)sample_points_;
    cloud->height = 1;
    cloud->is_dense = true;
  }

 private:
  int sample_points_;
  std::vector<float> verts_;
  std::vector<int> faces_;
};

#endif
