// stub of <pcl/point_types.h>: the point types the reference's tests and the drop-in headers use, PCL's memory layout
#pragma once
#include <cstdint>
namespace pcl {
struct alignas(16) PointXYZ {
  float x = 0.f, y = 0.f, z = 0.f, data_w = 1.f;
};
struct alignas(16) PointXYZRGB {
  float x = 0.f, y = 0.f, z = 0.f, data_w = 1.f;
  union {
    float rgb;
    uint32_t rgba;
    struct { uint8_t b, g, r, a; };
  };
  float pad_[3] = {0.f, 0.f, 0.f};
  PointXYZRGB() : rgba(0xff000000u) {}  // r = g = b = 0, a = 255
};
static_assert(sizeof(PointXYZRGB) == 32, "pcl::PointXYZRGB is 32 bytes");
struct alignas(16) Normal {
  float normal_x = 0.f, normal_y = 0.f, normal_z = 0.f, data_n_w = 0.f;
  float curvature = 0.f, pad_[3] = {0.f, 0.f, 0.f};
};
static_assert(sizeof(Normal) == 32, "pcl::Normal is 32 bytes");
}  // namespace pcl
