// cloud_io.hpp - PointCloud2 unpacking and PCD reading (cloud_io.cu); SURVEY section 8f row 4.
#pragma once
#include "engine.hpp"

namespace gicpb {

// out[2i] = (x, y, z, 1), out[2i+1] = (rgba bits, 0, 0, 0): pcl::PointXYZRGB rows from a PointCloud2 payload.
// Point i sits at data + (i / width) * row_step + (i % width) * point_step; off_rgb < 0: no colour field (rgba = a 255).
void launch_pc2_unpack(const unsigned char* data, int64_t n, int64_t width, int64_t point_step, int64_t row_step, int off_x,
                       int off_y, int off_z, int off_rgb, float4* out, cudaStream_t stream);

struct PcdField {
  std::string name;
  int offset, size;
  char type;  // 'I', 'U', 'F'
  int count;
};

// PCL 1.8.1 PCDReader restated: header -> field table, body -> point-major blob (PCLPointCloud2::data)
struct PcdFile {
  std::vector<PcdField> fields;
  int64_t width = 0, height = 0, points = 0, data_offset = 0;
  int point_step = 0;
  int data_kind = -1;  // 0 ascii, 1 binary, 2 binary_compressed
  int off_x = -1, off_y = -1, off_z = -1, off_rgb = -1;  // fields a PointXYZRGB maps (pcl::FieldMatches), -1 = absent
  bool is_dense = true;
  void read_header(const std::string& path);
  void read_body(const std::string& path, unsigned char* blob);  // blob: points * point_step bytes
};

}  // namespace gicpb
