// CloudIO_b200.hpp - header-only C++ drop-ins for the wire / on-disk formats at the boundary of the registration
// path, on top of the C ABI of libgicp_b200.so (include/gicp_b200.h):
//
//   gicpb_shim::fromROSMsg(msg, cloud)     pcl::fromROSMsg(sensor_msgs::PointCloud2, PointCloud<PointXYZRGB>)
//                                          reference src/node.cpp:37,41 (the source and target clouds the node receives)
//   gicpb_shim::toROSMsg(cloud, msg)       pcl::toROSMsg, reference Utils::cloudToROSMsg src/Utils.cpp:100-105
//   gicpb_shim::loadPCDFile(path, cloud)   pcl::io::loadPCDFile<pcl::PointXYZRGB>, reference
//                                          src/load_and_publish_clouds.cpp:75
//
// `Msg` is any type with the members of sensor_msgs::PointCloud2 / pcl::PCLPointCloud2 (height, width, fields[] with
// name / offset / datatype / count, point_step, row_step, data, is_dense, is_bigendian): the real ROS message under
// -DGICPB_WITH_ROS, the stand-in below otherwise.  The field lookup (pcl::FieldMatches) happens here on the host; the
// gather into 32-byte pcl::PointXYZRGB rows runs on the GPU (gicpb_pointcloud2_to_xyzrgb).
#pragma once
#ifndef CLOUD_IO_B200_HPP_
#define CLOUD_IO_B200_HPP_

#include "GICPAlignment_b200.hpp"

namespace gicpb_shim {

// sensor_msgs::PointField datatypes
enum PointFieldType { kINT8 = 1, kUINT8 = 2, kINT16 = 3, kUINT16 = 4, kINT32 = 5, kUINT32 = 6, kFLOAT32 = 7, kFLOAT64 = 8 };

#ifndef GICPB_WITH_ROS
struct PointField {
  std::string name;
  uint32_t offset = 0;
  uint8_t datatype = 0;
  uint32_t count = 1;
};
struct PointCloud2 {
  uint32_t height = 0, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};
#endif

// pcl::FieldMatches<PointXYZRGB, ...>: same name, datatype and count; "rgb" FLOAT32 and "rgba" UINT32 match each other.
// Returns false (and logs what PCL warns) when x, y or z has no match.
template <class Msg>
inline bool pointcloud2Layout(const Msg& msg, gicpb_pc2_layout* lay) {
  lay->width = msg.width;
  lay->height = msg.height;
  lay->point_step = msg.point_step;
  lay->row_step = msg.row_step;
  lay->off_x = lay->off_y = lay->off_z = lay->off_rgb = -1;
  for (size_t i = 0; i < msg.fields.size(); ++i) {
    const auto& f = msg.fields[i];
    const bool one = f.count == 1 || f.count == 0;
    const bool f32 = f.datatype == kFLOAT32 && one;
    if (f.name == "x" && f32) lay->off_x = (int32_t)f.offset;
    if (f.name == "y" && f32) lay->off_y = (int32_t)f.offset;
    if (f.name == "z" && f32) lay->off_z = (int32_t)f.offset;
    if ((f.name == "rgb" && f32) || (f.name == "rgba" && f.datatype == kUINT32 && one)) lay->off_rgb = (int32_t)f.offset;
  }
  bool ok = true;
  if (lay->off_x < 0) { log(kWarn, "Failed to find match for field 'x'."); ok = false; }
  if (lay->off_y < 0) { log(kWarn, "Failed to find match for field 'y'."); ok = false; }
  if (lay->off_z < 0) { log(kWarn, "Failed to find match for field 'z'."); ok = false; }
  if (lay->off_rgb < 0) log(kWarn, "Failed to find match for field 'rgb'.");
  return ok;
}

template <class Msg, class Cloud>
inline void fromROSMsg(const Msg& msg, Cloud& cloud, Context* shared = nullptr) {
  gicpb_pc2_layout lay;
  if (!pointcloud2Layout(msg, &lay)) throw std::runtime_error("fromROSMsg: the message has no FLOAT32 x / y / z fields");
  std::shared_ptr<Context> hold;
  shared = resolve(shared, hold);
  cloud.points.resize((size_t)msg.width * msg.height);
  cloud.width = msg.width;
  cloud.height = msg.height;
  cloud.is_dense = msg.is_dense;
  if (cloud.points.empty()) return;
  shared->check(gicpb_pointcloud2_to_xyzrgb(shared->get(), msg.data.data(), 0, &lay, &cloud.points[0].x, 0),
                "gicpb_pointcloud2_to_xyzrgb");
}

// the message pcl::toROSMsg builds for PointXYZRGB: fields x@0 y@4 z@8 rgb@16, the 32-byte rows copied as they are
template <class Cloud, class Msg>
inline void toROSMsg(const Cloud& cloud, Msg& msg) {
  msg.height = cloud.height;
  msg.width = cloud.width;
  if (msg.width == 0 && msg.height == 0) {  // pcl::toPCLPointCloud2: an unsized cloud becomes one row
    msg.width = (uint32_t)cloud.points.size();
    msg.height = 1;
  }
  static const char* names[4] = {"x", "y", "z", "rgb"};
  static const uint32_t offs[4] = {0, 4, 8, 16};
  msg.fields.resize(4);
  for (int i = 0; i < 4; ++i) {
    msg.fields[i].name = names[i];
    msg.fields[i].offset = offs[i];
    msg.fields[i].datatype = kFLOAT32;
    msg.fields[i].count = 1;
  }
  msg.is_bigendian = false;
  msg.point_step = (uint32_t)sizeof(cloud.points[0]);
  msg.row_step = msg.point_step * msg.width;
  msg.is_dense = cloud.is_dense;
  msg.data.resize(cloud.points.size() * sizeof(cloud.points[0]));
  if (!cloud.points.empty()) std::memcpy(msg.data.data(), &cloud.points[0], msg.data.size());
}

// returns 0, or -1 like pcl::io::loadPCDFile when the file cannot be read (the reason is logged)
template <class Cloud>
inline int loadPCDFile(const std::string& file_name, Cloud& cloud, Context* shared = nullptr) {
  std::shared_ptr<Context> hold;
  shared = resolve(shared, hold);
  gicpb_pcd_info info;
  if (gicpb_pcd_load_xyzrgb(shared->get(), file_name.c_str(), nullptr, 0, 0, &info) != GICPB_OK) {
    log(kError, "[pcl::PCDReader::read] %s", gicpb_last_error(shared->get()));
    return -1;
  }
  cloud.points.resize((size_t)info.points);
  gicpb_pcd_info body;
  void* rows = cloud.points.empty() ? static_cast<void*>(&body) : static_cast<void*>(&cloud.points[0].x);
  if (gicpb_pcd_load_xyzrgb(shared->get(), file_name.c_str(), rows, info.points, 0, &body) != GICPB_OK) {
    log(kError, "[pcl::PCDReader::read] %s", gicpb_last_error(shared->get()));
    return -1;
  }
  cloud.width = (uint32_t)body.width;
  cloud.height = (uint32_t)body.height;
  cloud.is_dense = body.is_dense != 0;
  return 0;
}

}  // namespace gicpb_shim
#endif  // CLOUD_IO_B200_HPP_
