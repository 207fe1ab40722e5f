// stub of <pcl/PointIndices.h>
#pragma once
#include <vector>
namespace pcl {
struct PointIndices {
  std::vector<int> indices;
};
}  // namespace pcl
