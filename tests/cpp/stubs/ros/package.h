// stub of <ros/package.h>: the package path comes from GICPB_STUB_PKG_PATH (the directory that holds test/cube.ply)
#pragma once
#include <cstdlib>
#include <string>
namespace ros {
namespace package {
inline std::string getPath(const std::string&) {
  const char* p = std::getenv("GICPB_STUB_PKG_PATH");
  return p ? p : ".";
}
}  // namespace package
}  // namespace ros
