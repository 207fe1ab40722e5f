// stub of <ros/ros.h>: ros::Time::init() and the logging macros
#pragma once
#include <cstdio>
#include <string>
namespace ros {
struct Time {
  static void init() {}
};
}  // namespace ros
#define ROS_INFO(...) (std::fprintf(stderr, "[ INFO] " __VA_ARGS__), std::fputc('\n', stderr))
#define ROS_WARN(...) (std::fprintf(stderr, "[ WARN] " __VA_ARGS__), std::fputc('\n', stderr))
#define ROS_ERROR(...) (std::fprintf(stderr, "[ERROR] " __VA_ARGS__), std::fputc('\n', stderr))
