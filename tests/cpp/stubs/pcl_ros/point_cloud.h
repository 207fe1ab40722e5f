// stub of <pcl_ros/point_cloud.h>
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
