// stands in for the reference's include/Viewer.h (visualisation only); keeps its global typedef (Viewer.h:27)
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
typedef pcl::PointCloud<pcl::PointXYZRGB> PointCloudRGB;
