#pragma once
#include <pcl/point_cloud.h>
#include <sensor_msgs/PointCloud2.h>
namespace pcl {
template <typename T> void toROSMsg(const PointCloud<T> &, sensor_msgs::PointCloud2 &) {}
template <typename T> void fromROSMsg(const sensor_msgs::PointCloud2 &, PointCloud<T> &) {}
}
