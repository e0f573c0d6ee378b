#pragma once
#include <ros/ros.h>
namespace sensor_msgs {
struct Imu { struct { ros::Time stamp; } header; struct { double x = 0, y = 0, z = 0; } angular_velocity, linear_acceleration; };
}
