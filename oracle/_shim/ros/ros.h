// stub (oracle/_ref build): the factor sources include <ros/ros.h> through d2vins_params.hpp but use nothing of it
#pragma once
#include <cstdio>
#include <string>
#include "assert.h"
namespace ros {
struct Time { double t = 0; Time() {} explicit Time(double s) : t(s) {} double toSec() const { return t; } static Time now() { return Time(); } };
struct Duration { double t = 0; Duration() {} explicit Duration(double s) : t(s) {} double toSec() const { return t; } };
class NodeHandle {};
}  // namespace ros
#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)
#define ROS_DEBUG(...) do { } while (0)
