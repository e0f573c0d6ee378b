#pragma once
#include <ros/ros.h>
#include "lcm_gen/IMUData_t.hpp"
inline ros::Time toROSTime(const Time_t &t) { return ros::Time(t.sec + 1e-9 * t.nsec); }
inline Time_t toLCMTime(const ros::Time &t) { Time_t o; o.sec = (int32_t)t.toSec(); o.nsec = (int32_t)((t.toSec() - o.sec) * 1e9); return o; }
