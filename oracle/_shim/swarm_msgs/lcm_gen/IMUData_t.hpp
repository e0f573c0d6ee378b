#pragma once
#include <cstdint>
struct Time_t { int32_t sec = 0, nsec = 0; };
struct Vector3d_t { double x = 0, y = 0, z = 0; };
struct IMUData_t { Time_t timestamp; double dt = 0; Vector3d_t acc, gyro; };
