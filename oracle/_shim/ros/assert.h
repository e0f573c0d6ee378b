#pragma once
#include <cassert>
#include <cstdlib>
#define ROS_ASSERT(x) assert(x)
#define ROS_BREAK() abort()
