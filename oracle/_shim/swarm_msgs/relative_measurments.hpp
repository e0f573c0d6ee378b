// Stand-in for swarm_msgs/relative_measurments.hpp (un-vendored) -- Swarm::LoopEdge lives in the Pose.h stand-in.
#pragma once
#include <swarm_msgs/Pose.h>
