#pragma once
#include "Pose.h"
namespace Swarm { class Odometry { public: Odometry() {} }; }
