// GICPAlignment.h - the header swap of INTEGRATION.md: put `include/dropin` in front of the reference's `include/` on the
// include path and every `#include "GICPAlignment.h"` / `#include <GICPAlignment.h>` (reference
// src/LeicaStateMachine.cpp, test/test_gicp_alignment.cpp) picks up the B200 drop-in with PCL and ROS types.
#pragma once
#ifndef GICPB_WITH_PCL
#define GICPB_WITH_PCL 1
#endif
#ifndef GICPB_WITH_ROS
#define GICPB_WITH_ROS 1
#endif
#include <Utils.h>

#include "../GICPAlignment_b200.hpp"
