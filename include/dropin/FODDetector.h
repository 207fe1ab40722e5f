// FODDetector.h - the header swap for the reference's include/FODDetector.h (src/LeicaStateMachine.cpp:200-205,
// test/test_fod_detector.cpp); see dropin/GICPAlignment.h.
#pragma once
#ifndef GICPB_WITH_PCL
#define GICPB_WITH_PCL 1
#endif
#ifndef GICPB_WITH_ROS
#define GICPB_WITH_ROS 1
#endif
#include <Utils.h>

#include "../FODDetector_b200.hpp"
