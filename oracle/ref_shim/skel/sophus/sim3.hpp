// oracle/ref_shim/skel/sophus/sim3.hpp -- TEST INFRASTRUCTURE: resolved by the reference's include/ORBmatcher.h; everything is in ref_frame_skel.h
#pragma once
#include "../ref_frame_skel.h"
