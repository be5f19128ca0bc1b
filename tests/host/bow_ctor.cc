// tests/host/bow_ctor.cc -- TEST INFRASTRUCTURE: ORBmatcher's constructor and constants for bow_cpu_mine (in a deployment they come from
// host/ORBmatcher_b200.cc, which also holds the device-frame searches and so cannot be linked without the library).
#include "ORBmatcher.h"
namespace ORB_SLAM3 {
const int ORBmatcher::TH_HIGH = 100;
const int ORBmatcher::TH_LOW = 50;
const int ORBmatcher::HISTO_LENGTH = 30;
ORBmatcher::ORBmatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}
}
