// see sophus/se3.hpp in refshim
#pragma once
#include "sophus/se3.hpp"
