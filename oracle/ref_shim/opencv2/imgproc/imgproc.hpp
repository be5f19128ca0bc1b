// TEST INFRASTRUCTURE: see core/core.hpp in this directory.
#pragma once
#include "opencv2/core/core.hpp"
