// oracle/ref_shim/dbow/boost/serialization/map.hpp -- TEST INFRASTRUCTURE (see serialization.hpp).
#pragma once
#include "serialization.hpp"
