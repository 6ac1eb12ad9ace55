// Stand-in (test infrastructure): the reference's common.h only includes the logger, the device headers never call it.  The real
// header drags in the standard headers common.h relies on.
#pragma once
#include <algorithm>
#include <cassert>
#include <sstream>
#include <string>
#include <vector>
