// Stand-in (test infrastructure): the reference's common.h only includes the logger, the device headers never call it.
#pragma once
