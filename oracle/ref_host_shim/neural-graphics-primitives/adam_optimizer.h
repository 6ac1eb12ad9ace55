// Stand-in (test infrastructure): the kernel part of src/testbed_nerf.cu (lines 1-2396) uses nothing from this header.
#pragma once
