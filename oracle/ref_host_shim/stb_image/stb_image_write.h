// Stand-in (test infrastructure): the kernel part of src/marching_cubes.cu writes no images.
#pragma once
