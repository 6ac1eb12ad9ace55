// Stand-in (test infrastructure): the kernel part of src/marching_cubes.cu uses no thread pool.
#pragma once
