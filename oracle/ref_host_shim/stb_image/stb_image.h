// Stand-in (test infrastructure): nothing of this header is used by the kernels compiled for the oracle pins.
#pragma once
