// pairing kernels compiled for 4 resident 64-thread blocks per SM
#define B200_PAIR_VARIANT v4
#define B200_PAIR_MINB 4
#include "pairing_kernels.inc"
