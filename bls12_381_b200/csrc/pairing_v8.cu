// pairing kernels compiled for 8 resident 64-thread blocks per SM
#define B200_PAIR_VARIANT v8
#define B200_PAIR_MINB 8
#include "pairing_kernels.inc"
