// pairing kernels, experimental variant: Fp2 multiply with the Fp products inlined side by side
#define B200_PAIR_VARIANT v8
#define B200_PAIR_MINB 4
#define B200_FP2_KINLINE 1
#include "pairing_kernels.inc"
