// pairing kernels: Fp2 multiply = Karatsuba over three fp_mul_c calls (measured best for these latency-bound kernels;
// also tried and measured slower: inlined products, lazy reduction, 168/128-register builds, called Fp2 add/sub)
#define B200_PAIR_VARIANT v4
#define B200_PAIR_MINB 4
#define B200_FP2_KCALL 1
#include "pairing_kernels.inc"
