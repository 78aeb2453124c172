// EXPERIMENTAL pairing kernels (round 2 candidate, tuning key pairing_variant = 6; default stays v4): as pairing_v5.cu
// but with ALL THREE Karatsuba products of an Fp2 multiplication in one row-alternating routine (fp2.cuh
// B200_FP2_KTRIPLE) — three dependent streams per warp.  CPU-validated (variant "ktriple"); not yet measured.
#define B200_PAIR_VARIANT v6
#define B200_PAIR_MINB 4
#define B200_FP2_KTRIPLE 1
#include "pairing_kernels.inc"
