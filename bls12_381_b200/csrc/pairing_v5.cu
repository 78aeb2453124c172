// EXPERIMENTAL pairing kernels (round 2 candidate, selected with the tuning key pairing_variant = 5; default stays v4):
// the same kernels as pairing_v4.cu with the dual-stream Fp2 multiply (fp2.cuh B200_FP2_KDUAL: the two independent Fp
// products of an Fp2 multiplication / squaring run as row-alternated carry chains in one routine).  Motivation: ncu of
// v4 shows the `wait` stall (fixed-latency dependency) at 3.3-3.6 cycles per issue with 2 warps per scheduler.
// CPU-validated (tests/test_device_source_cpu.py, variant "kdual"); not yet measured on hardware.
#define B200_PAIR_VARIANT v5
#define B200_PAIR_MINB 4
#define B200_FP2_KDUAL 1
#include "pairing_kernels.inc"
