// EXPERIMENTAL second build of the MSM unit (round 2 candidate; nothing calls it by default): capi_msm.cu compiled again
// with the row-alternated lazy Fp2 multiply (fp2.cuh B200_FP2_LAZY3: the three wide products and the two Montgomery
// reductions of an Fp2 multiplication run as alternated carry-chain rows).  Motivation: the G2 bucket kernel reaches
// 65 % of the integer multiplier in round 1 (G1: 88 %), its `wait` stall is 2.1 cycles per issue.
// The source is the validated capi_msm.cu, untouched; only names differ: the exported entry points get the prefix
// b200x_lazy3_ (they are NOT part of include/bls12381_b200.h) and the kernels of msm_affine.cuh that have external
// linkage are renamed so the two builds cannot be merged by the linker.  The arithmetic is CPU-validated
// (tests/test_device_source_cpu.py, variant "lazy3"); tools/bench_g2_msm_variant.py times both builds side by side and
// tests/test_gpu_zz_g2_lazy3.py checks that they agree.  Only the G2 entry points are of interest (G1 does not use Fp2).
#define B200_FP2_LAZY3 1
#define k_aff_counts k_aff_counts_lazy3
#define k_aff_level k_aff_level_lazy3
#define k_aff_leftover k_aff_leftover_lazy3
#define k_msm_accumulate_buf k_msm_accumulate_buf_lazy3
#define b200_g1_msm_shard_dev b200x_lazy3_g1_msm_shard_dev
#define b200_g2_msm_shard_dev b200x_lazy3_g2_msm_shard_dev
#define b200_glv_decompose b200x_lazy3_glv_decompose
#define b200_g1_msm_dev b200x_lazy3_g1_msm_dev
#define b200_g2_msm_dev b200x_lazy3_g2_msm_dev
#define b200_g1_msm b200x_lazy3_g1_msm
#define b200_g2_msm b200x_lazy3_g2_msm
#include "capi_msm.cu"
