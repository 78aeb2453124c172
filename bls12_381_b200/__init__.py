"""bls12_381_b200 — B200-native hot path of zkcrypto/bls12_381 (batched scalar-mul, G1/G2 MSM,
batched multi_miller_loop + final_exponentiation) behind the C ABI of include/bls12381_b200.h."""
from .engine import B200Error, Engine, MultiEngine  # noqa: F401

__all__ = ["Engine", "MultiEngine", "B200Error"]
