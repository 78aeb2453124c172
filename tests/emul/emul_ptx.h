// TEST INFRASTRUCTURE.  Bit-exact C models of the PTX extended-precision instructions used by fp.cuh / fr.cuh.
// One thread-local condition-code register (CC.CF) per host thread, like one per GPU thread.
#pragma once
#include <cstdint>
// (included from inside namespace b200)
inline thread_local uint32_t emul_cf = 0;   // one CC.CF per host thread, shared by all translation units
#define EMUL_DEV static inline
EMUL_DEV void ptx_add_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; d = (uint32_t)t; emul_cf = (uint32_t)(t >> 32); }
EMUL_DEV void ptx_addc_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + emul_cf; d = (uint32_t)t; emul_cf = (uint32_t)(t >> 32); }
EMUL_DEV void ptx_addc(uint32_t &d, uint32_t a, uint32_t b) { d = a + b + emul_cf; }
EMUL_DEV void ptx_sub_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; d = (uint32_t)t; emul_cf = (uint32_t)(t >> 63); }
EMUL_DEV void ptx_subc_cc(uint32_t &d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - emul_cf; d = (uint32_t)t; emul_cf = (uint32_t)(t >> 63); }
EMUL_DEV void ptx_subc(uint32_t &d, uint32_t a, uint32_t b) { d = a - b - emul_cf; }
EMUL_DEV void ptx_mul_lo(uint32_t &d, uint32_t a, uint32_t b) { d = a * b; }
EMUL_DEV void ptx_mul_hi(uint32_t &d, uint32_t a, uint32_t b) { d = (uint32_t)(((uint64_t)a * b) >> 32); }
EMUL_DEV void ptx_mad_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; d = (uint32_t)t; emul_cf = (uint32_t)(t >> 32); }
EMUL_DEV void ptx_madc_lo_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + emul_cf; d = (uint32_t)t; emul_cf = (uint32_t)(t >> 32); }
EMUL_DEV void ptx_madc_hi_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + emul_cf; d = (uint32_t)t; emul_cf = (uint32_t)(t >> 32); }
EMUL_DEV void ptx_madc_hi(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { d = (uint32_t)(((uint64_t)a * b) >> 32) + c + emul_cf; }
EMUL_DEV void ptx_mad_hi_cc(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c; d = (uint32_t)t; emul_cf = (uint32_t)(t >> 32); }
EMUL_DEV void ptx_madc_lo(uint32_t &d, uint32_t a, uint32_t b, uint32_t c) { d = a * b + c + emul_cf; }

