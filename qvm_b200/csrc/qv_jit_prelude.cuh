// qv_jit_prelude.cuh -- first include of every generated pass (qv_jit_gen.cpp): the micro-op templates and the
// handful of macros that let the same text compile for the device (NVRTC) and, under QVJ_HOST, for the test emulator.
#pragma once
#define QV_NATURAL_DENSE 1
#if defined(QVJ_HOST)
#include "qv_ops.h"
#define QVJ_FN static inline
#define QVJ_RESTRICT
#define QVJ_UNROLL
#define QVJ_NOUNROLL
#define QVJ_FENCE() do { } while (0)
#define QVJ_SYNC() do { } while (0)
#else
#include "qv_tile_common.cuh"
#define QVJ_FN __device__ __forceinline__
#define QVJ_RESTRICT __restrict__
#define QVJ_UNROLL _Pragma("unroll")
#define QVJ_NOUNROLL _Pragma("unroll 1")
#define QVJ_FENCE() asm volatile("" ::: "memory")
#define QVJ_SYNC() __syncthreads()
#endif
// the pass's shared-memory swizzle (see wants_wide_swizzle in qv_jit_gen.cpp); both are XOR-linear
#if QVJ_WIDE_SWZ
QVJ_FN uint32_t qvj_swz(uint32_t e) { return e ^ ((e >> 3) & 7u) ^ ((e >> 6) & 7u) ^ ((e >> 9) & 7u); }
// slot in this pass's layout of a slot given in the default layout (host-precomputed store-permutation constants)
QVJ_FN uint32_t qvj_from_swz1(uint32_t s) { return qvj_swz(qv_swz(s)); }
#else
QVJ_FN uint32_t qvj_swz(uint32_t e) { return qv_swz(e); }
QVJ_FN uint32_t qvj_from_swz1(uint32_t s) { return s; }
#endif
// a 32-bit field of the control program at a fixed offset (constant bank on the device)
QVJ_FN uint32_t qvj_u32(const uint8_t* blob, uint32_t off) { return *reinterpret_cast<const uint32_t*>(blob + off); }
