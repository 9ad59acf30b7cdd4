// qv_tile_common.cuh -- device helpers shared by the interpreter tile kernel (qv_tile_kernel.cuh) and the kernels the
// pass compiler generates (qv_jit_kernel.cuh): streaming / asynchronous accesses, peer pointer table, remap index.
#pragma once
#if !defined(__CUDACC_RTC__)
#include <cuda_runtime.h>
#endif

#include "qv_ops.h"

// Streaming 128-bit accesses that do not allocate in L1: L1 is kept for the tile
// program (ops, matrices, diagonal tables), which every CTA re-reads.
__device__ __forceinline__ qvc qv_ld_stream(const qvc* p) {
    qvc v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ qvc qv_scaled(qvc v, bool on, double f) {
    if (on) {
        v.x *= f;
        v.y *= f;
    }
    return v;
}
__device__ __forceinline__ void qv_st_stream(qvc* p, qvc v) {
    asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
}

// Asynchronous 16-byte global -> shared copy (LDGSTS): no register staging, so a thread keeps all 16 of
// its tile loads in flight at once (64 KiB per CTA) and the swizzled shared-memory slot is free to choose.
__device__ __forceinline__ void qv_cp_async16(qvc* smem_dst, const qvc* gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

// The same copy with a run-time source size of 16 or 0 bytes: 0 reads nothing and fills the 16 bytes with zeros (amplitudes of
// a shard that is known to be all zeros are not fetched over NVLink).
__device__ __forceinline__ void qv_cp_async16_z(qvc* smem_dst, const qvc* gsrc, uint32_t src_bytes) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}

struct QvProgSmall { uint8_t bytes[QV_PROG_SMALL_BYTES]; };
struct QvProgLarge { uint8_t bytes[QV_PROG_LARGE_BYTES]; };

QV_HD uint64_t qv_remap_index(uint64_t P, const QvRemap& rm) {
    uint64_t S = P;
    for (uint32_t i = 0; i < rm.n_pairs; i++) {
        const uint64_t x = ((P >> rm.local_bit[i]) ^ (P >> rm.global_bit[i])) & 1ull;
        S ^= (x << rm.local_bit[i]) | (x << rm.global_bit[i]);
    }
    return S;
}

