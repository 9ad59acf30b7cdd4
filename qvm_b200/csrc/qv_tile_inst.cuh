// qv_tile_inst.cuh -- body of one tile-kernel translation unit: define QV_INST_MODE and QV_INST_M, then include.
#include <atomic>
#include <cstring>

#include "qv_tile_kernel.cuh"
#include "qv_tile_launch.h"

namespace {

template <typename PROG, bool FULL>
const char* qv_launch_one(const QvTileLaunch& L) {
    constexpr int MODE = QV_INST_MODE, M = QV_INST_M;
    // function attributes are per device: one flag per (instantiation, device ordinal)
    static std::atomic<bool> attr_set[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return "cannot identify the current device";
    if (!attr_set[dev].exchange(true)) {
        cudaError_t e = cudaFuncSetAttribute(qv_tile_kernel<PROG, MODE, FULL, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(qv_tile_kernel<PROG, MODE, FULL, M>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
        if (e != cudaSuccess) {
            attr_set[dev] = false;
            return cudaGetErrorString(e);
        }
    }
    static thread_local PROG prog;   // up to 28 KiB: keep it off the stack
    std::memcpy(prog.bytes, L.blob, L.blob_bytes);
    const int threads = M == 4 ? QV_THREADS_WIDE : QV_THREADS;
    qv_tile_kernel<PROG, MODE, FULL, M><<<L.grid, threads, L.smem, L.stream>>>(prog, *L.peers, L.tables, L.alt_own);
    const cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

}  // namespace

#define QV_CAT2(a, b, c) qv_launch_tile_##a##_##b
#define QV_CAT(a, b) QV_CAT2(a, b, 0)

const char* QV_CAT(QV_INST_MODE, QV_INST_M)(const QvTileLaunch& L) {
    const bool small = L.blob_bytes <= QV_PROG_SMALL_BYTES;
#if QV_INST_M == 4
    if (!L.full) return "4 register bits need a full 12-bit tile";
    return small ? qv_launch_one<QvProgSmall, true>(L) : qv_launch_one<QvProgLarge, true>(L);
#else
    if (L.full) return small ? qv_launch_one<QvProgSmall, true>(L) : qv_launch_one<QvProgLarge, true>(L);
    return small ? qv_launch_one<QvProgSmall, false>(L) : qv_launch_one<QvProgLarge, false>(L);
#endif
}
