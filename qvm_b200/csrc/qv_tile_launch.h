// qv_tile_launch.h -- launchers of the tile-kernel instantiations (one translation unit per (mode, register bits)).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

struct QvPeers;
struct qvc;

struct QvTileLaunch {
    const uint8_t* blob;        // control program (copied into the kernel's __grid_constant__ parameter)
    size_t blob_bytes;
    bool full;                  // T == 12
    int grid;
    size_t smem;                // dynamic shared memory: the tile
    cudaStream_t stream;
    const QvPeers* peers;
    const qvc* tables;
    qvc* alt_own;
};

// mode 0 = local, 1 = peer (in place over NVLink), 2 = pull (remap fused into the loads); return nullptr or an error string
const char* qv_launch_tile_0_3(const QvTileLaunch& L);
const char* qv_launch_tile_1_3(const QvTileLaunch& L);
const char* qv_launch_tile_2_3(const QvTileLaunch& L);
const char* qv_launch_tile_0_4(const QvTileLaunch& L);
