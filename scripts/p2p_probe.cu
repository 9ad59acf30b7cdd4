// p2p_probe.cu -- NVLink peer-access microbenchmark (one process, two GPUs).  Build on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/p2p_probe scripts/p2p_probe.cu
// Measures, for a buffer of N bytes on the PEER device, kernels running on device 0 that
//   (a) read peer memory with 128-bit LDG and write local memory,
//   (b) read peer memory with cp.async into shared memory and write local memory,
//   (c) read local memory and write peer memory,
//   (d) exchange: read peer + write peer (what a remap pass does), register-staged and cp.async staged,
// each on 1 GPU alone and on both GPUs at once (full duplex).
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("%s failed: %s\n", #x, cudaGetErrorString(e));                      \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

struct alignas(16) v16 {
    double x, y;
};

__global__ void __launch_bounds__(256) k_copy_ldg(const v16* __restrict__ src, v16* __restrict__ dst, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x * 8;
    for (size_t i = (size_t)blockIdx.x * blockDim.x * 8 + threadIdx.x; i < n; i += stride) {
        v16 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = src[i + (size_t)j * 256];
#pragma unroll
        for (int j = 0; j < 8; j++) dst[i + (size_t)j * 256] = v[j];
    }
}

// tile-like: 4096 elements staged in smem by cp.async, then stored
__global__ void __launch_bounds__(256, 3) k_copy_cpasync(const v16* __restrict__ src, v16* __restrict__ dst, size_t n) {
    extern __shared__ __align__(16) unsigned char raw[];
    v16* tile = (v16*)raw;
    const size_t n_tiles = n / 4096;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t base = t * 4096;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            unsigned d = (unsigned)__cvta_generic_to_shared(tile + threadIdx.x + i * 256);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src + base + threadIdx.x + i * 256) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 16; i++) dst[base + threadIdx.x + i * 256] = tile[threadIdx.x + i * 256];
        __syncthreads();
    }
}

// in-place exchange between a local and a peer region (what a 1-bit remap does): a <-> b
__global__ void __launch_bounds__(256, 3) k_swap_cpasync(v16* a, v16* b, size_t n) {
    extern __shared__ __align__(16) unsigned char raw[];
    v16* tile = (v16*)raw;   // 2048 from a, 2048 from b
    const size_t n_tiles = n / 2048;
    for (size_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const size_t base = t * 2048;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            unsigned d0 = (unsigned)__cvta_generic_to_shared(tile + threadIdx.x + i * 256);
            unsigned d1 = (unsigned)__cvta_generic_to_shared(tile + 2048 + threadIdx.x + i * 256);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d0), "l"(a + base + threadIdx.x + i * 256) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d1), "l"(b + base + threadIdx.x + i * 256) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; i++) {
            b[base + threadIdx.x + i * 256] = tile[threadIdx.x + i * 256];
            a[base + threadIdx.x + i * 256] = tile[2048 + threadIdx.x + i * 256];
        }
        __syncthreads();
    }
}

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) {
        printf("need 2 GPUs\n");
        return 0;
    }
    const size_t bytes = (argc > 1 ? atoll(argv[1]) : 4096ll) << 20;   // MiB
    const size_t n = bytes / 16;
    v16* buf[2][2];
    cudaStream_t st[2];
    cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        for (int k = 0; k < 2; k++) {
            CK(cudaMalloc(&buf[d][k], bytes));
            CK(cudaMemset(buf[d][k], 0, bytes));
        }
        CK(cudaStreamCreate(&st[d]));
        CK(cudaEventCreate(&e0[d]));
        CK(cudaEventCreate(&e1[d]));
        CK(cudaFuncSetAttribute(k_copy_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        CK(cudaFuncSetAttribute(k_swap_cpasync, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    }
    const int grid = 148 * 3 * 4;
    auto run = [&](const char* name, int which, bool both, double bytes_per_dir) {
        for (int rep = 0; rep < 3; rep++) {
            for (int d = 0; d < (both ? 2 : 1); d++) {
                CK(cudaSetDevice(d));
                v16* local0 = buf[d][0];
                v16* local1 = buf[d][1];
                v16* peer0 = buf[1 - d][0];
                v16* peer1 = buf[1 - d][1];
                CK(cudaEventRecord(e0[d], st[d]));
                switch (which) {
                    case 0: k_copy_ldg<<<grid, 256, 0, st[d]>>>(local0, local1, n); break;                 // local copy
                    case 1: k_copy_ldg<<<grid, 256, 0, st[d]>>>(peer0, local1, n); break;                  // peer read (LDG)
                    case 2: k_copy_cpasync<<<grid, 256, 65536, st[d]>>>(peer0, local1, n); break;          // peer read (cp.async)
                    case 3: k_copy_ldg<<<grid, 256, 0, st[d]>>>(local0, peer1, n); break;                  // peer write
                    case 4: k_copy_cpasync<<<grid, 256, 65536, st[d]>>>(local0, peer1, n); break;          // peer write from smem
                    case 5: k_swap_cpasync<<<grid, 256, 65536, st[d]>>>(local0 + (d ? n / 2 : 0), peer0 + (d ? n / 2 : 0), n / 2); break;   // exchange halves
                    case 6: k_copy_cpasync<<<grid, 256, 65536, st[d]>>>(local0, local1, n); break;         // local copy cp.async
                }
                CK(cudaGetLastError());
                CK(cudaEventRecord(e1[d], st[d]));
            }
            for (int d = 0; d < (both ? 2 : 1); d++) {
                CK(cudaSetDevice(d));
                CK(cudaStreamSynchronize(st[d]));
            }
        }
        float ms = 0;
        CK(cudaSetDevice(0));
        CK(cudaEventElapsedTime(&ms, e0[0], e1[0]));
        printf("%-44s %s  %8.3f ms  %8.1f GB/s per direction per GPU\n", name, both ? "both GPUs" : "GPU0 only", ms, bytes_per_dir / ms / 1e6);
    };
    printf("buffer %zu MiB\n", bytes >> 20);
    run("local copy (LDG/STG), bytes read", 0, false, (double)bytes);
    run("local copy (cp.async tile), bytes read", 6, false, (double)bytes);
    for (int both = 0; both < 2; both++) {
        run("peer read  LDG.128 -> local store", 1, both, (double)bytes);
        run("peer read  cp.async -> smem -> local store", 2, both, (double)bytes);
        run("local read -> peer write (STG)", 3, both, (double)bytes);
        run("local read cp.async -> smem -> peer write", 4, both, (double)bytes);
        run("exchange halves in place (cp.async both, STG both)", 5, both, (double)bytes / 2);
    }
    return 0;
}
