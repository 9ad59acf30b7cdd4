// scripts/fp64_probe.cu -- what bounds the dense-gate passes on B200: the FP64 vector pipe (DFMA), and whether the FP64
// tensor path (DMMA, mma.sync.m8n8k4.f64) is a second, concurrently usable resource or the same one.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/fp64_probe scripts/fp64_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// mode 0: DFMA only (8 chains); 1: DMMA only (4 accumulator pairs); 2: both interleaved in every warp;
// 3: even warps DFMA, odd warps DMMA
__global__ void __launch_bounds__(1024) probe(int mode, int iters, double* out, long long* cycles) {
    const long long c_begin = clock64();
    double x[8], c0[4], c1[4];
    const double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
    for (int i = 0; i < 8; i++) x[i] = threadIdx.x + i;
    for (int i = 0; i < 4; i++) { c0[i] = i; c1[i] = -i; }
    const int warp = threadIdx.x >> 5;
    const bool do_fma = mode == 0 || mode == 2 || (mode == 3 && !(warp & 1));
    const bool do_mma = mode == 1 || mode == 2 || (mode == 3 && (warp & 1));
    for (int it = 0; it < iters; it++) {
        if (do_fma) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = fma(x[i], a, b);
        }
        if (do_mma) {
#pragma unroll
            for (int i = 0; i < 4; i++) dmma(c0[i], c1[i], a, b);
        }
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += x[i];
    for (int i = 0; i < 4; i++) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0 && cycles) cycles[blockIdx.x] = clock64() - c_begin;
}

int main() {
    double* d;
    cudaMalloc(&d, 8);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d_cyc;
    cudaMalloc(&d_cyc, sizeof(long long) * sms * 8);
    const int iters = 20000, blocks = sms * 8;
    // (a) clock-independent: ONE block of 8 warps per SM, cycles from clock64 -> FMA per clock per SM
    for (int mode = 0; mode < 2; mode++) {
        for (int warps = 4; warps <= 32; warps *= 2) {
            probe<<<sms, warps * 32>>>(mode, 2000, d, d_cyc);
            cudaDeviceSynchronize();
            long long c = 0;
            cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost);
            const double fma_per_thread = mode == 0 ? 8.0 * 2000 : 0.0, mma_per_warp = mode == 1 ? 4.0 * 2000 * 256.0 : 0.0;
            const double fmas = fma_per_thread * warps * 32 + mma_per_warp * warps;
            printf("%-10s %2d warps on one SM: %9lld cycles, %6.1f FP64 FMA per clock per SM\n", mode == 0 ? "DFMA" : "DMMA", warps, c, fmas / (double)c);
        }
    }
    const char* names[4] = {"DFMA only", "DMMA only", "DFMA + DMMA interleaved per warp", "DFMA warps beside DMMA warps"};
    for (int mode = 0; mode < 4; mode++) {
        probe<<<blocks, 256>>>(mode, 100, d, nullptr);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<<<blocks, 256>>>(mode, iters, d, nullptr);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double warps = (double)blocks * 8, thr = warps * 32;
        double fma_flops = 0, mma_flops = 0;
        if (mode == 0 || mode == 2) fma_flops = thr * 8.0 * iters * 2;
        if (mode == 3) fma_flops = thr / 2 * 8.0 * iters * 2;
        if (mode == 1 || mode == 2) mma_flops = warps * 4.0 * iters * (2.0 * 8 * 8 * 4);
        if (mode == 3) mma_flops = warps / 2 * 4.0 * iters * (2.0 * 8 * 8 * 4);
        printf("%-36s %8.3f ms  DFMA %6.2f TFLOP/s  DMMA %6.2f TFLOP/s  total %6.2f TFLOP/s\n", names[mode], ms, fma_flops / ms / 1e9,
               mma_flops / ms / 1e9, (fma_flops + mma_flops) / ms / 1e9);
    }
    printf("error state: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
