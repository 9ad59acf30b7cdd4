// qv_kernels.cuh -- the streaming sm_100a kernels of libqvmcuda (reductions, element-wise passes, sampler,
// generic dense gates, pull remap).  The tile kernel lives in qv_tile_kernel.cuh and is instantiated in its own
// translation units; this header is included by qvmcuda.cu only.
#pragma once
#include "qv_tile_kernel.cuh"

// ---------------------------------------------------------------------------
// Generic dense k-qubit gate (k >= 3 mixing qubits; rare on the benchmark path):
// a CTA stages G groups of 2^k amplitudes (G*2^k = 2048) in shared memory and
// every thread produces output amplitudes as a row-times-column sum in the
// reference's left-to-right order (src/linear-algebra.lisp:97-131).
// ---------------------------------------------------------------------------
#define QV_BIG_ELEMS 2048
__global__ void __launch_bounds__(QV_THREADS)
qv_big_kernel(qvc* __restrict__ psi, QvBigGate g, const qvc* __restrict__ mat, uint32_t n_bits) {
    __shared__ qvc in[QV_BIG_ELEMS];
    const uint32_t k = g.k;
    const uint32_t d = 1u << k;
    const uint32_t G = QV_BIG_ELEMS >> k;
    const uint64_t n_groups = 1ull << (n_bits - k);
    // sorted target positions for zero insertion
    uint32_t sorted[16];
    for (uint32_t j = 0; j < k; j++) sorted[j] = g.pos[j];
    for (uint32_t i = 1; i < k; i++) {
        uint32_t v = sorted[i];
        int j = (int)i - 1;
        while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; j--; }
        sorted[j + 1] = v;
    }
    for (uint64_t gb = (uint64_t)blockIdx.x * G; gb < n_groups; gb += (uint64_t)gridDim.x * G) {
        for (uint32_t idx = threadIdx.x; idx < QV_BIG_ELEMS; idx += QV_THREADS) {
            const uint64_t grp = gb + (idx >> k);
            const uint32_t c = idx & (d - 1);
            if (grp < n_groups) {
                uint64_t base = grp;
                for (uint32_t j = 0; j < k; j++) {
                    const uint64_t lo = base & ((1ull << sorted[j]) - 1ull);
                    base = ((base >> sorted[j]) << (sorted[j] + 1)) | lo;
                }
                uint64_t a = base;
                for (uint32_t j = 0; j < k; j++)
                    if (c >> j & 1) a |= 1ull << g.pos[j];
                in[idx] = psi[a];
            }
        }
        __syncthreads();
        for (uint32_t idx = threadIdx.x; idx < QV_BIG_ELEMS; idx += QV_THREADS) {
            const uint64_t grp = gb + (idx >> k);
            const uint32_t r = idx & (d - 1);
            if (grp < n_groups) {
                uint64_t base = grp;
                for (uint32_t j = 0; j < k; j++) {
                    const uint64_t lo = base & ((1ull << sorted[j]) - 1ull);
                    base = ((base >> sorted[j]) << (sorted[j] + 1)) | lo;
                }
                if (((base | g.fixed_bits) & g.ctrl_mask) == g.ctrl_val) {
                    const qvc* row = mat + (size_t)r * d;
                    const qvc* col = in + ((idx >> k) << k);
                    qvc acc;
                    acc.x = 0.0;
                    acc.y = 0.0;
                    for (uint32_t c = 0; c < d; c++) acc = qv_cmadd(acc, row[c], col[c]);
                    uint64_t a = base;
                    for (uint32_t j = 0; j < k; j++)
                        if (r >> j & 1) a |= 1ull << g.pos[j];
                    psi[a] = acc;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Dense k-qubit gate (3 <= k <= 8 mixing qubits) on the FP64 TENSOR path: mma.sync.m8n8k4.f64 (SASS: DMMA).
// The complex 2^k x 2^k matrix is applied as the real (2d x 2d) block matrix W = [[Re, -Im], [Im, Re]] with rows and
// columns interleaved (re, im) per amplitude, to a panel of G groups staged in shared memory:
//     Y^T (groups x outputs) = X^T (groups x inputs) . W^T (inputs x outputs)
// so that the A fragment of a lane is one real of one group (row-major panel in shared memory, padded by 4 doubles per row
// against bank conflicts), the B fragment is W[output][input] straight from the row-major matrix, and the two accumulators of
// a lane are (re, im) of ONE output amplitude of ONE group: results go back to the panel as 16-byte stores, no shuffles.
// Why: measured on B200 (scripts/fp64_probe.cu, gpurun_out/r2d_fp64_probe.txt) the tensor path sustains 55 FP64 FMA per clock
// per SM against 38 for DFMA; a dense k >= 5 gate is bound by FP64, not by HBM (4 * 2^k FMA per amplitude).
// ---------------------------------------------------------------------------
#define QV_MMA_ELEMS 2048
__device__ __forceinline__ void qv_dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MT>   // m-tiles (8 groups each) per work item: they share every B (matrix) fragment
__device__ __forceinline__ void qv_mma_item(const double* __restrict__ xa, const double* __restrict__ wb, uint32_t D2, uint32_t ld,
                                            double* __restrict__ yo) {
    double c0[MT], c1[MT];
#pragma unroll
    for (int m = 0; m < MT; m++) c0[m] = c1[m] = 0.0;
    for (uint32_t k0 = 0; k0 < D2; k0 += 4) {
        const double b = wb[k0];
#pragma unroll
        for (int m = 0; m < MT; m++) qv_dmma(c0[m], c1[m], xa[(size_t)m * 8 * ld + k0], b);
    }
#pragma unroll
    for (int m = 0; m < MT; m++) {
        qvc v;
        v.x = c0[m];
        v.y = c1[m];
        *reinterpret_cast<qvc*>(yo + (size_t)m * 8 * ld) = v;
    }
}

__global__ void __launch_bounds__(QV_THREADS)
qv_bigmma_kernel(qvc* __restrict__ psi, QvBigGate g, const double* __restrict__ Wg, uint32_t n_bits, uint32_t w_in_smem) {
    extern __shared__ __align__(16) double qv_mma_smem[];
    const uint32_t k = g.k, d = 1u << k, D2 = 2u * d;
    const uint32_t G = QV_MMA_ELEMS >> k;             // groups per panel (>= 8 for k <= 8)
    const uint32_t ld = D2 + 4;                       // padded row of the panel, in doubles
    double* X = qv_mma_smem;
    double* Y = X + (size_t)G * ld;
    double* Ws = Y + (size_t)G * ld;
    const uint32_t ldw = w_in_smem ? D2 + 4 : D2;
    const double* W = w_in_smem ? Ws : Wg;
    if (w_in_smem)
        for (uint32_t i = threadIdx.x; i < D2 * D2; i += QV_THREADS) Ws[(i / D2) * ldw + (i % D2)] = Wg[i];
    const uint64_t n_groups = 1ull << (n_bits - k);
    uint32_t sorted[16];
    for (uint32_t j = 0; j < k; j++) sorted[j] = g.pos[j];
    for (uint32_t i = 1; i < k; i++) {
        uint32_t v = sorted[i];
        int j = (int)i - 1;
        while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; j--; }
        sorted[j + 1] = v;
    }
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n_tiles = D2 / 8, m_tiles = G / 8;
    // Address arithmetic off the per-element path: element idx = threadIdx.x + 256*i of a panel is (group idx >> k, member
    // idx & (d-1)); d <= 256 divides the block size, so a thread always handles the SAME member c -- its offset
    // sum_j bit_j(c) << pos[j] is computed once -- and the group bases (zeros inserted at the gate's positions) are computed
    // once per panel by G threads into shared memory.
    __shared__ uint64_t s_base[QV_MMA_ELEMS >> 3];
    const uint32_t my_c = threadIdx.x & (d - 1);
    uint64_t my_dep = 0;
    for (uint32_t j = 0; j < k; j++)
        if (my_c >> j & 1) my_dep |= 1ull << g.pos[j];
    for (uint64_t gb = (uint64_t)blockIdx.x * G; gb < n_groups; gb += (uint64_t)gridDim.x * G) {
        for (uint32_t t = threadIdx.x; t < G; t += QV_THREADS) {
            uint64_t base = gb + t;
            for (uint32_t j = 0; j < k; j++) {
                const uint64_t lo = base & ((1ull << sorted[j]) - 1ull);
                base = ((base >> sorted[j]) << (sorted[j] + 1)) | lo;
            }
            // bit 63 marks groups outside the state or not selected by the gate's controls (left untouched)
            const bool live = gb + t < n_groups;
            const bool sel = live && (((base | g.fixed_bits) & g.ctrl_mask) == g.ctrl_val);
            s_base[t] = base | (live ? 0ull : (1ull << 63)) | (sel ? 0ull : (1ull << 62));
        }
        __syncthreads();
        // gather the panel (coalesced where the gate's positions allow)
        for (uint32_t idx = threadIdx.x; idx < QV_MMA_ELEMS; idx += QV_THREADS) {
            const uint64_t b = s_base[idx >> k];
            qvc v;
            v.x = 0.0;
            v.y = 0.0;
            if (!(b >> 63)) v = qv_ld_stream(psi + ((b & ~(3ull << 62)) | my_dep));
            *reinterpret_cast<qvc*>(X + (size_t)(idx >> k) * ld + 2 * my_c) = v;
        }
        __syncthreads();
        // work items: (n-tile of 8 output reals) x (block of MT m-tiles); consecutive warps take consecutive n-tiles
        if (m_tiles >= 4) {
            const uint32_t m_blocks = m_tiles / 4;
            for (uint32_t item = warp; item < n_tiles * m_blocks; item += QV_THREADS / 32) {
                const uint32_t nt = item % n_tiles, mb = item / n_tiles;
                qv_mma_item<4>(X + (size_t)(mb * 32 + lane / 4) * ld + lane % 4, W + (size_t)(nt * 8 + lane / 4) * ldw + lane % 4, D2, ld,
                               Y + (size_t)(mb * 32 + lane / 4) * ld + nt * 8 + 2 * (lane % 4));
            }
        } else {
            for (uint32_t item = warp; item < n_tiles * m_tiles; item += QV_THREADS / 32) {
                const uint32_t nt = item % n_tiles, mt = item / n_tiles;
                qv_mma_item<1>(X + (size_t)(mt * 8 + lane / 4) * ld + lane % 4, W + (size_t)(nt * 8 + lane / 4) * ldw + lane % 4, D2, ld,
                               Y + (size_t)(mt * 8 + lane / 4) * ld + nt * 8 + 2 * (lane % 4));
            }
        }
        __syncthreads();
        for (uint32_t idx = threadIdx.x; idx < QV_MMA_ELEMS; idx += QV_THREADS) {
            const uint64_t b = s_base[idx >> k];
            if (!(b >> 62))
                qv_st_stream(psi + (b | my_dep), *reinterpret_cast<const qvc*>(Y + (size_t)(idx >> k) * ld + 2 * my_c));
        }
        __syncthreads();
    }
}

// Diagonal gate on k > QV_MAX_CHUNK_BITS qubits: psi_i *= table[bits of i at the gate's positions].  Element-wise, 32 B per
// amplitude; positions may be rank bits (their value comes from fixed_bits).
__global__ void __launch_bounds__(QV_THREADS)
qv_bigdiag_kernel(qvc* __restrict__ psi, QvBigGate g, const qvc* __restrict__ table, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * QV_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QV_THREADS + threadIdx.x; i < count; i += stride) {
        const uint64_t full = i | g.fixed_bits;
        uint32_t idx = 0;
        for (uint32_t j = 0; j < g.k; j++) idx |= (uint32_t)((full >> g.pos[j]) & 1ull) << j;
        qvc a = qv_ld_stream(psi + i);
        qv_cmul_ip(a, table[idx]);
        qv_st_stream(psi + i, a);
    }
}

// ---------------------------------------------------------------------------
// Reductions: per-thread strided partial sums, warp-shuffle + shared-memory
// block reduction, one partial per CTA; a second single-CTA kernel adds the
// partials in a fixed order (run-to-run deterministic, unlike the reference's
// completion-order PSUM-DOTIMES, src/utilities.lisp:395-425).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double qv_block_sum(double s) {
    __shared__ double warp_part[QV_THREADS / 32];
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < QV_THREADS / 32; w++) tot += warp_part[w];
    }
    __syncthreads();
    return tot;   // valid in thread 0
}

// mode 0: sum |psi_i|^2 over all i                      (NORM, src/wavefunction.lisp:333-347)
// mode 1: sum |psi_a|^2, a = inject(i, q) | 1<<q        (WAVEFUNCTION-EXCITED-STATE-PROBABILITY :64-70)
// mode 3: sum |psi_a|^2, a = inject(i, q)               (WAVEFUNCTION-GROUND-STATE-PROBABILITY :54-60)
// mode 2: sum Re rho[a*dim + a], a = inject(i, q)|1<<q  (density GET-EXCITED-STATE-PROBABILITY, measurement.lisp:77-85)
__global__ void __launch_bounds__(QV_THREADS)
qv_reduce_kernel(const qvc* __restrict__ psi, uint64_t count, int mode, uint32_t q, uint64_t dim, double* __restrict__ partial) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * QV_THREADS;
    const uint64_t lowmask = (1ull << q) - 1ull;
    auto addr = [&](uint64_t i) -> uint64_t {
        if (mode == 0) return i;
        const uint64_t a = ((i & ~lowmask) << 1) | (i & lowmask) | (mode == 3 ? 0ull : (1ull << q));
        return mode == 2 ? a * dim + a : a;
    };
    uint64_t i = (uint64_t)blockIdx.x * QV_THREADS + threadIdx.x;
    if (mode == 2) {
        for (; i < count; i += stride) s0 += psi[addr(i)].x;
    } else {
        for (; i + 3 * stride < count; i += 4 * stride) {
            const qvc a = psi[addr(i)], b = psi[addr(i + stride)], c = psi[addr(i + 2 * stride)], d = psi[addr(i + 3 * stride)];
            s0 += a.x * a.x + a.y * a.y;
            s1 += b.x * b.x + b.y * b.y;
            s2 += c.x * c.x + c.y * c.y;
            s3 += d.x * d.x + d.y * d.y;
        }
        for (; i < count; i += stride) {
            const qvc a = psi[addr(i)];
            s0 += a.x * a.x + a.y * a.y;
        }
    }
    const double tot = qv_block_sum((s0 + s1) + (s2 + s3));
    if (threadIdx.x == 0) partial[blockIdx.x] = tot;
}

// <a|b> = sum conj(a_i) b_i (PURE-STATE-EXPECTATION's INNER-PRODUCT, app/src/api/expectation.lisp:79-84): per-CTA partial
// sums of the real and the imaginary part (partial[2*blockIdx], partial[2*blockIdx + 1]); reads 32*count bytes.
__global__ void __launch_bounds__(QV_THREADS)
qv_inner_kernel(const qvc* __restrict__ a, const qvc* __restrict__ b, uint64_t count, double* __restrict__ partial) {
    double re0 = 0.0, im0 = 0.0, re1 = 0.0, im1 = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * QV_THREADS;
    uint64_t i = (uint64_t)blockIdx.x * QV_THREADS + threadIdx.x;
    for (; i + stride < count; i += 2 * stride) {
        const qvc x0 = qv_ld_stream(a + i), y0 = qv_ld_stream(b + i);
        const qvc x1 = qv_ld_stream(a + i + stride), y1 = qv_ld_stream(b + i + stride);
        re0 += x0.x * y0.x + x0.y * y0.y;
        im0 += x0.x * y0.y - x0.y * y0.x;
        re1 += x1.x * y1.x + x1.y * y1.y;
        im1 += x1.x * y1.y - x1.y * y1.x;
    }
    for (; i < count; i += stride) {
        const qvc x0 = qv_ld_stream(a + i), y0 = qv_ld_stream(b + i);
        re0 += x0.x * y0.x + x0.y * y0.y;
        im0 += x0.x * y0.y - x0.y * y0.x;
    }
    const double re = qv_block_sum(re0 + re1);
    const double im = qv_block_sum(im0 + im1);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = re;
        partial[2 * blockIdx.x + 1] = im;
    }
}

// tr(Q rho) = sum_{i,j} Q[i][j] rho[j][i] (MIXED-STATE-EXPECTATION, app/src/api/expectation.lisp:91-107): rho = vec(rho)
// row-major on 2n bits, Q row-major dim x dim on the device; per-CTA partial sums (re, im) like qv_inner_kernel.
__global__ void __launch_bounds__(QV_THREADS)
qv_trace_product_kernel(const qvc* __restrict__ rho, const qvc* __restrict__ Q, uint32_t n_qubits, double* __restrict__ partial) {
    double re = 0.0, im = 0.0;
    const uint64_t dim = 1ull << n_qubits, count = dim * dim;
    const uint64_t stride = (uint64_t)gridDim.x * QV_THREADS;
    for (uint64_t idx = (uint64_t)blockIdx.x * QV_THREADS + threadIdx.x; idx < count; idx += stride) {
        const uint64_t j = idx >> n_qubits, i = idx & (dim - 1);      // rho[j][i]
        const qvc r = qv_ld_stream(rho + idx), q = Q[i * dim + j];
        re += q.x * r.x - q.y * r.y;
        im += q.x * r.y + q.y * r.x;
    }
    const double tre = qv_block_sum(re);
    const double tim = qv_block_sum(im);
    if (threadIdx.x == 0) {
        partial[2 * blockIdx.x] = tre;
        partial[2 * blockIdx.x + 1] = tim;
    }
}

// vec(I) on 2n bits: the zero state of UNITARY-STATE (src/unitary-qvm.lisp:57-61)
__global__ void __launch_bounds__(QV_THREADS) qv_set_identity_kernel(qvc* __restrict__ psi, uint64_t dim) {
    const uint64_t i = (uint64_t)blockIdx.x * QV_THREADS + threadIdx.x;
    if (i < dim) {
        psi[i * dim + i].x = 1.0;
        psi[i * dim + i].y = 0.0;
    }
}

// out[0] = sum partial[2k], out[1] = sum partial[2k+1], fixed order
__global__ void __launch_bounds__(QV_THREADS) qv_final_sum2_kernel(const double* __restrict__ partial, uint32_t n, double* out) {
    double re = 0.0, im = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += QV_THREADS) {
        re += partial[2 * i];
        im += partial[2 * i + 1];
    }
    const double tre = qv_block_sum(re);
    const double tim = qv_block_sum(im);
    if (threadIdx.x == 0) {
        out[0] = tre;
        out[1] = tim;
    }
}

__global__ void __launch_bounds__(QV_THREADS) qv_final_sum_kernel(const double* __restrict__ partial, uint32_t n, double* out) {
    double s = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += QV_THREADS) s += partial[i];
    const double tot = qv_block_sum(s);
    if (threadIdx.x == 0) *out = tot;
}

// ---------------------------------------------------------------------------
// Pull remap (multi-GPU): dst[p] = current[src_rank][src_off] where (src_rank, src_off) is the destination
// index (rank, p) with every (local_bit, global_bit) pair swapped.  Local bits are >= 4, so 16 consecutive
// destinations share a 256-byte source run: coalesced 128-bit NVLink reads, local streaming writes.
// Bound: NVLink ingress, (1 - 2^-pairs) of the shard (measured peer-read peak ~ 675-785 GB/s, scripts/p2p_probe.cu).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(QV_THREADS)
qv_remap_pull_kernel(const __grid_constant__ QvPeers cur, qvc* __restrict__ dst, const __grid_constant__ QvRemap rm) {
    const uint32_t n_local = rm.n_local_bits;
    const uint64_t n = 1ull << n_local;
    const uint64_t local_mask = n - 1ull;
    const uint64_t rank_bits = (uint64_t)rm.rank << n_local;
    const uint64_t stride = (uint64_t)gridDim.x * QV_THREADS * 4;
    for (uint64_t p0 = ((uint64_t)blockIdx.x * QV_THREADS * 4) + threadIdx.x; p0 < n; p0 += stride) {
        qvc v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t pj = p0 + (uint64_t)j * QV_THREADS;
            if (pj >= n) continue;
            const uint64_t P = rank_bits | pj;
            uint64_t S = P;
            for (uint32_t i = 0; i < rm.n_pairs; i++) {
                const uint64_t lb = (P >> rm.local_bit[i]) & 1ull, gb = (P >> rm.global_bit[i]) & 1ull;
                const uint64_t x = lb ^ gb;
                S ^= (x << rm.local_bit[i]) | (x << rm.global_bit[i]);
            }
            v[j] = qv_ld_stream(cur.base[(S >> n_local) & (QV_MAX_PEERS - 1)] + (S & local_mask));
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (p0 + (uint64_t)j * QV_THREADS < n) qv_st_stream(dst + p0 + (uint64_t)j * QV_THREADS, v[j]);
    }
}

// ---------------------------------------------------------------------------
// Element-wise passes.
// ---------------------------------------------------------------------------
// mode 0: psi *= f                                                 (NORMALIZE-WAVEFUNCTION wavefunction.lisp:349-364)
// mode 1: pure collapse on bit q  (FORCE-MEASUREMENT measurement.lisp:10-41): zero when bit != keep, else *= f
// mode 2: density collapse on bits q and q2 (measurement.lisp:43-68)
// mode 3: density measure-discard: zero when bit q != bit q2 (measurement.lisp:111-120)
__global__ void __launch_bounds__(QV_THREADS)
qv_elementwise_kernel(qvc* __restrict__ psi, uint64_t count, int mode, uint32_t q, uint32_t q2, uint32_t keep, double f) {
    const uint64_t stride = (uint64_t)gridDim.x * QV_THREADS;
    const qvc zero = {0.0, 0.0};
    for (uint64_t i = (uint64_t)blockIdx.x * QV_THREADS + threadIdx.x; i < count; i += stride) {
        bool kill = false, touch = true;
        if (mode == 1) kill = ((i >> q) & 1ull) != keep;
        else if (mode == 2) kill = (((i >> q) & 1ull) != keep) || (((i >> q2) & 1ull) != keep);
        else if (mode == 3) { kill = ((i >> q) & 1ull) != ((i >> q2) & 1ull); touch = kill; }
        if (kill) qv_st_stream(psi + i, zero);       // the annihilated half is never read
        else if (touch) {
            qvc a = qv_ld_stream(psi + i);
            a.x = f * a.x;
            a.y = f * a.y;
            qv_st_stream(psi + i, a);
        }
    }
}

// out[i] = |psi[first + i]|^2, the PROBABILITY of every basis state (src/wavefunction.lisp:44-50) for the
// :probabilities export path (app/src/api/probabilities.lisp, handle-request.lisp:155-176): 8 instead of 16 bytes
// per amplitude leave the device.  Explicitly rounded products and sum: bit-exact against the oracle.
__global__ void __launch_bounds__(QV_THREADS)
qv_probs_kernel(const qvc* __restrict__ psi, uint64_t first, uint64_t count, double* __restrict__ out) {
    const uint64_t stride = (uint64_t)gridDim.x * QV_THREADS;
    for (uint64_t i = (uint64_t)blockIdx.x * QV_THREADS + threadIdx.x; i < count; i += stride) {
        const qvc a = qv_ld_stream(psi + first + i);
        out[i] = __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y));
    }
}

__global__ void qv_set_one_kernel(qvc* psi, uint64_t index) {
    psi[index].x = 1.0;
    psi[index].y = 0.0;
}

__global__ void __launch_bounds__(QV_THREADS) qv_diag_probs_kernel(const qvc* __restrict__ rho, uint64_t dim, double* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * QV_THREADS + threadIdx.x;
    if (i < dim) out[i] = rho[i + i * dim].x;
}

// ---------------------------------------------------------------------------
// Sampler: blocked prefix structure + per-shot descent.  The summation order is
// the one restated in oracle/qvm_oracle.c (orc_sample_tree); all sums use
// explicitly rounded, never-contracted adds/multiplies so the indices are
// bit-exact against that oracle for identical uniforms.
// ---------------------------------------------------------------------------
#define QV_SB 1024
__device__ __forceinline__ double qv_prob_rn(qvc a) { return __dadd_rn(__dmul_rn(a.x, a.x), __dmul_rn(a.y, a.y)); }

// one warp per leaf block of 1024 amplitudes (level 1) or 1024 level-1 sums (level 2)
__global__ void __launch_bounds__(QV_THREADS)
qv_sample_build_kernel(const qvc* __restrict__ psi, const double* __restrict__ src, uint64_t count, double* __restrict__ dst, uint64_t n_blocks) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warp = ((uint64_t)blockIdx.x * QV_THREADS + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * QV_THREADS) >> 5;
    for (uint64_t b = warp; b < n_blocks; b += n_warps) {
        double s = 0.0;
#pragma unroll 8
        for (int j = 0; j < 32; j++) {
            const uint64_t idx = b * QV_SB + lane + 32u * (uint32_t)j;
            double v = 0.0;
            if (idx < count) v = psi ? qv_prob_rn(psi[idx]) : src[idx];
            s = __dadd_rn(s, v);
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s = __dadd_rn(s, __shfl_xor_sync(0xffffffffu, s, off));
        if (lane == 0) dst[b] = s;
    }
}

__global__ void qv_sample_top_kernel(const double* __restrict__ l2, uint64_t n2, double* __restrict__ top) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (uint64_t c = 0; c < n2; c++) {
            s = __dadd_rn(s, l2[c]);
            top[c] = s;
        }
    }
}

__device__ __forceinline__ bool qv_rule_hit(double c, double p, int strict) { return strict ? (c > p) : (c >= p); }

__global__ void __launch_bounds__(128)
qv_sample_descend_kernel(const qvc* __restrict__ psi, uint64_t n_amps, const double* __restrict__ l1, uint64_t n1,
                         const double* __restrict__ top, uint64_t n2, const double* __restrict__ u, uint64_t n_shots,
                         int strict, double base, uint64_t* __restrict__ out) {
    // base: probability mass in front of this vector (0 for a whole state; the preceding shards' totals for a shard).
    // Adding 0.0 is exact, so the single-device indices do not depend on it (oracle: sample_tree_base).
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_shots) return;
    const double p = u[t];
    uint64_t lo = 0, hi = n2 - 1;
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (qv_rule_hit(__dadd_rn(base, top[mid]), p, strict)) hi = mid;
        else lo = mid + 1;
    }
    const uint64_t c = lo;
    const double acc = (c == 0) ? base : __dadd_rn(base, top[c - 1]);
    const uint64_t b0 = c * QV_SB;
    uint64_t bcnt = n1 - b0;
    if (bcnt > QV_SB) bcnt = QV_SB;
    double run = 0.0, before = 0.0;
    uint64_t b = b0 + bcnt - 1;
    bool found = false;
    for (uint64_t j = 0; j < bcnt; j++) {
        const double nr = __dadd_rn(run, l1[b0 + j]);
        if (qv_rule_hit(__dadd_rn(acc, nr), p, strict)) {
            b = b0 + j;
            before = run;
            found = true;
            break;
        }
        if (j + 1 < bcnt) run = nr;   // keep the sum of the first bcnt-1 entries if nothing hits
    }
    if (!found) before = run;
    const double acc2 = __dadd_rn(acc, before);
    const uint64_t i0 = b * QV_SB;
    uint64_t icnt = n_amps - i0;
    if (icnt > QV_SB) icnt = QV_SB;
    uint64_t r = i0 + icnt - 1;
    run = 0.0;
    for (uint64_t j = 0; j < icnt; j++) {
        run = __dadd_rn(run, qv_prob_rn(psi[i0 + j]));
        if (qv_rule_hit(__dadd_rn(acc2, run), p, strict)) {
            r = i0 + j;
            break;
        }
    }
    out[t] = r;
}
