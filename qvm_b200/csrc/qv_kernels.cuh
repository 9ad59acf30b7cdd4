// qv_kernels.cuh -- sm_100a kernels of libqvmcuda.
//
// Roofline of every kernel here is HBM bandwidth: a gate pass
// reads and writes each 16-byte amplitude once (32 B/amplitude) whatever the gate
// arity, so the design goal is (1) full-width coalesced 128-bit accesses for every
// qubit position and (2) as many gates as possible per pass.
#pragma once
#include <cuda_runtime.h>

#include "qv_ops.h"

struct QvPeers {
    qvc* base[QV_MAX_PEERS];   // shard base pointer of every rank (own pointer at [rank])
};

// Streaming 128-bit accesses that do not allocate in L1: L1 is kept for the tile
// program (ops, matrices, diagonal tables), which every CTA re-reads.
__device__ __forceinline__ qvc qv_ld_stream(const qvc* p) {
    qvc v;
    asm volatile("ld.global.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
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

// ---------------------------------------------------------------------------
// The tile kernel: one CTA = one tile of 2^T amplitudes staged in shared memory
// (XOR-swizzled, see qv_swz), rounds of register-resident groups, write back.
//   M = 3 : 256 threads, 8 amplitudes per thread per round, <= 80 registers  }  64 KiB tile + 10.5 KiB of
//   M = 4 : 128 threads, 16 amplitudes per thread per round, <= 168 registers }  per-tile tables -> 3 CTAs / SM
//           (fewer, fatter threads: micro-op decode is amortised over twice the amplitudes and a
//           pass over 8 tile bits needs 2 rounds instead of 3; used for passes that carry many gates)
//   FULL  = true : T == 12 (every state of >= 12 qubits): all loop bounds are compile-time, the
//                  tile-local -> physical address of element tid + THREADS*i is
//                  (base | gather(tid)) | hi_off[i] with hi_off precomputed by the host.
//   PEERS = true : tile bits include rank bits, amplitudes come from / go to peer
//                  shards over NVLink (P2P loads/stores on IPC-mapped pointers).
// ---------------------------------------------------------------------------
struct QvProgSmall { uint8_t bytes[QV_PROG_SMALL_BYTES]; };
struct QvProgLarge { uint8_t bytes[QV_PROG_LARGE_BYTES]; };

template <typename PROG, bool PEERS, bool FULL, int M>
__global__ void __launch_bounds__((M == 4 ? QV_THREADS_WIDE : QV_THREADS), 3)
qv_tile_kernel(const __grid_constant__ PROG prog, const __grid_constant__ QvPeers peers,
               const qvc* __restrict__ tables) {
    constexpr int THREADS = (M == 4 ? QV_THREADS_WIDE : QV_THREADS);
    constexpr int NS = 1 << M;
    constexpr int ITERS = 4096 / THREADS;
    extern __shared__ __align__(16) uint8_t qv_smem_raw[];
    qvc* tile = reinterpret_cast<qvc*>(qv_smem_raw);
    __shared__ qvc s_slice[QV_SLICE_ENTRIES];
    __shared__ uint32_t s_srcext[QV_MAX_SOURCES];
    __shared__ uint8_t s_pred[QV_MAX_PREDS];

    // The control program sits in the constant bank (kernel parameters): every read below is a
    // uniform constant load, matrices reach the FP64 pipe through uniform registers.
    const uint8_t* blob = prog.bytes;
    const QvPassHeader* h = reinterpret_cast<const QvPassHeader*>(blob);
    const uint32_t T = FULL ? 12u : h->T;
    const uint32_t tile_n = 1u << T;
    const uint64_t fixed_bits = h->fixed_bits;
    const uint64_t n_tiles = h->n_tiles;
    const uint32_t n_local = h->n_local_bits;
    const uint64_t local_mask = (1ull << n_local) - 1ull;
    const uint32_t n_rounds = h->n_rounds;
    const QvRound* rounds = reinterpret_cast<const QvRound*>(blob + h->off_rounds);
    const QvUop* uops = reinterpret_cast<const QvUop*>(blob + h->off_uops);
    const QvSource* sources = reinterpret_cast<const QvSource*>(blob + h->off_sources);
    const QvSlice* slices = reinterpret_cast<const QvSlice*>(blob + h->off_slices);
    const uint8_t* slice_of = blob + h->off_slice_of;
    const QvPred* preds = reinterpret_cast<const QvPred*>(blob + h->off_preds);

    const uint32_t tid = threadIdx.x;
    const uint32_t iters = FULL ? (uint32_t)ITERS : (tile_n + THREADS - 1) / THREADS;
    // tile-local e = tid + THREADS*i: the gather is bitwise linear, so split it.
    const uint64_t glo = qv_gather((uint64_t)tid, h->tile_segs, h->n_tile_segs);
    // qv_swz only mixes bits 3..5 into bits 0..2, so the slot of tid + THREADS*i is qv_swz(tid) + THREADS*i
    qvc* const my_tile = tile + qv_swz(tid);
    qvc* const own = peers.base[PEERS ? 0 : (fixed_bits >> n_local) & (QV_MAX_PEERS - 1)];
    // store permutation (trailing X / CNOT / SWAP gates): slot read for destination e = tid + THREADS*i
    const bool store_perm = h->store_perm != 0;
    uint32_t st_lo = h->st_const;
    if (store_perm)
        for (uint32_t k = 0; k < T; k++)
            if (tid >> k & 1) st_lo ^= h->st_col[k];

    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t base = qv_gather(t, h->base_segs, h->n_base_segs) | fixed_bits;
        const uint64_t pbase = PEERS ? (base | glo) : ((base | glo) & local_mask);

        // ---- HBM -> shared memory: asynchronous 16-byte copies, all of a thread's copies in flight at once
        if (FULL) {
#pragma unroll
            for (int i = 0; i < ITERS; i++) {
                const uint64_t p = pbase | h->hi_off[i];
                const qvc* src = PEERS ? peers.base[(p >> n_local) & (QV_MAX_PEERS - 1)] + (p & local_mask) : own + p;
                qv_cp_async16(my_tile + i * THREADS, src);
            }
        } else {
            for (uint32_t i = 0; i < iters; i++) {
                const uint32_t e = tid + i * THREADS;
                if (e < tile_n) {
                    const uint64_t p = pbase | h->hi_off[i];
                    const qvc* src = PEERS ? peers.base[(p >> n_local) & (QV_MAX_PEERS - 1)] + (p & local_mask) : own + p;
                    qv_cp_async16(my_tile + i * THREADS, src);
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");

        // ---- per-tile tables, built while the copies fly: source offsets, control predicates,
        //      then the diagonal slices (all factors whose external bits are constant over this tile
        //      collapse into small shared-memory tables)
        if (h->n_sources | h->n_preds) {
            for (uint32_t i = tid; i < h->n_sources; i += THREADS)
                s_srcext[i] = (uint32_t)qv_gather(base, sources[i].esegs, sources[i].n_esegs) << sources[i].nl;
            for (uint32_t i = tid; i < h->n_preds; i += THREADS)
                s_pred[i] = (base & preds[i].mask) == preds[i].val ? 1 : 0;
            __syncthreads();
            for (uint32_t f = tid; f < h->n_slice_entries; f += THREADS) {
                const QvSlice& sl = slices[slice_of[f]];
                s_slice[f] = qv_slice_entry(sl, sources, s_srcext, tables, f - sl.off);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();

        // ---- rounds: 2^m amplitudes per thread in registers, every micro-op of the round applied there
        for (uint32_t r = 0; r < n_rounds; r++) {
            const QvRound& rd = rounds[r];
            const uint32_t m = FULL ? (uint32_t)M : rd.m;
            const uint32_t nslots = 1u << m;
            const uint32_t ngroups = tile_n >> m;
            for (uint32_t g = tid; g < ngroups; g += THREADS) {
                uint32_t e0 = g;
                if (m > 0) e0 = qv_insert_zero(e0, rd.regpos[0]);
                if (m > 1) e0 = qv_insert_zero(e0, rd.regpos[1]);
                if (m > 2) e0 = qv_insert_zero(e0, rd.regpos[2]);
                if (M > 3 && m > 3) e0 = qv_insert_zero(e0, rd.regpos[3]);
                const uint32_t se0 = qv_swz(e0);
                qvc a[NS];
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    if (FULL || (uint32_t)s < nslots) a[s] = tile[se0 ^ rd.slot_xor[s]];
                    else { a[s].x = 0.0; a[s].y = 0.0; }
                }
                // Micro-op loop: the list ends with a QV_K_END sentinel, so the loop condition is the kind that is
                // dispatched on anyway; the next header is fetched while the current micro-op runs.
                const QvUopHead* hp = reinterpret_cast<const QvUopHead*>(uops + rd.first_uop);
                QvUopHead nh = hp[0];
                while ((nh.w0 & 0xffu) != QV_K_END) {
                    const QvUopHead ch = nh;
                    const QvUop& cu = *reinterpret_cast<const QvUop*>(hp);
                    hp += sizeof(QvUop) / sizeof(QvUopHead);
                    nh = hp[0];
                    qv_run_uop<NS>(a, ch, cu, g, blob, tables, s_slice, s_pred);
                }
#pragma unroll
                for (int s = 0; s < NS; s++)
                    if (FULL || (uint32_t)s < nslots) tile[se0 ^ rd.slot_xor[s]] = a[s];
            }
            __syncthreads();
        }

        // ---- shared memory -> HBM
        if (FULL && !store_perm) {
#pragma unroll
            for (int i = 0; i < ITERS; i++) {
                const uint64_t p = pbase | h->hi_off[i];
                qvc* dst = PEERS ? peers.base[(p >> n_local) & (QV_MAX_PEERS - 1)] + (p & local_mask) : own + p;
                qv_st_stream(dst, my_tile[i * THREADS]);
            }
        } else if (FULL) {
#pragma unroll
            for (int i = 0; i < ITERS; i++) {
                const uint64_t p = pbase | h->hi_off[i];
                qvc* dst = PEERS ? peers.base[(p >> n_local) & (QV_MAX_PEERS - 1)] + (p & local_mask) : own + p;
                qv_st_stream(dst, tile[st_lo ^ h->st_hi[i]]);
            }
        } else {
            for (uint32_t i = 0; i < iters; i++) {
                const uint32_t e = tid + i * THREADS;
                if (e < tile_n) {
                    const uint64_t p = pbase | h->hi_off[i];
                    qvc* dst = PEERS ? peers.base[(p >> n_local) & (QV_MAX_PEERS - 1)] + (p & local_mask) : own + p;
                    qv_st_stream(dst, store_perm ? tile[st_lo ^ h->st_hi[i]] : my_tile[i * THREADS]);
                }
            }
        }
        __syncthreads();
    }
}

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
                         int strict, uint64_t* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_shots) return;
    const double p = u[t];
    uint64_t lo = 0, hi = n2 - 1;
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (qv_rule_hit(top[mid], p, strict)) hi = mid;
        else lo = mid + 1;
    }
    const uint64_t c = lo;
    const double acc = (c == 0) ? 0.0 : top[c - 1];
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
