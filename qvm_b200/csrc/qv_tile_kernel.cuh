// qv_tile_kernel.cuh -- the sm_100a tile kernel (the gate pass) and the device helpers it shares with the
// streaming kernels of qv_kernels.cuh.
//
// Roofline of every kernel here is HBM bandwidth: a gate pass
// reads and writes each 16-byte amplitude once (32 B/amplitude) whatever the gate
// arity, so the design goal is (1) full-width coalesced 128-bit accesses for every
// qubit position and (2) as many gates as possible per pass.
#pragma once
#include "qv_tile_common.cuh"

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
//   MODE = 0 : local pass, in place.
//   MODE = 1 : peer pass: tile bits include rank bits, amplitudes come from / go to peer shards over NVLink
//              (P2P loads/stores on IPC-mapped pointers), in place.
//   MODE = 2 : pull pass: loads go through a pending qubit remap (from every rank's CURRENT buffer, over
//              NVLink where the source is remote), results are written to this rank's alternate buffer.
// ---------------------------------------------------------------------------
template <typename PROG, int MODE, bool FULL, int M>
__global__ void __launch_bounds__((M == 4 ? QV_THREADS_WIDE : QV_THREADS), 3)
qv_tile_kernel(const __grid_constant__ PROG prog, const __grid_constant__ QvPeers peers,
               const qvc* __restrict__ tables, qvc* __restrict__ alt_own) {
    constexpr bool PEERS = MODE == 1;
    constexpr bool PULL = MODE == 2;
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
    const uint32_t zero_ranks = PULL ? h->zero_ranks : 0u;      // ranks whose current buffer is known to be all zeros
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
    qvc* const own = PULL ? alt_own : peers.base[PEERS ? 0 : (fixed_bits >> n_local) & (QV_MAX_PEERS - 1)];
    // store permutation (trailing X / CNOT / SWAP gates): slot read for destination e = tid + THREADS*i
    const bool store_perm = h->store_perm != 0;
    const bool has_scale = h->has_scale != 0;       // write-back scale (factors of the gates that ran as unscaled butterflies)
    const double out_scale = h->out_scale;
    uint32_t st_lo = h->st_const;
    if (store_perm)
        for (uint32_t k = 0; k < T; k++)
            if (tid >> k & 1) st_lo ^= h->st_col[k];

    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t base = qv_gather(t, h->base_segs, h->n_base_segs) | fixed_bits;
        const uint64_t pbase = PEERS ? (base | glo) : ((base | glo) & local_mask);
        const uint64_t sbase = PULL ? qv_remap_index(base | glo, h->pull_remap) : 0;     // source index of pbase

        // ---- HBM -> shared memory: asynchronous 16-byte copies, all of a thread's copies in flight at once
        if (FULL && MODE == 0) {
            // local pass: pbase and hi_off[i] have disjoint bits, so the address is one 64-bit add per element
            const char* tsrc = reinterpret_cast<const char*>(own + pbase);
#pragma unroll
            for (int i = 0; i < ITERS; i++) qv_cp_async16(my_tile + i * THREADS, reinterpret_cast<const qvc*>(tsrc + h->hi_byte[i]));
        } else if (FULL) {
#pragma unroll
            for (int i = 0; i < ITERS; i++) {
                const uint64_t p = PULL ? (sbase ^ h->hi_src[i]) : (pbase | h->hi_off[i]);
                const uint32_t pr = (uint32_t)(p >> n_local) & (QV_MAX_PEERS - 1);
                const qvc* src = (PEERS || PULL) ? peers.base[pr] + (p & local_mask) : own + p;
                if (PULL) qv_cp_async16_z(my_tile + i * THREADS, src, (zero_ranks >> pr & 1u) ? 0u : 16u);
                else qv_cp_async16(my_tile + i * THREADS, src);
            }
        } else {
            for (uint32_t i = 0; i < iters; i++) {
                const uint32_t e = tid + i * THREADS;
                if (e < tile_n) {
                    const uint64_t p = PULL ? (sbase ^ h->hi_src[i]) : (pbase | h->hi_off[i]);
                    const uint32_t pr = (uint32_t)(p >> n_local) & (QV_MAX_PEERS - 1);
                    const qvc* src = (PEERS || PULL) ? peers.base[pr] + (p & local_mask) : own + p;
                    if (PULL) qv_cp_async16_z(my_tile + i * THREADS, src, (zero_ranks >> pr & 1u) ? 0u : 16u);
                    else qv_cp_async16(my_tile + i * THREADS, src);
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
            // (unrolled x2 by nvcc: measured 63.1 ms vs 73.5 ms for QFT-30 with "#pragma unroll 1", which spills)
            for (uint32_t g_ = tid; g_ < ngroups; g_ += THREADS) {
                // keep the group counter in a register: ptxas otherwise re-derives it from SR_TID.X (a 20-cycle S2R)
                // in front of every micro-op
                uint32_t g = g_;
                asm volatile("" : "+r"(g));
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
        if (FULL && !PEERS && !store_perm) {
            char* tdst = reinterpret_cast<char*>(own + pbase);
#pragma unroll
            for (int i = 0; i < ITERS; i++) qv_st_stream(reinterpret_cast<qvc*>(tdst + h->hi_byte[i]), qv_scaled(my_tile[i * THREADS], has_scale, out_scale));
        } else if (FULL && !PEERS) {
            char* tdst = reinterpret_cast<char*>(own + pbase);
#pragma unroll
            for (int i = 0; i < ITERS; i++) qv_st_stream(reinterpret_cast<qvc*>(tdst + h->hi_byte[i]), qv_scaled(tile[st_lo ^ h->st_hi[i]], has_scale, out_scale));
        } else if (FULL && !store_perm) {
#pragma unroll
            for (int i = 0; i < ITERS; i++) {
                const uint64_t p = pbase | h->hi_off[i];
                qvc* dst = PEERS ? peers.base[(p >> n_local) & (QV_MAX_PEERS - 1)] + (p & local_mask) : own + p;
                qv_st_stream(dst, qv_scaled(my_tile[i * THREADS], has_scale, out_scale));
            }
        } else if (FULL) {
#pragma unroll
            for (int i = 0; i < ITERS; i++) {
                const uint64_t p = pbase | h->hi_off[i];
                qvc* dst = PEERS ? peers.base[(p >> n_local) & (QV_MAX_PEERS - 1)] + (p & local_mask) : own + p;
                qv_st_stream(dst, qv_scaled(tile[st_lo ^ h->st_hi[i]], has_scale, out_scale));
            }
        } else {
            for (uint32_t i = 0; i < iters; i++) {
                const uint32_t e = tid + i * THREADS;
                if (e < tile_n) {
                    const uint64_t p = pbase | h->hi_off[i];
                    qvc* dst = PEERS ? peers.base[(p >> n_local) & (QV_MAX_PEERS - 1)] + (p & local_mask) : own + p;
                    qv_st_stream(dst, qv_scaled(store_perm ? tile[st_lo ^ h->st_hi[i]] : my_tile[i * THREADS], has_scale, out_scale));
                }
            }
        }
        __syncthreads();
    }
}

