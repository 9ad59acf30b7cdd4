// qv_jit_kernel.cuh -- skeleton of a COMPILED gate pass.
//
// The pass compiler (qv_jit_gen.cpp) turns one tile program (qv_program.h) into CUDA C++ in which everything the
// interpreter kernel (qv_tile_kernel.cuh) decodes at run time is a literal: the rounds are unrolled, every
// micro-op is a direct call of its qv_ops.h template with compile-time register bits, control masks, index fields and
// blob / table offsets, and the matrices are read from the constant bank at fixed addresses (they reach the FP64 pipe
// as c[bank][offset] operands).  Gate matrices, diagonal tables and tile geometry stay DATA (the kernel parameter and
// the table pool), so passes with the same structure but other angles or qubit positions share one cubin.  This is the
// device-side analogue of the reference's per-gate compiled lambdas and their cache
// (src/compile-gate.lisp:156-209, 315-334), one level up: one compiled function per fused pass.
//
// The generated translation unit defines, before including this file:
//   QVJ_M, QVJ_THREADS, QVJ_MODE (0 local, 2 pull), QVJ_PROG_BYTES, QVJ_HAS_SCALE, QVJ_STORE_PERM, QVJ_HAS_TABLES,
//   the round functions qvj_round_<r>() and QVJ_RUN_ROUNDS (the sequence of rounds with barriers in between).
// With QVJ_HOST defined the same text compiles as plain C++ (tests/support: the round functions are run by the CPU
// emulator in place of its interpreter, which checks the generator without a GPU).
#pragma once

#if !defined(QVJ_HOST)

struct QvjProg { uint8_t bytes[QVJ_PROG_BYTES]; };

#if !defined(QVJ_TMA_LOAD)
#define QVJ_TMA_LOAD 0
#endif
#if !defined(QVJ_SRC_BASIS)
#define QVJ_SRC_BASIS 0
#endif
// Variant QVJ_SRC_BASIS: the state is a basis vector that was never written to HBM (SET-TO-ZERO-STATE is lazy): nothing is
// loaded.  A tile that does not contain the non-zero amplitude stays zero under any gate, so it is written back as zeros
// without running the rounds; the one tile that does is synthesised in shared memory.  The pass costs its 16 B/amplitude of
// writes only, and the reset itself costs nothing.
#if QVJ_TMA_LOAD
// Variant: the tile is loaded by the tensor-memory accelerator -- ONE cp.async.bulk.tensor per tile, issued by one thread,
// completion on an mbarrier -- instead of 16 LDGSTS per thread; everything else as below (three CTAs per SM).  See
// qv_jit_kernel_tma.cuh for the tile-as-a-box view and why SWIZZLE_128B reproduces qv_swz.
struct alignas(64) QvjTensorMap { unsigned long long opaque[16]; };      // CUtensorMap
__device__ __forceinline__ uint32_t qvj_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void qvj_mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "QVJ_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra QVJ_DONE;\n"
        "bra QVJ_WAIT;\n"
        "QVJ_DONE:\n"
        "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
#endif

extern "C" __global__ void __launch_bounds__(QVJ_THREADS, QVJ_MIN_CTAS)
qvj_kernel(const __grid_constant__ QvjProg prog, const __grid_constant__ QvPeers peers,
           const qvc* __restrict__ tables, qvc* __restrict__ alt_own
#if QVJ_TMA_LOAD
           , const __grid_constant__ QvjTensorMap tmap, const __grid_constant__ QvTmaGeom geom
#endif
           ) {
    constexpr bool PULL = QVJ_MODE == 2;
    constexpr int THREADS = QVJ_THREADS;
    constexpr int ITERS = 4096 / THREADS;
#if QVJ_TMA_LOAD
    extern __shared__ __align__(1024) uint8_t qv_smem_raw[];      // SWIZZLE_128B wants the box on a 1024-byte boundary
    __shared__ __align__(8) unsigned long long s_mbar;
#else
    extern __shared__ __align__(16) uint8_t qv_smem_raw[];
#endif
    qvc* tile = reinterpret_cast<qvc*>(qv_smem_raw);
#if QVJ_HAS_TABLES
    __shared__ qvc s_slice[QV_SLICE_ENTRIES];
    __shared__ uint32_t s_srcext[QV_MAX_SOURCES];
    __shared__ uint8_t s_pred[QV_MAX_PREDS];
#else
    const qvc* s_slice = nullptr;
    const uint8_t* s_pred = nullptr;
#endif
    const uint8_t* blob = prog.bytes;
    const QvPassHeader* h = reinterpret_cast<const QvPassHeader*>(blob);
    const uint64_t fixed_bits = h->fixed_bits;
    const uint64_t n_tiles = h->n_tiles;
    const uint32_t n_local = h->n_local_bits;
    const uint64_t local_mask = (1ull << n_local) - 1ull;
    const uint32_t tid = threadIdx.x;
    const uint64_t glo = qv_gather((uint64_t)tid, h->tile_segs, h->n_tile_segs);
    // element tid + THREADS*i of the tile: the default swizzle leaves bits >= 8 alone (slot = swz(tid) + THREADS*i); the wide
    // one also folds them into the column, a compile-time constant per i
    const uint32_t my_slot = qvj_swz(tid);
#if QVJ_WIDE_SWZ
#define QVJ_SLOT(i) ((my_slot ^ (qvj_swz((uint32_t)(i) * THREADS) & 7u)) + (uint32_t)(i) * THREADS)
#else
#define QVJ_SLOT(i) (my_slot + (uint32_t)(i) * THREADS)
#endif
    qvc* const own = PULL ? alt_own : peers.base[(fixed_bits >> n_local) & (QV_MAX_PEERS - 1)];
    // store permutation (trailing X / CNOT / SWAP gates folded into the write-back).  A compile-time switch: as a uniform
    // run-time branch the write-back of the QFT's fourth pass went from 7.6 to 10.5 ms (r2d_summary.md vs r2a_summary.md).
#if QVJ_STORE_PERM
    uint32_t st_lo = h->st_const;
#pragma unroll
    for (uint32_t k = 0; k < 12; k++)
        if (tid >> k & 1) st_lo ^= h->st_col[k];
    st_lo = qvj_from_swz1(st_lo);       // the host constants are in the default layout; the map is XOR-linear
#endif
#if QVJ_HAS_SCALE
    const double out_scale = h->out_scale;
#endif

#if QVJ_TMA_LOAD
    const uint32_t mbar = qvj_smem_u32(&s_mbar);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t tile_no = 0;
#endif

    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const uint64_t base = qv_gather(t, h->base_segs, h->n_base_segs) | fixed_bits;
        const uint64_t pbase = (base | glo) & local_mask;

        // ---- HBM -> shared memory (asynchronous 16-byte copies, all in flight at once)
#if QVJ_SRC_BASIS
        if (((h->basis_index ^ base) & ~h->tile_mask & local_mask) != 0) {      // uniform over the CTA
            char* zdst = reinterpret_cast<char*>(own + pbase);
            qvc z;
            z.x = 0.0;
            z.y = 0.0;
#pragma unroll
            for (int i = 0; i < ITERS; i++) qv_st_stream(reinterpret_cast<qvc*>(zdst + h->hi_byte[i]), z);
            continue;
        }
#pragma unroll
        for (int i = 0; i < ITERS; i++) {
            qvc v;
            v.x = ((pbase | h->hi_off[i]) == (h->basis_index & local_mask)) ? 1.0 : 0.0;
            v.y = 0.0;
            tile[QVJ_SLOT(i)] = v;
        }
#elif QVJ_TMA_LOAD
        if (tid == 0) {
            const uint64_t lbase = base & local_mask;
            int32_t c1 = geom.is_tile[0] ? 0 : (int32_t)((lbase >> geom.start[0]) & ((1ull << geom.len[0]) - 1ull));
            int32_t c2 = geom.is_tile[1] ? 0 : (int32_t)((lbase >> geom.start[1]) & ((1ull << geom.len[1]) - 1ull));
            int32_t c3 = geom.is_tile[2] ? 0 : (int32_t)((lbase >> geom.start[2]) & ((1ull << geom.len[2]) - 1ull));
            int32_t c4 = geom.is_tile[3] ? 0 : (int32_t)((lbase >> geom.start[3]) & ((1ull << geom.len[3]) - 1ull));
            // the previous tile's write-back read this buffer through the generic proxy (it ended with a barrier)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(65536u) : "memory");
            asm volatile(
                "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                    qvj_smem_u32(tile)),
                "l"(&tmap), "r"(0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(mbar)
                : "memory");
        }
#else
        if (!PULL) {
            const char* tsrc = reinterpret_cast<const char*>(own + pbase);
#pragma unroll
            for (int i = 0; i < ITERS; i++) qv_cp_async16(tile + QVJ_SLOT(i), reinterpret_cast<const qvc*>(tsrc + h->hi_byte[i]));
        } else {
            const uint64_t sbase = qv_remap_index(base | glo, h->pull_remap);
            const uint32_t zero_ranks = h->zero_ranks;
#pragma unroll
            for (int i = 0; i < ITERS; i++) {
                const uint64_t p = sbase ^ h->hi_src[i];
                const uint32_t pr = (uint32_t)(p >> n_local) & (QV_MAX_PEERS - 1);
                // a shard known to hold only zeros (right after a reset) is not fetched: the copy zero-fills
                qv_cp_async16_z(tile + QVJ_SLOT(i), peers.base[pr] + (p & local_mask), (zero_ranks >> pr & 1u) ? 0u : 16u);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
#endif

#if QVJ_HAS_TABLES
        {   // per-tile tables, built while the copies fly (same construction as the interpreter kernel)
            const QvSource* sources = reinterpret_cast<const QvSource*>(blob + h->off_sources);
            const QvSlice* slices = reinterpret_cast<const QvSlice*>(blob + h->off_slices);
            const uint8_t* slice_of = blob + h->off_slice_of;
            const QvPred* preds = reinterpret_cast<const QvPred*>(blob + h->off_preds);
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
#endif
#if QVJ_SRC_BASIS
        __syncthreads();
#elif QVJ_TMA_LOAD
        qvj_mbar_wait(mbar, tile_no & 1u);
        tile_no++;
#if QVJ_HAS_TABLES
        __syncthreads();      // the slices were written by other threads
#endif
#else
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
#endif

        // ---- the rounds: straight-line code emitted by the pass compiler
        QVJ_RUN_ROUNDS(tile, tid, blob, tables, s_slice, s_pred)

        // ---- shared memory -> HBM
        {
            char* tdst = reinterpret_cast<char*>(own + pbase);
#pragma unroll
            for (int i = 0; i < ITERS; i++) {
#if QVJ_STORE_PERM
                qvc v = tile[st_lo ^ qvj_from_swz1(h->st_hi[i])];
#else
                qvc v = tile[QVJ_SLOT(i)];
#endif
#if QVJ_HAS_SCALE
                v.x *= out_scale;
                v.y *= out_scale;
#endif
                qv_st_stream(reinterpret_cast<qvc*>(tdst + h->hi_byte[i]), v);
            }
        }
        __syncthreads();
    }
}

#endif  // !QVJ_HOST
