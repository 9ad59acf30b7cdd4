// qv_jit_kernel_tma.cuh -- skeleton of a compiled gate pass, persistent / TMA-fed variant (local passes only).
//
// One CTA per SM (512 threads) walks over its tiles with TWO 64 KiB tile buffers: while the rounds and the write-back of tile k
// run, the tensor-memory accelerator loads tile k+1 into the other buffer (ONE cp.async.bulk.tensor per tile, issued by one
// thread, completion signalled on an mbarrier).  The HBM read phase of a tile therefore overlaps the compute of the previous one
// INSIDE a CTA; in the classic variant (qv_jit_kernel.cuh) the overlap only comes from three independent CTAs per SM.
//
// A tile is a box of a 5-dimensional view of the amplitude vector: dimension 0 = eight amplitudes (16 doubles = 128 bytes),
// the others = the runs of consecutive index bits above bit 2 that are all inside / all outside the tile (QvTmaGeom, built from
// the pass header at launch time: geometry stays data).  The box covers the tile runs completely and sits at the coordinates the
// tile id gives the other runs.  With CU_TENSOR_MAP_SWIZZLE_128B the hardware writes amplitude e of the tile to 16-byte slot
// e ^ ((e >> 3) & 7) -- exactly qv_swz, the layout every round expects.
//
// The generated translation unit defines QVJ_M, QVJ_THREADS (512), QVJ_PROG_BYTES, QVJ_HAS_SCALE, QVJ_STORE_PERM, QVJ_HAS_TABLES,
// the round functions and QVJ_RUN_ROUNDS before including this file.
#pragma once

#if !defined(QVJ_HOST)

struct QvjProg { uint8_t bytes[QVJ_PROG_BYTES]; };
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

extern "C" __global__ void __launch_bounds__(QVJ_THREADS, 1)
qvj_kernel(const __grid_constant__ QvjProg prog, const __grid_constant__ QvPeers peers,
           const qvc* __restrict__ tables, qvc* __restrict__ alt_own,
           const __grid_constant__ QvjTensorMap tmap, const __grid_constant__ QvTmaGeom geom) {
    constexpr int THREADS = QVJ_THREADS;                 // 512: one group of 8 amplitudes per thread and round
    constexpr int PER_THREAD = 4096 / THREADS;           // tile elements a thread writes back
    extern __shared__ uint8_t qv_smem_raw[];
    // SWIZZLE_128B needs the tile buffers on a 1024-byte boundary
    uint8_t* const aligned = qv_smem_raw + ((1024u - (qvj_smem_u32(qv_smem_raw) & 1023u)) & 1023u);
    qvc* const tile0 = reinterpret_cast<qvc*>(aligned);
    qvc* const tile1 = reinterpret_cast<qvc*>(aligned + 65536);
    __shared__ __align__(8) unsigned long long s_mbar[2];
#if QVJ_HAS_TABLES
    __shared__ qvc s_slice[QV_SLICE_ENTRIES];
    __shared__ uint32_t s_srcext[QV_MAX_SOURCES];
    __shared__ uint8_t s_pred[QV_MAX_PREDS];
#else
    const qvc* s_slice = nullptr;
    const uint8_t* s_pred = nullptr;
#endif
    (void)alt_own;
    const uint8_t* blob = prog.bytes;
    const QvPassHeader* h = reinterpret_cast<const QvPassHeader*>(blob);
    const uint64_t fixed_bits = h->fixed_bits;
    const uint64_t n_tiles = h->n_tiles;
    const uint32_t n_local = h->n_local_bits;
    const uint64_t local_mask = (1ull << n_local) - 1ull;
    const uint32_t tid = threadIdx.x;
    // write-back addressing uses the header's tables, which are laid out for 256 threads: thread T handles the elements
    // (T & 255) + 256 * (2 j + (T >> 8)), j = 0 .. 7
    const uint32_t tid8 = tid & 255u, half = tid >> 8;
    const uint64_t glo = qv_gather((uint64_t)tid8, h->tile_segs, h->n_tile_segs);
    const uint32_t my_slot = qv_swz(tid8);
    qvc* const own = peers.base[(fixed_bits >> n_local) & (QV_MAX_PEERS - 1)];
#if QVJ_STORE_PERM
    uint32_t st_lo = h->st_const;
#pragma unroll
    for (uint32_t k = 0; k < 12; k++)
        if (tid8 >> k & 1) st_lo ^= h->st_col[k];
#endif
#if QVJ_HAS_SCALE
    const double out_scale = h->out_scale;
#endif
    const uint32_t mbar0 = qvj_smem_u32(&s_mbar[0]), mbar1 = qvj_smem_u32(&s_mbar[1]);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar0));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // one thread: arm the buffer's mbarrier with the tile's byte count and start the tensor copy
    auto issue_load = [&](uint64_t t, qvc* dst, uint32_t mbar) {
        const uint64_t base = (qv_gather(t, h->base_segs, h->n_base_segs) | fixed_bits) & local_mask;
        int32_t c[5];
        c[0] = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) c[k + 1] = geom.is_tile[k] ? 0 : (int32_t)((base >> geom.start[k]) & ((1ull << geom.len[k]) - 1ull));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(65536u) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                qvj_smem_u32(dst)),
            "l"(&tmap), "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "r"(c[4]), "r"(mbar)
            : "memory");
    };

    if (tid == 0 && (uint64_t)blockIdx.x < n_tiles) issue_load(blockIdx.x, tile0, mbar0);
    uint32_t k = 0;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x, k++) {
        qvc* const tile = (k & 1) ? tile1 : tile0;
        const uint64_t tn = t + gridDim.x;
        if (tid == 0 && tn < n_tiles) {
            // the other buffer was last read (generic proxy) by the write-back of the previous iteration, which ended with
            // a barrier; order those reads before the asynchronous proxy's writes
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_load(tn, (k & 1) ? tile0 : tile1, (k & 1) ? mbar0 : mbar1);
        }
        const uint64_t base = qv_gather(t, h->base_segs, h->n_base_segs) | fixed_bits;
        const uint64_t pbase = (base | glo) & local_mask;
#if QVJ_HAS_TABLES
        {   // per-tile tables, built while the copy flies (same construction as the interpreter kernel)
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
            __syncthreads();
        }
#endif
        qvj_mbar_wait((k & 1) ? mbar1 : mbar0, (k >> 1) & 1u);

        // ---- the rounds: straight-line code emitted by the pass compiler
        QVJ_RUN_ROUNDS(tile, tid, blob, tables, s_slice, s_pred)

        // ---- shared memory -> HBM
        {
            char* tdst = reinterpret_cast<char*>(own + pbase);
#pragma unroll
            for (int j = 0; j < PER_THREAD; j++) {
                const uint32_t i = 2u * (uint32_t)j + half;
#if QVJ_STORE_PERM
                qvc v = tile[st_lo ^ h->st_hi[i]];
#else
                qvc v = tile[my_slot + 256u * i];
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
