// qv_ops.h -- register-level op application shared by the CUDA tile kernel
// (qv_kernels.cuh) and the TEST-ONLY CPU emulator (tests/support/qv_emulator.cpp).
// Everything here works on a GROUP: 2^m amplitudes held in a[0..7], slot r being
// the amplitude at tile-local index e0 | rd.slot_dep[r].
#pragma once
#include "qv_program.h"

#if defined(__CUDACC__)
#define QV_HD __host__ __device__ __forceinline__
#else
#define QV_HD inline
#endif

struct alignas(16) qvc {
    double x, y;
};

// Shared-memory swizzle: XOR the 16-byte column inside a 128-byte row with the
// row number so that groups whose register bits are the low bits (stride 128 B
// between lanes) still hit 8 distinct columns per quarter-warp.  It is XOR-linear:
// qv_swz(a ^ b) == qv_swz(a) ^ qv_swz(b), which the kernel uses to turn per-slot
// address math into one XOR with a host-precomputed constant.
QV_HD uint32_t qv_swz(uint32_t e) { return e ^ ((e >> 3) & 7u); }

QV_HD uint64_t qv_gather(uint64_t x, const QvSeg* segs, uint32_t n) {
    uint64_t r = 0;
    for (uint32_t i = 0; i < n; i++)
        r |= ((x >> segs[i].src) & ((1ull << segs[i].len) - 1ull)) << segs[i].dst;
    return r;
}

QV_HD uint32_t qv_gather32(uint32_t x, const QvSeg* segs, uint32_t n) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < n; i++)
        r |= ((x >> segs[i].src) & ((1u << segs[i].len) - 1u)) << segs[i].dst;
    return r;
}

QV_HD uint32_t qv_insert_zero(uint32_t g, uint32_t pos) {
    uint32_t lo = g & ((1u << pos) - 1u);
    return ((g >> pos) << (pos + 1)) | lo;
}

QV_HD qvc qv_cmul(qvc a, qvc b) {
    qvc r;
    r.x = a.x * b.x - a.y * b.y;
    r.y = a.x * b.y + a.y * b.x;
    return r;
}

// acc + m*a, accumulated in the reference's left-to-right order
// (src/linear-algebra.lisp:53-72, src/serial-kernels.lisp:60-66).
QV_HD qvc qv_cmadd(qvc acc, qvc m, qvc a) {
    qvc r;
    r.x = acc.x + (m.x * a.x - m.y * a.y);
    r.y = acc.y + (m.x * a.y + m.y * a.x);
    return r;
}

template <int RB, bool REAL, bool CTRL>
QV_HD void qv_dense1(qvc a[8], const qvc* M, const QvRound& rd, uint32_t e0, uint32_t cm, uint32_t cv) {
    const qvc m00 = M[0], m01 = M[1], m10 = M[2], m11 = M[3];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r & (1 << RB)) continue;
        const int r1 = r | (1 << RB);
        const qvc a0 = a[r], a1 = a[r1];
        qvc n0, n1;
        if (REAL) {
            n0.x = m00.x * a0.x + m01.x * a1.x;
            n0.y = m00.x * a0.y + m01.x * a1.y;
            n1.x = m10.x * a0.x + m11.x * a1.x;
            n1.y = m10.x * a0.y + m11.x * a1.y;
        } else {
            n0 = qv_cmadd(qv_cmul(m00, a0), m01, a1);
            n1 = qv_cmadd(qv_cmul(m10, a0), m11, a1);
        }
        bool ok = true;
        if (CTRL) ok = (((e0 | rd.slot_dep[r]) & cm) == cv);
        if (ok) {
            a[r] = n0;
            a[r1] = n1;
        }
    }
}

template <int RB0, int RB1, bool REAL, bool CTRL>
QV_HD void qv_dense2(qvc a[8], const qvc* M, const QvRound& rd, uint32_t e0, uint32_t cm, uint32_t cv) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r & ((1 << RB0) | (1 << RB1))) continue;
        const int i0 = r, i1 = r | (1 << RB0), i2 = r | (1 << RB1), i3 = r | (1 << RB0) | (1 << RB1);
        const qvc v0 = a[i0], v1 = a[i1], v2 = a[i2], v3 = a[i3];
        qvc o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const qvc* row = M + 4 * i;
            if (REAL) {
                o[i].x = ((row[0].x * v0.x + row[1].x * v1.x) + row[2].x * v2.x) + row[3].x * v3.x;
                o[i].y = ((row[0].x * v0.y + row[1].x * v1.y) + row[2].x * v2.y) + row[3].x * v3.y;
            } else {
                qvc acc = qv_cmul(row[0], v0);
                acc = qv_cmadd(acc, row[1], v1);
                acc = qv_cmadd(acc, row[2], v2);
                acc = qv_cmadd(acc, row[3], v3);
                o[i] = acc;
            }
        }
        bool ok = true;
        if (CTRL) ok = (((e0 | rd.slot_dep[r]) & cm) == cv);
        if (ok) {
            a[i0] = o[0];
            a[i1] = o[1];
            a[i2] = o[2];
            a[i3] = o[3];
        }
    }
}

template <int RB, bool REAL>
QV_HD void qv_dense1_dispatch(qvc a[8], const qvc* M, const QvRound& rd, const QvOp& op, uint32_t e0) {
    if (op.flags & QV_F_CTRL_LOCAL) qv_dense1<RB, REAL, true>(a, M, rd, e0, op.cm_local, op.cv_local);
    else qv_dense1<RB, REAL, false>(a, M, rd, e0, 0, 0);
}

template <int RB0, int RB1>
QV_HD void qv_dense2_dispatch(qvc a[8], const qvc* M, const QvRound& rd, const QvOp& op, uint32_t e0) {
    const bool ctrl = (op.flags & QV_F_CTRL_LOCAL) != 0;
    if (op.flags & QV_F_REAL) {
        if (ctrl) qv_dense2<RB0, RB1, true, true>(a, M, rd, e0, op.cm_local, op.cv_local);
        else qv_dense2<RB0, RB1, true, false>(a, M, rd, e0, 0, 0);
    } else {
        if (ctrl) qv_dense2<RB0, RB1, false, true>(a, M, rd, e0, op.cm_local, op.cv_local);
        else qv_dense2<RB0, RB1, false, false>(a, M, rd, e0, 0, 0);
    }
}

// Table lookups of one chunk for the 8 slots of a group.  TAB is either a global-memory table or a
// shared-memory slice (separate instantiations keep the address spaces explicit for the compiler).
//   gated by register bit RB: every entry with that bit clear is exactly 1, so only the slots with the
//   bit set are touched (controlled-phase ladders: half the multiplies and lookups).
template <int RB, typename TAB>
QV_HD void qv_diag_gated(qvc a[8], TAB tab, uint32_t g0, const QvChunk& ch) {
    if ((ch.reg_mask & (ch.reg_mask - 1)) == 0) {
        const qvc t1 = tab[g0 | ch.slot_off[1 << RB]];
#pragma unroll
        for (int r = 0; r < 8; r++)
            if (r & (1 << RB)) a[r] = qv_cmul(a[r], t1);
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++)
            if (r & (1 << RB)) a[r] = qv_cmul(a[r], tab[g0 | ch.slot_off[r]]);
    }
}

// One register bit feeds the chunk: two table entries serve the whole group.
template <int RB>
QV_HD void qv_diag_1bit(qvc a[8], qvc t0, qvc t1) {
#pragma unroll
    for (int r = 0; r < 8; r++) a[r] = qv_cmul(a[r], (r & (1 << RB)) ? t1 : t0);
}

template <typename TAB>
QV_HD void qv_diag_chunk(qvc a[8], TAB tab, uint32_t g0, const QvChunk& ch, qvc& common, bool& have_common) {
    const uint32_t rm = ch.reg_mask;
    if (rm == 0) {
        // no register bit: one factor for the whole group
        common = have_common ? qv_cmul(common, tab[g0]) : tab[g0];
        have_common = true;
    } else if (ch.gate_rb) {
        if (ch.gate_rb == 1) qv_diag_gated<0>(a, tab, g0, ch);
        else if (ch.gate_rb == 2) qv_diag_gated<1>(a, tab, g0, ch);
        else qv_diag_gated<2>(a, tab, g0, ch);
    } else if ((rm & (rm - 1)) == 0) {
        const qvc t0 = tab[g0];
        if (rm == 1) qv_diag_1bit<0>(a, t0, tab[g0 | ch.slot_off[1]]);
        else if (rm == 2) qv_diag_1bit<1>(a, t0, tab[g0 | ch.slot_off[2]]);
        else qv_diag_1bit<2>(a, t0, tab[g0 | ch.slot_off[4]]);
    } else {
#pragma unroll
        for (int r = 0; r < 8; r++) a[r] = qv_cmul(a[r], tab[g0 | ch.slot_off[r]]);
    }
}

// Merged diagonal: every amplitude is multiplied by the product of its chunk table entries.
//   chunk_ext[c] : the part of a GLOBAL chunk's table index that comes from the bits outside the tile
//                  (constant per tile, computed once per tile);
//   slices       : the per-tile slice area (SLICE chunks), built once per tile by qv_build_slices.
QV_HD void qv_diag(qvc a[8], const QvOp& op, const QvChunk* chunks, const qvc* tables, const uint32_t* chunk_ext,
                   const qvc* slices, uint32_t e0) {
    qvc common;
    common.x = 1.0;
    common.y = 0.0;
    bool have_common = false;
    for (uint32_t c = 0; c < op.n_chunks; c++) {
        const uint32_t ci = op.data_off + c;
        const QvChunk& ch = chunks[ci];
        const uint32_t gl = qv_gather32(e0, ch.lsegs, ch.n_lsegs);
        if (ch.kind) qv_diag_chunk(a, slices + ch.table_off, gl, ch, common, have_common);
        else qv_diag_chunk(a, tables + ch.table_off, chunk_ext[ci] | gl, ch, common, have_common);
    }
    if (have_common) {
#pragma unroll
        for (int r = 0; r < 8; r++) a[r] = qv_cmul(a[r], common);
    }
}

// Entry x of SLICE chunk ch for the tile whose base index is `base`: the product of its sources.
QV_HD qvc qv_slice_entry(const QvChunk& ch, const QvSource* sources, const qvc* tables, uint64_t base, uint32_t x) {
    qvc prod;
    prod.x = 1.0;
    prod.y = 0.0;
    for (uint32_t s = 0; s < ch.n_src; s++) {
        const QvSource& src = sources[ch.first_src + s];
        const uint32_t ext = (uint32_t)qv_gather(base, src.esegs, src.n_esegs);
        const qvc t = tables[src.table_off + ((ext << ch.nl) | x)];
        prod = s ? qv_cmul(prod, t) : t;
    }
    return prod;
}

// Apply every op of a round to one register group.
QV_HD void qv_apply_round(qvc a[8], const QvRound& rd, const QvOp* ops, const QvChunk* chunks,
                          const qvc* mats, const qvc* tables, const uint32_t* chunk_ext, const qvc* slices,
                          uint32_t e0, uint64_t base) {
    for (uint32_t i = 0; i < rd.n_ops; i++) {
        const QvOp& op = ops[rd.first_op + i];
        if ((op.flags & QV_F_CTRL_EXT) && ((base & op.cm_ext) != op.cv_ext)) continue;
        if (op.type == QV_OP_DENSE1) {
            const qvc* M = mats + op.data_off;
            if (op.flags & QV_F_REAL) {
                switch (op.rb0) {
                    case 0: qv_dense1_dispatch<0, true>(a, M, rd, op, e0); break;
                    case 1: qv_dense1_dispatch<1, true>(a, M, rd, op, e0); break;
                    default: qv_dense1_dispatch<2, true>(a, M, rd, op, e0); break;
                }
            } else {
                switch (op.rb0) {
                    case 0: qv_dense1_dispatch<0, false>(a, M, rd, op, e0); break;
                    case 1: qv_dense1_dispatch<1, false>(a, M, rd, op, e0); break;
                    default: qv_dense1_dispatch<2, false>(a, M, rd, op, e0); break;
                }
            }
        } else if (op.type == QV_OP_DENSE2) {
            const qvc* M = mats + op.data_off;
            const uint32_t sel = op.rb0 + op.rb1;   // (0,1)->1 (0,2)->2 (1,2)->3
            switch (sel) {
                case 1: qv_dense2_dispatch<0, 1>(a, M, rd, op, e0); break;
                case 2: qv_dense2_dispatch<0, 2>(a, M, rd, op, e0); break;
                default: qv_dense2_dispatch<1, 2>(a, M, rd, op, e0); break;
            }
        } else {
            qv_diag(a, op, chunks, tables, chunk_ext, slices, e0);
        }
    }
}
