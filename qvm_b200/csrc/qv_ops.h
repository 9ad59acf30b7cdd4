// qv_ops.h -- register-level op application shared by the CUDA tile kernel
// (qv_kernels.cu) and the TEST-ONLY CPU emulator (tests/support/qv_emulator.cpp).
// Everything here works on a GROUP: 2^m amplitudes held in a[0..7], slot r being
// the amplitude at tile-local index e0 | dep[r].
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
// between lanes) still hit 8 distinct columns per quarter-warp.
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

// tile-local offset of register slot s for register-bit positions p0 < p1 < p2
QV_HD uint32_t qv_dep(int s, uint32_t p0, uint32_t p1, uint32_t p2) {
    return ((s & 1) ? (1u << p0) : 0u) | ((s & 2) ? (1u << p1) : 0u) | ((s & 4) ? (1u << p2) : 0u);
}

struct QvRegPos {
    uint32_t p0, p1, p2;
};

template <int RB>
QV_HD void qv_dense1(qvc a[8], const qvc* M, uint32_t flags, uint32_t e0, const QvRegPos dep,
                     uint32_t cm, uint32_t cv) {
    const qvc m00 = M[0], m01 = M[1], m10 = M[2], m11 = M[3];
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r & (1 << RB)) continue;
        const int r1 = r | (1 << RB);
        const qvc a0 = a[r], a1 = a[r1];
        qvc n0, n1;
        if (flags & QV_F_REAL) {
            n0.x = m00.x * a0.x + m01.x * a1.x;
            n0.y = m00.x * a0.y + m01.x * a1.y;
            n1.x = m10.x * a0.x + m11.x * a1.x;
            n1.y = m10.x * a0.y + m11.x * a1.y;
        } else {
            n0 = qv_cmadd(qv_cmul(m00, a0), m01, a1);
            n1 = qv_cmadd(qv_cmul(m10, a0), m11, a1);
        }
        bool ok = true;
        if (flags & QV_F_CTRL_LOCAL) ok = (((e0 | qv_dep(r, dep.p0, dep.p1, dep.p2)) & cm) == cv);
        if (ok) {
            a[r] = n0;
            a[r1] = n1;
        }
    }
}

template <int RB0, int RB1>
QV_HD void qv_dense2(qvc a[8], const qvc* M, uint32_t flags, uint32_t e0, const QvRegPos dep,
                     uint32_t cm, uint32_t cv) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
        if (r & ((1 << RB0) | (1 << RB1))) continue;
        const int i0 = r, i1 = r | (1 << RB0), i2 = r | (1 << RB1), i3 = r | (1 << RB0) | (1 << RB1);
        const qvc v0 = a[i0], v1 = a[i1], v2 = a[i2], v3 = a[i3];
        qvc o[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const qvc* row = M + 4 * i;
            if (flags & QV_F_REAL) {
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
        if (flags & QV_F_CTRL_LOCAL) ok = (((e0 | qv_dep(r, dep.p0, dep.p1, dep.p2)) & cm) == cv);
        if (ok) {
            a[i0] = o[0];
            a[i1] = o[1];
            a[i2] = o[2];
            a[i3] = o[3];
        }
    }
}

// Merged diagonal: every amplitude is multiplied by the product of its chunk
// table entries.  Chunks that do not read a register bit give one factor for
// the whole group, folded into `common` (one complex multiply per group instead
// of one per amplitude).
QV_HD void qv_diag(qvc a[8], const QvOp& op, const QvChunk* chunks, const qvc* tables, uint32_t e0,
                   const QvRegPos dep, uint64_t base) {
    qvc common;
    common.x = 1.0;
    common.y = 0.0;
    bool have_common = false;
    for (uint32_t c = 0; c < op.n_chunks; c++) {
        const QvChunk& ch = chunks[op.data_off + c];
        const uint32_t g0 = (uint32_t)qv_gather(base, ch.esegs, ch.n_esegs) | qv_gather32(e0, ch.lsegs, ch.n_lsegs);
        const qvc* tab = tables + ch.table_off;
        if (ch.reg_mask == 0) {
            common = have_common ? qv_cmul(common, tab[g0]) : tab[g0];
            have_common = true;
        } else {
#pragma unroll
            for (int r = 0; r < 8; r++) {
                const uint32_t g = g0 | qv_gather32(qv_dep(r, dep.p0, dep.p1, dep.p2), ch.lsegs, ch.n_lsegs);
                a[r] = qv_cmul(a[r], tab[g]);
            }
        }
    }
    if (have_common) {
#pragma unroll
        for (int r = 0; r < 8; r++) a[r] = qv_cmul(a[r], common);
    }
}

// Apply every op of a round to one register group.
QV_HD void qv_apply_round(qvc a[8], const QvRound& rd, const QvOp* ops, const QvChunk* chunks,
                          const qvc* mats, const qvc* tables, uint32_t e0, const QvRegPos dep,
                          uint64_t base) {
    for (uint32_t i = 0; i < rd.n_ops; i++) {
        const QvOp& op = ops[rd.first_op + i];
        if ((op.flags & QV_F_CTRL_EXT) && ((base & op.cm_ext) != op.cv_ext)) continue;
        if (op.type == QV_OP_DENSE1) {
            const qvc* M = mats + op.data_off;
            switch (op.rb0) {
                case 0: qv_dense1<0>(a, M, op.flags, e0, dep, op.cm_local, op.cv_local); break;
                case 1: qv_dense1<1>(a, M, op.flags, e0, dep, op.cm_local, op.cv_local); break;
                default: qv_dense1<2>(a, M, op.flags, e0, dep, op.cm_local, op.cv_local); break;
            }
        } else if (op.type == QV_OP_DENSE2) {
            const qvc* M = mats + op.data_off;
            const uint32_t sel = op.rb0 + op.rb1;   // (0,1)->1 (0,2)->2 (1,2)->3
            switch (sel) {
                case 1: qv_dense2<0, 1>(a, M, op.flags, e0, dep, op.cm_local, op.cv_local); break;
                case 2: qv_dense2<0, 2>(a, M, op.flags, e0, dep, op.cm_local, op.cv_local); break;
                default: qv_dense2<1, 2>(a, M, op.flags, e0, dep, op.cm_local, op.cv_local); break;
            }
        } else {
            qv_diag(a, op, chunks, tables, e0, dep, base);
        }
    }
}
