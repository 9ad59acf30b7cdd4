// qv_ops.h -- register-level micro-op application shared by the CUDA tile kernel
// (qv_kernels.cuh) and the TEST-ONLY CPU emulator (tests/support/qv_emulator.cpp).
// Everything here works on a GROUP: NS = 2^m amplitudes held in a[0..NS-1], slot r being the
// amplitude whose register bits spell r; g is the group counter (the tile-local index with the
// register bits squeezed out).
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

// ---------------------------------------------------------------------------------------------
// In-place FP64 primitives.  A register group is 16 complex amplitudes = 64 registers that stay live
// across a data-dependent dispatch (one code path per micro-op kind).  If every path produced its
// results in fresh registers, the compiler would have to shuffle up to 64 registers wherever paths
// merge (measured: more register moves than FP64 instructions).  On the device these primitives are
// inline PTX with the destination TIED to a source operand, so an amplitude lives in the same
// register pair on every path and no moves are needed.  On the host (test emulator) they are the
// same expressions in plain C++.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ void qv_mul_ip(double& x, double b) { asm("mul.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b)); }
__device__ __forceinline__ void qv_add_ip(double& x, double b) { asm("add.rn.f64 %0, %0, %1;" : "+d"(x) : "d"(b)); }
__device__ __forceinline__ void qv_fma_acc(double& acc, double a, double b) { asm("fma.rn.f64 %0, %1, %2, %0;" : "+d"(acc) : "d"(a), "d"(b)); }
__device__ __forceinline__ void qv_fnma_acc(double& acc, double a, double b) {
    asm("{\n\t.reg .f64 n;\n\tneg.f64 n, %1;\n\tfma.rn.f64 %0, n, %2, %0;\n\t}" : "+d"(acc) : "d"(a), "d"(b));
}
__device__ __forceinline__ void qv_fma_self(double& x, double b, double c) { asm("fma.rn.f64 %0, %0, %1, %2;" : "+d"(x) : "d"(b), "d"(c)); }
#else
inline void qv_mul_ip(double& x, double b) { x = x * b; }
inline void qv_add_ip(double& x, double b) { x = x + b; }
inline void qv_fma_acc(double& acc, double a, double b) { acc = a * b + acc; }
inline void qv_fnma_acc(double& acc, double a, double b) { acc = acc - a * b; }
inline void qv_fma_self(double& x, double b, double c) { x = x * b + c; }
#endif

// a <- a * t
QV_HD void qv_cmul_ip(qvc& a, qvc t) {
    const double w = a.x * t.y;
    qv_mul_ip(a.x, t.x);
    qv_fnma_acc(a.x, a.y, t.y);
    qv_fma_self(a.y, t.x, w);
}

// a <- a * d + p   (d, p complex; the in-place tail of a matrix row whose other terms are already in p)
QV_HD void qv_cmul_add_ip(qvc& a, qvc d, qvc p) {
    const double w = a.x * d.y;
    qv_fma_self(a.x, d.x, p.x);
    qv_fnma_acc(a.x, d.y, a.y);
    qv_fma_self(a.y, d.x, p.y);
    qv_add_ip(a.y, w);
}

// ---------------------------------------------------------------------------------------------
// Dense micro-ops: psi_group <- M psi_group on one or two register bits (APPLY-1Q/2Q-OPERATOR,
// src/serial-kernels.lisp:174-267).  Every output is the same sum of products as the reference's row
// times column; the terms are associated so that the update happens in place (differences are at the
// level of one rounding per term, far inside the 1e-12 parity tolerance).  CTRL: only the slots whose
// bit is set in slot_ok are touched (controls that sit on register bits).
// ---------------------------------------------------------------------------------------------
template <int NS, int RB, bool REAL, bool CTRL>
QV_HD void qv_dense1(qvc (&a)[NS], const qvc* M, uint32_t slot_ok) {
    const qvc m00 = M[0], m01 = M[1], m10 = M[2], m11 = M[3];
#pragma unroll
    for (int r = 0; r < NS; r++) {
        if (r & (1 << RB)) continue;
        if (CTRL && !((slot_ok >> r) & 1u)) continue;
        const int r1 = r | (1 << RB);
        if (REAL) {
            const double ux = m10.x * a[r].x, uy = m10.x * a[r].y;
            qv_mul_ip(a[r].x, m00.x);
            qv_fma_acc(a[r].x, m01.x, a[r1].x);
            qv_mul_ip(a[r].y, m00.x);
            qv_fma_acc(a[r].y, m01.x, a[r1].y);
            qv_fma_self(a[r1].x, m11.x, ux);
            qv_fma_self(a[r1].y, m11.x, uy);
        } else {
            const qvc p = qv_cmul(m10, a[r]);          // the part of the second row that needs the old a[r]
            const double w = a[r].x * m00.y;
            qv_mul_ip(a[r].x, m00.x);
            qv_fnma_acc(a[r].x, m00.y, a[r].y);
            qv_fma_acc(a[r].x, m01.x, a[r1].x);
            qv_fnma_acc(a[r].x, m01.y, a[r1].y);
            qv_fma_self(a[r].y, m00.x, w);
            qv_fma_acc(a[r].y, m01.x, a[r1].y);
            qv_fma_acc(a[r].y, m01.y, a[r1].x);
            qv_cmul_add_ip(a[r1], m11, p);
        }
    }
}

template <int NS, int RB0, int RB1, bool REAL, bool CTRL>
QV_HD void qv_dense2(qvc (&a)[NS], const qvc* M, uint32_t slot_ok) {
#pragma unroll
    for (int r = 0; r < NS; r++) {
        if (r & ((1 << RB0) | (1 << RB1))) continue;
        if (CTRL && !((slot_ok >> r) & 1u)) continue;
        const int ix[4] = {r, r | (1 << RB0), r | (1 << RB1), r | (1 << RB0) | (1 << RB1)};
        // off-diagonal part of every row first (needs all four old amplitudes), then the diagonal term in place
        qvc p[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const qvc* row = M + 4 * i;
            bool first = true;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (j == i) continue;
                const qvc v = a[ix[j]];
                if (REAL) {
                    if (first) { p[i].x = row[j].x * v.x; p[i].y = row[j].x * v.y; }
                    else { p[i].x = p[i].x + row[j].x * v.x; p[i].y = p[i].y + row[j].x * v.y; }
                } else {
                    p[i] = first ? qv_cmul(row[j], v) : qv_cmadd(p[i], row[j], v);
                }
                first = false;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const qvc d = M[4 * i + i];
            if (REAL) {
                qv_fma_self(a[ix[i]].x, d.x, p[i].x);
                qv_fma_self(a[ix[i]].y, d.x, p[i].y);
            } else {
                qv_cmul_add_ip(a[ix[i]], d, p[i]);
            }
        }
    }
}

template <int NS, bool CTRL>
QV_HD void qv_dense_dispatch(qvc (&a)[NS], uint32_t kind, const qvc* M, uint32_t ok) {
    switch (kind) {
        case QV_K_DENSE1 + 0: qv_dense1<NS, 0, true, CTRL>(a, M, ok); break;
        case QV_K_DENSE1 + 1: qv_dense1<NS, 0, false, CTRL>(a, M, ok); break;
        case QV_K_DENSE1 + 2: if (NS > 2) qv_dense1<NS, (NS > 2 ? 1 : 0), true, CTRL>(a, M, ok); break;
        case QV_K_DENSE1 + 3: if (NS > 2) qv_dense1<NS, (NS > 2 ? 1 : 0), false, CTRL>(a, M, ok); break;
        case QV_K_DENSE1 + 4: if (NS > 4) qv_dense1<NS, (NS > 4 ? 2 : 0), true, CTRL>(a, M, ok); break;
        case QV_K_DENSE1 + 5: if (NS > 4) qv_dense1<NS, (NS > 4 ? 2 : 0), false, CTRL>(a, M, ok); break;
        case QV_K_DENSE1 + 6: if (NS > 8) qv_dense1<NS, (NS > 8 ? 3 : 0), true, CTRL>(a, M, ok); break;
        case QV_K_DENSE1 + 7: if (NS > 8) qv_dense1<NS, (NS > 8 ? 3 : 0), false, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 0: if (NS > 2) qv_dense2<NS, 0, (NS > 2 ? 1 : 1), true, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 1: if (NS > 2) qv_dense2<NS, 0, (NS > 2 ? 1 : 1), false, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 2: if (NS > 4) qv_dense2<NS, 0, (NS > 4 ? 2 : 1), true, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 3: if (NS > 4) qv_dense2<NS, 0, (NS > 4 ? 2 : 1), false, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 4: if (NS > 4) qv_dense2<NS, (NS > 4 ? 1 : 0), (NS > 4 ? 2 : 1), true, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 5: if (NS > 4) qv_dense2<NS, (NS > 4 ? 1 : 0), (NS > 4 ? 2 : 1), false, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 6: if (NS > 8) qv_dense2<NS, 0, (NS > 8 ? 3 : 1), true, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 7: if (NS > 8) qv_dense2<NS, 0, (NS > 8 ? 3 : 1), false, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 8: if (NS > 8) qv_dense2<NS, (NS > 8 ? 1 : 0), (NS > 8 ? 3 : 1), true, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 9: if (NS > 8) qv_dense2<NS, (NS > 8 ? 1 : 0), (NS > 8 ? 3 : 1), false, CTRL>(a, M, ok); break;
        case QV_K_DENSE2 + 10: if (NS > 8) qv_dense2<NS, (NS > 8 ? 2 : 0), (NS > 8 ? 3 : 1), true, CTRL>(a, M, ok); break;
        default: if (NS > 8) qv_dense2<NS, (NS > 8 ? 2 : 0), (NS > 8 ? 3 : 1), false, CTRL>(a, M, ok); break;
    }
}

// ---------------------------------------------------------------------------------------------
// Diagonal micro-ops: every amplitude is multiplied by one table entry.  The table index is
// idx (the part that depends on the group and the tile) | slot_off[r] (the register bits).
//   gated by RB: the table only holds the entries with that bit SET (every entry with the bit
//   clear is exactly 1: controlled-phase ladders), so only half of the slots are touched.
// ---------------------------------------------------------------------------------------------
template <int NS, int RB>
QV_HD void qv_diag_gated1(qvc (&a)[NS], qvc t) {
#pragma unroll
    for (int r = 0; r < NS; r++)
        if (r & (1 << RB)) qv_cmul_ip(a[r], t);
}

template <int NS, int RB>
QV_HD void qv_diag_gatedn(qvc (&a)[NS], const qvc* tab, uint32_t idx, const QvUop& u) {
#pragma unroll
    for (int r = 0; r < NS; r++)
        if (r & (1 << RB)) qv_cmul_ip(a[r], tab[idx | u.slot_off[r]]);
}

template <int NS, int RB>
QV_HD void qv_diag_onebit(qvc (&a)[NS], const qvc* tab, uint32_t idx, const QvUop& u) {
    const qvc t0 = tab[idx], t1 = tab[idx | u.slot_off[1 << RB]];
#pragma unroll
    for (int r = 0; r < NS; r++) qv_cmul_ip(a[r], (r & (1 << RB)) ? t1 : t0);
}

template <int NS>
QV_HD void qv_diag_dispatch(qvc (&a)[NS], uint32_t kind, const qvc* tab, uint32_t idx, const QvUop& u, const qvc* slices) {
    if (kind <= QV_K_DIAG_GATED1 + 3) {
        // one lookup for the whole group, optionally times a per-tile scalar
        qvc t = tab[idx];
        if (u.flags & QV_UF_SCALE) t = qv_cmul(t, slices[u.scale]);
        switch (kind) {
            case QV_K_DIAG_COMMON: {
#pragma unroll
                for (int r = 0; r < NS; r++) qv_cmul_ip(a[r], t);
                break;
            }
            case QV_K_DIAG_GATED1 + 0: qv_diag_gated1<NS, 0>(a, t); break;
            case QV_K_DIAG_GATED1 + 1: if (NS > 2) qv_diag_gated1<NS, (NS > 2 ? 1 : 0)>(a, t); break;
            case QV_K_DIAG_GATED1 + 2: if (NS > 4) qv_diag_gated1<NS, (NS > 4 ? 2 : 0)>(a, t); break;
            default: if (NS > 8) qv_diag_gated1<NS, (NS > 8 ? 3 : 0)>(a, t); break;
        }
        return;
    }
    switch (kind) {
        case QV_K_DIAG_GATEDN + 0: qv_diag_gatedn<NS, 0>(a, tab, idx, u); break;
        case QV_K_DIAG_GATEDN + 1: if (NS > 2) qv_diag_gatedn<NS, (NS > 2 ? 1 : 0)>(a, tab, idx, u); break;
        case QV_K_DIAG_GATEDN + 2: if (NS > 4) qv_diag_gatedn<NS, (NS > 4 ? 2 : 0)>(a, tab, idx, u); break;
        case QV_K_DIAG_GATEDN + 3: if (NS > 8) qv_diag_gatedn<NS, (NS > 8 ? 3 : 0)>(a, tab, idx, u); break;
        case QV_K_DIAG_ONEBIT + 0: qv_diag_onebit<NS, 0>(a, tab, idx, u); break;
        case QV_K_DIAG_ONEBIT + 1: if (NS > 2) qv_diag_onebit<NS, (NS > 2 ? 1 : 0)>(a, tab, idx, u); break;
        case QV_K_DIAG_ONEBIT + 2: if (NS > 4) qv_diag_onebit<NS, (NS > 4 ? 2 : 0)>(a, tab, idx, u); break;
        case QV_K_DIAG_ONEBIT + 3: if (NS > 8) qv_diag_onebit<NS, (NS > 8 ? 3 : 0)>(a, tab, idx, u); break;
        default: {
#pragma unroll
            for (int r = 0; r < NS; r++) qv_cmul_ip(a[r], tab[idx | u.slot_off[r]]);
            break;
        }
    }
}

// Apply one micro-op to a register group.
//   tables : the pass's global-memory table pool        slices : the per-tile slice area
//   s_ext  : per-tile external index parts (QvExt)      s_pred : per-tile control predicates (QvPred)
template <int NS>
QV_HD void qv_run_uop(qvc (&a)[NS], const QvUop& u, uint32_t g, const uint8_t* blob, const qvc* tables,
                      const qvc* slices, const uint32_t* s_ext, const uint8_t* s_pred) {
    const uint32_t kind = u.kind;
    const uint32_t flags = u.flags;
    if (kind < QV_K_DIAG_COMMON) {
        if ((flags & QV_UF_PRED) && !s_pred[u.pred]) return;
        const qvc* M = reinterpret_cast<const qvc*>(blob + u.data);
        if (flags & QV_UF_CTRL) {
            if ((g & u.cm) != u.cv) return;
            qv_dense_dispatch<NS, true>(a, kind, M, u.slot_ok);
        } else {
            qv_dense_dispatch<NS, false>(a, kind, M, 0xffffu);
        }
    } else {
        uint32_t idx;
        if (flags & QV_UF_GENERIC) {
            const QvSegList* sl = reinterpret_cast<const QvSegList*>(blob + u.segs);
            idx = qv_gather32(g, sl->segs, sl->n);
        } else {
            idx = (g >> (u.cm & 0xffu)) & (u.cm >> 8);
            if (flags & QV_UF_FIELD2) idx |= (g >> (u.cv & 0xffu)) & (u.cv >> 8);
        }
        if (flags & QV_UF_EXT) idx |= s_ext[u.ext];
        const qvc* tab = ((flags & QV_UF_SLICE) ? slices : tables) + u.data;
        qv_diag_dispatch<NS>(a, kind, tab, idx, u, slices);
    }
}

// Entry x of a slice for the tile at hand: the product of its sources (src_ext[s] = the source's
// external index part, already shifted left by its nl).
QV_HD qvc qv_slice_entry(const QvSlice& sl, const QvSource* sources, const uint32_t* src_ext, const qvc* tables, uint32_t x) {
    qvc prod;
    prod.x = 1.0;
    prod.y = 0.0;
    for (uint32_t s = 0; s < sl.n_src; s++) {
        const QvSource& src = sources[sl.first_src + s];
        const qvc t = tables[src.table_off + (src_ext[sl.first_src + s] | qv_gather32(x, src.lsegs, src.n_lsegs))];
        prod = s ? qv_cmul(prod, t) : t;
    }
    return prod;
}
