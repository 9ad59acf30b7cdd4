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

// 1 in the translation units the pass compiler generates (qv_jit_prelude.cuh): dense micro-ops use plain sums over
// temporaries instead of the in-place forms the interpreter kernel needs (see "In-place FP64 primitives" below)
#if !defined(QV_NATURAL_DENSE)
#define QV_NATURAL_DENSE 0
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
        } else if (QV_NATURAL_DENSE) {
            // compiled passes: plain row-times-column sums, 1 multiply + 3 fused multiply-adds per real output (16 FP64
            // instructions per pair, the minimum for a general complex 2x2); the copies cost no FP64-pipe slot
            const double a0x = a[r].x, a0y = a[r].y, a1x = a[r1].x, a1y = a[r1].y;
            double x0 = m00.x * a0x, y0 = m00.x * a0y, x1 = m10.x * a0x, y1 = m10.x * a0y;
            qv_fnma_acc(x0, m00.y, a0y);
            qv_fma_acc(y0, m00.y, a0x);
            qv_fnma_acc(x1, m10.y, a0y);
            qv_fma_acc(y1, m10.y, a0x);
            qv_fma_acc(x0, m01.x, a1x);
            qv_fma_acc(y0, m01.x, a1y);
            qv_fma_acc(x1, m11.x, a1x);
            qv_fma_acc(y1, m11.x, a1y);
            qv_fnma_acc(x0, m01.y, a1y);
            qv_fma_acc(y0, m01.y, a1x);
            qv_fnma_acc(x1, m11.y, a1y);
            qv_fma_acc(y1, m11.y, a1x);
            a[r].x = x0;
            a[r].y = y0;
            a[r1].x = x1;
            a[r1].y = y1;
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

// The unscaled butterfly (a0, a1) <- (a0 + a1, a0 - a1) on register bit RB, in place and without temporaries:
// a0 += a1, then a1 = a0 - 2 a1.
template <int NS, int RB>
QV_HD void qv_bfly(qvc (&a)[NS]) {
#pragma unroll
    for (int r = 0; r < NS; r++) {
        if (r & (1 << RB)) continue;
        const int r1 = r | (1 << RB);
        qv_add_ip(a[r].x, a[r1].x);
        qv_add_ip(a[r].y, a[r1].y);
        qv_fma_self(a[r1].x, -2.0, a[r].x);
        qv_fma_self(a[r1].y, -2.0, a[r].y);
    }
}

template <int NS, int RB0, int RB1, bool REAL, bool CTRL>
QV_HD void qv_dense2(qvc (&a)[NS], const qvc* M, uint32_t slot_ok) {
#pragma unroll
    for (int r = 0; r < NS; r++) {
        if (r & ((1 << RB0) | (1 << RB1))) continue;
        if (CTRL && !((slot_ok >> r) & 1u)) continue;
        const int ix[4] = {r, r | (1 << RB0), r | (1 << RB1), r | (1 << RB0) | (1 << RB1)};
        if (QV_NATURAL_DENSE && !REAL) {
            // compiled passes: 16 FP64 instructions per output amplitude (1 multiply + 7 fused multiply-adds per real part)
            const qvc v0 = a[ix[0]], v1 = a[ix[1]], v2 = a[ix[2]], v3 = a[ix[3]];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const qvc* row = M + 4 * i;
                double x = row[0].x * v0.x, y = row[0].x * v0.y;
                qv_fnma_acc(x, row[0].y, v0.y);
                qv_fma_acc(y, row[0].y, v0.x);
                qv_fma_acc(x, row[1].x, v1.x);
                qv_fma_acc(y, row[1].x, v1.y);
                qv_fnma_acc(x, row[1].y, v1.y);
                qv_fma_acc(y, row[1].y, v1.x);
                qv_fma_acc(x, row[2].x, v2.x);
                qv_fma_acc(y, row[2].x, v2.y);
                qv_fnma_acc(x, row[2].y, v2.y);
                qv_fma_acc(y, row[2].y, v2.x);
                qv_fma_acc(x, row[3].x, v3.x);
                qv_fma_acc(y, row[3].x, v3.y);
                qv_fnma_acc(x, row[3].y, v3.y);
                qv_fma_acc(y, row[3].y, v3.x);
                a[ix[i]].x = x;
                a[ix[i]].y = y;
            }
            continue;
        }
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
// Diagonal micro-ops: every amplitude (or every amplitude whose gate bit is set) is multiplied by one
// table entry.
// ---------------------------------------------------------------------------------------------
// slot number r with bit RB squeezed out
template <int RB>
QV_HD constexpr int qv_squeeze(int r) { return ((r >> (RB + 1)) << RB) | (r & ((1 << RB) - 1)); }

// GATE = 0: all slots; GATE = 1 + RB: the slots with register bit RB set
template <int NS, int GATE>
QV_HD void qv_diag1(qvc (&a)[NS], qvc t) {
#pragma unroll
    for (int r = 0; r < NS; r++)
        if (GATE == 0 || (r & (1 << (GATE > 0 ? GATE - 1 : 0)))) qv_cmul_ip(a[r], t);
}

// tab points at the group's run of per-slot entries
template <int NS, int GATE>
QV_HD void qv_diagr(qvc (&a)[NS], const qvc* tab) {
#pragma unroll
    for (int r = 0; r < NS; r++) {
        if (GATE == 0) qv_cmul_ip(a[r], tab[r]);
        else if (r & (1 << (GATE > 0 ? GATE - 1 : 0))) qv_cmul_ip(a[r], tab[qv_squeeze<(GATE > 0 ? GATE - 1 : 0)>(r)]);
    }
}

// log2 of the per-group run of a DIAGR table
template <int NS>
QV_HD constexpr uint32_t qv_slot_field(uint32_t gate) {
    return (NS == 16 ? 4u : 3u) - (gate ? 1u : 0u);
}

// The first 16 bytes of a micro-op (kind | flags << 8 | pred << 16, data, cm, cv): the kernel fetches the
// NEXT micro-op's header while the current one runs, so that the decode latency is off the critical path.
struct alignas(16) QvUopHead {
    uint32_t w0, data, cm, cv;
};

// Rare variants (controls, predicates, scattered index fields): decoded step by step.
template <int NS>
QV_HD void qv_run_uop_slow(qvc (&a)[NS], const QvUopHead& h, const QvUop& u, uint32_t g, const uint8_t* blob,
                           const qvc* tables, const qvc* slices, const uint8_t* s_pred) {
    const uint32_t kind = h.w0 & 0xffu, flags = (h.w0 >> 8) & 0xffu;
    if (kind < QV_K_DIAG_BASE) {
        const qvc* M = reinterpret_cast<const qvc*>(blob + h.data);
        if ((flags & QV_UF_PRED) && !s_pred[(h.w0 >> 16) & 0xffu]) return;
        if (flags & QV_UF_CTRL) {
            if ((g & h.cm) != h.cv) return;
            qv_dense_dispatch<NS, true>(a, kind, M, u.slot_ok);
        } else {
            qv_dense_dispatch<NS, false>(a, kind, M, 0xffffu);
        }
        return;
    }
    // diagonal with a scattered index (QV_UF_GENERIC)
    const QvSegList* sl = reinterpret_cast<const QvSegList*>(blob + u.segs);
    const uint32_t idx = qv_gather32(g, sl->segs, sl->n);
    const uint32_t gate = (kind - QV_K_DIAG_BASE) % 5u, space = (kind - QV_K_DIAG_BASE) / 5u;
    qvc t;
    const qvc* tab = nullptr;
    if (space == 0) t = slices[h.data + idx];
    else if (space == 1) t = tables[h.data + idx];
    else if (space == 2) tab = slices + h.data + (idx << qv_slot_field<NS>(gate));
    else if (space == 3) tab = tables + h.data + (idx << qv_slot_field<NS>(gate));
    else tab = reinterpret_cast<const qvc*>(blob + h.data);
    if (space < 2 && (flags & QV_UF_SCALE)) t = qv_cmul(t, slices[u.scale]);
    switch (gate) {
        case 0: if (space < 2) qv_diag1<NS, 0>(a, t); else qv_diagr<NS, 0>(a, tab); break;
        case 1: if (space < 2) qv_diag1<NS, 1>(a, t); else qv_diagr<NS, 1>(a, tab); break;
        case 2: if (space < 2) qv_diag1<NS, 2>(a, t); else qv_diagr<NS, 2>(a, tab); break;
        case 3: if (space < 2) qv_diag1<NS, 3>(a, t); else qv_diagr<NS, 3>(a, tab); break;
        default:
            if (NS > 8) {
                if (space < 2) qv_diag1<NS, (NS > 8 ? 4 : 0)>(a, t);
                else qv_diagr<NS, (NS > 8 ? 4 : 0)>(a, tab);
            }
            break;
    }
}

// Apply one micro-op to a register group: ONE flat jump on the kind, every operand already resolved.
//   tables : the pass's global-memory table pool        slices : the per-tile slice area (shared memory)
//   s_pred : per-tile control predicates (QvPred)
template <int NS>
QV_HD void qv_run_uop(qvc (&a)[NS], const QvUopHead& h, const QvUop& u, uint32_t g, const uint8_t* blob,
                      const qvc* tables, const qvc* slices, const uint8_t* s_pred) {
    const uint32_t kind = h.w0 & 0xffu;
    if (h.w0 & ((QV_UF_CTRL | QV_UF_PRED | QV_UF_GENERIC) << 8)) {
        qv_run_uop_slow<NS>(a, h, u, g, blob, tables, slices, s_pred);
        return;
    }
    const qvc* M = reinterpret_cast<const qvc*>(blob + h.data);
    // table index of a diagonal micro-op: two fields of the group counter (evaluated in the diagonal cases only)
#define QV_IDX (((g >> (h.cm & 0xffu)) & (h.cm >> 8)) | ((g >> (h.cv & 0xffu)) & (h.cv >> 8)))
#define QV_D1(RB) \
    case QV_K_DENSE1 + 2 * RB: qv_dense1<NS, RB, true, false>(a, M, 0xffffu); break; \
    case QV_K_DENSE1 + 2 * RB + 1: qv_dense1<NS, RB, false, false>(a, M, 0xffffu); break;
#define QV_D2(P, RB0, RB1) \
    case QV_K_DENSE2 + 2 * P: qv_dense2<NS, RB0, RB1, true, false>(a, M, 0xffffu); break; \
    case QV_K_DENSE2 + 2 * P + 1: qv_dense2<NS, RB0, RB1, false, false>(a, M, 0xffffu); break;
#define QV_BD(LBL, RB) \
    case QV_K_BFLY_DIAG1_S + LBL: { \
        qv_bfly<NS, RB>(a); \
        qvc t = slices[h.data + QV_IDX]; \
        if (h.w0 & (QV_UF_SCALE << 8)) t = qv_cmul(t, slices[u.scale]); \
        qv_diag1<NS, RB + 1>(a, t); \
        break; \
    } \
    case QV_K_BFLY_DIAG1_G + LBL: { \
        qv_bfly<NS, RB>(a); \
        qvc t = tables[h.data + QV_IDX]; \
        if (h.w0 & (QV_UF_SCALE << 8)) t = qv_cmul(t, slices[u.scale]); \
        qv_diag1<NS, RB + 1>(a, t); \
        break; \
    }
#define QV_DG(G) \
    case QV_K_DIAG1_S + G: { \
        qvc t = slices[h.data + QV_IDX]; \
        if (h.w0 & (QV_UF_SCALE << 8)) t = qv_cmul(t, slices[u.scale]); \
        qv_diag1<NS, G>(a, t); \
        break; \
    } \
    case QV_K_DIAG1_G + G: { \
        qvc t = tables[h.data + QV_IDX]; \
        if (h.w0 & (QV_UF_SCALE << 8)) t = qv_cmul(t, slices[u.scale]); \
        qv_diag1<NS, G>(a, t); \
        break; \
    } \
    case QV_K_DIAGR_S + G: qv_diagr<NS, G>(a, slices + h.data + (QV_IDX << qv_slot_field<NS>(G))); break; \
    case QV_K_DIAGR_G + G: qv_diagr<NS, G>(a, tables + h.data + (QV_IDX << qv_slot_field<NS>(G))); break; \
    case QV_K_DIAGR_C + G: qv_diagr<NS, G>(a, M); break;
    if (NS == 16) {
        switch (kind) {
            case QV_K_BFLY + 0: qv_bfly<NS, 0>(a); break;
            case QV_K_BFLY + 1: qv_bfly<NS, 1>(a); break;
            case QV_K_BFLY + 2: qv_bfly<NS, 2>(a); break;
            case QV_K_BFLY + 3: qv_bfly<NS, (NS > 8 ? 3 : 0)>(a); break;
            QV_BD(0, 0) QV_BD(1, 1) QV_BD(2, 2) QV_BD(3, (NS > 8 ? 3 : 0))
            QV_D1(0) QV_D1(1) QV_D1(2) QV_D1(3)
            QV_D2(0, 0, 1) QV_D2(1, 0, 2) QV_D2(2, 1, 2) QV_D2(3, 0, 3) QV_D2(4, 1, 3) QV_D2(5, 2, 3)
            QV_DG(0) QV_DG(1) QV_DG(2) QV_DG(3) QV_DG(4)
            default: break;
        }
    } else {
        switch (kind) {
            case QV_K_BFLY + 0: qv_bfly<NS, 0>(a); break;
            case QV_K_BFLY + 1: qv_bfly<NS, 1>(a); break;
            case QV_K_BFLY + 2: qv_bfly<NS, 2>(a); break;
            QV_BD(0, 0) QV_BD(1, 1) QV_BD(2, 2)
            QV_D1(0) QV_D1(1) QV_D1(2)
            QV_D2(0, 0, 1) QV_D2(1, 0, 2) QV_D2(2, 1, 2)
            QV_DG(0) QV_DG(1) QV_DG(2) QV_DG(3)
            default: break;
        }
    }
#undef QV_D1
#undef QV_D2
#undef QV_DG
#undef QV_BD
#undef QV_IDX
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
