// qv_jit_gen.cpp -- the pass compiler's front end: tile program -> CUDA C++ (see qv_jit_kernel.cuh for the why).
//
// Everything structural becomes a literal; the numbers (matrices, tables, tile geometry) stay in the kernel
// parameter / table pool.  The arithmetic is NOT re-implemented here: every micro-op is emitted as a call of the
// qv_ops.h template the interpreter kernel dispatches to, so interpreter, compiled pass and the test emulator
// execute the same expression for every amplitude.
#include "qv_jit.h"

#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace qv {
namespace {

struct Out {
    std::string s;
    void operator()(const char* fmt, ...) __attribute__((format(printf, 2, 3))) {
        char buf[1024];
        va_list ap;
        va_start(ap, fmt);
        const int n = vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        if (n > 0) s.append(buf, (size_t)std::min<int>(n, (int)sizeof(buf) - 1));
    }
};

uint64_t fnv1a(const std::string& s, uint64_t h = 1469598103934665603ull) {
    for (unsigned char c : s) {
        h ^= c;
        h *= 1099511628211ull;
    }
    return h;
}

inline uint32_t swz1(uint32_t e) { return e ^ ((e >> 3) & 7u); }
inline uint32_t swz2(uint32_t e) { return e ^ ((e >> 3) & 7u) ^ ((e >> 6) & 7u) ^ ((e >> 9) & 7u); }

// rank over GF(2) of three 3-bit vectors
int rank3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t v[3] = {a & 7u, b & 7u, c & 7u};
    int rank = 0;
    for (int bit = 2; bit >= 0; bit--) {
        int piv = -1;
        for (int i = rank; i < 3; i++)
            if (v[i] >> bit & 1) { piv = i; break; }
        if (piv < 0) continue;
        std::swap(v[rank], v[piv]);
        for (int i = 0; i < 3; i++)
            if (i != rank && (v[i] >> bit & 1)) v[i] ^= v[rank];
        rank++;
    }
    return rank;
}

// Shared-memory swizzle of a pass.  The default (qv_swz: 16-byte column ^= bits 3..5) keeps every ROUND free of bank
// conflicts; a store permutation that sends the three lowest tile-local bits to positions >= 6 (the QFT's bit reversal) makes
// the eight lanes of a quarter-warp read the same column in the write-back (measured: 1.0e9 conflicts, 7.6 ms instead of
// ~6 for the QFT's fourth pass, profiles/r02_a_compiled_passes.md).  The wide swizzle folds bits 6..8 and 9..11 into the
// column as well; it is chosen only when it removes that conflict (it costs one XOR per element in the tile load / store).
bool wants_wide_swizzle(const QvPassHeader& h) {
    static const bool enabled = !(getenv("QVMCUDA_JIT_WIDE_SWZ") && atoi(getenv("QVMCUDA_JIT_WIDE_SWZ")) == 0);
    if (!h.store_perm || !enabled) return false;
    // st_col[k] = qv_swz(A e_k): the swizzle is an involution, so A e_k = qv_swz(st_col[k])
    const uint32_t c0 = swz1(h.st_col[0]), c1 = swz1(h.st_col[1]), c2 = swz1(h.st_col[2]);
    const int r1 = rank3(swz1(c0), swz1(c1), swz1(c2)), r2 = rank3(swz2(c0), swz2(c1), swz2(c2));
    return r1 < 3 && r2 > r1;
}

// index expression of a diagonal micro-op over the group counter g
std::string index_expr(const QvUop& u, const uint8_t* blob) {
    std::string e;
    auto add = [&](uint32_t shift, uint32_t mask) {      // ((g >> shift) & mask)
        if (!mask) return;
        char b[96];
        if (shift) snprintf(b, sizeof(b), "((g >> %uu) & 0x%xu)", shift, mask);
        else snprintf(b, sizeof(b), "(g & 0x%xu)", mask);
        if (!e.empty()) e += " | ";
        e += b;
    };
    if (u.flags & QV_UF_GENERIC) {
        QvSegList sl;
        std::memcpy(&sl, blob + u.segs, sizeof(sl));
        for (uint32_t i = 0; i < sl.n; i++) {
            const QvSeg& q = sl.segs[i];
            const uint32_t lenmask = (1u << q.len) - 1u;
            // ((g >> src) & lenmask) << dst  ==  (g >> (src-dst)) & (lenmask << dst)   when src >= dst
            if (q.src >= q.dst) add((uint32_t)(q.src - q.dst), lenmask << q.dst);
            else {
                char b[96];
                snprintf(b, sizeof(b), "((g << %uu) & 0x%xu)", (uint32_t)(q.dst - q.src), lenmask << q.dst);
                if (!e.empty()) e += " | ";
                e += b;
            }
        }
    } else {
        add(u.cm & 0xffu, u.cm >> 8);
        add(u.cv & 0xffu, u.cv >> 8);
    }
    return e.empty() ? std::string("0u") : "(" + e + ")";
}

}  // namespace

bool jit_tma_geometry(const QvPassHeader& h, QvTmaGeom& geom) {
    std::memset(&geom, 0, sizeof(geom));
    if (h.T != QV_MAX_TILE_BITS || h.uses_peers || h.pull) return false;
    uint64_t tilemask = 0;
    for (uint32_t k = 0; k < h.n_tile_segs; k++)
        for (uint32_t b = 0; b < h.tile_segs[k].len; b++) tilemask |= 1ull << (h.tile_segs[k].dst + b);
    if ((tilemask & 15ull) != 15ull) return false;                  // the low four bits are always inside
    const uint32_t n = h.n_local_bits;
    uint32_t runs = 0;
    uint32_t b = 3;
    while (b < n) {
        const bool in = (tilemask >> b) & 1ull;
        uint32_t e = b;
        while (e < n && (((tilemask >> e) & 1ull) != 0) == in && (!in || e - b < 8)) e++;     // a box side holds <= 256
        if (runs == 4) return false;
        geom.start[runs] = (uint8_t)b;
        geom.len[runs] = (uint8_t)(e - b);
        geom.is_tile[runs] = in ? 1 : 0;
        runs++;
        b = e;
    }
    geom.n_runs = runs;
    return true;
}

JitSource jit_generate(const Step& st, int variant) {
    JitSource js;
    if (st.kind != Step::TILE) {
        js.why_not = "not a tile pass";
        return js;
    }
    const uint8_t* blob = st.blob.data();
    QvPassHeader h;
    std::memcpy(&h, blob, sizeof(h));
    if (h.T != QV_MAX_TILE_BITS) {
        js.why_not = "partial tile";
        return js;
    }
    if (h.uses_peers) {
        js.why_not = "in-place peer pass";
        return js;
    }
    const int M = (int)h.reg_bits;
    if (!((M == 3 && h.threads_log2 == 8) || (M == 4 && h.threads_log2 == 7))) {
        js.why_not = "unsupported round format";
        return js;
    }
    const int NS = 1 << M;
    QvTmaGeom geom;
    const bool tma_ok = M == 3 && !h.pull && !wants_wide_swizzle(h) && jit_tma_geometry(h, geom);
    // 32 = the source is a basis state that was never written to HBM (lazy reset): no tile loads at all
    const bool src_basis = (variant & 32) != 0;
    if (src_basis && h.pull) {
        js.why_not = "pull pass from a lazy basis state";
        return js;
    }
    const bool tma = !src_basis && (variant & 8) != 0 && tma_ok;           // persistent, double-buffered
    const bool tma_load = !src_basis && !tma && (variant & 16) != 0 && tma_ok;   // classic kernel, tile loaded by one tensor copy
    const int threads = tma ? 512 : (1 << h.threads_log2);
    const int iters = (4096 >> M) / threads;
    js.tma = tma ? 2 : tma_load ? 1 : 0;
    const QvRound* rounds = reinterpret_cast<const QvRound*>(blob + h.off_rounds);
    const QvUop* uops = reinterpret_cast<const QvUop*>(blob + h.off_uops);
    js.mode = h.pull ? 2 : 0;
    js.prog_bytes = st.blob.size() <= QV_PROG_SMALL_BYTES ? QV_PROG_SMALL_BYTES : QV_PROG_LARGE_BYTES;
    js.threads = threads;

    Out o;
    o("// compiled gate pass (qv_jit_gen.cpp); variant %d\n", variant);
    // variant bits (experiments, QVMCUDA_JIT_VARIANT): 1 = keep the two group iterations of a round rolled (fewer live
    // registers, half the code), 2 = two CTAs per SM (up to 128 registers per thread)
    const bool rolled = (variant & 1) != 0;
    const int min_ctas = (variant & 2) ? 2 : 3;
    const bool fences = (variant & 4) != 0;        // 4 = compiler fence after every micro-op (table loads are not hoisted across micro-ops)
    o("#define QVJ_M %d\n#define QVJ_THREADS %d\n#define QVJ_MIN_CTAS %d\n#define QVJ_MODE %d\n#define QVJ_PROG_BYTES %d\n", M, threads,
      min_ctas, js.mode, js.prog_bytes);
    const bool wide = wants_wide_swizzle(h);
    if (tma) o("#define QVJ_TMA 1\n");
    if (tma_load) o("#define QVJ_TMA_LOAD 1\n");
    if (src_basis) o("#define QVJ_SRC_BASIS 1\n");
    o("#define QVJ_HAS_SCALE %d\n#define QVJ_STORE_PERM %d\n#define QVJ_HAS_TABLES %d\n#define QVJ_WIDE_SWZ %d\n", h.has_scale ? 1 : 0,
      h.store_perm ? 1 : 0, (h.n_sources | h.n_preds | h.n_slice_entries) ? 1 : 0, wide ? 1 : 0);
    o("#include \"qv_jit_prelude.cuh\"\n\n");

    static const int pairs[6][2] = {{0, 1}, {0, 2}, {1, 2}, {0, 3}, {1, 3}, {2, 3}};
    for (uint32_t r = 0; r < h.n_rounds; r++) {
        const QvRound& rd = rounds[r];
        if ((int)rd.m != M) {
            js.why_not = "short round";
            return js;
        }
        o("QVJ_FN void qvj_round_%u(qvc* QVJ_RESTRICT tile, const uint32_t tid, const uint8_t* QVJ_RESTRICT blob,\n"
          "                        const qvc* QVJ_RESTRICT tables, const qvc* QVJ_RESTRICT s_slice, const uint8_t* QVJ_RESTRICT s_pred) {\n",
          r);
        o("    (void)blob; (void)tables; (void)s_slice; (void)s_pred;\n");
        o("%s\n    for (uint32_t it = 0; it < %du; it++) {\n", rolled ? "QVJ_NOUNROLL" : "QVJ_UNROLL", iters);
        o("        const uint32_t g = tid + %du * it;\n", threads);
        // e0 = g with a zero inserted at every register position (ascending)
        o("        uint32_t e0 = g;\n");
        for (int i = 0; i < M; i++) {
            // inserting at a position at or above the width of the counter so far is the identity
            const uint32_t width_before = (uint32_t)(12 - M + i);
            if (rd.regpos[i] < width_before) o("        e0 = qv_insert_zero(e0, %uu);\n", rd.regpos[i]);
        }
        o("        const uint32_t se0 = qvj_swz(e0);\n");
        o("        qvc a[%d];\n", NS);
        // slot address: the register bits are zero in e0 and the swizzle only rewrites the three column bits, so the slot of
        // (e0 | dep) is (se0 ^ x) + (dep & ~7) with x = the column bits of swizzle(dep): one XOR (none when x = 0) and an
        // immediate offset
        for (int pass = 0; pass < 2; pass++) {
            if (pass == 1) {
                // ---- micro-ops
                for (uint32_t k = rd.first_uop; uops[k].kind != QV_K_END; k++) {
                    const QvUop& u = uops[k];
                    const uint32_t kind = u.kind, flags = u.flags;
                    o("        {   // uop %u kind %u flags %u\n", k - rd.first_uop, kind, flags);
                    if (kind < QV_K_DIAG_BASE) {
                        std::string cond;
                        if (flags & QV_UF_PRED) {
                            char b[64];
                            snprintf(b, sizeof(b), "s_pred[%u]", (unsigned)u.pred);
                            cond = b;
                        }
                        const bool ctrl = (flags & QV_UF_CTRL) != 0;
                        if (ctrl && u.cm) {
                            char b[96];
                            snprintf(b, sizeof(b), "(g & 0x%xu) == 0x%xu", u.cm, u.cv);
                            cond = cond.empty() ? b : cond + " && " + b;
                        }
                        if (!cond.empty()) o("            if (%s)\n    ", cond.c_str());
                        const uint32_t ok = ctrl ? ((uint32_t)u.slot_ok & ((1u << NS) - 1u)) : 0xffffu;
                        if (kind < QV_K_DENSE2) {
                            const int rb = (int)(kind - QV_K_DENSE1) / 2, cplx = (int)(kind - QV_K_DENSE1) & 1;
                            o("            qv_dense1<%d, %d, %s, %s>(a, reinterpret_cast<const qvc*>(blob + %uu), 0x%xu);\n", NS, rb,
                              cplx ? "false" : "true", ctrl ? "true" : "false", u.data, ok);
                        } else {
                            const int p = (int)(kind - QV_K_DENSE2) / 2, cplx = (int)(kind - QV_K_DENSE2) & 1;
                            o("            qv_dense2<%d, %d, %d, %s, %s>(a, reinterpret_cast<const qvc*>(blob + %uu), 0x%xu);\n", NS,
                              pairs[p][0], pairs[p][1], cplx ? "false" : "true", ctrl ? "true" : "false", u.data, ok);
                        }
                    } else if (kind >= QV_K_BFLY && kind < QV_K_BFLY + 4) {
                        o("            qv_bfly<%d, %d>(a);\n", NS, (int)(kind - QV_K_BFLY));
                    } else {
                        // diagonal kinds (optionally with a leading butterfly)
                        uint32_t dk = kind;
                        int bfly_rb = -1;
                        if (kind >= QV_K_BFLY_DIAG1_S && kind < QV_K_BFLY_DIAG1_G) {
                            bfly_rb = (int)(kind - QV_K_BFLY_DIAG1_S);
                            dk = QV_K_DIAG1_S + 1 + (uint32_t)bfly_rb;
                        } else if (kind >= QV_K_BFLY_DIAG1_G && kind < QV_K_COUNT) {
                            bfly_rb = (int)(kind - QV_K_BFLY_DIAG1_G);
                            dk = QV_K_DIAG1_G + 1 + (uint32_t)bfly_rb;
                        }
                        if (bfly_rb >= 0) o("            qv_bfly<%d, %d>(a);\n", NS, bfly_rb);
                        const uint32_t space = (dk - QV_K_DIAG_BASE) / 5u, gate = (dk - QV_K_DIAG_BASE) % 5u;
                        const std::string idx = index_expr(u, blob);
                        const uint32_t field = (NS == 16 ? 4u : 3u) - (gate ? 1u : 0u);
                        // (Reading table offsets from the micro-op record instead -- so that passes differing only in table
                        // sizes share a kernel -- was measured: +3 % instructions, +5 % time on the QFT's heavy passes,
                        // gpurun_out/r2d_summary.md vs r2a_summary.md.  Literals it is; a pass structure costs 0.35 s to compile.)
                        if (space < 2) {
                            o("            qvc t = %s[%uu + %s];\n", space == 0 ? "s_slice" : "tables", u.data, idx.c_str());
                            if (flags & QV_UF_SCALE) o("            t = qv_cmul(t, s_slice[%uu]);\n", (unsigned)u.scale);
                            o("            qv_diag1<%d, %u>(a, t);\n", NS, gate);
                        } else if (space < 4) {
                            o("            qv_diagr<%d, %u>(a, %s + %uu + (%s << %uu));\n", NS, gate, space == 2 ? "s_slice" : "tables", u.data,
                              idx.c_str(), field);
                        } else {
                            o("            qv_diagr<%d, %u>(a, reinterpret_cast<const qvc*>(blob + %uu));\n", NS, gate, u.data);
                        }
                    }
                    o("        }\n");
                    if (fences) o("        QVJ_FENCE();\n");
                }
            }
            for (int s = 0; s < NS; s++) {
                uint32_t dep = 0;
                for (int i = 0; i < M; i++)
                    if (s >> i & 1) dep |= 1u << rd.regpos[i];
                const uint32_t sw = wide ? swz2(dep) : swz1(dep);
                const uint32_t xorpart = sw & 7u, addpart = sw & ~7u;      // == dep & ~7
                char addr[96];
                if (xorpart && addpart) snprintf(addr, sizeof(addr), "(se0 ^ 0x%xu) + 0x%xu", xorpart, addpart);
                else if (xorpart) snprintf(addr, sizeof(addr), "se0 ^ 0x%xu", xorpart);
                else if (addpart) snprintf(addr, sizeof(addr), "se0 + 0x%xu", addpart);
                else snprintf(addr, sizeof(addr), "se0");
                if (pass == 0) o("        a[%d] = tile[%s];\n", s, addr);
                else o("        tile[%s] = a[%d];\n", addr, s);
            }
        }
        o("    }\n}\n\n");
    }
    o("#define QVJ_RUN_ROUNDS(tile, tid, blob, tables, s_slice, s_pred)");
    for (uint32_t r = 0; r < h.n_rounds; r++) o(" \\\n    qvj_round_%u(tile, tid, blob, tables, s_slice, s_pred); QVJ_SYNC();", r);
    o("\n\n");
    o("#if defined(QVJ_HOST)\n// slot of the pass's layout for a slot of the default layout (the emulator stages and writes back through it)\n"
      "extern \"C\" uint32_t qvj_host_slot(uint32_t s) { return qvj_from_swz1(s); }\n#endif\n");
    o("#if defined(QVJ_HOST)\nextern \"C\" void qvj_host_round(int r, qvc* tile, uint32_t tid, const uint8_t* blob, const qvc* tables,\n"
      "                               const qvc* s_slice, const uint8_t* s_pred) {\n    switch (r) {\n");
    for (uint32_t r = 0; r < h.n_rounds; r++) o("        case %u: qvj_round_%u(tile, tid, blob, tables, s_slice, s_pred); break;\n", r, r);
    o("        default: break;\n    }\n}\n#endif\n");
    o("#include \"%s\"\n", tma ? "qv_jit_kernel_tma.cuh" : "qv_jit_kernel.cuh");
    js.text = std::move(o.s);
    js.sig = fnv1a(js.text, fnv1a(QVJIT_VERSION));
    js.ok = true;
    return js;
}

}  // namespace qv
