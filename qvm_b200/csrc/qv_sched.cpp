// qv_sched.cpp -- see qv_sched.h.
//
// Vocabulary
//   logical qubit : what the gate list names.
//   wire          : one bit-stream of the amplitude index.  Absorbed SWAP gates only exchange which
//                   wire a logical qubit rides on (wire_of), no data moves.
//   physical bit  : the position of a wire in the amplitude index (w2p).  It changes only when a
//                   REMAP pass physically swaps a global (rank-selecting) bit with a local bit --
//                   dqvm's "record the permutation instead of undoing it"
//                   (dqvm/src/apply-distributed-gate.lisp:38-42), here batched and deferred until a
//                   gate really needs the qubit to be local.
//   atom          : one controlled dense block or one diagonal of a gate, in wire space.
#include "qv_sched.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace qv {
namespace {

struct Atom {
    enum Kind { DIAG, DENSE, BIG } kind = DIAG;
    std::vector<int> tw;        // DENSE/BIG: wire of matrix index bit j
    std::vector<cd> mat;        // DENSE/BIG: 2^kt x 2^kt row-major; DIAG: 2^k diagonal entries
    std::vector<int> dw;        // DIAG: wire of entry index bit j
    uint64_t cmask = 0, cval = 0;   // control wires / required values
    uint64_t mix = 0;           // wires the atom mixes (its targets)
    uint64_t touch = 0;         // every wire the atom reads
};

inline int popc(uint64_t x) { return __builtin_popcountll(x); }

bool is_exact_swap(const Gate& g) {
    if (g.qubits.size() != 2) return false;
    static const int one[4] = {0, 2, 1, 3};
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) {
            cd want = (one[r] == c) ? cd(1.0, 0.0) : cd(0.0, 0.0);
            if (g.mat[r * 4 + c] != want) return false;
        }
    return true;
}

// deposit the bits of s into the set positions of mask (ascending)
uint32_t deposit_bits(uint32_t s, uint32_t mask) {
    uint32_t r = 0;
    int j = 0;
    for (int b = 0; b < 32; b++)
        if (mask >> b & 1) {
            if (s >> j & 1) r |= 1u << b;
            j++;
        }
    return r;
}

// Split a gate into atoms.  A qubit is NON-MIXING when the matrix is block diagonal with respect to
// its bit (controls, diagonal gates): such a qubit never has to be inside a tile.
void analyze(const Gate& g, const std::vector<int>& wire_of, std::vector<Atom>& out) {
    const int k = (int)g.qubits.size();
    if (k < 1 || k > 16) throw std::runtime_error("gate arity out of range (1..16)");
    const uint32_t d = 1u << k;
    if (g.mat.size() != (size_t)d * d) throw std::runtime_error("gate matrix has the wrong size");
    for (int j = 0; j < k; j++)
        for (int i = 0; i < j; i++)
            if (g.qubits[i] == g.qubits[j]) throw std::runtime_error("gate repeats a qubit");
    uint32_t mixbits = 0;
    for (uint32_t r = 0; r < d; r++)
        for (uint32_t c = 0; c < d; c++)
            if (g.mat[(size_t)r * d + c] != cd(0.0, 0.0)) mixbits |= (r ^ c);
    auto wire = [&](int j) { return wire_of[g.qubits[j]]; };
    uint64_t touch = 0;
    for (int j = 0; j < k; j++) touch |= 1ull << wire(j);

    if (mixbits == 0 && k <= QV_MAX_CHUNK_BITS) {
        Atom a;
        a.kind = Atom::DIAG;
        a.touch = touch;
        a.mat.resize(d);
        bool ident = true;
        for (uint32_t r = 0; r < d; r++) {
            a.mat[r] = g.mat[(size_t)r * d + r];
            if (a.mat[r] != cd(1.0, 0.0)) ident = false;
        }
        for (int j = 0; j < k; j++) a.dw.push_back(wire(j));
        if (!ident) out.push_back(std::move(a));
        return;
    }
    if (mixbits == 0) mixbits = d - 1;   // oversized diagonal: run it as a dense gate
    const uint32_t nonmix = (d - 1) & ~mixbits;
    const int km = popc(mixbits);
    const uint32_t dm = 1u << km;
    uint32_t v = 0;
    for (;;) {   // every value of the non-mixing bits selects one block
        std::vector<cd> sub((size_t)dm * dm);
        bool ident = true;
        for (uint32_t r = 0; r < dm; r++)
            for (uint32_t c = 0; c < dm; c++) {
                const uint32_t fr = deposit_bits(r, mixbits) | v, fc = deposit_bits(c, mixbits) | v;
                const cd e = g.mat[(size_t)fr * d + fc];
                sub[(size_t)r * dm + c] = e;
                if (e != ((r == c) ? cd(1.0, 0.0) : cd(0.0, 0.0))) ident = false;
            }
        if (!ident) {
            Atom a;
            a.kind = (km <= 2) ? Atom::DENSE : Atom::BIG;
            a.touch = touch;
            for (int j = 0; j < k; j++)
                if (mixbits >> j & 1) a.tw.push_back(wire(j));
            for (int j = 0; j < k; j++)
                if (nonmix >> j & 1) {
                    a.cmask |= 1ull << wire(j);
                    if (v >> j & 1) a.cval |= 1ull << wire(j);
                }
            a.mat = std::move(sub);
            for (int w : a.tw) a.mix |= 1ull << w;
            out.push_back(std::move(a));
        }
        if (v == nonmix) break;
        v = (v - nonmix) & nonmix;
    }
}

// ---------------------------------------------------------------- segments
std::vector<QvSeg> make_segs(const std::vector<int>& srcpos, const std::vector<int>& dstpos) {
    std::vector<QvSeg> segs;
    for (size_t i = 0; i < srcpos.size(); i++) {
        if (!segs.empty()) {
            QvSeg& s = segs.back();
            if (srcpos[i] == s.src + s.len && dstpos[i] == s.dst + s.len) {
                s.len++;
                continue;
            }
        }
        QvSeg s{};
        s.src = (uint8_t)srcpos[i];
        s.len = 1;
        s.dst = (uint8_t)dstpos[i];
        segs.push_back(s);
    }
    return segs;
}

// ---------------------------------------------------------------- diagonal chunks (physical space)
struct Chunk {
    std::vector<int> bits;      // sorted physical positions; table index bit i <-> bits[i]
    std::vector<cd> table;
};

struct DiagFactor {
    std::vector<int> pos;       // physical bit of entry index bit j
    std::vector<cd> diag;
};

struct TileMap {
    int T = 0;
    std::vector<int> tilebits;          // sorted physical bits inside the tile
    std::vector<int> local_of;          // physical bit -> tile-local position or -1
};

void chunk_multiply(Chunk& c, const DiagFactor& f) {
    const size_t n = c.table.size();
    std::vector<int> idx_of(f.pos.size());
    for (size_t j = 0; j < f.pos.size(); j++)
        idx_of[j] = (int)(std::find(c.bits.begin(), c.bits.end(), f.pos[j]) - c.bits.begin());
    for (size_t t = 0; t < n; t++) {
        uint32_t fi = 0;
        for (size_t j = 0; j < f.pos.size(); j++)
            if (t >> idx_of[j] & 1) fi |= 1u << j;
        c.table[t] *= f.diag[fi];
    }
}

void chunk_extend(Chunk& c, const std::vector<int>& newbits) {
    std::vector<int> old = c.bits;
    c.bits = newbits;
    std::vector<cd> nt((size_t)1 << newbits.size());
    std::vector<int> idx_of(old.size());
    for (size_t j = 0; j < old.size(); j++)
        idx_of[j] = (int)(std::find(newbits.begin(), newbits.end(), old[j]) - newbits.begin());
    for (size_t t = 0; t < nt.size(); t++) {
        uint32_t oi = 0;
        for (size_t j = 0; j < old.size(); j++)
            if (t >> idx_of[j] & 1) oi |= 1u << j;
        nt[t] = c.table[oi];
    }
    c.table.swap(nt);
}

// Physical bits b of a factor such that every entry with bit b clear is exactly 1 (controlled phases:
// CPHASE/CZ are gated by both of their qubits, T / PHASE by their only qubit, RZ by none).
uint64_t gating_bits(const DiagFactor& f) {
    uint64_t g = 0;
    for (size_t j = 0; j < f.pos.size(); j++) {
        bool all_one = true;
        for (size_t t = 0; t < f.diag.size() && all_one; t++)
            if (!(t >> j & 1) && f.diag[t] != cd(1.0, 0.0)) all_one = false;
        if (all_one) g |= 1ull << f.pos[j];
    }
    return g;
}

// Merge the factors of one diagonal group into chunk tables of <= QV_MAX_CHUNK_BITS bits.  Factors that
// are gated by the same REGISTER bit of the round (reg_phys) are kept together, so that the whole chunk
// stays gated by it and the kernel only touches the slots where that bit is set.
std::vector<Chunk> build_chunks(const std::vector<DiagFactor>& factors, uint64_t reg_phys) {
    std::vector<Chunk> chunks;
    std::vector<int> chunk_gate;      // physical gate bit of the chunk or -1
    std::vector<uint64_t> gates(factors.size());
    int cnt[64] = {0};
    for (size_t i = 0; i < factors.size(); i++) {
        gates[i] = gating_bits(factors[i]) & reg_phys;
        for (int b = 0; b < 64; b++)
            if (gates[i] >> b & 1) cnt[b]++;
    }
    for (size_t fi = 0; fi < factors.size(); fi++) {
        const DiagFactor& f = factors[fi];
        std::vector<int> fb = f.pos;
        std::sort(fb.begin(), fb.end());
        // preferred gate: the candidate register bit shared by most factors of the group
        int want_gate = -1;
        for (int b = 0; b < 64; b++)
            if ((gates[fi] >> b & 1) && (want_gate < 0 || cnt[b] > cnt[want_gate])) want_gate = b;
        int best = -1;
        size_t best_size = 1000;
        std::vector<int> best_union;
        for (size_t ci = 0; ci < chunks.size(); ci++) {
            if (chunk_gate[ci] != want_gate) continue;
            std::vector<int> u;
            std::set_union(chunks[ci].bits.begin(), chunks[ci].bits.end(), fb.begin(), fb.end(), std::back_inserter(u));
            if (u.size() > QV_MAX_CHUNK_BITS) continue;
            if (u.size() < best_size) {
                best_size = u.size();
                best = (int)ci;
                best_union.swap(u);
            }
        }
        if (best < 0) {
            Chunk c;
            c.bits = fb;
            c.table.assign((size_t)1 << fb.size(), cd(1.0, 0.0));
            chunks.push_back(std::move(c));
            chunk_gate.push_back(want_gate);
            best = (int)chunks.size() - 1;
        } else if (best_union.size() != chunks[best].bits.size()) {
            chunk_extend(chunks[best], best_union);
        }
        chunk_multiply(chunks[best], f);
    }
    return chunks;
}

// ---------------------------------------------------------------- pass builder
struct RoundOp {
    bool is_diag = false;
    const Atom* dense = nullptr;
    std::vector<DiagFactor> factors;
    uint64_t mix = 0, touch = 0;        // wire space
};

struct BlobWriter {
    std::vector<QvRound> rounds;
    std::vector<QvOp> ops;
    std::vector<QvChunk> chunks;
    std::vector<QvSource> sources;
    std::vector<cd> mats;
    std::vector<cd> tables;
    size_t slice_entries = 0;
};

struct Layout {
    const std::vector<int>* w2p;        // wire -> physical bit
    int phys(int w) const { return (*w2p)[w]; }
};

void emit_round(BlobWriter& w, const std::vector<RoundOp>& rops, const std::vector<int>& regpos_local,
                const TileMap& tm, const Layout& lay) {
    QvRound rd{};
    rd.m = (uint32_t)regpos_local.size();
    for (size_t i = 0; i < regpos_local.size(); i++) rd.regpos[i] = (uint32_t)regpos_local[i];
    for (uint32_t sl = 0; sl < 8; sl++) {
        uint32_t dep = 0;
        for (size_t i = 0; i < regpos_local.size(); i++)
            if (sl >> i & 1) dep |= 1u << regpos_local[i];
        rd.slot_dep[sl] = dep;
        rd.slot_xor[sl] = dep ^ ((dep >> 3) & 7u);     // qv_swz
    }
    rd.first_op = (uint32_t)w.ops.size();
    for (const RoundOp& ro : rops) {
        QvOp op{};
        if (!ro.is_diag) {
            const Atom& a = *ro.dense;
            auto rb_of = [&](int wire) {
                const int lp = tm.local_of[lay.phys(wire)];
                for (size_t i = 0; i < regpos_local.size(); i++)
                    if (regpos_local[i] == lp) return (int)i;
                throw std::runtime_error("scheduler bug: target bit is not a register bit");
            };
            std::vector<cd> mat = a.mat;
            if (a.tw.size() == 1) {
                op.type = QV_OP_DENSE1;
                op.rb0 = (uint8_t)rb_of(a.tw[0]);
            } else {
                op.type = QV_OP_DENSE2;
                int r0 = rb_of(a.tw[0]), r1 = rb_of(a.tw[1]);
                if (r0 > r1) {   // matrix index bit 0 must be the lower register bit
                    std::swap(r0, r1);
                    static const int sw[4] = {0, 2, 1, 3};
                    for (int r = 0; r < 4; r++)
                        for (int c = 0; c < 4; c++) mat[sw[r] * 4 + sw[c]] = a.mat[r * 4 + c];
                }
                op.rb0 = (uint8_t)r0;
                op.rb1 = (uint8_t)r1;
            }
            bool real = true;
            for (const cd& e : mat)
                if (e.imag() != 0.0) real = false;
            if (real) op.flags |= QV_F_REAL;
            for (int wq = 0; wq < 64; wq++) {
                if (!(a.cmask >> wq & 1)) continue;
                const bool one = a.cval >> wq & 1;
                const int b = lay.phys(wq);
                if (tm.local_of[b] >= 0) {
                    op.flags |= QV_F_CTRL_LOCAL;
                    op.cm_local |= 1u << tm.local_of[b];
                    if (one) op.cv_local |= 1u << tm.local_of[b];
                } else {
                    op.flags |= QV_F_CTRL_EXT;
                    op.cm_ext |= 1ull << b;
                    if (one) op.cv_ext |= 1ull << b;
                }
            }
            op.data_off = (uint32_t)w.mats.size();
            w.mats.insert(w.mats.end(), mat.begin(), mat.end());
        } else {
            op.type = QV_OP_DIAG;
            uint64_t reg_phys = 0;
            for (int lp : regpos_local) reg_phys |= 1ull << tm.tilebits[lp];
            op.data_off = (uint32_t)w.chunks.size();

            // How the tile-local bits `lbits` (index bit i <-> lbits[i]) reach a table / slice index:
            // register bits through slot_off, the others through lsegs.
            auto map_local = [&](QvChunk& qc, const std::vector<int>& lbits) {
                std::vector<int> lsrc, ldst;
                for (size_t i = 0; i < lbits.size(); i++) {
                    const int lp = tm.local_of[lbits[i]];
                    bool is_reg = false;
                    for (size_t r = 0; r < regpos_local.size(); r++)
                        if (regpos_local[r] == lp) {
                            qc.reg_mask |= (uint8_t)(1u << r);
                            for (uint32_t sl = 0; sl < 8; sl++)
                                if (sl >> r & 1) qc.slot_off[sl] |= 1u << i;
                            is_reg = true;
                        }
                    if (is_reg) continue;
                    lsrc.push_back(lp);
                    ldst.push_back((int)i);
                }
                std::vector<QvSeg> ls = make_segs(lsrc, ldst);
                if (ls.size() > QV_CHUNK_SEGS) throw std::runtime_error("scheduler bug: chunk needs too many segments");
                qc.n_lsegs = (uint8_t)ls.size();
                std::copy(ls.begin(), ls.end(), qc.lsegs);
            };
            // A chunk is gated by register bit r when every table entry with that bit clear is exactly 1.
            auto detect_gate = [&](QvChunk& qc, const std::vector<const Chunk*>& tabs) {
                for (size_t r = 0; r < regpos_local.size() && !qc.gate_rb; r++) {
                    if (!(qc.reg_mask >> r & 1)) continue;
                    bool all_one = true;
                    for (const Chunk* c : tabs) {
                        size_t ibit = 0;
                        for (size_t i = 0; i < c->bits.size(); i++)
                            if (tm.local_of[c->bits[i]] == regpos_local[r]) ibit = i;
                        for (size_t t = 0; t < c->table.size() && all_one; t++)
                            if (!(t >> ibit & 1) && c->table[t] != cd(1.0, 0.0)) all_one = false;
                    }
                    if (all_one) qc.gate_rb = (uint8_t)(r + 1);
                }
            };
            auto emit_global = [&](const Chunk& c) {
                QvChunk qc{};
                qc.table_off = (uint32_t)w.tables.size();
                w.tables.insert(w.tables.end(), c.table.begin(), c.table.end());
                // table index bit i <-> c.bits[i] (ascending physical); local and external bits interleave
                std::vector<int> lsrc, ldst, esrc, edst;
                for (size_t i = 0; i < c.bits.size(); i++) {
                    const int b = c.bits[i];
                    if (tm.local_of[b] >= 0) {
                        bool is_reg = false;
                        for (size_t r = 0; r < regpos_local.size(); r++)
                            if (regpos_local[r] == tm.local_of[b]) {
                                qc.reg_mask |= (uint8_t)(1u << r);
                                for (uint32_t sl = 0; sl < 8; sl++)
                                    if (sl >> r & 1) qc.slot_off[sl] |= 1u << i;
                                is_reg = true;
                            }
                        if (is_reg) continue;
                        lsrc.push_back(tm.local_of[b]);
                        ldst.push_back((int)i);
                    } else {
                        esrc.push_back(b);
                        edst.push_back((int)i);
                    }
                }
                detect_gate(qc, {&c});
                std::vector<QvSeg> ls = make_segs(lsrc, ldst), es = make_segs(esrc, edst);
                if (ls.size() > QV_CHUNK_SEGS || es.size() > QV_CHUNK_SEGS)
                    throw std::runtime_error("scheduler bug: chunk needs too many segments");
                qc.n_lsegs = (uint8_t)ls.size();
                qc.n_esegs = (uint8_t)es.size();
                std::copy(ls.begin(), ls.end(), qc.lsegs);
                std::copy(es.begin(), es.end(), qc.esegs);
                w.chunks.push_back(qc);
            };
            // SLICE chunk over the local bits `lbits` whose sources are `srcs` (bits = lbits then external bits)
            auto emit_slice = [&](const std::vector<int>& lbits, const std::vector<Chunk>& srcs) {
                QvChunk qc{};
                qc.kind = 1;
                qc.nl = (uint16_t)lbits.size();
                qc.table_off = (uint32_t)w.slice_entries;
                w.slice_entries += (size_t)1 << lbits.size();
                qc.first_src = (uint16_t)w.sources.size();
                qc.n_src = (uint16_t)srcs.size();
                map_local(qc, lbits);
                std::vector<const Chunk*> tabs;
                for (const Chunk& c : srcs) {
                    tabs.push_back(&c);
                    QvSource src{};
                    src.table_off = (uint32_t)w.tables.size();
                    w.tables.insert(w.tables.end(), c.table.begin(), c.table.end());
                    std::vector<int> esrc, edst;
                    for (size_t i = lbits.size(); i < c.bits.size(); i++) {
                        esrc.push_back(c.bits[i]);
                        edst.push_back((int)(i - lbits.size()));
                    }
                    std::vector<QvSeg> es = make_segs(esrc, edst);
                    if (es.size() > QV_CHUNK_SEGS) throw std::runtime_error("scheduler bug: source needs too many segments");
                    src.n_esegs = (uint8_t)es.size();
                    std::copy(es.begin(), es.end(), src.esegs);
                    w.sources.push_back(src);
                }
                detect_gate(qc, tabs);
                w.chunks.push_back(qc);
            };

            // 1. split the factors: purely tile-local ones, ones with external bits and few local bits
            //    (grouped by their exact local bit set), and the rest.
            std::vector<DiagFactor> local_pool, global_pool;
            std::vector<std::pair<std::vector<int>, std::vector<DiagFactor>>> ext_groups;
            for (const DiagFactor& f : ro.factors) {
                std::vector<int> L, E;
                for (int b : f.pos) (tm.local_of[b] >= 0 ? L : E).push_back(b);
                std::sort(L.begin(), L.end());
                if (E.empty()) local_pool.push_back(f);
                else if (L.size() <= 3) {
                    size_t gi = 0;
                    while (gi < ext_groups.size() && ext_groups[gi].first != L) gi++;
                    if (gi == ext_groups.size()) ext_groups.push_back({L, {}});
                    ext_groups[gi].second.push_back(f);
                } else global_pool.push_back(f);
            }
            // 2. external groups -> one SLICE each: all their tables collapse per tile into 2^|L| entries
            for (auto& grp : ext_groups) {
                const std::vector<int>& L = grp.first;
                if (w.slice_entries + ((size_t)1 << L.size()) > QV_SLICE_ENTRIES ||
                    w.sources.size() + grp.second.size() > 60000) {
                    global_pool.insert(global_pool.end(), grp.second.begin(), grp.second.end());
                    continue;
                }
                const size_t cap_ext = 10 - L.size();
                std::vector<Chunk> srcs;
                std::vector<std::vector<int>> src_ext;
                for (const DiagFactor& f : grp.second) {
                    std::vector<int> fe;
                    for (int b : f.pos)
                        if (tm.local_of[b] < 0) fe.push_back(b);
                    std::sort(fe.begin(), fe.end());
                    int best = -1;
                    size_t best_size = 1000;
                    std::vector<int> best_union;
                    for (size_t si = 0; si < srcs.size(); si++) {
                        std::vector<int> u;
                        std::set_union(src_ext[si].begin(), src_ext[si].end(), fe.begin(), fe.end(), std::back_inserter(u));
                        if (u.size() > cap_ext) continue;
                        if (u.size() < best_size) {
                            best_size = u.size();
                            best = (int)si;
                            best_union.swap(u);
                        }
                    }
                    if (best < 0) {
                        Chunk c;
                        c.bits = L;
                        c.bits.insert(c.bits.end(), fe.begin(), fe.end());
                        c.table.assign((size_t)1 << c.bits.size(), cd(1.0, 0.0));
                        srcs.push_back(std::move(c));
                        src_ext.push_back(fe);
                        best = (int)srcs.size() - 1;
                    } else if (best_union.size() != src_ext[best].size()) {
                        std::vector<int> nb = L;
                        nb.insert(nb.end(), best_union.begin(), best_union.end());
                        chunk_extend(srcs[best], nb);
                        src_ext[best] = best_union;
                    }
                    chunk_multiply(srcs[best], f);
                }
                emit_slice(L, srcs);
            }
            // 3. tile-local factors -> chunks of <= 8 bits; small ones are staged as slices (shared-memory
            //    lookups), big ones stay in global memory (L1-resident) so the slice area is kept for the
            //    external groups, where the per-tile collapse saves whole lookups
            for (const Chunk& c : build_chunks(local_pool, reg_phys)) {
                if (c.table.size() <= 16 && w.slice_entries + c.table.size() <= QV_SLICE_ENTRIES) emit_slice(c.bits, {c});
                else emit_global(c);
            }
            // 4. everything else: tables in global memory, indexed by local and external bits
            for (const Chunk& c : build_chunks(global_pool, reg_phys)) emit_global(c);
            op.n_chunks = (uint32_t)w.chunks.size() - op.data_off;
            if (getenv("QV_SCHED_DEBUG")) {
                fprintf(stderr, "  DIAG op: %zu factors (%zu local, %zu ext groups, %zu global) -> %u chunks:", ro.factors.size(),
                        local_pool.size(), ext_groups.size(), global_pool.size(), op.n_chunks);
                for (uint32_t c = op.data_off; c < w.chunks.size(); c++)
                    fprintf(stderr, " [%s nl=%u rm=%u gate=%u src=%u]", w.chunks[c].kind ? "S" : "G", w.chunks[c].nl,
                            w.chunks[c].reg_mask, w.chunks[c].gate_rb, w.chunks[c].n_src);
                fprintf(stderr, "\n");
            }
        }
        w.ops.push_back(op);
    }
    rd.n_ops = (uint32_t)w.ops.size() - rd.first_op;
    w.rounds.push_back(rd);
}

// Split the ordered atoms of one pass into register rounds (same greedy + commutation look-ahead
// as the pass level, one level down: <= QV_REG_BITS target bits per round).
void build_rounds(BlobWriter& w, const std::vector<const Atom*>& atoms, const TileMap& tm, const Layout& lay) {
    const int m_max = std::min(QV_REG_BITS, tm.T);
    std::vector<const Atom*> pending = atoms;
    while (!pending.empty()) {
        std::vector<const Atom*> deferred;
        uint64_t dmix = 0, dtouch = 0;
        uint64_t regwires = 0;
        std::vector<RoundOp> rops;
        for (const Atom* a : pending) {
            bool blocked = (a->mix & dtouch) != 0 || (dmix & a->touch) != 0;
            if (!blocked && a->kind == Atom::DENSE) {
                const uint64_t need = regwires | a->mix;
                if (popc(need) <= m_max) regwires = need;
                else blocked = true;
            }
            if (blocked) {
                deferred.push_back(a);
                dmix |= a->mix;
                dtouch |= a->touch;
                continue;
            }
            if (a->kind == Atom::DENSE) {
                RoundOp ro;
                ro.dense = a;
                ro.mix = a->mix;
                ro.touch = a->touch;
                rops.push_back(std::move(ro));
            } else {
                // hoist the diagonal backwards over commuting dense ops and merge it into the
                // nearest earlier diagonal group it can reach.
                DiagFactor f;
                for (int wq : a->dw) f.pos.push_back(lay.phys(wq));
                f.diag = a->mat;
                int j = (int)rops.size() - 1;
                bool merged = false;
                while (j >= 0) {
                    if (rops[j].is_diag) {
                        rops[j].factors.push_back(f);
                        rops[j].touch |= a->touch;
                        merged = true;
                        break;
                    }
                    if ((rops[j].mix & a->touch) != 0) break;
                    j--;
                }
                if (!merged) {
                    RoundOp ro;
                    ro.is_diag = true;
                    ro.factors.push_back(f);
                    ro.touch = a->touch;
                    rops.insert(rops.begin() + (j + 1), std::move(ro));
                }
            }
        }
        std::vector<int> regpos;
        for (int wq = 0; wq < 64; wq++)
            if (regwires >> wq & 1) regpos.push_back(tm.local_of[lay.phys(wq)]);
        for (int lp = tm.T - 1; lp >= 0 && (int)regpos.size() < m_max; lp--)
            if (std::find(regpos.begin(), regpos.end(), lp) == regpos.end()) regpos.push_back(lp);
        std::sort(regpos.begin(), regpos.end());
        emit_round(w, rops, regpos, tm, lay);
        pending.swap(deferred);
    }
}

struct Geometry {
    int n_bits, n_local, T, lmin, rank;
};

// tile_targets: physical bits that must be inside the tile.
Step build_tile_step(const std::vector<const Atom*>& atoms, uint64_t tile_targets, const Geometry& geo,
                     const Layout& lay) {
    TileMap tm;
    tm.T = geo.T;
    uint64_t tb = tile_targets;
    for (int b = 0; b < geo.lmin; b++) tb |= 1ull << b;
    for (int b = 0; b < geo.n_local && popc(tb) < tm.T; b++) tb |= 1ull << b;
    if (popc(tb) != tm.T) throw std::runtime_error("scheduler bug: tile bit count");
    tm.local_of.assign(64, -1);
    for (int b = 0; b < 64; b++)
        if (tb >> b & 1) {
            tm.local_of[b] = (int)tm.tilebits.size();
            tm.tilebits.push_back(b);
        }
    const uint64_t local_mask = (1ull << geo.n_local) - 1ull;
    const uint64_t tile_global = tb & ~local_mask;          // rank bits that vary inside the tile
    const int s = popc(tile_global);

    BlobWriter w;
    build_rounds(w, atoms, tm, lay);

    QvPassHeader h{};
    h.T = (uint32_t)tm.T;
    {
        std::vector<int> src(tm.T), dst(tm.T);
        for (int i = 0; i < tm.T; i++) {
            src[i] = i;
            dst[i] = tm.tilebits[i];
        }
        std::vector<QvSeg> sg = make_segs(src, dst);
        if (sg.size() > QV_MAX_SEGS) throw std::runtime_error("scheduler bug: too many tile segments");
        h.n_tile_segs = (uint32_t)sg.size();
        std::copy(sg.begin(), sg.end(), h.tile_segs);
    }
    // Non-tile local bits enumerate the tiles.  In a peer pass the 2^s ranks that share the tiles
    // split them: the top s non-tile local bits are pinned to this rank's value on the tile's rank bits.
    std::vector<int> nontile;
    for (int b = 0; b < geo.n_local; b++)
        if (!(tb >> b & 1)) nontile.push_back(b);
    if ((int)nontile.size() < s) throw std::runtime_error("scheduler bug: shard too small for a peer pass");
    uint64_t fixed = 0;
    for (int b = geo.n_local; b < geo.n_bits; b++)
        if (!(tb >> b & 1) && (geo.rank >> (b - geo.n_local) & 1)) fixed |= 1ull << b;
    {
        int j = 0;
        for (int b = geo.n_local; b < geo.n_bits; b++)
            if (tb >> b & 1) {
                const int pinned = nontile[nontile.size() - s + j];
                if (geo.rank >> (b - geo.n_local) & 1) fixed |= 1ull << pinned;
                j++;
            }
        nontile.resize(nontile.size() - s);
    }
    {
        std::vector<int> src, dst;
        for (size_t i = 0; i < nontile.size(); i++) {
            src.push_back((int)i);
            dst.push_back(nontile[i]);
        }
        std::vector<QvSeg> sg = make_segs(src, dst);
        if (sg.size() > QV_MAX_SEGS) throw std::runtime_error("scheduler bug: too many base segments");
        h.n_base_segs = (uint32_t)sg.size();
        std::copy(sg.begin(), sg.end(), h.base_segs);
    }
    for (int i = 0; i < 16; i++) {
        uint64_t off = 0;
        const uint32_t e = (uint32_t)i * QV_THREADS;
        for (int t = 0; t < tm.T; t++)
            if (e >> t & 1) off |= 1ull << tm.tilebits[t];
        h.hi_off[i] = off;
    }
    h.fixed_bits = fixed;
    h.n_tiles = 1ull << nontile.size();
    h.n_local_bits = (uint32_t)geo.n_local;
    h.uses_peers = s > 0 ? 1u : 0u;
    h.n_rounds = (uint32_t)w.rounds.size();
    h.n_ops = (uint32_t)w.ops.size();
    h.n_chunks = (uint32_t)w.chunks.size();
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    size_t off = align16(sizeof(QvPassHeader));
    h.off_rounds = (uint32_t)off;
    off = align16(off + w.rounds.size() * sizeof(QvRound));
    h.off_ops = (uint32_t)off;
    off = align16(off + w.ops.size() * sizeof(QvOp));
    h.off_chunks = (uint32_t)off;
    off = align16(off + w.chunks.size() * sizeof(QvChunk));
    h.off_sources = (uint32_t)off;
    off = align16(off + w.sources.size() * sizeof(QvSource));
    h.n_sources = (uint32_t)w.sources.size();
    h.n_slice_entries = (uint32_t)w.slice_entries;
    h.off_matrices = (uint32_t)off;
    off = align16(off + w.mats.size() * sizeof(cd));
    h.n_table_entries = (uint32_t)w.tables.size();
    h.blob_bytes = (uint32_t)off;
    if (off > QV_PROG_LARGE_BYTES) throw std::length_error("pass control program too large");
    if (w.chunks.size() > QV_MAX_PASS_CHUNKS) throw std::length_error("too many diagonal chunks in a pass");

    Step st;
    st.kind = Step::TILE;
    st.uses_peers = s > 0;
    st.blob.assign(off, 0);
    std::memcpy(st.blob.data(), &h, sizeof(h));
    if (!w.rounds.empty()) std::memcpy(st.blob.data() + h.off_rounds, w.rounds.data(), w.rounds.size() * sizeof(QvRound));
    if (!w.ops.empty()) std::memcpy(st.blob.data() + h.off_ops, w.ops.data(), w.ops.size() * sizeof(QvOp));
    if (!w.chunks.empty()) std::memcpy(st.blob.data() + h.off_chunks, w.chunks.data(), w.chunks.size() * sizeof(QvChunk));
    if (!w.sources.empty()) std::memcpy(st.blob.data() + h.off_sources, w.sources.data(), w.sources.size() * sizeof(QvSource));
    if (!w.mats.empty()) std::memcpy(st.blob.data() + h.off_matrices, w.mats.data(), w.mats.size() * sizeof(cd));
    st.tables = std::move(w.tables);
    st.n_gates = (int)atoms.size();
    return st;
}

Step build_big_step(const Atom& a, const Geometry& geo, const Layout& lay) {
    Step st;
    st.kind = Step::BIG;
    st.big.k = (uint32_t)a.tw.size();
    for (size_t j = 0; j < a.tw.size(); j++) st.big.pos[j] = (uint32_t)lay.phys(a.tw[j]);
    for (int wq = 0; wq < 64; wq++)
        if (a.cmask >> wq & 1) {
            st.big.ctrl_mask |= 1ull << lay.phys(wq);
            if (a.cval >> wq & 1) st.big.ctrl_val |= 1ull << lay.phys(wq);
        }
    st.big.fixed_bits = (uint64_t)geo.rank << geo.n_local;
    st.bigmat = a.mat;
    st.n_gates = 1;
    return st;
}

}  // namespace

Tape compile(const std::vector<Gate>& gates, int n_bits, const CompileOptions& opt,
             const std::vector<int>& l2p_in) {
    if (n_bits < 1 || n_bits > 40) throw std::runtime_error("qubit count out of range");
    Tape tape;
    tape.n_bits = n_bits;
    std::vector<int> w2p = l2p_in;          // wire q starts as logical qubit q
    if (w2p.empty()) {
        w2p.resize(n_bits);
        for (int i = 0; i < n_bits; i++) w2p[i] = i;
    }
    if ((int)w2p.size() != n_bits) throw std::runtime_error("l2p has the wrong length");
    std::vector<int> wire_of(n_bits);
    for (int i = 0; i < n_bits; i++) wire_of[i] = i;

    Geometry geo;
    geo.n_bits = n_bits;
    geo.n_local = opt.n_local_bits > 0 ? opt.n_local_bits : n_bits;
    geo.rank = opt.rank;
    geo.T = std::min(opt.tile_bits, geo.n_local);
    if (geo.n_local > n_bits) throw std::runtime_error("n_local_bits exceeds the qubit count");
    const int g_bits = n_bits - geo.n_local;
    if (geo.T < 1 || geo.T > QV_MAX_TILE_BITS || (geo.T < geo.n_local && geo.T < 2))
        throw std::runtime_error("tile_bits out of range");
    // keep room for a 2-target gate above the always-resident low bits
    geo.lmin = (geo.T < geo.n_local) ? std::max(0, std::min(opt.min_low_bits, geo.T - 2)) : geo.T;
    if (g_bits > 0 && geo.T - geo.lmin < 2)
        throw std::runtime_error("shard too small for the requested number of ranks");
    const uint64_t lowmask = (1ull << geo.lmin) - 1;
    const uint64_t local_mask = (1ull << geo.n_local) - 1ull;
    const int cap_high = geo.T - geo.lmin;
    Layout lay{&w2p};

    // 1. analyse gates into atoms (wire space)
    std::vector<Atom> atoms;
    std::vector<int> gate_of;
    for (size_t gi = 0; gi < gates.size(); gi++) {
        const Gate& g = gates[gi];
        for (int q : g.qubits)
            if (q < 0 || q >= n_bits) throw std::runtime_error("gate qubit out of range");
        if (opt.absorb_swaps && is_exact_swap(g)) {
            std::swap(wire_of[g.qubits[0]], wire_of[g.qubits[1]]);
            continue;
        }
        const size_t before = atoms.size();
        analyze(g, wire_of, atoms);
        for (size_t i = before; i < atoms.size(); i++) gate_of.push_back((int)gi);
    }
    tape.n_gates = (int)gates.size();
    tape.n_atoms = (int)atoms.size();

    auto phys_mask = [&](uint64_t wires) {
        uint64_t m = 0;
        for (int wq = 0; wq < n_bits; wq++)
            if (wires >> wq & 1) m |= 1ull << w2p[wq];
        return m;
    };
    auto fits = [&](uint64_t need_phys) {
        return (need_phys & ~local_mask) == 0 && popc(need_phys & ~lowmask) <= cap_high;
    };

    // The SWAP matrix used by remap passes.
    Atom swap_proto;
    swap_proto.kind = Atom::DENSE;
    swap_proto.mat.assign(16, cd(0.0, 0.0));
    swap_proto.mat[0] = swap_proto.mat[6] = swap_proto.mat[9] = swap_proto.mat[15] = cd(1.0, 0.0);
    std::vector<Atom> remap_atoms;     // storage that outlives the step builders

    // Bring the wires in `need` (currently on rank bits) onto local bits: one peer pass of physical
    // SWAPs between each needed global bit and a victim local bit (the local wire whose next use as
    // a gate target is farthest away), then relabel.
    auto remap = [&](uint64_t need_wires, const std::vector<const Atom*>& pending) {
        std::vector<int> globals;
        for (int wq = 0; wq < n_bits; wq++)
            if ((need_wires >> wq & 1) && w2p[wq] >= geo.n_local) globals.push_back(wq);
        if (globals.empty()) throw std::runtime_error("scheduler bug: remap without a global wire");
        // next use (as a mixing target) of every wire
        std::vector<size_t> next_use(n_bits, pending.size() + 1);
        for (size_t i = 0; i < pending.size(); i++)
            for (int wq = 0; wq < n_bits; wq++)
                if ((pending[i]->mix >> wq & 1) && next_use[wq] > i) next_use[wq] = i;
        std::vector<int> cand;
        for (int wq = 0; wq < n_bits; wq++)
            if (w2p[wq] < geo.n_local && w2p[wq] >= geo.lmin && !(need_wires >> wq & 1)) cand.push_back(wq);
        std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
            if (next_use[a] != next_use[b]) return next_use[a] > next_use[b];
            return w2p[a] > w2p[b];     // prefer high local bits: longer contiguous runs stay put
        });
        if (cand.size() < globals.size()) throw std::runtime_error("scheduler: no local bit available for a remap");
        if (opt.remap_pull) {
            // one out-of-place pull moves every needed pair at once
            Step st;
            st.kind = Step::REMAP;
            st.uses_peers = true;
            st.is_remap = true;
            st.remap.n_pairs = (uint32_t)globals.size();
            st.remap.n_local_bits = (uint32_t)geo.n_local;
            st.remap.rank = (uint32_t)geo.rank;
            for (size_t i = 0; i < globals.size(); i++) {
                st.remap.local_bit[i] = (uint32_t)w2p[cand[i]];
                st.remap.global_bit[i] = (uint32_t)w2p[globals[i]];
            }
            tape.steps.push_back(std::move(st));
            for (size_t i = 0; i < globals.size(); i++) std::swap(w2p[cand[i]], w2p[globals[i]]);
            return;
        }
        // in place: one peer pass exchanges as many (global, victim) pairs as the tile has room for
        const size_t per_pass = (size_t)std::max(1, cap_high / 2);
        for (size_t first = 0; first < globals.size(); first += per_pass) {
            const size_t last = std::min(globals.size(), first + per_pass);
            remap_atoms.clear();
            remap_atoms.reserve(last - first);
            uint64_t targets = 0;
            for (size_t i = first; i < last; i++) {
                Atom a = swap_proto;
                a.tw = {cand[i], globals[i]};
                a.mix = a.touch = (1ull << cand[i]) | (1ull << globals[i]);
                targets |= (1ull << w2p[cand[i]]) | (1ull << w2p[globals[i]]);
                remap_atoms.push_back(std::move(a));
            }
            std::vector<const Atom*> ptrs;
            for (const Atom& a : remap_atoms) ptrs.push_back(&a);
            Step st = build_tile_step(ptrs, targets, geo, lay);
            st.n_gates = 0;
            st.is_remap = true;
            tape.steps.push_back(std::move(st));
            for (size_t i = first; i < last; i++) std::swap(w2p[cand[i]], w2p[globals[i]]);
        }
    };

    std::vector<const Atom*> pending;
    for (const Atom& a : atoms) pending.push_back(&a);

    if (!opt.fuse) {
        // every gate is its own pass (its controlled blocks share it when they fit)
        size_t i = 0;
        while (i < atoms.size()) {
            size_t j = i;
            while (j < atoms.size() && gate_of[j] == gate_of[i]) j++;
            uint64_t need = 0;
            for (size_t a = i; a < j; a++) need |= atoms[a].mix;
            if (phys_mask(need) & ~local_mask) {
                std::vector<const Atom*> rest(pending.begin() + i, pending.end());
                remap(need, rest);
            }
            std::vector<const Atom*> cur;
            uint64_t targets = 0;
            for (size_t a = i; a < j; a++) {
                if (atoms[a].kind == Atom::BIG) {
                    if (!cur.empty()) {
                        tape.steps.push_back(build_tile_step(cur, targets, geo, lay));
                        cur.clear();
                        targets = 0;
                    }
                    tape.steps.push_back(build_big_step(atoms[a], geo, lay));
                } else {
                    const uint64_t pm = phys_mask(atoms[a].mix);
                    if (!cur.empty() && !fits(targets | pm)) {
                        tape.steps.push_back(build_tile_step(cur, targets, geo, lay));
                        cur.clear();
                        targets = 0;
                    }
                    targets |= pm;
                    cur.push_back(&atoms[a]);
                }
            }
            if (!cur.empty()) tape.steps.push_back(build_tile_step(cur, targets, geo, lay));
            i = j;
        }
        tape.l2p.resize(n_bits);
        for (int q = 0; q < n_bits; q++) tape.l2p[q] = w2p[wire_of[q]];
        return tape;
    }

    // 2. greedy pass formation with commutation look-ahead.
    while (!pending.empty()) {
        if (pending.front()->kind == Atom::BIG) {
            if (phys_mask(pending.front()->mix) & ~local_mask) {
                remap(pending.front()->mix, pending);
                continue;
            }
            tape.steps.push_back(build_big_step(*pending.front(), geo, lay));
            pending.erase(pending.begin());
            continue;
        }
        std::vector<const Atom*> in_pass, deferred;
        uint64_t dmix = 0, dtouch = 0, targets = 0;
        size_t est_bytes = sizeof(QvPassHeader);
        size_t n_diag = 0;
        for (const Atom* a : pending) {
            bool blocked = (a->mix & dtouch) != 0 || (dmix & a->touch) != 0;
            if (!blocked && a->kind == Atom::DIAG && n_diag + 1 > 4096) blocked = true;
            // conservative size of the atom in the control program (round + op + chunk + matrix)
            // rough size of the atom in the control program; diagonals merge into shared chunks, so they
            // are cheap -- the real size is checked when the pass is built (see the retry below)
            const size_t need_bytes = a->kind == Atom::DENSE ? sizeof(QvOp) + a->mat.size() * sizeof(cd) + 32 : 16;
            if (!blocked && est_bytes + need_bytes > QV_PROG_LARGE_BYTES - 2048) blocked = true;
            if (!blocked) {
                if (a->kind == Atom::BIG) blocked = true;
                else if (a->kind == Atom::DENSE) {
                    const uint64_t pm = phys_mask(a->mix);
                    if (fits(targets | pm)) targets |= pm;
                    else blocked = true;
                }
            }
            if (blocked) {
                deferred.push_back(a);
                dmix |= a->mix;
                dtouch |= a->touch;
            } else {
                in_pass.push_back(a);
                est_bytes += need_bytes;
                if (a->kind == Atom::DIAG) n_diag++;
            }
        }
        if (in_pass.empty()) {
            // the front atom needs wires that sit on rank bits: bring them (and the other global
            // targets of the atoms queued right behind it, while there is room) home first.
            uint64_t need = pending.front()->mix;
            if (!(phys_mask(need) & ~local_mask)) throw std::runtime_error("scheduler bug: no atom fits an empty pass");
            for (const Atom* a : pending) {
                const uint64_t gl = a->mix & ~need;
                uint64_t extra = 0;
                for (int wq = 0; wq < n_bits; wq++)
                    if ((gl >> wq & 1) && w2p[wq] >= geo.n_local) extra |= 1ull << wq;
                int n_glob = 0;
                for (int wq = 0; wq < n_bits; wq++)
                    if (((need | extra) >> wq & 1) && w2p[wq] >= geo.n_local) n_glob++;
                if (n_glob <= g_bits) need |= extra;
            }
            remap(need, pending);
            continue;
        }
        // Build the pass; if its control program or chunk count overflows, keep the first half of the
        // atoms and send the rest back (original order restored: pointers into `atoms` are ordered).
        for (;;) {
            try {
                tape.steps.push_back(build_tile_step(in_pass, targets, geo, lay));
                break;
            } catch (const std::length_error&) {
                if (in_pass.size() < 2) throw std::runtime_error("scheduler: a single atom overflows a pass");
                const size_t keep = in_pass.size() / 2;
                deferred.insert(deferred.end(), in_pass.begin() + keep, in_pass.end());
                std::sort(deferred.begin(), deferred.end());
                in_pass.resize(keep);
                targets = 0;
                for (const Atom* a : in_pass)
                    if (a->kind == Atom::DENSE) targets |= phys_mask(a->mix);
            }
        }
        pending.swap(deferred);
    }
    tape.l2p.resize(n_bits);
    for (int q = 0; q < n_bits; q++) tape.l2p[q] = w2p[wire_of[q]];
    return tape;
}

std::string describe(const Tape& t) {
    std::ostringstream os;
    os << "tape: n_bits=" << t.n_bits << " gates=" << t.n_gates << " atoms=" << t.n_atoms
       << " steps=" << t.steps.size() << "\n";
    for (size_t i = 0; i < t.steps.size(); i++) {
        const Step& s = t.steps[i];
        if (s.kind == Step::BIG) {
            os << "  [" << i << "] BIG k=" << s.big.k << "\n";
            continue;
        }
        if (s.kind == Step::REMAP) {
            os << "  [" << i << "] REMAP(pull)";
            for (uint32_t k = 0; k < s.remap.n_pairs; k++) os << " " << s.remap.local_bit[k] << "<->" << s.remap.global_bit[k];
            os << "\n";
            continue;
        }
        QvPassHeader h;
        std::memcpy(&h, s.blob.data(), sizeof(h));
        os << "  [" << i << "] " << (s.is_remap ? "REMAP" : (s.uses_peers ? "PEER" : "TILE")) << " T=" << h.T
           << " atoms=" << s.n_gates << " rounds=" << h.n_rounds << " ops=" << h.n_ops << " chunks=" << h.n_chunks << " sources=" << h.n_sources << " slice_entries=" << h.n_slice_entries
           << " bytes=" << h.blob_bytes << " tables=" << h.n_table_entries << " tilebits=";
        for (uint32_t k = 0; k < h.n_tile_segs; k++)
            os << (int)h.tile_segs[k].dst << "+" << (int)h.tile_segs[k].len << (k + 1 < h.n_tile_segs ? "," : "");
        os << "\n";
    }
    os << "  l2p:";
    for (int p : t.l2p) os << " " << p;
    os << "\n";
    return os.str();
}

}  // namespace qv
