// qv_sched.cpp -- see qv_sched.h.
#include "qv_sched.h"

#include <algorithm>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace qv {
namespace {

struct Atom {
    enum Kind { DIAG, DENSE, BIG } kind = DIAG;
    std::vector<int> tpos;      // DENSE/BIG: physical target bit of matrix index bit j (DENSE: ascending)
    std::vector<cd> mat;        // DENSE/BIG: 2^kt x 2^kt row-major; DIAG: 2^k diagonal entries
    std::vector<int> dpos;      // DIAG: physical bit of entry index bit j
    uint64_t cmask = 0, cval = 0;
    uint64_t mix = 0;           // physical bits the atom mixes (its targets)
    uint64_t touch = 0;         // every physical bit the atom reads
};

inline int popc(uint64_t x) { return __builtin_popcountll(x); }

inline bool commute(const Atom& a, const Atom& b) {
    // Shared bits must be non-mixing (control / diagonal) in both atoms.
    return (a.mix & b.touch) == 0 && (b.mix & a.touch) == 0;
}

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

void analyze(const Gate& g, const std::vector<int>& l2p, std::vector<Atom>& out) {
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
    auto phys = [&](int j) { return l2p[g.qubits[j]]; };
    uint64_t touch = 0;
    for (int j = 0; j < k; j++) touch |= 1ull << phys(j);

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
        for (int j = 0; j < k; j++) a.dpos.push_back(phys(j));
        if (!ident) out.push_back(std::move(a));
        return;
    }
    if (mixbits == 0) mixbits = d - 1;   // oversized diagonal: run it as a dense gate
    const uint32_t nonmix = (d - 1) & ~mixbits;
    const int km = popc(mixbits);
    const uint32_t dm = 1u << km;
    // enumerate the values of the non-mixing bits
    uint32_t v = 0;
    for (;;) {
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
            std::vector<int> tp;
            for (int j = 0; j < k; j++)
                if (mixbits >> j & 1) tp.push_back(phys(j));
            for (int j = 0; j < k; j++)
                if (nonmix >> j & 1) {
                    a.cmask |= 1ull << phys(j);
                    if (v >> j & 1) a.cval |= 1ull << phys(j);
                }
            if (a.kind == Atom::DENSE && km == 2 && tp[0] > tp[1]) {
                // make matrix bit 0 the lower physical bit
                std::swap(tp[0], tp[1]);
                static const int sw[4] = {0, 2, 1, 3};
                std::vector<cd> m2(16);
                for (int r = 0; r < 4; r++)
                    for (int c = 0; c < 4; c++) m2[sw[r] * 4 + sw[c]] = sub[r * 4 + c];
                sub.swap(m2);
            }
            a.tpos = tp;
            a.mat = std::move(sub);
            for (int p : a.tpos) a.mix |= 1ull << p;
            out.push_back(std::move(a));
        }
        if (v == nonmix) break;
        v = (v - nonmix) & nonmix;   // next subset of nonmix
    }
}

// ---------------------------------------------------------------- segments
std::vector<QvSeg> make_segs(const std::vector<int>& srcpos, const std::vector<int>& dstpos) {
    // bit srcpos[i] of the source goes to bit dstpos[i]; merge runs.
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

// ---------------------------------------------------------------- diagonal chunks
struct Chunk {
    std::vector<int> bits;      // sorted physical positions; table index bit i <-> bits[i]
    std::vector<cd> table;
};

struct DiagFactor {
    std::vector<int> pos;       // physical bit of entry index bit j
    std::vector<cd> diag;
};

int count_runs(const std::vector<int>& v) {
    int runs = 0;
    for (size_t i = 0; i < v.size(); i++)
        if (i == 0 || v[i] != v[i - 1] + 1) runs++;
    return runs;
}

struct TileMap {
    int T = 0;
    std::vector<int> tilebits;          // sorted physical bits inside the tile
    std::vector<int> local_of;          // physical bit -> tile-local position or -1
};

bool chunk_segs_ok(const std::vector<int>& bits, const TileMap& tm) {
    std::vector<int> loc, ext;
    for (int b : bits) {
        if (tm.local_of[b] >= 0) loc.push_back(tm.local_of[b]);
        else ext.push_back(b);
    }
    return count_runs(loc) <= QV_CHUNK_SEGS && count_runs(ext) <= QV_CHUNK_SEGS;
}

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

std::vector<Chunk> build_chunks(const std::vector<DiagFactor>& factors, const TileMap& tm) {
    std::vector<Chunk> chunks;
    for (const DiagFactor& f : factors) {
        std::vector<int> fb = f.pos;
        std::sort(fb.begin(), fb.end());
        int best = -1;
        size_t best_size = 1000;
        std::vector<int> best_union;
        for (size_t ci = 0; ci < chunks.size(); ci++) {
            std::vector<int> u;
            std::set_union(chunks[ci].bits.begin(), chunks[ci].bits.end(), fb.begin(), fb.end(),
                           std::back_inserter(u));
            if (u.size() > QV_MAX_CHUNK_BITS) continue;
            if (u.size() != chunks[ci].bits.size() && !chunk_segs_ok(u, tm)) continue;
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
    const Atom* dense = nullptr;            // DENSE atom
    std::vector<DiagFactor> factors;        // merged DIAG atoms
    uint64_t mix = 0, touch = 0;
};

struct BlobWriter {
    std::vector<QvRound> rounds;
    std::vector<QvOp> ops;
    std::vector<QvChunk> chunks;
    std::vector<cd> mats;
    std::vector<cd> tables;
};

void emit_round(BlobWriter& w, const std::vector<RoundOp>& rops, const std::vector<int>& regpos_local,
                const TileMap& tm) {
    QvRound rd{};
    rd.m = (uint32_t)regpos_local.size();
    for (size_t i = 0; i < regpos_local.size(); i++) rd.regpos[i] = (uint32_t)regpos_local[i];
    rd.first_op = (uint32_t)w.ops.size();
    uint64_t tile_mask = 0;
    for (int b : tm.tilebits) tile_mask |= 1ull << b;
    for (const RoundOp& ro : rops) {
        QvOp op{};
        if (!ro.is_diag) {
            const Atom& a = *ro.dense;
            op.type = a.tpos.size() == 1 ? QV_OP_DENSE1 : QV_OP_DENSE2;
            auto rb_of = [&](int physbit) {
                const int lp = tm.local_of[physbit];
                for (size_t i = 0; i < regpos_local.size(); i++)
                    if (regpos_local[i] == lp) return (int)i;
                throw std::runtime_error("scheduler bug: target bit not a register bit");
            };
            op.rb0 = (uint8_t)rb_of(a.tpos[0]);
            if (a.tpos.size() == 2) {
                op.rb1 = (uint8_t)rb_of(a.tpos[1]);
                if (op.rb0 >= op.rb1) throw std::runtime_error("scheduler bug: register bits not ascending");
            }
            bool real = true;
            for (const cd& e : a.mat)
                if (e.imag() != 0.0) real = false;
            if (real) op.flags |= QV_F_REAL;
            for (int b = 0; b < 64; b++) {
                if (!(a.cmask >> b & 1)) continue;
                const bool one = a.cval >> b & 1;
                if (tile_mask >> b & 1) {
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
            w.mats.insert(w.mats.end(), a.mat.begin(), a.mat.end());
        } else {
            op.type = QV_OP_DIAG;
            std::vector<Chunk> chunks = build_chunks(ro.factors, tm);
            op.data_off = (uint32_t)w.chunks.size();
            op.n_chunks = (uint32_t)chunks.size();
            for (const Chunk& c : chunks) {
                QvChunk qc{};
                qc.table_off = (uint32_t)w.tables.size();
                w.tables.insert(w.tables.end(), c.table.begin(), c.table.end());
                std::vector<int> lsrc, ldst, esrc, edst;
                for (size_t i = 0; i < c.bits.size(); i++) {
                    const int b = c.bits[i];
                    if (tm.local_of[b] >= 0) {
                        lsrc.push_back(tm.local_of[b]);
                        ldst.push_back((int)i);
                        for (size_t r = 0; r < regpos_local.size(); r++)
                            if (regpos_local[r] == tm.local_of[b]) qc.reg_mask |= (uint8_t)(1u << r);
                    } else {
                        esrc.push_back(b);
                        edst.push_back((int)i);
                    }
                }
                std::vector<QvSeg> ls = make_segs(lsrc, ldst), es = make_segs(esrc, edst);
                if (ls.size() > QV_CHUNK_SEGS || es.size() > QV_CHUNK_SEGS)
                    throw std::runtime_error("scheduler bug: chunk needs too many segments");
                qc.n_lsegs = (uint8_t)ls.size();
                qc.n_esegs = (uint8_t)es.size();
                std::copy(ls.begin(), ls.end(), qc.lsegs);
                std::copy(es.begin(), es.end(), qc.esegs);
                w.chunks.push_back(qc);
            }
        }
        w.ops.push_back(op);
    }
    rd.n_ops = (uint32_t)w.ops.size() - rd.first_op;
    w.rounds.push_back(rd);
}

// Split the ordered atoms of one pass into register rounds.
void build_rounds(BlobWriter& w, const std::vector<const Atom*>& atoms, const TileMap& tm) {
    const int m_max = std::min(QV_REG_BITS, tm.T);
    std::vector<const Atom*> pending = atoms;
    while (!pending.empty()) {
        std::vector<const Atom*> deferred;
        uint64_t dmix = 0, dtouch = 0;
        uint64_t regbits = 0;   // physical bits that must be register bits
        std::vector<RoundOp> rops;
        for (const Atom* a : pending) {
            bool blocked = (a->mix & dtouch) != 0 || (dmix & a->touch) != 0;
            if (!blocked && a->kind == Atom::DENSE) {
                const uint64_t need = regbits | a->mix;
                if (popc(need) <= m_max) regbits = need;
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
                // hoist the diagonal backwards over commuting dense ops and merge it
                // into the nearest earlier diagonal group it can reach.
                DiagFactor f{a->dpos, a->mat};
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
        // register bit positions: required ones, padded with unused tile-local
        // positions (highest first) up to m_max.
        std::vector<int> regpos;
        for (int b = 0; b < 64; b++)
            if (regbits >> b & 1) regpos.push_back(tm.local_of[b]);
        for (int lp = tm.T - 1; lp >= 0 && (int)regpos.size() < m_max; lp--)
            if (std::find(regpos.begin(), regpos.end(), lp) == regpos.end()) regpos.push_back(lp);
        std::sort(regpos.begin(), regpos.end());
        emit_round(w, rops, regpos, tm);
        pending.swap(deferred);
    }
}

Step build_tile_step(const std::vector<const Atom*>& atoms, uint64_t tile_targets, int n_bits,
                     const CompileOptions& opt) {
    const int n_local = opt.n_local_bits > 0 ? opt.n_local_bits : n_bits;
    TileMap tm;
    tm.T = std::min(opt.tile_bits, n_local);
    const int lmin = (tm.T < n_local) ? std::max(0, std::min(opt.min_low_bits, tm.T - 2)) : tm.T;
    uint64_t tb = tile_targets;
    for (int b = 0; b < lmin; b++) tb |= 1ull << b;
    for (int b = 0; b < n_local && popc(tb) < tm.T; b++) tb |= 1ull << b;
    if (popc(tb) != tm.T) throw std::runtime_error("scheduler bug: tile bit count");
    tm.local_of.assign(64, -1);
    for (int b = 0; b < 64; b++)
        if (tb >> b & 1) {
            tm.local_of[b] = (int)tm.tilebits.size();
            tm.tilebits.push_back(b);
        }
    BlobWriter w;
    build_rounds(w, atoms, tm);

    QvPassHeader h{};
    h.T = (uint32_t)tm.T;
    {
        std::vector<int> src(tm.T), dst(tm.T);
        for (int i = 0; i < tm.T; i++) {
            src[i] = i;
            dst[i] = tm.tilebits[i];
        }
        std::vector<QvSeg> s = make_segs(src, dst);
        if (s.size() > QV_MAX_SEGS) throw std::runtime_error("scheduler bug: too many tile segments");
        h.n_tile_segs = (uint32_t)s.size();
        std::copy(s.begin(), s.end(), h.tile_segs);
    }
    {
        std::vector<int> src, dst;
        for (int b = 0; b < n_local; b++)
            if (!(tb >> b & 1)) {
                src.push_back((int)src.size());
                dst.push_back(b);
            }
        std::vector<QvSeg> s = make_segs(src, dst);
        if (s.size() > QV_MAX_SEGS) throw std::runtime_error("scheduler bug: too many base segments");
        h.n_base_segs = (uint32_t)s.size();
        std::copy(s.begin(), s.end(), h.base_segs);
    }
    h.fixed_bits = (uint64_t)opt.rank << n_local;
    h.n_tiles = 1ull << (n_local - tm.T);
    h.n_local_bits = (uint32_t)n_local;
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
    h.off_matrices = (uint32_t)off;
    off = align16(off + w.mats.size() * sizeof(cd));
    h.n_table_entries = (uint32_t)w.tables.size();
    h.blob_bytes = (uint32_t)off;
    if (off > QV_PROG_LARGE_BYTES) throw std::runtime_error("scheduler bug: pass control program too large");

    Step st;
    st.kind = Step::TILE;
    st.blob.assign(off, 0);
    std::memcpy(st.blob.data(), &h, sizeof(h));
    if (!w.rounds.empty()) std::memcpy(st.blob.data() + h.off_rounds, w.rounds.data(), w.rounds.size() * sizeof(QvRound));
    if (!w.ops.empty()) std::memcpy(st.blob.data() + h.off_ops, w.ops.data(), w.ops.size() * sizeof(QvOp));
    if (!w.chunks.empty()) std::memcpy(st.blob.data() + h.off_chunks, w.chunks.data(), w.chunks.size() * sizeof(QvChunk));
    if (!w.mats.empty()) std::memcpy(st.blob.data() + h.off_matrices, w.mats.data(), w.mats.size() * sizeof(cd));
    st.tables = std::move(w.tables);
    st.n_gates = (int)atoms.size();
    return st;
}

Step build_big_step(const Atom& a) {
    Step st;
    st.kind = Step::BIG;
    st.big.k = (uint32_t)a.tpos.size();
    for (size_t j = 0; j < a.tpos.size(); j++) st.big.pos[j] = (uint32_t)a.tpos[j];
    st.big.ctrl_mask = a.cmask;
    st.big.ctrl_val = a.cval;
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
    tape.l2p = l2p_in;
    if (tape.l2p.empty()) {
        tape.l2p.resize(n_bits);
        for (int i = 0; i < n_bits; i++) tape.l2p[i] = i;
    }
    if ((int)tape.l2p.size() != n_bits) throw std::runtime_error("l2p has the wrong length");
    const int n_local = opt.n_local_bits > 0 ? opt.n_local_bits : n_bits;
    const int T = std::min(opt.tile_bits, n_local);
    if (T < 1 || T > QV_MAX_TILE_BITS || (T < n_local && T < 2)) throw std::runtime_error("tile_bits out of range");
    // keep room for a 2-target gate above the always-resident low bits
    const int lmin = (T < n_local) ? std::max(0, std::min(opt.min_low_bits, T - 2)) : T;
    const uint64_t lowmask = (1ull << lmin) - 1;
    const int cap_high = T - lmin;

    // 1. analyse gates into atoms; group boundaries matter only when !fuse.
    std::vector<Atom> atoms;
    std::vector<int> gate_of;   // atom -> gate index
    for (size_t gi = 0; gi < gates.size(); gi++) {
        const Gate& g = gates[gi];
        for (int q : g.qubits)
            if (q < 0 || q >= n_bits) throw std::runtime_error("gate qubit out of range");
        if (opt.absorb_swaps && is_exact_swap(g)) {
            std::swap(tape.l2p[g.qubits[0]], tape.l2p[g.qubits[1]]);
            continue;
        }
        const size_t before = atoms.size();
        analyze(g, tape.l2p, atoms);
        for (size_t i = before; i < atoms.size(); i++) gate_of.push_back((int)gi);
    }
    tape.n_gates = (int)gates.size();
    tape.n_atoms = (int)atoms.size();
    for (const Atom& a : atoms)
        if (a.kind != Atom::BIG)
            for (int b = 0; b < 64; b++)
                if ((a.mix >> b & 1) && b >= n_local)
                    throw std::runtime_error("gate mixes a physical bit that is not local to this device");

    auto fits = [&](uint64_t need) { return popc(need & ~lowmask) <= cap_high; };

    if (!opt.fuse) {
        size_t i = 0;
        while (i < atoms.size()) {
            size_t j = i;
            while (j < atoms.size() && gate_of[j] == gate_of[i]) j++;
            // atoms of one gate: tile passes for DIAG/DENSE (one pass if they fit), BIG on their own
            std::vector<const Atom*> cur;
            uint64_t targets = 0;
            for (size_t a = i; a < j; a++) {
                if (atoms[a].kind == Atom::BIG) {
                    if (!cur.empty()) {
                        tape.steps.push_back(build_tile_step(cur, targets, n_bits, opt));
                        cur.clear();
                        targets = 0;
                    }
                    tape.steps.push_back(build_big_step(atoms[a]));
                } else {
                    if (!fits(targets | atoms[a].mix)) {
                        tape.steps.push_back(build_tile_step(cur, targets, n_bits, opt));
                        cur.clear();
                        targets = 0;
                    }
                    targets |= atoms[a].mix;
                    cur.push_back(&atoms[a]);
                }
            }
            if (!cur.empty()) tape.steps.push_back(build_tile_step(cur, targets, n_bits, opt));
            i = j;
        }
        return tape;
    }

    // 2. greedy pass formation with commutation look-ahead.
    std::vector<const Atom*> pending;
    for (const Atom& a : atoms) pending.push_back(&a);
    while (!pending.empty()) {
        if (pending.front()->kind == Atom::BIG) {
            tape.steps.push_back(build_big_step(*pending.front()));
            pending.erase(pending.begin());
            continue;
        }
        std::vector<const Atom*> in_pass, deferred;
        uint64_t dmix = 0, dtouch = 0, targets = 0;
        size_t est_bytes = sizeof(QvPassHeader);
        for (const Atom* a : pending) {
            bool blocked = (a->mix & dtouch) != 0 || (dmix & a->touch) != 0;
            // conservative size of the atom in the control program (round + op + chunk + matrix)
            const size_t need_bytes = sizeof(QvRound) + sizeof(QvOp) +
                                      (a->kind == Atom::DENSE ? a->mat.size() * sizeof(cd) : sizeof(QvChunk));
            if (!blocked && est_bytes + need_bytes > QV_PROG_LARGE_BYTES - 1024) blocked = true;
            if (!blocked) {
                if (a->kind == Atom::BIG) blocked = true;
                else if (a->kind == Atom::DENSE) {
                    if (fits(targets | a->mix)) targets |= a->mix;
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
            }
        }
        if (in_pass.empty()) throw std::runtime_error("scheduler bug: no atom fits an empty pass");
        tape.steps.push_back(build_tile_step(in_pass, targets, n_bits, opt));
        pending.swap(deferred);
    }
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
        QvPassHeader h;
        std::memcpy(&h, s.blob.data(), sizeof(h));
        os << "  [" << i << "] TILE T=" << h.T << " atoms=" << s.n_gates << " rounds=" << h.n_rounds
           << " ops=" << h.n_ops << " chunks=" << h.n_chunks << " bytes=" << h.blob_bytes << " tables=" << h.n_table_entries << " tilebits=";
        for (uint32_t k = 0; k < h.n_tile_segs; k++)
            os << (int)h.tile_segs[k].dst << "+" << (int)h.tile_segs[k].len << (k + 1 < h.n_tile_segs ? "," : "");
        os << "\n";
    }
    return os.str();
}

}  // namespace qv
