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
#include <deque>
#include <stdexcept>

namespace qv {
namespace {

struct Atom {
    enum Kind { DIAG, DENSE, BIG } kind = DIAG;
    std::vector<int> tw;        // DENSE/BIG: wire of matrix index bit j
    std::vector<cd> mat;        // DENSE/BIG: 2^kt x 2^kt row-major; DIAG: 2^k diagonal entries
    std::vector<int> dw;        // DIAG: wire of entry index bit j
    uint64_t cmask = 0, cval = 0;   // control wires / required values
    bool bigdiag = false;       // BIG: mat holds the 2^k diagonal entries (a diagonal on more than QV_MAX_CHUNK_BITS qubits)
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
    if (mixbits == 0) {
        // A diagonal on more than QV_MAX_CHUNK_BITS qubits: its own element-wise pass with a 2^k-entry table.  It mixes
        // nothing, so its wires may sit anywhere (rank bits included) and other atoms commute past it like past any diagonal.
        Atom a;
        a.kind = Atom::BIG;
        a.bigdiag = true;
        a.touch = touch;
        a.mat.resize(d);
        for (uint32_t r = 0; r < d; r++) a.mat[r] = g.mat[(size_t)r * d + r];
        for (int j = 0; j < k; j++) a.tw.push_back(wire(j));
        out.push_back(std::move(a));
        return;
    }
    const uint32_t nonmix = (d - 1) & ~mixbits;
    const int km = popc(mixbits);
    // checked here, before any step exists: a failure at launch time would leave earlier steps of the batch applied
    if (km > 11) throw std::runtime_error("dense gates on more than 11 mixing qubits are not supported");
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


// ---------------------------------------------------------------- host-side matrix fusion
// The reference multiplies the matrices of neighbouring gates on the same qubits before it compiles them
// (quil::fuse-gates-in-executable-code, called from src/qvm.lisp:166-175).  Same idea at atom level: an uncontrolled
// dense atom absorbs the dense atoms and 1- or 2-qubit diagonals on a SUBSET of its target wires that sit next to it in the
// dependency order (RZ.RY.RZ on one qubit becomes one 2x2; U* on the column bit, U on the row bit and the 4x4
// depolarizing superoperator that follows them on vec(rho) become one 4x4), as long as the merged block stays within two
// target wires.  X / CNOT / SWAP atoms are left alone (they fold into store permutations or relabelings for free), and so
// are diagonals that reach outside the block (CZ / CPHASE stay table lookups).  Products are formed in double precision
// on the host: the result differs from gate-by-gate application by a few ulp per merged gate, inside the 1e-12 parity
// tolerance.
bool is_linear_perm(const Atom& a);

std::vector<int> atom_wires(const Atom& a) { return a.kind == Atom::DIAG ? a.dw : a.tw; }

// full matrix of a (dense or diagonal) atom on the wire list U (index bit j <-> U[j])
std::vector<cd> expand_atom(const Atom& a, const std::vector<int>& U) {
    const std::vector<int> w = atom_wires(a);
    const size_t d = (size_t)1 << U.size();
    std::vector<int> pos(w.size());
    for (size_t j = 0; j < w.size(); j++) pos[j] = (int)(std::find(U.begin(), U.end(), w[j]) - U.begin());
    uint32_t own = 0;
    for (int p : pos) own |= 1u << p;
    auto sub = [&](size_t x) {
        uint32_t r = 0;
        for (size_t j = 0; j < w.size(); j++)
            if (x >> pos[j] & 1) r |= 1u << j;
        return r;
    };
    const size_t da = (size_t)1 << w.size();
    std::vector<cd> M(d * d, cd(0.0, 0.0));
    for (size_t r = 0; r < d; r++)
        for (size_t c = 0; c < d; c++) {
            if ((r & ~(size_t)own) != (c & ~(size_t)own)) continue;
            if (a.kind == Atom::DIAG) {
                if (r == c) M[r * d + c] = a.mat[sub(r)];
            } else {
                M[r * d + c] = a.mat[(size_t)sub(r) * da + sub(c)];
            }
        }
    return M;
}

bool fusable(const Atom& a) {
    if (a.kind == Atom::BIG || a.cmask != 0) return false;
    if (a.kind == Atom::DENSE && is_linear_perm(a)) return false;
    return atom_wires(a).size() <= 2;
}

// `later` applied after `earlier`; returns false when the pair should stay apart
bool fuse_pair(const Atom& earlier, const Atom& later, Atom& out) {
    if (!fusable(earlier) || !fusable(later)) return false;
    if (earlier.kind == Atom::DIAG && later.kind == Atom::DIAG) return false;      // diagonals merge into tables later
    const Atom& dense = earlier.kind == Atom::DENSE ? earlier : later;
    std::vector<int> U = dense.tw;
    for (const Atom* x : {&earlier, &later})
        for (int w : atom_wires(*x))
            if (std::find(U.begin(), U.end(), w) == U.end()) {
                if (x->kind == Atom::DIAG) return false;      // a diagonal that reaches outside the dense block stays a lookup
                U.push_back(w);
            }
    if (U.size() > 2) return false;
    const size_t d = (size_t)1 << U.size();
    const std::vector<cd> A = expand_atom(earlier, U), B = expand_atom(later, U);
    std::vector<cd> P(d * d, cd(0.0, 0.0));
    for (size_t r = 0; r < d; r++)
        for (size_t k = 0; k < d; k++) {
            const cd b = B[r * d + k];
            if (b == cd(0.0, 0.0)) continue;
            for (size_t c = 0; c < d; c++) P[r * d + c] += b * A[k * d + c];
        }
    out = Atom();
    bool diagonal = true;
    for (size_t r = 0; r < d; r++)
        for (size_t c = 0; c < d; c++)
            if (r != c && P[r * d + c] != cd(0.0, 0.0)) diagonal = false;
    for (int w : U) out.touch |= 1ull << w;
    if (diagonal) {
        out.kind = Atom::DIAG;
        out.dw = U;
        out.mat.resize(d);
        for (size_t r = 0; r < d; r++) out.mat[r] = P[r * d + r];
    } else {
        out.kind = Atom::DENSE;
        out.tw = U;
        out.mat = std::move(P);
        out.mix = out.touch;
    }
    return true;
}

// atoms in, atoms out (same order semantics); returns the number of merges
int fuse_matrices(std::vector<Atom>& atoms) {
    std::vector<Atom> out;
    std::vector<char> alive;
    int merges = 0;
    for (Atom& b : atoms) {
        Atom cur = std::move(b);
        long pos = -1;
        for (;;) {
            long k = (pos < 0 ? (long)out.size() : pos) - 1;
            for (; k >= 0; k--)
                if (alive[k] && (out[k].touch & cur.touch)) break;
            Atom merged;
            if (k < 0 || !fuse_pair(out[k], cur, merged)) break;
            // the merged atom takes the place of the EARLIER one: everything between the two leaves cur's wires alone
            if (pos >= 0) alive[pos] = 0;
            out[k] = std::move(merged);
            cur = out[k];
            pos = k;
            merges++;
        }
        if (pos < 0) {
            out.push_back(std::move(cur));
            alive.push_back(1);
        }
    }
    atoms.clear();
    for (size_t i = 0; i < out.size(); i++) {
        if (!alive[i]) continue;
        Atom& a = out[i];
        if (a.kind == Atom::DIAG) {      // drop products that came out as the identity
            bool ident = true;
            for (const cd& e : a.mat)
                if (e != cd(1.0, 0.0)) ident = false;
            if (ident) continue;
        }
        atoms.push_back(std::move(a));
    }
    return merges;
}

// ---------------------------------------------------------------- Euler split of complex 1-qubit unitaries
// On B200 the dense-heavy passes (layers of general 1-qubit gates) are bound by the FP64 pipe, not by HBM.  A general
// complex 2x2 costs 8 FP64 instructions per amplitude; written as  U = diag(1, l) . R . diag(r0, r1)  with R a REAL
// rotation [[c, -s], [s, c]] it costs 4, plus two diagonals -- and diagonals merge: the right factor of one layer, the
// CZ / CPHASE gates between the layers and the left factor of the previous layer all land in the same table lookup
// (one complex multiply per amplitude for up to 8 qubits).  compile() schedules both forms and keeps the cheaper one
// (tape_cost below), so an isolated gate stays a single dense micro-op.
bool euler_split_atom(const Atom& a, Atom out[3]) {
    if (a.kind != Atom::DENSE || a.tw.size() != 1 || a.cmask != 0) return false;
    const cd u00 = a.mat[0], u01 = a.mat[1], u10 = a.mat[2], u11 = a.mat[3];
    bool real = true;
    for (const cd& e : a.mat)
        if (e.imag() != 0.0) real = false;
    if (real) return false;
    // unitary?  (Kraus operators and arbitrary DEFGATE matrices are not: they stay dense)
    const cd g00 = std::conj(u00) * u00 + std::conj(u10) * u10, g01 = std::conj(u00) * u01 + std::conj(u10) * u11;
    const cd g11 = std::conj(u01) * u01 + std::conj(u11) * u11;
    if (std::abs(g00 - 1.0) > 1e-13 || std::abs(g11 - 1.0) > 1e-13 || std::abs(g01) > 1e-13) return false;
    const double c = std::abs(u00), s = std::abs(u10);
    if (c < 1e-8 || s < 1e-8) return false;     // (anti)diagonal up to rounding: nothing to gain
    const cd l0 = u00 / c, l1 = u10 / s;        // r0 = 1
    const cd r1 = u11 / (l1 * c);
    // U = diag(l0, l1) R diag(1, r1) = diag(1, l1 / l0) R diag(l0, l0 r1)
    const cd dl = l1 / l0, d0 = l0, d1 = l0 * r1;
    // reconstruction check in double precision
    const cd w00 = d0 * c, w01 = -s * d1, w10 = dl * s * d0, w11 = dl * c * d1;
    if (std::abs(w00 - u00) + std::abs(w01 - u01) + std::abs(w10 - u10) + std::abs(w11 - u11) > 4e-15) return false;
    const int w = a.tw[0];
    Atom right, rot, left;
    right.kind = Atom::DIAG;
    right.dw = {w};
    right.mat = {d0, d1};
    right.touch = 1ull << w;
    rot.kind = Atom::DENSE;
    rot.tw = {w};
    rot.mat = {cd(c, 0.0), cd(-s, 0.0), cd(s, 0.0), cd(c, 0.0)};
    rot.mix = rot.touch = 1ull << w;
    left.kind = Atom::DIAG;
    left.dw = {w};
    left.mat = {cd(1.0, 0.0), dl};
    left.touch = 1ull << w;
    out[0] = std::move(right);      // applied first
    out[1] = std::move(rot);
    out[2] = std::move(left);
    return true;
}

int euler_split(std::vector<Atom>& atoms) {
    std::vector<Atom> out;
    out.reserve(atoms.size() * 2);
    int n = 0;
    for (Atom& a : atoms) {
        Atom parts[3];
        if (euler_split_atom(a, parts)) {
            for (Atom& p : parts) out.push_back(std::move(p));
            n++;
        } else {
            out.push_back(std::move(a));
        }
    }
    atoms.swap(out);
    return n;
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

// ---------------------------------------------------------------- diagonal factors (physical space)
struct DiagFactor {
    std::vector<int> pos;       // physical bit of entry index bit j
    std::vector<cd> diag;
};

struct TileMap {
    int T = 0;
    std::vector<int> tilebits;          // sorted physical bits inside the tile
    std::vector<int> local_of;          // physical bit -> tile-local position or -1
};

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

// ---------------------------------------------------------------- pass builder
struct RoundOp {
    bool is_diag = false;
    const Atom* dense = nullptr;
    std::vector<DiagFactor> factors;
    uint64_t mix = 0, touch = 0;        // wire space
};

struct BlobWriter {
    std::vector<QvRound> rounds;
    std::vector<QvUop> uops;
    std::vector<QvSource> sources;
    std::vector<QvSlice> slices;
    std::vector<QvPred> preds;
    std::vector<QvSegList> seglists;
    std::vector<cd> mats;
    std::vector<cd> tables;
    double out_scale = 1.0;             // product of the factors of the gates that run as unscaled butterflies
    bool has_scale = false;
    size_t slice_entries = 0;
    size_t slice_build = 0;             // sum of 2^nl * n_src: per-tile construction work
    size_t n_diag_uops = 0;
};

struct Layout {
    const std::vector<int>* w2p;        // wire -> physical bit
    int phys(int w) const { return (*w2p)[w]; }
};

// f with physical bit b fixed to 1 (b disappears from the factor).
DiagFactor restrict_to_one(const DiagFactor& f, int b) {
    DiagFactor r;
    size_t jb = 0;
    for (size_t j = 0; j < f.pos.size(); j++) {
        if (f.pos[j] == b) jb = j;
        else r.pos.push_back(f.pos[j]);
    }
    r.diag.resize(f.diag.size() / 2);
    for (size_t t = 0; t < r.diag.size(); t++) {
        const size_t lo = t & (((size_t)1 << jb) - 1);
        const size_t full = ((t >> jb) << (jb + 1)) | ((size_t)1 << jb) | lo;
        r.diag[t] = f.diag[full];
    }
    return r;
}

// Table of the product of `facs` over the index layout `bits` (index bit i <-> physical bit bits[i]).
std::vector<cd> product_table(const std::vector<DiagFactor>& facs, const std::vector<int>& bits) {
    std::vector<cd> tab((size_t)1 << bits.size(), cd(1.0, 0.0));
    for (const DiagFactor& f : facs) {
        std::vector<int> idx_of(f.pos.size());
        for (size_t j = 0; j < f.pos.size(); j++) {
            const auto it = std::find(bits.begin(), bits.end(), f.pos[j]);     // layout entries of -1 are don't-care bits
            if (it == bits.end()) throw std::runtime_error("scheduler bug: factor bit missing from a table layout");
            idx_of[j] = (int)(it - bits.begin());
        }
        // plain (ac - bd, ad + bc): std::complex's operator*= goes through __muldc3 (NaN recovery) and made this loop 60 % of
        // a QFT-30 schedule; the products are the same numbers
        for (size_t t = 0; t < tab.size(); t++) {
            uint32_t fi = 0;
            for (size_t j = 0; j < f.pos.size(); j++)
                if (t >> idx_of[j] & 1) fi |= 1u << j;
            const double ar = tab[t].real(), ai = tab[t].imag(), br = f.diag[fi].real(), bi = f.diag[fi].imag();
            tab[t] = cd(ar * br - ai * bi, ar * bi + ai * br);
        }
    }
    return tab;
}

// One planned diagonal micro-op: the factors it multiplies together and the bits its table is indexed by.
struct PlanChunk {
    int gate = -1;                      // physical register bit the chunk is gated by (its factors are restricted to it), or -1
    std::vector<int> lbits;             // sorted physical tile-local bits of the index (gate excluded)
    std::vector<int> ebits;             // sorted physical bits outside the tile
    std::vector<DiagFactor> facs;
    bool as_slice = false;
};

void emit_round(BlobWriter& w, const std::vector<RoundOp>& rops, const std::vector<int>& regpos_local,
                const TileMap& tm, const Layout& lay, int reg_bits, bool butterflies) {
    const int m = (int)regpos_local.size();
    QvRound rd{};
    rd.m = (uint32_t)m;
    for (int i = 0; i < m; i++) rd.regpos[i] = (uint32_t)regpos_local[i];
    for (uint32_t sl = 0; sl < QV_MAX_SLOTS; sl++) {
        uint32_t dep = 0;
        for (int i = 0; i < m; i++)
            if (sl >> i & 1) dep |= 1u << regpos_local[i];
        rd.slot_xor[sl] = (uint16_t)(dep ^ ((dep >> 3) & 7u));     // qv_swz
    }
    rd.first_uop = (uint32_t)w.uops.size();

    // register index of a tile-local position, or -1
    auto reg_of = [&](int lp) {
        for (int i = 0; i < m; i++)
            if (regpos_local[i] == lp) return i;
        return -1;
    };
    // position in the group counter g (tile-local index with the register bits squeezed out)
    auto gpos_of = [&](int lp) {
        int below = 0;
        for (int i = 0; i < m; i++)
            if (regpos_local[i] < lp) below++;
        return lp - below;
    };
    uint64_t reg_phys = 0;
    for (int lp : regpos_local) reg_phys |= 1ull << tm.tilebits[lp];
    static const int pair_index[4][4] = {{-1, 0, 1, 3}, {0, -1, 2, 4}, {1, 2, -1, 5}, {3, 4, 5, -1}};

    for (const RoundOp& ro : rops) {
        if (!ro.is_diag) {
            const Atom& a = *ro.dense;
            QvUop u{};
            auto rb_of = [&](int wire) {
                const int r = reg_of(tm.local_of[lay.phys(wire)]);
                if (r < 0) throw std::runtime_error("scheduler bug: target bit is not a register bit");
                return r;
            };
            std::vector<cd> mat = a.mat;
            int r0 = rb_of(a.tw[0]), r1 = -1;
            if (a.tw.size() == 2) {
                r1 = rb_of(a.tw[1]);
                if (r0 > r1) {   // matrix index bit 0 must be the lower register bit
                    std::swap(r0, r1);
                    static const int sw[4] = {0, 2, 1, 3};
                    for (int r = 0; r < 4; r++)
                        for (int c = 0; c < 4; c++) mat[sw[r] * 4 + sw[c]] = a.mat[r * 4 + c];
                }
            }
            bool real = true;
            for (const cd& e : mat)
                if (e.imag() != 0.0) real = false;
            u.kind = (uint8_t)(r1 < 0 ? QV_K_DENSE1 + 2 * r0 + (real ? 0 : 1) : QV_K_DENSE2 + 2 * pair_index[r0][r1] + (real ? 0 : 1));
            // s * [[1,1],[1,-1]] without controls (Hadamard): unscaled butterfly, s goes into the write-back scale
            if (butterflies && r1 < 0 && real && a.cmask == 0 && mat[0].real() != 0.0 && mat[0] == mat[1] && mat[0] == mat[2] &&
                mat[3] == -mat[0]) {
                u.kind = (uint8_t)(QV_K_BFLY + r0);
                w.out_scale *= mat[0].real();
                w.has_scale = true;
                w.uops.push_back(u);
                continue;
            }
            uint32_t slot_ok = 0xffffu;
            QvPred pred{};
            for (int wq = 0; wq < 64; wq++) {
                if (!(a.cmask >> wq & 1)) continue;
                const bool one = a.cval >> wq & 1;
                const int b = lay.phys(wq);
                const int lp = tm.local_of[b];
                if (lp >= 0) {
                    u.flags |= QV_UF_CTRL;
                    const int r = reg_of(lp);
                    if (r >= 0) {
                        for (uint32_t sl = 0; sl < QV_MAX_SLOTS; sl++)
                            if (((sl >> r) & 1u) != (one ? 1u : 0u)) slot_ok &= ~(1u << sl);
                    } else {
                        u.cm |= 1u << gpos_of(lp);
                        if (one) u.cv |= 1u << gpos_of(lp);
                    }
                } else {
                    u.flags |= QV_UF_PRED;
                    pred.mask |= 1ull << b;
                    if (one) pred.val |= 1ull << b;
                }
            }
            u.slot_ok = (uint16_t)slot_ok;
            if (u.flags & QV_UF_PRED) {
                size_t pi = 0;
                while (pi < w.preds.size() && !(w.preds[pi].mask == pred.mask && w.preds[pi].val == pred.val)) pi++;
                if (pi == w.preds.size()) {
                    if (w.preds.size() >= QV_MAX_PREDS) throw std::length_error("too many external control predicates in a pass");
                    w.preds.push_back(pred);
                }
                u.pred = (uint8_t)pi;
            }
            u.data = (uint32_t)(w.mats.size() * sizeof(cd));   // made blob-relative when the blob is laid out
            w.mats.insert(w.mats.end(), mat.begin(), mat.end());
            w.uops.push_back(u);
            continue;
        }

        // ---------------- a merged diagonal: plan its chunks
        // 1. choose a gate for every factor: the register bit (if any) that gates most factors of the group
        const size_t nf = ro.factors.size();
        std::vector<uint64_t> gates(nf);
        int cnt[64] = {0};
        for (size_t i = 0; i < nf; i++) {
            gates[i] = gating_bits(ro.factors[i]) & reg_phys;
            for (int b = 0; b < 64; b++)
                if (gates[i] >> b & 1) cnt[b]++;
        }
        std::vector<int> want(nf, -1);
        for (size_t i = 0; i < nf; i++)
            for (int b = 0; b < 64; b++)
                if ((gates[i] >> b & 1) && (want[i] < 0 || cnt[b] > cnt[want[i]])) want[i] = b;
        // a gate whose factors would leave no other tile-local bit in the index yields a one-entry table:
        // two or more of those are cheaper as ONE ungated chunk over their gate bits
        {
            uint64_t bucket_l[64] = {0};
            for (size_t i = 0; i < nf; i++)
                if (want[i] >= 0)
                    for (int b : ro.factors[i].pos)
                        if (b != want[i] && tm.local_of[b] >= 0) bucket_l[want[i]] |= 1ull << b;
            int n_scalar = 0;
            for (int b = 0; b < 64; b++)
                if (cnt[b] && bucket_l[b] == 0) {
                    bool used = false;
                    for (size_t i = 0; i < nf; i++)
                        if (want[i] == b) used = true;
                    if (used) n_scalar++;
                }
            if (n_scalar >= 2)
                for (size_t i = 0; i < nf; i++)
                    if (want[i] >= 0 && bucket_l[want[i]] == 0) want[i] = -1;
        }
        // 2. greedy packing: a factor joins the chunk of its gate whose tile-local index grows least;
        //    external bits do not count (they are frozen per tile when the chunk becomes a slice).
        //    Index width: the non-register bits, plus -- as soon as any register bit takes part -- the whole
        //    slot field (per-slot table offsets are compile-time constants in the kernel).
        auto eff_bits = [&](const std::vector<int>& lbits, int gate) {
            size_t nonreg = 0;
            bool any_reg = false;
            for (int b : lbits) {
                if (reg_of(tm.local_of[b]) >= 0) any_reg = true;
                else nonreg++;
            }
            return nonreg + (any_reg ? (size_t)(reg_bits - (gate >= 0 ? 1 : 0)) : 0);
        };
        // Several gates at once (one-qubit phases and CZ-like factors hanging off different register bits) cost 2 FP64
        // instructions per amplitude PER GATE BIT; when the whole group fits ONE table (slot field + a few other tile-local
        // bits) a single ungated lookup does it for 4.  Measured (gpurun_out/r2g_*): no gain on random 1q/CZ layers or the density
        // circuit (5.47 vs 5.50 ms, 22.83 vs 22.85 ms) and a LOSS on the QFT (39.7 vs 37.1 ms per circuit: per-slot table loads
        // replace the gated lookups that fuse with the butterflies), so it is off unless QVMCUDA_DIAG_SINGLE_CHUNK=1.
        {
            std::vector<int> all_l;
            bool any_plain = false;
            std::vector<int> gates_used;
            for (size_t i = 0; i < nf; i++) {
                for (int b : ro.factors[i].pos)
                    if (tm.local_of[b] >= 0 && std::find(all_l.begin(), all_l.end(), b) == all_l.end()) all_l.push_back(b);
                if (want[i] < 0) any_plain = true;
                else if (std::find(gates_used.begin(), gates_used.end(), want[i]) == gates_used.end()) gates_used.push_back(want[i]);
            }
            const size_t gated_cost = 2 * gates_used.size() + (any_plain ? 4 : 0);
            static const bool single_chunk = getenv("QVMCUDA_DIAG_SINGLE_CHUNK") && atoi(getenv("QVMCUDA_DIAG_SINGLE_CHUNK")) == 1;
            if (single_chunk && gated_cost > 4 && eff_bits(all_l, -1) <= QV_MAX_CHUNK_BITS)
                for (size_t i = 0; i < nf; i++) want[i] = -1;
        }
        std::vector<PlanChunk> plan;
        for (size_t i = 0; i < nf; i++) {
            DiagFactor f = want[i] >= 0 ? restrict_to_one(ro.factors[i], want[i]) : ro.factors[i];
            std::vector<int> L, E;
            for (int b : f.pos) (tm.local_of[b] >= 0 ? L : E).push_back(b);
            std::sort(L.begin(), L.end());
            std::sort(E.begin(), E.end());
            int best = -1;
            size_t best_size = 1000;
            std::vector<int> best_union;
            for (size_t ci = 0; ci < plan.size(); ci++) {
                if (plan[ci].gate != want[i]) continue;
                std::vector<int> u;
                std::set_union(plan[ci].lbits.begin(), plan[ci].lbits.end(), L.begin(), L.end(), std::back_inserter(u));
                if (eff_bits(u, want[i]) > QV_MAX_CHUNK_BITS) continue;
                if (u.size() < best_size) {
                    best_size = u.size();
                    best = (int)ci;
                    best_union.swap(u);
                }
            }
            if (best < 0) {
                PlanChunk pc;
                pc.gate = want[i];
                pc.lbits = L;
                plan.push_back(std::move(pc));
                best = (int)plan.size() - 1;
            } else {
                plan[best].lbits = best_union;
            }
            std::vector<int> eu;
            std::set_union(plan[best].ebits.begin(), plan[best].ebits.end(), E.begin(), E.end(), std::back_inserter(eu));
            plan[best].ebits.swap(eu);
            plan[best].facs.push_back(std::move(f));
        }
        // 3. emit
        // A slice over the index layout `lay_bits` (index bit i <-> physical tile-local bit lay_bits[i]): its
        // factors are packed into source tables over (a subset of the layout, external bits).
        auto make_slice = [&](const std::vector<int>& lay_bits, const std::vector<DiagFactor>& facs) -> uint32_t {
            struct Src {
                std::vector<int> l, e;
                std::vector<DiagFactor> facs;
            };
            std::vector<Src> srcs;
            for (const DiagFactor& f : facs) {
                std::vector<int> L, E;
                for (int b : f.pos) (tm.local_of[b] >= 0 ? L : E).push_back(b);
                std::sort(L.begin(), L.end());
                std::sort(E.begin(), E.end());
                int best = -1;
                size_t best_size = 1000;
                std::vector<int> bl, be;
                for (size_t si = 0; si < srcs.size(); si++) {
                    std::vector<int> ul, ue;
                    std::set_union(srcs[si].l.begin(), srcs[si].l.end(), L.begin(), L.end(), std::back_inserter(ul));
                    std::set_union(srcs[si].e.begin(), srcs[si].e.end(), E.begin(), E.end(), std::back_inserter(ue));
                    if (ul.size() + ue.size() > QV_MAX_SOURCE_BITS) continue;
                    if (ul.size() + ue.size() < best_size) {
                        best_size = ul.size() + ue.size();
                        best = (int)si;
                        bl.swap(ul);
                        be.swap(ue);
                    }
                }
                if (best < 0) {
                    if (L.size() + E.size() > QV_MAX_SOURCE_BITS) throw std::runtime_error("scheduler bug: diagonal factor too wide");
                    srcs.push_back(Src{L, E, {}});
                    best = (int)srcs.size() - 1;
                } else {
                    srcs[best].l.swap(bl);
                    srcs[best].e.swap(be);
                }
                srcs[best].facs.push_back(f);
            }
            const size_t entries = (size_t)1 << lay_bits.size();     // -1 layout entries are don't-care bits (replicated)
            if (w.slice_entries + entries > QV_SLICE_ENTRIES || w.slices.size() >= QV_MAX_SLICES ||
                w.sources.size() + srcs.size() > QV_MAX_SOURCES || w.slice_build + entries * srcs.size() > QV_MAX_SLICE_BUILD)
                throw std::length_error("per-tile slice area exhausted");
            QvSlice qs{};
            qs.off = (uint16_t)w.slice_entries;
            qs.nl = (uint16_t)lay_bits.size();
            qs.first_src = (uint16_t)w.sources.size();
            qs.n_src = (uint16_t)srcs.size();
            for (const Src& sc : srcs) {
                QvSource qsrc{};
                // source index = its local bits in slice-layout order, then its external bits
                std::vector<int> sl_local, lsrc, ldst;
                for (size_t j = 0; j < lay_bits.size(); j++)
                    if (std::find(sc.l.begin(), sc.l.end(), lay_bits[j]) != sc.l.end()) {
                        lsrc.push_back((int)j);
                        ldst.push_back((int)sl_local.size());
                        sl_local.push_back(lay_bits[j]);
                    }
                if (sl_local.size() != sc.l.size()) throw std::runtime_error("scheduler bug: slice source bit outside the slice layout");
                std::vector<int> bits = sl_local;
                bits.insert(bits.end(), sc.e.begin(), sc.e.end());
                qsrc.nl = (uint8_t)sl_local.size();
                qsrc.table_off = (uint32_t)w.tables.size();
                const std::vector<cd> tab = product_table(sc.facs, bits);
                w.tables.insert(w.tables.end(), tab.begin(), tab.end());
                std::vector<QvSeg> ls = make_segs(lsrc, ldst);
                std::vector<int> esrc, edst;
                for (size_t j = 0; j < sc.e.size(); j++) {
                    esrc.push_back(sc.e[j]);
                    edst.push_back((int)j);
                }
                std::vector<QvSeg> es = make_segs(esrc, edst);
                if (ls.size() > QV_CHUNK_SEGS || es.size() > QV_CHUNK_SEGS) throw std::length_error("slice source needs too many segments");
                qsrc.n_lsegs = (uint8_t)ls.size();
                qsrc.n_esegs = (uint8_t)es.size();
                std::copy(ls.begin(), ls.end(), qsrc.lsegs);
                std::copy(es.begin(), es.end(), qsrc.esegs);
                w.sources.push_back(qsrc);
            }
            w.slice_build += entries * srcs.size();
            w.slice_entries += entries;
            w.slices.push_back(qs);
            return qs.off;
        };
        // One diagonal micro-op over the tile-local bits `lbits` (sorted physical), gated by `gate` (or -1).
        //   facs       : what the table holds (as a slice when as_slice, else a static table: global memory, or
        //                constants in the blob when only register bits index it)
        //   scale_facs : factors with external bits only; their per-tile product (a one-entry slice) is
        //                multiplied into the looked-up entry (DIAG1 kinds only)
        auto emit_diag = [&](int gate, const std::vector<int>& lbits, const std::vector<DiagFactor>& facs, bool as_slice,
                             const std::vector<DiagFactor>& scale_facs) {
            std::vector<int> nonreg;
            bool any_reg = false;
            for (int b : lbits) {
                if (reg_of(tm.local_of[b]) >= 0) any_reg = true;
                else nonreg.push_back(b);
            }
            const int gate_r = gate >= 0 ? reg_of(tm.local_of[gate]) : -1;
            const uint32_t gate_code = gate_r >= 0 ? (uint32_t)gate_r + 1 : 0;
            // index layout: the slot field (register bits in slot order, gate bit squeezed out; -1 = a slot bit this
            // round does not use or the table does not depend on), then the non-register bits (ascending)
            std::vector<int> lay_bits;
            if (any_reg)
                for (int r = 0; r < reg_bits; r++) {
                    if (r == gate_r) continue;
                    int phys = -1;
                    if (r < m) {
                        const int b = tm.tilebits[regpos_local[r]];
                        if (std::find(lbits.begin(), lbits.end(), b) != lbits.end()) phys = b;
                    }
                    lay_bits.push_back(phys);
                }
            lay_bits.insert(lay_bits.end(), nonreg.begin(), nonreg.end());
            QvUop u{};
            // index fields over the group counter g
            std::vector<int> src, dst;
            for (size_t j = 0; j < nonreg.size(); j++) {
                src.push_back(gpos_of(tm.local_of[nonreg[j]]));
                dst.push_back((int)j);
            }
            std::vector<QvSeg> sg = make_segs(src, dst);
            auto field = [](const QvSeg& q) { return (uint32_t)(q.src - q.dst) | ((((1u << q.len) - 1u) << q.dst) << 8); };
            if (sg.size() > QV_CHUNK_SEGS) throw std::runtime_error("scheduler bug: chunk needs too many segments");
            if (sg.size() >= 1) u.cm = field(sg[0]);
            if (sg.size() == 2) u.cv = field(sg[1]);
            if (sg.size() > 2) {
                QvSegList sl{};
                sl.n = (uint32_t)sg.size();
                std::copy(sg.begin(), sg.end(), sl.segs);
                u.flags |= QV_UF_GENERIC;
                u.cm = u.cv = 0;
                u.segs = (uint16_t)w.seglists.size();      // made blob-relative when the blob is laid out
                w.seglists.push_back(sl);
            }
            const char* where = "GLOBAL";
            if (as_slice) {
                u.kind = (uint8_t)((any_reg ? QV_K_DIAGR_S : QV_K_DIAG1_S) + gate_code);
                u.data = make_slice(lay_bits, facs);
                where = "SLICE";
            } else if (any_reg && nonreg.empty()) {
                u.kind = (uint8_t)(QV_K_DIAGR_C + gate_code);
                u.data = (uint32_t)(w.mats.size() * sizeof(cd));   // made blob-relative when the blob is laid out
                const std::vector<cd> tab = product_table(facs, lay_bits);
                w.mats.insert(w.mats.end(), tab.begin(), tab.end());
                where = "CONST";
            } else {
                u.kind = (uint8_t)((any_reg ? QV_K_DIAGR_G : QV_K_DIAG1_G) + gate_code);
                u.data = (uint32_t)w.tables.size();
                const std::vector<cd> tab = product_table(facs, lay_bits);
                w.tables.insert(w.tables.end(), tab.begin(), tab.end());
            }
            if (!scale_facs.empty()) {
                if (any_reg) throw std::runtime_error("scheduler bug: scaled diagonal with per-slot lookups");
                u.flags |= QV_UF_SCALE;
                u.scale = (uint16_t)make_slice({}, scale_facs);
            }
            w.uops.push_back(u);
            w.n_diag_uops++;
            if (getenv("QV_SCHED_DEBUG"))
                fprintf(stderr, "    uop kind=%u gate=%d nonreg=%zu any_reg=%d facs=%zu %s scale_facs=%zu\n", (unsigned)u.kind, gate,
                        nonreg.size(), (int)any_reg, facs.size(), where, scale_facs.size());
        };
        auto local_bits_of = [&](const std::vector<DiagFactor>& facs) {
            std::vector<int> bits;
            for (const DiagFactor& f : facs)
                for (int b : f.pos)
                    if (tm.local_of[b] >= 0 && std::find(bits.begin(), bits.end(), b) == bits.end()) bits.push_back(b);
            std::sort(bits.begin(), bits.end());
            return bits;
        };
        for (PlanChunk& pc : plan) {
            std::vector<DiagFactor> loc, ext;
            for (DiagFactor& f : pc.facs) {
                bool has_ext = false;
                for (int b : f.pos)
                    if (tm.local_of[b] < 0) has_ext = true;
                (has_ext ? ext : loc).push_back(f);
            }
            const size_t entries = (size_t)1 << eff_bits(pc.lbits, pc.gate);
            bool any_reg = false, any_nonreg = false;
            for (int b : pc.lbits) (reg_of(tm.local_of[b]) >= 0 ? any_reg : any_nonreg) = true;
            if (ext.empty()) {
                // static table: register-only ones are constants in the blob, tiny ones are staged per tile in
                // shared memory, the others stay in global memory (L1-resident)
                emit_diag(pc.gate, pc.lbits, loc, any_nonreg && entries <= 16, {});
            } else if (loc.empty() || entries <= 32) {
                emit_diag(pc.gate, pc.lbits, pc.facs, true, {});
            } else {
                const std::vector<int> le = local_bits_of(ext);
                if (le.empty() && !any_reg) {
                    // static table x per-tile scalar in one micro-op
                    emit_diag(pc.gate, pc.lbits, loc, false, ext);
                } else {
                    emit_diag(pc.gate, local_bits_of(loc), loc, false, {});
                    emit_diag(pc.gate, le, ext, true, {});
                }
            }
        }
        if (getenv("QV_SCHED_DEBUG")) fprintf(stderr, "  DIAG group: %zu factors -> %zu chunks (slice entries so far %zu, build %zu)\n", nf, plan.size(), w.slice_entries, w.slice_build);
    }
    // Peephole: a DIAG1 gated by register bit r commutes with every micro-op that is diagonal or acts on other
    // register bits, so it can move back to the last dense micro-op on r; if that one is an unscaled butterfly, the
    // pair becomes ONE micro-op ("H, then the controlled phases hanging off that qubit": the QFT's inner step).
    if (butterflies) {
        for (size_t i = rd.first_uop; i < w.uops.size(); i++) {
            const QvUop& d = w.uops[i];
            int space = -1;
            if (d.kind > QV_K_DIAG1_S && d.kind < QV_K_DIAG1_S + 5) space = 0;
            else if (d.kind > QV_K_DIAG1_G && d.kind < QV_K_DIAG1_G + 5) space = 1;
            if (space < 0 || (d.flags & QV_UF_GENERIC)) continue;
            const int r = d.kind - (space == 0 ? QV_K_DIAG1_S : QV_K_DIAG1_G) - 1;
            // last micro-op before i that mixes register bit r
            int j = (int)i - 1;
            for (; j >= (int)rd.first_uop; j--) {
                const QvUop& e = w.uops[j];
                bool mixes = false;
                if (e.kind >= QV_K_BFLY && e.kind < QV_K_BFLY + 4) mixes = (int)(e.kind - QV_K_BFLY) == r;
                else if (e.kind >= QV_K_BFLY_DIAG1_S && e.kind < QV_K_COUNT) mixes = (int)((e.kind - QV_K_BFLY_DIAG1_S) & 3) == r;
                else if (e.kind < QV_K_DENSE2) mixes = (int)((e.kind - QV_K_DENSE1) / 2) == r;
                else if (e.kind < QV_K_DIAG_BASE) {
                    static const int pr[6][2] = {{0, 1}, {0, 2}, {1, 2}, {0, 3}, {1, 3}, {2, 3}};
                    const int p = (e.kind - QV_K_DENSE2) / 2;
                    mixes = pr[p][0] == r || pr[p][1] == r;
                }
                if (mixes) break;
            }
            if (j < (int)rd.first_uop) continue;
            QvUop& b = w.uops[j];
            if (!(b.kind >= QV_K_BFLY && b.kind < QV_K_BFLY + 4)) continue;
            QvUop fused = d;
            fused.kind = (uint8_t)((space == 0 ? QV_K_BFLY_DIAG1_S : QV_K_BFLY_DIAG1_G) + r);
            b = fused;
            w.uops.erase(w.uops.begin() + i);
            i--;
        }
    }
    rd.n_uops = (uint32_t)w.uops.size() - rd.first_uop;
    {
        QvUop end{};
        end.kind = QV_K_END;        // the kernel's micro-op loop stops on it
        w.uops.push_back(end);
    }
    w.rounds.push_back(rd);
    if (getenv("QV_SCHED_DEBUG")) {
        fprintf(stderr, " ROUND regpos:");
        for (int lp : regpos_local) fprintf(stderr, " %d(phys %d)", lp, tm.tilebits[lp]);
        fprintf(stderr, " uops=%u\n", rd.n_uops);
    }
}

// Split the ordered atoms of one pass into register rounds (same greedy + commutation look-ahead
// as the pass level, one level down: <= reg_bits target bits per round).
void build_rounds(BlobWriter& w, const std::vector<const Atom*>& atoms, const TileMap& tm, const Layout& lay, int reg_bits,
                  bool butterflies) {
    const int m_max = std::min(reg_bits, tm.T);
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
                // Nearest group, not the earliest reachable one: merging "as soon as possible" (diagonals commute with each
                // other) was measured on the 30-qubit QFT and LOST 4.5 ms per circuit (gpurun_out/r2c_bench.json vs
                // r2a_bench_jit.json): it moves ladder phases away from the butterfly they fuse with.
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
        emit_round(w, rops, regpos, tm, lay, reg_bits, butterflies);
        pending.swap(deferred);
    }
}

struct Geometry {
    int n_bits, n_local, T, lmin, rank;
    bool store_perm;    // fold trailing X / CNOT / SWAP gates into the write-back addressing
    bool butterflies;   // run s*[[1,1],[1,-1]] gates as unscaled butterflies, s folded into the write-back scale
    int reg_bits;       // 0 = choose per pass (4 for passes that carry many gates, else 3)
};

// Amplitudes per thread per round.  Measured on B200 (gpurun_out/configs_j_m{3,4}.jsonl, QFT-30 launch lists):
// the 256-thread / 8-amplitude kernel beats the 128-thread / 16-amplitude one on every configuration (QFT-30
// heavy passes 14.4 vs 16.3 ms, random layers 25q 9.6 vs 12.6 ms) although it executes 35 % more instructions:
// the interpreter is latency-bound, and 24 resident warps per SM hide more of it than 12.  The 16-amplitude
// format stays available (CompileOptions::reg_bits = 4, QVMCUDA_REG_BITS=4) and tested.
int choose_reg_bits(const std::vector<const Atom*>&, const Geometry& geo) {
    return geo.reg_bits ? geo.reg_bits : 3;
}

// tile_targets: physical bits that must be inside the tile.
// X, singly-controlled X (CNOT, also control-on-zero) or SWAP with exact 0/1 entries.
bool is_linear_perm(const Atom& a) {
    if (a.kind != Atom::DENSE) return false;
    const cd one(1.0, 0.0), zero(0.0, 0.0);
    if (a.tw.size() == 1) {
        if (popc(a.cmask) > 1) return false;
        return a.mat[0] == zero && a.mat[1] == one && a.mat[2] == one && a.mat[3] == zero;
    }
    if (a.tw.size() == 2 && a.cmask == 0) {
        static const int img[4] = {0, 2, 1, 3};
        for (int r = 0; r < 4; r++)
            for (int c = 0; c < 4; c++)
                if (a.mat[r * 4 + c] != (img[r] == c ? one : zero)) return false;
        return true;
    }
    return false;
}

Step build_tile_step(const std::vector<const Atom*>& atoms_in, uint64_t tile_targets, const Geometry& geo,
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

    // Trailing X / CNOT / SWAP atoms whose qubits all sit inside the tile are GF(2)-affine maps of the
    // tile-local index: they are folded into the write-back addressing (store permutation) instead of
    // running as arithmetic.  An atom may be pulled out from behind later atoms it commutes with.
    std::vector<const Atom*> kept;
    std::vector<const Atom*> stripped;      // in circuit order
    {
        uint64_t later_mix = 0, later_touch = 0;
        std::vector<char> strip(atoms_in.size(), 0);
        for (size_t i = atoms_in.size(); i-- > 0;) {
            const Atom* a = atoms_in[i];
            bool lin = geo.store_perm && is_linear_perm(*a);
            if (lin)
                for (int wq = 0; wq < 64; wq++)
                    if ((a->touch >> wq & 1) && tm.local_of[lay.phys(wq)] < 0) lin = false;
            if (lin && ((a->mix & later_touch) != 0 || (later_mix & a->touch) != 0)) lin = false;
            if (lin) strip[i] = 1;
            else {
                later_mix |= a->mix;
                later_touch |= a->touch;
            }
        }
        for (size_t i = 0; i < atoms_in.size(); i++) (strip[i] ? stripped : kept).push_back(atoms_in[i]);
    }
    const std::vector<const Atom*>& atoms = kept;
    // rows of the affine map src = A e ^ b that the write-back applies: pi_1 o ... o pi_k (every pi an involution)
    uint32_t st_row[QV_MAX_TILE_BITS], st_b = 0;
    for (int k = 0; k < QV_MAX_TILE_BITS; k++) st_row[k] = 1u << k;
    for (size_t i = stripped.size(); i-- > 0;) {
        const Atom& a = *stripped[i];
        auto lp = [&](int wire) { return tm.local_of[lay.phys(wire)]; };
        if (a.tw.size() == 2) {
            const int x = lp(a.tw[0]), y = lp(a.tw[1]);
            std::swap(st_row[x], st_row[y]);
            const uint32_t bx = st_b >> x & 1, by = st_b >> y & 1;
            st_b = (st_b & ~((1u << x) | (1u << y))) | (by << x) | (bx << y);
        } else {
            const int t = lp(a.tw[0]);
            if (a.cmask == 0) st_b ^= 1u << t;
            else {
                int cw = 0;
                while (!(a.cmask >> cw & 1)) cw++;
                const int c = lp(cw);
                st_row[t] ^= st_row[c];
                st_b ^= ((st_b >> c & 1) ^ ((a.cval >> cw & 1) ? 0u : 1u)) << t;
            }
        }
    }

    const int reg_bits = choose_reg_bits(atoms, geo);
    const int threads_log2 = reg_bits == 4 ? 7 : 8;
    BlobWriter w;
    build_rounds(w, atoms, tm, lay, reg_bits, geo.butterflies);

    QvPassHeader h{};
    h.T = (uint32_t)tm.T;
    h.reg_bits = (uint32_t)reg_bits;
    h.threads_log2 = (uint32_t)threads_log2;
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
    for (int i = 0; i < 32; i++) {
        uint64_t off = 0;
        const uint32_t e = (uint32_t)i << threads_log2;
        for (int t = 0; t < tm.T; t++)
            if (e >> t & 1) off |= 1ull << tm.tilebits[t];
        h.hi_off[i] = off;
        h.hi_byte[i] = off * 16;
    }
    if (!stripped.empty()) {
        auto image = [&](uint32_t e) {      // A e (no constant)
            uint32_t r = 0;
            for (int k = 0; k < tm.T; k++)
                if (__builtin_popcount(st_row[k] & e) & 1) r |= 1u << k;
            return r;
        };
        auto swz = [](uint32_t e) { return e ^ ((e >> 3) & 7u); };
        h.store_perm = 1;
        for (int k = 0; k < tm.T; k++) h.st_col[k] = (uint16_t)swz(image(1u << k));
        h.st_const = (uint16_t)swz(st_b);
        for (int i = 0; i < 32; i++) h.st_hi[i] = (uint16_t)swz(image(((uint32_t)i << threads_log2) & ((1u << tm.T) - 1u)));
    }
    h.fixed_bits = fixed;
    h.tile_mask = tb;
    h.n_tiles = 1ull << nontile.size();
    h.n_local_bits = (uint32_t)geo.n_local;
    h.uses_peers = s > 0 ? 1u : 0u;
    h.n_rounds = (uint32_t)w.rounds.size();
    h.n_uops = (uint32_t)w.uops.size();
    h.n_diag_uops = (uint32_t)w.n_diag_uops;
    h.out_scale = w.out_scale;
    h.has_scale = w.has_scale ? 1u : 0u;
    h.n_sources = (uint32_t)w.sources.size();
    h.n_slices = (uint32_t)w.slices.size();
    h.n_slice_entries = (uint32_t)w.slice_entries;
    h.n_preds = (uint32_t)w.preds.size();
    std::vector<uint8_t> slice_of(w.slice_entries);
    for (size_t i = 0; i < w.slices.size(); i++)
        for (size_t x = 0; x < ((size_t)1 << w.slices[i].nl); x++) slice_of[w.slices[i].off + x] = (uint8_t)i;
    auto align16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
    // Layout: header, rounds, micro-ops, matrices first -- their offsets depend only on the STRUCTURE of the pass (round and
    // micro-op counts, matrix sizes), which lets the pass compiler (qv_jit_gen.cpp) address matrices as literals and still
    // share one kernel between passes that differ in tile geometry / number of slice sources -- then the per-tile tables.
    size_t off = align16(sizeof(QvPassHeader));
    h.off_rounds = (uint32_t)off;
    off = align16(off + w.rounds.size() * sizeof(QvRound));
    h.off_uops = (uint32_t)off;
    off = align16(off + w.uops.size() * sizeof(QvUop));
    h.off_matrices = (uint32_t)off;
    off = align16(off + w.mats.size() * sizeof(cd));
    const size_t off_seglists = off;
    off = align16(off + w.seglists.size() * sizeof(QvSegList));
    h.off_sources = (uint32_t)off;
    off = align16(off + w.sources.size() * sizeof(QvSource));
    h.off_slices = (uint32_t)off;
    off = align16(off + w.slices.size() * sizeof(QvSlice));
    h.off_slice_of = (uint32_t)off;
    off = align16(off + slice_of.size());
    h.off_preds = (uint32_t)off;
    off = align16(off + w.preds.size() * sizeof(QvPred));
    h.n_table_entries = (uint32_t)w.tables.size();
    h.blob_bytes = (uint32_t)off;
    if (off > QV_PROG_LARGE_BYTES) throw std::length_error("pass control program too large");
    for (QvUop& u : w.uops) {
        if (u.kind == QV_K_END || u.kind >= QV_K_BFLY) continue;      // no blob-relative operands
        if (u.kind < QV_K_DIAG_BASE || u.kind >= QV_K_DIAGR_C) u.data += h.off_matrices;
        if (u.kind >= QV_K_DIAG_BASE && (u.flags & QV_UF_GENERIC)) u.segs = (uint16_t)(off_seglists + u.segs * sizeof(QvSegList));
    }

    Step st;
    st.kind = Step::TILE;
    st.uses_peers = s > 0;
    st.blob.assign(off, 0);
    std::memcpy(st.blob.data(), &h, sizeof(h));
    auto put = [&](size_t at, const void* src, size_t bytes) {
        if (bytes) std::memcpy(st.blob.data() + at, src, bytes);
    };
    put(h.off_rounds, w.rounds.data(), w.rounds.size() * sizeof(QvRound));
    put(h.off_uops, w.uops.data(), w.uops.size() * sizeof(QvUop));
    put(h.off_sources, w.sources.data(), w.sources.size() * sizeof(QvSource));
    put(h.off_slices, w.slices.data(), w.slices.size() * sizeof(QvSlice));
    put(h.off_slice_of, slice_of.data(), slice_of.size());
    put(h.off_preds, w.preds.data(), w.preds.size() * sizeof(QvPred));
    put(off_seglists, w.seglists.data(), w.seglists.size() * sizeof(QvSegList));
    put(h.off_matrices, w.mats.data(), w.mats.size() * sizeof(cd));
    st.tables = std::move(w.tables);
    st.n_gates = (int)atoms_in.size();
    return st;
}

Step build_big_step(const Atom& a, const Geometry& geo, const Layout& lay) {
    Step st;
    st.kind = Step::BIG;
    st.big.k = (uint32_t)a.tw.size();
    st.big.diag = a.bigdiag ? 1u : 0u;
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

// A pull remap followed by a local tile pass becomes ONE pass that loads through the remap (from every rank's
// current buffer, over NVLink where the source is remote) and stores into the alternate buffer.
void fuse_pulls(Tape& tape, const CompileOptions& opt) {
    if (!opt.remap_pull || !opt.fuse_pull) return;
    std::vector<Step> out;
    for (size_t i = 0; i < tape.steps.size(); i++) {
        Step& st = tape.steps[i];
        if (st.kind == Step::REMAP && i + 1 < tape.steps.size() && tape.steps[i + 1].kind == Step::TILE &&
            !tape.steps[i + 1].uses_peers) {
            Step nx = std::move(tape.steps[i + 1]);
            QvPassHeader h;
            std::memcpy(&h, nx.blob.data(), sizeof(h));
            h.pull = 1;
            h.pull_remap = st.remap;
            for (int k = 0; k < 32; k++) {
                uint64_t S = h.hi_off[k];
                for (uint32_t p = 0; p < st.remap.n_pairs; p++) {
                    const uint64_t x = ((h.hi_off[k] >> st.remap.local_bit[p]) ^ (h.hi_off[k] >> st.remap.global_bit[p])) & 1ull;
                    S ^= (x << st.remap.local_bit[p]) | (x << st.remap.global_bit[p]);
                }
                h.hi_src[k] = S;
            }
            std::memcpy(nx.blob.data(), &h, sizeof(h));
            nx.uses_peers = true;       // reads peer shards: the host barriers around it
            nx.is_remap = true;
            out.push_back(std::move(nx));
            i++;
            continue;
        }
        out.push_back(std::move(st));
    }
    tape.steps.swap(out);
}

}  // namespace

// Cost model of a tape in FP64 instructions per amplitude (the unit the dense-heavy passes are bound by on B200).  A pass
// cannot be faster than its HBM traffic: 32 B per amplitude at ~6.2 TB/s against 64 FP64 lanes x 148 SMs x 1.965 GHz is worth
// about 96 FP64 instructions per amplitude; exchange passes cost about eight times that (NVLink ingress).
double tape_cost(const Tape& t) {
    double total = 0.0;
    for (const Step& st : t.steps) {
        if (st.kind == Step::REMAP) { total += 800.0; continue; }
        if (st.kind == Step::BIG) { total += std::max(96.0, 8.0 * (double)(1u << st.big.k)); continue; }
        QvPassHeader h;
        std::memcpy(&h, st.blob.data(), sizeof(h));
        const QvUop* uops = reinterpret_cast<const QvUop*>(st.blob.data() + h.off_uops);
        double fp = h.has_scale ? 2.0 : 0.0, other = 10.0 + 3.0 * h.n_rounds;
        for (uint32_t k = 0; k < h.n_uops; k++) {
            const uint32_t kind = uops[k].kind;
            if (kind == QV_K_END) continue;
            const double share = (uops[k].flags & (QV_UF_CTRL | QV_UF_PRED)) ? 0.5 : 1.0;
            if (kind < QV_K_DENSE2) fp += share * ((kind & 1) ? 8.5 : 4.0);
            else if (kind < QV_K_DIAG_BASE) fp += share * ((kind & 1) ? 17.0 : 8.0);
            else if (kind >= QV_K_BFLY && kind < QV_K_BFLY + 4) fp += 2.0;
            else if (kind >= QV_K_BFLY_DIAG1_S) { fp += 4.0; other += 1.0; }
            else {
                const uint32_t gate = (kind - QV_K_DIAG_BASE) % 5u;
                fp += gate ? 2.0 : 4.0;
                other += 1.0;
            }
        }
        const double floor_cost = (h.pull || st.uses_peers) ? 800.0 : 96.0;
        total += std::max(floor_cost, fp / 0.8 + 0.25 * other);
    }
    return total;
}

Tape compile(const std::vector<Gate>& gates, int n_bits, const CompileOptions& opt,
             const std::vector<int>& l2p_in) {
    if (n_bits < 1 || n_bits > 40) throw std::runtime_error("qubit count out of range");
    if (opt.fuse && opt.euler_split < 0) {
        // schedule with fused complex 1q matrices; if any could be written diag . rotation . diag, schedule that form too
        // and keep the cheaper tape (both are exact to rounding; the choice is a cost model, not a semantic one)
        CompileOptions o = opt;
        o.euler_split = 0;
        Tape fused = compile(gates, n_bits, o, l2p_in);
        if (fused.n_splittable == 0) return fused;
        o.euler_split = 1;
        Tape split = compile(gates, n_bits, o, l2p_in);
        const double cf = tape_cost(fused), cs = tape_cost(split);
        if (getenv("QV_SCHED_DEBUG")) fprintf(stderr, "euler split: cost fused %.1f split %.1f\n", cf, cs);
        return cs < cf ? split : fused;
    }
    if (opt.fuse && opt.hoist_remaps < 0) {
        // sharded, fused pulls: schedule with and without remap hoisting, keep the cheaper tape (ties: hoisted)
        CompileOptions o = opt;
        o.hoist_remaps = 0;
        const int nl = opt.n_local_bits > 0 ? opt.n_local_bits : n_bits;
        if (nl == n_bits || !opt.remap_pull || !opt.fuse_pull) return compile(gates, n_bits, o, l2p_in);
        Tape plain = compile(gates, n_bits, o, l2p_in);
        o.hoist_remaps = 1;
        Tape hoisted = compile(gates, n_bits, o, l2p_in);
        if (hoisted.n_hoisted == 0) return plain;
        const double cp = tape_cost(plain), ch = tape_cost(hoisted);
        if (getenv("QV_SCHED_DEBUG")) fprintf(stderr, "remap hoisting: cost plain %.1f (%zu steps) hoisted %.1f (%zu steps)\n", cp,
                                              plain.steps.size(), ch, hoisted.steps.size());
        return ch <= cp ? hoisted : plain;
    }
    if (opt.fuse && opt.route_swaps < 0) {
        // Exact SWAP gates either run where they stand (folded into a pass's write-back when they trail it, else a pass of
        // their own) or are absorbed as relabelings whose physical permutation is spread over the spare tile slots of the
        // gate passes (route_swaps = 1).  From the canonical layout both tapes end in the canonical layout; from another one
        // (left behind by an absorb_swaps run) the plain tape keeps it and the routed tape canonicalises.  Keep the cheaper one.
        CompileOptions o = opt;
        o.route_swaps = 0;
        const int nl = opt.n_local_bits > 0 ? opt.n_local_bits : n_bits;
        bool any = false;
        for (const Gate& g : gates)
            if (is_exact_swap(g)) any = true;
        for (size_t q = 0; q < l2p_in.size(); q++)
            if (l2p_in[q] != (int)q) any = true;
        if (nl != n_bits || opt.absorb_swaps || !any) return compile(gates, n_bits, o, l2p_in);
        Tape plain = compile(gates, n_bits, o, l2p_in);
        o.route_swaps = 1;
        Tape routed = compile(gates, n_bits, o, l2p_in);
        const double cp = tape_cost(plain), cr = tape_cost(routed);
        if (getenv("QV_SCHED_DEBUG")) fprintf(stderr, "swap routing: cost plain %.1f (%zu steps) routed %.1f (%zu steps)\n", cp,
                                              plain.steps.size(), cr, routed.steps.size());
        return cr < cp ? routed : plain;
    }
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
    geo.reg_bits = opt.reg_bits;
    geo.store_perm = opt.store_perm;
    geo.butterflies = opt.butterflies;
    if (opt.reg_bits != 0 && opt.reg_bits != 3 && opt.reg_bits != 4) throw std::runtime_error("reg_bits must be 0 (auto), 3 or 4");
    geo.T = std::min(opt.tile_bits, geo.n_local);
    if (geo.n_local > n_bits) throw std::runtime_error("n_local_bits exceeds the qubit count");
    const int g_bits = n_bits - geo.n_local;
    if (geo.T < 1 || geo.T > QV_MAX_TILE_BITS || (geo.T < geo.n_local && geo.T < 2))
        throw std::runtime_error("tile_bits out of range");
    // keep room for a 2-target gate above the always-resident low bits
    // (a shard that fits one tile still needs room for rank bits in a peer pass, hence the g_bits test)
    geo.lmin = (geo.T < geo.n_local || g_bits > 0) ? std::max(0, std::min(opt.min_low_bits, geo.T - 2)) : geo.T;
    if (g_bits > 0 && geo.T - geo.lmin < 2)
        throw std::runtime_error("shard too small for the requested number of ranks");
    const uint64_t lowmask = (1ull << geo.lmin) - 1;
    const uint64_t local_mask = (1ull << geo.n_local) - 1ull;
    const int cap_high = geo.T - geo.lmin;
    Layout lay{&w2p};

    // Swap routing (single device, fused): exact SWAP gates are absorbed as relabelings and the physical permutation that brings
    // the state back to the canonical layout (logical qubit q on physical bit q) is executed for free by the write-back of the
    // gate passes: a pass may permute the index bits inside its tile, so a wire whose final position lies in the tile goes there,
    // and a pass reserves a tile slot for the final position of a wire it mixes while there is room.  What is left at the end
    // (if anything) takes permutation-only passes.  The 30-qubit QFT runs in 5 passes this way instead of 4 + 2.
    const bool routing = opt.fuse && opt.route_swaps > 0 && g_bits == 0 && !opt.absorb_swaps && opt.store_perm;

    // 1. analyse gates into atoms (wire space)
    std::vector<Atom> atoms;
    std::vector<int> gate_of;
    for (size_t gi = 0; gi < gates.size(); gi++) {
        const Gate& g = gates[gi];
        for (int q : g.qubits)
            if (q < 0 || q >= n_bits) throw std::runtime_error("gate qubit out of range");
        if ((opt.absorb_swaps || routing) && is_exact_swap(g)) {
            std::swap(wire_of[g.qubits[0]], wire_of[g.qubits[1]]);
            continue;
        }
        const size_t before = atoms.size();
        analyze(g, wire_of, atoms);
        for (size_t i = before; i < atoms.size(); i++) gate_of.push_back((int)gi);
    }
    if (opt.fuse && opt.fuse_matrices) tape.n_fused = fuse_matrices(atoms);
    if (opt.fuse && opt.euler_split > 0) tape.n_split = euler_split(atoms);
    else if (opt.fuse) {
        Atom parts[3];
        for (const Atom& a : atoms)
            if (euler_split_atom(a, parts)) tape.n_splittable++;
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
        if (cand.size() < globals.size()) {
            // tiny shards: fall back to the always-resident low bits as well (correct, only shorter contiguous runs)
            std::vector<int> low;
            for (int wq = 0; wq < n_bits; wq++)
                if (w2p[wq] < geo.lmin && !(need_wires >> wq & 1)) low.push_back(wq);
            std::stable_sort(low.begin(), low.end(), [&](int a, int b) {
                if (next_use[a] != next_use[b]) return next_use[a] > next_use[b];
                return w2p[a] > w2p[b];
            });
            cand.insert(cand.end(), low.begin(), low.end());
        }
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

    // ---- swap routing helpers
    std::vector<int> dest(n_bits, -1);      // final physical bit of every wire: logical qubit q ends on bit q
    if (routing)
        for (int q = 0; q < n_bits; q++) dest[wire_of[q]] = q;
    std::deque<Atom> route_atoms;           // synthesised SWAPs (addresses must stay valid until the pass is built)
    auto wire_at = [&]() {
        std::vector<int> p2w(n_bits, -1);
        for (int wq = 0; wq < n_bits; wq++) p2w[w2p[wq]] = wq;
        return p2w;
    };
    // Spend the spare slots of a tile (bit set tb) on final positions: first those of wires already inside the tile (following
    // the chain of displaced wires; a slot that brings two wires home goes first), then -- seeds -- whole 2-cycles outside it.
    auto add_route_slots = [&](uint64_t tb, bool seeds) {
        const std::vector<int> p2w = wire_at();
        auto spare = [&]() { return cap_high - popc(tb & ~lowmask); };
        for (;;) {
            int best = -1, best_score = 0;
            for (int p = 0; p < n_bits && spare() > 0; p++) {
                if (!(tb >> p & 1)) continue;
                const int d = dest[p2w[p]];
                if (tb >> d & 1) continue;
                const int score = (tb >> dest[p2w[d]] & 1) ? 2 : 1;
                if (score > best_score) {
                    best_score = score;
                    best = d;
                }
            }
            if (best >= 0) {
                tb |= 1ull << best;
                continue;
            }
            if (!seeds || spare() < 2) break;
            int seed = -1;
            for (int p = 0; p < n_bits && seed < 0; p++)
                if (!(tb >> p & 1) && dest[p2w[p]] != p) seed = p;
            if (seed < 0) break;
            tb |= 1ull << seed;
        }
        for (int b = 0; b < geo.n_local && popc(tb) < geo.T; b++) tb |= 1ull << b;    // pad like build_tile_step
        return tb;
    };
    // The write-back permutation of a pass whose tile holds the bits tb: every wire inside the tile whose final position is
    // inside too goes there; a wire that loses its place takes the lowest free one (a wire on an always-resident low bit is
    // inside every later tile).  Appends the SWAP atoms that realise it to `out` and returns the new position of every wire.
    auto route_in_tile = [&](uint64_t tb, std::vector<const Atom*>& out) {
        const std::vector<int> p2w = wire_at();
        std::vector<int> newpos(w2p), ends_at(n_bits, -1);
        std::vector<int> pos;
        for (int b = 0; b < n_bits; b++)
            if (tb >> b & 1) pos.push_back(b);
        std::vector<char> seated(n_bits, 0);
        for (int p : pos) {
            const int wq = p2w[p];
            if (tb >> dest[wq] & 1) {
                ends_at[dest[wq]] = wq;
                seated[wq] = 1;
            }
        }
        for (int p : pos)
            if (!seated[p2w[p]] && ends_at[p] < 0) {
                ends_at[p] = p2w[p];
                seated[p2w[p]] = 1;
            }
        for (int p : pos) {
            const int wq = p2w[p];
            if (seated[wq]) continue;
            for (int f : pos)
                if (ends_at[f] < 0) {
                    ends_at[f] = wq;
                    seated[wq] = 1;
                    break;
                }
        }
        // transpositions of tile positions, applied in order: afterwards position q holds what was at w2p[ends_at[q]]
        std::vector<int> cur(n_bits);
        for (int b = 0; b < n_bits; b++) cur[b] = b;
        for (int q : pos) {
            const int want = w2p[ends_at[q]];
            if (cur[q] == want) continue;
            int q2 = -1;
            for (int x : pos)
                if (cur[x] == want) q2 = x;
            Atom a = swap_proto;
            a.tw = {p2w[q], p2w[q2]};
            a.mix = a.touch = (1ull << p2w[q]) | (1ull << p2w[q2]);
            route_atoms.push_back(std::move(a));
            out.push_back(&route_atoms.back());
            std::swap(cur[q], cur[q2]);
        }
        for (int q : pos) newpos[ends_at[q]] = q;
        return newpos;
    };

    // Sharded states never return to the canonical layout (the host maps indices through the layout), so an exact
    // SWAP gate with an operand on a rank bit is a pure relabeling: the two wires exchange their physical bits
    // and nothing crosses NVLink.  Only the atom at the head of the queue may do this (every earlier atom has been
    // scheduled with the old mapping).
    auto relabels = [&](const Atom& a) {
        if (!opt.relabel_global_swaps || g_bits == 0) return false;
        if (!(a.kind == Atom::DENSE && a.tw.size() == 2 && a.cmask == 0 && is_linear_perm(a))) return false;
        return w2p[a.tw[0]] >= geo.n_local || w2p[a.tw[1]] >= geo.n_local;
    };

    if (!opt.fuse) {
        // every gate is its own pass (its controlled blocks share it when they fit)
        size_t i = 0;
        while (i < atoms.size()) {
            size_t j = i;
            while (j < atoms.size() && gate_of[j] == gate_of[i]) j++;
            if (j == i + 1 && relabels(atoms[i])) {
                std::swap(w2p[atoms[i].tw[0]], w2p[atoms[i].tw[1]]);
                tape.n_relabeled++;
                i = j;
                continue;
            }
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
        fuse_pulls(tape, opt);
        return tape;
    }

    // One pass: the atoms of `pend` (in order) that fit a tile together; an atom that does not fit is deferred together with
    // everything behind it that does not commute with it.
    auto form_pass = [&](const std::vector<const Atom*>& pend, std::vector<const Atom*>& in_pass, std::vector<const Atom*>& deferred,
                         uint64_t& targets) {
        uint64_t dmix = 0, dtouch = 0;
        size_t est_bytes = sizeof(QvPassHeader);
        size_t n_diag = 0;
        for (const Atom* a : pend) {
            bool blocked = (a->mix & dtouch) != 0 || (dmix & a->touch) != 0;
            if (!blocked && a->kind == Atom::DIAG && n_diag + 1 > 4096) blocked = true;
            // rough size of the atom in the control program; diagonals merge into shared chunks, so they
            // are cheap -- the real size is checked when the pass is built (see the retry below)
            const size_t need_bytes = a->kind == Atom::DENSE ? sizeof(QvUop) + a->mat.size() * sizeof(cd) + 32 : 16;
            if (!blocked && est_bytes + need_bytes > QV_PROG_LARGE_BYTES - 2048) blocked = true;
            if (!blocked) {
                if (a->kind == Atom::BIG) blocked = true;
                else if (a->kind == Atom::DENSE) {
                    const uint64_t pm = phys_mask(a->mix);
                    uint64_t em = pm;       // routing: with the final positions of the wires it mixes, while there is room
                    if (routing)
                        for (int wq = 0; wq < n_bits; wq++)
                            if (a->mix >> wq & 1) em |= 1ull << dest[wq];
                    if (fits(targets | em)) targets |= em;
                    else if (fits(targets | pm)) targets |= pm;
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
    };
    // wires to bring home in one exchange: `need` plus the other rank-bit targets of the queued atoms, while there is room
    auto globals_wanted = [&](uint64_t need) {
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
        return need;
    };
    const bool hoisting = opt.hoist_remaps > 0 && g_bits > 0 && opt.remap_pull && opt.fuse_pull;

    // 2. greedy pass formation with commutation look-ahead.
    while (!pending.empty()) {
        if (relabels(*pending.front())) {
            const Atom& a = *pending.front();
            std::swap(w2p[a.tw[0]], w2p[a.tw[1]]);
            tape.n_relabeled++;
            pending.erase(pending.begin());
            continue;
        }
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
        uint64_t targets = 0;
        form_pass(pending, in_pass, deferred, targets);
        if (in_pass.empty()) {
            // the front atom needs wires that sit on rank bits: bring them (and the other global
            // targets of the atoms queued right behind it, while there is room) home first.
            const uint64_t need = pending.front()->mix;
            if (!(phys_mask(need) & ~local_mask)) throw std::runtime_error("scheduler bug: no atom fits an empty pass");
            remap(globals_wanted(need), pending);
            continue;
        }
        // Remap hoisting (fused pulls): an exchange pass is bound by NVLink, whatever it computes.  If an atom this pass had to
        // leave behind waits for a wire on a rank bit, try the exchange NOW: when the pass formed after it holds everything this
        // one holds and more, the exchange rides on this pass instead of costing one of its own later.
        if (hoisting && !deferred.empty()) {
            const Atom* trigger = nullptr;
            for (const Atom* a : deferred)
                if (phys_mask(a->mix) & ~local_mask) {
                    trigger = a;
                    break;
                }
            if (trigger) {
                const std::vector<int> w2p_saved = w2p;
                const size_t n_steps = tape.steps.size();
                bool adopted = false;
                try {
                    remap(globals_wanted(trigger->mix), pending);
                    std::vector<const Atom*> in2, def2;
                    uint64_t t2 = 0;
                    form_pass(pending, in2, def2, t2);
                    adopted = in2.size() > in_pass.size() && std::includes(in2.begin(), in2.end(), in_pass.begin(), in_pass.end());
                    if (adopted) {
                        in_pass.swap(in2);
                        deferred.swap(def2);
                        targets = t2;
                        tape.n_hoisted++;
                    }
                } catch (const std::runtime_error&) {
                    adopted = false;
                }
                if (!adopted) {
                    w2p = w2p_saved;
                    tape.steps.resize(n_steps);
                }
            }
        }
        // Build the pass; if its control program or chunk count overflows, keep the first half of the
        // atoms and send the rest back (original order restored: pointers into `atoms` are ordered).
        for (;;) {
            try {
                if (routing) {
                    const uint64_t tb = add_route_slots(targets | lowmask, false);
                    std::vector<const Atom*> with_swaps = in_pass;
                    const std::vector<int> newpos = route_in_tile(tb, with_swaps);
                    tape.steps.push_back(build_tile_step(with_swaps, tb, geo, lay));
                    tape.steps.back().n_gates = (int)in_pass.size();
                    tape.n_routed += (int)(with_swaps.size() - in_pass.size());
                    w2p = newpos;
                } else {
                    tape.steps.push_back(build_tile_step(in_pass, targets, geo, lay));
                }
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
    // routing: whatever is not home yet takes permutation-only passes (every one places at least one wire)
    while (routing) {
        bool home = true;
        for (int wq = 0; wq < n_bits; wq++) home = home && w2p[wq] == dest[wq];
        if (home) break;
        const uint64_t tb = add_route_slots(lowmask, true);
        std::vector<const Atom*> swaps;
        const std::vector<int> newpos = route_in_tile(tb, swaps);
        if (swaps.empty()) throw std::runtime_error("scheduler bug: swap routing makes no progress");
        tape.steps.push_back(build_tile_step(swaps, tb, geo, lay));
        tape.steps.back().n_gates = 0;
        tape.n_routed += (int)swaps.size();
        w2p = newpos;
    }
    tape.l2p.resize(n_bits);
    for (int q = 0; q < n_bits; q++) tape.l2p[q] = w2p[wire_of[q]];
    fuse_pulls(tape, opt);
    return tape;
}

std::string describe(const Tape& t) {
    std::ostringstream os;
    os << "tape: n_bits=" << t.n_bits << " gates=" << t.n_gates << " atoms=" << t.n_atoms
       << " steps=" << t.steps.size() << (t.n_fused ? " matrix_merges=" + std::to_string(t.n_fused) : std::string())
       << (t.n_split ? " euler_splits=" + std::to_string(t.n_split) : std::string()) << (t.n_relabeled ? " relabeled_swaps=" + std::to_string(t.n_relabeled) : std::string())
       << (t.n_hoisted ? " hoisted_remaps=" + std::to_string(t.n_hoisted) : std::string())
       << (t.n_routed ? " routed_transpositions=" + std::to_string(t.n_routed) : std::string()) << "\n";
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
        os << "  [" << i << "] " << (h.pull ? "PULL+TILE" : s.is_remap ? "REMAP" : (s.uses_peers ? "PEER" : "TILE")) << " T=" << h.T
           << " m=" << h.reg_bits << " atoms=" << s.n_gates << " rounds=" << h.n_rounds << " uops=" << h.n_uops << " diag_uops=" << h.n_diag_uops
           << " slices=" << h.n_slices << " sources=" << h.n_sources << " slice_entries=" << h.n_slice_entries
           << (h.store_perm ? " store_perm" : "") << (h.has_scale ? " out_scale" : "") << " bytes=" << h.blob_bytes << " tables=" << h.n_table_entries << " tilebits=";
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
