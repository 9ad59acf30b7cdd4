// tests/support/qv_emulator.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Interprets the tile programs produced by qvm_b200/csrc/qv_sched.cpp on a host
// array, tile by tile and group by group, through the SAME op code
// (qv_ops.h) the CUDA kernel compiles.  It exists so the scheduler and the op
// semantics can be checked against the oracle in the GPU-less build container
// (`pytest -m "not gpu"`).  It is never linked into libqvmcuda and is not a
// fallback: the product fails loudly without a GPU.
#include <cstring>
#include <string>
#include <vector>

#include "../../qvm_b200/csrc/qv_ops.h"
#include "../../qvm_b200/csrc/qv_sched.h"

namespace {

void run_tile_step(qvc* psi, const qv::Step& st) {
    const uint8_t* blob = st.blob.data();
    QvPassHeader h;
    std::memcpy(&h, blob, sizeof(h));
    const QvRound* rounds = (const QvRound*)(blob + h.off_rounds);
    const QvOp* ops = (const QvOp*)(blob + h.off_ops);
    const QvChunk* chunks = (const QvChunk*)(blob + h.off_chunks);
    const qvc* mats = (const qvc*)(blob + h.off_matrices);
    const qvc* tables = (const qvc*)st.tables.data();
    const uint32_t tile_n = 1u << h.T;
    const uint64_t local_mask = (1ull << h.n_local_bits) - 1ull;
    std::vector<qvc> smem(tile_n);
    for (uint64_t tile = 0; tile < h.n_tiles; tile++) {
        const uint64_t base = qv_gather(tile, h.base_segs, h.n_base_segs) | h.fixed_bits;
        for (uint32_t e = 0; e < tile_n; e++) {
            const uint64_t p = base | qv_gather(e, h.tile_segs, h.n_tile_segs);
            smem[qv_swz(e)] = psi[p & local_mask];
        }
        for (uint32_t r = 0; r < h.n_rounds; r++) {
            const QvRound& rd = rounds[r];
            QvRegPos dep{rd.regpos[0], rd.regpos[1], rd.regpos[2]};
            const uint32_t ngroups = tile_n >> rd.m;
            for (uint32_t g = 0; g < ngroups; g++) {
                uint32_t e0 = g;
                for (uint32_t j = 0; j < rd.m; j++) e0 = qv_insert_zero(e0, rd.regpos[j]);
                qvc a[8];
                for (uint32_t s = 0; s < 8; s++) {
                    if (s < (1u << rd.m)) a[s] = smem[qv_swz(e0 | qv_dep(s, dep.p0, dep.p1, dep.p2))];
                    else { a[s].x = 0.0; a[s].y = 0.0; }
                }
                qv_apply_round(a, rd, ops, chunks, mats, tables, e0, dep, base);
                for (uint32_t s = 0; s < (1u << rd.m); s++) smem[qv_swz(e0 | qv_dep(s, dep.p0, dep.p1, dep.p2))] = a[s];
            }
        }
        for (uint32_t e = 0; e < tile_n; e++) {
            const uint64_t p = base | qv_gather(e, h.tile_segs, h.n_tile_segs);
            psi[p & local_mask] = smem[qv_swz(e)];
        }
    }
}

void run_big_step(qvc* psi, int n_bits, const qv::Step& st) {
    const uint32_t k = st.big.k;
    const uint64_t d = 1ull << k;
    uint64_t tmask = 0;
    for (uint32_t j = 0; j < k; j++) tmask |= 1ull << st.big.pos[j];
    std::vector<qvc> in(d), out(d);
    for (uint64_t base = 0; base < (1ull << n_bits); base++) {
        if (base & tmask) continue;
        if ((base & st.big.ctrl_mask) != st.big.ctrl_val) continue;
        for (uint64_t c = 0; c < d; c++) {
            uint64_t a = base;
            for (uint32_t j = 0; j < k; j++)
                if (c >> j & 1) a |= 1ull << st.big.pos[j];
            in[c] = psi[a];
        }
        for (uint64_t r = 0; r < d; r++) {
            qvc acc{0.0, 0.0};
            for (uint64_t c = 0; c < d; c++) {
                qvc m{st.bigmat[r * d + c].real(), st.bigmat[r * d + c].imag()};
                acc = qv_cmadd(acc, m, in[c]);
            }
            out[r] = acc;
        }
        for (uint64_t c = 0; c < d; c++) {
            uint64_t a = base;
            for (uint32_t j = 0; j < k; j++)
                if (c >> j & 1) a |= 1ull << st.big.pos[j];
            psi[a] = out[c];
        }
    }
}

}  // namespace

extern "C" int qvtest_run(double* psi, int n_bits, int n_gates, const int* ks, const int* qubits_flat,
                          const double* mats_flat, int fuse, int tile_bits, int absorb_swaps,
                          int* l2p_inout, char* desc, int desc_len) {
    try {
        std::vector<qv::Gate> gates(n_gates);
        size_t qo = 0, mo = 0;
        for (int g = 0; g < n_gates; g++) {
            const int k = ks[g];
            gates[g].qubits.assign(qubits_flat + qo, qubits_flat + qo + k);
            qo += k;
            const size_t d = (size_t)1 << k;
            gates[g].mat.resize(d * d);
            for (size_t i = 0; i < d * d; i++) gates[g].mat[i] = qv::cd(mats_flat[mo + 2 * i], mats_flat[mo + 2 * i + 1]);
            mo += 2 * d * d;
        }
        qv::CompileOptions opt;
        opt.fuse = fuse != 0;
        opt.tile_bits = tile_bits;
        opt.absorb_swaps = absorb_swaps != 0;
        std::vector<int> l2p;
        if (l2p_inout) l2p.assign(l2p_inout, l2p_inout + n_bits);
        qv::Tape tape = qv::compile(gates, n_bits, opt, l2p);
        for (const qv::Step& st : tape.steps) {
            if (st.kind == qv::Step::TILE) run_tile_step((qvc*)psi, st);
            else run_big_step((qvc*)psi, n_bits, st);
        }
        if (l2p_inout) std::memcpy(l2p_inout, tape.l2p.data(), sizeof(int) * n_bits);
        if (desc && desc_len > 0) {
            std::string s = qv::describe(tape);
            std::strncpy(desc, s.c_str(), desc_len - 1);
            desc[desc_len - 1] = 0;
        }
        return (int)tape.steps.size();
    } catch (const std::exception& e) {
        if (desc && desc_len > 0) {
            std::strncpy(desc, e.what(), desc_len - 1);
            desc[desc_len - 1] = 0;
        }
        return -1;
    }
}
