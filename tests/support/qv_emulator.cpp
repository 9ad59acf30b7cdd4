// tests/support/qv_emulator.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Interprets the tile programs produced by qvm_b200/csrc/qv_sched.cpp on a host
// array, tile by tile and group by group, through the SAME op code
// (qv_ops.h) the CUDA kernel compiles.  It exists so the scheduler and the op
// semantics can be checked against the oracle in the GPU-less build container
// (`pytest -m "not gpu"`).  It is never linked into libqvmcuda and is not a
// fallback: the product fails loudly without a GPU.
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <dlfcn.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <map>

#include "../../qvm_b200/csrc/qv_jit.h"
#include "../../qvm_b200/csrc/qv_ops.h"
#include "../../qvm_b200/csrc/qv_sched.h"

namespace {

// ---- compiled passes on the host: the text the pass compiler (qv_jit_gen.cpp) would hand to NVRTC is compiled with
// g++ (-DQVJ_HOST) and its round functions replace the interpreter loop below.  This checks the generator -- literal
// offsets, index expressions, slot addressing, template arguments -- against the oracle without a GPU.
bool g_jit_host = false;
int g_jit_host_used = 0;
std::string g_csrc_dir;
typedef void (*qvj_round_fn)(int, qvc*, uint32_t, const uint8_t*, const qvc*, const qvc*, const uint8_t*);
typedef uint32_t (*qvj_slot_fn)(uint32_t);
struct HostPass {
    qvj_round_fn round = nullptr;
    qvj_slot_fn slot = nullptr;       // default-layout slot -> slot of the pass's own shared-memory swizzle
};
std::map<uint64_t, HostPass> g_jit_fns;

HostPass host_compiled_rounds(const qv::Step& st) {
    const qv::JitSource src = qv::jit_generate(st);
    if (!src.ok) return HostPass();
    auto it = g_jit_fns.find(src.sig);
    if (it != g_jit_fns.end()) return it->second;
    char base[128];
    snprintf(base, sizeof(base), "/tmp/qvj_host_%ld_%016llx", (long)getpid(), (unsigned long long)src.sig);
    const std::string cu = std::string(base) + ".cpp", so = std::string(base) + ".so";
    FILE* f = fopen(cu.c_str(), "w");
    if (!f) throw std::runtime_error("emulator: cannot write the generated pass");
    fwrite(src.text.data(), 1, src.text.size(), f);
    fclose(f);
    const std::string cmd = "/usr/bin/g++ -O1 -std=c++17 -fPIC -shared -DQVJ_HOST -Wno-unknown-pragmas -I" + g_csrc_dir + " -o " + so + " " + cu + " 2>" + base + ".log";
    if (system(cmd.c_str()) != 0) throw std::runtime_error("emulator: g++ failed on a generated pass, see " + std::string(base) + ".log");
    void* lib = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) throw std::runtime_error(std::string("emulator: dlopen of a generated pass: ") + dlerror());
    HostPass fn;
    fn.round = (qvj_round_fn)dlsym(lib, "qvj_host_round");
    fn.slot = (qvj_slot_fn)dlsym(lib, "qvj_host_slot");
    if (!fn.round || !fn.slot) throw std::runtime_error("emulator: generated pass lacks qvj_host_round / qvj_host_slot");
    unlink(cu.c_str());
    unlink(so.c_str());
    unlink((std::string(base) + ".log").c_str());
    g_jit_fns[src.sig] = fn;
    return fn;
}

// peers[r] = base pointer of rank r's shard (peers[0] for a single device)
// pull passes (header.pull): loads go through the remap from the current buffers, stores into alt_own
void run_tile_step(qvc* const* peers, const qv::Step& st, qvc* alt_own = nullptr) {
    const uint8_t* blob = st.blob.data();
    QvPassHeader h;
    std::memcpy(&h, blob, sizeof(h));
    const QvRound* rounds = (const QvRound*)(blob + h.off_rounds);
    const QvUop* uops = (const QvUop*)(blob + h.off_uops);
    const QvSource* sources = (const QvSource*)(blob + h.off_sources);
    const QvSlice* slices = (const QvSlice*)(blob + h.off_slices);
    const uint8_t* slice_of = blob + h.off_slice_of;
    const QvPred* preds = (const QvPred*)(blob + h.off_preds);
    const qvc* tables = (const qvc*)st.tables.data();
    const uint32_t tile_n = 1u << h.T;
    const uint32_t threads = 1u << h.threads_log2;
    const uint64_t local_mask = (1ull << h.n_local_bits) - 1ull;
    std::vector<qvc> smem(tile_n);
    std::vector<uint32_t> s_srcext(QV_MAX_SOURCES);
    std::vector<uint8_t> s_pred(QV_MAX_PREDS);
    std::vector<qvc> s_slice(QV_SLICE_ENTRIES);
    if (h.n_sources > QV_MAX_SOURCES || h.n_preds > QV_MAX_PREDS || h.n_slice_entries > QV_SLICE_ENTRIES ||
        h.n_slices > QV_MAX_SLICES)
        throw std::runtime_error("emulator: per-tile table limits exceeded");
    HostPass hp;
    if (g_jit_host && h.n_rounds > 0) {
        hp = host_compiled_rounds(st);
        if (hp.round) g_jit_host_used++;
    }
    const qvj_round_fn compiled = hp.round;
    auto slot_of = [&](uint32_t s1) { return hp.slot ? hp.slot(s1) : s1; };
    auto addr = [&](uint64_t p) { return peers[p >> h.n_local_bits] + (p & local_mask); };
    if (h.pull && !alt_own) throw std::runtime_error("emulator: pull pass without an alternate buffer");
    // the kernel's split source index: S(base | gather(tid)) ^ hi_src[i]
    auto src_index = [&](uint64_t base, uint32_t e) {
        const uint64_t P = base | qv_gather(e & (threads - 1), h.tile_segs, h.n_tile_segs);
        uint64_t S = P;
        for (uint32_t i = 0; i < h.pull_remap.n_pairs; i++) {
            const uint64_t x = ((P >> h.pull_remap.local_bit[i]) ^ (P >> h.pull_remap.global_bit[i])) & 1ull;
            S ^= (x << h.pull_remap.local_bit[i]) | (x << h.pull_remap.global_bit[i]);
        }
        return S ^ h.hi_src[e >> h.threads_log2];
    };
    // the split address computation of the kernel: (base | gather(tid)) | hi_off[i] for e = tid + threads*i
    auto phys = [&](uint64_t base, uint32_t e) {
        return base | qv_gather(e & (threads - 1), h.tile_segs, h.n_tile_segs) | h.hi_off[e >> h.threads_log2];
    };
    for (uint64_t tile = 0; tile < h.n_tiles; tile++) {
        const uint64_t base = qv_gather(tile, h.base_segs, h.n_base_segs) | h.fixed_bits;
        for (uint32_t i = 0; i < h.n_sources; i++)
            s_srcext[i] = (uint32_t)qv_gather(base, sources[i].esegs, sources[i].n_esegs) << sources[i].nl;
        for (uint32_t i = 0; i < h.n_preds; i++) s_pred[i] = (base & preds[i].mask) == preds[i].val ? 1 : 0;
        for (uint32_t f = 0; f < h.n_slice_entries; f++) {
            const QvSlice& sl = slices[slice_of[f]];
            s_slice[f] = qv_slice_entry(sl, sources, s_srcext.data(), tables, f - sl.off);
        }
        for (uint32_t e = 0; e < tile_n; e++) smem[slot_of(qv_swz(e))] = h.pull ? *addr(src_index(base, e)) : *addr(phys(base, e));
        if (compiled) {     // the generated round functions, one virtual thread after the other, round by round
            for (uint32_t r = 0; r < h.n_rounds; r++)
                for (uint32_t tid = 0; tid < threads; tid++)
                    compiled((int)r, smem.data(), tid, blob, tables, s_slice.data(), s_pred.data());
        } else
        for (uint32_t r = 0; r < h.n_rounds; r++) {
            const QvRound& rd = rounds[r];
            if (rd.m > h.reg_bits) throw std::runtime_error("emulator: round uses more register bits than the pass declares");
            if (uops[rd.first_uop + rd.n_uops].kind != QV_K_END) throw std::runtime_error("emulator: round without an end sentinel");
            const uint32_t ngroups = tile_n >> rd.m;
            for (uint32_t g = 0; g < ngroups; g++) {
                uint32_t e0 = g;
                for (uint32_t j = 0; j < rd.m; j++) e0 = qv_insert_zero(e0, rd.regpos[j]);
                const uint32_t se0 = qv_swz(e0);
                if (h.reg_bits <= 3) {      // the 8-slot instantiation, as the 256-thread kernel
                    qvc a[8];
                    for (uint32_t s = 0; s < 8; s++) {
                        if (s < (1u << rd.m)) a[s] = smem[se0 ^ rd.slot_xor[s]];
                        else { a[s].x = 0.0; a[s].y = 0.0; }
                    }
                    for (uint32_t u = rd.first_uop; uops[u].kind != QV_K_END; u++) {   // sentinel-terminated, as in the kernel
                        QvUopHead hd;
                        std::memcpy(&hd, &uops[u], sizeof(hd));
                        qv_run_uop<8>(a, hd, uops[u], g, blob, tables, s_slice.data(), s_pred.data());
                    }
                    for (uint32_t s = 0; s < (1u << rd.m); s++) smem[se0 ^ rd.slot_xor[s]] = a[s];
                } else {                    // the 16-slot instantiation, as the 128-thread kernel
                    qvc a[16];
                    for (uint32_t s = 0; s < 16; s++) {
                        if (s < (1u << rd.m)) a[s] = smem[se0 ^ rd.slot_xor[s]];
                        else { a[s].x = 0.0; a[s].y = 0.0; }
                    }
                    for (uint32_t u = rd.first_uop; uops[u].kind != QV_K_END; u++) {   // sentinel-terminated, as in the kernel
                        QvUopHead hd;
                        std::memcpy(&hd, &uops[u], sizeof(hd));
                        qv_run_uop<16>(a, hd, uops[u], g, blob, tables, s_slice.data(), s_pred.data());
                    }
                    for (uint32_t s = 0; s < (1u << rd.m); s++) smem[se0 ^ rd.slot_xor[s]] = a[s];
                }
            }
        }
        for (uint32_t e = 0; e < tile_n; e++) {
            uint32_t slot = qv_swz(e);
            if (h.store_perm) {     // as the kernel: st_lo(tid) ^ st_hi[i] for e = tid + threads*i
                slot = h.st_const;
                for (uint32_t k = 0; k < h.T; k++)
                    if ((e & (threads - 1)) >> k & 1) slot ^= h.st_col[k];
                slot ^= h.st_hi[e >> h.threads_log2];
            }
            qvc v = smem[slot_of(slot)];
            if (h.has_scale) {
                v.x *= h.out_scale;
                v.y *= h.out_scale;
            }
            if (h.pull) alt_own[phys(base, e) & local_mask] = v;
            else *addr(phys(base, e)) = v;
        }
    }
}

void run_big_step(qvc* psi, int n_bits, const qv::Step& st) {   // psi = this rank's shard, n_bits = its log2 size
    const uint32_t k = st.big.k;
    const uint64_t d = 1ull << k;
    if (st.big.diag) {      // wide diagonal: element-wise table lookup (qv_bigdiag_kernel)
        for (uint64_t i = 0; i < (1ull << n_bits); i++) {
            const uint64_t full = i | st.big.fixed_bits;
            uint32_t idx = 0;
            for (uint32_t j = 0; j < k; j++) idx |= (uint32_t)((full >> st.big.pos[j]) & 1ull) << j;
            const qvc t{st.bigmat[idx].real(), st.bigmat[idx].imag()};
            psi[i] = qv_cmul(psi[i], t);
        }
        return;
    }
    uint64_t tmask = 0;
    for (uint32_t j = 0; j < k; j++) tmask |= 1ull << st.big.pos[j];
    std::vector<qvc> in(d), out(d);
    for (uint64_t base = 0; base < (1ull << n_bits); base++) {
        if (base & tmask) continue;
        if (((base | st.big.fixed_bits) & st.big.ctrl_mask) != st.big.ctrl_val) continue;
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

// Pull remap (Step::REMAP): dst shard of every rank gathers from all current shards, then the buffers flip.
// Same index algebra as qv_remap_pull_kernel.
void run_remap_step(std::vector<qvc*>& cur, std::vector<qvc*>& alt, int n_local, const qv::Step& st_of_rank0) {
    const QvRemap& rm = st_of_rank0.remap;
    const uint64_t n = 1ull << n_local, local_mask = n - 1ull;
    for (size_t r = 0; r < cur.size(); r++) {
        for (uint64_t p = 0; p < n; p++) {
            const uint64_t P = ((uint64_t)r << n_local) | p;
            uint64_t S = P;
            for (uint32_t i = 0; i < rm.n_pairs; i++) {
                const uint64_t x = ((P >> rm.local_bit[i]) ^ (P >> rm.global_bit[i])) & 1ull;
                S ^= (x << rm.local_bit[i]) | (x << rm.global_bit[i]);
            }
            alt[r][p] = cur[S >> n_local][S & local_mask];
        }
    }
    std::swap(cur, alt);
}

bool g_remap_pull = false;
int g_reg_bits = 0;

}  // namespace

extern "C" void qvtest_set_remap_pull(int on) { g_remap_pull = on != 0; }
extern "C" void qvtest_set_reg_bits(int m) { g_reg_bits = m; }
int g_route_swaps = -1;
extern "C" void qvtest_set_route_swaps(int m) { g_route_swaps = m; }
// compiled-pass mode: csrc_dir = where qv_jit_prelude.cuh & co. live; returns the number of passes run compiled so far
extern "C" int qvtest_set_jit_host(int on, const char* csrc_dir) {
    g_jit_host = on != 0;
    if (csrc_dir) g_csrc_dir = csrc_dir;
    return g_jit_host_used;
}

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
        opt.reg_bits = g_reg_bits;
        opt.route_swaps = g_route_swaps;
        std::vector<int> l2p;
        if (l2p_inout) l2p.assign(l2p_inout, l2p_inout + n_bits);
        qv::Tape tape = qv::compile(gates, n_bits, opt, l2p);
        for (const qv::Step& st : tape.steps) {
            qvc* peers[1] = {(qvc*)psi};
            if (st.kind == qv::Step::TILE) run_tile_step(peers, st);
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

// Emulates `world` ranks in one process: psi holds the shards back to back (rank r at r << n_local).
// Every rank compiles its own tape (the schedules must agree step for step); step i of all ranks
// runs before step i+1 of any rank, which is what the barriers around peer passes guarantee on GPUs.
extern "C" int qvtest_run_sharded(double* psi, int n_bits, int world, int n_gates, const int* ks,
                                  const int* qubits_flat, const double* mats_flat, int fuse, int tile_bits,
                                  int absorb_swaps, int* l2p_inout, char* desc, int desc_len) {
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
        int gb = 0;
        while ((1 << gb) < world) gb++;
        const int n_local = n_bits - gb;
        std::vector<int> l2p;
        if (l2p_inout) l2p.assign(l2p_inout, l2p_inout + n_bits);
        std::vector<qv::Tape> tapes(world);
        for (int r = 0; r < world; r++) {
            qv::CompileOptions opt;
            opt.fuse = fuse != 0;
            opt.tile_bits = tile_bits;
            opt.absorb_swaps = absorb_swaps != 0;
            opt.reg_bits = g_reg_bits;
            opt.n_local_bits = n_local;
            opt.rank = r;
            opt.remap_pull = g_remap_pull;
            tapes[r] = qv::compile(gates, n_bits, opt, l2p);
            if (tapes[r].steps.size() != tapes[0].steps.size() || tapes[r].l2p != tapes[0].l2p)
                throw std::runtime_error("ranks disagree on the schedule");
        }
        std::vector<qvc*> peers(world);
        for (int r = 0; r < world; r++) peers[r] = (qvc*)psi + ((size_t)r << n_local);
        std::vector<qvc> altbuf(g_remap_pull ? ((size_t)1 << n_bits) : 0);
        std::vector<qvc*> alts(world);
        for (int r = 0; r < world; r++) alts[r] = altbuf.data() + (g_remap_pull ? ((size_t)r << n_local) : 0);
        int peer_steps = 0;
        for (size_t i = 0; i < tapes[0].steps.size(); i++) {
            if (tapes[0].steps[i].kind == qv::Step::REMAP) {
                for (int r = 0; r < world; r++)
                    if (tapes[r].steps[i].kind != qv::Step::REMAP || tapes[r].steps[i].remap.rank != (uint32_t)r)
                        throw std::runtime_error("ranks disagree on a remap step");
                run_remap_step(peers, alts, n_local, tapes[0].steps[i]);
                peer_steps++;
                continue;
            }
            bool pull = false;
            for (int r = 0; r < world; r++) {
                const qv::Step& st = tapes[r].steps[i];
                if (st.kind != tapes[0].steps[i].kind || st.uses_peers != tapes[0].steps[i].uses_peers)
                    throw std::runtime_error("ranks disagree on a step");
                if (st.kind == qv::Step::TILE) {
                    QvPassHeader hh;
                    std::memcpy(&hh, st.blob.data(), sizeof(hh));
                    if (r == 0) pull = hh.pull != 0;
                    else if (pull != (hh.pull != 0)) throw std::runtime_error("ranks disagree on a pull pass");
                    run_tile_step(peers.data(), st, pull ? alts[r] : nullptr);
                } else run_big_step(peers[r], n_local, st);
            }
            if (pull) std::swap(peers, alts);   // every rank flips after the same step
            if (tapes[0].steps[i].uses_peers) peer_steps++;
        }
        if (peers[0] != (qvc*)psi)   // an odd number of flips: the result sits in the alternate buffers
            std::memcpy(psi, altbuf.data(), sizeof(qvc) << n_bits);
        if (l2p_inout) std::memcpy(l2p_inout, tapes[0].l2p.data(), sizeof(int) * n_bits);
        if (desc && desc_len > 0) {
            std::string s = qv::describe(tapes[0]);
            std::strncpy(desc, s.c_str(), desc_len - 1);
            desc[desc_len - 1] = 0;
        }
        return (int)tapes[0].steps.size() * 1000 + peer_steps;
    } catch (const std::exception& e) {
        if (desc && desc_len > 0) {
            std::strncpy(desc, e.what(), desc_len - 1);
            desc[desc_len - 1] = 0;
        }
        return -1;
    }
}

// ---- step-wise sharded emulation for the gloo world-size-2 tests: each PROCESS owns one rank's tape and
// runs its own steps on shards that live in POSIX shared memory (standing in for NVLink peer access).
struct EmuTape {
    qv::Tape tape;
    int n_local = 0;
};

extern "C" void* qvtest_shard_compile(int n_bits, int world, int rank, int n_gates, const int* ks, const int* qubits_flat,
                                      const double* mats_flat, int fuse, int tile_bits, int absorb_swaps,
                                      const int* l2p_in, char* err, int errlen) {
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
        int gb = 0;
        while ((1 << gb) < world) gb++;
        qv::CompileOptions opt;
        opt.fuse = fuse != 0;
        opt.tile_bits = tile_bits;
        opt.absorb_swaps = absorb_swaps != 0;
        opt.reg_bits = g_reg_bits;
        opt.n_local_bits = n_bits - gb;
        opt.rank = rank;
        opt.remap_pull = g_remap_pull;      // schedule inspection only: the step-wise engine runs in-place schedules
        std::vector<int> l2p(l2p_in, l2p_in + n_bits);
        EmuTape* t = new EmuTape();
        t->n_local = n_bits - gb;
        t->tape = qv::compile(gates, n_bits, opt, l2p);
        return t;
    } catch (const std::exception& e) {
        if (err && errlen > 0) {
            std::strncpy(err, e.what(), errlen - 1);
            err[errlen - 1] = 0;
        }
        return nullptr;
    }
}
extern "C" int qvtest_shard_num_steps(void* t) { return (int)((EmuTape*)t)->tape.steps.size(); }
extern "C" int qvtest_shard_step_flags(void* t, int i) {
    const qv::Step& st = ((EmuTape*)t)->tape.steps[i];
    return (st.uses_peers ? 1 : 0) | (st.is_remap ? 2 : 0);
}
extern "C" int qvtest_shard_run_step(void* t, int i, double** peers, int rank) {
    EmuTape* et = (EmuTape*)t;
    const qv::Step& st = et->tape.steps[i];
    if (st.kind == qv::Step::TILE) run_tile_step((qvc* const*)peers, st);
    else run_big_step((qvc*)peers[rank], et->n_local, st);
    return 0;
}
extern "C" void qvtest_shard_l2p(void* t, int* out) {
    EmuTape* et = (EmuTape*)t;
    for (size_t i = 0; i < et->tape.l2p.size(); i++) out[i] = et->tape.l2p[i];
}
extern "C" void qvtest_shard_describe(void* t, char* out, int len) {
    const std::string s = qv::describe(((EmuTape*)t)->tape);
    std::strncpy(out, s.c_str(), len - 1);
    out[len - 1] = 0;
}
extern "C" void qvtest_shard_free(void* t) { delete (EmuTape*)t; }
