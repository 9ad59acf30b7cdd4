// qvmcuda.cu -- C ABI of libqvmcuda (see include/qvmcuda.h for the contract and
// the reference interfaces each entry point stands behind).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qvmcuda.h"
#include "qv_kernels.cuh"
#include "qv_tile_launch.h"
#include "qv_sched.h"
#include "qv_jit.h"

namespace {

thread_local std::string g_err;
std::atomic<uint64_t> g_launches{0};

int fail(const std::string& msg) {
    g_err = msg;
    return 1;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(std::string(#call) + ": " + cudaGetErrorString(e_));                  \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
        active = dev;
    }
    ~DeviceGuard() {
        if (prev >= 0 && prev != active) cudaSetDevice(prev);
    }
    int active = -1;
};

int log2_exact(uint64_t n) {
    if (n == 0 || (n & (n - 1))) return -1;
    int b = 0;
    while ((1ull << b) < n) b++;
    return b;
}

constexpr uint32_t kReduceBlocks = 148 * 8;
constexpr uint64_t kLazyAllZero = ~0ull;

}  // namespace

struct qvmcuda_tape;
// Immediate-mode schedule cache: the reference compiles a loaded program once and runs it many times (multishot loops,
// COMPILE-LOADED-PROGRAM src/qvm.lisp:166-175); qvmcuda_apply_gates sees the same gate list over and over in that case.  The
// last few schedules are kept per state, keyed by the exact gate list, flags and starting layout.
struct QvTapeCacheEntry {
    uint64_t key = 0;
    uint32_t flags = 0;
    std::vector<int> l2p_in;
    std::vector<qv::Gate> gates;
    qvmcuda_tape* tape = nullptr;
};

struct qvmcuda_state {
    std::mutex mu;
    int device = 0;
    uint64_t n_amps = 0;
    int n_bits = 0;
    qvc* d_amps = nullptr;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    double* d_partial = nullptr;   // kReduceBlocks + 1 doubles
    int sm_count = 148;
    std::vector<int> l2p;          // logical -> physical qubit (identity unless swaps were absorbed)
    // Lazy reset (single device): SET-TO-ZERO-STATE only records that the state is the basis vector |lazy_index>.  If the next
    // thing that happens is a compiled gate pass, that pass synthesises its tiles instead of loading them (no 16 B/amplitude
    // reset write, no 16 B/amplitude read in the first pass); anything else writes the vector first (materialize_locked).
    bool lazy_basis = false;
    uint64_t lazy_index = 0;       // kLazyAllZero: the vector holds only zeros (a shard of a freshly reset sharded state)
    // sharded: ranks (bit r) whose CURRENT buffer is known to hold only zeros -- set by the host right after a collective reset
    // (qvmcuda_shard_set_zero_ranks), cleared by the first exchange step.  Pull passes do not fetch those ranks' amplitudes.
    uint32_t zero_ranks = 0;
    std::vector<QvTapeCacheEntry> tape_cache;      // most recent first
    // multi-GPU
    int rank = 0, world = 1;
    QvPeers peers{};
    std::vector<void*> opened;     // IPC-opened peer pointers
    // alternate shard buffer for pull remaps (out of place, then every rank flips)
    qvc* d_alt = nullptr;
    QvPeers peers_alt{};
    bool remap_pull = false;       // every rank has attached every rank's alternate buffer
    // sampler scratch: block sums (two levels) + top prefix; per-shot uniforms / results
    double* d_sample_tree = nullptr;
    void* d_shots = nullptr;
    uint64_t shot_cap = 0;
    // export scratch (probabilities, diagonal of rho): persistent, grown on demand -- no allocation per call
    double* d_aux = nullptr;
    uint64_t aux_cap = 0;
    // immediate-mode program upload
    uint8_t* d_scratch = nullptr;
    uint8_t* h_scratch = nullptr;  // pinned
    size_t scratch_cap = 0;
    uint64_t scratch_tape = 0;     // id of the tape whose step data d_scratch holds
    cudaEvent_t upload_done = nullptr;
};

static uint64_t next_tape_id() {
    static std::atomic<uint64_t> n{1};
    return n++;
}

struct qvmcuda_tape {
    std::mutex mu;
    qv::Tape tape;
    uint32_t flags = 0;
    std::map<int, uint8_t*> d_blobs;      // device -> one buffer holding every step's blob / matrix
    std::vector<size_t> offsets;          // per step offset into the buffer
    size_t total_bytes = 0;
    int n_local = 0, rank = 0, world = 1; // geometry the tape was compiled for
    uint64_t id = next_tape_id();         // identifies the tape whose data sits in a state's scratch buffer
    bool ephemeral = false;               // shard tapes are compiled per run: their data goes through the state's scratch
    bool state_owned = false;             // lives in a state's schedule cache: qvmcuda_tape_destroy only returns it
    std::atomic<int> checked_out{0};      // handles given out by qvmcuda_shard_compile and not yet destroyed
    // compiled passes (qv_jit.h) per device: jit[dev][step] = kernel or nullptr (interpreter); complete = no step is
    // still waiting for the asynchronous compiler
    std::map<int, std::vector<qv::JitKernel*>> jit;
    std::map<int, bool> jit_complete;
    // no-load variant of the first pass (lazy reset), per device; looked up once
    std::map<int, qv::JitKernel*> jit_basis;
};

namespace {

bool l2p_is_identity(const std::vector<int>& l2p) {
    for (size_t i = 0; i < l2p.size(); i++)
        if (l2p[i] != (int)i) return false;
    return true;
}

// Dense gates on 3..8 mixing qubits run on the FP64 tensor path (qv_bigmma_kernel); QVMCUDA_BIG=scalar keeps the
// scalar kernel (A/B measurements, and the kernel for k > 8).
bool big_uses_mma(const qv::Step& st) {
    static const bool scalar = getenv("QVMCUDA_BIG") && std::string(getenv("QVMCUDA_BIG")) == "scalar";
    return st.kind == qv::Step::BIG && !st.big.diag && st.big.k >= 3 && st.big.k <= 8 && !scalar;
}

// host bytes of a step's device data: tables (tile pass) or matrix (+ real block form W[2r+p][2c+q], see qv_bigmma_kernel)
void fill_step_data(const qv::Step& st, uint8_t* dst) {
    const std::vector<qv::cd>& src = st.kind == qv::Step::TILE ? st.tables : st.bigmat;
    if (src.empty()) return;
    std::memcpy(dst, src.data(), src.size() * sizeof(qv::cd));
    if (big_uses_mma(st)) {
        const size_t d = (size_t)1 << st.big.k, D2 = 2 * d;
        double* W = reinterpret_cast<double*>(dst + src.size() * sizeof(qv::cd));
        for (size_t r = 0; r < d; r++)
            for (size_t c = 0; c < d; c++) {
                const qv::cd m = st.bigmat[r * d + c];
                W[(2 * r) * D2 + 2 * c] = m.real();
                W[(2 * r) * D2 + 2 * c + 1] = -m.imag();
                W[(2 * r + 1) * D2 + 2 * c] = m.imag();
                W[(2 * r + 1) * D2 + 2 * c + 1] = m.real();
            }
    }
}

void layout_tape(qvmcuda_tape* t) {
    t->offsets.clear();
    size_t off = 0;
    for (const qv::Step& st : t->tape.steps) {
        t->offsets.push_back(off);
        size_t bytes = (st.kind == qv::Step::TILE ? st.tables.size() : st.bigmat.size()) * sizeof(qv::cd);
        if (big_uses_mma(st)) bytes += 4 * bytes;      // + the real (2d x 2d) block form the tensor-path kernel reads
        off += (bytes + 255) & ~(size_t)255;
    }
    t->total_bytes = off;
}

int tile_grid(const qvmcuda_state* s, uint64_t n_tiles) {
    const uint64_t cap = (uint64_t)s->sm_count * 3 * 4;   // 3 resident CTAs per SM, 4 waves of work per CTA slot
    return (int)(n_tiles < cap ? n_tiles : cap);
}

// The tile-kernel instantiations are compiled in their own translation units (qv_tile_inst_*.cu), one per
// (mode, register bits), so that the build runs in parallel; each exports one launcher.
int launch_tile_p(qvmcuda_state* s, const qv::Step& st, const QvPassHeader& h, const uint8_t* d_tables, qv::JitKernel* jk) {
    const bool full = h.T == QV_MAX_TILE_BITS;
    int mode = 0;
    if (h.pull) {
        if (!s->remap_pull || !s->d_alt) return fail("pull pass without an attached alternate buffer");
        if (h.uses_peers) return fail("malformed pass header: pull pass whose tile spans rank bits");
        mode = 2;
    } else if (h.uses_peers) {
        mode = 1;
    }
    std::vector<uint8_t> patched;
    const uint8_t* blob = st.blob.data();
    if (mode == 2 && s->zero_ranks) {      // known-zero shards: the pass zero-fills instead of fetching them
        patched = st.blob;
        QvPassHeader hp = h;
        hp.zero_ranks = s->zero_ranks;
        std::memcpy(patched.data(), &hp, sizeof(hp));
        blob = patched.data();
    }
    QvTileLaunch L;
    L.blob = blob;
    L.blob_bytes = st.blob.size();
    L.full = full;
    L.grid = tile_grid(s, h.n_tiles);
    L.smem = (size_t)sizeof(qvc) << h.T;
    L.stream = s->stream;
    L.peers = &s->peers;
    L.tables = (const qvc*)d_tables;
    L.alt_own = s->d_alt;
    const char* err = nullptr;
    if (jk && full && mode != 1) {
        // compiled pass: the same tile program, executed by its own straight-line kernel
        qv::JitLaunch J;
        J.blob = L.blob;
        J.blob_bytes = L.blob_bytes;
        J.grid = L.grid;
        J.smem = L.smem;
        J.sm_count = s->sm_count;
        J.stream = (void*)s->stream;
        J.peers = L.peers;
        J.tables = L.tables;
        J.alt_own = L.alt_own;
        err = qv::jit_launch(jk, J);
    } else if (h.reg_bits == 4) {
        // the 16-amplitudes-per-thread kernel exists for full local tiles only (the scheduler never asks otherwise)
        if (!full || h.threads_log2 != 7 || mode != 0) return fail("4 register bits need a full 12-bit local tile");
        err = qv_launch_tile_0_4(L);
    } else {
        if (h.reg_bits != 3 || h.threads_log2 != 8) return fail("malformed pass header");
        err = mode == 0 ? qv_launch_tile_0_3(L) : mode == 1 ? qv_launch_tile_1_3(L) : qv_launch_tile_2_3(L);
    }
    if (err) return fail(std::string("tile kernel launch: ") + err);
    g_launches++;
    if (mode == 2) {    // pull pass: results sit in the alternate buffers; every rank flips after the same step
        std::swap(s->d_amps, s->d_alt);
        std::swap(s->peers, s->peers_alt);
    }
    return 0;
}

int launch_tile(qvmcuda_state* s, const qv::Step& st, const uint8_t* d_tables, qv::JitKernel* jk) {
    QvPassHeader h;
    std::memcpy(&h, st.blob.data(), sizeof(h));
    if (st.blob.size() > QV_PROG_LARGE_BYTES) return fail("pass control program too large");
    if ((h.uses_peers || h.pull) && s->world < 2) return fail("peer pass on a state without attached peers");
    return launch_tile_p(s, st, h, d_tables, jk);
}

int launch_big(qvmcuda_state* s, const qv::Step& st, const uint8_t* d_mat) {
    if (st.big.diag) {
        uint64_t blocks = (s->n_amps + QV_THREADS - 1) / QV_THREADS;
        const uint64_t cap = (uint64_t)s->sm_count * 8 * 4;
        if (blocks > cap) blocks = cap;
        qv_bigdiag_kernel<<<(int)blocks, QV_THREADS, 0, s->stream>>>(s->d_amps, st.big, (const qvc*)d_mat, s->n_amps);
        g_launches++;
        CK(cudaGetLastError());
        return 0;
    }
    if (st.big.k > 11) return fail("dense gates on more than 11 mixing qubits are not supported");
    if (big_uses_mma(st) && s->n_bits >= (int)st.big.k) {
        const uint32_t k = st.big.k, D2 = 2u << k, G = QV_MMA_ELEMS >> k, ld = D2 + 4;
        const uint32_t w_in_smem = k <= 4 ? 1u : 0u;
        const size_t smem = ((size_t)2 * G * ld + (w_in_smem ? (size_t)D2 * (D2 + 4) : 0)) * sizeof(double);
        static std::atomic<bool> attr_set[64];
        if (s->device >= 0 && s->device < 64 && !attr_set[s->device].exchange(true))
            CK(cudaFuncSetAttribute(qv_bigmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        const uint64_t n_groups = 1ull << (s->n_bits - (int)k);
        uint64_t blocks = (n_groups + G - 1) / G;
        const uint64_t cap = (uint64_t)s->sm_count * 2 * 4;
        if (blocks > cap) blocks = cap;
        const double* W = reinterpret_cast<const double*>(d_mat + st.bigmat.size() * sizeof(qv::cd));
        qv_bigmma_kernel<<<(int)blocks, QV_THREADS, smem, s->stream>>>(s->d_amps, st.big, W, (uint32_t)s->n_bits, w_in_smem);
        g_launches++;
        CK(cudaGetLastError());
        return 0;
    }
    const uint64_t groups = 1ull << (s->n_bits - (int)st.big.k);
    const uint64_t G = QV_BIG_ELEMS >> st.big.k;
    uint64_t blocks = (groups + G - 1) / G;
    const uint64_t cap = (uint64_t)s->sm_count * 16;
    if (blocks > cap) blocks = cap;
    qv_big_kernel<<<(int)blocks, QV_THREADS, 0, s->stream>>>(s->d_amps, st.big, (const qvc*)d_mat, (uint32_t)s->n_bits);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

// run a compiled tape whose step data already sits in d_buf
// Pull remap: gather this rank's future shard from every rank's current buffer into the alternate
// buffer, then flip (every rank runs the same step, so all pointer tables flip together).
int launch_remap(qvmcuda_state* s, const qv::Step& st) {
    if (!s->remap_pull || !s->d_alt) return fail("pull remap without an attached alternate buffer");
    uint64_t blocks = (s->n_amps + QV_THREADS * 4 - 1) / (QV_THREADS * 4);
    const uint64_t cap = (uint64_t)s->sm_count * 8 * 4;
    if (blocks > cap) blocks = cap;
    qv_remap_pull_kernel<<<(int)blocks, QV_THREADS, 0, s->stream>>>(s->peers, s->d_alt, st.remap);
    g_launches++;
    CK(cudaGetLastError());
    std::swap(s->d_amps, s->d_alt);
    std::swap(s->peers, s->peers_alt);
    return 0;
}

int launch_step(qvmcuda_state* s, const qv::Step& st, const uint8_t* d_data, qv::JitKernel* jk = nullptr) {
    int rc;
    if (st.kind == qv::Step::TILE) rc = launch_tile(s, st, d_data, jk);
    else if (st.kind == qv::Step::BIG) rc = launch_big(s, st, d_data);
    else rc = launch_remap(s, st);
    if (st.uses_peers || st.kind == qv::Step::REMAP) s->zero_ranks = 0;     // after an exchange nobody is known to be zero
    return rc;
}

// kernels of the compiled passes of a tape on the state's device (nullptr entries run through the interpreter kernel)
void prepare_jit(qvmcuda_state* s, const qv::Tape& tape, std::vector<qv::JitKernel*>& out) {
    std::vector<const qv::Step*> ptrs;
    for (const qv::Step& st : tape.steps) ptrs.push_back(&st);
    qv::jit_prepare(ptrs, s->device, out);
}

const std::vector<qv::JitKernel*>& tape_jit(qvmcuda_state* s, qvmcuda_tape* t) {
    std::vector<qv::JitKernel*>& v = t->jit[s->device];
    if (v.size() != t->tape.steps.size() || !t->jit_complete[s->device]) {
        prepare_jit(s, t->tape, v);
        // in asynchronous mode kernels arrive later: ask again on the next run until every eligible step has one
        t->jit_complete[s->device] = qv::jit_policy() != qv::JitPolicy::ASYNC;
    }
    return v;
}

int materialize_locked(qvmcuda_state* s);

int run_steps(qvmcuda_state* s, const qv::Tape& tape, const std::vector<size_t>& offsets, const uint8_t* d_buf,
              const std::vector<qv::JitKernel*>* jit = nullptr, qvmcuda_tape* owner = nullptr) {
    std::vector<qv::JitKernel*> local;
    if (!jit) {
        prepare_jit(s, tape, local);
        jit = &local;
    }
    size_t first = 0;
    if (s->lazy_basis && !tape.steps.empty()) {
        // first pass after a lazy reset: its compiled no-load variant, with the basis index patched into the header copy
        const qv::Step& st0 = tape.steps[0];
        qv::JitKernel* jk = nullptr;
        if (s->world == 1 && st0.kind == qv::Step::TILE && !st0.uses_peers && (*jit)[0]) {
            auto cached = owner ? owner->jit_basis.find(s->device) : std::map<int, qv::JitKernel*>::iterator();
            if (owner && cached != owner->jit_basis.end()) {
                jk = cached->second;
            } else {
                std::vector<const qv::Step*> one = {&st0};
                std::vector<qv::JitKernel*> got;
                qv::jit_prepare(one, s->device, got, qv::kVariantSrcBasis);
                jk = got[0];
                if (owner && qv::jit_policy() != qv::JitPolicy::ASYNC) owner->jit_basis[s->device] = jk;
            }
        }
        if (jk) {
            qv::Step patched;               // the control program only: the tables already sit in d_buf
            patched.kind = qv::Step::TILE;
            patched.blob = st0.blob;
            QvPassHeader h;
            std::memcpy(&h, patched.blob.data(), sizeof(h));
            h.src_basis = 1;
            h.basis_index = s->lazy_index;
            std::memcpy(patched.blob.data(), &h, sizeof(h));
            s->lazy_basis = false;
            if (int rc = launch_step(s, patched, d_buf + offsets[0], jk)) return rc;
            first = 1;
        } else if (int rc = materialize_locked(s)) {
            return rc;
        }
    }
    for (size_t i = first; i < tape.steps.size(); i++) {
        int rc = launch_step(s, tape.steps[i], d_buf + offsets[i], (*jit)[i]);
        if (rc) return rc;
    }
    return 0;
}

int gates_from_flat(int n_gates, const int32_t* ks, const int32_t* qubits, const double* matrices, std::vector<qv::Gate>& out) {
    if (n_gates < 0) return fail("negative gate count");
    out.resize((size_t)n_gates);
    size_t qo = 0, mo = 0;
    for (int g = 0; g < n_gates; g++) {
        const int k = ks[g];
        if (k < 1 || k > 16) return fail("gate arity out of range (1..16)");
        out[g].qubits.assign(qubits + qo, qubits + qo + k);
        qo += (size_t)k;
        const size_t d = (size_t)1 << k;
        out[g].mat.resize(d * d);
        for (size_t i = 0; i < d * d; i++) out[g].mat[i] = qv::cd(matrices[mo + 2 * i], matrices[mo + 2 * i + 1]);
        mo += 2 * d * d;
    }
    return 0;
}

qv::CompileOptions make_options(const qvmcuda_state* s, uint32_t flags) {
    qv::CompileOptions opt;
    opt.fuse = (flags & QVMCUDA_FUSE) != 0;
    opt.absorb_swaps = (flags & QVMCUDA_ABSORB_SWAPS) != 0;
    static const int forced_reg_bits = getenv("QVMCUDA_REG_BITS") ? atoi(getenv("QVMCUDA_REG_BITS")) : 0;   // profiling knob: 3 or 4
    if (forced_reg_bits == 3 || forced_reg_bits == 4) opt.reg_bits = forced_reg_bits;
    static const int euler = getenv("QVMCUDA_EULER") ? atoi(getenv("QVMCUDA_EULER")) : 0;     // 1 always, 0 never, -1 cost model
    opt.euler_split = euler;
    static const int fuse_mats = getenv("QVMCUDA_FUSE_MATRICES") ? atoi(getenv("QVMCUDA_FUSE_MATRICES")) : 1;
    opt.fuse_matrices = fuse_mats != 0;
    static const int route = getenv("QVMCUDA_ROUTE_SWAPS") ? atoi(getenv("QVMCUDA_ROUTE_SWAPS")) : -1;   // 1 always, 0 never, -1 cost model
    opt.route_swaps = route;
    static const int hoist = getenv("QVMCUDA_HOIST_REMAPS") ? atoi(getenv("QVMCUDA_HOIST_REMAPS")) : -1;  // 1 always, 0 never, -1 cost model
    opt.hoist_remaps = hoist;
    if (s) {
        opt.rank = s->rank;
        opt.n_local_bits = s->n_bits;
        opt.remap_pull = s->remap_pull;
    }
    return opt;
}

int materialize_locked(qvmcuda_state* s) {
    if (!s->lazy_basis) return 0;
    s->lazy_basis = false;
    CK(cudaMemsetAsync(s->d_amps, 0, s->n_amps * sizeof(qvc), s->stream));
    if (s->lazy_index == kLazyAllZero) return 0;
    qv_set_one_kernel<<<1, 1, 0, s->stream>>>(s->d_amps, s->lazy_index);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

// Program data (diagonal tables / big matrices) of a tape go through a persistent pinned staging buffer and a persistent
// device scratch buffer of the STATE: no allocation on the gate path.  The staging buffer is reused only after the previous
// upload has completed (event), the device buffer is protected by stream order.  scratch_tape remembers whose data it holds.
int upload_to_scratch_locked(qvmcuda_state* s, const qvmcuda_tape& t) {
    const size_t need = t.total_bytes ? t.total_bytes : 256;
    if (need > s->scratch_cap) {
        CK(cudaStreamSynchronize(s->stream));
        if (s->d_scratch) cudaFree(s->d_scratch);
        if (s->h_scratch) cudaFreeHost(s->h_scratch);
        s->d_scratch = nullptr;
        s->h_scratch = nullptr;
        s->scratch_cap = 0;
        const size_t cap = std::max<size_t>(need * 2, 1 << 20);
        CK(cudaMalloc((void**)&s->d_scratch, cap));
        CK(cudaMallocHost((void**)&s->h_scratch, cap));
        if (!s->upload_done) CK(cudaEventCreateWithFlags(&s->upload_done, cudaEventDisableTiming));
        s->scratch_cap = cap;
    } else {
        CK(cudaEventSynchronize(s->upload_done));
    }
    for (size_t i = 0; i < t.tape.steps.size(); i++) fill_step_data(t.tape.steps[i], s->h_scratch + t.offsets[i]);
    CK(cudaMemcpyAsync(s->d_scratch, s->h_scratch, need, cudaMemcpyHostToDevice, s->stream));
    CK(cudaEventRecord(s->upload_done, s->stream));
    s->scratch_tape = t.id;
    return 0;
}

void drop_tape_cache_locked(qvmcuda_state* s) {
    for (size_t i = s->tape_cache.size(); i-- > 0;)
        if (s->tape_cache[i].tape->checked_out.load() == 0) {
            delete s->tape_cache[i].tape;
            s->tape_cache.erase(s->tape_cache.begin() + i);
        }
}

// Schedule of `gates` from the state's current layout: from the state's cache when the exact gate list, flags and starting
// layout were seen before, else compiled now (and cached).  The tape stays owned by the state.
int cached_tape_locked(qvmcuda_state* s, const std::vector<qv::Gate>& gates, uint32_t flags, qvmcuda_tape** out) {
    static const bool trace = getenv("QVMCUDA_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    static const size_t cache_slots = getenv("QVMCUDA_TAPE_CACHE") ? (size_t)atoi(getenv("QVMCUDA_TAPE_CACHE")) : 4;
    uint64_t key = 1469598103934665603ull;
    auto mix = [&key](const void* p, size_t n) {
        const uint8_t* raw = static_cast<const uint8_t*>(p);
        for (size_t i = 0; i < n / 8; i++) {
            uint64_t w;
            std::memcpy(&w, raw + 8 * i, 8);      // the qubit lists are only 4-byte aligned
            key = (key ^ w) * 1099511628211ull;
        }
        const uint8_t* b = raw + (n & ~(size_t)7);
        for (size_t i = 0; i < (n & 7); i++) key = (key ^ b[i]) * 1099511628211ull;
    };
    for (const qv::Gate& g : gates) {
        mix(g.qubits.data(), g.qubits.size() * sizeof(int));
        mix(g.mat.data(), g.mat.size() * sizeof(qv::cd));
    }
    mix(&flags, sizeof(flags));
    mix(s->l2p.data(), s->l2p.size() * sizeof(int));
    for (size_t i = 0; i < s->tape_cache.size(); i++) {
        QvTapeCacheEntry& e = s->tape_cache[i];
        if (e.key != key || e.flags != flags || e.l2p_in != s->l2p || e.gates.size() != gates.size()) continue;
        bool same = true;
        for (size_t g = 0; g < gates.size() && same; g++)
            same = e.gates[g].qubits == gates[g].qubits && e.gates[g].mat == gates[g].mat;
        if (!same) continue;
        *out = e.tape;
        if (i) std::rotate(s->tape_cache.begin(), s->tape_cache.begin() + i, s->tape_cache.begin() + i + 1);
        return 0;
    }
    std::unique_ptr<qvmcuda_tape> fresh(new qvmcuda_tape());
    try {
        const int total_bits = s->n_bits + log2_exact((uint64_t)s->world);
        fresh->tape = qv::compile(gates, total_bits, make_options(s, flags), s->l2p);
    } catch (const std::exception& e) {
        return fail(std::string("schedule: ") + e.what());
    }
    fresh->flags = flags;
    fresh->n_local = s->n_bits;
    fresh->rank = s->rank;
    fresh->world = s->world;
    fresh->ephemeral = true;
    fresh->state_owned = true;
    layout_tape(fresh.get());
    if (trace) fprintf(stderr, "[qvmcuda] schedule: %zu steps in %.3f ms\n", fresh->tape.steps.size(),
                       std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    QvTapeCacheEntry e;
    e.key = key;
    e.flags = flags;
    e.l2p_in = s->l2p;
    e.gates = gates;
    e.tape = fresh.release();
    *out = e.tape;
    s->tape_cache.insert(s->tape_cache.begin(), std::move(e));
    // evict the least recently used schedules nobody holds a handle to (the newest one always stays)
    for (size_t i = s->tape_cache.size(); i-- > 1 && s->tape_cache.size() > std::max<size_t>(cache_slots, 1);)
        if (s->tape_cache[i].tape->checked_out.load() == 0) {
            delete s->tape_cache[i].tape;
            s->tape_cache.erase(s->tape_cache.begin() + i);
        }
    return 0;
}

// compile + upload + run in immediate mode (state mutex held)
int run_gates_locked(qvmcuda_state* s, const std::vector<qv::Gate>& gates, uint32_t flags) {
    static const bool trace = getenv("QVMCUDA_TRACE") != nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    struct Tracer {
        bool on; std::chrono::steady_clock::time_point t0; size_t n;
        ~Tracer() {
            if (on) fprintf(stderr, "[qvmcuda] apply_gates: %zu gates, host time %.3f ms\n", n,
                            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
    } tracer{trace, t_begin, gates.size()};
    qvmcuda_tape* tp = nullptr;
    if (int rc = cached_tape_locked(s, gates, flags, &tp)) return rc;
    qvmcuda_tape& t = *tp;
    if (t.tape.steps.empty()) {
        s->l2p = t.tape.l2p;
        return 0;
    }
    if (s->scratch_tape != t.id)
        if (int rc = upload_to_scratch_locked(s, t)) return rc;
    int rc = run_steps(s, t.tape, t.offsets, s->d_scratch, &tape_jit(s, &t), &t);
    if (rc) return rc;
    s->l2p = t.tape.l2p;
    return 0;
}

// Undo absorbed swaps so that physical bit q holds logical qubit q again.
int canonicalize_locked(qvmcuda_state* s) {
    if (int rc = materialize_locked(s)) return rc;      // every caller is about to touch the amplitudes
    if (l2p_is_identity(s->l2p)) return 0;
    if (s->world > 1) return 0;   // shards keep their layout; the host maps indices through qvmcuda_state_layout
    std::vector<int> l2p = s->l2p;
    std::vector<qv::Gate> swaps;
    static const double sw[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
    for (int q = 0; q < (int)l2p.size(); q++) {
        while (l2p[q] != q) {
            const int p = l2p[q];
            int r = -1;
            for (int x = 0; x < (int)l2p.size(); x++)
                if (l2p[x] == q) r = x;
            qv::Gate g;
            g.qubits = {p, q};   // physical bits
            g.mat.resize(16);
            for (int i = 0; i < 16; i++) g.mat[i] = qv::cd(sw[i], 0.0);
            swaps.push_back(g);
            l2p[r] = p;
            l2p[q] = q;
        }
    }
    s->l2p.clear();   // the swap gates address physical bits: identity map while they run
    int rc = run_gates_locked(s, swaps, QVMCUDA_FUSE);
    if (rc) return rc;
    s->l2p = l2p;
    return 0;
}

int reduce_locked(qvmcuda_state* s, uint64_t count, int mode, uint32_t q, uint64_t dim, double* out) {
    if (int rc = materialize_locked(s)) return rc;
    uint64_t blocks = (count + QV_THREADS - 1) / QV_THREADS;
    if (blocks > kReduceBlocks) blocks = kReduceBlocks;
    if (blocks == 0) blocks = 1;
    qv_reduce_kernel<<<(int)blocks, QV_THREADS, 0, s->stream>>>(s->d_amps, count, mode, q, dim, s->d_partial);
    qv_final_sum_kernel<<<1, QV_THREADS, 0, s->stream>>>(s->d_partial, (uint32_t)blocks, s->d_partial + kReduceBlocks);
    g_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, s->d_partial + kReduceBlocks, sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

// Probability that physical bit `pbit` equals `value`, restricted to this shard (the host adds the
// shards' partial results).  A rank bit is constant over the shard.
int prob_bit_locked(qvmcuda_state* s, int pbit, int value, double* p) {
    if (pbit < s->n_bits) return reduce_locked(s, s->n_amps / 2, value ? 1 : 3, (uint32_t)pbit, 0, p);
    const int mine = (s->rank >> (pbit - s->n_bits)) & 1;
    if (mine != value) {
        *p = 0.0;
        return 0;
    }
    return reduce_locked(s, s->n_amps, 0, 0, 0, p);
}

int aux_locked(qvmcuda_state* s, uint64_t n_doubles, double** out) {
    if (n_doubles > s->aux_cap) {
        CK(cudaStreamSynchronize(s->stream));
        if (s->d_aux) cudaFree(s->d_aux);
        s->d_aux = nullptr;
        s->aux_cap = 0;
        CK(cudaMalloc((void**)&s->d_aux, n_doubles * sizeof(double)));
        s->aux_cap = n_doubles;
    }
    *out = s->d_aux;
    return 0;
}

int elementwise_locked(qvmcuda_state* s, int mode, uint32_t q, uint32_t q2, uint32_t keep, double f) {
    if (int rc = materialize_locked(s)) return rc;
    uint64_t blocks = (s->n_amps + QV_THREADS - 1) / QV_THREADS;
    const uint64_t cap = (uint64_t)s->sm_count * 8 * 4;
    if (blocks > cap) blocks = cap;
    qv_elementwise_kernel<<<(int)blocks, QV_THREADS, 0, s->stream>>>(s->d_amps, s->n_amps, mode, q, q2, keep, f);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

const char* qvmcuda_last_error(void) { return g_err.c_str(); }

int qvmcuda_device_count(int* count) {
    if (!count) return fail("null argument");
    CK(cudaGetDeviceCount(count));
    return 0;
}

int qvmcuda_launch_count(uint64_t* count) {
    if (!count) return fail("null argument");
    *count = g_launches.load();
    return 0;
}

int qvmcuda_state_create(uint64_t n_amplitudes, int device, qvmcuda_state** out) {
    if (!out) return fail("null argument");
    *out = nullptr;
    const int nb = log2_exact(n_amplitudes);
    if (nb < 1) return fail("state length must be a power of two >= 2");
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (count < 1) return fail("no CUDA device: libqvmcuda has no CPU fallback");
    if (device < 0 || device >= count) return fail("device ordinal out of range");
    DeviceGuard dg(device);
    if (!dg.ok) return fail("cannot select device");
    qvmcuda_state* s = new qvmcuda_state();
    s->device = device;
    s->n_amps = n_amplitudes;
    s->n_bits = nb;
    cudaError_t e = cudaMalloc((void**)&s->d_amps, n_amplitudes * sizeof(qvc));
    if (e != cudaSuccess) {
        delete s;
        return fail(std::string("cudaMalloc of the amplitude vector: ") + cudaGetErrorString(e));
    }
    e = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_partial, (kReduceBlocks + 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMemsetAsync(s->d_amps, 0, n_amplitudes * sizeof(qvc), s->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
    if (e != cudaSuccess) {
        cudaFree(s->d_amps);
        cudaFree(s->d_partial);
        if (s->stream) cudaStreamDestroy(s->stream);
        delete s;
        return fail(std::string("state setup: ") + cudaGetErrorString(e));
    }
    s->own_stream = true;
    cudaDeviceGetAttribute(&s->sm_count, cudaDevAttrMultiProcessorCount, device);
    s->l2p.resize(nb);
    for (int i = 0; i < nb; i++) s->l2p[i] = i;
    s->peers.base[0] = s->d_amps;
    *out = s;
    return 0;
}

int qvmcuda_state_destroy(qvmcuda_state* s) {
    if (!s) return 0;
    {
        std::lock_guard<std::mutex> lk(s->mu);
        DeviceGuard dg(s->device);
        cudaStreamSynchronize(s->stream);
        for (void* p : s->opened) cudaIpcCloseMemHandle(p);
        if (s->d_alt) cudaFree(s->d_alt);
        cudaFree(s->d_amps);
        cudaFree(s->d_partial);
        if (s->d_scratch) cudaFree(s->d_scratch);
        if (s->h_scratch) cudaFreeHost(s->h_scratch);
        if (s->d_sample_tree) cudaFree(s->d_sample_tree);
        if (s->d_shots) cudaFree(s->d_shots);
        if (s->d_aux) cudaFree(s->d_aux);
        if (s->upload_done) cudaEventDestroy(s->upload_done);
        if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
        s->d_amps = nullptr;
        for (QvTapeCacheEntry& e : s->tape_cache) delete e.tape;      // ephemeral tapes own no device memory
        s->tape_cache.clear();
    }
    delete s;
    return 0;
}

int qvmcuda_state_length(qvmcuda_state* s, uint64_t* n) {
    if (!s || !n) return fail("null argument");
    *n = s->n_amps;
    return 0;
}

int qvmcuda_state_set_stream(qvmcuda_state* s, uint64_t cuda_stream) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    CK(cudaStreamSynchronize(s->stream));
    if (s->own_stream && s->stream) cudaStreamDestroy(s->stream);
    s->stream = (cudaStream_t)(uintptr_t)cuda_stream;
    s->own_stream = false;
    return 0;
}

int qvmcuda_synchronize(qvmcuda_state* s) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int qvmcuda_download(qvmcuda_state* s, double* dst, uint64_t offset, uint64_t count) {
    if (!s || !dst) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (offset > s->n_amps || count > s->n_amps - offset) return fail("download range out of bounds");
    DeviceGuard dg(s->device);
    if (int rc = canonicalize_locked(s)) return rc;
    CK(cudaMemcpyAsync(dst, s->d_amps + offset, count * sizeof(qvc), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int qvmcuda_upload(qvmcuda_state* s, const double* src, uint64_t offset, uint64_t count) {
    if (!s || !src) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (offset > s->n_amps || count > s->n_amps - offset) return fail("upload range out of bounds");
    DeviceGuard dg(s->device);
    if (offset == 0 && count == s->n_amps) s->lazy_basis = false;      // fully overwritten: nothing to write first
    if (int rc = canonicalize_locked(s)) return rc;
    CK(cudaMemcpyAsync(s->d_amps + offset, src, count * sizeof(qvc), cudaMemcpyHostToDevice, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int qvmcuda_set_basis_state(qvmcuda_state* s, uint64_t basis) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (basis >= s->n_amps) return fail("basis state out of range");
    DeviceGuard dg(s->device);
    for (size_t i = 0; i < s->l2p.size(); i++) s->l2p[i] = (int)i;   // content is replaced: layout resets
    static const bool lazy = !(getenv("QVMCUDA_LAZY_RESET") && atoi(getenv("QVMCUDA_LAZY_RESET")) == 0);
    s->lazy_basis = true;
    s->lazy_index = basis;
    // shards are read by their peers, and small states gain nothing: write those now
    if (!lazy || s->world > 1 || s->n_bits < QV_MAX_TILE_BITS) return materialize_locked(s);
    return 0;
}

int qvmcuda_set_zero_state(qvmcuda_state* s) { return qvmcuda_set_basis_state(s, 0); }

int qvmcuda_copy(qvmcuda_state* dst, qvmcuda_state* src) {
    if (!dst || !src) return fail("null argument");
    if (dst == src) return 0;
    std::lock(dst->mu, src->mu);
    std::lock_guard<std::mutex> l1(dst->mu, std::adopt_lock), l2(src->mu, std::adopt_lock);
    DeviceGuard dg(src->device);
    if (int rc = canonicalize_locked(src)) return rc;
    CK(cudaStreamSynchronize(src->stream));
    const uint64_t n = dst->n_amps < src->n_amps ? dst->n_amps : src->n_amps;
    if (n == dst->n_amps) {
        for (size_t i = 0; i < dst->l2p.size(); i++) dst->l2p[i] = (int)i;
        dst->lazy_basis = false;       // fully overwritten
    } else {
        DeviceGuard dgd(dst->device);
        if (int rc = materialize_locked(dst)) return rc;
    }
    CK(cudaMemcpyAsync(dst->d_amps, src->d_amps, n * sizeof(qvc), cudaMemcpyDefault, dst->stream));
    CK(cudaStreamSynchronize(dst->stream));
    return 0;
}

int qvmcuda_apply_gates(qvmcuda_state* s, int n_gates, const int32_t* ks, const int32_t* qubits,
                        const double* matrices, uint32_t flags) {
    if (!s || (n_gates > 0 && (!ks || !qubits || !matrices))) return fail("null argument");
    std::vector<qv::Gate> gates;
    if (int rc = gates_from_flat(n_gates, ks, qubits, matrices, gates)) return rc;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    return run_gates_locked(s, gates, flags);
}

int qvmcuda_apply_matrix(qvmcuda_state* s, int k, const int32_t* qubits, const double* matrix) {
    const int32_t ks[1] = {k};
    return qvmcuda_apply_gates(s, 1, ks, qubits, matrix, 0);
}

int qvmcuda_tape_compile(int n_qubits, int n_gates, const int32_t* ks, const int32_t* qubits,
                         const double* matrices, uint32_t flags, qvmcuda_tape** out) {
    if (!out) return fail("null argument");
    *out = nullptr;
    std::vector<qv::Gate> gates;
    if (int rc = gates_from_flat(n_gates, ks, qubits, matrices, gates)) return rc;
    qvmcuda_tape* t = new qvmcuda_tape();
    t->flags = flags;
    t->n_local = n_qubits;
    try {
        t->tape = qv::compile(gates, n_qubits, make_options(nullptr, flags));
    } catch (const std::exception& e) {
        delete t;
        return fail(std::string("schedule: ") + e.what());
    }
    layout_tape(t);
    *out = t;
    return 0;
}

static int tape_device_buffer(qvmcuda_state* s, qvmcuda_tape* t, uint8_t** out) {
    if (t->ephemeral) {      // no allocation, no blocking copy: one asynchronous upload per tape
        if (s->scratch_tape != t->id)
            if (int rc = upload_to_scratch_locked(s, *t)) return rc;
        *out = s->d_scratch;
        return 0;
    }
    uint8_t*& d_buf = t->d_blobs[s->device];
    if (!d_buf) {
        std::vector<uint8_t> host(t->total_bytes ? t->total_bytes : 256, 0);
        for (size_t i = 0; i < t->tape.steps.size(); i++) fill_step_data(t->tape.steps[i], host.data() + t->offsets[i]);
        CK(cudaMalloc((void**)&d_buf, host.size()));
        CK(cudaMemcpy(d_buf, host.data(), host.size(), cudaMemcpyHostToDevice));
    }
    *out = d_buf;
    return 0;
}

int qvmcuda_shard_compile(qvmcuda_state* s, int n_gates, const int32_t* ks, const int32_t* qubits,
                          const double* matrices, uint32_t flags, qvmcuda_tape** out) {
    if (!s || !out) return fail("null argument");
    *out = nullptr;
    std::vector<qv::Gate> gates;
    if (int rc = gates_from_flat(n_gates, ks, qubits, matrices, gates)) return rc;
    std::lock_guard<std::mutex> lk(s->mu);
    // served from the state's schedule cache: a sharded state keeps the layout a circuit leaves behind, so a circuit run in a
    // loop cycles through a few (layout, schedule) pairs; the handle is returned with qvmcuda_tape_destroy as usual
    qvmcuda_tape* t = nullptr;
    if (int rc = cached_tape_locked(s, gates, flags, &t)) return rc;
    if (s->lazy_basis) {
        // A shard that holds only zeros stays unwritten if nothing will read it: local passes map zeros to zeros (skipped),
        // and a fused pull pass does not fetch ranks flagged in zero_ranks.  Anything else (in-place peer pass, stand-alone
        // remap kernel, dense k >= 3 gate) reads memory: write the zeros NOW, before the host's barrier in front of the first
        // exchange step, because peers read this buffer.
        bool stay = s->lazy_index == kLazyAllZero && (s->zero_ranks >> s->rank & 1u);
        for (const qv::Step& st : t->tape.steps) {
            if (!stay) break;
            if (st.kind != qv::Step::TILE) stay = false;
            else if (st.uses_peers) {
                QvPassHeader h;
                std::memcpy(&h, st.blob.data(), sizeof(h));
                stay = h.pull != 0;
                break;
            }
        }
        if (!stay) {
            DeviceGuard dg(s->device);
            if (int rc = materialize_locked(s)) return rc;
        }
    }
    t->checked_out++;
    *out = t;
    return 0;
}

int qvmcuda_shard_plan(int n_total, int world, int rank, int remap_pull, int32_t* l2p, int n_gates, const int32_t* ks,
                       const int32_t* qubits, const double* matrices, uint32_t flags, qvmcuda_tape** out) {
    if (!out || !l2p) return fail("null argument");
    *out = nullptr;
    if (world < 1 || (world & (world - 1)) || rank < 0 || rank >= world) return fail("world must be a power of two, 0 <= rank < world");
    const int g = log2_exact((uint64_t)world);
    if (n_total <= g) return fail("fewer qubits than rank bits");
    std::vector<qv::Gate> gates;
    if (int rc = gates_from_flat(n_gates, ks, qubits, matrices, gates)) return rc;
    qvmcuda_tape* t = new qvmcuda_tape();
    t->flags = flags;
    t->n_local = n_total - g;
    t->rank = rank;
    t->world = world;
    t->ephemeral = true;
    try {
        qv::CompileOptions opt = make_options(nullptr, flags);
        opt.rank = rank;
        opt.n_local_bits = n_total - g;
        opt.remap_pull = remap_pull != 0;
        t->tape = qv::compile(gates, n_total, opt, std::vector<int>(l2p, l2p + n_total));
    } catch (const std::exception& e) {
        delete t;
        return fail(std::string("schedule: ") + e.what());
    }
    layout_tape(t);
    for (int q = 0; q < n_total; q++) l2p[q] = t->tape.l2p[q];
    *out = t;
    return 0;
}

int qvmcuda_tape_num_steps(qvmcuda_tape* t, int* n_steps) {
    if (!t || !n_steps) return fail("null argument");
    *n_steps = (int)t->tape.steps.size();
    return 0;
}

int qvmcuda_tape_step_flags(qvmcuda_tape* t, int step, uint32_t* flags) {
    if (!t || !flags) return fail("null argument");
    if (step < 0 || step >= (int)t->tape.steps.size()) return fail("step out of range");
    const qv::Step& st = t->tape.steps[step];
    *flags = (st.uses_peers ? QVMCUDA_STEP_PEER : 0u) | (st.is_remap ? QVMCUDA_STEP_REMAP : 0u);
    return 0;
}

int qvmcuda_tape_step_info(qvmcuda_tape* t, int step, int64_t info[8]) {
    if (!t || !info) return fail("null argument");
    if (step < 0 || step >= (int)t->tape.steps.size()) return fail("step out of range");
    const qv::Step& st = t->tape.steps[step];
    std::memset(info, 0, 8 * sizeof(int64_t));
    info[0] = (st.uses_peers ? QVMCUDA_STEP_PEER : 0u) | (st.is_remap ? QVMCUDA_STEP_REMAP : 0u);
    info[1] = (int64_t)st.kind;
    info[2] = st.n_gates;
    if (st.kind == qv::Step::REMAP) info[3] = st.remap.n_pairs;
    if (st.kind == qv::Step::TILE) {
        QvPassHeader h;
        std::memcpy(&h, st.blob.data(), sizeof(h));
        info[4] = h.n_uops - h.n_rounds;
        info[5] = h.n_rounds;
        if (h.pull) info[3] = h.pull_remap.n_pairs;
        else if (h.uses_peers) {      // in-place exchange: rank bits inside the tile
            int sbits = 0;
            for (uint32_t k = 0; k < h.n_tile_segs; k++)
                for (uint32_t b = 0; b < h.tile_segs[k].len; b++)
                    if ((uint32_t)h.tile_segs[k].dst + b >= h.n_local_bits) sbits++;
            info[3] = sbits;
            info[6] = 1;
        }
    }
    return 0;
}

int qvmcuda_tape_run_step(qvmcuda_state* s, qvmcuda_tape* t, int step) {
    if (!s || !t) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    std::lock_guard<std::mutex> lt(t->mu);
    if (step < 0 || step >= (int)t->tape.steps.size()) return fail("step out of range");
    if (t->n_local != s->n_bits || t->rank != s->rank || t->world != s->world)
        return fail("tape was compiled for a different shard geometry");
    DeviceGuard dg(s->device);
    const qv::Step& st = t->tape.steps[step];
    if (s->lazy_basis && s->lazy_index == kLazyAllZero && st.kind == qv::Step::TILE) {
        QvPassHeader h;
        std::memcpy(&h, st.blob.data(), sizeof(h));
        if (!st.uses_peers) return 0;                  // zeros in, zeros out: nothing to run, nothing to write
        if (h.pull && s->remap_pull && (s->zero_ranks >> s->rank & 1u)) {
            s->lazy_basis = false;                     // nobody fetches this buffer; the pass writes the alternate one in full
        } else if (int rc = materialize_locked(s)) {
            return rc;
        }
    } else if (int rc = materialize_locked(s)) {
        return rc;
    }
    uint8_t* d_buf = nullptr;
    if (int rc = tape_device_buffer(s, t, &d_buf)) return rc;
    return launch_step(s, st, d_buf + t->offsets[step], tape_jit(s, t)[step]);
}

int qvmcuda_tape_commit(qvmcuda_state* s, qvmcuda_tape* t) {
    if (!s || !t) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (t->tape.l2p.size() != s->l2p.size()) return fail("tape layout does not match the state");
    s->l2p = t->tape.l2p;
    return 0;
}

int qvmcuda_state_layout(qvmcuda_state* s, int32_t* l2p, int n) {
    if (!s || !l2p) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (n != (int)s->l2p.size()) return fail("layout length mismatch");
    for (int i = 0; i < n; i++) l2p[i] = s->l2p[i];
    return 0;
}

int qvmcuda_tape_run(qvmcuda_state* s, qvmcuda_tape* t) {
    if (!s || !t) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    std::lock_guard<std::mutex> lt(t->mu);
    if (t->tape.n_bits != s->n_bits) return fail("tape was compiled for a different number of qubits");
    if (s->world != 1) return fail("precompiled tapes are single-device; use qvmcuda_apply_gates on shards");
    DeviceGuard dg(s->device);
    // a tape is compiled against the identity layout (a lazily reset state is in it: keep it lazy for the first pass)
    if (!s->lazy_basis)
        if (int rc = canonicalize_locked(s)) return rc;
    if (t->tape.steps.empty()) {
        s->l2p = t->tape.l2p;
        return 0;
    }
    uint8_t* d_buf = nullptr;
    if (int rc = tape_device_buffer(s, t, &d_buf)) return rc;
    if (int rc = run_steps(s, t->tape, t->offsets, d_buf, &tape_jit(s, t), t)) return rc;
    s->l2p = t->tape.l2p;
    return 0;
}

int qvmcuda_jit_stats(int64_t out[8]) {
    if (!out) return fail("null argument");
    const qv::JitStats st = qv::jit_stats();
    std::memset(out, 0, 8 * sizeof(int64_t));
    out[0] = (int64_t)st.compiled;
    out[1] = (int64_t)st.cache_hits;
    out[2] = (int64_t)st.disk_hits;
    out[3] = (int64_t)st.failed;
    out[4] = (int64_t)st.launches;
    out[5] = (int64_t)(st.compile_seconds * 1e3);
    out[6] = (int64_t)qv::jit_policy();
    out[7] = (int64_t)qv::jit_min_uops();
    return 0;
}

int qvmcuda_tape_jit_source(qvmcuda_tape* t, int step, char* buf, uint64_t buflen, uint64_t* sig) {
    if (!t || !buf || buflen == 0) return fail("null argument");
    if (step < 0 || step >= (int)t->tape.steps.size()) return fail("step out of range");
    const qv::JitSource src = qv::jit_generate(t->tape.steps[step]);
    if (!src.ok) return fail("step has no compiled form: " + src.why_not);
    if (src.text.size() + 1 > buflen) return fail("buffer too small for the generated source");
    std::memcpy(buf, src.text.c_str(), src.text.size() + 1);
    if (sig) *sig = src.sig;
    return 0;
}

int qvmcuda_tape_jit_precompile(qvmcuda_tape* t, int* n_eligible, int* n_ok, char* log, uint64_t loglen) {
    if (!t) return fail("null argument");
    std::vector<const qv::Step*> ptrs;
    for (const qv::Step& st : t->tape.steps) ptrs.push_back(&st);
    std::string text;
    int ne = 0, nk = 0;
    qv::jit_precompile(ptrs, ne, nk, text, t->world == 1);
    if (n_eligible) *n_eligible = ne;
    if (n_ok) *n_ok = nk;
    if (log && loglen) {
        std::strncpy(log, text.c_str(), loglen - 1);
        log[loglen - 1] = 0;
    }
    return 0;
}

int qvmcuda_tape_info(qvmcuda_tape* t, int64_t info[8]) {
    if (!t || !info) return fail("null argument");
    std::memset(info, 0, 8 * sizeof(int64_t));
    info[0] = (int64_t)t->tape.steps.size();
    info[1] = t->tape.n_gates;
    info[2] = t->tape.n_atoms;
    for (const qv::Step& st : t->tape.steps) info[st.kind == qv::Step::TILE ? 3 : 4]++;
    info[5] = (int64_t)t->total_bytes;
    return 0;
}

int qvmcuda_tape_describe(qvmcuda_tape* t, char* buf, uint64_t buflen) {
    if (!t || !buf || buflen == 0) return fail("null argument");
    const std::string s = qv::describe(t->tape);
    std::strncpy(buf, s.c_str(), buflen - 1);
    buf[buflen - 1] = 0;
    return 0;
}

int qvmcuda_tape_destroy(qvmcuda_tape* t) {
    if (!t) return 0;
    if (t->state_owned) {      // a handle from qvmcuda_shard_compile: the tape stays in its state's schedule cache (destroy the
                               // handle before the state it came from)
        int c = t->checked_out.load();
        while (c > 0 && !t->checked_out.compare_exchange_weak(c, c - 1)) {}
        return 0;
    }
    for (auto& kv : t->d_blobs) {
        DeviceGuard dg(kv.first);
        cudaFree(kv.second);
    }
    delete t;
    return 0;
}

int qvmcuda_prob_excited(qvmcuda_state* s, int qubit, double* p) {
    if (!s || !p) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (qubit < 0 || qubit >= (int)s->l2p.size()) return fail("qubit out of range");
    DeviceGuard dg(s->device);
    return prob_bit_locked(s, s->l2p[qubit], 1, p);
}

int qvmcuda_prob_ground(qvmcuda_state* s, int qubit, double* p) {
    if (!s || !p) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (qubit < 0 || qubit >= (int)s->l2p.size()) return fail("qubit out of range");
    DeviceGuard dg(s->device);
    return prob_bit_locked(s, s->l2p[qubit], 0, p);
}

int qvmcuda_norm2(qvmcuda_state* s, double* out) {
    if (!s || !out) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    return reduce_locked(s, s->n_amps, 0, 0, 0, out);
}

int qvmcuda_inner_product(qvmcuda_state* a, qvmcuda_state* b, double out[2]) {
    if (!a || !b || !out) return fail("null argument");
    if (a->n_amps != b->n_amps) return fail("inner product of states of different length");
    if (a->device != b->device) return fail("inner product of states on different devices");
    if (a == b) {
        std::lock_guard<std::mutex> lk(a->mu);
        DeviceGuard dg(a->device);
        out[1] = 0.0;
        return reduce_locked(a, a->n_amps, 0, 0, 0, out);
    }
    std::lock(a->mu, b->mu);
    std::lock_guard<std::mutex> l1(a->mu, std::adopt_lock), l2(b->mu, std::adopt_lock);
    DeviceGuard dg(a->device);
    if (a->world > 1 || b->world > 1) {
        if (a->l2p != b->l2p) return fail("inner product of shards with different qubit layouts");
    } else {
        if (int rc = canonicalize_locked(a)) return rc;
        if (int rc = canonicalize_locked(b)) return rc;
    }
    CK(cudaStreamSynchronize(b->stream));      // b's pending work runs on its own stream
    uint64_t blocks = (a->n_amps + QV_THREADS - 1) / QV_THREADS;
    if (blocks > (kReduceBlocks - 2) / 2) blocks = (kReduceBlocks - 2) / 2;      // d_partial holds kReduceBlocks + 1 doubles: partials + the two results
    if (blocks == 0) blocks = 1;
    qv_inner_kernel<<<(int)blocks, QV_THREADS, 0, a->stream>>>(a->d_amps, b->d_amps, a->n_amps, a->d_partial);
    qv_final_sum2_kernel<<<1, QV_THREADS, 0, a->stream>>>(a->d_partial, (uint32_t)blocks, a->d_partial + 2 * blocks);
    g_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, a->d_partial + 2 * blocks, 2 * sizeof(double), cudaMemcpyDeviceToHost, a->stream));
    CK(cudaStreamSynchronize(a->stream));
    return 0;
}

int qvmcuda_probabilities(qvmcuda_state* s, double* out, uint64_t offset, uint64_t count) {
    if (!s || (count && !out)) return fail("null argument");
    if (offset > s->n_amps || count > s->n_amps - offset) return fail("range outside the state");
    if (count == 0) return 0;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    if (int rc = canonicalize_locked(s)) return rc;
    // chunks of at most 2^24 basis states through one temporary device buffer (128 MiB)
    const uint64_t chunk = count < (1ull << 24) ? count : (1ull << 24);
    double* d_tmp = nullptr;
    if (int rc0 = aux_locked(s, chunk, &d_tmp)) return rc0;
    int rc = 0;
    for (uint64_t done = 0; done < count && !rc; done += chunk) {
        const uint64_t n = count - done < chunk ? count - done : chunk;
        uint64_t blocks = (n + QV_THREADS - 1) / QV_THREADS;
        const uint64_t cap = (uint64_t)s->sm_count * 8 * 4;
        if (blocks > cap) blocks = cap;
        qv_probs_kernel<<<(int)blocks, QV_THREADS, 0, s->stream>>>(s->d_amps, offset + done, n, d_tmp);
        g_launches++;
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out + done, d_tmp, n * sizeof(double), cudaMemcpyDeviceToHost, s->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s->stream);
        if (e != cudaSuccess) rc = fail(std::string("probabilities: ") + cudaGetErrorString(e));
    }
    return rc;
}

int qvmcuda_scale(qvmcuda_state* s, double factor) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    return elementwise_locked(s, 0, 0, 0, 0, factor);
}

int qvmcuda_normalize(qvmcuda_state* s) {
    double n2 = 0.0;
    if (int rc = qvmcuda_norm2(s, &n2)) return rc;
    return qvmcuda_scale(s, 1.0 / sqrt(n2));
}

int qvmcuda_collapse(qvmcuda_state* s, int qubit, int keep_bit, double inv_norm) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (qubit < 0 || qubit >= (int)s->l2p.size()) return fail("qubit out of range");
    DeviceGuard dg(s->device);
    const int pbit = s->l2p[qubit];
    if (pbit >= s->n_bits) {
        // the qubit selects the rank: the whole shard is either kept (scaled) or annihilated
        const int mine = (s->rank >> (pbit - s->n_bits)) & 1;
        if (mine == (keep_bit ? 1 : 0)) return elementwise_locked(s, 0, 0, 0, 0, inv_norm);
        s->lazy_basis = false;
        CK(cudaMemsetAsync(s->d_amps, 0, s->n_amps * sizeof(qvc), s->stream));
        return 0;
    }
    return elementwise_locked(s, 1, (uint32_t)pbit, 0, keep_bit ? 1u : 0u, inv_norm);
}

// Blocked-order sampler on this handle's vector.  base = mass in front of it (sharded states); total_out (optional)
// receives the vector's own mass in tree order.  n_shots may be 0 (total only).
static int sample_locked(qvmcuda_state* s, const double* uniforms, uint64_t n_shots, uint64_t* out, int strict, double base,
                         double* total_out) {
    if (int rc = materialize_locked(s)) return rc;
    const uint64_t n1 = (s->n_amps + QV_SB - 1) / QV_SB;
    const uint64_t n2 = (n1 + QV_SB - 1) / QV_SB;
    // Persistent scratch (no allocation on the measurement path): the two summation levels + top prefix live
    // with the state, the per-shot buffers grow on demand.
    if (!s->d_sample_tree) CK(cudaMalloc((void**)&s->d_sample_tree, (n1 + 2 * n2) * sizeof(double)));
    if (n_shots > s->shot_cap) {
        CK(cudaStreamSynchronize(s->stream));
        if (s->d_shots) cudaFree(s->d_shots);
        s->d_shots = nullptr;
        s->shot_cap = 0;
        const uint64_t cap = n_shots < 4096 ? 4096 : n_shots + n_shots / 2;
        CK(cudaMalloc((void**)&s->d_shots, cap * (sizeof(double) + sizeof(uint64_t))));
        s->shot_cap = cap;
    }
    double* d_l1 = s->d_sample_tree;
    double* d_l2 = d_l1 + n1;
    double* d_top = d_l2 + n2;
    uint64_t* d_out = reinterpret_cast<uint64_t*>(s->d_shots);
    double* d_u = reinterpret_cast<double*>(d_out + s->shot_cap);
    if (n_shots) CK(cudaMemcpyAsync(d_u, uniforms, n_shots * sizeof(double), cudaMemcpyHostToDevice, s->stream));
    auto warps_grid = [&](uint64_t blocks) {
        uint64_t g = (blocks * 32 + QV_THREADS - 1) / QV_THREADS;
        const uint64_t cap = (uint64_t)s->sm_count * 8 * 4;
        return (int)(g > cap ? cap : (g ? g : 1));
    };
    qv_sample_build_kernel<<<warps_grid(n1), QV_THREADS, 0, s->stream>>>(s->d_amps, nullptr, s->n_amps, d_l1, n1);
    qv_sample_build_kernel<<<warps_grid(n2), QV_THREADS, 0, s->stream>>>(nullptr, d_l1, n1, d_l2, n2);
    qv_sample_top_kernel<<<1, 32, 0, s->stream>>>(d_l2, n2, d_top);
    g_launches += 3;
    if (n_shots) {
        qv_sample_descend_kernel<<<(int)((n_shots + 127) / 128), 128, 0, s->stream>>>(s->d_amps, s->n_amps, d_l1, n1, d_top, n2, d_u, n_shots,
                                                                                   strict ? 1 : 0, base, d_out);
        g_launches++;
    }
    CK(cudaGetLastError());
    if (n_shots) CK(cudaMemcpyAsync(out, d_out, n_shots * sizeof(uint64_t), cudaMemcpyDeviceToHost, s->stream));
    if (total_out) CK(cudaMemcpyAsync(total_out, d_top + (n2 - 1), sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int qvmcuda_sample(qvmcuda_state* s, const double* uniforms, uint64_t n_shots, uint64_t* out, int strict) {
    if (!s || (n_shots && (!uniforms || !out))) return fail("null argument");
    if (n_shots == 0) return 0;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    if (int rc = canonicalize_locked(s)) return rc;
    return sample_locked(s, uniforms, n_shots, out, strict, 0.0, nullptr);
}

int qvmcuda_sample_total(qvmcuda_state* s, double* total) {
    if (!s || !total) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    return sample_locked(s, nullptr, 0, nullptr, 0, 0.0, total);
}

int qvmcuda_sample_shard(qvmcuda_state* s, const double* uniforms, uint64_t n_shots, uint64_t* out, int strict, double base) {
    if (!s || (n_shots && (!uniforms || !out))) return fail("null argument");
    if (n_shots == 0) return 0;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    return sample_locked(s, uniforms, n_shots, out, strict, base, nullptr);
}

// ------------------------------------------------------------------ density matrix
// Gates on the 2n index bits of vec(rho) for one operator rho <- sum_j K_j rho K_j^dagger on QUBITS (nat-tuple order).
static int density_gates_for_op(int n_qubits, int k, const int32_t* qubits, int m, const double* kraus, std::vector<qv::Gate>& gates) {
    if (k < 1 || k > 8) return fail("kraus operator arity out of range (1..8)");
    for (int j = 0; j < k; j++)
        if (qubits[j] < 0 || qubits[j] >= n_qubits) return fail("qubit out of range");
    if (m == 0) return 0;
    const size_t d = (size_t)1 << k;
    const qv::cd* K = reinterpret_cast<const qv::cd*>(kraus);
    if (m == 1) {
        // single-kraus: conj(K) on the column bits, then K on the row bits (src/apply-gate.lisp:56-65)
        qv::Gate gc, gr;
        gc.mat.resize(d * d);
        gr.mat.resize(d * d);
        for (size_t i = 0; i < d * d; i++) {
            gc.mat[i] = std::conj(K[i]);
            gr.mat[i] = K[i];
        }
        for (int j = 0; j < k; j++) {
            gc.qubits.push_back(qubits[j]);
            gr.qubits.push_back(qubits[j] + n_qubits);
        }
        gates.push_back(std::move(gc));
        gates.push_back(std::move(gr));
    } else {
        // kraus-list: ONE superoperator S = sum_j K_j (x) conj(K_j) on (row bits, column bits)
        // instead of the reference's copy / apply / add / restore loop (src/apply-gate.lisp:79-99).
        qv::Gate g;
        const size_t D = d * d;
        g.mat.assign(D * D, qv::cd(0.0, 0.0));
        for (int j = 0; j < m; j++) {
            const qv::cd* Kj = K + (size_t)j * d * d;
            for (size_t r1 = 0; r1 < d; r1++)
                for (size_t c1 = 0; c1 < d; c1++)
                    for (size_t r2 = 0; r2 < d; r2++)
                        for (size_t c2 = 0; c2 < d; c2++)
                            g.mat[((r1 << k) | c1) * D + ((r2 << k) | c2)] += Kj[r1 * d + r2] * std::conj(Kj[c1 * d + c2]);
        }
        for (int j = 0; j < k; j++) g.qubits.push_back(qubits[j]);
        for (int j = 0; j < k; j++) g.qubits.push_back(qubits[j] + n_qubits);
        gates.push_back(std::move(g));
    }
    return 0;
}

int qvmcuda_density_apply_kraus(qvmcuda_state* s, int n_qubits, int k, const int32_t* qubits, int m,
                                const double* kraus, uint32_t flags) {
    if (!s || !qubits || (m > 0 && !kraus)) return fail("null argument");
    if (s->n_bits != 2 * n_qubits) return fail("state length is not 4^n_qubits");
    std::vector<qv::Gate> gates;
    if (int rc = density_gates_for_op(n_qubits, k, qubits, m, kraus, gates)) return rc;
    if (gates.empty()) return 0;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    return run_gates_locked(s, gates, flags | QVMCUDA_FUSE);
}

int qvmcuda_density_apply_ops(qvmcuda_state* s, int n_qubits, int n_ops, const int32_t* ks, const int32_t* qubits,
                              const int32_t* ms, const double* kraus, uint32_t flags) {
    if (!s || (n_ops > 0 && (!ks || !qubits || !ms || !kraus))) return fail("null argument");
    if (n_ops < 0) return fail("negative operator count");
    if (s->n_bits != 2 * n_qubits) return fail("state length is not 4^n_qubits");
    std::vector<qv::Gate> gates;
    size_t qo = 0, ko = 0;
    for (int i = 0; i < n_ops; i++) {
        if (ks[i] < 1 || ks[i] > 8) return fail("kraus operator arity out of range (1..8)");
        if (ms[i] < 0) return fail("negative kraus count");
        if (int rc = density_gates_for_op(n_qubits, ks[i], qubits + qo, ms[i], kraus + ko, gates)) return rc;
        qo += (size_t)ks[i];
        ko += (size_t)2 * ((size_t)1 << (2 * ks[i])) * (size_t)ms[i];
    }
    if (gates.empty()) return 0;
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    return run_gates_locked(s, gates, flags);
}

int qvmcuda_density_prob_excited(qvmcuda_state* s, int n_qubits, int qubit, double* p) {
    if (!s || !p) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->n_bits != 2 * n_qubits) return fail("state length is not 4^n_qubits");
    if (qubit < 0 || qubit >= n_qubits) return fail("qubit out of range");
    DeviceGuard dg(s->device);
    if (int rc = canonicalize_locked(s)) return rc;
    return reduce_locked(s, (1ull << n_qubits) / 2, 2, (uint32_t)qubit, 1ull << n_qubits, p);
}

int qvmcuda_density_collapse(qvmcuda_state* s, int n_qubits, int qubit, int keep_bit, double inv_norm) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->n_bits != 2 * n_qubits) return fail("state length is not 4^n_qubits");
    if (qubit < 0 || qubit >= n_qubits) return fail("qubit out of range");
    DeviceGuard dg(s->device);
    if (int rc = canonicalize_locked(s)) return rc;
    return elementwise_locked(s, 2, (uint32_t)qubit, (uint32_t)(qubit + n_qubits), keep_bit ? 1u : 0u, inv_norm);
}

int qvmcuda_density_measure_discard(qvmcuda_state* s, int n_qubits, int qubit) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->n_bits != 2 * n_qubits) return fail("state length is not 4^n_qubits");
    if (qubit < 0 || qubit >= n_qubits) return fail("qubit out of range");
    DeviceGuard dg(s->device);
    if (int rc = canonicalize_locked(s)) return rc;
    return elementwise_locked(s, 3, (uint32_t)qubit, (uint32_t)(qubit + n_qubits), 0, 1.0);
}

int qvmcuda_density_diag_probs(qvmcuda_state* s, int n_qubits, double* out) {
    if (!s || !out) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->n_bits != 2 * n_qubits) return fail("state length is not 4^n_qubits");
    DeviceGuard dg(s->device);
    if (int rc = canonicalize_locked(s)) return rc;
    const uint64_t dim = 1ull << n_qubits;
    double* d_out = nullptr;
    if (int rc = aux_locked(s, dim, &d_out)) return rc;
    qv_diag_probs_kernel<<<(int)((dim + QV_THREADS - 1) / QV_THREADS), QV_THREADS, 0, s->stream>>>(s->d_amps, dim, d_out);
    g_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, d_out, dim * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int qvmcuda_density_expectation(qvmcuda_state* s, int n_qubits, const double* op_matrix, double out[2]) {
    if (!s || !op_matrix || !out) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->n_bits != 2 * n_qubits) return fail("state length is not 4^n_qubits");
    DeviceGuard dg(s->device);
    if (int rc = canonicalize_locked(s)) return rc;
    double* d_q = nullptr;
    if (int rc = aux_locked(s, 2 * s->n_amps, &d_q)) return rc;      // the operator matrix, as large as rho itself
    CK(cudaMemcpyAsync(d_q, op_matrix, s->n_amps * sizeof(qvc), cudaMemcpyHostToDevice, s->stream));
    uint64_t blocks = (s->n_amps + QV_THREADS - 1) / QV_THREADS;
    if (blocks > (kReduceBlocks - 2) / 2) blocks = (kReduceBlocks - 2) / 2;
    qv_trace_product_kernel<<<(int)blocks, QV_THREADS, 0, s->stream>>>(s->d_amps, reinterpret_cast<const qvc*>(d_q), (uint32_t)n_qubits, s->d_partial);
    qv_final_sum2_kernel<<<1, QV_THREADS, 0, s->stream>>>(s->d_partial, (uint32_t)blocks, s->d_partial + 2 * blocks);
    g_launches += 2;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, s->d_partial + 2 * blocks, 2 * sizeof(double), cudaMemcpyDeviceToHost, s->stream));
    CK(cudaStreamSynchronize(s->stream));
    return 0;
}

int qvmcuda_set_identity_matrix(qvmcuda_state* s, int n_qubits) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->n_bits != 2 * n_qubits) return fail("state length is not 4^n_qubits");
    DeviceGuard dg(s->device);
    for (size_t i = 0; i < s->l2p.size(); i++) s->l2p[i] = (int)i;
    const uint64_t dim = 1ull << n_qubits;
    s->lazy_basis = false;         // content is replaced
    CK(cudaMemsetAsync(s->d_amps, 0, s->n_amps * sizeof(qvc), s->stream));
    qv_set_identity_kernel<<<(int)((dim + QV_THREADS - 1) / QV_THREADS), QV_THREADS, 0, s->stream>>>(s->d_amps, dim);
    g_launches++;
    CK(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------ multi-GPU shards
int qvmcuda_shard_export(qvmcuda_state* s, uint8_t handle[64]) {
    if (!s || !handle) return fail("null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    if (int rc = materialize_locked(s)) return rc;     // peers will read this memory
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->d_amps));
    std::memcpy(handle, &h, 64);
    return 0;
}

int qvmcuda_shard_attach(qvmcuda_state* s, int rank, int world, const uint8_t* handles) {
    if (!s || !handles) return fail("null argument");
    if (world < 1 || world > QV_MAX_PEERS || (world & (world - 1))) return fail("world size must be 1, 2, 4 or 8");
    if (rank < 0 || rank >= world) return fail("rank out of range");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    drop_tape_cache_locked(s);        // schedules depend on rank / world / the alternate buffer
    for (int r = 0; r < world; r++) {
        if (r == rank) {
            s->peers.base[r] = s->d_amps;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->opened.push_back(p);
        s->peers.base[r] = (qvc*)p;
    }
    s->rank = rank;
    s->world = world;
    const int total = s->n_bits + log2_exact((uint64_t)world);
    s->l2p.resize(total);
    for (int i = 0; i < total; i++) s->l2p[i] = i;
    return 0;
}

int qvmcuda_shard_export_alt(qvmcuda_state* s, uint8_t handle[64]) {
    if (!s || !handle) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    if (!s->d_alt) {
        size_t free_b = 0, total_b = 0;
        CK(cudaMemGetInfo(&free_b, &total_b));
        const size_t need = s->n_amps * sizeof(qvc);
        if (free_b < need + (size_t(2) << 30)) return fail("not enough device memory for an alternate shard buffer");
        CK(cudaMalloc((void**)&s->d_alt, need));
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, s->d_alt));
    std::memcpy(handle, &h, 64);
    return 0;
}

int qvmcuda_shard_attach_alt(qvmcuda_state* s, const uint8_t* handles) {
    if (!s || !handles) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->world < 2 || !s->d_alt) return fail("attach the shards and export the alternate buffer first");
    DeviceGuard dg(s->device);
    drop_tape_cache_locked(s);        // schedules depend on rank / world / the alternate buffer
    for (int r = 0; r < s->world; r++) {
        if (r == s->rank) {
            s->peers_alt.base[r] = s->d_alt;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        s->opened.push_back(p);
        s->peers_alt.base[r] = (qvc*)p;
    }
    s->remap_pull = true;
    return 0;
}

int qvmcuda_shard_clear(qvmcuda_state* s) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    DeviceGuard dg(s->device);
    for (size_t i = 0; i < s->l2p.size(); i++) s->l2p[i] = (int)i;   // content is replaced: layout resets
    static const bool lazy = !(getenv("QVMCUDA_LAZY_RESET") && atoi(getenv("QVMCUDA_LAZY_RESET")) == 0);
    s->lazy_basis = true;
    s->lazy_index = kLazyAllZero;
    if (!lazy || s->world < 2 || !s->remap_pull) return materialize_locked(s);
    return 0;
}

int qvmcuda_shard_set_zero_ranks(qvmcuda_state* s, uint32_t mask) {
    if (!s) return fail("null argument");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->world < 2) return fail("not a shard");
    s->zero_ranks = mask & ((1u << s->world) - 1u);
    return 0;
}

// One host process driving several devices (the shape of a single Lisp image with N GPUs): the shards are states of the SAME
// process, so peers are reached through plain peer access instead of IPC handles.  With want_alt every state gets the alternate
// buffer (pull remaps), all or none.
int qvmcuda_shard_attach_local(qvmcuda_state* const* states, int world, int want_alt) {
    if (!states) return fail("null argument");
    if (world < 2 || world > QV_MAX_PEERS || (world & (world - 1))) return fail("world size must be 2, 4 or 8");
    for (int r = 0; r < world; r++) {
        if (!states[r]) return fail("null shard");
        if (states[r]->n_amps != states[0]->n_amps) return fail("shards differ in length");
        for (int q = 0; q < r; q++)
            if (states[q] == states[r] || states[q]->device == states[r]->device) return fail("every shard needs its own device");
    }
    for (int r = 0; r < world; r++) {
        qvmcuda_state* s = states[r];
        std::lock_guard<std::mutex> lk(s->mu);
        DeviceGuard dg(s->device);
        if (int rc = materialize_locked(s)) return rc;
        for (int q = 0; q < world; q++) {
            if (q == r) continue;
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, s->device, states[q]->device));
            if (!can) return fail("devices cannot access each other's memory");
            const cudaError_t e = cudaDeviceEnablePeerAccess(states[q]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(cudaGetErrorString(e));
            cudaGetLastError();
        }
        if (want_alt && !s->d_alt) {
            const cudaError_t e = cudaMalloc((void**)&s->d_alt, s->n_amps * sizeof(qvc));
            if (e != cudaSuccess) {
                cudaGetLastError();
                want_alt = 0;       // all or none: fall back to in-place exchanges
            }
        }
    }
    for (int r = 0; r < world; r++) {
        qvmcuda_state* s = states[r];
        std::lock_guard<std::mutex> lk(s->mu);
        drop_tape_cache_locked(s);
        for (int q = 0; q < world; q++) {
            s->peers.base[q] = states[q]->d_amps;
            s->peers_alt.base[q] = want_alt ? states[q]->d_alt : nullptr;
        }
        s->rank = r;
        s->world = world;
        s->remap_pull = want_alt != 0;
        const int total = s->n_bits + log2_exact((uint64_t)world);
        s->l2p.resize(total);
        for (int i = 0; i < total; i++) s->l2p[i] = i;
    }
    return 0;
}

}  // extern "C"
