// qv_jit.cpp -- back end of the pass compiler: NVRTC (source -> sm_100a cubin), the kernel cache (memory + disk) and
// the launcher (driver API through cudaGetDriverEntryPoint, so libqvmcuda keeps loading on machines without libcuda).
// See qv_jit.h.
#include "qv_jit.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "qv_program.h"

namespace qv {

struct JitKernel {
    CUmodule mod = nullptr;
    CUfunction fn = nullptr;
    int threads = 0;
    int prog_bytes = 0;
    int tma = 0;
};

namespace {

// ------------------------------------------------------------------------------------------ embedded headers
// The headers a generated pass includes travel inside the library (build/*.inc are the files wrapped in raw
// string literals by the Makefile).
const char kHdrProgram[] =
#include "build/qv_program.h.inc"
    ;
const char kHdrOps[] =
#include "build/qv_ops.h.inc"
    ;
const char kHdrCommon[] =
#include "build/qv_tile_common.cuh.inc"
    ;
const char kHdrPrelude[] =
#include "build/qv_jit_prelude.cuh.inc"
    ;
const char kHdrKernel[] =
#include "build/qv_jit_kernel.cuh.inc"
    ;
const char kHdrKernelTma[] =
#include "build/qv_jit_kernel_tma.cuh.inc"
    ;

// ------------------------------------------------------------------------------------------ NVRTC (dlopen)
struct Nvrtc {
    void* lib = nullptr;
    decltype(&nvrtcCreateProgram) create = nullptr;
    decltype(&nvrtcCompileProgram) compile = nullptr;
    decltype(&nvrtcGetCUBINSize) cubin_size = nullptr;
    decltype(&nvrtcGetCUBIN) cubin = nullptr;
    decltype(&nvrtcGetProgramLogSize) log_size = nullptr;
    decltype(&nvrtcGetProgramLog) log = nullptr;
    decltype(&nvrtcDestroyProgram) destroy = nullptr;
    decltype(&nvrtcGetErrorString) errstr = nullptr;
    std::string error;
};

Nvrtc& nvrtc() {
    static Nvrtc n;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("QVMCUDA_NVRTC");
        const char* names[] = {env, "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so",
                               "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* nm : names) {
            if (!nm || !*nm) continue;
            n.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
            if (n.lib) break;
        }
        if (!n.lib) {
            n.error = "libnvrtc.so.12 not found (set QVMCUDA_NVRTC)";
            return;
        }
#define QV_SYM(field, name)                                                         \
    n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.lib, name));              \
    if (!n.field) n.error = std::string("libnvrtc lacks ") + name;
        QV_SYM(create, "nvrtcCreateProgram")
        QV_SYM(compile, "nvrtcCompileProgram")
        QV_SYM(cubin_size, "nvrtcGetCUBINSize")
        QV_SYM(cubin, "nvrtcGetCUBIN")
        QV_SYM(log_size, "nvrtcGetProgramLogSize")
        QV_SYM(log, "nvrtcGetProgramLog")
        QV_SYM(destroy, "nvrtcDestroyProgram")
        QV_SYM(errstr, "nvrtcGetErrorString")
#undef QV_SYM
    });
    return n;
}

// ------------------------------------------------------------------------------------------ driver API
struct Driver {
    CUresult (*moduleLoadData)(CUmodule*, const void*) = nullptr;
    CUresult (*moduleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*funcSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
    CUresult (*launchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**,
                             void**) = nullptr;
    CUresult (*getErrorString)(CUresult, const char**) = nullptr;
    CUresult (*tensorMapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill) = nullptr;
    std::string error;
};

Driver& driver() {
    static Driver d;
    static std::once_flag once;
    std::call_once(once, [] {
        auto get = [&](const char* name, void** fn) {
            cudaDriverEntryPointQueryResult q;
            const cudaError_t e = cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q);
            if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) d.error = std::string("no driver entry point ") + name;
        };
        get("cuModuleLoadData", (void**)&d.moduleLoadData);
        get("cuModuleGetFunction", (void**)&d.moduleGetFunction);
        get("cuFuncSetAttribute", (void**)&d.funcSetAttribute);
        get("cuLaunchKernel", (void**)&d.launchKernel);
        get("cuGetErrorString", (void**)&d.getErrorString);
        {   // optional: without it the tiles are loaded with LDGSTS
            const std::string keep = d.error;
            get("cuTensorMapEncodeTiled", (void**)&d.tensorMapEncodeTiled);
            if (d.error != keep) {
                d.error = keep;
                d.tensorMapEncodeTiled = nullptr;
            }
        }
    });
    return d;
}

std::string cu_err(CUresult r) {
    const char* s = nullptr;
    if (driver().getErrorString) driver().getErrorString(r, &s);
    return s ? s : "CUDA driver error " + std::to_string((int)r);
}

// ------------------------------------------------------------------------------------------ cache
enum class St { PENDING, READY, FAILED };

struct Entry {
    St st = St::PENDING;
    std::vector<char> cubin;
    std::string log;
    int threads = 0, prog_bytes = 0;
    int tma = 0;
    std::map<int, std::unique_ptr<JitKernel>> loaded;     // device -> kernel
    bool load_failed = false;
};

std::mutex g_mu;
std::condition_variable g_cv;
std::map<uint64_t, std::shared_ptr<Entry>> g_cache;
JitStats g_stats;
std::string g_last_log;

std::string cache_dir() {
    static std::string dir = [] {
        if (const char* e = getenv("QVMCUDA_JIT_CACHE")) return std::string(e);
        Dl_info info;
        if (dladdr((void*)&cache_dir, &info) && info.dli_fname) {
            std::string p = info.dli_fname;
            const size_t slash = p.rfind('/');
            p = slash == std::string::npos ? std::string(".") : p.substr(0, slash);
            return p + "/jit_cache";
        }
        return std::string();
    }();
    return dir;
}

std::string cache_path(uint64_t sig) {
    const std::string d = cache_dir();
    if (d.empty() || d == "off") return std::string();
    char name[64];
    snprintf(name, sizeof(name), "/qvj_%016llx.cubin", (unsigned long long)sig);
    return d + name;
}

bool disk_read(uint64_t sig, std::vector<char>& out) {
    const std::string p = cache_path(sig);
    if (p.empty()) return false;
    FILE* f = fopen(p.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    bool ok = n > 0;
    if (ok) {
        out.resize((size_t)n);
        ok = fread(out.data(), 1, (size_t)n, f) == (size_t)n;
    }
    fclose(f);
    return ok;
}

void disk_write(uint64_t sig, const std::vector<char>& cubin) {
    const std::string p = cache_path(sig);
    if (p.empty()) return;
    mkdir(cache_dir().c_str(), 0755);
    const std::string tmp = p + ".tmp" + std::to_string((long)getpid()) + "_" + std::to_string((unsigned long)(uintptr_t)&cubin);
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    fclose(f);
    if (ok) rename(tmp.c_str(), p.c_str());
    else unlink(tmp.c_str());
}

// ------------------------------------------------------------------------------------------ compile pool
struct Pool {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    std::vector<std::thread> workers;
    bool stop = false;
    void start() {
        if (!workers.empty()) return;
        int n = getenv("QVMCUDA_JIT_THREADS") ? atoi(getenv("QVMCUDA_JIT_THREADS")) : 0;
        if (n <= 0) n = (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency()));
        for (int i = 0; i < n; i++)
            workers.emplace_back([this] {
                for (;;) {
                    std::function<void()> job;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [this] { return stop || !q.empty(); });
                        if (stop && q.empty()) return;
                        job = std::move(q.front());
                        q.pop_front();
                    }
                    job();
                }
            });
    }
    void submit(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> lk(mu);
            start();
            q.push_back(std::move(f));
        }
        cv.notify_one();
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv.notify_all();
        for (std::thread& t : workers)
            if (t.joinable()) t.join();
    }
};

Pool& pool() {
    static Pool* p = new Pool();     // leaked on purpose: worker threads must not be joined from a static destructor at exit
    return *p;
}

constexpr int kTmaSmemBytes = 2 * 65536 + 1024;      // two tile buffers + slack for the 1024-byte alignment SWIZZLE_128B needs

bool trace_on() {
    static const bool t = getenv("QVMCUDA_TRACE") != nullptr;
    return t;
}

// Generator variant of a step: QVMCUDA_JIT_VARIANT forces one; otherwise
//  * the tile is loaded by ONE tensor copy (bit 16) wherever the driver offers cuTensorMapEncodeTiled and the tile is a box of a
//    5-d view of the state -- measured 36.8 vs 37.1 ms per QFT-30, 3 % fewer instructions (gpurun_out/r2j_*); the persistent
//    double-buffered TMA kernel (bit 8) lost: 53.0 ms, one 16-warp CTA per SM cannot hide the latencies of the rounds that three
//    independent CTAs hide (profiles/r02_tma_variants.md), so it is never chosen automatically;
//  * passes whose FP64 work dominates (layers of dense gates: bound by the FP64 pipe, not by HBM) get two CTAs per SM with up to
//    128 registers per thread (bit 2) -- 6 % faster on 25-qubit random layers (gpurun_out/r2b_configs_v{0,2}.jsonl) -- everything
//    else three CTAs with 80.
int variant_of(const Step& st) {
    static const int forced = getenv("QVMCUDA_JIT_VARIANT") ? atoi(getenv("QVMCUDA_JIT_VARIANT")) : -1;
    if (forced >= 0) return forced;
    Tape one;
    one.steps.push_back(st);
    int v = tape_cost(one) > 180.0 ? 2 : 0;
    if (driver().error.empty() && driver().tensorMapEncodeTiled) v |= 16;
    return v;
}

uint32_t real_uops(const Step& st) {
    QvPassHeader h;
    std::memcpy(&h, st.blob.data(), sizeof(h));
    return h.n_uops - h.n_rounds;       // every round's list ends with a sentinel
}

JitKernel* load_locked(Entry& e, int device) {
    auto it = e.loaded.find(device);
    if (it != e.loaded.end()) return it->second.get();
    if (e.load_failed) return nullptr;
    Driver& d = driver();
    if (!d.error.empty()) {
        e.load_failed = true;
        g_last_log = d.error;
        return nullptr;
    }
    auto k = std::make_unique<JitKernel>();
    CUresult r = d.moduleLoadData(&k->mod, e.cubin.data());
    if (r == CUDA_SUCCESS) r = d.moduleGetFunction(&k->fn, k->mod, "qvj_kernel");
    if (r == CUDA_SUCCESS)
        r = d.funcSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, e.tma == 2 ? kTmaSmemBytes : 65536);
    if (r == CUDA_SUCCESS) r = d.funcSetAttribute(k->fn, CU_FUNC_ATTRIBUTE_PREFERRED_SHARED_MEMORY_CARVEOUT, 100);
    if (r != CUDA_SUCCESS) {
        e.load_failed = true;
        g_last_log = "loading a compiled pass: " + cu_err(r);
        if (trace_on()) fprintf(stderr, "[qvjit] %s\n", g_last_log.c_str());
        return nullptr;
    }
    k->threads = e.threads;
    k->prog_bytes = e.prog_bytes;
    k->tma = e.tma;
    JitKernel* out = k.get();
    e.loaded[device] = std::move(k);
    return out;
}

}  // namespace

bool jit_compile_cubin(const JitSource& src, std::vector<char>& cubin, std::string& log) {
    Nvrtc& n = nvrtc();
    if (!n.error.empty()) {
        log = n.error;
        return false;
    }
    const char* hdr_src[] = {kHdrProgram, kHdrOps, kHdrCommon, kHdrPrelude, kHdrKernel, kHdrKernelTma};
    const char* hdr_name[] = {"qv_program.h", "qv_ops.h", "qv_tile_common.cuh", "qv_jit_prelude.cuh", "qv_jit_kernel.cuh",
                              "qv_jit_kernel_tma.cuh"};
    nvrtcProgram prog = nullptr;
    nvrtcResult r = n.create(&prog, src.text.c_str(), "qvj_pass.cu", 6, hdr_src, hdr_name);
    if (r != NVRTC_SUCCESS) {
        log = std::string("nvrtcCreateProgram: ") + n.errstr(r);
        return false;
    }
    std::vector<const char*> opts = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo", "--ptxas-options=-v", "-default-device"};
    if (const char* extra = getenv("QVMCUDA_JIT_MAXREG")) {
        static thread_local std::string o;
        o = std::string("--maxrregcount=") + extra;
        opts.push_back(o.c_str());
    }
    r = n.compile(prog, (int)opts.size(), opts.data());
    size_t ls = 0;
    n.log_size(prog, &ls);
    if (ls > 1) {
        log.resize(ls);
        n.log(prog, &log[0]);
    }
    bool ok = r == NVRTC_SUCCESS;
    if (ok) {
        size_t cs = 0;
        ok = n.cubin_size(prog, &cs) == NVRTC_SUCCESS && cs > 0;
        if (ok) {
            cubin.resize(cs);
            ok = n.cubin(prog, cubin.data()) == NVRTC_SUCCESS;
        }
        if (!ok) log += "\nnvrtcGetCUBIN failed";
    } else {
        log += std::string("\nnvrtcCompileProgram: ") + n.errstr(r);
    }
    n.destroy(&prog);
    return ok;
}

JitPolicy jit_policy() {
    static const JitPolicy p = [] {
        const char* e = getenv("QVMCUDA_JIT");
        if (!e) return JitPolicy::SYNC;
        const std::string s = e;
        if (s == "0" || s == "off") return JitPolicy::OFF;
        if (s == "async") return JitPolicy::ASYNC;
        return JitPolicy::SYNC;
    }();
    return p;
}

uint32_t jit_min_uops() {
    static const uint32_t m = getenv("QVMCUDA_JIT_MIN_UOPS") ? (uint32_t)atoi(getenv("QVMCUDA_JIT_MIN_UOPS")) : 3u;
    return m;
}

void jit_prepare(const std::vector<const Step*>& steps, int device, std::vector<JitKernel*>& out, int extra_variant) {
    out.assign(steps.size(), nullptr);
    const JitPolicy pol = jit_policy();
    if (pol == JitPolicy::OFF) return;
    std::vector<std::shared_ptr<Entry>> ent(steps.size());
    std::vector<uint64_t> sigs(steps.size(), 0);
    for (size_t i = 0; i < steps.size(); i++) {
        const Step& st = *steps[i];
        if (st.kind != Step::TILE || real_uops(st) < jit_min_uops()) continue;
        JitSource src = jit_generate(st, variant_of(st) | extra_variant);
        if (!src.ok) continue;
        sigs[i] = src.sig;
        std::unique_lock<std::mutex> lk(g_mu);
        auto it = g_cache.find(src.sig);
        if (it != g_cache.end()) {
            ent[i] = it->second;
            g_stats.cache_hits++;
            continue;
        }
        auto e = std::make_shared<Entry>();
        e->threads = src.threads;
        e->prog_bytes = src.prog_bytes;
        e->tma = src.tma;
        g_cache[src.sig] = e;
        ent[i] = e;
        lk.unlock();
        if (disk_read(src.sig, e->cubin)) {
            std::lock_guard<std::mutex> l2(g_mu);
            e->st = St::READY;
            g_stats.disk_hits++;
            continue;
        }
        auto job = [e, src = std::move(src)]() {
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<char> cubin;
            std::string log;
            const bool ok = jit_compile_cubin(src, cubin, log);
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (ok) disk_write(src.sig, cubin);
            if (trace_on() || !ok)
                fprintf(stderr, "[qvjit] pass %016llx %s in %.2f s%s%s\n", (unsigned long long)src.sig, ok ? "compiled" : "FAILED", dt,
                        (ok && !trace_on()) ? "" : "\n", (ok && !trace_on()) ? "" : log.c_str());
            {
                std::lock_guard<std::mutex> lk(g_mu);
                e->cubin = std::move(cubin);
                e->log = log;
                e->st = ok ? St::READY : St::FAILED;
                g_last_log = log;
                g_stats.compile_seconds += dt;
                if (ok) g_stats.compiled++;
                else g_stats.failed++;
            }
            g_cv.notify_all();
        };
        pool().submit(std::move(job));
    }
    std::unique_lock<std::mutex> lk(g_mu);
    for (size_t i = 0; i < steps.size(); i++) {
        if (!ent[i]) continue;
        if (pol == JitPolicy::SYNC) g_cv.wait(lk, [&] { return ent[i]->st != St::PENDING; });
        if (ent[i]->st == St::READY) out[i] = load_locked(*ent[i], device);
    }
}

void jit_precompile(const std::vector<const Step*>& steps, int& n_eligible, int& n_ok, std::string& log, bool basis_first) {
    n_eligible = n_ok = 0;
    struct Job {
        JitSource src;
        bool ok = false, done = false;
        std::string log;
    };
    std::vector<std::shared_ptr<Job>> jobs;
    std::map<uint64_t, bool> seen;
    std::mutex mu;
    std::condition_variable cv;
    for (const Step* st : steps) {
        if (st->kind != Step::TILE || real_uops(*st) < jit_min_uops()) continue;
        // without a driver (build container) the variant the GPU box will pick is unknown: compile with and without the
        // single-copy tile load
        std::vector<int> variants = {variant_of(*st)};
        if (!(variants[0] & 16) && !getenv("QVMCUDA_JIT_VARIANT")) variants.push_back(variants[0] | 16);
        // the first pass of a tape may start from a lazily reset state (no tile loads)
        if (basis_first && st == steps.front()) variants.push_back(variants[0] | kVariantSrcBasis);
        for (int v : variants) {
        JitSource src = jit_generate(*st, v);
        if (!src.ok || seen.count(src.sig)) continue;
        seen[src.sig] = true;
        n_eligible++;
        std::vector<char> have;
        if (disk_read(src.sig, have)) {
            n_ok++;
            continue;
        }
        auto j = std::make_shared<Job>();
        j->src = std::move(src);
        jobs.push_back(j);
        pool().submit([j, &mu, &cv] {
            std::vector<char> cubin;
            std::string lg;
            const bool ok = jit_compile_cubin(j->src, cubin, lg);
            if (ok) disk_write(j->src.sig, cubin);
            {
                std::lock_guard<std::mutex> lk(mu);
                j->ok = ok;
                j->log = std::move(lg);
                j->done = true;
            }
            cv.notify_all();
        });
        }
    }
    std::unique_lock<std::mutex> lk(mu);
    for (auto& j : jobs) {
        cv.wait(lk, [&] { return j->done; });
        if (j->ok) n_ok++;
        char head[64];
        snprintf(head, sizeof(head), "== pass %016llx: %s\n", (unsigned long long)j->src.sig, j->ok ? "ok" : "FAILED");
        log += head;
        log += j->log;
        log += "\n";
    }
}

const char* jit_launch(JitKernel* k, const JitLaunch& L) {
    static thread_local std::vector<uint8_t> prog;
    static thread_local std::string err;
    if (L.blob_bytes > (size_t)k->prog_bytes) return "control program larger than the compiled pass expects";
    if (prog.size() < (size_t)k->prog_bytes) prog.resize((size_t)k->prog_bytes);
    std::memcpy(prog.data(), L.blob, L.blob_bytes);
    const QvPeers* peers = L.peers;
    const qvc* tables = L.tables;
    qvc* alt = L.alt_own;
    CUresult r;
    if (k->tma) {
        // the tile as a box of a 5-d view of this device's amplitudes (QvTmaGeom: geometry is data, not part of the kernel)
        QvPassHeader h;
        std::memcpy(&h, L.blob, sizeof(h));
        QvTmaGeom geom;
        if (!jit_tma_geometry(h, geom)) return "pass geometry does not fit a tensor map";
        alignas(64) CUtensorMap tmap;
        cuuint64_t gdim[5] = {16, 1, 1, 1, 1}, gstride[4] = {128, 128, 128, 128};
        cuuint32_t box[5] = {16, 1, 1, 1, 1}, estride[5] = {1, 1, 1, 1, 1};
        const cuuint64_t total_bytes = (cuuint64_t)16 << h.n_local_bits;
        for (int i = 0; i < 4; i++) {
            if (geom.len[i]) {
                gdim[i + 1] = (cuuint64_t)1 << geom.len[i];
                gstride[i] = (cuuint64_t)16 << geom.start[i];
                box[i + 1] = geom.is_tile[i] ? (cuuint32_t)1 << geom.len[i] : 1u;
            } else {
                gstride[i] = total_bytes;      // unused dimension of extent 1
            }
        }
        void* base = (void*)peers->base[(h.fixed_bits >> h.n_local_bits) & (QV_MAX_PEERS - 1)];
        r = driver().tensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 5, base, gdim, gstride, box, estride,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            err = "cuTensorMapEncodeTiled: " + cu_err(r);
            return err.c_str();
        }
        void* params[6] = {prog.data(), const_cast<QvPeers*>(peers), &tables, &alt, &tmap, &geom};
        const unsigned grid = k->tma == 2 ? (unsigned)std::min<uint64_t>(h.n_tiles, (uint64_t)L.sm_count) : (unsigned)L.grid;
        const unsigned smem = k->tma == 2 ? (unsigned)kTmaSmemBytes : (unsigned)L.smem;
        r = driver().launchKernel(k->fn, grid, 1, 1, (unsigned)k->threads, 1, 1, smem, (CUstream)L.stream, params, nullptr);
    } else {
        void* params[4] = {prog.data(), const_cast<QvPeers*>(peers), &tables, &alt};
        r = driver().launchKernel(k->fn, (unsigned)L.grid, 1, 1, (unsigned)k->threads, 1, 1, (unsigned)L.smem, (CUstream)L.stream, params,
                                  nullptr);
    }
    if (r != CUDA_SUCCESS) {
        err = cu_err(r);
        return err.c_str();
    }
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_stats.launches++;
    }
    return nullptr;
}

JitStats jit_stats() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_stats;
}

std::string jit_last_log() {
    std::lock_guard<std::mutex> lk(g_mu);
    return g_last_log;
}

}  // namespace qv
