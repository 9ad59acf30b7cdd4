// qv_jit.h -- the pass compiler: tile program -> specialised sm_100a kernel (NVRTC), cached by structure.
//
// The reference compiles every gate of a loaded program into a Lisp function at run time and caches the result by
// operator and qubit tuple (src/compile-gate.lisp:156-209, 315-361; driven by COMPILE-LOADED-PROGRAM, src/qvm.lisp:166-175).
// This is the same idea one level up and on the device: a whole fused pass becomes one straight-line kernel whose only
// run-time inputs are numbers (matrices, tables, tile geometry).  The interpreter kernel (qv_tile_kernel.cuh) stays as
// tier 0: it runs every pass the compiler does not cover (partial tiles, in-place peer passes), light passes that are
// HBM-bound anyway, and everything while a kernel is still being compiled in asynchronous mode.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "qv_sched.h"

#define QVJIT_VERSION "qvjit-10"

struct QvPeers;
struct qvc;

namespace qv {

struct JitSource {
    bool ok = false;
    std::string why_not;
    std::string text;       // the translation unit (includes qv_jit_prelude.cuh / qv_jit_kernel.cuh by name)
    uint64_t sig = 0;       // hash of text + compiler version: the cache key
    int mode = 0;           // 0 local pass, 2 pull pass
    int prog_bytes = 0;     // size of the kernel's control-program parameter
    int threads = 0;
    int tma = 0;            // 0 LDGSTS tile loads; 1 classic kernel with one cp.async.bulk.tensor per tile; 2 persistent double-buffered
                            // kernel fed by cp.async.bulk.tensor (1 and 2 need a tensor map at launch)
};

// Front end (no CUDA needed; also used by the test emulator, which compiles the text for the host).
// variant bits: 1 rolled group loop, 2 two CTAs per SM, 4 fences between micro-ops (experiments), 8 persistent TMA-fed kernel,
// 16 classic kernel with the tile loaded by one tensor copy (8 / 16: local passes whose tile is a box of a <= 5-dimensional
// view of the state; other passes fall back to the LDGSTS loads), 32 no tile loads: the source is a basis state that only exists
// as a flag (lazy reset).
JitSource jit_generate(const Step& st, int variant = 0);

// Tile geometry for the tensor-memory accelerator; false when the tile's bit runs do not fit five dimensions.
bool jit_tma_geometry(const QvPassHeader& h, QvTmaGeom& geom);

// NVRTC: source -> cubin for sm_100a.  Works without a GPU (build-time prewarming, CPU tests).  log receives the
// compiler's output (ptxas -v statistics included).
bool jit_compile_cubin(const JitSource& src, std::vector<char>& cubin, std::string& log);

enum class JitPolicy { OFF = 0, SYNC = 1, ASYNC = 2 };
JitPolicy jit_policy();                 // QVMCUDA_JIT = off | sync (default) | async
uint32_t jit_min_uops();                // passes with fewer micro-ops stay on the interpreter (QVMCUDA_JIT_MIN_UOPS)

struct JitKernel;                       // one loaded kernel on one device

// Make the kernels of these steps available on `device` according to the policy: SYNC compiles the missing ones now
// (in parallel) and returns when all are loaded; ASYNC queues them and returns at once.  out[i] = the kernel of
// steps[i] or nullptr (not eligible, below the threshold, still compiling, or failed -- the caller interprets).
// extra_variant: generator variant bits OR'd into every step's own choice (kVariantSrcBasis for the first pass after a lazy reset).
constexpr int kVariantSrcBasis = 32;
void jit_prepare(const std::vector<const Step*>& steps, int device, std::vector<JitKernel*>& out, int extra_variant = 0);

// Compile the eligible steps into the disk cache only (no device): n_eligible / n_ok count distinct kernels.
void jit_precompile(const std::vector<const Step*>& steps, int& n_eligible, int& n_ok, std::string& log, bool basis_first = false);

struct JitLaunch {
    const uint8_t* blob;
    size_t blob_bytes;
    int grid;
    size_t smem;
    int sm_count;
    void* stream;
    const QvPeers* peers;
    const qvc* tables;
    qvc* alt_own;
};
// nullptr on success, else an error string
const char* jit_launch(JitKernel* k, const JitLaunch& L);

struct JitStats {
    uint64_t compiled = 0, cache_hits = 0, disk_hits = 0, failed = 0, launches = 0;
    double compile_seconds = 0.0;
};
JitStats jit_stats();
std::string jit_last_log();

}  // namespace qv
