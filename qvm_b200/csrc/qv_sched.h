// qv_sched.h -- host-side gate scheduler: logical gates -> tile programs.
//
// Replaces, for the GPU path, what the reference does in
// src/compile-gate.lisp:315-361,409-526 (per-gate compiled lambdas) and the
// external quil::fuse-gates-in-executable-code (called src/qvm.lisp:166-175):
// it packs runs of gates into passes/rounds of the tile kernel (see qv_program.h); gates that act on the same one or
// two qubits back to back are first multiplied together on the host, as the reference's fusion does.
#pragma once
#include <complex>
#include <cstdint>
#include <string>
#include <vector>

#include "qv_program.h"

namespace qv {

using cd = std::complex<double>;

// A gate on LOGICAL qubits.  qubits[j] is attached to bit j of the matrix index
// (NAT-TUPLE order, src/utilities.lisp:43-51); mat is row-major 2^k x 2^k.
struct Gate {
    std::vector<int> qubits;
    std::vector<cd> mat;
};

struct CompileOptions {
    int tile_bits = QV_MAX_TILE_BITS;
    int min_low_bits = QV_MIN_LOW_BITS;
    bool fuse = true;            // false: every gate is its own HBM pass
    bool absorb_swaps = false;   // true: exact SWAP gates only relabel qubits (l2p changes)
    int n_local_bits = 0;        // log2(amplitudes on this device); 0 = n_bits (single device)
    int rank = 0;                // value of the physical bits >= n_local_bits on this device
    int reg_bits = 0;            // amplitudes per thread per round = 2^reg_bits: 3, 4, or 0 = chosen per pass
    bool butterflies = true;     // Hadamard-like gates as unscaled butterflies + one write-back scale per pass
    bool store_perm = true;      // fold trailing X / CNOT / SWAP gates of a pass into its write-back addressing
    bool fuse_pull = true;       // with remap_pull: merge a pull remap into the tile pass that follows it
    bool relabel_global_swaps = true;   // sharded: an exact SWAP touching a rank bit only exchanges the two wires' physical bits
    bool fuse_matrices = true;   // with fuse: multiply neighbouring dense / diagonal atoms on the same <= 2 wires together on the host
    int euler_split = 0;         // complex 1q unitaries as diag . real rotation . diag: 1 always, 0 never, -1 = whichever tape the
                                 // cost model prefers.  Off: measured slower on B200 (25-qubit random layers 6.3 vs 5.5 ms,
                                 // gpurun_out/r2c_configs_euler*.jsonl) -- the extra table lookups cost more than the FP64 saved
    int route_swaps = -1;        // single device, fused: absorb exact SWAP gates as relabelings and execute the permutation back to the
                                 // canonical layout in the write-back of the gate passes (spare tile slots): 1 always, 0 never (SWAPs
                                 // run where they stand), -1 = whichever tape the cost model prefers
    int hoist_remaps = -1;       // sharded, fused pulls: do an exchange a later atom needs as part of the CURRENT pass when that pass then
                                 // holds everything it held before and more (the exchange is NVLink-bound whatever it computes):
                                 // 1 always, 0 never, -1 = whichever tape the cost model prefers
    bool remap_pull = false;     // remaps as out-of-place pulls into the alternate buffer (needs 2x shard memory)
};

struct Step {
    enum Kind { TILE = 0, BIG = 1, REMAP = 2 } kind = TILE;
    std::vector<uint8_t> blob;   // TILE: QvPassHeader followed by rounds/micro-ops/descriptors/matrices (<= QV_PROG_LARGE_BYTES)
    std::vector<cd> tables;      // TILE: diagonal factor tables (global memory)
    QvBigGate big{};             // BIG
    QvRemap remap{};             // REMAP (pull remap into the alternate buffer; all ranks flip afterwards)
    std::vector<cd> bigmat;      // BIG: row-major 2^k x 2^k
    int n_gates = 0;             // logical gates (atoms) folded into this step
    bool uses_peers = false;     // TILE: the tile spans rank bits -> peer shards are read/written (needs barriers)
    bool is_remap = false;       // TILE: a global<->local qubit exchange inserted by the scheduler
};

struct Tape {
    int n_bits = 0;
    std::vector<Step> steps;
    std::vector<int> l2p;        // logical -> physical qubit map after the tape ran
    int n_gates = 0;
    int n_atoms = 0;
    int n_fused = 0;             // atoms merged away by host-side matrix fusion
    int n_split = 0;             // complex 1q unitaries written as diag . rotation . diag (Euler split)
    int n_splittable = 0;        // ... that could have been
    int n_routed = 0;            // transpositions of index bits executed by write-backs on behalf of absorbed SWAP gates (swap routing)
    int n_hoisted = 0;           // exchanges moved forward into a gate pass (remap hoisting)
    int n_relabeled = 0;         // SWAP gates on rank bits executed as relabelings (sharded)
};

// l2p_in: current logical->physical map (empty = identity).
Tape compile(const std::vector<Gate>& gates, int n_bits, const CompileOptions& opt,
             const std::vector<int>& l2p_in = {});

std::string describe(const Tape& t);

// estimated cost in FP64 instructions per amplitude (used to choose between equivalent schedules)
double tape_cost(const Tape& t);

}  // namespace qv
