// qv_program.h -- the "tile program" format executed by the sm_100a tile kernel.
//
// One PASS = one sweep over the (shard of the) amplitude vector in HBM.  Every
// CTA stages a TILE of 2^T amplitudes in shared memory: T physical index bits
// (the low L bits, so every HBM access is a run of 2^L*16 contiguous bytes, plus
// arbitrary higher bits) vary inside the tile, the remaining bits select the
// tile.  Inside the tile a pass runs ROUNDS: in a round each thread keeps a
// GROUP of 2^m (m<=3) amplitudes in registers, the m "register bits" being
// tile-local bit positions, and applies every OP of the round to them before
// the group goes back to shared memory.  So one HBM pass can apply many gates,
// and one shared-memory pass applies several of them.
//
// Everything in here is physical: the host scheduler (qv_sched.cpp) has already
// translated logical qubits into physical index bits.  All structs are PODs
// copied verbatim to the device.
#pragma once
#include <stdint.h>

#define QV_MAX_TILE_BITS 12      // 2^12 amplitudes * 16 B = 64 KiB of shared memory per CTA
#define QV_MIN_LOW_BITS 4        // low bits always inside the tile: 256-byte HBM runs at worst
#define QV_REG_BITS 3            // 2^3 amplitudes per thread per round
#define QV_MAX_SEGS 12
#define QV_CHUNK_SEGS 8
#define QV_MAX_CHUNK_BITS 8      // diagonal factor tables have <= 256 entries
#define QV_THREADS 256
#define QV_MAX_PEERS 8
#define QV_MAX_PASS_CHUNKS 256   // per-tile chunk offsets are staged in shared memory
#define QV_SLICE_ENTRIES 544     // per-tile diagonal slices staged in shared memory (8.5 KiB; 3 CTAs/SM still fit)
// The control part of a pass (header, rounds, ops, chunk descriptors, matrices) is handed to the
// kernel as a __grid_constant__ parameter: it lives in the constant bank, so ptxas reads matrices
// through uniform registers instead of spending vector registers on them.  Two size classes.
#define QV_PROG_SMALL_BYTES 3584
#define QV_PROG_LARGE_BYTES 28672

enum QvOpType : uint32_t {
    QV_OP_DENSE1 = 1,   // 2x2 complex matrix on register bit rb0
    QV_OP_DENSE2 = 2,   // 4x4 complex matrix on register bits rb0 < rb1 (matrix bit0 <-> rb0)
    QV_OP_DIAG = 3,     // product of chunk-table lookups (merged diagonal gates)
};

enum QvOpFlags : uint32_t {
    QV_F_CTRL_LOCAL = 1u,   // cm_local/cv_local restrict the op to matching tile-local indices
    QV_F_CTRL_EXT = 2u,     // cm_ext/cv_ext restrict the op to matching tiles (CTA-uniform)
    QV_F_REAL = 4u,         // matrix has no imaginary parts (H, X, RY, CNOT, SWAP ...)
};

// gather/deposit of one contiguous bit field: ((x >> src) & ((1<<len)-1)) << dst
struct QvSeg {
    uint8_t src, len, dst, pad;
};

// One source table of a SLICE chunk: index = (gather_ext(base) << nl) | local_index.
struct QvSource {
    uint32_t table_off;             // offset in complex entries into the pass's table pool
    uint8_t n_esegs, pad[3];
    QvSeg esegs[QV_CHUNK_SEGS];
};

// One factor of a merged diagonal.
//   kind 0 (GLOBAL): phase = table[gather_local(e) | gather_ext(base)], table in global memory.
//   kind 1 (SLICE) : all source tables over the same tile-local bits are multiplied together ONCE PER
//                    TILE (their external bits are constant there) into a 2^nl-entry slice in shared
//                    memory; phase = slice[gather_local(e)].
struct QvChunk {
    uint32_t table_off;             // offset in complex entries into the pass's table pool
    uint8_t n_lsegs, n_esegs;       // fields gathered from the tile-local index / the tile base
    uint8_t reg_mask;               // which register bits of the op's round feed this chunk
    uint8_t gate_rb;                // 1 + register bit r such that every entry with that bit clear is exactly 1; 0 = none
    QvSeg lsegs[QV_CHUNK_SEGS];     // only the NON-register local bits (register bits go through slot_off)
    QvSeg esegs[QV_CHUNK_SEGS];
    uint32_t slot_off[8];           // table-index contribution of register slot r (host-precomputed)
    uint16_t kind;                  // 0 = GLOBAL, 1 = SLICE (table_off then counts entries into the slice area)
    uint16_t nl;                    // SLICE: log2(entries)
    uint16_t first_src, n_src;      // SLICE: its sources
};

struct QvOp {
    uint32_t type;
    uint32_t flags;
    uint8_t rb0, rb1, pad0, pad1;
    uint32_t cm_local, cv_local;    // control over the tile-local index e
    uint64_t cm_ext, cv_ext;        // control over the physical index bits outside the tile
    uint32_t data_off;              // DENSE: offset in complex entries into the matrix pool (row-major)
                                    // DIAG : index of the first chunk in the chunk array
    uint32_t n_chunks;
    uint32_t pad2[2];
};

struct QvRound {
    uint32_t m;                     // register bits in this round (<= QV_REG_BITS, <= T)
    uint32_t regpos[QV_REG_BITS];   // tile-local bit positions, ascending
    uint32_t first_op, n_ops;
    uint32_t pad[2];
    uint32_t slot_dep[8];           // tile-local index offset of register slot r
    uint32_t slot_xor[8];           // qv_swz(slot_dep[r]): the swizzle is XOR-linear, so the shared-memory
                                    // slot of (e0 | dep) is qv_swz(e0) ^ slot_xor[r]
};

struct QvPassHeader {
    uint32_t T;                     // tile bits
    uint32_t n_tile_segs;           // tile-local index e -> physical index bits
    QvSeg tile_segs[QV_MAX_SEGS];
    uint32_t n_base_segs;           // tile id -> physical index bits
    QvSeg base_segs[QV_MAX_SEGS];
    uint64_t fixed_bits;            // OR'd into every physical index of the pass (rank bits, group split)
    uint64_t n_tiles;               // tiles this device processes
    uint32_t n_local_bits;          // log2(amplitudes per shard): physical bits above it select the peer
    uint32_t n_rounds;
    uint32_t n_ops;
    uint32_t n_chunks;
    // byte offsets from the start of the control blob (the diagonal tables travel separately,
    // in global memory)
    uint32_t off_rounds, off_ops, off_chunks, off_sources, off_matrices, n_table_entries, n_sources, n_slice_entries;
    uint32_t blob_bytes;
    uint32_t uses_peers;            // tile bits include a physical bit >= n_local_bits
    uint32_t pad;
    uint64_t hi_off[16];            // physical-index bits of tile-local index 256*i (host-precomputed gather)
};

// Pull remap (multi-GPU): every rank gathers the amplitudes it will own AFTER the physical bit swaps
// (local_bit[i] <-> global_bit[i]) from the current buffers of all ranks into its alternate buffer; then all
// ranks flip buffers.  Each amplitude crosses NVLink exactly once (an in-place exchange through the tile
// kernel moves it twice: pulled by the rank that handles the tile and pushed back).
struct QvRemap {
    uint32_t n_pairs;
    uint32_t n_local_bits;
    uint32_t rank;
    uint32_t pad;
    uint32_t local_bit[8];
    uint32_t global_bit[8];
};

// A k>=3 dense gate runs as its own pass through the generic kernel.
struct QvBigGate {
    uint32_t k;
    uint32_t pad;
    uint32_t pos[16];               // physical bit of matrix index bit j
    uint64_t ctrl_mask, ctrl_val;   // physical control bits (identity when not matching)
    uint64_t fixed_bits;            // rank bits of this shard (for controls on global qubits)
};
