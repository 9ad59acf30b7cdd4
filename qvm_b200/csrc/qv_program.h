// qv_program.h -- the "tile program" format executed by the sm_100a tile kernel.
//
// One PASS = one sweep over the (shard of the) amplitude vector in HBM.  Every
// CTA stages a TILE of 2^T amplitudes in shared memory: T physical index bits
// (the low L bits, so every HBM access is a run of 2^L*16 contiguous bytes, plus
// arbitrary higher bits) vary inside the tile, the remaining bits select the
// tile.  Inside the tile a pass runs ROUNDS: in a round each thread keeps a
// GROUP of 2^m (m <= 4) amplitudes in registers, the m "register bits" being
// tile-local bit positions, and applies every MICRO-OP of the round to them
// before the group goes back to shared memory.  So one HBM pass can apply many
// gates, and one shared-memory pass applies several of them.
//
// A micro-op is fully resolved by the host: its `kind` selects one specialised
// code path (register bits, real/complex, gating are part of the kind), table
// addresses, index fields and slot offsets are precomputed, so that the device
// spends its instructions on FP64 math, not on decoding.
//
// Everything in here is physical: the host scheduler (qv_sched.cpp) has already
// translated logical qubits into physical index bits.  All structs are PODs
// copied verbatim to the device.
#pragma once
#include <stdint.h>

#define QV_MAX_TILE_BITS 12      // 2^12 amplitudes * 16 B = 64 KiB of shared memory per CTA
#define QV_MIN_LOW_BITS 4        // low bits always inside the tile: 256-byte HBM runs at worst
#define QV_MAX_REG_BITS 4        // up to 2^4 amplitudes per thread per round
#define QV_MAX_SLOTS 16
#define QV_MAX_SEGS 12
#define QV_CHUNK_SEGS 8
#define QV_MAX_CHUNK_BITS 8      // diagonal factor tables have <= 256 entries
#define QV_MAX_SOURCE_BITS 10    // source tables of a slice have <= 1024 entries
#define QV_THREADS 256           // block size of the streaming kernels and of the 3-register-bit tile kernel
#define QV_THREADS_WIDE 128      // block size of the 4-register-bit tile kernel (fewer, fatter threads)
#define QV_MAX_PEERS 8
#define QV_MAX_EXT 128           // per-tile external index parts staged in shared memory
#define QV_MAX_SOURCES 192       // per-tile source offsets staged in shared memory
#define QV_MAX_PREDS 64          // per-tile control predicates (controls on bits outside the tile)
#define QV_MAX_SLICES 64
#define QV_SLICE_ENTRIES 544     // per-tile diagonal slices staged in shared memory (8.5 KiB; 3 CTAs/SM still fit)
#define QV_MAX_SLICE_BUILD 6144  // bound on sum(2^nl * n_src) per tile (slice construction work)
// The control part of a pass (header, rounds, micro-ops, descriptors, matrices) is handed to the
// kernel as a __grid_constant__ parameter: it lives in the constant bank, so ptxas reads matrices
// through uniform registers instead of spending vector registers on them.  Two size classes.
#define QV_PROG_SMALL_BYTES 3584
#define QV_PROG_LARGE_BYTES 28672

// Micro-op kinds.  RB = register bit index (0..3), PAIR = index of the register-bit pair
// (0,1) (0,2) (1,2) (0,3) (1,3) (2,3).
enum QvUopKind : uint32_t {
    QV_K_DENSE1 = 0,         // + 2*RB + (complex ? 1 : 0)            2x2 matrix on register bit RB
    QV_K_DENSE2 = 8,         // + 2*PAIR + (complex ? 1 : 0)          4x4 matrix on a register-bit pair
    QV_K_DIAG_COMMON = 20,   // no register bit in the table index: one factor for the whole group
    QV_K_DIAG_GATED1 = 21,   // + RB: gated by RB, no other register bit in the index: one factor for the gated slots
    QV_K_DIAG_GATEDN = 25,   // + RB: gated by RB, other register bits in the index: one lookup per gated slot
    QV_K_DIAG_ONEBIT = 29,   // + RB: exactly one register bit in the index, not gating: two lookups
    QV_K_DIAG_ALL = 33,      // one lookup per slot
    QV_K_COUNT = 34,
};

enum QvUopFlags : uint32_t {
    QV_UF_CTRL = 1u,      // dense: restricted by cm/cv (group level) and slot_ok (slot level)
    QV_UF_PRED = 2u,      // dense: restricted to tiles whose predicate `pred` holds (controls outside the tile)
    QV_UF_SLICE = 4u,     // diag: table is a per-tile slice in shared memory (else a global-memory table)
    QV_UF_EXT = 8u,       // diag: add the per-tile external index part `ext`
    QV_UF_FIELD2 = 16u,   // diag: a second index field
    QV_UF_GENERIC = 32u,  // diag: index fields come from a segment list (more than two fields)
    QV_UF_SCALE = 64u,    // diag (single-lookup kinds): multiply the entry by the one-entry slice `scale`
};

// gather/deposit of one contiguous bit field: ((x >> src) & ((1<<len)-1)) << dst
struct QvSeg {
    uint8_t src, len, dst, pad;
};

// A micro-op.  The group counter g (the tile-local index with the register bits squeezed out)
// addresses everything that depends on the non-register bits.
struct QvUop {
    uint8_t kind;                   // QvUopKind
    uint8_t flags;
    uint8_t pred;                   // index of the per-tile predicate (QV_UF_PRED)
    uint8_t ext;                    // index of the per-tile external index part (QV_UF_EXT)
    uint32_t data;                  // dense: byte offset of the row-major matrix in the blob
                                    // diag : entry offset of the table (slice area or global table pool)
    uint32_t cm, cv;                // dense: control mask / value over g
                                    // diag : cm = field 0, cv = field 1; field = shift | (mask << 8), value = (g >> shift) & mask
    uint16_t slot_ok;               // dense: slots allowed by controls on register bits
    uint16_t segs;                  // diag (QV_UF_GENERIC): byte offset of a QvSegList in the blob
    uint8_t slot_off[QV_MAX_SLOTS]; // diag: table index contribution of register slot r
    uint16_t scale;                 // diag (QV_UF_SCALE): entry of the slice area holding the per-tile scalar
    uint16_t pad16;
    uint32_t pad[2];
};                                  // 48 bytes

struct QvSegList {
    uint32_t n;
    QvSeg segs[QV_CHUNK_SEGS];
};

// Per-tile external part of a global table index: gather(tile base) << shift.
struct QvExt {
    uint8_t n_esegs, shift, pad[2];
    QvSeg esegs[QV_CHUNK_SEGS];
};

// One source table of a slice: entry x of the slice takes the factor
// tables[table_off + ((gather_ext(base) << nl) | gather_local(x))].
struct QvSource {
    uint32_t table_off;             // offset in complex entries into the pass's table pool
    uint8_t n_esegs, n_lsegs, nl, pad;
    QvSeg esegs[QV_CHUNK_SEGS];     // from the tile base
    QvSeg lsegs[QV_CHUNK_SEGS];     // from the slice index x
};

// A per-tile slice: all diagonal factors over the same tile-local bits, with their external bits
// frozen to the tile's value, multiplied together ONCE PER TILE into 2^nl shared-memory entries.
struct QvSlice {
    uint16_t off;                   // first entry in the slice area
    uint16_t nl;                    // log2(entries)
    uint16_t first_src, n_src;
};

struct QvPred {
    uint64_t mask, val;             // predicate: (tile base & mask) == val
};

struct QvRound {
    uint32_t m;                     // register bits in this round (<= reg_bits of the pass, <= T)
    uint32_t regpos[QV_MAX_REG_BITS];   // tile-local bit positions, ascending
    uint32_t first_uop, n_uops;
    uint32_t pad;
    uint16_t slot_xor[QV_MAX_SLOTS];    // qv_swz(tile-local offset of slot r): the swizzle is XOR-linear, so the
                                        // shared-memory slot of (e0 | dep) is qv_swz(e0) ^ slot_xor[r]
};                                  // 64 bytes

struct QvPassHeader {
    uint32_t T;                     // tile bits
    uint32_t reg_bits;              // 3: 256-thread kernel, 8 amplitudes per thread; 4: 128-thread kernel, 16 per thread
    uint32_t threads_log2;          // log2(block size) the gather split below was computed for
    uint32_t n_tile_segs;           // tile-local index e -> physical index bits
    QvSeg tile_segs[QV_MAX_SEGS];
    uint32_t n_base_segs;           // tile id -> physical index bits
    QvSeg base_segs[QV_MAX_SEGS];
    uint64_t fixed_bits;            // OR'd into every physical index of the pass (rank bits, group split)
    uint64_t n_tiles;               // tiles this device processes
    uint32_t n_local_bits;          // log2(amplitudes per shard): physical bits above it select the peer
    uint32_t n_rounds, n_uops, n_ext, n_sources, n_slices, n_slice_entries, n_preds;
    // byte offsets from the start of the control blob (the diagonal tables travel separately,
    // in global memory)
    uint32_t off_rounds, off_uops, off_ext, off_sources, off_slices, off_slice_of, off_preds, off_matrices;
    uint32_t n_table_entries;
    uint32_t blob_bytes;
    uint32_t uses_peers;            // tile bits include a physical bit >= n_local_bits
    uint32_t n_diag_uops;           // statistics for describe()
    uint64_t hi_off[32];            // physical-index bits of tile-local index (block size)*i (host-precomputed gather)
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
