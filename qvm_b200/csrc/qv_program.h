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
#if defined(__CUDACC_RTC__)
// NVRTC (the pass compiler, qv_jit.cpp) has no system headers: the fixed-width types come from the compiler's own
typedef unsigned char uint8_t;
typedef unsigned short uint16_t;
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef long long int64_t;
#else
#include <stdint.h>
#endif

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
// (0,1) (0,2) (1,2) (0,3) (1,3) (2,3).  GATE = 0: every slot, 1 + RB: only the slots whose register bit RB
// is set (the table then only holds the entries with that bit set: every entry with the bit clear is
// exactly 1 -- controlled-phase ladders).
//   DIAG1: the table index has no register bit -> one lookup serves the whole group.
//   DIAGR: the index is (nonreg_index << field) + slot, the slot field being the full slot number (or the
//          slot number with the gate bit squeezed out), so per-slot offsets are compile-time constants.
//   table space: S = per-tile slice in shared memory, G = global-memory table, C = constants in the blob.
enum QvUopKind : uint32_t {
    QV_K_DENSE1 = 0,         // + 2*RB + (complex ? 1 : 0)            2x2 matrix on register bit RB
    QV_K_DENSE2 = 8,         // + 2*PAIR + (complex ? 1 : 0)          4x4 matrix on a register-bit pair
    QV_K_DIAG_BASE = 20,
    QV_K_DIAG1_S = 20,       // + GATE
    QV_K_DIAG1_G = 25,       // + GATE
    QV_K_DIAGR_S = 30,       // + GATE
    QV_K_DIAGR_G = 35,       // + GATE
    QV_K_DIAGR_C = 40,       // + GATE   (index has register bits only, no per-tile part)
    QV_K_END = 45,           // terminates the micro-op list of a round (the kernel loops on the kind alone)
    QV_K_BFLY = 46,          // + RB: the unscaled butterfly [[1,1],[1,-1]] on register bit RB (Hadamard-like gates; their
                             //       common scale factor is folded into the pass's write-back scale)
    QV_K_BFLY_DIAG1_S = 50,  // + RB: butterfly on RB, then the DIAG1 (slice table) gated by RB: one dispatch for the QFT's
    QV_K_BFLY_DIAG1_G = 54,  // + RB: "H, then the controlled phases hanging off that qubit"; same with a global table
    QV_K_COUNT = 58,
};

enum QvUopFlags : uint32_t {
    QV_UF_CTRL = 1u,      // dense: restricted by cm/cv (group level) and slot_ok (slot level)
    QV_UF_PRED = 2u,      // dense: restricted to tiles whose predicate `pred` holds (controls outside the tile)
    QV_UF_GENERIC = 4u,   // diag: index fields come from a segment list (more than two fields)
    QV_UF_SCALE = 8u,     // diag (DIAG1 kinds): multiply the entry by the one-entry slice `scale`
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
    uint8_t pad0;
    uint32_t data;                  // dense: byte offset of the row-major matrix in the blob
                                    // diag : entry offset of the table (slice area / global table pool), byte offset in the blob for C
    uint32_t cm, cv;                // dense: control mask / value over g
                                    // diag : cm = field 0, cv = field 1; field = shift | (mask << 8), value = (g >> shift) & mask
    uint16_t slot_ok;               // dense: slots allowed by controls on register bits
    uint16_t segs;                  // diag (QV_UF_GENERIC): byte offset of a QvSegList in the blob
    uint16_t scale;                 // diag (QV_UF_SCALE): entry of the slice area holding the per-tile scalar
    uint16_t pad1;
    uint32_t pad2[2];
};                                  // 32 bytes

struct QvSegList {
    uint32_t n;
    QvSeg segs[QV_CHUNK_SEGS];
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
    uint32_t n_rounds, n_uops, n_sources, n_slices, n_slice_entries, n_preds;
    // byte offsets from the start of the control blob (the diagonal tables travel separately,
    // in global memory)
    uint32_t off_rounds, off_uops, off_sources, off_slices, off_slice_of, off_preds, off_matrices;
    uint32_t n_table_entries;
    uint32_t blob_bytes;
    uint32_t uses_peers;            // tile bits include a physical bit >= n_local_bits
    uint32_t n_diag_uops;           // statistics for describe()
    uint64_t hi_off[32];            // physical-index bits of tile-local index (block size)*i (host-precomputed gather)
    uint64_t hi_byte[32];           // 16 * hi_off[i]: byte offsets for the local-pass fast path (one 64-bit add per element)
    // Store permutation: the trailing X / CNOT / SWAP gates of a pass are GF(2)-affine maps of the tile-local
    // index, so they cost no arithmetic at all: the element written to tile-local position e is read from
    // shared-memory slot qv_swz(A e ^ b) = st_lo(tid) ^ st_hi[i] for e = tid + (block size)*i.
    uint32_t store_perm;            // 0: identity (plain write-back)
    uint16_t st_col[QV_MAX_TILE_BITS];  // qv_swz(A column k): contribution of bit k of tid
    uint16_t st_const;              // qv_swz(b)
    uint16_t st_pad;
    uint16_t st_hi[32];             // qv_swz(A ((block size)*i))
    // Fused pull remap (multi-GPU): the pass READS through a pending qubit remap -- element p of this rank's
    // new shard is amplitude S(rank, p) of the ranks' CURRENT buffers, S = the index with every
    // (local_bit, global_bit) pair exchanged -- and WRITES its results into the alternate buffer; all ranks
    // flip buffers afterwards.  The exchange costs no pass of its own.  S is GF(2)-linear, so
    // S(pbase | hi_off[i]) = S(pbase) ^ hi_src[i] with hi_src host-precomputed.
    // Write-back scale: uncontrolled gates of the form s*[[1,1],[1,-1]] (H) run as unscaled butterflies (2 FP64 ops per
    // amplitude instead of 4); the product of their factors is applied once, while the tile is written back.
    double out_scale;
    uint32_t has_scale;
    uint32_t src_basis;             // patched at launch time: the state is the basis vector |basis_index> that was never written to
                                    // HBM (lazy SET-TO-ZERO-STATE): the pass synthesises its tiles instead of loading them
    uint32_t pull;                  // 0: in place
    uint32_t zero_ranks;            // pull passes, patched at launch: ranks whose CURRENT buffer is known to hold only zeros (a
                                    // sharded state right after a reset): their amplitudes are not fetched, the copy zero-fills
    QvRemap pull_remap;
    uint64_t hi_src[32];            // S(hi_off[i])
    uint64_t basis_index;           // src_basis: physical index of the one non-zero amplitude
    uint64_t tile_mask;             // physical index bits that vary inside a tile
};

struct qvc;
struct QvPeers {
    qvc* base[QV_MAX_PEERS];   // shard base pointer of every rank (own pointer at [rank])
};

// Geometry of a tile for the tensor-memory accelerator (qv_jit_kernel_tma.cuh): dimension 0 of the 5-d view = index bits 0..2
// (eight amplitudes = 16 doubles = 128 bytes); dimensions 1..4 = runs of consecutive index bits >= 3 that are all inside
// (is_tile) or all outside the tile; unused dimensions have len 0.
struct QvTmaGeom {
    uint8_t start[4], len[4], is_tile[4];
    uint32_t n_runs;
};

// A k>=3 dense gate runs as its own pass through the generic kernel.
struct QvBigGate {
    uint32_t k;
    uint32_t diag;                  // 1: the "matrix" is the 2^k diagonal of a diagonal gate too wide for the table micro-ops
    uint32_t pos[16];               // physical bit of matrix index bit j
    uint64_t ctrl_mask, ctrl_val;   // physical control bits (identity when not matching)
    uint64_t fixed_bits;            // rank bits of this shard (for controls on global qubits)
};
