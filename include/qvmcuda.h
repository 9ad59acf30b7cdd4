/* qvmcuda.h -- C ABI of libqvmcuda, the B200 (sm_100a) engine behind the QVM's
 * gate-application / probability / measurement / sampling hot path.
 *
 * The reference (quil-lang/qvm, Common Lisp) has no C plugin interface for this
 * path; the seam is a set of CLOS protocols.  Each entry point
 * below is what a CFFI binding for one of those protocol methods calls; the
 * reference interface it stands behind is cited as file:line (paths relative to
 * the reference checkout).  The Lisp-side bindings are in lisp/ and
 * INTEGRATION.md.
 *
 * Conventions
 *   - every function returns 0 on success, nonzero on failure;
 *     qvmcuda_last_error() gives the message for the calling thread (the
 *     reference's FFI style: C int status -> Lisp ERROR, src/shm.lisp:198-206).
 *   - amplitudes are interleaved (re, im) doubles = the reference's CFLONUM
 *     (src/floats.lisp:13-26); a state is a flat vector of 2^n of them
 *     (src/linear-algebra.lisp:9-11); operators are row-major 2^k x 2^k
 *     (src/linear-algebra.lisp:21-24).
 *   - qubit lists are in NAT-TUPLE order (src/utilities.lisp:43-51): entry j is
 *     the qubit attached to bit j of the matrix index, i.e. the Quil argument
 *     list REVERSED.
 *   - the library never draws random numbers: uniforms come from the host's
 *     mt19937 (src/measurement.lisp:99, :136, :257).
 *   - handles are thread-safe (one mutex per handle); destroy may be called
 *     from any thread (tg:finalize runs finalizers on arbitrary threads,
 *     src/state-representation.lisp:92-102).
 *   - there is no CPU fallback: without a CUDA device every call fails.
 */
#ifndef QVMCUDA_H
#define QVMCUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qvmcuda_state qvmcuda_state;   /* a device-resident amplitude vector */
typedef struct qvmcuda_tape qvmcuda_tape;     /* a compiled (fused) gate sequence */

/* flags for qvmcuda_tape_compile / qvmcuda_apply_gates */
#define QVMCUDA_FUSE          1u   /* pack runs of gates into shared HBM passes (*fuse-gates-during-compilation*, src/config.lisp:41-58) */
#define QVMCUDA_ABSORB_SWAPS  2u   /* exact SWAP gates relabel qubits instead of moving data */

const char *qvmcuda_last_error(void);
int qvmcuda_device_count(int *count);
/* number of kernels this process has launched through the library (bench.py's gpu_launches) */
int qvmcuda_launch_count(uint64_t *count);

/* ---- allocation protocol: ALLOCATE-VECTOR / finalizer (src/allocator.lisp:47-62),
 *      MAKE-PURE-STATE / MAKE-DENSITY-MATRIX-STATE (src/state-representation.lisp:76-102, 235-266).
 *      The vector is zero-initialised, as the protocol requires. */
int qvmcuda_state_create(uint64_t n_amplitudes, int device, qvmcuda_state **out);
int qvmcuda_state_destroy(qvmcuda_state *s);                 /* idempotent on NULL */
int qvmcuda_state_length(qvmcuda_state *s, uint64_t *n_amplitudes);
/* run the handle's work on an existing CUDA stream (cudaStream_t passed as an integer) */
int qvmcuda_state_set_stream(qvmcuda_state *s, uint64_t cuda_stream);
int qvmcuda_synchronize(qvmcuda_state *s);

/* ---- state protocol: STATE-ELEMENTS / (SETF STATE-ELEMENTS), QVM::AMPLITUDES
 *      (src/state-representation.lisp:126-135, src/qvm.lisp:63-69): the device<->host sync points. */
int qvmcuda_download(qvmcuda_state *s, double *dst, uint64_t offset, uint64_t count);
int qvmcuda_upload(qvmcuda_state *s, const double *src, uint64_t offset, uint64_t count);
/* SET-TO-ZERO-STATE / BRING-TO-ZERO-STATE (src/wavefunction.lisp:80-91) */
int qvmcuda_set_zero_state(qvmcuda_state *s);
/* COPY-WAVEFUNCTION (src/wavefunction.lisp:93-109): copies min(|dst|,|src|) amplitudes */
int qvmcuda_copy(qvmcuda_state *dst, qvmcuda_state *src);

/* ---- operator API: APPLY-MATRIX-OPERATOR (src/wavefunction.lisp:308-330) behind
 *      APPLY-GATE-TO-STATE on a pure state (src/apply-gate.lisp:106-160). */
int qvmcuda_apply_matrix(qvmcuda_state *s, int k, const int32_t *qubits, const double *matrix);
/* a run of gates in one call (one CFFI crossing per program, SURVEY.md section 7 "per-transition host
 * overhead"): ks[g] qubits and 2*4^ks[g] doubles per gate, concatenated. */
int qvmcuda_apply_gates(qvmcuda_state *s, int n_gates, const int32_t *ks, const int32_t *qubits,
                        const double *matrices, uint32_t flags);

/* ---- compile protocol: COMPILE-LOADED-PROGRAM (src/qvm.lisp:166-175, src/compile-gate.lisp:484-526):
 *      build the fused tape once, run it many times. n_qubits = log2(state length). */
int qvmcuda_tape_compile(int n_qubits, int n_gates, const int32_t *ks, const int32_t *qubits,
                         const double *matrices, uint32_t flags, qvmcuda_tape **out);
int qvmcuda_tape_run(qvmcuda_state *s, qvmcuda_tape *t);
/* info[0]=HBM passes, [1]=gates, [2]=atoms (controlled blocks/diagonals), [3]=tile passes, [4]=generic k>=3 passes */
int qvmcuda_tape_info(qvmcuda_tape *t, int64_t info[8]);
int qvmcuda_tape_describe(qvmcuda_tape *t, char *buf, uint64_t buflen);
int qvmcuda_tape_destroy(qvmcuda_tape *t);
/* The pass compiler (the device-side counterpart of the reference's run-time COMPILE of gate lambdas and their cache,
 * src/compile-gate.lisp:156-209, 315-361): every fused pass of a tape that carries enough work is turned into its own
 * straight-line sm_100a kernel (NVRTC), cached in memory and on disk by structure; gate angles and qubit positions stay
 * data.  QVMCUDA_JIT = sync (default) | async | off; passes without a compiled kernel run through the interpreter kernel.
 * stats: [0] kernels compiled, [1] in-memory cache hits, [2] disk cache hits, [3] failures, [4] compiled-pass launches,
 * [5] compile time in ms, [6] policy (0 off, 1 sync, 2 async), [7] micro-op threshold. */
int qvmcuda_jit_stats(int64_t stats[8]);
/* generated CUDA source of one step (diagnostics, tests) and its cache key */
int qvmcuda_tape_jit_source(qvmcuda_tape *t, int step, char *buf, uint64_t buflen, uint64_t *signature);
/* compile every eligible step of the tape into the on-disk kernel cache without touching a device (build-time
 * prewarming; needs no GPU).  log receives the compiler output (register counts, spills). */
int qvmcuda_tape_jit_precompile(qvmcuda_tape *t, int *n_eligible, int *n_compiled, char *log, uint64_t loglen);

/* ---- measurement protocol (src/measurement.lisp:7-150) */
/* WAVEFUNCTION-EXCITED-STATE-PROBABILITY src/wavefunction.lisp:64-70 */
int qvmcuda_prob_excited(qvmcuda_state *s, int qubit, double *p);
/* WAVEFUNCTION-GROUND-STATE-PROBABILITY src/wavefunction.lisp:54-60 (compiled MEASURE, src/compile-gate.lisp:231) */
int qvmcuda_prob_ground(qvmcuda_state *s, int qubit, double *p);
/* NORM / NORMALIZE-WAVEFUNCTION src/wavefunction.lisp:333-364 */
int qvmcuda_norm2(qvmcuda_state *s, double *sum_of_squares);
/* <a|b> = sum conj(a_i) b_i: the INNER-PRODUCT of PURE-STATE-EXPECTATION (app/src/api/expectation.lisp:78-91),
 * out[0] = real part, out[1] = imaginary part.  Both states on the same device, same length. */
int qvmcuda_inner_product(qvmcuda_state *a, qvmcuda_state *b, double out[2]);
/* probabilities of COUNT basis states starting at OFFSET (|psi_i|^2, PROBABILITY src/wavefunction.lisp:44-50) straight to a
 * host buffer: the :probabilities request of the app (app/src/api/probabilities.lisp; handle-request.lisp:155-176 streams
 * them as big-endian doubles).  Half the bytes of a wavefunction download. */
int qvmcuda_probabilities(qvmcuda_state *s, double *out, uint64_t offset, uint64_t count);
int qvmcuda_scale(qvmcuda_state *s, double factor);
int qvmcuda_normalize(qvmcuda_state *s);
/* FORCE-MEASUREMENT (pure-state) src/measurement.lisp:10-41: amplitudes whose QUBIT bit differs from
 * keep_bit become 0, the others are multiplied by inv_norm. */
int qvmcuda_collapse(qvmcuda_state *s, int qubit, int keep_bit, double inv_norm);
/* multi-shot sampling.  strict=0: smallest index whose inclusive CDF is >= u
 * (SAMPLE-WAVEFUNCTION-MULTIPLE-TIMES src/measurement.lisp:246-288);
 * strict=1: smallest index whose inclusive CDF is > u
 * (SAMPLE-WAVEFUNCTION-AS-DISTRIBUTION-IN-PARALLEL-TRULY src/measurement.lisp:179-225, used by MEASURE-ALL :128-143).
 * The CDF is accumulated in the blocked order documented in oracle/qvm_oracle.c (orc_sample_tree). */
int qvmcuda_sample(qvmcuda_state *s, const double *uniforms, uint64_t n_shots, uint64_t *out, int strict);
/* The same sampler on ONE SHARD of a sharded state (physical index order, rank-major): every rank asks for its shard's mass
 * (qvmcuda_sample_total, tree order), the host adds the totals left to right, and the rank that owns a draw resolves it with
 * the preceding shards' mass as the starting accumulator BASE -- indices are bit-exact against the oracle's
 * orc_sample_tree_sharded.  Indices returned are local to the shard. */
int qvmcuda_sample_total(qvmcuda_state *s, double *total);
int qvmcuda_sample_shard(qvmcuda_state *s, const double *uniforms, uint64_t n_shots, uint64_t *out, int strict, double base);
/* psi <- |basis> (second half of MEASURE-ALL-STATE src/measurement.lisp:137-141) */
int qvmcuda_set_basis_state(qvmcuda_state *s, uint64_t basis);

/* ---- density-matrix state: vec(rho) row-major on 2n index bits (src/state-representation.lisp:235-286) */
/* superoperator application src/apply-gate.lisp:42-99,196-212: rho <- sum_j K_j rho K_j^dagger for the m
 * Kraus operators (each 2^k x 2^k row-major); m == 1 is a plain gate (single-kraus). Applied as ONE
 * operator sum_j K_j (x) conj(K_j) on the 2k bits (qubits, qubits+n). */
int qvmcuda_density_apply_kraus(qvmcuda_state *s, int n_qubits, int k, const int32_t *qubits, int m,
                                const double *kraus, uint32_t flags);
/* A RUN of such operators in one call (one CFFI crossing per stretch of gate transitions of the DENSITY-QVM,
 * src/density-qvm.lisp:125-137): operator i acts on ks[i] qubits (nat-tuple order, concatenated in QUBITS) with ms[i] Kraus
 * matrices (2*4^ks[i] doubles each, concatenated in KRAUS).  The whole run is scheduled together: U, U* and the channel
 * that follows them on the same qubit become one 4x4 on the (column, row) bit pair, and operators on different qubits share
 * HBM passes (QVMCUDA_FUSE). */
int qvmcuda_density_apply_ops(qvmcuda_state *s, int n_qubits, int n_ops, const int32_t *ks, const int32_t *qubits,
                              const int32_t *ms, const double *kraus, uint32_t flags);
/* GET-EXCITED-STATE-PROBABILITY (density-matrix-state) src/measurement.lisp:77-85 */
int qvmcuda_density_prob_excited(qvmcuda_state *s, int n_qubits, int qubit, double *p);
/* FORCE-MEASUREMENT (density-matrix-state) src/measurement.lisp:43-68 */
int qvmcuda_density_collapse(qvmcuda_state *s, int n_qubits, int qubit, int keep_bit, double inv_norm);
/* APPLY-MEASURE-DISCARD-TO-STATE (density) src/measurement.lisp:111-120 */
int qvmcuda_density_measure_discard(qvmcuda_state *s, int n_qubits, int qubit);
/* DENSITY-MATRIX-STATE-MEASUREMENT-PROBABILITIES src/state-representation.lisp:268-286 (2^n doubles to host) */
int qvmcuda_density_diag_probs(qvmcuda_state *s, int n_qubits, double *out);

/* MIXED-STATE-EXPECTATION (app/src/api/expectation.lisp:91-107): tr(Q rho) for the Hermitian matrix Q of an operator program
 * (row-major 2^n x 2^n complex doubles on the host); out = (re, im).  One read of vec(rho) and of Q. */
int qvmcuda_density_expectation(qvmcuda_state *s, int n_qubits, const double *op_matrix, double out[2]);
/* SET-TO-ZERO-STATE of UNITARY-STATE (src/unitary-qvm.lisp:57-61): the 4^n vector becomes vec(identity).  The UNITARY-QVM is
 * the pure-state path on 2n index bits (gates act on the low n bits = the row index of the column-major matrix,
 * src/unitary-qvm.lisp:113-137): every other entry point applies unchanged. */
int qvmcuda_set_identity_matrix(qvmcuda_state *s, int n_qubits);

/* ---- multi-GPU sharding (one process per GPU; layout after dqvm/src/global-addresses.lisp:99-151:
 *      the top log2(P) physical index bits select the rank).  The 64-byte IPC handle of every rank's
 *      shard is exchanged by the host (torch.distributed) and attached here so that the tile kernel can
 *      read/write peer shards over NVLink. */
int qvmcuda_shard_export(qvmcuda_state *s, uint8_t handle[64]);
int qvmcuda_shard_attach(qvmcuda_state *s, int rank, int world, const uint8_t *handles /* world*64 */);
/* Optional second shard buffer: with it, remaps are out-of-place PULLS (every rank gathers the amplitudes it
 * will own from all ranks' current buffers, then all ranks flip buffers), which moves every amplitude over
 * NVLink once instead of twice.  Costs 2x shard memory; without it remaps run in place through the tile kernel.
 * Protocol: every rank calls export_alt, the host all-gathers the handles, every rank calls attach_alt. */
int qvmcuda_shard_export_alt(qvmcuda_state *s, uint8_t handle[64]);
int qvmcuda_shard_attach_alt(qvmcuda_state *s, const uint8_t *handles /* world*64 */);
/* Collective reset of a sharded state (BRING-TO-ZERO-STATE, src/wavefunction.lisp:80-91, on shards): the rank that owns the
 * non-zero amplitude calls qvmcuda_set_basis_state, every other rank qvmcuda_shard_clear (all zeros; with an alternate buffer
 * attached the zeros are not even written until something reads them), and EVERY rank then tells the library which ranks hold
 * only zeros (bit r of mask).  Until the first exchange step, pull passes do not fetch those ranks' amplitudes over NVLink
 * and local passes on an all-zero shard are skipped.  The host must reset the mask (0) on every rank when any rank's content
 * changes by other means (an upload); gate runs keep it valid by construction. */
int qvmcuda_shard_clear(qvmcuda_state *s);
int qvmcuda_shard_set_zero_ranks(qvmcuda_state *s, uint32_t mask);
/* The same attachment for ONE host process that drives several devices (a single Lisp image with N GPUs): states[r] becomes
 * rank r of world; peers are reached through CUDA peer access, no IPC handles.  want_alt != 0 also gives every shard the
 * alternate buffer for pull remaps (all shards or none: without the memory every shard falls back to in-place exchanges).
 * The host then runs every step on every shard and synchronises all of them around steps flagged QVMCUDA_STEP_PEER. */
int qvmcuda_shard_attach_local(qvmcuda_state *const *states, int world, int want_alt);
/* Schedule a gate run against the shard's current qubit layout (gate qubits are LOGICAL, 0 <= q <
 * log2(shard length) + log2(world); every rank must pass the same gate list).  The tape is run step by
 * step: a step whose flags have QVMCUDA_STEP_PEER set reads/writes peer shards over NVLink, so the host
 * must synchronise and barrier all ranks before and after it (dqvm's MPI_Waitall + barrier,
 * dqvm/src/apply-distributed-gate.lisp:24-86); other steps touch only the local shard.  Steps are
 * asynchronous on the handle's stream.  qvmcuda_tape_commit adopts the layout the tape leaves behind. */
#define QVMCUDA_STEP_PEER  1u
#define QVMCUDA_STEP_REMAP 2u
int qvmcuda_shard_compile(qvmcuda_state *s, int n_gates, const int32_t *ks, const int32_t *qubits,
                          const double *matrices, uint32_t flags, qvmcuda_tape **out);
/* The same schedule WITHOUT a device: what rank RANK of WORLD ranks would run for this gate list from the layout l2p
 * (n_total = log2(shard length) + log2(world) entries; overwritten with the layout the tape leaves behind).  remap_pull
 * says whether the ranks will have the alternate shard buffer.  For planning and for filling the pass-compiler's kernel
 * cache ahead of time (qvmcuda_tape_jit_precompile); such a tape cannot be run. */
int qvmcuda_shard_plan(int n_total, int world, int rank, int remap_pull, int32_t *l2p, int n_gates, const int32_t *ks,
                       const int32_t *qubits, const double *matrices, uint32_t flags, qvmcuda_tape **out);
int qvmcuda_tape_num_steps(qvmcuda_tape *t, int *n_steps);
int qvmcuda_tape_step_flags(qvmcuda_tape *t, int step, uint32_t *flags);
/* info[0] = step flags, [1] = kind (0 tile pass, 1 generic dense gate, 2 stand-alone pull remap), [2] = atoms in the step,
 * [3] = (global, local) qubit pairs exchanged by a pull pass (each rank then reads (1 - 2^-pairs) of its shard over NVLink),
 *        or rank bits inside the tile of an in-place peer pass, [4] = micro-ops, [5] = rounds, [6] = 1 for an in-place peer pass */
int qvmcuda_tape_step_info(qvmcuda_tape *t, int step, int64_t info[8]);
int qvmcuda_tape_run_step(qvmcuda_state *s, qvmcuda_tape *t, int step);
int qvmcuda_tape_commit(qvmcuda_state *s, qvmcuda_tape *t);
/* current logical -> physical qubit map (n entries; physical bits >= log2(shard length) select the rank) */
int qvmcuda_state_layout(qvmcuda_state *s, int32_t *l2p, int n);

#ifdef __cplusplus
}
#endif
#endif /* QVMCUDA_H */
