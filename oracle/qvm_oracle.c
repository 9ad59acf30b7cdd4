/* oracle/qvm_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain scalar C restatement of the reference QVM's (quil-lang/qvm v1.18.0)
 * gate-application / probability / measurement / sampling hot path.  It is the
 * checker the CUDA path is held to; it is never linked, imported or executed
 * by the product (qvm_b200/, libqvmcuda).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Parity pinning: the reference itself (Common Lisp, SBCL) cannot run in the
 * build container, so this restatement is pinned against the reference's own
 * known-answer tests, transcribed into tests/golden/reference_kats.json
 * (see tests/test_oracle_golden.py for the file:line of every vector).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference checkout).  Compile with -ffp-contract=off: the reference's
 * scalar SBCL path does not fuse multiply-adds.
 *
 * Layout contract (src/floats.lisp:13-26, src/linear-algebra.lisp:9-24):
 *   amplitude = interleaved (re, im) doubles; state = flat vector of 2^n
 *   amplitudes; operator = row-major 2^k x 2^k complex matrix.
 *   Qubit tuples arrive here in NAT-TUPLE order (src/utilities.lisp:43-51):
 *   qubits[j] is the qubit attached to bit j of the matrix index, i.e. the
 *   Quil argument list reversed (first Quil argument = MSB).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double re, im; } cplx;

static inline cplx cmul(cplx a, cplx b) {
    /* complex multiply as SBCL open-codes it: (ar*br - ai*bi, ar*bi + ai*br) */
    cplx r;
    r.re = a.re * b.re - a.im * b.im;
    r.im = a.re * b.im + a.im * b.re;
    return r;
}
static inline cplx cadd(cplx a, cplx b) { cplx r = { a.re + b.re, a.im + b.im }; return r; }

/* src/utilities.lisp:243-255  INJECT-BIT: insert a 0 at bit position n. */
uint64_t orc_inject_bit(uint64_t x, int n) {
    uint64_t right = x & ((1ULL << n) - 1);
    uint64_t left = (x & ~((1ULL << n) - 1)) << 1;
    return left | right;
}

/* src/utilities.lisp:257-270 EJECT-BIT: remove the bit at position n. */
uint64_t orc_eject_bit(uint64_t x, int n) {
    uint64_t right = x & ((1ULL << n) - 1);
    uint64_t left = (x >> (n + 1)) << n;
    return left | right;
}

/* src/wavefunction.lisp:24-40 INDEX-TO-ADDRESS */
uint64_t orc_index_to_address(uint64_t index, int qubit, int state) {
    uint64_t a = orc_inject_bit(index, qubit);
    if (state) a |= (1ULL << qubit);
    return a;
}

/* src/wavefunction.lisp:44-50 PROBABILITY */
static inline double prob(cplx a) { return a.re * a.re + a.im * a.im; }

static void sort_ints(int *a, int k) {
    for (int i = 1; i < k; i++) {
        int v = a[i], j = i - 1;
        while (j >= 0 && a[j] > v) { a[j + 1] = a[j]; j--; }
        a[j + 1] = v;
    }
}

/* Generic k-qubit operator application.
 * Follows APPLY-OPERATOR (src/wavefunction.lisp:271-306): MAP-COMPLEMENT
 * (:155-175) enumerates i in [0, 2^(n-k)) and injects a 0 at every gate qubit
 * in ascending order; WITH-MODIFIED-AMPLITUDES (:234-269) gathers the 2^k
 * amplitudes with SET-QUBIT-COMPONENTS-OF-AMPLITUDE-ADDRESS (:112-122: bit j
 * of the combo goes to qubit nt[j]); MATRIX-MULTIPLY (src/linear-algebra.lisp:
 * 97-131) computes result[r] = sum_c M[r][c]*col[c], accumulated left to right
 * from 0; the result is scattered back.  The 1q/2q/3q serial kernels
 * (src/serial-kernels.lisp:54-66) compute the same sums in the same order.
 */
static void apply_range(cplx *psi, int k, const int *qubits, const int *sorted,
                        const cplx *mat, uint64_t lo, uint64_t hi, cplx *col, cplx *res) {
    uint64_t d = 1ULL << k;
    for (uint64_t i = lo; i < hi; i++) {
        uint64_t base = i;
        for (int s = 0; s < k; s++) base = orc_inject_bit(base, sorted[s]);
        for (uint64_t c = 0; c < d; c++) {
            uint64_t a = base;
            for (int j = 0; j < k; j++)
                if ((c >> j) & 1) a |= (1ULL << qubits[j]);
            col[c] = psi[a];
        }
        for (uint64_t r = 0; r < d; r++) {
            cplx e = { 0.0, 0.0 };
            for (uint64_t c = 0; c < d; c++) e = cadd(e, cmul(mat[r * d + c], col[c]));
            res[r] = e;
        }
        for (uint64_t c = 0; c < d; c++) {
            uint64_t a = base;
            for (int j = 0; j < k; j++)
                if ((c >> j) & 1) a |= (1ULL << qubits[j]);
            psi[a] = res[c];
        }
    }
}

/* APPLY-MATRIX-OPERATOR src/wavefunction.lisp:308-330 (all branches compute
 * the same sums).  n_amps = 2^n.  qubits in NAT-TUPLE order. */
int orc_apply_matrix(double *psi_, uint64_t n_amps, int k, const int *qubits, const double *mat_) {
    cplx *psi = (cplx *)psi_;
    const cplx *mat = (const cplx *)mat_;
    if (k < 0 || k > 20) return -1;
    int sorted[32];
    for (int j = 0; j < k; j++) sorted[j] = qubits[j];
    sort_ints(sorted, k);
    uint64_t d = 1ULL << k;
    cplx *col = (cplx *)malloc(sizeof(cplx) * d * 2);
    if (!col) return -2;
    apply_range(psi, k, qubits, sorted, mat, 0, n_amps >> k, col, col + d);
    free(col);
    return 0;
}

/* Multi-threaded variant used ONLY as the timed CPU baseline: the same loop,
 * split into contiguous per-thread ranges exactly like WITH-PARALLEL-SUBDIVISIONS
 * / LPARALLEL:PDOTIMES over the complement index (src/utilities.lisp:346-383,
 * src/wavefunction.lisp:177-195). */
int orc_apply_matrix_mt(double *psi_, uint64_t n_amps, int k, const int *qubits,
                        const double *mat_, int n_threads) {
    cplx *psi = (cplx *)psi_;
    const cplx *mat = (const cplx *)mat_;
    if (k < 0 || k > 20) return -1;
    int sorted[32];
    for (int j = 0; j < k; j++) sorted[j] = qubits[j];
    sort_ints(sorted, k);
    uint64_t d = 1ULL << k;
    uint64_t groups = n_amps >> k;
    if (n_threads < 1) n_threads = 1;
#ifdef _OPENMP
#pragma omp parallel num_threads(n_threads)
    {
        int t = omp_get_thread_num(), nt = omp_get_num_threads();
        uint64_t lo = groups * (uint64_t)t / (uint64_t)nt;
        uint64_t hi = groups * (uint64_t)(t + 1) / (uint64_t)nt;
        cplx *col = (cplx *)malloc(sizeof(cplx) * d * 2);
        apply_range(psi, k, qubits, sorted, mat, lo, hi, col, col + d);
        free(col);
    }
#else
    cplx *col = (cplx *)malloc(sizeof(cplx) * d * 2);
    apply_range(psi, k, qubits, sorted, mat, 0, groups, col, col + d);
    free(col);
#endif
    return 0;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Permutation gate as compiled by GENERATE-PERMUTATION-GATE-APPLICATION-CODE
 * (src/compile-gate.lisp:259-309): within each group the transpositions from
 * PERMUTATION-TO-TRANSPOSITIONS are applied with ROTATEF, which leaves
 * psi'[i] = psi[perm[i]] (tests/utilities-tests.lisp:41-57). */
int orc_apply_permutation(double *psi_, uint64_t n_amps, int k, const int *qubits, const int *perm) {
    cplx *psi = (cplx *)psi_;
    int sorted[32];
    for (int j = 0; j < k; j++) sorted[j] = qubits[j];
    sort_ints(sorted, k);
    uint64_t d = 1ULL << k;
    cplx *col = (cplx *)malloc(sizeof(cplx) * d);
    for (uint64_t i = 0; i < (n_amps >> k); i++) {
        uint64_t base = i;
        for (int s = 0; s < k; s++) base = orc_inject_bit(base, sorted[s]);
        for (uint64_t c = 0; c < d; c++) {
            uint64_t a = base;
            for (int j = 0; j < k; j++) if ((c >> j) & 1) a |= (1ULL << qubits[j]);
            col[c] = psi[a];
        }
        for (uint64_t c = 0; c < d; c++) {
            uint64_t a = base;
            for (int j = 0; j < k; j++) if ((c >> j) & 1) a |= (1ULL << qubits[j]);
            psi[a] = col[perm[c]];
        }
    }
    free(col);
    return 0;
}

/* WAVEFUNCTION-EXCITED-STATE-PROBABILITY src/wavefunction.lisp:64-70 (serial
 * PSUM-DOTIMES branch, src/utilities.lisp:400-405: one running sum). */
double orc_prob_excited(const double *psi_, uint64_t n_amps, int q) {
    const cplx *psi = (const cplx *)psi_;
    double s = 0.0;
    for (uint64_t i = 0; i < n_amps / 2; i++) s += prob(psi[orc_index_to_address(i, q, 1)]);
    return s;
}

/* WAVEFUNCTION-GROUND-STATE-PROBABILITY src/wavefunction.lisp:54-60 */
double orc_prob_ground(const double *psi_, uint64_t n_amps, int q) {
    const cplx *psi = (const cplx *)psi_;
    double s = 0.0;
    for (uint64_t i = 0; i < n_amps / 2; i++) s += prob(psi[orc_inject_bit(i, q)]);
    return s;
}

/* %SERIAL-NORM src/wavefunction.lisp:333-338 (returns the squared norm too) */
double orc_norm2(const double *psi_, uint64_t n_amps) {
    const cplx *psi = (const cplx *)psi_;
    double s = 0.0;
    for (uint64_t i = 0; i < n_amps; i++) s += prob(psi[i]);
    return s;
}
double orc_norm(const double *psi_, uint64_t n_amps) { return sqrt(orc_norm2(psi_, n_amps)); }

/* PROBABILITY src/wavefunction.lisp:44-50 over every amplitude (what PERFORM-PROBABILITIES collects,
 * app/src/api/probabilities.lisp): out[i] = re^2 + im^2. */
void orc_probabilities(const double *psi_, uint64_t n_amps, double *out) {
    const cplx *psi = (const cplx *)psi_;
    for (uint64_t i = 0; i < n_amps; i++) out[i] = prob(psi[i]);
}

/* INNER-PRODUCT inside PURE-STATE-EXPECTATION app/src/api/expectation.lisp:79-84:
 * (loop :for ai :across a :for bi :across b :sum (* (conjugate ai) bi)), sequential, out = (re, im). */
void orc_inner_product(const double *a_, const double *b_, uint64_t n_amps, double *out) {
    const cplx *a = (const cplx *)a_, *b = (const cplx *)b_;
    double re = 0.0, im = 0.0;
    for (uint64_t i = 0; i < n_amps; i++) {
        re += a[i].re * b[i].re + a[i].im * b[i].im;
        im += a[i].re * b[i].im - a[i].im * b[i].re;
    }
    out[0] = re;
    out[1] = im;
}

/* NORMALIZE-WAVEFUNCTION src/wavefunction.lisp:349-364: psi *= 1/norm.
 * (* real complex) in SBCL scales both components. */
void orc_scale(double *psi_, uint64_t n_amps, double inv) {
    cplx *psi = (cplx *)psi_;
    for (uint64_t i = 0; i < n_amps; i++) { psi[i].re = inv * psi[i].re; psi[i].im = inv * psi[i].im; }
}
void orc_normalize(double *psi_, uint64_t n_amps) { orc_scale(psi_, n_amps, 1.0 / orc_norm(psi_, n_amps)); }

/* FORCE-MEASUREMENT (pure-state) src/measurement.lisp:10-41 */
void orc_force_measurement(double *psi_, uint64_t n_amps, int q, int measured_value, double excited_probability) {
    cplx *psi = (cplx *)psi_;
    int annihilated = 1 - measured_value;
    double inv = (annihilated == 0) ? 1.0 / sqrt(excited_probability)
                                    : 1.0 / sqrt(1.0 - excited_probability);
    for (uint64_t i = 0; i < n_amps; i++) {
        if ((int)((i >> q) & 1) == annihilated) { psi[i].re = 0.0; psi[i].im = 0.0; }
        else { psi[i].re = inv * psi[i].re; psi[i].im = inv * psi[i].im; }
    }
}

/* MEASURE (base-qvm) decision rule src/measurement.lisp:93-105: given the
 * uniform draw r, returns the classical bit and collapses. */
int orc_measure(double *psi, uint64_t n_amps, int q, double r) {
    double p1 = orc_prob_excited(psi, n_amps, q);
    int cbit = (p1 == 0.0) ? 0 : (r <= p1 ? 1 : 0);
    orc_force_measurement(psi, n_amps, q, cbit, p1);
    return cbit;
}

/* Compiled MEASURE src/compile-gate.lisp:221-254: p0 = ground probability,
 * keep 0 iff r < p0, scale by 1/sqrt(p0) or 1/sqrt(1-p0), multiply every
 * amplitude by inv_norm * [bit == kept]. */
int orc_measure_compiled(double *psi_, uint64_t n_amps, int q, double r) {
    cplx *psi = (cplx *)psi_;
    double p0 = orc_prob_ground(psi_, n_amps, q);
    int keep_zero = r < p0;
    int bit = keep_zero ? 0 : 1;
    double inv = keep_zero ? 1.0 / sqrt(p0) : 1.0 / sqrt(1.0 - p0);
    for (uint64_t i = 0; i < n_amps; i++) {
        double f = inv * (double)(((int)((i >> q) & 1) == bit) ? 1 : 0);
        psi[i].re = f * psi[i].re; psi[i].im = f * psi[i].im;
    }
    return bit;
}

/* CUMULATIVE-DISTRIBUTION-FUNCTION src/wavefunction.lisp:382-395 */
void orc_cdf(const double *psi_, uint64_t n_amps, double *cdf) {
    const cplx *psi = (const cplx *)psi_;
    double s = 0.0;
    for (uint64_t i = 0; i < n_amps; i++) { s += prob(psi[i]); cdf[i] = s; }
}

/* MIDPOINT src/measurement.lisp:172-177 */
static inline uint64_t midpoint(uint64_t a, uint64_t b) { return a + (b - a) / 2; }

/* SAMPLE-WAVEFUNCTION-MULTIPLE-TIMES src/measurement.lisp:246-288, with the
 * uniform draws supplied by the caller instead of (random 1.0d0). */
void orc_sample_multiple(const double *psi, uint64_t n_amps, const double *u, uint64_t n_shots, uint64_t *out) {
    double *cdf = (double *)malloc(sizeof(double) * n_amps);
    orc_cdf(psi, n_amps, cdf);
    for (uint64_t t = 0; t < n_shots; t++) {
        double p = u[t];
        uint64_t min = 0, max = n_amps - 1, r;
        for (;;) {
            uint64_t mid = midpoint(min, max);
            if (min == max) { r = min; break; }
            if (max - min == 1) { r = (cdf[min] < p) ? max : min; break; }
            if (cdf[mid] < p) min = mid; else max = mid;
        }
        out[t] = r;
    }
    free(cdf);
}

/* SAMPLE-WAVEFUNCTION-AS-DISTRIBUTION-IN-PARALLEL-TRULY src/measurement.lisp:
 * 179-225: smallest b with C(b) > p by bisection, re-summing the lower half
 * each round (serial PSUM-DOTIMES branch). */
uint64_t orc_sample_bisect(const double *psi_, uint64_t n_amps, double p) {
    const cplx *psi = (const cplx *)psi_;
    uint64_t min = 0, max = n_amps;
    double sum_lt = 0.0;
    while (min != max - 1) {
        uint64_t mid = midpoint(min, max);
        double part = 0.0;
        for (uint64_t i = 0; i < mid - min; i++) part += prob(psi[i + min]);
        double sum = sum_lt + part;
        if (sum <= p) { min = mid; sum_lt = sum; }
        else max = mid;
    }
    return min;
}

/* SAMPLE-WAVEFUNCTION-AS-DISTRIBUTION src/measurement.lisp:292-311: sequential
 * cumsum; for every p: last idx at which (p < cumsum) first became true...
 * restated literally: for each idx, advance i while ps[i] < cumsum. */
void orc_sample_as_distribution(const double *psi_, uint64_t n_amps, const double *ps, uint64_t len, uint64_t *out) {
    const cplx *psi = (const cplx *)psi_;
    double cumsum = 0.0;
    for (uint64_t i = 0; i < len; i++) out[i] = 0;
    for (uint64_t idx = 0; idx < n_amps; idx++) {
        cumsum += prob(psi[idx]);
        uint64_t i = 0;
        while (i < len && ps[i] < cumsum) { out[i] = idx; i++; }
    }
}

/* MEASURE-ALL-STATE (pure) src/measurement.lisp:128-143: sample one basis
 * state with draw p, then psi <- |b>. Returns b. */
uint64_t orc_measure_all(double *psi_, uint64_t n_amps, double p) {
    cplx *psi = (cplx *)psi_;
    uint64_t b = orc_sample_bisect(psi_, n_amps, p);
    memset(psi, 0, sizeof(cplx) * n_amps);      /* BRING-TO-ZERO-STATE wavefunction.lisp:80-91 */
    psi[0].re = 1.0;
    cplx t = psi[0]; psi[0] = psi[b]; psi[b] = t; /* ROTATEF */
    return b;
}

/* BRING-TO-ZERO-STATE src/wavefunction.lisp:80-91 */
void orc_zero_state(double *psi_, uint64_t n_amps) {
    memset(psi_, 0, sizeof(cplx) * n_amps);
    psi_[0] = 1.0;
}

/* ---------------------------------------------------------------- density */

/* vec(rho) is row-major: index = row*2^n + col (src/state-representation.lisp:
 * 235-286).  A k-qubit unitary U on QUBITS is applied as U* on the column bits
 * (qubits) then U on the row bits (ghosts = qubits + n), both through the
 * pure-state kernel on 2n bits (src/apply-gate.lisp:56-65,196-212;
 * CONJUGATE-ENTRYWISE src/linear-algebra.lisp:247-263). */
int orc_density_apply_unitary(double *rho, int n, int k, const int *qubits, const double *U_) {
    const cplx *U = (const cplx *)U_;
    uint64_t d = 1ULL << k;
    uint64_t len = 1ULL << (2 * n);
    cplx *Uc = (cplx *)malloc(sizeof(cplx) * d * d);
    for (uint64_t i = 0; i < d * d; i++) { Uc[i].re = U[i].re; Uc[i].im = -U[i].im; }
    int ghosts[32];
    for (int j = 0; j < k; j++) ghosts[j] = qubits[j] + n;
    orc_apply_matrix(rho, len, k, qubits, (const double *)Uc);
    orc_apply_matrix(rho, len, k, ghosts, U_);
    free(Uc);
    return 0;
}

/* KRAUS-LIST branch of %EVOLVE-DENSITY-MATRIX-WITH-SUPEROPERATOR
 * src/apply-gate.lisp:67-99: pristine copy, sum <- 0, for every Kraus operator:
 * apply, sum += vec, vec <- pristine; finally vec <- sum.  m == 1 degenerates
 * to the single-Kraus branch (:74-77). */
int orc_density_apply_kraus(double *rho_, int n, int k, const int *qubits, int m, const double *kraus) {
    uint64_t d = 1ULL << k;
    uint64_t len = 1ULL << (2 * n);
    if (m == 0) return 0;
    if (m == 1) return orc_density_apply_unitary(rho_, n, k, qubits, kraus);
    cplx *rho = (cplx *)rho_;
    cplx *pristine = (cplx *)malloc(sizeof(cplx) * len);
    cplx *sum = (cplx *)calloc(len, sizeof(cplx));
    memcpy(pristine, rho, sizeof(cplx) * len);
    for (int j = 0; j < m; j++) {
        orc_density_apply_unitary(rho_, n, k, qubits, kraus + 2 * d * d * (uint64_t)j);
        for (uint64_t i = 0; i < len; i++) sum[i] = cadd(sum[i], rho[i]);
        memcpy(rho, pristine, sizeof(cplx) * len);
    }
    memcpy(rho, sum, sizeof(cplx) * len);
    free(pristine); free(sum);
    return 0;
}

/* GET-EXCITED-STATE-PROBABILITY (density-matrix-state) src/measurement.lisp:77-85 */
double orc_density_prob_excited(const double *rho_, int n, int q) {
    const cplx *rho = (const cplx *)rho_;
    uint64_t dim = 1ULL << n;
    double s = 0.0;
    for (uint64_t k = 0; k < dim / 2; k++) {
        uint64_t i = orc_index_to_address(k, q, 1);
        s += rho[i * dim + i].re;
    }
    return s;
}

/* FORCE-MEASUREMENT (density-matrix-state) src/measurement.lisp:43-68 */
void orc_density_force_measurement(double *rho_, int n, int q, int measured_value, double excited_probability) {
    cplx *rho = (cplx *)rho_;
    int annihilated = 1 - measured_value;
    double inv = (annihilated == 0) ? 1.0 / excited_probability : 1.0 / (1.0 - excited_probability);
    uint64_t len = 1ULL << (2 * n);
    for (uint64_t k = 0; k < len; k++) {
        if ((int)((k >> q) & 1) == annihilated || (int)((k >> (q + n)) & 1) == annihilated) {
            rho[k].re = 0.0; rho[k].im = 0.0;
        } else { rho[k].re = inv * rho[k].re; rho[k].im = inv * rho[k].im; }
    }
}

/* APPLY-MEASURE-DISCARD-TO-STATE (density) src/measurement.lisp:111-120:
 * zero rho[i][j] unless bit q of i equals bit q of j. */
void orc_density_measure_discard(double *rho_, int n, int q) {
    cplx *rho = (cplx *)rho_;
    uint64_t dim = 1ULL << n;
    for (uint64_t i = 0; i < dim; i++)
        for (uint64_t j = 0; j < dim; j++)
            if (((i >> q) & 1) != ((j >> q) & 1)) { rho[i * dim + j].re = 0.0; rho[i * dim + j].im = 0.0; }
}

/* DENSITY-MATRIX-STATE-MEASUREMENT-PROBABILITIES src/state-representation.lisp:268-286 */
void orc_density_diag_probs(const double *rho_, int n, double *out) {
    const cplx *rho = (const cplx *)rho_;
    uint64_t dim = 1ULL << n;
    for (uint64_t i = 0; i < dim; i++) out[i] = rho[i + i * dim].re;
}

/* ------------------------------------------------- hierarchical sampler
 * The GPU sampler (qvm_b200/csrc/sampling.cuh) cannot reproduce a 2^n-long
 * sequential running sum, so "bit-exact given identical uniforms" is defined
 * against THIS restatement of its summation order (SURVEY.md section 7,
 * "Sampler bit-exactness"); tests additionally check that it only disagrees
 * with orc_sample_multiple / orc_sample_bisect for draws within 1e-12 of a CDF
 * step.  Order (B = 1024 amplitudes per leaf block, F = 1024 fan-out):
 *   leaf sum  L1[b]  : 32 lane partials, lane l adds |psi[b*B + l + 32*j]|^2 for
 *                      j = 0..31 in order; then a 5-step xor butterfly
 *                      (offsets 16,8,4,2,1), v = v + partner.
 *   level sum L2[c]  : same reduction shape over the 1024 L1 entries of a chunk
 *                      (missing entries count as 0).
 *   top prefix       : sequential inclusive running sum over L2.
 *   descent for a draw p with tie rule `strict` (C(b) > p) or not (C(b) >= p):
 *     c  = first L2 chunk whose inclusive top prefix satisfies the rule
 *          (last chunk if none); acc = exclusive top prefix of c;
 *     b  = first leaf block in c for which acc + running(L1) satisfies the rule
 *          (running is a sequential sum started at 0 inside the chunk; last
 *          block if none); acc2 = acc + running-before-b;
 *     i  = first amplitude in b for which acc2 + running(|psi|^2) satisfies the
 *          rule (sequential, started at 0; last amplitude if none).
 */
#define SB 1024
static double tree_sum_1024(const double *v, uint64_t count) {
    double lane[32];
    for (int l = 0; l < 32; l++) {
        double s = 0.0;
        for (int j = 0; j < 32; j++) {
            uint64_t idx = (uint64_t)l + 32u * (uint64_t)j;
            s += (idx < count) ? v[idx] : 0.0;
        }
        lane[l] = s;
    }
    for (int off = 16; off >= 1; off >>= 1) {
        double nxt[32];
        for (int l = 0; l < 32; l++) nxt[l] = lane[l] + lane[l ^ off];
        memcpy(lane, nxt, sizeof(lane));
    }
    return lane[0];
}

static inline int rule_hit(double c, double p, int strict) { return strict ? (c > p) : (c >= p); }

/* The descent with a starting accumulator BASE (the probability mass that precedes this vector: 0 for a whole state,
 * the sum of the preceding shards' totals for one shard of a sharded state).  Every comparison adds BASE first, in the
 * same place the GPU does (qv_sample_descend_kernel); BASE = 0.0 leaves every sum bit-identical.  *total = the vector's
 * own mass in tree order (last entry of the top prefix). */
static void sample_tree_base(const cplx *psi, uint64_t n_amps, const double *u, uint64_t n_shots,
                             int strict, double base, uint64_t *out, double *total) {
    uint64_t n1 = (n_amps + SB - 1) / SB;
    uint64_t n2 = (n1 + SB - 1) / SB;
    double *l1 = (double *)malloc(sizeof(double) * n1);
    double *l2 = (double *)malloc(sizeof(double) * n2);
    double *top = (double *)malloc(sizeof(double) * n2);
    double tmp[SB];
    for (uint64_t b = 0; b < n1; b++) {
        uint64_t cnt = n_amps - b * SB; if (cnt > SB) cnt = SB;
        for (uint64_t i = 0; i < cnt; i++) tmp[i] = prob(psi[b * SB + i]);
        l1[b] = tree_sum_1024(tmp, cnt);
    }
    for (uint64_t c = 0; c < n2; c++) {
        uint64_t cnt = n1 - c * SB; if (cnt > SB) cnt = SB;
        l2[c] = tree_sum_1024(l1 + c * SB, cnt);
    }
    double s = 0.0;
    for (uint64_t c = 0; c < n2; c++) { s += l2[c]; top[c] = s; }
    if (total) *total = top[n2 - 1];
    for (uint64_t t = 0; t < n_shots; t++) {
        double p = u[t];
        /* chunk: binary search over the monotone top prefix */
        uint64_t lo = 0, hi = n2 - 1;
        while (lo < hi) {
            uint64_t mid = lo + (hi - lo) / 2;
            if (rule_hit(base + top[mid], p, strict)) hi = mid; else lo = mid + 1;
        }
        uint64_t c = lo;
        double acc = (c == 0) ? base : base + top[c - 1];
        uint64_t b0 = c * SB, bcnt = n1 - b0; if (bcnt > SB) bcnt = SB;
        double run = 0.0, before = 0.0; uint64_t b = b0 + bcnt - 1; int found = 0;
        for (uint64_t j = 0; j < bcnt; j++) {
            double nr = run + l1[b0 + j];
            if (rule_hit(acc + nr, p, strict)) { b = b0 + j; before = run; found = 1; break; }
            run = nr;
        }
        if (!found) {           /* fell off the chunk: last block, running sum before it */
            run = 0.0;
            for (uint64_t j = 0; j + 1 < bcnt; j++) run += l1[b0 + j];
            before = run;
        }
        double acc2 = acc + before;
        uint64_t i0 = b * SB, icnt = n_amps - i0; if (icnt > SB) icnt = SB;
        uint64_t r = i0 + icnt - 1; run = 0.0;
        for (uint64_t j = 0; j < icnt; j++) {
            run += prob(psi[i0 + j]);
            if (rule_hit(acc2 + run, p, strict)) { r = i0 + j; break; }
        }
        out[t] = r;
    }
    free(l1); free(l2); free(top);
}

void orc_sample_tree(const double *psi_, uint64_t n_amps, const double *u, uint64_t n_shots,
                     int strict, uint64_t *out) {
    sample_tree_base((const cplx *)psi_, n_amps, u, n_shots, strict, 0.0, out, NULL);
}

/* one shard's pieces of the sharded order (used by the CPU stand-in engine of the gloo tests) */
double orc_sample_tree_total(const double *psi_, uint64_t n_amps) {
    uint64_t dummy_out;
    double dummy_u = 2.0, tot = 0.0;
    sample_tree_base((const cplx *)psi_, n_amps, &dummy_u, 0, 0, 0.0, &dummy_out, &tot);
    return tot;
}
void orc_sample_tree_base(const double *psi_, uint64_t n_amps, const double *u, uint64_t n_shots, int strict, double base,
                          uint64_t *out) {
    sample_tree_base((const cplx *)psi_, n_amps, u, n_shots, strict, base, out, NULL);
}

/* The SHARDED sampler's summation order (qvm_b200/dist.py ShardedState.sample): the vector is WORLD contiguous shards
 * (dqvm's layout: the top index bits select the rank, dqvm/src/global-addresses.lisp:99-151).  Every shard builds its
 * own tree; the shard totals are added left to right into a prefix; a draw belongs to the first shard whose inclusive
 * prefix hits it (the tie rule of the single-device sampler) and is resolved there with the exclusive prefix as the
 * starting accumulator.  Restates SAMPLE-WAVEFUNCTION-MULTIPLE-TIMES / the bisection sampler
 * (src/measurement.lisp:179-225, 246-288) for a partitioned CDF. */
void orc_sample_tree_sharded(const double *psi_, uint64_t n_amps, int world, const double *u, uint64_t n_shots,
                             int strict, uint64_t *out) {
    const cplx *psi = (const cplx *)psi_;
    const uint64_t len = n_amps / (uint64_t)world;
    double *prefix = (double *)malloc(sizeof(double) * ((size_t)world + 1));
    uint64_t dummy_out;
    double dummy_u = 2.0;
    prefix[0] = 0.0;
    for (int s = 0; s < world; s++) {
        double tot = 0.0;
        sample_tree_base(psi + (uint64_t)s * len, len, &dummy_u, 0, strict, 0.0, &dummy_out, &tot);
        prefix[s + 1] = prefix[s] + tot;
    }
    for (uint64_t t = 0; t < n_shots; t++) {
        int owner = world - 1;
        for (int s = 0; s < world; s++)
            if (rule_hit(prefix[s + 1], u[t], strict)) { owner = s; break; }
        uint64_t local;
        sample_tree_base(psi + (uint64_t)owner * len, len, u + t, 1, strict, prefix[owner], &local, NULL);
        out[t] = (uint64_t)owner * len + local;
    }
    free(prefix);
}

/* %EVOLVE-PURE-STATE-STOCHASTICALLY src/apply-gate.lisp:16-39 with the uniform draw r supplied by the
 * caller: for j = 0..m-1: trial <- psi; trial <- K_j trial; summed += |trial|^2; stop when summed >= r;
 * then psi <- trial and NORMALIZE-WAVEFUNCTION.  Returns the index of the Kraus operator applied. */
int orc_evolve_stochastic(double *psi_, uint64_t n_amps, int k, const int *qubits, int m, const double *kraus, double r) {
    uint64_t d = 1ULL << k;
    cplx *trial = (cplx *)malloc(sizeof(cplx) * n_amps);
    double summed = 0.0;
    int j = 0;
    for (j = 0; j < m; j++) {
        memcpy(trial, psi_, sizeof(cplx) * n_amps);
        orc_apply_matrix((double *)trial, n_amps, k, qubits, kraus + 2 * d * d * (uint64_t)j);
        summed += orc_norm2((const double *)trial, n_amps);
        if (summed >= r) break;
    }
    if (j == m) j = m - 1;
    memcpy(psi_, trial, sizeof(cplx) * n_amps);
    free(trial);
    orc_normalize(psi_, n_amps);
    return j;
}
