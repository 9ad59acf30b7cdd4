/* baseline/cpu_port.c -- CPU BASELINE (timing only; the checker is oracle/qvm_oracle.c, not this file).
 *
 * What the reference's fastest CPU configuration does for APPLY-MATRIX-OPERATOR, restated in C so that it can be
 * timed on the GPU box's host cores (the reference itself needs SBCL, which the image lacks):
 *   - 1q / 2q gates through AVX2 + FMA kernels with the instruction pattern of the reference's SBCL VOPs
 *     (src/impl/sbcl-avx-vops.lisp:210-355: matmul2-simd = vmulpd, vfmadd231pd, vfmaddsub231pd, vfmadd231pd on
 *     [row1 | row0] packed registers; matmul4-simd-half likewise with four columns), called as in
 *     src/serial-kernels.lisp:177-267;
 *   - the amplitude groups of a gate are split into one CONTIGUOUS range per worker, as WITH-PARALLEL-SUBDIVISIONS /
 *     LPARALLEL:PDOTIMES do (src/utilities.lisp:346-383), here with OpenMP static scheduling;
 *   - every gate is one full pass over the state (the reference's per-gate transitions), k >= 3 through the scalar
 *     gather / matvec / scatter of APPLY-OPERATOR (src/wavefunction.lisp:234-306).
 * Build: gcc -O3 -mavx2 -mfma -fopenmp (baseline/Makefile).  Checked against the oracle in tests/test_cpu_port.py.
 */
#include <immintrin.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double re, im; } cplx;

static inline uint64_t inject0(uint64_t i, int pos) {
    const uint64_t lo = i & ((1ull << pos) - 1ull);
    return ((i >> pos) << (pos + 1)) | lo;
}

/* [m_r1 m_r1 m_r0 m_r0] style packing of one matrix column (2x2matrix-to-simd, sbcl-avx-vops.lisp:150-198):
 * lanes (low to high) = row0, row0, row1, row1 */
typedef struct { __m256d c0r, c0i, c1r, c1i; } m2pack;

static m2pack pack2(const double* m) {   /* m = row-major 2x2 complex: m00 m01 m10 m11 */
    m2pack p;
    p.c0r = _mm256_set_pd(m[4], m[4], m[0], m[0]);
    p.c0i = _mm256_set_pd(m[5], m[5], m[1], m[1]);
    p.c1r = _mm256_set_pd(m[6], m[6], m[2], m[2]);
    p.c1i = _mm256_set_pd(m[7], m[7], m[3], m[3]);
    return p;
}

/* (p, q) = M (a0, a1): acc lanes = [p.re p.im q.re q.im] */
static inline __m256d matmul2(const m2pack* M, __m128d a0, __m128d a1) {
    const __m256d aa0 = _mm256_set_m128d(a0, a0), aa1 = _mm256_set_m128d(a1, a1);         /* [re im re im] */
    const __m256d sw0 = _mm256_permute_pd(aa0, 0x5), sw1 = _mm256_permute_pd(aa1, 0x5);    /* [im re im re] */
    __m256d acc = _mm256_mul_pd(M->c0i, sw0);
    acc = _mm256_fmadd_pd(M->c1i, sw1, acc);              /* imaginary parts of the matrix times swapped amplitudes */
    acc = _mm256_fmaddsub_pd(M->c0r, aa0, acc);           /* re lane: m.re*a.re - m.im*a.im ; im lane: m.re*a.im + m.im*a.re */
    {   /* second column: its imaginary contribution was added with +, fix the sign on the real lanes */
        /* fmaddsub negated BOTH imaginary products on the real lanes (acc held their sum), which is what we want */
    }
    acc = _mm256_fmadd_pd(M->c1r, aa1, acc);
    return acc;
}

void cpb_apply_1q(cplx* psi, int n, int q, const double* m, int threads) {
    const m2pack M = pack2(m);
    const uint64_t half = 1ull << (n - 1), stride = 1ull << q;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (uint64_t i = 0; i < half; i++) {
        const uint64_t a = inject0(i, q);
        double* p0 = (double*)(psi + a);
        double* p1 = (double*)(psi + a + stride);
        const __m256d r = matmul2(&M, _mm_loadu_pd(p0), _mm_loadu_pd(p1));
        _mm_storeu_pd(p0, _mm256_castpd256_pd128(r));
        _mm_storeu_pd(p1, _mm256_extractf128_pd(r, 1));
    }
}

/* two rows (r, r+1) of a 4x4 matrix packed per column (2x4matrix-to-simd) */
typedef struct { __m256d cr[4], ci[4]; } m4half;

static m4half pack4(const double* m, int row) {   /* rows row, row+1 */
    m4half p;
    for (int c = 0; c < 4; c++) {
        const double* e0 = m + 2 * (4 * row + c);
        const double* e1 = m + 2 * (4 * (row + 1) + c);
        p.cr[c] = _mm256_set_pd(e1[0], e1[0], e0[0], e0[0]);
        p.ci[c] = _mm256_set_pd(e1[1], e1[1], e0[1], e0[1]);
    }
    return p;
}

static inline __m256d matmul4_half(const m4half* M, const __m128d a[4]) {
    __m256d aa[4], sw[4];
    for (int c = 0; c < 4; c++) {
        aa[c] = _mm256_set_m128d(a[c], a[c]);
        sw[c] = _mm256_permute_pd(aa[c], 0x5);
    }
    __m256d acc = _mm256_mul_pd(M->ci[0], sw[0]);
    acc = _mm256_fmadd_pd(M->ci[1], sw[1], acc);
    acc = _mm256_fmadd_pd(M->ci[2], sw[2], acc);
    acc = _mm256_fmadd_pd(M->ci[3], sw[3], acc);
    acc = _mm256_fmaddsub_pd(M->cr[0], aa[0], acc);
    acc = _mm256_fmadd_pd(M->cr[1], aa[1], acc);
    acc = _mm256_fmadd_pd(M->cr[2], aa[2], acc);
    acc = _mm256_fmadd_pd(M->cr[3], aa[3], acc);
    return acc;
}

/* qubits in NAT-TUPLE order: q0 = matrix index bit 0, q1 = bit 1 */
void cpb_apply_2q(cplx* psi, int n, int q0, int q1, const double* m, int threads) {
    const m4half lo = pack4(m, 0), hi = pack4(m, 2);
    const int pa = q0 < q1 ? q0 : q1, pb = q0 < q1 ? q1 : q0;
    const uint64_t quarter = 1ull << (n - 2), s0 = 1ull << q0, s1 = 1ull << q1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (uint64_t i = 0; i < quarter; i++) {
        const uint64_t base = inject0(inject0(i, pa), pb);
        double* p[4] = {(double*)(psi + base), (double*)(psi + base + s0), (double*)(psi + base + s1), (double*)(psi + base + s0 + s1)};
        __m128d a[4];
        for (int c = 0; c < 4; c++) a[c] = _mm_loadu_pd(p[c]);
        const __m256d r01 = matmul4_half(&lo, a), r23 = matmul4_half(&hi, a);
        _mm_storeu_pd(p[0], _mm256_castpd256_pd128(r01));
        _mm_storeu_pd(p[1], _mm256_extractf128_pd(r01, 1));
        _mm_storeu_pd(p[2], _mm256_castpd256_pd128(r23));
        _mm_storeu_pd(p[3], _mm256_extractf128_pd(r23, 1));
    }
}

/* generic k: gather, scalar matvec in the reference's order, scatter (src/wavefunction.lisp:234-306) */
void cpb_apply_kq(cplx* psi, int n, int k, const int* qubits, const double* m, int threads) {
    const uint64_t groups = 1ull << (n - k), d = 1ull << k;
    int sorted[16];
    memcpy(sorted, qubits, sizeof(int) * (size_t)k);
    for (int i = 1; i < k; i++) {
        int v = sorted[i], j = i - 1;
        while (j >= 0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; j--; }
        sorted[j + 1] = v;
    }
#pragma omp parallel num_threads(threads)
    {
        cplx* col = (cplx*)malloc(sizeof(cplx) * d);
        cplx* out = (cplx*)malloc(sizeof(cplx) * d);
#pragma omp for schedule(static)
        for (uint64_t i = 0; i < groups; i++) {
            uint64_t base = i;
            for (int j = 0; j < k; j++) base = inject0(base, sorted[j]);
            for (uint64_t c = 0; c < d; c++) {
                uint64_t a = base;
                for (int j = 0; j < k; j++)
                    if (c >> j & 1) a |= 1ull << qubits[j];
                col[c] = psi[a];
            }
            for (uint64_t r = 0; r < d; r++) {
                double sr = 0.0, si = 0.0;
                for (uint64_t c = 0; c < d; c++) {
                    const double mr = m[2 * (r * d + c)], mi = m[2 * (r * d + c) + 1];
                    sr += mr * col[c].re - mi * col[c].im;
                    si += mr * col[c].im + mi * col[c].re;
                }
                out[r].re = sr;
                out[r].im = si;
            }
            for (uint64_t c = 0; c < d; c++) {
                uint64_t a = base;
                for (int j = 0; j < k; j++)
                    if (c >> j & 1) a |= 1ull << qubits[j];
                psi[a] = out[c];
            }
        }
        free(col);
        free(out);
    }
}

void cpb_apply_gate(double* psi, int n, int k, const int* qubits, const double* m, int threads) {
    if (threads < 1) threads = 1;
    if (k == 1) cpb_apply_1q((cplx*)psi, n, qubits[0], m, threads);
    else if (k == 2) cpb_apply_2q((cplx*)psi, n, qubits[0], qubits[1], m, threads);
    else cpb_apply_kq((cplx*)psi, n, k, qubits, m, threads);
}

/* first touch in parallel so that pages are spread like the workers' ranges */
void cpb_zero_state(double* psi, int n, int threads) {
    const uint64_t N = 1ull << n;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (uint64_t i = 0; i < N; i++) {
        psi[2 * i] = 0.0;
        psi[2 * i + 1] = 0.0;
    }
    psi[0] = 1.0;
}

/* all logical CPUs of the machine, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1) */
int cpb_all_cores(void) { return omp_get_num_procs(); }
