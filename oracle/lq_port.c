/*
 * lq_port.c -- TEST INFRASTRUCTURE / CPU BASELINE ONLY (never linked into the product).
 *
 * C + OpenMP restatement of the CPU path the reference is compared with, `lightning.qubit`
 * (pennylane-lightning >= 0.30, requirements.txt:7 of the reference; NOT under /root/reference and
 * not installable here, so this is kind = "port" in bench.py's cpu_baseline).  Algorithm restated:
 * in-place, pair-strided kernels over the 2^(n-k) amplitude groups with rev_wire = n - 1 - wire,
 * `#pragma omp parallel for` over groups; specialised kernels for diagonal gates, X-type
 * (swap-only) gates and generic 1-/2-/k-qubit matrices, as lightning.qubit's "LM" kernels do.
 * The op-level semantics (which matrix, which wires are controls) come from oracle/np_oracle.py,
 * which is pinned by the reference's golden vectors; oracle/lq_port.py checks this file against
 * np_oracle on random circuits (tests/test_oracle_lq_port.py).
 *
 * Call sites in the reference this stands in for: the comparisons of lightning.gpu results with
 * lightning.qubit at tests/test_adjoint_jacobian.py:805-881 and the CPU fallback class at
 * pennylane_lightning_gpu/lightning_gpu.py:975-998.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC lq_port.c -o _build/liblq_port.so
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef double complex cd;

int lq_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void lq_set_num_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* insert zero bits at the (ascending) positions holes[0..m) */
static inline uint64_t expand(uint64_t o, const int *holes, int m) {
    for (int j = 0; j < m; ++j) {
        const int p = holes[j];
        o = ((o >> p) << (p + 1)) | (o & ((1ull << p) - 1ull));
    }
    return o;
}

static int cmp_int(const void *a, const void *b) { return *(const int *)a - *(const int *)b; }

static int collect_holes(const int *tgt_bits, int k, uint64_t ctrl_mask, int *holes) {
    int m = 0;
    for (int i = 0; i < k; ++i) holes[m++] = tgt_bits[i];
    for (int b = 0; b < 64; ++b)
        if ((ctrl_mask >> b) & 1ull) holes[m++] = b;
    qsort(holes, m, sizeof(int), cmp_int);
    return m;
}

/* generic single-qubit gate, no controls: the inner loop of lightning.qubit's applySingleQubitOp */
void lq_apply_1q(cd *sv, int n, const cd *m, int bit) {
    const uint64_t half = 1ull << (n - 1);
    const uint64_t stride = 1ull << bit;
    const cd m00 = m[0], m01 = m[1], m10 = m[2], m11 = m[3];
#pragma omp parallel for schedule(static)
    for (uint64_t g = 0; g < half; ++g) {
        const uint64_t i0 = ((g >> bit) << (bit + 1)) | (g & (stride - 1));
        const uint64_t i1 = i0 | stride;
        const cd a = sv[i0], b = sv[i1];
        sv[i0] = m00 * a + m01 * b;
        sv[i1] = m10 * a + m11 * b;
    }
}

/* 2^k x 2^k row-major matrix on tgt_bits (tgt_bits[0] = most significant matrix bit), applied where
 * all ctrl_mask bits are 1 */
void lq_apply_dense(cd *sv, int n, const cd *mat, const int *tgt_bits, int k, uint64_t ctrl_mask) {
    if (k == 1 && ctrl_mask == 0) {
        lq_apply_1q(sv, n, mat, tgt_bits[0]);
        return;
    }
    int holes[64];
    const int m = collect_holes(tgt_bits, k, ctrl_mask, holes);
    const int dim = 1 << k;
    uint64_t offs[64];
    if (k > 6) return;
    for (int j = 0; j < dim; ++j) {
        uint64_t o = 0;
        for (int b = 0; b < k; ++b)
            if ((j >> (k - 1 - b)) & 1) o |= 1ull << tgt_bits[b];
        offs[j] = o;
    }
    const uint64_t groups = 1ull << (n - m);
#pragma omp parallel for schedule(static)
    for (uint64_t g = 0; g < groups; ++g) {
        const uint64_t base = expand(g, holes, m) | ctrl_mask;
        cd x[64];
        for (int j = 0; j < dim; ++j) x[j] = sv[base + offs[j]];
        for (int r = 0; r < dim; ++r) {
            cd y = 0;
            for (int c = 0; c < dim; ++c) y += mat[r * dim + c] * x[c];
            sv[base + offs[r]] = y;
        }
    }
}

/* amp *= diag[t], t = index bits at tgt_bits (MSB first); only where ctrl bits are 1 */
void lq_apply_diag(cd *sv, int n, const cd *diag, const int *tgt_bits, int k, uint64_t ctrl_mask) {
    int holes[64];
    const int m = collect_holes(tgt_bits, 0, ctrl_mask, holes);
    const uint64_t items = 1ull << (n - m);
#pragma omp parallel for schedule(static)
    for (uint64_t o = 0; o < items; ++o) {
        const uint64_t i = expand(o, holes, m) | ctrl_mask;
        int t = 0;
        for (int b = 0; b < k; ++b) t = (t << 1) | (int)((i >> tgt_bits[b]) & 1ull);
        sv[i] *= diag[t];
    }
}

/* amp *= (popcount(i & zmask) odd ? odd : even) */
void lq_apply_parity(cd *sv, int n, uint64_t zmask, const cd *even_odd, uint64_t ctrl_mask) {
    int holes[64];
    const int m = collect_holes(NULL, 0, ctrl_mask, holes);
    const uint64_t items = 1ull << (n - m);
    const cd e = even_odd[0], od = even_odd[1];
#pragma omp parallel for schedule(static)
    for (uint64_t o = 0; o < items; ++o) {
        const uint64_t i = expand(o, holes, m) | ctrl_mask;
        sv[i] *= (__builtin_popcountll(i & zmask) & 1) ? od : e;
    }
}

/* X-type gate (PauliX / CNOT / Toffoli): swap the pair where the controls are 1 */
void lq_apply_x(cd *sv, int n, int bit, uint64_t ctrl_mask) {
    int holes[64];
    const int m = collect_holes(&bit, 1, ctrl_mask, holes);
    const uint64_t groups = 1ull << (n - m);
    const uint64_t stride = 1ull << bit;
#pragma omp parallel for schedule(static)
    for (uint64_t g = 0; g < groups; ++g) {
        const uint64_t i0 = expand(g, holes, m) | ctrl_mask;
        const cd a = sv[i0];
        sv[i0] = sv[i0 | stride];
        sv[i0 | stride] = a;
    }
}

void lq_copy(cd *dst, const cd *src, int n) {
    const uint64_t len = 1ull << n;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < len; ++i) dst[i] = src[i];
}

void lq_set_basis(cd *sv, int n, uint64_t index) {
    const uint64_t len = 1ull << n;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < len; ++i) sv[i] = 0;
    sv[index] = 1;
}

/* <a|b> */
void lq_inner(const cd *a, const cd *b, int n, double *out) {
    const uint64_t len = 1ull << n;
    double re = 0, im = 0;
#pragma omp parallel for schedule(static) reduction(+ : re, im)
    for (uint64_t i = 0; i < len; ++i) {
        const cd v = conj(a[i]) * b[i];
        re += creal(v);
        im += cimag(v);
    }
    out[0] = re;
    out[1] = im;
}

/* <bra| P |ket>, (P ket)_i = i^ny (-1)^{popc((i^x)&z)} ket_{i^x} */
void lq_bra_pauli_ket(const cd *bra, const cd *ket, int n, uint64_t x, uint64_t z, int ny, double *out) {
    const uint64_t len = 1ull << n;
    double re = 0, im = 0;
#pragma omp parallel for schedule(static) reduction(+ : re, im)
    for (uint64_t i = 0; i < len; ++i) {
        const uint64_t j = i ^ x;
        const double s = (__builtin_popcountll(j & z) & 1) ? -1.0 : 1.0;
        const cd v = conj(bra[i]) * ket[j] * s;
        re += creal(v);
        im += cimag(v);
    }
    cd tot = re + im * I;
    for (int q = 0; q < (ny & 3); ++q) tot *= I;
    out[0] = creal(tot);
    out[1] = cimag(tot);
}

/* out += coeff * P in  (Hamiltonian accumulation) */
void lq_pauli_axpy(cd *out, const cd *in, int n, uint64_t x, uint64_t z, double cre, double cim) {
    const uint64_t len = 1ull << n;
    const cd c = cre + cim * I;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < len; ++i) {
        const uint64_t j = i ^ x;
        const double s = (__builtin_popcountll(j & z) & 1) ? -1.0 : 1.0;
        out[i] += c * s * in[j];
    }
}

void lq_zero(cd *sv, int n) {
    const uint64_t len = 1ull << n;
#pragma omp parallel for schedule(static)
    for (uint64_t i = 0; i < len; ++i) sv[i] = 0;
}
