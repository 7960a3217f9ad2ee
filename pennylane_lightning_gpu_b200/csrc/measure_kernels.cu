// Measurement / reduction kernels: every one reads the state once and reduces with warp shuffles,
// accumulating in FP64 regardless of the state precision.
//
// Replaces, in the reference,
//   custatevecComputeExpectation            simulator/StateVectorCudaManaged.hpp:1602-1638, :1674-1711
//   custatevecComputeExpectationsOnPauliBasis                          Managed.hpp:1117-1128
//   custatevecAbs2SumArray                                             Managed.hpp:957-967
//   custatevecSampler*                                                 Managed.hpp:1015-1035
//   cusparseSpMV + cublas<Z|C>dotc          Managed.hpp:854-916, algorithms/ObservablesGPU.hpp:523-584
//   cublas<Z|C>dotc in the adjoint loop     util/cuda_helpers.hpp:456-497, algorithms/AdjointDiffGPU.hpp:145
// The "bra-op-ket" kernels fuse  mu = G lambda ; <bra|mu>  into one read of bra and lambda, so the
// adjoint sweep never materialises mu (north_star item (d)).
#include <algorithm>
#include <cstdlib>

#include "device_utils.cuh"
#include "qsv_internal.h"

namespace qsv {

namespace {

constexpr int RNT = 256;                  // threads per reduction block
constexpr int RGRID = NUM_SMS * 8;        // persistent-style grid for grid-stride reductions

template <int K> struct Offs {
    uint64_t v[1 << K];
};
template <typename T, int K> struct MatP {
    T re[1 << (2 * K)];
    T im[1 << (2 * K)];
};

// sum over groups of  conj(bra[r]) * M[r][c] * ket[c]
template <typename T, int K, int V>
__global__ void __launch_bounds__(RNT)
    k_bra_dense_ket(const void *__restrict__ bra, const void *__restrict__ ket, uint64_t n_groups,
                    Holes holes, uint64_t ctrl, Offs<K> offs, MatP<T, K> m, double *out) {
    constexpr int D = 1 << K;
    const bool same = bra == ket;
    double acc_re = 0, acc_im = 0;
    const uint64_t stride = (uint64_t)gridDim.x * RNT;
    for (uint64_t g = (uint64_t)blockIdx.x * RNT + threadIdx.x; g < n_groups; g += stride) {
        const uint64_t base = expand_index(g, holes) | ctrl;
        T x[D][2 * V], b[D][2 * V];
#pragma unroll
        for (int d = 0; d < D; ++d) load_elem<T, V>(x[d], ket, base + offs.v[d]);
        if (!same) {
#pragma unroll
            for (int d = 0; d < D; ++d) load_elem<T, V>(b[d], bra, base + offs.v[d]);
        } else {
#pragma unroll
            for (int d = 0; d < D; ++d)
#pragma unroll
                for (int a = 0; a < 2 * V; ++a) b[d][a] = x[d][a];
        }
#pragma unroll
        for (int r = 0; r < D; ++r) {
#pragma unroll
            for (int a = 0; a < V; ++a) {
                double yr = 0, yi = 0;
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const double mr = m.re[r * D + c], mi = m.im[r * D + c];
                    yr += mr * (double)x[c][2 * a] - mi * (double)x[c][2 * a + 1];
                    yi += mr * (double)x[c][2 * a + 1] + mi * (double)x[c][2 * a];
                }
                const double br = b[r][2 * a], bi = b[r][2 * a + 1];
                acc_re += br * yr + bi * yi;
                acc_im += br * yi - bi * yr;
            }
        }
    }
    block_accumulate<RNT>(acc_re, acc_im, out);
}

// large-k dense: one CTA per group, warp per row (see k_apply_dense_large)
template <typename T>
__global__ void __launch_bounds__(256)
    k_bra_dense_ket_large(const void *__restrict__ bra, const void *__restrict__ ket, int k,
                          uint64_t n_groups, Holes holes, uint64_t ctrl, const uint64_t *__restrict__ offs,
                          const double2 *__restrict__ mat, double *out) {
    extern __shared__ double2 s_x[];
    const int D = 1 << k;
    double2 *s_b = s_x + D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double acc_re = 0, acc_im = 0;
    for (uint64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const uint64_t base = expand_index(g, holes) | ctrl;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            T c[2];
            load_elem<T, 1>(c, ket, base + offs[d]);
            s_x[d] = make_double2((double)c[0], (double)c[1]);
            load_elem<T, 1>(c, bra, base + offs[d]);
            s_b[d] = make_double2((double)c[0], (double)c[1]);
        }
        __syncthreads();
        for (int r = warp; r < D; r += 8) {
            double re = 0, im = 0;
            const double2 *row = mat + (size_t)r * D;
            for (int c = lane; c < D; c += 32) {
                const double2 mm = row[c];
                const double2 xx = s_x[c];
                re += mm.x * xx.x - mm.y * xx.y;
                im += mm.x * xx.y + mm.y * xx.x;
            }
            const double2 bb = s_b[r];
            acc_re += bb.x * re + bb.y * im;
            acc_im += bb.x * im - bb.y * re;
        }
        __syncthreads();
    }
    block_accumulate<256>(acc_re, acc_im, out);
}

template <typename T> struct DiagP {
    int k;
    int parity;
    unsigned char tbits[4];
    uint64_t zmask;
    T re[16];
    T im[16];
};

// sum over (control-selected) i of conj(bra[i]) * d(i) * ket[i]
template <typename T, int V>
__global__ void __launch_bounds__(RNT)
    k_bra_diag_ket(const void *__restrict__ bra, const void *__restrict__ ket, uint64_t n_items,
                   Holes holes, uint64_t ctrl, DiagP<double> d, double *out) {
    const bool same = bra == ket;
    double acc_re = 0, acc_im = 0;
    const uint64_t stride = (uint64_t)gridDim.x * RNT;
#pragma unroll 2
    for (uint64_t o = (uint64_t)blockIdx.x * RNT + threadIdx.x; o < n_items; o += stride) {
        const uint64_t e = expand_index(o, holes) | ctrl;
        T x[2 * V], b[2 * V];
        load_elem<T, V>(x, ket, e);
        if (!same) {
            load_elem<T, V>(b, bra, e);
        } else {
#pragma unroll
            for (int a = 0; a < 2 * V; ++a) b[a] = x[a];
        }
#pragma unroll
        for (int a = 0; a < V; ++a) {
            const uint64_t i = e * V + a;
            int t = 0;
            if (d.parity) {
                t = __popcll(i & d.zmask) & 1;
            } else {
                for (int q = 0; q < d.k; ++q) t = (t << 1) | (int)((i >> d.tbits[q]) & 1ull);
            }
            const double pr = d.re[t], pi = d.im[t];
            const double yr = pr * (double)x[2 * a] - pi * (double)x[2 * a + 1];
            const double yi = pr * (double)x[2 * a + 1] + pi * (double)x[2 * a];
            const double br = b[2 * a], bi = b[2 * a + 1];
            acc_re += br * yr + bi * yi;
            acc_im += br * yi - bi * yr;
        }
    }
    block_accumulate<RNT>(acc_re, acc_im, out);
}

// <bra| P |ket>,  (P ket)_i = i^ny (-1)^{popc((i^x) & z)} ket_{i^x}
template <typename T, int V>
__global__ void __launch_bounds__(RNT)
    k_bra_pauli_ket(const void *__restrict__ bra, const void *__restrict__ ket, uint64_t n_elems,
                    uint64_t xmask, uint64_t zmask, int ny, double *out) {
    double acc_re = 0, acc_im = 0;
    const uint64_t stride = (uint64_t)gridDim.x * RNT;
    const uint64_t xe = V == 2 ? xmask >> 1 : xmask;
    const bool swap = V == 2 && (xmask & 1ull);
#pragma unroll 2
    for (uint64_t e = (uint64_t)blockIdx.x * RNT + threadIdx.x; e < n_elems; e += stride) {
        T b[2 * V], x[2 * V];
        load_elem<T, V>(b, bra, e);
        load_elem<T, V>(x, ket, e ^ xe);
#pragma unroll
        for (int a = 0; a < V; ++a) {
            const int pa = swap ? (a ^ 1) : a;  // position of the partner inside its element
            const uint64_t j = ((e ^ xe) * V + pa);
            const double sgn = (__popcll(j & zmask) & 1) ? -1.0 : 1.0;
            double xr, xi;
            if constexpr (V == 2) {
                xr = swap ? x[2 * (a ^ 1)] : x[2 * a];
                xi = swap ? x[2 * (a ^ 1) + 1] : x[2 * a + 1];
            } else {
                xr = x[0];
                xi = x[1];
            }
            const double br = b[2 * a], bi = b[2 * a + 1];
            acc_re += sgn * (br * xr + bi * xi);
            acc_im += sgn * (br * xi - bi * xr);
        }
    }
    // multiply by i^ny
    double re = acc_re, im = acc_im;
    switch (ny & 3) {
    case 1: re = -acc_im; im = acc_re; break;
    case 2: re = -acc_re; im = -acc_im; break;
    case 3: re = acc_im; im = -acc_re; break;
    default: break;
    }
    block_accumulate<RNT>(re, im, out);
}

// out[i] = sum_t coeff_t * (-1)^{popc((i^x_t) & z_t)} * in[i ^ x_t]      (coeff_t includes i^ny_t)
template <typename T>
__global__ void __launch_bounds__(256)
    k_pauli_sum_apply(const void *__restrict__ in, void *__restrict__ out, uint64_t length, int n_terms,
                      const uint64_t *__restrict__ xm, const uint64_t *__restrict__ zm,
                      const double2 *__restrict__ cf, int accumulate) {
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < length; i += stride) {
        double yr = 0, yi = 0;
        for (int t = 0; t < n_terms; ++t) {
            const uint64_t j = i ^ xm[t];
            T c[2];
            load_elem<T, 1>(c, in, j);
            const double2 w = cf[t];
            const double s = (__popcll(j & zm[t]) & 1) ? -1.0 : 1.0;
            yr += s * (w.x * (double)c[0] - w.y * (double)c[1]);
            yi += s * (w.x * (double)c[1] + w.y * (double)c[0]);
        }
        if (accumulate) {
            T o[2];
            load_elem<T, 1>(o, out, i);
            yr += (double)o[0];
            yi += (double)o[1];
        }
        T y[2] = {(T)yr, (T)yi};
        store_elem<T, 1>(out, i, y);
    }
}

// Marginal probabilities at HBM speed, any measured wires.  The state is read in ROWS of 32 consecutive amplitudes (one
// per lane: 512 / 256 contiguous bytes), so the five lowest index bits are lane bits.  The measured bits above them
// select a GROUP of rows (= the high part of the output bin), the other high bits are summed: a warp takes one chunk of
// one group's rows, every lane adds up |psi|^2 of its column, lanes that differ only in SUMMED low bits are combined with
// shuffles, and the remaining lanes -- one per value of the measured low bits -- write / add their bin.  No shared-memory
// atomics (the first version spent 27 ms on 8 hot bins at 30 qubits), no strided reads.
struct ProbsPlan {
    Holes row_holes;            // measured high bits as positions in the row index (index bit - 5), ascending
    unsigned char hi_out[40];   // output-bin bit of the j-th measured high bit
    unsigned char hi_pos[40];   // its position in the row index
    int n_hi;
    unsigned lane_sum_mask;     // lane bits that are summed over
    unsigned char lane_out[5];  // output-bin bit of lane bit b (255: summed)
    int rows_log2;              // log2(rows per group)
    int chunk_log2;             // log2(rows per chunk)
    int use_atomics;
};

template <typename T>
__global__ void __launch_bounds__(256)
    k_probs_rows(const void *__restrict__ sv, double *__restrict__ out, uint64_t n_items, const __grid_constant__ ProbsPlan P) {
    using A = typename VecOf<T, 1>::type;
    const A *psi = reinterpret_cast<const A *>(sv);
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int chunks_log2 = P.rows_log2 - P.chunk_log2;
    uint64_t t_low = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b)
        if (P.lane_out[b] != 255 && ((lane >> b) & 1u)) t_low |= 1ull << P.lane_out[b];
    for (uint64_t w = (((uint64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; w < n_items; w += warps) {
        const uint64_t g = w >> chunks_log2, c = w & ((1ull << chunks_log2) - 1ull);
        uint64_t base_row = 0, t_high = 0;
        for (int j = 0; j < P.n_hi; ++j)
            if ((g >> j) & 1ull) {
                base_row |= 1ull << P.hi_pos[j];
                t_high |= 1ull << P.hi_out[j];
            }
        const uint64_t s0 = c << P.chunk_log2, s1 = s0 + (1ull << P.chunk_log2);
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
        uint64_t s = s0;
        for (; s + 4 <= s1; s += 4) {
            const A a0 = psi[((expand_index(s, P.row_holes) | base_row) << 5) + lane];
            const A a1 = psi[((expand_index(s + 1, P.row_holes) | base_row) << 5) + lane];
            const A a2 = psi[((expand_index(s + 2, P.row_holes) | base_row) << 5) + lane];
            const A a3 = psi[((expand_index(s + 3, P.row_holes) | base_row) << 5) + lane];
            acc0 += (double)a0.x * (double)a0.x + (double)a0.y * (double)a0.y;
            acc1 += (double)a1.x * (double)a1.x + (double)a1.y * (double)a1.y;
            acc2 += (double)a2.x * (double)a2.x + (double)a2.y * (double)a2.y;
            acc3 += (double)a3.x * (double)a3.x + (double)a3.y * (double)a3.y;
        }
        for (; s < s1; ++s) {
            const A a0 = psi[((expand_index(s, P.row_holes) | base_row) << 5) + lane];
            acc0 += (double)a0.x * (double)a0.x + (double)a0.y * (double)a0.y;
        }
        double acc = (acc0 + acc1) + (acc2 + acc3);
#pragma unroll
        for (int b = 0; b < 5; ++b)
            if ((P.lane_sum_mask >> b) & 1u) acc += __shfl_xor_sync(0xffffffffu, acc, 1 << b);
        if ((lane & P.lane_sum_mask) == 0u) {
            if (P.use_atomics)
                atomicAdd(out + (t_high | t_low), acc);
            else
                out[t_high | t_low] = acc;
        }
    }
}

// marginal probabilities, small output (<= 2^11 bins): shared-memory histogram per block (registers below 10 qubits)
template <typename T>
__global__ void __launch_bounds__(256)
    k_probs_small(const void *__restrict__ sv, uint64_t length, int k, const unsigned char *__restrict__ bits,
                  double *out) {
    extern __shared__ double s_bins[];
    const int nb = 1 << k;
    for (int i = threadIdx.x; i < nb; i += 256) s_bins[i] = 0.0;
    __syncthreads();
    unsigned char lb[16];
    for (int q = 0; q < k; ++q) lb[q] = bits[q];
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < length; i += stride) {
        T c[2];
        load_elem<T, 1>(c, sv, i);
        const double p = (double)c[0] * (double)c[0] + (double)c[1] * (double)c[1];
        int t = 0;
        for (int q = 0; q < k; ++q) t |= (int)((i >> lb[q]) & 1ull) << q;
        atomicAdd(&s_bins[t], p);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nb; i += 256)
        if (s_bins[i] != 0.0) atomicAdd(&out[i], s_bins[i]);
}

// marginal probabilities, large output: one thread per output bin, loop over the traced-out bits
template <typename T>
__global__ void __launch_bounds__(256)
    k_probs_large(const void *__restrict__ sv, int n, int k, const unsigned char *__restrict__ bits,
                  Holes kept_sorted, double *out) {
    const uint64_t nb = 1ull << k;
    const uint64_t n_rest = 1ull << (n - k);
    const uint64_t stride = (uint64_t)gridDim.x * 256;
    for (uint64_t t = (uint64_t)blockIdx.x * 256 + threadIdx.x; t < nb; t += stride) {
        uint64_t fixed = 0;
        for (int q = 0; q < k; ++q) fixed |= ((t >> q) & 1ull) << bits[q];
        double p = 0;
        for (uint64_t r = 0; r < n_rest; ++r) {
            const uint64_t i = expand_index(r, kept_sorted) | fixed;
            T c[2];
            load_elem<T, 1>(c, sv, i);
            p += (double)c[0] * (double)c[0] + (double)c[1] * (double)c[1];
        }
        out[t] = p;
    }
}

// per-block (SBLK amplitudes) probability mass, for the two-level inverse-CDF sampler
constexpr int SBLK = 1024;
template <typename T>
__global__ void __launch_bounds__(256) k_block_mass(const void *__restrict__ sv, uint64_t length, double *mass) {
    __shared__ double s_w[8];
    const uint64_t b0 = (uint64_t)blockIdx.x * SBLK;
    double p = 0;
    for (int j = threadIdx.x; j < SBLK; j += 256) {
        const uint64_t i = b0 + j;
        if (i < length) {
            T c[2];
            load_elem<T, 1>(c, sv, i);
            p += (double)c[0] * (double)c[0] + (double)c[1] * (double)c[1];
        }
    }
    p = warp_sum(p);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = p;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < 8; ++w) t += s_w[w];
        mass[blockIdx.x] = t;
    }
}

// one thread per shot: walk the amplitudes of the selected block until the running mass passes
// the residual target
template <typename T>
__global__ void k_sample_in_block(const void *__restrict__ sv, uint64_t length, int64_t shots,
                                  const uint64_t *__restrict__ block_of, const double *__restrict__ residual,
                                  uint64_t *__restrict__ index_out) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= shots) return;
    const uint64_t b0 = block_of[s] * SBLK;
    const uint64_t end = b0 + SBLK < length ? b0 + SBLK : length;
    const double target = residual[s];
    double run = 0;
    uint64_t pick = end - 1;
    uint64_t last_nonzero = b0;
    bool found = false;
    for (uint64_t i = b0; i < end; ++i) {
        T c[2];
        load_elem<T, 1>(c, sv, i);
        const double p = (double)c[0] * (double)c[0] + (double)c[1] * (double)c[1];
        run += p;
        if (p > 0) last_nonzero = i;
        if (run > target) {
            pick = i;
            found = true;
            break;
        }
    }
    index_out[s] = found ? pick : last_nonzero;
}

// CSR y = H x with LPR lanes per row; optionally accumulates conj(x[row]) * y[row]
template <typename T, typename I, int LPR>
__global__ void __launch_bounds__(256)
    k_csr(const void *__restrict__ x, void *y, const I *__restrict__ indptr, const I *__restrict__ indices,
          const double2 *__restrict__ values, int64_t n_rows, double *out) {
    const int sub = threadIdx.x % LPR;
    const int64_t rows_per_block = 256 / LPR;
    double acc_re = 0, acc_im = 0;
    for (int64_t row0 = (int64_t)blockIdx.x * rows_per_block; row0 < n_rows;
         row0 += (int64_t)gridDim.x * rows_per_block) {
        // block-uniform loop bound: every lane takes part in the shuffles below
        const int64_t row = row0 + threadIdx.x / LPR;
        const bool active = row < n_rows;
        const int64_t lo = active ? (int64_t)indptr[row] : 0, hi = active ? (int64_t)indptr[row + 1] : 0;
        double yr = 0, yi = 0;
#pragma unroll 4
        for (int64_t j = lo + sub; j < hi; j += LPR) {
            const double2 v = values[j];
            T c[2];
            load_elem<T, 1>(c, x, (uint64_t)indices[j]);
            yr += v.x * (double)c[0] - v.y * (double)c[1];
            yi += v.x * (double)c[1] + v.y * (double)c[0];
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            yr += __shfl_down_sync(0xffffffffu, yr, o, LPR);
            yi += __shfl_down_sync(0xffffffffu, yi, o, LPR);
        }
        if (sub == 0 && active) {
            if (y != nullptr) {
                T o2[2] = {(T)yr, (T)yi};
                store_elem<T, 1>(y, (uint64_t)row, o2);
            }
            if (out != nullptr) {
                T c[2];
                load_elem<T, 1>(c, x, (uint64_t)row);
                acc_re += (double)c[0] * yr + (double)c[1] * yi;
                acc_im += (double)c[0] * yi - (double)c[1] * yr;
            }
        }
    }
    if (out != nullptr) block_accumulate<256>(acc_re, acc_im, out);
}

// Row block of a sharded CSR product: the vector x is spread over shards (peer-mapped device pointers), column c
// lives in shard c >> n_local at offset c & mask; the gathers travel over NVLink as plain loads.
template <typename T, typename I, int LPR>
__global__ void __launch_bounds__(256)
    k_csr_sharded(void *const *__restrict__ x_shards, int n_local, const void *__restrict__ x_local, void *y,
                  const I *__restrict__ indptr, const I *__restrict__ indices, const double2 *__restrict__ values,
                  int64_t n_rows, double *out) {
    const int sub = threadIdx.x % LPR;
    const int64_t rows_per_block = 256 / LPR;
    const uint64_t mask = (1ull << n_local) - 1ull;
    double acc_re = 0, acc_im = 0;
    for (int64_t row0 = (int64_t)blockIdx.x * rows_per_block; row0 < n_rows;
         row0 += (int64_t)gridDim.x * rows_per_block) {
        const int64_t row = row0 + threadIdx.x / LPR;
        const bool active = row < n_rows;
        const int64_t lo = active ? (int64_t)indptr[row] : 0, hi = active ? (int64_t)indptr[row + 1] : 0;
        double yr = 0, yi = 0;
        for (int64_t j = lo + sub; j < hi; j += LPR) {
            const double2 v = values[j];
            const uint64_t c = (uint64_t)indices[j];
            T a[2];
            load_elem<T, 1>(a, x_shards[c >> n_local], c & mask);
            yr += v.x * (double)a[0] - v.y * (double)a[1];
            yi += v.x * (double)a[1] + v.y * (double)a[0];
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
            yr += __shfl_down_sync(0xffffffffu, yr, o, LPR);
            yi += __shfl_down_sync(0xffffffffu, yi, o, LPR);
        }
        if (sub == 0 && active) {
            if (y != nullptr) {
                T o2[2] = {(T)yr, (T)yi};
                store_elem<T, 1>(y, (uint64_t)row, o2);
            }
            if (out != nullptr) {
                T c[2];
                load_elem<T, 1>(c, x_local, (uint64_t)row);
                acc_re += (double)c[0] * yr + (double)c[1] * yi;
                acc_im += (double)c[0] * yi - (double)c[1] * yr;
            }
        }
    }
    if (out != nullptr) block_accumulate<256>(acc_re, acc_im, out);
}

unsigned red_grid(uint64_t items) {
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((items + RNT - 1) / RNT, RGRID));
}

template <typename T, int K, int V>
void bra_dense_t(State &sv, const void *bra, const void *ket, const LoweredGate &g, double *out) {
    const int shift = V == 2 ? 1 : 0;
    QSV_CHECK((int)g.holes.size() <= MAX_HOLES, "too many control/target wires for one operator");
    Holes holes = make_holes(g.holes.data(), (int)g.holes.size(), shift);
    const uint64_t n_groups = 1ull << (sv.n - shift - (int)g.holes.size());
    Offs<K> offs;
    for (int j = 0; j < (1 << K); ++j) offs.v[j] = g.offs[j] >> shift;
    MatP<T, K> m;
    for (int j = 0; j < (1 << (2 * K)); ++j) {
        m.re[j] = (T)g.mat[j].real();
        m.im[j] = (T)g.mat[j].imag();
    }
    k_bra_dense_ket<T, K, V><<<red_grid(n_groups), RNT, 0, sv.stream>>>(bra, ket, n_groups, holes,
                                                                       g.ctrl_mask >> shift, offs, m, out);
    QSV_CUDA(cudaGetLastError());
}

template <typename T, int V>
void bra_dense_k(State &sv, const void *bra, const void *ket, const LoweredGate &g, double *out) {
    switch (g.k) {
    case 1: bra_dense_t<T, 1, V>(sv, bra, ket, g, out); break;
    case 2: bra_dense_t<T, 2, V>(sv, bra, ket, g, out); break;
    case 3: bra_dense_t<T, 3, V>(sv, bra, ket, g, out); break;
    case 4: bra_dense_t<T, 4, V>(sv, bra, ket, g, out); break;
    default: fail("internal: dense reduction kernel supports 1..4 targets");
    }
}

template <typename T>
void bra_dense_large(State &sv, const void *bra, const void *ket, const LoweredGate &g, double *out) {
    const int k = g.k;
    QSV_CHECK(k <= 10, "dense observables on more than 10 wires are not supported");
    const size_t D = 1ull << k;
    const size_t bytes = D * sizeof(uint64_t) + D * D * sizeof(double2);
    char *scr = (char *)sv.scratch_buffer(bytes);
    QSV_CUDA(cudaMemcpyAsync(scr, g.offs.data(), D * sizeof(uint64_t), cudaMemcpyHostToDevice, sv.stream));
    QSV_CUDA(cudaMemcpyAsync(scr + D * sizeof(uint64_t), g.mat.data(), D * D * sizeof(double2),
                             cudaMemcpyHostToDevice, sv.stream));
    Holes holes = make_holes(g.holes.data(), (int)g.holes.size(), 0);
    const uint64_t n_groups = 1ull << (sv.n - (int)g.holes.size());
    k_bra_dense_ket_large<T><<<(unsigned)std::min<uint64_t>(n_groups, RGRID), 256, 2 * D * sizeof(double2),
                               sv.stream>>>(bra, ket, k, n_groups, holes, g.ctrl_mask, (const uint64_t *)scr,
                                            (const double2 *)(scr + D * sizeof(uint64_t)), out);
    QSV_CUDA(cudaGetLastError());
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

template <typename T, int V>
void bra_diag_t(State &sv, const void *bra, const void *ket, const LoweredGate &g, double *out) {
    const int shift = V == 2 ? 1 : 0;
    std::vector<int> cpos;
    for (int b = 0; b < 64; ++b)
        if (g.ctrl_mask >> b & 1) cpos.push_back(b);
    QSV_CHECK((int)cpos.size() <= MAX_HOLES, "too many control wires for one operator");
    Holes holes = make_holes(cpos.data(), (int)cpos.size(), shift);
    const uint64_t n_items = 1ull << (sv.n - shift - (int)cpos.size());
    DiagP<double> d;
    d.k = 0;
    d.parity = g.kind == LoweredGate::PARITY;
    d.zmask = g.zmask;
    for (int i = 0; i < 4; ++i) d.tbits[i] = 0;
    for (int i = 0; i < 16; ++i) d.re[i] = d.im[i] = 0.0;
    if (g.kind == LoweredGate::NOP) {
        d.re[0] = 1.0;  // identity: plain inner product
    } else {
        if (!d.parity) {
            QSV_CHECK(g.k <= 4, "internal: diagonal table limited to 4 target bits");
            d.k = g.k;
            for (int i = 0; i < g.k; ++i) d.tbits[i] = (unsigned char)g.tgt_bits[i];
        }
        for (size_t i = 0; i < g.mat.size(); ++i) {
            d.re[i] = g.mat[i].real();
            d.im[i] = g.mat[i].imag();
        }
    }
    k_bra_diag_ket<T, V><<<red_grid(n_items), RNT, 0, sv.stream>>>(bra, ket, n_items, holes,
                                                                  g.ctrl_mask >> shift, d, out);
    QSV_CUDA(cudaGetLastError());
}

}  // namespace

void launch_bra_op_ket(State &sv, const void *bra, const void *ket, const LoweredGate &op, double *out_dev,
                       int slot) {
    sv.use();
    double *out = out_dev + 2 * (size_t)slot;
    const bool f32 = sv.dtype == QSV_C64;
    sv.stat_launches += 1;
    if (op.kind == LoweredGate::DENSE) {
        if (op.k > 4) {
            if (f32)
                bra_dense_large<float>(sv, bra, ket, op, out);
            else
                bra_dense_large<double>(sv, bra, ket, op, out);
            return;
        }
        const bool bit0 = !op.holes.empty() && op.holes[0] == 0;
        if (!f32)
            bra_dense_k<double, 1>(sv, bra, ket, op, out);
        else if (bit0)
            bra_dense_k<float, 1>(sv, bra, ket, op, out);
        else
            bra_dense_k<float, 2>(sv, bra, ket, op, out);
        return;
    }
    const bool bit0 = op.ctrl_mask & 1ull;
    if (!f32)
        bra_diag_t<double, 1>(sv, bra, ket, op, out);
    else if (bit0 || sv.n < 1)
        bra_diag_t<float, 1>(sv, bra, ket, op, out);
    else
        bra_diag_t<float, 2>(sv, bra, ket, op, out);
}

void launch_bra_pauli_ket(State &sv, const void *bra, const void *ket, uint64_t xmask, uint64_t zmask, int ny,
                          double *out_dev, int slot) {
    sv.use();
    sv.stat_launches += 1;
    double *out = out_dev + 2 * (size_t)slot;
    if (sv.dtype == QSV_C128) {
        k_bra_pauli_ket<double, 1><<<red_grid(sv.length()), RNT, 0, sv.stream>>>(bra, ket, sv.length(), xmask,
                                                                                zmask, ny, out);
    } else if (sv.n >= 1) {
        k_bra_pauli_ket<float, 2><<<red_grid(sv.length() / 2), RNT, 0, sv.stream>>>(bra, ket, sv.length() / 2,
                                                                                   xmask, zmask, ny, out);
    } else {
        k_bra_pauli_ket<float, 1><<<1, RNT, 0, sv.stream>>>(bra, ket, 1, xmask, zmask, ny, out);
    }
    QSV_CUDA(cudaGetLastError());
}

void launch_pauli_sum_apply(State &sv, const void *in, void *out, int n_terms, const uint64_t *xmasks,
                            const uint64_t *zmasks, const cplx *coeffs, bool accumulate) {
    sv.use();
    QSV_CHECK(in != out, "internal: pauli-sum apply is out of place");
    static const bool tiled_ok = [] {
        const char *v = std::getenv("QSV_PAULI_SUM_TILED");
        return !(v && std::atoi(v) == 0);
    }();
    if (tiled_ok && launch_pauli_sum_apply_tiled(sv, in, out, n_terms, xmasks, zmasks, coeffs, accumulate)) {
        QSV_CUDA(cudaStreamSynchronize(sv.stream));
        return;
    }
    sv.stat_launches += 1;
    const size_t nt = (size_t)n_terms;
    char *scr = (char *)sv.scratch_buffer(nt * (2 * sizeof(uint64_t) + sizeof(double2)));
    QSV_CUDA(cudaMemcpyAsync(scr, xmasks, nt * 8, cudaMemcpyHostToDevice, sv.stream));
    QSV_CUDA(cudaMemcpyAsync(scr + nt * 8, zmasks, nt * 8, cudaMemcpyHostToDevice, sv.stream));
    QSV_CUDA(cudaMemcpyAsync(scr + nt * 16, coeffs, nt * 16, cudaMemcpyHostToDevice, sv.stream));
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((sv.length() + 255) / 256, NUM_SMS * 16));
    if (sv.dtype == QSV_C128)
        k_pauli_sum_apply<double><<<grid, 256, 0, sv.stream>>>(in, out, sv.length(), n_terms, (const uint64_t *)scr,
                                                              (const uint64_t *)(scr + nt * 8),
                                                              (const double2 *)(scr + nt * 16), accumulate ? 1 : 0);
    else
        k_pauli_sum_apply<float><<<grid, 256, 0, sv.stream>>>(in, out, sv.length(), n_terms, (const uint64_t *)scr,
                                                             (const uint64_t *)(scr + nt * 8),
                                                             (const double2 *)(scr + nt * 16), accumulate ? 1 : 0);
    QSV_CUDA(cudaGetLastError());
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

void reduction_zero(State &sv, double *dev, size_t count) {
    sv.use();
    QSV_CUDA(cudaMemsetAsync(dev, 0, count * sizeof(double), sv.stream));
}

void reduction_read(State &sv, const double *dev, double *host, size_t count) {
    sv.use();
    QSV_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(double), cudaMemcpyDeviceToHost, sv.stream));
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

void launch_probs(State &sv, const std::vector<int> &bits, double *out_host) {
    sv.use();
    const int k = (int)bits.size();
    QSV_CHECK(k >= 1 && k <= sv.n, "probability needs between 1 and n wires");
    sv.stat_launches += 1;
    const size_t nb = 1ull << k;
    // device buffers: bins + bit list
    char *scr = (char *)sv.scratch_buffer(nb * sizeof(double) + 64);
    double *bins = (double *)scr;
    unsigned char *dbits = (unsigned char *)(scr + nb * sizeof(double));
    unsigned char hb[64] = {0};
    for (int i = 0; i < k; ++i) hb[i] = (unsigned char)bits[i];
    QSV_CUDA(cudaMemcpyAsync(dbits, hb, 64, cudaMemcpyHostToDevice, sv.stream));
    const bool f32 = sv.dtype == QSV_C64;
    if (sv.n >= 10) {
        // rows of 32 amplitudes; measured bits >= 5 select the group of rows, measured bits < 5 the lane's bin
        ProbsPlan P;
        memset(&P, 0, sizeof(P));
        for (int b = 0; b < 5; ++b) P.lane_out[b] = 255;
        std::vector<std::pair<int, int>> hi;  // (row-index position, output bit)
        for (int q = 0; q < k; ++q) {
            QSV_CHECK(bits[q] >= 0 && bits[q] < sv.n, "probability wire out of range");
            if (bits[q] < 5)
                P.lane_out[bits[q]] = (unsigned char)q;
            else
                hi.push_back({bits[q] - 5, q});
        }
        std::sort(hi.begin(), hi.end());
        QSV_CHECK((int)hi.size() <= MAX_HOLES, "probability on more than 45 wires is not supported");
        int pos[MAX_HOLES];
        P.n_hi = (int)hi.size();
        for (int j = 0; j < P.n_hi; ++j) {
            pos[j] = hi[j].first;
            P.hi_pos[j] = (unsigned char)hi[j].first;
            P.hi_out[j] = (unsigned char)hi[j].second;
        }
        P.row_holes = make_holes(pos, P.n_hi, 0);
        for (int b = 0; b < 5; ++b)
            if (P.lane_out[b] == 255) P.lane_sum_mask |= 1u << b;
        P.rows_log2 = sv.n - 5 - P.n_hi;
        // enough chunks to fill the GPU, at least 64 rows (32 KiB) each when the groups are that long
        const int want_items_log2 = 14;  // ~16k warp work items
        P.chunk_log2 = std::max(std::min(P.rows_log2, 6), P.rows_log2 - std::max(0, want_items_log2 - P.n_hi));
        P.chunk_log2 = std::min(P.chunk_log2, P.rows_log2);
        P.use_atomics = P.chunk_log2 < P.rows_log2 ? 1 : 0;
        const uint64_t n_items = 1ull << (P.n_hi + P.rows_log2 - P.chunk_log2);
        if (P.use_atomics) QSV_CUDA(cudaMemsetAsync(bins, 0, nb * sizeof(double), sv.stream));
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_items + 7) / 8, NUM_SMS * 8));
        if (f32)
            k_probs_rows<float><<<grid, 256, 0, sv.stream>>>(sv.data, bins, n_items, P);
        else
            k_probs_rows<double><<<grid, 256, 0, sv.stream>>>(sv.data, bins, n_items, P);
    } else if (k <= 11) {
        QSV_CUDA(cudaMemsetAsync(bins, 0, nb * sizeof(double), sv.stream));
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((sv.length() + 255) / 256, NUM_SMS * 4));
        if (f32)
            k_probs_small<float><<<grid, 256, nb * sizeof(double), sv.stream>>>(sv.data, sv.length(), k, dbits, bins);
        else
            k_probs_small<double><<<grid, 256, nb * sizeof(double), sv.stream>>>(sv.data, sv.length(), k, dbits, bins);
    } else {
        // holes = measured bits (sorted) so that the loop counter enumerates the traced-out bits
        std::vector<int> sorted(bits);
        std::sort(sorted.begin(), sorted.end());
        QSV_CHECK(k <= MAX_HOLES, "probability on more than 40 wires is not supported");
        Holes kept = make_holes(sorted.data(), k, 0);
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nb + 255) / 256, NUM_SMS * 16));
        if (f32)
            k_probs_large<float><<<grid, 256, 0, sv.stream>>>(sv.data, sv.n, k, dbits, kept, bins);
        else
            k_probs_large<double><<<grid, 256, 0, sv.stream>>>(sv.data, sv.n, k, dbits, kept, bins);
    }
    QSV_CUDA(cudaGetLastError());
    QSV_CUDA(cudaMemcpyAsync(out_host, bins, nb * sizeof(double), cudaMemcpyDeviceToHost, sv.stream));
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

// block masses of the vector on the host (sequential FP64 prefix, deterministic)
static void block_cdf(State &sv, std::vector<double> &cdf, double *&mass_dev, uint64_t &nblk, size_t extra_bytes = 0) {
    const uint64_t len = sv.length();
    nblk = (len + SBLK - 1) / SBLK;
    mass_dev = (double *)sv.scratch_buffer(nblk * sizeof(double) + extra_bytes);  // extra area follows the masses
    if (sv.dtype == QSV_C64)
        k_block_mass<float><<<(unsigned)nblk, 256, 0, sv.stream>>>(sv.data, len, mass_dev);
    else
        k_block_mass<double><<<(unsigned)nblk, 256, 0, sv.stream>>>(sv.data, len, mass_dev);
    QSV_CUDA(cudaGetLastError());
    std::vector<double> h_mass(nblk);
    QSV_CUDA(cudaMemcpyAsync(h_mass.data(), mass_dev, nblk * sizeof(double), cudaMemcpyDeviceToHost, sv.stream));
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
    cdf.assign(nblk + 1, 0.0);
    for (uint64_t b = 0; b < nblk; ++b) cdf[b + 1] = cdf[b] + h_mass[b];
}

double state_mass(State &sv) {
    sv.use();
    std::vector<double> cdf;
    double *mass_dev;
    uint64_t nblk;
    block_cdf(sv, cdf, mass_dev, nblk);
    return cdf[nblk];
}

// two-level inverse CDF: index_host[s] = first index whose cumulative |amp|^2 exceeds the target mass
// (targets[s] * total when the targets are uniform numbers, targets[s] itself when they are masses)
void launch_sample_indices(State &sv, const double *targets, int64_t shots, uint64_t *index_host, bool targets_are_mass) {
    sv.use();
    if (shots <= 0) return;
    sv.stat_launches += 2;
    const uint64_t len = sv.length();
    std::vector<double> cdf;
    double *mass_dev;
    uint64_t nblk;
    block_cdf(sv, cdf, mass_dev, nblk, (size_t)shots * (sizeof(uint64_t) * 2 + sizeof(double)));
    const double total = cdf[nblk];
    std::vector<uint64_t> h_block(shots);
    std::vector<double> h_resid(shots);
    for (int64_t s = 0; s < shots; ++s) {
        const double target = targets_are_mass ? targets[s] : targets[s] * total;
        // first block whose cumulative mass exceeds the target
        uint64_t b = (uint64_t)(std::upper_bound(cdf.begin() + 1, cdf.end(), target) - (cdf.begin() + 1));
        if (b >= nblk) b = nblk - 1;
        h_block[s] = b;
        h_resid[s] = target - cdf[b];
    }
    uint64_t *d_block = (uint64_t *)(mass_dev + nblk);
    double *d_resid = (double *)(d_block + shots);
    uint64_t *d_index = (uint64_t *)(d_resid + shots);
    QSV_CUDA(cudaMemcpyAsync(d_block, h_block.data(), shots * sizeof(uint64_t), cudaMemcpyHostToDevice, sv.stream));
    QSV_CUDA(cudaMemcpyAsync(d_resid, h_resid.data(), shots * sizeof(double), cudaMemcpyHostToDevice, sv.stream));
    const unsigned grid = (unsigned)((shots + 127) / 128);
    if (sv.dtype == QSV_C64)
        k_sample_in_block<float><<<grid, 128, 0, sv.stream>>>(sv.data, len, shots, d_block, d_resid, d_index);
    else
        k_sample_in_block<double><<<grid, 128, 0, sv.stream>>>(sv.data, len, shots, d_block, d_resid, d_index);
    QSV_CUDA(cudaGetLastError());
    QSV_CUDA(cudaMemcpyAsync(index_host, d_index, shots * sizeof(uint64_t), cudaMemcpyDeviceToHost, sv.stream));
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

void launch_sample(State &sv, const double *uniforms, int64_t shots, uint64_t *out_host) {
    if (shots <= 0) return;
    std::vector<uint64_t> h_index(shots);
    launch_sample_indices(sv, uniforms, shots, h_index.data(), false);
    const int n = sv.n;
    for (int64_t s = 0; s < shots; ++s)
        for (int w = 0; w < n; ++w) out_host[s * n + w] = (h_index[s] >> (n - 1 - w)) & 1ull;
}

void launch_csr(State &sv, const void *x, void *y, const void *dev_indptr, const void *dev_indices,
                const void *dev_values, int64_t n_rows, int64_t nnz, int index_bytes, double *out_dev, int slot) {
    sv.use();
    sv.stat_launches += 1;
    QSV_CHECK(x != y, "internal: CSR product is out of place");
    double *out = out_dev ? out_dev + 2 * (size_t)slot : nullptr;
    const double avg = n_rows > 0 ? (double)nnz / (double)n_rows : 0.0;
    // lanes per row: about a quarter of the row length (measured on B200, 22-qubit config-4 matrix with 31 non-zeros per
    // row: 8 lanes 1.09 ms, 16 lanes 1.23 ms, 32 lanes 1.72 ms) -- several rows in flight per warp, a shuffle tree of 3 levels
    int lpr = 1;
    while (lpr < 32 && lpr * 4 < avg) lpr *= 2;
    // QSV_CSR_LPR: lanes per row (A/B; fewer lanes = more rows in flight per warp and shorter shuffle trees)
    static const int lpr_env = [] {
        const char *v = std::getenv("QSV_CSR_LPR");
        return v ? std::atoi(v) : 0;
    }();
    if (lpr_env == 1 || lpr_env == 2 || lpr_env == 4 || lpr_env == 8 || lpr_env == 16 || lpr_env == 32) lpr = lpr_env;
    const bool f32 = sv.dtype == QSV_C64;
    const bool i32 = index_bytes == 4;
    QSV_CHECK(index_bytes == 4 || index_bytes == 8, "CSR index width must be 4 or 8 bytes");
    const double2 *vals = (const double2 *)dev_values;
#define QSV_CSR_LAUNCH(T, I, L)                                                                              \
    k_csr<T, I, L><<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((n_rows + (256 / L) - 1) / (256 / L),  \
                                                                      NUM_SMS * 16)),                        \
                     256, 0, sv.stream>>>(x, y, (const I *)dev_indptr, (const I *)dev_indices, vals, n_rows, out)
#define QSV_CSR_L(T, I)                                                                                      \
    switch (lpr) {                                                                                           \
    case 1: QSV_CSR_LAUNCH(T, I, 1); break;                                                                  \
    case 2: QSV_CSR_LAUNCH(T, I, 2); break;                                                                  \
    case 4: QSV_CSR_LAUNCH(T, I, 4); break;                                                                  \
    case 8: QSV_CSR_LAUNCH(T, I, 8); break;                                                                  \
    case 16: QSV_CSR_LAUNCH(T, I, 16); break;                                                                \
    default: QSV_CSR_LAUNCH(T, I, 32); break;                                                                \
    }
    if (f32 && i32) {
        QSV_CSR_L(float, int32_t)
    } else if (f32) {
        QSV_CSR_L(float, int64_t)
    } else if (i32) {
        QSV_CSR_L(double, int32_t)
    } else {
        QSV_CSR_L(double, int64_t)
    }
#undef QSV_CSR_L
#undef QSV_CSR_LAUNCH
    QSV_CUDA(cudaGetLastError());
}

void launch_csr_sharded(State &sv, void *const *x_shards_dev, int n_local, const void *x_local, void *y,
                        const void *dev_indptr, const void *dev_indices, const void *dev_values, int64_t n_rows,
                        int64_t nnz, int index_bytes, double *out_dev, int slot) {
    sv.use();
    sv.stat_launches += 1;
    QSV_CHECK(index_bytes == 4 || index_bytes == 8, "CSR index width must be 4 or 8 bytes");
    double *out = out_dev ? out_dev + 2 * (size_t)slot : nullptr;
    const double avg = n_rows > 0 ? (double)nnz / (double)n_rows : 0.0;
    const int lpr = avg > 12 ? 32 : (avg > 3 ? 8 : 1);
    const double2 *vals = (const double2 *)dev_values;
    const bool f32 = sv.dtype == QSV_C64, i32 = index_bytes == 4;
#define QSV_CSRS(T, I, L)                                                                                              \
    k_csr_sharded<T, I, L><<<(unsigned)std::max<int64_t>(1, std::min<int64_t>((n_rows + (256 / L) - 1) / (256 / L),     \
                                                                              NUM_SMS * 16)),                           \
                             256, 0, sv.stream>>>(x_shards_dev, n_local, x_local, y, (const I *)dev_indptr,             \
                                                  (const I *)dev_indices, vals, n_rows, out)
#define QSV_CSRS_L(T, I)                                                                                               \
    if (lpr == 32) {                                                                                                   \
        QSV_CSRS(T, I, 32);                                                                                            \
    } else if (lpr == 8) {                                                                                             \
        QSV_CSRS(T, I, 8);                                                                                             \
    } else {                                                                                                           \
        QSV_CSRS(T, I, 1);                                                                                             \
    }
    if (f32 && i32) {
        QSV_CSRS_L(float, int32_t)
    } else if (f32) {
        QSV_CSRS_L(float, int64_t)
    } else if (i32) {
        QSV_CSRS_L(double, int32_t)
    } else {
        QSV_CSRS_L(double, int64_t)
    }
#undef QSV_CSRS_L
#undef QSV_CSRS
    QSV_CUDA(cudaGetLastError());
}

}  // namespace qsv
