// Gate-application kernels, one HBM sweep per gate (the un-fused path).
//
// Replaces custatevecApplyMatrix / custatevecApplyPauliRotation at
//   simulator/StateVectorCudaManaged.hpp:1339-1386, :1400-1472, :1484-1568   (reference).
//
// Design (sm_100a, HBM-bound):
//   * a gate only touches the amplitudes it changes: controls (explicit or implicit, e.g. the |1>
//     half of PhaseShift) are an index predicate realised by "hole insertion", so the sweep moves
//     2 * B * N / 2^c bytes;
//   * every thread owns complete groups of 2^k partner amplitudes in registers and moves them with
//     128-bit accesses (double2 for complex128; float4 = two neighbouring complex64 amplitudes
//     whenever index bit 0 is not touched), consecutive threads -> consecutive addresses, so every
//     warp request is a run of whole 32-byte sectors for every target bit, high or low;
//   * U independent groups per thread are loaded before any arithmetic to keep >= 8 128-bit loads
//     in flight per thread (Little's law: ~35 KB in flight per SM saturates HBM3e);
//   * the matrix travels in the kernel parameter bank (constant cache, uniform operands).
#include <algorithm>
#include <type_traits>

#include "device_utils.cuh"
#include "qsv_internal.h"

namespace qsv {

namespace {

template <int K> struct Offs {
    uint64_t v[1 << K];
};
template <typename T, int K> struct MatP {
    T re[1 << (2 * K)];
    T im[1 << (2 * K)];
};

template <typename T, int K, int V, int U, int NT>
__global__ void __launch_bounds__(NT)
    k_apply_dense(void *single, void *const *table, uint64_t n_groups, Holes holes, uint64_t ctrl,
                  Offs<K> offs, MatP<T, K> m) {
    constexpr int D = 1 << K;
    void *sv = table ? table[blockIdx.y] : single;
    const uint64_t g0 = (uint64_t)blockIdx.x * (uint64_t)(NT * U) + threadIdx.x;
    T x[U][D][2 * V];
    uint64_t base[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t g = g0 + (uint64_t)u * NT;
        if (g < n_groups) {
            base[u] = expand_index(g, holes) | ctrl;
#pragma unroll
            for (int d = 0; d < D; ++d) load_elem<T, V>(x[u][d], sv, base[u] + offs.v[d]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t g = g0 + (uint64_t)u * NT;
        if (g < n_groups) {
#pragma unroll
            for (int r = 0; r < D; ++r) {
                T y[2 * V];
#pragma unroll
                for (int a = 0; a < 2 * V; ++a) y[a] = T(0);
#pragma unroll
                for (int c = 0; c < D; ++c) {
                    const T mr = m.re[r * D + c], mi = m.im[r * D + c];
#pragma unroll
                    for (int a = 0; a < V; ++a) {
                        y[2 * a] = fma(mr, x[u][c][2 * a], y[2 * a]);
                        y[2 * a] = fma(-mi, x[u][c][2 * a + 1], y[2 * a]);
                        y[2 * a + 1] = fma(mr, x[u][c][2 * a + 1], y[2 * a + 1]);
                        y[2 * a + 1] = fma(mi, x[u][c][2 * a], y[2 * a + 1]);
                    }
                }
                store_elem<T, V>(sv, base[u] + offs.v[r], y);
            }
        }
    }
}

// 2x2 gate on index bit 0: both partners sit in one 32-byte (complex128) / 16-byte (complex64) element,
// moved with a single 256-bit (LDG.E.ENL2.256) / 128-bit access per thread and element.
__device__ __forceinline__ void load_pair(double (&c)[4], const void *base, uint64_t e) {
    const double *p = reinterpret_cast<const double *>(base) + 4 * e;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c[0]), "=d"(c[1]), "=d"(c[2]), "=d"(c[3]) : "l"(p));
}
__device__ __forceinline__ void store_pair(void *base, uint64_t e, const double (&c)[4]) {
    double *p = reinterpret_cast<double *>(base) + 4 * e;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]) : "memory");
}
__device__ __forceinline__ void load_pair(float (&c)[4], const void *base, uint64_t e) {
    const float4 v = reinterpret_cast<const float4 *>(base)[e];
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
}
__device__ __forceinline__ void store_pair(void *base, uint64_t e, const float (&c)[4]) {
    reinterpret_cast<float4 *>(base)[e] = make_float4(c[0], c[1], c[2], c[3]);
}

template <typename T, int U, int NT>
__global__ void __launch_bounds__(NT)
    k_apply_dense_bit0(void *single, void *const *table, uint64_t n_elems, Holes holes, uint64_t ctrl, MatP<T, 1> m) {
    void *sv = table ? table[blockIdx.y] : single;
    const uint64_t g0 = (uint64_t)blockIdx.x * (uint64_t)(NT * U) + threadIdx.x;
    T x[U][4];
    uint64_t e[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t g = g0 + (uint64_t)u * NT;
        if (g < n_elems) {
            e[u] = expand_index(g, holes) | ctrl;
            load_pair(x[u], sv, e[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t g = g0 + (uint64_t)u * NT;
        if (g < n_elems) {
            T y[4];
            y[0] = m.re[0] * x[u][0] - m.im[0] * x[u][1] + m.re[1] * x[u][2] - m.im[1] * x[u][3];
            y[1] = m.re[0] * x[u][1] + m.im[0] * x[u][0] + m.re[1] * x[u][3] + m.im[1] * x[u][2];
            y[2] = m.re[2] * x[u][0] - m.im[2] * x[u][1] + m.re[3] * x[u][2] - m.im[3] * x[u][3];
            y[3] = m.re[2] * x[u][1] + m.im[2] * x[u][0] + m.re[3] * x[u][3] + m.im[3] * x[u][2];
            store_pair(sv, e[u], y);
        }
    }
}

// Generic dense block for k > 4 targets: one CTA per group, amplitudes staged in shared memory,
// one warp per output row, matrix streamed from L2.  Correctness path for large QubitUnitary;
// such blocks are FP-bound, not HBM-bound.
template <typename T>
__global__ void __launch_bounds__(256)
    k_apply_dense_large(void *single, void *const *table, int k, uint64_t n_groups, Holes holes,
                        uint64_t ctrl, const uint64_t *__restrict__ offs, const double2 *__restrict__ mat) {
    extern __shared__ double2 s_x[];  // x[D] then y[D], always double precision in smem
    const int D = 1 << k;
    double2 *s_y = s_x + D;
    void *sv = table ? table[blockIdx.y] : single;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint64_t g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const uint64_t base = expand_index(g, holes) | ctrl;
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            T c[2];
            load_elem<T, 1>(c, sv, base + offs[d]);
            s_x[d] = make_double2((double)c[0], (double)c[1]);
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
            re = warp_sum(re);
            im = warp_sum(im);
            if (lane == 0) s_y[r] = make_double2(re, im);
        }
        __syncthreads();
        for (int d = threadIdx.x; d < D; d += blockDim.x) {
            T c[2] = {(T)s_y[d].x, (T)s_y[d].y};
            store_elem<T, 1>(sv, base + offs[d], c);
        }
        __syncthreads();
    }
}

template <typename T> struct DiagP {
    int k;        // number of table bits (0..4), ignored for parity
    int parity;   // 1: index = popc(i & zmask) & 1
    unsigned char tbits[4];  // tbits[0] = MSB of the table index
    uint64_t zmask;
    T re[16];
    T im[16];
};

template <typename T, int V, int U, int NT>
__global__ void __launch_bounds__(NT)
    k_apply_diag(void *single, void *const *table, uint64_t n_items, Holes holes, uint64_t ctrl,
                 DiagP<T> d) {
    void *sv = table ? table[blockIdx.y] : single;
    const uint64_t i0 = (uint64_t)blockIdx.x * (uint64_t)(NT * U) + threadIdx.x;
    T x[U][2 * V];
    uint64_t idx[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t o = i0 + (uint64_t)u * NT;
        if (o < n_items) {
            idx[u] = expand_index(o, holes) | ctrl;
            load_elem<T, V>(x[u], sv, idx[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint64_t o = i0 + (uint64_t)u * NT;
        if (o < n_items) {
            T y[2 * V];
#pragma unroll
            for (int a = 0; a < V; ++a) {
                const uint64_t i = idx[u] * V + a;
                int t = 0;
                if (d.parity) {
                    t = __popcll(i & d.zmask) & 1;
                } else {
                    for (int b = 0; b < d.k; ++b) t = (t << 1) | (int)((i >> d.tbits[b]) & 1ull);
                }
                const T pr = d.re[t], pi = d.im[t];
                y[2 * a] = pr * x[u][2 * a] - pi * x[u][2 * a + 1];
                y[2 * a + 1] = pr * x[u][2 * a + 1] + pi * x[u][2 * a];
            }
            store_elem<T, V>(sv, idx[u], y);
        }
    }
}

template <typename T, int V, int NT>
__global__ void __launch_bounds__(NT) k_fill_basis(void *sv, uint64_t n_elems, uint64_t index) {
    const uint64_t stride = (uint64_t)gridDim.x * NT;
    for (uint64_t e = (uint64_t)blockIdx.x * NT + threadIdx.x; e < n_elems; e += stride) {
        T c[2 * V];
#pragma unroll
        for (int a = 0; a < V; ++a) {
            c[2 * a] = (e * V + a == index) ? T(1) : T(0);
            c[2 * a + 1] = T(0);
        }
        store_elem<T, V>(sv, e, c);
    }
}

template <typename T>
__global__ void k_scatter(void *sv, const int64_t *__restrict__ idx, const void *__restrict__ vals,
                          size_t count, uint64_t length) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        T c[2];
        load_elem<T, 1>(c, vals, i);
        const uint64_t j = (uint64_t)idx[i];
        if (j < length) store_elem<T, 1>(sv, j, c);
    }
}

template <typename T, int V, int NT>
__global__ void __launch_bounds__(NT)
    k_axpy(T ar, T ai, const void *__restrict__ x, void *y, uint64_t n_elems) {
    const uint64_t stride = (uint64_t)gridDim.x * NT;
    for (uint64_t e = (uint64_t)blockIdx.x * NT + threadIdx.x; e < n_elems; e += stride) {
        T a[2 * V], b[2 * V];
        load_elem<T, V>(a, x, e);
        load_elem<T, V>(b, y, e);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            b[2 * v] += ar * a[2 * v] - ai * a[2 * v + 1];
            b[2 * v + 1] += ar * a[2 * v + 1] + ai * a[2 * v];
        }
        store_elem<T, V>(y, e, b);
    }
}

unsigned grid_for(uint64_t items, uint64_t per_block) {
    uint64_t g = (items + per_block - 1) / per_block;
    QSV_CHECK(g <= 0x7fffffffull, "state too large for a single launch");
    return (unsigned)std::max<uint64_t>(g, 1);
}

int env_shape(const char *name, int dflt) {
    const char *e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// launch shape of the single-target dense kernel; read on every launch so that an A/B can switch it inside one process
constexpr int DENSE1_DEFAULT_SHAPE = 5;  // U=8, NT=128: +14 % over U=4, NT=256 (profiles/r1_ab_dense1.txt)
int dense1_shape() { return env_shape("QSV_DENSE1_SHAPE", DENSE1_DEFAULT_SHAPE); }

template <typename T, int K, int V>
void launch_dense_t(State &sv, const LoweredGate &g, void *const *table, int n_vecs) {
    const int shift = V == 2 ? 1 : 0;
    const int n_eff = sv.n - shift;
    QSV_CHECK((int)g.holes.size() <= MAX_HOLES, "too many control/target wires for one gate");
    Holes holes = make_holes(g.holes.data(), (int)g.holes.size(), shift);
    const uint64_t n_groups = 1ull << (n_eff - (int)g.holes.size());
    Offs<K> offs;
    for (int j = 0; j < (1 << K); ++j) offs.v[j] = g.offs[j] >> shift;
    MatP<T, K> m;
    for (int j = 0; j < (1 << (2 * K)); ++j) {
        m.re[j] = (T)g.mat[j].real();
        m.im[j] = (T)g.mat[j].imag();
    }
    // groups per thread / threads per block by block size (register budget)
    constexpr int U = K == 0 ? 8 : (K == 1 ? 4 : (K == 2 ? 2 : 1));
    constexpr int NT = K <= 2 ? 256 : 128;
    const uint64_t ctrl = g.ctrl_mask >> shift;
    void *single = table ? nullptr : sv.data;
    auto go = [&](auto u_tag, auto nt_tag) {
        constexpr int UU = decltype(u_tag)::value, NN = decltype(nt_tag)::value;
        dim3 grid(grid_for(n_groups, (uint64_t)NN * UU), (unsigned)n_vecs);
        k_apply_dense<T, K, V, UU, NN><<<grid, NN, 0, sv.stream>>>(single, table, n_groups, holes, ctrl, offs, m);
    };
    using std::integral_constant;
    if constexpr (K == 1 && V == 2 && std::is_same_v<T, double>) {
        // A/B only (QSV_DENSE1_SHAPE=8): two neighbouring complex128 amplitudes per 256-bit access
        go(integral_constant<int, 2>{}, integral_constant<int, 256>{});
    } else if constexpr (K == 1) {
        // launch shape of the single-target kernel (QSV_DENSE1_SHAPE, tools/ab_dense1.py, profiles/r1_ab_dense1.txt)
        switch (dense1_shape()) {
        case 0: go(integral_constant<int, 4>{}, integral_constant<int, 256>{}); break;
        case 1: go(integral_constant<int, 2>{}, integral_constant<int, 256>{}); break;
        case 2: go(integral_constant<int, 8>{}, integral_constant<int, 256>{}); break;
        case 3: go(integral_constant<int, 4>{}, integral_constant<int, 128>{}); break;
        case 4: go(integral_constant<int, 2>{}, integral_constant<int, 512>{}); break;
        case 6: go(integral_constant<int, 1>{}, integral_constant<int, 512>{}); break;
        default: go(integral_constant<int, 8>{}, integral_constant<int, 128>{}); break;  // 5, and 7 / 8 after their re-routing
        }
    } else if constexpr (K == 2) {
        // QSV_DENSE2_SHAPE (tools/ab_shapes2.py): U=4, NT=128 is level with U=2, NT=256 on high bits, +12 % on bits 1 and 2
        switch (env_shape("QSV_DENSE2_SHAPE", 1)) {
        case 0: go(integral_constant<int, 2>{}, integral_constant<int, 256>{}); break;
        case 2: go(integral_constant<int, 1>{}, integral_constant<int, 256>{}); break;
        case 3: go(integral_constant<int, 2>{}, integral_constant<int, 128>{}); break;
        default: go(integral_constant<int, 4>{}, integral_constant<int, 128>{}); break;
        }
    } else {
        go(integral_constant<int, U>{}, integral_constant<int, NT>{});
    }
    QSV_CUDA(cudaGetLastError());
}

template <typename T, int V>
void launch_dense_k(State &sv, const LoweredGate &g, void *const *table, int n_vecs) {
    switch (g.k) {
    case 1: launch_dense_t<T, 1, V>(sv, g, table, n_vecs); break;
    case 2: launch_dense_t<T, 2, V>(sv, g, table, n_vecs); break;
    case 3: launch_dense_t<T, 3, V>(sv, g, table, n_vecs); break;
    case 4: launch_dense_t<T, 4, V>(sv, g, table, n_vecs); break;
    default: fail("internal: dense register kernel supports 1..4 targets");
    }
}

template <typename T>
void launch_dense_bit0(State &sv, const LoweredGate &g, void *const *table, int n_vecs) {
    // element = the (2g, 2g+1) pair; remaining holes are the control bits, shifted down by one
    std::vector<int> cpos;
    for (int h : g.holes)
        if (h != 0) cpos.push_back(h);
    QSV_CHECK((int)cpos.size() <= MAX_HOLES, "too many control wires for one gate");
    Holes holes = make_holes(cpos.data(), (int)cpos.size(), 1);
    const uint64_t n_elems = 1ull << (sv.n - 1 - (int)cpos.size());
    MatP<T, 1> m;
    for (int j = 0; j < 4; ++j) {
        m.re[j] = (T)g.mat[j].real();
        m.im[j] = (T)g.mat[j].imag();
    }
    void *single = table ? nullptr : sv.data;
    auto go = [&](auto u_tag, auto nt_tag) {
        constexpr int U = decltype(u_tag)::value, NT = decltype(nt_tag)::value;
        dim3 grid(grid_for(n_elems, (uint64_t)NT * U), (unsigned)n_vecs);
        k_apply_dense_bit0<T, U, NT><<<grid, NT, 0, sv.stream>>>(single, table, n_elems, holes, g.ctrl_mask >> 1, m);
    };
    using std::integral_constant;
    // QSV_BIT0_SHAPE (tools/ab_shapes2.py): U=8, NT=128 +7 % over U=4, NT=256
    switch (env_shape("QSV_BIT0_SHAPE", 1)) {
    case 0: go(integral_constant<int, 4>{}, integral_constant<int, 256>{}); break;
    case 2: go(integral_constant<int, 2>{}, integral_constant<int, 256>{}); break;
    case 3: go(integral_constant<int, 8>{}, integral_constant<int, 256>{}); break;
    default: go(integral_constant<int, 8>{}, integral_constant<int, 128>{}); break;
    }
    QSV_CUDA(cudaGetLastError());
}

template <typename T>
void launch_dense_large(State &sv, const LoweredGate &g, void *const *table, int n_vecs) {
    const int k = g.k;
    QSV_CHECK(k <= 10, "dense gates on more than 10 target wires are not supported");
    QSV_CHECK((int)g.holes.size() <= MAX_HOLES, "too many control/target wires for one gate");
    const size_t D = 1ull << k;
    // device copies of offsets and matrix (stream ordered, scratch is reused between calls)
    const size_t bytes = D * sizeof(uint64_t) + D * D * sizeof(double2);
    char *scr = (char *)sv.scratch_buffer(bytes);
    QSV_CUDA(cudaMemcpyAsync(scr, g.offs.data(), D * sizeof(uint64_t), cudaMemcpyHostToDevice, sv.stream));
    QSV_CUDA(cudaMemcpyAsync(scr + D * sizeof(uint64_t), g.mat.data(), D * D * sizeof(double2),
                             cudaMemcpyHostToDevice, sv.stream));
    Holes holes = make_holes(g.holes.data(), (int)g.holes.size(), 0);
    const uint64_t n_groups = 1ull << (sv.n - (int)g.holes.size());
    dim3 grid((unsigned)std::min<uint64_t>(n_groups, NUM_SMS * 8), (unsigned)n_vecs);
    k_apply_dense_large<T><<<grid, 256, 2 * D * sizeof(double2), sv.stream>>>(
        table ? nullptr : sv.data, table, k, n_groups, holes, g.ctrl_mask, (const uint64_t *)scr,
        (const double2 *)(scr + D * sizeof(uint64_t)));
    QSV_CUDA(cudaGetLastError());
    // the scratch buffer may be overwritten by the next call only after this kernel ran
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

template <typename T, int V>
void launch_diag_t(State &sv, const LoweredGate &g, void *const *table, int n_vecs) {
    const int shift = V == 2 ? 1 : 0;
    std::vector<int> cpos;
    for (int b = 0; b < 64; ++b)
        if (g.ctrl_mask >> b & 1) cpos.push_back(b);
    QSV_CHECK((int)cpos.size() <= MAX_HOLES, "too many control wires for one gate");
    Holes holes = make_holes(cpos.data(), (int)cpos.size(), shift);
    const uint64_t n_items = 1ull << (sv.n - shift - (int)cpos.size());
    DiagP<T> d;
    d.k = 0;
    d.parity = g.kind == LoweredGate::PARITY;
    d.zmask = g.zmask;
    for (int i = 0; i < 4; ++i) d.tbits[i] = 0;
    for (int i = 0; i < 16; ++i) d.re[i] = d.im[i] = T(0);
    if (!d.parity) {
        QSV_CHECK(g.k <= 4, "internal: diagonal table limited to 4 target bits");
        d.k = g.k;
        for (int i = 0; i < g.k; ++i) d.tbits[i] = (unsigned char)g.tgt_bits[i];
    }
    for (size_t i = 0; i < g.mat.size(); ++i) {
        d.re[i] = (T)g.mat[i].real();
        d.im[i] = (T)g.mat[i].imag();
    }
    void *single = table ? nullptr : sv.data;
    auto go = [&](auto u_tag, auto nt_tag) {
        constexpr int U = decltype(u_tag)::value, NT = decltype(nt_tag)::value;
        dim3 grid(grid_for(n_items, (uint64_t)NT * U), (unsigned)n_vecs);
        k_apply_diag<T, V, U, NT><<<grid, NT, 0, sv.stream>>>(single, table, n_items, holes, g.ctrl_mask >> shift, d);
    };
    using std::integral_constant;
    // QSV_DIAG_SHAPE (tools/ab_shapes2.py): U=8, NT=256 +1 % (RZ) / +4 % (CZ) over U=4, NT=256
    switch (env_shape("QSV_DIAG_SHAPE", 2)) {
    case 0: go(integral_constant<int, 4>{}, integral_constant<int, 256>{}); break;
    case 1: go(integral_constant<int, 8>{}, integral_constant<int, 128>{}); break;
    case 3: go(integral_constant<int, 2>{}, integral_constant<int, 256>{}); break;
    default: go(integral_constant<int, 8>{}, integral_constant<int, 256>{}); break;
    }
    QSV_CUDA(cudaGetLastError());
}

bool dense1_pairs_up() { return dense1_shape() == 7; }

// (2x2 gate on bit t) -> (gate (x) identity) on bits (t, s), s = the nearest free bit below t (else above), s >= lowest
LoweredGate pair_up(const LoweredGate &g, int n, int lowest) {
    const uint64_t tb = g.offs[1] ^ g.offs[0];
    int t = 0;
    while (!((tb >> t) & 1ull)) ++t;
    auto used = [&](int b) { return std::find(g.holes.begin(), g.holes.end(), b) != g.holes.end(); };
    int s = -1;
    for (int b = t - 1; b >= lowest && s < 0; --b)
        if (!used(b)) s = b;
    for (int b = t + 1; b < n && s < 0; ++b)
        if (!used(b)) s = b;
    QSV_CHECK(s >= 0, "internal: no free bit to pair a single-target gate with");
    LoweredGate r = g;
    r.k = 2;
    r.holes.push_back(s);
    std::sort(r.holes.begin(), r.holes.end());
    r.offs.assign(4, 0);
    for (int c = 0; c < 4; ++c) r.offs[c] = g.offs[c >> 1] | ((uint64_t)(c & 1) << s);
    r.mat.assign(16, cplx(0.0, 0.0));
    for (int a = 0; a < 2; ++a)
        for (int a2 = 0; a2 < 2; ++a2)
            for (int b = 0; b < 2; ++b) r.mat[((a << 1) | b) * 4 + ((a2 << 1) | b)] = g.mat[a * 2 + a2];
    r.tgt_bits.clear();
    return r;
}

void launch_any(State &sv, const LoweredGate &g, void *const *table, int n_vecs) {
    sv.use();
    if (g.kind == LoweredGate::NOP) return;
    sv.stat_launches += 1;
    sv.stat_sweeps += 1;
    const bool f32 = sv.dtype == QSV_C64;
    if (g.kind == LoweredGate::DENSE) {
        if (g.k > 4) {
            if (f32)
                launch_dense_large<float>(sv, g, table, n_vecs);
            else
                launch_dense_large<double>(sv, g, table, n_vecs);
            return;
        }
        const bool bit0 = !g.holes.empty() && g.holes[0] == 0;
        if (g.k == 1 && g.offs[0] == 0 && g.offs[1] == 1 && sv.n >= 2) {
            if (f32)
                launch_dense_bit0<float>(sv, g, table, n_vecs);
            else
                launch_dense_bit0<double>(sv, g, table, n_vecs);
            return;
        }
        if (g.k == 1 && dense1_pairs_up() && (int)g.holes.size() < sv.n - (f32 && !bit0 ? 1 : 0)) {
            // A/B (QSV_DENSE1_SHAPE=7): run the 2x2 gate as (gate (x) identity) on the target and a free neighbour bit
            const LoweredGate g2 = pair_up(g, sv.n, f32 && !bit0 ? 1 : 0);
            if (!f32)
                launch_dense_k<double, 1>(sv, g2, table, n_vecs);
            else if (bit0)
                launch_dense_k<float, 1>(sv, g2, table, n_vecs);
            else
                launch_dense_k<float, 2>(sv, g2, table, n_vecs);
            return;
        }
        if (!f32 && g.k == 1 && !bit0 && sv.n >= 2 && dense1_shape() == 8)
            launch_dense_t<double, 1, 2>(sv, g, table, n_vecs);  // A/B: two neighbouring amplitudes per 256-bit access
        else if (!f32)
            launch_dense_k<double, 1>(sv, g, table, n_vecs);
        else if (bit0 || sv.n < 1)
            launch_dense_k<float, 1>(sv, g, table, n_vecs);
        else
            launch_dense_k<float, 2>(sv, g, table, n_vecs);
        return;
    }
    // DIAG / PARITY
    const bool bit0 = g.ctrl_mask & 1ull;
    if (!f32)
        launch_diag_t<double, 1>(sv, g, table, n_vecs);
    else if (bit0)
        launch_diag_t<float, 1>(sv, g, table, n_vecs);
    else
        launch_diag_t<float, 2>(sv, g, table, n_vecs);
}

}  // namespace

void launch_gate(State &sv, const LoweredGate &g) { launch_any(sv, g, nullptr, 1); }

void launch_gate_multi(State &sv, const LoweredGate &g, void *const *dev_table, int n_vecs) {
    if (n_vecs <= 0) return;
    launch_any(sv, g, dev_table, n_vecs);
}

void launch_fill_basis(State &sv, uint64_t index) {
    sv.use();
    QSV_CHECK(index < sv.length(), "basis state index out of range");
    constexpr int NT = 256;
    if (sv.dtype == QSV_C128) {
        const uint64_t ne = sv.length();
        k_fill_basis<double, 1, NT><<<(unsigned)std::min<uint64_t>((ne + NT - 1) / NT, NUM_SMS * 16), NT, 0,
                                      sv.stream>>>(sv.data, ne, index);
    } else if (sv.n >= 1) {
        const uint64_t ne = sv.length() / 2;
        k_fill_basis<float, 2, NT><<<(unsigned)std::min<uint64_t>((ne + NT - 1) / NT, NUM_SMS * 16), NT, 0,
                                     sv.stream>>>(sv.data, ne, index);
    } else {
        k_fill_basis<float, 1, NT><<<1, NT, 0, sv.stream>>>(sv.data, 1, index);
    }
    QSV_CUDA(cudaGetLastError());
}

void launch_scatter(State &sv, const int64_t *dev_idx, const void *dev_vals, size_t count) {
    sv.use();
    if (count == 0) return;
    const unsigned grid = grid_for(count, 256);
    if (sv.dtype == QSV_C128)
        k_scatter<double><<<grid, 256, 0, sv.stream>>>(sv.data, dev_idx, dev_vals, count, sv.length());
    else
        k_scatter<float><<<grid, 256, 0, sv.stream>>>(sv.data, dev_idx, dev_vals, count, sv.length());
    QSV_CUDA(cudaGetLastError());
}

void launch_axpy(State &sv, cplx alpha, const void *x, void *y) {
    sv.use();
    constexpr int NT = 256;
    if (sv.dtype == QSV_C128) {
        const uint64_t ne = sv.length();
        k_axpy<double, 1, NT><<<(unsigned)std::min<uint64_t>((ne + NT - 1) / NT, NUM_SMS * 16), NT, 0, sv.stream>>>(
            alpha.real(), alpha.imag(), x, y, ne);
    } else if (sv.n >= 1) {
        const uint64_t ne = sv.length() / 2;
        k_axpy<float, 2, NT><<<(unsigned)std::min<uint64_t>((ne + NT - 1) / NT, NUM_SMS * 16), NT, 0, sv.stream>>>(
            (float)alpha.real(), (float)alpha.imag(), x, y, ne);
    } else {
        k_axpy<float, 1, NT><<<1, NT, 0, sv.stream>>>((float)alpha.real(), (float)alpha.imag(), x, y, 1);
    }
    QSV_CUDA(cudaGetLastError());
}

}  // namespace qsv
