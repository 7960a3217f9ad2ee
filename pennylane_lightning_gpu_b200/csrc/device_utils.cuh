// Device-side helpers shared by the kernel translation units.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

namespace qsv {

constexpr int MAX_HOLES = 40;
constexpr int NUM_SMS = 148;  // B200

// Sorted bit positions at which a dense counter gets a zero bit inserted.
struct Holes {
    int n;
    unsigned char pos[MAX_HOLES];
};

__host__ __device__ __forceinline__ uint64_t expand_index(uint64_t o, const Holes &h) {
    for (int j = 0; j < h.n; ++j) {
        const int p = h.pos[j];
        o = ((o >> p) << (p + 1)) | (o & ((1ull << p) - 1ull));
    }
    return o;
}

// V complex numbers of precision T moved with one (up to) 128-bit access.
template <typename T, int V> struct VecOf;
template <> struct VecOf<double, 1> { using type = double2; };
template <> struct VecOf<float, 1> { using type = float2; };
template <> struct VecOf<float, 2> { using type = float4; };
struct alignas(32) Double4 {
    double x, y, z, w;
};
template <> struct VecOf<double, 2> { using type = Double4; };  // 256-bit access (LDG.E.ENL2.256), see load_elem below

template <typename T, int V>
__device__ __forceinline__ void load_elem(T (&c)[2 * V], const void *base, uint64_t idx) {
    using VT = typename VecOf<T, V>::type;
    const VT v = reinterpret_cast<const VT *>(base)[idx];
    c[0] = v.x;
    c[1] = v.y;
    if constexpr (V == 2) {
        c[2] = v.z;
        c[3] = v.w;
    }
}

// two neighbouring complex128 amplitudes with one 256-bit access
template <>
__device__ __forceinline__ void load_elem<double, 2>(double (&c)[4], const void *base, uint64_t idx) {
    const double *p = reinterpret_cast<const double *>(base) + 4 * idx;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c[0]), "=d"(c[1]), "=d"(c[2]), "=d"(c[3]) : "l"(p));
}

template <typename T, int V>
__device__ __forceinline__ void store_elem(void *base, uint64_t idx, const T (&c)[2 * V]) {
    using VT = typename VecOf<T, V>::type;
    VT v;
    v.x = c[0];
    v.y = c[1];
    if constexpr (V == 2) {
        v.z = c[2];
        v.w = c[3];
    }
    reinterpret_cast<VT *>(base)[idx] = v;
}

template <>
__device__ __forceinline__ void store_elem<double, 2>(void *base, uint64_t idx, const double (&c)[4]) {
    double *p = reinterpret_cast<double *>(base) + 4 * idx;
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(c[0]), "d"(c[1]), "d"(c[2]), "d"(c[3]) : "memory");
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum (re, im) over the block with warp shuffles, then one atomicAdd pair per block.
template <int NT>
__device__ __forceinline__ void block_accumulate(double re, double im, double *out) {
    __shared__ double s_red[2][NT / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    re = warp_sum(re);
    im = warp_sum(im);
    if (lane == 0) {
        s_red[0][warp] = re;
        s_red[1][warp] = im;
    }
    __syncthreads();
    if (warp == 0) {
        re = lane < NT / 32 ? s_red[0][lane] : 0.0;
        im = lane < NT / 32 ? s_red[1][lane] : 0.0;
        re = warp_sum(re);
        im = warp_sum(im);
        if (lane == 0) {
            atomicAdd(out, re);
            atomicAdd(out + 1, im);
        }
    }
}

inline Holes make_holes(const int *pos, int n, int shift) {
    Holes h;
    h.n = n;
    for (int i = 0; i < MAX_HOLES; ++i) h.pos[i] = 0;
    for (int i = 0; i < n; ++i) h.pos[i] = (unsigned char)(pos[i] - shift);
    return h;
}

}  // namespace qsv
