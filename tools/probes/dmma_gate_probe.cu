// Probe (tools only): a complex 4x4 gate on the two lowest LANE bits of a warp through mma.sync.m8n8k4.f64.
// Lane 4m+t holds amplitude t (re, im) of quad m; the real 8x8 form R of the gate goes in as the B operand
// (lane holds R[lane/4][2*(lane%4)+s] for the k-step s = re / im inputs); the D fragment comes back as (re_t, im_t) of
// the same quad -- i.e. in place.  Checks the fragment-layout reasoning of DESIGN.md 4.2 before it goes into the kernel.
#include <cmath>
#include <complex>
#include <cstdio>
#include <cuda_runtime.h>
#include <random>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

__global__ void k(double2 *x, const double *R) {
    const int lane = threadIdx.x & 31;
    const double2 bb = *reinterpret_cast<const double2 *>(R + (lane >> 2) * 8 + 2 * (lane & 3));
    double2 v = x[threadIdx.x];
    double d0 = 0.0, d1 = 0.0;
    dmma(d0, d1, v.x, bb.x);
    dmma(d0, d1, v.y, bb.y);
    x[threadIdx.x] = make_double2(d0, d1);
}

int main() {
    std::mt19937 g(7);
    std::normal_distribution<double> nd;
    std::complex<double> M[16], in[32], want[32];
    for (auto &c : M) c = {nd(g), nd(g)};
    for (auto &c : in) c = {nd(g), nd(g)};
    for (int m = 0; m < 8; ++m)
        for (int t = 0; t < 4; ++t) {
            std::complex<double> s = 0;
            for (int u = 0; u < 4; ++u) s += M[t * 4 + u] * in[4 * m + u];
            want[4 * m + t] = s;
        }
    double R[64];
    for (int t = 0; t < 4; ++t)
        for (int u = 0; u < 4; ++u) {
            R[(2 * t) * 8 + 2 * u] = M[t * 4 + u].real();
            R[(2 * t) * 8 + 2 * u + 1] = -M[t * 4 + u].imag();
            R[(2 * t + 1) * 8 + 2 * u] = M[t * 4 + u].imag();
            R[(2 * t + 1) * 8 + 2 * u + 1] = M[t * 4 + u].real();
        }
    double2 h[32], *d;
    double *dR;
    for (int i = 0; i < 32; ++i) h[i] = make_double2(in[i].real(), in[i].imag());
    cudaMalloc(&d, sizeof(h));
    cudaMalloc(&dR, sizeof(R));
    cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    cudaMemcpy(dR, R, sizeof(R), cudaMemcpyHostToDevice);
    k<<<1, 32>>>(d, dR);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < 32; ++i) err = fmax(err, std::abs(std::complex<double>(h[i].x, h[i].y) - want[i]));
    printf("dmma gate probe: max abs err %.3e (%s) %s\n", err, err < 1e-13 ? "layout OK" : "LAYOUT WRONG",
           cudaGetErrorString(cudaDeviceSynchronize()));
    return err < 1e-13 ? 0 : 1;
}
