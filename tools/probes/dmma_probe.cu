// Microbenchmark (tools only, not part of the product): FP64 throughput of B200 through the vector pipe (DFMA), the
// tensor pipe (mma.sync.m8n8k4.f64 = DMMA) and both interleaved -- decides whether fused 4x4 complex128 blocks on lane
// bits are worth moving to DMMA in the register-tile kernel (DESIGN.md 4.2, round-2 plan).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE> __global__ void __launch_bounds__(256, 2) k(double *out, int iters, double a, double b) {
    double c[16], x[16];
    for (int i = 0; i < 16; ++i) {
        c[i] = threadIdx.x * 1e-9 + i;
        x[i] = i * 1e-3;
    }
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);  // 16 independent DFMA chains
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) dmma(c[i], c[i + 1], a, b);  // 8 independent DMMA chains
        }
    }
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i] + x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, double *out, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 2 * 8;
    k<MODE><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)blocks * 8;
    const double dfma = (MODE == 0 || MODE == 2) ? warps * iters * 16.0 * 32.0 : 0.0;          // FMAs
    const double mma = (MODE == 1 || MODE == 2) ? warps * iters * 8.0 * 256.0 : 0.0;           // FMAs (8x8x4 per warp)
    printf("%-28s %8.3f ms  DFMA %7.2f TFLOP/s  DMMA %7.2f TFLOP/s  total %7.2f TFLOP/s\n", name, ms, 2 * dfma / ms / 1e9,
           2 * mma / ms / 1e9, 2 * (dfma + mma) / ms / 1e9);
}

int main() {
    double *out;
    cudaMalloc(&out, sizeof(double) * 148 * 2 * 8 * 256);
    run<0>("DFMA only", out, 4000);
    run<1>("DMMA only", out, 4000);
    run<2>("DFMA + DMMA interleaved", out, 4000);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
