// micro-benchmark: dependent-issue latency of fp64/fp32 ops on one warp (cycles per op)
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(double* out, long long* cyc, double a, double b) {
    double x = a; float xf = (float)a, bf = (float)b;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) {
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (OP == 0) x = fma(x, b, a);
            if (OP == 1) x = x * b;
            if (OP == 2) x = x + b;
            if (OP == 3) xf = fmaf(xf, bf, bf);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[OP] = t1 - t0; out[OP] = x + xf; }
}
int main() {
    double* out; long long* cyc; cudaMallocManaged(&out, 64); cudaMallocManaged(&cyc, 64);
    k<0><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); k<1><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    k<2><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999); k<3><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    const char* n[] = {"DFMA", "DMUL", "DADD", "FFMA"};
    for (int i = 0; i < 4; i++) printf("%s dependent latency: %.2f cycles\n", n[i], cyc[i] / (1024.0 * 16));
    // throughput: 8 warps x independent chains
    return 0;
}
