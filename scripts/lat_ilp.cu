// micro-benchmark: one warp, ILP independent fp64 FMA chains -> cycles per DFMA (is a single warp issue-limited on the fp64 pipe?)
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP> __global__ void k(double* out, long long* cyc, double a, double b, int slot) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = a + j;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 1024; i++) {
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int j = 0; j < ILP; j++) x[j] = fma(x[j], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; j++) s += x[j];
    if (threadIdx.x == 0) { cyc[slot] = t1 - t0; out[slot] = s; }
}
int main() {
    double* out; long long* cyc; cudaMallocManaged(&out, 256); cudaMallocManaged(&cyc, 256);
    k<1><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 0); k<2><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 1);
    k<4><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 2); k<8><<<1, 32>>>(out, cyc, 1.0000001, 0.9999999, 3);
    // lanes: does an 8-lane warp issue faster than a 32-lane one?
    k<4><<<1, 8>>>(out, cyc, 1.0000001, 0.9999999, 4); k<1><<<1, 8>>>(out, cyc, 1.0000001, 0.9999999, 5);
    cudaDeviceSynchronize();
    const int ilp[] = {1, 2, 4, 8, 4, 1};
    const char* n[] = {"32 lanes", "32 lanes", "32 lanes", "32 lanes", "8 lanes", "8 lanes"};
    for (int i = 0; i < 6; i++) printf("ILP %d (%s): %.2f cycles per DFMA, %.2f per round\n", ilp[i], n[i], cyc[i] / (1024.0 * 8 * ilp[i]), cyc[i] / (1024.0 * 8));
    return 0;
}
