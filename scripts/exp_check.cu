// accuracy of the chain's exp variants against the library exp (ulp statistics)
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>
#include "../openmm_drudenose_b200/csrc/tgnh_device.cuh"
__global__ void k(const double* x, double* a, double* b, double* c, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool bad = false;
    a[i] = exp(x[i]); b[i] = tgnh::exp_full(x[i]); c[i] = tgnh::chain_exp<true>(x[i], bad);
}
int main() {
    const int n = 1 << 20;
    double *x, *a, *b, *c;
    cudaMallocManaged(&x, n * 8); cudaMallocManaged(&a, n * 8); cudaMallocManaged(&b, n * 8); cudaMallocManaged(&c, n * 8);
    srand(1);
    for (int i = 0; i < n; i++) {
        double u = rand() / (double)RAND_MAX * 2 - 1;
        int cls = i % 4;
        x[i] = cls == 0 ? u * 0.03125 : cls == 1 ? u * 2.0 : cls == 2 ? u * 50.0 : u * 720.0;
    }
    k<<<n / 256, 256>>>(x, a, b, c, n);
    cudaDeviceSynchronize();
    double maxFull = 0, maxFast = 0; int worst = 0;
    for (int i = 0; i < n; i++) {
        if (a[i] == 0 || std::isinf(a[i]) || a[i] < 2.3e-308 || std::fabs(x[i]) > 700.0) continue;
        double ulp = std::fabs(std::nextafter(a[i], INFINITY) - a[i]);
        double e = std::fabs(b[i] - a[i]) / ulp;
        if (e > maxFull) { maxFull = e; worst = i; }
        if (i % 4 == 0) maxFast = fmax(maxFast, std::fabs(c[i] - a[i]) / ulp);
    }
    printf("exp_full vs exp: max %.2f ulp (x=%g); small-argument polynomial vs exp: max %.2f ulp\n", maxFull, x[worst], maxFast);
    printf("exp_full(-800)=%g exp_full(800)=%g exp_full(0)=%.17g exp_full(-707)=%g lib=%g\n", b[0]*0, 0.0, 1.0, 0.0, 0.0);
    return 0;
}
