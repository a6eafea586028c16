// micro-benchmark of the device Nose-Hoover chain (tgnh_device.cuh: chain_phase): cycles per sub-step of one warp,
// near equilibrium (short exp polynomial) and far from it (full-range exp).  Build: nvcc -I<dir of tgnh_device.cuh>.
#include <cstdio>
#include <vector>
#include <cmath>
#include "tgnh_device.cuh"
using namespace tgnh;
__global__ void __launch_bounds__(32, 1) bench(const __grid_constant__ ChainView c, int mode, int reps, long long* cyc) {
    long long t0 = clock64();
    for (int r = 0; r < reps; r++) chain_phase(c, mode, threadIdx.x);
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    const int T = 6, M = 3;
    for (int S : {20, 200})
    for (int hot = 0; hot < 2; hot++) {
        ChainView c{};
        c.T = T; c.G = T - 2; c.M = M; c.S = S; c.useDrudeNH = 1;
        c.dt = 0.001; c.kT = 2.494; c.kTD = 0.0083; c.dtc = c.dt / S;
#ifdef HAVE_EXP_TABLES
        chain_exp_tables(c);
#endif
        const size_t TM = T * M, TM1 = T * (M + 1);
        std::vector<double> h(5 * TM + TM1 + 9 * T + 8, 0.0);
        double* d; cudaMalloc(&d, h.size() * 8);
        size_t o = 0;
        auto take = [&](size_t n) { double* p = d + o; o += n; return p; };
        double* etaMass = take(TM); double* inv = take(TM); c.etaMass = etaMass; c.invEtaMass = inv;
        c.eta = take(TM); c.etaDot = take(TM1); c.etaDotDot = take(TM);
        c.nkbt = take(T); c.ke2 = take(T); c.ke2Local = take(T); c.ke2Used = take(T); c.pending = take(T); c.scaleA = take(T); c.vscale = take(T); c.keSum = take(1);
#ifdef HAVE_EXP_TABLES
        c.expHint = take(T);
#endif
        for (int g = 0; g < T; g++) {
            const double kT = g == T - 1 ? c.kTD : c.kT, tau = g == T - 1 ? 0.005 : 0.1, dof = 7.5e6;
            for (int i = 0; i < M; i++) { const double Q = (i == 0 ? dof : 1.0) * kT * tau * tau; h[g * M + i] = Q; h[TM + g * M + i] = 1.0 / Q; }
            h[(c.nkbt - d) + g] = dof * kT;
            h[(c.ke2 - d) + g] = dof * kT * (hot ? 30.0 : 1.0003);
            h[(c.pending - d) + g] = 1.0;
        }
        cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        long long* cyc; cudaMallocManaged(&cyc, 8);
        bench<<<1, 32>>>(c, CHAIN_SECOND, 1, cyc); cudaDeviceSynchronize();
        cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        const int reps = 50;
        bench<<<1, 32>>>(c, CHAIN_SECOND, reps, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        std::vector<double> out(h.size()); cudaMemcpy(out.data(), d, h.size() * 8, cudaMemcpyDeviceToHost);
        printf("%s: %s  %.1f cycles per sub-step (M=3, T=6, S=%d), vscale[0]=%.15g etaDot[0]=%.15g etaDot[drude]=%.15g\n", hot ? "far from equilibrium" : "near equilibrium",
               cudaGetErrorString(e), (double)*cyc / (reps * S), S, out[(c.vscale - d)], out[(c.etaDot - d)], out[(c.etaDot - d) + (T - 1) * (M + 1)]);
    }
    return 0;
}
