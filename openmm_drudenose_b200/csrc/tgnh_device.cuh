// Device-side building blocks of the B200 (sm_100a) TGNH step: PTX wrappers for the TMA bulk-copy
// pipeline (cp.async.bulk + mbarrier), the per-particle descriptor word, and the fp64 Nose-Hoover
// chain that replaces the reference's host loop (CudaDrudeTGNHKernels.cpp:559-642).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tgnh {

constexpr int TILE = 512;          // particles per tile == threads per CTA
constexpr int PADW = 520;          // 4-aligned window that covers any tile: 3 + 512 + 3 rounded up
constexpr int MAX_T = 32;          // thermostats (G + 2)
constexpr int MAX_M = 16;          // Nose-Hoover chain length
constexpr int MAX_RES = 128;       // particles per residue handled in-tile

// ---- descriptor word: one uint32 per particle, built once in tgnh_create ----------------------
//  [7:0]   temperature group
//  [9:8]   role: 0 ordinary, 1 Drude particle (pairParticles.x), 2 parent (pairParticles.y)
//  [16:10] offset back to the first particle of the residue
//  [23:17] offset forward to the last particle of the residue
//  [31:24] signed offset to the pair partner
constexpr uint32_t ROLE_NORMAL = 0, ROLE_DRUDE = 1, ROLE_PARENT = 2;
__host__ __device__ inline uint32_t desc_pack(int tg, uint32_t role, int offFirst, int offLast, int partner) {
    return (uint32_t)(tg & 0xff) | (role << 8) | ((uint32_t)offFirst << 10) | ((uint32_t)offLast << 17) |
           ((uint32_t)(partner & 0xff) << 24);
}
__device__ __forceinline__ int desc_tg(uint32_t d) { return d & 0xff; }
__device__ __forceinline__ uint32_t desc_role(uint32_t d) { return (d >> 8) & 3; }
__device__ __forceinline__ int desc_off_first(uint32_t d) { return (d >> 10) & 0x7f; }
__device__ __forceinline__ int desc_off_last(uint32_t d) { return (d >> 17) & 0x7f; }
__device__ __forceinline__ int desc_partner(uint32_t d) { return ((int32_t)d) >> 24; }

// ---- mbarrier + bulk async copy (TMA, 1-D) -----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// L2 eviction policy for data that is streamed once per kernel
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// global -> shared bulk copy; completion is signalled on `bar` as transaction bytes (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// streaming 128-bit store (written once, read by the next kernel from HBM/L2)
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// programmatic dependent launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- device-resident thermostat state ----------------------------------------------------------
// One block of doubles in HBM; all sub-arrays are views into it (see tgnh.cu: layout_chain()).
struct ChainView {
    int T, G, M, S, useDrudeNH;
    double dt, kT, kTD;
    double* etaMass;    // [T][M]
    double* eta;        // [T][M]
    double* etaDot;     // [T][M+1]   last column: permanent 0 (CudaDrudeTGNHKernels.cpp:96,252)
    double* etaDotDot;  // [T][M]
    double* nkbt;       // [T]
    double* ke2;        // [T]  2*KE of the velocities as stored (global: after the all-reduce)
    double* ke2Local;   // [T]  this rank's share (sharded runs)
    double* ke2Used;    // [T]  2*KE consumed by the most recent chain update
    double* pending;    // [T]  scale factors computed but not yet applied to velm
    double* scaleA;     // [T]  factors the next streaming kernel applies
    double* vscale;     // [T]  factors of the most recent chain update (vscaleFactorsVec)
    double* keSum;      // [1]
};

enum ChainMode {
    CHAIN_NONE = 0,        // reduction only (sharded runs: all-reduce first, chain in its own launch)
    CHAIN_FIRST = 1,       // first half-step of a step: consumes pending^2 * ke2, scaleA = pending * s
    CHAIN_SECOND = 2,      // second half-step: consumes ke2, pending = s, scaleA = s
    CHAIN_SECOND_FIRST = 3 // second half-step immediately followed by the next step's first half-step
};

// exp(x) for the chain.  |x| is ~1e-3 here (dtc/8 * etaDot); for |x| <= 2^-5 a degree-9 Taylor
// polynomial is exact to < 0.25 ulp before the final rounding, which puts it in the same <= 1 ulp
// class as CUDA's exp() and glibc's, at a fraction of the dependent-instruction depth.  The chain is
// strictly serial (40 updates x ~3 dependent exps per step) and sits between two streaming kernels,
// so its latency is directly visible in the step time.
__device__ __forceinline__ double chain_exp(double x) {
    if (fabs(x) <= 0.03125) {
        const double x2 = x * x;
        // Estrin: pairs of terms, then powers of x^2
        const double p01 = 1.0 + x;
        const double p23 = fma(x, 1.0 / 6.0, 0.5);
        const double p45 = fma(x, 1.0 / 120.0, 1.0 / 24.0);
        const double p67 = fma(x, 1.0 / 5040.0, 1.0 / 720.0);
        const double p89 = fma(x, 1.0 / 362880.0, 1.0 / 40320.0);
        const double x4 = x2 * x2;
        const double q0 = fma(p23, x2, p01);
        const double q1 = fma(p67, x2, p45);
        const double x8 = x4 * x4;
        return fma(p89, x8, fma(q1, x4, q0));
    }
    return exp(x);
}

// One thermostat's half-step chain update; restates CudaDrudeTGNHKernels.cpp:560-595 (relative and COM
// groups) and :597-642 (Drude group) as one loop nest: the Drude group differs only in the number of
// live links (1 unless useDrudeNHChains) and in kT.  Returns the velocity scale factor.
// Deviation (documented in DESIGN.md): the Q0 > 0 guard of :561 is applied to the Drude group too, so a
// system without Drude pairs yields scale 1 instead of NaN.
template <int MC>  // MC > 0: chain length known at compile time (registers); MC == 0: runtime M <= MAX_M
__device__ double chain_update(const ChainView& c, int g, double ke) {
    constexpr int CAP = MC > 0 ? MC : MAX_M;
    const int M = MC > 0 ? MC : c.M;
    const bool isDrude = (g == c.T - 1);
    const int nl = (isDrude && !c.useDrudeNH) ? 1 : M;
    const double kTl = isDrude ? c.kTD : c.kT;
    const double dtc = c.dt / c.S, dtc2 = dtc / 2.0, dtc4 = dtc / 4.0, dtc8 = dtc / 8.0;
    double Q[CAP], invQ[CAP], eta[CAP], ed[CAP + 1], edd[CAP];
#pragma unroll
    for (int i = 0; i < CAP; i++) {
        invQ[i] = 1.0;
        if (i < M) {
            Q[i] = c.etaMass[g * M + i];
            invQ[i] = 1.0 / Q[i];
            eta[i] = c.eta[g * M + i];
            ed[i] = c.etaDot[g * (M + 1) + i];
            edd[i] = c.etaDotDot[g * M + i];
        } else {
            Q[i] = 1.0; eta[i] = 0.0; ed[i] = 0.0; edd[i] = 0.0;
        }
    }
    ed[CAP] = 0.0;
    // ed[M] is the permanent zero; for a Drude group without chains ed[1] may hold user-set state
    const double edTop = c.etaDot[g * (M + 1) + nl];
    const double nkbt = c.nkbt[g];
    const bool live = Q[0] > 0;
    const double invQ0 = live ? 1.0 / Q[0] : 0.0;
    double scale = 1.0, ef = 1.0;
    if (live) edd[0] = (ke - nkbt) * invQ0;
    for (int iter = 0; iter < c.S; iter++) {
#pragma unroll
        for (int i = CAP - 1; i >= 0; i--) {
            if (i < nl) {
                const double up = (i == nl - 1) ? edTop : ed[i + 1];
                ef = chain_exp(-dtc8 * up);
                ed[i] *= ef; ed[i] += edd[i] * dtc4; ed[i] *= ef;
            }
        }
        scale *= chain_exp(-dtc2 * ed[0]);
        ke *= chain_exp(-dtc * ed[0]);
#pragma unroll
        for (int i = 0; i < CAP; i++)
            if (i < nl) eta[i] += dtc2 * ed[i];
        if (live) edd[0] = (ke - nkbt) * invQ0;
        ed[0] *= ef; ed[0] += edd[0] * dtc4; ed[0] *= ef;
#pragma unroll
        for (int i = 1; i < CAP; i++) {
            if (i < nl) {
                const double up = (i == nl - 1) ? edTop : ed[i + 1];
                ef = chain_exp(-dtc8 * up);
                ed[i] *= ef;
                edd[i] = (Q[i - 1] * ed[i - 1] * ed[i - 1] - kTl) * invQ[i];
                ed[i] += edd[i] * dtc4; ed[i] *= ef;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CAP; i++) {
        if (i < M) {
            c.eta[g * M + i] = eta[i];
            c.etaDot[g * (M + 1) + i] = ed[i];
            c.etaDotDot[g * M + i] = edd[i];
        }
    }
    return scale;
}

__device__ __forceinline__ double chain_update_any(const ChainView& c, int g, double ke) {
    switch (c.M) {
        case 1: return chain_update<1>(c, g, ke);
        case 2: return chain_update<2>(c, g, ke);
        case 3: return chain_update<3>(c, g, ke);
        case 4: return chain_update<4>(c, g, ke);
        default: return chain_update<0>(c, g, ke);
    }
}

// Executed by ONE warp (lanes 0..T-1 own one thermostat each; T <= 32).  c.ke2 must hold the global
// 2*KE sums of the velocities as stored.
__device__ void chain_phase(const ChainView& c, int mode, int lane) {
    const bool on = lane < c.T;
    double ke = on ? c.ke2[lane] : 0.0;
    double pend = on ? c.pending[lane] : 1.0;
    double used = 0.0, s = 1.0;
    if (mode == CHAIN_SECOND || mode == CHAIN_SECOND_FIRST) {
        used = ke;
        if (on) { s = chain_update_any(c, lane, ke); pend = s; }
    }
    if (mode == CHAIN_FIRST || mode == CHAIN_SECOND_FIRST) {
        const double keEff = pend * pend * ke;
        used = keEff;
        if (on) { s = chain_update_any(c, lane, keEff); }
        if (on) { c.scaleA[lane] = pend * s; c.pending[lane] = 1.0; }
    } else if (on) {
        c.scaleA[lane] = pend;
        c.pending[lane] = pend;
    }
    if (on) { c.vscale[lane] = s; c.ke2Used[lane] = used; }
    // KESum = 0.5 * sum_g 2KE_g in group order (CudaDrudeTGNHKernels.cpp:493-497)
    double tot = 0.0;
    for (int g = 0; g < c.T; g++) tot += __shfl_sync(0xffffffffu, used, g);
    if (lane == 0) *c.keSum = 0.5 * tot;
}

}  // namespace tgnh
