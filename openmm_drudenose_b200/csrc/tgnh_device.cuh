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
//  [6:0]   temperature group (< MAX_T)
//  [7]     the particle belongs to a "big" residue (more than MAX_RES particles, e.g. a protein): its residue's COM
//          velocity comes from the pre-pass kernel's table instead of the tile; offsets [23:10] are 0
//  [9:8]   role: 0 ordinary, 1 Drude particle (pairParticles.x), 2 parent (pairParticles.y)
//  [16:10] offset back to the first particle of the residue
//  [23:17] offset forward to the last particle of the residue
//  [31:24] signed offset to the pair partner
constexpr uint32_t ROLE_NORMAL = 0, ROLE_DRUDE = 1, ROLE_PARENT = 2;
__host__ __device__ inline uint32_t desc_pack(int tg, uint32_t role, int offFirst, int offLast, int partner, bool big = false) {
    return (uint32_t)(tg & 0x7f) | (big ? 0x80u : 0u) | (role << 8) | ((uint32_t)offFirst << 10) | ((uint32_t)offLast << 17) |
           ((uint32_t)(partner & 0xff) << 24);
}
__device__ __forceinline__ int desc_tg(uint32_t d) { return d & 0x7f; }
__device__ __forceinline__ bool desc_big(uint32_t d) { return (d & 0x80u) != 0; }
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Departure counter of a pipeline stage (relaxed shared-memory atomic).  No fence is needed for the refill that
// follows the last departure: SMs issue in order and an instruction issues only once its source registers
// are ready, so every shared-memory load of the stage whose value is consumed before this atomic (all of
// them: nothing but accumulators is carried across tiles) has returned its data by the time the atomic
// issues, i.e. long before the TMA write of the next tile can land.  (A release/acquire atomic here makes
// every warp wait for its outstanding global stores: measured +13 % on the second-half kernel.)
__device__ __forceinline__ uint32_t atom_add_shared(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.relaxed.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
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
__device__ __forceinline__ uint64_t policy_evict_normal() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// asynchronous prefetch of a contiguous range into L2 (no register or shared-memory footprint)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
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
// 128-bit store with the default L2 policy (velm: re-read by the next launch, see "L2 hand-over")
__device__ __forceinline__ void st_global(float4* p, float4 v) {
    asm volatile("st.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// programmatic dependent launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- device-resident thermostat state ----------------------------------------------------------
// One block of doubles in HBM; all sub-arrays are views into it (see tgnh.cu: layout_chain()).
struct ChainView {
    int T, G, M, S, useDrudeNH;
    double dt, kT, kTD;
    double dtc;         // dt / S
    double* etaMass;    // [T][M]
    double* invEtaMass; // [T][M]   1 / etaMass (0 where the mass is 0)
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

// ---- sharded runs: the kinetic-energy exchange over NVLink peer memory -------------------------------------------
// Every rank owns an inbox in its HBM that all peers map (CUDA IPC).  The last CTA of a reducing launch writes the rank's partial
// sums into slot [seq & 1][rank] of EVERY inbox; the chain launch that follows waits for all ranks' words in its own inbox and
// adds the partials in rank order, so every rank forms bit-identical sums and no collective launch sits between the two kernels.
// The words are self-validating (the layout of NCCL's LL protocol): each 8-byte word carries 32 bits of payload and the 32-bit
// sequence number of the reduction, and an aligned 8-byte store is a single transaction, so a reader sees a word either
// old or complete.  No fence, no separate flag: the exchange costs one NVLink write latency.  Two slots suffice: a peer can only be
// one reduction ahead, because its next one needs this rank's contribution to the current one.
constexpr int MAX_PEERS = 16;
struct PeerInbox {
    unsigned long long word[2][MAX_PEERS][2 * MAX_T];   // [slot][sending rank][2 g + half]: (payload << 32) | seq
    unsigned int error;           // set when a wait timed out (a peer died): reported by the next host-side query
    unsigned int pad;
    unsigned long long stamp[4];  // %globaltimer of this rank's most recent publish / gather start / gather end (diagnostics)
};
struct PeerView {
    int world, rank;              // world <= 1: not sharded (or the NCCL path is in use)
    unsigned int seq;             // number of this reduction, 1, 2, ...
    PeerInbox* inbox[MAX_PEERS];  // inbox[r]: rank r's inbox as mapped into this process (inbox[rank] is local)
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// all threads of the publishing CTA: thread i writes word (i mod 2T) of this rank's sums part[0..T) into rank (i div 2T)'s inbox
__device__ __forceinline__ void peer_publish(const PeerView& pv, const double* part, int T, int tid, int nthreads) {
    const int b = pv.seq & 1, words = 2 * T;
    for (int i = tid; i < pv.world * words; i += nthreads) {
        const int r = i / words, w = i - r * words;
        const unsigned long long bits = (unsigned long long)__double_as_longlong(part[w >> 1]);
        const unsigned long long payload = (w & 1) ? (bits >> 32) : (bits & 0xffffffffull);
        st_relaxed_sys(&pv.inbox[r]->word[b][pv.rank][w], (payload << 32) | pv.seq);
    }
    if (tid == 0) pv.inbox[pv.rank]->stamp[0] = global_timer_ns();
}

// one warp: wait for every rank's partial sums of reduction `seq`, add them in rank order into out[0..T)
__device__ __forceinline__ void peer_gather(const PeerView& pv, double* out, int T, int lane) {
    PeerInbox* in = pv.inbox[pv.rank];
    const int b = pv.seq & 1, words = 2 * T;
    const unsigned long long t0 = global_timer_ns();
    for (int i = lane; i < pv.world * words; i += 32) {
        const int r = i / words, w = i - r * words;
        while ((unsigned int)ld_relaxed_sys(&in->word[b][r][w]) != pv.seq)
            if (global_timer_ns() - t0 > 10000000000ull) { in->error = 1u; break; }     // 10 s: a peer is gone
    }
    __syncwarp();
    if (lane < T) {
        double x = 0.0;
        for (int r = 0; r < pv.world; r++) {
            const unsigned long long lo = ld_relaxed_sys(&in->word[b][r][2 * lane]), hi = ld_relaxed_sys(&in->word[b][r][2 * lane + 1]);
            x += __longlong_as_double((long long)((hi & 0xffffffff00000000ull) | (lo >> 32)));
        }
        out[lane] = x;
    }
    if (lane == 0) { in->stamp[1] = t0; in->stamp[2] = global_timer_ns(); }
    __syncwarp();
}

enum ChainMode {
    CHAIN_NONE = 0,        // no chain update
    CHAIN_FIRST = 1,       // first half-step of a step: consumes pending^2 * ke2, scaleA = pending * s
    CHAIN_SECOND = 2,      // second half-step: consumes ke2, pending = s, scaleA = s
    CHAIN_SECOND_FIRST = 3 // second half-step immediately followed by the next step's first half-step
};

// exp(x) for the chain.  |x| is ~1e-3 here (dtc/8 * etaDot); for |x| <= 2^-5 a degree-9 Taylor
// polynomial is exact to < 0.25 ulp before the final rounding, which puts it in the same <= 1 ulp
// class as CUDA's exp() and glibc's, at a fraction of the dependent-instruction depth.  The chain is
// strictly serial (40 updates x ~3 dependent exps per step) and sits between two streaming kernels,
// so its latency is directly visible in the step time.  FAST = false is the library exp().
// Full-range exp with a short dependency chain (~11 levels against ~20 for the library's Horner form):
// Cody-Waite reduction x = k ln2 + r, |r| <= 0.347, degree-13 Taylor polynomial in Estrin form (truncation
// 4e-18), scaling by 2^k through the exponent field.  <= 2 ulp; results below the normal range flush to 0.
__device__ __forceinline__ double exp_full(double x) {
    const double MAGIC = 6755399441055744.0;                  // 1.5 * 2^52: rounds to nearest integer
    const double t = fma(x, 1.4426950408889634, MAGIC);
    const int k = __double2loint(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    const double r2 = r * r;
    const double p01 = 1.0 + r;
    const double p23 = fma(r, 1.0 / 6.0, 0.5);
    const double p45 = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    const double p67 = fma(r, 1.0 / 5040.0, 1.0 / 720.0);
    const double p89 = fma(r, 1.0 / 362880.0, 1.0 / 40320.0);
    const double pab = fma(r, 1.0 / 39916800.0, 1.0 / 3628800.0);
    const double pcd = fma(r, 1.0 / 6227020800.0, 1.0 / 479001600.0);
    const double r4 = r2 * r2;
    const double q0 = fma(p23, r2, p01), q1 = fma(p67, r2, p45), q2 = fma(pab, r2, p89);
    const double r8 = r4 * r4;
    const double s0 = fma(q1, r4, q0), s1 = fma(pcd, r4, q2);
    const double p = fma(s1, r8, s0);
    const int kc = max(-1022, min(1023, k));
    double y = p * __hiloint2double((kc + 1023) << 20, 0);
    if (x < -708.0) y = 0.0;
    if (x > 709.7) y = __longlong_as_double(0x7ff0000000000000LL);
    return y;
}

// exp(x) for the chain.  |x| is ~1e-3 here (dtc/8 * etaDot) when the thermostats are near equilibrium; for
// |x| <= 2^-5 a degree-9 Taylor polynomial is exact to < 0.25 ulp before the final rounding, which puts it in
// the same <= 1 ulp class as CUDA's exp() and glibc's at a fraction of the dependent-instruction depth.  The
// chain is strictly serial (40 sub-steps x ~4 dependent exps per step) and sits between two streaming kernels,
// so its latency is directly visible in the step time.  FAST = false is the full-range version.
template <bool FAST>
__device__ __forceinline__ double chain_exp(double x, bool& outOfRange) {
    if (!FAST) return exp_full(x);
    outOfRange |= fabs(x) > 0.03125;
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

struct ChainConst {
    double dtc, dtc2, dtc4, dtc8, kTl, nkbt, invQ0;
    bool live;
};

// One of the S sub-steps of a thermostat's half-step update (the body of the `iter` loops at
// CudaDrudeTGNHKernels.cpp:565-593 / :606-642).  Returns false when the FAST polynomial left its range.
template <int NL, bool FAST>
__device__ __forceinline__ bool chain_substep(const ChainConst& k, const double (&Q)[NL], const double (&invQ)[NL], double (&eta)[NL],
                                              double (&ed)[NL], double (&edd)[NL], double (&ef)[NL], double& ke, double& scale) {
    bool bad = false;
#pragma unroll
    for (int i = NL - 1; i >= 0; i--) {
        if (i < NL - 1) ef[i] = chain_exp<FAST>(-k.dtc8 * ed[i + 1], bad);
        ed[i] *= ef[i]; ed[i] += edd[i] * k.dtc4; ed[i] *= ef[i];
    }
    scale *= chain_exp<FAST>(-k.dtc2 * ed[0], bad);
    ke *= chain_exp<FAST>(-k.dtc * ed[0], bad);
#pragma unroll
    for (int i = 0; i < NL; i++) eta[i] += k.dtc2 * ed[i];
    if (k.live) edd[0] = (ke - k.nkbt) * k.invQ0;
    ed[0] *= ef[0]; ed[0] += edd[0] * k.dtc4; ed[0] *= ef[0];
#pragma unroll
    for (int i = 1; i < NL; i++) {
        ed[i] *= ef[i];
        edd[i] = (Q[i - 1] * ed[i - 1] * ed[i - 1] - k.kTl) * invQ[i];
        ed[i] += edd[i] * k.dtc4; ed[i] *= ef[i];
    }
    return !bad;
}

// One thermostat's half-step chain update; restates CudaDrudeTGNHKernels.cpp:560-595 (relative and COM
// groups) and :597-642 (Drude group) as one loop nest: the Drude group differs only in the number of
// live links (1 unless useDrudeNHChains) and in kT.  Returns the velocity scale factor.
//
// Same arithmetic as the reference with latency cuts that do not change any value beyond the last ulp:
//   * the factor exp(-dtc/8 * etaDot[top+1]) of the top live link is loop-invariant (etaDot[M] is the
//     permanent zero) and computed once;
//   * the second (upward) sweep re-evaluates exp(-dtc/8 * etaDot[i+1]) on values the first sweep left
//     unchanged, so the first sweep's factors are reused (":583" already reuses a stale expfac this way);
//   * divisions by the constant thermostat masses become multiplications by their reciprocals;
//   * every sub-step first runs branch-free with the small-argument exp polynomial; only if an argument
//     left the polynomial's range is the sub-step redone from the saved state with the full-range exp_full().
// Deviation (documented in DESIGN.md): the Q0 > 0 guard of :561 is applied to the Drude group too, so a
// system without Drude pairs yields scale 1 instead of NaN.
template <int NL>  // live links, known at compile time so that the state sits in registers
__device__ __forceinline__ double chain_update(const ChainView& c, int g, double ke) {
    const int M = c.M;
    const bool isDrude = (g == c.T - 1);
    ChainConst k;
    k.dtc = c.dtc; k.dtc2 = k.dtc / 2.0; k.dtc4 = k.dtc / 4.0; k.dtc8 = k.dtc / 8.0;
    k.kTl = isDrude ? c.kTD : c.kT;
    k.nkbt = c.nkbt[g];
    double Q[NL], invQ[NL], eta[NL], ed[NL], edd[NL], ef[NL];
#pragma unroll
    for (int i = 0; i < NL; i++) {
        Q[i] = c.etaMass[g * M + i];
        invQ[i] = c.invEtaMass[g * M + i];
        eta[i] = c.eta[g * M + i];
        ed[i] = c.etaDot[g * (M + 1) + i];
        edd[i] = c.etaDotDot[g * M + i];
        ef[i] = 1.0;
    }
    // etaDot[M] is the permanent zero; for a Drude group without chains etaDot[1] may hold user-set state
    ef[NL - 1] = exp(-k.dtc8 * c.etaDot[g * (M + 1) + NL]);
    k.live = Q[0] > 0;
    k.invQ0 = k.live ? invQ[0] : 0.0;
    double scale = 1.0;
    if (k.live) edd[0] = (ke - k.nkbt) * k.invQ0;
    // The lanes of the warp run in lockstep, so a lane that needs the full-range exp makes everybody pay for it:
    // the choice is therefore made once for all converged lanes and is sticky for the rest of this update.
    bool full = false;
    for (int iter = 0; iter < c.S; iter++) {
        if (!full) {
            double eta0[NL], ed0[NL], edd0[NL], ef0[NL];
            const double ke0 = ke, scale0 = scale;
#pragma unroll
            for (int i = 0; i < NL; i++) { eta0[i] = eta[i]; ed0[i] = ed[i]; edd0[i] = edd[i]; ef0[i] = ef[i]; }
            const bool ok = chain_substep<NL, true>(k, Q, invQ, eta, ed, edd, ef, ke, scale);
            full = __any_sync(__activemask(), !ok);
            if (!full) continue;
#pragma unroll
            for (int i = 0; i < NL; i++) { eta[i] = eta0[i]; ed[i] = ed0[i]; edd[i] = edd0[i]; ef[i] = ef0[i]; }
            ke = ke0; scale = scale0;
        }
        chain_substep<NL, false>(k, Q, invQ, eta, ed, edd, ef, ke, scale);
    }
#pragma unroll
    for (int i = 0; i < NL; i++) {
        c.eta[g * M + i] = eta[i];
        c.etaDot[g * (M + 1) + i] = ed[i];
        c.etaDotDot[g * M + i] = edd[i];
    }
    return scale;
}

// runtime chain length (M > 4): same algorithm on local arrays, library exp
__device__ __noinline__ double chain_update_generic(const ChainView& c, int g, double ke, int nl) {
    const int M = c.M;
    const bool isDrude = (g == c.T - 1);
    const double kTl = isDrude ? c.kTD : c.kT;
    const double dtc = c.dtc, dtc2 = dtc / 2.0, dtc4 = dtc / 4.0, dtc8 = dtc / 8.0;
    double Q[MAX_M], invQ[MAX_M], eta[MAX_M], ed[MAX_M], edd[MAX_M], ef[MAX_M];
    for (int i = 0; i < nl; i++) {
        Q[i] = c.etaMass[g * M + i];
        invQ[i] = c.invEtaMass[g * M + i];
        eta[i] = c.eta[g * M + i];
        ed[i] = c.etaDot[g * (M + 1) + i];
        edd[i] = c.etaDotDot[g * M + i];
    }
    ef[nl - 1] = exp(-dtc8 * c.etaDot[g * (M + 1) + nl]);
    const double nkbt = c.nkbt[g];
    const bool live = Q[0] > 0;
    const double invQ0 = live ? invQ[0] : 0.0;
    double scale = 1.0;
    if (live) edd[0] = (ke - nkbt) * invQ0;
    for (int iter = 0; iter < c.S; iter++) {
        for (int i = nl - 1; i >= 0; i--) {
            if (i < nl - 1) ef[i] = exp(-dtc8 * ed[i + 1]);
            ed[i] *= ef[i]; ed[i] += edd[i] * dtc4; ed[i] *= ef[i];
        }
        scale *= exp(-dtc2 * ed[0]);
        ke *= exp(-dtc * ed[0]);
        for (int i = 0; i < nl; i++) eta[i] += dtc2 * ed[i];
        if (live) edd[0] = (ke - nkbt) * invQ0;
        ed[0] *= ef[0]; ed[0] += edd[0] * dtc4; ed[0] *= ef[0];
        for (int i = 1; i < nl; i++) {
            ed[i] *= ef[i];
            edd[i] = (Q[i - 1] * ed[i - 1] * ed[i - 1] - kTl) * invQ[i];
            ed[i] += edd[i] * dtc4; ed[i] *= ef[i];
        }
    }
    for (int i = 0; i < nl; i++) {
        c.eta[g * M + i] = eta[i];
        c.etaDot[g * (M + 1) + i] = ed[i];
        c.etaDotDot[g * M + i] = edd[i];
    }
    return scale;
}

__device__ __forceinline__ double chain_update_any(const ChainView& c, int g, double ke) {
    const int nl = (g == c.T - 1 && !c.useDrudeNH) ? 1 : c.M;
    switch (nl) {
        case 1: return chain_update<1>(c, g, ke);
        case 2: return chain_update<2>(c, g, ke);
        case 3: return chain_update<3>(c, g, ke);
        case 4: return chain_update<4>(c, g, ke);
        default: return chain_update_generic(c, g, ke, nl);
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
