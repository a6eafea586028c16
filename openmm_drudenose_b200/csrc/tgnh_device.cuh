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
// 128-bit store with an explicit L2 eviction policy
__device__ __forceinline__ void st_global_hint(float4* p, float4 v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
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
    double expA[3][10]; // exp(c y) = sum_k expA[j][k] y^k for c = -dtc/8, -dtc/2, -dtc (j = 0, 1, 2): c^k / k!, filled by chain_exp_tables
    double expLim[3];   // |y| <= expLim[j]  <=>  |c y| <= 2^-5, the range of the polynomial
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
    double* expHint;    // [T]  1: the last chain update of this thermostat left the range of the short exp polynomial
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

// exp(c*y) for the chain, c one of the three constants -dtc/8, -dtc/2, -dtc.  |c*y| is ~1e-3 when the thermostats are near
// equilibrium; for |c*y| <= 2^-5 a degree-9 Taylor polynomial is exact to < 0.25 ulp before the final rounding, which puts it in
// the same <= 1 ulp class as CUDA's exp() and glibc's at a fraction of the dependent-instruction depth.  The chain is strictly
// serial (40 sub-steps x 3 dependent exps per step for M = 3) and sits between two streaming kernels, so its latency is directly
// visible in the step time.  The constant is folded into the coefficients (a_k = c^k / k!, ChainView::expA), so the polynomial is
// evaluated in y itself: Estrin form, 4 dependent operations from y to the result.
__host__ __device__ inline void chain_exp_tables(ChainView& c) {
    const double invFact[10] = {1.0, 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0, 1.0 / 362880.0};
    const double cs[3] = {-c.dtc / 8.0, -c.dtc / 2.0, -c.dtc};
    for (int j = 0; j < 3; j++) {
        double ck = 1.0;
        for (int k = 0; k < 10; k++) { c.expA[j][k] = ck * invFact[k]; ck *= cs[j]; }
        c.expLim[j] = cs[j] != 0.0 ? 0.03125 / (cs[j] < 0 ? -cs[j] : cs[j]) : 1e300;
    }
}
// (the coefficients sit in the kernel parameters, i.e. in the constant bank: they cost neither registers nor loads)
template <bool FAST>
__device__ __forceinline__ double chain_exp(const double (&a)[10], double lim, double c, double y, bool& outOfRange) {
    outOfRange |= !(fabs(y) <= lim);
    if (!FAST) return exp_full(c * y);
    const double y2 = y * y;
    const double p01 = fma(y, a[1], a[0]);
    const double p23 = fma(y, a[3], a[2]);
    const double p45 = fma(y, a[5], a[4]);
    const double p67 = fma(y, a[7], a[6]);
    const double p89 = fma(y, a[9], a[8]);
    const double y4 = y2 * y2;
    const double q0 = fma(p23, y2, p01);
    const double q1 = fma(p67, y2, p45);
    const double y8 = y4 * y4;
    return fma(p89, y8, fma(q1, y4, q0));
}

template <int NL>
struct ChainConst {
    double dtc2, dtc4, dtc8, inv4;     // dtc/2, dtc/4, dtc/8, 4/dtc
    // g_i = (dtc/4) etaDotDot_i = gA[i] z + gB[i],  z = 2KE (i = 0) or etaDot_{i-1}^2 (i >= 1)
    double gA[NL], gB[NL];
    bool live;                         // Q_0 > 0 (:561)
};
template <int NL>
struct ChainState {
    double eta[NL], ed[NL], g[NL], ef[NL];
};

// One of the S sub-steps of a thermostat's half-step update (the body of the `iter` loops at
// CudaDrudeTGNHKernels.cpp:565-593 / :606-642).  Returns true when an argument left the range of the short polynomial.
//
// The reference updates a link as  v *= a; v += G dt/4; v *= a  with a = exp(-dt/8 v_next) and G = (Q_prev v_prev^2 - kT) / Q.
// The same values are formed here with a shorter dependency chain and fewer operations (a dependent DFMA costs 8.2 cycles on
// B200 and one warp issues a DFMA every 2.1 cycles, scripts/lat.cu, scripts/lat_ilp.cu):
//   downward sweep   v = a * fma(v, a, g)               g = G dt/4, known since the previous upward sweep
//   upward sweep     v = fma(g, a, v a^2)               v a^2 formed while g is still on its way
//                    g = fma(z, gA, gB)                 (Q_prev / Q)(dt/4) z - (kT / Q)(dt/4): one operation after z = v_prev^2 (or 2KE)
//   scale factor     s = exp(-dt/2 v_0);  scale *= s;  2KE *= s^2      (the reference evaluates exp(-dt v_0) separately, :575-576)
//   etaDotDot        = g / (dt/4), formed once after the last sub-step
// Each rearrangement is exact algebra; results differ from the reference's operation order in the last bits only (tests: chain
// variables against the oracle after 1 and 1000 steps).
template <int NL, bool FAST>
__device__ __forceinline__ bool chain_substep(const ChainView& c, const ChainConst<NL>& k, ChainState<NL>& st, double& ke, double& scale) {
    bool bad = false;
    double ef2[NL];
#pragma unroll
    for (int i = NL - 1; i >= 0; i--) {
        if (i < NL - 1) st.ef[i] = chain_exp<FAST>(c.expA[0], c.expLim[0], -k.dtc8, st.ed[i + 1], bad);
        ef2[i] = st.ef[i] * st.ef[i];
        st.ed[i] = st.ef[i] * fma(st.ed[i], st.ef[i], st.g[i]);
    }
    const double sf = chain_exp<FAST>(c.expA[1], c.expLim[1], -k.dtc2, st.ed[0], bad);
    scale *= sf;
    ke *= sf * sf;
#pragma unroll
    for (int i = 0; i < NL; i++) st.eta[i] = fma(k.dtc2, st.ed[i], st.eta[i]);
    if (k.live) st.g[0] = fma(ke, k.gA[0], k.gB[0]);
    st.ed[0] = fma(st.g[0], st.ef[0], st.ed[0] * ef2[0]);
#pragma unroll
    for (int i = 1; i < NL; i++) {
        st.g[i] = fma(st.ed[i - 1] * st.ed[i - 1], k.gA[i], k.gB[i]);
        st.ed[i] = fma(st.g[i], st.ef[i], st.ed[i] * ef2[i]);
    }
    return bad;
}

// One thermostat's half-step chain update on state held in registers; restates CudaDrudeTGNHKernels.cpp:560-595 (relative and
// COM groups) and :597-642 (Drude group) as one loop nest: the Drude group differs only in the number of live links (1 unless
// useDrudeNHChains) and in kT.  Returns the velocity scale factor.
//
// Same arithmetic as the reference with latency cuts that do not change any value beyond the last bits:
//   * the factor exp(-dtc/8 * etaDot[top+1]) of the top live link is loop-invariant (etaDot[M] is the
//     permanent zero) and computed once;
//   * the second (upward) sweep re-evaluates exp(-dtc/8 * etaDot[i+1]) on values the first sweep left
//     unchanged, so the first sweep's factors are reused (":583" already reuses a stale expfac this way);
//   * divisions by the constant thermostat masses become multiplications by their reciprocals, folded with dtc/4;
//   * the link updates in the regrouped form described at chain_substep;
//   * all S sub-steps run as one straight dependent chain with the small-argument polynomial (no vote, no branch, no state copy
//     between sub-steps); the arguments are checked on the side and, if one left the polynomial's range, the whole update is
//     redone from the saved state with the full-range exp.  A thermostat that needed the full range last time starts there
//     (`hint`, kept in c.expHint between launches: e.g. the first picoseconds of an unequilibrated system); the lanes of the
//     warp run in lockstep, so the choice is made for all of them together.
// Deviation (documented in DESIGN.md): the Q0 > 0 guard of :561 is applied to the Drude group too, so a
// system without Drude pairs yields scale 1 instead of NaN.
template <int NL>
__device__ __forceinline__ double chain_update(const ChainView& c, const ChainConst<NL>& k, ChainState<NL>& st, double ke, bool& hint) {
    if (k.live) st.g[0] = fma(ke, k.gA[0], k.gB[0]);             // etaDotDot_0 = (2KE - N kT) / Q_0 (:563)
    double scale = 1.0;
    bool full = __any_sync(__activemask(), hint);
    bool bad = false;
    if (!full) {
        const ChainState<NL> st0 = st;
        const double ke0 = ke;
#pragma unroll 1
        for (int iter = 0; iter < c.S; iter++) bad |= chain_substep<NL, true>(c, k, st, ke, scale);
        full = __any_sync(__activemask(), bad);
        if (full) { st = st0; ke = ke0; scale = 1.0; }
    }
    if (full) {
        bad = false;
#pragma unroll 1
        for (int iter = 0; iter < c.S; iter++) bad |= chain_substep<NL, false>(c, k, st, ke, scale);
    }
    hint = full && bad;
    return scale;
}

// Thermostat `g` (one lane): state and constants into registers, the one or two chain updates `mode` asks for, state back.
// `liveLinks` < NL: a Drude thermostat without chains (useDrudeNHChains = false) riding along in the warp's NL-link code instead of
// running its own 1-link pass after the others.  Its upper links are frozen at zero chain velocity (chain_phase checks that), so
// with gA = gB = g = 0 for them every factor is exp(0) = 1 and every increment 0: the live link gets bit for bit what the 1-link
// code gives it, and the frozen links are neither changed nor written back.
template <int NL>
__device__ __forceinline__ void chain_lane(const ChainView& c, int mode, int g, int liveLinks, double ke, double& pend, double& used, double& s) {
    const int M = c.M;
    const bool isDrude = (g == c.T - 1);
    ChainConst<NL> k;
    k.dtc2 = c.dtc / 2.0; k.dtc4 = c.dtc / 4.0; k.dtc8 = c.dtc / 8.0; k.inv4 = 4.0 / c.dtc;
    const double kTl = isDrude ? c.kTD : c.kT;
    ChainState<NL> st;
    double edd0 = 0.0;
    k.live = c.etaMass[g * M] > 0;
#pragma unroll
    for (int i = 0; i < NL; i++) {
        const double invQ = c.invEtaMass[g * M + i];
        const double dA = i == 0 ? (k.live ? invQ : 0.0) : c.etaMass[g * M + i - 1] * invQ;
        const double dB = i == 0 ? (k.live ? -c.nkbt[g] * invQ : 0.0) : -kTl * invQ;
        const bool frozen = i >= liveLinks;
        k.gA[i] = frozen ? 0.0 : dA * k.dtc4; k.gB[i] = frozen ? 0.0 : dB * k.dtc4;
        st.eta[i] = c.eta[g * M + i];
        st.ed[i] = c.etaDot[g * (M + 1) + i];
        const double edd = c.etaDotDot[g * M + i];
        if (i == 0) edd0 = edd;
        st.g[i] = frozen ? 0.0 : edd * k.dtc4;
        st.ef[i] = 1.0;
    }
    // etaDot[M] is the permanent zero; for a Drude group without chains etaDot[1] may hold user-set state
    st.ef[NL - 1] = exp(-k.dtc8 * c.etaDot[g * (M + 1) + NL]);
    bool hint = c.expHint[g] != 0.0;
    // update 0: the half-step that ends a step (consumes 2KE, its factor stays pending); update 1: the half-step that begins
    // a step (consumes pending^2 2KE).  One copy of the code for both: this kernel is one warp running a long instruction
    // stream once, and what it fetches it pays for.
    const int first = mode == CHAIN_FIRST ? 1 : 0, last = mode == CHAIN_SECOND ? 0 : 1;
#pragma unroll 1
    for (int u = first; u <= last; u++) {
        const double keIn = u == 0 ? ke : pend * pend * ke;
        used = keIn;
        s = chain_update<NL>(c, k, st, keIn, hint);
        if (u == 0) pend = s;
    }
#pragma unroll
    for (int i = 0; i < NL; i++) {
        if (i >= liveLinks) continue;
        c.eta[g * M + i] = st.eta[i];
        c.etaDot[g * (M + 1) + i] = st.ed[i];
        c.etaDotDot[g * M + i] = (i == 0 && !k.live) ? edd0 : st.g[i] * k.inv4;
    }
    c.expHint[g] = hint ? 1.0 : 0.0;
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

// runtime chain length (M > 4)
__device__ __noinline__ void chain_lane_generic(const ChainView& c, int mode, int g, double ke, int nl, double& pend, double& used, double& s) {
    if (mode == CHAIN_SECOND || mode == CHAIN_SECOND_FIRST) {
        used = ke;
        s = chain_update_generic(c, g, ke, nl);
        pend = s;
    }
    if (mode == CHAIN_FIRST || mode == CHAIN_SECOND_FIRST) {
        const double keEff = pend * pend * ke;
        used = keEff;
        s = chain_update_generic(c, g, keEff, nl);
    }
}

// Executed by ONE warp (lanes 0..T-1 own one thermostat each; T <= 32).  c.ke2 must hold the global
// 2*KE sums of the velocities as stored.
__device__ void chain_phase(const ChainView& c, int mode, int lane) {
    const bool on = lane < c.T;
    const double ke = on ? c.ke2[lane] : 0.0;
    double pend = on ? c.pending[lane] : 1.0;
    double used = 0.0, s = 1.0;
    if (on) {
        const int nl = (lane == c.T - 1 && !c.useDrudeNH) ? 1 : c.M;      // live links of this thermostat (:597-642: the Drude group has one unless useDrudeNHChains)
        int run = nl;                                                      // links of the code path this lane takes
        if (nl < c.M && c.M <= 4) {
            // ride along with the other lanes if the links above the live one are at rest (always, unless a caller set them)
            bool rest = true;
            for (int i = 1; i <= c.M; i++) rest &= c.etaDot[lane * (c.M + 1) + i] == 0.0;
            if (rest) run = c.M;
        }
        switch (run) {
            case 1: chain_lane<1>(c, mode, lane, nl, ke, pend, used, s); break;
            case 2: chain_lane<2>(c, mode, lane, nl, ke, pend, used, s); break;
            case 3: chain_lane<3>(c, mode, lane, nl, ke, pend, used, s); break;
            case 4: chain_lane<4>(c, mode, lane, nl, ke, pend, used, s); break;
            default: chain_lane_generic(c, mode, lane, ke, nl, pend, used, s); break;
        }
    }
    __syncwarp();
    if (mode == CHAIN_FIRST || mode == CHAIN_SECOND_FIRST) {
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
