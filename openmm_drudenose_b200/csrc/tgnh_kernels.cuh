// The three streaming kernels of the TGNH step, as one persistent, TMA-fed template.
//
//   KIND_A  (first half)   = integrateDrudeTGNHChain + integrateDrudeTGNHVelocities(updatePosDelta) +
//                            integrateDrudeTGNHPositions + applyHardWallConstraints
//                            (platforms/cuda/src/kernels/drudeTGNH.cu:249-301, 307-365, 435-466, 471-574)
//   KIND_B  (second half)  = integrateDrudeTGNHVelocities + calcCOMVelocities + normalizeVelocities +
//                            computeNormalizedKineticEnergies + sumNormalizedKineticEnergies
//                            (drudeTGNH.cu:307-365, 82-133, 138-242)
//   KIND_BU                = KIND_B for systems whose residues each lie in ONE temperature group (the normal
//                            case): sum_u m_u |v_u - V|^2 = sum_u m_u |v_u|^2 - M |V|^2 per residue, so particles
//                            never need their residue's COM velocity; one thread per RESIDUE (densely packed
//                            lanes instead of every member redoing the residue sum) adds the M |V|^2 terms
//   KIND_A1 / KIND_A2      = the first half split around OpenMM's position constraints (CudaDrudeTGNHKernels.cpp:356-376):
//                            A1 = scaling + kick, emits posDelta = dt * v (drudeTGNH.cu:249-365); the caller runs
//                            integration.applyConstraints on posDelta; A2 = x += posDelta, v = posDelta / dt, hard wall
//                            (drudeTGNH.cu:435-574)
//   KIND_KE (reduce/flush) = the kinetic-energy reduction alone, optionally applying a pending scaling
//                            (drudeTGNH.cu:82-242, 249-301)
//
// Data movement: every CTA is persistent and walks residue-aligned tiles of <= 512 consecutive
// particles.  One elected thread streams each tile's velm / posq / force / descriptor slices into a
// ring of shared-memory stages with cp.async.bulk (TMA) completing on a "full" mbarrier; warps hand a
// stage back through an "empty" mbarrier, so no CTA-wide barrier sits on the streaming path.  Thread i
// owns particle i of the tile, reads its pair partner and its residue's members out of shared memory
// (that is where pair / residue indexing would break coalescing), and writes its own float4 results
// straight back with fully coalesced 128-bit stores.  No temporary arrays (normVelm, comVelm,
// posDelta, kineticEnergyBuffer of the reference) ever touch HBM.
//
// Arithmetic: the reference transforms each Drude pair to (centre of mass, relative) coordinates and
// back in every kernel.  Algebraically, with f_j = m_j / (m_i + m_j) and rel = v_j - v_i,
//     scaling:  v_i <- sT (v_i - V) + sCOM V + (sT - sDrude) f_j rel        (drudeTGNH.cu:270-300)
//               evaluated as v_i + [(sT-1)(v_i - V) + (sCOM-1) V + (sT - sDrude) f_j rel]: the thermostat factors are
//               1 + O(1e-4), and a factor rounded to fp32 would put the SAME 3e-8 relative error on every particle of
//               the group (6e-8 on its kinetic energy, every step); the differences s-1 are taken in double
//     kick:     v_i <- v_i + fscale w_i F_i                                 (drudeTGNH.cu:330-364; the cm/rel
//                                                                            form reduces to the direct kick)
// which holds for both members of a pair and, with "partner = self" (rel = 0), for ordinary particles:
// one branch-free code path, and no (f_i + f_j != 1) round-trip error in fp32.
#pragma once
#include "tgnh_device.cuh"

namespace tgnh {

enum { KIND_A = 0, KIND_B = 1, KIND_KE = 2, KIND_BU = 3, KIND_A1 = 4, KIND_A2 = 5 };
constexpr int NWARPS = TILE / 32;
constexpr int TLIST_CAP = 128;     // tile bounds cached in shared memory per CTA (later tiles are looked up in global memory)

struct StreamArgs {
    float4* velm;
    float4* posq;
    float4* posDelta;         // float4[paddedN] (dt*v, 0): integration.getPosDelta() (KIND_A1 writes, KIND_A2 reads)
    const void* force;        // SoA [3][paddedN], float or long long
    const uint32_t* desc;     // [roundup4(N)]
    const int* tileStart;     // [numTiles + 1]
    const int* resStart;      // [R + 1] first particle of every residue in particle order, then N   (KIND_BU)
    const int* tileFirstRes;  // [numTiles + 1] index into resStart of each tile's first residue      (KIND_BU)
    int numTiles;
    int paddedN;
    float dt;                 // step size
    float fscale;             // 0.5*dt (f32 forces) or 0.5*dt/2^32 (i64 forces)
    float rmax;               // maxDrudeDistance
    float hardwallScale;      // sqrt(kB * T_drude)
    int applyScale;           // KIND_KE: scale velocities by scaleA and write them back
    int useLocalKE;           // sharded: reduce into chain.ke2Local (all-reduced into ke2 afterwards)
    int reverse;              // walk the tiles from the last to the first (see "L2 hand-over" below)
    int prologuePrefetch;     // tiles per CTA whose read-only inputs are prefetched into L2 before griddepcontrol.wait
    double* partials;         // [gridDim.x][T]
    unsigned int* ticket;     // last-CTA-done counter (self-resetting)
    ChainView chain;
};

template <int KIND, int FFMT>
struct StageLayout {
    static constexpr bool HAS_X = (KIND == KIND_A || KIND == KIND_A2);
    static constexpr bool HAS_F = (KIND != KIND_KE && KIND != KIND_A2);
    static constexpr bool HAS_P = (KIND == KIND_A2);             // posDelta tile
    static constexpr int FBYTES = FFMT == 1 ? 8 : 4;
    static constexpr int OFF_V = 0;
    static constexpr int OFF_X = OFF_V + TILE * 16;
    static constexpr int OFF_F = OFF_X + (HAS_X ? TILE * 16 : 0);
    static constexpr int OFF_D = OFF_F + (HAS_F ? 3 * PADW * FBYTES : 0);
    static constexpr bool HAS_R = (KIND == KIND_BU);
    static constexpr int OFF_R = OFF_D + PADW * 4;              // residue starts of the tile (KIND_BU)
    static constexpr int OFF_P = OFF_R + (HAS_R ? PADW * 4 : 0);
    static constexpr int OFF_HDR = OFF_P + (HAS_P ? TILE * 16 : 0);  // int4 {first particle, count, first residue, residues}
    static constexpr int BYTES = OFF_HDR + 16;
};

template <int KIND, int FFMT, bool USE_COM>
struct SmemLayout {
    using Stage = StageLayout<KIND, FFMT>;
    static constexpr bool HAS_KE = (KIND == KIND_B || KIND == KIND_BU || KIND == KIND_KE);
    static constexpr int NSTAGE = (KIND == KIND_A || KIND == KIND_A2) ? 3 : 4;
    static constexpr int OFF_BAR = NSTAGE * Stage::BYTES;         // full[NS], empty[NS] mbarriers
    static constexpr int OFF_TLIST = OFF_BAR + 128;               // int4[TLIST_CAP] bounds of this CTA's first tiles
    static constexpr int OFF_SCALE = OFF_TLIST + TLIST_CAP * 16;  // double[MAX_T] s^2, float[MAX_T] s - 1
    static constexpr int OFF_MISC = OFF_SCALE + MAX_T * 12;       // int[4]
    static constexpr int OFF_WARP = OFF_MISC + 16;                // double[T][16]   (HAS_KE)
    static constexpr int OFF_KE = OFF_WARP + (HAS_KE ? MAX_T * 16 * 8 : 0);
    static int bytes(int T) { return OFF_KE + (HAS_KE ? T * TILE * 8 : 0); }
};

template <int FFMT>
__device__ __forceinline__ float3 load_force3(const unsigned char* sF, int idx) {
    if (FFMT == 1) {
        const long long* f = reinterpret_cast<const long long*>(sF);
        return make_float3((float)f[idx], (float)f[PADW + idx], (float)f[2 * PADW + idx]);
    }
    const float* f = reinterpret_cast<const float*>(sF);
    return make_float3(f[idx], f[PADW + idx], f[2 * PADW + idx]);
}

// approximate reciprocal / rsqrt: 1 MUFU each, <= 1 ulp-class error; exact IEEE division would cost ~10
// instructions plus a divergent slow path in kernels that are issue-bound, not DRAM-bound, without it
__device__ __forceinline__ float rcp_fast(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// second-order term of the reciprocal: 1/w = r + rcp_lo(w, r) to ~1e-14 for r = rcp_fast(w)
__device__ __forceinline__ float rcp_lo(float w, float r) { return r * fmaf(-w, r, 1.0f); }
// Masses that weight the kinetic-energy sums are formed in double from that pair.  A species' rounded fp32
// mass (all oxygens share one w) would otherwise shift its thermostat's energy by up to 6e-8 — systematically,
// so it does not average out over particles, and the Nose-Hoover chain integrates it.
__device__ __forceinline__ double mass_d(float w, float r) { return (double)r + (double)rcp_lo(w, r); }
// 1/x in double from a float seed: two Newton steps (1e-7 -> 1e-14 -> rounding)
__device__ __forceinline__ double rcp_d(double x, float seed) {
    double r = (double)seed;
    r = r * fma(-x, r, 2.0);
    return r * fma(-x, r, 2.0);
}
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// applyHardWallConstraints (drudeTGNH.cu:487-572) for one pair that is beyond the wall.
// 1 = Drude particle, 2 = parent.  `delta` is x1 - x2, r2 its squared length.
__device__ __forceinline__ void hard_wall(float3 delta, float r2, float3& x1, float3& x2, float3& v1, float3& v2, float w1, float w2,
                                       float rmax, float hardwallScale, float dt) {
    const float rInv = rsqrt_fast(r2);
    const float r = r2 * rInv;
    const float3 bondDir = make_float3(delta.x * rInv, delta.y * rInv, delta.z * rInv);
    const float mass1 = rcp_fast(w1);
    const float deltaR = r - rmax;
    float deltaT = dt;
    float dotvr1 = v1.x * bondDir.x + v1.y * bondDir.y + v1.z * bondDir.z;
    const float3 vp1 = make_float3(v1.x - bondDir.x * dotvr1, v1.y - bondDir.y * dotvr1, v1.z - bondDir.z * dotvr1);
    const float vBond = hardwallScale * sqrtf(w1);                    // hardwallscaleDrude / SQRT(mass1)
    if (w2 == 0.0f) {
        // massless parent: only the Drude particle moves (:504-526)
        if (dotvr1 != 0.0f) deltaT = __fdividef(deltaR, fabsf(dotvr1));
        if (deltaT > dt) deltaT = dt;
        dotvr1 = -copysignf(vBond, dotvr1);                           // -dotvr1*scale/(|dotvr1|*sqrt(m1))
        const float dr = -deltaR + deltaT * dotvr1;
        x1.x += bondDir.x * dr; x1.y += bondDir.y * dr; x1.z += bondDir.z * dr;
        v1 = make_float3(vp1.x + bondDir.x * dotvr1, vp1.y + bondDir.y * dotvr1, vp1.z + bondDir.z * dotvr1);
    } else {
        const float mass2 = rcp_fast(w2);
        const float invTotalMass = rcp_fast(mass1 + mass2);
        float dotvr2 = v2.x * bondDir.x + v2.y * bondDir.y + v2.z * bondDir.z;
        const float3 vp2 = make_float3(v2.x - bondDir.x * dotvr2, v2.y - bondDir.y * dotvr2, v2.z - bondDir.z * dotvr2);
        const float vbCMass = (mass1 * dotvr1 + mass2 * dotvr2) * invTotalMass;
        dotvr1 -= vbCMass;
        dotvr2 -= vbCMass;
        if (dotvr1 != dotvr2) deltaT = __fdividef(deltaR, fabsf(dotvr1 - dotvr2));
        if (deltaT > dt) deltaT = dt;
        dotvr1 = -copysignf(vBond * mass2 * invTotalMass, dotvr1);    // :542  (-dotvr1*vBond*m2/M/|dotvr1|)
        dotvr2 = -copysignf(vBond * mass1 * invTotalMass, dotvr2);    // :543
        const float dr1 = -deltaR * mass2 * invTotalMass + deltaT * dotvr1;
        const float dr2 = deltaR * mass1 * invTotalMass + deltaT * dotvr2;
        dotvr1 += vbCMass;
        dotvr2 += vbCMass;
        x1.x += bondDir.x * dr1; x1.y += bondDir.y * dr1; x1.z += bondDir.z * dr1;
        x2.x += bondDir.x * dr2; x2.y += bondDir.y * dr2; x2.z += bondDir.z * dr2;
        v1 = make_float3(vp1.x + bondDir.x * dotvr1, vp1.y + bondDir.y * dotvr1, vp1.z + bondDir.z * dotvr1);
        v2 = make_float3(vp2.x + bondDir.x * dotvr2, vp2.y + bondDir.y * dotvr2, vp2.z + bondDir.z * dotvr2);
    }
}

__device__ __forceinline__ float dot3(float3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }

// v + eT*r + eCOM*V + c*rel (e = s - 1) with a fixed evaluation order, so that the two threads of a Drude pair compute
// bit-identical values for each other's particle (both must take the same side of the hard-wall test)
__device__ __forceinline__ float3 scaled_velocity(float3 v, float eT, float3 r, float eCOM, float3 V, float c, float3 rel) {
    return make_float3(v.x + fmaf(c, rel.x, fmaf(eT, r.x, eCOM * V.x)), v.y + fmaf(c, rel.y, fmaf(eT, r.y, eCOM * V.y)),
                       v.z + fmaf(c, rel.z, fmaf(eT, r.z, eCOM * V.z)));
}
__device__ __forceinline__ float3 kicked(float3 v, float fw, float3 F) {
    return make_float3(fmaf(fw, F.x, v.x), fmaf(fw, F.y, v.y), fmaf(fw, F.z, v.z));
}

// L2 hand-over: velm is written by one streaming launch and read (then overwritten) by the next.  B200's L2
// holds 126 MB, so when consecutive launches walk the tiles in opposite directions the tail of what the
// previous launch wrote is still resident: those reads never reach HBM and the dirty lines are overwritten in
// L2 before they are evicted.  velm stores therefore use the default L2 policy while everything that is touched
// once per step (posq, forces, descriptors) is loaded evict-first / stored streaming.
template <int KIND, int FFMT, bool USE_COM, bool HARDWALL>
__global__ void __launch_bounds__(TILE, 2) tgnh_stream_kernel(const __grid_constant__ StreamArgs a) {
    using L = SmemLayout<KIND, FFMT, USE_COM>;
    using St = typename L::Stage;
    constexpr int NS = L::NSTAGE;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* empty = full + NS;
    int4* tlist = reinterpret_cast<int4*>(smem + L::OFF_TLIST);
    double* ssq = reinterpret_cast<double*>(smem + L::OFF_SCALE);         // s_g^2
    float* seps = reinterpret_cast<float*>(ssq + MAX_T);                  // s_g - 1
    int* smisc = reinterpret_cast<int*>(smem + L::OFF_MISC);
    double* ske = reinterpret_cast<double*>(smem + L::OFF_KE);
    double* swarp = reinterpret_cast<double*>(smem + L::OFF_WARP);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.chain.T, G = a.chain.G;
    const int myTiles = blockIdx.x < a.numTiles ? (a.numTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // bounds of this CTA's it-th tile: {first particle, end particle, first residue, end residue}
    auto tile_bounds = [&](int it) {
        const int t = blockIdx.x + it * gridDim.x;
        const int tile = a.reverse ? a.numTiles - 1 - t : t;
        return make_int4(a.tileStart[tile], a.tileStart[tile + 1], St::HAS_R ? a.tileFirstRes[tile] : 0, St::HAS_R ? a.tileFirstRes[tile + 1] : 0);
    };

    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWARPS); }
        fence_mbar_init();
    }
    for (int it = tid; it < myTiles && it < TLIST_CAP; it += TILE) tlist[it] = tile_bounds(it);   // static tables: safe before pdl_wait
    if (L::HAS_KE)
        for (int g = 0; g < T; g++) ske[g * TILE + tid] = 0.0;
    __syncthreads();

    const uint64_t polOnce = policy_evict_first();
    // L2 prefetch of a later tile's read-only inputs (posq is only touched by first-half launches, forces and
    // descriptors are never written by this library): used in the prologue, while HBM is idle during the chain launch
    auto prefetch = [&](int it) {
        if (it >= myTiles) return;
        const int4 b = it < TLIST_CAP ? tlist[it] : tile_bounds(it);
        const int start = b.x, end = b.y;
        const int n = end - start, a0 = start & ~3, na = ((end + 3) & ~3) - a0;
        if (St::HAS_X) bulk_prefetch_l2(a.posq + start, n * 16);
        if (St::HAS_F) {
            const unsigned char* f = static_cast<const unsigned char*>(a.force);
            for (int c = 0; c < 3; c++) bulk_prefetch_l2(f + ((size_t)c * a.paddedN + a0) * St::FBYTES, na * St::FBYTES);
        }
        bulk_prefetch_l2(a.desc + a0, na * 4);
    };
    // Request tile `it` of this CTA into its stage (thread 0 only).  `parts`: 1 = everything but velm (arms the barrier
    // with the full byte count), 2 = velm, 3 = both.  The split exists for the prologue: velm is the only input the
    // previous launch writes, so the rest can be requested before griddepcontrol.wait, while the chain launch that
    // precedes a first-half launch is still running.
    auto issue = [&](int it, int parts) {
        const int4 b = it < TLIST_CAP ? tlist[it] : tile_bounds(it);
        const int start = b.x, end = b.y, r0 = b.z, r1 = b.w;
        const int n = end - start, a0 = start & ~3, na = ((end + 3) & ~3) - a0;
        unsigned char* st = smem + (it % NS) * St::BYTES;
        uint64_t* bar = &full[it % NS];
        if (parts & 1) {
            *reinterpret_cast<int4*>(st + St::OFF_HDR) = make_int4(start, n, r0, r1 - r0);
            uint32_t bytes = n * 16 + na * 4;
            const int ra0 = r0 & ~3, rna = ((r1 + 1 + 3) & ~3) - ra0;     // residues r0..r1 inclusive (r1 = end marker)
            if (St::HAS_R) bytes += rna * 4;
            if (St::HAS_X) bytes += n * 16;
            if (St::HAS_P) bytes += n * 16;
            if (St::HAS_F) bytes += 3 * na * St::FBYTES;
            mbar_arrive_expect_tx(bar, bytes);
            if (St::HAS_X) bulk_g2s(st + St::OFF_X, a.posq + start, n * 16, bar, polOnce);
            if (St::HAS_F) {
                const unsigned char* f = static_cast<const unsigned char*>(a.force);
                for (int c = 0; c < 3; c++)
                    bulk_g2s(st + St::OFF_F + c * PADW * St::FBYTES, f + ((size_t)c * a.paddedN + a0) * St::FBYTES, na * St::FBYTES, bar,
                             polOnce);
            }
            bulk_g2s(st + St::OFF_D, a.desc + a0, na * 4, bar, polOnce);
            if (St::HAS_R) bulk_g2s(st + St::OFF_R, a.resStart + ra0, rna * 4, bar, polOnce);
        }
        if (parts & 2) {
            bulk_g2s(st + St::OFF_V, a.velm + start, n * 16, bar, polOnce);
            if (St::HAS_P) bulk_g2s(st + St::OFF_P, a.posDelta + start, n * 16, bar, polOnce);   // written by the caller's constraint kernels
        }
    };
    if (tid == 0) {
        for (int it = 0; it < NS && it < myTiles; it++) issue(it, 1);
        for (int it = NS; it < NS + a.prologuePrefetch; it++) prefetch(it);
    }
    pdl_wait();                                         // everything below reads what earlier launches wrote
    if (tid == 0)
        for (int it = 0; it < NS && it < myTiles; it++) issue(it, 2);
    if (tid < T) {
        const double sg = (KIND == KIND_B || KIND == KIND_BU || KIND == KIND_A2) ? 1.0 : a.chain.scaleA[tid];
        ssq[tid] = sg * sg;
        seps[tid] = (float)(sg - 1.0);
    }
    __syncthreads();

    const bool doScale = (KIND == KIND_A || KIND == KIND_A1) || (KIND == KIND_KE && a.applyScale);
    const float eCOM = doScale ? seps[G] : 0.0f;
    const float eDrude = doScale ? seps[G + 1] : 0.0f;
    const float rmax2 = a.rmax * a.rmax;
    double accCOM = 0.0, accDrude = 0.0;               // this thread's share of the COM-group and Drude-group sums

    for (int it = 0; it < myTiles; it++) {
        const int stg = it % NS;
        const uint32_t phase = (it / NS) & 1;
        unsigned char* st = smem + stg * St::BYTES;
        const float4* sv = reinterpret_cast<const float4*>(st + St::OFF_V);
        const float4* sx = reinterpret_cast<const float4*>(st + St::OFF_X);
        const unsigned char* sF = st + St::OFF_F;
        const uint32_t* sd = reinterpret_cast<const uint32_t*>(st + St::OFF_D);

        mbar_wait(&full[stg], phase);
        const int4 hdr = *reinterpret_cast<const int4*>(st + St::OFF_HDR);
        const int start = hdr.x, n = hdr.y;
        const int fo = start & 3;                     // offset of the tile inside its 4-aligned window

        const bool active = tid < n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t d = 0;                               // inactive lanes: ordinary particle, partner = self, residue = self
        if (active) { v = sv[tid]; d = sd[fo + tid]; }
        const int tg = desc_tg(d);
        const uint32_t role = desc_role(d);
        const bool massive = v.w != 0.0f;
        const int pj = tid + desc_partner(d);         // ordinary particles: partner == self
        const float4 vj = active ? sv[pj] : v;
        const float m = massive ? rcp_fast(v.w) : 0.0f;
        const float mj = rcp_fast(vj.w);
        const float invTot = rcp_fast(m + mj);
        const float fi = m * invTot, fj = mj * invTot;        // mass fractions of this particle and of its partner
        float3 rel = make_float3(vj.x - v.x, vj.y - v.y, vj.z - v.z);   // partner minus me (0 for ordinary particles)
        const float eT = doScale ? seps[tg] : 0.0f;
        const float coef = (eT - eDrude);            // sT - sDrude

        float3 F = make_float3(0.f, 0.f, 0.f), Fj = F;
        if (St::HAS_F && active) { F = load_force3<FFMT>(sF, fo + tid); Fj = load_force3<FFMT>(sF, fo + pj); }
        const float fw = a.fscale * v.w, fwj = a.fscale * vj.w;

        // residue centre-of-mass velocity (calcCOMVelocities, drudeTGNH.cu:86-105); for the second half it is taken
        // after the kick: sum_j m_j (v_j + fscale w_j F_j) = sum_j (m_j v_j + fscale F_j) over the massive members
        float3 V = make_float3(0.f, 0.f, 0.f);
        double keC = 0.0;                             // M |V|^2 of this particle's residue (kinds that reduce energies)
        if (USE_COM && active && KIND != KIND_BU && KIND != KIND_A2) {
            const int j0 = tid - desc_off_first(d), j1 = tid + desc_off_last(d);
            if (!L::HAS_KE) {
                // first half: V only feeds the small corrections (sT-1)(v - V) and (sCOM-1) V, fp32 sums are ample
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int j = j0; j <= j1; j++) {
                    const float4 q = sv[j];
                    const float mq = q.w != 0.0f ? rcp_fast(q.w) : 0.0f;
                    acc.x = fmaf(q.x, mq, acc.x); acc.y = fmaf(q.y, mq, acc.y); acc.z = fmaf(q.z, mq, acc.z);
                    acc.w += mq;
                }
                const float inv = rcp_fast(acc.w);
                V = make_float3(acc.x * inv, acc.y * inv, acc.z * inv);
            } else {
                // kinds that reduce energies: momentum and mass of the residue in double (see mass_d)
                double Px = 0.0, Py = 0.0, Pz = 0.0, M = 0.0;
                for (int j = j0; j <= j1; j++) {
                    const float4 q = sv[j];
                    const bool mass = q.w != 0.0f;
                    const float mq = mass ? rcp_fast(q.w) : 0.0f;
                    const double mqd = mass ? mass_d(q.w, mq) : 0.0;
                    float3 vq = make_float3(q.x, q.y, q.z);
                    if (KIND == KIND_B) vq = kicked(vq, a.fscale * q.w, load_force3<FFMT>(sF, fo + j));   // what member j stores
                    Px = fma(mqd, (double)vq.x, Px); Py = fma(mqd, (double)vq.y, Py); Pz = fma(mqd, (double)vq.z, Pz);
                    M += mqd;
                }
                const double invM = rcp_d(M, rcp_fast((float)M));
                const double Vx = Px * invM, Vy = Py * invM, Vz = Pz * invM;
                V = make_float3((float)Vx, (float)Vy, (float)Vz);
                keC = M * fma(Vx, Vx, fma(Vy, Vy, Vz * Vz));
            }
        }

        float3 vn;                                    // this particle's new velocity
        float3 r;                                     // ... relative to the residue (after scaling / kick)
        if (KIND == KIND_BU) {
            // kick (drudeTGNH.cu:314-364); kinetic energies in the lab frame, the residues' M |V|^2 is removed below
            vn = kicked(make_float3(v.x, v.y, v.z), fw, F);
            if (active && massive) st_global(a.velm + start + tid, make_float4(vn.x, vn.y, vn.z, v.w));
            const float3 vjn = kicked(make_float3(vj.x, vj.y, vj.z), fwj, Fj);
            rel = make_float3(vjn.x - vn.x, vjn.y - vn.y, vjn.z - vn.z);
            r = vn;
            // one thread per residue of the tile; the duty rotates over the warps from tile to tile so that no warp is
            // always the slow one (a stage is recycled only when all 16 warps have left it)
            const int ridx = (tid - it * 128) & (TILE - 1);
            if (USE_COM && ridx < hdr.w) {
                // P = sum_j m_j v_j (kicked velocities), M = sum_j m_j over the residue's massive members
                const int* sr = reinterpret_cast<const int*>(st + St::OFF_R) + (hdr.z & 3);
                const int j0 = sr[ridx] - start, j1 = sr[ridx + 1] - start;
                double Px = 0.0, Py = 0.0, Pz = 0.0, M = 0.0;
                for (int j = j0; j < j1; j++) {
                    const float4 q = sv[j];
                    const bool mass = q.w != 0.0f;
                    const float mq = mass ? rcp_fast(q.w) : 0.0f;
                    const double mqd = mass ? mass_d(q.w, mq) : 0.0;
                    const float3 vq = kicked(make_float3(q.x, q.y, q.z), a.fscale * q.w, load_force3<FFMT>(sF, fo + j));   // what member j stores
                    Px = fma(mqd, (double)vq.x, Px); Py = fma(mqd, (double)vq.y, Py); Pz = fma(mqd, (double)vq.z, Pz);
                    M += mqd;
                }
                const double keC = fma(Px, Px, fma(Py, Py, Pz * Pz)) * rcp_d(M, rcp_fast((float)M));   // M |V|^2 = |P|^2 / M
                accCOM += keC;
                ske[desc_tg(sd[fo + j0]) * TILE + tid] -= keC;
            }
        } else if (KIND == KIND_B) {
            // integrateDrudeTGNHVelocities (drudeTGNH.cu:314-364), updatePosDelta = false
            vn = kicked(make_float3(v.x, v.y, v.z), fw, F);
            if (active && massive) st_global(a.velm + start + tid, make_float4(vn.x, vn.y, vn.z, v.w));
            rel = make_float3(rel.x + fwj * Fj.x - fw * F.x, rel.y + fwj * Fj.y - fw * F.y, rel.z + fwj * Fj.z - fw * F.z);
            r = make_float3(vn.x - V.x, vn.y - V.y, vn.z - V.z);
        } else {
            // thermostat scaling (integrateDrudeTGNHChain, drudeTGNH.cu:255-300) in the unified form of the header comment
            r = make_float3(v.x - V.x, v.y - V.y, v.z - V.z);
            vn = scaled_velocity(make_float3(v.x, v.y, v.z), eT, r, eCOM, V, coef * fj, rel);
        }

        if (L::HAS_KE) {
            // computeNormalizedKineticEnergies (drudeTGNH.cu:152-188), branch-free: an ordinary particle is its own
            // "pair centre of mass" (rel = 0); the Drude particle of a pair carries the pair's two terms
            const float3 cm = make_float3(r.x + fj * rel.x, r.y + fj * rel.y, r.z + fj * rel.z);    // pair COM relative to the residue
            const bool isDrude = role == ROLE_DRUDE;
            const double md = massive ? mass_d(v.w, m) : 0.0, mjd = mass_d(vj.w, mj);
            const double Mp = md + mjd;
            const double massT = isDrude ? Mp : (role == ROLE_NORMAL ? md : 0.0);
            const double s2T = doScale ? ssq[tg] : 1.0, s2D = doScale ? ssq[G + 1] : 1.0, s2C = doScale ? ssq[G] : 1.0;
            const double keT = massT * s2T * (double)dot3(cm);
            const double keD = isDrude ? md * mjd * rcp_d(Mp, invTot) * s2D * (double)dot3(rel) : 0.0;   // reduced mass
            if (!(USE_COM && KIND != KIND_BU && active && desc_off_first(d) == 0)) keC = 0.0;        // the residue's first particle carries M |V|^2
            keC *= s2C;
            if (active && massT != 0.0) ske[tg * TILE + tid] += keT;
            accDrude += keD;
            accCOM += keC;
        }

        if (KIND == KIND_KE) {
            if (doScale && active && massive) st_global(a.velm + start + tid, make_float4(vn.x, vn.y, vn.z, v.w));
        } else if (KIND == KIND_A1) {
            // scaling + half kick; posDelta = dt * v for OpenMM's constraint kernels (drudeTGNH.cu:322-324, 360-363)
            vn = kicked(vn, fw, F);
            if (active && massive) {
                st_global(a.velm + start + tid, make_float4(vn.x, vn.y, vn.z, v.w));
                st_global(a.posDelta + start + tid, make_float4(a.dt * vn.x, a.dt * vn.y, a.dt * vn.z, 0.0f));
            }
        } else if (KIND == KIND_A2 && active) {
            // integrateDrudeTGNHPositions (drudeTGNH.cu:438-465): x += delta, v = delta / dt; then the hard wall (:474-573)
            const float4* sp = reinterpret_cast<const float4*>(st + St::OFF_P);
            const float invDt = 1.0f / a.dt;
            const float4 dl = sp[tid], x = sx[tid];
            float3 xn = make_float3(x.x + dl.x, x.y + dl.y, x.z + dl.z);
            vn = make_float3(dl.x * invDt, dl.y * invDt, dl.z * invDt);
            if (HARDWALL && role != ROLE_NORMAL) {
                const float4 dj = sp[pj], xj = sx[pj];
                float3 vjn = make_float3(dj.x * invDt, dj.y * invDt, dj.z * invDt);
                float3 delta = make_float3((x.x - xj.x) + (dl.x - dj.x), (x.y - xj.y) + (dl.y - dj.y), (x.z - xj.z) + (dl.z - dj.z));
                const float r2 = dot3(delta);
                if (r2 > rmax2) {
                    float3 xjn = make_float3(xj.x + dj.x, xj.y + dj.y, xj.z + dj.z);
                    if (role == ROLE_DRUDE) hard_wall(delta, r2, xn, xjn, vn, vjn, v.w, vj.w, a.rmax, a.hardwallScale, a.dt);
                    else hard_wall(make_float3(-delta.x, -delta.y, -delta.z), r2, xjn, xn, vjn, vn, vj.w, v.w, a.rmax, a.hardwallScale, a.dt);
                }
            }
            if (massive) {
                st_global(a.velm + start + tid, make_float4(vn.x, vn.y, vn.z, v.w));
                st_stream(a.posq + start + tid, make_float4(xn.x, xn.y, xn.z, x.w));
            }
        } else if (KIND == KIND_A && active) {
            // half kick + drift (+ hard wall) (drudeTGNH.cu:314-364, 438-465, 474-573)
            vn = kicked(vn, fw, F);
            const float4 x = sx[tid];
            float3 xn = make_float3(x.x + a.dt * vn.x, x.y + a.dt * vn.y, x.z + a.dt * vn.z);
            if (HARDWALL && role != ROLE_NORMAL) {
                // the partner's update, recomputed here so that both threads of a pair see the same wall test;
                // seen from the partner, rel changes sign and the mass fraction is this particle's
                const float3 rj = make_float3(vj.x - V.x, vj.y - V.y, vj.z - V.z);
                float3 vjn = kicked(scaled_velocity(make_float3(vj.x, vj.y, vj.z), eT, rj, eCOM, V, coef * fi, make_float3(-rel.x, -rel.y, -rel.z)), fwj, Fj);
                const float4 xj = sx[pj];
                // displacement from the exact difference of the old positions plus the relative drift: avoids the
                // cancellation of two rounded box-sized coordinates in the wall test
                float3 delta = make_float3((x.x - xj.x) + a.dt * (vn.x - vjn.x), (x.y - xj.y) + a.dt * (vn.y - vjn.y), (x.z - xj.z) + a.dt * (vn.z - vjn.z));
                const float r2 = dot3(delta);
                if (r2 > rmax2) {                     // rInv*maxDrudeDistance < 1  (drudeTGNH.cu:490)
                    float3 xjn = make_float3(xj.x + a.dt * vjn.x, xj.y + a.dt * vjn.y, xj.z + a.dt * vjn.z);
                    if (role == ROLE_DRUDE) hard_wall(delta, r2, xn, xjn, vn, vjn, v.w, vj.w, a.rmax, a.hardwallScale, a.dt);
                    else hard_wall(make_float3(-delta.x, -delta.y, -delta.z), r2, xjn, xn, vjn, vn, vj.w, v.w, a.rmax, a.hardwallScale, a.dt);
                }
            }
            if (massive) {
                st_global(a.velm + start + tid, make_float4(vn.x, vn.y, vn.z, v.w));
                st_stream(a.posq + start + tid, make_float4(xn.x, xn.y, xn.z, x.w));
            }
        }

        // hand the stage back: one arrival per warp; the producer refills it once all 16 warps are done with it
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stg]);
        if (tid == 0 && it + NS < myTiles) {
            mbar_wait(&empty[stg], phase);
            issue(it + NS, 3);
        }
    }
    pdl_launch_dependents();
    if (!L::HAS_KE) return;

    // ---- deterministic reduction: thread columns -> warp -> CTA -> (last CTA) grid ----
    ske[G * TILE + tid] += accCOM;
    ske[(G + 1) * TILE + tid] += accDrude;
    for (int g = 0; g < T; g++) {
        double x = ske[g * TILE + tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) swarp[g * 16 + warp] = x;
    }
    __syncthreads();
    if (tid < T) {
        double x = 0.0;
        for (int w = 0; w < NWARPS; w++) x += swarp[tid * 16 + w];
        a.partials[(size_t)blockIdx.x * T + tid] = x;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(a.ticket, 1u);
        smisc[0] = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!smisc[0]) return;
    __threadfence();
    // last CTA: warp g sums column g of the partials over all CTAs in a fixed order
    double* out = a.useLocalKE ? a.chain.ke2Local : a.chain.ke2;
    for (int g = warp; g < T; g += NWARPS) {
        double x = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) x += __ldcg(a.partials + (size_t)b * T + g);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) out[g] = x;
    }
    if (tid == 0) *a.ticket = 0u;
    if (KIND == KIND_KE && a.applyScale && tid < T) a.chain.pending[tid] = 1.0;   // the deferred scaling is now applied
}

// The Nose-Hoover chain update(s) between two streaming launches: one warp, lane g = thermostat g.
__global__ void __launch_bounds__(32, 1) tgnh_chain_kernel(ChainView c, int mode) {
    pdl_launch_dependents();      // the next streaming launch may start its prologue; it waits for our results
    pdl_wait();
    chain_phase(c, mode, threadIdx.x);
}

}  // namespace tgnh
