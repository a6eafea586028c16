// The three streaming kernels of the TGNH step, as one persistent, TMA-fed template.
//
//   KIND_A  (first half)   = integrateDrudeTGNHChain + integrateDrudeTGNHVelocities(updatePosDelta) +
//                            integrateDrudeTGNHPositions + applyHardWallConstraints
//                            (platforms/cuda/src/kernels/drudeTGNH.cu:249-301, 307-365, 435-466, 471-574)
//   KIND_B  (second half)  = integrateDrudeTGNHVelocities + calcCOMVelocities + normalizeVelocities +
//                            computeNormalizedKineticEnergies + sumNormalizedKineticEnergies + the chain
//                            (drudeTGNH.cu:307-365, 82-133, 138-242; CudaDrudeTGNHKernels.cpp:559-642)
//   KIND_KE (reduce/flush) = the kinetic-energy reduction alone, optionally applying a pending scaling
//                            (drudeTGNH.cu:82-242, 249-301)
//
// Data movement: every CTA is persistent and walks residue-aligned tiles of <= 512 consecutive
// particles.  One elected thread streams each tile's velm / posq / force / descriptor slices into a
// ring of shared-memory stages with cp.async.bulk (TMA) completing on an mbarrier; thread i owns
// particle i of the tile, reads its pair partner and its residue's momenta out of shared memory
// (that is where pair / residue indexing would break coalescing), and writes its own float4 results
// straight back with fully coalesced 128-bit stores.  No temporary arrays (normVelm, comVelm,
// posDelta, kineticEnergyBuffer of the reference) ever touch HBM.
#pragma once
#include "tgnh_device.cuh"

namespace tgnh {

enum { KIND_A = 0, KIND_B = 1, KIND_KE = 2 };

struct StreamArgs {
    float4* velm;
    float4* posq;
    const void* force;        // SoA [3][paddedN], float or long long
    const uint32_t* desc;     // [roundup4(N)]
    const int* tileStart;     // [numTiles + 1]
    int numTiles;
    int paddedN;
    float dt;                 // step size
    float fscale;             // 0.5*dt (f32 forces) or 0.5*dt/2^32 (i64 forces)
    float rmax;               // maxDrudeDistance
    float hardwallScale;      // sqrt(kB * T_drude)
    int applyScale;           // KIND_KE: scale velocities by scaleA and write them back
    int chainMode;            // ChainMode run by the last CTA to finish (KIND_B / KIND_KE)
    double* partials;         // [gridDim.x][T]
    unsigned int* ticket;     // last-CTA-done counter (self-resetting)
    ChainView chain;
};

template <int KIND, int FFMT>
struct StageLayout {
    static constexpr bool HAS_X = (KIND == KIND_A);
    static constexpr bool HAS_F = (KIND != KIND_KE);
    static constexpr int FBYTES = FFMT == 1 ? 8 : 4;
    static constexpr int OFF_V = 0;
    static constexpr int OFF_X = OFF_V + TILE * 16;
    static constexpr int OFF_F = OFF_X + (HAS_X ? TILE * 16 : 0);
    static constexpr int OFF_D = OFF_F + (HAS_F ? 3 * PADW * FBYTES : 0);
    static constexpr int BYTES = OFF_D + PADW * 4;
};

template <int KIND, int FFMT, bool USE_COM>
struct SmemLayout {
    using Stage = StageLayout<KIND, FFMT>;
    static constexpr bool HAS_KE = (KIND != KIND_A);
    static constexpr int NSTAGE = (KIND == KIND_A) ? 3 : 4;
    static constexpr int OFF_MOM = NSTAGE * Stage::BYTES;
    static constexpr int OFF_BAR = OFF_MOM + (USE_COM ? TILE * 16 : 0);
    static constexpr int OFF_SCALE = OFF_BAR + 64;                // float[MAX_T]
    static constexpr int OFF_MISC = OFF_SCALE + MAX_T * 4;        // int[4]
    static constexpr int OFF_WARP = OFF_MISC + 16;                // double[T][16]   (HAS_KE)
    static constexpr int OFF_KE = OFF_WARP + (HAS_KE ? MAX_T * 16 * 8 : 0);
    static int bytes(int T) { return OFF_KE + (HAS_KE ? T * TILE * 8 : 0); }
};

template <int FFMT>
__device__ __forceinline__ float load_force(const unsigned char* sF, int comp, int idx) {
    if (FFMT == 1) return (float)reinterpret_cast<const long long*>(sF)[comp * PADW + idx];
    return reinterpret_cast<const float*>(sF)[comp * PADW + idx];
}

struct PairOut { float3 v1, v2; };

// applyHardWallConstraints (drudeTGNH.cu:487-572).  1 = Drude particle, 2 = parent.  `delta` is x1 - x2.
__device__ __forceinline__ void hard_wall(float3 delta, float3& x1, float3& x2, float3& v1, float3& v2, float w1, float w2,
                                          float rmax, float hardwallScale, float dt, bool& moved) {
    const float r2 = delta.x * delta.x + delta.y * delta.y + delta.z * delta.z;
    moved = false;
    if (!(r2 > rmax * rmax)) return;      // rInv*maxDrudeDistance < 1  <=>  r > rmax
    moved = true;
    const float r = sqrtf(r2);
    const float rInv = 1.0f / r;
    const float3 bondDir = make_float3(delta.x * rInv, delta.y * rInv, delta.z * rInv);
    const float mass1 = 1.0f / w1;
    const float deltaR = r - rmax;
    float deltaT = dt;
    float dotvr1 = v1.x * bondDir.x + v1.y * bondDir.y + v1.z * bondDir.z;
    const float3 vp1 = make_float3(v1.x - bondDir.x * dotvr1, v1.y - bondDir.y * dotvr1, v1.z - bondDir.z * dotvr1);
    if (w2 == 0.0f) {
        // massless parent: only the Drude particle moves (:504-526)
        if (dotvr1 != 0.0f) deltaT = deltaR / fabsf(dotvr1);
        if (deltaT > dt) deltaT = dt;
        dotvr1 = -dotvr1 * hardwallScale / (fabsf(dotvr1) * sqrtf(mass1));
        const float dr = -deltaR + deltaT * dotvr1;
        x1.x += bondDir.x * dr; x1.y += bondDir.y * dr; x1.z += bondDir.z * dr;
        v1 = make_float3(vp1.x + bondDir.x * dotvr1, vp1.y + bondDir.y * dotvr1, vp1.z + bondDir.z * dotvr1);
    } else {
        const float mass2 = 1.0f / w2;
        const float invTotalMass = 1.0f / (mass1 + mass2);
        float dotvr2 = v2.x * bondDir.x + v2.y * bondDir.y + v2.z * bondDir.z;
        const float3 vp2 = make_float3(v2.x - bondDir.x * dotvr2, v2.y - bondDir.y * dotvr2, v2.z - bondDir.z * dotvr2);
        const float vbCMass = (mass1 * dotvr1 + mass2 * dotvr2) * invTotalMass;
        dotvr1 -= vbCMass;
        dotvr2 -= vbCMass;
        if (dotvr1 != dotvr2) deltaT = deltaR / fabsf(dotvr1 - dotvr2);
        if (deltaT > dt) deltaT = dt;
        const float vBond = hardwallScale / sqrtf(mass1);
        dotvr1 = -dotvr1 * vBond * mass2 * invTotalMass / fabsf(dotvr1);
        dotvr2 = -dotvr2 * vBond * mass1 * invTotalMass / fabsf(dotvr2);
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

__device__ __forceinline__ double sq3(float x, float y, float z) {
    const double dx = x, dy = y, dz = z;
    return fma(dx, dx, fma(dy, dy, dz * dz));
}

template <int KIND, int FFMT, bool USE_COM, bool HARDWALL>
__global__ void __launch_bounds__(TILE, 2) tgnh_stream_kernel(const __grid_constant__ StreamArgs a) {
    using L = SmemLayout<KIND, FFMT, USE_COM>;
    using St = typename L::Stage;
    constexpr int NS = L::NSTAGE;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    float* sscale = reinterpret_cast<float*>(smem + L::OFF_SCALE);
    int* smisc = reinterpret_cast<int*>(smem + L::OFF_MISC);
    float4* smom = reinterpret_cast<float4*>(smem + L::OFF_MOM);
    double* ske = reinterpret_cast<double*>(smem + L::OFF_KE);
    double* swarp = reinterpret_cast<double*>(smem + L::OFF_WARP);

    const int tid = threadIdx.x;
    const int T = a.chain.T, G = a.chain.G;
    const int myTiles = blockIdx.x < a.numTiles ? (a.numTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (tid == 0) {
        for (int s = 0; s < NS; s++) mbar_init(&bars[s], 1);
        fence_mbar_init();
    }
    if (L::HAS_KE)
        for (int g = 0; g < T; g++) ske[g * TILE + tid] = 0.0;
    if (tid < T) sscale[tid] = (KIND == KIND_B) ? 1.0f : (float)a.chain.scaleA[tid];
    __syncthreads();

    const uint64_t policy = policy_evict_first();
    auto issue = [&](int it) {
        const int tile = blockIdx.x + it * gridDim.x;
        const int start = a.tileStart[tile], end = a.tileStart[tile + 1];
        const int n = end - start, a0 = start & ~3, na = ((end + 3) & ~3) - a0;
        unsigned char* st = smem + (it % NS) * St::BYTES;
        uint64_t* bar = &bars[it % NS];
        uint32_t bytes = n * 16 + na * 4;
        if (St::HAS_X) bytes += n * 16;
        if (St::HAS_F) bytes += 3 * na * St::FBYTES;
        mbar_arrive_expect_tx(bar, bytes);
        bulk_g2s(st + St::OFF_V, a.velm + start, n * 16, bar, policy);
        if (St::HAS_X) bulk_g2s(st + St::OFF_X, a.posq + start, n * 16, bar, policy);
        if (St::HAS_F) {
            const unsigned char* f = static_cast<const unsigned char*>(a.force);
            for (int c = 0; c < 3; c++)
                bulk_g2s(st + St::OFF_F + c * PADW * St::FBYTES, f + ((size_t)c * a.paddedN + a0) * St::FBYTES, na * St::FBYTES, bar,
                         policy);
        }
        bulk_g2s(st + St::OFF_D, a.desc + a0, na * 4, bar, policy);
    };
    if (tid == 0)
        for (int it = 0; it < NS && it < myTiles; it++) issue(it);

    const float sCOM = (KIND == KIND_B) ? 1.0f : sscale[G];
    const float sDrude = (KIND == KIND_B) ? 1.0f : sscale[G + 1];
    const bool doScale = (KIND == KIND_A) || (KIND == KIND_KE && a.applyScale);

    for (int it = 0; it < myTiles; it++) {
        const int stg = it % NS;
        const uint32_t phase = (it / NS) & 1;
        const int tile = blockIdx.x + it * gridDim.x;
        const int start = a.tileStart[tile];
        const int n = a.tileStart[tile + 1] - start;
        const int fo = start & 3;                     // offset of the tile inside its 4-aligned window
        unsigned char* st = smem + stg * St::BYTES;
        const float4* sv = reinterpret_cast<const float4*>(st + St::OFF_V);
        const float4* sx = reinterpret_cast<const float4*>(st + St::OFF_X);
        const unsigned char* sF = st + St::OFF_F;
        const uint32_t* sd = reinterpret_cast<const uint32_t*>(st + St::OFF_D);

        mbar_wait(&bars[stg], phase);

        const bool active = tid < n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t d = 0;
        if (active) { v = sv[tid]; d = sd[fo + tid]; }
        const uint32_t role = desc_role(d);
        const int tg = desc_tg(d);
        const bool massive = v.w != 0.0f;
        const float m = massive ? 1.0f / v.w : 0.0f;
        const int pj = tid + desc_partner(d);

        // ---- per-thread velocity update that does not need the residue's COM velocity (KIND_B: the kick) ----
        float3 nv = make_float3(v.x, v.y, v.z);       // this particle's new velocity
        float3 cmv = make_float3(0.f, 0.f, 0.f), relv = cmv;   // pair: centre-of-mass / relative velocity (KIND_B: after the kick)
        float m1 = 0.f, m2 = 0.f, f1 = 0.f, f2 = 0.f, w1 = 0.f, w2 = 0.f;
        float4 vo = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool isPair = active && role != ROLE_NORMAL;
        const bool iAmDrude = role == ROLE_DRUDE;
        if (isPair) {
            vo = sv[pj];
            w1 = iAmDrude ? v.w : vo.w;               // 1 = Drude particle, 2 = parent (pairParticles.x / .y)
            w2 = iAmDrude ? vo.w : v.w;
            m1 = 1.0f / w1; m2 = 1.0f / w2;
            const float invTot = 1.0f / (m1 + m2);
            f1 = invTot * m1; f2 = invTot * m2;
        }
        if (KIND == KIND_B) {
            // integrateDrudeTGNHVelocities (drudeTGNH.cu:314-364), updatePosDelta = false
            if (active) {
                const float fx = load_force<FFMT>(sF, 0, fo + tid), fy = load_force<FFMT>(sF, 1, fo + tid), fz = load_force<FFMT>(sF, 2, fo + tid);
                if (!isPair) {
                    if (massive) {
                        nv.x = v.x + a.fscale * v.w * fx; nv.y = v.y + a.fscale * v.w * fy; nv.z = v.z + a.fscale * v.w * fz;
                    }
                } else {
                    const float ox = load_force<FFMT>(sF, 0, fo + pj), oy = load_force<FFMT>(sF, 1, fo + pj), oz = load_force<FFMT>(sF, 2, fo + pj);
                    const float3 v1 = iAmDrude ? make_float3(v.x, v.y, v.z) : make_float3(vo.x, vo.y, vo.z);
                    const float3 v2 = iAmDrude ? make_float3(vo.x, vo.y, vo.z) : make_float3(v.x, v.y, v.z);
                    const float3 F1 = iAmDrude ? make_float3(fx, fy, fz) : make_float3(ox, oy, oz);
                    const float3 F2 = iAmDrude ? make_float3(ox, oy, oz) : make_float3(fx, fy, fz);
                    const float invTot = 1.0f / (m1 + m2);
                    const float invRed = (m1 + m2) * w1 * w2;
                    cmv = make_float3(v1.x * f1 + v2.x * f2, v1.y * f1 + v2.y * f2, v1.z * f1 + v2.z * f2);
                    relv = make_float3(v2.x - v1.x, v2.y - v1.y, v2.z - v1.z);
                    cmv.x += a.fscale * invTot * (F1.x + F2.x); cmv.y += a.fscale * invTot * (F1.y + F2.y); cmv.z += a.fscale * invTot * (F1.z + F2.z);
                    relv.x += a.fscale * invRed * (F2.x * f1 - F1.x * f2);
                    relv.y += a.fscale * invRed * (F2.y * f1 - F1.y * f2);
                    relv.z += a.fscale * invRed * (F2.z * f1 - F1.z * f2);
                    nv = iAmDrude ? make_float3(cmv.x - relv.x * f2, cmv.y - relv.y * f2, cmv.z - relv.z * f2)
                                  : make_float3(cmv.x + relv.x * f1, cmv.y + relv.y * f1, cmv.z + relv.z * f1);
                }
                if (massive) st_stream(a.velm + start + tid, make_float4(nv.x, nv.y, nv.z, v.w));
            }
        }

        // ---- residue centre-of-mass velocity (calcCOMVelocities, drudeTGNH.cu:86-105) ----
        float3 V = make_float3(0.f, 0.f, 0.f);
        float Mres = 0.f;
        if (USE_COM) {
            smom[tid] = make_float4(nv.x * m, nv.y * m, nv.z * m, m);   // inactive lanes and massless particles: zeros
            __syncthreads();
            if (active) {
                const int j0 = tid - desc_off_first(d), j1 = tid + desc_off_last(d);
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int j = j0; j <= j1; j++) {
                    const float4 q = smom[j];
                    acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
                }
                Mres = acc.w;
                const float inv = 1.0f / acc.w;
                V = make_float3(acc.x * inv, acc.y * inv, acc.z * inv);
            }
        }

        if (KIND == KIND_A || KIND == KIND_KE) {
            // ---- thermostat scaling (integrateDrudeTGNHChain, drudeTGNH.cu:255-300) ----
            if (isPair) {
                const float3 v1 = iAmDrude ? make_float3(v.x, v.y, v.z) : make_float3(vo.x, vo.y, vo.z);
                const float3 v2 = iAmDrude ? make_float3(vo.x, vo.y, vo.z) : make_float3(v.x, v.y, v.z);
                // cm' = centre of mass of the pair relative to the residue, rel = v2 - v1
                cmv = make_float3((v1.x - V.x) * f1 + (v2.x - V.x) * f2, (v1.y - V.y) * f1 + (v2.y - V.y) * f2, (v1.z - V.z) * f1 + (v2.z - V.z) * f2);
                relv = make_float3(v2.x - v1.x, v2.y - v1.y, v2.z - v1.z);
            }
            float3 rv = make_float3(v.x - V.x, v.y - V.y, v.z - V.z);    // ordinary particle relative to the residue
            float3 Vs = V;
            if (doScale) {
                const float sT = sscale[tg];
                rv.x *= sT; rv.y *= sT; rv.z *= sT;
                cmv.x *= sT; cmv.y *= sT; cmv.z *= sT;
                relv.x *= sDrude; relv.y *= sDrude; relv.z *= sDrude;
                Vs.x *= sCOM; Vs.y *= sCOM; Vs.z *= sCOM;
            }
            if (L::HAS_KE && active) {
                // computeNormalizedKineticEnergies (drudeTGNH.cu:152-188) on the (scaled) velocities
                if (!isPair) {
                    if (massive) ske[tg * TILE + tid] += (double)m * sq3(rv.x, rv.y, rv.z);
                } else if (iAmDrude) {
                    ske[tg * TILE + tid] += (double)(m1 + m2) * sq3(cmv.x, cmv.y, cmv.z);
                    ske[(G + 1) * TILE + tid] += (double)(m1 * m2 / (m1 + m2)) * sq3(relv.x, relv.y, relv.z);
                }
                if (USE_COM && desc_off_first(d) == 0) ske[G * TILE + tid] += (double)Mres * sq3(Vs.x, Vs.y, Vs.z);
            }
            if (KIND == KIND_KE) {
                if (doScale && active && massive) {
                    if (!isPair) nv = make_float3(rv.x + Vs.x, rv.y + Vs.y, rv.z + Vs.z);
                    else nv = iAmDrude ? make_float3(cmv.x - relv.x * f2 + Vs.x, cmv.y - relv.y * f2 + Vs.y, cmv.z - relv.z * f2 + Vs.z)
                                       : make_float3(cmv.x + relv.x * f1 + Vs.x, cmv.y + relv.y * f1 + Vs.y, cmv.z + relv.z * f1 + Vs.z);
                    st_stream(a.velm + start + tid, make_float4(nv.x, nv.y, nv.z, v.w));
                }
            } else if (active) {
                // ---- half kick + drift (+ hard wall) (drudeTGNH.cu:314-364, 438-465, 474-573) ----
                const float4 x = sx[tid];
                const float fx = load_force<FFMT>(sF, 0, fo + tid), fy = load_force<FFMT>(sF, 1, fo + tid), fz = load_force<FFMT>(sF, 2, fo + tid);
                if (!isPair) {
                    if (massive) {
                        nv.x = rv.x + Vs.x + a.fscale * v.w * fx;
                        nv.y = rv.y + Vs.y + a.fscale * v.w * fy;
                        nv.z = rv.z + Vs.z + a.fscale * v.w * fz;
                        st_stream(a.velm + start + tid, make_float4(nv.x, nv.y, nv.z, v.w));
                        st_stream(a.posq + start + tid, make_float4(x.x + a.dt * nv.x, x.y + a.dt * nv.y, x.z + a.dt * nv.z, x.w));
                    }
                } else {
                    const float ox = load_force<FFMT>(sF, 0, fo + pj), oy = load_force<FFMT>(sF, 1, fo + pj), oz = load_force<FFMT>(sF, 2, fo + pj);
                    const float4 xo = sx[pj];
                    const float3 F1 = iAmDrude ? make_float3(fx, fy, fz) : make_float3(ox, oy, oz);
                    const float3 F2 = iAmDrude ? make_float3(ox, oy, oz) : make_float3(fx, fy, fz);
                    const float invTot = 1.0f / (m1 + m2);
                    const float invRed = (m1 + m2) * w1 * w2;
                    // pair centre of mass in the lab frame: scaled relative part + scaled residue COM
                    float3 cm = make_float3(cmv.x + Vs.x, cmv.y + Vs.y, cmv.z + Vs.z);
                    cm.x += a.fscale * invTot * (F1.x + F2.x); cm.y += a.fscale * invTot * (F1.y + F2.y); cm.z += a.fscale * invTot * (F1.z + F2.z);
                    relv.x += a.fscale * invRed * (F2.x * f1 - F1.x * f2);
                    relv.y += a.fscale * invRed * (F2.y * f1 - F1.y * f2);
                    relv.z += a.fscale * invRed * (F2.z * f1 - F1.z * f2);
                    float3 nv1 = make_float3(cm.x - relv.x * f2, cm.y - relv.y * f2, cm.z - relv.z * f2);
                    float3 nv2 = make_float3(cm.x + relv.x * f1, cm.y + relv.y * f1, cm.z + relv.z * f1);
                    const float4 x1o = iAmDrude ? x : xo, x2o = iAmDrude ? xo : x;
                    float3 x1 = make_float3(x1o.x + a.dt * nv1.x, x1o.y + a.dt * nv1.y, x1o.z + a.dt * nv1.z);
                    float3 x2 = make_float3(x2o.x + a.dt * nv2.x, x2o.y + a.dt * nv2.y, x2o.z + a.dt * nv2.z);
                    if (HARDWALL) {
                        // Drude displacement from the exact difference of the old positions plus the relative drift:
                        // avoids the cancellation of two rounded ~box-sized coordinates in the wall test
                        const float3 delta = make_float3((x1o.x - x2o.x) - a.dt * relv.x, (x1o.y - x2o.y) - a.dt * relv.y, (x1o.z - x2o.z) - a.dt * relv.z);
                        bool moved;
                        hard_wall(delta, x1, x2, nv1, nv2, w1, w2, a.rmax, a.hardwallScale, a.dt, moved);
                    }
                    const float3 mv = iAmDrude ? nv1 : nv2;
                    const float3 mx = iAmDrude ? x1 : x2;
                    if (massive) {
                        st_stream(a.velm + start + tid, make_float4(mv.x, mv.y, mv.z, v.w));
                        st_stream(a.posq + start + tid, make_float4(mx.x, mx.y, mx.z, x.w));
                    }
                }
            }
        } else {
            // KIND_B: kinetic energies of the kicked velocities relative to the residue COM
            if (active) {
                if (!isPair) {
                    if (massive) ske[tg * TILE + tid] += (double)m * sq3(nv.x - V.x, nv.y - V.y, nv.z - V.z);
                } else if (iAmDrude) {
                    ske[tg * TILE + tid] += (double)(m1 + m2) * sq3(cmv.x - V.x, cmv.y - V.y, cmv.z - V.z);
                    ske[(G + 1) * TILE + tid] += (double)(m1 * m2 / (m1 + m2)) * sq3(relv.x, relv.y, relv.z);
                }
                if (USE_COM && desc_off_first(d) == 0) ske[G * TILE + tid] += (double)Mres * sq3(V.x, V.y, V.z);
            }
        }

        __syncthreads();                               // every read of this stage (and of smom) is done
        if (tid == 0 && it + NS < myTiles) issue(it + NS);
    }

    if (!L::HAS_KE) return;

    // ---- deterministic reduction: thread columns -> warp -> CTA -> (last CTA) grid -> chain ----
    const int lane = tid & 31, warp = tid >> 5;
    for (int g = 0; g < T; g++) {
        double x = ske[g * TILE + tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) swarp[g * 16 + warp] = x;
    }
    __syncthreads();
    if (tid < T) {
        double x = 0.0;
        for (int w = 0; w < TILE / 32; w++) x += swarp[tid * 16 + w];
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
    double* out = (a.chainMode == CHAIN_NONE && a.chain.ke2Local) ? a.chain.ke2Local : a.chain.ke2;
    for (int g = warp; g < T; g += TILE / 32) {
        double x = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) x += __ldcg(a.partials + (size_t)b * T + g);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) out[g] = x;
    }
    if (tid == 0) *a.ticket = 0u;
    if (KIND == KIND_KE && a.applyScale && tid < T) a.chain.pending[tid] = 1.0;   // the deferred scaling is now applied
    __syncthreads();
    if (a.chainMode != CHAIN_NONE && warp == 0) chain_phase(a.chain, a.chainMode, lane);
}

// stand-alone chain update (sharded runs after the all-reduce; first step of a tgnh_step batch)
__global__ void tgnh_chain_kernel(ChainView c, int mode) { chain_phase(c, mode, threadIdx.x); }

}  // namespace tgnh
