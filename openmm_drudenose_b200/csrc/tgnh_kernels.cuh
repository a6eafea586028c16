// The streaming kernels of the TGNH step (nine kinds of one pass over the particles), as one persistent, TMA-fed template.
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
//   KIND_K  (kick)         = integrateDrudeTGNHVelocities alone (drudeTGNH.cu:307-365): the second half kick of constrained
//                            systems, whose energies are reduced only after OpenMM's velocity constraints
//   KIND_KU                = KIND_KE for residue-uniform groups, in the lab-frame form of KIND_BU (no kick, no store)
//   KIND_S  (scale)        = integrateDrudeTGNHChain alone (drudeTGNH.cu:249-301): applies the second thermostat half-step's
//                            factors to velm right away, for callers whose other kernels (CMMotionRemover, barostat,
//                            reporters) read velocities between steps; residue-uniform groups only, where the scaled
//                            kinetic energies are s_g^2 KE_g and need no second reduction
//
// (The two halves and the plain reduction of single-precision systems whose residues fit a warp run through the second
// generation of these kernels, tgnh_v2.cuh; this template serves every other case.)
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
//
// Precision (template parameter PREC): 0 = OpenMM's single-precision layout (float4 velm, float4 posq), fp32 arithmetic
// with fp64 energy sums; 1 = OpenMM's mixed-precision layout (double4 velm, float4 posq + float4 posqCorrection, double4
// posDelta; `mixed` = double in drudeTGNH.cu), all arithmetic in double; 2 = OpenMM's double-precision layout (as 1, but
// posq is double4 and there is no posqCorrection) — instantiated only for the kernels that touch positions.
#pragma once
#include "tgnh_device.cuh"

namespace tgnh {

enum { KIND_A = 0, KIND_B = 1, KIND_KE = 2, KIND_BU = 3, KIND_A1 = 4, KIND_A2 = 5, KIND_S = 6, KIND_K = 7, KIND_KU = 8 };
constexpr int NWARPS = TILE / 32;
constexpr int TLIST_CAP = 256;     // tile bounds cached in shared memory per CTA (later tiles are looked up in global memory)

template <int PREC> struct Prec;
template <> struct Prec<0> { typedef float real; typedef float4 real4; };
template <> struct Prec<1> { typedef double real; typedef double4 real4; };
template <> struct Prec<2> { typedef double real; typedef double4 real4; };   // differs from 1 only where positions are touched

struct StreamArgs {
    void* velm;               // real4[paddedN]
    void* posq;               // float4[paddedN] (single, mixed) or double4[paddedN] (double)
    float4* posqCorrection;   // float4[paddedN], mixed precision only (cu.getPosqCorrection())
    void* posDelta;           // real4[paddedN] (dt*v, 0): integration.getPosDelta() (KIND_A1 writes, KIND_A2 reads)
    const void* force;        // SoA [3][paddedN], float or long long
    const uint32_t* desc;     // [roundup4(N)]
    const int* tileStart;     // [numTiles + 1]
    const int* resStart;      // [R + 1] first particle of every residue in particle order, then N   (KIND_BU)
    const int* tileFirstRes;  // [numTiles + 1] index into resStart of each tile's first residue      (KIND_BU)
    const int* bigFirst;      // [numBig] first particle of every big residue (> MAX_RES particles), ascending
    const double4* bigCom;    // [numBig] {V_x, V_y, V_z, M} of the big residues, written by tgnh_bigcom_kernel before this launch
    int numBig;
    int numTiles;
    int paddedN;
    double dt;                // step size
    double fscale;            // 0.5*dt (f32 forces) or 0.5*dt/2^32 (i64 forces)
    double rmax;              // maxDrudeDistance
    double hardwallScale;     // sqrt(kB * T_drude)
    int applyScale;           // KIND_KE: scale velocities by scaleA and write them back
    int useLocalKE;           // sharded: reduce into chain.ke2Local (all-reduced into ke2 afterwards)
    int reverse;              // walk the tiles from the last to the first (see "L2 hand-over" below)
    int prologuePrefetch;     // tiles per CTA whose read-only inputs are prefetched into L2 before griddepcontrol.wait
    int fusedChainMode;       // tgnh_stream_chain_kernel: the chain update the last CTA runs (ChainMode)
    int earlyLoads;           // the launches that precede this one in the stream are this library's own and write neither
                              // posq, forces nor posqCorrection: those tiles may be requested before griddepcontrol.wait
    // warp-chunk kernels (tgnh_v2.cuh)
    int uniformGroups;           // every residue lies in one temperature group
    int lazyKick;                // second half: the kicked velocities are reduced but NOT stored; first half: velm still lacks the
                                 // previous step's second half kick, which is applied (same forces, same fp32 operation, bit-identical
                                 // velocities) before anything else.  Only between the steps of one tgnh_step(n) call.
    const unsigned char* spec;   // [roundup16(N) + 32] species-table row of every particle
    const int* chunkStart;       // [15 * numTiles + 1] first particle of every chunk (tail padded with N)
    const float4* specTable;     // [tableRows * 3] species table (q0, q1, pad); the last row is "no particle"
    int tableRows;
    int maxRes;                  // particles in the longest residue (<= 32)
    int butterfly;               // > 0: every residue has this many particles (a power of two) and every chunk is full or ends the system
    int resPerLane;              // > 0: reducing launch in the residue-per-lane form (tgnh_v2.cuh): every residue has this many particles (2..8)
    int tileBegin;               // first tile of this launch (launches over a sub-range of the tiles: the chunked host-buffer path)
    int accumulate;              // add this launch's energy sums to what the previous launch over another sub-range left
    double* partials;         // [gridDim.x][T]
    unsigned int* ticket;     // last-CTA-done counter (self-resetting)
    ChainView chain;
    PeerView peers;           // sharded runs with peer-mapped inboxes: where the last CTA publishes this rank's sums
};

template <int KIND, int FFMT, int PREC>
struct StageLayout {
    static constexpr bool HAS_X = (KIND == KIND_A || KIND == KIND_A2);
    static constexpr bool HAS_F = (KIND != KIND_KE && KIND != KIND_A2 && KIND != KIND_S && KIND != KIND_KU);
    static constexpr bool HAS_P = (KIND == KIND_A2);             // posDelta tile
    static constexpr bool HAS_R = (KIND == KIND_BU || KIND == KIND_KU);
    static constexpr int FBYTES = FFMT == 1 ? 8 : 4;
    static constexpr int VB = PREC ? 32 : 16;                    // bytes of one velm / posDelta element
    static constexpr int OFF_V = 0;
    static constexpr int OFF_X = OFF_V + TILE * VB;
    static constexpr int XB = PREC == 2 ? 32 : 16;               // bytes of one posq element
    static constexpr int OFF_XC = OFF_X + (HAS_X ? TILE * XB : 0);             // posqCorrection tile (mixed)
    static constexpr int OFF_F = OFF_XC + ((HAS_X && PREC == 1) ? TILE * 16 : 0);
    static constexpr int OFF_D = OFF_F + (HAS_F ? 3 * PADW * FBYTES : 0);
    static constexpr int OFF_R = OFF_D + PADW * 4;              // residue starts of the tile (KIND_BU)
    static constexpr int OFF_P = OFF_R + (HAS_R ? PADW * 4 : 0);
    static constexpr int OFF_HDR = OFF_P + (HAS_P ? TILE * VB : 0);   // int4 {first particle, count, first residue, residues}
    static constexpr int BYTES = OFF_HDR + 16;
};

template <int KIND, int FFMT, bool USE_COM, int PREC>
struct SmemLayout {
    using Stage = StageLayout<KIND, FFMT, PREC>;
    static constexpr bool HAS_KE = (KIND == KIND_B || KIND == KIND_BU || KIND == KIND_KE || KIND == KIND_KU);
    static constexpr int NSTAGE = (KIND == KIND_A || KIND == KIND_A2) ? 3 : 4;
    static constexpr int OFF_BAR = NSTAGE * Stage::BYTES;         // full[NS], empty[NS] mbarriers
    static constexpr int OFF_TLIST = OFF_BAR + 128;               // int4[TLIST_CAP] bounds of this CTA's first tiles
    static constexpr int OFF_SCALE = OFF_TLIST + TLIST_CAP * 16;  // double[MAX_T] s^2, double[MAX_T] s - 1
    static constexpr int OFF_MISC = OFF_SCALE + MAX_T * 16;       // int[4]
    static constexpr int OFF_WARP = OFF_MISC + 16;                // double[T][16]   (HAS_KE)
    static constexpr int OFF_KE = OFF_WARP + (HAS_KE ? MAX_T * 16 * 8 : 0);
    static int bytes(int T) { return OFF_KE + (HAS_KE ? T * TILE * 8 : 0); }
};

// ---- small vector helpers, generic in the arithmetic type ------------------------------------------------------
template <typename T> struct V3 { T x, y, z; };
template <typename T> __device__ __forceinline__ V3<T> v3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> __device__ __forceinline__ V3<T> operator-(V3<T> a) { return v3(-a.x, -a.y, -a.z); }
template <typename T> __device__ __forceinline__ V3<T> operator*(T s, V3<T> a) { return v3(s * a.x, s * a.y, s * a.z); }
template <typename T> __device__ __forceinline__ T dot3(V3<T> a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
template <typename T> __device__ __forceinline__ T dot3(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// a + s*b with one rounding per component
template <typename T> __device__ __forceinline__ V3<T> axpy(T s, V3<T> b, V3<T> a) { return v3(fma(s, b.x, a.x), fma(s, b.y, a.y), fma(s, b.z, a.z)); }
__device__ __forceinline__ V3<float> xyz(float4 q) { return v3(q.x, q.y, q.z); }
__device__ __forceinline__ V3<double> xyz(double4 q) { return v3(q.x, q.y, q.z); }
__device__ __forceinline__ float4 pack4(V3<float> a, float w) { return make_float4(a.x, a.y, a.z, w); }
__device__ __forceinline__ double4 pack4(V3<double> a, double w) { return make_double4(a.x, a.y, a.z, w); }
template <typename T> __device__ __forceinline__ V3<double> to_double(V3<T> a) { return v3((double)a.x, (double)a.y, (double)a.z); }

__device__ __forceinline__ void st_stream(double4* p, double4 v) {
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" ::"l"(reinterpret_cast<double2*>(p) + 1), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ void st_global(double4* p, double4 v) {
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    asm volatile("st.global.v2.f64 [%0], {%1, %2};" ::"l"(reinterpret_cast<double2*>(p) + 1), "d"(v.z), "d"(v.w) : "memory");
}

template <int FFMT, typename T>
__device__ __forceinline__ V3<T> load_force3(const unsigned char* sF, int idx) {
    if (FFMT == 1) {
        const long long* f = reinterpret_cast<const long long*>(sF);
        return v3((T)f[idx], (T)f[PADW + idx], (T)f[2 * PADW + idx]);
    }
    const float* f = reinterpret_cast<const float*>(sF);
    return v3((T)f[idx], (T)f[PADW + idx], (T)f[2 * PADW + idx]);
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
// 1/x in double from a float seed: two Newton steps (1e-7 -> 1e-14 -> rounding)
__device__ __forceinline__ double rcp_d(double x, float seed) {
    double r = (double)seed;
    r = r * fma(-x, r, 2.0);
    return r * fma(-x, r, 2.0);
}
__device__ __forceinline__ double rcp_fast(double x) { return rcp_d(x, rcp_fast((float)x)); }
// Masses that weight the kinetic-energy sums are formed in double.  In the fp32 layout a species' rounded mass (all
// oxygens share one w) would otherwise shift its thermostat's energy by up to 6e-8 — systematically, so it does not
// average out over particles, and the Nose-Hoover chain integrates it.
__device__ __forceinline__ double mass_d(float w, float r) { return (double)r + (double)rcp_lo(w, r); }
__device__ __forceinline__ double mass_d(double w, double r) { return r; }
__device__ __forceinline__ float rsqrt_fast(float x) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double rsqrt_fast(double x) { return rsqrt(x); }
__device__ __forceinline__ float fast_div(float a, float b) { return __fdividef(a, b); }
__device__ __forceinline__ double fast_div(double a, double b) { return a / b; }

// applyHardWallConstraints (drudeTGNH.cu:487-572) for one pair that is beyond the wall.
// 1 = Drude particle, 2 = parent.  `delta` is x1 - x2, r2 its squared length.
template <typename T>
__device__ __forceinline__ void hard_wall(V3<T> delta, T r2, V3<T>& x1, V3<T>& x2, V3<T>& v1, V3<T>& v2, T w1, T w2, T rmax, T hardwallScale, T dt) {
    const T rInv = rsqrt_fast(r2);
    const T r = r2 * rInv;
    const V3<T> bondDir = rInv * delta;
    const T mass1 = rcp_fast(w1);
    const T deltaR = r - rmax;
    T deltaT = dt;
    T dotvr1 = dot3(v1, bondDir);
    const V3<T> vp1 = axpy(-dotvr1, bondDir, v1);
    const T vBond = hardwallScale * sqrt(w1);                         // hardwallscaleDrude / SQRT(mass1)
    if (w2 == T(0)) {
        // massless parent: only the Drude particle moves (:504-526)
        if (dotvr1 != T(0)) deltaT = fast_div(deltaR, fabs(dotvr1));
        if (deltaT > dt) deltaT = dt;
        dotvr1 = -copysign(vBond, dotvr1);                            // -dotvr1*scale/(|dotvr1|*sqrt(m1))
        const T dr = -deltaR + deltaT * dotvr1;
        x1 = axpy(dr, bondDir, x1);
        v1 = axpy(dotvr1, bondDir, vp1);
    } else {
        const T mass2 = rcp_fast(w2);
        const T invTotalMass = rcp_fast(mass1 + mass2);
        T dotvr2 = dot3(v2, bondDir);
        const V3<T> vp2 = axpy(-dotvr2, bondDir, v2);
        const T vbCMass = (mass1 * dotvr1 + mass2 * dotvr2) * invTotalMass;
        dotvr1 -= vbCMass;
        dotvr2 -= vbCMass;
        if (dotvr1 != dotvr2) deltaT = fast_div(deltaR, fabs(dotvr1 - dotvr2));
        if (deltaT > dt) deltaT = dt;
        dotvr1 = -copysign(vBond * mass2 * invTotalMass, dotvr1);    // :542  (-dotvr1*vBond*m2/M/|dotvr1|)
        dotvr2 = -copysign(vBond * mass1 * invTotalMass, dotvr2);    // :543
        const T dr1 = -deltaR * mass2 * invTotalMass + deltaT * dotvr1;
        const T dr2 = deltaR * mass1 * invTotalMass + deltaT * dotvr2;
        dotvr1 += vbCMass;
        dotvr2 += vbCMass;
        x1 = axpy(dr1, bondDir, x1);
        x2 = axpy(dr2, bondDir, x2);
        v1 = axpy(dotvr1, bondDir, vp1);
        v2 = axpy(dotvr2, bondDir, vp2);
    }
}

// v + eT*r + eCOM*V + c*rel (e = s - 1) with a fixed evaluation order, so that the two threads of a Drude pair compute
// bit-identical values for each other's particle (both must take the same side of the hard-wall test)
template <typename T>
__device__ __forceinline__ V3<T> scaled_velocity(V3<T> v, T eT, V3<T> r, T eCOM, V3<T> V, T c, V3<T> rel) {
    return v3(v.x + fma(c, rel.x, fma(eT, r.x, eCOM * V.x)), v.y + fma(c, rel.y, fma(eT, r.y, eCOM * V.y)),
              v.z + fma(c, rel.z, fma(eT, r.z, eCOM * V.z)));
}
template <typename T> __device__ __forceinline__ V3<T> kicked(V3<T> v, T fw, V3<T> F) { return axpy(fw, F, v); }

// The same with the factors eT and eCOM as float PAIRS (hi + lo = the double s - 1).  A factor rounded to fp32 is off by up to
// 6e-8 |s - 1|: when a thermostat works hard (|s - 1| ~ 1e-2: start-up transients, the Drude thermostat) that is a scale error of
// 5e-10 on every particle of the group, with one sign, step after step — and a systematic scale error of 1e-9 per step moves the
// group's chain velocities by 1e-6 relative (measured: scripts/dev_bias.py, DESIGN.md "Parity").
struct F2 { float hi, lo; };
__device__ __forceinline__ F2 split2(double x) { F2 r; r.hi = (float)x; r.lo = (float)(x - (double)r.hi); return r; }
__device__ __forceinline__ V3<float> scaled_velocity2(V3<float> v, F2 eT, V3<float> r, F2 eCOM, V3<float> V, float c, V3<float> rel) {
    const V3<float> t = v3(fmaf(eT.lo, r.x, eCOM.lo * V.x), fmaf(eT.lo, r.y, eCOM.lo * V.y), fmaf(eT.lo, r.z, eCOM.lo * V.z));
    return v3(v.x + fmaf(c, rel.x, fmaf(eT.hi, r.x, fmaf(eCOM.hi, V.x, t.x))), v.y + fmaf(c, rel.y, fmaf(eT.hi, r.y, fmaf(eCOM.hi, V.y, t.y))),
              v.z + fmaf(c, rel.z, fmaf(eT.hi, r.z, fmaf(eCOM.hi, V.z, t.z))));
}

// positions: single = posq; mixed = posq + posqCorrection in double (drudeTGNH.cu:441-448, 476-486)
template <int PREC> struct PosTile;
template <> struct PosTile<0> {
    typedef float Q;                                   // type of the charge carried in posq.w
    const float4* sx;
    __device__ __forceinline__ V3<float> load(int i, float& q) const { const float4 p = sx[i]; q = p.w; return v3(p.x, p.y, p.z); }
    __device__ __forceinline__ static void store(const StreamArgs& a, int gi, V3<float> x, float q) {
        st_stream(static_cast<float4*>(a.posq) + gi, make_float4(x.x, x.y, x.z, q));
    }
};
template <> struct PosTile<2> {
    typedef double Q;
    const double4* sx;
    __device__ __forceinline__ V3<double> load(int i, double& q) const { const double4 p = sx[i]; q = p.w; return v3(p.x, p.y, p.z); }
    __device__ __forceinline__ static void store(const StreamArgs& a, int gi, V3<double> x, double q) {
        st_stream(static_cast<double4*>(a.posq) + gi, make_double4(x.x, x.y, x.z, q));
    }
};
template <> struct PosTile<1> {
    typedef float Q;
    const float4* sx;
    const float4* sc;
    __device__ __forceinline__ V3<double> load(int i, float& q) const {
        const float4 p = sx[i], c = sc[i];
        q = p.w;
        return v3(p.x + (double)c.x, p.y + (double)c.y, p.z + (double)c.z);
    }
    __device__ __forceinline__ static void store(const StreamArgs& a, int gi, V3<double> x, float q) {
        const float hx = (float)x.x, hy = (float)x.y, hz = (float)x.z;                       // :457-458
        st_stream(static_cast<float4*>(a.posq) + gi, make_float4(hx, hy, hz, q));
        st_stream(a.posqCorrection + gi, make_float4((float)(x.x - hx), (float)(x.y - hy), (float)(x.z - hz), 0.0f));
    }
};

template <int PREC> __device__ __forceinline__ PosTile<PREC> make_pos(const void* sx, const void* sc);
template <> __device__ __forceinline__ PosTile<0> make_pos<0>(const void* sx, const void*) { PosTile<0> p; p.sx = static_cast<const float4*>(sx); return p; }
template <> __device__ __forceinline__ PosTile<1> make_pos<1>(const void* sx, const void* sc) {
    PosTile<1> p; p.sx = static_cast<const float4*>(sx); p.sc = static_cast<const float4*>(sc); return p;
}
template <> __device__ __forceinline__ PosTile<2> make_pos<2>(const void* sx, const void*) { PosTile<2> p; p.sx = static_cast<const double4*>(sx); return p; }

// index of the big residue that holds `particle`: the last entry of the ascending table that is <= particle
__device__ __forceinline__ int big_index(const int* bigFirst, int numBig, int particle) {
    int lo = 0, hi = numBig - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(bigFirst + mid) <= particle) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Pre-pass for residues that do not fit a tile (calcCOMVelocities, drudeTGNH.cu:82-113, for those residues only): one
// CTA per big residue sums momentum and mass in double, in a fixed order, and leaves {V, M} in bigCom.  fscale != 0:
// velocities as the second-half kernel is about to store them (v + fscale w F).
struct BigComArgs {
    const void* velm;
    const void* force;
    const int* bigFirst;
    const int* bigLast;
    double4* bigCom;
    int paddedN;
    double fscale;
};
template <int FFMT, int PREC>
__global__ void __launch_bounds__(256) tgnh_bigcom_kernel(const __grid_constant__ BigComArgs a) {
    using real = typename Prec<PREC>::real;
    using real4 = typename Prec<PREC>::real4;
    __shared__ double red[4][256];
    pdl_wait();
    const int first = a.bigFirst[blockIdx.x], last = a.bigLast[blockIdx.x];
    const real4* velm = static_cast<const real4*>(a.velm);
    double px = 0.0, py = 0.0, pz = 0.0, M = 0.0;
    for (int i = first + threadIdx.x; i <= last; i += 256) {
        const real4 q = velm[i];
        if (q.w == real(0)) continue;
        const double md = mass_d(q.w, rcp_fast(q.w));
        double vx = q.x, vy = q.y, vz = q.z;
        if (a.fscale != 0.0) {
            const double fw = a.fscale * (double)q.w;
            if (FFMT == 1) {
                const long long* f = static_cast<const long long*>(a.force);
                vx += fw * (double)f[i]; vy += fw * (double)f[a.paddedN + i]; vz += fw * (double)f[2 * (size_t)a.paddedN + i];
            } else {
                const float* f = static_cast<const float*>(a.force);
                vx += fw * (double)f[i]; vy += fw * (double)f[a.paddedN + i]; vz += fw * (double)f[2 * (size_t)a.paddedN + i];
            }
        }
        px += md * vx; py += md * vy; pz += md * vz; M += md;
    }
    red[0][threadIdx.x] = px; red[1][threadIdx.x] = py; red[2][threadIdx.x] = pz; red[3][threadIdx.x] = M;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int c = 0; c < 4; c++) red[c][threadIdx.x] += red[c][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double Mt = red[3][0], inv = Mt > 0.0 ? 1.0 / Mt : 0.0;
        a.bigCom[blockIdx.x] = make_double4(red[0][0] * inv, red[1][0] * inv, red[2][0] * inv, Mt);
    }
    pdl_launch_dependents();
}

// L2 hand-over: velm is written by one streaming launch and read (then overwritten) by the next.  B200's L2
// holds 126 MB, so when consecutive launches walk the tiles in opposite directions the tail of what the
// previous launch wrote is still resident: those reads never reach HBM and the dirty lines are overwritten in
// L2 before they are evicted.  velm stores therefore use the default L2 policy while everything that is touched
// once per step (posq, forces, descriptors) is loaded evict-first / stored streaming.
// Body of every streaming kernel.  Returns true in the one CTA that finished the grid-wide energy reduction (all threads of it).
template <int KIND, int FFMT, bool USE_COM, bool HARDWALL, int PREC, bool BIG>
__device__ __forceinline__ bool stream_body(const StreamArgs& a) {
    using L = SmemLayout<KIND, FFMT, USE_COM, PREC>;
    using St = typename L::Stage;
    using real = typename Prec<PREC>::real;
    using real4 = typename Prec<PREC>::real4;
    using R3 = V3<real>;
    constexpr int NS = L::NSTAGE;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* empty = full + NS;
    int4* tlist = reinterpret_cast<int4*>(smem + L::OFF_TLIST);
    double* ssq = reinterpret_cast<double*>(smem + L::OFF_SCALE);         // s_g^2
    double* seps = ssq + MAX_T;                                           // s_g - 1
    int* smisc = reinterpret_cast<int*>(smem + L::OFF_MISC);
    double* ske = reinterpret_cast<double*>(smem + L::OFF_KE);
    double* swarp = reinterpret_cast<double*>(smem + L::OFF_WARP);
    real4* gvelm = static_cast<real4*>(a.velm);
    real4* gdelta = static_cast<real4*>(a.posDelta);

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
        if (St::HAS_X) bulk_prefetch_l2(static_cast<const unsigned char*>(a.posq) + (size_t)start * St::XB, n * St::XB);
        if (St::HAS_F) {
            const unsigned char* f = static_cast<const unsigned char*>(a.force);
            for (int c = 0; c < 3; c++) bulk_prefetch_l2(f + ((size_t)c * a.paddedN + a0) * St::FBYTES, na * St::FBYTES);
        }
        bulk_prefetch_l2(a.desc + a0, na * 4);
    };
    // Request tile `it` of this CTA into its stage (thread 0 only).  `parts`: 1 = everything but velm / posDelta (arms
    // the barrier with the full byte count), 2 = velm (+ posDelta), 3 = both.  The split exists for the prologue: velm
    // is the only input the previous launch writes, so the rest can be requested before griddepcontrol.wait, while
    // the chain launch that precedes a first-half launch is still running.
    auto issue = [&](int it, int parts) {
        const int4 b = it < TLIST_CAP ? tlist[it] : tile_bounds(it);
        const int start = b.x, end = b.y, r0 = b.z, r1 = b.w;
        const int n = end - start, a0 = start & ~3, na = ((end + 3) & ~3) - a0;
        unsigned char* st = smem + (it % NS) * St::BYTES;
        uint64_t* bar = &full[it % NS];
        if (parts & 1) {
            *reinterpret_cast<int4*>(st + St::OFF_HDR) = make_int4(start, n, r0, r1 - r0);
            uint32_t bytes = n * St::VB + na * 4;
            const int ra0 = r0 & ~3, rna = ((r1 + 1 + 3) & ~3) - ra0;     // residues r0..r1 inclusive (r1 = end marker)
            if (St::HAS_R) bytes += rna * 4;
            if (St::HAS_X) bytes += n * (PREC == 1 ? 32 : St::XB);
            if (St::HAS_P) bytes += n * St::VB;
            if (St::HAS_F) bytes += 3 * na * St::FBYTES;
            mbar_arrive_expect_tx(bar, bytes);
            if (St::HAS_X) {
                bulk_g2s(st + St::OFF_X, static_cast<const unsigned char*>(a.posq) + (size_t)start * St::XB, n * St::XB, bar, polOnce);
                if (PREC == 1) bulk_g2s(st + St::OFF_XC, a.posqCorrection + start, n * 16, bar, polOnce);
            }
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
            bulk_g2s(st + St::OFF_V, gvelm + start, n * St::VB, bar, polOnce);
            if (St::HAS_P) bulk_g2s(st + St::OFF_P, gdelta + start, n * St::VB, bar, polOnce);   // written by the caller's constraint kernels
        }
    };
    // Only inside tgnh_step's own launch sequence (a.earlyLoads) are forces / posq known not to be written by the launches
    // this one may overlap with; behind foreign kernels (OpenMM's force, constraint, virtual-site kernels) every input is
    // requested after the wait.
    if (tid == 0 && a.earlyLoads) {
        for (int it = 0; it < NS && it < myTiles; it++) issue(it, 1);
        for (int it = NS; it < NS + a.prologuePrefetch; it++) prefetch(it);
    }
    pdl_wait();                                         // everything below reads what earlier launches wrote
    if (tid == 0)
        for (int it = 0; it < NS && it < myTiles; it++) issue(it, a.earlyLoads ? 2 : 3);
    if (tid < T) {
        const double sg = (KIND == KIND_B || KIND == KIND_BU || KIND == KIND_A2 || KIND == KIND_K || KIND == KIND_KU) ? 1.0 : a.chain.scaleA[tid];
        ssq[tid] = sg * sg;
        seps[tid] = sg - 1.0;
    }
    __syncthreads();

    const bool doScale = (KIND == KIND_A || KIND == KIND_A1 || KIND == KIND_S) || (KIND == KIND_KE && a.applyScale);
    const real eCOM = doScale ? (real)seps[G] : real(0);
    const real eDrude = doScale ? (real)seps[G + 1] : real(0);
    const real dt = (real)a.dt, fscale = (real)a.fscale, rmax = (real)a.rmax;
    const real rmax2 = rmax * rmax;
    double accCOM = 0.0, accDrude = 0.0;               // this thread's share of the COM-group and Drude-group sums
    double accT = 0.0;                                 // ... and of the group curTg its recent particles belong to (lab-frame kinds)
    int curTg = -1;

    for (int it = 0; it < myTiles; it++) {
        const int stg = it % NS;
        const uint32_t phase = (it / NS) & 1;
        unsigned char* st = smem + stg * St::BYTES;
        const real4* sv = reinterpret_cast<const real4*>(st + St::OFF_V);
        const PosTile<PREC> pos = make_pos<PREC>(st + St::OFF_X, st + St::OFF_XC);
        const unsigned char* sF = st + St::OFF_F;
        const uint32_t* sd = reinterpret_cast<const uint32_t*>(st + St::OFF_D);

        mbar_wait(&full[stg], phase);
        const int4 hdr = *reinterpret_cast<const int4*>(st + St::OFF_HDR);
        const int start = hdr.x, n = hdr.y;
        const int fo = start & 3;                     // offset of the tile inside its 4-aligned window

        const bool active = tid < n;
        real4 v4;
        v4.x = v4.y = v4.z = v4.w = real(0);
        uint32_t d = 0;                               // inactive lanes: ordinary particle, partner = self, residue = self
        if (active) { v4 = sv[tid]; d = sd[fo + tid]; }
        const R3 v = xyz(v4);
        const real w = v4.w;
        const int tg = desc_tg(d);
        const uint32_t role = desc_role(d);
        const bool massive = w != real(0);
        const int pj = tid + desc_partner(d);         // ordinary particles: partner == self
        const real4 vj4 = active ? sv[pj] : v4;
        const R3 vj = xyz(vj4);
        const real wj = vj4.w;
        const real m = massive ? rcp_fast(w) : real(0);
        const real mj = rcp_fast(wj);
        const real invTot = rcp_fast(m + mj);
        const real fi = m * invTot, fj = mj * invTot;        // mass fractions of this particle and of its partner
        R3 rel = vj - v;                              // partner minus me (0 for ordinary particles)
        const real eT = doScale ? (real)seps[tg] : real(0);
        const real coef = (eT - eDrude);              // sT - sDrude

        R3 F = v3(real(0), real(0), real(0)), Fj = F;
        if (St::HAS_F && active) { F = load_force3<FFMT, real>(sF, fo + tid); Fj = load_force3<FFMT, real>(sF, fo + pj); }
        const real fw = fscale * w, fwj = fscale * wj;

        // residue centre-of-mass velocity (calcCOMVelocities, drudeTGNH.cu:86-105); for the second half it is taken
        // after the kick
        R3 V = v3(real(0), real(0), real(0));
        double keC = 0.0;                             // M |V|^2 of this particle's residue (kinds that reduce energies)
        bool firstOfRes = desc_off_first(d) == 0;       // the residue's first particle carries M |V|^2
        const bool big = BIG && desc_big(d);            // BIG: the system has residues larger than a tile (separate instantiation)
        constexpr bool LAB = (KIND == KIND_BU || KIND == KIND_KU);      // lab-frame energies, M |V|^2 per residue subtracted
        if (USE_COM && active && KIND != KIND_A2 && KIND != KIND_K && big) {
            // big residue (a protein, a polymer): V and M from the pre-pass table (tgnh_bigcom_kernel)
            const int b = big_index(a.bigFirst, a.numBig, start + tid);
            const double4 c = a.bigCom[b];
            V = v3((real)c.x, (real)c.y, (real)c.z);
            firstOfRes = a.bigFirst[b] == start + tid;
            keC = c.w * (c.x * c.x + c.y * c.y + c.z * c.z);
        } else if (USE_COM && active && !LAB && KIND != KIND_A2 && KIND != KIND_K) {
            const int j0 = tid - desc_off_first(d), j1 = tid + desc_off_last(d);
            if (!L::HAS_KE && !PREC) {
                // first half, fp32 layout: V only feeds the small corrections (sT-1)(v - V) and (sCOM-1) V, fp32 sums are ample
                R3 P = v3(real(0), real(0), real(0));
                real M = real(0);
                for (int j = j0; j <= j1; j++) {
                    const real4 q = sv[j];
                    const real mq = q.w != real(0) ? rcp_fast(q.w) : real(0);
                    P = axpy(mq, xyz(q), P);
                    M += mq;
                }
                V = rcp_fast(M) * P;
            } else {
                // momentum and mass of the residue in double (see mass_d)
                V3<double> P = v3(0.0, 0.0, 0.0);
                double M = 0.0;
                for (int j = j0; j <= j1; j++) {
                    const real4 q = sv[j];
                    const bool mass = q.w != real(0);
                    const real mq = mass ? rcp_fast(q.w) : real(0);
                    const double mqd = mass ? mass_d(q.w, mq) : 0.0;
                    R3 vq = xyz(q);
                    if (KIND == KIND_B) vq = kicked(vq, fscale * q.w, load_force3<FFMT, real>(sF, fo + j));   // what member j stores
                    P = axpy(mqd, to_double(vq), P);
                    M += mqd;
                }
                const V3<double> Vd = rcp_d(M, rcp_fast((float)M)) * P;
                V = v3((real)Vd.x, (real)Vd.y, (real)Vd.z);
                keC = M * dot3(Vd);
            }
        }

        R3 vn;                                        // this particle's new velocity
        R3 r;                                         // ... relative to the residue (after scaling / kick)
        if (KIND == KIND_K) {
            vn = kicked(v, fw, F);                     // drudeTGNH.cu:314-364
            if (active) st_global(gvelm + start + tid, massive ? pack4(vn, w) : v4);
            r = vn;
        } else if (LAB) {
            // kick (drudeTGNH.cu:314-364; KIND_KU: F = 0, nothing stored); kinetic energies in the lab frame, the residues'
            // M |V|^2 is removed below
            vn = kicked(v, fw, F);
            if (KIND == KIND_BU && active) st_global(gvelm + start + tid, massive ? pack4(vn, w) : v4);
            const R3 vjn = kicked(vj, fwj, Fj);
            rel = vjn - vn;
            r = vn;
            // one thread per residue of the tile; the duty rotates over the warps from tile to tile so that no warp is
            // always the slow one (a stage is recycled only when all 16 warps have left it)
            const int ridx = (tid - it * 128) & (TILE - 1);
            if (USE_COM && active && big && firstOfRes) {   // big residue: M |V|^2 from the pre-pass table
                accCOM += keC;
                ske[tg * TILE + tid] -= keC;
            }
            if (USE_COM && ridx < hdr.w) {
                // P = sum_j m_j v_j (kicked velocities), M = sum_j m_j over the residue's massive members
                const int* sr = reinterpret_cast<const int*>(st + St::OFF_R) + (hdr.z & 3);
                const int j0 = sr[ridx] - start;
                const uint32_t d0 = sd[fo + j0];
                const int j1 = (BIG && desc_big(d0)) ? j0 : sr[ridx + 1] - start;      // segments of big residues: nothing to do here
                V3<double> P = v3(0.0, 0.0, 0.0);
                double M = 0.0;
                for (int j = j0; j < j1; j++) {
                    const real4 q = sv[j];
                    const bool mass = q.w != real(0);
                    const real mq = mass ? rcp_fast(q.w) : real(0);
                    const double mqd = mass ? mass_d(q.w, mq) : 0.0;
                    R3 vq = xyz(q);
                    if (St::HAS_F) vq = kicked(vq, fscale * q.w, load_force3<FFMT, real>(sF, fo + j));   // what member j stores
                    P = axpy(mqd, to_double(vq), P);
                    M += mqd;
                }
                const double keRes = j1 > j0 ? dot3(P) * rcp_d(M, rcp_fast((float)M)) : 0.0;   // M |V|^2 = |P|^2 / M
                accCOM += keRes;
                ske[desc_tg(d0) * TILE + tid] -= keRes;
            }
        } else if (KIND == KIND_B) {
            // integrateDrudeTGNHVelocities (drudeTGNH.cu:314-364), updatePosDelta = false
            vn = kicked(v, fw, F);
            if (active) st_global(gvelm + start + tid, massive ? pack4(vn, w) : v4);
            rel = axpy(fwj, Fj, axpy(-fw, F, rel));
            r = vn - V;
        } else {
            // thermostat scaling (integrateDrudeTGNHChain, drudeTGNH.cu:255-300) in the unified form of the header comment
            r = v - V;
            if constexpr (PREC == 0) vn = scaled_velocity2(v, split2(seps[tg]), r, split2(seps[G]), V, coef * fj, rel);
            else vn = scaled_velocity(v, eT, r, eCOM, V, coef * fj, rel);
        }

        if (LAB) {
            // computeNormalizedKineticEnergies (drudeTGNH.cu:152-188) for residue-uniform groups, without the pair transform:
            // M_p |v_cm|^2 + mu |rel|^2 = m_i |v_i|^2 + m_j |v_j|^2 for a Drude pair, so EVERY particle adds its own
            // m |v|^2 to its group and only the Drude particle moves the pair's internal term mu |rel|^2 from the group
            // to the Drude thermostat; 1/mu = 1/m_i + 1/m_j = w_i + w_j needs no masses at all
            const double md = massive ? mass_d(w, m) : 0.0;
            double ke = md * (double)dot3(r);
            if (role == ROLE_DRUDE) {
                const double wsum = (double)w + (double)wj;
                const double mu = (massive && wj != real(0)) ? rcp_d(wsum, rcp_fast((float)wsum)) : 0.0;
                const double keD = mu * (double)dot3(rel);
                ke -= keD;
                accDrude += keD;
            }
            // this thread's particles usually stay in one group from tile to tile (molecule-periodic group patterns): the
            // running sum lives in a register and moves to its shared-memory column only when the group changes
            if (active && massive) {
                if (tg != curTg) {
                    if (curTg >= 0) ske[curTg * TILE + tid] += accT;
                    accT = 0.0;
                    curTg = tg;
                }
                accT += ke;
            }
        } else if (L::HAS_KE) {
            // computeNormalizedKineticEnergies (drudeTGNH.cu:152-188), branch-free: an ordinary particle is its own
            // "pair centre of mass" (rel = 0); the Drude particle of a pair carries the pair's two terms
            const R3 cm = axpy(fj, rel, r);           // pair COM relative to the residue
            const bool isDrude = role == ROLE_DRUDE;
            const double md = massive ? mass_d(w, m) : 0.0, mjd = mass_d(wj, mj);
            const double Mp = md + mjd;
            const double massT = isDrude ? Mp : (role == ROLE_NORMAL ? md : 0.0);
            const double s2T = doScale ? ssq[tg] : 1.0, s2D = doScale ? ssq[G + 1] : 1.0, s2C = doScale ? ssq[G] : 1.0;
            const double keT = massT * s2T * (double)dot3(cm);
            const double keD = isDrude ? md * mjd * rcp_d(Mp, (float)invTot) * s2D * (double)dot3(rel) : 0.0;   // reduced mass
            if (!(USE_COM && !LAB && active && firstOfRes)) keC = 0.0;        // the residue's first particle carries M |V|^2
            keC *= s2C;
            if (active && massT != 0.0) ske[tg * TILE + tid] += keT;
            accDrude += keD;
            accCOM += keC;
        }

        if (KIND == KIND_KE || KIND == KIND_S) {
            if (doScale && active) st_global(gvelm + start + tid, massive ? pack4(vn, w) : v4);
        } else if (KIND == KIND_A1) {
            // scaling + half kick; posDelta = dt * v for OpenMM's constraint kernels (drudeTGNH.cu:322-324, 360-363)
            vn = kicked(vn, fw, F);
            if (active && massive) {
                st_global(gvelm + start + tid, pack4(vn, w));
                st_global(gdelta + start + tid, pack4(dt * vn, real(0)));
            }
        } else if (KIND == KIND_A2 && active) {
            // integrateDrudeTGNHPositions (drudeTGNH.cu:438-465): x += delta, v = delta / dt; then the hard wall (:474-573)
            const real4* sp = reinterpret_cast<const real4*>(st + St::OFF_P);
            const real invDt = real(1) / dt;
            typename PosTile<PREC>::Q q;
            const R3 dl = xyz(sp[tid]), x = pos.load(tid, q);
            R3 xn = x + dl;
            vn = invDt * dl;
            if (HARDWALL && role != ROLE_NORMAL) {
                typename PosTile<PREC>::Q qj;
                const R3 dj = xyz(sp[pj]), xj = pos.load(pj, qj);
                R3 vjn = invDt * dj;
                const R3 delta = (x - xj) + (dl - dj);
                const real r2 = dot3(delta);
                if (r2 > rmax2) {
                    R3 xjn = xj + dj;
                    if (role == ROLE_DRUDE) hard_wall(delta, r2, xn, xjn, vn, vjn, w, wj, rmax, (real)a.hardwallScale, dt);
                    else hard_wall(-delta, r2, xjn, xn, vjn, vn, wj, w, rmax, (real)a.hardwallScale, dt);
                }
            }
            if (massive) {
                st_global(gvelm + start + tid, pack4(vn, w));
                PosTile<PREC>::store(a, start + tid, xn, q);
            } else if (PREC != 1) {
                // a massless particle is not integrated; its slots get their own bits back so that whole sectors are written
                // (tgnh_v2.cuh; not in the mixed layout, where re-splitting a position may move bits between posq and its correction)
                st_global(gvelm + start + tid, v4);
                PosTile<PREC>::store(a, start + tid, x, q);
            }
        } else if (KIND == KIND_A && active) {
            // half kick + drift (+ hard wall) (drudeTGNH.cu:314-364, 438-465, 474-573)
            vn = kicked(vn, fw, F);
            typename PosTile<PREC>::Q q;
            const R3 x = pos.load(tid, q);
            R3 xn = axpy(dt, vn, x);
            if (HARDWALL && role != ROLE_NORMAL) {
                // the partner's update, recomputed here so that both threads of a pair see the same wall test;
                // seen from the partner, rel changes sign and the mass fraction is this particle's
                const R3 rj = vj - V;
                R3 vjn;
                if constexpr (PREC == 0) vjn = kicked(scaled_velocity2(vj, split2(seps[tg]), rj, split2(seps[G]), V, coef * fi, -rel), fwj, Fj);
                else vjn = kicked(scaled_velocity(vj, eT, rj, eCOM, V, coef * fi, -rel), fwj, Fj);
                typename PosTile<PREC>::Q qj;
                const R3 xj = pos.load(pj, qj);
                // displacement from the exact difference of the old positions plus the relative drift: avoids the
                // cancellation of two rounded box-sized coordinates in the wall test
                const R3 delta = axpy(dt, vn - vjn, x - xj);
                const real r2 = dot3(delta);
                if (r2 > rmax2) {                     // rInv*maxDrudeDistance < 1  (drudeTGNH.cu:490)
                    R3 xjn = axpy(dt, vjn, xj);
                    if (role == ROLE_DRUDE) hard_wall(delta, r2, xn, xjn, vn, vjn, w, wj, rmax, (real)a.hardwallScale, dt);
                    else hard_wall(-delta, r2, xjn, xn, vjn, vn, wj, w, rmax, (real)a.hardwallScale, dt);
                }
            }
            if (massive) {
                st_global(gvelm + start + tid, pack4(vn, w));
                PosTile<PREC>::store(a, start + tid, xn, q);
            } else if (PREC != 1) {
                // a massless particle is not integrated; its slots get their own bits back so that whole sectors are written
                // (tgnh_v2.cuh; not in the mixed layout, where re-splitting a position may move bits between posq and its correction)
                st_global(gvelm + start + tid, v4);
                PosTile<PREC>::store(a, start + tid, x, q);
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
    if (KIND == KIND_S && blockIdx.x == 0 && tid < T) {
        // the factors are applied: the kinetic energies of the stored velocities are s_g^2 KE_g (residue-uniform groups)
        const double sg = a.chain.scaleA[tid];
        a.chain.ke2[tid] *= sg * sg;
        a.chain.pending[tid] = 1.0;
    }
    if (!L::HAS_KE) return false;

    // ---- deterministic reduction: thread columns -> warp -> CTA -> (last CTA) grid ----
    if (curTg >= 0) ske[curTg * TILE + tid] += accT;
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
    if (!smisc[0]) return false;
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
    if (a.peers.world > 1) {
        // sharded: hand this rank's sums to every rank (this one included) over NVLink
        __syncthreads();
        peer_publish(a.peers, out, T, tid, (int)blockDim.x);
    }
    if (tid == 0) *a.ticket = 0u;
    if (KIND == KIND_KE && a.applyScale && tid < T) a.chain.pending[tid] = 1.0;   // the deferred scaling is now applied
    return true;
}

template <int KIND, int FFMT, bool USE_COM, bool HARDWALL, int PREC, bool BIG>
__global__ void __launch_bounds__(TILE, PREC ? 1 : 2) tgnh_stream_kernel(const __grid_constant__ StreamArgs a) {
    stream_body<KIND, FFMT, USE_COM, HARDWALL, PREC, BIG>(a);
}

// Small systems (every CTA owns at most one tile; launch hand-over, not bandwidth, sets the step time): the CTA that
// finishes the energy reduction runs the Nose-Hoover chain update itself instead of a separate chain launch.  One CTA per
// SM: the chain needs more registers than the streaming body and occupancy is irrelevant at this size.
template <int KIND, int FFMT, bool USE_COM, int PREC>
__global__ void __launch_bounds__(TILE, 1) tgnh_stream_chain_kernel(const __grid_constant__ StreamArgs a) {
    if (!stream_body<KIND, FFMT, USE_COM, false, PREC, false>(a)) return;
    __syncthreads();                                   // the energy vector written by this CTA's warps
    if (threadIdx.x < 32) chain_phase(a.chain, a.fusedChainMode, threadIdx.x);
}

// The Nose-Hoover chain update(s) between two streaming launches: one warp, lane g = thermostat g.
__global__ void __launch_bounds__(32, 1) tgnh_chain_kernel(const __grid_constant__ ChainView c, const __grid_constant__ PeerView peers, int mode) {
    pdl_launch_dependents();      // the next streaming launch may start its prologue; it waits for our results
    pdl_wait();
    if (peers.world > 1) peer_gather(peers, c.ke2, c.T, threadIdx.x);    // sharded: global sums from all ranks' partials
    if (mode != CHAIN_NONE) chain_phase(c, mode, threadIdx.x);
}

}  // namespace tgnh
