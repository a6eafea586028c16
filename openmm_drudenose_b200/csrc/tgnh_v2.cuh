// Second generation of the streaming kernels of the TGNH step ("warp-chunk" kernels), single-precision layout.
//
// Same operators as tgnh_kernels.cuh (reference: platforms/cuda/src/kernels/drudeTGNH.cu)
//   V2_A   first half   = integrateDrudeTGNHChain + integrateDrudeTGNHVelocities(updatePosDelta) +
//                         integrateDrudeTGNHPositions + applyHardWallConstraints          (:249-301, 307-365, 435-466, 471-574)
//   V2_B   second half  = integrateDrudeTGNHVelocities + calcCOMVelocities + normalizeVelocities +
//                         computeNormalizedKineticEnergies + sumNormalizedKineticEnergies (:307-365, 82-133, 138-242)
//   V2_KE  reduce       = the kinetic-energy reduction alone                              (:82-242)
//   V2_S   scale        = integrateDrudeTGNHChain alone                                   (:249-301)
// with a different decomposition of the work, chosen from the round-1 profile of the second-half kernel (58.8 M warp
// instructions, 40 % XU pipe, 7.5 M shared-memory bank conflicts, all warps of a CTA coupled through the slowest one):
//
//  * WARP-CHUNKS.  The particles are cut into residue-aligned chunks of at most 32 consecutive particles, one per warp
//    and tile (14 chunks = at most 448 particles per tile).  A Drude pair lies inside a residue, a residue inside a
//    chunk, so the pair partner and all members of the residue are lanes of the same warp: partner values travel by
//    one shuffle, residue sums by a segmented warp scan.  No thread reads another warp's particles, nothing is gathered
//    from shared memory with a stride (no bank conflicts), and no warp waits for another one inside a tile.
//  * SPECIES TABLE.  Everything that is constant per kind of particle — mass, reduced mass of its Drude pair, mass
//    fraction of the partner, inverse mass of its residue, temperature group, role, partner offset, "first particle of
//    the residue" — sits in a table of at most 255 rows in shared memory, indexed by ONE BYTE per particle (instead
//    of a 4-byte descriptor word per particle and kernel).  Masses are stored as float pairs (hi + lo = the double the
//    caller passed in System::getParticleMass): the kinetic-energy sums are unbiased per species (a mass rounded to
//    fp32 would shift a thermostat's energy systematically and the Nose-Hoover chain integrates that), yet the inner
//    loop contains no reciprocal, no float<->double conversion and no fp64 instruction at all.
//  * PRODUCER WARPS.  Warps 14 and 15 only stream: a producer waits for a stage to be released by the 14 consumer warps and requests
//    the next tile (cp.async.bulk / UBLKCP completing on the stage's mbarrier).  No consumer ever blocks on a refill.  The ~60
//    serial instructions of a refill are on the critical path of the ring in the kernels that move little data per tile, so in
//    the second-half, reduction and scaling kernels the two producers take every other tile; the HBM-bound first-half kernel
//    runs best with one producer at its own pace (the other idles).
//  * fp32 running sums per thread (the thread's group sum moves to its fp32 column when the group changes); fp64 from
//    the end of the tile loop on (warp shuffles -> CTA -> fixed-order sum over the CTAs by the last one), bit-reproducible
//    run to run.
//  * Round-2 profile of the first version (57 M warp instructions, issue slots 62 % busy, DRAM 53 %): the position of a
//    particle inside its residue comes from the species row instead of a ballot / bit-scan chain per tile, residues of one
//    power-of-two size are summed by a butterfly (the sum lands in every lane), the energy columns are fp32 (shared
//    memory for a third resident CTA per SM).
//
// Kinetic energies, in the reference's own form (:152-188) with V the residue's centre-of-mass velocity:
//     group tg:   sum_i m_i |v_i - V|^2  -  sum_pairs mu |v_d - v_p|^2      (= sum (m_d+m_p) |cm - V|^2 + normal particles)
//     COM group:  sum_res |P|^2 / M ,  P = sum_i m_i v_i
//     Drude:      sum_pairs mu |v_d - v_p|^2
// valid whether or not a residue lies in one temperature group.
#pragma once
#include "tgnh_kernels.cuh"

namespace tgnh {

enum { V2_A = 0, V2_B = 1, V2_KE = 2, V2_S = 3 };
#ifndef TGNH_V2_NCONS
#define TGNH_V2_NCONS 14
#endif
#ifndef TGNH_V2_NS_A
#define TGNH_V2_NS_A 3
#endif
#ifndef TGNH_V2_NS_B
#define TGNH_V2_NS_B 6
#endif
#ifndef TGNH_V2_NS_KE
#define TGNH_V2_NS_KE 6
#endif
#ifndef TGNH_V2_PF
#define TGNH_V2_PF 4
#endif
#ifndef TGNH_V2_PF_A
#define TGNH_V2_PF_A 0
#endif
constexpr int V2_PF_A = TGNH_V2_PF_A;         // the same for the first-half kernel
constexpr int V2_PF = TGNH_V2_PF;             // the producer fetches the chunk bounds of a tile this many tiles ahead
constexpr int V2_NCONS = TGNH_V2_NCONS;       // consumer warps per CTA (the last warp is the producer)
#ifndef TGNH_V2_NPROD
#define TGNH_V2_NPROD 2
#endif
constexpr int V2_NPROD = TGNH_V2_NPROD;       // producer warps (each refills every V2_NPROD-th tile)
constexpr int V2_THREADS = (V2_NCONS + V2_NPROD) * 32;
constexpr int V2_CTAS = V2_NCONS > 15 ? 1 : 2;   // resident CTAs per SM the kernels are compiled for
#ifndef TGNH_V2_CTAS_RED
#define TGNH_V2_CTAS_RED V2_CTAS
#endif
constexpr int V2_CTAS_RED = TGNH_V2_CTAS_RED;    // the same for the reducing kinds (second half, reduction): with a ring deeper than half an SM's
                                                 // shared memory they run one CTA per SM and may use its whole register file
constexpr int V2_TILE = V2_NCONS * 32;        // particles per tile, at most
constexpr int V2_FW = V2_TILE + 8;            // 4-aligned window of force components that covers any tile
constexpr int V2_SW = V2_TILE + 32;           // 16-aligned window of species bytes that covers any tile
constexpr int V2_MAX_SPECIES = 255;           // species rows; the table's last row (index = number of species) is "no particle"
constexpr int V2_ROW_F4 = 3;                  // float4 per row: q0, q1, pad (48 B stride: neighbouring rows fall into different banks)

// species-table row
//   q0 = { m_hi, m_lo, 1/M_res (hi), meta }          meta: [4:0] temperature group, [6:5] role, [12:7] signed offset to the pair
//   q1 = { mu_hi, mu_lo, 1/M_res (lo), f_partner }         partner (0 = none), [17:13] offset back to the first particle of the
//                                                          residue, [22:18] offset forward to its last particle
__host__ __device__ inline uint32_t v2_meta_pack(int tg, uint32_t role, int partner, int offFirst, int offLast) {
    return (uint32_t)(tg & 31) | (role << 5) | ((uint32_t)(partner & 63) << 7) | ((uint32_t)offFirst << 13) | ((uint32_t)offLast << 18);
}
__device__ __forceinline__ int v2_tg(uint32_t m) { return m & 31; }
__device__ __forceinline__ uint32_t v2_role(uint32_t m) { return (m >> 5) & 3; }
__device__ __forceinline__ int v2_partner(uint32_t m) { return ((int32_t)(m << 19)) >> 26; }
__device__ __forceinline__ int v2_off_first(uint32_t m) { return (m >> 13) & 31; }
__device__ __forceinline__ int v2_off_last(uint32_t m) { return (m >> 18) & 31; }

template <int KIND, int FFMT>
struct V2Layout {
    static constexpr bool HAS_X = (KIND == V2_A);
    static constexpr bool HAS_F = (KIND == V2_A || KIND == V2_B);
    static constexpr bool HAS_KE = (KIND == V2_B || KIND == V2_KE);
    static constexpr int FBYTES = FFMT == 1 ? 8 : 4;
    // ring depth; 6 stages of the second half with 8-byte forces would leave room for only one CTA per SM
    static constexpr int NSTAGE = V2_NCONS > 15 ? (HAS_X ? 4 : 6) : (HAS_X ? TGNH_V2_NS_A : (HAS_F && FFMT == 1) ? 4 : HAS_F ? TGNH_V2_NS_B : TGNH_V2_NS_KE);
    static constexpr int OFF_V = 0;
    static constexpr int OFF_X = OFF_V + V2_TILE * 16;
    static constexpr int OFF_F = OFF_X + (HAS_X ? V2_TILE * 16 : 0);
    static constexpr int OFF_S = OFF_F + (HAS_F ? 3 * V2_FW * FBYTES : 0);
    static constexpr int OFF_HDR = OFF_S + V2_SW;            // uint2 per consumer warp: { first particle of the tile, chunk offset | chunk length << 16 }
    static constexpr int STAGE = OFF_HDR + ((8 * V2_NCONS + 15) & ~15);
    static constexpr int OFF_BAR = NSTAGE * STAGE;            // full[NS], empty[NS]
    static constexpr int OFF_SCALE = OFF_BAR + ((2 * NSTAGE * 8 + 127) & ~127);           // double[MAX_T] s^2 (unused), double[MAX_T] s - 1
    static constexpr int OFF_MISC = OFF_SCALE + MAX_T * 16;
    static constexpr int OFF_WARP = OFF_MISC + 16;            // double[T][32]
    static constexpr int OFF_KE = OFF_WARP + (HAS_KE ? MAX_T * 32 * 8 : 0);        // float[T][V2_TILE]: per-thread group sums
    __host__ __device__ static int off_table(int T) { return OFF_KE + (HAS_KE ? T * V2_TILE * 4 : 0); }   // then the species table: rows x 48 B
    static int bytes(int T, int rows) { return off_table(T) + rows * V2_ROW_F4 * 16; }
};

template <int FFMT>
__device__ __forceinline__ V3<float> v2_force(const unsigned char* sF, int idx) {
    if (FFMT == 1) {
        const long long* f = reinterpret_cast<const long long*>(sF);
        return v3((float)f[idx], (float)f[V2_FW + idx], (float)f[2 * V2_FW + idx]);
    }
    const float* f = reinterpret_cast<const float*>(sF);
    return v3(f[idx], f[V2_FW + idx], f[2 * V2_FW + idx]);
}

// (hi + lo) * x with one rounding of the large term: hi * x + lo * x
__device__ __forceinline__ float mul2(float hi, float lo, float x) { return fmaf(hi, x, lo * x); }

// Inclusive segmented scan of a vector over the lanes of a warp.  Segments are the residues of the chunk (contiguous
// lanes, the first lane of each flagged in firstMask); on return the LAST lane of every segment holds the segment's sum.
__device__ __forceinline__ void seg_scan(V3<float>& p, int lane, int segStart, int maxRes) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        if (o < maxRes) {                               // warp-uniform: no residue of this system is longer
            const float tx = __shfl_up_sync(0xffffffffu, p.x, o), ty = __shfl_up_sync(0xffffffffu, p.y, o), tz = __shfl_up_sync(0xffffffffu, p.z, o);
            if (lane - o >= segStart) { p.x += tx; p.y += ty; p.z += tz; }
        }
    }
}

// Sum of `p` over the residue of every lane, delivered to ALL its lanes.  bfly > 0: every residue of the system has `bfly` particles (a
// power of two) and chunks are full, so residues are aligned lane groups and a butterfly does it; otherwise a segmented scan
// followed by a read of the segment's last lane.
__device__ __forceinline__ V3<float> residue_sum(V3<float> p, int lane, int offFirst, int offLast, int maxRes, int bfly) {
    auto level = [&](int o) { p.x += __shfl_xor_sync(0xffffffffu, p.x, o); p.y += __shfl_xor_sync(0xffffffffu, p.y, o); p.z += __shfl_xor_sync(0xffffffffu, p.z, o); };
    if (bfly == 4) { level(1); level(2); return p; }          // the commonest case first: one warp-uniform test, no per-level tests
    if (bfly > 0) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
            if (o < bfly) level(o);
        return p;
    }
    seg_scan(p, lane, lane - offFirst, maxRes);
    const int last = lane + offLast;
    return v3(__shfl_sync(0xffffffffu, p.x, last), __shfl_sync(0xffffffffu, p.y, last), __shfl_sync(0xffffffffu, p.z, last));
}

// One residue of K consecutive particles of a stage, all of it in one lane (the residue-per-lane form of the reducing kinds, see
// v2_body): half kick of every member (second half; HAS_F) and the residue's energy terms
//   e = sum_i m_i |v_i|^2,  keC = |P|^2 / M with P = sum_i m_i v_i  (:86-105, :154),  keD = sum over its Drude particles of mu |rel|^2 (:185).
//
// Shared-memory banks.  Lane L reads particle K (32 slot + L) + m: for odd K the 16-byte velocity reads of a quarter-warp and the
// 4-byte force reads of the warp fall into distinct banks by themselves; for even K they would collide (K = 4: four-way for
// both; the first version of this path spent its time there: 9.8 M conflicts in 15.7 M wavefronts, mio_throttle the largest
// stall).  So for even K the lanes walk the members of their residue in ROTATED order, lane L starting at member rpl_rot<K>(L):
//   K = 2, 6:  ((L >> 2) + (L >> 4)) & 1        K = 4:  ((L >> 1) + (L >> 3)) & 3        K = 8:  (L ^ (L >> 2)) & 7
// which makes both kinds of reads conflict-free.  The lanes of a warp are then at different members at the same time, so the
// pair term is not formed where the Drude particle is met (divergent) but once after the loop, from the two velocities kept in
// registers as they go by (even K: at most one Drude pair per residue, checked by tgnh_create).
template <int K> __device__ __forceinline__ int rpl_rot(int lane) {
    if (K == 2 || K == 6) return ((lane >> 2) + (lane >> 4)) & 1;
    if (K == 4) return ((lane >> 1) + (lane >> 3)) & 3;
    if (K == 8) return (lane ^ (lane >> 2)) & 7;
    return 0;
}
struct RplOut { float e, keC, keD; int tg; };
template <int K, int FFMT, bool HAS_F>
__device__ __forceinline__ RplOut rpl_residue(const float4* __restrict__ sv, const unsigned char* __restrict__ sF, int fo, const unsigned char* __restrict__ sS,
                                              const float4* __restrict__ stab, float fscale, int lane) {
    constexpr bool ROT = (K % 2) == 0;
    V3<float> P = v3(0.f, 0.f, 0.f);
    RplOut o;
    o.e = 0.f; o.keD = 0.f; o.tg = 0;
    float invMhi = 0.f, invMlo = 0.f;
    V3<float> vd = v3(0.f, 0.f, 0.f), vp = vd;          // ROT: kicked velocities of the residue's Drude particle and of its parent
    int rowD = -1;
    int mm = ROT ? rpl_rot<K>(lane) : 0;
#pragma unroll
    for (int j = 0; j < K; j++) {
        const int m = ROT ? mm : j;                       // the member this lane handles now (a compile-time constant without rotation)
        const int row = sS[m] * V2_ROW_F4;
        const float4 r0 = stab[row];
        const float4 v4 = sv[m];
        V3<float> vn = xyz(v4);
        if (HAS_F) vn = kicked(vn, fscale * v4.w, v2_force<FFMT>(sF, fo + m));
        const uint32_t meta = __float_as_uint(r0.w);
        const uint32_t role = v2_role(meta);
        if (m == 0) { o.tg = v2_tg(meta); invMhi = r0.z; invMlo = stab[row + 1].z; }
        const V3<float> pm = v3(mul2(r0.x, r0.y, vn.x), mul2(r0.x, r0.y, vn.y), mul2(r0.x, r0.y, vn.z));
        P = P + pm;
        o.e += pm.x * vn.x + pm.y * vn.y + pm.z * vn.z;
        if (ROT) {
            if (role == ROLE_DRUDE) { vd = vn; rowD = row; }
            if (role == ROLE_PARENT) vp = vn;
            mm = (mm + 1 == K) ? 0 : mm + 1;
        } else if (role == ROLE_DRUDE) {
            // mu |v_partner - v_d|^2 (:185); the partner's velocity kicked by the same fp32 operation that its own turn applies
            const int jp = j + v2_partner(meta);
            const float4 u4 = sv[jp];
            V3<float> un = xyz(u4);
            if (HAS_F) un = kicked(un, fscale * u4.w, v2_force<FFMT>(sF, fo + jp));
            const float4 r1 = stab[row + 1];
            o.keD += mul2(r1.x, r1.y, dot3(un - vn));
        }
    }
    if (ROT && rowD >= 0) {
        const float4 r1 = stab[rowD + 1];
        o.keD = mul2(r1.x, r1.y, dot3(vp - vd));          // mu |v_parent - v_d|^2 (:185)
    }
    o.keC = mul2(invMhi, invMlo, dot3(P));          // |P|^2 / M  (:154)
    return o;
}

// Body of the warp-chunk kernels.  Returns true in the one CTA that finished the grid-wide energy reduction (all threads of it).
template <int KIND, int FFMT, bool USE_COM, bool HARDWALL, bool RPL_OK = true>
__device__ __forceinline__ bool v2_body(const StreamArgs& a) {
    using L = V2Layout<KIND, FFMT>;
    constexpr int NS = L::NSTAGE;
    constexpr bool IS_A = (KIND == V2_A || KIND == V2_S);        // the kinds that apply the thermostat factors
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + L::OFF_BAR);
    uint64_t* empty = full + NS;
    float2* seps = reinterpret_cast<float2*>(smem + L::OFF_SCALE);              // s_g - 1 as a float pair (hi, lo)
    float* scoef = reinterpret_cast<float*>(smem + L::OFF_SCALE) + 2 * MAX_T;    // s_g - s_Drude
    int* smisc = reinterpret_cast<int*>(smem + L::OFF_MISC);
    double* swarp = reinterpret_cast<double*>(smem + L::OFF_WARP);
    float* ske = reinterpret_cast<float*>(smem + L::OFF_KE);
    float4* gvelm = static_cast<float4*>(a.velm);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = a.chain.T, G = a.chain.G;
    float4* stab = reinterpret_cast<float4*>(smem + L::off_table(T));
    const int myTiles = blockIdx.x < a.numTiles ? (a.numTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto tile_of = [&](int it) {
        const int t = blockIdx.x + it * gridDim.x;
        return a.tileBegin + (a.reverse ? a.numTiles - 1 - t : t);
    };

    // Residue-per-lane form of the reducing kinds (see the consumer loop): a tile is consumed by `rplSlots` warps instead of all
    // of them, the groups of `rplSlots` warps take the tiles in turn.
    constexpr bool CAN_RPL = RPL_OK && USE_COM && (KIND == V2_B || KIND == V2_KE);      // (not compiled into the one-tile-per-CTA kernels of small systems)
    const int rpl = CAN_RPL ? a.resPerLane : 0;
    const int rplSlots = rpl ? (V2_NCONS * (32 / rpl) + 31) / 32 : V2_NCONS;       // warps that share the residues of one tile
    // (a stage must always be consumed by the same group — a group that met a stage whose previous tile another group has not even
    // seen arrive would take that tile's pending phase for its own — so the number of groups divides the ring depth)
    int rplGroups = V2_NCONS / rplSlots;
    while (NS % rplGroups) rplGroups--;
    if (tid == 0) {
        for (int s = 0; s < NS; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], rplSlots); }
        fence_mbar_init();
    }
    __syncthreads();                                    // the barriers exist: the producer may start streaming

    const uint64_t polOnce = policy_evict_first();
    // velm of a second half that stores nothing (lazy second kick): the next first half starts on the tiles this launch reads last.
    // Those — and only those: what the 126 MB L2 can still hold when the first half gets there, measured best at ~48 MB of velm —
    // are read with evict_last; the rest of velm streams through like everything else (evict_first), which alone makes this
    // kernel 4 us faster at 10 M particles.  Measured (C4, us per step / first half / second half): all evict_normal 182.7 / 115.2 /
    // 65.4; all evict_first 190.3 / 126.2 / 62.2; last 30 % evict_last 180.4 / 115.1 / 63.2; last 25 % evict_normal 183.6 / 121.0 / 62.0.
#ifndef TGNH_V2_LAZY_KEEP_MB
#define TGNH_V2_LAZY_KEEP_MB 48
#endif
    const uint64_t polKeep = (KIND == V2_B && a.lazyKick) ? policy_evict_last() : polOnce;
    const int keepTiles = (int)(((size_t)TGNH_V2_LAZY_KEEP_MB << 20) / ((size_t)gridDim.x * V2_TILE * 16)) + 1;     // per CTA
    const int keepFrom = myTiles - keepTiles;
#ifndef TGNH_V2_A_STORE_KEEP_MB
#define TGNH_V2_A_STORE_KEEP_MB 0
#endif
#if TGNH_V2_A_STORE_KEEP_MB > 0
    const uint64_t polStoreKeep = policy_evict_last();
    const int storeKeepFrom = myTiles - ((int)(((size_t)TGNH_V2_A_STORE_KEEP_MB << 20) / ((size_t)gridDim.x * V2_TILE * 16)) + 1);
#endif
    // Producer: request tile `it` of this CTA into its stage (the whole warp calls; lane 0 issues the copies).
    // parts: 1 = header + everything no launch of this library writes (posq, forces, species bytes), arms the barrier with
    // the full byte count; 2 = velm; 3 = both.  `cs` = chunkStart[V2_NCONS * tile + lane] for lanes 0..V2_NCONS.
    auto issue = [&](int it, int parts, int cs) {
        const int start = __shfl_sync(0xffffffffu, cs, 0), end = __shfl_sync(0xffffffffu, cs, V2_NCONS);
        const int n = end - start;
        unsigned char* st = smem + (it % NS) * L::STAGE;
        uint64_t* bar = &full[it % NS];
        const int f0 = start & ~3, fn = ((end + 3) & ~3) - f0;            // 16-byte aligned windows
        const int s0 = start & ~15, sn = ((end + 15) & ~15) - s0;
        if (parts & 1) {
            const int csNext = __shfl_down_sync(0xffffffffu, cs, 1);
            if (lane < V2_NCONS)            // one 8-byte word per consumer warp: a single shared-memory load tells it where its chunk is
                reinterpret_cast<uint2*>(st + L::OFF_HDR)[lane] = make_uint2((unsigned int)start, (unsigned int)(cs - start) | ((unsigned int)(csNext - cs) << 16));
            __syncwarp();
            if (lane == 0) {
                uint32_t bytes = n * 16 + sn;
                if (L::HAS_X) bytes += n * 16;
                if (L::HAS_F) bytes += 3 * fn * L::FBYTES;
                mbar_arrive_expect_tx(bar, bytes);
                if (L::HAS_X) bulk_g2s(st + L::OFF_X, static_cast<const float4*>(a.posq) + start, n * 16, bar, polOnce);
                if (L::HAS_F) {
                    const unsigned char* f = static_cast<const unsigned char*>(a.force);
                    for (int c = 0; c < 3; c++)
                        bulk_g2s(st + L::OFF_F + c * V2_FW * L::FBYTES, f + ((size_t)c * a.paddedN + f0) * L::FBYTES, fn * L::FBYTES, bar, polOnce);
                }
                bulk_g2s(st + L::OFF_S, a.spec + s0, sn, bar, polOnce);
            }
        }
        if ((parts & 2) && lane == 0) bulk_g2s(st + L::OFF_V, gvelm + start, n * 16, bar, it >= keepFrom ? polKeep : polOnce);
    };
    auto chunk_bounds = [&](int it) { return lane <= V2_NCONS ? __ldg(a.chunkStart + V2_NCONS * tile_of(it) + lane) : 0; };

    const bool producer = warp >= V2_NCONS;
    const int pidx = warp - V2_NCONS;                   // which producer
    // The first-half kernel is bound by HBM and runs best with ONE producer that refills at its own pace (any eagerness of the
    // producer costs it 3 us: prefetched chunk bounds +3.1, two producers +2.8); in the second-half, reduction and scaling kernels
    // the producer's ~60 serial instructions per tile are on the critical path of the ring, and two producers that take every
    // other tile are worth 4.7 us of 66.
    constexpr int NPROD = L::HAS_X ? 1 : V2_NPROD;
    // Every stage must have ONE producer (and one consumer group): a producer that polls a stage's `empty` barrier has then issued the
    // stage's previous tile itself, so the barrier is at most one phase behind and the parity test is unambiguous.  Producers take
    // every NPROD-th tile, so the ring depth has to be a multiple of NPROD (an odd depth with two producers refills stages that are
    // still being read: measured).
    static_assert(NS % NPROD == 0, "ring depth must be a multiple of the number of producer warps");
    const int preloaded = myTiles < NS ? myTiles : NS;
    if (producer && pidx > 0) pdl_wait();
    else if (producer) {
        // The first NS tiles are requested while the consumer warps still copy the species table and clear their energy columns.
        // Inputs that no earlier launch of this stream can still be writing may even be requested before griddepcontrol.wait (only
        // when the host knows that the preceding launches are this library's own and do not write them: a.earlyLoads).
        // (the bounds of all NS tiles are fetched before the first copy is issued: NS independent loads, one latency)
        int cs0[NS];
#pragma unroll
        for (int it = 0; it < NS; it++) cs0[it] = it < preloaded ? chunk_bounds(it) : 0;
        if (a.earlyLoads) {
#pragma unroll
            for (int it = 0; it < NS; it++)
                if (it < preloaded) issue(it, 1, cs0[it]);
        }
        pdl_wait();                                     // everything below reads what earlier launches wrote
#pragma unroll
        for (int it = 0; it < NS; it++)
            if (it < preloaded) issue(it, a.earlyLoads ? 2 : 3, cs0[it]);
    } else {
        for (int i = tid; i < a.tableRows * V2_ROW_F4; i += V2_TILE) stab[i] = __ldg(a.specTable + i);     // static table: safe before the wait
        if (L::HAS_KE)
            for (int g = 0; g < T; g++) ske[g * V2_TILE + tid] = 0.f;
        pdl_wait();
        if (tid < T) {
            const double e = (IS_A ? a.chain.scaleA[tid] : 1.0) - 1.0, eD = (IS_A ? a.chain.scaleA[T - 1] : 1.0) - 1.0;
            const F2 f = split2(e);
            seps[tid] = make_float2(f.hi, f.lo);
            scoef[tid] = (float)(e - eD);
        }
    }
    __syncthreads();

    float accT = 0.f, accCOM = 0.f, accDrude = 0.f;    // this thread's running sums: its current group, the COM group, the Drude group
    int curTg = -1;

    if (producer) {
        // The chunk bounds of a tile are a dependent global load (DRAM latency under load ~ the time a CTA spends on one tile of the
        // second half): they are fetched V2_PF tiles ahead so that the refill of a released stage never waits for them.
        constexpr int PF = L::HAS_X ? V2_PF_A : V2_PF;
        int csq[PF > 0 ? PF : 1];
#pragma unroll
        for (int k = 0; k < PF; k++) csq[k] = preloaded + pidx + k * NPROD < myTiles ? chunk_bounds(preloaded + pidx + k * NPROD) : 0;
        for (int it = preloaded + pidx; pidx < NPROD && it < myTiles; it += NPROD) {
            int cs;
            if (PF > 0) {
                cs = csq[0];
#pragma unroll
                for (int k = 0; k + 1 < PF; k++) csq[k] = csq[k + 1];
                csq[PF - 1] = it + PF * NPROD < myTiles ? chunk_bounds(it + PF * NPROD) : 0;
            } else cs = chunk_bounds(it);
            mbar_wait(&empty[it % NS], ((it / NS) - 1) & 1);
            issue(it, 3, cs);
        }
    } else if (CAN_RPL && rpl > 0) {
        // RESIDUE PER LANE (reducing kinds; systems whose residues all have the same number k = 2..8 of particles, every residue in
        // one temperature group: water boxes).  With one particle per lane the energies cost ~115 warp instructions per 32
        // particles — the residue sums are 12 shuffles + 12 adds, the pair partner 3 shuffles, and wait / header / ring / arrive
        // are paid per 32 particles — and the second half without stores and the plain reduction were bound by the issue slots
        // (77 % busy at 60 % of the HBM peak, profiles/ncu_full_r2.md).  Here a lane owns a whole residue: its members are k
        // consecutive particles of the stage, the momentum sum is three adds per member in registers, the pair partner is another
        // read of the stage, nothing is shuffled, and the per-iteration overhead is paid once per 32 k particles.  A tile
        // has at most V2_NCONS * (32 / k) residues = `rplSlots` warps' worth, so the consumer warps form groups of rplSlots that take
        // the tiles in turn (group g: tiles g, g + groups, ...); warps beyond the last whole group have nothing to do.
        // Same formulas as below: sum_i m_i |v_i|^2 - |P|^2 / M - sum_pairs mu |rel|^2 per group, |P|^2 / M for the COM group.
        const int k = rpl;
        const int grp = warp / rplSlots, slot = warp - grp * rplSlots;
        const float fscale = (float)a.fscale;
        const bool store = (KIND == V2_B) && !a.lazyKick;
        if (grp < rplGroups) {
            int stg = grp % NS;
            uint32_t phase = (uint32_t)(grp / NS) & 1u;
            for (int it = grp; it < myTiles; it += rplGroups) {
                unsigned char* st = smem + stg * L::STAGE;
                const float4* sv = reinterpret_cast<const float4*>(st + L::OFF_V);
                const unsigned char* sF = st + L::OFF_F;
                mbar_wait(&full[stg], phase);
                const uint2 hdr = reinterpret_cast<const uint2*>(st + L::OFF_HDR)[V2_NCONS - 1];     // the last chunk ends the tile
                const int start = (int)hdr.x, n = (int)((hdr.y & 0xffffu) + (hdr.y >> 16));
                const int base = (slot * 32 + lane) * k;                  // this lane's residue: particles base .. base + k - 1 of the tile
                if (store) {
                    // the storing second half: the kicked velocities leave in particle order (coalesced 16-byte stores; a lane storing
                    // the members of its own residue would write 16 bytes every 16 k: +9 us at 10 M particles), each formed by the
                    // same fp32 operation on the same operands as in the energy terms below
                    const int w0 = slot * 32 * k + lane;
                    for (int c = 0; c < k; c++) {
                        const int idx = w0 + 32 * c;
                        if (idx < n) {
                            const float4 v4 = sv[idx];
                            st_global(gvelm + (unsigned int)(start + idx), v4.w != 0.f ? pack4(kicked(xyz(v4), fscale * v4.w, v2_force<FFMT>(sF, (start & 3) + idx)), v4.w) : v4);
                        }
                    }
                }
                if (base < n) {
                    const unsigned char* sS = st + L::OFF_S + (start & 15) + base;
                    RplOut o;
                    switch (k) {          // warp-uniform; the member loop is unrolled for each size: every address is base + immediate
                        case 2: o = rpl_residue<2, FFMT, L::HAS_F>(sv + base, sF, (start & 3) + base, sS, stab, fscale, lane); break;
                        case 3: o = rpl_residue<3, FFMT, L::HAS_F>(sv + base, sF, (start & 3) + base, sS, stab, fscale, lane); break;
                        case 4: o = rpl_residue<4, FFMT, L::HAS_F>(sv + base, sF, (start & 3) + base, sS, stab, fscale, lane); break;
                        case 5: o = rpl_residue<5, FFMT, L::HAS_F>(sv + base, sF, (start & 3) + base, sS, stab, fscale, lane); break;
                        case 6: o = rpl_residue<6, FFMT, L::HAS_F>(sv + base, sF, (start & 3) + base, sS, stab, fscale, lane); break;
                        case 7: o = rpl_residue<7, FFMT, L::HAS_F>(sv + base, sF, (start & 3) + base, sS, stab, fscale, lane); break;
                        default: o = rpl_residue<8, FFMT, L::HAS_F>(sv + base, sF, (start & 3) + base, sS, stab, fscale, lane); break;
                    }
                    accDrude += o.keD;
                    accCOM += o.keC;
                    if (o.tg != curTg) {
                        if (curTg >= 0) { ske[curTg * V2_TILE + tid] += accT; accT = 0.f; }
                        curTg = o.tg;
                    }
                    accT += (o.e - o.keC) - o.keD;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stg]);
                stg += rplGroups;
                while (stg >= NS) { stg -= NS; phase ^= 1u; }
            }
        }
    } else {
        F2 eCOM; eCOM.hi = seps[G].x; eCOM.lo = seps[G].y;
        const float dt = (float)a.dt, fscale = (float)a.fscale, rmax = (float)a.rmax, rmax2 = rmax * rmax;
        const int maxRes = USE_COM ? a.maxRes : 1, bfly = a.butterfly;
        int spRow = -1;                                              // species row held in q0, q1
        float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), q1 = q0;
        int stg = 0;
        uint32_t phase = 0;                                          // (stage, parity) advance incrementally: no division / modulo per tile
        for (int it = 0; it < myTiles; it++, phase ^= (stg + 1 == NS), stg = (stg + 1 == NS) ? 0 : stg + 1) {
            unsigned char* st = smem + stg * L::STAGE;
            const float4* sv = reinterpret_cast<const float4*>(st + L::OFF_V);
            const float4* sx = reinterpret_cast<const float4*>(st + L::OFF_X);
            const unsigned char* sF = st + L::OFF_F;
            const unsigned char* sS = st + L::OFF_S;
            mbar_wait(&full[stg], phase);
            const uint2 hdr = reinterpret_cast<const uint2*>(st + L::OFF_HDR)[warp];
            const int start = (int)hdr.x, c0 = (int)(hdr.y & 0xffffu), cn = (int)(hdr.y >> 16);
            const bool active = lane < cn;
            const int i = c0 + lane;                                 // index inside the tile
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            int sp = a.tableRows - 1;                              // the "no particle" row
            V3<float> F = v3(0.f, 0.f, 0.f);
            if (active) {
                v4 = sv[i];
                sp = sS[(start & 15) + i];
                if (L::HAS_F) F = v2_force<FFMT>(sF, (start & 3) + i);
            }
            // The species row of a lane rarely changes from tile to tile (periodic layouts: every chunk of a water box has the same
            // composition), and the shared-memory pipe is what bounds the second half (LDS + SHFL: ~26 of its cycles per warp and
            // tile, 30 warps per SM): the row stays in registers and is fetched again only when the species byte differs.
            if (sp != spRow) {
                q0 = stab[sp * V2_ROW_F4];
                q1 = stab[sp * V2_ROW_F4 + 1];
                spRow = sp;
            }
            const uint32_t meta = __float_as_uint(q0.w);
            const int tg = v2_tg(meta);
            const uint32_t role = v2_role(meta);
            const int pl = lane + v2_partner(meta);                   // lane of the pair partner (this lane for ordinary particles)
            V3<float> v = xyz(v4);
            const float w = v4.w;
            const bool massive = w != 0.f;
            const float fw = fscale * w;
            if (KIND == V2_A && a.lazyKick) v = kicked(v, fw, F);     // the previous step's second half kick (see StreamArgs::lazyKick)
            const int offFirst = v2_off_first(meta), offLast = v2_off_last(meta);    // position inside the residue (all of it in this warp)
            const unsigned int gidx = (unsigned int)(start + i);                      // unsigned: one IMAD.WIDE per address

            if (IS_A) {
                // thermostat scaling (integrateDrudeTGNHChain, drudeTGNH.cu:255-300) in the unified form of tgnh_kernels.cuh,
                // half kick (:314-364), drift (:438-465), hard wall (:474-573)
                V3<float> V = v3(0.f, 0.f, 0.f);
                if (USE_COM) {
                    // V only feeds the corrections (sT-1)(v-V), (sCOM-1)V: fp32 masses are ample
                    V = q0.z * residue_sum(q0.x * v, lane, offFirst, offLast, maxRes, bfly);
                }
                const V3<float> vj = v3(__shfl_sync(0xffffffffu, v.x, pl), __shfl_sync(0xffffffffu, v.y, pl), __shfl_sync(0xffffffffu, v.z, pl));
                const V3<float> rel = vj - v;
                const float2 e2 = seps[tg];
                F2 eT; eT.hi = e2.x; eT.lo = e2.y;
                V3<float> vn = scaled_velocity2(v, eT, v - V, eCOM, V, scoef[tg] * q1.w, rel);
                if (KIND == V2_S) {
                    if (active) st_global(gvelm + gidx, massive ? pack4(vn, w) : v4);      // (massless: see the first half's stores)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[stg]);
                    continue;
                }
                vn = kicked(vn, fw, F);
                float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (active) x4 = sx[i];
                const V3<float> x = xyz(x4);
                V3<float> xn = axpy(dt, vn, x);
                if (HARDWALL) {
                    // the partner's new velocity and old position, as its own lane computed them: both lanes of a pair see
                    // bit-identical operands and take the same side of the wall test
                    V3<float> vjn = v3(__shfl_sync(0xffffffffu, vn.x, pl), __shfl_sync(0xffffffffu, vn.y, pl), __shfl_sync(0xffffffffu, vn.z, pl));
                    const V3<float> xj = v3(__shfl_sync(0xffffffffu, x.x, pl), __shfl_sync(0xffffffffu, x.y, pl), __shfl_sync(0xffffffffu, x.z, pl));
                    const float wj = __shfl_sync(0xffffffffu, w, pl);
                    if (role != ROLE_NORMAL) {
                        // displacement from the exact difference of the old positions plus the relative drift (no cancellation
                        // of two rounded box-sized coordinates); the parent forms -(that) with the operands swapped, bit for bit
                        const V3<float> dDP = role == ROLE_DRUDE ? axpy(dt, vn - vjn, x - xj) : axpy(dt, vjn - vn, xj - x);   // Drude minus parent
                        const float r2 = dot3(dDP);
                        if (r2 > rmax2) {                             // rInv*maxDrudeDistance < 1  (drudeTGNH.cu:490)
                            V3<float> xjn = axpy(dt, vjn, xj);
                            if (role == ROLE_DRUDE) hard_wall(dDP, r2, xn, xjn, vn, vjn, w, wj, rmax, (float)a.hardwallScale, dt);
                            else hard_wall(dDP, r2, xjn, xn, vjn, vn, wj, w, rmax, (float)a.hardwallScale, dt);
                        }
                    }
                }
                if (active) {
                    // A massless particle (virtual site: w = 0) is not integrated (:258, :318, :441); its slots get the bits they held
                    // back instead of being skipped, so that every 32-byte sector the warp touches is written whole (a box of 5-site
                    // SWM4 waters skips every fifth slot otherwise and its first half took 182 us per 10 M particles, against 114 us
                    // for the 4-site box)
                    const float4 vout = massive ? pack4(vn, w) : v4;
#if TGNH_V2_A_STORE_KEEP_MB > 0
                    // experiment: only the velm this launch writes last can still be in L2 when the next launch reads it
                    st_global_hint(gvelm + gidx, vout, it >= storeKeepFrom ? polStoreKeep : polOnce);
#else
                    st_global(gvelm + gidx, vout);
#endif
                    st_stream(static_cast<float4*>(a.posq) + gidx, massive ? make_float4(xn.x, xn.y, xn.z, x4.w) : x4);
                }
            } else {
                // half kick (integrateDrudeTGNHVelocities, :314-364; V2_KE: F = 0, nothing stored), then the energies of
                // what was stored
                const V3<float> vn = kicked(v, fw, F);
                if (KIND == V2_B && !a.lazyKick && active) st_global(gvelm + gidx, massive ? pack4(vn, w) : v4);
#ifndef TGNH_V2_ABLATE
#define TGNH_V2_ABLATE 0
#endif
#if TGNH_V2_ABLATE == 2            // measurement builds only (wrong energies): what the pipeline costs without the energy arithmetic
                accT += vn.x + vn.y + vn.z;
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[stg]);
                continue;
#endif
                float keC = 0.f, keRel;
                // computeNormalizedKineticEnergies (:161-186); no branches on the particle: a massless particle, a lane without
                // particle and a particle that is no Drude particle have m = 0 resp. contribute a term multiplied by 0
                if (USE_COM && TGNH_V2_ABLATE != 1) {
                    // calcCOMVelocities (:86-105): P = sum m v over the residue, V = P / M
                    const V3<float> pm = v3(mul2(q0.x, q0.y, vn.x), mul2(q0.x, q0.y, vn.y), mul2(q0.x, q0.y, vn.z));
                    const V3<float> P = residue_sum(pm, lane, offFirst, offLast, maxRes, bfly);
                    keC = offFirst == 0 ? mul2(q0.z, q1.z, dot3(P)) : 0.f;      // |P|^2 / M  (:154), carried by the residue's first particle
                    if (a.uniformGroups) {
                        // every residue lies in one temperature group: sum_i m_i |v_i - V|^2 = sum_i m_i |v_i|^2 - |P|^2 / M, the
                        // members need neither V nor their relative velocity
                        keRel = pm.x * vn.x + pm.y * vn.y + pm.z * vn.z - keC;
                    } else {
                        const V3<float> r = vn - q0.z * P;             // normalizeVelocities (:126-128)
                        keRel = mul2(q0.x, q0.y, dot3(r));
                    }
                } else
                    keRel = mul2(q0.x, q0.y, dot3(vn));
#if TGNH_V2_ABLATE == 1            // measurement build: no residue sums, no pair term
                const float keD = 0.f;
#else
                const V3<float> vjn = v3(__shfl_sync(0xffffffffu, vn.x, pl), __shfl_sync(0xffffffffu, vn.y, pl), __shfl_sync(0xffffffffu, vn.z, pl));
                const float keD = role == ROLE_DRUDE ? mul2(q1.x, q1.y, dot3(vjn - vn)) : 0.f;   // mu |rel|^2 (:185)
#endif
                // m_d |r_d|^2 + m_p |r_p|^2 - mu |rel|^2 = (m_d + m_p) |cm|^2 (:184; the identity holds in any frame)
                const float ke = keRel - keD;
                accDrude += keD;
                accCOM += keC;
                // this thread's particles usually stay in one group from tile to tile (molecule-periodic group patterns): the running
                // sum lives in a register and moves to its shared-memory column only when the group changes
                if (tg != curTg) {
                    if (curTg >= 0) { ske[curTg * V2_TILE + tid] += accT; accT = 0.f; }
                    curTg = tg;
                }
                accT += ke;
            }
            // hand the stage back: one arrival per consumer warp
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[stg]);
        }
    }
    pdl_launch_dependents();
    if (KIND == V2_S && blockIdx.x == 0 && tid < T) {
        // the factors are applied: the kinetic energies of the stored velocities are s_g^2 KE_g (residue-uniform groups)
        const double sg = a.chain.scaleA[tid];
        a.chain.ke2[tid] *= sg * sg;
        a.chain.pending[tid] = 1.0;
    }
    if (!L::HAS_KE) return false;

    // ---- deterministic reduction: thread columns -> warp -> CTA -> (last CTA) grid ----
    if (!producer) {
        if (curTg >= 0) ske[curTg * V2_TILE + tid] += accT;
        for (int g = 0; g < T; g++) {
            double x = g == G ? (double)accCOM : g == G + 1 ? (double)accDrude : (double)ske[g * V2_TILE + tid];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) swarp[g * 32 + warp] = x;
        }
    }
    __syncthreads();
    if (tid < T) {
        double x = 0.0;
        for (int w = 0; w < V2_NCONS; w++) x += swarp[tid * 32 + w];
        a.partials[(size_t)blockIdx.x * T + tid] = x;
        __threadfence();                               // only the writers: a fence in all 512 threads costs a microsecond per CTA
    }
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
    for (int g = warp; g < T; g += V2_NCONS + V2_NPROD) {
        double x = 0.0;
        for (int b = lane; b < (int)gridDim.x; b += 32) x += __ldcg(a.partials + (size_t)b * T + g);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
        if (lane == 0) out[g] = a.accumulate ? out[g] + x : x;
    }
    if (a.peers.world > 1) {
        // sharded: hand this rank's sums to every rank (this one included) over NVLink
        __syncthreads();
        peer_publish(a.peers, out, T, tid, (int)blockDim.x);
    }
    if (tid == 0) *a.ticket = 0u;
    return true;
}

template <int KIND, int FFMT, bool USE_COM, bool HARDWALL>
__global__ void __launch_bounds__(V2_THREADS, (KIND == V2_B || KIND == V2_KE) ? V2_CTAS_RED : V2_CTAS) tgnh_v2_kernel(const __grid_constant__ StreamArgs a) {
    v2_body<KIND, FFMT, USE_COM, HARDWALL>(a);
}


// Small systems (every CTA owns at most one tile): the CTA that finishes the energy reduction runs the Nose-Hoover chain
// update itself instead of a separate chain launch (see tgnh_stream_chain_kernel).
template <int KIND, int FFMT, bool USE_COM>
__global__ void __launch_bounds__(V2_THREADS, 1) tgnh_v2_chain_kernel(const __grid_constant__ StreamArgs a) {
    if (!v2_body<KIND, FFMT, USE_COM, false, false>(a)) return;
    __syncthreads();                                   // the energy vector written by this CTA's warps
    if (threadIdx.x < 32) chain_phase(a.chain, a.fusedChainMode, threadIdx.x);
}

}  // namespace tgnh
