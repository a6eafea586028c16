// libtgnh.so — C-ABI (include/tgnh.h) over the sm_100a TGNH kernels.
//
// Host half of the hot path: what CudaIntegrateDrudeTGNHStepKernel::initialize / execute /
// propagateNHChain do in the reference (platforms/cuda/src/CudaDrudeTGNHKernels.cpp:75-282, 284-408,
// 433-652), re-designed so that a step is two streaming launches plus the Nose-Hoover chain, resident
// on the device: no D2H/H2D on the step path, no runtime kernel compilation.
#include "../../include/tgnh.h"

#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tgnh_kernels.cuh"
#include "tgnh_v2.cuh"

#include <array>
#include <map>

using namespace tgnh;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_error;

static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess) return fail(TGNH_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

#define TGNH_STR2(x) #x
#define TGNH_STR(x) TGNH_STR2(x)
extern "C" const char* tgnh_last_error(void) { return g_error.c_str(); }
extern "C" const char* tgnh_build_info(void) { return "libtgnh sm_100a AOT (CUDA " TGNH_STR(CUDART_VERSION) "), TMA bulk pipeline, device-resident NH chain"; }

// ------------------------------------------------------------------------------------------------
// NCCL, bound at run time so that a process that already loaded a libnccl.so.2 (e.g. torch) shares it
// ------------------------------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load() {
    if (g_nccl.lib) return TGNH_OK;
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) return fail(TGNH_ERR_NCCL, "cannot load libnccl.so.2: %s", dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(lib, "ncclCommDestroy");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(lib, "ncclAllReduce");
    g_nccl.AllGather = (decltype(g_nccl.AllGather))dlsym(lib, "ncclAllGather");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllReduce || !g_nccl.AllGather || !g_nccl.GetErrorString)
        return fail(TGNH_ERR_NCCL, "libnccl.so.2 lacks a required symbol");
    g_nccl.lib = lib;
    return TGNH_OK;
}

#define NCCL_TRY(expr)                                                                               \
    do {                                                                                             \
        ncclResult_t r_ = (expr);                                                                    \
        if (r_ != ncclSuccess) return fail(TGNH_ERR_NCCL, "%s failed: %s", #expr, g_nccl.GetErrorString(r_)); \
    } while (0)

struct tgnh_comm {
    ncclComm_t comm = nullptr;
    int worldSize = 1, rank = 0, device = 0;
};

extern "C" int tgnh_comm_get_unique_id(void* id_out) {
    if (!id_out) return fail(TGNH_ERR_INVALID_ARGUMENT, "id_out is null");
    if (int rc = nccl_load()) return rc;
    static_assert(sizeof(ncclUniqueId) == TGNH_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return TGNH_OK;
}

extern "C" int tgnh_comm_create(const void* unique_id, int world_size, int rank, int device, tgnh_comm** out) {
    if (!unique_id || !out || world_size < 1 || rank < 0 || rank >= world_size)
        return fail(TGNH_ERR_INVALID_ARGUMENT, "bad communicator arguments");
    if (int rc = nccl_load()) return rc;
    if (device >= 0) CUDA_TRY(cudaSetDevice(device));
    else CUDA_TRY(cudaGetDevice(&device));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof id);
    tgnh_comm* c = new tgnh_comm();
    c->worldSize = world_size; c->rank = rank; c->device = device;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world_size, id, rank);
    if (r != ncclSuccess) { delete c; return fail(TGNH_ERR_NCCL, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); }
    *out = c;
    return TGNH_OK;
}

extern "C" void tgnh_comm_destroy(tgnh_comm* c) {
    if (!c) return;
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

// ------------------------------------------------------------------------------------------------
// handle
// ------------------------------------------------------------------------------------------------
struct tgnh_handle {
    int device = 0, numSMs = 0;
    int N = 0, paddedN = 0, P = 0, R = 0, G = 0, T = 0, M = 0, S = 0;
    int useDrudeNH = 0, useCOM = 0, ffmt = 0, hardwall = 0, prec = 0;
    void* posqCorrection = nullptr;   // mixed precision: cu.getPosqCorrection(), registered with tgnh_set_posq_correction
    bool uniformGroups = true;   // every residue lies in one temperature group (folding / KE carry-over legal)
    double dt = 0, rmax = 0, kT = 0, kTD = 0;
    int numTiles = 0;
    // device
    uint32_t* dDesc = nullptr;
    int* dTileStart = nullptr;
    int* dResStart = nullptr;     // first particle of every residue in particle order, then N (padded to a multiple of 4)
    int* dTileFirstRes = nullptr; // index into dResStart of each tile's first residue
    int numBig = 0;               // residues of more than MAX_RES particles
    int *dBigFirst = nullptr, *dBigLast = nullptr;
    double4* dBigCom = nullptr;   // {V, M} of the big residues (tgnh_bigcom_kernel)
    int kindB = KIND_B;           // KIND_BU when every residue lies in one temperature group
    double* dChain = nullptr;     // one allocation holding every ChainView array
    double* dPartials = nullptr;
    unsigned int* dTicket = nullptr;
    ChainView chain{};
    size_t chainDoubles = 0;
    int gridA = 0, gridB = 0, gridKE = 0, gridA1 = 0, gridA2 = 0, gridS = 0, gridK = 0;
    int smemA = 0, smemB = 0, smemKE = 0, smemA1 = 0, smemA2 = 0, smemS = 0, smemK = 0;
    int kindKE = KIND_KE;         // KIND_KU when every residue lies in one temperature group
    bool fuseChain = false;       // small system: reducing launches run the chain update in their last CTA
    // warp-chunk kernels (tgnh_v2.cuh): single-precision layout, every residue fits a warp, at most 255 species
    bool v2 = false;
    unsigned char* dSpec = nullptr;
    int* dChunkStart = nullptr;
    float4* dSpecTable = nullptr;
    int numTiles2 = 0, maxRes = 1, numSpecies = 0, butterfly = 0;
    int rpl = 0, rplMode = 2;     // residue-per-lane reduction: residue size (0 = off); TGNH_RPL = 0 off, 1 only the launches that store nothing, 2 all reducing launches
    int gridA2v = 0, gridB2v = 0, gridKE2v = 0, gridS2v = 0, smemA2v = 0, smemB2v = 0, smemKE2v = 0, smemS2v = 0;
    bool lazyKick = true;         // TGNH_LAZY_KICK=0 at tgnh_create switches the lazy second kick of tgnh_step off (tests, measurements)
    bool lazyNow = false;         // set by tgnh_step around the launches that leave / find the second half kick pending (StreamArgs::lazyKick)
    bool earlyOK = false;         // set by tgnh_step around launches whose predecessors in the stream are its own
    std::vector<int> hChunkStart; // host copy of dChunkStart (chunk boundaries of the pipelined host-buffer path)
    // pipelined host-buffer path (tgnh_step_host2)
    cudaStream_t hsIn = nullptr, hsOut = nullptr;
    std::vector<cudaEvent_t> hsEvents;
    void* hsCorr = nullptr;
    const void* hsForceHost = nullptr;   // host pointer whose contents hsForce holds
    // host copies of the thermostat parameters
    std::vector<double> dof, nkbt, etaMass;
    // state machine
    bool keValid = false;         // chain.ke2 describes the velocities as stored
    bool scalePending = false;    // chain.pending != 1 has not been applied to velm yet
    int64_t launches = 0;
    int nextReverse = 0;          // direction of the next streaming launch (alternates, see launch_stream)
    tgnh_comm* comm = nullptr;
    // sharded: kinetic-energy exchange through peer-mapped inboxes (NVLink); world <= 1 when NCCL carries it instead
    PeerInbox* dInbox = nullptr;
    PeerView peers{};
    unsigned int reduceSeq = 0;
    // optional per-launch device timing (bench.py's roofline leg)
    bool profiling = false;
    std::vector<cudaEvent_t> evPool;
    std::vector<int> evKind;      // kind of the launch bracketed by evPool[2i], evPool[2i+1]
    size_t evUsed = 0;
    // staging buffers of tgnh_step_host
    void *hsVelm = nullptr, *hsPosq = nullptr, *hsForce = nullptr;
    cudaStream_t hsStream = nullptr;
};

typedef void (*StreamKernel)(const StreamArgs);

// Map every rank's inbox into this process (CUDA IPC over NVLink / NVSwitch).  Collective over the communicator; all
// ranks end up with the same answer: either everybody uses the inboxes or (IPC unavailable, different nodes, more than
// MAX_PEERS ranks, TGNH_P2P=0) everybody keeps the NCCL all-reduce.
static int setup_peers(tgnh_handle* h) {
    const int W = h->comm->worldSize, R = h->comm->rank;
    h->peers = PeerView{};
    const char* env = getenv("TGNH_P2P");
    bool ok = W <= MAX_PEERS && !(env && atoi(env) == 0);
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof mine);
    if (ok) ok = cudaMalloc((void**)&h->dInbox, sizeof(PeerInbox)) == cudaSuccess && cudaMemset(h->dInbox, 0, sizeof(PeerInbox)) == cudaSuccess;
    if (ok) ok = cudaIpcGetMemHandle(&mine, h->dInbox) == cudaSuccess;
    (void)cudaGetLastError();
    // all-gather the handles (and whether each rank got this far)
    const size_t rec = sizeof(cudaIpcMemHandle_t) + 8;
    std::vector<unsigned char> all((size_t)W * rec, 0);
    memcpy(&all[R * rec], &mine, sizeof mine);
    all[R * rec + sizeof mine] = ok ? 1 : 0;
    unsigned char* dall = nullptr;
    if (cudaMalloc((void**)&dall, all.size()) != cudaSuccess) return fail(TGNH_ERR_CUDA, "cudaMalloc failed");
    cudaMemcpy(dall, all.data(), all.size(), cudaMemcpyHostToDevice);
    ncclResult_t r = g_nccl.AllGather(dall + R * rec, dall, rec, ncclChar, h->comm->comm, 0);
    cudaError_t e = cudaStreamSynchronize(0);
    cudaMemcpy(all.data(), dall, all.size(), cudaMemcpyDeviceToHost);
    if (r != ncclSuccess || e != cudaSuccess) { cudaFree(dall); return fail(TGNH_ERR_NCCL, "all-gather of the inbox handles failed"); }
    for (int q = 0; q < W; q++) ok = ok && all[q * rec + sizeof mine] == 1;
    if (ok) {
        for (int q = 0; q < W && ok; q++) {
            if (q == R) { h->peers.inbox[q] = h->dInbox; continue; }
            cudaIpcMemHandle_t hd;
            memcpy(&hd, &all[q * rec], sizeof hd);
            void* ptr = nullptr;
            ok = cudaIpcOpenMemHandle(&ptr, hd, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
            h->peers.inbox[q] = ok ? (PeerInbox*)ptr : nullptr;
        }
        (void)cudaGetLastError();
    }
    // agreement: one rank failing to map a peer sends everybody back to NCCL
    double flag = ok ? 1.0 : 0.0;
    cudaMemcpy(dall, &flag, 8, cudaMemcpyHostToDevice);
    r = g_nccl.AllReduce(dall, dall, 1, ncclDouble, ncclMin, h->comm->comm, 0);
    e = cudaStreamSynchronize(0);
    cudaMemcpy(&flag, dall, 8, cudaMemcpyDeviceToHost);
    cudaFree(dall);
    if (r != ncclSuccess || e != cudaSuccess) return fail(TGNH_ERR_NCCL, "all-reduce of the inbox status failed");
    if (flag < 1.0) {
        for (int q = 0; q < MAX_PEERS; q++) {
            if (h->peers.inbox[q] && h->peers.inbox[q] != h->dInbox) cudaIpcCloseMemHandle(h->peers.inbox[q]);
            h->peers.inbox[q] = nullptr;
        }
        cudaFree(h->dInbox);
        h->dInbox = nullptr;
        h->peers = PeerView{};
        return TGNH_OK;
    }
    h->peers.world = W;
    h->peers.rank = R;
    return TGNH_OK;
}

template <int KIND, int FFMT, int PREC, bool BIG>
static StreamKernel pick3(bool useCOM, bool hardwall) {
    if constexpr (KIND == KIND_A) {   // only the kernels that move positions contain the hard wall
        if (useCOM) return hardwall ? tgnh_stream_kernel<KIND_A, FFMT, true, true, PREC, BIG> : tgnh_stream_kernel<KIND_A, FFMT, true, false, PREC, BIG>;
        return hardwall ? tgnh_stream_kernel<KIND_A, FFMT, false, true, PREC, false> : tgnh_stream_kernel<KIND_A, FFMT, false, false, PREC, false>;
    } else if constexpr (KIND == KIND_A2) {
        return hardwall ? tgnh_stream_kernel<KIND_A2, 0, false, true, PREC, false> : tgnh_stream_kernel<KIND_A2, 0, false, false, PREC, false>;
    } else if constexpr (KIND == KIND_KE || KIND == KIND_S || KIND == KIND_KU) {      // no forces: one force format
        return useCOM ? tgnh_stream_kernel<KIND, 0, true, false, PREC, BIG> : tgnh_stream_kernel<KIND, 0, false, false, PREC, false>;
    } else if constexpr (KIND == KIND_K) {
        return tgnh_stream_kernel<KIND_K, FFMT, false, false, PREC, false>;
    } else {
        return useCOM ? tgnh_stream_kernel<KIND, FFMT, true, false, PREC, BIG> : tgnh_stream_kernel<KIND, FFMT, false, false, PREC, false>;
    }
}

// residues larger than a tile only matter to kernels that use the residues' COM velocity (USE_COM)
template <int KIND, int FFMT, int PREC>
static StreamKernel pick2(bool useCOM, bool hardwall, bool big) {
    return big ? pick3<KIND, FFMT, PREC, true>(useCOM, hardwall) : pick3<KIND, FFMT, PREC, false>(useCOM, hardwall);
}

// PREC 2 (double4 posq) exists only for the kernels that touch positions; elsewhere double and mixed are the same kernel
template <int KIND>
static StreamKernel pick1(int ffmt, int prec, bool useCOM, bool hardwall, bool big) {
    if (prec == 2 && (KIND == KIND_A || KIND == KIND_A2))
        return ffmt ? pick2<KIND, 1, (KIND == KIND_A || KIND == KIND_A2) ? 2 : 1>(useCOM, hardwall, big)
                    : pick2<KIND, 0, (KIND == KIND_A || KIND == KIND_A2) ? 2 : 1>(useCOM, hardwall, big);
    if (prec) return ffmt ? pick2<KIND, 1, 1>(useCOM, hardwall, big) : pick2<KIND, 0, 1>(useCOM, hardwall, big);
    return ffmt ? pick2<KIND, 1, 0>(useCOM, hardwall, big) : pick2<KIND, 0, 0>(useCOM, hardwall, big);
}

static StreamKernel pick(int kind, int ffmt, int prec, bool useCOM, bool hardwall, bool big) {
    switch (kind) {
        case KIND_A: return pick1<KIND_A>(ffmt, prec, useCOM, hardwall, big);
        case KIND_B: return pick1<KIND_B>(ffmt, prec, useCOM, hardwall, big);
        case KIND_BU: return pick1<KIND_BU>(ffmt, prec, useCOM, hardwall, big);
        case KIND_A1: return pick1<KIND_A1>(ffmt, prec, useCOM, hardwall, big);
        case KIND_A2: return pick1<KIND_A2>(ffmt, prec, useCOM, hardwall, big);
        case KIND_S: return pick1<KIND_S>(ffmt, prec, useCOM, hardwall, big);
        case KIND_K: return pick1<KIND_K>(ffmt, prec, useCOM, hardwall, big);
        case KIND_KU: return pick1<KIND_KU>(ffmt, prec, useCOM, hardwall, big);
        default: return pick1<KIND_KE>(ffmt, prec, useCOM, hardwall, big);
    }
}

// small systems: the reducing launch runs the chain update in its last CTA (tgnh_stream_chain_kernel)
template <int KIND, int FFMT>
static StreamKernel pick_fused2(int prec, bool useCOM) {
    if (prec) return useCOM ? tgnh_stream_chain_kernel<KIND, FFMT, true, 1> : tgnh_stream_chain_kernel<KIND, FFMT, false, 1>;
    return useCOM ? tgnh_stream_chain_kernel<KIND, FFMT, true, 0> : tgnh_stream_chain_kernel<KIND, FFMT, false, 0>;
}
static StreamKernel pick_fused(int kind, int ffmt, int prec, bool useCOM) {
    switch (kind) {
        case KIND_B: return ffmt ? pick_fused2<KIND_B, 1>(prec, useCOM) : pick_fused2<KIND_B, 0>(prec, useCOM);
        case KIND_BU: return ffmt ? pick_fused2<KIND_BU, 1>(prec, useCOM) : pick_fused2<KIND_BU, 0>(prec, useCOM);
        case KIND_KU: return pick_fused2<KIND_KU, 0>(prec, useCOM);
        default: return pick_fused2<KIND_KE, 0>(prec, useCOM);
    }
}

template <int KIND, int FFMT, int PREC>
static int smem2(bool useCOM, int T) {
    if (KIND == KIND_A2) return SmemLayout<KIND_A2, 0, false, PREC>::bytes(T);
    if (KIND == KIND_KE) return useCOM ? SmemLayout<KIND_KE, 0, true, PREC>::bytes(T) : SmemLayout<KIND_KE, 0, false, PREC>::bytes(T);
    if (KIND == KIND_S) return useCOM ? SmemLayout<KIND_S, 0, true, PREC>::bytes(T) : SmemLayout<KIND_S, 0, false, PREC>::bytes(T);
    if (KIND == KIND_KU) return useCOM ? SmemLayout<KIND_KU, 0, true, PREC>::bytes(T) : SmemLayout<KIND_KU, 0, false, PREC>::bytes(T);
    if (KIND == KIND_K) return SmemLayout<KIND_K, FFMT, false, PREC>::bytes(T);
    return useCOM ? SmemLayout<KIND, FFMT, true, PREC>::bytes(T) : SmemLayout<KIND, FFMT, false, PREC>::bytes(T);
}

template <int KIND>
static int smem1(int ffmt, int prec, bool useCOM, int T) {
    if (prec == 2 && (KIND == KIND_A || KIND == KIND_A2))
        return ffmt ? smem2<KIND, 1, (KIND == KIND_A || KIND == KIND_A2) ? 2 : 1>(useCOM, T) : smem2<KIND, 0, (KIND == KIND_A || KIND == KIND_A2) ? 2 : 1>(useCOM, T);
    if (prec) return ffmt ? smem2<KIND, 1, 1>(useCOM, T) : smem2<KIND, 0, 1>(useCOM, T);
    return ffmt ? smem2<KIND, 1, 0>(useCOM, T) : smem2<KIND, 0, 0>(useCOM, T);
}

static int smem_bytes(int kind, int ffmt, int prec, bool useCOM, int T) {
    switch (kind) {
        case KIND_A: return smem1<KIND_A>(ffmt, prec, useCOM, T);
        case KIND_B: return smem1<KIND_B>(ffmt, prec, useCOM, T);
        case KIND_BU: return smem1<KIND_BU>(ffmt, prec, useCOM, T);
        case KIND_A1: return smem1<KIND_A1>(ffmt, prec, useCOM, T);
        case KIND_A2: return smem1<KIND_A2>(ffmt, prec, useCOM, T);
        case KIND_S: return smem1<KIND_S>(ffmt, prec, useCOM, T);
        case KIND_K: return smem1<KIND_K>(ffmt, prec, useCOM, T);
        case KIND_KU: return smem1<KIND_KU>(ffmt, prec, useCOM, T);
        default: return smem1<KIND_KE>(ffmt, prec, useCOM, T);
    }
}

static int configure_kernel(tgnh_handle* h, int kind, int* grid, int* smem) {
    StreamKernel k = pick(kind, h->ffmt, h->prec, h->useCOM, h->hardwall, h->numBig > 0);
    *smem = smem_bytes(kind, h->ffmt, h->prec, h->useCOM, h->T);
    if (*smem > 227 * 1024)
        return fail(TGNH_ERR_UNSUPPORTED, "%d temperature groups need %d bytes of shared memory per CTA (limit 232448)", h->G, *smem);
    CUDA_TRY(cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, *smem));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k, TILE, *smem));
    if (occ < 1) return fail(TGNH_ERR_CUDA, "kernel kind %d cannot be resident (smem %d B)", kind, *smem);
    int g = occ * h->numSMs;
    if (g > h->numTiles) g = h->numTiles;
    if (g < 1) g = 1;
    *grid = g;
    return TGNH_OK;
}

// ------------------------------------------------------------------------------------------------
// create: index tables, DOF bookkeeping, thermostat masses (CudaDrudeTGNHKernels.cpp:75-235)
// ------------------------------------------------------------------------------------------------
// scalar arguments of tgnh_params (no table is read, no device is touched)
static int validate_params(const tgnh_params* p) {
    const int N = p->num_particles, P = p->num_pairs, R = p->num_residues, G = p->num_temp_groups, M = p->num_nh_chains;
    const int T = G + 2;
    if (N < 1) return fail(TGNH_ERR_INVALID_ARGUMENT, "num_particles must be positive");
    if (P < 0 || R < 1 || G < 1) return fail(TGNH_ERR_INVALID_ARGUMENT, "num_pairs/num_residues/num_temp_groups out of range");
    if (!p->masses || !p->particle_temp_group || !p->particle_res_id || (P > 0 && (!p->pair_drude || !p->pair_parent)))
        return fail(TGNH_ERR_INVALID_ARGUMENT, "a required host table is null");
    if (p->num_constraints < 0 || (p->num_constraints > 0 && (!p->constraint_p || !p->constraint_p1)))
        return fail(TGNH_ERR_INVALID_ARGUMENT, "constraint tables missing");
    if (T > MAX_T) return fail(TGNH_ERR_UNSUPPORTED, "at most %d temperature groups are supported (got %d)", MAX_T - 2, G);
    if (M < 1 || M > MAX_M) return fail(TGNH_ERR_UNSUPPORTED, "numNHChains must be in [1,%d] (got %d)", MAX_M, M);
    if (p->drude_steps_per_real_step < 1) return fail(TGNH_ERR_INVALID_ARGUMENT, "drudeStepsPerRealStep must be >= 1");
    if (p->padded_num_particles < N || (p->padded_num_particles & 3))
        return fail(TGNH_ERR_INVALID_ARGUMENT, "padded_num_particles must be a multiple of 4 and >= num_particles");
    if (p->force_format != TGNH_FORCE_F32_SOA && p->force_format != TGNH_FORCE_I64_SOA)
        return fail(TGNH_ERR_INVALID_ARGUMENT, "unknown force_format %d", p->force_format);
    if (p->precision != TGNH_PRECISION_SINGLE && p->precision != TGNH_PRECISION_MIXED && p->precision != TGNH_PRECISION_DOUBLE)
        return fail(TGNH_ERR_INVALID_ARGUMENT, "unknown precision %d", p->precision);
    if (p->max_drude_distance < 0) return fail(TGNH_ERR_INVALID_ARGUMENT, "setMaxDrudeDistance: Distance cannot be negative");
    if (!(p->step_size > 0)) return fail(TGNH_ERR_INVALID_ARGUMENT, "step_size must be positive");
    return TGNH_OK;
}

// Everything tgnh_create derives on the host from the System's tables: residue ranges, DOF bookkeeping, pair partners,
// per-particle descriptors, residue-aligned tiles (big residues cut where no Drude pair is separated), the per-tile residue
// lists.  No device is touched, so it is also what tgnh_plan_tiles exposes to tests that run without a GPU.
struct HostPlan {
    std::vector<int> resFirst, resLast;
    std::vector<double> resMass, dofv, redMass;
    std::vector<int> partner;
    std::vector<uint32_t> role, desc;
    std::vector<int> tileStart, resStart, tileFirstRes, bigFirst, bigLast;
    double drudeDof = 0, comDof = 0;
    bool uniform = true;
    // warp-chunk plan (tgnh_v2.cuh); v2 == false when the system does not qualify (v2why says why)
    bool v2 = false;
    const char* v2why = "";
    std::vector<int> chunkStart;
    std::vector<unsigned char> spec;
    std::vector<float> specTable;
    int maxRes = 1, numTiles2 = 0, numSpecies = 0, butterfly = 0;
    int uniformRes = 0;           // every residue has this many particles (2..8), else 0
};

// Chunks, species bytes and the species table of the warp-chunk kernels.  Needs the legacy plan's resFirst/resLast/partner/role.
static void build_plan_v2(const tgnh_params* p, HostPlan& hp) {
    const int N = p->num_particles, R = p->num_residues;
    hp.v2 = false;
    if (p->precision != TGNH_PRECISION_SINGLE) { hp.v2why = "mixed / double precision layout"; return; }
    int maxRes = 0;
    for (int r = 0; r < R; r++) maxRes = std::max(maxRes, hp.resLast[r] - hp.resFirst[r] + 1);
    if (maxRes > 32) { hp.v2why = "a residue has more than 32 particles"; return; }
    hp.maxRes = maxRes;
    // residue-aligned chunks of at most 32 particles
    hp.chunkStart.assign(1, 0);
    for (int i = 0, cur = 0; i < N;) {
        const int r = p->particle_res_id[i], len = hp.resLast[r] - hp.resFirst[r] + 1;
        if (cur + len > 32) { hp.chunkStart.push_back(i); cur = 0; }
        cur += len;
        i += len;
    }
    const int numChunks = (int)hp.chunkStart.size();
    hp.numTiles2 = (numChunks + V2_NCONS - 1) / V2_NCONS;
    hp.chunkStart.resize((size_t)V2_NCONS * hp.numTiles2 + 1, N);
    // species: particles that agree in everything the kernels look up per particle
    std::map<std::array<uint64_t, 4>, int> rows;
    hp.spec.assign((size_t)((N + 15) & ~15) + 32, (unsigned char)0);
    hp.specTable.assign((size_t)(V2_MAX_SPECIES + 1) * V2_ROW_F4 * 4, 0.f);
    auto bits = [](double x) { uint64_t b; memcpy(&b, &x, 8); return b; };
    auto put_row = [&](int row, double m, double mu, double invM, double fpartner, uint32_t meta) {
        float* q = &hp.specTable[(size_t)row * V2_ROW_F4 * 4];
        const float mh = (float)m, muh = (float)mu, ih = (float)invM;
        float metaf;
        memcpy(&metaf, &meta, 4);
        q[0] = mh; q[1] = (float)(m - (double)mh); q[2] = ih; q[3] = metaf;
        q[4] = muh; q[5] = (float)(mu - (double)muh); q[6] = (float)(invM - (double)ih); q[7] = (float)fpartner;
    };
    // a tiny direct-mapped cache in front of the map: consecutive molecules repeat the same few species (200 M particles in C5)
    struct Slot { std::array<uint64_t, 4> key; int row; };
    std::vector<Slot> cache(256, Slot{{~0ull, ~0ull, ~0ull, ~0ull}, -1});
    for (int i = 0; i < N; i++) {
        const int r = p->particle_res_id[i];
        const double m = p->masses[i], mj = hp.partner[i] ? p->masses[i + hp.partner[i]] : 0.0;
        const double M = hp.resMass[r];                                   // summed in particle order, as calcCOMVelocities does (:90-100)
        const uint32_t meta = v2_meta_pack(p->particle_temp_group[i], hp.role[i], hp.partner[i], i - hp.resFirst[r], hp.resLast[r] - i);
        const std::array<uint64_t, 4> key = {bits(m), bits(mj), bits(M), (uint64_t)meta};
        Slot& slot = cache[(size_t)((key[0] * 0x9E3779B97F4A7C15ull ^ key[1] * 0xC2B2AE3D27D4EB4Full ^ key[2] ^ key[3] * 0x165667B19E3779F9ull) >> 56)];
        if (slot.row < 0 || slot.key != key) {
            auto it = rows.find(key);
            if (it == rows.end()) {
                if ((int)rows.size() >= V2_MAX_SPECIES) { hp.v2why = "more than 255 particle species"; return; }
                const int row = (int)rows.size();
                it = rows.emplace(key, row).first;
                const bool pair = hp.partner[i] != 0;
                put_row(row, m, pair ? m * mj / (m + mj) : 0.0, M > 0.0 ? 1.0 / M : 0.0, pair ? mj / (m + mj) : 0.0, meta);
            }
            slot.key = key;
            slot.row = it->second;
        }
        hp.spec[i] = (unsigned char)slot.row;
    }
    hp.numSpecies = (int)rows.size();
    put_row(hp.numSpecies, 0.0, 0.0, 0.0, 0.0, v2_meta_pack(0, ROLE_NORMAL, 0, 0, 0));     // "no particle": massless, alone in its residue
    hp.specTable.resize((size_t)(hp.numSpecies + 1) * V2_ROW_F4 * 4);
    // residues of one power-of-two size: chunks hold 32 / size whole residues at aligned lanes, a butterfly sums them
    hp.butterfly = 0;
    {
        const int k = hp.resLast[0] - hp.resFirst[0] + 1;
        bool same = (k & (k - 1)) == 0 && k <= 32;
        for (int r = 0; r < R && same; r++) same = hp.resLast[r] - hp.resFirst[r] + 1 == k;
        if (same) hp.butterfly = k;
        // residues of one size (any size up to 8): the reducing kernels may take a whole residue per lane (tgnh_v2.cuh)
        bool one = k >= 2 && k <= 8;
        for (int r = 0; r < R && one; r++) one = hp.resLast[r] - hp.resFirst[r] + 1 == k;
        if (one && k % 2 == 0) {
            // even sizes: the lanes walk their residue in rotated order and keep ONE Drude pair in registers (tgnh_v2.cuh: rpl_residue)
            std::vector<unsigned char> pairs(R, 0);
            for (int i = 0; i < N && one; i++)
                if (hp.role[i] == ROLE_DRUDE && ++pairs[p->particle_res_id[i]] > 1) one = false;
        }
        hp.uniformRes = one ? k : 0;
    }
    hp.v2 = true;
}

static int build_plan_impl(const tgnh_params* p, HostPlan& hp);

// Without the COM temperature group the reference never looks at the residues: calcCOMVelocities leaves every centre-of-mass
// velocity at zero (drudeTGNH.cu:87-108) and residue masses only enter the COM group's bookkeeping
// (CudaDrudeTGNHKernels.cpp:186-212).  Here residues also decide where tiles and warp-chunks may be cut, which is why they must be
// contiguous and contain their Drude pairs; a caller whose residue ids do not satisfy that (non-contiguous molecules, a pair
// across two residues) is served by residues of our own when the COM group is off: the particle ranges spanned by
// overlapping Drude pairs, every other particle alone.
static int build_plan(const tgnh_params* p, HostPlan& hp) {
    if (p->use_com_temp_group) return build_plan_impl(p, hp);
    const int N = p->num_particles;
    std::vector<int> glued(N + 2, 0);                   // glued[i] > 0: particle i stays with particle i - 1
    for (int k = 0; k < p->num_pairs; k++) {
        const int d = p->pair_drude[k], q = p->pair_parent[k];
        if (d < 0 || d >= N || q < 0 || q >= N) continue;            // reported by build_plan_impl
        const int a = d < q ? d : q, b = d < q ? q : d;
        glued[a + 1]++; glued[b + 1]--;
    }
    std::vector<int> resId(N);
    int r = -1, open = 0;
    for (int i = 0; i < N; i++) {
        open += glued[i];
        if (i == 0 || open == 0) r++;
        resId[i] = r;
    }
    tgnh_params q = *p;
    q.particle_res_id = resId.data();
    q.num_residues = r + 1;
    return build_plan_impl(&q, hp);
}

static int build_plan_impl(const tgnh_params* p, HostPlan& hp) {
    const int N = p->num_particles, P = p->num_pairs, R = p->num_residues, G = p->num_temp_groups;
    const int T = G + 2;
    std::vector<int>&resFirst = hp.resFirst, &resLast = hp.resLast, &partner = hp.partner;
    std::vector<double>&resMass = hp.resMass, &dofv = hp.dofv, &redMass = hp.redMass;
    std::vector<uint32_t>&role = hp.role, &desc = hp.desc;
    std::vector<int>&tileStart = hp.tileStart, &resStart = hp.resStart, &tileFirstRes = hp.tileFirstRes, &bigFirst = hp.bigFirst, &bigLast = hp.bigLast;
    double& drudeDof = hp.drudeDof;
    bool& uniform = hp.uniform;
    // ---- residues: contiguous ranges (drudeTGNH.cu:86-101 assumes it silently; we check) ----
    resFirst.assign(R, -1); resLast.assign(R, -1);
    resMass.assign(R, 0.0);
    for (int i = 0; i < N; i++) {
        const int r = p->particle_res_id[i];
        if (r < 0 || r >= R) return fail(TGNH_ERR_INVALID_ARGUMENT, "particle %d has residue id %d outside [0,%d)", i, r, R);
        if (resFirst[r] < 0) resFirst[r] = i;
        else if (resLast[r] != i - 1)
            return fail(TGNH_ERR_UNSUPPORTED, "residue %d is not a contiguous particle range (particle %d)", r, i);
        resLast[r] = i;
        resMass[r] += p->masses[i];                                   // DrudeTGNHIntegrator.cpp:147-148
        const int tg = p->particle_temp_group[i];
        if (tg < 0 || tg >= G) return fail(TGNH_ERR_INVALID_ARGUMENT, "particle %d has temperature group %d outside [0,%d)", i, tg, G);
    }
    uniform = true;
    for (int r = 0; r < R; r++) {
        if (resFirst[r] < 0) return fail(TGNH_ERR_INVALID_ARGUMENT, "residue %d has no particles", r);
        for (int i = resFirst[r] + 1; i <= resLast[r]; i++)
            if (p->particle_temp_group[i] != p->particle_temp_group[resFirst[r]]) uniform = false;
    }

    // ---- DOF bookkeeping (CudaDrudeTGNHKernels.cpp:114-150, 186-212) ----
    dofv.assign(T, 0.0); redMass.assign(G + 1, 0.0);
    for (int i = 0; i < N; i++) {
        const int tg = p->particle_temp_group[i];
        const double mass = p->masses[i];
        if (mass != 0.0) {
            dofv[tg] += 3;
            if (p->use_com_temp_group) redMass[tg] += 3 * mass * (1.0 / resMass[p->particle_res_id[i]]);
        }
    }
    partner.assign(N, 0);
    role.assign(N, ROLE_NORMAL);
    drudeDof = 0;
    for (int i = 0; i < P; i++) {
        const int d = p->pair_drude[i], q = p->pair_parent[i];
        if (d < 0 || d >= N || q < 0 || q >= N || d == q) return fail(TGNH_ERR_INVALID_ARGUMENT, "Drude pair %d has a bad particle index", i);
        if (role[d] != ROLE_NORMAL || role[q] != ROLE_NORMAL)
            return fail(TGNH_ERR_UNSUPPORTED, "particle of Drude pair %d belongs to more than one pair", i);
        if (p->particle_temp_group[d] != p->particle_temp_group[q])
            return fail(TGNH_ERR_TEMP_GROUP, "Temperature group for drude particle must be the same as the parent particle");
        if (p->particle_res_id[d] != p->particle_res_id[q])
            return fail(TGNH_ERR_UNSUPPORTED, "Drude pair %d spans two residues", i);
        if (p->masses[d] == 0.0) return fail(TGNH_ERR_INVALID_ARGUMENT, "Drude particle %d is massless", d);
        if (p->masses[q] == 0.0)      // the reference's pair transform (drudeTGNH.cu:270-300, 330-364) divides by both masses: NaN there
            return fail(TGNH_ERR_UNSUPPORTED, "parent particle %d of Drude pair %d is massless; the pair transform is undefined for it", q, i);
        role[d] = ROLE_DRUDE; role[q] = ROLE_PARENT;
        partner[d] = q - d; partner[q] = d - q;
        dofv[p->particle_temp_group[d]] -= 3;
        drudeDof += 3;
    }
    for (int i = 0; i < p->num_constraints; i++) {
        const int a = p->constraint_p[i], b = p->constraint_p1[i];
        if (a < 0 || a >= N || b < 0 || b >= N) return fail(TGNH_ERR_INVALID_ARGUMENT, "constraint %d has a bad particle index", i);
        if (p->particle_temp_group[a] != p->particle_temp_group[b])
            return fail(TGNH_ERR_TEMP_GROUP, "Temperature group of constrained particles must be the same");
        dofv[p->particle_temp_group[a]] -= 1;
    }
    hp.comDof = p->use_com_temp_group ? 3.0 * R : 0.0;


    // ---- descriptors and residue-aligned tiles ----
    const int descLen = (N + 3) & ~3;
    desc.assign(descLen, 0u);
    for (int i = 0; i < N; i++) {
        const int r = p->particle_res_id[i];
        if (partner[i] < -128 || partner[i] > 127) return (fail(TGNH_ERR_UNSUPPORTED, "Drude pair partner of particle %d is %d particles away (limit 127)", i, partner[i]));
        const bool big = resLast[r] - resFirst[r] + 1 > MAX_RES;      // COM velocity from the pre-pass table, not from the tile
        desc[i] = big ? desc_pack(p->particle_temp_group[i], role[i], 0, 0, partner[i], true)
                      : desc_pack(p->particle_temp_group[i], role[i], i - resFirst[r], resLast[r] - i, partner[i]);
    }
    // split points inside big residues must not separate a Drude pair: unsafe[s] != 0 <=> some pair (a < b) has a < s <= b
    std::vector<int> unsafe(N + 2, 0);
    for (int i = 0; i < N; i++)
        if (partner[i] > 0) { unsafe[i + 1]++; unsafe[i + partner[i] + 1]--; }
    for (int i = 1; i <= N; i++) unsafe[i] += unsafe[i - 1];
    bigFirst.clear(); bigLast.clear();
    tileStart.assign(1, 0);
    {
        int cur = 0;                          // particles in the open tile
        int r = p->particle_res_id[0];
        int i = 0;
        while (i < N) {
            r = p->particle_res_id[i];
            const int len = resLast[r] - resFirst[r] + 1;
            if (len > MAX_RES) {
                // big residue: its own tiles, cut where no Drude pair is separated (nothing else ties its particles to a tile)
                bigFirst.push_back(i); bigLast.push_back(resLast[r]);
                if (cur > 0) { tileStart.push_back(i); cur = 0; }
                const int end = resLast[r] + 1;
                int segStart = i;
                while (end - segStart > TILE) {
                    int cut = segStart + TILE;
                    while (cut > segStart && unsafe[cut]) cut--;
                    if (cut == segStart)
                        return (fail(TGNH_ERR_UNSUPPORTED, "residue %d: no place to cut %d..%d without separating a Drude pair", r, segStart, segStart + TILE));
                    tileStart.push_back(cut);
                    segStart = cut;
                }
                if (end < N) tileStart.push_back(end);
                cur = 0;
                i = end;
                continue;
            }
            if (cur + len > TILE) { tileStart.push_back(i); cur = 0; }
            cur += len;
            i += len;
        }
        tileStart.push_back(N);
    }
    // residues in particle order (tiles are residue-aligned, so each tile owns a contiguous slice of this list)
    resStart.clear(); tileFirstRes.clear();
    {
        size_t t = 0;
        for (int i = 0; i < N;) {
            if (t < tileStart.size() && tileStart[t] == i) { tileFirstRes.push_back((int)resStart.size()); t++; }
            resStart.push_back(i);
            int next = resLast[p->particle_res_id[i]] + 1;
            if (t < tileStart.size() && tileStart[t] < next) next = tileStart[t];     // a big residue continues in the next tile
            i = next;
        }
        tileFirstRes.push_back((int)resStart.size());
        resStart.push_back(N);
        while (resStart.size() & 3) resStart.push_back(N);
    }
    (void)T;
    build_plan_v2(p, hp);
    return TGNH_OK;
}

static StreamKernel pick_v2(int kind, int ffmt, bool useCOM, bool hardwall) {
    if (kind == V2_A) {
        if (ffmt) return useCOM ? (hardwall ? tgnh_v2_kernel<V2_A, 1, true, true> : tgnh_v2_kernel<V2_A, 1, true, false>)
                                : (hardwall ? tgnh_v2_kernel<V2_A, 1, false, true> : tgnh_v2_kernel<V2_A, 1, false, false>);
        return useCOM ? (hardwall ? tgnh_v2_kernel<V2_A, 0, true, true> : tgnh_v2_kernel<V2_A, 0, true, false>)
                      : (hardwall ? tgnh_v2_kernel<V2_A, 0, false, true> : tgnh_v2_kernel<V2_A, 0, false, false>);
    }
    if (kind == V2_B) {
        if (ffmt) return useCOM ? tgnh_v2_kernel<V2_B, 1, true, false> : tgnh_v2_kernel<V2_B, 1, false, false>;
        return useCOM ? tgnh_v2_kernel<V2_B, 0, true, false> : tgnh_v2_kernel<V2_B, 0, false, false>;
    }
    if (kind == V2_S) return useCOM ? tgnh_v2_kernel<V2_S, 0, true, false> : tgnh_v2_kernel<V2_S, 0, false, false>;
    return useCOM ? tgnh_v2_kernel<V2_KE, 0, true, false> : tgnh_v2_kernel<V2_KE, 0, false, false>;
}
static StreamKernel pick_v2_fused(int kind, int ffmt, bool useCOM) {
    if (kind == V2_B) {
        if (ffmt) return useCOM ? tgnh_v2_chain_kernel<V2_B, 1, true> : tgnh_v2_chain_kernel<V2_B, 1, false>;
        return useCOM ? tgnh_v2_chain_kernel<V2_B, 0, true> : tgnh_v2_chain_kernel<V2_B, 0, false>;
    }
    return useCOM ? tgnh_v2_chain_kernel<V2_KE, 0, true> : tgnh_v2_chain_kernel<V2_KE, 0, false>;
}
static int smem_v2(int kind, int ffmt, int T, int rows) {
    if (kind == V2_A) return ffmt ? V2Layout<V2_A, 1>::bytes(T, rows) : V2Layout<V2_A, 0>::bytes(T, rows);
    if (kind == V2_B) return ffmt ? V2Layout<V2_B, 1>::bytes(T, rows) : V2Layout<V2_B, 0>::bytes(T, rows);
    if (kind == V2_S) return V2Layout<V2_S, 0>::bytes(T, rows);
    return V2Layout<V2_KE, 0>::bytes(T, rows);
}

extern "C" int tgnh_create(const tgnh_params* p, tgnh_handle** out) {
    if (!p || !out) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    *out = nullptr;
    if (int rc = validate_params(p)) return rc;
    const int N = p->num_particles, P = p->num_pairs, R = p->num_residues, G = p->num_temp_groups, M = p->num_nh_chains;
    const int T = G + 2;

    // ---- everything derived on the host (also reachable without a device through tgnh_plan_tiles) ----
    HostPlan hp;
    if (int rc = build_plan(p, hp)) return rc;
    std::vector<int>&resFirst = hp.resFirst, &resLast = hp.resLast, &tileStart = hp.tileStart, &resStart = hp.resStart, &tileFirstRes = hp.tileFirstRes,
                    &bigFirst = hp.bigFirst, &bigLast = hp.bigLast;
    std::vector<double>&dofv = hp.dofv, &redMass = hp.redMass;
    std::vector<uint32_t>& desc = hp.desc;
    double &drudeDof = hp.drudeDof, &comDof = hp.comDof;
    const bool uniform = hp.uniform;
    (void)resFirst; (void)resLast;

    int deviceCount = 0;
    if (cudaGetDeviceCount(&deviceCount) != cudaSuccess || deviceCount == 0)
        return fail(TGNH_ERR_NO_DEVICE, "no CUDA device: libtgnh has no CPU fallback");
    int device = p->device;
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(TGNH_ERR_NO_DEVICE, "device %d is sm_%d%d; libtgnh is built for sm_100a (B200) only", device, prop.major, prop.minor);

    tgnh_handle* h = new tgnh_handle();
    h->device = device; h->numSMs = prop.multiProcessorCount;
    h->N = N; h->paddedN = p->padded_num_particles; h->P = P; h->R = R; h->G = G; h->T = T; h->M = M;
    h->S = p->drude_steps_per_real_step;
    h->useDrudeNH = p->use_drude_nh_chains != 0; h->useCOM = p->use_com_temp_group != 0;
    h->ffmt = p->force_format; h->hardwall = p->max_drude_distance > 0; h->prec = p->precision;
    h->uniformGroups = uniform;
    h->dt = p->step_size; h->rmax = p->max_drude_distance;
    h->kT = TGNH_BOLTZ * p->temperature; h->kTD = TGNH_BOLTZ * p->drude_temperature;
    h->comm = p->comm;
    auto bail = [&](int rc) { tgnh_destroy(h); return rc; };

    // ---- sharded: DOF sums are global (the thermostats see the whole system) ----
    if (h->comm && h->comm->worldSize > 1) {
        std::vector<double> pack(T + G + 3, 0.0);
        for (int g = 0; g < T; g++) pack[g] = dofv[g];
        for (int g = 0; g <= G; g++) pack[T + g] = redMass[g];
        pack[T + G + 1] = drudeDof; pack[T + G + 2] = comDof;
        double* dpack = nullptr;
        if (cudaMalloc(&dpack, pack.size() * sizeof(double)) != cudaSuccess) return bail(fail(TGNH_ERR_CUDA, "cudaMalloc failed"));
        cudaMemcpy(dpack, pack.data(), pack.size() * sizeof(double), cudaMemcpyHostToDevice);
        ncclResult_t r = g_nccl.AllReduce(dpack, dpack, pack.size(), ncclDouble, ncclSum, h->comm->comm, 0);
        cudaError_t e = cudaStreamSynchronize(0);
        cudaMemcpy(pack.data(), dpack, pack.size() * sizeof(double), cudaMemcpyDeviceToHost);
        cudaFree(dpack);
        if (r != ncclSuccess || e != cudaSuccess) return bail(fail(TGNH_ERR_NCCL, "all-reduce of the DOF table failed"));
        for (int g = 0; g < T; g++) dofv[g] = pack[g];
        for (int g = 0; g <= G; g++) redMass[g] = pack[T + g];
        drudeDof = pack[T + G + 1]; comDof = pack[T + G + 2];
    }
    if (h->comm && h->comm->worldSize > 1)
        if (int rc = setup_peers(h)) return bail(rc);
    if (p->use_com_temp_group) {
        dofv[G] = comDof;                                             // :197-199
        if (p->has_cm_motion_remover) dofv[G] -= 3;                   // :204-212
    }
    dofv[G + 1] = drudeDof;                                           // :201

    // ---- thermostat masses (CudaDrudeTGNHKernels.cpp:215-235) ----
    const double realkbT = h->kT, drudekbT = h->kTD;
    const double realUnit = realkbT * p->coupling_time * p->coupling_time;
    const double drudeUnit = drudekbT * p->drude_coupling_time * p->drude_coupling_time;
    h->dof.assign(T, 0.0); h->nkbt.assign(T, 0.0); h->etaMass.assign((size_t)T * M, 0.0);
    std::vector<double> etaDotDot((size_t)T * M, 0.0);
    for (int g = 0; g <= G; g++) {
        h->dof[g] = dofv[g] - redMass[g];
        h->nkbt[g] = (dofv[g] - redMass[g]) * realkbT;
        h->etaMass[g * M] = (dofv[g] - redMass[g]) * realUnit;
        for (int i = 1; i < M; i++) {
            h->etaMass[g * M + i] = realUnit;
            etaDotDot[g * M + i] = (h->etaMass[g * M + i - 1] * 0.0 - realkbT) / h->etaMass[g * M + i];
        }
    }
    h->dof[G + 1] = drudeDof;
    h->nkbt[G + 1] = drudeDof * drudekbT;
    h->etaMass[(G + 1) * M] = drudeDof * drudeUnit;
    for (int i = 1; i < M; i++) {
        h->etaMass[(G + 1) * M + i] = drudeUnit;
        if (h->useDrudeNH) etaDotDot[(G + 1) * M + i] = (h->etaMass[(G + 1) * M + i - 1] * 0.0 - drudekbT) / h->etaMass[(G + 1) * M + i];
    }

    h->numTiles = (int)tileStart.size() - 1;
    h->kindB = uniform ? KIND_BU : KIND_B;
    h->kindKE = uniform ? KIND_KU : KIND_KE;

    // ---- device allocations ----
    auto dmalloc = [&](void** ptr, size_t bytes) { return cudaMalloc(ptr, bytes ? bytes : 16) == cudaSuccess; };
    if (!dmalloc((void**)&h->dDesc, desc.size() * 4) || !dmalloc((void**)&h->dTileStart, tileStart.size() * 4) ||
        !dmalloc((void**)&h->dResStart, resStart.size() * 4) || !dmalloc((void**)&h->dTileFirstRes, tileFirstRes.size() * 4) ||
        !dmalloc((void**)&h->dTicket, 4))
        return bail(fail(TGNH_ERR_CUDA, "cudaMalloc of the index tables failed"));
    cudaMemcpy(h->dDesc, desc.data(), desc.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(h->dTileStart, tileStart.data(), tileStart.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(h->dResStart, resStart.data(), resStart.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(h->dTileFirstRes, tileFirstRes.data(), tileFirstRes.size() * 4, cudaMemcpyHostToDevice);
    h->numBig = (int)bigFirst.size();
    if (h->numBig) {
        if (!dmalloc((void**)&h->dBigFirst, bigFirst.size() * 4) || !dmalloc((void**)&h->dBigLast, bigLast.size() * 4) ||
            !dmalloc((void**)&h->dBigCom, bigFirst.size() * sizeof(double4)))
            return bail(fail(TGNH_ERR_CUDA, "cudaMalloc of the big-residue tables failed"));
        cudaMemcpy(h->dBigFirst, bigFirst.data(), bigFirst.size() * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(h->dBigLast, bigLast.data(), bigLast.size() * 4, cudaMemcpyHostToDevice);
        cudaMemset(h->dBigCom, 0, bigFirst.size() * sizeof(double4));
    }
    cudaMemset(h->dTicket, 0, 4);

    // chain block: etaMass, invEtaMass, eta, etaDot, etaDotDot, nkbt, ke2, ke2Local, ke2Used, pending, scaleA, vscale, keSum
    const size_t TM = (size_t)T * M;
    h->chainDoubles = 4 * TM + (size_t)T * (M + 1) + 8 * T + 1;
    if (!dmalloc((void**)&h->dChain, h->chainDoubles * 8)) return bail(fail(TGNH_ERR_CUDA, "cudaMalloc of the chain state failed"));
    std::vector<double> init(h->chainDoubles, 0.0);
    {
        double* b = h->dChain;
        ChainView& c = h->chain;
        c.T = T; c.G = G; c.M = M; c.S = h->S; c.useDrudeNH = h->useDrudeNH;
        c.dt = h->dt; c.kT = h->kT; c.kTD = h->kTD;
        c.dtc = h->dt / h->S;                                         // CudaDrudeTGNHKernels.cpp:440
        chain_exp_tables(c);
        size_t o = 0;
        c.etaMass = b + o; memcpy(&init[o], h->etaMass.data(), TM * 8); o += TM;
        c.invEtaMass = b + o; for (size_t i = 0; i < TM; i++) init[o + i] = h->etaMass[i] != 0.0 ? 1.0 / h->etaMass[i] : 0.0; o += TM;
        c.eta = b + o; o += TM;
        c.etaDot = b + o; o += (size_t)T * (M + 1);
        c.etaDotDot = b + o; memcpy(&init[o], etaDotDot.data(), TM * 8); o += TM;
        c.nkbt = b + o; memcpy(&init[o], h->nkbt.data(), T * 8); o += T;
        c.ke2 = b + o; o += T;
        c.ke2Local = b + o; o += T;
        c.ke2Used = b + o; o += T;
        c.pending = b + o; for (int g = 0; g < T; g++) init[o + g] = 1.0; o += T;
        c.scaleA = b + o; for (int g = 0; g < T; g++) init[o + g] = 1.0; o += T;
        c.vscale = b + o; for (int g = 0; g < T; g++) init[o + g] = 1.0; o += T;
        c.keSum = b + o; o += 1;
        c.expHint = b + o; o += T;
        if (!(h->comm && h->comm->worldSize > 1)) c.ke2Local = nullptr;
    }
    cudaMemcpy(h->dChain, init.data(), h->chainDoubles * 8, cudaMemcpyHostToDevice);

    int rc;
    if ((rc = configure_kernel(h, KIND_A, &h->gridA, &h->smemA)) || (rc = configure_kernel(h, h->kindB, &h->gridB, &h->smemB)) ||
        (rc = configure_kernel(h, h->kindKE, &h->gridKE, &h->smemKE)) || (rc = configure_kernel(h, KIND_K, &h->gridK, &h->smemK)) || (rc = configure_kernel(h, KIND_A1, &h->gridA1, &h->smemA1)) ||
        (rc = configure_kernel(h, KIND_A2, &h->gridA2, &h->smemA2)) || (rc = configure_kernel(h, KIND_S, &h->gridS, &h->smemS)))
        return bail(rc);
    {
        // small systems (one tile per SM at most, nothing sharded, no big residues): chain update inside the reducing launch
        const char* e = getenv("TGNH_FUSE_CHAIN");
        h->fuseChain = h->numTiles <= h->numSMs && h->numBig == 0 && !(h->comm && h->comm->worldSize > 1) && !(e && atoi(e) == 0);
        if (h->fuseChain)
            for (int kind : {h->kindB, h->kindKE}) {
                StreamKernel k = pick_fused(kind, h->ffmt, h->prec, h->useCOM);
                const int smem = smem_bytes(kind, h->ffmt, h->prec, h->useCOM, h->T);
                if (cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
                    (void)cudaGetLastError();
                    h->fuseChain = false;
                }
            }
    }
    // ---- warp-chunk kernels ----
    {
        const char* e = getenv("TGNH_V2");
        h->v2 = hp.v2 && !(e && atoi(e) == 0);
        const char* lz = getenv("TGNH_LAZY_KICK");
        h->lazyKick = !(lz && atoi(lz) == 0);
        if (h->v2) {
            h->numTiles2 = hp.numTiles2; h->maxRes = hp.maxRes; h->numSpecies = hp.numSpecies; h->butterfly = hp.butterfly;
            const char* rp = getenv("TGNH_RPL");
            h->rplMode = rp ? atoi(rp) : 2;
            h->rpl = (h->rplMode > 0 && h->useCOM && uniform) ? hp.uniformRes : 0;
            h->hChunkStart = hp.chunkStart;
            if (!dmalloc((void**)&h->dSpec, hp.spec.size()) || !dmalloc((void**)&h->dChunkStart, hp.chunkStart.size() * 4) ||
                !dmalloc((void**)&h->dSpecTable, hp.specTable.size() * 4))
                return bail(fail(TGNH_ERR_CUDA, "cudaMalloc of the species tables failed"));
            cudaMemcpy(h->dSpec, hp.spec.data(), hp.spec.size(), cudaMemcpyHostToDevice);
            cudaMemcpy(h->dChunkStart, hp.chunkStart.data(), hp.chunkStart.size() * 4, cudaMemcpyHostToDevice);
            cudaMemcpy(h->dSpecTable, hp.specTable.data(), hp.specTable.size() * 4, cudaMemcpyHostToDevice);
            int* grids[4] = {&h->gridA2v, &h->gridB2v, &h->gridKE2v, &h->gridS2v};
            int* smems[4] = {&h->smemA2v, &h->smemB2v, &h->smemKE2v, &h->smemS2v};
            for (int kind = 0; kind < 4 && h->v2; kind++) {
                StreamKernel k = pick_v2(kind, h->ffmt, h->useCOM, h->hardwall);
                const int smem = smem_v2(kind, h->ffmt, h->T, h->numSpecies + 1);
                int occ = 0;
                if (smem > 227 * 1024 || cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess ||
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)k, V2_THREADS, smem) != cudaSuccess || occ < 1) {
                    (void)cudaGetLastError();
                    h->v2 = false;              // e.g. 30 temperature groups: the energy columns do not fit beside the ring
                    break;
                }
                int g = occ * h->numSMs;
                if (g > h->numTiles2) g = h->numTiles2;
                *grids[kind] = g < 1 ? 1 : g;
                *smems[kind] = smem;
            }
            if (h->v2 && h->fuseChain) {
                h->fuseChain = h->numTiles2 <= h->numSMs;
                for (int kind : {V2_B, V2_KE}) {
                    if (!h->fuseChain) break;
                    if (cudaFuncSetAttribute((const void*)pick_v2_fused(kind, h->ffmt, h->useCOM), cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             smem_v2(kind, h->ffmt, h->T, h->numSpecies + 1)) != cudaSuccess) {
                        (void)cudaGetLastError();
                        h->fuseChain = false;
                    }
                }
            }
        }
    }
    int maxGrid = h->gridA > h->gridB ? h->gridA : h->gridB;
    if (h->gridA2v > maxGrid) maxGrid = h->gridA2v;
    if (h->gridB2v > maxGrid) maxGrid = h->gridB2v;
    if (h->gridKE2v > maxGrid) maxGrid = h->gridKE2v;
    if (h->numTiles2 > maxGrid && h->fuseChain) maxGrid = h->numTiles2;
    if (h->numTiles > maxGrid && h->fuseChain) maxGrid = h->numTiles;
    if (h->gridKE > maxGrid) maxGrid = h->gridKE;
    if (h->gridA1 > maxGrid) maxGrid = h->gridA1;
    if (h->gridA2 > maxGrid) maxGrid = h->gridA2;
    if (!dmalloc((void**)&h->dPartials, (size_t)maxGrid * T * 8)) return bail(fail(TGNH_ERR_CUDA, "cudaMalloc of the partial sums failed"));
    {
        // the chain kernel is configured nowhere above: load it now (CUDA loads a kernel's code at its first use; that is 1-40 ms
        // of host time which would otherwise fall into the first step that needs the stand-alone chain launch)
        cudaFuncAttributes fa;
        if (cudaFuncGetAttributes(&fa, (const void*)tgnh_chain_kernel) != cudaSuccess) (void)cudaGetLastError();
        if (h->numBig) {
            const void* bk[4] = {(const void*)tgnh_bigcom_kernel<0, 0>, (const void*)tgnh_bigcom_kernel<0, 1>, (const void*)tgnh_bigcom_kernel<1, 0>, (const void*)tgnh_bigcom_kernel<1, 1>};
            for (const void* k : bk)
                if (cudaFuncGetAttributes(&fa, k) != cudaSuccess) (void)cudaGetLastError();
        }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return bail(fail(TGNH_ERR_CUDA, "device error during tgnh_create: %s", cudaGetErrorString(cudaGetLastError())));
    *out = h;
    return TGNH_OK;
}

extern "C" int tgnh_plan_tiles(const tgnh_params* p, int32_t* tile_start, int32_t capacity, int32_t* num_tiles, int32_t* num_big_residues,
                               int32_t* residue_uniform) {
    if (!p) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = validate_params(p)) return rc;
    HostPlan hp;
    if (int rc = build_plan(p, hp)) return rc;
    const int n = (int)hp.tileStart.size() - 1;
    if (num_tiles) *num_tiles = n;
    if (num_big_residues) *num_big_residues = (int)hp.bigFirst.size();
    if (residue_uniform) *residue_uniform = hp.uniform ? 1 : 0;
    if (tile_start) {
        if (capacity < n + 1) return fail(TGNH_ERR_INVALID_ARGUMENT, "tile_start holds %d entries, %d are needed", capacity, n + 1);
        for (int i = 0; i <= n; i++) tile_start[i] = hp.tileStart[i];
    }
    return TGNH_OK;
}

extern "C" int tgnh_plan_descriptors(const tgnh_params* p, uint32_t* desc_out) {
    if (!p || !desc_out) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = validate_params(p)) return rc;
    HostPlan hp;
    if (int rc = build_plan(p, hp)) return rc;
    for (int i = 0; i < p->num_particles; i++) desc_out[i] = hp.desc[i];
    return TGNH_OK;
}

extern "C" int tgnh_plan_chunks(const tgnh_params* p, int32_t* chunk_start, int32_t capacity, int32_t* num_chunks, uint8_t* species_out,
                                float* table_out, int32_t* num_species, int32_t* max_residue) {
    if (!p) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    if (int rc = validate_params(p)) return rc;
    HostPlan hp;
    if (int rc = build_plan(p, hp)) return rc;
    if (!hp.v2) return fail(TGNH_ERR_UNSUPPORTED, "the warp-chunk kernels do not cover this system: %s", hp.v2why);
    const int n = V2_NCONS * hp.numTiles2;
    if (num_chunks) *num_chunks = n;
    if (num_species) *num_species = hp.numSpecies;
    if (max_residue) *max_residue = hp.maxRes;
    if (chunk_start) {
        if (capacity < n + 1) return fail(TGNH_ERR_INVALID_ARGUMENT, "chunk_start holds %d entries, %d are needed", capacity, n + 1);
        for (int i = 0; i <= n; i++) chunk_start[i] = hp.chunkStart[i];
    }
    if (species_out) memcpy(species_out, hp.spec.data(), p->num_particles);
    if (table_out) {
        memset(table_out, 0, 256 * 8 * sizeof(float));
        for (int r = 0; r <= hp.numSpecies; r++) memcpy(table_out + 8 * r, &hp.specTable[(size_t)r * V2_ROW_F4 * 4], 32);
    }
    return TGNH_OK;
}

extern "C" int tgnh_chunks_per_tile(void) { return V2_NCONS; }
extern "C" int tgnh_kernel_generation(const tgnh_handle* h) { return h ? (h->v2 ? 2 : 1) : 0; }
extern "C" int tgnh_residue_per_lane(const tgnh_handle* h) { return h && h->v2 && !h->fuseChain ? h->rpl : 0; }
extern "C" int tgnh_lazy_second_kick(const tgnh_handle* h) { return h && h->lazyKick && h->v2 && !h->fuseChain && h->uniformGroups ? 1 : 0; }

extern "C" void tgnh_destroy(tgnh_handle* h) {
    if (!h) return;
    for (int r = 0; r < MAX_PEERS; r++)
        if (h->peers.inbox[r] && h->peers.inbox[r] != h->dInbox) cudaIpcCloseMemHandle(h->peers.inbox[r]);
    cudaFree(h->dInbox);
    cudaFree(h->dDesc); cudaFree(h->dTileStart); cudaFree(h->dResStart); cudaFree(h->dTileFirstRes); cudaFree(h->dBigFirst); cudaFree(h->dBigLast); cudaFree(h->dBigCom); cudaFree(h->dChain); cudaFree(h->dPartials); cudaFree(h->dTicket);
    cudaFree(h->dSpec); cudaFree(h->dChunkStart); cudaFree(h->dSpecTable);
    cudaFree(h->hsVelm); cudaFree(h->hsPosq); cudaFree(h->hsForce); cudaFree(h->hsCorr);
    for (cudaEvent_t e : h->hsEvents) cudaEventDestroy(e);
    if (h->hsIn) cudaStreamDestroy(h->hsIn);
    if (h->hsOut) cudaStreamDestroy(h->hsOut);
    for (cudaEvent_t e : h->evPool) cudaEventDestroy(e);
    if (h->hsStream) cudaStreamDestroy(h->hsStream);
    delete h;
}

// ------------------------------------------------------------------------------------------------
// launches
// ------------------------------------------------------------------------------------------------
static bool sharded(const tgnh_handle* h) { return h->comm && h->comm->worldSize > 1; }

static int check_ptrs(const tgnh_handle* h, const void* velm, const void* posq, const void* force, bool needX, bool needF) {
    if (!h) return fail(TGNH_ERR_INVALID_ARGUMENT, "null handle");
    if (!velm || ((uintptr_t)velm & 15)) return fail(TGNH_ERR_INVALID_ARGUMENT, "velm must be a 16-byte aligned device pointer");
    if (needX && (!posq || ((uintptr_t)posq & 15))) return fail(TGNH_ERR_INVALID_ARGUMENT, "posq must be a 16-byte aligned device pointer");
    if (needF && (!force || ((uintptr_t)force & 15))) return fail(TGNH_ERR_INVALID_ARGUMENT, "force must be a 16-byte aligned device pointer");
    return TGNH_OK;
}

// every launch carries the programmatic-stream-serialization attribute: the kernels call
// griddepcontrol.wait before touching global data, so the next launch's prologue (barrier init, smem
// clears, residency) overlaps the previous launch's tail instead of paying a full launch gap
template <typename... Args>
static cudaError_t launch_pdl(void (*kernel)(Args...), int grid, int block, int smem, cudaStream_t s, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

// `gather`: the launch follows a reducing launch of a sharded run and first forms the global sums from all ranks' partials
static int launch_chain(tgnh_handle* h, cudaStream_t s, int mode, bool gather = false) {
    PeerView pv = h->peers;
    pv.seq = h->reduceSeq;
    if (!gather) pv.world = 0;
    CUDA_TRY(launch_pdl(tgnh_chain_kernel, 1, 32, 0, s, h->chain, (const PeerView)pv, mode));
    h->launches++;
    return TGNH_OK;
}

// one streaming launch; for the reducing kinds (KIND_B / KIND_KE) followed by the all-reduce of the
// kinetic-energy vector when sharded and by the chain update `chainMode` asks for
static int launch_stream(tgnh_handle* h, cudaStream_t s, int kind, void* velm, void* posq, const void* force, int applyScale, int chainMode,
                         void* posDelta = nullptr) {
    StreamArgs a{};
    a.velm = velm; a.posq = posq; a.posqCorrection = (float4*)h->posqCorrection; a.force = force; a.posDelta = posDelta;
    if (h->prec == TGNH_PRECISION_MIXED && posq != nullptr && h->posqCorrection == nullptr)
        return fail(TGNH_ERR_INVALID_ARGUMENT, "mixed precision: register the posqCorrection array with tgnh_set_posq_correction first");
    a.desc = h->dDesc; a.tileStart = h->dTileStart; a.numTiles = h->numTiles; a.paddedN = h->paddedN;
    a.resStart = h->dResStart; a.tileFirstRes = h->dTileFirstRes;
    a.bigFirst = h->dBigFirst; a.bigCom = h->dBigCom; a.numBig = h->numBig;
    const bool firstHalf = kind == KIND_A || kind == KIND_A1 || kind == KIND_A2;
    const int prof = firstHalf ? KIND_A : (kind == KIND_S || kind == KIND_KE) ? KIND_KE : KIND_B;      // profiling slot (first half / second half / reduce+scale)
    const bool reduces = !firstHalf && kind != KIND_S && kind != KIND_K;
    if (kind == KIND_KE && applyScale == 0) kind = h->kindKE;
    if (kind == KIND_B) kind = h->kindB;
    a.dt = h->dt;
    a.fscale = h->ffmt == TGNH_FORCE_I64_SOA ? 0.5 * h->dt / 4294967296.0 : 0.5 * h->dt;             // CudaDrudeTGNHKernels.cpp:295
    a.rmax = h->rmax;
    a.hardwallScale = std::sqrt(h->kTD);                                                              // :299
    a.applyScale = applyScale;
    a.useLocalKE = sharded(h) ? 1 : 0;
    // L2 hand-over: every streaming launch walks the tiles in the direction opposite to the previous one, so it starts
    // on what that launch wrote last (first half forward, second half backward, a scaling pass forward again, ...)
    a.reverse = h->nextReverse;
    h->nextReverse ^= 1;
    static const int tunePrefetch = getenv("TGNH_TUNE_PREFETCH") ? atoi(getenv("TGNH_TUNE_PREFETCH")) : 0;                  // experiments only
    a.prologuePrefetch = (prof == KIND_A) ? tunePrefetch : 0;
    static const int tuneReverse = getenv("TGNH_TUNE_REVERSE") ? atoi(getenv("TGNH_TUNE_REVERSE")) : -1;   // experiments only
    if (tuneReverse == 0) a.reverse = 0;
    a.partials = h->dPartials; a.ticket = h->dTicket; a.chain = h->chain;
    a.peers = h->peers;
    const bool p2p = h->peers.world > 1 && reduces;
    if (p2p) a.peers.seq = ++h->reduceSeq; else a.peers.world = 0;
    const int grid = kind == KIND_K ? h->gridK : kind == KIND_S ? h->gridS : kind == KIND_A1 ? h->gridA1 : kind == KIND_A2 ? h->gridA2 : prof == KIND_A ? h->gridA : prof == KIND_B ? h->gridB : h->gridKE;
    const int smem = kind == KIND_K ? h->smemK : kind == KIND_S ? h->smemS : kind == KIND_A1 ? h->smemA1 : kind == KIND_A2 ? h->smemA2 : prof == KIND_A ? h->smemA : prof == KIND_B ? h->smemB : h->smemKE;
    StreamKernel k = pick(kind, h->ffmt, h->prec, h->useCOM, h->hardwall, h->numBig > 0);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->profiling) {
        if (h->evUsed + 2 > h->evPool.size()) {
            cudaEvent_t a0, a1;
            CUDA_TRY(cudaEventCreate(&a0));
            CUDA_TRY(cudaEventCreate(&a1));
            h->evPool.push_back(a0); h->evPool.push_back(a1);
        }
        e0 = h->evPool[h->evUsed]; e1 = h->evPool[h->evUsed + 1];
        h->evKind.resize(h->evUsed / 2 + 1);
        h->evKind[h->evUsed / 2] = prof;
        h->evUsed += 2;
        CUDA_TRY(cudaEventRecord(e0, s));
    }
    if (h->numBig && h->useCOM && kind != KIND_A2 && kind != KIND_K) {
        // residues that do not fit a tile: their COM velocity (as this launch will see / store the velocities) first
        BigComArgs b;
        b.velm = velm; b.force = force; b.bigFirst = h->dBigFirst; b.bigLast = h->dBigLast; b.bigCom = h->dBigCom; b.paddedN = h->paddedN;
        b.fscale = (prof == KIND_B) ? a.fscale : 0.0;
        void (*bk)(const BigComArgs) = h->prec ? (h->ffmt ? tgnh_bigcom_kernel<1, 1> : tgnh_bigcom_kernel<0, 1>)
                                               : (h->ffmt ? tgnh_bigcom_kernel<1, 0> : tgnh_bigcom_kernel<0, 0>);
        CUDA_TRY(launch_pdl(bk, h->numBig, 256, 0, s, (const BigComArgs)b));
        h->launches++;
    }
    const bool fused = h->fuseChain && reduces && chainMode != CHAIN_NONE && (kind == h->kindB || (kind == h->kindKE && !applyScale));
    a.fusedChainMode = fused ? chainMode : CHAIN_NONE;
    a.earlyLoads = h->earlyOK ? 1 : 0;
    a.lazyKick = h->lazyNow ? 1 : 0;
    a.uniformGroups = h->uniformGroups ? 1 : 0;
    // the two halves and the plain reduction run through the warp-chunk kernels where the system qualifies
    const int kind2 = !h->v2 ? -1 : kind == KIND_A ? V2_A : kind == h->kindB ? V2_B : (kind == h->kindKE && !applyScale) ? V2_KE : kind == KIND_S ? V2_S : -1;
    if (kind2 >= 0) {
        a.spec = h->dSpec; a.chunkStart = h->dChunkStart; a.specTable = h->dSpecTable; a.maxRes = h->maxRes; a.butterfly = h->butterfly; a.tableRows = h->numSpecies + 1;
        a.numTiles = h->numTiles2;
        a.resPerLane = (!fused && (kind2 == V2_KE || (kind2 == V2_B && (h->rplMode >= 2 || a.lazyKick)))) ? h->rpl : 0;
        const int grid2 = kind2 == V2_A ? h->gridA2v : kind2 == V2_B ? h->gridB2v : kind2 == V2_S ? h->gridS2v : h->gridKE2v;
        const int smem2 = kind2 == V2_A ? h->smemA2v : kind2 == V2_B ? h->smemB2v : kind2 == V2_S ? h->smemS2v : h->smemKE2v;
        if (fused) CUDA_TRY(launch_pdl(pick_v2_fused(kind2, h->ffmt, h->useCOM), h->numTiles2, V2_THREADS, smem2, s, (const StreamArgs)a));
        else CUDA_TRY(launch_pdl(pick_v2(kind2, h->ffmt, h->useCOM, h->hardwall), grid2, V2_THREADS, smem2, s, (const StreamArgs)a));
    } else if (fused) CUDA_TRY(launch_pdl(pick_fused(kind, h->ffmt, h->prec, h->useCOM), h->numTiles, TILE, smem, s, (const StreamArgs)a));
    else CUDA_TRY(launch_pdl(k, grid, TILE, smem, s, (const StreamArgs)a));
    if (e1) CUDA_TRY(cudaEventRecord(e1, s));
    h->launches++;
    if (!reduces || fused) return TGNH_OK;
    // the only exchange on the path: double[G+2] kinetic-energy partials.  Peer inboxes: published by the launch above,
    // gathered by the chain launch below (which therefore runs even without a chain update); otherwise an NCCL all-reduce
    if (p2p) return launch_chain(h, s, chainMode, true);
    if (sharded(h)) NCCL_TRY(g_nccl.AllReduce(h->chain.ke2Local, h->chain.ke2, h->T, ncclDouble, ncclSum, h->comm->comm, s));
    if (chainMode != CHAIN_NONE) return launch_chain(h, s, chainMode);
    return TGNH_OK;
}

// make chain.ke2 describe velm; runs CHAIN_FIRST in the same launch when asked to
static int ensure_ke(tgnh_handle* h, cudaStream_t s, void* velm, int chainMode) {
    if (h->keValid) return chainMode == CHAIN_NONE ? TGNH_OK : launch_chain(h, s, chainMode);
    int rc = launch_stream(h, s, KIND_KE, velm, nullptr, nullptr, 0, chainMode);
    if (rc == TGNH_OK) h->keValid = true;
    return rc;
}

// apply chain.pending to velm (integrateDrudeTGNHChain) and refresh ke2 from the scaled velocities
static int flush_scale(tgnh_handle* h, cudaStream_t s, void* velm) {
    if (!h->scalePending) return TGNH_OK;
    if (h->uniformGroups) {
        // residue-uniform groups: one scaling pass; the kernel also turns ke2 into s_g^2 ke2 and resets pending to 1
        int rc = launch_stream(h, s, KIND_S, velm, nullptr, nullptr, 1, CHAIN_NONE);
        if (rc) return rc;
        h->scalePending = false;
        h->keValid = true;
        return TGNH_OK;
    }
    // residues that span temperature groups: COM velocities do not simply scale; apply, then recompute from scratch
    int rc = launch_stream(h, s, KIND_KE, velm, nullptr, nullptr, 1, CHAIN_NONE);
    if (rc) return rc;
    h->scalePending = false;      // the kernel's last CTA resets pending to 1 (scaleA keeps the factors that were just applied)
    h->keValid = false;
    return ensure_ke(h, s, velm, CHAIN_NONE);
}

extern "C" int tgnh_half1(tgnh_handle* h, void* stream, void* velm, void* posq, const void* force) {
    if (int rc = check_ptrs(h, velm, posq, force, true, true)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(h->device));
    if (int rc = ensure_ke(h, s, velm, CHAIN_FIRST)) return rc;      // :336-337, chain on the device
    h->scalePending = false;                                         // folded into scaleA
    h->keValid = false;
    return launch_stream(h, s, KIND_A, velm, posq, force, 1, CHAIN_NONE);   // :351-376
}

extern "C" int tgnh_half1_kick(tgnh_handle* h, void* stream, void* velm, const void* force, void* pos_delta) {
    if (int rc = check_ptrs(h, velm, pos_delta, force, true, true)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(h->device));
    if (int rc = ensure_ke(h, s, velm, CHAIN_FIRST)) return rc;
    h->scalePending = false;
    h->keValid = false;
    return launch_stream(h, s, KIND_A1, velm, nullptr, force, 1, CHAIN_NONE, pos_delta);
}

extern "C" int tgnh_half1_drift(tgnh_handle* h, void* stream, void* velm, void* posq, void* pos_delta) {
    if (int rc = check_ptrs(h, velm, posq, pos_delta, true, true)) return rc;
    CUDA_TRY(cudaSetDevice(h->device));
    h->keValid = false;
    return launch_stream(h, (cudaStream_t)stream, KIND_A2, velm, posq, nullptr, 0, CHAIN_NONE, pos_delta);
}

extern "C" int tgnh_thermostat(tgnh_handle* h, void* stream, void* velm, int flags) {
    if (int rc = check_ptrs(h, velm, nullptr, nullptr, false, false)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(h->device));
    if (h->scalePending) return fail(TGNH_ERR_INVALID_ARGUMENT, "a deferred velocity scaling is pending; call tgnh_flush first");
    h->keValid = false;                                              // the caller's velocity constraints changed velm
    if (int rc = ensure_ke(h, s, velm, CHAIN_SECOND)) return rc;
    h->scalePending = true;
    if ((flags & TGNH_HALF2_DEFER_SCALE) && h->uniformGroups) return TGNH_OK;
    return flush_scale(h, s, velm);
}

extern "C" int tgnh_half2(tgnh_handle* h, void* stream, void* velm, const void* force, int flags) {
    if (int rc = check_ptrs(h, velm, nullptr, force, false, true)) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(h->device));
    if (flags & TGNH_HALF2_KICK_ONLY) {
        // constrained systems: OpenMM's velocity constraints (:391) come between the kick and the thermostat half-step;
        // the kinetic energies this launch reduces are discarded and tgnh_thermostat recomputes them
        if (int rc = launch_stream(h, s, KIND_K, velm, nullptr, force, 0, CHAIN_NONE)) return rc;
        h->keValid = false;
        return TGNH_OK;
    }
    if (int rc = launch_stream(h, s, KIND_B, velm, nullptr, force, 0, CHAIN_SECOND)) return rc;   // :384-395
    h->keValid = true;
    h->scalePending = true;
    if ((flags & TGNH_HALF2_DEFER_SCALE) && h->uniformGroups) return TGNH_OK;
    return flush_scale(h, s, velm);                                  // :402
}

extern "C" int tgnh_flush(tgnh_handle* h, void* stream, void* velm) {
    if (int rc = check_ptrs(h, velm, nullptr, nullptr, false, false)) return rc;
    CUDA_TRY(cudaSetDevice(h->device));
    return flush_scale(h, (cudaStream_t)stream, velm);
}

extern "C" int tgnh_step(tgnh_handle* h, void* stream, void* velm, void* posq, const void* force, int nsteps) {
    if (int rc = check_ptrs(h, velm, posq, force, true, true)) return rc;
    if (nsteps < 0) return fail(TGNH_ERR_INVALID_ARGUMENT, "nsteps must be >= 0");
    if (nsteps == 0) return TGNH_OK;
    cudaStream_t s = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(h->device));
    if (!h->uniformGroups) {
        for (int i = 0; i < nsteps; i++) {
            if (int rc = tgnh_half1(h, stream, velm, posq, force)) return rc;
            if (int rc = tgnh_half2(h, stream, velm, force, TGNH_HALF2_DEFAULT)) return rc;
        }
        return TGNH_OK;
    }
    if (int rc = ensure_ke(h, s, velm, CHAIN_FIRST)) return rc;
    struct EarlyGuard { tgnh_handle* h; ~EarlyGuard() { h->earlyOK = false; h->lazyNow = false; } } guard{h};
    // Lazy second kick (warp-chunk kernels): between two steps of this call nothing reads velm but the two halves themselves, and
    // the second half of step i and the first half of step i+1 see the same forces.  The second half then only REDUCES the
    // energies of v + (dt/2) F/m and stores nothing; the next first half repeats that fp32 operation on the same operands
    // (bit-identical velocities) before it scales, kicks and drifts: 16 B per particle and step less through HBM.  The last
    // step's second half stores as usual, so velm is complete when the call returns.
    const bool lazy = h->lazyKick && h->v2 && !h->fuseChain;
    for (int i = 0; i < nsteps; i++) {
        // from the second launch on, everything that can still be running ahead of a launch is this loop's own work, which
        // writes velm only (and posq in first-half launches that have completed by then): see StreamArgs::earlyLoads
        h->lazyNow = lazy && i > 0;
        if (int rc = launch_stream(h, s, KIND_A, velm, posq, force, 1, CHAIN_NONE)) return rc;
        h->earlyOK = true;
        // the thermostat half-step that ends step i and the one that begins step i+1 run back to back in one
        // chain launch; their scale factors are applied together by the next first-half pass
        const int mode = (i + 1 < nsteps) ? CHAIN_SECOND_FIRST : CHAIN_SECOND;
        h->lazyNow = lazy && i + 1 < nsteps;
        if (int rc = launch_stream(h, s, KIND_B, velm, nullptr, force, 0, mode)) return rc;
    }
    h->lazyNow = false;
    h->keValid = true;
    h->scalePending = true;
    return flush_scale(h, s, velm);
}

extern "C" int tgnh_set_posq_correction(tgnh_handle* h, void* posq_correction) {
    if (!h) return fail(TGNH_ERR_INVALID_ARGUMENT, "null handle");
    if (h->prec != TGNH_PRECISION_MIXED) return fail(TGNH_ERR_INVALID_ARGUMENT, "posqCorrection exists in mixed precision only");
    if (!posq_correction || ((uintptr_t)posq_correction & 15)) return fail(TGNH_ERR_INVALID_ARGUMENT, "posq_correction must be a 16-byte aligned device pointer");
    h->posqCorrection = posq_correction;
    return TGNH_OK;
}

extern "C" int tgnh_invalidate(tgnh_handle* h) {
    if (!h) return fail(TGNH_ERR_INVALID_ARGUMENT, "null handle");
    h->keValid = false;
    return TGNH_OK;
}

extern "C" int tgnh_step_host2(tgnh_handle* h, void* velm_host, void* posq_host, void* posq_correction_host, const void* force_host, int nsteps,
                               int flags, double* ke2_host);
extern "C" int tgnh_step_host(tgnh_handle* h, void* velm_host, void* posq_host, const void* force_host, int nsteps, double* ke2_host) {
    if (h && h->prec == TGNH_PRECISION_MIXED)
        return fail(TGNH_ERR_UNSUPPORTED, "mixed precision also needs the posqCorrection array: use tgnh_step_host2");
    return tgnh_step_host2(h, velm_host, posq_host, nullptr, force_host, nsteps, 0, ke2_host);
}

// One launch of a warp-chunk kernel over the tiles [tileBegin, tileBegin + tileCount): the building block of the pipelined
// host-buffer path.  `accumulate`: add the energy sums to those of the previous sub-range; `last`: the launch that completes the
// reduction (publishes to the peers when sharded).
static int launch_v2_range(tgnh_handle* h, cudaStream_t s, int kind2, void* velm, void* posq, const void* force, int tileBegin, int tileCount,
                           bool accumulate, bool last) {
    StreamArgs a{};
    a.velm = velm; a.posq = posq; a.force = force;
    a.desc = h->dDesc; a.tileStart = h->dTileStart; a.paddedN = h->paddedN;
    a.dt = h->dt;
    a.fscale = h->ffmt == TGNH_FORCE_I64_SOA ? 0.5 * h->dt / 4294967296.0 : 0.5 * h->dt;
    a.rmax = h->rmax;
    a.hardwallScale = std::sqrt(h->kTD);
    a.useLocalKE = sharded(h) ? 1 : 0;
    a.partials = h->dPartials; a.ticket = h->dTicket; a.chain = h->chain;
    a.peers = h->peers;
    const bool reduces = kind2 != V2_A;
    if (h->peers.world > 1 && reduces && last) a.peers.seq = ++h->reduceSeq; else a.peers.world = 0;
    a.spec = h->dSpec; a.chunkStart = h->dChunkStart; a.specTable = h->dSpecTable; a.maxRes = h->maxRes; a.butterfly = h->butterfly; a.tableRows = h->numSpecies + 1;
    a.numTiles = tileCount; a.tileBegin = tileBegin; a.accumulate = accumulate ? 1 : 0;
    a.resPerLane = (kind2 == V2_KE || (kind2 == V2_B && h->rplMode >= 2)) ? h->rpl : 0;
    a.uniformGroups = h->uniformGroups ? 1 : 0;
    int grid = kind2 == V2_A ? h->gridA2v : kind2 == V2_B ? h->gridB2v : h->gridKE2v;
    if (grid > tileCount) grid = tileCount;
    const int smem = kind2 == V2_A ? h->smemA2v : kind2 == V2_B ? h->smemB2v : h->smemKE2v;
    CUDA_TRY(launch_pdl(pick_v2(kind2, h->ffmt, h->useCOM, h->hardwall), grid, V2_THREADS, smem, s, (const StreamArgs)a));
    h->launches++;
    return TGNH_OK;
}

// One step with the state in HOST memory, pipelined over HS_CHUNKS particle ranges and three streams:
//   copy-in :  velm chunks first, then posq + force chunks                          (host -> device engine)
//   compute :  energies of chunk c as soon as its velm has landed; chain; then first half + second half of chunk c as soon
//              as its posq / forces have landed; chain; scaling
//   copy-out:  posq of chunk c as soon as its first half is done (device -> host engine, concurrent with the uploads that are
//              still running); velm after the scaling
// upload / download: whether this step starts from / ends in the host buffers (the inner steps of a multi-step call do neither).
static constexpr int HS_CHUNKS = 8;
static int pipelined_host_step(tgnh_handle* h, void* velm_host, void* posq_host, const void* force_host, bool upload, bool uploadForce, bool download,
                               double* ke2_host) {
    cudaStream_t sc = h->hsStream, si = h->hsIn, so = h->hsOut;
    const int nt = h->numTiles2, K = nt < HS_CHUNKS ? nt : HS_CHUNKS;
    const size_t fb = h->ffmt == TGNH_FORCE_I64_SOA ? 8 : 4;
    float4* dv = (float4*)h->hsVelm;
    float4* dx = (float4*)h->hsPosq;
    unsigned char* df = (unsigned char*)h->hsForce;
    auto tile0 = [&](int c) { return (int)((long long)nt * c / K); };
    auto part0 = [&](int c) { return h->hChunkStart[(size_t)V2_NCONS * tile0(c)]; };     // first particle of chunk c (c == K: N)
    cudaEvent_t* ev = h->hsEvents.data();            // [0,K) velm in, [K,2K) posq+force in, [2K,3K) first half done, [3K] compute ready, [3K+1] scaled
    if (upload) {
        CUDA_TRY(cudaEventRecord(ev[3 * K], sc));                      // the copies must not overtake earlier work on the compute stream
        CUDA_TRY(cudaStreamWaitEvent(si, ev[3 * K], 0));
        for (int c = 0; c < K; c++) {
            const int p0 = part0(c), p1 = part0(c + 1);
            CUDA_TRY(cudaMemcpyAsync(dv + p0, (const float4*)velm_host + p0, (size_t)(p1 - p0) * 16, cudaMemcpyHostToDevice, si));
            CUDA_TRY(cudaEventRecord(ev[c], si));
        }
        for (int c = 0; c < K; c++) {
            const int p0 = part0(c), p1 = part0(c + 1);
            CUDA_TRY(cudaMemcpyAsync(dx + p0, (const float4*)posq_host + p0, (size_t)(p1 - p0) * 16, cudaMemcpyHostToDevice, si));
            if (uploadForce) {
                // forces are SoA: three slices per chunk, rounded outwards to the 16-byte windows the kernels fetch
                const int f0 = p0 & ~3, f1 = c + 1 == K ? h->paddedN : ((p1 + 3) & ~3);
                for (int k = 0; k < 3; k++)
                    CUDA_TRY(cudaMemcpyAsync(df + ((size_t)k * h->paddedN + f0) * fb, (const unsigned char*)force_host + ((size_t)k * h->paddedN + f0) * fb,
                                             (size_t)(f1 - f0) * fb, cudaMemcpyHostToDevice, si));
            }
            CUDA_TRY(cudaEventRecord(ev[K + c], si));
        }
        h->keValid = false;
    }
    // thermostat half-step that begins the step
    if (!h->keValid) {
        for (int c = 0; c < K; c++) {
            if (upload) CUDA_TRY(cudaStreamWaitEvent(sc, ev[c], 0));
            if (int rc = launch_v2_range(h, sc, V2_KE, dv, nullptr, nullptr, tile0(c), tile0(c + 1) - tile0(c), c > 0, c + 1 == K)) return rc;
        }
        h->keValid = true;
        if (int rc = launch_chain(h, sc, CHAIN_FIRST, h->peers.world > 1)) return rc;
    } else if (int rc = launch_chain(h, sc, CHAIN_FIRST)) return rc;
    if (sharded(h) && h->peers.world <= 1) return fail(TGNH_ERR_UNSUPPORTED, "pipelined host path needs the peer-inbox exchange when sharded");
    h->scalePending = false;
    h->keValid = false;
    for (int c = 0; c < K; c++) {
        if (upload) CUDA_TRY(cudaStreamWaitEvent(sc, ev[K + c], 0));
        const int t0 = tile0(c), tn = tile0(c + 1) - t0;
        if (int rc = launch_v2_range(h, sc, V2_A, dv, dx, df, t0, tn, false, false)) return rc;
        if (download) {
            CUDA_TRY(cudaEventRecord(ev[2 * K + c], sc));
            CUDA_TRY(cudaStreamWaitEvent(so, ev[2 * K + c], 0));
            const int p0 = part0(c), p1 = part0(c + 1);
            CUDA_TRY(cudaMemcpyAsync((float4*)posq_host + p0, dx + p0, (size_t)(p1 - p0) * 16, cudaMemcpyDeviceToHost, so));
        }
        if (int rc = launch_v2_range(h, sc, V2_B, dv, nullptr, df, t0, tn, c > 0, c + 1 == K)) return rc;
    }
    if (int rc = launch_chain(h, sc, CHAIN_SECOND, h->peers.world > 1)) return rc;
    h->keValid = true;
    h->scalePending = true;
    if (int rc = flush_scale(h, sc, dv)) return rc;
    if (download) {
        CUDA_TRY(cudaMemcpyAsync(velm_host, dv, (size_t)h->N * 16, cudaMemcpyDeviceToHost, sc));
        if (ke2_host) CUDA_TRY(cudaMemcpyAsync(ke2_host, h->chain.ke2Used, h->T * 8, cudaMemcpyDeviceToHost, sc));
        CUDA_TRY(cudaStreamSynchronize(so));
    }
    return TGNH_OK;
}

extern "C" int tgnh_step_host2(tgnh_handle* h, void* velm_host, void* posq_host, void* posq_correction_host, const void* force_host, int nsteps,
                               int flags, double* ke2_host) {
    if (!h || !velm_host || !posq_host || !force_host) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    if (nsteps < 1) return fail(TGNH_ERR_INVALID_ARGUMENT, "nsteps must be >= 1");
    CUDA_TRY(cudaSetDevice(h->device));
    const bool mixed = h->prec == TGNH_PRECISION_MIXED;
    if (mixed && !posq_correction_host) return fail(TGNH_ERR_INVALID_ARGUMENT, "mixed precision: posq_correction_host is required");
    const size_t fbytes = (size_t)3 * h->paddedN * (h->ffmt == TGNH_FORCE_I64_SOA ? 8 : 4);
    const size_t vb = h->prec ? 32 : 16;                                      // bytes per velm element
    const size_t xb = h->prec == TGNH_PRECISION_DOUBLE ? 32 : 16;             // bytes per posq element
    if (!h->hsVelm) {
        if (!h->hsStream) CUDA_TRY(cudaStreamCreateWithFlags(&h->hsStream, cudaStreamNonBlocking));
        void *v = nullptr, *x = nullptr, *f = nullptr, *c = nullptr;
        if (cudaMalloc(&v, (size_t)h->paddedN * vb) != cudaSuccess || cudaMalloc(&x, (size_t)h->paddedN * xb) != cudaSuccess ||
            cudaMalloc(&f, fbytes) != cudaSuccess || (mixed && cudaMalloc(&c, (size_t)h->paddedN * 16) != cudaSuccess)) {
            cudaFree(v); cudaFree(x); cudaFree(f); cudaFree(c);
            (void)cudaGetLastError();
            return fail(TGNH_ERR_CUDA, "cudaMalloc of the staging buffers failed");
        }
        h->hsVelm = v; h->hsPosq = x; h->hsForce = f; h->hsCorr = c;
        h->hsForceHost = nullptr;
    }
    if (mixed) CUDA_TRY((cudaError_t)(tgnh_set_posq_correction(h, h->hsCorr) == TGNH_OK ? cudaSuccess : cudaErrorInvalidValue));
    // forces: skipped when the caller vouches that this host array has not changed since the call that uploaded it
    const bool uploadForce = !((flags & TGNH_HOST_FORCES_UNCHANGED) && h->hsForceHost == force_host);
    cudaStream_t s = h->hsStream;
    if (h->v2 && h->numTiles2 >= 2 * HS_CHUNKS && h->uniformGroups) {
        if (!h->hsIn) {
            CUDA_TRY(cudaStreamCreateWithFlags(&h->hsIn, cudaStreamNonBlocking));
            CUDA_TRY(cudaStreamCreateWithFlags(&h->hsOut, cudaStreamNonBlocking));
            h->hsEvents.resize(3 * HS_CHUNKS + 2);
            for (cudaEvent_t& e : h->hsEvents) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        }
        for (int i = 0; i < nsteps; i++) {
            const bool first = i == 0, last = i + 1 == nsteps;
            if (!first && !last) {
                if (int rc = tgnh_step(h, s, h->hsVelm, h->hsPosq, h->hsForce, nsteps - 2)) return rc;
                i = nsteps - 2;
                continue;
            }
            if (int rc = pipelined_host_step(h, velm_host, posq_host, force_host, first, first && uploadForce, last, ke2_host)) return rc;
        }
        h->hsForceHost = force_host;
        CUDA_TRY(cudaStreamSynchronize(s));
        return TGNH_OK;
    }
    // every other layout: upload, step, download on one stream
    CUDA_TRY(cudaMemcpyAsync(h->hsVelm, velm_host, (size_t)h->N * vb, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->hsPosq, posq_host, (size_t)h->N * xb, cudaMemcpyHostToDevice, s));
    if (mixed) CUDA_TRY(cudaMemcpyAsync(h->hsCorr, posq_correction_host, (size_t)h->N * 16, cudaMemcpyHostToDevice, s));
    if (uploadForce) CUDA_TRY(cudaMemcpyAsync(h->hsForce, force_host, fbytes, cudaMemcpyHostToDevice, s));
    h->hsForceHost = force_host;
    h->keValid = false;
    if (int rc = tgnh_step(h, s, h->hsVelm, h->hsPosq, h->hsForce, nsteps)) return rc;
    CUDA_TRY(cudaMemcpyAsync(velm_host, h->hsVelm, (size_t)h->N * vb, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(posq_host, h->hsPosq, (size_t)h->N * xb, cudaMemcpyDeviceToHost, s));
    if (mixed) CUDA_TRY(cudaMemcpyAsync(posq_correction_host, h->hsCorr, (size_t)h->N * 16, cudaMemcpyDeviceToHost, s));
    if (ke2_host) CUDA_TRY(cudaMemcpyAsync(ke2_host, h->chain.ke2Used, h->T * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return TGNH_OK;
}

// ------------------------------------------------------------------------------------------------
// thermostat state
// ------------------------------------------------------------------------------------------------
extern "C" int tgnh_num_thermostats(const tgnh_handle* h) { return h ? h->T : 0; }
extern "C" int tgnh_num_nh_chains(const tgnh_handle* h) { return h ? h->M : 0; }
extern "C" int64_t tgnh_launch_count(const tgnh_handle* h) { return h ? h->launches : 0; }
extern "C" int tgnh_exchange_kind(const tgnh_handle* h) {
    if (!h || !h->comm || h->comm->worldSize <= 1) return TGNH_EXCHANGE_NONE;
    return h->peers.world > 1 ? TGNH_EXCHANGE_PEER : TGNH_EXCHANGE_NCCL;
}

static int d2h(tgnh_handle* h, void* stream, double* dst, const double* src, size_t n) {
    if (!h || !dst) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaMemcpyAsync(dst, src, n * 8, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    unsigned int peerError = 0;
    if (h->dInbox) CUDA_TRY(cudaMemcpyAsync(&peerError, &h->dInbox->error, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    if (peerError) return fail(TGNH_ERR_NCCL, "a peer rank did not deliver its kinetic energies within 10 s");
    return TGNH_OK;
}

extern "C" int tgnh_get_exchange_timing(tgnh_handle* h, void* stream, double* us) {
    if (!h || !us) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    us[0] = us[1] = 0.0;
    if (!h->dInbox) return TGNH_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    unsigned long long st[4] = {0, 0, 0, 0};
    CUDA_TRY(cudaMemcpyAsync(st, h->dInbox->stamp, sizeof st, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CUDA_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    us[0] = 1e-3 * (double)(long long)(st[2] - st[1]);
    us[1] = 1e-3 * (double)(long long)(st[1] - st[0]);
    return TGNH_OK;
}

extern "C" int tgnh_get_kinetic_energies(tgnh_handle* h, void* stream, double* ke2) { return d2h(h, stream, ke2, h ? h->chain.ke2Used : nullptr, h ? h->T : 0); }
extern "C" int tgnh_kinetic_energy(tgnh_handle* h, void* stream, double* ke_sum) { return d2h(h, stream, ke_sum, h ? h->chain.keSum : nullptr, 1); }
extern "C" int tgnh_get_vscale(tgnh_handle* h, void* stream, double* vscale) { return d2h(h, stream, vscale, h ? h->chain.vscale : nullptr, h ? h->T : 0); }

extern "C" int tgnh_compute_kinetic_energies(tgnh_handle* h, void* stream, const void* velm, double* ke2) {
    if (int rc = check_ptrs(h, velm, nullptr, nullptr, false, false)) return rc;
    if (!ke2) return fail(TGNH_ERR_INVALID_ARGUMENT, "ke2 is null");
    if (h->scalePending) return fail(TGNH_ERR_INVALID_ARGUMENT, "a deferred velocity scaling is pending; call tgnh_flush first");
    CUDA_TRY(cudaSetDevice(h->device));
    h->keValid = false;
    if (int rc = ensure_ke(h, (cudaStream_t)stream, const_cast<void*>(velm), CHAIN_NONE)) return rc;
    return d2h(h, stream, ke2, h->chain.ke2, h->T);
}

extern "C" int tgnh_get_chain_state(tgnh_handle* h, void* stream, double* eta, double* eta_dot, double* eta_dot_dot) {
    if (!h || !eta || !eta_dot || !eta_dot_dot) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t TM = (size_t)h->T * h->M;
    CUDA_TRY(cudaMemcpyAsync(eta, h->chain.eta, TM * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(eta_dot, h->chain.etaDot, (size_t)h->T * (h->M + 1) * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(eta_dot_dot, h->chain.etaDotDot, TM * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return TGNH_OK;
}

extern "C" int tgnh_set_chain_state(tgnh_handle* h, void* stream, const double* eta, const double* eta_dot, const double* eta_dot_dot) {
    if (!h || !eta || !eta_dot || !eta_dot_dot) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t TM = (size_t)h->T * h->M;
    CUDA_TRY(cudaMemcpyAsync(h->chain.eta, eta, TM * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->chain.etaDot, eta_dot, (size_t)h->T * (h->M + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(h->chain.etaDotDot, eta_dot_dot, TM * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return TGNH_OK;
}

extern "C" int tgnh_set_profiling(tgnh_handle* h, int enabled) {
    if (!h) return fail(TGNH_ERR_INVALID_ARGUMENT, "null handle");
    h->profiling = enabled != 0;
    h->evUsed = 0;
    return TGNH_OK;
}

extern "C" int tgnh_get_profile(tgnh_handle* h, double* ms, int64_t* counts) {
    if (!h || !ms || !counts) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    CUDA_TRY(cudaSetDevice(h->device));
    for (int k = 0; k < 3; k++) { ms[k] = 0.0; counts[k] = 0; }
    for (size_t i = 0; i + 1 < h->evUsed; i += 2) {
        CUDA_TRY(cudaEventSynchronize(h->evPool[i + 1]));
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, h->evPool[i], h->evPool[i + 1]));
        ms[h->evKind[i / 2]] += t;
        counts[h->evKind[i / 2]]++;
    }
    h->evUsed = 0;
    return TGNH_OK;
}

extern "C" int tgnh_get_thermostat_params(const tgnh_handle* h, double* dof, double* nkbt, double* eta_mass) {
    if (!h || !dof || !nkbt || !eta_mass) return fail(TGNH_ERR_INVALID_ARGUMENT, "null argument");
    memcpy(dof, h->dof.data(), h->T * 8);
    memcpy(nkbt, h->nkbt.data(), h->T * 8);
    memcpy(eta_mass, h->etaMass.data(), (size_t)h->T * h->M * 8);
    return TGNH_OK;
}
