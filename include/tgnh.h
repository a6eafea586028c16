/*
 * tgnh.h — C-ABI of the B200-native DrudeTGNHIntegrator step (libtgnh.so, sm_100a, AOT).
 *
 * This is the drop-in boundary for ONE hot path of scychon/openmm_drudeNose: everything
 * CudaIntegrateDrudeTGNHStepKernel does on the device and on the host critical path
 * (reference paths relative to /root/reference):
 *     platforms/cuda/src/CudaDrudeTGNHKernels.cpp:75-282   initialize   -> tgnh_create
 *     platforms/cuda/src/CudaDrudeTGNHKernels.cpp:284-408  execute      -> tgnh_half1 / tgnh_half2 / tgnh_step
 *     platforms/cuda/src/CudaDrudeTGNHKernels.cpp:433-652  propagateNHChain (host fp64 chain, 2 blocking
 *                                                          D2H + 2 H2D per step) -> device-resident chain
 *     platforms/cuda/src/CudaDrudeTGNHKernels.cpp:654-661  computeKineticEnergy -> tgnh_kinetic_energy
 *     platforms/cuda/src/kernels/drudeTGNH.cu:82-574       the 8 live runtime-compiled kernels
 * It is called from the C++ OpenMM KernelImpl (plugin/), one host thread per handle (OpenMM
 * contexts are not thread-safe either).  Plain pointers and sizes only; no exceptions, STL or
 * torch types cross it.  Every function returns TGNH_OK or an error code; the message is
 * available from tgnh_last_error() (thread-local).
 *
 * Buffers handed to the step functions are DEVICE pointers owned by the caller, in the layouts
 * OpenMM's CudaContext uses; shown for single precision, see TGNH_PRECISION_MIXED below:
 *     velm   float4[paddedN]   (vx, vy, vz, 1/m)   cu.getVelm();  w == 0 marks an immovable particle
 *     posq   float4[paddedN]   (x, y, z, q)        cu.getPosq();  w is preserved
 *     force  SoA [3][paddedN]  force[i + k*paddedN]                cu.getForce()
 *            TGNH_FORCE_I64_SOA: long long fixed point, scale 2^32 (CudaDrudeTGNHKernels.cpp:295)
 *            TGNH_FORCE_F32_SOA: float (synthetic-force bench; same indexing)
 * Units: nm, ps, amu, kJ/mol, K.
 */
#ifndef TGNH_H_
#define TGNH_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* OpenMM's BOLTZ (RGAS/1000, SimTKOpenMMRealType.h; external constant, pinned here) */
#define TGNH_BOLTZ (1.380649e-23 * 6.02214076e23 / 1000.0)

enum {
    TGNH_OK = 0,
    TGNH_ERR_INVALID_ARGUMENT = 1,   /* bad sizes / indices / pointers */
    TGNH_ERR_TEMP_GROUP = 2,         /* Drude pair or constraint spans two temperature groups (CudaDrudeTGNHKernels.cpp:146,193) */
    TGNH_ERR_UNSUPPORTED = 3,        /* layout the AOT kernels do not cover (see message) */
    TGNH_ERR_CUDA = 4,
    TGNH_ERR_NCCL = 5,
    TGNH_ERR_NO_DEVICE = 6
};

enum { TGNH_FORCE_F32_SOA = 0, TGNH_FORCE_I64_SOA = 1 };

/* OpenMM CUDA precision modes (cu.getUseMixedPrecision() / cu.getUseDoublePrecision()):
 *   SINGLE  velm float4,  posq float4,                         posDelta float4;  fp32 arithmetic, fp64 energy sums + chain
 *   MIXED   velm double4, posq float4 + posqCorrection float4, posDelta double4; fp64 arithmetic ("mixed" = double in
 *           drudeTGNH.cu); register cu.getPosqCorrection() once with tgnh_set_posq_correction
 *   DOUBLE  velm double4, posq double4,                        posDelta double4; fp64 arithmetic */
enum { TGNH_PRECISION_SINGLE = 0, TGNH_PRECISION_MIXED = 1, TGNH_PRECISION_DOUBLE = 2 };

/* flags for tgnh_half2 */
enum {
    TGNH_HALF2_DEFAULT = 0,
    TGNH_HALF2_KICK_ONLY = 2,        /* constrained systems: only the half kick; the caller applies OpenMM's velocity
                                        constraints and then calls tgnh_thermostat */
    TGNH_HALF2_DEFER_SCALE = 1       /* leave the second thermostat half-step's velocity scaling pending; it is folded
                                        into the next tgnh_half1 (or applied by tgnh_flush).  Only legal when nothing
                                        reads or writes velm in between. */
};

typedef struct tgnh_handle tgnh_handle;
typedef struct tgnh_comm tgnh_comm;

typedef struct {
    /* sizes */
    int32_t num_particles;           /* N  (of THIS shard when comm != NULL) */
    int32_t padded_num_particles;    /* stride of the SoA force components; multiple of 4, >= N (OpenMM: multiple of 32) */
    int32_t num_pairs;               /* P  DrudeForce entries */
    int32_t num_residues;            /* R  molecules (ContextImpl::getMolecules) */
    int32_t num_temp_groups;         /* G  DrudeTGNHIntegrator::getNumTempGroups */
    int32_t num_constraints;         /* only used for DOF bookkeeping (constraints stay in OpenMM) */
    /* integrator parameters (openmmapi/include/openmm/DrudeTGNHIntegrator.h:71) */
    int32_t num_nh_chains;
    int32_t drude_steps_per_real_step;
    int32_t use_drude_nh_chains;
    int32_t use_com_temp_group;
    int32_t has_cm_motion_remover;   /* System contains a CMMotionRemover (CudaDrudeTGNHKernels.cpp:204-212) */
    int32_t force_format;            /* TGNH_FORCE_* */
    int32_t device;                  /* CUDA device ordinal; -1 = current device */
    int32_t precision;               /* TGNH_PRECISION_* */
    double temperature;
    double coupling_time;
    double drude_temperature;
    double drude_coupling_time;
    double step_size;
    double max_drude_distance;       /* 0 disables the hard wall */
    /* host tables, copied during tgnh_create */
    const double* masses;            /* [N] System::getParticleMass */
    const int32_t* pair_drude;       /* [P] DrudeForce::getParticleParameters particle  */
    const int32_t* pair_parent;      /* [P] DrudeForce::getParticleParameters particle1 */
    const int32_t* particle_temp_group; /* [N] */
    const int32_t* particle_res_id;  /* [N]; with use_com_temp_group each residue must be a contiguous index range (the reference assumes it,
                                        drudeTGNH.cu:86-101) that contains its Drude pairs; without it the reference never reads the residues
                                        (:87-108) and neither does this library: any ids are accepted */
    const int32_t* constraint_p;     /* [C] */
    const int32_t* constraint_p1;    /* [C] */
    /* sharding: NULL = single GPU.  With a communicator the tables above describe this rank's molecule-aligned
       particle range; DOF / thermostat masses are summed over ranks at create time and the per-group
       kinetic-energy vector is all-reduced (NCCL, double[G+2]) before every chain update. */
    tgnh_comm* comm;
} tgnh_params;

/* ---- lifetime ------------------------------------------------------------------------------ */
int tgnh_create(const tgnh_params* params, tgnh_handle** out);
void tgnh_destroy(tgnh_handle* h);
const char* tgnh_last_error(void);
/* "sm_100a" build info, for diagnostics */
const char* tgnh_build_info(void);

/* ---- the step (streams are cudaStream_t passed as void*) ---------------------------------- */
/* First half: thermostat half-step (KE -> chain -> scale), half kick, drift, hard wall.
 * = CudaDrudeTGNHKernels.cpp:336-376 without the OpenMM constraint call at :363. */
int tgnh_half1(tgnh_handle* h, void* stream, void* velm, void* posq, const void* force);
/* The first half split around OpenMM's position constraints (systems with SETTLE / SHAKE / CCMA):
 *   tgnh_half1_kick   thermostat half-step + half kick; writes pos_delta = (dt*v, 0) float4[paddedN], the layout of
 *                     integration.getPosDelta() in single precision            (= :336-360)
 *   ... caller: integration.applyConstraints(tol) on pos_delta                 (= :363)
 *   tgnh_half1_drift  x += pos_delta, v = pos_delta / dt, hard wall            (= :366-376) */
int tgnh_half1_kick(tgnh_handle* h, void* stream, void* velm, const void* force, void* pos_delta);
int tgnh_half1_drift(tgnh_handle* h, void* stream, void* velm, void* posq, void* pos_delta);
/* Thermostat half-step on the velocities as they are (after tgnh_half2(KICK_ONLY) + the caller's velocity constraints):
 * kinetic energies, chain update, scaling (= :394-402).  flags: TGNH_HALF2_DEFER_SCALE or 0. */
int tgnh_thermostat(tgnh_handle* h, void* stream, void* velm, int flags);
/* Second half: half kick with the new forces, thermostat half-step.  = :384-402 without :391. */
int tgnh_half2(tgnh_handle* h, void* stream, void* velm, const void* force, int flags);
/* Apply a pending (deferred) velocity scaling so that velm is what the reference would hold. */
int tgnh_flush(tgnh_handle* h, void* stream, void* velm);
/* nsteps full steps with the SAME force array in both halves (integrator-only path with fixed
 * synthetic forces).  Internally defers/folds the scaling between steps; velm is consistent on return. */
int tgnh_step(tgnh_handle* h, void* stream, void* velm, void* posq, const void* force, int nsteps);
/* Host-buffer convenience (end-to-end path): copies velm/posq/force from pinned or pageable HOST memory,
 * runs nsteps, copies velm/posq back and the 2*KE vector into ke2_host ([G+2], may be NULL). Blocking.
 * Single and double layouts (mixed also needs the posqCorrection array: use the device-buffer calls). */
int tgnh_step_host(tgnh_handle* h, void* velm_host, void* posq_host, const void* force_host, int nsteps, double* ke2_host);
/* The same with (a) the mixed layout's posqCorrection array (NULL otherwise), (b) flags:
 *   TGNH_HOST_FORCES_UNCHANGED  the caller vouches that force_host holds what it held in the previous call with this pointer
 *                               (fixed synthetic forces): the upload of the forces is skipped.
 * Single-precision systems that run through the warp-chunk kernels are pipelined: the state is cut into 8 particle ranges; the
 * kinetic energies of a range are reduced as soon as its velocities have landed, its two halves run as soon as its positions and
 * forces have landed, and its new positions travel back over the second copy engine while later ranges are still being uploaded. */
enum { TGNH_HOST_FORCES_UNCHANGED = 1 };
int tgnh_step_host2(tgnh_handle* h, void* velm_host, void* posq_host, void* posq_correction_host, const void* force_host, int nsteps, int flags,
                    double* ke2_host);
/* Mixed precision: the float4[paddedN] residual array of the positions (cu.getPosqCorrection()). */
int tgnh_set_posq_correction(tgnh_handle* h, void* posq_correction);
/* The velocities were changed behind the integrator's back (DrudeTGNHIntegrator::stateChanged,
 * openmmapi/src/DrudeTGNHIntegrator.cpp:166-170): cached kinetic energies are dropped. */
int tgnh_invalidate(tgnh_handle* h);

/* ---- thermostat state (blocking; they synchronise `stream`) -------------------------------- */
/* sizes: T = G+2 thermostats (G relative groups, COM group at G, Drude group at G+1) */
int tgnh_num_thermostats(const tgnh_handle* h);
/* M, the Nose-Hoover chain length the handle was created with */
int tgnh_num_nh_chains(const tgnh_handle* h);
/* 2*KE per thermostat as consumed by the most recent chain update (kineticEnergiesVec, :490) */
int tgnh_get_kinetic_energies(tgnh_handle* h, void* stream, double* ke2 /*[T]*/);
/* 0.5 * sum(2KE) cached by the last chain update (KESum, :493-497 / :654-658) */
int tgnh_kinetic_energy(tgnh_handle* h, void* stream, double* ke_sum);
/* 2*KE per thermostat of the velocities as they are now (runs the reduction kernel) */
int tgnh_compute_kinetic_energies(tgnh_handle* h, void* stream, const void* velm, double* ke2 /*[T]*/);
/* eta [T*M], eta_dot [T*(M+1)] (last column is the permanent 0), eta_dot_dot [T*M]  (CudaDrudeTGNHKernels.h:90-93) */
int tgnh_get_chain_state(tgnh_handle* h, void* stream, double* eta, double* eta_dot, double* eta_dot_dot);
int tgnh_set_chain_state(tgnh_handle* h, void* stream, const double* eta, const double* eta_dot, const double* eta_dot_dot);
/* velocity scale factors produced by the most recent chain update (vscaleFactorsVec) */
int tgnh_get_vscale(tgnh_handle* h, void* stream, double* vscale /*[T]*/);
/* dof[T] (dof - COM share), NkT[T], eta_mass[T*M]  (tempGroupDof/tempGroupNkbT/etaMass, :215-235) */
int tgnh_get_thermostat_params(const tgnh_handle* h, double* dof, double* nkbt, double* eta_mass);
/* The host-side plan tgnh_create would build for these parameters, without touching a device: every check of the tables
 * (contiguous residues, Drude pairs inside one residue and one temperature group, constraint partners in one group: the
 * reference's two exceptions, CudaDrudeTGNHKernels.cpp:146,193) and the tiling.  tile_start (may be NULL) receives
 * num_tiles + 1 particle indices: tile t covers [tile_start[t], tile_start[t+1]), at most 512 particles, never separating
 * a Drude pair nor a residue of up to 128 particles. */
int tgnh_plan_tiles(const tgnh_params* p, int32_t* tile_start, int32_t capacity, int32_t* num_tiles, int32_t* num_big_residues,
                    int32_t* residue_uniform);
/* The per-particle descriptor words tgnh_create uploads (temperature group, role, offsets inside the residue, offset to the
 * pair partner), host-only.  Two particles are interchangeable for this library exactly when their words and masses are
 * equal: what a CudaForceInfo::areParticlesIdentical must answer so that OpenMM's atom reordering (cu.reorderAtoms) only
 * swaps molecules whose tables agree ("Atom reordering"). */
int tgnh_plan_descriptors(const tgnh_params* p, uint32_t* desc_out /*[num_particles]*/);
/* The plan of the warp-chunk kernels (csrc/tgnh_v2.cuh), host-only: residue-aligned chunks of at most 32 consecutive particles
 * (chunk_start receives num_chunks + 1 particle indices, num_chunks a multiple of tgnh_chunks_per_tile(), the tail
 * padded with empty chunks), one species byte per particle, and the species table (256 rows of 8 floats: m_hi, m_lo,
 * 1/M_residue hi, meta bits, mu_hi, mu_lo, 1/M_residue lo, m_partner/(m + m_partner); the row after the last species = "no particle", all zero; meta: [4:0] temperature
 * group, [6:5] role, [12:7] signed offset to the pair partner, [17:13] / [22:18] offsets to the first / last particle of the residue).
 * TGNH_ERR_UNSUPPORTED when the system does not qualify (mixed / double layout, a residue of more than 32 particles, more than
 * 255 species): such systems run through the first-generation kernels (tgnh_plan_tiles).  Any output pointer may be NULL. */
int tgnh_plan_chunks(const tgnh_params* p, int32_t* chunk_start, int32_t capacity, int32_t* num_chunks, uint8_t* species_out /*[N]*/,
                     float* table_out /*[256*8]*/, int32_t* num_species, int32_t* max_residue);
/* chunks (= consumer warps) per tile of the warp-chunk kernels: tgnh_plan_chunks pads its chunk table to a multiple of it */
int tgnh_chunks_per_tile(void);
/* 2 when the handle's two halves run through the warp-chunk kernels, 1 otherwise (environment TGNH_V2=0 forces 1) */
int tgnh_kernel_generation(const tgnh_handle* h);
/* 1 when tgnh_step(n) leaves the second half kick of every step but the last to the next first half (the second half then only
 * reduces the kinetic energies: 28 instead of 44 bytes per particle; bit-identical results).  Warp-chunk kernels only, not for
 * systems small enough to run the chain in the reducing launch; environment TGNH_LAZY_KICK=0 at tgnh_create switches it off. */
int tgnh_lazy_second_kick(const tgnh_handle* h);
/* k > 0 when the handle's reducing launches (second half, kinetic-energy reduction) run in the residue-per-lane form: every
 * residue of the system has k particles (2..8) and lies in one temperature group, the COM temperature group is on, warp-chunk
 * kernels.  A lane then owns a whole residue (no shuffles, ~3 times fewer instructions per particle).  0 otherwise; environment
 * TGNH_RPL=0 at tgnh_create switches it off, TGNH_RPL=1 restricts it to the launches that store no velocities. */
int tgnh_residue_per_lane(const tgnh_handle* h);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t tgnh_launch_count(const tgnh_handle* h);
/* Per-launch device timing: while enabled every streaming launch is bracketed by CUDA events on its stream.
 * tgnh_get_profile synchronises, returns accumulated milliseconds and launch counts for
 * [0] first-half kernel, [1] second-half kernel, [2] reduce/flush kernel, and resets the accumulators. */
int tgnh_set_profiling(tgnh_handle* h, int enabled);
int tgnh_get_profile(tgnh_handle* h, double* ms /*[3]*/, int64_t* counts /*[3]*/);

/* ---- sharding over the GPUs of one node ----------------------------------------------------------
 * Each rank owns a contiguous, molecule-aligned particle range; the thermostats see the whole system.  The only data
 * exchanged per step is the double[T] vector of kinetic-energy partial sums.  When the ranks' GPUs can map each other's
 * memory (CUDA IPC over NVLink / NVSwitch, one node) it travels through peer-mapped inboxes written by the reducing
 * kernel's last CTA and read by the chain kernel — no collective launch; otherwise through ncclAllReduce.  The NCCL
 * communicator also carries the one-time set-up collectives of tgnh_create.  All calls on a sharded handle that launch
 * a kinetic-energy reduction are collective: every rank must make the same sequence of calls. */
#define TGNH_UNIQUE_ID_BYTES 128
int tgnh_comm_get_unique_id(void* id_out /*[128]*/);
int tgnh_comm_create(const void* unique_id, int world_size, int rank, int device, tgnh_comm** out);
void tgnh_comm_destroy(tgnh_comm* c);
enum { TGNH_EXCHANGE_NONE = 0, TGNH_EXCHANGE_NCCL = 1, TGNH_EXCHANGE_PEER = 2 };
/* how this handle exchanges the kinetic-energy partial sums (TGNH_EXCHANGE_*); environment TGNH_P2P=0 forces NCCL */
int tgnh_exchange_kind(const tgnh_handle* h);

/* Diagnostics of the peer-inbox exchange (device %globaltimer stamps of the most recent reduction on this rank), microseconds:
 *   us[0]  time the chain launch spent waiting for all ranks' partial sums (on the rank that finishes last this is the pure
 *          exchange latency, on the others it also contains the skew between the ranks)
 *   us[1]  time between this rank's publish (end of its reducing launch) and the start of its wait (launch hand-over)
 * Both are 0 for handles that do not use the inboxes. */
int tgnh_get_exchange_timing(tgnh_handle* h, void* stream, double* us /*[2]*/);

#ifdef __cplusplus
}
#endif
#endif /* TGNH_H_ */
