/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.
 *
 * CPU (fp64, serial) restatement of the reference's TGNH integrator step, used only as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs.  Nothing under openmm_drudenose_b200/ may link, import or call it.
 *
 * Two layers (SURVEY.md finding 2, paths relative to /root/reference):
 *   TGNH_ORACLE_TG   restates the CUDA platform's host+device algorithm in fp64:
 *                    platforms/cuda/src/CudaDrudeTGNHKernels.cpp:75-235 (init), :433-652 (chain)
 *                    platforms/cuda/src/kernels/drudeTGNH.cu:82-301,307-365,435-574
 *                    (temperature groups + molecular-COM thermostat live ONLY there).
 *   TGNH_ORACLE_REF  restates platforms/reference/src/ReferenceDrudeTGNHKernels.cpp
 *                    :104-219 (init), :221-415 (step), :426-546 (chain), :548-584 (kick)
 *                    (dual Nose-Hoover only; no temperature groups / COM group).
 *
 * Pinning status: the reference ships NO golden vectors.  This restatement is
 * pinned (a) against the reference platform's own sources compiled unmodified against a
 * self-written OpenMM API shim (oracle/_ref, see oracle/Makefile) where that build exists,
 * (b) by cross-checking the two layers on their overlap domain, and (c) by the reference
 * tests' statistical invariants (testSinglePair).  Constraints, virtual sites and force
 * evaluation are OpenMM-external and are not restated (integrator-only path).
 */
#ifndef TGNH_ORACLE_H_
#define TGNH_ORACLE_H_

#ifdef __cplusplus
extern "C" {
#endif

/* OpenMM's BOLTZ (SimTKOpenMMRealType.h, external): RGAS/1000 with the 2019 SI constants.
 * Both sides of every parity test use this literal. */
#define TGNH_ORACLE_BOLTZ (1.380649e-23 * 6.02214076e23 / 1000.0)

enum { TGNH_ORACLE_TG = 0, TGNH_ORACLE_REF = 1 };
enum { TGNH_ORACLE_FORCE_FIXED = 0, TGNH_ORACLE_FORCE_HARMONIC = 1 };

typedef struct {
    int num_particles;
    int num_pairs;
    int num_residues;
    int num_temp_groups;
    int num_nh_chains;
    int drude_steps_per_real_step;
    int use_drude_nh_chains;
    int use_com_temp_group;
    int has_cm_motion_remover;
    int num_constraints;
    double temperature;
    double coupling_time;
    double drude_temperature;
    double drude_coupling_time;
    double step_size;
    double max_drude_distance;
    const double* masses;            /* [N]  System::getParticleMass */
    const int* pair_drude;           /* [P]  DrudeForce particle  (pairParticles.x) */
    const int* pair_parent;          /* [P]  DrudeForce particle1 (pairParticles.y) */
    const int* particle_temp_group;  /* [N] */
    const int* particle_res_id;      /* [N] */
    const int* constraint_p;         /* [C] only used for DOF bookkeeping */
    const int* constraint_p1;        /* [C] */
} tgnh_oracle_params;

typedef struct tgnh_oracle tgnh_oracle;

/* returns 0 on success; on failure a message is available from tgnh_oracle_last_error() */
int tgnh_oracle_create(const tgnh_oracle_params* p, int which, tgnh_oracle** out);
void tgnh_oracle_destroy(tgnh_oracle* o);
const char* tgnh_oracle_last_error(void);

/* O(N) loops run on `n` OpenMP threads (default 1 = the reference platform's serial order).  Only the
 * CPU-baseline timing legs of bench.py raise it; parity tests keep 1. */
void tgnh_oracle_set_threads(int n);
int tgnh_oracle_get_threads(void);

/* Individual phases (pos/vel/force are [N][3] doubles, AoS like vector<RealVec>). */
int tgnh_oracle_propagate_nh_chain(tgnh_oracle* o, double* vel);   /* KE -> chain -> scale velocities */
int tgnh_oracle_half_kick(tgnh_oracle* o, double* vel, const double* force);
int tgnh_oracle_drift(tgnh_oracle* o, double* pos, double* vel);
int tgnh_oracle_hard_wall(tgnh_oracle* o, double* pos, double* vel);

/* Full steps.  force_model FIXED: `force` is used unchanged in both half kicks.
 * HARMONIC: after the drift, force = ext_force (may be NULL) + Drude springs -k_i (x_d - x_p)
 * (the isotropic DrudeForce term used as synthetic force by the reference's testSinglePair).
 * `force` must hold valid forces for the current positions on entry (the reference assumes this:
 * openmmapi/src/DrudeTGNHIntegrator.cpp:166-170) and holds them again on exit. */
int tgnh_oracle_step(tgnh_oracle* o, double* pos, double* vel, double* force, int nsteps,
                     int force_model, const double* ext_force, const double* k_spring);

/* State inspection. sizes: ke2 [G+2] (the 2*KE sums the last chain call consumed),
 * vscale [G+2], eta/eta_dot_dot [(G+2)*M], eta_dot [(G+2)*(M+1)].  For TGNH_ORACLE_REF,
 * G is reported as 1 with group 0 = real, 1 = (absent COM group, zeros), 2 = Drude. */
int tgnh_oracle_num_thermostats(const tgnh_oracle* o);
void tgnh_oracle_get_ke2(const tgnh_oracle* o, double* ke2);
void tgnh_oracle_get_vscale(const tgnh_oracle* o, double* vscale);
void tgnh_oracle_get_chain_state(const tgnh_oracle* o, double* eta, double* eta_dot, double* eta_dot_dot);
void tgnh_oracle_set_chain_state(tgnh_oracle* o, const double* eta, const double* eta_dot, const double* eta_dot_dot);
void tgnh_oracle_get_thermostat_params(const tgnh_oracle* o, double* dof, double* nkbt, double* eta_mass /*[(G+2)*M]*/);
double tgnh_oracle_get_ke_sum(const tgnh_oracle* o);
/* 2*KE sums of the CURRENT velocities without touching the chain. */
int tgnh_oracle_compute_ke2(tgnh_oracle* o, const double* vel, double* ke2);

#ifdef __cplusplus
}
#endif
#endif
