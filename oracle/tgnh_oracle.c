/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See tgnh_oracle.h for scope and pinning status.
 *
 * Plain C, fp64, serial (the reference platform is single-threaded).  Every function cites the
 * reference file:line it follows (paths relative to /root/reference).  The arithmetic keeps the
 * reference's operation order so that the double-precision results can be compared tightly with
 * oracle/_ref (the reference platform's own sources compiled against an API shim).
 */
#include "tgnh_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Threads used by the O(N) loops.  1 (default) = the reference platform's serial order; > 1 is only
 * used by bench.py's CPU-baseline legs ("all host cores"), where reduction order may differ. */
static int g_threads = 1;
void tgnh_oracle_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int tgnh_oracle_get_threads(void) { return g_threads; }
#define PAR_FOR _Pragma("omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1)")

static char g_err[512] = "";
const char* tgnh_oracle_last_error(void) { return g_err; }
#define FAIL(...) do { snprintf(g_err, sizeof g_err, __VA_ARGS__); return 1; } while (0)

struct tgnh_oracle {
    int which;
    int N, P, R, G, M, S;
    int useDrudeNH, useCOM;
    double kT, kTD, dt, rmax;
    double *mass, *invMass;
    int *pairD, *pairP, *tg, *resid, *normal;
    int numNormal;
    /* ---- TG layer (CudaDrudeTGNHKernels.h:88-101) ---- */
    int *resCount, *resFirst;               /* particlesInResidues (count, first) */
    double *tgDof, *tgNkbT;                 /* [G+2] (dof stored as dof - redMass, as double) */
    double *etaMass, *eta, *etaDot, *etaDotDot; /* [G+2][M], [G+2][M], [G+2][M+1], [G+2][M] */
    double *vscale, *ke2;                   /* [G+2] */
    double *comVel, *normVel;               /* [R][4], [N][3] scratch */
    double KESum;
    /* ---- REF layer (ReferenceDrudeTGNHKernels.h:94-112) ---- */
    double *pairInvTotalMass, *pairInvReducedMass;
    int numTempGroupRef, iNumNHChains, idxMaxNHChains, nFlat;
    double realNkbT, drudeNkbT;
    double *fEtaMass, *fEta, *fEtaDot, *fEtaDotDot;  /* flat chain arrays */
    double scaleReal, scaleDrude, realKE2, drudeKE2;
};

static void* xcalloc(size_t n, size_t s) { void* p = calloc(n ? n : 1, s); if (!p) abort(); return p; }

/* ------------------------------------------------------------------------------------------------
 * init, TG layer: platforms/cuda/src/CudaDrudeTGNHKernels.cpp:75-235
 * residue masses: openmmapi/src/DrudeTGNHIntegrator.cpp:136-153
 * ---------------------------------------------------------------------------------------------- */
static int init_tg(tgnh_oracle* o, const tgnh_oracle_params* p) {
    const int N = o->N, G = o->G, M = o->M, T = G + 2;
    double* resMass = xcalloc(o->R, sizeof(double));
    for (int i = 0; i < N; i++) {
        if (o->resid[i] < 0 || o->resid[i] >= o->R) { free(resMass); FAIL("particle %d has residue id %d outside [0,%d)", i, o->resid[i], o->R); }
        resMass[o->resid[i]] += o->mass[i];                       /* DrudeTGNHIntegrator.cpp:147-148 */
    }
    o->resCount = xcalloc(o->R, sizeof(int));
    o->resFirst = xcalloc(o->R, sizeof(int));
    for (int r = 0; r < o->R; r++) o->resFirst[r] = -1;            /* :88-89 */
    int* dofInt = xcalloc(T, sizeof(int));
    double* redMass = xcalloc(G + 1, sizeof(double));
    int prevResId = -1, drudeDof = 0;
    for (int i = 0; i < N; i++) {                                  /* :114-134 */
        int tg = o->tg[i];
        if (tg < 0 || tg >= G) { free(resMass); free(dofInt); free(redMass); FAIL("particle %d has temperature group %d outside [0,%d)", i, tg, G); }
        int resid = o->resid[i];
        o->resCount[resid] += 1;
        if (prevResId != resid) { o->resFirst[resid] = i; prevResId = resid; }
        double mass = o->mass[i];
        double resInvMass = 1.0 / resMass[resid];                  /* DrudeTGNHIntegrator.cpp:152-153 */
        if (mass != 0.0) {
            dofInt[tg] += 3;
            if (o->useCOM) redMass[tg] += 3 * mass * resInvMass;
        }
    }
    for (int i = 0; i < o->P; i++) {                               /* :135-150 */
        int tg = o->tg[o->pairD[i]], tg1 = o->tg[o->pairP[i]];
        if (tg != tg1) { free(resMass); free(dofInt); free(redMass); FAIL("Temperature group for drude particle must be the same as the parent particle"); }
        dofInt[tg] -= 3;
        drudeDof += 3;
    }
    for (int i = 0; i < p->num_constraints; i++) {                 /* :186-196 */
        int tg = o->tg[p->constraint_p[i]], tg1 = o->tg[p->constraint_p1[i]];
        if (tg != tg1) { free(resMass); free(dofInt); free(redMass); FAIL("Temperature group of constrained particles must be the same"); }
        dofInt[tg] -= 1;
    }
    if (o->useCOM) dofInt[G] = 3 * o->R;                            /* :197-199 */
    dofInt[G + 1] = drudeDof;                                      /* :201 */
    if (o->useCOM && p->has_cm_motion_remover) dofInt[G] -= 3;     /* :204-212 */

    o->tgDof = xcalloc(T, sizeof(double));
    o->tgNkbT = xcalloc(T, sizeof(double));
    o->etaMass = xcalloc((size_t)T * M, sizeof(double));
    o->eta = xcalloc((size_t)T * M, sizeof(double));
    o->etaDot = xcalloc((size_t)T * (M + 1), sizeof(double));
    o->etaDotDot = xcalloc((size_t)T * M, sizeof(double));
    o->vscale = xcalloc(T, sizeof(double));
    o->ke2 = xcalloc(T, sizeof(double));
    for (int i = 0; i < T; i++) o->vscale[i] = 1.0;
    const double drudekbT = o->kTD, realkbT = o->kT;
    const double drudeNkbT = drudeDof * drudekbT;                  /* :215 */
    const double drudeEtaMassUnit = drudekbT * pow(p->drude_coupling_time, 2);
    const double realEtaMassUnit = realkbT * pow(p->coupling_time, 2);
    for (int i = 0; i < G + 1; i++) {                              /* :218-225 */
        o->tgDof[i] = dofInt[i] - redMass[i];
        o->tgNkbT[i] = (dofInt[i] - redMass[i]) * realkbT;
        o->etaMass[i * M + 0] = (dofInt[i] - redMass[i]) * realEtaMassUnit;
        for (int ich = 1; ich < M; ich++) {
            o->etaMass[i * M + ich] = realEtaMassUnit;
            double ed = o->etaDot[i * (M + 1) + ich - 1];
            o->etaDotDot[i * M + ich] = (o->etaMass[i * M + ich - 1] * ed * ed - realkbT) / o->etaMass[i * M + ich];
        }
    }
    const int itg = G + 1;                                         /* :227-235 */
    o->tgDof[itg] = drudeDof;
    o->tgNkbT[itg] = drudeNkbT;
    o->etaMass[itg * M + 0] = drudeDof * drudeEtaMassUnit;
    for (int ich = 1; ich < M; ich++) {
        o->etaMass[itg * M + ich] = drudeEtaMassUnit;
        if (o->useDrudeNH) {
            double ed = o->etaDot[itg * (M + 1) + ich - 1];
            o->etaDotDot[itg * M + ich] = (o->etaMass[itg * M + ich - 1] * ed * ed - drudekbT) / o->etaMass[itg * M + ich];
        }
    }
    o->comVel = xcalloc((size_t)o->R * 4, sizeof(double));
    o->normVel = xcalloc((size_t)N * 3, sizeof(double));
    free(resMass); free(dofInt); free(redMass);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * init, REF layer: platforms/reference/src/ReferenceDrudeTGNHKernels.cpp:104-219
 * (D2: the out-of-bounds debug print at :218 is not restated.)
 * ---------------------------------------------------------------------------------------------- */
static int init_ref(tgnh_oracle* o, const tgnh_oracle_params* p) {
    int realDof = 0, drudeDof = 0;
    for (int i = 0; i < o->N; i++) realDof += (o->mass[i] == 0.0 ? 0 : 3);  /* :119 */
    o->pairInvTotalMass = xcalloc(o->P, sizeof(double));
    o->pairInvReducedMass = xcalloc(o->P, sizeof(double));
    for (int i = 0; i < o->P; i++) {                                         /* :121-135 */
        double m1 = o->mass[o->pairD[i]], m2 = o->mass[o->pairP[i]];
        o->pairInvTotalMass[i] = 1.0 / (m1 + m2);
        o->pairInvReducedMass[i] = (m1 + m2) / (m1 * m2);
        realDof -= 3;
        drudeDof += 3;
    }
    const int M = o->M;
    o->numTempGroupRef = o->useDrudeNH ? 2 : 1;                              /* :139-144 */
    o->iNumNHChains = M;
    if (o->useDrudeNH) {                                                     /* :146-154 */
        o->idxMaxNHChains = M * o->numTempGroupRef - 1;
        o->iNumNHChains = M * o->numTempGroupRef;
    } else {
        o->idxMaxNHChains = M * o->numTempGroupRef;
        o->iNumNHChains = M * o->numTempGroupRef + 1;
    }
    realDof -= p->num_constraints;                                           /* :157 */
    if (p->has_cm_motion_remover) realDof -= 3;                              /* :158-165 */
    const double realkbT = o->kT, drudekbT = o->kTD;
    o->realNkbT = realDof * realkbT;                                         /* :168-171 */
    o->drudeNkbT = drudeDof * drudekbT;
    /* flat arrays; capacity generous enough for the reference's i+2 reads in the second sweep */
    o->nFlat = 2 * M + 4;
    o->fEtaMass = xcalloc(o->nFlat, sizeof(double));
    o->fEta = xcalloc(o->nFlat, sizeof(double));
    o->fEtaDot = xcalloc(o->nFlat + 2, sizeof(double));
    o->fEtaDotDot = xcalloc(o->nFlat, sizeof(double));
    int n = 0;
    o->fEtaMass[n++] = o->realNkbT * pow(p->coupling_time, 2);
    o->fEtaMass[n++] = o->drudeNkbT * pow(p->drude_coupling_time, 2);
    const int ntg = o->numTempGroupRef;
    if (o->useDrudeNH) {                                                     /* :192-205 */
        for (int ich = 1; ich < M; ich++) {
            o->fEtaMass[n++] = realkbT * pow(p->coupling_time, 2);
            o->fEtaMass[n++] = drudekbT * pow(p->drude_coupling_time, 2);
            double e0 = o->fEtaDot[(ich - 1) * ntg], e1 = o->fEtaDot[(ich - 1) * ntg + 1];
            o->fEtaDotDot[ich * ntg] = (o->fEtaMass[(ich - 1) * ntg] * e0 * e0 - realkbT) / o->fEtaMass[ich * ntg];
            o->fEtaDotDot[ich * ntg + 1] = (o->fEtaMass[(ich - 1) * ntg + 1] * e1 * e1 - drudekbT) / o->fEtaMass[ich * ntg + 1];
        }
    } else {                                                                 /* :206-214 */
        for (int ich = 1; ich < M; ich++) {
            o->fEtaMass[n++] = realkbT * pow(p->coupling_time, 2);
            double e = o->fEtaDot[(ich - 1) * ntg + 1];
            o->fEtaDotDot[ich * ntg + 1] = (o->fEtaMass[(ich - 1) * ntg + 1] * e * e - realkbT) / o->fEtaMass[ich * ntg + 1];
        }
    }
    o->tgDof = xcalloc(3, sizeof(double));
    o->tgDof[0] = realDof; o->tgDof[2] = drudeDof;
    o->scaleReal = o->scaleDrude = 1.0;
    return 0;
}

int tgnh_oracle_create(const tgnh_oracle_params* p, int which, tgnh_oracle** out) {
    if (!p || !out) FAIL("null argument");
    if (p->num_nh_chains < 1) FAIL("numNHChains must be >= 1");
    if (p->drude_steps_per_real_step < 1) FAIL("drudeStepsPerRealStep must be >= 1");
    tgnh_oracle* o = xcalloc(1, sizeof *o);
    o->which = which;
    o->N = p->num_particles; o->P = p->num_pairs; o->R = p->num_residues;
    o->G = (which == TGNH_ORACLE_REF) ? 1 : p->num_temp_groups;
    o->M = p->num_nh_chains; o->S = p->drude_steps_per_real_step;
    o->useDrudeNH = p->use_drude_nh_chains != 0;
    o->useCOM = p->use_com_temp_group != 0;
    o->kT = TGNH_ORACLE_BOLTZ * p->temperature;         /* CudaDrudeTGNHKernels.cpp:80-81 */
    o->kTD = TGNH_ORACLE_BOLTZ * p->drude_temperature;
    o->dt = p->step_size; o->rmax = p->max_drude_distance;
    o->mass = xcalloc(o->N, sizeof(double));
    o->invMass = xcalloc(o->N, sizeof(double));
    o->tg = xcalloc(o->N, sizeof(int));
    o->resid = xcalloc(o->N, sizeof(int));
    o->pairD = xcalloc(o->P, sizeof(int));
    o->pairP = xcalloc(o->P, sizeof(int));
    char* inPair = xcalloc(o->N, 1);
    for (int i = 0; i < o->N; i++) {
        o->mass[i] = p->masses[i];
        o->invMass[i] = p->masses[i] == 0.0 ? 0.0 : 1.0 / p->masses[i];  /* velm.w / particleInvMass */
        o->tg[i] = p->particle_temp_group ? p->particle_temp_group[i] : 0;
        o->resid[i] = p->particle_res_id ? p->particle_res_id[i] : 0;
    }
    for (int i = 0; i < o->P; i++) {
        o->pairD[i] = p->pair_drude[i]; o->pairP[i] = p->pair_parent[i];
        if (o->pairD[i] < 0 || o->pairD[i] >= o->N || o->pairP[i] < 0 || o->pairP[i] >= o->N) {
            free(inPair); tgnh_oracle_destroy(o); FAIL("pair %d has particle index out of range", i);
        }
        inPair[o->pairD[i]] = 1; inPair[o->pairP[i]] = 1;
    }
    /* normalParticles = sorted set difference (CudaDrudeTGNHKernels.cpp:111,141-142,151) */
    o->normal = xcalloc(o->N, sizeof(int));
    for (int i = 0; i < o->N; i++) if (!inPair[i]) o->normal[o->numNormal++] = i;
    free(inPair);
    int rc = (which == TGNH_ORACLE_REF) ? init_ref(o, p) : init_tg(o, p);
    if (rc) { tgnh_oracle_destroy(o); return rc; }
    *out = o;
    return 0;
}

void tgnh_oracle_destroy(tgnh_oracle* o) {
    if (!o) return;
    free(o->mass); free(o->invMass); free(o->pairD); free(o->pairP); free(o->tg); free(o->resid); free(o->normal);
    free(o->resCount); free(o->resFirst); free(o->tgDof); free(o->tgNkbT);
    free(o->etaMass); free(o->eta); free(o->etaDot); free(o->etaDotDot); free(o->vscale); free(o->ke2);
    free(o->comVel); free(o->normVel);
    free(o->pairInvTotalMass); free(o->pairInvReducedMass);
    free(o->fEtaMass); free(o->fEta); free(o->fEtaDot); free(o->fEtaDotDot);
    free(o);
}

/* ------------------------------------------------------------------------------------------------
 * TG layer kernels
 * ---------------------------------------------------------------------------------------------- */

/* calcCOMVelocities (drudeTGNH.cu:82-113) + normalizeVelocities (:119-133) */
static void tg_com_and_norm(tgnh_oracle* o, const double* vel) {
    PAR_FOR
    for (int r = 0; r < o->R; r++) {
        double* c = o->comVel + 4 * r;
        c[0] = c[1] = c[2] = c[3] = 0;
        if (o->useCOM) {
            double comMass = 0.0;
            for (int j = 0; j < o->resCount[r]; j++) {
                int index = o->resFirst[r] + j;
                double w = o->invMass[index];
                if (w != 0) {
                    double mass = 1.0 / w;
                    c[0] += vel[3 * index] * mass;
                    c[1] += vel[3 * index + 1] * mass;
                    c[2] += vel[3 * index + 2] * mass;
                    comMass += mass;
                }
            }
            c[3] = 1.0 / comMass;
            c[0] *= c[3]; c[1] *= c[3]; c[2] *= c[3];
        } else {
            c[3] = 1.0;
        }
    }
    PAR_FOR
    for (int i = 0; i < o->N; i++) {
        const double* c = o->comVel + 4 * o->resid[i];
        o->normVel[3 * i] = vel[3 * i] - c[0];
        o->normVel[3 * i + 1] = vel[3 * i + 1] - c[1];
        o->normVel[3 * i + 2] = vel[3 * i + 2] - c[2];
    }
}

/* computeNormalizedKineticEnergies (drudeTGNH.cu:138-200) + sumNormalizedKineticEnergies (:202-242);
 * serial accumulation order (the device order depends on launch geometry, D7). */
static void tg_ke_range(const tgnh_oracle* o, double* ke2, int r0, int r1, int n0, int n1, int p0, int p1) {
    const int G = o->G;
    for (int r = r0; r < r1; r++) {
        const double* c = o->comVel + 4 * r;
        ke2[G] += (c[0] * c[0] + c[1] * c[1] + c[2] * c[2]) / c[3];
    }
    for (int i = n0; i < n1; i++) {
        int index = o->normal[i];
        double w = o->invMass[index];
        const double* v = o->normVel + 3 * index;
        if (w != 0) ke2[o->tg[index]] += (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / w;
    }
    for (int i = p0; i < p1; i++) {
        int px = o->pairD[i], py = o->pairP[i];
        const double* v1 = o->normVel + 3 * px;
        const double* v2 = o->normVel + 3 * py;
        double w1 = o->invMass[px], w2 = o->invMass[py];
        double mass1 = 1.0 / w1, mass2 = 1.0 / w2;
        double invTotalMass = 1.0 / (mass1 + mass2);
        double invReducedMass = (mass1 + mass2) * w1 * w2;
        double mass1fract = invTotalMass * mass1, mass2fract = invTotalMass * mass2;
        double cm[3], rel[3];
        for (int k = 0; k < 3; k++) { cm[k] = v1[k] * mass1fract + v2[k] * mass2fract; rel[k] = v2[k] - v1[k]; }
        ke2[o->tg[px]] += (cm[0] * cm[0] + cm[1] * cm[1] + cm[2] * cm[2]) * (mass1 + mass2);
        ke2[G + 1] += (rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]) * (1.0 / invReducedMass);
    }
}

static void tg_kinetic_energies(tgnh_oracle* o, double* ke2) {
    const int T = o->G + 2;
    for (int g = 0; g < T; g++) ke2[g] = 0;
    if (g_threads <= 1) { tg_ke_range(o, ke2, 0, o->R, 0, o->numNormal, 0, o->P); return; }
    const int nt = g_threads;
    double* part = xcalloc((size_t)nt * T, sizeof(double));
    _Pragma("omp parallel for schedule(static) num_threads(nt)")
    for (int t = 0; t < nt; t++)
        tg_ke_range(o, part + (size_t)t * T, (int)((long long)o->R * t / nt), (int)((long long)o->R * (t + 1) / nt),
                    (int)((long long)o->numNormal * t / nt), (int)((long long)o->numNormal * (t + 1) / nt),
                    (int)((long long)o->P * t / nt), (int)((long long)o->P * (t + 1) / nt));
    for (int t = 0; t < nt; t++) for (int g = 0; g < T; g++) ke2[g] += part[(size_t)t * T + g];
    free(part);
}

/* host chain: CudaDrudeTGNHKernels.cpp:559-642.  `ke` is consumed (scaled in place like kineticEnergiesVec). */
static void tg_chain(tgnh_oracle* o, double* ke) {
    const int G = o->G, M = o->M, S = o->S;
    const double dtc = o->dt / S, dtc2 = dtc / 2.0, dtc4 = dtc / 4.0, dtc8 = dtc / 8.0;
    const double realkbT = o->kT, drudekbT = o->kTD;
    for (int g = 0; g < G + 2; g++) o->vscale[g] = 1.0;                                    /* :446 */
    for (int itg = 0; itg < G + 1; itg++) {                                               /* :560-595 */
        double* Q = o->etaMass + itg * M; double* eta = o->eta + itg * M;
        double* ed = o->etaDot + itg * (M + 1); double* edd = o->etaDotDot + itg * M;
        double expfac = 1.0;                                                              /* :559 */
        if (Q[0] > 0) edd[0] = (ke[itg] - o->tgNkbT[itg]) / Q[0];
        for (int iter = 0; iter < S; iter++) {
            for (int i = M - 1; i >= 0; i--) {
                expfac = exp(-dtc8 * ed[i + 1]);
                ed[i] *= expfac; ed[i] += edd[i] * dtc4; ed[i] *= expfac;
            }
            o->vscale[itg] *= exp(-dtc2 * ed[0]);
            ke[itg] *= exp(-dtc * ed[0]);
            for (int i = 0; i < M; i++) eta[i] += dtc2 * ed[i];
            if (Q[0] > 0) edd[0] = (ke[itg] - o->tgNkbT[itg]) / Q[0];
            ed[0] *= expfac; ed[0] += edd[0] * dtc4; ed[0] *= expfac;
            for (int i = 1; i < M; i++) {
                expfac = exp(-dtc8 * ed[i + 1]);
                ed[i] *= expfac;
                edd[i] = (Q[i - 1] * ed[i - 1] * ed[i - 1] - realkbT) / Q[i];
                ed[i] += edd[i] * dtc4; ed[i] *= expfac;
            }
        }
    }
    {                                                                                     /* :597-642 */
        const int itg = G + 1;
        double* Q = o->etaMass + itg * M; double* eta = o->eta + itg * M;
        double* ed = o->etaDot + itg * (M + 1); double* edd = o->etaDotDot + itg * M;
        double expfac = 1.0;
        edd[0] = (ke[itg] - o->tgNkbT[itg]) / Q[0];
        for (int iter = 0; iter < S; iter++) {
            if (o->useDrudeNH)
                for (int i = M - 1; i > 0; i--) {
                    expfac = exp(-dtc8 * ed[i + 1]);
                    ed[i] *= expfac; ed[i] += edd[i] * dtc4; ed[i] *= expfac;
                }
            expfac = exp(-dtc8 * ed[1]);
            ed[0] *= expfac; ed[0] += edd[0] * dtc4; ed[0] *= expfac;
            o->vscale[itg] *= exp(-dtc2 * ed[0]);
            ke[itg] *= exp(-dtc * ed[0]);
            eta[0] += dtc2 * ed[0];
            if (o->useDrudeNH) for (int i = 1; i < M; i++) eta[i] += dtc2 * ed[i];
            edd[0] = (ke[itg] - o->tgNkbT[itg]) / Q[0];
            ed[0] *= expfac; ed[0] += edd[0] * dtc4; ed[0] *= expfac;
            if (o->useDrudeNH)
                for (int i = 1; i < M; i++) {
                    expfac = exp(-dtc8 * ed[i + 1]);
                    ed[i] *= expfac;
                    edd[i] = (Q[i - 1] * ed[i - 1] * ed[i - 1] - drudekbT) / Q[i];
                    ed[i] += edd[i] * dtc4; ed[i] *= expfac;
                }
        }
    }
}

/* integrateDrudeTGNHChain (drudeTGNH.cu:249-301) */
static void tg_scale(tgnh_oracle* o, double* vel) {
    const int G = o->G;
    const double vscaleCOM = o->vscale[G], vscaleDrude = o->vscale[G + 1];
    PAR_FOR
    for (int i = 0; i < o->numNormal; i++) {
        int index = o->normal[i];
        double* v = vel + 3 * index; const double* vr = o->normVel + 3 * index;
        double vscale = o->vscale[o->tg[index]];
        if (o->invMass[index] != 0)
            for (int k = 0; k < 3; k++) v[k] = vscale * vr[k] + vscaleCOM * (v[k] - vr[k]);
    }
    PAR_FOR
    for (int i = 0; i < o->P; i++) {
        int px = o->pairD[i], py = o->pairP[i];
        double vscaleCM = o->vscale[o->tg[px]];
        double* v1 = vel + 3 * px; double* v2 = vel + 3 * py;
        const double* r1 = o->normVel + 3 * px; const double* r2 = o->normVel + 3 * py;
        double mass1 = 1.0 / o->invMass[px], mass2 = 1.0 / o->invMass[py];
        double invTotalMass = 1.0 / (mass1 + mass2);
        double mass1fract = invTotalMass * mass1, mass2fract = invTotalMass * mass2;
        for (int k = 0; k < 3; k++) {
            double velCOM1 = v1[k] - r1[k], velCOM2 = v2[k] - r2[k];
            double cmVel = r1[k] * mass1fract + r2[k] * mass2fract;
            double relVel = r2[k] - r1[k];
            cmVel = vscaleCM * cmVel;
            relVel = vscaleDrude * relVel;
            v1[k] = cmVel - relVel * mass2fract + vscaleCOM * velCOM1;
            v2[k] = cmVel + relVel * mass1fract + vscaleCOM * velCOM2;
        }
    }
}

/* integrateDrudeTGNHVelocities (drudeTGNH.cu:307-365); fscale = 0.5*dt (the 2^32 fixed-point factor
 * of CudaDrudeTGNHKernels.cpp:295 belongs to the int64 force format, forces here are doubles). */
static void tg_half_kick(tgnh_oracle* o, double* vel, const double* force) {
    const double fscale = 0.5 * o->dt, fscaleDrude = fscale;
    PAR_FOR
    for (int i = 0; i < o->numNormal; i++) {
        int index = o->normal[i];
        double w = o->invMass[index];
        if (w != 0)
            for (int k = 0; k < 3; k++) vel[3 * index + k] = vel[3 * index + k] + fscale * w * force[3 * index + k];
    }
    PAR_FOR
    for (int i = 0; i < o->P; i++) {
        int px = o->pairD[i], py = o->pairP[i];
        double* v1 = vel + 3 * px; double* v2 = vel + 3 * py;
        const double* f1 = force + 3 * px; const double* f2 = force + 3 * py;
        double w1 = o->invMass[px], w2 = o->invMass[py];
        double mass1 = 1.0 / w1, mass2 = 1.0 / w2;
        double invTotalMass = 1.0 / (mass1 + mass2);
        double invReducedMass = (mass1 + mass2) * w1 * w2;
        double mass1fract = invTotalMass * mass1, mass2fract = invTotalMass * mass2;
        for (int k = 0; k < 3; k++) {
            double cmVel = v1[k] * mass1fract + v2[k] * mass2fract;
            double relVel = v2[k] - v1[k];
            double cmForce = f1[k] + f2[k];
            double relForce = f2[k] * mass1fract - f1[k] * mass2fract;
            cmVel = cmVel + fscale * invTotalMass * cmForce;
            relVel = relVel + fscaleDrude * invReducedMass * relForce;
            v1[k] = cmVel - relVel * mass2fract;
            v2[k] = cmVel + relVel * mass1fract;
        }
    }
}

/* posDelta = dt*v (drudeTGNH.cu:322-324,360-363) then integrateDrudeTGNHPositions (:435-466);
 * no constraints in between on the integrator-only path. */
static void tg_drift(tgnh_oracle* o, double* pos, double* vel) {
    const double invStepSize = 1.0 / o->dt;
    PAR_FOR
    for (int i = 0; i < o->N; i++) {
        if (o->invMass[i] != 0)
            for (int k = 0; k < 3; k++) {
                double delta = o->dt * vel[3 * i + k];
                pos[3 * i + k] += delta;
                vel[3 * i + k] = invStepSize * delta;
            }
    }
}

/* applyHardWallConstraints (drudeTGNH.cu:471-574); identical arithmetic in
 * ReferenceDrudeTGNHKernels.cpp:298-363 except for the throw at :311-312 (D5), restated only for REF. */
static int hard_wall(tgnh_oracle* o, double* pos, double* vel) {
    const double maxDrudeDistance = o->rmax;
    if (!(maxDrudeDistance > 0)) return 0;
    const double hardwallscaleDrude = sqrt(o->kTD);
    const double stepSize = o->dt;
    int tooFar = 0;
    _Pragma("omp parallel for schedule(static) num_threads(g_threads) if (g_threads > 1) reduction(|:tooFar)")
    for (int i = 0; i < o->P; i++) {
        int p1 = o->pairD[i], p2 = o->pairP[i];
        double* x1 = pos + 3 * p1; double* x2 = pos + 3 * p2;
        double delta[3] = { x1[0] - x2[0], x1[1] - x2[1], x1[2] - x2[2] };
        double r = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
        double rInv = 1.0 / r;
        if (rInv * maxDrudeDistance < 1) {
            if (o->which == TGNH_ORACLE_REF && rInv * maxDrudeDistance < 0.5) { tooFar |= 1; continue; }
            double bondDir[3] = { delta[0] * rInv, delta[1] * rInv, delta[2] * rInv };
            double* vel1 = vel + 3 * p1; double* vel2 = vel + 3 * p2;
            double mass1, mass2;
            int parentMassless;
            if (o->which == TGNH_ORACLE_REF) { mass1 = o->mass[p1]; mass2 = o->mass[p2]; parentMassless = (mass2 == 0); }
            else { mass1 = 1.0 / o->invMass[p1]; mass2 = 1.0 / o->invMass[p2]; parentMassless = (o->invMass[p2] == 0); }
            double deltaR = r - maxDrudeDistance;
            double deltaT = stepSize;
            double dotvr1 = vel1[0] * bondDir[0] + vel1[1] * bondDir[1] + vel1[2] * bondDir[2];
            double vp1[3] = { vel1[0] - bondDir[0] * dotvr1, vel1[1] - bondDir[1] * dotvr1, vel1[2] - bondDir[2] * dotvr1 };
            if (parentMassless) {
                if (dotvr1 != 0) deltaT = deltaR / fabs(dotvr1);
                if (deltaT > stepSize) deltaT = stepSize;
                dotvr1 = -dotvr1 * hardwallscaleDrude / (fabs(dotvr1) * sqrt(mass1));
                double dr = -deltaR + deltaT * dotvr1;
                for (int k = 0; k < 3; k++) { x1[k] += bondDir[k] * dr; vel1[k] = vp1[k] + bondDir[k] * dotvr1; }
            } else {
                double invTotalMass = (o->which == TGNH_ORACLE_REF) ? o->pairInvTotalMass[i] : 1.0 / (mass1 + mass2);
                double dotvr2 = vel2[0] * bondDir[0] + vel2[1] * bondDir[1] + vel2[2] * bondDir[2];
                double vp2[3] = { vel2[0] - bondDir[0] * dotvr2, vel2[1] - bondDir[1] * dotvr2, vel2[2] - bondDir[2] * dotvr2 };
                double vbCMass = (mass1 * dotvr1 + mass2 * dotvr2) * invTotalMass;
                dotvr1 -= vbCMass;
                dotvr2 -= vbCMass;
                if (dotvr1 != dotvr2) deltaT = deltaR / fabs(dotvr1 - dotvr2);
                if (deltaT > stepSize) deltaT = stepSize;
                double vBond = hardwallscaleDrude / sqrt(mass1);
                dotvr1 = -dotvr1 * vBond * mass2 * invTotalMass / fabs(dotvr1);
                dotvr2 = -dotvr2 * vBond * mass1 * invTotalMass / fabs(dotvr2);
                double dr1 = -deltaR * mass2 * invTotalMass + deltaT * dotvr1;
                double dr2 = deltaR * mass1 * invTotalMass + deltaT * dotvr2;
                dotvr1 += vbCMass;
                dotvr2 += vbCMass;
                for (int k = 0; k < 3; k++) {
                    x1[k] += bondDir[k] * dr1; x2[k] += bondDir[k] * dr2;
                    vel1[k] = vp1[k] + bondDir[k] * dotvr1; vel2[k] = vp2[k] + bondDir[k] * dotvr2;
                }
            }
        }
    }
    if (tooFar) FAIL("Drude particle moved too far beyond hard wall constraint");   /* ReferenceDrudeTGNHKernels.cpp:311-312 */
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * REF layer: ReferenceDrudeTGNHKernels.cpp:426-546 (chain+scale), :548-584 (kick), :253-284 (drift)
 * ---------------------------------------------------------------------------------------------- */
static void ref_propagate_nh_chain(tgnh_oracle* o, double* vel) {
    const int S = o->S;
    const double dt = o->dt, dtc = dt / S, dtc2 = dtc / 2.0, dtc4 = dtc / 4.0, dtc8 = dtc / 8.0;
    const double realkbT = o->kT, drudekbT = o->kTD;
    double realKE = 0.0, drudeKE = 0.0;
    for (int i = 0; i < o->numNormal; i++) {                                   /* :443-448 */
        int index = o->normal[i];
        const double* v = vel + 3 * index;
        if (o->invMass[index] != 0) realKE += (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) / o->invMass[index];
    }
    for (int i = 0; i < o->P; i++) {                                           /* :451-460 */
        int p1 = o->pairD[i], p2 = o->pairP[i];
        double mass1fract = o->pairInvTotalMass[i] / o->invMass[p1];
        double mass2fract = o->pairInvTotalMass[i] / o->invMass[p2];
        double cm[3], rel[3];
        for (int k = 0; k < 3; k++) { cm[k] = vel[3 * p1 + k] * mass1fract + vel[3 * p2 + k] * mass2fract; rel[k] = vel[3 * p2 + k] - vel[3 * p1 + k]; }
        realKE += (cm[0] * cm[0] + cm[1] * cm[1] + cm[2] * cm[2]) / o->pairInvTotalMass[i];
        drudeKE += (rel[0] * rel[0] + rel[1] * rel[1] + rel[2] * rel[2]) / o->pairInvReducedMass[i];
    }
    o->realKE2 = realKE; o->drudeKE2 = drudeKE;
    o->KESum = 0.5 * (realKE + drudeKE);
    double scaleReal = 1.0, scaleDrude = 1.0, expfac = 1.0;
    double* etaMass = o->fEtaMass; double* eta = o->fEta; double* etaDot = o->fEtaDot; double* etaDotDot = o->fEtaDotDot;
    const int ntg = o->numTempGroupRef;
    etaDotDot[0] = (realKE - o->realNkbT) / etaMass[0];                        /* :471-472 */
    etaDotDot[1] = (drudeKE - o->drudeNkbT) / etaMass[1];
    for (int iter = 0; iter < S; iter++) {                                     /* :474-504 (D1 restated as written) */
        for (int i = o->idxMaxNHChains; i >= 0; i--) {
            expfac = exp(-dtc8 * etaDot[i + ntg]);
            etaDot[i] *= expfac; etaDot[i] += etaDotDot[i] * dtc4; etaDot[i] *= expfac;
        }
        scaleReal *= exp(-dtc2 * etaDot[0]);
        scaleDrude *= exp(-dtc2 * etaDot[1]);
        realKE *= exp(-dtc * etaDot[0]);
        drudeKE *= exp(-dtc * etaDot[1]);
        for (int i = 0; i < o->iNumNHChains; i++) eta[i] += dtc2 * etaDot[i];
        etaDotDot[0] = (realKE - o->realNkbT) / etaMass[0];
        etaDotDot[1] = (drudeKE - o->drudeNkbT) / etaMass[1];
        for (int i = 0; i < o->iNumNHChains; i++) {
            expfac = exp(-dtc8 * etaDot[i + 2]);
            etaDot[i] *= expfac;
            if (i > 1) {
                double dofkbT = (i % 2 == 0 ? realkbT : drudekbT);
                etaDotDot[i] = (etaMass[i - 2] * etaDot[i - 2] * etaDot[i - 2] - dofkbT) / etaMass[i];
            }
            etaDot[i] += etaDotDot[i] * dtc4; etaDot[i] *= expfac;
        }
    }
    o->scaleReal = scaleReal; o->scaleDrude = scaleDrude;
    for (int i = 0; i < o->numNormal; i++) {                                   /* :517-524 */
        int index = o->normal[i];
        if (o->invMass[index] != 0.0) for (int k = 0; k < 3; k++) vel[3 * index + k] = scaleReal * vel[3 * index + k];
    }
    for (int i = 0; i < o->P; i++) {                                           /* :527-541 */
        int p1 = o->pairD[i], p2 = o->pairP[i];
        double mass1fract = o->pairInvTotalMass[i] / o->invMass[p1];
        double mass2fract = o->pairInvTotalMass[i] / o->invMass[p2];
        for (int k = 0; k < 3; k++) {
            double cmVel = vel[3 * p1 + k] * mass1fract + vel[3 * p2 + k] * mass2fract;
            double relVel = vel[3 * p2 + k] - vel[3 * p1 + k];
            cmVel = scaleReal * cmVel;
            relVel = scaleDrude * relVel;
            vel[3 * p1 + k] = cmVel - relVel * mass2fract;
            vel[3 * p2 + k] = cmVel + relVel * mass1fract;
        }
    }
}

static void ref_half_kick(tgnh_oracle* o, double* vel, const double* force) {
    const double dt = o->dt;
    for (int i = 0; i < o->numNormal; i++) {                                   /* :555-562 */
        int index = o->normal[i];
        double invMass = o->invMass[index];
        if (invMass != 0.0) for (int k = 0; k < 3; k++) vel[3 * index + k] += 0.5 * dt * invMass * force[3 * index + k];
    }
    for (int i = 0; i < o->P; i++) {                                           /* :565-583 */
        int p1 = o->pairD[i], p2 = o->pairP[i];
        double mass1fract = o->pairInvTotalMass[i] / o->invMass[p1];
        double mass2fract = o->pairInvTotalMass[i] / o->invMass[p2];
        for (int k = 0; k < 3; k++) {
            double cmVel = vel[3 * p1 + k] * mass1fract + vel[3 * p2 + k] * mass2fract;
            double relVel = vel[3 * p2 + k] - vel[3 * p1 + k];
            double cmForce = force[3 * p1 + k] + force[3 * p2 + k];
            double relForce = force[3 * p2 + k] * mass1fract - force[3 * p1 + k] * mass2fract;
            cmVel += 0.5 * dt * o->pairInvTotalMass[i] * cmForce;
            relVel += 0.5 * dt * o->pairInvReducedMass[i] * relForce;
            vel[3 * p1 + k] = cmVel - relVel * mass2fract;
            vel[3 * p2 + k] = cmVel + relVel * mass1fract;
        }
    }
}

static void ref_drift(tgnh_oracle* o, double* pos, double* vel) {
    const double dt = o->dt, dtInv = 1.0 / dt;                                  /* :253-284, constraints: none */
    for (int i = 0; i < o->N; i++)
        if (o->invMass[i] != 0.0)
            for (int k = 0; k < 3; k++) {
                double xPrime = pos[3 * i + k] + vel[3 * i + k] * dt;
                vel[3 * i + k] = (xPrime - pos[3 * i + k]) * dtInv;
                pos[3 * i + k] = xPrime;
            }
}

/* ------------------------------------------------------------------------------------------------
 * public phases
 * ---------------------------------------------------------------------------------------------- */
int tgnh_oracle_propagate_nh_chain(tgnh_oracle* o, double* vel) {
    if (o->which == TGNH_ORACLE_REF) { ref_propagate_nh_chain(o, vel); return 0; }
    /* CudaDrudeTGNHKernels.cpp:469-497 then :559-650, then the scaling launch :351-353/:402 */
    tg_com_and_norm(o, vel);
    tg_kinetic_energies(o, o->ke2);
    double KESum = 0.0;
    for (int g = 0; g < o->G + 2; g++) KESum += o->ke2[g];
    o->KESum = 0.5 * KESum;
    double* ke = xcalloc(o->G + 2, sizeof(double));
    memcpy(ke, o->ke2, (o->G + 2) * sizeof(double));
    tg_chain(o, ke);
    free(ke);
    tg_scale(o, vel);
    return 0;
}

int tgnh_oracle_compute_ke2(tgnh_oracle* o, const double* vel, double* ke2) {
    if (o->which == TGNH_ORACLE_REF) FAIL("compute_ke2 is a TG-layer call");
    tg_com_and_norm(o, vel);
    tg_kinetic_energies(o, ke2);
    return 0;
}

int tgnh_oracle_half_kick(tgnh_oracle* o, double* vel, const double* force) {
    if (o->which == TGNH_ORACLE_REF) ref_half_kick(o, vel, force); else tg_half_kick(o, vel, force);
    return 0;
}

int tgnh_oracle_drift(tgnh_oracle* o, double* pos, double* vel) {
    if (o->which == TGNH_ORACLE_REF) ref_drift(o, pos, vel); else tg_drift(o, pos, vel);
    return 0;
}

int tgnh_oracle_hard_wall(tgnh_oracle* o, double* pos, double* vel) { return hard_wall(o, pos, vel); }

static void harmonic_forces(tgnh_oracle* o, const double* pos, double* force, const double* ext, const double* k) {
    const size_t n3 = (size_t)o->N * 3;
    if (ext) memcpy(force, ext, n3 * sizeof(double)); else memset(force, 0, n3 * sizeof(double));
    for (int i = 0; i < o->P; i++) {
        int d = o->pairD[i], p = o->pairP[i];
        for (int c = 0; c < 3; c++) {
            double f = -k[i] * (pos[3 * d + c] - pos[3 * p + c]);
            force[3 * d + c] += f;
            force[3 * p + c] -= f;
        }
    }
}

/* step order: CudaDrudeTGNHKernels.cpp:336-402 == ReferenceDrudeTGNHKernels.cpp:231-406 without
 * constraints / virtual sites (integrator-only path). */
int tgnh_oracle_step(tgnh_oracle* o, double* pos, double* vel, double* force, int nsteps,
                     int force_model, const double* ext_force, const double* k_spring) {
    if (force_model == TGNH_ORACLE_FORCE_HARMONIC && !k_spring) FAIL("harmonic force model needs k_spring");
    for (int s = 0; s < nsteps; s++) {
        tgnh_oracle_propagate_nh_chain(o, vel);
        tgnh_oracle_half_kick(o, vel, force);
        tgnh_oracle_drift(o, pos, vel);
        if (hard_wall(o, pos, vel)) return 1;
        if (force_model == TGNH_ORACLE_FORCE_HARMONIC) harmonic_forces(o, pos, force, ext_force, k_spring);
        tgnh_oracle_half_kick(o, vel, force);
        tgnh_oracle_propagate_nh_chain(o, vel);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * inspection
 * ---------------------------------------------------------------------------------------------- */
int tgnh_oracle_num_thermostats(const tgnh_oracle* o) { return o->G + 2; }

/* REF flat index of (thermostat t in {0 real, 2 drude}, link i); -1 if the link does not exist */
static int ref_flat(const tgnh_oracle* o, int t, int i) {
    if (o->useDrudeNH) return i * 2 + (t == 2 ? 1 : 0);
    if (t == 2) return i == 0 ? 1 : -1;
    return i == 0 ? 0 : i + 1;
}

void tgnh_oracle_get_ke2(const tgnh_oracle* o, double* ke2) {
    if (o->which == TGNH_ORACLE_REF) { ke2[0] = o->realKE2; ke2[1] = 0; ke2[2] = o->drudeKE2; return; }
    memcpy(ke2, o->ke2, (o->G + 2) * sizeof(double));
}

void tgnh_oracle_get_vscale(const tgnh_oracle* o, double* vscale) {
    if (o->which == TGNH_ORACLE_REF) { vscale[0] = o->scaleReal; vscale[1] = 1.0; vscale[2] = o->scaleDrude; return; }
    memcpy(vscale, o->vscale, (o->G + 2) * sizeof(double));
}

void tgnh_oracle_get_chain_state(const tgnh_oracle* o, double* eta, double* eta_dot, double* eta_dot_dot) {
    const int T = o->G + 2, M = o->M;
    if (o->which == TGNH_ORACLE_REF) {
        memset(eta, 0, sizeof(double) * T * M); memset(eta_dot, 0, sizeof(double) * T * (M + 1)); memset(eta_dot_dot, 0, sizeof(double) * T * M);
        for (int t = 0; t < 3; t += 2)
            for (int i = 0; i < M; i++) {
                int f = ref_flat(o, t, i);
                if (f < 0) continue;
                eta[t * M + i] = o->fEta[f]; eta_dot[t * (M + 1) + i] = o->fEtaDot[f]; eta_dot_dot[t * M + i] = o->fEtaDotDot[f];
            }
        return;
    }
    memcpy(eta, o->eta, sizeof(double) * T * M);
    memcpy(eta_dot, o->etaDot, sizeof(double) * T * (M + 1));
    memcpy(eta_dot_dot, o->etaDotDot, sizeof(double) * T * M);
}

void tgnh_oracle_set_chain_state(tgnh_oracle* o, const double* eta, const double* eta_dot, const double* eta_dot_dot) {
    const int T = o->G + 2, M = o->M;
    if (o->which == TGNH_ORACLE_REF) {
        for (int t = 0; t < 3; t += 2)
            for (int i = 0; i < M; i++) {
                int f = ref_flat(o, t, i);
                if (f < 0) continue;
                o->fEta[f] = eta[t * M + i]; o->fEtaDot[f] = eta_dot[t * (M + 1) + i]; o->fEtaDotDot[f] = eta_dot_dot[t * M + i];
            }
        return;
    }
    memcpy(o->eta, eta, sizeof(double) * T * M);
    memcpy(o->etaDot, eta_dot, sizeof(double) * T * (M + 1));
    memcpy(o->etaDotDot, eta_dot_dot, sizeof(double) * T * M);
}

void tgnh_oracle_get_thermostat_params(const tgnh_oracle* o, double* dof, double* nkbt, double* eta_mass) {
    const int T = o->G + 2, M = o->M;
    if (o->which == TGNH_ORACLE_REF) {
        dof[0] = o->tgDof[0]; dof[1] = 0; dof[2] = o->tgDof[2];
        nkbt[0] = o->realNkbT; nkbt[1] = 0; nkbt[2] = o->drudeNkbT;
        memset(eta_mass, 0, sizeof(double) * T * M);
        for (int t = 0; t < 3; t += 2)
            for (int i = 0; i < M; i++) { int f = ref_flat(o, t, i); if (f >= 0) eta_mass[t * M + i] = o->fEtaMass[f]; }
        return;
    }
    memcpy(dof, o->tgDof, sizeof(double) * T);
    memcpy(nkbt, o->tgNkbT, sizeof(double) * T);
    memcpy(eta_mass, o->etaMass, sizeof(double) * T * M);
}

double tgnh_oracle_get_ke_sum(const tgnh_oracle* o) { return o->KESum; }
