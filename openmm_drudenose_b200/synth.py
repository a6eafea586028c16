"""Deterministic synthetic Drude-polarizable systems (BASELINE.json configs C1-C5).

A system is described the way the reference integrator sees it after
DrudeTGNHIntegrator::initialize (openmmapi/src/DrudeTGNHIntegrator.cpp:103-160): particle
masses, the DrudeForce (drude, parent) pairs, per-particle temperature groups and residue
(molecule) ids.  Random data comes from Philox streams keyed by (seed, chunk of molecules), so
any contiguous molecule range (a shard) reproduces exactly the same particles.

numpy only; no GPU code here.
"""
from dataclasses import dataclass, field

import numpy as np

BOLTZ = 1.380649e-23 * 6.02214076e23 / 1000.0      # kJ/mol/K, OpenMM's BOLTZ
ONE_4PI_EPS0 = 138.935456                          # OpenMM's ONE_4PI_EPS0 (older literal used by the tests' era)
SEED = 20261018
_CHUNK = 1 << 16                                    # molecules per Philox stream


@dataclass
class Template:
    """One molecule type: masses, local (drude, parent) pairs, local geometry offsets (nm)."""
    name: str
    masses: np.ndarray
    pairs: np.ndarray            # [p, 2] local (drude, parent)
    offsets: np.ndarray          # [k, 3] position of each particle relative to the first
    constraints: np.ndarray = field(default_factory=lambda: np.zeros((0, 2), np.int32))  # DOF bookkeeping only

    @property
    def size(self):
        return len(self.masses)


@dataclass
class DrudeSystem:
    masses: np.ndarray           # [N] f64
    pair_drude: np.ndarray       # [P] i32   (pairParticles.x, CudaDrudeTGNHKernels.cpp:147)
    pair_parent: np.ndarray      # [P] i32   (pairParticles.y)
    temp_group: np.ndarray       # [N] i32
    res_id: np.ndarray           # [N] i32
    positions: np.ndarray        # [N,3] f64 (values representable in f32)
    velocities: np.ndarray       # [N,3] f64 (values representable in f32)
    forces: np.ndarray           # [N,3] f64 (values representable in f32)
    k_spring: np.ndarray         # [P] f64
    constraints: np.ndarray      # [C,2] i32 (DOF bookkeeping only; never applied)
    num_temp_groups: int
    num_residues: int
    # integrator parameters (DrudeTGNHIntegrator.h:71)
    temperature: float = 300.0
    coupling_time: float = 0.1
    drude_temperature: float = 1.0
    drude_coupling_time: float = 0.005
    step_size: float = 0.001
    drude_steps: int = 20
    num_nh_chains: int = 3
    use_drude_nh_chains: bool = True
    use_com_temp_group: bool = True
    max_drude_distance: float = 0.02

    @property
    def num_particles(self):
        return len(self.masses)

    @property
    def num_pairs(self):
        return len(self.pair_drude)

    @property
    def inv_masses(self):
        with np.errstate(divide="ignore"):
            return np.where(self.masses == 0.0, 0.0, 1.0 / np.where(self.masses == 0.0, 1.0, self.masses))

    # ---- boundary layouts (OpenMM single-precision arrays) ----
    def velm_f32(self):
        """float4 (vx, vy, vz, 1/m) like cu.getVelm() in single precision."""
        out = np.zeros((self.num_particles, 4), np.float32)
        out[:, :3] = self.velocities
        out[:, 3] = self.inv_masses
        return out

    def posq_f32(self, charges=None):
        """float4 (x, y, z, q) like cu.getPosq()."""
        out = np.zeros((self.num_particles, 4), np.float32)
        out[:, :3] = self.positions
        if charges is not None:
            out[:, 3] = charges
        return out

    def force_f32_soa(self, padded=None):
        """float SoA [3, paddedN]: force[i + k*paddedN] (same indexing as cu.getForce(), fp32 instead of int64)."""
        n = self.num_particles
        padded = padded or n
        out = np.zeros((3, padded), np.float32)
        out[:, :n] = self.forces.T
        return out

    def force_i64_soa(self, padded=None):
        """OpenMM's fixed-point force buffer: long long [3, paddedN], scale 2^32 (CudaDrudeTGNHKernels.cpp:295)."""
        n = self.num_particles
        padded = padded or n
        out = np.zeros((3, padded), np.int64)
        out[:, :n] = np.rint(self.forces.T * 4294967296.0).astype(np.int64)
        return out


def _f32(x):
    return np.asarray(x, np.float32).astype(np.float64)


# ---- molecule templates -------------------------------------------------------------------
# masses from platforms/cuda/tests/TestCudaDrudeTGNHIntegrator.cpp:132-136; geometry :158-161
WATER4 = Template("water4", np.array([15.6, 0.4, 1.0, 1.0]), np.array([[1, 0]], np.int32),
                  np.array([[0, 0, 0], [0, 0, 0], [0.09572, 0, 0], [-0.023999, 0.092663, 0]]))
SWM4 = Template("swm4", np.array([15.6, 0.4, 1.0, 1.0, 0.0]), np.array([[1, 0]], np.int32),
                np.array([[0, 0, 0], [0, 0, 0], [0.09572, 0, 0], [-0.023999, 0.092663, 0], [0.0, 0.0247, 0.0]]),
                np.array([[0, 2], [0, 3], [2, 3]], np.int32))
SOD = Template("sod", np.array([22.59, 0.4]), np.array([[1, 0]], np.int32), np.zeros((2, 3)))
CLA = Template("cla", np.array([35.05, 0.4]), np.array([[1, 0]], np.int32), np.zeros((2, 3)))


def polymer_template(heavy_atoms=300, seed_tag=4242):
    """A chain molecule far larger than a tile residue slot (CHARMM-Drude style: every heavy atom is followed by its
    Drude particle and two hydrogens): 4 * heavy_atoms particles in ONE residue, as OpenMM's getMolecules() reports a
    protein or a polymer (openmmapi/src/DrudeTGNHIntegrator.cpp:121-141)."""
    masses, pairs = [], []
    for k in range(heavy_atoms):
        masses += [12.011 - 0.4 if k % 3 else 14.007 - 0.4, 0.4, 1.008, 1.008]
        pairs.append([4 * k + 1, 4 * k])
    rng = np.random.Generator(np.random.Philox(key=[SEED, seed_tag]))
    off = np.cumsum(rng.uniform(-0.08, 0.08, (4 * heavy_atoms, 3)), axis=0)
    for d, p in pairs:
        off[d] = off[p]
    return Template(f"polymer{heavy_atoms}", np.array(masses), np.array(pairs, np.int32), off - off[0])


def polymer_in_water(waters=1500, polymers=(300, 700), G=2, **kw):
    """Waters with a few big molecules in between (first, middle, last positions in particle order)."""
    templates = [WATER4] + [polymer_template(n, 4242 + n) for n in polymers]
    types = [1] + [0] * (waters // 2)
    for k in range(1, len(polymers)):
        types += [k + 1] + [0] * (waters // (2 * max(1, len(polymers) - 1)))
    types += [len(polymers)]                       # a big molecule last as well
    types = np.array(types, np.int32)
    return build(templates, types, np.arange(len(types)) % G, G, **kw)


def _ionic_templates():
    """[BMIM]+ (25 atoms, 10 heavy atoms carrying Drudes -> 35 particles) and [BF4]- (5 atoms, all
    polarizable -> 10 particles); masses are element masses minus 0.4 for Drude carriers."""
    heavy = [14.007, 12.011, 14.007, 12.011, 12.011, 12.011, 12.011, 12.011, 12.011, 12.011]
    masses, pairs = [], []
    for m in heavy:
        masses += [m - 0.4, 0.4]
        pairs.append([len(masses) - 1, len(masses) - 2])
    masses += [1.008] * 15
    rng = np.random.Generator(np.random.Philox(key=[SEED, 777]))
    off = rng.uniform(-0.3, 0.3, (35, 3))
    for d, p in pairs:
        off[d] = off[p]
    bmim = Template("bmim", np.array(masses), np.array(pairs, np.int32), off)
    am, ap = [], []
    for m in [10.811, 18.998, 18.998, 18.998, 18.998]:
        am += [m - 0.4, 0.4]
        ap.append([len(am) - 1, len(am) - 2])
    aoff = rng.uniform(-0.14, 0.14, (10, 3))
    for d, p in ap:
        aoff[d] = aoff[p]
    bf4 = Template("bf4", np.array(am), np.array(ap, np.int32), aoff)
    return bmim, bf4


def build(templates, mol_types, mol_groups, num_temp_groups, *, seed=SEED, first_molecule=0,
          density=33.4, box_molecules=None, drude_sigma=0.005, force_sigma=200.0,
          k_spring=100000 * 4.184, temperature=300.0, pair_force="common", cold_drudes=False,
          quantize_masses=False, **params):
    """Assemble a system from per-molecule template ids / temperature groups.

    mol_types[j], mol_groups[j] describe global molecule first_molecule + j.  Random data for a
    molecule depends only on (seed, its global index), never on the shard it is generated in.

    pair_force selects the fixed synthetic force on Drude pairs:
      "common"         (default) the pair's random total force is split by mass fraction, i.e. both members get
                       the same acceleration and the Drude displacement is force-free: pairs drift at the cold
                       (T_drude) relative velocity and meet the hard wall as the rare event it is in real MD
                       (measured ~1 % of pairs per step in steady state, Drude temperature on target);
      "frozen_spring"  SURVEY.md 8d's original choice: the Drude spring -k (x_d - x_p) evaluated once at t = 0 and
                       frozen.  A constant non-restoring force of ~2000 kJ/mol/nm drives EVERY pair into the wall
                       on every step (measured: hit fraction 1.000 after ~30 steps, Drude temperature 318 K for a
                       1 K target); kept as the hard-wall stress workload;
      "none"           zero force on pair members (for tests that recompute harmonic forces every step).

    cold_drudes: draw each pair's centre-of-mass velocity at `temperature` and its relative velocity at
    drude_temperature (an equilibrated dual-thermostat state) instead of independent thermal velocities.
    quantize_masses: replace every mass m by 1 / float32(1 / m), so that the fp32 inverse masses of the device
    layout and the fp64 masses given to a CPU oracle describe exactly the same system.
    """
    mol_types = np.asarray(mol_types, np.int32)
    mol_groups = np.asarray(mol_groups, np.int32)
    nmol = len(mol_types)
    sizes = np.array([t.size for t in templates], np.int64)
    kmax = int(sizes.max())
    msize = sizes[mol_types]
    start = np.concatenate([[0], np.cumsum(msize)])
    N = int(start[-1])

    # per-particle static tables
    res_id = np.repeat(np.arange(nmol, dtype=np.int32), msize)
    local = np.arange(N, dtype=np.int64) - np.repeat(start[:-1], msize)
    ptype = np.repeat(mol_types, msize)
    mass_tab = np.zeros((len(templates), kmax)); off_tab = np.zeros((len(templates), kmax, 3))
    for ti, t in enumerate(templates):
        mass_tab[ti, :t.size] = t.masses
        off_tab[ti, :t.size] = t.offsets
    masses = mass_tab[ptype, local]
    temp_group = np.repeat(mol_groups, msize).astype(np.int32)

    pd, pp, cons = [], [], []
    for ti, t in enumerate(templates):
        sel = np.nonzero(mol_types == ti)[0]
        if len(sel) == 0:
            continue
        base = start[sel]
        for d, p in t.pairs:
            pd.append(base + d); pp.append(base + p)
        for a, b in t.constraints:
            cons.append(np.stack([base + a, base + b], 1))
    if pd:
        pair_drude = np.concatenate(pd); pair_parent = np.concatenate(pp)
        order = np.argsort(pair_drude, kind="stable")
        pair_drude = pair_drude[order].astype(np.int32); pair_parent = pair_parent[order].astype(np.int32)
    else:
        pair_drude = np.zeros(0, np.int32); pair_parent = np.zeros(0, np.int32)
    constraints = np.concatenate(cons).astype(np.int32) if cons else np.zeros((0, 2), np.int32)

    # per-molecule random data: one Philox stream per chunk of _CHUNK global molecules
    total = box_molecules if box_molecules is not None else first_molecule + nmol
    box = (total / density) ** (1.0 / 3.0)
    u_center = np.empty((nmol, 3)); n_mol = np.empty((nmol, kmax, 9))
    g0 = first_molecule
    j = 0
    if kmax > 64:
        # big molecules (polymers): one Philox stream per global molecule instead of a [_CHUNK, kmax, 9] block per chunk
        for j in range(nmol):
            rng = np.random.Generator(np.random.Philox(key=[seed, (1 << 40) + g0 + j]))
            u_center[j] = rng.random(3)
            n_mol[j, :msize[j]] = rng.standard_normal((int(msize[j]), 9))
            n_mol[j, msize[j]:] = 0.0
        j = nmol
    while j < nmol:
        chunk = (g0 + j) // _CHUNK
        lo = chunk * _CHUNK
        cnt = min(nmol - j, lo + _CHUNK - (g0 + j))
        rng = np.random.Generator(np.random.Philox(key=[seed, chunk]))
        uc = rng.random((_CHUNK, 3)); nm = rng.standard_normal((_CHUNK, kmax, 9))
        a = g0 + j - lo
        u_center[j:j + cnt] = uc[a:a + cnt]; n_mol[j:j + cnt] = nm[a:a + cnt]
        j += cnt
    nrm = n_mol[res_id, local]                        # [N, 9]: 0-2 drude displacement, 3-5 velocity, 6-8 force

    positions = u_center[res_id] * box + off_tab[ptype, local]
    positions[pair_drude] = positions[pair_parent] + drude_sigma * nrm[pair_drude, 0:3]
    with np.errstate(divide="ignore", invalid="ignore"):
        vsig = np.where(masses > 0, np.sqrt(BOLTZ * temperature / np.where(masses > 0, masses, 1.0)), 0.0)
    if quantize_masses:
        with np.errstate(divide="ignore"):
            w32 = np.where(masses > 0, 1.0 / np.where(masses > 0, masses, 1.0), 0.0).astype(np.float32).astype(np.float64)
            masses = np.where(w32 > 0, 1.0 / np.where(w32 > 0, w32, 1.0), 0.0)
    velocities = vsig[:, None] * nrm[:, 3:6]
    if cold_drudes and len(pair_drude):
        md, mp = masses[pair_drude], masses[pair_parent]
        mt, mu = md + mp, md * mp / (md + mp)
        t_d = params.get("drude_temperature", 1.0)
        vcm = np.sqrt(BOLTZ * temperature / mt)[:, None] * nrm[pair_parent, 3:6]
        vrel = np.sqrt(BOLTZ * t_d / mu)[:, None] * nrm[pair_drude, 3:6]          # v_parent - v_drude
        velocities[pair_drude] = vcm - vrel * (mp / mt)[:, None]
        velocities[pair_parent] = vcm + vrel * (md / mt)[:, None]
    positions = _f32(positions); velocities = _f32(velocities)
    forces = force_sigma * nrm[:, 6:9]
    forces[masses == 0] = 0.0
    ks = np.full(len(pair_drude), float(k_spring))
    if len(pair_drude):
        if pair_force == "frozen_spring":
            fd = -ks[:, None] * (positions[pair_drude] - positions[pair_parent])
            forces[pair_drude] = fd
            forces[pair_parent] = -fd
        elif pair_force == "common":
            ftot = forces[pair_parent].copy()
            mt = masses[pair_drude] + masses[pair_parent]
            forces[pair_drude] = ftot * (masses[pair_drude] / mt)[:, None]
            forces[pair_parent] = ftot * (masses[pair_parent] / mt)[:, None]
        elif pair_force == "none":
            forces[pair_drude] = 0.0; forces[pair_parent] = 0.0
        else:
            raise ValueError(f"unknown pair_force {pair_force!r}")
    forces = _f32(forces)
    return DrudeSystem(masses=masses, pair_drude=pair_drude, pair_parent=pair_parent, temp_group=temp_group,
                       res_id=res_id, positions=positions, velocities=velocities, forces=forces, k_spring=ks,
                       constraints=constraints, num_temp_groups=int(num_temp_groups), num_residues=nmol,
                       temperature=temperature, **params)


def tile(block, reps):
    """The index tables of `reps` copies of `block` laid end to end (masses, pairs, temperature groups, residue ids, constraints);
    state arrays stay those of ONE block (callers tile them where they live, e.g. on the device).  For benchmark systems far larger
    than the host generator can produce in reasonable time."""
    n, r = block.num_particles, block.num_residues
    off = (np.arange(reps, dtype=np.int64) * n)[:, None]
    cons = block.constraints
    if len(cons):
        cons = (cons[None, :, :].astype(np.int64) + off[:, :, None]).reshape(-1, 2).astype(np.int32)
    return DrudeSystem(
        masses=np.tile(block.masses, reps), pair_drude=(block.pair_drude[None, :] + off).reshape(-1).astype(np.int32),
        pair_parent=(block.pair_parent[None, :] + off).reshape(-1).astype(np.int32), temp_group=np.tile(block.temp_group, reps),
        res_id=(block.res_id[None, :].astype(np.int64) + (np.arange(reps, dtype=np.int64) * r)[:, None]).reshape(-1).astype(np.int32),
        positions=block.positions, velocities=block.velocities, forces=block.forces, k_spring=np.tile(block.k_spring, reps), constraints=cons,
        num_temp_groups=block.num_temp_groups, num_residues=r * reps, temperature=block.temperature, coupling_time=block.coupling_time,
        drude_temperature=block.drude_temperature, drude_coupling_time=block.drude_coupling_time, step_size=block.step_size,
        drude_steps=block.drude_steps, num_nh_chains=block.num_nh_chains, use_drude_nh_chains=block.use_drude_nh_chains,
        use_com_temp_group=block.use_com_temp_group, max_drude_distance=block.max_drude_distance)


def water_box(num_molecules, num_temp_groups=4, *, first_molecule=0, box_molecules=None, **kw):
    """C4 / C5: 4-particle molecules [parent 15.6, drude 0.4, a 1.0, b 1.0]; group of molecule k = k mod G."""
    gidx = np.arange(first_molecule, first_molecule + num_molecules)
    return build([WATER4], np.zeros(num_molecules, np.int32), gidx % num_temp_groups, num_temp_groups,
                 first_molecule=first_molecule, box_molecules=box_molecules, **kw)


def swm4_box(num_molecules=10000, **kw):
    """C2: SWM4-NDP waters incl. the massless M site (5 particles), 1 temperature group."""
    kw.setdefault("num_nh_chains", 1)
    return build([SWM4], np.zeros(num_molecules, np.int32), np.zeros(num_molecules, np.int32), 1, **kw)


def nacl_box(**kw):
    """C1 (example/nacl_1m_pos.pdb shape): 492 SWM4 waters + 10 Na+ + 10 Cl-, N = 2500, P = 512, G = 2."""
    types = np.array([0] * 492 + [1] * 10 + [2] * 10, np.int32)
    groups = np.array([0] * 492 + [1] * 20, np.int32)
    kw.setdefault("num_nh_chains", 1)
    kw.setdefault("use_drude_nh_chains", False)
    return build([SWM4, SOD, CLA], types, groups, 2, **kw)


def ionic_liquid(num_ion_pairs=1000, **kw):
    """C3: [BMIM][BF4]-like, 35 + 10 particles per ion pair, G = 3 (cations, anions, spare solvent group)."""
    bmim, bf4 = _ionic_templates()
    types = np.tile(np.array([0, 1], np.int32), num_ion_pairs)
    kw.setdefault("density", 3.0)
    return build([bmim, bf4], types, types.copy(), 3, **kw)


def single_pair():
    """The 2-particle system of testSinglePair (platforms/reference/tests/TestReferenceDrudeTGNHIntegrator.cpp:54-109)."""
    k = ONE_4PI_EPS0 * 1.5
    return DrudeSystem(masses=np.array([1.0, 0.1]), pair_drude=np.array([1], np.int32), pair_parent=np.array([0], np.int32),
                       temp_group=np.zeros(2, np.int32), res_id=np.zeros(2, np.int32),
                       positions=np.array([[0, 0, 0], [0, 0, 0.01]], np.float64),
                       velocities=np.array([[1, 0, 0], [1, 0, 0.01]], np.float64),
                       forces=np.zeros((2, 3)), k_spring=np.array([k]), constraints=np.zeros((0, 2), np.int32),
                       num_temp_groups=1, num_residues=1, temperature=300.0, coupling_time=0.1,
                       drude_temperature=10.0, drude_coupling_time=0.005, step_size=0.003, drude_steps=20,
                       num_nh_chains=2, use_drude_nh_chains=False, use_com_temp_group=True, max_drude_distance=0.05)
