"""The driving loop of the reference's example/nacl_tg.py (lines 30-75) on this repo's B200 build.

OpenMM's app layer (PDB reader, force field, reporters) is not available here, so the system comes from the synthetic
generator (the shape of example/nacl_1m_pos.pdb: 492 SWM4 waters + 10 Na+ + 10 Cl-, 2500 particles, 512 Drude pairs)
and the forces from the shim's host model (harmonic Drude springs); everything from `DrudeTGNHIntegrator(...)` down is the
real stack: integrator class -> kernel interface -> libtgnh.so -> sm_100a kernels.  With OpenMM installed the same lines run
with `from simtk.openmm import *` in place of `mm = dp.shim`.

    python example/nacl_tg_b200.py [steps]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "plugin", "python"))
import drudetgnhplugin as dp                      # noqa: E402
from openmm_drudenose_b200 import synth           # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
s = synth.nacl_box(pair_force="none", cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0)
mm = dp.shim

system = mm.System()
for m in s.masses:
    system.addParticle(float(m))
drude = mm.DrudeForce()
for d, p in zip(s.pair_drude, s.pair_parent):
    drude.addParticle(int(d), int(p), -1, -1, -1, -1.0, 1.0, 1.0, 1.0)
system.addForce(drude)
bonds = mm.BondForce()
for i in range(1, s.num_particles):
    if s.res_id[i] == s.res_id[i - 1]:
        bonds.addBond(i - 1, i)
system.addForce(bonds)

# example/nacl_tg.py:47-57: 300 K / 1 K, 0.1 ps / 0.005 ps, 1 fs, 20 Drude steps, hard wall at 0.02 nm, ions in their own group
integrator = dp.DrudeTGNHIntegrator(300.0, 0.1, 1.0, 0.005, 0.001, 20, 1, 0, 1)
integrator.setMaxDrudeDistance(0.02)
integrator.addTempGroup()          # group 0: water
integrator.addTempGroup()          # group 1: ions
for g in s.temp_group:
    integrator.addParticleTempGroup(int(g))

platform = mm.Platform.getPlatformByName("CUDA")
context = mm.Context(system, integrator, platform, {"Precision": "mixed"})      # example/nacl_tg.py:60
context.setPositions(s.positions)
context.setVelocities(s.velocities)
context.setForceModel(np.zeros_like(s.forces), [int(x) for x in s.pair_drude], [int(x) for x in s.pair_parent], [float(k) for k in s.k_spring])

kB = 0.0083144626
massive, pairs = int(np.count_nonzero(s.masses)), len(s.pair_drude)
target = 0.5 * kB * ((3 * (massive - pairs) - 3) * 300.0 + 3 * pairs * 1.0)         # atoms at 300 K, Drude pairs' relative motion at 1 K
for block in range(5):
    integrator.step(steps // 5)
    state = context.getState(getEnergy=True, getPositions=True)
    d = state.getPositions()[s.pair_drude] - state.getPositions()[s.pair_parent]
    print(f"step {(block + 1) * (steps // 5):6d}  KE = {state.getKineticEnergy():10.2f} kJ/mol (dual-thermostat target {target:.2f})  "
          f"max Drude distance = {np.linalg.norm(d, axis=1).max():.4f} nm")
