"""Generates tests/golden/*.npz — known-answer vectors for the TGNH step.

The reference ships no golden vectors, so these are produced by the CPU oracle
(oracle/, the fp64 restatement of the reference algorithm) on seeded inputs of the deterministic generator
and committed, together with this script.  Re-run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from openmm_drudenose_b200 import synth  # noqa: E402
from oracle import oracle as O           # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # 1. full steps, temperature groups + COM thermostat + Drude chains + hard wall, fixed forces
    mol, groups, steps = 256, 3, 5
    s = synth.water_box(mol, groups, quantize_masses=True)
    o = O.Oracle(s, O.TG)
    p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    ke0 = o.compute_ke2(v.copy())
    o.step(p, v, f, steps)
    eta, ed, edd = o.chain_state()
    dof, nkbt, q = o.thermostat_params()
    np.savez_compressed(os.path.join(HERE, "tgnh_golden.npz"), molecules=mol, groups=groups, steps=steps,
                        positions0=s.positions, velocities0=s.velocities, forces=s.forces, ke2_initial=ke0,
                        positions=p, velocities=v, ke2=o.ke2, vscale=o.vscale, ke_sum=o.ke_sum,
                        eta=eta, eta_dot=ed, eta_dot_dot=edd, dof=dof, nkbt=nkbt, eta_mass=q)

    # 2. overlap domain (G = 1, no COM group, Drude chains): both oracle layers, harmonic Drude springs
    s2 = synth.water_box(128, 1, use_com_temp_group=False, quantize_masses=True, pair_force="none", cold_drudes=True,
                         drude_sigma=1.4e-4, force_sigma=0.0)
    out = {}
    for name, which in (("tg", O.TG), ("ref", O.REF)):
        o2 = O.Oracle(s2, which)
        p2, v2 = s2.positions.copy(), s2.velocities.copy()
        f2 = O.harmonic_forces(s2, p2)
        o2.step(p2, v2, f2, 50, O.FORCE_HARMONIC, None, s2.k_spring)
        out[f"positions_{name}"] = p2; out[f"velocities_{name}"] = v2
        out[f"vscale_{name}"] = o2.vscale; out[f"eta_dot_{name}"] = o2.chain_state()[1]
    np.savez_compressed(os.path.join(HERE, "overlap_golden.npz"), molecules=128, steps=50, **out)

    # 3. the chain alone: scale factors and chain variables for prescribed kinetic energies (known-answer test of
    #    CudaDrudeTGNHKernels.cpp:559-642); the velocities are scaled copies so KE is controlled exactly
    s3 = synth.water_box(64, 2, quantize_masses=True, num_nh_chains=4)
    o3 = O.Oracle(s3, O.TG)
    v3 = s3.velocities.copy()
    rec = []
    for i in range(6):
        v3 *= (1.0 + 0.25 * i)
        o3.propagate_nh_chain(v3)
        rec.append(np.concatenate([o3.ke2, o3.vscale, o3.chain_state()[1].ravel()]))
    np.savez_compressed(os.path.join(HERE, "chain_golden.npz"), molecules=64, groups=2, chains=4, records=np.array(rec))
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
