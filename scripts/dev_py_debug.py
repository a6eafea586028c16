import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R + "/tests"); sys.path.insert(0, R + "/plugin/python")
import numpy as np
from openmm_drudenose_b200 import synth
from oracle import oracle as O
import drudetgnhplugin as dp
import test_plugin as T
from plugin_driver import PluginSim
s = synth.water_box(1500, 3, quantize_masses=True)
s.positions = (s.positions - s.positions.mean(0)).astype(np.float32).astype(np.float64)
s.velocities = s.velocities.astype(np.float32).astype(np.float64)
for steps in (1, 2, 20):
    p, v, ke, ext = T._script_style_run(dp, s, "single", steps, False)
    sim = PluginSim(s, force_model=0)
    pa, va, fa = s.positions.copy(), s.velocities.copy(), ext.copy()
    sim.step(pa, va, fa, steps, None)
    o = O.Oracle(s, O.TG); pb, vb, fb = s.positions.copy(), s.velocities.copy(), ext.copy(); o.step(pb, vb, fb, steps, 0, None, None)
    print(steps, "script vs oracle", np.abs(v - vb).max(), "driver vs oracle", np.abs(va - vb).max(), "script vs driver", np.abs(v - va).max(), "ke", ke, sim.kinetic_energy(), o.ke_sum)
