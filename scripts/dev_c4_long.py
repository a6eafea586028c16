"""C4 (10 M particles, the bench workload) over 5000 steps with the bench's fixed forces: do the thermostats hold their targets at full
size, how far do Drude pairs stretch (hard wall at 0.02 nm), is anything non-finite?  Prints one line per 500 steps."""
import os, sys
import numpy as np, torch
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import bench
from openmm_drudenose_b200 import capi, synth
from util import DeviceState
dev = torch.device("cuda:0")
s = bench.make_system("c4", 0, 1)
st = DeviceState(s, dev)
h = capi.Handle(s, padded=st.padded)
dof, nkbt, _ = h.thermostat_params()
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); sp = stream.cuda_stream
n = s.num_particles
done = 0
for block in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.step(*st.ptrs, nsteps=500, stream=sp); e1.record(); torch.cuda.synchronize()
    done += 500
    ke2 = h.kinetic_energies()
    T = np.where(nkbt > 0, ke2 / np.where(nkbt > 0, nkbt, 1.0), 0.0)          # in units of the group's own target: KE2_g / (N_g k T_g)
    x = st.posq[:n, :3]
    d = (x[1::4] - x[0::4]).norm(dim=1)                                        # molecules are [parent, drude, a, b]
    finite = bool(torch.isfinite(st.velm[:n]).all() and torch.isfinite(x).all())
    print(f"steps {done:5d}  {e0.elapsed_time(e1) / 500 * 1e3:6.1f} us/step  KE2/(N k T_target) per thermostat {np.round(T, 4).tolist()}  "
          f"Drude-parent distance max {float(d.max()):.5f} mean {float(d.mean()):.5f} nm  finite {finite}  vscale {np.round(h.vscale(), 6).tolist()}", flush=True)
h.close()
