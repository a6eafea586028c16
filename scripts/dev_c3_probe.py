import sys, os
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
import bench
from openmm_drudenose_b200 import capi, synth
dev = torch.device("cuda:0")
s = synth.ionic_liquid(1000)
padded, _, (velm, posq, force) = bench.device_buffers(torch, s, dev)
h = capi.Handle(s, padded=padded)
ptrs = [velm.data_ptr(), posq.data_ptr(), force.data_ptr()]
h.step(*ptrs, 10); torch.cuda.synchronize()
if os.environ.get("WARM2"): h.step(*ptrs, 10); torch.cuda.synchronize()
import time; t0=time.perf_counter(); h.step(*ptrs, 10); t1=time.perf_counter(); torch.cuda.synchronize(); t2=time.perf_counter(); print("host time of the launch call %.2f ms, until sync %.2f ms" % ((t1-t0)*1e3, (t2-t0)*1e3))
ts = []
for rep in range(int(os.environ.get("REPS", "8"))):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); h.step(*ptrs, int(os.environ.get("BLOCK", "100"))); e1.record(); torch.cuda.synchronize()
    ts.append(round(e0.elapsed_time(e1) * 1000 / int(os.environ.get("BLOCK", "100")), 1))
print(os.environ.get("TGNH_LIB", "default")[-20:], "fuse", os.environ.get("TGNH_FUSE_CHAIN"), "us/step per 100-step block:", ts, "ke2", h.kinetic_energies(), "vscale", h.vscale(), "M", s.num_nh_chains, s.use_drude_nh_chains)
