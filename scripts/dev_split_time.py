"""C4: per-step time of the constraint-split call sequence (kick / drift / kick / thermostat), per-kind launch times."""
import sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from util import DeviceState
dev = torch.device("cuda:0")
s = synth.water_box(2_500_000, 4)
st = DeviceState(s, dev)
h = capi.Handle(s)
delta = torch.zeros_like(st.velm)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); sp = stream.cuda_stream
def run(n):
    for _ in range(n):
        h.half1_kick(st.velm.data_ptr(), st.force.data_ptr(), delta.data_ptr(), stream=sp)
        h.half1_drift(st.velm.data_ptr(), st.posq.data_ptr(), delta.data_ptr(), stream=sp)
        h.half2(st.velm.data_ptr(), st.force.data_ptr(), capi.HALF2_KICK_ONLY, stream=sp)
        h.thermostat(st.velm.data_ptr(), stream=sp)
run(5); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(50); e1.record(); torch.cuda.synchronize()
print("constraint split, immediate scaling: us/step", round(e0.elapsed_time(e1) / 50 * 1e3, 1))
h.set_profiling(True); run(20); torch.cuda.synchronize(); print({k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in h.profile().items()})
