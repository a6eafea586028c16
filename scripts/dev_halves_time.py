"""C4: per-step time of the OpenMM-facing call sequence (tgnh_half1 + tgnh_half2 with immediate / deferred scaling)
against the fused tgnh_step, with per-kind launch times."""
import sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from util import DeviceState
dev = torch.device("cuda:0")
s = synth.water_box(2_500_000, 4)
st = DeviceState(s, dev)
h = capi.Handle(s)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); sp = stream.cuda_stream
def timed(fn, n=50):
    fn(5); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(n); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def halves(flag):
    def run(n):
        for _ in range(n):
            h.half1(*st.ptrs, stream=sp); h.half2(st.velm.data_ptr(), st.force.data_ptr(), flag, stream=sp)
        h.flush(st.velm.data_ptr(), stream=sp)
    return run
print("fused tgnh_step        us/step", round(timed(lambda n: h.step(*st.ptrs, nsteps=n, stream=sp)), 1))
print("half1+half2 (deferred) us/step", round(timed(halves(capi.HALF2_DEFER_SCALE)), 1))
print("half1+half2 (default)  us/step", round(timed(halves(capi.HALF2_DEFAULT)), 1))
h.set_profiling(True); halves(capi.HALF2_DEFAULT)(20); torch.cuda.synchronize(); print({k: (round(v[0] / max(v[1], 1) * 1e3, 1), v[1]) for k, v in h.profile().items()})
