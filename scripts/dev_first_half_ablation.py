"""C4: first-half kernel time with the hard wall and the COM group switched off (what the per-residue COM velocity loop and
the partner recomputation for the wall test cost)."""
import sys
import numpy as np, torch
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
from openmm_drudenose_b200 import capi, synth
from util import DeviceState
dev = torch.device("cuda:0")
for name, kw in (("COM group + hard wall", {}), ("COM group, no wall", dict(max_drude_distance=0.0)), ("no COM group, hard wall", dict(use_com_temp_group=False)),
                 ("no COM group, no wall", dict(use_com_temp_group=False, max_drude_distance=0.0))):
    s = synth.water_box(2_500_000, 4, **kw)
    st = DeviceState(s, dev)
    h = capi.Handle(s)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        h.step(*st.ptrs, nsteps=5, stream=stream.cuda_stream); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); h.step(*st.ptrs, nsteps=40, stream=stream.cuda_stream); e1.record(stream); torch.cuda.synchronize()
        h.set_profiling(True); h.step(*st.ptrs, nsteps=40, stream=stream.cuda_stream); torch.cuda.synchronize()
        prof = h.profile()
    print(f"{name}: {e0.elapsed_time(e1) / 40 * 1e3:.1f} us/step; first half {prof['half1'][0] / prof['half1'][1] * 1e3:.1f} us, second half {prof['half2'][0] / prof['half2'][1] * 1e3:.1f} us")
    h.close(); del st; torch.cuda.empty_cache()
