import sys, os, json
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np, torch
from openmm_drudenose_b200 import synth, capi
dev=torch.device('cuda:0')
def run(label, **kw):
    s = synth.water_box(2_500_000, 4, **kw)
    n=s.num_particles; padded=((n+31)//32)*32
    velm=torch.zeros((padded,4),dtype=torch.float32); velm[:n]=torch.from_numpy(s.velm_f32())
    posq=torch.zeros((padded,4),dtype=torch.float32); posq[:n]=torch.from_numpy(s.posq_f32())
    force=torch.zeros((3,padded),dtype=torch.float32); force[:,:n]=torch.from_numpy(np.ascontiguousarray(s.forces.T,np.float32))
    velm,posq,force=velm.to(dev),posq.to(dev),force.to(dev)
    h=capi.Handle(s,padded=padded)
    h.step(velm.data_ptr(),posq.data_ptr(),force.data_ptr(),5)
    torch.cuda.synchronize(); h.set_profiling(True)
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); h.step(velm.data_ptr(),posq.data_ptr(),force.data_ptr(),50); e1.record(); torch.cuda.synchronize()
    p=h.profile(); print(f"{label:28s} step {e0.elapsed_time(e1)/50*1000:7.1f} us  A {p['half1'][0]/p['half1'][1]*1000:6.1f}  B {p['half2'][0]/p['half2'][1]*1000:6.1f}")
    h.close()
run("default (COM, wall, M3)")
run("no COM group", use_com_temp_group=False)
run("no hard wall", max_drude_distance=0.0)
run("M=1 no drude chain", num_nh_chains=1, use_drude_nh_chains=False)
