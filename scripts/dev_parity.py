import sys, time
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np, torch
from openmm_drudenose_b200 import synth, capi
from oracle import oracle as O
from util import DeviceState, rel_err
dev = torch.device('cuda:0')
print(capi.lib().tgnh_build_info().decode())
for name, s in [("water G=4", synth.water_box(5000, 4, drude_sigma=0.001)),
                ("nacl", synth.nacl_box(drude_sigma=0.001)),
                ("ionic", synth.ionic_liquid(100, drude_sigma=0.001)),
                ("water noCOM G=1", synth.water_box(3000, 1, use_com_temp_group=False, drude_sigma=0.001))]:
    print("==", name, s.num_particles)
    st = DeviceState(s, dev)
    h = capi.Handle(s)
    o = O.Oracle(s, O.TG)
    ke_gpu = h.compute_kinetic_energies(st.velm.data_ptr())
    ke_cpu = o.compute_ke2(s.velocities.copy())
    print(" ke2 rel", np.max(np.abs(ke_gpu-ke_cpu)/np.maximum(np.abs(ke_cpu),1e-300)), ke_cpu)
    print(" dof", np.abs(np.array(h.thermostat_params()[0]) - o.thermostat_params()[0]).max(), np.abs(h.thermostat_params()[2]-o.thermostat_params()[2]).max())
    p, v, f = s.positions.copy(), s.velocities.copy(), s.forces.copy()
    for nst in (1, 1, 3, 20):
        h.step(*st.ptrs, nsteps=nst); torch.cuda.synchronize()
        o.step(p, v, f, nst)
        print(f" after +{nst}: v rel {rel_err(st.vel(), v):.3e} x rel {rel_err(st.pos(), p):.3e}",
              "ke2 rel", np.max(np.abs(h.kinetic_energies()-o.ke2)/np.maximum(np.abs(o.ke2),1e-300)),
              "vscale", np.abs(h.vscale()-o.vscale).max(), "etadot", np.abs(h.chain_state()[1]-o.chain_state()[1]).max()/np.abs(o.chain_state()[1]).max())
    print(" KESum", h.kinetic_energy(), o.ke_sum, "launches", h.launch_count)
    assert np.array_equal(st.posq[:s.num_particles,3].cpu().numpy(), st.charges)
