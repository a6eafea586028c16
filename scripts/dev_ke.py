import sys, os
R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, R+'/tests')
import numpy as np, torch
from openmm_drudenose_b200 import synth, capi
from oracle import oracle as O
from util import DeviceState
dev=torch.device('cuda:0')
np.set_printoptions(linewidth=200, precision=3)
for mol in (25000, 250000):
    s = synth.water_box(mol, 4, quantize_masses=True, cold_drudes=True, drude_sigma=1.4e-4, force_sigma=2.0, max_drude_distance=2.0, use_drude_nh_chains=False)
    st = DeviceState(s, dev); h = capi.Handle(s); o = O.Oracle(s, O.TG)
    ke_g = h.compute_kinetic_energies(st.velm.data_ptr()); ke_r = o.compute_ke2(s.velocities.copy())
    print(mol, "KE kernel signed rel err:", (ke_g-ke_r)/ke_r)
    # second-half kernel: one half2 with zero forces on identical state -> ke consumed by chain = KE of same velocities
    st.force.zero_()
    h2 = capi.Handle(s); h2.half1(*st.ptrs)   # scaled+drifted; now compare B's KE against oracle KE of the device velocities
    vdev = st.vel()
    h2.half2(st.velm.data_ptr(), st.force.data_ptr(), capi.HALF2_DEFER_SCALE)
    ke_b = h2.kinetic_energies(); ke_rb = o.compute_ke2(vdev.copy())
    print(mol, "BU kernel signed rel err:", (ke_b-ke_rb)/ke_rb)
