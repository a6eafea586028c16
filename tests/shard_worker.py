"""torchrun worker for tests/test_sharded.py: the particle-range sharded TGNH step (exchange of the double[G+2]
kinetic-energy partial sums through peer-mapped inboxes, or NCCL with TGNH_P2P=0) against the same system on one GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist

from openmm_drudenose_b200 import capi, synth
from util import DeviceState

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ids = [capi.Comm.unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
comm = capi.Comm(ids[0], world, rank, local)

MOL, G, STEPS = 40000, 4, 25
per = MOL // world
shard = synth.water_box(per, G, first_molecule=rank * per, box_molecules=MOL, quantize_masses=True)
st = DeviceState(shard, dev)
h = capi.Handle(shard, device=local, comm=comm)
kind = h.exchange_kind
assert kind == (1 if os.environ.get("TGNH_P2P") == "0" else int(os.environ.get("EXPECT_EXCHANGE", kind))), kind
h.step(*st.ptrs, nsteps=STEPS - 5)
# the OpenMM-facing call sequence, a kinetic-energy query and a deferred-scale flush in between (all collective)
for i in range(5):
    h.half1(*st.ptrs)
    h.half2(st.velm.data_ptr(), st.force.data_ptr(), capi.HALF2_DEFER_SCALE if i == 2 else capi.HALF2_DEFAULT)
    if i == 2:
        h.flush(st.velm.data_ptr())
    if i == 3:
        ke_now = h.compute_kinetic_energies(st.velm.data_ptr())
torch.cuda.synchronize()
ke, vs, ed = h.kinetic_energies(), h.vscale(), h.chain_state()[1]
dof = h.thermostat_params()[0]
vel = st.vel()

# every rank holds the same global thermostat state
gathered = [None] * world
dist.all_gather_object(gathered, (ke.tolist(), vs.tolist(), ed.tolist(), dof.tolist()))
if rank == 0:
    for other in gathered[1:]:
        assert other == gathered[0], "ranks disagree on the thermostat state"
    whole = synth.water_box(MOL, G, quantize_masses=True)
    st1 = DeviceState(whole, dev)
    h1 = capi.Handle(whole, device=local)
    h1.step(*st1.ptrs, nsteps=STEPS - 5)
    for i in range(5):
        h1.half1(*st1.ptrs)
        h1.half2(st1.velm.data_ptr(), st1.force.data_ptr(), capi.HALF2_DEFER_SCALE if i == 2 else capi.HALF2_DEFAULT)
        if i == 2:
            h1.flush(st1.velm.data_ptr())
        if i == 3:
            np.testing.assert_allclose(ke_now, h1.compute_kinetic_energies(st1.velm.data_ptr()), rtol=1e-9)
    # The two runs differ in the order of the energy sums (last bits of the scale factors); over 25 steps that flips the fp32 rounding of
    # a few velocities by one ulp, which the energies see at the 1e-10 level: bounds are those, far inside the 1e-6 of the parity tests.
    np.testing.assert_allclose(dof, h1.thermostat_params()[0], rtol=1e-12)         # DOF tables summed over ranks at create (order of the COM-share sum differs)
    np.testing.assert_allclose(ke, h1.kinetic_energies(), rtol=1e-9)
    np.testing.assert_allclose(vs, h1.vscale(), rtol=1e-10)
    np.testing.assert_allclose(ed, h1.chain_state()[1], rtol=1e-8, atol=1e-10)
    n = shard.num_particles
    np.testing.assert_allclose(vel, st1.vel()[:n], rtol=0, atol=1e-6)
    print("SHARD_OK", world, "exchange", kind)
h.close()
comm.close()
dist.destroy_process_group()
