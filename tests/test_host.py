"""CPU tests of the host side: generator, C-ABI library loading / validation, shard logic (gloo, world_size 2)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

from openmm_drudenose_b200 import capi, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_generator_is_shard_invariant():
    whole = synth.water_box(70000, 4)
    lo = synth.water_box(30000, 4, first_molecule=0, box_molecules=70000)
    hi = synth.water_box(40000, 4, first_molecule=30000, box_molecules=70000)
    n = lo.num_particles
    for name in ("positions", "velocities", "forces", "masses", "temp_group"):
        np.testing.assert_array_equal(getattr(whole, name)[:n], getattr(lo, name))
        np.testing.assert_array_equal(getattr(whole, name)[n:], getattr(hi, name))
    assert whole.num_pairs == lo.num_pairs + hi.num_pairs


def test_config_shapes():
    c1 = synth.nacl_box()
    assert (c1.num_particles, c1.num_pairs, c1.num_residues, c1.num_temp_groups) == (2500, 512, 512, 2)
    assert (c1.masses == 0).sum() == 492 and len(c1.constraints) == 1476
    c2 = synth.swm4_box(10000)
    assert (c2.num_particles, c2.num_pairs, len(c2.constraints)) == (50000, 10000, 30000)
    c3 = synth.ionic_liquid(1000)
    assert (c3.num_particles, c3.num_pairs, c3.num_residues, c3.num_temp_groups) == (45000, 15000, 2000, 3)
    v = c1.velm_f32(); f = c1.force_i64_soa(2528)
    assert v.shape == (2500, 4) and v.dtype == np.float32 and f.shape == (3, 2528) and f.dtype == np.int64


def test_library_exports_every_declared_symbol():
    """include/tgnh.h <-> libtgnh.so: every declared function is exported (no compute calls without a GPU)."""
    hdr = open(os.path.join(ROOT, "include", "tgnh.h")).read()
    import re
    declared = set(re.findall(r"\b(tgnh_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"tgnh_params", "tgnh_handle", "tgnh_comm"}
    lib = ctypes.CDLL(capi.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    assert declared == set(capi.SYMBOLS)
    assert b"sm_100a" in capi.lib().tgnh_build_info()


def test_sass_is_blackwell_native():
    """The streaming kernels use TMA bulk copies (UBLKCP) and mbarriers (SYNCS) and are built for sm_100a only."""
    out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and "sm_90" not in out and "sm_80" not in out
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "UBLKCP" in sass and "SYNCS" in sass


def _params_error(**kw):
    s = synth.water_box(8, 2)
    for k, v in list(kw.items()):
        if hasattr(s, k):
            setattr(s, k, v); kw.pop(k)
    with pytest.raises(capi.TgnhError) as e:
        capi.Handle(s, **kw)
    return e.value


def test_create_validates_before_touching_the_device():
    """Argument errors are reported even without a GPU; a valid request without a GPU fails loudly (no CPU fallback)."""
    assert _params_error(num_nh_chains=0).code == capi.ERR_UNSUPPORTED
    assert _params_error(padded=30).code == capi.ERR_INVALID_ARGUMENT           # not a multiple of 4 / < N
    assert _params_error(max_drude_distance=-1.0).code == capi.ERR_INVALID_ARGUMENT
    assert _params_error(num_temp_groups=40).code == capi.ERR_UNSUPPORTED
    assert _params_error(force_format=7).code == capi.ERR_INVALID_ARGUMENT
    import torch
    if not torch.cuda.is_available():
        s = synth.water_box(8, 2)
        with pytest.raises(capi.TgnhError) as e:
            capi.Handle(s)
        assert e.value.code == capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_bench_reference_arm_runs_on_cpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    import json
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] in ("port", "reference")
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    # both arms print the same `config` object for the same command line (nothing measured, nothing about the implementation)
    import bench
    assert line["config"] == bench.workload_config("c4", 1) and line["config"]["total_particles"] == 10_000_000


def test_bench_config_names_every_workload():
    """`config` of the bench line: workload text, particle counts and the L2 statement, for every workload and GPU count, without a device."""
    import bench
    for w in bench.WORKLOADS:
        for world in (1, 2, 8):
            cfg = bench.workload_config(w, world)
            assert set(cfg) == {"workload", "particles_per_gpu", "total_particles", "l2"} and cfg["workload"] == bench.WORKLOADS[w]
            assert cfg["total_particles"] == (200_000_000 if w == "c5" else cfg["particles_per_gpu"] * world)
            assert ("larger than L2" in cfg["l2"]) == (w.startswith("c4") or w == "c5")
    assert bench.workload_config("c1", 1)["particles_per_gpu"] == 2500 and bench.workload_config("c2", 1)["particles_per_gpu"] == 50000


_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from openmm_drudenose_b200 import synth
from oracle import oracle as O
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
mol = 600
per = mol // world
shard = synth.water_box(per, 3, first_molecule=rank * per, box_molecules=mol, quantize_masses=True)
# what tgnh_create does when sharded: local DOF contributions summed over ranks
o_local = O.Oracle(shard, O.TG)
# the per-step collective: local 2*KE vectors summed (stands in for ncclAllReduce(double[G+2]))
ke = torch.from_numpy(o_local.compute_ke2(shard.velocities.copy()))
dist.all_reduce(ke)
if rank == 0:
    whole = synth.water_box(mol, 3, quantize_masses=True)
    ref = O.Oracle(whole, O.TG).compute_ke2(whole.velocities.copy())
    np.testing.assert_allclose(ke.numpy(), ref, rtol=1e-12)
    print("OK")
dist.destroy_process_group()
"""


def test_sharded_kinetic_energy_reduction_gloo(tmp_path):
    """world_size 2 on CPU (gloo): molecule-aligned shards + a sum all-reduce of the double[G+2] vector reproduce
    the single-process kinetic energies — the host-side logic of the sharded path."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "OK" in out.stdout


# ---- the host-side plan (tgnh_plan_tiles): table checks and tiling, no device needed -----------------------------------
def _assert_plan_invariants(s, ts, nbig):
    n = s.num_particles
    assert ts[0] == 0 and ts[-1] == n and np.all(np.diff(ts) > 0) and np.all(np.diff(ts) <= 512)
    first = np.full(s.num_residues, n, np.int64); last = np.full(s.num_residues, -1, np.int64)
    np.minimum.at(first, s.res_id, np.arange(n)); np.maximum.at(last, s.res_id, np.arange(n))
    size = last - first + 1
    tile_of = np.searchsorted(ts, np.arange(n), side="right") - 1
    small = size <= 128
    assert np.all(tile_of[first[small]] == tile_of[last[small]]), "a residue of <= 128 particles was split"
    assert nbig == int(np.sum(~small))
    if len(s.pair_drude):
        assert np.all(tile_of[s.pair_drude] == tile_of[s.pair_parent]), "a Drude pair was separated"
    for r in np.nonzero(~small)[0]:          # a big residue starts its own tile and ends one
        assert first[r] in ts and (last[r] + 1) in ts


def test_plan_tiles_on_the_synthetic_systems():
    for s in (synth.water_box(3000, 4), synth.nacl_box(), synth.ionic_liquid(200), synth.polymer_in_water(1500, (300, 700), 2),
              synth.build([synth.WATER4, synth.SOD, synth.SWM4], np.arange(1501) % 3, np.arange(1501) % 2, 2)):
        ts, nbig, uniform = capi.plan_tiles(s)
        _assert_plan_invariants(s, ts, nbig)
        assert uniform


def test_plan_tiles_random_topologies():
    """Random mixtures of residue sizes from 1 to 1500 particles with random Drude pairs up to 100 indices apart."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.sampled_from([1, 2, 3, 4, 5, 7, 35, 127, 128, 129, 300, 513, 1500]), min_size=1, max_size=40), st.integers(0, 2 ** 31 - 1))
    def run(sizes, seed):
        import types
        rng = np.random.default_rng(seed)
        res_id, pd, pp, base = [], [], [], 0
        for k, size in enumerate(sizes):
            used = set()
            for _ in range(size // 3):
                a = int(rng.integers(0, size)); b = a + int(rng.integers(1, 101))
                if b < size and a not in used and b not in used:
                    used.update((a, b)); pd.append(base + b); pp.append(base + a)
            res_id += [k] * size
            base += size
        res_id = np.array(res_id, np.int32)
        s = types.SimpleNamespace(
            num_particles=base, masses=rng.uniform(1.0, 20.0, base), pair_drude=np.array(pd, np.int32), pair_parent=np.array(pp, np.int32),
            temp_group=(res_id % 2).astype(np.int32), res_id=res_id, constraints=np.zeros((0, 2), np.int32), num_residues=len(sizes),
            num_temp_groups=2, num_nh_chains=3, drude_steps=20, use_drude_nh_chains=True, use_com_temp_group=True, temperature=300.0,
            coupling_time=0.1, drude_temperature=1.0, drude_coupling_time=0.005, step_size=0.001, max_drude_distance=0.02)
        try:
            ts, nbig, _ = capi.plan_tiles(s)
        except capi.TgnhError as e:
            assert e.code == capi.ERR_UNSUPPORTED and "without separating a Drude pair" in str(e)     # legal refusal, never a bad plan
            return
        _assert_plan_invariants(s, ts, nbig)
    run()


def test_plan_reports_the_reference_table_errors():
    """The reference's two exceptions (CudaDrudeTGNHKernels.cpp:146, 193) and this library's own checks, all before any device is touched."""
    def err(mut):
        s = synth.water_box(64, 2)
        mut(s)
        with pytest.raises(capi.TgnhError) as e:
            capi.plan_tiles(s)
        return e.value

    def drude_group(s): s.temp_group = s.temp_group.copy(); s.temp_group[s.pair_drude[0]] ^= 1
    e = err(drude_group)
    assert e.code == capi.ERR_TEMP_GROUP and "Temperature group for drude particle must be the same as the parent particle" in str(e)

    def constraint_group(s):
        s.temp_group = s.temp_group.copy(); s.temp_group[2] ^= 1
        s.constraints = np.array([[0, 2]], np.int32)
    assert err(constraint_group).code == capi.ERR_TEMP_GROUP

    def scattered(s): s.res_id = s.res_id.copy(); s.res_id[[1, 5]] = s.res_id[[5, 1]]
    assert "not a contiguous particle range" in str(err(scattered))

    def far_pair(s): s.res_id = np.zeros_like(s.res_id); s.num_residues = 1; s.pair_drude = s.pair_drude.copy(); s.pair_drude[0] = 250
    assert err(far_pair).code == capi.ERR_UNSUPPORTED

    def cross_residue(s): s.pair_parent = s.pair_parent.copy(); s.pair_parent[0] = 8        # molecule 2: same group, other residue
    assert "spans two residues" in str(err(cross_residue))

    def twice(s): s.pair_parent = s.pair_parent.copy(); s.pair_parent[1] = s.pair_parent[0]
    assert "more than one pair" in str(err(twice))

    def massless_parent(s): s.masses = s.masses.copy(); s.masses[s.pair_parent[0]] = 0.0
    assert "massless" in str(err(massless_parent))       # the reference's pair transform yields NaN there; refused instead

    def bad_group(s): s.temp_group = s.temp_group.copy(); s.temp_group[3] = 7
    assert err(bad_group).code == capi.ERR_INVALID_ARGUMENT


def test_residues_are_ignored_without_the_com_group():
    """Without the COM temperature group the reference never reads the residue table (drudeTGNH.cu:87-108): scattered residue ids
    and Drude pairs across residues are accepted and planned over residues of the library's own (the ranges spanned by
    overlapping pairs); with the COM group the same table is refused."""
    s = synth.water_box(700, 2, use_com_temp_group=False)
    ref_tiles, _, _ = capi.plan_tiles(s)
    s.res_id = np.random.default_rng(3).integers(0, s.num_residues, s.num_particles).astype(np.int32)
    ts, nbig, uniform = capi.plan_tiles(s)
    assert nbig == 0 and uniform and ts[0] == 0 and ts[-1] == s.num_particles and np.all(np.diff(ts) <= 512)
    tile_of = np.searchsorted(ts, np.arange(s.num_particles), side="right") - 1
    assert np.all(tile_of[s.pair_drude] == tile_of[s.pair_parent])
    cs, spec, table, nspecies, max_res = capi.plan_chunks(s)
    assert max_res == 2 and nspecies == 6                        # parent, Drude, ordinary particle x 2 temperature groups
    chunk_of = np.searchsorted(cs, np.arange(s.num_particles), side="right") - 1
    assert np.all(chunk_of[s.pair_drude] == chunk_of[s.pair_parent]) and np.all(np.diff(cs) <= 32)
    s.use_com_temp_group = True
    with pytest.raises(capi.TgnhError) as e:
        capi.plan_tiles(s)
    assert "not a contiguous particle range" in str(e.value)


def test_descriptors_say_which_particles_are_interchangeable():
    """What a CudaForceInfo built on tgnh_plan_descriptors would tell OpenMM's atom reordering: molecules of the same kind and
    temperature group have equal words particle by particle, molecules in different groups do not."""
    s = synth.water_box(8, 2)                     # molecule k -> group k % 2
    d = capi.plan_descriptors(s).reshape(8, 4)
    assert np.array_equal(d[0], d[2]) and np.array_equal(d[1], d[3])
    assert not np.array_equal(d[0], d[1])
    tg, role, partner = d & 0x7f, (d >> 8) & 3, d.astype(np.int32) >> 24
    assert np.array_equal(tg, np.repeat((np.arange(8) % 2)[:, None], 4, 1))
    assert np.array_equal(role[0], [2, 1, 0, 0]) and np.array_equal(partner[0], [1, -1, 0, 0])       # parent, Drude, H, H
    assert np.array_equal((d[0] >> 10) & 0x7f, [0, 1, 2, 3]) and np.array_equal((d[0] >> 17) & 0x7f, [3, 2, 1, 0])


# ---- the warp-chunk plan (csrc/tgnh_v2.cuh): chunks, species bytes, species table ----------------------------------------------------------
def _assert_chunk_invariants(s, cs, spec, table, nspec, maxres):
    n = s.num_particles
    assert cs[0] == 0 and cs[-1] == n and (len(cs) - 1) % capi.lib().tgnh_chunks_per_tile() == 0 and np.all(np.diff(cs) >= 0) and np.diff(cs).max() <= 32
    res = np.asarray(s.res_id)
    starts = np.unique(cs[cs < n])
    assert np.all(res[starts[1:]] != res[starts[1:] - 1])                       # every chunk starts a residue ...
    sizes = np.bincount(res)
    assert maxres == sizes.max()
    # ... and a pair's partner lies in its own chunk
    chunk_of = np.searchsorted(cs, np.arange(n), side="right") - 1
    chunk_of = np.searchsorted(starts, np.arange(n), side="right") - 1
    assert np.array_equal(chunk_of[s.pair_drude], chunk_of[s.pair_parent])
    # species rows: masses as hi + lo floats, reduced mass, residue inverse mass, mass fraction of the partner, meta bits
    assert spec.max() < nspec <= 255
    m = np.asarray(s.masses, np.float64)
    row = table[spec].astype(np.float64)
    meta = table[spec, 3].view(np.uint32)
    np.testing.assert_allclose(row[:, 0] + row[:, 1], m, rtol=1e-14, atol=0)
    resmass = np.bincount(res, weights=m)
    with np.errstate(divide="ignore"):
        invm = np.where(resmass[res] > 0, 1.0 / resmass[res], 0.0)
    np.testing.assert_allclose(row[:, 2] + row[:, 6], invm, rtol=1e-13, atol=0)
    partner = np.zeros(n, np.int64); role = np.zeros(n, np.int64)
    partner[s.pair_drude] = s.pair_parent - s.pair_drude; partner[s.pair_parent] = s.pair_drude - s.pair_parent
    role[s.pair_drude] = 1; role[s.pair_parent] = 2
    assert np.array_equal(meta & 31, s.temp_group) and np.array_equal((meta >> 5) & 3, role)
    p6 = ((meta.astype(np.int64) >> 7) & 63)
    assert np.array_equal(np.where(p6 >= 32, p6 - 64, p6), partner)
    first_idx = np.r_[0, np.nonzero(res[1:] != res[:-1])[0] + 1]
    off_first = np.arange(n) - np.repeat(first_idx, sizes[res[first_idx]])
    assert np.array_equal((meta >> 13) & 31, off_first) and np.array_equal((meta >> 18) & 31, sizes[res] - 1 - off_first)
    mj = np.where(partner != 0, m[np.arange(n) + partner], 0.0)
    with np.errstate(invalid="ignore", divide="ignore"):
        mu = np.where(partner != 0, m * mj / (m + mj), 0.0)
        fj = np.where(partner != 0, mj / (m + mj), 0.0)
    np.testing.assert_allclose(row[:, 4] + row[:, 5], mu, rtol=1e-13, atol=0)
    np.testing.assert_allclose(row[:, 7], fj, rtol=1e-7, atol=0)
    assert np.all(table[255] == 0)                                                              # the "no particle" row


def test_plan_chunks_on_the_synthetic_systems():
    for s in (synth.water_box(3000, 4), synth.nacl_box(), synth.swm4_box(777),
              synth.build([synth.WATER4, synth.SOD, synth.SWM4], np.arange(1501) % 3, np.arange(1501) % 2, 2)):
        _assert_chunk_invariants(s, *capi.plan_chunks(s))
    for s, why in ((synth.ionic_liquid(20), "more than 32 particles"), (synth.polymer_in_water(100, (300,), 2), "more than 32 particles")):
        with pytest.raises(capi.TgnhError) as e:
            capi.plan_chunks(s)
        assert e.value.code == capi.ERR_UNSUPPORTED and why in str(e.value)
    with pytest.raises(capi.TgnhError) as e:
        capi.plan_chunks(synth.water_box(100, 2), precision=capi.PRECISION_MIXED)
    assert "mixed / double" in str(e.value)


def test_plan_chunks_random_topologies():
    """Random mixtures of residues of 1..32 particles with random masses, random in-residue Drude pairs and random groups: either a
    valid plan or the documented refusal (more than 255 species)."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.lists(st.integers(1, 32), min_size=1, max_size=200), st.integers(0, 2 ** 31 - 1), st.integers(1, 6))
    def run(sizes, seed, kinds):
        import types
        rng = np.random.default_rng(seed)
        palette = rng.uniform(1.0, 20.0, kinds)
        res_id, pd, pp, base = [], [], [], 0
        for k, size in enumerate(sizes):
            used = set()
            for _ in range(size // 3):
                a = int(rng.integers(0, size)); b = int(rng.integers(0, size))
                if a != b and a not in used and b not in used:
                    used.update((a, b)); pd.append(base + b); pp.append(base + a)
            res_id += [k] * size
            base += size
        res_id = np.array(res_id, np.int32)
        masses = palette[rng.integers(0, kinds, base)]
        order = np.argsort(pd) if pd else []
        s = types.SimpleNamespace(
            num_particles=base, masses=masses, pair_drude=np.array(pd, np.int32)[order], pair_parent=np.array(pp, np.int32)[order],
            temp_group=(res_id % 3).astype(np.int32), res_id=res_id, constraints=np.zeros((0, 2), np.int32), num_residues=len(sizes),
            num_temp_groups=3, num_nh_chains=3, drude_steps=20, use_drude_nh_chains=True, use_com_temp_group=True, temperature=300.0,
            coupling_time=0.1, drude_temperature=1.0, drude_coupling_time=0.005, step_size=0.001, max_drude_distance=0.02)
        try:
            plan = capi.plan_chunks(s)
        except capi.TgnhError as e:
            assert e.code == capi.ERR_UNSUPPORTED and "more than 255 particle species" in str(e)
            return
        _assert_chunk_invariants(s, *plan)
    run()
