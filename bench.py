#!/usr/bin/env python
"""bench.py — TGNH step throughput on B200 (BASELINE.json: particle-steps/s, HBM roofline fraction).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c5|c1|c2|c3]

N = 1   : config C4 — synthetic 10M-particle Drude system, 4 temperature groups, integrator-only with
          fixed synthetic fp32 forces.  One "step" = one full TGNH step over all particles.
N > 1   : the same C4 shard (10M particles) on every GPU, molecule-aligned particle ranges of ONE system of
          N x 10M particles (weak scaling); the only collective is the NCCL all-reduce of the double[G+2]
          kinetic-energy vector before each chain update.  `--workload c5` runs the 200M-particle system
          split over the N ranks instead (strong scaling).
--impl reference : the reference algorithm's CPU implementation (oracle/, a restatement of the plugin's
          platforms; the real OpenMM Reference platform cannot be built here) on the host cores.

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for what every key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from openmm_drudenose_b200 import synth  # noqa: E402

ALG_BYTES_STEP = 120      # 2*(16 + 12 + 16) + (16 + 16) with fp32 SoA forces
ALG_BYTES_HALF1 = 76      # first-half kernel: reads velm, force, posq; writes velm, posq
ALG_BYTES_HALF2 = 44      # second-half kernel: reads velm, force; writes velm
C4_MOLECULES = 2_500_000  # x 4 particles = 10M
C5_MOLECULES = 50_000_000 # x 4 particles = 200M

MOLECULES_OVERRIDE = 0
# generator options of the C4 family (synth.build): "c4" is a state the thermostats can hold for the length of a bench run
C4_STATE = {"c4": dict(pair_force="common", cold_drudes=True, force_sigma=2.0), "c4-hot": dict(pair_force="common"),
            "c4-wall": dict(pair_force="frozen_spring")}

WORKLOADS = {
    "c4": "C4 synthetic 10M-particle Drude system (2.5M 4-particle molecules, 2.5M Drude pairs), G=4, M=3, S=20, "
          "COM group on, hard wall 0.02 nm, fixed fp32 SoA forces (sigma 2 kJ/mol/nm), equilibrated dual-thermostat start "
          "(300 K / Drude 1 K): thermostats hold their targets, ~0.5 % of the pairs meet the wall per step",
    "c4-hot": "round 1's C4 state: independent 300 K velocities on every particle and fixed forces of sigma 200 kJ/mol/nm (the Drude "
              "thermostat quenches all relative motion, group temperatures run to 2e4 K, the chain takes its full-range exp path)",
    "c4-wall": "C4 with the frozen-spring pair forces of SURVEY.md 8d: every Drude pair hits the hard wall on every step "
               "(hard-wall stress case)",
    "c5": "C5 synthetic 200M-particle Drude system sharded over the ranks, G=4",
    "c1": "C1 NaCl 1M box shape, N=2500, G=2, equilibrated start", "c2": "C2 SWM4-NDP 10k waters, N=50000, G=1, equilibrated start",
    "c1-hot": "C1 from round 1's unequilibrated state", "c2-hot": "C2 from round 1's unequilibrated state", "c3-hot": "C3 from round 1's unequilibrated state",
    "c3": "C3 [BMIM][BF4]-like 1000 ion pairs, N=45000, G=3, equilibrated start",
}


def workload_config(workload, world):
    """The `config` object of the JSON line: what the workload IS, nothing measured and nothing about the implementation, so that
    both arms (`--impl ours` / `--impl reference`) print the same object for the same command line.  Everything measured beside the
    headline (C5 strong split, replicas, shard check, small systems, call patterns ...) goes to `details`."""
    if workload in ("c4", "c4-wall", "c4-hot"):
        per = (MOLECULES_OVERRIDE or C4_MOLECULES) * 4
    elif workload == "c5":
        per = (C5_MOLECULES // world) * 4
    else:
        per = make_system(workload, 0, 1).num_particles
    return {"workload": WORKLOADS[workload], "particles_per_gpu": per, "total_particles": per * world,
            "l2": "inputs larger than L2 (44 B of state and forces per particle against 126 MB of L2)" if per * 44 > 2 * 126e6
                  else "inputs smaller than L2 (latency workload: the step is bound by launches and the serial chain, not by bytes)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,clocks.mem")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 6 and r[2 + i].lower() == "active" for r in self.rows)]
        mem = [float(r[6]) for r in self.rows if len(r) >= 7 and r[6].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "sm_min_mhz": min(sm) if sm else None, "mem_mhz": float(np.median(mem)) if mem else None,
                "reasons": reasons, "samples": len(sm)}


def make_system(workload, rank, world):
    if workload in ("c4", "c4-wall", "c4-hot"):
        mol = MOLECULES_OVERRIDE or C4_MOLECULES
        return synth.water_box(mol, 4, first_molecule=rank * mol, box_molecules=world * mol, **C4_STATE[workload])
    if workload == "c5":
        per = C5_MOLECULES // world
        return synth.water_box(per, 4, first_molecule=rank * per, box_molecules=C5_MOLECULES, **C4_STATE["c4"])
    # C1-C3: like C4 from an equilibrated dual-thermostat state (thermostats hold their targets, the chain runs its short-polynomial
    # path); "-hot" = round 1's state (independent 300 K velocities on the Drude particles: the Drude thermostat is far from its
    # 1 K target for the whole run and the chain takes its full-range exp path every step)
    small = {} if workload.endswith("-hot") else C4_STATE["c4"]
    if workload in ("c1", "c1-hot"):
        return synth.nacl_box(**small)
    if workload in ("c2", "c2-hot"):
        return synth.swm4_box(10000, **small)
    if workload in ("c3", "c3-hot"):
        return synth.ionic_liquid(1000, **small)
    raise SystemExit(f"unknown workload {workload}")


# ---------------------------------------------------------------------------------------------------
# CPU legs (the only place besides tests/ and smoke() that touches oracle/)
# ---------------------------------------------------------------------------------------------------
def cpu_run(system, steps, warmup, threads):
    """Times the reference algorithm on the host.  Preferred: oracle/_ref — the reference's OWN reference-platform
    sources compiled unmodified against the OpenMM API shim (serial like the original; it implements the dual
    Nose-Hoover part only and ignores temperature groups, SURVEY.md finding 1).  Otherwise the oracle port
    (temperature groups + COM group, OpenMP over `threads`).  Returns (seconds, kind, cores)."""
    from oracle import ref as R
    if R.available() and not os.environ.get("TGNH_BENCH_FORCE_PORT"):
        sim = R.ReferenceSim(system)
        p, v, f = system.positions.copy(), system.velocities.copy(), system.forces.copy()
        if warmup:
            sim.step(p, v, f, warmup)
        t0 = time.perf_counter()
        sim.step(p, v, f, steps)
        return time.perf_counter() - t0, "reference", 1
    from oracle import oracle as O
    O.lib().tgnh_oracle_set_threads(threads)
    o = O.Oracle(system, O.TG, constraints=system.constraints)
    p, v, f = system.positions.copy(), system.velocities.copy(), system.forces.copy()
    if warmup:
        o.step(p, v, f, warmup)
    t0 = time.perf_counter()
    o.step(p, v, f, steps)
    dt = time.perf_counter() - t0
    O.lib().tgnh_oracle_set_threads(1)
    return dt, "port", threads


def cpu_port_all_cores(system, steps):
    """the OpenMP-over-molecules variant of the oracle port (temperature groups + COM group, fp64) on all
    host cores, beside the serial reference platform.  Returns a dict for the JSON line."""
    from oracle import oracle as O
    threads = os.cpu_count() or 1
    O.lib().tgnh_oracle_set_threads(threads)
    o = O.Oracle(system, O.TG, constraints=system.constraints)
    p, v, f = system.positions.copy(), system.velocities.copy(), system.forces.copy()
    o.step(p, v, f, 2)
    t0 = time.perf_counter()
    o.step(p, v, f, steps)
    dt = time.perf_counter() - t0
    O.lib().tgnh_oracle_set_threads(1)
    return {"value": system.num_particles * steps / dt, "unit": "particle-steps/s", "cores": threads, "kind": "port",
            "sample": f"{system.num_particles} particles x {steps} steps ({dt:.2f} s), oracle-tg port, OpenMP over molecules"}


def cpu_sample_system(workload):
    """Bounded sample of the workload for the CPU legs: 1M particles of the same generator (C4/C5), else the config itself."""
    if workload in ("c4", "c4-wall", "c4-hot", "c5"):
        return synth.water_box(250_000, 4, box_molecules=C4_MOLECULES, **C4_STATE.get(workload, C4_STATE["c4"]))
    return make_system(workload, 0, 1)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on the host cores, bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    system = cpu_sample_system(args.workload)
    dt, kind, cores = cpu_run(system, args.steps, args.warmup, os.cpu_count() or 1)
    value = system.num_particles * args.steps / dt
    what = ("/root/reference platforms/reference + openmmapi sources built against the OpenMM API shim (oracle/_ref), serial"
            if kind == "reference" else "oracle port (fp64 restatement, OpenMP)")
    sample = f"{system.num_particles} particles of the same generator x {args.steps} steps ({dt:.2f} s); {what}"
    out = {
        "impl": "reference", "metric": "TGNH step particle-steps/s", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, max(args.gpus, 1)),
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "cpu_port_all_cores": cpu_port_all_cores(system, max(args.steps, 10)),
    }
    print(json.dumps(out))


# ---------------------------------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from openmm_drudenose_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the TGNH path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ids = [capi.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = capi.Comm(ids[0], world, rank, local)

    system = make_system(args.workload, rank, world)
    n = system.num_particles
    padded, (h_velm, h_posq, h_force), (velm, posq, force) = device_buffers(torch, system, dev, pinned=True)
    h = capi.Handle(system, padded=padded, device=local, comm=comm)
    tstream = torch.cuda.Stream(device=dev)            # the step runs on its own (non-default) stream
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    ptrs = [velm.data_ptr(), posq.data_ptr(), force.data_ptr()]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput: W warm-up steps, then exactly K timed steps ----
    # The GPU has idled for seconds while the host generated the system; 3 warm-up steps are < 1 ms and do not bring
    # clocks and power state back (one fresh box measured 0.275 ms/step instead of 0.239 that way).  Spin-up: 0.5 s of
    # device-to-device copies that do not touch the system's state, with the clock sampler already running; no idle gap
    # between it, the warm-up steps and the timed region.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    scratch = torch.empty_like(velm)
    t_end = time.perf_counter() + 0.5
    while time.perf_counter() < t_end:
        for _ in range(50):
            scratch.copy_(velm)
        torch.cuda.synchronize()
    del scratch
    h.step(*ptrs, max(args.warmup, 3), stream)
    launches0 = h.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h.step(*ptrs, args.steps, stream)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = h.launch_count - launches0
    # per-kernel durations for the roofline object: the same K steps once more, every streaming launch bracketed by
    # CUDA events on the launching stream (kept out of the pass above: an event between two launches serialises
    # them and removes the programmatic-dependent-launch overlap the product path runs with)
    h.set_profiling(True)
    h.step(*ptrs, args.steps, stream)
    barrier()
    prof = h.profile()
    h.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    total_particles = n * world
    value = total_particles * args.steps / (ms * 1e-3)
    generation = h.kernel_generation
    lazy = h.lazy_second_kick
    rpl = h.residue_per_lane

    # ---- end to end through the C-ABI with HOST buffers: every step copies velm/posq/force in from pinned
    #      memory, runs one step, and copies velm/posq + the 2*KE vector back (tgnh_step_host) ----
    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 10))
    ke2 = h.kinetic_energies()

    def e2e_leg(forces_unchanged):
        if not e2e_steps:
            return 0.0
        k = h.step_host2(h_velm.data_ptr(), h_posq.data_ptr(), h_force.data_ptr(), 1, forces_unchanged=forces_unchanged)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            k = h.step_host2(h_velm.data_ptr(), h_posq.data_ptr(), h_force.data_ptr(), 1, forces_unchanged=forces_unchanged)
        barrier()
        nonlocal ke2
        ke2 = k
        return total_particles * e2e_steps / max_over_ranks(max(time.perf_counter() - t0, 1e-9))

    e2e_full = e2e_leg(False)          # velm + posq + forces up, velm + posq + energies down, every step
    e2e_value = e2e_leg(True)          # the bench's forces are fixed: uploaded once, the caller vouches they are unchanged
    h2d = n * 16 * 2
    d2h = n * 16 * 2 + 8 * h.T

    # ---- further legs, all after the headline's timed region (none of them changes `value`) ----
    extra = {}
    if args.workload == "c4" and not args.quick:
        extra["openmm_call_pattern_ms"] = call_pattern_leg(torch, capi, h, ptrs, stream, barrier)
        if world > 1:
            # independent replicas, one full system per GPU, no communication (the reference's only multi-GPU mode,
            # CudaDrudeTGNHKernelFactory.cpp:62 takes contexts[0])
            hr = capi.Handle(system, padded=padded, device=local)
            rms = max_over_ranks(timed_steps(torch, hr, ptrs, args.steps, stream, barrier))
            hr.close()
            extra["replicas"] = {"replicas": world, "particles_each": n, "ms_per_step": rms / args.steps, "value": total_particles * args.steps / (rms * 1e-3),
                                 "what": "independent replicas, one handle per GPU, no exchange"}
    h_T = h.T
    exchange_kind = h.exchange_kind
    if world > 1 and exchange_kind == 2 and not args.quick:
        # where does the time between the second half and the next first half go?  50 single steps, device timer stamps of the
        # exchange on every rank: the rank that publishes last waits only for the NVLink latency, the others also for the skew
        waits, gaps = [], []
        for _ in range(50):
            h.step(*ptrs, 1, stream)
            w, g = h.exchange_timing(stream)
            waits.append(w); gaps.append(g)
        allw = [None] * world
        dist.all_gather_object(allw, (waits, gaps))
        if rank == 0:
            wm = np.array([a[0] for a in allw]); gm = np.array([a[1] for a in allw])        # [rank][step]
            extra["exchange_timing_us"] = {
                "steps": 50, "what": "wait of the chain launch for all ranks' energy partials (peer inboxes), per step",
                "latency_min_over_ranks_median": float(np.median(wm.min(axis=0))), "wait_max_over_ranks_median": float(np.median(wm.max(axis=0))),
                "skew_median": float(np.median(wm.max(axis=0) - wm.min(axis=0))), "per_rank_median": [float(x) for x in np.median(wm, axis=1)],
                "publish_to_wait_start_median": float(np.median(gm))}
    h.close()
    if args.workload == "c4" and not args.quick:
        del velm, posq, force
        torch.cuda.empty_cache()
        if world > 1:
            extra["shard_check"] = shard_check_leg(torch, capi, dev, local, comm, rank, world, stream, dist)
        extra["c5_strong"] = c5_leg(torch, capi, dev, local, comm, rank, world, stream, barrier, dist, min(args.steps, 20))
        torch.cuda.empty_cache()
        if world == 1:
            extra["small_systems"] = small_systems_leg(torch, capi, dev, local, stream, barrier)

    if rank == 0:
        peak, peak_src = peaks()
        cfg = workload_config(args.workload, world)
        assert cfg["particles_per_gpu"] == n, (cfg, n)
        a_ms, a_cnt = prof["half1"]
        b_ms, b_cnt = prof["half2"]
        ach = ALG_BYTES_HALF1 * n / (a_ms / max(a_cnt, 1) * 1e-3) / 1e9 if a_cnt else None
        # the lazy second kick (tgnh.h: tgnh_lazy_second_kick): all second halves of a tgnh_step call but the last store nothing
        half2_bytes = (ALG_BYTES_HALF2 - 16 * (b_cnt - 1) / b_cnt) if (lazy and b_cnt) else ALG_BYTES_HALF2
        moved = ALG_BYTES_STEP - (16 * (args.steps - 1) / args.steps if lazy else 0)
        traffic, traffic_src = ncu_traffic(args.workload, n)
        out = {
            "metric": "TGNH step particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": cfg,
            "details": {"parallelism": f"particle-range shards x{world}" if world > 1 else "single GPU",
                        "host_numa_binding": None if world == 1 else f"rank 0 on node {numa_node}",
                        "exchange": {0: "none", 1: "ncclAllReduce of double[T] per step", 2: "peer-mapped inboxes over NVLink (no collective launch)"}[exchange_kind],
                        "kernels": "warp-chunk kernels (tgnh_v2.cuh)" if generation == 2 else "first-generation kernels (tgnh_kernels.cuh)",
                        "spin_up": "0.5 s of device-to-device copies before the warm-up steps (clock ramp after the host-side set-up)",
                        "accumulation": "fp32 state (OpenMM single-precision layouts), fp64 KE reductions and NH chain",
                        "step_achieved_gbs": ALG_BYTES_STEP * n * args.steps / (ms * 1e-3) / 1e9,
                        "step_frac_of_peak": ALG_BYTES_STEP * n * args.steps / (ms * 1e-3) / 1e9 / peak,
                        "lazy_second_kick": bool(lazy),
                        "residue_per_lane": int(rpl),
                        "step_bytes_moved_per_particle": moved,
                        "step_frac_of_peak_on_bytes_moved": moved * n * args.steps / (ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_full, "unit": "particle-steps/s", "h2d_bytes_per_step": n * 16 * 2 + 3 * padded * 4, "d2h_bytes_per_step": d2h,
                    "what": "tgnh_step_host2 per step: pinned host velm/posq/forces -> device in 8 pipelined particle ranges, 1 step, velm/posq/2KE "
                            "back; H2D and D2H overlap on the two copy engines",
                    "forces_uploaded_once": {"value": e2e_value, "h2d_bytes_per_step": h2d,
                                             "what": "the same with TGNH_HOST_FORCES_UNCHANGED: the bench's forces are fixed, the caller vouches for it"}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": ("tgnh_v2_kernel<V2_A>" if generation == 2 else "tgnh_stream_kernel<KIND_A>") + " (scale+kick+drift+hard wall)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_particle": ALG_BYTES_HALF1,
                         "avg_launch_ms": a_ms / max(a_cnt, 1),
                         "half2_avg_launch_ms": b_ms / max(b_cnt, 1),
                         "half2_algorithmic_bytes_per_particle": half2_bytes,
                         "half2_achieved": (half2_bytes * n / (b_ms / max(b_cnt, 1) * 1e-3) / 1e9) if b_cnt else None,
                         "note": "achieved = algorithmic bytes / launch time; above the measured copy peak when part of velm is served by the L2 "
                                 "(every launch starts on the tiles the previous one touched last); traffic = DRAM bytes per launch from ncu"},
            "ke2_last": [float(x) for x in ke2],
        }
        out["details"].update(extra)
        if world == 1 and args.workload == "c4" and not args.quick and not args.no_reference_cuda:
            # the reference's own CUDA kernels on the same system, same box (an extra measured baseline; never on the product path)
            try:
                out["reference_cuda"] = reference_cuda_leg(system, 10)
            except Exception as e:      # noqa: BLE001 - a baseline leg must not take the bench line down
                out["reference_cuda"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        # CPU baseline beside it (rank 0, N == 1 only): bounded sample of the same generator, 1 core (the
        # reference platform is serial)
        if world == 1 and not args.no_cpu_baseline:
            sysb = cpu_sample_system(args.workload)
            bsteps = 20
            dt, kind, cores = cpu_run(sysb, bsteps, 2, 1)
            out["cpu_baseline"] = {"value": sysb.num_particles * bsteps / dt, "unit": "particle-steps/s", "cores": cores, "kind": kind,
                                   "sample": f"{sysb.num_particles} particles of the same generator x {bsteps} steps ({dt:.1f} s), "
                                             + ("the reference platform's own sources (oracle/_ref), serial fp64" if kind == "reference"
                                                else "oracle-tg port, serial fp64"),
                                   "port_all_cores": cpu_port_all_cores(sysb, 20)}
        print(json.dumps(out))
    if comm is not None:
        comm.close()
        dist.destroy_process_group()


def bind_to_gpu_numa_node(torch, local):
    """Run this rank (and first-touch its pinned host buffers) on the NUMA node its GPU hangs off: with 8 ranks the host-buffer leg
    otherwise funnels every rank's PCIe traffic through whichever sockets the scheduler picked.  Returns the node or None."""
    try:
        props = torch.cuda.get_device_properties(local)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except (OSError, ValueError, AttributeError):
        pass
    return None


def device_buffers(torch, system, dev, pinned=False):
    n = system.num_particles
    padded = ((n + 31) // 32) * 32
    mk = (lambda *shape: torch.zeros(shape, dtype=torch.float32).pin_memory()) if pinned else (lambda *shape: torch.zeros(shape, dtype=torch.float32))
    h_velm, h_posq, h_force = mk(padded, 4), mk(padded, 4), mk(3, padded)
    h_velm[:n] = torch.from_numpy(system.velm_f32())
    h_posq[:n] = torch.from_numpy(system.posq_f32())
    h_force[:, :n] = torch.from_numpy(np.ascontiguousarray(system.forces.T, np.float32))
    return padded, (h_velm, h_posq, h_force), (h_velm.to(dev), h_posq.to(dev), h_force.to(dev))


def timed_steps(torch, h, ptrs, steps, stream, barrier, warmup=3):
    """W warm-up steps, then `steps` steps in one tgnh_step call between CUDA events on the launching stream."""
    h.step(*ptrs, warmup, stream)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h.step(*ptrs, steps, stream)
    e1.record()
    barrier()
    return e0.elapsed_time(e1)


def small_systems_leg(torch, capi, dev, local, stream, barrier):
    """C1-C3 (BASELINE.json configs[0..2]): launch- and chain-latency bound; microseconds and launches per step."""
    out = {}
    for w in ("c1", "c2", "c3", "c1-hot", "c2-hot", "c3-hot"):
        system = make_system(w, 0, 1)
        padded, _, bufs = device_buffers(torch, system, dev)
        h = capi.Handle(system, padded=padded, device=local)
        ptrs = [b.data_ptr() for b in bufs]
        l0 = None
        h.step(*ptrs, 20, stream)
        l0 = h.launch_count
        ms = timed_steps(torch, h, ptrs, 500, stream, barrier, warmup=20)
        launches = (h.launch_count - l0 - 0) / 520.0
        out[w] = {"particles": system.num_particles, "us_per_step": 1e3 * ms / 500, "launches_per_step": round(launches, 2),
                  "kernel_generation": h.kernel_generation}
        h.close()
    return out


def call_pattern_leg(torch, capi, h, ptrs, stream, barrier, steps=20):
    """The OpenMM-facing call sequence tgnh_half1 / tgnh_half2 per step on the C4 buffers (what the KernelImpl issues), in its three
    modes: reference semantics (energies reduced from velm at the start of every step, scaling applied at its end), energies carried
    over, and carried over + scaling deferred into the next step.  ms per step."""
    velm, posq, force = ptrs
    out = {}
    for name, invalidate, flags in (("reference_semantics", True, capi.HALF2_DEFAULT), ("carry_over", False, capi.HALF2_DEFAULT),
                                    ("carry_over_deferred", False, capi.HALF2_DEFER_SCALE)):
        def run(k):
            for _ in range(k):
                if invalidate:
                    h.invalidate()
                h.half1(velm, posq, force, stream)
                h.half2(velm, force, flags, stream)
            h.flush(velm, stream)
        run(3)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        run(steps)
        e1.record()
        barrier()
        out[name] = e0.elapsed_time(e1) / steps
    return out


def c5_leg(torch, capi, dev, local, comm, rank, world, stream, barrier, dist, steps):
    """BASELINE.json configs[4]: ONE 200M-particle system split over the ranks (strong scaling).  Every rank tiles its 200M / N
    particles from a 10M-particle block of the C4 generator (the host generator makes 1M particles per second; the throughput of the
    integrator does not depend on the values) and builds the full index tables for its range."""
    per_mol = C5_MOLECULES // world
    reps = max(1, -(-per_mol // C4_MOLECULES))
    while per_mol % reps:
        reps += 1
    block_mol = per_mol // reps
    block = synth.water_box(block_mol, 4, first_molecule=rank * per_mol, box_molecules=C5_MOLECULES, **C4_STATE["c4"])
    nb = block.num_particles
    n = nb * reps
    system = synth.tile(block, reps)
    padded = ((n + 31) // 32) * 32
    velm = torch.zeros((padded, 4), dtype=torch.float32, device=dev)
    posq = torch.zeros((padded, 4), dtype=torch.float32, device=dev)
    force = torch.zeros((3, padded), dtype=torch.float32, device=dev)
    bv, bx = torch.from_numpy(block.velm_f32()).to(dev), torch.from_numpy(block.posq_f32()).to(dev)
    bf = torch.from_numpy(np.ascontiguousarray(block.forces.T, np.float32)).to(dev)
    for r in range(reps):
        velm[r * nb:(r + 1) * nb] = bv
        posq[r * nb:(r + 1) * nb] = bx
        force[:, r * nb:(r + 1) * nb] = bf
    del bv, bx, bf
    h = capi.Handle(system, padded=padded, device=local, comm=comm)
    del system
    ptrs = [velm.data_ptr(), posq.data_ptr(), force.data_ptr()]
    ms = timed_steps(torch, h, ptrs, steps, stream, barrier)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    h.close()
    total = n * world
    peak, _ = peaks()
    return {"total_particles": total, "particles_per_gpu": n, "ms_per_step": ms / steps, "value": total * steps / (ms * 1e-3),
            "frac_of_peak_per_gpu": ALG_BYTES_STEP * n * steps / (ms * 1e-3) / 1e9 / peak, "steps": steps, "scaling": "strong"}


def shard_check_leg(torch, capi, dev, local, comm, rank, world, stream, dist):
    """Untimed: a 160k-particle system sharded over the ranks for 25 steps; every rank must hold the same thermostat state, and it must
    equal the state of the same system stepped on ONE GPU (rank 0) up to the summation order of the energy partials."""
    mol, g, steps = 40000, 4, 25
    per = mol // world
    shard = synth.water_box(per, g, first_molecule=rank * per, box_molecules=per * world, quantize_masses=True)
    padded, _, bufs = device_buffers(torch, shard, dev)
    h = capi.Handle(shard, padded=padded, device=local, comm=comm)
    h.step(*[b.data_ptr() for b in bufs], steps, stream)
    torch.cuda.synchronize()
    state = (h.kinetic_energies().tolist(), h.vscale().tolist(), h.chain_state()[1].tolist())
    vel = bufs[0][:shard.num_particles, :3].double().cpu().numpy()
    h.close()
    gathered = [None] * world
    dist.all_gather_object(gathered, state)
    res = None
    if rank == 0:
        same = all(other == gathered[0] for other in gathered[1:])
        whole = synth.water_box(per * world, g, quantize_masses=True)
        padded1, _, bufs1 = device_buffers(torch, whole, dev)
        h1 = capi.Handle(whole, padded=padded1, device=local)
        h1.step(*[b.data_ptr() for b in bufs1], steps, stream)
        torch.cuda.synchronize()
        ke1, vs1, ed1 = h1.kinetic_energies(), h1.vscale(), h1.chain_state()[1]
        rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - b) / np.maximum(np.abs(b), 1e-300 + 1e-12 * np.abs(b).max())))
        vel1 = bufs1[0][:shard.num_particles, :3].double().cpu().numpy()
        dv = float(np.max(np.abs(vel - vel1)))
        # in units of the fp32 spacing at the particle's largest velocity component (the kernels' operations mix the components
        # of a particle and of its molecule, so a small component carries the absolute error of the large ones)
        big = np.maximum(np.abs(vel1).max(axis=1, keepdims=True), 1e-3)
        ulps = np.abs(vel - vel1) / np.spacing(big.astype(np.float32)).astype(np.float64)
        dv_ulp = float(np.max(ulps))
        res = {"ranks_identical": bool(same), "ke_vs_single_gpu": rel(state[0], ke1), "vscale_vs_single_gpu": rel(state[1], vs1),
               "eta_dot_vs_single_gpu": rel(state[2], ed1), "max_abs_dv_rank0": dv, "max_dv_in_fp32_ulps": dv_ulp,
               "components_differing": float(np.mean(ulps > 0)), "components_differing_by_more_than_1_ulp": float(np.mean(ulps > 1)), "particles": per * world * 4, "steps": steps}
        # (summation order differs -> scale factors differ in their last bits -> a few fp32 velocities round the other way, and a
        # velocity that did carries its ulp along while later steps may add another: 1e-10 on the energies; on velocities the MAXIMUM
        # over 480k components after 25 steps was between 0 and 3 ulps in the runs at 2, 4 and 8 GPUs kept under profiles/.  Bound: 8 ulps = 4.8e-7 relative, inside the
        # 1e-6 bar of the thermostat variables and 20x inside the per-step 1e-5 tolerance of x, v.)
        res["ok"] = bool(same and res["ke_vs_single_gpu"] < 1e-9 and res["vscale_vs_single_gpu"] < 1e-10 and res["eta_dot_vs_single_gpu"] < 1e-8 and dv_ulp <= 8.0)
        h1.close()
    return res


def reference_cuda_leg(system, steps):
    """The same-box GPU baseline: the reference's OWN CUDA platform (oracle/_refcuda: platforms/cuda sources and kernel strings compiled
    unmodified, NVRTC, OpenMM's launch rule; mixed precision, int64 forces, ~16 launches + 2 blocking downloads + 2 uploads per step) on the
    same C4 system, through the reference's DrudeTGNHIntegrator.step.  Test infrastructure used as a measured baseline, like cpu_baseline."""
    from oracle import refcuda as RC
    if not RC.available("reference"):
        return {"unavailable": "oracle/_refcuda/librefcuda.so was not built (no /root/reference at build time)"}
    t0 = time.perf_counter()
    sim = RC.CudaSim(system, "reference", "mixed")
    sim.set_state(system.positions, system.velocities, np.rint(system.forces * 4294967296.0) / 4294967296.0)
    setup = time.perf_counter() - t0
    sim.step(3)
    c0 = sim.counters()["launches"]
    ms = sim.time_steps(steps)
    launches = (sim.counters()["launches"] - c0) / steps
    sim.close()
    n = system.num_particles
    return {"value": n * steps / (ms * 1e-3), "unit": "particle-steps/s", "ms_per_step": ms / steps, "launches_per_step": launches, "steps": steps,
            "precision": "mixed (the reference's single mode reads its double scale factors as floats)", "setup_s": round(setup, 1),
            "what": "reference platforms/cuda sources + kernel strings, unmodified, behind the CUDA-platform stand-in (shim/cuda)"}


def ncu_traffic(workload, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE first-half launch, from the committed `ncu --set full` capture
    of this workload (profiles/ncu_full_rNN.json, written by scripts/summarise_profiles.py).  Never measured in this
    process: a run under a profiler is not a bench run.  None when no capture matches the workload's size."""
    import glob
    root = os.path.dirname(os.path.abspath(__file__))
    for path in sorted(glob.glob(os.path.join(root, "profiles", "ncu_full_r*.json")), reverse=True):
        try:
            cap = json.load(open(path))
            if workload == "c4" and cap.get("half1") and cap.get("particles", 10_000_000) == n:
                return cap["half1"]["dram_bytes"], os.path.relpath(path, root) + " (half1, bytes per launch)"
        except (OSError, ValueError, KeyError):
            continue
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--quick", action="store_true", help="headline only: skip the call-pattern, replica, shard-check, C5 and small-system legs")
    ap.add_argument("--no-reference-cuda", action="store_true", help="skip the reference-CUDA-kernel baseline leg")
    ap.add_argument("--molecules", type=int, default=0, help="override molecules per GPU of the c4 generator (profiling runs)")
    args = ap.parse_args()
    global MOLECULES_OVERRIDE
    MOLECULES_OVERRIDE = args.molecules
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
