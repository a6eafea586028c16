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

WORKLOADS = {
    "c4": "C4 synthetic 10M-particle Drude system (2.5M 4-particle molecules, 2.5M Drude pairs), G=4, M=3, S=20, "
          "COM group on, hard wall 0.02 nm, fixed fp32 SoA forces",
    "c4-wall": "C4 with the frozen-spring pair forces of SURVEY.md 8d: every Drude pair hits the hard wall on every step "
               "(hard-wall stress case)",
    "c5": "C5 synthetic 200M-particle Drude system sharded over the ranks, G=4",
    "c1": "C1 NaCl 1M box shape, N=2500, G=2", "c2": "C2 SWM4-NDP 10k waters, N=50000, G=1",
    "c3": "C3 [BMIM][BF4]-like 1000 ion pairs, N=45000, G=3",
}


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
    if workload in ("c4", "c4-wall"):
        mol = MOLECULES_OVERRIDE or C4_MOLECULES
        return synth.water_box(mol, 4, first_molecule=rank * mol, box_molecules=world * mol,
                               pair_force="frozen_spring" if workload == "c4-wall" else "common")
    if workload == "c5":
        per = C5_MOLECULES // world
        return synth.water_box(per, 4, first_molecule=rank * per, box_molecules=C5_MOLECULES)
    if workload == "c1":
        return synth.nacl_box()
    if workload == "c2":
        return synth.swm4_box(10000)
    if workload == "c3":
        return synth.ionic_liquid(1000)
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
    if workload in ("c4", "c4-wall", "c5"):
        return synth.water_box(250_000, 4, box_molecules=C4_MOLECULES, pair_force="frozen_spring" if workload == "c4-wall" else "common")
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
        "config": {"workload": WORKLOADS[args.workload], "sample": sample},
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
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        ids = [capi.Comm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        comm = capi.Comm(ids[0], world, rank, local)

    system = make_system(args.workload, rank, world)
    n = system.num_particles
    padded = ((n + 31) // 32) * 32
    h_velm = torch.zeros((padded, 4), dtype=torch.float32).pin_memory()
    h_posq = torch.zeros((padded, 4), dtype=torch.float32).pin_memory()
    h_force = torch.zeros((3, padded), dtype=torch.float32).pin_memory()
    h_velm[:n] = torch.from_numpy(system.velm_f32())
    h_posq[:n] = torch.from_numpy(system.posq_f32())
    h_force[:, :n] = torch.from_numpy(np.ascontiguousarray(system.forces.T, np.float32))
    velm, posq, force = h_velm.to(dev), h_posq.to(dev), h_force.to(dev)
    h = capi.Handle(system, padded=padded, device=local, comm=comm)
    tstream = torch.cuda.Stream(device=dev)            # the step runs on its own (non-default) stream
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

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
    h.step(velm.data_ptr(), posq.data_ptr(), force.data_ptr(), max(args.warmup, 3), stream)
    launches0 = h.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    h.step(velm.data_ptr(), posq.data_ptr(), force.data_ptr(), args.steps, stream)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = h.launch_count - launches0
    # per-kernel durations for the roofline object: the same K steps once more, every streaming launch bracketed by
    # CUDA events on the launching stream (kept out of the pass above: an event between two launches serialises
    # them and removes the programmatic-dependent-launch overlap the product path runs with)
    h.set_profiling(True)
    h.step(velm.data_ptr(), posq.data_ptr(), force.data_ptr(), args.steps, stream)
    barrier()
    prof = h.profile()
    h.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    total_particles = n * world
    value = total_particles * args.steps / (ms * 1e-3)

    # ---- end to end through the C-ABI with HOST buffers: every step copies velm/posq/force in from pinned
    #      memory, runs one step, and copies velm/posq + the 2*KE vector back (tgnh_step_host) ----
    e2e_steps = 0 if args.no_e2e else max(3, min(args.steps, 10))
    ke2 = h.kinetic_energies()
    if e2e_steps:
        h.step_host(h_velm.data_ptr(), h_posq.data_ptr(), h_force.data_ptr(), 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ke2 = h.step_host(h_velm.data_ptr(), h_posq.data_ptr(), h_force.data_ptr(), 1)
    barrier()
    e2e_s = max(time.perf_counter() - t0, 1e-9)
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = total_particles * e2e_steps / e2e_s
    h2d = n * 16 * 2 + 3 * padded * 4
    d2h = n * 16 * 2 + 8 * h.T

    if rank == 0:
        peak, peak_src = peaks()
        a_ms, a_cnt = prof["half1"]
        b_ms, b_cnt = prof["half2"]
        ach = ALG_BYTES_HALF1 * n / (a_ms / max(a_cnt, 1) * 1e-3) / 1e9 if a_cnt else None
        traffic, traffic_src = ncu_traffic(args.workload, n)
        out = {
            "metric": "TGNH step particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.workload == "c5" else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload], "particles_per_gpu": n, "total_particles": total_particles,
                       "parallelism": f"particle-range shards x{world}" if world > 1 else "single GPU",
                       "exchange": {0: "none", 1: "ncclAllReduce of double[T] per step", 2: "peer-mapped inboxes over NVLink (no collective launch)"}[h.exchange_kind],
                       "l2": "inputs larger than L2 (>=440 MB working set per GPU vs 126 MB L2)",
                       "spin_up": "0.5 s of device-to-device copies before the warm-up steps (clock ramp after the host-side set-up)",
                       "accumulation": "fp32 state (OpenMM single-precision layouts), fp64 KE reductions and NH chain",
                       "step_achieved_gbs": ALG_BYTES_STEP * n * args.steps / (ms * 1e-3) / 1e9,
                       "step_frac_of_peak": ALG_BYTES_STEP * n * args.steps / (ms * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "tgnh_step_host: pinned host velm/posq/force -> device, 1 step, velm/posq/KE back, per step"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "tgnh_stream_kernel<KIND_A> (scale+kick+drift+hard wall)",
                         "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_particle": ALG_BYTES_HALF1,
                         "avg_launch_ms": a_ms / max(a_cnt, 1),
                         "half2_avg_launch_ms": b_ms / max(b_cnt, 1),
                         "half2_achieved": (ALG_BYTES_HALF2 * n / (b_ms / max(b_cnt, 1) * 1e-3) / 1e9) if b_cnt else None},
            "ke2_last": [float(x) for x in ke2],
        }
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
    h.close()
    if comm is not None:
        comm.close()
        dist.destroy_process_group()


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
