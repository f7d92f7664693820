#!/usr/bin/env python
"""bench.py — particle-steps/s of the SPH step engine on synthetic Keplerian-disc ICs (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--particles P] [--impl b200|reference]

One "step" = one body of the reference loop (SUMMER_SPH - Variable.f90:1120-1162): two full evaluations
(tree + density + EOS + gravity + sinks + SPH pairs), two half kicks, drift, dt ladder, h Newton-Raphson,
sink creation, accretion, bounds cull.  N = 1 workload: BASELINE.json configs[3] (Keplerian disc, 16M gas
particles + central sink, variable h, FP64).  Prints ONE JSON line (see README / DESIGN.md §Measurement); it is
the LAST line of stdout: with NCCL_DEBUG=VERSION in the environment (the GPU boxes set it) NCCL itself prints its
version banner on stdout once when the first communicator of a multi-rank run is created.

--impl reference times the CPU implementation of the same path (the oracle port of the Fortran loops —
no Fortran compiler exists in this image, so the reference itself cannot be built) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

BYTES_PER_PARTICLE_STEP = 1960.0        # compulsory HBM bytes per particle-step
BYTES_SPH_PER_LAUNCH = 144.0            # SPH force pass, per particle per launch
BYTES_DENSITY_PER_LAUNCH = 80.0
BYTES_GRAVITY_PER_LAUNCH = 120.0
FLOPS = {"density_candidate": 44, "sph_pair": 173, "grav_opened": 13, "grav_accepted": 35, "sink_gas": 20}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def make_ics(n):
    from summersph_b200 import ics
    b, s = ics.keplerian_disc(n, seed=20251018)
    return b, s


def cpu_reference_rate(n_sample, steps, warmup, threads):
    """Oracle port timed on the host cores: particle-steps/s on a bounded disc sample."""
    from summersph_b200 import default_params, MODE_VARIABLE_H
    from oracle.oracle import Oracle
    p = default_params(MODE_VARIABLE_H)
    b, s = make_ics(n_sample)
    o = Oracle(p, threads=threads)
    o.upload(b, s)
    dt, t = 0.01, 0.0
    for _ in range(warmup):
        dt, t = o.step(dt, t)
    t0 = time.perf_counter()
    for _ in range(steps):
        dt, t = o.step(dt, t)
    el = time.perf_counter() - t0
    n, _ = o.sizes()
    return n * steps / el, el / steps


REF_THREADS = 16      # fixed, so that records taken on boxes with different core counts are comparable


def run_reference(args, rank):
    if rank != 0:
        return
    threads = min(REF_THREADS, os.cpu_count() or 1)
    n_sample = args.ref_particles
    rate, sec = cpu_reference_rate(n_sample, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "particle-steps/s", "value": rate, "unit": "particle-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"Keplerian disc {args.particles} gas + 1 sink, variable h (BASELINE configs[3]); each step timed on a {n_sample}-particle sample of it"},
        "cpu_baseline": {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                         "sample": f"{n_sample}-particle Keplerian disc, {args.steps} full steps, OpenMP on density/gravity/h-iteration/SPH loops; restatement, not gfortran (no Fortran compiler in the image)"},
        "e2e": {"value": rate, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--particles", type=int, default=16_000_000)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ref-particles", type=int, default=1_000_000)
    ap.add_argument("--cpu-baseline-particles", type=int, default=100_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (large multi-GPU sizing runs)")
    ap.add_argument("--drift", action="store_true", help="add the energy / momentum / angular-momentum drift over the timed steps "
                    "(sph_conserved before and after them, outside the timed region) as a \"drift\" key")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from summersph_b200 import default_params, MODE_VARIABLE_H, Bodies, Sinks
    from summersph_b200.state import GAS_FIELDS, SINK_FIELDS
    from summersph_b200.engine import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the SPH step)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    p = default_params(MODE_VARIABLE_H)
    p.n_ranks = world
    n = args.particles
    b, s = make_ics(n)
    s.radius[:] = p.sink_radius
    # pinned host buffers (torch is only the allocator here)
    pin = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in GAS_FIELDS}
    for k in GAS_FIELDS:
        pin[k].numpy()[:] = getattr(b, k)
    hb = Bodies(*[pin[k].numpy() for k in GAS_FIELDS])
    out_pin = {k: torch.empty(n, dtype=torch.float64).pin_memory() for k in GAS_FIELDS}
    ob = Bodies(*[out_pin[k].numpy() for k in GAS_FIELDS])

    e = Engine(p, device=local)
    if world > 1:
        from summersph_b200.parallel import init_comm, torch_broadcast_bytes
        init_comm(e, rank, world, torch_broadcast_bytes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ---------------------------------------------------------
    e.upload(hb, s)
    dt, t = 0.01, 0.0
    for _ in range(args.warmup):
        dt, t = e.step(dt, t)
    stage_acc = {}
    cons_first = e.conserved() if args.drift else None      # builds / reuses the tree; never changes a later step
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = e.launch_count()
    e.timer_start()
    for _ in range(args.steps):
        dt, t = e.step(dt, t)
        for k, v in e.stage_times().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    ms = e.timer_stop()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = e.launch_count() - l0
    drift = None
    if args.drift:
        from summersph_b200._abi import drift_report
        drift = drift_report(cons_first, e.conserved())
    # interaction counts of the reference algorithm (every leaf-box candidate): one untimed evaluation with exact
    # counters on the state the timed steps ended in; the timed steps cull exact-zero pairs before counting them
    e.set_exact_counters(True); e.evaluate(); counters = e.counters(); e.set_exact_counters(False)
    n_now, ns_now = e.sizes()
    tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
    per_rank = [ms]
    if world > 1:
        allms = [torch.zeros_like(tm) for _ in range(world)]
        dist.all_gather(allms, tm)
        per_rank = [float(x.item()) for x in allms]
        walk = torch.tensor([stage_acc.get("density", 0) + stage_acc.get("gravity", 0) + stage_acc.get("sph", 0)], dtype=torch.float64, device="cuda")
        allw = [torch.zeros_like(walk) for _ in range(world)]
        dist.all_gather(allw, walk)
        per_rank_walk = [float(x.item()) / args.steps for x in allw]
        comm = torch.tensor([stage_acc.get("comm", 0)], dtype=torch.float64, device="cuda")
        allc = [torch.zeros_like(comm) for _ in range(world)]
        dist.all_gather(allc, comm)
        per_rank_comm = [float(x.item()) / args.steps for x in allc]
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms = float(tm.item())
    value = n * args.steps / (ms * 1e-3)

    # ---- end to end through the C-ABI with host buffers ("e2e") -----------------------------------------
    # The host owns the state between steps (what the Fortran loop does): every step uploads the state it
    # got back from the previous step, advances it, and downloads the result.
    e2e_steps = 0 if args.no_e2e else max(1, min(args.steps, 3))
    hs = Sinks.empty(e.sizes()[1])
    if e2e_steps:
        e.download(into=(ob, hs))             # current state -> pinned host (untimed)
    dt2, t2 = dt, t
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        nb_now = e.sizes()[0]
        view = Bodies(*[getattr(ob, k)[:nb_now] for k in GAS_FIELDS])
        e.upload(view, hs)                    # H2D of the step's inputs from pinned host memory
        dt2, t2 = e.step(dt2, t2)
        hs = Sinks.empty(e.sizes()[1])
        e.download(into=(ob, hs))             # D2H of the step's result (the new state)
    barrier()
    e2e_sec = time.perf_counter() - t0
    te = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n * e2e_steps / float(te.item()) if e2e_steps else None
    h2d = 10 * 8 * n + 8 * 8 * len(s)
    d2h = 10 * 8 * n + 8 * 8 * len(s)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel = the stage with the largest share of the step
        per_launch = {"sph": (BYTES_SPH_PER_LAUNCH, "k_force"), "density": (BYTES_DENSITY_PER_LAUNCH, "k_density"), "gravity": (BYTES_GRAVITY_PER_LAUNCH, "k_gravity")}
        dom = max(per_launch, key=lambda k: stage_acc.get(k, 0.0))
        dom_ms = stage_acc[dom] / (2 * args.steps)           # two launches per step
        achieved = per_launch[dom][0] * (n / world) / (dom_ms * 1e-3) / 1e9       # per-rank launch processes n/world targets
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
            # ncu capture was taken at tr["particles"]; DRAM traffic of the walk kernels scales with N
            traffic = tr["dram_bytes_per_launch"][per_launch[dom][1]] / tr["particles"] * n / max(world, 1)
        except Exception:
            pass
        fp64_peak = e.fp64_peak() * world
        flops_eval = (counters["density_candidates"] * FLOPS["density_candidate"] + counters["sph_pairs"] * FLOPS["sph_pair"]
                      + counters["grav_opened"] * FLOPS["grav_opened"] + counters["grav_accepted"] * FLOPS["grav_accepted"]
                      + n * ns_now * FLOPS["sink_gas"])
        step_flops = 2.0 * flops_eval
        line = {
            "metric": "particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"Keplerian disc {n} gas + 1 central sink, variable h, eta 1.2, theta 0.5 (BASELINE configs[3])",
                       "particles": n, "mode": "variable_h", "parallelism": f"{world} rank(s)",
                       "l2": "inputs (>=1.3 GB state + tree) exceed the 126 MB L2; no flush needed"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "what": "sph_upload (pinned host SoA) + sph_step + sph_download (ascending number order) per step"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": per_launch[dom][1], "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "traffic_note": "bytes/launch; dram__bytes_read+write from the 2M-particle ncu capture scaled linearly to this N", "peak_source": peak_src,
                         "algorithmic_bytes_per_particle_per_launch": per_launch[dom][0], "launch_ms": dom_ms,
                         "note": "walk kernels are FP64-pipe / latency bound (SURVEY.md §8(d)); see fp64"},
            "fp64": {"achieved_tflops": step_flops / (ms / args.steps * 1e-3) / 1e12, "peak_tflops": fp64_peak,
                     "frac": step_flops / (ms / args.steps * 1e-3) / 1e12 / fp64_peak, "algorithmic_flops_per_step": step_flops,
                     "how": "reference-expression flop counts x interaction counters (SURVEY.md §8(d)); peak = in-library FMA probe"},
            "hbm_step": {"achieved": BYTES_PER_PARTICLE_STEP * n / (ms / args.steps * 1e-3) / 1e9, "peak": hbm_peak,
                         "frac": BYTES_PER_PARTICLE_STEP * n / (ms / args.steps * 1e-3) / 1e9 / hbm_peak, "unit": "GB/s"},
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_acc.items()},
            "per_rank_ms_per_step": [v / args.steps for v in per_rank],
            "counters": counters,
        }
        if drift is not None:
            line["drift"] = dict(drift, over_steps=args.steps)
        if world > 1:
            line["per_rank_walk_ms_per_step"] = per_rank_walk
            line["per_rank_comm_ms_per_step"] = per_rank_comm
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            nb = args.cpu_baseline_particles
            rate, sec = cpu_reference_rate(nb, 1, 0, threads)
            line["cpu_baseline"] = {"value": rate, "unit": "particle-steps/s", "cores": threads, "kind": "port",
                                    "sample": f"{nb}-particle Keplerian disc (same generator/seed), 1 full step, oracle port with OpenMP; restatement, not gfortran"}
        print(json.dumps(line), flush=True)
    e.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
