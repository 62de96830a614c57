#!/usr/bin/env python
"""Benchmark of the per-tick agent hot path on the BASELINE.json headline configuration.

Workload: 1,000,000 agents x 32 rays x 1,000 wall segments in a
4096 x 4096 arena, synthetic (numpy PCG64, seeds 0/1).  A *step* is one tick of the hot
path over all agents.  Metric: ray-wall tests/sec = agents x rays x walls x ticks / seconds
(algorithmic pairs).  With N GPUs the agents are sharded by index (strong scaling: the
1M agents are split), walls are replicated.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the
reference path (oracle/c, OpenMP on all host cores): the reference itself needs jax, which
is not installed in this image.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_AGENTS = 1_000_000
N_RAYS = 32
N_WALLS = 1000
ARENA = 4096.0
RAY_SPAN = float(np.float32(np.pi / 4.0))
RAY_LEN = 60.0
FLOP_PER_TEST = 20  # 4 orientation determinants x 5 FP32 ops, no FMA credit
METRIC = "ray-wall tests/sec"
UNIT = "tests/s"


def make_workload(n_agents=N_AGENTS, n_walls=N_WALLS):
    """walls seed 0 (4 boundary + interior segments), agents seed 1."""
    rw = np.random.default_rng(0)
    k = n_walls - 4
    mx, my = rw.uniform(0, ARENA, k), rw.uniform(0, ARENA, k)
    ln, th = rw.uniform(20, 200, k), rw.uniform(0, np.pi, k)
    inter = np.stack([mx - ln / 2 * np.cos(th), my - ln / 2 * np.sin(th), mx + ln / 2 * np.cos(th),
                      my + ln / 2 * np.sin(th)], axis=1)
    bound = np.array([[0, 0, ARENA, 0], [ARENA, 0, ARENA, ARENA], [0, 0, 0, ARENA], [0, ARENA, ARENA, ARENA]])
    w = np.concatenate([bound, inter]).astype(np.float32)
    ra = np.random.default_rng(1)
    x = ra.uniform(1, ARENA - 1, n_agents).astype(np.float32)
    y = ra.uniform(1, ARENA - 1, n_agents).astype(np.float32)
    ang = ra.uniform(-np.pi, np.pi, n_agents).astype(np.float32)
    rad = ra.uniform(5, 10, n_agents).astype(np.float32)
    active = np.zeros(n_agents, np.float32)
    active[: int(0.9 * n_agents)] = 1.0
    return {"x": x, "y": y, "ang": ang, "rad": rad, "active": active,
            "walls": [np.ascontiguousarray(w[:, i]) for i in range(4)]}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, smax, power, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [s.strip() for s in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0]))
                smax.append(float(c[1]))
                power.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if sm:
            # "under load" = samples whose power draw is in the upper half of what was seen
            thr = (max(power) + min(power)) / 2.0
            load = [s for s, pw in zip(sm, power) if pw >= thr] or sm
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), power_w_max=float(max(power)))
        return out


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


def cpu_baseline(work, n_sample, repeats=1):
    """The C restatement of the reference path (oracle/c), OpenMP on every host core, on the
    first `n_sample` agents of the same workload."""
    from oracle import c_oracle
    w = work["walls"]
    def sample(k):
        return (work["x"][:k], work["y"][:k], work["ang"][:k], work["active"][:k], N_RAYS, RAY_SPAN, RAY_LEN, *w)

    args = sample(n_sample)
    c_oracle.raycast_walls(*sample(256))  # load the library, spin up the OpenMP pool
    best = None
    for _ in range(repeats):
        t = time.perf_counter()
        c_oracle.raycast_walls(*args)
        dt = time.perf_counter() - t
        best = dt if best is None else min(best, dt)
    tests = float(n_sample) * N_RAYS * len(w[0])
    return {"value": tests / best, "unit": UNIT, "cores": c_oracle.threads(), "kind": "port",
            "sample": f"first {n_sample} agents x {N_RAYS} rays x {len(w[0])} walls of the same workload "
                      f"({best:.2f} s), C restatement of the reference path with OpenMP (jax is not installed, "
                      "so the reference itself cannot run)",
            "seconds": best}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    work = make_workload()
    n_sample = 20_000
    times = []
    for i in range(args.warmup + args.steps):
        b = cpu_baseline(work, n_sample)
        if i >= args.warmup:
            times.append(b["seconds"])
    tests = float(n_sample) * N_RAYS * N_WALLS
    sec = float(np.mean(times))
    val = tests / sec
    from oracle import c_oracle
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C4: 1M agents x 32 rays x 1000 walls, 4096^2 arena (SURVEY.md 8d)",
                   "tick": ["ray_generator", "ray_wall_sensor min/argmin"],
                   "step_sample": f"each step = first {n_sample} agents of the workload"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": c_oracle.threads(), "kind": "port",
                         "sample": f"first {n_sample} agents x {N_RAYS} rays x {N_WALLS} walls per step; C restatement "
                                   "of the reference path (oracle/c, OpenMP); the reference needs jax, absent here"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: foragax_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import foragax_b200.base.space_methods as sm
    from foragax_b200 import _lib as L
    from foragax_b200.base.space_classes import Point, Wall

    work = make_workload()
    lo, hi = rank * N_AGENTS // world, (rank + 1) * N_AGENTS // world
    n_local = hi - lo
    up = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    x, y, ang, act = (up(work[k][lo:hi]) for k in ("x", "y", "ang", "active"))
    walls = Wall(Point(up(work["walls"][0]), up(work["walls"][1])), Point(up(work["walls"][2]), up(work["walls"][3])))
    dist_out = torch.empty((n_local, N_RAYS), dtype=torch.float32, device=dev)
    hit_out = torch.empty((n_local, N_RAYS), dtype=torch.int32, device=dev)
    launches_per_step = 1

    def tick():
        sm.raycast_walls((x, y, ang), RAY_SPAN, RAY_LEN, walls, active_state=act, n_rays=N_RAYS,
                         out=(dist_out, hit_out))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        tick()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        tick()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    tests_per_step = float(N_AGENTS) * N_RAYS * N_WALLS
    value = tests_per_step / (ms_step * 1e-3)

    # dominant kernel (raycast_walls_kernel): per-launch duration with CUDA events on the launching stream
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms = []
    for _ in range(args.steps):
        k0.record(stream)
        tick()
        k1.record(stream)
        k1.synchronize()
        kms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kms))

    # end to end through the public API with HOST buffers: pinned H2D of the step's inputs, the tick,
    # D2H of the step's result (distances + hit indices)
    hx, hy, ha, hact = (torch.from_numpy(work[k][lo:hi].copy()).pin_memory() for k in ("x", "y", "ang", "active"))
    h_dist = torch.empty((n_local, N_RAYS), dtype=torch.float32).pin_memory()
    h_hit = torch.empty((n_local, N_RAYS), dtype=torch.int32).pin_memory()

    def e2e_step():
        x.copy_(hx, non_blocking=True)
        y.copy_(hy, non_blocking=True)
        ang.copy_(ha, non_blocking=True)
        act.copy_(hact, non_blocking=True)
        tick()
        h_dist.copy_(dist_out, non_blocking=True)
        h_hit.copy_(hit_out, non_blocking=True)

    for _ in range(2):
        e2e_step()
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    e1.record(stream)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms_step = float(t.item()) / args.steps
    h2d = 4 * 4 * n_local * world
    d2h = 2 * 4 * N_RAYS * n_local * world
    clocks = sampler.stop() if sampler else None

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        peak_tflops = sms * 128 * sm_max * 1e6 / 1e12  # non-FMA FP32 lanes x clock
        ach = FLOP_PER_TEST * (float(n_local) * N_RAYS * N_WALLS) / (kernel_ms * 1e-3) / 1e12
        roof = {"bound": "fp32", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
                "traffic": None, "kernel": "raycast_walls_kernel<8>", "kernel_ms": kernel_ms,
                "peak_basis": f"{sms} SMs x 128 FP32 lanes x {sm_max:.0f} MHz (sm_max_mhz, {peak_src} "
                              "MEASURED_PEAKS.json), non-FMA; 20 FLOP per ray-wall test (SURVEY.md 8d)"}
        if clocks and clocks.get("sm_mhz"):
            roof["frac_at_sampled_clock"] = ach / (sms * 128 * clocks["sm_mhz"] * 1e6 / 1e12)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C4: 1M agents x 32 rays x 1000 walls, 4096^2 arena (SURVEY.md 8d)",
                       "agents": N_AGENTS, "rays": N_RAYS, "walls": N_WALLS, "sharding": f"agents by index / {world}",
                       "tick": ["ray_generator", "ray_wall_sensor min/argmin (fgx_raycast_walls)"],
                       "l2": "per-step working set 272 MB (16 MB in, 256 MB out) > 126 MB L2; kernel is FP32-bound"},
            "agent_steps_per_sec": N_AGENTS / (ms_step * 1e-3),
            "roofline": roof,
            "e2e": {"value": tests_per_step / (e2e_ms_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_step},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(work, args.cpu_sample).items() if k != "seconds"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=40_000, help="agents in the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
