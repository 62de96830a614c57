#!/usr/bin/env python
"""Benchmark of the per-tick agent hot path on the BASELINE.json headline configuration.

Workload: 1,000,000 agents x 32 rays x 1,000 wall segments in a
4096 x 4096 arena, synthetic (numpy PCG64 seeds 0/1 for walls / positions, threefry PRNGKey(0)
for the per-agent policy weights), 90 % of the slots active.  A *step* is one tick of the hot
path over all agents:

    K1  fgx_raycast_walls   ray_generator + ray_wall_sensor + min/argmin   (FP32-bound)
    K2  fgx_forager_step    WilsonCowan policy (40 neurons, obs = 32 ray distances + own
                            energy, 2 actions, per-agent weights) + physics   (HBM-bound)

Metric: ray-wall tests/sec = agents x rays x walls x ticks / seconds (algorithmic pairs); the same
line also carries agent-steps/sec.  With N GPUs the agents are sharded by index (strong scaling:
the 1M agents are split), walls are replicated, no data-path collective is needed for this tick.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference path
(oracle/c, OpenMP on all host cores): the reference itself needs jax, which is not installed in
this image.
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
N_NEURONS, N_OBS, N_ACT = 40, N_RAYS + 1, 2
DT = 0.1
FLOP_PER_TEST = 20  # 4 orientation determinants x 5 FP32 ops, no FMA credit
# K2 algorithmic bytes per ACTIVE agent: weights J,E,D,tau,B + Z read, obs row + 6 state scalars read,
# Z + 11 state scalars + age written (SURVEY.md 8(d) "Forager step"); 4 B (the flag) per inactive slot
K2_BYTES_ACTIVE = 4 * (N_NEURONS * N_NEURONS + N_NEURONS * N_OBS + N_ACT * N_NEURONS + 3 * N_NEURONS  # weights + Z
                       + N_RAYS + 8                                                                  # obs, state in
                       + N_NEURONS + 12) + 4                                                         # Z, state out
METRIC = "ray-wall tests/sec"
UNIT = "tests/s"
WORKLOAD = ("C4: 1M agents x 32 rays x 1000 walls, 4096^2 arena, WilsonCowan forager 40 neurons / 33 obs "
            "(SURVEY.md 8d)")


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


def wall_array(work):
    w = work["walls"]
    return np.stack([np.stack([w[0], w[1]], 1), np.stack([w[2], w[3]], 1)], 1)  # [W, 2, 2]


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


def measured_traffic(kernel, n_local):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r01_traffic.json),
    scaled to this rank's shard; None if there is no capture."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(p):
        return None
    t = json.load(open(p)).get(kernel)
    return None if not t else t["dram_bytes_per_launch"] * n_local / t["agents"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class CpuTick:
    """The same tick on the CPU: the C restatement of the reference path (oracle/c, OpenMP on every
    host core) over the first `n_sample` agents of the workload."""

    def __init__(self, work, n_sample):
        from oracle import agent_methods as oam
        from oracle import c_oracle
        from oracle import foraging as ofo
        from oracle import jax_random as jr
        self.c = c_oracle
        self.n = n_sample
        self.cfg = ofo.Config("sim", n_neurons=N_NEURONS, n_obs=N_OBS, n_act=N_ACT, x_max=ARENA, y_max=ARENA)
        self.fg = oam.create_agents(lambda p, *a: ofo.forager_create(self.cfg, *a), None, n_sample, n_sample, 1,
                                    jr.PRNGKey(0))
        st = self.fg["state"]
        st["X"], st["Y"], st["ang"] = (work[k][:n_sample].reshape(-1, 1).copy() for k in ("x", "y", "ang"))
        self.fg["active_state"] = work["active"][:n_sample].copy()
        self.walls = work["walls"]
        self.zero = np.zeros(n_sample, np.float32)
        self.tests = float(n_sample) * N_RAYS * len(self.walls[0])
        c_oracle.raycast_walls(work["x"][:64], work["y"][:64], work["ang"][:64], None, N_RAYS, RAY_SPAN, RAY_LEN,
                               *self.walls)  # load the library, spin up the OpenMP pool

    def step(self):
        st = self.fg["state"]
        t = time.perf_counter()
        dist, _ = self.c.raycast_walls(st["X"], st["Y"], st["ang"], self.fg["active_state"], N_RAYS, RAY_SPAN,
                                       RAY_LEN, *self.walls)
        self.c.forager_step(self.cfg, dist, self.zero, self.fg)
        return time.perf_counter() - t


def cpu_baseline(work, n_sample):
    tick = CpuTick(work, n_sample)
    sec = tick.step()
    return {"value": tick.tests / sec, "unit": UNIT, "cores": tick.c.threads(), "kind": "port",
            "sample": f"one tick (ray cast + forager step) over the first {n_sample} agents x {N_RAYS} rays x "
                      f"{N_WALLS} walls of the same workload ({sec:.2f} s); C restatement of the reference path "
                      "with OpenMP (jax is not installed, so the reference itself cannot run)"}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    work = make_workload()
    n_sample = 20_000
    tick = CpuTick(work, n_sample)
    times = []
    for i in range(args.warmup + args.steps):
        dt = tick.step()
        if i >= args.warmup:
            times.append(dt)
    sec = float(np.mean(times))
    val = tick.tests / sec
    sample = (f"each step = one tick (ray cast + forager step) over the first {n_sample} agents x {N_RAYS} rays x "
              f"{N_WALLS} walls; C restatement of the reference path (oracle/c, OpenMP); the reference needs jax, "
              "absent from this image")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "tick": ["raycast_walls", "forager_step"], "step_sample": sample},
        "agent_steps_per_sec": n_sample / sec,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": tick.c.threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def build_foragers(work, rank, world, dev):
    """The local shard (slots rank, rank+world, ...: round-robin, so that active and inactive slots are
    spread evenly over the ranks) of the 1M-forager set: native create (threefry weights, PRNGKey(0)),
    then positions / headings of the synthetic workload.  Every rank builds exactly its rows of the
    single-GPU set."""
    import torch

    from foragax_b200 import random as fgx_random
    from foragax_b200.base.agent_classes import Agent_Set, Params
    from foragax_b200.base.agent_methods import create_agents
    from foragax_b200.base.space_methods import create_space
    from foragax_b200.examples.foraging import Forager
    space = create_space(0.0, ARENA, 0.0, ARENA, False, wall_array(work))
    fset = Agent_Set(agent=Forager, num_total_agents=N_AGENTS, num_active_agents=int(0.9 * N_AGENTS), agent_type=1)
    cp = Params(content={'policy_params': {'num_neurons': N_NEURONS, 'num_obs': N_OBS, 'num_actions': N_ACT, 'dt': DT,
                                           'action_scale': 1.0, 'deterministic': True}, 'dt': DT, 'space': space})
    n_local = len(range(rank, N_AGENTS, world))
    fset.agents = create_agents(cp, fset, fgx_random.PRNGKey(0), shard=(rank, n_local, world))
    st = fset.agents.state.content
    for k, src in (("X", "x"), ("Y", "y"), ("ang", "ang")):
        st[k].copy_(torch.from_numpy(np.ascontiguousarray(work[src][rank::world])).to(dev).reshape(-1, 1))
    return fset, space


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
    from foragax_b200.base.agent_classes import Params, Signal
    from foragax_b200.base.agent_methods import step_agents

    work = make_workload()
    mine = slice(rank, None, world)  # round-robin sharding by slot index
    n_local = len(range(rank, N_AGENTS, world))
    n_active_local = int(work["active"][mine].sum())
    fset, space = build_foragers(work, rank, world, dev)
    F = fset.agents
    st = F.state.content
    dist_out = torch.empty((n_local, N_RAYS), dtype=torch.float32, device=dev)
    hit_out = torch.empty((n_local, N_RAYS), dtype=torch.int32, device=dev)
    energy_in = torch.zeros(n_local, dtype=torch.float32, device=dev)
    step_params = Params(content={'space': space, 'repr_thresh': 5.0, 'die_thresh': 0.0, 'energy_min': -1.0,
                                  'energy_max': 15.0})
    step_input = Signal(content={'obs': dist_out, 'energy_in': energy_in})
    launches_per_step = 2  # culled ray cast (its wall grid is built once and cached), forager step
    BOUNDS = (0.0, ARENA, 0.0, ARENA)

    def k1():
        # default product path: the spatially culled kernel (bit-identical to all-pairs; tests/test_gpu_raycast.py)
        sm.raycast_walls((st["X"], st["Y"], st["ang"]), RAY_SPAN, RAY_LEN, space.walls, active_state=F.active_state,
                         n_rays=N_RAYS, out=(dist_out, hit_out), bounds=BOUNDS)

    def k1_all_pairs():
        sm.raycast_walls((st["X"], st["Y"], st["ang"]), RAY_SPAN, RAY_LEN, space.walls, active_state=F.active_state,
                         n_rays=N_RAYS, out=(dist_out, hit_out))

    def k2():
        fset.agents = step_agents(step_params, step_input, fset)

    def tick():
        k1()
        k2()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    stream = torch.cuda.current_stream()
    sampler = ClockSampler(local) if rank == 0 else None
    warm = max(args.warmup, 3)
    for _ in range(warm):
        tick()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        tick()
    e1.record(stream)
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    tests_per_step = float(N_AGENTS) * N_RAYS * N_WALLS
    value = tests_per_step / (ms_step * 1e-3)

    # per-kernel durations, live, with CUDA events on the launching stream
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    k1_ms, k2_ms = [], []
    for _ in range(args.steps):
        ev[0].record(stream)
        k1()
        ev[1].record(stream)
        k2()
        ev[2].record(stream)
        ev[2].synchronize()
        k1_ms.append(ev[0].elapsed_time(ev[1]))
        k2_ms.append(ev[1].elapsed_time(ev[2]))
    k1_ms, k2_ms = float(np.mean(k1_ms)), float(np.mean(k2_ms))

    # the same tick with the all-pairs ray cast (every (ray, wall) pair evaluates all four determinants): the
    # FP32-bound kernel whose roofline fraction is reported next to the culled default
    ap_k1, ap_tick = [], []
    k1_all_pairs()
    for _ in range(max(3, min(args.steps, 10))):
        ev[0].record(stream)
        k1_all_pairs()
        ev[1].record(stream)
        k2()
        ev[2].record(stream)
        ev[2].synchronize()
        ap_k1.append(ev[0].elapsed_time(ev[1]))
        ap_tick.append(ev[0].elapsed_time(ev[2]))
    ap_k1_ms, ap_tick_ms = max_over_ranks(float(np.mean(ap_k1))), max_over_ranks(float(np.mean(ap_tick)))

    # end to end through the public API with HOST buffers: every step copies the agents' pose and active
    # flags from pinned host memory (H2D), runs the tick and reads back its result (ray distances, hit
    # indices and the stepped positions, D2H).  The result read-back of step i runs on a side stream
    # into pinned memory while step i+1 computes (double-buffered distance / hit outputs); every byte is
    # still copied every step and the timed region ends only when the last read-back has landed.
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()  # noqa: E731
    hx, hy, ha, hact = (pin(work[k][mine]) for k in ("x", "y", "ang", "active"))
    h_dist = torch.empty((n_local, N_RAYS), dtype=torch.float32).pin_memory()
    h_hit = torch.empty((n_local, N_RAYS), dtype=torch.int32).pin_memory()
    h_x, h_y = torch.empty((n_local, 1)).pin_memory(), torch.empty((n_local, 1)).pin_memory()
    outs = [(dist_out, hit_out), (torch.empty_like(dist_out), torch.empty_like(hit_out))]
    side = torch.cuda.Stream()
    ev_done = [torch.cuda.Event() for _ in range(2)]
    ev_xy = torch.cuda.Event()
    ev_out = [torch.cuda.Event() for _ in range(2)]
    state = {"i": 0}

    def e2e_step():
        i = state["i"]
        b = i & 1
        if i >= 1:
            stream.wait_event(ev_xy)          # the previous step's positions have been read back
        st["X"].copy_(hx.view(-1, 1), non_blocking=True)
        st["Y"].copy_(hy.view(-1, 1), non_blocking=True)
        st["ang"].copy_(ha.view(-1, 1), non_blocking=True)
        F.active_state.copy_(hact, non_blocking=True)
        if i >= 2:
            stream.wait_event(ev_out[b])      # this output buffer's previous read-back has landed
        d_o, h_o = outs[b]
        sm.raycast_walls((st["X"], st["Y"], st["ang"]), RAY_SPAN, RAY_LEN, space.walls, active_state=F.active_state,
                         n_rays=N_RAYS, out=(d_o, h_o), bounds=BOUNDS)
        fset.agents = step_agents(step_params, Signal(content={'obs': d_o, 'energy_in': energy_in}), fset)
        ev_done[b].record(stream)
        with torch.cuda.stream(side):
            side.wait_event(ev_done[b])
            h_x.copy_(st["X"], non_blocking=True)
            h_y.copy_(st["Y"], non_blocking=True)
            ev_xy.record(side)
            h_dist.copy_(d_o, non_blocking=True)
            h_hit.copy_(h_o, non_blocking=True)
            ev_out[b].record(side)
        state["i"] = i + 1

    for _ in range(2):
        e2e_step()
    stream.wait_stream(side)
    barrier()
    state["i"] = 0
    e0.record(stream)
    for _ in range(args.steps):
        e2e_step()
    stream.wait_stream(side)
    e1.record(stream)
    barrier()
    e2e_ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    h2d = 4 * 4 * N_AGENTS
    d2h = (2 * 4 * N_RAYS + 2 * 4) * N_AGENTS
    clocks = sampler.stop() if sampler else None

    # extra stage (not part of `value`): agent-agent ray sensing -- NCCL all-gather of float4 positions
    # (N > 1), uniform-grid build, ray_agent_sensor over the 3x3 neighbourhood, merge with the wall result
    agent_sensing = None
    if not args.no_agent_sensing:
        from foragax_b200 import distributed as fd

        def sense():
            k1()
            tg = fd.all_gather_positions(st["X"], st["Y"], st["rad"], F.active_state)
            sm.raycast_agents((st["X"], st["Y"], st["ang"]), RAY_SPAN, RAY_LEN, tg, (0.0, ARENA, 0.0, ARENA), 15.0,
                              active_state=F.active_state, n_rays=N_RAYS, merge_walls=True, out=(dist_out, hit_out))

        sense()
        barrier()
        reps = max(2, min(5, args.steps))
        e0.record(stream)
        for _ in range(reps):
            sense()
        e1.record(stream)
        barrier()
        sense_ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        frac_agent = float(((hit_out >= 0) & (hit_out < N_AGENTS)).float().mean())
        agent_sensing = {"ms_per_tick_walls_plus_agents": sense_ms, "ms_agents_only": sense_ms - k1_ms,
                         "rays_hitting_an_agent": frac_agent,
                         "note": "fgx_raycast_walls + all_gather_positions + fgx_raycast_agents (grid, exact "
                                 "vs all-pairs); 1M agents sensing 1M targets, not included in `value`"}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        peak_tflops = sms * 128 * sm_max * 1e6 / 1e12  # non-FMA FP32 lanes x clock
        # all-pairs ray cast: FP32 / issue bound, 20 algorithmic FLOP per (ray, wall) test
        ach = FLOP_PER_TEST * (float(n_local) * N_RAYS * N_WALLS) / (ap_k1_ms * 1e-3) / 1e12
        roof_ap = {"bound": "fp32", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
                   "traffic": measured_traffic("raycast_walls_kernel", n_local),
                   "algorithmic_bytes": float(n_local) * (16 + 8 * N_RAYS),
                   "kernel": "raycast_walls_kernel<16,1>", "kernel_ms": ap_k1_ms,
                   "share_of_step": ap_k1_ms / ap_tick_ms,
                   "peak_basis": f"{sms} SMs x 128 FP32 lanes x {sm_max:.0f} MHz (sm_max_mhz, {peak_src} "
                                 "MEASURED_PEAKS.json), non-FMA; 20 FLOP per ray-wall test (SURVEY.md 8d)"}
        if clocks and clocks.get("sm_mhz"):
            roof_ap["frac_at_sampled_clock"] = ach / (sms * 128 * clocks["sm_mhz"] * 1e6 / 1e12)
        all_pairs = {"ms_per_step": ap_tick_ms, "value": tests_per_step / (ap_tick_ms * 1e-3), "unit": UNIT,
                     "agent_steps_per_sec": N_AGENTS / (ap_tick_ms * 1e-3), "roofline": roof_ap,
                     "note": "same tick with fgx_raycast_walls (no cull: every (ray, wall) pair evaluates all four "
                             "orientation determinants); outputs identical to the default path"}
        # the default tick's dominant kernel is the forager step (HBM bound)
        k2_bytes = K2_BYTES_ACTIVE * n_active_local + 4 * (n_local - n_active_local)
        k2_gbs = k2_bytes / (k2_ms * 1e-3) / 1e9
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        roof = {"bound": "hbm", "achieved": k2_gbs, "peak": hbm, "unit": "GB/s", "frac": k2_gbs / hbm,
                "traffic": measured_traffic("forager_step_kernel", n_local), "algorithmic_bytes": float(k2_bytes),
                "kernel": "forager_step_kernel<40,33,2>", "kernel_ms": k2_ms,
                "share_of_step": k2_ms / (k1_ms + k2_ms),
                "bytes_per_active_agent": K2_BYTES_ACTIVE, "peak_basis": f"hbm_gbs, {peak_src} MEASURED_PEAKS.json"}
        # the culled ray cast is instruction-issue bound: warp instructions (ncu, committed under profiles/) against
        # the issue rate of 4 schedulers per SM
        rc_prof = None
        tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tp):
            rc_prof = json.load(open(tp)).get("raycast_walls_grid_kernel")
        rc_roof = None
        if rc_prof:
            winst = rc_prof["warp_instructions_per_launch"] * n_local / rc_prof["agents"]
            issue_peak = sms * 4 * sm_max * 1e6
            rc_roof = {"bound": "issue", "achieved": winst / (k1_ms * 1e-3) / 1e9, "peak": issue_peak / 1e9,
                       "unit": "G warp-instructions/s", "frac": winst / (k1_ms * 1e-3) / issue_peak,
                       "traffic": rc_prof["dram_bytes_per_launch"] * n_local / rc_prof["agents"],
                       "peak_basis": f"{sms} SMs x 4 warp schedulers x {sm_max:.0f} MHz; instruction count from "
                                     "profiles/r01_traffic.json (ncu smsp__inst_executed.sum of this workload)"}
        roof_rc = {"kernel": "raycast_walls_grid_kernel (wall lists built once, cached)", "kernel_ms": k1_ms,
                   "roofline": rc_roof,
                   "share_of_step": k1_ms / (k1_ms + k2_ms),
                   "algorithmic_tests_per_sec": float(n_local) * N_RAYS * N_WALLS / (k1_ms * 1e-3),
                   "speedup_vs_all_pairs_kernel": ap_k1_ms / k1_ms,
                   "note": "spatial cull: per grid cell only the walls that can matter are tested (~14 of 1000 here); "
                           "bit-identical to the all-pairs kernel, so a FLOP roofline no longer describes it"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "agents": N_AGENTS, "active_agents": int(0.9 * N_AGENTS), "rays": N_RAYS,
                       "walls": N_WALLS, "sharding": f"agents by slot index, round-robin over {world} rank(s), walls replicated",
                       "tick": ["fgx_raycast_walls_grid (ray_generator + ray_wall_sensor + min/argmin over the walls a "
                                "uniform grid lists for the agent's cell; results bit-identical to all-pairs)",
                                "fgx_forager_step (WilsonCowan policy + physics)"],
                       "l2": "per-step working set 11.6 GB of per-agent weights + 272 MB ray I/O >> 126 MB L2"},
            "agent_steps_per_sec": N_AGENTS / (ms_step * 1e-3),
            "roofline": roof, "raycast_stage": roof_rc, "all_pairs": all_pairs,
            "e2e": {"value": tests_per_step / (e2e_ms_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_step,
                    "agent_steps_per_sec": N_AGENTS / (e2e_ms_step * 1e-3)},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if agent_sensing:
            line["agent_sensing"] = agent_sensing
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(work, args.cpu_sample)
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
    ap.add_argument("--no-agent-sensing", action="store_true", help="skip the extra agent-agent sensing stage")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
