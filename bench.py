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
# `value` counts ALGORITHMIC pairs (agents x rays x walls per tick, inactive slots included: the reference's vmap
# evaluates every slot and every pair) whether or not the kernel culls them; the unit says so, and the
# line also carries the executed figure (`executed_tests_per_sec`) and the un-culled tick (`all_pairs`).
UNIT = "all-pairs-equivalent tests/s"
PARITY_TICKS = 3
PARITY_GOLDEN = os.path.join(ROOT, "tests", "golden", "bench_parity_digest.json")
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
    """SM clock, power and throttle reasons sampled DURING the timed regions: NVML polled every ~2 ms from a thread
    (an `nvidia-smi -lms 20` child, round 1's sampler, saw one sample of the 6 ms timed loop at N = 8); falls back to
    nvidia-smi when NVML cannot be loaded."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        import threading
        self.rows, self.p, self.t, self.via = [], None, None, None   # rows: (sm MHz, max MHz, W, reason bitmask)
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else gpu_index
            h = nv.nvmlDeviceGetHandleByIndex(phys)
            smax = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                    "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                    "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
            self._bits = bits
            self._stop = threading.Event()

            def poll():
                while not self._stop.is_set():
                    try:
                        self.rows.append((float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), smax,
                                          nv.nvmlDeviceGetPowerUsage(h) / 1000.0,
                                          int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)),
                                          float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_MEM))))
                    except nv.NVMLError:
                        pass
                    self._stop.wait(0.002)
            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            self.via = "nvml, 2 ms"
        except Exception:  # noqa: BLE001 -- any NVML problem: fall back to the nvidia-smi child
            self.t = None
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            try:
                self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}",
                                           "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                          stderr=subprocess.DEVNULL)
                self.via = "nvidia-smi -lms 20"
            except OSError:
                self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "via": self.via}
        sm, smax, power, reasons, mem = [], [], [], set(), []
        if self.t is not None:
            self._stop.set()
            self.t.join(timeout=2)
            for s_, m_, w_, r_, mc_ in self.rows:
                mem.append(mc_)
                sm.append(s_)
                smax.append(m_)
                power.append(w_)
                for name, bit in self._bits.items():
                    if r_ & bit:
                        reasons.add(name)
        elif self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
            self.f.flush()
            self.f.seek(0)
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
        else:
            return out
        if sm:
            # "under load" = samples whose power draw is in the upper half of what was seen (90th percentile instead of
            # the maximum: NVML's instantaneous power reading has isolated spikes far above the board limit)
            thr = (float(np.percentile(power, 90)) + min(power)) / 2.0
            load = [s for s, pw in zip(sm, power) if pw >= thr] or sm
            out.update(sm_mhz=float(np.median(load)), sm_max_mhz=float(max(smax)), reasons=sorted(reasons),
                       samples=len(sm), samples_under_load=len(load), power_w_p90=float(np.percentile(power, 90)),
                       power_w_max=float(max(power)))
            if mem:
                out["mem_mhz_min"] = float(min(mem))
                out["mem_mhz_median"] = float(np.median(mem))
        return out


def measured_traffic(kernel, n_local):
    """DRAM bytes per launch of `kernel` from the committed ncu capture (profiles/r02_traffic.json, else the
    round-1 file), scaled to this rank's shard; None if there is no capture."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not os.path.exists(p):
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
        self.raycast_sec = time.perf_counter() - t
        self.c.forager_step(self.cfg, dist, self.zero, self.fg)
        return time.perf_counter() - t


def cpu_baseline(work, n_sample):
    use_all_host_cores()
    tick = CpuTick(work, n_sample)
    sec = tick.step()
    return {"value": tick.tests / sec, "unit": UNIT, "cores": tick.c.threads(), "kind": "port",
            "raycast_only_tests_per_sec": tick.tests / tick.raycast_sec,
            "sample": f"one tick (ray cast + forager step) over the first {n_sample} agents x {N_RAYS} rays x "
                      f"{N_WALLS} walls of the same workload ({sec:.2f} s); C restatement of the reference path "
                      "with OpenMP (jax is not installed, so the reference itself cannot run)"}


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must use every host core at every N, so the
    variable is reset BEFORE the OpenMP runtime of the C oracle loads and the pool size is set explicitly."""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0)) or cores
    except (AttributeError, OSError):
        pass
    os.environ["OMP_NUM_THREADS"] = str(cores)
    from oracle import c_oracle
    c_oracle.set_threads(cores)
    return cores


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    use_all_host_cores()
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
    emit({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "tick": ["raycast_walls", "forager_step"], "step_sample": sample},
        "agent_steps_per_sec": n_sample / sec,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": tick.c.threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


def build_foragers(work, rank, world, dev, n_agents=None):
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
    n_agents = N_AGENTS if n_agents is None else int(n_agents)
    space = create_space(0.0, ARENA, 0.0, ARENA, False, wall_array(work))
    fset = Agent_Set(agent=Forager, num_total_agents=n_agents, num_active_agents=int(0.9 * n_agents), agent_type=1)
    cp = Params(content={'policy_params': {'num_neurons': N_NEURONS, 'num_obs': N_OBS, 'num_actions': N_ACT, 'dt': DT,
                                           'action_scale': 1.0, 'deterministic': True}, 'dt': DT, 'space': space})
    n_local = len(range(rank, n_agents, world))
    fset.agents = create_agents(cp, fset, fgx_random.PRNGKey(0), shard=(rank, n_local, world))
    st = fset.agents.state.content
    for k, src in (("X", "x"), ("Y", "y"), ("ang", "ang")):
        st[k].copy_(torch.from_numpy(np.ascontiguousarray(work[src][rank::world])).to(dev).reshape(-1, 1))
    return fset, space


def np_forager_rows(F, rows):
    """Oracle-layout dict (oracle/foraging.py) of the first `rows` local rows of a Forager pytree."""
    from oracle import foraging as ofo
    c, pc = F.state.content, F.policy.params.content
    cp = lambda t: t[:rows].cpu().numpy()  # noqa: E731
    return {"state": {k: cp(c[k]) for k in ofo.FSTATE},
            "policy": {**{k: cp(pc[k]) for k in ("J", "tau", "E", "B", "D")}, "Z": cp(F.policy.state.content["Z"])},
            "unique_id": cp(F.unique_id), "agent_type": cp(F.agent_type), "active_state": cp(F.active_state),
            "age": cp(F.age)}


def oracle_slice_check(work, F, space, dist_out, hit_out, tick, rows, dev):
    """Rank 0, BEFORE the parity ticks: one tick over this rank's first `rows` agents on the oracle -- ray cast
    bit-exact (the oracle is fed the device's ray directions, as in the parity tests), forager step within the
    1e-5 relative tolerance of north_star -- against the tick the device runs on the full shard."""
    import torch

    import foragax_b200.base.space_methods as sm
    from oracle import foraging as ofo
    from oracle import space_methods as osm
    st = F.state.content
    before = np_forager_rows(F, rows)
    old_res = sm.RESOLUTION
    sm.RESOLUTION = N_RAYS
    try:
        rays = sm.ray_generator((st["X"][:rows], st["Y"][:rows], st["ang"][:rows]), RAY_SPAN, RAY_LEN)
    finally:
        sm.RESOLUTION = old_res
    dx, dy = rays.ray_direction.x.cpu().numpy(), rays.ray_direction.y.cpu().numpy()
    x0, y0 = before["state"]["X"].reshape(-1), before["state"]["Y"].reshape(-1)
    w = work["walls"]
    ref_d = np.empty((rows, N_RAYS), np.float32)
    ref_h = np.empty((rows, N_RAYS), np.int32)
    for lo in range(0, rows, 500):  # chunks bound the [agents, rays, walls] temporaries of the NumPy oracle
        hi = min(rows, lo + 500)
        d = osm.ray_wall_sensor(x0[lo:hi, None, None], y0[lo:hi, None, None], dx[lo:hi, :, None], dy[lo:hi, :, None],
                                RAY_LEN, *[a[None, None, :] for a in w])
        ref_d[lo:hi] = d.min(-1)
        ref_h[lo:hi] = np.where(ref_d[lo:hi] < np.float32(RAY_LEN), d.argmin(-1), -1)
    inactive = before["active_state"] == 0
    ref_d[inactive] = 0
    ref_h[inactive] = -1
    tick()  # the device tick over the whole shard (first of the PARITY_TICKS)
    got_d, got_h = dist_out[:rows].cpu().numpy(), hit_out[:rows].cpu().numpy()
    cfg = ofo.Config("sim", n_neurons=N_NEURONS, n_obs=N_OBS, n_act=N_ACT, x_max=ARENA, y_max=ARENA)
    ref = ofo.forager_step(cfg, got_d, np.zeros(rows, np.float32), before)
    after = np_forager_rows(F, rows)

    def rel(name, a, b):
        scale = max(1.0, float(np.abs(b).max()))
        return float(np.abs(a.astype(np.float64) - b.astype(np.float64)).max() / scale)

    errs = {k: rel(k, after["state"][k], ref["state"][k]) for k in ("X", "Y", "X_vel", "Y_vel", "ang", "energy")}
    errs["Z"] = rel("Z", after["policy"]["Z"], ref["policy"]["Z"])
    return {"agents": rows, "raycast_dist_bit_exact": bool(np.array_equal(got_d, ref_d)),
            "raycast_hit_bit_exact": bool(np.array_equal(got_h, ref_h)),
            "rays_hitting_a_wall": float((ref_h >= 0).mean()),
            "step_max_rel_err": errs, "step_within_1e-5": bool(max(errs.values()) <= 1e-5),
            "age_bit_exact": bool(np.array_equal(after["age"], ref["age"])),
            "oracle": "oracle/space_methods.py ray_wall_sensor (device ray directions) + oracle/foraging.py forager_step"}


def c5_churn_leg(args, rank, world, dev):
    """BASELINE config 5 (10 M Dice slots, D20: ~5 % + 5 % turnover per tick) -- per-op CUDA-event times,
    algorithmic bytes and fraction of measured HBM (N = 1); with N > 1 the set is block-sharded and the tick runs
    through the peer-memory exchange (foragax_b200/peer.py)."""
    import torch
    import torch.distributed as dist

    from foragax_b200.examples import churn
    n, ticks = args.c5_slots, args.c5_ticks
    peaks, peak_src = measured_peaks()
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    if world == 1:
        w = churn.ChurnWorld(n)
        eager_ms, ops, n_act = churn.time_ticks(w, ticks)
        tick_ms, graph = churn.time_graph_ticks(w, ticks)  # the same 13 launches per tick, two ticks per graph launch
        del graph
        alg = churn.algorithmic_bytes(n, n_act, 0.05 * n_act)
        rows = {}
        for name in churn.OPS:
            gbs = alg[name] / (ops[name] * 1e-3) / 1e9
            rows[name] = {"ms": ops[name], "algorithmic_bytes": float(alg[name]), "gbs": gbs, "frac_of_hbm": gbs / hbm}
        tot = float(sum(alg.values()))
        out = {"workload": f"C5: {n} Dice slots (28 B rows), D20, ~{n_act:.0f} active, ~5% + 5% turnover per tick; tick = "
                           "step -> select -> remove -> sort(active_state, desc) -> select -> clip -> add "
                           "(hello_world.ipynb:104-144)",
               "n_gpus": 1, "ms_per_tick": tick_ms, "eager_ms_per_tick": eager_ms,
               "mode": "CUDA graph of two ticks, replayed (per-op times: eager, events around every op)",
               "slot_steps_per_sec": n / (tick_ms * 1e-3),
               "active_agent_steps_per_sec": n_act / (tick_ms * 1e-3), "ops": rows,
               "algorithmic_bytes_per_tick": tot, "tick_gbs": tot / (tick_ms * 1e-3) / 1e9,
               "tick_frac_of_hbm": tot / (tick_ms * 1e-3) / 1e9 / hbm, "hbm_peak_gbs": hbm, "peak_basis": peak_src,
               "launches_per_tick": 13, "final_active": int(churn.fgx_methods.count_active(w.agents))}
        del w
        torch.cuda.empty_cache()
        return out
    from foragax_b200.examples import churn_sharded
    w = churn_sharded.ShardedChurnWorld(n)
    res = w.time_ticks(ticks)
    t = torch.tensor([res["ms_per_tick"], res["eager_ms_per_tick"]], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    cnt = torch.tensor([res["final_active_local"], res["rows_sent_remote_per_tick"]], dtype=torch.float64, device=dev)
    dist.all_reduce(cnt)
    out = {"workload": f"C5: {n} Dice slots (28 B rows) in contiguous blocks over {world} GPUs, D20; tick = local step / "
                       "select / remove -> GLOBAL pack of the actives (one kernel: compaction + peer-memory stores over "
                       "NVLink, device-side flag barriers, no NCCL call in the tick) -> local select -> global add range "
                       "-> local add",
           "n_gpus": world, "ms_per_tick": float(t[0]), "eager_ms_per_tick": float(t[1]),
           "slot_steps_per_sec": n / (float(t[0]) * 1e-3), "final_active": int(cnt[0]),
           "nvlink_row_bytes_per_tick": float(cnt[1]) * churn.ROW_BYTES, "mode": res["mode"]}
    del w
    torch.cuda.empty_cache()
    return out


def d2h_ceiling(n_local, dev, barrier, max_over_ranks):
    """Pinned device->host copy bandwidth with every rank copying at once (what bounds the e2e leg): one 35 MB-class
    buffer per rank, 10 copies, GB/s summed over ranks."""
    import torch
    nbytes = (2 * 4 * N_RAYS + 8) * n_local
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    dst = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dst.copy_(src, non_blocking=True)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / 10
    return nbytes, ms


def run_ours(args):
    import hashlib

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
    from foragax_b200 import distributed as fd
    from foragax_b200.base.agent_classes import Params, Signal
    from foragax_b200.base.agent_methods import step_agents
    from foragax_b200.examples.foraging import Forager

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

    def tick_two():
        k1()
        k2()

    def tick_fused():
        # the same tick as ONE kernel: fgx_forager_sense_step -- four ray-casting warps beside the eight step warps of
        # every CTA; same bits as k1 + k2 (parity block below, tests/test_gpu_foraging.py)
        Forager.sense_step(step_params, energy_in, fset.agents, space.walls, BOUNDS, RAY_SPAN, RAY_LEN,
                           out=(dist_out, hit_out))

    fused = args.fused_tick
    tick = tick_fused if fused else tick_two
    launches_per_step = 1 if fused else 2

    peer_positions = None
    if world > 1 and not args.no_agent_sensing and not args.nccl_positions:
        from foragax_b200.peer import PeerPositions
        peer_positions = PeerPositions(n_local, layout="round_robin", device=dev)

    def sense():
        # agent-agent ray sensing: the float4 positions of all agents IN SLOT ORDER (round-robin shards are
        # un-interleaved) -- through peer memory (stage, mailbox barrier, pull over NVLink; `--nccl-positions`: NCCL
        # all-gather) --, uniform-grid build, ray_agent_sensor, merge with the wall result; hit = global entity id
        k1()
        if peer_positions is not None:
            tg = peer_positions.gather(st["X"], st["Y"], st["rad"], F.active_state)
        else:
            tg = fd.all_gather_positions(st["X"], st["Y"], st["rad"], F.active_state, layout="round_robin")
        sm.raycast_agents((st["X"], st["Y"], st["ang"]), RAY_SPAN, RAY_LEN, tg, BOUNDS, 15.0,
                          active_state=F.active_state, n_rays=N_RAYS, merge_walls=True, out=(dist_out, hit_out))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def slot_order(t):
        """The local rows of every rank reassembled in GLOBAL slot order (slot i = row i // world of rank i % world)."""
        t = t.contiguous()
        if world == 1:
            return t
        assert N_AGENTS % world == 0
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=dev)
        dist.all_gather_into_tensor(out, t)
        return out.transpose(0, 1).reshape((N_AGENTS,) + tuple(t.shape[1:]))

    def digest(t):
        g = slot_order(t)
        h = hashlib.sha256(g.cpu().numpy().tobytes()).hexdigest() if rank == 0 else None
        del g
        return h

    # ---- parity block: PARITY_TICKS ticks from the fixed seed, BEFORE any timing.  Digests of the outputs in global
    # slot order are independent of N (every agent is stepped by the same kernel on the same operands), so the
    # N-GPU line must carry the digests committed from the 1-GPU run; rank 0 also checks a slice against the oracle.
    parity = None
    if not args.no_parity:
        oracle_slice = None
        ticks_left = PARITY_TICKS
        if rank == 0:
            oracle_slice = oracle_slice_check(work, F, space, dist_out, hit_out, tick, args.parity_rows, dev)
        else:
            tick()
        ticks_left -= 1
        for _ in range(ticks_left):
            tick()
        dg = {"dist": digest(dist_out), "hit": digest(hit_out), "X": digest(st["X"]),
              "Z": digest(F.policy.state.content["Z"])}
        if not args.no_agent_sensing:
            sense()
            dg["agent_sensing_dist"] = digest(dist_out)
            dg["agent_sensing_hit"] = digest(hit_out)
        other_dg = None
        other_name = "two_kernel" if fused else "fused"
        if not args.no_other_tick:
            # the same PARITY_TICKS ticks through the OTHER form of the tick (two kernels: fgx_raycast_walls_grid +
            # fgx_forager_step / one kernel: fgx_forager_sense_step) on a second set built from the same seed: its
            # digests must equal the same golden
            fset2, _ = build_foragers(work, rank, world, dev)
            d2, h2 = torch.empty_like(dist_out), torch.empty_like(hit_out)
            s2 = fset2.agents.state.content
            for _ in range(PARITY_TICKS):
                if fused:
                    sm.raycast_walls((s2["X"], s2["Y"], s2["ang"]), RAY_SPAN, RAY_LEN, space.walls,
                                     active_state=fset2.agents.active_state, n_rays=N_RAYS, out=(d2, h2), bounds=BOUNDS)
                    fset2.agents = step_agents(step_params, Signal(content={'obs': d2, 'energy_in': energy_in}), fset2)
                else:
                    Forager.sense_step(step_params, energy_in, fset2.agents, space.walls, BOUNDS, RAY_SPAN, RAY_LEN,
                                       out=(d2, h2))
            other_dg = {"dist": digest(d2), "hit": digest(h2), "X": digest(s2["X"]),
                        "Z": digest(fset2.agents.policy.state.content["Z"])}
            del fset2, d2, h2, s2
            torch.cuda.empty_cache()
        if rank == 0:
            golden = json.load(open(PARITY_GOLDEN)) if os.path.exists(PARITY_GOLDEN) else None
            same = None
            if golden:
                same = {k: (golden["sha256"].get(k) == v) for k, v in dg.items()}
                if other_dg:
                    same.update({other_name + "_" + k: (golden["sha256"].get(k) == v) for k, v in other_dg.items()})
            parity = {"ticks": PARITY_TICKS, "order": "global slot order (round-robin shards un-interleaved)",
                      "sha256": dg, "golden": "tests/golden/bench_parity_digest.json (1-GPU run)" if golden else None,
                      "equals_1gpu_golden": same, "all_equal": (all(same.values()) if same else None),
                      "oracle_slice": oracle_slice}
            if args.write_parity_golden:
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                with open(os.path.join(ROOT, "gpurun_out", "bench_parity_digest.json"), "w") as f:
                    json.dump({"sha256": dg, "ticks": PARITY_TICKS, "n_gpus": world,
                               "made_by": "python bench.py --write-parity-golden (see DESIGN.md, Measurement)"}, f,
                              indent=1)

    stream = torch.cuda.current_stream()
    sampler = ClockSampler(local) if rank == 0 else None
    warm = max(args.warmup, 3)
    for _ in range(warm):
        tick()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed_ticks(one_kernel):
        """K ticks between e0 and e1 (barrier + synchronize on both sides; max over ranks).  The two-kernel form also
        records an event between and after its kernels INSIDE the timed loop, without any host sync: the per-kernel
        durations then come from the very launches the step time comes from (a separate later loop can run at a
        different power / clock state -- one round-2 line had K1c + K2 above its own step time that way -- and a
        sync per iteration would put Python's launch latency into the short kernel's interval, which is how round 1
        reported K1c ~45 us too long)."""
        marks = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
        barrier()
        e0.record(stream)
        for i in range(args.steps):
            if one_kernel:
                tick_fused()
            else:
                k1()
                marks[i][0].record(stream)
                k2()
                marks[i][1].record(stream)
        e1.record(stream)
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / args.steps
        if one_kernel:
            return ms, None, None
        a = [(e0 if i == 0 else marks[i - 1][1]).elapsed_time(marks[i][0]) for i in range(args.steps)]
        b = [marks[i][0].elapsed_time(marks[i][1]) for i in range(args.steps)]
        return ms, float(np.mean(a)), float(np.mean(b))

    ms_step, k1_ms, k2_ms = timed_ticks(fused)
    tests_per_step = float(N_AGENTS) * N_RAYS * N_WALLS
    value = tests_per_step / (ms_step * 1e-3)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)]]
    ev = evs[0]
    listed = sm.wall_grid_listed((st["X"], st["Y"]), space.walls, BOUNDS, RAY_LEN, active_state=F.active_state)

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

    # the other form of the tick, timed the same way as the default one
    other_ms = None
    if not args.no_other_tick:
        other_tick = tick_two if fused else tick_fused
        for _ in range(3):
            other_tick()
        other_ms, o1, o2 = timed_ticks(not fused)
        if fused:
            k1_ms, k2_ms = o1, o2
    if k1_ms is None:  # --fused-tick --no-other-tick: the two kernels are still needed for the stage blocks
        for _ in range(2):
            tick_two()
        _, k1_ms, k2_ms = timed_ticks(False)

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
        if fused:
            Forager.sense_step(step_params, energy_in, fset.agents, space.walls, BOUNDS, RAY_SPAN, RAY_LEN, out=(d_o, h_o))
        else:
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
    # the ceiling of that leg: every rank's pinned read-back running at once, nothing else on the GPUs
    ceil_bytes, ceil_ms = d2h_ceiling(n_local, dev, barrier, max_over_ranks)

    # extra stage (not part of `value`): agent-agent ray sensing.  Timed eagerly and -- every launch of the leg being a
    # kernel on the current stream (no NCCL with peer-memory positions) -- replayed as ONE CUDA graph, which removes the
    # ~15 Python launches whose CPU latency is exposed once a rank holds 1/8 of the agents
    agent_sensing = None
    if not args.no_agent_sensing:
        sense()
        barrier()
        reps = max(2, min(5, args.steps))
        e0.record(stream)
        for _ in range(reps):
            sense()
        e1.record(stream)
        barrier()
        sense_ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        graph_ms = None
        if peer_positions is not None or world == 1:
            graph = torch.cuda.CUDAGraph()
            captured = 1
            try:
                with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                    sense()
            except RuntimeError:
                captured = 0
            ok = torch.tensor([captured], device=dev)
            if world > 1:
                dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok):
                graph.replay()
                barrier()
                e0.record(stream)
                for _ in range(reps):
                    graph.replay()
                e1.record(stream)
                barrier()
                graph_ms = max_over_ranks(e0.elapsed_time(e1)) / reps
        if peer_positions is not None:
            peer_positions.check()
        frac_agent = float(((hit_out >= 0) & (hit_out < N_AGENTS)).float().mean())
        best = sense_ms if graph_ms is None else min(sense_ms, graph_ms)
        agent_sensing = {"ms_per_tick_walls_plus_agents": best, "eager_ms": sense_ms, "graph_ms": graph_ms,
                         "ms_agents_only": best - k1_ms,
                         "rays_hitting_an_agent": frac_agent,
                         "positions_via": ("peer memory (fgx_peer_gather_positions, no NCCL)" if peer_positions is not None
                                           else ("NCCL all-gather" if world > 1 else "local")),
                         "note": "fgx_raycast_walls + positions of all agents in slot order + fgx_raycast_agents (grid, exact "
                                 "vs all-pairs); 1M agents sensing 1M targets, not included in `value`; hit indices "
                                 "are global slot ids (digest in `parity`)"}

    # free the 12.5 GB forager set before the other configs run
    listed_host = [int(v) for v in listed.cpu().tolist()]
    c5 = None
    if not args.no_c5:
        del fset, F, st, outs, step_input
        torch.cuda.empty_cache()
        c5 = c5_churn_leg(args, rank, world, dev)

    if rank == 0:
        peaks, peak_src = measured_peaks()
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        peak_tflops = sms * 128 * sm_max * 1e6 / 1e12  # non-FMA FP32 lanes x clock
        # all-pairs ray cast: FP32 / issue bound, 20 algorithmic FLOP per (ray, wall) test
        ap_tests = float(n_local) * N_RAYS * N_WALLS
        ach = FLOP_PER_TEST * ap_tests / (ap_k1_ms * 1e-3) / 1e12
        ach_exec = 16 * ap_tests / (ap_k1_ms * 1e-3) / 1e12
        roof_ap = {"bound": "fp32", "achieved": ach, "peak": peak_tflops, "unit": "TFLOP/s", "frac": ach / peak_tflops,
                   "executed": {"flop_per_test": 16, "achieved": ach_exec, "frac": ach_exec / peak_tflops,
                                "note": "o4 reuses the two differences of o2 and o3 is hoisted per (thread, wall): 16 "
                                        "FP32 ops execute per test against the 20 of SURVEY.md 8d"},
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
        roof_k2 = {"bound": "hbm", "achieved": k2_gbs, "peak": hbm, "unit": "GB/s", "frac": k2_gbs / hbm,
                   "traffic": measured_traffic("forager_step_kernel", n_local), "algorithmic_bytes": float(k2_bytes),
                   "kernel": "forager_step_kernel<40,33,2>", "kernel_ms": k2_ms,
                   "share_of_step": k2_ms / (k1_ms + k2_ms),
                   "frac_of_nominal_8tbs": k2_gbs / 8000.0,
                   "bytes_per_active_agent": K2_BYTES_ACTIVE, "peak_basis": f"hbm_gbs, {peak_src} MEASURED_PEAKS.json"}
        if fused:
            # one launch per step: the kernel's duration is the step's; its algorithmic bytes are the step kernel's
            # (weights, policy state, scalars) plus the ray stage's (pose in, dist + hit out)
            fk_bytes = k2_bytes + float(n_local) * 16 + float(n_local) * N_RAYS * 8
            fk_gbs = fk_bytes / (ms_step * 1e-3) / 1e9
            roof = {"bound": "hbm", "achieved": fk_gbs, "peak": hbm, "unit": "GB/s", "frac": fk_gbs / hbm,
                    "traffic": measured_traffic("forager_sense_step_kernel", n_local),
                    "algorithmic_bytes": float(fk_bytes),
                    "kernel": "forager_step_kernel<40,33,2,RW=4> (fgx_forager_sense_step: 8 step warps + 4 ray-casting "
                              "warps per CTA, 2 CTAs per SM)",
                    "kernel_ms": ms_step, "share_of_step": 1.0, "frac_of_nominal_8tbs": fk_gbs / 8000.0,
                    "bytes_per_active_agent": K2_BYTES_ACTIVE, "bytes_per_agent_ray_stage": 16 + 8 * N_RAYS,
                    "peak_basis": f"hbm_gbs, {peak_src} MEASURED_PEAKS.json",
                    "step_kernel_alone": {"kernel": "forager_step_kernel<40,33,2>", "kernel_ms": k2_ms,
                                          "achieved": k2_gbs, "frac": k2_gbs / hbm,
                                          "note": "the HBM-bound half of the two-kernel tick, timed in the same run "
                                                  "(two_kernel_tick); the fused kernel hides the ray cast in its idle "
                                                  "issue slots at the price of some of this fraction"}}
        else:
            roof = roof_k2
        # the culled ray cast against its two floors: bytes (pose in, dist + hit out) and instructions (the walls the
        # grid lists for the agents' cells x 16 executed FP32 ops per (ray, wall) + the output), not against its own
        # instruction count
        rc_bytes = float(n_local) * 16 + float(n_local) * N_RAYS * 8
        listed_pairs = float(listed_host[0]) * N_RAYS            # (ray, wall) tests the kernel has to evaluate
        floor_bytes_ms = rc_bytes / (hbm * 1e9) * 1e3
        floor_inst_ms = listed_pairs * 16 / (peak_tflops * 1e12) * 1e3
        roof_rc = {"kernel": "raycast_walls_grid_warp32_kernel (fgx_raycast_walls_grid, 32 rays: a warp per agent; wall lists built "
                             "once, cached)", "kernel_ms": k1_ms,
                   "share_of_step": k1_ms / (k1_ms + k2_ms),
                   "algorithmic_bytes": rc_bytes, "gbs": rc_bytes / (k1_ms * 1e-3) / 1e9,
                   "traffic": measured_traffic("raycast_walls_grid_kernel", n_local),
                   "frac_of_hbm": rc_bytes / (k1_ms * 1e-3) / 1e9 / hbm,
                   "listed_walls_per_active_agent": listed_host[0] / max(listed_host[1], 1),
                   "floors_ms": {"bytes": floor_bytes_ms, "fp32_of_listed_tests": floor_inst_ms},
                   "frac_of_larger_floor": max(floor_bytes_ms, floor_inst_ms) / k1_ms,
                   "executed_tests_per_sec": listed_pairs / (k1_ms * 1e-3),
                   "algorithmic_tests_per_sec": float(n_local) * N_RAYS * N_WALLS / (k1_ms * 1e-3),
                   "speedup_vs_all_pairs_kernel": ap_k1_ms / k1_ms,
                   "note": "spatial cull: per grid cell only the walls that can matter are tested; bit-identical to the "
                           "all-pairs kernel. Floors: algorithmic bytes / measured HBM, and listed (ray, wall) tests x "
                           "16 FP32 ops / non-FMA FP32 peak"}
        ceil_gbs = ceil_bytes * world / (ceil_ms * 1e-3) / 1e9
        e2e_floor_ms = (h2d + d2h) / world / (ceil_bytes / (ceil_ms * 1e-3)) * 1e3
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "agents": N_AGENTS, "active_agents": int(0.9 * N_AGENTS), "rays": N_RAYS,
                       "walls": N_WALLS, "sharding": f"agents by slot index, round-robin over {world} rank(s), walls replicated",
                       "tick": (["fgx_forager_sense_step = fgx_raycast_walls_grid (ray_generator + ray_wall_sensor + min/argmin "
                                 "over the walls a uniform grid lists for the agent's cell; bit-identical to all-pairs) "
                                 "and fgx_forager_step (WilsonCowan policy + physics) as ONE kernel; the two-kernel "
                                 "form is timed in `two_kernel_tick`"] if fused else
                                ["fgx_raycast_walls_grid (ray_generator + ray_wall_sensor + min/argmin over the walls a "
                                 "uniform grid lists for the agent's cell; results bit-identical to all-pairs)",
                                 "fgx_forager_step (WilsonCowan policy + physics)"]),
                       "value_counts": "agents x rays x walls per tick = algorithmic pairs (SURVEY.md 8d), all slots, "
                                       "whether or not the kernel culls them; see executed_tests_per_sec and all_pairs",
                       "l2": "per-step working set 11.6 GB of per-agent weights + 272 MB ray I/O >> 126 MB L2"},
            "agent_steps_per_sec": N_AGENTS / (ms_step * 1e-3),
            "executed_tests_per_sec": listed_pairs * world / (ms_step * 1e-3),
            "roofline": roof, "raycast_stage": roof_rc, "all_pairs": all_pairs,
            "e2e": {"value": tests_per_step / (e2e_ms_step * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms_step,
                    "agent_steps_per_sec": N_AGENTS / (e2e_ms_step * 1e-3),
                    "copy_ceiling": {"pinned_d2h_gbs_all_ranks": ceil_gbs, "bytes_per_rank": ceil_bytes,
                                     "ms_for_this_steps_bytes": e2e_floor_ms,
                                     "e2e_ms_over_ceiling_ms": e2e_ms_step / e2e_floor_ms,
                                     "note": "measured here: every rank copies its result-sized buffer device -> "
                                             "pinned host at once, nothing else running; (H2D + D2H bytes of a step) "
                                             "/ that rate is the floor of the e2e leg"}},
            "gpu_launches": launches_per_step * args.steps,
            "clocks": clocks,
        }
        if other_ms is not None:
            line["two_kernel_tick" if fused else "fused_tick"] = {
                "ms_per_step": other_ms, "value": tests_per_step / (other_ms * 1e-3), "unit": UNIT,
                "agent_steps_per_sec": N_AGENTS / (other_ms * 1e-3), "default_tick_speedup": other_ms / ms_step,
                "roofline": roof_k2 if fused else None,
                "note": ("fgx_raycast_walls_grid + fgx_forager_step as two launches; same bits (parity.equals_1gpu_golden "
                         "two_kernel_*)" if fused else
                         "fgx_forager_sense_step: the same tick as ONE launch -- 4 ray-casting warps beside the 8 step "
                         "warps of every CTA (forager_step_kernel<40,33,2,RW=4>); same bits (parity.equals_1gpu_golden "
                         "fused_*)")}
            tick_bytes = float(k2_bytes) + rc_bytes
            line["tick_frac_of_hbm"] = {
                "algorithmic_bytes_per_tick": tick_bytes,
                "two_kernel": tick_bytes / ((other_ms if fused else ms_step) * 1e-3) / 1e9 / hbm,
                "fused": tick_bytes / ((ms_step if fused else other_ms) * 1e-3) / 1e9 / hbm,
                "note": "whole tick (step kernel's bytes + pose in / dist + hit out of the ray stage) against measured HBM"}
        if parity:
            line["parity"] = parity
        if agent_sensing:
            line["agent_sensing"] = agent_sensing
        if c5:
            line["c5_churn"] = c5
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(work, args.cpu_sample)
            cb = line["cpu_baseline"]
            line["all_pairs_vs_cpu"] = {"gpu_all_pairs_tests_per_sec": all_pairs["value"],
                                        "cpu_tests_per_sec": cb["value"], "ratio": all_pairs["value"] / cb["value"],
                                        "note": "like for like: neither side culls (the CPU port is all-pairs)"}
        if world == 1 and not args.no_extra:
            line["other_configs"] = other_configs()
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def other_configs():
    """BASELINE configs 2 and 3 (scripts/bench_foraging.py, scripts/bench_wolf_sheep.py), each in its own process on
    the same GPU after the main measurement; their last stdout line is a JSON record."""
    out = {}
    for name, script in (("c2_foraging_world", "bench_foraging.py"), ("c3_wolf_sheep_grass", "bench_wolf_sheep.py")):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", script)], capture_output=True, text=True,
                               timeout=240, cwd=ROOT)
            last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            out[name] = json.loads(last[-1]) if r.returncode == 0 and last else {"error": r.stderr[-300:]}
        except (subprocess.TimeoutExpired, ValueError) as e:
            out[name] = {"error": str(e)[:300]}
    return out


_STDOUT = sys.stdout


def emit(line):
    """The ONE JSON line of the contract, on the real stdout (everything else -- e.g. the reference API's own
    "AgentSet initialized" print, agent_classes.py:83 -- is sent to stderr while the bench runs)."""
    _STDOUT.write(json.dumps(line) + "\n")
    _STDOUT.flush()


def main():
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=40_000, help="agents in the cpu_baseline sample")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--fused-tick", action="store_true",
                    help="time the one-kernel fgx_forager_sense_step as the tick (default: fgx_raycast_walls_grid + "
                         "fgx_forager_step, two launches; the other form is always timed as a leg)")
    ap.add_argument("--no-other-tick", action="store_true", help="skip the leg / parity digests of the other tick form")
    ap.add_argument("--no-agent-sensing", action="store_true", help="skip the extra agent-agent sensing stage")
    ap.add_argument("--nccl-positions", action="store_true",
                    help="agent sensing at N > 1: exchange positions with an NCCL all-gather instead of peer memory")
    ap.add_argument("--no-parity", action="store_true", help="skip the parity block (digests + oracle slice)")
    ap.add_argument("--parity-rows", type=int, default=2000, help="agents of the oracle slice of the parity block")
    ap.add_argument("--write-parity-golden", action="store_true",
                    help="also write gpurun_out/bench_parity_digest.json (copy it to tests/golden/ from a 1-GPU run)")
    ap.add_argument("--no-c5", action="store_true", help="skip the config-5 churn leg")
    ap.add_argument("--c5-slots", type=int, default=10_000_000)
    ap.add_argument("--c5-ticks", type=int, default=20)
    ap.add_argument("--no-extra", action="store_true", help="skip the config-2 / config-3 subprocess legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
