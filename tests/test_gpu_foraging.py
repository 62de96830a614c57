"""K2 parity: the foraging example (Forager / Resource / interaction / sensor_model /
Neuroevolution) on the CUDA path vs the oracle and vs the reference's golden vector G2.

Tolerances (north star): RNG draws, keys, active masks, selected indices, sort order and counts
bit-exact; fp32 state within 1e-5 relative (the policy's dot products are summed in a different
order and tanh/exp/atan2 differ by a few ulp between CUDA and NumPy)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import agent_methods as oam
from oracle import foraging as ofo
from oracle import jax_random as jr

pytestmark = pytest.mark.gpu

G2 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "g2_foraging.json")))["ticks"]
RTOL, ATOL = 1e-5, 2e-5


def np_forager(a):
    c = a.state.content
    pc = a.policy.params.content
    return {"state": {k: c[k].cpu().numpy() for k in ofo.FSTATE},
            "policy": {**{k: pc[k].cpu().numpy() for k in ("J", "tau", "E", "B", "D")},
                       "Z": a.policy.state.content["Z"].cpu().numpy()},
            "unique_id": a.unique_id.cpu().numpy(), "agent_type": a.agent_type.cpu().numpy(),
            "active_state": a.active_state.cpu().numpy(), "age": a.age.cpu().numpy()}


def np_resource(a):
    c = a.state.content
    return {"state": {k: c[k].cpu().numpy() for k in ("X", "Y", "energy", "rad", "blossom", "key", "time_above_blossom")},
            "params": {k: a.params.content[k].cpu().numpy() for k in ("growth_rate", "decay_rate")},
            "unique_id": a.unique_id.cpu().numpy(), "agent_type": a.agent_type.cpu().numpy(),
            "active_state": a.active_state.cpu().numpy(), "age": a.age.cpu().numpy()}


def assert_tree(got, ref, exact=(), path="", rel=RTOL):
    """Integer / bool / key leaves (and those named in `exact`) bit-exact; float leaves within
    `rel` relative to the leaf's scale (the policy state Z reaches O(100), so an absolute floor
    of rel * max|leaf| is the meaningful form of "1e-5 relative" for sums of 68 products)."""
    for k, v in ref.items():
        if isinstance(v, dict):
            assert_tree(got[k], v, exact, path + k + ".", rel)
        elif v.dtype.kind in "iub" or (path + k) in exact:
            np.testing.assert_array_equal(got[k], v, err_msg=path + k)
        else:
            scale = max(1.0, float(np.abs(v[np.isfinite(v)]).max()) if v.size else 1.0)
            np.testing.assert_allclose(got[k], v, rtol=rel, atol=rel * scale, err_msg=path + k)


def make_world(variant, cuda):
    from foragax_b200 import random as fgx_random
    from foragax_b200.examples import foraging as fg
    fg.use_variant(variant)
    cfg = ofo.Config(variant)
    w = fg.Foraging_world(fg.default_params(variant), fgx_random.PRNGKey(cfg.seed), 10)
    return fg, w, cfg


@pytest.mark.parametrize("variant", ["sim", "nb"])
def test_create_matches_oracle(cuda, variant):
    fg, w, cfg = make_world(variant, cuda)
    ow = ofo.World(cfg)
    exact = {"state." + k for k in ofo.FSTATE if k != "ang"} | {"policy." + k for k in ("J", "tau", "E", "B", "D", "Z")}
    assert_tree(np_forager(w.foragers_set.agents), ow.fg, exact)
    assert_tree(np_resource(w.resources_set.agents), ow.rs,
                {"state.X", "state.Y", "state.energy", "state.rad", "state.time_above_blossom", "params.growth_rate",
                 "params.decay_rate"})
    np.testing.assert_array_equal(w.key.cpu().numpy(), ow.key)
    fg.use_variant("sim")


def test_tick_stages_match_oracle(cuda):
    """interaction, sensor_model, forager step, resource step -- each stage against the oracle
    on the oracle's own inputs (so that errors do not compound)."""
    from foragax_b200.base.agent_classes import Params, Signal
    fg, w, cfg = make_world("sim", cuda)
    ow = ofo.World(cfg)
    for _ in range(3):  # advance both worlds so that positions / Z are non-trivial
        w.step()
        ow.step()
    F, Rs = w.foragers_set.agents, w.resources_set.agents
    assert_tree(np_forager(F), ow.fg, rel=1e-4)  # three compounded ticks
    # interaction
    _, e_in, e_out = fg.jit_forager_resource_interaction(F, Rs)
    r_in, r_out = ofo.interaction(np_forager(F), np_resource(Rs))
    np.testing.assert_allclose(e_in.cpu().numpy(), r_in, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(e_out.cpu().numpy(), r_out, rtol=1e-6, atol=1e-6)
    assert (r_in > 0).sum() > 5
    # sensor model (device sincos within 2 ulp of the oracle's: distances to 1e-5, rare grazing flips)
    obs = fg.jit_sensor_model(F, Rs, None, w.space.walls).cpu().numpy()
    robs = ofo.sensor_model(cfg, np_forager(F), np_resource(Rs), cfg.walls)
    same = np.isclose(obs, robs, rtol=RTOL, atol=ATOL * 60)
    assert same.mean() > 0.9995, f"sensor mismatch fraction {1 - same.mean():.2e}"
    assert (robs.reshape(len(robs), -1, 3)[:, :, 1] > 0).sum() > 50  # rays do hit agents
    # forager step on the oracle's observations
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(cuda)  # noqa: E731
    ref = ofo.forager_step(cfg, robs, r_in, np_forager(F))
    sp = Params(content={'space': w.space, 'repr_thresh': 5.0, 'die_thresh': 0.0, 'energy_min': -1.0,
                         'energy_max': 15.0})
    out = fg.Forager.step_agent(sp, Signal(content={'obs': d(robs), 'energy_in': d(r_in)}), F)
    assert_tree(np_forager(out), ref)
    # resource step
    ref_r = ofo.resource_step(cfg, r_out, np_resource(Rs))
    out_r = fg.Resource.step_agent(Params(content={'blossom_thresh': 3.0}), Signal(content={'energy_out': d(r_out)}), Rs)
    got_r = np_resource(out_r)
    assert_tree(got_r, ref_r, {"state.energy", "state.rad", "state.time_above_blossom"})


def test_wilson_cowan_step_policy(cuda):
    from foragax_b200 import random as fgx_random
    from foragax_b200.base.agent_classes import Params, Signal
    from foragax_b200.policy.wilson_cowan import WilsonCowan
    # (40, 33, 2) is the instantiation bench.py times (odd-length E rows: the `len & 1` tail of row_dot2<33>);
    # (22, 18, 2) is even but not a multiple of 4 (generic path, float2 rows without a float4 body)
    for nn, nobs, nact in ((40, 28, 2), (30, 28, 2), (40, 33, 2), (22, 18, 2), (17, 5, 3), (64, 97, 4)):
        cfg = ofo.Config("sim", n_neurons=nn, n_obs=nobs, n_act=nact)
        keys = jr.split(jr.PRNGKey(nn), 300)
        dk = torch.from_numpy(keys).to(cuda)
        pol = WilsonCowan.create_policy(Params(content={'num_neurons': nn, 'num_obs': nobs, 'num_actions': nact,
                                                        'deterministic': True, 'dt': 0.1, 'action_scale': 1.0}), dk)
        ref = ofo.wc_create(cfg, keys)
        for k in ("J", "tau", "E", "B", "D"):
            np.testing.assert_array_equal(pol.params.content[k].cpu().numpy(), ref[k], err_msg=k)
        rng = np.random.default_rng(0)
        for _ in range(3):
            obs = rng.standard_normal((300, nobs)).astype(np.float32)
            act, ref_Z = ofo.wc_step(cfg, ref, obs)
            ref["Z"] = ref_Z
            sig, pol = WilsonCowan.step_policy(Signal(content={'obs': torch.from_numpy(obs).to(cuda)}), pol)
            scale = max(1.0, float(np.abs(ref_Z).max()))
            np.testing.assert_allclose(pol.state.content['Z'].cpu().numpy(), ref_Z, rtol=RTOL, atol=RTOL * scale)
            np.testing.assert_allclose(sig.content['actions'].cpu().numpy(), act, rtol=1e-4, atol=1e-4)
            # keep the oracle on the device's state so that the per-step comparison does not compound
            ref["Z"] = pol.state.content['Z'].cpu().numpy().copy()


def test_world_ticks_match_oracle_sim_variant(cuda):
    """Whole ticks (interaction -> sensor -> steps -> Neuroevolution) against the oracle world:
    counts, masks, ids and keys exact; float state within tolerance."""
    fg, w, cfg = make_world("sim", cuda)
    ow = ofo.World(cfg)
    for t in range(12):
        nf, nr, le = w.step()
        onf, onr, ole = ow.step()
        assert (int(nf), int(nr)) == (onf, onr), f"tick {t}"
        np.testing.assert_allclose(float(le), ole, rtol=1e-4)
        gf, gr = np_forager(w.foragers_set.agents), np_resource(w.resources_set.agents)
        np.testing.assert_array_equal(gf["active_state"], ow.fg["active_state"])
        np.testing.assert_array_equal(gf["unique_id"], ow.fg["unique_id"])
        np.testing.assert_array_equal(gr["active_state"], ow.rs["active_state"])
        np.testing.assert_array_equal(gr["unique_id"], ow.rs["unique_id"])
        np.testing.assert_array_equal(gr["state"]["key"], ow.rs["state"]["key"])
        np.testing.assert_array_equal(w.key.cpu().numpy(), ow.key)
    assert_tree(gf, ow.fg, rel=1e-3)  # twelve compounded ticks of an expanding RNN
    assert_tree(gr, ow.rs, rel=1e-4)


def test_g2_golden_on_the_cuda_path(cuda):
    """The reference notebook's printed run (golden vector G2, PRNGKey(0)) reproduced by the CUDA
    path: (num_active_foragers, num_active_resources, life_expectancy) for ticks 0..40."""
    fg, w, cfg = make_world("nb", cuda)
    try:
        for t in range(41):
            nf, nr, le = w.step()
            assert (int(nf), int(nr)) == (G2[t][0], G2[t][1]), f"tick {t}: {(int(nf), int(nr))} != {G2[t][:2]}"
            assert abs(float(le) - G2[t][2]) <= 1e-4 * max(1.0, abs(G2[t][2])), f"tick {t}"
    finally:
        fg.use_variant("sim")


def test_cuda_graph_replay_equals_eager_ticks(cuda):
    """A whole tick captured into a CUDA graph and replayed (one launch per tick) walks the same
    trajectory as the eager calls -- same kernels, same order -- and still reproduces G2."""
    fg, w_eager, cfg = make_world("nb", cuda)
    _, w_graph, _ = make_world("nb", cuda)
    try:
        g = w_graph.capture(warmup=2)          # ticks 0 and 1 run eagerly inside capture()
        for t in range(2):
            w_eager.step()
        for t in range(2, 30):
            nf, nr, le = g.replay()
            enf, enr, ele = w_eager.step()
            assert (int(nf), int(nr)) == (int(enf), int(enr)) == (G2[t][0], G2[t][1]), f"tick {t}"
            assert float(le) == float(ele)
        a, b = np_forager(w_graph.foragers_set.agents), np_forager(w_eager.foragers_set.agents)
        assert_tree(a, b, exact={"state." + k for k in ofo.FSTATE} | {"policy." + k for k in ("J", "tau", "E", "B", "D", "Z")})
        np.testing.assert_array_equal(w_graph.key.cpu().numpy(), w_eager.key.cpu().numpy())
    finally:
        fg.use_variant("sim")


@pytest.mark.parametrize("mode", ["specialised-2", "specialised-4", "same-warp"])
def test_fused_sense_step_equals_raycast_then_step(cuda, mode, monkeypatch):
    """fgx_forager_sense_step must give the same bits as fgx_raycast_walls_grid followed by fgx_forager_step:
    distances, hit indices and every state / policy leaf, over several ticks, with inactive agents and agents
    outside the bounds.  Modes: ray warps beside the step warps of every CTA (2 or 4 of them; rounds published
    through shared-memory flags), and the older variant where the warp that steps an agent casts its rays first."""
    monkeypatch.setenv("FGX_SENSE_STEP_MODE", "1" if mode == "same-warp" else "2")
    monkeypatch.setenv("FGX_SENSE_RW", "4" if mode.endswith("4") else "2")

    import bench
    from foragax_b200.base.agent_classes import Params, Signal
    from foragax_b200.base.agent_methods import step_agents
    import foragax_b200.base.space_methods as sm
    from foragax_b200 import struct
    from foragax_b200.examples.foraging import Forager
    n = 30_000
    work = bench.make_workload(n)
    fa, space = bench.build_foragers(work, 0, 1, cuda, n_agents=n)
    fb, _ = bench.build_foragers(work, 0, 1, cuda, n_agents=n)
    for f in (fa, fb):   # some agents outside the arena (all-walls loop) and on the boundary
        f.agents.state.content["X"][:50] -= 300.0
        f.agents.state.content["Y"][50:100] = 0.0
    bounds = (0.0, bench.ARENA, 0.0, bench.ARENA)
    params = Params(content={'space': space, 'repr_thresh': 5.0, 'die_thresh': 0.0, 'energy_min': -1.0, 'energy_max': 15.0})
    energy_in = torch.rand(n, device=cuda) * 0.1
    for _ in range(4):
        A, B = fa.agents, fb.agents
        sa = A.state.content
        d0, h0 = sm.raycast_walls((sa["X"], sa["Y"], sa["ang"]), bench.RAY_SPAN, bench.RAY_LEN, space.walls,
                                  active_state=A.active_state, n_rays=bench.N_RAYS, bounds=bounds)
        fa.agents = step_agents(params, Signal(content={'obs': d0, 'energy_in': energy_in}), fa)
        fb.agents, d1, h1 = Forager.sense_step(params, energy_in, B, space.walls, bounds, bench.RAY_SPAN, bench.RAY_LEN)
        assert torch.equal(d0, d1) and torch.equal(h0, h1)
        for x, y in zip(struct.tree_leaves(fa.agents), struct.tree_leaves(fb.agents)):
            assert torch.equal(x, y)
    assert float((h0 >= 0).float().mean()) > 0.01


@pytest.mark.parametrize("n", [1, 7, 8, 9, 63, 300])
def test_fused_sense_step_small_sets(cuda, n, monkeypatch):
    """The warp-specialised fused tick on tiny agent sets (fewer agents than one round of one CTA, partial last tile,
    ray warps without any round): same bits as the two-kernel tick, no hang."""
    import bench
    from foragax_b200.base.agent_classes import Params, Signal
    from foragax_b200.base.agent_methods import step_agents
    import foragax_b200.base.space_methods as sm
    from foragax_b200 import struct
    from foragax_b200.examples.foraging import Forager
    monkeypatch.setenv("FGX_SENSE_STEP_MODE", "2")
    work = bench.make_workload(n, 200)
    fa, space = bench.build_foragers(work, 0, 1, cuda, n_agents=n)
    fb, _ = bench.build_foragers(work, 0, 1, cuda, n_agents=n)
    bounds = (0.0, bench.ARENA, 0.0, bench.ARENA)
    params = Params(content={'space': space, 'repr_thresh': 5.0, 'die_thresh': 0.0, 'energy_min': -1.0, 'energy_max': 15.0})
    energy_in = torch.zeros(n, device=cuda)
    for _ in range(3):
        sa = fa.agents.state.content
        d0, h0 = sm.raycast_walls((sa["X"], sa["Y"], sa["ang"]), bench.RAY_SPAN, bench.RAY_LEN, space.walls,
                                  active_state=fa.agents.active_state, n_rays=bench.N_RAYS, bounds=bounds)
        fa.agents = step_agents(params, Signal(content={'obs': d0, 'energy_in': energy_in}), fa)
        fb.agents, d1, h1 = Forager.sense_step(params, energy_in, fb.agents, space.walls, bounds, bench.RAY_SPAN,
                                               bench.RAY_LEN)
        assert torch.equal(d0, d1) and torch.equal(h0, h1)
        for x, y in zip(struct.tree_leaves(fa.agents), struct.tree_leaves(fb.agents)):
            assert torch.equal(x, y)


def test_bench_instantiation_forager_step_matches_oracle(cuda):
    """The kernel instantiation bench.py times -- forager_step_kernel<40,33,2> on the C4 workload (32 ray
    distances + own energy = 33 observations) -- against oracle.foraging.forager_step (Forager.step_agent,
    sim.py:74-136, with WilsonCowan.step_policy, wilson_cowan.py:34-51) on 50,000 agents of the bench workload,
    three ticks: ray cast on the device, the same distances fed to both sides, the oracle restarted from the
    device's state every tick.

    Tolerances: positions, energy, timers, age and the policy state Z within 1e-5 relative to the leaf scale
    (north_star).  The ACTIONS (X_acc / Y_acc = tanh(D . tanh Z')) and what is linear in them (velocities) are
    ill-conditioned in float32 at this workload -- observations reach 60, so E . obs sums ~300-sized terms and a
    different summation order moves Z' by ~1e-4, which tanh(Z') near 0 passes on undamped -- so for them the bar is
    set by a float64 evaluation of the same formula: the device must be as close to it as the float32 oracle is
    (4x its worst error)."""
    import bench
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.agent_classes import Params, Signal
    from foragax_b200.base.agent_methods import step_agents
    n = 50_000
    work = bench.make_workload(n)
    fs, space = bench.build_foragers(work, 0, 1, cuda, n_agents=n)
    cfg = ofo.Config("sim", n_neurons=bench.N_NEURONS, n_obs=bench.N_OBS, n_act=bench.N_ACT, x_max=bench.ARENA,
                     y_max=bench.ARENA)
    params = Params(content={'space': space, 'repr_thresh': 5.0, 'die_thresh': 0.0, 'energy_min': -1.0,
                             'energy_max': 15.0})
    bounds = (0.0, bench.ARENA, 0.0, bench.ARENA)
    rng = np.random.default_rng(5)
    dt = np.float64(cfg.dt)
    for tick in range(3):
        A = fs.agents
        st = A.state.content
        dist, _ = sm.raycast_walls((st["X"], st["Y"], st["ang"]), bench.RAY_SPAN, bench.RAY_LEN, space.walls,
                                   active_state=A.active_state, n_rays=bench.N_RAYS, bounds=bounds)
        e_in = (rng.random(n) * 0.2).astype(np.float32)
        before = np_forager(A)
        d_np = dist.cpu().numpy()
        ref = ofo.forager_step(cfg, d_np, e_in, before)
        fs.agents = step_agents(params, Signal(content={'obs': dist, 'energy_in': torch.from_numpy(e_in).to(cuda)}), fs)
        got = np_forager(fs.agents)
        # float64 evaluation of step_policy on the same float32 inputs
        P = {k: v.astype(np.float64) for k, v in before["policy"].items()}
        obs64 = np.concatenate([d_np, before["state"]["energy"]], axis=1).astype(np.float64)
        dz = (np.einsum("nij,nj->ni", P["J"], np.tanh(P["Z"])) + np.einsum("nij,nj->ni", P["E"], obs64) + P["B"]
              - P["Z"]) * (10.0 / (1.0 + np.exp(-P["tau"])))
        z64 = P["Z"] + dt * dz
        act64 = cfg.action_scale * np.tanh(np.einsum("nij,nj->ni", P["D"], np.tanh(z64)))
        on = before["active_state"] != 0
        # Z: 1e-5 of the leaf scale, and no further from the float64 value than the float32 oracle is (x4)
        zscale = max(1.0, float(np.abs(ref["policy"]["Z"]).max()))
        np.testing.assert_allclose(got["policy"]["Z"], ref["policy"]["Z"], rtol=RTOL, atol=RTOL * zscale)
        err_ref = float(np.abs(np.stack([ref["state"]["X_acc"], ref["state"]["Y_acc"]], 1)[on, :, 0] - act64[on]).max())
        err_dev = float(np.abs(np.stack([got["state"]["X_acc"], got["state"]["Y_acc"]], 1)[on, :, 0] - act64[on]).max())
        assert err_dev <= 4.0 * err_ref + 1e-6, (tick, err_dev, err_ref)
        atol_v = float(dt) * (err_dev + err_ref) + 1e-6
        for k in ("X_vel", "Y_vel"):
            np.testing.assert_allclose(got["state"][k], ref["state"][k], rtol=RTOL, atol=atol_v, err_msg=k)
        # everything that does not hang on the ill-conditioned actions: the usual bar
        well = {"state": {k: ref["state"][k] for k in ("X", "Y", "energy", "rad", "time_above_rep_thresh",
                                                        "time_below_die_thresh")},
                "unique_id": ref["unique_id"], "agent_type": ref["agent_type"], "active_state": ref["active_state"],
                "age": ref["age"]}
        assert_tree(got, well, rel=max(RTOL, 0.3 * float(dt) * (err_dev + err_ref)))
        # inactive slots pass through untouched
        for k in ofo.FSTATE:
            np.testing.assert_array_equal(got["state"][k][~on], before["state"][k][~on], err_msg=k)
    assert float((dist < bench.RAY_LEN).float().mean()) > 0.01  # the observations are not all "no hit"


def test_recording_ring_and_checkpoint(cuda, tmp_path):
    """Foraging_world.run with the recording ring (sim.py:555-595,668-701): frame f of the ring holds the positions
    after tick rec_start_frame + f, the full ring is checkpointed once, and the ring write wraps like .at[frame]."""
    from foragax_b200 import random as fgx_random
    from foragax_b200.examples import foraging as fg
    from foragax_b200.recording import CheckpointManager, init_rec_data, record_data
    fg.use_variant("sim")
    frames, start = 4, 3
    w = fg.Foraging_world(fg.default_params("sim"), fgx_random.PRNGKey(1), 8, num_rec_frames=frames,
                          rec_dir=str(tmp_path / "rec"))
    w2 = fg.Foraging_world(fg.default_params("sim"), fgx_random.PRNGKey(1), 8)
    w.run(rec_start_frame=start)
    want = {k: [] for k in ("X", "Y", "ang", "rad")}
    blossom = []
    for t in range(1, 9):
        w2.step()
        if start <= t < start + frames:
            for k in want:
                want[k].append(w2.foragers_set.agents.state.content[k].clone())
            blossom.append(w2.resources_set.agents.state.content["blossom"].reshape(-1).float().clone())
    mgr = CheckpointManager(str(tmp_path / "rec"))
    assert mgr.all_steps() == [start + frames - 1]
    got = mgr.restore()
    for k, name in (("X", "forager_xs"), ("Y", "forager_ys"), ("ang", "forager_angs"), ("rad", "forager_rads")):
        assert torch.equal(got[f"rec_data.{name}"], torch.stack(want[k]).cpu()), name
    assert torch.equal(got["rec_data.resource_blossoms"], torch.stack(blossom).cpu())
    assert torch.equal(w.rec_data.forager_xs.cpu(), got["rec_data.forager_xs"])
    # frame indices wrap (device scalar index, as inside a captured tick)
    rec = init_rec_data(3, w.foragers_set, w.resources_set)
    record_data(w.foragers_set.agents, w.resources_set.agents, rec, torch.tensor(4, dtype=torch.int32, device=cuda))
    assert torch.equal(rec.forager_xs[1], w.foragers_set.agents.state.content["X"])
    assert float(rec.forager_xs[0].abs().sum()) == 0.0 and float(rec.forager_xs[2].abs().sum()) == 0.0


def test_state_checkpoint_resume_is_bit_exact(cuda, tmp_path):
    """save_state after 6 ticks, 6 more ticks; a fresh world that loads the checkpoint and runs 6 ticks must hold
    exactly the same leaves, key and counters (the reference has no resume)."""
    from foragax_b200 import random as fgx_random
    from foragax_b200 import struct
    from foragax_b200.examples import foraging as fg
    from foragax_b200.recording import CheckpointManager
    fg.use_variant("nb")
    mgr = CheckpointManager(str(tmp_path / "state"))
    a = fg.Foraging_world(fg.default_params("nb"), fgx_random.PRNGKey(0), 6)
    a.run()
    a.save_state(mgr)
    out_a = a.run()
    b = fg.Foraging_world(fg.default_params("nb"), fgx_random.PRNGKey(5), 6)   # different seed: everything is overwritten
    assert b.load_state(mgr) == 6 and b.sim_step == 6
    out_b = b.run()
    for x, y in zip(struct.tree_leaves(a.foragers_set.agents) + struct.tree_leaves(a.resources_set.agents),
                    struct.tree_leaves(b.foragers_set.agents) + struct.tree_leaves(b.resources_set.agents)):
        assert torch.equal(x, y)
    assert torch.equal(a.key, b.key) and torch.equal(a.life_expectancy, b.life_expectancy)
    assert [(int(f), int(r)) for f, r, _ in out_a] == [(int(f), int(r)) for f, r, _ in out_b]
    assert [(int(f), int(r)) for f, r, _ in out_a] == [tuple(G2[t][:2]) for t in range(6, 12)]  # and they are G2's ticks
