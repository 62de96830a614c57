"""K3 + Dice K2 parity through the foragax.base-shaped API (which calls the C ABI):
G1 end to end, multi-tick churn vs the oracle, select / sort edge cases -- all bit-exact."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import agent_methods as oam
from oracle import dice as odice
from oracle import jax_random as jr

pytestmark = pytest.mark.gpu

G1 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "g1_hello_world.json")))


def to_np(agents):
    c = agents.state.content
    return {"state": {"draw": c["draw"].cpu().numpy(), "key": c["key"].cpu().numpy()},
            "unique_id": agents.unique_id.cpu().numpy(), "agent_type": agents.agent_type.cpu().numpy(),
            "active_state": agents.active_state.cpu().numpy(), "age": agents.age.cpu().numpy()}


def assert_same(dev_agents, ref):
    got = to_np(dev_agents)
    for k in ("unique_id", "agent_type", "active_state", "age"):
        np.testing.assert_array_equal(got[k], ref[k], err_msg=k)
    np.testing.assert_array_equal(got["state"]["draw"], ref["state"]["draw"])
    np.testing.assert_array_equal(got["state"]["key"], ref["state"]["key"])


def test_g1_end_to_end(cuda):
    """The hello-world script of the reference (README.md:93-129), same calls, same prints."""
    import foragax_b200.base.agent_classes as fgx_classes
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import random as fgx_random
    from foragax_b200.examples.hello_world import Dice, is_one, is_six

    printed = []
    draws = lambda s: s.agents.state.content['draw'].reshape(-1).cpu().numpy().tolist()  # noqa: E731
    Dice_set = fgx_classes.Agent_Set(agent=Dice, num_total_agents=10, num_active_agents=5, agent_type=0)
    Dice_set.agents = fgx_methods.create_agents(params=None, agent_set=Dice_set, key=fgx_random.PRNGKey(0))
    printed.append(draws(Dice_set))
    Dice_set.agents = fgx_methods.step_agents(params=None, agent_set=Dice_set, input=None)
    printed.append(draws(Dice_set))
    num_agents_dead, remove_indices = fgx_methods.jit_select_agents(select_func=is_one, select_params=None,
                                                                    agents=Dice_set.agents)
    dice_remove_params = fgx_classes.Params(content={'remove_ids': remove_indices})
    Dice_set.agents = fgx_methods.jit_remove_agents(remove_func=Dice.remove_agent, num_agents_remove=num_agents_dead,
                                                    remove_params=dice_remove_params, agents=Dice_set.agents)
    printed.append(draws(Dice_set))
    Dice_set.agents, sorted_indices = fgx_methods.jit_sort_agents(quantity=Dice_set.agents.active_state, ascend=False,
                                                                  agents=Dice_set.agents)
    printed.append(draws(Dice_set))
    printed.append(Dice_set.agents.active_state.cpu().numpy().tolist())
    num_agents_add, add_indices = fgx_methods.jit_select_agents(select_func=is_six, select_params=None,
                                                                agents=Dice_set.agents)
    num_active_agents = fgx_methods.count_active(Dice_set.agents)
    num_agents_add = torch.minimum(num_agents_add, Dice_set.num_total_agents - num_active_agents)
    Dice_set.agents, key = fgx_methods.jit_add_agents(add_func=Dice.add_agent, num_agents_add=num_agents_add,
                                                      add_params=None, agents=Dice_set.agents, key=None)
    printed.append(draws(Dice_set))
    printed.append(Dice_set.agents.active_state.cpu().numpy().tolist())
    assert printed == G1["printed"]
    assert sorted_indices.cpu().numpy().tolist() == [0, 1, 2, 4, 3, 5, 6, 7, 8, 9]
    assert int(num_agents_dead) == 1


@pytest.mark.parametrize("n,n_active,sides,ticks", [(10, 5, 6, 6), (1000, 500, 6, 5), (5000, 2500, 20, 4),
                                                    (20000, 9000, 20, 3)])
def test_churn_ticks_match_oracle(cuda, n, n_active, sides, ticks):
    """create -> (step -> select -> remove -> sort -> select -> clip -> add) x ticks, compared
    leaf by leaf with the oracle after every tick."""
    import foragax_b200.base.agent_classes as fgx_classes
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import random as fgx_random
    from foragax_b200.examples.hello_world import is_value, make_dice
    import functools

    D = make_dice(sides)
    s = fgx_classes.Agent_Set(agent=D, num_total_agents=n, num_active_agents=n_active, agent_type=3)
    s.agents = fgx_methods.create_agents(None, s, fgx_random.PRNGKey(n))
    ref = oam.create_agents(functools.partial(odice.create_agent, sides=sides), None, n, n_active, 3, jr.PRNGKey(n))
    assert_same(s.agents, ref)
    for _ in range(ticks):
        s.agents = fgx_methods.step_agents(None, None, s)
        n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, s.agents)
        s.agents = fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}),
                                                 s.agents)
        s.agents, _ = fgx_methods.jit_sort_agents(s.agents.active_state, False, s.agents)
        n_add, _ = fgx_methods.jit_select_agents(is_value(sides), None, s.agents)
        n_act = fgx_methods.count_active(s.agents)
        n_add = torch.minimum(n_add, n - n_act)
        s.agents, _ = fgx_methods.jit_add_agents(D.add_agent, n_add, None, s.agents, None)
        ref = odice.churn_tick(ref, n, sides)
        assert_same(s.agents, ref)


@pytest.mark.parametrize("n", [1, 2, 31, 32, 33, 1024, 1025, 2047, 2048, 2049, 6000, 8193, 16384, 16385, 100003,
                               1 << 20, 3_000_017, 10_000_019])  # the last two: several 256-slot groups per warp
@pytest.mark.parametrize("frac", [0.0, 0.3, 1.0])
def test_select_partition(cuda, n, frac):
    import foragax_b200.base.agent_methods as fgx_methods
    rng = np.random.default_rng(n)
    mask = rng.random(n) < frac
    for arr in (mask, mask.astype(np.int32) * 7, mask.astype(np.float32) * 0.5):
        t = torch.from_numpy(arr).to(cuda)
        cnt, perm = fgx_methods.select_agents(lambda a, p: t, None, None)
        rc, rp = oam.select_from_mask(mask)
        assert int(cnt) == int(rc)
        np.testing.assert_array_equal(perm.cpu().numpy(), rp)


def test_select_fused_predicates(cuda):
    """The compound predicates of the reference's examples (sim.py:477-481,513-515;
    wolf_sheep_grass/sim.ipynb select_dead_sheeps), evaluated in-kernel."""
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200.base.predicates import P
    rng = np.random.default_rng(3)
    n = 50021
    active = (rng.random(n) < 0.6).astype(np.float32)
    tb = rng.random(n).astype(np.float32) * 6
    age = rng.random(n).astype(np.float32) * 1200
    energy = rng.integers(-3, 5, n).astype(np.int32)
    eaten = rng.random(n) < 0.2
    d = lambda a: torch.from_numpy(a).to(cuda)  # noqa: E731
    A, TB, AGE, EN, EAT = d(active), d(tb).reshape(n, 1), d(age), d(energy).reshape(n, 1), d(eaten)
    cases = [
        (((P(A) == 1.0) & (P(TB) > 3.0)) | (P(AGE) > 1000.0), ((active == 1.0) & (tb > 3.0)) | (age > 1000.0)),
        ((P(A) == 1.0) & (P(TB) > 0.4), (active == 1.0) & (tb > np.float32(0.4))),
        (((P(EN) <= 0) | P(EAT).nonzero()) & P(A).nonzero(), ((energy <= 0) | eaten) & (active != 0)),
        (~(P(EN) == 2), energy != 2),
    ]
    for pred, mask in cases:
        cnt, perm = fgx_methods.select_agents(lambda a, p: pred, None, None)
        rc, rp = oam.select_from_mask(mask)
        assert int(cnt) == int(rc)
        np.testing.assert_array_equal(perm.cpu().numpy(), rp)


@pytest.mark.parametrize("n", [20_000, 70_001])
def test_select_single_term_operators(cuda, n):
    """The one-term fast path of fgx_select (a closed-range test per operator, compact.cu::fast_range) for every
    operator, on float and int32 leaves with the edge immediates: +-0, +-inf, NaN, INT_MIN / INT_MAX; leaves contain
    NaN, +-inf, +-0 and the extreme integers.  Both the cooperative one-pass kernel (n > 16384) and its partial last
    group are exercised; the reference is NumPy's comparison."""
    import operator

    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200.base.predicates import P
    rng = np.random.default_rng(n)
    f = rng.standard_normal(n).astype(np.float32)
    f[rng.integers(0, n, 50)] = np.nan
    f[rng.integers(0, n, 50)] = np.inf
    f[rng.integers(0, n, 50)] = -np.inf
    f[rng.integers(0, n, 50)] = 0.0
    f[rng.integers(0, n, 50)] = -0.0
    f[rng.integers(0, n, 50)] = 1.5
    i = rng.integers(-5, 6, n).astype(np.int32)
    i[rng.integers(0, n, 50)] = np.iinfo(np.int32).min
    i[rng.integers(0, n, 50)] = np.iinfo(np.int32).max
    F, I = torch.from_numpy(f).to(cuda), torch.from_numpy(i).to(cuda)
    ops = [operator.eq, operator.ne, operator.lt, operator.le, operator.gt, operator.ge]

    def check(pred, mask):
        cnt, perm = fgx_methods.select_agents(lambda a, p: pred, None, None)
        rc, rp = oam.select_from_mask(mask)
        assert int(cnt) == int(rc)
        np.testing.assert_array_equal(perm.cpu().numpy(), rp)

    with np.errstate(invalid="ignore"):
        for c in (0.0, -0.0, 1.5, -0.25, float("inf"), float("-inf"), float("nan")):
            for op in ops:
                check(op(P(F), c), op(f, np.float32(c)))
                check(~op(P(F), c), ~op(f, np.float32(c)))
        for c in (0, 2, -5, 5, int(np.iinfo(np.int32).min), int(np.iinfo(np.int32).max)):
            for op in ops:
                check(op(P(I), c), op(i, np.int32(c)))
    check(P(F).nonzero(), f != 0)
    check(P(I).nonzero(), i != 0)


@pytest.mark.parametrize("n", [1, 2, 100, 256, 257, 1000, 1025, 4096, 6000, 16384, 16385, 50000, 1 << 20, 3_000_017])
@pytest.mark.parametrize("kind", ["flag", "float", "float_special", "int", "int_small", "const"])
@pytest.mark.parametrize("ascend", [True, False])
def test_argsort(cuda, n, kind, ascend):
    import foragax_b200.base.agent_methods as fgx_methods
    rng = np.random.default_rng(n + len(kind))
    if kind == "flag":
        q = (rng.random(n) < 0.5).astype(np.float32)
    elif kind == "float":
        q = rng.standard_normal(n).astype(np.float32) * 100
    elif kind == "float_special":
        q = rng.standard_normal(n).astype(np.float32)
        q[rng.random(n) < 0.2] = 0.0
        q[rng.random(n) < 0.1] = -0.0
        q[rng.random(n) < 0.05] = np.nan
        q[rng.random(n) < 0.05] = np.inf
        q[rng.random(n) < 0.05] = -np.inf
        q = np.round(q, 1)  # many ties
    elif kind == "int":
        q = rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)
        q[0] = -2**31
    elif kind == "int_small":
        q = rng.integers(0, 2, n).astype(np.int32)
    else:
        q = np.full(n, 3.5, np.float32)
    got = fgx_methods.argsort(torch.from_numpy(q).to(cuda).reshape(n, 1), ascend).cpu().numpy()
    np.testing.assert_array_equal(got, oam.sort_ids(q, ascend))


def test_sort_agents_moves_every_leaf(cuda):
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200.base.agent_classes import Agent, Params, Policy, State
    n = 3001
    rng = np.random.default_rng(0)
    leaves = {
        "a": rng.standard_normal((n, 1)).astype(np.float32),
        "w": rng.standard_normal((n, 40, 40)).astype(np.float32),   # 6400-byte rows
        "k": rng.integers(0, 2**32, (n, 2), dtype=np.uint32),
        "b": rng.random(n) < 0.5,                                   # 1-byte rows
        "odd": rng.integers(0, 255, (n, 3)).astype(np.uint8),       # 3-byte rows
    }
    d = {k: torch.from_numpy(v).to(cuda) for k, v in leaves.items()}
    ag = Agent(params=Params(content={"w": d["w"]}), state=State(content={"a": d["a"], "k": d["k"], "b": d["b"]}),
               unique_id=torch.arange(n, dtype=torch.int32, device=cuda), agent_type=torch.zeros(n, dtype=torch.int32, device=cuda),
               active_state=torch.ones(n, device=cuda), age=torch.zeros(n, device=cuda),
               policy=Policy(state=State(content={"odd": d["odd"]}), params=None, deterministic=True))
    new, ids = fgx_methods.jit_sort_agents(ag.state.content["a"], True, ag)
    ref_ids = oam.sort_ids(leaves["a"], True)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)
    np.testing.assert_array_equal(new.state.content["a"].cpu().numpy(), leaves["a"][ref_ids])
    np.testing.assert_array_equal(new.params.content["w"].cpu().numpy(), leaves["w"][ref_ids])
    np.testing.assert_array_equal(new.state.content["k"].cpu().numpy(), leaves["k"][ref_ids])
    np.testing.assert_array_equal(new.state.content["b"].cpu().numpy(), leaves["b"][ref_ids])
    np.testing.assert_array_equal(new.policy.state.content["odd"].cpu().numpy(), leaves["odd"][ref_ids])
    np.testing.assert_array_equal(new.unique_id.cpu().numpy(), ref_ids)
    assert new.policy.deterministic is True


def test_non_native_callbacks_are_rejected(cuda):
    import foragax_b200.base.agent_methods as fgx_methods
    with pytest.raises(TypeError, match="no CPU fallback"):
        fgx_methods.jit_remove_agents(lambda p, a, i: a, 1, None, None)


@pytest.mark.parametrize("world,stride_mode", [(3, "round_robin"), (4, "blocks")])
def test_sharded_create_equals_global_rows(cuda, world, stride_mode):
    """Every rank of a multi-GPU run builds exactly its rows of the single-GPU set (fgx_create_base_shard)."""
    import foragax_b200.base.agent_classes as fgx_classes
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import random as fgx_random
    from foragax_b200.examples.hello_world import Dice
    n, n_active = 10007, 6001
    s = fgx_classes.Agent_Set(agent=Dice, num_total_agents=n, num_active_agents=n_active, agent_type=2)
    full = to_np(fgx_methods.create_agents(None, s, fgx_random.PRNGKey(9)))
    for r in range(world):
        if stride_mode == "round_robin":
            rows = np.arange(r, n, world)
            shard = (r, len(rows), world)
        else:
            lo, hi = r * n // world, (r + 1) * n // world
            rows = np.arange(lo, hi)
            shard = (lo, hi - lo)
        part = to_np(fgx_methods.create_agents(None, s, fgx_random.PRNGKey(9), shard=shard))
        for k in ("unique_id", "agent_type", "active_state", "age"):
            np.testing.assert_array_equal(part[k], full[k][rows], err_msg=k)
        np.testing.assert_array_equal(part["state"]["draw"], full["state"]["draw"][rows])
        np.testing.assert_array_equal(part["state"]["key"], full["state"]["key"][rows])


def test_argsort_one_launch_per_phase_variant(cuda):
    """FGX_SORT_MULTI=1 (one launch per radix phase instead of the cooperative single launch) gives the same
    permutation; the switch is read once per process, hence the subprocess."""
    import subprocess
    import sys
    code = ("import torch, foragax_b200.base.agent_methods as am\n"
            "g = torch.Generator(device='cuda').manual_seed(1)\n"
            "for n in (20000, 300001):\n"
            "    k = torch.randint(-1000, 1000, (n,), device='cuda', generator=g).float() * 0.25\n"
            "    for asc in (True, False):\n"
            "        ref = torch.sort(k if asc else -k, stable=True).indices.to(torch.int32)\n"
            "        assert torch.equal(am.argsort(k, ascend=asc), ref)\n"
            "print('ok')\n")
    env = dict(os.environ, FGX_SORT_MULTI="1", PYTHONPATH=os.path.dirname(os.path.dirname(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("n", [16385, 100003, 2_500_037])
@pytest.mark.parametrize("key_kind", ["active", "flag_int", "three_valued", "all_equal"])
@pytest.mark.parametrize("ascend", [True, False])
def test_sort_agents_fused_rows(cuda, n, key_kind, ascend):
    """jit_sort_agents = fgx_sort_rows: on a two-valued key the cooperative kernel partitions AND moves the
    one-word rows (4 / 8 / 16 bytes) in one pass; other leaves (bytes, 12-byte rows, wide rows) and keys with
    more values take the gather.  Every leaf must equal leaf[argsort] (agent_methods.py:74-79)."""
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200.base.agent_classes import Agent, Params, Policy, State
    rng = np.random.default_rng(n)
    act = (rng.random(n) < 0.6).astype(np.float32)
    leaves = {
        "draw": rng.integers(0, 21, (n, 1)).astype(np.int32),            # 4-byte rows
        "key": rng.integers(0, 2**32, (n, 2), dtype=np.uint32),          # 8-byte rows
        "quad": rng.standard_normal((n, 4)).astype(np.float32),          # 16-byte rows
        "tri": rng.standard_normal((n, 3)).astype(np.float32),           # 12-byte rows (3 chunks)
        "b": rng.random(n) < 0.5,                                        # 1-byte rows
        "w": rng.standard_normal((n, 40)).astype(np.float32),            # 160-byte rows (wide)
    }
    d = {k: torch.from_numpy(v).to(cuda) for k, v in leaves.items()}
    uid = np.arange(n, dtype=np.int32)
    ag = Agent(params=Params(content={"w": d["w"], "quad": d["quad"]}),
               state=State(content={"draw": d["draw"], "key": d["key"], "tri": d["tri"], "b": d["b"]}),
               unique_id=torch.from_numpy(uid).to(cuda), agent_type=torch.full((n,), 3, dtype=torch.int32, device=cuda),
               active_state=torch.from_numpy(act).to(cuda), age=torch.arange(n, dtype=torch.float32, device=cuda),
               policy=None)
    if key_kind == "active":
        q, qt = act, ag.active_state
    elif key_kind == "flag_int":
        q = (leaves["draw"][:, 0] > 10).astype(np.int32)
        qt = torch.from_numpy(q).to(cuda)
    elif key_kind == "three_valued":
        q = (leaves["draw"][:, 0] % 3).astype(np.float32)
        qt = torch.from_numpy(q).to(cuda).reshape(n, 1)
    else:
        q = np.full(n, 1.0, np.float32)
        qt = torch.from_numpy(q).to(cuda)
    new, ids = fgx_methods.jit_sort_agents(qt, ascend, ag)
    ref_ids = oam.sort_ids(q, ascend)
    np.testing.assert_array_equal(ids.cpu().numpy(), ref_ids)
    np.testing.assert_array_equal(new.unique_id.cpu().numpy(), uid[ref_ids])
    np.testing.assert_array_equal(new.active_state.cpu().numpy(), act[ref_ids])
    np.testing.assert_array_equal(new.age.cpu().numpy(), np.arange(n, dtype=np.float32)[ref_ids])
    np.testing.assert_array_equal(new.agent_type.cpu().numpy(), np.full(n, 3, np.int32))
    for name, tree in (("draw", new.state.content), ("key", new.state.content), ("tri", new.state.content),
                       ("b", new.state.content), ("w", new.params.content), ("quad", new.params.content)):
        np.testing.assert_array_equal(tree[name].cpu().numpy(), leaves[name][ref_ids], err_msg=name)
    # the input set is untouched (sort_agents gathers into fresh buffers)
    np.testing.assert_array_equal(ag.state.content["draw"].cpu().numpy(), leaves["draw"])


def test_two_pass_variants_of_select_and_sort(cuda):
    """FGX_SELECT_TWO_PASS=1 / FGX_SORT_TWO_PASS=1 (count pass + scatter pass instead of the cooperative
    single pass; also what sets beyond ~77 M slots use) give the same permutations; the switches are read
    once per process, hence the subprocess."""
    import subprocess
    import sys
    code = ("import torch, foragax_b200.base.agent_methods as am\n"
            "g = torch.Generator(device='cuda').manual_seed(2)\n"
            "for n in (20000, 300001, 2000003):\n"
            "    f = (torch.rand(n, device='cuda', generator=g) < 0.4)\n"
            "    cnt, perm = am.select_agents(lambda a, p: f, None, None)\n"
            "    ref = torch.sort((~f).to(torch.int32), stable=True).indices.to(torch.int32)\n"
            "    assert int(cnt) == int(f.sum()) and torch.equal(perm, ref)\n"
            "    k = f.float()\n"
            "    for asc in (True, False):\n"
            "        ref = torch.sort(k if asc else -k, stable=True).indices.to(torch.int32)\n"
            "        assert torch.equal(am.argsort(k, ascend=asc), ref)\n"
            "print('ok')\n")
    env = dict(os.environ, FGX_SELECT_TWO_PASS="1", FGX_SORT_TWO_PASS="1",
               PYTHONPATH=os.path.dirname(os.path.dirname(__file__)))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


@pytest.mark.parametrize("n", [40_000, 300_017])
def test_churn_world_matches_oracle(cuda, n):
    """examples/churn.py (the C5 tick bench.py times: one-pass select, fused sort + row move, count + clip in one
    launch, add with the count passed in) against the oracle's hello-world tick, leaf by leaf, every tick."""
    import functools
    from foragax_b200.examples.churn import ChurnWorld
    w = ChurnWorld(n, sides=20, n_active=n // 2, key=None)
    ref = oam.create_agents(functools.partial(odice.create_agent, sides=20), None, n, n // 2, 0, jr.PRNGKey(0))
    assert_same(w.agents, ref)
    for _ in range(4):
        n_act = w.tick()
        ref = odice.churn_tick(ref, n, 20)
        assert_same(w.agents, ref)
    assert 0 < int(n_act) <= n  # the count before the add of the last tick


def test_churn_world_graph_replay_matches_oracle(cuda):
    """ChurnWorld.capture_two_ticks: the sort alternates between two fixed buffer sets (sort_agents(out=)), two ticks
    are one CUDA graph; after 2 warm-up ticks and k replays the state equals 2 + 2k oracle ticks leaf by leaf."""
    import functools
    from foragax_b200.examples.churn import ChurnWorld
    n = 120_011
    w = ChurnWorld(n, sides=20, n_active=n // 2, key=None)
    ref = oam.create_agents(functools.partial(odice.create_agent, sides=20), None, n, n // 2, 0, jr.PRNGKey(0))
    graph = w.capture_two_ticks()
    for _ in range(2):
        ref = odice.churn_tick(ref, n, 20)
    assert_same(w.agents, ref)  # capturing does not execute
    for _ in range(3):
        graph.replay()
        for _ in range(2):
            ref = odice.churn_tick(ref, n, 20)
        torch.cuda.synchronize()
        assert_same(w.agents, ref)
    # an eager tick after the replays continues from the same state (and flips the buffer sets)
    w.tick()
    assert_same(w.agents, odice.churn_tick(ref, n, 20))


def test_sort_agents_out_buffers(cuda):
    """sort_agents(out=): rows land in the caller's buffers, bit-identical to the fresh-buffer call; misuse raises."""
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import struct
    from foragax_b200._lib import FgxError
    from foragax_b200.examples.churn import ChurnWorld
    w = ChurnWorld(50_003, sides=6, n_active=20_000, key=None)
    w.tick()
    ag = w.agents
    want, perm = fgx_methods.jit_sort_agents(ag.active_state, False, ag)
    out = struct.tree_map(torch.empty_like, ag)
    got, perm2 = fgx_methods.jit_sort_agents(ag.active_state, False, ag, out=out)
    assert torch.equal(perm, perm2)
    for a, b, c in zip(struct.tree_leaves(want), struct.tree_leaves(got), struct.tree_leaves(out)):
        assert torch.equal(a, b) and b.data_ptr() == c.data_ptr()
    with pytest.raises(FgxError):
        fgx_methods.jit_sort_agents(ag.active_state, False, ag, out=ag)


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 1023, 100_003, 2_000_001])
def test_count_active_and_clip(cuda, n):
    """fgx_count_active / fgx_count_active_clip (agent_methods.py:27 + README.md:124-125): float4 body + scalar tail,
    16-byte-misaligned views (scalar path), the last-CTA clip epilogue."""
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200.base.agent_classes import Agent
    rng = np.random.default_rng(n)
    flags = (rng.random(n + 3) < 0.37).astype(np.float32)
    base = torch.from_numpy(flags).to(cuda)
    for off in (0, 1, 3):  # data pointers at +0, +4, +12 bytes
        act = base[off:off + n]
        ag = Agent(params=None, state=None, unique_id=torch.zeros(n, dtype=torch.int32, device=cuda), agent_type=None,
                   active_state=act, age=None, policy=None)
        want = int(flags[off:off + n].sum())
        assert int(fgx_methods.count_active(ag)) == want
        for n_add in (0, 5, n, 2 * n + 7):
            n_act, clipped = fgx_methods.count_active_clip(ag, n_add)
            assert int(n_act) == want and int(clipped) == min(n_add, n - want)
            n_act, clipped = fgx_methods.count_active_clip(ag, torch.tensor([n_add], dtype=torch.int32, device=cuda))
            assert int(n_act) == want and int(clipped) == min(n_add, n - want)


@pytest.mark.parametrize("n", [0, 1, 257, 4096, 1_000_003])
def test_deterministic_float_sum(cuda, n):
    """fgx_sum_f32 (the age sums of Neuroevolution, sim.py:486,491): exact for integer-valued leaves (ages are tick
    counts), float32-accurate and bit-identical run to run otherwise."""
    from foragax_b200.examples.foraging import _sum_f32
    rng = np.random.default_rng(n + 1)
    ages = rng.integers(0, 16, n).astype(np.float32)  # sums stay below 2^24: exactly representable
    assert float(_sum_f32(torch.from_numpy(ages).to(cuda))) == float(ages.astype(np.float64).sum())
    x = rng.standard_normal(n).astype(np.float32)
    xt = torch.from_numpy(x).to(cuda)
    a, b = _sum_f32(xt), _sum_f32(xt)
    assert torch.equal(a, b)
    ref = float(x.astype(np.float64).sum())
    assert abs(float(a) - ref) <= 1e-6 * float(np.abs(x).astype(np.float64).sum()) + 1e-6
