"""K2 parity for the wolf-sheep-grass example (BASELINE config 3): Animal / Grass / interaction /
Ecosystem on the CUDA path vs the oracle.  All state is integer, flags or RNG keys -> bit-exact, tick by
tick, including the cell-equality interaction (histogram join vs the oracle's equality matrices) and the
float->int32 truncation of `energy / 2`."""
import numpy as np
import pytest
import torch

from oracle import wolf_sheep as ows

pytestmark = pytest.mark.gpu


def np_animal(a):
    c = a.state.content
    return {"state": {k: c[k].cpu().numpy() for k in ("X_pos", "Y_pos", "energy", "reproduce", "key")},
            "policy_key": a.policy.state.content["key"].cpu().numpy(), "unique_id": a.unique_id.cpu().numpy(),
            "agent_type": a.agent_type.cpu().numpy(), "active_state": a.active_state.cpu().numpy(),
            "age": a.age.cpu().numpy()}


def np_grass(g):
    c = g.state.content
    return {"state": {k: c[k].cpu().numpy() for k in ("X_pos", "Y_pos", "fully_grown", "count_down")},
            "unique_id": g.unique_id.cpu().numpy(), "active_state": g.active_state.cpu().numpy()}


def assert_equal_tree(got, ref, path=""):
    for k, v in ref.items():
        if isinstance(v, dict):
            assert_equal_tree(got[k], v, path + k + ".")
        elif k in got:
            np.testing.assert_array_equal(got[k], v, err_msg=path + k)


@pytest.mark.parametrize("cfg,ticks", [
    (dict(X_max=25, Y_max=25, max_animals=300, init_sheeps=100, init_wolves=50), 80),      # the reference's sizes
    (dict(X_max=120, Y_max=90, max_animals=20000, init_sheeps=9000, init_wolves=3000), 25),
])
def test_ecosystem_ticks_bit_exact(cuda, cfg, ticks):
    from foragax_b200 import random as fgx_random
    from foragax_b200.examples.wolf_sheep_grass import Ecosystem
    eco = Ecosystem(grass_regrowth_time=30, wolf_reproduction_rate=0.08, wolf_energy=30, init_wolves=cfg["init_wolves"],
                    sheep_reproduction_rate=0.2, sheep_energy=5, init_sheeps=cfg["init_sheeps"], X_max=cfg["X_max"],
                    Y_max=cfg["Y_max"], sim_steps=ticks, key=fgx_random.PRNGKey(0), max_animals=cfg["max_animals"])
    ref = ows.Ecosystem(init_wolves=cfg["init_wolves"], init_sheeps=cfg["init_sheeps"], X_max=cfg["X_max"],
                        Y_max=cfg["Y_max"], max_animals=cfg["max_animals"])
    assert_equal_tree(np_animal(eco.sheeps.agents), ref.sheeps)
    assert_equal_tree(np_animal(eco.wolves.agents), ref.wolves)
    assert_equal_tree(np_grass(eco.grasses.agents), ref.grasses)
    seen = set()
    for t in range(ticks):
        ns, nw = eco.step()
        rs, rw = ref.step()
        assert (int(ns), int(nw)) == (rs, rw), f"tick {t}"
        seen.add((rs, rw))
        if t % 5 == 4 or t == ticks - 1:
            assert_equal_tree(np_animal(eco.sheeps.agents), ref.sheeps)
            assert_equal_tree(np_animal(eco.wolves.agents), ref.wolves)
            assert_equal_tree(np_grass(eco.grasses.agents), ref.grasses)
    assert len(seen) > ticks // 3  # the populations actually move (births and deaths every tick)


def test_interaction_matches_equality_matrices(cuda):
    from foragax_b200.examples.wolf_sheep_grass import interaction
    from foragax_b200.base.agent_classes import Agent, Params, Policy, State
    rng = np.random.default_rng(1)
    xm, ym, nw, ns = 40, 30, 5000, 7000
    mk = lambda n: {"X_pos": rng.integers(-1, xm, (n, 1)).astype(np.int32), "Y_pos": rng.integers(-1, ym, (n, 1)).astype(np.int32)}  # noqa: E731
    w, s = mk(nw), mk(ns)
    w["X_pos"][:50], w["Y_pos"][:50] = -1, -1       # inactive animals sit at (-1,-1) and match each other
    s["X_pos"][:70], s["Y_pos"][:70] = -1, -1
    ng = xm * ym
    g = {"X_pos": (np.arange(ng) % xm).astype(np.int32)[:, None], "Y_pos": (np.arange(ng) // xm).astype(np.int32)[:, None],
         "fully_grown": rng.random(ng) < 0.5}
    d = lambda a: torch.from_numpy(a).to(cuda)  # noqa: E731
    mkag = lambda st, n: Agent(params=None, state=State(content={k: d(v) for k, v in st.items()}),  # noqa: E731
                               unique_id=torch.arange(n, dtype=torch.int32, device=cuda), agent_type=None,
                               active_state=None, age=None, policy=None)
    from oracle.wolf_sheep import interaction as oint
    got = interaction(mkag(w, nw), mkag(s, ns), mkag(g, ng))
    ref = oint({"state": w}, {"state": s}, {"state": g})
    for a, b in zip(got, ref):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    assert ref[0].max() > 3 and ref[2].sum() > 100


def test_cuda_graph_replay_equals_eager_ticks(cuda):
    """The whole Ecosystem tick captured once and replayed (Ecosystem.capture) stays on the oracle's
    trajectory: counts every tick, full state at the end."""
    from foragax_b200 import random as fgx_random
    from foragax_b200.examples.wolf_sheep_grass import Ecosystem
    cfg = dict(X_max=60, Y_max=45, max_animals=6000, init_sheeps=2500, init_wolves=800)
    eco = Ecosystem(grass_regrowth_time=30, wolf_reproduction_rate=0.08, wolf_energy=30, init_wolves=cfg["init_wolves"],
                    sheep_reproduction_rate=0.2, sheep_energy=5, init_sheeps=cfg["init_sheeps"], X_max=cfg["X_max"],
                    Y_max=cfg["Y_max"], sim_steps=30, key=fgx_random.PRNGKey(0), max_animals=cfg["max_animals"])
    ref = ows.Ecosystem(init_wolves=cfg["init_wolves"], init_sheeps=cfg["init_sheeps"], X_max=cfg["X_max"],
                        Y_max=cfg["Y_max"], max_animals=cfg["max_animals"])
    g = eco.capture(warmup=2)
    for _ in range(2):
        ref.step()
    for t in range(2, 30):
        ns, nw = g.replay()
        rs, rw = ref.step()
        assert (int(ns), int(nw)) == (rs, rw), f"tick {t}"
    assert_equal_tree(np_animal(eco.sheeps.agents), ref.sheeps)
    assert_equal_tree(np_animal(eco.wolves.agents), ref.wolves)
    assert_equal_tree(np_grass(eco.grasses.agents), ref.grasses)
