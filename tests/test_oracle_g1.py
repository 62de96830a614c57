"""Oracle pinning on the reference's own golden vector G1: the seven arrays printed by
examples/basic/hello_world/hello_world.ipynb:92-99 (run under jax.random.PRNGKey(0))."""
import json
import os

import numpy as np

from oracle import agent_methods as am
from oracle import dice
from oracle import jax_random as jr

G1 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "g1_hello_world.json")))


def run_g1():
    out = []
    a = am.create_agents(dice.create_agent, None, 10, 5, 0, jr.PRNGKey(0))
    out.append(a["state"]["draw"].reshape(-1).tolist())
    a = dice.step_agent(None, None, a)
    out.append(a["state"]["draw"].reshape(-1).tolist())
    n, rem = am.select_agents(dice.is_value(1), None, a)
    a = am.remove_agents(dice.remove_agent, n, {"remove_ids": rem}, a)
    out.append(a["state"]["draw"].reshape(-1).tolist())
    a, ids = am.sort_agents(a["active_state"], False, a)
    out.append(a["state"]["draw"].reshape(-1).tolist())
    out.append(a["active_state"].tolist())
    n6, _ = am.select_agents(dice.is_value(6), None, a)
    n6 = min(int(n6), 10 - int(a["active_state"].sum()))
    a, _ = am.add_agents(dice.add_agent, n6, None, a, None)
    out.append(a["state"]["draw"].reshape(-1).tolist())
    out.append(a["active_state"].tolist())
    return out, ids


def test_g1_bit_exact():
    out, ids = run_g1()
    assert out == G1["printed"]
    assert ids.tolist() == [0, 1, 2, 4, 3, 5, 6, 7, 8, 9]


def test_select_is_stable_partition():
    rng = np.random.default_rng(0)
    for n in (1, 2, 31, 257, 5000):
        mask = rng.random(n) < 0.3
        cnt, perm = am.select_from_mask(mask)
        idx = np.arange(n)
        assert cnt == mask.sum()
        np.testing.assert_array_equal(perm, np.concatenate([idx[mask], idx[~mask]]))


def test_sort_semantics():
    q = np.array([0.0, -0.0, np.nan, 1.0, -np.inf, np.inf, 1.0, -1.0], np.float32)
    # ascending: -inf, -1, (+-0 tie: slot order), 1, 1, inf, nan
    assert am.sort_ids(q, True).tolist() == [4, 7, 0, 1, 3, 6, 5, 2]
    # descending = argsort(-q): ties keep ascending slot order, NaN still last
    assert am.sort_ids(q, False).tolist() == [5, 3, 6, 0, 1, 7, 4, 2]
    qi = np.array([3, -1, 3, 0, -2147483648], np.int32)
    assert am.sort_ids(qi, True).tolist() == [4, 1, 3, 0, 2]
    assert am.sort_ids(qi, False).tolist() == [4, 0, 2, 3, 1]  # -INT_MIN wraps to INT_MIN
