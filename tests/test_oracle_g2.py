"""Oracle pinning on the reference's golden vector G2: the (num_active_foragers,
num_active_resources, life_expectancy) triples printed by EcoEvoJax/sim.ipynb under PRNGKey(0)
(tests/golden/g2_foraging.json, extracted by tests/golden/make_g2.py).

The full restatement -- create (uniform with bounds, WilsonCowan weights), interaction, the ray
sensor composition (ray_generator, ray_wall_sensor, ray_agent_sensor, argmin), the policy,
Resource blossom (bernoulli), select / remove / stable descending sort / add with the serial key
chain and normal() mutation / set -- reproduces ticks 0..40 exactly (births start at tick 6, deaths
at tick 40).  From tick 41 the trajectory is chaotic (one forager crosses a threshold a tick
earlier), so later ticks are not a parity criterion."""
import json
import os

from oracle import foraging as fo

G2 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "g2_foraging.json")))["ticks"]
PINNED_TICKS = 41


def test_g2_ticks_0_to_40():
    w = fo.World(fo.Config("nb"))
    for i in range(PINNED_TICKS):
        nf, nr, life = w.step()
        assert (nf, nr) == (G2[i][0], G2[i][1]), f"tick {i}: {(nf, nr)} != {G2[i][:2]}"
        assert abs(life - G2[i][2]) <= 1e-4 * max(1.0, abs(G2[i][2])), f"tick {i}: life {life} != {G2[i][2]}"
    assert G2[6][0] == 873 and G2[40][0] < G2[39][0]  # the window covers births and the first deaths
