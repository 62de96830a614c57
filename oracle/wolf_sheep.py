"""NumPy restatement of the wolf-sheep-grass example
(``examples/basic/wolf_sheep_grass/sim.ipynb``: ``Animal`` :37-170, ``Grass`` :174-218, ``interaction``
:238-290, ``Ecosystem`` :365-517), the churn workload of BASELINE config 3.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED by reference data (the notebook's
only output is a PNG plot); all state is integer / RNG, restated from source.  vmapped functions are
batched, ``fori_loop`` bodies stay serial.  As in the reference the interaction has no active-state
check (inactive animals sit at (-1,-1) and match each other there).
"""
import numpy as np

from . import agent_methods as am
from . import jax_random as jr

F = np.float32
I = np.int32


def animal_create(cfg, uid, active, atype, keys):
    """vmapped ``Animal.create_agent`` (sim.ipynb:40-70); cfg = dict(x_max, y_max, delta_energy, reproduction_rate)."""
    n = len(uid)
    k = jr.split(keys, 2)
    key, pkey = k[:, 0], k[:, 1]
    s4 = jr.split(key, 4)
    a = (active != 0)[:, None]
    X = jr.randint(s4[:, 1], 1, 0, cfg["x_max"] - 1)
    Y = jr.randint(s4[:, 2], 1, 0, cfg["y_max"] - 1)
    E = jr.randint(s4[:, 3], 1, 1, 2 * cfg["delta_energy"])
    st = {"X_pos": np.where(a, X, I(-1)).astype(I), "Y_pos": np.where(a, Y, I(-1)).astype(I),
          "energy": np.where(a, E, I(-1)).astype(I), "reproduce": np.zeros(n, I),
          "key": np.where(a, s4[:, 0], key).astype(np.uint32)}
    return {"state": st, "policy_key": pkey.astype(np.uint32), "unique_id": uid.astype(I), "agent_type": atype.astype(I),
            "active_state": active.astype(F), "age": np.zeros(n, F)}


def animal_step(cfg, energy_in, an):
    """vmapped ``Animal.step_agent`` (sim.ipynb:72-122) with ``Random_Policy.step_policy``
    (random_policy.py:17-23); energy_in int32[n]."""
    s = an["state"]
    pk = jr.split(an["policy_key"], 2)
    action = jr.uniform(pk[:, 0], 1)[:, 0]
    X, Y = s["X_pos"][:, 0], s["Y_pos"][:, 0]
    Xn = np.where(action < F(0.25), X + 1, X)
    Xn = np.where((action >= F(0.25)) & (action < F(0.5)), X - 1, Xn)
    Yn = np.where((action >= F(0.5)) & (action < F(0.75)), Y + 1, Y)
    Yn = np.where(action >= F(0.75), Y - 1, Yn)
    xmin, xmax, ymin, ymax = 0, cfg["x_max"], 0, cfg["y_max"]
    tor = cfg.get("torous", True)
    Xn = np.where(tor & (Xn < xmin), xmax - 1, Xn)
    Xn = np.where(tor & (Xn > xmax), xmin, Xn)
    Yn = np.where(tor & (Yn < ymin), ymax - 1, Yn)
    Yn = np.where(tor & (Yn > ymax), ymin, Yn)
    Xn, Yn = np.clip(Xn, xmin, xmax - 1), np.clip(Yn, ymin, ymax - 1)
    en = s["energy"][:, 0] - 1 + I(cfg["delta_energy"]) * np.asarray(energy_in, I)
    sk = jr.split(s["key"], 2)
    rf = jr.uniform(sk[:, 1], 1)[:, 0]
    rep = np.where(rf < F(cfg["reproduction_rate"]), I(1), s["reproduce"])
    a = an["active_state"] != 0
    out = dict(an)
    out["state"] = {"X_pos": np.where(a, Xn, X).astype(I)[:, None], "Y_pos": np.where(a, Yn, Y).astype(I)[:, None],
                    "energy": np.where(a, en, s["energy"][:, 0]).astype(I)[:, None],
                    "reproduce": np.where(a, rep, s["reproduce"]).astype(I),
                    "key": np.where(a[:, None], sk[:, 0], s["key"]).astype(np.uint32)}
    out["policy_key"] = np.where(a[:, None], pk[:, 1], an["policy_key"]).astype(np.uint32)
    out["age"] = np.where(a, an["age"] + F(1.0), an["age"]).astype(F)
    return out


def animal_remove(params, an, i):          # sim.ipynb:124-133
    row = am.tree_row(an, i)
    st = dict(row["state"])
    st.update(X_pos=np.array([-1], I), Y_pos=np.array([-1], I), energy=np.array([-1], I), reproduce=I(0))
    row["state"] = st
    row["age"] = F(0.0)
    row["active_state"] = False
    return row


def animal_add(params, an, idx, key):       # sim.ipynb:135-156
    copy_ids, na = params["copy_ids"], int(params["num_active_agents"])
    new = am.tree_row(an, idx)
    src = am.tree_row(an, copy_ids[min(max(idx - na, 0), len(copy_ids) - 1)])
    st = dict(new["state"])
    st.update(X_pos=src["state"]["X_pos"], Y_pos=src["state"]["Y_pos"],
              energy=(src["state"]["energy"].astype(F) / F(2)), reproduce=I(0))   # float -> int32 cast truncates
    new["state"] = st
    new["age"] = F(0.0)
    new["active_state"] = True
    return new, key


def animal_set(params, an, i, key):         # sim.ipynb:158-170
    row = am.tree_row(an, i)
    st = dict(row["state"])
    st.update(energy=(row["state"]["energy"].astype(F) / F(2)), reproduce=I(0))
    row["state"] = st
    return row, key


def grass_create(cfg, uid, active, atype, keys):
    """vmapped ``Grass.create_agent`` (sim.ipynb:177-199); cfg = dict(x_max, regrowth_time)."""
    n = len(uid)
    s3 = jr.split(keys, 3)
    ht = jr.uniform(s3[:, 2], 1)[:, 0]
    full = ~(ht < F(0.5))
    cd = np.where(full, I(0), jr.randint(s3[:, 1], 1, 1, cfg["regrowth_time"])[:, 0]).astype(I)
    return {"state": {"X_pos": (uid % cfg["x_max"]).astype(I)[:, None], "Y_pos": (uid // cfg["x_max"]).astype(I)[:, None],
                      "fully_grown": full, "count_down": cd[:, None]},
            "unique_id": uid.astype(I), "agent_type": atype.astype(I), "active_state": active.astype(F),
            "age": np.zeros(n, F)}


def grass_step(cfg, energy_out, gr):
    """vmapped ``Grass.step_agent`` (sim.ipynb:202-218): no active-state check."""
    s = gr["state"]
    eaten = np.asarray(energy_out).reshape(-1) != 0
    cd = np.where(eaten, I(cfg["regrowth_time"]), s["count_down"][:, 0])
    full = np.where(eaten, False, s["fully_grown"])
    cd = np.where(full, cd, cd - 1)
    full = np.where(cd <= 0, True, full)
    out = dict(gr)
    out["state"] = {"X_pos": s["X_pos"], "Y_pos": s["Y_pos"], "fully_grown": full, "count_down": cd.astype(I)[:, None]}
    return out


def interaction(wolves, sheeps, grasses):
    """sim.ipynb:238-290 -> (wolves_energy_in i32, sheeps_energy_in i32, sheeps_eaten f32, grasses_eaten f32)."""
    wx, wy = wolves["state"]["X_pos"][:, 0], wolves["state"]["Y_pos"][:, 0]
    sx, sy = sheeps["state"]["X_pos"][:, 0], sheeps["state"]["Y_pos"][:, 0]
    gx, gy = grasses["state"]["X_pos"][:, 0], grasses["state"]["Y_pos"][:, 0]
    ws = (wx[:, None] == sx[None, :]) & (wy[:, None] == sy[None, :])
    sg = (sx[:, None] == gx[None, :]) & (sy[:, None] == gy[None, :]) & grasses["state"]["fully_grown"][None, :]
    return (ws.sum(1).astype(I), sg.sum(1).astype(I), ws.max(0, initial=False).astype(F),
            sg.max(0, initial=False).astype(F))


class Ecosystem:
    """sim.ipynb:365-517 (without recording).  Sizes are parameters (reference: 300 + 300 animals)."""

    def __init__(self, grass_regrowth_time=30, wolf_reproduction_rate=0.08, wolf_energy=30, init_wolves=50,
                 sheep_reproduction_rate=0.2, sheep_energy=5, init_sheeps=100, X_max=25, Y_max=25,
                 max_animals=300, key=None):
        key = jr.PRNGKey(0) if key is None else key
        self.key = key
        self.gcfg = {"x_max": X_max, "regrowth_time": grass_regrowth_time}
        self.scfg = {"x_max": X_max, "y_max": Y_max, "delta_energy": sheep_energy,
                     "reproduction_rate": sheep_reproduction_rate}
        self.wcfg = {"x_max": X_max, "y_max": Y_max, "delta_energy": wolf_energy,
                     "reproduction_rate": wolf_reproduction_rate}
        self.max_animals = max_animals
        k = jr.split(key, 2)
        key, gk = k[0], k[1]
        ng = X_max * Y_max
        self.grasses = am.create_agents(lambda p, *a: grass_create(self.gcfg, *a), None, ng, ng, 0, gk)
        k = jr.split(key, 2)
        key, sk = k[0], k[1]
        self.sheeps = am.create_agents(lambda p, *a: animal_create(self.scfg, *a), None, max_animals, init_sheeps, 1, sk)
        k = jr.split(key, 2)
        key, wk = k[0], k[1]
        self.wolves = am.create_agents(lambda p, *a: animal_create(self.wcfg, *a), None, max_animals, init_wolves, 2, wk)

    def _add_animals(self, an):
        n_rep, rep_idx = am.select_agents(lambda a, p: (a["state"]["reproduce"] != 0) & (a["active_state"] != 0), None, an)
        n_act = int(np.sum(an["active_state"], dtype=F))
        n_add = min(int(n_rep), self.max_animals - n_act)
        an, self.key = am.add_agents(animal_add, n_add, {"copy_ids": rep_idx, "num_active_agents": n_act}, an, self.key)
        an, self.key = am.set_agents(animal_set, n_add, {"set_ids": rep_idx}, an, self.key)
        an, _ = am.sort_agents(an["state"]["reproduce"], False, an)
        return an, n_act

    def step(self):
        w_in, s_in, s_eaten, g_eaten = interaction(self.wolves, self.sheeps, self.grasses)
        self.grasses = grass_step(self.gcfg, g_eaten, self.grasses)
        self.sheeps = animal_step(self.scfg, s_in, self.sheeps)
        self.wolves = animal_step(self.wcfg, w_in, self.wolves)
        n_dead, idx = am.select_agents(lambda a, p: ((a["state"]["energy"][:, 0] <= 0) | (s_eaten != 0)) &
                                       (a["active_state"] != 0), None, self.sheeps)
        self.sheeps = am.remove_agents(animal_remove, n_dead, {"remove_ids": idx}, self.sheeps)
        self.sheeps, _ = am.sort_agents(self.sheeps["active_state"], False, self.sheeps)
        n_dead, idx = am.select_agents(lambda a, p: (a["state"]["energy"][:, 0] <= 0) & (a["active_state"] != 0), None,
                                       self.wolves)
        self.wolves = am.remove_agents(animal_remove, n_dead, {"remove_ids": idx}, self.wolves)
        self.wolves, _ = am.sort_agents(self.wolves["active_state"], False, self.wolves)
        self.sheeps, ns = self._add_animals(self.sheeps)
        self.wolves, nw = self._add_animals(self.wolves)
        return ns, nw
