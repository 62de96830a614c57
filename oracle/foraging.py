"""NumPy restatement of the EcoEvoJax foraging example
(``examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.py:24-553,597-666`` and its
notebook twin ``sim.ipynb``), the workload of BASELINE config 2 and the source of golden
vector G2 (500 printed ticks, ``sim.ipynb`` cell output).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  ``Config.variant`` selects the constants of
``sim.py`` ("sim") or of the notebook that produced G2 ("nb"); the differences are listed in
SURVEY.md section 4.  vmapped functions are batched, ``fori_loop`` bodies stay serial.

Reference quirk reproduced on purpose: ``sensor_model`` orders entities as
[foragers, resources, walls] (sim.py:390-408) but the distance matrix it prefilters with is
ordered [resources, foragers, walls-zero-pad] (sim.py:381,416), so entity ``e`` is culled by the
distance to a *different* agent (column ``e`` of that matrix).
"""
import numpy as np

from . import agent_methods as am
from . import jax_random as jr
from . import space_methods as osm

F = np.float32


class Config:
    def __init__(self, variant="sim", **kw):
        self.variant = variant
        if variant == "sim":      # sim.py:16-19,704-715
            self.x_max, self.y_max = 1000.0, 500.0
            self.n_neurons = 40
            self.res_total, self.res_active = 6000, 5000
            self.max_age, self.time_to_die = 1000.0, 3.0
            self.repr_thresh, self.time_to_reproduce = 5.0, 0.4
            self.seed = 1
        else:                     # sim.ipynb driver cell
            self.x_max, self.y_max = 2000.0, 500.0
            self.n_neurons = 30
            self.res_total, self.res_active = 4000, 3000
            self.max_age, self.time_to_die = 50.0, 1.0
            self.repr_thresh, self.time_to_reproduce = 5.5, 0.5
            self.seed = 0
        self.x_min = self.y_min = 0.0
        self.for_total, self.for_active = 1000, 500
        self.n_obs, self.n_act = 28, 2
        self.dt = 0.1
        self.action_scale = 1.0
        self.noise = 0.02
        self.n_rays = 9
        self.ray_span = np.pi / 4.0
        self.ray_len = 60.0
        self.die_thresh, self.energy_min, self.energy_max = 0.0, -1.0, 15.0
        self.blossom_thresh = 3.0
        self.res_energy_min = 0.1
        for k, v in kw.items():
            setattr(self, k, v)

    @property
    def walls(self):
        X, Y = self.x_max, self.y_max
        w = np.array([[0, 0, X, 0], [X, 0, X, Y], [0, 0, 0, Y], [0, Y, X, Y]], F)
        return [np.ascontiguousarray(w[:, i]) for i in range(4)]


def sigmoid(x):
    x = np.asarray(x, F)
    return (F(1.0) / (F(1.0) + np.exp(-x).astype(F))).astype(F)


# ---- WilsonCowan (foragax/policy/wilson_cowan.py) ----------------------------
def wc_create(cfg, keys):
    """``create_policy`` (wilson_cowan.py:10-31) for a key batch [N,2]."""
    n, o, a = cfg.n_neurons, cfg.n_obs, cfg.n_act
    N = keys.shape[0]
    ik = jr.split(keys, 6)[:, 1:]
    return {
        "J": jr.uniform(ik[:, 0], n * n, -1, 1).reshape(N, n, n),
        "tau": jr.uniform(ik[:, 1], n, -1, 1),
        "E": jr.uniform(ik[:, 2], n * o, -1, 1).reshape(N, n, o),
        "B": jr.uniform(ik[:, 3], n, -1, 1),
        "D": jr.uniform(ik[:, 4], a * n, -1, 1).reshape(N, a, n),
        "Z": np.zeros((N, n), F),
    }


def wc_step(cfg, pol, obs):
    """``step_policy`` (wilson_cowan.py:34-51), batched: obs [N, n_obs] -> (actions [N, n_act], new Z)."""
    Z = pol["Z"]
    tz = np.tanh(Z).astype(F)
    dz = (np.einsum("nij,nj->ni", pol["J"], tz).astype(F) + np.einsum("nij,nj->ni", pol["E"], obs).astype(F)
          + pol["B"] - Z).astype(F)
    dz = (dz * (F(10.0) * sigmoid(pol["tau"])).astype(F)).astype(F)
    newZ = (Z + (F(cfg.dt) * dz).astype(F)).astype(F)
    act = (F(cfg.action_scale) * np.tanh(np.einsum("nij,nj->ni", pol["D"], np.tanh(newZ).astype(F)).astype(F))).astype(F)
    return act, newZ


# ---- Forager (sim.py:24-225) -------------------------------------------------
FSTATE = ("X", "X_vel", "X_acc", "Y", "Y_vel", "Y_acc", "ang", "energy", "rad", "time_above_rep_thresh",
          "time_below_die_thresh")


def forager_create(cfg, uid, active, atype, keys):
    N = len(uid)
    k = jr.split(keys, 2)
    pol = wc_create(cfg, k[:, 1])
    sk = jr.split(k[:, 0], 8)[:, 1:]
    xlo, xhi = cfg.x_min + (cfg.x_min + 1.0) / 10.0, cfg.x_max - cfg.x_max / 10.0
    ylo, yhi = cfg.y_min + (cfg.y_min + 1.0) / 10.0, cfg.y_max - cfg.y_max / 10.0
    X, Xv, Xa = jr.uniform(sk[:, 0], 1, xlo, xhi), jr.uniform(sk[:, 1], 1, -1, 1), jr.uniform(sk[:, 2], 1, -1, 1)
    Y, Yv, Ya = jr.uniform(sk[:, 3], 1, ylo, yhi), jr.uniform(sk[:, 4], 1, -1, 1), jr.uniform(sk[:, 5], 1, -1, 1)
    en = jr.uniform(sk[:, 6], 1, 5.0, 10.0)
    act = {"X": X, "X_vel": Xv, "X_acc": Xa, "Y": Y, "Y_vel": Yv, "Y_acc": Ya, "ang": np.arctan2(Yv, Xv).astype(F),
           "energy": en, "rad": en, "time_above_rep_thresh": np.zeros((N, 1), F),
           "time_below_die_thresh": np.zeros((N, 1), F)}
    ina = {"X": -1.0, "X_vel": 0.0, "X_acc": 0.0, "Y": -1.0, "Y_vel": 0.0, "Y_acc": 0.0, "ang": 0.0, "energy": -1.0,
           "rad": 0.0, "time_above_rep_thresh": -1.0, "time_below_die_thresh": -1.0}
    a = (active != 0)[:, None]
    st = {f: np.where(a, act[f], F(ina[f])).astype(F) for f in FSTATE}
    return {"state": st, "policy": pol, "unique_id": uid.astype(np.int32), "agent_type": atype.astype(np.int32),
            "active_state": active.astype(F), "age": np.zeros(N, F)}


def forager_step(cfg, obs, energy_in, fg):
    """vmapped ``Forager.step_agent`` (sim.py:74-136); obs [N,27], energy_in [N]."""
    s = fg["state"]
    dt = F(cfg.dt)
    full_obs = np.concatenate([obs.astype(F), s["energy"]], axis=1)
    actions, newZ = wc_step(cfg, fg["policy"], full_obs)
    Xa, Ya = actions[:, 0:1], actions[:, 1:2]
    X, Y, Xv, Yv = s["X"], s["Y"], s["X_vel"], s["Y_vel"]
    Xn, Yn = (X + (Xv * dt).astype(F)).astype(F), (Y + (Yv * dt).astype(F)).astype(F)
    Xvn, Yvn = (Xv + (Xa * dt).astype(F)).astype(F), (Yv + (Ya * dt).astype(F)).astype(F)
    bx = (Xn > F(cfg.x_max)) | (Xn < F(cfg.x_min))
    by = (Yn > F(cfg.y_max)) | (Yn < F(cfg.y_min))
    Xn, Xvn = np.where(bx, X, Xn), np.where(bx, -Xv, Xvn)
    Yn, Yvn = np.where(by, Y, Yn), np.where(by, -Yv, Yvn)
    ang = np.arctan2(Yvn, Xvn).astype(F)
    asum = (np.abs(actions[:, 0]) + np.abs(actions[:, 1])).astype(F)[:, None]
    e_in = np.asarray(energy_in, F).reshape(-1, 1)
    if cfg.variant == "sim":
        en = ((s["energy"] + e_in).astype(F) - ((F(0.3) * dt).astype(F) * asum).astype(F)).astype(F)
        en = np.clip(en, F(cfg.energy_min), F(cfg.energy_max)).astype(F)
    else:
        en = ((s["energy"] + e_in).astype(F) - (dt * asum).astype(F)).astype(F)
        en = np.minimum(en, F(cfg.energy_max)).astype(F)
    tar = np.where(en > F(cfg.repr_thresh), (s["time_above_rep_thresh"] + dt).astype(F), F(0.0)).astype(F)
    tbd = np.where(en < F(cfg.die_thresh), (s["time_below_die_thresh"] + dt).astype(F), F(0.0)).astype(F)
    new = {"X": Xn, "X_vel": Xvn, "X_acc": Xa, "Y": Yn, "Y_vel": Yvn, "Y_acc": Ya, "ang": ang, "energy": en, "rad": en,
           "time_above_rep_thresh": tar, "time_below_die_thresh": tbd}
    a = fg["active_state"] != 0
    out = dict(fg)
    out["state"] = {f: np.where(a[:, None], new[f], s[f]).astype(F) for f in FSTATE}
    pol = dict(fg["policy"])
    pol["Z"] = np.where(a[:, None], newZ, fg["policy"]["Z"]).astype(F)
    out["policy"] = pol
    out["age"] = np.where(a, (fg["age"] + dt).astype(F), fg["age"]).astype(F)
    return out


def forager_remove(cfg):
    def f(params, fg, i):                      # sim.py:139-151
        row = am.tree_row(fg, i)
        vals = {"X": -1.0, "X_vel": 0.0, "X_acc": 0.0, "Y": -1.0, "Y_vel": 0.0, "Y_acc": 0.0, "ang": -1.0,
                "energy": -1.0, "rad": 0.0, "time_above_rep_thresh": -1.0, "time_below_die_thresh": -1.0}
        row["state"] = {k: np.array([v], F) for k, v in vals.items()}
        row["policy"]["Z"] = np.zeros_like(row["policy"]["Z"])
        row["active_state"] = False
        row["age"] = F(0.0)
        return row
    return f


def forager_add(cfg):
    def f(params, fg, idx, key):               # sim.py:154-210
        good, na = params["good_ids"], int(params["num_active_foragers"])
        new = am.tree_row(fg, idx)
        src = am.tree_row(fg, good[min(max(idx - na, 0), len(good) - 1)])
        z = np.zeros(1, F)
        en = (src["state"]["energy"] / F(2.0)).astype(F)
        new["state"] = {"X": src["state"]["X"], "X_vel": z, "X_acc": z, "Y": src["state"]["Y"], "Y_vel": z, "Y_acc": z,
                        "ang": z, "energy": en, "rad": en, "time_above_rep_thresh": z, "time_below_die_thresh": z}
        if cfg.variant == "sim":
            ns = F(cfg.noise)
        else:
            ns = np.clip((F(0.1) / (src["age"] + F(0.01))).astype(F), F(0.001), F(0.1)).astype(F)
        ks = jr.split(key, 6)
        key, nk = ks[0], ks[1:]
        pol = dict(new["policy"])
        for j, name in enumerate(("J", "tau", "E", "B", "D")):
            w = src["policy"][name]
            noise = jr.normal(nk[j], w.size).reshape(w.shape)
            pol[name] = (w + (ns * noise).astype(F)).astype(F)
        new["policy"] = pol
        new["age"] = F(0.0)
        new["active_state"] = True
        return new, key
    return f


def forager_set(cfg):
    def f(params, fg, i, key):                 # sim.py:213-225
        row = am.tree_row(fg, i)
        en = (row["state"]["energy"] / F(2.0)).astype(F)
        st = dict(row["state"])
        st.update(energy=en, rad=en, time_above_rep_thresh=np.zeros(1, F))
        row["state"] = st
        return row, key
    return f


# ---- Resource (sim.py:228-355) -----------------------------------------------
def resource_create(cfg, uid, active, atype, keys):
    N = len(uid)
    k = jr.split(keys, 2)
    gr = jr.uniform(k[:, 1], 1, 0.01, 1.0)
    dr = (F(0.2) * gr).astype(F)
    s4 = jr.split(k[:, 0], 4)
    if cfg.variant == "sim":
        xlo, xhi, ylo, yhi = cfg.x_min + 10.0, cfg.x_max - 10.0, cfg.y_min + 10.0, cfg.y_max - 10.0
    else:
        xlo, xhi = cfg.x_min + (cfg.x_min + 1.0) / 10.0, cfg.x_max - cfg.x_max / 10.0
        ylo, yhi = cfg.y_min + (cfg.y_min + 1.0) / 10.0, cfg.y_max - cfg.y_max / 10.0
    X, Y = jr.uniform(s4[:, 1], 1, xlo, xhi), jr.uniform(s4[:, 2], 1, ylo, yhi)
    en = jr.uniform(s4[:, 3], 1, 3.0, 4.0)
    a = (active != 0)[:, None]
    st = {"X": np.where(a, X, F(-1.0)).astype(F), "Y": np.where(a, Y, F(-1.0)).astype(F),
          "energy": np.where(a, en, F(0.0)).astype(F), "rad": np.where(a, en, F(0.0)).astype(F),
          "blossom": np.zeros(N, bool), "key": np.where(a, s4[:, 0], k[:, 0]).astype(np.uint32),
          "time_above_blossom": np.where(a, F(0.0), F(-1.0)).astype(F)}
    return {"state": st, "params": {"growth_rate": gr, "decay_rate": dr}, "unique_id": uid.astype(np.int32),
            "agent_type": atype.astype(np.int32), "active_state": active.astype(F), "age": np.zeros(N, F)}


def resource_step(cfg, energy_out, rs):
    """vmapped ``Resource.step_agent`` (sim.py:264-303); energy_out [N]."""
    s, p = rs["state"], rs["params"]
    dt = F(cfg.dt)
    old = s["energy"]
    eaten = np.asarray(energy_out, F).reshape(-1, 1)
    growth = ((p["growth_rate"] * old).astype(F) - (p["decay_rate"] * (old * old).astype(F)).astype(F)).astype(F)
    en = ((old + (dt * growth).astype(F)).astype(F) - eaten).astype(F)
    en = np.clip(en, F(0.0), F(5.0)).astype(F)
    xmax, xmin = F(cfg.x_max), F(cfg.x_min)
    prob = sigmoid(((F(4.0) * (xmax - s["X"]).astype(F)).astype(F) / (xmax - xmin)).astype(F))
    k = jr.split(s["key"], 2)
    luck = jr.uniform(k[:, 1], 1) < prob
    change = luck[:, 0] & (en[:, 0] > F(cfg.blossom_thresh))
    blossom = np.where(s["blossom"], s["blossom"], change)
    tab = np.where(blossom[:, None], (s["time_above_blossom"] + dt).astype(F), F(0.0)).astype(F)
    a = rs["active_state"] != 0
    out = dict(rs)
    out["state"] = {"X": s["X"], "Y": s["Y"], "energy": np.where(a[:, None], en, old).astype(F),
                    "rad": np.where(a[:, None], en, s["rad"]).astype(F), "blossom": np.where(a, blossom, s["blossom"]),
                    "key": np.where(a[:, None], k[:, 0], s["key"]).astype(np.uint32),
                    "time_above_blossom": np.where(a[:, None], tab, s["time_above_blossom"]).astype(F)}
    out["age"] = np.where(a, (rs["age"] + dt).astype(F), rs["age"]).astype(F)
    return out


def resource_remove(cfg):
    def f(params, rs, i):                      # sim.py:306-313
        row = am.tree_row(rs, i)
        st = dict(row["state"])
        st.update(X=np.array([-1.0], F), Y=np.array([-1.0], F), energy=np.zeros(1, F), rad=np.zeros(1, F),
                  blossom=False, time_above_blossom=np.array([-1.0], F))
        row["state"] = st
        row["active_state"] = False
        row["age"] = F(0.0)
        return row
    return f


def resource_add(cfg):
    def f(params, rs, idx, key):               # sim.py:316-344
        good, na = params["good_ids"], int(params["num_active_resources"])
        new = am.tree_row(rs, idx)
        src = am.tree_row(rs, good[min(max(idx - na, 0), len(good) - 1)])
        ks = jr.split(key, 4)
        key, sk = ks[0], ks[1:]
        rad = src["state"]["rad"]
        lo, hi = (F(-2.0) * rad).astype(F), (F(2.0) * rad).astype(F)
        X = np.clip((src["state"]["X"] + jr.uniform(sk[0], 1, lo, hi)).astype(F), F(cfg.x_min), F(cfg.x_max)).astype(F)
        Y = np.clip((src["state"]["Y"] + jr.uniform(sk[1], 1, lo, hi)).astype(F), F(cfg.y_min), F(cfg.y_max)).astype(F)
        en = jr.uniform(sk[2], 1, 3.0, 4.0)
        new["state"] = {"X": X, "Y": Y, "energy": en, "rad": en, "blossom": False, "key": new["state"]["key"],
                        "time_above_blossom": np.zeros(1, F)}
        ks2 = jr.split(key, 2)
        key, pk = ks2[0], ks2[1]
        gr = np.clip((src["params"]["growth_rate"] + jr.uniform(pk, 1, -0.01, 0.01)).astype(F), F(0.001), F(1.0)).astype(F)
        new["params"] = {"growth_rate": gr, "decay_rate": (F(0.2) * gr).astype(F)}
        new["age"] = F(0.0)
        new["active_state"] = True
        return new, key
    return f


def resource_set(cfg):
    def f(params, rs, i, key):                 # sim.py:347-355
        row = am.tree_row(rs, i)
        en = (row["state"]["energy"] / F(2.0)).astype(F)
        st = dict(row["state"])
        st.update(energy=en, rad=en, blossom=False, time_above_blossom=np.zeros(1, F))
        row["state"] = st
        return row, key
    return f


# ---- forager_resource_interaction (sim.py:358-384) ----------------------------
def interaction(fg, rs):
    """Returns (energy_in [F], energy_out [R]); the distance matrix itself is recomputed
    where ``sensor_model`` needs it."""
    fx, fy = fg["state"]["X"][:, 0], fg["state"]["Y"][:, 0]
    rx, ry = rs["state"]["X"][:, 0], rs["state"]["Y"][:, 0]
    dx = (rx[None, :] - fx[:, None]).astype(F)
    dy = (ry[None, :] - fy[:, None]).astype(F)
    dist = np.sqrt(((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)).astype(F)
    hit = dist < (rs["state"]["rad"][:, 0] + F(0.1)).astype(F)[None, :]
    em = np.where(hit, rs["state"]["energy"][:, 0][None, :], F(0.0)).astype(F)
    return em.sum(axis=1, dtype=F), em.sum(axis=0, dtype=F)


def _pair_dist(ax, ay, bx, by):
    dx, dy = (bx - ax).astype(F), (by - ay).astype(F)
    return np.sqrt(((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)).astype(F)


# ---- sensor_model (sim.py:388-468) --------------------------------------------
def sensor_model(cfg, fg, rs, walls):
    """Returns obs float32[F, 3*n_rays]: per ray (energy, type, distance) of the first-minimum
    entity, entities ordered [foragers, resources, walls]; reproduces the misaligned prefilter."""
    nf, nr = len(fg["unique_id"]), len(rs["unique_id"])
    R, L = cfg.n_rays, F(cfg.ray_len)
    ex = np.concatenate([fg["state"]["X"][:, 0], rs["state"]["X"][:, 0]])
    ey = np.concatenate([fg["state"]["Y"][:, 0], rs["state"]["Y"][:, 0]])
    er = np.concatenate([fg["state"]["rad"][:, 0], rs["state"]["rad"][:, 0]])
    ea = np.concatenate([fg["active_state"], rs["active_state"]])
    et = np.concatenate([fg["agent_type"], rs["agent_type"], np.zeros(len(walls[0]), np.int32)]).astype(F)
    en = np.concatenate([fg["state"]["energy"][:, 0], rs["state"]["energy"][:, 0], np.zeros(len(walls[0]), F)])
    # column e of the reference's dist matrix: [resources, foragers]
    px = np.concatenate([rs["state"]["X"][:, 0], fg["state"]["X"][:, 0]])
    py = np.concatenate([rs["state"]["Y"][:, 0], fg["state"]["Y"][:, 0]])
    fx, fy, fa = fg["state"]["X"][:, 0], fg["state"]["Y"][:, 0], fg["state"]["ang"][:, 0]
    ox, oy, dx, dy, _ = osm.ray_generator(fx, fy, fa, cfg.ray_span, cfg.ray_len, R)
    out = np.zeros((nf, R, 3), F)
    nE = nf + nr
    for f in np.nonzero(fg["active_state"] != 0)[0]:
        near = _pair_dist(fx[f], fy[f], px, py) < (L + er).astype(F)
        cand = np.nonzero(near)[0]
        dw = osm.ray_wall_sensor(ox[f], oy[f], dx[f][:, None], dy[f][:, None], L, *[w[None, :] for w in walls])
        best = dw.min(axis=1)
        bidx = nE + dw.argmin(axis=1)
        if len(cand):
            da = osm.ray_agent_sensor(ox[f], oy[f], dx[f][:, None], dy[f][:, None], L, ex[cand][None, :],
                                      ey[cand][None, :], er[cand][None, :], ea[cand][None, :])
            amin = da.min(axis=1)
            aidx = cand[da.argmin(axis=1)]
            take = amin <= best           # agents come before walls in entity order: ties go to the agent
            take &= amin < L              # ... but an all-L agent row must not displace the wall argmin
            bidx = np.where(take, aidx, bidx)
            best = np.where(take, amin, best)
        hit = best < L
        # argmin over an all-L row is entity 0, but index -1 is forced when min >= L (sim.py:447)
        out[f, :, 0] = np.where(hit, en[np.where(hit, bidx, 0)], F(0.0))
        out[f, :, 1] = np.where(hit, et[np.where(hit, bidx, 0)], F(0.0))
        out[f, :, 2] = best
    return out.reshape(nf, R * 3)


# ---- Neuroevolution (sim.py:474-553) -------------------------------------------
def neuroevolution(cfg, fg, rs, key, life_old):
    t1 = lambda x: x.reshape(-1)  # noqa: E731

    def weak_or_old(a, p):
        return (a["age"] > F(cfg.max_age)) | ((a["active_state"] == 1.0) &
                                              (t1(a["state"]["time_below_die_thresh"]) > F(cfg.time_to_die)))
    n_rem, rem = am.select_agents(weak_or_old, None, fg)
    age_before = np.sum(fg["age"], dtype=F)
    fg = am.remove_agents(forager_remove(cfg), n_rem, {"remove_ids": rem}, fg)
    age_after = np.sum(fg["age"], dtype=F)
    life = F((age_before - age_after) / (F(n_rem) + F(0.01)))
    life = life_old if life == 0.0 else life

    n_rrem, rrem = am.select_agents(lambda a, p: (a["active_state"] == 1.0) &
                                    (t1(a["state"]["energy"]) < F(cfg.res_energy_min)), None, rs)
    rs = am.remove_agents(resource_remove(cfg), n_rrem, {"remove_ids": rrem}, rs)

    fg, _ = am.sort_agents(fg["state"]["time_above_rep_thresh"], False, fg)
    rs, _ = am.sort_agents(rs["state"]["time_above_blossom"], False, rs)

    n_fadd, fadd = am.select_agents(lambda a, p: (a["active_state"] == 1.0) &
                                    (t1(a["state"]["time_above_rep_thresh"]) > F(cfg.time_to_reproduce)), None, fg)
    n_fact = int(np.sum(fg["active_state"], dtype=F))
    n_fadd = min(int(n_fadd), len(fg["unique_id"]) - n_fact)
    fg, key = am.add_agents(forager_add(cfg), n_fadd, {"good_ids": fadd, "num_active_foragers": n_fact}, fg, key)
    fg, key = am.set_agents(forager_set(cfg), n_fadd, {"set_ids": fadd}, fg, key)

    n_radd, radd = am.select_agents(lambda a, p: (a["active_state"] == 1.0) & a["state"]["blossom"], None, rs)
    n_ract = int(np.sum(rs["active_state"], dtype=F))
    n_radd = min(int(n_radd), len(rs["unique_id"]) - n_ract)
    rs, key = am.add_agents(resource_add(cfg), n_radd, {"good_ids": radd, "num_active_resources": n_ract}, rs, key)
    rs, key = am.set_agents(resource_set(cfg), n_radd, {"set_ids": radd}, rs, key)
    return fg, rs, key, n_fact, n_ract, life


# ---- Foraging_world (sim.py:597-666) --------------------------------------------
class World:
    def __init__(self, cfg, key=None):
        self.cfg = cfg
        key = jr.PRNGKey(cfg.seed) if key is None else key
        k = jr.split(key, 2)
        self.key, sub = k[0], k[1]
        self.fg = am.create_agents(lambda p, *a: forager_create(cfg, *a), None, cfg.for_total, cfg.for_active, 1, sub)
        k = jr.split(self.key, 2)
        self.key, sub = k[0], k[1]
        self.rs = am.create_agents(lambda p, *a: resource_create(cfg, *a), None, cfg.res_total, cfg.res_active, 2, sub)
        self.life = F(50.0)
        self.walls = cfg.walls

    def step(self):
        cfg = self.cfg
        e_in, e_out = interaction(self.fg, self.rs)
        obs = sensor_model(cfg, self.fg, self.rs, self.walls)
        self.fg = forager_step(cfg, obs, e_in, self.fg)
        self.rs = resource_step(cfg, e_out, self.rs)
        self.fg, self.rs, self.key, nf, nr, self.life = neuroevolution(cfg, self.fg, self.rs, self.key, self.life)
        return nf, nr, float(self.life)
