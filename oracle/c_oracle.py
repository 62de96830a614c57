"""ctypes loader of the C oracle (``oracle/c/fgx_oracle.c`` -> ``oracle/_build/libfgx_oracle.so``).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``): the multi-threaded CPU baseline timed by
``bench.py`` and a second checker for the geometry.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libfgx_oracle.so")


def build():
    import subprocess
    os.makedirs(os.path.join(_HERE, "_build"), exist_ok=True)
    subprocess.check_call(["gcc", "-O3", "-mavx2", "-mno-fma", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
                           "-fno-fast-math", "-o", LIB, os.path.join(_HERE, "c", "fgx_oracle.c"), "-lm"])


def load():
    if not os.path.exists(LIB):
        build()
    lib = C.CDLL(LIB)
    lib.fgx_oracle_threads.restype = C.c_int
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def raycast_walls(x, y, ang, active, n_rays, span, length, w1x, w1y, w2x, w2y, want_hit=True):
    lib = load()
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32).reshape(-1)  # noqa: E731
    x, y, ang, active, w1x, w1y, w2x, w2y = (f(a) for a in (x, y, ang, active, w1x, w1y, w2x, w2y))
    n = x.shape[0]
    dist = np.empty((n, n_rays), np.float32)
    hit = np.empty((n, n_rays), np.int32) if want_hit else None
    lib.fgx_oracle_raycast_walls(_p(x), _p(y), _p(ang), _p(active), C.c_int64(n), C.c_int(n_rays), C.c_float(span),
                                 C.c_float(length), _p(w1x), _p(w1y), _p(w2x), _p(w2y), C.c_int64(w1x.shape[0]),
                                 _p(dist), _p(hit))
    return dist, hit


def dice_step(sides, active, draw, key, age):
    """In place on contiguous arrays: draw int32[n], key uint32[n,2], age float32[n]."""
    lib = load()
    n = active.shape[0]
    lib.fgx_oracle_dice_step(C.c_int64(n), C.c_int(sides), _p(active), _p(draw), _p(key), _p(age))


def threads():
    return load().fgx_oracle_threads()


def set_threads(n):
    """Size of the OpenMP pool (torchrun exports OMP_NUM_THREADS=1 to its workers: bench.py's CPU arm sets it back)."""
    load().fgx_oracle_set_threads(C.c_int(int(n)))


def forager_step(cfg, obs, energy_in, fg):
    """In place on the oracle's forager dict (oracle/foraging.py layout); sim.py variant only."""
    lib = load()
    st = [np.ascontiguousarray(fg["state"][k].reshape(-1)) for k in
          ("X", "X_vel", "X_acc", "Y", "Y_vel", "Y_acc", "ang", "energy", "rad", "time_above_rep_thresh",
           "time_below_die_thresh")]
    arr = (C.c_void_p * 11)(*[_p(a) for a in st])
    pol = fg["policy"]
    n = len(fg["unique_id"])
    f = C.c_float
    obs = np.ascontiguousarray(obs, np.float32)
    energy_in = np.ascontiguousarray(energy_in, np.float32)
    lib.fgx_oracle_forager_step(C.c_int64(n), C.c_int(cfg.n_neurons), C.c_int(cfg.n_obs), C.c_int(cfg.n_act),
                                f(cfg.dt), f(cfg.action_scale), f(cfg.x_min), f(cfg.x_max), f(cfg.y_min),
                                f(cfg.y_max), f(cfg.repr_thresh), f(cfg.die_thresh), f(cfg.energy_min),
                                f(cfg.energy_max), _p(fg["active_state"]), _p(obs), _p(energy_in), arr, _p(pol["J"]),
                                _p(pol["tau"]), _p(pol["E"]), _p(pol["B"]), _p(pol["D"]), _p(pol["Z"]), _p(fg["age"]))
    for k, a in zip(("X", "X_vel", "X_acc", "Y", "Y_vel", "Y_acc", "ang", "energy", "rad", "time_above_rep_thresh",
                     "time_below_die_thresh"), st):
        fg["state"][k] = a.reshape(-1, 1)
    return fg
