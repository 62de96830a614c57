"""NumPy restatement of ``foragax/base/space_methods.py:10-156``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  PARITY UNPINNED: the
reference holds no tests or golden vectors for these functions and cannot run
here (no jax); they are restated from source, operation by operation in
float32 without FMA contraction, and checked on hand-derived cases
(``tests/test_oracle_geometry.py``).

Vectorised form: every function broadcasts over leading axes, which is what
the reference's ``jax.vmap`` nests (agents x rays x entities,
``examples/.../EcoEvoJax/sim.py:444,456,467``) compute.
"""
import numpy as np

F = np.float32


def _f(x):
    return np.asarray(x, dtype=F)


def linspace(start, stop, num):
    """``jnp.linspace`` float32: ``start*(1-s_i) + stop*s_i``, last element = stop.

    start/stop: [...]; returns [..., num]."""
    start = _f(start)[..., None]
    stop = _f(stop)[..., None]
    if num == 1:
        return start.copy()
    div = F(num - 1)
    step = (np.arange(num - 1, dtype=F) / div).astype(F)
    out = ((start * (F(1.0) - step)).astype(F) + (stop * step).astype(F)).astype(F)
    return np.concatenate([out, np.broadcast_to(stop, out.shape[:-1] + (1,))], axis=-1).astype(F)


def ray_generator(x, y, angle, ray_span, ray_length, resolution):
    """``ray_generator`` (space_methods.py:56-69).  x, y, angle: [...].

    Returns (origin_x[...], origin_y[...], dir_x[..., R], dir_y[..., R], length).
    cos/sin are evaluated in float64 and rounded (a correctly-rounded float32
    cos/sin; XLA:CPU's and CUDA's differ from it by <= 2 ulp)."""
    x, y, angle = _f(x), _f(y), _f(angle)
    span = F(ray_span)
    ang = linspace((angle - span).astype(F), (angle + span).astype(F), resolution)
    dx = np.cos(ang.astype(np.float64)).astype(F)
    dy = np.sin(ang.astype(np.float64)).astype(F)
    return x, y, dx, dy, F(ray_length)


def _orientation(px, py, qx, qy, rx, ry):
    """space_methods.py:112-113 -> class 0 (==0), 1 (>0), 2 (otherwise)."""
    val = ((qy - py).astype(F) * (rx - qx).astype(F)).astype(F) - \
          ((qx - px).astype(F) * (ry - qy).astype(F)).astype(F)
    val = val.astype(F)
    return np.where(val == 0, 0, np.where(val > 0, 1, 2)).astype(np.int32)


def _on_segment(px, py, qx, qy, rx, ry):
    """space_methods.py:105-110 (same as :13-17)."""
    return ((qx <= np.maximum(px, rx)) & (qx >= np.minimum(px, rx)) &
            (qy <= np.maximum(py, ry)) & (qy >= np.minimum(py, ry)))


def _norm(dx, dy):
    return np.sqrt(((dx * dx).astype(F) + (dy * dy).astype(F)).astype(F)).astype(F)


def _min_nan(a, b):
    """``jnp.minimum`` (NaN-propagating)."""
    return np.minimum(a, b)


def ray_wall_sensor(ox, oy, dx, dy, length, w1x, w1y, w2x, w2y):
    """``ray_wall_sensor`` (space_methods.py:96-143), broadcasting."""
    ox, oy, dx, dy = _f(ox), _f(oy), _f(dx), _f(dy)
    L = _f(length)
    w1x, w1y, w2x, w2y = _f(w1x), _f(w1y), _f(w2x), _f(w2y)
    with np.errstate(all="ignore"):
        p1x, p1y = ox, oy
        q1x = (ox + (dx * L).astype(F)).astype(F)
        q1y = (oy + (dy * L).astype(F)).astype(F)
        p2x, p2y, q2x, q2y = w1x, w1y, w2x, w2y
        o1 = _orientation(p1x, p1y, q1x, q1y, p2x, p2y)
        o2 = _orientation(p1x, p1y, q1x, q1y, q2x, q2y)
        o3 = _orientation(p2x, p2y, q2x, q2y, p1x, p1y)
        o4 = _orientation(p2x, p2y, q2x, q2y, q1x, q1y)
        shape = np.broadcast(o1, o3).shape
        Lb = np.broadcast_to(L, shape)

        # line_intercept (space_methods.py:98-103)
        ax, ay = (q1x - p1x).astype(F), (q1y - p1y).astype(F)
        d1 = ((ax * (p2y - p1y).astype(F)).astype(F) - (ay * (p2x - p1x).astype(F)).astype(F)).astype(F)
        d2 = ((ax * (q2y - p1y).astype(F)).astype(F) - (ay * (q2x - p1x).astype(F)).astype(F)).astype(F)
        r = (d1 / ((d1 - d2).astype(F) + F(1e-6)).astype(F)).astype(F)
        ix = (p2x + (r * (q2x - p2x).astype(F)).astype(F)).astype(F)
        iy = (p2y + (r * (q2y - p2y).astype(F)).astype(F)).astype(F)
        inter = _norm((ix - p1x).astype(F), (iy - p1y).astype(F))

        c1 = (o1 != o2) & (o3 != o4)
        c2 = (o1 == 0) & _on_segment(p1x, p1y, p2x, p2y, q1x, q1y)
        c3 = (o2 == 0) & _on_segment(p1x, p1y, q2x, q2y, q1x, q1y)
        c4 = (o3 == 0) & _on_segment(p2x, p2y, p1x, p1y, q2x, q2y)
        c5 = (o4 == 0) & _on_segment(p2x, p2y, q1x, q1y, q2x, q2y)
        e1 = np.where(c1, inter, Lb)
        e2 = np.where(c2, _norm((p1x - p2x).astype(F), (p1y - p2y).astype(F)), Lb)
        e3 = np.where(c3, _norm((p1x - q2x).astype(F), (p1y - q2y).astype(F)), Lb)
        e4 = np.where(c4, F(0.0), Lb)
        e5 = np.where(c5, _norm((p1x - q1x).astype(F), (p1y - q1y).astype(F)), Lb)
        out = _min_nan(e1, _min_nan(e2, _min_nan(e3, _min_nan(e4, e5))))
    return out.astype(F)


def ray_agent_sensor(ox, oy, dx, dy, length, ax, ay, arad, aactive):
    """``ray_agent_sensor`` (space_methods.py:71-94), broadcasting."""
    ox, oy, dx, dy = _f(ox), _f(oy), _f(dx), _f(dy)
    L = _f(length)
    ax, ay, arad = _f(ax), _f(ay), _f(arad)
    with np.errstate(all="ignore"):
        sx, sy = (ox - ax).astype(F), (oy - ay).astype(F)
        b = ((sx * dx).astype(F) + (sy * dy).astype(F)).astype(F)
        c = (((sx * sx).astype(F) + (sy * sy).astype(F)).astype(F) - (arad * arad).astype(F)).astype(F)
        h = ((b * b).astype(F) - c).astype(F)
        hs = np.where(h < 0, F(-1.0), np.sqrt(np.where(h < 0, F(0.0), h))).astype(F)
        Lb = np.broadcast_to(L, hs.shape)
        t = np.where(hs >= 0, ((-b).astype(F) - hs).astype(F), Lb).astype(F)
        t = np.where(t < 0, Lb, t).astype(F)
        t = _min_nan(t, Lb)
        out = np.where(np.asarray(aactive) != 0, t, Lb)
    return out.astype(F)


def check_collision(w1x, w1y, w2x, w2y, oldx, oldy, newx, newy):
    """``check_collision`` (space_methods.py:10-43): walls [W] vs moves [...] -> bool[...]."""
    w1x, w1y, w2x, w2y = (_f(a) for a in (w1x, w1y, w2x, w2y))
    p2x, p2y = _f(oldx)[..., None], _f(oldy)[..., None]
    q2x, q2y = _f(newx)[..., None], _f(newy)[..., None]
    p1x, p1y, q1x, q1y = w1x, w1y, w2x, w2y
    with np.errstate(all="ignore"):
        o1 = _orientation(p1x, p1y, q1x, q1y, p2x, p2y)
        o2 = _orientation(p1x, p1y, q1x, q1y, q2x, q2y)
        o3 = _orientation(p2x, p2y, q2x, q2y, p1x, p1y)
        o4 = _orientation(p2x, p2y, q2x, q2y, q1x, q1y)
        c1 = (o1 != o2) & (o3 != o4)
        c2 = (o1 == 0) & _on_segment(p1x, p1y, p2x, p2y, q1x, q1y)
        c3 = (o2 == 0) & _on_segment(p1x, p1y, q2x, q2y, q1x, q1y)
        c4 = (o3 == 0) & _on_segment(p2x, p2y, p1x, p1y, q2x, q2y)
        c5 = (o4 == 0) & _on_segment(p2x, p2y, q1x, q1y, q2x, q2y)
    return np.any(c1 | c2 | c3 | c4 | c5, axis=-1)


def raycast_walls(x, y, angle, ray_span, ray_length, resolution, w1x, w1y, w2x, w2y,
                  active=None, chunk=4096):
    """agents x rays x walls: (min distance f32[N,R], first-min wall index i32[N,R], -1 = no hit).

    The wall-only slice of the example-level composition
    ``sim.py:418-457`` (min / argmin over entities, ``-1`` if ``min >= L``)."""
    x, y, angle = _f(x).reshape(-1), _f(y).reshape(-1), _f(angle).reshape(-1)
    n = x.shape[0]
    dist = np.empty((n, resolution), F)
    hit = np.empty((n, resolution), np.int32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        ox, oy, dx, dy, L = ray_generator(x[s:e], y[s:e], angle[s:e], ray_span, ray_length, resolution)
        d = ray_wall_sensor(ox[:, None, None], oy[:, None, None], dx[:, :, None], dy[:, :, None], L,
                            _f(w1x)[None, None, :], _f(w1y)[None, None, :],
                            _f(w2x)[None, None, :], _f(w2y)[None, None, :])
        dist[s:e] = d.min(axis=-1)
        idx = d.argmin(axis=-1).astype(np.int32)
        hit[s:e] = np.where(dist[s:e] < L, idx, -1)
    if active is not None:
        act = np.asarray(active).reshape(-1) != 0
        dist = np.where(act[:, None], dist, F(0.0)).astype(F)
        hit = np.where(act[:, None], hit, -1).astype(np.int32)
    return dist, hit


def wall_generator(wall_array):
    """``wall_generator`` (space_methods.py:45-53): [[b,e],...] -> (x1,y1,x2,y2) float32[W]."""
    w = np.asarray(wall_array, dtype=F).reshape(-1, 2, 2)
    return w[:, 0, 0].copy(), w[:, 0, 1].copy(), w[:, 1, 0].copy(), w[:, 1, 1].copy()
