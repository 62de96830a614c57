"""Ray casting / wall detection -- the B200 mirror of ``foragax/base/space_methods.py:10-156``.

Same names as the reference (``check_collision, wall_generator, RESOLUTION, ray_generator,
ray_agent_sensor, ray_wall_sensor, create_space`` and their ``jit_*`` twins).  The
reference functions handle ONE ray / ONE entity and are ``jax.vmap``-ed by the caller
(``EcoEvoJax/sim.py:444,456,467``); here the same functions take batches -- tensors with
leading ray / entity axes -- and run as CUDA kernels through the C ABI.  The fused
agents x rays x walls hot loop is :func:`raycast_walls`.
"""
import weakref

import numpy as np
import torch

from .. import _lib as L
from .space_classes import NetworkSpace, Point, Ray, Space, Wall  # noqa: F401

RESOLUTION = 9  # rays per agent; module-level knob as in the reference (space_methods.py:55)


def _f32(x, dev=None):
    if isinstance(x, torch.Tensor):
        t = x.to(dtype=torch.float32)
        if not t.is_cuda:
            t = t.to(L.device())
        return t.reshape(-1) if t.is_contiguous() else t.contiguous().reshape(-1)
    return torch.as_tensor(np.asarray(x, dtype=np.float32).reshape(-1), device=dev or L.device())


def _wall_ptrs(walls):
    if walls is None:
        return None, None, None, None, 0, ()
    x1, y1, x2, y2 = _f32(walls.p1.x), _f32(walls.p1.y), _f32(walls.p2.x), _f32(walls.p2.y)
    return L.ptr(x1), L.ptr(y1), L.ptr(x2), L.ptr(y2), x1.shape[0], (x1, y1, x2, y2)


# space_methods.py:10-43
def check_collision(walls, old_position, new_position):
    """Batched: ``old_position`` / ``new_position`` are Points of n moves -> bool[n]
    (a single move returns a 0-d tensor, like the reference's ``jnp.any``)."""
    ox, oy, nx, ny = _f32(old_position.x), _f32(old_position.y), _f32(new_position.x), _f32(new_position.y)
    n = ox.shape[0]
    out = torch.empty(n, dtype=torch.uint8, device=ox.device)
    w1, w2, w3, w4, nw, keep = _wall_ptrs(walls)
    L.check(L.lib.fgx_check_collision(w1, w2, w3, w4, nw, L.ptr(ox), L.ptr(oy), L.ptr(nx), L.ptr(ny), n, L.ptr(out),
                                      L.stream()))
    out = out.view(torch.bool)
    return out.reshape(()) if n == 1 and not (isinstance(old_position.x, torch.Tensor) and old_position.x.dim()) else out


jit_check_collision = check_collision


# space_methods.py:45-53
def wall_generator(wall_array):
    w = np.asarray(wall_array.detach().cpu().numpy() if isinstance(wall_array, torch.Tensor) else wall_array,
                   dtype=np.float32).reshape(-1, 2, 2)
    dev = L.device()
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=dev)  # noqa: E731
    return Point(t(w[:, 0, 0]), t(w[:, 0, 1])), Point(t(w[:, 1, 0]), t(w[:, 1, 1]))


# space_methods.py:56-69
def ray_generator(agent_pos, ray_span, ray_length):
    """``agent_pos = (x, y, angle)``, each ``[n]`` or ``[n,1]`` (or ``[1]`` for one agent).
    Returns ``Ray`` with ``ray_origin`` Point [n], ``ray_direction`` Point [n, RESOLUTION]."""
    x, y, ang = (_f32(a) for a in agent_pos)
    n = ang.shape[0]
    dx = torch.empty((n, RESOLUTION), dtype=torch.float32, device=ang.device)
    dy = torch.empty_like(dx)
    L.check(L.lib.fgx_ray_generator(L.ptr(ang), n, RESOLUTION, float(ray_span), L.ptr(dx), L.ptr(dy), L.stream()))
    return Ray(Point(x, y), Point(dx, dy), float(ray_length))


jit_ray_generator = ray_generator


def _flat_rays(ray):
    dx, dy = ray.ray_direction.x, ray.ray_direction.y
    lead = tuple(dx.shape)
    ox, oy = ray.ray_origin.x, ray.ray_origin.y
    if dx.dim() == 2 and ox.numel() == dx.shape[0]:
        ox = ox.reshape(-1, 1).expand_as(dx)
        oy = oy.reshape(-1, 1).expand_as(dy)
    return _f32(ox), _f32(oy), _f32(dx), _f32(dy), lead


# space_methods.py:71-94
def ray_agent_sensor(ray, data):
    """Batched: rays [...], entities ``data = (x, y, rad, active)`` each [E] -> float32[..., E]."""
    ox, oy, dx, dy, lead = _flat_rays(ray)
    ex, ey, er, ea = (_f32(d) for d in data)
    m, e = dx.shape[0], ex.shape[0]
    out = torch.empty((m, e), dtype=torch.float32, device=dx.device)
    L.check(L.lib.fgx_ray_agent_sensor(L.ptr(ox), L.ptr(oy), L.ptr(dx), L.ptr(dy), m, float(ray.ray_length),
                                       L.ptr(ex), L.ptr(ey), L.ptr(er), L.ptr(ea), e, L.ptr(out), L.stream()))
    return out.reshape(lead + (e,))


jit_ray_agent_sensor = ray_agent_sensor


# space_methods.py:96-143
def ray_wall_sensor(ray, data):
    """Batched: rays [...], walls ``data = (p1x, p1y, p2x, p2y)`` each [W] -> float32[..., W]."""
    ox, oy, dx, dy, lead = _flat_rays(ray)
    w1, w2, w3, w4 = (_f32(d) for d in data)
    m, w = dx.shape[0], w1.shape[0]
    out = torch.empty((m, w), dtype=torch.float32, device=dx.device)
    L.check(L.lib.fgx_ray_wall_sensor(L.ptr(ox), L.ptr(oy), L.ptr(dx), L.ptr(dy), m, float(ray.ray_length),
                                      L.ptr(w1), L.ptr(w2), L.ptr(w3), L.ptr(w4), w, L.ptr(out), L.stream()))
    return out.reshape(lead + (w,))


jit_ray_wall_sensor = ray_wall_sensor


# space_methods.py:146-156
def create_space(x_min, x_max, y_min, y_max, torous, wall_array):
    if wall_array is None:
        return Space(x_min, x_max, y_min, y_max, torous, None)
    wall_begin_points, wall_end_points = wall_generator(wall_array)
    return Space(x_min, x_max, y_min, y_max, torous, Wall(wall_begin_points, wall_end_points))


GRID_MIN_WALLS = 32  # below this the all-pairs kernel is as fast as building the wall grid
_wall_grids = weakref.WeakKeyDictionary()  # Wall -> (key, scratch holding its grid): walls are static (create_space)


class _GridTicket:
    """One use of a wall grid: ``reuse`` tells the native call whether the scratch already holds the lists;
    :meth:`done` is called AFTER the native call returned FGX_OK -- only then does the cache learn that the grid
    exists (a call that raised, e.g. on a bad argument before the build, must not leave a cache entry pointing
    at uninitialised scratch) -- and records the build on the launching stream so that a later call on another
    stream waits for it."""
    __slots__ = ("bounds", "scratch", "reuse", "_walls", "_key")

    def __init__(self, bounds, scratch, reuse, walls, key):
        self.bounds, self.scratch, self.reuse, self._walls, self._key = bounds, scratch, reuse, walls, key

    def done(self):
        if not self.reuse:
            ev = None
            if not torch.cuda.is_current_stream_capturing():  # a captured build is ordered by the graph itself
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream())
            _wall_grids[self._walls] = (self._key, self.scratch, ev, L.stream())


def wall_grid(walls, bounds, ray_length):
    """A :class:`_GridTicket` for ``fgx_raycast_walls_grid`` / ``fgx_forager_sense_step``: the wall lists are built
    by the first successful call (``reuse == 0``) and cached per Wall object."""
    w1, w2, w3, w4, nw, keep = _wall_ptrs(walls)
    b = [float(v) for v in bounds]
    key = (tuple(b), float(ray_length), nw, w1, w2, w3, w4, keep[0].device.index if keep else None)
    ent = _wall_grids.get(walls)
    if ent is not None and ent[0] == key:
        if ent[2] is not None and ent[3] != L.stream() and not torch.cuda.is_current_stream_capturing():
            torch.cuda.current_stream().wait_event(ent[2])  # built on another stream: order this one after it
        return _GridTicket(b, ent[1], 1, walls, key)
    scr = L.scratch(L.lib.fgx_raycast_walls_grid_scratch_bytes(nw, b[0], b[1], b[2], b[3], float(ray_length)))
    return _GridTicket(b, scr, 0, walls, key)


def invalidate_wall_grid(walls):
    """Forget the cached grid of ``walls`` (only needed after mutating its coordinate tensors IN PLACE; new
    tensors, bounds or ray lengths are detected)."""
    _wall_grids.pop(walls, None)



def raycast_walls(agent_pos, ray_span, ray_length, walls, active_state=None, n_rays=None, want_hit=True,
                  out=None, bounds=None):
    """The fused hot loop: ``ray_generator`` + ``ray_wall_sensor`` over agents x rays x walls,
    reduced per ray to (min distance, first wall index attaining it or -1) -- what
    ``EcoEvoJax/sim.py:418-457`` computes for wall entities.  Returns
    ``(dist float32[n,R], hit int32[n,R] | None)``; inactive agents get zeros / -1.

    ``bounds = (x_min, x_max, y_min, y_max)`` (the ``Space``'s extent) selects the spatially culled kernel
    when there are enough walls for it to pay: per grid cell only the walls that can matter are tested.
    The result is bit-identical to the all-pairs evaluation (``bounds=None``), agents outside the bounds
    included."""
    x, y, ang = (_f32(a) for a in agent_pos)
    n = x.shape[0]
    r = int(RESOLUTION if n_rays is None else n_rays)
    act = None if active_state is None else _f32(active_state)
    if out is None:
        dist = torch.empty((n, r), dtype=torch.float32, device=x.device)
        hit = torch.empty((n, r), dtype=torch.int32, device=x.device) if want_hit else None
    else:
        dist, hit = out
    w1, w2, w3, w4, nw, keep = _wall_ptrs(walls)
    if bounds is not None and nw >= GRID_MIN_WALLS:
        tk = wall_grid(walls, bounds, ray_length)
        b, scr = tk.bounds, tk.scratch
        L.check(L.lib.fgx_raycast_walls_grid(L.ptr(x), L.ptr(y), L.ptr(ang), L.ptr(act), n, r, float(ray_span),
                                             float(ray_length), w1, w2, w3, w4, nw, b[0], b[1], b[2], b[3],
                                             L.ptr(dist), L.ptr(hit), L.ptr(scr), scr.numel(), tk.reuse, L.stream()))
        tk.done()
        return dist, hit
    L.check(L.lib.fgx_raycast_walls(L.ptr(x), L.ptr(y), L.ptr(ang), L.ptr(act), n, r, float(ray_span),
                                    float(ray_length), w1, w2, w3, w4, nw, L.ptr(dist), L.ptr(hit), L.stream()))
    return dist, hit


def wall_grid_listed(agent_pos, walls, bounds, ray_length, active_state=None):
    """``(listed, active)`` int64[2] device tensor: the number of (agent, wall) pairs the culled kernel evaluates
    per ray for these agents (sum of the agents' cell lists) and the number of active agents."""
    x, y = _f32(agent_pos[0]), _f32(agent_pos[1])
    act = None if active_state is None else _f32(active_state)
    w1, w2, w3, w4, nw, keep = _wall_ptrs(walls)
    tk = wall_grid(walls, bounds, ray_length)
    b, scr = tk.bounds, tk.scratch
    out = torch.empty(2, dtype=torch.int64, device=x.device)
    L.check(L.lib.fgx_raycast_walls_grid_listed(L.ptr(x), L.ptr(y), L.ptr(act), x.shape[0], w1, w2, w3, w4, nw, b[0],
                                                b[1], b[2], b[3], float(ray_length), L.ptr(scr), scr.numel(), tk.reuse,
                                                L.ptr(out), L.stream()))
    tk.done()
    return out


def raycast_agents(agent_pos, ray_span, ray_length, targets, bounds, max_rad, active_state=None, n_rays=None,
                   merge_walls=False, out=None):
    """The fused agent-agent hot loop: ``ray_generator`` + ``ray_agent_sensor`` over sensing agents x rays x
    target circles, reduced per ray to (min distance, first target index or -1) -- what the all-pairs
    composition of ``EcoEvoJax/sim.py:418-457`` computes for agent entities, found through a uniform grid.

    ``targets``: ``(x, y, rad, active)`` tensors of M targets, or one ``float32[M,4]`` tensor of interleaved
    rows (the result of ``distributed.all_gather_positions``).  ``bounds = (x_min, x_max, y_min, y_max)``.
    ``merge_walls=True``: ``out=(dist, hit)`` holds the result of :func:`raycast_walls` and receives the
    merged result in entity order [targets, walls]."""
    x, y, ang = (_f32(a) for a in agent_pos)
    n = x.shape[0]
    r = int(RESOLUTION if n_rays is None else n_rays)
    act = None if active_state is None else _f32(active_state)
    if isinstance(targets, torch.Tensor):
        t = targets.to(torch.float32).contiguous()
        m, stride = t.shape[0], 4
        base, esz = t.data_ptr(), 4
        tp = (base, base + esz, base + 2 * esz, base + 3 * esz) if m else (None,) * 4
        keep = t
    else:
        keep = tuple(_f32(a) for a in targets)
        m, stride = keep[0].shape[0], 1
        tp = tuple(L.ptr(a) for a in keep) if m else (None,) * 4
    if out is None:
        if merge_walls:
            raise L.FgxError("merge_walls needs out=(dist, hit) from raycast_walls")
        dist = torch.empty((n, r), dtype=torch.float32, device=x.device)
        hit = torch.empty((n, r), dtype=torch.int32, device=x.device)
    else:
        dist, hit = out
    b = [float(v) for v in bounds]
    nbytes = L.lib.fgx_raycast_agents_scratch_bytes(n, m, b[0], b[1], b[2], b[3], float(ray_length), float(max_rad))
    scr = L.scratch(nbytes)
    L.check(L.lib.fgx_raycast_agents(L.ptr(x), L.ptr(y), L.ptr(ang), L.ptr(act), n, r, float(ray_span),
                                     float(ray_length), tp[0], tp[1], tp[2], tp[3], stride, m, b[0], b[1], b[2], b[3],
                                     float(max_rad), 1 if merge_walls else 0, L.ptr(dist), L.ptr(hit), L.ptr(scr),
                                     nbytes, L.stream()))
    del keep
    return dist, hit
