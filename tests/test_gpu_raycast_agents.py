"""K1 part 2 parity: grid-accelerated agent-agent ray casting vs the oracle's ALL-PAIRS evaluation of
ray_agent_sensor (space_methods.py:71-94) + min / first-argmin -- bit-exact (device ray directions are
fed to the oracle, as in test_gpu_raycast.py), including ties, inactive targets, targets outside the
arena, and the merge with the wall result in entity order [agents, walls]."""
import numpy as np
import pytest
import torch

from oracle import space_methods as osm

pytestmark = pytest.mark.gpu

SPAN = np.float32(np.pi / 4.0)
LEN = 60.0


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def device_dirs(x, y, ang, n_rays, cuda, span=SPAN):
    import foragax_b200.base.space_methods as sm
    sm.RESOLUTION = n_rays
    rays = sm.ray_generator((dev(x, cuda), dev(y, cuda), dev(ang, cuda)), span, LEN)
    return rays.ray_direction.x.cpu().numpy(), rays.ray_direction.y.cpu().numpy()


def oracle_agents(x, y, dx, dy, tx, ty, tr, ta, active=None, chunk=256):
    n, R = dx.shape
    dist = np.empty((n, R), np.float32)
    hit = np.empty((n, R), np.int32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        d = osm.ray_agent_sensor(x[s:e, None, None], y[s:e, None, None], dx[s:e, :, None], dy[s:e, :, None], LEN,
                                 tx[None, None, :], ty[None, None, :], tr[None, None, :], ta[None, None, :])
        dist[s:e] = d.min(-1)
        hit[s:e] = np.where(dist[s:e] < np.float32(LEN), d.argmin(-1), -1)
    if active is not None:
        dist[active == 0] = 0
        hit[active == 0] = -1
    return dist, hit


@pytest.mark.parametrize("n,n_rays,size,rmax", [(1, 9, 100.0, 5.0), (1500, 32, 600.0, 10.0), (3000, 9, 2000.0, 8.0),
                                                (800, 16, 150.0, 10.0), (2000, 8, 60.0, 3.0), (700, 5, 900.0, 30.0)])
def test_raycast_agents_self_sensing_bit_exact(cuda, n, n_rays, size, rmax):
    """Agents sense each other (targets == the agents themselves; own circle never hits)."""
    import foragax_b200.base.space_methods as sm
    rng = np.random.default_rng(n + n_rays)
    x, y = rng.uniform(0, size, n).astype(np.float32), rng.uniform(0, size, n).astype(np.float32)
    ang = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    rad = rng.uniform(rmax / 2, rmax, n).astype(np.float32)
    act = (rng.random(n) < 0.85).astype(np.float32)
    if n > 100:   # exact duplicates (ties -> lowest index), targets exactly ray_length away
        x[10:20], y[10:20], rad[10:20] = x[0:10], y[0:10], rad[0:10]
        x[30], y[30] = x[31] + 60.0, y[31]
    dx, dy = device_dirs(x, y, ang, n_rays, cuda)
    ref_d, ref_h = oracle_agents(x, y, dx, dy, x, y, rad, act, act)
    got_d, got_h = sm.raycast_agents((dev(x, cuda), dev(y, cuda), dev(ang, cuda)), SPAN, LEN,
                                     (dev(x, cuda), dev(y, cuda), dev(rad, cuda), dev(act, cuda)),
                                     (0.0, size, 0.0, size), rmax, active_state=dev(act, cuda), n_rays=n_rays)
    np.testing.assert_array_equal(got_d.cpu().numpy(), ref_d)
    np.testing.assert_array_equal(got_h.cpu().numpy(), ref_h)
    if n >= 800:
        assert (ref_h >= 0).mean() > 0.05


@pytest.mark.parametrize("n,n_rays,size,rmax,span", [
    (900, 40, 500.0, 10.0, np.pi / 4),    # two rays per lane
    (600, 100, 400.0, 8.0, np.pi / 3),    # four rays per lane
    (1200, 1, 400.0, 10.0, 0.3),          # a single ray: degenerate cones
    (1500, 32, 500.0, 10.0, 2.5),         # non-convex fan (no cell cull against the fan, per-group front test)
    (1000, 7, 300.0, 6.0, np.pi),         # full circle, fewer rays than lanes of a group
    (6000, 32, 330.0, 6.0, np.pi / 4),    # dense: survivor lists overflow and drain several times per agent
])
def test_raycast_agents_ray_layouts_and_fans(cuda, n, n_rays, size, rmax, span):
    import foragax_b200.base.space_methods as sm
    span = np.float32(span)
    rng = np.random.default_rng(7 * n + n_rays)
    x, y = rng.uniform(0, size, n).astype(np.float32), rng.uniform(0, size, n).astype(np.float32)
    ang = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    rad = rng.uniform(rmax / 3, rmax, n).astype(np.float32)
    act = (rng.random(n) < 0.9).astype(np.float32)
    dx, dy = device_dirs(x, y, ang, n_rays, cuda, span)
    ref_d, ref_h = oracle_agents(x, y, dx, dy, x, y, rad, act, act)
    got_d, got_h = sm.raycast_agents((dev(x, cuda), dev(y, cuda), dev(ang, cuda)), span, LEN,
                                     (dev(x, cuda), dev(y, cuda), dev(rad, cuda), dev(act, cuda)),
                                     (0.0, size, 0.0, size), rmax, active_state=dev(act, cuda), n_rays=n_rays)
    np.testing.assert_array_equal(got_d.cpu().numpy(), ref_d)
    np.testing.assert_array_equal(got_h.cpu().numpy(), ref_h)
    assert (ref_h >= 0).mean() > 0.05


def test_raycast_agents_interleaved_targets_out_of_bounds_and_wall_merge(cuda):
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.space_classes import Point, Wall
    rng = np.random.default_rng(4)
    n, m, n_rays, size = 1200, 2500, 32, 700.0
    x, y = rng.uniform(1, size - 1, n).astype(np.float32), rng.uniform(1, size - 1, n).astype(np.float32)
    ang = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    tx, ty = rng.uniform(-100, size + 100, m).astype(np.float32), rng.uniform(-100, size + 100, m).astype(np.float32)
    tr, ta = rng.uniform(2, 10, m).astype(np.float32), (rng.random(m) < 0.9).astype(np.float32)
    dx, dy = device_dirs(x, y, ang, n_rays, cuda)
    ref_d, ref_h = oracle_agents(x, y, dx, dy, tx, ty, tr, ta)
    targets = dev(np.stack([tx, ty, tr, ta], axis=1), cuda)      # float32[M,4], as from all_gather_positions
    pos = (dev(x, cuda), dev(y, cuda), dev(ang, cuda))
    got_d, got_h = sm.raycast_agents(pos, SPAN, LEN, targets, (0.0, size, 0.0, size), 10.0, n_rays=n_rays)
    np.testing.assert_array_equal(got_d.cpu().numpy(), ref_d)
    np.testing.assert_array_equal(got_h.cpu().numpy(), ref_h)
    # walls, then merge in entity order [targets, walls]
    w = np.array([[0, 0, size, 0], [size, 0, size, size], [0, 0, 0, size], [0, size, size, size],
                  [100, 100, 600, 500], [50, 650, 650, 40]], np.float32)
    walls = [np.ascontiguousarray(w[:, i]) for i in range(4)]
    W = Wall(Point(dev(walls[0], cuda), dev(walls[1], cuda)), Point(dev(walls[2], cuda), dev(walls[3], cuda)))
    wd = osm.ray_wall_sensor(x[:, None, None], y[:, None, None], dx[:, :, None], dy[:, :, None], LEN,
                             *[a[None, None, :] for a in walls])
    allw = wd.min(-1)
    all_d = np.minimum(ref_d, allw)
    take_agent = (ref_h >= 0) & (ref_d <= allw)
    all_h = np.where(take_agent, ref_h, np.where(allw < np.float32(LEN), m + wd.argmin(-1), -1)).astype(np.int32)
    out = sm.raycast_walls(pos, SPAN, LEN, W, n_rays=n_rays)
    d2, h2 = sm.raycast_agents(pos, SPAN, LEN, targets, (0.0, size, 0.0, size), 10.0, n_rays=n_rays,
                               merge_walls=True, out=out)
    np.testing.assert_array_equal(d2.cpu().numpy(), all_d)
    np.testing.assert_array_equal(h2.cpu().numpy(), all_h)
    assert (all_h >= m).sum() > 100 and ((all_h >= 0) & (all_h < m)).sum() > 1000


def test_raycast_agents_no_targets(cuda):
    import foragax_b200.base.space_methods as sm
    z = torch.zeros(0, device=cuda)
    x = torch.rand(50, device=cuda) * 100
    d, h = sm.raycast_agents((x, x, x), SPAN, LEN, (z, z, z, z), (0.0, 100.0, 0.0, 100.0), 5.0, n_rays=9)
    assert bool((d == LEN).all()) and bool((h == -1).all())
