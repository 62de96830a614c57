"""K1 parity: ray casting kernels (through the foragax.base.space_methods-shaped API, which
calls the C ABI) vs the oracle.

The orientation predicates are discrete, so the fused kernel is compared BIT-EXACTLY: ray
directions are taken from the device ``ray_generator`` (CUDA ``sincosf``; within 2 ulp of the
oracle's correctly rounded cos/sin, checked separately) and fed to the oracle, so that both
sides see the same rays.  The end-to-end comparison (oracle's own directions) uses the
north-star tolerance of 1e-5 relative on distances.
"""
import numpy as np
import pytest
import torch

from oracle import space_methods as osm

pytestmark = pytest.mark.gpu

SPAN = np.float32(np.pi / 4.0)
LEN = 60.0


def arena_walls(rng, n_walls, size=4096.0):
    """4 boundary segments + interior segments (midpoint U(arena), length U(20,200), angle U(0,pi))."""
    w = [[0, 0, size, 0], [size, 0, size, size], [0, 0, 0, size], [0, size, size, size]][:min(4, n_walls)]
    k = n_walls - len(w)
    mx, my = rng.uniform(0, size, k), rng.uniform(0, size, k)
    ln, th = rng.uniform(20, 200, k), rng.uniform(0, np.pi, k)
    inter = np.stack([mx - ln / 2 * np.cos(th), my - ln / 2 * np.sin(th), mx + ln / 2 * np.cos(th),
                      my + ln / 2 * np.sin(th)], axis=1)
    w = np.concatenate([np.array(w, np.float64).reshape(-1, 4), inter]).astype(np.float32)
    return [np.ascontiguousarray(w[:, i]) for i in range(4)]


def agents(rng, n, size=4096.0):
    return (rng.uniform(1, size - 1, n).astype(np.float32), rng.uniform(1, size - 1, n).astype(np.float32),
            rng.uniform(-np.pi, np.pi, n).astype(np.float32))


def dev(a, cuda):
    return torch.from_numpy(np.ascontiguousarray(a)).to(cuda)


def oracle_with_device_dirs(x, y, ang, n_rays, walls, cuda, active=None):
    """Oracle reduction over walls using the DEVICE's ray directions (bit-exact comparison)."""
    import foragax_b200.base.space_methods as sm
    sm.RESOLUTION = n_rays
    rays = sm.ray_generator((dev(x, cuda), dev(y, cuda), dev(ang, cuda)), SPAN, LEN)
    dx, dy = rays.ray_direction.x.cpu().numpy(), rays.ray_direction.y.cpu().numpy()
    n = len(x)
    dist = np.empty((n, n_rays), np.float32)
    hit = np.empty((n, n_rays), np.int32)
    nw = len(walls[0])
    for s in range(0, n, 512):
        e = min(n, s + 512)
        if nw == 0:
            dist[s:e] = LEN
            hit[s:e] = -1
            continue
        d = osm.ray_wall_sensor(x[s:e, None, None], y[s:e, None, None], dx[s:e, :, None], dy[s:e, :, None], LEN,
                                *[w[None, None, :] for w in walls])
        dist[s:e] = d.min(-1)
        hit[s:e] = np.where(dist[s:e] < np.float32(LEN), d.argmin(-1), -1)
    if active is not None:
        dist[active == 0] = 0
        hit[active == 0] = -1
    return dist, hit


def run_fused(x, y, ang, n_rays, walls, cuda, active=None):
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.space_classes import Point, Wall
    w = None if len(walls[0]) == 0 else Wall(Point(dev(walls[0], cuda), dev(walls[1], cuda)),
                                              Point(dev(walls[2], cuda), dev(walls[3], cuda)))
    dist, hit = sm.raycast_walls((dev(x, cuda), dev(y, cuda), dev(ang, cuda)), SPAN, LEN, w,
                                 active_state=None if active is None else dev(active, cuda), n_rays=n_rays)
    torch.cuda.synchronize()
    return dist.cpu().numpy(), hit.cpu().numpy()


def test_hand_cases(cuda):
    """SURVEY.md Appendix A.6, derived by hand from space_methods.py:71-143."""
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.space_classes import Point, Ray
    one = lambda v: dev(np.array([v], np.float32), cuda)  # noqa: E731
    ray = Ray(Point(one(0), one(0)), Point(one(1), one(0)), 60.0)
    walls = np.array([[30, -10, 30, 10], [70, -10, 70, 10], [10, 0, 20, 0], [0, -1, 0, 1]], np.float32)
    got = sm.ray_wall_sensor(ray, tuple(dev(walls[:, i].copy(), cuda) for i in range(4))).cpu().numpy()
    np.testing.assert_array_equal(got.reshape(-1), np.array([30, 60, 10, 0], np.float32))
    ents = np.array([[30, 0, 5, 1], [30, 6, 5, 1], [-30, 0, 5, 1], [0, 0, 5, 1], [30, 0, 5, 0]], np.float32)
    got = sm.ray_agent_sensor(ray, tuple(dev(ents[:, i].copy(), cuda) for i in range(4))).cpu().numpy()
    np.testing.assert_array_equal(got.reshape(-1), np.array([25, 60, 60, 60, 60], np.float32))


@pytest.mark.parametrize("n_rays", [1, 2, 9, 32])
def test_ray_generator(cuda, n_rays):
    import foragax_b200.base.space_methods as sm
    rng = np.random.default_rng(n_rays)
    x, y, ang = agents(rng, 4097)
    ang[:3] = [0.0, np.pi, -np.pi]
    sm.RESOLUTION = n_rays
    rays = sm.jit_ray_generator((dev(x, cuda), dev(y, cuda), dev(ang, cuda)), SPAN, LEN)
    _, _, dx, dy, _ = osm.ray_generator(x, y, ang, SPAN, LEN, n_rays)
    # CUDA sincosf: <= 2 ulp from the correctly rounded value
    np.testing.assert_allclose(rays.ray_direction.x.cpu().numpy(), dx, rtol=0, atol=2.5e-7)
    np.testing.assert_allclose(rays.ray_direction.y.cpu().numpy(), dy, rtol=0, atol=2.5e-7)
    assert float(rays.ray_direction.x.abs().max()) <= 1.0 and float(rays.ray_direction.y.abs().max()) <= 1.0


def test_elementwise_sensors_bit_exact(cuda):
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.space_classes import Point, Ray
    rng = np.random.default_rng(11)
    m, nw, ne = 3000, 257, 301
    ox, oy = rng.uniform(0, 500, m).astype(np.float32), rng.uniform(0, 500, m).astype(np.float32)
    th = rng.uniform(-np.pi, np.pi, m)
    dx, dy = np.cos(th).astype(np.float32), np.sin(th).astype(np.float32)
    walls = arena_walls(rng, nw, 500.0)
    # exact-zero orientations: axis-aligned rays from integer origins, walls on the integer grid
    ox[:500] = rng.integers(0, 50, 500) * 10
    oy[:500] = rng.integers(0, 50, 500) * 10
    dx[:500], dy[:500] = np.where(rng.random(500) < 0.5, 1, -1), 0
    dx[250:500], dy[250:500] = 0, np.where(rng.random(250) < 0.5, 1, -1)
    gw = rng.integers(0, 50, (100, 4)).astype(np.float32) * 10
    gw[:50, 3] = gw[:50, 1]   # horizontal walls
    gw[50:, 2] = gw[50:, 0]   # vertical walls
    gw[90:, 2:] = gw[90:, :2]  # degenerate zero-length walls
    for i in range(4):
        walls[i][4:104] = gw[:, i]
    ray = Ray(Point(dev(ox, cuda), dev(oy, cuda)), Point(dev(dx, cuda), dev(dy, cuda)), LEN)
    got = sm.jit_ray_wall_sensor(ray, tuple(dev(w, cuda) for w in walls)).cpu().numpy()
    ref = osm.ray_wall_sensor(ox[:, None], oy[:, None], dx[:, None], dy[:, None], LEN, *[w[None, :] for w in walls])
    np.testing.assert_array_equal(got, ref)
    assert (ref == 0).sum() > 0 and ((ref > 0) & (ref < LEN)).sum() > 1000
    ex, ey = rng.uniform(0, 500, ne).astype(np.float32), rng.uniform(0, 500, ne).astype(np.float32)
    er, ea = rng.uniform(0, 10, ne).astype(np.float32), (rng.random(ne) < 0.8).astype(np.float32)
    ex[:20], ey[:20] = ox[:20], oy[:20]  # origin inside the circle
    got = sm.jit_ray_agent_sensor(ray, (dev(ex, cuda), dev(ey, cuda), dev(er, cuda), dev(ea, cuda))).cpu().numpy()
    ref = osm.ray_agent_sensor(ox[:, None], oy[:, None], dx[:, None], dy[:, None], LEN, ex[None], ey[None], er[None],
                               ea[None])
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("n,n_rays,n_walls", [(1, 9, 4), (777, 9, 4), (2000, 32, 1000), (513, 32, 1001),
                                              (300, 32, 1003), (200, 8, 2048), (150, 4, 2049), (100, 16, 5000),
                                              (257, 1, 100), (300, 5, 64), (300, 3, 7), (300, 2, 3), (64, 32, 0)])
def test_fused_raycast_bit_exact(cuda, n, n_rays, n_walls):
    rng = np.random.default_rng(n * 31 + n_walls)
    x, y, ang = agents(rng, n)
    walls = arena_walls(rng, n_walls)
    # put some agents right next to / on the boundary walls
    k = min(n, 40)
    x[:k // 2] = rng.uniform(0, 70, k // 2)
    y[k // 2:k] = 4096 - rng.uniform(0, 70, k - k // 2)
    if n > 50:
        x[40:45] = 0.0      # exactly on the x=0 wall (c4: distance 0)
        y[45:50] = 4096.0
    ref_d, ref_h = oracle_with_device_dirs(x, y, ang, n_rays, walls, cuda)
    got_d, got_h = run_fused(x, y, ang, n_rays, walls, cuda)
    np.testing.assert_array_equal(got_d, ref_d)
    np.testing.assert_array_equal(got_h, ref_h)
    if n_walls >= 1000 and n >= 500:
        assert (ref_h >= 0).mean() > 0.02


def test_fused_raycast_dense_hits_and_degenerate_geometry(cuda):
    """A small arena where most rays hit several walls, integer-grid walls with exact-zero
    orientations, zero-length walls, and agents on wall end points."""
    rng = np.random.default_rng(5)
    n, n_rays = 1500, 32
    x = (rng.integers(0, 30, n) * 10).astype(np.float32)
    y = (rng.integers(0, 30, n) * 10).astype(np.float32)
    ang = (rng.integers(0, 8, n) * (np.pi / 4)).astype(np.float32)
    x[n // 2:] += rng.uniform(-3, 3, n - n // 2).astype(np.float32)
    gw = (rng.integers(0, 30, (600, 4)) * 10).astype(np.float32)
    gw[:200, 3] = gw[:200, 1]
    gw[200:400, 2] = gw[200:400, 0]
    gw[400:420, 2:] = gw[400:420, :2]
    walls = [np.ascontiguousarray(gw[:, i]) for i in range(4)]
    ref_d, ref_h = oracle_with_device_dirs(x, y, ang, n_rays, walls, cuda)
    got_d, got_h = run_fused(x, y, ang, n_rays, walls, cuda)
    np.testing.assert_array_equal(got_d, ref_d)
    np.testing.assert_array_equal(got_h, ref_h)
    assert (ref_h >= 0).mean() > 0.5 and (ref_d == 0).sum() > 100


def test_fused_raycast_inactive_agents_and_unaligned_walls(cuda):
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.space_classes import Point, Wall
    rng = np.random.default_rng(9)
    n, n_rays, n_walls = 1000, 32, 999
    x, y, ang = agents(rng, n)
    walls = arena_walls(rng, n_walls)
    active = (rng.random(n) < 0.7).astype(np.float32)
    ref_d, ref_h = oracle_with_device_dirs(x, y, ang, n_rays, walls, cuda, active)
    got_d, got_h = run_fused(x, y, ang, n_rays, walls, cuda, active)
    np.testing.assert_array_equal(got_d, ref_d)
    np.testing.assert_array_equal(got_h, ref_h)
    # wall arrays that are not 16-byte aligned (views at a 4-byte offset): non-TMA staging path
    big = [dev(np.concatenate([[0], w]).astype(np.float32), cuda)[1:] for w in walls]
    w = Wall(Point(big[0], big[1]), Point(big[2], big[3]))
    assert big[0].data_ptr() % 16 != 0
    # _f32 keeps contiguous views as they are, so the kernel sees the unaligned pointers
    d2, h2 = sm.raycast_walls((dev(x, cuda), dev(y, cuda), dev(ang, cuda)), SPAN, LEN, w, dev(active, cuda), n_rays)
    np.testing.assert_array_equal(d2.cpu().numpy(), ref_d)
    np.testing.assert_array_equal(h2.cpu().numpy(), ref_h)


def test_fused_raycast_end_to_end_tolerance(cuda):
    """Oracle with its OWN (correctly rounded) ray directions: distances within the north-star
    tolerance 1e-5 relative; hit indices equal except where a <=2-ulp direction change flips a
    grazing hit (bounded, reported)."""
    rng = np.random.default_rng(21)
    n, n_rays, n_walls = 4000, 32, 1000
    x, y, ang = agents(rng, n)
    walls = arena_walls(rng, n_walls)
    ref_d, ref_h = osm.raycast_walls(x, y, ang, SPAN, LEN, n_rays, *walls)
    got_d, got_h = run_fused(x, y, ang, n_rays, walls, cuda)
    same = got_h == ref_h
    assert same.mean() > 0.9995, f"hit index mismatch fraction {1 - same.mean():.2e}"
    np.testing.assert_allclose(got_d[same], ref_d[same], rtol=1e-5, atol=1e-5 * LEN)


def test_c_oracle_equals_numpy_oracle():
    """Not a GPU test per se, kept here to pin the timed CPU baseline to the checker."""
    from oracle import c_oracle
    rng = np.random.default_rng(2)
    x, y, ang = agents(rng, 300)
    walls = arena_walls(rng, 500)
    d, h = c_oracle.raycast_walls(x, y, ang, None, 32, SPAN, LEN, *walls)
    d2, h2 = osm.raycast_walls(x, y, ang, SPAN, LEN, 32, *walls)
    np.testing.assert_array_equal(d, d2)
    np.testing.assert_array_equal(h, h2)


@pytest.mark.parametrize("n_walls", [0, 1, 4, 100, 1000])
def test_check_collision(cuda, n_walls):
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.space_classes import Point, Wall
    rng = np.random.default_rng(n_walls)
    n = 5000
    walls = arena_walls(rng, n_walls, 500.0)
    ox, oy = rng.uniform(-5, 505, n).astype(np.float32), rng.uniform(-5, 505, n).astype(np.float32)
    nx = (ox + rng.uniform(-30, 30, n)).astype(np.float32)
    ny = (oy + rng.uniform(-30, 30, n)).astype(np.float32)
    ox[:100], oy[:100] = np.round(ox[:100] / 10) * 10, 0.0   # moves along / touching the y=0 wall
    ny[:100] = 0.0
    w = None if n_walls == 0 else Wall(Point(dev(walls[0], cuda), dev(walls[1], cuda)),
                                       Point(dev(walls[2], cuda), dev(walls[3], cuda)))
    got = sm.jit_check_collision(w, Point(dev(ox, cuda), dev(oy, cuda)), Point(dev(nx, cuda), dev(ny, cuda)))
    ref = osm.check_collision(*walls, ox, oy, nx, ny) if n_walls else np.zeros(n, bool)
    np.testing.assert_array_equal(got.cpu().numpy(), ref)
    if n_walls >= 100:
        assert ref.any() and not ref.all()


# ---- the spatially culled kernel (fgx_raycast_walls_grid) must equal the all-pairs kernel bit for bit ------------
def _both_paths(x, y, ang, n_rays, walls, cuda, bounds, active=None):
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.space_classes import Point, Wall
    w = Wall(Point(dev(walls[0], cuda), dev(walls[1], cuda)), Point(dev(walls[2], cuda), dev(walls[3], cuda)))
    pos = (dev(x, cuda), dev(y, cuda), dev(ang, cuda))
    act = None if active is None else dev(active, cuda)
    d0, h0 = sm.raycast_walls(pos, SPAN, LEN, w, active_state=act, n_rays=n_rays)
    d1, h1 = sm.raycast_walls(pos, SPAN, LEN, w, active_state=act, n_rays=n_rays, bounds=bounds)
    torch.cuda.synchronize()
    return d0.cpu().numpy(), h0.cpu().numpy(), d1.cpu().numpy(), h1.cpu().numpy()


@pytest.mark.parametrize("n,n_rays,n_walls,size", [(200_000, 32, 1000, 4096.0), (50_000, 9, 300, 1500.0),
                                                   (30_000, 33, 64, 400.0), (20_000, 1, 2000, 4096.0)])
def test_grid_cull_equals_all_pairs(cuda, n, n_rays, n_walls, size):
    rng = np.random.default_rng(n_walls + n_rays)
    walls = arena_walls(rng, n_walls, size)
    x, y, ang = agents(rng, n, size)
    act = (rng.random(n) < 0.9).astype(np.float32)
    d0, h0, d1, h1 = _both_paths(x, y, ang, n_rays, walls, cuda, (0.0, size, 0.0, size), act)
    np.testing.assert_array_equal(d1, d0)
    np.testing.assert_array_equal(h1, h0)
    assert (h0 >= 0).mean() > 0.01


def test_grid_cull_adversarial_geometry(cuda):
    """Agents ON the infinite lines of far walls looking along them (the float32 sign noise of the orientation
    tests can make such a wall 'hit'), agents on walls and wall endpoints, axis-aligned rays along the boundary
    walls, zero-length and duplicate walls, walls and agents outside the bounds: the cull must reproduce the
    all-pairs result exactly, whatever it is."""
    rng = np.random.default_rng(5)
    size = 2048.0
    walls = arena_walls(rng, 400, size)
    wx1, wy1, wx2, wy2 = [w.copy() for w in walls]
    wx2[10], wy2[10] = wx1[10], wy1[10]                       # zero-length wall
    wx1[11], wy1[11], wx2[11], wy2[11] = wx1[12], wy1[12], wx2[12], wy2[12]   # duplicate
    wx1[13], wx2[13] = -500.0, -300.0                          # outside the bounds
    wx1[14], wy1[14], wx2[14], wy2[14] = 100.0, 100.0, 1900.0, 1900.0         # long diagonal
    wx1[15], wy1[15], wx2[15], wy2[15] = 50.0, 700.0, 2000.0, 700.0           # long horizontal
    walls = [wx1, wy1, wx2, wy2]
    xs, ys, angs = [], [], []
    for w in range(0, 400):
        ax, ay, bx, by = wx1[w], wy1[w], wx2[w], wy2[w]
        th = np.arctan2(by - ay, bx - ax)
        for t in (-30.0, -9.0, -2.0, -0.4, 0.0, 0.5, 1.0, 1.3, 3.0, 12.0, 25.0):
            px, py = np.float32(ax + t * (bx - ax)), np.float32(ay + t * (by - ay))
            for da in (0.0, np.pi, SPAN, -SPAN, 1e-4, np.pi / 2):
                xs.append(px), ys.append(py), angs.append(np.float32(th + da))
    # axis-aligned positions and headings along the boundary walls
    for v in (0.0, 1.0, 60.0, 700.0, size):
        for u in np.linspace(-100.0, size + 100.0, 23):
            for a in (0.0, np.pi / 2, np.pi, -np.pi / 2):
                xs.append(np.float32(u)), ys.append(np.float32(v)), angs.append(np.float32(a))
                xs.append(np.float32(v)), ys.append(np.float32(u)), angs.append(np.float32(a))
    x, y, ang = np.array(xs, np.float32), np.array(ys, np.float32), np.array(angs, np.float32)
    for n_rays in (9, 32):
        d0, h0, d1, h1 = _both_paths(x, y, ang, n_rays, walls, cuda, (0.0, size, 0.0, size))
        np.testing.assert_array_equal(d1, d0)
        np.testing.assert_array_equal(h1, h0)
    assert (h0 >= 0).mean() > 0.2


def test_grid_cull_vs_oracle_and_edge_sizes(cuda):
    rng = np.random.default_rng(8)
    size = 600.0
    walls = arena_walls(rng, 120, size)
    x, y, ang = agents(rng, 3000, size)
    ref_d, ref_h = oracle_with_device_dirs(x, y, ang, 16, walls, cuda)
    _, _, d1, h1 = _both_paths(x, y, ang, 16, walls, cuda, (0.0, size, 0.0, size))
    np.testing.assert_array_equal(d1, ref_d)
    np.testing.assert_array_equal(h1, ref_h)
    # bounds that do not cover the agents or the walls at all: everything takes the all-walls loop
    d0, h0, d2, h2 = _both_paths(x, y, ang, 16, walls, cuda, (5000.0, 5100.0, 5000.0, 5100.0))
    np.testing.assert_array_equal(d2, d0)
    np.testing.assert_array_equal(h2, h0)


def test_grid_cache_follows_the_walls(cuda):
    """The wall lists are cached per Wall object: another Wall object, other bounds or another ray length rebuild;
    an in-place edit of the coordinates needs invalidate_wall_grid()."""
    import foragax_b200.base.space_methods as sm
    from foragax_b200.base.space_classes import Point, Wall
    rng = np.random.default_rng(21)
    size = 800.0
    x, y, ang = agents(rng, 5000, size)
    pos = (dev(x, cuda), dev(y, cuda), dev(ang, cuda))
    bounds = (0.0, size, 0.0, size)

    def wall_obj(seed):
        w = arena_walls(np.random.default_rng(seed), 100, size)
        return Wall(Point(dev(w[0], cuda), dev(w[1], cuda)), Point(dev(w[2], cuda), dev(w[3], cuda)))

    def check(walls, length=LEN, b=bounds):
        d0, h0 = sm.raycast_walls(pos, SPAN, length, walls, n_rays=16)
        d1, h1 = sm.raycast_walls(pos, SPAN, length, walls, n_rays=16, bounds=b)
        assert torch.equal(d0, d1) and torch.equal(h0, h1)

    wa, wb = wall_obj(1), wall_obj(2)
    for _ in range(2):      # second call reuses the cached grid
        check(wa)
    check(wb)               # different Wall object
    check(wa, length=35.0)  # different ray length
    check(wa, b=(0.0, size + 50.0, -20.0, size))  # different bounds
    wa.p1.x.add_(37.0)      # in-place edit: the cache cannot see it ...
    sm.invalidate_wall_grid(wa)  # ... so the caller says so
    check(wa)
