"""Oracle geometry: hand-derived cases (from space_methods.py:71-143)
and agreement of the C restatement (the timed CPU baseline) with the NumPy one.  The
reference holds no golden vectors for these functions: parity unpinned (see oracle/__init__.py)."""
import numpy as np

from oracle import space_methods as osm


def test_ray_wall_hand_cases():
    f = lambda *w: float(osm.ray_wall_sensor(0, 0, 1, 0, 60, *w))  # noqa: E731
    assert f(30, -10, 30, 10) == 30.0            # proper crossing: r = -600/(-1200+1e-6)
    assert f(70, -10, 70, 10) == 60.0            # beyond the ray
    assert f(10, 0, 20, 0) == 10.0               # collinear: min(|p1-p2|, |p1-q2|)
    assert f(0, -1, 0, 1) == 0.0                 # origin on the wall (c4)
    assert f(-5, 5, -5, -5) == 60.0              # behind the origin


def test_ray_agent_hand_cases():
    g = lambda *a: float(osm.ray_agent_sensor(0, 0, 1, 0, 60, *a))  # noqa: E731
    assert g(30, 0, 5, 1) == 25.0
    assert g(30, 6, 5, 1) == 60.0    # h < 0
    assert g(-30, 0, 5, 1) == 60.0   # t < 0
    assert g(0, 0, 5, 1) == 60.0     # origin inside the circle
    assert g(30, 0, 5, 0) == 60.0    # inactive target


def test_linspace_and_ray_generator():
    np.testing.assert_array_equal(osm.linspace(np.float32([0.0]), np.float32([1.0]), 5),
                                  np.array([[0, 0.25, 0.5, 0.75, 1.0]], np.float32))
    x, y, dx, dy, L = osm.ray_generator(np.float32([1.0]), np.float32([2.0]), np.float32([0.0]), np.pi / 4, 60.0, 9)
    assert dx.shape == (1, 9) and dy[0, 4] == 0.0 and dx[0, 4] == 1.0
    np.testing.assert_allclose(dy[0, 0], -np.sin(np.pi / 4), rtol=1e-6)
    np.testing.assert_allclose(dy[0, 8], np.sin(np.pi / 4), rtol=1e-6)


def test_check_collision_cases():
    w = (np.float32([0.0]), np.float32([0.0]), np.float32([10.0]), np.float32([0.0]))
    assert bool(osm.check_collision(*w, 5.0, -1.0, 5.0, 1.0))       # crosses
    assert not bool(osm.check_collision(*w, 5.0, 1.0, 6.0, 2.0))    # stays above
    assert bool(osm.check_collision(*w, 5.0, 0.0, 6.0, 3.0))        # starts on the wall
    assert not bool(osm.check_collision(*w, 11.0, -1.0, 11.0, 1.0))  # passes beyond the end


def test_c_oracle_equals_numpy_oracle():
    from oracle import c_oracle
    rng = np.random.default_rng(2)
    n, nw = 200, 300
    x, y = rng.uniform(1, 499, n).astype(np.float32), rng.uniform(1, 499, n).astype(np.float32)
    ang = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    x[:10] = 0.0
    w = rng.uniform(0, 500, (nw, 4)).astype(np.float32)
    w[:4] = [[0, 0, 500, 0], [500, 0, 500, 500], [0, 0, 0, 500], [0, 500, 500, 500]]
    walls = [np.ascontiguousarray(w[:, i]) for i in range(4)]
    act = (rng.random(n) < 0.8).astype(np.float32)
    d, h = c_oracle.raycast_walls(x, y, ang, act, 32, np.float32(np.pi / 4), 60.0, *walls)
    d2, h2 = osm.raycast_walls(x, y, ang, np.float32(np.pi / 4), 60.0, 32, *walls, active=act)
    np.testing.assert_array_equal(d, d2)
    np.testing.assert_array_equal(h, h2)
    assert (h >= 0).mean() > 0.1
