"""The conservativeness invariant of the wall grid (csrc/raycast_walls_grid.cu::wall_column_cells), checked on a
NumPy restatement of the rasterisation: the cell of ANY point that lies within `thick` of a wall's segment, or within
`eps` of its infinite line, lists that wall.  (That the CUDA code implements this rasterisation and that the cull
reproduces the all-pairs result bit for bit is what tests/test_gpu_raycast.py::test_grid_cull_* check on the GPU.)"""
import numpy as np

F = np.float32


def grid_dims(x_min, x_max, y_min, y_max, L, div=8.0):
    ext = max(x_max - x_min, y_max - y_min)
    scale = 2.0 * max(abs(x_min), abs(x_max), abs(y_min), abs(y_max))
    cs = L / div
    if ext / cs > 1024.0:
        cs = ext / 1024.0
    gx, gy = max(1, int(np.ceil((x_max - x_min) / cs))), max(1, int(np.ceil((y_max - y_min) / cs)))
    return dict(x0=x_min, y0=y_min, cs=cs, gx=gx, gy=gy, scale=scale, eps=1e-4 * scale + 1e-6,
                thick=L * 1.0001 + 5e-4 * scale + 1e-6)


def coord(v, v0, cs, n):
    f = np.floor((v - v0) / cs)
    return int(min(max(f, 0), n - 1))


def wall_cells(g, ax, ay, bx, by):
    xmajor = abs(bx - ax) >= abs(by - ay)
    nu, nv = (g["gx"], g["gy"]) if xmajor else (g["gy"], g["gx"])
    au, av, bu, bv = (ax, ay, bx, by) if xmajor else (ay, ax, by, bx)
    if au > bu:
        au, av, bu, bv = bu, bv, au, av
    du = bu - au
    m = (bv - av) / du if du > 0 else 0.0
    wv = np.sqrt(1.0 + m * m)
    u0, v0 = (g["x0"], g["y0"]) if xmajor else (g["y0"], g["x0"])
    pad = 0.02 * g["cs"] + 1e-5 * g["scale"]
    out = set()
    for cu in range(nu):
        ulo, uhi = u0 + cu * g["cs"] - pad, u0 + (cu + 1) * g["cs"] + pad
        thick = uhi >= au - g["thick"] and ulo <= bu + g["thick"]
        hw = (g["thick"] if thick else g["eps"]) * wv + pad
        v1, v2 = av + m * (ulo - au), av + m * (uhi - au)
        vlo, vhi = min(v1, v2) - hw, max(v1, v2) + hw
        if vhi < v0 or vlo > v0 + nv * g["cs"]:
            continue
        for cv in range(coord(vlo, v0, g["cs"], nv), coord(vhi, v0, g["cs"], nv) + 1):
            out.add(cv * g["gx"] + cu if xmajor else cu * g["gx"] + cv)
    return out


def seg_dist(px, py, ax, ay, bx, by):
    vx, vy = bx - ax, by - ay
    l2 = vx * vx + vy * vy
    t = 0.0 if l2 == 0 else min(1.0, max(0.0, ((px - ax) * vx + (py - ay) * vy) / l2))
    return np.hypot(px - (ax + t * vx), py - (ay + t * vy))


def line_dist(px, py, ax, ay, bx, by):
    vx, vy = bx - ax, by - ay
    ln = np.hypot(vx, vy)
    return np.inf if ln == 0 else abs(vx * (py - ay) - vy * (px - ax)) / ln


def test_every_point_near_a_wall_or_on_its_line_finds_it_in_its_cell():
    rng = np.random.default_rng(0)
    size, L = 600.0, 60.0
    g = grid_dims(0.0, size, 0.0, size, L)
    checked_seg = checked_line = 0
    for _ in range(60):
        kind = rng.integers(0, 4)
        mx, my, ln, th = rng.uniform(0, size), rng.uniform(0, size), rng.uniform(0, 250), rng.uniform(0, np.pi)
        if kind == 0:     # axis-aligned
            th = rng.choice([0.0, np.pi / 2])
        elif kind == 1:   # a point
            ln = 0.0
        ax, ay, bx, by = (mx - ln / 2 * np.cos(th), my - ln / 2 * np.sin(th), mx + ln / 2 * np.cos(th),
                          my + ln / 2 * np.sin(th))
        cells = wall_cells(g, ax, ay, bx, by)
        # points around the segment (within reach) and along the infinite line (anywhere in the arena)
        pts = []
        for _ in range(60):
            t, r, a = rng.uniform(0, 1), rng.uniform(0, g["thick"]), rng.uniform(0, 2 * np.pi)
            pts.append((ax + t * (bx - ax) + r * np.cos(a), ay + t * (by - ay) + r * np.sin(a)))
        if ln > 0:
            ux, uy = (bx - ax) / ln, (by - ay) / ln
            for _ in range(60):
                t, off = rng.uniform(-2 * size, 2 * size), rng.uniform(-g["eps"], g["eps"])
                pts.append((ax + t * ux - off * uy, ay + t * uy + off * ux))
        for px, py in pts:
            if not (0.0 <= px < g["gx"] * g["cs"] and 0.0 <= py < g["gy"] * g["cs"]):
                continue   # outside the grid: the kernel scans all walls for such agents
            near_seg = seg_dist(px, py, ax, ay, bx, by) <= g["thick"]
            near_line = line_dist(px, py, ax, ay, bx, by) <= g["eps"]
            if near_seg or near_line:
                cell = coord(py, g["y0"], g["cs"], g["gy"]) * g["gx"] + coord(px, g["x0"], g["cs"], g["gx"])
                assert cell in cells, (ax, ay, bx, by, px, py)
                checked_seg += near_seg
                checked_line += near_line and not near_seg
    assert checked_seg > 2000 and checked_line > 300


def test_lists_stay_short():
    """The cull is only useful if a cell lists few walls: ~8 of 1,000 for the C4 arena (thin line bands included)."""
    rng = np.random.default_rng(1)
    size, L, n_walls = 4096.0, 60.0, 120
    g = grid_dims(0.0, size, 0.0, size, L)
    total = 0
    for _ in range(n_walls):
        mx, my, ln, th = rng.uniform(0, size), rng.uniform(0, size), rng.uniform(20, 200), rng.uniform(0, np.pi)
        total += len(wall_cells(g, mx - ln / 2 * np.cos(th), my - ln / 2 * np.sin(th), mx + ln / 2 * np.cos(th),
                                my + ln / 2 * np.sin(th)))
    per_cell_per_1000_walls = total / n_walls * 1000 / (g["gx"] * g["gy"])
    assert per_cell_per_1000_walls < 12.0
