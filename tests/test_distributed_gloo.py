"""The N>1 host logic on CPU: two gloo ranks shard a slot range, compact locally (with the oracle,
no GPU here) and must reassemble exactly the single-process stable partition; positions all-gather
in slot order."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import agent_methods as oam


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from foragax_b200 import distributed as fd
    rng = np.random.default_rng(seed)
    mask = rng.random(n) < 0.37
    x = rng.random(n).astype(np.float32)
    lo, hi = fd.shard_bounds(n)
    cnt, perm = oam.select_from_mask(mask[lo:hi])
    total, gperm = fd.global_partition(torch.from_numpy(perm), torch.tensor(int(cnt)), lo, n)
    so, uo, tot = fd.global_partition_offsets(torch.tensor(int(cnt)), hi - lo)
    pos = fd.all_gather_positions(torch.from_numpy(x[lo:hi]), torch.from_numpy(x[lo:hi]) * 2,
                                  torch.ones(hi - lo), torch.from_numpy(mask[lo:hi].astype(np.float32)))
    q.put((rank, int(total), gperm.numpy(), int(so), int(uo), pos.numpy(), (lo, hi)))
    dist.destroy_process_group()


def _run(world, n, seed):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out, key=lambda t: t[0])


def test_two_rank_partition_equals_single_process():
    n, seed = 1001, 5
    res = _run(2, n, seed)
    rng = np.random.default_rng(seed)
    mask = rng.random(n) < 0.37
    x = rng.random(n).astype(np.float32)
    rc, rp = oam.select_from_mask(mask)
    sel_seen = 0
    for rank, total, gperm, so, uo, pos, (lo, hi) in res:
        assert total == int(rc)
        np.testing.assert_array_equal(gperm, rp)          # identical on every rank, equal to 1-process
        assert so == sel_seen                              # exclusive prefix of selected counts
        assert uo == int(rc) + (lo - sel_seen)
        sel_seen += int(mask[lo:hi].sum())
        np.testing.assert_array_equal(pos[:, 0], x)        # slot order preserved by the all-gather
        np.testing.assert_array_equal(pos[:, 1], x * 2)
        np.testing.assert_array_equal(pos[:, 3], mask.astype(np.float32))
    assert res[0][6] == (0, 500) and res[1][6] == (500, 1001)


def test_shard_bounds_cover_range():
    from foragax_b200.distributed import shard_bounds
    for n in (0, 1, 7, 1_000_000):
        for w in (1, 2, 4, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
