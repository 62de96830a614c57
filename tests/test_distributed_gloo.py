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


# ---- global_pack_active: sort(active_state, descending) over a sharded set -----------------------------
def _torch_sort_active(agents):
    """Test stand-in for the native local sort: stable 'actives first' partition + row gather."""
    from foragax_b200 import struct
    ids = torch.from_numpy(np.argsort(-agents.active_state.numpy(), kind="stable"))
    n = agents.active_state.shape[0]
    return struct.tree_map(lambda x: x[ids] if x.dim() and x.shape[0] == n else x, agents)


def _make_set(n, seed):
    from foragax_b200.base.agent_classes import Agent, State
    rng = np.random.default_rng(seed)
    active = (rng.random(n) < 0.55).astype(np.float32)
    return Agent(params=None, unique_id=torch.arange(n, dtype=torch.int32), agent_type=torch.zeros(n, dtype=torch.int32),
                 active_state=torch.from_numpy(active), age=torch.from_numpy(rng.random(n).astype(np.float32)),
                 state=State(content={'w': torch.from_numpy(rng.random((n, 3)).astype(np.float32)),
                                      'flag': torch.from_numpy(rng.random(n) < 0.5),
                                      'key': torch.from_numpy(rng.integers(0, 2**32, (n, 2), dtype=np.uint32))}),
                 policy=None)


def _pack_worker(rank, world, port, n, seed, sizes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from foragax_b200 import distributed as fd
    from foragax_b200 import struct
    full = _make_set(n, seed)
    lo, hi = sum(sizes[:rank]), sum(sizes[:rank + 1])
    mine = struct.tree_map(lambda x: x[lo:hi].clone(), full)
    out = fd.global_pack_active(mine, _torch_sort_active)
    q.put((rank, [x.numpy() for x in struct.tree_leaves(out)]))
    dist.destroy_process_group()


def _run_pack(world, n, seed, sizes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pack_worker, args=(r, world, port, n, seed, sizes, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out, key=lambda t: t[0])


def test_global_pack_active_equals_single_process_sort():
    """2 ranks (even shards) and 3 ranks (ragged shards, one of them with no active row to keep): every rank
    ends up with exactly its block of the single-process sort_agents(active_state, descending)."""
    from foragax_b200 import struct
    for world, n, sizes in ((2, 1000, [500, 500]), (3, 611, [300, 11, 300])):
        seed = 7 + world
        ref = _torch_sort_active(_make_set(n, seed))
        ref_leaves = [x.numpy() for x in struct.tree_leaves(ref)]
        res = _run_pack(world, n, seed, sizes)
        for rank, leaves in res:
            lo, hi = sum(sizes[:rank]), sum(sizes[:rank + 1])
            assert len(leaves) == len(ref_leaves)
            for got, want in zip(leaves, ref_leaves):
                np.testing.assert_array_equal(got, want[lo:hi])


def test_plan_pack_splits_is_a_bijection():
    from foragax_b200.distributed import plan_pack_splits
    rng = np.random.default_rng(0)
    for _ in range(200):
        w = int(rng.integers(1, 9))
        n = [int(v) for v in rng.integers(0, 50, w)]
        a = [int(rng.integers(0, v + 1)) for v in n]
        sp = plan_pack_splits(a, n)
        assert [sum(row) for row in sp] == n                          # every rank sends all its rows
        assert [sum(sp[s][d] for s in range(w)) for d in range(w)] == n  # and receives a full shard


# ---- distributed add: which rows each rank adds, and the parent rows it needs ---------------------------
def _add_worker(rank, world, port, n, seed, sizes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from foragax_b200 import distributed as fd
    from foragax_b200 import struct
    full = _torch_sort_active(_make_set(n, seed))          # globally packed set, as global_pack_active leaves it
    lo, hi = sum(sizes[:rank]), sum(sizes[:rank + 1])
    mine = struct.tree_map(lambda x: x[lo:hi].clone(), full)
    sel_mask = (mine.state.content['flag'] & (mine.active_state != 0)).numpy()
    cnt, perm = oam.select_from_mask(sel_mask)
    n_act = (mine.active_state != 0).sum()
    k, local_count, j0 = fd.global_add_range(torch.tensor(int(cnt)), n_act, hi - lo, lo, n)
    take = lambda a, ids: struct.tree_map(lambda x: x[ids.long()] if x.dim() and x.shape[0] == hi - lo else x, a)  # noqa: E731
    parents = fd.fetch_parent_rows(mine, torch.from_numpy(perm), torch.tensor(int(cnt)), n_act, take)
    q.put((rank, int(k), int(local_count), int(j0), [x.numpy() for x in struct.tree_leaves(parents)]))
    dist.destroy_process_group()


def test_distributed_add_range_and_parent_rows():
    from foragax_b200 import struct
    ctx = mp.get_context("spawn")
    for world, n, sizes in ((2, 400, [200, 200]), (3, 301, [100, 1, 200])):
        seed = 11 + world
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_add_worker, args=(r, world, port, n, seed, sizes, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        full = _torch_sort_active(_make_set(n, seed))
        act = full.active_state.numpy() != 0
        a = int(act.sum())
        sel = full.state.content['flag'].numpy() & act
        _, gperm = oam.select_from_mask(sel)
        k = min(int(sel.sum()), n - a)
        ref_parents = struct.tree_map(lambda x: x[torch.from_numpy(gperm[:k]).long()], full)  # parent of new slot a + j
        ref_leaves = [x.numpy() for x in struct.tree_leaves(ref_parents)]
        assert k > 0
        for rank, kk, local_count, j0, leaves in res:
            lo, hi = sum(sizes[:rank]), sum(sizes[:rank + 1])
            want_lo, want_hi = min(max(a, lo), hi), min(max(a + k, lo), hi)     # [a, a+k) cut by this block
            assert kk == k and local_count == want_hi - want_lo
            if local_count:
                assert j0 == want_lo - a
            for got, want in zip(leaves, ref_leaves):
                np.testing.assert_array_equal(got, want[want_lo - a:want_hi - a] if local_count else want[:0])


def _rr_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from foragax_b200 import distributed as fd
    x = np.arange(n, dtype=np.float32)
    mine = slice(rank, None, world)  # round-robin: slot i lives on rank i % world (bench.py's sharding)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    pos = fd.all_gather_positions(t(x[mine]), t(x[mine] * 2), t(x[mine] + 0.5), t((x[mine] % 3 == 0).astype(np.float32)),
                                  layout="round_robin")
    bad = None
    try:  # a block-sharded set passed off as round-robin must be refused when the shard sizes give it away
        if n % world:
            lo, hi = fd.shard_bounds(n)
            fd.all_gather_positions(t(x[lo:hi]), t(x[lo:hi]), t(x[lo:hi]), t(x[lo:hi]), layout="round_robin")
            bad = "ragged block shards were accepted as round_robin"
    except ValueError:
        pass
    q.put((rank, pos.numpy(), bad))
    dist.destroy_process_group()


def test_round_robin_positions_come_back_in_slot_order():
    """with slots dealt round-robin the gathered table must be un-interleaved, otherwise
    the hit indices of agent-agent sensing index the gathered order instead of the global slot ids."""
    ctx = mp.get_context("spawn")
    for world, n in ((2, 1000), (2, 1001), (3, 1000)):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_rr_worker, args=(r, world, port, n, q)) for r in range(world)]
        for p in procs:
            p.start()
        out = [q.get(timeout=120) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        x = np.arange(n, dtype=np.float32)
        for rank, pos, bad in out:
            assert bad is None, bad
            assert pos.shape == (n, 4)
            np.testing.assert_array_equal(pos[:, 0], x)
            np.testing.assert_array_equal(pos[:, 1], x * 2)
            np.testing.assert_array_equal(pos[:, 2], x + 0.5)
            np.testing.assert_array_equal(pos[:, 3], (x % 3 == 0).astype(np.float32))
