"""The row-moving exchange of SURVEY.md 8(e)(3) on real GPUs: two NCCL ranks, the native local sort, and
an all-to-all of agent rows must give every rank its block of the single-GPU sort_agents(active_state,
descending).  Needs 2 visible GPUs (skipped otherwise; the same logic runs under gloo in
tests/test_distributed_gloo.py)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _collect(procs, q, world, timeout=240):
    """Results of all workers; fails as soon as one of them dies instead of waiting for the queue to time out."""
    import queue as _queue
    import time
    out, t0 = [], time.time()
    while len(out) < world:
        try:
            out.append(q.get(timeout=2))
        except _queue.Empty:
            dead = [p.exitcode for p in procs if p.exitcode not in (None, 0)]
            if dead or time.time() - t0 > timeout:
                for p in procs:
                    if p.is_alive():
                        p.kill()
                raise AssertionError(f"worker failed (exit codes {dead}) or timed out")
    return sorted(out)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import foragax_b200.base.agent_classes as fgx_classes
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import distributed as fd
    from foragax_b200 import random as fgx_random
    from foragax_b200 import struct
    from foragax_b200.examples.hello_world import is_value, make_dice
    D = make_dice(6)
    s = fgx_classes.Agent_Set(agent=D, num_total_agents=n, num_active_agents=n * 3 // 5, agent_type=0)
    s.agents = fgx_methods.create_agents(None, s, fgx_random.PRNGKey(3))
    s.agents = fgx_methods.step_agents(None, None, s)
    n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, s.agents)
    full = fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}),
                                         s.agents)
    lo, hi = fd.shard_bounds(n)
    mine = struct.tree_map(lambda x: x[lo:hi].clone() if x.dim() and x.shape[0] == n else x, full)
    sort_active = lambda a: fgx_methods.jit_sort_agents(a.active_state, False, a)[0]  # noqa: E731
    got = fd.global_pack_active(mine, sort_active, fgx_methods.count_active(mine))
    ref = sort_active(full)
    ok = all(torch.equal(g, r[lo:hi]) for g, r in zip(struct.tree_leaves(got), struct.tree_leaves(ref))
             if r.dim() and r.shape[0] == n)
    n_act = int(fgx_methods.count_active(got))
    q.put((rank, ok, n_act, int(fgx_methods.count_active(full))))
    dist.destroy_process_group()


def test_two_gpu_pack_active_equals_single_gpu_sort():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n = 2, 200_001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q, world)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    total_active = res[0][3]
    # actives are packed at the GLOBAL front: rank 0's shard fills up before rank 1 gets any
    assert res[0][2] == min(total_active, n // 2) and res[0][2] + res[1][2] == total_active


def _churn_worker(rank, world, port, n, ticks, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import foragax_b200.base.agent_classes as fgx_classes
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import distributed as fd
    from foragax_b200 import random as fgx_random
    from foragax_b200 import struct
    from foragax_b200.examples.hello_world import is_value, make_dice
    sides = 20
    D = make_dice(sides)
    rows = lambda t, f: struct.tree_map(lambda x: f(x) if x.dim() and x.shape[0] == n else x, t)  # noqa: E731
    s = fgx_classes.Agent_Set(agent=D, num_total_agents=n, num_active_agents=n // 2, agent_type=0)
    s.agents = fgx_methods.create_agents(None, s, fgx_random.PRNGKey(9))
    lo, hi = fd.shard_bounds(n)
    n_local = hi - lo
    mine = rows(s.agents, lambda x: x[lo:hi].clone())       # this rank's block of the same initial set
    local = fgx_classes.Agent_Set(agent=D, num_total_agents=n_local, num_active_agents=0, agent_type=0)
    sort_active = lambda a: fgx_methods.jit_sort_agents(a.active_state, False, a)[0]  # noqa: E731
    ok = True
    for _ in range(ticks):
        # single-GPU tick on the full set (every rank computes it: the reference result)
        s.agents = fgx_methods.step_agents(None, None, s)
        n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, s.agents)
        s.agents = fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}),
                                                 s.agents)
        s.agents = sort_active(s.agents)
        n_add, _ = fgx_methods.jit_select_agents(is_value(sides), None, s.agents)
        n_add = torch.minimum(n_add, n - fgx_methods.count_active(s.agents))
        s.agents, _ = fgx_methods.jit_add_agents(D.add_agent, n_add, None, s.agents, None)
        # the same tick on the sharded set: local step / select / remove, global pack, global add range
        local.agents = mine
        mine = fgx_methods.step_agents(None, None, local)
        n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, mine)
        mine = fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}),
                                             mine)
        mine = fd.global_pack_active(mine, sort_active, fgx_methods.count_active(mine))
        n_sel, _ = fgx_methods.jit_select_agents(is_value(sides), None, mine)
        k, local_count, _ = fd.global_add_range(n_sel, fgx_methods.count_active(mine), n_local, lo, n)
        mine, _ = fgx_methods.jit_add_agents(D.add_agent, local_count, None, mine, None)
        ok = ok and all(torch.equal(g, r[lo:hi]) for g, r in zip(struct.tree_leaves(mine), struct.tree_leaves(s.agents))
                        if r.dim() and r.shape[0] == n)
    q.put((rank, ok, int(fgx_methods.count_active(mine)), int(fgx_methods.count_active(s.agents))))
    dist.destroy_process_group()


def test_two_gpu_sharded_churn_ticks_equal_single_gpu():
    """step -> select -> remove -> GLOBAL sort (all-to-all) -> select -> clip -> GLOBAL add, 6 ticks on two
    ranks: every rank's block equals the single-GPU set after every tick, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n = 2, 100_000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_churn_worker, args=(r, world, port, n, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q, world)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] + res[1][2] == res[0][3]


def _peer_worker(rank, world, port, n, ticks, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import foragax_b200.base.agent_classes as fgx_classes
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import distributed as fd
    from foragax_b200 import random as fgx_random
    from foragax_b200 import struct
    from foragax_b200.examples.hello_world import is_value, make_dice
    from foragax_b200.peer import PeerPackedSet
    sides = 20
    D = make_dice(sides)
    s = fgx_classes.Agent_Set(agent=D, num_total_agents=n, num_active_agents=n // 2, agent_type=0)
    s.agents = fgx_methods.create_agents(None, s, fgx_random.PRNGKey(4))
    lo, hi = fd.shard_bounds(n)
    n_local = hi - lo
    mine = struct.tree_map(lambda x: x[lo:hi].clone() if x.dim() and x.shape[0] == n else x, s.agents)
    st = PeerPackedSet(mine)
    local = fgx_classes.Agent_Set(agent=D, num_total_agents=n_local, num_active_agents=0, agent_type=0)
    sort_active = lambda a: fgx_methods.jit_sort_agents(a.active_state, False, a)[0]  # noqa: E731
    ok = True
    for _ in range(ticks):
        s.agents = fgx_methods.step_agents(None, None, s)
        n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, s.agents)
        s.agents = fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}),
                                                 s.agents)
        s.agents = sort_active(s.agents)
        n_add, _ = fgx_methods.jit_select_agents(is_value(sides), None, s.agents)
        n_add = torch.minimum(n_add, n - fgx_methods.count_active(s.agents))
        s.agents, _ = fgx_methods.jit_add_agents(D.add_agent, n_add, None, s.agents, None)
        # sharded tick, state in peer-mapped memory: in-place local ops + ONE scatter kernel for the global pack
        local.agents = st.agents
        fgx_methods.step_agents(None, None, local)
        n_dead, rem = fgx_methods.jit_select_agents(is_value(1), None, st.agents)
        fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}), st.agents)
        packed = st.pack_active()
        n_sel, _ = fgx_methods.jit_select_agents(is_value(sides), None, packed)
        k, local_count, _ = fd.global_add_range(n_sel, fgx_methods.count_active(packed), n_local, lo, n)
        fgx_methods.jit_add_agents(D.add_agent, local_count, None, packed, None)
        ok = ok and all(torch.equal(g, r[lo:hi]) for g, r in zip(struct.tree_leaves(st.agents), struct.tree_leaves(s.agents))
                        if r.dim() and r.shape[0] == n)
    q.put((rank, ok, int(fgx_methods.count_active(st.agents)), int(fgx_methods.count_active(s.agents))))
    dist.destroy_process_group()


def test_two_gpu_peer_memory_pack_equals_single_gpu():
    """The same sharded churn ticks with the state in symmetric memory and the global pack done by
    fgx_pack_rows_peer (one kernel: compaction + stores into the peer's buffers over NVLink)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n = 2, 100_001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, world, port, n, 6, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q, world)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    assert res[0][2] + res[1][2] == res[0][3]


def _sort_worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import foragax_b200.base.agent_classes as fgx_classes
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import distributed as fd
    from foragax_b200 import random as fgx_random
    from foragax_b200 import struct
    from foragax_b200.examples.hello_world import make_dice
    from foragax_b200.peer import PeerPackedSet
    D = make_dice(20)
    s = fgx_classes.Agent_Set(agent=D, num_total_agents=n, num_active_agents=n * 2 // 3, agent_type=0)
    s.agents = fgx_methods.create_agents(None, s, fgx_random.PRNGKey(6))
    for _ in range(3):
        s.agents = fgx_methods.step_agents(None, None, s)
    g = torch.Generator(device="cuda").manual_seed(5)   # same on every rank
    keyf = torch.randint(-50, 50, (n,), device="cuda", generator=g).float() * 0.5   # many ties
    keyf[::97] = float("nan")
    keyf[5::101] = -0.0
    s.agents = s.agents.replace(age=keyf)
    lo, hi = fd.shard_bounds(n)
    mine = struct.tree_map(lambda x: x[lo:hi].clone() if x.dim() and x.shape[0] == n else x, s.agents)
    st = PeerPackedSet(mine)
    ok = True
    for leaf, asc in (("age", True), ("age", False), ("draw", True), ("draw", False)):
        get = (lambda a: a.age) if leaf == "age" else (lambda a: a.state.content['draw'])  # noqa: E731
        ref, _ = fgx_methods.jit_sort_agents(get(s.agents), asc, s.agents)
        got = st.sort(get, asc)
        for gl, rl in zip(struct.tree_leaves(got), struct.tree_leaves(ref)):
            if rl.dim() and rl.shape[0] == n:
                a, b = gl, rl[lo:hi]
                same = torch.equal(a, b) if a.dtype != torch.float32 else torch.equal(a.view(torch.int32), b.view(torch.int32))
                ok = ok and same
        s.agents = ref   # keep both sides in the same state for the next sort
    q.put((rank, ok))
    dist.destroy_process_group()


def test_two_gpu_general_sort_equals_single_gpu():
    """PeerPackedSet.sort: float keys with ties, NaN and -0.0 (ascending / descending) and an int key, on ragged
    shards -- every rank ends with its block of the single-GPU sort_agents, bit for bit."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n = 2, 60_001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sort_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q, world)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def _forager_add_worker(rank, world, port, n, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import foragax_b200.base.agent_methods as fgx_methods
    from foragax_b200 import distributed as fd
    from foragax_b200 import random as fgx_random
    from foragax_b200 import struct
    from foragax_b200.base.agent_classes import Agent_Set, Params
    from foragax_b200.base.predicates import P
    from foragax_b200.base.space_methods import create_space
    from foragax_b200.examples import foraging as fg
    fg.use_variant("sim")
    space = create_space(0.0, 1000.0, 0.0, 500.0, False, None)
    fset = Agent_Set(agent=fg.Forager, num_total_agents=n, num_active_agents=n * 3 // 5, agent_type=1)
    cp = Params(content={'policy_params': {'num_neurons': 40, 'num_obs': 28, 'num_actions': 2, 'dt': 0.1,
                                           'action_scale': 1.0, 'deterministic': True}, 'dt': 0.1, 'space': space})
    full = fgx_methods.create_agents(cp, fset, fgx_random.PRNGKey(2))
    g = torch.Generator(device="cuda").manual_seed(3)
    full.state.content['time_above_rep_thresh'].copy_(torch.rand((n, 1), device="cuda", generator=g))
    full.state.content['energy'].copy_(torch.rand((n, 1), device="cuda", generator=g) * 10)
    full.age.copy_(torch.rand(n, device="cuda", generator=g) * 50)

    def good(a, p):
        return (P(a.active_state) == 1.0) & (P(a.state.content['time_above_rep_thresh']) > 0.55)

    lo, hi = fd.shard_bounds(n)
    mine = struct.tree_map(lambda x: x[lo:hi].clone() if x.dim() and x.shape[0] == n else x, full)
    key = fgx_random.PRNGKey(11)
    # single-GPU reference (sim.py:518-524): select, clip, add, set
    n_add, ids = fgx_methods.jit_select_agents(good, None, full)
    n_add = torch.minimum(n_add, n - fgx_methods.count_active(full))
    ref, key_ref = fgx_methods.jit_add_agents(fg.Forager.add_agent, n_add, Params(content={'good_ids': ids}), full, key)
    ref, _ = fgx_methods.jit_set_agents(fg.Forager.set_agent, n_add, Params(content={'set_ids': ids}), ref, key_ref)
    # the same on the sharded set
    n_sel, lids = fgx_methods.jit_select_agents(good, None, mine)
    def parent_leaves(a):   # what Forager.add_agent reads from a parent (sim.py:158-200) -- and overwrites in the child
        c, pc = a.state.content, a.policy.params.content
        return [c['X'], c['Y'], c['energy'], a.age, pc['J'], pc['tau'], pc['E'], pc['B'], pc['D']]

    got, key_got, n_par = fd.global_add_agents(fg.Forager.add_agent, lambda i: Params(content={'good_ids': i}), mine,
                                               parent_leaves, lids, n_sel, key, lo, n, chain_mode=0)
    got, _ = fgx_methods.jit_set_agents(fg.Forager.set_agent, n_par, Params(content={'set_ids': lids}), got, key_got)
    ok = torch.equal(key_got, key_ref)
    for a, b in zip(struct.tree_leaves(got), struct.tree_leaves(ref)):
        if b.dim() and b.shape[0] == n:
            ok = ok and torch.equal(a, b[lo:hi])
    q.put((rank, ok, int(n_add), int(fgx_methods.count_active(got))))
    dist.destroy_process_group()


def test_two_gpu_forager_add_with_parents_equals_single_gpu():
    """Neuroevolution's add step (select good foragers -> clip -> add_agents with mutated copies of the parents'
    networks, serial key chain -> set_agents on the parents) on two ranks: parents' 12 KB rows travel to the ranks
    that own the new slots; every leaf and the returned key equal the single-GPU result."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n = 2, 3001
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_forager_add_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q, world)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res) and res[0][2] > 100


def _sense_worker(rank, world, port, n, layout, q):
    """Agent-agent sensing over a sharded set: all-gather of the float4 positions in SLOT order, then every
    rank senses its own agents; the hit indices must be the global slot ids the single-GPU call returns."""
    import numpy as np
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import foragax_b200.base.space_methods as sm
    from foragax_b200 import distributed as fd
    rng = np.random.default_rng(11)
    arena, L, R = 600.0, 40.0, 16
    x = rng.uniform(1, arena - 1, n).astype(np.float32)
    y = rng.uniform(1, arena - 1, n).astype(np.float32)
    ang = rng.uniform(-np.pi, np.pi, n).astype(np.float32)
    rad = rng.uniform(2, 6, n).astype(np.float32)
    act = (rng.random(n) < 0.8).astype(np.float32)
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    bounds = (0.0, arena, 0.0, arena)
    full_d, full_h = sm.raycast_agents((up(x), up(y), up(ang)), 0.7, L, (up(x), up(y), up(rad), up(act)), bounds, 6.0,
                                       active_state=up(act), n_rays=R)
    if layout == "round_robin":
        mine = slice(rank, None, world)
    else:
        lo, hi = fd.shard_bounds(n)
        mine = slice(lo, hi)
    tg = fd.all_gather_positions(up(x[mine]), up(y[mine]), up(rad[mine]), up(act[mine]), layout=layout)
    d, h = sm.raycast_agents((up(x[mine]), up(y[mine]), up(ang[mine])), 0.7, L, tg, bounds, 6.0,
                             active_state=up(act[mine]), n_rays=R)
    ok = torch.equal(d, full_d[mine]) and torch.equal(h, full_h[mine])
    # the same table through peer memory (no NCCL: stage, mailbox barrier, pull over NVLink), several rounds
    from foragax_b200.peer import PeerPositions
    pp = PeerPositions(int(up(x[mine]).shape[0]), layout=layout)
    for rnd in range(3):
        shift = float(rnd)
        t2 = pp.gather(up(x[mine]) + shift, up(y[mine]), up(rad[mine]), up(act[mine]))
        want = tg.clone()
        want[:, 0] += shift
        ok = ok and torch.equal(t2, want)
    pp.check()
    q.put((rank, bool(ok), float((full_h >= 0).float().mean())))
    dist.destroy_process_group()


@pytest.mark.parametrize("layout,n", [("round_robin", 30_001), ("block", 30_001)])
def test_two_gpu_agent_sensing_hit_indices_are_global_slots(layout, n):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sense_worker, args=(r, world, port, n, layout, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = _collect(procs, q, world)
    for p in procs:
        p.join(timeout=60)
    for rank, ok, frac in out:
        assert ok, f"rank {rank}: sharded sensing differs from the single-GPU result ({layout})"
        assert frac > 0.05


def _sharded_world_worker(rank, world, port, n, q):
    """examples/churn_sharded.py (no NCCL call inside the tick: counts and ordering through peer-memory mailboxes)
    against examples/churn.py on the whole set: eager ticks, then the SAME ticks replayed from a CUDA graph."""
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from foragax_b200 import struct
    from foragax_b200.examples.churn import ChurnWorld
    from foragax_b200.examples.churn_sharded import ShardedChurnWorld
    whole = ChurnWorld(n)
    sh = ShardedChurnWorld(n)
    lo, hi = sh.lo, sh.hi

    def same():
        return all(torch.equal(g, r[lo:hi]) for g, r in zip(struct.tree_leaves(sh.agents), struct.tree_leaves(whole.agents))
                   if r.dim() and r.shape[0] == n)

    ok_eager = same()
    for _ in range(4):
        whole.tick()
        sh.tick()
        ok_eager = ok_eager and same()
    sh.packed.check()
    graph = sh.capture_two_ticks()          # warm-up runs 2 ticks, the capture itself runs none
    captured = graph is not None
    ok_graph = True
    if captured:
        for _ in range(2):
            whole.tick()
        for _ in range(3):
            graph.replay()                  # 6 ticks
            for _ in range(2):
                whole.tick()
            torch.cuda.synchronize()
            ok_graph = ok_graph and same()
        sh.packed.check()
    q.put((rank, bool(ok_eager), captured, bool(ok_graph)))
    dist.destroy_process_group()


def test_two_gpu_sharded_churn_world_eager_and_graph():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world, n = 2, 400_003
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_world_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = _collect(procs, q, world)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_eager, captured, ok_graph in res:
        assert ok_eager, f"rank {rank}: sharded eager ticks differ from the single-GPU set"
        assert captured, f"rank {rank}: the two-tick CUDA graph could not be captured"
        assert ok_graph, f"rank {rank}: graph-replayed ticks differ from the single-GPU set"


def test_one_rank_sharded_churn_world_eager_and_graph():
    """The same world with ONE rank (runs on a 1-GPU box): symmetric-memory state, mailbox exchange with itself,
    two-tick CUDA graph -- equal to the plain single-GPU ChurnWorld."""
    if torch.cuda.device_count() < 1:
        pytest.skip("needs a GPU")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sharded_world_worker, args=(0, 1, port, 250_001, q))]
    procs[0].start()
    res = _collect(procs, q, 1)
    procs[0].join(timeout=60)
    assert procs[0].exitcode == 0
    rank, ok_eager, captured, ok_graph = res[0]
    assert ok_eager and captured and ok_graph
