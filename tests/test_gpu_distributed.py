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
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res)
    total_active = res[0][3]
    # actives are packed at the GLOBAL front: rank 0's shard fills up before rank 1 gets any
    assert res[0][2] == min(total_active, n // 2) and res[0][2] + res[1][2] == total_active
