"""BASELINE config 5 sharded over the GPUs of one node: the Dice churn tick of ``churn.py`` on a set
split into contiguous blocks of slots, one process per GPU.  Per tick and rank:

    local step / select(draw == 1) / remove
    -> GLOBAL ``sort_agents(active_state, descending)``: local stable partition (``fgx_select``), counts exchanged
       through peer-memory mailboxes (``fgx_peer_exchange``), ONE kernel that compacts and scatters every row to its
       slot of the global order on whichever GPU owns it (``fgx_pack_rows_peer``: stores over NVLink / NVSwitch),
       a device-side flag barrier
    -> local select(draw == sides) -> global add range (``fgx_peer_add_range``) -> local add

There is no NCCL call and no host read-back inside the tick, so two consecutive ticks (buffer set A -> B -> A) are
captured once as a CUDA graph on every rank and replayed.  The result equals the single-GPU tick block by block
(tests/test_gpu_distributed.py).
"""
import torch
import torch.distributed as dist

from .. import distributed as fd
from ..base import agent_classes as fgx_classes
from ..base import agent_methods as fgx_methods
from .. import random as fgx_random
from ..peer import PeerPackedSet
from .hello_world import is_value, make_dice


class ShardedChurnWorld:
    def __init__(self, n_slots, sides=20, n_active=None, key=None):
        self.n, self.sides = int(n_slots), int(sides)
        self.rank, self.world = fd.world()
        self.D = make_dice(self.sides)
        n_active = self.n // 2 if n_active is None else int(n_active)
        self.lo, self.hi = fd.shard_bounds(self.n, self.world, self.rank)
        self.n_local = self.hi - self.lo
        whole = fgx_classes.Agent_Set(agent=self.D, num_total_agents=self.n, num_active_agents=n_active, agent_type=0)
        # this rank's rows of the single-GPU set (fgx_create_base_shard)
        block = fgx_methods.create_agents(None, whole, fgx_random.PRNGKey(0) if key is None else key,
                                          shard=(self.lo, self.n_local))
        self.set = fgx_classes.Agent_Set(agent=self.D, num_total_agents=self.n_local, num_active_agents=0, agent_type=0)
        self.packed = PeerPackedSet(block)
        self.set.agents = self.packed.agents
        self._dead, self._birth = is_value(1), is_value(self.sides)

    @property
    def agents(self):
        return self.set.agents

    def tick(self):
        s, D, st = self.set, self.D, self.packed
        fgx_methods.step_agents(None, None, s)
        n_dead, rem = fgx_methods.jit_select_agents(self._dead, None, s.agents)
        fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}), s.agents)
        s.agents = st.pack_active()
        n_sel, _ = fgx_methods.jit_select_agents(self._birth, None, s.agents)
        n_act = fgx_methods.count_active(s.agents)
        _, local_count, _ = st.add_range(n_sel, n_act, self.n)
        fgx_methods.jit_add_agents(D.add_agent, local_count, None, s.agents, None, num_active=n_act)

    def capture_two_ticks(self):
        """Two ticks (the state returns to the buffer set it started from) as one CUDA graph; ``None`` if the
        capture fails on this stack (the eager tick is then the product path)."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up on a side stream: allocator pools, lazy module loads
            self.tick()
            self.tick()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        try:
            # thread_local: the NCCL watchdog thread of the process group keeps polling events while we capture
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                self.tick()
                self.tick()
        except RuntimeError:
            return None
        return graph

    def time_ticks(self, ticks, warmup=3):
        """Eager and graph-replayed tick times on THIS rank (the caller takes the max over ranks)."""
        for _ in range(warmup):
            self.tick()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ticks = max(2, ticks // 2 * 2)
        e0.record()
        for _ in range(ticks):
            self.tick()
        e1.record()
        torch.cuda.synchronize()
        eager_ms = e0.elapsed_time(e1) / ticks
        self.packed.check()
        graph = self.capture_two_ticks()
        ok = torch.tensor([1 if graph is not None else 0], device=self.agents.active_state.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        ms, mode = eager_ms, "eager (graph capture failed)"
        if int(ok):
            for _ in range(2):
                graph.replay()
            torch.cuda.synchronize()
            dist.barrier()
            e0.record()
            for _ in range(ticks // 2):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            ms, mode = e0.elapsed_time(e1) / ticks, "CUDA graph of two ticks per rank, replayed"
            self.packed.check()
        # rows that cross NVLink in one pack: the exchange plan from the active counts AFTER the removals of a tick
        # (the state the pack sees); the tick is completed afterwards so that the set stays consistent
        s, D, st = self.set, self.D, self.packed
        fgx_methods.step_agents(None, None, s)
        n_dead, rem = fgx_methods.jit_select_agents(self._dead, None, s.agents)
        fgx_methods.jit_remove_agents(D.remove_agent, n_dead, fgx_classes.Params(content={'remove_ids': rem}), s.agents)
        acts = fd.all_gather_counts(fgx_methods.count_active(s.agents).reshape(()))
        sizes = [self.packed.bounds[i + 1] - self.packed.bounds[i] for i in range(self.world)]
        splits = fd.plan_pack_splits(acts.tolist(), sizes)
        remote = sum(v for d, v in enumerate(splits[self.rank]) if d != self.rank)
        s.agents = st.pack_active()
        n_sel, _ = fgx_methods.jit_select_agents(self._birth, None, s.agents)
        n_act = fgx_methods.count_active(s.agents)
        _, local_count, _ = st.add_range(n_sel, n_act, self.n)
        fgx_methods.jit_add_agents(D.add_agent, local_count, None, s.agents, None, num_active=n_act)
        n_act = fgx_methods.count_active(self.agents)
        return {"ms_per_tick": ms, "eager_ms_per_tick": eager_ms, "mode": mode, "final_active_local": int(n_act),
                "rows_sent_remote_per_tick": remote}
