"""Block-sharded agent state in peer-mapped (symmetric) device memory, and the global "actives first" pack as
ONE kernel per rank that compacts and exchanges at once (``fgx_pack_rows_peer``: stores over NVLink /
NVSwitch straight into the destination ranks' buffers).

``distributed.global_pack_active`` does the same job with NCCL: local sort, counts to the HOST (the
collective needs split sizes), one all-to-all per leaf, local sort again.  Here nothing leaves the device:
the local stable partition (``fgx_select``) is never materialised as sorted rows -- each row is written
once, directly to its slot of the global order on whichever GPU owns it -- and the only collectives are an
all-gather of one count per rank and the all-reduce that orders the ranks afterwards.

The state is double-buffered (the pack reads buffer set ``cur`` and writes set ``1 - cur`` on every rank),
so a rank never overwrites rows a peer may still be reading: set ``1 - cur`` was last READ by the pack two
generations ago, and every rank passed the ordering all-reduce since.

Requires ``active_state`` to hold flags (1.0 / 0.0, agent_methods.py:14): the order is the stable partition
by ``active_state != 0``, which is what ``argsort(-active_state)`` gives for flags.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _lib as L
from . import struct
from .base.predicates import P


class PeerPackedSet:
    def __init__(self, agents, group=None, row_mask=None):
        """``row_mask`` (optional, one bool per leaf in tree order): which leaves are row leaves; default = every
        tensor whose leading dimension is the shard size (see ``distributed.row_leaf_mask``)."""
        import torch.distributed._symmetric_memory as symm
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > L.MAX_PEERS:
            raise L.FgxError(f"at most {L.MAX_PEERS} ranks of one node")
        dev = agents.active_state.device
        self.n_local = int(agents.active_state.shape[0])
        sizes = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([self.n_local], dtype=torch.int64, device=dev), group=self.group)
        self.bounds = [0]
        for v in sizes.tolist():  # once, at construction
            self.bounds.append(self.bounds[-1] + int(v))
        self.n_max = max(self.bounds[i + 1] - self.bounds[i] for i in range(self.world))
        self._template = agents
        leaves = struct.tree_leaves(agents)
        from .distributed import check_row_leaves_agree, row_leaf_mask
        self._is_row = row_leaf_mask(agents, self.n_local, row_mask)
        check_row_leaves_agree(self._is_row, dev)
        # one symmetric allocation per buffer set; every rank uses the layout of the LARGEST shard so that a
        # leaf sits at the same offset on every rank
        self._offsets, off = [], 0
        for x, row in zip(leaves, self._is_row):
            self._offsets.append(off)
            if row:
                row_bytes = x.element_size() * (x.numel() // max(self.n_local, 1)) if self.n_local else x.element_size()
                off += (row_bytes * self.n_max + 255) // 256 * 256
        self._row_bytes = [x.element_size() * (x.numel() // self.n_local) if row and self.n_local else 0
                           for x, row in zip(leaves, self._is_row)]
        total = max(off, 256)
        self._bufs, self._ptrs, self._handles = [], [], []
        for _ in range(2):
            buf = symm.empty(total, dtype=torch.uint8, device=dev)
            hdl = symm.rendezvous(buf, self.group)
            self._bufs.append(buf)
            self._handles.append(hdl)
            self._ptrs.append([int(p) for p in hdl.buffer_ptrs])
        self._views = [self._make_views(b, leaves) for b in self._bufs]
        for v, x, row in zip(self._views[0], leaves, self._is_row):
            if row:
                v.copy_(x)
        self.cur = 0
        self._token = torch.zeros(1, dtype=torch.int32, device=dev)
        self._counts = torch.empty(self.world, dtype=torch.int64, device=dev)
        self._bounds_c = (C.c_int64 * (self.world + 1))(*self.bounds)
        # device-side exchange state (fgx_peer_exchange): one zeroed mailbox per rank in symmetric memory, the epoch
        # counter and the error flag in local memory.  After this constructor no NCCL call is needed by a tick.
        self._mailbox = symm.empty(int(L.lib.fgx_peer_mailbox_bytes()), dtype=torch.uint8, device=dev)
        self._mailbox.zero_()
        self._mailbox_hdl = symm.rendezvous(self._mailbox, self.group)
        self._epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self._error = torch.zeros(1, dtype=torch.int32, device=dev)
        self._gathered = torch.zeros(2 * L.MAX_PEERS, dtype=torch.int64, device=dev)
        self._sync = L.PeerSync()
        for d in range(self.world):
            self._sync.mailbox[d] = int(self._mailbox_hdl.buffer_ptrs[d])
        self._sync.epoch, self._sync.error = self._epoch.data_ptr(), self._error.data_ptr()
        self._sync.rank, self._sync.world = self.rank, self.world
        torch.cuda.synchronize()
        dist.all_reduce(self._token, group=self.group)  # every rank's buffers exist (and mailboxes are zero) before anyone writes to them

    def _make_views(self, buf, leaves):
        out = []
        for x, row, off in zip(leaves, self._is_row, self._offsets):
            if not row:
                out.append(x)
                continue
            nbytes = x.numel() * x.element_size()
            out.append(buf[off:off + nbytes].view(x.dtype).reshape(x.shape))
        return out

    @property
    def agents(self):
        """The agent pytree over the CURRENT buffer set (in-place ops update it directly)."""
        it = iter(self._views[self.cur])
        return struct.tree_map(lambda x: next(it), self._template)

    def _peer_table(self, src, dst, part):
        arr = (L.PeerLeaf * len(part))()
        for k, (v, off, rb) in enumerate(part):
            arr[k].src = v.data_ptr()
            arr[k].row_bytes = rb
            for d in range(self.world):
                arr[k].dst[d] = self._ptrs[dst][d] + off
        return arr

    def sort(self, quantity, ascend):
        """``sort_agents(quantity, ascend, agents)`` (agent_methods.py:74-79) over the whole sharded set for a
        GENERAL float32 / int32 key: local native argsort, all-gather of every rank's sorted key images (4 B per
        slot), global positions by binary search (ties: rank order, then local order = the stable order of the
        concatenated set), rows written once to their slots on the owning GPUs.  ``quantity``: a row leaf of
        ``self.agents`` or a callable ``agents -> leaf``.  Returns the new local block."""
        from .base.agent_methods import argsort
        agents = self.agents
        q = quantity(agents) if callable(quantity) else quantity
        q = q.reshape(-1)
        if q.dtype == torch.bool:
            q = q.to(torch.int32)
        q = q.contiguous()
        n, dev = self.n_local, q.device
        perm = argsort(q, ascend)
        images = torch.full((self.n_max,), -1, dtype=torch.int32, device=dev)  # padding = 0xFFFFFFFF
        L.check(L.lib.fgx_sort_key_images(L.ptr(q), L.dtype_code(q), n, 0 if ascend else 1, L.ptr(perm), L.ptr(images),
                                          L.stream()))
        all_images = torch.empty(self.world * self.n_max, dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(all_images, images, group=self.group)
        gpos = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        src, dst = self._views[self.cur], 1 - self.cur
        rows = [(v, off, rb) for v, off, rb, row in zip(src, self._offsets, self._row_bytes, self._is_row) if row]
        for s in range(0, len(rows), L.MAX_PEER_LEAVES):
            part = rows[s:s + L.MAX_PEER_LEAVES]
            L.check(L.lib.fgx_sort_rows_peer(self._peer_table(src, dst, part), len(part), L.ptr(perm), n,
                                             L.ptr(all_images), self.n_max, self._bounds_c, self.rank, self.world,
                                             L.ptr(gpos), L.stream()))
        dist.all_reduce(self._token, group=self.group)  # orders the ranks: every scatter has landed
        self.cur = dst
        return self.agents

    def exchange(self, v0=None, v1=None):
        """All-gather of up to two int32 device scalars per rank through the mailboxes (``fgx_peer_exchange``: one
        one-warp kernel, peer stores + flags, no host, no NCCL); with no values a barrier between the ranks' streams
        that also publishes every peer store enqueued before it.  Returns int64[2, MAX_PEERS] (rows = values)."""
        L.check(L.lib.fgx_peer_exchange(C.byref(self._sync), L.ptr(v0), L.ptr(v1), L.ptr(self._gathered), L.stream()))
        return self._gathered.view(2, L.MAX_PEERS)

    def add_range(self, n_selected_local, n_active_local, n_total):
        """``distributed.global_add_range`` without NCCL or torch arithmetic: one launch (``fgx_peer_add_range``).
        Returns int32 device scalars ``(k, local_count, j0)``."""
        out = torch.empty(3, dtype=torch.int32, device=self._epoch.device)
        L.check(L.lib.fgx_peer_add_range(C.byref(self._sync), L.ptr(L.device_i32(n_selected_local)),
                                         L.ptr(L.device_i32(n_active_local)), self.n_local, self.bounds[self.rank],
                                         int(n_total), L.ptr(out), L.stream()))
        return out[0:1], out[1:2], out[2:3]

    def check(self):
        """Raises if a device-side exchange timed out (a peer never arrived); synchronises the stream."""
        e = int(self._error.item())
        if e:
            raise L.FgxError(f"peer exchange timed out waiting for rank {e - 1}")

    def pack_active(self, sync="peer"):
        """``sort_agents(active_state, descending)`` over the whole sharded set; returns the new local block.
        ``sync="peer"`` (default): the counts travel and the ranks are ordered by ``fgx_peer_exchange`` -- no NCCL
        call, so the tick can be captured in a CUDA graph on every rank; ``sync="nccl"``: the round-1 path
        (all-gather of the counts + an all-reduce as the barrier)."""
        from .base.agent_methods import select_agents
        agents = self.agents
        count, perm = select_agents(lambda a, p: P(a.active_state).nonzero(), None, agents)
        if sync == "nccl":
            dist.all_gather_into_tensor(self._counts, count.reshape(1).to(torch.int64), group=self.group)
            counts = self._counts
        else:
            counts = self.exchange(count.reshape(1))[0]
        src, dst = self._views[self.cur], 1 - self.cur
        rows = [(v, off, rb) for v, off, rb, row in zip(src, self._offsets, self._row_bytes, self._is_row) if row]
        for s in range(0, len(rows), L.MAX_PEER_LEAVES):
            part = rows[s:s + L.MAX_PEER_LEAVES]
            arr = (L.PeerLeaf * len(part))()
            for k, (v, off, rb) in enumerate(part):
                arr[k].src = v.data_ptr()
                arr[k].row_bytes = rb
                for d in range(self.world):
                    arr[k].dst[d] = self._ptrs[dst][d] + off
            L.check(L.lib.fgx_pack_rows_peer(arr, len(part), L.ptr(perm), L.ptr(count), self.n_local,
                                             L.ptr(counts), self._bounds_c, self.rank, self.world, L.stream()))
        if sync == "nccl":
            dist.all_reduce(self._token, group=self.group)  # orders the ranks: every scatter has landed
        else:
            self.exchange()  # system fence + flags: every rank's scatter has landed before anyone reads set `dst`
        self.cur = dst
        return self.agents


class PeerPositions:
    """``distributed.all_gather_positions`` without NCCL (``fgx_peer_gather_positions``): each rank stages its
    ``float4 (x, y, rad, active)`` rows in symmetric memory, the ranks meet in a mailbox barrier and every rank pulls
    all rows over NVLink into a table IN GLOBAL SLOT ORDER (so the hit indices of ``fgx_raycast_agents`` are global
    slot ids).  Three small launches on the current stream, no host sync: the whole sensing leg can be captured in a
    CUDA graph.  ``layout``: ``"block"`` (contiguous blocks, ``distributed.shard_bounds``) or ``"round_robin"``."""

    def __init__(self, n_local, layout="block", group=None, device=None):
        import torch.distributed._symmetric_memory as symm
        if layout not in ("block", "round_robin"):
            raise ValueError(f"unknown shard layout {layout!r}")
        self.group = dist.group.WORLD if group is None else group
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > L.MAX_PEERS:
            raise L.FgxError(f"at most {L.MAX_PEERS} ranks of one node")
        dev = L.device() if device is None else device
        self.layout, self.n_local = layout, int(n_local)
        sizes = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sizes, torch.tensor([self.n_local], dtype=torch.int64, device=dev), group=self.group)
        sizes = [int(v) for v in sizes.tolist()]
        self.n_total, self.n_max = sum(sizes), max(sizes)
        bounds = [0]
        for v in sizes:
            bounds.append(bounds[-1] + v)
        self._bounds_c = (C.c_int64 * (self.world + 1))(*bounds)
        self._stage = symm.empty(max(2 * self.n_max * 16, 256), dtype=torch.uint8, device=dev)
        self._stage_hdl = symm.rendezvous(self._stage, self.group)
        self._stage_c = (C.c_void_p * self.world)(*[int(p) for p in self._stage_hdl.buffer_ptrs])
        self._mailbox = symm.empty(int(L.lib.fgx_peer_mailbox_bytes()), dtype=torch.uint8, device=dev)
        self._mailbox.zero_()
        self._mailbox_hdl = symm.rendezvous(self._mailbox, self.group)
        self._epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        self._error = torch.zeros(1, dtype=torch.int32, device=dev)
        self._sync = L.PeerSync()
        for d in range(self.world):
            self._sync.mailbox[d] = int(self._mailbox_hdl.buffer_ptrs[d])
        self._sync.epoch, self._sync.error = self._epoch.data_ptr(), self._error.data_ptr()
        self._sync.rank, self._sync.world = self.rank, self.world
        self.table = torch.empty((self.n_total, 4), dtype=torch.float32, device=dev)
        token = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize()
        dist.all_reduce(token, group=self.group)  # every rank's staging buffer and zeroed mailbox exist

    def gather(self, x, y, rad, active):
        """float32[n_total, 4] rows ``(x, y, rad, active)`` of all agents in slot order (a buffer owned by this
        object, overwritten by the next call)."""
        f = lambda t: t.reshape(-1) if t.is_contiguous() else t.contiguous().reshape(-1)  # noqa: E731
        x, y, rad, active = f(x), f(y), f(rad), f(active)
        if x.shape[0] != self.n_local:
            raise L.FgxError("PeerPositions.gather: leaf length differs from the shard size given at construction")
        L.check(L.lib.fgx_peer_gather_positions(C.byref(self._sync), self._stage_c, self.n_max, L.ptr(x), L.ptr(y), L.ptr(rad),
                                                L.ptr(active), self.n_local, self.n_total,
                                                1 if self.layout == "round_robin" else 0, self._bounds_c,
                                                L.ptr(self.table), L.stream()))
        return self.table

    def check(self):
        e = int(self._error.item())
        if e:
            raise L.FgxError(f"peer position gather timed out waiting for rank {e - 1}")
