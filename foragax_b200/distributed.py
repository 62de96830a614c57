"""Multi-GPU plumbing for the agent hot path: one process per GPU (``torch.distributed``, NCCL over
NVLink on the GPU box; gloo in the CPU tests).

The reference is single-process / single-device.  Sharding follows SURVEY.md
8(e): agents are split by slot index into contiguous blocks, walls and broadcast params are
replicated.  Ray-wall sensing and the fused step need no exchange; the two exchange steps are

* :func:`all_gather_positions` -- ``float4 (x, y, rad, active)`` of every agent, needed only by
  agent-agent sensing (16 B per agent);
* :func:`global_partition_offsets` -- one int32 count per rank, all-gathered, whose exclusive
  prefix sum places each rank's locally compacted (selected / unselected) slot lists inside the
  global stable partition that ``select_agents`` returns on one GPU.

Those two move 16 B/agent and 4 B/rank.  The third exchange moves agent rows:

* :func:`global_pack_active` -- ``sort_agents(active_state, descending)`` over the WHOLE sharded set:
  every rank packs its actives locally, the active counts are all-gathered, and one all-to-all per leaf
  sends each row to the rank that owns its slot of the global stable order (all actives in slot order,
  then all inactives), so that ``add_agents``' "actives are packed at the front" precondition
  (agent_methods.py:26-38) holds globally.
"""
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def shard_bounds(n_total, world_size=None, rank=None):
    """Contiguous block ``[lo, hi)`` of slot indices owned by ``rank``."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return rank * n_total // world_size, (rank + 1) * n_total // world_size


def row_leaf_mask(agents, n_local, row_mask=None):
    """Which leaves of an agent pytree are ROW leaves (slot axis first, one row per local slot) and therefore move
    when rows are exchanged.  Default: every tensor leaf whose leading dimension equals the shard size -- which can
    misfire for a replicated table that happens to have that many rows (a per-wall or parameter table; with ragged
    shards the coincidence may even differ between ranks), so callers holding such leaves pass ``row_mask`` (one
    bool per leaf, tree order) and :func:`check_row_leaves_agree` verifies that all ranks made the same choice."""
    from . import struct
    leaves = struct.tree_leaves(agents)
    if row_mask is not None:
        mask = [bool(m) for m in row_mask]
        if len(mask) != len(leaves):
            raise ValueError(f"row_mask has {len(mask)} entries for {len(leaves)} leaves")
        for m, x in zip(mask, leaves):
            if m and not (isinstance(x, torch.Tensor) and x.dim() > 0 and x.shape[0] == n_local):
                raise ValueError("row_mask marks a leaf whose leading dimension is not the shard size")
        return mask
    return [isinstance(x, torch.Tensor) and x.dim() > 0 and x.shape[0] == n_local for x in leaves]


def check_row_leaves_agree(mask, device):
    """All ranks must treat the same leaves as rows, otherwise the per-leaf collectives mismatch (or silently
    permute a replicated table on some ranks only).  One tiny all-gather; raises on disagreement."""
    r, w = world()
    if w == 1:
        return
    code = 0
    for k, m in enumerate(mask):
        code = (code * 3 + (2 if m else 1) * (k + 1)) % ((1 << 61) - 1)
    mine = torch.tensor([len(mask), code], dtype=torch.int64, device=device)
    allc = torch.empty(2 * w, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(allc, mine)
    allc = allc.view(w, 2)
    if not bool((allc == allc[0]).all()):
        raise RuntimeError("ranks disagree on which agent leaves are row leaves; pass an explicit row_mask")


def all_gather_counts(local_count):
    """int32/int64 scalar tensor per rank -> int64[world] (same device / backend as the input)."""
    r, w = world()
    c = local_count.reshape(1).to(torch.int64)
    if w == 1:
        return c
    out = torch.empty(w, dtype=torch.int64, device=c.device)
    dist.all_gather_into_tensor(out, c)
    return out


def global_partition_offsets(local_selected, n_local):
    """Where this rank's locally compacted lists go in the GLOBAL stable partition
    ("all selected slots ascending, then all unselected ascending", agent_methods.py:66-71).

    Returns ``(sel_offset, unsel_offset, total_selected)`` as int64 0-d tensors: the rank's selected
    slots occupy ``[sel_offset, sel_offset + local_selected)`` and its unselected ones
    ``[unsel_offset, unsel_offset + n_local - local_selected)`` of the global permutation.
    """
    r, w = world()
    sel = all_gather_counts(local_selected)
    tot = all_gather_counts(torch.as_tensor(n_local, device=local_selected.device))
    sel_before = sel[:r].sum()
    unsel_before = (tot[:r] - sel[:r]).sum()
    total_sel = sel.sum()
    return sel_before, total_sel + unsel_before, total_sel


def global_partition(local_perm, local_selected, lo, n_total):
    """Assemble the global permutation from every rank's local stable partition.

    ``local_perm`` int32[n_local] (local slot indices: selected ascending, then unselected),
    ``local_selected`` its selected count, ``lo`` the rank's first global slot.  Returns
    ``(total_selected, global_perm int32[n_total])`` -- identical on every rank and equal to the
    single-GPU result.  (Used when the permutation itself must be materialised; the compaction
    kernels only need :func:`global_partition_offsets`.)
    """
    r, w = world()
    n_local = local_perm.shape[0]
    so, uo, total = global_partition_offsets(local_selected, n_local)
    k = int(local_selected)
    gperm = torch.zeros(n_total, dtype=torch.int32, device=local_perm.device)
    gids = local_perm + int(lo)
    gperm[int(so):int(so) + k] = gids[:k]
    gperm[int(uo):int(uo) + n_local - k] = gids[k:]
    if w > 1:
        dist.all_reduce(gperm, op=dist.ReduceOp.SUM)  # disjoint segments: the sum is a concatenation
    return total.to(torch.int32), gperm


def all_gather_positions(x, y, rad, active, layout="block"):
    """``float4 (x, y, rad, active)`` of all agents IN GLOBAL SLOT ORDER: float32[n_total, 4] -- row ``i`` is slot
    ``i`` of the unsharded set, so the hit indices ``fgx_raycast_agents`` returns against this table are global
    slot ids (entity order of ``sim.py:402-408``).

    ``layout`` names how the slots are dealt to the ranks: ``"block"`` = contiguous blocks (rank order = slot
    order, :func:`shard_bounds`), ``"round_robin"`` = slot ``i`` on rank ``i % world`` at local row ``i // world``
    (what ``fgx_create_base_shard(first=rank, stride=world)`` builds and ``bench.py`` uses): the gathered blocks
    are then un-interleaved back to slot order."""
    if layout not in ("block", "round_robin"):
        raise ValueError(f"unknown shard layout {layout!r}")
    r, w = world()
    local = torch.stack([x.reshape(-1), y.reshape(-1), rad.reshape(-1), active.reshape(-1)], dim=1).contiguous()
    if w == 1:
        return local
    counts = all_gather_counts(torch.as_tensor(local.shape[0], device=local.device))
    m = int(counts.max())
    n_total = int(counts.sum())
    if int(counts.min()) == m:
        out = torch.empty((w * m, 4), dtype=torch.float32, device=local.device)
        dist.all_gather_into_tensor(out, local)
    else:
        # ragged shards: pad to the largest one, gather (the padding is dropped below)
        padded = torch.zeros((m, 4), dtype=torch.float32, device=local.device)
        padded[: local.shape[0]] = local
        out = torch.empty((w * m, 4), dtype=torch.float32, device=local.device)
        dist.all_gather_into_tensor(out, padded)
    if layout == "round_robin":
        # rank q holds slots q, q + w, ...: row j of rank q is slot j * w + q.  Ranks q < n_total % w hold one row
        # more, so the padded rows are exactly the slots >= n_total of the interleaved table.
        if any(int(c) != len(range(q, n_total, w)) for q, c in enumerate(counts)):
            raise ValueError("round_robin layout: shard sizes do not match slot i -> rank i % world")
        return out.view(w, m, 4).transpose(0, 1).reshape(w * m, 4)[:n_total].contiguous()
    if int(counts.min()) == m:
        return out
    return torch.cat([out[i * m: i * m + int(c)] for i, c in enumerate(counts)], dim=0)


def plan_pack_splits(active_counts, shard_sizes):
    """Integer plan of :func:`global_pack_active` (pure host logic).

    Rank s holds ``[a_s actives, n_s - a_s inactives]`` after its local stable partition.  In the global
    stable order its actives occupy positions ``[A_s, A_s + a_s)`` (``A`` = exclusive prefix of the active
    counts) and its inactives ``[A_tot + I_s, ...)``; rank d owns positions ``[B_d, B_{d+1})`` (``B`` = prefix
    of the shard sizes).  Returns ``splits[s][d]`` = rows rank s sends to rank d.  The rows for one
    destination are contiguous in the sender's packed buffer (only the LAST active piece and the FIRST
    inactive piece can share a destination, and they are adjacent), so the packed leaves are the send
    buffers as they are."""
    w = len(shard_sizes)
    a = [int(v) for v in active_counts]
    n = [int(v) for v in shard_sizes]
    bounds = [0]
    for v in n:
        bounds.append(bounds[-1] + v)
    a_tot = sum(a)

    def overlap(lo, hi, d):
        return max(0, min(hi, bounds[d + 1]) - max(lo, bounds[d]))

    splits = []
    a_before = i_before = 0
    for s_rank in range(w):
        act = (a_before, a_before + a[s_rank])
        ina = (a_tot + i_before, a_tot + i_before + n[s_rank] - a[s_rank])
        splits.append([overlap(*act, d) + overlap(*ina, d) for d in range(w)])
        a_before += a[s_rank]
        i_before += n[s_rank] - a[s_rank]
    return splits


def global_pack_active(agents, sort_active, n_active_local=None, row_mask=None):
    """``sort_agents(quantity=active_state, ascend=False)`` of a set sharded in contiguous blocks: returns
    this rank's block of the globally sorted set (bit-identical to slicing the single-GPU result).

    ``sort_active(agents) -> agents`` is the LOCAL stable "actives first" partition (on the GPU path:
    ``lambda a: jit_sort_agents(a.active_state, False, a)[0]``, i.e. the native two-valued argsort + row
    gather).  Leaves whose leading dimension is the shard size are exchanged; everything else is
    replicated state and left alone.  The split sizes have to be host integers for the collective, so
    this is the one place the counts are read back (one int64 per rank)."""
    from . import struct
    r, w = world()
    n_local = agents.active_state.shape[0]
    agents = sort_active(agents)
    if w == 1:
        return agents
    if n_active_local is None:
        n_active_local = (agents.active_state != 0).sum()
    a = all_gather_counts(torch.as_tensor(n_active_local, device=agents.active_state.device)).tolist()
    n = all_gather_counts(torch.as_tensor(n_local, device=agents.active_state.device)).tolist()
    splits = plan_pack_splits(a, n)
    send = splits[r]
    recv = [splits[s_rank][r] for s_rank in range(w)]
    assert sum(recv) == n_local and sum(send) == n_local
    mask = row_leaf_mask(agents, n_local, row_mask)
    check_row_leaves_agree(mask, agents.active_state.device)
    is_row = iter(mask)

    def exchange(leaf):
        if not next(is_row):
            return leaf
        src = leaf.contiguous()
        row = 1
        for d in src.shape[1:]:
            row *= int(d)
        # rows as raw bytes: one code path for every dtype (gloo / nccl lack bool and the unsigned types)
        flat = src.reshape(n_local, row).view(torch.uint8)
        out = torch.empty_like(flat)
        dist.all_to_all_single(out, flat, recv, send)
        return out.view(src.dtype).reshape(src.shape)

    return sort_active(struct.tree_map(exchange, agents))


def global_add_range(n_selected_local, n_active_local, n_local, lo, n_total):
    """``add_agents`` over a globally packed, block-sharded set (agent_methods.py:26-38 + the caller's clip
    ``min(n_selected, N - n_active)``, README.md:124-125): the new rows are the GLOBAL slots
    ``[a, a + k)`` with ``a`` = total active, ``k`` = clipped total of selected parents.

    Returns device scalars ``(k, local_count, j0)``: this rank adds ``local_count`` rows -- they start at its
    local active count, because :func:`global_pack_active` left its actives at the local front -- and its
    first new row is number ``j0`` of the global add order (index into the global parent list / key
    chain).  One all-gather of two counts; nothing is read back to the host."""
    r, w = world()
    dev = n_active_local.device
    pair = torch.stack([n_selected_local.reshape(()).to(torch.int64), n_active_local.reshape(()).to(torch.int64)])
    if w > 1:
        flat = torch.empty(w * 2, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(flat, pair)
        allp = flat.view(w, 2)
        sel, act = allp[:, 0].sum(), allp[:, 1].sum()
    else:
        sel, act = pair[0], pair[1]
    k = torch.minimum(sel, n_total - act)
    start = torch.clamp(act - lo, 0, n_local)
    end = torch.clamp(act + k - lo, 0, n_local)
    return k.to(torch.int32), (end - start).to(torch.int32), (start + lo - act).to(torch.int32)


def plan_parent_splits(selected_counts, active_counts, shard_sizes):
    """Integer plan of :func:`fetch_parent_rows`: ``splits[s][d]`` = parent rows rank s sends to rank d.
    Rank s holds the parents number ``[S_s, S_s + c_s)`` of the global selection (``S`` = exclusive prefix
    of the selected counts); parent j seeds the new row at global slot ``a + j`` (``j < k``), owned by the
    rank whose block contains it.  Both maps are monotone in j, so every (s, d) piece is one contiguous run."""
    w = len(shard_sizes)
    c = [int(v) for v in selected_counts]
    n = [int(v) for v in shard_sizes]
    a = sum(int(v) for v in active_counts)
    k = min(sum(c), sum(n) - a)
    bounds = [0]
    for v in n:
        bounds.append(bounds[-1] + v)
    splits, before = [], 0
    for s_rank in range(w):
        j_lo, j_hi = min(before, k), min(before + c[s_rank], k)
        splits.append([max(0, min(a + j_hi, bounds[d + 1]) - max(a + j_lo, bounds[d])) for d in range(w)])
        before += c[s_rank]
    return splits


def fetch_parent_rows(agents, selected_ids, n_selected_local, n_active_local, take_rows, n_local=None):
    """Parent rows for this rank's new slots: every rank gathers the rows of its selected
    parents (``take_rows(agents, ids) -> agents`` = the native row gather; ``selected_ids`` is the local select
    permutation, parents first), and one all-to-all per leaf delivers them to the ranks that own the new
    slots ``[a, a + k)``.  Returns a pytree whose row leaves hold, in order, the parent of each of this rank's
    new rows (``global_add_range``'s ``local_count`` rows).  Split sizes are host integers (one read-back of
    2 counts per rank)."""
    from . import struct
    r, w = world()
    if n_local is None:  # `agents` may be any pytree of row leaves when n_local is given
        n_local = agents.active_state.shape[0]
    dev = selected_ids.device
    trio = torch.stack([torch.as_tensor(n_selected_local, device=dev).reshape(()).to(torch.int64),
                        torch.as_tensor(n_active_local, device=dev).reshape(()).to(torch.int64),
                        torch.as_tensor(n_local, device=dev, dtype=torch.int64)])
    if w > 1:
        flat3 = torch.empty(w * 3, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(flat3, trio)
        allt = flat3.view(w, 3)
    else:
        allt = trio.reshape(1, 3)
    sel, act, sizes = (allt[:, i].tolist() for i in range(3))
    splits = plan_parent_splits(sel, act, sizes)
    send = splits[r]
    recv = [splits[s_rank][r] for s_rank in range(w)]
    n_send, n_recv = sum(send), sum(recv)
    # parents this rank contributes: the first n_send of its selected list (the clip drops the global tail)
    skip = sum(sel[:r])
    a_tot, k = sum(act), min(sum(sel), sum(sizes) - sum(act))
    assert n_send == max(0, min(skip + sel[r], k) - min(skip, k))
    parents = take_rows(agents, selected_ids[:n_send].contiguous())

    def exchange(leaf):
        if not isinstance(leaf, torch.Tensor) or leaf.dim() == 0 or leaf.shape[0] != n_send:
            return leaf
        src = leaf.contiguous()
        row = 1
        for d in src.shape[1:]:
            row *= int(d)
        flat = src.reshape(n_send, row).view(torch.uint8)
        out = torch.empty((n_recv, flat.shape[1]), dtype=torch.uint8, device=src.device)
        if w > 1:
            dist.all_to_all_single(out, flat, recv, send)
        else:
            out.copy_(flat)
        return out.view(src.dtype).reshape((n_recv,) + tuple(src.shape[1:]))

    return struct.tree_map(exchange, parents)


def key_chain_advance(mode, key, steps):
    """The key ``add_agents``' serial chain holds after ``steps`` added rows (``fgx_key_chain_advance``; mode 0 =
    Forager, 1 = Resource).  ``key`` uint32[2] device tensor, ``steps`` device scalar or int."""
    from . import _lib as L
    out = torch.empty_like(key)
    st = L.device_i32(steps)
    L.check(L.lib.fgx_key_chain_advance(int(mode), L.ptr(key), L.ptr(st), L.ptr(out), L.stream()))
    return out


def global_add_agents(add_func, make_add_params, agents, parent_leaves, selected_ids, n_selected_local, key, lo,
                      n_total, chain_mode):
    """``add_agents`` (agent_methods.py:26-38) + the caller's clip over a globally packed, block-sharded set, for
    add functions that copy from PARENT rows and thread a key (Forager / Resource, sim.py:154-210,316-344).

    The new rows are the global slots ``[a, a + k)``; child j descends from the j-th selected agent of the GLOBAL
    select order and uses link j of the serial key chain.  Per rank: :func:`global_add_range` gives the local
    part of the range and the global index ``j0`` of its first child; :func:`fetch_parent_rows` brings the rows
    of the leaves the add function READS from a parent (``parent_leaves(agents)`` -> list of row leaves, all of
    which it overwrites in the child) and they are written INTO the new slots, so the unmodified native add runs
    with ``good_ids = identity`` (its kernels read a parent element before they overwrite it) from the chain key
    advanced by ``j0`` links.  Returns ``(agents, key_out, n_parents_local)``: ``key_out`` is the key after all k
    links (the same on every rank) and the first ``n_parents_local`` of ``selected_ids`` are this rank's parents
    that did reproduce (what the caller's ``set_agents`` applies to)."""
    from .base.agent_methods import add_agents, count_active, take_rows
    n_local = agents.active_state.shape[0]
    dev = agents.active_state.device
    n_act = count_active(agents)
    k, local_count, j0 = global_add_range(n_selected_local, n_act, n_local, lo, n_total)
    leaves = list(parent_leaves(agents))
    parents = fetch_parent_rows(leaves, selected_ids, n_selected_local, n_act, take_rows, n_local=n_local)
    n_new = int(parents[0].shape[0]) if parents else 0
    start = int(n_act)  # the counts were read back by fetch_parent_rows already
    if n_new:
        for dst, src in zip(leaves, parents):
            dst[start:start + n_new].copy_(src)
    ids = torch.arange(start, start + max(n_new, 1), dtype=torch.int32, device=dev)
    key_r = key_chain_advance(chain_mode, key, j0)
    agents, _ = add_agents(add_func, local_count, make_add_params(ids), agents, key_r)
    # parents of this rank that reproduced: its selected agents whose global index is below k
    r, w = world()
    sel_all = all_gather_counts(torch.as_tensor(n_selected_local, device=dev).reshape(()))
    before = sel_all[:r].sum()
    n_parents_local = torch.minimum(torch.clamp(k.to(torch.int64) - before, min=0), sel_all[r]).to(torch.int32)
    return agents, key_chain_advance(chain_mode, key, k), n_parents_local
