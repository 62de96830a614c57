"""Agent manipulation -- the B200 mirror of ``foragax/base/agent_methods.py:9-79``.

Same function names, argument order and return values as the reference
(``create_agents, step_agents, (jit_)add_agents, (jit_)remove_agents, (jit_)set_agents,
(jit_)select_agents, (jit_)sort_agents``).  Every function enqueues CUDA kernels of
``libforagax_b200.so`` on the current stream and returns without synchronising; counts
(``num_agents_*``, the count returned by ``select_agents``) are int32 *device* scalars so a
select -> remove -> sort -> select -> add chain never round-trips to the host.

Where the reference runs a serial ``lax.fori_loop`` over a per-agent callback
(``agent_methods.py:29-35,43-49,55-61``) the native op processes all ``k`` rows in one
launch; it receives the rows as a :class:`RowRange` / :class:`RowIds` instead of one
``idx`` at a time.  Rows touched by one call are distinct (they come from a permutation),
so the parallel result equals the serial one -- with ONE precondition for ``add_agents`` with an
add function that copies from parent rows (``Forager.add_agent`` / ``Resource.add_agent``,
``sim.py:158,320``): every parent slot ``good_ids[j]`` must lie BELOW the first new slot
``a = sum(active_state)``.  The reference's serial loop would let a later child copy a row an
earlier iteration already overwrote; the parallel kernels read parents while other lanes write
children, so a parent inside ``[a, a + k)`` is a data race here.  The shipped examples meet it
by construction (parents are selected among the actives, which the preceding descending sort
packs below ``a``); ``check_add_parents`` verifies it on demand (host sync; ``FGX_CHECK_ADD=1``
makes the example add functions call it).

Buffers are updated in place where the reference would return an updated copy (the usual
``set.agents = step_agents(..., set)`` pattern drops the old value); ``sort_agents``
gathers into fresh buffers.
"""
import ctypes as C

import torch

from .. import _lib as L
from .. import random as fgx_random
from .. import struct
from .agent_classes import Agent, Agent_Set, Params, Policy, Signal, State, native, require_native  # noqa: F401
from .predicates import P, Pred, as_pred  # noqa: F401


class RowRange:
    """Rows ``[start, start+count)``; both are int32 device scalars (add_agents)."""

    def __init__(self, start, count):
        self.start, self.count = start, count


class RowIds:
    """Rows ``ids[0:count]``; ``count`` is an int32 device scalar (remove / set)."""

    def __init__(self, ids, count):
        self.ids, self.count = ids, count


def num_slots(agents):
    return int(agents.unique_id.shape[0])


# agent_methods.py:9-17
def create_agents(params, agent_set, key, shard=None):
    """``shard=(first, n_local[, stride])`` creates only the slots first, first+stride, ... of the set
    (multi-GPU: every rank builds exactly its rows of the single-GPU result); default: the whole set."""
    n = int(agent_set.num_total_agents)
    n_active = int(agent_set.num_active_agents)
    first, n_local = (0, n) if shard is None else (int(shard[0]), int(shard[1]))
    stride = 1 if shard is None or len(shard) < 3 else int(shard[2])
    dev = L.device()
    k0, k1 = fgx_random.key_to_host(key)
    key_host = (C.c_uint32 * 2)(k0, k1)
    uniq_ids = torch.empty(n_local, dtype=torch.int32, device=dev)
    create_keys = torch.empty((n_local, 2), dtype=torch.uint32, device=dev)
    active_states = torch.empty(n_local, dtype=torch.float32, device=dev)
    agent_types = torch.empty(n_local, dtype=torch.int32, device=dev)
    L.check(L.lib.fgx_create_base_shard(key_host, n, n_active, first, stride, n_local, int(agent_set.agent_type),
                                        L.ptr(uniq_ids), L.ptr(create_keys), L.ptr(active_states),
                                        L.ptr(agent_types), L.stream()))
    return agent_set.create_agents(params, uniq_ids, active_states, agent_types, create_keys)


# agent_methods.py:20-24
def step_agents(params, input, agent_set):
    try:
        return agent_set.step_agents(params, input, agent_set.agents)
    except ValueError:
        print("Error in step_agents")


def count_active(agents):
    """``jnp.sum(agents.active_state, dtype=jnp.int32)`` as an int32[1] device tensor."""
    a = agents.active_state
    out = torch.empty(1, dtype=torch.int32, device=a.device)
    L.check(L.lib.fgx_count_active(L.ptr(a), a.shape[0], L.ptr(out), L.stream()))
    return out


def check_add_parents(good_ids, rows):
    """Raises unless every parent of an ``add_agents`` call lies below the first new row (see the module
    docstring).  Reads three scalars back: debugging aid, not for the hot path."""
    a, k = int(rows.start.reshape(-1)[0]), int(rows.count.reshape(-1)[0])
    if k <= 0:
        return
    worst = int(good_ids.reshape(-1)[:k].max())
    if worst >= a:
        raise L.FgxError(f"add_agents: parent slot {worst} lies inside the new rows [{a}, {a + k}): the parallel add "
                         "would race where the reference's serial loop re-reads an overwritten row")


def count_active_clip(agents, num_agents_add):
    """``(n_active, min(num_agents_add, N - n_active))`` as int32[1] device tensors in ONE launch: the count of
    ``add_agents`` (agent_methods.py:27) and the caller's clip to the free slots (README.md:124-125,
    hello_world.ipynb:138-139, sim.py:520-521)."""
    a = agents.active_state
    out = torch.empty(3, dtype=torch.int32, device=a.device)
    k = L.device_i32(num_agents_add)
    L.check(L.lib.fgx_count_active_clip(L.ptr(a), a.shape[0], L.ptr(k), L.ptr(out), L.stream()))
    return out[0:1], out[1:2]


# agent_methods.py:26-38
def add_agents(add_func, num_agents_add, add_params, agents, key, num_active=None):
    """``num_active`` (optional, not in the reference signature): ``jnp.sum(agents.active_state)`` as an int32
    device scalar when the caller already holds it (the clip before an add computes it, README.md:124-125);
    otherwise it is counted here as the reference does (agent_methods.py:27)."""
    require_native(add_func, "add_func")
    id_last_active = count_active(agents) if num_active is None else L.device_i32(num_active)
    rows = RowRange(id_last_active, L.device_i32(num_agents_add))
    new_agents, key = add_func(add_params, agents, rows, key)
    return new_agents, key


jit_add_agents = add_agents


# agent_methods.py:41-52
def remove_agents(remove_func, num_agents_remove, remove_params, agents):
    require_native(remove_func, "remove_func")
    rows = RowIds(remove_params.content['remove_ids'], L.device_i32(num_agents_remove))
    return remove_func(remove_params, agents, rows)


jit_remove_agents = remove_agents


# agent_methods.py:54-63
def set_agents(set_func, num_agents_set, set_params, agents, key):
    require_native(set_func, "set_func")
    rows = RowIds(set_params.content['set_ids'], L.device_i32(num_agents_set))
    new_agents, key = set_func(set_params, agents, rows, key)
    return new_agents, key


jit_set_agents = set_agents


# agent_methods.py:66-72
def select_agents(select_func, select_params, agents):
    """Returns ``(selected_ids_len int32[], sort_selected_ids int32[N])``: the stable
    partition "selected ascending, then unselected ascending" (== ``argsort(-mask)``)."""
    pred = as_pred(select_func(agents, select_params))
    if isinstance(agents, Agent) or hasattr(agents, "unique_id"):
        n = num_slots(agents)
    else:
        n = int(pred.terms[0].tensor.shape[0])
    dev = L.device()
    pstruct, keep = pred.to_struct(n)
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    count = torch.empty(1, dtype=torch.int32, device=dev)
    nbytes = L.lib.fgx_select_scratch_bytes(n)
    scr = L.scratch(nbytes)
    L.check(L.lib.fgx_select(C.byref(pstruct), n, L.ptr(perm), L.ptr(count), L.ptr(scr), nbytes, L.stream()))
    del keep
    return count.reshape(()), perm


jit_select_agents = select_agents


def argsort(quantity, ascend=True):
    """Stable ``jnp.argsort(q)`` / ``jnp.argsort(-q)`` of a float32 / int32 key -> int32[N]."""
    q = quantity.reshape(-1)
    if q.dtype == torch.bool:
        q = q.to(torch.int32)
    if not q.is_contiguous():
        q = q.contiguous()
    n = q.shape[0]
    if isinstance(ascend, torch.Tensor):
        ascend = bool(ascend.item())
    perm = torch.empty(n, dtype=torch.int32, device=q.device)
    nbytes = L.lib.fgx_argsort_scratch_bytes(n)
    scr = L.scratch(nbytes)
    L.check(L.lib.fgx_argsort(L.ptr(q), L.dtype_code(q), n, 0 if ascend else 1, L.ptr(perm), L.ptr(scr), nbytes,
                              L.stream()))
    return perm


def _leaf_table(agents, n_out, out=None):
    """(leaves, destination buffers of n_out rows, [(src, dst, row_bytes)]) of an agent pytree; the destinations are
    fresh unless ``out`` (a pytree of the same structure holding contiguous n_out-row buffers) supplies them."""
    leaves = struct.tree_leaves(agents)
    n_src = leaves[0].shape[0] if leaves else 0
    given = struct.tree_leaves(out) if out is not None else None
    if given is not None and len(given) != len(leaves):
        raise L.FgxError("out= must have the structure of the agent pytree")
    outs, table = [], []
    for k, x in enumerate(leaves):
        if x.shape[0] != n_src:
            raise L.FgxError("every agent leaf must have the slot axis first")
        src = x if x.is_contiguous() else x.contiguous()
        shape = (n_out,) + tuple(src.shape[1:])
        if given is None:
            dst = torch.empty(shape, dtype=src.dtype, device=src.device)
        else:
            dst = given[k]
            if (tuple(dst.shape) != shape or dst.dtype != src.dtype or not dst.is_contiguous()
                    or dst.data_ptr() == src.data_ptr()):
                raise L.FgxError("out= leaves must be contiguous buffers of the sorted shape, distinct from the input")
        outs.append(dst)
        row_bytes = src.element_size() * (src.numel() // n_src) if n_src else 0
        table.append((src, dst, row_bytes))
    return n_src, outs, table


def _leaf_array(part):
    arr = (L.Leaf * len(part))()
    for k, (src, dst, rb) in enumerate(part):
        arr[k].src, arr[k].dst, arr[k].row_bytes = L.ptr(src), L.ptr(dst), rb
    return arr


def take_rows(agents, ids):
    """``tree_map(lambda x: jnp.take(x, ids, axis=0), agents)`` in one launch per 32 leaves; ``ids`` may be shorter
    (or longer) than the slot axis: the result has ``len(ids)`` rows."""
    n = ids.shape[0]
    n_src, outs, table = _leaf_table(agents, n)
    if n and n_src:
        for s in range(0, len(table), L.MAX_LEAVES):
            part = table[s:s + L.MAX_LEAVES]
            L.check(L.lib.fgx_gather_rows_from(_leaf_array(part), len(part), L.ptr(ids), n, n_src, L.stream()))
    it = iter(outs)
    return struct.tree_map(lambda x: next(it), agents)


# agent_methods.py:74-79
def sort_agents(quantity, ascend, agents, out=None):
    """argsort + ``take`` of every leaf as ONE native call (``fgx_sort_rows``): a two-valued key (the sort on
    ``active_state``) is partitioned and its rows are moved by a single kernel that reads the key once.
    ``out`` (optional, beyond the reference signature): an agent pytree of preallocated destination buffers -- a
    world that alternates between two buffer sets returns to its first set every second tick, which is what lets
    a pair of ticks be captured as one CUDA graph."""
    q = quantity.reshape(-1)
    if q.dtype == torch.bool:
        q = q.to(torch.int32)
    if not q.is_contiguous():
        q = q.contiguous()
    n = q.shape[0]
    n_src, outs, table = _leaf_table(agents, n, out)
    if len(table) > L.MAX_LEAVES or n_src != n:
        if out is not None:
            raise L.FgxError("sort_agents(out=): the fused row move handles at most %d leaves" % L.MAX_LEAVES)
        sorted_ids = argsort(quantity, ascend)
        return take_rows(agents, sorted_ids), sorted_ids
    if isinstance(ascend, torch.Tensor):
        ascend = bool(ascend.item())
    perm = torch.empty(n, dtype=torch.int32, device=q.device)
    nbytes = L.lib.fgx_argsort_scratch_bytes(n)
    scr = L.scratch(nbytes)
    L.check(L.lib.fgx_sort_rows(L.ptr(q), L.dtype_code(q), n, 0 if ascend else 1, L.ptr(perm), _leaf_array(table),
                                len(table), L.ptr(scr), nbytes, L.stream()))
    it = iter(outs)
    return struct.tree_map(lambda x: next(it), agents), perm


jit_sort_agents = sort_agents
