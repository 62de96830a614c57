"""NumPy restatement of ``foragax/base/agent_methods.py:9-79``.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Agents are nested dicts of
NumPy arrays whose leading axis is the slot axis (the struct-of-arrays pytree
``jax.vmap`` produces in the reference).  Per-agent callbacks (``add_func``,
``remove_func``, ``set_func``, ``select_func``) are plain Python functions over
those dicts, applied one row at a time in the same serial order as the
reference's ``lax.fori_loop`` bodies.
"""
import numpy as np

from . import jax_random as jr


# --- pytree helpers (nested dict / None leaves) ------------------------------
def tree_map(f, tree, *rest):
    if tree is None:
        return None
    if isinstance(tree, dict):
        return {k: tree_map(f, tree[k], *[r[k] for r in rest]) for k in tree}
    return f(tree, *rest)


def tree_leaves(tree):
    if tree is None:
        return []
    if isinstance(tree, dict):
        out = []
        for k in tree:
            out += tree_leaves(tree[k])
        return out
    return [tree]


def tree_row(tree, i):
    """``tree_map(lambda x: x[i], agents)`` with jnp's clamp-on-gather rule."""
    def get(x):
        j = int(i)
        if j < 0:
            j += x.shape[0]
        j = min(max(j, 0), x.shape[0] - 1)
        return x[j].copy() if isinstance(x[j], np.ndarray) else x[j]
    return tree_map(get, tree)


def _set_row(x, i, y):
    """``x.at[i].set(y)``: cast to the leaf dtype (float->int truncates), drop OOB."""
    i = int(i)
    if i < 0:
        i += x.shape[0]
    if 0 <= i < x.shape[0]:
        x[i] = np.asarray(y).astype(x.dtype)
    return x


# --- argsort with lax.sort's float total order --------------------------------
def _sort_key_bits(q):
    """Monotone uint32 image of an int32 / float32 key under ``lax.sort``'s order.

    Floats: -0.0 -> +0.0, every NaN sorts last (after +inf); ints: bias by 2^31.
    """
    q = np.asarray(q)
    if q.dtype.kind == "f":
        q = q.astype(np.float32)
        nan = np.isnan(q)
        q = np.where(q == 0, np.float32(0.0), q)
        b = q.view(np.uint32)
        neg = (b >> np.uint32(31)).astype(bool)
        m = np.where(neg, ~b, b | np.uint32(0x80000000)).astype(np.uint32)
        return np.where(nan, np.uint32(0xFFFFFFFF), m).astype(np.uint32)
    if q.dtype.kind == "b":
        return q.astype(np.uint32)
    return (q.astype(np.int64) + (1 << 31)).astype(np.uint32)


def argsort(q):
    """Stable ``jnp.argsort`` of a 1-D int32 / float32 / bool key -> int32[N]."""
    return np.argsort(_sort_key_bits(q), kind="stable").astype(np.int32)


def _neg(q):
    q = np.asarray(q)
    if q.dtype.kind == "f":
        return (-q.astype(np.float32)).astype(np.float32)
    if q.dtype.kind == "b":
        raise TypeError("cannot negate a bool key")
    with np.errstate(over="ignore"):
        return (-q.astype(np.int32)).astype(np.int32)  # two's complement wrap, as XLA


# --- agent_methods.py:9-17 ----------------------------------------------------
def create_agents(create_agent, params, num_total, num_active, agent_type, key):
    """``create_agents`` (agent_methods.py:9-17) with the vmapped user
    ``create_agent(params, unique_ids[N], active_states[N], agent_types[N], keys[N,2])``
    supplied as a batched NumPy function."""
    uniq_ids = np.arange(num_total, dtype=np.int32)
    create_keys = jr.split(key, num_total + 1)[1:]
    active = np.concatenate([np.ones(num_active, np.float32),
                             np.zeros(num_total - num_active, np.float32)])
    types = np.full(num_total, agent_type, dtype=np.int32)
    return create_agent(params, uniq_ids, active, types, create_keys)


# --- agent_methods.py:26-38 ---------------------------------------------------
def add_agents(add_func, num_agents_add, add_params, agents, key):
    a = int(np.sum(agents["active_state"], dtype=np.float32).astype(np.int32))
    agents = tree_map(lambda x: x.copy(), agents)
    for idx in range(a, a + int(num_agents_add)):
        new_agent, key = add_func(add_params, agents, idx, key)
        tree_map(lambda x, y: _set_row(x, idx, y), agents, new_agent)
    return agents, key


# --- agent_methods.py:41-52 ---------------------------------------------------
def remove_agents(remove_func, num_agents_remove, remove_params, agents):
    agents = tree_map(lambda x: x.copy(), agents)
    ids = remove_params["remove_ids"]
    for j in range(int(num_agents_remove)):
        i = int(ids[min(j, len(ids) - 1)])
        new_agent = remove_func(remove_params, agents, i)
        tree_map(lambda x, y: _set_row(x, i, y), agents, new_agent)
    return agents


# --- agent_methods.py:54-63 ---------------------------------------------------
def set_agents(set_func, num_agents_set, set_params, agents, key):
    agents = tree_map(lambda x: x.copy(), agents)
    ids = set_params["set_ids"]
    for j in range(int(num_agents_set)):
        i = int(ids[min(j, len(ids) - 1)])
        new_agent, key = set_func(set_params, agents, i, key)
        tree_map(lambda x, y: _set_row(x, i, y), agents, new_agent)
    return agents, key


# --- agent_methods.py:66-72 ---------------------------------------------------
def select_agents(select_func, select_params, agents):
    mask = np.asarray(select_func(agents, select_params))
    s = np.where(mask, np.float32(1.0), np.float32(0.0)).reshape(-1)
    order = argsort((np.float32(-1.0) * s).astype(np.float32))
    return np.int32(np.sum(s, dtype=np.float32)), order


def select_from_mask(mask):
    """The data part of ``select_agents`` for an already-evaluated predicate."""
    return select_agents(lambda a, p: mask, None, None)


# --- agent_methods.py:74-79 ---------------------------------------------------
def sort_ids(quantity, ascend):
    q = np.asarray(quantity).reshape(-1)
    return argsort(q) if ascend else argsort(_neg(q))


def sort_agents(quantity, ascend, agents):
    ids = sort_ids(quantity, ascend)
    return tree_map(lambda x: np.take(x, ids, axis=0), agents), ids
