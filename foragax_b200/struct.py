"""Minimal stand-in for ``flax.struct.dataclass`` (flax is a third-party dependency of
the reference, ``pyproject.toml:20-23``, and is not part of this package's runtime).

``@struct.dataclass`` gives a frozen dataclass with ``.replace(**changes)``; the pytree
helpers below walk dataclasses, dicts, lists/tuples and ``None`` the way
``jax.tree_util`` walks the reference's ``Agent`` pytrees (``agent_methods.py:32,47,59,77``).
Leaves are device tensors (``torch.Tensor``) or Python scalars.
"""
import dataclasses

import torch


def dataclass(cls):
    cls = dataclasses.dataclass(frozen=True, eq=False)(cls)

    def replace(self, **changes):
        return dataclasses.replace(self, **changes)

    cls.replace = replace
    cls._fgx_struct = True
    return cls


def _is_struct(x):
    return dataclasses.is_dataclass(x) and not isinstance(x, type)


def tree_map(f, tree, *rest):
    """Apply ``f`` to every tensor leaf; containers are rebuilt, other leaves pass through."""
    if isinstance(tree, torch.Tensor):
        return f(tree, *rest)
    if tree is None:
        return None
    if _is_struct(tree):
        kw = {fl.name: tree_map(f, getattr(tree, fl.name), *[getattr(r, fl.name) for r in rest])
              for fl in dataclasses.fields(tree)}
        return type(tree)(**kw)
    if isinstance(tree, dict):
        return {k: tree_map(f, v, *[r[k] for r in rest]) for k, v in tree.items()}
    if isinstance(tree, (list, tuple)):
        return type(tree)(tree_map(f, v, *[r[i] for r in rest]) for i, v in enumerate(tree))
    return tree


def tree_leaves(tree):
    """All tensor leaves, in field / key order."""
    out = []
    if isinstance(tree, torch.Tensor):
        out.append(tree)
    elif tree is None:
        pass
    elif _is_struct(tree):
        for fl in dataclasses.fields(tree):
            out += tree_leaves(getattr(tree, fl.name))
    elif isinstance(tree, dict):
        for v in tree.values():
            out += tree_leaves(v)
    elif isinstance(tree, (list, tuple)):
        for v in tree:
            out += tree_leaves(v)
    return out
