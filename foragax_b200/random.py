"""Device ``jax.random`` look-alike over the K4 threefry kernels (``csrc/rng.cu``).

Replaces the third-party ``jax.random`` calls on the reference's hot path
(``agent_methods.py:11``, ``random_policy.py:19-20``, ``wilson_cowan.py:19-26``,
``README.md:39,42``); results are bit-identical to JAX's legacy threefry layout.
Keys are ``uint32[2]`` (or batches ``uint32[..., 2]``) device tensors; a batch of keys is
processed "vmapped", each key drawing ``shape`` samples.
"""
import numpy as np
import torch

from . import _lib as L


def PRNGKey(seed):
    """``jax.random.PRNGKey`` (x64 off): ``[hi32(seed), lo32(seed)]`` -> ``[0, seed]``."""
    seed = int(seed)
    host = np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=np.uint32)
    return torch.from_numpy(host).to(L.device())


def key_to_host(key):
    """uint32[2] key (device tensor / numpy / list) -> (k0, k1) Python ints."""
    if isinstance(key, torch.Tensor):
        key = key.detach().cpu().numpy()
    key = np.asarray(key, dtype=np.uint32).reshape(-1)
    return int(key[0]), int(key[1])


def _keys2d(key):
    if key.dtype != torch.uint32:
        raise L.FgxError("PRNG keys must be uint32")
    k = key.contiguous()
    return k.reshape(-1, 2), tuple(key.shape[:-1])


def _n(shape):
    if isinstance(shape, int):
        shape = (shape,)
    n = 1
    for s in shape:
        n *= int(s)
    return n, tuple(int(s) for s in shape)


def split(key, num=2):
    k, lead = _keys2d(key)
    out = torch.empty((k.shape[0], num, 2), dtype=torch.uint32, device=k.device)
    L.check(L.lib.fgx_threefry_split(L.ptr(k), k.shape[0], num, L.ptr(out), L.stream()))
    return out.reshape(lead + (num, 2))


def bits(key, shape=()):
    k, lead = _keys2d(key)
    n, shape = _n(shape)
    out = torch.empty((k.shape[0], n), dtype=torch.uint32, device=k.device)
    L.check(L.lib.fgx_threefry_random_bits(L.ptr(k), k.shape[0], n, L.ptr(out), L.stream()))
    return out.reshape(lead + shape)


def uniform(key, shape=(), minval=0.0, maxval=1.0):
    k, lead = _keys2d(key)
    n, shape = _n(shape)
    out = torch.empty((k.shape[0], n), dtype=torch.float32, device=k.device)
    L.check(L.lib.fgx_threefry_uniform(L.ptr(k), k.shape[0], n, float(minval), float(maxval), L.ptr(out),
                                       L.stream()))
    return out.reshape(lead + shape)


def randint(key, shape, minval, maxval):
    k, lead = _keys2d(key)
    n, shape = _n(shape)
    out = torch.empty((k.shape[0], n), dtype=torch.int32, device=k.device)
    L.check(L.lib.fgx_threefry_randint(L.ptr(k), k.shape[0], n, int(minval), int(maxval), L.ptr(out), L.stream()))
    return out.reshape(lead + shape)


def normal(key, shape=()):
    k, lead = _keys2d(key)
    n, shape = _n(shape)
    out = torch.empty((k.shape[0], n), dtype=torch.float32, device=k.device)
    L.check(L.lib.fgx_threefry_normal(L.ptr(k), k.shape[0], n, L.ptr(out), L.stream()))
    return out.reshape(lead + shape)
