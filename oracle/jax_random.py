"""NumPy restatement of the slice of ``jax.random`` the reference's hot path uses.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  The algorithm lives in the
third-party dependency ``jax`` (``pyproject.toml:20-23`` pins only
``jax>=0.4.13``; absent from ``/root/reference`` and from this image), so this
restates its published algorithm -- threefry2x32 (Salmon et al., Random123)
with JAX's *legacy* counter layout (``jax_threefry_partitionable=False``, the
default before JAX 0.5) -- and is pinned by the Random123 known-answer vectors
and by the reference notebook output G1, which only this layout reproduces.

Reference call sites: ``foragax/base/agent_methods.py:11`` (split),
``foragax/policy/random_policy.py:19-20`` (split, uniform),
``foragax/policy/wilson_cowan.py:19-26`` (split, uniform with bounds),
``README.md:39,42,59,61,75-76`` (split, randint),
``examples/.../EcoEvoJax/sim.py:28,37-49,181-199,234-235,247-250,284-285,322-336``
(split(k,n), uniform, bernoulli, normal).

All functions take a *batch* of keys ``uint32[..., 2]`` (the vmapped form the
reference produces) or a single key ``uint32[2]``.
"""
import numpy as np

U32 = np.uint32
_ROT = ((13, 15, 26, 6), (17, 29, 16, 24))


def PRNGKey(seed):
    """``jax.random.PRNGKey`` with x64 disabled: ``[0, seed]`` (uint32)."""
    seed = int(seed)
    return np.array([(seed >> 32) & 0xFFFFFFFF, seed & 0xFFFFFFFF], dtype=U32)


def _rotl(x, r):
    return (x << U32(r)) | (x >> U32(32 - r))


def threefry2x32(k0, k1, x0, x1):
    """One threefry2x32 block, 20 rounds, element-wise over uint32 arrays."""
    with np.errstate(over="ignore"):
        k0 = np.asarray(k0, dtype=U32)
        k1 = np.asarray(k1, dtype=U32)
        x0 = np.asarray(x0, dtype=U32).copy()
        x1 = np.asarray(x1, dtype=U32).copy()
        ks = (k0, k1, k0 ^ k1 ^ U32(0x1BD11BDA))
        x0 = x0 + ks[0]
        x1 = x1 + ks[1]
        for g in range(5):
            for r in _ROT[g & 1]:
                x0 = x0 + x1
                x1 = _rotl(x1, r)
                x1 = x1 ^ x0
            x0 = x0 + ks[(g + 1) % 3]
            x1 = x1 + ks[(g + 2) % 3] + U32(g + 1)
    return x0, x1


def _bits(keys, n):
    """Legacy ``threefry_2x32(key, iota(n))`` for a batch of keys.

    keys: uint32[..., 2] -> uint32[..., n].  Odd n is padded with one zero
    counter, the counters are split into halves (x0 = first half, x1 = second
    half) and the outputs concatenated, exactly as ``jax._src.prng``.
    """
    keys = np.asarray(keys, dtype=U32)
    lead = keys.shape[:-1]
    h = (n + 1) // 2
    cnt = np.arange(2 * h, dtype=U32)
    if n % 2:
        cnt[-1] = 0  # the pad element is a literal zero, not the next iota value
    x0 = np.broadcast_to(cnt[:h], lead + (h,))
    x1 = np.broadcast_to(cnt[h:], lead + (h,))
    o0, o1 = threefry2x32(keys[..., 0:1], keys[..., 1:2], x0, x1)
    return np.concatenate([o0, o1], axis=-1)[..., :n]


def split(keys, num=2):
    """``jax.random.split``: uint32[..., 2] -> uint32[..., num, 2]."""
    keys = np.asarray(keys, dtype=U32)
    flat = _bits(keys, 2 * num)
    return flat.reshape(keys.shape[:-1] + (num, 2))


def random_bits(keys, n):
    """``jax.random.bits(key, (n,))`` (32-bit): uint32[..., 2] -> uint32[..., n]."""
    return _bits(keys, int(n))


def uniform(keys, n=1, minval=0.0, maxval=1.0):
    """``jax.random.uniform(key, (n,), float32, minval, maxval)``.

    minval / maxval broadcast against the output ``[..., n]``.
    """
    bits = random_bits(keys, n)
    fl = ((bits >> U32(9)) | U32(0x3F800000)).view(np.float32) - np.float32(1.0)
    lo = np.asarray(minval, dtype=np.float32)
    hi = np.asarray(maxval, dtype=np.float32)
    out = (fl * (hi - lo)).astype(np.float32) + lo
    return np.maximum(lo, out.astype(np.float32)).astype(np.float32)


def randint(keys, n, minval, maxval):
    """``jax.random.randint(key, (n,), minval, maxval)`` (int32 scalar bounds)."""
    k = split(keys, 2)
    hb = random_bits(k[..., 0, :], n)
    lb = random_bits(k[..., 1, :], n)
    minval = int(minval)
    maxval = int(maxval)
    span = U32((maxval - minval) & 0xFFFFFFFF) if maxval > minval else U32(1)
    with np.errstate(over="ignore"):
        m = U32(65536) % span
        m = (m * m) % span
        off = (hb % span) * m + (lb % span)
        off = off % span
    return (np.int32(minval) + off.astype(np.int32)).astype(np.int32)


def bernoulli(keys, p, n=1):
    """``jax.random.bernoulli(key, p)`` with ``p`` of shape ``[..., n]``."""
    return uniform(keys, n) < np.asarray(p, dtype=np.float32)


def erf_inv_f32(x):
    """XLA's float32 ``erf_inv`` (Giles' single-precision polynomial)."""
    x = np.asarray(x, dtype=np.float32)
    w = (-np.log1p((-x * x).astype(np.float32))).astype(np.float32)
    lt = w < np.float32(5.0)
    w1 = (w - np.float32(2.5)).astype(np.float32)
    w2 = (np.sqrt(np.maximum(w, 0)) - np.float32(3.0)).astype(np.float32)
    c1 = [2.81022636e-08, 3.43273939e-07, -3.5233877e-06, -4.39150654e-06,
          0.00021858087, -0.00125372503, -0.00417768164, 0.246640727, 1.50140941]
    c2 = [-0.000200214257, 0.000100950558, 0.00134934322, -0.00367342844,
          0.00573950773, -0.0076224613, 0.00943887047, 1.00167406, 2.83297682]
    ww = np.where(lt, w1, w2).astype(np.float32)
    p = np.where(lt, np.float32(c1[0]), np.float32(c2[0])).astype(np.float32)
    for a, b in zip(c1[1:], c2[1:]):
        p = (np.where(lt, np.float32(a), np.float32(b)) + p * ww).astype(np.float32)
    out = (p * x).astype(np.float32)
    return np.where(np.abs(x) == 1, np.float32(np.inf) * x, out).astype(np.float32)


def normal(keys, n=1):
    """``jax.random.normal(key, (n,))`` float32 (tolerance-only: erf_inv polynomial)."""
    lo = np.nextafter(np.float32(-1.0), np.float32(0.0))
    u = uniform(keys, n, lo, np.float32(1.0))
    return (np.float32(np.sqrt(2)) * erf_inv_f32(u)).astype(np.float32)
