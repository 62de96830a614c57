"""Select predicates evaluated inside the partition kernel.

The reference's ``select_func(agents, params) -> bool[N]`` (``agent_methods.py:67``) is
traced JAX code.  Here a ``select_func`` returns either

* a :class:`Pred` built from leaf comparisons -- ``P(leaf) == 1``, ``&``, ``|``, ``~`` --
  which ``fgx_select`` evaluates in-kernel straight from the agent leaves (up to four
  comparisons, combined through a 16-entry truth table), or
* an already-materialised mask tensor (bool / int32 / float32; non-zero = selected).
"""
import math
import struct as _struct

import torch

from .. import _lib as L


class Term:
    __slots__ = ("tensor", "op", "imm")

    def __init__(self, tensor, op, imm):
        self.tensor, self.op, self.imm = tensor, op, imm


class Pred:
    """Boolean function of comparison terms: ``fn(list_of_bools) -> bool``."""

    def __init__(self, terms, fn):
        self.terms, self.fn = terms, fn

    def _combine(self, other, op):
        other = as_pred(other)
        na = len(self.terms)
        fa, fb = self.fn, other.fn
        return Pred(self.terms + other.terms, lambda v: op(fa(v[:na]), fb(v[na:])))

    def __and__(self, other):
        return self._combine(other, lambda a, b: a and b)

    def __or__(self, other):
        return self._combine(other, lambda a, b: a or b)

    def __invert__(self):
        f = self.fn
        return Pred(self.terms, lambda v: not f(v))

    def to_struct(self, n):
        if len(self.terms) > 4:
            raise L.FgxError("a native select predicate has at most 4 comparison terms")
        p = L.Predicate()
        p.n_terms = len(self.terms)
        keep = []
        for k, t in enumerate(self.terms):
            x = t.tensor
            if x.dtype == torch.bool:
                x = x.view(torch.uint8)
            x = x.reshape(-1) if x.is_contiguous() else x.contiguous().reshape(-1)
            if x.numel() % n:
                raise L.FgxError("predicate leaf size is not a multiple of the number of slots")
            keep.append(x)
            code = L.dtype_code(x)
            op, c = t.op, t.imm
            if isinstance(c, torch.Tensor):
                raise L.FgxError("a predicate immediate must be a Python scalar (a tensor would force a host sync)")
            if code == L.FGX_F32:
                imm = _struct.unpack("<I", _struct.pack("<f", float(c)))[0]
            else:
                op, c = _int_compare(op, c)
                imm = int(c) & 0xFFFFFFFF
            p.term[k].data = L.ptr(x)
            p.term[k].dtype = code
            p.term[k].op = op
            p.term[k].imm_bits = imm
            p.term[k].stride = x.numel() // n
        lut = 0
        nt = len(self.terms)
        for idx in range(1 << nt):
            if self.fn([bool((idx >> b) & 1) for b in range(nt)]):
                lut |= 1 << idx
        # replicate over the unused high term bits (they are always 0 in-kernel, but keep it total)
        p.lut = lut
        return p, keep


_I32_MIN = -(1 << 31)


def _int_compare(op, c):
    """``int_leaf <op> c`` for a possibly non-integral immediate, as an integer comparison with the same truth
    table.  JAX promotes the integer leaf and compares in floating point (``x > 0.5`` is ``x >= 1``, not
    ``x > 0``); truncating the immediate would silently select different agents."""
    if op == L.OP_NONZERO or float(c) == math.floor(float(c)):
        return op, int(c)
    lo, hi = math.floor(c), math.ceil(c)
    if op == L.OP_GT:
        return L.OP_GT, lo          # x > 2.5  <=>  x > 2
    if op == L.OP_GE:
        return L.OP_GE, hi          # x >= 2.5 <=>  x >= 3
    if op == L.OP_LT:
        return L.OP_LT, hi          # x < 2.5  <=>  x < 3
    if op == L.OP_LE:
        return L.OP_LE, lo          # x <= 2.5 <=>  x <= 2
    if op == L.OP_EQ:
        return L.OP_LT, _I32_MIN    # never
    return L.OP_GE, _I32_MIN        # OP_NE: always


class P:
    """A leaf to compare: ``P(agents.state.content['draw']) == 1``."""

    def __init__(self, tensor):
        self.tensor = tensor

    def _t(self, op, imm):
        return Pred([Term(self.tensor, op, imm)], lambda v: v[0])

    def __eq__(self, c):
        return self._t(L.OP_EQ, c)

    def __ne__(self, c):
        return self._t(L.OP_NE, c)

    def __lt__(self, c):
        return self._t(L.OP_LT, c)

    def __le__(self, c):
        return self._t(L.OP_LE, c)

    def __gt__(self, c):
        return self._t(L.OP_GT, c)

    def __ge__(self, c):
        return self._t(L.OP_GE, c)

    def nonzero(self):
        return self._t(L.OP_NONZERO, 0)

    __hash__ = None


def as_pred(x):
    if isinstance(x, Pred):
        return x
    if isinstance(x, P):
        return x.nonzero()
    if isinstance(x, torch.Tensor):
        return P(x).nonzero()
    raise TypeError(f"cannot use {type(x).__name__} as a select predicate")
