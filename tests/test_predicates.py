"""Host logic of the in-kernel select predicates (no GPU needed): an integer leaf compared with a non-integral
immediate must select what JAX's promoted floating-point comparison selects (``P(int_leaf) > 0.5`` used
to truncate to ``> 0``)."""
import numpy as np
import pytest


def test_integer_leaf_with_fractional_immediate():
    from foragax_b200 import _lib as L
    from foragax_b200.base.predicates import _int_compare
    x = np.arange(-6, 7)
    ops = {L.OP_EQ: np.equal, L.OP_NE: np.not_equal, L.OP_LT: np.less, L.OP_LE: np.less_equal, L.OP_GT: np.greater,
           L.OP_GE: np.greater_equal}
    for c in (0.5, -0.5, 2.5, -2.5, 3.0, -3, 0.999, 1e-9, -1e-9, 5.999):
        for op, f in ops.items():
            op2, c2 = _int_compare(op, c)
            assert isinstance(c2, int)
            np.testing.assert_array_equal(f(x.astype(np.float64), c), ops[op2](x, c2), err_msg=f"{op} {c}")


def test_tensor_immediates_are_rejected():
    import torch
    from foragax_b200 import _lib as L
    from foragax_b200.base.predicates import P
    pred = P(torch.zeros(4, dtype=torch.int32)) > torch.tensor(1)
    with pytest.raises(L.FgxError):
        pred.to_struct(4)
