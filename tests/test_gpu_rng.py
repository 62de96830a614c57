"""K4 parity: CUDA threefry / jax.random ops vs the oracle, bit-exact (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import jax_random as jr

pytestmark = pytest.mark.gpu


def dev_keys(keys, cuda):
    return torch.from_numpy(np.ascontiguousarray(keys, dtype=np.uint32)).to(cuda)


def test_kats_on_device(cuda):
    from foragax_b200 import random as fr
    k0 = fr.PRNGKey(0)
    assert fr.split(k0).cpu().numpy().tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert fr.bits(fr.PRNGKey(1), (4,)).cpu().numpy().tolist() == [0x918CA911, 0x528C7AEA, 0xE3AB1C6B, 0x78C0C24D]
    assert fr.bits(fr.PRNGKey(1), (3,)).cpu().numpy().tolist() == [0x918CA911, 0x4233DC34, 0xE3AB1C6B]
    assert fr.randint(fr.PRNGKey(1), (4,), 1, 7).cpu().numpy().tolist() == [3, 2, 6, 6]
    assert fr.uniform(k0, (1,)).cpu().numpy()[0] == np.float32(0.41845703)


@pytest.mark.parametrize("num", [2, 3, 6, 8, 11])
def test_split_batched(cuda, num):
    from foragax_b200 import random as fr
    keys = jr.split(jr.PRNGKey(123), 1000)
    got = fr.split(dev_keys(keys, cuda), num).cpu().numpy()
    np.testing.assert_array_equal(got, jr.split(keys, num))


@pytest.mark.parametrize("n", [1, 2, 3, 7, 40, 1600])
def test_bits_uniform_randint_normal(cuda, n):
    from foragax_b200 import random as fr
    keys = jr.split(jr.PRNGKey(5), 257)
    dk = dev_keys(keys, cuda)
    np.testing.assert_array_equal(fr.bits(dk, (n,)).cpu().numpy(), jr.random_bits(keys, n))
    np.testing.assert_array_equal(fr.uniform(dk, (n,), -1.0, 1.0).cpu().numpy(), jr.uniform(keys, n, -1.0, 1.0))
    np.testing.assert_array_equal(fr.uniform(dk, (n,), 5.0, 10.0).cpu().numpy(), jr.uniform(keys, n, 5.0, 10.0))
    np.testing.assert_array_equal(fr.randint(dk, (n,), 1, 7).cpu().numpy(), jr.randint(keys, n, 1, 7))
    np.testing.assert_array_equal(fr.randint(dk, (n,), -5, 1000003).cpu().numpy(), jr.randint(keys, n, -5, 1000003))
    np.testing.assert_array_equal(fr.randint(dk, (n,), 3, 3).cpu().numpy(), jr.randint(keys, n, 3, 3))
    # normal: XLA's erf_inv polynomial restated on both sides; tolerance 1e-5 relative (fp32 contraction differs)
    np.testing.assert_allclose(fr.normal(dk, (n,)).cpu().numpy(), jr.normal(keys, n), rtol=1e-5, atol=1e-6)


def test_one_key_many_draws_1e6(cuda):
    from foragax_b200 import random as fr
    n = 1_000_001
    got = fr.bits(fr.PRNGKey(42), (n,)).cpu().numpy()
    np.testing.assert_array_equal(got, jr.random_bits(jr.PRNGKey(42), n))
