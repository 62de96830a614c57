"""Oracle pinning: threefry known-answer vectors (Random123) and the jax.random values
quoted in SURVEY.md Appendix A.3/A.4 (legacy counter layout)."""
import numpy as np

from oracle import jax_random as jr


def test_threefry_kats():
    assert [int(v) for v in jr.threefry2x32(0, 0, 0, 0)] == [0x6B200159, 0x99BA4EFE]
    o = jr.threefry2x32(0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF)
    assert [int(v) for v in o] == [0x1CB996FC, 0xBB002BE7]
    o = jr.threefry2x32(0x13198A2E, 0x03707344, 0x243F6A88, 0x85A308D3)
    assert [int(v) for v in o] == [0xC4923A9C, 0x483DF7A0]


def test_split_layout():
    assert jr.split(jr.PRNGKey(0)).tolist() == [[4146024105, 967050713], [2718843009, 1272950319]]
    assert jr.split(jr.PRNGKey(1)).tolist() == [[2441914641, 1384938218], [3819641963, 2025898573]]
    assert jr.split(jr.PRNGKey(0), 11)[:3].tolist() == [[255289810, 1154544838], [2600436532, 3355744386],
                                                        [2335412307, 2417518091]]


def test_bits_uniform_randint():
    assert int(jr.random_bits(jr.PRNGKey(0), 1)[0]) == 0x6B200159
    assert jr.random_bits(jr.PRNGKey(1), 4).tolist() == [0x918CA911, 0x528C7AEA, 0xE3AB1C6B, 0x78C0C24D]
    assert jr.random_bits(jr.PRNGKey(1), 3).tolist() == [0x918CA911, 0x4233DC34, 0xE3AB1C6B]
    assert jr.uniform(jr.PRNGKey(0), 1)[0] == np.float32(0.41845703)
    assert jr.randint(jr.PRNGKey(1), 4, 1, 7).tolist() == [3, 2, 6, 6]
    np.testing.assert_array_equal(jr.uniform(jr.PRNGKey(1), 3, -1, 1),
                                  np.array([0.137104988, -0.482792377, 0.778659344], np.float32))


def test_normal_matches_published_jax_output():
    # jax.random.normal(jax.random.PRNGKey(0), (5,)) on JAX < 0.5 (legacy threefry)
    ref = np.array([0.18784384, -1.2833426, -0.2710917, 1.2490593, 0.24447003], np.float32)
    np.testing.assert_allclose(jr.normal(jr.PRNGKey(0), 5), ref, rtol=2e-6)


def test_batched_keys_equal_loop():
    keys = jr.split(jr.PRNGKey(7), 9)
    b = jr.randint(keys, 1, 1, 7)
    for i in range(9):
        assert b[i, 0] == jr.randint(keys[i], 1, 1, 7)[0]
    s = jr.split(keys, 3)
    for i in range(9):
        np.testing.assert_array_equal(s[i], jr.split(keys[i], 3))
