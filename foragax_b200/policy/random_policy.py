"""Random policy -- the B200 mirror of ``foragax/policy/random_policy.py:9-28`` (batched over a
key batch ``uint32[N,2]``; the draws come from the K4 threefry kernels)."""
from .. import random as fgx_random
from ..base.agent_classes import Params, Policy, Signal, State


class Random_Policy(Policy):
    # random_policy.py:9-12
    @staticmethod
    def create_policy(params: Params, key):
        return Policy(state=State(content={'key': key}), params=params, deterministic=False)

    # random_policy.py:17-23: key, subkey = split(key); action = uniform(key, (1,)); keep subkey
    @staticmethod
    def step_policy(input: Signal, policy: Policy):
        ks = fgx_random.split(policy.state.content['key'])
        action = fgx_random.uniform(ks[..., 0, :].contiguous(), (1,))
        new_policy = policy.replace(state=State(content={'key': ks[..., 1, :].contiguous()}))
        return Signal(content={'action': action}), new_policy

    # random_policy.py:25-28
    @staticmethod
    def reset_policy(key, policy: Policy):
        return policy.replace(state=State(content={'key': key}))
