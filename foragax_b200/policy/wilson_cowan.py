"""WilsonCowan rate-RNN policy -- the B200 mirror of ``foragax/policy/wilson_cowan.py:8-66``.

Batched over agents: ``create_policy`` / ``reset_policy`` take a key batch ``uint32[N,2]`` and
draw every agent's weights with the K4 threefry kernels; ``step_policy`` runs the fused
GEMV kernel (``fgx_wilson_cowan_step``; one warp per agent, the agent's weight rows pulled
into shared memory by TMA).  The Forager agent fuses this step with its physics
(``fgx_forager_step``).
"""
import torch

from .. import _lib as L
from .. import random as fgx_random
from ..base.agent_classes import Params, Policy, Signal, State


def wc_struct(policy):
    c = policy.params.content
    s = L.WcPolicy()
    s.J, s.tau, s.E, s.B, s.D = (L.ptr(c[k]) for k in ("J", "tau", "E", "B", "D"))
    s.Z = L.ptr(policy.state.content['Z'])
    return s


def _draw(keys, num_neurons, num_obs, num_actions):
    """wilson_cowan.py:19-26: key, *init_keys = split(key, 6); J,tau,E,B,D ~ U(-1,1)."""
    n = keys.shape[0]
    ik = fgx_random.split(keys, 6)
    k = lambda i: ik[:, 1 + i, :].contiguous()  # noqa: E731
    J = fgx_random.uniform(k(0), (num_neurons, num_neurons), -1, 1)
    tau = fgx_random.uniform(k(1), (num_neurons,), -1, 1)
    E = fgx_random.uniform(k(2), (num_neurons, num_obs), -1, 1)
    B = fgx_random.uniform(k(3), (num_neurons,), -1, 1)
    D = fgx_random.uniform(k(4), (num_actions, num_neurons), -1, 1)
    return J.reshape(n, num_neurons, num_neurons), tau, E.reshape(n, num_neurons, num_obs), B, \
        D.reshape(n, num_actions, num_neurons)


class WilsonCowan(Policy):
    # wilson_cowan.py:10-31
    @staticmethod
    def create_policy(params: Params, key):
        c = params.content
        n = key.shape[0]
        J, tau, E, B, D = _draw(key, c['num_neurons'], c['num_obs'], c['num_actions'])
        Z = torch.zeros((n, c['num_neurons']), dtype=torch.float32, device=key.device)
        pp = Params(content={'J': J, 'tau': tau, 'E': E, 'B': B, 'D': D, 'dt': c['dt'],
                             'action_scale': c['action_scale']})
        return Policy(state=State(content={'Z': Z}), params=pp, deterministic=c['deterministic'])

    # wilson_cowan.py:34-51 (Z updated in place)
    @staticmethod
    def step_policy(input: Signal, policy: Policy, active_state=None):
        c = policy.params.content
        obs = input.content['obs']
        n, nn = policy.state.content['Z'].shape
        nobs, nact = c['E'].shape[2], c['D'].shape[1]
        obs = obs.reshape(n, nobs).contiguous()
        if active_state is None:
            active_state = torch.ones(n, dtype=torch.float32, device=obs.device)
        actions = torch.zeros((n, nact), dtype=torch.float32, device=obs.device)
        s = wc_struct(policy)
        L.check(L.lib.fgx_wilson_cowan_step(n, nn, nobs, nact, float(c['dt']), float(c['action_scale']),
                                            L.ptr(active_state), L.ptr(obs), s, L.ptr(actions), L.stream()))
        return Signal(content={'actions': actions}), policy

    # wilson_cowan.py:54-66
    @staticmethod
    def reset_policy(key, policy: Policy):
        c = policy.params.content
        nn, nobs, nact = c['J'].shape[1], c['E'].shape[2], c['D'].shape[1]
        J, tau, E, B, D = _draw(key, nn, nobs, nact)
        pp = Params(content={'J': J, 'tau': tau, 'E': E, 'B': B, 'D': D, 'dt': c['dt'],
                             'action_scale': c['action_scale']})
        return policy.replace(params=pp)
