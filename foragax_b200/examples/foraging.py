"""The EcoEvoJax foraging example (reference
``examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.py:24-553,597-666``) on native ops:
``Forager`` (WilsonCowan policy + point-mass physics + energy), ``Resource`` (logistic growth +
blossom), ``forager_resource_interaction``, ``sensor_model``, ``Neuroevolution`` and the
``Foraging_world`` driver -- same names, arguments and return values.  BASELINE config 2.

``HYPER`` holds the module constants of ``sim.py:16-19`` and the constants scattered through its
functions; ``use_variant('nb')`` switches to the values of the notebook twin ``sim.ipynb`` (the run
that produced golden vector G2).  The recording ring and its checkpoints (``sim.py:555-595,643-647,695-700``)
are in ``foragax_b200/recording.py`` (a native ring-write kernel; orbax replaced by a small on-disk format).
"""
import os

import numpy as np
import torch

from .. import _lib as L
from .. import random as fgx_random
from .. import struct
from ..base import space_methods as sm
from ..base.agent_classes import Agent, Agent_Set, Params, Policy, Signal, State, native
from ..base.agent_methods import (count_active, count_active_clip, create_agents, jit_add_agents, jit_remove_agents,
                                  jit_select_agents, jit_set_agents, jit_sort_agents, step_agents)
from ..base.predicates import P
from ..base.space_methods import create_space

_CHECK_ADD = os.environ.get("FGX_CHECK_ADD", "0") == "1"  # verify add_agents' "parents below the new rows" precondition
from ..policy.wilson_cowan import wc_struct

f32 = np.float32


class _Hyper:
    def __init__(self):
        self.use('sim')

    def use(self, variant):
        self.variant = variant
        if variant == 'sim':       # sim.py:16-19,109-111,483,517,659
            self.MAX_AGE, self.TIME_TO_DIE, self.NOISE, self.NUMBER_OF_NEURONS = 1000.0, 3.0, 0.02, 40
            self.adaptive_noise, self.energy_cost, self.clip_min = False, 0.3, True
            self.repr_thresh, self.time_to_reproduce = 5.0, 0.4
            self.resource_margin = True      # resource spawn range +-10 (sim.py:248-249)
        elif variant == 'nb':      # sim.ipynb
            self.MAX_AGE, self.TIME_TO_DIE, self.NOISE, self.NUMBER_OF_NEURONS = 50.0, 1.0, 0.02, 30
            self.adaptive_noise, self.energy_cost, self.clip_min = True, 1.0, False
            self.repr_thresh, self.time_to_reproduce = 5.5, 0.5
            self.resource_margin = False
        else:
            raise ValueError(variant)
        self.die_thresh, self.energy_min, self.energy_max = 0.0, -1.0, 15.0
        self.blossom_thresh, self.resource_energy_min = 3.0, 0.1
        self.ray_span, self.ray_max_length = float(np.pi / 4.0), 60.0


HYPER = _Hyper()
use_variant = HYPER.use

FSTATE = L.FORAGER_STATE_FIELDS
RSTATE = ("X", "Y", "energy", "rad", "time_above_blossom", "blossom", "key")


def _fstate(agents):
    s = L.ForagerState()
    for k in FSTATE:
        setattr(s, k, L.ptr(agents.state.content[k]))
    return s


def _rstate(agents):
    s = L.ResourceState()
    c = agents.state.content
    for k in RSTATE:
        setattr(s, k, L.ptr(c[k].view(torch.uint8) if c[k].dtype == torch.bool else c[k]))
    return s


def _forager_params(space, dt, policy_params, step=None):
    p = L.ForagerParams()
    p.n_neurons, p.n_obs, p.n_act = policy_params['num_neurons'], policy_params['num_obs'], policy_params['num_actions']
    p.dt = p.policy_dt = float(dt)
    p.action_scale = float(policy_params['action_scale'])
    p.x_min, p.x_max, p.y_min, p.y_max = (float(v) for v in (space.x_min, space.x_max, space.y_min, space.y_max))
    step = step or {}
    p.repr_thresh = float(step.get('repr_thresh', HYPER.repr_thresh))
    p.die_thresh = float(step.get('die_thresh', HYPER.die_thresh))
    p.energy_min = float(step.get('energy_min', HYPER.energy_min))
    p.energy_max = float(step.get('energy_max', HYPER.energy_max))
    p.energy_cost_dt = float(f32(dt) if HYPER.energy_cost == 1.0 else f32(HYPER.energy_cost) * f32(dt))
    p.clip_min = 1 if HYPER.clip_min else 0
    return p


def _policy_shape(agents):
    c = agents.policy.params.content
    return {'num_neurons': c['J'].shape[1], 'num_obs': c['E'].shape[2], 'num_actions': c['D'].shape[1],
            'action_scale': c['action_scale'], 'dt': c['dt']}


@struct.dataclass
class Forager(Agent):
    # sim.py:27-71
    @staticmethod
    @native
    def create_agent(params, unique_id, active_state, agent_type, key):
        n, dev = unique_id.shape[0], unique_id.device
        pp, space, dt = params.content['policy_params'], params.content['space'], params.content['dt']
        # create bounds of X / Y (sim.py:39,43), evaluated in float64 like the reference's Python floats
        cp = _forager_params(space, dt, pp)
        cp.x_min = float(space.x_min + (space.x_min + 1.0) / 10.0)
        cp.x_max = float(space.x_max - space.x_max / 10.0)
        cp.y_min = float(space.y_min + (space.y_min + 1.0) / 10.0)
        cp.y_max = float(space.y_max - space.y_max / 10.0)
        nn, nobs, nact = pp['num_neurons'], pp['num_obs'], pp['num_actions']
        st = {k: torch.empty((n, 1), dtype=torch.float32, device=dev) for k in FSTATE}
        Z = torch.empty((n, nn), dtype=torch.float32, device=dev)
        ik = torch.empty((5, n, 2), dtype=torch.uint32, device=dev)
        age = torch.empty(n, dtype=torch.float32, device=dev)
        s = L.ForagerState()
        for k in FSTATE:
            setattr(s, k, L.ptr(st[k]))
        L.check(L.lib.fgx_forager_create(cp, n, L.ptr(key), L.ptr(active_state), s, L.ptr(Z), L.ptr(ik), L.ptr(age),
                                         L.stream()))
        # WilsonCowan.create_policy (wilson_cowan.py:19-26) from the five init keys
        J = fgx_random.uniform(ik[0], (nn * nn,), -1, 1).reshape(n, nn, nn)
        tau = fgx_random.uniform(ik[1], (nn,), -1, 1)
        E = fgx_random.uniform(ik[2], (nn * nobs,), -1, 1).reshape(n, nn, nobs)
        B = fgx_random.uniform(ik[3], (nn,), -1, 1)
        D = fgx_random.uniform(ik[4], (nact * nn,), -1, 1).reshape(n, nact, nn)
        policy = Policy(state=State(content={'Z': Z}),
                        params=Params(content={'J': J, 'tau': tau, 'E': E, 'B': B, 'D': D, 'dt': pp['dt'],
                                               'action_scale': pp['action_scale']}),
                        deterministic=pp['deterministic'])
        return Forager(state=State(content=st), policy=policy, params=Params(content={'dt': dt}),
                       unique_id=unique_id, agent_type=agent_type, age=age, active_state=active_state)

    # sim.py:74-136 (in place)
    @staticmethod
    @native
    def step_agent(params, input, forager):
        n = forager.unique_id.shape[0]
        pshape = _policy_shape(forager)
        p = _forager_params(params.content['space'], forager.params.content['dt'], pshape, params.content)
        obs = input.content['obs'].reshape(n, -1)
        if obs.shape[1] != pshape['num_obs'] - 1:
            raise ValueError(f"obs has {obs.shape[1]} columns, the policy expects {pshape['num_obs'] - 1} (+ energy)")
        energy_in = input.content['energy_in'].reshape(n)
        L.check(L.lib.fgx_forager_step(p, n, L.ptr(forager.active_state), L.ptr(obs.contiguous()),
                                       L.ptr(energy_in.contiguous()), _fstate(forager), wc_struct(forager.policy),
                                       L.ptr(forager.age), L.stream()))
        return forager

    @staticmethod
    def sense_step(params, energy_in, forager, walls, bounds, ray_span, ray_length, out=None):
        """``sensor -> step_agents`` of a tick (sim.py:652-660) in ONE kernel, for a Forager that observes its wall
        distances + own energy (``num_obs == n_rays + 1``): each agent's rays are cast against the wall grid while
        its weights are in flight and fed to the policy from shared memory.  Returns ``(forager, dist, hit)`` --
        the same bits as ``raycast_walls(..., bounds=bounds)`` followed by ``step_agent`` with ``obs = dist``."""
        from ..base import space_methods as sm
        n = forager.unique_id.shape[0]
        pshape = _policy_shape(forager)
        p = _forager_params(params.content['space'], forager.params.content['dt'], pshape, params.content)
        r = pshape['num_obs'] - 1
        if out is None:
            dist = torch.empty((n, r), dtype=torch.float32, device=forager.age.device)
            hit = torch.empty((n, r), dtype=torch.int32, device=forager.age.device)
        else:
            dist, hit = out
        w1, w2, w3, w4, nw, keep = sm._wall_ptrs(walls)
        tk = sm.wall_grid(walls, bounds, ray_length)
        b, scr = tk.bounds, tk.scratch
        L.check(L.lib.fgx_forager_sense_step(p, n, L.ptr(forager.active_state), L.ptr(energy_in.reshape(n).contiguous()),
                                             _fstate(forager), wc_struct(forager.policy), L.ptr(forager.age), r,
                                             float(ray_span), float(ray_length), w1, w2, w3, w4, nw, b[0], b[1], b[2],
                                             b[3], L.ptr(scr), scr.numel(), tk.reuse, L.ptr(dist), L.ptr(hit),
                                             L.stream()))
        tk.done()
        return forager, dist, hit

    # sim.py:139-151
    @staticmethod
    @native
    def remove_agent(params, foragers, rows):
        n = foragers.unique_id.shape[0]
        Z = foragers.policy.state.content['Z']
        L.check(L.lib.fgx_forager_remove(n, Z.shape[1], L.ptr(rows.ids), L.ptr(rows.count), _fstate(foragers),
                                         L.ptr(Z), L.ptr(foragers.active_state), L.ptr(foragers.age), L.stream()))
        return foragers

    # sim.py:154-210; `key` is a uint32[2] device key, threaded through the rows serially
    @staticmethod
    @native
    def add_agent(params, foragers, rows, key):
        n = foragers.unique_id.shape[0]
        ps = _policy_shape(foragers)
        good_ids = params.content['good_ids']
        if _CHECK_ADD:
            from ..base.agent_methods import check_add_parents
            check_add_parents(good_ids, rows)
        key_out = torch.empty_like(key)
        chain = torch.empty((n, 2), dtype=torch.uint32, device=key.device)
        L.check(L.lib.fgx_forager_add(n, ps['num_neurons'], ps['num_obs'], ps['num_actions'], L.ptr(good_ids),
                                      L.ptr(rows.start), L.ptr(rows.count), L.ptr(key), L.ptr(key_out),
                                      float(HYPER.NOISE), 1 if HYPER.adaptive_noise else 0, _fstate(foragers),
                                      wc_struct(foragers.policy), L.ptr(foragers.active_state), L.ptr(foragers.age),
                                      L.ptr(chain), L.stream()))
        return foragers, key_out

    # sim.py:213-225
    @staticmethod
    @native
    def set_agent(params, foragers, rows, key):
        n = foragers.unique_id.shape[0]
        L.check(L.lib.fgx_forager_set(n, L.ptr(rows.ids), L.ptr(rows.count), _fstate(foragers), L.stream()))
        return foragers, key


@struct.dataclass
class Resource(Agent):
    # sim.py:231-261
    @staticmethod
    @native
    def create_agent(params, unique_id, active_state, agent_type, key):
        n, dev = unique_id.shape[0], unique_id.device
        space, dt = params.content['space'], params.content['dt']
        if HYPER.resource_margin:
            lo_x, hi_x = space.x_min + 10.0, space.x_max - 10.0
            lo_y, hi_y = space.y_min + 10.0, space.y_max - 10.0
        else:
            lo_x, hi_x = space.x_min + (space.x_min + 1.0) / 10.0, space.x_max - space.x_max / 10.0
            lo_y, hi_y = space.y_min + (space.y_min + 1.0) / 10.0, space.y_max - space.y_max / 10.0
        f = lambda: torch.empty((n, 1), dtype=torch.float32, device=dev)  # noqa: E731
        st = {'X': f(), 'Y': f(), 'energy': f(), 'rad': f(), 'blossom': torch.empty(n, dtype=torch.bool, device=dev),
              'key': torch.empty((n, 2), dtype=torch.uint32, device=dev), 'time_above_blossom': f()}
        gr, dr = f(), f()
        age = torch.empty(n, dtype=torch.float32, device=dev)
        s = L.ResourceState()
        for k in RSTATE:
            setattr(s, k, L.ptr(st[k].view(torch.uint8) if st[k].dtype == torch.bool else st[k]))
        L.check(L.lib.fgx_resource_create(n, float(lo_x), float(hi_x), float(lo_y), float(hi_y), L.ptr(key),
                                          L.ptr(active_state), s, L.ptr(gr), L.ptr(dr), L.ptr(age), L.stream()))
        agent_params = Params(content={'growth_rate': gr, 'decay_rate': dr, 'dt': dt, 'x_min': space.x_min,
                                       'x_max': space.x_max, 'y_min': space.y_min, 'y_max': space.y_max})
        return Resource(state=State(content=st), policy=None, params=agent_params, unique_id=unique_id,
                        agent_type=agent_type, age=age, active_state=active_state)

    # sim.py:264-303 (in place)
    @staticmethod
    @native
    def step_agent(params, input, res):
        n = res.unique_id.shape[0]
        c = res.params.content
        p = L.ResourceParams(float(c['dt']), float(c['x_min']), float(c['x_max']),
                             float(params.content['blossom_thresh']))
        eo = input.content['energy_out'].reshape(n).contiguous()
        L.check(L.lib.fgx_resource_step(p, n, L.ptr(res.active_state), L.ptr(eo), _rstate(res), L.ptr(c['growth_rate']),
                                        L.ptr(c['decay_rate']), L.ptr(res.age), L.stream()))
        return res

    # sim.py:306-313
    @staticmethod
    @native
    def remove_agent(remove_params, agents, rows):
        n = agents.unique_id.shape[0]
        L.check(L.lib.fgx_resource_remove(n, L.ptr(rows.ids), L.ptr(rows.count), _rstate(agents),
                                          L.ptr(agents.active_state), L.ptr(agents.age), L.stream()))
        return agents

    # sim.py:316-344
    @staticmethod
    @native
    def add_agent(params, ress, rows, key):
        n = ress.unique_id.shape[0]
        c = ress.params.content
        if _CHECK_ADD:
            from ..base.agent_methods import check_add_parents
            check_add_parents(params.content['good_ids'], rows)
        key_out = torch.empty_like(key)
        chain = torch.empty((n, 2), dtype=torch.uint32, device=key.device)
        L.check(L.lib.fgx_resource_add(n, float(c['x_min']), float(c['x_max']), float(c['y_min']), float(c['y_max']),
                                       L.ptr(params.content['good_ids']), L.ptr(rows.start), L.ptr(rows.count),
                                       L.ptr(key), L.ptr(key_out), _rstate(ress), L.ptr(c['growth_rate']),
                                       L.ptr(c['decay_rate']), L.ptr(ress.active_state), L.ptr(ress.age), L.ptr(chain),
                                       L.stream()))
        return ress, key_out

    # sim.py:347-355
    @staticmethod
    @native
    def set_agent(params, ress, rows, key):
        n = ress.unique_id.shape[0]
        L.check(L.lib.fgx_resource_set(n, L.ptr(rows.ids), L.ptr(rows.count), _rstate(ress), L.stream()))
        return ress, key


# sim.py:358-384.  The reference also returns the [F, R+F, 1] distance matrix that sensor_model
# prefilters with; the native sensor recomputes those distances on the fly, so the first return
# value is None.
def forager_resource_interaction(foragers, resources):
    nf, nr = foragers.unique_id.shape[0], resources.unique_id.shape[0]
    fs, rs = foragers.state.content, resources.state.content
    energy_in = torch.empty(nf, dtype=torch.float32, device=fs['X'].device)
    energy_out = torch.empty(nr, dtype=torch.float32, device=fs['X'].device)
    L.check(L.lib.fgx_forager_resource_interaction(nf, L.ptr(fs['X']), L.ptr(fs['Y']), nr, L.ptr(rs['X']),
                                                   L.ptr(rs['Y']), L.ptr(rs['rad']), L.ptr(rs['energy']),
                                                   L.ptr(energy_in), L.ptr(energy_out), L.stream()))
    return None, energy_in, energy_out


jit_forager_resource_interaction = forager_resource_interaction

sm.RESOLUTION = 9  # sim.py:387


# sim.py:388-468
def sensor_model(foragers, resources, dist_mat, walls):
    nf, nr = foragers.unique_id.shape[0], resources.unique_id.shape[0]
    fs, rs = foragers.state.content, resources.state.content
    r = sm.RESOLUTION
    obs = torch.empty((nf, r * 3), dtype=torch.float32, device=fs['X'].device)
    w1, w2, w3, w4, nw, keep = sm._wall_ptrs(walls)
    L.check(L.lib.fgx_sensor_model(nf, L.ptr(fs['X']), L.ptr(fs['Y']), L.ptr(fs['ang']), L.ptr(fs['rad']),
                                   L.ptr(fs['energy']), L.ptr(foragers.active_state), L.ptr(foragers.agent_type), nr,
                                   L.ptr(rs['X']), L.ptr(rs['Y']), L.ptr(rs['rad']), L.ptr(rs['energy']),
                                   L.ptr(resources.active_state), L.ptr(resources.agent_type), nw, w1, w2, w3, w4, r,
                                   HYPER.ray_span, HYPER.ray_max_length, L.ptr(obs), L.stream()))
    return obs


jit_sensor_model = sensor_model


# sim.py:474-553
def _sum_f32(x):
    """``jnp.sum(x)`` of a float32 leaf as a 1-element device tensor (``fgx_sum_f32``: one launch, fixed order)."""
    x = x.reshape(-1)
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    nbytes = L.lib.fgx_sum_f32_scratch_bytes()
    scr = L.scratch(nbytes)
    L.check(L.lib.fgx_sum_f32(L.ptr(x.contiguous()), x.shape[0], L.ptr(out), L.ptr(scr), nbytes, L.stream()))
    return out


def Neuroevolution(foragers, resources, key, max_foragers, max_resources, space, life_expectancy_old):
    def select_weak_or_old_foragers(foragers, select_params):
        cond = (P(foragers.active_state) == 1.0) & (P(foragers.state.content['time_below_die_thresh'])
                                                    > select_params.content['time_to_die'])
        return (P(foragers.age) > select_params.content['max_age']) | cond

    select_params = Params(content={'time_to_die': HYPER.TIME_TO_DIE, 'max_age': HYPER.MAX_AGE})
    num_foragers_remove, remove_indx = jit_select_agents(select_weak_or_old_foragers, select_params, foragers)
    age_before_death = _sum_f32(foragers.age)
    foragers = jit_remove_agents(Forager.remove_agent, num_foragers_remove, Params(content={'remove_ids': remove_indx}),
                                 foragers)
    age_after_death = _sum_f32(foragers.age)
    # sim.py:486-493 on device scalars, one launch (no host read-back, no framework elementwise kernels)
    if not isinstance(life_expectancy_old, torch.Tensor):
        life_expectancy_old = torch.as_tensor(life_expectancy_old, dtype=torch.float32, device=foragers.age.device)
    life_expectancy = torch.empty((), dtype=torch.float32, device=foragers.age.device)
    L.check(L.lib.fgx_life_expectancy(L.ptr(age_before_death), L.ptr(age_after_death),
                                      L.ptr(L.device_i32(num_foragers_remove)),
                                      L.ptr(life_expectancy_old.reshape(1).to(torch.float32)), L.ptr(life_expectancy),
                                      L.stream()))

    def select_depleted_resources(resources, select_params):
        return (P(resources.active_state) == 1.0) & (P(resources.state.content['energy'])
                                                     < select_params.content['energy_min'])

    num_resources_remove, remove_indx = jit_select_agents(select_depleted_resources,
                                                          Params(content={'energy_min': HYPER.resource_energy_min}),
                                                          resources)
    resources = jit_remove_agents(Resource.remove_agent, num_resources_remove,
                                  Params(content={'remove_ids': remove_indx}), resources)

    foragers, sort_ids = jit_sort_agents(foragers.state.content['time_above_rep_thresh'], False, foragers)
    resources, sort_ids = jit_sort_agents(resources.state.content['time_above_blossom'], False, resources)

    def select_good_foragers(foragers, select_params):
        return (P(foragers.active_state) == 1.0) & (P(foragers.state.content['time_above_rep_thresh'])
                                                    > select_params.content['time_to_reproduce'])

    num_foragers_add, add_indx = jit_select_agents(select_good_foragers,
                                                   Params(content={'time_to_reproduce': HYPER.time_to_reproduce}),
                                                   foragers)
    if int(max_foragers) != int(foragers.active_state.shape[0]):
        raise L.FgxError("Neuroevolution: max_foragers must be the slot count of the forager set (sim.py:664)")
    num_active_foragers, num_foragers_add = count_active_clip(foragers, num_foragers_add)  # sim.py:520-521
    add_params = Params(content={'good_ids': add_indx, 'num_active_foragers': num_active_foragers})
    foragers, key = jit_add_agents(Forager.add_agent, num_foragers_add, add_params, foragers, key,
                                   num_active=num_active_foragers)
    foragers, key = jit_set_agents(Forager.set_agent, num_foragers_add, Params(content={'set_ids': add_indx}), foragers,
                                   key)

    def select_blossom_resources(resources, select_params):
        return (P(resources.active_state) == 1.0) & P(resources.state.content['blossom']).nonzero()

    num_resources_add, add_indx = jit_select_agents(select_blossom_resources, None, resources)
    if int(max_resources) != int(resources.active_state.shape[0]):
        raise L.FgxError("Neuroevolution: max_resources must be the slot count of the resource set (sim.py:664)")
    num_active_resources, num_resources_add = count_active_clip(resources, num_resources_add)  # sim.py:539-540
    add_params = Params(content={'good_ids': add_indx, 'num_active_resources': num_active_resources})
    resources, key = jit_add_agents(Resource.add_agent, num_resources_add, add_params, resources, key,
                                    num_active=num_active_resources)
    resources, key = jit_set_agents(Resource.set_agent, num_resources_add, Params(content={'set_ids': add_indx}),
                                    resources, key)
    return foragers, resources, key, num_active_foragers, num_active_resources, life_expectancy


jit_Neuroevolution = Neuroevolution


# sim.py:597-701
class Foraging_world:
    def __init__(self, params: Params, key, steps, num_rec_frames=0, rec_dir='./rec_data'):
        self.sim_steps = steps
        space_params = params.content['space_params']
        forager_params = params.content['forager_params']
        resource_params = params.content['resource_params']
        self.space = create_space(x_min=space_params['x_min'], x_max=space_params['x_max'],
                                  y_min=space_params['y_min'], y_max=space_params['y_max'],
                                  torous=space_params['torous'], wall_array=space_params['walls'])
        foragers_create_params = Params(content={
            'policy_params': {k: forager_params[k] for k in ('num_neurons', 'num_obs', 'num_actions', 'dt',
                                                             'action_scale', 'deterministic')},
            'dt': forager_params['dt'], 'space': self.space})
        resources_create_params = Params(content={'space': self.space, 'dt': resource_params['dt']})
        self.foragers_set = Agent_Set(agent=Forager, num_total_agents=forager_params['total_num'],
                                      num_active_agents=forager_params['active_num'], agent_type=forager_params['type'])
        ks = fgx_random.split(key)
        self.key, subkey = ks[0].contiguous(), ks[1].contiguous()
        self.foragers_set.agents = create_agents(foragers_create_params, self.foragers_set, subkey)
        self.resources_set = Agent_Set(agent=Resource, num_total_agents=resource_params['total_num'],
                                       num_active_agents=resource_params['active_num'],
                                       agent_type=resource_params['type'])
        ks = fgx_random.split(self.key)
        self.key, subkey = ks[0].contiguous(), ks[1].contiguous()
        self.resources_set.agents = create_agents(resources_create_params, self.resources_set, subkey)
        self.life_expectancy = torch.tensor(50.0, dtype=torch.float32, device=self.key.device)
        # sim.py:640-647: the recording ring and its checkpoint manager (orbax in the reference; see recording.py)
        self.num_rec_frames = int(num_rec_frames)
        self.rec_data, self.checkpoint_manager = None, None
        if self.num_rec_frames > 0:
            from ..recording import CheckpointManager, init_rec_data
            self.rec_data = init_rec_data(self.num_rec_frames, self.foragers_set, self.resources_set)
            self.checkpoint_manager = CheckpointManager(rec_dir, max_to_keep=100, create=True)
        self.sim_step = 0

    def step(self):
        dist_mat, energy_in, energy_out = jit_forager_resource_interaction(self.foragers_set.agents,
                                                                           self.resources_set.agents)
        sensor_data = jit_sensor_model(self.foragers_set.agents, self.resources_set.agents, dist_mat, self.space.walls)
        forager_input = Signal(content={'obs': sensor_data, 'energy_in': energy_in})
        resource_input = Signal(content={'energy_out': energy_out})
        forager_step_params = Params(content={'space': self.space, 'repr_thresh': HYPER.repr_thresh,
                                              'die_thresh': HYPER.die_thresh, 'energy_min': HYPER.energy_min,
                                              'energy_max': HYPER.energy_max})
        resource_step_params = Params(content={'blossom_thresh': HYPER.blossom_thresh})
        self.foragers_set.agents = step_agents(forager_step_params, forager_input, self.foragers_set)
        self.resources_set.agents = step_agents(resource_step_params, resource_input, self.resources_set)
        (self.foragers_set.agents, self.resources_set.agents, self.key, num_active_foragers, num_active_resources,
         self.life_expectancy) = jit_Neuroevolution(self.foragers_set.agents, self.resources_set.agents, self.key,
                                                    self.foragers_set.num_total_agents,
                                                    self.resources_set.num_total_agents, self.space,
                                                    self.life_expectancy)
        return num_active_foragers, num_active_resources, self.life_expectancy

    def capture(self, warmup=2):
        """Capture one whole tick into a CUDA graph (``foragax_b200.graphs.GraphedTick``): ``warmup`` real
        ticks are run first; afterwards every ``replay()`` is one tick and returns the static
        ``(num_active_foragers, num_active_resources, life_expectancy)`` tensors."""
        from ..graphs import GraphedTick
        state = {'foragers': self.foragers_set.agents, 'resources': self.resources_set.agents, 'key': self.key,
                 'life': self.life_expectancy}

        def bind(st):
            self.foragers_set.agents, self.resources_set.agents = st['foragers'], st['resources']
            self.key, self.life_expectancy = st['key'], st['life']

        def tick(st):
            bind(st)
            out = self.step()
            return {'foragers': self.foragers_set.agents, 'resources': self.resources_set.agents, 'key': self.key,
                    'life': self.life_expectancy}, out

        g = GraphedTick(tick, state, warmup)
        bind(state)
        return g

    def run(self, rec_start_frame=None, print_freq=0):
        """sim.py:668-701: ``sim_steps`` ticks; from tick ``rec_start_frame`` on, ``num_rec_frames`` frames of
        positions / headings / radii / blossoms are recorded into the ring and the full ring is checkpointed.
        Returns the per-tick ``(num_active_foragers, num_active_resources, life_expectancy)`` device scalars."""
        from ..recording import init_rec_data, jit_record_data
        out = []
        record_flag, record_frame = False, 0
        average_life_expectancy = 0.0
        for _ in range(self.sim_steps):
            nf, nr, le = self.step()
            out.append((nf, nr, le))
            sim_step = self.sim_step
            if print_freq:
                average_life_expectancy += float(le)
                if sim_step % print_freq == 0 and sim_step != 0:
                    print(sim_step)
                    print("num_active_foragers: ", int(nf))
                    print("num_active_resources: ", int(nr))
                    print("avg life_expectancy: ", average_life_expectancy / print_freq)
                    average_life_expectancy = 0.0
            self.sim_step = sim_step = sim_step + 1
            if self.rec_data is not None and rec_start_frame is not None:
                if sim_step == rec_start_frame:
                    record_flag, record_frame = True, 0
                    self.rec_data = init_rec_data(self.num_rec_frames, self.foragers_set, self.resources_set)
                if record_flag:
                    self.rec_data = jit_record_data(self.foragers_set.agents, self.resources_set.agents, self.rec_data,
                                                    record_frame)
                    record_frame += 1
                    if record_frame == self.num_rec_frames:
                        self.checkpoint_manager.save(sim_step, {'rec_data': self.rec_data})
                        record_flag = False
        return out

    def save_state(self, manager, step=None):
        """Checkpoint of the simulation STATE (not in the reference, which cannot resume): both agent sets, the
        world key, the life-expectancy average and the tick counter."""
        from ..recording import save_state
        step = self.sim_step if step is None else step
        return save_state(manager, step, foragers=self.foragers_set.agents, resources=self.resources_set.agents,
                          key=self.key, life_expectancy=self.life_expectancy, sim_step=int(self.sim_step))

    def load_state(self, manager, step=None):
        """Resume from :meth:`save_state`: the checkpointed leaves are copied into this world's buffers."""
        from ..recording import load_state
        template = {'foragers': self.foragers_set.agents, 'resources': self.resources_set.agents, 'key': self.key,
                    'life_expectancy': self.life_expectancy}
        step, scalars = load_state(manager, template, step)
        self.sim_step = int(scalars.get('sim_step', step))
        return step


def default_params(variant='sim'):
    """The driver constants of sim.py:704-715 ('sim') or of the notebook ('nb')."""
    if variant == 'sim':
        X_max, Y_max, neurons, res = 1000.0, 500.0, 40, (6000, 5000)
    else:
        X_max, Y_max, neurons, res = 2000.0, 500.0, 30, (4000, 3000)
    wall_array = [[[0.0, 0.0], [X_max, 0.0]], [[X_max, 0.0], [X_max, Y_max]], [[0.0, 0.0], [0.0, Y_max]],
                  [[0.0, Y_max], [X_max, Y_max]]]
    dt = 0.1
    return Params(content={
        'space_params': {'x_min': 0.0, 'x_max': X_max, 'y_min': 0.0, 'y_max': Y_max, 'torous': False,
                         'walls': wall_array},
        'forager_params': {'total_num': 1000, 'active_num': 500, 'type': 1, 'dt': dt, 'num_neurons': neurons,
                           'num_obs': 28, 'num_actions': 2, 'action_scale': 1.0, 'deterministic': True},
        'resource_params': {'total_num': res[0], 'active_num': res[1], 'type': 2, 'dt': dt}})
