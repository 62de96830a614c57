"""The wolf-sheep-grass example (reference ``examples/basic/wolf_sheep_grass/sim.ipynb``: ``Animal``
:37-170, ``Grass`` :174-218, ``interaction`` :238-290, ``Ecosystem`` :365-517) on native ops -- same
names, arguments and return values.  BASELINE config 3 (birth/death churn).  Recording / orbax
checkpointing of the reference driver is visualisation I/O, not part of the hot path.
"""
import torch

from .. import _lib as L
from .. import random as fgx_random
from .. import struct
from ..base.agent_classes import Agent, Agent_Set, Params, Policy, Signal, State, native
from ..base.agent_methods import (count_active, create_agents, jit_add_agents, jit_remove_agents,
                                  jit_select_agents, jit_set_agents, jit_sort_agents, step_agents)
from ..base.predicates import P
from ..base.space_methods import create_space


def _astate(an):
    s = L.AnimalState()
    c = an.state.content
    for k in ("X_pos", "Y_pos", "energy", "reproduce", "key"):
        setattr(s, k, L.ptr(c[k]))
    s.policy_key = L.ptr(an.policy.state.content['key'])
    return s


@struct.dataclass
class Animal(Agent):
    # sim.ipynb:40-70
    @staticmethod
    @native
    def create_agent(params, unique_id, active_state, agent_type, key):
        n, dev = unique_id.shape[0], unique_id.device
        c = params.content
        space = c['space']
        i1 = lambda: torch.empty((n, 1), dtype=torch.int32, device=dev)  # noqa: E731
        st = {'X_pos': i1(), 'Y_pos': i1(), 'energy': i1(), 'reproduce': torch.empty(n, dtype=torch.int32, device=dev),
              'key': torch.empty((n, 2), dtype=torch.uint32, device=dev)}
        pkey = torch.empty((n, 2), dtype=torch.uint32, device=dev)
        age = torch.empty(n, dtype=torch.float32, device=dev)
        s = L.AnimalState()
        for k in st:
            setattr(s, k, L.ptr(st[k]))
        s.policy_key = L.ptr(pkey)
        L.check(L.lib.fgx_animal_create(n, int(space.x_max), int(space.y_max), int(c['delta_energy']), L.ptr(key),
                                        L.ptr(active_state), s, L.ptr(age), L.stream()))
        policy = Policy(state=State(content={'key': pkey}), params=params, deterministic=False)  # Random_Policy
        agent_params = Params(content={'reproduction_rate': c['reproduction_rate'], 'delta_energy': c['delta_energy']})
        return Animal(unique_id=unique_id, active_state=active_state, agent_type=agent_type, state=State(content=st),
                      policy=policy, params=agent_params, age=age)

    # sim.ipynb:72-122 (in place)
    @staticmethod
    @native
    def step_agent(params, input, animal):
        n = animal.unique_id.shape[0]
        space = params.content['space']
        p = L.AnimalParams(int(space.x_min), int(space.x_max), int(space.y_min), int(space.y_max),
                           1 if space.torous else 0, int(animal.params.content['delta_energy']),
                           float(animal.params.content['reproduction_rate']))
        e_in = input.content['energy_in'].reshape(n).to(torch.int32).contiguous()
        L.check(L.lib.fgx_animal_step(p, n, L.ptr(animal.active_state), L.ptr(e_in), _astate(animal),
                                      L.ptr(animal.age), L.stream()))
        return animal

    # sim.ipynb:124-133
    @staticmethod
    @native
    def remove_agent(params, animals, rows):
        n = animals.unique_id.shape[0]
        L.check(L.lib.fgx_animal_remove(n, L.ptr(rows.ids), L.ptr(rows.count), _astate(animals),
                                        L.ptr(animals.active_state), L.ptr(animals.age), L.stream()))
        return animals

    # sim.ipynb:135-156
    @staticmethod
    @native
    def add_agent(params, animals, rows, key):
        n = animals.unique_id.shape[0]
        L.check(L.lib.fgx_animal_add(n, L.ptr(params.content['copy_ids']), L.ptr(rows.start), L.ptr(rows.count),
                                     _astate(animals), L.ptr(animals.active_state), L.ptr(animals.age), L.stream()))
        return animals, key

    # sim.ipynb:158-170
    @staticmethod
    @native
    def set_agent(params, animals, rows, key):
        n = animals.unique_id.shape[0]
        L.check(L.lib.fgx_animal_set(n, L.ptr(rows.ids), L.ptr(rows.count), _astate(animals), L.stream()))
        return animals, key


@struct.dataclass
class Grass(Agent):
    # sim.ipynb:177-199
    @staticmethod
    @native
    def create_agent(params, unique_id, active_state, agent_type, key):
        n, dev = unique_id.shape[0], unique_id.device
        c = params.content
        i1 = lambda: torch.empty((n, 1), dtype=torch.int32, device=dev)  # noqa: E731
        st = {'X_pos': i1(), 'Y_pos': i1(), 'fully_grown': torch.empty(n, dtype=torch.bool, device=dev),
              'count_down': i1()}
        age = torch.empty(n, dtype=torch.float32, device=dev)
        L.check(L.lib.fgx_grass_create(n, int(c['space'].x_max), int(c['regrowth_time']), L.ptr(key),
                                       L.ptr(unique_id), L.ptr(st['X_pos']), L.ptr(st['Y_pos']),
                                       L.ptr(st['fully_grown'].view(torch.uint8)), L.ptr(st['count_down']), L.ptr(age),
                                       L.stream()))
        return Grass(unique_id=unique_id, active_state=active_state, agent_type=agent_type, state=State(content=st),
                     age=age, params=Params(content={'regrowth_time': c['regrowth_time']}), policy=None)

    # sim.ipynb:202-218 (in place)
    @staticmethod
    @native
    def step_agent(params, input, grass):
        n = grass.unique_id.shape[0]
        c = grass.state.content
        eo = input.content['energy_out'].reshape(n).to(torch.float32).contiguous()
        L.check(L.lib.fgx_grass_step(n, int(grass.params.content['regrowth_time']), L.ptr(eo),
                                     L.ptr(c['fully_grown'].view(torch.uint8)), L.ptr(c['count_down']), L.stream()))
        return grass


# sim.ipynb:238-290
def interaction(wolves, sheeps, grasses, space=None):
    ws, ss, gs = wolves.state.content, sheeps.state.content, grasses.state.content
    nw, ns, ng = wolves.unique_id.shape[0], sheeps.unique_id.shape[0], grasses.unique_id.shape[0]
    dev = ws['X_pos'].device
    if space is not None:
        xm, ym = int(space.x_max), int(space.y_max)
    else:  # the grass lattice spans the arena
        xm = int(gs['X_pos'].max()) + 1
        ym = int(gs['Y_pos'].max()) + 1
    w_in = torch.empty(nw, dtype=torch.int32, device=dev)
    s_in = torch.empty(ns, dtype=torch.int32, device=dev)
    s_eaten = torch.empty(ns, dtype=torch.float32, device=dev)
    g_eaten = torch.empty(ng, dtype=torch.float32, device=dev)
    nbytes = L.lib.fgx_wolf_sheep_interaction_scratch_bytes(xm, ym)
    scr = L.scratch(nbytes)
    L.check(L.lib.fgx_wolf_sheep_interaction(xm, ym, nw, L.ptr(ws['X_pos']), L.ptr(ws['Y_pos']), ns,
                                             L.ptr(ss['X_pos']), L.ptr(ss['Y_pos']), ng, L.ptr(gs['X_pos']),
                                             L.ptr(gs['Y_pos']), L.ptr(gs['fully_grown'].view(torch.uint8)),
                                             L.ptr(w_in), L.ptr(s_in), L.ptr(s_eaten), L.ptr(g_eaten), L.ptr(scr),
                                             nbytes, L.stream()))
    return w_in, s_in, s_eaten, g_eaten


jit_interaction = interaction


# sim.ipynb:365-517 (without the recording / orbax part)
class Ecosystem:
    def __init__(self, grass_regrowth_time, wolf_reproduction_rate, wolf_energy, init_wolves, sheep_reproduction_rate,
                 sheep_energy, init_sheeps, X_max, Y_max, sim_steps, key, max_animals=300):
        self.sim_steps = sim_steps
        self.space = create_space(x_min=0, x_max=X_max, y_min=0, y_max=Y_max, torous=True, wall_array=None)
        split = lambda k: [c.contiguous() for c in fgx_random.split(k)]  # noqa: E731
        key, grass_key = split(key)
        grass_num = self.space.x_max * self.space.y_max
        self.grasses = Agent_Set(agent=Grass, num_total_agents=grass_num, num_active_agents=grass_num, agent_type=0)
        grass_params = Params(content={'regrowth_time': grass_regrowth_time, 'space': self.space})
        self.grasses.agents = create_agents(params=grass_params, agent_set=self.grasses, key=grass_key)
        key, sheep_key = split(key)
        self.sheeps = Agent_Set(agent=Animal, num_total_agents=max_animals, num_active_agents=init_sheeps, agent_type=1)
        sheep_params = Params(content={'reproduction_rate': sheep_reproduction_rate, 'delta_energy': sheep_energy,
                                       'space': self.space})
        self.sheeps.agents = create_agents(params=sheep_params, agent_set=self.sheeps, key=sheep_key)
        key, wolf_key = split(key)
        self.wolves = Agent_Set(agent=Animal, num_total_agents=max_animals, num_active_agents=init_wolves, agent_type=2)
        wolf_params = Params(content={'reproduction_rate': wolf_reproduction_rate, 'delta_energy': wolf_energy,
                                      'space': self.space})
        self.wolves.agents = create_agents(params=wolf_params, agent_set=self.wolves, key=wolf_key)
        self.key = key  # the reference keeps the ORIGINAL key in self.key (sim.ipynb:378); see step()

    @staticmethod
    def select_dead_sheeps(sheeps_agents, select_params):
        return ((P(sheeps_agents.state.content['energy']) <= 0) | P(select_params.content['sheeps_eaten']).nonzero()) \
            & P(sheeps_agents.active_state).nonzero()

    @staticmethod
    def select_dead_wolves(wolves_agents, select_params):
        return (P(wolves_agents.state.content['energy']) <= 0) & P(wolves_agents.active_state).nonzero()

    @staticmethod
    def select_reproduce_animals(animals_agents, select_params):
        return P(animals_agents.state.content['reproduce']).nonzero() & P(animals_agents.active_state).nonzero()

    @staticmethod
    def add_animals(animal_agents, max_num, key):
        num_rep, rep_indx = jit_select_agents(Ecosystem.select_reproduce_animals, None, animal_agents)
        num_active = count_active(animal_agents)
        num_to_add = torch.minimum(max_num - num_active, num_rep)  # get_num_animals_to_add (sim.ipynb:433-438)
        add_params = Params(content={'copy_ids': rep_indx, 'num_active_agents': num_active})
        animal_agents, key = jit_add_agents(Animal.add_agent, num_agents_add=num_to_add, add_params=add_params,
                                            agents=animal_agents, key=key)
        animal_agents, key = jit_set_agents(Animal.set_agent, num_agents_set=num_to_add,
                                            set_params=Params(content={'set_ids': rep_indx}), agents=animal_agents,
                                            key=key)
        animal_agents, sorted_ids = jit_sort_agents(quantity=animal_agents.state.content['reproduce'], ascend=False,
                                                    agents=animal_agents)
        return animal_agents, num_active, key

    def step(self):
        w_in, s_in, s_eaten, g_eaten = jit_interaction(self.wolves.agents, self.sheeps.agents, self.grasses.agents,
                                                       self.space)
        self.grasses.agents = step_agents(params=None, agent_set=self.grasses, input=Signal(content={'energy_out': g_eaten}))
        sp = Params(content={'space': self.space})
        self.sheeps.agents = step_agents(params=sp, agent_set=self.sheeps, input=Signal(content={'energy_in': s_in}))
        self.wolves.agents = step_agents(params=sp, agent_set=self.wolves, input=Signal(content={'energy_in': w_in}))
        n_dead, idx = jit_select_agents(Ecosystem.select_dead_sheeps, Params(content={'sheeps_eaten': s_eaten}),
                                        self.sheeps.agents)
        self.sheeps.agents = jit_remove_agents(Animal.remove_agent, num_agents_remove=n_dead,
                                               remove_params=Params(content={'remove_ids': idx}),
                                               agents=self.sheeps.agents)
        self.sheeps.agents, _ = jit_sort_agents(quantity=self.sheeps.agents.active_state, ascend=False,
                                                agents=self.sheeps.agents)
        n_dead, idx = jit_select_agents(Ecosystem.select_dead_wolves, None, self.wolves.agents)
        self.wolves.agents = jit_remove_agents(Animal.remove_agent, num_agents_remove=n_dead,
                                               remove_params=Params(content={'remove_ids': idx}),
                                               agents=self.wolves.agents)
        self.wolves.agents, _ = jit_sort_agents(quantity=self.wolves.agents.active_state, ascend=False,
                                                agents=self.wolves.agents)
        self.sheeps.agents, num_active_sheeps, self.key = Ecosystem.add_animals(self.sheeps.agents,
                                                                                self.sheeps.num_total_agents, self.key)
        self.wolves.agents, num_active_wolves, self.key = Ecosystem.add_animals(self.wolves.agents,
                                                                                self.wolves.num_total_agents, self.key)
        return num_active_sheeps, num_active_wolves

    def capture(self, warmup=2):
        """One whole tick as a CUDA graph (``foragax_b200.graphs.GraphedTick``): ``warmup`` real ticks run
        first; every ``replay()`` is one tick and returns the static ``(num_active_sheeps, num_active_wolves)``."""
        from ..graphs import GraphedTick
        state = {'grass': self.grasses.agents, 'sheep': self.sheeps.agents, 'wolves': self.wolves.agents,
                 'key': self.key}

        def bind(st):
            self.grasses.agents, self.sheeps.agents, self.wolves.agents = st['grass'], st['sheep'], st['wolves']
            self.key = st['key']

        def tick(st):
            bind(st)
            out = self.step()
            return {'grass': self.grasses.agents, 'sheep': self.sheeps.agents, 'wolves': self.wolves.agents,
                    'key': self.key}, out

        g = GraphedTick(tick, state, warmup)
        bind(state)
        return g

    def run(self):
        num_sheep, num_wolf = [], []
        for _ in range(self.sim_steps):
            s, w = self.step()
            num_sheep.append(s)
            num_wolf.append(w)
        return num_sheep, num_wolf
