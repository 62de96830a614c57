"""CPU oracle for the Foragax per-tick agent hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or the timed
CPU baseline.  ``foragax_b200`` never imports this package.

It restates, in NumPy (and C for the timed baseline, ``oracle/c``), the
algorithm of the reference's hot path (``/root/reference/foragax/base``,
``foragax/policy``) and the slice of third-party ``jax.random`` / ``jnp``
semantics whose results that path exposes (jax is absent from
``/root/reference`` and from this image; the reference pins only
``jax>=0.4.13`` in ``pyproject.toml:20-23``).

Parity pinning status
---------------------
* pinned: ``jax_random`` (threefry KATs, legacy counter layout) and the
  select / remove / sort / add chain, by the reference's own printed notebook
  output G1 (``examples/basic/hello_world/hello_world.ipynb:92-99``) and by G2
  ticks 0-3 (``examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.ipynb``).
* parity unpinned: ``space_methods`` geometry (``ray_wall_sensor``,
  ``ray_agent_sensor``, ``ray_generator``, ``check_collision``), the
  WilsonCowan policy and ``normal`` -- the reference has no tests or golden
  vectors for them and cannot be executed here (no jax); they are restated
  from source and checked on hand-derived cases only.
"""
