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
* pinned by the reference's own printed notebook outputs:
  G1 (``examples/basic/hello_world/hello_world.ipynb:92-99``; ``tests/test_oracle_g1.py``) pins
  ``jax_random`` (legacy counter layout) and the select / remove / stable sort / add chain
  bit-exactly; G2 (``examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.ipynb`` cell
  output; ``tests/test_oracle_g2.py``) is reproduced for ticks 0-40 by the full restatement in
  ``foraging.py`` and thereby pins -- through integer counts that depend on them -- ``uniform``
  with bounds, ``bernoulli``, ``normal``, the serial key chain of ``add_agents``, the sensor
  composition (``ray_generator``, ``ray_wall_sensor``, ``ray_agent_sensor``, first-argmin) and the
  WilsonCowan policy.  Threefry is also pinned by the Random123 known-answer vectors.
* parity unpinned (no reference golden data exists; restated from source, hand-derived cases
  only): ``check_collision``, the collinear special cases of ``ray_wall_sensor`` that G2 does not
  exercise, NaN / -0.0 / int-key sort edge cases.
"""
