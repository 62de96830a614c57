"""Drop-in import alias for reference user code: ``import foragax.base.agent_methods as fgx_methods``,
``from foragax.policy.wilson_cowan import WilsonCowan`` ... resolve to the ``foragax_b200`` modules that
mirror the reference's ``foragax/base`` and ``foragax/policy`` (same names, B200-native ops underneath)."""
import importlib
import sys

import foragax_b200  # noqa: F401  (loads libforagax_b200.so; raises if it is missing)

for _name in ("base", "base.agent_classes", "base.agent_methods", "base.space_classes", "base.space_methods",
              "policy", "policy.random_policy", "policy.wilson_cowan"):
    _mod = importlib.import_module("foragax_b200." + _name)
    sys.modules["foragax." + _name] = _mod
    _parent, _, _leaf = _name.rpartition(".")
    setattr(sys.modules["foragax." + _parent] if _parent else sys.modules[__name__], _leaf, _mod)
