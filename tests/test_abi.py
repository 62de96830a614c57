"""The C-ABI library loads and exports every symbol include/foragax_b200.h declares
(no compute calls: runs without a GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "foragax_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fgx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header_symbols():
    from foragax_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 15
    missing = [s for s in syms if not hasattr(_lib.lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_version_and_error_text_without_gpu():
    from foragax_b200 import _lib
    assert _lib.lib.fgx_version() >= 1
    # argument validation happens on the host before any launch
    rc = _lib.lib.fgx_argsort(None, 99, 10, 0, None, None, 0, None)
    assert rc < 0
    assert b"dtype" in _lib.lib.fgx_last_error()


def test_every_entry_point_has_ctypes_signature():
    from foragax_b200 import _lib
    for s in declared_symbols():
        assert getattr(_lib.lib, s).argtypes is not None or s in ("fgx_last_error", "fgx_version"), s


def test_product_code_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "foragax_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), os.path.join(dirpath, f)


def test_reference_import_paths_resolve_to_the_native_mirror():
    """User code written against the reference (`import foragax.base.agent_methods as fgx_methods`) imports as is."""
    import foragax.base.agent_classes as fgx_classes
    import foragax.base.agent_methods as fgx_methods
    import foragax.base.space_methods as space_methods
    from foragax.policy.wilson_cowan import WilsonCowan
    import foragax_b200.base.agent_methods as native
    assert fgx_methods is native
    for name in ("create_agents", "step_agents", "jit_select_agents", "jit_remove_agents", "jit_add_agents",
                 "jit_set_agents", "jit_sort_agents"):
        assert callable(getattr(fgx_methods, name)), name
    for name in ("Signal", "Params", "State", "Policy", "Agent", "Agent_Set"):
        assert hasattr(fgx_classes, name), name
    for name in ("check_collision", "wall_generator", "RESOLUTION", "ray_generator", "ray_agent_sensor",
                 "ray_wall_sensor", "create_space", "jit_ray_generator", "jit_ray_wall_sensor", "jit_ray_agent_sensor",
                 "jit_check_collision"):
        assert hasattr(space_methods, name), name
    assert hasattr(WilsonCowan, "step_policy")
