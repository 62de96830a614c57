"""Compile-only check of the XLA-FFI shim (north_star: "thin jax.ffi C-ABI layer").  jaxlib is not installed in
this image, so the shim is checked against a stand-in header that types the FFI binding builder: a handler whose
parameter list drifts from its binding, or a launcher call that drifts from include/foragax_b200.h, fails here."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "foragax_b200", "csrc", "xla_ffi", "xla_ffi_shim.cc")


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_shim_compiles_against_the_stand_in_header():
    r = subprocess.run(["make", "-s", "xla-ffi-check"], cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


def test_shim_covers_the_per_tick_path():
    src = open(SHIM).read()
    handlers = set(re.findall(r"XLA_FFI_DEFINE_HANDLER_SYMBOL\((\w+),", src))
    want = {"FgxSelect", "FgxArgsort", "FgxTakeRows", "FgxRaycastWalls", "FgxRaycastWallsGrid", "FgxRaycastAgents",
            "FgxThreefrySplit", "FgxThreefryUniform", "FgxThreefryRandint", "FgxCountActive", "FgxDiceStep",
            "FgxDiceRemove", "FgxDiceAdd", "FgxForagerStep", "FgxForagerRemove", "FgxForagerAdd", "FgxForagerSet",
            "FgxSensorModel", "FgxForagerSenseStep"}
    assert want <= handlers, sorted(want - handlers)
    # every launcher the handlers forward to is declared in the public header
    hdr = open(os.path.join(ROOT, "include", "foragax_b200.h")).read()
    for fn in set(re.findall(r"\b(fgx_[a-z0-9_]+)\(", src)):
        assert re.search(r"\b%s\(" % fn, hdr), fn
