"""Runs last (file name): verdict of guard mode over the whole GPU test session.

tests/conftest.py sets SB_GUARD=1, so every device array of every problem the tests created was allocated between two
4 KiB canary regions; the library checks them when an array is freed.  No kernel of the session may have written outside
its arrays -- the substitute for a compute-sanitizer memcheck run on pools where that tool is not available.
"""
import gc

import pytest

import skeelberzins_jl_b200 as sb


@pytest.mark.gpu
def test_no_kernel_wrote_outside_its_arrays():
    gc.collect()
    sb.clear_pool()
    gc.collect()
    violations, checked = sb.guard_violations()
    assert checked > 0, "guard mode was not active (SB_GUARD must be set before the first library call)"
    assert violations == 0, "%d of %d device arrays had their canary bytes overwritten" % (violations, checked)


@pytest.mark.gpu
def test_guard_mode_detects_a_spoiled_canary():
    """Positive control in a fresh process: SB_GUARD_SELFTEST makes the library overwrite one canary byte of the first
    array it allocates, as an out-of-bounds store would; the check at free time must count it."""
    import os
    import subprocess
    import sys
    code = ("import skeelberzins_jl_b200 as sb\n"
            "from skeelberzins_jl_b200 import examples as ex\n"
            "c = ex.example101()\n"
            "fn = sb.Functor(c.functor, c.params)\n"
            "sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, c.tspan, tstep=c.tstep)\n"
            "sb.clear_pool()\n"
            "v, n = sb.guard_violations()\n"
            "assert v == 1 and n > 1, (v, n)\n")
    env = dict(os.environ, SB_GUARD="1", SB_GUARD_SELFTEST="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
