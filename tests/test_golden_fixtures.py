"""Committed fixtures under tests/golden/.

* reference_goldens.json -- the literals the reference's own tests and examples assert (file:line inside);
* oracle_vectors.npz     -- small input/output vectors of the CPU oracle, written by tests/golden/make_oracle_vectors.py.

CPU: the oracle built here reproduces the committed vectors and their sums are the reference's literals.
GPU: the CUDA path (through the C ABI) reproduces the committed vectors -- no oracle call on that side.
"""
import importlib.util
import json
import os

import numpy as np
import pytest

import skeelberzins_jl_b200 as sb
from skeelberzins_jl_b200 import examples as ex

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_goldens.json")))
VEC = np.load(os.path.join(HERE, "golden", "oracle_vectors.npz"))
_spec = importlib.util.spec_from_file_location("make_oracle_vectors", os.path.join(HERE, "golden", "make_oracle_vectors.py"))
gen = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(gen)

TRANSIENT = [("ex101", ex.example101, "example101_sum"), ("ex102", ex.example102, "example102_sum"),
             ("ex103", ex.example103, "example103_sum"), ("ex106", ex.example106, "example106_sum"),
             ("ex107", ex.example107, None)]
STATIONARY = [("ex104", lambda: ex.example104(c0=1.0), "example104_sum"), ("ex104c0", lambda: ex.example104(c0=0.0), "example104_sum"),
              ("ex105", ex.example105, "example105_sum")]


def test_committed_states_carry_the_reference_literals():
    for key, _, gold in TRANSIENT + STATIONARY:
        if gold is not None:
            assert VEC[key + "_u"].sum() == pytest.approx(GOLD[gold]["value"], rel=1e-12), key
    c = ex.example107()
    assert np.linalg.norm(VEC["ex107_u"][:, 0] - c.extra["exact"](c.xmesh, 0.8)) < GOLD["example107_err_bound"]["value"]
    # the examples module quotes the same literals
    for make, gold in ((ex.example101, "example101_sum"), (ex.example102, "example102_sum"), (ex.example103, "example103_sum"),
                       (ex.example105, "example105_sum"), (ex.example106, "example106_sum")):
        assert make().golden == GOLD[gold]["value"]


def test_oracle_reproduces_committed_vectors(oracle):
    for key, make, _ in TRANSIENT:
        c = make()
        tg, tau = sb.time_grid(c.tspan, c.tstep)
        r = oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params, c.initial(), tg, tau)
        assert np.allclose(r["u"][0], VEC[key + "_u"], rtol=1e-13, atol=0)
        assert np.array_equal(r["iters"][0], VEC[key + "_iters"])
    for i, (fid, m, s) in enumerate(gen.KERNEL_CASES):
        x, prm, u, b = gen.kernel_case(fid, m, s, 1000 + i)
        key = "k_f%d_m%d_s%d" % (fid, m, int(s))
        du = oracle.assemble(fid, m, x, prm, u, 0.3)
        assert np.allclose(du, VEC[key + "_du"], rtol=1e-13, atol=1e-13 * np.abs(VEC[key + "_du"]).max())
        mass = oracle.mass(fid, m, x, prm, u, 0.3)
        F, Jl, Jd, Ju = oracle.residual_jacobian(fid, m, x, prm, u, b, 0.01, mass, 0.3)
        J = np.stack([Jl, Jd, Ju])
        assert np.allclose(F, VEC[key + "_F"], rtol=1e-13, atol=1e-13 * np.abs(VEC[key + "_F"]).max())
        assert np.allclose(J, VEC[key + "_J"], rtol=1e-13, atol=1e-13 * np.abs(VEC[key + "_J"]).max())


@pytest.fixture(scope="module")
def ctx():
    return sb.Context(0)


@pytest.mark.gpu
@pytest.mark.parametrize("design", ["fused_tma", "fused", "split"])
def test_gpu_reproduces_committed_states(ctx, design):
    for key, make, _ in TRANSIENT:
        c = make()
        fn = sb.Functor(c.functor, c.params)
        sol, dev = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, c.tspan, ctx=ctx, design=design, tstep=c.tstep, squeeze=False,
                            save="last", return_device=True)
        u = dev.download()[0]
        ref = VEC[key + "_u"]
        assert np.abs(u - ref).max() < 1e-10 * np.abs(ref).max(), key
        assert dev.total_iters()[0] == VEC[key + "_iters"].sum(), key
        dev.close()
    for key, make, _ in STATIONARY:
        c = make()
        fn = sb.Functor(c.functor, c.params)
        u, hist = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, ctx=ctx, design=design, hist=True)
        ref = VEC[key + "_u"][:, 0]
        assert np.abs(u - ref).max() < 1e-10 * np.abs(ref).max(), key
        assert len(hist) == VEC[key + "_iters"][0], key


@pytest.mark.gpu
def test_gpu_reproduces_committed_kernel_vectors(ctx):
    """assemble!, residual and block-tridiagonal Jacobian of seeded random states: every registry functor, every
    coordinate system, against the committed oracle output."""
    for i, (fid, m, s) in enumerate(gen.KERNEL_CASES):
        x, prm, u, b = gen.kernel_case(fid, m, s, 1000 + i)
        key = "k_f%d_m%d_s%d" % (fid, m, int(s))
        dev = sb.DeviceProblem(ctx, m, x, sb.Functor(fid, prm), design="split")
        du = dev.assemble(u, 0.3)
        ref = VEC[key + "_du"]
        assert np.abs(du - ref).max() <= 1e-12 * np.abs(ref).max(), key
        dev.close()
