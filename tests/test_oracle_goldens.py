"""Pin the CPU oracle against EVERY golden value the reference's own tests hold for the internal
implicit-Euler / stationary-Newton path.

The reference compares with `≈` (rtol = sqrt(eps) ~ 1.5e-8); the oracle is held to 1e-12.
Iteration counts are checked against the structure the algorithm dictates (linear problems: exactly 2
Newton iterations per step, src/utils.jl:323-333) and the survey probe's counts.
"""
import numpy as np
import pytest

import skeelberzins_jl_b200 as sb
from skeelberzins_jl_b200 import examples as ex

RTOL = 1e-12


def _transient(oracle, c, **kw):
    tg, tau = sb.time_grid(c.tspan, c.tstep)
    return tg, oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params, c.initial(), tg, tau, **kw)


def test_example101_linear_diffusion(oracle):
    # examples/Example101_LinearDiffusion.jl:64,69
    c = ex.example101()
    tg, r = _transient(oracle, c)
    assert tg.size == 101 and tg[37] == 37 / 100
    assert r["u"].sum() == pytest.approx(3.721004873950427, rel=RTOL)
    assert (r["iters"] == 2).all()


def test_transient_tstep_variants(oracle):
    # test/test_solveTransient.jl:38-40,45-46,63-65: tstep=1e-3 scalar and vector collect(0:1e-3:1), hist sums equal
    c = ex.example101(tstep=1e-3)
    tg, r1 = _transient(oracle, c, hist=True)
    assert tg.size == 1001
    c2 = ex.example101(tstep=sb.julia_range(0, 1e-3, 1))
    tg2, r2 = _transient(oracle, c2, hist=True)
    assert np.array_equal(tg, tg2)
    assert r1["u"].sum() == pytest.approx(3.721004873950427, rel=RTOL)
    assert r2["u"].sum() == pytest.approx(3.721004873950427, rel=RTOL)
    # banded variant literal (test/test_solveTransient.jl:64)
    assert r1["u"].sum() == pytest.approx(3.7210048739504265, rel=RTOL)
    assert np.nansum(r1["hist"]) == pytest.approx(np.nansum(r2["hist"]), rel=1e-9)


def test_example102_nonlinear_diffusion(oracle):
    # examples/Example102_NonlinearDiffusion.jl:84
    c = ex.example102()
    tg, r = _transient(oracle, c)
    assert tg.size == 91 and tg[0] == 0.001 and tg[-1] == 0.01
    assert r["u"].sum() == pytest.approx(46.66666666678757, rel=RTOL)
    assert r["iters"].sum() == 294  # SURVEY section 4 probe: 4 x 21 then 3


def test_example103_cylindrical(oracle):
    # examples/Example103_LinearDiffusionCylindrical.jl:75
    c = ex.example103()
    _, r = _transient(oracle, c)
    assert r["u"].sum() == pytest.approx(0.0463188424523652, rel=RTOL)
    assert (r["iters"] == 2).all()


@pytest.mark.parametrize("c0", [1.0, 0.0])
def test_example104_poisson_and_c0_variant(oracle, c0):
    # examples/Example104_Poisson.jl:58 ; test/test_solveStationary.jl:37 (c = 0)
    c = ex.example104(c0=c0)
    r = oracle.pdepe_stationary(c.functor, c.m, c.xmesh, c.params, c.initial(), hist=True)
    assert r["u"].sum() == pytest.approx(3.7624999999999997, rel=RTOL)
    assert r["iters"][0] == 2


def test_example105_stationary_nonlinear(oracle):
    # examples/Example105_StationaryNonlinearDiffusion.jl:61
    c = ex.example105()
    r = oracle.pdepe_stationary(c.functor, c.m, c.xmesh, c.params, c.initial(), hist=True)
    assert r["u"].sum() == pytest.approx(6.025575019008793, rel=RTOL)
    assert r["iters"][0] == 7
    x = c.xmesh
    assert np.allclose(r["u"][0, :, 0], np.sqrt(0.01 + x * (1 - x) / 2), rtol=0, atol=1e-15)


def test_example106_reaction_diffusion_system(oracle):
    # examples/Example106_SystemReactionDiffusion.jl:84
    c = ex.example106()
    _, r = _transient(oracle, c)
    assert r["u"].sum() == pytest.approx(29.034702247833415, rel=RTOL)
    assert (r["iters"] == 2).all()


def test_example107_spherical(oracle):
    # examples/Example107_LinearDiffusionSpherical.jl:67-68: norm(u(T) - (x^2 + 6T)) < 1e-2
    c = ex.example107()
    _, r = _transient(oracle, c)
    err = np.linalg.norm(r["u"][0, :, 0] - c.extra["exact"](c.xmesh, 0.8))
    assert err < 1.0e-2
    assert err < 1e-13  # the scheme is exact for this solution


def test_example201_boundary_flux_error(oracle):
    # examples/Example201_PartialDerivativeApprox.jl:63-110: err3 = ||seriesI - (Ip*D/K)*dudx(x=0,t_k)||_inf
    c = ex.example201()
    n = c.xmesh.size
    tg, r = _transient(oracle, c, snapshots=True)
    assert tg.size == n
    e = c.extra
    t = ex.linspace(0, 1, n)
    It = np.zeros(n)
    for k in range(1, 41):  # analytical(t), :68-75
        mm = (k * np.pi) ** 2 + 0.25 * e["eta"] ** 2
        It = It + ((k * np.pi) ** 2 / mm) * np.exp(-(e["D"] / e["L"] ** 2) * mm * t)
    It = 2 * e["Ip"] * ((1 - np.exp(-e["eta"])) / e["eta"]) * It
    S = r["snapshots"][0, :, :, 0]                      # [nt, Nx]
    h = c.xmesh[1] - c.xmesh[0]
    dudx = (S[:, 1] - S[:, 0]) / h                       # pdeval at x=0: src/utils.jl:404-405 -> :26-30
    dudx = (e["Ip"] * e["D"] / e["K"]) * dudx
    err = np.abs(It[1:n] - dudx[1:n]).max()
    assert err == pytest.approx(0.6621164947313215, rel=1e-10)


def test_quadrature_closed_forms(oracle):
    # src/utils.jl:683-718
    x = ex.linspace(0, 1, 11)
    xi, ze, sing = oracle.quadrature(0, x)
    assert not sing and np.array_equal(xi, (x[:-1] + x[1:]) / 2) and np.array_equal(xi, ze)
    xi, ze, sing = oracle.quadrature(1, x)
    assert sing and ze[0] == 0.0 and xi[0] == pytest.approx(2 / 3 * x[1])
    xi, ze, sing = oracle.quadrature(2, x)
    assert sing and ze[0] == 0.0
    a, b = x[1:-1], x[2:]
    assert np.allclose(ze[1:], (a * b * (a + b) / 2) ** (1 / 3), rtol=1e-15)
    xr = ex.linspace(1, 2, 11)
    xi, ze, sing = oracle.quadrature(1, xr)
    a, b = xr[:-1], xr[1:]
    assert not sing and np.allclose(xi, (b - a) / np.log(b / a), rtol=1e-15)
    xi, ze, sing = oracle.quadrature(2, xr)
    assert np.allclose(xi, a * b * np.log(b / a) / (b - a), rtol=1e-15)


def test_manufactured_regular_m1_m2(oracle):
    # parity-UNPINNED branches (regular m=1,2: src/utils.jl:32-45, :700-701,:712): validated against mathematics only.
    # u = x^2 + 2(m+1) t solves u_t = x^-m (x^m u_x)_x; Dirichlet data both ends (SURVEY section 4 KAT (i),(ii)).
    for m, g in ((1, 4.0), (2, 6.0)):
        for N in (21, 41):
            x = ex.linspace(1, 2, N)
            prm = ex.generic_params(al=1.0, gl0=x[0] ** 2, gl1=g, ql=0.0, ar=1.0, gr0=x[-1] ** 2, gr1=g, qr=0.0)
            tg = sb.julia_range(0, 1e-2, 0.5)
            r = oracle.pdepe_transient(ex.GENERIC_SCALAR, m, x, prm[None, :], (x ** 2)[None, :, None], tg, 1e-2)
            assert np.abs(r["u"][0, :, 0] - (x ** 2 + g * 0.5)).max() < 5e-14
            assert (r["iters"] == 2).all()


def test_manufactured_regular_m2_stationary_flux_bc(oracle):
    # SURVEY section 4 KAT (iii): c = 0, left flux BC p = -1/xl^2, q = 1; right Dirichlet 3 - 1/xr -> u = 3 - 1/x
    x = ex.linspace(1, 2, 33)
    prm = ex.generic_params(c0=0.0, al=0.0, gl0=1.0 / x[0] ** 2, ql=1.0, ar=1.0, gr0=3 - 1 / x[-1], qr=0.0)
    r = oracle.pdepe_stationary(ex.GENERIC_SCALAR, 2, x, prm[None, :], np.ones((1, x.size, 1)))
    assert np.abs(r["u"][0, :, 0] - (3 - 1 / x)).max() < 1e-14
    assert r["iters"][0] == 2


def test_convergence_failure_reported(oracle):
    # src/utils.jl:336: throw("convergence failed") after maxit
    c = ex.example105()
    r = oracle.pdepe_stationary(c.functor, c.m, c.xmesh, c.params, c.initial(), maxit=3)
    assert r["status"][0] == -1
