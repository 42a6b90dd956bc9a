"""post-processing of stored solutions -- pdeval (src/utils.jl:396-438), sol(t) (:575-597),
sol(x, t, pb) (:460-561) -- evaluated on the GPU, against the oracle's restatement and the reference's golden."""
import numpy as np
import pytest

import skeelberzins_jl_b200 as sb
from skeelberzins_jl_b200 import examples as ex

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    return sb.Context(0)


def test_example201_golden_through_pdeval_and_sol_call(ctx):
    """examples/Example201_PartialDerivativeApprox.jl:63-110: err3 (pdeval) ~ err4 (sol(x,t,pb)) ~ 0.6621164947313215."""
    c = ex.example201()
    n = c.xmesh.size
    fn = sb.Functor(c.functor, c.params)
    sol, pb = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, c.tspan, ctx=ctx, data=True, tstep=c.tstep)
    assert np.array_equal(sol.t, sb.julia_range(0, 1 / (n - 1), 1))
    e = c.extra
    t = ex.linspace(0, 1, n)
    It = np.zeros(n)
    for k in range(1, 41):
        mm = (k * np.pi) ** 2 + 0.25 * e["eta"] ** 2
        It = It + ((k * np.pi) ** 2 / mm) * np.exp(-(e["D"] / e["L"] ** 2) * mm * t)
    It = 2 * e["Ip"] * ((1 - np.exp(-e["eta"])) / e["eta"]) * It
    d3, d4 = np.zeros(n), np.zeros(n)
    for i in range(n):
        _, d3[i] = sb.pdeval(pb.m, pb.xmesh, sol.u[i], 0, pb)
        _, d4[i] = sol(0, sol.t[i], pb)
    scale = e["Ip"] * e["D"] / e["K"]
    err3 = np.abs(It[1:n] - scale * d3[1:n]).max()
    err4 = np.abs(It[1:n] - scale * d4[1:n]).max()
    assert err3 == pytest.approx(0.6621164947313215, rel=1e-10)
    assert err4 == pytest.approx(0.6621164947313215, rel=1e-10)


@pytest.mark.parametrize("m,singular", [(0, False), (1, True), (1, False), (2, True), (2, False)])
def test_pdeval_matches_oracle(ctx, oracle, m, singular):
    rng = np.random.default_rng(5 + m)
    Nx, B = 300, 4
    lo = 0.0 if (singular or m == 0) else 0.5
    x = np.linspace(lo, lo + 2.0, Nx)
    x[1:-1] += (rng.random(Nx - 2) - 0.5) * 0.4 * (x[1] - x[0])
    fid = ex.LIN_DIFF
    u = rng.random((B, Nx, 1))
    dev = sb.DeviceProblem(ctx, m, x, sb.Functor(fid, np.ones((B, 1))))
    xq = np.concatenate([[x[0], x[-1], x[7], x[0] + 1e-13 * (x[1] - x[0])], x[0] + (x[-1] - x[0]) * rng.random(50)])
    uo, dxo = dev.pdeval(u, xq)
    ro, rdx = oracle.pdeval(fid, m, x, u, xq)
    assert np.abs(uo - ro).max() < 1e-13 * np.abs(ro).max()
    assert np.abs(dxo - rdx).max() < 1e-12 * np.abs(rdx).max()
    dev.upload(u)
    uo2, dxo2 = dev.pdeval(None, xq)                                 # resident state
    assert np.array_equal(uo, uo2) and np.array_equal(dxo, dxo2)
    with pytest.raises(ValueError, match="oustide"):
        dev.pdeval(u, [x[-1] + 1.0])
    with pytest.raises(ValueError, match="oustide"):
        dev.pdeval(u, [x[0] - 1.0])


def test_sol_time_interpolation_and_system(ctx, oracle):
    c = ex.example106(41)
    fn = sb.Functor(c.functor, c.params)
    sol, pb = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (0.0, 0.2), ctx=ctx, data=True, tstep=1e-2)
    tq = 0.1234
    ref = oracle.interpolate_sol_time(sol.t, sol.u, tq)
    assert np.array_equal(sol(tq), ref)
    assert np.array_equal(sol(sol.t[0]), sol.u[0])
    with pytest.raises(ValueError, match="not within the interval"):
        sol(5.0)
    xq = np.array([0.0, 0.3, 0.77, 1.0])
    u, du = sol(xq, tq, pb)                                          # npde = 2: [nq, npde]
    st = np.transpose(ref, (1, 0))[None]                             # [1, Nx, npde]
    ro, rdx = oracle.pdeval(c.functor, c.m, c.xmesh, st, xq)
    assert np.abs(u - ro[0]).max() < 1e-14 and np.abs(du - rdx[0]).max() < 1e-12
    only_val = sol(xq, tq, pb, deriv=False)
    assert np.array_equal(only_val, u)


def test_diffeq_operator_and_alias_compat(ctx, oracle):
    """the ODEFunction/ODEProblem adapter (src/diffEq.jl:18-57) and the reference's stored-value quirk."""
    c = ex.example103(60)
    fn = sb.Functor(c.functor, c.params)
    pb = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, c.tspan, ctx=ctx, solver="DiffEq")     # returns pb, src/pdepe.jl:123-126
    f, u0, tspan = sb.ODEProblem(pb)
    assert tspan == (0.0, 1.0) and u0.shape == (1, 60)
    du = np.empty_like(u0)
    f(du, u0, None, 0.25)
    ref = oracle.assemble(c.functor, c.m, c.xmesh, c.params, u0.reshape(1, 60, 1), 0.25)
    assert np.abs(du.reshape(-1) - ref.reshape(-1)).max() < 1e-12 * np.abs(ref).max()
    assert f.mass_matrix is not None and f.mass_matrix[0, -1] == 0 and f.mass_matrix[0, 0] == 1     # Dirichlet right end
    J = f.jac_prototype
    assert J.shape == (60, 60) and J.nnz == 3 * 60 - 2                                              # src/utils.jl:725-759
    assert sb.jac_prototype(21, 2).nnz == 4 * (3 * 21 - 2)
    assert sb.reshape_solution(np.arange(12.0).reshape(1, 12), type("P", (), {"Nx": 6, "npde": 2})).shape == (1, 2, 6)
    true = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, c.tspan, ctx=ctx)
    ali = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, c.tspan, ctx=ctx, alias_compat=True)
    assert np.array_equal(true.u[-1], ali.u[-1])
    assert ali.u[5][-1] == 0.0 and true.u[5][-1] != 0.0 and np.array_equal(ali.u[5][:-1], true.u[5][:-1])


def test_nonuniform_vector_tstep(ctx, oracle):
    """Params.tstep as a vector is the time grid itself, tau_i = diff (src/pdepe.jl:85,92)."""
    c = ex.baseline_case("2", B=6, Nx=200)
    grid = np.array([0.001, 0.0011, 0.0013, 0.0018, 0.002, 0.0027])
    fn = sb.Functor(c.functor, c.params)
    sol = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (grid[0], grid[-1]), ctx=ctx, tstep=grid, squeeze=False)
    r = oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params, c.initial(), grid, None, snapshots=True)
    assert len(sol) == grid.size
    for k in range(grid.size):
        assert np.abs(sol.u[k] - r["snapshots"][:, k, :, 0]).max() < 1e-10 * np.abs(r["snapshots"]).max()
