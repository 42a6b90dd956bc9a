"""GPU parity at BASELINE.json's full sizes (every instance, every mesh node, every time step).

The whole batch runs through the public `pdepe`; the oracle re-solves an evenly spaced sample of the
instances (1e-10 relative, identical total Newton iteration counts), and the complete batch is checked
through size-independent properties of the discretisation:

* the implicit-Euler residual of the last step, evaluated with the separate `assemble!` kernel, vanishes;
* cfg2 (zero-flux boundaries): sum_i vol_i u_i is conserved, and the solution stays even in x;
* cfg3a/3b (Dirichlet at x = 1): the boundary value is met; cfg3b's exact solution x^2 + 6 D t is
  reproduced (the singular m = 2 interpolation is exact for it);
* cfg4: boundary conditions met, 0 <= u2 <= u1 <= 1.
"""
import numpy as np
import pytest
from scipy.special import j0

import skeelberzins_jl_b200 as sb
from skeelberzins_jl_b200 import examples as ex

pytestmark = pytest.mark.gpu

RTOL = 1e-10


@pytest.fixture(scope="module")
def ctx():
    return sb.Context(0)


@pytest.mark.parametrize("cfg", ["2", "3a", "3b", "4"])
def test_full_size_sampled_parity_and_properties(ctx, oracle, cfg):
    c = ex.baseline_case(cfg)
    B, Nx, P = c.params.shape[0], c.xmesh.size, c.npde
    assert (B, Nx) == {"2": (4096, 1024), "3a": (4096, 512), "3b": (4096, 512), "4": (16384, 256)}[cfg]
    tg, _ = sb.time_grid(c.tspan, c.tstep)
    assert tg.size - 1 == {"2": 90, "3a": 100, "3b": 80, "4": 1000}[cfg]
    fn = sb.Functor(c.functor, c.params)

    # all time steps but the last through pdepe, the last one by hand so that u_n and u_{n+1} are both at hand
    sol, pb, dev = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-2]), ctx=ctx, tstep=tg[:-1], save="last",
                            squeeze=False, data=True, return_device=True)
    try:
        u_n = dev.download().copy()                                  # [B, Nx, P]
        tau = tg[-1] - tg[-2]
        sb.newton(pb, tau, tg[-1])
        u = dev.download().copy()
        iters = dev.total_iters()
        assert (dev.status() == 0).all()
        assert np.isfinite(u).all()

        # -- sampled instances against the oracle -----------------------------------------------------------
        ks = np.unique(np.append(np.linspace(0, B - 1, 12).astype(int), B // 2))
        r = oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params[ks], c.initial()[ks], tg, None)
        uo = r["u"].reshape(ks.size, Nx, P)
        err = np.abs(u[ks] - uo).reshape(ks.size, -1).max(axis=1) / np.abs(uo).reshape(ks.size, -1).max(axis=1)
        assert err.max() < RTOL
        assert (iters[ks] == r["iters"].sum(axis=1)).all()

        # -- residual of the last implicit-Euler step with the stand-alone assemble! kernel -----------------
        M = dev.mass().reshape(B, Nx, P)
        rhs = dev.assemble(u, tg[-1]).reshape(B, Nx, P)
        udot = M * (u - u_n) / tau
        res = np.abs(udot - rhs).reshape(B, -1).max(axis=1)
        scale = np.maximum(np.abs(udot).reshape(B, -1).max(axis=1), np.abs(rhs).reshape(B, -1).max(axis=1))
        scale = np.maximum(scale, 1e-4 * np.abs(u).reshape(B, -1).max(axis=1) / tau)   # near a steady state udot -> 0
        assert (res / scale).max() < 1e-7, (res / scale).max()

        x = c.xmesh
        t_end = tg[-1]
        if cfg == "2":
            zeta = 0.5 * (x[1:] + x[:-1])
            vol = np.diff(np.concatenate([x[:1], zeta, x[-1:]]))
            mass0 = (c.initial()[:, :, 0] * vol).sum(axis=1)
            mass1 = (u[:, :, 0] * vol).sum(axis=1)
            assert np.abs(mass1 / mass0 - 1.0).max() < 1e-9
            assert np.abs(u[:, :, 0] - u[:, ::-1, 0]).max() < 1e-9 * np.abs(u).max()
        elif cfg == "3a":
            D = c.params[:, 0]
            bnd = j0(ex.BESSEL_ZERO) * np.exp(-D * ex.BESSEL_ZERO ** 2 * t_end)
            assert np.abs(u[:, -1, 0] - bnd).max() < 1e-12
            exact = j0(ex.BESSEL_ZERO * x)[None, :] * np.exp(-D * ex.BESSEL_ZERO ** 2 * t_end)[:, None]
            assert np.abs(u[:, :, 0] - exact).max() < 2e-2              # O(tau) accurate, Example103's own tolerance
        elif cfg == "3b":
            D = c.params[:, 0]
            exact = x[None, :] ** 2 + 6.0 * D[:, None] * t_end
            assert np.abs(u[:, :, 0] - exact).max() < 1e-8 * np.abs(exact).max()
        else:
            assert np.abs(u[:, 0, 0] - 1.0).max() < 1e-12             # u1(0) = 1
            assert np.abs(u[:, -1, 1]).max() < 1e-12                  # u2(1) = 0
            assert u.min() > -1e-12 and u.max() < 1.0 + 1e-12
            assert (u[:, :, 0] - u[:, :, 1]).min() > -1e-12
    finally:
        pb.device = None
        dev.close()
