"""Plugin slot for user coefficient functors: CUDA source compiled at run time (NVRTC, sm_100a) behind the same
kernels.  Checked against the oracle, which carries the same callbacks as an oracle-only registry entry."""
import numpy as np
import pytest

import skeelberzins_jl_b200 as sb
from skeelberzins_jl_b200 import examples as ex

pytestmark = pytest.mark.gpu

EXP_DIFF_SRC = r"""
struct ExpDiff {
    static constexpr int npde = 1, nparams = 3;
    static constexpr bool c_unit = true, s_zero = false;
    template <class S> static __device__ void pde(double x, double t, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[0] * exp(u[0]) * ux[0]; s[0] = -(p[1] * u[0] * u[0]) + x;
    }
    template <class S> static __device__ void bc(double, const S* ul, double, const S* ur, double t, const double* p,
                                                 S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0] - p[2] * (1.0 + t); ql[0] = 0.0; pr[0] = S(0.0); qr[0] = 1.0;
    }
};
"""
ORACLE_EXP_DIFF = 10

NONLIN_CLONE_SRC = r"""
struct MyNonlinDiff {   // same callbacks as examples/Example102_NonlinearDiffusion.jl:44-65, supplied by the user
    static constexpr int npde = 1, nparams = 1;
    static constexpr bool c_unit = false, s_zero = false;   // no traits claimed: generic code path
    template <class S> static __device__ void pde(double, double, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); f[0] = p[0] * u[0] * ux[0]; s[0] = S(0.0);
    }
    template <class S> static __device__ void bc(double, const S*, double, const S*, double, const double*, S* pl, double* ql, S* pr, double* qr) {
        pl[0] = S(0.0); ql[0] = 1.0; pr[0] = S(0.0); qr[0] = 1.0;
    }
};
"""


@pytest.fixture(scope="module")
def ctx():
    return sb.Context(0)


@pytest.mark.parametrize("design", ["fused_tma", "fused", "split"])
@pytest.mark.parametrize("m,Nx", [(0, 200), (2, 700)])
def test_user_functor_matches_oracle(ctx, oracle, design, m, Nx):
    fid = sb.register_user_functor(EXP_DIFF_SRC, "ExpDiff", 1, 3)
    assert fid >= 1000 and sb.functor_info(fid) == (1, 3, "ExpDiff")
    B = 6
    k = np.arange(B)
    prm = np.stack([0.5 + 0.1 * k, 1.0 + 0.2 * k, 0.3 + 0.05 * k], axis=1)
    x = ex.linspace(0, 1, Nx)
    ic = 0.3 * np.cos(x) ** 2
    tg = sb.julia_range(0.0, 1e-2, 0.1)
    fn = sb.Functor(fid, prm)
    sol, hist, dev = sb.pdepe(m, fn, ic, fn, x, (0.0, 0.1), ctx=ctx, design=design, tstep=1e-2, hist=True, squeeze=False,
                              save="last", return_device=True)
    u0 = np.broadcast_to(ic[None, :, None], (B, Nx, 1))
    r = oracle.pdepe_transient(ORACLE_EXP_DIFF, m, x, prm, u0, tg, 1e-2)
    err = np.abs(sol.u[-1] - r["u"][:, :, 0]).max() / np.abs(r["u"]).max()
    assert err < 1e-10
    its = np.array([[len(h[i]) for h in hist] for i in range(B)])
    assert np.array_equal(its, r["iters"])
    du = dev.assemble(r["u"], 0.05)
    assert np.abs(du - oracle.assemble(ORACLE_EXP_DIFF, m, x, prm, r["u"], 0.05)).max() < 1e-11 * np.abs(du).max()


def test_user_clone_of_registry_functor_is_bitwise_close(ctx):
    """A user-supplied copy of a registry functor (without traits) reproduces the built-in one."""
    fid = sb.register_user_functor(NONLIN_CLONE_SRC, "MyNonlinDiff", 1, 1)
    c = ex.baseline_case("2", B=8, Nx=300)
    tg, _ = sb.time_grid(c.tspan, c.tstep)
    tg = tg[:5]
    out = []
    for f in (sb.Functor(fid, c.params), sb.Functor(c.functor, c.params)):
        s = sb.pdepe(c.m, f, c.icfun, f, c.xmesh, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False, save="last")
        out.append(s.u[-1])
    assert np.abs(out[0] - out[1]).max() < 1e-12 * np.abs(out[1]).max()


def test_user_functor_compile_error_is_reported(ctx):
    fid = sb.register_user_functor("struct Broken { static constexpr int npde = 1, nparams = 1; };", "Broken", 1, 1)
    with pytest.raises(sb.SBError, match="NVRTC"):
        sb.DeviceProblem(ctx, 0, ex.linspace(0, 1, 50), sb.Functor(fid, np.ones((2, 1))))


def test_user_functor_long_mesh(ctx, oracle):
    fid = sb.register_user_functor(EXP_DIFF_SRC, "ExpDiff", 1, 3)
    Nx = 6000
    x = ex.linspace(0, 1, Nx)
    prm = np.array([[0.7, 1.5, 0.4]])
    ic = 0.3 * np.cos(x) ** 2
    fn = sb.Functor(fid, prm)
    sol = sb.pdepe(0, fn, ic, fn, x, (0.0, 0.03), ctx=ctx, tstep=1e-2, squeeze=False, save="last")
    r = oracle.pdepe_transient(ORACLE_EXP_DIFF, 0, x, prm, ic[None, :, None], sb.julia_range(0.0, 1e-2, 0.03), 1e-2)
    assert np.abs(sol.u[-1] - r["u"][:, :, 0]).max() < 1e-10 * np.abs(r["u"]).max()
