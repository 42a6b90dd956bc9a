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


# ---- npde = 3 and 4: reference src/utils.jl:67-104, :725-759 are generic in npde; here wide systems come through the plugin
CHAIN3_SRC = r"""
struct Chain3 {   // A -> B (rate k1*A^2), B -> C (rate k2*B); prm = {D1, D2, D3, k1, k2}
    static constexpr int npde = 3, nparams = 5;
    static constexpr bool c_unit = true, s_zero = false;
    template <class S> static __device__ void pde(double, double, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); c[1] = S(1.0); c[2] = S(1.0);
        f[0] = p[0] * ux[0]; f[1] = p[1] * ux[1]; f[2] = p[2] * ux[2];
        const S r1 = p[3] * u[0] * u[0], r2 = p[4] * u[1];
        s[0] = -r1; s[1] = r1 - r2; s[2] = r2;
    }
    template <class S> static __device__ void bc(double, const S* ul, double, const S* ur, double, const double*,
                                                 S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0] - 1.0; ql[0] = 0.0; pl[1] = S(0.0); ql[1] = 1.0; pl[2] = S(0.0); ql[2] = 1.0;
        pr[0] = S(0.0); qr[0] = 1.0; pr[1] = S(0.0); qr[1] = 1.0; pr[2] = ur[2]; qr[2] = 0.0;
    }
};
"""
CHAIN4_SRC = r"""
struct Chain4 {   // A -> B -> C -> D; species 2 has capacity 1 + u2 and a flux that feels u1; prm = {D1..D4, k1, k2, k3}
    static constexpr int npde = 4, nparams = 7;
    static constexpr bool c_unit = false, s_zero = false;
    template <class S> static __device__ void pde(double, double, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = S(1.0); c[1] = 1.0 + u[1]; c[2] = S(1.0); c[3] = S(1.0);
        f[0] = p[0] * ux[0]; f[1] = p[1] * (1.0 + u[0]) * ux[1]; f[2] = p[2] * ux[2]; f[3] = p[3] * ux[3];
        const S r1 = p[4] * u[0], r2 = p[5] * u[1], r3 = p[6] * u[2];
        s[0] = -r1; s[1] = r1 - r2; s[2] = r2 - r3; s[3] = r3;
    }
    template <class S> static __device__ void bc(double, const S* ul, double, const S* ur, double t, const double*,
                                                 S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0] - (1.0 + 0.1 * t); ql[0] = 0.0;
        for (int j = 1; j < 4; ++j) { pl[j] = S(0.0); ql[j] = 1.0; }
        for (int j = 0; j < 3; ++j) { pr[j] = S(0.0); qr[j] = 1.0; }
        pr[3] = ur[3]; qr[3] = 0.0;
    }
};
"""
WIDE = {3: (CHAIN3_SRC, "Chain3", 11, [0.5, 0.2, 0.1, 2.0, 1.0]), 4: (CHAIN4_SRC, "Chain4", 12, [0.5, 0.2, 0.1, 0.05, 2.0, 1.0, 0.5])}


@pytest.mark.parametrize("npde,Nx,design", [(3, 100, "resident"), (3, 100, "fused_tma"), (3, 100, "split"), (3, 1000, "resident"),
                                            (3, 1000, "fused"), (4, 50, "resident"), (4, 50, "split"), (4, 500, "resident"),
                                            (4, 500, "fused_tma")])
def test_three_and_four_species_systems(ctx, oracle, npde, Nx, design):
    """Wide systems through the plugin slot, every design, warp and CTA decompositions, against the oracle: final
    state to 1e-10, the Newton iteration count of every instance in every step, residual and Jacobian blocks."""
    src, name, oracle_id, base = WIDE[npde]
    fid = sb.register_user_functor(src, name, npde, len(base))
    B = 5
    prm = np.tile(np.array(base), (B, 1)) * (1.0 + 0.1 * np.arange(B))[:, None]
    x = ex.linspace(0, 1, Nx)
    u0 = np.zeros((B, Nx, npde))
    u0[:, :, 0] = np.exp(-5 * x)
    tg = np.linspace(0.0, 0.3, 7)
    fn = sb.Functor(fid, prm)
    dev = sb.DeviceProblem(ctx, 0, x, fn, design=design)
    try:
        dev.upload(u0)
        dev.mass_flags(tg[0])
        M = dev.mass()
        assert np.array_equal(M, oracle.mass(oracle_id, 0, x, prm, u0, tg[0]))
        its = np.empty((B, tg.size - 1), dtype=np.int32)
        for i in range(tg.size - 1):
            dev.begin_step(tg[i + 1], 1.0 / (tg[i + 1] - tg[i]))
            if i % 2:
                dev.newton_solve(1e-10, 100)
            else:
                sb.api._newton_loop(dev, 1e-10, 100)
            its[:, i] = dev.step_iters()
        u = dev.download().copy()
        r = oracle.pdepe_transient(oracle_id, 0, x, prm, u0, tg, None)
        assert np.abs(u - r["u"]).max() < 1e-10 * np.abs(r["u"]).max()
        assert np.array_equal(its, r["iters"])
        assert (dev.status() == 0).all()
        if design == "split":                                # F and the block-tridiagonal Jacobian entry by entry
            dev.begin_step(0.4, 10.0)
            b = M * u
            dev.residual_jacobian()
            F, Jl, Jd, Ju = dev.debug_system()
            Fo, Jlo, Jdo, Juo = oracle.residual_jacobian(oracle_id, 0, x, prm, b, b, 0.1, M, 0.4)
            scale = np.abs(Jdo).max()
            assert np.abs(F - Fo).max() <= 1e-12 * max(np.abs(Fo).max(), 1.0)
            for A, Ao in ((Jl, Jlo), (Jd, Jdo), (Ju, Juo)):
                assert np.abs(A - Ao).max() <= 1e-12 * scale
        if design == "resident":                             # the whole solve in one launch, same answer
            dev.upload(u0)
            dev.mass_flags(tg[0])
            it1 = dev.solve_steps_opt(tg, None, 1e-10, 100, step_iters=True)
            # (bit-identical for npde <= 2, tests/test_gpu_resident.py; the generic Gauss-Jordan block inverse of wider
            # systems is contracted into FMAs differently in the two kernels: agreement to rounding)
            assert np.abs(dev.download() - u).max() < 1e-13 * np.abs(u).max() and np.array_equal(it1, its)
    finally:
        dev.close()


@pytest.mark.parametrize("npde,Nx,top_rows", [(3, 1500, None), (3, 3000, 256), (4, 700, None)])
def test_wide_systems_on_long_meshes(ctx, oracle, monkeypatch, npde, Nx, top_rows):
    """Meshes beyond the batched range of a wide system (npde = 3: 1024 nodes, npde = 4: 512) take the long-mesh path
    with P x P blocks (k6s_* kernels compiled through the plugin): host-owned and C Newton loops against the oracle;
    with the top level lowered to 256 rows the 750 chunk rows of the 3000-node mesh pass through a stored level."""
    if top_rows:
        monkeypatch.setenv("SB_LONG_TOP_ROWS", str(top_rows))
    src, name, oracle_id, base = WIDE[npde]
    fid = sb.register_user_functor(src, name, npde, len(base))
    B = 2
    prm = np.tile(np.array(base), (B, 1)) * (1.0 + 0.1 * np.arange(B))[:, None]
    x = ex.linspace(0, 1, Nx)
    u0 = np.zeros((B, Nx, npde))
    u0[:, :, 0] = np.exp(-5 * x)
    tg = np.linspace(0.0, 0.15, 4)
    dev = sb.DeviceProblem(ctx, 0, x, sb.Functor(fid, prm), design="fused")
    try:
        assert dev.longmesh()
        dev.upload(u0)
        dev.mass_flags(tg[0])
        assert np.array_equal(dev.mass(), oracle.mass(oracle_id, 0, x, prm, u0, tg[0]))
        its = np.empty((B, tg.size - 1), dtype=np.int32)
        for i in range(tg.size - 1):
            dev.begin_step(tg[i + 1], 1.0 / (tg[i + 1] - tg[i]))
            if i % 2:
                dev.newton_solve(1e-10, 100)
            else:
                sb.api._newton_loop(dev, 1e-10, 100)
            its[:, i] = dev.step_iters()
        u = dev.download().copy()
        r = oracle.pdepe_transient(oracle_id, 0, x, prm, u0, tg, None)
        assert np.abs(u - r["u"]).max() < 1e-10 * np.abs(r["u"]).max()
        assert np.array_equal(its, r["iters"])
    finally:
        dev.close()
