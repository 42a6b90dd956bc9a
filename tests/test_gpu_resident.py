"""GPU tests of the whole-solve kernel (csrc/sb_resident.cuh) and what rides on it: per-step iteration counts,
snapshots of every state, per-instance time grids, the pivoted fallback solve, sharding over devices inside pdepe.

The bar is the north_star's: 1e-10 relative against the oracle and identical Newton iteration counts; against the
host-owned loops (one C call per Newton iteration) the states must be bit-identical, because every instance goes
through the same arithmetic whichever loop drives it."""
import numpy as np
import pytest

import skeelberzins_jl_b200 as sb
from skeelberzins_jl_b200 import examples as ex

pytestmark = pytest.mark.gpu
RTOL = 1e-10


@pytest.fixture(scope="module")
def ctx():
    return sb.Context(0)


def _relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("cfg,B,Nx,nsteps", [("2", 70, 1024, 6), ("2", 33, 1000, 4), ("2", 700, 100, 5), ("2", 9, 2048, 3),
                                             ("3a", 50, 512, 8), ("3b", 50, 500, 8), ("4", 40, 256, 12), ("4", 21, 300, 6),
                                             ("2", 5, 33, 4), ("4", 7, 1500, 3)])
def test_whole_solve_kernel_equals_host_loops(ctx, oracle, cfg, B, Nx, nsteps):
    c = ex.baseline_case(cfg, B=B, Nx=Nx)
    tg, tau = sb.time_grid(c.tspan, c.tstep)
    tg = tg[:nsteps + 1]
    tol, maxit = 1e-10, 100
    dev = sb.DeviceProblem(ctx, c.m, c.xmesh, sb.Functor(c.functor, c.params))
    try:
        # host-owned loops: per-step iteration counts from the device counters
        dev.upload(c.initial())
        dev.mass_flags(tg[0])
        it_host = np.empty((B, nsteps), dtype=np.int32)
        for i in range(nsteps):
            dev.begin_step(tg[i + 1], 1.0 / tau)
            sb.api._newton_loop(dev, tol, maxit)
            it_host[:, i] = dev.step_iters()
        u_host = dev.download().copy()
        # one launch
        dev.upload(c.initial())
        dev.mass_flags(tg[0])
        it_dev = dev.solve_steps_opt(tg, tau, tol, maxit, step_iters=True)
        u_dev = dev.download().copy()
        assert np.array_equal(u_dev, u_host)
        assert np.array_equal(it_dev, it_host)
        assert np.array_equal(dev.total_iters(), it_host.sum(axis=1))
        assert np.array_equal(dev.step_iters(), it_host[:, -1])
        assert (dev.status() == 0).all()
        # the Newton loop of single steps (sb_newton_solve) through the same kernel
        dev.upload(c.initial())
        dev.mass_flags(tg[0])
        for i in range(nsteps):
            dev.begin_step(tg[i + 1], 1.0 / tau)
            assert dev.newton_solve(tol, maxit) == it_host[:, i].max()
        assert np.array_equal(dev.download(), u_host)
    finally:
        dev.close()
    r = oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params, c.initial(), tg, tau)
    assert _relerr(u_dev, r["u"].reshape(u_dev.shape)) < RTOL
    assert np.array_equal(it_dev, r["iters"])


@pytest.mark.parametrize("cfg,B,Nx,nsteps", [("2", 1300, 512, 5), ("3a", 40, 512, 7), ("4", 30, 256, 9)])
def test_snapshots_of_every_state(ctx, oracle, cfg, B, Nx, nsteps):
    """save="all" (what the reference stores, src/pdepe.jl:99-107) through the whole-solve kernel: every state is written
    by the kernel and streamed out round by round; 1300 instances are several rounds of the persistent grid."""
    c = ex.baseline_case(cfg, B=B, Nx=Nx)
    tg, tau = sb.time_grid(c.tspan, c.tstep)
    tg = tg[:nsteps + 1]
    fn = sb.Functor(c.functor, c.params)
    sol = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False, c_loop=True)
    ref = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False)      # host loops
    assert len(sol.u) == nsteps + 1 and np.array_equal(sol.t, tg)
    for a, b in zip(sol.u, ref.u):
        assert np.array_equal(a, b)
    ks = np.unique(np.linspace(0, B - 1, 7).astype(int))
    r = oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params[ks], c.initial()[ks], tg, None, snapshots=True)
    for s in range(nsteps + 1):
        so = r["snapshots"][:, s]                         # [k, Nx, P]
        so = so[:, :, 0] if c.npde == 1 else np.transpose(so, (0, 2, 1))
        assert _relerr(np.asarray(sol.u[s])[ks], so) < RTOL
    # the reference's aliased storage (quirk Q1) on top of the same snapshots
    al = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False, c_loop=True, alias_compat=True)
    al_ref = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False, alias_compat=True)
    for a, b in zip(al.u, al_ref.u):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("cfg,B,Nx", [("2", 12, 256), ("3a", 10, 128), ("4", 9, 64)])
def test_time_grid_per_instance(ctx, oracle, cfg, B, Nx):
    """The reference's vector tstep (src/pdepe.jl:84-85,92), one vector per instance: instance k steps through its own
    non-uniform grid.  Oracle: the reference loop per instance with that instance's grid."""
    c = ex.baseline_case(cfg, B=B, Nx=Nx)
    rng = np.random.default_rng(11)
    nsteps = 7
    t0, t1 = c.tspan[0], c.tspan[0] + 8 * c.tstep
    grids = np.empty((B, nsteps + 1))
    for k in range(B):
        inner = np.sort(t0 + (t1 - t0) * rng.random(nsteps - 1))
        grids[k] = np.concatenate([[t0], inner, [t1]])
    fn = sb.Functor(c.functor, c.params)
    sol = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (t0, t1), ctx=ctx, tstep=grids, squeeze=False)
    assert len(sol.u) == nsteps + 1 and sol.t.shape == (B, nsteps + 1)
    for k in range(B):
        r = oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params[k:k + 1], c.initial()[k:k + 1], grids[k], None, snapshots=True)
        for s in (1, nsteps // 2, nsteps):
            so = r["snapshots"][0, s]
            so = so[:, 0] if c.npde == 1 else so.T
            assert _relerr(np.asarray(sol.u[s])[k], so) < RTOL, (k, s)
    last = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (t0, t1), ctx=ctx, tstep=grids, squeeze=False, save="last")
    assert np.array_equal(last.u[-1], sol.u[-1])


def _zero_diagonal_case(B, Nx):
    """c = 1, f = -a u_x with 1/tau = 2a/h^2: the interior diagonal of the Newton matrix is EXACTLY zero (the
    off-diagonals are a/h^2).  Nonsense as physics, a fair test of the linear algebra: elimination without row
    interchanges divides by zero at the first interior row, the reference's pivoted LU (src/utils.jl:317-319) solves it."""
    x = np.linspace(0.0, 1.0, Nx)
    h = x[1] - x[0]
    a = 1.0
    tau = h * h / (2 * a)
    prm = np.tile(ex.generic_params(c0=1.0, c1=0, d0=-a, d1=0, s0=0.3, s1=0, s2=0, al=1, gl0=0.1, gl1=0, gl2=0, ll=0, ql=0,
                                    ar=1, gr0=0.2, gr1=0, gr2=0, lr=0, qr=0), (B, 1))
    prm[:, 4] = 0.3 + 0.01 * np.arange(B)                  # a different source per instance
    u0 = np.tile((0.1 + 0.1 * x + 0.05 * np.sin(7 * x)).reshape(1, Nx, 1), (B, 1, 1))
    return x, tau, prm, u0


@pytest.mark.parametrize("design", ["resident", "fused_tma", "fused", "split"])
@pytest.mark.parametrize("Nx,tol", [(64, 1e-10), (200, 1e-10), (1024, 1e-7), (2048, 1e-6)])
def test_pivoted_fallback_on_a_zero_diagonal(ctx, oracle, design, Nx, tol):
    """Warp-per-system (Nx <= 256) and CTA-per-system decompositions.  The matrix has condition number ~ Nx and the
    solution grows with Nx, so the rounding floor of ||h||_2 reaches 1e-10 near Nx = 1000: the long meshes use a
    tolerance above it (the same for the oracle)."""
    B = 6
    x, tau, prm, u0 = _zero_diagonal_case(B, Nx)
    tg = np.array([0.0, tau, 2 * tau, 3 * tau])
    fn = sb.Functor(ex.GENERIC_SCALAR, prm)
    for c_loop in (False, True):
        sol, dev = sb.pdepe(0, fn, u0, fn, x, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False, save="last", design=design,
                            c_loop=c_loop, return_device=True, tol=tol)
        try:
            r = oracle.pdepe_transient(ex.GENERIC_SCALAR, 0, x, prm, u0, tg, None, tol=tol)
            assert (r["status"] == 0).all()
            assert _relerr(np.asarray(sol.u[-1]), r["u"][:, :, 0]) < 1e-9 * max(1.0, Nx / 200)   # cond(J) ~ Nx here
            assert np.array_equal(dev.total_iters(), r["iters"].sum(axis=1))
            assert (dev.status() == 3).all()               # solved, by the pivoted fallback
        finally:
            dev.close()


@pytest.mark.parametrize("P", [1, 2])
@pytest.mark.parametrize("Nx", [21, 100, 256, 1024, 2048])
def test_block_tridiagonal_solve_without_dominance(ctx, oracle, P, Nx):
    """The solver where it can fail: random block-tridiagonal systems whose off-diagonal blocks are as large as the
    diagonal ones (no dominance), half of the instances with exact zeros on the diagonal.  The monitor must hand the
    instances with pivot growth to the pivoted fallback; every instance must satisfy its equations."""
    rng = np.random.default_rng(Nx * 10 + P + 7)
    B = 8
    fid = ex.REACT_DIFF if P == 2 else ex.LIN_DIFF
    x = np.linspace(0, 1, Nx)
    prm = np.tile(np.array([0.5, 0.1, 1.0] if P == 2 else [1.0]), (B, 1))
    dev = sb.DeviceProblem(ctx, 0, x, sb.Functor(fid, prm), design="split")
    Jl = rng.standard_normal((B, Nx, P, P))
    Ju = rng.standard_normal((B, Nx, P, P))
    Jd = rng.standard_normal((B, Nx, P, P))
    Jl[:, 0] = 0
    Ju[:, -1] = 0
    for p in range(P):
        Jd[B // 2:, ::3, p, p] = 0.0
    F = rng.standard_normal((B, Nx, P))
    xg = dev.debug_solve(Jl, Jd, Ju, F)
    st = dev.status()
    dev.close()
    xo = oracle.block_tridiag_solve(Jl, Jd, Ju, F)
    # residual of every instance: r_i = Jl_i x_{i-1} + Jd_i x_i + Ju_i x_{i+1} - F_i
    res = np.einsum("bipq,biq->bip", Jd, xg) - F
    res[:, 1:] += np.einsum("bipq,biq->bip", Jl[:, 1:], xg[:, :-1])
    res[:, :-1] += np.einsum("bipq,biq->bip", Ju[:, :-1], xg[:, 1:])
    scale = np.abs(xg).reshape(B, -1).max(axis=1) * 4.0 + np.abs(F).reshape(B, -1).max(axis=1)
    assert (np.abs(res).reshape(B, -1).max(axis=1) / scale).max() < 1e-9
    # and against the oracle's pivoted LU where the system is not hopelessly conditioned
    err = np.abs(xg - xo).reshape(B, -1).max(axis=1) / np.abs(xo).reshape(B, -1).max(axis=1)
    assert np.median(err) < 1e-8
    if P == 1:
        assert (st[B // 2:] == 3).all()                    # zeros on the diagonal: cannot pass without interchanges
    # (P = 2: the pivoting INSIDE the 2x2 blocks already copes with a zero diagonal entry)


def test_convection_dominated_problem(ctx, oracle):
    """Cell Peclet number 25: the sub- and super-diagonal of the Newton matrix have opposite signs and the matrix is far
    from diagonally dominant (stationary limit, large time step) -- benign for elimination without interchanges or not,
    the result and the iteration counts must be the oracle's."""
    B, Nx = 10, 400
    x = np.linspace(0.0, 1.0, Nx)
    h = x[1] - x[0]
    d0 = 1e-3
    prm = np.tile(ex.generic_params(c0=1.0, c1=0, d0=d0, d1=0.0, s0=0.0, s1=0.0, s2=-50.0 * d0 / h * 0.5, al=1, gl0=1.0, gl1=0,
                                    gl2=0, ll=0, ql=0, ar=1, gr0=0.0, gr1=0, gr2=0, lr=0, qr=0), (B, 1))
    prm[:, 6] *= np.linspace(0.5, 1.5, B)
    u0 = np.tile((1.0 - x).reshape(1, Nx, 1), (B, 1, 1))
    tg = np.array([0.0, 10.0, 20.0, 1000.0])
    fn = sb.Functor(ex.GENERIC_SCALAR, prm)
    r = oracle.pdepe_transient(ex.GENERIC_SCALAR, 0, x, prm, u0, tg, None)
    assert (r["status"] == 0).all()
    for design in ("resident", "fused_tma", "split"):
        sol, dev = sb.pdepe(0, fn, u0, fn, x, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False, save="last", design=design,
                            c_loop=True, return_device=True)
        assert _relerr(np.asarray(sol.u[-1]), r["u"][:, :, 0]) < RTOL
        assert np.array_equal(dev.total_iters(), r["iters"].sum(axis=1))
        assert np.isin(dev.status(), (0, 3)).all()
        dev.close()


def test_convergence_failure_of_the_whole_solve_kernel(ctx):
    c = ex.baseline_case("2", B=20, Nx=256)
    tg, tau = sb.time_grid(c.tspan, c.tstep)
    dev = sb.DeviceProblem(ctx, c.m, c.xmesh, sb.Functor(c.functor, c.params))
    try:
        dev.upload(c.initial())
        dev.mass_flags(tg[0])
        with pytest.raises(sb.ConvergenceFailed, match="convergence failed"):      # src/utils.jl:336
            dev.solve_steps(tg[:6], tau, 1e-10, 3)
        assert (dev.status() == 1).any()
    finally:
        dev.close()


def test_devices_argument_shards_the_batch(ctx, oracle):
    """pdepe(..., devices=[...]): contiguous shards by instance index, one host thread per shard.  With a single GPU in
    the box the same device is named twice -- the sharding, threading and gathering are what is under test."""
    import torch
    n = torch.cuda.device_count()
    devices = [0, 1 % n, 2 % n]
    c = ex.baseline_case("2", B=50, Nx=256)
    tg, _ = sb.time_grid(c.tspan, c.tstep)
    tg = tg[:6]
    fn = sb.Functor(c.functor, c.params)
    one = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False, c_loop=True)
    many = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), tstep=tg, c_loop=True, devices=devices)
    assert len(one.u) == len(many.u)
    for a, b in zip(one.u, many.u):
        assert np.array_equal(a, b)
    last = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), tstep=tg, c_loop=True, devices=devices, save="last")
    assert np.array_equal(last.u[-1], one.u[-1])


@pytest.mark.parametrize("cfg,B,Nx,nsteps", [("2", 3000, 64, 4), ("4", 900, 100, 5), ("3a", 40, 512, 6), ("2", 700, 1024, 3), ("4", 2500, 256, 2)])
def test_final_state_streamed_out_during_the_solve(ctx, cfg, B, Nx, nsteps):
    """sb_solve_steps_final: u(t_end) arrives in host memory round by round of instances while later rounds compute
    (3000 instances of 64 nodes are several rounds of the persistent grid); bit-identical to a solve followed by a
    download, which is also what the device state holds afterwards, and what pdepe(save="last", c_loop=True) returns.
    sb_solve_from_host pipelines the upload as well (chunks of about 2 MB: 1, 1, 1 and 3 chunks here): same bits."""
    c = ex.baseline_case(cfg, B=B, Nx=Nx)
    tg, tau = sb.time_grid(c.tspan, c.tstep)
    tg = tg[:nsteps + 1]
    fn = sb.Functor(c.functor, c.params)
    dev = sb.DeviceProblem(ctx, c.m, c.xmesh, fn)
    try:
        dev.upload(c.initial())
        dev.mass_flags(tg[0])
        dev.solve_steps_opt(tg, None, 1e-10, 100)             # (vector-tstep semantics, as the pdepe call below)
        ref = dev.download().copy()
        dev.upload(c.initial())
        dev.mass_flags(tg[0])
        out = dev.solve_steps_final(tg, None, 1e-10, 100)
        assert np.array_equal(out, ref)
        assert np.array_equal(dev.download(), ref)
        for _ in range(2):                                    # (twice: the arrival words of the first call must not satisfy the second)
            u0p = sb.api.pinned_empty((B, Nx, c.npde))
            u0p[:] = c.initial()
            out2 = dev.solve_from_host(u0p, tg[0], tg, None, 1e-10, 100)
            assert np.array_equal(out2, ref)
            assert np.array_equal(dev.download(), ref)
        assert np.array_equal(dev.solve_from_host(np.array(c.initial()), tg[0], tg, None, 1e-10, 100), ref)   # pageable source
    finally:
        dev.close()
    sol = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), ctx=ctx, tstep=tg, squeeze=False, save="last", c_loop=True)
    got = np.asarray(sol.u[-1])
    want = ref[:, :, 0] if c.npde == 1 else np.transpose(ref, (0, 2, 1))
    assert np.array_equal(got, want)


def test_host_loops_after_a_whole_solve_on_the_same_problem(ctx):
    """A problem (as the pool hands it out again) that goes host-owned loops -> whole-solve kernel -> host-owned loops: the
    ring of completion counters of the per-iteration launches must move past the launches of the first phase when the
    whole-solve kernel clears the step's launch counter (it did not: uncleaned slots never reported)."""
    c = ex.baseline_case("2", B=40, Nx=512)
    tg, _ = sb.time_grid(c.tspan, c.tstep)
    tg = tg[:5]
    dev = sb.DeviceProblem(ctx, c.m, c.xmesh, sb.Functor(c.functor, c.params))

    def host_loops():
        dev.upload(c.initial())
        dev.mass_flags(tg[0])
        for i in range(tg.size - 1):
            dev.begin_step(tg[i + 1], 1.0 / (tg[i + 1] - tg[i]))
            sb.api._newton_loop(dev, 1e-10, 100)
        return dev.download().copy()

    try:
        a = host_loops()
        dev.upload(c.initial())
        dev.mass_flags(tg[0])
        dev.solve_steps_opt(tg, None, 1e-10, 100)
        assert np.array_equal(dev.download(), a)
        for _ in range(3):
            assert np.array_equal(host_loops(), a)
            assert np.array_equal(dev.solve_from_host(c.initial(), tg[0], tg, None, 1e-10, 100), a)
    finally:
        dev.close()
