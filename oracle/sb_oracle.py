"""ctypes wrapper of the CPU oracle (oracle/sb_oracle.cpp).

TEST INFRASTRUCTURE ONLY -- importable from tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package never imports this module.

All arrays use the reference layout (src/utils.jl:632): [instance][node][component].
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsb_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    """Compile libsb_oracle.so with the system g++ (see oracle/Makefile)."""
    src = os.path.join(_HERE, "sb_oracle.cpp")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-shared",
           "-o", _LIB_PATH, src]
    subprocess.run(cmd, check=True, cwd=_HERE)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def _d(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _i(a):
    return None if a is None else a.ctypes.data_as(_ip)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def functor_info(fid):
    npde, nprm = ctypes.c_int(), ctypes.c_int()
    if lib().sbo_functor_info(int(fid), ctypes.byref(npde), ctypes.byref(nprm)) != 0:
        raise ValueError("unknown functor id %r" % (fid,))
    return npde.value, nprm.value


def num_threads():
    return lib().sbo_num_threads()


def _prep(fid, xmesh, prm):
    npde, nprm = functor_info(fid)
    x = _c(xmesh)
    prm = _c(prm).reshape(-1, nprm)
    return npde, nprm, x, prm, prm.shape[0]


def quadrature(m, xmesh):
    x = _c(xmesh)
    xi = np.empty(x.size - 1)
    ze = np.empty(x.size - 1)
    sing = ctypes.c_int()
    lib().sbo_quadrature(int(m), x.size, _d(x), _d(xi), _d(ze), ctypes.byref(sing))
    return xi, ze, bool(sing.value)


def assemble(fid, m, xmesh, prm, u, t):
    npde, nprm, x, prm, B = _prep(fid, xmesh, prm)
    u = _c(u).reshape(B, x.size, npde)
    du = np.empty_like(u)
    rc = lib().sbo_assemble(int(fid), int(m), x.size, _d(x), ctypes.c_int64(B), _d(prm), _d(u),
                            ctypes.c_double(t), _d(du))
    assert rc == 0
    return du


def mass(fid, m, xmesh, prm, u0, t0):
    npde, nprm, x, prm, B = _prep(fid, xmesh, prm)
    u0 = _c(u0).reshape(B, x.size, npde)
    M = np.empty_like(u0)
    rc = lib().sbo_mass(int(fid), int(m), x.size, _d(x), ctypes.c_int64(B), _d(prm), _d(u0),
                        ctypes.c_double(t0), _d(M))
    assert rc == 0
    return M


def residual_jacobian(fid, m, xmesh, prm, u, b, tau, mass_, t):
    """F[B,Nx,P], Jl/Jd/Ju[B,Nx,P,P] of one Newton iteration; mass_=None -> stationary residual."""
    npde, nprm, x, prm, B = _prep(fid, xmesh, prm)
    u = _c(u).reshape(B, x.size, npde)
    b = _c(b).reshape(B, x.size, npde)
    mm = None if mass_ is None else _c(mass_).reshape(B, x.size, npde)
    F = np.empty_like(u)
    Jl = np.empty((B, x.size, npde, npde))
    Jd = np.empty_like(Jl)
    Ju = np.empty_like(Jl)
    rc = lib().sbo_residual_jacobian(int(fid), int(m), x.size, _d(x), ctypes.c_int64(B), _d(prm),
                                     _d(u), _d(b), ctypes.c_double(tau), _d(mm), ctypes.c_double(t),
                                     _d(F), _d(Jl), _d(Jd), _d(Ju))
    assert rc == 0
    return F, Jl, Jd, Ju


def block_tridiag_solve(Jl, Jd, Ju, F):
    Jd = _c(Jd)
    B, Nx, P, _ = Jd.shape
    Jl = _c(Jl).reshape(Jd.shape)
    Ju = _c(Ju).reshape(Jd.shape)
    F = _c(F).reshape(B, Nx, P)
    x = np.empty_like(F)
    rc = lib().sbo_block_tridiag_solve(P, Nx, ctypes.c_int64(B), _d(Jl), _d(Jd), _d(Ju), _d(F), _d(x))
    assert rc == 0
    return x


def pdepe_transient(fid, m, xmesh, prm, u0, tgrid, tau=None, tol=1e-10, maxit=100, hist=False,
                    snapshots=False, nthreads=0, hist_cap=0):
    """Reference pdepe (src/pdepe.jl:42-128) per instance.  tau=None -> vector-tstep semantics
    (tau_i = diff(tgrid)); a float -> constant tstep.  Returns dict(u, iters, status[, hist, snapshots])."""
    npde, nprm, x, prm, B = _prep(fid, xmesh, prm)
    u = _c(u0).reshape(B, x.size, npde).copy()
    tg = _c(tgrid)
    nt = tg.size
    iters = np.zeros((B, max(nt - 1, 1)), dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    hist_cap = int(hist_cap) or int(maxit)        # norms kept per step (the first hist_cap iterations)
    H = np.full((B, max(nt - 1, 1), hist_cap), np.nan) if hist else None
    S = np.empty((B, nt, x.size, npde)) if snapshots else None
    rc = lib().sbo_pdepe_transient(int(fid), int(m), x.size, _d(x), ctypes.c_int64(B), _d(prm), _d(u), nt,
                                   _d(tg), ctypes.c_double(-1.0 if tau is None else tau),
                                   ctypes.c_double(tol), int(maxit), _i(iters), _d(H), hist_cap, _d(S),
                                   _i(status), int(nthreads))
    assert rc == 0
    out = dict(u=u, iters=iters[:, :nt - 1], status=status)
    if hist:
        out["hist"] = H
    if snapshots:
        out["snapshots"] = S
    return out


def pdepe_stationary(fid, m, xmesh, prm, u0, tol=1e-10, maxit=100, hist=False, nthreads=0):
    """Reference stationary pdepe (src/pdepe.jl:156-192) per instance."""
    npde, nprm, x, prm, B = _prep(fid, xmesh, prm)
    u = _c(u0).reshape(B, x.size, npde).copy()
    iters = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    H = np.full((B, maxit), np.nan) if hist else None
    rc = lib().sbo_pdepe_stationary(int(fid), int(m), x.size, _d(x), ctypes.c_int64(B), _d(prm), _d(u),
                                    ctypes.c_double(tol), int(maxit), _i(iters), _d(H), _i(status),
                                    int(nthreads))
    assert rc == 0
    out = dict(u=u, iters=iters, status=status)
    if hist:
        out["hist"] = H
    return out


def pdeval(fid, m, xmesh, u, x_eval):
    """pdeval (src/utils.jl:396-438) per instance: returns (u[B,nq,P], dudx[B,nq,P])."""
    npde, nprm = functor_info(fid)
    x = _c(xmesh)
    u = _c(u).reshape(-1, x.size, npde)
    xq = np.atleast_1d(_c(x_eval))
    uo = np.empty((u.shape[0], xq.size, npde))
    dxo = np.empty_like(uo)
    rc = lib().sbo_pdeval(int(fid), int(m), x.size, _d(x), ctypes.c_int64(u.shape[0]), _d(u), xq.size, _d(xq), _d(uo), _d(dxo))
    if rc == -3:
        raise ValueError("Evaluation point oustide of the spatial discretization.")
    assert rc == 0
    return uo, dxo


def interpolate_sol_time(ts, us, t):
    """interpolate_sol_time(u_approx, t), src/utils.jl:575-595 (us: sequence of snapshots, ts: their times)."""
    ts = np.asarray(ts, dtype=np.float64)
    if t == ts[0] or abs(t - ts[0]) <= 1.0e-10 * abs(ts[1] - ts[0]):
        return np.asarray(us[0])
    idx = int(np.searchsorted(ts, t, side="left")) + 1                  # searchsortedfirst, 1-based
    if idx == 1 or idx > len(us):
        raise ValueError("The chosen time is not within the interval.")
    dt = ts[idx - 1] - ts[idx - 2]
    x1 = (ts[idx - 1] - t) / dt
    x0 = (t - ts[idx - 2]) / dt
    return x1 * np.asarray(us[idx - 2]) + x0 * np.asarray(us[idx - 1])
