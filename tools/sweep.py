"""One line of timing per (library build, BASELINE config, thread decomposition) -- the inner loop of kernel tuning.

    [SB_LIB_PATH=...] [SB_ROWS_PER_THREAD=R SB_THREADS_PER_SYSTEM=T] python tools/sweep.py --config 2 [--streams 2]
        [--batch B] [--time-steps N] [--reps 3] [--tag name]

No torch: the library's own stream, host wall clock around synchronised solves (each solve is 10+ ms of GPU
work) and the library's event-timed profile pass for the per-launch kernel time.  Numbers from here steer
tuning; bench.py produces the ones that are reported.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skeelberzins_jl_b200 as sb                                   # noqa: E402
from skeelberzins_jl_b200 import examples as ex                     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="2")
    ap.add_argument("--streams", type=int, default=0, help="0: 1 for the resident design, 2 otherwise")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--time-steps", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--tag", default="")
    ap.add_argument("--design", default="resident")
    ap.add_argument("--check", action="store_true", help="compare a sample of instances with the CPU oracle")
    a = ap.parse_args()
    c = ex.baseline_case(a.config, B=a.batch or None)
    B, Nx, P = c.params.shape[0], c.xmesh.size, c.npde
    stationary = c.tspan is None
    if stationary:
        tg, tau = np.array([1.0, 1.0]), np.inf
    else:
        tg, tau = sb.time_grid(c.tspan, c.tstep)
        if a.time_steps:
            tg = tg[:a.time_steps + 1]
    tol, maxit = 1e-10, 100
    n_str = 1 if stationary else (a.streams or (1 if a.design == "resident" else 2))
    ctxs = [sb.Context(0) for _ in range(n_str)]
    bounds = [int(round(i * B / n_str)) for i in range(n_str + 1)]
    u0 = c.initial()
    devs = []
    for i in range(n_str):
        d = sb.DeviceProblem(ctxs[i], c.m, c.xmesh, sb.Functor(c.functor, c.params[bounds[i]:bounds[i + 1]]), design=a.design)
        d.upload(u0[bounds[i]:bounds[i + 1]])
        d.checkpoint()
        devs.append(d)

    def one(d):
        if stationary:
            d.begin_step(1.0, 0.0)
            d.newton_solve(tol, maxit)
        else:
            d.solve_steps(tg, tau, tol, maxit)

    def solve():
        for d in devs:
            d.rollback()
            d.mass_flags(tg[0], stationary=stationary)
        th = [threading.Thread(target=one, args=(d,)) for d in devs[1:]]
        for t_ in th:
            t_.start()
        one(devs[0])
        for t_ in th:
            t_.join()
        for x in ctxs:
            x.synchronize()

    for _ in range(2):
        solve()
    units = sum(d.total_active_iterations() for d in devs) * Nx
    best = 1e30
    for _ in range(a.reps):
        t0 = time.perf_counter()
        solve()
        best = min(best, time.perf_counter() - t0)
    bad = int(sum((d.status() != 0).sum() for d in devs))
    lay = devs[0].layout()
    err = None
    if a.check:
        from oracle import sb_oracle
        ks = np.unique(np.linspace(0, B - 1, 9).astype(int))
        u = np.concatenate([d.download() for d in devs])[ks]
        if stationary:
            r = sb_oracle.pdepe_stationary(c.functor, c.m, c.xmesh, c.params[ks], u0[ks])
        else:
            r = sb_oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params[ks], u0[ks], tg, tau)
        uo = r["u"].reshape(u.shape)
        err = float((np.abs(u - uo).reshape(ks.size, -1).max(axis=1) / np.abs(uo).reshape(ks.size, -1).max(axis=1)).max())
        tot = np.concatenate([d.total_iters() for d in devs])[ks]
        if not (tot == np.asarray(r["iters"]).reshape(ks.size, -1).sum(axis=1)).all():
            err = -1.0   # iteration counts differ
    # kernel time per launch: whole batch in one problem, event pair around every launch
    for d in devs[1:]:
        d.close()
    if n_str > 1:
        devs[0].close()
        d = sb.DeviceProblem(ctxs[0], c.m, c.xmesh, sb.Functor(c.functor, c.params), design=a.design)
        d.upload(u0)
        d.checkpoint()
    else:
        d = devs[0]
    d.rollback()
    d.mass_flags(tg[0], stationary=stationary)
    d.profile_begin()
    if a.design == "resident" and not stationary:
        d.solve_steps(tg, tau, tol, maxit)
    else:
        for i in range(tg.size - 1):
            tau_i = tau if tau is not None else tg[i + 1] - tg[i]
            d.begin_step(tg[i + 1], 0.0 if np.isinf(tau_i) else 1.0 / tau_i)
            d.newton_solve(tol, maxit)
    prof = d.profile_end()
    nl = max(prof["launches"], 1)
    us = 1e3 * (prof["ms_solve"] + prof["ms_assemble"]) / nl
    upl = prof["instance_iterations"] * Nx / nl
    print(json.dumps(dict(tag=a.tag, cfg=a.config, R=lay["R"], T=lay["T"], streams=n_str, B=B, Nx=Nx,
                          ms_per_solve=round(best * 1e3, 3), units_per_s=float("%.4g" % (units / best)),
                          kernel_us_per_launch=round(us, 2), units_per_launch=round(upl),
                          kernel_GBps_at_24P=round(upl * 24 * P / (us * 1e-6) / 1e9, 1), not_ok=bad, oracle_err=err)))


if __name__ == "__main__":
    main()
