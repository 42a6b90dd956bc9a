"""Where the end-to-end time of a solve goes: the public pdepe() call taken apart, host wall clock around
synchronised phases (cfg5 by default; transient configs: the whole-solve kernel).  Steers tuning; bench.py produces
the reported numbers.

    python tools/e2e_breakdown.py [--config 5 | 2 | 3a | 3b | 4]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skeelberzins_jl_b200 as sb                                   # noqa: E402
from skeelberzins_jl_b200 import api, examples as ex                # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="5")
    a = ap.parse_args()
    c = ex.baseline_case(a.config)
    ctx = sb.Context(0)
    fn = sb.Functor(c.functor, c.params)
    u0 = api.pinned_empty((c.params.shape[0], c.xmesh.size, c.npde))
    u0[:] = c.initial()
    stationary = c.tspan is None

    def call():
        if stationary:
            return sb.pdepe(c.m, fn, u0, fn, c.xmesh, ctx=ctx, c_loop=True, squeeze=False)
        return sb.pdepe(c.m, fn, u0, fn, c.xmesh, c.tspan, ctx=ctx, tstep=c.tstep, save="last", c_loop=True, squeeze=False)

    t0 = time.perf_counter()
    call()
    ctx.synchronize()
    cold = (time.perf_counter() - t0) * 1e3          # first call: problem set-up (mesh weights, allocations) included
    for _ in range(3):
        call()
    T = {}

    def lap(name, t0):
        ctx.synchronize()
        T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()

    n = 5
    for _ in range(n):
        t = time.perf_counter()
        Nx, npde, inival, pb = api.problem_init(c.m, c.xmesh, (0, 1) if stationary else c.tspan, fn, u0, fn, api.Params(), ctx, "resident")
        t = lap("problem_init (pool lookup, mesh compare)", t)
        dev = pb.device
        dev.upload(inival)
        t = lap("upload (H2D + layout)", t)
        if stationary:
            dev.mass_flags(1.0, stationary=True)
            t = lap("mass_flags", t)
            api.newton_stat(pb, np.inf, 1.0, 1e-10, 100, False, True)
            t = lap("newton_stat (C loop)", t)
        else:
            tg, tau = sb.time_grid(c.tspan, c.tstep)
            t = lap("time_grid", t)
            dev.mass_flags(tg[0])
            t = lap("mass_flags", t)
            dev.solve_steps_opt(tg, tau, 1e-10, 100)
            t = lap("solve_steps_opt (one launch)", t)
        u = dev.download()
        t = lap("download (layout + D2H)", t)
        pb.device = None
        api._release(dev)
        del u
        t = lap("release", t)
    t0 = time.perf_counter()
    for _ in range(n):
        call()
    ctx.synchronize()
    whole = (time.perf_counter() - t0) * 1e3 / n
    for k, v in T.items():
        print("%-45s %8.3f ms" % (k, v / n))
    print("%-45s %8.3f ms" % ("sum", sum(T.values()) / n))
    print("%-45s %8.3f ms" % ("pdepe() as a whole", whole))
    print("%-45s %8.3f ms" % ("first pdepe() call (problem set-up included)", cold))


if __name__ == "__main__":
    main()
