"""Kernel timeline of one solve (CUPTI through torch.profiler): per-kernel durations and the gaps between
consecutive kernels in the warm stream -- what `ncu` cannot show because it serialises and flushes caches.

    python tools/trace_kernels.py --config 5 [--time-steps N] [--design fused_tma]
"""
import argparse
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skeelberzins_jl_b200 as sb                                   # noqa: E402
from skeelberzins_jl_b200 import examples as ex                     # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="5")
    ap.add_argument("--time-steps", type=int, default=3)
    ap.add_argument("--design", default="fused_tma")
    ap.add_argument("--last", type=int, default=40, help="print the last N kernel records")
    ap.add_argument("--c-loop", action="store_true", help="time loop and Newton loop in C (sb_solve_steps)")
    args = ap.parse_args()
    c = ex.baseline_case(args.config)
    ctx = sb.Context(0)
    fn = sb.Functor(c.functor, c.params)
    stationary = c.tspan is None

    def solve():
        if stationary:
            return sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, ctx=ctx, design=args.design, squeeze=False)
        tg, _ = sb.time_grid(c.tspan, c.tstep)
        tg = tg[:args.time_steps + 1]
        return sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), ctx=ctx, design=args.design, tstep=tg,
                        save="last", squeeze=False, c_loop=args.c_loop)

    solve()
    solve()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        solve()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    rows = []
    prev_end = None
    for e in evs:
        st, en = e.time_range.start, e.time_range.end
        rows.append((e.name[:70], en - st, (st - prev_end) if prev_end is not None else 0.0))
        prev_end = en
    for name, dur, gap in rows[-args.last:]:
        print("%-72s %9.2f us   gap before %7.2f us" % (name, dur, gap))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, dur, gap in rows:
        agg[name][0] += 1
        agg[name][1] += dur
    print("---- totals")
    for name, (n, tot) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s n=%5d  total %10.1f us  avg %8.2f us" % (name, n, tot, tot / n))
    print("sum of gaps: %.1f us over %d kernels" % (sum(r[2] for r in rows[1:]), len(rows)))


if __name__ == "__main__":
    main()
