"""Small driver for compute-sanitizer runs (memcheck / racecheck): exercises every kernel family once on tiny inputs.
Run on a GPU box:  compute-sanitizer --tool memcheck python tests/sanitizer_smoke.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skeelberzins_jl_b200 as sb                     # noqa: E402
from skeelberzins_jl_b200 import examples as ex       # noqa: E402

ctx = sb.Context(0)


def run(c, designs=("fused_tma", "fused", "split"), nsteps=2, **kw):
    for d in designs:
        fn = sb.Functor(c.functor, c.params)
        if c.tspan is None:
            u = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, ctx=ctx, design=d, squeeze=False, **kw)
            print(c.name, d, float(np.asarray(u).sum()))
        else:
            tg, _ = sb.time_grid(c.tspan, c.tstep)
            tg = tg[:nsteps + 1]
            s = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (tg[0], tg[-1]), ctx=ctx, design=d, tstep=tg, squeeze=False,
                         save="last", **kw)
            print(c.name, d, float(s.u[-1].sum()))


run(ex.baseline_case("2", B=5, Nx=21))            # warp per system, R = 1, padding lanes
run(ex.baseline_case("2", B=6, Nx=300))           # CTA of 64 threads
run(ex.baseline_case("2", B=3, Nx=1024))          # CTA of 128 threads, no padding
run(ex.baseline_case("3a", B=3, Nx=513))          # non-SYM0 weights, padding
run(ex.baseline_case("4", B=5, Nx=256))           # npde = 2, warp per system
run(ex.baseline_case("4", B=2, Nx=600))           # npde = 2, CTA per system
run(ex.example105(50))                            # stationary
os.environ["SB_FORCE_LONGMESH"] = "1"
run(ex.example105(5000), designs=("fused_tma",))  # long-mesh path, two levels
run(ex.baseline_case("3b", B=2, Nx=4500), designs=("fused_tma",), nsteps=2)
del os.environ["SB_FORCE_LONGMESH"]
c = ex.example201()
fn = sb.Functor(c.functor, c.params)
sol, pb = sb.pdepe(c.m, fn, c.icfun, fn, c.xmesh, (0.0, 0.1), ctx=ctx, data=True, tstep=c.tstep)
print("pdeval", sb.pdeval(pb.m, pb.xmesh, sol.u[-1], [0.0, 0.5, 1.0], pb))
print("sanitizer smoke done")
