"""Writes tests/golden/oracle_vectors.npz: small input/output vectors of the CPU oracle (oracle/sb_oracle.cpp).

The reference is Julia and cannot run in this image, so these are not outputs of the reference itself: they are
outputs of the restatement after it was pinned to the reference's literals (reference_goldens.json).  They freeze
(a) the final states behind those literals, node by node, and (b) residuals / block-tridiagonal Jacobians of seeded
random states for every registry functor and coordinate system, so that the CUDA path can be checked against
committed numbers as well as against the oracle built on the box.

    python tests/golden/make_oracle_vectors.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import skeelberzins_jl_b200 as sb                        # noqa: E402  (host-side helpers only: examples, time grid)
from skeelberzins_jl_b200 import examples as ex          # noqa: E402
from oracle import sb_oracle as oracle                   # noqa: E402

NX_KERNEL = 33


def kernel_case(fid, m, singular, seed):
    """Seeded inputs of the kernel-level vectors (shared with the tests)."""
    rng = np.random.default_rng(seed)
    npde, nprm = oracle.functor_info(fid)
    lo = 0.0 if (singular or m == 0) else 0.5
    if m == 0 and not singular:
        lo = -1.0
    x = np.linspace(lo, lo + 2.0, NX_KERNEL)
    x[1:-1] += (rng.random(NX_KERNEL - 2) - 0.5) * 0.4 * (x[1] - x[0])
    prm = 0.5 + rng.random((2, nprm))
    u = 0.5 + rng.random((2, NX_KERNEL, npde))
    b = 0.5 + rng.random((2, NX_KERNEL, npde))
    return x, prm, u, b


KERNEL_CASES = [(fid, m, s) for fid in range(10) for (m, s) in ((0, False), (1, True), (1, False), (2, True), (2, False))]


def main():
    out = {}
    oracle.build()
    for name, make in (("ex101", ex.example101), ("ex102", ex.example102), ("ex103", ex.example103),
                       ("ex106", ex.example106), ("ex107", ex.example107)):
        c = make()
        tg, tau = sb.time_grid(c.tspan, c.tstep)
        r = oracle.pdepe_transient(c.functor, c.m, c.xmesh, c.params, c.initial(), tg, tau)
        out[name + "_u"] = r["u"][0]
        out[name + "_iters"] = r["iters"][0]
    for name, c in (("ex104", ex.example104(c0=1.0)), ("ex104c0", ex.example104(c0=0.0)), ("ex105", ex.example105())):
        r = oracle.pdepe_stationary(c.functor, c.m, c.xmesh, c.params, c.initial())
        out[name + "_u"] = r["u"][0]
        out[name + "_iters"] = r["iters"]
    for i, (fid, m, s) in enumerate(KERNEL_CASES):
        x, prm, u, b = kernel_case(fid, m, s, 1000 + i)
        key = "k_f%d_m%d_s%d" % (fid, m, int(s))
        out[key + "_du"] = oracle.assemble(fid, m, x, prm, u, 0.3)
        mass = oracle.mass(fid, m, x, prm, u, 0.3)
        F, Jl, Jd, Ju = oracle.residual_jacobian(fid, m, x, prm, u, b, 0.01, mass, 0.3)
        out[key + "_F"] = F
        out[key + "_J"] = np.stack([Jl, Jd, Ju])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
