"""The reference's example problems (examples/Example10x_*.jl, Example201) as batched cases.

Every ``Case`` carries what ``pdepe`` needs: symmetry m, registry functor id, mesh, tspan/tstep,
the host-side initial condition ``icfun`` (the reference evaluates icfun on the host too,
src/utils.jl:629-632) and the per-instance functor parameters.  ``golden`` is the literal the
reference's own test compares ``sum(sol.u[end])`` (or ``sum(sol)``) against (file:line cited).

``baseline_case(cfg, B, Nx)`` builds the batched parameter sweeps of BASELINE.json / SURVEY 8(d).
"""
from dataclasses import dataclass, field
from typing import Callable, Optional

import numpy as np
from scipy.special import j0

# registry ids (include/skeelberzins_b200.h, enum sb_functor_id)
LIN_DIFF, NONLIN_DIFF, LIN_DIFF_CYL, POISSON, STAT_NONLIN, REACT_DIFF = 0, 1, 2, 3, 4, 5
LIN_DIFF_SPH, CONV_DIFF, LIN_REACT, GENERIC_SCALAR = 6, 7, 8, 9

FUNCTOR_NAMES = {
    LIN_DIFF: "LIN_DIFF", NONLIN_DIFF: "NONLIN_DIFF", LIN_DIFF_CYL: "LIN_DIFF_CYL", POISSON: "POISSON",
    STAT_NONLIN: "STAT_NONLIN", REACT_DIFF: "REACT_DIFF", LIN_DIFF_SPH: "LIN_DIFF_SPH",
    CONV_DIFF: "CONV_DIFF", LIN_REACT: "LIN_REACT", GENERIC_SCALAR: "GENERIC_SCALAR",
}
FUNCTOR_NPDE = {0: 1, 1: 1, 2: 1, 3: 1, 4: 1, 5: 2, 6: 1, 7: 1, 8: 1, 9: 1}
FUNCTOR_NPARAMS = {0: 1, 1: 1, 2: 3, 3: 3, 4: 3, 5: 3, 6: 1, 7: 3, 8: 2, 9: 19}

BESSEL_ZERO = 2.404825557695773  # examples/Example103_LinearDiffusionCylindrical.jl:44,50


def generic_params(c0=1.0, c1=0.0, d0=1.0, d1=0.0, s0=0.0, s1=0.0, s2=0.0,
                   al=0.0, gl0=0.0, gl1=0.0, gl2=0.0, ll=0.0, ql=1.0,
                   ar=0.0, gr0=0.0, gr1=0.0, gr2=0.0, lr=0.0, qr=1.0):
    """Parameter vector of SB_GENERIC_SCALAR (order fixed by include/skeelberzins_b200.h)."""
    return np.array([c0, c1, d0, d1, s0, s1, s2, al, gl0, gl1, gl2, ll, ql, ar, gr0, gr1, gr2, lr, qr],
                    dtype=np.float64)


def barenblatt(x, t, p):
    """examples/Example102_NonlinearDiffusion.jl:33-42"""
    x = np.asarray(x, dtype=np.float64)
    tx = t ** (-1.0 / (p + 1.0))
    xx = x * tx
    xx = xx * xx
    xx = 1 - xx * (p - 1) / (2.0 * p * (p + 1))
    xx = np.where(xx < 0.0, 0.0, xx)
    return tx * xx ** (1.0 / (p - 1.0))


@dataclass
class Case:
    name: str
    m: int
    functor: int
    xmesh: np.ndarray
    icfun: Callable                      # x[Nx] -> u0[Nx] or u0[Nx, npde]
    params: np.ndarray                   # [B, nparams]
    tspan: Optional[tuple] = None        # None -> stationary
    tstep: object = 1e-2
    golden: Optional[float] = None       # reference literal for instance `golden_instance`
    golden_src: str = ""
    golden_instance: int = 0
    extra: dict = field(default_factory=dict)

    @property
    def npde(self):
        return FUNCTOR_NPDE[self.functor]

    @property
    def B(self):
        return self.params.shape[0]

    def initial(self):
        """[B, Nx, npde] initial values (icfun.(xmesh), identical for all instances)."""
        u0 = np.asarray(self.icfun(self.xmesh), dtype=np.float64).reshape(self.xmesh.size, self.npde)
        return np.ascontiguousarray(np.broadcast_to(u0, (self.B,) + u0.shape))


def _rows(*vals):
    return np.array([vals], dtype=np.float64)


def linspace(a, b, n):
    """Julia `collect(range(a, b; length=n))` / LinRange: a + i*(b-a)/(n-1) evaluated as in Base
    (lerpi): ((n-1-i)*a + i*b)/(n-1)."""
    i = np.arange(n, dtype=np.float64)
    return ((n - 1 - i) * float(a) + i * float(b)) / (n - 1)


def example101(Nx=21, tstep=1e-2):
    """examples/Example101_LinearDiffusion.jl:41-57; golden :64"""
    return Case("Example101_LinearDiffusion", 0, LIN_DIFF, linspace(0, 1, Nx),
                lambda x: np.exp(-100 * (x - 0.25) ** 2), _rows(1.0), (0.0, 1.0), tstep,
                3.721004873950427, "examples/Example101_LinearDiffusion.jl:64")


def example102(Nx=21):
    """examples/Example102_NonlinearDiffusion.jl:18-77; golden :84"""
    return Case("Example102_NonlinearDiffusion", 0, NONLIN_DIFF, linspace(-1, 1, Nx),
                lambda x: barenblatt(x, 0.001, 2), _rows(2.0), (0.001, 0.01), 1e-4,
                46.66666666678757, "examples/Example102_NonlinearDiffusion.jl:84")


def example103(Nx=21):
    """examples/Example103_LinearDiffusionCylindrical.jl:21-68; golden :75"""
    n = BESSEL_ZERO
    return Case("Example103_LinearDiffusionCylindrical", 1, LIN_DIFF_CYL, linspace(0, 1, Nx),
                lambda x: j0(n * x), _rows(1.0, n, float(j0(n))), (0.0, 1.0), 1e-2,
                0.0463188424523652, "examples/Example103_LinearDiffusionCylindrical.jl:75")


def example104(Nx=21, c0=1.0):
    """examples/Example104_Poisson.jl:20-52; golden :58 (c0=0: test/test_solveStationary.jl:6-37)"""
    return Case("Example104_Poisson", 0, POISSON, linspace(0, 1, Nx),
                lambda x: np.full_like(x, 0.1), _rows(c0, 1.0, 0.1), None, np.inf,
                3.7624999999999997, "examples/Example104_Poisson.jl:58")


def example105(Nx=21):
    """examples/Example105_StationaryNonlinearDiffusion.jl:17-56; golden :61"""
    return Case("Example105_StationaryNonlinearDiffusion", 0, STAT_NONLIN, linspace(0, 1, Nx),
                lambda x: np.full_like(x, 0.1), _rows(2.0, 1.0, 0.1), None, np.inf,
                6.025575019008793, "examples/Example105_StationaryNonlinearDiffusion.jl:61")


def example106(Nx=21):
    """examples/Example106_SystemReactionDiffusion.jl:30-77; golden :84"""
    return Case("Example106_SystemReactionDiffusion", 0, REACT_DIFF, linspace(0, 1, Nx),
                lambda x: np.zeros((x.size, 2)), _rows(0.5, 0.1, 1.0), (0.0, 10.0), 1e-2,
                29.034702247833415, "examples/Example106_SystemReactionDiffusion.jl:84")


def example107(Nx=21):
    """examples/Example107_LinearDiffusionSpherical.jl:19-63; test :67-68 is an error-norm bound"""
    return Case("Example107_LinearDiffusionSpherical", 2, LIN_DIFF_SPH, linspace(0, 1, Nx),
                lambda x: x ** 2, _rows(1.0), (0.0, 0.8), 1e-2, None,
                "examples/Example107_LinearDiffusionSpherical.jl:67-68",
                extra={"exact": lambda x, t: x ** 2 + 6 * t, "err_bound": 1.0e-2})


def example201(n=50):
    """examples/Example201_PartialDerivativeApprox.jl:21-63 (the internal-Euler solve only)"""
    D, eta, L, K = 0.1, 10.0, 1.0, 1.0
    return Case("Example201_PartialDerivativeApprox", 0, CONV_DIFF, linspace(0, 1, n),
                lambda x: (K * L / D) * (1 - np.exp(-eta * (1 - x / L))) / eta, _rows(D, eta, L),
                (0.0, 1.0), 1.0 / (n - 1), 0.6621164947313215,
                "examples/Example201_PartialDerivativeApprox.jl:108",
                extra={"D": D, "eta": eta, "L": L, "K": K, "Ip": 1.0})


ALL_EXAMPLES = [example101, example102, example103, example104, example105, example106, example107,
                example201]


def baseline_case(cfg, B=None, Nx=None):
    """Batched workloads of BASELINE.json `configs` made concrete.

    cfg: "1", "2", "3a", "3b", "4", "5".  Instance k = B//2 carries the example's own parameter.
    """
    cfg = str(cfg)
    if cfg == "1":
        return example101(Nx or 21)
    if cfg == "2":
        B = B or 4096
        c = example102(Nx or 1024)
        k = np.arange(B, dtype=np.float64)
        c.params = (1.0 + 2.0 * k / B).reshape(B, 1)
        c.golden_instance = B // 2
        c.name = "cfg2_Example102_sweep"
        return c
    if cfg in ("3a", "3b"):
        B = B or 4096
        k = np.arange(B, dtype=np.float64)
        D = 0.5 + k / B
        if cfg == "3a":
            c = example103(Nx or 512)
            c.params = np.stack([D, np.full(B, BESSEL_ZERO), np.full(B, float(j0(BESSEL_ZERO)))], axis=1)
        else:
            c = example107(Nx or 512)
            c.params = D.reshape(B, 1)
        c.golden_instance = B // 2
        c.name = "cfg%s_%s_sweep" % (cfg, c.name)
        return c
    if cfg == "4":
        B = B or 16384
        c = example106(Nx or 256)
        k = np.arange(B, dtype=np.float64)
        c.params = np.stack([np.full(B, 0.5), np.full(B, 0.1), 0.5 + k / B], axis=1)
        c.golden_instance = B // 2
        c.name = "cfg4_Example106_sweep"
        return c
    if cfg == "5":
        c = example105(Nx or 2 ** 22)
        c.name = "cfg5_Example105_longmesh"
        return c
    raise ValueError("unknown BASELINE config %r" % (cfg,))
