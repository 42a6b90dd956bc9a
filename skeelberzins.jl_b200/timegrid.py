"""Time grid of the transient driver.

The reference builds it as ``collect(tspan[1]:tstep:tspan[2])`` (src/pdepe.jl:84), a Julia float
range.  Julia lifts the three floats to rationals when they have exact small rational
approximations (Base ``(:)(start::T, step::T, stop::T)`` in twiceprecision.jl) and then every
element is the correctly rounded value of ``(start_n + i*step_n)/den`` -- e.g. exactly
``fl(i/100)`` for ``0:0.01:1`` -- and NOT the result of repeated addition.
This module restates that construction so the Python host produces the same grid bit for bit.
"""
from fractions import Fraction
import functools
import math

import numpy as np

_MAXINT_F32 = 16777216  # maxintfloat(Float32): bound used by Base.rat for Float64 inputs


def _rat(x):
    """Base.rat: continued-fraction rational approximation with small terms."""
    y = float(x)
    a, d = 1, 1
    b, c = 0, 0
    m = _MAXINT_F32
    while abs(y) <= m:
        f = int(math.trunc(y))
        y -= f
        a, c = f * a + c, a
        b, d = f * b + d, b
        if max(abs(a), abs(b)) > m:
            return c, d
        if b != 0 and a / b == x:
            break
        if y == 0:
            break
        y = 1.0 / y
    return a, b


def julia_range(start, step, stop):
    """``collect(start:step:stop)`` for Float64 arguments (a fresh array; the exact-rational construction is cached per
    argument triple: 65 us for 90 steps, 0.5 ms for 1000, on every pdepe() call of a parameter study otherwise)."""
    return _julia_range(float(start), float(step), float(stop)).copy()


@functools.lru_cache(maxsize=64)
def _julia_range(start, step, stop):
    if step == 0:
        raise ValueError("range step cannot be zero")
    sn, sd = _rat(step)
    if sd != 0 and sn / sd == step:
        an, ad = _rat(start)
        bn, bd = _rat(stop)
        if ad != 0 and bd != 0 and an / ad == start and bn / bd == stop:
            den = ad * sd // math.gcd(ad, sd)
            mflt = 2.0 ** 53
            if den != 0 and abs(start * den) <= mflt and abs(step * den) <= mflt:
                start_n = round(start * den)
                step_n = round(step * den)
                # len = div(den*stop_n - stop_d*start_n + step_n*stop_d, step_n*stop_d), truncating
                num = den * bn - bd * start_n + step_n * bd
                dd = step_n * bd
                q = abs(num) // abs(dd)
                if (num < 0) != (dd < 0):
                    q = -q
                length = max(0, q)
                last = start + (length - 1) * step
                nxt = start + length * step

                def between(a, x, b):
                    return a <= x <= b or b <= x <= a

                if between(start, last, stop + step / 2) and not between(start, nxt, stop):
                    return np.array([float(Fraction(start_n + i * step_n, den)) for i in range(length)])
    # fallback: take start and step literally (steprangelen_hp): start + i*step in extended precision
    lf = (stop - start) / step
    if lf < 0:
        length = 0
    elif lf == 0:
        length = 1
    else:
        length = int(round(lf)) + 1
        stop2 = start + (length - 1) * step
        length -= int(start < stop < stop2) + int(start > stop > stop2)
    fs, fstep = Fraction(start), Fraction(step)
    return np.array([float(fs + i * fstep) for i in range(length)])


def time_grid(tspan, tstep):
    """(timeSteps, tau) as in src/pdepe.jl:82-85: scalar tstep -> range grid and constant tau;
    vector tstep -> the vector is the grid, tau = None (meaning diff of the grid, :85,:92)."""
    if np.isscalar(tstep):
        return julia_range(tspan[0], tstep, tspan[1]), float(tstep)
    grid = np.ascontiguousarray(tstep, dtype=np.float64)
    return grid, None
