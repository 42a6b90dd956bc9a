"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/skeelberzins_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("sb_build", os.path.join(ROOT, "skeelberzins.jl_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "skeelberzins_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in the header is not exported" % n


def test_python_binding_covers_header(built_lib):
    import skeelberzins_jl_b200 as sb
    from skeelberzins_jl_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    _lib.load()
    assert b"sm_100a" in _lib.load().sb_version()
    assert sb.functor_info(5) == (2, 3, "REACT_DIFF")
    for fid in range(10):
        npde, nprm, _ = sb.functor_info(fid)
        from skeelberzins_jl_b200 import examples as ex
        assert npde == ex.FUNCTOR_NPDE[fid] and nprm == ex.FUNCTOR_NPARAMS[fid]


def test_no_cpu_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import skeelberzins_jl_b200 as sb
    with pytest.raises(sb.SBError, match="no CPU fallback"):
        sb.Context(0)
    from skeelberzins_jl_b200 import examples as ex
    c = ex.example101()
    with pytest.raises(sb.SBError):
        sb.pdepe(c.m, sb.Functor(c.functor, c.params), c.icfun, None, c.xmesh, c.tspan)


def test_oracle_is_not_imported_by_product():
    pkg = os.path.join(ROOT, "skeelberzins.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "sb_oracle" not in src and "libsb_oracle" not in src, f


def test_time_grid_matches_julia_ranges():
    import numpy as np
    import skeelberzins_jl_b200 as sb
    g = sb.julia_range(0, 1e-2, 1)
    assert g.size == 101 and all(g[i] == i / 100 for i in range(101))
    g = sb.julia_range(0.001, 1e-4, 0.01)
    assert g.size == 91 and g[0] == 0.001 and g[-1] == 0.01 and g[45] == 0.0055
    g = sb.julia_range(0, 1 / 49, 1)          # Example201: tstep = 1/(n-1)
    assert g.size == 50 and all(g[i] == i / 49 for i in range(50))
    g = sb.julia_range(0.0, 1e-2, 0.8)
    assert g.size == 81 and g[-1] == 0.8
    grid, tau = sb.time_grid((0, 1), np.array([0.0, 0.1, 0.3, 1.0]))
    assert tau is None and grid.tolist() == [0.0, 0.1, 0.3, 1.0]


USER_SRC = r"""
struct CpuCheckFn {
    static constexpr int npde = 1, nparams = 2;
    static constexpr bool c_unit = false, s_zero = false;
    template <class S> static __device__ void pde(double x, double t, const S* u, const S* ux, const double* p, S* c, S* f, S* s) {
        c[0] = 1.0 + p[1] * u[0] * u[0]; f[0] = p[0] * sqrt(1.0 + u[0] * u[0]) * ux[0]; s[0] = log(2.0 + u[0] * u[0]) - x * t;
    }
    template <class S> static __device__ void bc(double, const S* ul, double, const S* ur, double t, const double* p,
                                                 S* pl, double* ql, S* pr, double* qr) {
        pl[0] = ul[0] - t; ql[0] = 0.0; pr[0] = ur[0] * ur[0] - 1.0; qr[0] = 1.0;
    }
};
"""


@pytest.mark.parametrize("Nx,m", [(21, 0), (256, 1), (1024, 0), (1024, 2), (100000, 0)])
def test_user_functor_plugin_compiles_without_gpu(built_lib, Nx, m):
    """The NVRTC stage of the plugin slot runs on the CPU box: the kernel headers must stay NVRTC-clean for every
    problem shape (warp / CTA / long-mesh decompositions, both symmetry specialisations)."""
    import skeelberzins_jl_b200 as sb
    fid = sb.register_user_functor(USER_SRC, "CpuCheckFn", 1, 2)
    assert sb.user_functor_compile_check(fid, Nx, m) > 10000
    bad = sb.register_user_functor("struct Nope { int x; };", "Nope", 1, 1)
    with pytest.raises(sb.SBError, match="NVRTC"):
        sb.user_functor_compile_check(bad, 64, 0)


def _julia_ccalls():
    """(symbol, return type, [argument types]) of every ccall in julia/SkeelBerzinsB200.jl."""
    src = open(os.path.join(ROOT, "julia", "SkeelBerzinsB200.jl")).read()
    out = []
    for m in re.finditer(r"ccall\(\(:(sb_[a-z_0-9]+),\s*libsb\),\s*(\w+),\s*\(", src):
        i, depth = m.end(), 1
        while depth:                                  # balanced scan over the argument-type tuple
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        tup = src[m.end():i - 1]
        args, cur, d = [], "", 0
        for ch in tup:
            if ch == "{":
                d += 1
            if ch == "}":
                d -= 1
            if ch == "," and d == 0:
                args.append(cur.strip())
                cur = ""
            else:
                cur += ch
        if cur.strip():
            args.append(cur.strip())
        out.append((m.group(1), m.group(2), args))
    return out


def test_julia_shim_matches_the_c_abi():
    """julia cannot run in this image, so the shim is checked statically: every ccall names an exported symbol and
    passes the number and kind of arguments the header (through the ctypes table) declares."""
    from skeelberzins_jl_b200 import _lib
    kind = {  # Julia ccall type -> class of C type
        "Cint": "int", "Cdouble": "double", "Int64": "int64", "Csize_t": "size", "Cstring": "ptr",
    }

    def jl_kind(t):
        return "ptr" if t.startswith(("Ptr{", "Ref{")) else kind[t]

    def c_kind(t):
        if t in (ctypes.c_int,):
            return "int"
        if t in (ctypes.c_double,):
            return "double"
        if t in (ctypes.c_int64,):
            return "int64"
        if t in (ctypes.c_size_t,):
            return "size"
        return "ptr"

    calls = _julia_ccalls()
    assert len(calls) >= 25
    for name, ret, args in calls:
        assert name in _lib.SIGNATURES, "%s is not part of the C ABI" % name
        restype, argtypes = _lib.SIGNATURES[name]
        assert (ret == "Cstring") == (restype is ctypes.c_char_p), name
        assert len(args) == len(argtypes), "%s: the shim passes %d arguments, the header declares %d" % (name, len(args), len(argtypes))
        for a, c in zip(args, argtypes):
            assert jl_kind(a) == c_kind(c), "%s: %s vs %s" % (name, a, c)
