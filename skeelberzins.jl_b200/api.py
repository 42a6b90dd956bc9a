"""Host-side mirror of the reference interface for the internal implicit-Euler / Newton path.

Same names, argument meaning and error behaviour as SkeelBerzins.jl (citations are reference
file:line), batched over B independent PDE instances:

    pdepe(m, pdefun, icfun, bdfun, xmesh, tspan; params=Params(), kwargs...)   src/pdepe.jl:42
    pdepe(m, pdefun, icfun, bdfun, xmesh; ...)            (stationary)          src/pdepe.jl:156
    Params                                                                      src/utils.jl:191
    ProblemDefinition (alias SkeelBerzinsDiscretization)                        src/utils.jl:113
    assemble_(du, u, pb, t)                (assemble!)                          src/assembler.jl:19
    mass_matrix(pb)                                                             src/assembler.jl:138
    newton(...), newton_stat(...)                                               src/utils.jl:303,344

The host owns the time loop and the Newton loop exactly as the reference does; each Newton
iteration is one (fused) or two (split) kernel launches through the C ABI.  ``pdefun`` / ``bdfun``
are registry functors (``Functor``): Julia closures cannot run on the device, so the example
callbacks are device-compiled (include/skeelberzins_b200.h, enum sb_functor_id) and carry a
per-instance parameter matrix -- that is the parameter sweep.  ``icfun`` stays a host callable, as
in the reference (src/utils.jl:629-632).
"""
import ctypes
import weakref
from dataclasses import dataclass, field, replace
from typing import Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import SBError
from .timegrid import time_grid


class ConvergenceFailed(RuntimeError):
    """reference: throw("convergence failed")  (src/utils.jl:336,377)"""


@dataclass
class Params:
    """Keyword arguments of the solver, field for field (src/utils.jl:191-238)."""
    solver: str = "euler"                      # :euler | :DiffEq
    tstep: Union[float, Sequence[float]] = 1e-2
    hist: bool = False
    sparsity: str = "sparseArrays"             # :sparseArrays | :banded (storage plug; block-tridiagonal on device)
    linsolve: object = "KLUFactorization"      # accepted for source compatibility; the GPU solver is fixed
    maxit: int = 100
    tol: float = 1e-10
    data: bool = False
    nb_design_var: int = 0


@dataclass
class Functor:
    """Registry coefficient functor (pdefun + bdfun) with per-instance parameters [B, nparams]."""
    id: int
    params: np.ndarray

    def __post_init__(self):
        npde, nprm, name = functor_info(self.id)
        p = np.ascontiguousarray(self.params, dtype=np.float64)
        if p.ndim == 1:
            p = p.reshape(1, -1)
        if p.shape[1] != nprm:
            raise ValueError("functor %s takes %d parameters per instance, got %d" % (name, nprm, p.shape[1]))
        self.params = p
        self.npde, self.nparams, self.name = npde, nprm, name

    @property
    def batch(self):
        return self.params.shape[0]


def functor_info(fid):
    lib = _lib.load()
    npde, nprm, name = ctypes.c_int(), ctypes.c_int(), ctypes.c_char_p()
    _lib.check(lib.sb_functor_info(int(fid), ctypes.byref(npde), ctypes.byref(nprm), ctypes.byref(name)))
    return npde.value, nprm.value, name.value.decode()


def guard_violations():
    """(violations, arrays_checked) of guard mode (SB_GUARD=1): device arrays whose canary bytes were overwritten."""
    v, n = ctypes.c_int64(), ctypes.c_int64()
    _lib.check(_lib.load().sb_debug_guard_violations(ctypes.byref(v), ctypes.byref(n)))
    return v.value, n.value


def kernel_launches():
    """CUDA kernels launched by the library so far in this process."""
    n = ctypes.c_int64()
    _lib.check(_lib.load().sb_kernel_launches(ctypes.byref(n)))
    return n.value


def register_user_functor(cuda_source, struct_name, npde, nparams):
    """Plugin slot (sb_register_user_functor): CUDA C++ source of a functor struct

        struct <struct_name> {
            static constexpr int npde = ..., nparams = ...;
            static constexpr bool c_unit = ..., s_zero = ...;    // c == 1 identically / s == 0 identically
            static constexpr bool s_const = ...;                 // optional: s does not depend on u or dudx
            template <class S> static __device__ void pde(double x, double t, const S* u, const S* ux,
                                                          const double* prm, S* c, S* f, S* s);
            template <class S> static __device__ void bc(double xl, const S* ul, double xr, const S* ur, double t,
                                                         const double* prm, S* pl, double* ql, S* pr, double* qr);
        };

    i.e. the reference's pdefun / bdfun written generically in the scalar type S (double or dual number).  The
    kernels are compiled with NVRTC for sm_100a when a problem that uses the functor is created.  Returns the
    functor id to pass to ``Functor``."""
    fid = ctypes.c_int()
    _lib.check(_lib.load().sb_register_user_functor(None, cuda_source.encode(), struct_name.encode(), int(npde),
                                                    int(nparams), ctypes.byref(fid)))
    return fid.value


def user_functor_compile_check(functor_id, Nx, m=0):
    """NVRTC dry run of a registered user functor for a problem shape (no GPU needed); returns the CUBIN size."""
    n = ctypes.c_size_t()
    _lib.check(_lib.load().sb_user_functor_compile_check(int(functor_id), int(Nx), int(m), ctypes.byref(n)))
    return n.value


class _PinnedPool:
    """Page-locked host buffers for results.  A buffer goes back to the pool when the numpy array that wraps
    it (and every view of it) has been garbage-collected, so repeated solves re-use the same pinned pages."""

    def __init__(self, max_cached=4):
        self.free, self.max_cached = {}, max_cached

    def empty(self, shape):
        n = int(np.prod(shape))
        nbytes = max(n, 1) * 8
        lst = self.free.get(nbytes)
        if lst:
            ptr = lst.pop()
        else:
            p = ctypes.c_void_p()
            _lib.check(_lib.load().sb_host_alloc(nbytes, ctypes.byref(p)))
            ptr = p.value
        buf = (ctypes.c_double * max(n, 1)).from_address(ptr)
        weakref.finalize(buf, self._release, ptr, nbytes)
        return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)

    def _release(self, ptr, nbytes):
        lst = self.free.setdefault(nbytes, [])
        if sum(len(v) for v in self.free.values()) < self.max_cached:
            lst.append(ptr)
        else:
            try:
                _lib.load().sb_host_free(ctypes.c_void_p(ptr))
            except Exception:
                pass


_pinned = _PinnedPool()


def pinned_empty(shape):
    """float64 array in page-locked host memory (fast H2D/D2H); freed/re-used when garbage-collected."""
    return _pinned.empty(shape)


def _is_device_array(a):
    return hasattr(a, "__cuda_array_interface__")


def _device_ptr(a, n):
    """device pointer of a float64 device array with n elements (or of a raw integer address)"""
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    cai = a.__cuda_array_interface__
    if cai["typestr"] not in ("<f8", "=f8", "|f8"):
        raise TypeError("device arrays must be float64")
    if int(np.prod(cai["shape"])) != n:
        raise ValueError("expected a device array of %d values, got shape %r" % (n, tuple(cai["shape"])))
    if cai.get("strides") is not None:
        raise ValueError("device arrays must be contiguous")
    return ctypes.c_void_p(cai["data"][0])


class Context:
    """One per GPU (sb_ctx).  stream: a cudaStream_t handle (int) to run on, or None for an own stream."""

    def __init__(self, device=0, stream=None):
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.sb_ctx_create(int(device), ctypes.c_void_p(stream) if stream else None, ctypes.byref(h)))
        self._h, self.device = h, int(device)

    def synchronize(self):
        _lib.check(_lib.load().sb_ctx_synchronize(self._h))

    def close(self):
        if self._h:
            _lib.load().sb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


class DeviceProblem:
    """Thin object wrapper of sb_problem: the per-iteration entry points of the C ABI."""

    def __init__(self, ctx, m, xmesh, functor, design="resident"):
        self.lib = _lib.load()
        self.ctx = ctx
        x = np.ascontiguousarray(xmesh, dtype=np.float64)
        h = ctypes.c_void_p()
        _lib.check(self.lib.sb_problem_create(ctx._h, int(m), x.size, ctypes.c_int64(functor.batch), _lib.ptr(x),
                                              int(functor.id), _lib.ptr(functor.params), functor.nparams,
                                              ctypes.byref(h)))
        self._h = h
        self.B, self.Nx, self.npde = functor.batch, x.size, functor.npde
        self.nparams, self.functor_id, self.m = functor.nparams, int(functor.id), int(m)
        self._mesh_copy = x.copy()
        self._mesh_sample = _mesh_sample(self._mesh_copy).copy()
        self._mesh_key = _mesh_fingerprint(x, self._mesh_sample)
        self.set_design(design)

    def close(self):
        if self._h:
            self.lib.sb_problem_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_params(self, params):
        p = np.ascontiguousarray(params, dtype=np.float64)
        if p.shape != (self.B, self.nparams):
            raise ValueError("expected params of shape %r" % ((self.B, self.nparams),))
        _lib.check(self.lib.sb_problem_set_params(self._h, _lib.ptr(p)))

    def set_design(self, design):
        names = {"split": _lib.DESIGN_SPLIT, "fused": _lib.DESIGN_FUSED, "fused_tma": _lib.DESIGN_FUSED_TMA,
                 "resident": _lib.DESIGN_RESIDENT}
        d = names[design] if isinstance(design, str) else int(design)
        _lib.check(self.lib.sb_problem_set_design(self._h, d))
        self.design = {v: k for k, v in names.items()}[d]

    def layout(self):
        R, T, w = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        _lib.check(self.lib.sb_problem_layout(self._h, ctypes.byref(R), ctypes.byref(T), ctypes.byref(w)))
        return dict(R=R.value, T=T.value, warp=bool(w.value))

    def longmesh(self):
        """True when the mesh is solved by the long-mesh path (tiles; more nodes than one thread group holds)."""
        lay = self.layout()
        return lay["R"] * lay["T"] < self.Nx

    def info(self):
        npde, Nx, B, s = ctypes.c_int(), ctypes.c_int(), ctypes.c_int64(), ctypes.c_int()
        _lib.check(self.lib.sb_problem_info(self._h, ctypes.byref(npde), ctypes.byref(Nx), ctypes.byref(B), ctypes.byref(s)))
        return dict(npde=npde.value, Nx=Nx.value, B=B.value, singular=bool(s.value))

    def quadrature(self):
        xi, ze = np.empty(self.Nx - 1), np.empty(self.Nx - 1)
        _lib.check(self.lib.sb_problem_quadrature(self._h, _lib.ptr(xi), _lib.ptr(ze)))
        return xi, ze

    def _nodes(self, a, w=None):
        a = np.ascontiguousarray(a, dtype=np.float64)
        n = self.B * self.Nx * (self.npde if w is None else w)
        if a.size != n:
            raise ValueError("expected %d values ([B=%d][Nx=%d][%d]), got %d" % (n, self.B, self.Nx, n // (self.B * self.Nx), a.size))
        return a

    def upload(self, u):
        u = self._nodes(u)
        _lib.check(self.lib.sb_state_upload(self._h, _lib.ptr(u)))

    def download(self, out=None):
        if out is None:
            out = pinned_empty((self.B, self.Nx, self.npde))
        _lib.check(self.lib.sb_state_download(self._h, _lib.ptr(out)))
        return out

    def checkpoint(self):
        _lib.check(self.lib.sb_state_checkpoint(self._h))

    def rollback(self):
        _lib.check(self.lib.sb_state_rollback(self._h))

    def profile_begin(self):
        _lib.check(self.lib.sb_profile_begin(self._h))

    def profile_end(self):
        a, s = ctypes.c_double(), ctypes.c_double()
        n, z, ii = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(self.lib.sb_profile_end(self._h, ctypes.byref(a), ctypes.byref(s), ctypes.byref(n), ctypes.byref(z),
                                           ctypes.byref(ii)))
        return dict(ms_assemble=a.value, ms_solve=s.value, launches=n.value, noop_launches=z.value,
                    instance_iterations=ii.value)

    def mass_flags(self, t0, stationary=False):
        _lib.check(self.lib.sb_mass_flags(self._h, float(t0), int(bool(stationary))))

    def mass(self):
        out = np.empty((self.B, self.Nx, self.npde))
        _lib.check(self.lib.sb_mass_download(self._h, _lib.ptr(out)))
        return out

    def assemble(self, u, t):
        du = np.empty((self.B, self.Nx, self.npde))
        uu = None if u is None else self._nodes(u)
        _lib.check(self.lib.sb_assemble(self._h, _lib.ptr(uu), float(t), _lib.ptr(du)))
        return du

    def assemble_device(self, d_u, t, d_du):
        """assemble! on DEVICE memory (sb_assemble_device): d_u / d_du are device pointers (int) or objects with
        ``__cuda_array_interface__`` (torch.cuda tensors, cupy arrays) holding float64 [B, Nx, npde] in reference
        layout; d_u None: the resident state.  Asynchronous on the context's stream."""
        _lib.check(self.lib.sb_assemble_device(self._h, _device_ptr(d_u, self.B * self.Nx * self.npde), float(t),
                                               _device_ptr(d_du, self.B * self.Nx * self.npde)))

    def begin_step(self, t_next, inv_tau):
        _lib.check(self.lib.sb_begin_step(self._h, float(t_next), float(inv_tau)))

    def residual_jacobian(self):
        _lib.check(self.lib.sb_residual_jacobian(self._h))

    def solve_update_norm(self, tol):
        _lib.check(self.lib.sb_solve_update_norm(self._h, float(tol)))

    def newton_iteration(self, tol):
        _lib.check(self.lib.sb_newton_iteration(self._h, float(tol)))

    def poll_active(self, blocking=True):
        n = ctypes.c_int64()
        _lib.check(self.lib.sb_poll_active(self._h, int(bool(blocking)), ctypes.byref(n)))
        return n.value

    def wait_iteration(self, it):
        n = ctypes.c_int64()
        _lib.check(self.lib.sb_wait_iteration(self._h, int(it), ctypes.byref(n)))
        return n.value

    def newton_solve(self, tol, maxit):
        """C-side pipelined Newton loop; raises ConvergenceFailed like the reference's throw."""
        it = ctypes.c_int()
        rc = self.lib.sb_newton_solve(self._h, float(tol), int(maxit), ctypes.byref(it))
        if rc == _lib.SB_ERR_CONVERGENCE:
            raise ConvergenceFailed("convergence failed")
        _lib.check(rc)
        return it.value

    def solve_steps(self, tgrid, tau, tol, maxit):
        """C-side time loop over the grid (no snapshots in between); raises ConvergenceFailed like the reference."""
        tg = np.ascontiguousarray(tgrid, dtype=np.float64)
        done = ctypes.c_int()
        rc = self.lib.sb_solve_steps(self._h, tg.size - 1, _lib.ptr(tg), float(tau) if tau is not None else -1.0,
                                     float(tol), int(maxit), ctypes.byref(done))
        if rc == _lib.SB_ERR_CONVERGENCE:
            raise ConvergenceFailed("convergence failed")
        _lib.check(rc)
        return done.value

    def solve_steps_opt(self, tgrid, tau, tol, maxit, step_iters=False, snapshots=None, tgrid_per_instance=None):
        """sb_solve_steps_opt: the whole-solve kernel with its per-instance extras.  tgrid_per_instance [B, nsteps+1]:
        one time grid per instance; step_iters=True: returns the [B, nsteps] Newton iteration counts; snapshots: a
        float64 array [nsteps+1, B, Nx, npde] (pinned: ``pinned_empty``) that receives every state."""
        tgi = None
        if tgrid_per_instance is not None:
            tgi = np.ascontiguousarray(tgrid_per_instance, dtype=np.float64)
            if tgi.ndim != 2 or tgi.shape[0] != self.B:
                raise ValueError("tgrid_per_instance must be [B, nsteps+1]")
            nsteps, tg = tgi.shape[1] - 1, None
        else:
            tg = np.ascontiguousarray(tgrid, dtype=np.float64)
            nsteps = tg.size - 1
        its = np.empty((self.B, nsteps), dtype=np.int32) if step_iters else None
        if snapshots is not None and (snapshots.dtype != np.float64 or snapshots.size != (nsteps + 1) * self.B * self.Nx * self.npde
                                      or not snapshots.flags.c_contiguous):
            raise ValueError("snapshots must be a contiguous float64 array [nsteps+1, B, Nx, npde]")
        done = ctypes.c_int()
        rc = self.lib.sb_solve_steps_opt(self._h, nsteps, _lib.ptr(tg), _lib.ptr(tgi), float(tau) if tau is not None else -1.0,
                                         float(tol), int(maxit), _lib.ptr(its), _lib.ptr(snapshots), ctypes.byref(done))
        if rc == _lib.SB_ERR_CONVERGENCE:
            raise ConvergenceFailed("convergence failed")
        _lib.check(rc)
        return its

    def solve_steps_final(self, tgrid, tau, tol, maxit, out=None):
        """sb_solve_steps_final: the whole-solve kernel, u(t_end) delivered to host memory [B, Nx, npde] round by round
        of instances while later rounds compute (the download hidden behind the solve).  Returns the array."""
        tg = np.ascontiguousarray(tgrid, dtype=np.float64)
        if out is None:
            out = pinned_empty((self.B, self.Nx, self.npde))
        elif out.dtype != np.float64 or out.size != self.B * self.Nx * self.npde or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array [B, Nx, npde]")
        done = ctypes.c_int()
        rc = self.lib.sb_solve_steps_final(self._h, tg.size - 1, _lib.ptr(tg), float(tau) if tau is not None else -1.0,
                                           float(tol), int(maxit), _lib.ptr(out), ctypes.byref(done))
        if rc == _lib.SB_ERR_CONVERGENCE:
            raise ConvergenceFailed("convergence failed")
        _lib.check(rc)
        return out

    def solve_from_host(self, u0, t0, tgrid, tau, tol, maxit, out=None):
        """sb_solve_from_host: upload, mass flags at t0, the whole-solve kernel and the download of u(t_end) as one
        pipelined operation -- chunks of instances are uploaded while the kernel already iterates on those that have
        arrived, finished rounds are downloaded while later ones iterate.  u0 [B, Nx, npde] (pinned for real overlap:
        ``pinned_empty``); returns u(t_end) [B, Nx, npde]."""
        u0 = self._nodes(u0)
        tg = np.ascontiguousarray(tgrid, dtype=np.float64)
        if out is None:
            out = pinned_empty((self.B, self.Nx, self.npde))
        elif out.dtype != np.float64 or out.size != self.B * self.Nx * self.npde or not out.flags.c_contiguous:
            raise ValueError("out must be a contiguous float64 array [B, Nx, npde]")
        done = ctypes.c_int()
        rc = self.lib.sb_solve_from_host(self._h, _lib.ptr(u0), float(t0), tg.size - 1, _lib.ptr(tg),
                                         float(tau) if tau is not None else -1.0, float(tol), int(maxit), _lib.ptr(out),
                                         ctypes.byref(done))
        if rc == _lib.SB_ERR_CONVERGENCE:
            raise ConvergenceFailed("convergence failed")
        _lib.check(rc)
        return out

    def _get(self, fn, dtype):
        out = np.empty(self.B, dtype=dtype)
        _lib.check(fn(self._h, _lib.ptr(out)))
        return out

    def step_iters(self):
        return self._get(self.lib.sb_get_step_iters, np.int32)

    def total_iters(self):
        return self._get(self.lib.sb_get_total_iters, np.int64)

    def last_norm(self):
        return self._get(self.lib.sb_get_last_norm, np.float64)

    def status(self):
        return self._get(self.lib.sb_get_status, np.int32)

    def enable_history(self, cap):
        _lib.check(self.lib.sb_enable_history(self._h, int(cap)))
        self._hist_cap = int(cap)

    def history(self):
        out = np.empty((self.B, self._hist_cap))
        _lib.check(self.lib.sb_get_history(self._h, _lib.ptr(out), self._hist_cap))
        return out

    def total_active_iterations(self):
        n = ctypes.c_int64()
        _lib.check(self.lib.sb_total_active_iterations(self._h, ctypes.byref(n)))
        return n.value

    def pdeval(self, u, x_eval):
        """(u, dudx)[B, nq, npde] at the points x_eval; u: states [B, Nx, npde] in reference layout or None (resident state)."""
        xq = np.atleast_1d(np.ascontiguousarray(x_eval, dtype=np.float64))
        uu = None if u is None else self._nodes(u)
        uo = np.empty((self.B, xq.size, self.npde))
        dxo = np.empty_like(uo)
        rc = self.lib.sb_pdeval(self._h, _lib.ptr(uu), xq.size, _lib.ptr(xq), _lib.ptr(uo), _lib.ptr(dxo))
        if rc == _lib.SB_ERR_INVALID and "oustide" in self.lib.sb_last_error().decode():
            raise ValueError("Evaluation point oustide of the spatial discretization.")   # src/utils.jl:413 (sic)
        _lib.check(rc)
        return uo, dxo

    def debug_system(self):
        P = self.npde
        F = np.empty((self.B, self.Nx, P))
        Jl, Jd, Ju = (np.empty((self.B, self.Nx, P, P)) for _ in range(3))
        _lib.check(self.lib.sb_debug_download_system(self._h, _lib.ptr(F), _lib.ptr(Jl), _lib.ptr(Jd), _lib.ptr(Ju)))
        return F, Jl, Jd, Ju

    def debug_solve(self, Jl, Jd, Ju, F):
        P = self.npde
        Jl, Jd, Ju = (self._nodes(a, P * P) for a in (Jl, Jd, Ju))
        F = self._nodes(F)
        x = np.empty((self.B, self.Nx, P))
        _lib.check(self.lib.sb_debug_solve(self._h, _lib.ptr(Jl), _lib.ptr(Jd), _lib.ptr(Ju), _lib.ptr(F), _lib.ptr(x)))
        return x


# Pool of device problems that pdepe() created and no caller holds: a parameter study calls pdepe() again and
# again with the same mesh / batch / functor, and re-using the device allocations (cudaMalloc/cudaFree
# synchronise the device) is what keeps the end-to-end call close to the kernel time.
_pool = {}
_POOL_MAX = 4          # pooled device problems (a call with streams=n holds n of them)


_MESH_EXACT_MAX = 1 << 16     # meshes up to this size are compared node by node before a pooled problem is reused


def _mesh_sample(x):
    """Evenly strided sample of a long mesh (16k points plus 1024 at both ends): what pooled problems are matched on
    when an exact comparison would cost more than the solve (a 2^22-node mesh: 5 ms for two passes over it, the solve
    takes 0.6 ms).  A caller who changes a long mesh between the sampled points must not rely on the pool: pass
    `return_device=True` / hold the DeviceProblem, or call clear_pool()."""
    if x.size <= _MESH_EXACT_MAX:
        return x
    step = x.size // (1 << 14)
    return np.concatenate([x[:1024], x[::step], x[-1024:]])


def _mesh_fingerprint(x, sample=None):
    """Cheap key of a mesh for the problem pool; a pooled problem is only reused after a comparison with its own copy
    of the mesh (exact up to 65536 nodes, on the sample beyond)."""
    xs = _mesh_sample(x) if sample is None else sample
    return (x.size, float(x[0]), float(x[-1]), float(xs.sum()), float(np.dot(xs, xs)))


def _same_mesh(dev, x, sample):
    return dev._mesh_copy.size == x.size and np.array_equal(dev._mesh_sample, sample)


def _pool_key(ctx, m, x, functor, sample=None):
    return (id(ctx), int(m), int(functor.id), functor.batch, _mesh_fingerprint(x, sample))


def _acquire(ctx, m, x, functor, design):
    sample = _mesh_sample(x)                     # taken once per call: key and comparison both use it
    lst = _pool.get(_pool_key(ctx, m, x, functor, sample))
    if lst and _same_mesh(lst[-1], x, sample):
        dev = lst.pop()
        dev.set_params(functor.params)
        dev.set_design(design)
        return dev
    return DeviceProblem(ctx, m, x, functor, design)


def _release(dev):
    key = (id(dev.ctx), dev.m, dev.functor_id, dev.B, dev._mesh_key)
    lst = _pool.setdefault(key, [])
    if len(lst) < 1 and sum(len(v) for v in _pool.values()) < _POOL_MAX:
        if getattr(dev, "_hist_cap", 0):
            dev.enable_history(0)
        lst.append(dev)
    else:
        dev.close()


def clear_pool():
    for lst in _pool.values():
        for dev in lst:
            dev.close()
    _pool.clear()


@dataclass
class ProblemDefinition:
    """Problem record (src/utils.jl:113-182), same field names; `jac` lives on the device as
    block-tridiagonal scratch (Jl, Jd, Ju) and is rewritten every Newton iteration like pb.jac."""
    npde: int
    Nx: int
    xmesh: np.ndarray
    tspan: tuple
    singular: bool
    m: int
    nb_design_var: int
    inival: np.ndarray          # [B, Nx*npde] flattened, component fastest (src/utils.jl:632)
    pdefunction: Functor
    icfunction: object
    bdfunction: Functor
    device: DeviceProblem = field(repr=False, default=None)
    _xi_zeta: tuple = field(repr=False, default=None)

    @property
    def jac(self):
        return self.device.debug_system()[1:]

    def _quadrature(self):
        if self._xi_zeta is None:
            self._xi_zeta = self.device.quadrature()
        return self._xi_zeta

    @property
    def xi(self):       # ξ (src/utils.jl:683-718); fetched from the library on first use
        return self._quadrature()[0]

    @property
    def zeta(self):     # ζ
        return self._quadrature()[1]

    def __len__(self):  # Base.length(pb) = nb_design_var, src/utils.jl:665
        return self.nb_design_var


SkeelBerzinsDiscretization = ProblemDefinition  # name used by the north_star for this record


class DiffEqArray:
    """Result container of the transient solve (RecursiveArrayTools.DiffEqArray, src/pdepe.jl:110):
    `u[k]` is the state at `t[k]`: [B, Nx] for npde == 1, [B, npde, Nx] otherwise (squeezed to the
    reference's Vector[Nx] / Matrix(npde, Nx) when the batch axis was not requested).
    `sol(t)` interpolates linearly in time (src/utils.jl:597); `sol(x_eval, t, pb)` in time and space (:561)."""

    def __init__(self, u, t):
        self.u, self.t = u, np.asarray(t)

    def __call__(self, *args, val=True, deriv=True):
        if len(args) == 1:
            return interpolate_sol_time(self, args[0])
        x_eval, t, pb = args
        return interpolate_sol_space(self, x_eval, t, pb, val=val, deriv=deriv)

    def __len__(self):
        return len(self.u)

    def __getitem__(self, k):
        return self.u[k]


def _to_reference_layout(u_approx, pb):
    """states as pdepe returns them ([Nx] | [B, Nx] | [npde, Nx] | [B, npde, Nx] | flat) -> [B, Nx, npde]"""
    B, Nx, P = pb.device.B, pb.Nx, pb.npde
    a = np.asarray(u_approx, dtype=np.float64)
    if a.size != B * Nx * P:
        raise ValueError("u_approx has %d values, expected %d" % (a.size, B * Nx * P))
    if P == 1 or a.ndim <= 1 or a.shape[-1] == Nx * P:
        return np.ascontiguousarray(a.reshape(B, Nx, P))          # component fastest already (or scalar PDE)
    return np.ascontiguousarray(np.transpose(a.reshape(B, P, Nx), (0, 2, 1)))


def _shape_eval(a, pb, scalar_x):
    """[B, nq, npde] -> drop the axes the caller did not ask for (reference: scalars for a scalar x_eval)"""
    if pb.npde == 1:
        a = a[:, :, 0]
    if scalar_x:
        a = a[:, 0]
    return a[0] if pb.device.B == 1 else a


def pdeval(m, xmesh, u_approx, x_eval, pb):
    """pdeval(m, xmesh, u_approx, x_eval, pb) (src/utils.jl:396-438): (u, dudx) at x_eval, evaluated on the GPU with
    the interpolation of the discretisation.  Batched: one result row per instance."""
    if m != pb.m:
        raise ValueError("The symmetry argument is different from the one used to solve the PDE problem.")   # :398
    scalar_x = np.isscalar(x_eval)
    uo, dxo = pb.device.pdeval(_to_reference_layout(u_approx, pb), x_eval)
    return _shape_eval(uo, pb, scalar_x), _shape_eval(dxo, pb, scalar_x)


def interpolate_sol_time(u_approx, t):
    """interpolate_sol_time(u_approx, t) (src/utils.jl:575-595): linear interpolation between stored time steps."""
    ts = u_approx.t
    if t == ts[0] or abs(t - ts[0]) <= 1.0e-10 * abs(ts[1] - ts[0]):
        return u_approx.u[0]
    idx = int(np.searchsorted(ts, t, side="left")) + 1            # searchsortedfirst (1-based)
    if idx == 1 or idx > len(u_approx):
        raise ValueError("The chosen time is not within the interval.")          # :583
    dt = ts[idx - 1] - ts[idx - 2]
    x1 = (ts[idx - 1] - t) / dt
    x0 = (t - ts[idx - 2]) / dt
    return x1 * np.asarray(u_approx.u[idx - 2]) + x0 * np.asarray(u_approx.u[idx - 1])


def interpolate_sol_space(u_approx, x_eval, times, pb, val=True, deriv=True):
    """interpolate_sol_space(u_approx, x_eval, times, pb; val, deriv) (src/utils.jl:460-559)."""
    scalar_t = np.isscalar(times)
    us, dus = [], []
    for t in np.atleast_1d(times):
        u, du = pdeval(pb.m, pb.xmesh, interpolate_sol_time(u_approx, t), x_eval, pb)
        us.append(u)
        dus.append(du)
    if scalar_t:
        us, dus = us[0], dus[0]
    if val and not deriv:
        return us
    if deriv and not val:
        return dus
    return us, dus


class ODEFunction:
    """DifferentialEquations.jl-facing operator (src/diffEq.jl:18-27): f(du, u, p, t) = assemble!(du, u, pb, t) on the
    GPU, the diagonal mass matrix when the problem is a DAE, and the block-tridiagonal Jacobian prototype.
    Vectors are flat [B, Nx*npde] (component fastest), like pb.inival: numpy arrays (copied in and out) or device
    arrays (anything with ``__cuda_array_interface__``, e.g. torch.cuda tensors: sb_assemble_device, asynchronous on the
    context's stream, which should then be the integrator's stream -- Context(stream=...))."""

    def __init__(self, pb):
        self.pb = pb
        M, flag_dae = mass_matrix(pb)
        self.mass_matrix = M if flag_dae else None               # src/diffEq.jl:21-26
        self.jac_prototype = jac_prototype(pb.Nx, pb.npde)

    def __call__(self, du, u, p, t):
        if _is_device_array(du):                     # integrator state on the GPU: no host round trip, asynchronous
            self.pb.device.assemble_device(u, t, du)
            return du
        return assemble_(du, u, self.pb, t)


def jac_prototype(Nx, npde):
    """All-ones block-tridiagonal sparsity prototype (get_sparsity_pattern, src/utils.jl:725-759) as scipy CSC."""
    import scipy.sparse as sp
    blk = np.ones((npde, npde))
    return sp.kron(sp.diags([np.ones(Nx - 1), np.ones(Nx), np.ones(Nx - 1)], [-1, 0, 1]), blk, format="csc")


def ODEProblem(pb):
    """ODEProblem(pb) (src/diffEq.jl:40-43): (ODEFunction, u0, tspan) for an external ODE/DAE integrator."""
    return ODEFunction(pb), pb.inival.copy(), pb.tspan


def reshape_solution(sol_u, pb):
    """reshape(sol, pb) (src/diffEq.jl:54-57): flat states -> [..., npde, Nx]."""
    a = np.asarray(sol_u)
    return np.transpose(a.reshape(a.shape[:-1] + (pb.Nx, pb.npde)), tuple(range(a.ndim - 1)) + (a.ndim, a.ndim - 1))


def _merge_params(params, kwargs, stationary):
    # src/pdepe.jl:44-52 / :158-166 -- kwargs override the fields of params
    names = ("solver", "tstep", "hist", "sparsity", "linsolve", "maxit", "tol", "data", "nb_design_var")
    over = {k: kwargs.pop(k) for k in names if k in kwargs}
    p = replace(params, **over)
    if stationary and "tstep" not in over:
        p = replace(p, tstep=np.inf)                      # :159
    return p


def problem_init(m, xmesh, tspan, pdefun, icfun, bdfun, params, ctx=None, design="resident"):
    """problem_init (src/utils.jl:609-663) for the whole batch."""
    if not isinstance(pdefun, Functor):
        raise TypeError("pdefun must be a registry Functor (device-compiled); arbitrary host closures cannot run "
                        "on the GPU and there is no CPU fallback")
    if bdfun is not None and bdfun is not pdefun:
        if not (isinstance(bdfun, Functor) and bdfun.id == pdefun.id):
            raise TypeError("bdfun must be the same registry Functor as pdefun (the registry couples them)")
    if params.sparsity not in ("sparseArrays", "banded"):
        # src/utils.jl:641
        raise ValueError("Error: Invalid sparsity pattern selected. Please choose from the available options: "
                         ":sparseArrays, :banded")
    x = np.ascontiguousarray(xmesh, dtype=np.float64)
    ctx = ctx or default_context(0)
    dev = _acquire(ctx, m, x, pdefun, design)
    B, Nx, npde = pdefun.batch, x.size, pdefun.npde
    if callable(icfun):
        u0 = np.asarray(icfun(x), dtype=np.float64)
    else:
        u0 = np.asarray(icfun, dtype=np.float64)
    if u0.size == Nx * npde:
        u0 = np.broadcast_to(u0.reshape(1, Nx, npde), (B, Nx, npde))
    inival = np.ascontiguousarray(u0.reshape(B, Nx * npde))
    info = dev.info()
    pb = ProblemDefinition(npde, Nx, x, tuple(tspan), info["singular"], int(m), params.nb_design_var, inival,
                           pdefun, icfun, pdefun, dev)
    return Nx, npde, inival, pb


def mass_matrix(pb, t0=None):
    """mass_matrix(pb) (src/assembler.jl:138-174): (diag[B, Nx*npde], flag_DAE) from the initial value."""
    pb.device.upload(pb.inival)
    pb.device.mass_flags(pb.tspan[0] if t0 is None else t0)
    M = pb.device.mass().reshape(pb.inival.shape)
    return M, bool((M == 0).any())


def assemble_(du, u, pb, t):
    """assemble!(du, u, pb, t) (src/assembler.jl:19-125): in place on flat vectors [B, Nx*npde]; returns the
    reshaped du like the reference does (:124)."""
    out = pb.device.assemble(u, t)
    du = np.asarray(du)
    du.reshape(-1)[:] = out.reshape(-1)
    return du.reshape(pb.device.B, pb.Nx, pb.npde)


def _newton_loop(dev, tol, maxit, depth=2):
    """The loop of src/utils.jl:311-336 for the batch, host-owned.  Up to `depth` iterations are kept
    enqueued behind the one whose "still active" counter the host waits for, so the GPU never idles on
    the host round trip (converged instances are frozen on the device: surplus launches are no-ops and
    per-instance iteration counts stay exact)."""
    launched = done = 0
    while True:
        while launched < maxit and launched - done < depth:
            dev.newton_iteration(tol)
            launched += 1
        n_active = dev.wait_iteration(done)
        done += 1
        if n_active == 0:
            return launched
        if done >= maxit:
            raise ConvergenceFailed("convergence failed")      # src/utils.jl:336


def newton(pb, tau, timeStep, tol=1.0e-10, maxit=100, hist_flag=False, c_loop=False):
    """newton(b, tau, timeStep, pb, mass, cache, rhs; ...) (src/utils.jl:303-337) on the resident state:
    b = M .* u_n is formed on the device (src/pdepe.jl:94), unP1 = copy(b) (:309), then the Newton loop.
    Returns per-instance history [B, maxit] when hist_flag."""
    dev = pb.device
    dev.begin_step(timeStep, 0.0 if np.isinf(tau) else 1.0 / tau)
    if c_loop:
        dev.newton_solve(tol, maxit)
    else:
        _newton_loop(dev, tol, maxit)
    if hist_flag:
        return dev.history()
    return None


newton_stat = newton  # src/utils.jl:344-378: same loop, residual without mass matrix (flags all 1, inv_tau = 0)


_stream_ctx = {}


def _stream_contexts(ctx, n):
    """ctx plus n-1 cached contexts (= CUDA streams) on the same GPU."""
    lst = _stream_ctx.setdefault(id(ctx), [])
    while len(lst) < n - 1:
        lst.append(Context(ctx.device))
    return [ctx] + lst[:n - 1]


def _pdepe_streams(m, pdefun, icfun, bdfun, x, tspan, params, ctx, design, squeeze, streams):
    """pdepe(save="last") for a batch split into `streams` contiguous sub-batches.  Every sub-batch is an ordinary
    solve -- problem_init, mass flags, the time loop of src/pdepe.jl:90-108 with the Newton loop inside (in C,
    sb_solve_steps) -- on its own stream and host thread; instances are independent, so the results are those of
    the single-batch call."""
    import threading
    ctx = ctx or default_context(0)
    B, npde, Nx = pdefun.batch, pdefun.npde, x.size
    timeSteps, tau = time_grid(tspan, params.tstep)
    if callable(icfun):
        u0 = np.asarray(icfun(x), dtype=np.float64)
    else:
        u0 = np.asarray(icfun, dtype=np.float64)
    shared = u0.size == Nx * npde
    u0 = u0.reshape((1 if shared else B), Nx, npde)
    bounds = [int(round(i * B / streams)) for i in range(streams + 1)]
    out = pinned_empty((B, Nx, npde))
    devs, errors = [], []
    for i, c in enumerate(_stream_contexts(ctx, streams)):
        lo, hi = bounds[i], bounds[i + 1]
        fn = Functor(pdefun.id, pdefun.params[lo:hi])
        devs.append((_acquire(c, m, x, fn, design), lo, hi))

    def work(dev, lo, hi):
        # upload, solve and download of a sub-batch on its own stream: the copies of one sub-batch overlap the
        # other's kernels
        try:
            dev.upload(np.broadcast_to(u0, (hi - lo, Nx, npde)) if shared else u0[lo:hi])
            dev.mass_flags(timeSteps[0])
            if timeSteps.size > 1:
                dev.solve_steps(timeSteps, tau, params.tol, params.maxit)
            dev.download(out[lo:hi])
        except Exception as e:          # re-raised on the calling thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=d) for d in devs[1:]]
    for t in threads:
        t.start()
    work(*devs[0])
    for t in threads:
        t.join()
    for dev, _, _ in devs:
        _release(dev)
    if errors:
        raise errors[0]
    u = out[:, :, 0] if npde == 1 else np.transpose(out, (0, 2, 1))
    return DiffEqArray([u], timeSteps[-1:])


def _pdepe_devices(devices, m, pdefun, icfun, bdfun, x, tspan, kw):
    """pdepe for a batch sharded over the GPUs `devices` of this process: contiguous blocks of instances
    (sharding.shard_range), one host thread and one context per GPU, no exchange between the shards (the instances are
    independent); the results are put together in instance order."""
    import threading
    from .sharding import shard_indices
    B, Nx, npde = pdefun.batch, x.size, pdefun.npde
    u0 = None if callable(icfun) else np.asarray(icfun, dtype=np.float64)
    per_inst_ic = u0 is not None and u0.size == B * Nx * npde and B > 1
    tstep = kw.get("tstep", None)
    parts, errors = [None] * len(devices), []

    shards = [shard_indices(B, len(devices), r) for r in range(len(devices))]   # strided: every GPU gets the same mix

    def work(r, d):
        idx = shards[r]
        if idx.size == 0:
            return
        try:
            fn = Functor(pdefun.id, pdefun.params[idx])
            ic = u0.reshape(B, Nx, npde)[idx] if per_inst_ic else icfun
            kw_r = dict(kw)
            if tstep is not None and np.ndim(tstep) == 2:
                kw_r["tstep"] = np.asarray(tstep)[idx]
            parts[r] = pdepe(m, fn, ic, fn, x, tspan, ctx=default_context(int(d)), squeeze=False, **kw_r)
        except Exception as e:          # re-raised on the calling thread
            errors.append(e)

    threads = [threading.Thread(target=work, args=(r, d)) for r, d in enumerate(devices)]
    for t_ in threads:
        t_.start()
    for t_ in threads:
        t_.join()
    if errors:
        raise errors[0]
    live = [(shards[r], p) for r, p in enumerate(parts) if p is not None]

    def gather(pieces):                  # back into instance order
        first = np.asarray(pieces[0])
        out = np.empty((B,) + first.shape[1:], dtype=first.dtype)
        for (idx, _), a in zip(live, pieces):
            out[idx] = a
        return out

    if tspan is None:
        return gather([p for _, p in live])
    p0 = live[0][1]
    us = [gather([p.u[i] for _, p in live]) for i in range(len(p0.u))]
    ts = p0.t if np.ndim(p0.t) == 1 else gather([p.t for _, p in live])
    return DiffEqArray(us, ts)


def pdepe(m, pdefun, icfun, bdfun, xmesh, tspan=None, *, params=None, ctx=None, design="resident", save="all",
          squeeze=None, c_loop=False, return_device=False, alias_compat=False, streams=1, devices=None, **kwargs):
    """Batched pdepe.  With `tspan`: transient solve by implicit Euler (src/pdepe.jl:42-128); without:
    stationary solve by one Newton solve with tau = Inf, t = 1 (src/pdepe.jl:156-192).

    Extra keyword arguments of this implementation: ctx (Context / GPU), design ("resident" | "fused_tma" | "fused" |
    "split"), save ("all": every time step as the reference does | "last": only u(t_end), for very large batches),
    c_loop (drive the time loop and the Newton loop from C: with the resident design the whole solve is one kernel
    launch; default False = the loops are owned by this host code, one C call per Newton iteration, as in the reference),
    tstep may also be a 2-D array [B, nsteps+1]: one time grid per instance (implies the whole-solve kernel),
    devices (list of GPU ordinals: the batch is sharded over them by instance index, one host thread per GPU, no
    exchange between shards; returns the solution only),
    squeeze (drop the batch axis; default: when the functor has a single instance),
    streams (with save="last" and c_loop: split the batch into that many contiguous sub-batches, each with its own
    CUDA stream, time loop and Newton loop on its own host thread -- the sub-batches' launches interleave on the
    GPU, so one sub-batch's nearly converged iterations no longer leave SMs idle).
    Returns what the reference returns: sol | (sol, pb) | (sol, storage) | (sol, pb, storage); stationary:
    u | (u, hist) | (u, pb) | (u, pb, hist).  Unlike the reference stored states are the
    true states, not the in-place masked aliases; alias_compat=True reproduces the reference's stored values
    (every sol.u[k] but the last -- and, for npde > 1, the first -- multiplied by the mass flags, src/pdepe.jl:87-107).
    """
    stationary = tspan is None
    if devices is not None and len(devices) > 1:
        if return_device or kwargs.get("data") or kwargs.get("hist") or (params is not None and (params.data or params.hist)):
            raise ValueError("devices=[...] returns the solution only (no data / hist / return_device)")
        kw = dict(kwargs, design=design, save=save, c_loop=c_loop, alias_compat=alias_compat, streams=streams)
        if params is not None:
            kw["params"] = params
        sol = _pdepe_devices(list(devices), m, pdefun, icfun, bdfun, np.asarray(xmesh, dtype=np.float64), tspan, kw)
        return sol
    if devices is not None and len(devices) == 1 and ctx is None:
        ctx = default_context(int(devices[0]))
    params = _merge_params(params or Params(tstep=np.inf if stationary else 1e-2), kwargs, stationary)
    if kwargs:
        raise TypeError("unknown keyword arguments: %s" % sorted(kwargs))
    x = np.asarray(xmesh, dtype=np.float64)
    # src/pdepe.jl:55-59 / :168-169
    assert m in (0, 1, 2), "Parameter m invalid"
    assert m == 0 or (m > 0 and x[0] >= 0), "Non-conforming mesh"
    if stationary:
        assert params.solver == "euler", "Elliptic problems can only be solved using the Euler method"
        assert np.isscalar(params.tstep) and np.isinf(params.tstep), \
            "Time step should be set to Inf to obtain the stationary solution"
    else:
        assert not (np.isscalar(params.tstep) and np.isinf(params.tstep)), \
            "Adjust time step or try the alternative method for solving elliptic problems"

    if (streams > 1 and not stationary and save != "all" and c_loop and not params.hist and not params.data
            and not return_device and params.solver == "euler" and isinstance(pdefun, Functor) and pdefun.batch >= 2 * streams):
        return _pdepe_streams(m, pdefun, icfun, bdfun, x, tspan, params, ctx, design, squeeze, int(streams))

    Nx, npde, inival, pb = problem_init(m, x, (0, 1) if stationary else tspan, pdefun, icfun, bdfun, params, ctx, design)
    dev = pb.device
    B = dev.B
    if squeeze is None:
        squeeze = (B == 1)
    if params.solver == "DiffEq":
        return pb                                              # src/pdepe.jl:123-126
    if params.hist:
        dev.enable_history(params.maxit)

    def shape(u):  # [B,Nx,npde] -> reference shapes ([npde, Nx] per instance for systems: a transposed VIEW of the
        # downloaded buffer -- copying 67 MB into the other order costs 15 ms on cfg4, 6 % of the whole call)
        u = u[:, :, 0] if npde == 1 else np.transpose(u, (0, 2, 1))
        return u[0] if squeeze else u

    if stationary:
        dev.upload(inival)
        dev.mass_flags(1.0, stationary=True)
        hist = newton_stat(pb, np.inf, 1.0, params.tol, params.maxit, params.hist, c_loop)   # src/pdepe.jl:180
        u = dev.download()
        flat = u.reshape(B, Nx * npde)
        flat = flat[0] if squeeze else flat                    # reference returns the flat vector
        if params.hist:
            it = dev.step_iters()
            hist = [hist[k, :it[k]] for k in range(B)]
            hist = hist[0] if squeeze else hist
        out = [flat]
        if params.data:
            out.append(pb)
        if params.hist:
            out.append(hist)
        if return_device:
            out.append(dev)
        elif not params.data:
            pb.device = None
            _release(dev)
        return out[0] if len(out) == 1 else tuple(out)

    per_instance = np.ndim(params.tstep) == 2                   # one time grid per instance: tstep[B, nsteps+1]
    if per_instance:
        grids = np.ascontiguousarray(params.tstep, dtype=np.float64)
        if grids.shape[0] != B or grids.shape[1] < 2:
            raise ValueError("a 2-D tstep must hold one time grid per instance: [B, nsteps+1]")
        if params.hist:
            raise ValueError("hist=True needs a time grid shared by the batch")
        timeSteps, tau = grids, None
        t0 = float(tspan[0])
    else:
        timeSteps, tau = time_grid(tspan, params.tstep)        # src/pdepe.jl:82-85
        t0 = timeSteps[0]
    nsteps = timeSteps.shape[-1] - 1
    storage = []
    whole = (per_instance or (c_loop and not params.hist)) and nsteps >= 1 and dev.design == "resident" and not dev.longmesh()
    if per_instance and not whole:
        raise ValueError("per-instance time grids need the resident design and a mesh in the batched range")
    pipelined = whole and save != "all" and not per_instance
    if not pipelined:
        dev.upload(inival)
        dev.mass_flags(t0)                                     # src/pdepe.jl:78-79
    if whole:
        # the time loop of src/pdepe.jl:90-108 in one launch (sb_solve_steps_opt); with save="all" every state is
        # written by the kernel and streamed to pinned host memory round by round while later rounds compute
        if pipelined:
            # only u(t_end) is wanted: upload, mass flags, solve and download as one pipelined operation -- instances
            # are uploaded while earlier ones iterate, finished rounds travel back while later ones iterate
            results = [shape(dev.solve_from_host(inival, t0, timeSteps, tau, params.tol, params.maxit))]
        else:
            snaps = pinned_empty((nsteps + 1, B, Nx, npde)) if save == "all" else None
            dev.solve_steps_opt(None if per_instance else timeSteps, tau, params.tol, params.maxit, snapshots=snaps,
                                tgrid_per_instance=grids if per_instance else None)
            if save == "all":
                results = [shape(snaps[i]) for i in range(nsteps + 1)]
            else:
                results = [shape(dev.download())]
    else:
        results = [shape(inival.reshape(B, Nx, npde))] if save == "all" else []
        if save != "all" and not params.hist and c_loop and timeSteps.size > 1:
            dev.solve_steps(timeSteps, tau, params.tol, params.maxit)        # same loop, driven from C
            timeSteps_loop = timeSteps[:1]
        else:
            timeSteps_loop = timeSteps
        for i in range(timeSteps_loop.size - 1):                   # src/pdepe.jl:90
            time_val = timeSteps[i + 1]
            tstep_val = tau if tau is not None else timeSteps[i + 1] - timeSteps[i]
            h = newton(pb, tstep_val, time_val, params.tol, params.maxit, params.hist, c_loop)      # :94-96
            if params.hist:
                it = dev.step_iters()
                hs = [h[k, :it[k]] for k in range(B)]
                storage.append(hs[0] if squeeze else hs)
            if save == "all":
                results.append(shape(dev.download()))
        if save != "all":
            results.append(shape(dev.download()))
    if save != "all":
        sol = DiffEqArray(results, timeSteps[..., -1:])
    else:
        if alias_compat:
            Mm = shape(dev.mass())
            first = 1 if npde > 1 else 0
            results = [r if (k < first or k == len(results) - 1) else Mm * r for k, r in enumerate(results)]
        sol = DiffEqArray(results, timeSteps)
    out = [sol]
    if params.data:
        out.append(pb)
    if params.hist:
        out.append(storage)
    if return_device:
        out.append(dev)
    elif not params.data:
        pb.device = None
        _release(dev)
    return out[0] if len(out) == 1 else tuple(out)
