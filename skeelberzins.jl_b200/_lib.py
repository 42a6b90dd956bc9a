"""ctypes binding of libskeelberzins_b200.so (C ABI: include/skeelberzins_b200.h).

The library is the product; there is no fallback.  ``load()`` raises if it has not been built and
every compute call raises ``SBError`` with the library's message if no GPU is present.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SB_LIB_PATH", os.path.join(HERE, "libskeelberzins_b200.so"))  # SB_LIB_PATH: tuning builds

SB_OK = 0
SB_ERR_INVALID, SB_ERR_CUDA, SB_ERR_UNSUPPORTED, SB_ERR_CONVERGENCE, SB_ERR_PLUGIN = -1, -2, -3, -4, -5
DESIGN_SPLIT, DESIGN_FUSED, DESIGN_FUSED_TMA, DESIGN_RESIDENT = 0, 1, 2, 3

_c = ctypes
_dp = _c.POINTER(_c.c_double)
_vp = _c.c_void_p

# name -> (restype, argtypes); every symbol include/skeelberzins_b200.h declares
SIGNATURES = {
    "sb_version": (_c.c_char_p, []),
    "sb_last_error": (_c.c_char_p, []),
    "sb_build_digest": (_c.c_char_p, []),
    "sb_device_count": (_c.c_int, [_c.POINTER(_c.c_int)]),
    "sb_kernel_launches": (_c.c_int, [_c.POINTER(_c.c_int64)]),
    "sb_debug_guard_violations": (_c.c_int, [_c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64)]),
    "sb_host_alloc": (_c.c_int, [_c.c_size_t, _c.POINTER(_vp)]),
    "sb_host_free": (_c.c_int, [_vp]),
    "sb_ctx_create": (_c.c_int, [_c.c_int, _vp, _c.POINTER(_vp)]),
    "sb_ctx_destroy": (_c.c_int, [_vp]),
    "sb_ctx_synchronize": (_c.c_int, [_vp]),
    "sb_functor_info": (_c.c_int, [_c.c_int, _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.POINTER(_c.c_char_p)]),
    "sb_register_user_functor": (_c.c_int, [_vp, _c.c_char_p, _c.c_char_p, _c.c_int, _c.c_int, _c.POINTER(_c.c_int)]),
    "sb_user_functor_compile_check": (_c.c_int, [_c.c_int, _c.c_int, _c.c_int, _c.POINTER(_c.c_size_t)]),
    "sb_problem_create": (_c.c_int, [_vp, _c.c_int, _c.c_int, _c.c_int64, _vp, _c.c_int, _vp, _c.c_int, _c.POINTER(_vp)]),
    "sb_problem_destroy": (_c.c_int, [_vp]),
    "sb_problem_info": (_c.c_int, [_vp, _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int)]),
    "sb_problem_layout": (_c.c_int, [_vp, _c.POINTER(_c.c_int), _c.POINTER(_c.c_int), _c.POINTER(_c.c_int)]),
    "sb_problem_quadrature": (_c.c_int, [_vp, _vp, _vp]),
    "sb_problem_set_design": (_c.c_int, [_vp, _c.c_int]),
    "sb_problem_set_params": (_c.c_int, [_vp, _vp]),
    "sb_state_upload": (_c.c_int, [_vp, _vp]),
    "sb_state_download": (_c.c_int, [_vp, _vp]),
    "sb_state_checkpoint": (_c.c_int, [_vp]),
    "sb_state_rollback": (_c.c_int, [_vp]),
    "sb_profile_begin": (_c.c_int, [_vp]),
    "sb_profile_end": (_c.c_int, [_vp, _c.POINTER(_c.c_double), _c.POINTER(_c.c_double), _c.POINTER(_c.c_int64),
                                  _c.POINTER(_c.c_int64), _c.POINTER(_c.c_int64)]),
    "sb_mass_flags": (_c.c_int, [_vp, _c.c_double, _c.c_int]),
    "sb_mass_download": (_c.c_int, [_vp, _vp]),
    "sb_assemble": (_c.c_int, [_vp, _vp, _c.c_double, _vp]),
    "sb_assemble_device": (_c.c_int, [_vp, _vp, _c.c_double, _vp]),
    "sb_pdeval": (_c.c_int, [_vp, _vp, _c.c_int, _vp, _vp, _vp]),
    "sb_begin_step": (_c.c_int, [_vp, _c.c_double, _c.c_double]),
    "sb_residual_jacobian": (_c.c_int, [_vp]),
    "sb_solve_update_norm": (_c.c_int, [_vp, _c.c_double]),
    "sb_newton_iteration": (_c.c_int, [_vp, _c.c_double]),
    "sb_poll_active": (_c.c_int, [_vp, _c.c_int, _c.POINTER(_c.c_int64)]),
    "sb_wait_iteration": (_c.c_int, [_vp, _c.c_int, _c.POINTER(_c.c_int64)]),
    "sb_newton_solve": (_c.c_int, [_vp, _c.c_double, _c.c_int, _c.POINTER(_c.c_int)]),
    "sb_solve_steps": (_c.c_int, [_vp, _c.c_int, _vp, _c.c_double, _c.c_double, _c.c_int, _c.POINTER(_c.c_int)]),
    "sb_solve_steps_opt": (_c.c_int, [_vp, _c.c_int, _vp, _vp, _c.c_double, _c.c_double, _c.c_int, _vp, _vp, _c.POINTER(_c.c_int)]),
    "sb_solve_steps_final": (_c.c_int, [_vp, _c.c_int, _vp, _c.c_double, _c.c_double, _c.c_int, _vp, _c.POINTER(_c.c_int)]),
    "sb_solve_from_host": (_c.c_int, [_vp, _vp, _c.c_double, _c.c_int, _vp, _c.c_double, _c.c_double, _c.c_int, _vp, _c.POINTER(_c.c_int)]),
    "sb_get_step_iters": (_c.c_int, [_vp, _vp]),
    "sb_get_total_iters": (_c.c_int, [_vp, _vp]),
    "sb_get_last_norm": (_c.c_int, [_vp, _vp]),
    "sb_get_status": (_c.c_int, [_vp, _vp]),
    "sb_enable_history": (_c.c_int, [_vp, _c.c_int]),
    "sb_get_history": (_c.c_int, [_vp, _vp, _c.c_int]),
    "sb_total_active_iterations": (_c.c_int, [_vp, _c.POINTER(_c.c_int64)]),
    "sb_debug_download_system": (_c.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "sb_debug_solve": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None


class SBError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("skeelberzins_b200 error %d: %s" % (code, msg))
        self.code = code


def _build_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_sb_build", os.path.join(HERE, "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load():
    """Load the shared library (build it first with skeelberzins.jl_b200/build.py or __graft_entry__.build()).

    The library carries the digest of the sources it was built from; if csrc/ or include/ have changed since, this
    raises instead of silently running old kernels (SB_ALLOW_STALE=1 overrides; tuning builds named by SB_LIB_PATH are
    not checked)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            # not a fallback: the only implementation is the CUDA library; build it in-tree if nvcc is here
            try:
                _build_module().build()
            except Exception as exc:
                raise ImportError("%s is not built and building it failed (%s): run `python "
                                  "skeelberzins.jl_b200/build.py` (there is no CPU fallback)" % (LIB_PATH, exc))
        lib = _c.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if "SB_LIB_PATH" not in os.environ and not os.environ.get("SB_ALLOW_STALE"):
            have = lib.sb_build_digest().decode()
            want = _build_module().sources_digest()
            if have != want:
                raise ImportError("%s is stale: built from sources with digest %s..., the tree has %s... -- run `python "
                                  "skeelberzins.jl_b200/build.py` (SB_ALLOW_STALE=1 to load it anyway)"
                                  % (LIB_PATH, have[:12], want[:12]))
        _lib = lib
    return _lib


def check(rc):
    if rc != SB_OK:
        raise SBError(rc, load().sb_last_error().decode("utf-8", "replace"))


def ptr(a):
    return None if a is None else a.ctypes.data_as(_vp)
