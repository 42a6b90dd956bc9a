"""skeelberzins.jl_b200 -- B200-native batched implicit-Euler / stationary-Newton path of SkeelBerzins.jl.

Host-side mirror of the reference interface for this path (``pdepe``, ``Params``,
``ProblemDefinition``/``SkeelBerzinsDiscretization``, ``assemble_``, ``newton``) over the C-ABI
library ``libskeelberzins_b200.so`` (include/skeelberzins_b200.h).  All compute runs in
hand-written sm_100a CUDA kernels; there is no CPU fallback -- importing the host logic works
anywhere, but every solve raises if the CUDA library or a GPU is missing.
"""
from . import examples, sharding, timegrid  # noqa: F401
from .api import (ODEFunction, ODEProblem, jac_prototype, reshape_solution,  # noqa: F401
                  ConvergenceFailed, Context, DeviceProblem, DiffEqArray, Functor, Params,  # noqa: F401
                  ProblemDefinition, SBError, SkeelBerzinsDiscretization, assemble_, default_context,
                  clear_pool, functor_info, guard_violations, interpolate_sol_space, interpolate_sol_time, kernel_launches, mass_matrix, pdeval, newton, newton_stat, pdepe, pinned_empty, problem_init, register_user_functor, user_functor_compile_check)
from .timegrid import julia_range, time_grid  # noqa: F401

__version__ = "0.1.0"
