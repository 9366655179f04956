"""ctypes binding of libroberts_b200.so (the C ABI of include/roberts_b200.h).

The library is built in-tree by ``python -m superfluid_dynamics_b200.build`` (or ``__graft_entry__.build()``).
There is no Python / NumPy / CPU fallback: if the shared library is missing the import fails loudly.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_bool, c_char_p, c_double, c_float, c_int, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libroberts_b200.so")


class rb_props(Structure):
    _fields_ = [("rho", c_double), ("U", c_double), ("kappa", c_double), ("depth", c_double),
                ("use_expansions", c_int), ("expansion_order", c_int), ("infinite_depth", c_int),
                ("physics", c_int), ("solve_mode", c_int), ("guess_mode", c_int), ("max_iterations", c_int),
                ("compute_energies", c_int), ("tolerance", c_double)]


class SimProperties(Structure):          # L/ExportTypes.cuh:8-20 ; P/integration/rhs.py:25-37
    _fields_ = [("L", c_double), ("rho", c_double), ("kappa", c_double), ("depth", c_double),
                ("use_expansions", c_bool), ("expansion_order", c_int), ("infinite_depth", c_bool)]


class rb_rk45_options(Structure):        # RK45_Options, L/RK45.cuh:21-27
    _fields_ = [("atol", c_double), ("rtol", c_double), ("h_min", c_double), ("h_max", c_double), ("initial_timestep", c_double)]


# AutonomousProblem<T,N>::run as a C callback: (user, state_dev, rhs_dev, cuda_stream)
RB_RHS_FN = ctypes.CFUNCTYPE(None, c_void_p, c_void_p, c_void_p, c_void_p)


class RK4SolverOptions(Structure):       # L/ExportTypes.cuh:35-40 ; P/integration/rhs.py:48-53
    _fields_ = [("timeStep", c_double), ("t0", c_double), ("t1", c_double), ("returnTrajectory", c_bool)]


class rb_opto(Structure):                # OptomechanicalVariables, L/OptomechanicalVariables.h:3-28 (+ the drive strength)
    _fields_ = [(n, c_double) for n in ("detuning", "gamma", "G", "Tau", "max_intensity", "initial_time", "location_x0_mode",
                                        "sigma_optical_mode", "Beta", "DampingStrength", "drive_strength")]


class rb_gl2_options(Structure):         # GaussLegendre2Options, L/GaussLegendre.cuh:70-90
    _fields_ = [("stepSize", c_double), ("newtonTolerance", c_double), ("maxNewtonIterations", c_size_t),
                ("allowSimplifiedFallback", c_int), ("returnTrajectory", c_int), ("armijo_c", c_double), ("backtrack", c_double),
                ("minAlpha", c_double), ("maxStepsHalves", c_size_t)]


class rb_gl2_stats(Structure):           # StepResult, L/GaussLegendre.cuh:176-181 (+ totals)
    _fields_ = [("numberIterations", c_size_t), ("converged", c_int), ("residualNorm", c_double), ("simplifiedFallbackUsed", c_int),
                ("steps_accepted", c_size_t), ("steps_halved", c_size_t), ("newton_iterations", c_size_t),
                ("rhs_evaluations", c_size_t), ("jacobians", c_size_t), ("linear_solves", c_size_t)]


class GaussLegendreOptions(Structure):   # L/ExportTypes.cuh:20-33 ; P/integration/rhs.py
    _fields_ = [("t0", c_double), ("t1", c_double), ("stepSize", c_double), ("newtonTolerance", c_double),
                ("maxNewtonIterations", c_size_t), ("allowSimplifiedFallback", c_bool), ("returnTrajectory", c_bool),
                ("armijo_c", c_double), ("backtrack", c_double), ("minAlpha", c_double), ("maxStepsHalves", c_size_t)]


class COptomechanicalVariables(Structure):   # L/ExportTypes.cuh:41-57
    _fields_ = [(n, c_double) for n in ("detuning", "gamma", "G", "tau", "max_intensity", "initial_time", "location_x0_mode",
                                        "sigma_optical_mode", "beta", "damping_strength")]


RB_WATER, RB_HELIUM, RB_HELIUM_INF = 0, 1, 2
RB_SOLVE_MATRIX_FREE, RB_SOLVE_DENSE_LU = 0, 1
RB_GUESS_COLD, RB_GUESS_WARM = 0, 1

# name -> (restype, argtypes); every symbol include/roberts_b200.h declares
_P = c_void_p
_D = POINTER(c_double)
SIGNATURES = {
    "rb_last_error": (c_char_p, []),
    "rb_version": (c_int, []),
    "rb_device_count": (c_int, []),
    "rb_set_device": (c_int, [c_int]),
    "rb_default_props": (None, [POINTER(rb_props)]),
    "rb_create": (_P, [c_int, c_int, POINTER(rb_props)]),
    "rb_destroy": (c_int, [_P]),
    "rb_set_stream": (c_int, [_P, _P]),
    "rb_get_stream": (_P, [_P]),
    "rb_get_props": (c_int, [_P, POINTER(rb_props), POINTER(c_int), POINTER(c_int)]),
    "rb_rhs": (c_int, [_P, _P, _P]),
    "rb_vorticities": (c_int, [_P, _P]),
    "rb_dev_a": (_P, [_P]),
    "rb_dev_zp": (_P, [_P]),
    "rb_dev_zpp": (_P, [_P]),
    "rb_dev_velocities_upper": (_P, [_P]),
    "rb_dev_phi_prime": (_P, [_P]),
    "rb_synchronize": (c_int, [_P]),
    "rb_energies": (c_int, [_P, _D]),
    "rb_solve_stats": (c_int, [_P, _D]),
    "rb_solve_status": (c_int, [_P, _D]),
    "rb_debug_set_row_range": (c_int, [_P, c_int, c_int]),
    "rb_sweep_plan": (c_int, [_P, POINTER(c_int)]),
    "rb_set_strict": (c_int, [_P, c_int]),
    "rb_zphi_derivative": (c_int, [_P, _P, _P, _P, _P, _P]),
    "rb_fft_derivative": (c_int, [_P, _P, _P, c_int, c_double]),
    "rb_create_M": (c_int, [_P, _P, _P, _P, c_double, c_int, c_size_t, _P]),
    "rb_create_finite_depth_M": (c_int, [_P, _P, _P, _P, c_double, c_int, c_size_t, c_int, _P]),
    "rb_velocity_matrices": (c_int, [_P, _P, _P, c_int, _P, _P, c_int, c_size_t, _P]),
    "rb_helium_velocity_matrices": (c_int, [_P, _P, _P, c_double, c_int, _P, _P, c_int, c_size_t, c_int, _P]),
    "rb_rhs_phi_water": (c_int, [_P, _P, _P, _P, c_double, c_int, _P]),
    "rb_rhs_phi_helium": (c_int, [_P, _P, _P, c_double, c_int, _P]),
    "rb_rhs_phi_helium_surface_tension": (c_int, [_P, _P, _P, _P, _P, c_double, c_double, c_int, _P]),
    "rb_rhs_phi_helium_expansion": (c_int, [_P, _P, _P, c_double, c_int, c_int, _P]),
    "rb_cotangent_sum": (c_int, [_P, _P, _P, _P]),
    "rb_rk4_create": (_P, [_P, c_double]),
    "rb_rk4_destroy": (c_int, [_P]),
    "rb_rk4_set_time_step": (c_int, [_P, c_double]),
    "rb_rk4_initialize": (c_int, [_P, _P, c_int]),
    "rb_rk4_step": (c_int, [_P]),
    "rb_rk4_evolve": (c_int, [_P, c_double, c_double, POINTER(c_size_t)]),
    "rb_rk4_run_steps": (c_int, [_P, c_size_t]),
    "rb_rk4_dev_state": (_P, [_P]),
    "rb_rk4_get_state": (c_int, [_P, _P]),
    "rb_rk4_current_time": (c_double, [_P]),
    "rb_rk4_stats": (c_int, [_P, _D]),
    "rb_rk4_chunk_stats": (c_int, [_P, _D]),
    "rb_rk4_guess_stats": (c_int, [_P, _D]),
    "rb_rk4_set_optimistic": (c_int, [_P, c_int]),
    "rb_rk4_set_guess": (c_int, [_P, c_int, c_int]),
    "rb_rk45_create": (_P, [_P, POINTER(rb_rk45_options)]),
    "rb_rk45_create_generic": (_P, [c_size_t, RB_RHS_FN, _P, POINTER(rb_rk45_options), _P]),
    "rb_rk45_destroy": (c_int, [_P]),
    "rb_rk45_set_options": (c_int, [_P, POINTER(rb_rk45_options)]),
    "rb_rk45_set_tolerance": (c_int, [_P, c_double, c_double]),
    "rb_rk45_set_max_rejected": (c_int, [_P, c_size_t]),
    "rb_rk45_initialize": (c_int, [_P, _P, c_int]),
    "rb_rk45_step": (c_int, [_P, POINTER(c_int)]),
    "rb_rk45_evolve": (c_int, [_P, c_double, c_double, POINTER(c_int)]),
    "rb_rk45_dev_state": (_P, [_P]),
    "rb_rk45_get_state": (c_int, [_P, _P]),
    "rb_rk45_current_time": (c_double, [_P]),
    "rb_rk45_current_timestep": (c_double, [_P]),
    "rb_rk45_stats": (c_int, [_P, _D]),
    "rb_rk4_set_logging": (c_int, [_P, c_size_t, c_size_t]),
    "rb_rk4_copy_trajectory": (c_int, [_P, POINTER(_D), POINTER(c_size_t), POINTER(_P), POINTER(c_size_t)]),
    "rb_free": (None, [_P]),
    "rb_rk4_stage_update": (c_int, [_P, _P, _P, c_double, c_size_t, _P]),
    "rb_rk4_final_update": (c_int, [_P, _P, _P, _P, _P, c_double, c_size_t, _P]),
    "rb_comm_handle_bytes": (c_int, []),
    "rb_comm_export": (c_int, [_P, c_char_p]),
    "rb_comm_init": (c_int, [_P, c_int, c_int, c_char_p]),
    "rb_comm_row_range": (c_int, [c_int, c_int, c_int, POINTER(c_int)]),
    "rb_comm_error": (c_int, [_P]),
    "rb_comm_destroy": (c_int, [_P]),
    "rb_launch_count": (ctypes.c_ulonglong, []),
    "rb_measure_fp64_peak": (c_int, [_D, _P]),
    "rb_measure_fp64_rate_3operand": (c_int, [_D, _P]),
    "rb_measure_fp64_tensor_overlap": (c_int, [_D, _P]),
    "rb_bench_sweep": (c_int, [_P, _P, c_int, POINTER(c_float), _D]),
    "calculateRHSFromVectors": (c_int, [_D, _D, _D, _D, _D, _D, c_double, c_double, c_double, c_double, c_size_t]),
    "calculateRHS256FromVectors": (c_int, [_D, _D, _D, _D, _D, _D, c_double, c_double, c_double, c_double]),
    "calculateRHS2048FromVectors": (c_int, [_D, _D, _D, _D, _D, _D, c_double, c_double, c_double, c_double]),
    "calculateRHS256FromVectorsBatched": (c_int, [_D, _D, _D, _D, _D, _D, c_double, c_double, c_double, c_double, c_int]),
    "calculateVorticities256FromVectors": (c_int, [_P, _P, _D, _P, _P, c_double, c_double, c_double, c_double]),
    "calculateDerivativeFFT256": (c_int, [_P, _P]),
    "integrateSimulationRK4": (c_int, [_D, POINTER(_D), POINTER(c_size_t), POINTER(_D), POINTER(c_size_t),
                                       POINTER(SimProperties), POINTER(RK4SolverOptions), c_size_t]),
    "integrateSimulationRK4_freeMemory": (c_int, [_D, _D]),
    "rb_integrate_rk4_host": (c_int, [_D, _D, c_size_t, c_size_t, POINTER(rb_props), c_double, c_size_t]),
    "rb_default_opto": (None, [POINTER(rb_opto)]),
    "rb_opto_drive_strength": (c_double, [POINTER(rb_opto), c_double, c_double, c_double]),
    "rb_light_intensity": (c_int, [_P, _P, POINTER(rb_opto), c_size_t, _P]),
    "rb_augmented_rhs": (c_int, [_P, POINTER(rb_opto), _P, _P]),
    "rb_aug_rk4_create": (_P, [_P, POINTER(rb_opto), c_double]),
    "rb_aug_rk4_destroy": (c_int, [_P]),
    "rb_aug_rk4_set_time_step": (c_int, [_P, c_double]),
    "rb_aug_rk4_initialize": (c_int, [_P, _P, c_int]),
    "rb_aug_rk4_step": (c_int, [_P]),
    "rb_aug_rk4_run_steps": (c_int, [_P, c_size_t]),
    "rb_aug_rk4_evolve": (c_int, [_P, c_double, c_double, POINTER(c_size_t)]),
    "rb_aug_rk4_dev_state": (_P, [_P]),
    "rb_aug_rk4_get_state": (c_int, [_P, _P]),
    "rb_aug_rk4_current_time": (c_double, [_P]),
    "rb_timed_rk4_create": (_P, [_P, POINTER(rb_opto), c_double]),
    "rb_timed_rk4_destroy": (c_int, [_P]),
    "rb_timed_rk4_set_time_step": (c_int, [_P, c_double]),
    "rb_timed_rk4_initialize": (c_int, [_P, _P, c_int]),
    "rb_timed_rk4_set_starting_time": (c_int, [_P, c_double]),
    "rb_timed_rhs": (c_int, [_P, c_double, c_int, _P, _P]),
    "rb_timed_rk4_step": (c_int, [_P, c_int]),
    "rb_timed_rk4_evolve": (c_int, [_P, c_double, c_double, POINTER(c_size_t)]),
    "rb_timed_rk4_set_logging": (c_int, [_P, c_int]),
    "rb_timed_rk4_copy_trajectory": (c_int, [_P, POINTER(_D), POINTER(c_size_t), POINTER(_P), POINTER(c_size_t)]),
    "rb_timed_rk4_dev_state": (_P, [_P]),
    "rb_timed_rk4_dev_delayed_intensity": (_P, [_P]),
    "rb_timed_rk4_get_state": (c_int, [_P, _P]),
    "rb_timed_rk4_current_time": (c_double, [_P]),
    "integrateOptomechanicalSimulationRK4": (c_int, [_D, POINTER(_D), POINTER(c_size_t), POINTER(_D), POINTER(c_size_t),
                                                     POINTER(SimProperties), POINTER(RK4SolverOptions),
                                                     POINTER(COptomechanicalVariables), c_size_t]),
    "integrateOptomechanicalSimulationRK4_freeMemory": (c_int, [_D, _D]),
    "calculateRhsAugmentedOptomechanical": (c_int, [_D, _D, POINTER(SimProperties), POINTER(COptomechanicalVariables), c_size_t]),
    "integrateAugmentedOptomechanicalSimulationRK4": (c_int, [_D, POINTER(_D), POINTER(c_size_t), POINTER(_D), POINTER(c_size_t),
                                                              POINTER(SimProperties), POINTER(RK4SolverOptions),
                                                              POINTER(COptomechanicalVariables), c_size_t]),
    "integrateAugmentedOptomechanicalSimulationRK4_freeMemory": (c_int, [_D, _D]),
    "rb_integrate_aug_rk4_host": (c_int, [_D, _D, c_size_t, POINTER(rb_props), POINTER(rb_opto), c_double, c_size_t]),
    "rb_real_rhs": (c_int, [_P, _P, _P]),
    "rb_perturbed_states": (c_int, [_P, _P, c_double, c_int, _P]),
    "rb_jacobian_create": (_P, [c_int, POINTER(rb_props)]),
    "rb_jacobian_destroy": (c_int, [_P]),
    "rb_jacobian_set_epsilon": (c_int, [_P, c_double]),
    "rb_jacobian_set_stream": (c_int, [_P, _P]),
    "rb_jacobian_solver": (_P, [_P]),
    "rb_jacobian_calculate": (c_int, [_P, _P, _P]),
    "rb_lu_solve": (c_int, [_P, _P, c_int, c_int, POINTER(c_int), _P]),
    "rb_gl2_default_options": (None, [POINTER(rb_gl2_options)]),
    "rb_gl2_create": (_P, [_P, _P, POINTER(rb_gl2_options)]),
    "rb_gl2_destroy": (c_int, [_P]),
    "rb_gl2_set_options": (c_int, [_P, POINTER(rb_gl2_options)]),
    "rb_gl2_initialize": (c_int, [_P, _P, c_int]),
    "rb_gl2_step": (c_int, [_P, c_double, POINTER(c_int)]),
    "rb_gl2_evolve": (c_int, [_P, c_double, c_double]),
    "rb_gl2_copy_trajectory": (c_int, [_P, POINTER(_D), POINTER(c_size_t), POINTER(_D), POINTER(c_size_t)]),
    "rb_gl2_dev_state": (_P, [_P]),
    "rb_gl2_get_state": (c_int, [_P, _D]),
    "rb_gl2_get_stats": (c_int, [_P, POINTER(rb_gl2_stats)]),
    "calculateJacobian": (c_int, [_D, _D, c_double, c_double, c_double, c_double, c_double, c_size_t]),
    "calculatePerturbedStates256": (c_int, [_D, _D, _D, _P, c_double, c_double, c_double, c_double, c_double]),
    "integrateSimulationGL2": (c_int, [_D, POINTER(_D), POINTER(c_size_t), POINTER(_D), POINTER(c_size_t), POINTER(SimProperties),
                                       POINTER(GaussLegendreOptions), c_size_t]),
    "integrateSimulationGL2_freeMemory": (c_int, [_D, _D]),
}

_lib = None


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m superfluid_dynamics_b200.build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class RobertsError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = load().rb_last_error()
        raise RobertsError(f"{what}: {msg.decode() if msg else 'error'}")
