"""ctypes binding of the C ABI declared in include/ilqc_b200.h.

The product path loads ilqc_b200/libilqc_b200.so (hand-written sm_100a CUDA).  There is NO CPU fallback:
if the library is missing or no CUDA device is present, using the engine raises.  (tests/hostsim builds
a CPU emulation of the same kernels for the `-m "not gpu"` tests; it is loaded only from tests/.)
"""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(PKG, "libilqc_b200.so")

ABI_VERSION = 3
# enums of include/ilqc_b200.h
MODEL_PENDULUM, MODEL_CARTPOLE, MODEL_CAR_SIMPLE, MODEL_CAR_BICYCLE = 0, 1, 2, 3
INT_EULER, INT_RK4_CST, INT_RK4 = 0, 1, 2
COST_CONTOURING, COST_EXACT = 0, 1
TP_VREF_SQUARED, TP_TREF, TP_DREF, TP_DREF_SQUARED = 0, 1, 2, 3
TERM_LOOP, TERM_REPEAT_END = 0, 1
ALGO_GD, ALGO_CLASSIC, ALGO_DDP = 0, 1, 2
APPROX_LINQUAD, APPROX_QUAD = 0, 1
STEP_DIR, STEP_REG, STEP_DIRVAR, STEP_REGVAR = 0, 1, 2, 3
RUNNING, CONVERGED, DIVERGED, STUCK, NO_DIRECTION = 0, 1, 2, 3, 4
BWD_LINQUAD, BWD_QUAD_NEWTON, BWD_QUAD_DDP = 0, 1, 2
BAD_DIR = {"check_total_value": 0, "flag": 1, "let_it_be": 2}
ROLL_EXACT, ROLL_LIN, ROLL_GD = 0, 1, 2
E_ARG, E_UNSUPPORTED, E_WORKSPACE = -1, -2, -3
STATUS_NAMES = {RUNNING: "running", CONVERGED: "converged", DIVERGED: "diverged", STUCK: "got stuck",
                NO_DIRECTION: "no direction"}

_d, _i, _p = C.c_double, C.c_int32, C.c_void_p


class ModelDesc(C.Structure):
    """struct ilqc_model_desc (include/ilqc_b200.h)."""
    _fields_ = (
        [(n, _i) for n in ("model", "integrator", "cost", "time_penalty", "horizon", "t0", "stay_put_start",
                           "has_x_limits", "n_seg", "termination", "reg_ctrl_is_tuple", "reserved0")]
        + [("dt", _d)]
        + [(n, _d) for n in ("reg_speed", "reg_ctrl", "reg_ctrl_tref", "reg_barrier", "x_min", "x_max")]
        + [("phys", _d * 8)]
        + [(n, _d) for n in ("reg_cont", "reg_lag", "reg_bar", "reg_obs", "vref", "constrain_angle", "acc_lo",
                             "acc_hi")]
        + [(n, _d) for n in ("Cm1", "Cm2", "Cr0", "Cr2", "Br", "Cr", "Dr", "Bf", "Cf", "Df", "m", "Iz", "lf",
                             "lr", "car_l", "car_w")]
        + [(n, _d) for n in ("obs_x", "obs_y", "track_len")]
        + [("knots", _p), ("coef", _p), ("pp_weights", _p), ("pp_t0", _p)]
    )


class AlgoDesc(C.Structure):
    """struct ilqc_algo_desc (include/ilqc_b200.h)."""
    _fields_ = ([(n, _i) for n in ("algo_type", "approx", "step_mode", "n_candidates", "line_search",
                                   "max_ls_trips", "compute_grad_norm", "scheduler", "handle_bad_dir", "expansion",
                                   "reserved1", "reserved2")] + [("pp_cost0", _p)])


_MP, _AP = C.POINTER(ModelDesc), C.POINTER(AlgoDesc)

# name -> (restype, argtypes) ; every symbol include/ilqc_b200.h declares
PROTOTYPES = {
    "ilqc_abi_version": (C.c_int, []),
    "ilqc_dims": (C.c_int, [_MP, C.POINTER(_i), C.POINTER(_i)]),
    "ilqc_lin_stride": (C.c_int, [_MP]),
    "ilqc_gain_stride": (C.c_int, [_MP]),
    "ilqc_rollout_cost": (C.c_int, [_MP, C.c_int, _p, _p, _p, _p, _p, _p]),
    "ilqc_linearize": (C.c_int, [_MP, C.c_int, _p, _p, _p, _p, _p, _p]),
    "ilqc_expand_dense": (C.c_int, [_MP, C.c_int, _p, _p, C.c_int, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "ilqc_lq_backward_dense": (C.c_int, [C.c_int] * 5 + [_p] * 8 + [_d, _p, _p, _p, _p]),
    "ilqc_rollout_lin_dense": (C.c_int, [C.c_int] * 5 + [_p] * 7),
    "ilqc_backward": (C.c_int, [_MP, C.c_int, C.c_int, C.c_int, _p, _p, _d, _p, _p, _p, _p, _p, _p, _p]),
    "ilqc_rollout": (C.c_int, [_MP, C.c_int, C.c_int, C.c_int] + [_p] * 13),
    "ilqc_grad_obj": (C.c_int, [_MP, C.c_int, _p, _p, _p, _p, _p]),
    "ilqc_solve_workspace_bytes": (C.c_size_t, [_MP, _AP, C.c_int]),
    "ilqc_solve": (C.c_int, [_MP, _AP, C.c_int, _p, _p, _p, C.c_int, _p, _p, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "ilqc_solve_host_workspace_bytes": (C.c_size_t, [_MP, _AP, C.c_int, C.c_int]),
    "ilqc_solve_host": (C.c_int, [_MP, _AP, C.c_int, _p, _p, _p, _p, C.c_int, _p, _p, _p, _p, _p, _p, C.c_size_t, _p]),
    "ilqc_mpc_num_windows": (C.c_int, [C.c_int] * 3),
    "ilqc_mpc_workspace_bytes": (C.c_size_t, [_MP, _AP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ilqc_mpc": (C.c_int, [_MP, _AP, C.c_int, _p] + [C.c_int] * 7 + [_p, C.POINTER(_i), _p, _p, _p, _p, C.c_size_t, _p]),
    "ilqc_newton_dense_workspace_bytes": (C.c_size_t, [_MP, C.c_int]),
    "ilqc_newton_dense": (C.c_int, [_MP, C.c_int] + [_p] * 8 + [C.c_size_t, _p]),
    "ilqc_traj_jacobian_workspace_bytes": (C.c_size_t, [_MP, C.c_int]),
    "ilqc_traj_jacobian": (C.c_int, [_MP, C.c_int, _p, _p, _p, _p, C.c_size_t, _p]),
    "ilqc_to_soa": (C.c_int, [_p, _p, C.c_int, C.c_int, _p]),
    "ilqc_from_soa": (C.c_int, [_p, _p, C.c_int, C.c_int, _p]),
    "ilqc_profile_enable": (C.c_int, [C.c_int]),
    "ilqc_set_tuning": (C.c_int, [C.c_char_p, C.c_longlong]),
    "ilqc_profile_read": (C.c_int, [C.POINTER(_d), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int]),
    "ilqc_launch_count": (C.c_longlong, [C.c_int]),
    "ilqc_fp64_peak": (C.c_int, [C.c_int, C.POINTER(_d), _p]),
    "ilqc_fastmath_eval": (C.c_int, [C.c_int, C.c_int, _p, _p, _p, _p, _p]),
}


def bind(path):
    """dlopen `path` and attach the prototypes; raises OSError/AttributeError if anything is missing."""
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.ilqc_abi_version() != ABI_VERSION:
        raise RuntimeError("ilqc_b200: ABI version mismatch in %s" % path)
    return lib


_cuda_lib = None


def cuda_lib():
    """The product library.  Fails loudly when it has not been built (python -m ilqc_b200.build)."""
    global _cuda_lib
    if _cuda_lib is None:
        if not os.path.exists(CUDA_LIB_PATH):
            raise RuntimeError(
                "ilqc_b200: %s not found. Build it with `python -m ilqc_b200.build` (nvcc, sm_100a). "
                "There is no CPU fallback." % CUDA_LIB_PATH)
        _cuda_lib = bind(CUDA_LIB_PATH)
    return _cuda_lib


class IlqcError(RuntimeError):
    pass


def check(rc, what):
    if rc == 0:
        return
    names = {E_ARG: "invalid argument", E_UNSUPPORTED: "unsupported configuration", E_WORKSPACE: "workspace too small"}
    if rc < 0:
        raise IlqcError("%s: %s (%d)" % (what, names.get(rc, "error"), rc))
    raise IlqcError("%s: CUDA error %d" % (what, rc))
