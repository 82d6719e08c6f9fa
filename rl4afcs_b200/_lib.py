"""ctypes binding of librl4afcs_b200.so (include/rl4afcs_b200.h).

There is deliberately no fallback: if the CUDA library is missing, fails to load, or the
device is not an sm_100 GPU, the calls raise.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RL4AFCS_LIB") or os.path.join(_HERE, "librl4afcs_b200.so")   # override: tuning experiments only

# enums of include/rl4afcs_b200.h
FP64, FP32, MIXED = 0, 1, 2
POLICY = {"fp64": FP64, "fp32": FP32, "mixed": MIXED}
ELIG = {None: 0, "none": 0, "accumulating": 1, "replacing": 2}
FAULT = {None: 0, "none": 0, "invert_elevator": 1, "damp_elevator": 2, "shift_cg": 3}
FAULT_COUNT = 4

SPE = dict(X=0, XPREV=2, THETA=4, COV=10, CGRAD_PREV=19, EPS=20, EPS_NORM=22, SUM_C=23, SUM_ABS_E=24,
           EA=25, EC_H=33, EC_W1R0=37, EC_W1R1=41, COUNT=45)
SPN = dict(A=0, APREV=1, W1A=2, W2A=6, W1C=10, W2C=14, W1T=22, W2T=26, MPREV=34, ETA_A=38, ETA_C=39, COUNT=40)
SPI = dict(COOLDOWN=0, FLAGS=1, DIVERGED_STEP=2, CONV_STEP=3, COUNT=4)
SPF = dict(CHANGED=1, LR_INIT=2, LAMBDA_LOW=4, X_NAN=8)
HP = dict(ETA_A_H=0, ETA_A_L=1, ETA_C_H=2, ETA_C_L=3, LAMBDA_H=4, LAMBDA_L=5, GAMMA=6, GAMMA_SQ=7, TAU=8,
          KAPPA=9, RLS_GAMMA=10, RLS_COV0=11, ERROR_THRESH_DEG=12, REF_AMP=13, COUNT=14)
HPI = dict(MULTISTEP=0, WARMUP_STEPS=1, COOLDOWN_STEPS=2, FAULT_STEP=3, FAULT_KIND=4, ELIG_A=5, ELIG_C=6, TRACKED_Q=7,
           COUNT=8)
LOG_NONE, LOG_BASIC, LOG_FULL = 0, 1, 2
LB = dict(X=0, A=2, C=3, REF=4, E=5, COUNT=6)
LF = dict(AW1=6, AW2=10, CW1=14, CW2=18, AE=26, CE=34, AGRAD=46, CGRAD=54, PARAMS=66, COV=72, EPS_NORM=81,
          EPS_ABS=82, LAM=84, LAM_T=86, TD=88, DADZ=90, M=91, LOSS_GRAD=95, COUNT=96)


class SpState(ctypes.Structure):
    _fields_ = [("env", ctypes.c_void_p), ("net", ctypes.c_void_p), ("ints", ctypes.c_void_p),
                ("stride", ctypes.c_int64)]


class SpParams(ctypes.Structure):
    _fields_ = [
        ("A", (ctypes.c_double * 4) * FAULT_COUNT),
        ("B", (ctypes.c_double * 2) * FAULT_COUNT),
        ("dt", ctypes.c_double),
        ("hp", ctypes.c_double * HP["COUNT"]),
        ("hpi", ctypes.c_int32 * HPI["COUNT"]),
        ("q3_alias", ctypes.c_int32),
        ("q7_numpy1", ctypes.c_int32),
        ("hp_agent", ctypes.c_void_p * HP["COUNT"]),
        ("hpi_agent", ctypes.c_void_p * HPI["COUNT"]),
    ]


# ---- nonlinear path ----
NLE = dict(XFULL=0, XACT=12, XLON=15, XPREVLON=18, THETA=21, COV=33, CGRAD_PREV=49, EPS=50, EPS_NORM=53, RSE=54,
           NZ_PEAK=56, ETA_A=57, ETA_C=58, LAMBDAA=59, GL=60, EA=61, RSE_FLIGHT=111, COUNT=113)
NLN = dict(S=0, SPREV=4, A=8, APREV=9, W1A=10, W2A=50, W1C=60, W2C=100, W1T=130, W2T=170, MPREV=200, LR_A=209,
           LR_C=210, COUNT=211)
NLI = dict(COOLDOWN=0, DIVERGED_STEP=1, STEPP=2, PYFLOAT_MASK=3, COUNT=4)
NHP = dict(ETA_A_H=0, ETA_A_L=1, ETA_C_H=2, ETA_C_L=3, LAMBDA_H=4, LAMBDA_L=5, GAMMA=6, GAMMA_SQ=7, TAU=8, LR_DECAY=9,
           RLS_GAMMA=10, RLS_COV0=11, Q_SYM=12, LAMBDA_T=13, LAMBDA_S=14, DAMP_FACTOR=15, CG_SHIFT=16, COUNT=17)
NHPI = dict(MULTISTEP=0, WARMUP_STEPS=1, COOLDOWN_STEPS=2, FAULT_STEP=3, FAULT_DAMP=4, FAULT_SAT=5, ELIG_A=6, FLIGHT_STEP=7,
            NUMPY2=8, COUNT=9)
NL_DAMP = {None: 0, "none": 0, "damp_elevator": 1, "damp_aileron": 2, "damp_rudder": 3, "damp_all": 4, "shift_cg": 5,
           "slow_all": 6}
NL_SAT = {None: 0, "none": 0, "saturate_elevator": 1, "saturate_aileron": 2, "saturate_rudder": 3}
NLL = dict(XFULL=0, A=12, E_THETA=13, REWARD=14, SURF=15, COUNT=18)
# enum rl4_nl_fulllog_field (name -> (offset, width))
NLF_FIELDS = dict(ETA_A=(0, 1), XFULL=(1, 12), RSE=(13, 2), X=(15, 3), A_CMD=(18, 1), A_EFF=(19, 1), S=(20, 1), YREF=(21, 1),
                  E=(22, 1), A_W1=(23, 40), A_W2=(63, 10), C_W1=(73, 40), C_W2=(113, 30), A_GRAD=(143, 50), C_GRAD=(193, 70),
                  RLS_PARAMS=(263, 12), RLS_COV=(275, 16), RLS_EPS=(291, 3), RLS_EPS_NORM=(294, 1), A=(295, 1), REWARD=(296, 1))
NLF = {k: v[0] for k, v in NLF_FIELDS.items()}
NLF["COUNT"] = 297
NLM = dict(E=0, THETA=1, ALPHA=2, Q=3, V=4, H=5, A_CMD=6, A_EFF=7, WA_NORM=8, WC_NORM=9, RLS_EPS=10, COUNT=11)
INTEGRATOR = {"rk4": 0, "ode5": 1}
CIT_FIELDS = ["m", "S", "c", "b", "Ixx", "Iyy", "Izz", "Ixz", "g", "CL0", "CLa", "CLq", "CLde", "CLflap", "al_stall",
              "CD0", "CDk", "CDgear", "CDflap", "CDstall", "Cm0", "Cma", "Cmq", "Cmde", "Cmflap", "Cmstall",
              "CYb", "CYp", "CYr", "CYda", "CYdr", "Clb", "Clp", "Clr", "Clda", "Cldr",
              "Cnb", "Cnp", "Cnr", "Cnda", "Cndr", "Tstatic", "TV", "Vref", "xcg_gain",
              "inv_m", "inv_Iyy", "inv_gam", "inv_al_stall", "inv_c", "inv_b"]


CIT_NPOLY = 21


class CitParams(ctypes.Structure):
    _fields_ = [(f, ctypes.c_double) for f in CIT_FIELDS] + [("zeta_per_m", ctypes.c_double),
                                                              ("rho_poly", ctypes.c_double * CIT_NPOLY),
                                                              ("lapse_poly", ctypes.c_double * CIT_NPOLY)]


class NlState(ctypes.Structure):
    _fields_ = [("env", ctypes.c_void_p), ("net", ctypes.c_void_p), ("ints", ctypes.c_void_p), ("stride", ctypes.c_int64)]


class NlParams(ctypes.Structure):
    _fields_ = [
        ("plant", CitParams), ("trim_input", ctypes.c_double * 11), ("dt", ctypes.c_double),
        ("hp", ctypes.c_double * NHP["COUNT"]), ("noise_std", ctypes.c_double * 4),
        ("omega0", ctypes.c_double), ("omega_slow", ctypes.c_double), ("rate_limit", ctypes.c_double),
        ("limit_deg", ctypes.c_double * 3), ("sat_limit", ctypes.c_double * 3),
        ("hpi", ctypes.c_int32 * NHPI["COUNT"]), ("integrator", ctypes.c_int32),
        ("hp_agent", ctypes.c_void_p * NHP["COUNT"]), ("hpi_agent", ctypes.c_void_p * NHPI["COUNT"]),
    ]


class SpLog(ctypes.Structure):
    _fields_ = [("buf", ctypes.c_void_p), ("level", ctypes.c_int32), ("every", ctypes.c_int32),
                ("n_agents_logged", ctypes.c_int64)]


class SpHostIO(ctypes.Structure):
    _fields_ = [("x0", ctypes.c_void_p), ("w1a", ctypes.c_void_p), ("w2a", ctypes.c_void_p),
                ("w1c", ctypes.c_void_p), ("w2c", ctypes.c_void_p), ("ref_base", ctypes.c_void_p),
                ("out_env", ctypes.c_void_p), ("out_net", ctypes.c_void_p), ("out_ints", ctypes.c_void_p),
                ("out_mask", ctypes.c_int32), ("reserved", ctypes.c_int32)]


class NlHostIO(ctypes.Structure):
    _fields_ = [("w1a", ctypes.c_void_p), ("w2a", ctypes.c_void_p), ("w1c", ctypes.c_void_p), ("w2c", ctypes.c_void_p),
                ("theta_ref", ctypes.c_void_p), ("noise", ctypes.c_void_p), ("noise_seed", ctypes.c_uint64), ("noise_agent0", ctypes.c_int64),
                ("out_env", ctypes.c_void_p), ("out_net", ctypes.c_void_p), ("out_ints", ctypes.c_void_p),
                ("out_mask", ctypes.c_int32), ("reserved", ctypes.c_int32)]


ABI_VERSION = 3          # RL4_ABI_VERSION of include/rl4afcs_b200.h these struct layouts were written against
OUT = dict(STATS=1, WEIGHTS=2, RLS=4, STATE=8, TRACES=16, TARGET=32, ALL=63)
SPS = dict(SUM_C=0, CONV_TIME=1, DIVERGED=2, UNSTEADY=3, MEAN_ABS_E=4, NMAE=5, COUNT=6)
NLS = dict(RSE_WARMUP=0, RSE_FLIGHT=1, RSE_LAT=2, NZ_PEAK=3, DIVERGED=4, COUNT=5)


EXPORTS = [
    "rl4_abi_version", "rl4_last_error", "rl4_device_check",
    "rl4_sp_init", "rl4_sp_run", "rl4_sp_env_step", "rl4_sp_rls_update",
    "rl4_sp_critic_forward", "rl4_sp_actor_forward", "rl4_sp_critic_weight_update",
    "rl4_ctx_create", "rl4_ctx_destroy", "rl4_sp_episode_host",
    "rl4_nl_rls_update", "rl4_peak_fma", "rl4_launch_count", "rl4_test_math", "rl4_test_t13_div_f32",
    "rl4_nl_init", "rl4_nl_run", "rl4_nl_env_step", "rl4_nl_default_params", "rl4_nl_critic_forward", "rl4_nl_actor_forward",
    "rl4_sizeof_sp_params", "rl4_sizeof_nl_params", "rl4_soft_update", "rl4_actor_weight_update", "rl4_sp_agent_stats",
    "rl4_nl_agent_stats", "rl4_stats_reduce_work_doubles", "rl4_stats_reduce", "rl4_nl_noise_fill", "rl4_nl_episode_host", "rl4_test_rcp_f32",
    "rl4_nl_trim_state",
    "rl4_dasmat_available", "rl4_dasmat_image_bytes", "rl4_dasmat_state_words", "rl4_dasmat_word_x", "rl4_dasmat_word_engine",
    "rl4_dasmat_initialize", "rl4_dasmat_reset", "rl4_dasmat_step", "rl4_dasmat_broadcast", "rl4_nl_env_step_dasmat", "rl4_nl_run_dasmat",
]
_NOT_STATUS = ("rl4_last_error", "rl4_launch_count", "rl4_abi_version", "rl4_sizeof_sp_params", "rl4_sizeof_nl_params",
               "rl4_stats_reduce_work_doubles", "rl4_dasmat_available", "rl4_dasmat_image_bytes", "rl4_dasmat_state_words",
               "rl4_dasmat_word_x", "rl4_dasmat_word_engine")

_lib = None


class Rl4Error(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the CUDA library; raises when it is absent (no CPU fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise Rl4Error(f"{LIB_PATH} is missing: build it with `python -m rl4afcs_b200.build` "
                       "(nvcc, sm_100a). rl4afcs_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    L.rl4_abi_version.restype = ctypes.c_int
    # a stale library from an older checkout would take by-value structs of another layout: refuse it
    try:
        got = int(L.rl4_abi_version())
        L.rl4_sizeof_sp_params.restype = i64
        L.rl4_sizeof_nl_params.restype = i64
        sizes = (int(L.rl4_sizeof_sp_params()), int(L.rl4_sizeof_nl_params()))
    except AttributeError as e:
        raise Rl4Error(f"{LIB_PATH} predates ABI {ABI_VERSION} ({e}); rebuild it with `python -m rl4afcs_b200.build --force`")
    want = (ctypes.sizeof(SpParams), ctypes.sizeof(NlParams))
    if got != ABI_VERSION or sizes != want:
        raise Rl4Error(f"{LIB_PATH}: ABI version {got} / struct sizes {sizes}, this binding expects {ABI_VERSION} / {want}; "
                       "rebuild the library with `python -m rl4afcs_b200.build --force`")
    L.rl4_last_error.restype = ctypes.c_char_p
    L.rl4_device_check.argtypes = [ctypes.c_int]
    L.rl4_sp_init.argtypes = [ctypes.c_int, ctypes.POINTER(SpParams), vp, vp, vp, vp, vp, i64, SpState, i64, vp]
    L.rl4_sp_run.argtypes = [ctypes.c_int, ctypes.POINTER(SpParams), vp, i32, i32, SpState, i64, i32, SpLog, vp]
    L.rl4_sp_env_step.argtypes = [ctypes.c_int, ctypes.POINTER(SpParams), vp, i32, vp, vp, vp, vp, vp, i64, i64, vp]
    L.rl4_sp_rls_update.argtypes = [ctypes.c_int, ctypes.POINTER(SpParams), vp, vp, vp, vp, vp, vp, vp, i64, i64, vp]
    L.rl4_sp_critic_forward.argtypes = [ctypes.c_int, vp, vp, vp, vp, vp, dbl, i32, i64, i64, vp]
    L.rl4_sp_actor_forward.argtypes = [ctypes.c_int, vp, vp, vp, vp, vp, vp, dbl, i32, i64, i64, vp]
    L.rl4_sp_critic_weight_update.argtypes = [ctypes.c_int, vp, vp, vp, i64, i64, vp]
    L.rl4_ctx_create.argtypes = [ctypes.c_int, ctypes.c_int, i64, i32, ctypes.POINTER(vp)]
    L.rl4_ctx_destroy.argtypes = [vp]
    L.rl4_sp_episode_host.argtypes = [vp, ctypes.POINTER(SpParams), ctypes.POINTER(SpHostIO), i64, i32, i32]
    L.rl4_peak_fma.argtypes = [ctypes.c_int, ctypes.POINTER(dbl), vp]
    L.rl4_launch_count.restype = i64
    L.rl4_test_math.argtypes = [ctypes.c_int, vp, vp, vp, i64, vp]
    L.rl4_nl_init.argtypes = [ctypes.c_int, ctypes.POINTER(NlParams), vp, vp, vp, vp, i64, NlState, i64, vp]
    L.rl4_nl_run.argtypes = [ctypes.c_int, ctypes.POINTER(NlParams), vp, vp, i64, i32, i32, NlState, i64, SpLog, vp]
    L.rl4_nl_env_step.argtypes = [ctypes.POINTER(NlParams), vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp]
    L.rl4_nl_default_params.argtypes = [ctypes.POINTER(NlParams)]
    L.rl4_nl_trim_state.argtypes = [ctypes.POINTER(NlParams), vp, vp]
    L.rl4_nl_rls_update.argtypes = [ctypes.POINTER(NlParams), vp, vp, vp, vp, vp, vp, vp, i64, i64, vp]
    L.rl4_nl_critic_forward.argtypes = [ctypes.c_int, vp, vp, vp, vp, i64, i64, vp]
    L.rl4_nl_actor_forward.argtypes = [ctypes.c_int, vp, vp, vp, vp, vp, vp, dbl, i32, i32, i64, i64, vp]
    L.rl4_test_t13_div_f32.argtypes = [ctypes.c_uint32, ctypes.c_uint32, vp, vp]
    L.rl4_test_rcp_f32.argtypes = [ctypes.c_uint32, ctypes.c_uint32, vp, vp]
    L.rl4_soft_update.argtypes = [ctypes.c_int, vp, vp, dbl, i32, i64, i64, vp]
    L.rl4_actor_weight_update.argtypes = [ctypes.c_int, vp, vp, vp, i32, i64, i64, vp]
    L.rl4_sp_agent_stats.argtypes = [ctypes.c_int, ctypes.POINTER(SpParams), SpState, i64, i32, dbl, dbl, vp, i64, vp]
    L.rl4_nl_agent_stats.argtypes = [NlState, i64, vp, i64, vp]
    L.rl4_stats_reduce_work_doubles.restype = i64
    L.rl4_stats_reduce.argtypes = [vp, i64, i32, vp, i64, vp, vp, i64, vp]
    L.rl4_nl_noise_fill.argtypes = [ctypes.c_uint64, i64, i32, i32, i64, vp, i64, vp]
    L.rl4_nl_episode_host.argtypes = [vp, ctypes.POINTER(NlParams), ctypes.POINTER(NlHostIO), i64, i32]
    L.rl4_dasmat_image_bytes.restype = i64
    L.rl4_dasmat_initialize.argtypes = [vp, vp]
    L.rl4_dasmat_reset.argtypes = [vp, vp, i64, i64, vp]
    L.rl4_dasmat_step.argtypes = [vp, vp, i64, i64, vp, i64, i32, vp, i64, vp, vp, vp]
    L.rl4_dasmat_broadcast.argtypes = [vp, i64, i64, vp, i64, i64, vp]
    L.rl4_nl_env_step_dasmat.argtypes = [ctypes.POINTER(NlParams), vp, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i64, vp, vp, i64, vp, vp]
    L.rl4_nl_run_dasmat.argtypes = [ctypes.c_int, ctypes.POINTER(NlParams), vp, vp, i64, i32, i32, NlState, i64, SpLog, vp, vp, i64, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if name not in _NOT_STATUS:
            fn.restype = ctypes.c_int
    _lib = L
    return L


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().rl4_last_error().decode("utf-8", "replace")
        raise Rl4Error(f"{what} failed (rc={rc}): {msg}")
