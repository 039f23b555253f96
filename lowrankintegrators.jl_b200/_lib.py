"""ctypes binding of libdlra.so (include/dlra.h).  There is NO fallback: if the shared library is missing or
cannot be loaded this module raises, and every product entry point fails loudly."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdlra.so")

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)
handle_t = C.c_void_p

OK, EINVAL, ECUDA, ENCCL, ESTATE, ENOMEM, EUNSUPPORTED, EMAXITERS = range(8)
RANK_ADAPTIVE, FORCE_GENERIC, AUG_BASIS_FIRST = 1, 2, 4
KSL_PRIMAL, KSL_DUAL, KSL_STRANG = 0, 1, 2
DATA_SNAPSHOT, DATA_DELTA = 0, 1
FLOW_K, FLOW_S, FLOW_L = 0, 1, 2
GREEDY_DATA, GREEDY_HYBRID = 0, 1
ODE_EULER, ODE_RK4, ODE_TSIT5_FIXED, ODE_TSIT5 = 0, 1, 2, 3
OP_NONE, OP_DENSE, OP_CSR, OP_IDENTITY_SCALED = 0, 1, 2, 3


class Operator(C.Structure):
    """struct dlra_operator"""
    _fields_ = [("kind", C.c_int), ("rows", C.c_int64), ("cols", C.c_int64), ("dense", C.c_void_p), ("ld", C.c_int64),
                ("rowptr", C.c_void_p), ("colind", C.c_void_p), ("values", C.c_void_p), ("scale", C.c_double)]


# every symbol include/dlra.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "dlra_create": (C.c_int, [C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(handle_t)]),
    "dlra_destroy": (C.c_int, [handle_t]),
    "dlra_last_error": (C.c_char_p, [handle_t]),
    "dlra_version": (C.c_char_p, []),
    "dlra_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "dlra_comm_init": (C.c_int, [handle_t, C.c_int, C.c_int, C.c_void_p]),
    "dlra_p2p_export": (C.c_int, [handle_t, C.c_void_p]),
    "dlra_p2p_import": (C.c_int, [handle_t, C.c_int, C.c_int, C.c_void_p]),
    "dlra_set_factors_host": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]),
    "dlra_set_factors": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int]),
    "dlra_get_factors_host": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, c_int_p]),
    "dlra_get_factors": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, c_int_p]),
    "dlra_get_rank": (C.c_int, [handle_t, c_int_p]),
    "dlra_save_factors_async": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, c_int_p]),
    "dlra_save_wait": (C.c_int, [handle_t]),
    "dlra_truncated_svd": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_int, C.c_double, C.c_int, C.c_int, C.c_uint64]),
    "dlra_factor_ptrs": (C.c_int, [handle_t, C.POINTER(C.c_void_p), c_i64_p, C.POINTER(C.c_void_p), c_i64_p,
                                   C.POINTER(C.c_void_p), c_i64_p, c_int_p]),
    "dlra_data_init": (C.c_int, [handle_t, C.c_void_p, C.c_int64]),
    "dlra_data_init_host": (C.c_int, [handle_t, C.c_void_p, C.c_int64]),
    "dlra_data_push": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_int]),
    "dlra_data_push_host": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_int]),
    "dlra_rhs_set": (C.c_int, [handle_t, C.POINTER(Operator), C.POINTER(Operator), C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                               C.c_int, C.POINTER(Operator), C.POINTER(Operator), C.c_double]),
    "dlra_rhs_add_term": (C.c_int, [handle_t, C.POINTER(Operator), C.POINTER(Operator)]),
    "dlra_set_substepper": (C.c_int, [handle_t, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]),
    "dlra_set_substepper_maxiters": (C.c_int, [handle_t, C.c_int, C.c_int64]),
    "dlra_step_ksl": (C.c_int, [handle_t, C.c_int, C.c_double, C.c_double]),
    "dlra_step_bug": (C.c_int, [handle_t, C.c_double, C.c_double]),
    "dlra_step_rabug": (C.c_int, [handle_t, C.c_double, C.c_double, C.c_double, C.c_int64, c_int_p, c_int_p]),
    "dlra_step_greedy": (C.c_int, [handle_t, C.c_double, C.c_double]),
    "dlra_step_greedy_two_factor": (C.c_int, [handle_t, C.c_int, C.c_int, C.c_double, C.c_double]),
    "dlra_normal_component": (C.c_int, [handle_t, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_int64,
                                        c_double_p]),
    "dlra_sync": (C.c_int, [handle_t]),
    "dlra_wait_stream": (C.c_int, [handle_t, C.c_void_p]),
    "dlra_get_stream": (C.c_int, [handle_t, C.POINTER(C.c_void_p)]),
    "dlra_progress": (C.c_int, [handle_t, c_i64_p, c_i64_p, C.c_int64]),
    "dlra_reconstruct_error": (C.c_int, [handle_t, C.c_void_p, C.c_int64, c_double_p]),
    "dlra_reconstruct": (C.c_int, [handle_t, C.c_void_p, C.c_int64]),
    "dlra_stats": (C.c_int, [handle_t, c_i64_p, c_i64_p, c_double_p, c_double_p, C.c_int]),
    "dlra_set_profiling": (C.c_int, [handle_t, C.c_int]),
    "dlra_pass_breakdown": (C.c_int, [handle_t, c_i64_p, c_double_p, c_double_p, c_double_p]),
    "dlra_event_record": (C.c_int, [handle_t, C.c_int]),
    "dlra_event_elapsed_ms": (C.c_int, [handle_t, C.c_int, C.c_int, c_double_p]),
}

_lib = None


class DLRAError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libdlra error {code}: {msg}")
        self.code = code


def load():
    """Load libdlra.so once.  Raises (never falls back) when the CUDA extension is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or `make -C {os.path.join(_HERE, 'csrc')}`; there is no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(h, rc):
    if rc != OK:
        msg = load().dlra_last_error(h)
        raise DLRAError(rc, msg.decode() if msg else "unknown")
