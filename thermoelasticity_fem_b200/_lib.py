"""ctypes binding of libtefem_b200.so (C ABI declared in include/tefem.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is visible, every
entry point of the product path raises.  (Importing this module without the library only raises
when a function is first used, so that host-side logic stays importable in a GPU-less container.)
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TEFEM_LIB") or os.path.join(_PKG, "libtefem_b200.so")    # TEFEM_LIB: build variants (experiments)

OK = 0
ERR_INVALID, ERR_CUDA, ERR_NEG_JACOBIAN, ERR_ZERO_JACOBIAN, ERR_NO_CONVERGE, ERR_BREAKDOWN, ERR_NCCL = -1, -2, -3, -4, -5, -6, -7
OP_M, OP_K, OP_D = 1, 2, 4
KIND = {"Muu": 1, "Dtu": 2, "Dtt": 4, "Kuu": 8, "Kut": 16, "Ktt": 32}
INT32_MAX = 2 ** 31 - 1

c_void_p, c_int, c_int32, c_int64, c_double, c_size_t, c_char_p = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_double, ctypes.c_size_t, ctypes.c_char_p)
P = c_void_p

_SIGNATURES = {
    "tefem_last_error": [c_char_p, c_size_t],
    "tefem_version": [c_char_p, c_size_t],
    "tefem_quadrature_tables": [P, P, P],
    "tefem_element_matrices": [P, P, P, P, c_int, P, c_int64, c_int, P, P, P, P, P, P, P, P],
    "tefem_assemble": [P, c_int, c_int32, P, P, P, P, P, P, P, P, P, P, c_int64, c_int, c_double, c_double,
                       P, P, P, P, P],
    "tefem_assemble_set_threads": [c_int],
    "tefem_check_jacobian": [P, P, P],
    "tefem_schedule_bank_numbering": [c_int32, P, P, P, P, c_int, c_int, c_int, P, P, P],
    "tefem_planes_spmv": [P, P, c_int32, c_int64, P, P, P, c_double, c_double, P, P],
    "tefem_form_system": [P, P, c_int32, c_int64, P, c_double, P, c_int, c_double, P, c_double, P, P, P],
    "tefem_block_jacobi_scale": [P, P, c_int32, P, P, P, c_int, P],
    "tefem_block_jacobi_apply": [P, c_int32, P, P, P, P],
    "tefem_spmv": [P, P, c_int32, P, P, P, P],
    "tefem_spmv_set_config": [c_int, c_int, c_int],
    "tefem_solver_create": [P, c_int32, c_int32, P, P],
    "tefem_solver_destroy": [P],
    "tefem_solver_profile": [P, c_int, P, P],
    "tefem_solver_halo_exchange": [P, P, P],
    "tefem_solver_set_precond": [P, P],
    "tefem_solver_set_interior_range": [P, c_int32, c_int32, c_int],
    "tefem_precond_create": [P, c_int32, c_int32, P, P, P, P, c_double, P, P, P, P, c_double, c_int, c_int, c_int32, P, P, P, P, P, P,
                             P, P, P],
    "tefem_precond_destroy": [P],
    "tefem_precond_apply": [P, P, P, P],
    "tefem_solver_set_halo": [P, c_int32, c_int, P, P, P, P, P, P, P, c_int32, P, c_int32],
    "tefem_precond_set_agg_ranks": [P, P],
    "tefem_precond_set_agg_list": [P, P, c_int32],
    "tefem_precond_set_level1_half": [P, P],
    "tefem_precond_set_cycle": [P, c_int],
    "tefem_precond_set_coarse_mode": [P, c_int, P],
    "tefem_precond_set_cycle_masks": [P, P, P],
    "tefem_bicgstab": [P, P, P, P, P, P, c_double, c_int, c_int, c_double, P, P],
    "tefem_cg": [P, P, P, P, P, P, P, c_double, c_int, c_int, c_double, P, P],
    "tefem_newmark_predict": [c_int64, P, P, P, c_double, c_double, c_double, P, P, P, P],
    "tefem_newmark_correct": [c_int64, P, P, P, P, c_double, c_double, c_double, P, P],
    "tefem_load_axpy": [c_int64, P, P, P, P, P],
    "tefem_load_axpy_dev": [c_int64, P, P, P, P, P],
    "tefem_lincomb3": [c_int64, c_double, P, c_double, P, c_double, P, P, P],
    "tefem_comm_unique_id": [c_char_p, P],
    "tefem_comm_create": [P, c_char_p, P, c_int, c_int],
    "tefem_comm_destroy": [P],
    "tefem_comm_allreduce_sum": [P, P, c_int, P],
    "tefem_comm_shm_alloc": [P, c_int64, P],
    "tefem_comm_shm_open": [P, P],
    "tefem_comm_check": [P],
    "tefem_comm_reset_error": [P],
    "tefem_comm_set_timeout": [P, c_double],
    "tefem_comm_vec_alloc": [P, c_int, c_int64, P],
    "tefem_comm_vec_open": [P, P],
    "tefem_solver_peer_slots": [],
}
_RESTYPES = {"tefem_solver_work_doubles": (c_int64, [c_int32, c_int32]), "tefem_launch_count": (c_int64, []),
             "tefem_precond_work_doubles": (c_int64, [c_int32, c_int32]),
             "tefem_precond_peer_doubles": (c_int64, [c_int32, c_int32, c_int, c_int]),
             "tefem_precond_cycle_doubles": (c_int64, [c_int32, c_int32, c_int, c_int])}

_lib = None


class TefemError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"[tefem {code}] {message}")
        self.code = code
        self.message = message


class ConvergenceError(TefemError):
    pass


def declared_symbols():
    return sorted(list(_SIGNATURES) + list(_RESTYPES))


def load():
    """Loads the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -m thermoelasticity_fem_b200.build or __graft_entry__.build()). "
            "thermoelasticity_fem_b200 has no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    for name, (res, argtypes) in _RESTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = res
    _lib = lib
    return lib


def last_error():
    buf = ctypes.create_string_buffer(512)
    load().tefem_last_error(buf, 512)
    return buf.value.decode()


def check(rc):
    """Maps a status code to the exception the reference would have raised where it has one."""
    if rc == OK:
        return
    msg = last_error()
    if rc in (ERR_NEG_JACOBIAN, ERR_ZERO_JACOBIAN):
        raise ValueError(msg)                      # elements.py:207-210
    if rc in (ERR_NO_CONVERGE, ERR_BREAKDOWN):
        raise ConvergenceError(rc, msg)
    raise TefemError(rc, msg)


def version():
    buf = ctypes.create_string_buffer(128)
    check(load().tefem_version(buf, 128))
    return buf.value.decode()


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array, or NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous()
        return t.data_ptr()
    return t.ctypes.data


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("thermoelasticity_fem_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
