"""ctypes binding of libdjets_b200.so (include/djets.h).

This is the Python twin of the Julia ``ccall`` shim (julia/DistributedJetsB200.jl): the same
symbols, the same argument order, nothing else.  There is no CPU fallback: if the shared library
has not been built the import fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdjets_b200.so")

F32, F64, C64, C128 = 0, 1, 2, 3
BLOCK_ZERO, BLOCK_DENSE, BLOCK_DIAG = 0, 1, 2
OK, ERR_INVALID, ERR_CUDA, ERR_NCCL, ERR_MISMATCH, ERR_UNSUPPORTED, ERR_NOT_LOCAL = range(7)
(RED_SUM, RED_MIN, RED_MAX, RED_ABSSUM, RED_ABSMAX, RED_ABSMIN, RED_SUMSQ, RED_NNZ, RED_SUMPOW,
 RED_SUMSQ_SHIFT) = range(10)
K_GEMV_N, K_GEMV_T, K_FINISH, K_REDUCE, K_ELTWISE, K_COUNT = range(6)
# operation codes of djets_vec_eval (DJETS_EV_*) and djets_scalar_op (DJETS_SC_*)
(EV_IN, EV_SC, EV_ADD, EV_SUB, EV_MUL, EV_DIV, EV_POW, EV_MIN, EV_MAX, EV_NEG, EV_ABS, EV_SQRT, EV_SQUARE, EV_CUBE,
 EV_EXP, EV_LOG, EV_ADDS, EV_SUBS, EV_RSUBS, EV_MULS, EV_RMULS, EV_DIVS, EV_RDIVS, EV_POWS, EV_MINS,
 EV_MAXS) = range(1, 27)
SC_ADD, SC_SUB, SC_MUL, SC_DIV, SC_NEG, SC_SQRT, SC_COPY, SC_ABS, SC_MAX, SC_MIN = range(10)
KERNEL_NAMES = {K_GEMV_N: "gemv_n", K_GEMV_T: "gemv_t", K_FINISH: "finish", K_REDUCE: "reduce",
                K_ELTWISE: "eltwise"}


# int32 (*djets_allgather_fn)(void* user, const void* send, void* recv, int64 bytes)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


class DJetsError(RuntimeError):
    """Any failure reported by the library (status != 0)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"[djets status {code}] {msg}")
        self.code = code


class DimensionMismatch(DJetsError):
    """Layout / dtype mismatch (the reference raises Julia's DimensionMismatch / MethodError)."""


class NotLocal(DJetsError):
    """The block lives on a rank driven by another process."""


if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C distributedjets.jl_b200/csrc`. This backend has no CPU fallback.")

lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)

i32, i64, f64, f32 = C.c_int32, C.c_int64, C.c_double, C.c_float
p_i32, p_i64, p_f64, p_f32 = C.POINTER(i32), C.POINTER(i64), C.POINTER(f64), C.POINTER(f32)
vp = C.c_void_p
p_vp = C.POINTER(vp)

# name -> argument types; every function returns int32 status except version / last_error
SIGNATURES = {
    "djets_even_cuts": [i64, i32, p_i64],
    "djets_grid_rank": [i32, i32, i32, i32, p_i32, p_i32],
    "djets_plan_tiles": [i32, i64, i64, i32, p_i64, p_i64, p_i64],
    "djets_plan_chains": [i32, i32, i64, i64, i64, i64, i64, i32, i32, p_i32, p_i32],
    "djets_plan_exchange": [i32, i32, i32, i32, i32, p_i32, i32, p_i32],
    "djets_plan_eval": [C.POINTER(C.c_uint8), i32, i32, i32, p_i32, p_i32],
    "djets_device_count": [p_i32],
    "djets_nccl_unique_id": [C.POINTER(C.c_uint8)],
    "djets_host_alloc": [i64, p_vp],
    "djets_host_free": [vp],
    "djets_ctx_create": [i32, i32, p_i32, p_i32, C.POINTER(C.c_uint8), p_vp],
    "djets_ctx_destroy": [vp],
    "djets_ctx_set_host_allgather": [vp, vp, vp],
    "djets_ctx_sync": [vp],
    "djets_ctx_info": [vp, p_i32, p_i32, p_i32],
    "djets_ctx_is_local": [vp, i32, p_i32],
    "djets_timer_start": [vp],
    "djets_timer_stop": [vp, p_f32],
    "djets_ctx_kernel_timing": [vp, i32],
    "djets_ctx_kernel_stats": [vp, i32, p_i64, p_f64],
    "djets_ctx_reset_stats": [vp],
    "djets_ctx_set_param": [vp, C.c_char_p, i64],
    "djets_vec_create": [vp, i32, i32, p_i32, p_i64, p_i64, p_vp],
    "djets_vec_similar": [vp, i32, p_vp],
    "djets_vec_destroy": [vp],
    "djets_vec_info": [vp, p_i32, p_i64, p_i64, p_i32],
    "djets_vec_part_info": [vp, i32, p_i32, p_i64, p_i64, p_i64, p_i64],
    "djets_vec_block_info": [vp, i64, p_i32, p_i64, p_i64],
    "djets_vec_setblock": [vp, i64, vp],
    "djets_vec_getblock": [vp, i64, vp],
    "djets_vec_getindex": [vp, i64, p_f64],
    "djets_vec_getindex_c": [vp, i64, p_f64],
    "djets_vec_fill_c": [vp, f64, f64],
    "djets_vec_dot_c": [vp, vp, p_f64],
    "djets_vec_collect": [vp, vp],
    "djets_vec_scatter": [vp, vp],
    "djets_vec_fill": [vp, f64],
    "djets_vec_lincomb": [vp, i32, p_f64, p_vp, i32],
    "djets_vec_binary": [vp, i32, vp, vp],
    "djets_vec_dot": [vp, vp, p_f64],
    "djets_vec_norm": [vp, f64, p_f64],
    "djets_vec_reduce": [vp, i32, f64, p_f64],
    "djets_vec_reduce_parts": [vp, i32, f64, p_f64],
    "djets_vec_scatter_async": [vp, vp],
    "djets_vec_collect_async": [vp, vp],
    "djets_vec_eval": [vp, C.POINTER(C.c_uint8), i32, p_vp, i32, p_f64, p_vp, i32, i32],
    "djets_vec_dot_s": [vp, vp, vp],
    "djets_vec_norm_s": [vp, f64, vp],
    "djets_vec_reduce_s": [vp, i32, f64, vp],
    "djets_ctx_mark": [vp, i32],
    "djets_ctx_wait_mark": [vp, i32],
    "djets_ctx_capture_begin": [vp],
    "djets_ctx_capture_end": [vp, p_vp],
    "djets_graph_launch": [vp],
    "djets_graph_destroy": [vp],
    "djets_scalar_create": [vp, p_vp],
    "djets_scalar_destroy": [vp],
    "djets_scalar_set": [vp, f64],
    "djets_scalar_get": [vp, p_f64],
    "djets_scalar_op": [vp, i32, vp, vp, i32],
    "djets_op_create": [vp, i32, i64, i64, p_i64, p_i64, i32, i32, p_i32, p_i64, p_i64, i32, p_vp],
    "djets_op_set_dense": [vp, i64, i64, vp, i64],
    "djets_op_set_diag": [vp, i64, i64, vp],
    "djets_op_finalize": [vp],
    "djets_op_replan": [vp],
    "djets_op_destroy": [vp],
    "djets_op_new_vec": [vp, i32, p_vp],
    "djets_op_mul": [vp, vp, vp],
    "djets_op_mul_adjoint": [vp, vp, vp],
    "djets_op_blockmap": [vp, i32, i32, p_i64, p_i64, p_i64, p_i64],
    "djets_op_block_owner": [vp, i64, i64, p_i32, p_i32, p_i32],
    "djets_op_getblock": [vp, i64, i64, p_i32, vp, i64],
    "djets_op_algorithmic_bytes": [vp, p_i64],
    "djets_op_perfstat_enable": [vp, i32],
    "djets_op_perfstat": [vp, i32, i32, p_f64],
    "djets_op_schedule_info": [vp, i32, p_i64, p_i64],
}

for _name, _args in SIGNATURES.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = i32
lib.djets_version.argtypes = []
lib.djets_version.restype = i32
lib.djets_last_error.argtypes = []
lib.djets_last_error.restype = C.c_char_p


def check(status: int) -> None:
    if status == OK:
        return
    msg = (lib.djets_last_error() or b"").decode("utf-8", "replace")
    if status == ERR_MISMATCH:
        raise DimensionMismatch(status, msg)
    if status == ERR_NOT_LOCAL:
        raise NotLocal(status, msg)
    raise DJetsError(status, msg)


def call(name: str, *args) -> None:
    check(getattr(lib, name)(*args))


def arr_i64(values):
    values = [int(v) for v in values]
    return (i64 * max(len(values), 1))(*values)


def arr_i32(values):
    values = [int(v) for v in values]
    return (i32 * max(len(values), 1))(*values)
