"""ctypes binding of libminidiff_b200.so (the C ABI declared in include/minidiff_b200.h).

There is no CPU path: importing this module fails loudly if the shared library is missing, and the
first device call fails loudly (RuntimeError) if no B200 is present.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libminidiff_b200.so")
MAX_DIMS = 8

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python scripts/build_lib.py` "
        "(or __graft_entry__.build()). minidiff_b200 has no CPU fallback."
    )
lib = C.CDLL(LIB_PATH)


class MdbArray(C.Structure):
    """mirror of `mdb_array` (include/minidiff_b200.h)"""

    _fields_ = [
        ("ptr", C.c_void_p),
        ("dtype", C.c_int32),
        ("ndim", C.c_int32),
        ("shape", C.c_int64 * MAX_DIMS),
        ("strides", C.c_int64 * MAX_DIMS),
        ("imm", C.c_double),
        ("imm_i", C.c_int64),
    ]


# status codes
OK, EINVAL, ECUDA, ENOMEM, ENOTSUP, ECOMM, EINDEX = range(7)

# dtype codes (mdb_dtype)
BOOL, U8, I8, I16, I32, I64, F32, F64, U16, U32, U64, F16 = range(12)

# elementwise op ids (mdb_op)
OP = dict(
    COPY=0, NEG=1, ABS=2, SIGN=3, CEIL=4, FLOOR=5, SIN=6, COS=7, TAN=8, SINH=9, COSH=10, TANH=11,
    EXP=12, LOG=13, SQRT=14, RECIP=15, SQUARE=16, LOGICAL_NOT=17, INVERT=18, ISNAN=19, RELU=20,
    ADD=32, SUB=33, MUL=34, DIV=35, POW=36, MOD=37, FLOORDIV=38, MAXIMUM=39, MINIMUM=40,
    EQ=41, NE=42, GT=43, GE=44, LT=45, LE=46, AND=47, OR=48, XOR=49,
    WHERE=64, CLIP=65, FMA=66,
    SIN_BWD=96, COS_BWD=97, EXP_BWD=98, LOG_BWD=99, TANH_BWD=100, POW_BWD=101, DIV_BWD_Y=102,
    RELU_MASK_BWD=103, POW_BWD_LIN=104,
)
RED = dict(SUM=0, MEAN=1, MAX=2, MIN=3, PROD=4, ANY=5, ALL=6, ARGMAX=7, ARGMIN=8)

_P = C.POINTER
_A = _P(MdbArray)
_SIGS = {
    "mdb_abi_version": (C.c_int, []),
    "mdb_last_error": (C.c_char_p, []),
    "mdb_device_count": (C.c_int, [_P(C.c_int)]),
    "mdb_init": (C.c_int, [C.c_int]),
    "mdb_shutdown": (C.c_int, []),
    "mdb_device_info": (C.c_int, [_P(C.c_int), _P(C.c_size_t), _P(C.c_int), _P(C.c_int)]),
    "mdb_stream": (C.c_void_p, []),
    "mdb_sync": (C.c_int, []),
    "mdb_alloc": (C.c_int, [C.c_size_t, _P(C.c_void_p)]),
    "mdb_free": (C.c_int, [C.c_void_p]),
    "mdb_empty_cache": (C.c_int, []),
    "mdb_mem_stats": (C.c_int, [_P(C.c_size_t), _P(C.c_size_t), _P(C.c_size_t), _P(C.c_uint64)]),
    "mdb_host_alloc": (C.c_int, [C.c_size_t, _P(C.c_void_p)]),
    "mdb_host_free": (C.c_int, [C.c_void_p]),
    "mdb_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mdb_d2h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mdb_d2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mdb_prefetch_h2d": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mdb_prefetch_wait": (C.c_int, []),
    "mdb_event_create": (C.c_int, [_P(C.c_void_p)]),
    "mdb_event_record": (C.c_int, [C.c_void_p]),
    "mdb_event_elapsed_ms": (C.c_int, [C.c_void_p, C.c_void_p, _P(C.c_float)]),
    "mdb_event_destroy": (C.c_int, [C.c_void_p]),
    "mdb_launch_count": (C.c_uint64, []),
    "mdb_graph_begin": (C.c_int, []),
    "mdb_graph_end": (C.c_int, [_P(C.c_void_p)]),
    "mdb_graph_launch": (C.c_int, [C.c_void_p]),
    "mdb_graph_info": (C.c_int, [C.c_void_p, _P(C.c_uint64), _P(C.c_size_t)]),
    "mdb_graph_destroy": (C.c_int, [C.c_void_p]),
    "mdb_prof_enable": (C.c_int, [C.c_int]),
    "mdb_prof_read": (C.c_int, [C.c_int, _P(C.c_double), _P(C.c_uint64), _P(C.c_double)]),
    "mdb_fill": (C.c_int, [_A, C.c_double]),
    "mdb_copy": (C.c_int, [_A, _A]),
    "mdb_elementwise": (C.c_int, [C.c_int, _A, C.c_int, _A]),
    "mdb_elementwise_new": (C.c_int, [C.c_int, _A, C.c_int, _A, _A, _A]),
    "mdb_reduce": (C.c_int, [C.c_int, _A, _A, C.c_uint32]),
    "mdb_elementwise_reduce": (C.c_int, [C.c_int, _A, C.c_int, _A, C.c_int]),
    "mdb_gemm": (C.c_int, [_A, _A, _A, C.c_int]),
    "mdb_gemm_batched": (C.c_int, [_A, _A, _A]),
    "mdb_gemm_config": (C.c_int, [C.c_int]),
    "mdb_gemm_tune": (C.c_int, [C.c_int]),
    "mdb_gemm_stats": (C.c_int, [_P(C.c_uint64), C.c_int]),
    "mdb_gemm_knob": (C.c_int, [C.c_int, C.c_int]),
    "mdb_gemm_last_plan": (C.c_int, [_P(C.c_int)]),
    "mdb_gemm_fused": (C.c_int, [_A, _A, _A, C.c_int, _A, C.c_int, _A]),
    "mdb_gather_rows": (C.c_int, [_A, _A, _A]),
    "mdb_scatter_rows": (C.c_int, [_A, _A, _A, C.c_int]),
    "mdb_random": (C.c_int, [_A, C.c_int, C.c_uint64, C.c_uint64]),
    "mdb_random_reset": (C.c_int, [C.c_uint64]),
    "mdb_random_bits": (C.c_int, [_A, C.c_uint64, C.c_uint64]),
    "mdb_randint": (C.c_int, [_A, C.c_int64, C.c_int64, C.c_uint64, C.c_uint64]),
    "mdb_binomial": (C.c_int, [_A, C.c_int64, _A, C.c_uint64, C.c_uint64]),
    "mdb_permutation": (C.c_int, [_A, _A]),
    "mdb_arange": (C.c_int, [_A, C.c_double, C.c_double, C.c_int64, C.c_int64, C.c_int]),
    "mdb_cumsum_f64": (C.c_int, [_A, _A]),
    "mdb_searchsorted_cdf": (C.c_int, [_A, _A, _A]),
    "mdb_index_offsets": (C.c_int, [_A, _A, C.c_int64, C.c_int64, C.c_int]),
    "mdb_nonzero": (C.c_int, [_A, _A, _P(C.c_int64)]),
    "mdb_unravel_index": (C.c_int, [_A, _A, C.c_int, _P(C.c_int64)]),
    "mdb_isin": (C.c_int, [_A, _A, _A, C.c_int]),
    "mdb_comm_unique_id": (C.c_int, [C.c_void_p, C.c_char_p]),
    "mdb_comm_init": (C.c_int, [C.c_int, C.c_int, C.c_void_p, C.c_char_p]),
    "mdb_comm_allreduce_f32": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int]),
    "mdb_comm_allreduce_multi_f32": (C.c_int, [_P(C.c_void_p), _P(C.c_size_t), C.c_int, C.c_int]),
    "mdb_comm_wait": (C.c_int, []),
    "mdb_comm_last_seq": (C.c_uint64, []),
    "mdb_comm_wait_seq": (C.c_int, [C.c_uint64]),
    "mdb_comm_destroy": (C.c_int, []),
}
EXPORTS = tuple(_SIGS)
for _name, (_res, _args) in _SIGS.items():
    _fn = getattr(lib, _name)  # AttributeError here == the .so does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args

if lib.mdb_abi_version() != 1:
    raise ImportError("libminidiff_b200.so ABI version mismatch; rebuild the library")


def last_error() -> str:
    return (lib.mdb_last_error() or b"").decode("utf-8", "replace")


def check(rc: int) -> None:
    """Map a C status to the exception type NumPy would raise through the reference
    (shape / broadcast / argument problems -> ValueError, everything else -> RuntimeError)."""
    if rc == OK:
        return
    msg = last_error()
    if rc == EINVAL:
        raise ValueError(msg)
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == ENOTSUP:
        raise NotImplementedError(msg)
    if rc == EINDEX:
        raise IndexError(msg)
    raise RuntimeError(f"minidiff_b200: {msg}")


_initialised = False


def ensure_device() -> None:
    """Bind this process to its GPU (LOCAL_RANK under torchrun) -- raises if there is none."""
    global _initialised
    if _initialised:
        return
    dev = int(os.environ.get("MINIDIFF_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    check(lib.mdb_init(dev))
    _initialised = True


def device_available() -> bool:
    n = C.c_int(0)
    return lib.mdb_device_count(C.byref(n)) == OK and n.value > 0
