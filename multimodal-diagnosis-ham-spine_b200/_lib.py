"""ctypes binding of csrc/libmdhs_b200.so (the C ABI declared in include/mdhs_b200.h).

There is deliberately no fallback: if the library is missing or a kernel reports an error the
call raises, so a GPU box can never silently run on another code path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmdhs_b200.so")
_lib = None


class MdhsError(RuntimeError):
    pass


class GemmArgs(ctypes.Structure):
    _fields_ = [
        ("A", ctypes.c_void_p), ("lda", ctypes.c_int64), ("a_mn_major", ctypes.c_int32),
        ("B", ctypes.c_void_p), ("ldb", ctypes.c_int64), ("b_mn_major", ctypes.c_int32),
        ("D", ctypes.c_void_p), ("ldd", ctypes.c_int64), ("d_dtype", ctypes.c_int32),
        ("accumulate", ctypes.c_int32),
        ("M", ctypes.c_int32), ("N", ctypes.c_int32), ("K", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("aux_out", ctypes.c_void_p), ("ld_aux_out", ctypes.c_int64),
        ("aux_in", ctypes.c_void_p), ("ld_aux_in", ctypes.c_int64),
        ("act", ctypes.c_int32), ("dact", ctypes.c_int32),
        ("residual", ctypes.c_void_p), ("ldr", ctypes.c_int64), ("r_dtype", ctypes.c_int32),
        ("split_k", ctypes.c_int32), ("bn_hint", ctypes.c_int32),
        ("colsum", ctypes.c_void_p), ("colsumsq", ctypes.c_void_p),
    ]


def lib():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MdhsError(
                f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU or eager fallback for the hot path)")
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.mdhs_abi_version.restype = ctypes.c_int
        _lib.mdhs_launch_count.restype = ctypes.c_int64
    return _lib


def check(rc, what):
    if rc != 0:
        raise MdhsError(f"{what} failed with status {rc}")


def launch_count():
    return int(lib().mdhs_launch_count())
