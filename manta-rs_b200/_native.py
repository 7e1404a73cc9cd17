"""ctypes binding of libmantaprover.so (the C ABI declared in include/mantaprover.h).

The library is the product path: if it is missing or cannot be loaded this module raises — there is no
Python / CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MP_LIB_PATH") or os.path.join(_HERE, "libmantaprover.so")   # MP_LIB_PATH: A/B builds of the same library

FR_LIMBS = 4
G1_BYTES, G2_BYTES, PROOF_BYTES = 96, 192, 192


class PkView(ctypes.Structure):
    _fields_ = [
        ("alpha_g1", ctypes.c_void_p), ("beta_g2", ctypes.c_void_p), ("gamma_g2", ctypes.c_void_p),
        ("delta_g2", ctypes.c_void_p), ("gamma_abc_g1", ctypes.c_void_p), ("gamma_abc_len", ctypes.c_uint64),
        ("beta_g1", ctypes.c_void_p), ("delta_g1", ctypes.c_void_p),
        ("a_query", ctypes.c_void_p), ("a_len", ctypes.c_uint64),
        ("b_g1_query", ctypes.c_void_p), ("b_g1_len", ctypes.c_uint64),
        ("b_g2_query", ctypes.c_void_p), ("b_g2_len", ctypes.c_uint64),
        ("h_query", ctypes.c_void_p), ("h_len", ctypes.c_uint64),
        ("l_query", ctypes.c_void_p), ("l_len", ctypes.c_uint64),
    ]


class R1csView(ctypes.Structure):
    _fields_ = [
        ("num_instance", ctypes.c_uint64), ("num_witness", ctypes.c_uint64), ("num_constraints", ctypes.c_uint64),
        ("a_row_ptr", ctypes.c_void_p), ("a_col", ctypes.c_void_p), ("a_coeff", ctypes.c_void_p),
        ("b_row_ptr", ctypes.c_void_p), ("b_col", ctypes.c_void_p), ("b_coeff", ctypes.c_void_p),
        ("c_row_ptr", ctypes.c_void_p), ("c_col", ctypes.c_void_p), ("c_coeff", ctypes.c_void_p),
    ]


# every symbol include/mantaprover.h declares: name -> (restype, argtypes)
_V, _I, _SZ, _U = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint
SYMBOLS = {
    "mp_strerror": (ctypes.c_char_p, [_I]),
    "mp_last_error_detail": (ctypes.c_char_p, []),
    "mp_device_count": (_I, [_V]),
    "mp_pk_parse": (_I, [_V, _SZ, _V]),
    "mp_ctx_create": (_I, [_V, _V, _I, _V]),
    "mp_ctx_destroy": (None, [_V]),
    "mp_ctx_info": (_I, [_V, _V, _V, _V, _V]),
    "mp_prove": (_I, [_V, _V, _V, _V, _V]),
    "mp_prove_from_abc": (_I, [_V, _V, _V, _V, _V, _V, _V, _V]),
    "mp_prove_batch": (_I, [_V, _SZ, _V, _V, _V, _V]),
    "mp_batch_create": (_I, [_V, _SZ, _V]),
    "mp_batch_create_ex": (_I, [_V, _SZ, _I, _V]),
    "mp_batch_destroy": (None, [_V]),
    "mp_batch_upload": (_I, [_V, _SZ, _V, _V, _V]),
    "mp_batch_run": (_I, [_V, _V]),
    "mp_batch_download": (_I, [_V, _V]),
    "mp_batch_run_async": (_I, [_V]),
    "mp_batch_submit": (_I, [_V, _SZ, _V, _V, _V, _V]),
    "mp_batch_wait": (_I, [_V, _V]),
    "mp_batch_phase_ms": (_I, [_V, _V, _I]),
    "mp_phase_name": (ctypes.c_char_p, [_I]),
    "mp_batch_dominant_kernel": (_I, [_V, _V, _V]),
    "mp_batch_kernel_launches": (ctypes.c_uint64, [_V]),
    "mp_batch_device_bytes": (ctypes.c_uint64, [_V]),
    "mp_batch_set_overlap": (_I, [_V, _I]),
    "mp_msm_g1": (_I, [_I, _V, _V, _SZ, _V, _V]),
    "mp_msm_g2": (_I, [_I, _V, _V, _SZ, _V, _V]),
    "mp_msm_bases_create": (_I, [_I, _I, _V, _SZ, _V]),
    "mp_msm_bases_run": (_I, [_V, _V, _SZ, _V, _V]),
    "mp_msm_bases_destroy": (None, [_V]),
    "mp_points_sum_g1": (_I, [_I, _V, _SZ, _V]),
    "mp_points_sum_g2": (_I, [_I, _V, _SZ, _V]),
    "mp_ntt": (_I, [_I, _V, _U, _I, _I, _V]),
    "mp_witness_map": (_I, [_V, _V, _V]),
    "mp_fixed_base_g1": (_I, [_I, _V, _SZ, _V]),
    "mp_fixed_base_g2": (_I, [_I, _V, _SZ, _V]),
    "mp_keygen": (_I, [_I, _V, _V, ctypes.c_uint64, _V, _SZ, _V]),
    "mp_mpc_initialize": (_I, [_I, _V, _V, _SZ, _V, _V, _V, _V, _V, _SZ, _V]),
    "mp_group_ntt": (_I, [_I, _I, _V, _U, _I]),
    "mp_poseidon_permute": (_I, [_I, _I, _I, _I, _V, _V, _V, _SZ, _V]),
    "mp_debug_field_op": (_I, [_I, _I, _I, _V, _V, _V, _SZ]),
    "mp_debug_group_op": (_I, [_I, _I, _I, _V, _V, _V, _V, _SZ]),
    "mp_debug_ba_geometry": (_I, [_I, _I, _U, _I, _V, _V, _V, _V]),
    "mp_debug_prove_ba_demand": (_I, [_U, _U, _SZ, _SZ, _I, _V]),
    "mp_debug_int_pipe_rate": (_I, [_I, _V, _V]),
}

_lib = None


class NativeError(RuntimeError):
    def __init__(self, code, what, detail):
        super().__init__(f"libmantaprover: {what} (code {code}){': ' + detail if detail else ''}")
        self.code = code


def lib():
    """The loaded library (raises if it has not been built: run `python manta-rs_b200/build.py`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing — build it with `python manta-rs_b200/build.py`; "
                               "there is no CPU fallback for the proving path")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int):
    if rc != 0:
        l = lib()
        raise NativeError(rc, l.mp_strerror(rc).decode(), l.mp_last_error_detail().decode())


def pack_scalars(vals, limbs=FR_LIMBS) -> bytes:
    n = 8 * limbs
    return b"".join(int(v).to_bytes(n, "little") for v in vals)


def unpack_scalars(buf: bytes, limbs=FR_LIMBS):
    n = 8 * limbs
    return [int.from_bytes(buf[i:i + n], "little") for i in range(0, len(buf), n)]
