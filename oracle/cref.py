"""TEST INFRASTRUCTURE ONLY — ctypes loader for the C++ oracle (oracle/csrc/oracle.cpp -> oracle/_build/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def _host_tag() -> str:
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def build(force=False):
    """(Re)build with -march=native; rebuilt when the source changed or the host CPU differs from the one the
    library was built on (the built .so travels to the GPU box with the repo snapshot)."""
    stamp = os.path.join(_HERE, "_build", "host.txt")
    tag = _host_tag()
    stale = (not os.path.exists(LIB_PATH)
             or os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "csrc", "oracle.cpp"))
             or not os.path.exists(stamp) or open(stamp).read() != tag)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
        with open(stamp, "w") as f:
            f.write(tag)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        l = ctypes.CDLL(LIB_PATH)
        V, I, SZ = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
        l.oracle_max_threads.restype = I
        l.oracle_ctx_create.restype = V
        l.oracle_ctx_create.argtypes = [V, SZ, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, V, V, V]
        l.oracle_ctx_destroy.argtypes = [V]
        l.oracle_ctx_domain_size.restype = ctypes.c_uint64
        l.oracle_ctx_domain_size.argtypes = [V]
        l.oracle_prove.argtypes = [V, V, V, V, V, I]
        l.oracle_witness_map.argtypes = [V, V, V]
        l.oracle_msm_g1.argtypes = [V, V, SZ, V, I]
        l.oracle_msm_g2.argtypes = [V, V, SZ, V, I]
        l.oracle_fixed_base.argtypes = [I, V, SZ, V]
        l.oracle_ntt.argtypes = [V, ctypes.c_uint, I, I]
        l.oracle_field_op.argtypes = [I, I, V, V, V, SZ]
        _lib = l
    return _lib


def pack(vals, limbs=4) -> bytes:
    return b"".join(int(v).to_bytes(8 * limbs, "little") for v in vals)


def unpack(buf, limbs=4):
    n = 8 * limbs
    return [int.from_bytes(buf[i:i + n], "little") for i in range(0, len(buf), n)]


def _csr(rows):
    row_ptr, cols, coeffs = [0], [], []
    for row in rows:
        for coeff, col in row:
            cols.append(col)
            coeffs.append(coeff)
        row_ptr.append(len(cols))
    return ((ctypes.c_uint64 * len(row_ptr))(*row_ptr), (ctypes.c_uint32 * max(1, len(cols)))(*cols),
            ctypes.create_string_buffer(pack(coeffs), max(1, len(coeffs)) * 32))


class OracleProver:
    """CPU restatement of `ark_groth16::create_proof` bound to one proving key + circuit."""

    def __init__(self, pk_bytes: bytes, p: int, w: int, a, b, c):
        l = lib()
        self._keep = [_csr(m) for m in (a, b, c)]
        rp = (ctypes.c_void_p * 3)(*[ctypes.cast(k[0], ctypes.c_void_p) for k in self._keep])
        cl = (ctypes.c_void_p * 3)(*[ctypes.cast(k[1], ctypes.c_void_p) for k in self._keep])
        cf = (ctypes.c_void_p * 3)(*[ctypes.cast(k[2], ctypes.c_void_p) for k in self._keep])
        self.n = p + w
        self.h = l.oracle_ctx_create(pk_bytes, len(pk_bytes), p, w, len(a), rp, cl, cf)
        if not self.h:
            raise ValueError("oracle: malformed proving key / shape mismatch")
        self.m = l.oracle_ctx_domain_size(self.h)

    def prove(self, z, r, s, threads=1) -> bytes:
        zb = z if isinstance(z, (bytes, bytearray)) else pack(z)
        out = ctypes.create_string_buffer(192)
        rc = lib().oracle_prove(self.h, zb, pack([r]), pack([s]), out, threads)
        assert rc == 0
        return out.raw

    def witness_map(self, z):
        out = ctypes.create_string_buffer(self.m * 32)
        lib().oracle_witness_map(self.h, pack(z), out)
        return unpack(out.raw)

    def close(self):
        if self.h:
            lib().oracle_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def msm(group: int, bases_bytes: bytes, scalars, threads=1) -> bytes:
    pb = 96 if group == 1 else 192
    n = len(scalars)
    out = ctypes.create_string_buffer(pb)
    fn = lib().oracle_msm_g1 if group == 1 else lib().oracle_msm_g2
    fn(bases_bytes, pack(scalars), n, out, threads)
    return out.raw


def fixed_base(group: int, scalars) -> bytes:
    pb = 96 if group == 1 else 192
    out = ctypes.create_string_buffer(max(1, len(scalars)) * pb)
    lib().oracle_fixed_base(group, pack(scalars), len(scalars), out)
    return out.raw[: len(scalars) * pb]


def ntt(values, log_n, inverse, coset):
    buf = ctypes.create_string_buffer(pack(values), len(values) * 32)
    lib().oracle_ntt(buf, log_n, int(inverse), int(coset))
    return unpack(buf.raw)


def field_op(field: int, op: int, a, b=None):
    limbs = 6 if field == 0 else 4
    out = ctypes.create_string_buffer(len(a) * limbs * 8)
    lib().oracle_field_op(field, op, pack(a, limbs), pack(b, limbs) if b is not None else None, out, len(a))
    return unpack(out.raw, limbs)
