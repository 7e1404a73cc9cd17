"""The C++ host mirror of the reference interface (include/mantaprover.hpp) against the CPU oracle.

The test writes a fixture (proving key, matrices, assignments, rng seeds, expected proof bytes from the oracle with the
r, s that the Python mirror of `create_random_proof` draws from the same seeds), builds tests/cpp/host_mirror_test.cpp with
g++ against the in-tree libmantaprover.so and runs it.  CPU: the binary must report the opaque Error (no device, no CPU
fallback).  GPU: every proof bit-exact."""
import os
import struct
import subprocess

import pytest

from helpers import ROOT, cref, oracle_keygen
import manta_rs_b200.workload as wl
from manta_rs_b200.rng import ChaCha20Rng, field_rand


def _matrix(rows):
    cols, coeffs, row_ptr = [], [], [0]
    for row in rows:
        for coeff, col in row:
            cols.append(col)
            coeffs.append(coeff)
        row_ptr.append(len(cols))
    return (struct.pack("<Q", len(cols)) + struct.pack(f"<{len(row_ptr)}Q", *row_ptr) + struct.pack(f"<{len(cols)}I", *cols)
            + b"".join(c.to_bytes(32, "little") for c in coeffs))


def _build(tmp_path):
    exe = str(tmp_path / "host_mirror_test")
    libdir = os.path.join(ROOT, "manta-rs_b200")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"),
                    "-o", exe, "-L", libdir, "-lmantaprover", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    return exe


def _fixture(tmp_path):
    cs = wl.make_r1cs(3, 70, dist="R")
    pk, _ = oracle_keygen(cs, wl.sample_trapdoor(12))
    op = cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c)
    count = 3
    zs = [wl.make_assignment(cs, 40 + i) for i in range(count)]
    seeds = [bytes([7 * i + j & 0xFF for j in range(32)]) for i in range(count)]
    singles = []
    for z, seed in zip(zs, seeds):
        rng = ChaCha20Rng(seed)
        r = field_rand(rng, cs.modulus)
        s = field_rand(rng, cs.modulus)
        singles.append(op.prove(z, r, s))
    rng = ChaCha20Rng(seeds[0])
    many = []
    for z in zs:
        r = field_rand(rng, cs.modulus)
        s = field_rand(rng, cs.modulus)
        many.append(op.prove(z, r, s))
    blob = struct.pack("<Q", len(pk)) + pk + struct.pack("<QQQ", cs.p, cs.w, len(cs.a))
    blob += _matrix(cs.a) + _matrix(cs.b) + _matrix(cs.c) + struct.pack("<Q", count)
    for z, seed, e1, e2 in zip(zs, seeds, singles, many):
        blob += b"".join(int(v).to_bytes(32, "little") for v in z) + seed + e1 + e2
    path = str(tmp_path / "fixture.bin")
    open(path, "wb").write(blob)
    return path


def test_cpp_host_mirror_builds_and_fails_loudly_without_device(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu test")
    exe = _build(tmp_path)
    r = subprocess.run([exe, _fixture(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 3 and "Error" in r.stdout, (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_cpp_host_mirror_proofs_bit_exact(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, _fixture(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "bit-exact" in r.stdout, (r.returncode, r.stdout, r.stderr)
