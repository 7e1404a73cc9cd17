#!/usr/bin/env python3
"""Regenerates the committed fixtures under tests/golden/ (run in the build container, where /root/reference
exists; the GPU box never reads the reference).

  poseidon_bls381_width3.json  <- manta-pay/src/crypto/poseidon/permutation_hardcoded_test/{poseidonperm_bls381_width3.sage,width3}
                                  (round constants, MDS matrix, expected output of the reference's own KAT,
                                  `hash.rs:248-258`): known-answer vector for BLS12-381 Fr arithmetic.
  bn254_vk_kat.json            <- manta-parameters/data/pay/verifying/*.dat: compressed alpha_g1 / beta_g2 / gamma_g2 /
                                  delta_g2 and the stored pairing value alpha_g1_beta_g2 (groth16.rs:338-356):
                                  known-answer vector for the ark-serialize point encoding (y-sign flags, Fq2 order).
  groth16_proofs.json          <- ORACLE-generated proof bytes on small seeded circuits (the reference holds no
                                  known-answer proofs — SURVEY.md §8c); each was checked against the trapdoor
                                  closed form when generated.  Regression vectors for the CUDA path.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"


def poseidon():
    d = os.path.join(REF, "manta-pay/src/crypto/poseidon/permutation_hardcoded_test")
    sage = open(os.path.join(d, "poseidonperm_bls381_width3.sage")).read()
    rc = re.search(r"round_constants = \[(.*?)\]", sage, re.S).group(1)
    rc = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", rc)]
    mds = re.search(r"MDS_matrix = \[(.*?)\]\]", sage, re.S).group(1)
    mds = [int(x, 16) for x in re.findall(r"0x[0-9a-fA-F]+", mds)]
    assert len(rc) == 3 * (8 + 55) and len(mds) == 9
    exp = [int(x) for x in re.findall(r'"(\d+)"', open(os.path.join(d, "width3")).read())]
    assert len(exp) == 3
    return {"source": "manta-pay/src/crypto/poseidon/permutation_hardcoded_test", "full_rounds": 8, "partial_rounds": 55,
            "input": [3, 1, 2], "round_constants": [hex(x) for x in rc], "mds": [hex(x) for x in mds],
            "expected": [str(x) for x in exp]}


def bn254_vk():
    out = {}
    for name in ("to-private", "private-transfer", "to-public"):
        data = open(os.path.join(REF, "manta-parameters/data/pay/verifying", name + ".dat"), "rb").read()
        n = int.from_bytes(data[224:232], "little")
        pos = 232 + 32 * n
        out[name] = {"file_size": len(data), "gamma_abc_len": n, "alpha_g1": data[0:32].hex(), "beta_g2": data[32:96].hex(),
                     "gamma_g2": data[96:160].hex(), "delta_g2": data[160:224].hex(),
                     "gamma_abc_g1": [data[232 + 32 * i:264 + 32 * i].hex() for i in range(n)],
                     "alpha_g1_beta_g2": data[pos:pos + 384].hex()}
    return out


def proofs():
    from helpers import oracle_keygen, trapdoor_proof_bytes, cref
    import manta_rs_b200.workload as wl
    cases = []
    for (p, w, dist, seed, r, s) in [(2, 1, "U", 1, 7, 11), (3, 40, "R", 2, 0, 5), (4, 300, "R", 3, (1 << 200) + 9, (1 << 254) + 3),
                                      (5, 1100, "U", 4, 123456789, 987654321)]:
        cs = wl.make_r1cs(p, w, seed=seed, dist=dist)
        z = wl.make_assignment(cs, seed)
        trapdoor = wl.sample_trapdoor(seed)
        pk, trap = oracle_keygen(cs, trapdoor)
        op = cref.OracleProver(pk, cs.p, cs.w, cs.a, cs.b, cs.c)
        proof = op.prove(z, r, s)
        assert proof == trapdoor_proof_bytes(cs, trap, z, r, s)
        cases.append({"p": p, "w": w, "dist": dist, "seed": seed, "r": str(r), "s": str(s), "proof": proof.hex()})
    return {"note": "oracle-generated (no reference KAT exists for proof bytes)", "cases": cases}


if __name__ == "__main__":
    for fname, fn in (("poseidon_bls381_width3.json", poseidon), ("bn254_vk_kat.json", bn254_vk), ("groth16_proofs.json", proofs)):
        with open(os.path.join(HERE, fname), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", fname)
