"""Shared test helpers (CPU oracle side).  The oracle is the checker; the product path never imports it."""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cref, trapdoor as _trap  # noqa: E402
from oracle.pyref.fields import BLS12_381  # noqa: E402
import manta_rs_b200  # noqa: E402,F401


def oracle_keygen(cs, trapdoor, h_len=None):
    """Known-trapdoor proving key built entirely on the CPU by the oracle: QAP evaluation with Python integers
    (oracle/trapdoor.py) and fixed-base multiplications by the C++ oracle.  Returns (`ProvingContext` bytes, trapdoor dict)."""
    g1s, g2s, trap = _trap.key_scalars(cs, trapdoor, h_len)
    n, p = cs.n, cs.p
    h_len = len(g1s["h"])
    order = ("alpha", "gamma_abc", "beta", "delta", "a", "b", "h", "l")
    g1 = cref.fixed_base(1, [x for k in order for x in g1s[k]])
    g2 = cref.fixed_base(2, g2s["beta"] + g2s["gamma"] + g2s["delta"] + g2s["b"])
    pos = [0]

    def take(k):
        s = g1[pos[0] * 96:(pos[0] + k) * 96]
        pos[0] += k
        return s

    def vec(data, k):
        return k.to_bytes(8, "little") + data

    alpha_g1, gamma_abc, beta_g1, delta_g1 = take(1), take(p), take(1), take(1)
    a_q, b1_q, h_q, l_q = take(n), take(n), take(h_len), take(n - p)
    pk = (alpha_g1 + g2[:192] + g2[192:384] + g2[384:576] + vec(gamma_abc, p) + beta_g1 + delta_g1 + vec(a_q, n)
          + vec(b1_q, n) + vec(g2[576:], n) + vec(h_q, h_len) + vec(l_q, n - p))
    return pk, trap


def trapdoor_proof_bytes(cs, trap, z, r_rand, s_rand) -> bytes:
    """Proof bytes from the closed form (three oracle scalar multiplications + oracle compression)."""
    key = id(trap)
    chk = _CHECKERS.get(key)
    if chk is None or chk[0] is not trap:
        chk = (trap, _trap.TrapdoorChecker(cs, trap))
        _CHECKERS[key] = chk
    return chk[1].proof_bytes(z, r_rand, s_rand)


_CHECKERS = {}
