"""Small end-to-end run for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`:
one tiny proof (smoke), a skewed and a degenerate stand-alone MSM, a 2^11 NTT, a Poseidon batch."""
import ctypes, os, random, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge
ge.smoke()
import manta_rs_b200  # noqa: F401
from manta_rs_b200 import _native as nat, workload as wl
lib = nat.lib()
rng = random.Random(1)
n = 3000
ks = [rng.randrange(1, wl.FR_BLS12_381) for _ in range(n)]
pts = ctypes.create_string_buffer(n * 96)
nat.check(lib.mp_fixed_base_g1(0, nat.pack_scalars(ks), n, pts))
out = ctypes.create_string_buffer(96)
for sc in ([rng.randrange(wl.FR_BLS12_381) for _ in range(n)], [5] * n, [0] * n):
    nat.check(lib.mp_msm_g1(0, pts.raw, nat.pack_scalars(sc), n, out, None))
pts2 = ctypes.create_string_buffer(500 * 192)
nat.check(lib.mp_fixed_base_g2(0, nat.pack_scalars(ks[:500]), 500, pts2))
out2 = ctypes.create_string_buffer(192)
nat.check(lib.mp_msm_g2(0, pts2.raw, nat.pack_scalars([rng.randrange(wl.FR_BLS12_381) for _ in range(500)]), 500, out2, None))
buf = ctypes.create_string_buffer(nat.pack_scalars([rng.randrange(wl.FR_BLS12_381) for _ in range(1 << 11)]), (1 << 11) * 32)
nat.check(lib.mp_ntt(0, buf, 11, 0, 1, None))
print("sanitize smoke done")
