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
# round 2 additions: a batch large enough for the tree reduction path (> 8 proofs), resident-bases MSM, device keygen,
# trusted-setup initialize on the reference's dummy circuit, group-valued transform
sys.path.insert(0, os.path.join(ROOT, "tests"))
from manta_rs_b200 import groth16 as g16, keygen
from test_keygen import dummy_circuit, phase1_powers
cs = wl.make_r1cs(3, 40, dist="R")
pk = keygen.generate(cs, wl.sample_trapdoor(9))
ctx = g16.ProvingContext.decode(pk)
zs = [wl.make_assignment(cs, s) for s in range(10)]
proofs = g16.Groth16.prove_many_with_randomness(ctx, [g16.R1CS.from_workload(cs, z) for z in zs], list(range(1, 11)), list(range(11, 21)))
assert len(proofs) == 10
ctx.close()
h = ctypes.c_void_p()
nat.check(lib.mp_msm_bases_create(0, 1, pts.raw[:96 * 200], 200, ctypes.byref(h)))
nat.check(lib.mp_msm_bases_run(h, nat.pack_scalars([rng.randrange(wl.FR_BLS12_381) for _ in range(150)]), 150, out, None))
lib.mp_msm_bases_destroy(h)
dcs, _ = dummy_circuit()
keygen.mpc_initialize(dcs, *phase1_powers(dcs.m, 0x1234567, 0x89ABCDE, 0xF0F0F0F1))
keygen.group_ntt(2, pts2.raw[:192 * 8], True)
print("sanitize smoke (round 2 additions) done")
