"""Extra GPU measurements: single-proof latency per shape, stand-alone MSM sweep (BASELINE config 3)."""
import ctypes, json, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manta_rs_b200
from manta_rs_b200 import workload as wl, keygen, groth16 as g16, _native as nat
lib = nat.lib()
out = {}
if "--latency" in sys.argv:
    for shape in ("to_private", "to_public", "private_transfer"):
        cs = wl.make_shape(shape)
        pk, trap = keygen.generate(cs, wl.sample_trapdoor(21))
        ctx = g16.ProvingContext.decode(pk)
        z = wl.make_assignment(cs, 0)
        comp = g16.R1CS.from_workload(cs, z)
        h = ctx.native(comp.matrices)
        batch = ctypes.c_void_p()
        nat.check(lib.mp_batch_create(h, 1, ctypes.byref(batch)))
        zb = nat.pack_scalars(z)
        rr_ = random.Random(99)   # full-width r, s: the two 255-bit scalar multiplications of the finishing kernel are part of the latency
        rb = nat.pack_scalars([rr_.randrange(wl.FR_BLS12_381)]); sb = nat.pack_scalars([rr_.randrange(wl.FR_BLS12_381)])
        res = []
        for it in range(6):
            t0 = time.perf_counter()
            nat.check(lib.mp_batch_upload(batch, 1, zb, rb, sb))
            ms = ctypes.c_float(); nat.check(lib.mp_batch_run(batch, ctypes.byref(ms)))
            o = ctypes.create_string_buffer(192); nat.check(lib.mp_batch_download(batch, o))
            res.append(((time.perf_counter() - t0) * 1e3, ms.value))
        buf = (ctypes.c_float * 8)(); lib.mp_batch_phase_ms(batch, buf, 8)
        out["latency_" + shape] = {"e2e_ms_min": min(r[0] for r in res[2:]), "device_ms_min": min(r[1] for r in res[2:]),
                                  "phases": {lib.mp_phase_name(i).decode(): round(buf[i], 3) for i in range(8)}}
        print(shape, out["latency_" + shape], flush=True)
        lib.mp_batch_destroy(batch); ctx.close()
if "--poseidon" in sys.argv:
    from manta_rs_b200 import poseidon
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "poseidon_bls381_width3.json")))
    perm = poseidon.Permutation(3, 8, 55, [int(x, 16) for x in gold["round_constants"]], [int(x, 16) for x in gold["mds"]])
    rr = random.Random(3)
    n = 1 << 18
    states = [[rr.randrange(wl.FR_BLS12_381) for _ in range(3)] for _ in range(n)]
    perm.permute_many(states[:1024])
    perm.permute_many(states)
    out["poseidon_width3"] = {"count": n, "device_ms": perm.last_device_ms, "Mperm_per_s": n / perm.last_device_ms / 1e3}
    print("poseidon width 3:", out["poseidon_width3"], flush=True)
if "--sweep" in sys.argv:
    from oracle import cref
    rng = random.Random(1)
    sizes = [int(x) for x in os.environ.get('SWEEP', '16,18,20,22,24').split(',')]
    for logn in sizes:
        n = 1 << logn
        base_k = [rng.randrange(1, wl.FR_BLS12_381) for _ in range(1 << 12)]
        pts = ctypes.create_string_buffer((1 << 12) * 96)
        nat.check(lib.mp_fixed_base_g1(0, nat.pack_scalars(base_k), 1 << 12, pts))
        reps = n >> 12
        bases = pts.raw * reps                       # repeated bases: closed form still holds
        t0 = time.time()
        rr = random.Random(logn)
        sc_bytes = b''.join(rr.randbytes(1 << 20) for _ in range(n * 32 >> 20)) if n * 32 >= (1 << 20) else rr.randbytes(n * 32)
        sc_arr = bytearray(sc_bytes)
        for i in range(n):                            # clear top bits so that scalars < r
            sc_arr[32 * i + 31] &= 0x3F
        sc_bytes = bytes(sc_arr)
        ms = ctypes.c_float(); res = ctypes.create_string_buffer(96)
        best = 1e9
        for it in range(3):
            nat.check(lib.mp_msm_g1(0, bases, sc_bytes, n, res, ctypes.byref(ms))); best = min(best, ms.value)
        # closed form: sum_j k_{j mod 4096} * s_j
        acc = [0] * (1 << 12)
        for i in range(n):
            acc[i & 4095] += int.from_bytes(sc_bytes[32 * i:32 * i + 32], "little")
        tot = sum(k * a for k, a in zip(base_k, acc)) % wl.FR_BLS12_381
        ok = res.raw == cref.fixed_base(1, [tot])
        c_ref = (logn * 69 // 100) + 2
        credited = n * ((255 + c_ref - 1) // c_ref) * 11
        cpu_ms = None
        if logn <= 20 and "--sweep-cpu" in sys.argv:   # BASELINE configs[2]: the same MSM on the host cores (CPU oracle, all threads)
            sc_int = [int.from_bytes(sc_bytes[32 * i:32 * i + 32], "little") for i in range(n)]
            tc = time.perf_counter()
            cpu_res = cref.msm(1, bases, sc_int, threads=cref.lib().oracle_max_threads())
            cpu_ms = (time.perf_counter() - tc) * 1e3
            ok = ok and cpu_res == res.raw
        out["msm_g1_2^%d" % logn] = {"device_ms": best, "ok": ok, "credited_GFqmul_per_s": credited / best / 1e6, "cpu_oracle_ms": cpu_ms}
        print("msm 2^%d: %.2f ms ok=%s credited %.1f GFq-mul/s%s (prep %.0fs)" % (
            logn, best, ok, credited / best / 1e6, "" if cpu_ms is None else ", CPU oracle (all threads) %.0f ms = %.0fx" % (cpu_ms, cpu_ms / best),
            time.time() - t0), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "probe_extra.json"), "w"), indent=1)
