import ctypes, sys, random, time
sys.path.insert(0,'/root/repo')
lib = ctypes.CDLL('/root/repo/manta-rs_b200/libmantaprover.so')
lib.mp_strerror.restype = ctypes.c_char_p
lib.mp_last_error_detail.restype = ctypes.c_char_p
V=ctypes.c_void_p
lib.mp_msm_g1.argtypes=[ctypes.c_int,V,V,ctypes.c_size_t,V,V]
lib.mp_msm_g2.argtypes=[ctypes.c_int,V,V,ctypes.c_size_t,V,V]
lib.mp_fixed_base_g1.argtypes=[ctypes.c_int,V,ctypes.c_size_t,V]
lib.mp_fixed_base_g2.argtypes=[ctypes.c_int,V,ctypes.c_size_t,V]
lib.mp_ntt.argtypes=[ctypes.c_int,V,ctypes.c_uint,ctypes.c_int,ctypes.c_int,V]
from oracle.pyref.fields import BLS12_381 as C
from oracle.pyref.curves import Group, msm_pippenger
from oracle.pyref.poly import Radix2Domain
def chk(rc): assert rc==0,(rc,lib.mp_strerror(rc),lib.mp_last_error_detail())
def pack(vals, limbs=4):
    return b''.join(v.to_bytes(8*limbs,'little') for v in vals)
rng=random.Random(7)
for gid,fb,msm,pb in ((1,lib.mp_fixed_base_g1,lib.mp_msm_g1,96),(2,lib.mp_fixed_base_g2,lib.mp_msm_g2,192)):
    G=Group(C,gid)
    for n in ([0,1,2,33,500,3000] if gid==1 else [0,1,40,700]):
        ks=[rng.randrange(1,C.r) for _ in range(n)]
        out=ctypes.create_string_buffer(max(n,1)*pb)
        chk(fb(0,pack(ks),n,out))
        bases=[G.deserialize_uncompressed(out.raw[i*pb:(i+1)*pb]) for i in range(n)]
        for i in range(min(n,3)): assert bases[i]==G.mul(G.gen,ks[i]),"fixed base mismatch"
        sc=[rng.randrange(C.r) for _ in range(n)]
        if n>5: sc[0]=0; sc[1]=1; sc[2]=C.r-1; sc[3]=2**128-1; sc[4]=1<<254
        raw=bytearray(out.raw[:n*pb])
        if n>40:  # an infinity base and a duplicated base
            raw[7*pb:8*pb]=G.serialize_uncompressed(None); bases[7]=None
            raw[9*pb:10*pb]=raw[8*pb:9*pb]; bases[9]=bases[8]; sc[9]=sc[8]
        res=ctypes.create_string_buffer(pb); ms=ctypes.c_float()
        t=time.time(); chk(msm(0,bytes(raw),pack(sc),n,res,ctypes.byref(ms)))
        got=G.deserialize_uncompressed(res.raw)
        # closed form: sum k_i s_i (fast) cross-check, plus oracle pippenger on small n
        tot=0
        for i in range(n):
            if bases[i] is not None: tot+= (ks[i] if not (n>40 and i==9) else ks[8])*sc[i]
        exp=G.mul(G.gen, tot % C.r)
        ok=(got==exp)
        if n<=700: ok = ok and (G.to_affine(msm_pippenger(G,bases,sc))==got)
        print("msm G%d n=%d ok=%s dev_ms=%.3f"%(gid,n,ok,ms.value)); assert ok
for logn in [0,1,2,3,5,8,10,11,12,13,16]:
    n=1<<logn
    dom=Radix2Domain(C,n)
    x=[rng.randrange(C.r) for _ in range(n)]
    for inv in (0,1):
        for coset in (0,1):
            if logn>12 and (inv,coset) not in ((0,0),(1,1)): continue
            buf=ctypes.create_string_buffer(pack(x),n*32); ms=ctypes.c_float()
            chk(lib.mp_ntt(0,buf,logn,inv,coset,ctypes.byref(ms)))
            got=[int.from_bytes(buf.raw[i*32:(i+1)*32],'little') for i in range(n)]
            exp={(0,0):dom.fft,(1,0):dom.ifft,(0,1):dom.coset_fft,(1,1):dom.coset_ifft}[(inv,coset)](x)
            print("ntt logn=%d inv=%d coset=%d ok=%s dev_ms=%.3f"%(logn,inv,coset,got==exp,ms.value)); assert got==exp
print("ALL OK")
