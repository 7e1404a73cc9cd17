import ctypes, sys, random, time
sys.path.insert(0,'/root/repo')
lib = ctypes.CDLL('/root/repo/manta-rs_b200/libmantaprover.so')
lib.mp_strerror.restype = ctypes.c_char_p
lib.mp_last_error_detail.restype = ctypes.c_char_p
V=ctypes.c_void_p
lib.mp_debug_field_op.argtypes=[ctypes.c_int,ctypes.c_int,ctypes.c_int,V,V,V,ctypes.c_size_t]
lib.mp_debug_group_op.argtypes=[ctypes.c_int,ctypes.c_int,ctypes.c_int,V,V,V,V,ctypes.c_size_t]
lib.mp_debug_int_pipe_rate.argtypes=[ctypes.c_int,V,V]
from oracle.pyref.fields import BLS12_381 as C
from oracle.pyref.curves import Group
def pack(vals, limbs):
    arr=(ctypes.c_uint64*(len(vals)*limbs))()
    for i,v in enumerate(vals):
        for j in range(limbs): arr[i*limbs+j]=(v>>(64*j))&(2**64-1)
    return arr
def unpack(arr, n, limbs):
    return [sum(arr[i*limbs+j]<<(64*j) for j in range(limbs)) for i in range(n)]
rng=random.Random(1)
for field,(p,limbs) in enumerate([(C.q,6),(C.r,4)]):
    n=1000
    a=[rng.randrange(p) for _ in range(n)]; b=[rng.randrange(p) for _ in range(n)]
    a[0]=0; b[1]=0; a[2]=p-1; b[2]=p-1; a[3]=1
    for op,fn in enumerate([lambda x,y:(x+y)%p, lambda x,y:(x-y)%p, lambda x,y:x*y%p, lambda x,y:x*x%p, lambda x,y:pow(x,-1,p) if x else 0, lambda x,y:(-x)%p]):
        out=(ctypes.c_uint64*(n*limbs))()
        rc=lib.mp_debug_field_op(0,field,op,pack(a,limbs),pack(b,limbs),out,n)
        assert rc==0,(rc,lib.mp_strerror(rc),lib.mp_last_error_detail())
        got=unpack(out,n,limbs); exp=[fn(x,y) for x,y in zip(a,b)]
        bad=[i for i in range(n) if got[i]!=exp[i]]
        print("field",field,"op",op,"mismatches",len(bad), bad[:5])
G1=Group(C,1); G2=Group(C,2)
for gid,G in ((1,G1),(2,G2)):
    n=40
    pts=[G.mul(G.gen, rng.randrange(1,C.r)) for _ in range(n)]
    pts2=[G.mul(G.gen, rng.randrange(1,C.r)) for _ in range(n)]
    pts2[0]=pts[0]; pts2[1]=G.neg(pts[1]); pts[2]=None; pts2[3]=None; pts[4]=None; pts2[4]=None
    ks=[rng.randrange(C.r) for _ in range(n)]; ks[5]=0; ks[6]=1; ks[7]=C.r-1
    pb=96*gid
    A=b''.join(G.serialize_uncompressed(P) for P in pts); B=b''.join(G.serialize_uncompressed(P) for P in pts2)
    for op in range(3):
        out=ctypes.create_string_buffer(n*pb)
        rc=lib.mp_debug_group_op(0,gid,op,A,B,pack(ks,4),out,n)
        assert rc==0,(rc,lib.mp_strerror(rc),lib.mp_last_error_detail())
        got=[G.deserialize_uncompressed(out.raw[i*pb:(i+1)*pb]) for i in range(n)]
        if op==0: exp=[G.to_affine(G.jac_add(G.to_jac(x),G.to_jac(y))) for x,y in zip(pts,pts2)]
        elif op==1: exp=[G.to_affine(G.jac_double(G.to_jac(x))) for x in pts]
        else: exp=[G.mul(x,k) for x,k in zip(pts,ks)]
        bad=[i for i in range(n) if got[i]!=exp[i]]
        print("group",gid,"op",op,"mismatches",len(bad),bad[:8])
w=ctypes.c_double(); f=ctypes.c_double()
rc=lib.mp_debug_int_pipe_rate(0,ctypes.byref(w),ctypes.byref(f)); print(rc, "wide MAC/s %.4g  Fq mul/s %.4g  ratio %.1f"%(w.value,f.value,w.value/f.value))
