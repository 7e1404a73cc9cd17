import sys, time, random
sys.path.insert(0,'/root/repo')
import manta_rs_b200
from manta_rs_b200 import workload as wl, keygen, groth16 as g16, _native as nat
from oracle.pyref.fields import BLS12_381 as C
from oracle.pyref import groth16 as og
from oracle.pyref.curves import Group
G1=Group(C,1); G2=Group(C,2)
def run(p,w,dist,nproofs,full_oracle):
    cs=wl.make_r1cs(p,w,dist=dist)
    t=time.time(); pkb,trap=keygen.generate(cs, wl.sample_trapdoor(3)); t_key=time.time()-t
    ctx=g16.ProvingContext.decode(pkb)
    zs=[wl.make_assignment(cs,s) for s in range(nproofs)]
    rng=random.Random(5)
    rs=[rng.randrange(C.r) for _ in zs]; ss=[rng.randrange(C.r) for _ in zs]
    if nproofs>1: rs[1]=0
    t=time.time(); proofs=g16.Groth16.prove_many_with_randomness(ctx,[g16.R1CS.from_workload(cs,z) for z in zs],rs,ss); t_prove=time.time()-t
    ok=True
    for i,(z,r,s,pr) in enumerate(zip(zs,rs,ss,proofs)):
        a_s,b_s,c_s=keygen.trapdoor_proof_scalars(cs,trap,z,r,s)
        exp=og.proof_to_bytes(C,(G1.mul(C.g1,a_s),G2.mul(C.g2,b_s),G1.mul(C.g1,c_s)))
        if exp!=pr.to_bytes(): ok=False; print("  MISMATCH proof",i, [exp[k:k+48]==pr.to_bytes()[k:k+48] for k in (0,48,96,144)])
    if full_oracle:
        pk=og.pk_from_bytes(C,pkb)
        pr=og.create_proof(C,pk,cs.as_dict(),zs[0],rs[0],ss[0])
        if og.proof_to_bytes(C,pr)!=proofs[0].to_bytes(): ok=False; print("  MISMATCH vs oracle create_proof")
    print("shape p=%d w=%d m=%d dist=%s proofs=%d ok=%s keygen=%.2fs prove=%.3fs"%(p,w,cs.m,dist,nproofs,ok,t_key,t_prove))
    ctx.close()
    return ok
allok=True
allok&=run(2,5,"U",3,True)
allok&=run(3,60,"R",4,True)
allok&=run(5,1200,"R",3,False)
allok&=run(13,8240,"U",2,False)
print("ALL OK" if allok else "FAILED")
