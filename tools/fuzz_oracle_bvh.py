"""CPU fuzzing of the ORACLE: its BVH traversal against its own brute-force loop on the scenes / rays of tools/fuzz_emul.py.
python tools/fuzz_oracle_bvh.py SEED0 SEED1"""
import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/oracle'); sys.path.insert(0,'/root/repo/tests')
sys.argv_bak=sys.argv[:]; a0,a1=int(sys.argv[1]),int(sys.argv[2]); sys.argv=['x','0','0']
exec(open('/root/repo/tools/fuzz_emul.py').read().split("bad = 0")[0])
import oracle.oracle as om
from pgi_raytracing_b200 import scenes
INVALID=0xFFFFFFFF
bad=0
for seed in range(a0,a1):
    rng = np.random.default_rng(seed); kind = seed % 5; n = int(rng.choice([1, 2, 3, 7, 33, 200, 1500, 6000]))
    pos = gen_scene(rng, kind, n); rays = gen_rays(rng, pos, 600)
    P3 = pos.reshape(n,3,3)
    nrm = np.tile(np.array([0,0,1],np.float32),(n,3,1)); uv=np.zeros((n,3,2),np.float32)
    sc = scenes.Scene("f",[scenes.Mesh("m",P3,nrm,uv,0)],[scenes.Material("m")],camera=scenes.Camera(8,8))
    o = om.Oracle(sc)
    rh = om.make_rayhits(rays[:,:3], rays[:,4:7], tnear=0.0)
    rh["tnear"]=rays[:,3]; rh["tfar"]=rays[:,7]
    a = o.intersect(rh, brute=False); b = o.intersect(rh, brute=True)
    for f in ("tfar","u","v","geomID","primID"):
        if not np.array_equal(a[f].view(np.uint32) if a[f].dtype==np.float32 else a[f], b[f].view(np.uint32) if b[f].dtype==np.float32 else b[f]):
            idx=np.nonzero(a[f]!=b[f])[0]
            print("MISMATCH seed",seed,"kind",kind,"n",n,f,len(idx),"ray",rays[idx[0]],"bvh",a["tfar"][idx[0]],a["primID"][idx[0]],"brute",b["tfar"][idx[0]],b["primID"][idx[0]]); bad+=1; break
print("done",a0,a1,"bad",bad)
