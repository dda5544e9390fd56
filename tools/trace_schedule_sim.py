"""Offline model of wf_trace's warp scheduling (CPU only; numpy): how many of a warp's issue slots do DDA steps on live lanes?

Part 1 samples shadow / bounce rays on scene_fall.vox the way the path tracer produces them (origins on exposed faces, cosine-weighted and
sky-biased directions), steps them with a plain DDA and records, per ray, its true number of iterations and the number PREDICTED from the DDA
set-up alone (iterations to the volume boundary). Part 2 replays wf_trace's loop for one warp -- ranges of `grab` rays, chunks of `chunk`
iterations, refill when fewer than `live_min` lanes run, assumed costs for the refill / retire rounds -- and reports useful step
instructions / issued instructions x 32 lanes.

Findings: the model lands on the measured figure (0.46 vs roofline.issue 0.35 / 0.80 issue slots busy = 0.44) and on the
measured optima (chunk 16, live_min ~26); sorting the rays of a range by predicted length would give +11 % (bounce) / +6 % (shadow rays), a
perfect sort +22 %; the largest loss is the refill / retire code itself (0.46 -> 0.68 if it were free).

    python tools/trace_schedule_sim.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util
vol = util.scene_fall_volume()
res = vol['res']; grid = np.asarray(vol['grid']).reshape(res[2],res[1],res[0])  # z,y,x
occ = grid >= 0
X,Y,Z = res
print("res",res,"solid",occ.sum())
rng = np.random.RandomState(0)
# exposed faces
faces=[]
pad = np.pad(occ,1,constant_values=False)
for axis,(dz,dy,dx) in enumerate([(0,0,1),(0,0,-1),(0,1,0),(0,-1,0),(1,0,0),(-1,0,0)]):
    nb = pad[1+dz:1+dz+Z,1+dy:1+dy+Y,1+dx:1+dx+X]
    z,y,x = np.nonzero(occ & ~nb)
    n = np.array([dx,dy,dz],float)
    faces.append(np.stack([x,y,z,np.full_like(x,axis)],1))
faces=np.concatenate(faces)
normals=np.array([(1,0,0),(-1,0,0),(0,1,0),(0,-1,0),(0,0,1),(0,0,-1)],float)
# weight: top faces more visible
w = np.where(faces[:,3]==2,3.0,1.0); w/=w.sum()
N=200000
sel = faces[rng.choice(len(faces),N,p=w)]
nrm = normals[sel[:,3]]
org = sel[:,:3]+0.5+nrm*0.501 + (rng.rand(N,3)-0.5)*(1-np.abs(nrm))
def cos_dirs(n):
    u1,u2=rng.rand(len(n)),rng.rand(len(n))
    r=np.sqrt(u1); th=2*np.pi*u2
    lx,ly,lz=r*np.cos(th),r*np.sin(th),np.sqrt(1-u1)
    # basis
    a=np.where(np.abs(n[:,[0]])>0.5,[[0,1,0]],[[1,0,0]])
    t=np.cross(n,a); t/=np.linalg.norm(t,axis=1)[:,None]; b=np.cross(n,t)
    return t*lx[:,None]+b*ly[:,None]+n*lz[:,None]
def sphere_dirs(k):
    v=rng.normal(size=(k,3)); v[:,1]=np.abs(v[:,1])*np.where(rng.rand(k)<0.85,1,-1)  # env sampling favours the sky
    return v/np.linalg.norm(v,axis=1)[:,None]
def trace(o,d):
    d=np.where(np.abs(d)<1e-5,1e-5,d)
    inc=1/np.abs(d); sg=np.sign(d).astype(int)
    v=np.floor(o).astype(int)
    dis=np.where(sg>0,(v+1-o),(o-v))*inc
    steps=np.zeros(len(o),int); alive=np.ones(len(o),bool)
    # predicted exit steps
    nb = np.where(sg>0,np.array([X,Y,Z])-1-v, v)          # voxels to boundary per axis
    texit = np.min(dis+nb*inc,axis=1)
    pred = np.sum(np.floor(np.maximum(texit[:,None]-dis,0)/inc)+1,axis=1)
    for it in range(X+Y+Z+5):
        inside=(v>=0).all(1)&(v[:,0]<X)&(v[:,1]<Y)&(v[:,2]<Z)
        alive&=inside
        if not alive.any(): break
        idx=np.nonzero(alive)[0]
        hit=occ[v[idx,2],v[idx,1],v[idx,0]]
        steps[idx]+=1
        alive[idx[hit]]=False
        idx=idx[~hit]
        if len(idx)==0: break
        m=dis[idx].min(1,keepdims=True); ax=dis[idx]==m
        dis[idx]+=ax*inc[idx]; v[idx]+=ax*sg[idx]
    return steps,pred
sb,pb=trace(org,cos_dirs(nrm))
ss,ps=trace(org,sphere_dirs(N))
print("bounce mean steps",sb.mean(),"shadow",ss.mean(),"corr bounce",np.corrcoef(sb,pb)[0,1],"corr shadow",np.corrcoef(ss,ps)[0,1])
d = dict(sb=sb, pb=pb, ss=ss, ps=ps)


# ---- part 2: the warp loop ------------------------------------------------------------------------------------------
def simulate(steps, live_min=26, chunk=16, grab=256, step_instr=25, chunk_over=8, refill_instr=140, retire_instr=60):
    """one warp processing ranges of `grab` rays in order; returns (useful lane-steps, issued warp-instr, mean lanes in step loop)"""
    n=len(steps); pos=0
    lanes=np.zeros(32,int); have=np.zeros(32,bool)
    instr=0; useful=0; lane_steps_issued=0; step_chunks=0
    exhausted=False
    while True:
        need=~have
        if need.any() and not exhausted:
            k=min(need.sum(), n-pos)
            idx=np.nonzero(need)[0][:k]
            lanes[idx]=steps[pos:pos+k]; have[idx]=True; pos+=k
            instr+=refill_instr
            if pos>=n: exhausted=True
        if not have.any(): break
        while True:
            run=have&(lanes>0)
            adv=np.minimum(lanes,chunk)*run
            mx=adv.max() if run.any() else 0
            # the unrolled chunk: lanes break out when done; the warp issues until the slowest lane of the chunk ends
            instr+=mx*step_instr+chunk_over
            useful+=adv.sum(); lanes-=adv
            lane_steps_issued+=mx*32; step_chunks+=1
            run=have&(lanes>0)
            if exhausted:
                if not run.any(): break
            elif run.sum()<live_min: break
        fin=have&(lanes<=0)
        if fin.any(): instr+=retire_instr; have[fin]=False
    return useful, instr, useful/max(1,lane_steps_issued)*32
def run(steps, pred=None, sort=False, **kw):
    rng=np.random.RandomState(1)
    tot_u=tot_i=0; lanesum=0; W=40
    per=len(steps)//W
    for w in range(W):
        s=steps[w*per:(w+1)*per]
        if sort:
            p=pred[w*per:(w+1)*per]
            out=[]
            for b in range(0,len(s),kw.get('grab',256)):
                o=np.argsort(-p[b:b+kw.get('grab',256)],kind='stable'); out.append(s[b:b+kw.get('grab',256)][o])
            s=np.concatenate(out)
        u,i,l=simulate(s,**kw); tot_u+=u; tot_i+=i; lanesum+=l
    return tot_u*25/(tot_i*32), lanesum/W
for name,(s,p) in {"bounce":(d['sb'],d['pb']),"shadow":(d['ss'],d['ps'])}.items():
    s=s[:80000]; p=p[:80000]
    print(name,"random order  eff %.3f lanes-in-steps %.1f"%run(s))
    print(name,"sorted by pred eff %.3f lanes-in-steps %.1f"%run(s,p,sort=True))
    print(name,"sorted by TRUE eff %.3f lanes-in-steps %.1f"%run(s,s,sort=True))
    for lm in (20,26,30):
        print("   live_min",lm,"random %.3f"%run(s,live_min=lm)[0],"sorted %.3f"%run(s,p,sort=True,live_min=lm)[0])
print("---- grab size (sorted by pred) / chunk size")
s=d['sb'][:80000]; p=d['pb'][:80000]
for g in (128,256,512,1024,4096):
    print("bounce grab",g,"sorted %.3f"%run(s,p,sort=True,grab=g)[0], "random %.3f"%run(s,grab=g)[0])
for c in (8,12,16,24):
    print("bounce chunk",c,"random %.3f"%run(s,chunk=c)[0],"sorted %.3f"%run(s,p,sort=True,chunk=c)[0])
# mixing both kinds in one stream (interleaved) vs separate
s2=np.concatenate([d['sb'][:40000],d['ss'][:40000]]); p2=np.concatenate([d['pb'][:40000],d['ps'][:40000]])
perm=np.random.RandomState(2).permutation(len(s2))
print("mixed kinds random %.3f sorted %.3f"%(run(s2[perm])[0], run(s2[perm],p2[perm],sort=True)[0]))
print("---- where the loss is (bounce, random order)")
print("no refill/retire cost   %.3f"%run(s,refill_instr=0,retire_instr=0)[0])
print("half refill/retire cost %.3f"%run(s,refill_instr=70,retire_instr=30)[0])
print("chunk 1 no chunk overhead %.3f"%run(s,chunk=1,chunk_over=0)[0])
print("chunk 1, no overheads, live 32 %.3f"%run(s,chunk=1,chunk_over=0,refill_instr=0,retire_instr=0,live_min=32)[0])
print("chunk 4 over 2 %.3f"%run(s,chunk=4,chunk_over=2)[0])
