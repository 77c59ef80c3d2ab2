import numpy as np, sys
from scipy.spatial import cKDTree
dp=0.0125
nx,ny,nz=60,int(1.0/dp),int(0.5/dp)
rng=np.random.default_rng(0)
X=np.stack(np.meshgrid((np.arange(nx)+.5)*dp,(np.arange(ny)+.5)*dp,(np.arange(nz)+.5)*dp,indexing='ij'),-1).reshape(-1,3)
jit=float(sys.argv[1]) if len(sys.argv)>1 else 0.0
X=X+jit*dp*rng.uniform(-1,1,X.shape)
N=len(X); h=1.3*dp; rc=2*h; lower=-4*dp-2*rc
c=np.floor((X-lower)/rc).astype(np.int64); ncell=c.max(0)+2
lin=(c[:,0]*ncell[1]+c[:,1])*ncell[2]+c[:,2]
order=np.argsort(lin,kind='stable'); Xs=X[order]
tree=cKDTree(Xs)
nw=N//32
sel=rng.choice(np.arange(nw//4,3*nw//4),size=400,replace=False)
def wf128(rows):  # rows: list of per-lane slot or -1 ; LDG.128 on 16B records: per quarter-warp max multiplicity of (slot mod 8) over distinct addresses
    tot=0
    for q in range(4):
        s=[x for x in rows[8*q:8*q+8] if x>=0]
        if not s: continue
        s=set(s)
        cnt=np.bincount([x%8 for x in s],minlength=8)
        tot+=cnt.max()
    return tot
def wf256(rows): # 32B records: assume processed in 8 phases of 4 lanes, 4 groups (slot mod 4)
    tot=0
    for q in range(8):
        s=set(x for x in rows[4*q:4*q+4] if x>=0)
        if not s: continue
        cnt=np.bincount([x%4 for x in s],minlength=4); tot+=cnt.max()
    return tot
def wf32(rows):
    s=set(x for x in rows if x>=0)
    if not s: return 0
    return np.bincount([x%32 for x in s],minlength=32).max()
res={k:[] for k in ['cur128','cur256','cur32','new128','new256','new32','rows_cur','rows_new','aligned']}
STRIDE=128
for w in sel:
    slots=np.arange(w*32,w*32+32)
    nb=tree.query_ball_point(Xs[slots],rc*(1-1e-9))
    lists=[sorted(j for j in l if j!=s) for l,s in zip(nb,slots)]
    mx=max(len(l) for l in lists)
    res['rows_cur'].append(mx)
    for k in range(mx):
        row=[l[k] if k<len(l) else -1 for l in lists]
        res['cur128'].append(wf128(row)); res['cur256'].append(wf256(row)); res['cur32'].append(wf32(row))
    # new ordering
    newl=[]
    al=0;tot=0
    for lane,l in enumerate(lists):
        rows=[-1]*STRIDE; cnt=[0]*8
        for s in l:
            r=(s-lane)&7; m=cnt[r]; cnt[r]+=1
            row=8*m+r
            if m>=16 or row>=STRIDE or rows[row]>=0:
                row=max(i for i in range(STRIDE) if rows[i]<0)
            rows[row]=s
        n=len(l)
        # compaction: move entries at rows>=n into holes < n
        hi=[i for i in range(n,STRIDE) if rows[i]>=0]; lo=[i for i in range(n) if rows[i]<0]
        assert len(hi)==len(lo)
        for a,b in zip(lo,hi[::-1]):
            rows[a]=rows[b]; rows[b]=-1
        al+=sum(1 for i in range(n) if (rows[i]-lane)&7==(i&7)); tot+=n
        newl.append(rows[:n])
    res['aligned'].append(al/tot)
    res['rows_new'].append(max(len(l) for l in newl))
    for k in range(mx):
        row=[l[k] if k<len(l) else -1 for l in newl]
        res['new128'].append(wf128(row)); res['new256'].append(wf256(row)); res['new32'].append(wf32(row))
for k,v in res.items(): print(k, round(float(np.mean(v)),3))
