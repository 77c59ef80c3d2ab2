import numpy as np, sys
from scipy.spatial import cKDTree
dp=0.0125
nx,ny,nz=60,int(1.0/dp),int(0.5/dp)
rng=np.random.default_rng(0)
X=np.stack(np.meshgrid((np.arange(nx)+.5)*dp,(np.arange(ny)+.5)*dp,(np.arange(nz)+.5)*dp,indexing='ij'),-1).reshape(-1,3)
N=len(X); h=1.3*dp; rc=2*h; lower=-4*dp-2*rc
c=np.floor((X-lower)/rc).astype(np.int64); ncell=c.max(0)+2
lin=(c[:,0]*ncell[1]+c[:,1])*ncell[2]+c[:,2]
order=np.argsort(lin,kind='stable'); Xs=X[order]
tree=cKDTree(Xs)
nw=N//32
sel=rng.choice(np.arange(nw//4,3*nw//4),size=300,replace=False)
def wf128(rows):
    tot=0
    for q in range(4):
        s=set(x for x in rows[8*q:8*q+8] if x>=0)
        if not s: continue
        tot+=np.bincount([x%8 for x in s],minlength=8).max()
    return tot
def local(l,lane,G):
    out=[]
    for g in range(0,len(l),G):
        grp=l[g:g+G]; n=len(grp)
        rows=[-1]*n; left=[]
        for s in grp:
            r=(s-lane-g)&7   # class relative to row index: row k wants (s - lane - k)%8==0  -> position p in group: (g+p)%8 == (s-lane)%8
            placed=False
            p=((s-lane)-g)%8
            while p<n:
                if rows[p]<0: rows[p]=s; placed=True; break
                p+=8
            if not placed: left.append(s)
        for i in range(n):
            if rows[i]<0: rows[i]=left.pop()
        out+=rows
    return out
res={}
for G in (8,16,32,128):
    W=[];AL=[]
    for w in sel:
        slots=np.arange(w*32,w*32+32)
        nb=tree.query_ball_point(Xs[slots],rc*(1-1e-9))
        lists=[sorted(j for j in l if j!=s) for l,s in zip(nb,slots)]
        newl=[local(l,lane,G) for lane,l in enumerate(lists)]
        mx=max(len(l) for l in lists)
        for k in range(mx):
            W.append(wf128([l[k] if k<len(l) else -1 for l in newl]))
        AL.append(np.mean([((l[k]-lane-k)%8==0) for lane,l in enumerate(newl) for k in range(len(l))]))
    print("G",G,"wavefronts/row",np.mean(W),"aligned",np.mean(AL))
