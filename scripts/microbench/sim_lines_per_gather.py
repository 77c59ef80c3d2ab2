import numpy as np, sys
from scipy.spatial import cKDTree
dp=0.0125
# fluid block 2.0 x 1.0 x 0.5 (LL, LH, LW) like dam break water; lattice centres
nx,ny,nz=int(2.0/dp),int(1.0/dp),int(0.5/dp)
nx=60  # subsample in x to bound cost
rng=np.random.default_rng(0)
X=np.stack(np.meshgrid((np.arange(nx)+.5)*dp,(np.arange(ny)+.5)*dp,(np.arange(nz)+.5)*dp,indexing='ij'),-1).reshape(-1,3)
jit=float(sys.argv[1]) if len(sys.argv)>1 else 0.0
X=X+jit*dp*rng.uniform(-1,1,X.shape)
N=len(X)
h=1.3*dp; rc=2*h
lower=-4*dp-2*rc
c=np.floor((X-lower)/rc).astype(np.int64)
ncell=c.max(0)+2
lin=(c[:,0]*ncell[1]+c[:,1])*ncell[2]+c[:,2]
order=np.argsort(lin,kind='stable')   # slot -> id
Xs=X[order]; lins=lin[order]; cs=c[order]
tree=cKDTree(Xs)
print("N",N)
# sample warps in the interior
nw=N//32
sel=rng.choice(np.arange(nw//4,3*nw//4),size=600,replace=False)
def lines(idx,rec): return len(set((np.asarray(idx)*rec)//128))
res={'cur16':[], 'cur32':[], 'cur4':[],'al16':[],'al32':[],'al4':[],'it_cur':[],'it_al':[],'sal16':[],'sal32':[],'it_sal':[]}
for w in sel:
    slots=np.arange(w*32,w*32+32)
    nb=tree.query_ball_point(Xs[slots],rc*(1-1e-9))
    lists=[np.array(sorted(j for j in l if j!=s)) for l,s in zip(nb,slots)]
    cnt=[len(l) for l in lists]; mx=max(cnt)
    # current: k-th of each lane
    for k in range(mx):
        idx=[l[k] for l in lists if k<len(l)]
        res['cur16'].append(lines(idx,16)); res['cur32'].append(lines(idx,32)); res['cur4'].append(lines(idx,4))
    res['it_cur'].append(mx)
    # run aligned: group each lane's list by column (cx,cy) of neighbour
    def colkey(j): return cs[j,0]*ncell[1]+cs[j,1]
    for mode,name in ((0,'al'),(1,'sal')):
        per=[]
        keys=set()
        for l in lists:
            d={}
            for j in l:
                kx = colkey(j) if mode==0 else cs[j,0]
                d.setdefault(kx,[]).append(j)
            per.append(d); keys|=set(d)
        tot=0
        for kx in sorted(keys):
            m=max(len(d.get(kx,[])) for d in per)
            tot+=m
            for k in range(m):
                idx=[d[kx][k] for d in per if kx in d and k<len(d[kx])]
                res[name+'16'].append(lines(idx,16)); res[name+'32'].append(lines(idx,32))
                if mode==0: res['al4'].append(lines(idx,4))
        res['it_'+name].append(tot)
for k,v in res.items(): print(k, np.mean(v))
print("mean count", np.mean([len(l) for l in lists]))
