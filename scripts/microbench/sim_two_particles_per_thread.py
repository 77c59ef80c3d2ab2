import numpy as np, sys
from scipy.spatial import cKDTree
dp=0.0125
nx,ny,nz=60,int(1.0/dp),int(0.5/dp)
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
# in-cell order: by id (random-ish) or by z
mode=sys.argv[2] if len(sys.argv)>2 else 'id'
if mode=='z':
    order=np.lexsort((X[:,2],lin))
else:
    order=np.argsort(lin,kind='stable')
Xs=X[order]
tree=cKDTree(Xs)
nw=N//64
sel=rng.choice(np.arange(nw//4,3*nw//4),size=300,replace=False)
def lines(idx,rec): return len(set((np.asarray(idx)*rec)//128))
L16=[];L32=[];L4=[];IT=[];U=[]
for w in sel:
    slots=np.arange(w*64,w*64+64)
    nb=tree.query_ball_point(Xs[slots],rc*(1-1e-9))
    lists=[]
    for m in range(32):
        a=set(nb[2*m])-{slots[2*m]}; b=set(nb[2*m+1])-{slots[2*m+1]}
        lists.append(np.array(sorted(a|b))); U.append(len(a|b))
    mx=max(len(l) for l in lists)
    IT.append(mx)
    for k in range(mx):
        idx=[l[k] for l in lists if k<len(l)]
        L16.append(lines(idx,16));L32.append(lines(idx,32));L4.append(lines(idx,4))
print("union",np.mean(U),"iters/warp",np.mean(IT),"lines16",np.mean(L16),"lines32",np.mean(L32),"lines4",np.mean(L4))
print("wavefronts per particle 32B:", np.mean(IT)*(np.mean(L32)+1)/64*32, "vs single ~", 79.4*(20.8+1))
