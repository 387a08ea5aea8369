"""How much of c2 (events of x,y ~ N(0,1) in (x+y, sqrt(x^2+y^2)), 203x203 bins) a shared-memory hot window of W bins
can hold: best rectangle vs one interval per row (what fillable.choose_ragged places) vs any W bins."""
import numpy as np, heapq
rs=np.random.RandomState(1)
n=4*10**6
x=rs.normal(size=n); y=rs.normal(size=n)
def b(v,lo,hi,nb):
    t=np.floor((v-lo)*(nb/(hi-lo)))+1
    return np.clip(t,0,nb+1).astype(int)
i2=b(x+y,-5,5,200); i8=b(np.sqrt(x*x+y*y),0,5,200)
H=np.zeros((203,203)); np.add.at(H,(i2,i8),1)
tot=H.sum()
def ragged(H,W):
    # per row contiguous interval; greedy trim of lowest-density end
    R=H.shape[0]
    lo=np.zeros(R,int); hi=np.zeros(R,int)
    heap=[]
    vol=0
    for r in range(R):
        nz=np.flatnonzero(H[r])
        if len(nz)==0: continue
        lo[r]=nz[0]; hi[r]=nz[-1]+1
        vol+=hi[r]-lo[r]
        heapq.heappush(heap,(H[r,lo[r]],r,0)); heapq.heappush(heap,(H[r,hi[r]-1],r,1))
    while vol>W and heap:
        v,r,end=heapq.heappop(heap)
        if hi[r]-lo[r]<=0: continue
        if end==0:
            if H[r,lo[r]]!=v: continue
            lo[r]+=1
            if hi[r]>lo[r]: heapq.heappush(heap,(H[r,lo[r]],r,0))
        else:
            if H[r,hi[r]-1]!=v: continue
            hi[r]-=1
            if hi[r]>lo[r]: heapq.heappush(heap,(H[r,hi[r]-1],r,1))
        vol-=1
    cov=sum(H[r,lo[r]:hi[r]].sum() for r in range(R))/tot
    return cov
def rect(H,W):
    best=0
    m0=H.sum(1); m1=H.sum(0)
    # brute force over widths
    for n0 in range(10,203,2):
        n1=W//n0
        if n1<1 or n1>203: continue
        # best position: use 2D prefix sums
        P=np.zeros((204,204)); P[1:,1:]=H.cumsum(0).cumsum(1)
        S=P[n0:,n1:]-P[:-n0,n1:]-P[n0:,:-n1]+P[:-n0,:-n1]
        best=max(best,S.max())
    return best/tot
top=np.sort(H.ravel())[::-1].cumsum()/tot
for W in (3200,4266,6400,8533,9600,12800,14000):
    print(W,"rect %.3f ragged %.3f anybins %.3f"%(rect(H,W),ragged(H,W),top[W-1]))
