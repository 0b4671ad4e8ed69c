"""Checks the tie rule the device uses for ConvexTotalSplitter{<:ConstrainedCost} (dynamic.cu, k_dp_convex_constrained) against the
restated reference algorithm (oracle: chunk_convex_constrained! per layer, ConvexTotalChunker.jl:167-265) on random small cases.
CPU only:  python tools/convex_k_rule.py   ->  "<cases> [<cases matching>, 0]"."""
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import numpy as np
import chainb200 as cp, pyoracle as ref
from helpers import sprand
INF=float('inf')

def solve_rule(A, K, f, wcoef, wmax, variant):
    n=A.n
    pos=A.colptr
    def w(j,jp): return wcoef[0]+(jp-j)*wcoef[1]+int(pos[jp-1]-pos[j-1])*wcoef[2]
    over=lambda j,jp: w(j,jp)>wmax
    # column_constraints
    lo=[0]*(K+1); hi=[0]*(K+1)
    jp=n+1
    for k in range(K,0,-1):
        lo[k]=jp; j=jp
        while j-1>=1 and not over(j-1,jp): j-=1
        jp=j
    j=1
    for k in range(1,K+1):
        jp=j
        while jp+1<=n+1 and not over(j,jp+1): jp+=1
        hi[k]=jp; j=jp
    if hi[K]<n+1: return [1]*K+[n+1]
    # all costs
    js=[];jps=[]
    for a in range(1,n+2):
        for b in range(a,n+2): js.append(a); jps.append(b)
    C={}
    vals=ref.oracle_query(f,A,js,jps)
    for a,b,v in zip(js,jps,vals): C[(a,b)]=v
    cst=[None]*(K+1); ptr=[None]*(K+1)
    cst[1]=[INF]*(n+2); ptr[1]=[0]*(n+2)
    for jp in range(lo[1],hi[1]+1): cst[1][jp]=C[(1,jp)]; ptr[1][jp]=1
    for k in range(2,K+1):
        pc=cst[k-1]
        cur=[INF]*(n+2); pt=[0]*(n+2)
        def inw(x): return lo[k]<=x<=hi[k]
        for jp in range(lo[k],hi[k]+1):
            cur[jp]=pc[jp]+C[(jp,jp)]; pt[jp]=jp
        J0=lo[k-1]; JP1=hi[k]
        jp1=J0+1
        while jp1<JP1 and not over(J0,jp1+1): jp1+=1
        j0=J0; t=0
        prev_j0=None
        while True:
            for jp in range(j0+1,jp1+1):
                if not inw(jp): continue
                cand_in=[(pc[j]+C[(j,jp)],j) for j in range(j0,jp)]
                vin=min(c for c,_ in cand_in)
                jin=min(j for c,j in cand_in if c==vin)   # leftmost
                if vin<=cur[jp]:
                    cur[jp]=vin; pt[jp]=jin
            if jp1==JP1: break
            # staircase: next block
            nj0=jp1
            reach=jp1
            while reach<JP1 and not over(jp1,reach+1): reach+=1
            for jp in range(jp1+1,reach+1):
                cand=[(pc[j]+C[(j,jp)],j) for j in range(j0+1,jp1+1) if not over(j,jp)]
                v=min(c for c,_ in cand)
                if variant==0: jj=max(j for c,j in cand if c==v)   # rightmost
                else: jj=min(j for c,j in cand if c==v)
                if inw(jp):
                    cur[jp]=v; pt[jp]=jj
            j0=nj0; jp1=reach
        cst[k]=cur; ptr[k]=pt
    spl=[0]*(K+2)
    spl[K+1]=n+1
    jp=n+1
    for k in range(K,0,-1):
        jp=ptr[k][jp]; spl[k]=jp
    return spl[1:]

rng=np.random.default_rng(11)
tot=0; ok=[0,0]
for trial in range(2500):
    n=int(rng.integers(1,26)); m=n if trial%5==4 else int(rng.integers(1,14))
    A=sprand(rng,m,n,float(rng.choice([0.1,0.3,0.6])))
    K=int(rng.integers(1,9))
    f=[cp.AffineConnectivityModel(0,0,0,1),cp.AffineConnectivityModel(0,3,1,3),cp.AffineWorkModel(0,0,0),cp.AffineConnectivityModel(0,10,1,100),cp.AffineMonotonizedSymmetricConnectivityModel(0,3,1,3,2)][trial%5] if trial%7 else cp.AffineConnectivityModel(0.5,0.25,1.5,3.0)
    wmax=int(rng.choice([2,3,4,8]))
    wc=[(0,1,0),(0,1,0),(1,1,1),(0,2,1)][int(rng.integers(0,4))]
    if wc!=(0,1,0): wmax=int(rng.choice([4,8,15,30]))
    spec=cp.ConstrainedCost(f, cp.AffineWorkModel(*wc), wmax)
    r=ref.partition_stripe(A,K,cp.ConvexTotalSplitter(spec)).spl.tolist()
    tot+=1
    for v in (0,):
        g=solve_rule(A,K,f,wc,wmax,v)
        if g==r: ok[v]+=1
        elif v==0 and tot-ok[0]<4: print('diff',m,n,K,wc,wmax,f.coef,r,g)
print(tot,ok)
