import numpy as np, sys
pts=np.array([[0,0.55],[1024,0.70],[8192,0.93],[32768,1.17],[65536,1.9],[98304,2.78],[131072,3.38],[262144,6.7]])
OVER=0.12
def T(n): return float(np.interp(n,pts[:,0],pts[:,1]))+OVER
def dp_policy(hist, scale, cap0=132608, cap=75776, maxc=160):
    """hist[k] = number of problems needing exactly k trips (k>=1). returns list of cut points."""
    K=len(hist)-1
    surv=np.array([hist[k+1:].sum() for k in range(K+1)])*scale   # surv[c] = problems needing more than c trips
    INF=1e18
    best=[INF]*(K+1); nxt=[None]*(K+1)
    best[K]=0.0
    for c in range(K-1,-1,-1):
        n=surv[c]
        if n==0: best[c]=0.0; continue
        for c2 in range(c+1,K+1):
            L=c2-c
            if L>128: break
            slots=n*L
            if slots>(cap0 if c==0 else cap) and L>1: break
            v=T(slots)+best[c2]
            if v<best[c]: best[c]=v; nxt[c]=c2
    cuts=[]; c=0
    while c<K and nxt[c] is not None: c=nxt[c]; cuts.append(c)
    return cuts,best[0]
def run_cost(trips_row, cuts, scale):
    rem=trips_row[trips_row>0]; tot=0; c_prev=0; slots=0; rounds=0
    i=0
    while len(rem)>0:
        c=cuts[i] if i<len(cuts) else c_prev+128
        L=c-c_prev
        tot+=T(len(rem)*scale*L); slots+=len(rem)*scale*L; rounds+=1
        rem=rem[rem>c]; c_prev=c; i+=1
    return tot,slots,rounds
if __name__=="__main__":
    d=np.load('/root/repo/gpurun_out/trips.npz'); t=d['trips']
    if t.shape[0]!=20: t=t.T
    scale=4
    tot_oracle=tot_prev=0; slots_prev=0
    prev_cuts=[4,7,23]
    for it in range(20):
        h=np.bincount(t[it][t[it]>0],minlength=130)
        cuts,b=dp_policy(h,scale)
        tot_oracle+=run_cost(t[it],cuts,scale)[0]
        r=run_cost(t[it],prev_cuts,scale); tot_prev+=r[0]; slots_prev+=r[1]
        print(it, "oracle cuts",cuts[:6],"cost %.2f"%b, "| with prev-iteration cuts %.2f"%r[0], "rounds",r[2])
        prev_cuts=cuts
    print("oracle total",tot_oracle,"prev-iteration policy total",tot_prev, "cands/iter", slots_prev/(32768*20))
    # MPC data
    for f in ('mpc_trips_3','mpc_trips_12','mpc_trips_25'):
        m=np.load(f'/root/repo/gpurun_out/{f}.npy').T   # (10, 4096)
        def cur(n,r):
            if r==0: c=int(np.clip(49152//n,4,16)); capx=132608
            elif n<=148 or r>=3: c=128; capx=75776
            else: c=3 if r==1 else 16; capx=75776
            return int(np.clip(capx//n,1,c))
        tot_cur=tot_dp=0
        prev=None
        for it in range(10):
            row=m[it]; rem=row[row>0]; r=0; cuts=[]; c=0
            # current policy
            while len(rem)>0:
                L=cur(len(rem),r); tot_cur+=T(len(rem)*L); c+=L; rem=rem[rem>c]; r+=1
            h=np.bincount(row[row>0],minlength=130)
            if prev is None: prev=dp_policy(h,1)[0]   # (first iteration: previous window's histogram - same distribution)
            tot_dp+=run_cost(row,prev,1)[0]
            prev=dp_policy(h,1)[0]
        print(f,"current policy %.1f ms per window, adaptive (prev-iteration histogram) %.1f"%(tot_cur,tot_dp))
