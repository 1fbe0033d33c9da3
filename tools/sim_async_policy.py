import numpy as np
d=np.load('/root/repo/gpurun_out/trips.npz'); t=d['trips']
if t.shape[0]==20: t=t.T   # (B, 20)
B,N=t.shape; SCALE=32768//B
pts=np.array([[0,0.55],[1024,0.70],[8192,0.93],[32768,1.17],[65536,1.9],[98304,2.78],[131072,3.38],[262144,6.7]])
def T_ls(n): return np.interp(n,pts[:,0],pts[:,1])+0.12 if n>0 else 0.0
def T_lin(n): return (0.9+0.8*min(n,32768)/32768) if n>0 else 0.0   # expansion+adjoint+status, floor-ish
def cand(tr, pol): 
    a,b,c,dd=pol
    return np.where(tr<a,a,np.where(tr<a+b,b,np.where(tr<a+b+c,c,dd)))
def sim(lag_frac, pol=(4,3,16,128), verbose=False):
    it=np.zeros(B,int); tr=np.zeros(B,int); phase=np.ones(B,int)  # 1=LS (after initial lin)
    total=T_lin(32768); ticks=0; slots_total=0
    while (it<N).any():
        ls=(phase==1)&(it<N)
        n_ls=ls.sum()
        if n_ls>0:
            c=cand(tr[ls],pol)
            need=t[ls,it[ls]]-tr[ls]
            total+=T_ls(c.sum()*SCALE); slots_total+=c.sum()*SCALE
            done=need<=c
            idx=np.nonzero(ls)[0]
            tr[idx]+=c
            fin=idx[done]
            phase[fin]=0; tr[fin]=0
        rem=((phase==1)&(it<N)).sum()
        lin=(phase==0)&(it<N)
        if lin.sum()>0 and rem<=lag_frac*B:
            total+=T_lin(lin.sum()*SCALE)
            it[lin]+=1; phase[lin]=1
        ticks+=1
    return total, ticks, slots_total/(B*SCALE*N)
print("lockstep-like (lag 0):", sim(0.0))
for lf in (0.002,0.02,0.05,0.1,0.2,0.3,0.5,1.0):
    print(lf, sim(lf))
for pol in ((4,3,16,128),(4,4,16,128),(4,8,128,128),(5,4,16,128),(4,16,128,128),(3,3,16,128)):
    print(pol, sim(0.2,pol), sim(1.0,pol))
print("---- lag-aware candidate counts")
def sim2(lag_frac, pol, pol_behind, lin_every=False):
    it=np.zeros(B,int); tr=np.zeros(B,int); phase=np.ones(B,int)
    total=T_lin(32768); ticks=0; slots_total=0
    while (it<N).any():
        ls=(phase==1)&(it<N)
        lead=it[it<N].max() if (it<N).any() else N
        if ls.sum()>0:
            idx=np.nonzero(ls)[0]
            behind=it[idx]<lead
            c=np.where(behind,cand(tr[idx],pol_behind),cand(tr[idx],pol))
            need=t[idx,it[idx]]-tr[idx]
            total+=T_ls(c.sum()*SCALE); slots_total+=c.sum()*SCALE
            done=need<=c
            tr[idx]+=c
            fin=idx[done]; phase[fin]=0; tr[fin]=0
        rem=((phase==1)&(it<N)).sum()
        lin=(phase==0)&(it<N)
        if lin.sum()>0 and rem<=lag_frac*B:
            total+=T_lin(lin.sum()*SCALE)
            it[lin]+=1; phase[lin]=1
        ticks+=1
    return round(float(total),1), ticks, round(slots_total/(B*SCALE*N),2)
for lf in (0.05,0.2,1.0):
  for pol in ((4,3,16,128),(4,4,16,128),(5,4,16,128)):
    for pb in ((8,16,128,128),(6,10,128,128),(16,128,128,128),(5,8,32,128)):
        print(lf,pol,pb,sim2(lf,pol,pb))
