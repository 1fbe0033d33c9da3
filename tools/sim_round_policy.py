"""Offline simulation of line-search round policies (candidates per problem per round) on recorded trip counts
(tools/dump_trips.py -> gpurun_out/trips.npz) with the measured launch-cost model of the backward + roll-out kernels."""
import numpy as np
d=np.load('/root/repo/gpurun_out/trips.npz')
t=d['trips']; 
if t.shape[0]!=20: t=t.T
SCALE=4  # 8192 -> 32768
# measured (slots, ms) backward+rollout (direct feed; dense async for first round handled separately)
pts=np.array([[0,0.55],[1024,0.70],[8192,0.93],[32768,1.17],[65536,1.9],[98304,2.78],[131072,3.38],[262144,6.7]])
def cost(n): return np.interp(n,pts[:,0],pts[:,1])
OVER=0.12
def simulate(policy, trips):
    total=0; slots=0; rounds=0
    for it in range(trips.shape[0]):
        rem=trips[it].copy().astype(int); rem=rem[rem>0]
        done=np.zeros(len(rem),int)
        r=0
        while len(rem)>0:
            n_act=len(rem)*SCALE
            L=policy(r,n_act)
            total+=cost(n_act*L)+OVER; slots+=n_act*L; rounds+=1
            rem=rem-L; rem=rem[rem>0]; r+=1
    return total, slots/ (trips.shape[0]*trips.shape[1]*SCALE), rounds/trips.shape[0]
def cur(cap,Lmax=32):
    return lambda r,n: int(np.clip(cap//n,1,Lmax))
print("current cap75776:", simulate(cur(75776),t))
print("cap 66304:", simulate(cur(66304),t))
print("cap 151552:", simulate(cur(151552),t))
print("cap75776 Lmax128:", simulate(cur(75776,128),t))
def pol(caps,Lmaxs):
    def f(r,n):
        i=min(r,len(caps)-1)
        return int(np.clip(caps[i]//n,1,Lmaxs[i]))
    return f
best=None
import itertools
for c0 in (65536,98304,131072,140000):
  for c1 in (32768,49152,65536,75776):
    for c2 in (16384,32768,65536):
      for l1 in (1,2,3,4):
        for l2 in (2,4,8,16):
          p=pol([c0,c1,c2,75776],[8,l1,l2,128])
          res=simulate(p,t)
          if best is None or res[0]<best[0][0]: best=(res,(c0,c1,c2,l1,l2))
print("best",best)
print("---- refine")
res=[]
for c0 in (98304,131072,132608):
  for l0 in (4,6,8):
    for c1 in (49152,65536,75776,98304):
      for l1 in (2,3,4):
        for c2 in (32768,49152,65536,75776):
          for l2 in (8,16,32):
            for l3 in (64,128,256):
              p=pol([c0,c1,c2,75776],[l0,l1,l2,l3])
              r=simulate(p,t)
              res.append((r[0],r[1],r[2],(c0,l0,c1,l1,c2,l2,l3)))
res.sort()
for r in res[:12]: print(r)
# simpler family: same cap every round, per-round Lmax
print("---- same cap 75776, per-round Lmax")
for lm in ([2,2,4,128],[2,2,2,128],[2,2,4,16,128],[4,3,16,128],[2,2,3,16,128],[2,2,2,8,128]):
    print(lm, simulate(pol([75776]*len(lm),lm),t))
print("---- cap 132608 first then 75776")
for lm in ([4,3,16,128],[4,2,8,128],[4,3,8,128],[4,4,16,128],[4,3,16,256]):
    print(lm, simulate(pol([132608,75776,75776,75776],lm),t))
