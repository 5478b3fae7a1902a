"""CPU-side look at the fold's work distribution on the C4 workload (no GPU): 40 consecutive synthetic keyframes are binned with
numpy; for chunks of M consecutive bucket records it reports the lane efficiency of a count-sorted hand-out of cells to 8 warps
("LPT") against a fixed thread = cell mapping, the mean run length, and the heaviest cell per chunk (the serial chain behind a
chunk's barrier).  python profiles/fold_balance_sim.py"""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from traversability_mapping_b200 import synth
true, drift = bench.c4_poses(2000)
res=0.1
recs=[]  # (bin, op, cell) arrays
K0, NK = 300, 40
for k in range(K0, K0+NK):
    c = synth.make_keyframe(7, k, true[k], max_range=bench.C4.get("max_range",20.0), rings=128, azimuths=2048)
    p = c[:, :3]
    ok = np.isfinite(p).all(1) & ((p!=0).any(1))
    p = p[ok]
    T = true[k].astype(np.float64)
    pb = p.astype(np.float64) + np.array([0,0,1.0])
    r = np.linalg.norm(pb,axis=1); pb = pb[(r>=bench.C4.get("min_range",1.0))&(r<=bench.C4.get("max_range",20.0))]
    pm = pb @ T[:, :3].T + T[:, 3]
    ci = np.rint(pm[:,0]/res).astype(np.int64); cj = np.rint(pm[:,1]/res).astype(np.int64)
    b = (ci>>4)*4096+(cj>>4); cell=(ci&15)+16*(cj&15)
    recs.append(np.stack([b, np.full_like(b,k), cell],1))
R = np.concatenate(recs)
order = np.argsort(R[:,0], kind='stable')
R = R[order]
ub, st, cn = np.unique(R[:,0], return_index=True, return_counts=True)
print("bins", len(ub), "records", len(R), "heaviest", cn.max())
def analyze(M):
    tot_ideal=0; tot_lpt=0; tot_fixed=0; heads=0; nrec=0; distinct=0; rows=0; hrows=0; hdist=0
    for s,n in zip(st,cn):
        S = R[s:s+n]
        for a in range(0,n,M):
            ch = S[a:a+M]
            key = ch[:,1]*256+ch[:,2]
            cnt = np.bincount(ch[:,2], minlength=256)
            srt = np.sort(cnt)[::-1]
            lpt = sum(srt[g*32] for g in range(8))
            fixed = sum(cnt[g*32:(g+1)*32].max() for g in range(8))
            tot_ideal += len(ch)/32; tot_lpt += lpt; tot_fixed += fixed
            h = np.ones(len(ch),bool); h[1:] = key[1:]!=key[:-1]; h[::32]=True
            heads += h.sum(); nrec+=len(ch)
    print("M",M,"eff LPT %.3f fixed %.3f  avg run %.2f"%(tot_ideal/tot_lpt, tot_ideal/tot_fixed, nrec/heads))
for M in [2048,4096,8192,16384]:
    analyze(M)
def crit(M):
    mx=[]; avg=[]; w=[]
    for s,n in zip(st,cn):
        S = R[s:s+n]
        for a in range(0,n,M):
            ch = S[a:a+M]
            cnt = np.bincount(ch[:,2], minlength=256)
            mx.append(cnt.max()); avg.append(len(ch)/256); w.append(len(ch))
    mx=np.array(mx); w=np.array(w)
    print("M",M,"chunks",len(mx),"record-weighted mean max count %.1f"%( (mx*w).sum()/w.sum()), "mean max %.1f"%mx.mean(), "sum max / sum records*256 = %.2f"%(mx.sum()*256/w.sum()))
for M in [1024,2048,3072]:
    crit(M)
