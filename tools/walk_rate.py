"""Estimate, with scipy on the CPU, how often the graph-walk certificate of csrc/walk.cuh holds (needs /tmp/sim_scene.npz from
tools/sim_run.py): pass rate of 4 D^2 < r2 at converged / slightly stale / far poses for several list sizes, and the number of
moves a walk from a stale seed needs."""
import numpy as np, time
from scipy.spatial import cKDTree
z=np.load('/tmp/sim_scene.npz'); p1,p2,gt,init=z['p1'],z['p2'],z['gt'],z['init']
def q_of(poses): # src p1 into dst p2 local
    Ps,Pd=poses[0],poses[1]
    g=p1@Ps[:3,:3].T+Ps[:3,3]; return (g-Pd[:3,3])@np.linalg.inv(Pd[:3,:3]).T
T=cKDTree(p2)
K=9
dk,ik=T.query(p2,k=K+1)      # self + 9 others
for k in (6,8,12):
    dkk,ikk=T.query(p2,k=k+2)
    R=dkk[:,k+1]              # distance to the (k+1)-th other point = first outside a list of k others
    for name,poses in (('gt',gt),('near: trans 0.1mm off',None),('half',None)):
        if name=='gt': P=gt
        elif name.startswith('near'):
            P=gt.copy(); P[0,:3,3]+=1e-4
        else:
            P=init.copy(); P[:, :3,3]=0.5*(init[:, :3,3]+gt[:, :3,3])
        q=q_of(P); D,m=T.query(q)
        ok=(2*D<R[m])
        print(f'k={k} {name:24s} median D {np.median(D):.2e}  median R {np.median(R):.2e}  certificate pass rate {ok.mean():.3f}')
# hops: seed = NN under slightly different pose, walk greedy on k=8 graph
k=8; dkk,ikk=T.query(p2,k=k+2); nb=ikk[:,1:k+1]; R=dkk[:,k+1]
P0=gt.copy(); P0[0,:3,3]+=3e-4     # previous round pose 0.3 mm off
q0=q_of(P0); _,seed=T.query(q0)
q=q_of(gt); D,m=T.query(q)
cur=seed.copy(); hops=np.zeros(len(q),int); active=np.ones(len(q),bool)
for h in range(6):
    dcur=np.linalg.norm(q-p2[cur],axis=1)
    dn=np.linalg.norm(q[:,None,:]-p2[nb[cur]],axis=2)
    j=np.argmin(dn,axis=1); better=dn[np.arange(len(q)),j]<dcur
    mv=better&active
    cur[mv]=nb[cur[mv],j[mv]]; hops[mv]+=1; active&=better
    if not active.any(): break
dcur=np.linalg.norm(q-p2[cur],axis=1)
cert=(2*dcur<R[cur])
print('walk from 0.3mm-stale seeds: exact NN reached', np.mean(cur==m), 'certified', cert.mean(), 'certified&correct', np.mean(cert&(cur==m)), 'certified but wrong', np.mean(cert&(cur!=m)), 'hops hist', np.bincount(hops)[:7])
