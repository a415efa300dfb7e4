#!/usr/bin/env python
"""How much would combining the contributions of B consecutive (Morton-ordered) P2 elements in shared
memory save before the scatter reaches L2? (CPU only, no GPU needed.)

  python tools/combine_study.py [cells per side, default 14] [basis order, default 2]

For clusters of B elements: contributions per unique values[] entry, distinct 32-byte sectors touched per
element today (one RED instruction = one column component of one column node, 10 runs of 3 doubles) against
the sectors of the cluster's unique entries, and the shared memory the unique entries need."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from polyfem_b200 import mesh as M, dist as D
n=int(sys.argv[1]) if len(sys.argv)>1 else 14
p=int(sys.argv[2]) if len(sys.argv)>2 else 2
mesh=M.kuhn_cube(n,p)
conn=mesh.conn.astype(np.int64); ne,nl=conn.shape; nb=mesh.n_bases
adj_off,adj=D.block_pattern_numpy(mesh.conn, nb)
deg=np.diff(adj_off)
# morton order of centroids
cen=mesh.vertices.mean(axis=1)
q=np.clip(((cen-cen.min(0))/(cen.max(0)-cen.min(0))*1023).astype(np.int64),0,1023)
def spread(v):
    r=np.zeros_like(v)
    for b in range(10): r|=((v>>b)&1)<<(3*b)
    return r
key=spread(q[:,0])|(spread(q[:,1])<<1)|(spread(q[:,2])<<2)
perm=np.argsort(key,kind='stable')
connp=conn[perm]
# pair key -> position in adj
pair_key=np.repeat(np.arange(nb),deg)*nb+adj
def pair_pos(b,a):
    return np.searchsorted(pair_key,b*nb+a)
# interior elements only (skip boundary effects): use all
def study(B):
    tot_contrib=0; tot_unique=0; tot_sec_now=0; tot_sec_comb=0; smem=[]
    for c0 in range(0,ne-B+1,B*7):  # sample clusters
        c=connp[c0:c0+B]
        b=np.repeat(c,nl,axis=1).reshape(-1)   # column node
        a=np.tile(c,(1,nl)).reshape(-1)        # row node
        pos=pair_pos(b,a)
        k=pos-adj_off[b]
        # scalar index of (row (a,m), col (b,nn)) = 9*off[b] + nn*3*deg[b] + 3k + m
        base=9*adj_off[b].astype(np.int64)+3*k
        # now: per element, per (j,nn): instruction covers 10 runs -> distinct sectors per element
        sec_now=0
        idx_all=[]
        for nn in range(3):
            idx=(base+nn*3*deg[b])[:,None]+np.arange(3)[None,:]   # [contrib,3]
            idx_all.append(idx)
        idx_all=np.stack(idx_all,1)  # [contrib, nn, m]
        per_el=idx_all.reshape(B,nl,nl,3,3)  # e,i,j,nn,m  (b=c[e,j] repeated? check ordering)
        # ordering: b=repeat(c,nl) -> for element e: b index = i*nl+j? repeat(c,nl,axis=1): [c0,c0,..(nl times),c1,...] so first index is column node index jj, second (tile) is row ii
        for e in range(B):
            for jj in range(nl):
                for nn in range(3):
                    s=np.unique(per_el[e,jj,:,nn,:]//4)
                    sec_now+=s.size
        uniq=np.unique(idx_all.reshape(-1))
        sec_comb=np.unique(uniq//4).size
        tot_contrib+=B*9*nl*nl; tot_unique+=uniq.size; tot_sec_now+=sec_now; tot_sec_comb+=sec_comb
        smem.append(uniq.size*8)
    print(f"B={B:3d}: contributions/unique doubles = {tot_contrib/tot_unique:.2f}, sector-ops/element now {tot_sec_now/(len(smem)*B):.0f} -> combined {tot_sec_comb/(len(smem)*B):.0f} ({tot_sec_now/tot_sec_comb:.2f}x), smem for unique doubles {np.mean(smem)/1024:.0f} KB (max {np.max(smem)/1024:.0f})")
for B in ((1,2,4,6,8,12,16,24,48) if p==2 else (1,6,12,24,48,96,192,384)):
    study(B)
