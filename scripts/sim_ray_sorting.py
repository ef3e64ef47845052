#!/usr/bin/env python
"""Design exploration (not part of the product or the tests): how much lane efficiency ray ordering
could buy in the lock-step traversal loop, replayed on real C2 path segments with their origins and
directions (tests/hostsim: hs_traversal_events_rays).  Results:
profiles/experiments/r01_lane_efficiency_simulations.txt.  Run from the repo root; ~10 minutes."""
import ctypes as C, numpy as np, random, sys
sys.path.insert(0,'/root/repo')
from raytracer_go_b200 import scenes
from raytracer_go_b200 import api
hs = C.CDLL('/root/repo/tests/hostsim/libhostsim.so'); hs.hs_traversal_events_rays.restype = C.c_int64
s = scenes.random_scene(); desc, keep = s.to_desc()
cam = api.camera_from_options(scenes.camera_options(1200, 8))
N = 30_000_000; NR=400_000
tok = np.zeros(N, np.int8); rr = np.zeros((NR,7), np.float32)
rays=[]; meta=[]
# a contiguous image block: 40 rows in the middle (so that neighbouring pixels are present, as in a wavefront batch)
for r in range(300, 340, 1):
    nr = C.c_int64()
    n = hs.hs_traversal_events_rays(C.byref(desc), C.byref(cam), C.c_uint64(7), 8, C.c_int64(r*cam.width), C.c_int64(cam.width), 4, tok.ctypes.data_as(C.c_void_p), C.c_int64(N), rr.ctypes.data_as(C.c_void_p), C.c_int64(NR), C.byref(nr))
    t = tok[:n]; ends = np.nonzero(t==0)[0]; start=0
    for i,e in enumerate(ends):
        rays.append(t[start:e].copy()); start=e+1
    meta.append(rr[:nr.value].copy())
meta=np.concatenate(meta); print('rays', len(rays), len(meta))
CI, CS, CL = 50, 35, 10
def ideal(r): return int(r[r>0].sum())*CI + int(-r[r<0].sum())*CS + int((r<0).sum())*CL
ID = np.array([ideal(r) for r in rays])
def sim_warp(idx):
    lanes=[[rays[i],0] for i in idx]; cost=0
    while True:
        act=[l for l in lanes if l[1]<len(l[0])]
        if not act: break
        m=0
        for l in act:
            r,p=l
            if r[p]>0: m=max(m,int(r[p])); l[1]+=1
        cost+=m*CI
        ms=0; anyl=False
        for l in act:
            r,p=l
            if p<len(r) and r[p]<0: ms=max(ms,int(-r[p])); l[1]+=1; anyl=True
        if anyl: cost+=ms*CS+CL
    return cost
def eff(order, nw=1500):
    tot=0; ide=0
    starts = np.random.default_rng(0).choice(len(order)//32-1, nw, replace=False)*32
    for st in starts:
        idx = order[st:st+32]; tot+=sim_warp(idx); ide+=ID[idx].sum()
    return ide/(32*tot)
n=len(rays)
sec = meta[:,6]>0
print('baseline random order (all rays):', round(eff(np.random.default_rng(1).permutation(n)),3))
print('generation order (path order):', round(eff(np.arange(n)),3))
# secondary rays only, random vs sorted
sidx = np.nonzero(sec)[0]
print('secondary random:', round(eff(np.random.default_rng(2).permutation(sidx)),3))
o = meta[:,:3]; d = meta[:,3:6]; dn = d/np.linalg.norm(d,axis=1,keepdims=True)
def keys(cell, dirbits):
    g = np.floor((o - o.min(0)) / cell).astype(np.int64); g = np.clip(g,0,1023)
    if dirbits==3:
        oc = (dn[:,0]>0).astype(np.int64) | ((dn[:,1]>0).astype(np.int64)<<1) | ((dn[:,2]>0).astype(np.int64)<<2)
    else:
        # finer direction bins: quantise each component to 4 levels
        q = np.clip(((dn+1)*2).astype(np.int64),0,3); oc = q[:,0] | (q[:,1]<<2) | (q[:,2]<<4)
    return ((g[:,0]*1024 + g[:,1])*1024 + g[:,2])*64 + oc
for cell in (4.0, 2.0, 1.0, 0.5):
    for db in (3,6):
        k = keys(cell, db)
        order = sidx[np.argsort(k[sidx], kind='stable')]
        print(f'secondary sorted cell={cell} dirbits={db}:', round(eff(order),3))
k = keys(1.0,6); order=np.argsort(k, kind='stable'); print('all rays sorted cell=1 dir6:', round(eff(order),3))
pidx=np.nonzero(~sec)[0]; print('primary in generation order:', round(eff(pidx),3))

# ---- CTA-local sorting: a batch = one segment of each of B consecutive paths (steady-state pool) ----
depth = meta[:,6].astype(int)
path_start = np.nonzero(depth==0)[0]
path_len = np.diff(np.append(path_start, n))
rng = np.random.default_rng(5)
def batch_eff(B, sort, cell=2.0, db=6, nb=60):
    tot=0; ide=0
    kk = keys(cell, db) if sort else None
    for b in range(nb):
        p0 = rng.integers(0, len(path_start)-B)
        ps = path_start[p0:p0+B]; pl = path_len[p0:p0+B]
        idx = ps + (rng.random(B)*pl).astype(int)
        if sort: idx = idx[np.argsort(kk[idx], kind='stable')]
        for w in range(0, B, 32):
            g = idx[w:w+32]; tot += sim_warp(g); ide += ID[g].sum()
    return ide/(32*tot)
for B in (128, 512, 2048, 8192):
    print('batch', B, 'unsorted', round(batch_eff(B, False),3), 'sorted c2 d6', round(batch_eff(B, True),3), 'sorted c4 d3', round(batch_eff(B, True, 4.0, 3),3))
# sort by depth==0 first then key (primaries together)
def batch_eff2(B, nb=60):
    tot=0; ide=0; kk = keys(2.0,6)
    for b in range(nb):
        p0 = rng.integers(0, len(path_start)-B)
        ps = path_start[p0:p0+B]; pl = path_len[p0:p0+B]
        idx = ps + (rng.random(B)*pl).astype(int)
        key2 = np.where(depth[idx]==0, -1, kk[idx])
        idx = idx[np.argsort(key2, kind='stable')]
        for w in range(0, B, 32):
            g = idx[w:w+32]; tot += sim_warp(g); ide += ID[g].sum()
    return ide/(32*tot)
for B in (512, 2048, 8192): print('batch', B, 'primaries first + sorted', round(batch_eff2(B),3))
