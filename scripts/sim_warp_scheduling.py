#!/usr/bin/env python
"""Design exploration (not part of the product or the tests): replays the traversal event traces of
real C2 path segments (tests/hostsim: hs_traversal_events) under different warp-scheduling policies
and prints the lane efficiency each would reach.  Results: profiles/experiments/r01_lane_efficiency_simulations.txt.
Run from the repo root after `make -C tests/hostsim`; takes a few minutes (pure Python)."""
import ctypes as C
import os
import pickle
import random
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
TRACES = "/tmp/rays_events.pkl"


def record_traces():
    from raytracer_go_b200 import scenes
    from raytracer_go_b200 import api
    hs = C.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/hostsim/libhostsim.so"))
    hs.hs_traversal_events.restype = C.c_int64
    s = scenes.random_scene()
    desc, keep = s.to_desc()
    cam = api.camera_from_options(scenes.camera_options(1200, 8))
    n_tok = 40_000_000
    tok = np.zeros(n_tok, np.int8)
    out = []
    for r in np.linspace(0, cam.height - 1, 24).astype(int):
        nr = C.c_int64()
        n = hs.hs_traversal_events(C.byref(desc), C.byref(cam), C.c_uint64(7), 8, C.c_int64(int(r) * cam.width),
                                   C.c_int64(cam.width), 4, tok.ctypes.data_as(C.c_void_p), C.c_int64(n_tok), C.byref(nr))
        t = tok[:n].copy()
        start = 0
        for e in np.nonzero(t == 0)[0]:
            out.append(t[start:e])
            start = e + 1
    pickle.dump(out, open(TRACES, "wb"))


if not os.path.exists(TRACES):
    record_traces()
rays = pickle.load(open('/tmp/rays_events.pkl','rb'))
rays = [r for r in rays if len(r)]
random.seed(1)
CI, CS, CL = 50, 35, 10   # inner step, per sphere test, leaf overhead
def ideal(rs): return sum(int(r[r>0].sum())*CI + int(-r[r<0].sum())*CS + int((r<0).sum())*CL for r in rs)

def sim_whilewhile(pool_rays, K, dynamic):
    """pool of 32*K rays; dynamic fetch when lane finishes (checked once per outer iteration)."""
    P = len(pool_rays); nxt = 0
    lanes = [None]*32  # (ray, pos)
    cost = 0
    # static: lane i gets rays i, i+32,... sequentially (no fetch from others)
    queues = [list(range(i, P, 32)) for i in range(32)]
    while True:
        # fetch
        for i in range(32):
            if lanes[i] is None:
                if dynamic:
                    if nxt < P: lanes[i] = [pool_rays[nxt], 0]; nxt += 1
                else:
                    if queues[i]: lanes[i] = [pool_rays[queues[i].pop(0)], 0]
        act = [l for l in lanes if l is not None]
        if not act: break
        cost += 8  # fetch check overhead
        # inner phase: each active lane runs its next positive run (if token positive)
        m = 0
        for l in act:
            r,pos = l
            if pos < len(r) and r[pos] > 0:
                m = max(m, int(r[pos])); l[1] += 1
        cost += m*CI
        # leaf phase
        ms = 0; any_leaf=False
        for l in act:
            r,pos = l
            if pos < len(r) and r[pos] < 0:
                ms = max(ms, int(-r[pos])); l[1] += 1; any_leaf=True
        if any_leaf: cost += ms*CS + CL
        for i in range(32):
            l = lanes[i]
            if l is not None and l[1] >= len(l[0]): lanes[i] = None
    return cost

def sim_ifif(pool_rays, dynamic=True):
    P=len(pool_rays); nxt=0; lanes=[None]*32; cost=0
    while True:
        for i in range(32):
            if lanes[i] is None and nxt < P: lanes[i]=[pool_rays[nxt],0,0]; nxt+=1   # ray,pos,sub
        act=[l for l in lanes if l is not None]
        if not act: break
        cost += 8
        anyI=False; ms=0; anyL=False
        for l in act:
            r,pos,sub=l
            if r[pos]>0:
                anyI=True; l[2]+=1
                if l[2]>=r[pos]: l[1]+=1; l[2]=0
            else:
                anyL=True; ms=max(ms,int(-r[pos])); l[1]+=1
        cost += (CI if anyI else 0) + ((ms*CS+CL) if anyL else 0)
        for i in range(32):
            l=lanes[i]
            if l is not None and l[1]>=len(l[0]): lanes[i]=None
    return cost

def sim_postpone(pool_rays, nleaf_wait=1, thresh=0.5):
    """speculative: a lane that reaches a leaf keeps it pending and waits; inner loop continues while
    > thresh of busy lanes are still descending; then leaf phase processes all pending leaves."""
    P=len(pool_rays); nxt=0; lanes=[None]*32; cost=0
    while True:
        for i in range(32):
            if lanes[i] is None and nxt < P: lanes[i]=[pool_rays[nxt],0,0]; nxt+=1
        act=[l for l in lanes if l is not None]
        if not act: break
        cost += 8
        # inner loop: step lanes whose current token is positive; stop when fraction of descending lanes < thresh
        while True:
            desc=[l for l in act if l[1]<len(l[0]) and l[0][l[1]]>0]
            if len(desc)==0 or len(desc) < thresh*len(act): break
            cost += CI
            for l in desc:
                l[2]+=1
                if l[2]>=l[0][l[1]]: l[1]+=1; l[2]=0
        ms=0; anyL=False
        for l in act:
            r,pos,sub=l
            if pos<len(r) and r[pos]<0:
                anyL=True; ms=max(ms,int(-r[pos])); l[1]+=1
        if anyL: cost += ms*CS+CL
        for i in range(32):
            l=lanes[i]
            if l is not None and l[1]>=len(l[0]): lanes[i]=None
    return cost

def run(name, f, K, trials=300):
    tot=0; ide=0
    for t in range(trials):
        pr = random.sample(rays, 32*K)
        tot += f(pr); ide += ideal(pr)
    print(f"{name:40s} K={K} efficiency {32*0+ide/ (32*tot):.3f}" )
for K in (1,2,4):
    run('whilewhile static', lambda pr: sim_whilewhile(pr,K,False), K)
    run('whilewhile dynamic', lambda pr: sim_whilewhile(pr,K,True), K)
for K in (1,4):
    run('if-if dynamic', lambda pr: sim_ifif(pr), K)
for K in (4,):
    for th in (0.25,0.5,0.75):
        run(f'postpone thresh {th}', lambda pr: sim_postpone(pr,1,th), K)

def sim_spec(pool_rays, M=1, thresh=0.5, fetch_every_iter=True):
    """speculative traversal: a lane may keep up to M untested leaves and go on descending; the warp
    runs inner steps while >= thresh of its busy lanes can step, then tests pending leaves."""
    P=len(pool_rays); nxt=0; lanes=[None]*32; cost=0
    def fetch():
        nonlocal nxt
        for i in range(32):
            if lanes[i] is None and nxt < P: lanes[i]=[pool_rays[nxt],0,0,[]]; nxt+=1  # ray,pos,sub,pending
    while True:
        fetch()
        act=[l for l in lanes if l is not None]
        if not act: break
        cost += 8
        while True:
            # lanes whose current token is a leaf move it to pending if room
            for l in act:
                r=l[0]
                while l[1]<len(r) and r[l[1]]<0 and len(l[3])<M:
                    l[3].append(int(-r[l[1]])); l[1]+=1
            desc=[l for l in act if l[1]<len(l[0]) and l[0][l[1]]>0]
            if len(desc)==0 or len(desc) < thresh*len(act): break
            cost += CI
            for l in desc:
                l[2]+=1
                if l[2]>=l[0][l[1]]: l[1]+=1; l[2]=0
        # leaf phase: up to M passes
        while True:
            pend=[l for l in act if l[3]]
            if not pend: break
            ms=max(l[3][0] for l in pend)
            cost += ms*CS+CL
            for l in pend: l[3].pop(0)
        for i in range(32):
            l=lanes[i]
            if l is not None and l[1]>=len(l[0]) and not l[3]: lanes[i]=None
    return cost
for K in (1,4,8):
    for M in (1,2):
        for th in (0.35,0.5,0.65):
            run(f'spec M={M} thresh {th}', lambda pr: sim_spec(pr,M,th), K, trials=150)

