#!/usr/bin/env python
"""CPU-only probe (oracle): far-range grazing.  The reference's vertex / edge quadratics (collision_detect.rs:44-112)
lose about 3e-7 * D^2 of miss distance at range D (cancellation in b^2 - 4ac), so for a sphere that passes a far
triangle at distance 1.0002 .. 1 + 3e-7 D^2 the reference can report a hit although the path never enters the
triangle's box inflated by 1.0001 -- pairs the broadphase predicate P does not offer.  This script aims such rays
(miss direction along a coordinate axis, where the box is tight) and counts them per range.  See DESIGN.md section 2.
"""
import sys, numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'xp-game-engine_b200')]
from oracle import xp_oracle as oracle
rng=np.random.default_rng(4)
tot=0; viol=0; ex=None
for D in (30.0, 100.0, 300.0, 1000.0, 3000.0):
    v_found=0; n_hit=0
    for rep in range(40):
        # small triangle at the origin region, axis-aligned-ish miss direction
        tri=(rng.normal(0,0.05,(3,3))).astype(np.float32)
        ns=4000
        axis=rng.integers(0,3,ns); sgn=rng.choice([-1.0,1.0],ns)
        band=6e-8*D*D/2
        d=1.0+rng.uniform(-0.2*band, 1.5*band+2e-4, ns)
        off=np.zeros((ns,3)); off[np.arange(ns),axis]=sgn*d                 # closest approach point relative to vertex 0
        dirs=rng.normal(0,1,(ns,3)); dirs[np.arange(ns),axis]=0; dirs/= np.linalg.norm(dirs,axis=1,keepdims=True)
        speed=rng.uniform(0.05,2.0,(ns,1))
        c=tri[0].astype(np.float64)+off-dirs*D
        resps=oracle.make_responses(c.astype(np.float32),(dirs*speed).astype(np.float32))
        want=oracle.detect_batch(resps,tri[None]); pi,pj=oracle.broadphase_pairs(resps,tri[None])
        inP=np.zeros(ns,bool); inP[pi]=True
        bad=want[1].astype(bool)&~inP
        n_hit+=int(want[1].sum()); v_found+=int(bad.sum())
    print(f"D={D:7.0f}: band eps*D^2/2 = {6e-8*D*D/2:.2e}; hits {n_hit}; hits outside P: {v_found}")
