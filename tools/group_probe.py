"""Premise check: G independent handles (B/G trees each, own stream) driven round-robin on one GPU, hash priors."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi

n, b, steps = 19, 4096, 300
for g in (1, 4, 16, 32, 64):
    bg = b // g
    hs = []
    for i in range(g):
        cfg = capi.default_config(n, bg, prior_mode=capi.PRIOR_HASH, max_steps=steps + 60, first_root=i * bg)
        h = capi.Handle(cfg)
        p, m = capi.generate_roots(0, i * bg, bg, n)
        h.set_counter_mode(False)
        h.set_roots(p, m)
        h.init_trees()
        h.step(50)
        hs.append(h)
    t0 = time.perf_counter()
    for s in range(steps):
        for h in hs:
            h.step_enqueue(1)
    t1 = time.perf_counter()
    for h in hs:
        h.step(0)
    t2 = time.perf_counter()
    live = sum(h.counters()["n_live"] for h in hs)
    print(f"G={g:3d} trees/group={bg:5d} enqueue {1e6*(t1-t0)/steps:7.1f} us/step  total {1e6*(t2-t0)/steps:7.1f} us/step")
    for h in hs:
        h.close()
