"""One asynchronous launch for ncu: B roots, MLP priors, library layout.  usage: ncu_async_target.py [roots] [steps] [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi

b = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
n = int(sys.argv[3]) if len(sys.argv) > 3 else 19
cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=3 * steps + 24, async_workers=capi.ASYNC_AUTO)
p, m = capi.generate_roots(0, 0, b, n)
with capi.Handle(cfg) as h:
    h.set_counter_mode(False)
    h.mlp_init(1)
    h.set_roots(p, m)
    h.init_trees()
    h.step(steps)      # launch 0 (warm-up)
    ms, _ = h.step_timed(steps)  # launch 1: the one to capture (-s 1 -c 1)
    print(f"N={n} B={b} steps={steps} us/step={ms / steps * 1e3:.1f} workers={int(h.cfg.async_workers)}")
