"""Small fixed workload for ncu: B roots, hash priors, `steps` lock-step launches of the search kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 120
me = int(sys.argv[3]) if len(sys.argv) > 3 else 0
cfg = capi.default_config(19, b, prior_mode=capi.PRIOR_HASH, max_steps=steps + 8, max_episodes=me)
p, m = capi.generate_roots(0, 0, b, 19)
with capi.Handle(cfg) as h:
    h.set_counter_mode(False)
    h.set_roots(p, m)
    h.init_trees()
    ms, _ = h.step_timed(steps)
    print(f"B={b} steps={steps} us/step={ms / steps * 1e3:.1f}")
