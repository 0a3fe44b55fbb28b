"""Bounded-episode launches: device time per launch and per step of the search kernel alone (hash priors)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
for me in (0, 1, 2, 3, 4, 6):
    cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_HASH, max_steps=400, max_episodes=me)
    p, m = capi.generate_roots(0, 0, b, n)
    with capi.Handle(cfg) as h:
        h.set_counter_mode(False)
        h.set_roots(p, m)
        h.init_trees()
        h.step(50)
        h.reset_counters()
        l0 = h.kernel_launches()
        ms, _ = h.step_timed(300)
        l1 = h.kernel_launches()
        k = h.counters()
        nl = l1 - l0
        print(f"me={me} B={b} us/step={ms/300*1e3:8.1f} launches={nl} us/launch={ms*1e3/nl:7.1f} sims/s={k['n_live']/(ms*1e-3):.3e}", flush=True)
