"""Latency probe: ms per step of the search kernel alone (hash priors) for several batch sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
me = int(sys.argv[2]) if len(sys.argv) > 2 else 0
for b in (1, 128, 4096, 16384, 65536):
    cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_HASH, max_steps=400, max_episodes=me)
    p, m = capi.generate_roots(0, 0, b, n)
    with capi.Handle(cfg) as h:
        h.set_roots(p, m)
        h.init_trees()
        h.step(50)
        h.reset_counters()
        ms, _ = h.step_timed(300)
        k = h.counters()
        ep = k["n_arc"] / max(k["n_live"] + k["n_noop"], 1)
        print(f"me={me} B={b:6d} us/step={ms/300*1e3:8.1f} episodes/tree-step={ep:.2f} ins/tree-step={k['n_ins']/(b*300):.2f} "
              f"sel/tree-step={k['n_sel']/(b*300):.1f} cn/tree-step={k['n_cn']/(b*300):.1f} sims/s={k['n_live']/(ms*1e-3):.3e}")
