"""Workload counters per tree-step for hash vs MLP priors at B=4096 (full counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi
n, b, steps = 19, 4096, 400
p, m = capi.generate_roots(0, 0, b, n)
for mode in ("hash", "mlp"):
    kw = dict(max_steps=steps + 8)
    if mode == "hash":
        kw.update(prior_mode=capi.PRIOR_HASH)
    else:
        kw.update(prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC)
    with capi.Handle(capi.default_config(n, b, **kw)) as h:
        if mode == "mlp":
            h.mlp_init(1)
        h.set_roots(p, m)
        h.init_trees()
        h.step(16)
        h.reset_counters()
        ms, _ = h.step_timed(steps - 16)
        k = h.counters()
        per = b * (steps - 16)
        print(mode, f"us/step={ms/(steps-16)*1e3:.1f}", {kk: round(v / per, 2) for kk, v in k.items()})
        sz = [h.tree_sizes(i) for i in range(0, b, 64)]
        print("   mean nodes/arcs/preds", [sum(s[j] for s in sz) / len(sz) for j in range(3)], "max", [max(s[j] for s in sz) for j in range(3)])
