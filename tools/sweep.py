"""Rollout throughput sweep (BASELINE configs[3], [4]): search step with the tensor-core MLP, several batch sizes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024, 4096, 16384, 65536]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 200
mode = sys.argv[4] if len(sys.argv) > 4 else "async"  # "async" (bench.py's default search kernel) or "lock"
out = []
for b in sizes:
    aw = 0 if mode == "lock" or b < 1024 else capi.ASYNC_AUTO  # the library's layout for the root count
    cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps + 24, async_workers=aw)
    p, m = capi.generate_roots(0, 0, b, n)
    with capi.Handle(cfg) as h:
        h.mlp_init(1)
        h.set_counter_mode(False)
        h.set_roots(p, m)
        h.init_trees()
        h.step(16)
        h.reset_counters()
        ms, _ = h.step_timed(steps)
        k = h.counters()
        rec = dict(n=n, roots=b, steps=steps, async_workers=int(h.cfg.async_workers), us_per_step=ms / steps * 1e3, sims_per_s=k["n_live"] / (ms * 1e-3),
                   cost_evals_per_s=k["n_ins"] / (ms * 1e-3), hbm_mb=h.device_bytes() / 1e6)
        out.append(rec)
        print(json.dumps(rec), flush=True)
