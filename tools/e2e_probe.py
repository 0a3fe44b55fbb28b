"""Where the wall clock of a short epoch goes (the driver's bench call: 20 steps after 5 warm-up): every piece of the
e2e leg of bench.py timed on its own, synchronised (host buffers -> set_roots -> init_trees -> step(K) -> argmin)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from azdopt_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
p, m = capi.generate_roots(0, 0, b, n)
cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps + 140, async_workers=capi.ASYNC_AUTO)
with capi.Handle(cfg) as h:
    h.mlp_init(1)
    h.set_roots(p, m)
    h.init_trees()
    h.step(5)
    h.set_counter_mode(False)
    ms, _ = h.step_timed(steps)
    print(f"device-timed: {steps} steps {ms:.3f} ms ({ms/steps*1e3:.1f} us/step)")
    ms, _ = h.step_timed(100)
    print(f"device-timed: 100 more steps {ms:.3f} ms ({ms/100*1e3:.1f} us/step)")
    for rep in range(3):
        t = [time.perf_counter()]
        h.set_roots(p, m); t.append(time.perf_counter())
        h.init_trees(); t.append(time.perf_counter())
        k0 = h.counters()["n_live"]; t.append(time.perf_counter())
        h.step(steps, cap=steps); t.append(time.perf_counter())
        h.argmin(); t.append(time.perf_counter())
        k1 = h.counters()["n_live"]
        names = ["set_roots", "init_trees", "counters", f"step({steps})", "argmin"]
        tot = t[-1] - t[0]
        print("  ".join(f"{nm} {1e3*(t[i+1]-t[i]):.3f} ms" for i, nm in enumerate(names)), f"| total {tot*1e3:.3f} ms  {(k1-k0)/tot/1e6:.2f} M sims/s e2e")
