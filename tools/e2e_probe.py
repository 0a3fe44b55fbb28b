"""Where the end-to-end epoch's time goes (bench.py's e2e leg): wall time of each C-ABI call, synchronised after each."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi

n, b, steps = 19, int(sys.argv[1]) if len(sys.argv) > 1 else 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 800
cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps + 8, async_workers=20)
p, m = capi.generate_roots(0, 0, b, n)
with capi.Handle(cfg) as h:
    h.mlp_init(1)
    h.set_counter_mode(False)
    for rep in range(3):
        t = [time.perf_counter()]
        h.set_roots(p, m); h.counters(); t.append(time.perf_counter())
        h.init_trees(); h.counters(); t.append(time.perf_counter())
        h.step(steps, cap=steps); h.counters(); t.append(time.perf_counter())
        a = h.argmin(); h.counters(); t.append(time.perf_counter())
        d = [1e3 * (t[i + 1] - t[i]) for i in range(4)]
        print(f"rep {rep}: set_roots {d[0]:.2f} ms  init_trees {d[1]:.2f} ms  step({steps}) {d[2]:.2f} ms  argmin {d[3]:.2f} ms  total {sum(d):.2f} ms")
