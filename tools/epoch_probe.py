"""The example's whole epoch loop on the device (04-c21-tree.rs:140-208): K fused steps, par_update_model,
par_reset_trees — wall time of each part, loss and best cost per epoch."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from azdopt_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
epochs = int(sys.argv[3]) if len(sys.argv) > 3 else 3
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 800
aw = int(sys.argv[5]) if len(sys.argv) > 5 else (20 if b <= 4096 else 48 if b < 16384 else 32)  # 0 = lock step
cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps, async_workers=aw)
p, m = capi.generate_roots(0, 0, b, n)
with capi.Handle(cfg) as h:
    h.set_counter_mode(False)
    h.mlp_init(1)
    h.set_roots(p, m)
    h.init_trees()
    print(f"N={n} B={b} async_workers={aw}: root argmin eval {h.argmin()['eval']:.5f}", flush=True)
    for e in range(1, epochs + 1):
        t0 = time.perf_counter(); n_imp, _ = h.step(steps); t1 = time.perf_counter()
        loss = h.update_model(200); t2 = time.perf_counter()
        h.reset_trees(1234); t3 = time.perf_counter()
        a = h.argmin()
        print(f"epoch {e}: rollout {1e3*(t1-t0):8.1f} ms ({n_imp} improvements)  update_model {1e3*(t2-t1):7.2f} ms  loss {loss:.6f}  "
              f"reset_trees {1e3*(t3-t2):7.2f} ms  best eval {a['eval']:.5f} (lambda_1 {a['lambda1']:.4f}, mu {a['mu']})", flush=True)
