"""What predicts a tree's walking time?  Profile flavour, asynchronous kernel: per-tree walking cycles against the
root's permitted-action count, its legal-action count and the tree's final size."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["AZB_LIB"] = os.path.join(ROOT, "azdopt_b200", "lib", "libazb_prof.so")
import ctypes as C
import numpy as np
from azdopt_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
w = int(sys.argv[4]) if len(sys.argv) > 4 else 20
p, m = capi.generate_roots(0, 0, b, n)
k = np.array([sum(bin(int(x)).count("1") for x in row) for row in m])
cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=2 * steps + 40, async_workers=w)
L = capi.lib()
L.azb_debug_tree_prof.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32]
with capi.Handle(cfg) as h:
    h.set_counter_mode(False)
    h.mlp_init(1)
    h.set_roots(p, m)
    h.init_trees()
    legal = np.array([h.tree_sizes(i)[2] for i in range(b)])  # root predictions = legal actions of the root
    runs = []
    for rnd in range(2):
        ms, _ = h.step_timed(steps)
        buf = np.zeros((b, 4), dtype=np.uint32)
        L.azb_debug_tree_prof(h._h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), b)
        runs.append(buf.copy())
        print(f"round {rnd}: {ms / steps * 1e3:.1f} us/step")
    sizes = np.array([h.tree_sizes(i) for i in range(b)])
for rnd, buf in enumerate(runs):
    run, wait = buf[:, 0].astype(float) * 1024 / 1965 / steps, buf[:, 1].astype(float) * 1024 / 1965 / steps
    print(f"round {rnd}: walking us/step mean {run.mean():.1f} p50 {np.median(run):.1f} p90 {np.percentile(run, 90):.1f} "
          f"p99 {np.percentile(run, 99):.1f} max {run.max():.1f}; waiting mean {wait.mean():.1f} max {wait.max():.1f}; "
          f"run+wait max {(run + wait).max():.1f}")
    for name, x in (("k (permitted at root)", k), ("legal actions at root", legal), ("final nodes", sizes[:, 0]),
                    ("final predictions", sizes[:, 2])):
        print(f"   corr(walking, {name}) = {np.corrcoef(run, x)[0, 1]:.3f}")
    order = np.argsort(-run)
    print("   slowest 8 trees:", [(int(i), round(run[i], 1), int(k[i]), int(legal[i])) for i in order[:8]])
    if rnd == 1:
        print(f"   corr(walking round 0, walking round 1) = {np.corrcoef(runs[0][:, 0], runs[1][:, 0])[0, 1]:.3f}")
