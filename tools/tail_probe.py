"""Who is the slowest tree of a launch?  (profile flavour)"""
import ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["AZB_LIB"] = os.path.join(ROOT, "azdopt_b200", "lib", "libazb_prof.so")
from azdopt_b200 import capi
n, b = 19, 4096
L = capi.lib()
L.azb_debug_tree_prof.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32]
mode = sys.argv[1] if len(sys.argv) > 1 else "mlp"
kw = dict(max_steps=400)
if mode == "hash":
    kw.update(prior_mode=capi.PRIOR_HASH)
else:
    kw.update(prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC)
p, m = capi.generate_roots(0, 0, b, n)
with capi.Handle(capi.default_config(n, b, **kw)) as h:
    if mode != "hash":
        h.mlp_init(1)
    h.set_roots(p, m)
    h.init_trees()
    for warm in (20, 100, 300):
        h.step(warm - h.counters()["n_live"] * 0 - (0 if warm == 20 else 0)) if False else None
    done = 0
    for target in (20, 100, 300):
        h.step(target - done - 1); done = target - 1
        h.step(1); done += 1
        buf = np.zeros((b, 4), dtype=np.uint32)
        L.azb_debug_tree_prof(h._h, buf.ctypes.data_as(C.POINTER(C.c_uint32)), b)
        cyc = buf[:, 0].astype(np.int64)
        order = np.argsort(-cyc)
        print(f"step {target}: mean cycles {cyc.mean():.0f}  p50 {np.median(cyc):.0f} p99 {np.percentile(cyc,99):.0f} max {cyc.max()}  (mean episodes {buf[:,1].mean():.2f}, evals {buf[:,2].mean():.2f}, sqrt terms {buf[:,3].mean():.0f})")
        for t in order[:6]:
            print(f"     tree {t}: cycles {buf[t,0]} episodes {buf[t,1]} cost evals {buf[t,2]} sqrt terms {buf[t,3]}")
