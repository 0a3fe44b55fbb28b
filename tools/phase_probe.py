"""Per-phase lane-0 cycles of the search kernel (needs lib/libazb_prof.so: azdopt_b200.build.build_profile_flavour())."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["AZB_LIB"] = os.environ.get("PROF_LIB", os.path.join(ROOT, "azdopt_b200", "lib", "libazb_prof.so"))
from azdopt_b200 import capi  # noqa: E402

PH = ["sel", "cur", "probe", "arc", "cascade", "cost", "insert", "reset", "add", "pack", "load", "store"]
n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
aw = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # > 0: the asynchronous kernel with that many worker SMs (MLP priors)
L = capi.lib()
L.azb_debug_phase_cycles.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
for b in ((int(os.environ.get("PROBE_B", "4096")),) if aw else (1, 4096)):
    cfg = (capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=400, async_workers=aw) if aw
           else capi.default_config(n, b, prior_mode=capi.PRIOR_HASH, max_steps=400))
    p, m = capi.generate_roots(0, 0, b, n)
    with capi.Handle(cfg) as h:
        if aw:
            h.mlp_init(1)
        h.set_roots(p, m)
        h.init_trees()
        h.step(50)
        h.reset_counters()
        buf = (C.c_uint64 * 16)()
        L.azb_debug_phase_cycles(h._h, buf)
        steps = 200
        ms, _ = h.step_timed(steps)
        k = h.counters()
        L.azb_debug_phase_cycles(h._h, buf)
        tot = sum(buf[i] for i in range(12))
        per = b * steps
        print(f"B={b} us/step={ms / steps * 1e3:.1f}  lane-0 cycles per tree-step: total {tot / per:.0f}")
        ev = {"sel": k["n_sel"], "cur": k["n_cur"], "probe": k["n_probe"], "arc": k["n_hit"], "cascade": k["n_hit"] + k["n_term"],
              "cost": k["n_ins"], "insert": k["n_ins"], "reset": k["n_reset"], "add": k["n_live"], "pack": k["n_live"],
              "load": per, "store": per}
        print(f"   cost split: init {buf[12] / max(k['n_ins'],1):.0f}  dp {buf[13] / max(k['n_ins'],1):.0f}  coef {buf[14] / max(k['n_ins'],1):.0f}  laguerre+polish {buf[15] / max(k['n_ins'],1):.0f} cyc/eval")
        for i, name in enumerate(PH):
            print(f"   {name:8s} {buf[i] / per:9.0f} cyc/tree-step  {buf[i] / max(ev[name], 1):9.0f} cyc/event  ({ev[name] / per:.2f} events/tree-step)")
