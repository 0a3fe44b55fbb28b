"""Lock step vs the asynchronous kernel: simulations/s of K timed steps for several worker counts."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("PROBE_PROF"):  # cycle counters need the -DAZB_PROFILE flavour
    os.environ["AZB_LIB"] = os.path.join(ROOT, "azdopt_b200", "lib", "libazb_prof.so")
from azdopt_b200 import capi
import ctypes as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
workers = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 8, 16, 24, 32]
p, m = capi.generate_roots(0, 0, b, n)
for w in workers:
    cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps + int(os.environ.get("PROBE_SLACK", "40")), async_workers=w)
    with capi.Handle(cfg) as h:
        h.set_counter_mode(False)
        h.mlp_init(1)
        h.set_roots(p, m)
        h.init_trees()
        h.step(16)
        h.reset_counters()
        ms, _ = h.step_timed(steps)
        k = h.counters()
        print(f"N={n} B={b} workers={w:3d}: {ms:9.2f} ms for {steps} steps ({ms/steps*1e3:7.1f} us/step)  "
              f"{k['n_live']/(ms*1e-3)/1e6:7.2f} M sims/s  argmin {h.argmin()['eval']:.5f}", flush=True)
        if w and os.environ.get("PROBE_PROF"):
            L = capi.lib()
            L.azb_debug_async.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
            d = (C.c_uint64 * 32)()
            L.azb_debug_async(h._h, d)
            tiles = max(d[4], 1)
            us = lambda c: c / tiles / 1965.0
            print(f"      per tile (us): producer acquire {us(d[0]):6.1f} wait-empty {us(d[1]):6.1f} wait-layer {us(d[2]):6.1f} tile {us(d[3]):6.1f} | "
                  f"MMA wait-full {us(d[6]):6.1f} wait-acc {us(d[7]):6.1f} | epilogue wait-acc {us(d[11]):6.1f} busy {us(d[13]):6.1f} (tmem-ld {us(d[10]):6.1f} fence {us(d[12]):6.1f}) | "
                  f"tiles {d[4]} rows real {d[15] >> 32} dummy {d[15] & 0xffffffff}", flush=True)
            print(f"      trees: kernel {d[18]/1965e3:.1f} ms; cycles advancing a tree: mean {d[16]/b/1965e3:.1f} ms, max {d[17]/1965e3:.1f} ms "
                  f"(per step: mean {d[16]/b/steps/1965:.1f} us, slowest tree {d[17]/steps/1965:.1f} us); waiting for priors per step: "
                  f"mean {d[19]/b/steps/1965:.1f} us, max tree {d[20]/steps/1965:.1f} us; max (run+wait) {d[21]/1965e3:.1f} ms; "
                  f"pick-up latency (answer flag -> step starts): mean {d[24]/max(d[26],1)/1e3:.2f} us, max {d[25]/1e3:.1f} us over {d[26]} pick-ups", flush=True)
