"""Shared-SM form of the asynchronous step (AZB_ASYNC_SHARED) against the whole-SM form: simulations/s of K timed steps.
usage: shared_probe.py N B steps mode[,mode...]   (mode = 'shared' or a whole-SM worker count)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("PROBE_PROF"):
    os.environ["AZB_LIB"] = os.path.join(ROOT, "azdopt_b200", "lib", "libazb_prof.so")
from azdopt_b200 import capi
import ctypes as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 19
b = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 400
modes = sys.argv[4].split(",") if len(sys.argv) > 4 else ["shared", "20"]
p, m = capi.generate_roots(0, 0, b, n)
for mode in modes:
    w = capi.ASYNC_SHARED if mode == "shared" else int(mode)
    cfg = capi.default_config(n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps + 40, async_workers=w)
    with capi.Handle(cfg) as h:
        h.set_counter_mode(False)
        h.mlp_init(1)
        h.set_roots(p, m)
        h.init_trees()
        h.step(16)
        h.reset_counters()
        ms, _ = h.step_timed(steps)
        k = h.counters()
        print(f"N={n} B={b} mode={mode:>6s} G={os.environ.get('AZB_ASYNC_GROUP', '-')}: {ms:9.2f} ms for {steps} steps ({ms/steps*1e3:7.1f} us/step)  "
              f"{k['n_live']/(ms*1e-3)/1e6:7.2f} M sims/s  argmin {h.argmin()['eval']:.5f}", flush=True)
        if os.environ.get("PROBE_PROF"):
            L = capi.lib()
            L.azb_debug_async.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
            d = (C.c_uint64 * 32)()
            L.azb_debug_async(h._h, d)
            tiles = max(d[4], 1)
            us = lambda c: c / tiles / 1965.0
            print(f"      per tile and member (us): acquire {us(d[0]):6.1f} wait-layer {us(d[2]):6.1f} busy {us(d[3]):6.1f} | producer wait-empty {us(d[1]):6.1f} MMA wait-full {us(d[6]):6.1f} | warp 2: wait-acc {us(d[11]):6.1f} epilogue {us(d[13]):6.1f} fence+barrier {us(d[12]):6.1f} | member-tiles {d[4]} rows real {d[15] >> 32} dummy {d[15] & 0xffffffff}", flush=True)
            print(f"      trees: per step: walking mean {d[16]/b/steps/1965:.1f} us, slowest tree {d[17]/steps/1965:.1f} us; waiting for priors mean {d[19]/b/steps/1965:.1f} us, max tree {d[20]/steps/1965:.1f} us", flush=True)
