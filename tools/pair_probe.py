"""CTA pairs (AZB_ASYNC_PAIR=1): parity against the lock step on small batches, then simulations/s against the
single-CTA model workers.  usage: pair_probe.py [parity|speed|all] [roots,...] [workers,...] [steps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
if os.environ.get("PROBE_PROF"):
    os.environ["AZB_LIB"] = os.path.join(ROOT, "azdopt_b200", "lib", "libazb_prof.so")
from azdopt_b200 import capi
from test_gpu_async import _same
import ctypes as C

what = sys.argv[1] if len(sys.argv) > 1 else "all"
roots = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4096, 8192, 32768]
workers = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [20, 28, 36]
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 200


def mk(n, b, **kw):
    return capi.Handle(capi.default_config(n, b, **kw))


if what in ("parity", "all"):
    os.environ["AZB_ASYNC_PAIR"] = "1"
    for n, b, w, mode in [(19, 300, 4, capi.MLP_TC), (19, 129, 2, capi.MLP_TC), (19, 1, 2, capi.MLP_TC), (12, 77, 4, capi.MLP_TC),
                          (19, 1024, 8, capi.MLP_TC), (19, 200, 4, capi.MLP_TC3), (33, 160, 4, capi.MLP_TC), (64, 96, 4, capi.MLP_TC),
                          (19, 4096, 20, capi.MLP_TC)]:
        st = 36
        p, m = capi.generate_roots(5, 0, b, n)
        kw = dict(prior_mode=capi.PRIOR_MLP, mlp_mode=mode, max_steps=2 * st + 2)
        with mk(n, b, **kw) as lock, mk(n, b, async_workers=w, **kw) as asy:
            for h in (lock, asy):
                h.mlp_init(3)
                h.set_roots(p, m)
                h.init_trees()
            n1, log1 = lock.step(st, cap=256)
            n2, log2 = asy.step(st, cap=256)
            assert n1 == n2 and [tuple(x) for x in log1] == [tuple(x) for x in log2]
            _same(lock, asy, b)
            for h in (lock, asy):
                h.step(1)
                h.step(st)
            _same(lock, asy, b)
        print(f"parity ok: N={n} B={b} workers={w} mode={mode}", flush=True)

if what in ("speed", "all"):
    for b in roots:
        n = 19
        p, m = capi.generate_roots(0, 0, b, n)
        for pair in ([int(x) for x in os.environ["PROBE_PAIR"].split(",")] if os.environ.get("PROBE_PAIR") else (0, 1)):
            os.environ["AZB_ASYNC_PAIR"] = str(pair)
            for w in workers:
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
                    print(f"B={b} pair={pair} workers={w:3d}: {ms/steps*1e3:7.1f} us/step {k['n_live']/(ms*1e-3)/1e6:7.2f} M sims/s argmin {h.argmin()['eval']:.5f}", flush=True)
                    if os.environ.get("PROBE_PROF"):
                        L = capi.lib()
                        L.azb_debug_async.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
                        d = (C.c_uint64 * 32)()
                        L.azb_debug_async(h._h, d)
                        tiles = max(d[4], 1)
                        us = lambda c: c / tiles / 1965.0
                        print(f"      per tile (us): acquire {us(d[0]):6.1f} wait-empty {us(d[1]):6.1f} wait-layer {us(d[2]):6.1f} tile {us(d[3]):6.1f} | "
                              f"MMA wait-full {us(d[6]):6.1f} wait-acc {us(d[7]):6.1f} | epi wait-acc {us(d[11]):6.1f} busy {us(d[13]):6.1f} (tmem-ld {us(d[10]):6.1f} fence {us(d[12]):6.1f}) | "
                              f"tiles {d[4]} rows real {d[15] >> 32} dummy {d[15] & 0xffffffff}", flush=True)
