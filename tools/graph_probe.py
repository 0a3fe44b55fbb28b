"""Throughput of azb_eval_graph_costs beside the oracle's restatement on the host.
usage: python tools/graph_probe.py [m]      prints one JSON line per (n, density)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from azdopt_b200 import build, capi  # noqa: E402
from graphs_util import random_connected_graph  # noqa: E402
from oracle import oracle as orc  # noqa: E402  (the checker / CPU baseline only)

build()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
h = capi.Handle(capi.default_config(6, 2, prior_mode=capi.PRIOR_HASH))
for n, p in ((19, 0.0), (19, 0.15), (32, 0.0), (32, 0.06), (32, 0.5)):
    rng = np.random.default_rng(n)
    base = np.stack([random_connected_graph(rng, n, p) for _ in range(2048)])
    graphs = np.ascontiguousarray(np.tile(base, (m // 2048, 1)))
    h.eval_graph_costs(graphs[:4096])
    best = min(h.eval_graph_costs(graphs)[3] for _ in range(3))
    call_ms = 1e30
    for _ in range(3):  # the whole call through the C ABI with host buffers: H2D, kernel, D2H of all three outputs
        t0 = time.perf_counter()
        l1, mu, kinds, _ = h.eval_graph_costs(graphs)
        call_ms = min(call_ms, (time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter()
    k = 256
    for i in range(k):
        lo, mo, _ = orc.graph_cost(base[i])
        ko = orc.graph_action_kinds(base[i])
        assert mo == mu[i] and abs(lo - l1[i]) <= 1e-12 * lo and np.array_equal(ko, kinds[i])
    cpu = k / (time.perf_counter() - t0)
    print(json.dumps({"n": n, "p": p, "graphs": len(graphs), "kernel_ms": best, "graphs_per_s": len(graphs) / best * 1e3,
                      "call_ms_host_buffers": call_ms, "graphs_per_s_through_the_call": len(graphs) / call_ms * 1e3,
                      "algorithmic_bytes_per_graph": 4 * n + 12 + 4 * ((n * (n - 1) + 31) // 32),
                      "oracle_graphs_per_s_1_thread_incl_ctypes": cpu}), flush=True)
