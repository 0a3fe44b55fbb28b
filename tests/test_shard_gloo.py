"""The N>1 host path on CPU: world_size-2 gloo.  Roots shard by contiguous blocks, each rank's synthetic roots and
hash priors depend only on GLOBAL root indices (so a sharded run is the same problem as the unsharded one), and the
epoch-boundary argmin gather picks the first minimum.  The per-rank search itself runs on the oracle here (CPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, steps, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from azdopt_b200 import capi, shard
    from oracle import oracle

    n = 19
    lo, hi = shard.shard_range(total, rank, world)
    b = hi - lo
    parents, masks = capi.generate_roots(0, lo, b, n)  # the product's generator, global indices
    o = oracle.Optimizer(n, b, lambda_method=oracle.LAMBDA_MULTISECTION)
    o.set_roots(parents, masks)
    o.init_trees(oracle.hash_priors(0, lo, b, oracle.action_dim(n), 0))
    o.steps_hash(0, lo, 1, steps)
    am = o.argmin()
    ev, who, p, m = shard.global_argmin(float(am["eval"]), am["parents"], am["permitted"], dist)
    k = shard.sum_counters(o.counters(), dist)
    digests = [o.tree_sizes(i) for i in range(b)]
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([ev, who, k["n_live"], k["n_ins"]], dtype=np.float64))
    np.save(os.path.join(out_dir, f"s{rank}.npy"), np.array(digests, dtype=np.int64))
    np.save(os.path.join(out_dir, f"p{rank}.npy"), p)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from azdopt_b200 import shard

    for total in (1, 7, 4096, 65536, 1000):
        for world in (1, 2, 3, 8):
            edges = [shard.shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_matches_single_rank(tmp_path):
    from oracle import oracle

    oracle.build()
    total, steps, world = 24, 60, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, total, steps, str(tmp_path)), nprocs=world, join=True)
    # single-rank run of the same global problem
    n = 19
    parents, masks = oracle.generate_roots(0, 0, total, n)
    o = oracle.Optimizer(n, total, lambda_method=oracle.LAMBDA_MULTISECTION)
    o.set_roots(parents, masks)
    o.init_trees(oracle.hash_priors(0, 0, total, oracle.action_dim(n), 0))
    o.steps_hash(0, 0, 1, steps)
    k = o.counters()
    am = o.argmin()
    r0 = np.load(tmp_path / "r0.npy")
    r1 = np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1)  # both ranks hold the same gathered result
    assert np.float32(r0[0]) == am["eval"]
    assert int(r0[2]) == k["n_live"] and int(r0[3]) == k["n_ins"]
    sizes = np.concatenate([np.load(tmp_path / "s0.npy"), np.load(tmp_path / "s1.npy")])
    assert np.array_equal(sizes, np.array([o.tree_sizes(i) for i in range(total)]))
    assert np.array_equal(np.load(tmp_path / "p0.npy"), am["parents"])
