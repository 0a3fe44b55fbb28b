"""Epoch boundary over 2 GPUs (needs >= 2 CUDA devices; skipped on a 1-GPU box): a 2-rank sharded run — roots in
contiguous blocks, no collective inside a step, then the NCCL all-reduce of weight sum / gradient / loss inside
azb_update_model and the NCCL all-gather inside azb_comm_argmin — must equal ONE handle holding all the roots."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N, TOTAL, STEPS, TOL = 19, 64, 40, 3


def _run(capi, h, lo, b):
    parents, masks = capi.generate_roots(3, lo, b, N)
    h.mlp_init(4)
    h.set_roots(parents, masks)
    h.init_trees()
    h.step(STEPS)


def _worker(rank, world, id_path, out_dir):
    sys.path.insert(0, ROOT)
    from azdopt_b200 import capi, shard

    lo, hi = shard.shard_range(TOTAL, rank, world)
    cfg = capi.default_config(N, hi - lo, prior_mode=capi.PRIOR_HASH, prior_seed=9, max_steps=STEPS, device=rank,
                              first_root=lo)
    with capi.Handle(cfg) as h:
        _run(capi, h, lo, hi - lo)
        if rank == 0:
            uid = capi.comm_unique_id()
            with open(id_path + ".tmp", "wb") as f:
                f.write(uid)
            os.replace(id_path + ".tmp", id_path)
        else:
            import time
            for _ in range(600):
                if os.path.exists(id_path):
                    break
                time.sleep(0.05)
            uid = open(id_path, "rb").read()
        h.comm_init(uid, rank, world)
        loss = h.update_model(TOL)
        p, m, lam, mu, ev, owner = h.comm_argmin()
        assert h.comm_allreduce_bench(3) > 0.0  # the gradient-sized all-reduce alone, timed with events
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), loss=loss, params=h.mlp_get_params(), parents=p, permitted=m,
                 lam=lam, mu=mu, ev=ev, owner=owner, local_ev=h.argmin()["eval"])


def test_two_gpu_update_model_and_argmin_match_one_handle(capi, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    mp.spawn(_worker, args=(world, str(tmp_path / "nccl_id"), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{k}.npz") for k in range(world)]
    # every rank applied the same global step
    assert np.array_equal(r[0]["params"], r[1]["params"]) and r[0]["loss"] == r[1]["loss"]
    with capi.Handle(capi.default_config(N, TOTAL, prior_mode=capi.PRIOR_HASH, prior_seed=9, max_steps=STEPS)) as h:
        _run(capi, h, 0, TOTAL)
        loss = h.update_model(TOL)
        one = h.mlp_get_params()
        am = h.argmin()
    assert abs(float(r[0]["loss"]) - loss) <= 1e-5 * abs(loss)           # f32 sums in a different order: 1e-5 relative
    assert np.abs(r[0]["params"] - one).max() <= 1e-6
    # global argmin: the same state and value as the unsharded run; owner = the rank holding the better local best
    for k in range(world):
        assert np.array_equal(r[k]["parents"], am["parents"]) and np.array_equal(r[k]["permitted"], am["permitted"])
        assert np.float32(r[k]["ev"]) == am["eval"] and int(r[k]["mu"]) == am["mu"]
    evs = [float(r[k]["local_ev"]) for k in range(world)]
    assert int(r[0]["owner"]) == int(np.argmin(evs)) == int(r[1]["owner"])
