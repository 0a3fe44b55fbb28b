"""The oracle's search-DAG logic: against the independent pure-Python restatement (tests/pyref.py), against the
invariants of SURVEY.md Appendix A.6, and against the committed golden digests.  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

import pyref

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "oracle_trees.json")


def _same(d1, d2):
    for k in ("nodes", "keys", "preds", "arcs"):
        a, b = np.asarray(d1[k]), np.asarray(d2[k])
        if a.shape != b.shape or not np.array_equal(a, b):
            return k
    return None


@pytest.mark.parametrize("n,b,steps,tol,tol_default", [(7, 6, 60, (200, 50, 50), 25), (9, 5, 50, (3, 2), 1),
                                                       (19, 3, 40, (200, 50, 50), 25), (12, 4, 60, (4, 2, 2), 2)])
def test_oracle_matches_python_restatement(orc, n, b, steps, tol, tol_default):
    a_dim = orc.action_dim(n)
    parents, masks = orc.generate_roots(1, 0, b, n, k_min=min(5, a_dim // 2), k_max=a_dim // 2)
    o = orc.Optimizer(n, b, n_as_tol=tol, n_as_tol_default=tol_default, lambda_method=orc.LAMBDA_DENSE)
    o.set_roots(parents, masks)
    py = pyref.Optimizer(n, parents, [orc.actions_from_mask(m) for m in masks], tol=tol, tol_default=tol_default)
    pri = orc.hash_priors(9, 0, b, a_dim, 0)
    o.init_trees(pri)
    py.init_trees(pri)
    for s in range(steps):
        pri = orc.hash_priors(9, 0, b, a_dim, s + 1)
        o.rollout()
        imp_o = o.add_actions(pri)
        imp_p = py.step(pri)
        assert imp_o == (imp_p is not None), f"step {s}"
    w = o.walkers()
    for i in range(b):
        bad = _same(o.dump_tree(i), py.dump(i, orc.mask_words(n)))
        assert bad is None, f"tree {i}: {bad} differ"
        assert int(w["pos"][i]) == py.pos[i] and int(w["path_len"][i]) == len(py.paths[i])
        assert list(w["parents"][i]) == py.states[i].parents
        assert orc.actions_from_mask(w["permitted"][i]) == sorted(py.states[i].permitted)
    assert o.argmin()["eval"] == py.best


def _check_invariants(orc, o, n, b, roots, masks):
    for i in range(b):
        d = o.dump_tree(i)
        nodes, keys, preds, arcs = d["nodes"], d["keys"], d["preds"], d["arcs"]
        c = nodes[:, 0].view(np.float32)
        cs = nodes[:, 1].view(np.float32)
        assert len(arcs) >= len(nodes) - 1
        assert (cs <= c).all()
        active = nodes[:, 4] + nodes[:, 3] < nodes[:, 5]
        ex = np.zeros(len(nodes), dtype=np.int64)
        for src, dst, ppos in arcs:
            ex[src] += 0 if active[dst] else 1
            ka = set(orc.actions_from_mask(keys[src]))
            kb = set(orc.actions_from_mask(keys[dst]))
            a = int(preds[ppos, 0])
            assert kb == ka | {a} and a not in ka
            assert nodes[src, 4] <= ppos < nodes[src, 5]
        assert np.array_equal(ex, nodes[:, 3].astype(np.int64))
        depth = np.array([len(orc.actions_from_mask(k)) for k in keys])
        assert depth[0] == 0 and depth.max() <= min(n - 3, len(orc.actions_from_mask(masks[i])))
        # replaying a key from the root is legal in ascending order and reproduces the node's cost
        for k in range(0, len(nodes), max(1, len(nodes) // 8)):
            p, m = roots[i].copy(), masks[i].copy()
            for a in orc.actions_from_mask(keys[k]):
                assert a in orc.action_data(p, m)
                p, m = orc.act(p, m, a)
            assert orc.cost(p)[2].view(np.uint32) == nodes[k, 0]


@pytest.mark.parametrize("n,b,steps", [(19, 12, 300), (10, 8, 200), (64, 2, 60)])
def test_oracle_invariants(orc, n, b, steps):
    parents, masks = orc.generate_roots(0, 0, b, n)
    o = orc.Optimizer(n, b, n_threads=4)
    o.set_roots(parents, masks)
    o.init_trees(orc.hash_priors(0, 0, b, orc.action_dim(n), 0))
    o.steps_hash(0, 0, 1, steps)
    _check_invariants(orc, o, n, b, parents, masks)
    k = o.counters()
    assert k["n_live"] + k["n_noop"] == b * steps
    assert k["n_ins"] >= k["n_live"] and k["n_arc"] == k["n_ins"] + k["n_hit"]
    assert k["n_reset"] == k["n_term"] + k["n_hit"]


def test_oracle_thread_count_does_not_change_results(orc):
    n, b = 19, 16
    parents, masks = orc.generate_roots(3, 0, b, n)
    dumps = []
    for threads in (1, 5):
        o = orc.Optimizer(n, b, n_threads=threads)
        o.set_roots(parents, masks)
        o.init_trees(orc.hash_priors(3, 0, b, orc.action_dim(n), 0))
        imp = o.steps_hash(3, 0, 1, 120)
        dumps.append((imp, [o.dump_tree(i) for i in range(b)], o.counters()))
    assert dumps[0][0] == dumps[1][0] and dumps[0][2] == dumps[1][2]
    for d0, d1 in zip(dumps[0][1], dumps[1][1]):
        assert _same(d0, d1) is None


def digest(dump):
    h = hashlib.sha256()
    for k in ("nodes", "keys", "preds", "arcs"):
        h.update(np.ascontiguousarray(dump[k], dtype=np.uint32).tobytes())
    return h.hexdigest()


def run_golden_case(orc, case):
    n, b = case["n"], case["b"]
    parents, masks = orc.generate_roots(case["seed"], case["first_root"], b, n)
    o = orc.Optimizer(n, b, lambda_method=orc.LAMBDA_MULTISECTION)
    o.set_roots(parents, masks)
    o.init_trees(orc.hash_priors(case["seed"], case["first_root"], b, orc.action_dim(n), 0))
    imp = o.steps_hash(case["seed"], case["first_root"], 1, case["steps"])
    am = o.argmin()
    return dict(improved_steps=imp, counters=o.counters(), digests=[digest(o.dump_tree(i)) for i in range(b)],
                argmin_eval_bits=int(np.float32(am["eval"]).view(np.uint32)), argmin_mu=int(am["mu"]),
                argmin_parents=[int(x) for x in am["parents"]])


def test_oracle_against_golden_digests(orc):
    # tests/golden/make_golden.py wrote these from the oracle after it passed every test above; they pin the oracle
    # (and, in test_gpu_parity.py, the CUDA path) against silent drift
    with open(GOLDEN) as f:
        golden = json.load(f)
    for case in golden["cases"]:
        got = run_golden_case(orc, case["config"])
        assert got == case["expect"], case["config"]
