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


def _run_both(orc, n, parents, masks, tol, tol_default, steps, prior_fn):
    """The oracle and the pure-Python restatement on the same roots and priors; every tree must come out identical."""
    b = len(parents)
    o = orc.Optimizer(n, b, n_as_tol=tol, n_as_tol_default=tol_default, lambda_method=orc.LAMBDA_DENSE)
    o.set_roots(parents, masks)
    py = pyref.Optimizer(n, parents, [orc.actions_from_mask(m) for m in masks], tol=tol, tol_default=tol_default)
    pri = prior_fn(0)
    o.init_trees(pri)
    py.init_trees(pri)
    for s in range(steps):
        pri = prior_fn(s + 1)
        o.rollout()
        assert o.add_actions(pri) == (py.step(pri) is not None), f"step {s}"
    for i in range(b):
        bad = _same(o.dump_tree(i), py.dump(i, orc.mask_words(n)))
        assert bad is None, f"tree {i}: {bad} differ"
    assert o.argmin()["eval"] == py.best
    return o


@pytest.mark.parametrize("n,tol,tol_default,seed", [(6, (2, 1), 1, 0), (8, (3, 2), 1, 1), (9, (1,), 1, 2),
                                                    (10, (4, 2, 2), 2, 3), (12, (2, 2), 1, 4), (7, (200, 50, 50), 25, 5)])
def test_tie_stress_oracle_matches_python_restatement(orc, n, tol, tol_default, seed):
    """Priors quantised to three values and tiny revisit budgets: ties in (n_t, c*) (first minimum, newest arc first),
    in the curiosity sums (last maximum) and in the no-children argmin (first minimum) at almost every selection —
    nabla/tree/next_action.rs:28-88.  Continuous hash priors almost never tie outside c*."""
    b, steps = 6, 70
    a_dim = orc.action_dim(n)
    rng = np.random.default_rng(100 + seed)
    parents, masks = orc.generate_roots(seed, 0, b, n, k_min=min(5, a_dim // 2), k_max=a_dim // 2)
    vals = np.array([0.25, 0.5, 0.75], dtype=np.float32)
    _run_both(orc, n, parents, masks, tol, tol_default, steps, lambda s: rng.choice(vals, size=(b, a_dim)))


def test_boolean_lattice_dag_oracle_matches_python_restatement(orc):
    """Every child has one permitted action, so states are SETS of moved children: the densest transposition DAG of the
    space (2^8 nodes, every node of depth d has d parents).  Cascades (nabla/tree/empty_transitions.rs:50-127) then
    merge many contributions per parent and per level; both restatements must exhaust the lattice identically."""
    n, b = 11, 2
    a_dim = orc.action_dim(n)
    parents = np.zeros((b, n), dtype=np.uint8)
    acts = [c * (c - 1) // 2 for c in range(2, n - 1)]
    masks = np.stack([orc.mask_from_actions(n, acts) for _ in range(b)])
    o = _run_both(orc, n, parents, masks, (4, 2, 2), 2, 300, lambda s: orc.hash_priors(31, 0, b, a_dim, s))
    assert o.tree_sizes(0)[0] == 2 ** (n - 3) and o.tree_sizes(0)[1] == 2 ** (n - 3) * (n - 3) // 2
