"""Parity at the configurations bench.py times ("parity at the timed configurations"), and the stress
cases the reference's tie rules and unbounded cascade frontier call for.

  * C2: 4096 roots, N = 19, asynchronous kernel with 20 model SMs, one worker per tile — every tree equal to the lock
    step, and 64 of them equal to the oracle fed with the priors the device produced;
  * the C3 per-GPU shape (8192 roots, 48 model SMs in pairs) and the C4 shape (N = 64, 32 model SMs);
  * whole trees against the oracle with the DENSE symmetric eigensolve (what the reference's faer call is:
    rooted_tree/ordered_edge.rs:72-82) instead of the kernel's own sectioning / matching-polynomial method;
  * tie stress: priors quantised to {0.25, 0.5, 0.75} and tiny revisit budgets force constant ties in (n_t, c*) and in the
    curiosity maximum (nabla/tree/next_action.rs:28-88) — pure-Python restatement, oracle and GPU must agree;
  * a Boolean-lattice DAG whose cascades visit hundreds of ancestors per wave (nabla/tree/empty_transitions.rs:50-127).
"""
import numpy as np
import pytest

import pyref
from test_oracle_tree import digest

pytestmark = pytest.mark.gpu


def _mk(capi, n, b, **kw):
    return capi.Handle(capi.default_config(n, b, **kw))


def _digests(h, idx):
    return [digest(h.dump_tree(i)) for i in idx]


def _assert_same_runs(ha, hb, b):
    idx = range(b)
    assert _digests(ha, idx) == _digests(hb, idx)
    wa, wb = ha.walkers(), hb.walkers()
    for k in ("parents", "permitted", "path", "pos", "path_len"):
        assert np.array_equal(wa[k], wb[k]), k
    aa, ab = ha.argmin(), hb.argmin()
    assert aa["eval"] == ab["eval"] and np.array_equal(aa["parents"], ab["parents"])
    assert ha.counters() == hb.counters()


def _oracle_follows_device(orc, o, h, idx, steps):
    """Lock step on the device, one step per call; the oracle advances the trees `idx` with the priors the device's
    model produced for them (the MLP is bf16 on the tensor cores: parity is defined on identical priors)."""
    for _ in range(steps):
        h.step(1)
        o.rollout()
        o.add_actions(np.ascontiguousarray(h.priors()[idx]))


def test_c2_async_w20_4096_roots_equals_lock_step_and_oracle(capi, orc):
    n, b, steps, workers = 19, 4096, 200, 20  # bench.py's default launch shape: 128 tree SMs + 20 model SMs, G = 1
    parents, masks = capi.generate_roots(0, 0, b, n)
    kw = dict(prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps + 2)
    idx = list(range(0, b, b // 64))[:64]
    o = orc.Optimizer(n, len(idx), lambda_method=orc.LAMBDA_MULTISECTION, n_threads=4)
    o.set_roots(parents[idx], masks[idx])
    with _mk(capi, n, b, **kw) as lock, _mk(capi, n, b, async_workers=workers, **kw) as asy:
        for h in (lock, asy):
            h.mlp_init(1)
            h.set_roots(parents, masks)
            h.init_trees()
        o.init_trees(np.ascontiguousarray(lock.priors()[idx]))
        n2, log2 = asy.step(steps, cap=1024)  # ONE launch of the persistent kernel, like bench.py's timed call
        _oracle_follows_device(orc, o, lock, idx, steps)
        _assert_same_runs(lock, asy, b)
        for j, i in enumerate(idx):  # the asynchronous kernel's trees against the oracle's
            d_o, d_g = o.dump_tree(j), asy.dump_tree(i)
            for key in ("nodes", "keys", "preds", "arcs"):
                assert np.array_equal(d_o[key], d_g[key]), (i, key)
        k = asy.counters()
        assert k["n_live"] + k["n_noop"] == b * steps
        assert n2 == len(log2) and all(log2[j][3] > log2[j + 1][3] for j in range(len(log2) - 1))  # strictly improving


def test_c3_shape_8192_roots_w48_pairs_equals_lock_step(capi):
    n, b, steps, workers = 19, 8192, 100, 48  # one GPU's share of 65 536 roots over 8 GPUs; G = 2 from 40 workers
    parents, masks = capi.generate_roots(2, 8192, b, n)  # the second rank's roots
    kw = dict(prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps + 2, first_root=8192)
    with _mk(capi, n, b, **kw) as lock, _mk(capi, n, b, async_workers=workers, **kw) as asy:
        for h in (lock, asy):
            h.mlp_init(1)
            h.set_roots(parents, masks)
            h.init_trees()
            h.step(steps)
        _assert_same_runs(lock, asy, b)


def test_c4_shape_n64_w32_equals_lock_step_and_oracle(capi, orc):
    n, b, steps, workers = 64, 1024, 60, 32
    parents, masks = capi.generate_roots(4, 0, b, n)
    kw = dict(prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps + 2)
    idx = list(range(0, b, b // 8))[:8]
    o = orc.Optimizer(n, len(idx), lambda_method=orc.LAMBDA_MULTISECTION, n_threads=4)
    o.set_roots(parents[idx], masks[idx])
    with _mk(capi, n, b, **kw) as lock, _mk(capi, n, b, async_workers=workers, **kw) as asy:
        for h in (lock, asy):
            h.mlp_init(2)
            h.set_roots(parents, masks)
            h.init_trees()
        o.init_trees(np.ascontiguousarray(lock.priors()[idx]))
        asy.step(steps)
        _oracle_follows_device(orc, o, lock, idx, steps)
        _assert_same_runs(lock, asy, b)
        for j, i in enumerate(idx):
            assert digest(o.dump_tree(j)) == digest(asy.dump_tree(i)), i


@pytest.mark.parametrize("n,b,steps", [(19, 48, 250), (64, 6, 70)])
def test_whole_trees_against_the_dense_eigensolve_oracle(capi, orc, n, b, steps):
    """The oracle in LAMBDA_DENSE mode computes lambda_1 the way the reference does (a dense symmetric eigensolve of the
    adjacency matrix); the device uses the matching polynomial (N <= 22) or Sturm sections.  Every f32 the trees hold,
    every selected action, visit count and the argmin must still agree."""
    seed = 21
    a_dim = orc.action_dim(n)
    parents, masks = orc.generate_roots(seed, 0, b, n)
    o = orc.Optimizer(n, b, lambda_method=orc.LAMBDA_DENSE, n_threads=4)
    o.set_roots(parents, masks)
    o.init_trees(orc.hash_priors(seed, 0, b, a_dim, 0))
    imp_o = o.steps_hash(seed, 0, 1, steps)
    with _mk(capi, n, b, prior_mode=capi.PRIOR_HASH, prior_seed=seed, max_steps=steps) as h:
        h.set_roots(parents, masks)
        h.init_trees()
        _, imps = h.step(steps, cap=1024)
        for i in range(b):
            d_o, d_g = o.dump_tree(i), h.dump_tree(i)
            for key in ("nodes", "keys", "preds", "arcs"):
                assert np.array_equal(d_o[key], d_g[key]), (i, key)
        assert [s - 1 for s in imp_o] == [s for (s, _, _, _) in imps]
        assert h.counters() == o.counters()
        ao, ag = o.argmin(), h.argmin()
        assert ag["eval"] == ao["eval"] and ag["mu"] == ao["mu"] and abs(ag["lambda1"] - ao["lambda1"]) <= 1e-12 * ao["lambda1"]


def quantised_priors(rng, b, a_dim):
    return rng.choice(np.array([0.25, 0.5, 0.75], dtype=np.float32), size=(b, a_dim))


@pytest.mark.parametrize("n,tol,tol_default,seed", [(6, (2, 1), 1, 0), (8, (3, 2), 1, 1), (9, (1,), 1, 2), (10, (4, 2, 2), 2, 3),
                                                    (12, (2, 2), 1, 4), (7, (200, 50, 50), 25, 5)])
def test_tie_stress_python_oracle_gpu_agree(capi, orc, n, tol, tol_default, seed):
    b, steps = 6, 70
    a_dim = orc.action_dim(n)
    rng = np.random.default_rng(100 + seed)
    parents, masks = orc.generate_roots(seed, 0, b, n, k_min=min(5, a_dim // 2), k_max=a_dim // 2)
    o = orc.Optimizer(n, b, n_as_tol=tol, n_as_tol_default=tol_default, lambda_method=orc.LAMBDA_DENSE)
    o.set_roots(parents, masks)
    py = pyref.Optimizer(n, parents, [orc.actions_from_mask(m) for m in masks], tol=tol, tol_default=tol_default)
    pri = quantised_priors(rng, b, a_dim)
    o.init_trees(pri)
    py.init_trees(pri)
    with _mk(capi, n, b, prior_mode=capi.PRIOR_INJECTED, n_as_tol=tol, n_as_tol_default=tol_default, max_steps=steps) as h:
        h.set_roots(parents, masks)
        h.set_priors(pri)
        h.init_trees()
        sv = np.zeros((b, 2 * a_dim), dtype=np.float32)
        for s in range(steps):
            pri = quantised_priors(rng, b, a_dim)
            o.rollout()
            h.rollout_host(sv)
            imp_o = o.add_actions(pri)
            imp_g = h.add_actions_host(pri)
            imp_p = py.step(pri)
            assert imp_o == imp_g == (imp_p is not None), f"step {s}"
        for i in range(b):
            d_o, d_g, d_p = o.dump_tree(i), h.dump_tree(i), py.dump(i, orc.mask_words(n))
            for key in ("nodes", "keys", "preds", "arcs"):
                assert np.array_equal(d_o[key], d_g[key]), (i, key, "oracle vs gpu")
                assert np.array_equal(np.asarray(d_p[key]), d_o[key]), (i, key, "python vs oracle")
        assert h.argmin()["eval"] == o.argmin()["eval"] == py.best


def lattice_roots(orc, n, b):
    """Roots whose reachable states form a Boolean lattice: every child 2..N-2 hangs off vertex 0 and has exactly one
    permitted action (re-parent to vertex 1), so a state is the SET of children moved and every order of the same moves
    reaches the same node: the densest transposition DAG this space can build."""
    parents = np.zeros((b, n), dtype=np.uint8)
    acts = [c * (c - 1) // 2 for c in range(2, n - 1)]  # action (parent 1, child c): ordered_edge.rs:35-38
    masks = np.stack([orc.mask_from_actions(n, acts) for _ in range(b)])
    return parents, masks


@pytest.mark.parametrize("frontier_cap", [None, 4])
def test_dense_transposition_dag_and_the_cascade_continuation_lists(capi, orc, monkeypatch, frontier_cap):
    """The cascade's work lists keep AZB_FRONTIER_CAP (128) entries on chip and continue in HBM, so no DAG shape can
    overflow them.  With the on-chip part shrunk to 4 entries (AZB_TEST_FRONTIER_CAP, read by azb_create) almost every
    cascade of the lattice runs through the continuation; both settings must reproduce the oracle bit for bit."""
    n, b, steps, seed = 14, 3, 2300, 31  # 11 movable children: all 2048 lattice nodes get inserted and exhausted
    if frontier_cap:
        monkeypatch.setenv("AZB_TEST_FRONTIER_CAP", str(frontier_cap))
    a_dim = orc.action_dim(n)
    parents, masks = lattice_roots(orc, n, b)
    tol, tol_default = (4, 2, 2), 2
    o = orc.Optimizer(n, b, n_as_tol=tol, n_as_tol_default=tol_default, lambda_method=orc.LAMBDA_MULTISECTION, n_threads=3)
    o.set_roots(parents, masks)
    o.init_trees(orc.hash_priors(seed, 0, b, a_dim, 0))
    o.steps_hash(seed, 0, 1, steps)
    with _mk(capi, n, b, prior_mode=capi.PRIOR_HASH, prior_seed=seed, n_as_tol=tol, n_as_tol_default=tol_default,
             max_steps=steps) as h:
        h.set_roots(parents, masks)
        h.init_trees()
        h.step(steps)
        if frontier_cap:
            assert h.cascade_spills() > 1000, "the shrunken work list did not push cascades into the continuation"
        for i in range(b):
            d_o, d_g = o.dump_tree(i), h.dump_tree(i)
            for key in ("nodes", "keys", "preds", "arcs"):
                assert np.array_equal(d_o[key], d_g[key]), (i, key)
        assert h.counters() == o.counters()
        assert int(h.tree_sizes(0)[0]) == 2 ** (n - 3)  # the whole lattice was inserted and exhausted


def test_ordinary_trees_through_the_cascade_continuation(capi, orc, monkeypatch):
    """The example's root distribution with a 2-entry on-chip list: the walker's own path already spills."""
    monkeypatch.setenv("AZB_TEST_FRONTIER_CAP", "2")
    n, b, steps, seed = 19, 24, 150, 8
    parents, masks = orc.generate_roots(seed, 0, b, n)
    o = orc.Optimizer(n, b, lambda_method=orc.LAMBDA_MULTISECTION, n_threads=4)
    o.set_roots(parents, masks)
    o.init_trees(orc.hash_priors(seed, 0, b, orc.action_dim(n), 0))
    o.steps_hash(seed, 0, 1, steps)
    with _mk(capi, n, b, prior_mode=capi.PRIOR_HASH, prior_seed=seed, max_steps=steps) as h:
        h.set_roots(parents, masks)
        h.init_trees()
        h.step(steps)
        assert h.cascade_spills() > 0
        assert [digest(o.dump_tree(i)) for i in range(b)] == [digest(h.dump_tree(i)) for i in range(b)]
        assert h.counters() == o.counters()


def test_node_state_replays_the_action_set_on_the_root(capi, orc):
    """azb_get_node_state = the root with the node's key replayed (optimizer/mod.rs:224-239); checked against the
    oracle's act() for nodes of several trees, and against azb_get_argmin for the winning node."""
    n, b, steps, seed = 19, 16, 60, 5
    parents, masks = orc.generate_roots(seed, 0, b, n)
    with _mk(capi, n, b, prior_mode=capi.PRIOR_HASH, prior_seed=seed, max_steps=steps) as h:
        h.set_roots(parents, masks)
        h.init_trees()
        n_imp, imps = h.step(steps, cap=256)
        for t in (0, 7, 15):
            d = h.dump_tree(t)
            for node in range(0, len(d["nodes"]), 9):
                p, m = parents[t].copy(), masks[t].copy()
                for a in orc.actions_from_mask(d["keys"][node]):
                    p, m = orc.act(p, m, a)
                gp, gm = h.node_state(t, node)
                assert np.array_equal(gp, p) and np.array_equal(gm, m), (t, node)
        assert n_imp > 0
        _, tree, node, ev = imps[-1]
        am = h.argmin()
        gp, gm = h.node_state(tree, node)
        assert np.array_equal(gp, am["parents"]) and np.array_equal(gm, am["permitted"]) and am["eval"] == np.float32(ev)


def test_update_model_without_observations_is_refused(capi):
    """No root child is exhausted or has n_t >= n_obs_tol right after init_trees: the weight sum is 0 and the reference's
    loss is 0/0 (nabla/model/dfdx.rs:105-110).  The step must be refused with a status, and the model left untouched."""
    n, b = 19, 32
    parents, masks = capi.generate_roots(1, 0, b, n)
    with _mk(capi, n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=8) as h:
        h.mlp_init(3)
        h.set_roots(parents, masks)
        h.init_trees()
        before = h.mlp_get_params().copy()
        with pytest.raises(capi.AzbError) as e:
            h.update_model(200)
        assert e.value.code == capi.ERR_STATE and "no observations" in str(e.value)
        assert np.array_equal(before, h.mlp_get_params())
        h.step(4)  # still usable, priors still finite
        assert np.isfinite(h.priors()).all()
