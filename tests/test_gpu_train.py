"""GPU parity of the epoch-boundary training step (azb_model_gradients / azb_model_update / azb_update_model)
against the numpy oracle (oracle.model_gradients / adam_step, pinned to torch autograd + torch Adam in
tests/test_oracle_train.py).  Floating point: tolerance 1e-5 relative (north_star), written at each assert."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mk(capi, n, b, **kw):
    return capi.Handle(capi.default_config(n, b, **kw))


def _dims(a):
    return [2 * a, 512, 1024, 512, a]


def _rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / max(np.linalg.norm(b.astype(np.float64)), 1e-30))


@pytest.mark.parametrize("n,rows", [(19, 96), (8, 33)])
def test_gradients_match_oracle(capi, orc, n, rows):
    a = orc.action_dim(n)
    rng = np.random.default_rng(n)
    x = (rng.random((rows, 2 * a)) < 0.3).astype(np.float32)
    o = rng.random((rows, a)).astype(np.float32)
    w = (rng.random((rows, a)) < 0.2).astype(np.float32)
    with _mk(capi, n, rows, prior_mode=capi.PRIOR_HASH) as h:
        h.mlp_init(3)
        params = h.mlp_get_params()
        loss, g = h.model_gradients(x, o, w)
        want_loss, want_g = orc.model_gradients(params, _dims(a), x, o, w)
        assert abs(loss - want_loss) <= 1e-5 * abs(want_loss)      # 1e-5 relative
        assert _rel(g, want_g) <= 1e-5                             # 1e-5 relative, norm-wise
        assert np.allclose(g, want_g, rtol=1e-3, atol=1e-6 * np.abs(want_g).max())
        # the hook leaves the model untouched and the gradient buffer zeroed: same answer twice
        loss2, g2 = h.model_gradients(x, o, w)
        assert loss2 == loss and np.array_equal(g, g2)
        assert np.array_equal(h.mlp_get_params(), params)


def test_model_update_matches_oracle_adam(capi, orc):
    n, rows = 19, 64
    a = orc.action_dim(n)
    rng = np.random.default_rng(5)
    with _mk(capi, n, rows, prior_mode=capi.PRIOR_HASH) as h:
        h.mlp_init(11)
        cur = h.mlp_get_params()
        st = orc.AdamState(cur.size)
        for it in range(4):
            x = (rng.random((rows, 2 * a)) < 0.3).astype(np.float32)
            o = rng.random((rows, a)).astype(np.float32)
            w = (rng.random((rows, a)) < 0.2).astype(np.float32)
            loss = h.model_update(x, o, w)
            want_loss, cur = orc.update_model(cur, _dims(a), x, o, w, st)
            assert abs(loss - want_loss) <= 1e-5 * abs(want_loss)
            got = h.mlp_get_params()
            # one Adam step moves a parameter by at most ~lr = 1e-4; agree to 1% of that and 1e-5 relative overall
            assert np.abs(got - cur).max() <= 1e-6
            assert _rel(got, cur) <= 1e-5
        assert np.abs(cur - h.mlp_get_params()).max() <= 1e-6


@pytest.mark.parametrize("mode", ["fp32", "tc"])
def test_update_model_on_device_trees(capi, orc, mode):
    """par_update_model (optimizer/mod.rs:249-281) fused on the device == write_observations + the oracle's update."""
    n, b, steps, tol = 19, 48, 60, 3
    a = orc.action_dim(n)
    parents, masks = orc.generate_roots(6, 0, b, n)
    kw = dict(prior_mode=capi.PRIOR_HASH, prior_seed=9, max_steps=steps)
    if mode == "tc":
        kw["mlp_mode"] = capi.MLP_TC
    with _mk(capi, n, b, **kw) as h:
        h.mlp_init(2)
        h.set_roots(parents, masks)
        h.init_trees()
        h.step(steps)
        sv, obs, w = h.write_observations(tol)
        assert w.sum() > 0
        params = h.mlp_get_params()
        st = orc.AdamState(params.size)
        want_loss, want = orc.update_model(params, _dims(a), sv, obs, w, st)
        loss = h.update_model(tol)
        assert abs(loss - want_loss) <= 1e-5 * abs(want_loss)
        got = h.mlp_get_params()
        assert np.abs(got - want).max() <= 1e-6 and _rel(got, want) <= 1e-5
        assert np.abs(got - params).max() > 5e-5  # it did train
        # the rollout model sees the new parameters (the tensor-core copy is reloaded)
        x = (np.random.default_rng(0).random((b, 2 * a)) < 0.2).astype(np.float32)
        y = h.model_write_predictions(x)
        ref = orc.mlp_forward(got, _dims(a), x, n_threads=4)
        assert np.abs(y - ref).max() < (2e-2 if mode == "tc" else 1e-5)
        # a second epoch boundary keeps the Adam state (t = 2)
        want_loss2, want2 = orc.update_model(want, _dims(a), sv, obs, w, st)
        loss2 = h.update_model(tol)
        assert abs(loss2 - want_loss2) <= 2e-5 * abs(want_loss2)
        assert np.abs(h.mlp_get_params() - want2).max() <= 2e-6


def _cmp_trees(orc_opt, h, b):
    for i in range(b):
        d_o, d_g = orc_opt.dump_tree(i), h.dump_tree(i)
        for k in ("nodes", "keys", "preds", "arcs"):
            assert d_o[k].shape == d_g[k].shape and np.array_equal(d_o[k], d_g[k]), f"tree {i} {k}"


@pytest.mark.parametrize("n,k_max", [(19, 9), (12, 0), (33, 0)])
def test_reset_trees_policy_matches_oracle(capi, orc, n, k_max):
    """azb_reset_trees (device modify_root + init) against the oracle with the same counter draws: the re-selected
    roots, the rebuilt trees and the next epoch's search are bit-exact, over several epochs."""
    b, steps = 72, 30
    a = orc.action_dim(n)
    kmx = k_max or a // 2
    parents, masks = orc.generate_roots(3, 0, b, n, k_min=5, k_max=kmx)
    o = orc.Optimizer(n, b, lambda_method=orc.LAMBDA_MULTISECTION)
    o.set_roots(parents, masks)
    o.init_trees(orc.hash_priors(5, 0, b, a, 0))
    with _mk(capi, n, b, prior_mode=capi.PRIOR_HASH, prior_seed=5, max_steps=steps) as h:
        h.set_roots(parents, masks)
        h.init_trees()
        moved = 0
        for epoch in range(3):
            h.step(steps)
            o.steps_hash(5, 0, 1, steps)
            _cmp_trees(o, h, b)
            h.reset_trees(21, 5, kmx)
            o.modify_roots(21, epoch, 0, 5, kmx)
            pg, mg = h.get_roots()
            po, mo = o.get_roots()
            assert np.array_equal(pg, po) and np.array_equal(mg, mo)
            moved += int((pg != parents).any(axis=1).sum())
            parents = pg
            o.reinit_trees(orc.hash_priors(5, 0, b, a, 0))
            _cmp_trees(o, h, b)
            ag, ao = h.argmin(), o.argmin()
            assert ag["eval"] == ao["eval"] and np.array_equal(ag["parents"], ao["parents"])
        h.step(steps)
        o.steps_hash(5, 0, 1, steps)
        _cmp_trees(o, h, b)
        assert h.argmin()["eval"] == o.argmin()["eval"]
        assert moved > b  # roots really move


def test_python_mirror_runs_the_examples_epoch_loop(capi, orc):
    """NablaOptimizer (azdopt_b200/nabla.py) with the reference's method names: par_new, the fused steps,
    par_update_model, par_reset_trees, argmin_data — two epochs equal the same calls made on the raw handle."""
    from azdopt_b200 import nabla, observe

    n, b, steps = 19, 64, 40
    roots = capi.generate_roots(4, 0, b, n)
    opt = nabla.NablaOptimizer.par_new(nabla.ROTModifyParentsOnce(n=n), roots, nabla.ActionModel(seed=5, arithmetic="tc"),
                                       b, max_steps=steps)
    with _mk(capi, n, b, prior_mode=capi.PRIOR_MLP, mlp_mode=capi.MLP_TC, max_steps=steps) as h:
        h.mlp_init(5)
        h.set_roots(*roots)
        h.init_trees()
        for epoch in range(2):
            imp = opt.roll_out(steps)
            n_imp, log = h.step(steps, cap=steps)
            assert [tuple(x) for x in imp] == [tuple(x) for x in log]
            assert opt.par_update_model(3) == h.update_model(3)
            opt.par_reset_trees(seed=9)
            h.reset_trees(9)
            a, g = opt.argmin_data(), h.argmin()
            assert a.eval == g["eval"] and a.mu == g["mu"] and np.array_equal(a.parents, g["parents"])
        assert opt.argmin_graph6() == observe.graph6_of_state(g["parents"])
        assert opt.tree_dot(0).startswith("graph search_tree {")
    opt.close()
