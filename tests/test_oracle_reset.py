"""par_reset_trees + the example's modify_root policy (optimizer/mod.rs:284-360; 04-c21-tree.rs:172-206) in the
oracle: invariants of the policy (the reference has no test for it and draws from an unseeded RNG, so there are no
golden trajectories: parity unpinned by the reference)."""
import numpy as np


def _popc(m):
    return np.array([sum(bin(int(w)).count("1") for w in row) for row in m])


def test_modify_roots_policy_invariants(orc):
    n, b, steps = 12, 64, 30
    a = orc.action_dim(n)
    k_min, k_max = 5, 9
    parents, masks = orc.generate_roots(3, 0, b, n, k_min=k_min, k_max=k_max)
    o = orc.Optimizer(n, b, lambda_method=orc.LAMBDA_MULTISECTION)
    o.set_roots(parents, masks)
    o.init_trees(orc.hash_priors(1, 0, b, a, 0))
    seen = set()
    for epoch in range(4):
        o.steps_hash(1, 0, 1, steps)
        dumps = [o.dump_tree(i) for i in range(b)]
        p0, m0 = o.get_roots()
        best_before = o.argmin()["eval"]
        o.modify_roots(11, epoch, 0, k_min, k_max)
        p1, m1 = o.get_roots()
        k0, k1 = _popc(m0), _popc(m1)
        assert np.all((k1 >= k_min) & (k1 <= k_max))
        assert np.all(p1[:, 0] == 0) and np.all(p1[:, 1] == 0) and np.all(p1[:, n - 1] == 0)
        for v in range(2, n - 1):
            assert np.all(p1[:, v] < v)
        o.reinit_trees(orc.hash_priors(1, 0, b, a, 0))
        assert o.argmin()["eval"] == best_before  # par_reset_trees leaves argmin_data alone
        for i in range(b):
            nodes = dumps[i]["nodes"]
            c = nodes[:, 0].copy().view(np.float32)
            c_root, c_star = c[0], nodes[0, 1:2].copy().view(np.float32)[0]
            new_c = o.dump_tree(i)["nodes"][0, 0:1].copy().view(np.float32)[0]
            if c_root == c_star:
                if k0[i] == k_max:
                    seen.add("regenerate")
                else:
                    seen.add("widen")
                    assert new_c == c_root and k1[i] >= k0[i]   # moved to an equal-cost node, never fewer permissions
            else:
                seen.add("descend")
                thr = (np.float32(c_root) + np.float32(3.0) * np.float32(c_star)) / np.float32(4.0)
                assert new_c <= thr and new_c in c                # the new root is one of the tree's nodes below the bar
    assert seen == {"regenerate", "widen", "descend"}


def test_modify_roots_is_shard_invariant(orc):
    """Draws are keyed by the GLOBAL root index: two half-batches re-select exactly like the whole batch."""
    n, b, steps = 10, 24, 20
    a = orc.action_dim(n)
    parents, masks = orc.generate_roots(2, 0, b, n)

    def run(lo, hi):
        o = orc.Optimizer(n, hi - lo, lambda_method=orc.LAMBDA_MULTISECTION)
        o.set_roots(parents[lo:hi], masks[lo:hi])
        o.init_trees(orc.hash_priors(4, lo, hi - lo, a, 0))
        o.steps_hash(4, lo, 1, steps)
        o.modify_roots(9, 0, lo)
        return o.get_roots()

    pw, mw = run(0, b)
    pa, ma = run(0, b // 2)
    pb, mb = run(b // 2, b)
    assert np.array_equal(pw, np.concatenate([pa, pb])) and np.array_equal(mw, np.concatenate([ma, mb]))
