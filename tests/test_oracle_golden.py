"""The oracle against every known answer the reference's own tests hold for this path, plus
cross-checks against numpy / networkx.  CPU only."""
import math

import numpy as np
import pytest

import pyref


def test_colex_bijection_ten_vertices(orc):
    # graph-state/src/simple_graph/edge.rs:76-83
    i = 0
    for v in range(10):
        for u in range(v):
            assert orc.colex_position(v, u) == i
            assert orc.from_colex_position(i) == (v, u)
            assert pyref.colex(v, u) == i and pyref.from_colex(i) == (v, u)
            i += 1
    assert i == 45


def test_action_index_table(orc):
    # graph-state/src/rooted_tree/ordered_edge.rs:139-161: index 0..8 <-> (parent, child)
    expected = [(0, 2), (1, 2), (0, 3), (1, 3), (2, 3), (0, 4), (1, 4), (2, 4), (3, 4)]
    for idx, (parent, child) in enumerate(expected):
        assert orc.colex_position(parent, child) - 1 == idx
        mx, mn = orc.from_colex_position(idx + 1)
        assert (mn, mx) == (parent, child)
        assert pyref.action_edge(idx) == (parent, child)


def test_action_dim_and_c_upper(orc):
    # rooted_tree/space.rs:46-48; 04-c21-tree.rs:33-35,59-68
    assert orc.action_dim(19) == 152 and orc.action_dim(64) == 1952
    assert orc.c_upper(19) == 15 and orc.c_upper(64) == 40
    for n in range(5, 65):
        assert orc.c_upper(n) == pyref.c_upper(n)


def test_star5_legal_reparentings(orc):
    # ordered_edge.rs:164-172: the star's possible modifications ignoring the last vertex are (2,1)... i.e. children
    # 2 and 3 may move to any other smaller vertex: (parent, child) in {(1,2), (1,3), (2,3)}
    parents = np.array([0, 0, 0, 0, 0], dtype=np.uint8)
    n = 5
    full = orc.mask_from_actions(n, range(orc.action_dim(n)))
    legal = orc.action_data(parents, full)
    got = sorted(pyref.action_edge(a) for a in legal)
    assert got == [(1, 2), (1, 3), (2, 3)]


def test_generator_support_five_vertices(orc):
    # ordered_edge.rs:175-195: exactly these six parent arrays
    expected = {(0, 0, 0, 0, 0), (0, 0, 0, 1, 0), (0, 0, 0, 2, 0), (0, 0, 1, 0, 0), (0, 0, 1, 1, 0), (0, 0, 1, 2, 0)}
    p, _ = orc.generate_roots(7, 0, 2000, 5, k_min=1, k_max=2)
    assert {tuple(int(x) for x in row) for row in p} == expected


@pytest.mark.parametrize("method", [0, 1, 2])
def test_star5_cost(orc, method):
    # ordered_edge.rs:198-213: lambda_1 = 2 (1e-6), one matching edge
    lam, mu, c, rc = orc.cost([0, 0, 0, 0, 0], method=method)
    assert rc == 0 and abs(lam - 2.0) < 1e-6 and mu == 1
    assert abs(lam - 2.0) < 1e-12


@pytest.mark.parametrize("method", [0, 1, 2])
def test_path5_cost(orc, method):
    # ordered_edge.rs:216-234: lambda_1 = 2 cos(pi/6) (1e-6), two matching edges
    lam, mu, c, rc = orc.cost([0, 0, 1, 2, 3], method=method)
    assert rc == 0 and abs(lam - 2.0 * math.cos(math.pi / 6.0)) < 1e-6 and mu == 2
    assert abs(lam - 2.0 * math.cos(math.pi / 6.0)) < 1e-12


def _bfs_parents(n, edges):
    adj = {v: [] for v in range(n)}
    for u, v in edges:
        adj[u].append(v)
        adj[v].append(u)
    order, parent_of = [0], {0: None}
    for v in order:
        for w in sorted(adj[v]):
            if w not in parent_of:
                parent_of[w] = v
                order.append(w)
    new = {v: i for i, v in enumerate(order)}
    parents = [0] * n
    for v in order[1:]:
        parents[new[v]] = new[parent_of[v]]
    return parents


def test_twenty_vertex_tree_matching_nine(orc):
    # simple_graph/connected_bitset_graph/mod.rs:395-422 (an extra known answer for any matching routine)
    edges = [(0, 11), (0, 16), (0, 19), (1, 15), (1, 17), (2, 13), (3, 14), (4, 13), (4, 14), (5, 9), (5, 10), (5, 18),
             (6, 15), (7, 17), (7, 19), (8, 10), (9, 12), (10, 13), (16, 18)]
    parents = _bfs_parents(20, edges)
    assert all(parents[v] < v for v in range(1, 20))
    _, mu, _, _ = orc.cost(parents)
    assert mu == 9
    assert orc.matching_greedy(parents) == 9
    assert pyref.maximum_matching(20, parents) == 9


def _random_parents(rng, n):
    p = np.zeros(n, dtype=np.uint8)
    for i in range(2, n - 1):
        p[i] = rng.integers(0, i)
    return p


@pytest.mark.parametrize("n", [5, 8, 19, 22, 33, 64])
def test_lambda1_methods_agree_with_lapack(orc, n):
    rng = np.random.default_rng(n)
    flips = 0
    methods = [orc.LAMBDA_DENSE, orc.LAMBDA_JACOBI, orc.LAMBDA_MULTISECTION, orc.LAMBDA_SECTION_ONLY]
    if n <= 22:
        methods.append(orc.LAMBDA_POLY)
    trees = [_random_parents(rng, n) for _ in range(150)]
    trees.append(np.maximum(np.arange(n) - 1, 0).astype(np.uint8))  # path: clustered eigenvalues, worst conditioning
    trees.append(np.zeros(n, dtype=np.uint8))  # star
    broom = np.zeros(n, dtype=np.uint8)
    broom[3::2] = 1
    trees.append(broom)  # two hubs: lambda_1 and lambda_2 close
    for p in trees:
        ref = pyref.lambda1(n, p)
        vals = [orc.cost(p, method=m) for m in methods]
        for lam, mu, c, rc in vals:
            assert rc == 0
            assert abs(lam - ref) <= 1e-12 * ref
        assert vals[0][1] == vals[2][1] == pyref.maximum_matching(n, list(p)) == orc.matching_greedy(p)
        if n <= 22:
            assert orc.matching_poly(p) == vals[0][1]  # degree of the matching polynomial = matching number
            assert abs(vals[4][0] - ref) <= 4e-15 * ref
        flips += sum(int(vals[0][2].view(np.uint32) != v[2].view(np.uint32)) for v in vals[2:])
        assert vals[0][2] == pyref.evaluate(n, vals[0][1], vals[0][0])
    assert flips == 0  # the f32 cost of the kernels' methods equals the dense one's


def test_matching_against_networkx(orc):
    nx = pytest.importorskip("networkx")
    rng = np.random.default_rng(5)
    for n in (6, 19, 40, 64):
        for _ in range(40):
            p = _random_parents(rng, n)
            g = nx.Graph([(v, int(p[v])) for v in range(1, n)])
            want = len(nx.max_weight_matching(g, maxcardinality=True))
            assert orc.cost(p)[1] == want


def test_act_action_data_write_vec_against_pyref(orc):
    rng = np.random.default_rng(11)
    for n in (6, 9, 19, 64):
        a_dim = orc.action_dim(n)
        for _ in range(20):
            p = _random_parents(rng, n)
            k = int(rng.integers(1, a_dim // 2 + 1))
            acts = rng.choice(a_dim, size=k, replace=False)
            mask = orc.mask_from_actions(n, acts)
            st = pyref.State(p, acts)
            for _ in range(6):
                legal = orc.action_data(p, mask)
                assert legal == pyref.action_data(n, st)
                assert np.array_equal(orc.write_vec(p, mask), pyref.write_vec(n, st))
                if not legal:
                    break
                a = int(rng.choice(legal))
                p, mask = orc.act(p, mask, a)
                pyref.act(n, st, a)
                assert list(p) == st.parents and orc.actions_from_mask(mask) == sorted(st.permitted)


def test_root_generator_distribution(orc):
    # 04-c21-tree.rs:85,108-112; rooted_tree/mod.rs:14-20; modify_parent_once.rs:14-25
    n = 19
    p, m = orc.generate_roots(0, 0, 512, n)
    assert (p[:, 0] == 0).all() and (p[:, 1] == 0).all() and (p[:, n - 1] == 0).all()
    for i in range(2, n - 1):
        assert (p[:, i] < i).all()
    ks = np.array([len(orc.actions_from_mask(row)) for row in m])
    assert ks.min() >= 5 and ks.max() <= orc.action_dim(n) // 2
    p2, m2 = orc.generate_roots(0, 100, 412, n)
    assert np.array_equal(p[100:], p2) and np.array_equal(m[100:], m2)  # sharding by first_root


def test_mlp_forward_against_numpy(orc):
    rng = np.random.default_rng(3)
    dims = [304, 512, 1024, 512, 152]
    params = []
    for l in range(4):
        b = 1.0 / math.sqrt(dims[l])
        params.append(rng.uniform(-b, b, size=dims[l] * dims[l + 1] + dims[l + 1]).astype(np.float32))
    x = (rng.random((7, 304)) < 0.3).astype(np.float32)
    y = orc.mlp_forward(np.concatenate(params), dims, x)
    cur = x.astype(np.float64)
    for l in range(4):
        w = params[l][: dims[l] * dims[l + 1]].reshape(dims[l + 1], dims[l]).astype(np.float64)
        b = params[l][dims[l] * dims[l + 1]:].astype(np.float64)
        cur = cur @ w.T + b
        cur = np.maximum(cur, 0) if l < 3 else 1.0 / (1.0 + np.exp(-cur))
    assert np.allclose(y, cur, rtol=1e-4, atol=1e-5)
