"""SURVEY 8(f) row 3 on the GPU: azb_eval_graph_costs (through the C ABI) against the oracle's restatement of
ConnectedBitsetGraph::{conjecture_2_1_cost, matching_number, action_kinds}.  mu and the action-kind masks bit-exact;
lambda_1 within 1e-12 relative (north_star allows 1e-5)."""
import numpy as np
import pytest

from graphs_util import named_graphs, random_connected_graph

pytestmark = pytest.mark.gpu


def _mk(capi):
    return capi.Handle(capi.default_config(6, 2, prior_mode=capi.PRIOR_HASH))


def test_named_graphs_closed_forms(capi, orc):
    with _mk(capi) as h:
        for name, n, edges, lam, mu in named_graphs():
            g = orc.graph_from_edges(n, edges)
            if lam + 1e-4 <= 1.4:  # connected_bitset_graph/mod.rs:333 asserts
                with pytest.raises(capi.AzbError) as e:
                    h.eval_graph_costs(g[None, :])
                assert e.value.code == capi.ERR_LAMBDA, name
                continue
            l1, m, kinds, _ = h.eval_graph_costs(g[None, :])
            assert m[0] == mu, name
            assert abs(l1[0] - (lam + 1e-4)) <= 1e-12 * max(1.0, lam), name
            assert np.array_equal(kinds[0], orc.graph_action_kinds(g)), name


@pytest.mark.parametrize("n,p,m", [(4, 0.5, 200), (7, 0.2, 300), (13, 0.1, 300), (19, 0.0, 400), (19, 0.12, 400),
                                   (26, 0.5, 200), (31, 0.08, 300), (32, 0.0, 300), (32, 0.04, 600), (32, 0.3, 300),
                                   (32, 0.95, 100)])
def test_random_graphs_match_oracle(capi, orc, n, p, m):
    rng = np.random.default_rng(77 * n + int(1000 * p))
    graphs = np.stack([random_connected_graph(rng, n, p) for _ in range(m)])
    with _mk(capi) as h:
        l1, mu, kinds, _ = h.eval_graph_costs(graphs)
    for i in range(m):
        lo, mo, rc = orc.graph_cost(graphs[i])
        assert rc == 0
        assert mu[i] == mo, f"graph {i}: mu {mu[i]} vs oracle {mo}"
        assert abs(l1[i] - lo) <= 1e-12 * lo, f"graph {i}: lambda_1 {l1[i]!r} vs oracle {lo!r}"
        assert np.array_equal(kinds[i], orc.graph_action_kinds(graphs[i])), f"graph {i}: action kinds"


def test_invalid_graphs_are_rejected(capi, orc):
    with _mk(capi) as h:
        two_parts = orc.graph_from_edges(4, [(0, 1), (2, 3)])  # to_connected() returns None
        one_way = np.array([0b010, 0b000, 0b000], dtype=np.uint32)
        loop = np.array([0b011, 0b001], dtype=np.uint32)
        for bad in (two_parts, one_way, loop, np.zeros(33, dtype=np.uint32), np.zeros(1, dtype=np.uint32)):
            with pytest.raises(capi.AzbError) as e:
                h.eval_graph_costs(bad[None, :])
            assert e.value.code == capi.ERR_INVALID


def test_one_bad_graph_in_a_batch_is_named(capi, orc):
    rng = np.random.default_rng(3)
    graphs = np.stack([random_connected_graph(rng, 12, 0.2) for _ in range(64)])
    graphs[37, 4] ^= np.uint32(1 << 9)  # one direction of an edge only
    with _mk(capi) as h:
        with pytest.raises(capi.AzbError) as e:
            h.eval_graph_costs(graphs)
        assert e.value.code == capi.ERR_INVALID and "graph 37" in str(e.value)
        graphs[37, 4] ^= np.uint32(1 << 9)
        l1, mu, _, _ = h.eval_graph_costs(graphs)  # the handle stays usable
        assert mu[37] == orc.graph_matching_number(graphs[37])


def test_full_size_properties(capi, orc):
    """65 536 graphs on 32 vertices: relabelling invariance (the cost is a graph invariant, the kinds permute with the
    vertices), counts of the kind masks, and the oracle on a sample."""
    n, m = 32, 65536
    rng = np.random.default_rng(5)
    base = np.stack([random_connected_graph(rng, n, 0.06) for _ in range(512)])
    graphs = np.empty((m, n), dtype=np.uint32)
    perms = np.empty((m, n), dtype=np.int64)
    for i in range(m):
        g = base[i % 512]
        perm = rng.permutation(n) if i >= 512 else np.arange(n)
        perms[i] = perm
        # vertex v -> perm[v]
        bits = (g[:, None] >> np.arange(n, dtype=np.uint32)[None, :]) & 1  # [v, u]
        pb = np.zeros_like(bits)
        pb[np.ix_(perm, perm)] = bits
        graphs[i] = (pb.astype(np.uint64) << np.arange(n, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint32)
    with _mk(capi) as h:
        l1, mu, kinds, ms = h.eval_graph_costs(graphs)
    e2 = n * (n - 1) // 2
    n_edges = np.array([sum(bin(int(x)).count("1") for x in g) // 2 for g in graphs[:512]])
    for i in range(m):
        b = i % 512
        assert mu[i] == mu[b]
        assert abs(l1[i] - l1[b]) <= 1e-12 * l1[b]
    pop = np.array([[bin(int(w)).count("1") for w in row] for row in kinds[:2048]])
    bits = np.unpackbits(kinds[:2048].view(np.uint8), axis=1, bitorder="little")
    adds, dels = bits[:, :e2].sum(axis=1), bits[:, e2:2 * e2].sum(axis=1)
    assert pop.sum(axis=1).tolist() == (adds + dels).tolist()
    for i in range(2048):
        assert adds[i] == e2 - n_edges[i % 512] and dels[i] <= n_edges[i % 512]
        assert dels[i] == dels[i % 512]
    for i in range(0, 512, 8):
        lo, mo, _ = orc.graph_cost(graphs[i])
        assert mu[i] == mo and abs(l1[i] - lo) <= 1e-12 * lo
        assert np.array_equal(kinds[i], orc.graph_action_kinds(graphs[i]))
