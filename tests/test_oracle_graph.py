"""SURVEY 8(f) row 3 -- the oracle's restatement of ConnectedBitsetGraph's cost and action kinds against the
reference's own known answers (connected_bitset_graph/mod.rs:374-422) and against networkx / LAPACK.  CPU only."""
import networkx as nx
import numpy as np
import pytest

from graphs_util import edges_of, named_graphs, random_connected_graph

TWENTY = [(0, 11), (0, 16), (0, 19), (1, 15), (1, 17), (2, 13), (3, 14), (4, 13), (4, 14), (5, 9), (5, 10), (5, 18),
          (6, 15), (7, 17), (7, 19), (8, 10), (9, 12), (10, 13), (16, 18)]


def test_reference_known_answers(orc):
    # mod.rs:374-382, :384-392, :394-422
    k4 = orc.graph_from_edges(4, [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)])
    c5 = orc.graph_from_edges(5, [(0, 1), (1, 2), (2, 3), (3, 4), (4, 0)])
    t20 = orc.graph_from_edges(20, TWENTY)
    assert orc.graph_matching_number(k4) == 2
    assert orc.graph_matching_number(c5) == 2
    assert orc.graph_matching_number(t20) == 9


def test_named_graphs_closed_forms(orc):
    for name, n, edges, lam, mu in named_graphs():
        g = orc.graph_from_edges(n, edges)
        l1, m, rc = orc.graph_cost(g)
        assert m == mu, name
        assert abs(l1 - (lam + 1e-4)) <= 1e-12 * max(1.0, lam), name  # adjacency_matrix's 1e-4 diagonal (mod.rs:205-209)
        assert rc == (0 if l1 > 1.4 else 1), name  # mod.rs:333


@pytest.mark.parametrize("n,p", [(5, 0.3), (9, 0.1), (16, 0.15), (20, 0.0), (24, 0.5), (32, 0.05), (32, 0.3), (32, 0.9)])
def test_random_graphs_against_networkx_and_lapack(orc, n, p):
    rng = np.random.default_rng(1000 * n + int(100 * p))
    for _ in range(25):
        g = random_connected_graph(rng, n, p)
        edges = edges_of(g)
        G = nx.Graph(edges)
        l1, mu, _ = orc.graph_cost(g)
        assert mu == len(nx.max_weight_matching(G, maxcardinality=True))
        a = np.zeros((n, n))
        for v, u in edges:
            a[v, u] = a[u, v] = 1.0
        assert abs(l1 - (np.linalg.eigvalsh(a)[-1] + 1e-4)) <= 1e-12 * l1
        bridges = {(max(e), min(e)) for e in nx.bridges(G)}
        kinds = orc.graph_action_kinds(g)
        e2 = n * (n - 1) // 2
        for v in range(n):
            for u in range(v):
                pos = orc.colex_position(v, u)
                add = int(kinds[pos >> 5]) >> (pos & 31) & 1
                dele = int(kinds[(e2 + pos) >> 5]) >> ((e2 + pos) & 31) & 1
                is_edge = (v, u) in set(edges)
                assert add == (0 if is_edge else 1)
                assert dele == (1 if is_edge and (v, u) not in bridges else 0)
                if is_edge:
                    assert orc.graph_is_cut_edge(g, v, u) == ((v, u) in bridges)


def _brute_force_matching(n, edges):
    """Exact matching number by exhaustive search over edge subsets (small graphs only)."""
    best = 0

    def go(i, used, size):
        nonlocal best
        best = max(best, size)
        if size + (len(edges) - i) <= best:
            return
        for k in range(i, len(edges)):
            a, b = edges[k]
            if not (used >> a & 1) and not (used >> b & 1):
                go(k + 1, used | 1 << a | 1 << b, size + 1)

    go(0, 0, 0)
    return best


def test_small_graphs_against_exhaustive_search(orc):
    """Every connected graph the generator yields on <= 7 vertices: the restated branch and bound equals an exhaustive
    search, relabelling does not change the cost, and the kinds permute with the vertices."""
    rng = np.random.default_rng(11)
    for n in range(2, 8):
        for _ in range(60):
            g = random_connected_graph(rng, n, float(rng.random()) * 0.8)
            edges = edges_of(g)
            assert orc.graph_matching_number(g) == _brute_force_matching(n, edges)
            perm = rng.permutation(n)
            g2 = orc.graph_from_edges(n, [(int(perm[a]), int(perm[b])) for a, b in edges])
            l1, mu, _ = orc.graph_cost(g)
            l2, mu2, _ = orc.graph_cost(g2)
            assert mu == mu2 and abs(l1 - l2) <= 1e-12 * max(1.0, l1)
            k1, k2 = orc.graph_action_kinds(g), orc.graph_action_kinds(g2)
            e2 = n * (n - 1) // 2
            for v in range(n):
                for u in range(v):
                    p1 = orc.colex_position(v, u)
                    p2 = orc.colex_position(int(perm[v]), int(perm[u]))
                    for off in (0, e2):
                        b1 = int(k1[(off + p1) >> 5]) >> ((off + p1) & 31) & 1
                        b2 = int(k2[(off + p2) >> 5]) >> ((off + p2) & 31) & 1
                        assert b1 == b2
