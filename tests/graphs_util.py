"""Random connected graphs as neighbourhood masks (test helper; shared by the CPU and GPU tests of SURVEY 8(f) row 3)."""
import numpy as np


def random_connected_graph(rng, n, p):
    """A random spanning tree plus G(n, p) edges: u32[n] neighbourhood masks."""
    nbr = np.zeros(n, dtype=np.uint32)
    order = rng.permutation(n)
    for i in range(1, n):
        a, b = int(order[i]), int(order[rng.integers(0, i)])
        nbr[a] |= np.uint32(1 << b)
        nbr[b] |= np.uint32(1 << a)
    extra = np.triu(rng.random((n, n)) < p, 1)
    for a, b in zip(*np.nonzero(extra)):
        nbr[a] |= np.uint32(1 << int(b))
        nbr[b] |= np.uint32(1 << int(a))
    return nbr


def edges_of(nbr):
    n = len(nbr)
    return [(v, u) for v in range(n) for u in range(v) if int(nbr[v]) >> u & 1]


def named_graphs():
    """(name, n, edges, lambda_1 of A, mu) with closed forms."""
    import math
    out = []
    for n in (2, 3, 5, 8, 20, 32):
        out.append((f"path{n}", n, [(i, i + 1) for i in range(n - 1)], 2 * math.cos(math.pi / (n + 1)), n // 2))
        out.append((f"star{n}", n, [(0, i) for i in range(1, n)], math.sqrt(n - 1), 1))
        out.append((f"complete{n}", n, [(i, j) for i in range(n) for j in range(i)], float(n - 1), n // 2))
        if n >= 3:
            out.append((f"cycle{n}", n, [(i, (i + 1) % n) for i in range(n)], 2.0, n // 2))
    for a, b in ((3, 17), (1, 31), (16, 16), (5, 6)):
        out.append((f"K{a},{b}", a + b, [(i, a + j) for i in range(a) for j in range(b)], math.sqrt(a * b), min(a, b)))
    # Petersen graph: 3-regular, perfect matching
    pet = [(i, (i + 1) % 5) for i in range(5)] + [(i, i + 5) for i in range(5)] + [(5 + i, 5 + (i + 2) % 5) for i in range(5)]
    out.append(("petersen", 10, pet, 3.0, 5))
    return out
