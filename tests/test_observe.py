"""Host-side outputs (azdopt_b200/observe.py) against the reference's own golden vectors and public check values."""
import io

import numpy as np

from azdopt_b200 import observe


def test_graph6_golden_vectors_of_the_reference():
    # graph-state/src/simple_graph/connected_bitset_graph/graph6.rs:47-68
    assert observe.graph6_from_edges(2, [(0, 1)]) == b"A_"
    edges = [
        [(0, 4), (4, 1), (1, 5), (5, 0), (2, 6), (6, 3), (3, 7), (7, 2)],
        [(0, 4), (4, 1), (1, 6), (6, 3), (3, 7), (7, 2), (2, 5), (5, 0)],
        [(0, 3), (3, 5), (5, 0), (1, 4), (4, 7), (7, 2), (2, 6), (6, 1)],
    ]
    assert [observe.graph6_from_edges(8, e) for e in edges] == [b"G?r@`_", b"G?qa`_", b"GCQR@O"]


def test_graph6_of_a_rooted_tree_state():
    # star on 5 vertices (ordered_edge.rs:164-172's tree): vertex 0 adjacent to all -> columns 1..4 each start with a 1
    assert observe.graph6_of_state([0, 0, 0, 0, 0]) == observe.graph6_from_edges(5, [(0, 1), (0, 2), (0, 3), (0, 4)])
    assert observe.graph6_of_state([0, 0, 0, 0, 0]) == b"Ds_"  # bits 1 10 100 1000 -> 110100 100000 (+63)
    path = observe.graph6_of_state([0, 0, 1, 2, 3])
    assert path == observe.graph6_from_edges(5, [(0, 1), (1, 2), (2, 3), (3, 4)])


def test_crc32c_check_value_and_tfrecord_round_trip():
    assert observe.crc32c(b"123456789") == 0xE3069283  # the CRC-32C check value (RFC 3720 B.4)
    buf = io.BytesIO()
    w = observe.TensorboardWriter(buf)
    w.write_file_version()
    w.write_cost(800, 2.7229, 3)
    w.write_loss(800, 0.00321)
    ev = observe.read_events(buf.getvalue())
    assert ev[0][2] == "brain.Event:2" and ev[0][1] == {}
    assert ev[1][0] == 800 and set(ev[1][1]) == {"cost/cost", "cost/lambda_1", "cost/mu"}
    assert abs(ev[1][1]["cost/cost"] - 5.7229) < 1e-6 and ev[1][1]["cost/mu"] == 3.0
    assert ev[2][1] == {"loss": np.float32(0.00321)}


def test_search_tree_dot_labels_and_attributes(orc):
    n, b = 9, 1
    a = orc.action_dim(n)
    parents, masks = orc.generate_roots(1, 0, b, n)
    o = orc.Optimizer(n, b, lambda_method=orc.LAMBDA_MULTISECTION)
    o.set_roots(parents, masks)
    o.init_trees(orc.hash_priors(1, 0, b, a, 0))
    o.steps_hash(1, 0, 1, 12)
    d = o.dump_tree(0)
    dot = observe.search_tree_dot(d)
    assert dot.startswith("graph search_tree {") and dot.rstrip().endswith("}")
    assert f"s0n{int(d['nodes'][0][2])}x{int(d['nodes'][0][3])}" in dot          # graphviz.rs:11-13
    assert dot.count(" -- ") == len(d["arcs"])
    inactive = sum(1 for r in d["nodes"] if int(r[4]) + int(r[3]) >= int(r[5]))
    assert dot.count("shape=doublecircle") == inactive                            # graphviz.rs:19-25
