"""Independent pure-Python restatement of the reference's search step, for SMALL cases only.

Written directly from the reference sources (not from oracle/azb_oracle.cpp) so that the two restatements check
each other: az-discrete-opt/src/nabla/tree/{mod,next_action,graph_operations,empty_transitions}.rs,
nabla/optimizer/mod.rs, graph-state/src/rooted_tree/{space,ordered_edge,mod}.rs, examples/04-c21-tree.rs.
All f32 arithmetic goes through numpy.float32 scalars, one rounding per operation.
"""
from __future__ import annotations

import numpy as np

F = np.float32


def action_dim(n):
    return (n - 1) * (n - 2) // 2 - 1


def colex(mx, mn):  # simple_graph/edge.rs:48-53
    return mx * (mx + 1) // 2 - (mx - mn)


def from_colex(pos):  # simple_graph/edge.rs:55-65
    v = 1
    while True:
        last = v * (v + 1) // 2
        if pos < last:
            return v, v - (last - pos)
        v += 1


def action_edge(a):  # ordered_edge.rs:40-42 -> (parent, child)
    mx, mn = from_colex(a + 1)
    return mn, mx


def edge_action(parent, child):  # ordered_edge.rs:35-38
    return colex(child, parent) - 1


class State:
    def __init__(self, parents, permitted):
        self.parents = list(int(p) for p in parents)
        self.permitted = set(int(a) for a in permitted)

    def clone(self):
        return State(self.parents, self.permitted)


def act(n, s, a):  # rooted_tree/space.rs:56-73
    parent, child = action_edge(a)
    s.parents[child] = parent
    for u in range(child):
        s.permitted.discard(edge_action(u, child))


def current_edges(n, s):  # rooted_tree/mod.rs:60-72
    return [edge_action(s.parents[c], c) for c in range(2, n - 1)]


def action_data(n, s):  # rooted_tree/space.rs:75-89
    cur = current_edges(n, s)
    return [a for a in sorted(s.permitted) if a not in cur]


def write_vec(n, s):  # rooted_tree/space.rs:91-101
    a_dim = action_dim(n)
    v = np.zeros(2 * a_dim, dtype=np.float32)
    for e in current_edges(n, s):
        v[e] = 1.0
    for a in s.permitted:
        v[a_dim + a] = 1.0
    return v


def maximum_matching(n, parents):  # ordered_edge.rs:94-124
    available = [True] * n
    m = 0
    while True:
        next_leaf = list(available)
        for i in range(1, n):
            if available[i]:
                next_leaf[parents[i]] = False
        for i in range(1, n):
            if next_leaf[i]:
                available[i] = False
                p = parents[i]
                if available[p]:
                    available[p] = False
                    m += 1
        if sum(available) < 2:
            break
    return m


def lambda1(n, parents):  # ordered_edge.rs:72-91 (dense symmetric eigenvalues; numpy/LAPACK here)
    a = np.zeros((n, n))
    for i in range(1, n):
        a[i, parents[i]] = 1.0
        a[parents[i], i] = 1.0
    return float(np.linalg.eigvalsh(a).max())


def c_upper(n):  # 04-c21-tree.rs:59-68
    import math

    s = math.isqrt(n - 1)
    sq = s if s * s == n - 1 else s + 1
    return sq + (n + 1) // 2


def evaluate(n, mu, lam):  # 04-c21-tree.rs:70-74,98-102
    slope = F(1.0) / F(c_upper(n) - 2)
    x = F(mu) + F(lam)
    x = x - F(2)
    return F(slope * x)


class Node:
    def __init__(self, c):  # state_weight.rs:13-21
        self.c = F(c)
        self.cstar = F(c)
        self.nt = 0
        self.ex = 0
        self.lo = 0
        self.hi = 0
        self.out = []  # arc ids, newest first (petgraph)
        self.inn = []

    def active(self):  # state_weight.rs:31-33
        return self.lo + self.ex < self.hi


class Tree:
    def __init__(self):
        self.positions = {}
        self.nodes = []
        self.arcs = []  # (src, dst, prediction_pos)
        self.preds = []  # [a_id, g, arc or None]
        self.keys = []

    def add_node(self, key, c):  # graph_operations.rs:8-16
        self.nodes.append(Node(c))
        idx = len(self.nodes) - 1
        assert key not in self.positions
        self.positions[key] = idx
        self.keys.append(key)
        return idx

    def add_arc(self, src, dst, ppos):  # graph_operations.rs:18-30
        self.arcs.append((src, dst, ppos))
        e = len(self.arcs) - 1
        self.nodes[src].out.insert(0, e)
        self.nodes[dst].inn.insert(0, e)
        self.preds[ppos][2] = e
        return e

    def add_actions(self, n, idx, state, h):  # graph_operations.rs:32-56
        node = self.nodes[idx]
        start = len(self.preds)
        for a in action_data(n, state):
            self.preds.append([a, F(node.c - F(h[a])), None])
        node.lo, node.hi = start, len(self.preds)

    def revisit_choice(self, pos):  # next_action.rs:28-53
        best = None
        for e in self.nodes[pos].out:
            ch = self.nodes[self.arcs[e][1]]
            if not ch.active():
                continue
            cand = (e, ch.nt, ch.cstar)
            if best is None or (cand[1], cand[2]) < (best[1], best[2]):  # min_by keeps the first minimum
                best = cand
        return best

    def max_curiosity(self, pos):  # next_action.rs:55-88
        node = self.nodes[pos]
        cs = [self.nodes[self.arcs[e][1]].cstar for e in node.out]
        cands = [(j, F(node.c - self.preds[j][1])) for j in range(node.lo, node.hi) if self.preds[j][2] is None]
        if not cands:
            return None
        if not cs:
            best = cands[0]
            for c in cands[1:]:
                if c[1] < best[1]:
                    best = c
            return best[0]
        best = None
        for j, v in cands:
            cur = F(0)
            for x in cs:
                cur = F(cur + np.sqrt(np.abs(F(x - v))))
            if best is None or cur >= best[1]:  # max_by keeps the last maximum
                best = (j, cur)
        return best[0]

    def next_action(self, pos, tol):  # next_action.rs:11-26
        if not self.nodes[pos].active():
            return None
        r = self.revisit_choice(pos)
        if r is not None and r[1] < tol:
            return ("V", r[0])
        j = self.max_curiosity(pos)
        if j is not None:
            return ("U", j)
        return ("V", r[0]) if r is not None else None

    def cascade(self, arc, old):  # empty_transitions.rs:50-127
        src, dst, _ = self.arcs[arc]
        t = self.nodes[dst]
        ntt = t.nt
        cst = t.cstar
        e0 = (0 if t.active() else 1) if old else 1
        cur = {src: [cst, e0]}
        nxt = {}
        while True:
            if not cur:
                cur, nxt = nxt, {}
                if not cur:
                    break
            idx = min(cur)
            cstar, e = cur.pop(idx)
            nd = self.nodes[idx]
            nd.ex += e
            if nd.cstar > cstar:
                nd.cstar = cstar
            else:
                nd.nt += 1
            if old:
                nd.nt = max(nd.nt, ntt)
            up = [cstar, 0 if nd.active() else 1]
            for ie in nd.inn:
                p = self.arcs[ie][0]
                if p in nxt:
                    nxt[p][0] = min(nxt[p][0], up[0])
                    nxt[p][1] += up[1]
                else:
                    nxt[p] = list(up)


class Optimizer:
    """NablaOptimizer (optimizer/mod.rs) with injected priors instead of a model."""

    def __init__(self, n, parents, permitted_lists, tol=(200, 50, 50), tol_default=25):
        self.n = n
        self.roots = [State(p, a) for p, a in zip(parents, permitted_lists)]
        self.tol, self.tol_default = list(tol), tol_default

    def cost_eval(self, s):
        return evaluate(self.n, maximum_matching(self.n, s.parents), lambda1(self.n, s.parents))

    def init_trees(self, priors):  # optimizer/mod.rs:62-101
        self.states = [r.clone() for r in self.roots]
        self.paths = [set() for _ in self.roots]
        self.pos = [0 for _ in self.roots]
        self.trees = []
        evals = []
        for i, r in enumerate(self.roots):
            t = Tree()
            c = self.cost_eval(r)
            evals.append(c)
            t.add_node((), c)
            t.add_actions(self.n, 0, r, priors[i])
            self.trees.append(t)
        self.inspected = [0] * len(self.roots)
        self.best = min(evals)

    def n_as_tol(self, depth):
        return self.tol[depth] if depth < len(self.tol) else self.tol_default

    def rollout(self, i):  # tree/mod.rs:113-232
        t, root, n = self.trees[i], self.roots[i], self.n
        while True:
            na = t.next_action(self.pos[i], self.n_as_tol(len(self.paths[i])))
            if na is None:
                assert not self.paths[i]
                return
            kind, x = na
            if kind == "V":
                a = t.preds[t.arcs[x][2]][0]
                self.paths[i].add(a)
                act(n, self.states[i], a)
                self.pos[i] = t.arcs[x][1]
                continue
            a = t.preds[x][0]
            self.paths[i].add(a)
            key = tuple(sorted(self.paths[i]))
            if key in t.positions:
                arc = t.add_arc(self.pos[i], t.positions[key], x)
                t.cascade(arc, True)
                self.states[i] = root.clone()
                self.paths[i] = set()
                self.pos[i] = 0
                continue
            act(n, self.states[i], a)
            c = self.cost_eval(self.states[i])
            idx = t.add_node(key, c)
            arc = t.add_arc(self.pos[i], idx, x)
            if not action_data(n, self.states[i]):
                t.cascade(arc, False)
                self.states[i] = root.clone()
                self.paths[i] = set()
                self.pos[i] = 0
                continue
            self.pos[i] = idx
            return

    def step(self, priors):  # optimizer/mod.rs:121-246
        for i in range(len(self.roots)):
            self.rollout(i)
        for i in range(len(self.roots)):
            if self.paths[i]:
                self.trees[i].add_actions(self.n, self.pos[i], self.states[i], priors[i])
        improved = None
        for i, t in enumerate(self.trees):
            best_t = None
            for j in range(self.inspected[i], len(t.nodes)):
                c = t.nodes[j].c
                if c < self.best and (best_t is None or c < best_t[0]):
                    best_t = (c, j)
            self.inspected[i] = len(t.nodes)
            if best_t is not None and (improved is None or best_t[0] < improved[0]):
                improved = (best_t[0], i, best_t[1])
        if improved is not None:
            self.best = improved[0]
        return improved

    def dump(self, i, words):
        t = self.trees[i]
        nodes = np.zeros((len(t.nodes), 6), dtype=np.uint32)
        keys = np.zeros((len(t.nodes), words), dtype=np.uint32)
        for k, nd in enumerate(t.nodes):
            nodes[k] = [np.float32(nd.c).view(np.uint32), np.float32(nd.cstar).view(np.uint32), nd.nt, nd.ex, nd.lo, nd.hi]
            for a in t.keys[k]:
                keys[k, a >> 5] |= np.uint32(1 << (a & 31))
        preds = np.zeros((len(t.preds), 3), dtype=np.uint32)
        for k, (a, g, e) in enumerate(t.preds):
            preds[k] = [a, np.float32(g).view(np.uint32), 0xFFFFFFFF if e is None else e]
        arcs = np.array(t.arcs, dtype=np.uint32).reshape(-1, 3)
        return dict(nodes=nodes, keys=keys, preds=preds, arcs=arcs)
