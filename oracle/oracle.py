"""ctypes binding of the CPU oracle (oracle/libazb_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Nothing under azdopt_b200/
imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libazb_oracle.so")

NONE = 0xFFFFFFFF
LAMBDA_DENSE, LAMBDA_JACOBI, LAMBDA_MULTISECTION, LAMBDA_SECTION_ONLY, LAMBDA_POLY = 0, 1, 2, 3, 4

COUNTER_FIELDS = [
    "n_sel", "d_sel", "n_cur", "n_cand", "n_probe", "n_ins", "n_term", "n_hit", "n_arc", "n_pred",
    "n_cn", "d_cn", "n_reset", "n_live", "n_noop", "n_visit",
]


class Counters(C.Structure):
    _fields_ = [(f, C.c_uint64) for f in COUNTER_FIELDS]

    def as_dict(self):
        return {f: int(getattr(self, f)) for f in COUNTER_FIELDS}


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ only)."""
    src = [os.path.join(_HERE, f) for f in ("azb_oracle.cpp", "azb_oracle.h", "Makefile")]
    stale = not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "libazb_oracle.so"])
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    u8p, u32p, f32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_float)
    L.orc_colex_position.restype = C.c_uint32
    L.orc_colex_position.argtypes = [C.c_uint32, C.c_uint32]
    L.orc_from_colex_position.argtypes = [C.c_uint32, u32p, u32p]
    L.orc_action_dim.restype = C.c_uint32
    L.orc_action_dim.argtypes = [C.c_uint32]
    L.orc_c_upper.restype = C.c_uint32
    L.orc_c_upper.argtypes = [C.c_uint32]
    L.orc_cost.restype = C.c_int
    L.orc_cost.argtypes = [C.c_uint32, u8p, C.c_int, C.c_float, C.c_float, C.POINTER(C.c_double), u32p, f32p]
    L.orc_graph_is_cut_edge.restype = C.c_int
    L.orc_graph_is_cut_edge.argtypes = [C.c_uint32, u32p, C.c_uint32, C.c_uint32]
    L.orc_graph_action_kinds.argtypes = [C.c_uint32, u32p, u32p]
    L.orc_graph_matching_number.restype = C.c_uint32
    L.orc_graph_matching_number.argtypes = [C.c_uint32, u32p]
    L.orc_graph_cost.restype = C.c_int
    L.orc_graph_cost.argtypes = [C.c_uint32, u32p, C.POINTER(C.c_double), u32p]
    L.orc_matching_greedy.restype = C.c_uint32
    L.orc_matching_greedy.argtypes = [C.c_uint32, u8p]
    L.orc_matching_poly.restype = C.c_uint32
    L.orc_matching_poly.argtypes = [C.c_uint32, u8p]
    L.orc_action_data.restype = C.c_uint32
    L.orc_action_data.argtypes = [C.c_uint32, u8p, u32p, u32p]
    L.orc_act.argtypes = [C.c_uint32, u8p, u32p, C.c_uint32]
    L.orc_write_vec.argtypes = [C.c_uint32, u8p, u32p, f32p]
    L.orc_generate_roots.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, u8p, u32p]
    L.orc_hash_priors.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, f32p]
    L.orc_mlp_forward.argtypes = [f32p, u32p, C.c_uint32, f32p, f32p, C.c_int]
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_uint32, C.c_uint32, C.c_float, C.c_float, u32p, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_set_roots.argtypes = [C.c_void_p, u8p, u32p]
    L.orc_init_trees.restype = C.c_int
    L.orc_init_trees.argtypes = [C.c_void_p, f32p]
    L.orc_root_vecs.argtypes = [C.c_void_p, f32p]
    L.orc_reinit_trees.restype = C.c_int
    L.orc_reinit_trees.argtypes = [C.c_void_p, f32p]
    L.orc_modify_roots.restype = C.c_int
    L.orc_modify_roots.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32]
    L.orc_get_roots.argtypes = [C.c_void_p, u8p, u32p]
    L.orc_rollout.restype = C.c_int
    L.orc_rollout.argtypes = [C.c_void_p, f32p]
    L.orc_add_actions.restype = C.c_int
    L.orc_add_actions.argtypes = [C.c_void_p, f32p, C.POINTER(C.c_int)]
    L.orc_steps_hash.restype = C.c_int
    L.orc_steps_hash.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, u32p, C.c_uint32, u32p]
    L.orc_get_counters.argtypes = [C.c_void_p, C.POINTER(Counters)]
    L.orc_reset_counters.argtypes = [C.c_void_p]
    L.orc_get_argmin.argtypes = [C.c_void_p, u8p, u32p, C.POINTER(C.c_double), u32p, f32p]
    L.orc_get_walkers.argtypes = [C.c_void_p, u8p, u32p, u32p, u32p, u32p]
    L.orc_tree_sizes.argtypes = [C.c_void_p, C.c_uint32, u32p, u32p, u32p]
    L.orc_dump_tree.argtypes = [C.c_void_p, C.c_uint32, u32p, u32p, u32p, u32p]
    L.orc_write_observations.argtypes = [C.c_void_p, C.c_uint32, f32p, f32p, f32p]
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def action_dim(n: int) -> int:
    return (n - 1) * (n - 2) // 2 - 1


def mask_words(n: int) -> int:
    return (action_dim(n) + 31) // 32


def c_upper(n: int) -> int:
    return int(lib().orc_c_upper(n))


def colex_position(u, v):
    return int(lib().orc_colex_position(u, v))


def from_colex_position(pos):
    mx, mn = C.c_uint32(), C.c_uint32()
    lib().orc_from_colex_position(pos, C.byref(mx), C.byref(mn))
    return mx.value, mn.value


def cost(parents, method=LAMBDA_DENSE, c_lower=2.0, c_up=None):
    """(lambda_1 f64, mu, c f32, rc) for one parent array."""
    p = np.ascontiguousarray(parents, dtype=np.uint8)
    n = p.shape[0]
    if c_up is None:
        c_up = c_upper(n)
    lam, mu, c = C.c_double(), C.c_uint32(), C.c_float()
    rc = lib().orc_cost(n, _p(p, C.c_uint8), method, c_lower, float(c_up), C.byref(lam), C.byref(mu), C.byref(c))
    return lam.value, mu.value, np.float32(c.value), rc


# ---- connected graphs as neighbourhood bit sets (connected_bitset_graph/mod.rs) ----
def graph_from_edges(n, edges):
    """Neighbourhood masks u32[n] of an edge list (try_from.rs: BitsetGraph from &[(usize, usize)])."""
    nbr = np.zeros(n, dtype=np.uint32)
    for u, v in edges:
        nbr[u] |= np.uint32(1 << v)
        nbr[v] |= np.uint32(1 << u)
    return nbr


def graph_kind_words(n: int) -> int:
    return (n * (n - 1) + 31) // 32


def graph_is_cut_edge(nbr, v, u) -> bool:
    a = np.ascontiguousarray(nbr, dtype=np.uint32)
    return bool(lib().orc_graph_is_cut_edge(a.shape[0], _p(a, C.c_uint32), v, u))


def graph_action_kinds(nbr):
    """Bit colex(e): Add(e) available; bit C(n,2)+colex(e): Delete(e) available (mod.rs:134-154, action.rs:10-19)."""
    a = np.ascontiguousarray(nbr, dtype=np.uint32)
    out = np.zeros(graph_kind_words(a.shape[0]), dtype=np.uint32)
    lib().orc_graph_action_kinds(a.shape[0], _p(a, C.c_uint32), _p(out, C.c_uint32))
    return out


def graph_matching_number(nbr) -> int:
    a = np.ascontiguousarray(nbr, dtype=np.uint32)
    return int(lib().orc_graph_matching_number(a.shape[0], _p(a, C.c_uint32)))


def graph_cost(nbr):
    """(lambda_1 of A + 1e-4 I, mu, rc) for one connected graph (mod.rs:319-337)."""
    a = np.ascontiguousarray(nbr, dtype=np.uint32)
    lam, mu = C.c_double(), C.c_uint32()
    rc = lib().orc_graph_cost(a.shape[0], _p(a, C.c_uint32), C.byref(lam), C.byref(mu))
    return lam.value, mu.value, rc


def matching_greedy(parents):
    p = np.ascontiguousarray(parents, dtype=np.uint8)
    return int(lib().orc_matching_greedy(p.shape[0], _p(p, C.c_uint8)))


def matching_poly(parents):
    p = np.ascontiguousarray(parents, dtype=np.uint8)
    return int(lib().orc_matching_poly(p.shape[0], _p(p, C.c_uint8)))


def mask_from_actions(n, actions):
    m = np.zeros(mask_words(n), dtype=np.uint32)
    for a in actions:
        m[a >> 5] |= np.uint32(1 << (a & 31))
    return m


def actions_from_mask(mask):
    out = []
    for w, word in enumerate(np.asarray(mask, dtype=np.uint32)):
        word = int(word)
        while word:
            b = (word & -word).bit_length() - 1
            out.append(w * 32 + b)
            word &= word - 1
    return out


def action_data(parents, mask):
    p = np.ascontiguousarray(parents, dtype=np.uint8)
    m = np.ascontiguousarray(mask, dtype=np.uint32)
    out = np.zeros(action_dim(p.shape[0]), dtype=np.uint32)
    k = lib().orc_action_data(p.shape[0], _p(p, C.c_uint8), _p(m, C.c_uint32), _p(out, C.c_uint32))
    return out[:k].tolist()


def act(parents, mask, action):
    p = np.array(parents, dtype=np.uint8)
    m = np.array(mask, dtype=np.uint32)
    lib().orc_act(p.shape[0], _p(p, C.c_uint8), _p(m, C.c_uint32), action)
    return p, m


def write_vec(parents, mask):
    p = np.ascontiguousarray(parents, dtype=np.uint8)
    m = np.ascontiguousarray(mask, dtype=np.uint32)
    v = np.zeros(2 * action_dim(p.shape[0]), dtype=np.float32)
    lib().orc_write_vec(p.shape[0], _p(p, C.c_uint8), _p(m, C.c_uint32), _p(v, C.c_float))
    return v


def generate_roots(seed, first_root, count, n, k_min=5, k_max=None):
    if k_max is None:
        k_max = action_dim(n) // 2
    parents = np.zeros((count, n), dtype=np.uint8)
    masks = np.zeros((count, mask_words(n)), dtype=np.uint32)
    lib().orc_generate_roots(seed, first_root, count, n, k_min, k_max, _p(parents, C.c_uint8), _p(masks, C.c_uint32))
    return parents, masks


def hash_priors(seed, first_root, count, a_dim, step):
    out = np.zeros((count, a_dim), dtype=np.float32)
    lib().orc_hash_priors(seed, first_root, count, a_dim, step, _p(out, C.c_float))
    return out


def mlp_forward(params, dims, x, n_threads=1):
    params = np.ascontiguousarray(params, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    d = np.asarray(dims, dtype=np.uint32)
    y = np.zeros((x.shape[0], int(d[4])), dtype=np.float32)
    lib().orc_mlp_forward(_p(params, C.c_float), _p(d, C.c_uint32), x.shape[0], _p(x, C.c_float), _p(y, C.c_float), n_threads)
    return y


class Optimizer:
    """Mirror of NablaOptimizer (optimizer/mod.rs) over the oracle."""

    def __init__(self, n, n_roots, n_as_tol=(200, 50, 50), n_as_tol_default=25, c_lower=2.0, c_up=None,
                 lambda_method=LAMBDA_DENSE, n_threads=1):
        self.n, self.b = n, n_roots
        self.a = action_dim(n)
        self.w = mask_words(n)
        if c_up is None:
            c_up = c_upper(n)
        tol = np.asarray(n_as_tol, dtype=np.uint32)
        self._h = lib().orc_create(n, n_roots, c_lower, float(c_up), _p(tol, C.c_uint32), len(tol), n_as_tol_default,
                                   lambda_method, n_threads)
        if not self._h:
            raise ValueError("orc_create failed")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_roots(self, parents, masks):
        p = np.ascontiguousarray(parents, dtype=np.uint8)
        m = np.ascontiguousarray(masks, dtype=np.uint32)
        assert p.shape == (self.b, self.n) and m.shape == (self.b, self.w)
        lib().orc_set_roots(self._h, _p(p, C.c_uint8), _p(m, C.c_uint32))

    def init_trees(self, priors):
        pr = np.ascontiguousarray(priors, dtype=np.float32)
        assert pr.shape == (self.b, self.a)
        rc = lib().orc_init_trees(self._h, _p(pr, C.c_float))
        if rc:
            raise RuntimeError(f"orc_init_trees rc={rc}")

    def reinit_trees(self, priors):
        pr = np.ascontiguousarray(priors, dtype=np.float32)
        assert pr.shape == (self.b, self.a)
        rc = lib().orc_reinit_trees(self._h, _p(pr, C.c_float))
        if rc:
            raise RuntimeError(f"orc_reinit_trees rc={rc}")

    def modify_roots(self, seed, epoch, first_root=0, k_min=5, k_max=None):
        """par_reset_trees' modify_root half (example policy); follow with reinit_trees."""
        rc = lib().orc_modify_roots(self._h, seed, epoch, first_root, k_min, self.a // 2 if k_max is None else k_max)
        if rc:
            raise RuntimeError(f"orc_modify_roots rc={rc}")

    def get_roots(self):
        p = np.zeros((self.b, self.n), dtype=np.uint8)
        m = np.zeros((self.b, self.w), dtype=np.uint32)
        lib().orc_get_roots(self._h, _p(p, C.c_uint8), _p(m, C.c_uint32))
        return p, m

    def root_vecs(self):
        v = np.zeros((self.b, 2 * self.a), dtype=np.float32)
        lib().orc_root_vecs(self._h, _p(v, C.c_float))
        return v

    def rollout(self, state_vecs=None):
        ptr = _p(state_vecs, C.c_float) if state_vecs is not None else None
        rc = lib().orc_rollout(self._h, ptr)
        if rc:
            raise RuntimeError(f"orc_rollout rc={rc}")

    def add_actions(self, priors):
        pr = np.ascontiguousarray(priors, dtype=np.float32)
        imp = C.c_int()
        lib().orc_add_actions(self._h, _p(pr, C.c_float), C.byref(imp))
        return bool(imp.value)

    def steps_hash(self, seed, first_root, step0, n_steps):
        imp = np.zeros(max(1, n_steps), dtype=np.uint32)
        n_imp = C.c_uint32()
        rc = lib().orc_steps_hash(self._h, seed, first_root, step0, n_steps, _p(imp, C.c_uint32), len(imp), C.byref(n_imp))
        if rc:
            raise RuntimeError(f"orc_steps_hash rc={rc}")
        return imp[: n_imp.value].tolist()

    def counters(self):
        c = Counters()
        lib().orc_get_counters(self._h, C.byref(c))
        return c.as_dict()

    def reset_counters(self):
        lib().orc_reset_counters(self._h)

    def argmin(self):
        p = np.zeros(self.n, dtype=np.uint8)
        m = np.zeros(self.w, dtype=np.uint32)
        lam, mu, ev = C.c_double(), C.c_uint32(), C.c_float()
        lib().orc_get_argmin(self._h, _p(p, C.c_uint8), _p(m, C.c_uint32), C.byref(lam), C.byref(mu), C.byref(ev))
        return dict(parents=p, permitted=m, lambda1=lam.value, mu=mu.value, eval=np.float32(ev.value))

    def walkers(self):
        p = np.zeros((self.b, self.n), dtype=np.uint8)
        m = np.zeros((self.b, self.w), dtype=np.uint32)
        k = np.zeros((self.b, self.w), dtype=np.uint32)
        pos = np.zeros(self.b, dtype=np.uint32)
        ln = np.zeros(self.b, dtype=np.uint32)
        lib().orc_get_walkers(self._h, _p(p, C.c_uint8), _p(m, C.c_uint32), _p(k, C.c_uint32), _p(pos, C.c_uint32),
                              _p(ln, C.c_uint32))
        return dict(parents=p, permitted=m, path=k, pos=pos, path_len=ln)

    def tree_sizes(self, tree):
        a, b, c = C.c_uint32(), C.c_uint32(), C.c_uint32()
        lib().orc_tree_sizes(self._h, tree, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def dump_tree(self, tree):
        nn, na, npred = self.tree_sizes(tree)
        nodes = np.zeros((nn, 6), dtype=np.uint32)
        keys = np.zeros((nn, self.w), dtype=np.uint32)
        preds = np.zeros((npred, 3), dtype=np.uint32)
        arcs = np.zeros((na, 3), dtype=np.uint32)
        lib().orc_dump_tree(self._h, tree, _p(nodes, C.c_uint32), _p(keys, C.c_uint32), _p(preds, C.c_uint32),
                            _p(arcs, C.c_uint32))
        return dict(nodes=nodes, keys=keys, preds=preds, arcs=arcs)

    def write_observations(self, n_obs_tol):
        v = np.zeros((self.b, 2 * self.a), dtype=np.float32)
        obs = np.zeros((self.b, self.a), dtype=np.float32)
        w = np.zeros((self.b, self.a), dtype=np.float32)
        lib().orc_write_observations(self._h, n_obs_tol, _p(v, C.c_float), _p(obs, C.c_float), _p(w, C.c_float))
        return v, obs, w


# ---- epoch boundary: the training step (numpy restatement; TEST INFRASTRUCTURE like the rest of this file) -------
class AdamState:
    """dfdx 0.13 Adam (tensor_ops/adam; optim/adam): moments per parameter and the step counter t."""

    def __init__(self, n_params, lr=1e-4, beta1=0.9, beta2=0.999, eps=1e-8, l2=1e-6):  # 04-c21-tree.rs:87-92
        self.m = np.zeros(n_params, dtype=np.float32)
        self.v = np.zeros(n_params, dtype=np.float32)
        self.t = 0
        self.lr, self.beta1, self.beta2, self.eps, self.l2 = (np.float32(x) for x in (lr, beta1, beta2, eps, l2))


def _split_params(params, dims):
    out, off = [], 0
    for l in range(4):
        din, dout = int(dims[l]), int(dims[l + 1])
        w = params[off:off + din * dout].reshape(dout, din)  # weight[out][in] (dfdx Linear)
        b = params[off + din * dout:off + din * dout + dout]
        out.append((w, b, off))
        off += din * dout + dout
    return out


def model_gradients(params, dims, states, observations, weights):
    """loss and d loss / d params of ActionModel::update_model's objective (nabla/model/dfdx.rs:104-126), f32:
    w_n = w / sum(w); loss = sum (forward(states) - observations)^2 * w_n."""
    params = np.ascontiguousarray(params, dtype=np.float32)
    x = np.ascontiguousarray(states, dtype=np.float32)
    o = np.ascontiguousarray(observations, dtype=np.float32)
    w = np.ascontiguousarray(weights, dtype=np.float32)
    layers = _split_params(params, dims)
    acts = [x]
    for l, (wt, b, _) in enumerate(layers):
        z = acts[-1] @ wt.T + b
        acts.append(np.maximum(z, np.float32(0)) if l < 3 else (np.float32(1) / (np.float32(1) + np.exp(-z))).astype(np.float32))
    wsum = np.float32(w.sum(dtype=np.float64))  # dfdx.rs:105 (exact for 0/1 weights below 2^24)
    wn = (w / wsum).astype(np.float32)
    p = acts[-1]
    d = p - o
    loss = np.float32((d * d * wn).sum(dtype=np.float64))
    dz = (np.float32(2) * d * wn * (p * (np.float32(1) - p))).astype(np.float32)
    grads = np.zeros_like(params)
    for l in range(3, -1, -1):
        wt, b, off = layers[l]
        din, dout = wt.shape[1], wt.shape[0]
        grads[off:off + din * dout] = (dz.T @ acts[l]).reshape(-1)
        grads[off + din * dout:off + din * dout + dout] = dz.sum(axis=0, dtype=np.float32)
        if l > 0:
            dz = ((dz @ wt) * (acts[l] > 0)).astype(np.float32)
    return loss, grads


def adam_step(params, grads, st: AdamState):
    """One dfdx Adam update with WeightDecay::L2 (g += l2 p before the moments)."""
    one = np.float32(1)
    st.t += 1
    g = (grads + st.l2 * params).astype(np.float32)
    st.m = (st.m * st.beta1 + g * (one - st.beta1)).astype(np.float32)
    st.v = (st.v * st.beta2 + g * g * (one - st.beta2)).astype(np.float32)
    mh = st.m * np.float32(1.0 / (1.0 - float(st.beta1) ** st.t))
    vh = st.v * np.float32(1.0 / (1.0 - float(st.beta2) ** st.t))
    return (params - st.lr * mh / (np.sqrt(vh) + st.eps)).astype(np.float32)


def update_model(params, dims, states, observations, weights, st: AdamState):
    """NablaModel::update_model (nabla/model/dfdx.rs:86-131): returns (loss, new params)."""
    loss, grads = model_gradients(params, dims, states, observations, weights)
    return loss, adam_step(np.ascontiguousarray(params, dtype=np.float32), grads, st)
