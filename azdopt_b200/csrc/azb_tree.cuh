// azb_tree.cuh — the search step: one warp owns one root's search DAG.
//
// Replaces, per root, SearchTree::roll_out_episodes (az-discrete-opt/src/nabla/tree/mod.rs:113-232) with
// next_action / revisit_choice / max_curiosity (nabla/tree/next_action.rs:11-88), add_node / add_arc / add_actions
// (nabla/tree/graph_operations.rs:8-56), cascade_new_terminal / cascade_old_node
// (nabla/tree/empty_transitions.rs:50-127), the ROTModifyParentsOnce space (graph-state/src/rooted_tree/space.rs:
// 37-125), write_vec (space.rs:91-101) and the argmin scan (nabla/optimizer/mod.rs:194-246).
//
// Design notes:
//  * out-arcs of a node live in the node's own prediction range, in creation order: kid[lo + t] is the t-th arc.
//    petgraph iterates newest-first, so "first minimum" == highest t among equal keys and the curiosity sum runs
//    t = n_out-1 .. 0.  One coalesced read replaces a linked-list walk.
//  * in-arcs (parents) of a node live in `depth` reserved slots (a node of depth d has at most d parents, because
//    keys are action sets and every parent key is the child key minus one action).
//  * the transposition map BTreeMap<ActionSet, NodeIndex> is an open-addressing table keyed by the W-word mask.
//  * walker state (parents, permitted mask, path mask) sits in shared memory while the warp runs.
#pragma once
#include "azb_common.cuh"
#include "azb_cost.cuh"

enum { AZB_F_ADD = 1, AZB_F_ROLLOUT = 2, AZB_F_INIT = 4, AZB_F_FIRST = 8 };

struct WarpCtx {
    uint32_t *wk;      // walker block copy [WS]
    uint8_t *par;      // parents (bytes) inside wk
    uint32_t *perm;    // permitted mask [W]
    uint32_t *keym;    // path mask [W]
    uint8_t *rpar;     // root parents
    uint32_t *rperm;   // root permitted
    uint32_t *cur;     // scratch mask [W]: current-edge mask
    uint32_t *pfx;     // scratch [W+1]
    float *lbuf;       // children's c*, creation order [LCAP]
    uint32_t *fr;      // frontier buffers [4][FRONTIER_CAP]
    CostScratch *cs;   // lambda_1 program scratch
    // per-tree slabs
    uint32_t *node;    // 8 words per node
    uint2 *pred;
    uint2 *kid;
    uint32_t *arcseq;
    uint32_t *inl;
    uint32_t *key;
    uint32_t *hash;
    int lane;
    uint32_t err;
    uint32_t ct[16];
};

__device__ __forceinline__ uint32_t nth_set_bit(uint32_t mask, uint32_t r) {
    for (uint32_t i = 0; i < r; ++i) mask &= mask - 1;
    return __ffs(mask) - 1;
}

__device__ __forceinline__ uint32_t lanemask_lt(int lane) { return (1u << lane) - 1u; }

// hash of a W-word action-set mask held as (k0 = word lane, k1 = word lane+32) across the warp
__device__ __forceinline__ uint32_t key_hash(uint32_t k0, uint32_t k1, int lane, uint32_t W) {
    uint32_t hv = (k0 * 0x9E3779B1u + (uint32_t)lane * 0x85EBCA77u) ^ ((k1 + 0x7F4A7C15u) * 0xC2B2AE3Du);
    hv ^= hv >> 15;
    hv *= 0x2C1B3C6Du;
    hv = (uint32_t)lane < W ? hv : 0u;
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) hv ^= __shfl_xor_sync(0xffffffffu, hv, m);
    hv ^= hv >> 13;
    hv *= 0x297A2D39u;
    hv ^= hv >> 16;
    return hv;
}

// current-edge mask of the walker state into cx.cur (rooted_tree/mod.rs:60-72)
__device__ __forceinline__ void build_cur_mask(const AzbLayout &L, WarpCtx &cx) {
    for (uint32_t w = cx.lane; w < L.W; w += 32) cx.cur[w] = 0u;
    __syncwarp();
    for (uint32_t c = 2 + cx.lane; c + 1 < L.N; c += 32) {
        uint32_t e = azb_action_index(cx.par[c], c);
        atomicOr(&cx.cur[e >> 5], 1u << (e & 31));
    }
    __syncwarp();
}

// act (rooted_tree/space.rs:56-73): set the parent, drop every action of that child, extend the path set
__device__ __forceinline__ void walker_act(const AzbLayout &L, WarpCtx &cx, uint32_t a) {
    uint32_t child = azb_action_child(a);
    uint32_t first = azb_child_first_action(child);
    uint32_t last = first + child;  // exclusive
    if (cx.lane == 0) cx.par[child] = (uint8_t)(a - first);
    for (uint32_t w = cx.lane; w < L.W; w += 32) {
        uint32_t b0 = w * 32, b1 = b0 + 32;
        uint32_t s = max(first, b0), e = min(last, b1);
        if (s < e) {
            uint32_t len = e - s;
            uint32_t m = (len == 32 ? 0xffffffffu : ((1u << len) - 1u)) << (s - b0);
            cx.perm[w] &= ~m;
        }
        if (w == (a >> 5)) cx.keym[w] |= 1u << (a & 31);
    }
    __syncwarp();
}

__device__ __forceinline__ void walker_reset(const AzbLayout &L, WarpCtx &cx) {  // tree/mod.rs:176-178
    for (uint32_t i = cx.lane; i < L.PW; i += 32) ((uint32_t *)cx.par)[i] = ((const uint32_t *)cx.rpar)[i];
    for (uint32_t w = cx.lane; w < L.W; w += 32) {
        cx.perm[w] = cx.rperm[w];
        cx.keym[w] = 0u;
    }
    __syncwarp();
}

// the 8 words of node `id`, broadcast to every lane
struct NodeRec {
    float c, cstar;
    uint32_t nt, ex, cnt, lo, n_out, n_in, in_off, depth;
};
__device__ __forceinline__ NodeRec load_node(const WarpCtx &cx, uint32_t id) {
    uint32_t w = cx.lane < 8 ? cx.node[(size_t)id * 8 + cx.lane] : 0u;
    NodeRec r;
    r.c = __uint_as_float(__shfl_sync(0xffffffffu, w, ND_C));
    r.cstar = __uint_as_float(__shfl_sync(0xffffffffu, w, ND_CSTAR));
    r.nt = __shfl_sync(0xffffffffu, w, ND_NT);
    uint32_t excnt = __shfl_sync(0xffffffffu, w, ND_EXCNT);
    r.ex = excnt & 0xffffu;
    r.cnt = excnt >> 16;
    r.lo = __shfl_sync(0xffffffffu, w, ND_LO);
    uint32_t outin = __shfl_sync(0xffffffffu, w, ND_OUTIN);
    r.n_out = outin & 0xffffu;
    r.n_in = outin >> 16;
    r.in_off = __shfl_sync(0xffffffffu, w, ND_INOFF);
    r.depth = __shfl_sync(0xffffffffu, w, ND_DEPTH);
    return r;
}

__device__ __forceinline__ float prior_of(const AzbLayout &L, uint32_t tree, uint32_t a, uint32_t prior_step) {
    if (L.prior_mode == 1) return azb_hash_prior(L.prior_seed, L.first_root + tree, prior_step, a);
    return L.h[(size_t)tree * L.h_ld + a];
}

// add_actions (graph_operations.rs:32-56): one prediction per legal action, ascending; g = c_s - h (04-c21-tree.rs:103)
__device__ void tree_add_actions(const AzbLayout &L, WarpCtx &cx, uint32_t tree, uint32_t prior_step) {
    const int lane = cx.lane;
    uint32_t pos = cx.wk[WK_POS];
    float c_s = __uint_as_float(cx.wk[10]);
    build_cur_mask(L, cx);
    // legal = permitted minus current edges (space.rs:75-89); counts per word -> exclusive prefix
    uint32_t l0 = (uint32_t)lane < L.W ? (cx.perm[lane] & ~cx.cur[lane]) : 0u;
    uint32_t l1 = (uint32_t)lane + 32 < L.W ? (cx.perm[lane + 32] & ~cx.cur[lane + 32]) : 0u;
    uint32_t s0 = __popc(l0), s1 = __popc(l1);
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t0 = __shfl_up_sync(0xffffffffu, s0, d), t1 = __shfl_up_sync(0xffffffffu, s1, d);
        if (lane >= d) {
            s0 += t0;
            s1 += t1;
        }
    }
    uint32_t tot0 = __shfl_sync(0xffffffffu, s0, 31);
    uint32_t cnt = tot0 + __shfl_sync(0xffffffffu, s1, 31);
    __syncwarp();
    if ((uint32_t)lane < L.W) {
        cx.cur[lane] = l0;
        cx.pfx[lane + 1] = s0;
    }
    if ((uint32_t)lane + 32 < L.W) {
        cx.cur[lane + 32] = l1;
        cx.pfx[lane + 33] = tot0 + s1;
    }
    if (lane == 0) cx.pfx[0] = 0u;
    __syncwarp();
    uint32_t lo = cx.wk[WK_NPREDS];
    if (lo + cnt > L.cap_preds || cnt > 0xffffu) {
        cx.err = 3;
        return;
    }
    for (uint32_t j = lane; j < cnt; j += 32) {
        // word holding the j-th legal action: largest w with pfx[w] <= j
        uint32_t a0 = 0, a1 = L.W;  // invariant pfx[a0] <= j < pfx[a1]
        while (a1 - a0 > 1) {
            uint32_t mid = (a0 + a1) >> 1;
            if (cx.pfx[mid] <= j)
                a0 = mid;
            else
                a1 = mid;
        }
        uint32_t a = a0 * 32 + nth_set_bit(cx.cur[a0], j - cx.pfx[a0]);
        float h = prior_of(L, tree, a, prior_step);
        if (h != h) cx.err = 4;
        float g = __fsub_rn(c_s, h);
        cx.pred[lo + j] = make_uint2(__float_as_uint(g), a);
    }
    cx.err = __reduce_or_sync(0xffffffffu, cx.err);
    if (lane == 0) {
        cx.node[(size_t)pos * 8 + ND_EXCNT] = cnt << 16;  // exhausted_children = 0
        cx.node[(size_t)pos * 8 + ND_LO] = lo;
        cx.wk[WK_NPREDS] = lo + cnt;
        cx.wk[WK_FLAGS] &= ~1u;
    }
    cx.ct[CT_PRED] += cnt;
    __syncwarp();
}

// cascade_new_terminal / cascade_old_node (empty_transitions.rs:50-127).  The value carried upward is the start
// node's c* unchanged (:71-74,111-114), so a level's merge only has to sum the newly-exhausted counts per parent.
__device__ void tree_cascade(const AzbLayout &L, WarpCtx &cx, uint32_t src, float cstar, uint32_t ntt, uint32_t e0,
                             bool old) {
    const int lane = cx.lane;
    uint32_t *curN = cx.fr, *curE = cx.fr + AZB_FRONTIER_CAP;
    uint32_t *nxtN = cx.fr + 2 * AZB_FRONTIER_CAP, *nxtE = cx.fr + 3 * AZB_FRONTIER_CAP;
    if (lane == 0) {
        curN[0] = src;
        curE[0] = e0;
    }
    uint32_t ncur = 1;
    __syncwarp();
    uint4 *node4 = reinterpret_cast<uint4 *>(cx.node);
    while (ncur > 0 && cx.err == 0) {
        uint32_t nnxt = 0;
        cx.ct[CT_CN] += ncur;
        for (uint32_t base = 0; base < ncur; base += 32) {
            uint32_t i = base + lane;
            bool valid = i < ncur;
            uint32_t nin = 0, off = 0, up = 0;
            if (valid) {
                uint32_t p = curN[i], e = curE[i];
                uint4 r0 = node4[(size_t)p * 2], r1 = node4[(size_t)p * 2 + 1];
                uint32_t ex = (r0.w & 0xffffu) + e, cnt = r0.w >> 16;
                float cs = __uint_as_float(r0.y);
                uint32_t nt = r0.z;
                if (cs > cstar)
                    cs = cstar;
                else
                    nt += 1;
                if (old) nt = max(nt, ntt);
                r0.y = __float_as_uint(cs);
                r0.z = nt;
                r0.w = ex | (cnt << 16);
                node4[(size_t)p * 2] = r0;
                up = (ex < cnt) ? 0u : 1u;
                nin = r1.y >> 16;
                off = r1.z;
            }
            uint32_t nvalid = min(32u, ncur - base);
            for (uint32_t l = 0; l < nvalid; ++l) {
                uint32_t nin_l = __shfl_sync(0xffffffffu, nin, l);
                uint32_t off_l = __shfl_sync(0xffffffffu, off, l);
                uint32_t up_l = __shfl_sync(0xffffffffu, up, l);
                cx.ct[CT_DCN] += nin_l;
                for (uint32_t q0 = 0; q0 < nin_l; q0 += 32) {
                    bool has = q0 + lane < nin_l;
                    uint32_t q = has ? cx.inl[off_l + q0 + lane] : 0u;
                    int idx = -1;
                    if (has)
                        for (uint32_t k = 0; k < nnxt; ++k)
                            if (nxtN[k] == q) {
                                idx = (int)k;
                                break;
                            }
                    bool fresh = has && idx < 0;
                    uint32_t bal = __ballot_sync(0xffffffffu, fresh);
                    if (fresh) {
                        uint32_t at = nnxt + __popc(bal & lanemask_lt(lane));
                        if (at < AZB_FRONTIER_CAP) {
                            nxtN[at] = q;
                            nxtE[at] = up_l;
                        }
                    } else if (has) {
                        nxtE[idx] += up_l;
                    }
                    nnxt += __popc(bal);
                    if (nnxt > AZB_FRONTIER_CAP) {
                        cx.err = 3;
                        nnxt = AZB_FRONTIER_CAP;
                    }
                    __syncwarp();
                }
            }
        }
        uint32_t *t = curN;
        curN = nxtN;
        nxtN = t;
        t = curE;
        curE = nxtE;
        nxtE = t;
        ncur = nnxt;
        __syncwarp();
    }
}

// add_arc (graph_operations.rs:18-30): the new arc becomes kid[lo + n_out] of `src`, and `src` joins dst's parents
__device__ __forceinline__ void tree_add_arc(WarpCtx &cx, uint32_t src, const NodeRec &s, uint32_t j, uint32_t a,
                                             float g_bits_as_float, uint32_t dst, uint32_t dst_in_off,
                                             uint32_t dst_n_in, uint32_t dst_outin_word) {
    if (cx.lane == 0) {
        uint32_t slot = s.lo + s.n_out;
        cx.kid[slot] = make_uint2(dst, j | (a << 16));
        cx.arcseq[slot] = cx.wk[WK_NARCS];
        cx.wk[WK_NARCS] += 1;
        cx.pred[s.lo + j] = make_uint2(__float_as_uint(g_bits_as_float), a | (1u << 16));
        cx.node[(size_t)src * 8 + ND_OUTIN] = (s.n_out + 1) | (s.n_in << 16);
        cx.inl[dst_in_off + dst_n_in] = src;
        cx.node[(size_t)dst * 8 + ND_OUTIN] = (dst_outin_word & 0xffffu) | ((dst_n_in + 1) << 16);
    }
    cx.ct[CT_ARC] += 1;
    __syncwarp();
}

template <int DEPTH>
__device__ void tree_rollout(const AzbLayout &L, WarpCtx &cx, uint32_t tree, uint32_t best_c_start) {
    const int lane = cx.lane;
    uint32_t pos = cx.wk[WK_POS], depth = cx.wk[WK_DEPTH];
    uint4 *node4 = reinterpret_cast<uint4 *>(cx.node);
    uint32_t guard = 0;
    for (;;) {
        if (cx.err) break;
        if (++guard > 4u * L.cap_preds + 64u) {
            cx.err = 6;
            break;
        }
        NodeRec s = load_node(cx, pos);
        if (!(s.ex < s.cnt)) {  // next_action.rs:12-14
            if (depth != 0) cx.err = 6;  // tree/mod.rs:227 unreachable!()
            break;
        }
        cx.ct[CT_SEL] += 1;
        cx.ct[CT_DSEL] += s.n_out;
        // ---- revisit_choice (next_action.rs:28-53): first minimum of (n_t, c*) over active children, newest first
        unsigned long long best_key = ~0ull;
        int best_t = -1;
        uint32_t best_child = 0, best_aid = 0;
        for (uint32_t base = 0; base < s.n_out; base += 32) {
            uint32_t t = base + lane;
            bool valid = t < s.n_out;
            unsigned long long k = ~0ull;
            uint32_t child = 0, kd_y = 0;
            if (valid) {
                uint2 kd = cx.kid[s.lo + t];
                child = kd.x;
                kd_y = kd.y;
                uint4 r0 = node4[(size_t)child * 2];
                cx.lbuf[t] = __uint_as_float(r0.y);
                bool act = (r0.w & 0xffffu) < (r0.w >> 16);
                if (act) k = ((unsigned long long)r0.z << 32) | azb_f2ord(__uint_as_float(r0.y));
            }
            unsigned long long mk = warp_min_u64(k);
            if (mk != ~0ull && mk <= best_key) {  // a later chunk holds newer arcs: it wins ties
                uint32_t bal = __ballot_sync(0xffffffffu, k == mk);
                int wl = 31 - __clz(bal);  // newest arc among equals
                best_key = mk;
                best_t = (int)base + wl;
                best_child = __shfl_sync(0xffffffffu, child, wl);
                best_aid = __shfl_sync(0xffffffffu, kd_y, wl) >> 16;
            }
        }
        __syncwarp();
        bool have_r = best_t >= 0;
        uint32_t tol = depth < L.tol_len ? L.tol[depth] : L.tol_default;  // 04-c21-tree.rs:136-138
        bool visit = have_r && (uint32_t)(best_key >> 32) < tol;          // next_action.rs:16-20
        int chosen_j = -1;
        uint32_t chosen_a = 0;
        float chosen_g = 0.f;
        if (!visit) {
            // ---- max_curiosity (next_action.rs:55-88)
            cx.ct[CT_CUR] += 1;
            cx.ct[CT_CAND] += s.cnt;
            const bool no_kids = s.n_out == 0;
            unsigned long long bk = no_kids ? ~0ull : 0ull;
            for (uint32_t base = 0; base < s.cnt; base += 32) {
                uint32_t j = base + lane;
                bool valid = j < s.cnt;
                unsigned long long kk = no_kids ? ~0ull : 0ull;
                uint2 pr = make_uint2(0u, 0u);
                if (valid) {
                    pr = cx.pred[s.lo + j];
                    if (((pr.y >> 16) & 1u) == 0u) {  // no arc yet (:68-71)
                        float v = __fsub_rn(s.c, __uint_as_float(pr.x));
                        if (no_kids) {
                            if (v != v) cx.err = 4;
                            kk = ((unsigned long long)azb_f2ord(v) << 32) | j;  // first minimum (:73-75)
                        } else {
                            float cur = 0.f;
                            for (int t = (int)s.n_out - 1; t >= 0; --t)  // newest first, left fold (:79-82)
                                cur = __fadd_rn(cur, __fsqrt_rn(fabsf(__fsub_rn(cx.lbuf[t], v))));
                            if (cur != cur) cx.err = 4;
                            kk = ((unsigned long long)azb_f2ord(cur) << 32) | j;  // last maximum (:85)
                        }
                    }
                }
                unsigned long long red = no_kids ? warp_min_u64(kk) : warp_max_u64(kk);
                bool better = no_kids ? (red < bk) : (red > bk);
                if (better) {
                    bk = red;
                    uint32_t bal = __ballot_sync(0xffffffffu, kk == red);
                    int wl = __ffs(bal) - 1;
                    chosen_j = (int)(red & 0xffffffffu);
                    chosen_a = __shfl_sync(0xffffffffu, pr.y, wl) & 0xffffu;
                    chosen_g = __uint_as_float(__shfl_sync(0xffffffffu, pr.x, wl));
                }
            }
            cx.err = __reduce_or_sync(0xffffffffu, cx.err);
            if (cx.err) break;
            if (chosen_j < 0) {
                if (have_r)
                    visit = true;  // next_action.rs:24
                else {
                    cx.err = 6;
                    break;
                }
            }
        }
        if (visit) {  // tree/mod.rs:139-159
            cx.ct[CT_VISIT] += 1;
            walker_act(L, cx, best_aid);
            pos = best_child;
            depth += 1;
            continue;
        }
        // ---- Unvisited(j) (tree/mod.rs:160-218)
        const uint32_t a = chosen_a, j = (uint32_t)chosen_j;
        cx.ct[CT_PROBE] += 1;
        // key of the successor = path + a; hash it
        uint32_t k0 = (uint32_t)lane < L.W ? (cx.keym[lane] | (((a >> 5) == (uint32_t)lane) ? 1u << (a & 31) : 0u)) : 0u;
        uint32_t k1 = (uint32_t)lane + 32 < L.W
                          ? (cx.keym[lane + 32] | (((a >> 5) == (uint32_t)lane + 32) ? 1u << (a & 31) : 0u))
                          : 0u;
        uint32_t hv = key_hash(k0, k1, lane, L.W);
        uint32_t slot = hv & (L.cap_hash - 1);
        int hit = -1;
        for (uint32_t probes = 0; probes < L.cap_hash; ++probes) {
            uint32_t e = cx.hash[slot];
            if (e == 0u) break;
            uint32_t idx = e - 1;
            bool eq = true;
            if ((uint32_t)lane < L.W) eq = cx.key[(size_t)idx * L.W + lane] == k0;
            if ((uint32_t)lane + 32 < L.W) eq = eq && (cx.key[(size_t)idx * L.W + lane + 32] == k1);
            if (__all_sync(0xffffffffu, eq)) {
                hit = (int)idx;
                break;
            }
            slot = (slot + 1) & (L.cap_hash - 1);
        }
        if (hit >= 0) {  // transposition (tree/mod.rs:172-179)
            cx.ct[CT_HIT] += 1;
            NodeRec o = load_node(cx, (uint32_t)hit);
            if (o.n_in >= o.depth) {
                cx.err = 3;
                break;
            }
            tree_add_arc(cx, pos, s, j, a, chosen_g, (uint32_t)hit, o.in_off, o.n_in, o.n_out);
            tree_cascade(L, cx, pos, o.cstar, o.nt, (o.ex < o.cnt) ? 0u : 1u, true);
            walker_reset(L, cx);
            pos = 0;
            depth = 0;
            cx.ct[CT_RESET] += 1;
            continue;
        }
        // ---- new node (tree/mod.rs:181-216)
        walker_act(L, cx, a);
        const uint32_t ndepth = depth + 1;
        double l1 = azb_lambda1_warp<DEPTH>(L.N, cx.par, cx.cs, lane);
        uint32_t mu = azb_matching(L.N, cx.par);
        if (!(l1 >= 1.4)) {  // ordered_edge.rs:79
            cx.err = 5;
            break;
        }
        float c_new = azb_evaluate(mu, l1, L.c_lower, L.slope);
        cx.ct[CT_INS] += 1;
        const uint32_t nn = cx.wk[WK_NNODES], in_off = cx.wk[WK_INTOP];
        if (nn >= L.cap_nodes || in_off + ndepth > L.cap_in) {
            cx.err = 3;
            break;
        }
        __syncwarp();
        if (lane < 8) {
            uint32_t w = 0u;
            if (lane == ND_C || lane == ND_CSTAR) w = __float_as_uint(c_new);  // StateWeight::new (state_weight.rs:13-21)
            if (lane == ND_OUTIN) w = 0u;  // n_out = 0, n_in = 0 (add_arc below makes it 1)
            if (lane == ND_INOFF) w = in_off;
            if (lane == ND_DEPTH) w = ndepth;
            cx.node[(size_t)nn * 8 + lane] = w;
        }
        if ((uint32_t)lane < L.W) cx.key[(size_t)nn * L.W + lane] = k0;
        if ((uint32_t)lane + 32 < L.W) cx.key[(size_t)nn * L.W + lane + 32] = k1;
        if (lane == 0) {
            cx.hash[slot] = nn + 1;
            cx.wk[WK_NNODES] = nn + 1;
            cx.wk[WK_INTOP] = in_off + ndepth;
        }
        __syncwarp();
        tree_add_arc(cx, pos, s, j, a, chosen_g, nn, in_off, 0u, 0u);
        // argmin candidate: first minimum of c over this step's new nodes that beats the best (optimizer/mod.rs:208-213)
        uint32_t oc = azb_f2ord(c_new);
        if (oc < best_c_start && oc < cx.wk[WK_CAND_C]) {
            if (lane == 0) {
                cx.wk[WK_CAND_C] = oc;
                cx.wk[WK_CAND_NODE] = nn;
            }
            __syncwarp();
        }
        // is_terminal (nabla/space/mod.rs:27-29)
        build_cur_mask(L, cx);
        bool any = false;
        for (uint32_t w = lane; w < L.W; w += 32) any = any || ((cx.perm[w] & ~cx.cur[w]) != 0u);
        any = __any_sync(0xffffffffu, any);
        if (!any) {
            cx.ct[CT_TERM] += 1;
            tree_cascade(L, cx, pos, c_new, 0u, 1u, false);
            walker_reset(L, cx);
            pos = 0;
            depth = 0;
            cx.ct[CT_RESET] += 1;
            continue;
        }
        pos = nn;
        depth = ndepth;
        if (lane == 0) {
            cx.wk[WK_FLAGS] |= 1u;
            cx.wk[10] = __float_as_uint(c_new);
        }
        __syncwarp();
        break;  // tree/mod.rs:212-215
    }
    if (lane == 0) {
        cx.wk[WK_POS] = pos;
        cx.wk[WK_DEPTH] = depth;
    }
    __syncwarp();
}

// write_vec (rooted_tree/space.rs:91-101): [A one-hot of current edges | A permitted mask] as f32
__device__ __forceinline__ void tree_pack(const AzbLayout &L, WarpCtx &cx, uint32_t tree) {
    build_cur_mask(L, cx);
    float *row = L.sv + (size_t)tree * L.sv_ld;
    for (uint32_t i = cx.lane; i < 2 * L.A; i += 32) {
        uint32_t b = i < L.A ? (cx.cur[i >> 5] >> (i & 31)) : (cx.perm[(i - L.A) >> 5] >> ((i - L.A) & 31));
        row[i] = (b & 1u) ? 1.0f : 0.0f;
    }
}

// replay a node's action set on its tree's root (optimizer/mod.rs:224-239) into g->argmin_state; one warp
__device__ void finalize_argmin_state(const AzbLayout &L, uint32_t *scratch, uint32_t tree, uint32_t node, int lane) {
    WarpCtx cx;
    cx.lane = lane;
    cx.wk = scratch;
    cx.par = (uint8_t *)(scratch + WK_HDR);
    cx.perm = scratch + WK_HDR + L.PW;
    cx.keym = cx.perm + L.W;
    const uint32_t *src = L.walker + (size_t)tree * L.WS;
    for (uint32_t i = lane; i < L.PW; i += 32) ((uint32_t *)cx.par)[i] = src[WK_HDR + L.PW + 2 * L.W + i];
    for (uint32_t w = lane; w < L.W; w += 32) {
        cx.perm[w] = src[WK_HDR + 2 * L.PW + 2 * L.W + w];
        cx.keym[w] = 0u;
    }
    __syncwarp();
    const uint32_t *key = L.key + ((size_t)tree * L.cap_nodes + node) * L.W;
    for (uint32_t w = 0; w < L.W; ++w) {
        uint32_t word = key[w];
        while (word) {
            uint32_t b = __ffs(word) - 1;
            word &= word - 1;
            walker_act(L, cx, w * 32 + b);
        }
    }
    for (uint32_t i = lane; i < 16; i += 32) L.g->argmin_state[i] = i < L.PW ? ((uint32_t *)cx.par)[i] : 0u;
    for (uint32_t w = lane; w < 61; w += 32) L.g->argmin_state[16 + w] = w < L.W ? cx.perm[w] : 0u;
}

template <int DEPTH>
__global__ void __launch_bounds__(AZB_WARPS_PER_BLOCK * 32)
    azb_tree_kernel(const AzbLayout L, const uint32_t flags, const uint32_t smem_words_per_warp, const uint32_t lcap) {
    extern __shared__ __align__(16) uint32_t smem[];
    __shared__ uint32_t s_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tree = blockIdx.x * AZB_WARPS_PER_BLOCK + warp;
    uint32_t *base = smem + (size_t)warp * smem_words_per_warp;
    const uint32_t best_c_start = L.g->best_c;
    const uint32_t prior_step = (flags & AZB_F_INIT) ? 0u : L.g->step;  // priors of step t feed the add_actions after rollout t-1
    if (tree < L.B) {
        WarpCtx cx;
        cx.lane = lane;
        cx.err = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) cx.ct[i] = 0;
        cx.wk = base;
        cx.par = (uint8_t *)(base + WK_HDR);
        cx.perm = base + WK_HDR + L.PW;
        cx.keym = cx.perm + L.W;
        cx.rpar = (uint8_t *)(cx.keym + L.W);
        cx.rperm = cx.keym + L.W + L.PW;
        uint32_t *p = base + L.WS;
        cx.cur = p;
        p += 64;
        cx.pfx = p;
        p += 64;
        cx.lbuf = (float *)p;
        p += lcap;
        cx.fr = p;
        p += 4 * AZB_FRONTIER_CAP;
        cx.cs = reinterpret_cast<CostScratch *>(p);
        cx.node = reinterpret_cast<uint32_t *>(L.node) + (size_t)tree * L.cap_nodes * 8;
        cx.pred = L.pred + (size_t)tree * L.cap_preds;
        cx.kid = L.kid + (size_t)tree * L.cap_preds;
        cx.arcseq = L.arcseq + (size_t)tree * L.cap_preds;
        cx.inl = L.inl + (size_t)tree * L.cap_in;
        cx.key = L.key + (size_t)tree * L.cap_nodes * L.W;
        cx.hash = L.hash + (size_t)tree * L.cap_hash;
        uint32_t *gw = L.walker + (size_t)tree * L.WS;
        for (uint32_t i = lane; i < L.WS; i += 32) cx.wk[i] = gw[i];
        __syncwarp();
        if (lane == 0) cx.wk[WK_CAND_C] = 0xffffffffu;
        __syncwarp();

        if (flags & AZB_F_INIT) {
            // tail of par_new / par_reset_trees (optimizer/mod.rs:62-101, 340-359): state <- root, root cost, root node
            walker_reset(L, cx);
            double l1 = azb_lambda1_warp<DEPTH>(L.N, cx.par, cx.cs, lane);
            uint32_t mu = azb_matching(L.N, cx.par);
            if (!(l1 >= 1.4)) cx.err = 5;
            float c0 = azb_evaluate(mu, l1, L.c_lower, L.slope);
            if (lane < 8) {
                uint32_t w = 0u;
                if (lane == ND_C || lane == ND_CSTAR) w = __float_as_uint(c0);
                cx.node[lane] = w;
            }
            if ((uint32_t)lane < L.W) cx.key[lane] = 0u;
            if ((uint32_t)lane + 32 < L.W) cx.key[lane + 32] = 0u;
            // empty path hashes like every other key
            uint32_t hv = key_hash(0u, 0u, lane, L.W);
            if (lane == 0) {
                cx.hash[hv & (L.cap_hash - 1)] = 1u;
                cx.wk[WK_POS] = 0;
                cx.wk[WK_DEPTH] = 0;
                cx.wk[WK_NNODES] = 1;
                cx.wk[WK_NPREDS] = 0;
                cx.wk[WK_NARCS] = 0;
                cx.wk[WK_INTOP] = 0;
                cx.wk[10] = __float_as_uint(c0);
                if (flags & AZB_F_FIRST) {
                    // par_new: silent argmin over the roots (optimizer/mod.rs:95-101)
                    cx.wk[WK_FLAGS] = 1u;
                    cx.wk[WK_CAND_C] = azb_f2ord(c0);
                    cx.wk[WK_CAND_NODE] = 0;
                } else {
                    // par_reset_trees zeroes num_inspected_nodes (optimizer/mod.rs:359): the new root is looked at by
                    // the first step's argmin scan
                    cx.wk[WK_FLAGS] = 3u;
                }
            }
            __syncwarp();
            tree_pack(L, cx, tree);
        }
        if ((flags & AZB_F_ADD) && (cx.wk[WK_FLAGS] & 1u)) tree_add_actions(L, cx, tree, prior_step);
        if (flags & AZB_F_ROLLOUT) {
            if (cx.wk[WK_FLAGS] & 2u) {  // root not yet inspected by par_update_argmmim_data (optimizer/mod.rs:203-219)
                uint32_t oc = azb_f2ord(__uint_as_float(cx.node[ND_C]));
                __syncwarp();
                if (lane == 0) {
                    cx.wk[WK_FLAGS] &= ~2u;
                    if (oc < best_c_start) {
                        cx.wk[WK_CAND_C] = oc;
                        cx.wk[WK_CAND_NODE] = 0;
                    }
                }
                __syncwarp();
            }
            tree_rollout<DEPTH>(L, cx, tree, best_c_start);
            if (cx.err == 0) {
                if (cx.wk[WK_DEPTH] != 0) {  // optimizer/mod.rs:171-173
                    cx.ct[CT_LIVE] += 1;
                    tree_pack(L, cx, tree);
                } else {
                    cx.ct[CT_NOOP] += 1;
                }
            }
        }
        __syncwarp();
        // publish: walker block (without the root part), the step's argmin candidate, counters, errors
        const uint32_t live_words = WK_HDR + L.PW + 2 * L.W;
        for (uint32_t i = lane; i < live_words; i += 32) gw[i] = cx.wk[i];
        if (lane == 0 && cx.wk[WK_CAND_C] != 0xffffffffu)
            atomicMin(&L.g->step_best, ((unsigned long long)cx.wk[WK_CAND_C] << 32) | tree);
        if (lane < 16) {
            uint32_t v = 0;
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (lane == i) v = cx.ct[i];
            if (v) atomicAdd(&L.g->counters.v[lane], (unsigned long long)v);
        }
        uint32_t e = __reduce_or_sync(0xffffffffu, cx.err);
        if (e && lane == 0) {
            if (atomicCAS(&L.g->err, 0u, e) == 0u) L.g->err_tree = tree;
        }
    }
    // ---- last block done: par_update_argmmim_data's cross-tree minimum (optimizer/mod.rs:221-245)
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&L.g->blocks_done, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (s_last && warp == 0) {
        __threadfence();
        AzbGlobals *g = L.g;
        unsigned long long sb = *((volatile unsigned long long *)&g->step_best);
        uint32_t improved = 0;
        if (sb != ~0ull) {
            uint32_t oc = (uint32_t)(sb >> 32), t = (uint32_t)sb;
            if (oc < g->best_c) {
                uint32_t node = *((volatile uint32_t *)&L.walker[(size_t)t * L.WS + WK_CAND_NODE]);
                improved = (flags & AZB_F_INIT) ? 0u : 1u;
                if (lane == 0) {
                    g->best_c = oc;
                    if (improved) {
                        uint32_t n = g->n_improved;
                        if (n < L.log_cap) {
                            L.log[n].step = g->step;
                            L.log[n].tree = t;
                            L.log[n].node = node;
                            L.log[n].eval = azb_ord2f(oc);
                        }
                        g->n_improved = n + 1;
                    }
                }
                finalize_argmin_state(L, base, t, node, lane);
            }
        }
        __syncwarp();
        if (lane == 0) {
            g->step_best = ~0ull;
            g->improved_last = improved;
            if (flags & AZB_F_ROLLOUT) g->step += 1;
            g->blocks_done = 0u;
            __threadfence();
        }
    }
}

// write_observations (tree/mod.rs:242-264) for every tree, plus the root vectors par_update_model packs first
// (optimizer/mod.rs:253-259).  One warp per tree; h_sa = c*_as (04-c21-tree.rs:104).
__global__ void __launch_bounds__(128) azb_observe_kernel(const AzbLayout L, const uint32_t n_obs_tol,
                                                          float *__restrict__ obs, float *__restrict__ wts,
                                                          float *__restrict__ root_vecs) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tree = blockIdx.x * 4 + warp;
    if (tree >= L.B) return;
    float *o = obs + (size_t)tree * L.A, *w = wts + (size_t)tree * L.A;
    for (uint32_t a = lane; a < L.A; a += 32) {
        o[a] = 0.f;
        w[a] = 0.f;
    }
    __syncwarp();
    const uint4 *node4 = L.node + (size_t)tree * L.cap_nodes * 2;
    const uint2 *kid = L.kid + (size_t)tree * L.cap_preds;
    const uint4 r0 = node4[0], r1 = node4[1];
    const uint32_t lo = r1.x, n_out = r1.y & 0xffffu;
    for (uint32_t t = lane; t < n_out; t += 32) {
        const uint2 kd = kid[lo + t];
        const uint4 c0 = node4[(size_t)kd.x * 2];
        const bool active = (c0.w & 0xffffu) < (c0.w >> 16);
        if (!active || c0.z >= n_obs_tol) {
            const uint32_t a = kd.y >> 16;
            o[a] = __uint_as_float(c0.y);
            w[a] = 1.0f;
        }
    }
    (void)r0;
    if (root_vecs) {
        const uint32_t *wk = L.walker + (size_t)tree * L.WS;
        const uint8_t *rpar = (const uint8_t *)(wk + WK_HDR + L.PW + 2 * L.W);
        const uint32_t *rperm = wk + WK_HDR + 2 * L.PW + 2 * L.W;
        float *row = root_vecs + (size_t)tree * L.sv_ld;
        for (uint32_t i = lane; i < 2 * L.A; i += 32) {
            bool one;
            if (i < L.A) {
                const uint32_t child = azb_action_child(i);
                one = (uint32_t)rpar[child] == i - azb_child_first_action(child);
            } else {
                one = (rperm[(i - L.A) >> 5] >> ((i - L.A) & 31)) & 1u;
            }
            row[i] = one ? 1.0f : 0.0f;
        }
    }
}
