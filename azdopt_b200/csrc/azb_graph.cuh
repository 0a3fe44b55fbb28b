// azb_graph.cuh -- SURVEY 8(f) row 3: the c21 cost and the action kinds of CONNECTED GRAPHS held as neighbourhood
// bit sets (`ConnectedBitsetGraph<N, B32>`, graph-state/src/simple_graph/connected_bitset_graph/mod.rs), batched.
// One warp per graph, lane v = vertex v (N <= 32), everything in shared memory and registers:
//   * action kinds (`action_kinds` :134-154 over `is_cut_edge` :45-71) in the index space of
//     `AddOrDeleteEdge::action_index` (bitset_graph/space/action.rs:10-19): the bit-set BFS test on the N - 1 edges
//     of a spanning tree (the only possible cut edges), then one lane per vertex pair;
//   * matching number (`maximum_matching` :226-317 is a branch and bound; only its SIZE is read, so any exact
//     algorithm gives the same number): greedy start + Edmonds' blossom augmentation, serial on lane 0;
//   * lambda_1 (`adjacency_matrix` :200-216 with its 1e-4 diagonal, `conjecture_2_1_cost` :319-337 takes the
//     largest eigenvalue): warp-parallel Householder tridiagonalisation in f64 (lane = row), then the largest
//     eigenvalue by Sturm counts on a 32-point section per round (lane = section point).
// The snapshot has no NablaStateActionSpace over these graphs, so this is the stand-alone cost row only.
#pragma once
#include <cstdint>

#define AZG_WARPS 4

struct AzgWarpScratch {  // one per warp, carved out of dynamic shared memory: the matrix is a packed lower triangle
    double *a;
    double *u, *w, *d, *e;
    uint32_t *nbr, *kinds;
    int8_t *match, *par, *base, *queue;
};
__host__ __device__ inline uint32_t azg_tri(uint32_t n) { return (n * (n + 1) / 2 + 1u) & ~1u; }  // doubles, 16 B multiple
__host__ __device__ inline uint32_t azg_ld(uint32_t n) { return n | 1u; }  // square layout: odd row stride
__host__ __device__ inline bool azg_packed(uint32_t n) { return n > 22u; }  // the triangle costs index arithmetic and a
// divergent second loop in A u; it pays once the square would hold the kernel below ~40 resident warps per SM
__host__ __device__ inline uint32_t azg_mat_doubles(uint32_t n) { return azg_packed(n) ? azg_tri(n) : n * azg_ld(n); }
__host__ __device__ inline uint32_t azg_warp_bytes(uint32_t n) { return 8u * azg_mat_doubles(n) + 4u * 32u * 8u + 2u * 32u * 4u + 4u * 32u; }

__device__ __forceinline__ double azg_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double azg_warp_max(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// connected_bitset_graph/mod.rs:45-71
__device__ __forceinline__ bool azg_is_cut_edge(const uint32_t *nbr, uint32_t v, uint32_t u) {
    uint32_t fresh = nbr[v] ^ (1u << u), explored = 1u << v;
    while (fresh) {
        if (fresh >> u & 1u) return false;
        explored |= fresh;
        uint32_t next = 0u;
        for (uint32_t r = fresh; r; r &= r - 1) next |= nbr[__ffs(r) - 1];
        fresh = next & ~explored;
    }
    return true;
}

// Edmonds' blossom algorithm (array form), serial, on the vertices of `alive` with neighbourhoods nbr[] (already
// restricted to `alive`); match/par/base/queue live in the warp's shared scratch
__device__ uint32_t azg_matching_number(const AzgWarpScratch &s, const uint32_t *nbr, uint32_t alive, uint32_t n) {
    int8_t *match = s.match, *par = s.par, *base = s.base, *queue = s.queue;
    uint32_t size = 0;
    for (uint32_t v = 0; v < n; ++v) match[v] = -1;
    for (uint32_t av = alive; av; av &= av - 1) {  // greedy start
        const uint32_t v = __ffs(av) - 1;
        if (match[v] >= 0) continue;
        for (uint32_t r = nbr[v]; r; r &= r - 1) {
            const int t = __ffs(r) - 1;
            if (match[t] < 0) {
                match[v] = (int8_t)t;
                match[t] = (int8_t)v;
                ++size;
                break;
            }
        }
    }
    for (uint32_t ar = alive; ar; ar &= ar - 1) {
        const uint32_t root = __ffs(ar) - 1;
        if (match[root] >= 0) continue;
        for (uint32_t i = 0; i < n; ++i) {
            par[i] = -1;
            base[i] = (int8_t)i;
        }
        uint32_t used = 1u << root, qh = 0, qt = 0;
        queue[qt++] = (int8_t)root;
        int found = -1;
        while (qh < qt && found < 0) {
            const int v = queue[qh++];
            for (uint32_t r = nbr[v]; r && found < 0; r &= r - 1) {
                int to = __ffs(r) - 1;
                if (base[v] == base[to] || match[v] == to) continue;
                if (to == (int)root || (match[to] >= 0 && par[match[to]] >= 0)) {
                    // an odd cycle: contract it onto the lowest common ancestor of v and to
                    uint32_t seen = 0u;
                    int a = v, b = to, cur;
                    for (;;) {
                        a = base[a];
                        seen |= 1u << a;
                        if (match[a] < 0) break;
                        a = par[match[a]];
                    }
                    for (;;) {
                        b = base[b];
                        if (seen >> b & 1u) break;
                        b = par[match[b]];
                    }
                    cur = b;
                    uint32_t blossom = 0u;
                    for (int side = 0; side < 2; ++side) {
                        int x = side ? to : v, child = side ? v : to;
                        while (base[x] != cur) {
                            blossom |= 1u << base[x];
                            blossom |= 1u << base[match[x]];
                            par[x] = (int8_t)child;
                            child = match[x];
                            x = par[match[x]];
                        }
                    }
                    for (uint32_t i = 0; i < n; ++i)
                        if (blossom >> base[i] & 1u) {
                            base[i] = (int8_t)cur;
                            if (!(used >> i & 1u)) {
                                used |= 1u << i;
                                queue[qt++] = (int8_t)i;
                            }
                        }
                } else if (par[to] < 0) {
                    par[to] = (int8_t)v;
                    if (match[to] < 0) {
                        found = to;
                    } else {
                        to = match[to];
                        used |= 1u << to;
                        queue[qt++] = (int8_t)to;
                    }
                }
            }
        }
        if (found >= 0) {
            ++size;
            for (int v = found; v >= 0;) {
                const int pv = par[v], ppv = match[pv];
                match[v] = (int8_t)pv;
                match[pv] = (int8_t)v;
                v = ppv;
            }
        }
    }
    return size;
}

// x > lambda_max of the symmetric tridiagonal (d, e)  <=>  every leading principal minor of x I - T is positive:
// p_0 = 1, p_1 = x - d_0, p_i = (x - d_{i-1}) p_{i-1} - e_{i-2}^2 p_{i-2}.  No division; |p_i| <= (2 ||T||)^32 < 1e60.
__device__ __forceinline__ bool azg_above_all(const double *d, const double *e2, uint32_t n, double x) {
    double pm = 1.0, p = __dsub_rn(x, d[0]);
    bool all = p > 0.0;
    // no early exit: the warp runs as long as its highest section point anyway, and that one never leaves early
    for (uint32_t i = 1; i < n; ++i) {
        const double pn = __fma_rn(__dsub_rn(x, d[i]), p, -__dmul_rn(e2[i - 1], pm));
        pm = p;
        p = pn;
        all = all && p > 0.0;
    }
    return all;
}

template <bool PACKED>
__global__ void __launch_bounds__(AZG_WARPS * 32)
azb_graph_cost_kernel(const uint32_t *__restrict__ nbr_g, uint32_t m, uint32_t n, uint32_t kw, double *__restrict__ l1_out,
                      uint32_t *__restrict__ mu_out, uint32_t *__restrict__ kinds_out, uint32_t *err) {
    extern __shared__ __align__(16) uint8_t azg_smem[];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t g = blockIdx.x * AZG_WARPS + warp;
    if (g >= m) return;
    AzgWarpScratch s;
    {
        uint8_t *base = azg_smem + (size_t)warp * azg_warp_bytes(n);
        s.a = reinterpret_cast<double *>(base);
        s.u = s.a + azg_mat_doubles(n);
        s.w = s.u + 32;
        s.d = s.w + 32;
        s.e = s.d + 32;
        s.nbr = reinterpret_cast<uint32_t *>(s.e + 32);
        s.kinds = s.nbr + 32;
        s.match = reinterpret_cast<int8_t *>(s.kinds + 32);
        s.par = s.match + 32;
        s.base = s.par + 32;
        s.queue = s.base + 32;
    }
    const uint32_t mine = lane < n ? nbr_g[(size_t)g * n + lane] : 0u;
    s.nbr[lane] = mine;
    s.kinds[lane] = 0u;
    __syncwarp();
    // ---- what BitsetGraph::try_from / to_connected accept: no loops, neighbours < N, symmetric (column v of the
    // relation, gathered by ballots, must equal row v); connectedness falls out of the spanning tree below
    const uint32_t full = n == 32 ? 0xffffffffu : (1u << n) - 1u;
    bool bad = (mine & ~full) || (mine >> lane & 1u);
    for (uint32_t v = 0; v < n; ++v) {
        const uint32_t column = __ballot_sync(0xffffffffu, mine >> v & 1u);
        if (lane == v) bad = bad || column != mine;
    }

    // ---- cut edges.  Only the N - 1 edges of a spanning tree can be cut edges (any other edge closes a cycle with
    // the tree), so: a breadth-first spanning tree by ballots (lane v learns its parent), then ONE round of the
    // reference's test (`is_cut_edge`), lane v on the edge to its parent, instead of one test per edge of the graph.
    uint32_t *brg = reinterpret_cast<uint32_t *>(s.u);  // brg[v]: bit u <=> vu is a cut edge (the doubles are free here)
    brg[lane] = 0u;
    int parent = -1;
    uint32_t seen = 1u;
    for (uint32_t frontier = 1u; frontier;) {
        const uint32_t cand = mine & frontier;
        const bool join = lane < n && !(seen >> lane & 1u) && cand;
        if (join) parent = __ffs(cand) - 1;
        frontier = __ballot_sync(0xffffffffu, join);
        seen |= frontier;
    }
    if (__any_sync(0xffffffffu, bad) || seen != full) {  // not a connected simple graph: report the first such input
        if (lane == 0) {
            atomicOr(err, 1u << 1);
            atomicMin(err + 1, g);
            l1_out[g] = 0.0;
            mu_out[g] = 0u;
        }
        return;
    }
    __syncwarp();
    if (parent >= 0 && azg_is_cut_edge(s.nbr, lane, (uint32_t)parent)) {
        atomicOr(&brg[lane], 1u << parent);
        atomicOr(&brg[parent], 1u << lane);
    }
    __syncwarp();

    // ---- action kinds: pair p = colex(v, u) = v (v - 1) / 2 + u, 32 pairs per round
    const uint32_t e2 = n * (n - 1) / 2;
    for (uint32_t p0 = 0; p0 < e2; p0 += 32) {
        const uint32_t p = p0 + lane;
        bool add = false, del = false;
        if (p < e2) {
            uint32_t v = (uint32_t)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
            while (v * (v - 1) / 2 > p) --v;
            while (v * (v + 1) / 2 <= p) ++v;
            const uint32_t u = p - v * (v - 1) / 2;
            if (s.nbr[v] >> u & 1u) del = !(brg[v] >> u & 1u);
            else add = true;
        }
        const uint32_t wa = __ballot_sync(0xffffffffu, add), wd = __ballot_sync(0xffffffffu, del);
        if (lane == 0) {
            s.kinds[p0 >> 5] |= wa;
            const uint32_t bit = e2 + p0, sh = bit & 31u;
            s.kinds[bit >> 5] |= wd << sh;
            if (sh && (wd >> (32u - sh))) s.kinds[(bit >> 5) + 1] |= wd >> (32u - sh);
        }
        __syncwarp();
    }
    if (kinds_out && lane < kw) kinds_out[(size_t)g * kw + lane] = s.kinds[lane];

    // ---- matching number: pendant edges first (some maximum matching contains any given pendant edge, so matching a
    // leaf with its neighbour and deleting both is exact), one ballot per leaf; what is left has minimum degree 2 and
    // goes through the blossom search on lane 0.  A tree never gets there.
    uint32_t mu = 0, alive = n == 32 ? 0xffffffffu : (1u << n) - 1u;
    for (;;) {
        const bool me = alive >> lane & 1u;
        const uint32_t mynb = mine & alive;
        const uint32_t leaves = __ballot_sync(0xffffffffu, me && __popc(mynb) == 1);
        alive &= ~__ballot_sync(0xffffffffu, me && mynb == 0u);
        if (!leaves) break;
        const int v = __ffs(leaves) - 1;
        const int t = __ffs(__shfl_sync(0xffffffffu, mynb, v)) - 1;
        alive &= ~((1u << v) | (1u << t));
        ++mu;
    }
    if (alive) {
        __syncwarp();
        s.kinds[lane] = (alive >> lane & 1u) ? (mine & alive) : 0u;  // the kinds are out: their words hold the core now
        __syncwarp();
        uint32_t core = 0;
        if (lane == 0) core = azg_matching_number(s, s.kinds, alive, n);
        mu += __shfl_sync(0xffffffffu, core, 0);
    }
    __syncwarp();

    // ---- lambda_1 of A + 1e-4 I.  The matrix is symmetric: lane i keeps the lower-triangle row i (entries j <= i) at
    // a + i (i + 1) / 2; the upper part of a row is read down its column (consecutive words across lanes).
    // (n <= 22: plain square rows of n | 1 doubles, every row whole -- fewer instructions, and it fits anyway)
    double *__restrict__ arow = s.a + (PACKED ? lane * (lane + 1) / 2 : lane * azg_ld(n));
    const uint32_t jmax = PACKED ? lane : n - 1;  // last column this lane stores
    if (lane < n)
        for (uint32_t j = 0; j <= jmax; ++j) arow[j] = lane == j ? 0.0001 : ((mine >> j & 1u) ? 1.0 : 0.0);
    __syncwarp();
    for (uint32_t k = 0; k + 2 < n; ++k) {
        const bool act = lane > k && lane < n;
        const double x = act ? arow[k] : 0.0;
        const double x0 = __shfl_sync(0xffffffffu, x, k + 1);
        const double rest2 = azg_warp_sum(lane == k + 1 ? 0.0 : __dmul_rn(x, x));
        const double norm2 = __fma_rn(x0, x0, rest2);
        if (rest2 == 0.0) {  // the column is already tridiagonal
            if (lane == 0) s.e[k] = x0;
            continue;
        }
        const double alpha = x0 > 0.0 ? -__dsqrt_rn(norm2) : __dsqrt_rn(norm2);
        double u = lane == k + 1 ? __dsub_rn(x, alpha) : x;
        const double u0 = __dsub_rn(x0, alpha);
        const double un2 = __fma_rn(u0, u0, rest2);
        u = __dmul_rn(u, __ddiv_rn(1.0, __dsqrt_rn(un2)));
        s.u[lane] = u;
        __syncwarp();
        const double *__restrict__ uu = s.u;
        double q = 0.0, q1 = 0.0;
        if (act) {
            uint32_t j = k + 1;
            for (; j + 1 <= jmax; j += 2) {  // own row (up to the diagonal when packed), two partial sums
                q = __fma_rn(arow[j], uu[j], q);
                q1 = __fma_rn(arow[j + 1], uu[j + 1], q1);
            }
            if (j <= jmax) q = __fma_rn(arow[j], uu[j], q);
            if (PACKED) {
                const double *col = s.a + lane;  // A[i][j] = A[j][i] for j > i: row j starts at j (j + 1) / 2
                j = lane + 1;
                uint32_t off = j * (j + 1) / 2;
                for (; j + 1 < n; j += 2) {
                    q = __fma_rn(col[off], uu[j], q);
                    off += j + 1;
                    q1 = __fma_rn(col[off], uu[j + 1], q1);
                    off += j + 2;
                }
                if (j < n) q = __fma_rn(col[off], uu[j], q);
            }
            q = __dadd_rn(q, q1);
        }
        const double uq = azg_warp_sum(__dmul_rn(u, q));
        const double w = __fma_rn(-uq, u, q);
        s.w[lane] = w;
        __syncwarp();  // also: every column read of this step is done before the rows change
        const double *__restrict__ ww = s.w;
        if (act) {
            const double u2 = __dmul_rn(-2.0, u), w2 = __dmul_rn(-2.0, w);  // A' = A - 2 u w^T - 2 w u^T
#pragma unroll 4
            for (uint32_t j = k + 1; j <= jmax; ++j) arow[j] = __fma_rn(u2, ww[j], __fma_rn(w2, uu[j], arow[j]));
        }
        if (lane == 0) s.e[k] = alpha;
        __syncwarp();
    }
    if (lane < n) s.d[lane] = arow[lane];
    if (lane == n - 1 && n >= 2) s.e[n - 2] = arow[n - 2];
    __syncwarp();
    if (lane + 1 < n) s.u[lane] = __dmul_rn(s.e[lane], s.e[lane]);  // squared off-diagonal, once
    __syncwarp();
    // Gershgorin bracket of the spectrum, then 33-fold sections of [lo, hi] with  not above(lo), above(hi)
    double rad = 0.0, dd = -1e300;
    if (lane < n) {
        dd = s.d[lane];
        rad = (lane > 0 ? fabs(s.e[lane - 1]) : 0.0) + (lane + 1 < n ? fabs(s.e[lane]) : 0.0);
    }
    double hi = azg_warp_max(lane < n ? dd + rad : -1e300), lo = -azg_warp_max(lane < n ? rad - dd : -1e300);
    const double scale = fmax(fabs(hi), fabs(lo));
    hi = hi + fmax(scale, 1.0) * 1e-9;
    for (int round = 0; round < 16; ++round) {
        const double step = __ddiv_rn(__dsub_rn(hi, lo), 33.0);
        const double x = __dadd_rn(lo, __dmul_rn(step, (double)(lane + 1)));
        const uint32_t above = __ballot_sync(0xffffffffu, azg_above_all(s.d, s.u, n, x));
        const int first = above ? __ffs(above) - 1 : 32;
        const double nlo = first > 0 ? __shfl_sync(0xffffffffu, x, first - 1) : lo;
        const double nhi = first < 32 ? __shfl_sync(0xffffffffu, x, first & 31) : hi;
        lo = nlo;
        hi = nhi;
        if (!(__dsub_rn(hi, lo) > 4.0 * 2.220446049250313e-16 * fmax(fabs(hi), fabs(lo)))) break;
    }
    const double l1 = __dmul_rn(0.5, __dadd_rn(lo, hi));
    if (lane == 0) {
        l1_out[g] = l1;
        mu_out[g] = mu;
        if (!(l1 > 1.4)) atomicOr(err, 1u << 5);  // connected_bitset_graph/mod.rs:333 `assert!(lambda_1 > 1.4)`
    }
}
