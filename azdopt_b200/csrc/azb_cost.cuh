// azb_cost.cuh — the c21 cost: lambda_1 (adjacency spectral radius, f64) + matching number of a rooted tree.
// Replaces RootedOrderedTree::conjecture_2_1_cost (graph-state/src/rooted_tree/ordered_edge.rs:72-124), which
// calls a dense symmetric eigensolver (faer) and a leaf-stripping matching, plus the example's evaluate/squish
// closures (graph-state/examples/04-c21-tree.rs:70-74,96-102).
//
// lambda_1 by sectioning.  For a tree, x > lambda_1(T) iff the characteristic polynomial
// P_v(x) of every rooted subtree T_v is positive.  With Q_v = prod_{children c} P_c and
// S_v = sum_c Q_c prod_{c' != c} P_c', P_v = x Q_v - S_v, and folding a child c into its parent's pair is
//     S_v <- S_v P_c + Q_v Q_c ;  Q_v <- Q_v P_c            (no division; only signs are ever used).
// One warp owns one tree; lane l evaluates grid point l+1 of a 32-interval grid over the current bracket, the
// ballot of the 31 signs is walked as a 5-level binary search, and 11 rounds take the bracket
// [sqrt(max degree), sqrt(max #2-walks)] down to f64 resolution.
//
// Register-resident evaluation.  The recursion is run as a stack program in Sethi–Ullman order: a vertex first
// evaluates its "heaviest" internal child in place, then folds its leaves, then pushes one slot per remaining
// internal child.  A subtree that needs k slots has at least m(k) = 2 m(k-1) + 1 vertices (m(1) = 2), so 3 slots
// suffice for N <= 22, 4 for N <= 46 and 5 for N <= 94: the (Q, S) pairs of a whole evaluation live in 12-20
// registers, the program (3 bits per op, the same for all lanes and all rounds) in shared memory.
// The oracle (oracle/azb_oracle.cpp, lambda1_multisection) builds the same program and performs the same IEEE
// operations in the same order, so lambda_1 is bit-identical on both sides.
#pragma once
#include "azb_common.cuh"

#define AZB_SECTION_ROUNDS 11

#ifdef AZB_PROFILE
__device__ unsigned long long g_cost_prof[8];
#define CPROF_T0() long long cprof_t = clock64()
#define CPROF(i)                                                            \
    do {                                                                    \
        long long t1_ = clock64();                                          \
        if (lane == 0) atomicAdd(&g_cost_prof[i], (unsigned long long)(t1_ - cprof_t)); \
        cprof_t = t1_;                                                      \
    } while (0)
#else
#define CPROF_T0() do {} while (0)
#define CPROF(i) do {} while (0)
#endif

enum { OP_END = 0, OP_L0 = 1, OP_L = 2, OP_T = 3, OP_PUSH = 4, OP_POPF = 5 };

#define AZB_SECTION_MAX_DEPTH 5
struct __align__(16) CostScratch {  // per-warp shared memory
    uint32_t cnt[64];               // number of children
    uint32_t w2[64];                // sum of the children's degrees
    uint8_t su[64], s1[64], s2[64], heavy[64], len[64], start[64], cur[64];
    uint8_t ops[128];  // the program, one op per byte (at most ~2 n ops), read eight at a time
    // saved (Q, S) pairs of the section evaluation: [level][Q | S][lane].  LAST: a tree of n vertices uses
    // azb_stack_depth(n) - 1 levels, and the host sizes the scratch accordingly (azb_cost_scratch_words)
    double stk[(AZB_SECTION_MAX_DEPTH - 1) * 2 * 32];
};


__host__ __device__ __forceinline__ int azb_stack_depth(uint32_t n) { return n <= 22 ? 3 : (n <= 46 ? 4 : 5); }

// Degrees, the bracket, and the stack program of the tree `par` (n bytes in shared memory).
// Returns the number of ops; lo/hi receive the bracket.  Warp-collective.
__device__ __forceinline__ uint32_t azb_cost_prepare(uint32_t n, const uint8_t *par, CostScratch *cs, int lane,
                                                     double &lo, double &hi) {
    const uint32_t FULL = 0xffffffffu;
    cs->cnt[lane] = 0u;
    cs->cnt[lane + 32] = 0u;
    cs->w2[lane] = 0u;
    cs->w2[lane + 32] = 0u;
    cs->s1[lane] = 0;
    cs->s1[lane + 32] = 0;
    cs->s2[lane] = 0;
    cs->s2[lane + 32] = 0;
    cs->heavy[lane] = 0xff;
    cs->heavy[lane + 32] = 0xff;
    cs->len[lane] = 0;
    cs->len[lane + 32] = 0;
    __syncwarp();
    const bool a0 = lane >= 1 && (uint32_t)lane < n, a1 = (uint32_t)lane + 32 < n;
    const uint32_t p0 = a0 ? par[lane] : 0u, p1 = a1 ? par[lane + 32] : 0u;
    {   // children counts: one shared-memory add per (parent, half)
        uint32_t m = __match_any_sync(FULL, a0 ? p0 : (0x10000u | (uint32_t)lane));
        if (a0 && (__ffs(m) - 1) == lane) atomicAdd(&cs->cnt[p0], (uint32_t)__popc(m));
        if (n > 32) {
            m = __match_any_sync(FULL, a1 ? p1 : (0x10000u | (uint32_t)lane));
            if (a1 && (__ffs(m) - 1) == lane) atomicAdd(&cs->cnt[p1], (uint32_t)__popc(m));
        }
    }
    __syncwarp();
    // degree of v = children + (v has a parent); 2-walks from v = sum of the neighbours' degrees
    if (a0) atomicAdd(&cs->w2[p0], cs->cnt[lane] + 1u);
    if (a1) atomicAdd(&cs->w2[p1], cs->cnt[lane + 32] + 1u);
    __syncwarp();
    uint32_t maxdeg = 0, maxw2 = 0;
    if ((uint32_t)lane < n) {
        uint32_t d = cs->cnt[lane] + (lane >= 1 ? 1u : 0u);
        uint32_t w = cs->w2[lane] + (lane >= 1 ? cs->cnt[p0] + (p0 >= 1 ? 1u : 0u) : 0u);
        maxdeg = d;
        maxw2 = w;
    }
    if (a1) {
        uint32_t d = cs->cnt[lane + 32] + 1u;
        uint32_t w = cs->w2[lane + 32] + cs->cnt[p1] + (p1 >= 1 ? 1u : 0u);
        maxdeg = max(maxdeg, d);
        maxw2 = max(maxw2, w);
    }
    maxdeg = __reduce_max_sync(FULL, maxdeg);
    maxw2 = __reduce_max_sync(FULL, maxw2);
    // sqrt(max degree) <= lambda_1 <= sqrt(max 2-walk count); widened by 2^-30 relative
    lo = __dmul_rn(__dsqrt_rn((double)maxdeg), 1.0 - 9.313225746154785e-10);
    hi = __dmul_rn(__dsqrt_rn((double)maxw2), 1.0 + 9.313225746154785e-10);

    if (lane == 0) {
        // pass 1 (leaves first; the children of v have larger indices, so v's own entries are final when v comes up):
        // slots needed by each internal vertex, heaviest internal child of each vertex, program length of every subtree.
        // A child contributes 1 op (leaf), len + 1 (the heaviest internal child: OP_T) or len + 2 (PUSH ... POPF); which
        // child is the heaviest is only known once all children are in, so every internal child is entered as len + 2 and
        // the vertex takes the 1 back when its own turn comes.
#pragma unroll 1
        for (uint32_t v = n - 1; v >= 1; --v) {
            const uint32_t p = par[v];
            if (cs->cnt[v] == 0u) {
                cs->len[p] = (uint8_t)(cs->len[p] + 1u);
                continue;
            }
            const uint32_t lenv = cs->len[v] - (cs->heavy[v] != 0xff ? 1u : 0u);
            cs->len[v] = (uint8_t)lenv;
            const uint32_t suv = max((uint32_t)cs->s1[v], 1u + cs->s2[v]);
            cs->su[v] = (uint8_t)suv;
            if (suv > cs->s1[p]) {
                cs->s2[p] = cs->s1[p];
                cs->s1[p] = (uint8_t)suv;
                cs->heavy[p] = (uint8_t)v;
            } else if (suv > cs->s2[p]) {
                cs->s2[p] = (uint8_t)suv;
            }
            cs->len[p] = (uint8_t)(cs->len[p] + lenv + 2u);
        }
        if (cs->heavy[0] != 0xff) cs->len[0] = (uint8_t)(cs->len[0] - 1u);
        // pass 3 (root first): slot offsets; every vertex writes its own op(s)
        cs->start[0] = 0;
        cs->cur[0] = cs->heavy[0] != 0xff ? (uint8_t)(cs->len[cs->heavy[0]] + 1u) : 0;
#pragma unroll 1
        for (uint32_t v = 1; v < n; ++v) {
            const uint32_t p = par[v];
            const bool internal = cs->cnt[v] != 0u, is_heavy = cs->heavy[p] == v;
            const uint32_t lenv = cs->len[v];
            uint32_t slot;
            if (is_heavy) {
                slot = cs->start[p];
            } else {
                slot = cs->cur[p];
                cs->cur[p] = (uint8_t)(slot + (internal ? lenv + 2u : 1u));
            }
            if (!internal) {
                cs->ops[slot] = (cs->heavy[p] == 0xff && slot == cs->start[p]) ? OP_L0 : OP_L;
            } else {
                uint32_t st = slot;
                if (is_heavy) {
                    cs->ops[slot + lenv] = OP_T;
                } else {
                    cs->ops[slot] = OP_PUSH;
                    st = slot + 1u;
                    cs->ops[st + lenv] = OP_POPF;
                }
                cs->start[v] = (uint8_t)st;
                cs->cur[v] = (uint8_t)(st + (cs->heavy[v] != 0xff ? cs->len[cs->heavy[v]] + 1u : 0u));
            }
        }
        cs->ops[cs->len[0]] = OP_END;
    }
    __syncwarp();
    return (uint32_t)cs->len[0] + 1u;
}

// run the stack program at x: true iff P_v(x) > 0 for every vertex v.
// The pair on top of the stack lives in registers, the saved pairs below it in the warp's shared memory ([level][Q | S]
// [lane]: conflict-free), so PUSH and POPF are one store / load of a pair.  (The first version kept all DEPTH pairs in
// registers and shifted them on every PUSH / POPF, and decoded 3-bit ops from packed words: ~40 dependent instructions per
// op, 283 K cycles per evaluation at N = 64 — 86 % of a tree-step.)  The IEEE operations and their order per grid point
// are unchanged, so lambda_1 stays bit-identical to the oracle's lambda1_multisection.
template <int DEPTH>
__device__ __forceinline__ bool azb_section_positive(CostScratch *cs, int lane, double x) {
    double Q0 = 0., S0 = 0.;
    bool ok = true;
    const uint2 *ops8 = reinterpret_cast<const uint2 *>(cs->ops);
    double *sp = cs->stk + lane;
    uint2 next_word = ops8[0];
#pragma unroll 1
    for (uint32_t w = 1;; ++w) {
        // eight ops per shared-memory load, the next eight requested while these run; the eight are decoded by constant
        // byte extracts in an unrolled body (the walkers of N >= 23 are issue-bound in this loop: 16 warps per SM interpret
        // the same kind of program, so what counts is instructions per op)
        const uint2 word = next_word;
        next_word = ops8[w];  // (ops[] is 128 bytes for at most ~113 ops: the look-ahead stays inside it)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t op = ((k < 4 ? word.x : word.y) >> (8 * (k & 3))) & 0xffu;
            if (op == OP_L) {  // fold a leaf child (P = x, Q = 1)
                S0 = __fma_rn(S0, x, Q0);
                Q0 = __dmul_rn(Q0, x);
            } else if (op == OP_POPF) {  // finish the vertex on top, fold it into the pair below
                const double P = __fma_rn(x, Q0, -S0);
                ok = ok && (P > 0.0);
                sp -= 64;
                const double Q1 = sp[0], S1 = sp[32];
                const double t = __dmul_rn(Q1, Q0);
                S0 = __fma_rn(S1, P, t);
                Q0 = __dmul_rn(Q1, P);
            } else if (op == OP_PUSH) {  // (the pair on top is dead until the pushed subtree's first leaf sets it)
                sp[0] = Q0;
                sp[32] = S0;
                sp += 64;
            } else if (op == OP_T) {  // the slot held the heaviest child; it becomes the parent's pair
                const double P = __fma_rn(x, Q0, -S0);
                ok = ok && (P > 0.0);
                S0 = Q0;
                Q0 = P;
            } else if (op == OP_L0) {  // first child is a leaf
                Q0 = x;
                S0 = 1.0;
            } else {  // OP_END: the root
                const double P = __fma_rn(x, Q0, -S0);
                return ok && (P > 0.0);
            }
        }
    }
}

// A whole warp on one tree.  par: n bytes of shared memory.  Warp-collective; every lane returns lambda_1.
template <int DEPTH>
__device__ __forceinline__ double azb_lambda1_warp(uint32_t n, const uint8_t *par, CostScratch *cs, int lane) {
    double lo, hi;
    CPROF_T0();
    azb_cost_prepare(n, par, cs, lane, lo, hi);
    CPROF(0);
#pragma unroll 1
    for (int round = 0; round < AZB_SECTION_ROUNDS; ++round) {
        const double w = __dmul_rn(__dsub_rn(hi, lo), 0.03125);
        const double x = __dadd_rn(lo, __dmul_rn((double)(lane + 1), w));
        const bool pos = azb_section_positive<DEPTH>(cs, lane, x);
        const uint32_t bal = __ballot_sync(0xffffffffu, pos);  // bit l <-> grid point l+1 (bit 31 is never used)
        int L = 0, H = 32;
#pragma unroll
        for (int lev = 0; lev < 5; ++lev) {
            const int mid = (L + H) >> 1;
            if ((bal >> (mid - 1)) & 1u)
                H = mid;
            else
                L = mid;
        }
        const double nlo = __dadd_rn(lo, __dmul_rn((double)L, w));
        const double nhi = H == 32 ? hi : __dadd_rn(lo, __dmul_rn((double)H, w));
        lo = nlo;
        hi = nhi;
    }
    __syncwarp();
    CPROF(1);
    return __dmul_rn(0.5, __dadd_rn(lo, hi));
}

// ---------------------------------------------------------------------------------------------------------------
// N <= 22: lambda_1 AND mu from the matching polynomial.
// The characteristic polynomial of a forest is its matching polynomial: phi(x) = sum_k (-1)^k m_k x^(N-2k), with
// m_k the number of k-matchings; so lambda_1^2 is the largest root of p(y) = sum_k (-1)^k m_k y^(K-k) and K, the
// largest k with m_k > 0, IS the matching number.  m_k <= C(22-k, k) <= 6435, exact in u32.
//   1. m_k by a leaves-first DP over the tree, lane = k:  a_v = matchings of T_v, b_v = matchings of T_v - v;
//      folding child c into v:  b' = b * a_c,  a' = a * a_c + shift(b * b_c)   (polynomial products in k);
//   2. Laguerre's iteration from y0 = max #2-walks >= lambda_1^2: for a real-rooted polynomial the iterates from the
//      right of the largest root decrease monotonically onto it, cubically (plain Horner for p, p', p'');
//   3. two Newton steps with a compensated Horner evaluation of p (error-free TwoProd/TwoSum), which removes the
//      cancellation error of the alternating sum: worst relative error 1.3e-16 over random trees, paths, stars and
//      brooms (tests/test_oracle_golden.py), i.e. the f32 cost equals the dense eigensolver's.
// Every lane runs steps 2-3 redundantly (uniform); all operations are single IEEE f64 ops, mirrored by the oracle.
#define AZB_POLY_KMAX 11
#define AZB_POLY_NV 22


struct __align__(16) PolyScratch {  // per-warp shared memory, lives in the same bytes as CostScratch
    uint32_t A[AZB_POLY_NV][AZB_POLY_KMAX + 1];
    uint32_t B[AZB_POLY_NV][AZB_POLY_KMAX + 1];
    uint32_t cnt[32];
    uint32_t w2[32];
    double cf[AZB_POLY_KMAX + 1];  // signed coefficients of p(y), highest power first
};

// (tried as __noinline__ single copies to shrink the walkers' hot code: the calls cost more than the 0.9 KB saved)
__device__ __forceinline__ double azb_dsqrt(double x) { return __dsqrt_rn(x); }
__device__ __forceinline__ double azb_ddiv(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ void azb_two_sum(double a, double b, double &s, double &e) {
    s = __dadd_rn(a, b);
    const double bb = __dsub_rn(s, a);
    e = __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb));
}

// n <= 22.  par: n bytes of shared memory.  Warp-collective; every lane returns lambda_1 and *mu_out.
__device__ __forceinline__ double azb_lambda1_poly(uint32_t n, const uint8_t *par, PolyScratch *ps, int lane,
                                                uint32_t *mu_out) {
    const uint32_t FULL = 0xffffffffu;
    constexpr int KM = AZB_POLY_KMAX;
    CPROF_T0();
    // ---- children counts, max #2-walks (the start of the root iteration) ----
    ps->cnt[lane] = 0u;
    ps->w2[lane] = 0u;
    __syncwarp();
    const bool act = lane >= 1 && (uint32_t)lane < n;
    const uint32_t p0 = act ? par[lane] : 0u;
    {
        const uint32_t m = __match_any_sync(FULL, act ? p0 : (0x10000u | (uint32_t)lane));
        if (act && (__ffs(m) - 1) == lane) ps->cnt[p0] = (uint32_t)__popc(m);
    }
    __syncwarp();
    const uint32_t mycnt = ps->cnt[lane];
    if (act) atomicAdd(&ps->w2[p0], mycnt + 1u);
    const uint32_t leaf_mask = __ballot_sync(FULL, mycnt == 0u);  // bit v: vertex v has no children
    // leaf children fold as (a, b) -> (a + x b, b); the folds commute (polynomial products), so all L leaf
    // children of a vertex are folded at once by starting it at a = 1 + L x, b = 1
    uint32_t nleaf = 0;
    {
        const bool lf = act && mycnt == 0u;
        const uint32_t m = __match_any_sync(FULL, lf ? p0 : (0x10000u | (uint32_t)lane));
        // every leaf child of p0 sees the same group; the parent's lane needs the count: publish through cnt's
        // sibling array w2 is taken, so reuse A[p][1] directly below
        if (lf && (__ffs(m) - 1) == lane) nleaf = (uint32_t)__popc(m);
    }
#pragma unroll 1
    for (uint32_t i = lane; i < n * (KM + 1); i += 32) {
        const uint32_t one = (i % (KM + 1)) == 0u ? 1u : 0u;
        (&ps->A[0][0])[i] = one;
        (&ps->B[0][0])[i] = one;
    }
    __syncwarp();
    if (nleaf) ps->A[p0][1] = nleaf;
    uint32_t maxw2 = 0;
    if ((uint32_t)lane < n) maxw2 = ps->w2[lane] + (lane >= 1 ? ps->cnt[p0] + (p0 >= 1 ? 1u : 0u) : 0u);
    maxw2 = __reduce_max_sync(FULL, maxw2);
    __syncwarp();
    CPROF(0);
    // ---- 1. matching counts: fold the internal children, leaves first (parents[v] < v); lane = k ----
    const int k = lane;
    uint32_t todo = ~leaf_mask & ((n >= 32 ? 0u : (1u << n)) - 1u) & ~1u;  // internal vertices except the root
    while (todo) {
        const uint32_t v = 31u - __clz(todo);
        todo &= ~(1u << v);
        const uint32_t p = par[v];
        uint32_t na = 0, nb = 0;
        // the child's polynomials sit one coefficient per lane; their degree bounds the product loop
        const uint32_t av = k <= KM ? ps->A[v][k] : 0u, bv = k <= KM ? ps->B[v][k] : 0u;
        const int dav = 31 - __clz(__ballot_sync(FULL, av != 0u));
        // product loop in registers: lane k holds the parent's A_k and B_k; SA / SB are those polynomials shifted up by
        // i lanes (zero-filled from lane 0: one shuffle per iteration instead of clamped shared-memory reads), so
        //   nb_k = sum_i a_i B_(k-i),   na_k = sum_i a_i A_(k-i) + b_i B_(k-i-1)
        // (exact u32 arithmetic: the order of the terms does not matter)
        uint32_t SA = k <= KM ? ps->A[p][k] : 0u, SB = k <= KM ? ps->B[p][k] : 0u;
#pragma unroll 1
        for (int i = 0; i <= dav; ++i) {
            const uint32_t ac = __shfl_sync(FULL, av, i), bc = __shfl_sync(FULL, bv, i);
            const uint32_t ta = __shfl_up_sync(FULL, SA, 1), tb = __shfl_up_sync(FULL, SB, 1);
            const uint32_t SB1 = k == 0 ? 0u : tb;  // B shifted by i + 1
            nb += ac * SB;
            na += ac * SA + bc * SB1;
            SA = k == 0 ? 0u : ta;
            SB = SB1;
        }
        __syncwarp();
        if (k <= KM) {
            ps->A[p][k] = na;
            ps->B[p][k] = nb;
        }
        __syncwarp();
    }
    CPROF(1);
    // the matching number is the largest k with m_k > 0
    const uint32_t mk = k <= KM ? ps->A[0][k] : 0u;
    const uint32_t nz = __ballot_sync(FULL, mk != 0u);
    const uint32_t K = 31u - __clz(nz);
    *mu_out = K;
    // ---- coefficients: position j of a fixed 12-term Horner holds (-1)^(j - sh) m_(j - sh), sh = KM - K leading zeros, so
    //      the fixed-length Horner IS the K-term Horner.  m_k <= 6435 is exact in f64.  The signed values sit in shared
    //      memory, one per lane, and the Horner loops below are ROLLED: a dozen instructions instead of 5 KB of unrolled
    //      unpack-convert-fma code in the hottest loop of the walkers (instruction fetch is their largest stall).  A zero
    //      leading coefficient keeps the sign its position gives it (-0.0 or +0.0), as the oracle's mirror does.
    const int sh = KM - (int)K;
    if (k <= KM) {
        const uint32_t u = k >= sh ? ps->A[0][k - sh] : 0u;
        const bool neg = ((k & 1) != 0) != ((sh & 1) != 0);
        ps->cf[k] = neg ? -(double)u : (double)u;
    }
    __syncwarp();
    const double *cf = ps->cf;
    CPROF(2);
    // ---- 2. Laguerre from the right (plain Horner for p, p', p''/2): for a real-rooted polynomial the iterates
    //         decrease monotonically onto the largest root, cubically
    // (the Horner loops start at the first non-zero coefficient, position sh: through the leading zeros the fixed-length
    // form only carries signed zeros, and a signed zero added to the first non-zero term leaves it exact — s, d, h and
    // with them every iterate are bit for bit what the 12-term form gives, for 40 % fewer Horner steps at K = 6 … 7)
    double y = (double)maxw2;
    const double nn = (double)K, nm1 = (double)(K - 1u);
#pragma unroll 1
    for (int it = 0; it < 40; ++it) {
        double s = cf[sh], d = 0.0, h = 0.0;
#pragma unroll 1
        for (int j = sh + 1; j <= KM; ++j) {
            h = __fma_rn(h, y, d);
            d = __fma_rn(d, y, s);
            s = __fma_rn(s, y, cf[j]);
        }
        const double dd = __dmul_rn(2.0, h);
        double disc = __dmul_rn(nm1, __dsub_rn(__dmul_rn(nm1, __dmul_rn(d, d)), __dmul_rn(nn, __dmul_rn(s, dd))));
        if (!(disc >= 0.0)) disc = 0.0;
        const double den = __dadd_rn(d, azb_dsqrt(disc));
        if (!(den > 0.0)) break;
        const double yn = __dsub_rn(y, azb_ddiv(__dmul_rn(nn, s), den));
        if (!(yn < y)) break;
        const bool close = __dsub_rn(y, yn) < __dmul_rn(1e-5, yn);  // the two polishing steps finish from here
        y = yn;
        if (close) break;
    }
    CPROF(3);
    // ---- 3. polish: compensated Horner for p ----
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        double s = cf[sh], e = 0.0, t = s, d = 0.0;
#pragma unroll 1
        for (int j = sh + 1; j <= KM; ++j) {
            const double cj = cf[j];
            d = __fma_rn(d, y, t);
            t = __fma_rn(t, y, cj);
            const double pr = __dmul_rn(s, y);
            const double pi = __fma_rn(s, y, -pr);
            double sg;
            azb_two_sum(pr, cj, s, sg);
            e = __dadd_rn(__dmul_rn(e, y), __dadd_rn(pi, sg));
        }
        const double pv = __dadd_rn(s, e);
        if (!(d > 0.0)) break;
        y = __dsub_rn(y, azb_ddiv(pv, d));
    }
    CPROF(4);
    __syncwarp();
    return azb_dsqrt(y);
}

// cost of the tree in `par`: picks the method by size; every lane gets (lambda_1, mu)
template <int DEPTH>
__device__ __forceinline__ double azb_cost_warp(uint32_t n, const uint8_t *par, CostScratch *cs, int lane, uint32_t *mu);

// Maximum matching of a tree: leaves-first greedy over v = N-1..1 (a leaves-first order because parents[v] < v).
// The reference strips leaves round by round (ordered_edge.rs:94-124); both are maximum matchings, so the sizes
// agree and only the size enters the cost (04-c21-tree.rs:100).
__device__ __forceinline__ uint32_t azb_matching(uint32_t n, const uint8_t *parents) {
    unsigned long long used = 0ull;
    uint32_t m = 0;
#pragma unroll 1
    for (uint32_t v = n - 1; v >= 1; --v) {
        uint32_t p = parents[v];
        unsigned long long pair = (1ull << v) | (1ull << p);
        if ((used & pair) == 0ull) {
            used |= pair;
            ++m;
        }
    }
    return m;
}

template <int DEPTH>
__device__ __forceinline__ double azb_cost_warp(uint32_t n, const uint8_t *par, CostScratch *cs, int lane, uint32_t *mu) {
    // DEPTH == 3 <=> N <= 22 (azb_stack_depth): each instantiation carries only the method it uses
    if constexpr (DEPTH == 3) {
        return azb_lambda1_poly(n, par, reinterpret_cast<PolyScratch *>(cs), lane, mu);
    } else {
        const double l1 = azb_lambda1_warp<DEPTH>(n, par, cs, lane);
        *mu = azb_matching(n, par);
        return l1;
    }
}

// evaluate + squish: 04-c21-tree.rs:70-74,98-102
__device__ __forceinline__ float azb_evaluate(uint32_t mu, double lambda1, float c_lower, float slope) {
    float x = __fadd_rn((float)mu, (float)lambda1);
    x = __fsub_rn(x, c_lower);
    return __fmul_rn(slope, x);
}

#define AZB_COST_SCRATCH_BYTES (sizeof(CostScratch) > sizeof(PolyScratch) ? sizeof(CostScratch) : sizeof(PolyScratch))
#define AZB_COST_SCRATCH_WORDS ((uint32_t)(AZB_COST_SCRATCH_BYTES / 4))
// what a tree of n vertices needs: the matching-polynomial scratch up to 22 vertices, the section scratch beyond
__host__ __device__ __forceinline__ uint32_t azb_cost_scratch_words(uint32_t n) {
    return (uint32_t)((n <= 22 ? sizeof(PolyScratch)
                               : offsetof(CostScratch, stk) + (size_t)(azb_stack_depth(n) - 1) * 2 * 32 * sizeof(double)) / 4);
}

// Stand-alone batched cost kernel (azb_eval_costs): one warp per tree, 8 trees per block; the parent arrays of a
// block are one contiguous, coalesced read; outputs are SoA.
template <int DEPTH>
__global__ void __launch_bounds__(256) azb_cost_kernel(const uint8_t *__restrict__ parents, uint32_t m, uint32_t n,
                                                       float c_lower, float slope, double *__restrict__ lambda1,
                                                       uint32_t *__restrict__ mu, float *__restrict__ c,
                                                       uint32_t *__restrict__ err) {
    __shared__ __align__(16) uint8_t scratch_bytes[8][AZB_COST_SCRATCH_BYTES];
    __shared__ __align__(16) uint8_t sm_par[8 * 64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t first = blockIdx.x * 8u;
    const uint32_t count = min(8u, m - first);
    for (uint32_t i = threadIdx.x; i < count * n; i += blockDim.x) sm_par[(i / n) * 64 + (i % n)] = parents[(size_t)first * n + i];
    __syncthreads();
    if ((uint32_t)warp >= count) return;
    const uint8_t *p = sm_par + warp * 64;
    uint32_t k = 0;
    const double l1 = azb_cost_warp<DEPTH>(n, p, reinterpret_cast<CostScratch *>(scratch_bytes[warp]), lane, &k);
    if (lane == 0) {
        lambda1[first + warp] = l1;
        mu[first + warp] = k;
        c[first + warp] = azb_evaluate(k, l1, c_lower, slope);
        if (!(l1 >= 1.4)) atomicMax(err, 5u);  // ordered_edge.rs:79
    }
}
