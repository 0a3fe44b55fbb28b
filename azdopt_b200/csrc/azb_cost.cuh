// azb_cost.cuh — the c21 cost: lambda_1 (adjacency spectral radius, f64) + matching number of a rooted tree.
// Replaces RootedOrderedTree::conjecture_2_1_cost (graph-state/src/rooted_tree/ordered_edge.rs:72-124), which
// calls a dense symmetric eigensolver (faer) and a leaf-stripping matching, plus the example's evaluate/squish
// closures (graph-state/examples/04-c21-tree.rs:70-74,96-102).
//
// lambda_1 by sectioning.  For a tree, x > lambda_1(T) iff the characteristic polynomial
// P_v(x) of every rooted subtree T_v is positive.  With Q_v = prod_{children c} P_c and
// P_v = x Q_v - sum_c Q_c prod_{c' != c} P_c' the pair (P_v, Q_v) is carried leaves-first without a division;
// parents[v] < v (rooted_tree/mod.rs:6) makes v = N-1..1 a leaves-first order.  The bracket
// [sqrt(max degree), sqrt(max #2-walks)] is narrowed by 11 rounds of a 5-level binary search on a 32-point grid.
// A warp evaluates the 31 interior grid points at once and then walks the binary search over the ballot; a single
// thread evaluates the 5 points the search visits.  Both produce bit-identical f64 results, and so does the
// oracle (oracle/azb_oracle.cpp lambda1_multisection), because every operation is a single IEEE f64 mul/add.
#pragma once
#include "azb_common.cuh"

#define AZB_SECTION_ROUNDS 11

// P_v(x) > 0 for every v?  parents: N bytes (shared or local memory).
template <int MAXV>
__device__ __forceinline__ bool azb_section_positive(uint32_t n, const uint8_t *parents, double x) {
    double Q[MAXV], S[MAXV];
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
        Q[v] = 1.0;
        S[v] = 0.0;
    }
    bool ok = true;
    for (uint32_t v = n - 1; v >= 1; --v) {
        double P = __dsub_rn(__dmul_rn(x, Q[v]), S[v]);
        ok = ok && (P > 0.0);
        uint32_t p = parents[v];
        double t0 = __dmul_rn(S[p], P);
        double t1 = __dmul_rn(Q[p], Q[v]);
        S[p] = __dadd_rn(t0, t1);
        Q[p] = __dmul_rn(Q[p], P);
    }
    double P0 = __dsub_rn(__dmul_rn(x, Q[0]), S[0]);
    return ok && (P0 > 0.0);
}

// bracket from vertex degrees; thread-serial (used by the thread-per-tree kernel)
__device__ __forceinline__ void azb_bracket_serial(uint32_t n, const uint8_t *parents, double &lo, double &hi) {
    uint32_t maxdeg = 0, maxw2 = 0;
    for (uint32_t v = 0; v < n; ++v) {
        uint32_t deg = v >= 1 ? 1u : 0u;
        uint32_t w2 = 0;
        for (uint32_t c = 1; c < n; ++c) deg += (parents[c] == v) ? 1u : 0u;
        // neighbours' degrees
        if (v >= 1) {
            uint32_t p = parents[v];
            uint32_t dp = p >= 1 ? 1u : 0u;
            for (uint32_t c = 1; c < n; ++c) dp += (parents[c] == p) ? 1u : 0u;
            w2 += dp;
        }
        for (uint32_t c = 1; c < n; ++c) {
            if (parents[c] == v) {
                uint32_t dc = 1;
                for (uint32_t c2 = 1; c2 < n; ++c2) dc += (parents[c2] == c) ? 1u : 0u;
                w2 += dc;
            }
        }
        maxdeg = max(maxdeg, deg);
        maxw2 = max(maxw2, w2);
    }
    lo = __dmul_rn(__dsqrt_rn((double)maxdeg), 1.0 - 9.313225746154785e-10);
    hi = __dmul_rn(__dsqrt_rn((double)maxw2), 1.0 + 9.313225746154785e-10);
}

// one thread, lazily evaluated binary search on the 32-grid
template <int MAXV>
__device__ double azb_lambda1_thread(uint32_t n, const uint8_t *parents) {
    double lo, hi;
    azb_bracket_serial(n, parents, lo, hi);
    for (int round = 0; round < AZB_SECTION_ROUNDS; ++round) {
        double w = __dmul_rn(__dsub_rn(hi, lo), 0.03125);
        int L = 0, H = 32;
#pragma unroll 1
        for (int lev = 0; lev < 5; ++lev) {
            int mid = (L + H) >> 1;
            double x = __dadd_rn(lo, __dmul_rn((double)mid, w));
            if (azb_section_positive<MAXV>(n, parents, x))
                H = mid;
            else
                L = mid;
        }
        double nlo = __dadd_rn(lo, __dmul_rn((double)L, w));
        double nhi = H == 32 ? hi : __dadd_rn(lo, __dmul_rn((double)H, w));
        lo = nlo;
        hi = nhi;
    }
    return __dmul_rn(0.5, __dadd_rn(lo, hi));
}

// a whole warp on one tree: lane l evaluates grid point l+1 (lane 31 idles), then everybody walks the search
// over the ballot.  parents in shared memory; scratch: 2*n u32 of shared memory (degrees).
template <int MAXV>
__device__ double azb_lambda1_warp(uint32_t n, const uint8_t *parents, uint32_t *scratch, int lane) {
    // degrees: lane v (and v+32) counts its children
    uint32_t maxdeg = 0, maxw2 = 0;
    for (uint32_t v = lane; v < n; v += 32) {
        uint32_t deg = v >= 1 ? 1u : 0u;
        for (uint32_t c = 1; c < n; ++c) deg += (parents[c] == v) ? 1u : 0u;
        scratch[v] = deg;
        maxdeg = max(maxdeg, deg);
    }
    __syncwarp();
    for (uint32_t v = lane; v < n; v += 32) {
        uint32_t w2 = v >= 1 ? scratch[parents[v]] : 0u;
        for (uint32_t c = 1; c < n; ++c) w2 += (parents[c] == v) ? scratch[c] : 0u;
        maxw2 = max(maxw2, w2);
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        maxdeg = max(maxdeg, __shfl_xor_sync(0xffffffffu, maxdeg, m));
        maxw2 = max(maxw2, __shfl_xor_sync(0xffffffffu, maxw2, m));
    }
    __syncwarp();
    double lo = __dmul_rn(__dsqrt_rn((double)maxdeg), 1.0 - 9.313225746154785e-10);
    double hi = __dmul_rn(__dsqrt_rn((double)maxw2), 1.0 + 9.313225746154785e-10);
    for (int round = 0; round < AZB_SECTION_ROUNDS; ++round) {
        double w = __dmul_rn(__dsub_rn(hi, lo), 0.03125);
        double x = __dadd_rn(lo, __dmul_rn((double)(lane + 1), w));
        bool pos = azb_section_positive<MAXV>(n, parents, x);
        uint32_t bal = __ballot_sync(0xffffffffu, pos);  // bit l <-> grid point l+1
        int L = 0, H = 32;
#pragma unroll
        for (int lev = 0; lev < 5; ++lev) {
            int mid = (L + H) >> 1;
            if ((bal >> (mid - 1)) & 1u)
                H = mid;
            else
                L = mid;
        }
        double nlo = __dadd_rn(lo, __dmul_rn((double)L, w));
        double nhi = H == 32 ? hi : __dadd_rn(lo, __dmul_rn((double)H, w));
        lo = nlo;
        hi = nhi;
    }
    return __dmul_rn(0.5, __dadd_rn(lo, hi));
}

// Maximum matching of a tree: leaves-first greedy over v = N-1..1 (a leaves-first order because parents[v] < v).
// The reference strips leaves round by round (ordered_edge.rs:94-124); both are maximum matchings, so the sizes
// agree and only the size enters the cost (04-c21-tree.rs:100).
__device__ __forceinline__ uint32_t azb_matching(uint32_t n, const uint8_t *parents) {
    unsigned long long used = 0ull;
    uint32_t m = 0;
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

// evaluate + squish: 04-c21-tree.rs:70-74,98-102
__device__ __forceinline__ float azb_evaluate(uint32_t mu, double lambda1, float c_lower, float slope) {
    float x = __fadd_rn((float)mu, (float)lambda1);
    x = __fsub_rn(x, c_lower);
    return __fmul_rn(slope, x);
}

// Stand-alone batched cost kernel: one thread per tree, parents staged through shared memory so that the
// global read is coalesced; outputs are SoA.  (azb_eval_costs)
template <int MAXV>
__global__ void __launch_bounds__(128) azb_cost_kernel(const uint8_t *__restrict__ parents, uint32_t m, uint32_t n,
                                                       float c_lower, float slope, double *__restrict__ lambda1,
                                                       uint32_t *__restrict__ mu, float *__restrict__ c,
                                                       uint32_t *__restrict__ err) {
    extern __shared__ uint8_t sm_par[];  // [128][n]
    const uint32_t first = blockIdx.x * blockDim.x;
    const uint32_t count = min((uint32_t)blockDim.x, m - first);
    for (uint32_t i = threadIdx.x; i < count * n; i += blockDim.x) sm_par[i] = parents[(size_t)first * n + i];
    __syncthreads();
    if (threadIdx.x >= count) return;
    const uint8_t *p = sm_par + threadIdx.x * n;
    double l1 = azb_lambda1_thread<MAXV>(n, p);
    uint32_t k = azb_matching(n, p);
    lambda1[first + threadIdx.x] = l1;
    mu[first + threadIdx.x] = k;
    c[first + threadIdx.x] = azb_evaluate(k, l1, c_lower, slope);
    if (!(l1 >= 1.4)) atomicMax(err, 5u);  // ordered_edge.rs:79
}
