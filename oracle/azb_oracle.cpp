/*
 * oracle/azb_oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference's batched search step for the c21 example
 * (az-discrete-opt `NablaOptimizer::par_roll_out_episodes` over
 * graph-state `ROTModifyParentsOnce`).  Written from the reference's
 * behaviour, citing the file:line each function follows; no reference source
 * is copied.  See azb_oracle.h for who may use it and for the parity status
 * ("parity unpinned" for the search-DAG part).
 *
 * Build: make -C oracle   (g++ -O2 -ffp-contract=off; IEEE f32/f64, no FMA
 * contraction, so that the multisection lambda_1 is bit-identical to the CUDA
 * kernels' arithmetic).
 */
#include "azb_oracle.h"

#include <algorithm>
#include <array>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <thread>
#include <vector>

namespace {

constexpr uint32_t MAXN = 64;

// ----------------------------------------------------------------------------
// A.1 indexing
// ----------------------------------------------------------------------------

// simple_graph/edge.rs:48-53
inline uint32_t colex_position(uint32_t u, uint32_t v) {
    uint32_t mx = u > v ? u : v, mn = u > v ? v : u;
    uint32_t last_pos = mx * (mx + 1) / 2;
    return last_pos - (mx - mn);
}

// simple_graph/edge.rs:55-65 (the reference searches linearly; so do we)
inline void from_colex_position(uint32_t pos, uint32_t &mx, uint32_t &mn) {
    uint32_t v = 1;
    for (;;) {
        uint32_t last_position = v * (v + 1) / 2;
        if (pos < last_position) {
            uint32_t diff = last_position - pos;
            mx = v;
            mn = v - diff;
            return;
        }
        ++v;
    }
}

// rooted_tree/ordered_edge.rs:35-38
inline uint32_t index_ignoring_edge_0_1(uint32_t parent, uint32_t child) {
    return colex_position(parent, child) - 1;
}
// rooted_tree/ordered_edge.rs:40-42
inline void from_index_ignoring_edge_0_1(uint32_t index, uint32_t &parent, uint32_t &child) {
    from_colex_position(index + 1, child, parent);
}

inline uint32_t action_dim(uint32_t n) { return (n - 1) * (n - 2) / 2 - 1; }  // rooted_tree/space.rs:48
inline uint32_t mask_words(uint32_t n) { return (action_dim(n) + 31) / 32; }

// examples/04-c21-tree.rs:59-68
inline uint32_t c_upper_bound(uint32_t n) {
    uint32_t s = 0;
    while ((s + 1) * (s + 1) <= n - 1) ++s;  // isqrt
    uint32_t sq = (s * s == n - 1) ? s : s + 1;
    return sq + (n + 1) / 2;
}

// ----------------------------------------------------------------------------
// A.2 state (rooted_tree/mod.rs:8-10, modify_parent_once.rs:8-11)
// The BTreeSet<usize> of permitted actions is held as a bit mask plus ordered
// iteration; ascending bit order == BTreeSet iteration order.
// ----------------------------------------------------------------------------
struct State {
    std::array<uint8_t, MAXN> parents{};
    std::array<uint32_t, 61> permitted{};  // W <= 61 words for N <= 64
};

inline bool mask_get(const uint32_t *m, uint32_t i) { return (m[i >> 5] >> (i & 31)) & 1u; }
inline void mask_set(uint32_t *m, uint32_t i) { m[i >> 5] |= 1u << (i & 31); }
inline void mask_clear(uint32_t *m, uint32_t i) { m[i >> 5] &= ~(1u << (i & 31)); }

// rooted_tree/space.rs:56-73 — act: set the parent, then forget every action of that child
inline void space_act(uint32_t /*n*/, State &s, uint32_t action) {
    uint32_t parent, child;
    from_index_ignoring_edge_0_1(action, parent, child);
    s.parents[child] = (uint8_t)parent;  // ordered_edge.rs:46-50
    for (uint32_t u = 0; u < child; ++u) mask_clear(s.permitted.data(), index_ignoring_edge_0_1(u, child));
}

// rooted_tree/mod.rs:60-72
inline uint32_t current_edge_indices(uint32_t n, const State &s, uint32_t *out) {
    uint32_t k = 0;
    for (uint32_t child = 2; child + 1 < n; ++child) out[k++] = index_ignoring_edge_0_1(s.parents[child], child);
    return k;
}

// rooted_tree/space.rs:75-89 — permitted actions, ascending, that are not a current edge
inline uint32_t space_action_data(uint32_t n, const State &s, uint32_t *out) {
    uint32_t cur[MAXN];
    uint32_t ncur = current_edge_indices(n, s, cur);
    uint32_t a_dim = action_dim(n), k = 0;
    for (uint32_t a = 0; a < a_dim; ++a) {
        if (!mask_get(s.permitted.data(), a)) continue;
        bool is_current = false;
        for (uint32_t i = 0; i < ncur; ++i) is_current |= (cur[i] == a);
        if (!is_current) out[k++] = a;
    }
    return k;
}

// nabla/space/mod.rs:27-29
inline bool space_is_terminal(uint32_t n, const State &s) {
    uint32_t tmp[2048];
    return space_action_data(n, s, tmp) == 0;
}

// rooted_tree/space.rs:91-101
inline void space_write_vec(uint32_t n, const State &s, float *v) {
    uint32_t a_dim = action_dim(n);
    std::fill(v, v + 2 * a_dim, 0.0f);
    uint32_t cur[MAXN];
    uint32_t ncur = current_edge_indices(n, s, cur);
    for (uint32_t i = 0; i < ncur; ++i) v[cur[i]] = 1.0f;
    for (uint32_t a = 0; a < a_dim; ++a)
        if (mask_get(s.permitted.data(), a)) v[a_dim + a] = 1.0f;
}

// ----------------------------------------------------------------------------
// A.3 cost (rooted_tree/ordered_edge.rs:72-124)
// ----------------------------------------------------------------------------

// ordered_edge.rs:84-91 + :74-78.  faer 0.15 `selfadjoint_eigenvalues` is an
// un-vendored dependency; its published algorithm for small matrices is
// Householder tridiagonalisation followed by a QR/QL eigenvalue iteration.  Any
// backward-stable f64 symmetric eigensolver gives lambda_1 to ~1e-15 relative.
double lambda1_dense(uint32_t n, const uint8_t *parents) {
    double a[MAXN][MAXN];
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < n; ++j) a[i][j] = 0.0;
    for (uint32_t i = 1; i < n; ++i) {
        a[i][parents[i]] = 1.0;
        a[parents[i]][i] = 1.0;
    }
    double d[MAXN], e[MAXN];
    // Householder reduction to tridiagonal form, eigenvalues only
    for (uint32_t k = 0; k + 2 < n; ++k) {
        uint32_t m = n - k - 1;  // length of the column below the diagonal
        double u[MAXN];
        double norm2 = 0.0;
        for (uint32_t i = 0; i < m; ++i) {
            u[i] = a[k + 1 + i][k];
            norm2 += u[i] * u[i];
        }
        double norm = std::sqrt(norm2);
        double rest2 = norm2 - u[0] * u[0];
        if (rest2 == 0.0) {  // already tridiagonal in this column
            e[k] = u[0];
            continue;
        }
        double alpha = u[0] > 0.0 ? -norm : norm;
        u[0] -= alpha;
        double un2 = 0.0;
        for (uint32_t i = 0; i < m; ++i) un2 += u[i] * u[i];
        double inv = 1.0 / std::sqrt(un2);
        for (uint32_t i = 0; i < m; ++i) u[i] *= inv;
        // B' = B - 2 u w^T - 2 w u^T with q = B u, w = q - (u.q) u
        double q[MAXN];
        double uq = 0.0;
        for (uint32_t i = 0; i < m; ++i) {
            double acc = 0.0;
            for (uint32_t j = 0; j < m; ++j) acc += a[k + 1 + i][k + 1 + j] * u[j];
            q[i] = acc;
            uq += u[i] * acc;
        }
        for (uint32_t i = 0; i < m; ++i) q[i] -= uq * u[i];
        for (uint32_t i = 0; i < m; ++i)
            for (uint32_t j = 0; j < m; ++j) a[k + 1 + i][k + 1 + j] -= 2.0 * (u[i] * q[j] + q[i] * u[j]);
        e[k] = alpha;
    }
    for (uint32_t i = 0; i < n; ++i) d[i] = a[i][i];
    if (n >= 2) e[n - 2] = a[n - 1][n - 2];
    e[n - 1] = 0.0;
    // implicit QL with Wilkinson shifts, eigenvalues only.  Deflation uses an absolute floor
    // eps*||T|| as well: only lambda_1 (= ||T||_2) is wanted, and trees have large null spaces.
    double anorm = 0.0;
    for (uint32_t i = 0; i < n; ++i) anorm = std::max(anorm, std::fabs(d[i]) + std::fabs(e[i]));
    for (uint32_t l = 0; l < n; ++l) {
        for (int iter = 0; iter < 200; ++iter) {
            uint32_t m = l;
            for (; m + 1 < n; ++m) {
                double dd = std::fabs(d[m]) + std::fabs(d[m + 1]);
                if (std::fabs(e[m]) <= 2.2204460492503131e-16 * dd || std::fabs(e[m]) <= 1e-18 * anorm) break;
            }
            if (m == l) break;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
            double r = std::hypot(g, 1.0);
            g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
            double s = 1.0, c = 1.0, p = 0.0;
            bool early = false;
            for (int i = (int)m - 1; i >= (int)l; --i) {
                double f = s * e[i], b = c * e[i];
                r = std::hypot(f, g);
                e[i + 1] = r;
                if (r == 0.0) {
                    d[i + 1] -= p;
                    e[m] = 0.0;
                    early = true;
                    break;
                }
                s = f / r;
                c = g / r;
                g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b;
                p = s * r;
                d[i + 1] = g + p;
                g = c * r - b;
            }
            if (early) continue;
            d[l] -= p;
            e[l] = g;
            e[m] = 0.0;
        }
    }
    double best = d[0];
    for (uint32_t i = 1; i < n; ++i) best = std::max(best, d[i]);  // ordered_edge.rs:76-78
    return best;
}

// independent cross-check: cyclic Jacobi
double jacobi_max_eigenvalue(uint32_t n, double (*a)[MAXN]);
double lambda1_jacobi(uint32_t n, const uint8_t *parents) {
    double a[MAXN][MAXN];
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < n; ++j) a[i][j] = 0.0;
    for (uint32_t i = 1; i < n; ++i) {
        a[i][parents[i]] = 1.0;
        a[parents[i]][i] = 1.0;
    }
    return jacobi_max_eigenvalue(n, a);
}

double jacobi_max_eigenvalue(uint32_t n, double (*a)[MAXN]) {
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (uint32_t p = 0; p < n; ++p)
            for (uint32_t q = p + 1; q < n; ++q) off += a[p][q] * a[p][q];
        if (off < 1e-32) break;
        for (uint32_t p = 0; p < n; ++p)
            for (uint32_t q = p + 1; q < n; ++q) {
                if (a[p][q] == 0.0) continue;
                double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (uint32_t k = 0; k < n; ++k) {
                    double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (uint32_t k = 0; k < n; ++k) {
                    double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
            }
    }
    double best = a[0][0];
    for (uint32_t i = 1; i < n; ++i) best = std::max(best, a[i][i]);
    return best;
}

// The tree-specific method the CUDA kernels use, restated serially with the same IEEE operations in the same
// order (DESIGN.md "lambda_1 by 32-ary section").  x > lambda_1(T) iff every subtree characteristic polynomial
// P_v(x) is positive; P_v = x Q_v - S_v with Q_v = prod_children P_c, carried without division, leaves first
// (parents[v] < v, rooted_tree/mod.rs:6).  The recursion is run as a stack program in Sethi-Ullman order (heaviest
// internal child first and in place; one extra slot per further internal child), exactly the program
// azdopt_b200/csrc/azb_cost.cuh builds: ops END, L0, L, T, PUSH, POPF.
enum { OP_END = 0, OP_L0 = 1, OP_L = 2, OP_T = 3, OP_PUSH = 4, OP_POPF = 5 };

inline uint32_t build_program(uint32_t n, const uint8_t *par, uint8_t *ops) {
    uint32_t cnt[MAXN] = {0}, s1[MAXN] = {0}, s2[MAXN] = {0}, heavy[MAXN], len[MAXN] = {0}, start[MAXN] = {0}, cur[MAXN] = {0};
    for (uint32_t v = 0; v < n; ++v) heavy[v] = 0xff;
    for (uint32_t v = 1; v < n; ++v) cnt[par[v]]++;
    for (uint32_t v = n - 1; v >= 1; --v) {  // slots needed; heaviest internal child (ties: the larger index)
        if (cnt[v] == 0) continue;
        uint32_t suv = std::max(s1[v], 1u + s2[v]);
        uint32_t p = par[v];
        if (suv > s1[p]) {
            s2[p] = s1[p];
            s1[p] = suv;
            heavy[p] = v;
        } else if (suv > s2[p]) {
            s2[p] = suv;
        }
    }
    for (uint32_t v = n - 1; v >= 1; --v) {  // program length of every subtree
        uint32_t p = par[v];
        len[p] += cnt[v] == 0 ? 1u : len[v] + (heavy[p] == v ? 1u : 2u);
    }
    start[0] = 0;
    cur[0] = heavy[0] != 0xff ? len[heavy[0]] + 1u : 0u;
    for (uint32_t v = 1; v < n; ++v) {  // slot offsets, root first; children of a vertex in ascending order
        uint32_t p = par[v];
        bool internal = cnt[v] != 0, is_heavy = heavy[p] == v;
        uint32_t slot;
        if (is_heavy) {
            slot = start[p];
        } else {
            slot = cur[p];
            cur[p] = slot + (internal ? len[v] + 2u : 1u);
        }
        if (!internal) {
            ops[slot] = (heavy[p] == 0xff && slot == start[p]) ? OP_L0 : OP_L;
        } else {
            uint32_t st = slot;
            if (is_heavy) {
                ops[slot + len[v]] = OP_T;
            } else {
                ops[slot] = OP_PUSH;
                st = slot + 1;
                ops[st + len[v]] = OP_POPF;
            }
            start[v] = st;
            cur[v] = st + (heavy[v] != 0xff ? len[heavy[v]] + 1u : 0u);
        }
    }
    ops[len[0]] = OP_END;
    return len[0] + 1;
}

inline bool section_positive(const uint8_t *ops, double x) {
    double Q[8] = {0}, S[8] = {0};  // Q[0], S[0] = top of the stack
    bool ok = true;
    for (uint32_t i = 0;; ++i) {
        switch (ops[i]) {
            case OP_L:
                S[0] = std::fma(S[0], x, Q[0]);
                Q[0] = Q[0] * x;
                break;
            case OP_L0:
                Q[0] = x;
                S[0] = 1.0;
                break;
            case OP_T: {
                double P = std::fma(x, Q[0], -S[0]);
                ok = ok && (P > 0.0);
                S[0] = Q[0];
                Q[0] = P;
                break;
            }
            case OP_PUSH:
                for (int k = 7; k >= 1; --k) {
                    Q[k] = Q[k - 1];
                    S[k] = S[k - 1];
                }
                break;
            case OP_POPF: {
                double P = std::fma(x, Q[0], -S[0]);
                ok = ok && (P > 0.0);
                double t = Q[1] * Q[0];
                S[0] = std::fma(S[1], P, t);
                Q[0] = Q[1] * P;
                for (int k = 1; k < 7; ++k) {
                    Q[k] = Q[k + 1];
                    S[k] = S[k + 1];
                }
                break;
            }
            default: {
                double P = std::fma(x, Q[0], -S[0]);
                return ok && (P > 0.0);
            }
        }
    }
}

double lambda1_multisection(uint32_t n, const uint8_t *parents) {
    uint32_t deg[MAXN] = {0};
    for (uint32_t v = 1; v < n; ++v) {
        deg[v]++;
        deg[parents[v]]++;
    }
    uint32_t w2[MAXN] = {0};  // number of 2-walks from v
    for (uint32_t v = 1; v < n; ++v) {
        w2[v] += deg[parents[v]];
        w2[parents[v]] += deg[v];
    }
    uint32_t maxdeg = 0, maxw2 = 0;
    for (uint32_t v = 0; v < n; ++v) {
        maxdeg = std::max(maxdeg, deg[v]);
        maxw2 = std::max(maxw2, w2[v]);
    }
    // sqrt(max degree) <= lambda_1 <= sqrt(max 2-walk count); widen by 2^-30 relative
    double lo = std::sqrt((double)maxdeg) * (1.0 - 9.313225746154785e-10);
    double hi = std::sqrt((double)maxw2) * (1.0 + 9.313225746154785e-10);
    // 11 rounds of a 5-level binary search on the 32-point grid lo + k*w (grid point 32 is `hi`, positive by
    // construction).  The CUDA warp evaluates all 31 interior points and walks the same search over its ballot.
    uint8_t ops[2 * MAXN];
    build_program(n, parents, ops);
    for (int round = 0; round < 11; ++round) {
        double w = (hi - lo) * 0.03125;
        int L = 0, H = 32;
        for (int lev = 0; lev < 5; ++lev) {
            int mid = (L + H) >> 1;
            double x = lo + (double)mid * w;
            if (section_positive(ops, x))
                H = mid;
            else
                L = mid;
        }
        double nlo = lo + (double)L * w;
        double nhi = H == 32 ? hi : lo + (double)H * w;
        lo = nlo;
        hi = nhi;
    }
    return 0.5 * (lo + hi);
}

// N <= 22, the method azdopt_b200/csrc/azb_cost.cuh uses there: lambda_1^2 is the largest root of the matching
// polynomial p(y) = sum_k (-1)^k m_k y^(K-k) (a forest's characteristic polynomial is its matching polynomial).
// m_k by a leaves-first DP, Newton from y0 = max #2-walks (plain Horner), three Newton steps with a compensated
// Horner.  Same IEEE operations in the same order as the kernel => bit-identical lambda_1.
inline void two_sum(double a, double b, double &s, double &e) {
    s = a + b;
    double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}

double lambda1_poly(uint32_t n, const uint8_t *par, uint32_t *mu_out) {
    constexpr int KM = 11;
    uint32_t A[22][KM + 1], B[22][KM + 1];
    uint32_t dA[22] = {0}, dB[22] = {0};
    for (uint32_t v = 0; v < n; ++v)
        for (int k = 0; k <= KM; ++k) A[v][k] = B[v][k] = k == 0 ? 1u : 0u;
    for (uint32_t v = n - 1; v >= 1; --v) {
        uint32_t p = par[v];
        uint32_t na[KM + 1], nb[KM + 1];
        for (int k = 0; k <= KM; ++k) {
            uint32_t a = 0, b = 0;
            if (dA[v] == 0) {
                a = A[p][k] + (k >= 1 ? B[p][k - 1] : 0u);
                b = B[p][k];
            } else {
                for (uint32_t i = 0; i <= dA[v] && (int)i <= k; ++i) {
                    b += A[v][i] * B[p][k - i];
                    a += A[v][i] * A[p][k - i];
                    if ((int)i < k) a += B[v][i] * B[p][k - 1 - i];
                }
            }
            na[k] = a;
            nb[k] = b;
        }
        uint32_t nda = std::max(dA[p] + dA[v], dB[p] + dB[v] + 1u), ndb = dB[p] + dA[v];
        for (int k = 0; k <= KM; ++k) {
            A[p][k] = na[k];
            B[p][k] = nb[k];
        }
        dA[p] = std::min<uint32_t>(nda, KM);
        dB[p] = std::min<uint32_t>(ndb, KM);
    }
    uint32_t K = 0;
    for (int k = 0; k <= KM; ++k)
        if (A[0][k] != 0) K = (uint32_t)k;
    if (mu_out) *mu_out = K;
    double c[KM + 1];
    for (int j = 0; j <= KM; ++j) {
        int kk = j - (KM - (int)K);
        double v = kk < 0 ? 0.0 : (double)A[0][kk];
        c[j] = kk < 0 ? 0.0 : ((kk & 1) ? -v : v);
    }
    uint32_t deg[22] = {0}, w2[22] = {0};
    for (uint32_t v = 1; v < n; ++v) {
        deg[v]++;
        deg[par[v]]++;
    }
    for (uint32_t v = 1; v < n; ++v) {
        w2[v] += deg[par[v]];
        w2[par[v]] += deg[v];
    }
    uint32_t maxw2 = 0;
    for (uint32_t v = 0; v < n; ++v) maxw2 = std::max(maxw2, w2[v]);
    double y = (double)maxw2;
    const double nn = (double)K, nm1 = (double)(K - 1u);
    for (int it = 0; it < 40; ++it) {  // Laguerre from the right of the largest root
        double s = c[0], d = 0.0, h = 0.0;
        for (int j = 1; j <= KM; ++j) {
            h = std::fma(h, y, d);
            d = std::fma(d, y, s);
            s = std::fma(s, y, c[j]);
        }
        double dd = 2.0 * h;
        double disc = nm1 * ((nm1 * (d * d)) - (nn * (s * dd)));
        if (!(disc >= 0.0)) disc = 0.0;
        double den = d + std::sqrt(disc);
        if (!(den > 0.0)) break;
        double yn = y - (nn * s) / den;
        if (!(yn < y)) break;
        bool close = (y - yn) < 1e-5 * yn;  // the two polishing steps finish from here
        y = yn;
        if (close) break;
    }
    for (int it = 0; it < 2; ++it) {
        double s = c[0], e = 0.0, t = c[0], d = 0.0;
        for (int j = 1; j <= KM; ++j) {
            d = std::fma(d, y, t);
            t = std::fma(t, y, c[j]);
            double pr = s * y;
            double pi = std::fma(s, y, -pr);
            double sg;
            two_sum(pr, c[j], s, sg);
            e = e * y + (pi + sg);
        }
        double pv = s + e;
        if (!(d > 0.0)) break;
        y = y - pv / d;
    }
    return std::sqrt(y);
}

// ordered_edge.rs:94-124 — repeated leaf stripping; only the size is used by the cost
uint32_t maximum_matching(uint32_t n, const uint8_t *parents) {
    bool available[MAXN];
    for (uint32_t i = 0; i < n; ++i) available[i] = true;
    uint32_t m = 0;
    for (;;) {
        bool next_leaf[MAXN];
        for (uint32_t i = 0; i < n; ++i) next_leaf[i] = available[i];
        for (uint32_t i = 1; i < n; ++i)
            if (available[i]) next_leaf[parents[i]] = false;
        for (uint32_t i = 1; i < n; ++i) {
            if (next_leaf[i]) {
                available[i] = false;
                uint32_t parent = parents[i];
                if (available[parent]) {
                    available[parent] = false;
                    ++m;
                }
            }
        }
        uint32_t num_available = 0;
        for (uint32_t i = 0; i < n; ++i) num_available += available[i];
        if (num_available < 2) break;
    }
    return m;
}

// independent check: leaves-first greedy (v = N-1..1 is a leaves-first order)
uint32_t matching_greedy(uint32_t n, const uint8_t *parents) {
    uint64_t used = 0;
    uint32_t m = 0;
    for (uint32_t v = n - 1; v >= 1; --v) {
        uint32_t p = parents[v];
        if (!((used >> v) & 1) && !((used >> p) & 1)) {
            used |= (1ull << v) | (1ull << p);
            ++m;
        }
    }
    return m;
}

struct Cost {  // connected_bitset_graph/mod.rs:340-344 (only matching.len() is ever read on this path)
    double lambda_1 = 0.0;
    uint32_t mu = 0;
};

struct Space {  // rooted_tree/space.rs:13-19 with the example's closures (04-c21-tree.rs:96-105) as data
    uint32_t n = 0, a_dim = 0, words = 0;
    float c_lower = 2.0f, slope = 0.0f;
    int lambda_method = ORC_LAMBDA_DENSE;
    std::vector<uint32_t> tol;
    uint32_t tol_default = 25;

    Cost cost(const State &s, bool *bad = nullptr) const {  // ordered_edge.rs:72-82
        Cost c;
        switch (lambda_method) {
            case ORC_LAMBDA_JACOBI: c.lambda_1 = lambda1_jacobi(n, s.parents.data()); break;
            case ORC_LAMBDA_MULTISECTION:  // "what the CUDA kernels do": matching polynomial up to 22 vertices
                c.lambda_1 = n <= 22 ? lambda1_poly(n, s.parents.data(), nullptr) : lambda1_multisection(n, s.parents.data());
                break;
            case ORC_LAMBDA_SECTION_ONLY: c.lambda_1 = lambda1_multisection(n, s.parents.data()); break;
            case ORC_LAMBDA_POLY: c.lambda_1 = lambda1_poly(n, s.parents.data(), nullptr); break;
            default: c.lambda_1 = lambda1_dense(n, s.parents.data()); break;
        }
        if (bad && !(c.lambda_1 >= 1.4)) *bad = true;  // ordered_edge.rs:79 assert
        c.mu = maximum_matching(n, s.parents.data());
        return c;
    }
    // 04-c21-tree.rs:98-102 + squish :70-74
    float evaluate(const Cost &c) const {
        float x = (float)c.mu + (float)c.lambda_1;
        x = x - c_lower;
        return slope * x;
    }
    // 04-c21-tree.rs:103
    float g_theta_star_sa(float c_s, float h_theta_sa) const { return c_s - h_theta_sa; }
    // 04-c21-tree.rs:104
    float h_sa(float /*c_s*/, float c_as_star) const { return c_as_star; }
    // 04-c21-tree.rs:136-138
    uint32_t n_as_tol(size_t len) const { return len < tol.size() ? tol[len] : tol_default; }
};

// ----------------------------------------------------------------------------
// A.4 the search DAG (az-discrete-opt/src/nabla/tree)
// ----------------------------------------------------------------------------
struct StateWeight {  // state_weight.rs:4-10
    float c, c_t_star;
    uint32_t n_t, exhausted_children;
    uint32_t lo, hi;  // actions: Range<u32>
    explicit StateWeight(float c_) : c(c_), c_t_star(c_), n_t(0), exhausted_children(0), lo(0), hi(0) {}  // :13-21
    bool is_active() const { return lo + exhausted_children < hi; }                                       // :31-33
};

struct ActionPrediction {  // arc_weight.rs:11-16
    uint32_t a_id;
    float g_theta_sa;
    uint32_t edge_id;  // Option<EdgeIndex>
};

// petgraph 0.6 DiGraph storage (un-vendored dependency): every node heads two
// singly linked lists (outgoing, incoming) of edges; add_edge pushes at the
// head, so iteration is newest edge first.
struct GNode {
    StateWeight w;
    uint32_t next[2];
};
struct GEdge {
    uint32_t prediction_pos;  // arc_weight.rs:4-8
    uint32_t next[2];
    uint32_t node[2];  // source, target
};

using ActionSet = std::vector<uint16_t>;  // path/set.rs:5-8: BTreeSet<usize>, kept sorted ascending; Ord = lexicographic

struct SearchTree {  // tree/mod.rs:28-32
    std::map<ActionSet, uint32_t> positions;
    std::vector<GNode> nodes;
    std::vector<GEdge> edges;
    std::vector<ActionPrediction> predictions;
    std::vector<ActionSet> node_keys;  // for dumps only

    void clear() {  // tree/mod.rs:45-49
        positions.clear();
        nodes.clear();
        edges.clear();
        predictions.clear();
        node_keys.clear();
    }
    uint32_t graph_add_node(const StateWeight &w) {
        nodes.push_back(GNode{w, {ORC_NONE, ORC_NONE}});
        return (uint32_t)nodes.size() - 1;
    }
    uint32_t graph_add_edge(uint32_t a, uint32_t b, uint32_t prediction_pos) {
        uint32_t idx = (uint32_t)edges.size();
        GEdge e;
        e.prediction_pos = prediction_pos;
        e.node[0] = a;
        e.node[1] = b;
        e.next[0] = nodes[a].next[0];
        e.next[1] = nodes[b].next[1];
        nodes[a].next[0] = idx;
        nodes[b].next[1] = idx;
        edges.push_back(e);
        return idx;
    }
    // graph_operations.rs:8-16
    uint32_t add_node(const ActionSet &p, const StateWeight &w) {
        uint32_t index = graph_add_node(w);
        positions.emplace(p, index);
        node_keys.push_back(p);
        return index;
    }
    // graph_operations.rs:18-30
    uint32_t add_arc(uint32_t parent, uint32_t child, uint32_t prediction_pos) {
        uint32_t arc = graph_add_edge(parent, child, prediction_pos);
        predictions[prediction_pos].edge_id = arc;
        return arc;
    }
};

struct Walker {
    orc_counters k{};
};

// graph_operations.rs:32-56
void add_actions(SearchTree &t, uint32_t id, const Space &space, const State &state, const float *h_theta, orc_counters &k) {
    float c = t.nodes[id].w.c;
    uint32_t start = (uint32_t)t.predictions.size();
    uint32_t acts[2048];
    uint32_t na = space_action_data(space.n, state, acts);
    for (uint32_t i = 0; i < na; ++i) {
        float h = h_theta[acts[i]];
        t.predictions.push_back(ActionPrediction{acts[i], space.g_theta_star_sa(c, h), ORC_NONE});
    }
    uint32_t end = (uint32_t)t.predictions.size();
    t.nodes[id].w.lo = start;
    t.nodes[id].w.hi = end;
    k.n_pred += na;
}

// next_action.rs:28-53 — first minimum of (n_t, c*) over active children, newest arc first
bool revisit_choice(const SearchTree &t, uint32_t pos, uint32_t &edge, uint32_t &n_t_as, orc_counters &k) {
    bool have = false;
    float best_c = 0.0f;
    for (uint32_t e = t.nodes[pos].next[0]; e != ORC_NONE; e = t.edges[e].next[0]) {
        ++k.d_sel;
        const StateWeight &cw = t.nodes[t.edges[e].node[1]].w;
        if (!cw.is_active()) continue;
        // min_by keeps the first of equal minima: replace only on strictly less
        bool less = !have || cw.n_t < n_t_as || (cw.n_t == n_t_as && cw.c_t_star < best_c);
        if (less) {
            have = true;
            edge = e;
            n_t_as = cw.n_t;
            best_c = cw.c_t_star;
        }
    }
    return have;
}

// next_action.rs:55-88
bool max_curiosity(const SearchTree &t, uint32_t pos, uint32_t &prediction_pos, orc_counters &k) {
    const StateWeight &w = t.nodes[pos].w;
    float c_s = w.c;
    std::vector<float> c_t_star_values;  // :57-61, newest arc first
    for (uint32_t e = t.nodes[pos].next[0]; e != ORC_NONE; e = t.edges[e].next[0])
        c_t_star_values.push_back(t.nodes[t.edges[e].node[1]].w.c_t_star);
    ++k.n_cur;
    k.n_cand += w.hi - w.lo;
    bool have = false;
    float best = 0.0f;
    for (uint32_t j = w.lo; j < w.hi; ++j) {
        const ActionPrediction &p = t.predictions[j];
        if (p.edge_id != ORC_NONE) continue;  // :68-71
        float c_theta_star = c_s - p.g_theta_sa;
        if (c_t_star_values.empty()) {
            // :73-75 min_by keeps the first minimum
            if (!have || c_theta_star < best) {
                have = true;
                best = c_theta_star;
                prediction_pos = j;
            }
        } else {
            // :78-86 left-fold f32 sum in list order; max_by keeps the last maximum
            float curiosity = 0.0f;
            for (float c_t_star : c_t_star_values) curiosity += std::sqrt(std::fabs(c_t_star - c_theta_star));
            if (!have || curiosity >= best) {
                have = true;
                best = curiosity;
                prediction_pos = j;
            }
        }
    }
    return have;
}

enum class Next { None, Visited, Unvisited };

// next_action.rs:11-26
Next next_action(const SearchTree &t, uint32_t pos, uint32_t n_as_tol, uint32_t &out, orc_counters &k) {
    if (!t.nodes[pos].w.is_active()) return Next::None;
    ++k.n_sel;
    uint32_t e = 0, n_t_as = 0;
    bool have_r = revisit_choice(t, pos, e, n_t_as, k);
    if (have_r && n_t_as < n_as_tol) {
        out = e;
        return Next::Visited;
    }
    uint32_t j = 0;
    if (max_curiosity(t, pos, j, k)) {
        out = j;
        return Next::Unvisited;
    }
    if (have_r) {
        out = e;
        return Next::Visited;
    }
    return Next::None;
}

// empty_transitions.rs:7-41: two ordered maps; pop the lowest node of the
// current level, upsert into the next level
struct Info {
    float c_t_star;
    uint32_t newly_exhausted_children;
};

// empty_transitions.rs:50-87 (old = false) and :89-127 (old = true)
void cascade(SearchTree &t, uint32_t edge_id, bool old, orc_counters &k) {
    const GEdge &a_t = t.edges[edge_id];
    const StateWeight &s_t = t.nodes[a_t.node[1]].w;
    uint32_t n_t_s_t = s_t.n_t;
    Info info;
    info.c_t_star = s_t.c_t_star;
    info.newly_exhausted_children = old ? (s_t.is_active() ? 0u : 1u) : 1u;
    std::map<uint32_t, Info> current, next;
    current.emplace(a_t.node[0], info);
    for (;;) {
        if (current.empty()) {
            std::swap(current, next);
            if (current.empty()) break;
        }
        auto it = current.begin();
        uint32_t child_index = it->first;
        Info ai = it->second;
        current.erase(it);
        ++k.n_cn;
        StateWeight &child = t.nodes[child_index].w;
        child.exhausted_children += ai.newly_exhausted_children;
        if (child.c_t_star > ai.c_t_star)
            child.c_t_star = ai.c_t_star;
        else
            child.n_t += 1;
        if (old) child.n_t = std::max(child.n_t, n_t_s_t);
        Info up;
        up.c_t_star = ai.c_t_star;
        up.newly_exhausted_children = child.is_active() ? 0u : 1u;
        for (uint32_t e = t.nodes[child_index].next[1]; e != ORC_NONE; e = t.edges[e].next[1]) {
            ++k.d_cn;
            uint32_t parent_id = t.edges[e].node[0];
            auto f = next.find(parent_id);
            if (f == next.end()) {
                next.emplace(parent_id, up);
            } else {
                f->second.c_t_star = std::min(f->second.c_t_star, up.c_t_star);
                f->second.newly_exhausted_children += up.newly_exhausted_children;
            }
        }
    }
}

inline void path_insert(ActionSet &p, uint32_t a) {  // path/set.rs:23-26
    p.insert(std::lower_bound(p.begin(), p.end(), (uint16_t)a), (uint16_t)a);
}

// tree/mod.rs:113-232
int roll_out_episodes(SearchTree &t, const Space &space, const State &root, State &state, Cost &cost, ActionSet &path,
                      uint32_t &state_pos, orc_counters &k) {
    for (;;) {
        uint32_t out = 0;
        Next na = next_action(t, state_pos, space.n_as_tol(path.size()), out, k);
        if (na == Next::Visited) {  // :139-159
            ++k.n_visit;
            uint32_t prediction_pos = t.edges[out].prediction_pos;
            uint32_t action_id = t.predictions[prediction_pos].a_id;
            path_insert(path, action_id);
            space_act(space.n, state, action_id);
            state_pos = t.edges[out].node[1];
        } else if (na == Next::Unvisited) {  // :160-218
            uint32_t prediction_pos = out;
            uint32_t action_id = t.predictions[prediction_pos].a_id;
            path_insert(path, action_id);
            ++k.n_probe;
            auto f = t.positions.find(path);
            if (f != t.positions.end()) {  // :172-179
                ++k.n_hit;
                ++k.n_arc;
                uint32_t arc = t.add_arc(state_pos, f->second, prediction_pos);
                cascade(t, arc, true, k);
                state = root;
                path.clear();
                state_pos = 0;
                ++k.n_reset;
            } else {  // :181-216
                space_act(space.n, state, action_id);
                bool bad = false;
                cost = space.cost(state, &bad);
                if (bad) return 1;
                float c_as = space.evaluate(cost);
                ++k.n_ins;
                uint32_t next_pos = t.add_node(path, StateWeight(c_as));
                ++k.n_arc;
                uint32_t arc = t.add_arc(state_pos, next_pos, prediction_pos);
                if (space_is_terminal(space.n, state)) {
                    ++k.n_term;
                    cascade(t, arc, false, k);
                    state = root;
                    path.clear();
                    state_pos = 0;
                    ++k.n_reset;
                } else {
                    state_pos = next_pos;
                    return 0;
                }
            }
        } else {  // :220-229
            if (path.empty()) return 0;
            return 2;  // the reference's unreachable!()
        }
    }
}

// ----------------------------------------------------------------------------
// synthetic inputs: counter-based, identical in the CUDA library
// ----------------------------------------------------------------------------
inline uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
inline uint32_t bounded(uint64_t r, uint32_t n) { return (uint32_t)(((r >> 32) * (uint64_t)n) >> 32); }

void generate_root(uint64_t seed, uint64_t root, uint32_t n, uint32_t k_min, uint32_t k_max, uint8_t *parents,
                   uint32_t *permitted) {
    uint32_t a_dim = action_dim(n), words = mask_words(n);
    uint64_t s = mix64(seed ^ mix64(root + 0x5851F42D4C957F2Dull));
    uint64_t ctr = 0;
    auto next = [&]() { return mix64(s + (ctr++) * 0xD1342543DE82EF95ull); };
    // rooted_tree/mod.rs:14-20: parents[0] = parents[1] = parents[N-1] = 0, parents[i] ~ U{0..i-1}
    for (uint32_t i = 0; i < n; ++i) parents[i] = 0;
    for (uint32_t i = 2; i + 1 < n; ++i) parents[i] = (uint8_t)bounded(next(), i);
    // 04-c21-tree.rs:110: k ~ U{k_min..=k_max}; modify_parent_once.rs:14-25: k actions without replacement
    uint32_t kk = k_min + bounded(next(), k_max - k_min + 1);
    std::vector<uint32_t> perm(a_dim);
    for (uint32_t i = 0; i < a_dim; ++i) perm[i] = i;
    for (uint32_t i = 0; i < words; ++i) permitted[i] = 0;
    for (uint32_t t = 0; t < kk; ++t) {
        uint32_t j = t + bounded(next(), a_dim - t);
        std::swap(perm[t], perm[j]);
        mask_set(permitted, perm[t]);
    }
}

inline float hash_prior(uint64_t seed, uint64_t root, uint64_t step, uint32_t a) {
    uint64_t r = mix64(mix64(seed ^ 0xA0761D6478BD642Full) + root * 0x9E3779B97F4A7C15ull + step * 0xE7037ED1A0B428DBull +
                       (uint64_t)a * 0x8EBC6AF09C88C6E3ull);
    return (float)(r >> 40) * 5.9604644775390625e-08f;  // 24 bits -> [0,1)
}

template <class F>
void parallel_for(uint32_t n, int n_threads, F f) {
    if (n_threads <= 1 || n < 2) {
        for (uint32_t i = 0; i < n; ++i) f(i, 0);
        return;
    }
    std::atomic<uint32_t> next{0};
    const uint32_t chunk = std::max<uint32_t>(1, n / (uint32_t)(n_threads * 16));
    std::vector<std::thread> ts;
    for (int t = 0; t < n_threads; ++t)
        ts.emplace_back([&, t]() {
            for (;;) {
                uint32_t b = next.fetch_add(chunk);
                if (b >= n) break;
                uint32_t e = std::min(n, b + chunk);
                for (uint32_t i = b; i < e; ++i) f(i, t);
            }
        });
    for (auto &t : ts) t.join();
}


// ----------------------------------------------------------------------------
// the c21 cost and the action kinds of a connected graph held as neighbourhood bit sets
// (simple_graph/connected_bitset_graph/mod.rs; B32: one u32 per vertex, N <= 32)
// ----------------------------------------------------------------------------

// connected_bitset_graph/mod.rs:45-71: breadth-first search from v in G - vu; a cut edge never reaches u
bool graph_is_cut_edge(const uint32_t *nbr, uint32_t v, uint32_t u) {
    uint32_t new_vertices = nbr[v] ^ (1u << u);  // :52-55 add_or_remove(u)
    uint32_t explored = 1u << v;                 // :56-57
    while (new_vertices) {                       // :58
        if (new_vertices >> u & 1u) return false;  // :59-61
        explored |= new_vertices;                // :62
        uint32_t recent = new_vertices;          // :63
        new_vertices = 0u;                       // :64
        for (uint32_t r = recent; r; r &= r - 1) new_vertices |= nbr[__builtin_ctz(r)];  // :65-67
        new_vertices &= ~explored;               // :68
    }
    return true;  // :70
}

// connected_bitset_graph/mod.rs:134-154 in the index space of AddOrDeleteEdge::action_index
// (bitset_graph/space/action.rs:10-19): bit colex(e) = Add(e) is available, bit C(N,2) + colex(e) = Delete(e) is
void graph_action_kinds(uint32_t n, const uint32_t *nbr, uint32_t *kinds) {
    const uint32_t e2 = n * (n - 1) / 2;
    for (uint32_t w = 0; w < (2 * e2 + 31) / 32; ++w) kinds[w] = 0u;
    for (uint32_t v = 0; v < n; ++v)
        for (uint32_t u = 0; u < v; ++u) {  // :139-140
            const uint32_t pos = colex_position(v, u);
            if (nbr[v] >> u & 1u) {                                          // :141
                if (!graph_is_cut_edge(nbr, v, u)) mask_set(kinds, e2 + pos);  // :143-147 (None for a cut edge)
            } else {
                mask_set(kinds, pos);  // :149
            }
        }
}

// connected_bitset_graph/mod.rs:226-317: depth-first branch and bound over (matched edges, unvisited vertices);
// only the size of the matching is ever read on this path, so the edge lists are kept as counts
uint32_t graph_matching_number(uint32_t n, const uint32_t *nbr) {
    if (n == 0) return 0;  // :235-237
    struct Search {
        uint32_t edges, unvisited;
    };
    std::vector<Search> queue;  // a VecDeque used as a stack: push_back / pop_back (:242-246)
    queue.push_back({0u, n >= 32 ? 0xffffffffu : (1u << n) - 1u});  // :238-243
    uint32_t matching_number = 0;                                   // :244
    while (!queue.empty()) {
        Search m = queue.back();
        queue.pop_back();
        uint32_t unvisited = m.unvisited;
        uint32_t max_future = (uint32_t)__builtin_popcount(unvisited) / 2;  // :261
        if (m.edges + max_future <= matching_number) continue;               // :262-264
        const uint32_t next_v = 31u - (uint32_t)__builtin_clz(unvisited);    // :269 max_unchecked
        unvisited ^= 1u << next_v;                                           // :274
        max_future = (uint32_t)__builtin_popcount(unvisited) / 2;            // :275
        if (m.edges + max_future > matching_number) queue.push_back({m.edges, unvisited});  // :279-285
        for (uint32_t nb = unvisited & nbr[next_v]; nb; nb &= nb - 1) {      // :286-288 ascending
            const uint32_t next_u = (uint32_t)__builtin_ctz(nb);
            const uint32_t new_edges = m.edges + 1;                          // :289-290
            const uint32_t new_unvisited = unvisited ^ (1u << next_u);       // :291-296
            if (new_edges > matching_number) matching_number = new_edges;    // :297-300
            max_future = (uint32_t)__builtin_popcount(new_unvisited) / 2;    // :302
            if (new_edges + max_future > matching_number) queue.push_back({new_edges, new_unvisited});  // :303-309
        }
    }
    return matching_number;
}

// connected_bitset_graph/mod.rs:200-216 + :324-333: the largest real part of the eigenvalues of A + 1e-4 I.  faer's
// general `eigenvalues()` is un-vendored; the matrix is symmetric, so its spectrum is real and any backward-stable
// symmetric solver gives the same lambda_1 to ~1e-15 (cyclic Jacobi here)
double graph_lambda1(uint32_t n, const uint32_t *nbr) {
    double a[MAXN][MAXN];
    for (uint32_t i = 0; i < n; ++i)
        for (uint32_t j = 0; j < n; ++j) a[i][j] = 0.0;
    for (uint32_t i = 0; i < n; ++i) a[i][i] = 0.0001;  // :205-209 "ZERO"
    for (uint32_t v = 0; v < n; ++v)
        for (uint32_t r = nbr[v]; r; r &= r - 1) a[v][__builtin_ctz(r)] = 1.0;  // :210-214
    return jacobi_max_eigenvalue(n, a);
}

}  // namespace

// ----------------------------------------------------------------------------
// optimizer/mod.rs:7-22
// ----------------------------------------------------------------------------
struct orc_optimizer {
    Space space;
    uint32_t batch = 0;
    int n_threads = 1;
    std::vector<State> roots, states;
    std::vector<Cost> costs;
    std::vector<ActionSet> paths;
    std::vector<uint32_t> last_positions;
    std::vector<SearchTree> trees;
    std::vector<uint32_t> num_inspected_nodes;
    State argmin_state;
    Cost argmin_cost;
    float argmin_eval = 0.0f;
    std::vector<orc_counters> thread_counters;
    std::vector<int> thread_err;

    void merge(orc_counters &dst) const {
        std::memset(&dst, 0, sizeof(dst));
        for (const auto &c : thread_counters) {
            const uint64_t *s = reinterpret_cast<const uint64_t *>(&c);
            uint64_t *d = reinterpret_cast<uint64_t *>(&dst);
            for (size_t i = 0; i < sizeof(orc_counters) / 8; ++i) d[i] += s[i];
        }
    }
};

extern "C" {

uint32_t orc_colex_position(uint32_t u, uint32_t v) { return colex_position(u, v); }
void orc_from_colex_position(uint32_t pos, uint32_t *mx, uint32_t *mn) { from_colex_position(pos, *mx, *mn); }
uint32_t orc_action_dim(uint32_t n) { return action_dim(n); }
uint32_t orc_c_upper(uint32_t n) { return c_upper_bound(n); }

int orc_cost(uint32_t n, const uint8_t *parents, int method, float c_lower, float c_upper, double *lambda1, uint32_t *mu,
             float *c) {
    Space sp;
    sp.n = n;
    sp.a_dim = action_dim(n);
    sp.words = mask_words(n);
    sp.c_lower = c_lower;
    sp.slope = 1.0f / (c_upper - c_lower);
    sp.lambda_method = method;
    State s;
    for (uint32_t i = 0; i < n; ++i) s.parents[i] = parents[i];
    bool bad = false;
    Cost k = sp.cost(s, &bad);
    *lambda1 = k.lambda_1;
    *mu = k.mu;
    *c = sp.evaluate(k);
    return bad ? 1 : 0;
}


/* ---- connected bitset graphs (connected_bitset_graph/mod.rs) ---- */
int orc_graph_is_cut_edge(uint32_t n, const uint32_t *nbr, uint32_t v, uint32_t u) {
    (void)n;
    return graph_is_cut_edge(nbr, v, u) ? 1 : 0;
}
void orc_graph_action_kinds(uint32_t n, const uint32_t *nbr, uint32_t *kinds) { graph_action_kinds(n, nbr, kinds); }
uint32_t orc_graph_matching_number(uint32_t n, const uint32_t *nbr) { return graph_matching_number(n, nbr); }
/* conjecture_2_1_cost (:319-337): returns 1 where the reference's `assert!(lambda_1 > 1.4)` would abort */
int orc_graph_cost(uint32_t n, const uint32_t *nbr, double *lambda1, uint32_t *mu) {
    *lambda1 = graph_lambda1(n, nbr);
    *mu = graph_matching_number(n, nbr);
    return *lambda1 > 1.4 ? 0 : 1;
}

uint32_t orc_matching_greedy(uint32_t n, const uint8_t *parents) { return matching_greedy(n, parents); }
uint32_t orc_matching_poly(uint32_t n, const uint8_t *parents) {
    uint32_t mu = 0;
    if (n <= 22) lambda1_poly(n, parents, &mu);
    return mu;
}

static State make_state(uint32_t n, const uint8_t *parents, const uint32_t *mask) {
    State s;
    for (uint32_t i = 0; i < n; ++i) s.parents[i] = parents[i];
    for (uint32_t i = 0; i < mask_words(n); ++i) s.permitted[i] = mask[i];
    return s;
}

uint32_t orc_action_data(uint32_t n, const uint8_t *parents, const uint32_t *mask, uint32_t *out) {
    State s = make_state(n, parents, mask);
    return space_action_data(n, s, out);
}

void orc_act(uint32_t n, uint8_t *parents, uint32_t *mask, uint32_t action) {
    State s = make_state(n, parents, mask);
    space_act(n, s, action);
    for (uint32_t i = 0; i < n; ++i) parents[i] = s.parents[i];
    for (uint32_t i = 0; i < mask_words(n); ++i) mask[i] = s.permitted[i];
}

void orc_write_vec(uint32_t n, const uint8_t *parents, const uint32_t *mask, float *vec) {
    State s = make_state(n, parents, mask);
    space_write_vec(n, s, vec);
}

void orc_generate_roots(uint64_t seed, uint64_t first_root, uint32_t count, uint32_t n, uint32_t k_min, uint32_t k_max,
                        uint8_t *parents, uint32_t *permitted_mask) {
    uint32_t words = mask_words(n);
    for (uint32_t i = 0; i < count; ++i)
        generate_root(seed, first_root + i, n, k_min, k_max, parents + (size_t)i * n, permitted_mask + (size_t)i * words);
}

void orc_hash_priors(uint64_t seed, uint64_t first_root, uint32_t count, uint32_t a_dim, uint64_t step, float *out) {
    for (uint32_t i = 0; i < count; ++i)
        for (uint32_t a = 0; a < a_dim; ++a) out[(size_t)i * a_dim + a] = hash_prior(seed, first_root + i, step, a);
}

// nabla/model/dfdx.rs:69-84 with the example's module stack (04-c21-tree.rs:46-52):
// three Linear+ReLU and one Linear+Sigmoid, f32, y = x W^T + b with W stored [out][in] as dfdx does.
void orc_mlp_forward(const float *params, const uint32_t *dims, uint32_t rows, const float *x, float *y, int n_threads) {
    parallel_for(rows, n_threads, [&](uint32_t r, int) {
        std::vector<float> cur(x + (size_t)r * dims[0], x + (size_t)(r + 1) * dims[0]), nxt;
        const float *p = params;
        for (int l = 0; l < 4; ++l) {
            uint32_t din = dims[l], dout = dims[l + 1];
            const float *w = p, *b = p + (size_t)din * dout;  // weight[out][in] (dfdx Linear), then bias[out]
            nxt.assign(dout, 0.0f);
            for (uint32_t o = 0; o < dout; ++o) {
                const float *wr = w + (size_t)o * din;
                float acc = 0.0f;
                for (uint32_t i = 0; i < din; ++i) acc += cur[i] * wr[i];
                nxt[o] = acc;
            }
            for (uint32_t o = 0; o < dout; ++o) {
                float v = nxt[o] + b[o];
                nxt[o] = l < 3 ? (v > 0.0f ? v : 0.0f) : 1.0f / (1.0f + std::exp(-v));
            }
            cur.swap(nxt);
            p += (size_t)din * dout + dout;
        }
        std::copy(cur.begin(), cur.end(), y + (size_t)r * dims[4]);
    });
}

orc_optimizer *orc_create(uint32_t n, uint32_t n_roots, float c_lower, float c_upper, const uint32_t *n_as_tol,
                          uint32_t n_as_tol_len, uint32_t n_as_tol_default, int lambda_method, int n_threads) {
    if (n < 5 || n > MAXN) return nullptr;
    auto *o = new orc_optimizer();
    o->space.n = n;
    o->space.a_dim = action_dim(n);
    o->space.words = mask_words(n);
    o->space.c_lower = c_lower;
    o->space.slope = 1.0f / (c_upper - c_lower);
    o->space.lambda_method = lambda_method;
    o->space.tol.assign(n_as_tol, n_as_tol + n_as_tol_len);
    o->space.tol_default = n_as_tol_default;
    o->batch = n_roots;
    o->n_threads = n_threads <= 0 ? (int)std::max(1u, std::thread::hardware_concurrency()) : n_threads;
    o->roots.resize(n_roots);
    o->states.resize(n_roots);
    o->costs.resize(n_roots);
    o->paths.resize(n_roots);
    o->last_positions.assign(n_roots, 0);
    o->trees.resize(n_roots);
    o->num_inspected_nodes.assign(n_roots, 0);
    o->thread_counters.assign(o->n_threads, orc_counters{});
    o->thread_err.assign(o->n_threads, 0);
    return o;
}

void orc_destroy(orc_optimizer *o) { delete o; }

void orc_set_roots(orc_optimizer *o, const uint8_t *parents, const uint32_t *mask) {
    uint32_t n = o->space.n, w = o->space.words;
    for (uint32_t i = 0; i < o->batch; ++i) o->roots[i] = make_state(n, parents + (size_t)i * n, mask + (size_t)i * w);
}

// optimizer/mod.rs:62-101 (and the tail of par_reset_trees :340-359)
static int init_trees_impl(orc_optimizer *o, const float *priors, bool first);
int orc_init_trees(orc_optimizer *o, const float *priors) { return init_trees_impl(o, priors, true); }
// the tail of par_reset_trees (optimizer/mod.rs:340-359): as above, but argmin_data is left alone
int orc_reinit_trees(orc_optimizer *o, const float *priors) { return init_trees_impl(o, priors, false); }
static int init_trees_impl(orc_optimizer *o, const float *priors, bool first) {
    const Space &sp = o->space;
    std::fill(o->thread_err.begin(), o->thread_err.end(), 0);
    parallel_for(o->batch, o->n_threads, [&](uint32_t i, int t) {
        o->states[i] = o->roots[i];
        bool bad = false;
        o->costs[i] = sp.cost(o->roots[i], &bad);
        if (bad) o->thread_err[t] = 1;
        o->paths[i].clear();
        o->last_positions[i] = 0;
        SearchTree &tr = o->trees[i];
        tr.clear();
        float c = sp.evaluate(o->costs[i]);
        uint32_t root_id = tr.add_node(ActionSet{}, StateWeight(c));
        add_actions(tr, root_id, sp, o->roots[i], priors + (size_t)i * sp.a_dim, o->thread_counters[t]);
        o->num_inspected_nodes[i] = 0;
    });
    if (!first) {
        for (int e : o->thread_err)
            if (e) return e;
        return 0;
    }
    // :95-101 min_by over roots; first minimum (lowest index) on ties
    uint32_t best = 0;
    float best_e = sp.evaluate(o->costs[0]);
    for (uint32_t i = 1; i < o->batch; ++i) {
        float e = sp.evaluate(o->costs[i]);
        if (e < best_e) {
            best_e = e;
            best = i;
        }
    }
    o->argmin_state = o->roots[best];
    o->argmin_cost = o->costs[best];
    o->argmin_eval = best_e;
    for (int e : o->thread_err)
        if (e) return e;
    return 0;
}

// par_reset_trees' first half (optimizer/mod.rs:284-339) with the example's modify_root (04-c21-tree.rs:172-206).
// The reference draws from an unseeded thread_rng; the draws here come from the counter generator keyed by
// (seed, epoch, global root index), and `n.choose(rng)` enumerates the candidates in node-index order (a uniform
// choice does not depend on the enumeration order).  Returns 6 where the reference would panic.  The caller then
// runs orc_init_trees (the shared tail :340-359).
int orc_modify_roots(orc_optimizer *o, uint64_t seed, uint64_t epoch, uint64_t first_root, uint32_t k_min, uint32_t k_max) {
    const Space &sp = o->space;
    const uint32_t n = sp.n, a_dim = sp.a_dim;
    int rc = 0;
    for (uint32_t i = 0; i < o->batch; ++i) {
        const SearchTree &t = o->trees[i];
        State &root = o->roots[i];
        uint64_t s = mix64(seed ^ mix64(first_root + i + 0x5851F42D4C957F2Dull) ^
                           mix64(epoch * 0xA24BAED4963EE407ull + 0x9FB21C651E98DF25ull));
        uint64_t ctr = 0;
        auto next = [&]() { return mix64(s + (ctr++) * 0xD1342543DE82EF95ull); };
        auto randomize = [&](uint32_t num) {  // modify_parent_once.rs:27-37 (choose_multiple as partial Fisher-Yates)
            std::vector<uint32_t> perm(a_dim);
            for (uint32_t q = 0; q < a_dim; ++q) perm[q] = q;
            root.permitted.fill(0);
            for (uint32_t q = 0; q < num; ++q) {
                uint32_t j = q + bounded(next(), a_dim - q);
                std::swap(perm[q], perm[j]);
                mask_set(root.permitted.data(), perm[q]);
            }
        };
        auto move_to = [&](uint32_t node) {  // p.actions_taken().for_each(|a| space.act(state, a))
            for (uint16_t a : t.node_keys[node]) space_act(n, root, a);
        };
        const float c_root = t.nodes[0].w.c, c_star = t.nodes[0].w.c_t_star;  // n[0]: the empty path sorts first
        uint32_t kcur = 0;
        for (uint32_t w = 0; w < sp.words; ++w) kcur += (uint32_t)__builtin_popcount(root.permitted[w]);
        std::vector<uint32_t> cands;
        if (c_root == c_star) {
            if (kcur < k_min || kcur > k_max) {  // unreachable!()
                rc = 6;
                continue;
            }
            if (kcur == k_max) {
                uint32_t num = k_min + bounded(next(), k_max - k_min + 1);
                for (uint32_t v = 0; v < n; ++v) root.parents[v] = 0;
                for (uint32_t v = 2; v + 1 < n; ++v) root.parents[v] = (uint8_t)bounded(next(), v);
                randomize(num);
            } else {
                for (uint32_t q = 0; q < t.nodes.size(); ++q)
                    if (t.nodes[q].w.c == c_root) cands.push_back(q);
                move_to(cands[bounded(next(), (uint32_t)cands.size())]);
                randomize(kcur + bounded(next(), k_max - kcur + 1));
            }
        } else {
            const float thr = (c_root + 3.0f * c_star) / 4.0f;
            for (uint32_t q = 0; q < t.nodes.size(); ++q)
                if (t.nodes[q].w.c <= thr) cands.push_back(q);
            if (cands.empty()) {
                rc = 6;
                continue;
            }
            move_to(cands[bounded(next(), (uint32_t)cands.size())]);
            randomize(k_min + bounded(next(), k_max - k_min + 1));
        }
    }
    return rc;
}

void orc_get_roots(orc_optimizer *o, uint8_t *parents, uint32_t *mask) {
    uint32_t n = o->space.n, w = o->space.words;
    for (uint32_t i = 0; i < o->batch; ++i) {
        std::memcpy(parents + (size_t)i * n, o->roots[i].parents.data(), n);
        std::memcpy(mask + (size_t)i * w, o->roots[i].permitted.data(), (size_t)w * 4);
    }
}

void orc_root_vecs(orc_optimizer *o, float *state_vecs) {
    const Space &sp = o->space;
    parallel_for(o->batch, o->n_threads,
                 [&](uint32_t i, int) { space_write_vec(sp.n, o->roots[i], state_vecs + (size_t)i * 2 * sp.a_dim); });
}

// optimizer/mod.rs:159-174
int orc_rollout(orc_optimizer *o, float *state_vecs) {
    const Space &sp = o->space;
    std::fill(o->thread_err.begin(), o->thread_err.end(), 0);
    parallel_for(o->batch, o->n_threads, [&](uint32_t i, int t) {
        orc_counters &k = o->thread_counters[t];
        int rc = roll_out_episodes(o->trees[i], sp, o->roots[i], o->states[i], o->costs[i], o->paths[i],
                                   o->last_positions[i], k);
        if (rc) o->thread_err[t] = rc;
        if (!o->paths[i].empty()) {
            ++k.n_live;
            if (state_vecs) space_write_vec(sp.n, o->states[i], state_vecs + (size_t)i * 2 * sp.a_dim);
        } else {
            ++k.n_noop;
        }
    });
    for (int e : o->thread_err)
        if (e) return e;
    return 0;
}

// optimizer/mod.rs:177-190 then :194-246
int orc_add_actions(orc_optimizer *o, const float *priors, int *improved) {
    const Space &sp = o->space;
    parallel_for(o->batch, o->n_threads, [&](uint32_t i, int t) {
        if (!o->paths[i].empty())
            add_actions(o->trees[i], o->last_positions[i], sp, o->states[i], priors + (size_t)i * sp.a_dim,
                        o->thread_counters[t]);
    });
    // par_update_argmmim_data: per tree, first minimum of c over nodes added since the last look that beat the
    // current best; across trees the lowest tree index wins ties (rayon leaves this unspecified)
    float min_eval = o->argmin_eval;
    bool have = false;
    uint32_t best_tree = 0, best_node = 0;
    float best_c = 0.0f;
    for (uint32_t i = 0; i < o->batch; ++i) {
        const SearchTree &t = o->trees[i];
        uint32_t num = o->num_inspected_nodes[i];
        if (num < t.nodes.size()) {
            bool have_t = false;
            uint32_t node_t = 0;
            float c_t = 0.0f;
            for (uint32_t j = num; j < t.nodes.size(); ++j) {
                float c = t.nodes[j].w.c;
                if (c < min_eval && (!have_t || c < c_t)) {
                    have_t = true;
                    c_t = c;
                    node_t = j;
                }
            }
            o->num_inspected_nodes[i] = (uint32_t)t.nodes.size();
            if (have_t && (!have || c_t < best_c)) {
                have = true;
                best_c = c_t;
                best_tree = i;
                best_node = node_t;
            }
        }
    }
    if (improved) *improved = have ? 1 : 0;
    if (have) {
        // :224-241 rebuild the state by replaying the node's action set (ascending) from its tree's root
        State st = o->roots[best_tree];
        for (uint16_t a : o->trees[best_tree].node_keys[best_node]) space_act(sp.n, st, a);
        o->argmin_state = st;
        o->argmin_cost = sp.cost(st);
        o->argmin_eval = sp.evaluate(o->argmin_cost);
    }
    return 0;
}

int orc_steps_hash(orc_optimizer *o, uint64_t seed, uint64_t first_root, uint64_t step0, uint32_t n_steps,
                   uint32_t *improved_steps, uint32_t cap, uint32_t *n_improved) {
    const Space &sp = o->space;
    std::vector<float> priors((size_t)o->batch * sp.a_dim);
    uint32_t ni = 0;
    for (uint32_t s = 0; s < n_steps; ++s) {
        int rc = orc_rollout(o, nullptr);
        if (rc) return rc;
        parallel_for(o->batch, o->n_threads, [&](uint32_t i, int) {
            for (uint32_t a = 0; a < sp.a_dim; ++a)
                priors[(size_t)i * sp.a_dim + a] = hash_prior(seed, first_root + i, step0 + s, a);
        });
        int imp = 0;
        orc_add_actions(o, priors.data(), &imp);
        if (imp) {
            if (improved_steps && ni < cap) improved_steps[ni] = (uint32_t)(step0 + s);
            ++ni;
        }
    }
    if (n_improved) *n_improved = ni;
    return 0;
}

void orc_get_counters(orc_optimizer *o, orc_counters *out) { o->merge(*out); }
void orc_reset_counters(orc_optimizer *o) { std::fill(o->thread_counters.begin(), o->thread_counters.end(), orc_counters{}); }

void orc_get_argmin(orc_optimizer *o, uint8_t *parents, uint32_t *mask, double *lambda1, uint32_t *mu, float *eval) {
    for (uint32_t i = 0; i < o->space.n; ++i) parents[i] = o->argmin_state.parents[i];
    for (uint32_t i = 0; i < o->space.words; ++i) mask[i] = o->argmin_state.permitted[i];
    *lambda1 = o->argmin_cost.lambda_1;
    *mu = o->argmin_cost.mu;
    *eval = o->argmin_eval;
}

void orc_get_walkers(orc_optimizer *o, uint8_t *parents, uint32_t *mask, uint32_t *path_mask, uint32_t *pos,
                     uint32_t *path_len) {
    uint32_t n = o->space.n, w = o->space.words;
    for (uint32_t i = 0; i < o->batch; ++i) {
        for (uint32_t v = 0; v < n; ++v) parents[(size_t)i * n + v] = o->states[i].parents[v];
        for (uint32_t k = 0; k < w; ++k) {
            mask[(size_t)i * w + k] = o->states[i].permitted[k];
            path_mask[(size_t)i * w + k] = 0;
        }
        for (uint16_t a : o->paths[i]) mask_set(path_mask + (size_t)i * w, a);
        pos[i] = o->last_positions[i];
        path_len[i] = (uint32_t)o->paths[i].size();
    }
}

void orc_tree_sizes(orc_optimizer *o, uint32_t tree, uint32_t *n_nodes, uint32_t *n_arcs, uint32_t *n_preds) {
    const SearchTree &t = o->trees[tree];
    *n_nodes = (uint32_t)t.nodes.size();
    *n_arcs = (uint32_t)t.edges.size();
    *n_preds = (uint32_t)t.predictions.size();
}

void orc_dump_tree(orc_optimizer *o, uint32_t tree, uint32_t *nodes, uint32_t *keys, uint32_t *preds, uint32_t *arcs) {
    const SearchTree &t = o->trees[tree];
    uint32_t w = o->space.words;
    auto bits = [](float f) {
        uint32_t u;
        std::memcpy(&u, &f, 4);
        return u;
    };
    for (size_t i = 0; i < t.nodes.size(); ++i) {
        const StateWeight &s = t.nodes[i].w;
        uint32_t *r = nodes + i * 6;
        r[0] = bits(s.c);
        r[1] = bits(s.c_t_star);
        r[2] = s.n_t;
        r[3] = s.exhausted_children;
        r[4] = s.lo;
        r[5] = s.hi;
        for (uint32_t k = 0; k < w; ++k) keys[i * w + k] = 0;
        for (uint16_t a : t.node_keys[i]) mask_set(keys + i * w, a);
    }
    for (size_t j = 0; j < t.predictions.size(); ++j) {
        preds[j * 3 + 0] = t.predictions[j].a_id;
        preds[j * 3 + 1] = bits(t.predictions[j].g_theta_sa);
        preds[j * 3 + 2] = t.predictions[j].edge_id;
    }
    for (size_t e = 0; e < t.edges.size(); ++e) {
        arcs[e * 3 + 0] = t.edges[e].node[0];
        arcs[e * 3 + 1] = t.edges[e].node[1];
        arcs[e * 3 + 2] = t.edges[e].prediction_pos;
    }
}

// optimizer/mod.rs:262-278 + tree/mod.rs:242-264
void orc_write_observations(orc_optimizer *o, uint32_t n_obs_tol, float *state_vecs, float *observations, float *weights) {
    const Space &sp = o->space;
    std::fill(observations, observations + (size_t)o->batch * sp.a_dim, 0.0f);
    std::fill(weights, weights + (size_t)o->batch * sp.a_dim, 0.0f);
    parallel_for(o->batch, o->n_threads, [&](uint32_t i, int) {
        if (state_vecs) space_write_vec(sp.n, o->roots[i], state_vecs + (size_t)i * 2 * sp.a_dim);
        const SearchTree &t = o->trees[i];
        float c_s = t.nodes[0].w.c;
        for (uint32_t e = t.nodes[0].next[0]; e != ORC_NONE; e = t.edges[e].next[0]) {
            const StateWeight &cw = t.nodes[t.edges[e].node[1]].w;
            if (!cw.is_active() || cw.n_t >= n_obs_tol) {
                float h = sp.h_sa(c_s, cw.c_t_star);
                uint32_t a = t.predictions[t.edges[e].prediction_pos].a_id;
                observations[(size_t)i * sp.a_dim + a] = h;
                weights[(size_t)i * sp.a_dim + a] = 1.0f;
            }
        }
    });
}

}  // extern "C"
