// micro-benchmark: cycles per op of three formulations of the lambda_1 stack machine (single warp, DEPTH 3)
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
enum { OP_END = 0, OP_L0 = 1, OP_L = 2, OP_T = 3, OP_PUSH = 4, OP_POPF = 5 };

__device__ __forceinline__ bool run_ifchain(const unsigned *prog, double x) {
    double Q0 = 0., S0 = 0., Q1 = 0., S1 = 0., Q2 = 0., S2 = 0.;
    bool ok = true;
    unsigned word = 0;
#pragma unroll 1
    for (unsigned i = 0;; ++i) {
        const unsigned r = i % 10u;
        if (r == 0u) word = prog[i / 10u];
        const unsigned op = word & 7u;
        word >>= 3;
        if (op == OP_L) { S0 = __fma_rn(S0, x, Q0); Q0 = __dmul_rn(Q0, x); }
        else if (op == OP_L0) { Q0 = x; S0 = 1.0; }
        else if (op == OP_T) { const double P = __fma_rn(x, Q0, -S0); ok = ok && (P > 0.0); S0 = Q0; Q0 = P; }
        else if (op == OP_PUSH) { Q2 = Q1; S2 = S1; Q1 = Q0; S1 = S0; }
        else if (op == OP_POPF) {
            const double P = __fma_rn(x, Q0, -S0); ok = ok && (P > 0.0);
            const double t = __dmul_rn(Q1, Q0); S0 = __fma_rn(S1, P, t); Q0 = __dmul_rn(Q1, P); Q1 = Q2; S1 = S2;
        } else { const double P = __fma_rn(x, Q0, -S0); ok = ok && (P > 0.0); break; }
    }
    return ok;
}
__device__ __forceinline__ bool run_switch(const unsigned *prog, double x) {
    double Q0 = 0., S0 = 0., Q1 = 0., S1 = 0., Q2 = 0., S2 = 0.;
    bool ok = true;
    unsigned word = 0;
#pragma unroll 1
    for (unsigned i = 0;; ++i) {
        if (i % 10u == 0u) word = prog[i / 10u];
        const unsigned op = word & 7u;
        word >>= 3;
        bool done = false;
        switch (op) {
            case OP_L: S0 = __fma_rn(S0, x, Q0); Q0 = __dmul_rn(Q0, x); break;
            case OP_L0: Q0 = x; S0 = 1.0; break;
            case OP_T: { const double P = __fma_rn(x, Q0, -S0); ok = ok && (P > 0.0); S0 = Q0; Q0 = P; break; }
            case OP_PUSH: Q2 = Q1; S2 = S1; Q1 = Q0; S1 = S0; break;
            case OP_POPF: { const double P = __fma_rn(x, Q0, -S0); ok = ok && (P > 0.0);
                const double t = __dmul_rn(Q1, Q0); S0 = __fma_rn(S1, P, t); Q0 = __dmul_rn(Q1, P); Q1 = Q2; S1 = S2; break; }
            default: { const double P = __fma_rn(x, Q0, -S0); ok = ok && (P > 0.0); done = true; }
        }
        if (done) break;
    }
    return ok;
}
// branch-free: ops re-encoded 2 bits: 0 L, 1 T, 2 PUSH, 3 POPF ; count known; fresh slot = (1,0)
__device__ __forceinline__ bool run_select(const unsigned *prog2, unsigned nops, double x) {
    double Q0 = 1., S0 = 0., Q1 = 1., S1 = 0., Q2 = 1., S2 = 0.;
    bool ok = true;
    unsigned word = 0;
#pragma unroll 1
    for (unsigned i = 0; i < nops; ++i) {
        if ((i & 15u) == 0u) word = prog2[i >> 4];
        const unsigned op = word & 3u;
        word >>= 2;
        const bool isL = op == 0u, isT = op == 1u, isPush = op == 2u, isPop = op == 3u;
        const double P = __fma_rn(x, Q0, -S0);
        const double bq = isT ? 1.0 : Q1, bs = isT ? 0.0 : S1;   // T == POPF against a fresh slot, without popping
        const double A = isL ? S0 : bs, B = isL ? x : P;
        const double C = isL ? Q0 : __dmul_rn(bq, Q0), D = isL ? Q0 : bq;
        const double Sn = __fma_rn(A, B, C), Qn = __dmul_rn(D, B);
        ok = ok && (isL || isPush || (P > 0.0));
        const double nQ1 = isPush ? Q0 : (isPop ? Q2 : Q1), nS1 = isPush ? S0 : (isPop ? S2 : S1);
        Q2 = isPush ? Q1 : Q2; S2 = isPush ? S1 : S2;
        Q1 = nQ1; S1 = nS1;
        Q0 = isPush ? 1.0 : Qn; S0 = isPush ? 0.0 : Sn;
    }
    const double P = __fma_rn(x, Q0, -S0);
    return ok && (P > 0.0);
}
__global__ void k(const unsigned *prog, const unsigned *prog2, unsigned nops, double x0, long long *cyc, int *out, int reps) {
    __shared__ unsigned sp[16], sp2[16];
    if (threadIdx.x < 16) { sp[threadIdx.x] = prog[threadIdx.x]; sp2[threadIdx.x] = prog2[threadIdx.x]; }
    __syncwarp();
    double x = x0 + threadIdx.x * 1e-3;
    int acc = 0;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) acc += run_ifchain(sp, x + r * 1e-6);
    long long t1 = clock64();
    for (int r = 0; r < reps; ++r) acc += run_switch(sp, x + r * 1e-6);
    long long t2 = clock64();
    for (int r = 0; r < reps; ++r) acc += run_select(sp2, nops, x + r * 1e-6);
    long long t3 = clock64();
    if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; }
    out[threadIdx.x] = acc;
}
int main() {
    // a plausible N=19 program: ops for a random-ish tree (27 ops incl. END)
    int ops[] = {OP_L0, OP_L, OP_T, OP_L, OP_PUSH, OP_L0, OP_L, OP_POPF, OP_T, OP_L, OP_L, OP_PUSH, OP_L0, OP_T, OP_L, OP_POPF, OP_L,
                 OP_T, OP_L, OP_PUSH, OP_L0, OP_L, OP_L, OP_POPF, OP_L, OP_L, OP_END};
    int n = sizeof(ops) / sizeof(int);
    std::vector<unsigned> p(16, 0), p2(16, 0);
    for (int i = 0; i < n; ++i) p[i / 10] |= (unsigned)ops[i] << (3 * (i % 10));
    int n2 = 0;
    for (int i = 0; i < n - 1; ++i) {
        int o = ops[i];
        unsigned c = (o == OP_L || o == OP_L0) ? 0u : (o == OP_T ? 1u : (o == OP_PUSH ? 2u : 3u));
        p2[n2 / 16] |= c << (2 * (n2 % 16));
        ++n2;
    }
    unsigned *dp, *dp2; long long *cyc; int *out;
    cudaMalloc(&dp, 64); cudaMalloc(&dp2, 64); cudaMallocManaged(&cyc, 64); cudaMalloc(&out, 128);
    cudaMemcpy(dp, p.data(), 64, cudaMemcpyHostToDevice); cudaMemcpy(dp2, p2.data(), 64, cudaMemcpyHostToDevice);
    int reps = 2000;
    for (int it = 0; it < 2; ++it) { k<<<1, 32>>>(dp, dp2, n2, 4.3, cyc, out, reps); cudaDeviceSynchronize(); }
    printf("ops per program: %d\n", n);
    const char *nm[] = {"if-chain", "switch", "branch-free select"};
    for (int i = 0; i < 3; ++i) printf("%-20s %.1f cycles/program  %.1f cycles/op\n", nm[i], (double)cyc[i] / reps, (double)cyc[i] / reps / n);
    return 0;
}
