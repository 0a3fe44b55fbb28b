// micro-benchmarks: dependent-chain latencies on B200 (one warp), cycles per op
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *cyc, double x, float xf, int n) {
    __shared__ unsigned smem[64];
    smem[threadIdx.x] = threadIdx.x ^ 1;
    __syncwarp();
    double a = x + threadIdx.x, b = x * 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) a = __fma_rn(a, x, b);
    long long t1 = clock64();
    for (int i = 0; i < n; ++i) a = __dmul_rn(a, x);
    long long t2 = clock64();
    float f = xf + threadIdx.x;
    for (int i = 0; i < n; ++i) f = __fmaf_rn(f, xf, 0.5f);
    long long t3 = clock64();
    unsigned idx = threadIdx.x;
    for (int i = 0; i < n; ++i) idx = smem[idx & 31];
    long long t4 = clock64();
    // two independent DFMA chains (ILP 2)
    double c = x - threadIdx.x, d = x * 0.25;
    for (int i = 0; i < n; ++i) { a = __fma_rn(a, x, b); c = __fma_rn(c, x, d); }
    long long t5 = clock64();
    // dependent chain fma -> compare -> select (like the sign test)
    bool ok = true;
    for (int i = 0; i < n; ++i) { a = __fma_rn(a, x, b); ok = ok && (a > 0.0); }
    long long t6 = clock64();
    unsigned v = threadIdx.x;
    for (int i = 0; i < n; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1;
    long long t7 = clock64();
    float g = xf;
    for (int i = 0; i < n; ++i) g = __fsqrt_rn(g + 1.0f);
    long long t8 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0; cyc[1] = t2 - t1; cyc[2] = t3 - t2; cyc[3] = t4 - t3; cyc[4] = t5 - t4; cyc[5] = t6 - t5; cyc[6] = t7 - t6; cyc[7] = t8 - t7;
    }
    out[threadIdx.x] = a + f + idx + c + (ok ? 1 : 0) + v + g;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, 32 * 8); cudaMallocManaged(&cyc, 8 * 8);
    int n = 4096;
    k<<<1, 32>>>(out, cyc, 1.0000001, 1.0000001f, n);
    cudaDeviceSynchronize();
    k<<<1, 32>>>(out, cyc, 1.0000001, 1.0000001f, n);
    cudaDeviceSynchronize();
    const char *nm[] = {"DFMA chain", "DMUL chain", "FFMA chain", "LDS chain", "2x DFMA ILP (per pair)", "DFMA+DSETP chain", "SHFL+IADD chain", "FSQRT+FADD chain"};
    for (int i = 0; i < 8; ++i) printf("%-26s %.1f cycles/iter\n", nm[i], (double)cyc[i] / n);
    return 0;
}
