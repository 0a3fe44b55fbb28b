// FP64 pipe throughput on B200 vs the number of active lanes: cycles per DFMA warp-instruction with 32 warps per SM,
// 8 independent accumulators per thread (throughput bound, not latency bound).
// build: nvcc -O3 -cudart shared -gencode arch=compute_100a,code=sm_100a -o tools/ubench/fp64 tools/ubench/fp64.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, int active, int iters, long long *cyc) {
    const int lane = threadIdx.x & 31;
    double a0 = threadIdx.x, a1 = 1, a2 = 2, a3 = 3, a4 = 4, a5 = 5, a6 = 6, a7 = 7;
    const double x = 1.0000001, y = 1e-9;
    __syncthreads();
    const long long t0 = clock64();
    if (lane < active) {
        for (int i = 0; i < iters; ++i) {
            a0 = fma(a0, x, y); a1 = fma(a1, x, y); a2 = fma(a2, x, y); a3 = fma(a3, x, y);
            a4 = fma(a4, x, y); a5 = fma(a5, x, y); a6 = fma(a6, x, y); a7 = fma(a7, x, y);
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    double *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int warps : {4, 32})
        for (int active : {32, 16, 8, 4, 2, 1}) {
            k<<<148, warps * 32>>>(out, active, iters, cyc);
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            // warp-instructions per SMSP = warps/4 * iters * 8
            printf("warps/SM %2d active lanes %2d: %.2f cycles per DFMA warp-instruction per SMSP\n", warps, active,
                   (double)h / ((double)warps / 4 * iters * 8));
        }
    return 0;
}
