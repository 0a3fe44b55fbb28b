// Can ONE launch be cooperative (all CTAs co-resident, grid.sync legal) AND clustered (CTA pairs on one TPC)?
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(1024, 1) k(unsigned *out) {
    extern __shared__ unsigned char smem[];
    cg::grid_group g = cg::this_grid();
    unsigned rank, smid, ncta;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(ncta));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) {
        smem[0] = 1;
        out[blockIdx.x * 4 + 0] = rank;
        out[blockIdx.x * 4 + 1] = smid;
        out[blockIdx.x * 4 + 2] = ncta;
    }
    g.sync();
    if (threadIdx.x == 0) out[blockIdx.x * 4 + 3] = out[((blockIdx.x + 1) % gridDim.x) * 4 + 1];
}
int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int grid = p.multiProcessorCount, smem = 187 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    unsigned *out;
    cudaMalloc(&out, grid * 16);
    cudaMemset(out, 0xff, grid * 16);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(1024);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    at[1].id = cudaLaunchAttributeClusterDimension;
    at[1].val.clusterDim.x = 2;
    at[1].val.clusterDim.y = 1;
    at[1].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    int ncl = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, k, &cfg);
    printf("SMs %d, max active clusters of 2: %d (%s)\n", grid, ncl, cudaGetErrorString(e));
    e = cudaLaunchKernelEx(&cfg, k, out);
    printf("launch: %s\n", cudaGetErrorString(e));
    e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    unsigned *h = new unsigned[grid * 4];
    cudaMemcpy(h, out, grid * 16, cudaMemcpyDeviceToHost);
    int pairs_ok = 0;
    for (int i = 0; i + 1 < grid; i += 2) pairs_ok += (h[i * 4] == 0 && h[i * 4 + 4] == 1 && (h[i * 4 + 1] ^ h[i * 4 + 5]) == 1);
    printf("cta 0: rank %u smid %u ncta %u; cta 1: rank %u smid %u; pairs on adjacent SMs: %d of %d\n", h[0], h[1], h[2], h[4], h[5], pairs_ok, grid / 2);
    for (int i = 0; i < 16; ++i) printf("%u:%u ", h[i * 4], h[i * 4 + 1]);
    printf("\n");
    return 0;
}
