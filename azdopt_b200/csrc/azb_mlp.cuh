// azb_mlp.cuh — the prior model forward: ActionModel::write_predictions (az-discrete-opt/src/nabla/model/dfdx.rs:
// 69-84) for the example's stack (graph-state/examples/04-c21-tree.rs:46-52): three Linear+ReLU, one Linear+Sigmoid.
// Parameters in dfdx order: weight[out][in] (K contiguous) then bias[out], f32.
//
// This file holds the fp32 CUDA-core path (AZB_MLP_FP32): the arithmetic cuBLAS sgemm gives the reference.
// One kernel per layer, bias + activation fused into the epilogue.
#pragma once
#include "azb_common.cuh"

enum { AZB_ACT_RELU = 0, AZB_ACT_SIGMOID = 1 };

// Y[M][ldy] = act(X[M][ldx] * W[Nout][K]^T + b).  64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread.
template <int ACT>
__global__ void __launch_bounds__(256) azb_linear_fp32_kernel(const float *__restrict__ X, uint32_t ldx,
                                                              const float *__restrict__ Wt, const float *__restrict__ bias,
                                                              float *__restrict__ Y, uint32_t ldy, uint32_t M, uint32_t K,
                                                              uint32_t Nout) {
    __shared__ float sx[16][64 + 4];
    __shared__ float sw[16][64 + 4];
    const uint32_t m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const uint32_t tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    // each thread loads 4 elements of each tile: row r = tid/4 (0..63), k4 = (tid%4)*4
    const uint32_t lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 4;
    for (uint32_t k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t k = k0 + lk + q;
            float xv = 0.f, wv = 0.f;
            if (k < K) {
                if (m0 + lr < M) xv = X[(size_t)(m0 + lr) * ldx + k];
                if (n0 + lr < Nout) wv = Wt[(size_t)(n0 + lr) * K + k];
            }
            sx[lk + q][lr] = xv;
            sw[lk + q][lr] = wv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sx[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sw[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t n = n0 + tx * 4 + j;
            if (n >= Nout) continue;
            float v = acc[i][j] + bias[n];
            if (ACT == AZB_ACT_RELU)
                v = v > 0.f ? v : 0.f;
            else
                v = 1.0f / (1.0f + expf(-v));
            Y[(size_t)m * ldy + n] = v;
        }
    }
}
