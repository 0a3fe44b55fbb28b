// azb_train.cuh — the epoch-boundary training step: NablaModel::update_model for ActionModel
// (az-discrete-opt/src/nabla/model/dfdx.rs:86-131) with the example's AdamConfig
// (graph-state/examples/04-c21-tree.rs:87-92: lr 1e-4, betas .9/.999, eps 1e-8, WeightDecay::L2(1e-6)).
//
//   w_n   = action_weights / sum(action_weights)                      (dfdx.rs:106-110)
//   loss  = sum_rows sum_actions (forward(states) - observations)^2 * w_n   (dfdx.rs:119-124)
//   grads = d loss / d params;  Adam step;  gradients zeroed          (dfdx.rs:127-129)
//
// Everything is f32 like the reference (dfdx on Cuda = cuBLAS sgemm + elementwise kernels).  The gradient of all
// four layers lives in ONE contiguous buffer in parameter order, so a sharded run all-reduces it with a single NCCL
// call before the Adam step; the loss and the weight sum are reduced in a fixed order (no float
// atomics), so a step is reproducible.
#pragma once
#include "azb_common.cuh"

// C[M][N] (ldc) = sum_k A(m,k) B(k,n), 64x64 tile, BK = 16, 256 threads, 4x4 outputs per thread.
//   A_KC: A is stored [M][K] (k contiguous), else [K][M] (m contiguous); lda is the row pitch of the stored matrix.
//   B_NC: B is stored [K][N] (n contiguous), else [N][K] (k contiguous).
//   RELU_MASK: multiply C by (mask[m][n] > 0)  — back-propagation through the ReLU whose OUTPUT is `mask`.
template <bool A_KC, bool B_NC, bool RELU_MASK>
__global__ void __launch_bounds__(256) azb_gemm_fp32_kernel(const float *__restrict__ A, uint32_t lda,
                                                            const float *__restrict__ Bm, uint32_t ldb,
                                                            float *__restrict__ C, uint32_t ldc, uint32_t M, uint32_t N,
                                                            uint32_t K, const float *__restrict__ mask, uint32_t ldm) {
    __shared__ float sa[16][64 + 4];
    __shared__ float sb[16][64 + 4];
    const uint32_t m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const uint32_t tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (uint32_t k0 = 0; k0 < K; k0 += 16) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t e = threadIdx.x + 256u * q;  // 1024 elements per tile
            float av = 0.f, bv = 0.f;
            if (A_KC) {
                const uint32_t kk = e & 15u, mm = e >> 4;
                if (m0 + mm < M && k0 + kk < K) av = A[(size_t)(m0 + mm) * lda + k0 + kk];
                sa[kk][mm] = av;
            } else {
                const uint32_t mm = e & 63u, kk = e >> 6;
                if (m0 + mm < M && k0 + kk < K) av = A[(size_t)(k0 + kk) * lda + m0 + mm];
                sa[kk][mm] = av;
            }
            if (B_NC) {
                const uint32_t nn = e & 63u, kk = e >> 6;
                if (n0 + nn < N && k0 + kk < K) bv = Bm[(size_t)(k0 + kk) * ldb + n0 + nn];
                sb[kk][nn] = bv;
            } else {
                const uint32_t kk = e & 15u, nn = e >> 4;
                if (n0 + nn < N && k0 + kk < K) bv = Bm[(size_t)(n0 + nn) * ldb + k0 + kk];
                sb[kk][nn] = bv;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sa[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sb[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t m = m0 + ty * 4 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t n = n0 + tx * 4 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (RELU_MASK) v = mask[(size_t)m * ldm + n] > 0.f ? v : 0.f;
            C[(size_t)m * ldc + n] = v;
        }
    }
}

// fixed-order sum of `n` floats into *out (as double and float): per-block partials, then one block
__global__ void __launch_bounds__(256) azb_sum_partial_kernel(const float *__restrict__ x, size_t n,
                                                              double *__restrict__ part) {
    __shared__ double red[256];
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) s += (double)x[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void __launch_bounds__(256) azb_sum_final_kernel(const double *__restrict__ part, uint32_t n,
                                                            double *__restrict__ out) {
    __shared__ double red[256];
    double s = 0.0;
    for (uint32_t i = threadIdx.x; i < n; i += 256) s += part[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = red[0];
}

// dZ4 = dloss/dp * sigmoid'(z) with p = sigmoid(z):  2 (p - o) (w / wsum) p (1 - p);  per-block loss partials.
// `scal[0]` = the (global) weight sum as f64; the division w / wsum is the reference's f32 elementwise division.
__global__ void __launch_bounds__(256) azb_loss_grad_kernel(const float *__restrict__ P, const float *__restrict__ O,
                                                            const float *__restrict__ Wt, size_t n,
                                                            const double *__restrict__ scal, float *__restrict__ dZ,
                                                            double *__restrict__ part) {
    __shared__ double red[256];
    const float wsum = (float)scal[0];
    double s = 0.0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float p = P[i], wn = __fdiv_rn(Wt[i], wsum);
        const float d = __fsub_rn(p, O[i]);
        const float e = __fmul_rn(__fmul_rn(d, d), wn);
        s += (double)e;
        const float gp = __fmul_rn(__fmul_rn(2.0f, d), wn);
        dZ[i] = __fmul_rn(gp, __fmul_rn(p, __fsub_rn(1.0f, p)));
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int d = 128; d >= 1; d >>= 1) {
        if ((int)threadIdx.x < d) red[threadIdx.x] += red[threadIdx.x + d];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

// db[n] = sum_rows dZ[r][n]: one block per 32 columns, 8 row groups, fixed order
__global__ void __launch_bounds__(256) azb_colsum_kernel(const float *__restrict__ dZ, uint32_t ld, uint32_t rows,
                                                         uint32_t N, float *__restrict__ db) {
    __shared__ float red[8][33];
    const uint32_t c = blockIdx.x * 32 + (threadIdx.x & 31), g = threadIdx.x >> 5;
    float s = 0.f;
    if (c < N)
        for (uint32_t r = g; r < rows; r += 8) s = __fadd_rn(s, dZ[(size_t)r * ld + c]);
    red[g][threadIdx.x & 31] = s;
    __syncthreads();
    if (g == 0 && c < N) {
        float t = red[0][threadIdx.x];
#pragma unroll
        for (int q = 1; q < 8; ++q) t = __fadd_rn(t, red[q][threadIdx.x]);
        db[c] = t;
    }
}

struct AzbAdam {
    float lr, beta1, beta2, eps, l2;
    float bc1, bc2;  // 1 / (1 - beta^t)
};

// dfdx 0.13 Adam (tensor_ops/adam): g += l2 p; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
// p -= lr (m / (1 - b1^t)) / (sqrt(v / (1 - b2^t)) + eps).  The gradient buffer is zeroed (model.zero_grads).
__global__ void __launch_bounds__(256) azb_adam_kernel(float *__restrict__ p, float *__restrict__ g,
                                                       float *__restrict__ m, float *__restrict__ v, size_t n,
                                                       const AzbAdam cfg) {
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float pi = p[i];
        const float gi = __fadd_rn(g[i], __fmul_rn(cfg.l2, pi));
        const float mi = __fadd_rn(__fmul_rn(m[i], cfg.beta1), __fmul_rn(gi, 1.0f - cfg.beta1));
        const float vi = __fadd_rn(__fmul_rn(v[i], cfg.beta2), __fmul_rn(__fmul_rn(gi, gi), 1.0f - cfg.beta2));
        m[i] = mi;
        v[i] = vi;
        const float mh = __fmul_rn(mi, cfg.bc1), vh = __fmul_rn(vi, cfg.bc2);
        const float upd = __fdiv_rn(__fmul_rn(cfg.lr, mh), __fadd_rn(__fsqrt_rn(vh), cfg.eps));
        p[i] = __fsub_rn(pi, upd);
        g[i] = 0.f;
    }
}
