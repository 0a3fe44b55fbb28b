// azb_pipe.cuh — the prior model of the asynchronous search step as a WEIGHT-STATIONARY PIPELINE of SMs.
//
// Replaces ActionModel::write_predictions (az-discrete-opt/src/nabla/model/dfdx.rs:69-84) for the rows the tree warps
// of azb_async_kernel (azb_async.cuh) publish into the tile ring.  The in-kernel workers of azb_async.cuh stream all
// 2.5 MB of weights through one SM per 128-row tile (65 us per tile: a tree waits ~100 us for priors it needs after a
// ~45 us walk).  Here the weights never move: every CTA of this kernel owns one column block of ONE Linear layer and
// keeps it in shared memory for the whole launch (N = 19: 304->512 on 2 SMs x 256 columns, 512->1024 on 8 x 128,
// 1024->512 on 8 x 64, 512->152 on 2 x 80 = 20 SMs, <= 160 KB each).  Tiles flow through the layers in ring order:
//   stage l, tile q:  wait until every member of stage l-1 has stored its columns of tile q (monotonic counters in
//   global memory, one per activation slot)  ->  TMA the tile's activations [128 x K] through a small ring  ->
//   tcgen05.mma against the resident weights, accumulator in TMEM (double buffered: the epilogue of tile q overlaps
//   the MMAs of tile q + 1)  ->  epilogue warps: tcgen05.ld, bias, ReLU, bf16 rows into activation slot q % PIPE_D
//   (or Sigmoid -> the owning trees' prior rows, then their answer flags)  ->  bump the stage's counter.
// Per tile an SM ingests only the activations (80-256 KB) instead of the weights, and the four layers of a tile are
// spread over four SMs' tensor cores, so a tile's latency is the sum of four short stages instead of one long pass.
// Same arithmetic as the lock-step forward (azb_mlp_tc.cuh): bf16 operands, one fp32 TMEM accumulator per output
// over the whole K in k-block order, the same epilogue expressions — priors are bit-identical, so are the trees.
// Launched beside azb_async_kernel on a second stream (its CTAs need > 113 KB of shared memory each, so none shares an
// SM with a tree CTA); every spin loop watches `abort` and the %globaltimer watchdog like the tree warps do.
#pragma once
#include "azb_async.cuh"

#define PIPE_D 8            // activation slots between two stages (tiles in flight per stage boundary)
#define PIPE_MAX_STAGES 8   // A-ring depth (16 KB each); the host picks what fits beside the weights
#define PIPE_EPI_WARPS 8    // per epilogue group: two warps per TMEM lane quadrant
#define PIPE_EPI_GROUPS 2   // group g drains the tiles q = g (mod 2), i.e. TMEM buffer g: the fixed per-tile cost of an
                            // epilogue (stores -> fences -> counter, ~3 us) is overlapped between the groups
#define PIPE_THREADS ((2 + PIPE_EPI_GROUPS * PIPE_EPI_WARPS) * 32)

struct AzbPipeParams {
    uint32_t S[4], BN[4], first_cta[4];  // members per stage, columns per member, first CTA of the stage
    uint32_t stages[4];                  // A-ring depth per stage
    uint32_t n_ctas;
};

struct AzbPipeMaps {
    CUtensorMap ring;    // stage-0 input: [NT*128 rows][kpad0] bf16, box 64 x 128
    CUtensorMap act[3];  // activation slots [PIPE_D*128 rows][kpad[l+1]], box 64 x 128
    CUtensorMap w[4];    // weights [rows padded to 128][kpad[l]], box 64 x BN[l]
};

__device__ __forceinline__ void pipe_tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void pipe_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(PIPE_THREADS, 1)
    azb_pipe_kernel(const AzbLayout L, const AzbAsyncParams P, const AzbPipeParams Q, const __grid_constant__ AzbPipeMaps M) {
    extern __shared__ __align__(1024) uint8_t pipe_smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[PIPE_MAX_STAGES], empty_bar[PIPE_MAX_STAGES], acc_full[2], acc_empty[2], w_bar;
    __shared__ uint32_t tmem_slot, s_last[2];
    __shared__ volatile uint32_t s_final;  // first tile this CTA will NOT process (set by the producer thread on exit)
    __shared__ uint32_t s_rowtree[2][AS_TILE];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    AzbAsyncState *st = P.st;
    uint8_t *smem = (uint8_t *)(((uintptr_t)pipe_smem_raw + 1023) & ~(uintptr_t)1023);

    // ---- role: stage l, member j
    uint32_t l = 0;
    while (l < 3 && blockIdx.x >= Q.first_cta[l + 1]) ++l;
    const uint32_t j = blockIdx.x - Q.first_cta[l];
    const uint32_t bn = Q.BN[l], n0 = j * bn, S = Q.S[l];
    const uint32_t k_blocks = P.kpad[l] / TC_BK, NS = Q.stages[l];
    const uint32_t w_kb_bytes = bn * TC_BK * 2u, a_bytes = AS_TILE * TC_BK * 2u;
    uint8_t *w_smem = smem;                                     // [k_blocks][bn rows x 128 B], SWIZZLE_128B
    uint8_t *a_ring = smem + (size_t)k_blocks * w_kb_bytes;     // [NS][128 rows x 128 B]
    float *s_bias = reinterpret_cast<float *>(a_ring + (size_t)NS * a_bytes);
    const uint32_t acc_stride = bn <= 32 ? 32u : (bn <= 64 ? 64u : (bn <= 128 ? 128u : 256u));  // TMEM columns per buffer

    if (warp == 0 && lane == 0) {
        for (uint32_t s = 0; s < PIPE_MAX_STAGES; ++s) {
            tc_mbar_init(&full_bar[s], 1);
            tc_mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            tc_mbar_init(&acc_full[b], 1);
            tc_mbar_init(&acc_empty[b], PIPE_EPI_WARPS);
        }
        tc_mbar_init(&w_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_final = 0xffffffffu;
        asm volatile("prefetch.tensormap [%0];" ::"l"(l == 0 ? &M.ring : &M.act[l - 1]) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&M.w[l]) : "memory");
        // the stage's weights, once: k_blocks boxes of [bn rows x 64]
        tc_mbar_expect_tx(&w_bar, k_blocks * w_kb_bytes);
        const uint64_t pol = as_policy_evict_last();
        for (uint32_t kb = 0; kb < k_blocks; ++kb)
            as_tma_load_2d_hint(w_smem + (size_t)kb * w_kb_bytes, &M.w[l], &w_bar, (int)(kb * TC_BK), (int)n0, pol);
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)),
                     "r"(2u * acc_stride)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (uint32_t i = threadIdx.x; i < bn; i += PIPE_THREADS) s_bias[i] = n0 + i < P.npad[l] ? P.bias[l][n0 + i] : 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_slot;
    const unsigned long long t_start = as_now();
    uint32_t *done_in = l ? st->pipe_done + (l - 1) * PIPE_D : nullptr;  // the previous stage's counters
    uint32_t *done_me = st->pipe_done + l * PIPE_D;
    uint32_t *done_out = l < 3 ? st->pipe_done + (l + 1) * PIPE_D : nullptr;  // the next stage's (slot reuse)
    const uint32_t S_in = l ? Q.S[l - 1] : 0u, S_out = l < 3 ? Q.S[l + 1] : 0u;

    if (warp == 0) {
        // ===== producer: waits for tile q's input, then streams its activations through the A ring =====
        if (tc_elect_one()) {
            const CUtensorMap *ma = l == 0 ? &M.ring : &M.act[l - 1];
            uint32_t kbc = 0;
            long long d_in = 0, d_empty = 0, d_tiles = 0;
            for (uint32_t q = 0;; ++q) {
                bool stop = false;
                const long long t_in0 = AS_CLK();
                if (l == 0) {
                    const uint32_t *cnt_p = P.tile_count + (q % P.NT);
                    const uint32_t want = AS_TILE * (q / P.NT + 1u);
                    unsigned long long t_partial = 0;
                    bool flushed = false;
                    for (uint32_t spins = 0;; ++spins) {
                        if (as_ld_volatile(cnt_p) >= want) break;
                        if (as_ld_volatile(&st->abort)) {
                            stop = true;
                            break;
                        }
                        if (j == 0) {
                            // the leader ends the run (every submitted row was answered before its tree could finish) and
                            // tops a stale partial tile up with dummy rows
                            if (as_ld_volatile(&st->done_trees) >= L.B) {
                                st->final_q1 = q + 1u;
                                __threadfence();
                                stop = true;
                                break;
                            }
                            const uint32_t tail = as_ld_volatile(&st->row_tail);
                            if (!flushed && tail > q * AS_TILE && tail < (q + 1u) * AS_TILE) {
                                const unsigned long long now = as_now();
                                if (t_partial == 0) t_partial = now;
                                if (now - t_partial > P.flush_ns) {
                                    const uint32_t k = (q + 1u) * AS_TILE - tail;
                                    const uint32_t old = atomicAdd(&st->row_tail, k);
                                    for (uint32_t i = 0; i < k; ++i) P.slot_tree[(old + i) % (P.NT * AS_TILE)] = AS_NONE;
                                    __threadfence();
                                    for (uint32_t i = 0; i < k; ++i) atomicAdd(P.tile_count + (((old + i) / AS_TILE) % P.NT), 1u);
                                    atomicAdd(&st->rows_dummy, k);
                                    flushed = true;
                                }
                            }
                        } else {
                            const uint32_t f = as_ld_volatile(&st->final_q1);
                            if (f && q + 1u >= f) {
                                stop = true;
                                break;
                            }
                        }
                        __nanosleep(64);
                        if ((spins & 255u) == 255u && as_now() - t_start > P.timeout_ns) {
                            atomicExch(&st->abort, 1u);
                            stop = true;
                            break;
                        }
                    }
                } else {
                    const uint32_t *cnt_p = done_in + (q % PIPE_D);
                    const uint32_t want = S_in * (q / PIPE_D + 1u);
                    for (uint32_t spins = 0;; ++spins) {
                        if (as_ld_volatile(cnt_p) >= want) break;
                        const uint32_t f = as_ld_volatile(&st->final_q1);
                        if ((f && q + 1u >= f) || as_ld_volatile(&st->abort)) {
                            stop = true;
                            break;
                        }
                        if ((spins & 1023u) == 1023u && as_now() - t_start > P.timeout_ns) {
                            atomicExch(&st->abort, 1u);
                            stop = true;
                            break;
                        }
                    }
                }
                d_in += AS_CLK() - t_in0;
                if (stop) {
                    if (P.dbg) {
                        atomicAdd(P.dbg + 24 + l * 8 + 0, (unsigned long long)d_in);
                        atomicAdd(P.dbg + 24 + l * 8 + 1, (unsigned long long)d_empty);
                        atomicAdd(P.dbg + 24 + l * 8 + 2, (unsigned long long)d_tiles);
                    }
                    // wake the MMA warp (and through it the epilogue) on a tile that will not come
                    s_final = q;
                    __threadfence_block();
                    const uint32_t s = kbc % NS, ph = (kbc / NS) & 1u;
                    as_mbar_spin(&empty_bar[s], ph ^ 1u);
                    as_mbar_arrive(&full_bar[s]);
                    break;
                }
                as_fence_proxy_async();  // the rows were written through the generic proxy (tree warps / epilogue warps)
                const int arow = (int)(l == 0 ? (q % P.NT) * AS_TILE : (q % PIPE_D) * AS_TILE);
                for (uint32_t kb = 0; kb < k_blocks; ++kb, ++kbc) {
                    const uint32_t s = kbc % NS, ph = (kbc / NS) & 1u;
                    const long long te0 = AS_CLK();
                    as_mbar_spin(&empty_bar[s], ph ^ 1u);
                    d_empty += AS_CLK() - te0;
                    if (AS_DBG(4u)) {  // timing experiment: no loads, the MMAs run on stale operands
                        as_mbar_arrive(&full_bar[s]);
                        continue;
                    }
                    tc_mbar_expect_tx(&full_bar[s], a_bytes);
                    tc_tma_load_2d(a_ring + (size_t)s * a_bytes, ma, &full_bar[s], (int)(kb * TC_BK), arow);
                }
                d_tiles += 1;
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: tile q accumulates into TMEM buffer q & 1 =====
        tc_mbar_wait(&w_bar, 0);  // weights resident
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((bn >> 3) << 17) | ((AS_TILE >> 4) << 24);
        const uint32_t w_addr = tc_smem_u32(w_smem), a_addr0 = tc_smem_u32(a_ring);
        uint32_t kbc = 0;
        long long d_acc = 0, d_full = 0, d_first = 0, d_issue = 0;
        for (uint32_t q = 0;; ++q) {
            const uint32_t b = q & 1u;
            const long long ta0 = AS_CLK();
            as_mbar_spin(&acc_empty[b], ((q >> 1) & 1u) ^ 1u);  // the epilogue has drained tile q - 2
            d_acc += AS_CLK() - ta0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            bool stop = false;
            for (uint32_t kb = 0; kb < k_blocks; ++kb, ++kbc) {
                const uint32_t s = kbc % NS, ph = (kbc / NS) & 1u;
                const long long tf0 = AS_CLK();
                as_mbar_spin(&full_bar[s], ph);
                if (kb == 0 && s_final <= q) {
                    stop = true;
                    break;
                }
                if (kb == 0) d_first += AS_CLK() - tf0; else d_full += AS_CLK() - tf0;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const long long ti0 = AS_CLK();
                if (tc_elect_one()) {
                    const uint32_t a_addr = a_addr0 + s * a_bytes, b_addr = w_addr + kb * w_kb_bytes;
#pragma unroll
                    for (uint32_t k = 0; k < TC_BK / 16; ++k)
                        tc_umma_f16(tmem_base + b * acc_stride, tc_smem_desc(a_addr + k * 32u), tc_smem_desc(b_addr + k * 32u), idesc,
                                    (kb | k) != 0u ? 1u : 0u);
                    tc_umma_commit(&empty_bar[s]);
                    if (kb + 1 == k_blocks) tc_umma_commit(&acc_full[b]);
                }
                __syncwarp();
                d_issue += AS_CLK() - ti0;
            }
            if (stop) {
                if (P.dbg && lane == 0) atomicAdd(P.dbg + 56 + l, (unsigned long long)d_issue);
                // wake both epilogue groups behind every MMA issued so far (commit, not arrive: the groups' current
                // phases must complete in order): s_final tells them there is no tile q
                if (tc_elect_one()) {
                    tc_umma_commit(&acc_full[b]);
                    tc_umma_commit(&acc_full[b ^ 1u]);
                }
                if (P.dbg && lane == 0) {
                    atomicAdd(P.dbg + 24 + l * 8 + 3, (unsigned long long)d_acc);
                    atomicAdd(P.dbg + 24 + l * 8 + 4, (unsigned long long)d_full);
                    atomicAdd(P.dbg + 24 + l * 8 + 7, (unsigned long long)d_first);
                }
                break;
            }
        }
    } else {
        // ===== epilogue (warps 2..9): TMEM -> registers -> bias + activation -> activation slot / prior rows =====
        const uint32_t grp = (warp - 2u) / PIPE_EPI_WARPS, gw = (warp - 2u) % PIPE_EPI_WARPS;  // group, warp in the group
        const uint32_t q4 = warp & 3u, half = gw >> 2, row = q4 * 32u + lane;
        const uint32_t et = gw * 32u + lane, bar_id = 2u + grp;
        const bool leader = et == 0u;
        const uint32_t n_slices = bn / 16u, sl_lo = half ? (n_slices + 1u) / 2u : 0u, sl_hi = half ? n_slices : (n_slices + 1u) / 2u;
        long long d_wait = 0, d_busy = 0, d_slot = 0, d_fence = 0;
        for (uint32_t q = grp;; q += PIPE_EPI_GROUPS) {
            const uint32_t b = q & 1u, slot = q % PIPE_D;
            const long long tw0 = AS_CLK();
            as_mbar_spin(&acc_full[b], (q >> 1) & 1u);
            if (s_final <= q) break;
            const long long tb0 = AS_CLK();
            d_wait += tb0 - tw0;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            uint32_t my_tree = AS_NONE;
            if (l == 3) {
                // owners of the tile's rows (published by the tree warps before the tile counted as full)
                if (et < AS_TILE) s_rowtree[b][et] = __ldcg(P.slot_tree + (q % P.NT) * AS_TILE + et);
                as_named_bar(bar_id, PIPE_EPI_WARPS * 32);
                my_tree = s_rowtree[b][row];
            } else if (q >= PIPE_D) {
                // the slot still holds tile q - PIPE_D until every member of the next stage has consumed it
                const long long ts0 = AS_CLK();
                if (leader) {
                    const uint32_t want = S_out * (q / PIPE_D);
                    for (uint32_t spins = 0; as_ld_volatile(done_out + slot) < want; ++spins) {
                        if (as_ld_volatile(&st->abort)) break;
                        if ((spins & 1023u) == 1023u && as_now() - t_start > P.timeout_ns) atomicExch(&st->abort, 1u);
                    }
                }
                as_named_bar(bar_id, PIPE_EPI_WARPS * 32);
                d_slot += AS_CLK() - ts0;
            }
            const uint32_t t_row = tmem_base + ((q4 * 32u) << 16) + b * acc_stride;
            for (uint32_t s0 = sl_lo; s0 < sl_hi; s0 += 4u) {
                uint32_t r[4][16];
#pragma unroll
                for (uint32_t i = 0; i < 4; ++i)
                    if (s0 + i < sl_hi) pipe_tmem_ld16_issue(t_row + (s0 + i) * 16u, r[i]);
                pipe_tmem_ld_wait();
#pragma unroll
                for (uint32_t i = 0; i < 4; ++i) {
                    if (s0 + i >= sl_hi) continue;
                    const uint32_t c0 = (s0 + i) * 16u;
                    float bq[16];
#pragma unroll
                    for (int t = 0; t < 16; t += 4) {
                        const float4 b4 = *reinterpret_cast<const float4 *>(s_bias + c0 + t);
                        bq[t] = b4.x;
                        bq[t + 1] = b4.y;
                        bq[t + 2] = b4.z;
                        bq[t + 3] = b4.w;
                    }
                    if (l < 3) {
                        __nv_bfloat16 *dst = P.act[l] + (size_t)(slot * AS_TILE + row) * P.kpad[l + 1] + n0 + c0;
                        uint32_t pk[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            float v0 = __uint_as_float(r[i][2 * t]) + bq[2 * t];
                            float v1 = __uint_as_float(r[i][2 * t + 1]) + bq[2 * t + 1];
                            v0 = v0 > 0.f ? v0 : 0.f;
                            v1 = v1 > 0.f ? v1 : 0.f;
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(v0, v1);
                            pk[t] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        *reinterpret_cast<uint4 *>(dst + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                    } else if (my_tree != AS_NONE) {
                        float *dst = L.h + (size_t)my_tree * L.h_ld;
                        const bool vec = (L.h_ld & 3u) == 0u;
#pragma unroll
                        for (int t0 = 0; t0 < 16; t0 += 4) {
                            float o[4];
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const float v = __uint_as_float(r[i][t0 + t]) + bq[t0 + t];
                                o[t] = __fdividef(1.0f, 1.0f + __expf(-v));
                            }
                            const uint32_t n = n0 + c0 + t0;
                            if (vec && n + 3u < L.A) {
                                *reinterpret_cast<float4 *>(dst + n) = make_float4(o[0], o[1], o[2], o[3]);
                            } else {
#pragma unroll
                                for (int t = 0; t < 4; ++t)
                                    if (n + t < L.A) dst[n + t] = o[t];
                            }
                        }
                    }
                }
            }
            // TMEM buffer b is free for tile q + 2
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) as_mbar_arrive(&acc_empty[b]);
            // this member's columns of tile q are stored: tell the next stage (or, last member of the head: the trees)
            const long long tf0 = AS_CLK();
            if (l < 3) as_fence_proxy_async();  // the next stage reads these rows through TMA
            __threadfence();
            as_named_bar(bar_id, PIPE_EPI_WARPS * 32);
            if (leader) {
                const uint32_t v = atomicAdd(done_me + slot, 1u) + 1u;
                s_last[grp] = (l == 3 && v == S * (q / PIPE_D + 1u)) ? 1u : 0u;
            }
            d_fence += AS_CLK() - tf0;
            if (l == 3) {
                as_named_bar(bar_id, PIPE_EPI_WARPS * 32);
                if (s_last[grp]) {
                    __threadfence();
                    if (half == 0u && my_tree != AS_NONE) atomicAdd(P.h_flag + my_tree, 1u);
                    if (leader) {
                        atomicAdd(P.tile_retired + (q % P.NT), 1u);
                        atomicAdd(&st->tiles_done, 1u);
                    }
                }
            }
            d_busy += AS_CLK() - tb0;
        }
        if (P.dbg && leader) {
            atomicAdd(P.dbg + 24 + l * 8 + 5, (unsigned long long)d_wait);
            atomicAdd(P.dbg + 24 + l * 8 + 6, (unsigned long long)d_busy);
            atomicAdd(P.dbg + 60, (unsigned long long)(l == 0 ? d_fence : 0));
            atomicAdd(P.dbg + 61, (unsigned long long)(l == 1 ? d_slot : 0));
            atomicAdd(P.dbg + 62, (unsigned long long)(l == 1 ? d_fence : 0));
            atomicAdd(P.dbg + 63, (unsigned long long)(l == 2 ? d_fence : 0));
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * acc_stride) : "memory");
    }
}
