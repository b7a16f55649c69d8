// Persistent warp-specialised bf16 GEMM for sm_100a:  C[M,N] = epi(A[M,K] * W[N,K]^T)
//   * TMA (SWIZZLE_128B) stages A/W tiles into a 4-deep shared-memory ring,
//   * one elected thread issues tcgen05.mma (UMMA 128x256x16, fp32 accumulators in TMEM),
//   * TMEM accumulators are double-buffered (2 x 256 columns) so the epilogue of tile i
//     overlaps the main loop of tile i+1,
//   * 4 epilogue warps read TMEM (tcgen05.ld 32x32b: thread = row), apply bias / activation / gate in registers, then
//     transpose 32x32 fp32 chunks through a swizzled shared-memory staging tile so that the residual read and the bf16
//     store are row-contiguous (8 rows x 64 B per warp instruction instead of 32 rows x 16 B: every 32-byte sector is
//     touched once and completely; the per-thread-row accesses of round 1 made the out-projection epilogue as long as its
//     mainloop).
// Replaces the reference's F.linear calls (sat/mpu/layers.py:230-243, :425-444) together with
// the elementwise ops that follow them (bias, GELU-tanh, gate*out + residual).
#pragma once
#include "sm100.cuh"

namespace scail {

enum GemmEpilogue : int {
    EPI_BIAS = 0,           // C = acc + bias
    EPI_BIAS_GELU = 1,      // C = gelu_tanh(acc + bias)            (sat/transformer_defaults.py:173-174)
    EPI_BIAS_GATE_RES = 2,  // C = res + gate[b] * (acc + bias)     (dit_video_crossattn_sc_xc.py:1036,1050)
    EPI_BIAS_RES = 3,       // C = res + (acc + bias)               (dit_video_crossattn_sc_xc.py:1042)
    EPI_BIAS_SILU = 4,      // C = silu(acc + bias)
    EPI_BIAS_GELU_ERF = 5,  // C = gelu(acc + bias), exact erf form (MLPProj, dit_video_crossattn_sc_xc.py:38)
};

struct GemmParams {
    int M, N, K;
    const __nv_bfloat16* bias;      // [N] or null
    const __nv_bfloat16* gate;      // [B, gate_stride] (row b = m / rows_per_batch) or null
    const __nv_bfloat16* residual;  // [M, ldr] or null
    __nv_bfloat16* C;               // [M, ldc]
    float* C32;                     // optional fp32 output instead of bf16
    int64_t ldc, ldr, gate_stride;
    int rows_per_batch;
    int epilogue;
    int group_m;  // rasterisation: m-blocks per L2 group
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;  // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;  // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;
constexpr int GEMM_EPI_STAGE_BYTES = 32 * 32 * 4;  // one 32x32 fp32 chunk per epilogue warp
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 4 * GEMM_EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
constexpr int GEMM_THREADS = 256;

__device__ __forceinline__ void gemm_tile_coords(int tile, int num_m, int num_n, int group_m, int& m_blk, int& n_blk) {
    int per_group = group_m * num_n;
    int g = tile / per_group;
    int first_m = g * group_m;
    int gsize = min(group_m, num_m - first_m);
    int r = tile - g * per_group;
    n_blk = r / gsize;
    m_blk = first_m + (r - n_blk * gsize);
}

__device__ __forceinline__ float epi_act(float v, int epilogue) {
    if (epilogue == EPI_BIAS_GELU) return gelu_tanh(v);
    if (epilogue == EPI_BIAS_SILU) return v / (1.0f + __expf(-v));
    if (epilogue == EPI_BIAS_GELU_ERF) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    return v;
}

__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t epi_base = smem_base + GEMM_STAGES * GEMM_STAGE_BYTES;
    const uint32_t bar_base = epi_base + 4 * GEMM_EPI_STAGE_BYTES;
    // barrier layout (8 B each): full[S], empty[S], tmem_full[2], tmem_empty[2], then tmem ptr slot
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (GEMM_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * GEMM_STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * GEMM_STAGES + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (p.M + GEMM_BM - 1) / GEMM_BM;
    const int num_n = (p.N + GEMM_BN - 1) / GEMM_BN;
    const int num_tiles = num_m * num_n;
    const int num_k = (p.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_w);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(tfull_bar(s), 1);
            mbar_init(tempty_bar(s), 4);  // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp == 2) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                int m_blk, n_blk;
                gemm_tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1, 1);
                    const uint32_t sa = smem_base + stage * GEMM_STAGE_BYTES;
                    mbar_expect_tx(full_bar(stage), GEMM_STAGE_BYTES);
                    tma_load_2d(sa, &tmap_a, full_bar(stage), kb * GEMM_BK, m_blk * GEMM_BM);
                    tma_load_2d(sa + GEMM_A_BYTES, &tmap_w, full_bar(stage), kb * GEMM_BK, n_blk * GEMM_BN);
                    if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // whole warp runs the (warp-uniform) control flow; the elected lane issues tcgen05.mma / commit
        {
            const bool leader = elect_one_sync();
            constexpr uint32_t idesc = umma_idesc_bf16(GEMM_BM, GEMM_BN, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(tempty_bar(acc), acc_phase ^ 1, 2);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * GEMM_BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(full_bar(stage), phase, 3);
                    tc_fence_after();
                    const uint32_t sa = smem_base + stage * GEMM_STAGE_BYTES;
                    const uint64_t da = umma_desc_kmajor_sw128(sa);
                    const uint64_t db = umma_desc_kmajor_sw128(sa + GEMM_A_BYTES);
                    if (leader) {
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k) {
                            // +32 B per K=16 step inside the 128-B swizzle atom (descriptor units of 16 B)
                            umma_ss<1>(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                        }
                        umma_commit(empty_bar(stage));  // frees the smem stage when these MMAs have read it
                    }
                    if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
                }
                if (leader) umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue warps =====================
        const int sub = warp & 3;  // TMEM sub-partition this warp may access: lanes [32*sub, 32*sub+32)
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            int m_blk, n_blk;
            gemm_tile_coords(tile, num_m, num_n, p.group_m, m_blk, n_blk);
            mbar_wait(tfull_bar(acc), acc_phase, 4);
            tc_fence_after();
            const int row = m_blk * GEMM_BM + sub * 32 + lane;  // phase 1: this thread's accumulator row
            const int bidx = row < p.M ? row / p.rows_per_batch : 0;
            const uint32_t t_row = tmem_base + (static_cast<uint32_t>(sub * 32) << 16) + acc * GEMM_BN;
            const uint32_t stage = epi_base + sub * GEMM_EPI_STAGE_BYTES;
            // phase 2 mapping: 4 lanes per row (8 columns each), 8 rows per instruction
            const int r2 = lane >> 2, cg = lane & 3;
#pragma unroll 1
            for (int c = 0; c < GEMM_BN / 32; ++c) {
                const int col0 = n_blk * GEMM_BN + c * 32;
                if (col0 >= p.N) break;  // warp-uniform
                uint32_t v[32];
                tmem_ld_32x32(t_row + c * 32, v);
                tmem_ld_wait();
                // ---- phase 1 (thread = row): bias, activation, gate in fp32; park the chunk in the staging tile ----
#pragma unroll
                for (int g = 0; g < 4; ++g) {  // groups of 8 columns
                    const int col = col0 + g * 8;
                    float f[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[g * 8 + j]);
                    if (col < p.N) {
                        if (p.bias) {
                            uint4 bv = *reinterpret_cast<const uint4*>(p.bias + col);  // warp-uniform address: broadcast
                            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 b2 = unpack_bf16(bw[j]);
                                f[2 * j] += b2.x;
                                f[2 * j + 1] += b2.y;
                            }
                        }
                        if (p.epilogue == EPI_BIAS_GELU || p.epilogue == EPI_BIAS_SILU || p.epilogue == EPI_BIAS_GELU_ERF) {
#pragma unroll
                            for (int j = 0; j < 8; ++j) f[j] = epi_act(f[j], p.epilogue);
                        }
                        if (p.epilogue == EPI_BIAS_GATE_RES) {
                            uint4 gv = *reinterpret_cast<const uint4*>(p.gate + bidx * p.gate_stride + col);
                            const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 g2 = unpack_bf16(gw[j]);
                                f[2 * j] *= g2.x;
                                f[2 * j + 1] *= g2.y;
                            }
                        }
                    }
                    // row `lane` = 128 B = 8 chunks of 16 B; chunk q lives at position q ^ (lane & 7): conflict-free both ways
                    const uint32_t a0 = stage + lane * 128 + (((2 * g) ^ (lane & 7)) << 4);
                    const uint32_t a1 = stage + lane * 128 + (((2 * g + 1) ^ (lane & 7)) << 4);
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a0), "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]) : "memory");
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a1), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]) : "memory");
                }
                __syncwarp();
                // ---- phase 2 (4 lanes per row): residual add, one bf16 rounding, row-contiguous stores ----
                const int col = col0 + cg * 8;
#pragma unroll
                for (int it = 0; it < 4; ++it) {
                    const int rr = it * 8 + r2;
                    const int grow = m_blk * GEMM_BM + sub * 32 + rr;
                    float f[8];
                    const uint32_t a0 = stage + rr * 128 + (((2 * cg) ^ (rr & 7)) << 4);
                    const uint32_t a1 = stage + rr * 128 + (((2 * cg + 1) ^ (rr & 7)) << 4);
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[0]), "=f"(f[1]), "=f"(f[2]), "=f"(f[3]) : "r"(a0) : "memory");
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(f[4]), "=f"(f[5]), "=f"(f[6]), "=f"(f[7]) : "r"(a1) : "memory");
                    if (grow < p.M && col < p.N) {
                        if (p.epilogue == EPI_BIAS_GATE_RES || p.epilogue == EPI_BIAS_RES) {
                            uint4 rv = *reinterpret_cast<const uint4*>(p.residual + static_cast<int64_t>(grow) * p.ldr + col);
                            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float2 r2f = unpack_bf16(rw[j]);
                                f[2 * j] += r2f.x;
                                f[2 * j + 1] += r2f.y;
                            }
                        }
                        if (p.C32) {
                            float4* o = reinterpret_cast<float4*>(p.C32 + static_cast<int64_t>(grow) * p.ldc + col);
                            o[0] = make_float4(f[0], f[1], f[2], f[3]);
                            o[1] = make_float4(f[4], f[5], f[6], f[7]);
                        } else {
                            uint4 o;
                            o.x = pack_bf16(f[0], f[1]);
                            o.y = pack_bf16(f[2], f[3]);
                            o.z = pack_bf16(f[4], f[5]);
                            o.w = pack_bf16(f[6], f[7]);
                            *reinterpret_cast<uint4*>(p.C + static_cast<int64_t>(grow) * p.ldc + col) = o;
                        }
                    }
                }
                __syncwarp();  // the staging tile is rewritten by the next chunk
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

}  // namespace scail
