// Warp-specialised flash attention forward for sm_100a (head_dim 128, bf16, non-causal).
// Replaces F.scaled_dot_product_attention in sat/transformer_defaults.py:67-72 for both the
// spatiotemporal self-attention (dit_video_crossattn_sc_xc.py:1092-1094) and the two short-KV
// cross-attention calls (:1159-1197).
//
// One CTA owns 256 query rows of one (batch, head): two 128-row Q tiles that ping-pong on the
// tensor core.  Roles (12 warps = 3 warpgroups, registers re-split with setmaxnreg: 208 / 208 / 88):
//   warps 0-3  softmax warpgroup for Q tile 0      warps 4-7  softmax warpgroup for Q tile 1
//   warp  8    TMA producer (Q once, K/V rings)     warp  9    tcgen05.mma issuer + TMEM owner   (10, 11 idle)
// The MMA warp keeps its control flow warp-uniform (all lanes wait on the mbarriers; descriptors are computed in
// uniform registers) and only the tcgen05 instructions are issued by the elected lane: with the issue loop inside a
// divergent `if (lane == 0)` each tcgen05.mma cost ~80 cycles to issue, more than a 64-cycle 128x128x16 MMA runs.
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each; P (bf16) aliases the first 64
// columns of its S buffer and feeds the PV product as the TMEM A operand (no smem round trip).
// S = Q K^T : UMMA 128x128x16, A/B K-major from smem (TMA SWIZZLE_128B).
// O += P V  : UMMA 128x128x16, A from TMEM, B = V tile read MN-major straight from its [kv, d] layout.
// Online softmax keeps a (possibly stale) running max; O is rescaled only when the max grew by > 2^8.
#pragma once
#include "sm100.cuh"

// Perf-experiment hooks (ablation modes via SCAIL_ATTN_DEBUG, clock64 traces via scail_debug_set_attention_trace) are
// compiled only with -DSCAIL_ATTN_EXPERIMENTS (see scripts/trace_attn.py); the product build carries none of them.
#ifdef SCAIL_ATTN_EXPERIMENTS
#define SCAIL_ATTN_IF_NOT_DEBUG2 if (p.debug != 2)
#define SCAIL_ATTN_TRACE_DECL(...) __VA_ARGS__
#define SCAIL_ATTN_TRACE(IDX) if (tr) p.trace[IDX] = clock64()
#else
#define SCAIL_ATTN_IF_NOT_DEBUG2
#define SCAIL_ATTN_TRACE_DECL(...)
#define SCAIL_ATTN_TRACE(IDX)
#endif

namespace scail {

constexpr int ATT_D = 128;
constexpr int ATT_BQ = 128;   // rows per Q tile (2 tiles per CTA)
constexpr int ATT_BKV = 128;
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;  // 32 KB: one 128x128 bf16 tile (two 64-column halves)
constexpr int ATT_HALF_BYTES = ATT_TILE_BYTES / 2;
constexpr int ATT_THREADS = 384;  // 3 warpgroups: softmax0, softmax1, {TMA, MMA, 2 idle warps}
constexpr int ATT_POLY_EVERY = 0;  // every 4th column pair takes the FMA-pipe exp2 (0 = never)
constexpr int ATT_SMEM_BYTES = (2 + 2 * ATT_KV_STAGES) * ATT_TILE_BYTES + 1024 + 256;

struct AttnParams {
    __nv_bfloat16* out;  // [B*q_rows_per_batch, ldo]; head h written at columns [h*128, h*128+128)
    int64_t ldo;
    int q_len;           // valid query rows per batch
    int kv_len;          // valid key rows per batch
    int q_batch_rows;    // row stride between batches in the Q matrix / out matrix
    int kv_batch_rows;   // row stride between batches in the K/V matrices
    float scale_log2;    // softmax scale * log2(e)
    int accumulate;      // out += result (second cross-attention pass, dit_video_crossattn_sc_xc.py:1197)
    long long* trace;    // perf experiments only: per-iteration clock64 stamps of CTA (0,0,0), or null
    int debug;           // perf experiments only (SCAIL_ATTN_DEBUG): 1 = softmax skips its math, 2 = MMA ignores P barriers
};

// One 128x128 score tile of one softmax warp (thread = row): TMEM S -> running max (lazy O rescale) -> exp2 ->
// bf16 P back into TMEM.  MASK = this is the partial last KV tile.
template <bool MASK>
__device__ __forceinline__ void softmax_tile(uint32_t s_tmem, uint32_t o_tmem, float scale_log2, int valid, int j,
                                             float& m_run, float& l_run, uint32_t bar_pfull) {
    uint32_t s0[32], s1[32], s2[32], s3[32];
    tmem_ld_32x32(s_tmem + 0, s0);
    tmem_ld_32x32(s_tmem + 32, s1);
    tmem_ld_32x32(s_tmem + 64, s2);
    tmem_ld_32x32(s_tmem + 96, s3);
    tmem_ld_wait();
    if constexpr (MASK) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            if (c >= valid) s0[c] = 0xff800000u;
            if (c + 32 >= valid) s1[c] = 0xff800000u;
            if (c + 64 >= valid) s2[c] = 0xff800000u;
            if (c + 96 >= valid) s3[c] = 0xff800000u;
        }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 32; ++c) {
        mx = fmaxf(mx, fmaxf(fmaxf(__uint_as_float(s0[c]), __uint_as_float(s1[c])),
                             fmaxf(__uint_as_float(s2[c]), __uint_as_float(s3[c]))));
    }
    const float m_new = fmaxf(m_run, mx * scale_log2);
    const bool need = (m_new - m_run) > 8.0f;  // also true on the first tile (m_run = -inf)
    if (__any_sync(0xffffffffu, need)) {
        const float alpha = fast_exp2(m_run - m_new);  // 0 on the first tile
        m_run = m_new;
        l_run *= alpha;
        if (j > 0) {
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                uint32_t o[32];
                tmem_ld_32x32(o_tmem + c * 32, o);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; ++k) o[k] = __float_as_uint(__uint_as_float(o[k]) * alpha);
                tmem_st_32x32(o_tmem + c * 32, o);
            }
        }
    }
    const float neg_m = -m_run;
    const uint64_t sc2 = pack_f32x2(scale_log2, scale_log2), nm2 = pack_f32x2(neg_m, neg_m);
    uint64_t sum_a = 0ull, sum_b = 0ull;  // packed (0.f, 0.f)
#define SCAIL_ATT_P_CHUNK(SRC, CH)                                                                          \
    {                                                                                               \
        uint32_t pk[16];                                                                            \
        _Pragma("unroll") for (int c = 0; c < 16; ++c) {                                            \
            const uint64_t y2 = fma_f32x2(pack_f32x2(__uint_as_float(SRC[2 * c]), __uint_as_float(SRC[2 * c + 1])), sc2, nm2); \
            float e0, e1;                                                                           \
            if (ATT_POLY_EVERY > 0 && (c % (ATT_POLY_EVERY > 0 ? ATT_POLY_EVERY : 1)) == ATT_POLY_EVERY - 1) {                   \
                poly_exp2_x2(y2, e0, e1);                                                           \
            } else {                                                                                \
                float y0, y1;                                                                       \
                unpack_f32x2(y2, y0, y1);                                                           \
                e0 = fast_exp2(y0);                                                                 \
                e1 = fast_exp2(y1);                                                                 \
            }                                                                                       \
            if (c & 1) sum_b = add_f32x2(sum_b, pack_f32x2(e0, e1));                                \
            else sum_a = add_f32x2(sum_a, pack_f32x2(e0, e1));                                      \
            pk[c] = pack_bf16(e0, e1);                                                              \
        }                                                                                           \
        tmem_st_32x16(s_tmem + (CH) * 16, pk);                                                      \
    }
    SCAIL_ATT_P_CHUNK(s0, 0)
    SCAIL_ATT_P_CHUNK(s1, 1)
    SCAIL_ATT_P_CHUNK(s2, 2)
    SCAIL_ATT_P_CHUNK(s3, 3)
#undef SCAIL_ATT_P_CHUNK
    float la, lb, lc, ld;
    unpack_f32x2(sum_a, la, lb);
    unpack_f32x2(sum_b, lc, ld);
    const float lsum = (la + lb) + (lc + ld);
    l_run += lsum;
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_arrive(bar_pfull);
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t q_smem = smem_base;                                     // 2 tiles
    const uint32_t k_smem = smem_base + 2 * ATT_TILE_BYTES;                // ATT_KV_STAGES tiles
    const uint32_t v_smem = k_smem + ATT_KV_STAGES * ATT_TILE_BYTES;       // ATT_KV_STAGES tiles
    const uint32_t bar_base = v_smem + ATT_KV_STAGES * ATT_TILE_BYTES;
    enum { B_QFULL = 0, B_KFULL = 1, B_KEMPTY = 3, B_VFULL = 5, B_VEMPTY = 7, B_SFULL = 9, B_PFULL = 11, B_OFULL = 13, B_PHALF = 15, B_COUNT = 17 };
    auto bar = [&](int i) { return bar_base + 8u * i; };
    const uint32_t tmem_slot = bar_base + 8u * B_COUNT;
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int batch = blockIdx.z;
    const int q0 = blockIdx.x * (2 * ATT_BQ);
    const int n_kv = (p.kv_len + ATT_BKV - 1) / ATT_BKV;

    if (warp == 8 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_k);
        tma_prefetch_desc(&tmap_v);
        mbar_init(bar(B_QFULL), 1);
        for (int s = 0; s < ATT_KV_STAGES; ++s) {
            mbar_init(bar(B_KFULL + s), 1);
            mbar_init(bar(B_KEMPTY + s), 1);
            mbar_init(bar(B_VFULL + s), 1);
            mbar_init(bar(B_VEMPTY + s), 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(B_SFULL + i), 1);
            mbar_init(bar(B_PFULL + i), 4);  // one arrive per softmax warp
            mbar_init(bar(B_OFULL + i), 1);
        }
        fence_barrier_init();
    }
    if (warp == 9) {
        tmem_alloc<1>(tmem_slot, 512);
        tmem_relinquish<1>();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    if (warp >= 8) setmaxnreg_dec<88>();
    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int col = head * ATT_D;
            const int qrow = batch * p.q_batch_rows + q0;
            mbar_expect_tx(bar(B_QFULL), 2 * ATT_TILE_BYTES);
            for (int t = 0; t < 2; ++t)
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(q_smem + t * ATT_TILE_BYTES + h * ATT_HALF_BYTES, &tmap_q, bar(B_QFULL), col + h * 64,
                                qrow + t * ATT_BQ);
            const int kvrow = batch * p.kv_batch_rows;
            for (int j = 0; j < n_kv; ++j) {
                const int s = j % ATT_KV_STAGES;
                const uint32_t ph = (j / ATT_KV_STAGES) & 1;
                mbar_wait(bar(B_KEMPTY + s), ph ^ 1, 10);
                mbar_expect_tx(bar(B_KFULL + s), ATT_TILE_BYTES);
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(k_smem + s * ATT_TILE_BYTES + h * ATT_HALF_BYTES, &tmap_k, bar(B_KFULL + s), col + h * 64,
                                kvrow + j * ATT_BKV);
                mbar_wait(bar(B_VEMPTY + s), ph ^ 1, 11);
                mbar_expect_tx(bar(B_VFULL + s), ATT_TILE_BYTES);
                for (int h = 0; h < 2; ++h)
                    tma_load_2d(v_smem + s * ATT_TILE_BYTES + h * ATT_HALF_BYTES, &tmap_v, bar(B_VFULL + s), col + h * 64,
                                kvrow + j * ATT_BKV);
            }
        }
    } else if (warp == 9) {
        // ===================== MMA issuer =====================
        // The whole warp runs the control flow (waits, descriptor arithmetic stay warp-uniform => uniform
        // datapath); only the tcgen05 instructions themselves are issued by the elected lane.
        {
            const bool leader = elect_one_sync();
            constexpr uint32_t idesc_qk = umma_idesc_bf16(128, 128, 0, 0);
            constexpr uint32_t idesc_pv = umma_idesc_bf16(128, 128, 0, 1);  // B (=V) is MN-major
            const uint64_t q_desc0 = umma_desc_kmajor_sw128(q_smem), q_desc1 = umma_desc_kmajor_sw128(q_smem + ATT_TILE_BYTES);
            const uint64_t k_desc0 = umma_desc_kmajor_sw128(k_smem);
            const uint64_t v_desc0 = umma_desc_mnmajor_sw128(v_smem, ATT_HALF_BYTES);
            constexpr uint64_t STAGE_STEP = ATT_TILE_BYTES >> 4;  // descriptor address units are 16 B
            auto issue_qk = [&](int tile, int ks) {
#ifdef SCAIL_ATTN_EXPERIMENTS
                if (p.debug == 5) return;
#endif
                const uint32_t d = tmem_base + tile * 128;
                const uint64_t qa = tile ? q_desc1 : q_desc0, kb = k_desc0 + ks * STAGE_STEP;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t off = ((k >> 2) * ATT_HALF_BYTES + (k & 3) * 32) >> 4;
                        umma_ss<1>(d, qa + off, kb + off, idesc_qk, k != 0);
                    }
                }
            };
            auto issue_pv = [&](int tile, int vs, bool acc) {
#ifdef SCAIL_ATTN_EXPERIMENTS
                if (p.debug == 6) return;
#endif
                const uint32_t d = tmem_base + 256 + tile * 128;
                const uint32_t a = tmem_base + tile * 128;  // P aliases S columns [0,64)
                const uint64_t vb = v_desc0 + vs * STAGE_STEP;
                if (leader) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        // 16 kv rows per step = 2048 B inside each 64-column half; halves are 16 KB apart (LBO)
                        umma_ts(d, a + k * 8, vb + k * (2048 >> 4), idesc_pv, acc || k != 0);
                    }
                }
            };
            auto commit = [&](int b) {
                if (leader) umma_commit(bar(b));
            };
            mbar_wait(bar(B_QFULL), 0, 20);
            mbar_wait(bar(B_KFULL + 0), 0, 21);
            tc_fence_after();
            issue_qk(0, 0);
            commit(B_SFULL + 0);
            issue_qk(1, 0);
            commit(B_SFULL + 1);
            commit(B_KEMPTY + 0);
            for (int j = 0; j < n_kv; ++j) {
                const int vs = j % ATT_KV_STAGES;
                const uint32_t vph = (j / ATT_KV_STAGES) & 1;
                const bool more = j + 1 < n_kv;
                const int ks = (j + 1) % ATT_KV_STAGES;
                const uint32_t kph = ((j + 1) / ATT_KV_STAGES) & 1;
                mbar_wait(bar(B_VFULL + vs), vph, 22);
                SCAIL_ATTN_TRACE_DECL(const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 64 && leader;)
                SCAIL_ATTN_TRACE(j * 8 + 0);
                SCAIL_ATTN_IF_NOT_DEBUG2 mbar_wait(bar(B_PFULL + 0), j & 1, 23);
                SCAIL_ATTN_TRACE(j * 8 + 1);
                tc_fence_after();
                issue_pv(0, vs, j > 0);
                if (more) {
                    mbar_wait(bar(B_KFULL + ks), kph, 24);
                    tc_fence_after();
                    issue_qk(0, ks);
                    commit(B_SFULL + 0);
                    SCAIL_ATTN_TRACE(j * 8 + 2);
                } else {
                    commit(B_OFULL + 0);
                }
                SCAIL_ATTN_IF_NOT_DEBUG2 mbar_wait(bar(B_PFULL + 1), j & 1, 25);
                SCAIL_ATTN_TRACE(j * 8 + 3);
                tc_fence_after();
                issue_pv(1, vs, j > 0);
                commit(B_VEMPTY + vs);
                if (more) {
                    issue_qk(1, ks);
                    commit(B_SFULL + 1);
                    commit(B_KEMPTY + ks);
                } else {
                    commit(B_OFULL + 1);
                }
            }
        }
    } else if (warp < 8) {
        // ===================== softmax warpgroups (+ O rescale + epilogue) =====================
        setmaxnreg_inc<208>();
        const int tile = warp >> 2;  // 0 or 1
        const int sub = warp & 3;
        const uint32_t lane_off = static_cast<uint32_t>(sub * 32) << 16;
        const uint32_t s_tmem = tmem_base + lane_off + tile * 128;
        const uint32_t o_tmem = tmem_base + lane_off + 256 + tile * 128;
        float m_run = -INFINITY;  // running max, already multiplied by scale_log2
        float l_run = 0.f;
        const int n_full = p.kv_len / ATT_BKV;  // full tiles; an optional partial tile follows (peeled: no per-iteration branch)
        for (int j = 0; j < n_full; ++j) {
            SCAIL_ATTN_TRACE_DECL(const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && j < 64 && warp == 0 && lane == 0;)
            SCAIL_ATTN_TRACE(j * 8 + 4);
            mbar_wait(bar(B_SFULL + tile), j & 1, 30 + tile);
            SCAIL_ATTN_TRACE(j * 8 + 5);
            tc_fence_after();
            constexpr int valid = ATT_BKV;
#ifdef SCAIL_ATTN_EXPERIMENTS
            if (p.debug == 3) {  // TMEM read only
                uint32_t t0[32], t1[32], t2[32], t3[32];
                tmem_ld_32x32(s_tmem + 0, t0);
                tmem_ld_32x32(s_tmem + 32, t1);
                tmem_ld_32x32(s_tmem + 64, t2);
                tmem_ld_32x32(s_tmem + 96, t3);
                tmem_ld_wait();
                uint32_t x = 0;
#pragma unroll
                for (int c = 0; c < 32; ++c) x ^= t0[c] ^ t1[c] ^ t2[c] ^ t3[c];
                if (x == 0x12345678u) l_run += 1.f;
            }
            if (p.debug == 1 || p.debug == 3 || p.debug >= 5) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(B_PFULL + tile));
                l_run = 1.f;
                continue;
            }
#endif
            softmax_tile<false>(s_tmem, o_tmem, p.scale_log2, valid, j, m_run, l_run, bar(B_PFULL + tile));
        }
        if (n_full < n_kv) {  // partial last KV tile: masked instantiation
            mbar_wait(bar(B_SFULL + tile), n_full & 1, 32 + tile);
            tc_fence_after();
            softmax_tile<true>(s_tmem, o_tmem, p.scale_log2, p.kv_len - n_full * ATT_BKV, n_full, m_run, l_run, bar(B_PFULL + tile));
        }
        // ---- epilogue: O / l -> bf16 -> global ----
        mbar_wait(bar(B_OFULL + tile), 0, 40 + tile);
        tc_fence_after();
        const int qi = q0 + tile * ATT_BQ + sub * 32 + lane;
        const bool row_ok = qi < p.q_len;
        const float inv_l = 1.0f / l_run;
        __nv_bfloat16* orow = p.out + (static_cast<int64_t>(batch) * p.q_batch_rows + qi) * p.ldo + head * ATT_D;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld_32x32(o_tmem + c * 32, o);
            tmem_ld_wait();
            if (row_ok) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float f[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(o[g * 8 + k]) * inv_l;
                    uint4* dst = reinterpret_cast<uint4*>(orow + c * 32 + g * 8);
                    if (p.accumulate) {
                        uint4 prev = *dst;
                        const uint32_t pw[4] = {prev.x, prev.y, prev.z, prev.w};
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            float2 t = unpack_bf16(pw[k]);
                            f[2 * k] += t.x;
                            f[2 * k + 1] += t.y;
                        }
                    }
                    uint4 ov;
                    ov.x = pack_bf16(f[0], f[1]);
                    ov.y = pack_bf16(f[2], f[3]);
                    ov.z = pack_bf16(f[4], f[5]);
                    ov.w = pack_bf16(f[6], f[7]);
                    *dst = ov;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc<1>(tmem_base, 512);
    }
}

}  // namespace scail
