// Operand-bandwidth microbenchmark for tcgen05.mma on sm_100a: cycles per UMMA for SS / TS operand modes and N in
// {64, 128, 256}, with and without concurrent TMA-like shared-memory writes.  One CTA per SM; one thread issues a long
// dependency-free chain of MMAs that rotates through K-slices of shared-memory operand tiles exactly like the attention and
// GEMM main loops do; clock64 around (issue ... commit ... mbarrier wait).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/umma_microbench scripts/umma_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../scail_b200/csrc/sm100.cuh"

using namespace scail;

struct Case { int ts; int n; int m256; const char* name; };

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// TS: 0 = A,B from smem (K-major); 1 = A from TMEM, B K-major; 2 = A from TMEM, B MN-major (attention's P V)
// filler > 0: LSU st.shared traffic; filler < 0: a TMA warp streams -filler KB per 32 MMAs from global (L2) into a smem ring
template <int TS, int N>
__global__ void __launch_bounds__(256, 1) bench(long long* out, int iters, int filler, const uint8_t* gsrc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_smem = base;               // 128 x 128 bf16 (two 64-col halves, 32 KB)
    const uint32_t b_smem = base + 32768;       // up to 256 x 128 bf16 (64 KB)
    const uint32_t bar = base + 32768 + 65536 + 65536;
    const uint32_t slot = bar + 16;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (32768 + 65536 + 65536) / 4; i += blockDim.x)
        reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) { tmem_alloc<1>(slot, 512); tmem_relinquish<1>(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    if (warp == 0) {
        const bool leader = elect_one_sync();
        constexpr uint32_t idesc = umma_idesc_bf16(128, N, 0, 0);
        const uint64_t ad = umma_desc_kmajor_sw128(a_smem);
        const uint64_t bd = TS == 2 ? umma_desc_mnmajor_sw128(b_smem, 16384) : umma_desc_kmajor_sw128(b_smem);
        constexpr uint32_t idesc_mn = umma_idesc_bf16(128, N, 0, 1);
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (leader) {
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint64_t off = ((k >> 2) * 16384 + (k & 3) * 32) >> 4;
                    const uint64_t boff = ((k >> 2) * (N * 128) + (k & 3) * 32) >> 4;
                    if (TS == 2) umma_ts(tmem + 256, tmem + k * 8, bd + k * (2048 >> 4), idesc_mn, 1);
                    else if (TS) umma_ts(tmem + 256, tmem + k * 8, bd + boff, idesc, 1);
                    else umma_ss<1>(tmem + 256 * 0 + (it & 1) * 0, ad + off, bd + boff, idesc, 1);
                }
            }
        }
        if (leader) umma_commit(bar);
        mbar_wait(bar, 0);
        long long t1 = clock64();
        if (lane == 0) out[blockIdx.x] = t1 - t0;
    } else if (filler < 0 && warp == 1) {
        // TMA traffic: per 32 MMAs (= 4 iterations), -filler KB in 16 KB bulk copies into a 64 KB ring; completion tracked on a
        // second mbarrier that this lane polls before reusing a slot (keeps at most 4 copies in flight)
        if (lane == 0) {
            const uint32_t ring = base + 32768 + 65536, tbar = bar + 8;
            mbar_init(tbar, 1);
            fence_barrier_init();
            const int chunks = (-filler) / 16;
            uint32_t phase = 0;
            int slot_i = 0;
            for (int it = 0; it < iters / 4; ++it) {
                mbar_expect_tx(tbar, chunks * 16384);
                for (int c = 0; c < chunks; ++c) {
                    bulk_g2s(ring + (slot_i & 3) * 16384, gsrc + ((blockIdx.x * 7 + slot_i) & 63) * 16384, 16384, tbar);
                    ++slot_i;
                }
                mbar_wait(tbar, phase);
                phase ^= 1;
            }
        }
    } else if (filler > 0 && warp >= 4) {
        // emulate TMA fill traffic: plain shared-memory stores into a scratch region at ~filler bytes per 64 cycles
        const uint32_t scratch = base + 32768 + 65536;
        const int tid = threadIdx.x - 128;
        for (int it = 0; it < iters * 8; ++it) {
            for (int r = 0; r < filler; ++r)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %2, %2};" ::"r"(scratch + (((tid + r * 128) & 4095) << 4)), "r"(it), "r"(r) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<1>(tmem, 512); }
}

template <int TS, int N>
void run(const char* name, int filler) {
    const int iters = 2000, smem = 32768 + 65536 + 65536 + 1024 + 64;
    long long* d;
    cudaMalloc(&d, 148 * sizeof(long long));
    static uint8_t* gsrc = nullptr;
    if (!gsrc) { cudaMalloc(&gsrc, 64 * 16384); cudaMemset(gsrc, 0x3c, 64 * 16384); }
    cudaFuncSetAttribute(bench<TS, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    bench<TS, N><<<148, 256, smem>>>(d, 100, filler, gsrc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<TS, N><<<148, 256, smem>>>(d, iters, filler, gsrc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(148);
    cudaMemcpy(h.data(), d, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto v : h) avg += v;
    avg /= 148;
    const double per = avg / (iters * 8.0);
    const double ideal = 128.0 * N / 256.0;
    const double flops = 148.0 * iters * 8 * 2.0 * 128 * N * 16;
    printf("{\"case\": \"%s\", \"filler\": %d, \"cycles_per_mma\": %.1f, \"ideal\": %.0f, \"eff\": %.3f, \"ms\": %.3f, \"tflops\": %.0f, \"err\": \"%s\"}\n",
           name, filler, per, ideal, ideal / per, ms, flops / ms / 1e9, cudaGetErrorString(err));
    cudaFree(d);
}

int main() {
    run<0, 64>("SS 128x64x16", 0);
    run<0, 128>("SS 128x128x16", 0);
    run<0, 256>("SS 128x256x16", 0);
    run<1, 64>("TS 128x64x16", 0);
    run<1, 128>("TS 128x128x16", 0);
    run<1, 256>("TS 128x256x16", 0);
    run<2, 128>("TS-MNmajorB 128x128x16", 0);
    run<0, 128>("SS 128x128x16 + TMA 32KB/32mma", -32);
    run<0, 128>("SS 128x128x16 + TMA 64KB/32mma", -64);
    run<0, 128>("SS 128x128x16 + TMA 128KB/32mma", -128);
    run<1, 128>("TS 128x128x16 + TMA 64KB/32mma", -64);
    run<2, 128>("TS-MNmajorB 128x128x16 + TMA 64KB/32mma", -64);
    run<0, 256>("SS 128x256x16 + TMA 48KB/32mma", -48);
    run<0, 256>("SS 128x256x16 + TMA 96KB/32mma", -96);
    run<0, 256>("SS 128x256x16 + TMA 384KB/32mma (GEMM ratio)", -384);
    run<0, 128>("SS 128x128x16 + TMA 256KB/32mma", -256);
    return 0;
}
