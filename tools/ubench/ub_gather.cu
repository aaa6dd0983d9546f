// ub_gather.cu — microbenchmark behind the conv_tc2 producer design: how fast can one SM pull random 640-byte table rows
// from L2 into a SWIZZLE_128B K-major tile, (a) with TMA tile::gather4, (b) with 16-byte cp.async, (c) with both at once?
// One CTA per SM, stage = 128 rows x 128 B (one K block), consumer = one warp that recycles the stage immediately.
//   build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I../../review-based-recommender_b200/csrc ub_gather.cu -o ub_gather -lcuda
//   run:   ./ub_gather            (prints one line per configuration)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "tc_ptx.cuh"
using namespace rbr;

constexpr int NST = 8, ROWS = 128, STAGE = ROWS * 128, NKB = 5;

__device__ __forceinline__ void tma_gather4_1(uint32_t dst, const void* tmap, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst), "l"(tmap), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar)
        : "memory");
}

// RT rows by TMA (4 warps), the other 128 - RT rows by LW cp.async warps
template <int LW, int DEPTH>
__global__ void __launch_bounds__((5 + LW) * 32, 1)
gather_kernel(const __grid_constant__ CUtensorMap tmap, const __nv_bfloat16* __restrict__ table, const int* __restrict__ ids, int tiles,
              int RT, int emb_pad) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t sbase = (raw + 1023u) & ~1023u;
    const uint32_t bar_full = sbase + NST * STAGE, bar_empty = bar_full + 8 * NST;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lsu_rows = ROWS - RT;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST; ++i) { mbar_init(bar_full + 8 * i, (RT > 0 ? 1 : 0) + (lsu_rows > 0 ? LW : 0)); mbar_init(bar_empty + 8 * i, 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int* my_ids = ids + (size_t)blockIdx.x * tiles * ROWS;
    if (warp < 4) {                                             // TMA producers: warp w owns rows [w*RT/4, (w+1)*RT/4)
        if (RT == 0) return;
        const int rpw = RT / 4, gpw = rpw / 4;
        int stage = 0; uint32_t ph = 0;
        for (int g = 0; g < tiles; ++g) {
            int mine = lane < rpw ? __ldg(my_ids + g * ROWS + warp * rpw + lane) : -1;
            int idx[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) idx[i] = __shfl_sync(0xffffffffu, mine, i);
            for (int kb = 0; kb < NKB; ++kb) {
                mbar_wait(bar_empty + 8 * stage, ph ^ 1);
                const uint32_t fb = bar_full + 8 * stage;
                if (warp == 0 && lane == 0) mbar_expect_tx(fb, (uint32_t)RT * 128u);
                if (elect_one()) {
#pragma unroll
                    for (int gi = 0; gi < 8; ++gi)
                        if (gi < gpw)
                            tma_gather4_1(sbase + stage * STAGE + (uint32_t)(warp * gpw + gi) * 512u, &tmap, kb * 64, idx[4 * gi], idx[4 * gi + 1],
                                          idx[4 * gi + 2], idx[4 * gi + 3], fb);
                }
                __syncwarp();
                if (++stage == NST) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp < 4 + LW) {                                 // cp.async producers: thread -> piece j of rows RT + q + (32*LW/8)*i
        if (lsu_rows == 0) return;
        const int lw = warp - 4, t = lw * 32 + lane;
        const int j = t & 7, q = t >> 3, rstep = LW * 4;        // LW*32/8 rows per sweep
        const int sweeps = (lsu_rows + rstep - 1) / rstep;      // <= 8 for LW = 4
        int stage = 0; uint32_t ph = 0;
        int pend_stage[DEPTH]; int n_pend = 0;
        for (int g = 0; g < tiles; ++g) {
            int rid[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = RT + q + rstep * i;
                rid[i] = (i < sweeps && r < ROWS) ? __ldg(my_ids + g * ROWS + r) : -1;
            }
            for (int kb = 0; kb < NKB; ++kb) {
                mbar_wait(bar_empty + 8 * stage, ph ^ 1);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = RT + q + rstep * i;
                    if (i < sweeps && r < ROWS) {
                        const uint32_t dst = sbase + stage * STAGE + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
                        const bool ok = rid[i] >= 0;
                        cp_async16(dst, table + (size_t)(ok ? rid[i] : 0) * emb_pad + kb * 64 + j * 8, ok ? 16u : 0u);
                    }
                }
                cp_async_commit();
                pend_stage[n_pend % DEPTH] = stage; ++n_pend;
                if (n_pend >= DEPTH) {
                    cp_async_wait<DEPTH - 1>();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * pend_stage[(n_pend - DEPTH) % DEPTH]);
                }
                if (++stage == NST) { stage = 0; ph ^= 1; }
            }
        }
        cp_async_wait_all();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0)
            for (int k = (n_pend >= DEPTH ? n_pend - DEPTH + 1 : 0); k < n_pend; ++k) mbar_arrive(bar_full + 8 * pend_stage[k % DEPTH]);
    } else {                                                    // consumer
        int stage = 0; uint32_t ph = 0;
        for (int g = 0; g < tiles * NKB; ++g) {
            mbar_wait(bar_full + 8 * stage, ph);
            if (lane == 0) mbar_arrive(bar_empty + 8 * stage);
            __syncwarp();
            if (++stage == NST) { stage = 0; ph ^= 1; }
        }
    }
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int LW, int DEPTH>
static void run(const CUtensorMap& tm, const __nv_bfloat16* table, const int* ids, int tiles, int RT, int emb_pad, int sms, double ghz) {
    auto k = gather_kernel<LW, DEPTH>;
    const int smem = NST * STAGE + 16 * NST + 1024;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k<<<sms, (5 + LW) * 32, smem>>>(tm, table, ids, tiles, RT, emb_pad);
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const double bytes = (double)sms * tiles * NKB * STAGE;
    printf("RT=%3d (TMA rows) LSU rows=%3d warps=%d depth=%d : %.3f ms  %.2f TB/s  %.1f B/clk/SM @%.2f GHz  (%.2f clk per 128-B row segment)\n", RT,
           128 - RT, LW, DEPTH, best, bytes / best / 1e9, bytes / sms / (best * 1e-3) / (ghz * 1e9), ghz,
           (best * 1e-3 * ghz * 1e9) / ((double)tiles * NKB * ROWS));
}

int main() {
    int sms = 0, khz = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const double ghz = khz / 1e6;
    const int V = 50000, emb_pad = 320, tiles = 110;
    std::vector<__nv_bfloat16> h((size_t)V * emb_pad);
    for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16((float)(i % 13));
    std::vector<int> hid((size_t)sms * tiles * ROWS);
    unsigned s = 12345u;
    for (auto& x : hid) { s = s * 1664525u + 1013904223u; x = (int)((s >> 8) % V); }
    __nv_bfloat16* table; int* ids;
    CK(cudaMalloc(&table, h.size() * 2)); CK(cudaMalloc(&ids, hid.size() * 4));
    CK(cudaMemcpy(table, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ids, hid.data(), hid.size() * 4, cudaMemcpyHostToDevice));
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q));
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)emb_pad, (cuuint64_t)V}, strides[1] = {(cuuint64_t)emb_pad * 2};
    cuuint32_t box[2] = {64, 1}, es[2] = {1, 1};
    CUresult r = ((EncodeFn)f)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, table, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    printf("sms=%d clock=%.3f GHz (nominal), tiles/CTA=%d, stage=16 KB x %d\n", sms, ghz, tiles, NST);
    run<4, 3>(tm, table, ids, tiles, 128, emb_pad, sms, ghz);
    run<4, 3>(tm, table, ids, tiles, 0, emb_pad, sms, ghz);
    run<4, 6>(tm, table, ids, tiles, 0, emb_pad, sms, ghz);
    run<8, 3>(tm, table, ids, tiles, 0, emb_pad, sms, ghz);
    run<8, 6>(tm, table, ids, tiles, 0, emb_pad, sms, ghz);
    run<4, 3>(tm, table, ids, tiles, 96, emb_pad, sms, ghz);
    run<4, 3>(tm, table, ids, tiles, 80, emb_pad, sms, ghz);
    run<4, 3>(tm, table, ids, tiles, 64, emb_pad, sms, ghz);
    run<4, 6>(tm, table, ids, tiles, 64, emb_pad, sms, ghz);
    run<8, 3>(tm, table, ids, tiles, 64, emb_pad, sms, ghz);
    run<8, 3>(tm, table, ids, tiles, 48, emb_pad, sms, ghz);
    run<8, 3>(tm, table, ids, tiles, 32, emb_pad, sms, ghz);
    return 0;
}
