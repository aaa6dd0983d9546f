// probe_umma.cu — hardware probe (B200) for the assumptions behind conv_tc v2:
//   P1  TMA tile::gather4 into SWIZZLE_128B shared memory: physical layout, destination at 512-byte offsets inside a
//       1024-byte swizzle atom, out-of-bounds row indices (-1 and >= rows) zero-fill and still complete the mbarrier.
//   P2  tcgen05.mma with a K-major SWIZZLE_128B A operand whose descriptor start address is advanced by j ROWS (j*128 B)
//       — the "tap shift" of the implicit-GEMM conv — with base_offset = 0 and with base_offset = j & 7.
//   P3  the same under cta_group::2 (CTA pair, M = 256, each CTA holding half of B's rows), with the peer CTA's TMA
//       completing on the leader's mbarrier and tcgen05.commit multicast to both CTAs.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o probe_umma probe_umma.cu   (no -lcuda needed)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { printf("no cuTensorMapEncodeTiled\n"); exit(2); }
    return reinterpret_cast<EncodeTiledFn>(fn);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tmbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ int* g_err;
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity, int code, int* err) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 400000000ll) { atomicExch(err, code); return false; }
    }
    return true;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int CTAS>
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tm, int col, int r0, int r1, int r2, int r3, uint32_t bar) {
    if (CTAS == 1)
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::1 [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(dst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar) : "memory");
    else
        asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.cta_group::2 [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                     ::"r"(dst), "l"(tm), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar & 0xFEFFFFFFu) : "memory");
}

template <int CTAS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot, uint32_t cols) {
    if (CTAS == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot), "r"(cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CTAS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void umma(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
    if (CTAS == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <int CTAS>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    if (CTAS == 1) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---------------------------------------------------------------- P1
__global__ void probe_gather4(const __grid_constant__ CUtensorMap tm, int r0, int r1, int r2, int r3, int col, uint16_t* out /*[2048]*/, int* err) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_s;
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t bar = smem_u32(&bar_s);
    uint8_t* tile = smem + (base - smem_u32(smem));
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tile[i] = 0xEE;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(bar, 1024);
        tma_gather4<1>(base, &tm, col, r0, r1, r2, r3, bar);              // rows 0-3 of atom 0
        tma_gather4<1>(base + 512, &tm, col, r3, r2, r1, r0, bar);        // rows 4-7 of atom 0 (512-byte offset)
    }
    mbar_wait(bar, 0, 11, err);
    __syncthreads();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<uint16_t*>(tile)[i];
}

// ---------------------------------------------------------------- P2 / P3
struct MmaArgs {
    const int32_t* rows;       // [CTAS][ROWS_STAGED] row index per staged tile row (may be OOB → zeros)
    const __nv_bfloat16* wpk;  // [CTAS][TAPS][8 chunks][NLOC][8]  no-swizzle K-major B, per CTA
    float* D;                  // [CTAS*128][N]
    int base_offset_mode;      // 0: base_offset = 0, 1: base_offset = j & 7
    int* err;
};
constexpr int TAPS = 3;
constexpr int ROWS_STAGED = 136;      // 34 gather4 groups
constexpr int KB = 64;                // one 128-byte K block

template <int CTAS, int N>
__global__ void __launch_bounds__(192, 1) probe_mma(const __grid_constant__ CUtensorMap tm, MmaArgs a) {
    constexpr int NLOC = N / CTAS;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t rank = CTAS == 1 ? 0u : cluster_ctarank();
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t a_s = base;                                  // [ROWS_STAGED][128 B] SW128
    const uint32_t b_s = base + ROWS_STAGED * 128;              // [TAPS][8][NLOC][16 B]
    uint8_t* b_ptr = smem + ROWS_STAGED * 128;
    const uint32_t bar_full = smem_u32(&bars[0]), bar_done = smem_u32(&bars[1]);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) { mbar_init(bar_full, CTAS); mbar_init(bar_done, 1); fence_barrier_init(); }
    if (warp == 1) tmem_alloc<CTAS>(smem_u32(&tmem_slot), 32);
    // B: generic-proxy stores, then make them visible to the async proxy (tcgen05.mma reads smem through it)
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.wpk) + (size_t)rank * TAPS * 8 * NLOC;
        uint4* dst = reinterpret_cast<uint4*>(b_ptr);
        for (int i = threadIdx.x; i < TAPS * 8 * NLOC; i += blockDim.x) dst[i] = src[i];
    }
    fence_proxy_async();
    tc_fence_before();
    if (CTAS == 1) __syncthreads(); else cluster_sync();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (threadIdx.x == 0) {
        // producer: 34 gather4 of 4 rows x 128 B each; both CTAs complete on the LEADER's barrier
        if (rank == 0) mbar_expect_tx(bar_full, (uint32_t)(CTAS * ROWS_STAGED * 128));
        else mbar_arrive_remote(bar_full, 0);
        const int32_t* rows = a.rows + rank * ROWS_STAGED;
        for (int g = 0; g < ROWS_STAGED / 4; ++g)
            tma_gather4<CTAS>(a_s + g * 512, &tm, 0, rows[4 * g], rows[4 * g + 1], rows[4 * g + 2], rows[4 * g + 3], bar_full);
    }
    if (warp == 2 && rank == 0) {
        if (lane == 0) {
            mbar_wait(bar_full, 0, 21, a.err);
            tc_fence_after();
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)((128 * CTAS) >> 4) << 24);
            for (int ks = 0; ks < KB / 16; ++ks) {
                for (int j = 0; j < TAPS; ++j) {
                    const uint32_t a_addr = a_s + j * 128 + ks * 32;
                    const uint64_t bo = a.base_offset_mode ? (uint64_t)(j & 7) : 0ull;
                    const uint64_t ad = (uint64_t)((a_addr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
                                        (bo << 49) | (2ull << 61);
                    const uint32_t b_addr = b_s + (j * 8 + ks * 2) * NLOC * 16;
                    const uint64_t bd = (uint64_t)((b_addr >> 4) & 0x3FFFu) | ((uint64_t)((NLOC * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
                                        (1ull << 46);
                    umma<CTAS>(tmem_base, ad, bd, idesc, (uint32_t)((ks | j) != 0));
                }
            }
            umma_commit<CTAS>(bar_done);
        }
        __syncwarp();
    }
    if (warp < 4) {
        mbar_wait(bar_done, 0, 31, a.err);
        tc_fence_after();
        for (int c0 = 0; c0 < N; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
            float* drow = a.D + ((size_t)rank * 128 + warp * 32 + lane) * N + c0;
            for (int i = 0; i < 16; ++i) drow[i] = __uint_as_float(v[i]);
        }
    }
    tc_fence_before();
    if (CTAS == 1) __syncthreads(); else cluster_sync();
    if (warp == 1) tmem_dealloc<CTAS>(tmem_base, 32);
}

static float bf(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }
static uint16_t tobf(float f) { __nv_bfloat16 h = __float2bfloat16(f); uint16_t b; memcpy(&b, &h, 2); return b; }

int main() {
    EncodeTiledFn encode = get_encode();
    const int V = 200, EP = 320;
    std::vector<uint16_t> tab((size_t)V * EP);
    srand(1);
    for (int v = 0; v < V; ++v)
        for (int e = 0; e < EP; ++e) tab[(size_t)v * EP + e] = tobf((float)((rand() % 255) - 127) / 64.f);
    uint16_t* d_tab; CK(cudaMalloc(&d_tab, tab.size() * 2)); CK(cudaMemcpy(d_tab, tab.data(), tab.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)EP, (cuuint64_t)V};
    cuuint64_t gstr[1] = {(cuuint64_t)EP * 2};
    cuuint32_t box[2] = {64, 1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_tab, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
    int* d_err; CK(cudaMalloc(&d_err, 4)); CK(cudaMemset(d_err, 0, 4));
    int h_err = 0;

    // ---- P1
    {
        uint16_t* d_out; CK(cudaMalloc(&d_out, 2048));
        const int rr[4] = {3, 17, -1, V};      // two valid rows, one negative, one == rows (both OOB)
        const int col = 64;
        probe_gather4<<<1, 128, 4096 + 1024>>>(tm, rr[0], rr[1], rr[2], rr[3], col, d_out, d_err);
        cudaError_t e = cudaDeviceSynchronize();
        CK(cudaMemcpy(&h_err, d_err, 4, cudaMemcpyDeviceToHost));
        printf("P1 gather4: sync=%s err=%d\n", cudaGetErrorString(e), h_err);
        if (e != cudaSuccess) return 3;
        std::vector<uint16_t> out(1024);
        CK(cudaMemcpy(out.data(), d_out, 2048, cudaMemcpyDeviceToHost));
        // smem row p (0..7) holds source row: p<4 → rr[p], else rr[7-p]
        int bad_sw = 0, bad_lin = 0;
        for (int p = 0; p < 8; ++p) {
            const int src = p < 4 ? rr[p] : rr[7 - p];
            for (int c = 0; c < 8; ++c)
                for (int i = 0; i < 8; ++i) {
                    const float want = (src >= 0 && src < V) ? bf(tab[(size_t)src * EP + col + c * 8 + i]) : 0.f;
                    const float got_sw = bf(out[p * 64 + ((c ^ (p & 7)) * 8) + i]);
                    const float got_lin = bf(out[p * 64 + c * 8 + i]);
                    bad_sw += got_sw != want;
                    bad_lin += got_lin != want;
                }
        }
        printf("P1 layout: mismatches assuming chunk^(row&7) swizzle = %d, assuming linear = %d  (OOB rows -1 and V must read 0)\n", bad_sw, bad_lin);
        cudaFree(d_out);
    }

    // ---- P2 / P3
    auto run_mma = [&](int ctas, int bom) {
        const int N = 32, NLOC = N / ctas;
        std::vector<int32_t> rows((size_t)ctas * ROWS_STAGED);
        for (size_t i = 0; i < rows.size(); ++i) rows[i] = rand() % V;
        rows[5] = -1; rows[77] = V; if (ctas == 2) rows[ROWS_STAGED + 9] = V + 5;
        // W[j][n][k], k < 64
        std::vector<float> W((size_t)TAPS * N * KB);
        for (auto& w : W) w = bf(tobf((float)((rand() % 255) - 127) / 128.f));
        std::vector<uint16_t> wpk((size_t)ctas * TAPS * 8 * NLOC * 8);
        for (int c = 0; c < ctas; ++c)
            for (int j = 0; j < TAPS; ++j)
                for (int ch = 0; ch < 8; ++ch)
                    for (int n = 0; n < NLOC; ++n)
                        for (int i = 0; i < 8; ++i)
                            wpk[((((size_t)c * TAPS + j) * 8 + ch) * NLOC + n) * 8 + i] = tobf(W[((size_t)j * N + c * NLOC + n) * KB + ch * 8 + i]);
        int32_t* d_rows; uint16_t* d_w; float* d_D;
        CK(cudaMalloc(&d_rows, rows.size() * 4)); CK(cudaMemcpy(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_w, wpk.size() * 2)); CK(cudaMemcpy(d_w, wpk.data(), wpk.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&d_D, (size_t)ctas * 128 * N * 4)); CK(cudaMemset(d_D, 0xFF, (size_t)ctas * 128 * N * 4));
        CK(cudaMemset(d_err, 0, 4));
        MmaArgs a{d_rows, reinterpret_cast<const __nv_bfloat16*>(d_w), d_D, bom, d_err};
        const size_t smem = 1024 + ROWS_STAGED * 128 + TAPS * 8 * NLOC * 16;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = ctas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e;
        if (ctas == 1) e = cudaLaunchKernelEx(&cfg, probe_mma<1, 32>, tm, a);
        else e = cudaLaunchKernelEx(&cfg, probe_mma<2, 32>, tm, a);
        if (e != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(e)); exit(3); }
        e = cudaDeviceSynchronize();
        CK(cudaMemcpy(&h_err, d_err, 4, cudaMemcpyDeviceToHost));
        std::vector<float> D((size_t)ctas * 128 * N);
        if (e == cudaSuccess) CK(cudaMemcpy(D.data(), d_D, D.size() * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        for (int c = 0; c < ctas; ++c)
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < N; ++n) {
                    double acc = 0;
                    for (int j = 0; j < TAPS; ++j) {
                        const int src = rows[(size_t)c * ROWS_STAGED + m + j];
                        if (src < 0 || src >= V) continue;
                        for (int k = 0; k < KB; ++k) acc += (double)bf(tab[(size_t)src * EP + k]) * W[((size_t)j * N + n) * KB + k];
                    }
                    const double got = D[((size_t)c * 128 + m) * N + n];
                    maxerr = fmax(maxerr, fabs(got - acc));
                    maxref = fmax(maxref, fabs(acc));
                }
        printf("P%d mma ctas=%d base_offset_mode=%d: sync=%s err=%d max|err|=%.4g (max|ref|=%.4g) → %s\n", ctas == 1 ? 2 : 3, ctas, bom,
               cudaGetErrorString(e), h_err, maxerr, maxref, (e == cudaSuccess && h_err == 0 && maxerr < 1e-2 * fmax(maxref, 1.0)) ? "OK" : "WRONG");
        if (e != cudaSuccess) exit(4);
        cudaFree(d_rows); cudaFree(d_w); cudaFree(d_D);
    };
    CK(cudaFuncSetAttribute(probe_mma<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(probe_mma<2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    CK(cudaFuncSetAttribute(probe_mma<2, 32>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    run_mma(1, 0);
    run_mma(1, 1);
    run_mma(2, 0);
    run_mma(2, 1);
    return 0;
}
