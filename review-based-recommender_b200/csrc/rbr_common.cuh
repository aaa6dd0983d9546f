// rbr_common.cuh — shared helpers for the rbr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/rbr_b200.h"

namespace rbr {

// ---- error reporting -------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define RBR_REQUIRE(cond, code, ...)                    \
    do {                                                \
        if (!(cond)) {                                  \
            rbr::set_error(__VA_ARGS__);                \
            return (code);                              \
        }                                               \
    } while (0)

#define RBR_CUDA(call)                                                   \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) return rbr::cuda_fail(e__, #call);       \
    } while (0)

void count_launch();
#define RBR_LAUNCH_CHECK(name)                                           \
    do {                                                                 \
        rbr::count_launch();                                             \
        cudaError_t e__ = cudaGetLastError();                            \
        if (e__ != cudaSuccess) return rbr::cuda_fail(e__, name);        \
    } while (0)

// out-of-range id counter (ids outside [0, rows) are treated as padding rows and counted).
// One counter per translation unit (no relocatable device code needed); each .cu that can count defines
// an accessor with RBR_DEFINE_OOB_ACCESSOR and api.cu sums them in rbr_consume_oob_count().
static __device__ unsigned int g_oob_count = 0;
__device__ __forceinline__ void note_oob() { atomicAdd(&g_oob_count, 1u); }
#define RBR_DEFINE_OOB_ACCESSOR(name)                                                                              \
    namespace rbr {                                                                                                \
    int oob_consume_##name(cudaStream_t s, unsigned int* host_out) {                                               \
        unsigned int zero = 0;                                                                                     \
        if (cudaMemcpyFromSymbolAsync(host_out, g_oob_count, sizeof(unsigned int), 0, cudaMemcpyDeviceToHost, s) != \
            cudaSuccess)                                                                                           \
            return RBR_ECUDA;                                                                                      \
        if (cudaMemcpyToSymbolAsync(g_oob_count, &zero, sizeof(zero), 0, cudaMemcpyHostToDevice, s) != cudaSuccess) \
            return RBR_ECUDA;                                                                                      \
        return cudaStreamSynchronize(s) == cudaSuccess ? RBR_OK : RBR_ECUDA;                                       \
    }                                                                                                              \
    }

// ---- packed conv weight buffer layout (rbr_conv_pack) -------------------------------------------
// [0]            fp32  Wkeh  [k][E][Hpad4]      forward fp32 conv (filter index fastest)
// [off_hke]      fp32  Whke  [H][k][Epad4]      backward (embedding index fastest)
// [off_umma]     bf16  UMMA B operand, K-major no-swizzle core-matrix tiles:
//                       [P passes][k][Epad16/8 chunks][Nb rows][8 bf16]   (filter h = pass*Nb + row)
// (P, Nb) come from tc_pass_split (conv_tc.cu): the largest filter block whose weights stay resident in
// shared memory next to the activation ring; P == 0 means the tensor-core variant cannot take this shape.
void tc_pass_split(int64_t E, int64_t H, int64_t K, int64_t* P, int64_t* Nb);
// [off_umma2]    bf16  B operand of the CTA-pair kernel (conv_tc2.cu): each CTA of a pair keeps HALF of a pass's filters:
//                       [P2 passes][2 halves][k][Epad16/8 chunks][Nb2/2 rows][8 bf16]   (filter h = pass*Nb2 + half*Nb2/2 + row)
void tc2_pass_split(int64_t E, int64_t H, int64_t K, int64_t* P, int64_t* Nb);
struct PackLayout {
    int64_t E, H, k;
    int64_t P, Nb;      // tensor-core filter passes and filters per pass
    int64_t P2, Nb2;    // the same for the CTA-pair kernel (P2 == 0: unavailable)
    int64_t off_umma2;
    int64_t off_hke16;  // bf16 copy of Whke [H][k][Epad4] (table-gradient kernel in bf16 mode)
    int64_t Hpad4;      // H rounded up to 4
    int64_t Epad4;      // E rounded up to 4
    int64_t Epad16;     // E rounded up to 16 (UMMA K granularity for bf16)
    int64_t Npad;       // P * Nb
    int64_t off_keh, off_hke, off_umma, off_zero, total;   // off_zero: a zero row of rbr_emb_pad(E) bf16 (conv_tc zero source)
    // dense tensor-core backward (conv_bwd_tc.cu): bf16 Wt2 [NT * n_tiles rows = embedding index e][2 * HJp columns]:
    // Wt2[e][c] = W[h][e][j] for c = h*k + j and again for c = HJp + h*k + j (the K-major B operand of the table-gradient GEMM,
    // stacked twice because the coefficient matrix comes as bf16 hi | lo halves); zero elsewhere
    int64_t HJp;        // H * k rounded up to 64
    int64_t NT, n_tiles;
    int64_t off_wt2;
};
__host__ __device__ inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline int64_t cmat_hjp(int64_t H, int64_t k) { return round_up(H * k, 64); }
inline void cmat_ntile(int64_t E, int64_t* NT, int64_t* tiles) {
    *NT = E > 128 ? 160 : round_up(E, 16);            // UMMA N of the table-gradient GEMM (M = 128: N % 16 == 0, N <= 256)
    *tiles = (E + *NT - 1) / *NT;
}
inline PackLayout pack_layout(int64_t E, int64_t H, int64_t k) {
    PackLayout p;
    p.E = E; p.H = H; p.k = k;
    p.Hpad4 = round_up(H, 4);
    p.Epad4 = round_up(E, 4);
    p.Epad16 = round_up(E, 16);
    tc_pass_split(E, H, k, &p.P, &p.Nb);
    p.Npad = p.P * p.Nb;
    p.off_keh = 0;
    int64_t b = k * E * p.Hpad4 * 4;
    p.off_hke = round_up(b, 256);
    b = p.off_hke + H * k * p.Epad4 * 4;
    p.off_umma = round_up(b, 256);
    b = p.off_umma + k * p.Epad16 * p.Npad * 2;
    p.off_zero = round_up(b, 256);
    b = p.off_zero + round_up(E, 64) * 2 + 256;
    p.off_umma2 = round_up(b, 256);
    tc2_pass_split(E, H, k, &p.P2, &p.Nb2);
    b = p.off_umma2 + k * p.Epad16 * p.P2 * p.Nb2 * 2;
    p.off_hke16 = round_up(b, 256);
    b = p.off_hke16 + H * k * p.Epad4 * 2;
    p.off_wt2 = round_up(b, 1024);
    p.HJp = cmat_hjp(H, k);
    cmat_ntile(E, &p.NT, &p.n_tiles);
    b = p.off_wt2 + p.NT * p.n_tiles * 2 * p.HJp * 2;
    p.total = round_up(b, 256);
    return p;
}

// ---- small device helpers ----------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// counter-based keep mask for FM dropout: splitmix64 of (seed, index) → uniform in [0,1)
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (idx + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (float)(z >> 40) * (1.0f / 16777216.0f);
}

// relu written as a select so that NaN propagates like torch.relu (fmaxf would return 0 for NaN)
__device__ __forceinline__ float act_apply(int act, float x) { return act == RBR_ACT_RELU ? (x < 0.f ? 0.f : x) : tanhf(x); }
// derivative expressed through the activation OUTPUT y (relu: y>0, tanh: 1-y^2)
__device__ __forceinline__ float act_grad_from_out(int act, float y) {
    return act == RBR_ACT_RELU ? (y > 0.f ? 1.f : 0.f) : (1.f - y * y);
}

// token ids are int64 (torch.LongTensor) or int32 (RBR_IDS_I32: the staged input pipeline); a NULL mask means "all true"
// or, with RBR_MASK_FROM_IDS, mask = (id != 0) — what the reference's collate_fn computes on the host (utils.py:30-42)
struct IdView {
    const void* p;
    int i32;         // ids are int32
    int u16;         // ids are uint16
    int mask_ids;    // NULL mask → id != 0
};
__host__ __device__ inline IdView id_view(const void* ids, int flags) {
    IdView v;
    v.p = ids; v.i32 = (flags & RBR_IDS_I32) ? 1 : 0; v.u16 = (flags & RBR_IDS_U16) ? 1 : 0; v.mask_ids = (flags & RBR_MASK_FROM_IDS) ? 1 : 0;
    return v;
}
__device__ __forceinline__ int64_t ld_id(const IdView& v, int64_t i) {
    if (v.u16) return (int64_t)__ldg(reinterpret_cast<const uint16_t*>(v.p) + i);
    return v.i32 ? (int64_t)__ldg(reinterpret_cast<const int32_t*>(v.p) + i) : __ldg(reinterpret_cast<const int64_t*>(v.p) + i);
}
__device__ __forceinline__ bool ld_mask(const IdView& v, const uint8_t* mask, int64_t i, int64_t id) {
    return mask ? (__ldg(mask + i) != 0) : (v.mask_ids ? id != 0 : true);
}

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

}  // namespace rbr
