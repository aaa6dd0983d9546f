// avgpool.cu — K9: the SimpleSiamese encoder's masked average pooling fused with the embedding gather:
//     out[n, :] = sum_t mask[n,t] * table[ids[n,t], :] / (sum_t mask[n,t] + 1e-8)
// Replaces word_embedding(...) → transpose → MaskedAvgPooling1d.forward (reference models/simple_siamese/layers.py:90-110,
// called from simple_siamese.py:59-64) without materialising the [N, T, E] embeddings.  (The reference's VariationalDropout,
// one mask per (document, embedding channel) shared by all time steps, commutes with this mean: the module applies it to
// the pooled rows.)  One warp per document; backward: table_grad[ids[n,t], :] += mask * grad[n, :] / (len + 1e-8).
#include "rbr_common.cuh"

namespace rbr {

template <int NQ>
__global__ void __launch_bounds__(256) masked_avg_pool_fwd_kernel(const float4* __restrict__ table, int64_t vocab, int e4, const IdView ids,
                                                                  const uint8_t* __restrict__ mask, int64_t n_docs, int L,
                                                                  float4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_docs) return;
    float4 acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    float len = 0.f;
    for (int t0 = 0; t0 < L; t0 += 32) {
        // lane l resolves token t0 + l (id, mask) once; the warp then walks the 32 tokens with shuffles
        int64_t my = -1;
        float on = 0.f;
        if (t0 + lane < L) {
            const int64_t id = ld_id(ids, n * L + t0 + lane);
            if (ld_mask(ids, mask, n * L + t0 + lane, id)) {
                on = 1.f;                                               // the mask counts even if the id is out of range
                if (id >= 0 && id < vocab) my = id; else note_oob();
            }
        }
        len += on;
        const int cnt = min(32, L - t0);
        for (int d = 0; d < cnt; ++d) {
            const int64_t id = __shfl_sync(0xffffffffu, my, d);
            if (id < 0) continue;                                       // warp-uniform
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int c = lane + 32 * q;
                if (c < e4) {
                    const float4 v = __ldg(table + id * e4 + c);
                    acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w;
                }
            }
        }
    }
    len = warp_sum(len) + 1e-8f;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int c = lane + 32 * q;
        if (c < e4) out[n * e4 + c] = make_float4(acc[q].x / len, acc[q].y / len, acc[q].z / len, acc[q].w / len);
    }
}

template <int NQ>
__global__ void __launch_bounds__(256) masked_avg_pool_bwd_kernel(const IdView ids, const uint8_t* __restrict__ mask, int64_t n_docs, int L,
                                                                  int64_t vocab, int e4, int64_t padding_idx,
                                                                  const float4* __restrict__ out_grad, float4* __restrict__ table_grad) {
    const int lane = threadIdx.x & 31;
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (n >= n_docs) return;
    float len = 0.f;
    for (int t = lane; t < L; t += 32) {
        const int64_t id = ld_id(ids, n * L + t);
        if (ld_mask(ids, mask, n * L + t, id)) len += 1.f;
    }
    len = warp_sum(len) + 1e-8f;
    float4 g[NQ];
    bool any = false;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const int c = lane + 32 * q;
        g[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c < e4) {
            const float4 v = __ldg(out_grad + n * e4 + c);
            g[q] = make_float4(v.x / len, v.y / len, v.z / len, v.w / len);
            any |= (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f);
        }
    }
    if (!__any_sync(0xffffffffu, any)) return;
    for (int t0 = 0; t0 < L; t0 += 32) {
        int64_t my = -1;
        if (t0 + lane < L) {
            const int64_t id = ld_id(ids, n * L + t0 + lane);
            if (ld_mask(ids, mask, n * L + t0 + lane, id) && id >= 0 && id < vocab && id != padding_idx) my = id;
        }
        const int cnt = min(32, L - t0);
        for (int d = 0; d < cnt; ++d) {
            const int64_t id = __shfl_sync(0xffffffffu, my, d);
            if (id < 0) continue;
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                const int c = lane + 32 * q;
                if (c < e4) atomicAdd(table_grad + id * e4 + c, g[q]);
            }
        }
    }
}

}  // namespace rbr

using namespace rbr;

extern "C" int rbr_masked_avg_pool_fwd(const float* table, int64_t vocab, int64_t emb, const void* ids, const uint8_t* mask, int64_t n_docs,
                                       int64_t doc_len, float* out, int flags, void* stream) {
    RBR_REQUIRE(table && ids && out, RBR_EINVAL, "rbr_masked_avg_pool_fwd: null pointer");
    RBR_REQUIRE(vocab > 0 && emb > 0 && n_docs >= 0 && doc_len > 0, RBR_EINVAL, "rbr_masked_avg_pool_fwd: bad sizes");
    RBR_REQUIRE(emb % 4 == 0 && emb <= 512 && (uintptr_t)table % 16 == 0 && (uintptr_t)out % 16 == 0, RBR_EUNSUPPORTED,
                "rbr_masked_avg_pool_fwd: emb must be a multiple of 4 (<= 512) and the buffers 16-byte aligned");
    if (n_docs == 0) return RBR_OK;
    const int e4 = (int)(emb / 4), nq = (e4 + 31) / 32;
    const unsigned blocks = (unsigned)((n_docs * 32 + 255) / 256);
    const IdView iv = id_view(ids, flags);
#define RBR_AP(NQ) masked_avg_pool_fwd_kernel<NQ><<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(table), vocab, e4, iv, \
                                                                                         mask, n_docs, (int)doc_len, reinterpret_cast<float4*>(out))
    if (nq == 1) RBR_AP(1); else if (nq == 2) RBR_AP(2); else if (nq == 3) RBR_AP(3); else RBR_AP(4);
#undef RBR_AP
    RBR_LAUNCH_CHECK("masked_avg_pool_fwd_kernel");
    return RBR_OK;
}

extern "C" int rbr_masked_avg_pool_bwd(const void* ids, const uint8_t* mask, int64_t n_docs, int64_t doc_len, int64_t vocab, int64_t emb,
                                       int64_t padding_idx, const float* out_grad, float* table_grad, int flags, void* stream) {
    RBR_REQUIRE(ids && out_grad && table_grad, RBR_EINVAL, "rbr_masked_avg_pool_bwd: null pointer");
    RBR_REQUIRE(emb % 4 == 0 && emb <= 512 && (uintptr_t)out_grad % 16 == 0 && (uintptr_t)table_grad % 16 == 0, RBR_EUNSUPPORTED,
                "rbr_masked_avg_pool_bwd: emb must be a multiple of 4 (<= 512) and the buffers 16-byte aligned");
    if (n_docs == 0) return RBR_OK;
    const int e4 = (int)(emb / 4), nq = (e4 + 31) / 32;
    const unsigned blocks = (unsigned)((n_docs * 32 + 255) / 256);
    const IdView iv = id_view(ids, flags);
#define RBR_AP(NQ) masked_avg_pool_bwd_kernel<NQ><<<blocks, 256, 0, as_stream(stream)>>>(iv, mask, n_docs, (int)doc_len, vocab, e4, padding_idx, \
                                                                                         reinterpret_cast<const float4*>(out_grad),           \
                                                                                         reinterpret_cast<float4*>(table_grad))
    if (nq == 1) RBR_AP(1); else if (nq == 2) RBR_AP(2); else if (nq == 3) RBR_AP(3); else RBR_AP(4);
#undef RBR_AP
    RBR_LAUNCH_CHECK("masked_avg_pool_bwd_kernel");
    return RBR_OK;
}

RBR_DEFINE_OOB_ACCESSOR(avgpool)
