// hierpool.cu — K8: the reference's alternate encoder arch="HierPooling" (models/deepconn/layers.py:62-98, 110-114):
//     x = mask(table[ids])            [N, L, E]
//     y[n, t, e] = mean_{j<k} x[n, t+j, e]        F.avg_pool1d(kernel k, stride 1), t in [0, L-k+1)
//     pooled[n, e] = max_t y[n, t, e]             F.max_pool1d over the whole pooled length (first arg-max)
// followed (in the module, by library GEMMs) by the optional Linear(E → out) and the ReLU.  No weights: a gather plus a
// sliding-window reduction, bound by the row reads like K1 — the [N, L, E] embeddings, their masked and transposed copies and
// the [N, E, L-k+1] averages are never materialised.  Backward: the arg-max window of every (doc, channel) receives
// grad / k on its k token rows (scalar atomics into the dense table gradient).
#include "rbr_common.cuh"

namespace rbr {

constexpr int HP_MAXK = 8;

template <bool VEC>
__global__ void __launch_bounds__(128) hier_pool_fwd_kernel(const float* __restrict__ table, int64_t vocab, int E, const IdView ids,
                                                            const uint8_t* __restrict__ mask, int L, int K, float* __restrict__ pooled,
                                                            int32_t* __restrict__ argmax) {
    const int64_t n = blockIdx.x;
    constexpr int W = VEC ? 4 : 1;
    const int cols = VEC ? (E >> 2) : E;
    __shared__ int64_t row_s[128];                     // token rows of a chunk of positions (-1: reads as zeros)
    const float inv_k = 1.f / (float)K;
    for (int c = threadIdx.x; c < ((cols + 127) / 128) * 128; c += 128) {
        const bool live = c < cols;
        float ring[HP_MAXK][W];
        float best[W];
        int best_t[W];
#pragma unroll
        for (int w = 0; w < W; ++w) { best[w] = -INFINITY; best_t[w] = 0; }
        for (int t0 = 0; t0 < L; t0 += 128) {
            __syncthreads();
            {
                const int t = t0 + threadIdx.x;
                int64_t src = -1;
                if (t < L) {
                    const int64_t id = ld_id(ids, n * L + t);
                    if (ld_mask(ids, mask, n * L + t, id)) {
                        if (id >= 0 && id < vocab) src = id;
                        else if (c == (int)threadIdx.x) note_oob();
                    }
                }
                row_s[threadIdx.x] = src;
            }
            __syncthreads();
            if (!live) continue;
            const int tend = min(128, L - t0);
            for (int i = 0; i < tend; ++i) {
                const int t = t0 + i;
                const int64_t src = row_s[i];
                float x[W];
                if (src >= 0) {
                    if (VEC) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(table + src * E) + c);
                        x[0] = v.x; if (W > 1) { x[1 % W] = v.y; x[2 % W] = v.z; x[3 % W] = v.w; }
                    } else {
                        x[0] = __ldg(table + src * E + c);
                    }
                } else {
#pragma unroll
                    for (int w = 0; w < W; ++w) x[w] = 0.f;
                }
#pragma unroll
                for (int w = 0; w < W; ++w) {
#pragma unroll
                    for (int j = 0; j < HP_MAXK - 1; ++j) ring[j][w] = ring[j + 1][w];      // ring[HP_MAXK-1] = newest
                    ring[HP_MAXK - 1][w] = x[w];
                }
                if (t >= K - 1) {
#pragma unroll
                    for (int w = 0; w < W; ++w) {
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < HP_MAXK; ++j)
                            if (j >= HP_MAXK - K) s += ring[j][w];                          // oldest first, as avg_pool1d sums the window
                        const float v = s * inv_k;
                        if (!(v <= best[w]) && !(best[w] != best[w])) { best[w] = v; best_t[w] = t - (K - 1); }
                    }
                }
            }
        }
        if (live) {
#pragma unroll
            for (int w = 0; w < W; ++w) {
                pooled[n * E + c * W + w] = best[w];
                argmax[n * E + c * W + w] = best_t[w];
            }
        }
    }
}

__global__ void __launch_bounds__(256) hier_pool_bwd_kernel(const IdView ids, const uint8_t* __restrict__ mask, int64_t n_docs, int L,
                                                            int64_t vocab, int E, int K, int64_t padding_idx,
                                                            const int32_t* __restrict__ argmax, const float* __restrict__ pooled_grad,
                                                            float* __restrict__ table_grad) {
    const int64_t total = n_docs * E;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const float inv_k = 1.f / (float)K;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const float g = pooled_grad[q] * inv_k;
        if (g == 0.f) continue;
        const int64_t n = q / E;
        const int e = (int)(q - n * E);
        const int t0 = argmax[q];
        for (int j = 0; j < K; ++j) {
            const int t = t0 + j;
            if (t < 0 || t >= L) continue;
            const int64_t id = ld_id(ids, n * L + t);
            if (!ld_mask(ids, mask, n * L + t, id) || id < 0 || id >= vocab || id == padding_idx) continue;
            atomicAdd(table_grad + id * E + e, g);
        }
    }
}

}  // namespace rbr

using namespace rbr;

extern "C" int rbr_hier_pool_fwd(const float* table, int64_t vocab, int64_t emb, const void* ids, const uint8_t* mask, int64_t n_docs,
                                 int64_t doc_len, int64_t ksize, float* pooled, int32_t* argmax, int flags, void* stream) {
    RBR_REQUIRE(table && ids && pooled && argmax, RBR_EINVAL, "rbr_hier_pool_fwd: null pointer");
    RBR_REQUIRE(vocab > 0 && emb > 0 && n_docs >= 0 && doc_len > 0, RBR_EINVAL, "rbr_hier_pool_fwd: bad sizes");
    RBR_REQUIRE(ksize >= 1 && ksize <= HP_MAXK && ksize <= doc_len, RBR_EUNSUPPORTED, "rbr_hier_pool_fwd: kernel size must be in [1, %d] and <= doc_len",
                HP_MAXK);
    RBR_REQUIRE(n_docs <= 0x7fffffff, RBR_EUNSUPPORTED, "rbr_hier_pool_fwd: too many documents");
    if (n_docs == 0) return RBR_OK;
    const IdView iv = id_view(ids, flags);
    if (emb % 4 == 0 && (uintptr_t)table % 16 == 0)
        hier_pool_fwd_kernel<true><<<(unsigned)n_docs, 128, 0, as_stream(stream)>>>(table, vocab, (int)emb, iv, mask, (int)doc_len, (int)ksize,
                                                                                   pooled, argmax);
    else
        hier_pool_fwd_kernel<false><<<(unsigned)n_docs, 128, 0, as_stream(stream)>>>(table, vocab, (int)emb, iv, mask, (int)doc_len, (int)ksize,
                                                                                    pooled, argmax);
    RBR_LAUNCH_CHECK("hier_pool_fwd_kernel");
    return RBR_OK;
}

extern "C" int rbr_hier_pool_bwd(const void* ids, const uint8_t* mask, int64_t n_docs, int64_t doc_len, int64_t vocab, int64_t emb,
                                 int64_t ksize, int64_t padding_idx, const int32_t* argmax, const float* pooled_grad, float* table_grad,
                                 int flags, void* stream) {
    RBR_REQUIRE(ids && argmax && pooled_grad && table_grad, RBR_EINVAL, "rbr_hier_pool_bwd: null pointer");
    RBR_REQUIRE(ksize >= 1 && ksize <= HP_MAXK, RBR_EUNSUPPORTED, "rbr_hier_pool_bwd: kernel size must be in [1, %d]", HP_MAXK);
    if (n_docs == 0) return RBR_OK;
    const int64_t total = n_docs * emb;
    int64_t blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    hier_pool_bwd_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(id_view(ids, flags), mask, n_docs, (int)doc_len, vocab, (int)emb,
                                                                          (int)ksize, padding_idx, argmax, pooled_grad, table_grad);
    RBR_LAUNCH_CHECK("hier_pool_bwd_kernel");
    return RBR_OK;
}

RBR_DEFINE_OOB_ACCESSOR(hierpool)
