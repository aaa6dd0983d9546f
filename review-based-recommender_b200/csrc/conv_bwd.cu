// conv_bwd.cu — K2b: arg-max-sparse backward of the fused conv → activation → max-over-time encoder.
//
// Replaces what autograd runs for the reverse of NgramFeat.forward (reference
// models/deepconn/layers.py:123-136) and of the embedding lookup (layers.py:23):
// max_pool1d backward, relu backward, aten::convolution_backward (dgrad + wgrad + bias grad),
// masked_fill backward and aten::embedding_dense_backward.  The reference spends 2x the forward FLOPs
// on a grad_output that is 1 non-zero per (doc, filter); here the work is
//   entry (n, h, j):  coef = feat_grad[n,h] * act'(feat[n,h]),   t = argmax[n,h] + j - pad
//     weight_grad[h, :, j] += coef * x[n, t, :]                          (conv_bwd_weight_kernel)
//     table_grad[ids[n,t], :] += coef * W[h, :, j]                       (conv_bwd_table_kernel)
//     bias_grad[h] += coef                                               (once per (n,h))
// The table gradient is a warp-segmented scatter-add: entries are counting-sorted by token id
// (token_sort.cuh) and each warp reduces a chunk of same-token entries in registers before one vector
// atomic per (chunk, token) — no [N,L,E] gradient tensor ever exists.
#include <stdlib.h>

#include "rbr_common.cuh"
#include "token_sort.cuh"

namespace rbr {

// ---- entries: key (token id or -1) and coefficient per (doc, filter, tap) ------------------------------
__global__ void __launch_bounds__(256) conv_bwd_entries_kernel(
    const IdView ids, const uint8_t* __restrict__ mask, int64_t n_docs, int L, int H, int K, int pad,
    int64_t vocab, int64_t padding_idx, const float* __restrict__ feat, const int32_t* __restrict__ argmax,
    const float* __restrict__ feat_grad, int feat_ld, int act, const float* __restrict__ gate, int gate_mode,
    const float* __restrict__ pool_raw, float* __restrict__ gate_grad,
    int32_t* __restrict__ keys, float* __restrict__ coef) {
    const int64_t total = n_docs * H * K;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const int j = (int)(q % K);
        const int64_t nh = q / K;
        const int h = (int)(nh % H);
        const int64_t n = nh / H;
        const float y = feat[n * feat_ld + h];
        float g = feat_grad[n * feat_ld + h] * act_grad_from_out(act, y);
        const int ts = argmax[n * feat_ld + h];
        const int t = ts + j - pad;
        if (gate_mode) {
            // y = gate * conv_nobias(x) + bias, pool_raw = gate * conv_nobias(x)  →  d/d gate = g * pool_raw / gate (no
            // cancellation; a gate that saturated to exactly 0 contributes 0: its own derivative gate*(1-gate) is 0 downstream);
            // d/d x and d/d W scale by the gate
            const int64_t gi = gate_mode == 1 ? n * L + ts : n;
            const float gv = gate[gi];
            if (j == 0 && gate_grad && g != 0.f && gv != 0.f) atomicAdd(gate_grad + gi, g * (pool_raw[n * feat_ld + h] / gv));
            g *= gv;
        }
        int32_t key = -1;
        if (g != 0.f && t >= 0 && t < L) {
            const int64_t id = ld_id(ids, n * L + t);
            if (ld_mask(ids, mask, n * L + t, id) && id >= 0 && id < vocab && id != padding_idx) key = (int32_t)id;
        }
        keys[q] = key;
        coef[q] = g;
    }
}

// ---- table gradient: warp-segmented reduction over token-sorted entries ----------------------------------
// WB16: weight rows are read from the bf16 copy of [H][K][Epad4] (bf16 precision mode: the same rounded weights the forward
// multiplied by, half the bytes per entry — this kernel is bound by L1/L2 reads of weight rows, 1.2 KB each in fp32).
template <int NQ, bool WB16>
__global__ void __launch_bounds__(256) conv_bwd_table_kernel(const int32_t* __restrict__ order, const int32_t* __restrict__ keys,
                                                             const float* __restrict__ coef, int64_t n_entries, const int32_t* __restrict__ n_kept,
                                                             int H, int K, const float4* __restrict__ whke, int e4w /* Epad4/4 */,
                                                             int e4 /* E/4 */, float4* __restrict__ table_grad, int chunk) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    n_entries = min(n_entries, (int64_t)__ldg(n_kept));   // entries with key < 0 were dropped by the sort
    const int64_t i0 = warp * 32;                        // one lane-batch of 32 sorted entries per warp
    if (i0 >= n_entries) return;
    (void)chunk;
    // lane l resolves entry i0+l (entry id → token key, coefficient, weight row) with independent loads; the warp
    // then walks the 32 entries with shuffles, so only the weight-row loads sit on the per-entry critical path.
    int key_l = -1, row_l = 0;
    float g_l = 0.f;
    if (i0 + lane < n_entries) {
        const int ent = __ldg(order + i0 + lane);
        key_l = __ldg(keys + ent);
        g_l = __ldg(coef + ent);
        row_l = ent % (H * K);                           // ent = (n*H + h)*K + j → packed weight row h*K + j
    }
    float4 acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = -1;
#pragma unroll 4
    for (int d = 0; d < 32; ++d) {
        const int key = __shfl_sync(0xffffffffu, key_l, d);
        if (key < 0) break;                              // skipped entries are sorted last (warp-uniform)
        const float g = __shfl_sync(0xffffffffu, g_l, d);
        const int64_t wrow_i = (int64_t)__shfl_sync(0xffffffffu, row_l, d) * e4w;
        float4 w[NQ];
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = lane + 32 * q;
            w[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < e4) {
                if (WB16) {
                    const uint2 u = __ldg(reinterpret_cast<const uint2*>(whke) + wrow_i + c);      // 4 bf16
                    w[q] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xFFFF0000u), __uint_as_float(u.y << 16),
                                       __uint_as_float(u.y & 0xFFFF0000u));
                } else {
                    w[q] = __ldg(whke + wrow_i + c);
                }
            }
        }
        if (key != cur) {
            if (cur >= 0) {
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int c = lane + 32 * q;
                    if (c < e4) atomicAdd(table_grad + (int64_t)cur * e4 + c, acc[q]);
                    acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            cur = key;
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            acc[q].x = fmaf(g, w[q].x, acc[q].x); acc[q].y = fmaf(g, w[q].y, acc[q].y);
            acc[q].z = fmaf(g, w[q].z, acc[q].z); acc[q].w = fmaf(g, w[q].w, acc[q].w);
        }
    }
    if (cur >= 0) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = lane + 32 * q;
            if (c < e4) atomicAdd(table_grad + (int64_t)cur * e4 + c, acc[q]);
        }
    }
}

// ---- weight + bias gradient: one warp per (filter h, chunk of documents) -------------------------------
// acc[j][:] += coef * x[n, t*+j-pad, :]; x rows are read from the fp32 table (fp32 mode) or the bf16 shadow.
template <int K, int NQ, bool BF16>
__global__ void __launch_bounds__(128) conv_bwd_weight_kernel(
    const float* __restrict__ table, const __nv_bfloat16* __restrict__ shadow, int emb_pad16, int64_t vocab, int E,
    const IdView ids, const uint8_t* __restrict__ mask, int64_t n_docs, int L, int H, int pad,
    const float* __restrict__ feat, const int32_t* __restrict__ argmax, const float* __restrict__ feat_grad, int feat_ld,
    int act, const float* __restrict__ gate, int gate_mode, int docs_per_warp, float* __restrict__ dw_hke /* [H][K][Epad4] */,
    int epad4, float* __restrict__ bias_grad) {
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    const int h = blockIdx.x;
    const int64_t chunk_id = (int64_t)blockIdx.y * (blockDim.x >> 5) + wib;
    const int64_t n0 = chunk_id * docs_per_warp;
    if (n0 >= n_docs) return;
    const int64_t n1 = min(n0 + (int64_t)docs_per_warp, n_docs);
    const int e4 = E >> 2;
    float4 acc[K][NQ];
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[j][q] = make_float4(0.f, 0.f, 0.f, 0.f);
    float bsum = 0.f;
    // Documents are taken 32 at a time: lane l resolves document nb+l's (coefficient, arg-max → K token ids) with
    // independent loads (one latency round for the metadata, one for the ids) instead of a dependent chain per
    // document; the warp then walks the 32 documents with shuffles, K row loads in flight per document.
    for (int64_t nb = n0; nb < n1; nb += 32) {
        const int64_t n = nb + lane;
        float g_l = 0.f;
        int ts_l = 0;
        if (n < n1) {
            const float y = __ldg(feat + n * feat_ld + h);
            g_l = __ldg(feat_grad + n * feat_ld + h) * act_grad_from_out(act, y);
            const int ts = __ldg(argmax + n * feat_ld + h);
            ts_l = ts - pad;
            bsum += g_l;                                             // the bias gradient is not gated
            if (gate_mode) g_l *= gate[gate_mode == 1 ? n * L + ts : n];
        }
        int64_t id_l[K];
#pragma unroll
        for (int j = 0; j < K; ++j) {
            id_l[j] = -1;
            const int t = ts_l + j;
            if (g_l != 0.f && t >= 0 && t < L) {
                const int64_t id = ld_id(ids, n * L + t);
                if (ld_mask(ids, mask, n * L + t, id) && id >= 0 && id < vocab) id_l[j] = id;
            }
        }
        const int cnt = (int)min((int64_t)32, n1 - nb);
#pragma unroll 2
        for (int d = 0; d < cnt; ++d) {
            const float g = __shfl_sync(0xffffffffu, g_l, d);
            if (g == 0.f) continue;                                  // warp-uniform
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int64_t id = __shfl_sync(0xffffffffu, id_l[j], d);
                if (id < 0) continue;                                // warp-uniform
#pragma unroll
                for (int q = 0; q < NQ; ++q) {
                    const int c = lane + 32 * q;
                    if (c < e4) {
                        float4 x;
                        if (BF16) {
                            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(shadow + id * emb_pad16) + c);
                            const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
                            const __nv_bfloat162 hi = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
                            const float2 a = __bfloat1622float2(lo), b = __bfloat1622float2(hi);
                            x = make_float4(a.x, a.y, b.x, b.y);
                        } else {
                            x = __ldg(reinterpret_cast<const float4*>(table + id * E) + c);
                        }
                        acc[j][q].x = fmaf(g, x.x, acc[j][q].x); acc[j][q].y = fmaf(g, x.y, acc[j][q].y);
                        acc[j][q].z = fmaf(g, x.z, acc[j][q].z); acc[j][q].w = fmaf(g, x.w, acc[j][q].w);
                    }
                }
            }
        }
    }
    bsum = warp_sum(bsum);
    float4* dst = reinterpret_cast<float4*>(dw_hke + (int64_t)h * K * epad4);
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const int c = lane + 32 * q;
            if (c < e4) atomicAdd(dst + j * (epad4 >> 2) + c, acc[j][q]);
        }
    if (lane == 0 && bsum != 0.f) atomicAdd(bias_grad + h, bsum);
}

// generic (any E, any K) scalar variant of the weight gradient, used when E % 4 != 0 or K is large
__global__ void __launch_bounds__(128) conv_bwd_weight_scalar_kernel(
    const float* __restrict__ table, int64_t vocab, int E, const IdView ids,
    const uint8_t* __restrict__ mask, int64_t n_docs, int L, int H, int K, int pad, const float* __restrict__ feat,
    const int32_t* __restrict__ argmax, const float* __restrict__ feat_grad, int feat_ld, int act,
    const float* __restrict__ gate, int gate_mode, float* __restrict__ dw_hke, int epad4, float* __restrict__ bias_grad) {
    // one CTA per (filter, doc): thread-strided over e; atomics per element.  Slow path, small shapes only.
    const int h = blockIdx.y;
    const int64_t n = blockIdx.x;
    const float y = feat[n * feat_ld + h];
    float g = feat_grad[n * feat_ld + h] * act_grad_from_out(act, y);
    if (g == 0.f) return;
    if (threadIdx.x == 0) atomicAdd(bias_grad + h, g);
    if (gate_mode) g *= gate[gate_mode == 1 ? n * L + argmax[n * feat_ld + h] : n];
    const int ts = argmax[n * feat_ld + h] - pad;
    for (int j = 0; j < K; ++j) {
        const int t = ts + j;
        if (t < 0 || t >= L) continue;
        const int64_t id = ld_id(ids, n * L + t);
        if (!ld_mask(ids, mask, n * L + t, id)) continue;
        if (id < 0 || id >= vocab) continue;
        for (int e = threadIdx.x; e < E; e += blockDim.x)
            atomicAdd(dw_hke + ((int64_t)h * K + j) * epad4 + e, g * table[id * E + e]);
    }
}

// scalar variant of the table gradient for E % 4 != 0 (plain atomics, no sort)
__global__ void __launch_bounds__(128) conv_bwd_table_scalar_kernel(const int32_t* __restrict__ keys, const float* __restrict__ coef,
                                                                    int64_t n_entries, int H, int K, const float* __restrict__ whke,
                                                                    int epad4, int E, float* __restrict__ table_grad) {
    const int64_t ent = blockIdx.x;
    if (ent >= n_entries) return;
    const int key = keys[ent];
    if (key < 0) return;
    const float g = coef[ent];
    const float* wrow = whke + (ent % ((int64_t)H * K)) * epad4;
    for (int e = threadIdx.x; e < E; e += blockDim.x) atomicAdd(table_grad + (int64_t)key * E + e, g * wrow[e]);
}

// weight_grad[h][e][j] += dw_hke[h][j][e]
__global__ void conv_bwd_unpack_kernel(const float* __restrict__ dw_hke, int H, int E, int K, int epad4, float* __restrict__ wgrad) {
    const int64_t total = (int64_t)H * E * K;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const int j = (int)(q % K);
        const int e = (int)((q / K) % E);
        const int h = (int)(q / ((int64_t)K * E));
        const float v = dw_hke[((int64_t)h * K + j) * epad4 + e];
        if (v != 0.f) atomicAdd(wgrad + q, v);          // the document sides may run concurrently on two streams
    }
}

}  // namespace rbr

using namespace rbr;

struct ConvBwdWs {
    int32_t* keys;
    float* coef;
    float* dw_hke;
    void* sort;
    int64_t total;
};
static ConvBwdWs conv_bwd_ws(void* base, int64_t n_docs, int64_t H, int64_t K, int64_t E, int64_t vocab) {
    const int64_t ne = n_docs * H * K;
    const int64_t epad4 = round_up(E, 4);
    char* p = reinterpret_cast<char*>(base);
    int64_t off = 0;
    ConvBwdWs w;
    w.keys = reinterpret_cast<int32_t*>(p + off); off += round_up(ne * 4, 256);
    w.coef = reinterpret_cast<float*>(p + off); off += round_up(ne * 4, 256);
    w.dw_hke = reinterpret_cast<float*>(p + off); off += round_up(H * K * epad4 * 4, 256);
    w.sort = p + off; off += token_sort_workspace_bytes(ne, vocab);
    w.total = off;
    return w;
}

extern "C" int64_t rbr_conv_bwd_workspace_bytes(int64_t n_docs, int64_t filters, int64_t ksize, int64_t emb, int64_t vocab) {
    return conv_bwd_ws(nullptr, n_docs, filters, ksize, emb, vocab).total;
}

template <int K, bool BF16>
static int launch_weight(int nq, dim3 grid, cudaStream_t s, const float* table, const __nv_bfloat16* shadow, int emb_pad16,
                         int64_t vocab, int E, IdView ids, const uint8_t* mask, int64_t n_docs, int L, int H, int pad,
                         const float* feat, const int32_t* argmax, const float* feat_grad, int feat_ld, int act,
                         const float* gate, int gate_mode, int dpw, float* dw_hke, int epad4, float* bias_grad) {
#define RBR_W(NQ)                                                                                                     \
    conv_bwd_weight_kernel<K, NQ, BF16><<<grid, 128, 0, s>>>(table, shadow, emb_pad16, vocab, E, ids, mask, n_docs, L, H, \
                                                             pad, feat, argmax, feat_grad, feat_ld, act, gate, gate_mode, dpw, dw_hke, \
                                                             epad4, bias_grad)
    if (nq == 1) RBR_W(1); else if (nq == 2) RBR_W(2); else if (nq == 3) RBR_W(3); else RBR_W(4);
#undef RBR_W
    RBR_LAUNCH_CHECK("conv_bwd_weight_kernel");
    return RBR_OK;
}

extern "C" int rbr_conv_act_maxpool_bwd(int precision, int activation, const void* table, const void* shadow_bf16,
                                        int64_t vocab, int64_t emb, const void* ids_raw, const uint8_t* mask,
                                        const float* gate, int gate_mode, int64_t n_docs, int64_t doc_len,
                                        const void* packed, int64_t filters, int64_t ksize, int64_t pad, const float* feat,
                                        const int32_t* argmax, const float* feat_grad, const float* pool_raw,
                                        int64_t feat_ld, int64_t padding_idx, float* weight_grad, float* bias_grad, float* table_grad, float* gate_grad, void* ws,
                                        int64_t ws_bytes, int flags, void* stream) {
    RBR_REQUIRE(table && ids_raw && packed && feat && argmax && feat_grad, RBR_EINVAL, "conv_bwd: null pointer");
    const IdView ids = id_view(ids_raw, flags);
    RBR_REQUIRE((weight_grad != nullptr) == (bias_grad != nullptr), RBR_EINVAL, "conv_bwd: weight_grad and bias_grad go together");
    RBR_REQUIRE(weight_grad || table_grad || gate_grad, RBR_EINVAL, "conv_bwd: nothing to compute");
    const bool do_weight = weight_grad != nullptr;            // NULL weight/bias grads: only the table (+ gate) part
    const bool do_entries = table_grad != nullptr || gate_grad != nullptr;
    RBR_REQUIRE(gate_mode >= 0 && gate_mode <= 2 && (gate_mode == 0) == (gate == nullptr), RBR_EINVAL, "conv_bwd: gate / gate_mode mismatch");
    RBR_REQUIRE(gate_mode == 0 || !gate_grad || pool_raw, RBR_EINVAL, "conv_bwd: the gate gradient of a gated conv needs pool_raw");
    RBR_REQUIRE(gate_mode != 1 || ksize == 1, RBR_EUNSUPPORTED, "conv_bwd: per-token gate needs ksize == 1");
    RBR_REQUIRE(precision == RBR_PREC_FP32 || shadow_bf16, RBR_EINVAL, "conv_bwd: bf16 precision needs the shadow table");
    RBR_REQUIRE(n_docs >= 0 && doc_len > 0 && filters > 0 && ksize > 0 && emb > 0, RBR_EINVAL, "conv_bwd: bad sizes");
    if (n_docs == 0) return RBR_OK;
    const int64_t ne = n_docs * filters * ksize;
    RBR_REQUIRE(ne < (1ll << 31) && vocab < (1ll << 31), RBR_EUNSUPPORTED, "conv_bwd: more than 2^31 entries");
    RBR_REQUIRE(ws && ws_bytes >= rbr_conv_bwd_workspace_bytes(n_docs, filters, ksize, emb, vocab), RBR_EWORKSPACE,
                "conv_bwd: workspace too small");
    cudaStream_t s = as_stream(stream);
    const PackLayout pl = pack_layout(emb, filters, ksize);
    const float* whke = reinterpret_cast<const float*>(reinterpret_cast<const char*>(packed) + pl.off_hke);
    ConvBwdWs w = conv_bwd_ws(ws, n_docs, filters, ksize, emb, vocab);
    const int E = (int)emb, H = (int)filters, K = (int)ksize, L = (int)doc_len;
    const int epad4 = (int)pl.Epad4;
    const bool vec = (emb % 4 == 0) && emb <= 512 && (K == 1 || K == 2 || K == 3 || K == 4 || K == 5 || K == 7);
    const int nq = ((E >> 2) + 31) / 32;

    if (do_weight) RBR_CUDA(cudaMemsetAsync(w.dw_hke, 0, (size_t)H * K * epad4 * 4, s));
    if (do_entries) {
        int blocks = (int)((ne + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        conv_bwd_entries_kernel<<<blocks, 256, 0, s>>>(ids, mask, n_docs, L, H, K, (int)pad, vocab, padding_idx, feat, argmax,
                                                       feat_grad, (int)feat_ld, activation, gate, gate_mode, pool_raw, gate_grad,
                                                       w.keys, w.coef);
        RBR_LAUNCH_CHECK("conv_bwd_entries_kernel");
    }
    // ---- weight + bias gradient
    if (!do_weight) {
    } else if (vec) {
        const int dpw = 64;                                    // documents per warp
        const int warps_per_cta = 4;
        const int64_t chunks = (n_docs + dpw - 1) / dpw;
        dim3 grid((unsigned)H, (unsigned)((chunks + warps_per_cta - 1) / warps_per_cta));
        const bool bf = (precision == RBR_PREC_BF16);
        const __nv_bfloat16* sh = reinterpret_cast<const __nv_bfloat16*>(shadow_bf16);
        const float* tb = reinterpret_cast<const float*>(table);
        const int ep16 = (int)rbr_emb_pad(emb);
        int rc = RBR_OK;
#define RBR_WK(K_)                                                                                                         \
    rc = bf ? launch_weight<K_, true>(nq, grid, s, tb, sh, ep16, vocab, E, ids, mask, n_docs, L, H, (int)pad, feat, argmax, \
                                      feat_grad, (int)feat_ld, activation, gate, gate_mode, dpw, w.dw_hke, epad4, bias_grad) \
            : launch_weight<K_, false>(nq, grid, s, tb, sh, ep16, vocab, E, ids, mask, n_docs, L, H, (int)pad, feat, argmax, \
                                       feat_grad, (int)feat_ld, activation, gate, gate_mode, dpw, w.dw_hke, epad4, bias_grad)
        switch (K) {
            case 1: RBR_WK(1); break;
            case 2: RBR_WK(2); break;
            case 3: RBR_WK(3); break;
            case 4: RBR_WK(4); break;
            case 5: RBR_WK(5); break;
            default: RBR_WK(7); break;
        }
#undef RBR_WK
        if (rc != RBR_OK) return rc;
    } else {
        RBR_REQUIRE(H <= 65535, RBR_EUNSUPPORTED, "conv_bwd: too many filters for the scalar path");
        dim3 grid((unsigned)n_docs, (unsigned)H);
        conv_bwd_weight_scalar_kernel<<<grid, 128, 0, s>>>(reinterpret_cast<const float*>(table), vocab, E, ids, mask, n_docs, L,
                                                           H, K, (int)pad, feat, argmax, feat_grad, (int)feat_ld, activation,
                                                           gate, gate_mode, w.dw_hke, epad4, bias_grad);
        RBR_LAUNCH_CHECK("conv_bwd_weight_scalar_kernel");
    }
    if (do_weight) {
        const int64_t tot = (int64_t)H * E * K;
        int blocks = (int)((tot + 255) / 256);
        if (blocks > 148 * 4) blocks = 148 * 4;
        conv_bwd_unpack_kernel<<<blocks, 256, 0, s>>>(w.dw_hke, H, E, K, epad4, weight_grad);
        RBR_LAUNCH_CHECK("conv_bwd_unpack_kernel");
    }
    // ---- table gradient (skipped when the embedding is frozen: table_grad == NULL)
    if (table_grad) {
        if (vec && ((uintptr_t)table_grad % 16 == 0)) {
            TokenSort ts;
            int rc = token_sort(w.keys, ne, vocab, w.sort, ts, s);
            if (rc != RBR_OK) return rc;
            const int chunk = 32;
            const int64_t warps = (ne + chunk - 1) / chunk;
            const int blocks = (int)((warps * 32 + 255) / 256);
            const bool wb16 = precision == RBR_PREC_BF16;
            const float4* wsrc = wb16 ? reinterpret_cast<const float4*>(reinterpret_cast<const char*>(packed) + pl.off_hke16)
                                      : reinterpret_cast<const float4*>(whke);
#define RBR_T(NQ)                                                                                                            \
    if (wb16) conv_bwd_table_kernel<NQ, true><<<blocks, 256, 0, s>>>(ts.order, w.keys, w.coef, ne, ts.start + vocab, H, K, wsrc, \
                                                     epad4 >> 2, E >> 2, reinterpret_cast<float4*>(table_grad), chunk);      \
    else conv_bwd_table_kernel<NQ, false><<<blocks, 256, 0, s>>>(ts.order, w.keys, w.coef, ne, ts.start + vocab, H, K, wsrc, \
                                                     epad4 >> 2, E >> 2, reinterpret_cast<float4*>(table_grad), chunk)
            if (nq == 1) RBR_T(1); else if (nq == 2) RBR_T(2); else if (nq == 3) RBR_T(3); else RBR_T(4);
#undef RBR_T
            RBR_LAUNCH_CHECK("conv_bwd_table_kernel");
        } else {
            conv_bwd_table_scalar_kernel<<<(unsigned)ne, 128, 0, s>>>(w.keys, w.coef, ne, H, K, whke, epad4, E, table_grad);
            RBR_LAUNCH_CHECK("conv_bwd_table_scalar_kernel");
        }
    }
    return RBR_OK;
}
