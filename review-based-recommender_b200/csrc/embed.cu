// embed.cu — K1 embedding-row gather, K1b warp-segmented dense embedding gradient, K0 operand staging.
//
// K1  replaces nn.Embedding forward   (reference models/deepconn/layers.py:22-24).
// K1b replaces aten::embedding_dense_backward (autograd reverse of layers.py:23): counting sort of the
//     token ids (histogram → scan → fill) followed by a warp-per-chunk segmented reduction, so a hot
//     token's rows are summed in registers and flushed with ONE vector atomic per (chunk, token).
#include "rbr_common.cuh"
#include "token_sort.cuh"

namespace rbr {

// ------------------------------------------------------------------------------------------------
// K1: out[t, :] = table[ids[t], :]   (bit-exact copy; HBM/L2-bandwidth bound)
// One thread moves one 16-byte piece per iteration; consecutive threads → consecutive pieces of the
// flat [n_tokens * E/4] output, so stores are fully coalesced and loads are coalesced per row.
// ------------------------------------------------------------------------------------------------
template <int UNROLL>
__global__ void __launch_bounds__(256) gather_rows_v4_kernel(const float4* __restrict__ table, int64_t vocab, int e4,
                                                             const int64_t* __restrict__ ids, int64_t n_tokens,
                                                             float4* __restrict__ out) {
    const int64_t total = n_tokens * (int64_t)e4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; q + (UNROLL - 1) * stride < total; q += UNROLL * stride) {
        float4 v[UNROLL];
        int64_t dst[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t p = q + u * stride;
            const int64_t tok = p / e4;
            const int c = (int)(p - tok * e4);
            const int64_t id = __ldg(ids + tok);
            dst[u] = p;
            if (id >= 0 && id < vocab) {
                v[u] = __ldg(table + id * e4 + c);
            } else {
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (c == 0) note_oob();
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) __stcs(out + dst[u], v[u]);   // streaming store: output is not re-read here
    }
    for (; q < total; q += stride) {
        const int64_t tok = q / e4;
        const int c = (int)(q - tok * e4);
        const int64_t id = __ldg(ids + tok);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (id >= 0 && id < vocab) v = __ldg(table + id * e4 + c);
        else if (c == 0) note_oob();
        __stcs(out + q, v);
    }
}

__global__ void __launch_bounds__(256) gather_rows_scalar_kernel(const float* __restrict__ table, int64_t vocab, int emb,
                                                                 const int64_t* __restrict__ ids, int64_t n_tokens,
                                                                 float* __restrict__ out) {
    const int64_t total = n_tokens * (int64_t)emb;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const int64_t tok = q / emb;
        const int c = (int)(q - tok * emb);
        const int64_t id = __ldg(ids + tok);
        float v = 0.f;
        if (id >= 0 && id < vocab) v = __ldg(table + id * emb + c);
        else if (c == 0) note_oob();
        out[q] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// K1b: table_grad[id, :] += sum of grad_rows over tokens with that id.
// `order` lists token positions grouped by id (token_sort.cuh).  Each warp owns CHUNK consecutive
// entries of `order`, keeps the running row sum in registers (lane owns float4 columns lane, lane+32, ...)
// and flushes with float4 atomics when the id changes or the chunk ends.
// ------------------------------------------------------------------------------------------------
template <int NQ>   // NQ = ceil(E/4 / 32) float4 accumulators per lane
__global__ void __launch_bounds__(256) embgrad_segment_reduce_kernel(const int32_t* __restrict__ order,
                                                                      const int32_t* __restrict__ keys,
                                                                      int64_t n_sorted, const int32_t* __restrict__ n_kept, const float4* __restrict__ grad_rows,
                                                                      int e4, float4* __restrict__ table_grad, int chunk) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    n_sorted = min(n_sorted, (int64_t)__ldg(n_kept));     // entries with key < 0 were dropped by the sort
    int64_t i = warp * chunk;
    if (i >= n_sorted) return;
    const int64_t end = min(i + (int64_t)chunk, n_sorted);
    float4 acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cur = -1;
    for (; i < end; ++i) {
        const int tok = __ldg(order + i);
        const int key = __ldg(keys + tok);
        if (key < 0) break;                       // skipped entries (padding / out of range) are sorted last
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
            const int c = lane + 32 * q;
            if (c < e4) {
                const float4 g = __ldcs(grad_rows + (int64_t)tok * e4 + c);
                acc[q].x += g.x; acc[q].y += g.y; acc[q].z += g.z; acc[q].w += g.w;
            }
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

// scalar fallback for emb % 4 != 0: one thread per (token, column) with plain atomics
__global__ void embgrad_atomic_scalar_kernel(const int64_t* __restrict__ ids, const float* __restrict__ grad_rows,
                                             int64_t n_tokens, int emb, int64_t vocab, int64_t padding_idx,
                                             float* __restrict__ table_grad) {
    const int64_t total = n_tokens * (int64_t)emb;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const int64_t tok = q / emb;
        const int c = (int)(q - tok * emb);
        const int64_t id = ids[tok];
        if (id < 0 || id >= vocab) { if (c == 0) note_oob(); continue; }
        if (id == padding_idx) continue;
        atomicAdd(table_grad + id * emb + c, grad_rows[q]);
    }
}

// keys for the sort: token id, or -1 for padding / out-of-range
__global__ void embgrad_keys_kernel(const int64_t* __restrict__ ids, int64_t n, int64_t vocab, int64_t padding_idx,
                                    int32_t* __restrict__ keys) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int64_t id = ids[i];
        int32_t k = -1;
        if (id < 0 || id >= vocab) note_oob();
        else if (id != padding_idx) k = (int32_t)id;
        keys[i] = k;
    }
}

// ------------------------------------------------------------------------------------------------
// K0: bf16 shadow table, rows zero-padded to emb_pad (multiple of 16 → UMMA K granularity)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) table_to_bf16_kernel(const float* __restrict__ table, int64_t vocab, int emb, int emb_pad,
                                                            __nv_bfloat16* __restrict__ shadow) {
    const int pieces = emb_pad >> 3;                                  // 8 bf16 = 16 B per piece
    const int64_t total = vocab * (int64_t)pieces;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        const int64_t row = q / pieces;
        const int c0 = (int)(q - row * pieces) * 8;
        const float* src = table + row * emb + c0;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = (c0 + j < emb) ? __ldg(src + j) : 0.f;
        uint4 pk;
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(f[0], f[1]); pk.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(f[2], f[3]); pk.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(f[4], f[5]); pk.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(f[6], f[7]); pk.w = *reinterpret_cast<uint32_t*>(&t);
        *reinterpret_cast<uint4*>(shadow + row * emb_pad + c0) = pk;
    }
}

// K0: conv weight packing, weight [H][E][k] (nn.Conv1d) → the three layouts of PackLayout
__global__ void conv_pack_kernel(const float* __restrict__ w, int E, int H, int k, int Hpad4, int Epad4, int Epad16, int Npad, int Nb,
                                 int Npad2, int Nb2, float* __restrict__ keh, float* __restrict__ hke,
                                 __nv_bfloat16* __restrict__ umma, __nv_bfloat16* __restrict__ umma2, __nv_bfloat16* __restrict__ hke16,
                                 __nv_bfloat16* __restrict__ wt2, int HJp, int wt2_rows) {
    const int64_t n_keh = (int64_t)k * E * Hpad4;
    const int64_t n_hke = (int64_t)H * k * Epad4;
    const int64_t n_umma = (int64_t)k * Epad16 * Npad;
    const int64_t n_umma2 = (int64_t)k * Epad16 * Npad2;
    const int64_t n_wt2 = (int64_t)wt2_rows * 2 * HJp;
    const int64_t total = n_keh + n_hke + n_umma + n_umma2 + n_wt2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += stride) {
        if (q >= n_keh + n_hke + n_umma + n_umma2) {
            // Wt2[e][c]: c < HJp → filter-tap c, c >= HJp → filter-tap c - HJp (hi | lo halves of the coefficient matrix)
            const int64_t r = q - (n_keh + n_hke + n_umma + n_umma2);
            const int e = (int)(r / (2 * HJp));
            int c = (int)(r - (int64_t)e * 2 * HJp);
            if (c >= HJp) c -= HJp;
            const int h = c / k, j = c - h * k;
            wt2[r] = __float2bfloat16_rn((e < E && h < H) ? w[((int64_t)h * E + e) * k + j] : 0.f);
        } else if (q < n_keh) {
            const int h = (int)(q % Hpad4);
            const int e = (int)((q / Hpad4) % E);
            const int j = (int)(q / ((int64_t)Hpad4 * E));
            keh[q] = (h < H) ? w[((int64_t)h * E + e) * k + j] : 0.f;
        } else if (q < n_keh + n_hke) {
            const int64_t r = q - n_keh;
            const int e = (int)(r % Epad4);
            const int j = (int)((r / Epad4) % k);
            const int h = (int)(r / ((int64_t)Epad4 * k));
            const float v = (e < E) ? w[((int64_t)h * E + e) * k + j] : 0.f;
            hke[r] = v;
            hke16[r] = __float2bfloat16_rn(v);
        } else if (q >= n_keh + n_hke + n_umma) {
            // CTA-pair layout: [pass][half][j][chunk c][row n < Nb2/2][e%8],  filter h = pass*Nb2 + half*Nb2/2 + n
            const int64_t r = q - n_keh - n_hke - n_umma;
            const int C = Epad16 >> 3, NL = Nb2 >> 1;
            const int e8 = (int)(r & 7);
            int64_t u = r >> 3;
            const int n = (int)(u % NL); u /= NL;
            const int c = (int)(u % C); u /= C;
            const int j = (int)(u % k); u /= k;
            const int half = (int)(u & 1); u >>= 1;
            const int h = (int)u * Nb2 + half * NL + n;
            const int e = c * 8 + e8;
            const float v = (h < H && e < E) ? w[((int64_t)h * E + e) * k + j] : 0.f;
            umma2[r] = __float2bfloat16_rn(v);
        } else {
            // [pass][j][chunk c = e/8][row n][e%8],  filter h = pass*Nb + n
            const int64_t r = q - n_keh - n_hke;
            const int C = Epad16 >> 3;
            const int e8 = (int)(r & 7);
            int64_t u = r >> 3;
            const int n = (int)(u % Nb); u /= Nb;
            const int c = (int)(u % C); u /= C;
            const int j = (int)(u % k); u /= k;
            const int h = (int)u * Nb + n;
            const int e = c * 8 + e8;
            const float v = (h < H && e < E) ? w[((int64_t)h * E + e) * k + j] : 0.f;
            umma[r] = __float2bfloat16_rn(v);
        }
    }
}

}  // namespace rbr

using namespace rbr;

static int grid_for(int64_t work_items, int threads, int max_blocks = 148 * 16) {
    int64_t b = (work_items + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > max_blocks) b = max_blocks;
    return (int)b;
}

extern "C" int rbr_gather_fwd(const float* table, int64_t vocab, int64_t emb, const int64_t* ids, int64_t n_tokens,
                              float* out, void* stream) {
    RBR_REQUIRE(vocab > 0 && emb > 0 && n_tokens >= 0, RBR_EINVAL, "rbr_gather_fwd: bad sizes");
    if (n_tokens == 0) return RBR_OK;                       // empty batch: pointers may legitimately be null
    RBR_REQUIRE(table && ids && out, RBR_EINVAL, "rbr_gather_fwd: null pointer");
    cudaStream_t s = as_stream(stream);
    const bool vec = (emb % 4 == 0) && ((uintptr_t)table % 16 == 0) && ((uintptr_t)out % 16 == 0);
    if (vec) {
        const int e4 = (int)(emb / 4);
        // grid: a multiple of the SM count; 8 CTAs of 256 threads per SM keeps 64 warps resident
        const int blocks = grid_for(n_tokens * e4 / 4, 256, 148 * 8);
        gather_rows_v4_kernel<4><<<blocks, 256, 0, s>>>(reinterpret_cast<const float4*>(table), vocab, e4, ids, n_tokens,
                                                        reinterpret_cast<float4*>(out));
    } else {
        const int blocks = grid_for(n_tokens * emb, 256, 148 * 8);
        gather_rows_scalar_kernel<<<blocks, 256, 0, s>>>(table, vocab, (int)emb, ids, n_tokens, out);
    }
    RBR_LAUNCH_CHECK("gather_rows");
    return RBR_OK;
}

extern "C" int64_t rbr_embgrad_workspace_bytes(int64_t n_tokens, int64_t vocab) {
    // keys int32[n] + token sort scratch
    return round_up(n_tokens * 4, 256) + token_sort_workspace_bytes(n_tokens, vocab);
}

extern "C" int rbr_embgrad_scatter_add(const int64_t* ids, const float* grad_rows, int64_t n_tokens, int64_t emb,
                                       int64_t vocab, int64_t padding_idx, float* table_grad, void* ws, int64_t ws_bytes,
                                       void* stream) {
    RBR_REQUIRE(vocab > 0 && emb > 0 && n_tokens >= 0, RBR_EINVAL, "rbr_embgrad_scatter_add: bad sizes");
    if (n_tokens == 0) return RBR_OK;
    RBR_REQUIRE(ids && grad_rows && table_grad, RBR_EINVAL, "rbr_embgrad_scatter_add: null pointer");
    RBR_REQUIRE(n_tokens < (1ll << 31) && vocab < (1ll << 31), RBR_EUNSUPPORTED, "rbr_embgrad_scatter_add: > 2^31 entries");
    if (n_tokens == 0) return RBR_OK;
    cudaStream_t s = as_stream(stream);
    const bool vec = (emb % 4 == 0) && emb <= 512 && ((uintptr_t)grad_rows % 16 == 0) && ((uintptr_t)table_grad % 16 == 0);
    if (!vec) {
        embgrad_atomic_scalar_kernel<<<grid_for(n_tokens * emb, 256), 256, 0, s>>>(ids, grad_rows, n_tokens, (int)emb, vocab,
                                                                                   padding_idx, table_grad);
        RBR_LAUNCH_CHECK("embgrad_atomic_scalar");
        return RBR_OK;
    }
    RBR_REQUIRE(ws && ws_bytes >= rbr_embgrad_workspace_bytes(n_tokens, vocab), RBR_EWORKSPACE,
                "rbr_embgrad_scatter_add: workspace too small");
    int32_t* keys = reinterpret_cast<int32_t*>(ws);
    char* sort_ws = reinterpret_cast<char*>(ws) + round_up(n_tokens * 4, 256);
    embgrad_keys_kernel<<<grid_for(n_tokens, 256), 256, 0, s>>>(ids, n_tokens, vocab, padding_idx, keys);
    RBR_LAUNCH_CHECK("embgrad_keys");
    TokenSort ts;
    int rc = token_sort(keys, n_tokens, vocab, sort_ws, ts, s);
    if (rc != RBR_OK) return rc;
    const int e4 = (int)(emb / 4);
    const int chunk = 32;
    // upper bound on sorted entries is n_tokens (device-side count is not read back): warps past the end exit
    const int64_t warps = (n_tokens + chunk - 1) / chunk;
    const int blocks = (int)((warps * 32 + 255) / 256);
    const int nq = (e4 + 31) / 32;
#define RBR_SEG(NQ)                                                                                              \
    embgrad_segment_reduce_kernel<NQ><<<blocks, 256, 0, s>>>(ts.order, keys, n_tokens, ts.start + vocab,           \
                                                             reinterpret_cast<const float4*>(grad_rows), e4,     \
                                                             reinterpret_cast<float4*>(table_grad), chunk)
    // entries with key -1 are placed at the END of `order` by token_sort, and the kernel stops at key < 0
    if (nq == 1) RBR_SEG(1); else if (nq == 2) RBR_SEG(2); else if (nq == 3) RBR_SEG(3); else RBR_SEG(4);
#undef RBR_SEG
    RBR_LAUNCH_CHECK("embgrad_segment_reduce");
    return RBR_OK;
}

// bf16 shadow rows are padded to a multiple of 64 elements = 128 bytes, so every 64-byte piece the conv kernel
// gathers lies inside one 128-byte line
extern "C" int64_t rbr_emb_pad(int64_t emb) { return round_up(emb, 64); }

extern "C" int rbr_table_to_bf16(const float* table, int64_t vocab, int64_t emb, void* shadow_bf16, void* stream) {
    RBR_REQUIRE(table && shadow_bf16, RBR_EINVAL, "rbr_table_to_bf16: null pointer");
    RBR_REQUIRE(vocab > 0 && emb > 0, RBR_EINVAL, "rbr_table_to_bf16: bad sizes");
    RBR_REQUIRE((uintptr_t)shadow_bf16 % 16 == 0, RBR_EINVAL, "rbr_table_to_bf16: shadow must be 16-byte aligned");
    const int emb_pad = (int)rbr_emb_pad(emb);
    table_to_bf16_kernel<<<grid_for(vocab * (emb_pad / 8), 256, 148 * 8), 256, 0, as_stream(stream)>>>(
        table, vocab, (int)emb, emb_pad, reinterpret_cast<__nv_bfloat16*>(shadow_bf16));
    RBR_LAUNCH_CHECK("table_to_bf16");
    return RBR_OK;
}

extern "C" int64_t rbr_conv_pack_bytes(int64_t emb, int64_t filters, int64_t ksize) {
    return pack_layout(emb, filters, ksize).total;
}

extern "C" int rbr_conv_pack(const float* weight, int64_t emb, int64_t filters, int64_t ksize, void* packed, void* stream) {
    RBR_REQUIRE(weight && packed, RBR_EINVAL, "rbr_conv_pack: null pointer");
    RBR_REQUIRE(emb > 0 && filters > 0 && ksize > 0, RBR_EINVAL, "rbr_conv_pack: bad sizes");
    RBR_REQUIRE((uintptr_t)packed % 256 == 0, RBR_EINVAL, "rbr_conv_pack: packed buffer must be 256-byte aligned");
    const PackLayout p = pack_layout(emb, filters, ksize);
    char* base = reinterpret_cast<char*>(packed);
    RBR_CUDA(cudaMemsetAsync(base + p.off_zero, 0, (size_t)(p.off_umma2 - p.off_zero), as_stream(stream)));
    const int64_t npad2 = p.P2 * p.Nb2;
    const int64_t total = ksize * emb * p.Hpad4 + filters * ksize * p.Epad4 + ksize * p.Epad16 * (p.Npad + npad2) +
                          p.NT * p.n_tiles * 2 * p.HJp;
    conv_pack_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
        weight, (int)emb, (int)filters, (int)ksize, (int)p.Hpad4, (int)p.Epad4, (int)p.Epad16, (int)p.Npad, (int)(p.Nb > 0 ? p.Nb : 1),
        (int)npad2, (int)(p.Nb2 > 0 ? p.Nb2 : 2), reinterpret_cast<float*>(base + p.off_keh), reinterpret_cast<float*>(base + p.off_hke),
        reinterpret_cast<__nv_bfloat16*>(base + p.off_umma), reinterpret_cast<__nv_bfloat16*>(base + p.off_umma2),
        reinterpret_cast<__nv_bfloat16*>(base + p.off_hke16), reinterpret_cast<__nv_bfloat16*>(base + p.off_wt2), (int)p.HJp,
        (int)(p.NT * p.n_tiles));
    RBR_LAUNCH_CHECK("conv_pack");
    return RBR_OK;
}

RBR_DEFINE_OOB_ACCESSOR(embed)
