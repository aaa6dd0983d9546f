// api.cu — C-ABI plumbing: version, thread-local error text, out-of-range counter, conv forward dispatch.
#include <stdlib.h>

#include "rbr_common.cuh"

#include <string.h>
#include <atomic>

namespace rbr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return RBR_ECUDA;
}

int conv_fp32_dispatch(const float* table, int64_t vocab, int E, IdView ids, const uint8_t* mask, const float* gate,
                       int gate_mode, int64_t n_docs, int L, const float* keh, int Hpad4, const float* bias, int H, int K,
                       int pad, int act, float* feat, int32_t* argmax, float* preact, int feat_ld, cudaStream_t s);
int conv_tc_dispatch(const __nv_bfloat16* shadow, int64_t vocab, int E, IdView ids, const uint8_t* mask,
                     const float* gate, int gate_mode, int64_t n_docs, int L, const __nv_bfloat16* umma_w, const void* zero_row,
                     const float* bias, int H, int K, int pad, int act, float* feat, int32_t* argmax, float* preact, int feat_ld,
                     cudaStream_t s);

int conv_tc2_dispatch(const __nv_bfloat16* shadow, int64_t vocab, int E, IdView ids, const uint8_t* mask,
                      const float* gate, int gate_mode, int64_t n_docs, int L, const __nv_bfloat16* umma_w2, const float* bias,
                      int H, int K, int pad, int act, float* feat, int32_t* argmax, float* preact, int feat_ld, void* ws, int64_t ws_bytes,
                      cudaStream_t s);
int64_t conv_tc2_select_bytes(int64_t n_docs);
int64_t conv_tc2_workspace_bytes(int64_t n_docs, int64_t L, int64_t K, int64_t pad);

int oob_consume_embed(cudaStream_t, unsigned int*);
int oob_consume_conv_fp32(cudaStream_t, unsigned int*);
int oob_consume_conv_tc(cudaStream_t, unsigned int*);
int oob_consume_conv_tc2(cudaStream_t, unsigned int*);
int oob_consume_head(cudaStream_t, unsigned int*);
int oob_consume_attn(cudaStream_t, unsigned int*);
int oob_consume_attn_tc(cudaStream_t, unsigned int*);
int oob_consume_hierpool(cudaStream_t, unsigned int*);
int oob_consume_avgpool(cudaStream_t, unsigned int*);
int oob_consume_datt(cudaStream_t, unsigned int*);

}  // namespace rbr

using namespace rbr;

extern "C" int64_t rbr_conv_fwd_workspace_bytes(int64_t n_docs) { return conv_tc2_select_bytes(n_docs); }
extern "C" int64_t rbr_conv_fwd_workspace_bytes2(int64_t n_docs, int64_t doc_len, int64_t ksize, int64_t pad) {
    return conv_tc2_workspace_bytes(n_docs, doc_len, ksize, pad);
}

extern "C" int rbr_version(void) { return 100; }   // 0.1.0

extern "C" const char* rbr_last_error(void) { return g_err; }

extern "C" int64_t rbr_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int rbr_consume_oob_count(void* stream) {
    cudaStream_t s = as_stream(stream);
    unsigned long long total = 0;
    int (*const fns[])(cudaStream_t, unsigned int*) = {oob_consume_embed, oob_consume_conv_fp32, oob_consume_conv_tc, oob_consume_conv_tc2,
                                                       oob_consume_head, oob_consume_attn, oob_consume_attn_tc, oob_consume_hierpool, oob_consume_avgpool, oob_consume_datt};
    for (auto fn : fns) {
        unsigned int h = 0;
        if (fn(s, &h) != RBR_OK) return cuda_fail(cudaGetLastError(), "rbr_consume_oob_count");
        total += h;
    }
    return (int)(total > 0x7fffffffull ? 0x7fffffffull : total);
}

extern "C" int rbr_conv_act_maxpool_fwd(int precision, int activation, const void* table, const void* shadow_bf16,
                                        int64_t vocab, int64_t emb, const void* ids_raw, const uint8_t* mask, const float* gate,
                                        int gate_mode, int64_t n_docs, int64_t doc_len, const void* packed, const float* bias,
                                        int64_t filters, int64_t ksize, int64_t pad, float* feat, int32_t* argmax,
                                        float* preact, int64_t feat_ld, void* ws, int64_t ws_bytes, int flags, void* stream) {
    const IdView ids = id_view(ids_raw, flags);
    RBR_REQUIRE(ids_raw && packed && bias && feat && argmax, RBR_EINVAL, "conv_fwd: null pointer");
    RBR_REQUIRE(n_docs >= 0 && doc_len > 0 && filters > 0 && ksize > 0 && emb > 0 && vocab > 0 && pad >= 0, RBR_EINVAL,
                "conv_fwd: bad sizes");
    RBR_REQUIRE(feat_ld >= filters, RBR_EINVAL, "conv_fwd: feat_ld < filters");
    RBR_REQUIRE(activation == RBR_ACT_RELU || activation == RBR_ACT_TANH, RBR_EINVAL, "conv_fwd: unknown activation");
    RBR_REQUIRE(gate_mode >= 0 && gate_mode <= 2 && (gate_mode == 0) == (gate == nullptr), RBR_EINVAL,
                "conv_fwd: gate / gate_mode mismatch");
    RBR_REQUIRE(gate_mode != 1 || ksize == 1, RBR_EUNSUPPORTED, "conv_fwd: per-token gate needs ksize == 1");
    if (n_docs == 0) return RBR_OK;
    const PackLayout pl = pack_layout(emb, filters, ksize);
    const char* pk = reinterpret_cast<const char*>(packed);
    cudaStream_t s = as_stream(stream);
    if (precision == RBR_PREC_FP32) {
        RBR_REQUIRE(table, RBR_EINVAL, "conv_fwd: fp32 precision needs the fp32 table");
        return conv_fp32_dispatch(reinterpret_cast<const float*>(table), vocab, (int)emb, ids, mask, gate, gate_mode, n_docs,
                                  (int)doc_len, reinterpret_cast<const float*>(pk + pl.off_keh), (int)pl.Hpad4, bias,
                                  (int)filters, (int)ksize, (int)pad, activation, feat, argmax, preact, (int)feat_ld, s);
    }
    if (precision == RBR_PREC_BF16) {
        RBR_REQUIRE(shadow_bf16, RBR_EINVAL, "conv_fwd: bf16 precision needs the bf16 shadow table");
        // CTA-pair / TMA-gather kernel first, then the single-CTA cp.async kernel, then (shapes outside both) the fp32 kernel
        const bool single_only = (flags & RBR_CONV_TC_SINGLE_CTA) != 0, pair_only = (flags & RBR_CONV_TC_PAIR_ONLY) != 0;
        RBR_REQUIRE(!(single_only && pair_only), RBR_EINVAL, "conv_fwd: RBR_CONV_TC_SINGLE_CTA and RBR_CONV_TC_PAIR_ONLY exclude each other");
        if (!single_only && pl.P2 > 0) {
            const int rc2 = conv_tc2_dispatch(reinterpret_cast<const __nv_bfloat16*>(shadow_bf16), vocab, (int)emb, ids, mask, gate,
                                              gate_mode, n_docs, (int)doc_len,
                                              reinterpret_cast<const __nv_bfloat16*>(pk + pl.off_umma2), bias, (int)filters,
                                              (int)ksize, (int)pad, activation, feat, argmax, preact, (int)feat_ld, ws, ws_bytes, s);
            if (rc2 != RBR_EUNSUPPORTED) return rc2;
        }
        RBR_REQUIRE(!pair_only, RBR_EUNSUPPORTED, "conv_fwd[bf16]: shape outside the CTA-pair variant");
        if (pl.P > 0) {
            const int rc1 = conv_tc_dispatch(reinterpret_cast<const __nv_bfloat16*>(shadow_bf16), vocab, (int)emb, ids, mask, gate,
                                             gate_mode, n_docs, (int)doc_len, reinterpret_cast<const __nv_bfloat16*>(pk + pl.off_umma),
                                             pk + pl.off_zero, bias, (int)filters, (int)ksize, (int)pad, activation, feat, argmax,
                                             preact, (int)feat_ld, s);
            if (rc1 != RBR_EUNSUPPORTED) return rc1;
        }
        RBR_REQUIRE(!single_only && table, RBR_EUNSUPPORTED, "conv_fwd[bf16]: shape outside the tensor-core kernels (%s)", rbr_last_error());
        static bool warned = false;
        if (!warned) {
            warned = true;
            fprintf(stderr, "rbr_b200: conv shape (emb %lld, filters %lld, k %lld, doc_len %lld) is outside the tensor-core kernels; "
                            "using the fp32 kernel for it\n", (long long)emb, (long long)filters, (long long)ksize, (long long)doc_len);
        }
        return conv_fp32_dispatch(reinterpret_cast<const float*>(table), vocab, (int)emb, ids, mask, gate, gate_mode, n_docs,
                                  (int)doc_len, reinterpret_cast<const float*>(pk + pl.off_keh), (int)pl.Hpad4, bias,
                                  (int)filters, (int)ksize, (int)pad, activation, feat, argmax, preact, (int)feat_ld, s);
    }
    set_error("conv_fwd: unknown precision %d", precision);
    return RBR_EINVAL;
}
