// conv_fp32.cu — K2 (fp32 variant): fused gather → mask → Conv1d → bias → activation → max-over-time.
//
// Replaces NgramFeat.forward (reference models/deepconn/layers.py:123-136) and the embedding gather that
// feeds it (models/deepconn/deepconn.py:43-47) with ONE kernel: the [N,L,E] embedded documents, the
// masked copy, the [N,H,L] conv output and its ReLU'd copy are never written to HBM.
//
// This is the fp32 CUDA-core variant (RBR_PREC_FP32): an implicit GEMM with 8x8 register tiles that
// accumulates in fp32 in the same order for every run, used for the 1e-5 parity mode and for shapes
// outside the tensor-core kernel's envelope.  The bf16 tcgen05 variant lives in conv_tc.cu.
//
// Work decomposition: one CTA per (document, 256-filter block).  The CTA walks the document in tiles of
// TM output positions; per tile and per EK-wide slice of the embedding dim it stages
//     Xs[e][r]  = x[doc, t0 - pad + r, e0 + e]       (gathered straight from the table, transposed)
//     Ws[j][e][h] = W[h, e0 + e, j]                    (from the [k][E][Hpad] packed copy)
// and each thread accumulates an 8 (positions) x 8 (filters) tile:  acc[i][c] += Xs[e][8ty+i+j] * Ws[j][e][8tx+c].
// The running (max, first arg-max) over positions lives in registers across tiles.
#include "rbr_common.cuh"

namespace rbr {

constexpr int CF_EK = 16;   // embedding-dim slice per stage

__host__ __device__ constexpr int cf_xs_stride(int tm, int k) {
    // smallest stride >= tm + k - 1 with stride % 8 == 2: conflict-free transposed stores (see load loop)
    int need = tm + k - 1;
    int s = (need / 8) * 8 + 2;
    return s >= need ? s : s + 8;
}

template <int TM, int K>
__global__ void __launch_bounds__(32 * (TM / 8)) conv_fp32_kernel(
    const float* __restrict__ table, int64_t vocab, int E, const IdView ids,
    const uint8_t* __restrict__ mask, const float* __restrict__ gate, int gate_mode, int L,
    const float* __restrict__ keh, int Hpad4, const float* __restrict__ bias, int H, int pad, int Lout, int act,
    float* __restrict__ feat, int32_t* __restrict__ argmax, float* __restrict__ preact, int feat_ld, int TX) {
    constexpr int TY = TM / 8;
    constexpr int XS = cf_xs_stride(TM, K);
    constexpr int XROWS = TM + K - 1;
    extern __shared__ __align__(16) float smem[];
    const int HB = TX * 8;
    float* Xs = smem;                       // [CF_EK][XS]
    float* Ws = smem + CF_EK * XS;          // [K][CF_EK][HB]
    int64_t* row_src = reinterpret_cast<int64_t*>(Ws + K * CF_EK * HB);   // [XROWS] table row or -1

    const int nthreads = TX * TY;
    const int tid = threadIdx.x;
    const bool active = tid < nthreads;
    const int tx = tid % TX, ty = tid / TX;
    const int64_t doc = blockIdx.x;
    const int hb0 = blockIdx.y * 256;
    const bool vec = (E % 4 == 0);

    float best_v[8];
    int best_t[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) { best_v[c] = -INFINITY; best_t[c] = 0; }
    const float doc_gate = (gate_mode == 2) ? gate[doc] : 1.f;

    for (int t0 = 0; t0 < Lout; t0 += TM) {
        // which table row feeds each staged position (−1 → zeros: conv padding, masked, or bad id)
        for (int r = tid; r < XROWS; r += blockDim.x) {
            const int t = t0 - pad + r;
            int64_t src = -1;
            if (t >= 0 && t < L) {
                const int64_t id = ld_id(ids, doc * L + t);
                if (ld_mask(ids, mask, doc * L + t, id)) {
                    if (id >= 0 && id < vocab) src = id; else note_oob();
                }
            }
            row_src[r] = src;
        }
        float acc[8][8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[i][c] = 0.f;
        __syncthreads();

        for (int e0 = 0; e0 < E; e0 += CF_EK) {
            // ---- stage Xs (transposed gather).  piece p = (row r, quarter q): 4 consecutive e's of one row.
            // lanes of a warp: r = lane/4 + ..., q = lane%4 → store address (4q+i)*XS + r; with XS % 8 == 2
            // the 32 lanes hit 32 distinct banks.
            for (int p = tid; p < XROWS * (CF_EK / 4); p += blockDim.x) {
                const int r = p >> 2, q = p & 3;
                const int e = e0 + q * 4;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int64_t src = row_src[r];
                if (src >= 0 && e < E) {
                    const float* rowp = table + src * E + e;
                    if (vec) {
                        v = __ldg(reinterpret_cast<const float4*>(rowp));
                    } else {
                        v.x = __ldg(rowp);
                        if (e + 1 < E) v.y = __ldg(rowp + 1);
                        if (e + 2 < E) v.z = __ldg(rowp + 2);
                        if (e + 3 < E) v.w = __ldg(rowp + 3);
                    }
                }
                float* d = Xs + (q * 4) * XS + r;
                d[0] = v.x; d[XS] = v.y; d[2 * XS] = v.z; d[3 * XS] = v.w;
            }
            // ---- stage Ws: [K][EK][HB] ← keh[(j*E + e)*Hpad4 + hb0 + c]
            for (int p = tid; p < K * CF_EK * (HB / 4); p += blockDim.x) {
                const int c4 = p % (HB / 4);
                const int e = (p / (HB / 4)) % CF_EK;
                const int j = p / ((HB / 4) * CF_EK);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                const int h = hb0 + c4 * 4;
                if (e0 + e < E && h < Hpad4)
                    v = __ldg(reinterpret_cast<const float4*>(keh + ((int64_t)j * E + e0 + e) * Hpad4 + h));
                *reinterpret_cast<float4*>(Ws + (j * CF_EK + e) * HB + c4 * 4) = v;
            }
            __syncthreads();
            if (active) {
#pragma unroll 4
                for (int e = 0; e < CF_EK; ++e) {
                    float a[8 + K - 1];
                    const float* xr = Xs + e * XS + ty * 8;
#pragma unroll
                    for (int i = 0; i < 8 + K - 1; ++i) a[i] = xr[i];
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        const float4 b0 = *reinterpret_cast<const float4*>(Ws + (j * CF_EK + e) * HB + tx * 8);
                        const float4 b1 = *reinterpret_cast<const float4*>(Ws + (j * CF_EK + e) * HB + tx * 8 + 4);
                        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int i = 0; i < 8; ++i)
#pragma unroll
                            for (int c = 0; c < 8; ++c) acc[i][c] = fmaf(a[i + j], b[c], acc[i][c]);
                    }
                }
            }
            __syncthreads();
        }
        // ---- tile epilogue: bias (+ gate), running max with FIRST arg-max (positions visited in order)
        if (active) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int t = t0 + ty * 8 + i;
                if (t < Lout) {
                    float g = doc_gate;
                    if (gate_mode == 1) g = gate[doc * L + t];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const float v = acc[i][c] * g;                        // gate * conv_nobias(x); the bias is added after the max
                        // `!(v <= best)`: a NaN wins and stays (torch's max_pool1d propagates NaN), ties keep the earlier position
                        if (!(v <= best_v[c]) && !(best_v[c] != best_v[c])) { best_v[c] = v; best_t[c] = t; }
                    }
                }
            }
        }
    }
    // ---- combine the TY partial (max, arg-max) per filter: max value, ties → smallest position
    __syncthreads();
    float* red_v = smem;                                   // [TY][HB]
    int* red_t = reinterpret_cast<int*>(smem + TY * HB);   // [TY][HB]   (fits: TY*HB*2 <= staged tile sizes)
    if (active) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            red_v[ty * HB + tx * 8 + c] = best_v[c];
            red_t[ty * HB + tx * 8 + c] = best_t[c];
        }
    }
    __syncthreads();
    for (int c = tid; c < HB; c += blockDim.x) {
        const int h = hb0 + c;
        if (h >= H) continue;
        float bv = red_v[c];
        int bt = red_t[c];
        for (int y = 1; y < TY; ++y) {
            const float v = red_v[y * HB + c];
            const int t = red_t[y * HB + c];
            const bool v_nan = v != v, b_nan = bv != bv;
            if ((v_nan && (!b_nan || t < bt)) || (!v_nan && !b_nan && (v > bv || (v == bv && t < bt)))) { bv = v; bt = t; }
        }
        feat[doc * feat_ld + h] = act_apply(act, bv + bias[h]);
        if (preact) preact[doc * feat_ld + h] = bv;                      // pool_raw: no bias
        argmax[doc * feat_ld + h] = bt;
    }
}

template <int TM, int K>
static int launch_conv_fp32(const float* table, int64_t vocab, int E, IdView ids, const uint8_t* mask,
                            const float* gate, int gate_mode, int64_t n_docs, int L, const float* keh, int Hpad4,
                            const float* bias, int H, int pad, int Lout, int act, float* feat, int32_t* argmax, float* preact,
                            int feat_ld, cudaStream_t s) {
    constexpr int TY = TM / 8;
    constexpr int XS = cf_xs_stride(TM, K);
    const int hblocks = (H + 255) / 256;
    const int hb = H < 256 ? H : 256;
    const int TX = (hb + 7) / 8;
    const int HB = TX * 8;
    size_t smem = (size_t)(CF_EK * XS + K * CF_EK * HB) * 4 + (size_t)(TM + K - 1) * 8;
    const size_t red = (size_t)TY * HB * 8;
    if (red > smem) smem = red;
    auto kern = conv_fp32_kernel<TM, K>;
    static bool attr_set = false;   // per instantiation
    if (!attr_set) {
        RBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
    }
    RBR_REQUIRE(smem <= 200 * 1024, RBR_EUNSUPPORTED, "conv_fp32: shared memory %zu too large", smem);
    dim3 grid((unsigned)n_docs, (unsigned)hblocks);
    kern<<<grid, 32 * TY, smem, s>>>(table, vocab, E, ids, mask, gate, gate_mode, L, keh, Hpad4, bias, H, pad, Lout, act,
                                     feat, argmax, preact, feat_ld, TX);
    RBR_LAUNCH_CHECK("conv_fp32_kernel");
    return RBR_OK;
}

int conv_fp32_dispatch(const float* table, int64_t vocab, int E, IdView ids, const uint8_t* mask,
                       const float* gate, int gate_mode, int64_t n_docs, int L, const float* keh, int Hpad4,
                       const float* bias, int H, int K, int pad, int act, float* feat, int32_t* argmax, float* preact, int feat_ld,
                       cudaStream_t s) {
    const int Lout = L + 2 * pad - K + 1;
    RBR_REQUIRE(Lout >= 1, RBR_EINVAL, "conv: doc_len %d too short for kernel size %d", L, K);
    RBR_REQUIRE(n_docs <= 0x7fffffff, RBR_EUNSUPPORTED, "conv: too many documents");
    // TX*TY threads must fit one CTA: TX <= 32, TY = TM/8
#define RBR_CF(TM_, K_)                                                                                          \
    return launch_conv_fp32<TM_, K_>(table, vocab, E, ids, mask, gate, gate_mode, n_docs, L, keh, Hpad4, bias, H, \
                                     pad, Lout, act, feat, argmax, preact, feat_ld, s)
#define RBR_CF_K(TM_)                                     \
    switch (K) {                                          \
        case 1: RBR_CF(TM_, 1);                           \
        case 2: RBR_CF(TM_, 2);                           \
        case 3: RBR_CF(TM_, 3);                           \
        case 4: RBR_CF(TM_, 4);                           \
        case 5: RBR_CF(TM_, 5);                           \
        case 7: RBR_CF(TM_, 7);                           \
        default: break;                                   \
    }
    if (Lout > 96) { RBR_CF_K(128) } else { RBR_CF_K(64) }
#undef RBR_CF_K
#undef RBR_CF
    set_error("conv: kernel size %d not supported (supported: 1,2,3,4,5,7)", K);
    return RBR_EUNSUPPORTED;
}

}  // namespace rbr

RBR_DEFINE_OOB_ACCESSOR(conv_fp32)
