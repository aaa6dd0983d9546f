// head.cu — K4: fused LastFeat (user + item) + FM interaction + MSE loss, forward and backward.
//
// Replaces, per mini-batch, LastFeat.forward x2 (reference models/deepconn/layers.py:156-165: text_feat @ W + b +
// ebd(id)), FM.forward (layers.py:188-209: relu(u*i) → dropout → @h + user_bias(id) + item_bias(id) + g_bias) and
// nn.MSELoss (trainer/train_deepconn_pp.py:140,164) — about ten cuBLAS/elementwise launches in the reference.
// Latency/HBM-bound: ~1 KB per sample.  One warp per sample; lane = latent dimension.
#include "rbr_common.cuh"

namespace rbr {

constexpr int HD_KQ = 4;          // latent dim <= 128
constexpr int HD_WARPS = 8;
constexpr int HD_BS = 32;         // samples per CTA iteration in the backward

__device__ __forceinline__ float keep_scale(float p, uint64_t seed, int64_t b, int K, int kk) {
    if (p <= 0.f) return 1.f;
    return hash_uniform(seed, (uint64_t)b * (uint64_t)K + (uint64_t)kk) >= p ? 1.f / (1.f - p) : 0.f;
}

// smem weight copy with row stride K+1 (conflict-free both for lane=k and for lane=h access)
__device__ __forceinline__ void stage_weight(float* dst, const float* __restrict__ src, int H, int K) {
    for (int i = threadIdx.x; i < H * K; i += blockDim.x) dst[(i / K) * (K + 1) + (i % K)] = src[i];
}

__global__ void __launch_bounds__(HD_WARPS * 32) head_fwd_kernel(
    const float* __restrict__ u_text, const float* __restrict__ i_text, const int64_t* __restrict__ u_id,
    const int64_t* __restrict__ i_id, int64_t B, int H, int K, const float* __restrict__ Wu, const float* __restrict__ bu,
    const float* __restrict__ ebd_u, const float* __restrict__ Wi, const float* __restrict__ bi, const float* __restrict__ ebd_i,
    const float* __restrict__ fm_h, const float* __restrict__ user_bias, const float* __restrict__ item_bias,
    const float* __restrict__ g_bias, int64_t users, int64_t items, float drop_p, uint64_t drop_seed,
    const uint64_t* __restrict__ drop_seed_dev, float* __restrict__ pred, float* __restrict__ u_lat, float* __restrict__ i_lat,
    const float* __restrict__ ratings, float grad_scale, float* __restrict__ loss_sum, float* __restrict__ pred_grad) {
    if (drop_seed_dev) drop_seed += *drop_seed_dev;          // device-resident step counter (CUDA-graph replay varies the mask)
    extern __shared__ __align__(16) float smem[];
    float* Wu_s = smem;
    float* Wi_s = smem + H * (K + 1);
    float* xrow = Wi_s + H * (K + 1);                 // [HD_WARPS][2][H] text rows of the warp's current sample
    stage_weight(Wu_s, Wu, H, K);
    stage_weight(Wi_s, Wi, H, K);
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* xu = xrow + (wib * 2) * H;
    float* xi = xu + H;
    const float gb = g_bias[0];
    float loss_acc = 0.f;
    for (int64_t b = (int64_t)blockIdx.x * HD_WARPS + wib; b < B; b += (int64_t)gridDim.x * HD_WARPS) {
        for (int h = lane; h < H; h += 32) { xu[h] = u_text[b * H + h]; xi[h] = i_text[b * H + h]; }
        __syncwarp();
        int64_t uid = u_id[b], iid = i_id[b];
        const bool u_ok = uid >= 0 && uid < users, i_ok = iid >= 0 && iid < items;
        if (lane == 0 && (!u_ok || !i_ok)) note_oob();
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < HD_KQ; ++q) {
            const int kk = lane + 32 * q;
            if (kk < K) {
                float su = 0.f, si = 0.f;
                for (int h = 0; h < H; ++h) {
                    su = fmaf(xu[h], Wu_s[h * (K + 1) + kk], su);
                    si = fmaf(xi[h], Wi_s[h * (K + 1) + kk], si);
                }
                // reference order: (text @ W + b) + ebd  (layers.py:163)
                const float ul = (su + bu[kk]) + (u_ok ? ebd_u[uid * K + kk] : 0.f);
                const float il = (si + bi[kk]) + (i_ok ? ebd_i[iid * K + kk] : 0.f);
                u_lat[b * K + kk] = ul;
                i_lat[b * K + kk] = il;
                const float fm = fmaxf(ul * il, 0.f) * keep_scale(drop_p, drop_seed, b, K, kk);
                part = fmaf(fm, fm_h[kk], part);
            }
        }
        part = warp_sum(part);
        if (lane == 0) {
            const float p = part + (u_ok ? user_bias[uid] : 0.f) + (i_ok ? item_bias[iid] : 0.f) + gb;
            pred[b] = p;
            if (ratings) {
                const float d = p - ratings[b];
                loss_acc = fmaf(d, d, loss_acc);
                if (pred_grad) pred_grad[b] = 2.f * d * grad_scale;
            }
        }
        __syncwarp();
    }
    if (ratings && loss_sum && lane == 0 && loss_acc != 0.f) atomicAdd(loss_sum, loss_acc * grad_scale);
}

__global__ void __launch_bounds__(HD_WARPS * 32) head_bwd_kernel(
    const float* __restrict__ u_text, const float* __restrict__ i_text, const int64_t* __restrict__ u_id,
    const int64_t* __restrict__ i_id, int64_t B, int H, int K, const float* __restrict__ Wu, const float* __restrict__ Wi,
    const float* __restrict__ fm_h, const float* __restrict__ u_lat, const float* __restrict__ i_lat, float drop_p,
    uint64_t drop_seed, const uint64_t* __restrict__ drop_seed_dev, int64_t pad_eu, int64_t pad_ei, int64_t pad_bu, int64_t pad_bi,
    int64_t users, int64_t items, const float* __restrict__ pred_grad,
    float* __restrict__ u_text_grad, float* __restrict__ i_text_grad, float* __restrict__ Wu_grad, float* __restrict__ bu_grad,
    float* __restrict__ ebd_u_grad, float* __restrict__ Wi_grad, float* __restrict__ bi_grad, float* __restrict__ ebd_i_grad,
    float* __restrict__ fm_h_grad, float* __restrict__ user_bias_grad, float* __restrict__ item_bias_grad,
    float* __restrict__ g_bias_grad) {
    if (drop_seed_dev) drop_seed += *drop_seed_dev;
    extern __shared__ __align__(16) float smem[];
    const int KS = K + 1;
    float* Wu_s = smem;                          // [H][K+1]
    float* Wi_s = Wu_s + H * KS;
    float* dWu_s = Wi_s + H * KS;                // [H][K]   CTA-partial weight grads (owner-exclusive updates)
    float* dWi_s = dWu_s + H * K;
    float* xu_s = dWi_s + H * K;                 // [HD_BS][H]
    float* xi_s = xu_s + HD_BS * H;
    float* du_s = xi_s + HD_BS * H;              // [HD_BS][K]
    float* di_s = du_s + HD_BS * K;
    stage_weight(Wu_s, Wu, H, K);
    stage_weight(Wi_s, Wi, H, K);
    for (int i = threadIdx.x; i < 2 * H * K; i += blockDim.x) dWu_s[i] = 0.f;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float dbu[HD_KQ], dbi[HD_KQ], dfh[HD_KQ];
#pragma unroll
    for (int q = 0; q < HD_KQ; ++q) dbu[q] = dbi[q] = dfh[q] = 0.f;
    float dgb = 0.f;
    __syncthreads();

    for (int64_t b0 = (int64_t)blockIdx.x * HD_BS; b0 < B; b0 += (int64_t)gridDim.x * HD_BS) {
        const int nb = (int)min((int64_t)HD_BS, B - b0);
        // ---- phase 1: per-sample latent gradients (one warp per sample, HD_BS/HD_WARPS samples per warp)
        for (int s = wib; s < HD_BS; s += HD_WARPS) {
            const int64_t b = b0 + s;
            if (s < nb) {
                for (int h = lane; h < H; h += 32) { xu_s[s * H + h] = u_text[b * H + h]; xi_s[s * H + h] = i_text[b * H + h]; }
                const float gp = pred_grad[b];
                const int64_t uid = u_id[b], iid = i_id[b];
                const bool u_ok = uid >= 0 && uid < users, i_ok = iid >= 0 && iid < items;
                const bool u_row = u_ok && uid != pad_eu, i_row = i_ok && iid != pad_ei;        // LastFeat.ebd padding rows
                const bool ub_row = u_ok && uid != pad_bu, ib_row = i_ok && iid != pad_bi;      // FM.user_bias / item_bias padding rows
#pragma unroll
                for (int q = 0; q < HD_KQ; ++q) {
                    const int kk = lane + 32 * q;
                    if (kk < K) {
                        const float ul = u_lat[b * K + kk], il = i_lat[b * K + kk];
                        const float prod = ul * il;
                        const float ks = keep_scale(drop_p, drop_seed, b, K, kk);
                        const float dfm = (prod > 0.f) ? gp * fm_h[kk] * ks : 0.f;
                        const float du = dfm * il, di = dfm * ul;
                        du_s[s * K + kk] = du;
                        di_s[s * K + kk] = di;
                        dbu[q] += du; dbi[q] += di;
                        dfh[q] = fmaf(fmaxf(prod, 0.f) * ks, gp, dfh[q]);
                        if (u_row && du != 0.f) atomicAdd(ebd_u_grad + uid * K + kk, du);
                        if (i_row && di != 0.f) atomicAdd(ebd_i_grad + iid * K + kk, di);
                    }
                }
                if (lane == 0) {
                    dgb += gp;
                    if (ub_row) atomicAdd(user_bias_grad + uid, gp);
                    if (ib_row) atomicAdd(item_bias_grad + iid, gp);
                }
            } else {
                for (int h = lane; h < H; h += 32) { xu_s[s * H + h] = 0.f; xi_s[s * H + h] = 0.f; }
                for (int kk = lane; kk < K; kk += 32) { du_s[s * K + kk] = 0.f; di_s[s * K + kk] = 0.f; }
            }
        }
        __syncthreads();
        // ---- phase 2a: text-feature gradients  d_text[b,h] = sum_k d_lat[b,k] * W[h,k]
        for (int o = threadIdx.x; o < nb * H; o += blockDim.x) {
            const int s = o / H, h = o - s * H;
            float au = 0.f, ai = 0.f;
            for (int kk = 0; kk < K; ++kk) {
                au = fmaf(du_s[s * K + kk], Wu_s[h * KS + kk], au);
                ai = fmaf(di_s[s * K + kk], Wi_s[h * KS + kk], ai);
            }
            u_text_grad[(b0 + s) * H + h] = au;
            i_text_grad[(b0 + s) * H + h] = ai;
        }
        // ---- phase 2b: weight gradients  dW[h,k] += sum_b text[b,h] * d_lat[b,k]   (thread owns (h,k))
        for (int o = threadIdx.x; o < H * K; o += blockDim.x) {
            const int h = o / K, kk = o - h * K;
            float au = 0.f, ai = 0.f;
#pragma unroll 8
            for (int s = 0; s < HD_BS; ++s) {
                au = fmaf(xu_s[s * H + h], du_s[s * K + kk], au);
                ai = fmaf(xi_s[s * H + h], di_s[s * K + kk], ai);
            }
            dWu_s[o] += au;
            dWi_s[o] += ai;
        }
        __syncthreads();
    }
    // ---- flush CTA partials
    for (int o = threadIdx.x; o < H * K; o += blockDim.x) {
        if (dWu_s[o] != 0.f) atomicAdd(Wu_grad + o, dWu_s[o]);
        if (dWi_s[o] != 0.f) atomicAdd(Wi_grad + o, dWi_s[o]);
    }
#pragma unroll
    for (int q = 0; q < HD_KQ; ++q) {
        const int kk = lane + 32 * q;
        if (kk < K) {
            if (dbu[q] != 0.f) atomicAdd(bu_grad + kk, dbu[q]);
            if (dbi[q] != 0.f) atomicAdd(bi_grad + kk, dbi[q]);
            if (dfh[q] != 0.f) atomicAdd(fm_h_grad + kk, dfh[q]);
        }
    }
    if (lane == 0 && dgb != 0.f) atomicAdd(g_bias_grad, dgb);
}

// test / debug aid: the keep-scale both kernels apply, keep[b,k] in {0, 1/(1-p)}
__global__ void head_dropout_mask_kernel(int64_t B, int K, float drop_p, uint64_t drop_seed, const uint64_t* __restrict__ drop_seed_dev,
                                         float* __restrict__ keep) {
    if (drop_seed_dev) drop_seed += *drop_seed_dev;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < B * K; q += stride)
        keep[q] = keep_scale(drop_p, drop_seed, q / K, K, (int)(q % K));
}

}  // namespace rbr

using namespace rbr;

extern "C" int rbr_head_dropout_mask(int64_t batch, int64_t latent, float drop_p, uint64_t drop_seed, const uint64_t* drop_seed_dev,
                                     float* keep, void* stream) {
    RBR_REQUIRE(keep && batch >= 0 && latent > 0, RBR_EINVAL, "rbr_head_dropout_mask: bad arguments");
    RBR_REQUIRE(drop_p >= 0.f && drop_p < 1.f, RBR_EINVAL, "rbr_head_dropout_mask: dropout p must be in [0,1)");
    if (batch == 0) return RBR_OK;
    int64_t blocks = (batch * latent + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    head_dropout_mask_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(batch, (int)latent, drop_p, drop_seed, drop_seed_dev, keep);
    RBR_LAUNCH_CHECK("head_dropout_mask_kernel");
    return RBR_OK;
}

extern "C" int rbr_head_fwd(const float* u_text, const float* i_text, const int64_t* u_id, const int64_t* i_id, int64_t batch,
                            int64_t hidden, int64_t latent, const float* Wu, const float* bu, const float* ebd_u,
                            const float* Wi, const float* bi, const float* ebd_i, const float* fm_h, const float* user_bias,
                            const float* item_bias, const float* g_bias, int64_t users, int64_t items, float drop_p,
                            uint64_t drop_seed, const uint64_t* drop_seed_dev, float* pred, float* u_lat, float* i_lat,
                            const float* ratings, float grad_scale, float* loss_sum, float* pred_grad, void* stream) {
    RBR_REQUIRE(u_text && i_text && u_id && i_id && Wu && bu && ebd_u && Wi && bi && ebd_i && fm_h && user_bias && item_bias &&
                    g_bias && pred && u_lat && i_lat,
                RBR_EINVAL, "rbr_head_fwd: null pointer");
    RBR_REQUIRE(batch >= 0 && hidden > 0 && latent > 0 && latent <= 32 * HD_KQ, RBR_EUNSUPPORTED,
                "rbr_head_fwd: latent_dim must be in [1,%d]", 32 * HD_KQ);
    RBR_REQUIRE(drop_p >= 0.f && drop_p < 1.f, RBR_EINVAL, "rbr_head_fwd: dropout p must be in [0,1)");
    if (batch == 0) return RBR_OK;
    const int H = (int)hidden, K = (int)latent;
    const size_t smem = ((size_t)2 * H * (K + 1) + (size_t)HD_WARPS * 2 * H) * 4;
    RBR_REQUIRE(smem <= 200 * 1024, RBR_EUNSUPPORTED, "rbr_head_fwd: hidden*latent too large for shared memory");
    RBR_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (batch + HD_WARPS - 1) / HD_WARPS;
    if (blocks > 148 * 2) blocks = 148 * 2;
    head_fwd_kernel<<<(unsigned)blocks, HD_WARPS * 32, smem, as_stream(stream)>>>(
        u_text, i_text, u_id, i_id, batch, H, K, Wu, bu, ebd_u, Wi, bi, ebd_i, fm_h, user_bias, item_bias, g_bias, users, items,
        drop_p, drop_seed, drop_seed_dev, pred, u_lat, i_lat, ratings, grad_scale, loss_sum, pred_grad);
    RBR_LAUNCH_CHECK("head_fwd_kernel");
    return RBR_OK;
}

extern "C" int rbr_head_bwd(const float* u_text, const float* i_text, const int64_t* u_id, const int64_t* i_id, int64_t batch,
                            int64_t hidden, int64_t latent, const float* Wu, const float* Wi, const float* fm_h,
                            const float* u_lat, const float* i_lat, float drop_p, uint64_t drop_seed,
                            const uint64_t* drop_seed_dev, int64_t ebd_u_padding_idx, int64_t ebd_i_padding_idx,
                            int64_t user_bias_padding_idx, int64_t item_bias_padding_idx, int64_t users, int64_t items, const float* pred_grad, float* u_text_grad, float* i_text_grad, float* Wu_grad, float* bu_grad,
                            float* ebd_u_grad, float* Wi_grad, float* bi_grad, float* ebd_i_grad, float* fm_h_grad,
                            float* user_bias_grad, float* item_bias_grad, float* g_bias_grad, void* stream) {
    RBR_REQUIRE(u_text && i_text && u_id && i_id && Wu && Wi && fm_h && u_lat && i_lat && pred_grad && u_text_grad &&
                    i_text_grad && Wu_grad && bu_grad && ebd_u_grad && Wi_grad && bi_grad && ebd_i_grad && fm_h_grad &&
                    user_bias_grad && item_bias_grad && g_bias_grad,
                RBR_EINVAL, "rbr_head_bwd: null pointer");
    RBR_REQUIRE(batch >= 0 && hidden > 0 && latent > 0 && latent <= 32 * HD_KQ, RBR_EUNSUPPORTED,
                "rbr_head_bwd: latent_dim must be in [1,%d]", 32 * HD_KQ);
    if (batch == 0) return RBR_OK;
    const int H = (int)hidden, K = (int)latent;
    const size_t smem = ((size_t)2 * H * (K + 1) + (size_t)2 * H * K + (size_t)2 * HD_BS * H + (size_t)2 * HD_BS * K) * 4;
    RBR_REQUIRE(smem <= 200 * 1024, RBR_EUNSUPPORTED, "rbr_head_bwd: hidden*latent too large for shared memory");
    RBR_CUDA(cudaFuncSetAttribute(head_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (batch + HD_BS - 1) / HD_BS;
    if (blocks > 148 * 2) blocks = 148 * 2;
    head_bwd_kernel<<<(unsigned)blocks, HD_WARPS * 32, smem, as_stream(stream)>>>(
        u_text, i_text, u_id, i_id, batch, H, K, Wu, Wi, fm_h, u_lat, i_lat, drop_p, drop_seed, drop_seed_dev, ebd_u_padding_idx,
        ebd_i_padding_idx, user_bias_padding_idx, item_bias_padding_idx, users, items, pred_grad, u_text_grad, i_text_grad, Wu_grad, bu_grad, ebd_u_grad, Wi_grad, bi_grad,
        ebd_i_grad, fm_h_grad, user_bias_grad, item_bias_grad, g_bias_grad);
    RBR_LAUNCH_CHECK("head_bwd_kernel");
    return RBR_OK;
}

RBR_DEFINE_OOB_ACCESSOR(head)
