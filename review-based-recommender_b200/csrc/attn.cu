// attn.cu — K3: fused NARRE review-level attention (score → softmax over reviews → weighted pool), fwd + bwd.
//
// Replaces LinearAttention.forward (reference models/narre/narre.py:40-64):
//     e      = ebd_vals(other_id)                                    [B,R,A]     (narre.py:53)
//     logit  = relu(feat@W_rv + e@W_id + b_1) @ h + b_2              [B,R]       (narre.py:55)
//     score  = exp(logit) / (sum_R exp(logit) + 1e-8)                            (narre.py:58; no mask, no max-shift)
//     out    = sum_R score * feat                                    [B,H]       (narre.py:60)
// — two batched GEMMs plus ~8 elementwise launches in the reference, ~12-16 KB per sample.  Here: one warp per
// sample, lane = attention dim, the sample's R x H feature block staged once in shared memory, softmax with
// warp shuffles.  exp is evaluated as exp(logit - m) with the epsilon rescaled by exp(-m): identical in exact
// arithmetic, but finite where the reference overflows (logits > 88; SURVEY.md §7 "NARRE softmax quirks").
#include "rbr_common.cuh"

namespace rbr {

constexpr int AT_AQ = 4;     // att_dim <= 128

struct AttnSmem {
    int wrv, wid, dwrv, dwid, per_warp, fs, es, hs, dhs, ls, total_floats;
};
__host__ __device__ inline AttnSmem attn_smem(int R, int H, int A, int warps, bool bwd) {
    AttnSmem s;
    int off = 0;
    s.wrv = off; off += H * (A + 1);
    s.wid = off; off += A * (A + 1);
    s.dwrv = off; if (bwd) off += H * A;
    s.dwid = off; if (bwd) off += A * A;
    int pw = 0;
    s.fs = pw; pw += R * H;
    s.es = pw; pw += R * A;
    s.hs = pw; pw += R * A;
    s.dhs = pw; if (bwd) pw += R * A;
    s.ls = pw; pw += 2 * R;
    s.per_warp = pw;
    s.total_floats = off + warps * pw;
    return s;
}

// hid[r][a] = relu(b1[a] + feat[r]·W_rv[:,a] + e[r]·W_id[:,a]); logit[r] = hid[r]·h + b2 → ls[r]
__device__ __forceinline__ void attn_logits(const float* fs, const float* es, float* hs, float* ls, const float* Wrv_s,
                                            const float* Wid_s, const float* __restrict__ hvec, const float* __restrict__ b1,
                                            float b2, int R, int H, int A, int lane) {
    for (int r = 0; r < R; ++r) {
        float part = 0.f;
#pragma unroll
        for (int q = 0; q < AT_AQ; ++q) {
            const int a = lane + 32 * q;
            if (a < A) {
                float s1 = 0.f, s2 = 0.f;
                for (int h = 0; h < H; ++h) s1 = fmaf(fs[r * H + h], Wrv_s[h * (A + 1) + a], s1);
                for (int a2 = 0; a2 < A; ++a2) s2 = fmaf(es[r * A + a2], Wid_s[a2 * (A + 1) + a], s2);
                const float hid = fmaxf((s1 + s2) + b1[a], 0.f);
                hs[r * A + a] = hid;
                part = fmaf(hid, hvec[a], part);
            }
        }
        part = warp_sum(part);
        if (lane == 0) ls[r] = part + b2;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(256) narre_attn_fwd_kernel(const float* __restrict__ feat, const int64_t* __restrict__ other_id,
                                                             int64_t B, int R, int H, int A, const float* __restrict__ W_rv,
                                                             const float* __restrict__ W_id, const float* __restrict__ hvec,
                                                             const float* __restrict__ b1, const float* __restrict__ b2,
                                                             const float* __restrict__ ebd, int64_t n_ids, float* __restrict__ out,
                                                             float* __restrict__ scores) {
    extern __shared__ __align__(16) float smem[];
    const int warps = blockDim.x >> 5;
    const AttnSmem L = attn_smem(R, H, A, warps, false);
    float* Wrv_s = smem + L.wrv;
    float* Wid_s = smem + L.wid;
    for (int i = threadIdx.x; i < H * A; i += blockDim.x) Wrv_s[(i / A) * (A + 1) + (i % A)] = W_rv[i];
    for (int i = threadIdx.x; i < A * A; i += blockDim.x) Wid_s[(i / A) * (A + 1) + (i % A)] = W_id[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* pw = smem + (L.total_floats - warps * L.per_warp) + wib * L.per_warp;
    float* fs = pw + L.fs; float* es = pw + L.es; float* hs = pw + L.hs; float* ls = pw + L.ls;
    const float b2v = b2[0];
    for (int64_t b = (int64_t)blockIdx.x * warps + wib; b < B; b += (int64_t)gridDim.x * warps) {
        for (int i = lane; i < R * H; i += 32) fs[i] = feat[b * R * H + i];
        for (int i = lane; i < R * A; i += 32) {
            const int r = i / A, a = i - r * A;
            const int64_t id = other_id[b * R + r];
            const bool ok = id >= 0 && id < n_ids;
            if (!ok && a == 0) note_oob();
            es[i] = ok ? ebd[id * A + a] : 0.f;
        }
        __syncwarp();
        attn_logits(fs, es, hs, ls, Wrv_s, Wid_s, hvec, b1, b2v, R, H, A, lane);
        float m = -INFINITY;
        for (int r = lane; r < R; r += 32) m = fmaxf(m, ls[r]);
        m = warp_max(m);
        float ssum = 0.f;
        for (int r = lane; r < R; r += 32) ssum += expf(ls[r] - m);
        ssum = warp_sum(ssum);
        const float denom = ssum + 1e-8f * expf(-m);
        for (int r = lane; r < R; r += 32) {
            const float sc = expf(ls[r] - m) / denom;
            ls[R + r] = sc;
            scores[b * R + r] = sc;
        }
        __syncwarp();
        for (int h = lane; h < H; h += 32) {
            float acc = 0.f;
            for (int r = 0; r < R; ++r) acc = fmaf(ls[R + r], fs[r * H + h], acc);
            out[b * H + h] = acc;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) narre_attn_bwd_kernel(
    const float* __restrict__ feat, const int64_t* __restrict__ other_id, int64_t B, int R, int H, int A,
    const float* __restrict__ W_rv, const float* __restrict__ W_id, const float* __restrict__ hvec,
    const float* __restrict__ b1, const float* __restrict__ b2, const float* __restrict__ ebd, int64_t n_ids,
    int64_t padding_idx, const float* __restrict__ scores, const float* __restrict__ out_grad,
    const float* __restrict__ scores_grad, float* __restrict__ feat_grad, float* __restrict__ W_rv_grad,
    float* __restrict__ W_id_grad, float* __restrict__ h_grad, float* __restrict__ b1_grad, float* __restrict__ b2_grad,
    float* __restrict__ ebd_grad) {
    extern __shared__ __align__(16) float smem[];
    const int warps = blockDim.x >> 5;
    const AttnSmem L = attn_smem(R, H, A, warps, true);
    float* Wrv_s = smem + L.wrv;
    float* Wid_s = smem + L.wid;
    float* dWrv_s = smem + L.dwrv;
    float* dWid_s = smem + L.dwid;
    for (int i = threadIdx.x; i < H * A; i += blockDim.x) { Wrv_s[(i / A) * (A + 1) + (i % A)] = W_rv[i]; dWrv_s[i] = 0.f; }
    for (int i = threadIdx.x; i < A * A; i += blockDim.x) { Wid_s[(i / A) * (A + 1) + (i % A)] = W_id[i]; dWid_s[i] = 0.f; }
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* pw = smem + (L.total_floats - warps * L.per_warp) + wib * L.per_warp;
    float* fs = pw + L.fs; float* es = pw + L.es; float* hs = pw + L.hs; float* dhs = pw + L.dhs; float* ls = pw + L.ls;
    const float b2v = b2[0];
    float dh_acc[AT_AQ], db1_acc[AT_AQ];
#pragma unroll
    for (int q = 0; q < AT_AQ; ++q) dh_acc[q] = db1_acc[q] = 0.f;
    float db2_acc = 0.f;

    for (int64_t b = (int64_t)blockIdx.x * warps + wib; b < B; b += (int64_t)gridDim.x * warps) {
        for (int i = lane; i < R * H; i += 32) fs[i] = feat[b * R * H + i];
        for (int i = lane; i < R * A; i += 32) {
            const int r = i / A, a = i - r * A;
            const int64_t id = other_id[b * R + r];
            es[i] = (id >= 0 && id < n_ids) ? ebd[id * A + a] : 0.f;
        }
        __syncwarp();
        attn_logits(fs, es, hs, ls, Wrv_s, Wid_s, hvec, b1, b2v, R, H, A, lane);     // recompute hid (ls[0..R) unused below)
        // ds[r] = out_grad · feat[r] (+ scores_grad[r])
        for (int r = 0; r < R; ++r) {
            float p = 0.f;
            for (int h = lane; h < H; h += 32) p = fmaf(out_grad[b * H + h], fs[r * H + h], p);
            p = warp_sum(p);
            if (lane == 0) ls[r] = p + (scores_grad ? scores_grad[b * R + r] : 0.f);
        }
        __syncwarp();
        float dot = 0.f;
        for (int r = lane; r < R; r += 32) dot = fmaf(ls[r], scores[b * R + r], dot);
        dot = warp_sum(dot);
        for (int r = lane; r < R; r += 32) {
            const float sc = scores[b * R + r];
            const float dl = sc * (ls[r] - dot);                  // d loss / d logit[r]
            ls[R + r] = dl;
            db2_acc += dl;
        }
        __syncwarp();
        // d hid, d h, d b1
        for (int r = 0; r < R; ++r) {
            const float dl = ls[R + r];
#pragma unroll
            for (int q = 0; q < AT_AQ; ++q) {
                const int a = lane + 32 * q;
                if (a < A) {
                    const float hid = hs[r * A + a];
                    dh_acc[q] = fmaf(dl, hid, dh_acc[q]);
                    const float dhid = hid > 0.f ? dl * hvec[a] : 0.f;
                    dhs[r * A + a] = dhid;
                    db1_acc[q] += dhid;
                }
            }
        }
        __syncwarp();
        // d feat[r][h] = score[r] * out_grad[h] + sum_a dhid[r][a] * W_rv[h][a]
        for (int r = 0; r < R; ++r) {
            const float sc = scores[b * R + r];
            for (int h = lane; h < H; h += 32) {
                float acc = sc * out_grad[b * H + h];
                for (int a = 0; a < A; ++a) acc = fmaf(dhs[r * A + a], Wrv_s[h * (A + 1) + a], acc);
                feat_grad[(b * R + r) * H + h] = acc;
            }
        }
        // d e[r][a2] = sum_a dhid[r][a] * W_id[a2][a] → id-embedding rows (padding row skipped)
        for (int i = lane; i < R * A; i += 32) {
            const int r = i / A, a2 = i - r * A;
            const int64_t id = other_id[b * R + r];
            if (id < 0 || id >= n_ids || id == padding_idx) continue;
            float acc = 0.f;
            for (int a = 0; a < A; ++a) acc = fmaf(dhs[r * A + a], Wid_s[a2 * (A + 1) + a], acc);
            if (acc != 0.f) atomicAdd(ebd_grad + id * A + a2, acc);
        }
        // d W_rv[h][a] += sum_r feat[r][h] * dhid[r][a];  d W_id[a2][a] += sum_r e[r][a2] * dhid[r][a]
#pragma unroll
        for (int q = 0; q < AT_AQ; ++q) {
            const int a = lane + 32 * q;
            if (a < A) {
                for (int h = 0; h < H; ++h) {
                    float acc = 0.f;
                    for (int r = 0; r < R; ++r) acc = fmaf(fs[r * H + h], dhs[r * A + a], acc);
                    if (acc != 0.f) atomicAdd(dWrv_s + h * A + a, acc);
                }
                for (int a2 = 0; a2 < A; ++a2) {
                    float acc = 0.f;
                    for (int r = 0; r < R; ++r) acc = fmaf(es[r * A + a2], dhs[r * A + a], acc);
                    if (acc != 0.f) atomicAdd(dWid_s + a2 * A + a, acc);
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < H * A; i += blockDim.x) if (dWrv_s[i] != 0.f) atomicAdd(W_rv_grad + i, dWrv_s[i]);
    for (int i = threadIdx.x; i < A * A; i += blockDim.x) if (dWid_s[i] != 0.f) atomicAdd(W_id_grad + i, dWid_s[i]);
#pragma unroll
    for (int q = 0; q < AT_AQ; ++q) {
        const int a = lane + 32 * q;
        if (a < A) {
            if (dh_acc[q] != 0.f) atomicAdd(h_grad + a, dh_acc[q]);
            if (db1_acc[q] != 0.f) atomicAdd(b1_grad + a, db1_acc[q]);
        }
    }
    db2_acc = warp_sum(db2_acc);
    if (lane == 0 && db2_acc != 0.f) atomicAdd(b2_grad, db2_acc);
}

static int pick_warps(int R, int H, int A, bool bwd) {
    for (int w = 8; w >= 1; w >>= 1)
        if ((size_t)attn_smem(R, H, A, w, bwd).total_floats * 4 <= 200 * 1024) return w;
    return 0;
}

}  // namespace rbr

using namespace rbr;

extern "C" int rbr_narre_attn_fwd(const float* feat, const int64_t* other_id, int64_t batch, int64_t reviews, int64_t hidden,
                                  int64_t att, const float* W_rv, const float* W_id, const float* h, const float* b_1,
                                  const float* b_2, const float* ebd_vals, int64_t n_ids, float* out, float* scores,
                                  void* stream) {
    RBR_REQUIRE(feat && other_id && W_rv && W_id && h && b_1 && b_2 && ebd_vals && out && scores, RBR_EINVAL,
                "rbr_narre_attn_fwd: null pointer");
    RBR_REQUIRE(batch >= 0 && reviews > 0 && hidden > 0 && att > 0 && att <= 32 * AT_AQ, RBR_EUNSUPPORTED,
                "rbr_narre_attn_fwd: att_dim must be in [1,%d]", 32 * AT_AQ);
    if (batch == 0) return RBR_OK;
    const int R = (int)reviews, H = (int)hidden, A = (int)att;
    const int warps = pick_warps(R, H, A, false);
    RBR_REQUIRE(warps > 0, RBR_EUNSUPPORTED, "rbr_narre_attn_fwd: reviews*hidden too large for shared memory");
    const size_t smem = (size_t)attn_smem(R, H, A, warps, false).total_floats * 4;
    RBR_CUDA(cudaFuncSetAttribute(narre_attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (batch + warps - 1) / warps;
    if (blocks > 148 * 2) blocks = 148 * 2;
    narre_attn_fwd_kernel<<<(unsigned)blocks, warps * 32, smem, as_stream(stream)>>>(feat, other_id, batch, R, H, A, W_rv, W_id, h,
                                                                                     b_1, b_2, ebd_vals, n_ids, out, scores);
    RBR_LAUNCH_CHECK("narre_attn_fwd_kernel");
    return RBR_OK;
}

extern "C" int rbr_narre_attn_bwd(const float* feat, const int64_t* other_id, int64_t batch, int64_t reviews, int64_t hidden,
                                  int64_t att, const float* W_rv, const float* W_id, const float* h, const float* b_1,
                                  const float* b_2, const float* ebd_vals, int64_t n_ids, int64_t padding_idx,
                                  const float* scores, const float* out_grad, const float* scores_grad, float* feat_grad,
                                  float* W_rv_grad, float* W_id_grad, float* h_grad, float* b_1_grad, float* b_2_grad,
                                  float* ebd_vals_grad, void* stream) {
    RBR_REQUIRE(feat && other_id && W_rv && W_id && h && b_1 && b_2 && ebd_vals && scores && out_grad && feat_grad &&
                    W_rv_grad && W_id_grad && h_grad && b_1_grad && b_2_grad && ebd_vals_grad,
                RBR_EINVAL, "rbr_narre_attn_bwd: null pointer");
    RBR_REQUIRE(batch >= 0 && reviews > 0 && hidden > 0 && att > 0 && att <= 32 * AT_AQ, RBR_EUNSUPPORTED,
                "rbr_narre_attn_bwd: att_dim must be in [1,%d]", 32 * AT_AQ);
    if (batch == 0) return RBR_OK;
    const int R = (int)reviews, H = (int)hidden, A = (int)att;
    const int warps = pick_warps(R, H, A, true);
    RBR_REQUIRE(warps > 0, RBR_EUNSUPPORTED, "rbr_narre_attn_bwd: reviews*hidden too large for shared memory");
    const size_t smem = (size_t)attn_smem(R, H, A, warps, true).total_floats * 4;
    RBR_CUDA(cudaFuncSetAttribute(narre_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (batch + warps - 1) / warps;
    if (blocks > 148 * 2) blocks = 148 * 2;
    narre_attn_bwd_kernel<<<(unsigned)blocks, warps * 32, smem, as_stream(stream)>>>(
        feat, other_id, batch, R, H, A, W_rv, W_id, h, b_1, b_2, ebd_vals, n_ids, padding_idx, scores, out_grad, scores_grad,
        feat_grad, W_rv_grad, W_id_grad, h_grad, b_1_grad, b_2_grad, ebd_vals_grad);
    RBR_LAUNCH_CHECK("narre_attn_bwd_kernel");
    return RBR_OK;
}

RBR_DEFINE_OOB_ACCESSOR(attn)
