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


// ------------------------------------------------------------------------------------------------------------------
// Register-tiled variant for att_dim <= 32 (the reference's att_dim is 32, default_narre.json:17).  Lane = attention
// dim a.  One warp per sample; the R review rows are processed RC at a time with RC accumulators per lane, so each
// W_rv / W_id element is loaded once per RC rows and the feature rows are read as broadcast float4s (≈1.4 shared-memory
// instructions per FMA-group instead of 2 per FMA).  Backward: the weight gradients dW_rv = F^T dHid, dW_id = E^T dHid are
// accumulated in REGISTERS by all 256 threads of the CTA in a cooperative phase after each round of 8 samples
// (thread (warp w, lane a) owns rows h in [20w, 20w+20) of dW_rv and a2 in [4w, 4w+4) of dW_id) — the per-sample
// shared-memory atomics of the generic kernel were 80 % of its time.
// ------------------------------------------------------------------------------------------------------------------
constexpr int A2_WARPS = 8;
constexpr int A2_RC = 5;          // review rows per register tile
constexpr int A2_HW = 20;         // dW_rv rows owned per warp in the cooperative phase (8 * 20 = 160 >= hidden)

struct Attn2Smem {
    int H4, wrv, wid, per_warp, fs, es, dhs, ls, og, total_floats;
};
__host__ __device__ inline Attn2Smem attn2_smem(int R, int H, int A, bool bwd) {
    Attn2Smem s;
    s.H4 = (H + 3) & ~3;
    const int Hc = s.H4 > A2_WARPS * A2_HW ? s.H4 : A2_WARPS * A2_HW;     // cooperative phase reads 20 columns per warp
    int off = 0;
    s.wrv = off; off += H * 33;
    s.wid = off; off += A * 33;
    off = (off + 3) & ~3;
    int pw = 0;
    s.fs = pw; pw += R * Hc;                  // row pitch Hc (16-byte aligned rows, zero padded)
    s.es = pw; pw += R * 32;
    s.dhs = pw; if (bwd) pw += R * 32;
    s.ls = pw; pw += ((2 * R + 3) & ~3);
    s.og = pw; if (bwd) pw += s.H4;
    s.per_warp = (pw + 3) & ~3;
    s.total_floats = off + A2_WARPS * s.per_warp;
    return s;
}
__host__ __device__ inline int attn2_pitch(int H) {
    const int H4 = (H + 3) & ~3;
    return H4 > A2_WARPS * A2_HW ? H4 : A2_WARPS * A2_HW;
}

// hid[r] (r < R, RC at a time) for lane a: relu(b1[a] + fs[r]·W_rv[:,a] + es[r]·W_id[:,a]); calls f(r, hid)
template <typename F>
__device__ __forceinline__ void attn2_hidden(const float* fs, int pitch, const float* es, const float* Wrv_s, const float* Wid_s,
                                             float b1a, int R, int H, int A, int lane, F&& f) {
    const int a = lane < A ? lane : 0;
    for (int r0 = 0; r0 < R; r0 += A2_RC) {
        float acc[A2_RC];
#pragma unroll
        for (int i = 0; i < A2_RC; ++i) acc[i] = 0.f;
        int h = 0;
        for (; h + 4 <= H; h += 4) {
            const float w0 = Wrv_s[h * 33 + a], w1 = Wrv_s[(h + 1) * 33 + a], w2 = Wrv_s[(h + 2) * 33 + a], w3 = Wrv_s[(h + 3) * 33 + a];
#pragma unroll
            for (int i = 0; i < A2_RC; ++i) {
                const int r = r0 + i < R ? r0 + i : R - 1;
                const float4 x = *reinterpret_cast<const float4*>(fs + r * pitch + h);
                acc[i] = fmaf(x.x, w0, acc[i]); acc[i] = fmaf(x.y, w1, acc[i]);
                acc[i] = fmaf(x.z, w2, acc[i]); acc[i] = fmaf(x.w, w3, acc[i]);
            }
        }
        for (; h < H; ++h) {
            const float w0 = Wrv_s[h * 33 + a];
#pragma unroll
            for (int i = 0; i < A2_RC; ++i) {
                const int r = r0 + i < R ? r0 + i : R - 1;
                acc[i] = fmaf(fs[r * pitch + h], w0, acc[i]);
            }
        }
        for (int a2 = 0; a2 < A; ++a2) {
            const float w0 = Wid_s[a2 * 33 + a];
#pragma unroll
            for (int i = 0; i < A2_RC; ++i) {
                const int r = r0 + i < R ? r0 + i : R - 1;
                acc[i] = fmaf(es[r * 32 + a2], w0, acc[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < A2_RC; ++i)
            if (r0 + i < R) f(r0 + i, lane < A ? fmaxf(acc[i] + b1a, 0.f) : 0.f);
    }
}

__device__ __forceinline__ void attn2_stage(const float* __restrict__ feat, const int64_t* __restrict__ other_id,
                                            const float* __restrict__ ebd, int64_t n_ids, int64_t b, int R, int H, int A, int pitch,
                                            float* fs, float* es, int lane, bool count_oob) {
    // feature rows: coalesced global reads, zero padding up to the pitch
    for (int r = 0; r < R; ++r) {
        const float* src = feat + (b * R + r) * H;
        for (int h = lane; h < pitch; h += 32) fs[r * pitch + h] = h < H ? src[h] : 0.f;
    }
    for (int i = lane; i < R * 32; i += 32) {
        const int r = i >> 5, a = i & 31;
        const int64_t id = other_id[b * R + r];
        const bool ok = id >= 0 && id < n_ids;
        if (count_oob && !ok && a == 0) note_oob();
        es[i] = (ok && a < A) ? ebd[id * A + a] : 0.f;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(A2_WARPS * 32) narre_attn2_fwd_kernel(
    const float* __restrict__ feat, const int64_t* __restrict__ other_id, int64_t B, int R, int H, int A,
    const float* __restrict__ W_rv, const float* __restrict__ W_id, const float* __restrict__ hvec, const float* __restrict__ b1,
    const float* __restrict__ b2, const float* __restrict__ ebd, int64_t n_ids, float* __restrict__ out, float* __restrict__ scores) {
    extern __shared__ __align__(16) float smem[];
    const Attn2Smem L = attn2_smem(R, H, A, false);
    const int pitch = attn2_pitch(H);
    float* Wrv_s = smem + L.wrv;
    float* Wid_s = smem + L.wid;
    for (int i = threadIdx.x; i < H * 33; i += blockDim.x) Wrv_s[i] = (i % 33 < A) ? W_rv[(i / 33) * A + i % 33] : 0.f;
    for (int i = threadIdx.x; i < A * 33; i += blockDim.x) Wid_s[i] = (i % 33 < A) ? W_id[(i / 33) * A + i % 33] : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* pw = smem + (L.total_floats - A2_WARPS * L.per_warp) + wib * L.per_warp;
    float* fs = pw + L.fs; float* es = pw + L.es; float* ls = pw + L.ls;
    const float b2v = b2[0];
    const float b1a = lane < A ? b1[lane] : 0.f, hva = lane < A ? hvec[lane] : 0.f;
    for (int64_t b = (int64_t)blockIdx.x * A2_WARPS + wib; b < B; b += (int64_t)gridDim.x * A2_WARPS) {
        attn2_stage(feat, other_id, ebd, n_ids, b, R, H, A, pitch, fs, es, lane, true);
        attn2_hidden(fs, pitch, es, Wrv_s, Wid_s, b1a, R, H, A, lane, [&](int r, float hid) {
            const float part = warp_sum(hid * hva);
            if (lane == 0) ls[r] = part + b2v;
        });
        __syncwarp();
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
            for (int r = 0; r < R; ++r) acc = fmaf(ls[R + r], fs[r * pitch + h], acc);
            out[b * H + h] = acc;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(A2_WARPS * 32) narre_attn2_bwd_kernel(
    const float* __restrict__ feat, const int64_t* __restrict__ other_id, int64_t B, int R, int H, int A,
    const float* __restrict__ W_rv, const float* __restrict__ W_id, const float* __restrict__ hvec, const float* __restrict__ b1,
    const float* __restrict__ ebd, int64_t n_ids, int64_t padding_idx, const float* __restrict__ scores,
    const float* __restrict__ out_grad, const float* __restrict__ scores_grad, float* __restrict__ feat_grad,
    float* __restrict__ W_rv_grad, float* __restrict__ W_id_grad, float* __restrict__ h_grad, float* __restrict__ b1_grad,
    float* __restrict__ b2_grad, float* __restrict__ ebd_grad) {
    extern __shared__ __align__(16) float smem[];
    const Attn2Smem L = attn2_smem(R, H, A, true);
    const int pitch = attn2_pitch(H);
    float* Wrv_s = smem + L.wrv;
    float* Wid_s = smem + L.wid;
    for (int i = threadIdx.x; i < H * 33; i += blockDim.x) Wrv_s[i] = (i % 33 < A) ? W_rv[(i / 33) * A + i % 33] : 0.f;
    for (int i = threadIdx.x; i < A * 33; i += blockDim.x) Wid_s[i] = (i % 33 < A) ? W_id[(i / 33) * A + i % 33] : 0.f;
    // the cooperative phase reads every warp's staging area, also of warps that never staged a sample: start from zeros
    for (int i = threadIdx.x; i < A2_WARPS * L.per_warp; i += blockDim.x) smem[L.total_floats - A2_WARPS * L.per_warp + i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    float* pw_base = smem + (L.total_floats - A2_WARPS * L.per_warp);
    float* pw = pw_base + wib * L.per_warp;
    float* fs = pw + L.fs; float* es = pw + L.es; float* dhs = pw + L.dhs; float* ls = pw + L.ls; float* og = pw + L.og;
    const float b1a = lane < A ? b1[lane] : 0.f, hva = lane < A ? hvec[lane] : 0.f;
    float dh_acc = 0.f, db1_acc = 0.f, db2_acc = 0.f;
    float dwrv[A2_HW], dwid[4];
#pragma unroll
    for (int i = 0; i < A2_HW; ++i) dwrv[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) dwid[i] = 0.f;

    const int64_t rounds = (B + (int64_t)gridDim.x * A2_WARPS - 1) / ((int64_t)gridDim.x * A2_WARPS);
    for (int64_t rd = 0; rd < rounds; ++rd) {
        const int64_t b = (rd * gridDim.x + blockIdx.x) * A2_WARPS + wib;
        const bool live = b < B;
        if (live) {
            attn2_stage(feat, other_id, ebd, n_ids, b, R, H, A, pitch, fs, es, lane, false);
            for (int h = lane; h < L.H4; h += 32) og[h] = h < H ? out_grad[b * H + h] : 0.f;
            __syncwarp();
            // ds[r] = out_grad · feat[r] (+ scores_grad[r]);  dl[r] = score[r] * (ds[r] - sum_r' score[r'] ds[r'])
            for (int r = 0; r < R; ++r) {
                float p = 0.f;
                for (int h = lane; h < H; h += 32) p = fmaf(og[h], fs[r * pitch + h], p);
                p = warp_sum(p);
                if (lane == 0) ls[r] = p + (scores_grad ? scores_grad[b * R + r] : 0.f);
            }
            __syncwarp();
            float dot = 0.f;
            for (int r = lane; r < R; r += 32) dot = fmaf(ls[r], scores[b * R + r], dot);
            dot = warp_sum(dot);
            for (int r = lane; r < R; r += 32) {
                const float sc = scores[b * R + r];
                const float dl = sc * (ls[r] - dot);
                ls[r] = sc;                                       // ls[0..R) = score, ls[R..2R) = d loss / d logit
                ls[R + r] = dl;
                db2_acc += dl;
            }
            __syncwarp();
            // recompute hid, then d hid[r][a] = (hid > 0) * dl[r] * h[a]  → dhs;  d h, d b1
            attn2_hidden(fs, pitch, es, Wrv_s, Wid_s, b1a, R, H, A, lane, [&](int r, float hid) {
                const float dl = ls[R + r];
                dh_acc = fmaf(dl, hid, dh_acc);
                const float dhid = hid > 0.f ? dl * hva : 0.f;
                dhs[r * 32 + lane] = dhid;
                db1_acc += dhid;
            });
            __syncwarp();
            // d feat[r][h] = score[r] * out_grad[h] + sum_a dhid[r][a] * W_rv[h][a]       (lane = h, RC rows at a time)
            for (int h0 = 0; h0 < H; h0 += 32) {
                const int h = h0 + lane < H ? h0 + lane : H - 1;
                const float ogh = og[h];
                for (int r0 = 0; r0 < R; r0 += A2_RC) {
                    float acc[A2_RC];
#pragma unroll
                    for (int i = 0; i < A2_RC; ++i) acc[i] = 0.f;
                    for (int a = 0; a < 32; a += 4) {              // dhs is zero for a >= A
                        const float w0 = Wrv_s[h * 33 + a], w1 = Wrv_s[h * 33 + a + 1], w2 = Wrv_s[h * 33 + a + 2], w3 = Wrv_s[h * 33 + a + 3];
#pragma unroll
                        for (int i = 0; i < A2_RC; ++i) {
                            const int r = r0 + i < R ? r0 + i : R - 1;
                            const float4 d = *reinterpret_cast<const float4*>(dhs + r * 32 + a);
                            acc[i] = fmaf(d.x, w0, acc[i]); acc[i] = fmaf(d.y, w1, acc[i]);
                            acc[i] = fmaf(d.z, w2, acc[i]); acc[i] = fmaf(d.w, w3, acc[i]);
                        }
                    }
                    if (h0 + lane < H) {
#pragma unroll
                        for (int i = 0; i < A2_RC; ++i)
                            if (r0 + i < R) feat_grad[(b * R + r0 + i) * H + h] = fmaf(ls[r0 + i], ogh, acc[i]);
                    }
                }
            }
            // d e[r][a2] = sum_a dhid[r][a] * W_id[a2][a] → id-embedding rows (padding row skipped)
            if (lane < A) {
                for (int r = 0; r < R; ++r) {
                    const int64_t id = other_id[b * R + r];
                    if (id < 0 || id >= n_ids || id == padding_idx) continue;
                    float acc = 0.f;
                    for (int a = 0; a < A; ++a) acc = fmaf(dhs[r * 32 + a], Wid_s[lane * 33 + a], acc);
                    if (acc != 0.f) atomicAdd(ebd_grad + id * A + lane, acc);
                }
            }
        } else {
            for (int i = lane; i < R * 32; i += 32) dhs[i] = 0.f;  // idle warp of a partial round contributes nothing below
        }
        __syncthreads();
        // cooperative weight-gradient phase over the (up to) 8 samples staged by the CTA's warps
        for (int sw = 0; sw < A2_WARPS; ++sw) {
            const float* sp = pw_base + sw * L.per_warp;
            const float* sfs = sp + L.fs; const float* ses = sp + L.es; const float* sdh = sp + L.dhs;
            for (int r = 0; r < R; ++r) {
                const float d = sdh[r * 32 + lane];
#pragma unroll
                for (int q = 0; q < A2_HW / 4; ++q) {
                    const float4 x = *reinterpret_cast<const float4*>(sfs + r * pitch + wib * A2_HW + 4 * q);
                    dwrv[4 * q] = fmaf(x.x, d, dwrv[4 * q]); dwrv[4 * q + 1] = fmaf(x.y, d, dwrv[4 * q + 1]);
                    dwrv[4 * q + 2] = fmaf(x.z, d, dwrv[4 * q + 2]); dwrv[4 * q + 3] = fmaf(x.w, d, dwrv[4 * q + 3]);
                }
                const float4 e = *reinterpret_cast<const float4*>(ses + r * 32 + wib * 4);
                dwid[0] = fmaf(e.x, d, dwid[0]); dwid[1] = fmaf(e.y, d, dwid[1]);
                dwid[2] = fmaf(e.z, d, dwid[2]); dwid[3] = fmaf(e.w, d, dwid[3]);
            }
        }
        __syncthreads();
    }
    if (lane < A) {
#pragma unroll
        for (int i = 0; i < A2_HW; ++i) {
            const int h = wib * A2_HW + i;
            if (h < H && dwrv[i] != 0.f) atomicAdd(W_rv_grad + h * A + lane, dwrv[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int a2 = wib * 4 + i;
            if (a2 < A && dwid[i] != 0.f) atomicAdd(W_id_grad + a2 * A + lane, dwid[i]);
        }
        if (dh_acc != 0.f) atomicAdd(h_grad + lane, dh_acc);
        if (db1_acc != 0.f) atomicAdd(b1_grad + lane, db1_acc);
    }
    db2_acc = warp_sum(db2_acc);
    if (lane == 0 && db2_acc != 0.f) atomicAdd(b2_grad, db2_acc);
}

static bool attn2_ok(int R, int H, int A, bool bwd) {
    return A <= 32 && H <= A2_WARPS * A2_HW && (size_t)attn2_smem(R, H, A, bwd).total_floats * 4 <= 200 * 1024;
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
    if (attn2_ok(R, H, A, false)) {
        const size_t smem2 = (size_t)attn2_smem(R, H, A, false).total_floats * 4;
        RBR_CUDA(cudaFuncSetAttribute(narre_attn2_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        int64_t blocks2 = (batch + A2_WARPS - 1) / A2_WARPS;
        if (blocks2 > 148 * 2) blocks2 = 148 * 2;
        narre_attn2_fwd_kernel<<<(unsigned)blocks2, A2_WARPS * 32, smem2, as_stream(stream)>>>(feat, other_id, batch, R, H, A, W_rv, W_id,
                                                                                           h, b_1, b_2, ebd_vals, n_ids, out, scores);
        RBR_LAUNCH_CHECK("narre_attn2_fwd_kernel");
        return RBR_OK;
    }
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
    if (attn2_ok(R, H, A, true)) {
        const size_t smem2 = (size_t)attn2_smem(R, H, A, true).total_floats * 4;
        RBR_CUDA(cudaFuncSetAttribute(narre_attn2_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        int64_t blocks2 = (batch + A2_WARPS - 1) / A2_WARPS;
        if (blocks2 > 148 * 2) blocks2 = 148 * 2;
        narre_attn2_bwd_kernel<<<(unsigned)blocks2, A2_WARPS * 32, smem2, as_stream(stream)>>>(
            feat, other_id, batch, R, H, A, W_rv, W_id, h, b_1, ebd_vals, n_ids, padding_idx, scores, out_grad, scores_grad, feat_grad,
            W_rv_grad, W_id_grad, h_grad, b_1_grad, b_2_grad, ebd_vals_grad);
        RBR_LAUNCH_CHECK("narre_attn2_bwd_kernel");
        return RBR_OK;
    }
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
