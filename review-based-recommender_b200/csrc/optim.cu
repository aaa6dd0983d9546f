// optim.cu — K7: fused global-norm gradient clipping + Adam over the flat parameter / gradient arenas.
//
// Replaces, per step, nn.utils.clip_grad_norm_(model.parameters(), max_norm) followed by torch.optim.Adam.step()
// (reference trainer/train_deepconn_pp.py:135,167-168): ~15 launches and ≈ 420 MB of HBM traffic in the reference
// (norm per tensor, stack, norm, clamp, mul_ per tensor, then the multi-tensor Adam passes).  Here every parameter is a view
// of ONE flat fp32 buffer laid out like the gradient arena, so the step is
//   sumsq_kernel      one pass over the gradients → Σ g²  (fp32 per thread, fp64 across blocks)
//   clip_adam_kernel  one pass: g·clip → m, v, p updated in place; the word-table slice also writes its bf16 shadow row
//                     (the operand the tensor-core conv gathers), which removes rbr_table_to_bf16 from the step.
// clip = min(1, max_norm / (‖g‖₂ + 1e-6)) exactly as clip_grad_norm_; Adam as torch.optim.Adam (no amsgrad, no weight decay):
//   m = β1 m + (1-β1) g;  v = β2 v + (1-β2) g²;  p -= lr / (1-β1^t) · m / (sqrt(v) / sqrt(1-β2^t) + eps)
// The step count t lives on the device (bumped by the kernel), so a CUDA-graph replay of the whole trainer step advances it.
#include "rbr_common.cuh"

namespace rbr {

__global__ void __launch_bounds__(512) sumsq_kernel(const float4* __restrict__ g, int64_t n4, double* __restrict__ out) {
    float acc = 0.f;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = __ldg(g + i);
        acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc); acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
    }
    double d = (double)warp_sum(acc);
    __shared__ double part[16];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) part[w] = d;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += part[i];
        if (s != 0.0) atomicAdd(out, s);
    }
}

struct AdamArgs {
    float lr, beta1, beta2, eps, max_norm;   // max_norm <= 0: no clipping
    const double* sumsq;                      // Σ g² (device); may be null when max_norm <= 0
    int64_t* step;                            // device step counter t (already bumped for this step by adam_tick_kernel)
    // word-table slice of the flat buffers → bf16 shadow (optional)
    int64_t table_off, table_numel;
    int emb, emb_pad;
    __nv_bfloat16* shadow;
};

__global__ void adam_tick_kernel(int64_t* step, float* gnorm_out, const double* sumsq) {
    *step += 1;
    if (gnorm_out && sumsq) *gnorm_out = (float)sqrt(*sumsq);
}

__global__ void __launch_bounds__(512) clip_adam_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m,
                                                        float4* __restrict__ v, int64_t n4, const AdamArgs a) {
    __shared__ float s_clip, s_step_size, s_inv_bc2;
    if (threadIdx.x == 0) {
        float clip = 1.f;
        if (a.max_norm > 0.f && a.sumsq) {
            const float norm = (float)sqrt(*a.sumsq);
            clip = fminf(1.f, a.max_norm / (norm + 1e-6f));
        }
        const double t = (double)*a.step;
        const double bc1 = 1.0 - pow((double)a.beta1, t), bc2 = 1.0 - pow((double)a.beta2, t);
        s_clip = clip;
        s_step_size = (float)((double)a.lr / bc1);
        s_inv_bc2 = (float)(1.0 / sqrt(bc2));
    }
    __syncthreads();
    const float clip = s_clip, step_size = s_step_size, inv_sqrt_bc2 = s_inv_bc2;
    const float b1 = a.beta1, b2 = a.beta2, ob1 = 1.f - a.beta1, ob2 = 1.f - a.beta2;
    const int64_t t4lo = a.table_off >> 2, t4hi = (a.table_off + a.table_numel) >> 2;
    const int e4 = a.emb >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 gv = __ldg(g + i);
        float4 pv = p[i], mv = m[i], vv = v[i];
#define RBR_ADAM1(c)                                                         \
    {                                                                        \
        const float gg = gv.c * clip;                                        \
        mv.c = b1 * mv.c + ob1 * gg;                                         \
        vv.c = b2 * vv.c + ob2 * gg * gg;                                    \
        pv.c -= step_size * (mv.c / (sqrtf(vv.c) * inv_sqrt_bc2 + a.eps));   \
    }
        RBR_ADAM1(x) RBR_ADAM1(y) RBR_ADAM1(z) RBR_ADAM1(w)
#undef RBR_ADAM1
        p[i] = pv; m[i] = mv; v[i] = vv;
        if (a.shadow && i >= t4lo && i < t4hi) {                  // table element (row, 4 columns) → its bf16 shadow
            const int64_t q = i - t4lo;
            const int64_t row = q / e4;
            const int c = (int)(q - row * e4) * 4;
            const __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo); pk.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(a.shadow + row * a.emb_pad + c) = pk;
        }
    }
}

}  // namespace rbr

using namespace rbr;

extern "C" int rbr_clip_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n_floats, float lr,
                                  float beta1, float beta2, float eps, float max_norm, double* sumsq_dev, int64_t* step_dev,
                                  float* grad_norm_out, int64_t table_off, int64_t table_rows, int64_t emb, void* shadow_bf16,
                                  void* stream) {
    RBR_REQUIRE(params && grads && exp_avg && exp_avg_sq && step_dev, RBR_EINVAL, "rbr_clip_adam_step: null pointer");
    RBR_REQUIRE(n_floats >= 0 && n_floats % 4 == 0, RBR_EINVAL, "rbr_clip_adam_step: the flat buffers must be a multiple of 4 floats");
    RBR_REQUIRE(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, RBR_EINVAL,
                "rbr_clip_adam_step: buffers must be 16-byte aligned");
    RBR_REQUIRE(max_norm <= 0.f || sumsq_dev, RBR_EINVAL, "rbr_clip_adam_step: clipping needs the sumsq scratch");
    RBR_REQUIRE(!shadow_bf16 || (emb > 0 && emb % 4 == 0 && table_off % 4 == 0 && table_off >= 0 &&
                                 table_off + table_rows * emb <= n_floats),
                RBR_EINVAL, "rbr_clip_adam_step: bad table slice");
    if (n_floats == 0) return RBR_OK;
    cudaStream_t s = as_stream(stream);
    const int64_t n4 = n_floats / 4;
    int64_t blocks = (n4 + 512 * 4 - 1) / (512 * 4);
    if (blocks > 148 * 4) blocks = 148 * 4;
    if (blocks < 1) blocks = 1;
    if (max_norm > 0.f || grad_norm_out) {
        RBR_REQUIRE(sumsq_dev, RBR_EINVAL, "rbr_clip_adam_step: the gradient norm needs the sumsq scratch");
        RBR_CUDA(cudaMemsetAsync(sumsq_dev, 0, sizeof(double), s));
        sumsq_kernel<<<(unsigned)blocks, 512, 0, s>>>(reinterpret_cast<const float4*>(grads), n4, sumsq_dev);
        RBR_LAUNCH_CHECK("sumsq_kernel");
    }
    adam_tick_kernel<<<1, 1, 0, s>>>(step_dev, grad_norm_out, sumsq_dev);
    RBR_LAUNCH_CHECK("adam_tick_kernel");
    AdamArgs a{};
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.max_norm = max_norm; a.sumsq = sumsq_dev; a.step = step_dev;
    a.table_off = table_off; a.table_numel = shadow_bf16 ? table_rows * emb : 0; a.emb = (int)(emb > 0 ? emb : 4);
    a.emb_pad = (int)rbr_emb_pad(emb > 0 ? emb : 4); a.shadow = reinterpret_cast<__nv_bfloat16*>(shadow_bf16);
    clip_adam_kernel<<<(unsigned)blocks, 512, 0, s>>>(reinterpret_cast<float4*>(params), reinterpret_cast<const float4*>(grads),
                                                      reinterpret_cast<float4*>(exp_avg), reinterpret_cast<float4*>(exp_avg_sq), n4, a);
    RBR_LAUNCH_CHECK("clip_adam_kernel");
    return RBR_OK;
}
