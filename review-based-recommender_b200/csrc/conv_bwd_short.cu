// conv_bwd_short.cu — K2b for SHORT documents (NARRE: 10 reviews x 60 tokens per side): the whole arg-max-sparse backward of
// conv → activation → max-over-time → mask → embedding (weight, bias AND table gradient) in ONE kernel, organised per document.
//
// Why a second formulation: with short documents and many filters (60 positions, 150 filters x 3 taps) every token row is used
// ~7.5 times per document, and there are 15x more (doc, filter, tap) entries per batch than in DeepCoNN.  The generic kernels
// (conv_bwd.cu) walk entries filter-major (weight gradient: each (doc, filter) re-reads 3 rows = 1.8 KB from L2, 6.5 GB per
// side → L2-bandwidth bound) and token-major (table gradient: 18 M entries through a counting sort, each reading a 1.2 KB
// weight row).  Here instead:
//   * the embedding dimension is sliced (64 elements per CTA) so that the conv-weight slice [H*k][64] (fp32, 115 KB for
//     H=150, k=3) stays resident in shared memory and the dW slice [H][k][64] stays in REGISTERS (warp w owns filters
//     w, w+16, ...; lane l owns elements 2l, 2l+1) for the whole kernel;
//   * each CTA streams documents: a document's rows are read ONCE per slice from the bf16 shadow / fp32 table into shared
//     memory (software-pipelined one document ahead); its filters are bucketed by arg-max position (150 integer shared
//     atomics + a 70-element scan per document), and the input gradient of each position is then summed in REGISTERS by one
//     warp (dX[p] = sum over taps j and filters h with argmax = p - j of g_h * W[h][j], weight rows from shared memory) and
//     added to the table gradient with one float2 vector atomic per lane — no global entry list, no sort, no per-entry
//     weight-row fetch from L2, no floating-point shared atomics (those compile to CAS loops).
// Same maths as conv_bwd.cu (reference: autograd of models/narre/layers.py:365-401 + nn.Embedding, narre.py:166-179).
#include "rbr_common.cuh"

namespace rbr {

constexpr int BS_WARPS = 16;
constexpr int BS_THREADS = BS_WARPS * 32;
constexpr int BS_ES = 64;                 // embedding elements per slice
constexpr int BS_FPW = 10;                // filters per warp (H <= 160)
constexpr int BS_LMAX = 128;

struct BsSmem {
    int ws, xs, gs, tss, keys, cnt, start, list, total;   // offsets in floats (4-byte words)
};
__host__ __device__ inline int bs_buckets(int K, int L) { return (L + K + 1 + 3) & ~3; }   // bucket b = (argmax - pad) + K
__host__ __device__ inline BsSmem bs_smem(int H, int K, int L) {
    BsSmem s;
    int off = 0;
    s.ws = off; off += H * K * BS_ES;
    s.xs = off; off += 2 * L * BS_ES;
    s.gs = off; off += 2 * ((H + 3) & ~3);
    s.tss = off; off += 2 * ((H + 3) & ~3);
    s.keys = off; off += 2 * ((L + 3) & ~3);
    s.cnt = off; off += bs_buckets(K, L);
    s.start = off; off += bs_buckets(K, L) + 4;
    s.list = off; off += (H + 3) & ~3;
    s.total = off;
    return s;
}

template <int K, bool BF16>
__global__ void __launch_bounds__(BS_THREADS, 1) conv_bwd_short_kernel(
    const float* __restrict__ table, const __nv_bfloat16* __restrict__ shadow, int emb_pad, int64_t vocab, int E,
    const int64_t* __restrict__ ids, const uint8_t* __restrict__ mask, int64_t n_docs, int L, int H, int pad, int64_t padding_idx,
    const float* __restrict__ feat, const int32_t* __restrict__ argmax, const float* __restrict__ feat_grad, int feat_ld, int act,
    const float* __restrict__ whke /* [H][K][epad4] fp32 */, int epad4, int n_groups, int do_weight, int do_table,
    float* __restrict__ dw_hke /* [H][K][epad4], += */, float* __restrict__ bias_grad, float* __restrict__ table_grad) {
    extern __shared__ __align__(16) float smem[];
    const BsSmem S = bs_smem(H, K, L);
    float* Ws = smem + S.ws;
    float* xs = smem + S.xs;
    float* gs = smem + S.gs;
    int* tss = reinterpret_cast<int*>(smem + S.tss);
    int* keys = reinterpret_cast<int*>(smem + S.keys);
    int* cnt = reinterpret_cast<int*>(smem + S.cnt);
    int* start = reinterpret_cast<int*>(smem + S.start);
    int* list = reinterpret_cast<int*>(smem + S.list);
    const int NB = bs_buckets(K, L);
    const int Hp = (H + 3) & ~3, Lp = (L + 3) & ~3;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int slice = blockIdx.x / n_groups, group = blockIdx.x - slice * n_groups;
    const int e0 = slice * BS_ES;                       // first embedding element of this slice
    const int ew = min(BS_ES, E - e0);                  // valid width (last slice may be narrower)

    // resident weight slice (zero beyond the valid width)
    for (int i = tid; i < H * K * BS_ES; i += BS_THREADS) {
        const int c = i % BS_ES, r = i / BS_ES;
        Ws[i] = (do_table && c < ew) ? whke[(int64_t)r * epad4 + e0 + c] : 0.f;
    }
    for (int i = tid; i < NB; i += BS_THREADS) cnt[i] = 0;

    float2 acc[BS_FPW][K];
#pragma unroll
    for (int i = 0; i < BS_FPW; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) acc[i][j] = make_float2(0.f, 0.f);
    float bsum = 0.f;                                   // thread h < H accumulates bias_grad[h] (slice 0 only)

    // staging roles: thread → (row t, 16-byte chunk c of the 128/256-byte row slice); rows beyond L idle
    const int st_row = tid >> 3, st_c = tid & 7;        // 8 threads per row: 8 elements (bf16: one uint4, fp32: two float4) each
    const bool st_on = st_row < L && st_c * 8 < ew;
    // Software pipeline, two documents deep, with NO use of a loaded value in the iteration that issues the load (a use
    // would park the whole CTA on the scoreboard for an L2/HBM round trip per document):
    //   iteration d issues   (a) the raw id / mask loads of doc d+2,  (b) the raw row-chunk and meta loads of doc d+1
    //                        (whose id was loaded one iteration earlier),
    //   and only AFTER computing doc d converts (b) and stores it to the other shared-memory buffer.
    int64_t raw_id = -1;                                // ids[(doc d+2), st_row]      (validated when used)
    uint8_t raw_mk = 0;                                 // mask byte of the same token (1 when there is no mask)
    auto issue_id = [&](int64_t n) {
        raw_id = -1; raw_mk = 0;
        if (n < n_docs && st_row < L) {
            const int64_t q = n * L + st_row;
            raw_id = ids[q];
            raw_mk = mask ? mask[q] : (uint8_t)1;
        }
    };
    auto resolve_id = [&]() -> int64_t { return (raw_mk && raw_id >= 0 && raw_id < vocab) ? raw_id : -1; };
    uint4 raw_a = make_uint4(0, 0, 0, 0), raw_b = make_uint4(0, 0, 0, 0);    // staged row chunk of doc d+1, unconverted
    auto issue_row = [&](int64_t id) {
        raw_a = make_uint4(0, 0, 0, 0); raw_b = make_uint4(0, 0, 0, 0);
        if (id < 0 || !st_on) return;
        if (BF16) {
            raw_a = __ldg(reinterpret_cast<const uint4*>(shadow + id * emb_pad + e0) + st_c);
        } else {
            const uint4* src = reinterpret_cast<const uint4*>(table + id * E + e0) + 2 * st_c;
            raw_a = __ldg(src);
            if (st_c * 8 + 4 < ew) raw_b = __ldg(src + 1);
        }
    };
    float raw_y = 0.f, raw_fg = 0.f;                    // meta of doc d+1 for filter h = tid, uncombined
    int raw_am = 0;
    auto issue_meta = [&](int64_t n) {
        raw_y = 0.f; raw_fg = 0.f; raw_am = 0;
        if (n < n_docs && tid < H) {
            raw_y = __ldg(feat + n * feat_ld + tid);
            raw_fg = __ldg(feat_grad + n * feat_ld + tid);
            raw_am = __ldg(argmax + n * feat_ld + tid);
        }
    };
    auto store_stage = [&](int buf, int64_t id) {
        if (st_row < L) {
            float4 lo, hi;
            if (BF16) {
                lo = make_float4(__uint_as_float(raw_a.x << 16), __uint_as_float(raw_a.x & 0xFFFF0000u), __uint_as_float(raw_a.y << 16),
                                 __uint_as_float(raw_a.y & 0xFFFF0000u));
                hi = make_float4(__uint_as_float(raw_a.z << 16), __uint_as_float(raw_a.z & 0xFFFF0000u), __uint_as_float(raw_a.w << 16),
                                 __uint_as_float(raw_a.w & 0xFFFF0000u));
            } else {
                lo = make_float4(__uint_as_float(raw_a.x), __uint_as_float(raw_a.y), __uint_as_float(raw_a.z), __uint_as_float(raw_a.w));
                hi = make_float4(__uint_as_float(raw_b.x), __uint_as_float(raw_b.y), __uint_as_float(raw_b.z), __uint_as_float(raw_b.w));
            }
            float* dst = xs + (buf * L + st_row) * BS_ES + st_c * 8;
            *reinterpret_cast<float4*>(dst) = lo;
            *reinterpret_cast<float4*>(dst + 4) = hi;
            if (st_c == 0) keys[buf * Lp + st_row] = (id >= 0 && id != padding_idx) ? (int)id : -1;
        }
        if (tid < H) {
            gs[buf * Hp + tid] = raw_fg * act_grad_from_out(act, raw_y);
            tss[buf * Hp + tid] = raw_am - pad;
        }
    };

    // prologue: doc0 staged, doc1's ids in flight
    const int64_t n0 = group;
    issue_id(n0);
    int64_t id_next = resolve_id();
    issue_row(id_next);
    issue_meta(n0);
    issue_id(n0 + n_groups);
    store_stage(0, id_next);
    // a document with no valid token (NARRE pads users/items to 10 reviews: ~45 % of the "documents" are all padding) only
    // contributes to the bias gradient: the barrier that publishes a staged document also tells everyone whether it is live
    int live = __syncthreads_or(id_next >= 0);

    int buf = 0;
    for (int64_t n = n0; n < n_docs; n += n_groups) {
        // ---- issue the loads of doc n + G (rows, meta) and of the ids of doc n + 2G; nothing below uses them before store_stage
        const int64_t id_cur_next = resolve_id();       // id of (doc n + G, st_row): loaded one iteration ago
        issue_row(id_cur_next);
        issue_meta(n + n_groups);
        issue_id(n + 2 * (int64_t)n_groups);
        // ---- compute doc n from buffer `buf`
        if (tid < H && slice == 0) bsum += gs[buf * Hp + tid];
        const float* xb = xs + buf * L * BS_ES;
        // (1) bucket the filters by arg-max start position (table part)
        int my_b = -1, my_slot = 0;
        if (do_table && live && tid < H && gs[buf * Hp + tid] != 0.f) {
            my_b = tss[buf * Hp + tid] + K;                        // ts >= -pad > -K  →  b >= 1
            my_slot = atomicAdd(cnt + my_b, 1);
        }
        __syncthreads();
        // (2) warp 0 scans the bucket counts; every warp accumulates the weight gradient of its filters meanwhile
        if (do_table && live && warp == 0) {
            int carry = 0;
            for (int base = 0; base < NB; base += 32) {
                const int c = (base + lane < NB) ? cnt[base + lane] : 0;
                int x = c;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int y = __shfl_up_sync(0xffffffffu, x, o);
                    if (lane >= o) x += y;
                }
                if (base + lane < NB) { start[base + lane] = carry + x - c; cnt[base + lane] = 0; }
                carry += __shfl_sync(0xffffffffu, x, 31);
            }
            if (lane == 0) start[NB] = carry;
        }
        if (do_weight && live) {
#pragma unroll
            for (int i = 0; i < BS_FPW; ++i) {
                const int h = warp + BS_WARPS * i;
                if (h < H) {
                    const float g = gs[buf * Hp + h];
                    if (g != 0.f) {
                        const int ts = tss[buf * Hp + h];
#pragma unroll
                        for (int j = 0; j < K; ++j) {
                            const int t = ts + j;
                            if (t >= 0 && t < L) {
                                const float2 x = *reinterpret_cast<const float2*>(xb + t * BS_ES + 2 * lane);
                                acc[i][j].x = fmaf(g, x.x, acc[i][j].x);
                                acc[i][j].y = fmaf(g, x.y, acc[i][j].y);
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
        if (my_b >= 0) list[start[my_b] + my_slot] = tid;
        __syncthreads();
        // (3) input gradient per position, summed in registers, one float2 vector atomic per lane into the table gradient
        if (do_table && live) {
            for (int p = warp; p < L; p += BS_WARPS) {
                const int key = keys[buf * Lp + p];
                if (key < 0) continue;                               // masked / padding / invalid token: no table row
                float2 dx = make_float2(0.f, 0.f);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    const int b = p - j + K;                         // filters whose window starts at p - j reach p with tap j
                    if (b < 1 || b >= NB) continue;
                    const int i0 = start[b], i1 = start[b + 1];
                    for (int q = i0; q < i1; ++q) {
                        const int h = list[q];
                        const float g = gs[buf * Hp + h];
                        const float2 w = *reinterpret_cast<const float2*>(Ws + (h * K + j) * BS_ES + 2 * lane);
                        dx.x = fmaf(g, w.x, dx.x);
                        dx.y = fmaf(g, w.y, dx.y);
                    }
                }
                if (2 * lane < ew && (dx.x != 0.f || dx.y != 0.f))
                    atomicAdd(reinterpret_cast<float2*>(table_grad + (int64_t)key * E + e0) + lane, dx);
            }
        }
        store_stage(buf ^ 1, id_cur_next);
        buf ^= 1;
        live = __syncthreads_or(id_cur_next >= 0);
    }
    // ---- flush the register accumulators
    if (do_weight) {
#pragma unroll
        for (int i = 0; i < BS_FPW; ++i) {
            const int h = warp + BS_WARPS * i;
            if (h < H) {
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    float* dst = dw_hke + ((int64_t)h * K + j) * epad4 + e0;
                    if (2 * lane < ew && acc[i][j].x != 0.f) atomicAdd(dst + 2 * lane, acc[i][j].x);
                    if (2 * lane + 1 < ew && acc[i][j].y != 0.f) atomicAdd(dst + 2 * lane + 1, acc[i][j].y);
                }
            }
        }
        if (slice == 0 && tid < H && bsum != 0.f) atomicAdd(bias_grad + tid, bsum);
    }
}

// true when the per-document kernel can take this shape
bool conv_bwd_short_ok(int E, int H, int K, int L, int gate_mode, int64_t vocab) {
    if (gate_mode != 0 || L > BS_LMAX || L * 8 > BS_THREADS || H > BS_WARPS * BS_FPW || H > BS_THREADS) return false;
    if (E % 8 != 0 && E % 4 != 0) return false;
    if (!(K == 1 || K == 2 || K == 3 || K == 4 || K == 5 || K == 7)) return false;
    if (vocab >= (1ll << 31)) return false;
    return (size_t)bs_smem(H, K, L).total * 4 <= 220 * 1024;
}

template <int K>
static int launch_short(bool bf16, const float* table, const __nv_bfloat16* shadow, int emb_pad, int64_t vocab, int E,
                        const int64_t* ids, const uint8_t* mask, int64_t n_docs, int L, int H, int pad, int64_t padding_idx,
                        const float* feat, const int32_t* argmax, const float* feat_grad, int feat_ld, int act, const float* whke,
                        int epad4, int do_weight, int do_table, float* dw_hke, float* bias_grad, float* table_grad, cudaStream_t s) {
    const int slices = (E + BS_ES - 1) / BS_ES;
    int groups = 148 / slices;                       // one CTA per SM (shared memory), a whole number of groups per slice
    if (groups > n_docs) groups = (int)n_docs;
    if (groups < 1) groups = 1;
    const size_t smem = (size_t)bs_smem(H, K, L).total * 4;
    auto kern = bf16 ? conv_bwd_short_kernel<K, true> : conv_bwd_short_kernel<K, false>;
    RBR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<slices * groups, BS_THREADS, smem, s>>>(table, shadow, emb_pad, vocab, E, ids, mask, n_docs, L, H, pad, padding_idx, feat,
                                                   argmax, feat_grad, feat_ld, act, whke, epad4, groups, do_weight, do_table, dw_hke,
                                                   bias_grad, table_grad);
    RBR_LAUNCH_CHECK("conv_bwd_short_kernel");
    return RBR_OK;
}

int conv_bwd_short_dispatch(bool bf16, const float* table, const __nv_bfloat16* shadow, int emb_pad, int64_t vocab, int E,
                            const int64_t* ids, const uint8_t* mask, int64_t n_docs, int L, int H, int K, int pad,
                            int64_t padding_idx, const float* feat, const int32_t* argmax, const float* feat_grad, int feat_ld,
                            int act, const float* whke, int epad4, int do_weight, int do_table, float* dw_hke, float* bias_grad,
                            float* table_grad, cudaStream_t s) {
#define RBR_BS(K_)                                                                                                            \
    case K_:                                                                                                                  \
        return launch_short<K_>(bf16, table, shadow, emb_pad, vocab, E, ids, mask, n_docs, L, H, pad, padding_idx, feat, argmax, \
                                feat_grad, feat_ld, act, whke, epad4, do_weight, do_table, dw_hke, bias_grad, table_grad, s);
    switch (K) {
        RBR_BS(1) RBR_BS(2) RBR_BS(3) RBR_BS(4) RBR_BS(5) RBR_BS(7)
        default: break;
    }
#undef RBR_BS
    set_error("conv_bwd_short: unsupported kernel size %d", K);
    return RBR_EUNSUPPORTED;
}

}  // namespace rbr
