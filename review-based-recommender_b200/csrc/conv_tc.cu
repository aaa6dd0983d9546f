// conv_tc.cu — K2 (bf16 tensor-core variant): gather → mask → Conv1d → bias → activation → max-over-time as an
// implicit GEMM on tcgen05 (5th-gen tensor cores), fp32 accumulators in TMEM.  sm_100a only.
//
// Replaces, for one batch of documents, nn.Embedding + masked_fill + transpose + nn.Conv1d + nn.ReLU +
// nn.MaxPool1d (reference models/deepconn/layers.py:22-24,123-136; cuDNN implicit GEMM + 4 elementwise passes over
// [N,L,E] / [N,H,L] tensors).  Nothing of size N*L*E or N*H*L is written to HBM.
//
// GEMM view (per 128-position tile of a document):  D[128 x Nb] = sum_{tap j} sum_{e}  X[t+j, e] * W_j[h, e]
//   A operand = gathered bf16 token rows, staged by cp.async into shared memory in the UMMA "K-major, no swizzle"
//               core-matrix layout  As[chunk c = e/8][row r][8 bf16]  (16 B per (c,r); row stride 16 B, chunk stride RS*16 B).
//               Because consecutive rows are 16 B apart, the operand for tap j is the SAME tile with the descriptor start
//               address advanced by j rows: the k taps need no im2col, no re-staging and no extra traffic.
//   B operand = conv weights, bf16, same layout  Ws[tap j][chunk c][filter n][8 bf16]; loaded ONCE per CTA (persistent
//               kernel) by cp.async.bulk (TMA bulk copy) and kept resident for every tile the CTA processes.
//   D         = 128 lanes x Nb columns fp32 in TMEM, double buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
// Warp roles (288 threads): warps 0-3 epilogue (TMEM → registers → max-over-time), warps 4-7 gather producers,
// warp 8 = TMEM allocator + single-thread tcgen05.mma issuer.  mbarrier pipelines: full/empty per ring stage,
// acc_full/acc_empty per TMEM buffer.
//
// Max-over-time without shuffling floats: positions are TMEM lanes, so the max over positions is a max over the 32 threads
// of a warp.  Each fp32 accumulator is mapped to an order-preserving uint32; one redux.sync.max.u32 per column gives the
// exact max, a ballot + ffs gives the FIRST position attaining it (nn.MaxPool1d's tie rule), and the (value, ~position)
// pair is merged across warps and across the tiles of a document with a 64-bit shared-memory atomicMax.  The pooled value
// is therefore the exact fp32 accumulator.  Bias and activation are applied once per (doc, filter) after the max (both
// monotone), not per position.
#include "rbr_common.cuh"
#include "tc_ptx.cuh"

namespace rbr {

// ------------------------------------------------------------------------------------------------
// host-side plan
// ------------------------------------------------------------------------------------------------
constexpr int TC_SMEM_MAX = 232448;        // 227 KB opt-in dynamic shared memory per CTA
constexpr int TC_M = 128;                  // UMMA M (positions per tile)
constexpr int TC_EPI_WARPS = 8;             // two per TMEM lane quadrant, each takes half of the 16-column chunks
constexpr int TC_EPI_THREADS = TC_EPI_WARPS * 32;
constexpr int TC_PROD_WARP0 = TC_EPI_WARPS;   // 4 producer warps
constexpr int TC_MMA_WARP = TC_EPI_WARPS + 4;
constexpr int TC_THREADS = (TC_MMA_WARP + 1) * 32;
// TC_KPS (template parameter KPS of the kernel): UMMA K-steps per ring stage — one mbarrier round trip feeds KPS * k
// MMAs.  2 when the ring can still hold >= 5 stages (small weight tiles), else 1.
constexpr int TC_MAX_SLOTS = 6;            // documents packed into one tile (short-document mode)

struct TcPlan {
    int E, H, K, L, pad, Lout, Lext;
    int C;            // 16-byte K chunks per row = Epad16 / 8
    int ksteps;       // UMMA K-steps (16 elements) = Epad16 / 16
    int P, Nb;        // filter passes, filters per pass (multiple of 16)
    int RS;           // staged rows per chunk column (>= 128 + K - 1, == 4 mod 8 → conflict-free cp.async stores)
    int rows;         // 128 + K - 1
    int kps;          // K-steps per ring stage (1 or 2)
    int stage_bytes;  // 2 * kps * RS * 16
    int nstages_k;    // ring stages consumed per tile = ceil(ksteps / TC_KPS)
    int nst;          // ring stages
    int w_bytes;      // K * C * Nb * 16
    int mode_b;       // 1 = several short docs per tile
    int D;            // docs per tile (mode B) or 1
    int tpu;          // tiles per unit (mode A: ceil(Lout/128); mode B: 1)
    int64_t n_units;
    int ib;           // index bits in the packed key
    int tmem_cols, acc_stride;
    int off_w, off_ring, off_bias, off_keys, off_bars, off_slot, smem_bytes;
    int act;
};

// rows per chunk column: >= 128 + K - 1, and == 4 (mod 8) for 2 chunk columns per stage / == 2 (mod 8) for 4, so that
// the chunk columns x rows written by a quarter-warp of cp.async lanes land in 8 distinct 16-byte bank groups
__host__ __device__ inline int tc_rs(int K, int kps) {
    const int need = TC_M + K - 1;
    const int rem = kps == 1 ? 4 : 2;
    int s = (need / 8) * 8 + rem;
    return s >= need ? s : s + 8;
}

// filters-per-pass decision shared with rbr_conv_pack (the packed B operand is laid out per pass)
void tc_pass_split(int64_t E, int64_t H, int64_t K, int64_t* P, int64_t* Nb) {
    const int64_t epad16 = round_up(E, 16);
    const int64_t bytes_per_n = K * epad16 * 2;
    const int64_t ring_min = 4 * 2 * tc_rs((int)K, 1) * 16;
    const int64_t budget = TC_SMEM_MAX - ring_min - 6144;
    int64_t nb_max = budget / bytes_per_n / 16 * 16;
    if (nb_max > 256) nb_max = 256;
    if (nb_max < 16) { *P = 0; *Nb = 0; return; }          // tensor-core variant unavailable for this shape
    const int64_t npad = round_up(H, 16);
    *P = (npad + nb_max - 1) / nb_max;
    *Nb = round_up((H + *P - 1) / *P, 16);
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
struct TcArgs {
    const __nv_bfloat16* shadow;
    int64_t vocab;
    IdView ids;
    const uint8_t* mask;
    int64_t n_docs;
    const __nv_bfloat16* wpack;      // [P][K][C][Nb][8]
    const void* zero_row;            // >= emb_pad*2 zero bytes (global): source of rows that must read as zeros
    const float* bias;
    float* feat;
    int32_t* argmax;
    float* preact;                   // optional pool_raw: pooled gate * conv_nobias(x), before bias and activation
    const float* gate;               // optional multiplicative gate (D-ATT): mode 1 per token (k == 1), mode 2 per doc
    int gate_mode;
    int feat_ld;
    int emb_pad;                     // shadow row pitch in elements
    TcPlan p;
};

// row r of the staged tile → (document, input position) or "zero row"
__device__ __forceinline__ bool tc_row_source(const TcPlan& p, int64_t unit, int tt, int r, int64_t n_docs, int64_t* doc, int* t_in) {
    if (r >= p.rows) return false;
    int64_t d;
    int ext;
    if (p.mode_b) {
        const int q = r / p.Lext;
        if (q >= p.D) return false;
        d = unit * p.D + q;
        ext = r - q * p.Lext;
    } else {
        d = unit;
        ext = tt * TC_M + r;
        if (ext >= p.Lext) return false;
    }
    if (d >= n_docs) return false;
    const int t = ext - p.pad;
    if (t < 0 || t >= p.L) return false;
    *doc = d;
    *t_in = t;
    return true;
}

template <int KT, int TC_KPS>
__global__ void __launch_bounds__(TC_THREADS, 1) conv_tc_kernel(const TcArgs a) {
    constexpr int TC_CPS = 2 * TC_KPS;          // 16-byte chunk columns per stage
    extern __shared__ __align__(128) uint8_t smem[];
    const TcPlan& p = a.p;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t w_s = sbase + p.off_w, ring_s = sbase + p.off_ring;
    float* bias_s = reinterpret_cast<float*>(smem + p.off_bias);
    unsigned long long* keys_s = reinterpret_cast<unsigned long long*>(smem + p.off_keys);
    const uint32_t bars = sbase + p.off_bars;
    // barrier slots: full[nst], empty[nst], acc_full[2], acc_empty[2], w_ready
    const uint32_t bar_full = bars, bar_empty = bars + 8 * p.nst, bar_accf = bars + 16 * p.nst, bar_acce = bar_accf + 16,
                   bar_w = bar_acce + 16;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + p.off_slot);

    const int pass = blockIdx.x % p.P;
    const int cta_in_pass = blockIdx.x / p.P, ctas_per_pass = gridDim.x / p.P;
    const int h0 = pass * p.Nb;
    // units this CTA owns: cta_in_pass, cta_in_pass + ctas_per_pass, ...
    const int64_t my_units = (p.n_units > cta_in_pass) ? (p.n_units - cta_in_pass + ctas_per_pass - 1) / ctas_per_pass : 0;
    const int64_t my_tiles = my_units * p.tpu;

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.nst; ++i) { mbar_init(bar_full + 8 * i, 128); mbar_init(bar_empty + 8 * i, 1); }
        mbar_init(bar_accf, 1); mbar_init(bar_accf + 8, 1);
        mbar_init(bar_acce, TC_EPI_THREADS); mbar_init(bar_acce + 8, TC_EPI_THREADS);
        mbar_init(bar_w, 1);
        fence_barrier_init();
    }
    for (int i = threadIdx.x; i < p.Nb; i += blockDim.x) bias_s[i] = (h0 + i < p.H) ? a.bias[h0 + i] : 0.f;
    for (int i = threadIdx.x; i < p.D * p.Nb; i += blockDim.x) keys_s[i] = 0ull;
    if (warp == TC_MMA_WARP) tmem_alloc(smem_u32((const void*)tmem_slot), (uint32_t)p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == TC_MMA_WARP) {
        // =========================== MMA issuer ===========================
        // The whole warp runs the (warp-uniform) control flow so every operand lives in uniform registers; one elected
        // lane issues tcgen05.mma / tcgen05.commit.  Descriptors are advanced by integer adds on their low word.
        if (my_tiles > 0) {
            const bool leader = elect_one();
            if (leader) {
                // resident B operand: one TMA bulk stream, <= 32 KB per copy
                mbar_expect_tx(bar_w, (uint32_t)p.w_bytes);
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(a.wpack) + (size_t)pass * p.w_bytes;
                for (int off = 0; off < p.w_bytes; off += 32768) {
                    const int n = min(32768, p.w_bytes - off);
                    bulk_g2s(w_s + off, wsrc + off, (uint32_t)n, bar_w);
                }
            }
            __syncwarp();
            mbar_wait(bar_w, 0);
            const uint32_t idesc = umma_idesc(TC_M, p.Nb);
            const uint64_t a_desc0 = umma_desc(ring_s, (uint32_t)p.RS * 16u, 128u);      // stage 0, tap 0
            const uint64_t b_desc0 = umma_desc(w_s, (uint32_t)p.Nb * 16u, 128u);         // tap 0, K-step 0
            const uint32_t a_stage_inc = (uint32_t)p.stage_bytes >> 4;                    // descriptor address units (16 B)
            const uint32_t b_tap_inc = (uint32_t)(p.C * p.Nb);                            // one tap = C chunk columns of Nb rows
            const uint32_t b_step_inc = (uint32_t)(2 * p.Nb);                             // one K-step = 2 chunk columns
            int stage = 0;
            uint32_t ph = 0;
            uint32_t full_bar = bar_full, empty_bar = bar_empty;
            uint64_t ad = a_desc0;
            const int last = p.nstages_k - 1;
            const uint32_t a_kstep_inc = (uint32_t)(2 * p.RS);                       // one K-step = 2 chunk columns of RS rows
            for (int64_t g = 0; g < my_tiles; ++g) {
                const int buf = (int)(g & 1);
                mbar_wait(bar_acce + 8 * buf, (uint32_t)(((g >> 1) & 1) ^ 1));      // epilogue drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * p.acc_stride);
                uint64_t bd = b_desc0;
                for (int s = 0; s <= last; ++s) {
                    // No proxy fence here: the ring is written by cp.async whose completion is what flips this mbarrier
                    // (cp.async.mbarrier.arrive), the same producer/consumer protocol CUTLASS's sm100 cp.async mainloop uses.
                    mbar_wait(full_bar, ph);
                    tc_fence_after();
                    if (leader) {
#pragma unroll
                        for (int ks = 0; ks < TC_KPS; ++ks) {
                            if (s * TC_KPS + ks < p.ksteps) {
#pragma unroll
                                for (int j = 0; j < KT; ++j)
                                    umma_bf16(d_tmem, ad + (uint64_t)(ks * a_kstep_inc + j),
                                              bd + (uint64_t)(ks * b_step_inc + j * b_tap_inc), idesc, (uint32_t)((s | ks | j) != 0));
                            }
                        }
                        umma_commit(empty_bar);                                      // frees the ring slot when the MMAs retire
                        if (s == last) umma_commit(bar_accf + 8 * buf);              // accumulator complete → epilogue
                    }
                    __syncwarp();
                    bd += (uint64_t)(TC_KPS * b_step_inc);
                    ad += a_stage_inc;
                    full_bar += 8;
                    empty_bar += 8;
                    if (++stage == p.nst) { stage = 0; ph ^= 1; ad = a_desc0; full_bar = bar_full; empty_bar = bar_empty; }
                }
            }
        }
        __syncwarp();
    } else if (warp >= TC_PROD_WARP0) {
        // =========================== gather producers (128 threads) ===========================
        // Each thread owns up to 3 fixed 16-byte pieces (row, half) of every stage.  The loop is issue-latency bound (one warp
        // per scheduler), so it is kept to a handful of instructions per stage: row pointers are resolved once per tile,
        // rows that must read as zeros (conv padding, masked tokens, bad ids) point at a zero row so every copy is an
        // unconditional 16-byte cp.async, and four stages are unrolled so the K offset is an immediate.
        const int ptid = threadIdx.x - TC_PROD_WARP0 * 32;
        // piece → (row = piece / TC_CPS, chunk column c = piece % TC_CPS): TC_CPS adjacent lanes fetch 16*TC_CPS contiguous
        // bytes of one token row.  Pieces ptid + 128*i, i < PP; a thread's pieces all share c and rows step by 128/TC_CPS.
        constexpr int PP = (TC_CPS * (TC_M + 8) + 127) / 128;          // covers rows up to 128 + 7 (k <= 7... see plan check)
        constexpr int ROWSTEP = 128 / TC_CPS;
        const int my_c = ptid % TC_CPS;
        const int my_r0 = ptid / TC_CPS;
        const uint32_t d0 = (uint32_t)(my_c * p.RS * 16 + my_r0 * 16);
        int npieces = 0;
#pragma unroll
        for (int i = 0; i < PP; ++i)
            if (my_r0 + ROWSTEP * i < p.rows) npieces = i + 1;
        // token ids / mask bytes of tile g+1 are fetched while tile g streams, so the id → row-address dependency
        // (two dependent global loads) never stalls the ring.
        int64_t id_next[PP];
        uint8_t ok_next[PP];
        auto prefetch_ids = [&](int64_t g) {
            const int64_t unit = cta_in_pass + (g / p.tpu) * ctas_per_pass;
            const int tt = (int)(g % p.tpu);
#pragma unroll
            for (int i = 0; i < PP; ++i) {
                id_next[i] = -1;
                ok_next[i] = 0;
                if (i >= npieces || g >= my_tiles) continue;
                int64_t doc;
                int t;
                if (tc_row_source(p, unit, tt, my_r0 + ROWSTEP * i, a.n_docs, &doc, &t)) {
                    id_next[i] = ld_id(a.ids, doc * p.L + t);
                    ok_next[i] = ld_mask(a.ids, a.mask, doc * p.L + t, id_next[i]) ? (uint8_t)1 : (uint8_t)0;
                }
            }
        };
        int stage = 0;
        uint32_t ph = 0;
        uint32_t sb = ring_s + d0;
        prefetch_ids(0);
        for (int64_t g = 0; g < my_tiles; ++g) {
            // rows that must read as zeros (conv padding, masked tokens, bad ids) keep src_bytes = 0: cp.async then
            // zero-fills the 16 bytes without issuing a memory request (a shared "zero row" would be an L2 hot spot)
            const char* src[PP];
            uint32_t nb[PP];
#pragma unroll
            for (int i = 0; i < PP; ++i) {
                src[i] = reinterpret_cast<const char*>(a.shadow);
                nb[i] = 0;
                if (ok_next[i]) {
                    const int64_t id = id_next[i];
                    if (id >= 0 && id < a.vocab) { src[i] = reinterpret_cast<const char*>(a.shadow + id * a.emb_pad) + my_c * 16; nb[i] = 16; }
                    else if (my_c == 0) note_oob();
                }
            }
            prefetch_ids(g + 1);
            // the last stage of an odd K-step count copies one K-step of row padding (zeros inside the 128-byte aligned
            // shadow row) that no MMA reads
            for (int s = 0; s < p.nstages_k; ++s) {
                mbar_wait(bar_empty + 8 * stage, ph ^ 1);
#pragma unroll
                for (int i = 0; i < PP; ++i)
                    if (i < npieces) cp_async16_imm<0>(sb + (uint32_t)(i * ROWSTEP * 16), src[i], nb[i]);
                cp_async_arrive_noinc(bar_full + 8 * stage);
#pragma unroll
                for (int i = 0; i < PP; ++i) src[i] += 16 * TC_CPS;
                sb += (uint32_t)p.stage_bytes;
                if (++stage == p.nst) { stage = 0; ph ^= 1; sb = ring_s + d0; }
            }
        }
        cp_async_wait_all();
    } else {
        // =========================== epilogue (warps 0-3; warp w owns TMEM lanes 32w..32w+31) ===========================
        const int quad = warp & 3;                            // TMEM lane quadrant this warp may read
        const int half = warp >> 2;                           // which share of the 16-column chunks it reduces
        const int m = quad * 32 + lane;                       // tile row = TMEM lane
        const int n_chunks = p.Nb / 16;
        const int chunk_lo = half ? (n_chunks + 1) / 2 : 0;
        const int chunk_hi = half ? n_chunks : (n_chunks + 1) / 2;
        for (int64_t g = 0; g < my_tiles; ++g) {
            const int64_t unit = cta_in_pass + (g / p.tpu) * ctas_per_pass;
            const int tt = (int)(g % p.tpu);
            const int buf = (int)(g & 1);
            // this row's (slot, position)
            int slot, t;
            if (p.mode_b) { slot = m / p.Lext; t = m - slot * p.Lext; } else { slot = 0; t = tt * TC_M + m; }
            const bool valid = (t < p.Lout) && (slot < p.D) && ((p.mode_b ? unit * p.D + slot : unit) < a.n_docs);
            float row_gate = 1.f;
            if (a.gate_mode == 1 && valid) row_gate = a.gate[(p.mode_b ? unit * p.D + slot : unit) * p.L + t];
            // slots this warp's 32 rows touch
            const int slot_lo = p.mode_b ? (quad * 32) / p.Lext : 0;
            int slot_hi = p.mode_b ? (quad * 32 + 31) / p.Lext : 0;
            if (slot_hi >= p.D) slot_hi = p.D - 1;

            mbar_wait(bar_accf + 8 * buf, (uint32_t)((g >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * p.acc_stride);
            for (int ch = chunk_lo; ch < chunk_hi; ++ch) {
                const int c0 = ch * 16;
                uint32_t v[16];
                tmem_ld16(taddr + (uint32_t)c0, v);
                tmem_ld_wait();
                if (ch == chunk_hi - 1) {                     // this warp's last chunk is in registers: release the accumulator
                    tc_fence_before();
                    mbar_arrive(bar_acce + 8 * buf);
                }
                if (a.gate_mode == 1) {                       // per-token gate (k == 1): conv(g_t * x_t) = g_t * conv(x_t)
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * row_gate);
                }
                for (int sl = slot_lo; sl <= slot_hi; ++sl) {
                    const bool mine = valid && slot == sl;
                    uint32_t keep_v = 0, keep_b = 0;
                    if (__all_sync(0xffffffffu, mine)) tc_colmax<true>(v, true, lane, keep_v, keep_b);
                    else tc_colmax<false>(v, mine, lane, keep_v, keep_b);
                    if (lane < 16 && keep_b) {
                        // first (smallest-position) row attaining the max; rows of one slot are consecutive lanes
                        const int first = __ffs(keep_b) - 1;
                        const int tf = p.mode_b ? (quad * 32 + first - sl * p.Lext) : (tt * TC_M + quad * 32 + first);
                        const unsigned long long key =
                            ((unsigned long long)f2ord(keep_v) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)tf);
                        atomicMax(keys_s + sl * p.Nb + c0 + lane, key);
                    }
                }
            }
            if (chunk_lo == chunk_hi) { tc_fence_before(); mbar_arrive(bar_acce + 8 * buf); }   // Nb == 16: upper half idle
            if (tt == p.tpu - 1) {
                // ---- unit finished: merge is complete once all 4 epilogue warps have posted their keys
                asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
                for (int o = threadIdx.x; o < p.D * p.Nb; o += TC_EPI_THREADS) {
                    const int sl = o / p.Nb, c = o - sl * p.Nb;
                    const int64_t doc = p.mode_b ? unit * p.D + sl : unit;
                    const unsigned long long key = keys_s[o];
                    keys_s[o] = 0ull;
                    if (doc < a.n_docs && h0 + c < p.H) {
                        const float raw = __uint_as_float(ord2f((uint32_t)(key >> 32)));
                        const int tbest = (int)(0xFFFFFFFFu - (uint32_t)key);
                        const float gated = a.gate_mode == 2 ? raw * a.gate[doc] : raw;                 // per-doc gate > 0: monotone
                        a.feat[doc * a.feat_ld + h0 + c] = act_apply(p.act, gated + bias_s[c]);
                        if (a.preact) a.preact[doc * a.feat_ld + h0 + c] = gated;                       // pool_raw: no bias
                        a.argmax[doc * a.feat_ld + h0 + c] = tbest;
                    }
                }
                asm volatile("bar.sync 1, %0;" ::"n"(TC_EPI_THREADS) : "memory");
            }
        }
    }
    // teardown
    tc_fence_before();
    __syncthreads();
    if (warp == TC_MMA_WARP) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

static bool tc_make_plan(int E, int H, int K, int L, int pad, int act, int64_t n_docs, TcPlan* out) {
    TcPlan p{};
    p.E = E; p.H = H; p.K = K; p.L = L; p.pad = pad; p.act = act;
    p.Lout = L + 2 * pad - K + 1;
    p.Lext = L + 2 * pad;
    if (p.Lout < 1) return false;
    const int epad16 = (int)round_up(E, 16);
    p.C = epad16 / 8;
    p.ksteps = epad16 / 16;
    int64_t P, Nb;
    tc_pass_split(E, H, K, &P, &Nb);
    if (P == 0) return false;
    p.P = (int)P; p.Nb = (int)Nb;
    p.rows = TC_M + K - 1;
    p.w_bytes = K * p.C * p.Nb * 16;
    if (p.Lext <= TC_M + K - 1 && p.Lext * 2 <= TC_M + K - 1) {
        p.mode_b = 1;
        p.D = (TC_M + K - 1) / p.Lext;
        int dmax = 4096 / (p.Nb * 8);
        if (dmax > TC_MAX_SLOTS) dmax = TC_MAX_SLOTS;
        if (dmax < 1) dmax = 1;
        if (p.D > dmax) p.D = dmax;
        p.tpu = 1;
        p.n_units = (n_docs + p.D - 1) / p.D;
    } else {
        p.mode_b = 0;
        p.D = 1;
        p.tpu = (p.Lout + TC_M - 1) / TC_M;
        p.n_units = n_docs;
    }
    p.ib = 0;                                    // (unused since the pooled key carries value and position separately)
    p.acc_stride = p.Nb <= 128 ? 128 : 256;
    p.tmem_cols = 2 * p.acc_stride;
    int off = 0;
    p.off_w = off; off += p.w_bytes;
    off = (off + 127) / 128 * 128;
    p.off_ring = off;
    const int fixed_tail = p.Nb * 4 + p.D * p.Nb * 8 + 8 * (2 * 8 + 5) + 16 + 256;
    // two K-steps per stage when at least 5 such stages fit, else one
    int nst = 0;
    for (int kps = 2; kps >= 1; --kps) {
        p.kps = kps;
        p.RS = tc_rs(K, kps);
        p.stage_bytes = 2 * kps * p.RS * 16;
        nst = (TC_SMEM_MAX - off - fixed_tail) / p.stage_bytes;
        if (nst > 8) nst = 8;
        if (nst >= 5) break;
    }
    if (nst < 3) return false;
    p.nstages_k = (p.ksteps + p.kps - 1) / p.kps;
    p.nst = nst;
    off += nst * p.stage_bytes;
    p.off_bias = off; off += p.Nb * 4;
    off = (off + 7) / 8 * 8;
    p.off_keys = off; off += p.D * p.Nb * 8;
    p.off_bars = off; off += 8 * (2 * nst + 5);
    p.off_slot = off; off += 16;
    p.smem_bytes = off;
    if (p.smem_bytes > TC_SMEM_MAX) return false;
    *out = p;
    return true;
}

int conv_tc_dispatch(const __nv_bfloat16* shadow, int64_t vocab, int E, IdView ids, const uint8_t* mask,
                     const float* gate, int gate_mode, int64_t n_docs, int L, const __nv_bfloat16* umma_w, const void* zero_row,
                     const float* bias, int H, int K, int pad, int act, float* feat, int32_t* argmax, float* preact, int feat_ld,
                     cudaStream_t s) {
    TcArgs a{};
    RBR_REQUIRE(tc_make_plan(E, H, K, L, pad, act, n_docs, &a.p), RBR_EUNSUPPORTED,
                "conv_fwd[bf16]: shape (E=%d H=%d k=%d L=%d) outside the tensor-core variant; use precision fp32", E, H, K, L);
    a.shadow = shadow; a.vocab = vocab; a.ids = ids; a.mask = mask; a.n_docs = n_docs; a.wpack = umma_w; a.zero_row = zero_row; a.bias = bias;
    a.feat = feat; a.argmax = argmax; a.preact = preact; a.gate = gate; a.gate_mode = gate_mode; a.feat_ld = feat_ld; a.emb_pad = (int)rbr_emb_pad(E);
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        RBR_CUDA(cudaGetDevice(&dev));
        RBR_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    }
    // persistent grid: one CTA per SM (a multiple of the pass count), never more CTAs than units
    int64_t grid = num_sms / a.p.P * a.p.P;
    if (grid < a.p.P) grid = a.p.P;
    const int64_t max_useful = a.p.n_units * a.p.P;
    if (grid > max_useful) grid = max_useful;
#define RBR_TC(KT_)                                                                                               \
    case KT_: {                                                                                                   \
        static bool attr = false;                                                                                 \
        if (!attr) {                                                                                              \
            RBR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KT_, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX)); \
            RBR_CUDA(cudaFuncSetAttribute(conv_tc_kernel<KT_, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_MAX)); \
            attr = true;                                                                                          \
        }                                                                                                         \
        if (a.p.kps == 2) conv_tc_kernel<KT_, 2><<<(unsigned)grid, TC_THREADS, a.p.smem_bytes, s>>>(a);           \
        else conv_tc_kernel<KT_, 1><<<(unsigned)grid, TC_THREADS, a.p.smem_bytes, s>>>(a);                        \
    } break;
    switch (K) {
        RBR_TC(1) RBR_TC(2) RBR_TC(3) RBR_TC(4) RBR_TC(5) RBR_TC(7)
        default:
            set_error("conv_fwd[bf16]: kernel size %d not supported (1,2,3,4,5,7)", K);
            return RBR_EUNSUPPORTED;
    }
#undef RBR_TC
    RBR_LAUNCH_CHECK("conv_tc_kernel");
    return RBR_OK;
}

}  // namespace rbr

RBR_DEFINE_OOB_ACCESSOR(conv_tc)
