// token_sort.cuh — counting sort of int32 keys in [0, n_bins); entries with key < 0 are dropped ("skip").
// Produces `order`: entry indices grouped by key, so that a warp walking a slice of `order` sees
// equal keys adjacent (the "segments" of the warp-segmented scatter-add).  Three launches:
// histogram → single-CTA exclusive scan → fill.
//
// Token ids are frequency-ranked by the reference's preprocessing (preprocess/_tokenizer.py:53-65: ids are
// assigned in descending corpus frequency), so the hot keys are the SMALL ids: a few ids receive several
// percent of all entries.  Global atomics on those addresses serialise (one L2 atomic unit per address), so
// both the histogram and the fill keep keys < TS_SMEM_BINS in CTA-private shared-memory counters and touch
// global memory once per (CTA, non-empty bin); cold keys go straight to global atomics.
#pragma once
#include "rbr_common.cuh"

namespace rbr {

struct TokenSort {
    int32_t* order;   // [n]      entry indices grouped by key; only the first start[bins] slots are written
    int32_t* start;   // [bins+2] start[b] = first slot of bin b; start[bins] = number of kept (key >= 0) entries
};

constexpr int TS_SMEM_BINS = 4096;     // hot bins privatised per CTA (16 KB of counters + 16 KB of bases)
constexpr int TS_THREADS = 512;

inline int64_t token_sort_workspace_bytes(int64_t n, int64_t bins) {
    return round_up((bins + 2) * 4, 256) * 2 + round_up(n * 4, 256);
}


static __global__ void __launch_bounds__(TS_THREADS) tsort_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int bins,
                                                                       int32_t* __restrict__ counts) {
    __shared__ int32_t hot[TS_SMEM_BINS];
    for (int i = threadIdx.x; i < TS_SMEM_BINS; i += TS_THREADS) hot[i] = 0;
    __syncthreads();
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;      // contiguous slice per CTA
    const int64_t lo = (int64_t)blockIdx.x * per, hi = min(lo + per, n);
    for (int64_t i = lo + threadIdx.x; i < hi; i += TS_THREADS) {
        const int b = keys[i];
        if (b < 0) continue;                       // skipped entry (padding / masked / zero coefficient): not sorted at all
        if (b < TS_SMEM_BINS) atomicAdd(hot + b, 1);
        else atomicAdd(counts + b, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TS_SMEM_BINS && i < bins; i += TS_THREADS) {
        const int c = hot[i];
        if (c) atomicAdd(counts + i, c);
    }
}

// single CTA, 1024 threads: exclusive scan of counts[0..bins] → start[], cursor[] (copy).
// The counts are staged through shared memory in super-chunks of 1024 x TS_RUN elements (coalesced global reads and
// writes); each thread scans its own contiguous run of TS_RUN elements in shared memory (odd run length → conflict
// free), one block-level scan of the 1024 run totals per super-chunk.
constexpr int TS_RUN = 47;
constexpr int TS_SUPER = 1024 * TS_RUN;          // 48128 elements = 188 KB of shared memory
static __global__ void __launch_bounds__(1024) tsort_scan_kernel(const int32_t* counts, int bins_p1, int32_t* __restrict__ start,
                                                                 int32_t* cursor /* may alias counts */) {
    extern __shared__ int32_t buf[];             // [TS_SUPER]
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < bins_p1; base += TS_SUPER) {
        const int cnt = min(TS_SUPER, bins_p1 - base);
        for (int i = tid; i < cnt; i += 1024) buf[i] = counts[base + i];
        __syncthreads();
        const int lo = min(tid * TS_RUN, cnt), hi = min(lo + TS_RUN, cnt);
        int sum = 0;
        for (int i = lo; i < hi; ++i) { const int c = buf[i]; buf[i] = sum; sum += c; }     // run-local exclusive scan
        int x = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int t = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            warp_tot[lane] = t;                   // inclusive scan of warp totals
        }
        __syncthreads();
        const int carry = carry_s;
        const int off = carry + (wid ? warp_tot[wid - 1] : 0) + x - sum;      // exclusive prefix of this thread's run
        for (int i = lo; i < hi; ++i) buf[i] += off;
        __syncthreads();
        for (int i = tid; i < cnt; i += 1024) { const int v = buf[i]; start[base + i] = v; cursor[base + i] = v; }
        if (tid == 0) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) start[bins_p1] = carry_s;
}

static __global__ void __launch_bounds__(TS_THREADS) tsort_fill_kernel(const int32_t* __restrict__ keys, int64_t n, int bins,
                                                                       int32_t* __restrict__ cursor, int32_t* __restrict__ order) {
    __shared__ int32_t hot[TS_SMEM_BINS];      // pass 1: CTA-local count; pass 2: running rank inside the CTA's range
    __shared__ int32_t base[TS_SMEM_BINS];     // first global slot reserved for this CTA in each hot bin
    for (int i = threadIdx.x; i < TS_SMEM_BINS; i += TS_THREADS) hot[i] = 0;
    __syncthreads();
    const int64_t per = (n + gridDim.x - 1) / gridDim.x;
    const int64_t lo = (int64_t)blockIdx.x * per, hi = min(lo + per, n);
    for (int64_t i = lo + threadIdx.x; i < hi; i += TS_THREADS) {
        const int b = keys[i];
        if (b >= 0 && b < TS_SMEM_BINS) atomicAdd(hot + b, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < TS_SMEM_BINS && i < bins; i += TS_THREADS) {
        const int c = hot[i];
        base[i] = c ? atomicAdd(cursor + i, c) : 0;           // one global atomic per (CTA, non-empty hot bin)
        hot[i] = 0;
    }
    __syncthreads();
    for (int64_t i = lo + threadIdx.x; i < hi; i += TS_THREADS) {
        const int b = keys[i];
        if (b < 0) continue;
        const int pos = (b < TS_SMEM_BINS) ? base[b] + atomicAdd(hot + b, 1) : atomicAdd(cursor + b, 1);
        order[pos] = (int32_t)i;
    }
}

static inline int token_sort(const int32_t* keys, int64_t n, int64_t bins, void* ws, TokenSort& out, cudaStream_t s) {
    char* p = reinterpret_cast<char*>(ws);
    int32_t* counts = reinterpret_cast<int32_t*>(p);          // reused as cursor after the scan
    out.start = reinterpret_cast<int32_t*>(p + round_up((bins + 2) * 4, 256));
    out.order = reinterpret_cast<int32_t*>(p + 2 * round_up((bins + 2) * 4, 256));
    RBR_CUDA(cudaMemsetAsync(counts, 0, (bins + 2) * 4, s));
    // 2 CTAs per SM: enough warps to hide latency, few enough that each CTA's flush of the hot bins amortises
    int blocks = (int)((n + 4095) / 4096);
    if (blocks > 148 * 2) blocks = 148 * 2;
    if (blocks < 1) blocks = 1;
    tsort_hist_kernel<<<blocks, TS_THREADS, 0, s>>>(keys, n, (int)bins, counts);
    RBR_LAUNCH_CHECK("tsort_hist");
    // scan reads counts and writes start + cursor; cursor aliases counts, which is safe because every
    // element is read (into a register) before its own slot is written by the same thread.
    static bool scan_attr = false;
    if (!scan_attr) {
        RBR_CUDA(cudaFuncSetAttribute(tsort_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TS_SUPER * 4));
        scan_attr = true;
    }
    tsort_scan_kernel<<<1, 1024, TS_SUPER * 4, s>>>(counts, (int)bins + 1, out.start, counts);
    RBR_LAUNCH_CHECK("tsort_scan");
    tsort_fill_kernel<<<blocks, TS_THREADS, 0, s>>>(keys, n, (int)bins, counts, out.order);
    RBR_LAUNCH_CHECK("tsort_fill");
    return RBR_OK;
}

}  // namespace rbr
