// token_sort.cuh — counting sort of int32 keys in [0, n_bins) (key < 0 = "skip", placed last).
// Produces `order`: entry indices grouped by key, so that a warp walking a slice of `order` sees
// equal keys adjacent (the "segments" of the warp-segmented scatter-add).  Three tiny launches:
// histogram (int atomics) → single-CTA exclusive scan → fill (atomic cursor per bin).
#pragma once
#include "rbr_common.cuh"

namespace rbr {

struct TokenSort {
    int32_t* order;   // [n]     entry indices, grouped by key; skipped entries at the end
    int32_t* start;   // [bins+2] start[b] = first slot of bin b; start[bins] = first skipped slot
};

inline int64_t token_sort_workspace_bytes(int64_t n, int64_t bins) {
    return round_up((bins + 2) * 4, 256) * 2 + round_up(n * 4, 256);
}

static __global__ void __launch_bounds__(256) tsort_hist_kernel(const int32_t* __restrict__ keys, int64_t n, int bins,
                                                         int32_t* __restrict__ counts) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int k = keys[i];
        atomicAdd(counts + (k < 0 ? bins : k), 1);
    }
}

// single CTA, 1024 threads: exclusive scan of counts[0..bins] → start[], cursor[] (copy)
static __global__ void __launch_bounds__(1024) tsort_scan_kernel(const int32_t* counts, int bins_p1, int32_t* __restrict__ start,
                                                          int32_t* cursor /* may alias counts */) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < bins_p1; base += 1024) {
        const int i = base + tid;
        const int v = (i < bins_p1) ? counts[i] : 0;
        int x = v;
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
            warp_tot[lane] = t;                       // inclusive scan of warp totals
        }
        __syncthreads();
        const int carry = carry_s;
        const int excl = carry + (wid ? warp_tot[wid - 1] : 0) + x - v;
        if (i < bins_p1) { start[i] = excl; cursor[i] = excl; }
        __syncthreads();
        if (tid == 1023) carry_s = carry + warp_tot[31];
        __syncthreads();
    }
    if (tid == 0) start[bins_p1] = carry_s;
}

static __global__ void __launch_bounds__(256) tsort_fill_kernel(const int32_t* __restrict__ keys, int64_t n, int bins,
                                                         int32_t* __restrict__ cursor, int32_t* __restrict__ order) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int k = keys[i];
        const int pos = atomicAdd(cursor + (k < 0 ? bins : k), 1);
        order[pos] = (int32_t)i;
    }
}

inline int token_sort(const int32_t* keys, int64_t n, int64_t bins, void* ws, TokenSort& out, cudaStream_t s) {
    char* p = reinterpret_cast<char*>(ws);
    int32_t* counts = reinterpret_cast<int32_t*>(p);          // reused as cursor after the scan
    out.start = reinterpret_cast<int32_t*>(p + round_up((bins + 2) * 4, 256));
    out.order = reinterpret_cast<int32_t*>(p + 2 * round_up((bins + 2) * 4, 256));
    RBR_CUDA(cudaMemsetAsync(counts, 0, (bins + 2) * 4, s));
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    tsort_hist_kernel<<<blocks, 256, 0, s>>>(keys, n, (int)bins, counts);
    RBR_LAUNCH_CHECK("tsort_hist");
    // scan reads counts and writes start + cursor; cursor aliases counts, which is safe because every
    // element is read (into a register) before its own slot is written by the same thread.
    tsort_scan_kernel<<<1, 1024, 0, s>>>(counts, (int)bins + 1, out.start, counts);
    RBR_LAUNCH_CHECK("tsort_scan");
    tsort_fill_kernel<<<blocks, 256, 0, s>>>(keys, n, (int)bins, counts, out.order);
    RBR_LAUNCH_CHECK("tsort_fill");
    return RBR_OK;
}

}  // namespace rbr
