// ub_redux.cu — issue cost of the warp-wide primitives the conv epilogue's column max is built from (per SM sub-partition,
// 4 warps each issuing independent operations): redux.sync.max, vote.ballot, shfl.bfly, and plain IMNMX for scale.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a ub_redux.cu -o ub_redux
#include <cuda_runtime.h>
#include <cstdio>
template <int OP>
__global__ void __launch_bounds__(512) k(int* out, int iters, long long* cyc) {
    int v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 31 + i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (OP == 0) v[i] = __reduce_max_sync(0xffffffffu, v[i] + it);
            if (OP == 1) v[i] = (int)__ballot_sync(0xffffffffu, v[i] + it > 7) + v[i];
            if (OP == 2) v[i] = max(v[i], __shfl_xor_sync(0xffffffffu, v[i] + it, 16));
            if (OP == 3) v[i] = max(v[i] + it, v[(i + 1) & 15]);
        }
    }
    const long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    int* out; long long* cyc; cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    const char* names[4] = {"redux.sync.max.s32", "vote.ballot", "shfl.bfly + max", "IMNMX (alu)"};
    for (int op = 0; op < 4; ++op) {
        for (int rep = 0; rep < 2; ++rep) {
            if (op == 0) k<0><<<148, 512>>>(out, iters, cyc);
            if (op == 1) k<1><<<148, 512>>>(out, iters, cyc);
            if (op == 2) k<2><<<148, 512>>>(out, iters, cyc);
            if (op == 3) k<3><<<148, 512>>>(out, iters, cyc);
            cudaDeviceSynchronize();
        }
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        // 16 warps per SM = 4 per sub-partition, each issued iters*16 ops
        printf("%-22s %.2f clk per warp-op per sub-partition (4 warps/SMSP, 16 independent ops in flight per warp)\n", names[op],
               (double)c / ((double)iters * 16 * 4));
    }
    return 0;
}
