// tma_util.cuh — host-side tensor-map construction (cuTensorMapEncodeTiled through the runtime's driver entry point: no
// -lcuda) and the device-side TMA tile load, shared by the tcgen05 kernels.  sm_100a only.
#pragma once
#include <cuda.h>   // CUtensorMap (types only)

#include "rbr_common.cuh"

namespace rbr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(f);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

// 2-D bf16 tensor [rows][cols] with a row pitch of `pitch_elems` elements; boxes of box_cols x box_rows, SWIZZLE_128B
// (box_cols * 2 bytes must be 128), out-of-range elements read as zeros.
inline bool make_tmap_bf16_2d(CUtensorMap* tm, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems,
                              uint32_t box_cols, uint32_t box_rows) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return false;
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)pitch_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// one box at (col, row) → shared memory at `dst`; bytes complete on the CTA-local mbarrier `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int col, int row, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(tmap), "r"(col), "r"(row), "r"(bar)
        : "memory");
}

}  // namespace rbr
