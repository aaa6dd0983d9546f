// conv_tc.cu — placeholder until the tcgen05 kernel lands (next commit).
#include "rbr_common.cuh"
namespace rbr {
int conv_tc_dispatch(const __nv_bfloat16*, int64_t, int, const int64_t*, const uint8_t*, const float*, int, int64_t, int,
                     const __nv_bfloat16*, int, const float*, int, int, int, int, float*, int32_t*, int, cudaStream_t) {
    set_error("conv_fwd: bf16 tensor-core variant not built");
    return RBR_EUNSUPPORTED;
}
}

RBR_DEFINE_OOB_ACCESSOR(conv_tc)
