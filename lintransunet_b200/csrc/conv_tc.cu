// tcgen05 implicit-GEMM convolution (bf16 in, fp32 accumulate in TMEM) -- placeholder entry
// points until the kernel lands; ltu_conv3d_tc_supported() == 0 keeps callers on conv_kernels.cu.
#include "common.cuh"
using namespace ltu;
extern "C" int ltu_conv3d_tc_supported(int, int, int, int, int) { return 0; }
extern "C" int ltu_conv3d_tc_tiles(int64_t out_voxels) { return (int)ceil_div64(out_voxels, 128); }
extern "C" int ltu_conv3d_tc(const void*, int, const void*, int, int, int, int, int, int, int, int, int, const void*,
                             const float*, int, void*, int, int, int, float*, ltu_stream_t) {
    set_error("conv3d_tc: not built");
    return LTU_ERR_ARG;
}
