// Training-mode dropout of the reference (p = 0.3 by default, model/trans_3DUnet.py:162):
//   nn.Dropout   (element-wise)  model/trans_block.py:205,:208,:209 and model/Unet_3Dblock.py:339,:382,:429,:556
//   nn.Dropout3d (whole channels of a sample)  model/trans_block.py:96 (Conv3dPosEmbedding)
// as one bandwidth-bound pass  y = keep ? x / (1 - p) : 0  over a channels-last tensor.  The keep decision of element i is
// a pure function of (seed, offset, i) -- Philox4x32-10, four 32-bit draws per counter -- so the backward pass applies the
// SAME kernel to the gradient with the same (seed, offset) instead of storing a mask.  Channel mode keys the draw on
// (sample, channel): every voxel of a channel shares it.  The stream differs from PyTorch's (its draws depend on launch
// geometry), as any re-implementation's does; the distribution (independent Bernoulli(1 - p), scale 1 / (1 - p)) is the same.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0; key.y += W1;
    }
    return ctr;
}

// draws for the VN consecutive indices starting at idx0 (idx0 % 4 == 0)
template <int VN>
__device__ __forceinline__ void draws(uint64_t seed, uint64_t offset, uint64_t idx0, uint32_t (&u)[VN]) {
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
#pragma unroll
    for (int g = 0; g < VN / 4; ++g) {
        const uint64_t c = offset + idx0 / 4 + g;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), 0u, 0u), key);
        u[4 * g] = r.x; u[4 * g + 1] = r.y; u[4 * g + 2] = r.z; u[4 * g + 3] = r.w;
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
dropout_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t nvec, int C, int64_t per_sample, uint32_t thresh,
               float scale, uint64_t seed, uint64_t offset, int channelwise) {
    constexpr int VN = Vec<T>::N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const int64_t e0 = i * VN;
        uint64_t idx0 = (uint64_t)e0;
        if (channelwise) idx0 = (uint64_t)(e0 / per_sample) * (uint64_t)C + (uint64_t)(e0 % C);
        uint32_t u[VN];
        draws<VN>(seed, offset, idx0, u);
        float v[VN];
        load_vec(x + e0, v);
#pragma unroll
        for (int j = 0; j < VN; ++j) v[j] = u[j] >= thresh ? v[j] * scale : 0.f;
        store_vec(y + e0, v);
    }
}

}  // namespace ltu

using namespace ltu;

// y = dropout(x) (y may alias x): n elements of a channels-last tensor with C channels (innermost) and `per_sample`
// elements per batch sample; p in [0, 1).  channelwise 0: nn.Dropout; 1: nn.Dropout3d (one draw per (sample, channel)).
// The draws consume ceil(n / 4) Philox counters starting at `offset` (channel mode: ceil(samples * C / 4)): callers
// advance their offset by that amount.  n and C must be multiples of the 16-byte vector (4 fp32 / 8 bf16).
extern "C" int ltu_dropout(const void* x, void* y, int64_t n, int C, int64_t per_sample, float p, uint64_t seed,
                           uint64_t offset, int channelwise, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && y && n > 0, "dropout: null pointer or empty tensor");
    LTU_ARG_CHECK(p >= 0.f && p < 1.f, "dropout: p must be in [0, 1) (got %f)", (double)p);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "dropout: bad dtype %d", dtype);
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(C > 0 && C % vn == 0 && n % vn == 0 && per_sample > 0 && per_sample % C == 0 && n % per_sample == 0,
                  "dropout: n = %lld, C = %d, per_sample = %lld must be multiples of the 16-byte vector and of each other",
                  (long long)n, C, (long long)per_sample);
    LTU_ARG_CHECK((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "dropout: pointers must be 16-byte aligned");
    const double t = (double)p * 4294967296.0;
    const uint32_t thresh = t >= 4294967295.0 ? 0xffffffffu : (uint32_t)t;      // keep iff draw >= p * 2^32
    const float scale = 1.f / (1.f - p);
    const int64_t nvec = n / vn;
    int64_t blocks = (nvec + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == LTU_F32)
        dropout_kernel<float><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, nvec, C, per_sample,
                                                                           thresh, scale, seed, offset, channelwise);
    else
        dropout_kernel<bf16><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, nvec, C, per_sample,
                                                                          thresh, scale, seed, offset, channelwise);
    LTU_LAUNCH_CHECK("dropout");
    count_launch(1);
    return LTU_OK;
}
