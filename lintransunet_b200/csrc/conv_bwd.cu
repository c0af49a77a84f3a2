// Backward of the normalisation stage of the conv blocks: y = act(InstanceNorm3d(x)), biased variance,
// eps 1e-5, no affine (model/Unet_3Dblock.py:325-336,:547-554: nn.InstanceNorm3d + LeakyReLU(0.01)).
// x is the RAW convolution output the forward normalised, stats = (mean, rstd) per (sample, channel) from the forward.
//   xhat = (x - mean) rstd ;  dxhat = dy * (xhat > 0 ? 1 : 0.01)
//   dx   = rstd (dxhat - mean_v(dxhat) - xhat mean_v(dxhat xhat))
// Two passes over [B, V, C] channels-last data (partials + ordered finalize, then the elementwise pass): 5 V C E bytes,
// fixed-order sums => bit-reproducible.  A residual added after the activation just receives dy (no kernel).
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

__device__ __forceinline__ float act_slope(float xhat, int act) { return (act == LTU_ACT_LRELU && !(xhat > 0.f)) ? 0.01f : 1.f; }

// grid (chunks, B); like chan_partials_kernel: thread = (4 channels, voxel phase); partials [B][chunks][C][2]
template <typename T>
__global__ void __launch_bounds__(256)
instnorm_bwd_partials_kernel(const T* __restrict__ x, const float* __restrict__ stats, const T* __restrict__ dy,
                             float* __restrict__ partials, int64_t V, int C, int chunks, int act) {
    __shared__ float red[2 * 256 * 4];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cg = C / 4;
    const int rows = 256 / cg;
    const int my_cg = threadIdx.x % cg, my_row = threadIdx.x / cg;
    const int64_t per = ceil_div64(V, chunks);
    const int64_t v0 = (int64_t)chunk * per, v1 = v0 + per < V ? v0 + per : V;
    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
    if (my_row < rows) {
        float mean[4], rstd[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            mean[i] = stats[((int64_t)b * C + my_cg * 4 + i) * 2];
            rstd[i] = stats[((int64_t)b * C + my_cg * 4 + i) * 2 + 1];
        }
        for (int64_t v = v0 + my_row; v < v1; v += rows) {
            float xv[4], g[4];
            load4(x + ((int64_t)b * V + v) * C + my_cg * 4, xv);
            load4(dy + ((int64_t)b * V + v) * C + my_cg * 4, g);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xh = (xv[i] - mean[i]) * rstd[i];
                const float d = g[i] * act_slope(xh, act);
                s1[i] += d;
                s2[i] = fmaf(d, xh, s2[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { red[(threadIdx.x * 4 + i) * 2] = s1[i]; red[(threadIdx.x * 4 + i) * 2 + 1] = s2[i]; }
    __syncthreads();
    if (threadIdx.x < C) {
        const int c = threadIdx.x, g = c / 4, i = c % 4;
        float a = 0.f, bb = 0.f;
        for (int r = 0; r < rows; ++r) {
            const int t = r * cg + g;
            a += red[(t * 4 + i) * 2];
            bb += red[(t * 4 + i) * 2 + 1];
        }
        float* dst = partials + (((int64_t)b * chunks + chunk) * C + c) * 2;
        dst[0] = a; dst[1] = bb;
    }
}

// msum [B][C][2] = (mean_v dxhat, mean_v dxhat*xhat): ordered sum over the chunks
__global__ void __launch_bounds__(256)
instnorm_bwd_finalize_kernel(const float* __restrict__ partials, float* __restrict__ msum, int chunks, int C, float inv_v) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = 0.f, bb = 0.f;
    for (int k = 0; k < chunks; ++k) {
        const float* src = partials + (((int64_t)b * chunks + k) * C + c) * 2;
        a += src[0]; bb += src[1];
    }
    msum[((int64_t)b * C + c) * 2] = a * inv_v;
    msum[((int64_t)b * C + c) * 2 + 1] = bb * inv_v;
}

template <typename T>
__global__ void __launch_bounds__(256)
instnorm_bwd_apply_kernel(const T* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ msum,
                          const T* __restrict__ dy, T* __restrict__ dx, int64_t V, int C, int act) {
    const int b = blockIdx.y;
    const int cg = C / 4;
    const int64_t total = V * cg;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % cg) * 4;
        const int64_t off = ((int64_t)b * V + idx / cg) * C + c4;
        float xv[4], g[4], o[4];
        load4(x + off, xv);
        load4(dy + off, g);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float mean = stats[((int64_t)b * C + c4 + i) * 2], rstd = stats[((int64_t)b * C + c4 + i) * 2 + 1];
            const float m1 = msum[((int64_t)b * C + c4 + i) * 2], m2 = msum[((int64_t)b * C + c4 + i) * 2 + 1];
            const float xh = (xv[i] - mean) * rstd;
            o[i] = rstd * (g[i] * act_slope(xh, act) - m1 - xh * m2);
        }
        store4(dx + off, o);
    }
}

// z [B, Hz, Wz, Dz, C] = 0 except z[b, h*sh, w*sw, d*sd, :] = y[b, h, w, d, :]: the zero-insertion that turns the input
// gradient of a strided convolution into a stride-1 convolution of z with the reversed, transposed filter
template <typename T>
__global__ void __launch_bounds__(256)
zero_insert_kernel(const T* __restrict__ y, T* __restrict__ z, int B, int H, int W, int D, int C, int Hz, int Wz, int Dz,
                   int sh, int sw, int sd) {
    const int cg = C / 4;
    const int64_t total = (int64_t)B * Hz * Wz * Dz * cg;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % cg) * 4;
        int64_t t = idx / cg;
        const int d = (int)(t % Dz); t /= Dz;
        const int w = (int)(t % Wz); t /= Wz;
        const int h = (int)(t % Hz);
        const int b = (int)(t / Hz);
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (h % sh == 0 && w % sw == 0 && d % sd == 0 && h / sh < H && w / sw < W && d / sd < D)
            load4(y + ((((int64_t)b * H + h / sh) * W + w / sw) * D + d / sd) * C + c4, v);
        store4(z + (idx / cg) * C + c4, v);
    }
}

// y [B, H, W, D, C] = sum of the 2x2x2 block of x [B, 2H, 2W, 2D, C]: backward of nn.Upsample(nearest, x2)
template <typename T>
__global__ void __launch_bounds__(256)
sumpool2_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int H, int W, int D, int C) {
    const int cg = C / 4;
    const int64_t total = (int64_t)B * H * W * D * cg;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % cg) * 4;
        int64_t t = idx / cg;
        const int d = (int)(t % D); t /= D;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v[4];
            load4(x + ((((int64_t)b * 2 * H + 2 * h + (k >> 2)) * 2 * W + 2 * w + ((k >> 1) & 1)) * 2 * D + 2 * d + (k & 1)) * C + c4, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) s[i] += v[i];
        }
        store4(y + (idx / cg) * C + c4, s);
    }
}

static inline int in_bwd_chunks(int64_t V) {
    int64_t c = V / 512;
    if (c > 256) c = 256;
    return (int)(c < 1 ? 1 : c);
}

}  // namespace ltu

using namespace ltu;

static bool a16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" size_t ltu_instnorm_bwd_workspace(int B, int64_t voxels, int C) {
    if (B <= 0 || voxels <= 0 || C <= 0) return 0;
    return ((size_t)B * in_bwd_chunks(voxels) * C * 2 + (size_t)B * C * 2) * sizeof(float);
}

extern "C" int ltu_instnorm_bwd(const void* x, const float* stats, const void* dy, void* dx, void* ws, size_t ws_bytes,
                                int B, int64_t voxels, int C, int act, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && stats && dy && dx && ws, "instnorm_bwd: null pointer");
    LTU_ARG_CHECK(B > 0 && B <= 65535 && voxels > 0, "instnorm_bwd: bad shape");
    LTU_ARG_CHECK(C % 4 == 0 && C >= 4 && C <= 256 && 256 % (C / 4) == 0, "instnorm_bwd: C=%d must be a power of two in [4, 256]", C);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "instnorm_bwd: bad dtype %d", dtype);
    LTU_ARG_CHECK(act == LTU_ACT_NONE || act == LTU_ACT_LRELU, "instnorm_bwd: bad act %d", act);
    LTU_ARG_CHECK(a16(x) && a16(dy) && a16(dx), "instnorm_bwd: pointers must be 16-byte aligned");
    LTU_ARG_CHECK(ws_bytes >= ltu_instnorm_bwd_workspace(B, voxels, C), "instnorm_bwd: workspace too small");
    const int chunks = in_bwd_chunks(voxels);
    float* partials = (float*)ws;
    float* msum = partials + (size_t)B * chunks * C * 2;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LTU_F32) instnorm_bwd_partials_kernel<float><<<dim3(chunks, B), 256, 0, st>>>((const float*)x, stats, (const float*)dy, partials, voxels, C, chunks, act);
    else instnorm_bwd_partials_kernel<bf16><<<dim3(chunks, B), 256, 0, st>>>((const bf16*)x, stats, (const bf16*)dy, partials, voxels, C, chunks, act);
    LTU_LAUNCH_CHECK("instnorm_bwd_partials");
    instnorm_bwd_finalize_kernel<<<dim3((C + 255) / 256, B), 256, 0, st>>>(partials, msum, chunks, C, (float)(1.0 / (double)voxels));
    LTU_LAUNCH_CHECK("instnorm_bwd_finalize");
    int64_t bx = ceil_div64(voxels * (C / 4), 256);
    const int64_t cap = ceil_div64((int64_t)sm_count() * 16, B);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    if (dtype == LTU_F32) instnorm_bwd_apply_kernel<float><<<dim3((unsigned)bx, B), 256, 0, st>>>((const float*)x, stats, msum, (const float*)dy, (float*)dx, voxels, C, act);
    else instnorm_bwd_apply_kernel<bf16><<<dim3((unsigned)bx, B), 256, 0, st>>>((const bf16*)x, stats, msum, (const bf16*)dy, (bf16*)dx, voxels, C, act);
    LTU_LAUNCH_CHECK("instnorm_bwd_apply");
    count_launch(3);
    return LTU_OK;
}

static unsigned ew_grid(int64_t items) {
    int64_t b = ceil_div64(items, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

extern "C" int ltu_zero_insert(const void* y, void* z, int B, int H, int W, int D, int C, int Hz, int Wz, int Dz, int sh,
                               int sw, int sd, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(y && z && B > 0 && H > 0 && W > 0 && D > 0 && C > 0 && C % 4 == 0, "zero_insert: bad arguments");
    LTU_ARG_CHECK(sh >= 1 && sw >= 1 && sd >= 1 && Hz >= (H - 1) * sh + 1 && Wz >= (W - 1) * sw + 1 && Dz >= (D - 1) * sd + 1,
                  "zero_insert: the target volume is too small for the stride");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "zero_insert: bad dtype %d", dtype);
    LTU_ARG_CHECK(a16(y) && a16(z), "zero_insert: pointers must be 16-byte aligned");
    const unsigned g = ew_grid((int64_t)B * Hz * Wz * Dz * (C / 4));
    if (dtype == LTU_F32) zero_insert_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)y, (float*)z, B, H, W, D, C, Hz, Wz, Dz, sh, sw, sd);
    else zero_insert_kernel<bf16><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, (bf16*)z, B, H, W, D, C, Hz, Wz, Dz, sh, sw, sd);
    LTU_LAUNCH_CHECK("zero_insert");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_sumpool2(const void* x, void* y, int B, int H, int W, int D, int C, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && y && B > 0 && H > 0 && W > 0 && D > 0 && C > 0 && C % 4 == 0, "sumpool2: bad arguments");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "sumpool2: bad dtype %d", dtype);
    LTU_ARG_CHECK(a16(x) && a16(y), "sumpool2: pointers must be 16-byte aligned");
    const unsigned g = ew_grid((int64_t)B * H * W * D * (C / 4));
    if (dtype == LTU_F32) sumpool2_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, B, H, W, D, C);
    else sumpool2_kernel<bf16><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, H, W, D, C);
    LTU_LAUNCH_CHECK("sumpool2");
    count_launch(1);
    return LTU_OK;
}
