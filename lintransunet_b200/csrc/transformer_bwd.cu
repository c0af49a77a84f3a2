// Backward of the elementwise / row-wise glue of SelfAttentionLayer (model/trans_block.py:203-211), SURVEY 8f-1:
//   add_layernorm_bwd : y = LayerNorm(x + r)      -> dz (= dx = dr), dgamma, dbeta
//   gelu_bwd          : y = gelu(x) (exact erf)   -> dx
// fp32 arithmetic, fp32 or bf16 storage, HBM bound (4 and 3 passes over rows*C elements), fixed-order column
// reductions (per-CTA partials + an ordered finalize) => bit-reproducible parameter gradients.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

// one warp per row, the lane owns the same 4-element chunks as add_layernorm_kernel; z = x + r is rebuilt from the
// forward's inputs (nothing is saved by the forward).  part [gridDim.x][2][C] = per-CTA (dgamma, dbeta) partial sums.
template <typename T, int C>
__global__ void __launch_bounds__(256)
add_layernorm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ r, const T* __restrict__ dy,
                         const float* __restrict__ gamma, T* __restrict__ dz, float* __restrict__ part, int64_t rows,
                         float eps) {
    constexpr int Q = C / 128;
    __shared__ float red[8][2][C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t warp_global = (int64_t)blockIdx.x * 8 + warp;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    float g[Q][4], dg[Q][4], db[Q][4];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        load4(gamma + q * 128 + lane * 4, g[q]);
#pragma unroll
        for (int i = 0; i < 4; ++i) { dg[q][i] = 0.f; db[q][i] = 0.f; }
    }
    for (int64_t row = warp_global; row < rows; row += nwarps) {
        float z[Q][4], d[Q][4];
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float a[4], b[4];
            load4(x + row * C + q * 128 + lane * 4, a);
            load4(r + row * C + q * 128 + lane * 4, b);
            load4(dy + row * C + q * 128 + lane * 4, d[q]);
#pragma unroll
            for (int i = 0; i < 4; ++i) { z[q][i] = a[i] + b[i]; sum += z[q][i]; }
        }
        const float mean = warp_sum(sum) * (1.f / C);
        float sq = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float c = z[q][i] - mean; sq = fmaf(c, c, sq); }
        const float rstd = rsqrtf(warp_sum(sq) * (1.f / C) + eps);
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float xh = (z[q][i] - mean) * rstd;
                const float gd = d[q][i] * g[q][i];
                z[q][i] = xh;                                // keep xhat
                s1 += gd;
                s2 = fmaf(gd, xh, s2);
                dg[q][i] = fmaf(d[q][i], xh, dg[q][i]);
                db[q][i] += d[q][i];
            }
        const float m1 = warp_sum(s1) * (1.f / C), m2 = warp_sum(s2) * (1.f / C);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = rstd * (d[q][i] * g[q][i] - m1 - z[q][i] * m2);
            store4(dz + row * C + q * 128 + lane * 4, o);
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            red[warp][0][q * 128 + lane * 4 + i] = dg[q][i];
            red[warp][1][q * 128 + lane * 4 + i] = db[q][i];
        }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += 256) {
        const int which = c / C, col = c - which * C;
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += red[w][which][col];
        part[((int64_t)blockIdx.x * 2 + which) * C + col] = s;
    }
}

// dgamma[c] = sum over CTAs (in order) of part[cta][0][c]; dbeta likewise
__global__ void __launch_bounds__(256)
ln_bwd_finalize_kernel(const float* __restrict__ part, int nblocks, int C, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= 2 * C) return;
    const int which = c / C, col = c - which * C;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += part[((int64_t)b * 2 + which) * C + col];
    (which == 0 ? dgamma : dbeta)[col] = s;
}

// d/dx [x Phi(x)] = Phi(x) + x phi(x)
template <typename T>
__global__ void __launch_bounds__(256)
gelu_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, int64_t nvec) {
    constexpr int VN = Vec<T>::N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float a[VN], d[VN];
        load_vec(x + i * VN, a);
        load_vec(dy + i * VN, d);
#pragma unroll
        for (int k = 0; k < VN; ++k) {
            const float cdf = 0.5f * (1.f + erff(a[k] * 0.70710678118654752f));
            const float pdf = 0.3989422804014327f * __expf(-0.5f * a[k] * a[k]);
            d[k] *= fmaf(a[k], pdf, cdf);
        }
        store_vec(dx + i * VN, d);
    }
}

static inline int ln_bwd_blocks(int64_t rows) {
    int64_t blocks = ceil_div64(rows, 8);
    const int64_t cap = (int64_t)sm_count() * 6;             // 6 CTAs / SM of loads in flight; the ordered finalize stays short
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace ltu

using namespace ltu;

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" size_t ltu_add_layernorm_bwd_workspace(int64_t rows, int C) {
    if (rows <= 0 || !(C == 128 || C == 256)) return 0;
    return (size_t)ln_bwd_blocks(rows) * 2 * C * sizeof(float);
}

extern "C" int ltu_add_layernorm_bwd(const void* x, const void* res, const void* dy, const float* gamma, void* dz,
                                     float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int64_t rows, int C,
                                     float eps, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && res && dy && gamma && dz && dgamma && dbeta && ws, "add_layernorm_bwd: null pointer");
    LTU_ARG_CHECK(rows > 0, "add_layernorm_bwd: rows=%lld", (long long)rows);
    LTU_ARG_CHECK(C == 128 || C == 256, "add_layernorm_bwd: C must be 128 or 256 (got %d)", C);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "add_layernorm_bwd: bad dtype %d", dtype);
    LTU_ARG_CHECK(al16(x) && al16(res) && al16(dy) && al16(dz) && al16(gamma), "add_layernorm_bwd: pointers must be 16-byte aligned");
    LTU_ARG_CHECK(ws_bytes >= ltu_add_layernorm_bwd_workspace(rows, C), "add_layernorm_bwd: workspace too small");
    const int blocks = ln_bwd_blocks(rows);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)ws;
#define LNB(T, CC) add_layernorm_bwd_kernel<T, CC><<<blocks, 256, 0, st>>>((const T*)x, (const T*)res, (const T*)dy, gamma, (T*)dz, part, rows, eps)
    if (dtype == LTU_F32) { if (C == 128) LNB(float, 128); else LNB(float, 256); }
    else                  { if (C == 128) LNB(bf16, 128);  else LNB(bf16, 256); }
#undef LNB
    LTU_LAUNCH_CHECK("add_layernorm_bwd");
    ln_bwd_finalize_kernel<<<(2 * C + 255) / 256, 256, 0, st>>>(part, blocks, C, dgamma, dbeta);
    LTU_LAUNCH_CHECK("ln_bwd_finalize");
    count_launch(2);
    return LTU_OK;
}

extern "C" int ltu_gelu_bwd(const void* x, const void* dy, void* dx, int64_t n, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && dy && dx && n > 0, "gelu_bwd: bad arguments");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "gelu_bwd: bad dtype %d", dtype);
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(n % vn == 0 && al16(x) && al16(dy) && al16(dx), "gelu_bwd: n must be a multiple of %d, pointers 16-byte aligned", vn);
    const int64_t nvec = n / vn;
    int64_t blocks = ceil_div64(nvec, 256);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == LTU_F32) gelu_bwd_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, (const float*)dy, (float*)dx, nvec);
    else gelu_bwd_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)dy, (bf16*)dx, nvec);
    LTU_LAUNCH_CHECK("gelu_bwd");
    count_launch(1);
    return LTU_OK;
}
