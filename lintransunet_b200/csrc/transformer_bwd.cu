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

// dgamma[c] = sum over the CTAs of part[cta][0][c]; dbeta likewise.  One warp per output: the lanes take the CTAs
// cta = lane, lane + 32, ... in order, then a butterfly -- the same summation tree every run.
__global__ void __launch_bounds__(256)
ln_bwd_finalize_kernel(const float* __restrict__ part, int nblocks, int C, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= 2 * C) return;                                  // warp-uniform
    const int which = c / C, col = c - which * C;
    float s = 0.f;
    for (int b = lane; b < nblocks; b += 32) s += part[((int64_t)b * 2 + which) * C + col];
    s = warp_sum(s);
    if (lane == 0) (which == 0 ? dgamma : dbeta)[col] = s;
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

// ---------------------------------------------------------------- Conv3dPosEmbedding weight / bias gradient
// y = x + bias + sum_tap w[tap][c] x[vox + off(tap)][c]  (model/trans_block.py:86-96 in native axes):
//   dw[tap][c] = sum_vox dy[vox][c] x[vox + off(tap)][c],  dbias[c] = sum_vox dy[vox][c]
// (dx needs no kernel of its own: it is the forward kernel applied to dy with the taps reversed and a zero bias).
// A thread owns 4 channels and every `phases`-th voxel of its CTA's slice; 28 x 4 accumulators in registers; the
// voxel phases are folded into shared memory one after the other (fixed order), then one partial row per CTA.
template <typename T>
__global__ void __launch_bounds__(256)
posenc_wgrad_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ part, int B, int H, int W,
                    int D, int C, int64_t vox_per_cta) {
    extern __shared__ float sacc[];                          // [28][C]
    const int cv = C / 4, phases = 256 / cv;
    const int cg = threadIdx.x % cv, ph = threadIdx.x / cv;
    const int c4 = cg * 4;
    const int64_t total = (int64_t)B * H * W * D;
    const int64_t v0 = (int64_t)blockIdx.x * vox_per_cta;
    int64_t v1 = v0 + vox_per_cta;
    if (v1 > total) v1 = total;
    float acc[28][4];
#pragma unroll
    for (int t = 0; t < 28; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[t][i] = 0.f;
    for (int64_t vox = v0 + ph; vox < v1; vox += phases) {
        const int d = (int)(vox % D);
        int64_t t = vox / D;
        const int ww = (int)(t % W);
        t /= W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        float g[4];
        load4(dy + vox * C + c4, g);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[27][i] += g[i];
        // clamped coordinates + a 0/1 mask instead of branches: the 27 neighbour loads are independent and issued together
        // (guarded by `continue`s they were serialised, one L2 round trip each)
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hh = h + kh - 1;
            const bool vh = hh >= 0 && hh < H;
            const int hc = vh ? hh : h;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int w2 = ww + kw - 1;
                const bool vw = w2 >= 0 && w2 < W;
                const int wc = vw ? w2 : ww;
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const int dd = d + kd - 1;
                    const bool vd = dd >= 0 && dd < D;
                    const int dc = vd ? dd : d;
                    const float m = (vh && vw && vd) ? 1.f : 0.f;
                    float xv[4];
                    const int64_t nv = (((int64_t)b * H + hc) * W + wc) * D + dc;
                    load4(x + nv * C + c4, xv);
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[kh * 9 + kw * 3 + kd][i] = fmaf(g[i] * m, xv[i], acc[kh * 9 + kw * 3 + kd][i]);
                }
            }
        }
    }
    for (int p = 0; p < phases; ++p) {                       // ordered fold of the voxel phases
        if (ph == p) {
#pragma unroll
            for (int t = 0; t < 28; ++t)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float* s = sacc + t * C + c4 + i;
                    *s = p == 0 ? acc[t][i] : *s + acc[t][i];
                }
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < 28 * C; i += 256) part[(int64_t)blockIdx.x * 28 * C + i] = sacc[i];
}

// out[i] = sum over CTAs (in order) of part[cta][i], i < 28*C: rows 0..26 = dw[tap][c], row 27 = dbias[c]
__global__ void __launch_bounds__(256)
posenc_wgrad_finalize_kernel(const float* __restrict__ part, int nblocks, int n, float* __restrict__ dw,
                             float* __restrict__ dbias, int C) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += part[(int64_t)b * n + i];
    if (i < 27 * C) dw[i] = s;
    else dbias[i - 27 * C] = s;
}

static inline int posenc_wgrad_blocks(int64_t voxels) {
    int64_t blocks = ceil_div64(voxels, 64);                 // at least 64 voxels per CTA
    const int64_t cap = (int64_t)sm_count() * 2;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

static inline int ln_bwd_blocks(int64_t rows) {
    int64_t blocks = ceil_div64(rows, 8);
    const int64_t cap = (int64_t)sm_count() * 4;             // ONE resident wave (48-64 registers, 9-17 KB: >= 4 CTAs / SM): the rows are
                                                             // grid-strided, a second partial wave only adds a tail
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
    ln_bwd_finalize_kernel<<<(2 * C + 7) / 8, 256, 0, st>>>(part, blocks, C, dgamma, dbeta);
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

extern "C" size_t ltu_posenc_wgrad_workspace(int B, int H, int W, int D, int C) {
    if (B <= 0 || H <= 0 || W <= 0 || D <= 0 || C <= 0) return 0;
    return (size_t)posenc_wgrad_blocks((int64_t)B * H * W * D) * 28 * C * sizeof(float);
}

extern "C" int ltu_posenc_wgrad(const void* x, const void* dy, float* dw, float* dbias, void* ws, size_t ws_bytes, int B,
                                int H, int W, int D, int C, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && dy && dw && dbias && ws, "posenc_wgrad: null pointer");
    LTU_ARG_CHECK(B > 0 && H > 0 && W > 0 && D > 0, "posenc_wgrad: bad shape");
    LTU_ARG_CHECK(C == 128 || C == 256 || C == 64 || C == 32, "posenc_wgrad: C must be 32, 64, 128 or 256 (got %d)", C);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "posenc_wgrad: bad dtype %d", dtype);
    LTU_ARG_CHECK(al16(x) && al16(dy), "posenc_wgrad: pointers must be 16-byte aligned");
    LTU_ARG_CHECK(ws_bytes >= ltu_posenc_wgrad_workspace(B, H, W, D, C), "posenc_wgrad: workspace too small");
    const int64_t voxels = (int64_t)B * H * W * D;
    const int blocks = posenc_wgrad_blocks(voxels);
    const int64_t vox_per_cta = ceil_div64(voxels, blocks);
    const size_t smem = (size_t)28 * C * sizeof(float);      // <= 28 KB
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LTU_F32) posenc_wgrad_kernel<float><<<blocks, 256, smem, st>>>((const float*)x, (const float*)dy, (float*)ws, B, H, W, D, C, vox_per_cta);
    else posenc_wgrad_kernel<bf16><<<blocks, 256, smem, st>>>((const bf16*)x, (const bf16*)dy, (float*)ws, B, H, W, D, C, vox_per_cta);
    LTU_LAUNCH_CHECK("posenc_wgrad");
    posenc_wgrad_finalize_kernel<<<(28 * C + 255) / 256, 256, 0, st>>>((const float*)ws, blocks, 28 * C, dw, dbias, C);
    LTU_LAUNCH_CHECK("posenc_wgrad_finalize");
    count_launch(2);
    return LTU_OK;
}
