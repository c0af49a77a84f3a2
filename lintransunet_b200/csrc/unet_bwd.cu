// Backward of the decoder glue: trilinear x(2,2,fd) upsample with align_corners (model/Unet_3Dblock.py:
// 1341-1345,:1375-1378), the mask-head softmax (:1380-1387) and the output head (windows_unembedding + softmax, :138-152,
// :1392-1394).  All are GATHERS (every gradient element is computed by exactly one thread from a fixed list of terms), so
// the results are bit-reproducible; fp32 arithmetic, HBM bound.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

// forward tap of unet_kernels.cu::lin_tap, restated: output o reads inputs i0, i1 with weights l0, l1
__device__ __forceinline__ void lin_tap_b(int o, int in_size, float ratio, int& i0, int& i1, float& l0, float& l1) {
    const float r = ratio * (float)o;
    i0 = (int)r;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = r - (float)i0;
    l0 = 1.f - l1;
}

// adjoint of one axis: the outputs o that read input i, with their weights (recomputed with the forward's own
// arithmetic, so the pair is the exact transpose).  At most 6 for a x2 upsample.
__device__ __forceinline__ int axis_adjoint(int i, int n, int m, float ratio, int (&idx)[8], float (&wt)[8]) {
    int cnt = 0;
    int lo = 0, hi = m - 1;
    if (ratio > 0.f) {
        lo = (int)floorf((float)(i - 1) / ratio) - 1;
        hi = (int)ceilf((float)(i + 1) / ratio) + 1;
        if (lo < 0) lo = 0;
        if (hi > m - 1) hi = m - 1;
    }
    for (int o = lo; o <= hi; ++o) {
        int i0, i1;
        float l0, l1;
        lin_tap_b(o, n, ratio, i0, i1, l0, l1);
        const float w = (i0 == i ? l0 : 0.f) + (i1 == i ? l1 : 0.f);
        if (w != 0.f && cnt < 8) { idx[cnt] = o; wt[cnt] = w; ++cnt; }
    }
    return cnt;
}

// dy [B,2H,2W,fd*D,C] -> dx [B,H,W,D,C]
template <typename T>
__global__ void __launch_bounds__(256)
upsample_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int H, int W, int D, int C, int fd) {
    const int Ho = 2 * H, Wo = 2 * W, Do = fd * D;
    const int cg = C / 4;
    const float rh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
    const float rw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    const float rd = Do > 1 ? (float)(D - 1) / (float)(Do - 1) : 0.f;
    const int64_t total = (int64_t)B * H * W * D * cg;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % cg) * 4;
        int64_t t = idx / cg;
        const int d = (int)(t % D); t /= D;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        int oh[8], ow[8], od[8];
        float wh[8], ww[8], wd[8];
        const int nh = axis_adjoint(h, H, Ho, rh, oh, wh);
        const int nw = axis_adjoint(w, W, Wo, rw, ow, ww);
        int nd;
        if (fd == 1) { nd = 1; od[0] = d; wd[0] = 1.f; }
        else nd = axis_adjoint(d, D, Do, rd, od, wd);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const T* yb = dy + (int64_t)b * Ho * Wo * Do * C + c4;
        for (int a = 0; a < nh; ++a)
            for (int bb = 0; bb < nw; ++bb) {
                const float whw = wh[a] * ww[bb];
                const T* row = yb + ((int64_t)oh[a] * Wo + ow[bb]) * Do * C;
                for (int cc = 0; cc < nd; ++cc) {
                    float v[4];
                    load4(row + (int64_t)od[cc] * C, v);
                    const float wt = whw * wd[cc];
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] = fmaf(wt, v[i], acc[i]);
                }
            }
        store4(dx + (idx / cg) * C + c4, acc);
    }
}

// mask head: p = softmax_c(logits [B,V,C]); mask [B,C,V] = p  ->  dlogits[v][c] = p_c (dmask_c - sum_k p_k dmask_k)
__global__ void __launch_bounds__(256)
mask_softmax_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dmask, float* __restrict__ dlogits,
                        int64_t V, int Cout) {
    const int b = blockIdx.y;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        const float* l = logits + ((int64_t)b * V + v) * Cout;
        float e[8];
        float mx = l[0];
        for (int c = 1; c < Cout; ++c) mx = fmaxf(mx, l[c]);
        float s = 0.f;
        for (int c = 0; c < Cout; ++c) { e[c] = expf(l[c] - mx); s += e[c]; }
        float dot = 0.f;
        for (int c = 0; c < Cout; ++c) {
            e[c] /= s;
            dot = fmaf(e[c], dmask[((int64_t)b * Cout + c) * V + v], dot);
        }
        for (int c = 0; c < Cout; ++c)
            dlogits[((int64_t)b * V + v) * Cout + c] = e[c] * (dmask[((int64_t)b * Cout + c) * V + v] - dot);
    }
}

// output head: logits [B,H2,W2,D,4*COUT] (in-channel = c*4 + kh*2 + kw) -> probs [B,COUT,2*H2,2*W2,D] = softmax over c
template <int COUT>
__global__ void __launch_bounds__(256)
head_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ dprobs, float* __restrict__ dlogits, int B,
                int H2, int W2, int D) {
    const int H = 2 * H2, W = 2 * W2;
    const int64_t total = (int64_t)B * H2 * W2 * D;
    const int64_t plane = (int64_t)H * W * D;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(idx % D);
        int64_t t = idx / D;
        const int w2 = (int)(t % W2); t /= W2;
        const int h2 = (int)(t % H2);
        const int b = (int)(t / H2);
        const float* src = logits + idx * (4 * COUT);
        float* dst = dlogits + idx * (4 * COUT);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int hh = 2 * h2 + (k >> 1), ww = 2 * w2 + (k & 1);
            float e[COUT];
            float mx = src[k];
#pragma unroll
            for (int c = 1; c < COUT; ++c) mx = fmaxf(mx, src[c * 4 + k]);
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < COUT; ++c) { e[c] = expf(src[c * 4 + k] - mx); s += e[c]; }
            float g[COUT], dot = 0.f;
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                e[c] /= s;
                g[c] = dprobs[((int64_t)b * COUT + c) * plane + ((int64_t)hh * W + ww) * D + d];
                dot = fmaf(e[c], g[c], dot);
            }
#pragma unroll
            for (int c = 0; c < COUT; ++c) dst[c * 4 + k] = e[c] * (g[c] - dot);
        }
    }
}

// ---------------------------------------------------------------- attention gate (a13)
// forward (unet_kernels.cu::gate_kernel): h = relu(norm(a) + norm(g)), z = psi.h + b, s = sigmoid(z), out = skip * s
// backward, given dout:   dskip = dout * s            (the direct path; the W_x path is added by the caller)
//                         dz = (sum_c dout_c skip_c) s (1 - s);  dh_c = dz psi_c [h_c > 0]   (= d norm(a) = d norm(g))
//                         dpsi_c = sum_v dz h_c;  dpsi_b = sum_v dz        (per-CTA partials, ordered finalize)
// G = Ci / 4 lanes share a voxel (4 channels each), the two per-voxel sums are butterflies inside the lane group.
// part [gridDim.y * gridDim.x][Ci + 1].
template <typename T>
__global__ void __launch_bounds__(256)
gate_bwd_kernel(const T* __restrict__ a, const float* __restrict__ sa, const T* __restrict__ g, const float* __restrict__ sg,
                const float* __restrict__ psi_w, const float* __restrict__ psi_b, const T* __restrict__ skip,
                const T* __restrict__ dout, T* __restrict__ dskip, T* __restrict__ dh, float* __restrict__ part, int64_t V,
                int Ci) {
    __shared__ float red[256][5];
    const int G = Ci / 4;
    const int b = blockIdx.y;
    const int64_t total = V * G;
    const int64_t total_pad = (total + 31) / 32 * 32;
    const float bias = psi_b[0];
    const float* st_a = sa + (int64_t)b * Ci * 2;
    const float* st_g = sg + (int64_t)b * Ci * 2;
    // the grid stride (gridDim.x * 256) is a multiple of G: a thread always owns the same 4 channels
    const int c0 = (int)(threadIdx.x % G) * 4;
    float ma[4], ra[4], mg[4], rg[4], pw[4], acc[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ma[i] = st_a[(c0 + i) * 2]; ra[i] = st_a[(c0 + i) * 2 + 1];
        mg[i] = st_g[(c0 + i) * 2]; rg[i] = st_g[(c0 + i) * 2 + 1];
        pw[i] = psi_w[c0 + i];
    }
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_pad; idx += (int64_t)gridDim.x * blockDim.x) {
        const bool live = idx < total;
        const int64_t cidx = live ? idx : total - 1;
        const int64_t off = (int64_t)b * V * Ci + (cidx / G) * Ci + c0;
        float va[4], vg[4], vs[4], vd[4], h[4];
        load4(a + off, va);
        load4(g + off, vg);
        load4(skip + off, vs);
        load4(dout + off, vd);
        float dot = 0.f, ds = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            h[i] = fmaxf((va[i] - ma[i]) * ra[i] + (vg[i] - mg[i]) * rg[i], 0.f);
            dot = fmaf(h[i], pw[i], dot);
            ds = fmaf(vd[i], vs[i], ds);
        }
        for (int o = G >> 1; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xffffffffu, dot, o);
            ds += __shfl_xor_sync(0xffffffffu, ds, o);
        }
        const float s = 1.f / (1.f + __expf(-(dot + bias)));
        const float dz = ds * s * (1.f - s);
        if (live) {
            float o1[4], o2[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                o1[i] = vd[i] * s;
                o2[i] = h[i] > 0.f ? dz * pw[i] : 0.f;
                acc[i] = fmaf(dz, h[i], acc[i]);
            }
            if (c0 == 0) acc[4] += dz;
            store4(dskip + off, o1);
            store4(dh + off, o2);
        }
    }
    // ordered fold over the threads that own the same channels: t, t + G, t + 2G, ...
#pragma unroll
    for (int i = 0; i < 5; ++i) red[threadIdx.x][i] = acc[i];
    __syncthreads();
    float* out = part + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * (Ci + 1);
    if (threadIdx.x < G) {
        float sum[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
        for (int t = threadIdx.x; t < 256; t += G)
#pragma unroll
            for (int i = 0; i < 5; ++i) sum[i] += red[t][i];
#pragma unroll
        for (int i = 0; i < 4; ++i) out[threadIdx.x * 4 + i] = sum[i];
        if (threadIdx.x == 0) out[Ci] = sum[4];
    }
}

// dpsi [Ci] and dpsi_b [1]: one warp per output, lanes take the partial rows in order, then a butterfly
__global__ void __launch_bounds__(256)
gate_bwd_finalize_kernel(const float* __restrict__ part, int nparts, int Ci, float* __restrict__ dpsi_w, float* __restrict__ dpsi_b) {
    const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c > Ci) return;
    float s = 0.f;
    for (int p = lane; p < nparts; p += 32) s += part[(int64_t)p * (Ci + 1) + c];
    s = warp_sum(s);
    if (lane == 0) { if (c < Ci) dpsi_w[c] = s; else dpsi_b[0] = s; }
}

static inline int gate_bwd_blocks(int B, int64_t V, int Ci) {
    const int G = Ci / 4;
    int64_t bx = ceil_div64(V * G, 256);
    const int64_t cap = ceil_div64((int64_t)sm_count() * 8, B);
    if (bx > cap) bx = cap;
    return (int)(bx < 1 ? 1 : bx);
}

static unsigned ub_grid(int64_t items, int waves) {
    int64_t b = ceil_div64(items, 256);
    const int64_t cap = (int64_t)sm_count() * waves;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}

}  // namespace ltu

using namespace ltu;

extern "C" int ltu_upsample_trilinear_bwd(const void* dy, void* dx, int B, int H, int W, int D, int C, int fd, int dtype,
                                          ltu_stream_t stream) {
    LTU_ARG_CHECK(dy && dx && B > 0 && H > 0 && W > 0 && D > 0 && C > 0 && C % 4 == 0, "upsample_trilinear_bwd: bad arguments");
    LTU_ARG_CHECK(fd == 1 || fd == 2, "upsample_trilinear_bwd: depth factor must be 1 or 2");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "upsample_trilinear_bwd: bad dtype %d", dtype);
    LTU_ARG_CHECK(((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0, "upsample_trilinear_bwd: pointers must be 16-byte aligned");
    const unsigned g = ub_grid((int64_t)B * H * W * D * (C / 4), 16);
    if (dtype == LTU_F32) upsample_bwd_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)dy, (float*)dx, B, H, W, D, C, fd);
    else upsample_bwd_kernel<bf16><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16*)dy, (bf16*)dx, B, H, W, D, C, fd);
    LTU_LAUNCH_CHECK("upsample_trilinear_bwd");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_mask_softmax_bwd(const float* logits, const float* dmask, float* dlogits, int B, int64_t voxels, int Cout,
                                    ltu_stream_t stream) {
    LTU_ARG_CHECK(logits && dmask && dlogits && B > 0 && B <= 65535 && voxels > 0, "mask_softmax_bwd: bad arguments");
    LTU_ARG_CHECK(Cout >= 1 && Cout <= 8, "mask_softmax_bwd: Cout must be in [1,8] (got %d)", Cout);
    int64_t bx = ceil_div64(voxels, 256);
    const int64_t cap = ceil_div64((int64_t)sm_count() * 16, B);
    if (bx > cap) bx = cap;
    mask_softmax_bwd_kernel<<<dim3((unsigned)bx, B), 256, 0, (cudaStream_t)stream>>>(logits, dmask, dlogits, voxels, Cout);
    LTU_LAUNCH_CHECK("mask_softmax_bwd");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_head_d2s_softmax_bwd(const float* logits, const float* dprobs, float* dlogits, int B, int H2, int W2, int D,
                                        int Cout, ltu_stream_t stream) {
    LTU_ARG_CHECK(logits && dprobs && dlogits && B > 0 && H2 > 0 && W2 > 0 && D > 0, "head_d2s_softmax_bwd: bad arguments");
    LTU_ARG_CHECK(Cout >= 1 && Cout <= 8, "head_d2s_softmax_bwd: Cout must be in [1,8] (got %d)", Cout);
    const unsigned g = ub_grid((int64_t)B * H2 * W2 * D, 16);
    cudaStream_t st = (cudaStream_t)stream;
    switch (Cout) {
#define HB_CASE(n) case n: head_bwd_kernel<n><<<g, 256, 0, st>>>(logits, dprobs, dlogits, B, H2, W2, D); break;
        HB_CASE(1) HB_CASE(2) HB_CASE(3) HB_CASE(4) HB_CASE(5) HB_CASE(6) HB_CASE(7) HB_CASE(8)
#undef HB_CASE
    }
    LTU_LAUNCH_CHECK("head_d2s_softmax_bwd");
    count_launch(1);
    return LTU_OK;
}

extern "C" size_t ltu_gate_bwd_workspace(int B, int64_t voxels, int Ci) {
    if (B <= 0 || voxels <= 0 || Ci <= 0 || Ci % 4) return 0;
    return (size_t)B * gate_bwd_blocks(B, voxels, Ci) * (Ci + 1) * sizeof(float);
}

extern "C" int ltu_gate_bwd(const void* a, const float* stats_a, const void* g, const float* stats_g, const float* psi_w,
                            const float* psi_b, const void* skip, const void* dout, void* dskip, void* dh, float* dpsi_w,
                            float* dpsi_b, void* ws, size_t ws_bytes, int B, int64_t voxels, int Ci, int dtype,
                            ltu_stream_t stream) {
    LTU_ARG_CHECK(a && stats_a && g && stats_g && psi_w && psi_b && skip && dout && dskip && dh && dpsi_w && dpsi_b && ws,
                  "gate_bwd: null pointer");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "gate_bwd: bad dtype %d", dtype);
    const int G = Ci / 4;
    LTU_ARG_CHECK(Ci % 4 == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0, "gate_bwd: Ci=%d unsupported (Ci/4 must be a power of two <= 32)", Ci);
    LTU_ARG_CHECK(B > 0 && B <= 65535 && voxels > 0, "gate_bwd: bad shape");
    LTU_ARG_CHECK(ws_bytes >= ltu_gate_bwd_workspace(B, voxels, Ci), "gate_bwd: workspace too small");
    const int bx = gate_bwd_blocks(B, voxels, Ci);
    cudaStream_t st = (cudaStream_t)stream;
    float* part = (float*)ws;
    if (dtype == LTU_F32) gate_bwd_kernel<float><<<dim3(bx, B), 256, 0, st>>>((const float*)a, stats_a, (const float*)g, stats_g, psi_w, psi_b, (const float*)skip, (const float*)dout, (float*)dskip, (float*)dh, part, voxels, Ci);
    else gate_bwd_kernel<bf16><<<dim3(bx, B), 256, 0, st>>>((const bf16*)a, stats_a, (const bf16*)g, stats_g, psi_w, psi_b, (const bf16*)skip, (const bf16*)dout, (bf16*)dskip, (bf16*)dh, part, voxels, Ci);
    LTU_LAUNCH_CHECK("gate_bwd");
    gate_bwd_finalize_kernel<<<(Ci + 1 + 7) / 8, 256, 0, st>>>(part, bx * B, Ci, dpsi_w, dpsi_b);
    LTU_LAUNCH_CHECK("gate_bwd_finalize");
    count_launch(2);
    return LTU_OK;
}
