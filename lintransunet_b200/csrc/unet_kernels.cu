// Bandwidth-bound U-Net plumbing around the convolutions and the transformer, all channels-last:
// stem space-to-depth, trilinear upsample, mask-head softmax, fused attention gate, output head
// (depth-to-space + softmax + argmax one-hot), sliding-window vote accumulation.
// Reference: model/Unet_3Dblock.py (:123-152, :217-221, :1341-1345, :1375-1394),
// model/trans_3DUnet.py:196-202.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

// ---------------------------------------------------------------- a10: space-to-depth of the input
template <typename T>
__global__ void __launch_bounds__(256)
s2d_kernel(const float* __restrict__ x, T* __restrict__ y, int B, int H, int W, int D, int cpad) {
    const int H2 = H / 2, W2 = W / 2;
    const int64_t total = (int64_t)B * H2 * W2 * D;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int d = (int)(idx % D);
        int64_t t = idx / D;
        int w2 = (int)(t % W2);
        t /= W2;
        int h2 = (int)(t % H2);
        int b = (int)(t / H2);
        const float* src = x + (((int64_t)b * H + 2 * h2) * W + 2 * w2) * D + d;
        float v[4];
        v[0] = src[0];                       // kh=0,kw=0
        v[1] = src[D];                       // kh=0,kw=1
        v[2] = src[(int64_t)W * D];          // kh=1,kw=0
        v[3] = src[(int64_t)W * D + D];      // kh=1,kw=1
        store4(y + idx * cpad, v);
        if (cpad == 8) {
            const float z[4] = {0.f, 0.f, 0.f, 0.f};
            store4(y + idx * 8 + 4, z);
        }
    }
}

// ---------------------------------------------------------------- a12: trilinear x(2,2,fd), align_corners
__device__ __forceinline__ void lin_tap(int o, int in_size, float ratio, int& i0, int& i1, float& l0, float& l1) {
    float r = ratio * (float)o;
    i0 = (int)r;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = r - (float)i0;
    l0 = 1.f - l1;
}

// One block = a 4 x 4 patch of output (oh, ow) positions x 64 consecutive 16-byte vectors along (od, c): the 16 positions
// read at most 4 x 4 input columns between them (scale 2, align_corners), so 3 of 4 tap loads hit L1 and the kernel moves
// about what it writes instead of 4-8x that through L2 (round 1: one output vector per thread in linear order, 64-bit
// div/mod per element, 1.0-1.5 TB/s of its bytes).  A warp = 32 consecutive vectors of one position: 512-byte
// segments.  The per-output arithmetic (tap weights, fma order) is unchanged, so results are bit-identical.
template <typename T>
__global__ void __launch_bounds__(256)
upsample_kernel(const T* __restrict__ x, T* __restrict__ y, int H, int W, int D, int C, int fd, int tiles_h) {
    constexpr int VN = Vec<T>::N;
    const int Ho = 2 * H, Wo = 2 * W, Do = fd * D;
    const int cv = C / VN;
    const int nvec = Do * cv;                                   // vectors of one output column
    const float rh = Ho > 1 ? (float)(H - 1) / (float)(Ho - 1) : 0.f;
    const float rw = Wo > 1 ? (float)(W - 1) / (float)(Wo - 1) : 0.f;
    const float rd = Do > 1 ? (float)(D - 1) / (float)(Do - 1) : 0.f;
    const int v = (int)blockIdx.x * 64 + (int)(threadIdx.x & 63);
    const int ow = (int)blockIdx.y * 4 + (int)(threadIdx.x >> 6);
    const int b = (int)blockIdx.z / tiles_h, oh0 = ((int)blockIdx.z % tiles_h) * 4;
    if (v >= nvec || ow >= Wo) return;
    const int od = v / cv, c0 = (v - od * cv) * VN;
    int w0, w1, d0, d1;
    float lw0, lw1, ld0, ld1;
    lin_tap(ow, W, rw, w0, w1, lw0, lw1);
    if (fd == 1) { d0 = d1 = od; ld0 = 1.f; ld1 = 0.f; }
    else lin_tap(od, D, rd, d0, d1, ld0, ld1);
    const T* xb = x + (int64_t)b * H * W * D * C + c0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int oh = oh0 + j;
        if (oh >= Ho) break;
        int h0, h1;
        float lh0, lh1;
        lin_tap(oh, H, rh, h0, h1, lh0, lh1);
        float acc[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] = 0.f;
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int bb = 0; bb < 2; ++bb) {
                const int hh = a ? h1 : h0, ww = bb ? w1 : w0;
                const float whw = (a ? lh1 : lh0) * (bb ? lw1 : lw0);
                const T* row = xb + ((int64_t)hh * W + ww) * D * C;
                float v0[VN];
                load_vec(row + (int64_t)d0 * C, v0);
                const float wt0 = whw * ld0;
#pragma unroll
                for (int i = 0; i < VN; ++i) acc[i] = fmaf(wt0, v0[i], acc[i]);
                if (ld1 != 0.f) {
                    float v1[VN];
                    load_vec(row + (int64_t)d1 * C, v1);
                    const float wt1 = whw * ld1;
#pragma unroll
                    for (int i = 0; i < VN; ++i) acc[i] = fmaf(wt1, v1[i], acc[i]);
                }
            }
        store_vec(y + ((((int64_t)b * Ho + oh) * Wo + ow) * Do + od) * C + c0, acc);
    }
}

// ---------------------------------------------------------------- a12: mask head softmax + foreground
__global__ void __launch_bounds__(256)
mask_softmax_kernel(const float* __restrict__ logits, float* __restrict__ mask, float* __restrict__ fg,
                    int64_t V, int Cout) {
    const int b = blockIdx.y;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        const float* l = logits + ((int64_t)b * V + v) * Cout;
        float e[8];
        float mx = l[0];
        for (int c = 1; c < Cout; ++c) mx = fmaxf(mx, l[c]);
        float s = 0.f;
        for (int c = 0; c < Cout; ++c) { e[c] = expf(l[c] - mx); s += e[c]; }
        float p0 = e[0] / s;
        if (mask != nullptr)
            for (int c = 0; c < Cout; ++c) mask[((int64_t)b * Cout + c) * V + v] = e[c] / s;
        fg[(int64_t)b * V + v] = 1.f - p0;
    }
}

// ---------------------------------------------------------------- a13: attention gate
template <typename T>
__global__ void __launch_bounds__(256)
gate_kernel(const T* __restrict__ a, const float* __restrict__ sa, const T* __restrict__ g,
            const float* __restrict__ sg, const float* __restrict__ psi_w, const float* __restrict__ psi_b,
            const T* __restrict__ skip, T* __restrict__ out, int64_t V, int Ci) {
    constexpr int VN = Vec<T>::N;
    const int G = Ci / VN;                       // lanes per voxel (power of two, <= 32)
    const int b = blockIdx.y;
    const int64_t total = V * G;
    const int64_t total_pad = (total + 31) / 32 * 32;
    const float bias = psi_b[0];
    const float* st_a = sa + (int64_t)b * Ci * 2;
    const float* st_g = sg + (int64_t)b * Ci * 2;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total_pad;
         idx += (int64_t)gridDim.x * blockDim.x) {
        bool live = idx < total;
        int64_t cidx = live ? idx : total - 1;
        int c0 = (int)(cidx % G) * VN;
        int64_t off = (int64_t)b * V * Ci + cidx * VN;
        float va[VN], vg[VN];
        load_vec(a + off, va);
        load_vec(g + off, vg);
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < VN; ++i) {
            float na = (va[i] - __ldg(st_a + (c0 + i) * 2)) * __ldg(st_a + (c0 + i) * 2 + 1);
            float ng = (vg[i] - __ldg(st_g + (c0 + i) * 2)) * __ldg(st_g + (c0 + i) * 2 + 1);
            float r = fmaxf(na + ng, 0.f);
            dot = fmaf(r, __ldg(psi_w + c0 + i), dot);
        }
        for (int o = G >> 1; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        float gate = 1.f / (1.f + __expf(-(dot + bias)));
        if (live) {
            float vs[VN];
            load_vec(skip + off, vs);
#pragma unroll
            for (int i = 0; i < VN; ++i) vs[i] *= gate;
            store_vec(out + off, vs);
        }
    }
}

// ---------------------------------------------------------------- a12/a15/a16: output head
template <int COUT>
__global__ void __launch_bounds__(256)
head_kernel(const float* __restrict__ logits, float* __restrict__ probs, float* __restrict__ onehot,
            uint8_t* __restrict__ labels, int B, int H2, int W2, int D) {
    const int H = 2 * H2, W = 2 * W2;
    const int64_t total = (int64_t)B * H2 * W2 * D;
    const int64_t plane = (int64_t)H * W * D;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int d = (int)(idx % D);
        int64_t t = idx / D;
        int w2 = (int)(t % W2);
        t /= W2;
        int h2 = (int)(t % H2);
        int b = (int)(t / H2);
        float l[4 * COUT];
        const float* src = logits + idx * (4 * COUT);
#pragma unroll
        for (int i = 0; i < COUT; ++i) {
            float4 v = *reinterpret_cast<const float4*>(src + 4 * i);
            l[4 * i] = v.x; l[4 * i + 1] = v.y; l[4 * i + 2] = v.z; l[4 * i + 3] = v.w;
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int hh = 2 * h2 + (k >> 1), ww = 2 * w2 + (k & 1);
            float mx = l[k];
#pragma unroll
            for (int c = 1; c < COUT; ++c) mx = fmaxf(mx, l[c * 4 + k]);
            float e[COUT], s = 0.f;
#pragma unroll
            for (int c = 0; c < COUT; ++c) { e[c] = expf(l[c * 4 + k] - mx); s += e[c]; }
            float best = -1.f;
            int arg = 0;
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                e[c] = e[c] / s;
                if (e[c] > best) { best = e[c]; arg = c; }
            }
            int64_t vox = ((int64_t)hh * W + ww) * D + d;
#pragma unroll
            for (int c = 0; c < COUT; ++c) {
                int64_t o = ((int64_t)b * COUT + c) * plane + vox;
                if (probs != nullptr) probs[o] = e[c];
                if (onehot != nullptr) onehot[o] = (c == arg) ? 1.f : 0.f;
            }
            if (labels != nullptr) labels[(int64_t)b * plane + vox] = (uint8_t)arg;
        }
    }
}

// ---------------------------------------------------------------- config 5: vote accumulation
__global__ void __launch_bounds__(256)
vote_accumulate_kernel(const uint8_t* __restrict__ labels, const int32_t* __restrict__ starts,
                       uint8_t* __restrict__ votes, int rh, int rw, int rd, int C, int H, int W, int D) {
    const int win = blockIdx.y;
    const int64_t wv = (int64_t)rh * rw * rd;
    const int sh = starts[win * 3], sw = starts[win * 3 + 1], sd = starts[win * 3 + 2];
    const int64_t plane = (int64_t)H * W * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < wv; i += (int64_t)gridDim.x * blockDim.x) {
        int d = (int)(i % rd);
        int64_t t = i / rd;
        int w = (int)(t % rw);
        int h = (int)(t / rw);
        int lab = labels[(int64_t)win * wv + i];
        if (lab >= C) continue;
        int64_t byte = (int64_t)lab * plane + ((int64_t)(sh + h) * W + (sw + w)) * D + (sd + d);
        unsigned int* word = reinterpret_cast<unsigned int*>(votes + (byte & ~(int64_t)3));
        atomicAdd(word, 1u << (8 * (int)(byte & 3)));   // counts stay <= 255: no carry between bytes
    }
}

__global__ void __launch_bounds__(256)
vote_argmax_kernel(const uint8_t* __restrict__ votes, uint8_t* __restrict__ labels, int C, int64_t V) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        int best = -1, arg = 0;
        for (int c = 0; c < C; ++c) {
            int n = votes[(int64_t)c * V + v];
            if (n > best) { best = n; arg = c; }
        }
        labels[v] = (uint8_t)arg;
    }
}

// votes uint8 [C,V] -> frac fp32 [C,V] = votes / sum_c votes  (MONAI: output_image / count_map with a
// constant importance map and one-hot window predictions)
__global__ void __launch_bounds__(256)
vote_fractions_kernel(const uint8_t* __restrict__ votes, float* __restrict__ frac, int C, int64_t V) {
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        int tot = 0;
        for (int c = 0; c < C; ++c) tot += votes[(int64_t)c * V + v];
        // IEEE division like MONAI's `output_image / count_map` (k * (1/n) can differ in the last bit when n is not a
        // power of two: overlap 0.6 gives n = 3, 6, ...); tot == 0 cannot happen: every voxel is covered by >= 1 window
        for (int c = 0; c < C; ++c) frac[(int64_t)c * V + v] = __fdiv_rn((float)votes[(int64_t)c * V + v], (float)tot);
    }
}

// out[n,0,h,w,d] = vol[sh+h, sw+w, sd+d]: batches sliding windows for the predictor
__global__ void __launch_bounds__(256)
gather_windows_kernel(const float* __restrict__ vol, const int32_t* __restrict__ starts, float* __restrict__ out,
                      int rh, int rw, int rd, int H, int W, int D) {
    const int win = blockIdx.y;
    const int64_t wv = (int64_t)rh * rw * rd;
    const int sh = starts[win * 3], sw = starts[win * 3 + 1], sd = starts[win * 3 + 2];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < wv; i += (int64_t)gridDim.x * blockDim.x) {
        int d = (int)(i % rd);
        int64_t t = i / rd;
        int w = (int)(t % rw);
        int h = (int)(t / rw);
        out[(int64_t)win * wv + i] = vol[((int64_t)(sh + h) * W + (sw + w)) * D + (sd + d)];
    }
}

static inline unsigned grid_for(int64_t items, int per_block, int waves) {
    int64_t b = ceil_div64(items, per_block);
    int64_t cap = (int64_t)sm_count() * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace ltu

using namespace ltu;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
#define LTU_DTYPE_CHECK(name) LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, name ": bad dtype %d", dtype)

extern "C" int ltu_s2d_input(const float* x, void* y, int B, int H, int W, int D, int cpad, int dtype,
                             ltu_stream_t stream) {
    LTU_ARG_CHECK(cpad == 4 || cpad == 8, "s2d_input: cpad must be 4 or 8");
    LTU_ARG_CHECK(x && y, "s2d_input: null pointer");
    LTU_DTYPE_CHECK("s2d_input");
    LTU_ARG_CHECK(B > 0 && H > 0 && W > 0 && D > 0 && H % 2 == 0 && W % 2 == 0, "s2d_input: H and W must be even");
    LTU_ARG_CHECK(aligned16(y), "s2d_input: output must be 16-byte aligned");
    int64_t total = (int64_t)B * (H / 2) * (W / 2) * D;
    unsigned g = grid_for(total, 256, 16);
    if (dtype == LTU_F32) s2d_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>(x, (float*)y, B, H, W, D, cpad);
    else s2d_kernel<bf16><<<g, 256, 0, (cudaStream_t)stream>>>(x, (bf16*)y, B, H, W, D, cpad);
    LTU_LAUNCH_CHECK("s2d_input");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_upsample_trilinear(const void* x, void* y, int B, int H, int W, int D, int C, int fd, int dtype,
                                      ltu_stream_t stream) {
    LTU_ARG_CHECK(x && y, "upsample_trilinear: null pointer");
    LTU_DTYPE_CHECK("upsample_trilinear");
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(B > 0 && H > 0 && W > 0 && D > 0 && C % vn == 0 && (fd == 1 || fd == 2), "upsample_trilinear: bad shape");
    LTU_ARG_CHECK(aligned16(x) && aligned16(y), "upsample_trilinear: pointers must be 16-byte aligned");
    const int nvec = fd * D * (C / vn), tiles_h = (2 * H + 3) / 4;
    LTU_ARG_CHECK((int64_t)B * tiles_h <= 65535 && (2 * W + 3) / 4 <= 65535, "upsample_trilinear: grid too large");
    const dim3 g((unsigned)((nvec + 63) / 64), (unsigned)((2 * W + 3) / 4), (unsigned)(B * tiles_h));
    if (dtype == LTU_F32) upsample_kernel<float><<<g, 256, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, H, W, D, C, fd, tiles_h);
    else upsample_kernel<bf16><<<g, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, H, W, D, C, fd, tiles_h);
    LTU_LAUNCH_CHECK("upsample_trilinear");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_mask_softmax(const float* logits, float* mask, float* fg, int B, int64_t voxels, int Cout,
                                ltu_stream_t stream) {
    LTU_ARG_CHECK(logits && fg && B > 0 && B <= 65535 && voxels > 0, "mask_softmax: bad arguments");
    LTU_ARG_CHECK(Cout >= 1 && Cout <= 8, "mask_softmax: Cout must be in [1,8] (got %d)", Cout);
    int64_t bx = ceil_div64(voxels, 256);
    int64_t cap = ceil_div64((int64_t)sm_count() * 16, B);
    if (bx > cap) bx = cap;
    mask_softmax_kernel<<<dim3((unsigned)bx, B), 256, 0, (cudaStream_t)stream>>>(logits, mask, fg, voxels, Cout);
    LTU_LAUNCH_CHECK("mask_softmax");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_gate_fused(const void* a, const float* stats_a, const void* g, const float* stats_g,
                              const float* psi_w, const float* psi_b, const void* skip, void* out, int B,
                              int64_t voxels, int Ci, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(a && stats_a && g && stats_g && psi_w && psi_b && skip && out, "gate_fused: null pointer");
    LTU_DTYPE_CHECK("gate_fused");
    const int vn = dtype == LTU_F32 ? 4 : 8;
    const int G = Ci / vn;
    LTU_ARG_CHECK(Ci % vn == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0,
                  "gate_fused: Ci=%d unsupported (Ci/%d must be a power of two <= 32)", Ci, vn);
    LTU_ARG_CHECK(B > 0 && B <= 65535 && voxels > 0, "gate_fused: bad shape");
    LTU_ARG_CHECK(aligned16(a) && aligned16(g) && aligned16(skip) && aligned16(out), "gate_fused: pointers must be 16-byte aligned");
    int64_t total = voxels * G;
    int64_t bx = ceil_div64(total, 256);
    int64_t cap = ceil_div64((int64_t)sm_count() * 16, B);
    if (bx > cap) bx = cap;
    dim3 grid((unsigned)bx, B);
    if (dtype == LTU_F32) gate_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)a, stats_a, (const float*)g, stats_g, psi_w, psi_b, (const float*)skip, (float*)out, voxels, Ci);
    else gate_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)a, stats_a, (const bf16*)g, stats_g, psi_w, psi_b, (const bf16*)skip, (bf16*)out, voxels, Ci);
    LTU_LAUNCH_CHECK("gate_fused");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_head_d2s_softmax(const float* logits, float* probs, float* onehot, uint8_t* labels, int B, int H2,
                                    int W2, int D, int Cout, ltu_stream_t stream) {
    LTU_ARG_CHECK(logits && (probs || onehot || labels), "head_d2s_softmax: nothing to do");
    LTU_ARG_CHECK(B > 0 && H2 > 0 && W2 > 0 && D > 0, "head_d2s_softmax: bad shape");
    LTU_ARG_CHECK(Cout >= 1 && Cout <= 8, "head_d2s_softmax: dim_output must be in [1,8] (got %d)", Cout);
    LTU_ARG_CHECK(aligned16(logits), "head_d2s_softmax: logits must be 16-byte aligned");
    int64_t total = (int64_t)B * H2 * W2 * D;
    unsigned g = grid_for(total, 256, 16);
    cudaStream_t st = (cudaStream_t)stream;
    switch (Cout) {
#define HEAD_CASE(n) case n: head_kernel<n><<<g, 256, 0, st>>>(logits, probs, onehot, labels, B, H2, W2, D); break;
        HEAD_CASE(1) HEAD_CASE(2) HEAD_CASE(3) HEAD_CASE(4) HEAD_CASE(5) HEAD_CASE(6) HEAD_CASE(7) HEAD_CASE(8)
#undef HEAD_CASE
    }
    LTU_LAUNCH_CHECK("head_d2s_softmax");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_vote_accumulate(const uint8_t* labels, const int32_t* starts, uint8_t* votes, int nwin, int rh,
                                   int rw, int rd, int C, int H, int W, int D, ltu_stream_t stream) {
    LTU_ARG_CHECK(labels && starts && votes, "vote_accumulate: null pointer");
    LTU_ARG_CHECK(nwin > 0 && nwin <= 65535 && rh > 0 && rw > 0 && rd > 0 && C > 0 && rh <= H && rw <= W && rd <= D,
                  "vote_accumulate: bad shape");
    LTU_ARG_CHECK((reinterpret_cast<uintptr_t>(votes) & 3) == 0 && ((int64_t)H * W * D) % 4 == 0,
                  "vote_accumulate: votes must be 4-byte aligned with H*W*D a multiple of 4");
    int64_t wv = (int64_t)rh * rw * rd;
    int64_t bx = ceil_div64(wv, 256);
    int64_t cap = ceil_div64((int64_t)sm_count() * 16, nwin);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    vote_accumulate_kernel<<<dim3((unsigned)bx, nwin), 256, 0, (cudaStream_t)stream>>>(labels, starts, votes, rh, rw, rd, C, H, W, D);
    LTU_LAUNCH_CHECK("vote_accumulate");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_vote_argmax(const uint8_t* votes, uint8_t* labels, int C, int64_t voxels, ltu_stream_t stream) {
    LTU_ARG_CHECK(votes && labels && C > 0 && voxels > 0, "vote_argmax: bad arguments");
    unsigned g = grid_for(voxels, 256, 16);
    vote_argmax_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(votes, labels, C, voxels);
    LTU_LAUNCH_CHECK("vote_argmax");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_vote_fractions(const uint8_t* votes, float* frac, int C, int64_t voxels, ltu_stream_t stream) {
    LTU_ARG_CHECK(votes && frac && C > 0 && voxels > 0, "vote_fractions: bad arguments");
    unsigned g = grid_for(voxels, 256, 16);
    vote_fractions_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(votes, frac, C, voxels);
    LTU_LAUNCH_CHECK("vote_fractions");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_gather_windows(const float* volume, const int32_t* starts, float* out, int nwin, int rh, int rw,
                                  int rd, int H, int W, int D, ltu_stream_t stream) {
    LTU_ARG_CHECK(volume && starts && out, "gather_windows: null pointer");
    LTU_ARG_CHECK(nwin > 0 && nwin <= 65535 && rh > 0 && rw > 0 && rd > 0 && rh <= H && rw <= W && rd <= D,
                  "gather_windows: bad shape");
    int64_t wv = (int64_t)rh * rw * rd;
    int64_t bx = ceil_div64(wv, 256);
    int64_t cap = ceil_div64((int64_t)sm_count() * 16, nwin);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    gather_windows_kernel<<<dim3((unsigned)bx, nwin), 256, 0, (cudaStream_t)stream>>>(volume, starts, out, rh, rw, rd, H, W, D);
    LTU_LAUNCH_CHECK("gather_windows");
    count_launch(1);
    return LTU_OK;
}

namespace ltu {
// out[r][0 : Ca) = a[r], out[r][Ca : Ca + Cb) = b[r]   (16-byte vectors; Ca, Cb multiples of 8 bf16 / 4 fp32 elements)
__global__ void __launch_bounds__(256)
concat2_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, uint4* __restrict__ out, int64_t rows, int va, int vb) {
    pdl_prologue();
    const int vo = va + vb;
    const int64_t total = rows * vo;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vo;
        const int c = (int)(i - r * vo);
        out[i] = c < va ? __ldg(a + r * va + c) : __ldg(b + r * vb + (c - va));
    }
}
}  // namespace ltu

// torch.cat([a, b], dim=channel) of two channels-last tensors with the same rows (model/Unet_3Dblock.py:553): lets a layer
// whose two inputs have 32 channels each run as ONE 64-channel input of the TMA-halo tcgen05 kernel (ltu_conv3d_tc3)
extern "C" int ltu_concat2(const void* a, int Ca, const void* b, int Cb, void* out, int64_t rows, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(a && b && out && rows > 0, "concat2: bad arguments");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "concat2: bad dtype %d", dtype);
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(Ca > 0 && Cb > 0 && Ca % vn == 0 && Cb % vn == 0, "concat2: channel counts must be multiples of %d", vn);
    LTU_ARG_CHECK((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0, "concat2: pointers must be 16-byte aligned");
    const int64_t total = rows * ((Ca + Cb) / vn);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaError_t e = launch_pdl(concat2_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, (const uint4*)a, (const uint4*)b,
                               (uint4*)out, rows, Ca / vn, Cb / vn);
    if (e != cudaSuccess) { set_error("concat2: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    count_launch(1);
    return LTU_OK;
}
