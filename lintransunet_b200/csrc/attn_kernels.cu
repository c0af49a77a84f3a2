// Family 1: the linear ("efficient") attention core and the bandwidth-bound glue of
// SelfAttentionLayer (reference model/trans_block.py:41-67, :203-211, :86-96).
//
// Layout: tokens are rows of a [B, N, C] matrix, C = heads*32, heads are the contiguous 32-wide
// column slices (trans_block.py:156 view(B,N,h,32)).
//
// kv_reduce  : one pass over K and V.  Every warp keeps, for ONE head, the online column-softmax
//              state (running max m[j], running sum s[j]) and the 32x32 context in registers
//              (lane e owns column e).  Partial states are merged by kv_combine in a fixed order
//              => bit-reproducible.
// q_readout  : one pass over Q, writes out.
#include <stdlib.h>

#include "common.cuh"

namespace ltu {

constexpr int kHeadDim = 32;
constexpr int kTileTokens = 32;      // tokens staged per cp.async stage
constexpr int kAttnThreads = 256;    // 8 warps
constexpr int kAttnWarps = 8;
constexpr int kPartialFloats = 32 * 32 + 64;   // ctx[32][32], m[32], s[32]

// How a sample's N tokens are split into chunks (one CTA each, one partial state per chunk).  The split
// depends on N ONLY -- never on the batch size -- so that a sample's result is bit-identical whatever
// batch it is processed in (the N-GPU sliding window must equal the 1-GPU one exactly): up to 37 chunks
// per sample (37 * 8 CTAs = one wave of 2 CTAs x 148 SMs at the benchmark's batch of 8; measured: 74 chunks
// cost 25 % more through extra partial states and pipeline fills), at least 4 tiles (128 tokens) each.
static inline int kv_chunks_per_batch(int /*B*/, int64_t N) {
    const int64_t tiles = ceil_div64(N, kTileTokens);
    int64_t tiles_per_chunk = ceil_div64(tiles, 37);
    if (tiles_per_chunk < 4) tiles_per_chunk = 4;
    return (int)ceil_div64(tiles, tiles_per_chunk);
}

// Stage `kTileTokens` rows x C columns of `src` (row stride ld) into smem with 16-byte cp.async.
template <typename T>
__device__ __forceinline__ void stage_tile(T* smem, const T* src, int64_t ld, int64_t row0,
                                           int64_t nrows_total, int C) {
    constexpr int VN = Vec<T>::N;
    const int chunks_per_row = C / VN;
    const int total = kTileTokens * chunks_per_row;
    for (int i = threadIdx.x; i < total; i += kAttnThreads) {
        int r = i / chunks_per_row, c = i - r * chunks_per_row;
        int64_t row = row0 + r;
        bool ok = row < nrows_total;
        const T* g = src + (ok ? row : row0) * ld + c * VN;
        cp_async16_zfill(smem + r * C + c * VN, g, ok ? 16 : 0);
    }
}

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
kv_reduce_kernel(const T* __restrict__ K, const T* __restrict__ V, int64_t ld, float* __restrict__ part,
                 int64_t N, int heads, int chunks, int tiles_per_chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = heads * kHeadDim;
    const int wph = kAttnWarps / heads;              // warps per head
    const int ttw = kTileTokens / wph;               // tokens per warp per tile
    T* sK = reinterpret_cast<T*>(smem_raw);          // [2][TT][C]
    T* sV = sK + 2 * kTileTokens * C;                // [2][TT][C]
    float* sP = reinterpret_cast<float*>(sV + 2 * kTileTokens * C);   // [8 warps][ttw][32]

    const int b = blockIdx.y, chunk = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = warp % heads, sub = warp / heads;
    const T* Kb = K + (int64_t)b * N * ld;
    const T* Vb = V + (int64_t)b * N * ld;
    float* myP = sP + warp * ttw * kHeadDim;

    const int64_t tile0 = (int64_t)chunk * tiles_per_chunk;
    int64_t ntiles = ceil_div64(N, kTileTokens) - tile0;
    if (ntiles > tiles_per_chunk) ntiles = tiles_per_chunk;
    if (ntiles < 0) ntiles = 0;

    float acc[kHeadDim];
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) acc[j] = 0.f;
    float m_run = -INFINITY, s_run = 0.f;

    if (ntiles > 0) {
        stage_tile(sK, Kb, ld, tile0 * kTileTokens, N, C);
        stage_tile(sV, Vb, ld, tile0 * kTileTokens, N, C);
    }
    cp_async_commit();
    for (int64_t t = 0; t < ntiles; ++t) {
        const int buf = (int)(t & 1);
        if (t + 1 < ntiles) {
            stage_tile(sK + (buf ^ 1) * kTileTokens * C, Kb, ld, (tile0 + t + 1) * kTileTokens, N, C);
            stage_tile(sV + (buf ^ 1) * kTileTokens * C, Vb, ld, (tile0 + t + 1) * kTileTokens, N, C);
        }
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const T* tk = sK + buf * kTileTokens * C + hd * kHeadDim;
        const T* tv = sV + buf * kTileTokens * C + hd * kHeadDim;
        int64_t row0 = (tile0 + t) * kTileTokens;
        int valid = (int)((N - row0) < kTileTokens ? (N - row0) : kTileTokens);
        // tokens of this warp inside the tile: sub*ttw .. sub*ttw+ttw-1
        int n_lo = sub * ttw;
        int n_cnt = valid - n_lo;
        if (n_cnt > ttw) n_cnt = ttw;
        if (n_cnt > 0) {
            // (A) column max over the warp's tokens, lane j <-> key feature j
            float mt = -INFINITY;
            for (int n = 0; n < n_cnt; ++n) mt = fmaxf(mt, to_f32(tk[(n_lo + n) * C + lane]));
            float m_new = fmaxf(m_run, mt);
            if (__any_sync(0xffffffffu, m_new > m_run)) {
                float sc = (m_run == -INFINITY) ? 0.f : __expf(m_run - m_new);
                s_run *= sc;
#pragma unroll
                for (int j = 0; j < kHeadDim; ++j) acc[j] *= __shfl_sync(0xffffffffu, sc, j);
                m_run = m_new;
            }
            // (B) p = exp(k - m) into smem, running sum
            for (int n = 0; n < n_cnt; ++n) {
                float p = __expf(to_f32(tk[(n_lo + n) * C + lane]) - m_run);
                myP[n * kHeadDim + lane] = p;
                s_run += p;
            }
            __syncwarp();
            // (C) ctx[j][e] += p[n][j] * v[n][e], lane e <-> value feature e
            for (int n = 0; n < n_cnt; ++n) {
                float v = to_f32(tv[(n_lo + n) * C + lane]);
                const float4* pr = reinterpret_cast<const float4*>(myP + n * kHeadDim);
#pragma unroll
                for (int q = 0; q < kHeadDim / 4; ++q) {
                    float4 p4 = pr[q];
                    acc[4 * q + 0] = fmaf(p4.x, v, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(p4.y, v, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(p4.z, v, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(p4.w, v, acc[4 * q + 3]);
                }
            }
        }
        __syncthreads();   // tile buffer `buf` is re-filled by the next iteration's prefetch
    }
    cp_async_wait<0>();

    // partial index: ((b*chunks + chunk)*wph + sub)*heads + hd
    float* out = part + ((((int64_t)b * chunks + chunk) * wph + sub) * heads + hd) * kPartialFloats;
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) out[j * kHeadDim + lane] = acc[j];
    out[1024 + lane] = m_run;
    out[1056 + lane] = s_run;
}

// grid (heads, B), 1024 threads: warp j, lane e own ctx[j][e].  Merges the partial states in a fixed order.
// The merge is latency bound (every part is one dependent L2 round trip per thread), so the lanes of a warp
// split the parts for the running-max / key-sum columns (exact max; butterfly sums) and the 32 weighted
// accumulator loads of a block of parts are all issued before the first one is consumed.
__global__ void __launch_bounds__(1024)
kv_combine_kernel(const float* __restrict__ part, float* __restrict__ ctx, int heads, int nparts,
                  const bf16* __restrict__ wo, bf16* __restrict__ wout) {
    const int hd = blockIdx.x, b = blockIdx.y;
    const int j = threadIdx.x >> 5, e = threadIdx.x & 31;
    // optional tail (wo != null): the output projection's weight with this head's context folded in,
    //   wout[b][n][32 hd + j'] = sum_e' ctx[b][hd][j'][e'] wo[n][32 hd + e'],   n < C = 32 heads
    // (ltu_linear_fused_ex with per-sample weights).  wo is a parameter: its loads do not wait for the previous kernel.
    const int C = heads * kHeadDim;
    __shared__ float cs[32][33];                     // cs[e'][j'] = ctx[j'][e']
    __shared__ __align__(16) float ws_[256][32];     // ws_[n][e'] = wo[n][32 hd + e']
    if (wo != nullptr) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            if (i < heads) ws_[j + 32 * i][e] = __bfloat162float(wo[(int64_t)(j + 32 * i) * C + 32 * hd + e]);
    }
    pdl_prologue();
    const float* base = part + ((int64_t)b * nparts * heads + hd) * kPartialFloats;
    const int64_t stride = (int64_t)heads * kPartialFloats;
    float M = -INFINITY;
    for (int p = e; p < nparts; p += 32) M = fmaxf(M, base[p * stride + 1024 + j]);
    M = warp_max(M);
    float S = 0.f;
    float A8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int p0 = 0; p0 < nparts; p0 += 32) {
        const int p = p0 + e;
        const float mp = p < nparts ? base[p * stride + 1024 + j] : -INFINITY;
        const float sp = p < nparts ? base[p * stride + 1056 + j] : 0.f;
        const float wl = (mp == -INFINITY) ? 0.f : __expf(mp - M);
        S += warp_sum(sp * wl);
        const int cnt = nparts - p0 < 32 ? nparts - p0 : 32;
        float a[32];
#pragma unroll
        for (int u = 0; u < 32; ++u) a[u] = u < cnt ? base[(p0 + u) * stride + j * kHeadDim + e] : 0.f;
#pragma unroll
        for (int u = 0; u < 32; ++u) A8[u & 7] = fmaf(a[u], __shfl_sync(0xffffffffu, wl, u), A8[u & 7]);
    }
    const float A = ((A8[0] + A8[1]) + (A8[2] + A8[3])) + ((A8[4] + A8[5]) + (A8[6] + A8[7]));
    const float cv = A / S;
    ctx[(((int64_t)b * heads + hd) * kHeadDim + j) * kHeadDim + e] = cv;
    if (wo == nullptr) return;
    cs[e][j] = cv;
    __syncthreads();
    float c[32];                                     // lane e holds row j' = e of the context: c[k] = ctx[e][k]
#pragma unroll
    for (int k = 0; k < 32; ++k) c[k] = cs[k][e];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        if (i < heads) {                             // thread (warp j, lane e): row n = j + 32 i, column 32 hd + e
            const float4* w4 = reinterpret_cast<const float4*>(ws_[j + 32 * i]);      // warp-uniform: broadcast reads
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float4 w = w4[k];
                acc = fmaf(w.x, c[4 * k], acc); acc = fmaf(w.y, c[4 * k + 1], acc);
                acc = fmaf(w.z, c[4 * k + 2], acc); acc = fmaf(w.w, c[4 * k + 3], acc);
            }
            wout[((int64_t)b * C + j + 32 * i) * C + 32 * hd + e] = __float2bfloat16_rn(acc);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(kAttnThreads)
q_readout_kernel(const T* __restrict__ Q, int64_t ldq, const float* __restrict__ ctx, T* __restrict__ O,
                 int64_t ldo, int64_t N, int heads, int tiles_per_cta) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int C = heads * kHeadDim;
    const int wph = kAttnWarps / heads;
    const int ttw = kTileTokens / wph;
    T* sQ = reinterpret_cast<T*>(smem_raw);                                 // [2][TT][C]
    float* sP = reinterpret_cast<float*>(sQ + 2 * kTileTokens * C);         // [8][ttw][32]

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int hd = warp % heads, sub = warp / heads;
    const T* Qb = Q + (int64_t)b * N * ldq;
    T* Ob = O + (int64_t)b * N * ldo;
    float* myP = sP + warp * ttw * kHeadDim;

    // ctx column e of this head, pre-scaled by 1/sqrt(d_k) (trans_block.py:50)
    float c[kHeadDim];
    const float* cb = ctx + ((int64_t)b * heads + hd) * kHeadDim * kHeadDim;
    const float inv_sqrt_d = 0.17677669529663687f;   // 1/sqrt(32)
#pragma unroll
    for (int j = 0; j < kHeadDim; ++j) c[j] = cb[j * kHeadDim + lane] * inv_sqrt_d;

    const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
    int64_t ntiles = ceil_div64(N, kTileTokens) - tile0;
    if (ntiles > tiles_per_cta) ntiles = tiles_per_cta;
    if (ntiles < 0) ntiles = 0;

    if (ntiles > 0) stage_tile(sQ, Qb, ldq, tile0 * kTileTokens, N, C);
    cp_async_commit();
    for (int64_t t = 0; t < ntiles; ++t) {
        const int buf = (int)(t & 1);
        if (t + 1 < ntiles)
            stage_tile(sQ + (buf ^ 1) * kTileTokens * C, Qb, ldq, (tile0 + t + 1) * kTileTokens, N, C);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();

        const T* tq = sQ + buf * kTileTokens * C + hd * kHeadDim;
        int64_t row0 = (tile0 + t) * kTileTokens;
        int valid = (int)((N - row0) < kTileTokens ? (N - row0) : kTileTokens);
        int n_lo = sub * ttw;
        int n_cnt = valid - n_lo;
        if (n_cnt > ttw) n_cnt = ttw;
        if (n_cnt > 0) {
            for (int n = 0; n < n_cnt; ++n) {
                float qv = to_f32(tq[(n_lo + n) * C + lane]);
                float mx = warp_max(qv);
                myP[n * kHeadDim + lane] = __expf(qv - mx);
            }
            __syncwarp();
            for (int n = 0; n < n_cnt; ++n) {
                const float4* pr = reinterpret_cast<const float4*>(myP + n * kHeadDim);
                float o = 0.f, den = 0.f;
#pragma unroll
                for (int q = 0; q < kHeadDim / 4; ++q) {
                    float4 p4 = pr[q];
                    o = fmaf(p4.x, c[4 * q + 0], o);
                    o = fmaf(p4.y, c[4 * q + 1], o);
                    o = fmaf(p4.z, c[4 * q + 2], o);
                    o = fmaf(p4.w, c[4 * q + 3], o);
                    den += (p4.x + p4.y) + (p4.z + p4.w);
                }
                Ob[(row0 + n_lo + n) * ldo + hd * kHeadDim + lane] = from_f32<T>(o / den);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
}

// ---------------------------------------------------------------- add + LayerNorm
// Split-bf16 token stream (bf16 path): the residual stream between encoder layers is kept as hi + lo, two bf16 words
// (hi = bf16(y), lo = bf16(y - hi)): the GEMMs read `hi` only -- exactly the bf16 cast autocast applies to a Linear's
// input -- while the residual add sees 16 significant bits, like the reference, whose LayerNorm output stays fp32 under
// autocast.  `xlo` / `ylo` may be null (plain bf16 stream).
template <typename T> __device__ __forceinline__ void split_store4(T* yhi, T* ylo, const float (&o)[4]) { store4(yhi, o); }
template <> __device__ __forceinline__ void split_store4<bf16>(bf16* yhi, bf16* ylo, const float (&o)[4]) {
    float h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { h[i] = __bfloat162float(__float2bfloat16_rn(o[i])); l[i] = o[i] - h[i]; }
    store4(yhi, h);
    if (ylo) store4(ylo, l);
}

template <typename T, int C>
__global__ void __launch_bounds__(256)
add_layernorm_kernel(const T* __restrict__ x, const T* __restrict__ xlo, const T* __restrict__ r,
                     const float* __restrict__ gamma, const float* __restrict__ beta, T* __restrict__ y,
                     T* __restrict__ ylo, int64_t rows, float eps) {
    constexpr int Q = C / 128;   // 4-element chunks per lane
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    float g[Q][4], bt[Q][4];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        load4(gamma + q * 128 + lane * 4, g[q]);
        load4(beta + q * 128 + lane * 4, bt[q]);
    }
    for (int64_t row = warp_global; row < rows; row += nwarps) {
        float v[Q][4];
        float sum = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float a[4], bb[4];
            load4(x + row * C + q * 128 + lane * 4, a);
            load4(r + row * C + q * 128 + lane * 4, bb);
            if (xlo) {
                float lo[4];
                load4(xlo + row * C + q * 128 + lane * 4, lo);
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] += lo[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { v[q][i] = a[i] + bb[i]; sum += v[q][i]; }
        }
        float mean = warp_sum(sum) * (1.f / C);
        float sq = 0.f;
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) { float d = v[q][i] - mean; sq = fmaf(d, d, sq); }
        float rstd = rsqrtf(warp_sum(sq) * (1.f / C) + eps);
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            float o[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) o[i] = (v[q][i] - mean) * rstd * g[q][i] + bt[q][i];
            split_store4<T>(y + row * C + q * 128 + lane * 4, ylo ? ylo + row * C + q * 128 + lane * 4 : nullptr, o);
        }
    }
}

template <typename T> __device__ __forceinline__ float gelu_of(float v);
template <> __device__ __forceinline__ float gelu_of<float>(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
template <> __device__ __forceinline__ float gelu_of<bf16>(float v) { return gelu_erf(v); }   // relative error 8e-5 (common.cuh), then rounded to bf16

template <typename T>
__global__ void __launch_bounds__(256) gelu_kernel(T* __restrict__ x, int64_t nvec) {
    constexpr int VN = Vec<T>::N;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + stride < nvec; i += 2 * stride) {              // two independent 16-byte vectors in flight
        float a[VN], b[VN];
        load_vec(x + i * VN, a);
        load_vec(x + (i + stride) * VN, b);
#pragma unroll
        for (int k = 0; k < VN; ++k) { a[k] = gelu_of<T>(a[k]); b[k] = gelu_of<T>(b[k]); }
        store_vec(x + i * VN, a);
        store_vec(x + (i + stride) * VN, b);
    }
    for (; i < nvec; i += stride) {
        float v[VN];
        load_vec(x + i * VN, v);
#pragma unroll
        for (int k = 0; k < VN; ++k) v[k] = gelu_of<T>(v[k]);
        store_vec(x + i * VN, v);
    }
}

// ---------------------------------------------------------------- depthwise positional conv
template <typename T>
__global__ void __launch_bounds__(256)
posenc_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
              T* __restrict__ y, int B, int H, int W, int D, int C) {
    const int cv = C / 4;
    const int64_t total = (int64_t)B * H * W * D * cv;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        int c4 = (int)(idx % cv) * 4;
        int64_t vox = idx / cv;
        int d = (int)(vox % D);
        int64_t t = vox / D;
        int ww = (int)(t % W);
        t /= W;
        int h = (int)(t % H);
        int b = (int)(t / H);
        float acc[4], ctr[4];
        load4(x + vox * C + c4, ctr);
        float bs[4];
        load4(bias + c4, bs);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = bs[i];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            int hh = h + kh - 1;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                int w2 = ww + kw - 1;
                if (w2 < 0 || w2 >= W) continue;
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    int dd = d + kd - 1;
                    if (dd < 0 || dd >= D) continue;
                    float xv[4], wv[4];
                    int64_t nv = (((int64_t)b * H + hh) * W + w2) * D + dd;
                    load4(x + nv * C + c4, xv);
                    load4(w + (kh * 9 + kw * 3 + kd) * C + c4, wv);
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[i] = fmaf(xv[i], wv[i], acc[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += ctr[i];
        store4(y + vox * C + c4, acc);
    }
}

// v2: one thread = one (b,h,w) column segment of LD outputs along D x 2 channels.  The 27x2 weights live in
// registers; every input value is loaded once per (kh,kw) neighbour and scattered to the three outputs it
// feeds (kd = 0,1,2), so an output costs ~10 loads instead of 27 + 27 weight loads.
__device__ __forceinline__ void load2(const float* p, float& a, float& b) { float2 v = *reinterpret_cast<const float2*>(p); a = v.x; b = v.y; }
__device__ __forceinline__ void load2(const bf16* p, float& a, float& b) {
    uint32_t v = *reinterpret_cast<const uint32_t*>(p);
    a = __uint_as_float(v << 16); b = __uint_as_float(v & 0xffff0000u);
}
__device__ __forceinline__ void store2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
__device__ __forceinline__ void store2(bf16* p, float a, float b) { *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(a, b); }

template <typename T, int LD>
__global__ void __launch_bounds__(256)
posenc2_kernel(const T* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               T* __restrict__ y, int B, int H, int W, int D, int C) {
    const int cp = C / 2, nd = (D + LD - 1) / LD;
    const int64_t total = (int64_t)B * H * W * nd * cp;
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int c0 = (int)(idx % cp) * 2;
    int64_t t = idx / cp;
    const int dc = (int)(t % nd); t /= nd;
    const int ww = (int)(t % W); t /= W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    float wr[27][2];
#pragma unroll
    for (int tp = 0; tp < 27; ++tp) load2(w + tp * C + c0, wr[tp][0], wr[tp][1]);
    float b0, b1;
    load2(bias + c0, b0, b1);
    const int d0 = dc * LD;
    const int dend = (d0 + LD < D) ? d0 + LD : D;                 // outputs d0 .. dend-1
    const T* xb = x + (int64_t)b * H * W * D * C + c0;
    T* yb = y + (int64_t)b * H * W * D * C + c0;
    float aP0 = 0.f, aP1 = 0.f, aC0 = 0.f, aC1 = 0.f, aN0 = 0.f, aN1 = 0.f;
    for (int dd = d0 - 1; dd <= dend; ++dd) {
        if (dd >= 0 && dd < D) {
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int hh = h + kh - 1;
                if (hh < 0 || hh >= H) continue;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw) {
                    const int w2 = ww + kw - 1;
                    if (w2 < 0 || w2 >= W) continue;
                    float v0, v1;
                    load2(xb + (((int64_t)hh * W + w2) * D + dd) * C, v0, v1);
                    const int t0 = kh * 9 + kw * 3;
                    aN0 = fmaf(wr[t0][0], v0, aN0);     aN1 = fmaf(wr[t0][1], v1, aN1);        // kd = 0 -> out dd+1
                    aC0 = fmaf(wr[t0 + 1][0], v0, aC0); aC1 = fmaf(wr[t0 + 1][1], v1, aC1);    // kd = 1 -> out dd
                    aP0 = fmaf(wr[t0 + 2][0], v0, aP0); aP1 = fmaf(wr[t0 + 2][1], v1, aP1);    // kd = 2 -> out dd-1
                }
            }
        }
        const int o = dd - 1;                                      // complete: planes o-1, o, o+1 were scattered
        if (o >= d0 && o < dend) {
            const int64_t off = (((int64_t)h * W + ww) * D + o) * C;
            float x0, x1;
            load2(xb + off, x0, x1);
            store2(yb + off, aP0 + b0 + x0, aP1 + b1 + x1);
        }
        aP0 = aC0; aP1 = aC1; aC0 = aN0; aC1 = aN1; aN0 = 0.f; aN1 = 0.f;
    }
}

void count_launch(int n = 1);

// bf16 tensor-pipe variants (attn_tc.cu)
int kv_reduce_bf16_mma(const void* k, const void* v, int64_t ld, float* ctx, void* ws, int B, int64_t N, int heads,
                       cudaStream_t st);
int q_readout_bf16_mma(const void* q, int64_t ldq, const float* ctx, void* out, int64_t ldo, int B, int64_t N,
                       int heads, cudaStream_t st);
int kv_chunks_per_batch_host(int B, int64_t N) { return kv_chunks_per_batch(B, N); }
int kv_combine_launch(const float* ws, float* ctx, int heads, int B, int nparts, cudaStream_t st, const void* wo, void* wout) {
    cudaError_t e = launch_pdl(kv_combine_kernel, dim3(heads, B), dim3(1024), 0, st, ws, ctx, heads, nparts, (const bf16*)wo, (bf16*)wout);
    if (e != cudaSuccess) { set_error("kv_combine: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return LTU_OK;
}
// LTU_ATTN_STREAM=0 falls back to the first-generation cp.async kernels of attn_tc.cu (A/B switch, read once)
static bool use_stream_attention() {
    static const bool v = [] { const char* e = getenv("LTU_ATTN_STREAM"); return !(e && e[0] == '0'); }();
    return v;
}
int kv_reduce_bf16_stream(const void* k, const void* v, int64_t ld, float* ctx, void* ws, int B, int64_t N, int heads,
                          cudaStream_t st, const void* wo = nullptr, void* wout = nullptr);                      // attn_stream.cu
int q_readout_bf16_stream(const void* q, int64_t ldq, const float* ctx, void* out, int64_t ldo, int B, int64_t N, int heads,
                          cudaStream_t st);
static bool use_mma_attention() {
    static const bool v = [] { const char* e = getenv("LTU_ATTN_FMA"); return !(e && e[0] == '1'); }();
    return v;
}

template <typename T>
static int kv_reduce_impl(const void* k, const void* v, int64_t ld, float* ctx, void* ws, size_t ws_bytes,
                          int B, int64_t N, int heads, cudaStream_t st) {
    const int C = heads * kHeadDim;
    const int wph = kAttnWarps / heads;
    const int chunks = kv_chunks_per_batch(B, N);
    const int64_t tiles = ceil_div64(N, kTileTokens);
    const int tiles_per_chunk = (int)ceil_div64(tiles, chunks);
    const size_t need = (size_t)B * chunks * wph * heads * kPartialFloats * sizeof(float);
    LTU_ARG_CHECK(ws_bytes >= need, "kv_reduce: workspace too small (%zu < %zu)", ws_bytes, need);
    size_t smem = (size_t)4 * kTileTokens * C * sizeof(T) + (size_t)kAttnWarps * (kTileTokens / wph) * kHeadDim * 4;
    static thread_local int configured_dev_f32 = -1, configured_dev_bf16 = -1;
    int dev;
    cudaGetDevice(&dev);
    int& conf = sizeof(T) == 4 ? configured_dev_f32 : configured_dev_bf16;
    if (conf != dev) {
        cudaFuncSetAttribute(kv_reduce_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        conf = dev;
    }
    kv_reduce_kernel<T><<<dim3(chunks, B), kAttnThreads, smem, st>>>(
        (const T*)k, (const T*)v, ld, (float*)ws, N, heads, chunks, tiles_per_chunk);
    LTU_LAUNCH_CHECK("kv_reduce");
    kv_combine_kernel<<<dim3(heads, B), 1024, 0, st>>>((const float*)ws, ctx, heads, chunks * wph, nullptr, nullptr);
    LTU_LAUNCH_CHECK("kv_combine");
    count_launch(2);
    return LTU_OK;
}

template <typename T>
static int q_readout_impl(const void* q, int64_t ldq, const float* ctx, void* out, int64_t ldo, int B,
                          int64_t N, int heads, cudaStream_t st) {
    const int C = heads * kHeadDim;
    const int wph = kAttnWarps / heads;
    const int64_t tiles = ceil_div64(N, kTileTokens);
    int64_t want = ceil_div64(4 * (int64_t)sm_count(), B);
    int64_t ctas = tiles < want ? tiles : want;
    if (ctas < 1) ctas = 1;
    int tiles_per_cta = (int)ceil_div64(tiles, ctas);
    ctas = ceil_div64(tiles, tiles_per_cta);
    size_t smem = (size_t)2 * kTileTokens * C * sizeof(T) + (size_t)kAttnWarps * (kTileTokens / wph) * kHeadDim * 4;
    static thread_local int configured_dev_f32 = -1, configured_dev_bf16 = -1;
    int dev;
    cudaGetDevice(&dev);
    int& conf = sizeof(T) == 4 ? configured_dev_f32 : configured_dev_bf16;
    if (conf != dev) {
        cudaFuncSetAttribute(q_readout_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        conf = dev;
    }
    q_readout_kernel<T><<<dim3((unsigned)ctas, B), kAttnThreads, smem, st>>>(
        (const T*)q, ldq, ctx, (T*)out, ldo, N, heads, tiles_per_cta);
    LTU_LAUNCH_CHECK("q_readout");
    count_launch(1);
    return LTU_OK;
}

}  // namespace ltu

using namespace ltu;

static bool heads_ok(int heads) { return heads == 1 || heads == 2 || heads == 4 || heads == 8; }
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" size_t ltu_kv_reduce_workspace(int B, int64_t N, int heads) {
    if (B <= 0 || N <= 0 || !heads_ok(heads)) return 0;
    return (size_t)B * kv_chunks_per_batch(B, N) * (kAttnWarps / heads) * heads * kPartialFloats * sizeof(float);
}

extern "C" int ltu_kv_reduce(const void* k, const void* v, int64_t ld, float* ctx, void* ws, size_t ws_bytes,
                             int B, int64_t N, int heads, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(k && v && ctx && ws, "kv_reduce: null pointer");
    LTU_ARG_CHECK(B > 0 && N > 0 && B <= 65535, "kv_reduce: bad B=%d N=%lld", B, (long long)N);
    LTU_ARG_CHECK(heads_ok(heads), "kv_reduce: heads must be 1,2,4 or 8 (got %d)", heads);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "kv_reduce: bad dtype %d", dtype);
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(ld >= heads * 32 && ld % vn == 0, "kv_reduce: row stride %lld not a multiple of %d", (long long)ld, vn);
    LTU_ARG_CHECK(aligned16(k) && aligned16(v) && aligned16(ws), "kv_reduce: pointers must be 16-byte aligned");
    if (dtype == LTU_F32) return kv_reduce_impl<float>(k, v, ld, ctx, ws, ws_bytes, B, N, heads, (cudaStream_t)stream);
    if (heads >= 2 && use_mma_attention()) {
        LTU_ARG_CHECK(ws_bytes >= ltu_kv_reduce_workspace(B, N, heads), "kv_reduce: workspace too small");
        if ((heads == 4 || heads == 8) && ld % 8 == 0 && N < ((int64_t)1 << 31) && use_stream_attention())
            return kv_reduce_bf16_stream(k, v, ld, ctx, ws, B, N, heads, (cudaStream_t)stream);
        return kv_reduce_bf16_mma(k, v, ld, ctx, ws, B, N, heads, (cudaStream_t)stream);
    }
    return kv_reduce_impl<bf16>(k, v, ld, ctx, ws, ws_bytes, B, N, heads, (cudaStream_t)stream);
}

// ltu_kv_reduce (bf16, 4 or 8 heads) whose merge kernel also writes W_b = blockdiag(ctx_b) Wo^T (see ltu_ctx_project)
extern "C" int ltu_kv_reduce_project(const void* k, const void* v, int64_t ld, float* ctx, void* ws, size_t ws_bytes, int B,
                                     int64_t N, int heads, const void* wo_bf16, void* w_out, ltu_stream_t stream) {
    LTU_ARG_CHECK(k && v && ctx && ws && wo_bf16 && w_out, "kv_reduce_project: null pointer");
    LTU_ARG_CHECK(B > 0 && N > 0 && B <= 65535 && N < ((int64_t)1 << 31), "kv_reduce_project: bad B=%d N=%lld", B, (long long)N);
    LTU_ARG_CHECK(heads == 4 || heads == 8, "kv_reduce_project: heads must be 4 or 8 (got %d)", heads);
    LTU_ARG_CHECK(ld >= heads * 32 && ld % 8 == 0, "kv_reduce_project: row stride %lld not a multiple of 8", (long long)ld);
    LTU_ARG_CHECK(aligned16(k) && aligned16(v) && aligned16(ws), "kv_reduce_project: pointers must be 16-byte aligned");
    LTU_ARG_CHECK(ws_bytes >= ltu_kv_reduce_workspace(B, N, heads), "kv_reduce_project: workspace too small");
    return kv_reduce_bf16_stream(k, v, ld, ctx, ws, B, N, heads, (cudaStream_t)stream, wo_bf16, w_out);
}

extern "C" int ltu_q_readout(const void* q, int64_t ldq, const float* ctx, void* out, int64_t ldo, int B,
                             int64_t N, int heads, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(q && ctx && out, "q_readout: null pointer");
    LTU_ARG_CHECK(B > 0 && N > 0 && B <= 65535, "q_readout: bad B=%d N=%lld", B, (long long)N);
    LTU_ARG_CHECK(heads_ok(heads), "q_readout: heads must be 1,2,4 or 8 (got %d)", heads);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "q_readout: bad dtype %d", dtype);
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(ldq >= heads * 32 && ldq % vn == 0 && ldo >= heads * 32, "q_readout: bad row strides");
    LTU_ARG_CHECK(aligned16(q), "q_readout: q must be 16-byte aligned");
    if (dtype == LTU_F32) return q_readout_impl<float>(q, ldq, ctx, out, ldo, B, N, heads, (cudaStream_t)stream);
    if ((heads == 4 || heads == 8) && ldq % 8 == 0 && ldo % 8 == 0 && aligned16(out) && N < ((int64_t)1 << 31) &&
        use_mma_attention() && use_stream_attention())
        return q_readout_bf16_stream(q, ldq, ctx, out, ldo, B, N, heads, (cudaStream_t)stream);
    if (heads >= 2 && ldo % 8 == 0 && aligned16(out) && use_mma_attention())
        return q_readout_bf16_mma(q, ldq, ctx, out, ldo, B, N, heads, (cudaStream_t)stream);
    return q_readout_impl<bf16>(q, ldq, ctx, out, ldo, B, N, heads, (cudaStream_t)stream);
}

static int add_layernorm_impl(const void* x, const void* xlo, const void* res, const float* gamma, const float* beta,
                              void* y, void* ylo, int64_t rows, int C, float eps, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && res && gamma && beta && y, "add_layernorm: null pointer");
    LTU_ARG_CHECK(rows > 0, "add_layernorm: rows=%lld", (long long)rows);
    LTU_ARG_CHECK(C == 128 || C == 256, "add_layernorm: C must be 128 or 256 (got %d)", C);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "add_layernorm: bad dtype %d", dtype);
    LTU_ARG_CHECK(dtype == LTU_BF16 || (!xlo && !ylo), "add_layernorm: the split stream exists for bf16 only");
    LTU_ARG_CHECK(aligned16(x) && aligned16(res) && aligned16(y) && aligned16(gamma) && aligned16(beta) &&
                  aligned16(xlo) && aligned16(ylo), "add_layernorm: pointers must be 16-byte aligned");
    int64_t blocks = ceil_div64(rows, 8);
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = (cudaStream_t)stream;
#define LN_LAUNCH(T, CC) add_layernorm_kernel<T, CC><<<(unsigned)blocks, 256, 0, st>>>((const T*)x, (const T*)xlo, (const T*)res, gamma, beta, (T*)y, (T*)ylo, rows, eps)
    if (dtype == LTU_F32) { if (C == 128) LN_LAUNCH(float, 128); else LN_LAUNCH(float, 256); }
    else                  { if (C == 128) LN_LAUNCH(bf16, 128);  else LN_LAUNCH(bf16, 256); }
#undef LN_LAUNCH
    LTU_LAUNCH_CHECK("add_layernorm");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_add_layernorm(const void* x, const void* res, const float* gamma, const float* beta, void* y,
                                 int64_t rows, int C, float eps, int dtype, ltu_stream_t stream) {
    return add_layernorm_impl(x, nullptr, res, gamma, beta, y, nullptr, rows, C, eps, dtype, stream);
}

extern "C" int ltu_add_layernorm_split(const void* x_hi, const void* x_lo, const void* res, const float* gamma,
                                       const float* beta, void* y_hi, void* y_lo, int64_t rows, int C, float eps,
                                       ltu_stream_t stream) {
    LTU_ARG_CHECK(y_lo, "add_layernorm_split: y_lo is null");
    return add_layernorm_impl(x_hi, x_lo, res, gamma, beta, y_hi, y_lo, rows, C, eps, LTU_BF16, stream);
}

extern "C" int ltu_gelu(void* x, int64_t n, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && n > 0, "gelu: bad arguments");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "gelu: bad dtype %d", dtype);
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(n % vn == 0 && aligned16(x), "gelu: n must be a multiple of %d and x 16-byte aligned", vn);
    int64_t nvec = n / vn;
    int64_t blocks = ceil_div64(nvec, 256);
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (dtype == LTU_F32) gelu_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((float*)x, nvec);
    else gelu_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)x, nvec);
    LTU_LAUNCH_CHECK("gelu");
    count_launch(1);
    return LTU_OK;
}

// bf16 storage: one thread = 8 channels x 4 consecutive outputs along D.  For each of the 9 (kh,kw) neighbours
// the six input planes d0-1 .. d0+4 are loaded once (16-byte loads) and feed the 3 taps of the 4 outputs, so an
// output costs 13.5 data + 13.5 weight loads per 8 channels instead of 54 + 54 (the plain kernel is LSU-issue
// bound: 219 us on [8,39,23,64,128]); the residual term reuses the centre planes.
__global__ void __launch_bounds__(256)
posenc_bf16_kernel(const bf16* __restrict__ x, const bf16* __restrict__ xlo, const float* __restrict__ w,
                   const float* __restrict__ bias, bf16* __restrict__ y, bf16* __restrict__ ylo, int B, int H, int W,
                   int D, int C) {
    const int c8n = C >> 3;
    const int dgn = (D + 3) >> 2;
    const int64_t total = (int64_t)B * H * W * dgn * c8n;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int c0 = (int)(idx % c8n) * 8;
        int64_t t = idx / c8n;
        const int d0 = (int)(t % dgn) * 4;
        t /= dgn;
        const int ww = (int)(t % W);
        t /= W;
        const int h = (int)(t % H);
        const int b = (int)(t / H);
        float acc[4][8];
        {
            float b8[8];
            load_vec(bias + c0, *reinterpret_cast<float(*)[4]>(b8));
            load_vec(bias + c0 + 4, *reinterpret_cast<float(*)[4]>(b8 + 4));
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[o][i] = b8[i];
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
            const int hh = h + kh - 1;
            if (hh < 0 || hh >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const int w2 = ww + kw - 1;
                if (w2 < 0 || w2 >= W) continue;
                const bf16* col = x + ((((int64_t)b * H + hh) * W + w2) * D) * C + c0;
                uint4 raw[6];
#pragma unroll
                for (int pl = 0; pl < 6; ++pl) {
                    const int dd = d0 + pl - 1;
                    raw[pl] = (dd >= 0 && dd < D) ? *reinterpret_cast<const uint4*>(col + (int64_t)dd * C) : make_uint4(0, 0, 0, 0);
                }
                float wt[3][8];
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const float* wp = w + (kh * 9 + kw * 3 + kd) * C + c0;
                    load_vec(wp, *reinterpret_cast<float(*)[4]>(wt[kd]));
                    load_vec(wp + 4, *reinterpret_cast<float(*)[4]>(wt[kd] + 4));
                }
#pragma unroll
                for (int pl = 0; pl < 6; ++pl) {
                    const uint32_t u[4] = {raw[pl].x, raw[pl].y, raw[pl].z, raw[pl].w};
                    float xv[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        xv[2 * i] = __uint_as_float(u[i] << 16);
                        xv[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
                    }
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        const int o = pl - kd;                  // output d0+o reads plane d0+o+kd-1 = d0+pl-1
                        if (o < 0 || o > 3) continue;
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[o][i] = fmaf(xv[i], wt[kd][i], acc[o][i]);
                    }
                    if (kh == 1 && kw == 1 && pl >= 1 && pl <= 4) {     // residual: x itself
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[pl - 1][i] += xv[i];
                    }
                }
            }
        }
        const int64_t off = ((((int64_t)b * H + h) * W + ww) * D + d0) * C + c0;
        if (xlo) {                                  // split stream: the residual term is hi + lo, the taps read hi
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (d0 + o < D) {
                    float l8[8];
                    load_vec(xlo + off + (int64_t)o * C, l8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[o][i] += l8[i];
                }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (d0 + o < D) {
                if (ylo) {
                    float h8[8], l8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { h8[i] = __bfloat162float(__float2bfloat16_rn(acc[o][i])); l8[i] = acc[o][i] - h8[i]; }
                    store_vec(y + off + (int64_t)o * C, h8);
                    store_vec(ylo + off + (int64_t)o * C, l8);
                } else {
                    store_vec(y + off + (int64_t)o * C, acc[o]);
                }
            }
    }
}

// Round 2: shared-memory tiled version of the same arithmetic for C % 32 == 0.  posenc_bf16_kernel reads the nine
// (kh, kw) neighbour columns of every output column through L1/L2 -- 9x the tensor through L2, 1.0 TB/s of its bytes on
// bridge 1.  Here a block owns a 4 x 4 x 32 (h, w, d) tile of a 32-channel slice: the 6 x 6 x 34 halo is staged once
// (16-byte cp.async, zero fill = the convolution's padding, so the tap loop has no bounds checks) in a
// [channel chunk][voxel] layout whose chunk pitch is 1 mod 8 sixteen-byte units (conflict-free 16-byte LDS for a
// quarter-warp of 4 chunks x 2 depth groups); the 27 x 32 weights and the bias sit in shared memory as well.  2.4x instead of
// 9x re-read, same thread micro-kernel (8 channels x 4 outputs along D, six planes feed three taps each), same fma order.
constexpr int kPeTH = 4, kPeTW = 4, kPeTD = 32, kPeCC = 32;
constexpr int kPeHD = kPeTD + 2, kPeHW = kPeTW + 2, kPeHH = kPeTH + 2;
constexpr int kPeNV = kPeHH * kPeHW * kPeHD;                  // 1224 halo voxels
constexpr int kPePitch = kPeNV + 1;                            // 1225 = 1 mod 8
constexpr int kPeSmem = 4 * kPePitch * 16 + 28 * kPeCC * 4;   // halo + weights [27][32] + bias [32]

__global__ void __launch_bounds__(256, 2)
posenc_tile_kernel(const bf16* __restrict__ x, const bf16* __restrict__ xlo, const float* __restrict__ w,
                   const float* __restrict__ bias, bf16* __restrict__ y, bf16* __restrict__ ylo, int H, int W, int D, int C,
                   int tiles_w, int tiles_d, int slices) {
    extern __shared__ __align__(16) unsigned char pe_smem[];
    uint4* sx = reinterpret_cast<uint4*>(pe_smem);                              // [4][kPePitch]
    float* sw = reinterpret_cast<float*>(pe_smem + 4 * kPePitch * 16);          // [27][32], then bias [32]
    const int tid = threadIdx.x;
    int t = blockIdx.x;
    const int cs = t % slices; t /= slices;
    const int td = t % tiles_d; t /= tiles_d;
    const int tw = t % tiles_w;
    const int th = t / tiles_w;
    const int b = blockIdx.y;
    const int h0 = th * kPeTH, w0 = tw * kPeTW, d0t = td * kPeTD, c0s = cs * kPeCC;
    const bf16* xb = x + (int64_t)b * H * W * D * C + c0s;
    for (int i = tid; i < kPeNV * 4; i += 256) {
        const int chunk = i & 3, v = i >> 2;
        const int hd = v % kPeHD, hw = (v / kPeHD) % kPeHW, hh = v / (kPeHD * kPeHW);
        const int gh = h0 - 1 + hh, gw = w0 - 1 + hw, gd = d0t - 1 + hd;
        const bool ok = gh >= 0 && gh < H && gw >= 0 && gw < W && gd >= 0 && gd < D;
        const bf16* src = ok ? xb + (((int64_t)gh * W + gw) * D + gd) * C + chunk * 8 : x;
        cp_async16_zfill(sx + chunk * kPePitch + v, src, ok ? 16 : 0);
    }
    for (int i = tid; i < 27 * kPeCC; i += 256) sw[i] = w[(i >> 5) * C + c0s + (i & 31)];
    if (tid < kPeCC) sw[27 * kPeCC + tid] = bias[c0s + tid];
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const int chunk = tid & 3, dg = (tid >> 2) & 7;
    const uint4* sxc = sx + chunk * kPePitch;
    const float* swc = sw + chunk * 8;
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        const int col = (tid >> 5) + it * 8;                    // 0..15
        const int lh = col >> 2, lw = col & 3;
        const int gh = h0 + lh, gw = w0 + lw, gd0 = d0t + dg * 4;
        if (gh >= H || gw >= W || gd0 >= D) continue;
        float acc[4][8];
        {
            float b8[8];
            load_vec(swc + 27 * kPeCC, *reinterpret_cast<float(*)[4]>(b8));
            load_vec(swc + 27 * kPeCC + 4, *reinterpret_cast<float(*)[4]>(b8 + 4));
#pragma unroll
            for (int o = 0; o < 4; ++o)
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[o][i] = b8[i];
        }
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                const uint4* colp = sxc + ((lh + kh) * kPeHW + (lw + kw)) * kPeHD + dg * 4;    // halo plane d0 - 1
                uint4 raw[6];
#pragma unroll
                for (int pl = 0; pl < 6; ++pl) raw[pl] = colp[pl];
                float wt[3][8];
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const float* wp = swc + (kh * 9 + kw * 3 + kd) * kPeCC;
                    load_vec(wp, *reinterpret_cast<float(*)[4]>(wt[kd]));
                    load_vec(wp + 4, *reinterpret_cast<float(*)[4]>(wt[kd] + 4));
                }
#pragma unroll
                for (int pl = 0; pl < 6; ++pl) {
                    const uint32_t u[4] = {raw[pl].x, raw[pl].y, raw[pl].z, raw[pl].w};
                    float xv[8];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        xv[2 * i] = __uint_as_float(u[i] << 16);
                        xv[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u);
                    }
#pragma unroll
                    for (int kd = 0; kd < 3; ++kd) {
                        const int o = pl - kd;
                        if (o < 0 || o > 3) continue;
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[o][i] = fmaf(xv[i], wt[kd][i], acc[o][i]);
                    }
                    if (kh == 1 && kw == 1 && pl >= 1 && pl <= 4) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[pl - 1][i] += xv[i];
                    }
                }
            }
        const int64_t off = ((((int64_t)b * H + gh) * W + gw) * D + gd0) * C + c0s + chunk * 8;
        if (xlo) {
#pragma unroll
            for (int o = 0; o < 4; ++o)
                if (gd0 + o < D) {
                    float l8[8];
                    load_vec(xlo + off + (int64_t)o * C, l8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[o][i] += l8[i];
                }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
            if (gd0 + o < D) {
                if (ylo) {
                    float h8[8], l8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { h8[i] = __bfloat162float(__float2bfloat16_rn(acc[o][i])); l8[i] = acc[o][i] - h8[i]; }
                    store_vec(y + off + (int64_t)o * C, h8);
                    store_vec(ylo + off + (int64_t)o * C, l8);
                } else {
                    store_vec(y + off + (int64_t)o * C, acc[o]);
                }
            }
    }
}

static int posenc_impl(const void* x, const void* xlo, const float* w, const float* bias, void* y, void* ylo, int B,
                       int H, int W, int D, int C, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && w && bias && y, "posenc_dwconv3: null pointer");
    LTU_ARG_CHECK(B > 0 && H > 0 && W > 0 && D > 0 && C > 0 && C % 4 == 0, "posenc_dwconv3: bad shape");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "posenc_dwconv3: bad dtype %d", dtype);
    LTU_ARG_CHECK(x != y, "posenc_dwconv3: in-place is not supported");
    const bool vec_ok = dtype == LTU_BF16 && C % 8 == 0 &&
                        (((uintptr_t)x | (uintptr_t)y | (uintptr_t)w | (uintptr_t)bias | (uintptr_t)xlo | (uintptr_t)ylo) & 15) == 0;
    LTU_ARG_CHECK(vec_ok || (!xlo && !ylo), "posenc_dwconv3: the split stream needs bf16, C %% 8 == 0 and 16-byte aligned pointers");
    // posenc2_kernel (register-sliding along D) measured 2x SLOWER than the plain 27-tap kernel on B200
    // (404 vs 219 us on [8,39,23,64,128]: a serial chain of dependent loads per thread at 114 registers);
    // kept for reference, not dispatched.
    static const bool tiled = [] { const char* e = getenv("LTU_POSENC_TILED"); return !(e && e[0] == '0'); }();   // A/B switch
    if (vec_ok && tiled && C % kPeCC == 0 && B <= 65535) {
        const int tiles_h = (H + kPeTH - 1) / kPeTH, tiles_w = (W + kPeTW - 1) / kPeTW, tiles_d = (D + kPeTD - 1) / kPeTD;
        const int slices = C / kPeCC;
        const int64_t blocks = (int64_t)tiles_h * tiles_w * tiles_d * slices;
        LTU_ARG_CHECK(blocks < ((int64_t)1 << 31), "posenc_dwconv3: tensor too large");
        static thread_local int conf = -1;
        int dev; cudaGetDevice(&dev);
        if (conf != dev) {
            cudaFuncSetAttribute(posenc_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPeSmem);
            conf = dev;
        }
        posenc_tile_kernel<<<dim3((unsigned)blocks, (unsigned)B), 256, kPeSmem, (cudaStream_t)stream>>>(
            (const bf16*)x, (const bf16*)xlo, w, bias, (bf16*)y, (bf16*)ylo, H, W, D, C, tiles_w, tiles_d, slices);
    } else if (vec_ok) {
        int64_t total = (int64_t)B * H * W * ((D + 3) / 4) * (C / 8);
        int64_t blocks = ceil_div64(total, 256);
        LTU_ARG_CHECK(blocks < ((int64_t)1 << 31), "posenc_dwconv3: tensor too large");
        posenc_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)xlo, w, bias, (bf16*)y,
                                                                              (bf16*)ylo, B, H, W, D, C);
    } else {
        int64_t total = (int64_t)B * H * W * D * (C / 4);
        int64_t blocks = ceil_div64(total, 256);
        int64_t cap = (int64_t)sm_count() * 32;
        if (blocks > cap) blocks = cap;
        if (dtype == LTU_F32) posenc_kernel<float><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const float*)x, w, bias, (float*)y, B, H, W, D, C);
        else posenc_kernel<bf16><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, w, bias, (bf16*)y, B, H, W, D, C);
    }
    LTU_LAUNCH_CHECK("posenc_dwconv3");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_posenc_dwconv3(const void* x, const float* w, const float* bias, void* y, int B, int H, int W,
                                  int D, int C, int dtype, ltu_stream_t stream) {
    return posenc_impl(x, nullptr, w, bias, y, nullptr, B, H, W, D, C, dtype, stream);
}

extern "C" int ltu_posenc_dwconv3_split(const void* x_hi, const void* x_lo, const float* w, const float* bias, void* y_hi,
                                        void* y_lo, int B, int H, int W, int D, int C, ltu_stream_t stream) {
    LTU_ARG_CHECK(y_lo, "posenc_dwconv3_split: y_lo is null");
    return posenc_impl(x_hi, x_lo, w, bias, y_hi, y_lo, B, H, W, D, C, LTU_BF16, stream);
}
