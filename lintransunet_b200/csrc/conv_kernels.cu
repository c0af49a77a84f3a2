// Family 2 (exact-fp32 path): direct implicit-GEMM 3-D convolution on CUDA cores with an
// InstanceNorm-statistics epilogue, plus the InstanceNorm finalize/apply passes.
// Replaces nn.Conv3d + nn.InstanceNorm3d + LeakyReLU of model/Unet_3Dblock.py (:310-320,
// :375-382, :422-429, :523-535, :588-594, gates :200-214, heads :1328,:1353).
// The bf16 tensor-core path of the same op is conv_tc.cu; both share the partials layout.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

struct ConvParams {
    const void* in0; const void* in1;
    int C0, C1, Cin;
    int B, Hi, Wi, Di, up2;
    int ks, sh, sw, sd, pad;
    const float* weight; const float* bias;
    int Cout;
    void* out; int out_f32;
    int Ho, Wo, Do;
    float* partials; int tiles;
};

constexpr int kBK = 16;          // input channels per k-step
constexpr int kAPad = 1;         // As row padding (floats)

template <typename T, int BM, int BN>
__global__ void __launch_bounds__(256)
conv3d_kernel(const ConvParams p) {
    constexpr int TX = BN / 4, TY = BM / 4;
    static_assert(TX * TY == 256, "tile must map onto 256 threads");
    constexpr int AV = BM * kBK / 4 / 256;          // 4-channel vectors of A per thread per k-step
    constexpr int LDA = kBK + kAPad;
    __shared__ __align__(16) float As[2][BM * LDA];
    __shared__ __align__(16) float Bs[2][kBK * BN];

    const int tid = threadIdx.x;
    const int tx = tid % TX, ty = tid / TX;
    const int b = blockIdx.y;
    const int64_t Vo = (int64_t)p.Ho * p.Wo * p.Do;
    const int64_t vox0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.z * BN;
    const T* in0 = reinterpret_cast<const T*>(p.in0);
    const T* in1 = reinterpret_cast<const T*>(p.in1);
    const int64_t in_sample = (int64_t)p.Hi * p.Wi * p.Di;

    // the voxels this thread gathers for the A tile
    int a_h[AV], a_w[AV], a_d[AV], a_m[AV], a_cg[AV];
    bool a_ok[AV];
#pragma unroll
    for (int r = 0; r < AV; ++r) {
        int v = tid + r * 256;
        int m = v >> 2;
        a_m[r] = m;
        a_cg[r] = (v & 3) * 4;
        int64_t id = vox0 + m;
        a_ok[r] = id < Vo;
        if (!a_ok[r]) id = 0;
        int dd = (int)(id % p.Do);
        int64_t t = id / p.Do;
        int ww = (int)(t % p.Wo);
        int hh = (int)(t / p.Wo);
        a_h[r] = hh * p.sh - p.pad;
        a_w[r] = ww * p.sw - p.pad;
        a_d[r] = dd * p.sd - p.pad;
    }
    const int He = p.up2 ? 2 * p.Hi : p.Hi, We = p.up2 ? 2 * p.Wi : p.Wi, De = p.up2 ? 2 * p.Di : p.Di;
    const int taps = p.ks * p.ks * p.ks;
    const int csteps = (p.Cin + kBK - 1) / kBK;
    const int nsteps = taps * csteps;
    const bool w_vec = (p.Cout % 4) == 0;

    float ra[AV][4];
    float rb[4];
    const int b_k = tid / (BN / 4), b_n = (tid % (BN / 4)) * 4;   // B tile: thread -> (k row, 4 cols)
    const bool b_active = tid < kBK * (BN / 4);

    auto load_step = [&](int step) {
        int tap = step / csteps, c0 = (step - tap * csteps) * kBK;
        int kh = tap / (p.ks * p.ks), kw = (tap / p.ks) % p.ks, kd = tap % p.ks;
#pragma unroll
        for (int r = 0; r < AV; ++r) {
            int hv = a_h[r] + kh, wv = a_w[r] + kw, dv = a_d[r] + kd;
            int c = c0 + a_cg[r];
            bool ok = a_ok[r] && hv >= 0 && hv < He && wv >= 0 && wv < We && dv >= 0 && dv < De && c < p.Cin;
            if (ok) {
                if (p.up2) { hv >>= 1; wv >>= 1; dv >>= 1; }
                int64_t vox = (int64_t)b * in_sample + ((int64_t)hv * p.Wi + wv) * p.Di + dv;
                if (c < p.C0) load4(in0 + vox * p.C0 + c, ra[r]);
                else          load4(in1 + vox * p.C1 + (c - p.C0), ra[r]);
            } else {
                ra[r][0] = ra[r][1] = ra[r][2] = ra[r][3] = 0.f;
            }
        }
        if (b_active) {
            int c = c0 + b_k;
            const float* wrow = p.weight + ((int64_t)tap * p.Cin + c) * p.Cout + n0 + b_n;
            if (c < p.Cin && w_vec && n0 + b_n + 3 < p.Cout) {
                load4(wrow, rb);
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) rb[j] = (c < p.Cin && n0 + b_n + j < p.Cout) ? wrow[j] : 0.f;
            }
        }
    };
    auto store_step = [&](int buf) {
#pragma unroll
        for (int r = 0; r < AV; ++r) {
            float* dst = &As[buf][a_m[r] * LDA + a_cg[r]];
            dst[0] = ra[r][0]; dst[1] = ra[r][1]; dst[2] = ra[r][2]; dst[3] = ra[r][3];
        }
        if (b_active) *reinterpret_cast<float4*>(&Bs[buf][b_k * BN + b_n]) = make_float4(rb[0], rb[1], rb[2], rb[3]);
    };

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    load_step(0);
    store_step(0);
    __syncthreads();
    for (int step = 0; step < nsteps; ++step) {
        const int buf = step & 1;
        if (step + 1 < nsteps) load_step(step + 1);
        const float* a = &As[buf][ty * 4 * LDA];
        const float* bb = &Bs[buf][tx * 4];
#pragma unroll
        for (int k = 0; k < kBK; ++k) {
            float4 bv = *reinterpret_cast<const float4*>(bb + k * BN);
            float av[4] = {a[k], a[LDA + k], a[2 * LDA + k], a[3 * LDA + k]};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[i][0] = fmaf(av[i], bv.x, acc[i][0]);
                acc[i][1] = fmaf(av[i], bv.y, acc[i][1]);
                acc[i][2] = fmaf(av[i], bv.z, acc[i][2]);
                acc[i][3] = fmaf(av[i], bv.w, acc[i][3]);
            }
        }
        if (step + 1 < nsteps) store_step(buf ^ 1);
        __syncthreads();
    }

    // ---- epilogue: bias, store, per-tile channel statistics
    float bs[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int n = n0 + tx * 4 + j;
        bs[j] = (p.bias != nullptr && n < p.Cout) ? p.bias[n] : 0.f;
    }
    float csum[4] = {0.f, 0.f, 0.f, 0.f}, csq[4] = {0.f, 0.f, 0.f, 0.f};
    const int nbase = n0 + tx * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t id = vox0 + ty * 4 + i;
        if (id >= Vo) continue;
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            o[j] = acc[i][j] + bs[j];
            // statistics are taken from the value AS STORED (rounded to the storage type), like
            // instance_norm on the conv output: a (near-)constant channel then normalises to ~0
            // instead of amplifying its own rounding error by rstd
            if (!p.out_f32) o[j] = to_f32(from_f32<T>(o[j]));
            csum[j] += o[j];
            csq[j] = fmaf(o[j], o[j], csq[j]);
        }
        int64_t off = ((int64_t)b * Vo + id) * p.Cout + nbase;
        if (w_vec && nbase + 3 < p.Cout) {
            if (p.out_f32) store4(reinterpret_cast<float*>(p.out) + off, o);
            else           store4(reinterpret_cast<T*>(p.out) + off, o);
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (nbase + j < p.Cout) {
                    if (p.out_f32) reinterpret_cast<float*>(p.out)[off + j] = o[j];
                    else           reinterpret_cast<T*>(p.out)[off + j] = from_f32<T>(o[j]);
                }
        }
    }
    if (p.partials != nullptr) {
        // deterministic column reduction over the TY voxel-groups through smem (reuses As)
        static_assert(2 * BM * LDA >= 2 * TY * BN, "stats staging must fit in the A tiles");
        float* red = &As[0][0];                     // needs 2*TY*BN floats
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            red[(ty * BN + tx * 4 + j) * 2 + 0] = csum[j];
            red[(ty * BN + tx * 4 + j) * 2 + 1] = csq[j];
        }
        __syncthreads();
        if (tid < BN && n0 + tid < p.Cout) {
            float s = 0.f, q = 0.f;
            for (int r = 0; r < TY; ++r) { s += red[(r * BN + tid) * 2]; q += red[(r * BN + tid) * 2 + 1]; }
            float* dst = p.partials + (((int64_t)b * p.tiles + blockIdx.x) * p.Cout + n0 + tid) * 2;
            dst[0] = s; dst[1] = q;
        }
    }
}

// grid (ceil(C/32), B), 1024 threads: lane <-> channel, the 32 warps stride over the tiles (coalesced
// float2 rows of the partials), fp64 accumulation, fixed-order cross-warp merge => deterministic
__global__ void __launch_bounds__(1024)
instnorm_finalize_kernel(const float* __restrict__ partials, float* __restrict__ stats, int B, int tiles, int C,
                         double inv_vox, float eps) {
    __shared__ double red[32][32][2];
    pdl_prologue();          // launched with programmatic stream serialisation: scheduled under the producing conv's tail
    const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 32 + lane;
    double s = 0.0, q = 0.0;
    if (c < C) {
        const float2* base = reinterpret_cast<const float2*>(partials) + (int64_t)b * tiles * C + c;
        // eight independent loads in flight per thread (the loop is a chain of L2 round trips otherwise)
        double s8[8] = {0, 0, 0, 0, 0, 0, 0, 0}, q8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int t = warp;
        for (; t + 7 * 32 < tiles; t += 8 * 32) {
            float2 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = base[(int64_t)(t + 32 * u) * C];
#pragma unroll
            for (int u = 0; u < 8; ++u) { s8[u] += (double)v[u].x; q8[u] += (double)v[u].y; }
        }
        for (; t < tiles; t += 32) {
            const float2 v = base[(int64_t)t * C];
            s8[0] += (double)v.x;
            q8[0] += (double)v.y;
        }
        s = ((s8[0] + s8[1]) + (s8[2] + s8[3])) + ((s8[4] + s8[5]) + (s8[6] + s8[7]));
        q = ((q8[0] + q8[1]) + (q8[2] + q8[3])) + ((q8[4] + q8[5]) + (q8[6] + q8[7]));
    }
    red[warp][lane][0] = s;
    red[warp][lane][1] = q;
    __syncthreads();
    if (warp == 0 && c < C) {
        s = 0.0; q = 0.0;
        for (int w = 0; w < 32; ++w) { s += red[w][lane][0]; q += red[w][lane][1]; }
        const double mean = s * inv_vox;
        double var = q * inv_vox - mean * mean;
        if (var < 0.0) var = 0.0;
        stats[((int64_t)b * C + c) * 2] = (float)mean;
        stats[((int64_t)b * C + c) * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
}

// grid (chunks, B): each CTA reduces a contiguous voxel range of a channels-last tensor
template <typename T>
__global__ void __launch_bounds__(256)
chan_partials_kernel(const T* __restrict__ x, float* __restrict__ partials, int64_t V, int C, int chunks) {
    __shared__ float red[2 * 256 * 4];
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int cg = C / 4;                       // channel groups of 4
    const int rows = 256 / cg;                  // voxels processed per sweep (C <= 1024, C%4==0)
    const int my_cg = threadIdx.x % cg, my_row = threadIdx.x / cg;
    int64_t per = ceil_div64(V, chunks);
    int64_t v0 = (int64_t)chunk * per, v1 = v0 + per < V ? v0 + per : V;
    float s[4] = {0, 0, 0, 0}, q[4] = {0, 0, 0, 0};
    if (my_row < rows) {
        for (int64_t v = v0 + my_row; v < v1; v += rows) {
            float r[4];
            load4(x + ((int64_t)b * V + v) * C + my_cg * 4, r);
#pragma unroll
            for (int i = 0; i < 4; ++i) { s[i] += r[i]; q[i] = fmaf(r[i], r[i], q[i]); }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { red[(threadIdx.x * 4 + i) * 2] = s[i]; red[(threadIdx.x * 4 + i) * 2 + 1] = q[i]; }
    __syncthreads();
    if (threadIdx.x < C) {
        int c = threadIdx.x, g = c / 4, i = c % 4;
        float ss = 0.f, qq = 0.f;
        for (int r = 0; r < rows; ++r) {
            int t = r * cg + g;
            ss += red[(t * 4 + i) * 2];
            qq += red[(t * 4 + i) * 2 + 1];
        }
        float* dst = partials + (((int64_t)b * chunks + chunk) * C + c) * 2;
        dst[0] = ss; dst[1] = qq;
    }
}

// The grid stride (gridDim.x * 256 vectors) is a multiple of the vectors per voxel (C / VN in {1,2,4,...,32}), so a thread
// always handles the same channel group: (mean, rstd) live in registers and the loop body is two 16-byte loads, the
// arithmetic and one 16-byte store, two independent iterations in flight.
template <typename T>
__global__ void __launch_bounds__(256)
instnorm_apply_kernel(const T* __restrict__ x, const float* __restrict__ stats, const T* __restrict__ res,
                      T* __restrict__ y, int64_t V, int C, int act) {
    constexpr int VN = Vec<T>::N;
    pdl_prologue();          // the statistics come from instnorm_finalize, which is still running when this grid is scheduled
    const int b = blockIdx.y;
    const int cv = C / VN;
    const int64_t total = V * cv;
    const float* st = stats + (int64_t)b * C * 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool fixed = (stride % cv) == 0;                    // true for every launch of the wrapper (256 % cv == 0)
    float mean[VN], rstd[VN];
    {
        const int c0 = (int)(first % cv) * VN;
#pragma unroll
        for (int i = 0; i < VN; ++i) { mean[i] = __ldg(st + (c0 + i) * 2); rstd[i] = __ldg(st + (c0 + i) * 2 + 1); }
    }
    const T* xb = x + (int64_t)b * V * C;
    const T* rb = res != nullptr ? res + (int64_t)b * V * C : nullptr;
    T* yb = y + (int64_t)b * V * C;
    auto body = [&](int64_t idx, const float (&v_in)[VN], const float (&r_in)[VN]) {
        float v[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) {
            float t = (v_in[i] - mean[i]) * rstd[i];
            if (act == LTU_ACT_LRELU) t = t > 0.f ? t : 0.01f * t;
            v[i] = t + r_in[i];
        }
        store_vec(yb + idx * VN, v);
    };
    int64_t idx = first;
    if (fixed) {
        for (; idx + stride < total; idx += 2 * stride) {
            float v0[VN], v1[VN], r0[VN], r1[VN];
            load_vec(xb + idx * VN, v0);
            load_vec(xb + (idx + stride) * VN, v1);
            if (rb != nullptr) { load_vec(rb + idx * VN, r0); load_vec(rb + (idx + stride) * VN, r1); }
            else {
#pragma unroll
                for (int i = 0; i < VN; ++i) { r0[i] = 0.f; r1[i] = 0.f; }
            }
            body(idx, v0, r0);
            body(idx + stride, v1, r1);
        }
    }
    for (; idx < total; idx += stride) {
        if (!fixed) {
            const int c0 = (int)(idx % cv) * VN;
#pragma unroll
            for (int i = 0; i < VN; ++i) { mean[i] = __ldg(st + (c0 + i) * 2); rstd[i] = __ldg(st + (c0 + i) * 2 + 1); }
        }
        float v0[VN], r0[VN];
        load_vec(xb + idx * VN, v0);
        if (rb != nullptr) load_vec(rb + idx * VN, r0);
        else {
#pragma unroll
            for (int i = 0; i < VN; ++i) r0[i] = 0.f;
        }
        body(idx, v0, r0);
    }
}

template <typename T>
static int conv_launch(const ConvParams& p, cudaStream_t st) {
    if (p.Cout >= 64) {
        dim3 grid((unsigned)p.tiles, p.B, (p.Cout + 63) / 64);
        conv3d_kernel<T, 64, 64><<<grid, 256, 0, st>>>(p);
    } else if (p.Cout >= 32) {
        dim3 grid((unsigned)p.tiles, p.B, (p.Cout + 31) / 32);
        conv3d_kernel<T, 128, 32><<<grid, 256, 0, st>>>(p);
    } else {
        dim3 grid((unsigned)p.tiles, p.B, (p.Cout + 15) / 16);
        conv3d_kernel<T, 256, 16><<<grid, 256, 0, st>>>(p);
    }
    LTU_LAUNCH_CHECK("conv3d");
    count_launch(1);
    return LTU_OK;
}

}  // namespace ltu

using namespace ltu;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int ltu_conv3d_tiles(int64_t out_voxels, int Cout) {
    int bm = Cout >= 64 ? 64 : (Cout >= 32 ? 128 : 256);
    return (int)ceil_div64(out_voxels, bm);
}

extern "C" int ltu_conv3d(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di, int up2,
                          int ksize, int sh, int sw, int sd, int pad, const float* weight, const float* bias,
                          int Cout, void* out, int out_f32, int Ho, int Wo, int Do, float* partials, int dtype,
                          ltu_stream_t stream) {
    LTU_ARG_CHECK(in0 && weight && out, "conv3d: null pointer");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "conv3d: bad dtype %d", dtype);
    LTU_ARG_CHECK(ksize == 1 || ksize == 3, "conv3d: kernel size must be 1 or 3 (got %d)", ksize);
    LTU_ARG_CHECK(pad == (ksize == 3 ? 1 : 0) || pad == 0, "conv3d: pad must be 0 or ksize/2");
    LTU_ARG_CHECK(B > 0 && B <= 65535 && Hi > 0 && Wi > 0 && Di > 0 && Cout > 0 && C0 > 0, "conv3d: bad shape");
    LTU_ARG_CHECK(C0 % 4 == 0 && C1 % 4 == 0 && (in1 != nullptr) == (C1 > 0), "conv3d: channels must be multiples of 4");
    LTU_ARG_CHECK(C1 == 0 || C0 % kBK == 0, "conv3d: with a second input C0 must be a multiple of %d", kBK);
    LTU_ARG_CHECK(sh >= 1 && sw >= 1 && sd >= 1 && sh <= 2 && sw <= 2 && sd <= 2, "conv3d: stride must be 1 or 2");
    LTU_ARG_CHECK(!up2 || (sh == 1 && sw == 1 && sd == 1 && ksize == 3 && pad == 1), "conv3d: up2 needs a stride-1 3x3x3 conv");
    const int He = up2 ? 2 * Hi : Hi, We = up2 ? 2 * Wi : Wi, De = up2 ? 2 * Di : Di;
    LTU_ARG_CHECK(Ho == (He + 2 * pad - ksize) / sh + 1 && Wo == (We + 2 * pad - ksize) / sw + 1 &&
                  Do == (De + 2 * pad - ksize) / sd + 1, "conv3d: output size does not match the geometry");
    LTU_ARG_CHECK(aligned16(in0) && aligned16(in1) && aligned16(weight) && aligned16(out), "conv3d: pointers must be 16-byte aligned");
    ConvParams p;
    p.in0 = in0; p.in1 = in1; p.C0 = C0; p.C1 = C1; p.Cin = C0 + C1;
    p.B = B; p.Hi = Hi; p.Wi = Wi; p.Di = Di; p.up2 = up2;
    p.ks = ksize; p.sh = sh; p.sw = sw; p.sd = sd; p.pad = pad;
    p.weight = weight; p.bias = bias; p.Cout = Cout; p.out = out; p.out_f32 = out_f32;
    p.Ho = Ho; p.Wo = Wo; p.Do = Do; p.partials = partials;
    p.tiles = ltu_conv3d_tiles((int64_t)Ho * Wo * Do, Cout);
    if (dtype == LTU_F32) return conv_launch<float>(p, (cudaStream_t)stream);
    return conv_launch<bf16>(p, (cudaStream_t)stream);
}

extern "C" int ltu_instnorm_finalize(const float* partials, float* stats, int B, int tiles, int C, int64_t voxels,
                                     float eps, ltu_stream_t stream) {
    LTU_ARG_CHECK(partials && stats && B > 0 && tiles > 0 && C > 0 && voxels > 0, "instnorm_finalize: bad arguments");
    LTU_ARG_CHECK(B <= 65535, "instnorm_finalize: B too large");
    cudaError_t le = launch_pdl(instnorm_finalize_kernel, dim3((C + 31) / 32, B), dim3(1024), 0, (cudaStream_t)stream,
                                partials, stats, B, tiles, C, 1.0 / (double)voxels, eps);
    if (le != cudaSuccess) { set_error("instnorm_finalize: launch failed: %s", cudaGetErrorString(le)); return (int)le; }
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_chan_partials(const void* x, float* partials, int B, int64_t voxels, int C, int chunks, int dtype,
                                 ltu_stream_t stream) {
    LTU_ARG_CHECK(x && partials && B > 0 && B <= 65535 && voxels > 0 && chunks > 0, "chan_partials: bad arguments");
    LTU_ARG_CHECK(C % 4 == 0 && C >= 4 && C <= 256, "chan_partials: C must be a multiple of 4 in [4,256] (got %d)", C);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "chan_partials: bad dtype %d", dtype);
    dim3 grid(chunks, B);
    if (dtype == LTU_F32) chan_partials_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)x, partials, voxels, C, chunks);
    else chan_partials_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, partials, voxels, C, chunks);
    LTU_LAUNCH_CHECK("chan_partials");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_instnorm_apply(const void* x, const float* stats, const void* residual, void* y, int B,
                                  int64_t voxels, int C, int act, int dtype, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && stats && y && B > 0 && B <= 65535 && voxels > 0, "instnorm_apply: bad arguments");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "instnorm_apply: bad dtype %d", dtype);
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(C % vn == 0, "instnorm_apply: C=%d must be a multiple of %d", C, vn);
    LTU_ARG_CHECK(act == LTU_ACT_NONE || act == LTU_ACT_LRELU, "instnorm_apply: bad act %d", act);
    LTU_ARG_CHECK(aligned16(x) && aligned16(y) && aligned16(residual), "instnorm_apply: pointers must be 16-byte aligned");
    int64_t total = voxels * (C / vn);
    int64_t bx = ceil_div64(total, 256);
    int64_t cap = ceil_div64((int64_t)sm_count() * 16, B);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    dim3 grid((unsigned)bx, B);
    cudaError_t le;
    if (dtype == LTU_F32) le = launch_pdl(instnorm_apply_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream, (const float*)x, stats, (const float*)residual, (float*)y, voxels, C, act);
    else le = launch_pdl(instnorm_apply_kernel<bf16>, grid, dim3(256), 0, (cudaStream_t)stream, (const bf16*)x, stats, (const bf16*)residual, (bf16*)y, voxels, C, act);
    if (le != cudaSuccess) { set_error("instnorm_apply: launch failed: %s", cudaGetErrorString(le)); return (int)le; }
    count_launch(1);
    return LTU_OK;
}
