// Dynamic-ROI plumbing of ROIBridge (reference model/Unet_3Dblock.py):
//   roi_bbox     : get_mask_boundary2 (:821-873) + get_min_max_indice (:37-49), fully on device
//                  (the reference does 18 host syncs per sample here);
//   roi_resample : get_transfer_index (:51-64) / get_transfer_back_index (:66-82) feeding a 2-D
//                  bilinear grid_sample(align_corners=True, zeros) per depth slice (:1034, :1112),
//                  restated as a separable resample on channels-last data without a grid tensor.
// All index arithmetic uses explicitly rounded fp32 ops (__f*_rn) in the reference's operation
// order so that coordinates are bit-identical to the PyTorch fp32 evaluation.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

// one warp per (b,h,w) line over d; integer atomics => deterministic
__global__ void __launch_bounds__(256)
roi_profile_kernel(const float* __restrict__ fg, int* __restrict__ prof, int h, int w, int d, float thr) {
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int lines = h * w;
    int* ph = prof + (int64_t)b * (h + w);
    int* pw = ph + h;
    for (int line = blockIdx.x * 8 + (threadIdx.x >> 5); line < lines; line += gridDim.x * 8) {
        const float* src = fg + ((int64_t)b * lines + line) * d;
        int cnt = 0;
        for (int i = lane; i < d; i += 32) cnt += (src[i] >= thr) ? 1 : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        if (lane == 0 && cnt > 0) {
            atomicAdd(ph + line / w, cnt);
            atomicAdd(pw + line % w, cnt);
        }
    }
}

// get_min_max_indice on one profile; returns (lo, hi, mid) as the reference's fp32 values
__device__ void quantiles(const int* prof, int S, float& lo, float& hi, float& mid) {
    long long tot = 0;
    for (int i = 0; i < S; ++i) tot += prof[i];
    if (tot == 0) {
        float m = (float)S / 2.f;
        lo = m - 1.f; hi = m + 1.f; mid = m;
        return;
    }
    const float lo_t = (float)0.001, hi_t = (float)(1.0 - 0.001), md_t = 0.5f;
    const float ftot = (float)tot;
    int ilo = S, ihi = S, imid = S;
    long long cum = 0;
    for (int i = 0; i < S; ++i) {
        cum += prof[i];
        float r = __fdiv_rn((float)cum, ftot);
        if (ilo == S && r >= lo_t) ilo = i;      // searchsorted(right=False)
        if (ihi == S && r > hi_t) ihi = i;       // searchsorted(right=True)
        if (imid == S && r > md_t) imid = i;
    }
    lo = (float)ilo; hi = (float)ihi; mid = (float)imid;
}

__device__ void clamp_axis(float& lo, float& hi, float mid, int S, int mn) {
    const float size = __fsub_rn(hi, lo);          // evaluated once, before either clamp (:847-848)
    if (size < (float)mn) {
        const float half = (float)mn / 2.f;
        lo = fmaxf(__fsub_rn(mid, half), 0.f);
        hi = fminf(__fadd_rn(mid, half), (float)S);
    }
    if (size > (float)(S - mn)) {
        const float half = (float)(S - mn) / 2.f;
        lo = fmaxf(__fsub_rn(mid, half), 0.f);
        hi = fminf(__fadd_rn(mid, half), (float)S);
    }
}

__global__ void roi_box_kernel(const int* __restrict__ prof, float* __restrict__ box, int h, int w, int d,
                               int min_h, int min_w) {
    const int b = blockIdx.x;
    if (threadIdx.x >= 2) return;
    const int* ph = prof + (int64_t)b * (h + w);
    float lo, hi, mid;
    if (threadIdx.x == 0) {
        quantiles(ph, h, lo, hi, mid);
        clamp_axis(lo, hi, mid, h, min_h);
        box[b * 6 + 0] = lo; box[b * 6 + 3] = hi;
        box[b * 6 + 2] = 0.f; box[b * 6 + 5] = (float)(d - 1);
    } else {
        quantiles(ph + h, w, lo, hi, mid);
        clamp_axis(lo, hi, mid, w, min_w);
        box[b * 6 + 1] = lo; box[b * 6 + 4] = hi;
    }
}

// normalised grid coordinate of output sample i (get_transfer_index, :51-64); h = size-1
__device__ __forceinline__ float fisheye_fwd(float x0, float x1, int h, int roi, int eroi, int i) {
    const float k2 = __fdiv_rn(__fsub_rn(x1, x0), (float)(roi - 1));
    const float k1 = __fdiv_rn(__fadd_rn(__fsub_rn((float)h, x1), x0), (float)(eroi - roi));
    float t = __fadd_rn(__fmul_rn((float)i, k2), __fmul_rn(x0, __fsub_rn(1.f, __fdiv_rn(k2, k1))));
    const float r = __fdiv_rn(k1, k2);
    const float om = __fsub_rn(1.f, r);
    if (t <= x0) t = __fadd_rn(__fmul_rn(t, r), __fmul_rn(x0, om));
    if (t >= x1) t = __fadd_rn(__fmul_rn(t, r), __fmul_rn(x1, om));
    return __fsub_rn(__fdiv_rn(__fmul_rn(t, 2.f), (float)h), 1.f);
}
// get_transfer_back_index (:66-82): source pixel p of the original map -> coordinate in the roi
__device__ __forceinline__ float fisheye_back(float x0, float x1, int h, int roi, int eroi, int p) {
    const float k2 = __fdiv_rn((float)roi, __fsub_rn(x1, x0));
    const float k1 = __fdiv_rn((float)(eroi - roi), __fadd_rn(__fsub_rn((float)h, x1), x0));
    const float p0 = __fmul_rn(x0, k1);
    const float p1 = __fsub_rn((float)eroi, __fmul_rn(__fsub_rn((float)h, x1), k1));
    float t = __fadd_rn(__fmul_rn((float)p, k2), __fmul_rn(p0, __fsub_rn(1.f, __fdiv_rn(k2, k1))));
    const float r = __fdiv_rn(k1, k2);
    const float om = __fsub_rn(1.f, r);
    if (t <= p0) t = __fadd_rn(__fmul_rn(t, r), __fmul_rn(p0, om));
    if (t >= p1) t = __fadd_rn(__fmul_rn(t, r), __fmul_rn(p1, om));
    return __fsub_rn(__fdiv_rn(__fmul_rn(t, 2.f), (float)eroi), 1.f);
}

struct Tap { int i0, i1; float w0, w1; };
// grid_sample(align_corners=True, padding_mode='zeros') along one axis
__device__ __forceinline__ Tap make_tap(float coord, int size) {
    float pos = __fmul_rn(__fdiv_rn(__fadd_rn(coord, 1.f), 2.f), (float)(size - 1));
    float f = floorf(pos);
    Tap t;
    t.w1 = pos - f;
    t.w0 = 1.f - t.w1;
    // NaN / huge coordinates: everything out of range => zero contribution
    if (!(f >= -2.f && f <= (float)size)) { t.i0 = t.i1 = 0; t.w0 = t.w1 = 0.f; return t; }
    t.i0 = (int)f;
    t.i1 = t.i0 + 1;
    if (t.i0 < 0 || t.i0 > size - 1) { t.w0 = 0.f; t.i0 = 0; }
    if (t.i1 < 0 || t.i1 > size - 1) { t.w1 = 0.f; t.i1 = 0; }
    return t;
}

// One block = one output row index oi x 4 column indices oj x every (depth, channel) vector of those four columns.  The
// fisheye coordinates and bilinear taps depend on (oi, oj) only -- about twenty IEEE divisions -- and are computed ONCE per
// block (five threads, shared memory) instead of once per 16-byte output vector as in round 1, where the kernel was bound
// by that arithmetic (0.5-1.0 TB/s of its bytes); a warp then streams 512-byte segments of one column.  Same device
// functions, same operands => the taps and the results are bit-identical to the per-element version.
template <typename T>
__global__ void __launch_bounds__(256)
roi_resample_kernel(const T* __restrict__ x, const float* __restrict__ box, T* __restrict__ y, int ih, int iw,
                    int oh, int ow, int d, int C, int h_full, int w_full, int roi_h, int roi_w, int eval_h,
                    int eval_w, int direction) {
    constexpr int VN = Vec<T>::N;
    __shared__ Tap taps[5];                                    // [0] = row tap, [1..4] = column taps
    const int b = blockIdx.z, oi = blockIdx.y, oj0 = (int)blockIdx.x * 4;
    const int cv = C / VN, nvec = d * cv;
    if (threadIdx.x < 5) {
        const float x0 = box[b * 6 + 0], y0 = box[b * 6 + 1], x1 = box[b * 6 + 3], y1 = box[b * 6 + 4];
        if (threadIdx.x == 0) {
            const float ch = direction == 0 ? fisheye_fwd(x0, x1, h_full - 1, roi_h, eval_h, oi)
                                            : fisheye_back(x0, x1, h_full - 1, roi_h, eval_h, oi);
            taps[0] = make_tap(ch, ih);
        } else {
            const int oj = oj0 + (int)threadIdx.x - 1;
            const float cw = direction == 0 ? fisheye_fwd(y0, y1, w_full - 1, roi_w, eval_w, oj)
                                            : fisheye_back(y0, y1, w_full - 1, roi_w, eval_w, oj);
            taps[threadIdx.x] = make_tap(cw, iw);
        }
    }
    __syncthreads();
    const int q = threadIdx.x >> 6, oj = oj0 + q;
    if (oj >= ow) return;
    const Tap th = taps[0], tw = taps[1 + q];
    const T* xb = x + (int64_t)b * ih * iw * d * C;
    T* yb = y + (((int64_t)b * oh + oi) * ow + oj) * (int64_t)nvec * VN;
    const T* src[4];
    float wt[4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int bb = 0; bb < 2; ++bb) {
            wt[a * 2 + bb] = (a ? th.w1 : th.w0) * (bb ? tw.w1 : tw.w0);
            const int hi = a ? th.i1 : th.i0, wi = bb ? tw.i1 : tw.i0;
            src[a * 2 + bb] = xb + ((int64_t)hi * iw + wi) * (int64_t)nvec * VN;
        }
    for (int v = (int)(threadIdx.x & 63); v < nvec; v += 64) {
        float acc[VN];
#pragma unroll
        for (int i = 0; i < VN; ++i) acc[i] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (wt[k] == 0.f) continue;
            float val[VN];
            load_vec(src[k] + (int64_t)v * VN, val);
#pragma unroll
            for (int i = 0; i < VN; ++i) acc[i] = fmaf(wt[k], val[i], acc[i]);
        }
        store_vec(yb + (int64_t)v * VN, acc);
    }
}

// ---------------------------------------------------------------- backward of the fisheye resample
// The forward is separable: y[o_i, o_j] = sum_{a,b} wh_a(o_i) ww_b(o_j) x[h_a(o_i), w_b(o_j)].  Its transpose is computed as a
// GATHER: roi_taps_kernel tabulates, per sample and axis, the forward's own taps (same device functions, so the pair is the
// exact transpose, degenerate / NaN boxes included) and, for every INPUT index, the first and last output index that reads it
// (the piecewise-linear maps are monotone, so the readers are contiguous; entries inside the range that do not read the index
// simply weigh 0); roi_resample_bwd_kernel then sums dy over that small rectangle.  No atomics: bit-reproducible.
// grid B, 256 threads.  tabh [B][oh], tabw [B][ow] (Tap), rngh [B][ih], rngw [B][iw] (int2 = first, last; empty: 0, -1)
__global__ void __launch_bounds__(256)
roi_taps_kernel(const float* __restrict__ box, Tap* __restrict__ tabh, Tap* __restrict__ tabw, int2* __restrict__ rngh,
                int2* __restrict__ rngw, int ih, int iw, int oh, int ow, int h_full, int w_full, int roi_h, int roi_w,
                int eval_h, int eval_w, int direction) {
    const int b = blockIdx.x;
    const float x0 = box[b * 6 + 0], y0 = box[b * 6 + 1], x1 = box[b * 6 + 3], y1 = box[b * 6 + 4];
    Tap* th = tabh + (int64_t)b * oh;
    Tap* tw = tabw + (int64_t)b * ow;
    for (int n = threadIdx.x; n < oh; n += blockDim.x) {
        const float c = direction == 0 ? fisheye_fwd(x0, x1, h_full - 1, roi_h, eval_h, n) : fisheye_back(x0, x1, h_full - 1, roi_h, eval_h, n);
        th[n] = make_tap(c, ih);
    }
    for (int n = threadIdx.x; n < ow; n += blockDim.x) {
        const float c = direction == 0 ? fisheye_fwd(y0, y1, w_full - 1, roi_w, eval_w, n) : fisheye_back(y0, y1, w_full - 1, roi_w, eval_w, n);
        tw[n] = make_tap(c, iw);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ih + iw; i += blockDim.x) {
        const bool isw = i >= ih;
        const int k = isw ? i - ih : i;
        const Tap* t = isw ? tw : th;
        const int n_out = isw ? ow : oh;
        int lo = 0, hi = -1;
        bool found = false;
        for (int n = 0; n < n_out; ++n) {
            const Tap q = t[n];
            if ((q.i0 == k && q.w0 != 0.f) || (q.i1 == k && q.w1 != 0.f)) {
                if (!found) { lo = n; found = true; }
                hi = n;
            }
        }
        (isw ? rngw : rngh)[(int64_t)b * (isw ? iw : ih) + k] = make_int2(lo, hi);
    }
}

// dy [B,oh,ow,d,C] -> dx [B,ih,iw,d,C]
template <typename T>
__global__ void __launch_bounds__(256)
roi_resample_bwd_kernel(const T* __restrict__ dy, const Tap* __restrict__ tabh, const Tap* __restrict__ tabw,
                        const int2* __restrict__ rngh, const int2* __restrict__ rngw, T* __restrict__ dx, int ih, int iw,
                        int oh, int ow, int d, int C) {
    const int b = blockIdx.y;
    const int cg = C / 4;
    const int64_t total = (int64_t)ih * iw * d * cg;
    const T* yb = dy + (int64_t)b * oh * ow * d * C;
    T* xb = dx + (int64_t)b * ih * iw * d * C;
    const Tap* th = tabh + (int64_t)b * oh;
    const Tap* tw = tabw + (int64_t)b * ow;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % cg) * 4;
        int64_t t = idx / cg;
        const int dd = (int)(t % d); t /= d;
        const int j = (int)(t % iw);
        const int i = (int)(t / iw);
        const int2 rh = rngh[(int64_t)b * ih + i], rw = rngw[(int64_t)b * iw + j];
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int n = rh.x; n <= rh.y; ++n) {
            const Tap a = th[n];
            const float wh = (a.i0 == i ? a.w0 : 0.f) + (a.i1 == i ? a.w1 : 0.f);
            if (wh == 0.f) continue;
            for (int m = rw.x; m <= rw.y; ++m) {
                const Tap q = tw[m];
                const float ww = (q.i0 == j ? q.w0 : 0.f) + (q.i1 == j ? q.w1 : 0.f);
                if (ww == 0.f) continue;
                float v[4];
                load4(yb + (((int64_t)n * ow + m) * d + dd) * C + c4, v);
                const float wt = wh * ww;
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[k] = fmaf(wt, v[k], acc[k]);
            }
        }
        store4(xb + (idx / cg) * C + c4, acc);
    }
}

}  // namespace ltu

using namespace ltu;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" size_t ltu_roi_resample_bwd_workspace(int B, int h, int w, int eval_h, int eval_w) {
    if (B <= 0 || h <= 0 || w <= 0 || eval_h <= 0 || eval_w <= 0) return 0;
    // either direction: one axis pair is (h, w), the other (eval_h, eval_w)
    return (size_t)B * ((size_t)(h + w + eval_h + eval_w) * 16);
}

extern "C" int ltu_roi_resample_bwd(const void* dy, const float* box, void* dx, void* ws, size_t ws_bytes, int B, int h, int w,
                                    int d, int C, int roi_h, int roi_w, int eval_h, int eval_w, int direction, int dtype,
                                    ltu_stream_t stream) {
    LTU_ARG_CHECK(dy && box && dx && ws, "roi_resample_bwd: null pointer");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "roi_resample_bwd: bad dtype %d", dtype);
    LTU_ARG_CHECK(B > 0 && B <= 65535 && h > 0 && w > 0 && d > 0 && C % 4 == 0, "roi_resample_bwd: bad shape");
    LTU_ARG_CHECK(direction == 0 || direction == 1, "roi_resample_bwd: direction must be 0 or 1");
    LTU_ARG_CHECK(roi_h > 1 && roi_w > 1 && eval_h > roi_h && eval_w > roi_w, "roi_resample_bwd: bad ROI constants");
    LTU_ARG_CHECK(aligned16(dy) && aligned16(dx) && aligned16(ws), "roi_resample_bwd: pointers must be 16-byte aligned");
    LTU_ARG_CHECK(ws_bytes >= ltu_roi_resample_bwd_workspace(B, h, w, eval_h, eval_w), "roi_resample_bwd: workspace too small");
    // forward: x [ih, iw] -> y [oh, ow]; here dy has the forward's OUTPUT extent and dx its INPUT extent
    const int ih = direction == 0 ? h : eval_h, iw = direction == 0 ? w : eval_w;
    const int oh = direction == 0 ? eval_h : h, ow = direction == 0 ? eval_w : w;
    Tap* tabh = reinterpret_cast<Tap*>(ws);
    Tap* tabw = tabh + (size_t)B * oh;
    int2* rngh = reinterpret_cast<int2*>(tabw + (size_t)B * ow);
    int2* rngw = rngh + (size_t)B * ih;
    cudaStream_t st = (cudaStream_t)stream;
    roi_taps_kernel<<<B, 256, 0, st>>>(box, tabh, tabw, rngh, rngw, ih, iw, oh, ow, h, w, roi_h, roi_w, eval_h, eval_w, direction);
    LTU_LAUNCH_CHECK("roi_taps");
    int64_t bx = ceil_div64((int64_t)ih * iw * d * (C / 4), 256);
    const int64_t cap = ceil_div64((int64_t)sm_count() * 16, B);
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    if (dtype == LTU_F32) roi_resample_bwd_kernel<float><<<dim3((unsigned)bx, B), 256, 0, st>>>((const float*)dy, tabh, tabw, rngh, rngw, (float*)dx, ih, iw, oh, ow, d, C);
    else roi_resample_bwd_kernel<bf16><<<dim3((unsigned)bx, B), 256, 0, st>>>((const bf16*)dy, tabh, tabw, rngh, rngw, (bf16*)dx, ih, iw, oh, ow, d, C);
    LTU_LAUNCH_CHECK("roi_resample_bwd");
    count_launch(2);
    return LTU_OK;
}

extern "C" size_t ltu_roi_bbox_scratch(int B, int h, int w) { return (size_t)B * (h + w) * sizeof(int); }

extern "C" int ltu_roi_bbox(const float* fg, float* box, void* scratch, size_t scratch_bytes, int B, int h, int w,
                            int d, int min_h, int min_w, float thr, ltu_stream_t stream) {
    LTU_ARG_CHECK(fg && box && scratch, "roi_bbox: null pointer");
    LTU_ARG_CHECK(B > 0 && B <= 65535 && h > 0 && w > 0 && d > 0, "roi_bbox: bad shape");
    LTU_ARG_CHECK(scratch_bytes >= ltu_roi_bbox_scratch(B, h, w), "roi_bbox: scratch too small");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(scratch, 0, ltu_roi_bbox_scratch(B, h, w), st);
    if (e != cudaSuccess) { set_error("roi_bbox: memset failed: %s", cudaGetErrorString(e)); return (int)e; }
    int lines = h * w;
    int bx = (lines + 7) / 8;
    int cap = (sm_count() * 8 + B - 1) / B;
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    roi_profile_kernel<<<dim3(bx, B), 256, 0, st>>>(fg, (int*)scratch, h, w, d, thr);
    LTU_LAUNCH_CHECK("roi_profile");
    roi_box_kernel<<<B, 32, 0, st>>>((const int*)scratch, box, h, w, d, min_h, min_w);
    LTU_LAUNCH_CHECK("roi_box");
    count_launch(2);
    return LTU_OK;
}

extern "C" int ltu_roi_resample(const void* x, const float* box, void* y, int B, int h, int w, int d, int C,
                                int roi_h, int roi_w, int eval_h, int eval_w, int direction, int dtype,
                                ltu_stream_t stream) {
    LTU_ARG_CHECK(x && box && y, "roi_resample: null pointer");
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "roi_resample: bad dtype %d", dtype);
    const int vn = dtype == LTU_F32 ? 4 : 8;
    LTU_ARG_CHECK(B > 0 && B <= 65535 && h > 0 && w > 0 && d > 0 && C % vn == 0, "roi_resample: bad shape");
    LTU_ARG_CHECK(direction == 0 || direction == 1, "roi_resample: direction must be 0 or 1");
    LTU_ARG_CHECK(roi_h > 1 && roi_w > 1 && eval_h > roi_h && eval_w > roi_w, "roi_resample: bad ROI constants");
    LTU_ARG_CHECK(aligned16(x) && aligned16(y), "roi_resample: pointers must be 16-byte aligned");
    const int ih = direction == 0 ? h : eval_h, iw = direction == 0 ? w : eval_w;
    const int oh = direction == 0 ? eval_h : h, ow = direction == 0 ? eval_w : w;
    LTU_ARG_CHECK(oh <= 65535, "roi_resample: too many output rows");
    dim3 grid((unsigned)((ow + 3) / 4), (unsigned)oh, (unsigned)B);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LTU_F32)
        roi_resample_kernel<float><<<grid, 256, 0, st>>>((const float*)x, box, (float*)y, ih, iw, oh, ow, d, C, h, w, roi_h, roi_w, eval_h, eval_w, direction);
    else
        roi_resample_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)x, box, (bf16*)y, ih, iw, oh, ow, d, C, h, w, roi_h, roi_w, eval_h, eval_w, direction);
    LTU_LAUNCH_CHECK("roi_resample");
    count_launch(1);
    return LTU_OK;
}
