// Shared device/host helpers for the lintransunet_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/ltu_b200.h"

namespace ltu {

// ---------------------------------------------------------------- error plumbing
void set_error(const char* fmt, ...);

#define LTU_ARG_CHECK(cond, ...)                    \
    do {                                            \
        if (!(cond)) {                              \
            ::ltu::set_error(__VA_ARGS__);          \
            return LTU_ERR_ARG;                     \
        }                                           \
    } while (0)

// never synchronises: only picks up launch-configuration errors
#define LTU_LAUNCH_CHECK(name)                                                   \
    do {                                                                         \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            ::ltu::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return (int)e__;                                                     \
        }                                                                        \
    } while (0)

int sm_count();   // cached per device

__host__ __device__ static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------- element types
using bf16 = __nv_bfloat16;

template <typename T> struct Vec;   // 16-byte vector of T
template <> struct Vec<float> { static constexpr int N = 4; };
template <> struct Vec<bf16>  { static constexpr int N = 8; };

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }

// load `Vec<T>::N` consecutive elements (16-byte aligned) into fp32 registers
__device__ __forceinline__ void load_vec(const float* p, float (&r)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
}
__device__ __forceinline__ void load_vec(const bf16* p, float (&r)[8]) {
    uint4 v = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r[2 * i]     = __uint_as_float(w[i] << 16);
        r[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ void store_vec(float* p, const float (&r)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void store_vec(bf16* p, const float (&r)[8]) {
    uint4 v;
    v.x = pack_bf16x2(r[0], r[1]); v.y = pack_bf16x2(r[2], r[3]);
    v.z = pack_bf16x2(r[4], r[5]); v.w = pack_bf16x2(r[6], r[7]);
    *reinterpret_cast<uint4*>(p) = v;
}
// 4-element (fp32: 16 B, bf16: 8 B) loads used by the conv gather
__device__ __forceinline__ void load4(const float* p, float (&r)[4]) { load_vec(p, r); }
__device__ __forceinline__ void load4(const bf16* p, float (&r)[4]) {
    uint2 v = *reinterpret_cast<const uint2*>(p);
    r[0] = __uint_as_float(v.x << 16); r[1] = __uint_as_float(v.x & 0xffff0000u);
    r[2] = __uint_as_float(v.y << 16); r[3] = __uint_as_float(v.y & 0xffff0000u);
}
__device__ __forceinline__ void store4(float* p, const float (&r)[4]) { store_vec(p, r); }
__device__ __forceinline__ void store4(bf16* p, const float (&r)[4]) {
    uint2 v;
    v.x = pack_bf16x2(r[0], r[1]); v.y = pack_bf16x2(r[2], r[3]);
    *reinterpret_cast<uint2*>(p) = v;
}

// erf-based GELU, x * Phi(x) = max(x,0) - |x| erfc(z) / 2, z = |x| / sqrt(2), with
//   erfc(z) = exp(-z^2) erfcx(z),   erfcx(z) ~ P9(z/2 - 1) on [0, 4]   (weighted Chebyshev fit, relative error 8.3e-5 in fp32
//   Horner form; z is clamped at 4, beyond which the term is below 6e-8 in absolute value).
// The exponential carries the decay, so the RELATIVE accuracy holds along the whole negative tail (the Abramowitz & Stegun
// 7.1.28 form used before, (1 + a1 z + ... + a6 z^6)^-16, is accurate to 3e-7 ABSOLUTE: 14 % relative at x = -5, 16 % of the
// bf16-rounded results off by an ulp against 0.4 % here) and it costs ONE MUFU (ex2) instead of two (lg2, ex2): the XU pipe,
// 4 lanes per clock and scheduler, was the busiest pipe of the GELU epilogues.  Used where the result is rounded to bf16
// (the fp32 path keeps erff).
#define LTU_GELU_C0 2.554074526e-01f
#define LTU_GELU_C1 -2.135173380e-01f
#define LTU_GELU_C2 1.665058434e-01f
#define LTU_GELU_C3 -1.244897023e-01f
#define LTU_GELU_C4 9.327740967e-02f
#define LTU_GELU_C5 -5.744498968e-02f
#define LTU_GELU_C6 1.998868026e-02f
#define LTU_GELU_C7 -2.029449306e-02f
#define LTU_GELU_C8 3.327577561e-02f
#define LTU_GELU_C9 -1.571550407e-02f
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    const float s = fminf(fmaf(ax, 0.35355339059327373f, -1.f), 1.f);        // z / 2 - 1
    float p = fmaf(LTU_GELU_C9, s, LTU_GELU_C8);
    p = fmaf(p, s, LTU_GELU_C7); p = fmaf(p, s, LTU_GELU_C6); p = fmaf(p, s, LTU_GELU_C5); p = fmaf(p, s, LTU_GELU_C4);
    p = fmaf(p, s, LTU_GELU_C3); p = fmaf(p, s, LTU_GELU_C2); p = fmaf(p, s, LTU_GELU_C1); p = fmaf(p, s, LTU_GELU_C0);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"((x * -0.7213475108f) * x));   // exp(-x^2 / 2)
    return fmaf((p * e) * ax, -0.5f, fmaxf(x, 0.f));
}

// Two elements per instruction on the packed fp32 pipe (fma.rn.f32x2 / mul.rn.f32x2): the same arithmetic as gelu_erf,
// bit for bit (every operation is the same IEEE fp32 operation, only issued in pairs).
__device__ __forceinline__ uint64_t f32x2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f32x2_get(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ void gelu_erf_pair(float x0, float x1, float& y0, float& y1) {
    const float a0 = fabsf(x0), a1 = fabsf(x1);
    const uint64_t S = f32x2(fminf(fmaf(a0, 0.35355339059327373f, -1.f), 1.f), fminf(fmaf(a1, 0.35355339059327373f, -1.f), 1.f));
    uint64_t P = fma2(0xBC80BDCDBC80BDCDULL, S, 0x3D084C2E3D084C2EULL);      // c9 s + c8
    P = fma2(P, S, 0xBCA640A3BCA640A3ULL);                                    // c7
    P = fma2(P, S, 0x3CA3BF4D3CA3BF4DULL);                                    // c6
    P = fma2(P, S, 0xBD6B4B70BD6B4B70ULL);                                    // c5
    P = fma2(P, S, 0x3DBF083A3DBF083AULL);                                    // c4
    P = fma2(P, S, 0xBDFEF475BDFEF475ULL);                                    // c3
    P = fma2(P, S, 0x3E2A80823E2A8082ULL);                                    // c2
    P = fma2(P, S, 0xBE5AA44ABE5AA44AULL);                                    // c1
    P = fma2(P, S, 0x3E82C4C43E82C4C4ULL);                                    // c0
    const uint64_t X = f32x2(x0, x1);
    float t0, t1, e0, e1;
    f32x2_get(mul2(mul2(X, 0xBF38AA3BBF38AA3BULL), X), t0, t1);               // (x * -log2(e)/2) * x
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(t0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(t1));
    const uint64_t W = mul2(mul2(P, f32x2(e0, e1)), f32x2(a0, a1));           // (p e) |x|
    f32x2_get(fma2(W, 0xBF000000BF000000ULL, f32x2(fmaxf(x0, 0.f), fmaxf(x1, 0.f))), y0, y1);
}

// ---------------------------------------------------------------- warp helpers
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ uint32_t smem_u32_generic(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
// src_bytes in {0,16}: 0 => zero-fill (used for padding / out-of-range rows)
__device__ __forceinline__ void cp_async16_zfill(void* smem, const void* gmem, int src_bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---------------------------------------------------------------- programmatic dependent launch
// A kernel that starts with pdl_prologue() and is launched through launch_pdl() may be scheduled while its
// predecessor in the stream is still draining (the predecessor must have executed griddepcontrol.launch_dependents,
// which pdl_prologue() also does): block scheduling, barrier / tensor-map set-up and the launch latency overlap the
// predecessor's tail, and griddepcontrol.wait returns once the predecessor has COMPLETED and its writes are visible.
// Nothing may touch global memory before the wait.  Both instructions are no-ops in an ordinary launch; the edges
// survive CUDA-graph stream capture.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    static const bool enabled = [] { const char* e = getenv("LTU_PDL"); return !(e && e[0] == '0'); }();   // A/B switch
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = enabled ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

}  // namespace ltu
