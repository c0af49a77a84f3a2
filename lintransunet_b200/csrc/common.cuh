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

// erf-based GELU, x * Phi(x), with erfc from Abramowitz & Stegun 7.1.28
//   erfc(z) = (1 + a1 z + ... + a6 z^6)^-16,  |error| <= 3e-7   (z = |x| / sqrt(2), folded into the a_k)
// gelu(x) = max(x,0) - |x| erfc(z) / 2.  The 16th power goes to the MUFU pipe, p^-16 = ex2(-16 lg2 p), and every
// multiply-add has an immediate operand: 6 FFMA + 2 FMUL + 1 FFMA on the FMA pipe, 2 MUFU, 1 FMNMX, no branches.
// Used where the result is rounded to bf16 (the fp32 path keeps erff).
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x);
    float p = fmaf(z, 0.0000430638f * 0.125f, 0.0002765672f * 0.17677669529663687f);
    p = fmaf(p, z, 0.0001520143f * 0.25f);
    p = fmaf(p, z, 0.0092705272f * 0.35355339059327373f);
    p = fmaf(p, z, 0.0422820123f * 0.5f);
    p = fmaf(p, z, 0.0705230784f * 0.70710678118654752f);
    p = fmaf(p, z, 1.0f);
    float l, r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(p));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(l * -16.f));
    return fmaf(-0.5f * z, r, fmaxf(x, 0.f));
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
