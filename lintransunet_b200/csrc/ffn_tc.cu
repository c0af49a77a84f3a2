// Fused feed-forward half of SelfAttentionLayer (model/trans_block.py:207-210) for d_model = 128:
//
//     y = LayerNorm2( t + W2 . gelu(W1 . t + b1) + b2 )           t, y : bf16 [rows][128]
//
// as ONE persistent, warp-specialised tcgen05 kernel.  The separate path moves 13 row-units of HBM
// traffic per token row (GEMM1 r1+w2, gelu r2+w2, GEMM2 r2+w1, add+LayerNorm r2+w1); this kernel
// moves 2 (read t once, write y once): the 256-wide hidden activation never leaves the SM.
//
//   grid = min(#row tiles, #SMs), 384 threads, one 128-row tile at a time per CTA
//   warp 0      TMA: W1 (64 KB) and W2 (64 KB) once per CTA, then the t tiles (2 x [128 x 64] boxes,
//               SWIZZLE_128B) into a 2-slot ring; it also TMA-stores the finished y tile from the same
//               slot (the epilogue overwrites its own residual rows in place)
//   warp 1      one lane issues tcgen05.mma:  acc1[128x256] = T . W1^T           (A, B from smem)
//                                             acc2[128x128] = H . W2^T           (A from TENSOR MEMORY)
//   warps 4-11  epilogue (2 warps per TMEM lane quarter, splitting the columns):
//               epi1: acc1 -> +b1 -> erf-GELU -> bf16 pairs -> tcgen05.st into the H region of TMEM
//                     (64-column chunks, each with its own mbarrier so GEMM2 starts on chunk 0
//                     while the later chunks are still being activated)
//               epi2: acc2 -> +b2 + residual (read back from the swizzled t tile in smem) ->
//                     two-pass LayerNorm (row halves exchanged through smem) -> bf16 -> same smem slot
//   TMEM: acc1 cols [0,256) | H (packed bf16) cols [256,384) | acc2 cols [384,512)
//
// GEMM1 of tile i+1 is issued as soon as epi1 of tile i has drained acc1, so it runs under epi2(i).
#include <cuda.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);

constexpr int kFfnThreads = 384;
constexpr int kFfnSlots = 2;
constexpr int kFfnEpiThreads = 256;
constexpr uint32_t kFfnW1Bytes = 256 * 128 * 2;     // [256 out][128 in] bf16
constexpr uint32_t kFfnW2Bytes = 128 * 256 * 2;     // [128 out][256 in] bf16
constexpr uint32_t kFfnXBytes = 128 * 128 * 2;      // one 128-row tile
constexpr uint32_t kFfnOffW1 = 0;
constexpr uint32_t kFfnOffW2 = kFfnOffW1 + kFfnW1Bytes;
constexpr uint32_t kFfnOffX = kFfnOffW2 + kFfnW2Bytes;
constexpr uint32_t kFfnOffTail = kFfnOffX + kFfnSlots * kFfnXBytes;
constexpr uint32_t kTmemAcc1 = 0, kTmemH = 256, kTmemAcc2 = 384;

struct FfnTail {
    uint64_t w_full, x_full[kFfnSlots], out_ready[kFfnSlots], acc1_full, acc1_empty, h_full[4], acc2_full;
    uint32_t tmem_slot, pad_;
    float b1[256], b2[128], gamma[128], beta[128];
    float xs[2][128], xq[2][128];
};

struct FfnParams {
    const float* b1; const float* b2; const float* gamma; const float* beta;
    float eps;
    int tiles;
};

// erf-based GELU, x * Phi(x), with erfc from Abramowitz & Stegun 7.1.28
//   erfc(z) = (1 + a1 z + ... + a6 z^6)^-16,  |error| <= 3e-7   (z = |x| / sqrt(2))
// gelu(x) = max(x,0) - |x| erfc(z) / 2.  One MUFU (rcp) per element, no branches.
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    float p = fmaf(z, 0.0000430638f, 0.0002765672f);
    p = fmaf(p, z, 0.0001520143f);
    p = fmaf(p, z, 0.0092705272f);
    p = fmaf(p, z, 0.0422820123f);
    p = fmaf(p, z, 0.0705230784f);
    p = fmaf(p, z, 1.0f);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
    r *= r; r *= r; r *= r; r *= r;
    return fmaxf(x, 0.f) - fabsf(0.5f * x * r);
}

__global__ void __launch_bounds__(kFfnThreads, 1)
ffn128_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_y,
              const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2, const FfnParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    FfnTail* tail = reinterpret_cast<FfnTail*>(smem + kFfnOffTail);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&tail->w_full), 1);
        for (int s = 0; s < kFfnSlots; ++s) {
            mbar_init(smem_u32(&tail->x_full[s]), 1);
            mbar_init(smem_u32(&tail->out_ready[s]), kFfnEpiThreads);
        }
        mbar_init(smem_u32(&tail->acc1_full), 1);
        mbar_init(smem_u32(&tail->acc1_empty), kFfnEpiThreads);
        for (int c = 0; c < 4; ++c) mbar_init(smem_u32(&tail->h_full[c]), kFfnEpiThreads / 2);
        mbar_init(smem_u32(&tail->acc2_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += kFfnThreads) tail->b1[i] = p.b1[i];
    for (int i = threadIdx.x; i < 128; i += kFfnThreads) {
        tail->b2[i] = p.b2[i]; tail->gamma[i] = p.gamma[i]; tail->beta[i] = p.beta[i];
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tail->tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_slot;

    if (warp == 0) {
        // =========================== TMA producer / store issuer ===========================
        if (lane == 0) {
            const uint32_t wbar = smem_u32(&tail->w_full);
            mbar_expect_tx(wbar, kFfnW1Bytes + kFfnW2Bytes);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + kFfnOffW1 + kb * 32768, &tm_w1, kb * 64, 0, wbar);
            for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + kFfnOffW2 + kb * 16384, &tm_w2, kb * 64, 0, wbar);
            for (int i = 0; i < kFfnSlots && i < n_my; ++i) {
                const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128;
                const uint32_t xb = smem_u32(&tail->x_full[i]), dst = sbase + kFfnOffX + i * kFfnXBytes;
                mbar_expect_tx(xb, kFfnXBytes);
                tma_load_2d(dst, &tm_x, 0, row0, xb);
                tma_load_2d(dst + 16384, &tm_x, 64, row0, xb);
            }
            for (int i = 0; i < n_my; ++i) {
                const int s = i % kFfnSlots;
                const int row0 = ((int)blockIdx.x + i * (int)gridDim.x) * 128;
                const uint32_t slot = sbase + kFfnOffX + s * kFfnXBytes;
                mbar_wait(smem_u32(&tail->out_ready[s]), (i / kFfnSlots) & 1);
                tma_store_2d(&tm_y, slot, 0, row0);
                tma_store_2d(&tm_y, slot + 16384, 64, row0);
                tma_store_commit();
                tma_store_wait_read();                           // the slot may be overwritten now
                if (i + kFfnSlots < n_my) {
                    const int nrow0 = ((int)blockIdx.x + (i + kFfnSlots) * (int)gridDim.x) * 128;
                    const uint32_t xb = smem_u32(&tail->x_full[s]);
                    mbar_expect_tx(xb, kFfnXBytes);
                    tma_load_2d(slot, &tm_x, 0, nrow0, xb);
                    tma_load_2d(slot + 16384, &tm_x, 64, nrow0, xb);
                }
            }
            tma_store_wait_all();
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        constexpr uint32_t idesc1 = umma_idesc_bf16(128, 256), idesc2 = umma_idesc_bf16(128, 128);
        mbar_wait(smem_u32(&tail->w_full), 0);
        for (int i = 0; i < n_my; ++i) {
            const int s = i % kFfnSlots;
            mbar_wait(smem_u32(&tail->x_full[s]), (i / kFfnSlots) & 1);
            mbar_wait(smem_u32(&tail->acc1_empty), (i & 1) ^ 1);           // epi1 of the previous tile drained acc1
            tc_fence_after();
            if (lane == 0) {
                const uint32_t xa = sbase + kFfnOffX + s * kFfnXBytes;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t adesc = make_desc(xa + kb * 16384), bdesc = make_desc(sbase + kFfnOffW1 + kb * 32768);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + kTmemAcc1, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc1, (kb | k) != 0);
                }
                umma_commit(smem_u32(&tail->acc1_full));
            }
            __syncwarp();
            // GEMM2 over the hidden chunks in the order the two epilogue halves finish them: 0,2,1,3
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int c = ((o & 1) << 1) | (o >> 1);
                mbar_wait(smem_u32(&tail->h_full[c]), i & 1);
                tc_fence_after();
                if (lane == 0) {
                    const uint64_t bdesc = make_desc(sbase + kFfnOffW2 + c * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_ts(tmem_base + kTmemAcc2, tmem_base + kTmemH + (uint32_t)(c * 32 + k * 8),
                                     bdesc + (uint64_t)(k * 2), idesc2, (o | k) != 0);
                    if (o == 3) umma_commit(smem_u32(&tail->acc2_full));
                }
                __syncwarp();
            }
        }
    } else if (warp >= 4) {
        // =========================== epilogue ===========================
        const int e = warp - 4;
        const int q = e & 3;                       // TMEM lane quarter (== warp % 4)
        const int hh = e >> 2;                     // column half
        const int row = q * 32 + lane;             // tile row == TMEM lane
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int swz = row & 7;
        for (int i = 0; i < n_my; ++i) {
            const int s = i % kFfnSlots;
            // ---- epi1: hidden activation into TMEM
            mbar_wait(smem_u32(&tail->acc1_full), i & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c = hh * 2 + cc;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int col0 = c * 64 + half * 32;
                    float v[32];
                    tmem_ld32(tmem_base + lane_off + kTmemAcc1 + (uint32_t)col0, v);
                    if (cc == 1 && half == 1) {                  // last read of acc1 by this thread
                        tc_fence_before();
                        mbar_arrive(smem_u32(&tail->acc1_empty));
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float a = gelu_erf(v[2 * j] + tail->b1[col0 + 2 * j]);
                        const float b = gelu_erf(v[2 * j + 1] + tail->b1[col0 + 2 * j + 1]);
                        pk[j] = pack_bf16x2(a, b);
                    }
                    tmem_st16(tmem_base + lane_off + kTmemH + (uint32_t)(col0 >> 1), pk);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&tail->h_full[c]));
            }
            // ---- epi2: + b2 + residual -> LayerNorm -> bf16, in place in the t tile
            mbar_wait(smem_u32(&tail->x_full[s]), (i / kFfnSlots) & 1);     // acquire the TMA-written tile
            mbar_wait(smem_u32(&tail->acc2_full), i & 1);
            tc_fence_after();
            float y[64];
            {
                float v[32];
                tmem_ld32(tmem_base + lane_off + kTmemAcc2 + (uint32_t)(hh * 64), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) y[j] = v[j];
                tmem_ld32(tmem_base + lane_off + kTmemAcc2 + (uint32_t)(hh * 64 + 32), v);
#pragma unroll
                for (int j = 0; j < 32; ++j) y[32 + j] = v[j];
            }
            tc_fence_before();
            unsigned char* xrow = smem + kFfnOffX + s * kFfnXBytes + hh * 16384 + row * 128;
            float s1 = 0.f;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float r8[8];
                load_vec(reinterpret_cast<const bf16*>(xrow + ((jj ^ swz) << 4)), r8);
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float val = y[jj * 8 + t] + tail->b2[hh * 64 + jj * 8 + t] + r8[t];
                    y[jj * 8 + t] = val;
                    s1 += val;
                }
            }
            tail->xs[hh][row] = s1;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float mean = (s1 + tail->xs[hh ^ 1][row]) * (1.f / 128.f);
            float s2 = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) { const float d = y[j] - mean; s2 = fmaf(d, d, s2); }
            tail->xq[hh][row] = s2;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float rstd = rsqrtf((s2 + tail->xq[hh ^ 1][row]) * (1.f / 128.f) + p.eps);
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                float o8[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const int col = hh * 64 + jj * 8 + t;
                    o8[t] = fmaf((y[jj * 8 + t] - mean) * rstd, tail->gamma[col], tail->beta[col]);
                }
                store_vec(reinterpret_cast<bf16*>(xrow + ((jj ^ swz) << 4)), o8);
            }
            fence_async_smem();                                  // generic-proxy writes -> visible to the TMA store
            mbar_arrive(smem_u32(&tail->out_ready[s]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return (EncodeTiledFn)ptr;
    }();
    return fn;
}

// Row-major bf16 matrix [rows][cols] -> map with a [box_rows x 64-column] box, SWIZZLE_128B (the K-major UMMA
// operand layout); out-of-range rows read as zeros and are not written.
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LTU_ERR_ARG; }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * 2};
    const cuuint32_t box[2] = {64, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return LTU_ERR_ARG; }
    return LTU_OK;
}

}  // namespace ltu

using namespace ltu;

extern "C" int ltu_ffn_fused_supported(int C) { return C == 128 ? 1 : 0; }

extern "C" int ltu_ffn_fused(const void* x, int64_t rows, int C, const void* w1_bf16, const float* b1, const void* w2_bf16,
                             const float* b2, const float* gamma, const float* beta, float eps, void* y,
                             ltu_stream_t stream) {
    LTU_ARG_CHECK(C == 128, "ffn_fused: d_model %d not supported (128)", C);
    LTU_ARG_CHECK(x && y && w1_bf16 && w2_bf16 && b1 && b2 && gamma && beta, "ffn_fused: null pointer");
    LTU_ARG_CHECK(rows > 0 && rows < ((int64_t)1 << 31) - 256, "ffn_fused: bad row count");
    LTU_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)w1_bf16 & 15) == 0 &&
                  ((uintptr_t)w2_bf16 & 15) == 0, "ffn_fused: pointers must be 16-byte aligned");
    CUtensorMap tx, ty, tw1, tw2;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tx, x, (uint64_t)rows, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&ty, y, (uint64_t)rows, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw1, w1_bf16, 256, 128, 256)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw2, w2_bf16, 128, 256, 128)) != LTU_OK) return rc;
    FfnParams p;
    p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.tiles = (int)((rows + 127) / 128);
    const size_t smem = 1024 + kFfnOffTail + sizeof(FfnTail);
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(ffn128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured_dev = dev;
    }
    int grid = sm_count();
    if (grid > p.tiles) grid = p.tiles;
    ffn128_kernel<<<grid, kFfnThreads, smem, (cudaStream_t)stream>>>(tx, ty, tw1, tw2, p);
    LTU_LAUNCH_CHECK("ffn_fused");
    count_launch(1);
    return LTU_OK;
}
