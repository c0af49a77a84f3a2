// Fused feed-forward half of SelfAttentionLayer (model/trans_block.py:207-210) for d_model = 128:
//
//     y = LayerNorm2( t + W2 . gelu(W1 . t + b1) + b2 )           t, y : bf16 [rows][128]
//
// as ONE persistent, warp-specialised tcgen05 kernel.  The separate path moves 13 row-units of HBM
// traffic per token row (GEMM1 r1+w2, gelu r2+w2, GEMM2 r2+w1, add+LayerNorm r2+w1); this kernel
// moves 2 (read t once, write y once): the 256-wide hidden activation never leaves the SM.
//
//   grid = min(#row tiles, #SMs), 512 threads; a CTA walks its 128-row tiles T, T+grid, ... (tile t of
//   the CTA uses smem slot / TMEM buffer t & 1).  Two specialised groups of 8 warps, two warps per TMEM
//   lane quarter, one row per thread:
//     warps 0-7   e1: acc1 -> +b1 -> erf-GELU -> bf16 pairs -> tcgen05.st over the thread's OWN acc1
//                 columns (128 hidden columns per thread)
//     warps 8-15  e2: acc2 -> +b2 + residual (the row's own x, read from the tile's shared-memory slot) ->
//                 LayerNorm (per-thread mean/M2 merged with the row partner by Chan's formula, one smem
//                 exchange) -> bf16 written over the residual in the slot -> one TMA store per warp
//                 (64 output columns per thread; no per-thread global access: a row per lane costs 32 L1
//                 wavefronts per instruction, which the GELU warps' shared-memory reads had to share)
//   The GELU warps carry ~75 % of the instructions and never wait for the tensor pipe: acc1 of tile t+1
//   is complete long before e1(t) ends.  There are no dedicated producer / MMA warps (they would only
//   spin on barriers and cost issue slots).  MMA issue is warp-uniform (the whole warp runs the descriptor
//   arithmetic, one elected lane issues):
//     warp 12 lane 0: TMA load of tile t+2 (2 x [128 x 64] boxes, SWIZZLE_128B) once the stores of y(t) have
//                     read the slot (the GELU of tile t+1 runs meanwhile)
//     warp 8        : GEMM2(t)  acc2[128x128] = H . W2^T  (A from TENSOR MEMORY) once H is published; a
//                     tcgen05.mma issue blocks for about the duration of the MMA, so this must not sit in
//                     a GELU warp (measured: +2300 cycles per tile on the critical path)
//     warp 12       : GEMM1(t+2)  acc1[128x256] = T . W1^T  (A, B from smem) once the e2 group has pulled
//                     acc2(t) into registers
//   W1 / W2 (64 KB each) are loaded once per CTA.
//   TMEM: two 256-column buffers.  acc1 fills a buffer; a thread turns 16 of its accumulator columns at a
//   time into 8 packed bf16 columns and writes them back over columns it has already consumed, so
//   H = [0,64) u [128,192) with no cross-warp hazard, and acc2 lands in the columns in between:
//   [64,128) (outputs 0-63) and [192,256) (outputs 64-127), as two N=64 GEMMs.
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);

constexpr int kFfnThreads = 512;
constexpr int kFfnGroupThreads = 256;
constexpr uint32_t kFfnW1Bytes = 256 * 128 * 2;     // [256 out][128 in] bf16
constexpr uint32_t kFfnW2Bytes = 128 * 256 * 2;     // [128 out][256 in] bf16
constexpr uint32_t kFfnXBytes = 128 * 128 * 2;      // one 128-row tile
constexpr uint32_t kFfnOffW1 = 0;
constexpr uint32_t kFfnOffW2 = kFfnOffW1 + kFfnW1Bytes;
constexpr uint32_t kFfnOffX = kFfnOffW2 + kFfnW2Bytes;
constexpr uint32_t kFfnOffTail = kFfnOffX + 2 * kFfnXBytes;

struct FfnTail {
    uint64_t w_full, x_full[2], x_free[2], acc1_full[2], h_full[2], acc2_full[2], buf_free[2];
    uint32_t tmem_slot, pad_;
    alignas(16) float b1[256], b2[128], gamma[128], beta[128];       // read as float4
    float2 xs[2][2][128];           // [tile parity][column half][row] = (mean, M2)
};

struct FfnParams {
    const bf16* x; bf16* y;
    const float* b1; const float* b2; const float* gamma; const float* beta;
    float eps;
    int tiles;
    int64_t rows;
    int mode;               // debug ablations: bit 0 = skip GELU math, bit 1 = skip LayerNorm math + stores
    long long* trace;       // debug: [2 roles][64 tiles][8 events] clock64 stamps of CTA 0 (nullptr = off)
};

// gelu_erf_pair(): common.cuh (erfc = exp(-z^2) P9(z/2 - 1), one MUFU, packed fp32 pipe)
__global__ void __launch_bounds__(kFfnThreads, 1)
ffn128_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
              const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_y, const FfnParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // keep the shared address space visible to the compiler (LDS/STS instead of generic accesses)
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    FfnTail* tail = reinterpret_cast<FfnTail*>(smem + kFfnOffTail);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&tail->w_full), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tail->x_full[s]), 1);
            mbar_init(smem_u32(&tail->x_free[s]), 8);
            mbar_init(smem_u32(&tail->acc1_full[s]), 1);
            mbar_init(smem_u32(&tail->h_full[s]), kFfnGroupThreads);
            mbar_init(smem_u32(&tail->acc2_full[s]), 1);
            mbar_init(smem_u32(&tail->buf_free[s]), kFfnGroupThreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += kFfnThreads) tail->b1[i] = p.b1[i];
    for (int i = threadIdx.x; i < 128; i += kFfnThreads) {
        tail->b2[i] = p.b2[i]; tail->gamma[i] = p.gamma[i]; tail->beta[i] = p.beta[i];
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tail->tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_slot;

    const int role = warp >> 3;                // 0: e1 (GELU) warps, 1: e2 (LayerNorm) warps
    const int e = warp & 7;
    const int q = e & 3;                       // TMEM lane quarter (== warp % 4)
    const int hh = e >> 2;                     // column half: hidden [128hh, +128) / output [64hh, +64)
    const int row = q * 32 + lane;             // tile row == TMEM lane
    const bool leader = (e == 0) && (lane == 0);
    auto stamp = [&](int t, int ev) {
        if (p.trace != nullptr && lane == 0 && (e == 0 || (role == 1 && e == 4 && (ev == 2 || ev == 3))) && blockIdx.x == 0 && t < 64)
            p.trace[(role * 64 + t) * 8 + ev] = clock64();
    };
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 256), idesc2 = umma_idesc_bf16(128, 64);
    auto load_tile = [&](int t) {
        const int b = t & 1;
        const int row0 = ((int)blockIdx.x + t * (int)gridDim.x) * 128;
        const uint32_t xb = smem_u32(&tail->x_full[b]), dst = sbase + kFfnOffX + b * kFfnXBytes;
        mbar_expect_tx(xb, kFfnXBytes);
        tma_load_2d(dst, &tm_x, 0, row0, xb);
        tma_load_2d(dst + 16384, &tm_x, 64, row0, xb);
    };
    auto gemm1 = [&](int b) {
        const uint32_t xslot = sbase + kFfnOffX + b * kFfnXBytes;
        if (!(p.mode & 8))
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            const uint64_t adesc = make_desc(xslot + kb * 16384), bdesc = make_desc(sbase + kFfnOffW1 + kb * 32768);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_elect(tmem_base + (uint32_t)(b * 256), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc1, (kb | k) != 0);
        }
        umma_commit_elect(smem_u32(&tail->acc1_full[b]));
    };

    if (role == 0) {
        // =========================== e1: GELU warps ===========================
        if (leader) {
            const uint32_t wbar = smem_u32(&tail->w_full);
            mbar_expect_tx(wbar, kFfnW1Bytes + kFfnW2Bytes);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + kFfnOffW1 + kb * 32768, &tm_w1, kb * 64, 0, wbar);
            for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + kFfnOffW2 + kb * 16384, &tm_w2, kb * 64, 0, wbar);
            load_tile(0);
            if (n_my > 1) load_tile(1);
        }
        __syncwarp();
        for (int t = 0; t < n_my; ++t) {
            const int b = t & 1;
            const uint32_t ph = (t >> 1) & 1;
            const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256);
            stamp(t, 0);
            mbar_wait(smem_u32(&tail->acc1_full[b]), ph);
            tc_fence_after();
            stamp(t, 1);
            // 16 accumulator columns per step (few live registers -> the 16 independent GELU chains of a step
            // interleave); the 8 packed result columns go back over columns this thread has already consumed
            {
                const uint32_t abase = tb + (uint32_t)(hh * 128);
                uint32_t cur[16], nxt[16];
                tmem_ld16_nowait(abase, cur);
                tmem_ld_wait();
#pragma unroll
                for (int s8 = 0; s8 < 8; ++s8) {
                    if (s8 < 7) tmem_ld16_nowait(abase + (uint32_t)(16 * (s8 + 1)), nxt);
                    const float4* b1v = reinterpret_cast<const float4*>(tail->b1 + hh * 128 + s8 * 16);
                    float4 bb[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) bb[j] = b1v[j];
                    uint32_t pk[8];
                    if (p.mode & 1) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) pk[j] = cur[2 * j] ^ cur[2 * j + 1];
                    } else
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float g0, g1, g2, g3;
                        gelu_erf_pair(__uint_as_float(cur[4 * j]) + bb[j].x, __uint_as_float(cur[4 * j + 1]) + bb[j].y, g0, g1);
                        gelu_erf_pair(__uint_as_float(cur[4 * j + 2]) + bb[j].z, __uint_as_float(cur[4 * j + 3]) + bb[j].w, g2, g3);
                        pk[2 * j] = pack_bf16x2(g0, g1);
                        pk[2 * j + 1] = pack_bf16x2(g2, g3);
                    }
                    tmem_st8(abase + (uint32_t)(8 * s8), pk);
                    if (s8 < 7) {
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 16; ++j) cur[j] = nxt[j];
                    }
                }
            }
            tmem_st_wait();
            tc_fence_before();
            stamp(t, 2);
            mbar_arrive(smem_u32(&tail->h_full[b]));
        }
    } else {
        // =========================== e2: residual + LayerNorm warps ===========================
        if (e == 0) {                                          // whole warp: the MMA issue is warp-uniform, one elected lane issues
            mbar_wait(smem_u32(&tail->w_full), 0);
            mbar_wait(smem_u32(&tail->x_full[0]), 0);
            gemm1(0);
            if (n_my > 1) { mbar_wait(smem_u32(&tail->x_full[1]), 0); gemm1(1); }
        }
        __syncwarp();
        for (int t = 0; t < n_my; ++t) {
            const int b = t & 1;
            const uint32_t ph = (t >> 1) & 1;
            const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * 256);
            const int64_t grow = ((int64_t)blockIdx.x + (int64_t)t * gridDim.x) * 128 + row;
            const bool row_ok = grow < p.rows;
            unsigned char* xrow = smem + kFfnOffX + b * kFfnXBytes + hh * 16384 + row * 128;     // this thread's half row of x
            if (e == 0) {                                       // GEMM2 as soon as the GELU group has published H
                const uint32_t tacc = tmem_base + (uint32_t)(b * 256);
                mbar_wait_sleep(smem_u32(&tail->h_full[b]), ph, 32);
                tc_fence_after();
                stamp(t, 7);
                if (!(p.mode & 4))
#pragma unroll
                for (int nh = 0; nh < 2; ++nh) {                // output columns [64nh, 64nh+64) -> TMEM [64+128nh, +64)
#pragma unroll
                    for (int k = 0; k < 16; ++k) {              // hidden [16k, 16k+16): packed at TMEM (k>>3)*128 + (k&7)*8
                        const uint64_t bdesc = make_desc(sbase + kFfnOffW2 + (k >> 2) * 16384 + nh * 8192) + (uint64_t)((k & 3) * 2);
                        umma_bf16_ts_elect(tacc + (uint32_t)(64 + 128 * nh), tacc + (uint32_t)((k >> 3) * 128 + (k & 7) * 8),
                                           bdesc, idesc2, k != 0);
                    }
                }
                umma_commit_elect(smem_u32(&tail->acc2_full[b]));
            }
            __syncwarp();
            stamp(t, 0);
            mbar_wait_sleep(smem_u32(&tail->acc2_full[b]), ph);
            tc_fence_after();
            stamp(t, 1);
            float y[64];
            {
                uint32_t v0[32], v1[32];
                tmem_ld32_nowait(tb + (uint32_t)(64 + 128 * hh), v0);
                tmem_ld32_nowait(tb + (uint32_t)(64 + 128 * hh + 32), v1);
                tmem_ld_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&tail->buf_free[b]));
#pragma unroll
                for (int j = 0; j < 32; ++j) { y[j] = __uint_as_float(v0[j]); y[32 + j] = __uint_as_float(v1[j]); }
            }
            if (p.mode & 2) {
                if (y[0] == 123.456f && row_ok) p.y[grow * 128 + hh * 64] = __float2bfloat16_rn(y[1]);
                if (lane == 0) mbar_arrive(smem_u32(&tail->x_free[b]));
            } else {
            mbar_wait(smem_u32(&tail->x_full[b]), ph);          // completed long ago: makes the TMA write visible to this thread
            float s1 = 0.f;
            const float4* b2v = reinterpret_cast<const float4*>(tail->b2 + hh * 64);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const uint4 rv = (p.mode & 16) ? make_uint4(0, 0, 0, 0) : *reinterpret_cast<const uint4*>(xrow + ((j ^ (row & 7)) << 4));
                const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
                const float4 ba = b2v[2 * j], bb = b2v[2 * j + 1];
                const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float lo = __uint_as_float(w[u] << 16), hi = __uint_as_float(w[u] & 0xffff0000u);
                    y[j * 8 + 2 * u] += bs[2 * u] + lo;
                    y[j * 8 + 2 * u + 1] += bs[2 * u + 1] + hi;
                    s1 += y[j * 8 + 2 * u] + y[j * 8 + 2 * u + 1];
                }
            }
            const float m_loc = s1 * (1.f / 64.f);
            float m2 = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) { const float d = y[j] - m_loc; m2 = fmaf(d, d, m2); }
            tail->xs[b][hh][row] = make_float2(m_loc, m2);
            stamp(t, 4);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            stamp(t, 5);
            const float2 other = tail->xs[b][hh ^ 1][row];
            const float mean = 0.5f * (m_loc + other.x);
            const float dm = m_loc - other.x;
            const float var = (m2 + other.y + dm * dm * 32.f) * (1.f / 128.f);      // Chan: n_a n_b / (n_a + n_b) = 32
            const float rstd = rsqrtf(var + p.eps);
            {
                const float4* gv = reinterpret_cast<const float4*>(tail->gamma + hh * 64);
                const float4* bv = reinterpret_cast<const float4*>(tail->beta + hh * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 g0 = gv[2 * j], g1 = gv[2 * j + 1], e0 = bv[2 * j], e1 = bv[2 * j + 1];
                    const float gs[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                    const float es[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
                    float o[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) o[u] = fmaf((y[j * 8 + u] - mean) * rstd, gs[u], es[u]);
                    uint4 ov;
                    ov.x = pack_bf16x2(o[0], o[1]); ov.y = pack_bf16x2(o[2], o[3]);
                    ov.z = pack_bf16x2(o[4], o[5]); ov.w = pack_bf16x2(o[6], o[7]);
                    *reinterpret_cast<uint4*>(xrow + ((j ^ (row & 7)) << 4)) = ov;        // over the residual it was read from
                }
            }
            // the warp's [32 rows x 64 columns] of y sit in the x slot in the TMA (SWIZZLE_128B) layout: one bulk store per warp,
            // rows past the end are clipped by the tensor map; the slot is free once the store has read it
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                tma_store_2d(&tm_y, sbase + kFfnOffX + b * kFfnXBytes + hh * 16384 + q * 4096, hh * 64,
                             ((int)blockIdx.x + t * (int)gridDim.x) * 128 + q * 32);
                tma_store_commit();
                tma_store_wait_read();
                mbar_arrive(smem_u32(&tail->x_free[b]));
            }
            }
            stamp(t, 6);
            if (e == 4 && t + 2 < n_my) {                       // whole warp: tile t+2 into the slot and the TMEM buffer of tile t
                mbar_wait(smem_u32(&tail->x_free[b]), ph);      // every warp's store has read the slot
                mbar_wait(smem_u32(&tail->buf_free[b]), ph);
                stamp(t, 2);
                if (lane == 0) load_tile(t + 2);
                __syncwarp();
                mbar_wait(smem_u32(&tail->x_full[b]), ph ^ 1);
                tc_fence_after();
                stamp(t, 3);
                gemm1(b);
            }
            __syncwarp();
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
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
int make_tmap_bf16_2d_w(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    return make_tmap_bf16_2d_w(map, base, rows, cols, box_rows, 64);
}
// box_cols in {16, 32, 64}: 32- / 64- / 128-byte rows with the matching TMA swizzle
int make_tmap_bf16_2d_w(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LTU_ERR_ARG; }
    const cuuint64_t gdim[2] = {cols, rows};
    const cuuint64_t gstride[1] = {cols * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapSwizzle sw = box_cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return LTU_ERR_ARG; }
    return LTU_OK;
}

// [batch][rows][cols] bf16 (dense) as a 3-D map with [1][box_rows][64] boxes, SWIZZLE_128B: rows past the end of a SAMPLE
// are zero-filled on loads and clipped on stores, so a kernel may tile every sample separately
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint64_t ld = 0) {
    if (ld == 0) ld = cols;                           // elements between two rows (a column slice of wider rows)
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LTU_ERR_ARG; }
    const cuuint64_t gdim[3] = {cols, rows, batch};
    const cuuint64_t gstride[2] = {ld * 2, rows * ld * 2};
    const cuuint32_t box[3] = {64, box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (3-D) failed (CUresult %d)", (int)r); return LTU_ERR_ARG; }
    return LTU_OK;
}

}  // namespace ltu

using namespace ltu;

namespace ltu {
int ffn256_launch(const void* x, int64_t rows, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                  const float* gamma, const float* beta, float eps, void* y, cudaStream_t stream);      // ffn256_tc.cu
}

// Dispatch hint for the model: d_model 128 always; d_model 256 only with LTU_FFN256=1 -- ffn256_kernel is correct (tests call
// it directly) but bound by the latency of its 128 KB weight ring (tensor pipe 21 % busy) and only matches cuBLAS + gelu +
// add_layernorm inside the step (profiles/r1_conv_variants.md).
extern "C" int ltu_ffn_fused_supported(int C) {
    if (C == 128) return 1;
    const char* e = getenv("LTU_FFN256");
    return (C == 256 && e && e[0] == '1') ? 1 : 0;
}

static int ffn_launch(const void* x, int64_t rows, int C, const void* w1_bf16, const float* b1, const void* w2_bf16,
                      const float* b2, const float* gamma, const float* beta, float eps, void* y, long long* trace,
                      ltu_stream_t stream, int mode = 0) {
    LTU_ARG_CHECK(C == 128 || C == 256, "ffn_fused: d_model %d not supported (128, 256)", C);
    LTU_ARG_CHECK(x && y && w1_bf16 && w2_bf16 && b1 && b2 && gamma && beta, "ffn_fused: null pointer");
    LTU_ARG_CHECK(rows > 0 && rows < ((int64_t)1 << 31) - 256, "ffn_fused: bad row count");
    LTU_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)w1_bf16 & 15) == 0 &&
                  ((uintptr_t)w2_bf16 & 15) == 0, "ffn_fused: pointers must be 16-byte aligned");
    if (C == 256) {
        LTU_ARG_CHECK(trace == nullptr && mode == 0, "ffn_fused: the trace / ablation hooks exist for d_model 128 only");
        return ffn256_launch(x, rows, w1_bf16, b1, w2_bf16, b2, gamma, beta, eps, y, (cudaStream_t)stream);
    }
    CUtensorMap tx, tw1, tw2;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tx, x, (uint64_t)rows, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw1, w1_bf16, 256, 128, 256)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw2, w2_bf16, 128, 256, 128)) != LTU_OK) return rc;
    CUtensorMap ty;
    if ((rc = make_tmap_bf16_2d(&ty, y, (uint64_t)rows, 128, 32)) != LTU_OK) return rc;
    FfnParams p;
    p.x = (const bf16*)x; p.y = (bf16*)y; p.rows = rows;
    p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.tiles = (int)((rows + 127) / 128);
    p.trace = trace;
    p.mode = mode;
    const size_t smem = 1024 + kFfnOffTail + sizeof(FfnTail);
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(ffn128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured_dev = dev;
    }
    int grid = sm_count();
    if (grid > p.tiles) grid = p.tiles;
    ffn128_kernel<<<grid, kFfnThreads, smem, (cudaStream_t)stream>>>(tx, tw1, tw2, ty, p);
    LTU_LAUNCH_CHECK("ffn_fused");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_ffn_fused(const void* x, int64_t rows, int C, const void* w1_bf16, const float* b1, const void* w2_bf16,
                             const float* b2, const float* gamma, const float* beta, float eps, void* y,
                             ltu_stream_t stream) {
    return ffn_launch(x, rows, C, w1_bf16, b1, w2_bf16, b2, gamma, beta, eps, y, nullptr, stream);
}

// Same launch with a pipeline trace: CTA 0 writes clock64() stamps into trace[2][64][8] (int64, device memory;
// role 0 = GELU group leader: wait acc1 | acc1 ready | GELU done | GEMM2 issued; role 1 = LayerNorm group leader:
// wait acc2 | acc2 ready | buffer drained | tile t+2 landed (GEMM1 issued) | stats ready | exchanged | stored).
extern "C" int ltu_ffn_fused_trace(const void* x, int64_t rows, int C, const void* w1_bf16, const float* b1,
                                   const void* w2_bf16, const float* b2, const float* gamma, const float* beta, float eps,
                                   void* y, long long* trace, ltu_stream_t stream) {
    // a trace pointer below 16 is not a pointer but an ablation mode (1: no GELU math, 2: no LayerNorm/stores, 3: both)
    if ((uintptr_t)trace < 64) return ffn_launch(x, rows, C, w1_bf16, b1, w2_bf16, b2, gamma, beta, eps, y, nullptr, stream, (int)(uintptr_t)trace);
    return ffn_launch(x, rows, C, w1_bf16, b1, w2_bf16, b2, gamma, beta, eps, y, trace, stream);
}
