// Hardware probe, a development tool built by tools/umma_probe.py into tools/libltu_probe.so (NOT part of libltu_b200.so): does a K-major SWIZZLE_128B UMMA operand descriptor accept a start
// address that is not 1024-byte aligned and a stride between 8-row groups (SBO) that is not a multiple of
// 1024 bytes?  This is what a shared-memory HALO needs: the A rows of filter tap (kh,kw,kd) are the halo rows
// shifted by a constant, and consecutive 8-row groups of the 4x4x8 output tile are (8+2) halo rows apart.
//
//   halo  : R rows x 128 B (64 bf16), written by ONE TMA box load with SWIZZLE_128B
//   A     : rows r = 8g + i  <->  halo row  off + g * sbo_rows + i      (M = 128)
//   B     : [64][64] bf16, TMA SWIZZLE_128B
//   out   : fp32 [128][64] = A . B^T
#include <cuda.h>

#include "tc_common.cuh"

namespace ltu {

int make_tmap_bf16_2d_w(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols);

__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_w, float* out,
                  int R, int off_rows, int sbo_rows, int use_base_offset, int CW) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_slot;
    const uint32_t sa = smem_u32(smem), sb = sa + 256 * 128;
    const uint32_t RB = (uint32_t)CW * 2u;                         // row bytes: 128 / 64 / 32 <-> SWIZZLE_128B / 64B / 32B
    const uint64_t layout = CW == 64 ? 2 : (CW == 32 ? 4 : 6);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar_full), 1);
        mbar_init(smem_u32(&bar_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&bar_full), (uint32_t)(R + 64) * RB);
        tma_load_2d(sa, &tm_g, 0, 0, smem_u32(&bar_full));
        tma_load_2d(sb, &tm_w, 0, 0, smem_u32(&bar_full));
        mbar_wait(smem_u32(&bar_full), 0);
        tc_fence_after();
        const uint32_t start = sa + (uint32_t)off_rows * RB;
        uint64_t adesc = 0;
        adesc |= (uint64_t)((start & 0x3FFFF) >> 4);
        adesc |= (uint64_t)1 << 16;
        adesc |= (uint64_t)(((uint32_t)sbo_rows * RB) >> 4) << 32;
        adesc |= (uint64_t)1 << 46;
        if (use_base_offset) adesc |= (uint64_t)((start >> 7) & 7) << 49;
        adesc |= layout << 61;
        uint64_t bdesc = (uint64_t)((sb & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8u * RB) >> 4) << 32) | ((uint64_t)1 << 46) | (layout << 61);
        const uint32_t idesc = umma_idesc_bf16(128, 64);
        for (int k = 0; k < CW / 16; ++k) umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k != 0);
        umma_commit(smem_u32(&bar_done));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 64; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 32; ++i) out[row * 64 + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
    }
}

// Second probe: MN-major operands.  out[128 j][128 c] = P^T . X with P [128 tokens][128 j] and X [128 tokens][128 c], both
// TMA-written as two [128 rows x 64] SWIZZLE_128B blocks (the layout a K-major A operand tile has): here the TOKEN axis is
// the K dimension of the MMA, i.e. both operands are read "transposed" (instruction-descriptor bits 15 / 16).
// Descriptor fields tried by the caller: lbo / sbo bytes, start advance per K-step of 16 tokens.
__global__ void __launch_bounds__(128, 1)
umma_probe_mn_kernel(const __grid_constant__ CUtensorMap tm_p, const __grid_constant__ CUtensorMap tm_x, float* out,
                     int lbo, int sbo, int kadv) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_slot;
    const uint32_t sp = smem_u32(smem), sx = sp + 32768;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar_full), 1);
        mbar_init(smem_u32(&bar_done), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (threadIdx.x == 0) {
        mbar_expect_tx(smem_u32(&bar_full), 65536);
        tma_load_2d(sp, &tm_p, 0, 0, smem_u32(&bar_full));
        tma_load_2d(sp + 16384, &tm_p, 64, 0, smem_u32(&bar_full));
        tma_load_2d(sx, &tm_x, 0, 0, smem_u32(&bar_full));
        tma_load_2d(sx + 16384, &tm_x, 64, 0, smem_u32(&bar_full));
        mbar_wait(smem_u32(&bar_full), 0);
        tc_fence_after();
        auto desc = [&](uint32_t addr) {
            return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((uint32_t)lbo >> 4) << 16) | ((uint64_t)((uint32_t)sbo >> 4) << 32) |
                   ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        };
        const uint32_t idesc = umma_idesc_bf16(128, 128) | (1u << 15) | (1u << 16);      // A and B MN-major
        for (int k = 0; k < 8; ++k) umma_bf16(tmem_base, desc(sp + (uint32_t)(k * kadv)), desc(sx + (uint32_t)(k * kadv)), idesc, k != 0);
        umma_commit(smem_u32(&bar_done));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar_done), 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < 128; c0 += 32) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        for (int i = 0; i < 32; ++i) out[row * 128 + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
    }
}

}  // namespace ltu

using namespace ltu;

// p, x bf16 [128][128]; out fp32 [128][128] = p^T x
extern "C" int ltu_debug_umma_probe_mn(const void* pm, const void* x, float* out, int lbo, int sbo, int kadv, ltu_stream_t stream) {
    CUtensorMap tp, tx;
    int rc;
    if ((rc = make_tmap_bf16_2d_w(&tp, pm, 128, 128, 128, 64)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d_w(&tx, x, 128, 128, 128, 64)) != LTU_OK) return rc;
    const size_t smem = 1024 + 65536;
    cudaFuncSetAttribute(umma_probe_mn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_probe_mn_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tp, tx, out, lbo, sbo, kadv);
    LTU_LAUNCH_CHECK("umma_probe_mn");
    return LTU_OK;
}

// g bf16 [R][CW] (R <= 256), w bf16 [64][CW], out fp32 [128][64]; CW = 64, 32 or 16 channels per row
extern "C" int ltu_debug_umma_probe(const void* g, int R, const void* w, float* out, int off_rows, int sbo_rows,
                                    int use_base_offset, int CW, ltu_stream_t stream) {
    LTU_ARG_CHECK(g && w && out && R > 0 && R <= 256 && (CW == 64 || CW == 32 || CW == 16), "umma_probe: bad arguments");
    LTU_ARG_CHECK(off_rows >= 0 && sbo_rows > 0 && off_rows + 15 * sbo_rows + 8 <= R, "umma_probe: rows out of the halo");
    CUtensorMap tg, tw;
    int rc;
    if ((rc = make_tmap_bf16_2d_w(&tg, g, (uint64_t)R, (uint64_t)CW, (uint32_t)R, (uint32_t)CW)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d_w(&tw, w, 64, (uint64_t)CW, 64, (uint32_t)CW)) != LTU_OK) return rc;
    const size_t smem = 1024 + 256 * 128 + 64 * 128;
    cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tg, tw, out, R, off_rows, sbo_rows, use_base_offset, CW);
    LTU_LAUNCH_CHECK("umma_probe");
    return LTU_OK;
}
