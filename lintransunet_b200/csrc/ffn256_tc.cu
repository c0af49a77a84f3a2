// Fused feed-forward half of SelfAttentionLayer (model/trans_block.py:207-210) for d_model = 256 (bridges 2-4):
//
//     y = LayerNorm2( t + W2 . gelu(W1 . t + b1) + b2 )           t, y : bf16 [rows][256], hidden width 512
//
// Same idea as ffn128_kernel (ffn_tc.cu) -- the hidden activation never leaves the SM, HBM traffic is one read of t
// and one write of y instead of 13 row-units -- but W1 and W2 (256 KB each) do not fit in shared memory, so they are
// STREAMED from L2 through a 4 x 32 KB TMA ring, once per 128-row tile, and the hidden dimension is processed in four
// quarters of 128 so that GELU arithmetic, GEMM1 of the next quarter and GEMM2 of the previous one overlap:
//
//   TMEM   A1[0] = [0,128)  A1[1] = [128,256)   acc1 quarter buffers (fp32) -> packed bf16 H written in place
//          acc2  = [256,512)                     output accumulator [128 x 256]
//   MMA order per tile:  G1(0) G1(1) G2(0) G1(2) G2(1) G1(3) G2(2) G2(3)
//          G1(q): A1[q&1] = T . W1[128q:128q+128, :]^T        K = 256 (A: the resident t tile, B: 2 ring stages)
//          G2(q): acc2   += H_q . W2[:, 128q:128q+128]^T      K = 128 (A: TENSOR MEMORY, B: 2 ring stages)
//   warps 0-15  epilogue, 4 per TMEM lane quarter: GELU of 32 accumulator columns per quarter and thread; at the end
//               of the tile + b2 + residual (re-read from L2) -> LayerNorm over 256 (four 64-column partials per row
//               merged by Chan's formula through smem) -> bf16 -> 16-byte global stores
//   warp 16     TMA producer (t tile: 4 boxes [128 x 64]; ring: W1 stages = 2 boxes [128 x 64], W2 stages = 1 box [256 x 64])
//   warp 17     tcgen05.mma issue (warp-uniform, one elected lane), TMEM allocation
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

constexpr int kF2Threads = 576;
constexpr int kF2Epi = 512;
constexpr int kF2Stages = 4;
constexpr uint32_t kF2StageBytes = 32768;
constexpr uint32_t kF2OffX = 0;                                   // 4 K-slabs of [128 rows x 128 B]
constexpr uint32_t kF2OffRing = 65536;
constexpr uint32_t kF2OffTail = kF2OffRing + kF2Stages * kF2StageBytes;

struct F2Tail {
    uint64_t x_full, x_empty, b_full[kF2Stages], b_empty[kF2Stages], a1_full[2], a1_free[2], h_full[2], acc2_full, acc2_free;
    uint32_t tmem_slot, pad_;
    alignas(16) float b1[512], b2[256], gamma[256], beta[256];      // read as float4
    float2 xs[4][128];              // [column group][row] = (mean, M2) of 64 columns
};

struct F2Params {
    const bf16* x; bf16* y;
    const float* b1; const float* b2; const float* gamma; const float* beta;
    float eps;
    int tiles;
    int64_t rows;
    int cl;                 // CTAs per cluster: 1, or 2 = every weight stage is fetched once per CTA pair (TMA multicast)
};

__global__ void __launch_bounds__(kF2Threads, 1)
ffn256_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w1,
              const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_w2h, const F2Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    F2Tail* tail = reinterpret_cast<F2Tail*>(smem + kF2OffTail);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Cluster mode: the two CTAs of a pair walk tile PAIRS in lock step (same trip count; a tile past the end computes on
    // zero rows and stores nothing) because both must take part in every weight stage.
    const int crank = p.cl == 2 ? (int)cluster_ctarank() : 0;
    const int ngroups = (int)gridDim.x / p.cl, group = (int)blockIdx.x / p.cl;
    const int gtiles = (p.tiles + p.cl - 1) / p.cl;                       // tile groups (pairs)
    const int n_my = (gtiles - group + ngroups - 1) / ngroups;
    auto tile_of = [&](int t) { return (group + t * ngroups) * p.cl + crank; };
    const uint16_t cmask = (uint16_t)((1u << p.cl) - 1);

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&tail->x_full), 1);
        mbar_init(smem_u32(&tail->x_empty), 1);
        for (int s = 0; s < kF2Stages; ++s) { mbar_init(smem_u32(&tail->b_full[s]), 1); mbar_init(smem_u32(&tail->b_empty[s]), (uint32_t)p.cl); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tail->a1_full[s]), 1);
            mbar_init(smem_u32(&tail->a1_free[s]), 1);
            mbar_init(smem_u32(&tail->h_full[s]), kF2Epi);
        }
        mbar_init(smem_u32(&tail->acc2_full), 1);
        mbar_init(smem_u32(&tail->acc2_free), kF2Epi);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 512; i += kF2Threads) tail->b1[i] = p.b1[i];
    for (int i = threadIdx.x; i < 256; i += kF2Threads) { tail->b2[i] = p.b2[i]; tail->gamma[i] = p.gamma[i]; tail->beta[i] = p.beta[i]; }
    if (warp == 17) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tail->tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.cl == 2) cluster_sync_all();                  // the peer's barriers exist before anything is multicast to them
    const uint32_t tmem_base = tail->tmem_slot;

    // order of the eight GEMM phases of a tile: (kind, quarter); kind 0 = G1, 1 = G2
    constexpr int kPhaseKind[8] = {0, 0, 1, 0, 1, 0, 1, 1};
    constexpr int kPhaseQ[8]    = {0, 1, 0, 2, 1, 3, 2, 3};

    if (warp == 16) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            uint32_t n = 0;
            for (int t = 0; t < n_my; ++t) {
                const int row0 = tile_of(t) * 128;
                mbar_wait(smem_u32(&tail->x_empty), (t & 1) ^ 1);            // every G1 of the previous tile has read the t tile
                mbar_expect_tx(smem_u32(&tail->x_full), 65536);
                for (int kb = 0; kb < 4; ++kb) tma_load_2d(sbase + kF2OffX + kb * 16384, &tm_x, kb * 64, row0, smem_u32(&tail->x_full));
#pragma unroll 1
                for (int ph = 0; ph < 8; ++ph) {
                    const int kind = kPhaseKind[ph], q = kPhaseQ[ph];
#pragma unroll 1
                    for (int s2 = 0; s2 < 2; ++s2, ++n) {
                        const int s = n % kF2Stages;
                        const uint32_t dst = sbase + kF2OffRing + s * kF2StageBytes, bar = smem_u32(&tail->b_full[s]);
                        mbar_wait(smem_u32(&tail->b_empty[s]), ((n / kF2Stages) & 1) ^ 1);
                        mbar_expect_tx(bar, kF2StageBytes);
                        if (p.cl == 2) {        // each CTA fetches one 16 KB half of the stage and multicasts it to both
                            if (kind == 0) tma_load_2d_mc(dst + crank * 16384, &tm_w1, (2 * s2 + crank) * 64, q * 128, bar, cmask);
                            else tma_load_2d_mc(dst + crank * 16384, &tm_w2h, q * 128 + s2 * 64, crank * 128, bar, cmask);
                        } else if (kind == 0) { // W1 rows [128q, +128), K-blocks 2*s2 and 2*s2 + 1
                            tma_load_2d(dst, &tm_w1, (2 * s2) * 64, q * 128, bar);
                            tma_load_2d(dst + 16384, &tm_w1, (2 * s2 + 1) * 64, q * 128, bar);
                        } else {                // W2 all 256 rows, hidden columns [128q + 64*s2, +64)
                            tma_load_2d(dst, &tm_w2, q * 128 + s2 * 64, 0, bar);
                        }
                    }
                }
            }
        }
    } else if (warp == 17) {
        // =========================== MMA issuer ===========================
        constexpr uint32_t idesc1 = umma_idesc_bf16(128, 128), idesc2 = umma_idesc_bf16(128, 256);
        uint32_t n = 0;
        uint32_t use[2] = {0, 0};                  // G1 uses of each A1 buffer so far
        uint32_t g2use[2] = {0, 0};
        for (int t = 0; t < n_my; ++t) {
            mbar_wait(smem_u32(&tail->x_full), t & 1);
            tc_fence_after();
#pragma unroll 1
            for (int ph = 0; ph < 8; ++ph) {
                const int kind = kPhaseKind[ph], q = kPhaseQ[ph], b = q & 1;
                const uint32_t a1 = tmem_base + (uint32_t)(b * 128);
                if (kind == 0) {
                    mbar_wait(smem_u32(&tail->a1_free[b]), (use[b] & 1) ^ 1);       // G2 of the buffer's previous quarter is done
                    tc_fence_after();
#pragma unroll 1
                    for (int s2 = 0; s2 < 2; ++s2, ++n) {
                        const int s = n % kF2Stages;
                        mbar_wait(smem_u32(&tail->b_full[s]), (n / kF2Stages) & 1);
                        tc_fence_after();
#pragma unroll
                        for (int kb = 0; kb < 2; ++kb) {
                            const uint64_t adesc = make_desc(sbase + kF2OffX + (2 * s2 + kb) * 16384);
                            const uint64_t bdesc = make_desc(sbase + kF2OffRing + s * kF2StageBytes + kb * 16384);
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                umma_bf16_elect(a1, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc1, (s2 | kb | k) != 0);
                        }
                        if (p.cl == 2) umma_commit_mc_elect(smem_u32(&tail->b_empty[s]), cmask);     // the stage is free when BOTH CTAs have consumed it
                        else umma_commit_elect(smem_u32(&tail->b_empty[s]));
                    }
                    umma_commit_elect(smem_u32(&tail->a1_full[b]));
                    ++use[b];
                    if (q == 3) umma_commit_elect(smem_u32(&tail->x_empty));         // the t tile may be replaced
                } else {
                    if (q == 0) { mbar_wait(smem_u32(&tail->acc2_free), (t & 1) ^ 1); }   // e2 of the previous tile drained acc2
                    mbar_wait(smem_u32(&tail->h_full[b]), g2use[b] & 1);
                    tc_fence_after();
#pragma unroll 1
                    for (int s2 = 0; s2 < 2; ++s2, ++n) {
                        const int s = n % kF2Stages;
                        mbar_wait(smem_u32(&tail->b_full[s]), (n / kF2Stages) & 1);
                        tc_fence_after();
                        const uint64_t bdesc = make_desc(sbase + kF2OffRing + s * kF2StageBytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int kk = s2 * 4 + k;                              // K-step (16 hidden) of the quarter: packed at 32*(kk>>1) + 8*(kk&1)
                            umma_bf16_ts_elect(tmem_base + 256, a1 + (uint32_t)(32 * (kk >> 1) + 8 * (kk & 1)),
                                               bdesc + (uint64_t)(k * 2), idesc2, (q | kk) != 0);
                        }
                        if (p.cl == 2) umma_commit_mc_elect(smem_u32(&tail->b_empty[s]), cmask);     // the stage is free when BOTH CTAs have consumed it
                        else umma_commit_elect(smem_u32(&tail->b_empty[s]));
                    }
                    umma_commit_elect(smem_u32(&tail->a1_free[b]));
                    ++g2use[b];
                    if (q == 3) umma_commit_elect(smem_u32(&tail->acc2_full));
                }
            }
        }
    } else {
        // =========================== epilogue ===========================
        const int ql = warp & 3;                   // TMEM lane quarter (== warp % 4)
        const int cg = warp >> 2;                  // column group 0..3
        const int row = ql * 32 + lane;
        const uint32_t lane_off = (uint32_t)(ql * 32) << 16;
        uint32_t use[2] = {0, 0};
        for (int t = 0; t < n_my; ++t) {
            // ---- e1: GELU of this thread's 32 columns of each hidden quarter, packed H written in place
#pragma unroll 1
            for (int q = 0; q < 4; ++q) {
                const int b = q & 1;
                const uint32_t a1 = tmem_base + lane_off + (uint32_t)(b * 128 + cg * 32);
                mbar_wait(smem_u32(&tail->a1_full[b]), use[b] & 1);
                ++use[b];
                tc_fence_after();
                uint32_t raw[32];
                tmem_ld32_nowait(a1, raw);
                tmem_ld_wait();
                const float4* bv = reinterpret_cast<const float4*>(tail->b1 + q * 128 + cg * 32);
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 bb = bv[j];
                    pk[2 * j] = pack_bf16x2(gelu_erf(__uint_as_float(raw[4 * j]) + bb.x), gelu_erf(__uint_as_float(raw[4 * j + 1]) + bb.y));
                    pk[2 * j + 1] = pack_bf16x2(gelu_erf(__uint_as_float(raw[4 * j + 2]) + bb.z), gelu_erf(__uint_as_float(raw[4 * j + 3]) + bb.w));
                }
                tmem_st16(a1, pk);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&tail->h_full[b]));
            }
            // ---- e2: + b2 + residual -> LayerNorm over 256 columns -> bf16 rows (this thread: columns [64cg, 64cg+64))
            const int64_t grow = (int64_t)tile_of(t) * 128 + row;
            const bool row_ok = grow < p.rows;
            const uint4* rsrc = reinterpret_cast<const uint4*>(p.x + (row_ok ? grow : 0) * 256 + cg * 64);
            mbar_wait_sleep(smem_u32(&tail->acc2_full), t & 1, 64);
            tc_fence_after();
            float y[64];
            float s1 = 0.f;
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
                uint4 res[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) res[j] = __ldg(rsrc + hf * 4 + j);
                uint32_t raw[32];
                tmem_ld32_nowait(tmem_base + lane_off + 256 + (uint32_t)(cg * 64 + hf * 32), raw);
                tmem_ld_wait();
                const float4* b2v = reinterpret_cast<const float4*>(tail->b2 + cg * 64 + hf * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t w[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
                    const float4 ba = b2v[2 * j], bb = b2v[2 * j + 1];
                    const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float lo = __uint_as_float(w[u] << 16), hi = __uint_as_float(w[u] & 0xffff0000u);
                        const float v0 = __uint_as_float(raw[j * 8 + 2 * u]) + bs[2 * u] + lo;
                        const float v1 = __uint_as_float(raw[j * 8 + 2 * u + 1]) + bs[2 * u + 1] + hi;
                        y[hf * 32 + j * 8 + 2 * u] = v0; y[hf * 32 + j * 8 + 2 * u + 1] = v1;
                        s1 += v0 + v1;
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&tail->acc2_free));
            const float m_loc = s1 * (1.f / 64.f);
            float m2 = 0.f;
#pragma unroll
            for (int j = 0; j < 64; ++j) { const float d = y[j] - m_loc; m2 = fmaf(d, d, m2); }
            tail->xs[cg][row] = make_float2(m_loc, m2);
            asm volatile("bar.sync 1, 512;" ::: "memory");
            const float2 p0 = tail->xs[0][row], p1 = tail->xs[1][row], p2 = tail->xs[2][row], p3 = tail->xs[3][row];
            const float mean = 0.25f * ((p0.x + p1.x) + (p2.x + p3.x));
            const float d0 = p0.x - mean, d1 = p1.x - mean, d2 = p2.x - mean, d3 = p3.x - mean;
            // Chan: M2 = sum M2_i + n_i * sum (mean_i - mean)^2, n_i = 64
            const float var = (((p0.y + p1.y) + (p2.y + p3.y)) + 64.f * ((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3))) * (1.f / 256.f);
            const float rstd = rsqrtf(var + p.eps);
            asm volatile("bar.sync 1, 512;" ::: "memory");                 // xs may be rewritten by the next tile
            if (row_ok) {
                uint4* dst = reinterpret_cast<uint4*>(p.y + grow * 256 + cg * 64);
                const float4* gv = reinterpret_cast<const float4*>(tail->gamma + cg * 64);
                const float4* bv = reinterpret_cast<const float4*>(tail->beta + cg * 64);
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
                    dst[j] = ov;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cl == 2) cluster_sync_all();                  // no CTA leaves while its peer may still multicast into it
    if (warp == 17) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// launcher used by ltu_ffn_fused (ffn_tc.cu) for C == 256
int ffn256_launch(const void* x, int64_t rows, const void* w1_bf16, const float* b1, const void* w2_bf16, const float* b2,
                  const float* gamma, const float* beta, float eps, void* y, cudaStream_t stream) {
    CUtensorMap tx, tw1, tw2, tw2h;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tx, x, (uint64_t)rows, 256, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw1, w1_bf16, 512, 256, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw2, w2_bf16, 256, 512, 256)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw2h, w2_bf16, 256, 512, 128)) != LTU_OK) return rc;     // half-stage boxes (cluster mode)
    F2Params p;
    p.x = (const bf16*)x; p.y = (bf16*)y; p.rows = rows;
    p.b1 = b1; p.b2 = b2; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.tiles = (int)((rows + 127) / 128);
    const size_t smem = 1024 + kF2OffTail + sizeof(F2Tail);
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(ffn256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured_dev = dev;
    }
    static const int want_cl = [] { const char* e = getenv("LTU_FFN256_CLUSTER"); return (e && e[0] == '0') ? 1 : 2; }();
    p.cl = (want_cl == 2 && p.tiles >= 2) ? 2 : 1;
    int grid = sm_count() / p.cl * p.cl;
    const int need = (p.tiles + p.cl - 1) / p.cl * p.cl;
    if (grid > need) grid = need;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kF2Threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, ffn256_kernel, tx, tw1, tw2, tw2h, p);
    if (le != cudaSuccess) { set_error("ffn_fused (d_model 256): launch failed: %s", cudaGetErrorString(le)); return (int)le; }
    LTU_LAUNCH_CHECK("ffn_fused (d_model 256)");
    count_launch(1);
    return LTU_OK;
}

}  // namespace ltu
