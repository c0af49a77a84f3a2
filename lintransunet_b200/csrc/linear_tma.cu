// nn.Linear of the encoder layers (model/trans_block.py:155-157 Q/K/V, :166 output projection, :208 linear1/linear2)
// with the bias, the exact-erf GELU (:208) or the residual add + LayerNorm (:205-206, :209-210) in the epilogue:
//
//     epi 0   y = x W^T + b                                   (fused QKV / KV projection)
//     epi 1   y = gelu(x W^T + b)                             (linear1)
//     epi 2   (y_hi, y_lo) = split(LayerNorm(x W^T + b + r_hi + r_lo) * gamma + beta)      (output projection, linear2)
//
// Two options carry the query half of linear_attention (:50, :65) through the same launches (ltu_linear_fused_ex):
//   * softmax_cols > 0 (epi 0): output columns [0, softmax_cols) -- the Q third of the QKV projection -- are written as
//     softmax over each head's 32 columns / sqrt(32) (one thread owns a row: the head is 32 of its registers);
//   * samples > 1: W is one [N][K] matrix PER SAMPLE.  With W_b = (blockdiag(ctx_b) Wo^T)^T (ltu_ctx_project) the output
//     projection of softmax(Q) IS the readout followed by the output projection: (P ctx_b) Wo^T = P (ctx_b Wo^T), so
//     q_readout never runs and its result is never written.  Row tiles are then aligned to samples (3-D tensor maps clip).
//
// One persistent, warp-specialised TMA + tcgen05 kernel; every byte that moves between HBM and the SM moves through TMA
// (coalesced by construction), the epilogue warps only touch tensor memory and shared memory:
//
//   grid = min(#tiles, #SMs) CTAs of 320 threads, a CTA walks its 128 x 256 output tiles T, T+grid, ... with the
//   n-tile index fastest, so that the CTAs working on one 128-row block of x at the same time share it through L2.
//     warp 0      producer: one lane issues cp.async.bulk.tensor.2d loads of x [128 x 64] and W [256 x 64] k-blocks
//                 (SWIZZLE_128B = the K-major UMMA operand layout) into a 3-stage ring (48 KB per stage)
//     warp 1      MMA issue (warp-uniform descriptor arithmetic, one elected lane): tcgen05.mma M128 N256 K16 into one of
//                 TWO 256-column TMEM accumulators, so the epilogue of tile i overlaps the loads and MMAs of tile i+1
//     warps 2-9   epilogue, two warps per TMEM lane quarter (128 columns each), one row per thread:
//                 tcgen05.ld -> bias -> (GELU | LayerNorm) -> bf16 -> per-warp swizzled staging tile -> TMA store
//   The residual of epi 2 is NOT read by the epilogue threads (one row per thread = one 128-byte line per lane and
//   instruction: the L1 wavefronts alone would cost as much as the tile's HBM time).  It rides the same TMA ring as extra
//   A-operand k-blocks and is added ON THE TENSOR PIPE:  acc[:, 64j:64j+64] += R[:, 64j:64j+64] . I_64  (N = 64 MMAs
//   against a shared-memory identity; products with 1.0 and 0.0 are exact, the accumulator is fp32), for r_hi and r_lo --
//   the split bf16 token stream of the bf16 path (include/ltu_b200.h, ltu_add_layernorm_split).  LayerNorm statistics:
//   per-thread (mean, M2) over 128 columns, merged with the row partner by Chan's formula.
#include <cuda.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);   // ffn_tc.cu
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint64_t ld = 0);

constexpr int kLinThreads = 320;
constexpr int kLinStages = 3;
constexpr int kLinBN = 256;
constexpr uint32_t kLinABytes = 128 * 128;                 // x k-block: 128 rows x 64 bf16
constexpr uint32_t kLinBBytes = kLinBN * 128;              // W k-block: 256 rows x 64 bf16
constexpr uint32_t kLinStageBytes = kLinABytes + kLinBBytes;
constexpr uint32_t kLinOffEye = kLinStages * kLinStageBytes;
constexpr uint32_t kLinOffStaging = kLinOffEye + 64 * 128;
constexpr uint32_t kLinOffTail = kLinOffStaging + 8 * 8192;
constexpr int kLinMaxN = 768;

enum : int { kLinBias = 0, kLinGelu = 1, kLinResLN = 2 };

struct LinTail {
    uint64_t full[kLinStages], empty[kLinStages], tfull[2], tempty[2];
    uint32_t tmem_slot, pad_;
    alignas(16) float bias[kLinMaxN];
    alignas(16) float gamma[kLinBN], beta[kLinBN];
    float2 xs[2][2][128];            // [accumulator][column half][row] = (mean, M2) of 128 columns
};

struct LinParams {
    const float* bias; const float* gamma; const float* beta;
    float eps;
    int N, nkb;                      // output columns, 64-wide k-blocks of x
    int tiles_m, tiles_n;
    int epi, has_lo, want_lo;
    int cl;                          // CTAs per cluster: 2 = a CTA pair walks two row blocks of the same n-tile in lock step
                                     // and every W k-block is fetched ONCE per pair (each CTA loads half, TMA multicast)
    int mode;                        // debug ablations (LTU_LIN_MODE): 1 no output stores, 2 no GELU math, 4 no MMAs, 8 no W loads
    int tps;                         // 128-row tiles per sample (x, residual and y are [samples][rows per sample][cols] maps)
    int w_rows;                      // rows between the weight matrices of two samples (0: one weight for all)
    int qsm_tiles;                   // n-tiles [0, qsm_tiles) are written as per-head softmax / sqrt(32) (epi 0)
};

// first row (inside its sample) and sample of m-tile `mt`
struct LinTile { int r0, b; };
__device__ __forceinline__ LinTile lin_tile(int mt, int tps) { LinTile t; t.b = mt / tps; t.r0 = (mt - t.b * tps) * 128; return t; }

__global__ void __launch_bounds__(kLinThreads, 1)
linear_tma_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                  const __grid_constant__ CUtensorMap tm_wh, const __grid_constant__ CUtensorMap tm_rhi, const __grid_constant__ CUtensorMap tm_rlo,
                  const __grid_constant__ CUtensorMap tm_yhi, const __grid_constant__ CUtensorMap tm_ylo,
                  const LinParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    LinTail* tail = reinterpret_cast<LinTail*>(smem + kLinOffTail);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Work items: (row-block group, n-tile), n fastest; a group is p.cl consecutive 128-row blocks, one per CTA of the
    // cluster.  Both CTAs of a pair run the same trip count (a row block past the end loads zeros and stores nothing).
    const int crank = p.cl == 2 ? (int)cluster_ctarank() : 0;
    const int ngroups = (int)gridDim.x / p.cl, group = (int)blockIdx.x / p.cl;
    const int total = (p.tiles_m + p.cl - 1) / p.cl * p.tiles_n;
    const uint16_t cmask = (uint16_t)((1u << p.cl) - 1);
    const int nres = p.epi == kLinResLN ? (p.has_lo ? 8 : 4) : 0;      // residual k-blocks (r_hi then r_lo)
    const int nkb_all = p.nkb + (nres + 2) / 3;                       // ring stages per tile: three residual k-blocks share a stage

    if (threadIdx.x == 0) {
        for (int s = 0; s < kLinStages; ++s) { mbar_init(smem_u32(&tail->full[s]), 1); mbar_init(smem_u32(&tail->empty[s]), (uint32_t)p.cl); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tail->tfull[i]), 1); mbar_init(smem_u32(&tail->tempty[i]), 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < p.N; i += kLinThreads) tail->bias[i] = p.bias[i];
    if (p.epi == kLinResLN)
        for (int i = threadIdx.x; i < kLinBN; i += kLinThreads) { tail->gamma[i] = p.gamma[i]; tail->beta[i] = p.beta[i]; }
    // I_64 as a K-major SWIZZLE_128B operand: element (n, k) at n*128 + ((k/8) ^ (n%8))*16 + (k%8)*2
    for (int i = threadIdx.x; i < 64 * 8; i += kLinThreads) {
        const int n = i >> 3, c = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if ((n >> 3) == c) {
            const uint32_t one = 0x3f80u << (16 * (n & 1));
            const int w = (n & 7) >> 1;
            v.x = w == 0 ? one : 0u; v.y = w == 1 ? one : 0u; v.z = w == 2 ? one : 0u; v.w = w == 3 ? one : 0u;
        }
        *reinterpret_cast<uint4*>(smem + kLinOffEye + n * 128 + ((c ^ (n & 7)) << 4)) = v;
    }
    fence_async_smem();
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tail->tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (p.cl == 2) cluster_sync_all();                  // the peer's barriers exist before anything is multicast to them
    const uint32_t tmem_base = tail->tmem_slot;

    if (warp == 0) {
        // =========================== producer ===========================
        if (lane == 0) {
            pdl_prologue();                                              // inputs come from the previous kernel in the stream
            uint32_t kbc = 0;
            for (int T = group; T < total; T += ngroups) {
                const LinTile tl = lin_tile((T / p.tiles_n) * p.cl + crank, p.tps);
                const int n0 = (T % p.tiles_n) * kLinBN, wrow = n0 + tl.b * p.w_rows;
                for (int kb = 0; kb < nkb_all; ++kb, ++kbc) {
                    const int stage = kbc % kLinStages;
                    mbar_wait(smem_u32(&tail->empty[stage]), ((kbc / kLinStages) & 1) ^ 1);     // free in BOTH CTAs of a pair
                    const uint32_t fb = smem_u32(&tail->full[stage]);
                    const uint32_t sa = sbase + stage * kLinStageBytes;
                    if (kb < p.nkb) {
                        mbar_expect_tx(fb, (p.mode & 8) ? kLinABytes : kLinStageBytes);
                        tma_load_3d(sa, &tm_x, kb * 64, tl.r0, tl.b, fb);
                        if (p.mode & 8) {
                        } else if (p.cl == 2) {         // this CTA's 128-row half of the W k-block, delivered to both CTAs
                            tma_load_2d_mc(sa + kLinABytes + (uint32_t)crank * (kLinBBytes / 2), &tm_wh, kb * 64, wrow + crank * 128, fb, cmask);
                        } else {
                            tma_load_2d(sa + kLinABytes, &tm_w, kb * 64, wrow, fb);
                        }
                    } else {                                             // up to three residual k-blocks fill one stage
                        const int j0 = (kb - p.nkb) * 3;
                        const int cnt = nres - j0 < 3 ? nres - j0 : 3;
                        mbar_expect_tx(fb, (uint32_t)cnt * kLinABytes);
                        for (int i = 0; i < cnt; ++i) {
                            const int j = j0 + i;
                            tma_load_3d(sa + (uint32_t)i * kLinABytes, j < 4 ? &tm_rhi : &tm_rlo, (j & 3) * 64, tl.r0, tl.b, fb);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issue ===========================
        constexpr uint32_t idesc = umma_idesc_bf16(128, kLinBN), idesc_eye = umma_idesc_bf16(128, 64);
        const uint64_t eye = make_desc(sbase + kLinOffEye);
        uint32_t kbc = 0, it = 0;
        for (int T = group; T < total; T += ngroups, ++it) {
            const uint32_t abuf = it & 1;
            mbar_wait(smem_u32(&tail->tempty[abuf]), ((it >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tacc = tmem_base + abuf * 256u;
            for (int kb = 0; kb < nkb_all; ++kb, ++kbc) {
                const int stage = kbc % kLinStages;
                mbar_wait(smem_u32(&tail->full[stage]), (kbc / kLinStages) & 1);
                tc_fence_after();
                const uint32_t sa = sbase + stage * kLinStageBytes;
                const uint64_t adesc = make_desc(sa);
                if (p.mode & 4) {
                } else if (kb < p.nkb) {
                    const uint64_t bdesc = make_desc(sa + kLinABytes);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_elect(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                } else {
                    const int j0 = (kb - p.nkb) * 3;
                    const int cnt = nres - j0 < 3 ? nres - j0 : 3;
                    for (int i = 0; i < cnt; ++i) {
                        const uint32_t dcol = (uint32_t)(((j0 + i) & 3) * 64);
                        const uint64_t rdesc = make_desc(sa + (uint32_t)i * kLinABytes);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            umma_bf16_elect(tacc + dcol, rdesc + (uint64_t)(k * 2), eye + (uint64_t)(k * 2), idesc_eye, 1u);
                    }
                }
                if (p.cl == 2) umma_commit_mc_elect(smem_u32(&tail->empty[stage]), cmask);     // arrives in both CTAs
                else umma_commit_elect(smem_u32(&tail->empty[stage]));
                if (kb == nkb_all - 1) umma_commit_elect(smem_u32(&tail->tfull[abuf]));
            }
        }
    } else {
        // =========================== epilogue ===========================
        const int e = warp - 2;
        const int q = warp & 3;                          // TMEM lane quarter this warp may access (warp id % 4)
        const int hh = e >> 2;                           // column half [128hh, 128hh + 128)
        const int row = q * 32 + lane;
        unsigned char* stg = smem + kLinOffStaging + e * 8192;
        const uint32_t stg_u = smem_u32(stg);
        const uint32_t swz = (uint32_t)(lane & 7);
        uint32_t it = 0;
        for (int T = group; T < total; T += ngroups, ++it) {
            const LinTile tl = lin_tile((T / p.tiles_n) * p.cl + crank, p.tps);
            const int n0 = (T % p.tiles_n) * kLinBN;
            const bool qsm = (T % p.tiles_n) < p.qsm_tiles;
            const uint32_t abuf = it & 1;
            const uint32_t tb = tmem_base + abuf * 256u + ((uint32_t)(q * 32) << 16) + (uint32_t)(hh * 128);
            const float* bs = tail->bias + n0 + hh * 128;
            mbar_wait_sleep(smem_u32(&tail->tfull[abuf]), (it >> 1) & 1, 64);
            tc_fence_after();
            if (p.epi != kLinResLN) {
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32_nowait(tb + (uint32_t)(c * 64), v0);
                    tmem_ld32_nowait(tb + (uint32_t)(c * 64 + 32), v1);
                    // the store issued from this staging buffer one tile ago must have been read out
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    tmem_ld_wait();
                    if (c == 1) {                         // last read of this accumulator
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&tail->tempty[abuf]));
                    }
                    unsigned char* buf = stg + c * 4096 + lane * 128;
                    if (qsm) {
                        // the thread's 64 columns are two complete heads: softmax over 32 registers, scaled by 1/sqrt(32)
#pragma unroll
                        for (int hx = 0; hx < 2; ++hx) {
                            const uint32_t* src = hx ? v1 : v0;
                            const float4* b4 = reinterpret_cast<const float4*>(bs + c * 64 + hx * 32);
                            float o[32];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const float4 b = b4[i];
                                o[4 * i] = __uint_as_float(src[4 * i]) + b.x; o[4 * i + 1] = __uint_as_float(src[4 * i + 1]) + b.y;
                                o[4 * i + 2] = __uint_as_float(src[4 * i + 2]) + b.z; o[4 * i + 3] = __uint_as_float(src[4 * i + 3]) + b.w;
                            }
                            float m = o[0];
#pragma unroll
                            for (int i = 1; i < 32; ++i) m = fmaxf(m, o[i]);
                            const float mneg = -m * 1.4426950408889634f;
                            float sum = 0.f;
#pragma unroll
                            for (int i = 0; i < 32; ++i) {
                                float ex;
                                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(o[i], 1.4426950408889634f, mneg)));
                                o[i] = ex;
                                sum += ex;
                            }
                            const float inv = 0.17677669529663687f / sum;            // 1 / (sqrt(32) * sum)
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj) {
                                uint4 ov;
                                ov.x = pack_bf16x2(o[8 * jj] * inv, o[8 * jj + 1] * inv); ov.y = pack_bf16x2(o[8 * jj + 2] * inv, o[8 * jj + 3] * inv);
                                ov.z = pack_bf16x2(o[8 * jj + 4] * inv, o[8 * jj + 5] * inv); ov.w = pack_bf16x2(o[8 * jj + 6] * inv, o[8 * jj + 7] * inv);
                                *reinterpret_cast<uint4*>(buf + (((uint32_t)(hx * 4 + jj) ^ swz) << 4)) = ov;
                            }
                        }
                    } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t* src = j < 4 ? v0 + 8 * j : v1 + 8 * (j - 4);
                        const float4 ba = *reinterpret_cast<const float4*>(bs + c * 64 + j * 8);
                        const float4 bb = *reinterpret_cast<const float4*>(bs + c * 64 + j * 8 + 4);
                        float o[8] = {__uint_as_float(src[0]) + ba.x, __uint_as_float(src[1]) + ba.y,
                                      __uint_as_float(src[2]) + ba.z, __uint_as_float(src[3]) + ba.w,
                                      __uint_as_float(src[4]) + bb.x, __uint_as_float(src[5]) + bb.y,
                                      __uint_as_float(src[6]) + bb.z, __uint_as_float(src[7]) + bb.w};
                        if (p.epi == kLinGelu && !(p.mode & 2)) {
#pragma unroll
                            for (int u = 0; u < 8; u += 2) gelu_erf_pair(o[u], o[u + 1], o[u], o[u + 1]);
                        }
                        uint4 ov;
                        ov.x = pack_bf16x2(o[0], o[1]); ov.y = pack_bf16x2(o[2], o[3]);
                        ov.z = pack_bf16x2(o[4], o[5]); ov.w = pack_bf16x2(o[6], o[7]);
                        *reinterpret_cast<uint4*>(buf + (((uint32_t)j ^ swz) << 4)) = ov;
                    }
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0 && !(p.mode & 1)) {
                        tma_store_3d(&tm_yhi, stg_u + c * 4096, n0 + hh * 128 + c * 64, tl.r0 + q * 32, tl.b);
                        tma_store_commit();
                    }
                }
            } else {
                // pass 1 / 2: mean and M2 of this thread's 128 columns of acc + bias (the residual is already in acc)
                float s1 = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    tmem_ld32(tb + (uint32_t)(c * 32), v);
                    const float4* b4 = reinterpret_cast<const float4*>(bs + c * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b4[i];
                        s1 += (v[4 * i] + b.x) + (v[4 * i + 1] + b.y) + (v[4 * i + 2] + b.z) + (v[4 * i + 3] + b.w);
                    }
                }
                const float m_loc = s1 * (1.f / 128.f);
                float m2 = 0.f;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    float v[32];
                    tmem_ld32(tb + (uint32_t)(c * 32), v);
                    const float4* b4 = reinterpret_cast<const float4*>(bs + c * 32);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b = b4[i];
                        const float d0 = v[4 * i] + b.x - m_loc, d1 = v[4 * i + 1] + b.y - m_loc;
                        const float d2 = v[4 * i + 2] + b.z - m_loc, d3 = v[4 * i + 3] + b.w - m_loc;
                        m2 = fmaf(d0, d0, m2); m2 = fmaf(d1, d1, m2); m2 = fmaf(d2, d2, m2); m2 = fmaf(d3, d3, m2);
                    }
                }
                tail->xs[abuf][hh][row] = make_float2(m_loc, m2);
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float2 other = tail->xs[abuf][hh ^ 1][row];
                const float mean = 0.5f * (m_loc + other.x);
                const float dm = m_loc - other.x;
                const float var = (m2 + other.y + dm * dm * 64.f) * (1.f / 256.f);     // Chan: n_a n_b / (n_a + n_b) = 64
                const float rstd = rsqrtf(var + p.eps);
                const float* gm = tail->gamma + hh * 128;
                const float* bt = tail->beta + hh * 128;
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t v0[32], v1[32];
                    tmem_ld32_nowait(tb + (uint32_t)(c * 64), v0);
                    tmem_ld32_nowait(tb + (uint32_t)(c * 64 + 32), v1);
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // hi / lo buffers free
                    __syncwarp();
                    tmem_ld_wait();
                    if (c == 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(smem_u32(&tail->tempty[abuf]));
                    }
                    unsigned char* bhi = stg + lane * 128;
                    unsigned char* blo = stg + 4096 + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const uint32_t* src = j < 4 ? v0 + 8 * j : v1 + 8 * (j - 4);
                        float h[8], l[8];
                        const int col0 = c * 64 + j * 8;
                        const float4 ba = *reinterpret_cast<const float4*>(bs + col0), bb = *reinterpret_cast<const float4*>(bs + col0 + 4);
                        const float4 ga = *reinterpret_cast<const float4*>(gm + col0), gb = *reinterpret_cast<const float4*>(gm + col0 + 4);
                        const float4 ea = *reinterpret_cast<const float4*>(bt + col0), eb = *reinterpret_cast<const float4*>(bt + col0 + 4);
                        const float b8[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
                        const float g8[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
                        const float e8[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const float o = fmaf((__uint_as_float(src[u]) + b8[u] - mean) * rstd, g8[u], e8[u]);
                            h[u] = __bfloat162float(__float2bfloat16_rn(o));
                            l[u] = o - h[u];
                        }
                        uint4 hv, lv;
                        hv.x = pack_bf16x2(h[0], h[1]); hv.y = pack_bf16x2(h[2], h[3]);
                        hv.z = pack_bf16x2(h[4], h[5]); hv.w = pack_bf16x2(h[6], h[7]);
                        lv.x = pack_bf16x2(l[0], l[1]); lv.y = pack_bf16x2(l[2], l[3]);
                        lv.z = pack_bf16x2(l[4], l[5]); lv.w = pack_bf16x2(l[6], l[7]);
                        *reinterpret_cast<uint4*>(bhi + (((uint32_t)j ^ swz) << 4)) = hv;
                        *reinterpret_cast<uint4*>(blo + (((uint32_t)j ^ swz) << 4)) = lv;
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0 && !(p.mode & 1)) {
                        tma_store_3d(&tm_yhi, stg_u, hh * 128 + c * 64, tl.r0 + q * 32, tl.b);
                        if (p.want_lo) tma_store_3d(&tm_ylo, stg_u + 4096, hh * 128 + c * 64, tl.r0 + q * 32, tl.b);
                        tma_store_commit();
                    }
                }
            }
        }
        if (lane == 0) tma_store_wait_all();
    }
    tc_fence_before();
    __syncthreads();
    if (p.cl == 2) cluster_sync_all();                  // no CTA leaves while its peer may still multicast into it
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace ltu

using namespace ltu;

// y = epi(x W^T + b):  x bf16 [rows][K] (K % 64 == 0, K <= 1024), w_bf16 = the nn.Linear weight [N][K] rounded to bf16
// (row-major, K innermost), bias fp32 [N], N % 256 == 0 and N <= 768; y_hi bf16 [rows][N].
//   epi 0: bias | 1: bias + exact-erf GELU | 2: LayerNorm(x W^T + b + res_hi + res_lo) * gamma + beta with N == 256,
//   res_hi / res_lo bf16 [rows][256] (res_lo may be null), result split into y_hi + y_lo (y_lo may be null).
//   ldx: elements between two rows of x (>= K; a column slice of wider rows is read in place).
//   softmax_cols (epi 0; 0 or a multiple of 256): columns [0, softmax_cols) are written as softmax over each group of 32
//   columns, divided by sqrt(32).  samples > 1: rows = samples x (rows / samples) tokens and w_bf16 is [samples][N][K].
extern "C" int ltu_linear_fused_ex(const void* x, int64_t ldx, int64_t rows, int K, const void* w_bf16, const float* bias, int N, int epi,
                                   const void* res_hi, const void* res_lo, const float* gamma, const float* beta, float eps,
                                   void* y_hi, void* y_lo, int softmax_cols, int samples, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && w_bf16 && bias && y_hi, "linear_fused: null pointer");
    LTU_ARG_CHECK(rows > 0 && rows < ((int64_t)1 << 31) - 256, "linear_fused: bad row count");
    LTU_ARG_CHECK(K >= 64 && K <= 1024 && K % 64 == 0, "linear_fused: K must be a multiple of 64 in [64,1024] (got %d)", K);
    LTU_ARG_CHECK(N >= kLinBN && N <= kLinMaxN && N % kLinBN == 0, "linear_fused: N must be 256, 512 or 768 (got %d)", N);
    LTU_ARG_CHECK(epi >= 0 && epi <= 2, "linear_fused: bad epilogue %d", epi);
    LTU_ARG_CHECK(epi != kLinResLN || (N == kLinBN && res_hi && gamma && beta),
                  "linear_fused: the LayerNorm epilogue needs N == 256, a residual, gamma and beta");
    LTU_ARG_CHECK(epi == kLinResLN || (!res_hi && !res_lo && !y_lo), "linear_fused: residual / y_lo only with epi 2");
    LTU_ARG_CHECK((((uintptr_t)x | (uintptr_t)w_bf16 | (uintptr_t)y_hi | (uintptr_t)y_lo | (uintptr_t)res_hi | (uintptr_t)res_lo) & 15) == 0,
                  "linear_fused: pointers must be 16-byte aligned");
    LTU_ARG_CHECK(softmax_cols >= 0 && softmax_cols <= N && softmax_cols % kLinBN == 0 && (softmax_cols == 0 || epi == kLinBias),
                  "linear_fused: softmax_cols must be a multiple of 256 within N, with epi 0 (got %d)", softmax_cols);
    LTU_ARG_CHECK(ldx >= K && ldx % 8 == 0, "linear_fused: ldx (%lld) must be >= K and a multiple of 8", (long long)ldx);
    LTU_ARG_CHECK(samples >= 1 && rows % samples == 0, "linear_fused: rows (%lld) must be a multiple of samples (%d)", (long long)rows, samples);
    const uint64_t nper = (uint64_t)(rows / samples);
    CUtensorMap tx, tw, twh, trh, trl, tyh, tyl;
    int rc;
    if ((rc = make_tmap_bf16_3d(&tx, x, (uint64_t)samples, nper, (uint64_t)K, 128, (uint64_t)ldx)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw, w_bf16, (uint64_t)N * samples, (uint64_t)K, kLinBN)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&twh, w_bf16, (uint64_t)N * samples, (uint64_t)K, kLinBN / 2)) != LTU_OK) return rc;   // half boxes (pairs)
    if ((rc = make_tmap_bf16_3d(&tyh, y_hi, (uint64_t)samples, nper, (uint64_t)N, 32)) != LTU_OK) return rc;
    trh = tx; trl = tx; tyl = tyh;
    if (res_hi && (rc = make_tmap_bf16_3d(&trh, res_hi, (uint64_t)samples, nper, kLinBN, 128)) != LTU_OK) return rc;
    if (res_lo && (rc = make_tmap_bf16_3d(&trl, res_lo, (uint64_t)samples, nper, kLinBN, 128)) != LTU_OK) return rc;
    if (y_lo && (rc = make_tmap_bf16_3d(&tyl, y_lo, (uint64_t)samples, nper, (uint64_t)N, 32)) != LTU_OK) return rc;
    LinParams p;
    p.bias = bias; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.N = N; p.nkb = K / 64;
    p.tps = (int)((nper + 127) / 128);
    p.tiles_m = p.tps * samples; p.tiles_n = N / kLinBN;
    p.w_rows = samples > 1 ? N : 0;
    p.qsm_tiles = softmax_cols / kLinBN;
    p.epi = epi; p.has_lo = res_lo != nullptr; p.want_lo = y_lo != nullptr;
    static const int dbg_mode = [] { const char* e = getenv("LTU_LIN_MODE"); return e ? atoi(e) : 0; }();
    p.mode = dbg_mode;
    const size_t smem = 1024 + kLinOffTail + sizeof(LinTail);
    static thread_local int configured_dev = -1;
    static thread_local int max_pairs = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(linear_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured_dev = dev;
        max_pairs = -1;
    }
    // CTA pairs (2-CTA clusters, W fetched once per pair) when the device can keep every pair resident at once
    // Measured (tools/linear_probe.py, profiles/r2_linear_probe.md): correct, but no faster than independent CTAs -- these
    // GEMMs sit at the HBM write rate, not at the L2 -> SM rate of the W re-reads -- so pairs are opt-in (LTU_LIN_CLUSTER=1).
    static const int want_cl = [] { const char* e = getenv("LTU_LIN_CLUSTER"); return (e && e[0] == '1') ? 2 : 1; }();
    if (want_cl == 2 && max_pairs < 0) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3((unsigned)(sm_count() / 2 * 2)); q.blockDim = dim3(kLinThreads); q.dynamicSmemBytes = smem;
        cudaLaunchAttribute a[1];
        a[0].id = cudaLaunchAttributeClusterDimension;
        a[0].val.clusterDim.x = 2; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
        q.attrs = a; q.numAttrs = 1;
        int n = 0;
        max_pairs = cudaOccupancyMaxActiveClusters(&n, linear_tma_kernel, &q) == cudaSuccess ? n : 0;
        cudaGetLastError();
    }
    p.cl = (want_cl == 2 && p.tiles_m >= 2 && max_pairs >= 16 && samples == 1) ? 2 : 1;
    const int groups = (p.tiles_m + p.cl - 1) / p.cl * p.tiles_n;
    int grid = p.cl == 2 ? max_pairs * 2 : sm_count();
    if (grid > groups * p.cl) grid = groups * p.cl;
    static const bool pdl = [] { const char* e = getenv("LTU_PDL"); return !(e && e[0] == '0'); }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kLinThreads); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (pdl) { attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[na].val.programmaticStreamSerializationAllowed = 1; ++na; }
    if (p.cl == 2) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1; ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = (unsigned)na;
    cudaError_t e = cudaLaunchKernelEx(&cfg, linear_tma_kernel, tx, tw, twh, trh, trl, tyh, tyl, p);
    if (e != cudaSuccess) { set_error("linear_fused: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_linear_fused(const void* x, int64_t rows, int K, const void* w_bf16, const float* bias, int N, int epi,
                                const void* res_hi, const void* res_lo, const float* gamma, const float* beta, float eps,
                                void* y_hi, void* y_lo, ltu_stream_t stream) {
    return ltu_linear_fused_ex(x, K, rows, K, w_bf16, bias, N, epi, res_hi, res_lo, gamma, beta, eps, y_hi, y_lo, 0, 1, stream);
}

namespace ltu {
// out[b][n][32h + j] = sum_e ctx[b][h][j][e] * Wo[n][32h + e]: the output projection with the sample's context folded in
__global__ void __launch_bounds__(1024)
ctx_project_kernel(const float* __restrict__ ctx, const bf16* __restrict__ wo, bf16* __restrict__ out, int C) {
    __shared__ float cs[32][33];                     // cs[e][j]
    __shared__ __align__(16) float ws_[256][32];     // ws_[n][e] = wo[n][32 h + e]
    const int h = blockIdx.x, b = blockIdx.y, heads = gridDim.x;
    const int wrp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = wrp; n < C; n += 32) ws_[n][lane] = __bfloat162float(wo[(int64_t)n * C + 32 * h + lane]);   // a parameter: no wait
    pdl_prologue();
    cs[lane][wrp] = ctx[(((int64_t)b * heads + h) * 32 + wrp) * 32 + lane];
    __syncthreads();
    float c[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) c[k] = cs[k][lane];
    for (int n = wrp; n < C; n += 32) {
        const float4* w4 = reinterpret_cast<const float4*>(ws_[n]);
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float4 w = w4[k];
            acc = fmaf(w.x, c[4 * k], acc); acc = fmaf(w.y, c[4 * k + 1], acc);
            acc = fmaf(w.z, c[4 * k + 2], acc); acc = fmaf(w.w, c[4 * k + 3], acc);
        }
        out[((int64_t)b * C + n) * C + 32 * h + lane] = __float2bfloat16_rn(acc);
    }
}
}  // namespace ltu

// ctx fp32 [B][heads][32][32] (ltu_kv_reduce), wo_bf16 [C][C] (C = 32 heads) -> out bf16 [B][C][C], the per-sample weight
// W_b[n][32h + j] = sum_e ctx[b][h][j][e] Wo[n][32h + e] of ltu_linear_fused_ex(samples = B)
extern "C" int ltu_ctx_project(const float* ctx, const void* wo_bf16, void* out, int B, int heads, ltu_stream_t stream) {
    LTU_ARG_CHECK(ctx && wo_bf16 && out && B > 0 && heads > 0 && heads <= 8, "ctx_project: bad arguments (heads <= 8)");
    cudaError_t e = launch_pdl(ctx_project_kernel, dim3(heads, B), dim3(1024), 0, (cudaStream_t)stream, ctx, (const bf16*)wo_bf16,
                               (bf16*)out, heads * 32);
    if (e != cudaSuccess) { set_error("ctx_project: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    count_launch(1);
    return LTU_OK;
}
