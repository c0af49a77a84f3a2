// Query half of the linear-attention encoder layer for d_model = 128 (4 heads x 32), fused into ONE
// persistent, warp-specialised tcgen05 kernel (model/trans_block.py:50,:65 readout, :155-166 projections,
// :205-206 residual + layer_norm1):
//
//     Q   = x Wq^T + bq                                   [rows x 128]
//     P   = softmax over each head's 32 columns of Q, scaled by 1/sqrt(32)       (bf16, tensor memory)
//     att = P . blockdiag(ctx_b)                          ctx_b[h][j][e] = softmax_N(K)^T V of sample b
//     y   = LayerNorm1( x + att Wo^T + bo )
//
// The separate path runs the Q third of the QKV GEMM, q_readout, the output projection and add_layernorm
// (6 row-units of HBM traffic: Q w1, readout r1+w1, projection r1+w1, LayerNorm r2+w1 ...); this kernel reads x
// once and writes y once.  Three chained GEMMs per 128-row tile, intermediates only in tensor memory:
//
//   warps 0-7   group A: (1) per-head softmax of Q -> bf16 pairs written over the thread's own columns,
//                        (2) att fp32 -> bf16 pairs, again in place
//   warps 8-15  group B: + bo + residual (re-read from L2) -> LayerNorm (Chan merge with the row partner) -> y
//   TMA (warp 0 lane 0): per tile the x tile (2 x [128 x 64] boxes, SWIZZLE_128B) and the sample's ctx operand
//                        ([128 x 64] box: row h*32+e holds ctx[h][0..31][e] and 32 zeros); Wq / Wo once per CTA
//   MMA issue (single lanes of group-B warps, off the softmax warps' critical path):
//       GQ   R1[128x128]        = X . Wq^T                 A, B from smem
//       GATT R2[128x32] x 4     = P_h . ctx_h              A from TENSOR MEMORY, one N=32 GEMM per head
//       GO   R1[128x128]        = att . Wo^T               A from TENSOR MEMORY
//   TMEM: two 256-column buffers (tile parity), each R1 = [0,128) and R2 = [128,256).
//
// Tiles never straddle samples (tile T -> sample T / tiles_per_sample); the rows of a ragged last tile that
// belong to the next sample are computed and discarded (every op is row-wise).  Work of a tile PAIR (a, b):
// group A  softmax(a) softmax(b) convert(a) convert(b);  group B  LayerNorm(a) LayerNorm(b).
#include <cuda.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint64_t ld = 0);

constexpr int kAoThreads = 512;
constexpr int kAoGroup = 256;
constexpr uint32_t kAoWBytes = 128 * 128 * 2;
constexpr uint32_t kAoXBytes = 128 * 128 * 2;
constexpr uint32_t kAoCBytes = 128 * 64 * 2;
constexpr uint32_t kAoOffWq = 0;
constexpr uint32_t kAoOffWo = kAoOffWq + kAoWBytes;
constexpr uint32_t kAoOffX = kAoOffWo + kAoWBytes;
constexpr uint32_t kAoOffC = kAoOffX + 2 * kAoXBytes;
constexpr uint32_t kAoOffTail = kAoOffC + 2 * kAoCBytes;

struct AoTail {
    uint64_t w_full, x_full[2], c_full[2], q_full[2], p_full[2], att_full[2], a_full[2], o_full[2], r1_free[2];
    uint32_t tmem_slot, pad_;
    alignas(16) float bq[128], bo[128], gamma[128], beta[128];       // read as float4
    float2 xs[2][2][128];           // [tile parity][column half][row] = (mean, M2)
};

struct AoParams {
    const bf16* x; bf16* y;
    const float* bq; const float* bo; const float* gamma; const float* beta;
    float eps;
    int tiles, tps;                 // total tiles, tiles per sample
    int64_t N;                      // tokens per sample
    long long* trace;               // debug (LTU_AO_TRACE_PTR, attn_out128w only): [4 roles][64 tiles][8 events] clock64 stamps of CTA 0
};

__global__ void __launch_bounds__(kAoThreads, 1)
attn_out128_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wq,
                   const __grid_constant__ CUtensorMap tm_wo, const __grid_constant__ CUtensorMap tm_ctx, const AoParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    AoTail* tail = reinterpret_cast<AoTail*>(smem + kAoOffTail);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&tail->w_full), 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tail->x_full[s]), 1);
            mbar_init(smem_u32(&tail->c_full[s]), 1);
            mbar_init(smem_u32(&tail->q_full[s]), 1);
            mbar_init(smem_u32(&tail->p_full[s]), kAoGroup);
            mbar_init(smem_u32(&tail->att_full[s]), 1);
            mbar_init(smem_u32(&tail->a_full[s]), kAoGroup);
            mbar_init(smem_u32(&tail->o_full[s]), 1);
            mbar_init(smem_u32(&tail->r1_free[s]), kAoGroup);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 128; i += kAoThreads) {
        tail->bq[i] = p.bq[i]; tail->bo[i] = p.bo[i]; tail->gamma[i] = p.gamma[i]; tail->beta[i] = p.beta[i];
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

    const int role = warp >> 3;                // 0: group A (softmax, convert), 1: group B (LayerNorm + MMA issue)
    const int e = warp & 7;
    const int q = e & 3;                       // TMEM lane quarter (== warp % 4)
    const int hh = e >> 2;                     // column half [64hh, 64hh+64): heads 2hh, 2hh+1
    const int row = q * 32 + lane;             // tile row == TMEM lane
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    constexpr uint32_t idescQ = umma_idesc_bf16(128, 128), idescA = umma_idesc_bf16(128, 32);

    auto tile_row0 = [&](int t) -> int64_t {   // first global row of the CTA's t-th tile
        const int T = (int)blockIdx.x + t * (int)gridDim.x;
        return (int64_t)(T / p.tps) * p.N + (int64_t)(T % p.tps) * 128;
    };
    auto load_x = [&](int t) {
        const int b = t & 1;
        const int row0 = (int)tile_row0(t);
        const uint32_t xb = smem_u32(&tail->x_full[b]), dst = sbase + kAoOffX + b * kAoXBytes;
        mbar_expect_tx(xb, kAoXBytes);
        tma_load_2d(dst, &tm_x, 0, row0, xb);
        tma_load_2d(dst + 16384, &tm_x, 64, row0, xb);
    };
    auto load_ctx = [&](int t) {
        const int b = t & 1;
        const int T = (int)blockIdx.x + t * (int)gridDim.x;
        const uint32_t cb = smem_u32(&tail->c_full[b]);
        mbar_expect_tx(cb, kAoCBytes);
        tma_load_2d(sbase + kAoOffC + b * kAoCBytes, &tm_ctx, 0, (T / p.tps) * 128, cb);
    };
    auto gemm_q = [&](int b) {
        const uint32_t xs = sbase + kAoOffX + b * kAoXBytes;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
            const uint64_t adesc = make_desc(xs + kb * 16384), bdesc = make_desc(sbase + kAoOffWq + kb * 16384);
#pragma unroll
            for (int k = 0; k < 4; ++k)
                umma_bf16_elect(tmem_base + (uint32_t)(b * 256), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idescQ, (kb | k) != 0);
        }
        umma_commit_elect(smem_u32(&tail->q_full[b]));
    };

    if (role == 0) {
        // =========================== group A: softmax of Q, then att -> bf16 ===========================
        const bool leader = (e == 0) && (lane == 0);
        if (leader) {
            const uint32_t wbar = smem_u32(&tail->w_full);
            mbar_expect_tx(wbar, 2 * kAoWBytes);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + kAoOffWq + kb * 16384, &tm_wq, kb * 64, 0, wbar);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + kAoOffWo + kb * 16384, &tm_wo, kb * 64, 0, wbar);
            load_x(0); load_ctx(0);
            if (n_my > 1) { load_x(1); load_ctx(1); }
        }
        __syncwarp();
        for (int a = 0; a < n_my; a += 2) {
            for (int t = a; t < a + 2 && t < n_my; ++t) {
                // ---- softmax over each head's 32 columns, scaled by 1/sqrt(32); P replaces Q in place
                const int b = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                const uint32_t tb = tmem_base + lane_off + (uint32_t)(b * 256);
                mbar_wait(smem_u32(&tail->q_full[b]), ph);
                tc_fence_after();
                if (leader && t + 2 < n_my) load_x(t + 2);      // GQ(t) is complete: the x slot is free
                __syncwarp();
#pragma unroll 1
                for (int i = 0; i < 2; ++i) {
                    const int col0 = hh * 64 + i * 32;
                    uint32_t raw[32];
                    tmem_ld32_nowait(tb + (uint32_t)col0, raw);
                    tmem_ld_wait();
                    float v[32];
                    const float4* bv = reinterpret_cast<const float4*>(tail->bq + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = bv[j];
                        v[4 * j] = __uint_as_float(raw[4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) + bb.y;
                        v[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) + bb.w;
                    }
                    float m = v[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
                    const float mneg = -m * 1.4426950408889634f;
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float ex;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(v[j], 1.4426950408889634f, mneg)));
                        v[j] = ex;
                        s += ex;
                    }
                    const float inv = 0.17677669529663687f / s;            // 1 / (sqrt(32) * sum)
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j] * inv, v[2 * j + 1] * inv);
                    tmem_st16(tb + (uint32_t)col0, pk);                    // head h = col0/32: packed columns [32h, 32h+16)
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&tail->p_full[b]));
            }
            for (int t = a; t < a + 2 && t < n_my; ++t) {
                // ---- att (fp32, R2) -> bf16 pairs over the thread's own columns: K-step k of GO at R2 + (k<4 ? 8k : 64 + 8(k-4))
                const int b = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                const uint32_t tb = tmem_base + lane_off + (uint32_t)(b * 256 + 128);
                mbar_wait(smem_u32(&tail->att_full[b]), ph);
                tc_fence_after();
                if (leader && t + 2 < n_my) load_ctx(t + 2);    // GATT(t) is complete: the ctx slot is free
                __syncwarp();
#pragma unroll
                for (int s4 = 0; s4 < 4; ++s4) {
                    uint32_t raw[16];
                    tmem_ld16_nowait(tb + (uint32_t)(hh * 64 + s4 * 16), raw);
                    tmem_ld_wait();
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) pk[j] = pack_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]));
                    tmem_st8(tb + (uint32_t)(hh * 64 + s4 * 8), pk);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&tail->a_full[b]));
            }
        }
    } else {
        // =========================== group B: MMA issue + residual + LayerNorm ===========================
        if (e == 4) {                                          // whole warp: warp-uniform MMA issue, one elected lane
            mbar_wait(smem_u32(&tail->w_full), 0);
            mbar_wait(smem_u32(&tail->x_full[0]), 0);
            gemm_q(0);
            if (n_my > 1) { mbar_wait(smem_u32(&tail->x_full[1]), 0); gemm_q(1); }
        }
        __syncwarp();
        for (int a = 0; a < n_my; a += 2) {
            // ---- issue duties of this pair, spread over single lanes of different warps
            if (e == 0) {                                       // GATT(a), GATT(b): one N=32 GEMM per head
                for (int t = a; t < a + 2 && t < n_my; ++t) {
                    const int b = t & 1;
                    const uint32_t ph = (t >> 1) & 1;
                    const uint32_t tacc = tmem_base + (uint32_t)(b * 256);
                    mbar_wait_sleep(smem_u32(&tail->c_full[b]), ph, 32);
                    mbar_wait_sleep(smem_u32(&tail->p_full[b]), ph, 32);
                    tc_fence_after();
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const uint64_t bdesc = make_desc(sbase + kAoOffC + b * kAoCBytes + h * 4096);
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_bf16_ts_elect(tacc + (uint32_t)(128 + 32 * h), tacc + (uint32_t)(32 * h + 8 * ks),
                                               bdesc + (uint64_t)(ks * 2), idescA, ks != 0);
                    }
                    umma_commit_elect(smem_u32(&tail->att_full[b]));
                }
            }
            if (e == 1 || e == 2) {                             // GO(a) by warp 9, GO(b) by warp 10
                const int t = a + (e - 1);
                if (t < n_my) {
                    const int b = t & 1;
                    const uint32_t ph = (t >> 1) & 1;
                    const uint32_t tacc = tmem_base + (uint32_t)(b * 256);
                    mbar_wait_sleep(smem_u32(&tail->a_full[b]), ph, 32);
                    tc_fence_after();
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint64_t bdesc = make_desc(sbase + kAoOffWo + (k >> 2) * 16384) + (uint64_t)((k & 3) * 2);
                        umma_bf16_ts_elect(tacc, tacc + 128 + (uint32_t)(k < 4 ? 8 * k : 64 + 8 * (k - 4)), bdesc, idescQ, k != 0);
                    }
                    umma_commit_elect(smem_u32(&tail->o_full[b]));
                }
            }
            __syncwarp();
            for (int t = a; t < a + 2 && t < n_my; ++t) {
                const int b = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                const uint32_t tb = tmem_base + lane_off + (uint32_t)(b * 256);
                const int T = (int)blockIdx.x + t * (int)gridDim.x;
                const int64_t nrow = (int64_t)(T % p.tps) * 128 + row;             // token index inside the sample
                const bool row_ok = nrow < p.N;
                const int64_t grow = (int64_t)(T / p.tps) * p.N + nrow;
                uint4 res[8];
                {
                    const uint4* src = reinterpret_cast<const uint4*>(p.x + (row_ok ? grow : 0) * 128 + hh * 64);
#pragma unroll
                    for (int j = 0; j < 8; ++j) res[j] = __ldg(src + j);
                }
                mbar_wait_sleep(smem_u32(&tail->o_full[b]), ph);
                tc_fence_after();
                float y[64];
                {
                    uint32_t v0[32], v1[32];
                    tmem_ld32_nowait(tb + (uint32_t)(64 * hh), v0);
                    tmem_ld32_nowait(tb + (uint32_t)(64 * hh + 32), v1);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(smem_u32(&tail->r1_free[b]));
                    if (e == 4 && t + 2 < n_my) {               // GQ of tile t+2 into the drained R1 (whole warp, elected issue)
                        mbar_wait(smem_u32(&tail->r1_free[b]), ph);
                        mbar_wait(smem_u32(&tail->x_full[b]), ph ^ 1);
                        tc_fence_after();
                        gemm_q(b);
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 32; ++j) { y[j] = __uint_as_float(v0[j]); y[32 + j] = __uint_as_float(v1[j]); }
                }
                float s1 = 0.f;
                const float4* b2v = reinterpret_cast<const float4*>(tail->bo + hh * 64);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t w[4] = {res[j].x, res[j].y, res[j].z, res[j].w};
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
                asm volatile("bar.sync 1, 256;" ::: "memory");
                const float2 other = tail->xs[b][hh ^ 1][row];
                const float mean = 0.5f * (m_loc + other.x);
                const float dm = m_loc - other.x;
                const float var = (m2 + other.y + dm * dm * 32.f) * (1.f / 128.f);      // Chan: n_a n_b / (n_a + n_b) = 32
                const float rstd = rsqrtf(var + p.eps);
                if (row_ok) {
                    uint4* dst = reinterpret_cast<uint4*>(p.y + grow * 128 + hh * 64);
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
                        dst[j] = ov;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Second form (ltu_attn_out_fused_w): the readout folded into the output projection.  (P ctx_b) Wo^T = P (ctx_b Wo^T), and
// W_b = blockdiag(ctx_b) Wo^T is one [128 x 128] bf16 matrix per SAMPLE (written by the kv_combine tail), so a tile needs TWO
// chained GEMMs instead of three and the attention output never exists, not even in tensor memory:
//
//       GQ   R1[128x128] = X . Wq^T            A, B from smem
//       GO   R2[128x128] = P . W_b^T           A = softmax(Q) as bf16 pairs in R1 (tensor memory), B = the sample's W_b (smem)
//
//   warps 0-7    softmax of Q per head (in place, R1)                        warp 16   MMA issue (GQ one tile ahead of GO)
//   warps 8-15   + bo + residual -> LayerNorm1 -> y (reads R2)               warp 17   TMA: Wq once; x and W_b per tile
// GQ(t+2) may overwrite R1 as soon as GO(t) has completed; R2 is free again when the LayerNorm warps have read it.
// An x tile serves three times in its shared-memory slot: A operand of GQ, residual of the LayerNorm (each thread reads its own
// row half) and, overwritten in place by the same thread, staging tile of the TMA store of y -- no per-thread global access.
constexpr int kAmThreads = 576;
constexpr uint32_t kAmOffWq = 0;
constexpr uint32_t kAmOffX = kAoWBytes;
constexpr int kAmXSlots = 3;                    // an x tile is the A operand of GQ and, later, the residual of the LayerNorm
constexpr uint32_t kAmOffM = kAmOffX + kAmXSlots * kAoXBytes;
constexpr uint32_t kAmOffTail = kAmOffM + 2 * kAoWBytes;

struct AmTail {
    uint64_t w_full, x_full[kAmXSlots], x_free[kAmXSlots], m_full[2], q_full[2], p_full[2], o_full[2], r2_free[2];
    uint32_t tmem_slot, pad_;
    alignas(16) float bq[128], bo[128], gamma[128], beta[128];
    float2 xs[2][2][128];
};

__global__ void __launch_bounds__(kAmThreads, 1)
attn_out128w_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wq,
                    const __grid_constant__ CUtensorMap tm_m, const __grid_constant__ CUtensorMap tm_y, const AoParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    AmTail* tail = reinterpret_cast<AmTail*>(smem + kAmOffTail);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_my = (p.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto stamp = [&](int role, int t, int ev) {
        if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && t < 64) p.trace[(role * 64 + t) * 8 + ev] = clock64();
    };

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&tail->w_full), 1);
        for (int s = 0; s < kAmXSlots; ++s) { mbar_init(smem_u32(&tail->x_full[s]), 1); mbar_init(smem_u32(&tail->x_free[s]), 8); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&tail->m_full[s]), 1);
            mbar_init(smem_u32(&tail->q_full[s]), 1);
            mbar_init(smem_u32(&tail->p_full[s]), kAoGroup);
            mbar_init(smem_u32(&tail->o_full[s]), 1);
            mbar_init(smem_u32(&tail->r2_free[s]), kAoGroup);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 128; i += kAmThreads) {        // parameters: no dependency on the previous kernel
        tail->bq[i] = p.bq[i]; tail->bo[i] = p.bo[i]; tail->gamma[i] = p.gamma[i]; tail->beta[i] = p.beta[i];
    }
    if (warp == 16) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(&tail->tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_slot;
    constexpr uint32_t idescQ = umma_idesc_bf16(128, 128);

    if (warp == 17) {
        // =========================== TMA producer ===========================
        if (lane == 0) {
            const uint32_t wbar = smem_u32(&tail->w_full);
            mbar_expect_tx(wbar, kAoWBytes);
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(sbase + kAmOffWq + kb * 16384, &tm_wq, kb * 64, 0, wbar);
            pdl_prologue();                                  // x and W_b come from the previous kernels in the stream
            for (int t = 0; t < n_my; ++t) {
                const int b = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                const int T = (int)blockIdx.x + t * (int)gridDim.x;
                const int smp = T / p.tps;
                const int row0 = (int)((int64_t)smp * p.N + (int64_t)(T % p.tps) * 128);
                const int xs = t % kAmXSlots;
                if (t >= kAmXSlots) mbar_wait(smem_u32(&tail->x_free[xs]), ((t / kAmXSlots) & 1) ^ 1);   // the stores of y(t-3) have read the slot
                stamp(0, t, 0);                                                 // x(t) load issued
                const uint32_t xb = smem_u32(&tail->x_full[xs]), xd = sbase + kAmOffX + xs * kAoXBytes;
                mbar_expect_tx(xb, kAoXBytes);
                tma_load_2d(xd, &tm_x, 0, row0, xb);
                tma_load_2d(xd + 16384, &tm_x, 64, row0, xb);
                if (t >= 2) mbar_wait(smem_u32(&tail->o_full[b]), ph ^ 1);      // GO(t-2) has read the W_b slot
                stamp(0, t, 1);                                                 // W_b(t) load issued
                const uint32_t mb = smem_u32(&tail->m_full[b]), md = sbase + kAmOffM + b * kAoWBytes;
                mbar_expect_tx(mb, kAoWBytes);
                tma_load_2d(md, &tm_m, 0, smp * 128, mb);
                tma_load_2d(md + 16384, &tm_m, 64, smp * 128, mb);
            }
        }
    } else if (warp == 16) {
        // =========================== MMA issue (whole warp, one elected lane) ===========================
        mbar_wait(smem_u32(&tail->w_full), 0);
        for (int t = 0; t <= n_my; ++t) {
            if (t < n_my) {                                                     // GQ(t)
                const int b = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                if (t >= 2) mbar_wait(smem_u32(&tail->o_full[b]), ph ^ 1);      // GO(t-2) has read P out of R1
                mbar_wait(smem_u32(&tail->x_full[t % kAmXSlots]), (t / kAmXSlots) & 1);
                stamp(1, t, 0);                                                 // GQ(t) issued
                tc_fence_after();
                const uint32_t xs = sbase + kAmOffX + (t % kAmXSlots) * kAoXBytes;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
                    const uint64_t adesc = make_desc(xs + kb * 16384), bdesc = make_desc(sbase + kAmOffWq + kb * 16384);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16_elect(tmem_base + (uint32_t)(b * 256), adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idescQ, (kb | k) != 0);
                }
                umma_commit_elect(smem_u32(&tail->q_full[b]));
            }
            if (t >= 1) {                                                       // GO(t-1)
                const int u = t - 1, b = u & 1;
                const uint32_t ph = (u >> 1) & 1;
                const uint32_t tacc = tmem_base + (uint32_t)(b * 256);
                mbar_wait(smem_u32(&tail->m_full[b]), ph);
                if (u >= 2) mbar_wait(smem_u32(&tail->r2_free[b]), ph ^ 1);     // LayerNorm(u-2) has read R2
                mbar_wait(smem_u32(&tail->p_full[b]), ph);
                stamp(1, u, 1);                                                 // GO(u) issued
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 8; ++k) {                                   // K-step k = head k/2, j in [16 (k%2), +16)
                    const uint64_t bdesc = make_desc(sbase + kAmOffM + b * kAoWBytes + (k >> 2) * 16384) + (uint64_t)((k & 3) * 2);
                    umma_bf16_ts_elect(tacc + 128u, tacc + (uint32_t)(32 * (k >> 1) + 8 * (k & 1)), bdesc, idescQ, k != 0);
                }
                umma_commit_elect(smem_u32(&tail->o_full[b]));
            }
        }
    } else {
        pdl_prologue();
        const int role = warp >> 3;                // 0: softmax warps, 1: LayerNorm warps
        const int e = warp & 7;
        const int q = e & 3;                       // TMEM lane quarter (== warp % 4)
        const int hh = e >> 2;                     // column half [64hh, 64hh+64): heads 2hh, 2hh+1
        const int row = q * 32 + lane;             // tile row == TMEM lane
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        if (role == 0) {
            // =========================== softmax over each head's 32 columns, scaled by 1/sqrt(32); P replaces Q in place
            for (int t = 0; t < n_my; ++t) {
                const int b = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                const uint32_t tb = tmem_base + lane_off + (uint32_t)(b * 256);
                if (e == 0) stamp(2, t, 0);                                     // softmax warp 0: waiting for Q(t)
                mbar_wait(smem_u32(&tail->q_full[b]), ph);
                if (e == 0) stamp(2, t, 1);                                     // Q(t) complete
                tc_fence_after();
#pragma unroll 1
                for (int i = 0; i < 2; ++i) {
                    const int col0 = hh * 64 + i * 32;
                    uint32_t raw[32];
                    tmem_ld32_nowait(tb + (uint32_t)col0, raw);
                    tmem_ld_wait();
                    float v[32];
                    const float4* bv = reinterpret_cast<const float4*>(tail->bq + col0);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 bb = bv[j];
                        v[4 * j] = __uint_as_float(raw[4 * j]) + bb.x; v[4 * j + 1] = __uint_as_float(raw[4 * j + 1]) + bb.y;
                        v[4 * j + 2] = __uint_as_float(raw[4 * j + 2]) + bb.z; v[4 * j + 3] = __uint_as_float(raw[4 * j + 3]) + bb.w;
                    }
                    float m = v[0];
#pragma unroll
                    for (int j = 1; j < 32; ++j) m = fmaxf(m, v[j]);
                    const float mneg = -m * 1.4426950408889634f;
                    float s = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        float ex;
                        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex) : "f"(fmaf(v[j], 1.4426950408889634f, mneg)));
                        v[j] = ex;
                        s += ex;
                    }
                    const float inv = 0.17677669529663687f / s;            // 1 / (sqrt(32) * sum)
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(v[2 * j] * inv, v[2 * j + 1] * inv);
                    tmem_st16(tb + (uint32_t)col0, pk);                    // head h = col0/32: packed columns [32h, 32h+16)
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(smem_u32(&tail->p_full[b]));
                if (e == 0) stamp(2, t, 2);                                     // P(t) published
            }
        } else {
            // =========================== + bo + residual -> LayerNorm1 -> y ===========================
            for (int t = 0; t < n_my; ++t) {
                const int b = t & 1;
                const uint32_t ph = (t >> 1) & 1;
                const uint32_t tb = tmem_base + lane_off + (uint32_t)(b * 256 + 128);
                const int T = (int)blockIdx.x + t * (int)gridDim.x;
                if (e == 0) stamp(3, t, 0);                                     // LayerNorm warp 8: waiting for GO(t)
                mbar_wait_sleep(smem_u32(&tail->o_full[b]), ph, 32);
                if (e == 0) stamp(3, t, 1);                                     // GO(t) complete
                mbar_wait(smem_u32(&tail->x_full[t % kAmXSlots]), (t / kAmXSlots) & 1);     // completed long ago: makes the TMA write visible here
                tc_fence_after();
                float y[64];
                {
                    uint32_t v0[32], v1[32];
                    tmem_ld32_nowait(tb + (uint32_t)(64 * hh), v0);
                    tmem_ld32_nowait(tb + (uint32_t)(64 * hh + 32), v1);
                    tmem_ld_wait();
                    tc_fence_before();
                    mbar_arrive(smem_u32(&tail->r2_free[b]));
#pragma unroll
                    for (int j = 0; j < 32; ++j) { y[j] = __uint_as_float(v0[j]); y[32 + j] = __uint_as_float(v1[j]); }
                }
                float s1 = 0.f;
                const float4* b2v = reinterpret_cast<const float4*>(tail->bo + hh * 64);
                // the residual is the row's own x, still in the tile's shared-memory slot (k-block hh, SWIZZLE_128B row)
                unsigned char* xrow = smem + kAmOffX + (t % kAmXSlots) * kAoXBytes + hh * 16384 + row * 128;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 rv = *reinterpret_cast<const uint4*>(xrow + ((j ^ (row & 7)) << 4));
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
                if (e == 0) stamp(3, t, 2);                                     // accumulator read, residual added, partial statistics
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (e == 0) stamp(3, t, 3);                                     // partner's statistics visible
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
                        *reinterpret_cast<uint4*>(xrow + ((j ^ (row & 7)) << 4)) = ov;       // over the residual it was read from
                    }
                }
                // the warp's [32 rows x 64 columns] of y sit in the x slot in the TMA (SWIZZLE_128B) layout: one bulk store per
                // warp, rows past the end of the sample are clipped by the 3-D map; the slot is free once the store has read it
                if (e == 0) stamp(3, t, 4);                                     // normalised row written to the slot
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tm_y, sbase + kAmOffX + (t % kAmXSlots) * kAoXBytes + hh * 16384 + q * 4096, hh * 64,
                                 (T % p.tps) * 128 + q * 32, T / p.tps);
                    tma_store_commit();
                    tma_store_wait_read();
                    mbar_arrive(smem_u32(&tail->x_free[t % kAmXSlots]));
                }
                if (e == 0) stamp(3, t, 5);                                     // store has read the slot
            }
            if (lane == 0) tma_store_wait_all();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 16) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ctx fp32 [B][4][32 j][32 e] -> the bf16 B operand of GATT: out[(b*128 + h*32 + e)*64 + j] = ctx[b][h][j][e], j < 32; 0 for j >= 32
__global__ void ctx_pack_kernel(const float* __restrict__ ctx, bf16* __restrict__ out, int total_rows) {
    const int r = blockIdx.x * (blockDim.x >> 6) + (threadIdx.x >> 6);       // row b*128 + h*32 + e
    const int j = threadIdx.x & 63;
    if (r >= total_rows) return;
    const int bh = r >> 5, ee = r & 31;
    out[(int64_t)r * 64 + j] = j < 32 ? __float2bfloat16_rn(ctx[((int64_t)bh * 32 + j) * 32 + ee]) : __float2bfloat16_rn(0.f);
}

}  // namespace ltu

using namespace ltu;

extern "C" int ltu_attn_out_fused_supported(int C, int heads) { return (C == 128 && heads == 4) ? 1 : 0; }

extern "C" int ltu_ctx_pack_bf16(const float* ctx, void* out, int B, int heads, ltu_stream_t stream) {
    LTU_ARG_CHECK(ctx && out && B > 0 && heads > 0, "ctx_pack_bf16: bad arguments");
    const int rows = B * heads * 32;
    ctx_pack_kernel<<<(rows + 3) / 4, 256, 0, (cudaStream_t)stream>>>(ctx, (bf16*)out, rows);
    LTU_LAUNCH_CHECK("ctx_pack_bf16");
    count_launch(1);
    return LTU_OK;
}

extern "C" int ltu_attn_out_fused(const void* x, int B, int64_t N, int C, int heads, const void* wq_bf16, const float* bq,
                                  const void* ctx_bf16, const void* wo_bf16, const float* bo, const float* gamma,
                                  const float* beta, float eps, void* y, ltu_stream_t stream) {
    LTU_ARG_CHECK(C == 128 && heads == 4, "attn_out_fused: d_model %d / heads %d not supported (128 / 4)", C, heads);
    LTU_ARG_CHECK(x && y && wq_bf16 && wo_bf16 && ctx_bf16 && bq && bo && gamma && beta, "attn_out_fused: null pointer");
    LTU_ARG_CHECK(B > 0 && N > 0 && (int64_t)B * N < ((int64_t)1 << 31) - 256, "attn_out_fused: bad shape");
    LTU_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)wq_bf16 & 15) == 0 &&
                  ((uintptr_t)wo_bf16 & 15) == 0 && ((uintptr_t)ctx_bf16 & 15) == 0, "attn_out_fused: pointers must be 16-byte aligned");
    CUtensorMap tx, twq, two, tc;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tx, x, (uint64_t)B * (uint64_t)N, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&twq, wq_bf16, 128, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&two, wo_bf16, 128, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tc, ctx_bf16, (uint64_t)B * 128, 64, 128)) != LTU_OK) return rc;
    AoParams p;
    p.x = (const bf16*)x; p.y = (bf16*)y; p.bq = bq; p.bo = bo; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.N = N; p.tps = (int)((N + 127) / 128); p.tiles = p.tps * B; p.trace = nullptr;
    const size_t smem = 1024 + kAoOffTail + sizeof(AoTail);
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(attn_out128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured_dev = dev;
    }
    int grid = sm_count();
    if (grid > p.tiles) grid = p.tiles;
    attn_out128_kernel<<<grid, kAoThreads, smem, (cudaStream_t)stream>>>(tx, twq, two, tc, p);
    LTU_LAUNCH_CHECK("attn_out_fused");
    count_launch(1);
    return LTU_OK;
}

// y = LayerNorm(x + softmax_d(x Wq^T + bq)/sqrt(32) . W_b^T + bo): ltu_attn_out_fused with the readout folded into the output
// projection; w_b bf16 [B][128][128] = blockdiag(ctx_b) Wo^T from ltu_kv_project_reduce / ltu_kv_reduce_project / ltu_ctx_project
extern "C" int ltu_attn_out_fused_w(const void* x, int B, int64_t N, int C, int heads, const void* wq_bf16, const float* bq,
                                    const void* w_b, const float* bo, const float* gamma, const float* beta, float eps, void* y,
                                    ltu_stream_t stream) {
    LTU_ARG_CHECK(C == 128 && heads == 4, "attn_out_fused_w: d_model %d / heads %d not supported (128 / 4)", C, heads);
    LTU_ARG_CHECK(x && y && wq_bf16 && w_b && bq && bo && gamma && beta, "attn_out_fused_w: null pointer");
    LTU_ARG_CHECK(B > 0 && N > 0 && (int64_t)B * N < ((int64_t)1 << 31) - 256, "attn_out_fused_w: bad shape");
    LTU_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0 && ((uintptr_t)wq_bf16 & 15) == 0 && ((uintptr_t)w_b & 15) == 0,
                  "attn_out_fused_w: pointers must be 16-byte aligned");
    CUtensorMap tx, twq, tm;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tx, x, (uint64_t)B * (uint64_t)N, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&twq, wq_bf16, 128, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tm, w_b, (uint64_t)B * 128, 128, 128)) != LTU_OK) return rc;
    CUtensorMap ty;
    if ((rc = make_tmap_bf16_3d(&ty, y, (uint64_t)B, (uint64_t)N, 128, 32)) != LTU_OK) return rc;
    AoParams p;
    p.x = (const bf16*)x; p.y = (bf16*)y; p.bq = bq; p.bo = bo; p.gamma = gamma; p.beta = beta; p.eps = eps;
    p.N = N; p.tps = (int)((N + 127) / 128); p.tiles = p.tps * B;
    { const char* e = getenv("LTU_AO_TRACE_PTR"); p.trace = e ? (long long*)strtoull(e, nullptr, 0) : nullptr; }   // tools/ao_trace.py
    const size_t smem = 1024 + kAmOffTail + sizeof(AmTail);
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(attn_out128w_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured_dev = dev;
    }
    int grid = sm_count();
    if (grid > p.tiles) grid = p.tiles;
    cudaError_t e = launch_pdl(attn_out128w_kernel, dim3(grid), dim3(kAmThreads), smem, (cudaStream_t)stream, tx, twq, tm, ty, p);
    if (e != cudaSuccess) { set_error("attn_out_fused_w: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    count_launch(1);
    return LTU_OK;
}
