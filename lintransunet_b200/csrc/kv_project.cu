// Key / value half of the linear-attention core for d_model 128 (bridge 1: 57 408 tokens per sample) as ONE kernel:
//
//     ctx[b][h] = softmax_N(x_b Wk^T + bk)_h^T (x_b Wv^T + bv)_h          (model/trans_block.py:155-156 and :59-60)
//
// K and V are never written to memory.  The separate path writes them (2 x rows x 128 bf16 from ltu_linear_fused) and reads
// them back (ltu_kv_reduce): 4 x rows x C x 2 bytes per layer, 470 MB at the benchmark's batch -- more than everything else
// the layer moves.  Here a persistent TMA + tcgen05 CTA computes the [128 rows x 256] tile  [K | V] = x W_kv^T  into tensor
// memory (W_kv = 64 KB, resident in shared memory; x streams through a 6-stage TMA ring), and the epilogue warps, which
// already turn accumulator columns into bf16 rows in a SWIZZLE_128B staging tile, run the kv_reduce warp routine of
// attn_stream.cu on those staging tiles instead of storing them: K -> P = 2^(k log2e - r_j) in place, ctx += P^T V with
// ldmatrix.trans + mma.sync m16n8k16, column sums from an all-ones B tile, per-warp reference with an exact rescale path.
//
//   warp 0      TMA producer (W once, then the x k-blocks of the CTA's tiles)
//   warp 1      tcgen05.mma M128 N256 K16 into one of two TMEM accumulators
//   warps 2-17  four warps per TMEM lane quarter q: warp (q, cg) reads head cg's 32 K and 32 V accumulator columns of its 32
//               lanes, stages them as bf16 in a PRIVATE SWIZZLE_128B chunk and reduces that head over the quarter's 32 rows
//               -- no barrier between epilogue warps (16 reducing warps per SM, as many as two resident kv_stream CTAs)
// A CTA owns a CONTIGUOUS range of row tiles, so it touches at most two or three samples; a warp keeps its 32 x 32
// state in registers and flushes a partial (ctx, reference, column sums) when its rows move to the next sample.  The
// partials (<= 4 per CTA and head, 80 per sample at the benchmark's batch) are merged by the same fixed-order kv_combine
// kernel as ltu_kv_reduce's; slots a warp never fills are written as neutral states, so the result does not depend on timing.
#include <cuda.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);   // ffn_tc.cu
int kv_combine_launch(const float* ws, float* ctx, int heads, int B, int nparts, cudaStream_t st, const void* wo = nullptr,
                      void* wout = nullptr);            // attn_kernels.cu

// second form (kv_project2.cu): G = P^T x on the tensor pipe, V never computed per token
bool kv_project2_enabled();
size_t kv_project2_workspace(int B, int64_t N);
int kv_project2_launch(const void* x, const void* w_kv, const float* bias, float* ctx, void* workspace, int B, int64_t N,
                       const void* wo_bf16, void* w_out, cudaStream_t stream);

namespace {

constexpr int kKvpThreads = 576;                              // 18 warps: producer, MMA issue, 16 epilogue / reduce warps
constexpr int kKvpStages = 6;                                 // x ring: three tiles of look-ahead
constexpr int kKvpPart = 32 * 32 + 64;                        // ctx[32][32], m[32], s[32] (the layout kv_combine reads)
constexpr int kKvpHeads = 4;
constexpr uint32_t kKvpWBytes = 256 * 128 * 2;                // W_kv: two k-blocks of [256 x 64] bf16
constexpr uint32_t kKvpABytes = 128 * 128;                    // x k-block: 128 rows x 64 bf16
constexpr uint32_t kKvpOffRing = kKvpWBytes;
constexpr uint32_t kKvpOffStaging = kKvpOffRing + kKvpStages * kKvpABytes;
constexpr uint32_t kKvpOffTail = kKvpOffStaging + 16 * 4096;  // one PRIVATE [32 rows x (32 K + 32 V) columns] bf16 chunk per epilogue warp
constexpr float kL2e = 1.4426950408889634f;

struct KvpTail {
    uint64_t w_full, full[kKvpStages], empty[kKvpStages], tfull[2], tempty[2];
    uint32_t tmem_slot, pad_;
    alignas(16) float bias[256];
};

struct KvpParams {
    const float* bias;
    float* part;                    // [B][nparts][4 heads][kKvpPart]
    int64_t rows, N;                // rows = B * N tokens; N % 32 == 0
    int B, tiles_m, nparts;
};

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void lds16(uint32_t addr, uint4& v) {
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
}
__device__ __forceinline__ void sts16(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void unpack(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
__device__ __forceinline__ uint4 pack(const float (&k)[8]) {
    uint4 o;
    o.x = pack_bf16x2(k[0], k[1]); o.y = pack_bf16x2(k[2], k[3]);
    o.z = pack_bf16x2(k[4], k[5]); o.w = pack_bf16x2(k[6], k[7]);
    return o;
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [32 rows][128 B] SWIZZLE_128B tile (1024-B aligned)
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// running state of one head in one warp: ctx fragments, column-sum fragments, reference (log2 units)
struct HeadState {
    float acc[2][4][4];
    float accs[2][4];
    float rj[8];
    bool have_ref;
};

__device__ __forceinline__ void reset(HeadState& s) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) s.acc[mt][nt][i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) s.accs[mt][i] = 0.f;
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) s.rj[c] = 0.f;
    s.have_ref = false;
}

// One head, the warp's 32 rows: `tile` = shared address of the warp's private [32 rows x 128 B] staging chunk, K of the head in
// columns 0-31, V in columns 32-63.  The K half is turned into P in place.  (attn_stream.cu, kv_stream_kernel, with sub = 0;
// the keys are read twice from shared memory -- check, then exponentials -- to keep the register footprint small.)
__device__ __forceinline__ void reduce_head(HeadState& s, uint32_t tile, int valid, int lane) {
    const uint32_t tK = tile, tV = tile;
    constexpr int half = 0, halfV = 1;                           // K in 16-byte chunks 0-3 of a row, V in chunks 4-7
    const int g = lane >> 2, cc = lane & 3, mi = lane >> 3, lr = lane & 7;
    const uint32_t ones = g == 0 ? 0x3F803F80u : 0u;
    uint32_t addr[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) addr[i] = tK + swz(g + 8 * i, half * 4 + cc);
    // pass A: column max of the lane's rows (log2 units)
    float mx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) mx[c] = -INFINITY;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (g + 8 * i < valid) {
            uint4 v;
            float k[8];
            lds16(addr[i], v);
            unpack(v, k);
#pragma unroll
            for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], k[c] * kL2e);
        }
    bool need_max = !s.have_ref;
    if (s.have_ref) {
        float dmax = -INFINITY;
#pragma unroll
        for (int c = 0; c < 8; ++c) dmax = fmaxf(dmax, mx[c] - s.rj[c]);
        need_max = __any_sync(0xffffffffu, dmax > 64.f);       // rare: the data ran away from the reference
    }
    if (need_max) {
        float f[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
            const float m = s.have_ref ? fmaxf(s.rj[c], mx[c]) : mx[c];
            f[c] = s.have_ref ? ex2(s.rj[c] - m) : 1.f;
            s.rj[c] = m;
        }
        if (s.have_ref) {                                        // rescale the running state by 2^(r_old - r_new)
            // after the butterfly every lane with the same cc holds the same f[]: row j = mt*16 + g (+8) wants column j's
            // factor = register g of a lane with cc = 2 mt (+1) -- eight shuffles instead of a shared-memory scratch
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                float f0 = 1.f, f1 = 1.f;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float a0 = __shfl_sync(0xffffffffu, f[c], mt * 2), a1 = __shfl_sync(0xffffffffu, f[c], mt * 2 + 1);
                    if (g == c) { f0 = a0; f1 = a1; }
                }
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    s.acc[mt][nt][0] *= f0; s.acc[mt][nt][1] *= f0; s.acc[mt][nt][2] *= f1; s.acc[mt][nt][3] *= f1;
                }
                s.accs[mt][0] *= f0; s.accs[mt][1] *= f0; s.accs[mt][2] *= f1; s.accs[mt][3] *= f1;
            }
        }
        s.have_ref = true;
    }
    // pass B: P = 2^(k log2e - r_j) in place, rows past the end are zero
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool ok = g + 8 * i < valid;
        uint4 v;
        float k[8];
        lds16(addr[i], v);
        unpack(v, k);
#pragma unroll
        for (int c = 0; c < 8; ++c) k[c] = ok ? ex2(fmaf(k[c], kL2e, -s.rj[c])) : 0.f;
        sts16(addr[i], pack(k));
    }
    __syncwarp();                                                // P visible to the whole warp
    // ctx[j][e] += sum_n P[n][j] V[n][e]   (M = j, N = e, K = the 32 rows)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        const int n0 = ks * 16;
        uint32_t a[2][4], bq[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
            ldsm4t(tK + swz(n0 + lr + 8 * (mi >> 1), half * 4 + mt * 2 + (mi & 1)), a[mt]);
#pragma unroll
        for (int np = 0; np < 2; ++np)
            ldsm4t(tV + swz(n0 + lr + 8 * (mi & 1), halfV * 4 + np * 2 + (mi >> 1)), bq[np]);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
                mma(s.acc[mt][nt], a[mt], bq[nt >> 1][(nt & 1) * 2], bq[nt >> 1][(nt & 1) * 2 + 1]);
            mma(s.accs[mt], a[mt], ones, ones);
        }
    }
}

// partial state of one head -> out[kKvpPart]: ctx[j][e], m[j] (natural-log units), s[j]
__device__ __forceinline__ void flush_head(const HeadState& s, float* out, int lane) {
    const int g = lane >> 2, cc = lane & 3, tq = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int j = mt * 16 + g, e = nt * 8 + 2 * tq;
            *reinterpret_cast<float2*>(out + j * 32 + e) = make_float2(s.acc[mt][nt][0], s.acc[mt][nt][1]);
            *reinterpret_cast<float2*>(out + (j + 8) * 32 + e) = make_float2(s.acc[mt][nt][2], s.acc[mt][nt][3]);
        }
    if (g == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) out[1024 + cc * 8 + c] = s.have_ref ? s.rj[c] * (1.f / kL2e) : -INFINITY;
    }
    if (tq == 0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            out[1056 + mt * 16 + g] = s.accs[mt][0];
            out[1056 + mt * 16 + g + 8] = s.accs[mt][2];
        }
    }
}

__device__ __forceinline__ void neutral_head(float* out, int lane) {
    for (int i = lane; i < 1024; i += 32) out[i] = 0.f;
    out[1024 + lane] = -INFINITY;
    out[1056 + lane] = 0.f;
}

// contiguous tile ranges: CTA c owns tiles [c * M / G, (c + 1) * M / G)
__host__ __device__ __forceinline__ int64_t range_begin(int64_t c, int64_t M, int64_t G) { return c * M / G; }
__host__ __device__ __forceinline__ int64_t tile_owner(int64_t t, int64_t M, int64_t G) { return ((t + 1) * G - 1) / M; }

// 576 threads are allocated as 20 warps of registers: 65536 / (20 * 32) = 102 registers per thread at most
__global__ void __launch_bounds__(kKvpThreads, 1)
kv_project_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w, const KvpParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    KvpTail* tail = reinterpret_cast<KvpTail*>(smem + kKvpOffTail);
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t G = gridDim.x, M = p.tiles_m;
    const int t0 = (int)range_begin(blockIdx.x, M, G), t1 = (int)range_begin(blockIdx.x + 1, M, G);

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&tail->w_full), 1);
        for (int s = 0; s < kKvpStages; ++s) { mbar_init(smem_u32(&tail->full[s]), 1); mbar_init(smem_u32(&tail->empty[s]), 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&tail->tfull[i]), 1); mbar_init(smem_u32(&tail->tempty[i]), 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < 256; i += kKvpThreads) tail->bias[i] = p.bias[i];
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
        // =========================== producer ===========================
        if (lane == 0) {
            const uint32_t wb = smem_u32(&tail->w_full);
            mbar_expect_tx(wb, kKvpWBytes);
            tma_load_2d(sbase, &tm_w, 0, 0, wb);                             // W_kv is a parameter: no dependency to wait for
            tma_load_2d(sbase + kKvpWBytes / 2, &tm_w, 64, 0, wb);
            pdl_prologue();                                                  // x comes from the previous kernel in the stream
            uint32_t kbc = 0;
            for (int T = t0; T < t1; ++T)
                for (int kb = 0; kb < 2; ++kb, ++kbc) {
                    const int stage = kbc % kKvpStages;
                    mbar_wait(smem_u32(&tail->empty[stage]), ((kbc / kKvpStages) & 1) ^ 1);
                    const uint32_t fb = smem_u32(&tail->full[stage]);
                    mbar_expect_tx(fb, kKvpABytes);
                    tma_load_2d(sbase + kKvpOffRing + stage * kKvpABytes, &tm_x, kb * 64, T * 128, fb);
                }
        }
    } else if (warp == 1) {
        // =========================== MMA issue ===========================
        constexpr uint32_t idesc = umma_idesc_bf16(128, 256);
        mbar_wait(smem_u32(&tail->w_full), 0);
        uint32_t kbc = 0, it = 0;
        for (int T = t0; T < t1; ++T, ++it) {
            const uint32_t abuf = it & 1;
            mbar_wait(smem_u32(&tail->tempty[abuf]), ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tacc = tmem_base + abuf * 256u;
            for (int kb = 0; kb < 2; ++kb, ++kbc) {
                const int stage = kbc % kKvpStages;
                mbar_wait(smem_u32(&tail->full[stage]), (kbc / kKvpStages) & 1);
                tc_fence_after();
                const uint64_t adesc = make_desc(sbase + kKvpOffRing + stage * kKvpABytes);
                const uint64_t bdesc = make_desc(sbase + kb * (kKvpWBytes / 2));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_bf16_elect(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                umma_commit_elect(smem_u32(&tail->empty[stage]));
                if (kb == 1) umma_commit_elect(smem_u32(&tail->tfull[abuf]));
            }
        }
    } else {
        // =========================== epilogue + kv_reduce ===========================
        const int e = warp - 2;                          // 0..15
        const int q = warp & 3;                          // TMEM lane quarter (warp id % 4) = rows [32q, 32q + 32) of the tile
        const int cg = e >> 2;                           // stages accumulator columns [64 cg, +64); reduces head cg
        // Warp (q, cg) owns head cg of rows [32q, 32q + 32): it reads the head's 32 K columns and 32 V columns of its TMEM lanes
        // itself and stages them in a PRIVATE chunk, so the epilogue has no cross-warp dependency at all: no block or named
        // barrier, no double buffering (first version: column groups staged by one warp and reduced by another -- two 128-thread
        // barriers per tile, 116 us; double-buffered staging with one barrier and a 2-stage x ring, 107 us).
        unsigned char* stg = smem + kKvpOffStaging + e * 4096 + lane * 128;
        const uint32_t tile_u = sbase + kKvpOffStaging + (uint32_t)(e * 4096);
        const uint32_t swzl = (uint32_t)(lane & 7);
        HeadState hs;
        reset(hs);
        int64_t cur_b = -1;
        const int64_t row_first = (int64_t)t0 * 128, row_last = ((int64_t)t1 * 128 < p.rows ? (int64_t)t1 * 128 : p.rows) - 1;
        const int64_t b_first = t1 > t0 ? row_first / p.N : 0, b_last = t1 > t0 ? row_last / p.N : -1;
        uint32_t done_mask = 0;                          // samples (relative to b_first) this warp has written a partial for
        auto slot_ptr = [&](int64_t b) {
            const int64_t c_lo = tile_owner(b * p.N / 128, M, G);
            const int64_t slot = ((int64_t)blockIdx.x - c_lo) * 4 + q;
            return p.part + (((int64_t)b * p.nparts + slot) * kKvpHeads + cg) * kKvpPart;
        };
        uint32_t it = 0;
        for (int T = t0; T < t1; ++T, ++it) {
            const uint32_t abuf = it & 1;
            const uint32_t tb = tmem_base + abuf * 256u + ((uint32_t)(q * 32) << 16);
            mbar_wait_sleep(smem_u32(&tail->tfull[abuf]), (it >> 1) & 1, 64);
            tc_fence_after();
#pragma unroll 1
            for (int c4 = 0; c4 < 4; ++c4) {              // 16 columns at a time: K [32 cg, +32) then V [128 + 32 cg, +32)
                const int kv = c4 >> 1, off = kv * 128 + cg * 32 + (c4 & 1) * 16;
                uint32_t v[16];
                tmem_ld16_nowait(tb + (uint32_t)off, v);
                tmem_ld_wait();
                if (c4 == 3) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&tail->tempty[abuf]));
                }
                const float* bs = tail->bias + off;
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) {
                    const uint32_t* src = v + 8 * jj;
                    const float4 ba = *reinterpret_cast<const float4*>(bs + jj * 8);
                    const float4 bb = *reinterpret_cast<const float4*>(bs + jj * 8 + 4);
                    uint4 ov;
                    ov.x = pack_bf16x2(__uint_as_float(src[0]) + ba.x, __uint_as_float(src[1]) + ba.y);
                    ov.y = pack_bf16x2(__uint_as_float(src[2]) + ba.z, __uint_as_float(src[3]) + ba.w);
                    ov.z = pack_bf16x2(__uint_as_float(src[4]) + bb.x, __uint_as_float(src[5]) + bb.y);
                    ov.w = pack_bf16x2(__uint_as_float(src[6]) + bb.z, __uint_as_float(src[7]) + bb.w);
                    const uint32_t j = (uint32_t)(c4 * 2 + jj);
                    *reinterpret_cast<uint4*>(stg + ((j ^ swzl) << 4)) = ov;
                }
            }
            __syncwarp();                                                     // the chunk is visible to the whole warp
            const int64_t row0w = (int64_t)T * 128 + q * 32;
            int valid = (int)(p.rows - row0w < 32 ? p.rows - row0w : 32);
            if (valid > 0) {
                const int64_t b = row0w / p.N;                                // N % 32 == 0: a warp's rows share a sample
                if (b != cur_b) {
                    if (cur_b >= 0) { flush_head(hs, slot_ptr(cur_b), lane); done_mask |= 1u << (int)(cur_b - b_first); }
                    reset(hs);
                    cur_b = b;
                }
                reduce_head(hs, tile_u, valid, lane);
            }
            __syncwarp();                                                     // every lane is done with the chunk
        }
        if (cur_b >= 0) { flush_head(hs, slot_ptr(cur_b), lane); done_mask |= 1u << (int)(cur_b - b_first); }
        // slots nobody fills: samples of this CTA's range this warp never touched, and the padding of short ranges
        for (int64_t b = b_first; b <= b_last; ++b) {
            if (!((done_mask >> (int)(b - b_first)) & 1u)) neutral_head(slot_ptr(b), lane);
            const int64_t c_lo = tile_owner(b * p.N / 128, M, G), c_hi = tile_owner(((b + 1) * p.N - 1) / 128, M, G);
            if ((int64_t)blockIdx.x == c_hi) {
                for (int64_t slot = (c_hi - c_lo + 1) * 4 + q; slot < p.nparts; slot += 4)
                    neutral_head(p.part + (((int64_t)b * p.nparts + slot) * kKvpHeads + cg) * kKvpPart, lane);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

struct KvpPlan { int grid, tiles_m, nparts; };

static KvpPlan kvp_plan(int B, int64_t N) {
    KvpPlan pl;
    const int64_t rows = (int64_t)B * N;
    pl.tiles_m = (int)((rows + 127) / 128);
    pl.grid = sm_count() < pl.tiles_m ? sm_count() : pl.tiles_m;
    int64_t worst = 1;
    for (int b = 0; b < B; ++b) {
        const int64_t c_lo = tile_owner((int64_t)b * N / 128, pl.tiles_m, pl.grid);
        const int64_t c_hi = tile_owner(((int64_t)(b + 1) * N - 1) / 128, pl.tiles_m, pl.grid);
        if (c_hi - c_lo + 1 > worst) worst = c_hi - c_lo + 1;
    }
    pl.nparts = (int)worst * 4;
    return pl;
}

}  // namespace
}  // namespace ltu

using namespace ltu;

extern "C" int ltu_kv_project_reduce_supported(int C, int heads, int64_t N) {
    static const bool on = [] { const char* e = getenv("LTU_KV_PROJECT"); return !(e && e[0] == '0'); }();   // A/B switch
    return (on && C == 128 && heads == 4 && N > 0 && N % 32 == 0) ? 1 : 0;
}

extern "C" size_t ltu_kv_project_reduce_workspace(int B, int64_t N) {
    const KvpPlan pl = kvp_plan(B, N);
    const size_t a = (size_t)B * pl.nparts * kKvpHeads * kKvpPart * sizeof(float), b = kv_project2_workspace(B, N);
    return a > b ? a : b;
}

// ctx fp32 [B][4][32][32] = softmax over the N tokens of (x Wk^T + bk), transposed, times (x Wv^T + bv), per head:
// x bf16 [B][N][128], w_kv bf16 [256][128] = rows 0-127 Wk, rows 128-255 Wv (the nn.Linear weights rounded to bf16),
// bias fp32 [256]; K and V are rounded to bf16 exactly as the separate projection stores them.  N % 32 == 0.
// wo_bf16 / w_out (both or neither): the merge kernel also writes W_b = blockdiag(ctx_b) Wo^T, bf16 [B][128][128]
extern "C" int ltu_kv_project_reduce(const void* x, const void* w_kv, const float* bias, float* ctx, void* workspace,
                                     size_t workspace_bytes, int B, int64_t N, const void* wo_bf16, void* w_out,
                                     ltu_stream_t stream) {
    LTU_ARG_CHECK((wo_bf16 == nullptr) == (w_out == nullptr), "kv_project_reduce: wo_bf16 and w_out go together");
    LTU_ARG_CHECK(x && w_kv && bias && ctx && workspace, "kv_project_reduce: null pointer");
    LTU_ARG_CHECK(B > 0 && N > 0 && N % 32 == 0 && (int64_t)B * N < ((int64_t)1 << 31) - 256, "kv_project_reduce: bad shape (N %% 32 == 0)");
    LTU_ARG_CHECK((((uintptr_t)x | (uintptr_t)w_kv | (uintptr_t)workspace) & 15) == 0, "kv_project_reduce: pointers must be 16-byte aligned");
    const KvpPlan pl = kvp_plan(B, N);
    LTU_ARG_CHECK(workspace_bytes >= ltu_kv_project_reduce_workspace(B, N), "kv_project_reduce: workspace too small");
    LTU_ARG_CHECK(B <= 30, "kv_project_reduce: at most 30 samples per call (sample mask of a warp)");
    if (kv_project2_enabled()) return kv_project2_launch(x, w_kv, bias, ctx, workspace, B, N, wo_bf16, w_out, (cudaStream_t)stream);
    CUtensorMap tx, tw;
    int rc;
    if ((rc = make_tmap_bf16_2d(&tx, x, (uint64_t)B * N, 128, 128)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw, w_kv, 256, 128, 256)) != LTU_OK) return rc;
    KvpParams p;
    p.bias = bias; p.part = (float*)workspace; p.rows = (int64_t)B * N; p.N = N; p.B = B;
    p.tiles_m = pl.tiles_m; p.nparts = pl.nparts;
    const size_t smem = 1024 + kKvpOffTail + sizeof(KvpTail);
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(kv_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured_dev = dev;
    }
    cudaError_t e = launch_pdl(kv_project_kernel, dim3(pl.grid), dim3(kKvpThreads), smem, (cudaStream_t)stream, tx, tw, p);
    if (e != cudaSuccess) { set_error("kv_project_reduce: launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    rc = kv_combine_launch((const float*)workspace, ctx, kKvpHeads, B, pl.nparts, (cudaStream_t)stream, wo_bf16, w_out);
    if (rc != LTU_OK) return rc;
    count_launch(2);
    return LTU_OK;
}
