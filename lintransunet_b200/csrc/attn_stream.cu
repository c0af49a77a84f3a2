// Family 1, bf16 path, heads in {4, 8}: the linear-attention core as TMA-fed STREAMING kernels in which every warp
// is autonomous -- no block-wide barrier anywhere in the token loop.
//
//   * one producer warp issues `cp.async.bulk.tensor.3d` box loads ([T tokens] x [64 channels = two heads], SWIZZLE_128B,
//     tensor map over the [B][N][C] view with row stride ld, so a fused QKV buffer is read in place and rows past the
//     end of a sample arrive as zeros) into an NSTAGE ring; one full/empty mbarrier pair per (stage, head pair);
//   * consumer warp (head, sub) owns the 32 rows x 64 bytes of its head inside a box: kv_reduce turns K into
//     P = 2^(k*log2e - r_j) IN PLACE (r_j: column max of the warp's first tile; softmax over N is shift invariant, bf16
//     shares fp32's exponent range; an exact warp-local rescale path covers data that runs away from the reference),
//     then ctx += P^T V with ldmatrix.trans + mma.sync m16n8k16 (column sums from an all-ones B tile); q_readout keeps
//     the row softmax in the A fragments and multiplies by register-resident ctx fragments, writes the tile back in
//     place and stores 64-byte row segments with 16-byte vector stores;
//   * the per-lane cp.async address generation, the three __syncthreads per tile and the lock-step of eight warps of
//     the first version (attn_tc.cu: 39 % issue slots, 24 % warps active, 55 % DRAM at 4.4 TB/s) are gone.
// mma.sync is kept on purpose: the state is 32x32 per head and must be rescaled in registers; a 128-row UMMA tile
// would be 3/4 padding.  Partial states are merged by the same fixed-order kv_combine kernel (attn_kernels.cu).
// Every kernel starts with griddepcontrol.launch_dependents / .wait, so a launch with programmatic stream
// serialisation overlaps its prologue with the tail of its predecessor (and is a no-op otherwise).
#include <cuda.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);
int kv_chunks_per_batch_host(int B, int64_t N);   // attn_kernels.cu (per-sample split: depends on N only)
int kv_combine_launch(const float* ws, float* ctx, int heads, int B, int nparts, cudaStream_t st, const void* wo = nullptr,
                      void* wout = nullptr);

namespace {

constexpr int kPart = 32 * 32 + 64;      // ctx[32][32], m[32], s[32]
constexpr float kL2e = 1.4426950408889634f;
constexpr int kStages = 3;
constexpr int kConsumers = 8;            // consumer warps; warp 8 is the producer
constexpr int kThreads = (kConsumers + 1) * 32;

__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a [rows][128 B] SWIZZLE_128B box (1024-B aligned)
__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// ---------------------------------------------------------------------------------------------------- kv_reduce
// grid (chunks, B); partial index ((b*chunks + chunk)*WPH + sub)*HEADS + hd, as in attn_tc.cu / attn_kernels.cu
template <int HEADS>
__global__ void __launch_bounds__(kThreads, 2)
kv_stream_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                 const bf16* __restrict__ Kraw, int64_t ld, float* __restrict__ part, int64_t N, int chunks,
                 int tiles32_per_chunk) {
    constexpr int WPH = kConsumers / HEADS, T = 32 * WPH, HP = HEADS / 2;
    constexpr uint32_t BOX = T * 128, STAGE = 2 * HP * BOX;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + kStages * STAGE;               // full[kStages][HP], empty[kStages][HP]
    float* scratch = reinterpret_cast<float*>(smem_raw + (bars - smem_u32(smem_raw)) + 2 * kStages * HP * 8);   // [8][32]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y, chunk = blockIdx.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kStages * HP; ++i) {
            mbar_init(bars + 8 * i, 1);
            mbar_init(bars + 8 * (kStages * HP + i), 2 * WPH);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_prologue();

    const int64_t row_begin = (int64_t)chunk * tiles32_per_chunk * 32;
    int64_t row_end = row_begin + (int64_t)tiles32_per_chunk * 32;
    if (row_end > N) row_end = N;
    const int ntiles = row_end > row_begin ? (int)((row_end - row_begin + T - 1) / T) : 0;

    if (warp == kConsumers) {                                    // ---------------- producer
        if (lane == 0) {
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages;
                const int row = (int)(row_begin + (int64_t)t * T);
                for (int hp = 0; hp < HP; ++hp) {
                    if (t >= kStages) mbar_wait(bars + 8 * (kStages * HP + s * HP + hp), ((t / kStages) - 1) & 1);
                    const uint32_t full = bars + 8 * (s * HP + hp);
                    mbar_expect_tx(full, 2 * BOX);
                    tma_load_3d(base + s * STAGE + hp * BOX, &tmK, hp * 64, row, b, full);
                    tma_load_3d(base + s * STAGE + (HP + hp) * BOX, &tmV, hp * 64, row, b, full);
                }
            }
        }
        return;
    }

    // ---------------- consumers: warp = (head, sub)
    const int hd = warp % HEADS, sub = warp / HEADS, hp = hd >> 1, half = hd & 1;
    const int g = lane >> 2, cc = lane & 3, mi = lane >> 3, lr = lane & 7;
    float* out = part + ((((int64_t)b * chunks + chunk) * WPH + sub) * HEADS + hd) * kPart;
    float* myscr = scratch + warp * 32;

    float acc[2][4][4], accs[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) accs[mt][i] = 0.f;
    }
    float rj[8];                                                 // reference (log2 units) of columns cc*8 .. cc*8+7
#pragma unroll
    for (int c = 0; c < 8; ++c) rj[c] = 0.f;
    bool have_ref = false;
    const uint32_t ones = g == 0 ? 0x3F803F80u : 0u;            // B tile whose column 0 is 1.0

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % kStages;
        const int64_t row0 = row_begin + (int64_t)t * T + sub * 32;
        int valid = (int)(row_end - row0);
        valid = valid < 0 ? 0 : (valid > 32 ? 32 : valid);
        mbar_wait(bars + 8 * (s * HP + hp), (t / kStages) & 1);
        const uint32_t tK = base + s * STAGE + hp * BOX, tV = base + s * STAGE + (HP + hp) * BOX;
        if (valid > 0) {
            // ---- K -> P in place: lane owns rows sub*32 + g + 8i (i < 4), 16-byte chunk cc of its head
            uint32_t addr[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) addr[i] = tK + swz(sub * 32 + g + 8 * i, half * 4 + cc);
            if (!have_ref) {                                     // reference = column max of the warp's first tile
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
                        for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], k[c]);
                    }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
                    rj[c] = mx[c] * kL2e;
                }
                have_ref = true;
            }
            float dmax = -INFINITY;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = g + 8 * i < valid;
                uint4 v;
                float k[8];
                lds16(addr[i], v);
                unpack(v, k);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float d = fmaf(k[c], kL2e, -rj[c]);
                    if (ok) dmax = fmaxf(dmax, d);
                    k[c] = ok ? ex2(d) : 0.f;
                }
                sts16(addr[i], pack(k));
            }
            if (__any_sync(0xffffffffu, dmax > 64.f)) {
                // ---- rare: the data ran away from the reference.  Raise r_j to this tile's column max, rescale the
                // running state by 2^(r_old - r_new) and redo the tile's P from K in global memory.
                const bf16* Kg = Kraw + ((int64_t)b * N + row0) * ld + hd * 32 + cc * 8;
                float mx[8], f[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) mx[c] = -INFINITY;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (g + 8 * i < valid) {
                        float k[8];
                        load_vec(Kg + (int64_t)(g + 8 * i) * ld, k);
#pragma unroll
                        for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], k[c] * kL2e);
                    }
#pragma unroll
                for (int c = 0; c < 8; ++c) {
#pragma unroll
                    for (int o = 4; o < 32; o <<= 1) mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
                    const float m = fmaxf(rj[c], mx[c]);
                    f[c] = ex2(rj[c] - m);
                    rj[c] = m;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const bool ok = g + 8 * i < valid;
                    float k[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) k[c] = 0.f;
                    if (ok) load_vec(Kg + (int64_t)(g + 8 * i) * ld, k);
#pragma unroll
                    for (int c = 0; c < 8; ++c) k[c] = ok ? ex2(fmaf(k[c], kL2e, -rj[c])) : 0.f;
                    sts16(addr[i], pack(k));
                }
                if (g == 0) {
#pragma unroll
                    for (int c = 0; c < 8; ++c) myscr[cc * 8 + c] = f[c];
                }
                __syncwarp();
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float f0 = myscr[mt * 16 + g], f1 = myscr[mt * 16 + g + 8];
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        acc[mt][nt][0] *= f0; acc[mt][nt][1] *= f0; acc[mt][nt][2] *= f1; acc[mt][nt][3] *= f1;
                    }
                    accs[mt][0] *= f0; accs[mt][1] *= f0; accs[mt][2] *= f1; accs[mt][3] *= f1;
                }
            }
            __syncwarp();                                        // P visible to the whole warp
            // ---- ctx[j][e] += sum_n P[n][j] V[n][e]   (M = j, N = e, K = the warp's 32 tokens)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
                const int n0 = sub * 32 + ks * 16;
                uint32_t a[2][4], bq[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)       // A^T: matrices (k lo, m lo), (k lo, m hi), (k hi, m lo), (k hi, m hi)
                    ldsm4t(tK + swz(n0 + lr + 8 * (mi >> 1), half * 4 + mt * 2 + (mi & 1)), a[mt]);
#pragma unroll
                for (int np = 0; np < 2; ++np)       // B: matrices (k lo, n lo), (k hi, n lo), (k lo, n hi), (k hi, n hi)
                    ldsm4t(tV + swz(n0 + lr + 8 * (mi & 1), half * 4 + np * 2 + (mi >> 1)), bq[np]);
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt)
                        mma(acc[mt][nt], a[mt], bq[nt >> 1][(nt & 1) * 2], bq[nt >> 1][(nt & 1) * 2 + 1]);
                    mma(accs[mt], a[mt], ones, ones);
                }
            }
            fence_async_smem();                                  // generic writes (P) before the next TMA into this box
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (kStages * HP + s * HP + hp));
    }

    // ---- partial state: ctx[j][e], m[j] (natural-log units), s[j]
    const int tq = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int j = mt * 16 + g, e = nt * 8 + 2 * tq;
            *reinterpret_cast<float2*>(out + j * 32 + e) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
            *reinterpret_cast<float2*>(out + (j + 8) * 32 + e) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
    if (g == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) out[1024 + cc * 8 + c] = have_ref ? rj[c] * (1.f / kL2e) : -INFINITY;
    }
    if (tq == 0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            out[1056 + mt * 16 + g] = accs[mt][0];
            out[1056 + mt * 16 + g + 8] = accs[mt][2];
        }
    }
}

// ---------------------------------------------------------------------------------------------------- q_readout
// grid (ctas, B): CTA x handles tiles [x*tiles_per_cta, ...) of T = 32*WPH rows
template <int HEADS>
__global__ void __launch_bounds__(kThreads, 2)
q_stream_kernel(const __grid_constant__ CUtensorMap tmQ, const float* __restrict__ ctx, bf16* __restrict__ O, int64_t ldo,
                int64_t N, int tiles_per_cta) {
    constexpr int WPH = kConsumers / HEADS, T = 32 * WPH, HP = HEADS / 2;
    constexpr uint32_t BOX = T * 128, STAGE = HP * BOX;
    constexpr int NST = 2 * kStages;                             // half the bytes per stage of kv_reduce: twice the stages
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bars = base + NST * STAGE;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NST * HP; ++i) {
            mbar_init(bars + 8 * i, 1);
            mbar_init(bars + 8 * (NST * HP + i), 2 * WPH);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_prologue();

    const int64_t total_tiles = (N + T - 1) / T;
    const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
    int ntiles = (int)(total_tiles - tile0 < tiles_per_cta ? total_tiles - tile0 : tiles_per_cta);
    if (ntiles < 0) ntiles = 0;

    if (warp == kConsumers) {                                    // ---------------- producer
        if (lane == 0) {
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % NST;
                const int row = (int)((tile0 + t) * T);
                for (int hp = 0; hp < HP; ++hp) {
                    if (t >= NST) mbar_wait(bars + 8 * (NST * HP + s * HP + hp), ((t / NST) - 1) & 1);
                    const uint32_t full = bars + 8 * (s * HP + hp);
                    mbar_expect_tx(full, BOX);
                    tma_load_3d(base + s * STAGE + hp * BOX, &tmQ, hp * 64, row, b, full);
                }
            }
        }
        return;
    }

    const int hd = warp % HEADS, sub = warp / HEADS, hp = hd >> 1, half = hd & 1;
    const int g = lane >> 2, tq = lane & 3, mi = lane >> 3, lr = lane & 7;
    bf16* Ob = O + (int64_t)b * N * ldo + hd * 32 + tq * 8;

    // B fragments of ctx' = ctx / sqrt(32) for this head: (K = j) x (N = e)
    uint32_t bf[2][4][2];
    {
        const float* cb = ctx + ((int64_t)b * HEADS + hd) * 1024;
        const float sc = 0.17677669529663687f;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int j = ks * 16 + 2 * tq, e = nt * 8 + g;
                bf[ks][nt][0] = pack_bf16x2(cb[j * 32 + e] * sc, cb[(j + 1) * 32 + e] * sc);
                bf[ks][nt][1] = pack_bf16x2(cb[(j + 8) * 32 + e] * sc, cb[(j + 9) * 32 + e] * sc);
            }
    }

    for (int t = 0; t < ntiles; ++t) {
        const int s = t % NST;
        const int64_t row0 = (tile0 + t) * T + sub * 32;
        int valid = (int)(N - row0);
        valid = valid < 0 ? 0 : (valid > 32 ? 32 : valid);
        mbar_wait(bars + 8 * (s * HP + hp), (t / NST) & 1);
        const uint32_t tQ = base + s * STAGE + hp * BOX;
        if (valid > 0) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const int r0 = sub * 32 + mt * 16;
                uint32_t a[2][4];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)   // A: matrices (rows lo, k lo), (rows hi, k lo), (rows lo, k hi), (rows hi, k hi)
                    ldsm4(tQ + swz(r0 + lr + 8 * (mi & 1), half * 4 + ks * 2 + (mi >> 1)), a[ks]);
                // rows g (a[.][0], a[.][2]) and g+8 (a[.][1], a[.][3]): 8 of the 32 head features each
                float x0[8], x1[8];
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    x0[ks * 4 + 0] = __uint_as_float(a[ks][0] << 16); x0[ks * 4 + 1] = __uint_as_float(a[ks][0] & 0xffff0000u);
                    x0[ks * 4 + 2] = __uint_as_float(a[ks][2] << 16); x0[ks * 4 + 3] = __uint_as_float(a[ks][2] & 0xffff0000u);
                    x1[ks * 4 + 0] = __uint_as_float(a[ks][1] << 16); x1[ks * 4 + 1] = __uint_as_float(a[ks][1] & 0xffff0000u);
                    x1[ks * 4 + 2] = __uint_as_float(a[ks][3] << 16); x1[ks * 4 + 3] = __uint_as_float(a[ks][3] & 0xffff0000u);
                }
                float m0 = x0[0], m1 = x1[0];
#pragma unroll
                for (int i = 1; i < 8; ++i) { m0 = fmaxf(m0, x0[i]); m1 = fmaxf(m1, x1[i]); }
                m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
                m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
                float s0 = 0.f, s1 = 0.f;
                const float c0 = m0 * kL2e, c1 = m1 * kL2e;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    x0[i] = ex2(fmaf(x0[i], kL2e, -c0)); s0 += x0[i];
                    x1[i] = ex2(fmaf(x1[i], kL2e, -c1)); s1 += x1[i];
                }
                s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
                s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    a[ks][0] = pack_bf16x2(x0[ks * 4 + 0], x0[ks * 4 + 1]); a[ks][2] = pack_bf16x2(x0[ks * 4 + 2], x0[ks * 4 + 3]);
                    a[ks][1] = pack_bf16x2(x1[ks * 4 + 0], x1[ks * 4 + 1]); a[ks][3] = pack_bf16x2(x1[ks * 4 + 2], x1[ks * 4 + 3]);
                }
                float d[4][4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    d[nt][0] = d[nt][1] = d[nt][2] = d[nt][3] = 0.f;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) mma(d[nt], a[ks], bf[ks][nt][0], bf[ks][nt][1]);
                }
                const float i0 = 1.f / s0, i1 = 1.f / s1;
                __syncwarp();                                    // every lane has read its A fragments of this m-tile
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {                 // back in place: rows g / g+8, columns nt*8 + 2*tq (+1)
                    const uint32_t p0 = tQ + swz(r0 + g, half * 4 + nt) + 4 * tq;
                    const uint32_t p1 = tQ + swz(r0 + g + 8, half * 4 + nt) + 4 * tq;
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(p0), "r"(pack_bf16x2(d[nt][0] * i0, d[nt][1] * i0)) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(p1), "r"(pack_bf16x2(d[nt][2] * i1, d[nt][3] * i1)) : "memory");
                }
            }
            __syncwarp();
            // ---- 64-byte row segments out: lane -> rows g + 8i, 16-byte chunk tq of this head
            bf16* go = Ob + row0 * ldo;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = g + 8 * i;
                if (r < valid) {
                    uint4 v;
                    lds16(tQ + swz(sub * 32 + r, half * 4 + tq), v);
                    *reinterpret_cast<uint4*>(go + (int64_t)r * ldo) = v;
                }
            }
            fence_async_smem();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (NST * HP + s * HP + hp));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// bf16 [B][N][C] view with row stride ld (elements): box = [1][box_rows][64 channels], SWIZZLE_128B, zero fill
int make_tmap_tokens(CUtensorMap* map, const void* base, int B, int64_t N, int C, int64_t ld, int box_rows) {
    static EncodeTiledFn fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return (EncodeTiledFn)ptr;
    }();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LTU_ERR_ARG; }
    const cuuint64_t gdim[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)B};
    const cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)N * (cuuint64_t)ld * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (token view) failed (CUresult %d)", (int)r); return LTU_ERR_ARG; }
    return LTU_OK;
}

template <int HEADS>
int kv_stream_launch(const bf16* k, const bf16* v, int64_t ld, float* ws, int B, int64_t N, int chunks, int tiles32_per_chunk,
                     cudaStream_t st) {
    constexpr int WPH = kConsumers / HEADS, T = 32 * WPH, HP = HEADS / 2;
    constexpr size_t STAGE = (size_t)2 * HP * T * 128;
    const size_t smem = 1024 + kStages * STAGE + 2 * kStages * HP * 8 + kConsumers * 32 * 4;
    CUtensorMap tk, tv;
    int rc;
    if ((rc = make_tmap_tokens(&tk, k, B, N, HEADS * 32, ld, T)) != LTU_OK) return rc;
    if ((rc = make_tmap_tokens(&tv, v, B, N, HEADS * 32, ld, T)) != LTU_OK) return rc;
    static thread_local int conf = -1;
    int dev; cudaGetDevice(&dev);
    if (conf != dev) { cudaFuncSetAttribute(kv_stream_kernel<HEADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); conf = dev; }
    cudaError_t e = launch_pdl(kv_stream_kernel<HEADS>, dim3(chunks, B), dim3(kThreads), smem, st, tk, tv, k, ld, ws, N, chunks,
                               tiles32_per_chunk);
    if (e != cudaSuccess) { set_error("kv_reduce (stream): launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    return LTU_OK;
}

template <int HEADS>
int q_stream_launch(const bf16* q, int64_t ldq, const float* ctx, bf16* out, int64_t ldo, int B, int64_t N, cudaStream_t st) {
    constexpr int WPH = kConsumers / HEADS, T = 32 * WPH, HP = HEADS / 2;
    constexpr size_t STAGE = (size_t)HP * T * 128;
    constexpr int NST = 2 * kStages;
    const size_t smem = 1024 + NST * STAGE + 2 * NST * HP * 8;
    const int64_t tiles = (N + T - 1) / T;
    int64_t want = ceil_div64(2 * (int64_t)sm_count(), B);       // two resident CTAs per SM
    int64_t ctas = tiles < want ? tiles : want;
    if (ctas < 1) ctas = 1;
    const int tiles_per_cta = (int)ceil_div64(tiles, ctas);
    ctas = ceil_div64(tiles, tiles_per_cta);
    CUtensorMap tq;
    int rc;
    if ((rc = make_tmap_tokens(&tq, q, B, N, HEADS * 32, ldq, T)) != LTU_OK) return rc;
    static thread_local int conf = -1;
    int dev; cudaGetDevice(&dev);
    if (conf != dev) { cudaFuncSetAttribute(q_stream_kernel<HEADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); conf = dev; }
    cudaError_t e = launch_pdl(q_stream_kernel<HEADS>, dim3((unsigned)ctas, B), dim3(kThreads), smem, st, tq, ctx, out, ldo, N,
                               tiles_per_cta);
    if (e != cudaSuccess) { set_error("q_readout (stream): launch failed: %s", cudaGetErrorString(e)); return (int)e; }
    count_launch(1);
    return LTU_OK;
}

}  // namespace

// entry points used by attn_kernels.cu for dtype == bf16, heads in {4, 8}, 16-byte aligned rows
int kv_reduce_bf16_stream(const void* k, const void* v, int64_t ld, float* ctx, void* ws, int B, int64_t N, int heads,
                          cudaStream_t st, const void* wo, void* wout) {
    const int chunks = kv_chunks_per_batch_host(B, N);
    const int tiles32_per_chunk = (int)ceil_div64(ceil_div64(N, 32), chunks);
    int rc = heads == 8 ? kv_stream_launch<8>((const bf16*)k, (const bf16*)v, ld, (float*)ws, B, N, chunks, tiles32_per_chunk, st)
                        : kv_stream_launch<4>((const bf16*)k, (const bf16*)v, ld, (float*)ws, B, N, chunks, tiles32_per_chunk, st);
    if (rc != LTU_OK) return rc;
    rc = kv_combine_launch((const float*)ws, ctx, heads, B, chunks * (kConsumers / heads), st, wo, wout);
    if (rc != LTU_OK) return rc;
    count_launch(2);
    return LTU_OK;
}

int q_readout_bf16_stream(const void* q, int64_t ldq, const float* ctx, void* out, int64_t ldo, int B, int64_t N, int heads,
                          cudaStream_t st) {
    return heads == 8 ? q_stream_launch<8>((const bf16*)q, ldq, ctx, (bf16*)out, ldo, B, N, st)
                      : q_stream_launch<4>((const bf16*)q, ldq, ctx, (bf16*)out, ldo, B, N, st);
}

}  // namespace ltu
