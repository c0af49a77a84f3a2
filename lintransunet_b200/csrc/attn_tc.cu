// Family 1, bf16 path: the linear-attention core with the 32x32 per-head contractions on the
// tensor pipe (mma.sync m16n8k16 bf16, fp32 accumulate) so that the kernels are bound by HBM, not
// by CUDA-core FMA issue (AI = 16 flop/B needs ~100 TFLOP/s at 6.5 TB/s).
// tcgen05 is not used here on purpose: the accumulator is a 32x32 state per head that must be
// rescaled/merged in registers, and M=64/128 UMMA tiles would be mostly padding.
//
// kv_reduce : K,V tiles (32 tokens x C) are staged with cp.async (double buffered, padded rows).
//             K is turned IN PLACE into P = exp2(k*log2e - r_j) in bf16, where r_j is a per-column
//             reference (column max of the CTA's first tile; softmax over N is shift invariant and
//             bf16/fp32 share the exponent range, so r_j only has to stay within 2^+-64 of the data --
//             checked per tile, with an exact rescale path).  ctx += P^T V by ldmatrix.trans + mma;
//             the column sums s_j come from an extra all-ones B tile.  Partials are merged by the
//             same fixed-order kv_combine kernel as the fp32 path.
// q_readout : Q tiles staged the same way; the row softmax lives in the A fragments (row max/sum
//             by two quad shuffles), out = P ctx' / rowsum by mma against register-resident ctx
//             fragments, written back in place and stored with coalesced 16-byte rows.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

constexpr int kTT = 32;                  // tokens per tile (== kTileTokens of attn_kernels.cu)
constexpr int kPartF = 32 * 32 + 64;     // ctx[32][32], m[32], s[32]
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

// stage kTT rows x C bf16 into padded smem rows (row stride LDS elements); rows >= nrows are zero
template <int C, int LDS>
__device__ __forceinline__ void stage_rows(bf16* smem, const bf16* src, int64_t ld, int64_t row0, int64_t nrows) {
    constexpr int CPR = C / 8;
    for (int i = threadIdx.x; i < kTT * CPR; i += 256) {
        const int r = i / CPR, c = i - r * CPR;
        const int64_t row = row0 + r;
        const bool ok = row < nrows;
        cp_async16_zfill(smem + r * LDS + c * 8, src + (ok ? row : row0) * ld + c * 8, ok ? 16 : 0);
    }
}

template <int HEADS, int NSTAGE>
__global__ void __launch_bounds__(256, 2)
kv_reduce_mma_kernel(const bf16* __restrict__ K, const bf16* __restrict__ V, int64_t ld, float* __restrict__ part,
                     int64_t N, int chunks, int tiles_per_chunk) {
    constexpr int C = HEADS * 32, CPR = C / 8, RPP = 256 / CPR, NCH = kTT / RPP, LDS = C + 8, WPH = 8 / HEADS;
    static_assert(NCH >= 1, "tile must give every thread at least one 16-byte chunk");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* sK = reinterpret_cast<bf16*>(smem_raw);                 // [NSTAGE][kTT][LDS]  (K, then P in place)
    bf16* sV = sK + NSTAGE * kTT * LDS;                           // [NSTAGE][kTT][LDS]
    float* sRed = reinterpret_cast<float*>(sV + NSTAGE * kTT * LDS);   // [RPP][C]
    float* sRef = sRed + RPP * C;                                 // [C] reference r_j * log2e
    float* sScale = sRef + C;                                     // [C] rescale factors (slow path)

    const int b = blockIdx.y, chunk = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hd = warp % HEADS, sub = warp / HEADS;
    const int cc = tid % CPR, r0 = tid / CPR;
    const bf16* Kb = K + (int64_t)b * N * ld;
    const bf16* Vb = V + (int64_t)b * N * ld;
    float* out = part + ((((int64_t)b * chunks + chunk) * WPH + sub) * HEADS + hd) * kPartF;

    const int64_t tile0 = (int64_t)chunk * tiles_per_chunk;
    int64_t ntiles = ceil_div64(N, kTT) - tile0;
    if (ntiles > tiles_per_chunk) ntiles = tiles_per_chunk;
    if (ntiles <= 0) {                                            // cannot happen with the host-side split
        for (int i = lane; i < 1024; i += 32) out[i] = 0.f;
        out[1024 + lane] = -INFINITY;
        out[1056 + lane] = 0.f;
        return;
    }

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
    float rj[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) rj[c] = 0.f;
    const uint32_t ones = (lane >> 2) == 0 ? 0x3F803F80u : 0u;     // B tile with column 0 = 1.0 (bf16 pairs)

    // NSTAGE-deep cp.async ring: NSTAGE-1 tiles are in flight while one is processed.  A thread stages
    // exactly the chunks it later transforms: rows r0 + i*RPP, 16-byte column cc (offsets are loop invariant).
    const int64_t goff0 = (int64_t)r0 * ld + cc * 8;
    const int64_t gstep = (int64_t)RPP * ld;
    const int soff0 = r0 * LDS + cc * 8;
    auto stage = [&](int nb, int64_t trow0) {
        const bf16* gk = Kb + trow0 * ld + goff0;
        const bf16* gv = Vb + trow0 * ld + goff0;
        bf16* dk = sK + nb * kTT * LDS + soff0;
        bf16* dv = sV + nb * kTT * LDS + soff0;
        const int64_t left = N - trow0 - r0;                   // rows r0 + i*RPP with i*RPP < left exist
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const bool ok = (int64_t)i * RPP < left;
            cp_async16_zfill(dk + i * RPP * LDS, ok ? gk + i * gstep : Kb, ok ? 16 : 0);
            cp_async16_zfill(dv + i * RPP * LDS, ok ? gv + i * gstep : Vb, ok ? 16 : 0);
        }
    };
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (s < ntiles) stage(s, (tile0 + s) * kTT);
        cp_async_commit();
    }

    for (int64_t t = 0; t < ntiles; ++t) {
        const int buf = (int)(t % NSTAGE);
        if (t + NSTAGE - 1 < ntiles) stage((int)((t + NSTAGE - 1) % NSTAGE), (tile0 + t + NSTAGE - 1) * kTT);
        cp_async_commit();
        cp_async_wait<NSTAGE - 1>();
        __syncthreads();                                          // tile t landed
        bf16* tK = sK + buf * kTT * LDS;
        const bf16* tV = sV + buf * kTT * LDS;
        const int64_t row0 = (tile0 + t) * kTT;
        const int valid = (int)((N - row0) < kTT ? (N - row0) : kTT);

        if (t == 0) {                                             // reference = column max of the first tile
            float mx[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) mx[c] = -INFINITY;
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int row = r0 + i * RPP;
                if (row < valid) {
                    float k[8];
                    unpack8(*reinterpret_cast<const uint4*>(tK + row * LDS + cc * 8), k);
#pragma unroll
                    for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], k[c]);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) sRed[r0 * C + cc * 8 + c] = mx[c];
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float m = -INFINITY;
                for (int r = 0; r < RPP; ++r) m = fmaxf(m, sRed[r * C + cc * 8 + c]);
                rj[c] = m * kLog2e;
                if (r0 == 0) sRef[cc * 8 + c] = rj[c];
            }
        }

        // ---- K -> P in place
        float dmax = -INFINITY;
        if (valid == kTT) {                                       // every tile but the last one of the sequence
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                uint4* ptr = reinterpret_cast<uint4*>(tK + (r0 + i * RPP) * LDS + cc * 8);
                float k[8];
                unpack8(*ptr, k);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float d = fmaf(k[c], kLog2e, -rj[c]);
                    dmax = fmaxf(dmax, d);
                    k[c] = ex2f(d);
                }
                uint4 o;
                o.x = pack_bf16x2(k[0], k[1]); o.y = pack_bf16x2(k[2], k[3]);
                o.z = pack_bf16x2(k[4], k[5]); o.w = pack_bf16x2(k[6], k[7]);
                *ptr = o;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int row = r0 + i * RPP;
                uint4* ptr = reinterpret_cast<uint4*>(tK + row * LDS + cc * 8);
                float k[8];
                unpack8(*ptr, k);
                const bool ok = row < valid;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float d = fmaf(k[c], kLog2e, -rj[c]);
                    if (ok) dmax = fmaxf(dmax, d);
                    k[c] = ok ? ex2f(d) : 0.f;
                }
                uint4 o;
                o.x = pack_bf16x2(k[0], k[1]); o.y = pack_bf16x2(k[2], k[3]);
                o.z = pack_bf16x2(k[4], k[5]); o.w = pack_bf16x2(k[6], k[7]);
                *ptr = o;
            }
        }
        if (__syncthreads_or(dmax > 64.f)) {
            // ---- rare: the data ran away from the reference.  Raise r_j to the tile's column max,
            // rescale the running state by 2^(r_old - r_new) and redo this tile's P from global K.
            float mx[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) mx[c] = -INFINITY;
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int row = r0 + i * RPP;
                if (row < valid) {
                    float k[8];
                    unpack8(*reinterpret_cast<const uint4*>(Kb + (row0 + row) * ld + cc * 8), k);
#pragma unroll
                    for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], k[c] * kLog2e);
                }
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) sRed[r0 * C + cc * 8 + c] = mx[c];
            __syncthreads();
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                float m = rj[c];
                for (int r = 0; r < RPP; ++r) m = fmaxf(m, sRed[r * C + cc * 8 + c]);
                if (r0 == 0) { sScale[cc * 8 + c] = ex2f(rj[c] - m); sRef[cc * 8 + c] = m; }
                rj[c] = m;
            }
#pragma unroll
            for (int i = 0; i < NCH; ++i) {
                const int row = r0 + i * RPP;
                const bool ok = row < valid;
                float k[8];
                if (ok) unpack8(*reinterpret_cast<const uint4*>(Kb + (row0 + row) * ld + cc * 8), k);
#pragma unroll
                for (int c = 0; c < 8; ++c) k[c] = ok ? ex2f(fmaf(k[c], kLog2e, -rj[c])) : 0.f;
                uint4 o;
                o.x = pack_bf16x2(k[0], k[1]); o.y = pack_bf16x2(k[2], k[3]);
                o.z = pack_bf16x2(k[4], k[5]); o.w = pack_bf16x2(k[6], k[7]);
                *reinterpret_cast<uint4*>(tK + row * LDS + cc * 8) = o;
            }
            __syncthreads();
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
                const float f0 = sScale[hd * 32 + mt * 16 + (lane >> 2)], f1 = sScale[hd * 32 + mt * 16 + (lane >> 2) + 8];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    acc[mt][nt][0] *= f0; acc[mt][nt][1] *= f0; acc[mt][nt][2] *= f1; acc[mt][nt][3] *= f1;
                }
                accs[mt][0] *= f0; accs[mt][1] *= f0; accs[mt][2] *= f1; accs[mt][3] *= f1;
            }
        }

        // ---- ctx[j][e] += sum_n P[n][j] V[n][e]   (M = j, N = e, K = tokens)
        const int mi = lane >> 3, lr = lane & 7;
#pragma unroll
        for (int ks = 0; ks < kTT / 16; ++ks) {
            if ((ks % WPH) != sub) continue;
            const int n0 = ks * 16;
            uint32_t a[2][4], bq[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)     // A^T: matrices (k lo, m lo), (k lo, m hi), (k hi, m lo), (k hi, m hi)
                ldsm_x4_t(smem_u32_generic(tK + (n0 + lr + 8 * (mi >> 1)) * LDS + hd * 32 + mt * 16 + 8 * (mi & 1)), a[mt]);
#pragma unroll
            for (int np = 0; np < 2; ++np)     // B: matrices (k lo, n lo), (k hi, n lo), (k lo, n hi), (k hi, n hi)
                ldsm_x4_t(smem_u32_generic(tV + (n0 + lr + 8 * (mi & 1)) * LDS + hd * 32 + np * 16 + 8 * (mi >> 1)), bq[np]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma16816(acc[mt][nt], a[mt], bq[nt >> 1][(nt & 1) * 2], bq[nt >> 1][(nt & 1) * 2 + 1]);
                mma16816(accs[mt], a[mt], ones, ones);
            }
        }
        __syncthreads();                                          // buffers of tile t are free again
    }
    cp_async_wait<0>();

    // ---- partial state: ctx[j][e], m[j] (natural-log units), s[j]
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int j = mt * 16 + g, e = nt * 8 + 2 * tq;
            *reinterpret_cast<float2*>(out + j * 32 + e) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
            *reinterpret_cast<float2*>(out + (j + 8) * 32 + e) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
    out[1024 + lane] = sRef[hd * 32 + lane] * (1.f / kLog2e);
    if (tq == 0) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
            out[1056 + mt * 16 + g] = accs[mt][0];
            out[1056 + mt * 16 + g + 8] = accs[mt][2];
        }
    }
}

template <int HEADS, int NSTAGE>
__global__ void __launch_bounds__(256, 2)
q_readout_mma_kernel(const bf16* __restrict__ Q, int64_t ldq, const float* __restrict__ ctx, bf16* __restrict__ O,
                     int64_t ldo, int64_t N, int tiles_per_cta) {
    constexpr int C = HEADS * 32, CPR = C / 8, LDS = C + 8, WPH = 8 / HEADS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    bf16* sQ = reinterpret_cast<bf16*>(smem_raw);                 // [NSTAGE][kTT][LDS]

    const int b = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int hd = warp % HEADS, sub = warp / HEADS;
    const int g = lane >> 2, tq = lane & 3, mi = lane >> 3, lr = lane & 7;
    const bf16* Qb = Q + (int64_t)b * N * ldq;
    bf16* Ob = O + (int64_t)b * N * ldo;

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

    const int64_t tile0 = (int64_t)blockIdx.x * tiles_per_cta;
    int64_t ntiles = ceil_div64(N, kTT) - tile0;
    if (ntiles > tiles_per_cta) ntiles = tiles_per_cta;
    if (ntiles < 0) ntiles = 0;

    // a thread stages / copies out the same chunks every tile: rows r0 + i*RPP, 16-byte column cc
    constexpr int RPP = 256 / CPR, NCH = kTT / RPP;
    const int cc = tid % CPR, r0 = tid / CPR;
    const int64_t goff0 = (int64_t)r0 * ldq + cc * 8, gstep = (int64_t)RPP * ldq;
    const int64_t ooff0 = (int64_t)r0 * ldo + cc * 8, ostep = (int64_t)RPP * ldo;
    const int soff0 = r0 * LDS + cc * 8;
    auto stage = [&](int nb, int64_t trow0) {
        const bf16* gq = Qb + trow0 * ldq + goff0;
        bf16* dq = sQ + nb * kTT * LDS + soff0;
        const int64_t left = N - trow0 - r0;
#pragma unroll
        for (int i = 0; i < NCH; ++i) {
            const bool ok = (int64_t)i * RPP < left;
            cp_async16_zfill(dq + i * RPP * LDS, ok ? gq + i * gstep : Qb, ok ? 16 : 0);
        }
    };
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) {
        if (s < ntiles) stage(s, (tile0 + s) * kTT);
        cp_async_commit();
    }
    for (int64_t t = 0; t < ntiles; ++t) {
        const int buf = (int)(t % NSTAGE);
        if (t + NSTAGE - 1 < ntiles) stage((int)((t + NSTAGE - 1) % NSTAGE), (tile0 + t + NSTAGE - 1) * kTT);
        cp_async_commit();
        cp_async_wait<NSTAGE - 1>();
        __syncthreads();
        bf16* tQ = sQ + buf * kTT * LDS;
        const int64_t row0 = (tile0 + t) * kTT;
        const int valid = (int)((N - row0) < kTT ? (N - row0) : kTT);

#pragma unroll
        for (int mt = 0; mt < kTT / 16; ++mt) {
            if ((mt % WPH) != sub) continue;
            uint32_t a[2][4];
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)   // A: matrices (rows lo, k lo), (rows hi, k lo), (rows lo, k hi), (rows hi, k hi)
                ldsm_x4(smem_u32_generic(tQ + (mt * 16 + lr + 8 * (mi & 1)) * LDS + hd * 32 + ks * 16 + 8 * (mi >> 1)), a[ks]);
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
            const float c0 = m0 * kLog2e, c1 = m1 * kLog2e;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                x0[i] = ex2f(fmaf(x0[i], kLog2e, -c0)); s0 += x0[i];
                x1[i] = ex2f(fmaf(x1[i], kLog2e, -c1)); s1 += x1[i];
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
                for (int ks = 0; ks < 2; ++ks) mma16816(d[nt], a[ks], bf[ks][nt][0], bf[ks][nt][1]);
            }
            const float i0 = 1.f / s0, i1 = 1.f / s1;
            __syncwarp();
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {                      // back in place: this warp owns these rows x head columns
                bf16* p0 = tQ + (mt * 16 + g) * LDS + hd * 32 + nt * 8 + 2 * tq;
                *reinterpret_cast<uint32_t*>(p0) = pack_bf16x2(d[nt][0] * i0, d[nt][1] * i0);
                *reinterpret_cast<uint32_t*>(p0 + 8 * LDS) = pack_bf16x2(d[nt][2] * i1, d[nt][3] * i1);
            }
        }
        __syncthreads();                                          // output tile complete in smem
        {
            bf16* go = Ob + row0 * ldo + ooff0;
            const bf16* so = tQ + soff0;
#pragma unroll
            for (int i = 0; i < NCH; ++i)
                if (r0 + i * RPP < valid)
                    *reinterpret_cast<uint4*>(go + i * ostep) = *reinterpret_cast<const uint4*>(so + i * RPP * LDS);
        }
        __syncthreads();                                          // before the buffer is refilled
    }
    cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------
int kv_chunks_per_batch_host(int B, int64_t N);   // attn_kernels.cu (same split as the fp32 path)
int kv_combine_launch(const float* ws, float* ctx, int heads, int B, int nparts, cudaStream_t st, const void* wo = nullptr,
                      void* wout = nullptr);

template <int HEADS>
static int kv_mma_launch(const bf16* k, const bf16* v, int64_t ld, float* ws, int B, int64_t N, int chunks,
                         int tiles_per_chunk, cudaStream_t st) {
    constexpr int C = HEADS * 32, LDS = C + 8, RPP = 256 / (C / 8);
    constexpr int NSTAGE = HEADS == 8 ? 3 : 4;                    // ~100 KB / ~70 KB / ~37 KB of tiles per CTA
    const size_t smem = (size_t)2 * NSTAGE * kTT * LDS * 2 + (size_t)(RPP * C + 2 * C) * 4;
    static thread_local int conf = -1;
    int dev; cudaGetDevice(&dev);
    if (conf != dev) { cudaFuncSetAttribute(kv_reduce_mma_kernel<HEADS, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024); conf = dev; }
    kv_reduce_mma_kernel<HEADS, NSTAGE><<<dim3(chunks, B), 256, smem, st>>>(k, v, ld, ws, N, chunks, tiles_per_chunk);
    LTU_LAUNCH_CHECK("kv_reduce_mma");
    return LTU_OK;
}

template <int HEADS>
static int q_mma_launch(const bf16* q, int64_t ldq, const float* ctx, bf16* out, int64_t ldo, int B, int64_t N,
                        cudaStream_t st) {
    constexpr int C = HEADS * 32, LDS = C + 8;
    constexpr int NSTAGE = 4;
    const int64_t tiles = ceil_div64(N, kTT);
    int64_t want = ceil_div64(4 * (int64_t)sm_count(), B);
    int64_t ctas = tiles < want ? tiles : want;
    if (ctas < 1) ctas = 1;
    const int tiles_per_cta = (int)ceil_div64(tiles, ctas);
    ctas = ceil_div64(tiles, tiles_per_cta);
    const size_t smem = (size_t)NSTAGE * kTT * LDS * 2;
    static thread_local int conf = -1;
    int dev; cudaGetDevice(&dev);
    if (conf != dev) { cudaFuncSetAttribute(q_readout_mma_kernel<HEADS, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024); conf = dev; }
    q_readout_mma_kernel<HEADS, NSTAGE><<<dim3((unsigned)ctas, B), 256, smem, st>>>(q, ldq, ctx, out, ldo, N, tiles_per_cta);
    LTU_LAUNCH_CHECK("q_readout_mma");
    count_launch(1);
    return LTU_OK;
}

// entry points used by attn_kernels.cu for dtype == bf16 and heads in {2,4,8}
int kv_reduce_bf16_mma(const void* k, const void* v, int64_t ld, float* ctx, void* ws, int B, int64_t N, int heads,
                       cudaStream_t st) {
    const int chunks = kv_chunks_per_batch_host(B, N);
    const int tiles_per_chunk = (int)ceil_div64(ceil_div64(N, kTT), chunks);
    int rc;
    if (heads == 8) rc = kv_mma_launch<8>((const bf16*)k, (const bf16*)v, ld, (float*)ws, B, N, chunks, tiles_per_chunk, st);
    else if (heads == 4) rc = kv_mma_launch<4>((const bf16*)k, (const bf16*)v, ld, (float*)ws, B, N, chunks, tiles_per_chunk, st);
    else rc = kv_mma_launch<2>((const bf16*)k, (const bf16*)v, ld, (float*)ws, B, N, chunks, tiles_per_chunk, st);
    if (rc != LTU_OK) return rc;
    rc = kv_combine_launch((const float*)ws, ctx, heads, B, chunks * (8 / heads), st);
    if (rc != LTU_OK) return rc;
    count_launch(2);
    return LTU_OK;
}

int q_readout_bf16_mma(const void* q, int64_t ldq, const float* ctx, void* out, int64_t ldo, int B, int64_t N,
                       int heads, cudaStream_t st) {
    if (heads == 8) return q_mma_launch<8>((const bf16*)q, ldq, ctx, (bf16*)out, ldo, B, N, st);
    if (heads == 4) return q_mma_launch<4>((const bf16*)q, ldq, ctx, (bf16*)out, ldo, B, N, st);
    return q_mma_launch<2>((const bf16*)q, ldq, ctx, (bf16*)out, ldo, B, N, st);
}

}  // namespace ltu
