// Backward of the linear-attention core, bf16 path on the tensor pipe (mma.sync m16n8k16, fp32 accumulate) -- the same
// mathematics and the same three-launch structure as attn_bwd.cu (reduce -> combine -> apply), with the 32x32 per-head
// contractions as MMAs so that the passes are bound by HBM, not by instruction issue (the CUDA-core version spends
// 150 / 300 warp instructions per token-head; here ~5 / ~16).  tcgen05 is not used for the same reason as in attn_tc.cu:
// the operands are 32x32 per-head states, an M=128 UMMA tile would be mostly padding.
//
// pass 1 (attn_bwd_reduce_mma): a warp owns (head, token phase); per 32-token tile one LANE owns one token row:
//   it loads the row's 32 q values (64 B), takes the row softmax entirely in registers (no shuffles), writes
//   Qs = P/sqrt(32) as bf16 into the warp's private shared-memory tile and copies the row of G next to it;
//   dctx[j][e] += Qs^T G by ldmatrix.trans + mma (tokens are the K dimension), exactly the fragment pattern of
//   kv_reduce_mma.  The K tile then reuses the Qs buffer and lane j folds column j into the running (max, sum).
// pass 2 (attn_bwd_apply_mma): per 16-token tile the A fragments of G, V, K, Q come straight from global memory
//   (32-bit loads in the mma A layout; a C fragment has the same (row, column) ownership, so P, Ks and the three
//   products combine element-wise in registers); ctx / dctx / dctx^T live in registers as bf16 B fragments.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);
int kv_chunks_per_batch_host(int B, int64_t N);

namespace {

constexpr int kPartF = 32 * 32 + 64;
constexpr float kRsqrtD = 0.17677669529663687f;
constexpr int kLds = 40;                                     // padded bf16 row of a 32-column tile (80 B: conflict-free)

__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float lo_f(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float hi_f(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void unpack8w(const uint4& v, float* f) {
    f[0] = lo_f(v.x); f[1] = hi_f(v.x); f[2] = lo_f(v.y); f[3] = hi_f(v.y);
    f[4] = lo_f(v.z); f[5] = hi_f(v.z); f[6] = lo_f(v.w); f[7] = hi_f(v.w);
}

}  // namespace

// ---------------------------------------------------------------- pass 1
// grid (chunks, B), 256 threads.  Partials in the layout of attn_bwd.cu (same combine kernel).
__global__ void __launch_bounds__(256)
attn_bwd_reduce_mma_kernel(const bf16* __restrict__ Q, int64_t ldq, const bf16* __restrict__ K, int64_t ldk,
                           const bf16* __restrict__ G, int64_t ldg, float* __restrict__ part, int64_t N, int heads,
                           int chunks, int64_t tokens_per_chunk) {
    __shared__ __align__(16) bf16 sA[8][32 * kLds];          // Qs tile, then the K tile
    __shared__ __align__(16) bf16 sB[8][32 * kLds];          // G tile
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wph = 8 / heads, hd = warp % heads, sub = warp / heads;
    const int64_t n0 = (int64_t)chunk * tokens_per_chunk;
    int64_t n1 = n0 + tokens_per_chunk;
    if (n1 > N) n1 = N;
    const bf16* q = Q + (int64_t)b * N * ldq + hd * 32;
    const bf16* k = K + (int64_t)b * N * ldk + hd * 32;
    const bf16* g = G + (int64_t)b * N * ldg + hd * 32;
    bf16* tA = sA[warp];
    bf16* tB = sB[warp];

    float acc[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;
    float m_run = -INFINITY, s_run = 0.f;
    const int mi = lane >> 3, lr = lane & 7;

    for (int64_t t0 = n0 + (int64_t)sub * 32; t0 < n1; t0 += (int64_t)wph * 32) {
        const int64_t n = t0 + lane;                         // this lane's token row
        const bool ok = n < n1;
        // ---- Qs row (softmax in registers) and G row into the tiles
        uint4 qw[4], gw[4], kw[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            qw[c] = ok ? *reinterpret_cast<const uint4*>(q + n * ldq + c * 8) : make_uint4(0, 0, 0, 0);
            gw[c] = ok ? *reinterpret_cast<const uint4*>(g + n * ldg + c * 8) : make_uint4(0, 0, 0, 0);
            kw[c] = ok ? *reinterpret_cast<const uint4*>(k + n * ldk + c * 8) : make_uint4(0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u, 0xFF80FF80u);
        }
        float v[32];
#pragma unroll
        for (int c = 0; c < 4; ++c) unpack8w(qw[c], v + c * 8);
        float mx = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) { v[i] = __expf(v[i] - mx); sum += v[i]; }
        const float sc = ok ? kRsqrtD / sum : 0.f;           // rows past the end contribute nothing
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            uint4 o;
            o.x = pack_bf16x2(v[c * 8] * sc, v[c * 8 + 1] * sc);     o.y = pack_bf16x2(v[c * 8 + 2] * sc, v[c * 8 + 3] * sc);
            o.z = pack_bf16x2(v[c * 8 + 4] * sc, v[c * 8 + 5] * sc); o.w = pack_bf16x2(v[c * 8 + 6] * sc, v[c * 8 + 7] * sc);
            *reinterpret_cast<uint4*>(tA + lane * kLds + c * 8) = o;
            *reinterpret_cast<uint4*>(tB + lane * kLds + c * 8) = gw[c];
        }
        __syncwarp();
        // ---- dctx[j][e] += sum_n Qs[n][j] G[n][e]   (M = j, N = e, K = tokens)
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int r = ks * 16;
            uint32_t a[2][4], bq[2][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)     // A^T: matrices (k lo, m lo), (k lo, m hi), (k hi, m lo), (k hi, m hi)
                ldsm4t(smem_u32_generic(tA + (r + lr + 8 * (mi >> 1)) * kLds + mt * 16 + 8 * (mi & 1)), a[mt]);
#pragma unroll
            for (int np = 0; np < 2; ++np)     // B: matrices (k lo, n lo), (k hi, n lo), (k lo, n hi), (k hi, n hi)
                ldsm4t(smem_u32_generic(tB + (r + lr + 8 * (mi & 1)) * kLds + np * 16 + 8 * (mi >> 1)), bq[np]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt)
                    mma_bf16(acc[mt][nt], a[mt], bq[nt >> 1][(nt & 1) * 2], bq[nt >> 1][(nt & 1) * 2 + 1]);
        }
        __syncwarp();
        // ---- column statistics of K: the K tile reuses the Qs buffer, lane j walks down column j
#pragma unroll
        for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(tA + lane * kLds + c * 8) = kw[c];
        __syncwarp();
        float kc[32];
        float tmax = -INFINITY;
#pragma unroll
        for (int r = 0; r < 32; ++r) { kc[r] = __bfloat162float(tA[r * kLds + lane]); tmax = fmaxf(tmax, kc[r]); }
        float ts = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) ts += __expf(kc[r] - tmax);     // padded rows are -inf: exp = 0
        if (tmax > m_run) { s_run = s_run * __expf(m_run - tmax) + ts; m_run = tmax; }
        else s_run += ts * __expf(tmax - m_run);
        __syncwarp();                                        // before the next tile overwrites the buffers
    }
    float* out = part + ((((int64_t)b * chunks + chunk) * wph + sub) * heads + hd) * kPartF;
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            const int j = mt * 16 + gq, e = nt * 8 + 2 * tq;
            *reinterpret_cast<float2*>(out + j * 32 + e) = make_float2(acc[mt][nt][0], acc[mt][nt][1]);
            *reinterpret_cast<float2*>(out + (j + 8) * 32 + e) = make_float2(acc[mt][nt][2], acc[mt][nt][3]);
        }
    out[1024 + lane] = m_run;
    out[1056 + lane] = s_run;
}

// ---------------------------------------------------------------- pass 2
// grid (ctas, B), 256 threads; every CTA owns a contiguous token range; warp: head w % heads, tile phase w / heads.
// Ownership in every fragment below: rows g = lane/4 (r0) and g + 8 (r1) of the 16-token tile, columns
// 8 n' + 2 t + {0,1}, n' = 0..3, t = lane%4.  A-fragment register (s, x): row r(x&1), n' = 2 s + (x >> 1).
__global__ void __launch_bounds__(256)
attn_bwd_apply_mma_kernel(const bf16* __restrict__ Q, int64_t ldq, const bf16* __restrict__ K, const bf16* __restrict__ V,
                          int64_t ldkv, const bf16* __restrict__ G, int64_t ldg, const float* __restrict__ ctx,
                          const float* __restrict__ dctx, const float* __restrict__ kst, bf16* __restrict__ dQ,
                          bf16* __restrict__ dK, bf16* __restrict__ dV, int64_t ldd, int64_t N, int heads,
                          int64_t tokens_per_cta) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wph = 8 / heads, hd = warp % heads, sub = warp / heads;
    const int gq = lane >> 2, tq = lane & 3;
    const int64_t n0 = (int64_t)blockIdx.x * tokens_per_cta;
    int64_t n1 = n0 + tokens_per_cta;
    if (n1 > N) n1 = N;
    const float* cb = ctx + ((int64_t)b * heads + hd) * 1024;
    const float* db = dctx + ((int64_t)b * heads + hd) * 1024;
    const float* ks = kst + ((int64_t)b * heads + hd) * 96;

    // B fragments (n-tile n', k-step s): b0b1 = B[k = 16s+2t, +1][n = 8n'+g], b2b3 = B[k = 16s+2t+8, +9][n = 8n'+g]
    uint32_t bc[4][2][2], bd[4][2][2], bt[4][2][2];
#pragma unroll
    for (int np = 0; np < 4; ++np)
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int nn = 8 * np + gq, kk = 16 * s + 2 * tq + 8 * h2;
                bc[np][s][h2] = pack_bf16x2(cb[nn * 32 + kk], cb[nn * 32 + kk + 1]);         // dQs = G ctx^T : B[k=e][n=j] = ctx[j][e]
                bd[np][s][h2] = pack_bf16x2(db[nn * 32 + kk], db[nn * 32 + kk + 1]);         // dKs = V dctx^T: B[k=e][n=j] = dctx[j][e]
                bt[np][s][h2] = pack_bf16x2(db[kk * 32 + nn], db[(kk + 1) * 32 + nn]);       // dV  = Ks dctx  : B[k=j][n=e] = dctx[j][e]
            }
    float cM[4][2], cI[4][2], cT[4][2];                      // column constants of K: max, 1/sum, t
#pragma unroll
    for (int np = 0; np < 4; ++np)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int c = 8 * np + 2 * tq + i;
            cM[np][i] = ks[c];
            cI[np][i] = 1.f / ks[32 + c];
            cT[np][i] = ks[64 + c];
        }
    const int64_t base_q = (int64_t)b * N * ldq + hd * 32 + 2 * tq;
    const int64_t base_kv = (int64_t)b * N * ldkv + hd * 32 + 2 * tq;
    const int64_t base_g = (int64_t)b * N * ldg + hd * 32 + 2 * tq;
    const int64_t base_d = (int64_t)b * N * ldd + hd * 32 + 2 * tq;

    // A fragments straight from global memory: register (s, x) = row x&1, columns 16 s + 8 (x>>1) + 2t, +1.
    // The NEXT tile's fragments are requested before the current tile is processed (one tile of loads always in flight).
    uint32_t nq[2][4], nk[2][4], nv[2][4], ng[2][4];
    auto fetch = [&](int64_t t0) {
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int64_t r = t0 + gq + 8 * (x & 1);
                const int co = 16 * s + 8 * (x >> 1);
                const bool o = r < n1;
                nq[s][x] = o ? *reinterpret_cast<const uint32_t*>(Q + base_q + r * ldq + co) : 0u;
                nk[s][x] = o ? *reinterpret_cast<const uint32_t*>(K + base_kv + r * ldkv + co) : 0u;
                nv[s][x] = o ? *reinterpret_cast<const uint32_t*>(V + base_kv + r * ldkv + co) : 0u;
                ng[s][x] = o ? *reinterpret_cast<const uint32_t*>(G + base_g + r * ldg + co) : 0u;
            }
    };
    const int64_t first = n0 + (int64_t)sub * 16;
    if (first < n1) fetch(first);
    for (int64_t t0 = first; t0 < n1; t0 += (int64_t)wph * 16) {
        const int64_t row[2] = {t0 + gq, t0 + gq + 8};
        const bool ok[2] = {row[0] < n1, row[1] < n1};
        uint32_t aq[2][4], ak[2][4], av[2][4], ag[2][4];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int x = 0; x < 4; ++x) { aq[s][x] = nq[s][x]; ak[s][x] = nk[s][x]; av[s][x] = nv[s][x]; ag[s][x] = ng[s][x]; }
        if (t0 + (int64_t)wph * 16 < n1) fetch(t0 + (int64_t)wph * 16);
        // ---- P = softmax_d(Q) per row (quad reduction), Ks = exp(K - M)/S; element [row][n'][i]
        float p[2][4][2], kf[2][4][2];
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int rr = x & 1, np = 2 * s + (x >> 1);
                p[rr][np][0] = lo_f(aq[s][x]); p[rr][np][1] = hi_f(aq[s][x]);
                mx[rr] = fmaxf(mx[rr], fmaxf(p[rr][np][0], p[rr][np][1]));
                kf[rr][np][0] = __expf(lo_f(ak[s][x]) - cM[np][0]) * cI[np][0];
                kf[rr][np][1] = __expf(hi_f(ak[s][x]) - cM[np][1]) * cI[np][1];
            }
        float sm[2] = {0.f, 0.f};
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 1));
            mx[rr] = fmaxf(mx[rr], __shfl_xor_sync(0xffffffffu, mx[rr], 2));
#pragma unroll
            for (int np = 0; np < 4; ++np)
#pragma unroll
                for (int i = 0; i < 2; ++i) { p[rr][np][i] = __expf(p[rr][np][i] - mx[rr]); sm[rr] += p[rr][np][i]; }
            sm[rr] += __shfl_xor_sync(0xffffffffu, sm[rr], 1);
            sm[rr] += __shfl_xor_sync(0xffffffffu, sm[rr], 2);
            const float inv = 1.f / sm[rr];
#pragma unroll
            for (int np = 0; np < 4; ++np) { p[rr][np][0] *= inv; p[rr][np][1] *= inv; }
        }
        uint32_t aks[2][4];                                  // Ks as an A operand
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                const int rr = x & 1, np = 2 * s + (x >> 1);
                aks[s][x] = pack_bf16x2(kf[rr][np][0], kf[rr][np][1]);
            }
        // ---- three products, 16 tokens x 32 columns each: acc[n'] = (row r0: cols 8n'+2t,+1 | row r1: same)
        float dqs[4][4], dks[4][4], dvv[4][4];
#pragma unroll
        for (int np = 0; np < 4; ++np) {
#pragma unroll
            for (int i = 0; i < 4; ++i) { dqs[np][i] = 0.f; dks[np][i] = 0.f; dvv[np][i] = 0.f; }
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                mma_bf16(dqs[np], ag[s], bc[np][s][0], bc[np][s][1]);
                mma_bf16(dks[np], av[s], bd[np][s][0], bd[np][s][1]);
                mma_bf16(dvv[np], aks[s], bt[np][s][0], bt[np][s][1]);
            }
        }
        // ---- epilogues: dQ = P (dP - sum_j P dP), dK = Ks (dKs - t), dV
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            float dot = 0.f;
#pragma unroll
            for (int np = 0; np < 4; ++np)
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    dqs[np][2 * rr + i] *= kRsqrtD;          // dP
                    dot = fmaf(p[rr][np][i], dqs[np][2 * rr + i], dot);
                }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            if (ok[rr]) {
#pragma unroll
                for (int np = 0; np < 4; ++np) {
                    const int64_t o = base_d + row[rr] * ldd + 8 * np;
                    *reinterpret_cast<uint32_t*>(dQ + o) = pack_bf16x2(p[rr][np][0] * (dqs[np][2 * rr] - dot),
                                                                       p[rr][np][1] * (dqs[np][2 * rr + 1] - dot));
                    *reinterpret_cast<uint32_t*>(dK + o) = pack_bf16x2(kf[rr][np][0] * (dks[np][2 * rr] - cT[np][0]),
                                                                       kf[rr][np][1] * (dks[np][2 * rr + 1] - cT[np][1]));
                    *reinterpret_cast<uint32_t*>(dV + o) = pack_bf16x2(dvv[np][2 * rr], dvv[np][2 * rr + 1]);
                }
            }
        }
    }
}

// fixed-order merge shared with the CUDA-core path (attn_bwd.cu)
int attn_bwd_combine_launch(const float* ws, const float* ctx, float* dctx, float* kst, int heads, int B, int nparts,
                            cudaStream_t st);

int attn_bwd_bf16_mma(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* g, int64_t ldg,
                      const float* ctx, void* dq, void* dk, void* dv, int64_t ldd, float* dctx, float* kst, float* ws,
                      int B, int64_t N, int heads, cudaStream_t st) {
    const int chunks = kv_chunks_per_batch_host(B, N);
    const int wph = 8 / heads;
    // chunk boundaries on multiples of 32 tokens so that a tile never straddles two chunks
    const int64_t tokens_per_chunk = ceil_div64(ceil_div64(N, 32), chunks) * 32;
    attn_bwd_reduce_mma_kernel<<<dim3(chunks, B), 256, 0, st>>>((const bf16*)q, ldq, (const bf16*)k, ldkv, (const bf16*)g, ldg,
                                                               ws, N, heads, chunks, tokens_per_chunk);
    LTU_LAUNCH_CHECK("attn_bwd_reduce_mma");
    int rc = attn_bwd_combine_launch(ws, ctx, dctx, kst, heads, B, chunks * wph, st);
    if (rc != LTU_OK) return rc;
    int64_t ctas = ceil_div64(8 * (int64_t)sm_count(), B);
    const int64_t max_ctas = ceil_div64(N, 16 * wph);        // at least one 16-token tile per warp
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    int64_t tokens_per_cta = ceil_div64(ceil_div64(N, ctas), 16) * 16;
    ctas = ceil_div64(N, tokens_per_cta);
    attn_bwd_apply_mma_kernel<<<dim3((unsigned)ctas, B), 256, 0, st>>>((const bf16*)q, ldq, (const bf16*)k, (const bf16*)v, ldkv,
                                                                      (const bf16*)g, ldg, ctx, dctx, kst, (bf16*)dq, (bf16*)dk,
                                                                      (bf16*)dv, ldd, N, heads, tokens_per_cta);
    LTU_LAUNCH_CHECK("attn_bwd_apply_mma");
    count_launch(3);
    return LTU_OK;
}

}  // namespace ltu
