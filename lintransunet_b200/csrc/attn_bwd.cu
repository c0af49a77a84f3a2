// Backward of the linear-attention core (first slice): gradients of
//   out = softmax_d(Q)/sqrt(32) . (softmax_N(K)^T V)          model/trans_block.py:41-67
// with respect to Q, K and V, per head (head_dim 32), on the same strided [B, N, C] views the forward uses.
//
// With P = softmax_d(Q), Qs = P/sqrt(32), Ks = softmax_N(K), ctx = Ks^T V (from the forward) and G = dOut:
//   dctx = Qs^T G                         [32x32]  reduction over the N tokens           (pass 1)
//   t_j  = sum_e dctx[j][e] ctx[j][e]     = sum_n Ks[n][j] dKs[n][j]: the column-softmax correction needs NO second
//                                           pass over the tokens
//   dV   = Ks dctx ;  dKs = V dctx^T ;  dK = Ks (dKs - t)
//   dQs  = G ctx^T ;  dP = dQs/sqrt(32) ;  dQ = P (dP - sum_j P_j dP_j)                   (pass 2)
// The column statistics of K (max m_j, sum s_j) are rebuilt in pass 1 from K itself (online max), so the op
// needs nothing saved from the forward beyond ctx.
//
// First version: fp32 arithmetic on CUDA cores, one warp per (token, head) with lane = column, the 32x32
// contractions as FMAs against shared-memory broadcasts (16-byte broadcast loads; the shuffle-broadcast variant ran
// at 637 GB/s, bound by the shuffle pipe); fixed-order merges => bit-reproducible.  The tensor-pipe (mma) form
// of the forward kernels is the next step.  Bytes: pass 1 reads Q, G, K,
// pass 2 reads Q, K, V, G and writes dQ, dK, dV: 10 N C E against a floor of 7 N C E.
#include <stdlib.h>

#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);
int kv_chunks_per_batch_host(int B, int64_t N);              // attn_kernels.cu: depends on N only

constexpr int kBwdPartF = 32 * 32 + 64;                      // dctx[32][32], m[32], s[32]
constexpr float kInvSqrtD = 0.17677669529663687f;            // 1/sqrt(32)
constexpr int kPD = 4;                                       // tokens in flight per warp (register prefetch ring)

template <typename T> __device__ __forceinline__ float ld1(const T* p) { return to_f32(*p); }

// ---------------------------------------------------------------- pass 1: dctx partials + K column statistics
// grid (chunks, B), 256 threads; warp w: head w % heads, token phase w / heads.
// partial index ((b*chunks + chunk)*wph + sub)*heads + hd, like the forward's kv_reduce.
template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_reduce_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ K, int64_t ldk,
                       const T* __restrict__ G, int64_t ldg, float* __restrict__ part, int64_t N, int heads,
                       int chunks, int64_t tokens_per_chunk) {
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wph = 8 / heads, hd = warp % heads, sub = warp / heads;
    const int64_t n0 = (int64_t)chunk * tokens_per_chunk;
    int64_t n1 = n0 + tokens_per_chunk;
    if (n1 > N) n1 = N;
    const int col = hd * 32 + lane;
    const T* q = Q + (int64_t)b * N * ldq + col;
    const T* k = K + (int64_t)b * N * ldk + col;
    const T* g = G + (int64_t)b * N * ldg + col;

    // The 32-wide broadcasts go through shared memory (one 16-byte broadcast load feeds 4 FMAs); with shuffles the
    // kernel was bound by the shuffle pipe (one warp shuffle per clock per SM).
    __shared__ __align__(16) float sp[8][32];
    float* myp = sp[warp];
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    float m_run = -INFINITY, s_run = 0.f;
    // kPD tokens of this warp are always in flight: slot u is reloaded (kPD tokens ahead) right after it is consumed,
    // so the global-load latency of a token overlaps the arithmetic of the kPD - 1 tokens before it.
    float rq[kPD], rk[kPD], rg[kPD];
#pragma unroll
    for (int u = 0; u < kPD; ++u) {
        const int64_t n = n0 + sub + (int64_t)u * wph;
        const bool ok = n < n1;
        rq[u] = ok ? ld1(q + n * ldq) : 0.f;
        rk[u] = ok ? ld1(k + n * ldk) : 0.f;
        rg[u] = ok ? ld1(g + n * ldg) : 0.f;
    }
    for (int64_t nb = n0 + sub; nb < n1; nb += (int64_t)kPD * wph) {
#pragma unroll
        for (int u = 0; u < kPD; ++u) {
            const int64_t n = nb + (int64_t)u * wph;
            if (n >= n1) break;                              // warp-uniform
            const float qv = rq[u], kv = rk[u], gv = rg[u];
            const int64_t nn = n + (int64_t)kPD * wph;
            if (nn < n1) { rq[u] = ld1(q + nn * ldq); rk[u] = ld1(k + nn * ldk); rg[u] = ld1(g + nn * ldg); }
            const float ex = __expf(qv - warp_max(qv));
            const float p = ex / warp_sum(ex) * kInvSqrtD;   // Qs[n][lane]
            __syncwarp();                                    // the previous token's reads are done
            myp[lane] = p;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 pj = *reinterpret_cast<const float4*>(myp + j);
                acc[j] = fmaf(pj.x, gv, acc[j]);
                acc[j + 1] = fmaf(pj.y, gv, acc[j + 1]);
                acc[j + 2] = fmaf(pj.z, gv, acc[j + 2]);
                acc[j + 3] = fmaf(pj.w, gv, acc[j + 3]);
            }
            if (kv > m_run) { s_run = s_run * __expf(m_run - kv) + 1.f; m_run = kv; }
            else s_run += __expf(kv - m_run);
        }
    }
    float* out = part + ((((int64_t)b * chunks + chunk) * wph + sub) * heads + hd) * kBwdPartF;
#pragma unroll
    for (int j = 0; j < 32; ++j) out[j * 32 + lane] = acc[j];
    out[1024 + lane] = m_run;
    out[1056 + lane] = s_run;
}

// grid (heads, B), 1024 threads: warp j, lane e own dctx[j][e].  Fixed-order merge of the partials.
// kst [B, heads, 3, 32] = (m_j, s_j, t_j).
__global__ void __launch_bounds__(1024)
attn_bwd_combine_kernel(const float* __restrict__ part, const float* __restrict__ ctx, float* __restrict__ dctx,
                        float* __restrict__ kst, int heads, int nparts) {
    const int hd = blockIdx.x, b = blockIdx.y;
    const int j = threadIdx.x >> 5, e = threadIdx.x & 31;
    const float* base = part + ((int64_t)b * nparts * heads + hd) * kBwdPartF;
    const int64_t stride = (int64_t)heads * kBwdPartF;
    float a = 0.f;
#pragma unroll 8
    for (int p = 0; p < nparts; ++p) a += base[p * stride + j * 32 + e];     // fixed order; 8 loads in flight
    const int64_t o = (((int64_t)b * heads + hd) * 32 + j) * 32 + e;
    dctx[o] = a;
    const float t = warp_sum(a * ctx[o]);
    // column statistics of K for column j: the lanes split the parts (exact max; per-lane sums in part order, then a
    // butterfly: the same order every run)
    float M = -INFINITY;
    for (int p = e; p < nparts; p += 32) M = fmaxf(M, base[p * stride + 1024 + j]);
    M = warp_max(M);
    float S = 0.f;
    for (int p = e; p < nparts; p += 32) {
        const float mp = base[p * stride + 1024 + j];
        if (mp != -INFINITY) S += base[p * stride + 1056 + j] * __expf(mp - M);
    }
    S = warp_sum(S);
    if (e == 0) {
        float* ks = kst + ((int64_t)b * heads + hd) * 96;
        ks[j] = M;
        ks[32 + j] = S;
        ks[64 + j] = t;
    }
}

// ---------------------------------------------------------------- pass 2: dQ, dK, dV
// grid (ctas, B), 256 threads; every CTA owns a contiguous token range, warp w: head w % heads, phase w / heads.
template <typename T>
__global__ void __launch_bounds__(256, 2)
attn_bwd_apply_kernel(const T* __restrict__ Q, int64_t ldq, const T* __restrict__ K, const T* __restrict__ V, int64_t ldkv,
                      const T* __restrict__ G, int64_t ldg, const float* __restrict__ ctx, const float* __restrict__ dctx,
                      const float* __restrict__ kst, T* __restrict__ dQ, T* __restrict__ dK, T* __restrict__ dV,
                      int64_t ldd, int64_t N, int heads, int64_t tokens_per_cta) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wph = 8 / heads, hd = warp % heads, sub = warp / heads;
    const int64_t n0 = (int64_t)blockIdx.x * tokens_per_cta;
    int64_t n1 = n0 + tokens_per_cta;
    if (n1 > N) n1 = N;
    const int col = hd * 32 + lane;
    const float* cb = ctx + ((int64_t)b * heads + hd) * 1024;
    const float* db = dctx + ((int64_t)b * heads + hd) * 1024;
    const float* ks = kst + ((int64_t)b * heads + hd) * 96;
    float cr[32], dr[32], dc[32];                            // ctx row `lane`, dctx row `lane`, dctx column `lane`
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        cr[i] = cb[lane * 32 + i];
        dr[i] = db[lane * 32 + i];
        dc[i] = db[i * 32 + lane];
    }
    const float M = ks[lane], invS = 1.f / ks[32 + lane], t = ks[64 + lane];
    __shared__ __align__(16) float sb[8][3][32];             // per-warp broadcast staging of G, V, Ks of one token
    float* sg = sb[warp][0];
    float* sv = sb[warp][1];
    float* sk = sb[warp][2];
    const T* q = Q + (int64_t)b * N * ldq + col;
    const T* k = K + (int64_t)b * N * ldkv + col;
    const T* v = V + (int64_t)b * N * ldkv + col;
    const T* g = G + (int64_t)b * N * ldg + col;
    T* dq = dQ + (int64_t)b * N * ldd + col;
    T* dk = dK + (int64_t)b * N * ldd + col;
    T* dv = dV + (int64_t)b * N * ldd + col;
    float rq[kPD], rk[kPD], rv[kPD], rg[kPD];                // kPD tokens in flight (see attn_bwd_reduce_kernel)
#pragma unroll
    for (int u = 0; u < kPD; ++u) {
        const int64_t n = n0 + sub + (int64_t)u * wph;
        const bool ok = n < n1;
        rq[u] = ok ? ld1(q + n * ldq) : 0.f;
        rk[u] = ok ? ld1(k + n * ldkv) : 0.f;
        rv[u] = ok ? ld1(v + n * ldkv) : 0.f;
        rg[u] = ok ? ld1(g + n * ldg) : 0.f;
    }
    for (int64_t nb = n0 + sub; nb < n1; nb += (int64_t)kPD * wph) {
#pragma unroll
        for (int u = 0; u < kPD; ++u) {
            const int64_t n = nb + (int64_t)u * wph;
            if (n >= n1) break;                              // warp-uniform
            const float qv = rq[u], kv = rk[u], vv = rv[u], gv = rg[u];
            const int64_t nn = n + (int64_t)kPD * wph;
            if (nn < n1) {
                rq[u] = ld1(q + nn * ldq); rk[u] = ld1(k + nn * ldkv); rv[u] = ld1(v + nn * ldkv); rg[u] = ld1(g + nn * ldg);
            }
            const float ex = __expf(qv - warp_max(qv));
            const float P = ex / warp_sum(ex);
            const float Ks = __expf(kv - M) * invS;
            __syncwarp();                                    // the previous token's reads are done
            sg[lane] = gv; sv[lane] = vv; sk[lane] = Ks;
            __syncwarp();
            float dqs = 0.f, dks = 0.f, dvv = 0.f;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                const float4 a = *reinterpret_cast<const float4*>(sg + i);   // sum_e G[e]  ctx[lane][e]
                const float4 b4 = *reinterpret_cast<const float4*>(sv + i);  // sum_e V[e]  dctx[lane][e]
                const float4 c = *reinterpret_cast<const float4*>(sk + i);   // sum_j Ks[j] dctx[j][lane]
                dqs = fmaf(a.x, cr[i], dqs); dqs = fmaf(a.y, cr[i + 1], dqs); dqs = fmaf(a.z, cr[i + 2], dqs); dqs = fmaf(a.w, cr[i + 3], dqs);
                dks = fmaf(b4.x, dr[i], dks); dks = fmaf(b4.y, dr[i + 1], dks); dks = fmaf(b4.z, dr[i + 2], dks); dks = fmaf(b4.w, dr[i + 3], dks);
                dvv = fmaf(c.x, dc[i], dvv); dvv = fmaf(c.y, dc[i + 1], dvv); dvv = fmaf(c.z, dc[i + 2], dvv); dvv = fmaf(c.w, dc[i + 3], dvv);
            }
            const float dP = dqs * kInvSqrtD;
            const float dot = warp_sum(P * dP);
            dq[n * ldd] = from_f32<T>(P * (dP - dot));
            dk[n * ldd] = from_f32<T>(Ks * (dks - t));
            dv[n * ldd] = from_f32<T>(dvv);
        }
    }
}

int attn_bwd_combine_launch(const float* ws, const float* ctx, float* dctx, float* kst, int heads, int B, int nparts,
                            cudaStream_t st) {
    attn_bwd_combine_kernel<<<dim3(heads, B), 1024, 0, st>>>(ws, ctx, dctx, kst, heads, nparts);
    LTU_LAUNCH_CHECK("attn_bwd_combine");
    return LTU_OK;
}

// bf16 tensor-pipe variant (attn_bwd_tc.cu)
int attn_bwd_bf16_mma(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* g, int64_t ldg,
                      const float* ctx, void* dq, void* dk, void* dv, int64_t ldd, float* dctx, float* kst, float* ws,
                      int B, int64_t N, int heads, cudaStream_t st);
static bool use_mma_attn_bwd() {
    static const bool v = [] { const char* e = getenv("LTU_ATTN_BWD_FMA"); return !(e && e[0] == '1'); }();
    return v;
}

template <typename T>
static int attn_bwd_impl(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* g, int64_t ldg,
                         const float* ctx, void* dq, void* dk, void* dv, int64_t ldd, float* dctx, float* kst, float* ws,
                         int B, int64_t N, int heads, cudaStream_t st) {
    const int chunks = kv_chunks_per_batch_host(B, N);
    const int64_t tokens_per_chunk = ceil_div64(N, chunks);
    const int wph = 8 / heads;
    attn_bwd_reduce_kernel<T><<<dim3(chunks, B), 256, 0, st>>>((const T*)q, ldq, (const T*)k, ldkv, (const T*)g, ldg, ws, N,
                                                              heads, chunks, tokens_per_chunk);
    LTU_LAUNCH_CHECK("attn_bwd_reduce");
    attn_bwd_combine_kernel<<<dim3(heads, B), 1024, 0, st>>>(ws, ctx, dctx, kst, heads, chunks * wph);
    LTU_LAUNCH_CHECK("attn_bwd_combine");
    int64_t ctas = ceil_div64(4 * (int64_t)sm_count(), B);
    const int64_t max_ctas = ceil_div64(N, 32);              // at least 32 tokens per CTA
    if (ctas > max_ctas) ctas = max_ctas;
    if (ctas < 1) ctas = 1;
    const int64_t tokens_per_cta = ceil_div64(N, ctas);
    ctas = ceil_div64(N, tokens_per_cta);
    attn_bwd_apply_kernel<T><<<dim3((unsigned)ctas, B), 256, 0, st>>>((const T*)q, ldq, (const T*)k, (const T*)v, ldkv,
                                                                     (const T*)g, ldg, ctx, dctx, kst, (T*)dq, (T*)dk, (T*)dv,
                                                                     ldd, N, heads, tokens_per_cta);
    LTU_LAUNCH_CHECK("attn_bwd_apply");
    count_launch(3);
    return LTU_OK;
}

}  // namespace ltu

using namespace ltu;

extern "C" size_t ltu_attn_bwd_workspace(int B, int64_t N, int heads) {
    if (B <= 0 || N <= 0 || !(heads == 1 || heads == 2 || heads == 4 || heads == 8)) return 0;
    return (size_t)B * kv_chunks_per_batch_host(B, N) * (8 / heads) * heads * kBwdPartF * sizeof(float);
}

extern "C" int ltu_attn_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv, const void* dout,
                            int64_t ldo, const float* ctx, void* dq, void* dk, void* dv, int64_t ldd, float* dctx,
                            float* kstats, void* ws, size_t ws_bytes, int B, int64_t N, int heads, int dtype,
                            ltu_stream_t stream) {
    LTU_ARG_CHECK(q && k && v && dout && ctx && dq && dk && dv && dctx && kstats && ws, "attn_bwd: null pointer");
    LTU_ARG_CHECK(B > 0 && N > 0 && B <= 65535, "attn_bwd: bad B=%d N=%lld", B, (long long)N);
    LTU_ARG_CHECK(heads == 1 || heads == 2 || heads == 4 || heads == 8, "attn_bwd: heads must be 1,2,4 or 8 (got %d)", heads);
    LTU_ARG_CHECK(dtype == LTU_F32 || dtype == LTU_BF16, "attn_bwd: bad dtype %d", dtype);
    const int C = heads * 32;
    LTU_ARG_CHECK(ldq >= C && ldkv >= C && ldo >= C && ldd >= C, "attn_bwd: row strides must be >= heads*32");
    LTU_ARG_CHECK(ws_bytes >= ltu_attn_bwd_workspace(B, N, heads), "attn_bwd: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == LTU_F32)
        return attn_bwd_impl<float>(q, ldq, k, v, ldkv, dout, ldo, ctx, dq, dk, dv, ldd, dctx, kstats, (float*)ws, B, N, heads, st);
    const bool vec_ok = ldq % 8 == 0 && ldkv % 8 == 0 && ldo % 8 == 0 && ldd % 2 == 0 &&
                        ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                          reinterpret_cast<uintptr_t>(dout)) & 15) == 0 &&
                        ((reinterpret_cast<uintptr_t>(dq) | reinterpret_cast<uintptr_t>(dk) | reinterpret_cast<uintptr_t>(dv)) & 3) == 0;
    if (vec_ok && use_mma_attn_bwd())
        return attn_bwd_bf16_mma(q, ldq, k, v, ldkv, dout, ldo, ctx, dq, dk, dv, ldd, dctx, kstats, (float*)ws, B, N, heads, st);
    return attn_bwd_impl<bf16>(q, ldq, k, v, ldkv, dout, ldo, ctx, dq, dk, dv, ldd, dctx, kstats, (float*)ws, B, N, heads, st);
}
