// Deep-supervision training loss of the reference: CrossEntroLoss (loss/criterions.py:696-718), DiceClassLoss
// (:35-69) and BalanceDiceLoss (:416-443) over the final probabilities and the four mask heads against max-pooled labels
// (utils/utils_3D_embed_full.py:64-80).  All three criteria are functions of four per-(sample, class) sums over the voxels,
//     A = sum p            T = sum onehot            X = sum p * onehot            S = sum -(1 - p) * onehot * log(max(p, 1e-6))
// so ONE bandwidth-bound pass over the probabilities (read once, in the reference's own [N][C][V] layout, labels as
// bytes) replaces the reference's flatten / transpose / stack / clamp / log / sum temporaries (about 12 V*C-sized fp32
// tensors per criterion pair), and its backward is one more pass:
//     dp = gA + onehot * (gX + gS * d/dp[-(1 - p) log max(p, 1e-6)]).
// The scalar algebra on the [N][C][4] sums stays in torch (lintransunet_b200/losses.py), where autograd differentiates it.
// Sums: fp32 per thread, fixed-order block tree, fp64 across blocks in a fixed order => bit-reproducible.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

constexpr int kLossThreads = 256;
constexpr int kLossPerThread = 16;                       // voxels per thread and chunk
constexpr int kLossChunk = kLossThreads * kLossPerThread;

__device__ __forceinline__ float ce_term(float p) { return -(1.f - p) * logf(fmaxf(p, 1e-6f)); }
__device__ __forceinline__ float ce_term_grad(float p) {
    // d/dp of -(1 - p) * log(clamp(p, 1e-6)): the clamp passes its gradient for p >= 1e-6 (torch.clamp backward)
    return p >= 1e-6f ? logf(p) - (1.f - p) / p : logf(1e-6f);
}

__global__ void __launch_bounds__(kLossThreads)
loss_sums_kernel(const float* __restrict__ p, const uint8_t* __restrict__ labels, double* __restrict__ ws, int C, int64_t V,
                 int chunks) {
    const int chunk = blockIdx.x, c = blockIdx.y, n = blockIdx.z;
    const float* pc = p + ((int64_t)n * C + c) * V;
    const uint8_t* lb = labels + (int64_t)n * V;
    const int64_t v0 = (int64_t)chunk * kLossChunk;
    float a = 0.f, t = 0.f, x = 0.f, s = 0.f;
#pragma unroll 4
    for (int i = 0; i < kLossPerThread; ++i) {
        const int64_t v = v0 + (int64_t)i * kLossThreads + threadIdx.x;
        if (v < V) {
            const float pv = __ldg(pc + v);
            const bool hit = __ldg(lb + v) == (uint8_t)c;
            a += pv;
            if (hit) { t += 1.f; x += pv; s += ce_term(pv); }
        }
    }
    __shared__ float red[4][kLossThreads / 32];
    a = warp_sum(a); t = warp_sum(t); x = warp_sum(x); s = warp_sum(s);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = a; red[1][warp] = t; red[2][warp] = x; red[3][warp] = s; }
    __syncthreads();
    if (threadIdx.x < 4) {
        double acc = 0.0;
#pragma unroll
        for (int w = 0; w < kLossThreads / 32; ++w) acc += (double)red[threadIdx.x][w];
        ws[(((int64_t)n * C + c) * chunks + chunk) * 4 + threadIdx.x] = acc;
    }
}

__global__ void __launch_bounds__(128)
loss_sums_finalize_kernel(const double* __restrict__ ws, float* __restrict__ sums, int chunks) {
    const int nc = blockIdx.x, k = threadIdx.x >> 5, lane = threadIdx.x & 31;    // warp k sums quantity k
    double acc = 0.0;
    for (int i = lane; i < chunks; i += 32) acc += ws[((int64_t)nc * chunks + i) * 4 + k];
    acc = warp_sum_d(acc);
    if (lane == 0) sums[nc * 4 + k] = (float)acc;
}

__global__ void __launch_bounds__(kLossThreads)
loss_sums_bwd_kernel(const float* __restrict__ p, const uint8_t* __restrict__ labels, const float* __restrict__ g,
                     float* __restrict__ dp, int C, int64_t V) {
    const int c = blockIdx.y, n = blockIdx.z;
    const float* pc = p + ((int64_t)n * C + c) * V;
    float* dc = dp + ((int64_t)n * C + c) * V;
    const uint8_t* lb = labels + (int64_t)n * V;
    const float gA = g[(n * C + c) * 4 + 0], gX = g[(n * C + c) * 4 + 2], gS = g[(n * C + c) * 4 + 3];
    const int64_t v0 = (int64_t)blockIdx.x * kLossChunk;
#pragma unroll 4
    for (int i = 0; i < kLossPerThread; ++i) {
        const int64_t v = v0 + (int64_t)i * kLossThreads + threadIdx.x;
        if (v < V) {
            const float pv = __ldg(pc + v);
            const bool hit = __ldg(lb + v) == (uint8_t)c;
            dc[v] = hit ? gA + gX + gS * ce_term_grad(pv) : gA;
        }
    }
}

// max-pool of a uint8 label volume [N][H][W][D] with kernel = stride = (kh, kw, kd) (the reference's label pyramid)
__global__ void __launch_bounds__(256)
label_pool_kernel(const uint8_t* __restrict__ x, uint8_t* __restrict__ y, int H, int W, int D, int kh, int kw, int kd,
                  int64_t total) {
    const int Ho = H / kh, Wo = W / kw, Do = D / kd;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int od = (int)(i % Do);
        int64_t r = i / Do;
        const int ow = (int)(r % Wo); r /= Wo;
        const int oh = (int)(r % Ho);
        const int64_t n = r / Ho;
        uint8_t m = 0;
        for (int a = 0; a < kh; ++a)
            for (int b = 0; b < kw; ++b)
                for (int cc = 0; cc < kd; ++cc) {
                    const uint8_t v = x[(((n * H + oh * kh + a) * W + ow * kw + b) * (int64_t)D) + od * kd + cc];
                    m = v > m ? v : m;
                }
        y[i] = m;
    }
}

}  // namespace ltu

using namespace ltu;

extern "C" size_t ltu_loss_sums_workspace(int N, int C, int64_t V) {
    const int64_t chunks = (V + kLossChunk - 1) / kLossChunk;
    return (size_t)N * C * chunks * 4 * sizeof(double);
}

// sums fp32 [N][C][4] = (sum p, sum onehot, sum p*onehot, sum -(1-p)*onehot*log(max(p,1e-6))) over the V voxels of every
// (sample, class); p fp32 [N][C][V] (the reference layout [N,C,H,W,D]), labels uint8 [N][V] with onehot_c = (label == c).
extern "C" int ltu_loss_sums(const float* p, const uint8_t* labels, float* sums, void* workspace, size_t workspace_bytes,
                             int N, int C, int64_t V, ltu_stream_t stream) {
    LTU_ARG_CHECK(p && labels && sums && workspace, "loss_sums: null pointer");
    LTU_ARG_CHECK(N > 0 && N <= 65535 && C > 0 && C <= 65535 && V > 0, "loss_sums: bad shape");
    LTU_ARG_CHECK(workspace_bytes >= ltu_loss_sums_workspace(N, C, V), "loss_sums: workspace too small");
    LTU_ARG_CHECK(((uintptr_t)workspace & 7) == 0, "loss_sums: workspace must be 8-byte aligned");
    const int64_t chunks = (V + kLossChunk - 1) / kLossChunk;
    LTU_ARG_CHECK(chunks < ((int64_t)1 << 31), "loss_sums: too many voxels");
    loss_sums_kernel<<<dim3((unsigned)chunks, C, N), kLossThreads, 0, (cudaStream_t)stream>>>(p, labels, (double*)workspace, C, V, (int)chunks);
    LTU_LAUNCH_CHECK("loss_sums");
    loss_sums_finalize_kernel<<<N * C, 128, 0, (cudaStream_t)stream>>>((const double*)workspace, sums, (int)chunks);
    LTU_LAUNCH_CHECK("loss_sums_finalize");
    count_launch(2);
    return LTU_OK;
}

// dp fp32 [N][C][V] = gA + onehot * (gX + gS * d/dp[-(1-p) log max(p,1e-6)]) for g = d(loss)/d(sums) fp32 [N][C][4]
extern "C" int ltu_loss_sums_bwd(const float* p, const uint8_t* labels, const float* gsums, float* dp, int N, int C, int64_t V,
                                 ltu_stream_t stream) {
    LTU_ARG_CHECK(p && labels && gsums && dp, "loss_sums_bwd: null pointer");
    LTU_ARG_CHECK(N > 0 && N <= 65535 && C > 0 && C <= 65535 && V > 0, "loss_sums_bwd: bad shape");
    const int64_t chunks = (V + kLossChunk - 1) / kLossChunk;
    loss_sums_bwd_kernel<<<dim3((unsigned)chunks, C, N), kLossThreads, 0, (cudaStream_t)stream>>>(p, labels, gsums, dp, C, V);
    LTU_LAUNCH_CHECK("loss_sums_bwd");
    count_launch(1);
    return LTU_OK;
}

// y uint8 [N][H/kh][W/kw][D/kd] = max over kh x kw x kd blocks of x uint8 [N][H][W][D] (F.max_pool3d with kernel = stride,
// utils/utils_3D_embed_full.py:65,:76-79); H, W, D must be multiples of the kernel.
extern "C" int ltu_label_pool(const uint8_t* x, uint8_t* y, int N, int H, int W, int D, int kh, int kw, int kd,
                              ltu_stream_t stream) {
    LTU_ARG_CHECK(x && y && N > 0 && H > 0 && W > 0 && D > 0, "label_pool: bad arguments");
    LTU_ARG_CHECK(kh >= 1 && kw >= 1 && kd >= 1 && H % kh == 0 && W % kw == 0 && D % kd == 0,
                  "label_pool: %dx%dx%d is not a multiple of the kernel %dx%dx%d", H, W, D, kh, kw, kd);
    const int64_t total = (int64_t)N * (H / kh) * (W / kw) * (D / kd);
    int64_t blocks = (total + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    label_pool_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, y, H, W, D, kh, kw, kd, total);
    LTU_LAUNCH_CHECK("label_pool");
    count_launch(1);
    return LTU_OK;
}
