// Weight gradient of nn.Conv3d on the tensor pipe, bf16 activations, fp32 accumulation:
//   dW[co][tap][ci] = sum over output voxels v of dy[v][co] * x[in(v, tap)][ci]      (zero outside the volume)
// i.e. per filter tap a GEMM  dy^T [Cout x V] . x_shifted [V x Cin]  whose K dimension is the voxel index.  Both operands
// are stored voxel-major (channels-last), so both are read "transposed": 32-voxel tiles are staged in shared memory with
// 16-byte cp.async (zero fill = padding / ragged tail / channel tail) and fed to mma.sync m16n8k16 through
// ldmatrix.trans -- the fragment pattern of kv_reduce_mma / attn_bwd_reduce_mma.
//
// grid (voxel chunks, taps, 64x64 channel blocks), 256 threads: warp (wm, wn) owns rows wm*16..+16 of the Cout block and
// columns wn*32..+32 of the Cin block.  Every CTA writes one fp32 partial tile; an ordered finalize sums the chunks, so
// the result is bit-reproducible.  First version: x is re-read once per tap (27x, mostly from L2) -- the halo form of
// conv3d_tc3 is the next step.  Any stride / padding, 1x1x1 and 3x3x3 kernels, Cin and Cout multiples of 8.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

namespace {

constexpr int kTile = 32;                                    // voxels per K tile
constexpr int kRow = 72;                                     // padded bf16 row of a 64-channel tile (144 B: conflict-free ldmatrix)

__device__ __forceinline__ void ldsm4t_w(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_w(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct WgradParams {
    const bf16* x; const bf16* dy; float* part;
    int B, Hi, Wi, Di, Cin, Ho, Wo, Do, Cout, k, sh, sw, sd, pad, up2;   // up2: the conv reads nearest-x2-upsampled x
    int64_t vout;                                            // B*Ho*Wo*Do
    int64_t vox_per_cta;                                     // multiple of kTile
    int cin_blocks, CoP, CiP;                                // channel counts padded to 64
};

}  // namespace

__global__ void __launch_bounds__(256)
conv_wgrad_mma_kernel(const WgradParams p) {
    __shared__ __align__(16) bf16 sD[2][kTile * kRow];       // dy tile  [voxel][64 output channels]
    __shared__ __align__(16) bf16 sX[2][kTile * kRow];       // x tile   [voxel][64 input channels], shifted by the tap
    const int tap = blockIdx.y;
    const int cob = blockIdx.z / p.cin_blocks, cib = blockIdx.z % p.cin_blocks;
    const int co0 = cob * 64, ci0 = cib * 64;
    const int kh = p.k == 3 ? tap / 9 : 0, kw = p.k == 3 ? (tap / 3) % 3 : 0, kd = p.k == 3 ? tap % 3 : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wm = warp & 3, wn = warp >> 2;
    const int mi = lane >> 3, lr = lane & 7;
    const int64_t v0 = (int64_t)blockIdx.x * p.vox_per_cta;
    int64_t v1 = v0 + p.vox_per_cta;
    if (v1 > p.vout) v1 = p.vout;
    const int ntiles = v1 > v0 ? (int)ceil_div64(v1 - v0, kTile) : 0;

    // staging role of this thread: voxel row r of the tile, 16-byte channel chunk c
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    const bool co_ok = co0 + c * 8 < p.Cout, ci_ok = ci0 + c * 8 < p.Cin;
    auto stage = [&](int buf, int tile) {
        const int64_t v = v0 + (int64_t)tile * kTile + r;
        const bool vok = v < v1;
        const bf16* gd = p.dy;
        const bf16* gx = p.x;
        bool xok = false;
        if (vok) {
            int64_t t = v;
            const int dO = (int)(t % p.Do); t /= p.Do;
            const int wO = (int)(t % p.Wo); t /= p.Wo;
            const int hO = (int)(t % p.Ho);
            const int b = (int)(t / p.Ho);
            gd = p.dy + v * p.Cout + co0 + c * 8;
            const int hi = hO * p.sh + kh - p.pad, wi = wO * p.sw + kw - p.pad, di = dO * p.sd + kd - p.pad;
            const int e = p.up2 ? 2 : 1;                     // bounds in the (virtually upsampled) input, then the source voxel
            xok = hi >= 0 && hi < e * p.Hi && wi >= 0 && wi < e * p.Wi && di >= 0 && di < e * p.Di;
            if (xok) gx = p.x + ((((int64_t)b * p.Hi + hi / e) * p.Wi + wi / e) * p.Di + di / e) * p.Cin + ci0 + c * 8;
        }
        cp_async16_zfill(sD[buf] + r * kRow + c * 8, (vok && co_ok) ? gd : p.dy, (vok && co_ok) ? 16 : 0);
        cp_async16_zfill(sX[buf] + r * kRow + c * 8, (xok && ci_ok) ? gx : p.x, (xok && ci_ok) ? 16 : 0);
    };

    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[nt][i] = 0.f;

    if (ntiles > 0) stage(0, 0);
    cp_async_commit();
    for (int t = 0; t < ntiles; ++t) {
        const int buf = t & 1;
        if (t + 1 < ntiles) stage(buf ^ 1, t + 1);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();                                     // tile t landed for every thread
        const bf16* tD = sD[buf];
        const bf16* tX = sX[buf];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int rr = ks * 16;
            uint32_t a[4], bq[2][4];
            // A^T: matrices (k lo, m lo), (k lo, m hi), (k hi, m lo), (k hi, m hi); B: (k lo, n lo), (k hi, n lo), (k lo, n hi), (k hi, n hi)
            ldsm4t_w(smem_u32_generic(tD + (rr + lr + 8 * (mi >> 1)) * kRow + wm * 16 + 8 * (mi & 1)), a);
#pragma unroll
            for (int np = 0; np < 2; ++np)
                ldsm4t_w(smem_u32_generic(tX + (rr + lr + 8 * (mi & 1)) * kRow + wn * 32 + np * 16 + 8 * (mi >> 1)), bq[np]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_w(acc[nt], a, bq[nt >> 1][(nt & 1) * 2], bq[nt >> 1][(nt & 1) * 2 + 1]);
        }
        __syncthreads();                                     // the buffer is free for tile t + 2
    }
    cp_async_wait<0>();

    // partial tile: part[((chunk * taps + tap) * CoP + co) * CiP + ci]
    const int gq = lane >> 2, tq = lane & 3;
    float* out = p.part + (((int64_t)blockIdx.x * gridDim.y + tap) * p.CoP + co0 + wm * 16 + gq) * p.CiP + ci0 + wn * 32 + 2 * tq;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        *reinterpret_cast<float2*>(out + nt * 8) = make_float2(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<float2*>(out + (int64_t)8 * p.CiP + nt * 8) = make_float2(acc[nt][2], acc[nt][3]);
    }
}

// Small layers (Cin, Cout <= CT in {16, 32}): a 64x64 channel block would leave 15/16 (3/4) of the MMAs multiplying zeros
// and a 32-voxel tile moves 2 KB per CTA step.  Here the tile is 256 voxels x CT channels (one voxel row per thread), the
// eight warps split the K dimension (32 voxels each) and their CT x CT accumulators are summed in a fixed order through
// shared memory at the end.  Single-stage tiles; 4-5 CTAs per SM hide the load latency.
template <int CT>
__global__ void __launch_bounds__(256)
conv_wgrad_small_kernel(const WgradParams p) {
    constexpr int ROW = CT + 8, NCH = CT / 8, MT = CT / 16, NT = CT / 8, VT = 256;
    __shared__ __align__(16) unsigned char smem_raw[2 * VT * ROW * 2];
    bf16* sD = reinterpret_cast<bf16*>(smem_raw);            // dy tile [voxel][CT]
    bf16* sX = sD + VT * ROW;                                // x tile  [voxel][CT], shifted by the tap
    static_assert(8 * CT * CT * 4 <= 2 * VT * ROW * 2, "the reduction buffer reuses the tiles");
    const int tap = blockIdx.y;
    const int kh = p.k == 3 ? tap / 9 : 0, kw = p.k == 3 ? (tap / 3) % 3 : 0, kd = p.k == 3 ? tap % 3 : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mi = lane >> 3, lr = lane & 7;
    const int64_t v0 = (int64_t)blockIdx.x * p.vox_per_cta;
    int64_t v1 = v0 + p.vox_per_cta;
    if (v1 > p.vout) v1 = p.vout;
    const int ntiles = v1 > v0 ? (int)ceil_div64(v1 - v0, VT) : 0;
    const int r = threadIdx.x;                               // this thread stages voxel row r of every tile

    float acc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[mt][nt][i] = 0.f;

    for (int t = 0; t < ntiles; ++t) {
        const int64_t v = v0 + (int64_t)t * VT + r;
        const bool vok = v < v1;
        const bf16* gd = p.dy;
        const bf16* gx = p.x;
        bool xok = false;
        if (vok) {
            int64_t q = v;
            const int dO = (int)(q % p.Do); q /= p.Do;
            const int wO = (int)(q % p.Wo); q /= p.Wo;
            const int hO = (int)(q % p.Ho);
            const int b = (int)(q / p.Ho);
            gd = p.dy + v * p.Cout;
            const int hi = hO * p.sh + kh - p.pad, wi = wO * p.sw + kw - p.pad, di = dO * p.sd + kd - p.pad;
            const int e = p.up2 ? 2 : 1;
            xok = hi >= 0 && hi < e * p.Hi && wi >= 0 && wi < e * p.Wi && di >= 0 && di < e * p.Di;
            if (xok) gx = p.x + ((((int64_t)b * p.Hi + hi / e) * p.Wi + wi / e) * p.Di + di / e) * p.Cin;
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const bool dok = vok && c * 8 < p.Cout, iok = xok && c * 8 < p.Cin;
            cp_async16_zfill(sD + r * ROW + c * 8, dok ? gd + c * 8 : p.dy, dok ? 16 : 0);
            cp_async16_zfill(sX + r * ROW + c * 8, iok ? gx + c * 8 : p.x, iok ? 16 : 0);
        }
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
            const int rr = warp * 32 + ks * 16;
            uint32_t a[MT][4], bq[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
                ldsm4t_w(smem_u32_generic(sD + (rr + lr + 8 * (mi >> 1)) * ROW + mt * 16 + 8 * (mi & 1)), a[mt]);
#pragma unroll
            for (int np = 0; np < MT; ++np)
                ldsm4t_w(smem_u32_generic(sX + (rr + lr + 8 * (mi & 1)) * ROW + np * 16 + 8 * (mi >> 1)), bq[np]);
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    mma_w(acc[mt][nt], a[mt], bq[nt >> 1][(nt & 1) * 2], bq[nt >> 1][(nt & 1) * 2 + 1]);
        }
        __syncthreads();                                     // the tiles are free again
    }
    // ordered sum of the eight warps' accumulators
    float* red = reinterpret_cast<float*>(smem_raw);         // [8][CT][CT]
    const int gq = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            float* dst = red + ((warp * CT) + mt * 16 + gq) * CT + nt * 8 + 2 * tq;
            dst[0] = acc[mt][nt][0]; dst[1] = acc[mt][nt][1];
            dst[8 * CT] = acc[mt][nt][2]; dst[8 * CT + 1] = acc[mt][nt][3];
        }
    __syncthreads();
    float* out = p.part + ((int64_t)blockIdx.x * gridDim.y + tap) * CT * CT;
    for (int i = threadIdx.x; i < CT * CT; i += 256) {
        float sum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) sum += red[w * CT * CT + i];
        out[i] = sum;
    }
}

static inline int wgrad_small_chunks(int64_t vout, int taps) {
    int64_t want = ceil_div64((int64_t)sm_count() * 8, taps);        // about two waves of 4 CTAs / SM
    const int64_t most = ceil_div64(vout, 4 * 256);                  // at least four 256-voxel tiles per CTA
    if (want > most) want = most;
    return (int)(want < 1 ? 1 : want);
}
static inline int wgrad_small_ct(int Cin, int Cout) {                // 0: use the 64x64 block kernel
    const int m = Cin > Cout ? Cin : Cout;
    return m <= 16 ? 16 : (m <= 32 ? 32 : 0);
}

// dw[tap][co][ci] = sum over the chunks (in order) of the partial tiles
__global__ void __launch_bounds__(256)
conv_wgrad_finalize_kernel(const float* __restrict__ part, float* __restrict__ dw, int chunks, int taps, int Cout, int Cin,
                           int CoP, int CiP) {
    const int64_t n = (int64_t)taps * Cout * Cin;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int ci = (int)(i % Cin);
        const int64_t t = i / Cin;
        const int co = (int)(t % Cout), tap = (int)(t / Cout);
        float s = 0.f;
        for (int ch = 0; ch < chunks; ++ch) s += part[(((int64_t)ch * taps + tap) * CoP + co) * CiP + ci];
        dw[i] = s;
    }
}

static inline int wgrad_chunks(int64_t vout, int taps, int blocks) {
    // about four resident waves of CTAs (3 CTAs / SM), at least 8 voxel tiles per CTA
    int64_t want = ceil_div64((int64_t)sm_count() * 12, (int64_t)taps * blocks);
    const int64_t most = ceil_div64(vout, 8 * kTile);
    if (want > most) want = most;
    if (want < 1) want = 1;
    return (int)want;
}

}  // namespace ltu

using namespace ltu;

extern "C" size_t ltu_conv3d_wgrad_workspace(int B, int Ho, int Wo, int Do, int Cin, int Cout, int ksize) {
    if (B <= 0 || Ho <= 0 || Wo <= 0 || Do <= 0 || Cin <= 0 || Cout <= 0 || !(ksize == 1 || ksize == 3)) return 0;
    const int taps = ksize * ksize * ksize, cob = (Cout + 63) / 64, cib = (Cin + 63) / 64;
    const int ct = wgrad_small_ct(Cin, Cout);
    if (ct) return (size_t)wgrad_small_chunks((int64_t)B * Ho * Wo * Do, taps) * taps * ct * ct * sizeof(float);
    const int chunks = wgrad_chunks((int64_t)B * Ho * Wo * Do, taps, cob * cib);
    return (size_t)chunks * taps * cob * 64 * cib * 64 * sizeof(float);
}

extern "C" int ltu_conv3d_wgrad(const void* x, const void* dy, float* dw, void* ws, size_t ws_bytes, int B, int Hi, int Wi,
                                int Di, int Cin, int Ho, int Wo, int Do, int Cout, int ksize, int sh, int sw, int sd, int pad,
                                int up2, ltu_stream_t stream) {
    LTU_ARG_CHECK(x && dy && dw && ws, "conv3d_wgrad: null pointer");
    LTU_ARG_CHECK(B > 0 && Hi > 0 && Wi > 0 && Di > 0 && Ho > 0 && Wo > 0 && Do > 0, "conv3d_wgrad: bad shape");
    LTU_ARG_CHECK(ksize == 1 || ksize == 3, "conv3d_wgrad: kernel size must be 1 or 3");
    LTU_ARG_CHECK(Cin % 8 == 0 && Cout % 8 == 0 && Cin <= 1024 && Cout <= 1024, "conv3d_wgrad: Cin=%d and Cout=%d must be multiples of 8", Cin, Cout);
    LTU_ARG_CHECK(sh >= 1 && sw >= 1 && sd >= 1 && pad >= 0, "conv3d_wgrad: bad stride / padding");
    const int ue = up2 ? 2 : 1;
    LTU_ARG_CHECK(Ho == (ue * Hi + 2 * pad - ksize) / sh + 1 && Wo == (ue * Wi + 2 * pad - ksize) / sw + 1 &&
                  Do == (ue * Di + 2 * pad - ksize) / sd + 1,
                  "conv3d_wgrad: output size does not match input size, kernel, stride and padding");
    LTU_ARG_CHECK(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(dy)) & 15) == 0, "conv3d_wgrad: pointers must be 16-byte aligned");
    LTU_ARG_CHECK(ws_bytes >= ltu_conv3d_wgrad_workspace(B, Ho, Wo, Do, Cin, Cout, ksize), "conv3d_wgrad: workspace too small");
    WgradParams p;
    p.x = (const bf16*)x; p.dy = (const bf16*)dy; p.part = (float*)ws;
    p.B = B; p.Hi = Hi; p.Wi = Wi; p.Di = Di; p.Cin = Cin; p.Ho = Ho; p.Wo = Wo; p.Do = Do; p.Cout = Cout;
    p.k = ksize; p.sh = sh; p.sw = sw; p.sd = sd; p.pad = pad; p.up2 = up2 ? 1 : 0;
    p.vout = (int64_t)B * Ho * Wo * Do;
    const int taps = ksize * ksize * ksize, cob = (Cout + 63) / 64, cib = (Cin + 63) / 64;
    cudaStream_t st = (cudaStream_t)stream;
    const int ct = wgrad_small_ct(Cin, Cout);
    int chunks;
    if (ct) {
        chunks = wgrad_small_chunks(p.vout, taps);
        p.vox_per_cta = ceil_div64(ceil_div64(p.vout, chunks), 256) * 256;
        p.cin_blocks = 1; p.CoP = ct; p.CiP = ct;
        if (ct == 16) conv_wgrad_small_kernel<16><<<dim3(chunks, taps), 256, 0, st>>>(p);
        else conv_wgrad_small_kernel<32><<<dim3(chunks, taps), 256, 0, st>>>(p);
    } else {
        chunks = wgrad_chunks(p.vout, taps, cob * cib);
        p.vox_per_cta = ceil_div64(ceil_div64(p.vout, chunks), kTile) * kTile;
        p.cin_blocks = cib; p.CoP = cob * 64; p.CiP = cib * 64;
        conv_wgrad_mma_kernel<<<dim3(chunks, taps, cob * cib), 256, 0, st>>>(p);
    }
    LTU_LAUNCH_CHECK("conv3d_wgrad");
    const int64_t n = (int64_t)taps * Cout * Cin;
    int64_t fb = ceil_div64(n, 256);
    if (fb > (int64_t)sm_count() * 8) fb = (int64_t)sm_count() * 8;
    conv_wgrad_finalize_kernel<<<(unsigned)fb, 256, 0, st>>>(p.part, dw, chunks, taps, Cout, Cin, p.CoP, p.CiP);
    LTU_LAUNCH_CHECK("conv3d_wgrad_finalize");
    count_launch(2);
    return LTU_OK;
}
