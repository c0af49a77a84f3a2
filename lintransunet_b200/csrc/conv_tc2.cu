// Persistent, warp-specialised version of the tcgen05 implicit-GEMM convolution (conv_tc.cu has
// the one-tile-per-CTA kernel; same math, same operand layouts, same partials).
//
//   grid  = min(#tiles, #SMs) CTAs of 288 threads, each looping over tiles T, T+grid, ...
//   warps 0-3  im2col / weight producers (16-byte cp.async into the K-major SWIZZLE_128B ring)
//   warp  4    one elected lane issues tcgen05.mma into one of TWO TMEM accumulator buffers
//   warps 5-8  epilogue: tcgen05.ld -> bias -> (GELU | residual+LayerNorm) -> store + IN partials
//
// Three mbarrier pipelines keep their phase across tiles: smem full/empty (producers <-> MMA),
// TMEM full/empty (MMA <-> epilogue).  The epilogue of tile i therefore overlaps the gather and the
// MMAs of tile i+1, and TMEM allocation / barrier set-up is paid once per CTA instead of once per
// 128 output voxels -- which is what short-K layers (1x1x1 convs, Linear layers, small Cin) need.
//
// With ksize = 1 the "convolution" is a Linear layer on a [rows, Cin] token matrix; the epilogue
// then optionally fuses the exact-erf GELU (trans_block.py:208) or the residual add + LayerNorm
// (trans_block.py:205-206, :209-210), for which one thread owns one complete output row.
#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);

constexpr int kTc2Threads = 288;
constexpr int kEpiFirstWarp = 5;

enum : int { kEpiPlain = 0, kEpiGelu = 1, kEpiResLN = 2 };

struct Tc2Params {
    TcParams c;                 // geometry / operands (stages, tmem_cols unused)
    int nstages;
    int ncols;                  // TMEM columns of ONE accumulator buffer (power of two >= 32)
    int nclass;                 // 8 in fold mode, else 1
    int tiles_x;                // row tiles per (sample, class)
    int64_t rows;               // GEMM rows per (sample, class)
    int total_tiles;
    int ld_out;                 // output row stride in elements (>= Cstore; lets a call fill a column slice)
    int epi;                    // kEpiPlain | kEpiGelu | kEpiResLN
    const bf16* residual;       // [rows][Cstore] (kEpiResLN)
    const float* gamma; const float* beta; float eps;
};

__global__ void __launch_bounds__(kTc2Threads, 1)
conv3d_tc2_kernel(const Tc2Params q) {
    const TcParams& p = q.c;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = kTcM * 128;
    const int b_bytes = p.Cout * 128;
    const int stage_bytes = a_bytes + b_bytes;
    unsigned char* tail = smem + (size_t)q.nstages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);            // [nstages]
    uint64_t* empty_bar = full_bar + q.nstages;                         // [nstages]
    uint64_t* tfull_bar = empty_bar + q.nstages;                        // [2]
    uint64_t* tempty_bar = tfull_bar + 2;                               // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* sred = reinterpret_cast<float*>(tmem_slot + 2);              // [2 buffers][4 warps][Cout][2]
    int* toff = reinterpret_cast<int*>(sred + 2 * 4 * p.Cout * 2);      // [8 classes][32 taps]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = p.Kpad / kTcBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < q.nstages; ++s) {
            mbar_init(smem_u32(full_bar + s), kTcProducers);
            mbar_init(smem_u32(empty_bar + s), 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(tfull_bar + i), 1); mbar_init(smem_u32(tempty_bar + i), 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 256) {                                            // tap offset table of every parity class
        const int z = threadIdx.x >> 5, t = threadIdx.x & 31;
        int off = 0;
        if (t < p.ntaps && z < q.nclass) {
            int dh, dw, dd;
            if (p.fold) { dh = ((t >> 2) & 1) - 1 + ((z >> 2) & 1); dw = ((t >> 1) & 1) - 1 + ((z >> 1) & 1); dd = (t & 1) - 1 + (z & 1); }
            else if (p.ks == 3) { dh = t / 9 - 1; dw = (t / 3) % 3 - 1; dd = t % 3 - 1; }
            else { dh = dw = dd = 0; }
            off = (dh * p.Wi + dw) * p.Di + dd;
        }
        toff[z * 32 + t] = off;
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * q.ncols)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =========================== producers ===========================
        const int jc = threadIdx.x & 7, rbase = threadIdx.x >> 3;
        const int Cin = p.C0 + p.C1;
        const int b_iters = p.Cout / 16;
        uint32_t kbc = 0;                                               // k-blocks issued by this CTA so far
        for (int T = blockIdx.x; T < q.total_tiles; T += gridDim.x) {
            const int tx = T % q.tiles_x;
            const int z = (T / q.tiles_x) % q.nclass;
            const int b = T / (q.tiles_x * q.nclass);
            const int pa = (z >> 2) & 1, pb = (z >> 1) & 1, pc = z & 1;
            const int64_t vox0 = (int64_t)tx * kTcM;
            int ctrv[8];
            uint32_t vmask[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int rr = rbase + 16 * i;
                int64_t rid = vox0 + rr;
                const bool ok = rid < q.rows;
                if (!ok) rid = 0;
                int ch, cw, cd;
                uint32_t m = 0;
                if (p.fold) {
                    cd = (int)(rid % p.Di);
                    const int64_t t2 = rid / p.Di;
                    cw = (int)(t2 % p.Wi); ch = (int)(t2 / p.Wi);
                    uint32_t hm = 0, wm = 0, dm = 0;
#pragma unroll
                    for (int t = 0; t < 2; ++t) {
                        const int hv = ch + t - 1 + pa, wv = cw + t - 1 + pb, dv = cd + t - 1 + pc;
                        hm |= (hv >= 0 && hv < p.Hi) ? 1u << t : 0u;
                        wm |= (wv >= 0 && wv < p.Wi) ? 1u << t : 0u;
                        dm |= (dv >= 0 && dv < p.Di) ? 1u << t : 0u;
                    }
#pragma unroll
                    for (int th = 0; th < 2; ++th)
#pragma unroll
                        for (int tw = 0; tw < 2; ++tw)
                            if (((hm >> th) & 1u) && ((wm >> tw) & 1u)) m |= dm << (th * 4 + tw * 2);
                } else if (p.ks == 3) {
                    const int od = (int)(rid % p.Do);
                    const int64_t t2 = rid / p.Do;
                    const int ow = (int)(t2 % p.Wo), oh = (int)(t2 / p.Wo);
                    ch = oh * p.sh; cw = ow * p.sw; cd = od * p.sd;
                    uint32_t hm = 0, wm = 0, dm = 0;
#pragma unroll
                    for (int t = 0; t < 3; ++t) {
                        const int hv = ch + t - 1, wv = cw + t - 1, dv = cd + t - 1;
                        hm |= (hv >= 0 && hv < p.Hi) ? 1u << t : 0u;
                        wm |= (wv >= 0 && wv < p.Wi) ? 1u << t : 0u;
                        dm |= (dv >= 0 && dv < p.Di) ? 1u << t : 0u;
                    }
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh)
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
                            if (((hm >> kh) & 1u) && ((wm >> kw) & 1u)) m |= dm << (kh * 9 + kw * 3);
                } else {                                                // 1x1x1 / Linear: row == source voxel
                    if (p.sh == 1 && p.sw == 1 && p.sd == 1) { ch = 0; cw = 0; cd = (int)rid; }
                    else {
                        const int od = (int)(rid % p.Do);
                        const int64_t t2 = rid / p.Do;
                        const int ow = (int)(t2 % p.Wo), oh = (int)(t2 / p.Wo);
                        ch = oh * p.sh; cw = ow * p.sw; cd = od * p.sd;
                    }
                    m = 1u;
                }
                vmask[i] = ok ? m : 0u;
                ctrv[i] = (int)((int64_t)b * p.Hi * p.Wi * p.Di + ((int64_t)ch * p.Wi + cw) * p.Di + cd);
            }
            const bf16* wbase = p.weight + (int64_t)z * p.w_class_stride;
            const int* tz = toff + z * 32;
            for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                const int stage = kbc % q.nstages;
                const uint32_t round = kbc / q.nstages;
                mbar_wait(smem_u32(empty_bar + stage), (round & 1) ^ 1);
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint32_t sb = sa + a_bytes;
                const int kk = kb * kTcBK + jc * 8;
                const int tap = kk >> p.log2cin;
                const int c = kk & (Cin - 1);
                const int tvo = tz[tap & 31];
                const bool first = c < p.C0;
                const bf16* cbase = first ? p.in0 + c : p.in1 + (c - p.C0);
                const int cstride = first ? p.C0 : p.C1;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int rr = rbase + 16 * i;
                    const bool ok = (vmask[i] >> tap) & 1u;
                    const bf16* src = cbase + (int64_t)(ctrv[i] + tvo) * cstride;
                    cp16(sa + rr * 128 + ((jc ^ (rr & 7)) << 4), ok ? src : p.in0, ok ? 16 : 0);
                }
                for (int i = 0; i < b_iters; ++i) {
                    const int qi = threadIdx.x + i * kTcProducers;
                    const int n = qi >> 3, j = qi & 7;
                    cp16(sb + n * 128 + ((j ^ (n & 7)) << 4), wbase + (int64_t)n * p.Kpad + kb * kTcBK + j * 8, 16);
                }
                cp_async_arrive_noinc(smem_u32(full_bar + stage));
            }
        }
    } else if (warp == 4) {
        // =========================== MMA issuer ===========================
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
        uint32_t kbc = 0, it = 0;
        for (int T = blockIdx.x; T < q.total_tiles; T += gridDim.x, ++it) {
            const uint32_t abuf = it & 1;
            mbar_wait(smem_u32(tempty_bar + abuf), ((it >> 1) & 1) ^ 1);       // epilogue drained this accumulator
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + abuf * (uint32_t)q.ncols;
            for (int kb = 0; kb < nkb; ++kb, ++kbc) {
                const int stage = kbc % q.nstages;
                const uint32_t round = kbc / q.nstages;
                mbar_wait(smem_u32(full_bar + stage), round & 1);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0) {
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + a_bytes);
#pragma unroll
                    for (int k = 0; k < kTcBK / 16; ++k)
                        umma_bf16(tacc, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit(smem_u32(empty_bar + stage));
                    if (kb == nkb - 1) umma_commit(smem_u32(tfull_bar + abuf));
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== epilogue ===========================
        const int qw = warp & 3;                                         // TMEM lane quarter of this warp
        const int r = qw * 32 + lane;                                    // tile row == TMEM lane
        const int et = threadIdx.x - kEpiFirstWarp * 32;                 // 0..127
        uint32_t it = 0;
        for (int T = blockIdx.x; T < q.total_tiles; T += gridDim.x, ++it) {
            const int tx = T % q.tiles_x;
            const int z = (T / q.tiles_x) % q.nclass;
            const int b = T / (q.tiles_x * q.nclass);
            const uint32_t abuf = it & 1;
            int64_t id = (int64_t)tx * kTcM + r;
            const bool row_ok = id < q.rows;
            if (!row_ok) id = 0;
            int64_t out_vox = id;
            if (p.fold) {
                const int c_ = (int)(id % p.Di);
                const int64_t t2 = id / p.Di;
                const int b_ = (int)(t2 % p.Wi), a_ = (int)(t2 / p.Wi);
                out_vox = ((int64_t)(2 * a_ + ((z >> 2) & 1)) * p.Wo + (2 * b_ + ((z >> 1) & 1))) * p.Do + (2 * c_ + (z & 1));
            }
            const int64_t vrow = (int64_t)b * (int64_t)p.Ho * p.Wo * p.Do + out_vox;
            const int64_t out_row = vrow * q.ld_out;
            const int64_t res_row = vrow * p.Cstore;
            float* red = sred + abuf * (4 * p.Cout * 2);

            mbar_wait(smem_u32(tfull_bar + abuf), (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + abuf * (uint32_t)q.ncols + ((uint32_t)(qw * 32) << 16);

            float ln_mean = 0.f, ln_rstd = 1.f;
            if (q.epi == kEpiResLN) {
                // pass 1: x = acc + bias + residual, written back to TMEM; row statistics are thread-local
                float s = 0.f, ss = 0.f;
                for (int c0 = 0; c0 < p.Cstore; c0 += 32) {
                    float v[32];
                    tmem_ld32(tacc + (uint32_t)c0, v);
                    float rres[32];
                    if (row_ok) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float t8[8];
                            load_vec(q.residual + res_row + c0 + g * 8, t8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) rres[g * 8 + i] = t8[i];
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) rres[i] = 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        v[i] = v[i] + __ldg(p.bias + c0 + i) + rres[i];
                        s += v[i];
                    }
                    tmem_st32(tacc + (uint32_t)c0, v);
                }
                ln_mean = s / (float)p.Cstore;
                for (int c0 = 0; c0 < p.Cstore; c0 += 32) {
                    float v[32];
                    tmem_ld32(tacc + (uint32_t)c0, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) { const float d = v[i] - ln_mean; ss = fmaf(d, d, ss); }
                }
                ln_rstd = rsqrtf(ss / (float)p.Cstore + q.eps);
            }

            for (int c0 = 0; c0 < p.Cstore; c0 += 32) {
                float v[32];
                tmem_ld32(tacc + (uint32_t)c0, v);
                const int ncol = (p.Cstore - c0) < 32 ? (p.Cstore - c0) : 32;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float o;
                    if (q.epi == kEpiResLN) {
                        o = (v[i] - ln_mean) * ln_rstd * __ldg(q.gamma + c0 + i) + __ldg(q.beta + c0 + i);
                    } else {
                        const float bsv = (p.bias != nullptr && i < ncol) ? __ldg(p.bias + c0 + i) : 0.f;
                        o = v[i] + bsv;
                        if (q.epi == kEpiGelu) o = 0.5f * o * (1.f + erff(o * 0.70710678118654752f));
                    }
                    if (!p.out_f32) o = __bfloat162float(__float2bfloat16_rn(o));
                    v[i] = (row_ok && i < ncol) ? o : 0.f;
                }
                if (row_ok) {
                    if (p.out_f32) {
                        float* dst = reinterpret_cast<float*>(p.out) + out_row + c0;
                        if (ncol == 32 && (p.Cstore & 3) == 0) {
#pragma unroll
                            for (int g = 0; g < 8; ++g)
                                *reinterpret_cast<float4*>(dst + g * 4) = make_float4(v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (i < ncol) dst[i] = v[i];
                        }
                    } else {
                        bf16* dst = reinterpret_cast<bf16*>(p.out) + out_row + c0;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (g * 8 < ncol) {
                                uint4 o;
                                o.x = pack_bf16x2(v[g * 8 + 0], v[g * 8 + 1]); o.y = pack_bf16x2(v[g * 8 + 2], v[g * 8 + 3]);
                                o.z = pack_bf16x2(v[g * 8 + 4], v[g * 8 + 5]); o.w = pack_bf16x2(v[g * 8 + 6], v[g * 8 + 7]);
                                *reinterpret_cast<uint4*>(dst + g * 8) = o;
                            }
                        }
                    }
                }
                if (p.partials != nullptr) {
                    float sq[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) sq[i] = v[i] * v[i];
                    const float s = transpose_reduce32(v, lane);
                    const float qq = transpose_reduce32(sq, lane);
                    if (lane < ncol) {
                        red[((qw * p.Cstore) + c0 + lane) * 2] = s;
                        red[((qw * p.Cstore) + c0 + lane) * 2 + 1] = qq;
                    }
                }
            }
            // all four epilogue warps are done reading this accumulator -> hand it back to the MMA warp
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (et == 0) mbar_arrive(smem_u32(tempty_bar + abuf));
            if (p.partials != nullptr) {
                for (int c = et; c < p.Cstore; c += 128) {
                    float s = 0.f, qq = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) { s += red[(w * p.Cstore + c) * 2]; qq += red[(w * p.Cstore + c) * 2 + 1]; }
                    float* dst = p.partials + (((int64_t)b * p.tiles + (int64_t)z * q.tiles_x + tx) * p.Cstore + c) * 2;
                    dst[0] = s; dst[1] = qq;
                }
                // `red` of this buffer is rewritten two tiles later, after the next tile's bar.sync
            }
        }
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * q.ncols)) : "memory");
    }
}

static inline int ilog2i(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

// Shared launcher: `c` is a fully populated TcParams (as built by ltu_conv3d_tc).
int conv3d_tc2_launch(const TcParams& c, int B, int epi, const void* residual, const float* gamma, const float* beta,
                      float eps, int ld_out, cudaStream_t st) {
    Tc2Params q;
    q.c = c;
    q.ld_out = ld_out > 0 ? ld_out : c.Cstore;
    q.nclass = c.fold ? 8 : 1;
    q.rows = c.fold ? (int64_t)c.Hi * c.Wi * c.Di : (int64_t)c.Ho * c.Wo * c.Do;
    q.tiles_x = (int)ceil_div64(q.rows, kTcM);
    q.total_tiles = q.tiles_x * q.nclass * B;
    int cols = 32; while (cols < c.Cout) cols <<= 1;
    q.ncols = cols;
    q.epi = epi; q.residual = (const bf16*)residual; q.gamma = gamma; q.beta = beta; q.eps = eps;
    const int stage_bytes = kTcM * 128 + c.Cout * 128;
    int stages = (196 * 1024) / stage_bytes;
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    q.nstages = stages;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 4) * 8 + 16 + (size_t)2 * 4 * c.Cout * 2 * 4 + 8 * 32 * 4;
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(conv3d_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured_dev = dev;
    }
    int grid = sm_count();
    if (grid > q.total_tiles) grid = q.total_tiles;
    conv3d_tc2_kernel<<<grid, kTc2Threads, smem, st>>>(q);
    LTU_LAUNCH_CHECK("conv3d_tc2");
    count_launch(1);
    return LTU_OK;
}

}  // namespace ltu

using namespace ltu;

// Linear layer on a token matrix with a fused epilogue:  y = epi(x W^T + b)
//   x bf16 [rows][Cin] (Cin a power of two in [8,1024]), weight_bf16 = ltu_conv3d_tc packing of the
//   [Cout][Cin] matrix ([Cout16][kpad(Cin,1)]), Cout <= 256 and % 8 == 0, y bf16 [rows][ld_y] (the first
//   Cout columns of every row are written, so a wide layer is computed as column slices).
//   epi 0: bias; 1: bias + exact-erf GELU; 2: LayerNorm(x W^T + b + residual) * gamma + beta.
extern "C" int ltu_linear_tc(const void* x, int Cin, int64_t rows, const void* weight_bf16, const float* bias, int Cout,
                             void* y, int ld_y, int epi, const void* residual, const float* gamma, const float* beta,
                             float eps, ltu_stream_t stream) {
    LTU_ARG_CHECK(ld_y >= Cout && ld_y % 8 == 0, "linear_tc: bad output row stride %d", ld_y);
    LTU_ARG_CHECK(x && weight_bf16 && y && bias, "linear_tc: null pointer");
    LTU_ARG_CHECK(rows > 0 && rows < ((int64_t)1 << 31), "linear_tc: bad row count");
    LTU_ARG_CHECK(Cin >= 8 && Cin <= 1024 && (Cin & (Cin - 1)) == 0, "linear_tc: Cin must be a power of two in [8,1024]");
    LTU_ARG_CHECK(Cout >= 8 && Cout <= 256 && Cout % 8 == 0, "linear_tc: Cout must be a multiple of 8, <= 256");
    LTU_ARG_CHECK(epi >= 0 && epi <= 2, "linear_tc: bad epilogue %d", epi);
    LTU_ARG_CHECK(epi != kEpiResLN || (residual && gamma && beta && Cout % 32 == 0), "linear_tc: LayerNorm epilogue needs residual/gamma/beta and Cout %% 32 == 0");
    LTU_ARG_CHECK(((uintptr_t)x & 15) == 0 && ((uintptr_t)weight_bf16 & 15) == 0 && ((uintptr_t)y & 15) == 0 &&
                  ((uintptr_t)residual & 15) == 0, "linear_tc: pointers must be 16-byte aligned");
    TcParams c;
    c.in0 = (const bf16*)x; c.in1 = nullptr; c.C0 = Cin; c.C1 = 0; c.log2cin = ilog2i(Cin);
    c.Hi = 1; c.Wi = 1; c.Di = (int)rows; c.up2 = 0; c.ks = 1; c.pad = 0; c.sh = c.sw = c.sd = 1;
    c.weight = (const bf16*)weight_bf16; c.ntaps = 1; c.Ktot = Cin; c.Kpad = (Cin + kTcBK - 1) / kTcBK * kTcBK;
    c.bias = bias; c.Cstore = Cout; c.naux = 0; c.aux = nullptr; c.Cout = (Cout + 15) / 16 * 16; c.out = y; c.out_f32 = 0;
    c.Ho = 1; c.Wo = 1; c.Do = (int)rows; c.partials = nullptr; c.tiles = 0; c.stages = 0; c.tmem_cols = 0;
    c.fold = 0; c.w_class_stride = 0;
    return conv3d_tc2_launch(c, 1, epi, residual, gamma, beta, eps, ld_y, (cudaStream_t)stream);
}
