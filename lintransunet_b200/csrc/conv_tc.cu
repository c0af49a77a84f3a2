// Family 2 (bf16 path): implicit-GEMM 3x3x3 convolution on the 5th-generation tensor cores.
//
//   D[M=128 output voxels, N=Cout] += A[M, K] * B[N, K]^T,   K = 27 taps x Cin  (bf16, fp32 accum)
//
// * one CTA = one tile of 128 consecutive output voxels of one sample x all Cout channels
//   (Cout <= 256 = one UMMA N, accumulator = Cout TMEM columns);
// * warps 0-3: im2col producers.  Thread r gathers row r of the A tile (one output voxel, 64 K
//   values = 128 B) and its share of the weight tile with 16-byte cp.async into the canonical
//   K-major SWIZZLE_128B shared-memory layout; zero-fill implements padding, the K tail and the
//   ragged last tile.  Completion is tracked by mbarriers (cp.async.mbarrier.arrive.noinc);
// * warp 4: one elected lane issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=Cout, K=16)
//   and releases pipeline stages with tcgen05.commit;
// * epilogue (warps 0-3 again): tcgen05.ld 32x32b, + bias, bf16 store, and the per-tile
//   InstanceNorm partial sums (sum, sum of squares from the fp32 accumulators) reduced with a
//   fixed-order transpose-butterfly => bit-reproducible.
//
// Input addressing covers stride 1/2 per axis and a second (channel-concatenated) input (reference
// model/Unet_3Dblock.py:553).  nn.Upsample(nearest x2) + 3x3x3 conv of up_embed (:421-422) is FOLDED:
// output voxels of parity class (pa,pb,pc) see only 2x2x2 distinct source voxels, so each class is a
// 2x2x2 convolution of the low-resolution input with pre-summed weights (8/27 of the flops and loads).
#include <stdlib.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);
int conv3d_tc2_launch(const TcParams& c, int B, int epi, const void* residual, const float* gamma, const float* beta,
                      float eps, int ld_out, cudaStream_t st);       // conv_tc2.cu: persistent variant
static bool use_persistent_tc() {
    // Round 2 (tools/layer_profile.py, batch 8 of 128^3): the four 1x1x1 gate convolutions of levels 2-3 take 46-85 us on
    // the persistent kernel (<= 2 tiles per CTA: its per-CTA pipeline set-up is never amortised) and 20-31 us on this one
    // -> the persistent variant is opt-in (LTU_TC_PERSISTENT=1) and keeps serving ltu_linear_tc.
    static const bool v = [] { const char* e = getenv("LTU_TC_PERSISTENT"); return e && e[0] == '1'; }();
    return v;
}

__global__ void __launch_bounds__(kTcThreads, 4)
conv3d_tc_kernel(const TcParams p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: instnorm_finalize / _apply may be scheduled under this grid's tail (no-op without a PDL dependent)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // carve: [stages][A 16 KB][B Cout*128 B] | barriers | tmem slot | stat staging
    unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int a_bytes = kTcM * 128;
    const int b_bytes = p.Cout * 128;
    const int stage_bytes = a_bytes + b_bytes;
    unsigned char* tail = smem + (size_t)p.stages * stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);            // [stages]
    uint64_t* empty_bar = full_bar + p.stages;                          // [stages]
    uint64_t* done_bar = empty_bar + p.stages;                          // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);
    float* sred = reinterpret_cast<float*>(tmem_slot + 2);              // [4][Cout][2]
    int* toff = reinterpret_cast<int*>(sred + 4 * p.Cout * 2);          // [32] voxel offset of every tap

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.y;
    // rows of the implicit GEMM: output voxels, or (fold) low-resolution voxels of one parity class
    const int64_t Vo = p.fold ? (int64_t)p.Hi * p.Wi * p.Di : (int64_t)p.Ho * p.Wo * p.Do;
    const int64_t vox0 = (int64_t)blockIdx.x * kTcM;
    const int nkb = p.Kpad / kTcBK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            mbar_init(smem_u32(full_bar + s), kTcProducers);
            mbar_init(smem_u32(empty_bar + s), 1);
        }
        mbar_init(smem_u32(done_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // parity class of this CTA in fold mode: output voxel (2a+pa, 2b+pb, 2c+pc)
    const int pa = p.fold ? (blockIdx.z >> 2) & 1 : 0, pb = p.fold ? (blockIdx.z >> 1) & 1 : 0, pc = p.fold ? blockIdx.z & 1 : 0;
    if (threadIdx.x < 32) {
        const int t = threadIdx.x;
        int off = 0;
        if (t < p.ntaps) {
            int dh, dw, dd;
            if (p.fold) { dh = ((t >> 2) & 1) - 1 + pa; dw = ((t >> 1) & 1) - 1 + pb; dd = (t & 1) - 1 + pc; }
            else if (p.ks == 3) { dh = t / 9 - 1; dw = (t / 3) % 3 - 1; dd = t % 3 - 1; }
            else { dh = dw = dd = 0; }
            off = (dh * p.Wi + dw) * p.Di + dd;
        }
        toff[t] = off;
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
        // =========================== producers: im2col gather ===========================
        // Thread t owns 16-byte chunk column j = t & 7 of rows (t >> 3) + 16*i, i < 8: the eight lanes
        // of a row fetch its 128 contiguous bytes (full 32-byte sectors), and the tap / channel decode
        // of (k-block, j) is shared by the thread's eight rows.
        const int jc = threadIdx.x & 7, rbase = threadIdx.x >> 3;
        const int Cin = p.C0 + p.C1;
        int ctrv[8];                                   // source voxel under the centre tap (whole batch index)
        uint32_t vmask[8];                             // valid-tap bits of the row
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int rr = rbase + 16 * i;
            int64_t rid = vox0 + rr;
            const bool ok = rid < Vo;
            if (!ok) rid = 0;
            int ch, cw, cd;
            uint32_t m = 0;
            if (p.fold) {
                cd = (int)(rid % p.Di);
                const int64_t t2 = rid / p.Di;
                cw = (int)(t2 % p.Wi); ch = (int)(t2 / p.Wi);
                // per-axis validity of source offsets (th-1+pa), th in {0,1}
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
            } else {
                const int od = (int)(rid % p.Do);
                const int64_t t2 = rid / p.Do;
                const int ow = (int)(t2 % p.Wo), oh = (int)(t2 / p.Wo);
                ch = oh * p.sh; cw = ow * p.sw; cd = od * p.sd;
                if (p.ks == 3) {
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
                } else {
                    m = 1u;
                }
            }
            vmask[i] = ok ? m : 0u;
            ctrv[i] = (int)((int64_t)b * p.Hi * p.Wi * p.Di + ((int64_t)ch * p.Wi + cw) * p.Di + cd);
        }
        const int b_iters = p.Cout / 16;              // 8*Cout chunks over 128 threads
        const bf16* wbase = p.weight + (p.fold ? (int64_t)blockIdx.z * p.w_class_stride : 0);

        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % p.stages;
            const uint32_t round = kb / p.stages;
            mbar_wait(smem_u32(empty_bar + stage), (round & 1) ^ 1);
            const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint32_t sb = sa + a_bytes;
            // ---- A tile: this thread's chunk column of 8 rows; tap >= ntaps (K padding) has a zero mask bit
            const int kk = kb * kTcBK + jc * 8;
            const int tap = kk >> p.log2cin;
            const int c = kk & (Cin - 1);
            const int tvo = toff[tap & 31];
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
            // ---- B tile: Cout rows x 8 chunks (weights are zero-padded to Kpad)
            for (int i = 0; i < b_iters; ++i) {
                const int q = threadIdx.x + i * kTcProducers;
                const int n = q >> 3, j = q & 7;
                cp16(sb + n * 128 + ((j ^ (n & 7)) << 4), wbase + (int64_t)n * p.Kpad + kb * kTcBK + j * 8, 16);
            }
            cp_async_arrive_noinc(smem_u32(full_bar + stage));
        }

        // epilogue row of this thread (TMEM lane = threadIdx.x)
        const int r = threadIdx.x;
        int64_t id = vox0 + r;
        const bool row_ok = id < Vo;
        if (!row_ok) id = 0;
        int64_t out_vox = id;
        if (p.fold) {
            const int c_ = (int)(id % p.Di);
            const int64_t t2 = id / p.Di;
            const int b_ = (int)(t2 % p.Wi), a_ = (int)(t2 / p.Wi);
            out_vox = ((int64_t)(2 * a_ + pa) * p.Wo + (2 * b_ + pb)) * p.Do + (2 * c_ + pc);
        }

        // =========================== epilogue ===========================
        mbar_wait(smem_u32(done_bar), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t vrow = (int64_t)b * (int64_t)p.Ho * p.Wo * p.Do + out_vox;
        const int64_t out_row = vrow * p.Cstore;
        const int ctot = p.Cstore + p.naux;                       // main channels, then the fused fp32 head
        for (int c0 = 0; c0 < ctot; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            int ncol = p.Cstore - c0;                              // main columns of this chunk
            ncol = ncol < 0 ? 0 : (ncol > 32 ? 32 : ncol);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float bsv = (p.bias != nullptr && c0 + i < ctot) ? __ldg(p.bias + c0 + i) : 0.f;
                v[i] += bsv;
            }
            if (p.naux > 0 && row_ok) {                            // auxiliary head: unrounded fp32 logits
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int ca = c0 + i - p.Cstore;
                    if (ca >= 0 && ca < p.naux) p.aux[vrow * p.naux + ca] = v[i];
                }
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float o = v[i];
                // bf16 output: round here so that the statistics describe the stored values
                if (!p.out_f32) o = __bfloat162float(__float2bfloat16_rn(o));
                v[i] = (row_ok && i < ncol) ? o : 0.f;
            }
            if (ncol == 0) continue;
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
                float s = transpose_reduce32(v, lane);
                float q = transpose_reduce32(sq, lane);
                if (lane < ncol) {
                    sred[((warp * p.Cstore) + c0 + lane) * 2] = s;
                    sred[((warp * p.Cstore) + c0 + lane) * 2 + 1] = q;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        if (p.partials != nullptr) {
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int c = threadIdx.x; c < p.Cstore; c += kTcProducers) {
                float s = 0.f, q = 0.f;
#pragma unroll
                for (int w = 0; w < 4; ++w) { s += sred[(w * p.Cstore + c) * 2]; q += sred[(w * p.Cstore + c) * 2 + 1]; }
                float* dst = p.partials + (((int64_t)b * p.tiles + (int64_t)blockIdx.z * gridDim.x + blockIdx.x) * p.Cstore + c) * 2;
                dst[0] = s; dst[1] = q;
            }
        }
    } else {
        // =========================== MMA issuer ===========================
        // kind::f16 instruction descriptor: D=F32 (bit 4), A=B=BF16 (bits 7, 10), K-major A and B,
        // N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.Cout >> 3) << 17) | ((uint32_t)(kTcM >> 4) << 24);
        for (int kb = 0; kb < nkb; ++kb) {
            const int stage = kb % p.stages;
            const uint32_t round = kb / p.stages;
            mbar_wait(smem_u32(full_bar + stage), round & 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // cp.async (generic proxy) -> UMMA (async proxy)
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0) {
                const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                const uint64_t adesc = make_desc(sa), bdesc = make_desc(sa + a_bytes);
#pragma unroll
                for (int k = 0; k < kTcBK / 16; ++k)                        // +32 B per K=16 inside the swizzle row
                    umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                umma_commit(smem_u32(empty_bar + stage));                   // frees the stage when the MMAs retire
                if (kb == nkb - 1) umma_commit(smem_u32(done_bar));
            }
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

static inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }
static inline int ilog2(int v) { int l = 0; while ((1 << l) < v) ++l; return l; }

}  // namespace ltu

using namespace ltu;

extern "C" int ltu_conv3d_tc_supported(int C0, int C1, int Cout, int ksize, int pad) {
    const int Cin = C0 + C1;
    if (!((ksize == 3 && pad == 1) || (ksize == 1 && pad == 0))) return 0;
    if (!is_pow2(Cin) || Cin < 8 || Cin > 1024) return 0;
    if (C0 % 8 != 0 || C1 % 8 != 0) return 0;
    if (Cout < 1 || Cout > 256) return 0;
    return 1;
}

extern "C" int ltu_conv3d_tc_tiles(int64_t out_voxels, int up2) {
    if (up2) return 8 * (int)ceil_div64(out_voxels / 8, kTcM);     // 8 parity classes of the low-res grid
    return (int)ceil_div64(out_voxels, kTcM);
}

// ksize 3 -> 27 taps, 1 -> 1 tap, 2 -> the 8 folded taps of (nearest x2 upsample o 3x3x3 conv)
extern "C" int ltu_conv3d_tc_kpad(int Cin, int ksize) {
    return (int)ceil_div64((int64_t)ksize * ksize * ksize * Cin, kTcBK) * kTcBK;
}

extern "C" int ltu_conv3d_tc(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di, int up2,
                             int ksize, int sh, int sw, int sd, int pad, const void* weight_bf16, const float* bias,
                             int Cout, void* out, int out_f32, int Ho, int Wo, int Do, float* partials, int n_aux,
                             float* aux_out, ltu_stream_t stream) {
    LTU_ARG_CHECK(in0 && weight_bf16 && out, "conv3d_tc: null pointer");
    LTU_ARG_CHECK(n_aux >= 0 && n_aux <= 16 && (n_aux == 0 || (aux_out && ksize == 3 && !out_f32)), "conv3d_tc: bad auxiliary head");
    LTU_ARG_CHECK(ltu_conv3d_tc_supported(C0, C1, Cout + n_aux, ksize, pad), "conv3d_tc: unsupported C0=%d C1=%d Cout=%d k=%d pad=%d", C0, C1, Cout + n_aux, ksize, pad);
    LTU_ARG_CHECK((in1 != nullptr) == (C1 > 0), "conv3d_tc: in1/C1 mismatch");
    LTU_ARG_CHECK(B > 0 && B <= 65535 && Hi > 0 && Wi > 0 && Di > 0, "conv3d_tc: bad shape");
    LTU_ARG_CHECK(sh >= 1 && sh <= 2 && sw >= 1 && sw <= 2 && sd >= 1 && sd <= 2, "conv3d_tc: stride must be 1 or 2");
    LTU_ARG_CHECK(!up2 || (sh == 1 && sw == 1 && sd == 1 && ksize == 3), "conv3d_tc: up2 needs a stride-1 3x3x3 conv");
    const int He = up2 ? 2 * Hi : Hi, We = up2 ? 2 * Wi : Wi, De = up2 ? 2 * Di : Di;
    LTU_ARG_CHECK(Ho == (He + 2 * pad - ksize) / sh + 1 && Wo == (We + 2 * pad - ksize) / sw + 1 &&
                  Do == (De + 2 * pad - ksize) / sd + 1, "conv3d_tc: output size does not match the geometry");
    LTU_ARG_CHECK(((uintptr_t)in0 & 15) == 0 && ((uintptr_t)in1 & 15) == 0 && ((uintptr_t)weight_bf16 & 15) == 0 &&
                  ((uintptr_t)out & 15) == 0, "conv3d_tc: pointers must be 16-byte aligned");
    LTU_ARG_CHECK(out_f32 || Cout % 8 == 0, "conv3d_tc: bf16 output needs Cout %% 8 == 0");
    LTU_ARG_CHECK((int64_t)Hi * Wi * Di * B < (int64_t)1 << 31, "conv3d_tc: input too large for 32-bit voxel offsets");
    TcParams p;
    const int Cin = C0 + C1;
    p.in0 = (const bf16*)in0; p.in1 = (const bf16*)in1; p.C0 = C0; p.C1 = C1; p.log2cin = ilog2(Cin);
    p.Hi = Hi; p.Wi = Wi; p.Di = Di; p.up2 = up2; p.ks = ksize; p.pad = pad; p.sh = sh; p.sw = sw; p.sd = sd;
    p.weight = (const bf16*)weight_bf16;
    p.fold = up2 ? 1 : 0;
    p.ntaps = up2 ? 8 : ksize * ksize * ksize;
    p.Ktot = p.ntaps * Cin;
    p.Kpad = ltu_conv3d_tc_kpad(Cin, up2 ? 2 : ksize);
    p.bias = bias; p.Cstore = Cout; p.naux = n_aux; p.aux = aux_out;
    p.Cout = (Cout + n_aux + 15) / 16 * 16; p.out = out; p.out_f32 = out_f32;
    p.w_class_stride = (int64_t)p.Cout * p.Kpad;
    p.Ho = Ho; p.Wo = Wo; p.Do = Do;
    p.partials = partials; p.tiles = ltu_conv3d_tc_tiles((int64_t)Ho * Wo * Do, up2);
    // The persistent kernel wins where a tile has very little K (1x1x1 convs: one k-block per tile, the
    // per-CTA set-up dominated); deep-K layers are gather-bound and prefer two resident CTAs of this
    // kernel (2 x 128 producer threads per SM).  Measured on B200: profiles/r1_conv_variants.md
    if (use_persistent_tc() && ksize == 1) { p.stages = 0; p.tmem_cols = 0; return conv3d_tc2_launch(p, B, 0, nullptr, nullptr, nullptr, 0.f, 0, (cudaStream_t)stream); }
    int cols = 32; while (cols < p.Cout) cols <<= 1;
    p.tmem_cols = cols;
    const int stage_bytes = kTcM * 128 + p.Cout * 128;
    const int nkb = p.Kpad / kTcBK;
    // The kernel is bound by the im2col gather (address generation + L2 latency), i.e. by the number of
    // resident producer threads: measured on B200, 4 CTAs/SM (<= 54 KB of pipeline each, 96 registers)
    // beats 2 CTAs/SM with deeper pipelines by 1.2-1.5x on every layer (profiles/r1_conv_variants.md).
    int budget = 54 * 1024;
    {   // tuning knob (KB of pipeline smem per CTA => CTAs per SM); unset in production
        static const int knob = [] { const char* e = getenv("LTU_TC_SMEM_KB"); return e ? atoi(e) : 0; }();
        if (knob > 0) budget = knob * 1024;
    }
    int stages = budget / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages > nkb) stages = nkb;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + (size_t)4 * p.Cout * 2 * 4 + 32 * 4;
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(conv3d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured_dev = dev;
    }
    const unsigned gx = (unsigned)(up2 ? p.tiles / 8 : p.tiles);
    conv3d_tc_kernel<<<dim3(gx, B, up2 ? 8 : 1), kTcThreads, smem, (cudaStream_t)stream>>>(p);
    LTU_LAUNCH_CHECK("conv3d_tc");
    count_launch(1);
    return LTU_OK;
}
