// Family 2, third generation: TMA-fed shared-memory HALO + tcgen05 implicit GEMM for the stride-1 3x3x3
// convolutions with >= 64 input channels per tensor, including the folded (nearest x2 upsample o 3x3x3)
// up_embed convolutions that carry 60 % of the model's convolution flops (model/Unet_3Dblock.py:419-429).
//
// conv3d_tc (conv_tc.cu) gathers an im2col tile PER TAP: every input voxel is fetched 27 times from L2 and the
// whole weight tensor once per 128 output voxels; on the big layers it sits at the L2 bandwidth, not at the
// tensor pipe (profiles/r1_ncu_full.md).  Here a CTA owns a 4 x 16 x 8 block of output voxels:
//
//   * ONE cp.async.bulk.tensor.5d (SWIZZLE_128B, out-of-range = zero = padding) brings the 6 x 18 x 10 voxel halo
//     of one 64-channel slice into shared memory (1080 rows of 128 B = 135 KB): each input voxel is fetched 2.1
//     times instead of 27;
//   * the A operand of filter tap (kh,kw,kd) for the 1 x 16 x 8 sub-block v is NOT copied: it is the halo itself
//     seen through a UMMA descriptor whose start is shifted by ((v+kh)*18 + kw)*10 + kd rows and whose 8-row
//     groups are 10 rows apart.  tcgen05 applies the 128-byte swizzle to the final shared-memory address, so
//     any 128-byte row offset / group stride is legal with base_offset 0 (measured: tools/umma_probe.py);
//   * the weight tile of a (tap, slice) is one TMA box [Cout x 64] and serves all four sub-blocks (4 x fewer
//     weight bytes per output voxel); four TMEM accumulators [128 x Cout] live side by side;
//   * folded mode: the SAME low-resolution halo feeds all 8 output-parity classes (2x2x2 taps each, offsets
//     tap + parity in {0,1,2}); 512 / (4 Cout) classes are accumulated per pass.
//
//   warp 0 lane 0: halo TMA | warp 2 lane 0: weight TMA ring | warp 1 lane 0: tcgen05.mma issue
//   warps 4-11: epilogue (tcgen05.ld -> +bias -> bf16 -> 16-byte stores, InstanceNorm partial sums with the
//   fixed-order transpose butterfly) -- same results contract as conv3d_tc.
// The tile axes (8, 16, 4) are mapped to (H, W, D) in the order that wastes the fewest padded voxels (e.g. the
// 8 x 8 x 64 decoder level uses D as the 16-axis); the permutation lives in the tensor map strides only.
#include <cuda.h>
#include <stdlib.h>

#include "tc_common.cuh"

namespace ltu {

void count_launch(int n = 1);
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);

constexpr int kT3Threads = 384;
constexpr int kT3Stages = 12;          // weight ring: deep enough to cover the TMA latency when a tile feeds few MMAs

struct T3Params {
    bf16* out; const float* bias; float* partials;
    float* aux; int naux;            // fused auxiliary head: naux extra output channels, unrounded fp32 [B][V][naux]
    int N, Cstore;                   // UMMA N (multiple of 32, <= 256), stored main channels (<= 128)
    int nchunk0, nchunk1, Cin;       // 64-channel slices of in0 / in1, total input channels
    int fold, ntaps, cpp, npass;     // folded mode, taps per class, classes per pass, passes
    int dim[3];                      // H, W, D of the (low-resolution) input = tile space
    int ax8, ax16, axT;              // which of H(0), W(1), D(2) carries the tile's 8-, 16- and 4-axis
    int n8, n16, nT;                 // tiles along them
    int Ho, Wo, Do;
    int tiles_per_sample;            // partial-sum slots per sample = n8*n16*nT * npass
    int TH;                          // tile extent along axT: 1, 2 or 4 sub-blocks of 1 x 16 x 8 voxels
    int nbuf, nstage;                // halo buffers (1 or 2), weight ring stages
    int nacc;                        // accumulator sets in tensor memory: 2 (256 columns each) when TH*cpp*N <= 256, else 1
    int tiles, total_tiles;          // tiles per sample, tiles over the batch
    uint32_t halo_bytes, halo_stride, off_b;   // bytes of one halo, distance between halo buffers, offset of the weight ring
    // super-voxel mode (ltu_conv3d_tc3_masked): bit ks of ksmask[t] = the 16-channel K block ks of filter tap t holds
    // non-zero weights (block-Toeplitz repacking, see the entry point); all-zero blocks are not issued.  0xF otherwise.
    uint8_t ksmask[27];
};

struct T3Tail {
    uint64_t halo_full[2], halo_empty[2], b_full[kT3Stages], b_empty[kT3Stages], acc_done[2], acc_free[2];
    uint32_t tmem_slot, pad_;
    float sred[8][128][2];
};

template <bool kMasked>
__global__ void __launch_bounds__(kT3Threads, 1)
conv3d_tc3_kernel_t(const __grid_constant__ CUtensorMap tm_in0, const __grid_constant__ CUtensorMap tm_in1,
                  const __grid_constant__ CUtensorMap tm_w, const T3Params p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: instnorm_finalize / _apply may be scheduled under this grid's tail (no-op without a PDL dependent)
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    const uint32_t b_bytes = (uint32_t)p.N * 128u;
    T3Tail* tail = reinterpret_cast<T3Tail*>(smem + p.off_b + p.nstage * b_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // persistent: this CTA walks tiles T = blockIdx.x, blockIdx.x + gridDim.x, ...; a work ITEM is one (tile, pass)
    const int n_my = (p.total_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto tile_coords = [&](int i, int& b, int& tix, int& x8, int& x16, int& xT) {
        const int T = (int)blockIdx.x + i * (int)gridDim.x;
        b = T / p.tiles; tix = T % p.tiles;
        int r = tix;
        x8 = (r % p.n8) * 8; r /= p.n8;
        x16 = (r % p.n16) * 16; r /= p.n16;
        xT = r * p.TH;
    };
    const int nchunks = p.nchunk0 + p.nchunk1;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tail->halo_full[s]), 1); mbar_init(smem_u32(&tail->halo_empty[s]), 1); }
        for (int s = 0; s < kT3Stages; ++s) { mbar_init(smem_u32(&tail->b_full[s]), 1); mbar_init(smem_u32(&tail->b_empty[s]), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&tail->acc_done[s]), 1); mbar_init(smem_u32(&tail->acc_free[s]), 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
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
        // =========================== halo producer ===========================
        if (lane == 0) {
            uint32_t n = 0;
            for (int i = 0; i < n_my; ++i) {
              int b, tix, x8, x16, xT;
              tile_coords(i, b, tix, x8, x16, xT);
              for (int pass = 0; pass < p.npass; ++pass)
                for (int j = 0; j < nchunks; ++j, ++n) {
                    const int hb = n % p.nbuf;
                    mbar_wait(smem_u32(&tail->halo_empty[hb]), ((n / p.nbuf) & 1) ^ 1);
                    mbar_expect_tx(smem_u32(&tail->halo_full[hb]), p.halo_bytes);
                    const bool first = j < p.nchunk0;
                    tma_load_5d(sbase + hb * p.halo_stride, first ? (const void*)&tm_in0 : (const void*)&tm_in1,
                                (first ? j : j - p.nchunk0) * 64, x8 - 1, x16 - 1, xT - 1, b, smem_u32(&tail->halo_full[hb]));
                }
            }
        }
    } else if (warp == 2) {
        // =========================== weight producer ===========================
        if (lane == 0) {
            uint32_t n = 0;
            for (int i = 0; i < n_my; ++i)
              for (int pass = 0; pass < p.npass; ++pass)
                for (int j = 0; j < nchunks; ++j)
                    for (int zl = 0; zl < p.cpp; ++zl)
                        for (int t = 0; t < p.ntaps; ++t, ++n) {
                            const int s = n % p.nstage;
                            mbar_wait(smem_u32(&tail->b_empty[s]), ((n / p.nstage) & 1) ^ 1);
                            mbar_expect_tx(smem_u32(&tail->b_full[s]), b_bytes);
                            tma_load_2d(sbase + p.off_b + s * b_bytes, &tm_w, t * p.Cin + j * 64, (pass * p.cpp + zl) * p.N,
                                        smem_u32(&tail->b_full[s]));
                        }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        const uint32_t idesc = umma_idesc_bf16(128, p.N);
        const int first_ks = kMasked ? __ffs((int)p.ksmask[0]) - 1 : 0;             // first issued MMA of an item overwrites
        uint32_t nb = 0, nh = 0, it = 0;
        for (int i = 0; i < n_my; ++i)
        for (int pass = 0; pass < p.npass; ++pass, ++it) {
            const int ab = it % p.nacc;                                             // accumulator set of this item
            const uint32_t acc_base = tmem_base + (uint32_t)(ab * 256);
            mbar_wait(smem_u32(&tail->acc_free[ab]), ((it / p.nacc) & 1) ^ 1);      // the epilogue drained its previous use
            tc_fence_after();
            for (int j = 0; j < nchunks; ++j, ++nh) {
                const int hb = nh % p.nbuf;
                const uint32_t hbase = sbase + hb * p.halo_stride;
                mbar_wait(smem_u32(&tail->halo_full[hb]), (nh / p.nbuf) & 1);
                tc_fence_after();
                for (int zl = 0; zl < p.cpp; ++zl) {
                    const int z = pass * p.cpp + zl;
                    for (int t = 0; t < p.ntaps; ++t, ++nb) {
                        const int s = nb % p.nstage;
                        mbar_wait(smem_u32(&tail->b_full[s]), (nb / p.nstage) & 1);
                        tc_fence_after();
                        {
                            int k[3];                                   // halo offset of this tap along H, W, D
                            if (p.fold) {
                                k[0] = ((t >> 2) & 1) + ((z >> 2) & 1); k[1] = ((t >> 1) & 1) + ((z >> 1) & 1); k[2] = (t & 1) + (z & 1);
                            } else {
                                k[0] = t / 9; k[1] = (t / 3) % 3; k[2] = t % 3;
                            }
                            auto pick = [&](int ax) { return ax == 0 ? k[0] : (ax == 1 ? k[1] : k[2]); };
                            const int row0 = (pick(p.axT) * 18 + pick(p.ax16)) * 10 + pick(p.ax8);
                            const uint32_t kmask = kMasked ? p.ksmask[t] : 0xFu;
                            const uint64_t bdesc = make_desc(sbase + p.off_b + s * b_bytes);
#pragma unroll
                            for (int v = 0; v < 4; ++v) {
                                if (v >= p.TH) break;
                                // A = halo rows ((v + oT)*18 + o16 + u)*10 + o8 + d, u = 0..15 (groups, 10 rows apart), d = 0..7
                                const uint32_t start = hbase + (uint32_t)(row0 + v * 180) * 128u;
                                const uint64_t adesc = (uint64_t)((start & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1280 >> 4) << 32) |
                                                       ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
                                const uint32_t tacc = acc_base + (uint32_t)((zl * p.TH + v) * p.N);
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks) {
                                    if (kMasked) {      // super-voxel form only: the plain kernel keeps its branch-free issue loop
                                        if (!((kmask >> ks) & 1u)) continue;
                                        umma_bf16_elect(tacc, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc,
                                                        (uint32_t)((j | t) != 0 || ks != first_ks));
                                    } else {
                                        umma_bf16_elect(tacc, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc, (j | t | ks) != 0);
                                    }
                                }
                            }
                            umma_commit_elect(smem_u32(&tail->b_empty[s]));
                        }
                    }
                }
                umma_commit_elect(smem_u32(&tail->halo_empty[hb]));
            }
            umma_commit_elect(smem_u32(&tail->acc_done[ab]));
        }
    } else if (warp >= 4) {
        // =========================== epilogue ===========================
        const int e = warp - 4;
        const int q = e & 3;                       // TMEM lane quarter (== warp % 4)
        const int set = e >> 2;                    // accumulators set, set+2, ...
        const int r = q * 32 + lane;               // tile row = u*8 + d
        const int i16 = r >> 3, i8 = r & 7;
        const uint32_t lane_off = (uint32_t)(q * 32) << 16;
        const int et = threadIdx.x - 128;          // 0..255
        const int nacc = p.cpp * p.TH;
        uint32_t it = 0;
        for (int i = 0; i < n_my; ++i) {
          int b, tix, x8, x16, xT;
          tile_coords(i, b, tix, x8, x16, xT);
          for (int pass = 0; pass < p.npass; ++pass, ++it) {
            const int ab = it % p.nacc;
            const uint32_t acc_base = tmem_base + (uint32_t)(ab * 256);
            mbar_wait(smem_u32(&tail->acc_done[ab]), (it / p.nacc) & 1);
            tc_fence_after();
            float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};       // column sums of this warp, 32 columns per slot
            for (int a = set; a < nacc; a += 2) {
                const int zl = a / p.TH, v = a % p.TH;
                const int z = pass * p.cpp + zl;
                const int q8 = x8 + i8, q16 = x16 + i16, qT = xT + v;
                auto coord = [&](int ax) { return p.ax8 == ax ? q8 : (p.ax16 == ax ? q16 : qT); };
                int oh = coord(0), ow = coord(1), od = coord(2);
                const bool ok = oh < p.dim[0] && ow < p.dim[1] && od < p.dim[2];
                if (p.fold) { oh = 2 * oh + ((z >> 2) & 1); ow = 2 * ow + ((z >> 1) & 1); od = 2 * od + (z & 1); }
                bf16* dst = p.out + ((((int64_t)b * p.Ho + oh) * p.Wo + ow) * p.Do + od) * p.Cstore;
                const int64_t vrow = (((int64_t)b * p.Ho + oh) * p.Wo + ow) * p.Do + od;
#pragma unroll
                for (int cb = 0; cb < 8; ++cb) {
                    const int c0 = cb * 32;
                    if (c0 >= p.N) break;
                    float vv[32];
                    tmem_ld32(acc_base + lane_off + (uint32_t)(a * p.N + c0), vv);
                    int ncol = p.Cstore - c0;                              // main columns of this block
                    ncol = ncol < 0 ? 0 : (ncol > 32 ? 32 : ncol);
                    const int ctot = p.Cstore + p.naux;
#pragma unroll
                    for (int i = 0; i < 32; ++i) vv[i] += (p.bias != nullptr && c0 + i < ctot) ? __ldg(p.bias + c0 + i) : 0.f;
                    if (p.naux > 0 && ok && c0 + 32 > p.Cstore) {          // auxiliary head: unrounded fp32 logits
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int ca = c0 + i - p.Cstore;
                            if (ca >= 0 && ca < p.naux) p.aux[vrow * p.naux + ca] = vv[i];
                        }
                    }
                    if (ncol == 0) continue;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float o = __bfloat162float(__float2bfloat16_rn(vv[i]));   // statistics describe the stored values
                        vv[i] = (ok && i < ncol) ? o : 0.f;
                    }
                    if (ok) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (g * 8 < ncol) {
                                uint4 o;
                                o.x = pack_bf16x2(vv[g * 8 + 0], vv[g * 8 + 1]); o.y = pack_bf16x2(vv[g * 8 + 2], vv[g * 8 + 3]);
                                o.z = pack_bf16x2(vv[g * 8 + 4], vv[g * 8 + 5]); o.w = pack_bf16x2(vv[g * 8 + 6], vv[g * 8 + 7]);
                                *reinterpret_cast<uint4*>(dst + c0 + g * 8) = o;
                            }
                        }
                    }
                    if (p.partials != nullptr && cb < 4) {
                        float sq[32];
#pragma unroll
                        for (int i = 0; i < 32; ++i) sq[i] = vv[i] * vv[i];
                        cs[cb & 3] += transpose_reduce32(vv, lane);
                        cq[cb & 3] += transpose_reduce32(sq, lane);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(smem_u32(&tail->acc_free[ab]));
            if (p.partials != nullptr) {
#pragma unroll
                for (int cb = 0; cb < 4; ++cb) {
                    if (cb * 32 < p.Cstore) {
                        tail->sred[e][cb * 32 + lane][0] = cs[cb];
                        tail->sred[e][cb * 32 + lane][1] = cq[cb];
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (et < p.Cstore) {
                    float s = 0.f, qq = 0.f;
#pragma unroll
                    for (int w = 0; w < 8; ++w) { s += tail->sred[w][et][0]; qq += tail->sred[w][et][1]; }
                    float* dstp = p.partials + (((int64_t)b * p.tiles_per_sample + (int64_t)tix * p.npass + pass) * p.Cstore + et) * 2;
                    dstp[0] = s; dstp[1] = qq;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
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

typedef CUresult (*EncodeTiledFn5)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 5-D map over a channels-last activation [B][H][W][D][C]: dims (C, ax8, ax16, axT, B), box (64, 10, 18, 6, 1)
static int make_tmap_halo(CUtensorMap* map, const void* base, int C, const int dim[3], int B, int ax8, int ax16, int axT, int TH) {
    static EncodeTiledFn5 fn = [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
            qres != cudaDriverEntryPointSuccess)
            ptr = nullptr;
        return (EncodeTiledFn5)ptr;
    }();
    if (!fn) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return LTU_ERR_ARG; }
    const uint64_t sH = (uint64_t)dim[1] * dim[2] * C * 2, sW = (uint64_t)dim[2] * C * 2, sD = (uint64_t)C * 2;
    const uint64_t st[3] = {sH, sW, sD};
    const cuuint64_t gdim[5] = {(cuuint64_t)C, (cuuint64_t)dim[ax8], (cuuint64_t)dim[ax16], (cuuint64_t)dim[axT], (cuuint64_t)B};
    const cuuint64_t gstride[4] = {st[ax8], st[ax16], st[axT], (cuuint64_t)dim[0] * sH};
    const cuuint32_t box[5] = {64, 10, 18, (cuuint32_t)(TH + 2), 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled (5-D halo) failed (CUresult %d)", (int)r); return LTU_ERR_ARG; }
    return LTU_OK;
}

// tile-axis permutation that pads the fewest voxels: returns tiles per sample and fills ax8/ax16/axT, n8/n16/nT
static int64_t choose_axes(const int dim[3], int TH, int& ax8, int& ax16, int& axT, int& n8, int& n16, int& nT) {
    static const int perms[6][3] = {{2, 1, 0}, {2, 0, 1}, {1, 2, 0}, {0, 2, 1}, {1, 0, 2}, {0, 1, 2}};   // (ax8, ax16, axT)
    int64_t best = -1;
    for (int i = 0; i < 6; ++i) {
        const int a8 = perms[i][0], a16 = perms[i][1], aT = perms[i][2];
        const int m8 = (dim[a8] + 7) / 8, m16 = (dim[a16] + 15) / 16, mT = (dim[aT] + TH - 1) / TH;
        const int64_t tiles = (int64_t)m8 * m16 * mT;
        if (best < 0 || tiles < best) { best = tiles; ax8 = a8; ax16 = a16; axT = aT; n8 = m8; n16 = m16; nT = mT; }
    }
    return best;
}

// Tile height and classes per pass (functions of the per-sample shape only).
static void choose_tile(const int dim[3], int B, int N, int up2, int& TH, int& cpp) {
    static const int forced = [] { const char* e = getenv("LTU_TC3_TH"); return e ? atoi(e) : 0; }();   // tuning knob
    if (forced == 1 || forced == 2 || forced == 4) {
        TH = forced;
        while (TH > 1 && TH * N > 512) TH >>= 1;
        cpp = 1;
        if (up2) {
            static const int fcpp = [] { const char* e = getenv("LTU_TC3_CPP"); return e ? atoi(e) : 0; }();
            cpp = 512 / (TH * N); if (cpp > 8) cpp = 8; while (8 % cpp) --cpp;
            if (fcpp > 0 && fcpp <= cpp) cpp = fcpp;
        }
        return;
    }
    if (up2) {
        // tallest tile; classes per pass so that two accumulator sets fit (the epilogue of a pass runs under the next pass)
        TH = 4;
        while (TH > 1 && TH * N > 256) TH >>= 1;
        cpp = 256 / (TH * N);
        if (cpp < 1) cpp = 1;
        if (cpp > 8) cpp = 8;
        while (8 % cpp) --cpp;
        return;
    }
    cpp = 1;
    const int lim = N * 4 <= 256 ? 256 : (N * 2 <= 256 ? 256 : 512);      // keep two accumulator sets when some tile height allows it
    for (TH = 4; TH > 1; TH >>= 1) {
        int a, b, c, d, e, f;
        // >= 16 tiles per sample (batch 8 then fills the GPU); a per-sample rule, never a function of the batch size, so that
        // a patch gives the same bits in any batch (the partial-sum grouping of the InstanceNorm statistics depends on TH)
        if (TH * N <= lim && choose_axes(dim, TH, a, b, c, d, e, f) >= 16) return;
    }
}

static bool tc3_enabled() {
    static const bool v = [] { const char* e = getenv("LTU_DISABLE_TC3"); return !(e && e[0] == '1'); }();
    return v;
}

}  // namespace ltu

using namespace ltu;

extern "C" int ltu_conv3d_tc3_supported(int C0, int C1, int Cout, int ksize, int sh, int sw, int sd, int pad, int up2,
                                        int out_f32, int n_aux) {
    if (!tc3_enabled()) return 0;
    if (ksize != 3 || pad != 1 || sh != 1 || sw != 1 || sd != 1 || out_f32 || n_aux < 0 || n_aux > 64) return 0;
    if (C0 < 64 || C0 % 64 != 0 || C1 % 64 != 0) return 0;
    const int N = (Cout + n_aux + 31) / 32 * 32;
    if (Cout < 0 || Cout % 8 != 0 || Cout > 128 || N > 256 || N < 32) return 0;
    // folded layers: measured faster than the im2col kernel only for narrow outputs (b1.up_embed, Cout 32)
    if (up2 && (N > 32 || n_aux)) return 0;
    return 1;
}

// partial-sum slots per sample written by ltu_conv3d_tc3 for an input of H x W x D voxels (low resolution if up2)
extern "C" int ltu_conv3d_tc3_tiles(int B, int Hi, int Wi, int Di, int Cout, int n_aux, int up2) {
    const int dim[3] = {Hi, Wi, Di};
    const int N = (Cout + n_aux + 31) / 32 * 32;
    int TH, cpp, a, b, c, d, e, f;
    choose_tile(dim, B, N, up2, TH, cpp);
    const int64_t tiles = choose_axes(dim, TH, a, b, c, d, e, f);
    return (int)(tiles * (up2 ? 8 / cpp : 1));
}

// Same contract as ltu_conv3d_tc for the shapes ltu_conv3d_tc3_supported accepts.  weight_bf16: the ltu_conv3d_tc
// packing with rows padded to a multiple of 32 ([rows32][Kpad], or [8][rows32][Kpad] folded).
static int conv3d_tc3_impl(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di, int up2,
                           const void* weight_bf16, int weight_rows, int Kpad, const float* bias, int Cout, void* out,
                           float* partials, int n_aux, float* aux_out, const uint8_t* ks_mask, ltu_stream_t stream);

extern "C" int ltu_conv3d_tc3(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di, int up2,
                              const void* weight_bf16, int weight_rows, int Kpad, const float* bias, int Cout, void* out,
                              float* partials, int n_aux, float* aux_out, ltu_stream_t stream) {
    return conv3d_tc3_impl(in0, C0, in1, C1, B, Hi, Wi, Di, up2, weight_bf16, weight_rows, Kpad, bias, Cout, out, partials,
                           n_aux, aux_out, nullptr, stream);
}

// Super-voxel form of the small-channel layers (Cin in {8, 16, 32} per input): g = 64 / Cin consecutive voxels along D
// are ONE row of 64 channels -- the same memory, viewed as [B][H][W][D/g][64] -- so the layer is a 3x3x3 convolution over
// super-voxels with 64 input channels per tensor and g * Cout output channels, whose weights are the block-Toeplitz
// repacking of the original taps (W'[(kh,kw,kg)][(p,c)][(delta,co)] = W[kh][kw][g(kg-1)+p-delta+1][c][co], zero outside
// 0..2).  It runs on this kernel unchanged: UMMA N = g * Cout (64 .. 128) instead of 16 .. 32, the A operand is read
// once per g output voxels, the output [B][H][W][D/g][g*Cout] IS the channels-last [B][H][W][D][Cout] tensor, and the
// auxiliary head / InstanceNorm columns come out in (delta, channel) order.  ks_mask[27] (host memory): bit ks of entry
// t = K block ks of tap t is not identically zero; the zero blocks (7 of 12 per (kh,kw) at g = 4) are skipped.
extern "C" int ltu_conv3d_tc3_masked(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di,
                                     const void* weight_bf16, int weight_rows, int Kpad, const float* bias, int Cout,
                                     void* out, float* partials, int n_aux, float* aux_out, const uint8_t* ks_mask,
                                     ltu_stream_t stream) {
    LTU_ARG_CHECK(ks_mask, "conv3d_tc3_masked: null mask");
    return conv3d_tc3_impl(in0, C0, in1, C1, B, Hi, Wi, Di, 0, weight_bf16, weight_rows, Kpad, bias, Cout, out, partials,
                           n_aux, aux_out, ks_mask, stream);
}

static int conv3d_tc3_impl(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di, int up2,
                           const void* weight_bf16, int weight_rows, int Kpad, const float* bias, int Cout, void* out,
                           float* partials, int n_aux, float* aux_out, const uint8_t* ks_mask, ltu_stream_t stream) {
    LTU_ARG_CHECK(in0 && weight_bf16 && out, "conv3d_tc3: null pointer");
    LTU_ARG_CHECK(ltu_conv3d_tc3_supported(C0, C1, Cout, 3, 1, 1, 1, 1, up2, 0, n_aux), "conv3d_tc3: unsupported C0=%d C1=%d Cout=%d n_aux=%d", C0, C1, Cout, n_aux);
    LTU_ARG_CHECK(n_aux == 0 || aux_out, "conv3d_tc3: auxiliary head without an output buffer");
    LTU_ARG_CHECK((in1 != nullptr) == (C1 > 0), "conv3d_tc3: in1/C1 mismatch");
    LTU_ARG_CHECK(B > 0 && B <= 65535 && Hi > 0 && Wi > 0 && Di > 0, "conv3d_tc3: bad shape");
    T3Params p;
    p.N = (Cout + n_aux + 31) / 32 * 32;
    p.aux = aux_out; p.naux = n_aux;
    LTU_ARG_CHECK(weight_rows == p.N, "conv3d_tc3: weight rows %d != %d (pad the packed weight to a multiple of 32 rows)", weight_rows, p.N);
    LTU_ARG_CHECK(((uintptr_t)in0 & 15) == 0 && ((uintptr_t)in1 & 15) == 0 && ((uintptr_t)weight_bf16 & 15) == 0 &&
                  ((uintptr_t)out & 15) == 0, "conv3d_tc3: pointers must be 16-byte aligned");
    p.out = (bf16*)out; p.bias = bias; p.partials = partials; p.Cstore = Cout;
    for (int t = 0; t < 27; ++t) {
        p.ksmask[t] = ks_mask ? (uint8_t)(ks_mask[t] & 0xF) : (uint8_t)0xF;
        LTU_ARG_CHECK(p.ksmask[t] != 0, "conv3d_tc3_masked: tap %d has an empty mask", t);
    }
    p.nchunk0 = C0 / 64; p.nchunk1 = C1 / 64; p.Cin = C0 + C1;
    p.fold = up2 ? 1 : 0; p.ntaps = up2 ? 8 : 27;
    p.dim[0] = Hi; p.dim[1] = Wi; p.dim[2] = Di;
    choose_tile(p.dim, B, p.N, up2, p.TH, p.cpp);
    p.npass = up2 ? 8 / p.cpp : 1;
    const int64_t tiles = choose_axes(p.dim, p.TH, p.ax8, p.ax16, p.axT, p.n8, p.n16, p.nT);
    LTU_ARG_CHECK(tiles < ((int64_t)1 << 31), "conv3d_tc3: too many tiles");
    p.Ho = up2 ? 2 * Hi : Hi; p.Wo = up2 ? 2 * Wi : Wi; p.Do = up2 ? 2 * Di : Di;
    p.tiles_per_sample = (int)tiles * p.npass;
    p.tiles = (int)tiles;
    LTU_ARG_CHECK(tiles * B < ((int64_t)1 << 31), "conv3d_tc3: too many tiles");
    p.total_tiles = (int)(tiles * B);
    p.nacc = (p.TH * p.cpp * p.N <= 256) ? 2 : 1;
    LTU_ARG_CHECK(Kpad >= p.ntaps * p.Cin, "conv3d_tc3: Kpad %d too small", Kpad);
    CUtensorMap t0, t1, tw;
    int rc;
    if ((rc = make_tmap_halo(&t0, in0, C0, p.dim, B, p.ax8, p.ax16, p.axT, p.TH)) != LTU_OK) return rc;
    if ((rc = make_tmap_halo(&t1, in1 ? in1 : in0, in1 ? C1 : C0, p.dim, B, p.ax8, p.ax16, p.axT, p.TH)) != LTU_OK) return rc;
    if ((rc = make_tmap_bf16_2d(&tw, weight_bf16, (uint64_t)(up2 ? 8 : 1) * p.N, (uint64_t)Kpad, (uint32_t)p.N)) != LTU_OK) return rc;
    // shared memory: halo buffer(s) | weight ring | barriers + statistics staging; two halos when they fit
    p.halo_bytes = (uint32_t)((p.TH + 2) * 180 * 128);
    p.halo_stride = (p.halo_bytes + 1023u) & ~1023u;
    const size_t budget = 227 * 1024 - 1024 - sizeof(T3Tail);
    const size_t b_bytes = (size_t)p.N * 128;
    p.nbuf = (2 * (size_t)p.halo_stride + 2 * b_bytes <= budget) ? 2 : 1;
    p.off_b = p.nbuf * p.halo_stride;
    size_t st = (budget - p.off_b) / b_bytes;
    p.nstage = st > (size_t)kT3Stages ? kT3Stages : (int)st;
    LTU_ARG_CHECK(p.nstage >= 2, "conv3d_tc3: shared memory budget");
    const size_t smem = 1024 + p.off_b + (size_t)p.nstage * b_bytes + sizeof(T3Tail);
    static thread_local int configured_dev = -1;
    int dev; cudaGetDevice(&dev);
    if (configured_dev != dev) {
        cudaFuncSetAttribute(conv3d_tc3_kernel_t<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaFuncSetAttribute(conv3d_tc3_kernel_t<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        configured_dev = dev;
    }
    int grid = sm_count();
    if (grid > p.total_tiles) grid = p.total_tiles;
    if (ks_mask) conv3d_tc3_kernel_t<true><<<grid, kT3Threads, smem, (cudaStream_t)stream>>>(t0, t1, tw, p);
    else conv3d_tc3_kernel_t<false><<<grid, kT3Threads, smem, (cudaStream_t)stream>>>(t0, t1, tw, p);
    LTU_LAUNCH_CHECK("conv3d_tc3");
    count_launch(1);
    return LTU_OK;
}
