// Family 2, bf16 path, small-channel layers: stride-1 3x3x3 convolution with Cin in {8,16,32} and
// Cout <= 32 at (near) full resolution -- stem, enc.block0/1 conv1, dec.block3, the finest mask head
// and final_block.  These layers are bandwidth-bound (AI 86-216 flop/B): an im2col
// pipeline re-reads every input voxel 27 times from L2 and drowns in address generation.  Here the
// input halo of a 3-D output tile is staged ONCE in shared memory (16-byte cp.async, zero fill =
// padding, XOR-swizzled rows) together with all weights, and im2col happens for free in the
// per-lane row addresses of ldmatrix: A fragments of mma.sync.m16n8k16 (bf16, fp32 accumulate) are
// 16 consecutive voxels along D x 16 channels of one tap.  B fragments (weights) are shared by four
// m-tiles per warp.  Epilogue: bias, bf16/fp32 store, per-CTA InstanceNorm partial sums (fixed order).
//
// tcgen05 is deliberately not used for these layers: UMMA wants the A operand as 128-byte K-major
// rows in shared memory, i.e. an explicit im2col copy (27x the smem traffic); ldmatrix gathers rows.
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

struct HaloParams {
    const bf16* in0; const bf16* in1;
    int C0, C1;
    int H, W, D;                 // input == output spatial size (stride 1, pad 1)
    const bf16* weight;          // [>=16 rows][wld] bf16, K index = tap*Cin + c (the ltu_conv3d_tc packing)
    int wld;
    const float* bias;
    int Cout;                    // main channels stored in `out` (and covered by the statistics)
    int naux; float* aux;        // fused fp32 head: channels [Cout, Cout+naux) go to aux [B][V][naux]
    void* out; int out_f32;
    float* partials; int tiles;  // CTAs per sample
    int tiles_w, tiles_d;
};

__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr) : "memory");
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// CIN: input channels; NT: n-tiles of 8 output channels (2 or 4); TH x TW x 32 output tile;
// KS: kernel size 3 (pad 1) or 1 (pad 0: the 1x1x1 gate convolutions, no halo)
template <int CIN, int NT, int TH, int TW, int KS>
__global__ void __launch_bounds__(256, (NT == 2 ? 2 : 1))
conv3d_halo_kernel(const HaloParams p) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: instnorm_finalize / _apply may be scheduled under this grid's tail (no-op without a PDL dependent)
    constexpr int TD = 32, PAD = KS / 2, HH = TH + 2 * PAD, HW = TW + 2 * PAD, HD = TD + 2 * PAD;
    constexpr int TAPS = KS * KS * KS;
    constexpr int VB = CIN * 2;                               // bytes per halo voxel (no padding: swizzled)
    constexpr int CPV = CIN / 8;                              // 16-byte chunks per voxel
    constexpr int NVOX = HH * HW * HD;
    constexpr int KTOT = TAPS * CIN, KSTEPS = (KTOT + 15) / 16;
    constexpr int WROW = KSTEPS * 16 * 2 + 16;                // weight row pitch in smem (odd multiple of 16 B)
    constexpr int NROWS = NT * 8;
    constexpr int MT = TH * TW * 2;                           // m-tiles (16 voxels along D) per CTA tile
    constexpr int MG = 4;                                     // m-tiles per warp pass (share B fragments)
    static_assert(MT % (8 * MG) == 0, "tile must split into passes of 4 m-tiles per warp");
    static_assert((WROW / 16) % 2 == 1, "weight rows must have an odd 16-byte pitch (ldmatrix bank spread)");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sH = smem;                                 // halo  [NVOX][VB]
    unsigned char* sW = smem + ((NVOX * VB + 127) / 128) * 128;   // weights [NROWS][WROW]
    float* sred = reinterpret_cast<float*>(sW + NROWS * WROW);    // [8 warps][NROWS][2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y;
    const int td_i = blockIdx.x % p.tiles_d;
    const int tw_i = (blockIdx.x / p.tiles_d) % p.tiles_w;
    const int th_i = blockIdx.x / (p.tiles_d * p.tiles_w);
    const int h0 = th_i * TH, w0 = tw_i * TW, d0 = td_i * TD;
    const int Cin = p.C0 + p.C1;
    const int64_t sample = (int64_t)b * p.H * p.W * p.D;

    // ---- stage the halo (zero fill outside the volume) and the weights
    // a warp per (h, w) halo row, lanes along depth x channel chunk: no integer divisions (round 2: IMAD was 17-21 % of
    // the kernel's issue samples, most of it the div / mod chain of a flat staging loop)
    for (int row = warp; row < HH * HW; row += 8) {
        const int hh = row / HW, hw = row - hh * HW;
        const int gh = h0 - PAD + hh, gw = w0 - PAD + hw;
        const bool okhw = gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
        const int64_t vrow = sample + ((int64_t)gh * p.W + gw) * p.D;
        for (int j = lane; j < HD * CPV; j += 32) {
            const int hd = j / CPV, cc = j - hd * CPV;                     // CPV is a power of two
            const int gd = d0 - PAD + hd;
            const bool ok = okhw && gd >= 0 && gd < p.D;
            const int c = cc * 8;
            const int64_t vox = vrow + gd;
            const bf16* src = c < p.C0 ? p.in0 + vox * p.C0 + c : p.in1 + vox * p.C1 + (c - p.C0);
            const int v = row * HD + hd;
            const int sw = CIN >= 64 ? (v & 7) : (CIN == 32 ? ((v >> 1) & 3) : (CIN == 16 ? ((v >> 2) & 1) : 0));
            cp_async16_zfill(sH + v * VB + ((cc ^ sw) << 4), ok ? src : p.in0, ok ? 16 : 0);
        }
    }
    {
        constexpr int WCH = KSTEPS * 2;                       // 16-byte chunks per weight row
        for (int i = tid; i < NROWS * WCH; i += 256) {
            const int n = i / WCH, c = i - n * WCH;
            cp_async16(sW + n * WROW + c * 16, p.weight + (int64_t)n * p.wld + c * 8);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    const uint32_t sH_u = smem_u32_generic(sH), sW_u = smem_u32_generic(sW);
    const int g = lane >> 2, tq = lane & 3;
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1);      // A: row inside the m-tile addressed by this lane
    const int lhi = lane >> 4;                                // A: upper 8 of the 16 K values
    // B: ldmatrix.x4 = (n 0-7,k lo) (n 0-7,k hi) (n 8-15,k lo) (n 8-15,k hi)
    const uint32_t b_lane = (uint32_t)(((lane & 7) + 8 * (lane >> 4)) * WROW + 16 * ((lane >> 3) & 1));

    float csum[NT][2], csq[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { csum[nt][0] = csum[nt][1] = csq[nt][0] = csq[nt][1] = 0.f; }
    float bias_r[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = nt * 8 + 2 * tq + j;
            bias_r[nt][j] = (p.bias != nullptr && c < p.Cout + p.naux) ? p.bias[c] : 0.f;
        }

    for (int pass = 0; pass < MT / (8 * MG); ++pass) {
        // the warp's four m-tiles: index -> (h, w, d-half)
        int vbase[MG];
#pragma unroll
        for (int m = 0; m < MG; ++m) {
            const int mt = (pass * 8 + warp) * MG + m;
            const int dh = mt & 1, w = (mt >> 1) % TW, h = (mt >> 1) / TW;
            vbase[m] = (h * HW + w) * HD + dh * 16 + lrow;   // halo voxel of tap (0,0,0) for this lane's row
        }
        float acc[MG][NT][4];
#pragma unroll
        for (int m = 0; m < MG; ++m)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;

#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
            uint32_t bfr[NT / 2][4];
#pragma unroll
            for (int np = 0; np < NT / 2; ++np) ldsm4(sW_u + b_lane + np * 16 * WROW + ks * 32, bfr[np]);
            // tap / channel chunk of this k-step
            int tapoff, chunk;
            if (CIN >= 16) {
                constexpr int SPT = CIN / 16;                 // k-steps per tap
                const int tap = ks / SPT;
                tapoff = KS == 3 ? ((tap / 9) * HW + (tap / 3) % 3) * HD + tap % 3 : 0;
                chunk = (ks % SPT) * 2 + lhi;
            } else {                                          // CIN == 8: two taps per k-step (tap 27 = zero weights)
                int tap = 2 * ks + lhi;
                if (tap > TAPS - 1) tap = TAPS - 1;
                tapoff = KS == 3 ? ((tap / 9) * HW + (tap / 3) % 3) * HD + tap % 3 : 0;
                chunk = 0;
            }
#pragma unroll
            for (int m = 0; m < MG; ++m) {
                const int v = vbase[m] + tapoff;
                const int sw = CIN >= 64 ? (v & 7) : (CIN == 32 ? ((v >> 1) & 3) : (CIN == 16 ? ((v >> 2) & 1) : 0));
                uint32_t a[4];
                ldsm4(sH_u + v * VB + ((chunk ^ sw) << 4), a);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
                    mma_bf16_16816(acc[m][nt], a, bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
            }
        }

        // ---- epilogue of these m-tiles: rows g and g+8 of each, columns nt*8 + 2*tq (+1)
#pragma unroll
        for (int m = 0; m < MG; ++m) {
            const int mt = (pass * 8 + warp) * MG + m;
            const int dh = mt & 1, w = (mt >> 1) % TW, h = (mt >> 1) / TW;
            const int gh = h0 + h, gw = w0 + w;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int gd = d0 + dh * 16 + g + 8 * half;
                const bool ok = gh < p.H && gw < p.W && gd < p.D;
                const int64_t vrow = sample + ((int64_t)gh * p.W + gw) * p.D + gd;
                const int64_t row = vrow * p.Cout;
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const int c = nt * 8 + 2 * tq;
                    float o0 = acc[m][nt][half * 2] + bias_r[nt][0], o1 = acc[m][nt][half * 2 + 1] + bias_r[nt][1];
                    if (p.naux > 0 && ok) {                        // fused head: unrounded fp32 logits
                        const int ca = c - p.Cout;
                        if (ca >= 0 && ca < p.naux) p.aux[vrow * p.naux + ca] = o0;
                        if (ca + 1 >= 0 && ca + 1 < p.naux) p.aux[vrow * p.naux + ca + 1] = o1;
                    }
                    if (!p.out_f32) {                          // statistics describe the stored (rounded) values
                        o0 = __bfloat162float(__float2bfloat16_rn(o0));
                        o1 = __bfloat162float(__float2bfloat16_rn(o1));
                    }
                    if (ok) {
                        if (c < p.Cout) { csum[nt][0] += o0; csq[nt][0] = fmaf(o0, o0, csq[nt][0]); }
                        if (c + 1 < p.Cout) { csum[nt][1] += o1; csq[nt][1] = fmaf(o1, o1, csq[nt][1]); }
                        if (p.out_f32) {
                            float* dst = reinterpret_cast<float*>(p.out) + row + c;
                            if (c < p.Cout) dst[0] = o0;
                            if (c + 1 < p.Cout) dst[1] = o1;
                        } else {
                            bf16* dst = reinterpret_cast<bf16*>(p.out) + row + c;
                            if (c + 1 < p.Cout) *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(o0, o1);
                            else if (c < p.Cout) dst[0] = __float2bfloat16_rn(o0);
                        }
                    }
                }
            }
        }
    }

    if (p.partials != nullptr) {
        // reduce over the 8 row groups g (lanes 4g+tq), then over the 8 warps in a fixed order
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
#pragma unroll
                for (int o = 4; o <= 16; o <<= 1) {
                    csum[nt][j] += __shfl_xor_sync(0xffffffffu, csum[nt][j], o);
                    csq[nt][j] += __shfl_xor_sync(0xffffffffu, csq[nt][j], o);
                }
                if (g == 0) {
                    const int c = nt * 8 + 2 * tq + j;
                    sred[(warp * NROWS + c) * 2] = csum[nt][j];
                    sred[(warp * NROWS + c) * 2 + 1] = csq[nt][j];
                }
            }
        __syncthreads();
        if (tid < p.Cout) {
            float s = 0.f, q = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { s += sred[(w * NROWS + tid) * 2]; q += sred[(w * NROWS + tid) * 2 + 1]; }
            float* dst = p.partials + (((int64_t)b * p.tiles + blockIdx.x) * p.Cout + tid) * 2;
            dst[0] = s; dst[1] = q;
        }
    }
}

// Round 2: PERSISTENT version of the same kernel (same tile, same fragments, same fma order, same partial sums).
// conv3d_halo_kernel loads a halo, waits, computes, exits: the load latency of a tile is only hidden by the second
// resident CTA, the two CTAs of an SM tend to run in lock step (ncu: HMMA pipe 47-50 % busy, stall_wait + stall_math on top)
// and every CTA re-reads 14-55 KB of weights for its 512 output voxels.  Here a CTA keeps the weights, walks the tiles
// T, T + grid, ... over the whole batch and stages the halo of its NEXT tile (cp.async, second buffer) before it computes
// the current one, so the loads of tile i+1 fly under the MMAs of tile i.  Staging is organised by halo rows (a warp per
// (h, w) row, lanes along depth x channel chunk): no integer divisions.
template <int CIN, int NT, int TH, int TW, int KS, int NW, int MG>
__global__ void __launch_bounds__(NW * 32, 1)
conv3d_halo_p_kernel(const HaloParams p, int total_tiles) {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // PDL: instnorm_finalize / _apply may be scheduled under this grid's tail (no-op without a PDL dependent)
    constexpr int TD = 32, PAD = KS / 2, HH = TH + 2 * PAD, HW = TW + 2 * PAD, HD = TD + 2 * PAD;
    constexpr int TAPS = KS * KS * KS;
    constexpr int VB = CIN * 2;
    constexpr int CPV = CIN / 8;
    constexpr int NVOX = HH * HW * HD;
    constexpr int HBYTES = ((NVOX * VB + 127) / 128) * 128;
    constexpr int KTOT = TAPS * CIN, KSTEPS = (KTOT + 15) / 16;
    constexpr int WROW = KSTEPS * 16 * 2 + 16;
    constexpr int NROWS = NT * 8;
    constexpr int MT = TH * TW * 2;
    static_assert(MT % (NW * MG) == 0, "tile must split into passes of MG m-tiles per warp");
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* sW = smem + 2 * HBYTES;                                 // weights [NROWS][WROW]
    float* sred = reinterpret_cast<float*>(sW + NROWS * WROW);             // [NW warps][NROWS][2]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    auto stage = [&](int buf, int T) {
        const int b = T / p.tiles, tile = T - b * p.tiles;
        const int td_i = tile % p.tiles_d;
        const int tw_i = (tile / p.tiles_d) % p.tiles_w;
        const int th_i = tile / (p.tiles_d * p.tiles_w);
        const int h0 = th_i * TH, w0 = tw_i * TW, d0 = td_i * TD;
        const int64_t sample = (int64_t)b * p.H * p.W * p.D;
        unsigned char* sH = smem + buf * HBYTES;
        for (int row = warp; row < HH * HW; row += NW) {
            const int hh = row / HW, hw = row - hh * HW;
            const int gh = h0 - PAD + hh, gw = w0 - PAD + hw;
            const bool okhw = gh >= 0 && gh < p.H && gw >= 0 && gw < p.W;
            const int64_t vrow = sample + ((int64_t)gh * p.W + gw) * p.D;
            for (int j = lane; j < HD * CPV; j += 32) {
                const int hd = j / CPV, cc = j - hd * CPV;                 // CPV is a power of two
                const int gd = d0 - PAD + hd;
                const bool ok = okhw && gd >= 0 && gd < p.D;
                const int c = cc * 8;
                const int64_t vox = vrow + gd;
                const bf16* src = c < p.C0 ? p.in0 + vox * p.C0 + c : p.in1 + vox * p.C1 + (c - p.C0);
                const int v = row * HD + hd;
                const int sw = CIN >= 64 ? (v & 7) : (CIN == 32 ? ((v >> 1) & 3) : (CIN == 16 ? ((v >> 2) & 1) : 0));
                cp_async16_zfill(sH + v * VB + ((cc ^ sw) << 4), ok ? src : p.in0, ok ? 16 : 0);
            }
        }
    };

    {
        constexpr int WCH = KSTEPS * 2;
        for (int i = tid; i < NROWS * WCH; i += NW * 32) {
            const int n = i / WCH, c = i - n * WCH;
            cp_async16(sW + n * WROW + c * 16, p.weight + (int64_t)n * p.wld + c * 8);
        }
    }
    int T = blockIdx.x;
    if (T < total_tiles) stage(0, T);
    cp_async_commit();

    const uint32_t sW_u = smem_u32_generic(sW);
    const int g = lane >> 2, tq = lane & 3;
    const int lrow = (lane & 7) + 8 * ((lane >> 3) & 1);
    const int lhi = lane >> 4;
    const uint32_t b_lane = (uint32_t)(((lane & 7) + 8 * (lane >> 4)) * WROW + 16 * ((lane >> 3) & 1));
    float bias_r[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int c = nt * 8 + 2 * tq + j;
            bias_r[nt][j] = (p.bias != nullptr && c < p.Cout + p.naux) ? p.bias[c] : 0.f;
        }

    for (int it = 0; T < total_tiles; T += gridDim.x, ++it) {
        const int buf = it & 1;
        const int Tn = T + gridDim.x;
        if (Tn < total_tiles) stage(buf ^ 1, Tn);          // the buffer was released by the barrier that ended the last tile
        cp_async_commit();
        cp_async_wait<1>();                                  // this tile's halo (and, first time, the weights) has landed
        __syncthreads();

        const int b = T / p.tiles, tile = T - b * p.tiles;
        const int td_i = tile % p.tiles_d;
        const int tw_i = (tile / p.tiles_d) % p.tiles_w;
        const int th_i = tile / (p.tiles_d * p.tiles_w);
        const int h0 = th_i * TH, w0 = tw_i * TW, d0 = td_i * TD;
        const int64_t sample = (int64_t)b * p.H * p.W * p.D;
        const uint32_t sH_u = smem_u32_generic(smem + buf * HBYTES);

        float csum[NT][2], csq[NT][2];
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) { csum[nt][0] = csum[nt][1] = csq[nt][0] = csq[nt][1] = 0.f; }

        for (int pass = 0; pass < MT / (NW * MG); ++pass) {
            int vbase[MG];
#pragma unroll
            for (int m = 0; m < MG; ++m) {
                const int mt = (pass * NW + warp) * MG + m;
                const int dh = mt & 1, w = (mt >> 1) % TW, h = (mt >> 1) / TW;
                vbase[m] = (h * HW + w) * HD + dh * 16 + lrow;
            }
            float acc[MG][NT][4];
#pragma unroll
            for (int m = 0; m < MG; ++m)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                    for (int i = 0; i < 4; ++i) acc[m][nt][i] = 0.f;

#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                uint32_t bfr[NT / 2][4];
#pragma unroll
                for (int np = 0; np < NT / 2; ++np) ldsm4(sW_u + b_lane + np * 16 * WROW + ks * 32, bfr[np]);
                int tapoff, chunk;
                if (CIN >= 16) {
                    constexpr int SPT = CIN / 16;
                    const int tap = ks / SPT;
                    tapoff = KS == 3 ? ((tap / 9) * HW + (tap / 3) % 3) * HD + tap % 3 : 0;
                    chunk = (ks % SPT) * 2 + lhi;
                } else {
                    int tap = 2 * ks + lhi;
                    if (tap > TAPS - 1) tap = TAPS - 1;
                    tapoff = KS == 3 ? ((tap / 9) * HW + (tap / 3) % 3) * HD + tap % 3 : 0;
                    chunk = 0;
                }
#pragma unroll
                for (int m = 0; m < MG; ++m) {
                    const int v = vbase[m] + tapoff;
                    const int sw = CIN >= 64 ? (v & 7) : (CIN == 32 ? ((v >> 1) & 3) : (CIN == 16 ? ((v >> 2) & 1) : 0));
                    uint32_t a[4];
                    ldsm4(sH_u + v * VB + ((chunk ^ sw) << 4), a);
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                        mma_bf16_16816(acc[m][nt], a, bfr[nt >> 1][(nt & 1) * 2], bfr[nt >> 1][(nt & 1) * 2 + 1]);
                }
            }

#pragma unroll
            for (int m = 0; m < MG; ++m) {
                const int mt = (pass * NW + warp) * MG + m;
                const int dh = mt & 1, w = (mt >> 1) % TW, h = (mt >> 1) / TW;
                const int gh = h0 + h, gw = w0 + w;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int gd = d0 + dh * 16 + g + 8 * half;
                    const bool ok = gh < p.H && gw < p.W && gd < p.D;
                    const int64_t vrow = sample + ((int64_t)gh * p.W + gw) * p.D + gd;
                    const int64_t row = vrow * p.Cout;
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt) {
                        const int c = nt * 8 + 2 * tq;
                        float o0 = acc[m][nt][half * 2] + bias_r[nt][0], o1 = acc[m][nt][half * 2 + 1] + bias_r[nt][1];
                        if (p.naux > 0 && ok) {
                            const int ca = c - p.Cout;
                            if (ca >= 0 && ca < p.naux) p.aux[vrow * p.naux + ca] = o0;
                            if (ca + 1 >= 0 && ca + 1 < p.naux) p.aux[vrow * p.naux + ca + 1] = o1;
                        }
                        if (!p.out_f32) {
                            o0 = __bfloat162float(__float2bfloat16_rn(o0));
                            o1 = __bfloat162float(__float2bfloat16_rn(o1));
                        }
                        if (ok) {
                            if (c < p.Cout) { csum[nt][0] += o0; csq[nt][0] = fmaf(o0, o0, csq[nt][0]); }
                            if (c + 1 < p.Cout) { csum[nt][1] += o1; csq[nt][1] = fmaf(o1, o1, csq[nt][1]); }
                            if (p.out_f32) {
                                float* dst = reinterpret_cast<float*>(p.out) + row + c;
                                if (c < p.Cout) dst[0] = o0;
                                if (c + 1 < p.Cout) dst[1] = o1;
                            } else {
                                bf16* dst = reinterpret_cast<bf16*>(p.out) + row + c;
                                if (c + 1 < p.Cout) *reinterpret_cast<uint32_t*>(dst) = pack_bf16x2(o0, o1);
                                else if (c < p.Cout) dst[0] = __float2bfloat16_rn(o0);
                            }
                        }
                    }
                }
            }
        }

        if (p.partials != nullptr) {
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int o = 4; o <= 16; o <<= 1) {
                        csum[nt][j] += __shfl_xor_sync(0xffffffffu, csum[nt][j], o);
                        csq[nt][j] += __shfl_xor_sync(0xffffffffu, csq[nt][j], o);
                    }
                    if (g == 0) {
                        const int c = nt * 8 + 2 * tq + j;
                        sred[(warp * NROWS + c) * 2] = csum[nt][j];
                        sred[(warp * NROWS + c) * 2 + 1] = csq[nt][j];
                    }
                }
            __syncthreads();
            if (tid < p.Cout) {
                float s = 0.f, q = 0.f;
#pragma unroll
                for (int w = 0; w < NW; ++w) { s += sred[(w * NROWS + tid) * 2]; q += sred[(w * NROWS + tid) * 2 + 1]; }
                float* dst = p.partials + (((int64_t)b * p.tiles + tile) * p.Cout + tid) * 2;
                dst[0] = s; dst[1] = q;
            }
        }
        __syncthreads();          // every warp is done with halo[buf] (and sred) before the next iteration overwrites them
    }
    cp_async_wait<0>();
}

template <int CIN, int NT, int TH, int TW, int KS, int NW, int MG>
static int halo_launch_p(HaloParams& p, int B, cudaStream_t st) {
    constexpr int TD = 32, PAD = KS / 2, NVOX = (TH + 2 * PAD) * (TW + 2 * PAD) * (TD + 2 * PAD);
    constexpr int KSTEPS = (KS * KS * KS * CIN + 15) / 16;
    constexpr int WROW = KSTEPS * 32 + 16, NROWS = NT * 8;
    constexpr size_t HBYTES = ((NVOX * CIN * 2 + 127) / 128) * 128;
    const size_t smem = 2 * HBYTES + (size_t)NROWS * WROW + (size_t)NW * NROWS * 2 * 4;
    const int tiles_h = (p.H + TH - 1) / TH;
    p.tiles_w = (p.W + TW - 1) / TW;
    p.tiles_d = (p.D + TD - 1) / TD;
    p.tiles = tiles_h * p.tiles_w * p.tiles_d;
    static thread_local int conf = -1;
    int dev; cudaGetDevice(&dev);
    if (conf != dev) {
        cudaFuncSetAttribute(conv3d_halo_p_kernel<CIN, NT, TH, TW, KS, NW, MG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        conf = dev;
    }
    const int64_t total = (int64_t)p.tiles * B;
    int grid = sm_count();
    if (grid > total) grid = (int)total;
    conv3d_halo_p_kernel<CIN, NT, TH, TW, KS, NW, MG><<<grid, NW * 32, smem, st>>>(p, (int)total);
    LTU_LAUNCH_CHECK("conv3d_halo");
    count_launch(1);
    return LTU_OK;
}

template <int CIN, int NT, int TH, int TW, int KS>
static int halo_launch(HaloParams& p, int B, cudaStream_t st) {
    constexpr int TD = 32, PAD = KS / 2, NVOX = (TH + 2 * PAD) * (TW + 2 * PAD) * (TD + 2 * PAD);
    constexpr int KSTEPS = (KS * KS * KS * CIN + 15) / 16;
    constexpr int WROW = KSTEPS * 32 + 16, NROWS = NT * 8;
    const size_t smem = ((NVOX * CIN * 2 + 127) / 128) * 128 + (size_t)NROWS * WROW + (size_t)8 * NROWS * 2 * 4;
    const int tiles_h = (p.H + TH - 1) / TH;
    p.tiles_w = (p.W + TW - 1) / TW;
    p.tiles_d = (p.D + TD - 1) / TD;
    p.tiles = tiles_h * p.tiles_w * p.tiles_d;
    static thread_local int conf = -1;
    int dev; cudaGetDevice(&dev);
    if (conf != dev) {
        cudaFuncSetAttribute(conv3d_halo_kernel<CIN, NT, TH, TW, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        conf = dev;
    }
    conv3d_halo_kernel<CIN, NT, TH, TW, KS><<<dim3(p.tiles, B), 256, smem, st>>>(p);
    LTU_LAUNCH_CHECK("conv3d_halo");
    count_launch(1);
    return LTU_OK;
}

static void halo_tile(int Cin, int& th, int& tw) {
    if (Cin >= 32) { th = 4; tw = 4; } else { th = 4; tw = 8; }
}

}  // namespace ltu

using namespace ltu;

extern "C" int ltu_conv3d_halo_supported(int C0, int C1, int Cout, int ksize, int sh, int sw, int sd, int pad, int up2) {
    const int Cin = C0 + C1;
    if (!((ksize == 3 && pad == 1) || (ksize == 1 && pad == 0)) || sh != 1 || sw != 1 || sd != 1 || up2) return 0;
    if (!(Cin == 8 || Cin == 16 || Cin == 32 || (Cin == 64 && ksize == 1)) || C0 % 8 != 0 || C1 % 8 != 0) return 0;
    if (ksize == 1 && Cin == 8) return 0;
    if (Cout < 1 || Cout > 32) return 0;
    return 1;
}

extern "C" int ltu_conv3d_halo_tiles(int H, int W, int D, int Cin) {
    int th, tw;
    halo_tile(Cin, th, tw);
    return ((H + th - 1) / th) * ((W + tw - 1) / tw) * ((D + 31) / 32);
}

extern "C" int ltu_conv3d_halo(const void* in0, int C0, const void* in1, int C1, int B, int H, int W, int D, int ksize,
                               const void* weight_bf16, int weight_ld, const float* bias, int Cout, void* out,
                               int out_f32, float* partials, int n_aux, float* aux_out, ltu_stream_t stream) {
    LTU_ARG_CHECK(in0 && weight_bf16 && out, "conv3d_halo: null pointer");
    LTU_ARG_CHECK(n_aux >= 0 && (n_aux == 0 || (aux_out && !out_f32)), "conv3d_halo: bad auxiliary head");
    LTU_ARG_CHECK(ltu_conv3d_halo_supported(C0, C1, Cout + n_aux, ksize, 1, 1, 1, ksize / 2, 0), "conv3d_halo: unsupported C0=%d C1=%d Cout=%d k=%d", C0, C1, Cout + n_aux, ksize);
    LTU_ARG_CHECK((in1 != nullptr) == (C1 > 0), "conv3d_halo: in1/C1 mismatch");
    LTU_ARG_CHECK(B > 0 && B <= 65535 && H > 0 && W > 0 && D > 0, "conv3d_halo: bad shape");
    const int Cin = C0 + C1;
    LTU_ARG_CHECK(weight_ld >= ((ksize * ksize * ksize * Cin + 15) / 16) * 16 && weight_ld % 8 == 0, "conv3d_halo: weight row stride too small");
    LTU_ARG_CHECK(((uintptr_t)in0 & 15) == 0 && ((uintptr_t)in1 & 15) == 0 && ((uintptr_t)weight_bf16 & 15) == 0 &&
                  ((uintptr_t)out & 3) == 0, "conv3d_halo: misaligned pointer");
    LTU_ARG_CHECK(out_f32 || Cout % 2 == 0, "conv3d_halo: bf16 output needs an even Cout");
    HaloParams p;
    p.in0 = (const bf16*)in0; p.in1 = (const bf16*)in1; p.C0 = C0; p.C1 = C1; p.H = H; p.W = W; p.D = D;
    p.weight = (const bf16*)weight_bf16; p.wld = weight_ld; p.bias = bias; p.Cout = Cout; p.out = out; p.out_f32 = out_f32;
    p.naux = n_aux; p.aux = aux_out;
    p.partials = partials;
    cudaStream_t st = (cudaStream_t)stream;
    const bool wide = Cout + n_aux > 16;
    // persistent, double-buffered variant for the 3x3x3 layers (LTU_HALO_PERSISTENT=0: the one-tile-per-CTA kernel, A/B)
    static const bool persistent = [] { const char* e = getenv("LTU_HALO_PERSISTENT"); return !(e && e[0] == '0'); }();
    if (ksize == 1) {
        if (Cin == 64) return wide ? halo_launch<64, 4, 4, 4, 1>(p, B, st) : halo_launch<64, 2, 4, 4, 1>(p, B, st);
        if (Cin == 16) return wide ? halo_launch<16, 4, 4, 8, 1>(p, B, st) : halo_launch<16, 2, 4, 8, 1>(p, B, st);
        return wide ? halo_launch<32, 4, 4, 4, 1>(p, B, st) : halo_launch<32, 2, 4, 4, 1>(p, B, st);
    }
    // Measured (tools/layer_profile.py, batch 8 of 128^3): the persistent kernel wins only where the one-tile kernel is
    // limited to one CTA per SM anyway (NT = 4: enc.block1.conv1 193 vs 226 us); for NT = 2 two independent CTAs per SM
    // overlap staging, MMAs and epilogue of different tiles better than one CTA prefetching its next halo (Cin 32: 537 vs
    // 422 us, Cin 16: 283 vs 246 us, Cin 8: 211 vs 166 us with 16 warps; 8 warps are slower still).
    if (persistent && wide && (int64_t)B * H * W * D < ((int64_t)1 << 31)) {
        if (Cin == 8)  return halo_launch_p<8, 4, 4, 8, 3, 16, 4>(p, B, st);
        if (Cin == 16) return halo_launch_p<16, 4, 4, 8, 3, 16, 4>(p, B, st);
        return halo_launch_p<32, 4, 4, 4, 3, 16, 2>(p, B, st);
    }
    if (Cin == 8)  return wide ? halo_launch<8, 4, 4, 8, 3>(p, B, st) : halo_launch<8, 2, 4, 8, 3>(p, B, st);
    if (Cin == 16) return wide ? halo_launch<16, 4, 4, 8, 3>(p, B, st) : halo_launch<16, 2, 4, 8, 3>(p, B, st);
    return wide ? halo_launch<32, 4, 4, 4, 3>(p, B, st) : halo_launch<32, 2, 4, 4, 3>(p, B, st);
}
