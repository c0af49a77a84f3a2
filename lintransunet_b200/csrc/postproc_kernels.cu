// Volume post-processing right after the sliding-window stitch (the caller side of the hot path),
// all exact integer / byte work on the uint8 vote volume, HBM-bound, no tensor cores:
//   vote_decide            threshold / round of the stitched vote fractions     inference_embed_attn.py:147,
//                                                                               inference_multi_classes.py:148
//   keep_largest_component MONAI 0.7.0 KeepLargestConnectedComponent            inference_multi_classes.py:104,:150
//                          (union-find labelling on the device; the reference leaves the GPU for skimage here)
//   overlap_counts         per-class, per-row TP / predicted / target counts    loss/criterions.py:46-69,:291-311,
//                          from which Dice, recall, precision and the          :359-379,:192-241 and their
//                          localisation profile loss follow on the host         loss/multi_criterions.py twins
#include "common.cuh"

namespace ltu {

void count_launch(int n = 1);

static inline unsigned pp_grid(int64_t items, int per_block, int waves) {
    int64_t b = ceil_div64(items, per_block);
    const int64_t cap = (int64_t)sm_count() * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ---------------------------------------------------------------- threshold / round of vote fractions
// frac = votes / sum_c votes in fp32 (IEEE division: what `output_image / count_map` computes), then
//   mode 0: frac >= thr            mode 1: rintf(frac)  (torch.round: half to even, so 0.5 -> 0)
template <int VEC>
__global__ void __launch_bounds__(256)
vote_decide_kernel(const uint8_t* __restrict__ votes, uint8_t* __restrict__ onehot, int C, int64_t V, int mode, float thr) {
    const int64_t n = V / VEC;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int tot[VEC];
#pragma unroll
        for (int k = 0; k < VEC; ++k) tot[k] = 0;
        for (int c = 0; c < C; ++c) {
            const uint32_t word = VEC == 4 ? *reinterpret_cast<const uint32_t*>(votes + (int64_t)c * V + i * 4)
                                           : (uint32_t)votes[(int64_t)c * V + i];
#pragma unroll
            for (int k = 0; k < VEC; ++k) tot[k] += (word >> (8 * k)) & 0xFFu;
        }
        for (int c = 0; c < C; ++c) {
            const uint32_t word = VEC == 4 ? *reinterpret_cast<const uint32_t*>(votes + (int64_t)c * V + i * 4)
                                           : (uint32_t)votes[(int64_t)c * V + i];
            uint32_t o = 0;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
                // an uncovered voxel (tot == 0) cannot occur in a stitched volume; 0/0 = NaN decides 0 in both modes
                const float f = __fdiv_rn((float)((word >> (8 * k)) & 0xFFu), (float)tot[k]);
                const uint32_t bit = mode == 0 ? (uint32_t)(f >= thr) : (uint32_t)(rintf(f) == 1.f);
                o |= bit << (8 * k);
            }
            if (VEC == 4) *reinterpret_cast<uint32_t*>(onehot + (int64_t)c * V + i * 4) = o;
            else onehot[(int64_t)c * V + i] = (uint8_t)o;
        }
    }
}

// Exact integer form of the two decisions the scripts use: with k votes of n, fp32(k/n) > 0.5 <=> 2k > n and
// fp32(k/n) >= 0.5 <=> 2k >= n (k/n differs from 0.5 by at least 1/(2n) >= 1/510 unless it IS 0.5, which fp32 holds
// exactly), and rintf(x) == 1 on [0,1] <=> x > 0.5 (half to even).  16 voxels per thread, packed-byte SIMD:
// k > n-k  (no overflow: n <= 255), masked by n != 0.
__device__ __forceinline__ uint32_t decide_word(uint32_t k, uint32_t n, bool ge) {
    const uint32_t rest = __vsub4(n, k);
    const uint32_t hit = ge ? __vcmpgeu4(k, rest) : __vcmpgtu4(k, rest);
    return hit & __vcmpne4(n, 0u) & 0x01010101u;
}

template <int CMAX>
__global__ void __launch_bounds__(256)
vote_decide_half_kernel(const uint8_t* __restrict__ votes, uint8_t* __restrict__ onehot, int C, int64_t V, int ge) {
    const int64_t n16 = V / 16;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 w[CMAX];
        uint4 tot = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
                w[c] = __ldg(reinterpret_cast<const uint4*>(votes + (int64_t)c * V) + i);
                tot.x = __vadd4(tot.x, w[c].x); tot.y = __vadd4(tot.y, w[c].y);
                tot.z = __vadd4(tot.z, w[c].z); tot.w = __vadd4(tot.w, w[c].w);
            }
        }
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < C) {
                uint4 o;
                o.x = decide_word(w[c].x, tot.x, ge); o.y = decide_word(w[c].y, tot.y, ge);
                o.z = decide_word(w[c].z, tot.z, ge); o.w = decide_word(w[c].w, tot.w, ge);
                reinterpret_cast<uint4*>(onehot + (int64_t)c * V)[i] = o;
            }
        }
    }
}

// ---------------------------------------------------------------- connected components (union-find)
// Labels live in an int32 volume: L[v] = parent voxel index (a root has L[v] == v), -1 = background.  Links only
// ever point to a SMALLER index of the same component (atomicMin), so the structure is a forest at every moment,
// the final root of a component is its smallest voxel index -- the first voxel a raster scan meets, i.e. the
// order in which skimage.measure.label numbers components -- and the result does not depend on thread timing.
__device__ __forceinline__ int cc_find(const int* L, int x) {
    for (;;) {
        const int p = __ldcg(L + x);
        if (p == x) return x;
        x = p;
    }
}

__device__ __forceinline__ void cc_union(int* L, int a, int b) {
    for (;;) {
        a = cc_find(L, a);
        b = cc_find(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }       // a > b: hang a under b
        const int old = atomicMin(L + a, b);
        if (old == a) return;                                // a was still a root: linked
        a = old;                                             // somebody re-parented a meanwhile: retry from there
    }
}

// Foreground test + run linking without atomics: the 32 voxels of a warp are consecutive in memory (along D);
// every foreground voxel points straight at the first voxel of its run inside the warp's segment (a run also ends at
// a row boundary, d == 0).  Runs that continue across a segment boundary are joined by one union in cc_merge.
__global__ void __launch_bounds__(256)
cc_init_kernel(const uint8_t* __restrict__ onehot, int C, unsigned applied, int64_t V, int D, int* __restrict__ L,
               int* __restrict__ cnt, unsigned long long* __restrict__ best) {
    if (blockIdx.x == 0 && threadIdx.x == 0) *best = 0ull;
    const int lane = threadIdx.x & 31;
    // block-uniform trip count: every lane reaches the ballots
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < V; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = base + threadIdx.x;
        bool fg = false;
        if (v < V)
            for (int c = 0; c < C; ++c)
                if ((applied >> c) & 1u) fg |= onehot[(int64_t)c * V + v] != 0;
        const unsigned fgbits = __ballot_sync(0xffffffffu, fg);
        const unsigned d0bits = __ballot_sync(0xffffffffu, v < V && (v % D) == 0);
        // bit s set: lane s begins a run segment (first lane, or its predecessor is background, or a new row starts)
        const unsigned breaks = 1u | (~fgbits << 1) | d0bits;
        if (v < V) {
            const int start = 31 - __clz(breaks & ((2u << lane) - 1u));
            L[v] = fg ? (int)(v - (lane - start)) : -1;
            cnt[v] = 0;
        }
    }
}

// 26-connectivity (the reference's setting), run based: between a run R of this row and a run U of one of the four
// rows that precede it in raster order, ONE union is enough, and it is issued by
//   * R's first voxel a, if U covers a-1, or starts at a or a+1;
//   * otherwise by the voxel of R that sits just before U's first voxel (U starts at c >= a+2: voxel c-1).
// A voxel with a foreground predecessor therefore only looks at column d+1 of the neighbouring rows, and only when
// a run starts there; the interior of a solid object issues no union at all.
__global__ void __launch_bounds__(256)
cc_merge26_kernel(int* __restrict__ L, int H, int W, int D) {
    const int64_t V = (int64_t)H * W * D;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        const int lv = __ldcg(L + v);
        if (lv < 0) continue;
        const int d = (int)(v % D);
        const int64_t t = v / D;
        const int w = (int)(t % W), h = (int)(t / W);
        const bool prev = d > 0 && __ldcg(L + v - 1) >= 0;
        if (prev && (v & 31) == 0) cc_union(L, (int)v, (int)v - 1);     // first lane of a segment (cc_init): the run continues
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int dh = r < 3 ? -1 : 0, dw = r < 3 ? r - 1 : -1;
            const int hh = h + dh, ww = w + dw;
            if (hh < 0 || ww < 0 || ww >= W) continue;
            const int* row = L + ((int64_t)hh * W + ww) * D;
            const bool um = d > 0 && __ldcg(row + d - 1) >= 0;           // U covers d-1 / d / d+1
            const bool u0 = __ldcg(row + d) >= 0;
            const bool up = d + 1 < D && __ldcg(row + d + 1) >= 0;
            if (!prev) {
                if (um) cc_union(L, (int)v, (int)(row - L) + d - 1);
                else if (u0) cc_union(L, (int)v, (int)(row - L) + d);
            }
            if (up && !u0) cc_union(L, (int)v, (int)(row - L) + d + 1);
        }
    }
}

// connectivity 1 and 2 (6 / 18 neighbours): every foreground voxel links itself to its foreground neighbours that
// precede it in raster order (3 of 6, 9 of 18: at most `connectivity` non-zero offsets)
__global__ void __launch_bounds__(256)
cc_merge_kernel(int* __restrict__ L, int H, int W, int D, int connectivity) {
    const int64_t V = (int64_t)H * W * D;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        if (__ldcg(L + v) < 0) continue;
        const int d = (int)(v % D);
        const int64_t t = v / D;
        const int w = (int)(t % W), h = (int)(t / W);
        for (int dh = -1; dh <= 0; ++dh)
            for (int dw = -1; dw <= 1; ++dw)
                for (int dd = -1; dd <= 1; ++dd) {
                    // keep offsets that are lexicographically negative: (dh,dw,dd) < (0,0,0)
                    if (!(dh < 0 || (dh == 0 && (dw < 0 || (dw == 0 && dd < 0))))) continue;
                    if ((dh != 0) + (dw != 0) + (dd != 0) > connectivity) continue;
                    const int hh = h + dh, ww = w + dw, d2 = d + dd;
                    if (hh < 0 || ww < 0 || ww >= W || d2 < 0 || d2 >= D) continue;
                    const int64_t u = ((int64_t)hh * W + ww) * D + d2;
                    if (__ldcg(L + u) < 0) continue;
                    cc_union(L, (int)v, (int)u);
                }
    }
}

__global__ void __launch_bounds__(256)
cc_count_kernel(int* __restrict__ L, int* __restrict__ cnt, int64_t V) {
    const int lane = threadIdx.x & 31;
    // block-uniform trip count: every lane reaches the warp-wide match below
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x; base < V; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t v = base + threadIdx.x;
        int r = -1;
        if (v < V && __ldcg(L + v) >= 0) {
            r = cc_find(L, (int)v);
            L[v] = r;                                        // path compression: still a valid forest for concurrent finds
        }
        const unsigned same = __match_any_sync(0xffffffffu, r);     // one atomic per distinct root in the warp
        if (r >= 0 && lane == __ffs(same) - 1) atomicAdd(cnt + r, __popc(same));
    }
}

// largest component; ties go to the smallest root index = the lowest skimage label = np.argmax's first maximum
__global__ void __launch_bounds__(256)
cc_best_kernel(const int* __restrict__ L, const int* __restrict__ cnt, int64_t V, unsigned long long* __restrict__ best) {
    unsigned long long mine = 0ull;
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        if (L[v] != (int)v) continue;
        const unsigned long long key = ((unsigned long long)(unsigned)cnt[v] << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)v);
        if (key > mine) mine = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, o);
        if (other > mine) mine = other;
    }
    if ((threadIdx.x & 31) == 0 && mine != 0ull) atomicMax(best, mine);
}

__global__ void __launch_bounds__(256)
cc_apply_kernel(uint8_t* __restrict__ onehot, int C, unsigned applied, int64_t V, const int* __restrict__ L,
                const unsigned long long* __restrict__ best) {
    const unsigned long long key = *best;
    if (key == 0ull) return;                                 // no foreground at all: nothing to remove
    const int root = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
    for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < V; v += (int64_t)gridDim.x * blockDim.x) {
        const int l = L[v];
        if (l < 0 || l == root) continue;
        for (int c = 0; c < C; ++c)
            if ((applied >> c) & 1u) onehot[(int64_t)c * V + v] = 0;
    }
}

// ---------------------------------------------------------------- overlap counts
// grid (H, C + 1): class c < C compares pred[c] with (target == c); pseudo-class C is the foreground
// (1 - pred[0]) vs (target != 0) of loss/multi_criterions.py:49-50,:239-240.  out int64 [C+1][H][3] = TP, P, T.
__global__ void __launch_bounds__(256)
overlap_counts_kernel(const uint8_t* __restrict__ pred, const uint8_t* __restrict__ target, int C, int H, int64_t row,
                      int vec_ok, long long* __restrict__ out) {
    const int h = blockIdx.x, c = blockIdx.y;
    const bool fgclass = c == C;
    const uint8_t* p = pred + ((int64_t)(fgclass ? 0 : c) * H + h) * row;
    const uint8_t* t = target + (int64_t)h * row;
    int tp = 0, np = 0, nt = 0;
    int64_t done = 0;
    if (vec_ok) {                                            // 16 voxels per load, packed-byte compares, popcounts
        const uint32_t cls = 0x01010101u * (uint32_t)(c & 0xFF);
        const int64_t n16 = row / 16;
        for (int64_t i = threadIdx.x; i < n16; i += 256) {
            const uint4 pw = __ldg(reinterpret_cast<const uint4*>(p) + i);
            const uint4 tw = __ldg(reinterpret_cast<const uint4*>(t) + i);
            const uint32_t pa[4] = {pw.x, pw.y, pw.z, pw.w}, ta[4] = {tw.x, tw.y, tw.z, tw.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t pv = (fgclass ? __vcmpeq4(pa[k], 0u) : __vcmpne4(pa[k], 0u)) & 0x01010101u;
                const uint32_t tv = (fgclass ? __vcmpne4(ta[k], 0u) : __vcmpeq4(ta[k], cls)) & 0x01010101u;
                tp += __popc(pv & tv); np += __popc(pv); nt += __popc(tv);
            }
        }
        done = n16 * 16;
    }
    for (int64_t i = done + threadIdx.x; i < row; i += 256) {
        const int pv = fgclass ? (p[i] == 0) : (p[i] != 0);
        const int tv = fgclass ? (t[i] != 0) : (t[i] == c);
        tp += pv & tv; np += pv; nt += tv;
    }
    __shared__ int red[3][8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        tp += __shfl_xor_sync(0xffffffffu, tp, o);
        np += __shfl_xor_sync(0xffffffffu, np, o);
        nt += __shfl_xor_sync(0xffffffffu, nt, o);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = tp; red[1][threadIdx.x >> 5] = np; red[2][threadIdx.x >> 5] = nt; }
    __syncthreads();
    if (threadIdx.x < 3) {
        long long s = 0;
        for (int w = 0; w < 8; ++w) s += red[threadIdx.x][w];
        out[((int64_t)c * H + h) * 3 + threadIdx.x] = s;
    }
}

}  // namespace ltu

using namespace ltu;

extern "C" int ltu_vote_decide(const uint8_t* votes, uint8_t* onehot, int C, int64_t voxels, int mode, float thr,
                               ltu_stream_t stream) {
    LTU_ARG_CHECK(votes && onehot && C > 0 && C <= 255 && voxels > 0, "vote_decide: bad arguments");
    LTU_ARG_CHECK(mode == LTU_DECIDE_THRESHOLD || mode == LTU_DECIDE_ROUND, "vote_decide: mode must be 0 (threshold) or 1 (round)");
    // the decisions of the reference scripts (round, threshold 0.5) in exact integer form, 16 voxels per thread
    const bool half = mode == LTU_DECIDE_ROUND || thr == 0.5f;
    if (half && C <= 8 && voxels % 16 == 0 && ((reinterpret_cast<uintptr_t>(votes) | reinterpret_cast<uintptr_t>(onehot)) & 15) == 0) {
        const int ge = mode == LTU_DECIDE_THRESHOLD;
        const unsigned g = pp_grid(voxels / 16, 256, 8);
        if (C <= 4) vote_decide_half_kernel<4><<<g, 256, 0, (cudaStream_t)stream>>>(votes, onehot, C, voxels, ge);
        else vote_decide_half_kernel<8><<<g, 256, 0, (cudaStream_t)stream>>>(votes, onehot, C, voxels, ge);
        LTU_LAUNCH_CHECK("vote_decide");
        count_launch(1);
        return LTU_OK;
    }
    const bool vec = voxels % 4 == 0 && ((reinterpret_cast<uintptr_t>(votes) | reinterpret_cast<uintptr_t>(onehot)) & 3) == 0;
    if (vec) vote_decide_kernel<4><<<pp_grid(voxels / 4, 256, 16), 256, 0, (cudaStream_t)stream>>>(votes, onehot, C, voxels, mode, thr);
    else vote_decide_kernel<1><<<pp_grid(voxels, 256, 16), 256, 0, (cudaStream_t)stream>>>(votes, onehot, C, voxels, mode, thr);
    LTU_LAUNCH_CHECK("vote_decide");
    count_launch(1);
    return LTU_OK;
}

extern "C" size_t ltu_keep_largest_component_workspace(int64_t voxels) {
    return voxels > 0 ? 16 + (size_t)voxels * 8 : 0;         // best key | labels int32[V] | counts int32[V]
}

extern "C" int ltu_keep_largest_component(uint8_t* onehot, int C, unsigned applied_mask, int H, int W, int D,
                                          int connectivity, void* workspace, size_t ws_bytes, ltu_stream_t stream) {
    const int64_t V = (int64_t)H * W * D;
    LTU_ARG_CHECK(onehot && workspace, "keep_largest_component: null pointer");
    LTU_ARG_CHECK(C > 0 && C <= 32 && H > 0 && W > 0 && D > 0 && V < 0x7FFFFFFFll, "keep_largest_component: bad shape");
    LTU_ARG_CHECK(connectivity >= 1 && connectivity <= 3, "keep_largest_component: connectivity must be 1, 2 or 3");
    LTU_ARG_CHECK(applied_mask != 0 && (C == 32 || (applied_mask >> C) == 0), "keep_largest_component: applied labels outside [0, C)");
    LTU_ARG_CHECK(ws_bytes >= ltu_keep_largest_component_workspace(V), "keep_largest_component: workspace too small");
    LTU_ARG_CHECK((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "keep_largest_component: workspace must be 16-byte aligned");
    unsigned long long* best = reinterpret_cast<unsigned long long*>(workspace);
    int* L = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(workspace) + 16);
    int* cnt = L + V;
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = pp_grid(V, 256, 16);
    cc_init_kernel<<<g, 256, 0, st>>>(onehot, C, applied_mask, V, D, L, cnt, best);
    LTU_LAUNCH_CHECK("cc_init");
    if (connectivity == 3) cc_merge26_kernel<<<g, 256, 0, st>>>(L, H, W, D);
    else cc_merge_kernel<<<g, 256, 0, st>>>(L, H, W, D, connectivity);
    LTU_LAUNCH_CHECK("cc_merge");
    cc_count_kernel<<<g, 256, 0, st>>>(L, cnt, V);
    LTU_LAUNCH_CHECK("cc_count");
    cc_best_kernel<<<g, 256, 0, st>>>(L, cnt, V, best);
    LTU_LAUNCH_CHECK("cc_best");
    cc_apply_kernel<<<g, 256, 0, st>>>(onehot, C, applied_mask, V, L, best);
    LTU_LAUNCH_CHECK("cc_apply");
    count_launch(5);
    return LTU_OK;
}

extern "C" int ltu_overlap_counts(const uint8_t* pred_onehot, const uint8_t* target, int C, int H, int W, int D,
                                  int64_t* counts, ltu_stream_t stream) {
    LTU_ARG_CHECK(pred_onehot && target && counts, "overlap_counts: null pointer");
    LTU_ARG_CHECK(C > 0 && C <= 255 && H > 0 && W > 0 && D > 0 && (int64_t)W * D < 0x7FFFFFFFll, "overlap_counts: bad shape");
    const int64_t row = (int64_t)W * D;
    // rows start 16-byte aligned when the bases are and the row length is a multiple of 16
    const int vec_ok = row % 16 == 0 && ((reinterpret_cast<uintptr_t>(pred_onehot) | reinterpret_cast<uintptr_t>(target)) & 15) == 0;
    overlap_counts_kernel<<<dim3((unsigned)H, (unsigned)(C + 1)), 256, 0, (cudaStream_t)stream>>>(
        pred_onehot, target, C, H, row, vec_ok, reinterpret_cast<long long*>(counts));
    LTU_LAUNCH_CHECK("overlap_counts");
    count_launch(1);
    return LTU_OK;
}
