/*
 * ltu_b200.h -- C ABI of libltu_b200.so: the sm_100a kernels behind the drop-in
 * LinTransUNet `MaskTransUnet` forward (lintransunet_b200/).
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer owned by the caller (PyTorch); the library never
 *     allocates, frees or retains memory, has no hidden global buffers, never syncs;
 *   - feature maps are channels-last  [B, H, W, D, C]  (the reference's [B,C,H,W,D] with the
 *     channel axis moved innermost; a token matrix [B, N=H*W*D, C] is the same memory);
 *   - `dtype` selects the STORAGE type of activations: LTU_F32 or LTU_BF16; all arithmetic and
 *     all statistics are fp32; parameters are always fp32 unless stated;
 *   - return 0 = ok, <0 = argument error (nothing launched), >0 = cudaError_t of the launch;
 *     ltu_last_error() gives the message (thread-local);
 *   - re-entrant; kernels go to `stream`.
 *
 * Each entry point names the reference code (paths relative to the reference repo) it
 * replaces.
 */
#ifndef LTU_B200_H
#define LTU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st* ltu_stream_t; /* == cudaStream_t */

#define LTU_F32 0
#define LTU_BF16 1
#define LTU_OK 0
#define LTU_ERR_ARG (-1)

#define LTU_ACT_NONE 0
#define LTU_ACT_LRELU 1 /* LeakyReLU(0.01) */

const char* ltu_last_error(void);
int ltu_version(void);
/* number of kernels this library has launched in the calling process (all threads) */
int64_t ltu_launch_count(void);

/* ---- a1: linear_attention, model/trans_block.py:41-67 ------------------------------------
 * kv_reduce : ctx[b,h,j,e] = sum_n softmax_over_n(K)[b,n,h*32+j] * V[b,n,h*32+e]   (:59-60)
 *             K,V are [B,N,heads*32] views with row stride ld (elements), e.g. slices of a fused
 *             QKV projection.  ctx is fp32 [B,heads,32,32].  Deterministic two-stage reduction.
 * q_readout : out[b,n,h*32+e] = sum_j softmax_over_j(Q[b,n,h*32+:])[j]/sqrt(32) * ctx[b,h,j,e]
 *             (:50,:65) written in the merged-head [B,N,C] order of :165.                      */
size_t ltu_kv_reduce_workspace(int B, int64_t N, int heads);
int ltu_kv_reduce(const void* k, const void* v, int64_t ld, float* ctx, void* workspace,
                  size_t workspace_bytes, int B, int64_t N, int heads, int dtype,
                  ltu_stream_t stream);
int ltu_q_readout(const void* q, int64_t ld_q, const float* ctx, void* out, int64_t ld_out, int B,
                  int64_t N, int heads, int dtype, ltu_stream_t stream);

/* ---- a3: SelfAttentionLayer glue, model/trans_block.py:205-210 ----------------------------
 * add_layernorm: y = LayerNorm(x + res) * gamma + beta over C in {128,256} (eps 1e-6, :183)
 * gelu        : in-place exact erf GELU (:201,:208); the bias is applied by the GEMM.          */
int ltu_add_layernorm(const void* x, const void* res, const float* gamma, const float* beta,
                      void* y, int64_t rows, int C, float eps, int dtype, ltu_stream_t stream);
int ltu_gelu(void* x, int64_t n, int dtype, ltu_stream_t stream);
/* bf16 path, split token stream: the residual stream between encoder layers is carried as two bf16 tensors
 * hi = bf16(y), lo = bf16(y - hi).  The Linear layers read `hi` (the bf16 cast autocast applies to a Linear's
 * input, trans_block.py:155-157,:208) while the residual adds of :205,:209 see hi + lo, i.e. ~16 significant bits:
 * the reference keeps that stream in fp32 under autocast (layer_norm autocasts to fp32).  x_lo may be null. */
int ltu_add_layernorm_split(const void* x_hi, const void* x_lo, const void* res, const float* gamma,
                            const float* beta, void* y_hi, void* y_lo, int64_t rows, int C, float eps,
                            ltu_stream_t stream);

/* ---- a4: Conv3dPosEmbedding, model/trans_block.py:86-96 -----------------------------------
 * y = x + bias + depthwise3x3x3(x), zero pad 1, channels-last; w is fp32 [27][C] with the tap
 * index kh*9+kw*3+kd in the NATIVE (H,W,D) axes (w'[c,kh,kw,kd] = w_ref[c,0,kd,kh,kw]).       */
int ltu_posenc_dwconv3(const void* x, const float* w27c, const float* bias, void* y, int B, int H,
                       int W, int D, int C, int dtype, ltu_stream_t stream);
/* the same on the split bf16 token stream: the 27 taps read x_hi (a conv's input is cast to bf16 under autocast),
 * the residual term is x_hi + x_lo (x_lo may be null), the result is stored as y_hi, y_lo. */
int ltu_posenc_dwconv3_split(const void* x_hi, const void* x_lo, const float* w27c, const float* bias,
                             void* y_hi, void* y_lo, int B, int H, int W, int D, int C, ltu_stream_t stream);

/* ---- a10-a14: nn.Conv3d (+ InstanceNorm3d statistics), model/Unet_3Dblock.py:310-316,
 * :375,:422,:523-531,:588,:1328,:1353, gates :200-214 ---------------------------------------
 * Implicit-GEMM direct convolution, kernel 1 or 3, zero padding `pad`, stride (sh,sw,sd).
 *   in0 [B,Hi,Wi,Di,C0] (+ optional in1 [B,Hi,Wi,Di,C1] = torch.cat((in0,in1),1), :553)
 *   up2 = 1: the input is nearest-upsampled x2 on H,W,D on the fly (nn.Upsample before the
 *            up_embed conv, :421-422); Hi,Wi,Di are the STORED sizes.
 *   weight fp32 packed [taps][C0+C1][Cout], tap = kh*k*k + kw*k + kd;  bias fp32 [Cout] or NULL
 *   out [B,Ho,Wo,Do,Cout] in `dtype` (or fp32 when out_f32 != 0)
 *   partials (nullable) fp32 [B][tiles][Cout][2] = per-tile (sum, sum of squares) of the fp32
 *            results, tiles = ltu_conv3d_tiles(Ho*Wo*Do, Cout); feed to ltu_instnorm_finalize. */
int ltu_conv3d_tiles(int64_t out_voxels, int Cout);
int ltu_conv3d(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di,
               int up2, int ksize, int sh, int sw, int sd, int pad, const float* weight,
               const float* bias, int Cout, void* out, int out_f32, int Ho, int Wo, int Do,
               float* partials, int dtype, ltu_stream_t stream);

/* tcgen05 (UTCHMMA) implicit-GEMM path of the same convolution for bf16 activations:
 * weight_bf16 packed [Cout16][Kpad] (K-major B operand), K index = tap*(C0+C1) + c, zero padded to
 * Kpad = ltu_conv3d_tc_kpad(C0+C1, ksize) (multiple of 64) and to Cout16 = Cout rounded up to 16 rows.
 * up2 = 1 (nearest x2 upsample before a 3x3x3 conv): the weight is the FOLDED tensor
 * [8 parity classes][Cout16][ltu_conv3d_tc_kpad(Cin, 2)], class q = pa*4+pb*2+pc, tap t = th*4+tw*2+td
 * reading source voxel (a+th-1+pa, ...), with per-axis folded weights  parity 0: {w0, w1+w2},
 * parity 1: {w0+w1, w2}.
 * Requirements (ltu_conv3d_tc_supported): (ksize,pad) in {(3,1),(1,0)}, C0+C1 a power of two >= 8,
 * C0 % 8 == 0, Cout <= 256 (and Cout % 8 == 0 for bf16 output).  Output bf16, or fp32 when out_f32.
 * Same partials layout with tiles = ltu_conv3d_tc_tiles(out_voxels, up2).                       */
int ltu_conv3d_tc_supported(int C0, int C1, int Cout, int ksize, int pad);
int ltu_conv3d_tc_tiles(int64_t out_voxels, int up2);
int ltu_conv3d_tc_kpad(int Cin, int ksize);
/* Fused head (n_aux > 0, 3x3x3 only): the weight/bias carry n_aux extra output rows after the Cout
 * main ones (e.g. the 3-class mask head of ROIDecoder, Unet_3Dblock.py:1380, which reads the same
 * input as UpBlock.conv1); their unrounded fp32 results go to aux_out [B][V][n_aux] and are not part of
 * the statistics.  supported() is then asked with Cout + n_aux.                                     */
int ltu_conv3d_tc(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di,
                  int up2, int ksize, int sh, int sw, int sd, int pad, const void* weight_bf16,
                  const float* bias, int Cout, void* out, int out_f32, int Ho, int Wo, int Do,
                  float* partials, int n_aux, float* aux_out, ltu_stream_t stream);

/* Third-generation tensor-core convolution: TMA-loaded shared-memory halo + tcgen05 (conv_tc3.cu).  Same
 * results contract as ltu_conv3d_tc for stride-1 3x3x3 convolutions (pad 1) of bf16 activations with C0 >= 64,
 * C0 % 64 == 0, C1 % 64 == 0, Cout % 8 == 0, Cout <= 128, bf16 output, optional fp32 auxiliary head (n_aux extra
 * weight rows, as in ltu_conv3d_tc); up2 = the folded nearest-x2 upsample (Cout <= 32 only: wider folded layers
 * measured faster on the im2col kernel).  A CTA owns TH x 16 x 8 output voxels (TH = 1, 2 or 4); one 5-D TMA box brings the
 * (TH+2) x 18 x 10 halo of a 64-channel slice (each input voxel is fetched 2-4x instead of 27x) and every
 * filter tap reads it through a shifted UMMA descriptor.  weight_bf16 = the ltu_conv3d_tc packing with weight_rows = Cout + n_aux rounded up to 32
 * ([rows][Kpad], folded: [8][rows][Kpad]).  partials: [B][ltu_conv3d_tc3_tiles(B,Hi,Wi,Di,Cout,n_aux,up2)][Cout][2].   */
int ltu_conv3d_tc3_supported(int C0, int C1, int Cout, int ksize, int sh, int sw, int sd, int pad,
                             int up2, int out_f32, int n_aux);
int ltu_conv3d_tc3_tiles(int B, int Hi, int Wi, int Di, int Cout, int n_aux, int up2);
int ltu_conv3d_tc3(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di,
                   int up2, const void* weight_bf16, int weight_rows, int Kpad, const float* bias,
                   int Cout, void* out, float* partials, int n_aux, float* aux_out,
                   ltu_stream_t stream);
/* The small-channel stride-1 3x3x3 layers (model/Unet_3Dblock.py:310,:523,:528,:588,:1353 and the finest mask head
 * :1328: stem, enc.block0/1 conv1, dec.block2/3, final_block; 8, 16 or 32 channels per input) on the same kernel in
 * SUPER-VOXEL form: g = 64 / Cin consecutive voxels along D are one 64-channel row (the same memory viewed as
 * [B][H][W][D/g][64]; D % g == 0), the weights are the block-Toeplitz repacking
 * W'[(kh,kw,kg)][(p,c)][(delta,co)] = W[kh][kw][g(kg-1)+p-delta+1][c][co] (zero outside 0..2), so UMMA N = g * Cout and
 * the output [B][H][W][D/g][g*Cout] IS the channels-last [B][H][W][D][Cout] tensor.  The caller passes the viewed shapes
 * (C0 = 64, C1 in {0, 64}, Di = D/g, Cout = g * Cout, n_aux = g * n_aux, aux rows ordered (delta, a) after the
 * (delta, co) main rows; Cout = 0 with an auxiliary head only = fp32 output).  ks_mask: 27 bytes in HOST memory, bit ks
 * of entry t = the 16-channel K block ks of tap t is not identically zero (zero blocks are not issued).             */
int ltu_conv3d_tc3_masked(const void* in0, int C0, const void* in1, int C1, int B, int Hi, int Wi, int Di,
                          const void* weight_bf16, int weight_rows, int Kpad, const float* bias, int Cout,
                          void* out, float* partials, int n_aux, float* aux_out, const uint8_t* ks_mask,
                          ltu_stream_t stream);

/* nn.Linear on a bf16 token matrix with a fused epilogue (model/trans_block.py:166,:187-189,:205-210),
 * run by the persistent tcgen05 kernel (a 1x1x1 "convolution"):  y = epi(x W^T + b)
 *   x bf16 [rows][Cin], Cin a power of two in [8,1024]; weight_bf16 = the ltu_conv3d_tc packing of the
 *   [Cout][Cin] matrix ([Cout16][ltu_conv3d_tc_kpad(Cin,1)]); Cout <= 256, % 8 == 0;
 *   y bf16 [rows][ld_y], columns [0,Cout) of every row written (wide layers = column slices);
 *   epi 0: bias | 1: bias + exact-erf GELU | 2: LayerNorm(x W^T + b + residual) * gamma + beta
 *   (residual bf16 [rows][Cout], Cout % 32 == 0: one thread owns one complete output row).        */
int ltu_linear_tc(const void* x, int Cin, int64_t rows, const void* weight_bf16, const float* bias,
                  int Cout, void* y, int ld_y, int epi, const void* residual, const float* gamma,
                  const float* beta, float eps, ltu_stream_t stream);

/* nn.Linear of the encoder layers with d_model 256 (and the K/V projection of d_model 128) as ONE persistent,
 * warp-specialised TMA + tcgen05 kernel (csrc/linear_tma.cu): replaces F.linear (model/trans_block.py:155-157 Q/K/V
 * projections as one [3C][C] GEMM, :166 output projection, :208 linear1 / linear2) together with the bias, the exact-erf
 * GELU (:208) or the residual add + LayerNorm (:205-206, :209-210) that follow it.
 *     epi 0: y = x W^T + b | 1: y = gelu(x W^T + b) | 2: (y_hi, y_lo) = split(LayerNorm(x W^T + b + res_hi + res_lo))
 *   x bf16 [rows][K], K % 64 == 0, K <= 1024; w_bf16 = the nn.Linear weight [N][K] rounded to bf16 (row-major);
 *   bias fp32 [N]; N in {256, 512, 768}; y_hi bf16 [rows][N].  epi 2: N == 256, res_hi / res_lo bf16 [rows][256]
 *   (the split token stream of ltu_add_layernorm_split; res_lo may be null), gamma / beta fp32 [256], y_lo may be
 *   null.  All global traffic is TMA (x, W and the residual through one 3-stage ring, outputs through per-warp
 *   staging tiles); the residual is added on the tensor pipe (acc += R . I_64).                                  */
int ltu_linear_fused(const void* x, int64_t rows, int K, const void* w_bf16, const float* bias, int N,
                     int epi, const void* res_hi, const void* res_lo, const float* gamma,
                     const float* beta, float eps, void* y_hi, void* y_lo, ltu_stream_t stream);

/* The same launch with the query half of linear_attention (model/trans_block.py:50 softmax(q) / sqrt(d_k), :65 q . ctx)
 * carried by the GEMMs around it, so that q_readout never runs and the attention output is never written:
 *   softmax_cols (epi 0; 0 or a multiple of 256 <= N): output columns [0, softmax_cols) -- the Q third of the fused
 *     [3C][C] QKV projection -- are written as softmax over each head's 32 columns, divided by sqrt(32);
 *   samples > 1: x / residual / y are [samples][rows / samples][cols] and w_bf16 holds ONE [N][K] matrix per sample,
 *     [samples][N][K].  With ltu_ctx_project's W_b the output projection (:166) of softmax(Q) equals readout + projection:
 *     (P ctx_b) Wo^T = P (blockdiag(ctx_b) Wo^T).  Row tiles are aligned to samples.
 *   ldx: elements between two rows of x (>= K, % 8 == 0): the operand may be a column slice of wider rows (P inside QKV).
 * ltu_linear_fused(x, rows, K, ...) == ltu_linear_fused_ex(x, K, rows, K, ..., 0, 1, stream).                         */
int ltu_linear_fused_ex(const void* x, int64_t ldx, int64_t rows, int K, const void* w_bf16, const float* bias, int N,
                        int epi, const void* res_hi, const void* res_lo, const float* gamma,
                        const float* beta, float eps, void* y_hi, void* y_lo, int softmax_cols, int samples,
                        ltu_stream_t stream);
/* ctx fp32 [B][heads][32][32] (ltu_kv_reduce: softmax_N(K)^T V per head) and the output projection's weight wo_bf16
 * [C][C] (C = 32 heads) -> out bf16 [B][C][C]:  W_b[n][32h + j] = sum_e ctx[b][h][j][e] Wo[n][32h + e].             */
int ltu_ctx_project(const float* ctx, const void* wo_bf16, void* out, int B, int heads, ltu_stream_t stream);
/* ltu_kv_reduce (bf16 k / v, 4 or 8 heads) whose fixed-order merge kernel also writes ltu_ctx_project's W_b into w_out
 * (bf16 [B][C][C]) -- no extra launch between the context reduction and the output projection.                      */
int ltu_kv_reduce_project(const void* k, const void* v, int64_t ld, float* ctx, void* workspace,
                          size_t workspace_bytes, int B, int64_t N, int heads, const void* wo_bf16,
                          void* w_out, ltu_stream_t stream);

/* Key / value half of linear_attention for d_model 128, 4 heads, as ONE launch (+ the fixed-order merge):
 *     ctx[b][h] = softmax_N(x_b Wk^T + bk)_h^T (x_b Wv^T + bv)_h
 * i.e. the K and V projections of MultihAttention.forward (model/trans_block.py:155-156) fused with the context
 * reduction of linear_attention (:59-60): K and V are never written to memory (the separate path, ltu_linear_fused +
 * ltu_kv_reduce, writes and re-reads 4 x rows x C x 2 bytes).  x bf16 [B][N][128]; w_kv bf16 [256][128] = Wk rows then Wv
 * rows; bias fp32 [256]; ctx fp32 [B][4][32][32] exactly as ltu_kv_reduce produces it (K, V rounded to bf16 the same way).
 * N % 32 == 0, B <= 30.  workspace: ltu_kv_project_reduce_workspace(B, N) bytes.  supported(): dispatch hint.
 * wo_bf16 [128][128] and w_out bf16 [B][128][128] (both or neither): the merge kernel also writes ltu_ctx_project's W_b. */
int ltu_kv_project_reduce_supported(int C, int heads, int64_t N);
size_t ltu_kv_project_reduce_workspace(int B, int64_t N);
int ltu_kv_project_reduce(const void* x, const void* w_kv, const float* bias, float* ctx, void* workspace,
                          size_t workspace_bytes, int B, int64_t N, const void* wo_bf16, void* w_out,
                          ltu_stream_t stream);

/* Fused feed-forward half of SelfAttentionLayer (model/trans_block.py:207-210: linear1 -> erf GELU ->
 * linear2 -> residual -> layer_norm2, dropouts are identity in eval) as ONE persistent tcgen05 kernel:
 *     y = LayerNorm(x + W2 gelu(W1 x + b1) + b2) * gamma + beta
 *   x, y bf16 [rows][C] (y may alias x); w1_bf16 [2C][C] and w2_bf16 [C][2C] are the nn.Linear weights
 *   rounded to bf16 (row-major, K innermost: the K-major UMMA operand, fetched by TMA); biases, gamma,
 *   beta fp32.  The 2C-wide hidden activation stays in tensor memory; HBM traffic is one read of x and
 *   one write of y.  C == 128: W1, W2 resident in shared memory.  C == 256: W1, W2 streamed from L2 through a TMA ring
 *   once per 128-row tile (optionally once per 2-CTA cluster, multicast), hidden dimension in four quarters.
 *   supported(C) is the dispatch hint of the model: 1 for 128; for 256 only with LTU_FFN256=1 (no faster yet).   */
int ltu_ffn_fused_supported(int C);
int ltu_ffn_fused(const void* x, int64_t rows, int C, const void* w1_bf16, const float* b1,
                  const void* w2_bf16, const float* b2, const float* gamma, const float* beta,
                  float eps, void* y, ltu_stream_t stream);
/* debugging aid: same launch, CTA 0 also records clock64() stamps of its pipeline events into
 * trace[2 roles][64 tiles][8 events] (int64, device memory, zero it first)                      */
int ltu_ffn_fused_trace(const void* x, int64_t rows, int C, const void* w1_bf16, const float* b1,
                        const void* w2_bf16, const float* b2, const float* gamma, const float* beta,
                        float eps, void* y, long long* trace, ltu_stream_t stream);

/* Query half of the encoder layer (model/trans_block.py:50,:65 readout, :155-166 projections, :205-206
 * residual + layer_norm1) for d_model 128 / 4 heads as ONE persistent tcgen05 kernel:
 *     y = LayerNorm(x + (softmax_d(x Wq^T + bq)/sqrt(32) . ctx_b) Wo^T + bo) * gamma + beta
 *   x, y bf16 [B][N][128]; wq_bf16, wo_bf16 the nn.Linear weights [128][128] rounded to bf16; biases, gamma,
 *   beta fp32; ctx_bf16 = ltu_ctx_pack_bf16 of the fp32 ctx [B][heads][32][32] produced by ltu_kv_reduce:
 *   bf16 [B*heads*32 rows (h*32+e)][64] with row[j] = ctx[b][h][j][e] for j < 32 and zeros above (the K-major
 *   tensor-core operand of the readout, fetched by TMA).  Q, the softmax and the attention output never leave
 *   tensor memory: HBM traffic is one read of x and one write of y.                                     */
int ltu_attn_out_fused_supported(int C, int heads);
int ltu_ctx_pack_bf16(const float* ctx, void* out, int B, int heads, ltu_stream_t stream);
int ltu_attn_out_fused(const void* x, int B, int64_t N, int C, int heads, const void* wq_bf16,
                       const float* bq, const void* ctx_bf16, const void* wo_bf16, const float* bo,
                       const float* gamma, const float* beta, float eps, void* y, ltu_stream_t stream);
/* The same half-layer with the readout folded into the output projection: (P ctx_b) Wo^T = P (blockdiag(ctx_b) Wo^T).
 * w_b bf16 [B][128][128] = ltu_ctx_project's per-sample weight (written by the merge kernel of ltu_kv_project_reduce /
 * ltu_kv_reduce_project).  Two chained GEMMs per tile instead of three; the residual is read from the x tile in shared
 * memory (x is read from HBM once and never re-read).                                                          */
int ltu_attn_out_fused_w(const void* x, int B, int64_t N, int C, int heads, const void* wq_bf16,
                         const float* bq, const void* w_b, const float* bo, const float* gamma,
                         const float* beta, float eps, void* y, ltu_stream_t stream);

/* Small-channel stride-1 3x3x3 convolution for bf16 activations (stem, enc.block0/1 conv1,
 * dec.block3, finest mask head, final_block): the input halo of a 3-D output tile and all weights are
 * staged once in shared memory and im2col happens in the ldmatrix row addresses of mma.sync
 * (bandwidth-bound layers).  Supported: (ksize,pad) = (3,1) or (1,0: the 1x1x1 gate convs,
 * Cin in {16,32,64}), stride 1, no up2, C0+C1 in {8,16,32} for 3x3x3, C0 % 8 == 0, Cout <= 32 (even for bf16 output).  weight_bf16 is the ltu_conv3d_tc packing
 * ([>=16 rows][weight_ld], K index = tap*Cin + c).  partials: [B][tiles][Cout][2] with
 * tiles = ltu_conv3d_halo_tiles(H, W, D, Cin).                                                  */
int ltu_conv3d_halo_supported(int C0, int C1, int Cout, int ksize, int sh, int sw, int sd, int pad,
                              int up2);
int ltu_conv3d_halo_tiles(int H, int W, int D, int Cin);
int ltu_conv3d_halo(const void* in0, int C0, const void* in1, int C1, int B, int H, int W, int D,
                    int ksize, const void* weight_bf16, int weight_ld, const float* bias, int Cout,
                    void* out, int out_f32, float* partials, int n_aux, float* aux_out,
                    ltu_stream_t stream);

/* InstanceNorm3d (no affine, eps 1e-5, biased variance):
 * finalize : partials [B][tiles][C][2] -> stats [B][C][2] = (mean, rstd), fixed summation order
 * partials : per-chunk (sum,sumsq) of an existing channels-last tensor, `chunks` per sample
 * apply    : y = act((x - mean) * rstd) (+ residual), DownBlock residual :330-331            */
int ltu_instnorm_finalize(const float* partials, float* stats, int B, int tiles, int C,
                          int64_t voxels, float eps, ltu_stream_t stream);
int ltu_chan_partials(const void* x, float* partials, int B, int64_t voxels, int C, int chunks,
                      int dtype, ltu_stream_t stream);
int ltu_instnorm_apply(const void* x, const float* stats, const void* residual, void* y, int B,
                       int64_t voxels, int C, int act, int dtype, ltu_stream_t stream);

/* ---- a10: windows_embedding, model/Unet_3Dblock.py:123-136 --------------------------------
 * x fp32 [B,1,H,W,D] -> y [B,H/2,W/2,D,cpad] channels-last, channel = kh*2+kw; cpad in {4,8}:
 * with cpad = 8 channels 4..7 are zero (pads the stem input to one 16-byte bf16 vector)        */
int ltu_s2d_input(const float* x, void* y, int B, int H, int W, int D, int cpad, int dtype,
                  ltu_stream_t stream);

/* torch.cat([a, b], dim=1) of the UpBlock (model/Unet_3Dblock.py:553) on channels-last rows: out [rows][Ca + Cb].  Used
 * where the two inputs have 32 channels each: the concatenation is ONE 64-channel input of ltu_conv3d_tc3 (TMA halo +
 * tcgen05), which the two-input form of that kernel cannot take (64-channel TMA boxes).                          */
int ltu_concat2(const void* a, int Ca, const void* b, int Cb, void* out, int64_t rows, int dtype, ltu_stream_t stream);

/* ---- a12: nn.Upsample(trilinear, align_corners=True), model/Unet_3Dblock.py:1341-1345 ------
 * scale factors (2,2,fd) with fd in {1,2}; channels-last                                     */
int ltu_upsample_trilinear(const void* x, void* y, int B, int H, int W, int D, int C, int fd,
                           int dtype, ltu_stream_t stream);

/* ---- a12: mask head softmax + foreground, model/Unet_3Dblock.py:1380-1387 ------------------
 * logits fp32 [B,V,Cout] -> mask fp32 [B,Cout,V] (nullable; the mask_list entry, reference
 * layout) and fg fp32 [B,V] = 1 - softmax[...,0]                                              */
int ltu_mask_softmax(const float* logits, float* mask, float* fg, int B, int64_t voxels, int Cout,
                     ltu_stream_t stream);

/* ---- a13: SpatialAttention3DBlock tail + skip gating, model/Unet_3Dblock.py:217-221,:1384-85
 * out = skip * sigmoid(psi_b + sum_c psi_w[c] * relu(IN(a)[c] + IN(g)[c])); a,g are the raw
 * 1x1x1 conv outputs [B,V,Ci] with their InstanceNorm stats [B][Ci][2]; Ci in {16,..,256}     */
int ltu_gate_fused(const void* a, const float* stats_a, const void* g, const float* stats_g,
                   const float* psi_w, const float* psi_b, const void* skip, void* out, int B,
                   int64_t voxels, int Ci, int dtype, ltu_stream_t stream);

/* ---- a8: ROIBridge.get_mask_boundary2 + get_min_max_indice, Unet_3Dblock.py:821-873,:37-49 -
 * fg fp32 [B,h,w,d] -> box fp32 [B,6] = [x0,y0,0,x1,y1,d-1]; on device, no host sync.
 * scratch: ltu_roi_bbox_scratch(B,h,w) bytes (int32 row/column profiles, zeroed by the call)  */
size_t ltu_roi_bbox_scratch(int B, int h, int w);
int ltu_roi_bbox(const float* fg, float* box, void* scratch, size_t scratch_bytes, int B, int h,
                 int w, int d, int min_h, int min_w, float thr, ltu_stream_t stream);

/* ---- a9: get_transfer_index/_back_index + grid_sample, Unet_3Dblock.py:51-82,:985-1039,
 * :1080-1117 -- separable piecewise-linear ("fisheye") bilinear resample, zeros padding,
 * align_corners=True.  direction 0: x [B,h,w,d,C] -> y [B,eval_h,eval_w,d,C];
 * direction 1: x [B,eval_h,eval_w,d,C] -> y [B,h,w,d,C].                                       */
int ltu_roi_resample(const void* x, const float* box, void* y, int B, int h, int w, int d, int C,
                     int roi_h, int roi_w, int eval_h, int eval_w, int direction, int dtype,
                     ltu_stream_t stream);

/* ---- a12/a15/a16: windows_unembedding + softmax + argmax one-hot, Unet_3Dblock.py:138-152,
 * :1392-1394, trans_3DUnet.py:196-202 -------------------------------------------------------
 * logits fp32 [B,H2,W2,D,4*Cout] (in-channel = c*4+kh*2+kw) -> any of (nullable):
 *   probs  fp32 [B,Cout,2*H2,2*W2,D], onehot fp32 (same shape), labels uint8 [B,2*H2,2*W2,D]  */
int ltu_head_d2s_softmax(const float* logits, float* probs, float* onehot, uint8_t* labels, int B,
                         int H2, int W2, int D, int Cout, ltu_stream_t stream);

/* ---- config 5: constant-blend sliding-window accumulation (MONAI 0.7.0
 * sliding_window_inference as called at inference_multi_classes.py:143) ----------------------
 * labels uint8 [nwin,rh,rw,rd] (argmax class of each window) are scattered as one-hot votes
 * into votes uint8 [C,H,W,D]; starts int32 [nwin,3].                                          */
int ltu_vote_accumulate(const uint8_t* labels, const int32_t* starts, uint8_t* votes, int nwin,
                        int rh, int rw, int rd, int C, int H, int W, int D, ltu_stream_t stream);
/* votes uint8 [C,H,W,D] -> labels uint8 [H,W,D] = argmax_c votes (first max wins, like
 * torch.argmax over vote fractions with a common denominator)                                */
int ltu_vote_argmax(const uint8_t* votes, uint8_t* labels, int C, int64_t voxels,
                    ltu_stream_t stream);
/* votes uint8 [C,V] -> frac fp32 [C,V] = votes / sum_c votes: the stitched output_image/count_map
 * of the constant-blend sliding window when every window prediction is one-hot                */
int ltu_vote_fractions(const uint8_t* votes, float* frac, int C, int64_t voxels,
                       ltu_stream_t stream);
/* out fp32 [nwin,1,rh,rw,rd] = windows of volume fp32 [H,W,D] at starts int32 [nwin,3]
 * (MONAI: torch.cat([inputs[win_slice] ...]))                                                  */
int ltu_gather_windows(const float* volume, const int32_t* starts, float* out, int nwin, int rh,
                       int rw, int rd, int H, int W, int D, ltu_stream_t stream);

/* ---- 8f-2: decisions on the stitched volume, inference_embed_attn.py:147 `(predict >= threshold)`
 * and inference_multi_classes.py:148 `torch.round(predict)` (half to even: 0.5 -> 0) ------------
 * votes uint8 [C,V] -> onehot uint8 [C,V] (0/1); the fraction votes / sum_c votes is formed in fp32
 * with an IEEE division exactly as ltu_vote_fractions does, without materialising it.           */
#define LTU_DECIDE_THRESHOLD 0
#define LTU_DECIDE_ROUND 1
int ltu_vote_decide(const uint8_t* votes, uint8_t* onehot, int C, int64_t voxels, int mode,
                    float thr, ltu_stream_t stream);

/* ---- 8f-4: monai.transforms.KeepLargestConnectedComponent(applied_labels, independent=False,
 * connectivity) as constructed at inference_multi_classes.py:104 and applied at :150 (MONAI 0.7.0,
 * one-hot branch: foreground = any applied channel; its largest connected component -- skimage
 * label order, first maximum wins -- is kept, every other foreground voxel is zeroed in the
 * applied channels).  onehot uint8 [C,H,W,D] is edited in place; applied_mask bit i = label i;
 * connectivity 1/2/3 = 6/18/26 neighbours.  independent=True is one call per label with a
 * single-bit mask.  workspace: ltu_keep_largest_component_workspace(H*W*D) bytes, 16-byte aligned. */
size_t ltu_keep_largest_component_workspace(int64_t voxels);
int ltu_keep_largest_component(uint8_t* onehot, int C, unsigned applied_mask, int H, int W, int D,
                               int connectivity, void* workspace, size_t ws_bytes,
                               ltu_stream_t stream);

/* ---- 8f-2/-4: the integer statistics behind the evaluation metrics of the inference scripts
 * (loss/criterions.py DiceClassLoss :46-69, Recall :291-311, Precision :359-379, LocalizationLoss
 * :192-241; loss/multi_criterions.py :41-55,:69-83,:232-281,:359-374) ---------------------------
 * pred_onehot uint8 [C,H,W,D] (0/1), target uint8 [H,W,D] (class index) ->
 * counts int64 [C+1][H][3] = (sum pred*target, sum pred, sum target) per class and H-row;
 * row C is the foreground pseudo-class (1 - pred[0]) vs (target != 0).                         */
int ltu_overlap_counts(const uint8_t* pred_onehot, const uint8_t* target, int C, int H, int W,
                       int D, int64_t* counts, ltu_stream_t stream);

/* ---- 8f-1 (first slice): backward of linear_attention, model/trans_block.py:41-67 (what autograd
 * does for `F.softmax(query,-1)/sqrt(d)`, `F.softmax(key,-2)`, and the two einsums) ------------
 * q,k,v,dout: [B,N,heads*32] views with row strides ldq / ldkv (k and v) / ldo; ctx fp32
 * [B,heads,32,32] from ltu_kv_reduce.  Writes dq,dk,dv (row stride ldd, activation dtype) and,
 * as by-products, dctx fp32 [B,heads,32,32] and kstats fp32 [B,heads,3,32] = (column max of K,
 * column sum of exp(K-max), sum_e dctx*ctx).  Nothing else has to be saved by the forward.
 * workspace: ltu_attn_bwd_workspace(B,N,heads) bytes; fixed-order merges (bit-reproducible).    */
size_t ltu_attn_bwd_workspace(int B, int64_t N, int heads);
int ltu_attn_bwd(const void* q, int64_t ldq, const void* k, const void* v, int64_t ldkv,
                 const void* dout, int64_t ldo, const float* ctx, void* dq, void* dk, void* dv,
                 int64_t ldd, float* dctx, float* kstats, void* workspace, size_t ws_bytes, int B,
                 int64_t N, int heads, int dtype, ltu_stream_t stream);

/* backward of `layer_norm(x + dropout(res))`, model/trans_block.py:205-206,:209-210: z = x + res is
 * rebuilt from the forward's inputs; dz [rows,C] is the gradient of BOTH x and res; dgamma, dbeta
 * fp32 [C] by per-CTA partials summed in a fixed order.  C in {128, 256}.
 * workspace: ltu_add_layernorm_bwd_workspace(rows, C) bytes.                                      */
size_t ltu_add_layernorm_bwd_workspace(int64_t rows, int C);
int ltu_add_layernorm_bwd(const void* x, const void* res, const void* dy, const float* gamma,
                          void* dz, float* dgamma, float* dbeta, void* workspace, size_t ws_bytes,
                          int64_t rows, int C, float eps, int dtype, ltu_stream_t stream);
/* ---- f4: deep-supervision training loss, loss/criterions.py:35-69 (DiceClassLoss), :416-443 (BalanceDiceLoss),
 * :696-718 (CrossEntroLoss) as driven by utils/utils_3D_embed_full.py:64-86.  Every criterion is a function of four sums
 * over the voxels of each (sample, class):
 *     sums[n][c] = ( sum p, sum onehot, sum p*onehot, sum -(1-p)*onehot*log(max(p,1e-6)) ),  onehot = (label == c)
 * p fp32 [N][C][V] (the reference layout [N,C,H,W,D]), labels uint8 [N][V].  One pass over p; fixed-order fp32 block
 * sums, fp64 across blocks.  workspace: ltu_loss_sums_workspace(N, C, V) bytes, 8-byte aligned.                  */
size_t ltu_loss_sums_workspace(int N, int C, int64_t V);
int ltu_loss_sums(const float* p, const uint8_t* labels, float* sums, void* workspace,
                  size_t workspace_bytes, int N, int C, int64_t V, ltu_stream_t stream);
/* backward: dp[n][c][v] = g[n][c][0] + onehot * (g[n][c][2] + g[n][c][3] * d/dp[-(1-p) log max(p,1e-6)]),
 * g = d(loss)/d(sums) fp32 [N][C][4] (the clamp passes its gradient for p >= 1e-6, like torch.clamp)              */
int ltu_loss_sums_bwd(const float* p, const uint8_t* labels, const float* gsums, float* dp, int N,
                      int C, int64_t V, ltu_stream_t stream);
/* label pyramid: F.max_pool3d(labels, kernel = stride = (kh,kw,kd)), utils/utils_3D_embed_full.py:65,:76-79, on uint8
 * labels [N][H][W][D] (H, W, D multiples of the kernel)                                                           */
int ltu_label_pool(const uint8_t* x, uint8_t* y, int N, int H, int W, int D, int kh, int kw, int kd,
                   ltu_stream_t stream);

/* Training-mode dropout (p = 0.3 by default, model/trans_3DUnet.py:162): nn.Dropout of
 * model/trans_block.py:205,:208,:209 and model/Unet_3Dblock.py:339,:382,:429,:556 (channelwise 0) and the nn.Dropout3d of
 * Conv3dPosEmbedding, model/trans_block.py:96 (channelwise 1: one draw per (sample, channel)) on a channels-last tensor:
 *     y = keep ? x / (1 - p) : 0,   keep(i) = Philox4x32-10(seed, offset + i / 4)[i % 4] >= p * 2^32
 * The decision is a pure function of (seed, offset, index): the backward applies the same call to the gradient.  n elements,
 * C channels innermost, per_sample elements per batch sample; consumes ceil(n / 4) counters (channel mode:
 * ceil(samples * C / 4)) starting at `offset`.  y may alias x.  n, C multiples of the 16-byte vector.              */
int ltu_dropout(const void* x, void* y, int64_t n, int C, int64_t per_sample, float p,
                uint64_t seed, uint64_t offset, int channelwise, int dtype, ltu_stream_t stream);

/* backward of the exact-erf F.gelu, model/trans_block.py:201,:208: dx = dy (Phi(x) + x phi(x)),
 * x = the pre-activation                                                                          */
int ltu_gelu_bwd(const void* x, const void* dy, void* dx, int64_t n, int dtype, ltu_stream_t stream);

/* backward of Conv3dPosEmbedding, model/trans_block.py:86-96, parameters: dw fp32 [27][C] in the
 * packing of ltu_posenc_dwconv3 (tap = (kh*3+kw)*3+kd), dbias fp32 [C].  The input gradient is
 * ltu_posenc_dwconv3 itself on dy with the 27 taps reversed and a zero bias.
 * workspace: ltu_posenc_wgrad_workspace(...) bytes; fixed-order sums.                              */
size_t ltu_posenc_wgrad_workspace(int B, int H, int W, int D, int C);
int ltu_posenc_wgrad(const void* x, const void* dy, float* dw, float* dbias, void* workspace,
                     size_t ws_bytes, int B, int H, int W, int D, int C, int dtype, ltu_stream_t stream);

/* backward of nn.InstanceNorm3d (+ LeakyReLU), model/Unet_3Dblock.py:325-336,:547-554: x = the RAW
 * convolution output [B,V,C] the forward normalised, stats fp32 [B,C,2] = (mean, rstd) of the
 * forward, dy the gradient of act(norm(x)); writes dx.  A residual added after the activation
 * simply receives dy.  workspace: ltu_instnorm_bwd_workspace(B, voxels, C) bytes; ordered sums.  */
size_t ltu_instnorm_bwd_workspace(int B, int64_t voxels, int C);
int ltu_instnorm_bwd(const void* x, const float* stats, const void* dy, void* dx, void* workspace,
                     size_t ws_bytes, int B, int64_t voxels, int C, int act, int dtype,
                     ltu_stream_t stream);

/* weight gradient of nn.Conv3d (what autograd computes for the convolutions of DownBlock / UpBlock /
 * the embed blocks / the gates, model/Unet_3Dblock.py:304-323,:519-538,:343-432,:195-215):
 * x bf16 [B,Hi,Wi,Di,Cin], dy bf16 [B,Ho,Wo,Do,Cout] -> dw fp32 [k^3][Cout][Cin] (tap = (kh*3+kw)*3+kd),
 * fp32 accumulation on the tensor pipe, ordered finalize.  k in {1,3}, any stride / padding, Cin and
 * Cout multiples of 8.  The INPUT gradient of a stride-1 convolution is the forward kernel applied
 * to dy with the filter reversed and its channel axes swapped (lintransunet_b200/backward.py).
 * workspace: ltu_conv3d_wgrad_workspace(...) bytes.                                               */
size_t ltu_conv3d_wgrad_workspace(int B, int Ho, int Wo, int Do, int Cin, int Cout, int ksize);
int ltu_conv3d_wgrad(const void* x, const void* dy, float* dw, void* workspace, size_t ws_bytes, int B,
                     int Hi, int Wi, int Di, int Cin, int Ho, int Wo, int Do, int Cout, int ksize, int sh,
                     int sw, int sd, int pad, int up2, ltu_stream_t stream);
/* up2 = 1: the convolution read nn.Upsample(nearest, x2) of x (UpEmbedBlock, Unet_3Dblock.py:419-429);
 * Ho/Wo/Do then refer to the upsampled extent.
 * Helpers for the INPUT gradients (both channels-last, C % 4 == 0):
 *  ltu_zero_insert: z[b, h*sh, w*sw, d*sd, :] = y[b,h,w,d,:], zero elsewhere -- the gradient of a strided
 *    convolution is the stride-1 forward kernel on z with the reversed, transposed filter;
 *  ltu_sumpool2: y = sum over 2x2x2 blocks of x -- backward of the nearest x2 upsample.               */
int ltu_zero_insert(const void* y, void* z, int B, int H, int W, int D, int C, int Hz, int Wz, int Dz,
                    int sh, int sw, int sd, int dtype, ltu_stream_t stream);
int ltu_sumpool2(const void* x, void* y, int B, int H, int W, int D, int C, int dtype, ltu_stream_t stream);

/* backward of the decoder glue (all gathers: bit-reproducible):
 *  ltu_upsample_trilinear_bwd: dy [B,2H,2W,fd*D,C] -> dx [B,H,W,D,C], the exact transpose of
 *    ltu_upsample_trilinear (nn.Upsample(trilinear, align_corners=True), Unet_3Dblock.py:1341-1345);
 *  ltu_mask_softmax_bwd: logits fp32 [B,V,Cout], dmask fp32 [B,Cout,V] -> dlogits fp32 [B,V,Cout]
 *    (torch.softmax(mask_conv(x), 1), :1380);
 *  ltu_head_d2s_softmax_bwd: logits fp32 [B,H2,W2,D,4*Cout], dprobs fp32 [B,Cout,2*H2,2*W2,D] ->
 *    dlogits like logits (windows_unembedding + F.softmax, :138-152,:1392-1394).                  */
int ltu_upsample_trilinear_bwd(const void* dy, void* dx, int B, int H, int W, int D, int C, int fd,
                               int dtype, ltu_stream_t stream);
int ltu_mask_softmax_bwd(const float* logits, const float* dmask, float* dlogits, int B,
                         int64_t voxels, int Cout, ltu_stream_t stream);
int ltu_head_d2s_softmax_bwd(const float* logits, const float* dprobs, float* dlogits, int B, int H2,
                             int W2, int D, int Cout, ltu_stream_t stream);

/* backward of ltu_gate_fused (SpatialAttention3DBlock + `encoded * attn`, Unet_3Dblock.py:217-221,:1385):
 * writes dskip = dout * gate (direct path), dh = the gradient of norm(W_x skip) AND of norm(W_g up)
 * (feed it to ltu_instnorm_bwd with LTU_ACT_NONE for each), dpsi_w fp32 [Ci], dpsi_b fp32 [1].
 * workspace: ltu_gate_bwd_workspace(B, voxels, Ci) bytes; ordered sums.                           */
size_t ltu_gate_bwd_workspace(int B, int64_t voxels, int Ci);
int ltu_gate_bwd(const void* a, const float* stats_a, const void* g, const float* stats_g,
                 const float* psi_w, const float* psi_b, const void* skip, const void* dout, void* dskip,
                 void* dh, float* dpsi_w, float* dpsi_b, void* workspace, size_t ws_bytes, int B,
                 int64_t voxels, int Ci, int dtype, ltu_stream_t stream);

/* backward of ltu_roi_resample (grid_sample of roi_alignment2 / post_processing2, Unet_3Dblock.py:
 * 985-1039,:1080-1117; the box is not differentiated, like the reference where it comes out of
 * searchsorted): dy has the forward's OUTPUT extent, dx its INPUT extent; same `direction`, ROI
 * constants and box as the forward call.  A gather over tabulated forward taps: exact transpose,
 * no atomics.  workspace: ltu_roi_resample_bwd_workspace(B, h, w, eval_h, eval_w) bytes.          */
size_t ltu_roi_resample_bwd_workspace(int B, int h, int w, int eval_h, int eval_w);
int ltu_roi_resample_bwd(const void* dy, const float* box, void* dx, void* workspace, size_t ws_bytes,
                         int B, int h, int w, int d, int C, int roi_h, int roi_w, int eval_h, int eval_w,
                         int direction, int dtype, ltu_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LTU_B200_H */
