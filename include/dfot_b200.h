/*
 * dfot_b200.h — C ABI of the B200-native DFoT denoising-sampling kernels (sm_100a).
 *
 * The reference (ktncktnc/diffusion-forcing-transformer) is pure PyTorch and has no FFI;
 * the boundary it exposes for this path is its Python module API.  These
 * entry points sit *beneath* that API: each one replaces a group of ATen calls made by the
 * reference functions cited next to it.  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a CUDA device pointer unless noted;
 *   - `stream` is a cudaStream_t passed as void*; kernels are launched on it, the call never
 *     synchronises, never allocates device memory and never throws;
 *   - return value: 0 = success, negative = error (DFOT_ERR_*); dfot_last_error() returns a
 *     thread-local, human-readable message for the last failing call;
 *   - bf16 tensors are raw uint16 storage; "f32" is IEEE float.
 */
#ifndef DFOT_B200_H_
#define DFOT_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFOT_ABI_VERSION 1

#if defined(__GNUC__)
#define DFOT_API __attribute__((visibility("default")))
#else
#define DFOT_API
#endif

#define DFOT_OK 0
#define DFOT_ERR_INVALID_ARG (-1)
#define DFOT_ERR_UNSUPPORTED (-2)
#define DFOT_ERR_CUDA (-3)
#define DFOT_ERR_DRIVER (-4)

/* dtype tags for arguments that accept either storage type */
#define DFOT_F32 0
#define DFOT_BF16 1
#define DFOT_I64 2

DFOT_API int dfot_abi_version(void);
DFOT_API const char* dfot_last_error(void);
/* number of kernel launches issued through this library since load (all threads) */
DFOT_API int64_t dfot_launch_count(void);
/* Latency mode (off by default; DFOT_LATENCY_MODE=1 turns it on at load).  By default every kernel choice that affects the
 * ARITHMETIC of a token row is independent of the batch, so a forward row is bit-identical whatever rows it is batched
 * with (sharded rollouts == single-GPU rollouts, bit for bit).  Latency mode trades that bit-level invariance (parity with
 * the reference is unchanged) for the latency of small-batch DiT sampling (<= ~1k token rows): K1 runs one block per row up
 * to 2048 rows, short attention problems that fit one wave as single query tiles use kernel 1, and the host takes the
 * split-K block loop (dfot_gemm_bf16_splitk + dfot_splitk_gate_resid_adaln) up to 1280 rows. */
DFOT_API int dfot_set_latency_mode(int on);
DFOT_API int dfot_get_latency_mode(void);

/* ------------------------------------------------------------------------------------------
 * K4 — fused per-frame sampler step + history-guidance combine (+ next-step prepare).
 * Replaces, per sampling step, the ATen elementwise chain of
 *   algorithms/dfot/dfot_video.py:675-752            (mask update, clone, revert `where`)
 *   algorithms/dfot/history_guidance.py:446-568, 929-982 (prepare / compose)
 *   algorithms/dfot/diffusion/discrete_diffusion.py:242-250 (q_sample), 454-538 (DDIM update)
 * in one pass over HBM.  All per-(row, frame) integers/scalars are precomputed on the host
 * for the whole window (they depend only on the scheduling matrix and the context mask).
 *
 * Layout: a "frame" is F = C*H*W contiguous values.  x: [B, T, F] f32 state (in/out).
 * Branch rows r = (b, j), j in [0, nfe) ordered as the reference's "(b h g)".
 *   model_out     [B*nfe, T, F]  backbone output of THIS step (NULL => prepare-only launch)
 *   model_in_next [B*nfe, T, F]  backbone input for the NEXT step (NULL => do not emit)
 *   upd  [B*nfe, T] dfot_frame_update  : DDIM/compose coefficients of this step
 *   prep [B*nfe, T] dfot_frame_prepare : how to build the next step's input per branch frame
 *   noise_ddim [B*nfe, T, F] f32 or NULL (only read where upd.sigma != 0)
 *   noise_hist [B*n_hist, T, F] f32 or NULL; prep.noise_row selects the row
 *   noise_excl [B*nfe, T, F] f32 or NULL
 * For every frame with upd(b,0,t).generate != 0:
 *     x' = sum_j w_j * ( a_j * x + b_j * g(out_j) + sigma_j * n_j ),  g = clamp(+-clip) or identity
 * else x' = x.  Then, per branch frame: mode 0: in = x'; 1: in = qa*x' + qb*noise_hist; 2: in = noise_excl.
 */
typedef struct {
  float a;        /* coefficient of x_t            */
  float b;        /* coefficient of g(model_out)   */
  float sigma;    /* coefficient of DDIM noise     */
  float w;        /* compose weight (0 drops the branch for this frame) */
  float clip;     /* >0: clamp model_out to +-clip before use (pred_noise) */
  int32_t generate; /* 1: frame is being generated (context_mask == 0): x is overwritten */
} dfot_frame_update;

typedef struct {
  int32_t mode;      /* 0 copy, 1 q_sample re-noise, 2 pure noise (excluded gen token) */
  int32_t noise_row; /* row of noise_hist to read in mode 1 */
  float qa;          /* sqrt(alpha_bar[level])     */
  float qb;          /* sqrt(1 - alpha_bar[level]) */
} dfot_frame_prepare;

DFOT_API int dfot_sampler_step_hg(float* x, const void* model_out, int model_out_dtype, void* model_in_next,
                         int model_in_dtype, const dfot_frame_update* upd, const dfot_frame_prepare* prep,
                         const float* noise_ddim, const float* noise_hist, const float* noise_excl,
                         int64_t B, int64_t nfe, int64_t T, int64_t F, void* stream);

/* ------------------------------------------------------------------------------------------
 * K1 — fused AdaLN modulate + LayerNorm with per-frame gather.
 * Replaces algorithms/dfot/backbones/dit/dit_blocks.py:15-16, 378-437 (AdaLayerNorm[Zero]):
 *   y[m, :] = LN_eps(x[m, :]) * (1 + scale[f(m), :]) + shift[f(m), :],   f(m) = m / tokens_per_frame
 * shift/scale are read from the per-frame modulation matrix mod[f, shift_col + d] / mod[f, scale_col + d]
 * (the reference recomputes them per token; only the frame varies).  Writes y as f32 (residual
 * base, may be NULL) and/or bf16 (GEMM operand, may be NULL).  D % 8 == 0, D <= 8192.
 */
DFOT_API int dfot_adaln_layernorm(const float* x, const float* mod, int64_t mod_ld, int64_t shift_col, int64_t scale_col,
                         float* y_f32, void* y_bf16, int64_t M, int64_t D, int64_t tokens_per_frame, float eps,
                         void* stream);
/* The same with a side output: stats[m] = (mean, rstd) of row m (float pairs, may be NULL).  With it the consumer of the
 * fp32 copy — the gated-residual GEMM epilogue of a DiT block half, whose residual base is the MODULATED tensor
 * (dit_blocks.py:504-509) — can rebuild y from x (DFOT_EPI_GATE_LNRESID_F32) and y_f32 need not be written. */
DFOT_API int dfot_adaln_layernorm_stats(const float* x, const float* mod, int64_t mod_ld, int64_t shift_col, int64_t scale_col,
                               float* y_f32, void* y_bf16, float* stats, int64_t M, int64_t D, int64_t tokens_per_frame,
                               float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * K2 — tcgen05/TMEM bf16 GEMM fed by TMA:  C[M,N] = epilogue(A[M,K] · W[N,K]^T + bias[N]).
 * Replaces every nn.Linear on the path (dit_blocks.py:72,76,387,417,525; timm Mlp; the
 * patch-embed conv dit3d.py:49-55; diffusers TimestepEmbedding): A row-major bf16 (lda elements),
 * W = nn.Linear.weight row-major bf16 (ldw elements), fp32 accumulation in tensor memory.
 * K % 8 == 0, lda % 8 == 0, ldw % 8 == 0 (16-byte TMA strides), N % 8 == 0.
 */
#define DFOT_EPI_F32 0            /* out f32  = acc + bias                                   */
#define DFOT_EPI_BF16 1           /* out bf16 = acc + bias                                   */
#define DFOT_EPI_GELU_BF16 2      /* out bf16 = gelu_tanh(acc + bias)      (timm Mlp fc1)     */
#define DFOT_EPI_SILU_BF16 3      /* out bf16 = silu(acc + bias)           (TimestepEmbedding) */
#define DFOT_EPI_GATE_RESID_F32 4 /* out f32  = resid + gate[f(m), n] * (acc + bias)          */
#define DFOT_EPI_QKV_ROPE_BF16 5  /* out bf16 = rope3d(acc + bias) on q,k columns; q pre-scaled */
#define DFOT_EPI_RESID_F32 6      /* out f32  = resid + acc + bias         (U-ViT residual adds) */
#define DFOT_EPI_QKNORM_ROPE_BF16 7 /* out bf16 = [rope3d(rmsnorm_d(acc + bias) * w_q) * q_scale | rope3d(rmsnorm_d(.) * w_k) | v]:
                                     q/k RMSNorm over head_dim from the fp32 accumulators, then RoPE-3D
                                     (u_vit_blocks.py:253-259); N = 3*model_dim, head_dim in {64, 128} */

#define DFOT_EPI_GATE_LNRESID_F32 8 /* out f32 = y[m, n] + gate[f(m), n] * (acc + bias), y = ((resid - mean[m]) * rstd[m]) *
                                     (1 + ln_scale[f(m), n]) + ln_shift[f(m), n]: the residual base of a DiT block half rebuilt
                                     from x and the (mean, rstd) side output of dfot_adaln_layernorm_stats, bit for bit the
                                     value K1 computes; out may be resid itself (dit_blocks.py:427-437, 504-509) */

typedef struct {
  const float* bias;          /* [N] or NULL */
  /* GATE_RESID */
  const float* resid;         /* [M, ld_resid] f32 */
  int64_t ld_resid;
  const float* gate;          /* gate[f * ld_gate + n] */
  int64_t ld_gate;
  int64_t tokens_per_frame;   /* f(m) = m / tokens_per_frame */
  /* QKV_ROPE: columns [0, D) = q, [D, 2D) = k, [2D, 3D) = v; head h owns columns h*head_dim.. */
  const float* rope_cs;       /* [tokens_per_sample, head_dim/2, 2] = (cos, sin) of the pair angle */
  int64_t tokens_per_sample;  /* token index = m % tokens_per_sample */
  int64_t model_dim;          /* D */
  int64_t head_dim;
  float q_scale;              /* multiplies q after rotation (softmax scale * log2 e) */
  /* QKNORM_ROPE: RMSNorm weights of q and k over head_dim, eps */
  const float* qn_w;
  const float* kn_w;
  float qk_eps;
  /* optional side output of the F32 / RESID_F32 / BF16 epilogues: GroupNorm statistics of the OUTPUT (as stored),
     so the next GroupNorm needs no extra pass over HBM.  gn_sums: the 3*n_img*groups-double workspace of
     dfot_groupnorm_stats (zeroed, accumulated and finalised to (mean, rstd) by the call); image = m / gn_rows_per_img
     (gn_rows_per_img % 32 == 0), group = n / (N / gn_groups) (power-of-two channels per group). NULL = off */
  double* gn_sums;
  int64_t gn_rows_per_img;
  int64_t gn_groups;
  float gn_eps;
  /* GATE_LNRESID: stats[m] = (mean, rstd) float pairs; shift / scale vectors addressed like the gate (f * ld_gate + n) */
  const float* ln_stats;
  const float* ln_shift;
  const float* ln_scale;
} dfot_gemm_epilogue;

DFOT_API int dfot_gemm_bf16(const void* A, int64_t lda, const void* W, int64_t ldw, void* C, int64_t ldc, int64_t M,
                   int64_t N, int64_t K, int epilogue, const dfot_gemm_epilogue* epi, void* stream);

/* 3x3 convolution (stride 1, zero padding 1) as an implicit GEMM on the same tcgen05 kernel — no im2col buffer:
 * the A tile of tap (dy, dx) is the pixel tile's 4-D TMA box shifted by (dy, dx), out-of-image rows/columns are
 * zero-filled by the TMA unit.  Replaces the nn.Conv2d(k=3, padding=1) of the U-ViT ResBlock / Downsample /
 * Upsample (algorithms/dfot/backbones/u_vit/u_vit_blocks.py:63-75, 277-314).
 *   x   [n_img, H, W, Cin]  bf16 channel-last;   w [Cout, 3, 3, Cin] bf16 (torch weight permuted (0, 2, 3, 1));
 *   out [n_img*H*W, ldc]    f32 or bf16 per epilogue (F32, BF16, SILU_BF16, RESID_F32), channel-last.
 * W must be a power of two below 128 or a multiple of 128; H a multiple of 128/W (or a power of two when H*W < 128).
 */
DFOT_API int dfot_conv3x3_bf16(const void* x, const void* w, void* out, int64_t ldc, int64_t n_img, int64_t H,
                      int64_t W, int64_t Cin, int64_t Cout, int epilogue, const dfot_gemm_epilogue* epi,
                      void* stream);

/* Causal kt x 3 x 3 convolution along a frame axis on the same kernel (K = kt*9 taps x Cin; the temporal tap is a
 * frame offset of the 4-D TMA box) — the PaddedConv3D of the reference's causal VideoVAE
 * (algorithms/vae/common/modules/conv.py:39-108: the first frame repeated kt-1 times in front, no other temporal
 * padding; spatial zero padding 1).  First building block of the VAE-decode row.
 *   x   [n_frames_out + kt - 1, H, W, Cin] bf16 channel-last, laid out by the caller so that frames j .. j+kt-1 are
 *       output frame j's causal window (per clip: kt-1 copies of its first frame, then its frames);
 *   w   [Cout, kt, 3, 3, Cin] bf16 (torch Conv3d weight permuted (0, 2, 3, 4, 1));
 *   out [n_frames_out*H*W, ldc], output frame j = sum_dt conv3x3(x[j + dt], w[:, dt]).  Same epilogues and size
 *       rules as dfot_conv3x3_bf16; kt = 1 is that function.
 */
DFOT_API int dfot_conv3d_causal_bf16(const void* x, const void* w, void* out, int64_t ldc, int64_t n_frames_out,
                            int64_t H, int64_t W, int64_t Cin, int64_t Cout, int64_t kt, int epilogue,
                            const dfot_gemm_epilogue* epi, void* stream);

/* Split-K pair for the latency regime of K2 (a few hundred token rows — small-batch DiT sampling: fewer output tiles than
 * SMs and a long serial k-loop per CTA, bound by one SM's TMA fill rate):
 *   dfot_gemm_bf16_splitk: parts[M, splits*N] f32, columns [s*N, (s+1)*N) = A[:, K_s] · W[:, K_s]^T for the s-th share of
 *       the k-blocks (no bias, no atomics).  N % 64 == 0, splits <= ceil(K / 64); same alignment rules as dfot_gemm_bf16.
 *   dfot_splitk_gate_resid_adaln: the consumer, fused with the AdaLN that follows a DiT block half
 *       (dit_blocks.py:504-509 then :427-437):
 *         x[m, :] = resid[m, :] + mod[f(m), gate_col + :] * (sum_s parts[m, s*D + :] + bias)     (fixed split order)
 *         y[m, :] = LN_eps(x[m, :]) * (1 + mod[f(m), scale_col + :]) + mod[f(m), shift_col + :]  -> y_f32 and / or y_bf16
 *       x_out (may be NULL) receives x; shift_col < 0: no norm (x_out is the output).  Outputs must not alias resid.
 *       D % 4 == 0, D <= 2048.
 */
DFOT_API int dfot_gemm_bf16_splitk(const void* A, int64_t lda, const void* W, int64_t ldw, float* parts, int64_t M, int64_t N,
                          int64_t K, int64_t splits, void* stream);
DFOT_API int dfot_splitk_gate_resid_adaln(const float* parts, int64_t splits, const float* bias, const float* resid,
                                 const float* mod, int64_t mod_ld, int64_t gate_col, int64_t shift_col, int64_t scale_col,
                                 float* x_out, float* y_f32, void* y_bf16, int64_t M, int64_t D, int64_t tokens_per_frame,
                                 float eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * K3 — attention over space-time latent tokens (non-causal, no mask: context frames are
 * expressed through per-frame noise levels, never an attention mask — SURVEY.md §8a Q9).
 * Replaces dit_blocks.py:21-44, 100-123.  qkv: [R*Ntok, 3*D] bf16 as written by the
 * QKV_ROPE epilogue (q already rotated and multiplied by scale*log2e, k rotated);
 * out: [R*Ntok, D] bf16 (heads concatenated, i.e. "transpose(1,2).reshape(B,N,C)").
 * Online softmax in fp32 (exp2), row max / sum via warp shuffles.  head_dim in {64, 72, 128}.
 */
DFOT_API int dfot_attention(const void* qkv, void* out, int64_t R, int64_t Ntok, int64_t heads, int64_t head_dim,
                   void* stream);
/* same, with an explicit row stride for `out` (elements): the U-ViT block writes the attention output straight into
   the [attention | SiLU(mlp_h)] operand of its fused attn_out + mlp_out GEMM (u_vit_blocks.py:262-271) */
DFOT_API int dfot_attention_strided(const void* qkv, void* out, int64_t ld_out, int64_t R, int64_t Ntok, int64_t heads,
                           int64_t head_dim, void* stream);
/* same, for BOUNDED scores: `score_bound` >= |q.k| of the pre-scaled operands (log2 units).  With QK-normalisation
   (u_vit_blocks.py:253-254) |q.k|*scale*log2e <= sqrt(head_dim)*max|w_q|*max|w_k|*log2e, typically ~12: the softmax
   then needs no running maximum at all — p = 2^s, out = (sum p v) / (sum p) is the identical function — which removes
   the max pass, the subtraction and every O rescale from the MUFU-bound inner loop.  0 < score_bound <= 96 selects the
   bounded path (N > 128); 0 = unknown (online softmax with a running maximum). */
DFOT_API int dfot_attention_bounded(const void* qkv, void* out, int64_t ld_out, int64_t R, int64_t Ntok, int64_t heads,
                           int64_t head_dim, float score_bound, void* stream);

/* ------------------------------------------------------------------------------------------
 * Small glue kernels of the DiT3D backbone (dit3d.py:153-192, embeddings.py:67-153).
 */
/* noise-level features: sinusoidal [cos|sin] (levels int64 or f32) or Fourier cos(k*f+phase)*sqrt2 → bf16 [n, dim] */
DFOT_API int dfot_noise_features(const void* levels, int levels_dtype, const float* fourier_freqs,
                        const float* fourier_phases, void* out_bf16, int64_t n, int64_t dim, void* stream);
/* c_act = silu(a + (row_mask[r] ? 0 : b)) → bf16; a,b f32 [n_rows, D]; b / row_mask may be NULL; rows_per_mask = T */
DFOT_API int dfot_silu_sum_bf16(const float* a, const float* b, const uint8_t* row_mask, int64_t rows_per_mask,
                       void* out_bf16, int64_t n_rows, int64_t D, void* stream);
/* patchify: x [R*T, C, H, W] (f32 or bf16) → tokens [R*T*(H/p)*(W/p), ld] bf16, columns (c, py, px) = conv-weight order;
   columns >= C*p*p are left untouched (the caller zero-fills the K padding once) */
DFOT_API int dfot_patchify_bf16(const void* x, int x_dtype, void* out_bf16, int64_t ld, int64_t frames, int64_t C, int64_t H,
                       int64_t W, int64_t p, void* stream);
/* unpatchify: tokens [frames*(H/p)*(W/p), ld] f32 with columns (py, px, c) → x [frames, C, H, W] (f32 or bf16) */
DFOT_API int dfot_unpatchify(const float* tok, int64_t ld, void* x, int x_dtype, int64_t frames, int64_t C, int64_t H,
                    int64_t W, int64_t p, void* stream);
/* elementwise f32 → bf16 (weights repack / activations), n % 8 == 0 not required */
DFOT_API int dfot_cast_bf16(const float* in, void* out_bf16, int64_t n, void* stream);

/* Matrix-attention DiT variants (full_matrix_attention / factorized_matrix_attention; dit_blocks.py:211-350
 * MatrixAttention, :549-652 MatrixDiTBlock): the attention tokens are FRAMES, a frame's [P patches, D] matrix is
 * projected as u^T X v.  The u factors (over the patches) run as the two kernels below, the v factors as K2 GEMMs on
 * R*Mc*L rows, the attention over the L frames as K3 (Mc = embed_col_dim column heads of one row each).
 *   dfot_patch_mix_bf16:  out[((r*Mc + c)*L + l), d] = sum_n u[n*Mc + c] * y[((r*L + l)*P + n), d]     f32 -> bf16
 *       (replaces the `qkv_u` contraction of matrix_mul, dit_blocks.py:211-212,301; y = the block's modulated tokens)
 *   dfot_patch_expand_gate_resid:  x[((r*L + l)*P + n), d] = y[same] + gate[(r*L + l)*ld_gate + d] *
 *           (sum_c pu[c*P + n] * z[((r*Mc + c)*L + l), d] + pb[n*D + d])                                 f32
 *       (replaces the `proj_u` contraction + proj_bias, dit_blocks.py:345-347, and the gated residual :640-644;
 *        z = attention output through proj_v; pb may be NULL; y NULL = no residual and gate NULL = gate 1 (the bare attention
 *        output a MatrixCrossDiTBlock attends to, :752-761); x must not alias y).  D % 4 == 0.
 */
DFOT_API int dfot_patch_mix_bf16(const float* y, const float* u, void* out_bf16, int64_t R, int64_t L, int64_t P,
                        int64_t Mc, int64_t D, void* stream);
DFOT_API int dfot_patch_expand_gate_resid(float* x, const float* y, const float* z, const float* pu, const float* pb,
                                 const float* gate, int64_t ld_gate, int64_t R, int64_t L, int64_t P, int64_t Mc,
                                 int64_t D, void* stream);

/* ------------------------------------------------------------------------------------------
 * U-ViT3DPose glue kernels (algorithms/dfot/backbones/u_vit/{u_vit3d_pose,u_vit3d,u_vit_blocks}.py).
 * Activations are channel-last everywhere: [n_img, H, W, C] == [tokens, C], so ResBlock levels and
 * transformer levels share one layout and the reference's rearranges (u_vit3d.py:199-235) vanish.
 *
 * FiLM inputs: scale/shift = per-image f32 part  mod_img[img*ld_img + {scale,shift}_col + c]   (noise-level
 * embedding through the block's emb_layer, one GEMM per forward for all blocks) plus an optional per-pixel
 * bf16 part  mod_pix[(img_map[img]*HW + pix) * 2C + {0, C} + c]  (camera-pose embedding through the same
 * emb_layer — linear, hence constant over all sampling steps and cached per window; img_map[img] < 0 means
 * the row's pose embedding is masked out, embeddings.py:336-361).
 */
/* GroupNorm statistics (u_vit_blocks.py:52-53: 32 groups): sums[img, g] = (sum, sum of squares) over H*W x C/G,
   accumulated as 64-bit FIXED-POINT integers (2^-32 units; integer atomics are associative, so the statistics are
   deterministic and do not depend on how many images share a launch), then finalised to f32 (mean, rstd) pairs stored
   right behind the sums: the workspace `sums` must hold 3*n_img*groups 8-byte words (declared double for alignment).
   The buffer is zeroed by the call (cudaMemsetAsync on the stream). x f32 or bf16; C/groups divides 8 or is a multiple. */
DFOT_API int dfot_groupnorm_stats(const void* x, int x_dtype, double* sums, int64_t n_img, int64_t HW, int64_t C,
                         int64_t groups, float eps, void* stream);
/* y = silu( GN(x) * gamma + beta [ * (1 + scale) + shift ] ) -> bf16 (the conv operand).  mod_img/mod_pix NULL: no FiLM */
DFOT_API int dfot_groupnorm_silu_bf16(const void* x, int x_dtype, const double* sums, const float* gamma,
                             const float* beta, const float* mod_img, int64_t ld_img, int64_t scale_col,
                             int64_t shift_col, const void* mod_pix, const int32_t* img_map, void* y_bf16,
                             int64_t n_img, int64_t HW, int64_t C, int64_t groups, void* stream);

/* ------------------------------------------------------------------------------------------
 * VAE-decode row: building blocks of the reference's causal VideoVAE decoder
 * (algorithms/vae/video_vae/model.py:153-270).  Clips are channel-last with a padded frame axis
 * [B, 2 + T, H, W, C]: the two leading slots of a clip hold copies of its first frame in bf16 conv inputs (the causal
 * window of dfot_conv3d_causal_bf16) and are unused in the fp32 stream.
 *   dfot_groupnorm_stats_strided / dfot_groupnorm_apply_bf16 — Normalize = GroupNorm(32, eps 1e-6) over a clip's valid
 *     frames (normalize.py:4-7): "image" i starts at x + i*img_stride and has HW rows of C channels; apply writes
 *     bf16 at the same offsets of y, with or without the x*sigmoid(x) nonlinearity (ops.py:21-22).
 *   dfot_vae_upsample2x_bf16 — temporal = 0: nearest x2 in (H, W) (SpatialUpsample2x, updownsample.py:73-80);
 *     temporal = 1: first frame bilinear x2, the others trilinear x(2,2,2), align_corners = False
 *     (Spatial2xTime2x3DUpsample, updownsample.py:131-147).  fp32 clip in -> bf16 clip out (pads filled).
 *   dfot_vae_fill_pad_frames — pad slots of a bf16 clip <- its first frame (PaddedConv3D.forward, conv.py:98-104).
 *   dfot_softmax_rows_bf16 — softmax(scale * s) per row, fp32 -> bf16 (AttnBlock3D, attention.py:138-140).
 */
DFOT_API int dfot_groupnorm_stats_strided(const void* x, int x_dtype, double* sums, int64_t n_img, int64_t HW,
                                 int64_t img_stride, int64_t C, int64_t groups, float eps, void* stream);
DFOT_API int dfot_groupnorm_apply_bf16(const float* x, const double* sums, const float* gamma, const float* beta,
                              void* y_bf16, int64_t n_img, int64_t HW, int64_t img_stride, int64_t C,
                              int64_t groups, int silu, void* stream);
DFOT_API int dfot_vae_upsample2x_bf16(const float* in, void* out_bf16, int64_t B, int64_t T_in, int64_t H, int64_t W,
                             int64_t C, int temporal, void* stream);
/* nearest x2 of a plain image batch, fp32 [n_img, H, W, C] -> bf16 [n_img, 2H, 2W, C]: the interpolation of `Upsample`
   in the reference's ImageVAE decoder (algorithms/vae/common/modules/updownsample.py:10-24) */
DFOT_API int dfot_upsample2x_nearest_bf16(const float* in, void* out_bf16, int64_t n_img, int64_t H, int64_t W, int64_t C,
                                 void* stream);
DFOT_API int dfot_vae_fill_pad_frames(void* x_bf16, int64_t B, int64_t T, int64_t frame_elems, void* stream);
DFOT_API int dfot_softmax_rows_bf16(const float* s, int64_t ld_s, void* p_bf16, int64_t ld_p, int64_t rows, int64_t n,
                           float scale, void* stream);

/* NormalizeWithCond (u_vit_blocks.py:98-121): y = RMSNorm(x) * weight * (1 + scale) + shift -> bf16; x [M, D] f32,
   image = m / tokens_per_img */
DFOT_API int dfot_rmsnorm_film_bf16(const float* x, const float* weight, float eps, const float* mod_img, int64_t ld_img,
                           int64_t scale_col, int64_t shift_col, const void* mod_pix, const int32_t* img_map,
                           void* y_bf16, int64_t M, int64_t D, int64_t tokens_per_img, void* stream);
/* q/k RMSNorm over head_dim (* weight) then RoPE-3D, q additionally * q_scale; in place on qkv [M, ld] bf16 with
   columns [q | k | v] of `heads*head_dim` each (u_vit_blocks.py:253-259).  head_dim in {64, 128} */
DFOT_API int dfot_qk_norm_rope(void* qkv, int64_t ld, const float* q_weight, const float* k_weight, float eps,
                      const float* rope_cs, int64_t tokens_per_sample, int64_t M, int64_t heads, int64_t head_dim,
                      float q_scale, void* stream);
/* 2x2 average pooling, channel-last: in [n_img, H, W, C] -> out [n_img, H/2, W/2, C]; dtypes f32|bf16 each */
DFOT_API int dfot_avgpool2x2(const void* in, int in_dtype, void* out, int out_dtype, int64_t n_img, int64_t H, int64_t W,
                    int64_t C, void* stream);
/* out_bf16 = a - b (f32 inputs): the operand of the Upsample conv (u_vit3d_pose.py:125) */
DFOT_API int dfot_sub_bf16(const float* a, const float* b, void* out_bf16, int64_t n, void* stream);
/* out[n, Y, X, c] = low[n, Y/2, X/2, c] + skip[n, Y, X, c]  (nearest 2x upsampling + skip, u_vit_blocks.py:299-314) */
DFOT_API int dfot_upsample2x_add(const float* low, const float* skip, float* out, int64_t n_img, int64_t H, int64_t W,
                        int64_t C, void* stream);
/* Camera rays -> sinusoidal ray encoding -> patch rows for the pose PatchEmbed GEMM, never materialising the
   [frames, 180, H, W] tensor (utils/geometry_utils.py:50-81, 244-295; dfot_video_pose.py:64-110).
   cams [frames, 16] f32 = (fx, fy, px, py) already multiplied by the resolution, R^-1 row-major (9), origin (3);
   out [frames*(res/p)^2, ld] bf16 with columns ((py*p + px)*6*2*n_freq + channel), channel order as the reference:
   [origin: sin(v*2^s*pi) (i s) | sin(. + pi/2) | direction: ... ].  freq_scale [n_freq] f32. */
DFOT_API int dfot_pose_ray_patches(const float* cams, const float* freq_scale, int64_t n_freq, void* out_bf16, int64_t ld,
                          int64_t frames, int64_t res, int64_t p, void* stream);

/* ------------------------------------------------------------------------------------------
 * DC-AE image decoder (the VAE of the DMLab / Minecraft latent configurations): glue kernels around
 * the GEMM / implicit-GEMM calls.  Reference: algorithms/vae/dc_ae/autoencoder_dc_model.py (Decoder :383-470,
 * DCUpBlock2d :222-260, ResBlock :109-136, EfficientViTBlock :139-172, SanaMultiscaleLinearAttention :46-106) and the
 * diffusers==0.32.2 modules it imports (GLUMBConv, RMSNorm, SanaMultiscaleAttnProcessor2_0).  Channel-last [n, H, W, C].
 *   dfot_relu_bf16            — max(x, 0) in place (ResBlock nonlinearity; n % 8 == 0).
 *   dfot_pixel_shuffle2x      — out[n, 2h+i, 2w+j, c] = conv[n, h, w, 4c+2i+j] (+ x[n, h, w, (4c+2i+j)/repeats], the
 *                               channel-repeat shortcut of DCUpBlock2d; x NULL: none) -> f32 and / or bf16.
 *   dfot_linear_attention_relu— ReLU linear attention over the tokens of an image: a token's [q|k|v] row is read as
 *                               `heads` groups of 3*head_dim channels (query, key, value of the group), per group
 *                               out = (relu(q) KV) / (relu(q) KV[:, d] + eps), KV = relu(k)^T [v | 1]; fp32, head_dim <= 32.
 *   dfot_dwconv3x3_glu_bf16   — GLUMBConv middle: y = depthwise3x3(x) + b over 2*Ch channels; out = y[:Ch] * silu(y[Ch:]).
 *   dfot_rmsnorm_affine       — y = x * rsqrt(mean_c(x^2) + eps) * w + b (+ resid) (ReLU if relu) -> f32 and / or bf16.
 */
DFOT_API int dfot_relu_bf16(void* x_bf16, int64_t n, void* stream);
DFOT_API int dfot_pixel_shuffle2x(const float* conv, int64_t ld_conv, const float* x, int64_t Cx, int64_t repeats,
                                  float* out_f32, void* out_bf16, int64_t n_img, int64_t H, int64_t W, int64_t C,
                                  void* stream);
DFOT_API int dfot_linear_attention_relu(const float* qkv, int64_t ld_qkv, float* out, int64_t ld_out, int64_t n_img,
                                        int64_t HW, int64_t heads, int64_t head_dim, float eps, void* stream);
DFOT_API int dfot_dwconv3x3_glu_bf16(const void* x_bf16, const float* w, const float* b, void* out_bf16, int64_t n_img,
                                     int64_t H, int64_t W, int64_t Ch, void* stream);
DFOT_API int dfot_rmsnorm_affine(const float* x, int64_t ld_x, const float* w, const float* b, float eps,
                                 const float* resid, int relu, float* out_f32, void* out_bf16, int64_t M, int64_t C,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DFOT_B200_H_ */
