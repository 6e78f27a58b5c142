/*
 * videogpa_b200 — C-ABI of the B200-native denoise-and-score hot path of VideoGPA.
 *
 * The reference (Hongyang-Du/VideoGPA) has no native code and no FFI: its seams are Python
 * callables. Each entry point below replaces the arithmetic behind one of those
 * callables; the Python mirror in videogpa_b200/ binds them with ctypes and keeps the reference's
 * signatures. INTEGRATION.md shows the binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; vgpa_last_error() gives the reason
 *     (thread-local, valid until the next failing call on that thread). Nothing aborts the process:
 *     the reference's CLIs wrap each item in try/except and continue (generate/CogVideoX-5B.py:69-80).
 *   - pointers named d_* are DEVICE pointers on the current CUDA device, h_* are HOST pointers.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream). Device-pointer entry
 *     points only enqueue work on that stream; they never synchronise unless stated.
 *   - bf16 tensors are raw uint16 storage (IEEE bfloat16), row-major, innermost dimension contiguous.
 */
#ifndef VIDEOGPA_B200_H
#define VIDEOGPA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VGPA_ABI_VERSION 4 /* 2: vgpa_attention_args gained `lse`; backward / training, T5 and transpose entry points added
                              3: vgpa_attention_args gained `workspace` (bounded-softmax fast path of the head_dim-64 forward)
                              4: vgpa_conv3d_args gained the fused GroupNorm-statistics outputs (gn_*) */

/* ------------------------------------------------------------------------------------------------
 * runtime
 * ---------------------------------------------------------------------------------------------- */
const char* vgpa_last_error(void);
int vgpa_abi_version(void);
int vgpa_device_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * K2 — DiT linears: out = epilogue(A[M,K] . W[N,K]^T), bf16 in, fp32 accumulate (tcgen05), bf16 out.
 * Replaces torch.nn.Linear (cuBLAS) inside diffusers' CogVideoXBlock: attn1.to_q/to_k/to_v/to_out.0
 * and ff.net.0.proj / ff.net.2 (reference call site
 * generate/CogVideoX-5B.py:72-77, train/CogVideoX-5B/03_train.py:134-151).
 * ---------------------------------------------------------------------------------------------- */
enum {
  VGPA_EPI_BIAS = 0,      /* out = acc + bias                                                      */
  VGPA_EPI_BIAS_GELU = 1, /* out = gelu_tanh(acc + bias)              (FeedForward net.0)          */
  VGPA_EPI_GATE_RES = 2,  /* out = out + gate[b, seg(row)] * (acc + bias)   (adaLN-zero residual)  */
  VGPA_EPI_QKV = 3,       /* fused to_q|to_k|to_v: per-head LayerNorm(64) on q,k + 3-D RoPE on     */
                          /* video rows; v passes through      (CogVideoXAttnProcessor2_0)         */
  VGPA_EPI_ACCUM = 4,     /* out = bf16(out + alpha * acc), one rounding (PEFT LoRA merge,         */
                          /* generate/CogVideoX-5B.py:29-30: W <- W + (lora_alpha/r) * B @ A)      */
  VGPA_EPI_GATE_RES_F32 = 5 /* fp32 residual stream: out is float [M, ldo];                        */
                          /* out = out + gate[b, seg(row)] * bf16(acc + bias) in fp32 (WanModel    */
                          /* under torch.autocast(bf16): bf16 linear output added to an fp32 x,    */
                          /* generate/Wan2.2-TI2V-5B.py:120-129)                                   */
};

typedef struct vgpa_linear_args {
  const void* A;   /* [M, lda] bf16 activations                                                    */
  const void* W;   /* [N, K]   bf16 weight, torch Linear layout (out_features, in_features)        */
  const void* bias; /* [N] bf16 or NULL                                                            */
  void* out;       /* [M, ldo] bf16; read-modify-write for VGPA_EPI_GATE_RES                       */
  int32_t M, N, K, lda, ldo;
  int32_t epilogue;
  /* rows are grouped per sample as [text_rows | video rows]; rows_per_sample = 0 means one sample */
  int32_t rows_per_sample;
  int32_t text_rows;
  /* VGPA_EPI_GATE_RES: gate vectors [N] per sample, batch stride in elements; both NULL = gate 1  */
  const void* gate_txt;
  const void* gate_vid;
  int64_t gate_stride_b;
  /* VGPA_EPI_QKV */
  const float* ln_q_w; /* [64] fp32 */
  const float* ln_q_b;
  const float* ln_k_w;
  const float* ln_k_b;
  float ln_eps;
  const float* rope_cos; /* [video rows, 64] fp32, repeat-interleaved pairs, or NULL (training    */
  const float* rope_sin; /*   step passes no image_rotary_emb, 03_train.py:134-139)               */
  int32_t model_dim;     /* D: q cols [0,D), k cols [D,2D), v cols [2D,3D); head_dim is 64        */
  float alpha;           /* VGPA_EPI_ACCUM scale                                                  */
} vgpa_linear_args;

int vgpa_linear_bf16(const vgpa_linear_args* args, void* stream);

/* ------------------------------------------------------------------------------------------------
 * K1 — joint text+video full attention: out = softmax(q k^T * scale) v per (batch, head), bf16 in/out,
 * fp32 softmax, head_dim 64 (CogVideoX) or 128 (Wan2.2). Replaces F.scaled_dot_product_attention inside diffusers'
 * CogVideoXAttnProcessor2_0.__call__ (reference call sites
 * generate/CogVideoX-5B.py:72-77, train/CogVideoX-5B/03_train.py:134-151). No mask, no dropout,
 * not causal. Heads are packed along the row: head h of row s lives at columns [h*head_dim, (h+1)*head_dim).
 * q/k/v may alias one fused [B*S, 3*H*64] projection buffer (pass three pointers into it).
 * ---------------------------------------------------------------------------------------------- */
typedef struct vgpa_attention_args {
  const void* q;   /* [B, Sq,  >=H*64] bf16 */
  const void* k;   /* [B, Skv, >=H*64] bf16 */
  const void* v;   /* [B, Skv, >=H*64] bf16 */
  void* out;       /* [B, Sq,  >=H*64] bf16 */
  int32_t B, H, Sq, Skv, head_dim;
  float scale;     /* softmax scale; <= 0 means 1/sqrt(head_dim) */
  int64_t q_row_stride, k_row_stride, v_row_stride, out_row_stride;         /* elements */
  int64_t q_batch_stride, k_batch_stride, v_batch_stride, out_batch_stride; /* elements */
  float* lse;      /* optional [B, H, Sq] fp32: log2-domain logsumexp of the scaled scores, saved for
                      vgpa_attention_bwd_bf16 (head_dim 64 only); NULL = not written */
  /* Optional device scratch of vgpa_attention_workspace_bytes(B, H, head_dim) bytes, 16-byte aligned. With it the
   * head_dim-64 forward first measures max|q|, max|k| per (batch, head) and serves every head whose score bound
   * max|q| max|k| scale log2(e) is <= 90 with the bounded-softmax kernel (softmax shifted by a fixed per-head offset
   * instead of the running row maximum: same function, no row statistics; attention_d64b_sm100.cu); the remaining
   * heads, and every head when workspace is NULL, take the exact online-softmax kernel. */
  void* workspace;
  size_t workspace_bytes;
} vgpa_attention_args;

size_t vgpa_attention_workspace_bytes(int B, int H, int head_dim);
int vgpa_attention_bf16(const vgpa_attention_args* args, void* stream);

/* Backward of vgpa_attention_bf16 (head_dim 64) — the DPO training step differentiates through
 * F.scaled_dot_product_attention (train/CogVideoX-5B/03_train.py:134-157).
 * Given q, k, v, the forward output `out`, its logsumexp `lse` (written by the forward call) and d_out:
 *   P = exp2(scale*log2e * q k^T - lse),  delta = rowsum(d_out * out),  dS = P * (d_out v^T - delta) * scale,
 *   dq = dS k,  dk = dS^T q,  dv = P^T d_out      (bf16 outputs, fp32 accumulation, deterministic).
 * workspace: vgpa_attention_bwd_workspace_bytes(B, H, Sq) bytes of device scratch (delta). */
typedef struct vgpa_attention_bwd_args {
  const void* q;
  const void* k;
  const void* v;
  const void* out;    /* forward output [B, Sq, >=H*64] */
  const void* d_out;  /* gradient of the output, same shape */
  const float* lse;   /* [B, H, Sq] from the forward call */
  void* dq;           /* [B, Sq,  >=H*64] bf16 */
  void* dk;           /* [B, Skv, >=H*64] bf16 */
  void* dv;           /* [B, Skv, >=H*64] bf16 */
  int32_t B, H, Sq, Skv, head_dim;
  float scale;
  int64_t q_row_stride, k_row_stride, v_row_stride, out_row_stride, dout_row_stride, dq_row_stride, dk_row_stride, dv_row_stride;
  int64_t q_batch_stride, k_batch_stride, v_batch_stride, out_batch_stride, dout_batch_stride, dq_batch_stride,
      dk_batch_stride, dv_batch_stride;
  void* workspace;
  size_t workspace_bytes;
} vgpa_attention_bwd_args;
size_t vgpa_attention_bwd_workspace_bytes(int B, int H, int Sq);
int vgpa_attention_bwd_bf16(const vgpa_attention_bwd_args* args, void* stream);


/* ------------------------------------------------------------------------------------------------
 * K3 — fused LayerNorm + adaLN modulation: out = LN(x) * (1 + scale[b, seg]) + shift[b, seg].
 * Replaces CogVideoXLayerNormZero / AdaLayerNorm / norm_final of CogVideoXTransformer3DModel
 * (SURVEY.md App. A.1). seg = text for the first text_rows rows of each sample, video otherwise.
 * All modulation pointers NULL = plain LayerNorm. ln_weight/ln_bias NULL = no elementwise affine.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vgpa_layernorm_args {
  const void* x;  /* [rows, ldx] bf16 */
  void* out;      /* [rows, ldo] bf16 */
  int32_t rows, D;
  int64_t ldx, ldo;
  const void* ln_weight; /* [D] bf16 or NULL */
  const void* ln_bias;   /* [D] bf16 or NULL */
  float eps;
  int32_t rows_per_sample, text_rows;
  const void* shift_txt; /* [B][D] bf16, batch stride mod_stride_b elements */
  const void* scale_txt;
  const void* shift_vid;
  const void* scale_vid;
  int64_t mod_stride_b;
  /* 1: x is float [rows, ldx] (fp32 residual stream of the Wan2.2 forward); the normalisation, affine and modulation are
   * then evaluated in fp32 and rounded to bf16 once (torch.autocast semantics) instead of torch's eager-bf16 rounding
   * after every op. 0: x is bf16. */
  int32_t x_is_f32;
} vgpa_layernorm_args;

int vgpa_layernorm_modulate_bf16(const vgpa_layernorm_args* args, void* stream);

/* Row-wise backward kernels of CogVideoXBlock for the DPO training step (train/CogVideoX-5B/03_train.py:116-157 back-
 * propagates into the LoRA factors of to_q / to_k / to_v / to_out.0). */
/* dx = [add +] d/dx of vgpa_layernorm_modulate_bf16 applied to dy: `fwd` describes the forward call (x, ln_weight, eps,
 * scale_txt / scale_vid, rows_per_sample, text_rows, mod_stride_b; out / shift / ln_bias are ignored). The modulation
 * vectors come from the frozen conditioning path and get no gradient. */
int vgpa_layernorm_modulate_bwd_bf16(const vgpa_layernorm_args* fwd, const void* dy, int64_t ld_dy, const void* add,
                                     int64_t ld_add, void* dx, int64_t ld_dx, void* stream);
/* LayerNorm over every 64-column head of x [rows, >=heads*64] with fp32 affine [64] (norm_q / norm_k of
 * CogVideoXAttnProcessor2_0). backward = 0: out = LN(x) * weight + bias. backward = 1: out = d/dx for the upstream
 * gradient dy (weight / bias are frozen). */
int vgpa_head_layernorm_bf16(const void* x, const void* dy, void* out, int64_t rows, int heads, int64_t ldx, int64_t ld_dy,
                             int64_t ldo, const float* weight, const float* bias, float eps, int backward, void* stream);
/* GELU(tanh) on n contiguous bf16 values. backward = 0: out = gelu(x). backward = 1: out = dy * gelu'(x). */
int vgpa_gelu_tanh_bf16(const void* x, const void* dy, void* out, int64_t n, int backward, void* stream);
/* out[c, r] = x[r, c] for r < rows and 0 for rows <= r < ldo: the [tokens, features] -> [features, tokens (padded to the
 * GEMM's K granularity)] operand of the LoRA weight-gradient GEMMs dB = dy^T u, dA = du^T a. ldx, ldo even. */
int vgpa_transpose_bf16(const void* x, void* out, int rows, int cols, int64_t ldx, int64_t ldo, void* stream);
/* out[r, :] = [add[r, :] +] x[r, :] * gate[b(r), seg(r)][:]: the adaLN-zero gated residual `hidden + gate * branch` (bf16
 * product, then bf16 sum) out of place for the training path, and its backward (add = NULL). */
int vgpa_scale_cols_bf16(const void* x, const void* add, void* out, int rows, int D, int64_t ldx, int64_t ld_add, int64_t ldo,
                         int rows_per_sample, int text_rows, const void* gate_txt, const void* gate_vid, int64_t gate_stride_b,
                         void* stream);

/* ------------------------------------------------------------------------------------------------
 * Conditioning path and patch (un)embedding helpers.
 * ---------------------------------------------------------------------------------------------- */
enum { VGPA_ACT_NONE = 0, VGPA_ACT_SILU = 1 };
/* out[M, N] = bias + act_in(x[M, K]) . W[N, K]^T, M <= 8 (time_embedding.linear_{1,2}, adaLN linears) */
int vgpa_linear_smallm_bf16(const void* x, const void* W, const void* bias, void* out, int M, int N, int K,
                            int64_t ldx, int64_t ldo, int act_in, void* stream);
/* diffusers get_timestep_embedding(flip_sin_to_cos=True, downscale_freq_shift=0): out[B, dim] bf16 */
int vgpa_timestep_embedding_bf16(const float* d_timesteps, void* out, int B, int dim, void* stream);
/* [BF, C, H, W] -> [BF*(H/2)*(W/2), C*4] (feature = c*4 + ph*2 + pw) and its inverse (input row stride ldi) */
int vgpa_patchify_bf16(const void* in, void* out, int BF, int C, int H, int W, void* stream);
int vgpa_unpatchify_bf16(const void* in, void* out, int BF, int C, int H, int W, int64_t ldi, void* stream);

/* ------------------------------------------------------------------------------------------------
 * CFG combine + v-prediction scheduler update in one pass (CogVideoXPipeline
 * denoise loop + CogVideoXDDIMScheduler.step / CogVideoXDPMScheduler.step).
 *   v        = pred_uncond + guidance * (pred_cond - pred_uncond)      (pred_uncond NULL: v = pred_cond)
 *   x0       = bf16(sqrt_alpha_t * x) - sqrt_beta_t * v
 *   DDIM:  prev = bf16(c_sample * x) + c_x0 * x0
 *   DPM :  prev = bf16(c_sample * x) + c_x0 * x0 + c_x0_old * x0_old + c_noise * noise
 * x0_out (fp32, optional) receives x0 for the multistep solver.
 * ---------------------------------------------------------------------------------------------- */
enum { VGPA_SCHED_DDIM = 0, VGPA_SCHED_DPM = 1 };
typedef struct vgpa_sched_args {
  const void* pred_uncond; /* [n] bf16 or NULL */
  const void* pred_cond;   /* [n] bf16 */
  const void* sample;      /* [n] bf16 */
  void* prev_sample;       /* [n] bf16 */
  const float* x0_old;     /* [n] fp32 or NULL */
  float* x0_out;           /* [n] fp32 or NULL */
  const void* noise;       /* [n] bf16 or NULL */
  int64_t n;
  int32_t mode;
  float guidance, sqrt_alpha_t, sqrt_beta_t, c_sample, c_x0, c_x0_old, c_noise;
} vgpa_sched_args;

int vgpa_cfg_scheduler_step(const vgpa_sched_args* args, void* stream);

/* ================================================================================================
 * Geometry-consistency scorer (SURVEY.md §8 rows a-10 ... a-15). All pointers are DEVICE pointers;
 * every entry point is asynchronous on `stream`. Workspaces are caller-owned, 256-byte aligned.
 * ============================================================================================== */

/* K5 — MVCS for a batch of clips: replaces MVCSMetric.compute (metrics/mvcs.py:12-114).
 *   depths [n_clips, T, H, W] fp32; intrinsics [n_clips, T, k_dim, k_dim] (k_dim 3 or 4, top-left 3x3 used);
 *   extrinsics [n_clips, T, e_rows, 4] world->camera (e_rows 3 or 4).
 *   pair_mse / pair_cnt [n_clips, T-1] (optional): masked MSE and mask size of pair (i, i+1);
 *   scores [n_clips] fp64 = exp(-mean of the non-empty pairs' MSE), 0.0 when no pair contributes. */
size_t vgpa_mvcs_workspace_bytes(int n_clips, int T, int H, int W);
int vgpa_mvcs_blocks_per_pair(int n_clips, int T, int H, int W);
int vgpa_mvcs_batch(const float* d_depths, const float* d_intrinsics, const float* d_extrinsics, int n_clips, int T,
                    int H, int W, int k_dim, int e_rows, void* d_workspace, size_t workspace_bytes,
                    double* d_pair_mse, int64_t* d_pair_cnt, double* d_scores, void* stream);

/* K6 — reprojection renderer: replaces batch_reproject / project_points (utils/projection_utils.py:12-101).
 *   points, colors [N, 3] fp32; intrinsics [T, 3, 3]; extrinsics [T, e_rows, 4]; out [T, 3, H, W] fp32 in [-1, 1].
 *   Nearest z wins; exact z ties go to the lowest point index. */
size_t vgpa_reproject_workspace_bytes(int T, int H, int W);
int vgpa_reproject_batch(const float* d_points, const float* d_colors, const float* d_intrinsics,
                         const float* d_extrinsics, int64_t n_points, int T, int H, int W, int e_rows,
                         void* d_workspace, size_t workspace_bytes, float* d_out, void* stream);

/* K7 — coloured point cloud: replaces get_colored_pointcloud (utils/pointcloud_utils.py:10-80).
 *   points [N, 3], conf [N], images [T, 3, H, W] (images_nhwc = 0) or [T, H, W, 3] (= 1) in [0, 1], N = T*H*W.
 *   Keeps finite conf > 1e-5 and, for conf_thres > 0, conf >= the k-th largest valid value with
 *   k = max(1, ceil(N_valid * (1 - conf_thres/100))). Survivors keep their order. out_count receives N'. */
size_t vgpa_pointcloud_workspace_bytes(int64_t n_points);
int vgpa_pointcloud_filter(const float* d_points, const float* d_images, const float* d_conf, int64_t n_points,
                           int hw_per_frame, int images_nhwc, double conf_thres, void* d_workspace,
                           size_t workspace_bytes, float* d_out_vertices, float* d_out_colors, int64_t* d_out_count,
                           float* d_out_threshold, void* stream);

/* K8 — fundamental matrix + Sampson distance per frame pair: replaces kornia find_fundamental +
 * sampson_epipolar_distance as used by metrics/epipolar.py:194-216.
 *   pts1, pts2 [n_pairs, max_matches, 2] fp32; counts [n_pairs] (NULL = all max_matches);
 *   F [n_pairs, 3, 3]; mean_dist [n_pairs] = mean sqrt(d^2 + 1e-8); valid [n_pairs] (0 when < 8 matches or NaN). */
int vgpa_epipolar_batch(const float* d_pts1, const float* d_pts2, const int32_t* d_counts, int n_pairs, int max_matches,
                        float* d_F, float* d_mean_dist, int32_t* d_valid, void* stream);

/* Consistency-score geometry: compute_motion_score_vectorized (metrics/consistency_score.py:8-38),
 * MSEMetric.compute with its range heuristics (metrics/mse.py:14-54), and the DA3 depth un-projection
 * (pipelines/process_video.py:151-156). kind: 0 = fp32, 1 = uint8; nhwc: source layout; numpy: the
 * reference's ndarray branch (only `max > 1 -> /255`). */
int vgpa_motion_score(const float* d_extrinsics, int T, int e_rows, float* d_out, void* stream);
size_t vgpa_mse_workspace_bytes(void);
int vgpa_mse_range_normalized(const void* d_gt, int gt_kind, int gt_nhwc, int gt_numpy, const void* d_rep, int rep_kind,
                              int rep_nhwc, int rep_numpy, int64_t N, int C, int H, int W, void* d_workspace,
                              size_t workspace_bytes, float* d_out, void* stream);
int vgpa_unproject_depth(const float* d_depth, const float* d_intrinsics, const float* d_extrinsics_w2c, int T, int H,
                         int W, int e_rows, float* d_out_points, void* stream);

/* K9 — DPO loss: replaces DPOLoss.forward (train/loss.py:53-121) and its autograd.
 *   tensors[6] = v_win, v_lose, v_win_ref, v_lose_ref, v_win_target, v_lose_target, each [B, n_per_sample],
 *   fp32 (is_bf16 = 0) or bf16 (= 1). out5 = loss, reward_margin, winner_reward, loser_reward, accuracy.
 *   err4 [4, B] (optional) = model_win, model_lose, ref_win, ref_lose MSEs. coef [2, B] (needed for
 *   backward) = dLoss/d(model_win_err), dLoss/d(model_lose_err). */
enum { VGPA_DPO_SIGMOID = 0, VGPA_DPO_HINGE = 1, VGPA_DPO_SFT = 2 /* loss = mean MSE(tensors[0], tensors[4]) */ };
typedef struct vgpa_dpo_args {
  const void* tensors[6];
  int32_t is_bf16[6];
  int32_t B;
  int64_t n_per_sample;
  float beta, label_smoothing;
  int32_t loss_type;
  float* d_out5;
  float* d_err4;
  float* d_coef;
  void* d_workspace;
  size_t workspace_bytes;
} vgpa_dpo_args;
size_t vgpa_dpo_workspace_bytes(int B, int64_t n_per_sample);
int vgpa_dpo_loss_forward(const vgpa_dpo_args* args, void* stream);
int vgpa_dpo_loss_backward(const vgpa_dpo_args* args, const float* d_grad_loss, void* d_grad_win, void* d_grad_lose,
                           void* stream);


/* ------------------------------------------------------------------------------------------------
 * Wan2.2 DiT helpers (generate/Wan2.2-TI2V-5B.py:120-129). The block itself runs on
 * vgpa_linear_bf16 / vgpa_attention_bf16 (head_dim 128) / vgpa_layernorm_modulate_bf16.
 * ---------------------------------------------------------------------------------------------- */
/* In place on x [rows, ldx]: WanRMSNorm over D columns (`_norm(x.float()).type_as(x) * weight`, weight fp32 [D]) and,
 * when rope tables are given, the complex-pair rotation of rope_apply on every head (tables [rows_per_sample, head_dim]
 * fp32, repeat-interleaved cos / sin; rows_per_sample = 0 means rows). */
int vgpa_rmsnorm_rope_bf16(void* x, int rows, int D, int64_t ldx, const float* weight, float eps, const float* rope_cos,
                           const float* rope_sin, int head_dim, int rows_per_sample, void* stream);
/* out[r, :] = bf16(a[r, :] + b[:]) for R rows of N columns (`modulation + e` of WanAttentionBlock / Head); a bf16 with
 * row stride lda, b fp32. */
int vgpa_add_rows_bf16(const void* a, const float* b, void* out, int R, int64_t N, int64_t lda, void* stream);

/* ------------------------------------------------------------------------------------------------
 * T5 prompt encoder helpers. Replace transformers' T5Attention / T5DenseGatedActDense
 * arithmetic behind `pipeline.text_encoder(input_ids)[0]` (train/CogVideoX-5B/02_encode.py:69-84, 226 tokens, no
 * attention mask); the projections run on vgpa_linear_bf16 and T5LayerNorm on vgpa_rmsnorm_rope_bf16.
 * ---------------------------------------------------------------------------------------------- */
/* Self-attention with d_kv = 64, NO 1/sqrt(d) scaling and an additive position bias [H, S, S] (bf16):
 * scores = bf16(q k^T); scores = bf16(scores + bias); P = bf16(softmax_fp32(scores)); out = bf16(P v).
 * q, k, v: rows [B*S, ld_qkv] with head h at columns [64h, 64h+64) (pointers into a fused projection are fine);
 * out [B*S, ldo]. 0 < S <= 512. */
int vgpa_t5_attention_bf16(const void* q, const void* k, const void* v, const void* bias, void* out, int B, int H, int S,
                           int64_t ld_qkv, int64_t ldo, void* stream);
/* out = bf16(a * b) elementwise on [rows, N] (T5DenseGatedActDense: hidden_gelu * hidden_linear); N % 8 == 0. */
int vgpa_gated_mul_bf16(const void* a, const void* b, void* out, int rows, int N, int64_t lda, int64_t ldb, int64_t ldo,
                        void* stream);

/* ================================================================================================
 * K4 — CogVideoX VAE decoder building blocks. Replace the cuDNN conv3d / GroupNorm / interpolate calls
 * behind diffusers' AutoencoderKLCogVideoX.decode (CogVideoXDecoder3D, CogVideoXResnetBlock3D,
 * CogVideoXSpatialNorm3D, CogVideoXUpsample3D, CogVideoXCausalConv3d; reference call
 * sites generate/CogVideoX-5B.py:20-21,72-77). Activations are channels-last [T, H, W, C] bf16.
 * ============================================================================================== */

/* Causal 3-D convolution as an implicit GEMM on tcgen05 (K = KT*3*3*Cin):
 *   out[t,h,w,co] = bias[co] + residual[t,h,w,co] + sum_{kt,kh,kw,ci} x[t+kt, h+kh-1, w+kw-1, ci] * w[co][(kt,kh,kw,ci)]
 * x is already padded in time (KT-1 leading frames: the conv_cache or the replicated first frame, written by
 * the caller); H/W are zero padded by the TMA unit. KT in {1, 3} (KT = 1 is the per-frame Conv2d 3x3 of
 * CogVideoXUpsample3D); the kernel is 3x3 in space. Cin % 64 == 0. Cout_pad (rows of w) in {16, 64, 128} or a
 * multiple of 256; only the first Cout channels are written. */
typedef struct vgpa_conv3d_args {
  const void* x;        /* [T + KT - 1, H, W, Cin] bf16                                              */
  const void* w;        /* [Cout_pad, KT*9*Cin] bf16, K ordered (kt, kh, kw, ci)                      */
  const void* bias;     /* [Cout_pad] bf16 or NULL                                                   */
  const void* residual; /* [T, H, W, ld_res] bf16 or NULL                                            */
  void* out;            /* [T, H, W, ldo] bf16                                                       */
  int32_t T, H, W, Cin, Cout, Cout_pad, KT;
  int32_t ldo, ld_res;
  /* Optional (ABI 4): GroupNorm statistics of the tensor being written, produced by the conv epilogue instead of a separate
   * pass over it (the next CogVideoXSpatialNorm3D / GroupNorm normalises exactly this tensor). gn_mean_rstd [2, gn_groups]
   * fp32 or NULL; gn_workspace >= vgpa_conv3d_gn_workspace_bytes(Cout), 16-byte aligned; Cout <= 512. */
  float* gn_mean_rstd;
  void* gn_workspace;
  int32_t gn_groups;
  float gn_eps;
} vgpa_conv3d_args;
size_t vgpa_conv3d_gn_workspace_bytes(int Cout);
int vgpa_conv3d_causal_bf16(const vgpa_conv3d_args* args, void* stream);

/* GroupNorm statistics over one [T, H, W, C] tensor (all of T, H, W: the statistics of one frame batch of one
 * tile, as torch.nn.GroupNorm sees it): mean_rstd [2, groups] fp32. workspace >= vgpa_groupnorm_workspace_bytes. */
size_t vgpa_groupnorm_workspace_bytes(int C);
int vgpa_groupnorm_stats_bf16(const void* x, int64_t n_pixels, int C, int groups, float eps, void* workspace,
                              size_t workspace_bytes, float* mean_rstd, void* stream);

/* SpatialNorm3D apply (+ optional SiLU): out = act((x - mean_g) * rstd_g * gamma_c + beta_c) * Y[z(t,h,w), c] + Bz[z(t,h,w), c])
 * where Y = conv_y(zq), Bz = conv_b(zq) are given at LATENT resolution [Tz, Hz, Wz, C] (pointwise convs commute
 * with the nearest-neighbour resize) and z(t,h,w) = (tz_of_t[t], h >> shift, w >> shift). */
typedef struct vgpa_spatialnorm_args {
  const void* x;          /* [T, H, W, C] bf16 */
  void* out;              /* [T, H, W, C] bf16 (may point 2 frames into a time-padded buffer) */
  const float* mean_rstd; /* [2, groups] */
  const void* gamma;      /* [C] bf16 */
  const void* beta;       /* [C] bf16 */
  const void* y_lat;      /* [Tz, Hz, Wz, >=C] bf16, pixel stride ld_lat elements */
  const void* b_lat;      /* same layout */
  int64_t ld_lat;
  int32_t T, H, W, C, groups;
  int32_t Hz, Wz, shift;
  int32_t tz_of_t[16];    /* T <= 16 */
  int32_t silu;
} vgpa_spatialnorm_args;
int vgpa_spatialnorm_apply_bf16(const vgpa_spatialnorm_args* args, void* stream);

/* Nearest-neighbour upsample x2 in H and W; output frame t reads input frame t_src[t] (T_out <= 16). */
int vgpa_upsample_nearest_bf16(const void* x, void* out, int T_out, int H, int W, int C, const int32_t* t_src_host,
                               void* stream);

/* Tiled-decode composition (AutoencoderKLCogVideoX.tiled_decode blending + crop): tiles[i*cols + j] is the decoded
 * tile [T, th_i, tw_j, ldc] bf16 (channels-last, first 3 channels used); out [3, T, H, W] bf16. */
typedef struct vgpa_compose_args {
  const void* tiles[16];
  int32_t rows, cols;
  int32_t th[4], tw[4];           /* decoded tile heights per tile row / widths per tile column */
  int32_t T, H, W, ldc;
  int32_t blend_h, blend_w;       /* blend extents (rows / columns) */
  int32_t limit_h, limit_w;       /* crop of every tile */
  void* out;
} vgpa_compose_args;
int vgpa_vae_compose_tiles_bf16(const vgpa_compose_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOGPA_B200_H */
