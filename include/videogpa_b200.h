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

#define VGPA_ABI_VERSION 1

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
  VGPA_EPI_QKV = 3        /* fused to_q|to_k|to_v: per-head LayerNorm(64) on q,k + 3-D RoPE on     */
                          /* video rows; v passes through      (CogVideoXAttnProcessor2_0)         */
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
} vgpa_linear_args;

int vgpa_linear_bf16(const vgpa_linear_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIDEOGPA_B200_H */
