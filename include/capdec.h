/* capdec.h -- C-ABI of libcapdec.so: the B200-native caption-decoder hot path.
 *
 * Drop-in boundary for the reference's `Decoder.call` + loss wiring
 * (ImageCaptioning/model/model_adaAttention_aic.py:40-135, 161-183, 205-212).  The reference has no FFI /
 * plugin interface of its own (it is a PaddlePaddle graph-builder call), so these entry points are what a
 * ctypes binding placed at model:178 (train) and model:188 (eval) would call; INTEGRATION.md shows that stub.
 *
 * Conventions
 *   - every function returns 0 on success, < 0 on error; capdec_last_error() gives the message (thread-local);
 *     no C++ exception crosses the boundary;
 *   - all tensor arguments are DEVICE pointers owned by the caller (e.g. torch tensors), plain row-major,
 *     fp32 / int64 / int32 as stated; the library never allocates caller-visible memory;
 *   - library-owned workspace lives in the handle and grows on demand (capdec_workspace_bytes reports it);
 *   - one handle per (process, GPU); a handle is not thread-safe; every call is asynchronous on `stream`
 *     (a cudaStream_t passed as void*), results are valid after that stream is synchronised.
 */
#ifndef CAPDEC_H_
#define CAPDEC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPDEC_DTYPE_F32 0  /* fp32 storage, SIMT GEMMs: validation precision (rel 1e-3 vs the oracle)        */
#define CAPDEC_DTYPE_BF16 1 /* bf16 storage, tcgen05 GEMMs with fp32 accumulation: the fast path (rel 2e-2)   */

#define CAPDEC_SOFTMAX_FEATURE 0  /* model:107 as written: softmax over the size-1 axis, alpha == 1 */
#define CAPDEC_SOFTMAX_POSITION 1 /* soft-max over the k+1 attention positions (Lu et al. 2017)                 */

/* Validation schedule: run the recurrence of the teacher-forced step as one GEMM + one cell kernel per time step (the
 * schedule every hidden size supports) instead of the two persistent kernels (hidden_dim % 128 == 0, <= 1024).  The
 * parity tests compare the two; there is no environment switch anywhere in the library. */
#define CAPDEC_FLAG_STEPWISE_RECURRENCE 1

/* Bits of the asynchronous input-error report (capdec_status) */
#define CAPDEC_ERR_SOURCE_ID 1 /* caption[:, :-1] holds an id outside [0, vocab_size)  (lookup_table_v2 enforces this) */
#define CAPDEC_ERR_TARGET_ID 2 /* caption[:, 1:] holds an id outside [0, vocab_size)   (softmax_with_cross_entropy)    */
#define CAPDEC_ERR_LENGTHS 4   /* lengths[] does not describe the captions (see capdec_train_fwd)                      */

/* config.py:22-24 (ids) and config.py:51-60 (dims) */
typedef struct capdec_config {
  int32_t hidden_dim;       /* config.py:59  (1024) multiple of 8          */
  int32_t embedding_size;   /* config.py:56  (256)  multiple of 8          */
  int32_t vocab_size;       /* config.py:55  (12295)                       */
  int32_t num_positions;    /* config.py:51  (49) <= 63                    */
  int32_t sentence_length;  /* config.py:57  (35): train runs sentence_length-1 steps (model:66) */
  int32_t infer_max_length; /* config.py:60  (35)                          */
  int32_t padding_idx;      /* config.py:24  (0)                           */
  int32_t start_idx;        /* config.py:22  (2)                           */
  int32_t stop_idx;         /* config.py:23  (3)                           */
  int32_t dtype;            /* CAPDEC_DTYPE_*                              */
  int32_t softmax_axis;     /* CAPDEC_SOFTMAX_*                            */
  int32_t flags;            /* CAPDEC_FLAG_*, normally 0                   */
} capdec_config;

/* The decoder's parameters, fp32, PaddlePaddle layouts (fc weight = [in, out]); names follow the construction
 * order of model:52, 53, 87-88, 89, 90, 99, 102, 104, 107, 115, 24, 18, 23. */
typedef struct capdec_params {
  const float* fc_v_w;  const float* fc_v_b;   /* model:52  [H,H],[H]      values = tanh(fc(P))  */
  const float* fc_k_w;  const float* fc_k_b;   /* model:53  [H,H],[H]      keys   = fc(P)        */
  const float* lstm_w;  const float* lstm_b;   /* model:87-88 [E+2H,4H],[4H]  rows: word_emb | global | hid ; gates i,f,o,g */
  const float* fc_xs_w; const float* fc_xs_b;  /* model:89  [E+H,H],[H]    p_word_emb            */
  const float* fc_hs_w; const float* fc_hs_b;  /* model:90  [H,H],[H]      p_hidden              */
  const float* fc_p_w;  const float* fc_p_b;   /* model:99  [H,H],[H]      p_hid (tanh)          */
  const float* fc_g_w;  const float* fc_g_b;   /* model:102 [H,H],[H]      hid_emb               */
  const float* fc_s_w;  const float* fc_s_b;   /* model:104 [H,H],[H]      sentinel_emb          */
  const float* fc_a_w;  const float* fc_a_b;   /* model:107 [H,1],[1]      alpha score           */
  const float* fc_o_w;  const float* fc_o_b;   /* model:115 [H,H],[H]      out (tanh)            */
  const float* fc_e_w;  const float* fc_e_b;   /* model:24  [H,E],[E]      proj_x                */
  const float* word_embedding;                  /* model:16-19, 29-32 [V,E] tied lookup / output  */
  const float* out_fc_bias;                     /* model:20-23 [V]                                */
} capdec_params;

/* Gradients, same names / shapes / layouts as capdec_params (device fp32, written by capdec_train_bwd). */
typedef struct capdec_grads {
  float* fc_v_w;  float* fc_v_b;
  float* fc_k_w;  float* fc_k_b;
  float* lstm_w;  float* lstm_b;
  float* fc_xs_w; float* fc_xs_b;
  float* fc_hs_w; float* fc_hs_b;
  float* fc_p_w;  float* fc_p_b;
  float* fc_g_w;  float* fc_g_b;
  float* fc_s_w;  float* fc_s_b;
  float* fc_a_w;  float* fc_a_b;
  float* fc_o_w;  float* fc_o_b;
  float* fc_e_w;  float* fc_e_b;
  float* word_embedding;
  float* out_fc_bias;
} capdec_grads;

/* The feature adaptor between the encoder and the decoder: ImageCaptionModel._img2feature from the encoder output on
 * (model:193-199).  fp32, PaddlePaddle [in, out] layouts, X = encoder_channel
 * (config.py:52, 1280). */
typedef struct capdec_adaptor_params {
  const float* fc_img_w;  const float* fc_img_b;   /* model:197 [X,H],[H]  img_embed  = relu(fc(map^T))          */
  const float* fc_glob_w; const float* fc_glob_b;  /* model:199 [X,H],[H]  img_global = relu(fc(mean_k(map^T)))  */
} capdec_adaptor_params;
typedef struct capdec_adaptor_grads {
  float* fc_img_w;  float* fc_img_b;
  float* fc_glob_w; float* fc_glob_b;
} capdec_adaptor_grads;

typedef struct capdec_handle_s* capdec_handle;

const char* capdec_last_error(void);
int capdec_version(void);

/* Decoder.__init__ (model:42-48).  Rejects unsupported configurations with an error (cf. ValueError model:43-44). */
int capdec_create(const capdec_config* cfg, int device, capdec_handle* out);
int capdec_destroy(capdec_handle h);
/* bytes of device workspace currently owned by the handle */
size_t capdec_workspace_bytes(capdec_handle h);
/* number of kernels launched by this handle since creation (bench accounting) */
int64_t capdec_launch_count(capdec_handle h);

/* Device-side phase timeline (profiling aid): while enabled, the schedules record CUDA events on the caller's stream
 * at phase boundaries; capdec_profile_report synchronises on the last one and writes "phase milliseconds\n" lines
 * (summed per phase since the previous report) into buf. */
int capdec_profile(capdec_handle h, int32_t enable);
int capdec_profile_report(capdec_handle h, char* buf, size_t n);

/* Converts the fp32 parameters into the handle's internal operand layouts.  Call again after every optimizer
 * update.  The bias / w_alpha pointers are read directly by later calls and must stay valid. */
int capdec_bind_params(capdec_handle h, const capdec_params* p, void* stream);

/* Teacher-forced forward: Decoder(mode='train').call + training_network loss wiring (model:50-130, 161-183).
 *   global_img_feat [B,H] f32, p_img_feat [B,k,H] f32, caption [B,sentence_length] int64
 *   (both feature pointers NULL: use the features staged in the handle by capdec_img2feature for the same B)
 *   lengths         HOST array [B] (the only host-side operand of the library) or NULL: the number of leading
 *                   non-padding tokens of caption[b] = [start, w_1 .. w_L, stop, pad ..] (ai_challenge_tokenizer.py:81-86),
 *                   i.e. L + 2.  The reference has no such input (every sample runs all 34 steps, the mask
 *                   is target != pad, model:163-171); rows past the last target contribute exactly nothing to the loss or
 *                   to any gradient, so with `lengths` (and logits_out == NULL) everything downstream of the recurrence runs
 *                   on the sum_b (lengths[b] - 1) live rows only.  Results equal the call without `lengths`.  The array is
 *                   read before the call returns and only sizes the work; the device takes the live steps from the captions
 *                   themselves and reports a total that differs (CAPDEC_ERR_LENGTHS, capdec_status).
 *   logits_out      [B,T,V] f32 batch-major (model:129-130) or NULL when only the loss is wanted
 *   loss_out        4 floats on the device: {sum(ce*mask), sum(mask), loss = ratio, 1/sum(mask)}            */
int capdec_train_fwd(capdec_handle h, const float* global_img_feat, const float* p_img_feat, const int64_t* caption,
                     const int32_t* lengths, int32_t B, float* logits_out, float* loss_out, void* stream);

/* Backward of the last capdec_train_fwd (Paddle's appended *_grad / while_grad ops).
 *   d(loss)/d(logits) = (softmax - onehot) * mask * inv_ntok * loss_scale, where inv_ntok is read from
 *   `inv_ntok_dev` (device float; pass NULL to use 1/sum(mask) of this handle's batch -- data-parallel callers
 *   pass the all-reduced global value).  All grads are overwritten (not accumulated).
 *   d_global_img_feat [B,H] f32, d_p_img_feat [B,k,H] f32 (either may be NULL). */
int capdec_train_bwd(capdec_handle h, float loss_scale, const float* inv_ntok_dev, float* d_global_img_feat,
                     float* d_p_img_feat, const capdec_grads* g, void* stream);

/* Data-parallel callers (the reference's ParallelExecutor wrapper, train.py:119-127): up to six
 * cudaEvent_t (as void*; n = 0 switches the marks off) that every following capdec_train_bwd records on its stream as
 * soon as a group of gradients is final, so that the caller can all-reduce that group while the rest of the backward pass
 * still runs.  Groups (rows refer to lstm_w [E+2H, 4H]), in the order they become final: 0, 2, 3, 4, 1, 5.  Only the
 * weight gradients of the head are final before the BPTT kernel; the gate-weight gradients, which exist only after it,
 * come next (largest first) and the vocabulary-side and per-image weight gradients -- independent of the recurrence -- are
 * computed last, so that every exchange but the last small one has work to hide behind:
 *   events[0]  fc_hs_w .. fc_e_b (fc_hs, fc_p, fc_g, fc_s, fc_a, fc_o, fc_e)                      -- before the BPTT kernel
 *   events[2]  lstm_w rows [E+H, E+2H)  (hidden-state rows)                                        -- first work after it
 *   events[3]  lstm_w rows [E, E+H)     (global-image rows)
 *   events[4]  lstm_b, fc_xs_w, fc_xs_b, lstm_w rows [0, E) (word rows)
 *   events[1]  word_embedding, out_fc_bias
 *   events[5]  fc_v_w .. fc_k_b                                                                    -- last work of the call */
int capdec_set_grad_ready_events(capdec_handle h, void* const* events, int32_t n);

/* Inference: Decoder(mode='eval').call (model:55-58, 66-69, 119-123, 132-133) generalised to beam search
 * (SURVEY.md 8c row C5).  beam = 1, finish_on_stop = 0 is the reference's greedy decode, bit for bit.
 *   ids_out [B,beam,T] int32 (beams sorted by score), scores_out [B,beam] f32 (either may be NULL),
 *   ids_f32_out [B,T] f32 = best beam as float32 ids, the reference's return type (model:122, 132-133), or NULL */
int capdec_decode(capdec_handle h, const float* global_img_feat, const float* p_img_feat, int32_t B, int32_t beam,
                  int32_t finish_on_stop, int32_t* ids_out, float* scores_out, float* ids_f32_out, void* stream);

/* Input checks that run on the device (token ids in range, lengths consistent with the captions) cannot fail the
 * asynchronous call that launched them.  Offending entries are treated as padding (no out-of-bounds access ever happens)
 * and the NEXT capdec_train_fwd / capdec_train_bwd on the handle fails with the report; capdec_status synchronises
 * `stream` and returns it right away (< 0 and capdec_last_error(), the report is cleared by being returned). */
int capdec_status(capdec_handle h, void* stream);

/* The selections of the last capdec_decode on this handle, step by step: trace_out [T, B, beam, 2] int32 =
 * (token, back-pointer into the previous step's beams) -- the beam lattice; ids_out is its back-traced form.  Parity
 * tests compare it with the oracle's selections step by step (the near-tie protocol). */
int capdec_decode_trace(capdec_handle h, int32_t* trace_out, void* stream);

/* ---- the callers either side of the decoder ---------------------------------------------------- */

/* Converts the adaptor's fp32 parameters into the handle's operand layout (call again after every optimizer update).
 * The bias pointers are read directly by later calls and must stay valid. */
int capdec_bind_adaptor(capdec_handle h, int32_t encoder_channel, const capdec_adaptor_params* p, void* stream);

/* ImageCaptionModel._img2feature after the encoder (model:195-199):
 *   enc_map [B, X, k] f32 = the encoder output [B,1280,7,7] reshaped (model:195; cast model:194 is a no-op for f32)
 *   -> p_img_feat [B,k,H] = relu(fc(transpose(enc_map)))  and  global_img_feat [B,H] = relu(fc(mean over k)).
 * The results are STAGED in the handle in its storage type: a following capdec_train_fwd / capdec_decode called with
 * NULL feature pointers and the same B consumes them without a round trip through fp32.  p_img_feat_out /
 * global_img_feat_out (f32, either may be NULL) receive copies for callers that want to look at them. */
int capdec_img2feature(capdec_handle h, const float* enc_map, int32_t B, float* p_img_feat_out,
                       float* global_img_feat_out, void* stream);

/* Backward of the staged adaptor, after capdec_img2feature -> capdec_train_fwd(NULL, NULL, ..) -> capdec_train_bwd:
 * the four parameter gradients are overwritten; d_enc_map [B, X, k] f32 (or NULL) is the gradient handed to the encoder
 * (config.py:50 encoder_trainable). */
int capdec_img2feature_bwd(capdec_handle h, float* d_enc_map, const capdec_adaptor_grads* g, void* stream);

/* One step of fluid.optimizer.Adam behind opt.minimize(loss) (train.py:26-31, 44-45) over a flat fp32 buffer of n
 * elements, PaddlePaddle 1.8 update rule:
 *   lr_t = lr * sqrt(1 - beta2_pow) / (1 - beta1_pow)       beta?_pow = beta?^t of THIS step (t = 1, 2, ..)
 *   m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2;  param -= lr_t * m / (sqrt(v) + epsilon)
 * clip_value > 0 applies GradientClipByValue(clip_value) to g first (train.py:41-42); 0 = off (config.py:36). */
int capdec_adam_step(float* param, const float* grad, float* moment1, float* moment2, int64_t n, float lr, float beta1,
                     float beta2, float epsilon, float beta1_pow, float beta2_pow, float clip_value, void* stream);

/* Data-parallel gradient exchange in bf16 (train.py:119-127 is the wrapper being replaced): a range of the
 * flat fp32 gradient buffer is rounded to bf16 (round to nearest even) for the NCCL all-reduce and widened back, so
 * that half the bytes cross NVLink.  n % 8 == 0, 16-byte aligned pointers.  to_bf16 != 0: src fp32 -> dst bf16;
 * to_bf16 == 0: src bf16 -> dst fp32. */
int capdec_cast_bf16(int32_t to_bf16, const void* src, void* dst, int64_t n, void* stream);

/* ---- kernel-level entry points (parity tests and micro-benchmarks) ------------------------------------------ */

/* C[M,N] = act(A.B^T + bias + addend).  dtype: element type of A, B, addend, dact_y and (unless out_f32) C.
 *   majors bit 0: A is stored [K,M] (MN-major) instead of [M,K];  bit 1: B is stored [K,N] instead of [N,K]
 *   (0 = data-grad form, 2 = forward fc with a PaddlePaddle [in,out] weight as B, 3 = weight-grad form)
 *   act: 0 none, 1 tanh, 2 multiply by (1 - y^2) with y = dact_y, 5 relu, 6 multiply by [y > 0];  atomic: fp32 atomicAdd into C (split-K)
 *   variant (bf16 only): 0 = the library's own choice, 1 = SIMT kernel (cross-check), 2 = tcgen05 CTA-pair kernel
 *   (cta_group::2) wherever the operands allow it, 3 = tcgen05 single-CTA kernel                              */
int capdec_gemm(int32_t dtype, int32_t majors, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                int64_t ldc, int32_t M, int32_t N, int32_t K, const float* bias, const void* addend, int64_t ld_add,
                const void* dact_y, int64_t ld_y, int32_t act, int32_t out_f32, int32_t atomic, int32_t variant,
                void* stream);

/* Fused additive attention, model:103-113 (+ "context + p_hid" of model:115 when p_hid != NULL).
 * rows are addressed as  row(img, q) = img * row0_stride + q * q_stride.
 *   keys, vals [n_img,k,H]; hid_emb, sent_emb, sentinel, p_hid [rows,H] (dtype); w_alpha [H], b_alpha [1] f32
 *   alpha [rows,k+1] f32; actx [rows,H] = context + p_hid; ctx [rows,H] = context (either may be NULL)         */
int capdec_attention_fwd(int32_t dtype, int32_t H, int32_t k, int32_t n_img, int32_t q_per_img, int64_t row0_stride,
                         int64_t q_stride, int32_t softmax_axis, const void* keys, const void* vals,
                         const void* hid_emb, const void* sent_emb, const void* sentinel, const void* p_hid,
                         const float* w_alpha, const float* b_alpha, float* alpha, void* actx, void* ctx, void* stream);
/* Backward of the above given d(context) = dctx [rows,H].
 *   de [rows,k+1] f32 (scratch/out); dS_ctx [rows,H]; dKV [n_img*k, 2H]: cols [0,H) = d(values) * (1 - v^2) (pre-tanh),
 *   cols [H,2H) = d keys; dhid_emb, dsent_emb [rows,H]; dw_alpha [H], db_alpha [1] f32 (accumulated: zero them first) */
int capdec_attention_bwd(int32_t dtype, int32_t H, int32_t k, int32_t n_img, int32_t q_per_img, int64_t row0_stride,
                         int64_t q_stride, int32_t softmax_axis, const void* keys, const void* vals,
                         const void* hid_emb, const void* sent_emb, const void* sentinel, const float* w_alpha,
                         const float* alpha, const void* dctx, float* de, void* dS_ctx, void* dKV, void* dhid_emb,
                         void* dsent_emb, float* dw_alpha, float* db_alpha, void* stream);

/* Paddle lstm_unit pointwise cell (gate order i,f,o,g): gates_pre [R,4H] f32, c_prev [R,H] f32 ->
 * h_out [R,H] (dtype), c_out [R,H] f32 */
int capdec_lstm_cell_fwd(int32_t dtype, int32_t R, int32_t H, const float* gates_pre, const float* c_prev, void* h_out,
                         float* c_out, void* stream);

/* One beam-search selection step over logits [n_img*beam, ldl] f32 (see capdec_decode). */
int capdec_beam_step(const float* logits, int64_t ldl, int32_t V, int32_t n_img, int32_t beam, int32_t step,
                     int32_t T_total, int32_t pad, int32_t stop, int32_t finish_on_stop, const float* score_in,
                     float* score_out, const int32_t* fin_in, int32_t* fin_out, const int32_t* hist_in,
                     int32_t* hist_out, int32_t* tok_out, int32_t* src_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CAPDEC_H_ */
