// Host-side launchers of the non-GEMM kernels (templated on the storage type T in {float, bf16}).
// Row index convention for training tensors: time-major, m = t * B + b.
#pragma once
#include "common.cuh"

namespace capdec {

struct LaunchCtx {
  cudaStream_t stream;
  long* launches;
};

// ---- pointwise.cu -------------------------------------------------------------------------------
// One launch for a whole list of fp32 -> T strided copies (capdec_bind_params) / of fp32 buffers to clear.
struct ConvertJob {
  const float* src; void* dst;
  long lds, ldd;      // row strides (elements)
  int rows, cols;
  int mode;           // 0: dst[r, c] = src[r, c];  1: the gate-column permutation of recurrent.cu (cols = 4H)
  int H;
};
constexpr int kMaxConvertJobs = 24;
struct ConvertTable {
  ConvertJob job[kMaxConvertJobs];
  int first_block[kMaxConvertJobs + 1];
  int n = 0;
  void add(const float* src, long lds, void* dst, long ldd, int rows, int cols, int mode = 0, int H = 0) {
    job[n++] = ConvertJob{src, dst, lds, ldd, rows, cols, mode, H};
  }
};
// returns 0 when launched, 1 when some job is not 16-byte vectorisable (nothing was launched)
template <typename T> int multi_convert(LaunchCtx c, ConvertTable& tab);
constexpr int kMaxZeroJobs = 32;
struct ZeroTable {
  float* ptr[kMaxZeroJobs];
  long count[kMaxZeroJobs];
  int vec[kMaxZeroJobs];
  int first_block[kMaxZeroJobs + 1];
  int n = 0;
  void add(float* p, long cnt) { ptr[n] = p; count[n] = cnt; ++n; }
};
void multi_zero(LaunchCtx c, ZeroTable& tab);
template <typename T> void convert_from_f32(LaunchCtx c, const float* src, T* dst, long n);
// dst[r * ldd + cc] = src[r * lds + cc]   (strided copy + convert)
template <typename T> void copy2d_from_f32(LaunchCtx c, const float* src, long lds, T* dst, long ldd, int R, int C);
// out[(t*B+b), :] = ids[b*ids_ld + t] == pad ? 0 : emb[id]      (model:28-32, 84)
//   err: the handle's error word; an id outside [0, V) sets CAPDEC_ERR_SOURCE_ID there and reads as padding
template <typename T>
void embed_gather(LaunchCtx c, const T* emb, int E, int V, const int64_t* ids, long ids_ld, int B, int steps, int pad,
                  T* out, long ldo, int* err);
// LSTM pointwise cell (Paddle lstm_unit op, gate order i,f,o,g; model:87-88).
//   pre = gpre[r] (+ xe[r] if xe) (+ g5[(r % g5_mod) / g5_div] if g5);  c = sig(f) c_prev + sig(i) tanh(g); h = sig(o) tanh(c)
template <typename T>
void lstm_cell_fwd(LaunchCtx c, int R, int H, const float* gpre, long ld_gpre, const T* xe, long ld_xe, const float* g5,
                   long ld_g5, int g5_mod, int g5_div, const float* c_prev, T* h_out, long ld_h, float* c_out,
                   T* gates_act);
// sentinel (model:89-92): sg = sig(ph + xe + g5);  s = sg * tanh(c)
//   rowmap (optional, packed rows): row r of ph / sgate / s_out is row rowmap[r] of xe / g5 / cc
template <typename T>
void sentinel_fwd(LaunchCtx c, long R, int H, const T* ph, long ld_ph, const T* xe, long ld_xe, const float* g5,
                  long ld_g5, int g5_mod, int g5_div, const float* cc, T* sgate, T* s_out, const int* rowmap = nullptr);
// inference step: cell + sentinel from one fp32 pre-activation row gpre5 [R, 5H] = [gates i,f,o,g | sentinel gate] and the
// per-image term g5 [R / beam, 5H] (model:87-92); h_out, s_out [R, H] (T), c_out [R, H] fp32
template <typename T>
void lstm_cell_sentinel_fwd(LaunchCtx c, int R, int H, const float* gpre5, const float* g5, int beam, const float* c_prev,
                            T* h_out, float* c_out, T* s_out);
// ---- packed rows of the time-batched head (see pointwise.cu) ----
//   out: steps [B] live steps per sample (from the captions); prow [T*B] padded -> packed row or -1; pm [Mp] packed ->
//   padded row (zero it first: entries the captions do not fill must stay valid rows).  Mp = the host's row count.
int build_row_map(LaunchCtx c, const int64_t* caption, long cap_ld, int B, int T, int pad, long Mp, int* steps, int* prow,
                  int* pm, int* err);
// dst{0,1}[r, :] = src{0,1}[pm[r], :]   (src1 / dst1 optional);  N % 8 == 0
template <typename T>
void gather_rows(LaunchCtx c, const T* src0, const T* src1, long ld_src, const int* pm, long Mp, int N, T* dst0, T* dst1,
                 long ld_dst);
// dst{0,1}[m, :] = prow[m] >= 0 ? src{0,1}[prow[m], :] : 0   (all pitches N)
template <typename T>
void scatter_rows_zero(LaunchCtx c, const T* src0, const T* src1, const int* prow, long M, int N, T* dst0, T* dst1);
// sentinel_bwd over the M padded rows with packed operands dS, sgate [*, H]; also writes the packed copy dsg_packed
template <typename T>
void sentinel_bwd_packed(LaunchCtx c, long M, int H, const int* prow, const T* dS, const T* sgate, const float* cc,
                         T* dsg_pre, long ld_dsg, float* dC_head, T* dsg_packed);
// dsg_pre = dS * tanh(c) * sg (1 - sg)  -> dG[:, 4H:5H];   dC_head = dS * sg * (1 - tanh(c)^2)
template <typename T>
void sentinel_bwd(LaunchCtx c, long R, int H, const T* dS, const T* sgate, const float* cc, T* dsg_pre, long ld_dsg,
                  float* dC_head);
// one BPTT step of the cell (see pointwise.cu)
template <typename T>
void lstm_cell_bwd(LaunchCtx c, int B, int H, const T* dH_head, const T* dHprev_head, const float* dh_rec,
                   const float* dC_head, float* dc_rec, const T* gates_act, const float* c_t, const float* c_prev,
                   T* dgates, long ld_dg);
// out[b, :] = sum_t x[(t*B+b), :]   (live [B] optional: only the first live[b] steps of sample b are read, the rest are zero)
template <typename T>
void sum_over_steps(LaunchCtx c, const T* x, long ld, int B, int steps, int N, T* out, long ldo, const int* live = nullptr);
// out[n] += sum_m x[m, n]    (fp32 atomics)
// out[n] += scale * sum_m x[m, n]   (fp32 atomics; scale read from a device scalar, or 1)
template <typename T> void colsum_add(LaunchCtx c, const T* x, long ld, long M, int N, float* out, const float* scale_dev = nullptr);
// several out[n] += sum_m x[m, n] in one launch (falls back to one colsum_add per job for small / unaligned operands)
struct ColsumJob {
  const void* x; long ld; long M; int N; float* out; long rows_per_chunk;
};
constexpr int kMaxColsumJobs = 8;
struct ColsumTable {
  ColsumJob job[kMaxColsumJobs];
  int first_block[kMaxColsumJobs + 1];
  int n = 0;
  void add(const void* x, long ld, long M, int N, float* out) { job[n++] = ColsumJob{x, ld, M, N, out, 0}; }
};
template <typename T> void colsum_multi(LaunchCtx c, ColsumTable& tab);
// the fused bias vectors of capdec_bind_params: b_vocab [Vp] = out_fc_bias padded with zeros; b5 = [lstm_b | fc_xs_b];
// b5d = [lstm_b | fc_xs_b + fc_hs_b]; b5p (optional) = b5 with unit-major gate columns
struct BiasBind {
  const float* out_fc_bias; int V, Vp; float* b_vocab;
  const float* lstm_b; const float* fc_xs_b; const float* fc_hs_b; int H;
  float* b5; float* b5d; float* b5p;
};
void bind_biases(LaunchCtx c, const BiasBind& a);
// dEmb[id] += dX[(t*B+b)]  for id != pad        (lookup_table_v2_grad, padding_idx)
// pm (optional): dX holds `rows` packed rows, row r belongs to (t, b) = pm[r] / B, pm[r] % B
template <typename T>
void embed_scatter_add(LaunchCtx c, const T* dX, long ld, int E, int V, const int64_t* ids, long ids_ld, int B, int steps,
                       int pad, float* dEmb, const int* pm = nullptr, long rows = 0);

// ---- adaptor.cu ----------------------------------------------------------------------------------
// feature adaptor layout kernels (model:193-199): enc_map [B,X,k] fp32 -> rows [B*k,X], mean [B,X]; and the backward
//   d map[b,x,j] = d rows[b*k+j, x] + d mean[b,x] / k.   X % 8 == 0, k <= 64.  Return 0 / -1.
template <typename T> int map_to_rows(LaunchCtx c, const float* map, int B, int X, int k, T* rows, T* mean);
template <typename T> int rows_to_map(LaunchCtx c, const T* drows, const float* dmean, int B, int X, int k, float* dmap);
template <typename T> void convert_to_f32(LaunchCtx c, const T* src, float* dst, long n);   // n % 8 == 0
// PaddlePaddle 1.8 Adam over one flat fp32 buffer (see adaptor.cu); clip > 0: GradientClipByValue(clip) on the gradient
int adam_step(cudaStream_t s, float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
              float eps, float beta1_pow, float beta2_pow, float clip);

// ---- attention.cu ------------------------------------------------------------------------------
struct AttnShape {
  int H, k;             // hidden size, number of image positions (k+1 attention positions)
  int n_img;            // images
  int q_per_img;        // query rows per image (train: T steps; decode: beam)
  long row0_stride;     // row of query 0 of image i = i * row0_stride
  long q_stride;        // row of query q = row0 + q * q_stride      (train: 1 / B; decode: beam / 1)
  int position_softmax; // 1 = softmax over the k+1 positions; 0 = reference as written (alpha == 1)
  // Packed rows (training with lengths): only queries q < nq[img] of an image are live, and the per-query [rows, H]
  // operands (hid_emb, sent_emb, sentinel, p_hid and their gradients) hold the live rows only, row prow[padded row];
  // alpha, alpha_T, de, dal, sadd and the operands of the batched GEMMs (the context output, d context) keep the padded
  // row index.  pm = inverse map, packed_rows = number of live rows.  All nullptr / 0: every row is live.
  const int* prow = nullptr;
  const int* pm = nullptr;
  const int* nq = nullptr;
  long packed_rows = 0;
};
// fused additive attention forward (model:103-113 + the "context + p_hid" of model:115)
//   alpha [rows, k+1] fp32;  actx [rows, H] = mean_j(alpha_j v_j) + p_hid;  ctx_out (optional) without p_hid
template <typename T>
void attention_fwd(LaunchCtx c, const AttnShape& s, const T* keys, const T* vals, const T* hid_emb, const T* sent_emb,
                   const T* sentinel, const T* p_hid, const float* w_alpha, const float* b_alpha, float* alpha, T* actx,
                   T* ctx_out);
struct GemmContext;
// training-shaped forward on the tensor cores (bf16; see attention.cu): scores + soft-max kernel, then the context as a
// batched GEMM.  scratch: alpha_T [rows, kAlphaTLd] bf16 (alpha / (k+1), reusable by the backward pass), sadd [rows, H]
bool attention_fwd_batched_supported(const AttnShape& s);
int attention_fwd_batched(GemmContext& gc, LaunchCtx c, const AttnShape& s, const bf16* keys, const bf16* vals,
                          const bf16* hid_emb, const bf16* sent_emb, const bf16* sentinel, const bf16* p_hid,
                          const float* w_alpha, const float* b_alpha, float* alpha, bf16* alpha_T, bf16* sadd,
                          bf16* actx);
// inference-shaped forward (q_per_img <= 8): the same scores kernel + a small context kernel
bool attention_fwd_small_supported(const AttnShape& s);
int attention_fwd_small(LaunchCtx c, const AttnShape& s, const bf16* keys, const bf16* vals, const bf16* hid_emb,
                        const bf16* sent_emb, const bf16* sentinel, const bf16* p_hid, const float* w_alpha,
                        const float* b_alpha, float* alpha, bf16* actx);
// backward (see attention.cu): two batched tensor-core GEMMs (d alpha = dctx.V^T, d values = alpha^T.dctx with tanh'
// fused), a row-wise soft-max-backward kernel and the column-wise d keys kernel.  dactx [rows,H] = d context.
//   scratch: dal [rows, kDalLd] fp32, alpha_T [rows, kAlphaTLd] T
//   de [rows,k+1] fp32;  dS_ctx [rows,H] = alpha_k dctx / (k+1);  dKV [n_img*k, 2H]: cols 0..H = d(values pre-tanh),
//   cols H..2H = d keys;  dhid_emb, dsent_emb [rows,H];  dw_alpha [H], db_alpha [1] fp32 (+=)
constexpr int kDalLd = 68;      // >= 63 positions, 16-byte-multiple pitch in fp32
constexpr int kAlphaTLd = 72;   // >= 63 positions, 16-byte-multiple pitch in bf16
template <typename T>
int attention_bwd(GemmContext& gc, LaunchCtx c, const AttnShape& s, const T* keys, const T* vals, const T* hid_emb,
                  const T* sent_emb, const T* sentinel, const float* w_alpha, const float* alpha, const T* dactx,
                  float* dal, T* alpha_T, float* de, T* dS_ctx, T* dKV, T* dhid_emb, T* dsent_emb, float* dw_alpha,
                  float* db_alpha);

// ---- recurrent.cu (bf16 only) -------------------------------------------------------------------
// The whole teacher-forced recurrence as one persistent cooperative kernel per direction (see recurrent.cu).
//   *_supported: the shape allows the persistent kernels; otherwise (or with CAPDEC_FLAG_STEPWISE_RECURRENCE) the
//   caller runs the per-step GEMM + cell schedule.  ctr: 64 unsigned arrival counters owned by the caller.
bool lstm_seq_supported(int B, int H);
// dst[k, 64*j + 4*u + gate] = src[k, gate*H + 16*j + u]: the forward kernel's unit-major weight copy
void permute_gate_cols(LaunchCtx c, const float* src, long lds, bf16* dst, int H);
// xe [T*B, ld_xe]: complete x-part of the gate pre-activations in columns [0, 4H), unit-major (col 4*unit + gate)
int lstm_seq_fwd(LaunchCtx c, int B, int H, int T, const bf16* w_perm, const bf16* xe, long ld_xe, bf16* hs, float* cs,
                 bf16* gact, unsigned* ctr);
// dhrec2 [2,B,H] and dcrec [B,H] fp32 scratch (zeroed here)
int lstm_seq_bwd(LaunchCtx c, int B, int H, int T, const bf16* w_hh, const bf16* dHhead, const bf16* dHprev,
                 const float* dChead, const bf16* gact, const float* cs, bf16* dG, long ld_dg, float* dhrec2,
                 float* dcrec, unsigned* ctr);

// ---- loss.cu -----------------------------------------------------------------------------------
// masked soft-max cross entropy (model:163-171, 180-183, 205-212): target[t*B+b] = caption[b, t+1]
//   lse[m];  acc[0] += sum(ce * mask), acc[1] += sum(mask);  and the UNSCALED gradient
//   dlogits[m, n] = (softmax - onehot) * mask  (columns [V, Vp) zeroed); the 1/sum(mask) * loss_scale factor is applied
//   by the consumers (GemmEpi::alpha_dev, colsum_add scale).  err: see embed_gather (CAPDEC_ERR_TARGET_ID)
//   rows / rowmap: the number of logit rows and (packed rows) the padded row t*B+b of each of them; nullptr: rows = B*steps
template <typename T>
void ce_fused(LaunchCtx c, const float* logits, long ldl, int V, int Vp, const int64_t* caption, long cap_ld, int B,
              int steps, int pad, float* lse, float* acc, T* dlogits, int* err, long rows = -1,
              const int* rowmap = nullptr);
// out[0] = in[0] * s
void scale_scalar(LaunchCtx c, const float* in, float s, float* out);
// acc[2] = acc[0] / acc[1] (the loss), acc[3] = 1 / acc[1]
void loss_finalize(LaunchCtx c, float* acc);
// out[b, t, :V] = logits[(t*B+b), :V]
void logits_to_batch_major(LaunchCtx c, const float* logits, long ldl, int V, int B, int steps, float* out);
// one beam-search step for every image; see loss.cu.  V <= beam_max_vocab(), beam <= 8.
//   scratch: beam_scratch_bytes(n_img, beam) bytes of device memory owned by the caller (the decoder handle)
//   trace (optional): [n_img*beam, 2] = (token, back-pointer) of this step's selection
size_t beam_scratch_bytes(int n_img, int beam);
int beam_max_vocab();
void beam_step(LaunchCtx c, const float* logits, long ldl, int V, int n_img, int beam, int step, int T_total, int pad,
               int stop, int finish_on_stop, const float* score_in, float* score_out, const int* fin_in, int* fin_out,
               const int* hist_in, int* hist_out, int* tok_out, int* src_out, void* scratch, int* trace);
// initial beam state: score = 0 for prefix 0 of every image, -inf for the others; tok = start; src = r % beam
void beam_init(LaunchCtx c, int R, int beam, int start, float* score, int* tok, int* src);
// next-step operands: xin[r] = [emb[tok[r]] | h_new[src_row(r)]], c_next[r] = c_new[src_row(r)]
template <typename T>
void beam_reorder(LaunchCtx c, int n_img, int beam, int E, int H, int pad, const T* emb, const int* tok, const int* src,
                  const T* h_new, const float* c_new, T* xin, long ld_xin, float* c_next);

}  // namespace capdec
