// The decoder handle: operand layouts, workspace, and the three schedules (teacher-forced forward, backward,
// greedy / beam inference) that replace the PaddlePaddle While-op interpreter loop of
// model_adaAttention_aic.py:72-127 and its auto-generated while_grad.
//
// Training is "recurrent core + time-batched head" (SURVEY.md F7): only h_{t-1}.W_hh and the cell are
// sequential; the word-embedding half of the gate GEMM is hoisted over all T steps and everything downstream of
// (h_t, c_t) -- sentinel, attention, context, out, tied vocabulary projection, loss -- runs once at
// M = B*T rows.  Rows of every [M, *] activation are time-major: m = t*B + b.
#include <climits>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/capdec.h"
#include "gemm.cuh"
#include "kernels.cuh"

namespace capdec {

thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof g_err, fmt, ap);
  va_end(ap);
}

#define CUDA_OK(expr)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return -1;                                                                              \
    }                                                                                         \
  } while (0)
#define OK(expr)                \
  do {                          \
    if ((expr) != 0) return -1; \
  } while (0)

// bumped whenever a workspace buffer is (re)allocated: captured decode graphs hold raw pointers and are dropped then
static unsigned long g_ws_epoch = 0;

// a growable device buffer owned by the handle
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes, size_t* total) {
    if (bytes <= cap) return 0;
    if (p) { cudaFree(p); *total -= cap; p = nullptr; cap = 0; }
    bytes = (bytes + 255) & ~size_t(255);
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) { set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return -1; }
    cap = bytes;
    *total += bytes;
    ++g_ws_epoch;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <typename U> U* as() const { return reinterpret_cast<U*>(p); }
};

// optional device-side phase timeline (capdec_profile): CUDA events recorded on the caller's stream at phase
// boundaries; report() sums the elapsed time per phase name.  Costs nothing when disabled.
struct PhaseTimer {
  bool on = false;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;
  std::vector<cudaEvent_t> pool;
  void mark(const char* name, cudaStream_t s) {
    if (!on) return;
    cudaEvent_t e;
    if (!pool.empty()) { e = pool.back(); pool.pop_back(); }
    else if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, s);
    marks.emplace_back(name, e);
  }
  // "name ms\n" per phase (first-seen order), the time of a phase = from the previous mark to its own mark
  std::string report() {
    std::vector<std::pair<const char*, double>> agg;
    if (!marks.empty()) cudaEventSynchronize(marks.back().second);
    for (size_t i = 1; i < marks.size(); ++i) {
      float ms = 0.f;
      if (marks[i].first[0] == '^') continue;  // a '^' mark only restarts the clock
      cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second);
      bool found = false;
      for (auto& a : agg)
        if (!strcmp(a.first, marks[i].first)) { a.second += ms; found = true; break; }
      if (!found) agg.emplace_back(marks[i].first, ms);
    }
    for (auto& m : marks) pool.push_back(m.second);
    marks.clear();
    std::string out;
    char line[160];
    for (auto& a : agg) { snprintf(line, sizeof line, "%s %.6f\n", a.first, a.second); out += line; }
    return out;
  }
  ~PhaseTimer() {
    for (auto& m : marks) cudaEventDestroy(m.second);
    for (auto e : pool) cudaEventDestroy(e);
  }
};

class DecoderBase {
 public:
  virtual ~DecoderBase() {}
  PhaseTimer prof;
  virtual int bind(const capdec_params* p, cudaStream_t s) = 0;
  virtual int train_fwd(const float* g, const float* P, const int64_t* cap, const int32_t* lengths, int B,
                        float* logits_out, float* loss_out, cudaStream_t s) = 0;
  virtual int train_bwd(float loss_scale, const float* inv_ntok, float* dg, float* dP, const capdec_grads* gr,
                        cudaStream_t s) = 0;
  virtual int decode(const float* g, const float* P, int B, int beam, int finish, int32_t* ids, float* scores,
                     float* ids_f32, cudaStream_t s) = 0;
  virtual int decode_trace(int32_t* out, cudaStream_t s) = 0;
  virtual int bind_adaptor(int X, const capdec_adaptor_params* p, cudaStream_t s) = 0;
  virtual int img2feature(const float* enc_map, int B, float* P_out, float* g_out, cudaStream_t s) = 0;
  virtual int img2feature_bwd(float* d_enc_map, const capdec_adaptor_grads* gr, cudaStream_t s) = 0;
  size_t ws_bytes = 0;
  long launches = 0;
  int device = 0;
  // Device-side input checks (token ids out of range, lengths that contradict the captions) cannot fail the asynchronous
  // call that launched them: the kernels write a bit mask into this word of MAPPED host memory (only when something is
  // wrong: no traffic otherwise), the next call on the handle and capdec_status report it.
  int* err_flag = nullptr;
  int take_async_error() {
    if (!err_flag) return 0;
    const int e = *reinterpret_cast<volatile int*>(err_flag);
    if (e) *reinterpret_cast<volatile int*>(err_flag) = 0;
    return e;
  }
  // caller-owned events recorded by train_bwd when a group of gradients is final (capdec_set_grad_ready_events)
  cudaEvent_t grad_ready[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  void mark_ready(int i, cudaStream_t s) {
    if (grad_ready[i]) cudaEventRecord(grad_ready[i], s);
  }
};

template <typename T> class Decoder : public DecoderBase {
 public:
  explicit Decoder(const capdec_config& c) : cfg(c) {
    H = c.hidden_dim; E = c.embedding_size; V = c.vocab_size; k = c.num_positions;
    S = c.sentence_length; Tt = S - 1; Ti = c.infer_max_length;
    Vp = ((V + 127) / 128) * 128;
    stepwise = (c.flags & CAPDEC_FLAG_STEPWISE_RECURRENCE) != 0;
  }
  ~Decoder() override {
    dgraph.reset();
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (side_stream) cudaStreamDestroy(side_stream);
    if (ev_fork) cudaEventDestroy(ev_fork);
    if (ev_join) cudaEventDestroy(ev_join);
    for (DevBuf* b : all_bufs()) b->release();
  }

  // A second, handle-owned stream: independent kernel chains of one schedule run beside each other.  Most GEMMs of the
  // time-batched head are 80 tiles on 74 CTA pairs (two waves, the second almost empty) and the weight gradients use 128
  // of the 148 SMs: a neighbour on the other stream fills what they leave idle.  fork(): the side stream waits for
  // everything enqueued on `s` so far; join(): `s` waits for the side stream.  Without the stream both are no-ops and the
  // side chain runs in order on `s`.  (Inside a stream capture the side stream joins the capture.)  Measured: switching
  // programmatic dependent launch off between fork and join changes nothing (2.429 vs 2.431 ms per step).
  cudaStream_t fork(cudaStream_t s) {
    if (!side_stream) {
      if (cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking) != cudaSuccess) side_stream = nullptr;
      if (cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess) ev_fork = nullptr;
      if (cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) ev_join = nullptr;
      cudaGetLastError();
    }
    if (!side_stream || !ev_fork || !ev_join) return s;
    if (cudaEventRecord(ev_fork, s) != cudaSuccess || cudaStreamWaitEvent(side_stream, ev_fork, 0) != cudaSuccess) {
      cudaGetLastError();
      return s;
    }
    return side_stream;
  }
  int join(cudaStream_t s, cudaStream_t side) {
    if (side == s) return 0;
    CUDA_OK(cudaEventRecord(ev_join, side));
    CUDA_OK(cudaStreamWaitEvent(s, ev_join, 0));
    return 0;
  }

  // ------------------------------------------------------------------------------------------------------------
  int bind(const capdec_params* p, cudaStream_t s) override {
    prm = *p;
    ++g_ws_epoch;  // the caller's bias / w_alpha pointers are baked into captured decode graphs
    LaunchCtx lc{s, &launches};
    const size_t e = sizeof(T);
    const int X = E + H;
    prof.mark("^bind", s);
    OK(w_kv_n.ensure(e * H * 2 * H, &ws_bytes));
    OK(w_x_n.ensure(e * E * 5 * H, &ws_bytes));
    OK(w_g_n.ensure(e * H * 5 * H, &ws_bytes));
    OK(w_hh_n.ensure(e * H * 4 * H, &ws_bytes));
    OK(w_hs_n.ensure(e * H * H, &ws_bytes));
    OK(w_p_n.ensure(e * H * H, &ws_bytes));
    OK(w_g2_n.ensure(e * H * H, &ws_bytes));
    OK(w_s_n.ensure(e * H * H, &ws_bytes));
    OK(w_o_n.ensure(e * H * H, &ws_bytes));
    OK(w_e_n.ensure(e * H * E, &ws_bytes));
    OK(emb.ensure(e * (size_t)V * E, &ws_bytes));
    OK(w_dec_n.ensure(e * (size_t)X * 5 * H, &ws_bytes));
    OK(b5.ensure(4 * 5 * H, &ws_bytes)); OK(b5d.ensure(4 * 5 * H, &ws_bytes));
    OK(b_vocab.ensure(4 * (size_t)Vp, &ws_bytes));

    // Every weight keeps PaddlePaddle's own [in, out] layout (one storage-type copy each, no transposes):
    //   forward  Y = X.W      reads it as the MN-major B operand,
    //   backward dX = dY.W^T  reads the same buffer as the K-major B operand (N = in, K = out).
    const float* lw_g = p->lstm_w + (size_t)E * 4 * H;   // global-image rows of lstm_w
    const float* lw_h = p->lstm_w + (size_t)X * 4 * H;   // hidden rows of lstm_w
    bool want_perm = false;
    if constexpr (std::is_same<T, bf16>::value) {
      if (!stepwise && lstm_seq_supported(1, H)) {  // unit-major copy for the persistent forward recurrence
        OK(w_hh_perm.ensure(e * H * 4 * H, &ws_bytes));
        OK(w_x_p.ensure(e * E * 5 * H, &ws_bytes));
        OK(w_g_p.ensure(e * H * 5 * H, &ws_bytes));
        OK(b5p.ensure(4 * 5 * H, &ws_bytes));
        want_perm = true;
      }
    }
    T* wd = w_dec_n.as<T>();
    ConvertTable tab;
    tab.add(p->fc_v_w, H, w_kv_n.as<T>(), 2 * H, H, H);                      // [H, 2H] = [fc_v | fc_k]
    tab.add(p->fc_k_w, H, w_kv_n.as<T>() + H, 2 * H, H, H);
    tab.add(p->lstm_w, 4 * H, w_x_n.as<T>(), 5 * H, E, 4 * H);              // word rows: [E, 5H]
    tab.add(p->fc_xs_w, H, w_x_n.as<T>() + 4 * H, 5 * H, E, H);
    tab.add(lw_g, 4 * H, w_g_n.as<T>(), 5 * H, H, 4 * H);                    // global rows: [H, 5H]
    tab.add(p->fc_xs_w + (size_t)E * H, H, w_g_n.as<T>() + 4 * H, 5 * H, H, H);
    tab.add(lw_h, 4 * H, w_hh_n.as<T>(), 4 * H, H, 4 * H);                   // hid rows: [H, 4H]
    tab.add(p->fc_hs_w, H, w_hs_n.as<T>(), H, H, H);
    tab.add(p->fc_p_w, H, w_p_n.as<T>(), H, H, H);
    tab.add(p->fc_g_w, H, w_g2_n.as<T>(), H, H, H);
    tab.add(p->fc_s_w, H, w_s_n.as<T>(), H, H, H);
    tab.add(p->fc_o_w, H, w_o_n.as<T>(), H, H, H);
    tab.add(p->fc_e_w, E, w_e_n.as<T>(), E, H, E);
    tab.add(p->word_embedding, E, emb.as<T>(), E, V, E);
    // inference operand [E+H, 5H]: rows = [word_emb ; h_{t-1}], cols = [lstm gates | sentinel pre-activation]
    tab.add(p->lstm_w, 4 * H, wd, 5 * H, E, 4 * H);
    tab.add(p->fc_xs_w, H, wd + 4 * H, 5 * H, E, H);
    tab.add(lw_h, 4 * H, wd + (size_t)E * 5 * H, 5 * H, H, 4 * H);
    tab.add(p->fc_hs_w, H, wd + (size_t)E * 5 * H + 4 * H, 5 * H, H, H);
    if (want_perm) {
      // unit-major gate columns (col 4*unit + gate) for the persistent recurrence: W_hh, and the x-side weights whose
      // GEMMs produce the pre-activation tile the kernel loads by TMA
      tab.add(lw_h, 4 * H, w_hh_perm.p, 4 * H, H, 4 * H, 1, H);
      tab.add(p->lstm_w, 4 * H, w_x_p.p, 5 * H, E, 4 * H, 1, H);
      tab.add(p->fc_xs_w, H, w_x_p.as<T>() + 4 * H, 5 * H, E, H);
      tab.add(lw_g, 4 * H, w_g_p.p, 5 * H, H, 4 * H, 1, H);
      tab.add(p->fc_xs_w + (size_t)E * H, H, w_g_p.as<T>() + 4 * H, 5 * H, H, H);
    }
    const int rc = multi_convert<T>(lc, tab);
    if (rc < 0) return -1;
    if (rc > 0) {  // unaligned caller buffers: per-tensor scalar kernels
      for (int j = 0; j < tab.n; ++j) {
        const ConvertJob& jb = tab.job[j];
        if (jb.mode == 0) copy2d_from_f32<T>(lc, jb.src, jb.lds, reinterpret_cast<T*>(jb.dst), jb.ldd, jb.rows, jb.cols);
        else permute_gate_cols(lc, jb.src, jb.lds, reinterpret_cast<bf16*>(jb.dst), H);
      }
    }
    // out_fc_bias padded to Vp columns so the GEMM epilogue may read it in whole groups of 8, and the fused bias vectors:
    // train [lstm_b | fc_xs_b]; decode [lstm_b | fc_xs_b + fc_hs_b]; the unit-major copy for the persistent recurrence
    BiasBind bb;
    bb.out_fc_bias = p->out_fc_bias; bb.V = V; bb.Vp = Vp; bb.b_vocab = b_vocab.as<float>();
    bb.lstm_b = p->lstm_b; bb.fc_xs_b = p->fc_xs_b; bb.fc_hs_b = p->fc_hs_b; bb.H = H;
    bb.b5 = b5.as<float>(); bb.b5d = b5d.as<float>(); bb.b5p = want_perm ? b5p.as<float>() : nullptr;
    bind_biases(lc, bb);
    CUDA_OK(cudaGetLastError());
    prof.mark("bind_params", s);
    bound = true;
    return 0;
  }

  // ------------------------------------------------------------------------------------------------------------
  int image_projections(GemmContext& gc, LaunchCtx lc, const float* g, const float* P, int B) {
    const size_t e = sizeof(T);
    if (!g && !P) {
      // features staged by capdec_img2feature: gT / PT already hold them in the storage type
      if (!staged || staged_B != B) {
        set_error("no staged features for %d images: call capdec_img2feature first (or pass the features)", B);
        return -1;
      }
    } else if (!g || !P) {
      set_error("global_img_feat and p_img_feat must both be given, or both be NULL (staged features)");
      return -1;
    }
    OK(gT.ensure(e * (size_t)B * H, &ws_bytes));
    OK(PT.ensure(e * (size_t)B * k * H, &ws_bytes));
    OK(keys.ensure(e * (size_t)B * k * H, &ws_bytes));
    OK(vals.ensure(e * (size_t)B * k * H, &ws_bytes));
    if (g) {
      convert_from_f32<T>(lc, g, gT.as<T>(), (long)B * H);
      convert_from_f32<T>(lc, P, PT.as<T>(), (long)B * k * H);
      staged = false;
    }
    GemmCall c;
    c.A = PT.p; c.lda = H; c.M = B * k; c.N = H; c.K = H; c.ldb = 2 * H; c.b_mn = true; c.ldc = H;
    c.B = w_kv_n.p; c.C = vals.p; c.epi.bias = prm.fc_v_b; c.epi.act = ACT_TANH;            // model:52
    OK(gemm<T>(gc, c));
    c.B = w_kv_n.as<T>() + H; c.C = keys.p; c.epi.bias = prm.fc_k_b; c.epi.act = ACT_NONE;  // model:53
    OK(gemm<T>(gc, c));
    return 0;
  }

  // Y[M,N] = act(X . W + b)  with W a (sub-block of a) [K = in, N = out] weight copy, row stride ldw
  int fc(GemmContext& gc, const void* X, long ldx, const T* W, long ldw, const float* bias, void* Y, long ldy, int M,
         int N, int K, int act, bool out_f32 = false) {
    GemmCall c;
    c.A = X; c.lda = ldx; c.B = W; c.ldb = ldw; c.b_mn = true; c.C = Y; c.ldc = ldy; c.M = M; c.N = N; c.K = K;
    c.epi.bias = bias; c.epi.act = act; c.epi.out_f32 = out_f32 ? 1 : 0;
    return gemm<T>(gc, c);
  }
  // dW[in,out] += X[rows,in]^T . dY[rows,out]   (fp32 atomics, split-K over rows)
  int wgrad(GemmContext& gc, const void* X, long ldx, int n_in, const void* dY, long ldy, int n_out, int rows,
            float* dW, long ldw, const float* alpha_dev = nullptr) {
    if (!dW) return 0;
    GemmCall c;
    c.a_mn = true; c.b_mn = true;
    c.A = X; c.lda = ldx; c.B = dY; c.ldb = ldy; c.C = dW; c.ldc = ldw; c.M = n_in; c.N = n_out; c.K = rows;
    c.epi.atomic = 1; c.epi.out_f32 = 1; c.splits = 0; c.epi.alpha_dev = alpha_dev;
    return gemm<T>(gc, c);
  }

  // ------------------------------------------------------------------------------------------------------------
  // Feature adaptor, ImageCaptionModel._img2feature from the encoder map on (model:193-199)
  int bind_adaptor(int X, const capdec_adaptor_params* p, cudaStream_t s) override {
    if (X <= 0 || X % 8) { set_error("capdec_bind_adaptor: encoder_channel must be a positive multiple of 8"); return -1; }
    LaunchCtx lc{s, &launches};
    const size_t e = sizeof(T);
    Xc = X; aprm = *p;
    OK(w_img_n.ensure(e * (size_t)X * H, &ws_bytes));
    OK(w_glob_n.ensure(e * (size_t)X * H, &ws_bytes));
    ConvertTable tab;   // [in = X, out = H] kept as is: MN-major B operand forward, K-major B operand backward
    tab.add(p->fc_img_w, H, w_img_n.as<T>(), H, X, H);
    tab.add(p->fc_glob_w, H, w_glob_n.as<T>(), H, X, H);
    const int rc = multi_convert<T>(lc, tab);
    if (rc < 0) return -1;
    if (rc > 0)
      for (int j = 0; j < tab.n; ++j)
        copy2d_from_f32<T>(lc, tab.job[j].src, tab.job[j].lds, reinterpret_cast<T*>(tab.job[j].dst), tab.job[j].ldd,
                           tab.job[j].rows, tab.job[j].cols);
    CUDA_OK(cudaGetLastError());
    adaptor_bound = true;
    return 0;
  }

  int img2feature(const float* enc_map, int B, float* P_out, float* g_out, cudaStream_t s) override {
    if (!adaptor_bound) { set_error("capdec_img2feature: adaptor parameters not bound"); return -1; }
    if (B <= 0) { set_error("capdec_img2feature: empty batch"); return -1; }
    GemmContext gc; gc.stream = s;
    LaunchCtx lc{s, &launches};
    const size_t e = sizeof(T);
    const int X = Xc;
    prof.mark("^adaptor", s);
    OK(gT.ensure(e * (size_t)B * H, &ws_bytes));
    OK(PT.ensure(e * (size_t)B * k * H, &ws_bytes));
    OK(imgT.ensure(e * (size_t)B * k * X, &ws_bytes));
    OK(meanT.ensure(e * (size_t)B * X, &ws_bytes));
    // reshape + transpose (model:195-196) and reduce_mean over the positions (model:198): the map is read once
    OK(map_to_rows<T>(lc, enc_map, B, X, k, imgT.as<T>(), meanT.as<T>()));
    OK(fc(gc, imgT.p, X, w_img_n.as<T>(), H, aprm.fc_img_b, PT.p, H, B * k, H, X, ACT_RELU));      // model:197
    OK(fc(gc, meanT.p, X, w_glob_n.as<T>(), H, aprm.fc_glob_b, gT.p, H, B, H, X, ACT_RELU));       // model:199
    if (P_out) convert_to_f32<T>(lc, PT.as<T>(), P_out, (long)B * k * H);
    if (g_out) convert_to_f32<T>(lc, gT.as<T>(), g_out, (long)B * H);
    prof.mark("adaptor.fwd", s);
    launches += gc.launches;
    CUDA_OK(cudaGetLastError());
    staged = true; staged_B = B; have_dz = false;
    return 0;
  }

  int img2feature_bwd(float* d_enc_map, const capdec_adaptor_grads* gr, cudaStream_t s) override {
    if (!have_dz || !staged) {
      set_error("capdec_img2feature_bwd: needs capdec_img2feature -> capdec_train_fwd(NULL features) -> capdec_train_bwd first");
      return -1;
    }
    GemmContext gc; gc.stream = s;
    LaunchCtx lc{s, &launches};
    const size_t e = sizeof(T);
    const int X = Xc, B = staged_B;
    const capdec_adaptor_grads& G = *gr;
    prof.mark("^adaptor_bwd", s);
    ZeroTable zt;
    zt.add(G.fc_img_w, (long)X * H); zt.add(G.fc_img_b, H); zt.add(G.fc_glob_w, (long)X * H); zt.add(G.fc_glob_b, H);
    multi_zero(lc, zt);
    colsum_add<T>(lc, dzp.as<T>(), H, (long)B * k, H, G.fc_img_b);
    OK(wgrad(gc, imgT.p, X, X, dzp.p, H, H, B * k, G.fc_img_w, H));
    colsum_add<T>(lc, dzg.as<T>(), H, (long)B, H, G.fc_glob_b);
    OK(wgrad(gc, meanT.p, X, X, dzg.p, H, H, B, G.fc_glob_w, H));
    if (d_enc_map) {
      // gradient handed to the encoder (config.py:50 encoder_trainable): through both fc layers, then back to [B, X, k]
      OK(dimg.ensure(e * (size_t)B * k * X, &ws_bytes));
      OK(dmean.ensure(4 * (size_t)B * X, &ws_bytes));
      GemmCall c;
      c.A = dzp.p; c.lda = H; c.B = w_img_n.p; c.ldb = H; c.C = dimg.p; c.ldc = X; c.M = B * k; c.N = X; c.K = H;
      OK(gemm<T>(gc, c));
      GemmCall d;
      d.A = dzg.p; d.lda = H; d.B = w_glob_n.p; d.ldb = H; d.C = dmean.p; d.ldc = X; d.M = B; d.N = X; d.K = H;
      d.epi.out_f32 = 1;
      OK(gemm<T>(gc, d));
      OK(rows_to_map<T>(lc, dimg.as<T>(), dmean.as<float>(), B, X, k, d_enc_map));
    }
    prof.mark("adaptor.bwd", s);
    launches += gc.launches;
    CUDA_OK(cudaGetLastError());
    return 0;
  }

  int train_fwd(const float* g, const float* P, const int64_t* cap, const int32_t* lengths, int B, float* logits_out,
                float* loss_out, cudaStream_t s) override {
    if (!bound) { set_error("capdec_train_fwd: parameters not bound"); return -1; }
    if (B <= 0) { set_error("capdec_train_fwd: empty batch"); return -1; }
    GemmContext gc; gc.stream = s;
    LaunchCtx lc{s, &launches};
    const size_t e = sizeof(T);
    const long M = (long)B * Tt;
    last_B = B; last_cap = cap;
    fwd_staged = (g == nullptr); have_dz = false; have_fwd = false;
    // Packed rows: with `lengths` the time-batched head runs only on the rows (t, b) that carry a target,
    // t < lengths[b] - 1 (captions are [start, w.., stop, pad..]): Mp of the M = B*T rows.  The per-step logits of the dead
    // rows would be needed for logits_out, so the full schedule runs when they are requested; it also runs when there is
    // (almost) nothing to drop.
    long Mp = M;
    if (lengths && !logits_out) {
      Mp = 0;
      for (int b = 0; b < B; ++b) Mp += std::min<long>(std::max<long>((long)lengths[b] - 1, 0), Tt);
    }
    packed = lengths && !logits_out && Mp > 0 && Mp * 20 <= M * 19;
    if (!packed) Mp = M;
    head_rows = Mp;
    prof.mark("^fwd", s);
    OK(image_projections(gc, lc, g, P, B));
    prof.mark("fwd.image_kv", s);
    OK(embx.ensure(e * M * E, &ws_bytes));
    OK(xe.ensure(e * M * 5 * H, &ws_bytes));
    OK(g5.ensure(4 * (size_t)B * 5 * H, &ws_bytes));
    OK(hs.ensure(e * (size_t)(Tt + 1) * B * H, &ws_bytes));
    OK(cs.ensure(4 * (size_t)(Tt + 1) * B * H, &ws_bytes));
    OK(gact.ensure(e * M * 4 * H, &ws_bytes));
    OK(gpre.ensure(4 * (size_t)B * 4 * H, &ws_bytes));
    OK(ph.ensure(e * M * H, &ws_bytes));
    OK(sgate.ensure(e * M * H, &ws_bytes));
    OK(sent.ensure(e * M * H, &ws_bytes));
    OK(phid.ensure(e * M * H, &ws_bytes));
    OK(hemb.ensure(e * M * H, &ws_bytes));
    OK(semb.ensure(e * M * H, &ws_bytes));
    OK(alpha.ensure(4 * M * (k + 1), &ws_bytes));
    OK(actx.ensure(e * M * H, &ws_bytes));
    OK(outv.ensure(e * M * H, &ws_bytes));
    OK(proj.ensure(e * M * E, &ws_bytes));
    OK(lse.ensure(4 * M, &ws_bytes));
    OK(acc.ensure(64, &ws_bytes));
    OK(dlogits.ensure(e * M * (size_t)Vp, &ws_bytes));
    const int* d_prow = nullptr;
    const int* d_pm = nullptr;
    const int* d_steps = nullptr;
    if (packed) {
      OK(row_steps.ensure(4 * (size_t)B, &ws_bytes));
      OK(row_prow.ensure(4 * (size_t)M, &ws_bytes));
      OK(row_pm.ensure(4 * (size_t)M, &ws_bytes));
      OK(hprev_c.ensure(e * M * H, &ws_bytes));
      OK(hcur_c.ensure(e * M * H, &ws_bytes));
      OK(actx_p.ensure(e * M * H, &ws_bytes));
      // the live steps of every sample come from the captions themselves (last non-padding target); `lengths` only sizes
      // the launches, a different total is reported (CAPDEC_ERR_LENGTHS) and can never index out of bounds
      CUDA_OK(cudaMemsetAsync(row_pm.p, 0, 4 * (size_t)Mp, s));
      OK(build_row_map(lc, cap, S, B, Tt, cfg.padding_idx, Mp, row_steps.as<int>(), row_prow.as<int>(), row_pm.as<int>(),
                       err_flag));
      d_prow = row_prow.as<int>(); d_pm = row_pm.as<int>(); d_steps = row_steps.as<int>();
    }

    // word embeddings of the source tokens caption[:, :-1] (model:84, 164), time-major
    embed_gather<T>(lc, emb.as<T>(), E, V, cap, S, B, Tt, cfg.padding_idx, embx.as<T>(), E, err_flag);
    bool persistent = false;
    if constexpr (std::is_same<T, bf16>::value) persistent = !stepwise && lstm_seq_supported(B, H) && w_hh_perm.p;
    if (persistent) {
      // Persistent recurrence: ONE activation holds the complete x-part of the gate / sentinel pre-activations,
      //   xe[t*B+b] = emb(word).[lstm_w(word) | fc_xs(word)] + (g[b].[lstm_w(global) | fc_xs(global)] + biases),
      // gate columns unit-major (col 4*unit + gate) so that a CTA's 16 units are one 128-byte line per row; the
      // per-image term is added by the GEMM epilogue (row-broadcast addend, row m -> m % B).
      OK(g5b.ensure(e * (size_t)B * 5 * H, &ws_bytes));
      OK(fc(gc, gT.p, H, w_g_p.as<T>(), 5 * H, b5p.as<float>(), g5b.p, 5 * H, B, 5 * H, H, ACT_NONE));
      prof.mark("fwd.embed_g5", s);
      GemmCall c;
      c.A = embx.p; c.lda = E; c.B = w_x_p.p; c.ldb = 5 * H; c.b_mn = true; c.C = xe.p; c.ldc = 5 * H; c.M = (int)M;
      c.N = 5 * H; c.K = E; c.epi.addend = g5b.p; c.epi.ld_add = 5 * H; c.epi.add_mod = B;
      OK(gemm<T>(gc, c));
    } else {
      // hoisted x-part of the gate GEMM and of p_word_emb (model:86-89): word-embedding columns for all T steps ...
      OK(fc(gc, embx.p, E, w_x_n.as<T>(), 5 * H, nullptr, xe.p, 5 * H, (int)M, 5 * H, E, ACT_NONE));
      // ... and the global-image columns (+ biases) once per image
      OK(fc(gc, gT.p, H, w_g_n.as<T>(), 5 * H, b5.as<float>(), g5.p, 5 * H, B, 5 * H, H, ACT_NONE, true));
    }
    // recurrent core (model:63, 87-88, 94-95)
    CUDA_OK(cudaMemsetAsync(hs.p, 0, e * (size_t)B * H, s));
    CUDA_OK(cudaMemsetAsync(cs.p, 0, 4 * (size_t)B * H, s));
    prof.mark("fwd.xe_gemm", s);
    if constexpr (std::is_same<T, bf16>::value) {
      if (persistent) {
        // all T steps in one persistent kernel: W_hh resident in shared memory, cell fused into the MMA epilogue
        OK(rec_ctr.ensure(256, &ws_bytes));
        OK(lstm_seq_fwd(lc, B, H, Tt, w_hh_perm.as<bf16>(), xe.as<bf16>(), 5 * H, hs.as<bf16>(), cs.as<float>(),
                        gact.as<bf16>(), rec_ctr.as<unsigned>()));
      }
    }
    fwd_persistent = persistent;
    for (int t = 0; t < Tt && !persistent; ++t) {
      OK(fc(gc, hs.as<T>() + (size_t)t * B * H, H, w_hh_n.as<T>(), 4 * H, nullptr, gpre.p, 4 * H, B, 4 * H, H, ACT_NONE,
              true));
      lstm_cell_fwd<T>(lc, B, H, gpre.as<float>(), 4 * H, xe.as<T>() + (size_t)t * B * 5 * H, 5 * H, g5.as<float>(),
                       5 * H, INT_MAX, 1, cs.as<float>() + (size_t)t * B * H, hs.as<T>() + (size_t)(t + 1) * B * H, H,
                       cs.as<float>() + (size_t)(t + 1) * B * H, gact.as<T>() + (size_t)t * B * 4 * H);
    }
    prof.mark("fwd.recurrence", s);
    // h_{t-1}, h_t of the head's rows: views of hs, or (packed rows) gathered copies of the live rows
    const T* Hprev = hs.as<T>();
    const T* Hcur = hs.as<T>() + (size_t)B * H;
    const float* Ccur = cs.as<float>() + (size_t)B * H;
    if (packed) {
      gather_rows<T>(lc, Hprev, Hcur, H, d_pm, Mp, H, hprev_c.as<T>(), hcur_c.as<T>(), H);
      Hprev = hprev_c.as<T>(); Hcur = hcur_c.as<T>();
    }
    const int Mh = (int)Mp;   // rows of the time-batched head
    {
      // the sentinel chain (h_{t-1} -> p_hidden -> sentinel -> sentinel_emb) beside the hidden chain (h_t -> p_hid -> hid_emb)
      cudaStream_t s2 = fork(s);
      GemmContext g2; g2.stream = s2;
      LaunchCtx l2{s2, &launches};
      OK(fc(g2, Hprev, H, w_hs_n.as<T>(), H, prm.fc_hs_b, ph.p, H, Mh, H, H, ACT_NONE));                         // model:90
      OK(fc(gc, Hcur, H, w_p_n.as<T>(), H, prm.fc_p_b, phid.p, H, Mh, H, H, ACT_TANH));                          // model:99
      sentinel_fwd<T>(l2, Mp, H, ph.as<T>(), H, xe.as<T>() + 4 * H, 5 * H,
                      persistent ? nullptr : g5.as<float>() + 4 * H, 5 * H, B, 1, Ccur, sgate.as<T>(), sent.as<T>(),
                      d_pm);                                                                          // model:91-92
      OK(fc(g2, sent.p, H, w_s_n.as<T>(), H, prm.fc_s_b, semb.p, H, Mh, H, H, ACT_NONE));                        // model:104
      OK(fc(gc, phid.p, H, w_g2_n.as<T>(), H, prm.fc_g_b, hemb.p, H, Mh, H, H, ACT_NONE));                       // model:102
      OK(join(s, s2));
      gc.launches += g2.launches;
    }
    prof.mark("fwd.head_fc", s);
    AttnShape as = train_shape(B);
    as.prow = d_prow; as.pm = d_pm; as.nq = d_steps; as.packed_rows = packed ? Mp : 0;
    // packed rows: the batched GEMMs of the attention (context forward, d alpha / d values backward) keep the padded row
    // index (rows of one image are B apart); the dead rows of their operand alpha_T stay zero from here to the backward pass
    OK(alpha_t.ensure(e * M * kAlphaTLd, &ws_bytes));
    if (packed) CUDA_OK(cudaMemsetAsync(alpha_t.p, 0, e * M * kAlphaTLd, s));
    bool attn_batched = false;
    if constexpr (std::is_same<T, bf16>::value) {
      if (attention_fwd_batched_supported(as)) {
        // scores + soft-max on MUFU / mma.sync, context as one batched tcgen05 GEMM (its output gathered into the packed
        // actx with packed rows)                                                                     // model:103-115
        OK(sadd.ensure(e * M * H, &ws_bytes));
        OK(attention_fwd_batched(gc, lc, as, keys.as<bf16>(), vals.as<bf16>(), hemb.as<bf16>(), semb.as<bf16>(),
                                 sent.as<bf16>(), phid.as<bf16>(), prm.fc_a_w, prm.fc_a_b, alpha.as<float>(),
                                 alpha_t.as<bf16>(), sadd.as<bf16>(), packed ? actx_p.as<bf16>() : actx.as<bf16>()));
        if (packed) gather_rows<T>(lc, actx_p.as<T>(), nullptr, H, d_pm, Mp, H, actx.as<T>(), nullptr, H);
        attn_batched = true;
      }
    }
    if (!attn_batched)
      attention_fwd<T>(lc, as, keys.as<T>(), vals.as<T>(), hemb.as<T>(), semb.as<T>(), sent.as<T>(), phid.as<T>(),
                       prm.fc_a_w, prm.fc_a_b, alpha.as<float>(), actx.as<T>(), nullptr);             // model:103-115
    prof.mark("fwd.attention", s);
    OK(fc(gc, actx.p, H, w_o_n.as<T>(), H, prm.fc_o_b, outv.p, H, Mh, H, H, ACT_TANH));                          // model:115
    OK(fc(gc, outv.p, H, w_e_n.as<T>(), E, prm.fc_e_b, proj.p, E, Mh, E, H, ACT_NONE));                          // model:24
    CUDA_OK(cudaMemsetAsync(acc.p, 0, 64, s));
    prof.mark("fwd.out_fc", s);
    OK(logits.ensure(4 * M * (size_t)Vp, &ws_bytes));
    {
      // Logits stay fp32 also on the bf16 path.  (Measured and dropped: bf16 logit storage -- 42 us faster, but soft-max
      // probabilities from rounded logits carry a relative error of |x| * 2^-9 and d out_fc_bias = sum_rows(p - onehot)
      // leaves the 2e-2 tolerance; and the cross entropy fused into two passes of this GEMM -- 289 + 202 us against
      // 161 + 118 us for GEMM + kernel: the K = 256 GEMM is epilogue-bound, a second pass costs more than the round trip.)
      GemmCall c;                                                                                     // model:25
      c.A = proj.p; c.lda = E; c.B = emb.p; c.ldb = E; c.C = logits.p; c.ldc = Vp; c.M = Mh; c.N = V; c.K = E;
      c.epi.bias = b_vocab.as<float>(); c.epi.aux_padded = 1; c.epi.out_f32 = 1;
      OK(gemm<T>(gc, c));
    }
    prof.mark("fwd.out_proj_vocab", s);
    // masked cross entropy (model:163-171, 180-183, 205-212) fused with the unscaled d logits = (softmax - onehot) * mask,
    // so the logits leave HBM only once
    ce_fused<T>(lc, logits.as<float>(), Vp, V, Vp, cap, S, B, Tt, cfg.padding_idx, lse.as<float>(), acc.as<float>(),
                dlogits.as<T>(), err_flag, Mp, d_pm);
    prof.mark("fwd.cross_entropy", s);
    loss_finalize(lc, acc.as<float>());
    if (logits_out) logits_to_batch_major(lc, logits.as<float>(), Vp, V, B, Tt, logits_out);          // model:129-130
    if (loss_out) CUDA_OK(cudaMemcpyAsync(loss_out, acc.p, 16, cudaMemcpyDeviceToDevice, s));
    prof.mark("fwd.loss_out", s);
    launches += gc.launches;
    CUDA_OK(cudaGetLastError());
    have_fwd = true;
    return 0;
  }

  long S_len() const { return S; }
  AttnShape train_shape(int B) const {
    AttnShape as;
    as.H = H; as.k = k; as.n_img = B; as.q_per_img = Tt; as.row0_stride = 1; as.q_stride = B;
    as.position_softmax = cfg.softmax_axis == CAPDEC_SOFTMAX_POSITION;
    return as;
  }

  // ------------------------------------------------------------------------------------------------------------
  int train_bwd(float loss_scale, const float* inv_ntok, float* dg, float* dP, const capdec_grads* gr,
                cudaStream_t s) override {
    if (!have_fwd) { set_error("capdec_train_bwd: no forward pass to differentiate"); return -1; }
    GemmContext gc; gc.stream = s;
    LaunchCtx lc{s, &launches};
    const size_t e = sizeof(T);
    const int B = last_B, X = E + H;
    const long M = (long)B * Tt;
    const long Mp = head_rows;         // rows of the time-batched head: M, or the live rows (packed, see train_fwd)
    const int Mh = (int)Mp;
    const int64_t* cap = last_cap;
    const capdec_grads& G = *gr;
    OK(dproj.ensure(e * M * E, &ws_bytes));
    OK(doutp.ensure(e * M * H, &ws_bytes));
    OK(dactx.ensure(e * M * H, &ws_bytes));
    OK(de.ensure(4 * M * (k + 1), &ws_bytes));
    OK(dsctx.ensure(e * M * H, &ws_bytes));
    OK(dkv.ensure(e * (size_t)B * k * 2 * H, &ws_bytes));
    OK(dhemb.ensure(e * M * H, &ws_bytes));
    OK(dsemb.ensure(e * M * H, &ws_bytes));
    OK(dphp.ensure(e * M * H, &ws_bytes));
    OK(dS.ensure(e * M * H, &ws_bytes));
    OK(dG.ensure(e * M * 5 * H, &ws_bytes));
    OK(dChead.ensure(4 * M * H, &ws_bytes));
    OK(dHhead.ensure(e * M * H, &ws_bytes));
    OK(dHprev.ensure(e * M * H, &ws_bytes));
    OK(dhrec.ensure(2 * 4 * (size_t)B * H, &ws_bytes));
    OK(dcrec.ensure(4 * (size_t)B * H, &ws_bytes));
    OK(dGsum.ensure(e * (size_t)B * 5 * H, &ws_bytes));
    OK(dembx.ensure(e * M * E, &ws_bytes));
    if (packed) {
      OK(dactx_p.ensure(e * M * H, &ws_bytes));
      OK(dsg_c.ensure(e * M * H, &ws_bytes));
      OK(dHhead_c.ensure(e * M * H, &ws_bytes));
      OK(dHprev_c.ensure(e * M * H, &ws_bytes));
    }
    const int* d_prow = packed ? row_prow.as<int>() : nullptr;
    const int* d_pm = packed ? row_pm.as<int>() : nullptr;
    const int* d_steps = packed ? row_steps.as<int>() : nullptr;

    prof.mark("^bwd", s);
    // gradients are overwritten: start from zero, every contribution below is an fp32 atomic add
    struct Z { float* p; size_t n; };
    const Z zs[] = {{G.fc_v_w, (size_t)H * H}, {G.fc_v_b, (size_t)H}, {G.fc_k_w, (size_t)H * H}, {G.fc_k_b, (size_t)H},
                    {G.lstm_w, (size_t)(X + H) * 4 * H}, {G.lstm_b, (size_t)4 * H}, {G.fc_xs_w, (size_t)X * H},
                    {G.fc_xs_b, (size_t)H}, {G.fc_hs_w, (size_t)H * H}, {G.fc_hs_b, (size_t)H},
                    {G.fc_p_w, (size_t)H * H}, {G.fc_p_b, (size_t)H}, {G.fc_g_w, (size_t)H * H}, {G.fc_g_b, (size_t)H},
                    {G.fc_s_w, (size_t)H * H}, {G.fc_s_b, (size_t)H}, {G.fc_a_w, (size_t)H}, {G.fc_a_b, (size_t)1},
                    {G.fc_o_w, (size_t)H * H}, {G.fc_o_b, (size_t)H}, {G.fc_e_w, (size_t)H * E}, {G.fc_e_b, (size_t)E},
                    {G.word_embedding, (size_t)V * E}, {G.out_fc_bias, (size_t)V}};
    ZeroTable zt;
    for (const Z& z : zs) {
      if (!z.p) { set_error("capdec_train_bwd: NULL gradient pointer"); return -1; }
      zt.add(z.p, (long)z.n);
    }
    multi_zero(lc, zt);
    prof.mark("bwd.zero_grads", s);
    // h_{t-1}, h_t, c_t of the head's rows (see train_fwd)
    const T* Hprev = packed ? hprev_c.as<T>() : hs.as<T>();
    const T* Hcur = packed ? hcur_c.as<T>() : hs.as<T>() + (size_t)B * H;
    const float* Ccur = cs.as<float>() + (size_t)B * H;

    // d logits = (softmax - onehot) * mask was written by the forward pass; the factor 1/sum(mask) * loss_scale
    // (softmax_with_cross_entropy_grad + the mask/scale ops) is applied by its three consumers
    float* gscale = acc.as<float>() + 4;
    scale_scalar(lc, inv_ntok ? inv_ntok : acc.as<float>() + 3, loss_scale, gscale);
    // Only what the BPTT kernel needs (the data-gradient chain down to d h, d c) and the weight gradients that are final
    // with it run before that kernel; the vocabulary-side and per-image weight gradients are independent of the recurrence
    // and run at the END of the pass, so that the gate-weight gradients -- which exist only after the BPTT -- are exchanged
    // behind them instead of behind nothing (data-parallel: the exposed tail was 334 us at 8 GPUs).
    // tied output projection (model:24-25): d proj, d fc_e
    {
      GemmCall c;
      c.A = dlogits.p; c.lda = Vp; c.B = emb.p; c.ldb = E; c.b_mn = true; c.C = dproj.p; c.ldc = E; c.M = Mh; c.N = E;
      c.K = V; c.epi.alpha_dev = gscale;
      OK(gemm<T>(gc, c));
    }
    ColsumTable bias_sums;  // the bias gradients of the time-batched head: one launch, just before their group is marked
    bias_sums.add(dproj.p, E, Mp, E, G.fc_e_b);
    prof.mark("bwd.vocab_proj", s);
    // out = tanh(fc_o(ctx + p_hid))   (model:115)
    {
      GemmCall c;
      c.A = dproj.p; c.lda = E; c.B = w_e_n.p; c.ldb = E; c.C = doutp.p; c.ldc = H; c.M = Mh; c.N = H; c.K = E;
      c.epi.act = ACT_DTANH; c.epi.dact_y = outv.p; c.epi.ld_y = H;
      OK(gemm<T>(gc, c));
    }
    bias_sums.add(doutp.p, H, Mp, H, G.fc_o_b);
    // The data-gradient chain (d proj -> d out -> d ctx -> attention -> .. -> d h) is what the BPTT kernel waits for; the
    // weight gradients of the head only have to be final when group 0 is marked: they run on the side stream beside it.
    GemmContext g2;
    cudaStream_t s2 = fork(s);
    g2.stream = s2;
    LaunchCtx l2{s2, &launches};
    OK(wgrad(g2, outv.p, H, H, dproj.p, E, E, Mh, G.fc_e_w, E));
    OK(wgrad(g2, actx.p, H, H, doutp.p, H, H, Mh, G.fc_o_w, H));
    {
      GemmCall c;  // d(ctx + p_hid)
      c.A = doutp.p; c.lda = H; c.B = w_o_n.p; c.ldb = H; c.C = dactx.p; c.ldc = H; c.M = Mh; c.N = H; c.K = H;
      OK(gemm<T>(gc, c));
    }
    prof.mark("bwd.out_fc", s);
    // attention (model:103-113)
    AttnShape as = train_shape(B);
    as.prow = d_prow; as.pm = d_pm; as.nq = d_steps; as.packed_rows = packed ? Mp : 0;
    OK(dal.ensure(4 * M * kDalLd, &ws_bytes));
    OK(alpha_t.ensure(e * M * kAlphaTLd, &ws_bytes));
    // packed rows: the two batched GEMMs of the attention gradient index rows by (image, step); they read d context in
    // the padded layout with zeros in the dead rows (alpha_T is zero there since the forward pass)
    const T* dctx_att = dactx.as<T>();
    if (packed) {
      scatter_rows_zero<T>(lc, dactx.as<T>(), nullptr, d_prow, M, H, dactx_p.as<T>(), nullptr);
      dctx_att = dactx_p.as<T>();
    }
    OK(attention_bwd<T>(gc, lc, as, keys.as<T>(), vals.as<T>(), hemb.as<T>(), semb.as<T>(), sent.as<T>(), prm.fc_a_w,
                        alpha.as<float>(), dctx_att, dal.as<float>(), alpha_t.as<T>(), de.as<float>(),
                        dsctx.as<T>(), dkv.as<T>(), dhemb.as<T>(), dsemb.as<T>(), G.fc_a_w, G.fc_a_b));
    prof.mark("bwd.attention", s);
    // two independent chains down to the BPTT kernel: hid_emb -> p_hid -> d h_t on the caller's stream,
    // sentinel_emb -> sentinel -> p_hidden -> d h_{t-1} on the side stream (behind the two weight gradients queued above)
    s2 = fork(s);
    g2.stream = s2;
    l2.stream = s2;
    bias_sums.add(dhemb.p, H, Mp, H, G.fc_g_b);
    bias_sums.add(dsemb.p, H, Mp, H, G.fc_s_b);
    {
      GemmCall c;  // d sentinel = alpha_k d ctx / (k+1) + d sentinel_emb . fc_s^T                (model:104, 111)
      c.A = dsemb.p; c.lda = H; c.B = w_s_n.p; c.ldb = H; c.C = dS.p; c.ldc = H; c.M = Mh; c.N = H; c.K = H;
      c.epi.addend = dsctx.p; c.epi.ld_add = H;
      OK(gemm<T>(g2, c));
    }
    {
      GemmCall c;  // d p_hid (pre-tanh) = (d actx + d hid_emb . fc_g^T) * (1 - p_hid^2)      (model:99, 102)
      c.A = dhemb.p; c.lda = H; c.B = w_g2_n.p; c.ldb = H; c.C = dphp.p; c.ldc = H; c.M = Mh; c.N = H; c.K = H;
      c.epi.addend = dactx.p; c.epi.ld_add = H; c.epi.act = ACT_DTANH; c.epi.dact_y = phid.p; c.epi.ld_y = H;
      OK(gemm<T>(gc, c));
    }
    bias_sums.add(dphp.p, H, Mp, H, G.fc_p_b);
    // sentinel = sigmoid(p_word_emb + p_hidden) * tanh(c)     (model:89-92); d(gate pre-activation) is one column block
    // of dG [M, 5H] (padded rows: the BPTT kernel and the x-side weight gradients read it), d c goes to dChead
    const T* dsg = dG.as<T>() + 4 * H;   // d(sentinel gate pre-activation) of the head's rows, and its pitch
    long ld_dsg = 5 * H;
    if (packed) {
      sentinel_bwd_packed<T>(l2, M, H, d_prow, dS.as<T>(), sgate.as<T>(), Ccur, dG.as<T>() + 4 * H, 5 * H,
                             dChead.as<float>(), dsg_c.as<T>());
      dsg = dsg_c.as<T>(); ld_dsg = H;
    } else {
      sentinel_bwd<T>(l2, M, H, dS.as<T>(), sgate.as<T>(), Ccur, dG.as<T>() + 4 * H, 5 * H, dChead.as<float>());
    }
    bias_sums.add(dsg, ld_dsg, Mp, H, G.fc_hs_b);
    {
      GemmCall c;  // d h_t from p_hid
      c.A = dphp.p; c.lda = H; c.B = w_p_n.p; c.ldb = H; c.C = packed ? dHhead_c.p : dHhead.p; c.ldc = H; c.M = Mh;
      c.N = H; c.K = H;
      OK(gemm<T>(gc, c));
    }
    {
      GemmCall c;  // d h_{t-1} from p_hidden
      c.A = dsg; c.lda = ld_dsg; c.B = w_hs_n.p; c.ldb = H; c.C = packed ? dHprev_c.p : dHprev.p; c.ldc = H; c.M = Mh;
      c.N = H; c.K = H;
      OK(gemm<T>(g2, c));
    }
    // the four weight gradients of these chains, two per stream, behind the data gradients
    OK(wgrad(gc, phid.p, H, H, dhemb.p, H, H, Mh, G.fc_g_w, H));
    OK(wgrad(g2, sent.p, H, H, dsemb.p, H, H, Mh, G.fc_s_w, H));
    OK(wgrad(gc, Hcur, H, H, dphp.p, H, H, Mh, G.fc_p_w, H));
    OK(wgrad(g2, Hprev, H, H, dsg, ld_dsg, H, Mh, G.fc_hs_w, H));
    OK(join(s, s2));
    gc.launches += g2.launches;
    g2.launches = 0;
    if (packed)   // the BPTT kernel reads both by (step, sample): back to padded rows, zeros where no target exists
      scatter_rows_zero<T>(lc, dHhead_c.as<T>(), dHprev_c.as<T>(), d_prow, M, H, dHhead.as<T>(), dHprev.as<T>());
    prof.mark("bwd.head_fc", s);
    colsum_multi<T>(lc, bias_sums);
    CUDA_OK(cudaMemcpyAsync(G.fc_xs_b, G.fc_hs_b, 4 * H, cudaMemcpyDeviceToDevice, s));  // same pre-activation
    prof.mark("bwd.head_bias", s);
    // fc_hs .. fc_e (incl. fc_a) are final: their all-reduce overlaps the BPTT
    mark_ready(0, s);
    // BPTT through the recurrent core (while_grad of model:87-88, 94-95)
    bool persistent = false;
    if constexpr (std::is_same<T, bf16>::value) {
      if (fwd_persistent) {  // same schedule as the forward pass (the gate-activation layouts differ)
        OK(rec_ctr.ensure(256, &ws_bytes));
        OK(lstm_seq_bwd(lc, B, H, Tt, w_hh_n.as<bf16>(), dHhead.as<bf16>(), dHprev.as<bf16>(), dChead.as<float>(),
                        gact.as<bf16>(), cs.as<float>(), dG.as<bf16>(), 5 * H, dhrec.as<float>(), dcrec.as<float>(),
                        rec_ctr.as<unsigned>()));
        persistent = true;
      }
    }
    if (!persistent) CUDA_OK(cudaMemsetAsync(dcrec.p, 0, 4 * (size_t)B * H, s));
    for (int t = Tt - 1; t >= 0 && !persistent; --t) {
      const bool last = (t == Tt - 1);
      lstm_cell_bwd<T>(lc, B, H, dHhead.as<T>() + (size_t)t * B * H,
                       last ? nullptr : dHprev.as<T>() + (size_t)(t + 1) * B * H, last ? nullptr : dhrec.as<float>(),
                       dChead.as<float>() + (size_t)t * B * H, dcrec.as<float>(), gact.as<T>() + (size_t)t * B * 4 * H,
                       cs.as<float>() + (size_t)(t + 1) * B * H, cs.as<float>() + (size_t)t * B * H,
                       dG.as<T>() + (size_t)t * B * 5 * H, 5 * H);
      if (t > 0) {
        GemmCall c;  // d h_{t-1} through the gates
        c.A = dG.as<T>() + (size_t)t * B * 5 * H; c.lda = 5 * H; c.B = w_hh_n.p; c.ldb = 4 * H; c.C = dhrec.p;
        c.ldc = H; c.M = B; c.N = H; c.K = 4 * H; c.epi.out_f32 = 1; c.epi.atomic = 1; c.splits = 0;  // K = 4H: split-K
        CUDA_OK(cudaMemsetAsync(dhrec.p, 0, 4 * (size_t)B * H, s));
        OK(gemm<T>(gc, c));
      }
    }
    prof.mark("bwd.recurrence", s);
    // After the BPTT everything left is independent; it is ordered so that the gradients become final in large groups
    // early and a small one last (capdec_set_grad_ready_events): each group's all-reduce runs behind the remaining work.
    // Rows without a target have d gates == 0 exactly (nothing flows into them), so with packed rows the contractions over
    // (step, sample) run on the live rows only: d gates and the word embeddings are gathered once, h_{t-1} already is.
    const T* HprevW = hs.as<T>();       // h_{t-1} of the rows the weight gradients contract over
    const T* dGw = dG.as<T>();          // d [gates | sentinel gate] of the same rows, pitch 5H
    const T* embW = embx.as<T>();
    const int Mw = packed ? Mh : (int)M;
    // (the per-image terms need only the sum over steps of d gates: they run on the side stream beside the large
    // hidden-state gradient)
    s2 = fork(s);
    g2.stream = s2;
    l2.stream = s2;
    if (packed) {
      OK(dG_c.ensure(e * Mp * 5 * H, &ws_bytes));
      OK(embx_c.ensure(e * Mp * E, &ws_bytes));
      gather_rows<T>(lc, dG.as<T>(), nullptr, 5 * H, d_pm, Mp, 5 * H, dG_c.as<T>(), nullptr, 5 * H);
      gather_rows<T>(lc, embx.as<T>(), nullptr, E, d_pm, Mp, E, embx_c.as<T>(), nullptr, E);
      HprevW = hprev_c.as<T>(); dGw = dG_c.as<T>(); embW = embx_c.as<T>();
    }
    // weight gradients of the gate GEMM, batched over all steps
    OK(wgrad(gc, HprevW, H, H, dGw, 5 * H, 4 * H, Mw, G.lstm_w + (size_t)X * 4 * H, 4 * H));
    mark_ready(2, s);  // lstm_w rows [E+H, E+2H) (hidden-state rows) are final
    sum_over_steps<T>(l2, dG.as<T>(), 5 * H, B, Tt, 5 * H, dGsum.as<T>(), 5 * H, d_steps);
    OK(wgrad(g2, gT.p, H, H, dGsum.p, 5 * H, 4 * H, B, G.lstm_w + (size_t)E * 4 * H, 4 * H));
    if (dg) {
      GemmCall c;  // gradient handed back to the encoder side: d global_img_feat
      c.A = dGsum.p; c.lda = 5 * H; c.B = w_g_n.p; c.ldb = 5 * H; c.C = dg; c.ldc = H; c.M = B; c.N = H; c.K = 5 * H;
      c.epi.out_f32 = 1;
      OK(gemm<T>(g2, c));
    }
    if (fwd_staged) {  // d(pre-ReLU img_global) for capdec_img2feature_bwd (model:199)
      OK(dzg.ensure(e * (size_t)B * H, &ws_bytes));
      GemmCall c;
      c.A = dGsum.p; c.lda = 5 * H; c.B = w_g_n.p; c.ldb = 5 * H; c.C = dzg.p; c.ldc = H; c.M = B; c.N = H; c.K = 5 * H;
      c.epi.act = ACT_DRELU; c.epi.dact_y = gT.p; c.epi.ld_y = H;
      OK(gemm<T>(g2, c));
      have_dz = true;
    }
    OK(wgrad(g2, gT.p, H, H, dGsum.as<T>() + 4 * H, 5 * H, H, B, G.fc_xs_w + (size_t)E * H, H));
    OK(join(s, s2));
    mark_ready(3, s);  // lstm_w rows [E, E+H) (global-image rows) are final
    // the vocabulary side of the tied table (deferred from the start of the pass; 49 tiles: a third of the SMs idle) beside
    // the word-row gradients
    s2 = fork(s);
    g2.stream = s2;
    l2.stream = s2;
    OK(wgrad(g2, dlogits.p, Vp, V, proj.p, E, E, Mh, G.word_embedding, E, gscale));
    colsum_add<T>(l2, dlogits.as<T>(), Vp, Mp, V, G.out_fc_bias, gscale);
    OK(wgrad(gc, embW, E, E, dGw + 4 * H, 5 * H, H, Mw, G.fc_xs_w, H));
    colsum_add<T>(lc, dGw, 5 * H, Mw, 4 * H, G.lstm_b);
    OK(wgrad(gc, embW, E, E, dGw, 5 * H, 4 * H, Mw, G.lstm_w, 4 * H));
    mark_ready(4, s);  // lstm_b, fc_xs_w, fc_xs_b and the lstm_w word rows are final
    prof.mark("bwd.gate_wgrad", s);
    // the tied table, lookup side (needs d gates): d word_emb -> scatter-add, padding_idx skipped
    {
      GemmCall c;
      c.A = dGw; c.lda = 5 * H; c.B = w_x_n.p; c.ldb = 5 * H; c.C = dembx.p; c.ldc = E; c.M = Mw; c.N = E;
      c.K = 5 * H;
      OK(gemm<T>(gc, c));
    }
    OK(join(s, s2));   // the scatter-add and the vocabulary-side product add into the same table: one after the other
    embed_scatter_add<T>(lc, dembx.as<T>(), E, E, V, cap, S, B, Tt, cfg.padding_idx, G.word_embedding, d_pm, Mw);
    mark_ready(1, s);  // word_embedding, out_fc_bias are final
    prof.mark("bwd.embedding", s);
    ColsumTable kv_sums;
    // once-per-image projections (model:52-53)
    kv_sums.add(dkv.p, 2 * H, (long)B * k, H, G.fc_v_b);
    kv_sums.add(dkv.as<T>() + H, 2 * H, (long)B * k, H, G.fc_k_b);
    s2 = fork(s);
    g2.stream = s2;
    l2.stream = s2;
    OK(wgrad(gc, PT.p, H, H, dkv.p, 2 * H, H, B * k, G.fc_v_w, H));
    OK(wgrad(g2, PT.p, H, H, dkv.as<T>() + H, 2 * H, H, B * k, G.fc_k_w, H));
    colsum_multi<T>(l2, kv_sums);
    if (dP) {
      GemmCall c;  // d p_img_feat
      c.A = dkv.p; c.lda = 2 * H; c.B = w_kv_n.p; c.ldb = 2 * H; c.C = dP; c.ldc = H; c.M = B * k; c.N = H; c.K = 2 * H;
      c.epi.out_f32 = 1;
      OK(gemm<T>(gc, c));
    }
    if (fwd_staged) {
      // the features came from the adaptor (capdec_img2feature): leave d(pre-ReLU img_embed) = d p_img_feat * [P > 0]
      // in the handle for capdec_img2feature_bwd (model:197)
      OK(dzp.ensure(e * (size_t)B * k * H, &ws_bytes));
      GemmCall c;
      c.A = dkv.p; c.lda = 2 * H; c.B = w_kv_n.p; c.ldb = 2 * H; c.C = dzp.p; c.ldc = H; c.M = B * k; c.N = H; c.K = 2 * H;
      c.epi.act = ACT_DRELU; c.epi.dact_y = PT.p; c.epi.ld_y = H;
      OK(gemm<T>(gc, c));
    }
    OK(join(s, s2));
    mark_ready(5, s);  // fc_v, fc_k: everything is final
    prof.mark("bwd.image_kv", s);
    launches += g2.launches;
    launches += gc.launches;
    CUDA_OK(cudaGetLastError());
    return 0;
  }

  // ------------------------------------------------------------------------------------------------------------
  int decode(const float* g, const float* P, int B, int beam, int finish, int32_t* ids, float* scores, float* ids_f32,
             cudaStream_t s) override {
    if (!bound) { set_error("capdec_decode: parameters not bound"); return -1; }
    if (B <= 0 || beam < 1 || beam > 8) { set_error("capdec_decode: need B > 0 and 1 <= beam <= 8"); return -1; }
    if (beam > V) { set_error("capdec_decode: beam wider than the vocabulary"); return -1; }
    GemmContext gc; gc.stream = s;
    LaunchCtx lc{s, &launches};
    const size_t e = sizeof(T);
    const int R = B * beam, X = E + H;
    have_fwd = false;  // the activation buffers of a pending train_fwd are reused below
    prof.mark("^dec", s);
    OK(image_projections(gc, lc, g, P, B));
    prof.mark("dec.image_kv", s);
    OK(g5.ensure(4 * (size_t)B * 5 * H, &ws_bytes));
    OK(xin.ensure(e * (size_t)R * X, &ws_bytes));
    OK(gpre.ensure(4 * (size_t)R * 5 * H, &ws_bytes));
    OK(d_hnew.ensure(e * (size_t)R * H, &ws_bytes));
    OK(d_cnew.ensure(4 * (size_t)R * H, &ws_bytes));
    OK(d_ccur.ensure(4 * (size_t)R * H, &ws_bytes));
    OK(sent.ensure(e * (size_t)R * H, &ws_bytes));
    OK(phid.ensure(e * (size_t)R * H, &ws_bytes));
    OK(hemb.ensure(e * (size_t)R * H, &ws_bytes));
    OK(semb.ensure(e * (size_t)R * H, &ws_bytes));
    OK(alpha.ensure(4 * (size_t)R * (k + 1), &ws_bytes));
    OK(actx.ensure(e * (size_t)R * H, &ws_bytes));
    OK(outv.ensure(e * (size_t)R * H, &ws_bytes));
    OK(proj.ensure(e * (size_t)R * E, &ws_bytes));
    // (Measured and dropped: soft-max statistics + top-k candidates emitted by the vocabulary GEMM's epilogue instead of the
    // fp32 logits -- 144 us per step against 62 (GEMM) + 38 (top-k kernel): the K = 256 GEMM is epilogue-bound and selection
    // logic in its register-capped epilogue costs more than the 76 MB round trip.)
    OK(logits.ensure(4 * (size_t)R * Vp, &ws_bytes));
    OK(d_state.ensure((size_t)R * (2 * 4 + 2 * 4 + 2 * 4 * Ti + 8) + 256, &ws_bytes));
    OK(d_beam.ensure(beam_scratch_bytes(B, beam), &ws_bytes));
    OK(d_trace.ensure(8 * (size_t)R * Ti, &ws_bytes));
    trace_B = B; trace_beam = beam;
    // beam bookkeeping (double-buffered)
    float* score[2]; int* fin[2]; int* hist[2]; int *tok, *src;
    {
      char* p = d_state.as<char>();
      score[0] = (float*)p; p += 4 * R; score[1] = (float*)p; p += 4 * R;
      fin[0] = (int*)p; p += 4 * R; fin[1] = (int*)p; p += 4 * R;
      hist[0] = (int*)p; p += 4 * (size_t)R * Ti; hist[1] = (int*)p; p += 4 * (size_t)R * Ti;
      tok = (int*)p; p += 4 * R; src = (int*)p;
    }
    AttnShape as;
    as.H = H; as.k = k; as.n_img = B; as.q_per_img = beam; as.row0_stride = beam; as.q_stride = 1;
    as.position_softmax = cfg.softmax_axis == CAPDEC_SOFTMAX_POSITION;
    // everything from here to the end of the step loop touches handle-owned memory only
    cudaStream_t rs = s;  // stream the step loop is enqueued on (a private stream while it is being captured)
    auto run_steps = [&]() -> int {
      CUDA_OK(cudaMemsetAsync(d_state.p, 0, d_state.cap, rs));
      // only prefix 0 of every image is live at step 0; every row starts from <start>, h = c = 0 (model:56-58, 63)
      beam_init(lc, R, beam, cfg.start_idx, score[0], tok, src);
      CUDA_OK(cudaMemsetAsync(d_hnew.p, 0, e * (size_t)R * H, rs));
      CUDA_OK(cudaMemsetAsync(d_cnew.p, 0, 4 * (size_t)R * H, rs));
      beam_reorder<T>(lc, B, beam, E, H, cfg.padding_idx, emb.as<T>(), tok, src, d_hnew.as<T>(), d_cnew.as<float>(),
                      xin.as<T>(), X, d_ccur.as<float>());
      // global-image half of the gate / sentinel pre-activations, once per image
      OK(fc(gc, gT.p, H, w_g_n.as<T>(), 5 * H, b5d.as<float>(), g5.p, 5 * H, B, 5 * H, H, ACT_NONE, true));
      prof.mark("dec.setup", rs);
      for (int step = 0; step < Ti; ++step) {
        const int cur = step & 1, nxt = cur ^ 1;
        // [word_emb | h_{t-1}] . [lstm_w(word) ; lstm_w(hid) | fc_xs(word) ; fc_hs]: gate and sentinel-gate pre-activations
        // in one fp32 product, then cell + sentinel in one kernel                                        (model:86-92)
        OK(fc(gc, xin.p, X, w_dec_n.as<T>(), 5 * H, nullptr, gpre.p, 5 * H, R, 5 * H, X, ACT_NONE, true));
        lstm_cell_sentinel_fwd<T>(lc, R, H, gpre.as<float>(), g5.as<float>(), beam, d_ccur.as<float>(), d_hnew.as<T>(),
                                  d_cnew.as<float>(), sent.as<T>());
        prof.mark("dec.gates_cell", rs);
        // sentinel_emb = fc(sentinel) (model:104) does not depend on p_hid / hid_emb (model:99-102): it runs on the side
        // stream beside them (M = rows is small: each of these GEMMs leaves a third of the SMs idle and is latency-bound)
        cudaStream_t s2 = fork(rs);
        GemmContext gs; gs.stream = s2;
        OK(fc(gs, sent.p, H, w_s_n.as<T>(), H, prm.fc_s_b, semb.p, H, R, H, H, ACT_NONE));
        gc.launches += gs.launches;
        OK(fc(gc, d_hnew.p, H, w_p_n.as<T>(), H, prm.fc_p_b, phid.p, H, R, H, H, ACT_TANH));
        OK(fc(gc, phid.p, H, w_g2_n.as<T>(), H, prm.fc_g_b, hemb.p, H, R, H, H, ACT_NONE));
        OK(join(rs, s2));
        prof.mark("dec.head_fc", rs);
        bool attn_small = false;
        if constexpr (std::is_same<T, bf16>::value) {
          if (attention_fwd_small_supported(as)) {
            OK(attention_fwd_small(lc, as, keys.as<bf16>(), vals.as<bf16>(), hemb.as<bf16>(), semb.as<bf16>(),
                                   sent.as<bf16>(), phid.as<bf16>(), prm.fc_a_w, prm.fc_a_b, alpha.as<float>(),
                                   actx.as<bf16>()));
            attn_small = true;
          }
        }
        if (!attn_small)
          attention_fwd<T>(lc, as, keys.as<T>(), vals.as<T>(), hemb.as<T>(), semb.as<T>(), sent.as<T>(), phid.as<T>(),
                           prm.fc_a_w, prm.fc_a_b, alpha.as<float>(), actx.as<T>(), nullptr);
        prof.mark("dec.attention", rs);
        OK(fc(gc, actx.p, H, w_o_n.as<T>(), H, prm.fc_o_b, outv.p, H, R, H, H, ACT_TANH));
        OK(fc(gc, outv.p, H, w_e_n.as<T>(), E, prm.fc_e_b, proj.p, E, R, E, H, ACT_NONE));
        {
          GemmCall c;                                                                                 // model:25
          c.A = proj.p; c.lda = E; c.B = emb.p; c.ldb = E; c.C = logits.p; c.ldc = Vp; c.M = R; c.N = V; c.K = E;
          c.epi.bias = b_vocab.as<float>(); c.epi.aux_padded = 1; c.epi.out_f32 = 1;
          OK(gemm<T>(gc, c));
        }
        prof.mark("dec.out_proj_vocab", rs);
        // top-k + back-pointers (model:119-123 generalised), then reorder the recurrent state for the next step
        beam_step(lc, logits.as<float>(), Vp, V, B, beam, step, Ti, cfg.padding_idx, cfg.stop_idx, finish, score[cur],
                  score[nxt], fin[cur], fin[nxt], hist[cur], hist[nxt], tok, src, d_beam.p,
                  d_trace.as<int>() + (size_t)step * R * 2);
        prof.mark("dec.beam_topk", rs);
        if (step + 1 < Ti)
          beam_reorder<T>(lc, B, beam, E, H, cfg.padding_idx, emb.as<T>(), tok, src, d_hnew.as<T>(), d_cnew.as<float>(),
                          xin.as<T>(), X, d_ccur.as<float>());
        prof.mark("dec.reorder", rs);
      }
      return 0;
    };
    // Small batches are launch-latency bound (14 dependent launches per step, ~9 us each at one image): the step loop
    // is captured once into a CUDA graph per (images, beam, finish) and replayed.  The first call of a shape runs
    // eagerly (it allocates the workspace and warms the tensor-map cache), the second captures.
#ifndef CAPDEC_DECODE_GRAPH_MAX_ROWS
#define CAPDEC_DECODE_GRAPH_MAX_ROWS 96   // build-time: larger decodes keep programmatic dependent launch instead
#endif
    const bool want_graph = R <= CAPDEC_DECODE_GRAPH_MAX_ROWS && !prof.on;
    bool done = false;
    if (want_graph) {
      const unsigned long key_epoch = g_ws_epoch;
      if (dgraph.exec && dgraph.B == B && dgraph.beam == beam && dgraph.finish == finish && dgraph.epoch == key_epoch &&
          dgraph.stream == s) {
        CUDA_OK(cudaGraphLaunch(dgraph.exec, s));
        launches += dgraph.launches;
        done = true;
      } else if (dgraph.seen_B == B && dgraph.seen_beam == beam && dgraph.seen_finish == finish &&
                 dgraph.seen_epoch == key_epoch) {
        dgraph.reset();
        const long l0 = launches + gc.launches;
        set_gemm_pdl(false);  // programmatic edges are not captured: plain launches inside the graph
        cudaGraph_t graph = nullptr;
        // captured on a private stream (the caller's may be the legacy default stream, which cannot capture); the graph
        // itself is launched into the caller's stream
        if (!cap_stream && cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking) != cudaSuccess) cap_stream = nullptr;
        if (cap_stream && cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          rs = cap_stream; gc.stream = cap_stream; lc.stream = cap_stream;
          const int rc = run_steps();
          const cudaError_t ce = cudaStreamEndCapture(cap_stream, &graph);
          rs = s; gc.stream = s; lc.stream = s;
          if (rc == 0 && ce == cudaSuccess && graph &&
              cudaGraphInstantiate(&dgraph.exec, graph, 0) == cudaSuccess) {
            dgraph.B = B; dgraph.beam = beam; dgraph.finish = finish; dgraph.epoch = key_epoch; dgraph.stream = s;
            dgraph.launches = (launches + gc.launches) - l0;
          } else {
            dgraph.exec = nullptr;
            cudaGetLastError();
          }
          if (graph) cudaGraphDestroy(graph);
        }
        cudaGetLastError();  // a refused capture must not leave its error behind for the eager path
        set_gemm_pdl(true);
        launches = l0 - gc.launches;  // the capture enqueued nothing; the replay below is what runs
        if (dgraph.exec) {
          CUDA_OK(cudaGraphLaunch(dgraph.exec, s));
          launches += dgraph.launches;
          done = true;
        }
      }
      dgraph.seen_B = B; dgraph.seen_beam = beam; dgraph.seen_finish = finish; dgraph.seen_epoch = g_ws_epoch;
    }
    if (!done) OK(run_steps());
    const int fin_buf = Ti & 1;
    if (ids) CUDA_OK(cudaMemcpyAsync(ids, hist[fin_buf], 4 * (size_t)R * Ti, cudaMemcpyDeviceToDevice, s));
    if (scores) CUDA_OK(cudaMemcpyAsync(scores, score[fin_buf], 4 * (size_t)R, cudaMemcpyDeviceToDevice, s));
    if (ids_f32) ids_to_f32(lc, hist[fin_buf], B, beam, Ti, ids_f32);
    launches += gc.launches;
    CUDA_OK(cudaGetLastError());
    return 0;
  }

  // (token, back-pointer) of every selection of the last decode: [T, B, beam, 2] int32
  int decode_trace(int32_t* out, cudaStream_t s) override {
    if (trace_B <= 0 || !d_trace.p) { set_error("capdec_decode_trace: no decode has run on this handle"); return -1; }
    CUDA_OK(cudaMemcpyAsync(out, d_trace.p, 8 * (size_t)trace_B * trace_beam * Ti, cudaMemcpyDeviceToDevice, s));
    return 0;
  }

  static void ids_to_f32(LaunchCtx lc, const int* hist, int B, int beam, int Ti, float* out);

 private:
  std::vector<DevBuf*> all_bufs() {
    return {&w_kv_n, &w_x_n, &w_g_n, &w_hh_n, &w_hs_n, &w_p_n, &w_g2_n, &w_s_n, &w_o_n, &w_e_n, &emb, &w_dec_n, &b5, &b5d, &b_vocab, &gT, &PT,
            &keys, &vals, &embx, &xe, &g5, &hs, &cs, &gact, &gpre, &ph, &sgate, &sent, &phid, &hemb, &semb, &alpha, &actx,
            &outv, &proj, &logits, &lse, &acc, &dlogits, &dproj, &doutp, &dactx, &de, &dsctx, &dkv, &dhemb, &dsemb, &dphp,
            &dS, &dG, &dChead, &dHhead, &dHprev, &dhrec, &dcrec, &dGsum, &dembx, &dal, &alpha_t, &w_hh_perm, &rec_ctr, &w_x_p, &w_g_p, &b5p, &g5b, &sadd, &xin, &d_hnew, &d_cnew, &d_ccur,
            &d_state, &d_beam, &d_trace, &row_steps, &row_prow, &row_pm, &hprev_c, &hcur_c, &actx_p, &dactx_p, &dsg_c,
            &dHhead_c, &dHprev_c, &dG_c, &embx_c, &w_img_n, &w_glob_n, &imgT, &meanT, &dzp, &dzg, &dimg, &dmean};
  }

  // captured step loop of capdec_decode for small batches
  struct DecodeGraph {
    cudaGraphExec_t exec = nullptr;
    int B = 0, beam = 0, finish = 0;
    unsigned long epoch = 0;
    cudaStream_t stream = nullptr;
    long launches = 0;
    int seen_B = -1, seen_beam = -1, seen_finish = -1;
    unsigned long seen_epoch = 0;
    void reset() {
      if (exec) cudaGraphExecDestroy(exec);
      exec = nullptr;
    }
  } dgraph;
  cudaStream_t cap_stream = nullptr;
  cudaStream_t side_stream = nullptr;                 // decode: independent GEMMs of a step run beside each other
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  capdec_config cfg;
  capdec_params prm{};
  int H, E, V, Vp, k, S, Tt, Ti;
  bool bound = false, have_fwd = false, stepwise = false, fwd_persistent = false;
  int last_B = 0;
  const int64_t* last_cap = nullptr;
  // operands
  DevBuf w_kv_n, w_x_n, w_g_n, w_hh_n, w_hs_n, w_p_n, w_g2_n, w_s_n, w_o_n, w_e_n, emb, w_dec_n, b5, b5d, b_vocab;
  // activations kept for the backward pass
  DevBuf gT, PT, keys, vals, embx, xe, g5, hs, cs, gact, gpre, ph, sgate, sent, phid, hemb, semb, alpha, actx, outv, proj,
      logits, lse, acc;
  // backward scratch
  DevBuf dlogits, dproj, doutp, dactx, de, dsctx, dkv, dhemb, dsemb, dphp, dS, dG, dChead, dHhead, dHprev, dhrec, dcrec,
      dGsum, dembx, dal, alpha_t, w_hh_perm, rec_ctr, w_x_p, w_g_p, b5p, g5b, sadd;
  // packed rows of the time-batched head (train_fwd with lengths)
  bool packed = false;
  long head_rows = 0;
  DevBuf row_steps, row_prow, row_pm, hprev_c, hcur_c, actx_p, dactx_p, dsg_c, dHhead_c, dHprev_c, dG_c, embx_c;
  // decode state
  DevBuf xin, d_hnew, d_cnew, d_ccur, d_state, d_beam, d_trace;
  int trace_B = 0, trace_beam = 0;
  // feature adaptor (capdec_img2feature)
  capdec_adaptor_params aprm{};
  int Xc = 0, staged_B = 0;
  bool adaptor_bound = false, staged = false, fwd_staged = false, have_dz = false;
  DevBuf w_img_n, w_glob_n, imgT, meanT, dzp, dzg, dimg, dmean;
};

namespace {
__global__ void ids_to_f32_kernel(const int* __restrict__ hist, int B, int beam, int Ti, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * Ti) return;
  const int b = i / Ti, t = i % Ti;
  out[i] = (float)hist[((long)b * beam) * Ti + t];  // best beam, cast like model:122
}
}  // namespace
template <typename T> void Decoder<T>::ids_to_f32(LaunchCtx lc, const int* hist, int B, int beam, int Ti, float* out) {
  ids_to_f32_kernel<<<(B * Ti + 255) / 256, 256, 0, lc.stream>>>(hist, B, beam, Ti, out);
  if (lc.launches) ++*lc.launches;
}

}  // namespace capdec

// ================================================================================================================
// C-ABI
// ================================================================================================================
using namespace capdec;

struct capdec_handle_s {
  DecoderBase* impl;
};

static int check_device(capdec_handle h) {
  if (!h || !h->impl) { set_error("null capdec handle"); return -1; }
  cudaError_t e = cudaSetDevice(h->impl->device);
  if (e != cudaSuccess) { set_error("cudaSetDevice(%d): %s", h->impl->device, cudaGetErrorString(e)); return -1; }
  return 0;
}

// input errors found by kernels of an EARLIER call on this handle (see DecoderBase::err_flag)
static int report_async_error(capdec_handle h) {
  const int e = h->impl->take_async_error();
  if (!e) return 0;
  set_error("an earlier call on this handle was given bad input:%s%s%s (offending token ids were treated as padding; with"
            " wrong lengths the step's results are undefined)",
            (e & CAPDEC_ERR_SOURCE_ID) ? " caption token id outside [0, vocab_size)" : "",
            (e & CAPDEC_ERR_TARGET_ID) ? " target token id outside [0, vocab_size)" : "",
            (e & CAPDEC_ERR_LENGTHS) ? " lengths[] does not describe the captions (sum of lengths[b] - 1 differs from the"
                                       " number of decode steps up to the last non-padding target)" : "");
  return -1;
}

namespace capdec { void set_gemm_trace(unsigned long long* buf); void set_rec_trace(unsigned long long* buf); }

extern "C" {

// debug hook (not part of include/capdec.h): device buffer of 148*32 u64 receiving per-CTA globaltimer stamps
// debug hook: device buffer of n_ctas * T * 8 u64 receiving the per-step timeline of the next persistent recurrence
// kernel (forward and backward write the same layout; trace one direction at a time)
void capdec_debug_rec_trace(void* buf) { capdec::set_rec_trace(reinterpret_cast<unsigned long long*>(buf)); }
void capdec_debug_gemm_trace(void* buf) { capdec::set_gemm_trace(reinterpret_cast<unsigned long long*>(buf)); }

const char* capdec_last_error(void) { return g_err; }
int capdec_version(void) { return 100; }

int capdec_create(const capdec_config* c, int device, capdec_handle* out) {
  if (!c || !out) { set_error("capdec_create: null argument"); return -1; }
  if (c->hidden_dim <= 0 || c->hidden_dim % 8 || c->embedding_size <= 0 || c->embedding_size % 8) {
    set_error("capdec_create: hidden_dim and embedding_size must be positive multiples of 8"); return -1;
  }
  if (c->num_positions < 1 || c->num_positions > 63) { set_error("capdec_create: num_positions must be in [1,63]"); return -1; }
  if (c->vocab_size < 4 || c->sentence_length < 2 || c->infer_max_length < 1) {
    set_error("capdec_create: bad vocab_size / sentence_length / infer_max_length"); return -1;
  }
  if (c->vocab_size > beam_max_vocab()) {   // the reference caps its dictionary at 15000 words (ai_challenge_tokenizer.py:21-50)
    set_error("capdec_create: vocab_size %d exceeds the %d tokens the top-k kernels hold in registers", c->vocab_size,
              beam_max_vocab());
    return -1;
  }
  if (c->padding_idx < 0 || c->padding_idx >= c->vocab_size || c->start_idx < 0 || c->start_idx >= c->vocab_size ||
      c->stop_idx < 0 || c->stop_idx >= c->vocab_size) {
    set_error("capdec_create: special token ids out of range"); return -1;
  }
  if (c->softmax_axis != CAPDEC_SOFTMAX_FEATURE && c->softmax_axis != CAPDEC_SOFTMAX_POSITION) {
    set_error("capdec_create: unsupported softmax_axis %d", c->softmax_axis); return -1;
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev <= 0) {
    set_error("capdec_create: no CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e)); return -1;
  }
  if (device < 0 || device >= ndev) { set_error("capdec_create: device %d out of range", device); return -1; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return -1; }
  if (prop.major != 10) {
    set_error("capdec_create: device %d is sm_%d%d; the kernels are built for sm_100a (B200) only", device, prop.major,
              prop.minor);
    return -1;
  }
  e = cudaSetDevice(device);
  if (e != cudaSuccess) { set_error("cudaSetDevice: %s", cudaGetErrorString(e)); return -1; }
  DecoderBase* impl = nullptr;
  if (c->dtype == CAPDEC_DTYPE_F32) impl = new (std::nothrow) Decoder<float>(*c);
  else if (c->dtype == CAPDEC_DTYPE_BF16) impl = new (std::nothrow) Decoder<bf16>(*c);
  else { set_error("capdec_create: unsupported dtype %d", c->dtype); return -1; }
  if (!impl) { set_error("capdec_create: out of host memory"); return -1; }
  impl->device = device;
  void* flag = nullptr;
  e = cudaHostAlloc(&flag, 64, cudaHostAllocMapped);
  if (e != cudaSuccess) { delete impl; set_error("capdec_create: cudaHostAlloc: %s", cudaGetErrorString(e)); return -1; }
  memset(flag, 0, 64);
  impl->err_flag = static_cast<int*>(flag);   // unified addressing: the same pointer is valid on the device
  capdec_handle h = new (std::nothrow) capdec_handle_s{impl};
  if (!h) { cudaFreeHost(flag); delete impl; set_error("capdec_create: out of host memory"); return -1; }
  *out = h;
  return 0;
}

int capdec_destroy(capdec_handle h) {
  if (!h) return 0;
  if (h->impl) {
    cudaSetDevice(h->impl->device);
    int* flag = h->impl->err_flag;
    delete h->impl;
    if (flag) { cudaDeviceSynchronize(); cudaFreeHost(flag); }
  }
  delete h;
  return 0;
}

int capdec_set_grad_ready_events(capdec_handle h, void* const* events, int32_t n) {
  if (!h || !h->impl) { set_error("null capdec handle"); return -1; }
  if (n < 0 || n > 6 || (n > 0 && !events)) { set_error("capdec_set_grad_ready_events: need 0 <= n <= 6 events"); return -1; }
  for (int i = 0; i < 6; ++i) h->impl->grad_ready[i] = (i < n) ? (cudaEvent_t)events[i] : nullptr;
  return 0;
}

int capdec_profile(capdec_handle h, int32_t enable) {
  if (!h || !h->impl) { set_error("null capdec handle"); return -1; }
  h->impl->prof.on = enable != 0;
  if (!enable) h->impl->prof.report();
  return 0;
}
int capdec_profile_report(capdec_handle h, char* buf, size_t n) {
  if (check_device(h)) return -1;
  if (!buf || n == 0) { set_error("capdec_profile_report: null buffer"); return -1; }
  const std::string r = h->impl->prof.report();
  snprintf(buf, n, "%s", r.c_str());
  return 0;
}

size_t capdec_workspace_bytes(capdec_handle h) { return (h && h->impl) ? h->impl->ws_bytes : 0; }
int64_t capdec_launch_count(capdec_handle h) { return (h && h->impl) ? h->impl->launches : 0; }

int capdec_bind_params(capdec_handle h, const capdec_params* p, void* stream) {
  if (check_device(h)) return -1;
  if (!p) { set_error("capdec_bind_params: null params"); return -1; }
  const float* const* pp = reinterpret_cast<const float* const*>(p);
  for (size_t i = 0; i < sizeof(capdec_params) / sizeof(float*); ++i)
    if (!pp[i]) { set_error("capdec_bind_params: parameter pointer %zu is NULL", i); return -1; }
  return h->impl->bind(p, (cudaStream_t)stream);
}

int capdec_train_fwd(capdec_handle h, const float* g, const float* P, const int64_t* caption, const int32_t* lengths,
                     int32_t B, float* logits_out, float* loss_out, void* stream) {
  if (check_device(h) || report_async_error(h)) return -1;
  if (!caption) { set_error("capdec_train_fwd: null caption"); return -1; }
  return h->impl->train_fwd(g, P, caption, lengths, B, logits_out, loss_out, (cudaStream_t)stream);
}

int capdec_train_bwd(capdec_handle h, float loss_scale, const float* inv_ntok_dev, float* dg, float* dP,
                     const capdec_grads* gr, void* stream) {
  if (check_device(h) || report_async_error(h)) return -1;
  if (!gr) { set_error("capdec_train_bwd: null grads"); return -1; }
  return h->impl->train_bwd(loss_scale, inv_ntok_dev, dg, dP, gr, (cudaStream_t)stream);
}

int capdec_decode(capdec_handle h, const float* g, const float* P, int32_t B, int32_t beam, int32_t finish_on_stop,
                  int32_t* ids_out, float* scores_out, float* ids_f32_out, void* stream) {
  if (check_device(h)) return -1;
  return h->impl->decode(g, P, B, beam, finish_on_stop, ids_out, scores_out, ids_f32_out, (cudaStream_t)stream);
}

int capdec_status(capdec_handle h, void* stream) {
  if (check_device(h)) return -1;
  cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
  if (e != cudaSuccess) { set_error("capdec_status: %s", cudaGetErrorString(e)); return -1; }
  return report_async_error(h);
}

int capdec_decode_trace(capdec_handle h, int32_t* trace_out, void* stream) {
  if (check_device(h)) return -1;
  if (!trace_out) { set_error("capdec_decode_trace: null output"); return -1; }
  return h->impl->decode_trace(trace_out, (cudaStream_t)stream);
}

int capdec_bind_adaptor(capdec_handle h, int32_t encoder_channel, const capdec_adaptor_params* p, void* stream) {
  if (check_device(h)) return -1;
  if (!p || !p->fc_img_w || !p->fc_img_b || !p->fc_glob_w || !p->fc_glob_b) {
    set_error("capdec_bind_adaptor: null parameter pointer"); return -1;
  }
  return h->impl->bind_adaptor(encoder_channel, p, (cudaStream_t)stream);
}

int capdec_img2feature(capdec_handle h, const float* enc_map, int32_t B, float* p_img_feat_out,
                       float* global_img_feat_out, void* stream) {
  if (check_device(h)) return -1;
  if (!enc_map) { set_error("capdec_img2feature: null encoder map"); return -1; }
  return h->impl->img2feature(enc_map, B, p_img_feat_out, global_img_feat_out, (cudaStream_t)stream);
}

int capdec_img2feature_bwd(capdec_handle h, float* d_enc_map, const capdec_adaptor_grads* gr, void* stream) {
  if (check_device(h)) return -1;
  if (!gr || !gr->fc_img_w || !gr->fc_img_b || !gr->fc_glob_w || !gr->fc_glob_b) {
    set_error("capdec_img2feature_bwd: null gradient pointer"); return -1;
  }
  return h->impl->img2feature_bwd(d_enc_map, gr, (cudaStream_t)stream);
}

int capdec_adam_step(float* param, const float* grad, float* moment1, float* moment2, int64_t n, float lr, float beta1,
                     float beta2, float epsilon, float beta1_pow, float beta2_pow, float clip_value, void* stream) {
  if (n < 0 || (n > 0 && (!param || !grad || !moment1 || !moment2))) { set_error("capdec_adam_step: null buffer"); return -1; }
  if (!(beta1_pow < 1.f) || !(beta2_pow < 1.f)) { set_error("capdec_adam_step: beta powers must be < 1 (step >= 1)"); return -1; }
  if (adam_step((cudaStream_t)stream, param, grad, moment1, moment2, (long)n, lr, beta1, beta2, epsilon, beta1_pow,
                beta2_pow, clip_value)) {
    set_error("capdec_adam_step: launch failed: %s", cudaGetErrorString(cudaGetLastError())); return -1;
  }
  return 0;
}

int capdec_cast_bf16(int32_t to_bf16, const void* src, void* dst, int64_t n, void* stream) {
  if (n < 0 || (n > 0 && (!src || !dst))) { set_error("capdec_cast_bf16: null buffer"); return -1; }
  if (n % 8 || ((uintptr_t)src & 15) || ((uintptr_t)dst & 15)) {
    set_error("capdec_cast_bf16: n must be a multiple of 8 and the buffers 16-byte aligned"); return -1;
  }
  LaunchCtx lc{(cudaStream_t)stream, nullptr};
  if (to_bf16) convert_from_f32<bf16>(lc, static_cast<const float*>(src), static_cast<bf16*>(dst), (long)n);
  else convert_to_f32<bf16>(lc, static_cast<const bf16*>(src), static_cast<float*>(dst), (long)n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("capdec_cast_bf16: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

int capdec_gemm(int32_t dtype, int32_t majors, const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                int64_t ldc, int32_t M, int32_t N, int32_t K, const float* bias, const void* addend, int64_t ld_add,
                const void* dact_y, int64_t ld_y, int32_t act, int32_t out_f32, int32_t atomic, int32_t variant,
                void* stream) {
  if (variant < GEMM_AUTO || variant > GEMM_SINGLE) { set_error("capdec_gemm: unknown kernel variant %d", variant); return -1; }
  GemmContext gc; gc.stream = (cudaStream_t)stream; gc.variant = variant;
  GemmCall c;
  c.A = A; c.lda = lda; c.B = B; c.ldb = ldb; c.C = C; c.ldc = ldc; c.M = M; c.N = N; c.K = K;
  c.a_mn = (majors & 1) != 0; c.b_mn = (majors & 2) != 0;
  c.epi.bias = bias; c.epi.addend = addend; c.epi.ld_add = ld_add; c.epi.dact_y = dact_y; c.epi.ld_y = ld_y;
  c.epi.act = act; c.epi.out_f32 = out_f32; c.epi.atomic = atomic; c.splits = atomic ? 0 : 1;
  if (dtype == CAPDEC_DTYPE_F32) return gemm<float>(gc, c);
  if (dtype == CAPDEC_DTYPE_BF16) return gemm<bf16>(gc, c);
  set_error("capdec_gemm: unsupported dtype %d", dtype);
  return -1;
}

static AttnShape make_shape(int H, int k, int n_img, int q, int64_t r0, int64_t qs, int axis) {
  AttnShape s;
  s.H = H; s.k = k; s.n_img = n_img; s.q_per_img = q; s.row0_stride = r0; s.q_stride = qs;
  s.position_softmax = axis == CAPDEC_SOFTMAX_POSITION;
  return s;
}

int capdec_attention_fwd(int32_t dtype, int32_t H, int32_t k, int32_t n_img, int32_t q_per_img, int64_t row0_stride,
                         int64_t q_stride, int32_t softmax_axis, const void* keys, const void* vals,
                         const void* hid_emb, const void* sent_emb, const void* sentinel, const void* p_hid,
                         const float* w_alpha, const float* b_alpha, float* alpha, void* actx, void* ctx, void* stream) {
  if (H % 8 || k < 1 || k > 63) { set_error("capdec_attention_fwd: need H %% 8 == 0 and 1 <= k <= 63"); return -1; }
  LaunchCtx lc{(cudaStream_t)stream, nullptr};
  AttnShape s = make_shape(H, k, n_img, q_per_img, row0_stride, q_stride, softmax_axis);
  if (dtype == CAPDEC_DTYPE_BF16 && !attention_fwd_batched_supported(s) && attention_fwd_small_supported(s) &&
      (actx || ctx)) {
    // the inference-shaped path (scores kernel + small context kernel)
    int rc = 0;
    if (actx)
      rc = attention_fwd_small(lc, s, (const bf16*)keys, (const bf16*)vals, (const bf16*)hid_emb, (const bf16*)sent_emb,
                               (const bf16*)sentinel, (const bf16*)p_hid, w_alpha, b_alpha, alpha, (bf16*)actx);
    if (!rc && ctx)
      rc = attention_fwd_small(lc, s, (const bf16*)keys, (const bf16*)vals, (const bf16*)hid_emb, (const bf16*)sent_emb,
                               (const bf16*)sentinel, nullptr, w_alpha, b_alpha, alpha, (bf16*)ctx);
    return rc ? -1 : 0;
  }
  if (dtype == CAPDEC_DTYPE_BF16 && attention_fwd_batched_supported(s) && (actx || ctx)) {
    // the training-shaped path (scores kernel + batched context GEMM); its scratch is owned by the decoder handle in
    // the train schedule, the stand-alone entry point allocates it
    GemmContext gc; gc.stream = (cudaStream_t)stream;
    const long rows = (long)n_img * q_per_img;
    void *at = nullptr, *sa = nullptr;
    if (cudaMalloc(&at, rows * kAlphaTLd * 2) != cudaSuccess || cudaMalloc(&sa, rows * (size_t)H * 2) != cudaSuccess) {
      if (at) cudaFree(at);
      set_error("capdec_attention_fwd: out of device memory");
      return -1;
    }
    int rc = 0;
    if (actx)
      rc = attention_fwd_batched(gc, lc, s, (const bf16*)keys, (const bf16*)vals, (const bf16*)hid_emb,
                                 (const bf16*)sent_emb, (const bf16*)sentinel, (const bf16*)p_hid, w_alpha, b_alpha, alpha,
                                 (bf16*)at, (bf16*)sa, (bf16*)actx);
    if (!rc && ctx)
      rc = attention_fwd_batched(gc, lc, s, (const bf16*)keys, (const bf16*)vals, (const bf16*)hid_emb,
                                 (const bf16*)sent_emb, (const bf16*)sentinel, nullptr, w_alpha, b_alpha, alpha, (bf16*)at,
                                 (bf16*)sa, (bf16*)ctx);
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(at);
    cudaFree(sa);
    return rc ? -1 : 0;
  }
  if (dtype == CAPDEC_DTYPE_F32)
    attention_fwd<float>(lc, s, (const float*)keys, (const float*)vals, (const float*)hid_emb, (const float*)sent_emb,
                         (const float*)sentinel, (const float*)p_hid, w_alpha, b_alpha, alpha, (float*)actx, (float*)ctx);
  else if (dtype == CAPDEC_DTYPE_BF16)
    attention_fwd<bf16>(lc, s, (const bf16*)keys, (const bf16*)vals, (const bf16*)hid_emb, (const bf16*)sent_emb,
                        (const bf16*)sentinel, (const bf16*)p_hid, w_alpha, b_alpha, alpha, (bf16*)actx, (bf16*)ctx);
  else { set_error("capdec_attention_fwd: unsupported dtype %d", dtype); return -1; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("capdec_attention_fwd: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

int capdec_attention_bwd(int32_t dtype, int32_t H, int32_t k, int32_t n_img, int32_t q_per_img, int64_t row0_stride,
                         int64_t q_stride, int32_t softmax_axis, const void* keys, const void* vals,
                         const void* hid_emb, const void* sent_emb, const void* sentinel, const float* w_alpha,
                         const float* alpha, const void* dctx, float* de, void* dS_ctx, void* dKV, void* dhid_emb,
                         void* dsent_emb, float* dw_alpha, float* db_alpha, void* stream) {
  if (H % 8 || k < 1 || k > 63) { set_error("capdec_attention_bwd: need H %% 8 == 0 and 1 <= k <= 63"); return -1; }
  LaunchCtx lc{(cudaStream_t)stream, nullptr};
  GemmContext gc; gc.stream = (cudaStream_t)stream;
  AttnShape s = make_shape(H, k, n_img, q_per_img, row0_stride, q_stride, softmax_axis);
  // scratch of the two batched GEMMs (the decoder handle owns these; the stand-alone entry point allocates them)
  const long rows = (long)n_img * q_per_img;
  const size_t esz = dtype == CAPDEC_DTYPE_F32 ? 4 : 2;
  void *dal = nullptr, *alpha_t = nullptr;
  if (cudaMalloc(&dal, rows * kDalLd * 4) != cudaSuccess || cudaMalloc(&alpha_t, rows * kAlphaTLd * esz) != cudaSuccess) {
    if (dal) cudaFree(dal);
    set_error("capdec_attention_bwd: out of device memory");
    return -1;
  }
  cudaMemsetAsync(alpha_t, 0, rows * kAlphaTLd * esz, (cudaStream_t)stream);
  int rc = 0;
  if (dtype == CAPDEC_DTYPE_F32)
    rc = attention_bwd<float>(gc, lc, s, (const float*)keys, (const float*)vals, (const float*)hid_emb,
                              (const float*)sent_emb, (const float*)sentinel, w_alpha, alpha, (const float*)dctx,
                              (float*)dal, (float*)alpha_t, de, (float*)dS_ctx, (float*)dKV, (float*)dhid_emb,
                              (float*)dsent_emb, dw_alpha, db_alpha);
  else if (dtype == CAPDEC_DTYPE_BF16)
    rc = attention_bwd<bf16>(gc, lc, s, (const bf16*)keys, (const bf16*)vals, (const bf16*)hid_emb, (const bf16*)sent_emb,
                             (const bf16*)sentinel, w_alpha, alpha, (const bf16*)dctx, (float*)dal, (bf16*)alpha_t, de,
                             (bf16*)dS_ctx, (bf16*)dKV, (bf16*)dhid_emb, (bf16*)dsent_emb, dw_alpha, db_alpha);
  else { set_error("capdec_attention_bwd: unsupported dtype %d", dtype); rc = -1; }
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(dal);
  cudaFree(alpha_t);
  if (rc) return -1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("capdec_attention_bwd: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

int capdec_lstm_cell_fwd(int32_t dtype, int32_t R, int32_t H, const float* gates_pre, const float* c_prev, void* h_out,
                         float* c_out, void* stream) {
  if (H % 8) { set_error("capdec_lstm_cell_fwd: H must be a multiple of 8"); return -1; }
  LaunchCtx lc{(cudaStream_t)stream, nullptr};
  if (dtype == CAPDEC_DTYPE_F32)
    lstm_cell_fwd<float>(lc, R, H, gates_pre, 4 * H, nullptr, 0, nullptr, 0, INT_MAX, 1, c_prev, (float*)h_out, H, c_out,
                         nullptr);
  else if (dtype == CAPDEC_DTYPE_BF16)
    lstm_cell_fwd<bf16>(lc, R, H, gates_pre, 4 * H, nullptr, 0, nullptr, 0, INT_MAX, 1, c_prev, (bf16*)h_out, H, c_out,
                        nullptr);
  else { set_error("capdec_lstm_cell_fwd: unsupported dtype %d", dtype); return -1; }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("capdec_lstm_cell_fwd: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

int capdec_beam_step(const float* logits, int64_t ldl, int32_t V, int32_t n_img, int32_t beam, int32_t step,
                     int32_t T_total, int32_t pad, int32_t stop, int32_t finish_on_stop, const float* score_in,
                     float* score_out, const int32_t* fin_in, int32_t* fin_out, const int32_t* hist_in,
                     int32_t* hist_out, int32_t* tok_out, int32_t* src_out, void* stream) {
  if (beam < 1 || beam > 8) { set_error("capdec_beam_step: 1 <= beam <= 8"); return -1; }
  if (V < 1 || V > beam_max_vocab()) { set_error("capdec_beam_step: 1 <= V <= %d", beam_max_vocab()); return -1; }
  if (n_img <= 0) return 0;
  LaunchCtx lc{(cudaStream_t)stream, nullptr};
  // the candidate scratch is owned by the decoder handle inside capdec_decode; the stand-alone entry point allocates it
  void* scratch = nullptr;
  if (cudaMalloc(&scratch, beam_scratch_bytes(n_img, beam)) != cudaSuccess) {
    set_error("capdec_beam_step: out of device memory");
    return -1;
  }
  beam_step(lc, logits, ldl, V, n_img, beam, step, T_total, pad, stop, finish_on_stop, score_in, score_out, fin_in,
            fin_out, hist_in, hist_out, tok_out, src_out, scratch, nullptr);
  cudaError_t e = cudaGetLastError();
  cudaStreamSynchronize((cudaStream_t)stream);
  cudaFree(scratch);
  if (e != cudaSuccess) { set_error("capdec_beam_step: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

}  // extern "C"
