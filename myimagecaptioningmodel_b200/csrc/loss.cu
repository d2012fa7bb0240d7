// Vocabulary-axis kernels: masked soft-max cross-entropy forward / backward (model:163-171, 180-183, 205-212),
// greedy arg-max and beam-search top-k with back-pointers (model:119-123; beam spec: SURVEY.md 8c row C5).
// One CTA per row / per image, warp-shuffle + shared-memory block reductions, 128-bit loads.
#include <cstdlib>

#include "kernels.cuh"

namespace capdec {

namespace {

constexpr int kCeThreads = 256;

__device__ __forceinline__ int64_t target_of(const int64_t* caption, long cap_ld, int B, long m) {
  const long t = m / B, b = m % B;
  return caption[b * cap_ld + t + 1];
}

// one pass for the online (max, sum), one L2-resident pass for the gradient: logits are read from HBM once.
// A target id outside [0, V) is reported (CAPDEC_ERR_TARGET_ID) and treated as padding.
template <typename T>
__global__ void __launch_bounds__(kCeThreads)
ce_fused_kernel(const float* __restrict__ logits, long ldl, int V, int Vp, const int64_t* __restrict__ caption,
                long cap_ld, int B, int pad, float* __restrict__ lse, float* __restrict__ acc, T* __restrict__ dlogits,
                int* __restrict__ err, const int* __restrict__ rowmap) {
  pdl_trigger();
  __shared__ float scratch[32];
  const long m = blockIdx.x;   // row of logits / lse / dlogits; with packed rows its (t, b) is rowmap[m] = t*B + b
  const float* row = logits + m * ldl;
  T* out = dlogits + m * (long)Vp;
  int64_t tgt = target_of(caption, cap_ld, B, rowmap ? rowmap[m] : m);
  if ((unsigned long)tgt >= (unsigned long)V) {
    if (threadIdx.x == 0 && err) atomicOr(err, 2 /* CAPDEC_ERR_TARGET_ID */);
    tgt = pad;
  }
  const int V8 = V & ~7;
  float mx = -INFINITY, sum = 0.f;
  for (int n = threadIdx.x * 8; n < V8; n += kCeThreads * 8) {
    float v[8];
    load8(row + n, v);
    float lm = v[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) lm = fmaxf(lm, v[i]);
    if (lm > mx) { sum *= __expf(mx - lm); mx = lm; }
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += __expf(v[i] - mx);
  }
  for (int n = V8 + threadIdx.x; n < V; n += kCeThreads) {
    const float x = to_f(row[n]);
    if (x > mx) { sum *= __expf(mx - x); mx = x; }
    sum += __expf(x - mx);
  }
  const float gmx = block_max(mx, scratch);
  const float gsum = block_sum(mx > -INFINITY ? sum * __expf(mx - gmx) : 0.f, scratch);
  const float l = gmx + logf(gsum);
  if (threadIdx.x == 0) {
    lse[m] = l;
    if (tgt != pad) {
      atomicAdd(acc + 0, l - to_f(row[tgt]));
      atomicAdd(acc + 1, 1.f);
    }
  }
  if (tgt == pad) {  // masked position: zero gradient row
    for (int n = threadIdx.x * 8; n < Vp; n += kCeThreads * 8) {
      float z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      store8(out + n, z);
    }
    return;
  }
  for (int n = threadIdx.x * 8; n < Vp; n += kCeThreads * 8) {
    float v[8];
    if (n + 8 <= V) {
      load8(row + n, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __expf(v[i] - l) - ((n + i) == tgt ? 1.f : 0.f);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = (n + i < V) ? __expf(to_f(row[n + i]) - l) - ((n + i) == tgt ? 1.f : 0.f) : 0.f;
    }
    store8(out + n, v);
  }
}

__global__ void scale_scalar_kernel(const float* in, float s, float* out) { out[0] = in[0] * s; }

__global__ void loss_finalize_kernel(float* acc) {
  acc[2] = acc[0] / acc[1];
  acc[3] = 1.f / acc[1];
}

__global__ void logits_bm_kernel(const float* __restrict__ logits, long ldl, int V, int B, int steps,
                                 float* __restrict__ out) {
  const long m = blockIdx.x;  // time-major row
  const long t = m / B, b = m % B;
  const float* src = logits + m * ldl;
  float* dst = out + (b * steps + t) * (long)V;
  for (int n = threadIdx.x; n < V; n += blockDim.x) dst[n] = src[n];
}

// ---- beam search step: one CTA per image --------------------------------------------------------------------
struct Cand { float v; int idx; };
__device__ __forceinline__ bool better(float v, int i, float bv, int bi) { return v > bv || (v == bv && i < bi); }
__device__ __forceinline__ Cand warp_best(Cand c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, c.v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, c.idx, o);
    if (better(ov, oi, c.v, c.idx)) { c.v = ov; c.idx = oi; }
  }
  return c;
}

constexpr int kMaxBeam = 8;

// ---- beam search step: one CTA per (image, prefix) ROW, then one warp per image -------------------------------------
// Within one prefix the candidate score  score_p + (x - lse_p)  is monotone in the raw logit x, so the per-prefix
// top-`beam` by x contains every global winner.
// beam_rows_kernel: the row's log-sum-exp and its `beam` best tokens (ties -> lower token), all logits of the row in
// registers (EPT per thread, every load in flight at once), one block barrier per selection round (every warp
// re-reduces the 8 warp winners from a double-buffered shared array).  A finished prefix offers exactly one candidate
// (itself, emitting pad, score unchanged).
// beam_merge_kernel: merges the beam*beam scored candidates of an image (ties -> lower prefix, then lower token) and
// writes tokens, back-pointers, scores, finished flags and histories.
constexpr int kRowThreads = 256;
constexpr int kRowWarps = kRowThreads / 32;

template <int EPT>
__global__ void __launch_bounds__(kRowThreads)
beam_rows_kernel(const float* __restrict__ logits, long ldl, int V, int beam, int pad, int finish_on_stop,
                 const float* __restrict__ score_in, const int* __restrict__ fin_in, float* __restrict__ cand_v,
                 int* __restrict__ cand_i, float* __restrict__ gscore) {
  pdl_trigger();
  __shared__ float scratch[32];
  __shared__ float s_v[2][kRowWarps];
  __shared__ int s_i[2][kRowWarps];
  const int r = blockIdx.x, p = r % beam;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int INT_BIG = 0x7fffffff;
  const bool greedy = (beam == 1 && !finish_on_stop);
  float* cv = cand_v + (long)r * kMaxBeam;
  int* ci = cand_i + (long)r * kMaxBeam;
  const float sc = score_in[r];
  if (!(sc > -INFINITY)) {  // dead prefix: no candidates
    if (tid < beam) { cv[tid] = -INFINITY; ci[tid] = INT_BIG; }
    return;
  }
  if (finish_on_stop && fin_in[r]) {  // finished: one candidate, itself + pad, score unchanged
    if (tid < beam) { cv[tid] = tid == 0 ? sc : -INFINITY; ci[tid] = tid == 0 ? p * V + pad : INT_BIG; }
    return;
  }
  const float* row = logits + (long)r * ldl;
  float x[EPT];
#pragma unroll
  for (int i = 0; i < EPT; ++i) {
    const int n = tid + i * kRowThreads;
    x[i] = (n < V) ? row[n] : -INFINITY;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < EPT; ++i) mx = fmaxf(mx, x[i]);
  const float gmx = block_max(mx, scratch);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < EPT; ++i) sum += __expf(x[i] - gmx);   // exp(-inf) = 0 for the padding slots
  const float lse = gmx + logf(block_sum(sum, scratch));
  for (int rr = 0; rr < beam; ++rr) {
    // local arg-max over the registers (ties -> lower token = lower i)
    float lv = -INFINITY;
    int li = INT_BIG;
#pragma unroll
    for (int i = 0; i < EPT; ++i)
      if (x[i] > lv) { lv = x[i]; li = tid + i * kRowThreads; }
    const Cand wb = warp_best(Cand{lv, li});
    if (lane == 0) { s_v[rr & 1][warp] = wb.v; s_i[rr & 1][warp] = wb.idx; }
    __syncthreads();
    Cand c{lane < kRowWarps ? s_v[rr & 1][lane] : -INFINITY, lane < kRowWarps ? s_i[rr & 1][lane] : INT_BIG};
    c = warp_best(c);  // every warp computes the same block winner
    if (tid == 0) {
      // greedy ranks the raw logits exactly like layers.argmax (model:120); the score is still accumulated
      cv[rr] = (c.idx == INT_BIG) ? -INFINITY : (greedy ? c.v : sc + (c.v - lse));
      ci[rr] = (c.idx == INT_BIG) ? INT_BIG : p * V + c.idx;
      if (greedy) gscore[r] = sc + (c.v - lse);
    }
    if (c.idx != INT_BIG && (c.idx % kRowThreads) == tid) {
      const int slot = c.idx / kRowThreads;
#pragma unroll
      for (int i = 0; i < EPT; ++i)
        if (i == slot) x[i] = -INFINITY;                  // remove the winner from this thread's pool
    }
  }
}

__global__ void __launch_bounds__(128)
beam_merge_kernel(int n_img, int V, int beam, int step, int T_total, int pad, int stop, int finish_on_stop,
                  const float* __restrict__ cand_v, const int* __restrict__ cand_i, const float* __restrict__ gscore,
                  float* __restrict__ score_out, const int* __restrict__ fin_in, int* __restrict__ fin_out,
                  const int* __restrict__ hist_in, int* __restrict__ hist_out, int* __restrict__ tok_out,
                  int* __restrict__ src_out, int* __restrict__ trace) {
  pdl_trigger();
  const int img = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (img >= n_img) return;
  const int INT_BIG = 0x7fffffff;
  const bool greedy = (beam == 1 && !finish_on_stop);
  // candidate (prefix p, rank rr) sits in lanes p * beam + rr (beam <= 8: two candidates per lane cover 64)
  float v0 = -INFINITY, v1 = -INFINITY;
  int i0 = INT_BIG, i1 = INT_BIG;
  {
    const int c0 = lane, c1 = lane + 32;
    if (c0 < beam * beam) {
      const long o = ((long)img * beam + c0 / beam) * kMaxBeam + c0 % beam;
      v0 = cand_v[o]; i0 = cand_i[o];
    }
    if (c1 < beam * beam) {
      const long o = ((long)img * beam + c1 / beam) * kMaxBeam + c1 % beam;
      v1 = cand_v[o]; i1 = cand_i[o];
    }
  }
  for (int r = 0; r < beam; ++r) {
    Cand best{-INFINITY, INT_BIG};
    if (i0 != INT_BIG && better(v0, i0, best.v, best.idx)) { best.v = v0; best.idx = i0; }
    if (i1 != INT_BIG && better(v1, i1, best.v, best.idx)) { best.v = v1; best.idx = i1; }
    best = warp_best(best);
    if (i0 == best.idx) i0 = INT_BIG;   // flat indices p*V + token are unique: remove the winner
    if (i1 == best.idx) i1 = INT_BIG;
    int idx = best.idx;
    float v = best.v;
    int src, tok;
    if (idx == INT_BIG) { src = 0; tok = pad; v = -INFINITY; }  // fewer live candidates than beams
    else { src = idx / V; tok = idx % V; }
    if (greedy && idx != INT_BIG) v = gscore[img * beam + src];
    const int was_fin = finish_on_stop ? fin_in[img * beam + src] : 0;
    if (lane == 0) {
      tok_out[img * beam + r] = tok;
      src_out[img * beam + r] = src;
      score_out[img * beam + r] = v;
      fin_out[img * beam + r] = (was_fin || (finish_on_stop && tok == stop)) ? 1 : 0;
      if (trace) {  // per-step record (token, back-pointer) of this selection
        trace[2 * (img * beam + r)] = tok;
        trace[2 * (img * beam + r) + 1] = src;
      }
    }
    const int* hin = hist_in + ((long)img * beam + src) * T_total;
    int* hout = hist_out + ((long)img * beam + r) * T_total;
    for (int t = lane; t < T_total; t += 32) hout[t] = (t == step) ? tok : (t < step ? hin[t] : 0);
  }
}

// initial beam state: only prefix 0 of every image is live, every row starts from <start> (model:56-58)
__global__ void beam_init_kernel(int R, int beam, int start, float* __restrict__ score, int* __restrict__ tok,
                                 int* __restrict__ src) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  score[r] = (r % beam == 0) ? 0.f : -INFINITY;
  tok[r] = start;
  src[r] = r % beam;
}

template <typename T>
__global__ void beam_reorder_kernel(int n_rows, int beam, int E, int H, int pad, const T* __restrict__ emb,
                                    const int* __restrict__ tok, const int* __restrict__ src,
                                    const T* __restrict__ h_new, const float* __restrict__ c_new, T* __restrict__ xin,
                                    long ld_xin, float* __restrict__ c_next) {
  pdl_trigger();
  const int ve = E / 8, vh = H / 8, per_row = ve + vh + vh;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = (long)n_rows * per_row;
  for (; i < n; i += stride) {
    const int r = (int)(i / per_row);
    const int v = (int)(i % per_row);
    const int srow = (r / beam) * beam + src[r];
    float x[8];
    if (v < ve) {
      const int id = tok[r];
      if (id == pad) {
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = 0.f;
      } else {
        load8(emb + (long)id * E + v * 8, x);
      }
      store8(xin + (long)r * ld_xin + v * 8, x);
    } else if (v < ve + vh) {
      const int h0 = (v - ve) * 8;
      load8(h_new + (long)srow * H + h0, x);
      store8(xin + (long)r * ld_xin + E + h0, x);
    } else {
      const int h0 = (v - ve - vh) * 8;
      load8(c_new + (long)srow * H + h0, x);
      store8(c_next + (long)r * H + h0, x);
    }
  }
}

}  // namespace

template <typename T>
void ce_fused(LaunchCtx c, const float* logits, long ldl, int V, int Vp, const int64_t* caption, long cap_ld, int B,
              int steps, int pad, float* lse, float* acc, T* dlogits, int* err, long rows, const int* rowmap) {
  if (rows < 0) rows = (long)B * steps;
  if (rows == 0) return;
  ce_fused_kernel<T><<<(unsigned)rows, kCeThreads, 0, c.stream>>>(logits, ldl, V, Vp, caption, cap_ld, B, pad, lse, acc,
                                                                 dlogits, err, rowmap);
  if (c.launches) ++*c.launches;
}
template void ce_fused<float>(LaunchCtx, const float*, long, int, int, const int64_t*, long, int, int, int, float*, float*,
                              float*, int*, long, const int*);
template void ce_fused<bf16>(LaunchCtx, const float*, long, int, int, const int64_t*, long, int, int, int, float*, float*,
                             bf16*, int*, long, const int*);
void scale_scalar(LaunchCtx c, const float* in, float s, float* out) {
  scale_scalar_kernel<<<1, 1, 0, c.stream>>>(in, s, out);
  if (c.launches) ++*c.launches;
}
void loss_finalize(LaunchCtx c, float* acc) {
  loss_finalize_kernel<<<1, 1, 0, c.stream>>>(acc);
  if (c.launches) ++*c.launches;
}
void logits_to_batch_major(LaunchCtx c, const float* logits, long ldl, int V, int B, int steps, float* out) {
  logits_bm_kernel<<<B * steps, 256, 0, c.stream>>>(logits, ldl, V, B, steps, out);
  if (c.launches) ++*c.launches;
}
size_t beam_scratch_bytes(int n_img, int beam) {
  const size_t rows = (size_t)n_img * beam;
  return rows * kMaxBeam * 8 + rows * 4;
}
int beam_max_vocab() { return kRowThreads * 64; }

void beam_step(LaunchCtx c, const float* logits, long ldl, int V, int n_img, int beam, int step, int T_total, int pad,
               int stop, int finish_on_stop, const float* score_in, float* score_out, const int* fin_in, int* fin_out,
               const int* hist_in, int* hist_out, int* tok_out, int* src_out, void* scratch, int* trace) {
  if (n_img <= 0) return;
  const long rows = (long)n_img * beam;
  char* sp = static_cast<char*>(scratch);
  float* cv = reinterpret_cast<float*>(sp);
  int* ci = reinterpret_cast<int*>(sp + (size_t)rows * kMaxBeam * 4);
  float* gs = reinterpret_cast<float*>(sp + (size_t)rows * kMaxBeam * 8);
#define LAUNCH_ROWS(EPTV)                                                                                           \
  beam_rows_kernel<EPTV><<<(unsigned)rows, kRowThreads, 0, c.stream>>>(logits, ldl, V, beam, pad, finish_on_stop,   \
                                                                       score_in, fin_in, cv, ci, gs)
  if (V <= kRowThreads * 8) LAUNCH_ROWS(8);
  else if (V <= kRowThreads * 49) LAUNCH_ROWS(49);      // config.py:55 (12295)
  else LAUNCH_ROWS(64);                                  // up to beam_max_vocab()
#undef LAUNCH_ROWS
  beam_merge_kernel<<<(n_img + 3) / 4, 128, 0, c.stream>>>(n_img, V, beam, step, T_total, pad, stop, finish_on_stop, cv, ci,
                                                          gs, score_out, fin_in, fin_out, hist_in, hist_out, tok_out,
                                                          src_out, trace);
  if (c.launches) *c.launches += 2;
}
void beam_init(LaunchCtx c, int R, int beam, int start, float* score, int* tok, int* src) {
  beam_init_kernel<<<(R + 255) / 256, 256, 0, c.stream>>>(R, beam, start, score, tok, src);
  if (c.launches) ++*c.launches;
}
template <typename T>
void beam_reorder(LaunchCtx c, int n_img, int beam, int E, int H, int pad, const T* emb, const int* tok, const int* src,
                  const T* h_new, const float* c_new, T* xin, long ld_xin, float* c_next) {
  const int n_rows = n_img * beam;
  const long n = (long)n_rows * (E / 8 + 2 * (H / 8));
  long blocks = (n + 255) / 256;
  if (blocks > 8L * kNumSMs) blocks = 8L * kNumSMs;
  beam_reorder_kernel<T><<<(int)blocks, 256, 0, c.stream>>>(n_rows, beam, E, H, pad, emb, tok, src, h_new, c_new, xin,
                                                           ld_xin, c_next);
  if (c.launches) ++*c.launches;
}

template void beam_reorder<float>(LaunchCtx, int, int, int, int, int, const float*, const int*, const int*,
                                  const float*, const float*, float*, long, float*);
template void beam_reorder<bf16>(LaunchCtx, int, int, int, int, int, const bf16*, const int*, const int*, const bf16*,
                                 const float*, bf16*, long, float*);

}  // namespace capdec
