// Bandwidth-bound pointwise kernels of the decoder path: conversions, embedding gather / scatter-add,
// the LSTM cell (Paddle lstm_unit, gate order i,f,o,g), the visual sentinel, their backward, and the small
// reductions around them.  All are grid-stride, 128-bit vectorised (8 elements per thread per trip).
#include "kernels.cuh"

namespace capdec {

void set_error(const char* fmt, ...);

namespace {

inline int grid_for(long work_items, int threads = 256, int max_waves = 8) {
  long blocks = (work_items + threads - 1) / threads;
  long cap = (long)kNumSMs * max_waves;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

template <typename T> __global__ void convert_kernel(const float* __restrict__ src, T* __restrict__ dst, long n) {
  pdl_trigger();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = from_f<T>(src[i]);
}

template <typename T>
__global__ void copy2d_kernel(const float* __restrict__ src, long lds, T* __restrict__ dst, long ldd, int R, int C) {
  pdl_trigger();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = (long)R * C;
  for (; i < n; i += stride) {
    long r = i / C, cc = i % C;
    dst[r * ldd + cc] = from_f<T>(src[r * lds + cc]);
  }
}

// An id outside [0, V) (a vocabulary that does not match the checkpoint) reads as padding and raises the handle's
// error word (include/capdec.h, capdec_status); PaddlePaddle's lookup_table_v2 enforces the same range.
template <typename T>
__global__ void embed_gather_kernel(const T* __restrict__ emb, int E, int V, const int64_t* __restrict__ ids, long ids_ld,
                                    int B, int steps, int pad, T* __restrict__ out, long ldo, int* __restrict__ err) {
  pdl_trigger();
  const int vec_per_row = E / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = (long)B * steps * vec_per_row;
  for (; i < n; i += stride) {
    const long m = i / vec_per_row;
    const int v = (int)(i % vec_per_row);
    const int t = (int)(m / B), b = (int)(m % B);
    const long id = ids[(long)b * ids_ld + t];
    const bool in_range = (unsigned long)id < (unsigned long)V;
    if (!in_range && v == 0 && err) atomicOr(err, 1 /* CAPDEC_ERR_SOURCE_ID */);
    float x[8];
    if (id == pad || !in_range) {
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.f;
    } else {
      load8(emb + id * E + v * 8, x);
    }
    store8(out + m * ldo + v * 8, x);
  }
}

template <typename T>
__global__ void lstm_cell_fwd_kernel(int R, int H, const float* __restrict__ gpre, long ld_gpre,
                                     const T* __restrict__ xe, long ld_xe, const float* __restrict__ g5, long ld_g5,
                                     int g5_mod, int g5_div, const float* __restrict__ c_prev, T* __restrict__ h_out,
                                     long ld_h, float* __restrict__ c_out, T* __restrict__ gates_act) {
  constexpr bool F = fast_math<T>::value;
  const int vec_per_row = H / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = (long)R * vec_per_row;
  for (; i < n; i += stride) {
    const long r = i / vec_per_row;
    const int h0 = (int)(i % vec_per_row) * 8;
    float pre[4][8];
#pragma unroll
    for (int g = 0; g < 4; ++g) load8(gpre + r * ld_gpre + g * H + h0, pre[g]);
    if (xe) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float x[8];
        load8(xe + r * ld_xe + g * H + h0, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) pre[g][j] += x[j];
      }
    }
    if (g5) {
      const long gr = (r % g5_mod) / g5_div;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float x[8];
        load8(g5 + gr * ld_g5 + g * H + h0, x);
#pragma unroll
        for (int j = 0; j < 8; ++j) pre[g][j] += x[j];
      }
    }
    float cp[8], cn[8], hn[8];
    load8(c_prev + r * H + h0, cp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ig = sigmoid_t<F>(pre[0][j]), fg = sigmoid_t<F>(pre[1][j]), og = sigmoid_t<F>(pre[2][j]);
      const float gg = tanh_t<F>(pre[3][j]);
      cn[j] = fg * cp[j] + ig * gg;
      hn[j] = og * tanh_t<F>(cn[j]);
      pre[0][j] = ig; pre[1][j] = fg; pre[2][j] = og; pre[3][j] = gg;
    }
    store8(c_out + r * H + h0, cn);
    store8(h_out + r * ld_h + h0, hn);
    if (gates_act) {
#pragma unroll
      for (int g = 0; g < 4; ++g) store8(gates_act + r * 4 * H + g * H + h0, pre[g]);
    }
  }
}

// Inference step: the LSTM cell and the visual sentinel from ONE fp32 pre-activation row [gates i,f,o,g | sentinel gate]
// (model:87-92): pre = gpre5[r] + g5[r / beam];  c = sig(f) c_prev + sig(i) tanh(g);  h = sig(o) tanh(c);
// s = sig(pre[4H:5H]) tanh(c).  (Was two kernels behind two GEMMs; the [R, 5H] product is one GEMM now.)
template <typename T>
__global__ void lstm_cell_sentinel_kernel(int R, int H, const float* __restrict__ gpre5, const float* __restrict__ g5,
                                          int beam, const float* __restrict__ c_prev, T* __restrict__ h_out,
                                          float* __restrict__ c_out, T* __restrict__ s_out) {
  pdl_trigger();
  constexpr bool F = fast_math<T>::value;
  const int vec_per_row = H / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = (long)R * vec_per_row;
  for (; i < n; i += stride) {
    const long r = i / vec_per_row;
    const int h0 = (int)(i % vec_per_row) * 8;
    const float* pr = gpre5 + r * 5 * H + h0;
    const float* gr = g5 + (r / beam) * 5 * H + h0;
    float pre[5][8];
#pragma unroll
    for (int g = 0; g < 5; ++g) {
      float x[8];
      load8(pr + (long)g * H, pre[g]);
      load8(gr + (long)g * H, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) pre[g][j] += x[j];
    }
    float cp[8], cn[8], hn[8], sn[8];
    load8(c_prev + r * H + h0, cp);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ig = sigmoid_t<F>(pre[0][j]), fg = sigmoid_t<F>(pre[1][j]), og = sigmoid_t<F>(pre[2][j]);
      const float gg = tanh_t<F>(pre[3][j]);
      cn[j] = fg * cp[j] + ig * gg;
      const float tc = tanh_t<F>(cn[j]);
      hn[j] = og * tc;
      sn[j] = sigmoid_t<F>(pre[4][j]) * tc;
    }
    store8(c_out + r * H + h0, cn);
    store8(h_out + r * H + h0, hn);
    store8(s_out + r * H + h0, sn);
  }
}

// rowmap (packed rows, see row_map_kernel): row r of ph / sgate / s_out is row rowmap[r] of xe / g5 / cc
template <typename T>
__global__ void sentinel_fwd_kernel(long R, int H, const T* __restrict__ ph, long ld_ph, const T* __restrict__ xe,
                                    long ld_xe, const float* __restrict__ g5, long ld_g5, int g5_mod, int g5_div,
                                    const float* __restrict__ cc, T* __restrict__ sgate, T* __restrict__ s_out,
                                    const int* __restrict__ rowmap) {
  pdl_trigger();
  constexpr bool F = fast_math<T>::value;
  const int vec_per_row = H / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = R * vec_per_row;
  for (; i < n; i += stride) {
    const long r = i / vec_per_row;
    const int h0 = (int)(i % vec_per_row) * 8;
    const long m = rowmap ? rowmap[r] : r;
    float a[8], x[8], c8[8], sg[8], s[8];
    load8(ph + r * ld_ph + h0, a);
    if (xe) {
      load8(xe + m * ld_xe + h0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += x[j];
    }
    if (g5) {
      load8(g5 + ((m % g5_mod) / g5_div) * ld_g5 + h0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += x[j];
    }
    load8(cc + m * H + h0, c8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sg[j] = sigmoid_t<F>(a[j]);
      s[j] = sg[j] * tanh_t<F>(c8[j]);
    }
    if (sgate) store8(sgate + r * H + h0, sg);
    store8(s_out + r * H + h0, s);
  }
}

template <typename T>
__global__ void sentinel_bwd_kernel(long R, int H, const T* __restrict__ dS, const T* __restrict__ sgate,
                                    const float* __restrict__ cc, T* __restrict__ dsg_pre, long ld_dsg,
                                    float* __restrict__ dC_head) {
  pdl_trigger();
  constexpr bool F = fast_math<T>::value;
  const int vec_per_row = H / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = R * vec_per_row;
  for (; i < n; i += stride) {
    const long r = i / vec_per_row;
    const int h0 = (int)(i % vec_per_row) * 8;
    float d[8], sg[8], c8[8], o1[8], o2[8];
    load8(dS + r * H + h0, d);
    load8(sgate + r * H + h0, sg);
    load8(cc + r * H + h0, c8);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float tc = tanh_t<F>(c8[j]);
      o1[j] = d[j] * tc * sg[j] * (1.f - sg[j]);
      o2[j] = d[j] * sg[j] * (1.f - tc * tc);
    }
    store8(dsg_pre + r * ld_dsg + h0, o1);
    store8(dC_head + r * H + h0, o2);
  }
}

// The same over PADDED rows when the head ran on packed rows: padded row m is packed row prow[m] (-1: a row past the
// caption's last target, no gradient).  d(sentinel gate pre-activation) goes to dG[:, 4H:5H] and d c to dC_head in the
// padded layout the BPTT kernel reads (zeros in dead rows), and to dsg_packed for the packed weight / data gradients.
template <typename T>
__global__ void sentinel_bwd_packed_kernel(long M, int H, const int* __restrict__ prow, const T* __restrict__ dS,
                                           const T* __restrict__ sgate, const float* __restrict__ cc,
                                           T* __restrict__ dsg_pre, long ld_dsg, float* __restrict__ dC_head,
                                           T* __restrict__ dsg_packed) {
  pdl_trigger();
  constexpr bool F = fast_math<T>::value;
  const int vec_per_row = H / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = M * vec_per_row;
  for (; i < n; i += stride) {
    const long m = i / vec_per_row;
    const int h0 = (int)(i % vec_per_row) * 8;
    const long r = prow[m];
    float o1[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o2[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r >= 0) {
      float d[8], sg[8], c8[8];
      load8(dS + r * H + h0, d);
      load8(sgate + r * H + h0, sg);
      load8(cc + m * H + h0, c8);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float tc = tanh_t<F>(c8[j]);
        o1[j] = d[j] * tc * sg[j] * (1.f - sg[j]);
        o2[j] = d[j] * sg[j] * (1.f - tc * tc);
      }
      store8(dsg_packed + r * H + h0, o1);
    }
    store8(dsg_pre + m * ld_dsg + h0, o1);
    store8(dC_head + m * H + h0, o2);
  }
}

// ---- packed rows of the time-batched head (capdec_train_fwd with lengths) -----------------------------
// Captions are [start, w.., stop, pad..] (ai_challenge_tokenizer.py:81-86), the loss mask is target != pad
// (model:163-171): a row (t, b) past the last non-padding target of caption b has d logits == 0 exactly, and nothing
// downstream of (h_t, c_t) feeds back into the recurrence, so the head only has to run on the rows
// t < steps[b], steps[b] = 1 + the last t with caption[b, t+1] != pad (rows in a hole of the mask stay: they are masked by
// the cross entropy as always).  Packed order = padded (time-major) order with the dead rows removed:
//   prow[t*B + b] = packed row of (t, b), or -1;   pm[r] = padded row of packed row r.
// The host sized every launch for Mp = sum_b (lengths[b] - 1) rows; the device works from the captions themselves, a
// different total raises CAPDEC_ERR_LENGTHS and surplus rows are dropped, so a wrong `lengths` cannot index out of bounds.
constexpr int kMapThreads = 1024;
__global__ void __launch_bounds__(kMapThreads)
caption_steps_kernel(const int64_t* __restrict__ caption, long cap_ld, int B, int T, int pad, long Mp,
                     int* __restrict__ steps, int* __restrict__ err) {
  __shared__ int wsum[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int part = 0;
  for (int b = tid; b < B; b += kMapThreads) {
    int n = 0;
    for (int t = 0; t < T; ++t)
      if (caption[(long)b * cap_ld + t + 1] != pad) n = t + 1;
    steps[b] = n;
    part += n;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) wsum[warp] = part;
  __syncthreads();
  if (tid == 0) {
    long total = 0;
    for (int w = 0; w < kMapThreads / 32; ++w) total += wsum[w];
    if (total != Mp && err) atomicOr(err, 4 /* CAPDEC_ERR_LENGTHS */);
  }
}
// one CTA per step t: its first packed row is sum_b min(steps[b], t), the rank of b inside the step a block scan
__global__ void __launch_bounds__(kMapThreads)
row_map_kernel(const int* __restrict__ steps, int B, int T, long Mp, int* __restrict__ prow, int* __restrict__ pm) {
  __shared__ int wsum[32];
  __shared__ int s_total;
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int part = 0;
  for (int b = tid; b < B; b += kMapThreads) part += min(steps[b], t);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if (lane == 0) wsum[warp] = part;
  __syncthreads();
  if (tid == 0) {
    int a = 0;
    for (int w = 0; w < kMapThreads / 32; ++w) a += wsum[w];
    s_total = a;
  }
  __syncthreads();
  int base = s_total;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += kMapThreads) {
    const int b = b0 + tid;
    const int live = (b < B && steps[b] > t) ? 1 : 0;
    int incl = live;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < kMapThreads / 32; ++w) {
      const int v = wsum[w];
      if (w < warp) before += v;
      total += v;
    }
    if (b < B) {
      const long m = (long)t * B + b;
      const long r = (long)base + before + incl - 1;
      if (live && r < Mp) {
        prow[m] = (int)r;
        pm[r] = (int)m;
      } else {
        prow[m] = -1;
      }
    }
    base += total;
    __syncthreads();
  }
}

// dst[r, :] = src[pm[r], :] for up to two tensors of the same shape (h_{t-1} and h_t rows of the packed head)
template <typename T>
__global__ void gather_rows_kernel(const T* __restrict__ src0, const T* __restrict__ src1, long ld_src,
                                   const int* __restrict__ pm, long Mp, int N, T* __restrict__ dst0,
                                   T* __restrict__ dst1, long ld_dst) {
  pdl_trigger();
  const int vec_per_row = N / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = Mp * vec_per_row;
  for (; i < n; i += stride) {
    const long r = i / vec_per_row;
    const int n0 = (int)(i % vec_per_row) * 8;
    const long m = pm[r];
    float v[8];
    load8(src0 + m * ld_src + n0, v);
    store8(dst0 + r * ld_dst + n0, v);
    if (src1) {
      load8(src1 + m * ld_src + n0, v);
      store8(dst1 + r * ld_dst + n0, v);
    }
  }
}
// dst[m, :] = prow[m] >= 0 ? src[prow[m], :] : 0 for up to two tensors [M, N] (pitch N)
template <typename T>
__global__ void scatter_rows_zero_kernel(const T* __restrict__ src0, const T* __restrict__ src1,
                                         const int* __restrict__ prow, long M, int N, T* __restrict__ dst0,
                                         T* __restrict__ dst1) {
  pdl_trigger();
  const int vec_per_row = N / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = M * vec_per_row;
  for (; i < n; i += stride) {
    const long m = i / vec_per_row;
    const int n0 = (int)(i % vec_per_row) * 8;
    const long r = prow[m];
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (r >= 0) load8(src0 + r * N + n0, v);
    store8(dst0 + m * N + n0, v);
    if (src1) {
      float w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (r >= 0) load8(src1 + r * N + n0, w);
      store8(dst1 + m * N + n0, w);
    }
  }
}

// One BPTT step (backward of Paddle lstm_unit + the state feedback of model:94-95):
//   dh = dH_head[t] + dHprev_head[t+1] + dh_rec          (p_hid path, p_hidden path of step t+1, gates of step t+1)
//   dc = dC_head[t] + dc_rec + dh * o * (1 - tanh(c_t)^2)
//   d(pre_i) = dc g i(1-i);  d(pre_f) = dc c_prev f(1-f);  d(pre_o) = dh tanh(c_t) o(1-o);  d(pre_g) = dc i (1-g^2)
//   dc_rec <- dc * f
template <typename T>
__global__ void lstm_cell_bwd_kernel(int B, int H, const T* __restrict__ dH_head, const T* __restrict__ dHprev_head,
                                     const float* __restrict__ dh_rec, const float* __restrict__ dC_head,
                                     float* __restrict__ dc_rec, const T* __restrict__ gates_act,
                                     const float* __restrict__ c_t, const float* __restrict__ c_prev,
                                     T* __restrict__ dgates, long ld_dg) {
  constexpr bool F = fast_math<T>::value;
  const int vec_per_row = H / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = (long)B * vec_per_row;
  for (; i < n; i += stride) {
    const long r = i / vec_per_row;
    const int h0 = (int)(i % vec_per_row) * 8;
    float dh[8], x[8], dc[8], ct[8], cp[8], ga[4][8], dg[4][8];
    load8(dH_head + r * H + h0, dh);
    if (dHprev_head) {
      load8(dHprev_head + r * H + h0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) dh[j] += x[j];
    }
    if (dh_rec) {
      load8(dh_rec + r * H + h0, x);
#pragma unroll
      for (int j = 0; j < 8; ++j) dh[j] += x[j];
    }
    load8(dC_head + r * H + h0, dc);
    load8(dc_rec + r * H + h0, x);
    load8(c_t + r * H + h0, ct);
    load8(c_prev + r * H + h0, cp);
#pragma unroll
    for (int g = 0; g < 4; ++g) load8(gates_act + r * 4 * H + g * H + h0, ga[g]);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ig = ga[0][j], fg = ga[1][j], og = ga[2][j], gg = ga[3][j];
      const float tc = tanh_t<F>(ct[j]);
      const float dcj = dc[j] + x[j] + dh[j] * og * (1.f - tc * tc);
      dg[0][j] = dcj * gg * ig * (1.f - ig);
      dg[1][j] = dcj * cp[j] * fg * (1.f - fg);
      dg[2][j] = dh[j] * tc * og * (1.f - og);
      dg[3][j] = dcj * ig * (1.f - gg * gg);
      x[j] = dcj * fg;
    }
    store8(dc_rec + r * H + h0, x);
#pragma unroll
    for (int g = 0; g < 4; ++g) store8(dgates + r * ld_dg + g * H + h0, dg[g]);
  }
}

template <typename T>
__global__ void sum_over_steps_kernel(const T* __restrict__ x, long ld, int B, int steps, int N, T* __restrict__ out,
                                      long ldo, const int* __restrict__ live) {
  pdl_trigger();
  const int vec_per_row = N / 8;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = (long)B * vec_per_row;
  for (; i < n; i += stride) {
    const long b = i / vec_per_row;
    const int n0 = (int)(i % vec_per_row) * 8;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int nt = live ? min(max(live[b], 0), steps) : steps;
    for (int t = 0; t < nt; ++t) {
      float v[8];
      load8(x + ((long)t * B + b) * ld + n0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
    store8(out + b * ldo + n0, acc);
  }
}

// grid (ceil(N/256), row_chunks), block = 32 column-threads x 16 row groups (one warp = one row group: 512
// contiguous bytes per row).  Each thread owns 8 adjacent columns of every 16th row of the chunk with four 128-bit
// loads in flight; the 16 row groups are reduced in shared memory and the block issues one set of fp32 atomics.
// Same-address atomics from different SMs serialise at ~100 ns each, so the chunk count (= atomics per address) is
// kept at <= 32 and parallelism comes from the column split instead.
template <typename T>
__global__ void __launch_bounds__(512)
colsum_kernel(const T* __restrict__ x, long ld, long M, int N, long rows_per_chunk, float* __restrict__ out,
              const float* __restrict__ scale_dev) {
  pdl_trigger();
  __shared__ float red[15][8][33];
  const float sc = scale_dev ? __ldg(scale_dev) : 1.f;
  const int ct = threadIdx.x & 31, rg = threadIdx.x >> 5;  // column thread, row group
  const int n0 = (blockIdx.x * 32 + ct) * 8;
  const long m0 = (long)blockIdx.y * rows_per_chunk;
  const long m1 = (m0 + rows_per_chunk < M) ? m0 + rows_per_chunk : M;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (n0 + 8 <= N) {
    long m = m0 + rg;
    for (; m + 48 < m1; m += 64) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(x + (m + 16 * u) * ld + n0, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
    for (; m < m1; m += 16) {
      float v[8];
      load8(x + m * ld + n0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  } else if (n0 < N) {  // ragged right edge
    for (long m = m0 + rg; m < m1; m += 16)
      for (int j = 0; j < 8 && n0 + j < N; ++j) acc[j] += to_f(x[m * ld + n0 + j]);
  }
  if (rg > 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[rg - 1][j][ct] = acc[j];
  }
  __syncthreads();
  if (rg == 0 && n0 < N) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[j];
#pragma unroll
      for (int g = 0; g < 15; ++g) v += red[g][j][ct];
      if (n0 + j < N) atomicAdd(out + n0 + j, v * sc);
    }
  }
}

// Narrow variant for N <= 4096: 64 columns per CTA (8 column threads x 64 row groups), so a [8704 x 1024] sum runs on
// 16 x 32 = 512 CTAs with every load of a thread in flight at once, while each output address still receives only
// 32 atomics (see above).  Row groups meet through two shuffles and one shared-memory step.
template <typename T>
__device__ __forceinline__ void colsum_narrow_body(const T* __restrict__ x, long ld, long M, int N, long rows_per_chunk,
                                                   float* __restrict__ out, float sc, int bx, int by) {
  __shared__ float red[16][8][9];
  const int ct = threadIdx.x & 7, rg = threadIdx.x >> 3, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = (bx * 8 + ct) * 8;
  const long m0 = (long)by * rows_per_chunk;
  const long m1 = (m0 + rows_per_chunk < M) ? m0 + rows_per_chunk : M;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (n0 + 8 <= N) {
    long m = m0 + rg;
    for (; m + 192 < m1; m += 256) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u) load8(x + (m + 64 * u) * ld + n0, v[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
    }
    for (; m < m1; m += 64) {
      float v[8];
      load8(x + m * ld + n0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  } else if (n0 < N) {  // ragged right edge
    for (long m = m0 + rg; m < m1; m += 64)
      for (int j = 0; j < 8 && n0 + j < N; ++j) acc[j] += to_f(x[m * ld + n0 + j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 8);
    acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int j = 0; j < 8; ++j) red[warp][j][ct] = acc[j];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int c = threadIdx.x & 7, j = threadIdx.x >> 3;
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 16; ++w) v += red[w][j][c];
    const int n = (bx * 8 + c) * 8 + j;
    if (n < N) atomicAdd(out + n, v * sc);
  }
}
template <typename T>
__global__ void __launch_bounds__(512)
colsum_narrow_kernel(const T* __restrict__ x, long ld, long M, int N, long rows_per_chunk, float* __restrict__ out,
                     const float* __restrict__ scale_dev) {
  pdl_trigger();
  colsum_narrow_body<T>(x, ld, M, N, rows_per_chunk, out, scale_dev ? __ldg(scale_dev) : 1.f, blockIdx.x, blockIdx.y);
}
// several independent column sums (the bias gradients of the time-batched head) in one launch: blockIdx.x runs over the
// 64-column blocks of all jobs.  Each of them alone is a 3 us kernel that costs 10 us as a launch of its own.
template <typename T>
__global__ void __launch_bounds__(512) colsum_multi_kernel(ColsumTable tab) {
  pdl_trigger();
  int j = 0;
  while (j + 1 < tab.n && (int)blockIdx.x >= tab.first_block[j + 1]) ++j;
  const ColsumJob& job = tab.job[j];
  colsum_narrow_body<T>(reinterpret_cast<const T*>(job.x), job.ld, job.M, job.N, job.rows_per_chunk, job.out, 1.f,
                        (int)blockIdx.x - tab.first_block[j], blockIdx.y);
}

template <typename T>
__global__ void embed_scatter_kernel(const T* __restrict__ dX, long ld, int E, int V, const int64_t* __restrict__ ids,
                                     long ids_ld, int B, long rows, int pad, float* __restrict__ dEmb,
                                     const int* __restrict__ pm) {
  pdl_trigger();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x, n = rows * E;
  for (; i < n; i += stride) {
    const long m = i / E;                         // row of dX
    const int e = (int)(i % E);
    const long mp = pm ? pm[m] : m;               // its (step, sample)
    const int t = (int)(mp / B), b = (int)(mp % B);
    const long id = ids[(long)b * ids_ld + t];
    // ids outside [0, V) were reported by the forward gather and read as padding there: no gradient either
    if (id != pad && (unsigned long)id < (unsigned long)V) atomicAdd(dEmb + id * E + e, to_f(dX[m * ld + e]));
  }
}

// every fused bias vector of capdec_bind_params in one launch (was: 2 memsets / 5 copies / an add / a permute launch)
__global__ void bind_bias_kernel(BiasBind a) {
  pdl_trigger();
  const int H = a.H, n5 = 5 * a.H;
  const int n = a.Vp > n5 ? a.Vp : n5;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (i < a.Vp) a.b_vocab[i] = i < a.V ? a.out_fc_bias[i] : 0.f;   // padded so the GEMM epilogue reads whole groups of 8
    if (i < n5) {
      const bool gate_col = i < 4 * H;
      const float v = gate_col ? a.lstm_b[i] : a.fc_xs_b[i - 4 * H];
      a.b5[i] = v;                                                    // train:  [lstm_b | fc_xs_b]
      a.b5d[i] = gate_col ? v : v + a.fc_hs_b[i - 4 * H];             // decode: [lstm_b | fc_xs_b + fc_hs_b]
      if (a.b5p) a.b5p[gate_col ? 4 * (i % H) + i / H : i] = v;       // unit-major gate columns (col 4*unit + gate)
    }
  }
}

// ---- multi-tensor kernels: one launch for the whole parameter re-layout / gradient clearing ----
constexpr int kChunk = 4096;  // elements per block

template <typename T>
__global__ void __launch_bounds__(256) multi_convert_kernel(const __grid_constant__ ConvertTable tab) {
  pdl_trigger();
  int j = 0;
  while (j + 1 < tab.n && (int)blockIdx.x >= tab.first_block[j + 1]) ++j;
  const ConvertJob& job = tab.job[j];
  const long total = (long)job.rows * job.cols;
  const long base = (long)(blockIdx.x - tab.first_block[j]) * kChunk;
  T* dst = reinterpret_cast<T*>(job.dst);
#pragma unroll 4
  for (int v = threadIdx.x; v < kChunk / 4; v += 256) {
    const long e = base + (long)v * 4;
    if (e >= total) break;
    const long r = e / job.cols;
    const int c = (int)(e % job.cols);
    float4 x;
    if (job.mode == 0) {
      x = *reinterpret_cast<const float4*>(job.src + r * job.lds + c);
    } else {  // gate permutation of recurrent.cu: dst col 64*jj + 4*u + gate <- src col gate*H + 16*jj + u
      const float* sp = job.src + r * job.lds + (c >> 6) * 16 + ((c >> 2) & 15);
      x = make_float4(sp[0], sp[job.H], sp[2 * (long)job.H], sp[3 * (long)job.H]);
    }
    T* d = dst + r * job.ldd + c;
    if (sizeof(T) == 4) {
      *reinterpret_cast<float4*>(d) = x;
    } else {
      uint2 u;
      *reinterpret_cast<__nv_bfloat162*>(&u.x) = __floats2bfloat162_rn(x.x, x.y);
      *reinterpret_cast<__nv_bfloat162*>(&u.y) = __floats2bfloat162_rn(x.z, x.w);
      *reinterpret_cast<uint2*>(d) = u;
    }
  }
}

__global__ void __launch_bounds__(256) multi_zero_kernel(const __grid_constant__ ZeroTable tab) {
  pdl_trigger();
  int j = 0;
  while (j + 1 < tab.n && (int)blockIdx.x >= tab.first_block[j + 1]) ++j;
  const long base = (long)(blockIdx.x - tab.first_block[j]) * kChunk;
  float* p = tab.ptr[j];
  const long n = tab.count[j];
  if (tab.vec[j]) {
#pragma unroll 4
    for (int v = threadIdx.x; v < kChunk / 4; v += 256) {
      const long e = base + (long)v * 4;
      if (e + 4 <= n) *reinterpret_cast<float4*>(p + e) = make_float4(0.f, 0.f, 0.f, 0.f);
      else for (long i = e; i < n; ++i) p[i] = 0.f;
    }
  } else {
    for (int v = threadIdx.x; v < kChunk; v += 256)
      if (base + v < n) p[base + v] = 0.f;
  }
}

}  // namespace

#define BUMP(c) do { if ((c).launches) ++*(c).launches; } while (0)

template <typename T> int multi_convert(LaunchCtx c, ConvertTable& tab) {
  int blocks = 0;
  for (int j = 0; j < tab.n; ++j) {
    const ConvertJob& job = tab.job[j];
    const size_t dst_al = sizeof(T) * 4;
    if ((job.cols % 4) || (job.lds % 4) || (job.ldd % 4) || ((uintptr_t)job.src % 16) || ((uintptr_t)job.dst % dst_al) ||
        (job.mode == 1 && (job.cols != 4 * job.H || job.H % 16)))
      return 1;  // not vectorisable: the caller falls back to the per-tensor kernels
    tab.first_block[j] = blocks;
    blocks += (int)(((long)job.rows * job.cols + kChunk - 1) / kChunk);
  }
  tab.first_block[tab.n] = blocks;
  if (blocks == 0) return 0;
  multi_convert_kernel<T><<<blocks, 256, 0, c.stream>>>(tab);
  BUMP(c);
  return 0;
}

void multi_zero(LaunchCtx c, ZeroTable& tab) {
  int blocks = 0;
  for (int j = 0; j < tab.n; ++j) {
    tab.vec[j] = ((uintptr_t)tab.ptr[j] % 16) == 0;
    tab.first_block[j] = blocks;
    blocks += (int)((tab.count[j] + kChunk - 1) / kChunk);
  }
  tab.first_block[tab.n] = blocks;
  if (blocks == 0) return;
  multi_zero_kernel<<<blocks, 256, 0, c.stream>>>(tab);
  BUMP(c);
}

template <typename T> void convert_from_f32(LaunchCtx c, const float* src, T* dst, long n) {
  if (n <= 0) return;
  convert_kernel<T><<<grid_for(n), 256, 0, c.stream>>>(src, dst, n);
  BUMP(c);
}
template <typename T> void copy2d_from_f32(LaunchCtx c, const float* src, long lds, T* dst, long ldd, int R, int C) {
  copy2d_kernel<T><<<grid_for((long)R * C), 256, 0, c.stream>>>(src, lds, dst, ldd, R, C);
  BUMP(c);
}
template <typename T>
void embed_gather(LaunchCtx c, const T* emb, int E, int V, const int64_t* ids, long ids_ld, int B, int steps, int pad,
                  T* out, long ldo, int* err) {
  embed_gather_kernel<T><<<grid_for((long)B * steps * (E / 8)), 256, 0, c.stream>>>(emb, E, V, ids, ids_ld, B, steps, pad,
                                                                                 out, ldo, err);
  BUMP(c);
}
template <typename T>
void lstm_cell_fwd(LaunchCtx c, int R, int H, const float* gpre, long ld_gpre, const T* xe, long ld_xe, const float* g5,
                   long ld_g5, int g5_mod, int g5_div, const float* c_prev, T* h_out, long ld_h, float* c_out,
                   T* gates_act) {
  lstm_cell_fwd_kernel<T><<<grid_for((long)R * (H / 8), 128), 128, 0, c.stream>>>(
      R, H, gpre, ld_gpre, xe, ld_xe, g5, ld_g5, g5_mod, g5_div, c_prev, h_out, ld_h, c_out, gates_act);
  BUMP(c);
}
template <typename T>
void sentinel_fwd(LaunchCtx c, long R, int H, const T* ph, long ld_ph, const T* xe, long ld_xe, const float* g5,
                  long ld_g5, int g5_mod, int g5_div, const float* cc, T* sgate, T* s_out, const int* rowmap) {
  if (R <= 0) return;
  sentinel_fwd_kernel<T><<<grid_for(R * (H / 8)), 256, 0, c.stream>>>(R, H, ph, ld_ph, xe, ld_xe, g5, ld_g5, g5_mod,
                                                                    g5_div, cc, sgate, s_out, rowmap);
  BUMP(c);
}
template <typename T>
void lstm_cell_sentinel_fwd(LaunchCtx c, int R, int H, const float* gpre5, const float* g5, int beam, const float* c_prev,
                            T* h_out, float* c_out, T* s_out) {
  if (R <= 0) return;
  lstm_cell_sentinel_kernel<T><<<grid_for((long)R * (H / 8), 128), 128, 0, c.stream>>>(R, H, gpre5, g5, beam, c_prev, h_out,
                                                                                     c_out, s_out);
  BUMP(c);
}
template <typename T>
void sentinel_bwd_packed(LaunchCtx c, long M, int H, const int* prow, const T* dS, const T* sgate, const float* cc,
                         T* dsg_pre, long ld_dsg, float* dC_head, T* dsg_packed) {
  if (M <= 0) return;
  sentinel_bwd_packed_kernel<T><<<grid_for(M * (H / 8)), 256, 0, c.stream>>>(M, H, prow, dS, sgate, cc, dsg_pre, ld_dsg,
                                                                           dC_head, dsg_packed);
  BUMP(c);
}
int build_row_map(LaunchCtx c, const int64_t* caption, long cap_ld, int B, int T, int pad, long Mp, int* steps, int* prow,
                  int* pm, int* err) {
  if (B <= 0 || T <= 0) return 0;
  caption_steps_kernel<<<1, kMapThreads, 0, c.stream>>>(caption, cap_ld, B, T, pad, Mp, steps, err);
  row_map_kernel<<<T, kMapThreads, 0, c.stream>>>(steps, B, T, Mp, prow, pm);
  if (c.launches) *c.launches += 2;
  return 0;
}
template <typename T>
void gather_rows(LaunchCtx c, const T* src0, const T* src1, long ld_src, const int* pm, long Mp, int N, T* dst0, T* dst1,
                 long ld_dst) {
  if (Mp <= 0) return;
  gather_rows_kernel<T><<<grid_for(Mp * (N / 8)), 256, 0, c.stream>>>(src0, src1, ld_src, pm, Mp, N, dst0, dst1, ld_dst);
  BUMP(c);
}
template <typename T>
void scatter_rows_zero(LaunchCtx c, const T* src0, const T* src1, const int* prow, long M, int N, T* dst0, T* dst1) {
  if (M <= 0) return;
  scatter_rows_zero_kernel<T><<<grid_for(M * (N / 8)), 256, 0, c.stream>>>(src0, src1, prow, M, N, dst0, dst1);
  BUMP(c);
}
template <typename T>
void sentinel_bwd(LaunchCtx c, long R, int H, const T* dS, const T* sgate, const float* cc, T* dsg_pre, long ld_dsg,
                  float* dC_head) {
  sentinel_bwd_kernel<T><<<grid_for(R * (H / 8)), 256, 0, c.stream>>>(R, H, dS, sgate, cc, dsg_pre, ld_dsg, dC_head);
  BUMP(c);
}
template <typename T>
void lstm_cell_bwd(LaunchCtx c, int B, int H, const T* dH_head, const T* dHprev_head, const float* dh_rec,
                   const float* dC_head, float* dc_rec, const T* gates_act, const float* c_t, const float* c_prev,
                   T* dgates, long ld_dg) {
  lstm_cell_bwd_kernel<T><<<grid_for((long)B * (H / 8), 128), 128, 0, c.stream>>>(
      B, H, dH_head, dHprev_head, dh_rec, dC_head, dc_rec, gates_act, c_t, c_prev, dgates, ld_dg);
  BUMP(c);
}
template <typename T>
void sum_over_steps(LaunchCtx c, const T* x, long ld, int B, int steps, int N, T* out, long ldo, const int* live) {
  sum_over_steps_kernel<T><<<grid_for((long)B * (N / 8), 128), 128, 0, c.stream>>>(x, ld, B, steps, N, out, ldo, live);
  BUMP(c);
}
template <typename T> void colsum_add(LaunchCtx c, const T* x, long ld, long M, int N, float* out, const float* scale_dev) {
  if (M <= 0 || N <= 0) return;
  // 128-bit loads need 16-byte (bf16) / 32-byte (fp32) aligned rows
  if ((ld % 8) || ((uintptr_t)x % (8 * sizeof(T)))) {
    set_error("colsum_add: operand must have an 8-element aligned pitch and base");
    return;
  }
  const int col_blocks = (N + 255) / 256;
  long chunks = 32;  // measured: 4x more (smaller) chunks is slower in stream order (+80 us per train step)
  if (chunks > (M + 15) / 16) chunks = (M + 15) / 16;
  if (chunks < 1) chunks = 1;
  const long rpc = (M + chunks - 1) / chunks;
  if (N <= 4096 && M >= 2048) {
    dim3 grid((N + 63) / 64, (unsigned)((M + rpc - 1) / rpc));
    colsum_narrow_kernel<T><<<grid, 512, 0, c.stream>>>(x, ld, M, N, rpc, out, scale_dev);
  } else {
    dim3 grid(col_blocks, (unsigned)((M + rpc - 1) / rpc));
    colsum_kernel<T><<<grid, 512, 0, c.stream>>>(x, ld, M, N, rpc, out, scale_dev);
  }
  BUMP(c);
}
template <typename T> void colsum_multi(LaunchCtx c, ColsumTable& tab) {
  if (tab.n == 0) return;
  const int chunks = 32;
  bool batched = true;
  for (int j = 0; j < tab.n; ++j) {
    const ColsumJob& job = tab.job[j];
    if (!(job.N <= 4096 && job.M >= 2048) || (job.ld % 8) || ((uintptr_t)job.x % (8 * sizeof(T)))) batched = false;
  }
  if (!batched) {  // small problems (tests): one launch each, colsum_add picks the kernel
    for (int j = 0; j < tab.n; ++j)
      colsum_add<T>(c, reinterpret_cast<const T*>(tab.job[j].x), tab.job[j].ld, tab.job[j].M, tab.job[j].N, tab.job[j].out,
                    nullptr);
    return;
  }
  int blocks = 0;
  for (int j = 0; j < tab.n; ++j) {
    tab.job[j].rows_per_chunk = (tab.job[j].M + chunks - 1) / chunks;
    tab.first_block[j] = blocks;
    blocks += (tab.job[j].N + 63) / 64;
  }
  tab.first_block[tab.n] = blocks;
  colsum_multi_kernel<T><<<dim3(blocks, chunks), 512, 0, c.stream>>>(tab);
  BUMP(c);
}
template <typename T>
void embed_scatter_add(LaunchCtx c, const T* dX, long ld, int E, int V, const int64_t* ids, long ids_ld, int B, int steps,
                       int pad, float* dEmb, const int* pm, long rows) {
  if (!pm) rows = (long)B * steps;
  if (rows <= 0) return;
  embed_scatter_kernel<T><<<grid_for(rows * E), 256, 0, c.stream>>>(dX, ld, E, V, ids, ids_ld, B, rows, pad, dEmb, pm);
  BUMP(c);
}
void bind_biases(LaunchCtx c, const BiasBind& a) {
  const int n = a.Vp > 5 * a.H ? a.Vp : 5 * a.H;
  bind_bias_kernel<<<(n + 255) / 256, 256, 0, c.stream>>>(a);
  BUMP(c);
}
template int multi_convert<float>(LaunchCtx, ConvertTable&);
template int multi_convert<bf16>(LaunchCtx, ConvertTable&);

#define INSTANTIATE(T)                                                                                                 \
  template void convert_from_f32<T>(LaunchCtx, const float*, T*, long);                                                \
  template void copy2d_from_f32<T>(LaunchCtx, const float*, long, T*, long, int, int);                                 \
  template void embed_gather<T>(LaunchCtx, const T*, int, int, const int64_t*, long, int, int, int, T*, long, int*);   \
  template void lstm_cell_fwd<T>(LaunchCtx, int, int, const float*, long, const T*, long, const float*, long, int,    \
                                 int, const float*, T*, long, float*, T*);                                             \
  template void sentinel_fwd<T>(LaunchCtx, long, int, const T*, long, const T*, long, const float*, long, int, int,    \
                                const float*, T*, T*, const int*);                                                     \
  template void lstm_cell_sentinel_fwd<T>(LaunchCtx, int, int, const float*, const float*, int, const float*, T*,      \
                                          float*, T*);                                                                 \
  template void sentinel_bwd_packed<T>(LaunchCtx, long, int, const int*, const T*, const T*, const float*, T*, long,   \
                                       float*, T*);                                                                    \
  template void gather_rows<T>(LaunchCtx, const T*, const T*, long, const int*, long, int, T*, T*, long);              \
  template void scatter_rows_zero<T>(LaunchCtx, const T*, const T*, const int*, long, int, T*, T*);                    \
  template void sentinel_bwd<T>(LaunchCtx, long, int, const T*, const T*, const float*, T*, long, float*);             \
  template void lstm_cell_bwd<T>(LaunchCtx, int, int, const T*, const T*, const float*, const float*, float*,          \
                                 const T*, const float*, const float*, T*, long);                                      \
  template void sum_over_steps<T>(LaunchCtx, const T*, long, int, int, int, T*, long, const int*);                          \
  template void colsum_add<T>(LaunchCtx, const T*, long, long, int, float*, const float*);                                           \
  template void colsum_multi<T>(LaunchCtx, ColsumTable&);                                           \
  template void embed_scatter_add<T>(LaunchCtx, const T*, long, int, int, const int64_t*, long, int, int, int, float*, \
                                     const int*, long);

INSTANTIATE(float)
INSTANTIATE(bf16)

}  // namespace capdec
