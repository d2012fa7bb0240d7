// GEMM interface used by the decoder schedule.
//
//   C[M,N] (=|+=) epi( sum_k A(m,k) * B(n,k) )
//
// Operand forms ("majors"), chosen per operand:
//   K-major  (a_mn / b_mn = false): A is stored [M,K] row-major (lda), B is stored [N,K] row-major (ldb).
//   MN-major (a_mn / b_mn = true) : A is stored [K,M] row-major (lda), B is stored [K,N] row-major (ldb).
//   forward fc   Y = X.W        : A = X K-major,  B = W [in,out] MN-major  (PaddlePaddle's own layout)
//   data grad    dX = dY.W^T    : A = dY K-major, B = W [in,out] K-major   (N = in, K = out)
//   weight grad  dW = X^T.dY    : A = X MN-major, B = dY MN-major          (contraction over the row index)
// T = bf16 -> tcgen05/TMEM/TMA kernel (gemm_tc.cu); T = float -> SIMT kernel (validation precision mode).
#pragma once
#include "common.cuh"

namespace capdec {

//   ACT_ROWSTATS     (tensor-core kernels only) no C output; per (row, 64-column group g) the partial max and
//                    sum exp(x - max) of x = acc + bias go to stats_m / stats_s [M, stats_ld] at column g
//   ACT_RELU         max(x, 0)            (the feature adaptor, model:196-198)
//   ACT_DRELU        x * [y > 0] with y = dact_y, the ReLU output
enum { ACT_NONE = 0, ACT_TANH = 1, ACT_DTANH = 2, ACT_ROWSTATS = 3, ACT_RELU = 5, ACT_DRELU = 6 };
// which kernel gemm<bf16> runs (GemmContext::variant; anything but AUTO is for the kernel-level parity tests)
enum { GEMM_AUTO = 0, GEMM_SIMT = 1, GEMM_PAIR = 2, GEMM_SINGLE = 3 };

struct GemmEpi {
  const float* bias = nullptr;   // [N] fp32, added to every row
  const void* addend = nullptr;  // T [M, ld_add], added element-wise
  const void* dact_y = nullptr;  // T [M, ld_y]: ACT_DTANH multiplies by (1 - y^2)
  long ld_add = 0, ld_y = 0;
  long batch_add = 0, batch_y = 0;  // element strides between batches of addend / dact_y
  int add_mod = 0;               // > 0: the addend row of output row m is m % add_mod (a [add_mod, ld_add] block broadcast
                                 // down the rows: the per-image term of a time-major [T*B, N] activation)
  int act = ACT_NONE;
  int out_f32 = 0;               // C is fp32 (otherwise T)
  int atomic = 0;                // fp32 atomicAdd into C (split-K / gradient accumulation)
  int aux_padded = 0;            // bias / addend / dact_y are readable up to the next multiple of 8 columns past N
  const float* alpha_dev = nullptr;  // device scalar multiplied into the accumulator before anything else
  float* stats_m = nullptr;      // ACT_ROWSTATS outputs
  float* stats_s = nullptr;
  long stats_ld = 0;
};

struct GemmCall {
  const void* A = nullptr; long lda = 0;
  const void* B = nullptr; long ldb = 0;
  void* C = nullptr; long ldc = 0;
  int M = 0, N = 0, K = 0;
  bool a_mn = false;             // A stored [K,M] (MN-major) instead of [M,K]
  bool b_mn = false;             // B stored [K,N] (MN-major) instead of [N,K]
  int splits = 1;                // split-K (requires epi.atomic); 0 = choose automatically
  int batch = 1;                 // independent problems of the same shape; element strides between them:
  long batch_a = 0, batch_b = 0, batch_c = 0;
  GemmEpi epi;
};

// element-wise epilogue shared by both kernels
template <typename T>
__device__ __forceinline__ float epi_value(float acc, int m, int n, const GemmEpi& e) {
  float v = e.alpha_dev ? acc * __ldg(e.alpha_dev) : acc;
  if (e.bias) v += __ldg(e.bias + n);
  if (e.addend) v += to_f(reinterpret_cast<const T*>(e.addend)[(long)(e.add_mod > 0 ? m % e.add_mod : m) * e.ld_add + n]);
  if (e.act == ACT_TANH) v = tanh_t<fast_math<T>::value>(v);
  if (e.act == ACT_RELU) v = fmaxf(v, 0.f);
  if (e.act == ACT_DTANH) {
    float y = to_f(reinterpret_cast<const T*>(e.dact_y)[(long)m * e.ld_y + n]);
    v *= (1.f - y * y);
  }
  if (e.act == ACT_DRELU) {
    float y = to_f(reinterpret_cast<const T*>(e.dact_y)[(long)m * e.ld_y + n]);
    v = y > 0.f ? v : 0.f;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ void epi_store(float v, void* C, long ldc, int m, int n, const GemmEpi& e) {
  if (e.atomic) {
    atomicAdd(reinterpret_cast<float*>(C) + (long)m * ldc + n, v);
  } else if (e.out_f32) {
    reinterpret_cast<float*>(C)[(long)m * ldc + n] = v;
  } else {
    reinterpret_cast<T*>(C)[(long)m * ldc + n] = from_f<T>(v);
  }
}

struct GemmContext {
  cudaStream_t stream = nullptr;
  int variant = GEMM_AUTO;       // GEMM_*: AUTO = cost model; SIMT routes bf16 GEMMs through the SIMT kernel (cross-check)
  long launches = 0;             // number of kernels launched (bench accounting)
};

// returns cudaError_t-style int (0 ok); message via capdec::set_error
int gemm_bf16_tc(GemmContext& ctx, const GemmCall& c);
template <typename T> int gemm_simt(GemmContext& ctx, const GemmCall& c);

template <typename T> inline int gemm(GemmContext& ctx, const GemmCall& c);
template <> inline int gemm<float>(GemmContext& ctx, const GemmCall& c) { return gemm_simt<float>(ctx, c); }
template <> inline int gemm<bf16>(GemmContext& ctx, const GemmCall& c) {
  return ctx.variant == GEMM_SIMT ? gemm_simt<bf16>(ctx, c) : gemm_bf16_tc(ctx, c);
}

void set_error(const char* fmt, ...);
// programmatic dependent launch of the tensor-core GEMMs on / off for the calling thread (off while a CUDA graph is captured)
void set_gemm_pdl(bool on);
bool gemm_pdl();

}  // namespace capdec
