// SIMT GEMM with generic operand strides: the validation-precision path (T = float) and the on-device
// cross-check for the tcgen05 kernel (T = bf16 with CAPDEC_FORCE_SIMT=1).  Not a performance path.
#include "gemm.cuh"

namespace capdec {

namespace {

constexpr int TS = 64;   // output tile
constexpr int TK = 16;

template <typename T>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const T* __restrict__ A, long sam, long sak, const T* __restrict__ B, long sbn, long sbk,
                 void* __restrict__ C, long ldc, int M, int N, int K, int kchunk, int splits, long batch_a, long batch_b,
                 long batch_c, GemmEpi epi) {
  __shared__ float sA[TK][TS + 1];
  __shared__ float sB[TK][TS + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * TS, n0 = blockIdx.x * TS;
  const int bt = blockIdx.z / splits, z = blockIdx.z % splits;
  A += bt * batch_a;
  B += bt * batch_b;
  if (epi.addend) epi.addend = reinterpret_cast<const T*>(epi.addend) + bt * epi.batch_add;
  if (epi.dact_y) epi.dact_y = reinterpret_cast<const T*>(epi.dact_y) + bt * epi.batch_y;
  C = (epi.out_f32 || epi.atomic) ? (void*)(reinterpret_cast<float*>(C) + bt * batch_c)
                                  : (void*)(reinterpret_cast<T*>(C) + bt * batch_c);
  const int k_begin = z * kchunk, k_end = min(K, k_begin + kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = k_begin; k0 < k_end; k0 += TK) {
    for (int idx = threadIdx.x; idx < TS * TK; idx += 256) {
      // pick the faster-varying index to match the contiguous operand dimension
      int r, kk;
      if (sak == 1) { kk = idx % TK; r = idx / TK; } else { r = idx % TS; kk = idx / TS; }
      int m = m0 + r, k = k0 + kk;
      sA[kk][r] = (m < M && k < k_end) ? to_f(A[(long)m * sam + (long)k * sak]) : 0.f;
      if (sbk == 1) { kk = idx % TK; r = idx / TK; } else { r = idx % TS; kk = idx / TS; }
      int n = n0 + r;
      k = k0 + kk;
      sB[kk][r] = (n < N && k < k_end) ? to_f(B[(long)n * sbn + (long)k * sbk]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float v = epi_value<T>(acc[i][j], m, n, epi);
      epi_store<T>(v, C, ldc, m, n, epi);
    }
  }
}

}  // namespace

template <typename T> int gemm_simt(GemmContext& ctx, const GemmCall& c) {
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) return 0;
  long sam, sak, sbn, sbk;
  if (!c.a_mn) { sam = c.lda; sak = 1; } else { sam = 1; sak = c.lda; }
  if (!c.b_mn) { sbn = c.ldb; sbk = 1; } else { sbn = 1; sbk = c.ldb; }
  int splits = 1;
  if (c.epi.atomic) {
    long tiles = (long)((c.M + TS - 1) / TS) * ((c.N + TS - 1) / TS);
    splits = (int)std::max(1L, std::min<long>((c.K + 255) / 256, (4 * kNumSMs) / std::max(1L, tiles)));
  }
  int kchunk = ((c.K + splits - 1) / splits + TK - 1) / TK * TK;
  splits = (c.K + kchunk - 1) / kchunk;
  dim3 grid((c.N + TS - 1) / TS, (c.M + TS - 1) / TS, splits * c.batch);
  gemm_simt_kernel<T><<<grid, 256, 0, ctx.stream>>>(reinterpret_cast<const T*>(c.A), sam, sak,
                                                   reinterpret_cast<const T*>(c.B), sbn, sbk, c.C, c.ldc, c.M, c.N,
                                                   c.K, kchunk, splits, c.batch_a, c.batch_b, c.batch_c, c.epi);
  ctx.launches++;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("gemm_simt launch: %s", cudaGetErrorString(err)); return -1; }
  return 0;
}

template int gemm_simt<float>(GemmContext&, const GemmCall&);
template int gemm_simt<bf16>(GemmContext&, const GemmCall&);

}  // namespace capdec
