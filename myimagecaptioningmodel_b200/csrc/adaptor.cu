// The two callers either side of the decoder, both HBM-bound:
//
//   feature adaptor  ImageCaptionModel._img2feature, model_adaAttention_aic.py:193-199, from the encoder map on:
//       enc_map [B, X, k] fp32 (MobileNetV2 output [B,1280,7,7] reshaped, model:195)
//         -> rows [B*k, X] (model:196 transpose) and their mean over the k positions [B, X] (model:198)
//       the two ReLU fc layers (model:197, 199) are tensor-core GEMMs with ACT_RELU / ACT_DRELU epilogues (gemm.cuh);
//       the kernels here are the layout changes around them, fused with the mean and with the fp32 <-> storage-type
//       conversions so the map is read once and written once.
//   Adam             fluid.optimizer.Adam(lr) behind opt.minimize(loss), train.py:26-31, 44-45, PaddlePaddle 1.8 rule
//       (epsilon is added to sqrt(v) un-corrected; the bias corrections scale the learning rate), with the optional
//       GradientClipByValue of train.py:41-42 fused in.  One launch over the flat parameter buffer.
#include "kernels.cuh"

namespace capdec {

void set_error(const char* fmt, ...);

namespace {

constexpr int kCh = 64;  // channels per CTA

// CTA (x-tile, image): reads the contiguous [64 channels x k positions] fp32 block of one image, writes the k rows
// (64 channels = one or two 128-byte lines each) and the 64 channel means.
// K > 0: k is the compile-time constant K and the tile is full (64 channels) -- the index arithmetic becomes
// multiply-shift and the loops unroll (all of a thread's loads are in flight together); the first version with run-time
// divisions was issue-bound (ncu: SM 47 %, DRAM 21 %).  K = 0: any k <= 64, ragged channel tiles.
template <typename T, int K>
__global__ void __launch_bounds__(256)
map_to_rows_kernel(const float* __restrict__ map, int X, int k_rt, T* __restrict__ rows, T* __restrict__ mean) {
  pdl_trigger();
  __shared__ float tile[64][kCh + 1];  // [position][channel]
  const int k = K > 0 ? K : k_rt;
  const int b = blockIdx.y, x0 = blockIdx.x * kCh;
  const int nx = K > 0 ? kCh : min(kCh, X - x0);
  const float* src = map + ((long)b * X + x0) * k;
  if (K > 0) {
    constexpr int kN = kCh * (K > 0 ? K : 1);
    constexpr int kIter = (kN + 255) / 256;
    float v[kIter];
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int i = it * 256 + threadIdx.x;
      v[it] = i < kN ? __ldcs(src + i) : 0.f;       // read once, never again
    }
#pragma unroll
    for (int it = 0; it < kIter; ++it) {
      const int i = it * 256 + threadIdx.x;
      if (i < kN) tile[i % (K > 0 ? K : 1)][i / (K > 0 ? K : 1)] = v[it];
    }
  } else {
    for (int i = threadIdx.x; i < nx * k; i += blockDim.x) tile[i % k][i / k] = src[i];
  }
  __syncthreads();
  if (threadIdx.x < nx) {
    float s = 0.f;
#pragma unroll 7
    for (int j = 0; j < k; ++j) s += tile[j][threadIdx.x];
    mean[(long)b * X + x0 + threadIdx.x] = from_f<T>(s / (float)k);   // reduce_mean(dim=1), model:198
  }
  const int groups = K > 0 ? kCh / 8 : nx / 8;  // X % 8 == 0
  for (int i = threadIdx.x; i < k * groups; i += blockDim.x) {
    const int j = i / groups, cg = i % groups;
    float v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = tile[j][cg * 8 + u];
    store8(rows + ((long)b * k + j) * X + x0 + cg * 8, v);
  }
}

// backward of the above: d map[b, x, j] = d rows[b*k + j, x] + d mean[b, x] / k
template <typename T, int K>
__global__ void __launch_bounds__(256)
rows_to_map_kernel(const T* __restrict__ drows, const float* __restrict__ dmean, int X, int k_rt, float* __restrict__ dmap) {
  pdl_trigger();
  __shared__ float tile[64][kCh + 1];
  __shared__ float dm[kCh];
  const int k = K > 0 ? K : k_rt;
  const int b = blockIdx.y, x0 = blockIdx.x * kCh;
  const int nx = K > 0 ? kCh : min(kCh, X - x0);
  if (threadIdx.x < nx) dm[threadIdx.x] = dmean[(long)b * X + x0 + threadIdx.x] / (float)k;
  const int groups = K > 0 ? kCh / 8 : nx / 8;
  for (int i = threadIdx.x; i < k * groups; i += blockDim.x) {
    const int j = i / groups, cg = i % groups;
    float v[8];
    load8(drows + ((long)b * k + j) * X + x0 + cg * 8, v);
#pragma unroll
    for (int u = 0; u < 8; ++u) tile[j][cg * 8 + u] = v[u];
  }
  __syncthreads();
  float* dst = dmap + ((long)b * X + x0) * k;
  if (K > 0) {
    constexpr int kN = kCh * (K > 0 ? K : 1);
#pragma unroll
    for (int it = 0; it < (kN + 255) / 256; ++it) {
      const int i = it * 256 + threadIdx.x;
      if (i < kN) __stcs(dst + i, tile[i % (K > 0 ? K : 1)][i / (K > 0 ? K : 1)] + dm[i / (K > 0 ? K : 1)]);
    }
  } else {
    for (int i = threadIdx.x; i < nx * k; i += blockDim.x) dst[i] = tile[i % k][i / k] + dm[i / k];
  }
}

template <typename T> __global__ void to_f32_kernel(const T* __restrict__ src, float* __restrict__ dst, long n8) {
  pdl_trigger();
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n8; i += stride) {
    float v[8];
    load8(src + i * 8, v);
    store8(dst + i * 8, v);
  }
}

// PaddlePaddle 1.8 adam_op (AdamFunctor):  lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t)
//   m = beta1 m + (1 - beta1) g;  v = beta2 v + (1 - beta2) g^2;  p -= lr_t * m / (sqrt(v) + epsilon)
struct AdamArgs {
  float lr_t, beta1, beta2, eps, clip;
};
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a) {
  if (a.clip > 0.f) g = fminf(fmaxf(g, -a.clip), a.clip);  // GradientClipByValue(max = clip, min = -clip)
  m = a.beta1 * m + (1.f - a.beta1) * g;
  v = a.beta2 * v + (1.f - a.beta2) * g * g;
  p -= a.lr_t * (m / (sqrtf(v) + a.eps));
}
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
            AdamArgs a) {
  const long n4 = n >> 2;
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    const float4 G = __ldcs(reinterpret_cast<const float4*>(g) + i);   // the gradient is not read again
    float4 M = reinterpret_cast<float4*>(m)[i], V = reinterpret_cast<float4*>(v)[i];
    adam_one(P.x, G.x, M.x, V.x, a);
    adam_one(P.y, G.y, M.y, V.y, a);
    adam_one(P.z, G.z, M.z, V.z, a);
    adam_one(P.w, G.w, M.w, V.w, a);
    reinterpret_cast<float4*>(p)[i] = P;
    __stcs(reinterpret_cast<float4*>(m) + i, M);
    __stcs(reinterpret_cast<float4*>(v) + i, V);
  }
  // ragged tail (n % 4 elements)
  const long t = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) adam_one(p[t], g[t], m[t], v[t], a);
}
__global__ void adam_scalar_kernel(float* p, const float* g, float* m, float* v, long n, AdamArgs a) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long stride = (long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) adam_one(p[i], g[i], m[i], v[i], a);
}

}  // namespace

#define BUMP(c) do { if ((c).launches) ++*(c).launches; } while (0)

template <typename T> int map_to_rows(LaunchCtx c, const float* map, int B, int X, int k, T* rows, T* mean) {
  if (X % 8 || k < 1 || k > 64) { set_error("feature adaptor: need encoder_channel %% 8 == 0 and 1 <= k <= 64"); return -1; }
  dim3 grid((X + kCh - 1) / kCh, B);
  if (k == 49 && X % kCh == 0) map_to_rows_kernel<T, 49><<<grid, 256, 0, c.stream>>>(map, X, k, rows, mean);   // 7 x 7 encoder map
  else map_to_rows_kernel<T, 0><<<grid, 256, 0, c.stream>>>(map, X, k, rows, mean);
  BUMP(c);
  return 0;
}
template <typename T> int rows_to_map(LaunchCtx c, const T* drows, const float* dmean, int B, int X, int k, float* dmap) {
  if (X % 8 || k < 1 || k > 64) { set_error("feature adaptor: need encoder_channel %% 8 == 0 and 1 <= k <= 64"); return -1; }
  dim3 grid((X + kCh - 1) / kCh, B);
  if (k == 49 && X % kCh == 0) rows_to_map_kernel<T, 49><<<grid, 256, 0, c.stream>>>(drows, dmean, X, k, dmap);
  else rows_to_map_kernel<T, 0><<<grid, 256, 0, c.stream>>>(drows, dmean, X, k, dmap);
  BUMP(c);
  return 0;
}
template <typename T> void convert_to_f32(LaunchCtx c, const T* src, float* dst, long n) {
  if (n <= 0) return;
  const long n8 = n / 8;  // callers pass multiples of 8 (H % 8 == 0)
  long blocks = (n8 + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;
  to_f32_kernel<T><<<(int)blocks, 256, 0, c.stream>>>(src, dst, n8);
  BUMP(c);
}

int adam_step(cudaStream_t s, float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
              float eps, float beta1_pow, float beta2_pow, float clip) {
  if (n <= 0) return 0;
  AdamArgs a;
  a.lr_t = lr * sqrtf(1.f - beta2_pow) / (1.f - beta1_pow);
  a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.clip = clip;
  const bool vec = (((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) % 16) == 0;
  long work = vec ? (n + 3) / 4 : n;
  long blocks = (work + 255) / 256;
  if (blocks > kNumSMs * 8) blocks = kNumSMs * 8;   // 8 CTAs of 256 threads per SM, grid-stride
  if (vec) adam_kernel<<<(int)blocks, 256, 0, s>>>(p, g, m, v, n, a);
  else adam_scalar_kernel<<<(int)blocks, 256, 0, s>>>(p, g, m, v, n, a);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

template int map_to_rows<float>(LaunchCtx, const float*, int, int, int, float*, float*);
template int map_to_rows<bf16>(LaunchCtx, const float*, int, int, int, bf16*, bf16*);
template int rows_to_map<float>(LaunchCtx, const float*, const float*, int, int, int, float*);
template int rows_to_map<bf16>(LaunchCtx, const bf16*, const float*, int, int, int, float*);
template void convert_to_f32<float>(LaunchCtx, const float*, float*, long);
template void convert_to_f32<bf16>(LaunchCtx, const bf16*, float*, long);

}  // namespace capdec
