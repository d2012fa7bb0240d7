// Fused additive (adaptive, visual-sentinel) attention -- model_adaAttention_aic.py:103-113 -- forward and backward.
//
//   z_j   = tanh(key_j + hid_emb)            j < k : per-image keys;  j = k : sentinel_emb (per query)
//   e_j   = z_j . w_alpha + b_alpha
//   alpha = softmax over the k+1 positions   ('position'), or == 1 as the reference graph is written ('feature')
//   ctx   = 1/(k+1) * sum_j alpha_j v_j      v_j: per-image values; v_k = sentinel          (reduce_mean, model:113)
//
// The reference materialises seven [B,k+1,H] tensors per step; here nothing of that size is written.
// Forward: one CTA per (image, chunk of query rows).  The image's key and value tiles ([k,H] each, time-invariant)
// are brought into shared memory ONCE per CTA with two 1-D TMA bulk copies (cp.async.bulk, 2 x 98 KB for bf16) and
// every query row of the chunk (34 teacher-forced steps, or the beams of one image at inference) streams through.
// Scores use one warp per position with 128-bit conflict-free shared loads and a shuffle reduction; the context
// pass gives each thread 4 adjacent columns.
#include <cstdlib>

#include "gemm.cuh"
#include "kernels.cuh"

namespace capdec {

namespace {

constexpr int kFwdThreads = 320;  // 10 warps: 50 positions = 5 rounds

__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
  float2 a = __bfloat1622float2(h[0]), b = __bfloat1622float2(h[1]);
  v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  uint2 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
  h[0] = __floats2bfloat162_rn(v[0], v[1]);
  h[1] = __floats2bfloat162_rn(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = u;
}

// Row addressing of query q of image img.  Padded row (alpha, alpha_T, de, ...): img * row0_stride + q * q_stride; the row
// of the per-query [rows, H] operands: the same, or with packed rows (AttnShape::prow) prow[padded row].
__device__ __forceinline__ long padded_row(const AttnShape& s, int img, int q) {
  return (long)img * s.row0_stride + (long)q * s.q_stride;
}
__device__ __forceinline__ long operand_row(const AttnShape& s, long prow_idx) {
  if (!s.prow) return prow_idx;
  const int r = s.prow[prow_idx];
  return r < 0 ? 0 : r;   // -1 only when `lengths` was wrong (reported, CAPDEC_ERR_LENGTHS): stay inside the buffers
}
__device__ __forceinline__ int live_queries(const AttnShape& s, int img) {
  return s.nq ? min(s.nq[img], s.q_per_img) : s.q_per_img;
}

// shared-memory carve-up shared by the forward and the backward-score kernels
struct SmemLayout {
  int off_e, off_alpha, off_bar, off_q, off_tile0, off_tile1, total;
};
// q_bytes: bytes reserved for the per-query vector ring of the forward kernel (0 for the backward kernels)
__host__ __device__ inline SmemLayout make_layout(int H, int k, int tile_elem_bytes, int n_tiles, int q_bytes = 0) {
  SmemLayout L;
  L.off_e = 0;            // 2 x 128 floats (double-buffered scores)
  L.off_alpha = 256 * 4;  // 10 warps x 64 floats (per-warp alpha copies)
  L.off_bar = 256 * 4 + 2560;
  L.off_q = 256 * 4 + 2560 + 64;
  L.off_tile0 = L.off_q + q_bytes;
  const int tile = ((k * H * tile_elem_bytes + 127) / 128) * 128;
  L.off_tile1 = L.off_tile0 + (n_tiles > 0 ? tile : 0);
  L.total = L.off_tile0 + n_tiles * tile;
  return L;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// sum_h tanh(key[h] + he[h]) * w[h] over this lane's 8-element slices (h = v*256 + lane*8 + i)
//   bf16: packed -- one HADD2 and ONE MUFU (tanh.approx.bf16x2) per two elements, fp32 accumulation
//   fp32: scalar tanhf (validation precision)
template <int NV>
__device__ __forceinline__ float score_dot(const bf16* kj, const uint32_t (&he2)[NV][4], const float (&w)[NV][8], int lane) {
  // four independent accumulators (one per 16-byte sub-chunk): no serial FFMA chain behind the MUFU results
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  uint4 kq[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) kq[v] = *reinterpret_cast<const uint4*>(kj + v * 256 + lane * 8);
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint32_t kk[4] = {kq[v].x, kq[v].y, kq[v].z, kq[v].w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t s2, z2;
      asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(s2) : "r"(kk[p]), "r"(he2[v][p]));
      asm("tanh.approx.bf16x2 %0, %1;" : "=r"(z2) : "r"(s2));
      acc[p] = fmaf(__uint_as_float(z2 << 16), w[v][2 * p], acc[p]);
      acc[p] = fmaf(__uint_as_float(z2 & 0xffff0000u), w[v][2 * p + 1], acc[p]);
    }
  }
  return (acc[0] + acc[1]) + (acc[2] + acc[3]);
}
template <int NV>
__device__ __forceinline__ float score_dot(const float* kj, const float (&he)[NV][8], const float (&w)[NV][8], int lane) {
  float acc = 0.f;
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    float kk[8];
    load8(kj + v * 256 + lane * 8, kk);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(tanhf(kk[i] + he[v][i]), w[v][i], acc);
  }
  return acc;
}
// per-lane query slice in the form score_dot wants
template <int NV> __device__ __forceinline__ void load_he(const bf16* he_s, uint32_t (&he2)[NV][4], int lane) {
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const uint4 q = *reinterpret_cast<const uint4*>(he_s + v * 256 + lane * 8);
    he2[v][0] = q.x; he2[v][1] = q.y; he2[v][2] = q.z; he2[v][3] = q.w;
  }
}
template <int NV> __device__ __forceinline__ void load_he(const float* he_s, float (&he)[NV][8], int lane) {
#pragma unroll
  for (int v = 0; v < NV; ++v) load8(he_s + v * 256 + lane * 8, he[v]);
}
template <typename T, int NV> struct HeRegs;
template <int NV> struct HeRegs<bf16, NV> { uint32_t r[NV][4]; };
template <int NV> struct HeRegs<float, NV> { float r[NV][8]; };

// NV > 0: H == NV * 256, per-lane operands live in registers.  NV == 0: generic H (multiple of 8).
//
// Per query row: (1) scores, one warp per position, e -> smem (double-buffered by query parity);
// (2) ONE block barrier; (3) every warp recomputes the (k+1)-way soft-max from smem with shuffles (no second
// barrier, no idle warps), warp 0 writes alpha; (4) context, 4 adjacent columns per thread, alpha broadcast by
// shuffle.  The query's four H-vectors (hid_emb, sentinel_emb, sentinel, p_hid) arrive through a 3-deep cp.async
// ring, prefetched one query ahead, so no global-load latency is exposed inside the loop.
template <typename T, int NV, bool STAGE>
__global__ void __launch_bounds__(kFwdThreads, 1)
attn_fwd_kernel(AttnShape s, int chunks, int q_per_chunk, const T* __restrict__ keys, const T* __restrict__ vals,
                const T* __restrict__ hid_emb, const T* __restrict__ sent_emb, const T* __restrict__ sentinel,
                const T* __restrict__ p_hid, const float* __restrict__ w_alpha, const float* __restrict__ b_alpha,
                float* __restrict__ alpha_out, T* __restrict__ actx, T* __restrict__ ctx_out) {
  constexpr bool F = fast_math<T>::value;
  constexpr int RV = NV > 0 ? NV : 1;
  extern __shared__ __align__(128) uint8_t smem[];
  const int H = s.H, k = s.k, P = s.k + 1;
  const int q_bytes = 3 * 4 * H * (int)sizeof(T);
  const SmemLayout L = make_layout(H, k, (int)sizeof(T), STAGE ? 2 : 0, q_bytes);
  float* e_s = reinterpret_cast<float*>(smem + L.off_e);
  float* a_s = reinterpret_cast<float*>(smem + L.off_alpha);  // [nwarps][64] per-warp alpha copies
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  T* qring = reinterpret_cast<T*>(smem + L.off_q);  // [3][4][H]: hid_emb, sent_emb, sentinel, p_hid
  const int img = blockIdx.x / chunks, chunk = blockIdx.x % chunks;
  const int q0 = chunk * q_per_chunk, q1 = min(live_queries(s, img), q0 + q_per_chunk);
  if (q0 >= q1) return;   // no live query in this chunk (packed rows)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  const T* Kt = keys + (long)img * k * H;
  const T* Vt = vals + (long)img * k * H;
  if (STAGE) {
    const uint32_t tile_bytes = (uint32_t)(k * H * sizeof(T));
    if (tid == 0) {
      mbar_init(bar, 1);
      mbar_fence_init();
      mbar_expect_tx(bar, 2 * tile_bytes);
      tma_bulk_g2s(smem + L.off_tile0, Kt, tile_bytes, bar);
      tma_bulk_g2s(smem + L.off_tile1, Vt, tile_bytes, bar);
    }
    Kt = reinterpret_cast<const T*>(smem + L.off_tile0);
    Vt = reinterpret_cast<const T*>(smem + L.off_tile1);
  }
  const int vec_per_row = H / 8;  // 16-byte chunks of T=bf16 (or half-chunks of fp32) per H-vector
  auto prefetch = [&](int q) {
    const long row = operand_row(s, padded_row(s, img, q));
    T* dst = qring + (size_t)(q % 3) * 4 * H;
    const T* src[4] = {hid_emb + row * H, sent_emb + row * H, sentinel + row * H, p_hid ? p_hid + row * H : sentinel + row * H};
    constexpr int EPC = 16 / (int)sizeof(T);  // elements per 16-byte cp.async
    const int n16 = H / EPC;
    for (int i = tid; i < 4 * n16; i += blockDim.x) {
      const int a = i / n16, c = i % n16;
      cp_async16(dst + a * H + c * EPC, src[a] + c * EPC);
    }
    cp_async_commit();
  };
  (void)vec_per_row;
  if (q0 < q1) prefetch(q0);
  // per-lane slice of w_alpha
  float w[RV][8];
  if (NV > 0) {
#pragma unroll
    for (int v = 0; v < RV; ++v) load8(w_alpha + v * 256 + lane * 8, w[v]);
  }
  const float b_a = b_alpha[0];
  const float inv_p = 1.f / (float)P;
  cp_async_wait_all();
  __syncthreads();  // barrier init + first query vectors visible
  if (STAGE) mbar_wait(bar, 0);
  for (int q = q0; q < q1; ++q) {
    const long row = padded_row(s, img, q);         // alpha
    const long orow = operand_row(s, row);          // actx / ctx_out (per-query [rows, H] operands)
    const T* qv = qring + (size_t)(q % 3) * 4 * H;
    const T* he_s = qv;
    const T* se_s = qv + H;
    const T* sn_s = qv + 2 * H;
    const T* ph_s = qv + 3 * H;
    float* eq = e_s + (q & 1) * 128;
    if (q + 1 < q1) prefetch(q + 1);  // ring slot (q+1)%3 was last read in query q-2: everyone is past it
    // ---- scores: one warp per position ----
    HeRegs<T, RV> he;
    if (NV > 0) load_he<RV>(he_s, he.r, lane);
    for (int j = warp; j < P; j += nwarps) {
      const T* kj = (j < k) ? (Kt + (long)j * H) : se_s;
      float acc = 0.f;
      if (NV > 0) {
        acc = score_dot<RV>(kj, he.r, w, lane);
      } else {
        for (int h0 = lane * 8; h0 < H; h0 += 256) {
          float kk[8], hh[8], ww[8];
          load8(kj + h0, kk);
          load8(he_s + h0, hh);
          load8(w_alpha + h0, ww);
#pragma unroll
          for (int i = 0; i < 8; ++i) acc = fmaf(tanh_t<F>(kk[i] + hh[i]), ww[i], acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) eq[j] = acc + b_a;
    }
    cp_async_wait_all();  // next query's vectors have landed (this thread's part); the barrier publishes them
    __syncthreads();
    // ---- soft-max over the k+1 positions, recomputed by every warp (P <= 64: two values per lane) ----
    float a_lo, a_hi;
    {
      const float e0 = (lane < P) ? eq[lane] : -INFINITY;
      const float e1 = (lane + 32 < P) ? eq[lane + 32] : -INFINITY;
      if (s.position_softmax) {
        const float m = warp_max(fmaxf(e0, e1));
        const float x0 = (lane < P) ? __expf(e0 - m) : 0.f;
        const float x1 = (lane + 32 < P) ? __expf(e1 - m) : 0.f;
        const float inv = 1.f / warp_sum(x0 + x1);
        a_lo = x0 * inv;
        a_hi = x1 * inv;
      } else {
        a_lo = 1.f;  // softmax over a size-1 axis (model:107 as written)
        a_hi = 1.f;
      }
      if (warp == 0) {
        if (lane < P) alpha_out[row * P + lane] = a_lo;
        if (lane + 32 < P) alpha_out[row * P + lane + 32] = a_hi;
      }
    }
    // ---- context: each thread owns 4 adjacent columns; alpha (already scaled by 1/(k+1)) sits in smem ----
    // written redundantly (same values) by every warp for its own use -> no extra barrier; positions are processed
    // 7 at a time with all loads issued first so that smem latency is paid once per group, not once per position
    float* aw = a_s + warp * 64;
    if (lane < P) aw[lane] = a_lo * inv_p;
    if (lane + 32 < P) aw[lane + 32] = a_hi * inv_p;
    __syncwarp();
    for (int hb = 0; hb < H; hb += blockDim.x * 4) {
      const int h0 = hb + tid * 4;
      if (h0 < H) {
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const T* vp = Vt + h0;
        int j = 0;
        for (; j + 7 <= k; j += 7) {
          float v[7][4], a[7];
#pragma unroll
          for (int u = 0; u < 7; ++u) {
            load4(vp + (long)(j + u) * H, v[u]);
            a[u] = aw[j + u];
          }
#pragma unroll
          for (int u = 0; u < 7; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(a[u], v[u][i], acc[i]);
        }
        for (; j < k; ++j) {
          float v[4];
          load4(vp + (long)j * H, v);
          const float a = aw[j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(a, v[i], acc[i]);
        }
        {
          float v[4];
          load4(sn_s + h0, v);
          const float a = aw[k];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] = fmaf(a, v[i], acc[i]);
        }
        if (ctx_out) store4(ctx_out + orow * H + h0, acc);
        if (actx) {
          float ph[4];
          load4(ph_s + h0, ph);
#pragma unroll
          for (int i = 0; i < 4; ++i) ph[i] += acc[i];
          store4(actx + orow * H + h0, ph);
        }
      }
    }
    __syncwarp();  // aw is rewritten by this warp in the next query
    // no trailing barrier: the next query writes the other score buffer and a ring slot nobody reads any more
  }
}

// ---- forward, training shape (many query rows per image): scores + soft-max kernel ---------------------------
// The score e[q,j] = sum_h w[h] tanh(key[j,h] + hid_emb[q,h]) is MUFU work (one tanh per element, 445 M per train
// step) wrapped in a dot product.  Here the tanh runs packed (tanh.approx.bf16x2, two elements per MUFU op) and its
// bf16x2 results ARE the A fragment of a warp-level tensor-core MMA (mma.sync m16n8k16: 16 positions x 16 columns of z
// against a B fragment that holds w in column 0), so the multiply-accumulate with w costs 1/4 instruction per pair
// instead of two unpacks + two FFMAs, and the kernel is MUFU-bound instead of issue-bound.
//   CTA = one image: key tile [k, H] in shared memory (row pitch 2H + 64 B: a quarter-warp's 2 rows x 4 chunks of
//   16 B hit all 32 banks once), 2 CTAs per SM.  Warp w owns columns [w*H/8, (w+1)*H/8) for every position and
//   query, w fragments resident in registers; the 8 partial sums of a score meet in shared memory, one barrier per
//   query, and one warp (rotating) finishes the soft-max while the others start the next query.
//   Position k (the sentinel) reads its "key" row from the staged sentinel_emb row of the query.
constexpr int kScThreads = 256;
constexpr int kScPad = 64;

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t tanh_add_bf16x2(uint32_t a, uint32_t b) {
  uint32_t s2, z2;
  asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(s2) : "r"(a), "r"(b));
  asm("tanh.approx.bf16x2 %0, %1;" : "=r"(z2) : "r"(s2));
  return z2;
}

inline size_t scores_smem_bytes(int H, int k) {
  const int pitch = 2 * H + kScPad, Ppad = (k + 1 + 3) & ~3;
  return (size_t)k * pitch + 4 * (size_t)pitch + 2 * 8 * Ppad * sizeof(float) + 64;
}

template <int NB>  // H = 256 * NB
__global__ void __launch_bounds__(kScThreads, 2)
attn_scores_kernel(AttnShape s, const bf16* __restrict__ keys, const bf16* __restrict__ hid_emb,
                   const bf16* __restrict__ sent_emb, const float* __restrict__ w_alpha,
                   const float* __restrict__ b_alpha, float* __restrict__ alpha_out, bf16* __restrict__ alpha_T,
                   int ld_at, int q_per_chunk) {
  pdl_trigger();
  // CTA = (image, chunk of its live queries): the block scheduler balances images of different caption lengths (packed
  // rows) over the SMs; a chunk past the image's last live query exits before it touches anything
  const int q_begin = blockIdx.y * q_per_chunk;
  const int nq = min(live_queries(s, blockIdx.x), q_begin + q_per_chunk);
  if (q_begin >= nq) return;
  constexpr int H = 256 * NB;
  constexpr int kWarpCols = 32 * NB;  // columns of one warp
  extern __shared__ __align__(128) uint8_t smem[];
  const int k = s.k, P = s.k + 1, Ppad = (P + 3) & ~3;
  const int pitch = 2 * H + kScPad;
  uint8_t* Kt = smem;
  uint8_t* ring = Kt + (size_t)k * pitch;                      // [2 slots][hid_emb row, sentinel_emb row]
  float* e_part = reinterpret_cast<float*>(ring + 4 * pitch);  // [2][8 warps][Ppad]
  uint64_t* bar = reinterpret_cast<uint64_t*>(e_part + 2 * 8 * Ppad);
  const int img = blockIdx.x;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int lr = lane >> 2, lc = lane & 3;  // fragment row / 16-byte column chunk of this lane
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_fence_init();
    mbar_expect_tx(bar, (uint32_t)(k * 2 * H));
    const bf16* src = keys + (long)img * k * H;
    for (int j = 0; j < k; ++j) tma_bulk_g2s(Kt + (size_t)j * pitch, src + (long)j * H, 2 * H, bar);
  }
  constexpr int cpr = H / 8;  // 16-byte chunks per row
  auto prefetch = [&](int q) {
    if (tid < 2 * cpr) {
      const long row = operand_row(s, padded_row(s, img, q));
      const int a = tid / cpr, c = tid % cpr;
      cp_async16(ring + (size_t)((q & 1) * 2 + a) * pitch + c * 16, (a == 0 ? hid_emb : sent_emb) + row * H + c * 8);
    }
    cp_async_commit();
  };
  prefetch(q_begin);
  // B fragments: w in column 0 (fragment lanes 0..3), zero elsewhere
  const int col0 = warp * kWarpCols + 8 * lc;  // first column of this lane's chunk in block 0
  uint32_t wq[NB][4];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    float wv[8];
    load8(w_alpha + col0 + 32 * b, wv);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h2 = __floats2bfloat162_rn(wv[2 * i], wv[2 * i + 1]);
      wq[b][i] = (lr == 0) ? *reinterpret_cast<uint32_t*>(&h2) : 0u;
    }
  }
  const float b_a = b_alpha[0];
  const float inv_p = 1.f / (float)P;
  cp_async_wait_all();
  __syncthreads();
  mbar_wait(bar, 0);
  for (int q = q_begin; q < nq; ++q) {
    if (q + 1 < nq) prefetch(q + 1);  // slot (q+1)&1 was last read in query q-1: everyone is past its barrier
    const uint8_t* he_row = ring + (size_t)((q & 1) * 2) * pitch;
    const uint8_t* se_row = he_row + pitch;
    uint4 heq[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) heq[b] = *reinterpret_cast<const uint4*>(he_row + (col0 + 32 * b) * 2);
    float* ep = e_part + ((q & 1) * 8 + warp) * Ppad;
    for (int pb = 0; 16 * pb <= k; ++pb) {
      const int jA = 16 * pb + lr, jB = jA + 8;
      const bool half_b = (16 * pb + 8 <= k);  // warp-uniform: the upper 8 rows hold at least one valid position
      const uint8_t* rowA = (jA < k) ? Kt + (size_t)jA * pitch : se_row;
      const uint8_t* rowB = (jB < k) ? Kt + (size_t)jB * pitch : se_row;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const uint4 ka = *reinterpret_cast<const uint4*>(rowA + (col0 + 32 * b) * 2);
        uint32_t za[4], zb[4] = {0u, 0u, 0u, 0u};
        za[0] = tanh_add_bf16x2(ka.x, heq[b].x);
        za[1] = tanh_add_bf16x2(ka.y, heq[b].y);
        za[2] = tanh_add_bf16x2(ka.z, heq[b].z);
        za[3] = tanh_add_bf16x2(ka.w, heq[b].w);
        if (half_b) {
          const uint4 kb = *reinterpret_cast<const uint4*>(rowB + (col0 + 32 * b) * 2);
          zb[0] = tanh_add_bf16x2(kb.x, heq[b].x);
          zb[1] = tanh_add_bf16x2(kb.y, heq[b].y);
          zb[2] = tanh_add_bf16x2(kb.z, heq[b].z);
          zb[3] = tanh_add_bf16x2(kb.w, heq[b].w);
        }
        // k index 2*lc + {0,1} <-> chunk elements {0,1}; k index 8 + 2*lc + {0,1} <-> elements {2,3}; then {4..7}
        mma_bf16_16816(c, za[0], zb[0], za[1], zb[1], wq[b][0], wq[b][1]);
        mma_bf16_16816(c, za[2], zb[2], za[3], zb[3], wq[b][2], wq[b][3]);
      }
      if (lc == 0) {
        if (jA <= k) ep[jA] = c[0];
        if (jB <= k) ep[jB] = c[2];
      }
    }
    cp_async_wait_all();  // next query's rows have landed (this thread's part); the barrier publishes them
    __syncthreads();
    if (warp == (q & 7)) {
      // ---- soft-max over the k+1 positions (P <= 64: two values per lane) ----
      const long row = padded_row(s, img, q);
      const float* eb = e_part + (q & 1) * 8 * Ppad;
      float e0 = -INFINITY, e1 = -INFINITY;
      if (lane < P) {
        e0 = b_a;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) e0 += eb[w8 * Ppad + lane];
      }
      if (lane + 32 < P) {
        e1 = b_a;
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) e1 += eb[w8 * Ppad + lane + 32];
      }
      const float m = warp_max(fmaxf(e0, e1));
      const float x0 = (lane < P) ? __expf(e0 - m) : 0.f;
      const float x1 = (lane + 32 < P) ? __expf(e1 - m) : 0.f;
      const float inv = 1.f / warp_sum(x0 + x1);
      const float a0 = x0 * inv, a1 = x1 * inv;
      if (lane < P) alpha_out[row * P + lane] = a0;
      if (lane + 32 < P) alpha_out[row * P + lane + 32] = a1;
      if (alpha_T) {
        if (lane < k) alpha_T[row * ld_at + lane] = __float2bfloat16_rn(a0 * inv_p);
        if (lane + 32 < k) alpha_T[row * ld_at + lane + 32] = __float2bfloat16_rn(a1 * inv_p);
      }
    }
  }
}

// sadd[row, :] = p_hid[row, :] + alpha[row, k] / (k+1) * sentinel[row, :]: the two per-query terms the context GEMM
// adds in its epilogue (p_hid may be NULL)
// pm (packed rows): operand row r is padded row pm[r] of alpha / sadd; dead rows of sadd are not written (the context rows
// they feed are never read)
__global__ void __launch_bounds__(256)
attn_sadd_kernel(long rows, int H, int P, const int* __restrict__ pm, const float* __restrict__ alpha,
                 const bf16* __restrict__ sentinel, const bf16* __restrict__ p_hid, bf16* __restrict__ sadd) {
  pdl_trigger();
  const int vec_per_row = H / 8;
  const float inv_p = 1.f / (float)P;
  const long n = rows * vec_per_row;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long r = i / vec_per_row;
    const int h0 = (int)(i % vec_per_row) * 8;
    const long m = pm ? pm[r] : r;
    const float a = alpha[m * P + (P - 1)] * inv_p;
    float v[8], ph[8];
    load8(sentinel + r * H + h0, v);
    if (p_hid) {
      load8(p_hid + r * H + h0, ph);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = fmaf(a, v[j], ph[j]);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] *= a;
    }
    store8(sadd + m * H + h0, v);
  }
}

// context for a few query rows per image (inference: the beams of one image):
//   actx[q, :] = 1/(k+1) * (sum_j alpha[q,j] v_j + alpha[q,k] sentinel[q]) + p_hid[q]
// one CTA per image, 4 adjacent columns per thread, the value rows stream from global memory once (coalesced 8-byte
// loads, all k of them independent) and are reused for every query row of the image.
template <int QMAX>
__global__ void __launch_bounds__(256)
attn_ctx_small_kernel(AttnShape s, const bf16* __restrict__ vals, const bf16* __restrict__ sentinel,
                      const bf16* __restrict__ p_hid, const float* __restrict__ alpha, bf16* __restrict__ actx) {
  pdl_trigger();
  extern __shared__ __align__(16) float aw[];  // [q_per_img][P], already scaled by 1/(k+1)
  const int H = s.H, k = s.k, P = s.k + 1, nq = s.q_per_img;
  const int img = blockIdx.x;
  const float inv_p = 1.f / (float)P;
  for (int i = threadIdx.x; i < nq * P; i += blockDim.x) {
    const int q = i / P, j = i % P;
    aw[i] = alpha[((long)img * s.row0_stride + (long)q * s.q_stride) * P + j] * inv_p;
  }
  __syncthreads();
  const bf16* Vt = vals + (long)img * k * H;
  for (int h0 = threadIdx.x * 4; h0 < H; h0 += blockDim.x * 4) {
    float acc[QMAX][4];
#pragma unroll
    for (int q = 0; q < QMAX; ++q)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[q][i] = 0.f;
#pragma unroll 7
    for (int j = 0; j < k; ++j) {
      float v[4];
      load4(Vt + (long)j * H + h0, v);
#pragma unroll
      for (int q = 0; q < QMAX; ++q) {
        if (q < nq) {
          const float a = aw[q * P + j];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[q][i] = fmaf(a, v[i], acc[q][i]);
        }
      }
    }
#pragma unroll
    for (int q = 0; q < QMAX; ++q) {
      if (q < nq) {
        const long row = (long)img * s.row0_stride + (long)q * s.q_stride;
        float sv[4], ph[4] = {0.f, 0.f, 0.f, 0.f};
        load4(sentinel + row * H + h0, sv);
        if (p_hid) load4(p_hid + row * H + h0, ph);
        const float a = aw[q * P + k];
#pragma unroll
        for (int i = 0; i < 4; ++i) ph[i] += fmaf(a, sv[i], acc[q][i]);
        store4(actx + row * H + h0, ph);
      }
    }
  }
}

// ---- backward, row-wise part (one warp per query row) -------------------------------------------------------
//   d alpha_j (j < k) arrives UNSCALED from the batched tensor-core GEMM  dal[row, j] = dctx[row, :] . v_j ;
//   here: d alpha_k = dctx . sentinel,  de = alpha * (d alpha - sum alpha d alpha) / (k+1)  (soft-max backward),
//   db_alpha += sum de,  dS_ctx = alpha_k / (k+1) * dctx,  and the storage-type copy alpha_T = alpha / (k+1)
//   that the second batched GEMM (d values) uses as its A operand.
//   pm (packed rows): the loop runs over the live rows; `orow` indexes sentinel / dS, row = pm[orow] everything else.
template <typename T>
__global__ void __launch_bounds__(256)
attn_bwd_rows_kernel(long rows, int H, int k, int position_softmax, const int* __restrict__ pm,
                     const T* __restrict__ sentinel, const float* __restrict__ alpha, const T* __restrict__ dactx,
                     const float* __restrict__ dal, int ld_dal, float* __restrict__ de, T* __restrict__ alpha_T, int ld_at,
                     T* __restrict__ dS, float* __restrict__ db_alpha) {
  pdl_trigger();
  const int lane = threadIdx.x & 31;
  const int P = k + 1;
  const float inv_p = 1.f / (float)P;
  float db_acc = 0.f;
  for (long orow = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); orow < rows;
       orow += (long)gridDim.x * (blockDim.x >> 5)) {
    const long row = pm ? pm[orow] : orow;
    const T* dc = dactx + row * H;
    const float a0 = (lane < P) ? alpha[row * P + lane] : 0.f;
    const float a1 = (lane + 32 < P) ? alpha[row * P + lane + 32] : 0.f;
    if (lane < k) alpha_T[row * ld_at + lane] = from_f<T>(a0 * inv_p);
    if (lane + 32 < k) alpha_T[row * ld_at + lane + 32] = from_f<T>(a1 * inv_p);
    const float ak = __shfl_sync(0xffffffffu, (k < 32) ? a0 : a1, k & 31);
    float dot_s = 0.f;
    for (int h0 = lane * 8; h0 < H; h0 += 256) {
      float d[8], sv[8];
      load8(dc + h0, d);
      if (position_softmax) {
        load8(sentinel + orow * H + h0, sv);
#pragma unroll
        for (int i = 0; i < 8; ++i) dot_s = fmaf(d[i], sv[i], dot_s);
      }
      const float sc = ak * inv_p;
#pragma unroll
      for (int i = 0; i < 8; ++i) d[i] *= sc;
      store8(dS + orow * H + h0, d);
    }
    if (position_softmax) {
      dot_s = warp_sum(dot_s);
      float d0 = 0.f, d1 = 0.f;  // d alpha for positions lane, lane + 32
      if (lane < k) d0 = dal[row * ld_dal + lane];
      if (lane + 32 < k) d1 = dal[row * ld_dal + lane + 32];
      if (lane == k) d0 = dot_s;
      if (lane + 32 == k) d1 = dot_s;
      d0 *= inv_p;
      d1 *= inv_p;
      const float dot = warp_sum(a0 * d0 + a1 * d1);
      const float e0 = a0 * (d0 - dot), e1 = a1 * (d1 - dot);
      if (lane < P) de[row * P + lane] = e0;
      if (lane + 32 < P) de[row * P + lane + 32] = e1;
      db_acc += e0 + e1;
    }
  }
  if (position_softmax) {
    db_acc = warp_sum(db_acc);
    if (lane == 0) atomicAdd(db_alpha, db_acc);
  }
}

// ---- backward (C): d keys, d hid_emb, d sentinel_emb, d w_alpha ----
//   u_j = 1 - z_j^2,  z_j = tanh(key_j[h] + hid_emb[q,h])  (recomputed, never stored; z_k uses sentinel_emb[q,h])
//   dK_j[h] = w[h] sum_q de[q,j] u_j;   d hid_emb[q,h] = w[h] sum_j de[q,j] u_j;   d sentinel_emb[q,h] = w[h] de[q,k] u_k;
//   dw[h] += sum_{q,j} de[q,j] z_j
// Two adjacent lanes share one column h of one image and split the k+1 positions (KH each, the sentinel is the last
// position of the upper half), so a thread keeps KH keys + KH accumulators in registers (3 CTAs per SM instead of 2
// with one thread per column) and the d hid_emb halves meet through one shuffle per query.  Per element: one add,
// one MUFU tanh and four FFMAs (w[h] is applied once per output, not per element); de comes from shared memory in
// 16-byte broadcasts.
template <typename T, int KH>
__global__ void __launch_bounds__(256, 3)
attn_bwd_dk_kernel(AttnShape s, const T* __restrict__ keys, const T* __restrict__ hid_emb,
                   const T* __restrict__ sent_emb, const float* __restrict__ w_alpha, const float* __restrict__ de,
                   T* __restrict__ dKV, T* __restrict__ dhid_emb, T* __restrict__ dsent_emb,
                   float* __restrict__ dw_alpha) {
  pdl_trigger();
  constexpr bool F = fast_math<T>::value;
  constexpr int KHP = (KH + 3) & ~3;  // each half of a query's de row starts 16-byte aligned
  extern __shared__ __align__(16) float de_s[];  // [q_per_img][2 halves][KHP], zero past the valid positions
  const int H = s.H, k = s.k, P = s.k + 1;
  const int img = blockIdx.x;
  const int nq = live_queries(s, img);
  // operand rows of this image's queries (packed rows: irregular), behind the de tile
  int* rows_s = reinterpret_cast<int*>(de_s + (size_t)s.q_per_img * 2 * (((KH + 3) & ~3)));
  for (int q = threadIdx.x; q < nq; q += blockDim.x) rows_s[q] = (int)operand_row(s, padded_row(s, img, q));
  const int h = blockIdx.y * (blockDim.x >> 1) + (threadIdx.x >> 1);
  const int half = threadIdx.x & 1;
  // positions of this thread: the upper half owns the LAST KH positions, so the sentinel (position k) always sits in
  // its slot KH - 1 (a compile-time register); the lower half owns [0, P - KH) and pads
  const int j0 = half ? P - KH : 0;
  const int nv = half ? KH : P - KH;  // valid slots
  for (int i = threadIdx.x; i < nq * 2 * KHP; i += blockDim.x) {
    const int q = i / (2 * KHP), r = i % (2 * KHP);
    const int hf = r / KHP, ii = r % KHP, j = (hf ? P - KH : 0) + ii;
    de_s[i] = (ii < (hf ? KH : P - KH)) ? de[padded_row(s, img, q) * P + j] : 0.f;
  }
  __syncthreads();
  const bool active = h < H;
  const int hc = active ? h : 0;
  float kc[KH], acc[KH];
#pragma unroll
  for (int i = 0; i < KH; ++i) {
    acc[i] = 0.f;
    kc[i] = (i < nv && j0 + i < k) ? to_f(keys[((long)img * k + j0 + i) * H + hc]) : 0.f;
  }
  const float wh = w_alpha[hc];
  float dw = 0.f;
  if (nq > 1) {
    asm volatile("prefetch.global.L1 [%0];" ::"l"(hid_emb + (long)rows_s[1] * H + hc));
    asm volatile("prefetch.global.L1 [%0];" ::"l"(sent_emb + (long)rows_s[1] * H + hc));
  }
  // The per-query operands are loaded one iteration ahead, and the line of the query after that is prefetched into L1:
  // under the 80-register cap ptxas sinks the look-ahead loads to ~30 instructions before their first use (26 % of
  // all stall samples sat there), the prefetch costs no register and turns that wait into an L1 hit.
  T he_raw = from_f<T>(0.f), se_raw = from_f<T>(0.f);
  if (nq > 0) { he_raw = hid_emb[(long)rows_s[0] * H + hc]; se_raw = sent_emb[(long)rows_s[0] * H + hc]; }
  for (int q = 0; q < nq; ++q) {
    const long row = rows_s[q];
    const float he = to_f(he_raw), se = to_f(se_raw);
    if (q + 2 < nq) {
      asm volatile("prefetch.global.L1 [%0];" ::"l"(hid_emb + (long)rows_s[q + 2] * H + hc));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(sent_emb + (long)rows_s[q + 2] * H + hc));
    }
    if (q + 1 < nq) {
      he_raw = hid_emb[(long)rows_s[q + 1] * H + hc];
      se_raw = sent_emb[(long)rows_s[q + 1] * H + hc];
    }
    const float4* dq4 = reinterpret_cast<const float4*>(de_s + (q * 2 + half) * KHP);
    float dhe = 0.f, g_s = 0.f;
#pragma unroll
    for (int i4 = 0; i4 < (KH + 3) / 4; ++i4) {
      const float4 d4 = dq4[i4];
      const float dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i4 * 4 + u;
        if (i < KH) {
          const float kv = (i == KH - 1 && half) ? se : kc[i];   // padded slots have de == 0
          const float z = tanh_t<F>(kv + he);
          const float uu = fmaf(-z, z, 1.f);
          acc[i] = fmaf(dd[u], uu, acc[i]);
          dhe = fmaf(dd[u], uu, dhe);
          dw = fmaf(dd[u], z, dw);
          if (i == KH - 1) g_s = dd[u] * uu;
        }
      }
    }
    dhe += __shfl_xor_sync(0xffffffffu, dhe, 1);
    if (active) {
      if (half == 0) dhid_emb[row * H + h] = from_f<T>(dhe * wh);
      if (half) dsent_emb[row * H + h] = from_f<T>(g_s * wh);
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < KH; ++i)
      if (i < nv && j0 + i < k) dKV[((long)img * k + j0 + i) * (2 * H) + H + h] = from_f<T>(acc[i] * wh);
    atomicAdd(dw_alpha + h, dw);
  }
}

// choose how many CTAs share one image's query rows: minimise waves x rows-per-CTA
int pick_chunks(int n_img, int nq) {
  int best = 1;
  long best_cost = -1;
  for (int c = 1; c <= nq && c <= 8; ++c) {
    const int qpc = (nq + c - 1) / c;
    const long waves = ((long)n_img * c + kNumSMs - 1) / kNumSMs;
    const long cost = waves * (qpc + 1);  // +1: tile load
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

template <typename T, bool STAGE>
void launch_fwd(LaunchCtx c, const AttnShape& s, int chunks, int qpc, size_t smem, const T* keys, const T* vals,
                const T* hid_emb, const T* sent_emb, const T* sentinel, const T* p_hid, const float* w_alpha,
                const float* b_alpha, float* alpha, T* actx, T* ctx_out) {
  const int grid = s.n_img * chunks;
#define LAUNCH_NV(NVV)                                                                                              \
  do {                                                                                                               \
    auto kern = attn_fwd_kernel<T, NVV, STAGE>;                                                                      \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                              \
    kern<<<grid, kFwdThreads, smem, c.stream>>>(s, chunks, qpc, keys, vals, hid_emb, sent_emb, sentinel, p_hid,      \
                                                w_alpha, b_alpha, alpha, actx, ctx_out);                             \
  } while (0)
  if (s.H == 1024) LAUNCH_NV(4);
  else if (s.H == 512) LAUNCH_NV(2);
  else if (s.H == 256) LAUNCH_NV(1);
  else LAUNCH_NV(0);
#undef LAUNCH_NV
  if (c.launches) ++*c.launches;
}

}  // namespace

template <typename T>
void attention_fwd(LaunchCtx c, const AttnShape& s, const T* keys, const T* vals, const T* hid_emb, const T* sent_emb,
                   const T* sentinel, const T* p_hid, const float* w_alpha, const float* b_alpha, float* alpha, T* actx,
                   T* ctx_out) {
  if (s.n_img <= 0 || s.q_per_img <= 0) return;
  const int chunks = pick_chunks(s.n_img, s.q_per_img);
  const int qpc = (s.q_per_img + chunks - 1) / chunks;
  const int q_bytes = 3 * 4 * s.H * (int)sizeof(T);
  const SmemLayout Ls = make_layout(s.H, s.k, (int)sizeof(T), 2, q_bytes);
  // stage both tiles in shared memory when they fit (bf16 at H=1024,k=49: 2 x 98 KB), otherwise read through L1/L2
  const bool stage = Ls.total <= 227 * 1024 && ((size_t)s.k * s.H * sizeof(T)) % 16 == 0;
  if (stage) {
    launch_fwd<T, true>(c, s, chunks, qpc, Ls.total, keys, vals, hid_emb, sent_emb, sentinel, p_hid, w_alpha, b_alpha,
                        alpha, actx, ctx_out);
  } else {
    const SmemLayout L0 = make_layout(s.H, s.k, (int)sizeof(T), 0, q_bytes);
    launch_fwd<T, false>(c, s, chunks, qpc, L0.total, keys, vals, hid_emb, sent_emb, sentinel, p_hid, w_alpha, b_alpha,
                         alpha, actx, ctx_out);
  }
}

bool attention_fwd_batched_supported(const AttnShape& s) {
  if (!s.position_softmax || s.q_per_img < 8 || s.k < 1 || s.k > 63) return false;
  if (s.H != 256 && s.H != 512 && s.H != 1024) return false;
  return scores_smem_bytes(s.H, s.k) <= 227 * 1024;
}

// Training-shaped forward: scores + soft-max kernel (MUFU + mma.sync), then the context as one batched tcgen05 GEMM
//   actx[q, :] = sum_j alpha[q, j] / (k+1) * v_j  +  (alpha[q, k] / (k+1) * sentinel[q] + p_hid[q])
// per image ([q_per_img x k] . [k x H], values as the MN-major B operand, the per-query terms as the epilogue addend).
// scratch: alpha_T [rows, kAlphaTLd] bf16 (kept: the backward pass reads it), sadd [rows, H] bf16.
int attention_fwd_batched(GemmContext& gc, LaunchCtx c, const AttnShape& s, const bf16* keys, const bf16* vals,
                          const bf16* hid_emb, const bf16* sent_emb, const bf16* sentinel, const bf16* p_hid,
                          const float* w_alpha, const float* b_alpha, float* alpha, bf16* alpha_T, bf16* sadd,
                          bf16* actx) {
  if (s.n_img <= 0 || s.q_per_img <= 0) return 0;
  const int H = s.H, k = s.k, P = s.k + 1;
  const long rows = (long)s.n_img * s.q_per_img;
  const size_t smem = scores_smem_bytes(H, k);
  // ~9 queries per CTA: 34 teacher-forced steps = 4 chunks per image, 1024 CTAs at 2 per SM; the key tile (98 KB from L2)
  // is re-loaded per chunk, ~2 us against ~25 us of MUFU work
  const int qpc = 9, chunks = (s.q_per_img + qpc - 1) / qpc;
#define LAUNCH_SC(NBV)                                                                                              \
  do {                                                                                                               \
    auto kern = attn_scores_kernel<NBV>;                                                                             \
    static bool attr_set = false;                                                                                    \
    if (!attr_set) {                                                                                                 \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);                           \
      attr_set = true;                                                                                               \
    }                                                                                                                \
    kern<<<dim3(s.n_img, chunks), kScThreads, smem, c.stream>>>(s, keys, hid_emb, sent_emb, w_alpha, b_alpha, alpha, \
                                                                alpha_T, kAlphaTLd, qpc);                            \
  } while (0)
  if (H == 1024) LAUNCH_SC(4);
  else if (H == 512) LAUNCH_SC(2);
  else LAUNCH_SC(1);
#undef LAUNCH_SC
  if (c.launches) ++*c.launches;
  {
    const long live = s.pm ? s.packed_rows : rows;
    long blocks = (live * (H / 8) + 255) / 256;
    if (blocks > 8L * kNumSMs) blocks = 8L * kNumSMs;
    if (blocks > 0) attn_sadd_kernel<<<(int)blocks, 256, 0, c.stream>>>(live, H, P, s.pm, alpha, sentinel, p_hid, sadd);
    if (c.launches) ++*c.launches;
  }
  GemmCall g;
  g.A = alpha_T; g.lda = s.q_stride * kAlphaTLd; g.batch_a = s.row0_stride * kAlphaTLd;
  g.B = vals; g.ldb = H; g.b_mn = true; g.batch_b = (long)k * H;
  g.C = actx; g.ldc = s.q_stride * H; g.batch_c = s.row0_stride * H;
  g.M = s.q_per_img; g.N = H; g.K = k; g.batch = s.n_img;
  g.epi.addend = sadd; g.epi.ld_add = s.q_stride * H; g.epi.batch_add = s.row0_stride * H;
  if (gemm<bf16>(gc, g)) return -1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attention_fwd_batched: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

// Inference-shaped forward (q_per_img <= 8 query rows per image): the scores + soft-max kernel above, then the small
// context kernel.  Versus the fused single-CTA-per-image kernel (both tiles staged, 1 CTA per SM): the key tile alone
// lets 2 CTAs share an SM and the tensor-core dot product frees the issue slots (70 -> ~30 us per step at 512 x 3 rows).
bool attention_fwd_small_supported(const AttnShape& s) {
  if (!s.position_softmax || s.q_per_img < 1 || s.q_per_img > 8 || s.k < 1 || s.k > 63) return false;
  if (s.H != 256 && s.H != 512 && s.H != 1024) return false;
  return scores_smem_bytes(s.H, s.k) <= 227 * 1024;
}
int attention_fwd_small(LaunchCtx c, const AttnShape& s, const bf16* keys, const bf16* vals, const bf16* hid_emb,
                        const bf16* sent_emb, const bf16* sentinel, const bf16* p_hid, const float* w_alpha,
                        const float* b_alpha, float* alpha, bf16* actx) {
  if (s.n_img <= 0) return 0;
  const int H = s.H, k = s.k, P = s.k + 1;
  const size_t smem = scores_smem_bytes(H, k);
#define LAUNCH_SC(NBV)                                                                                              \
  do {                                                                                                               \
    auto kern = attn_scores_kernel<NBV>;                                                                             \
    static bool attr_set = false;                                                                                    \
    if (!attr_set) {                                                                                                 \
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);                           \
      attr_set = true;                                                                                               \
    }                                                                                                                \
    kern<<<s.n_img, kScThreads, smem, c.stream>>>(s, keys, hid_emb, sent_emb, w_alpha, b_alpha, alpha, nullptr, 0,   \
                                                  s.q_per_img);                                                      \
  } while (0)
  if (H == 1024) LAUNCH_SC(4);
  else if (H == 512) LAUNCH_SC(2);
  else LAUNCH_SC(1);
#undef LAUNCH_SC
  attn_ctx_small_kernel<8><<<s.n_img, 256, (size_t)s.q_per_img * P * sizeof(float), c.stream>>>(s, vals, sentinel, p_hid,
                                                                                              alpha, actx);
  if (c.launches) *c.launches += 2;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("attention_fwd_small: %s", cudaGetErrorString(e)); return -1; }
  return 0;
}

template <typename T>
int attention_bwd(GemmContext& gc, LaunchCtx c, const AttnShape& s, const T* keys, const T* vals, const T* hid_emb,
                  const T* sent_emb, const T* sentinel, const float* w_alpha, const float* alpha, const T* dactx,
                  float* dal, T* alpha_T, float* de, T* dS_ctx, T* dKV, T* dhid_emb, T* dsent_emb, float* dw_alpha,
                  float* db_alpha) {
  if (s.n_img <= 0 || s.q_per_img <= 0) return 0;
  const int H = s.H, k = s.k, P = s.k + 1;
  const long rows = (long)s.n_img * s.q_per_img;
  const long r0 = s.row0_stride, qs = s.q_stride;
  if (s.position_softmax) {
    // d alpha[q, j] = dctx[q, :] . v_j  for the k image positions: one [q_per_img x k x H] GEMM per image on the
    // tensor cores (batched through 3-D tensor maps; rows of one image are q_stride apart)
    GemmCall g;
    g.A = dactx; g.lda = qs * H; g.batch_a = r0 * H;
    g.B = vals; g.ldb = H; g.batch_b = (long)k * H;
    g.C = dal; g.ldc = qs * kDalLd; g.batch_c = r0 * kDalLd;
    g.M = s.q_per_img; g.N = k; g.K = H; g.batch = s.n_img; g.epi.out_f32 = 1;
    if (gemm<T>(gc, g)) return -1;
  } else {
    // reference as written: alpha == 1 is constant -> no gradient reaches the scores
    cudaMemsetAsync(de, 0, rows * P * sizeof(float), c.stream);
    cudaMemsetAsync(dhid_emb, 0, rows * H * sizeof(T), c.stream);
    cudaMemsetAsync(dsent_emb, 0, rows * H * sizeof(T), c.stream);
    cudaMemset2DAsync(dKV + H, (size_t)2 * H * sizeof(T), 0, (size_t)H * sizeof(T), (size_t)s.n_img * k, c.stream);
  }
  {
    const long live = s.pm ? s.packed_rows : rows;
    long blocks = (live + 7) / 8;
    if (blocks > 16L * kNumSMs) blocks = 16L * kNumSMs;
    if (blocks > 0)
      attn_bwd_rows_kernel<T><<<(int)blocks, 256, 0, c.stream>>>(live, H, k, s.position_softmax, s.pm, sentinel, alpha,
                                                                dactx, dal, kDalLd, de, alpha_T, kAlphaTLd, dS_ctx,
                                                                db_alpha);
    if (c.launches) ++*c.launches;
  }
  {
    // d values (pre-tanh)[j, h] = (sum_q alpha[q, j] / (k+1) * dctx[q, h]) * (1 - v[j, h]^2): the weight-gradient form
    // (contraction over the query index), one [k x H x q_per_img] GEMM per image, tanh' fused in the epilogue
    GemmCall g;
    g.a_mn = true; g.b_mn = true;
    g.A = alpha_T; g.lda = qs * kAlphaTLd; g.batch_a = r0 * kAlphaTLd;
    g.B = dactx; g.ldb = qs * H; g.batch_b = r0 * H;
    g.C = dKV; g.ldc = 2 * H; g.batch_c = (long)k * 2 * H;
    g.M = k; g.N = H; g.K = s.q_per_img; g.batch = s.n_img;
    g.epi.act = ACT_DTANH; g.epi.dact_y = vals; g.epi.ld_y = H; g.epi.batch_y = (long)k * H;
    if (gemm<T>(gc, g)) return -1;
  }
  if (!s.position_softmax) return 0;
  // KH = positions per thread (two threads per column): the smallest instantiated size with 2 * KH >= k + 1
  dim3 grid_cols(s.n_img, (H + 127) / 128);
#define DISPATCH_KC(KERNEL, ...)                                                                                     \
  do {                                                                                                               \
    const int kh = P <= 8 ? 4 : P <= 16 ? 8 : P <= 32 ? 16 : P <= 50 ? 25 : 32;                                      \
    const size_t a_bytes = (size_t)s.q_per_img * (2 * ((kh + 3) & ~3) * sizeof(float) + sizeof(int));                \
    if (kh == 25) KERNEL<T, 25><<<grid_cols, 256, a_bytes, c.stream>>>(__VA_ARGS__);                                 \
    else if (kh == 4) KERNEL<T, 4><<<grid_cols, 256, a_bytes, c.stream>>>(__VA_ARGS__);                              \
    else if (kh == 8) KERNEL<T, 8><<<grid_cols, 256, a_bytes, c.stream>>>(__VA_ARGS__);                              \
    else if (kh == 16) KERNEL<T, 16><<<grid_cols, 256, a_bytes, c.stream>>>(__VA_ARGS__);                            \
    else KERNEL<T, 32><<<grid_cols, 256, a_bytes, c.stream>>>(__VA_ARGS__);                                          \
    if (c.launches) ++*c.launches;                                                                                   \
  } while (0)
  DISPATCH_KC(attn_bwd_dk_kernel, s, keys, hid_emb, sent_emb, w_alpha, de, dKV, dhid_emb, dsent_emb, dw_alpha);
#undef DISPATCH_KC
  return 0;
}

#define INSTANTIATE(T)                                                                                               \
  template void attention_fwd<T>(LaunchCtx, const AttnShape&, const T*, const T*, const T*, const T*, const T*,      \
                                 const T*, const float*, const float*, float*, T*, T*);                              \
  template int attention_bwd<T>(GemmContext&, LaunchCtx, const AttnShape&, const T*, const T*, const T*, const T*,    \
                                const T*, const float*, const float*, const T*, float*, T*, float*, T*, T*, T*, T*,    \
                                float*, float*);
INSTANTIATE(float)
INSTANTIATE(bf16)

}  // namespace capdec
