// bf16 GEMM on the 5th-gen tensor cores, persistent and warp-specialised:
//   TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> shared-memory ring -> tcgen05.mma (cta_group::1, M=128, N=BN,
//   K=16 per instruction) -> fp32 accumulators in TMEM (two buffers) -> tcgen05.ld -> fused epilogue -> global.
//
// One CTA per SM loops over (tile, k-split) work items.  Warp roles (192 threads): warp 0 = TMA producer (one
// elected lane), warp 1 = TMEM allocator + MMA issuer (one elected lane), warps 2..5 = epilogue (warp w owns TMEM
// lanes 32*(w%4) .. +31).  The TMEM accumulator is double-buffered, so the epilogue of tile i (bias / addend / tanh /
// tanh' / fp32 vector reductions for split-K) overlaps the MMAs of tile i+1, and the smem ring keeps streaming
// across tile boundaries.  Each operand is either K-major or MN-major (UMMA descriptor + instruction-descriptor
// major bits), which covers forward fc (B = Paddle's [in,out] weight, MN-major), data-grad and weight-grad forms
// without transposed weight copies.
//
// Replaces the cuBLAS SGEMMs behind every layers.fc / layers.matmul on the reference's decoder path
// (model_adaAttention_aic.py:24-25, 52-53, 87-90, 99-104, 115) and their *_grad ops.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "gemm.cuh"

namespace capdec {

namespace {

constexpr int BM = 128;  // rows of one UMMA; a CTA tile is MT of them stacked (MT = 1 or 2)
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kEpiWarps = 8;                     // two warps per TMEM lane quarter
constexpr int kThreads = 64 + kEpiWarps * 32;    // warp 0: TMA, warp 1: MMA, warps 2..9: epilogue
constexpr int kABytes = BM * BK * 2;  // 16 KB

// MT = 2: a 256-row CTA tile (two UMMA row blocks sharing every B k-block).  The decoder's big GEMMs are bound by the
// L2 -> shared-memory operand stream (all 148 SMs pull (BM + BN) * 128 B per k-block, ~7 TB/s chip-wide), not by the
// tensor pipe, and a 256 x 256 tile moves 1/3 less operand data per flop than 128 x 256.
template <int BN, int MT> struct TcCfg {
  static constexpr int kABytesT = MT * kABytes;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytesT + kBBytes;
  static constexpr int kStages = (MT == 2) ? (BN == 256 ? 3 : 4) : (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kAccBufs = (2 * MT * BN <= 512) ? 2 : 1;  // accumulator buffers that fit the 512 TMEM columns
  static constexpr int kTmemCols = kAccBufs * MT * BN;           // power of two >= 32
  static constexpr int kAlignSlack = 512;  // dynamic smem starts 1024-aligned in practice; checked at kernel entry
  static constexpr int kStagingBytes = kEpiWarps * 4096;  // per epilogue warp: one 32-row x 128-byte output tile
  static constexpr int kSmemBytes =
      kStages * kStageBytes + kStagingBytes + kAlignSlack + 256 /*barriers*/ + 2 * BN * 4 /*bias*/;
};

// 64-bit shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1      bits [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// descriptor of the k-th K=16 slice of an operand tile
//   K-major : rows of 128 B (64 k), 8-row groups 1024 B apart (SBO); a slice is 32 B further along the row
//   MN-major: every k-row is 128 B (64 mn elements), 8-k groups 1024 B apart (SBO), the next 64-wide mn group is one
//             TMA box (64 k-rows x 128 B = 8192 B) further (LBO); a slice is 16 k-rows = 2048 B further
template <bool MN> __device__ __forceinline__ uint64_t slice_desc(uint32_t tile_addr, int k) {
  return MN ? make_smem_desc(tile_addr + k * 2048, 8192, 1024) : make_smem_desc(tile_addr + k * 32, 16, 1024);
}

// TMA store / reduce of one [32 rows x 128 bytes] SWIZZLE_128B tile from shared memory (SASS UTMASTG / UTMAREDG)
// all tensor maps are 3-D: {innermost, row, batch}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// Shared-memory accesses of the epilogue by 32-bit shared-window address.  Going through generic pointers made ptxas
// rebuild the generic address (S2UR SR_CgaCtaId + ULEA) and convert it back at every pass: 8 % of the stall samples of the
// epilogue-bound vocabulary GEMM sat on that special-register read (ncu source page).
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void tma_store_3d_s(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d_s(const void* tmap, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(tmap),
               "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Programmatic dependent launch: a GEMM launched with the programmatic-stream-serialization attribute may start while
// its predecessor in the stream is still draining; everything up to pdl_wait() (barrier init, TMEM allocation, tensor-map
// prefetch) touches no global memory, so the launch latency and the prologue of GEMM n+1 hide behind the tail of
// kernel n.  pdl_wait() returns once the prerequisite grid has completed and its writes are visible.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

struct TcShape {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split;
  int batch;
  long ldc, batch_c;
  int vec_ok;
  int rotate_k;  // start every CTA's k-loop at a different k-block (spreads L2 hot lines among CTAs sharing a tile)
  int tma_out;   // epilogue writes through shared memory + TMA tensor stores (needs an aligned C)
  unsigned long long* trace;  // debug timeline (capdec_debug_gemm_trace), 32 slots per CTA, or nullptr
};
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TRACE(slot) do { if (sh.trace) sh.trace[blockIdx.x * 32 + (slot)] = gtime(); } while (0)

// Rare ragged-edge paths are kept out of line: the epilogue runs with ONE warp per scheduler, so its loop body has to
// stay inside the instruction cache (an inlined scalar fallback per unrolled group made the kernel 120 KB of SASS and
// the epilogue 3x slower than the MMAs it is supposed to hide behind).
__device__ __noinline__ void epi_math8_tail(float* v, const GemmEpi& epi, long row, int col, bool row_ok, int N) {
  for (int i = 0; i < 8; ++i)
    v[i] = (row_ok && col + i < N) ? epi_value<bf16>(v[i], (int)row, col + i, epi) : 0.f;
}
__device__ __noinline__ void epi_store8_tail(const float* v, void* C, long ldc, long row, int col, int N,
                                             const GemmEpi& epi) {
  for (int i = 0; i < 8 && col + i < N; ++i) epi_store<bf16>(v[i], C, ldc, (int)row, col + i, epi);
}

// 32 accumulator columns of one warp's 32 rows: TMEM -> registers (tcgen05.ld) with the addend / dact_y loads in
// flight behind it -> bias (smem) / addend / tanh / tanh' -> swizzled smem tile (16-byte chunks chunk0 .. of each
// 128-byte row).  FULL = all 32 columns are < N: straight-line code the scheduler can interleave freely (the
// epilogue has only two warps per scheduler, so exposed latencies are what it costs).
template <bool OUT_F32, bool FULL>
__device__ __forceinline__ void epilogue_cols32(uint32_t tmem_addr, uint32_t bias_t, const bf16* add_p,
                                                const bf16* y_p, int fact, float alpha, int left, uint32_t dst_row,
                                                uint32_t swz, uint32_t chunk0) {
  uint32_t r[32];
  uint4 aq[4], yq[4];
  __syncwarp();  // tcgen05.ld is .sync.aligned
  tmem_ld32(tmem_addr, r);
  if (add_p) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (FULL || g * 8 < left) aq[g] = *reinterpret_cast<const uint4*>(add_p + g * 8);
  }
  if (y_p) {
#pragma unroll
    for (int g = 0; g < 4; ++g)
      if (FULL || g * 8 < left) yq[g] = *reinterpret_cast<const uint4*>(y_p + g * 8);
  }
  tmem_ld_wait();
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    if (FULL || g * 8 < left) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]) * alpha;
      if (bias_t) {  // shared-window address of this pass's 32 bias values (0: no bias)
        const float4 b0 = lds_f4(bias_t + (uint32_t)(g * 32));
        const float4 b1 = lds_f4(bias_t + (uint32_t)(g * 32 + 16));
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
        v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
      }
      if (add_p) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&aq[g]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          v[2 * i] += f.x; v[2 * i + 1] += f.y;
        }
      }
      if (fact & 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = tanh_t<true>(v[i]);
      } else if (fact & 2) {  // ACT_RELU
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
      }
      if (y_p) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&yq[g]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          if (fact & 4) {  // ACT_DRELU: y is the ReLU output
            v[2 * i] = f.x > 0.f ? v[2 * i] : 0.f; v[2 * i + 1] = f.y > 0.f ? v[2 * i + 1] : 0.f;
          } else {
            v[2 * i] *= (1.f - f.x * f.x); v[2 * i + 1] *= (1.f - f.y * f.y);
          }
        }
      }
      if (OUT_F32) {
        sts_v4(dst_row + (((chunk0 + 2 * g) ^ swz) << 4), __float_as_uint(v[0]), __float_as_uint(v[1]),
               __float_as_uint(v[2]), __float_as_uint(v[3]));
        sts_v4(dst_row + (((chunk0 + 2 * g + 1) ^ swz) << 4), __float_as_uint(v[4]), __float_as_uint(v[5]),
               __float_as_uint(v[6]), __float_as_uint(v[7]));
      } else {
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        sts_v4(dst_row + (((chunk0 + g) ^ swz) << 4), u.x, u.y, u.z, u.w);
      }
    }
  }
}

// ACT_ROWSTATS: running (max, sum exp(x - max)) of x = acc + bias over 32 accumulator columns of this thread's row
__device__ __forceinline__ void epilogue_stats32(uint32_t tmem_addr, const float* bias_t, int left, float& m, float& s) {
  uint32_t r[32];
  __syncwarp();
  tmem_ld32(tmem_addr, r);
  tmem_ld_wait();
  float x[32];
  float cm = -INFINITY;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    x[i] = (i < left) ? __uint_as_float(r[i]) + (bias_t ? bias_t[i] : 0.f) : -INFINITY;
    cm = fmaxf(cm, x[i]);
  }
  if (cm > m) { s *= __expf(m - cm); m = cm; }   // m == -inf initially: s == 0, exp(-inf) = 0
  if (m > -INFINITY) {
#pragma unroll
    for (int i = 0; i < 32; ++i) s += __expf(x[i] - m);
  }
}

template <int BN, bool A_MN, bool B_MN, int MT>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, void* __restrict__ C, TcShape sh, GemmEpi epi) {
  using Cfg = TcCfg<BN, MT>;
  constexpr int BMT = BM * MT;          // rows of the CTA tile
  constexpr int NB = Cfg::kAccBufs;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t misalign = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (misalign > (uint32_t)Cfg::kAlignSlack) __trap();
  uint8_t* smem = smem_raw + misalign;
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;  // 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(staging + Cfg::kStagingBytes + 256);  // [2][BN]
  // the same two as shared-window addresses; the volatile move keeps ptxas from re-deriving the value (a special-register
  // read and the alignment arithmetic) at every use instead of holding it in a register
  uint32_t staging_sa;
  asm volatile("mov.u32 %0, %1;" : "=r"(staging_sa) : "r"(smem_u32(staging)));
  const uint32_t bias_sa = staging_sa + (uint32_t)Cfg::kStagingBytes + 256u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb_total = (sh.K + BK - 1) / BK;
  const int tiles = sh.tiles_m * sh.tiles_n;
  const int tb = tiles * sh.batch;           // work item w -> (split z, batch bt, tile)
  const int total_work = tb * sh.splits;

  if (threadIdx.x == 0) {
    TRACE(0);
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (sh.tma_out) tma_prefetch_desc(&tmC);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], kEpiWarps);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  if (threadIdx.x == 0) TRACE(1);
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int s = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int z = w / tb, bt = (w % tb) / tiles, tile = w % tiles;
        const int n0 = (tile % sh.tiles_n) * BN, m0 = (tile / sh.tiles_n) * BMT;
        const int kb0 = z * sh.kb_per_split, kb1 = min(nkb_total, kb0 + sh.kb_per_split);
        const int nk = kb1 - kb0;
        const int rot = sh.rotate_k ? (int)((blockIdx.x / sh.tiles_n) % nk) : 0;
        for (int i = 0; i < nk; ++i) {
          const int kb = kb0 + (i + rot < nk ? i + rot : i + rot - nk);
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sA = smem + s * Cfg::kStageBytes;
          uint8_t* sB = sA + Cfg::kABytesT;
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          if (!A_MN) {
#pragma unroll
            for (int mh = 0; mh < MT; ++mh) tma_load_3d(sA + mh * kABytes, &tmA, kb * BK, m0 + mh * BM, bt, &full_bar[s]);
          } else {
#pragma unroll
            for (int i = 0; i < BMT / 64; ++i) tma_load_3d(sA + i * 8192, &tmA, m0 + i * 64, kb * BK, bt, &full_bar[s]);
          }
          if (!B_MN) {
            tma_load_3d(sB, &tmB, kb * BK, n0, bt, &full_bar[s]);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_3d(sB + i * 8192, &tmB, n0 + i * 64, kb * BK, bt, &full_bar[s]);
          }
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
          if (w == (int)blockIdx.x && i == 0) TRACE(2);
        }
      }
      TRACE(3);
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // instruction descriptor: D=f32 (bit4), A=bf16 (bit7), B=bf16 (bit10), a/b major (bits 15/16),
      // N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                 ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
        const int z = w / tb;
        const int kb0 = z * sh.kb_per_split, kb1 = min(nkb_total, kb0 + sh.kb_per_split);
        const int buf = it % NB;
        const uint32_t acc_ph = (it / NB) & 1;
        mbar_wait(&tmem_empty_bar[buf], acc_ph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * MT * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          if (it == 0 && kb == kb0) TRACE(4);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + Cfg::kABytesT;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
#pragma unroll
            for (int mh = 0; mh < MT; ++mh)  // the row blocks share the B slice
              umma_bf16(tmem_d + (uint32_t)(mh * BN), slice_desc<A_MN>(a_addr + mh * kABytes, k),
                        slice_desc<B_MN>(b_addr, k), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full_bar[buf]);  // accumulator complete
        if (it < 4) TRACE(8 + it);         // all MMAs of tile `it` issued
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    const int ew = warp - 2;        // epilogue warp index 0..7
    const int half = ew >> 2;       // which of the two warps of the quarter
    int it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
      const int bt = (w % tb) / tiles, tile = w % tiles;
      const int n0 = (tile % sh.tiles_n) * BN, m0 = (tile / sh.tiles_n) * BMT;
      const int buf = it % NB;
      const uint32_t acc_ph = (it / NB) & 1;
      const bool has_bias = epi.bias != nullptr;
      if (sh.tma_out && has_bias) {
        // stage this tile's bias slice (zero past the readable range) while the MMAs are still running
        const int n_read = (sh.N + 7) & ~7;
        for (int i = threadIdx.x - 64; i < BN; i += kEpiWarps * 32)
          sts_f32(bias_sa + (uint32_t)(((it & 1) * BN + i) * 4), (n0 + i < n_read) ? epi.bias[n0 + i] : 0.f);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");  // the epilogue warps only
      }
      mbar_wait(&tmem_full_bar[buf], acc_ph);
      if (warp == 2 && lane == 0 && it < 4) TRACE(12 + it);   // accumulator of tile `it` complete
      tc_fence_after();
#pragma unroll 1
      for (int mh = 0; mh < MT; ++mh) {  // the 128-row blocks of the CTA tile, one after the other
        const bool last_block = (mh == MT - 1);
        const int row = m0 + mh * BM + quarter * 32 + lane;
        const bool row_ok = row < sh.M;
        const uint32_t tmem_d =
            tmem_base + (uint32_t)(buf * MT * BN + mh * BN) + ((uint32_t)(quarter * 32) << 16);
        if (sh.tma_out) {
          // ---- registers -> swizzled smem tile (32 rows x 128 B) -> TMA tensor store / reduce-add ----
          // Passes (128 bytes of output per row each) are dealt round-robin to the two warps of a lane quarter.  The
          // host guarantees (launch()) that bias / addend / dact_y may be read in whole groups of 8 columns up to the
          // next multiple of 8 past N; columns >= N are clipped by the TMA store, so there is no scalar edge path.
          const uint32_t my_stage = staging_sa + (uint32_t)(ew * 4096);
          const uint32_t swz = (uint32_t)(lane & 7);
          const uint32_t my_row = my_stage + (uint32_t)(lane * 128);
          const long add_row = epi.add_mod > 0 ? row % epi.add_mod : row;
          const bf16* add_p = (epi.addend && row_ok) ? reinterpret_cast<const bf16*>(epi.addend) + bt * epi.batch_add +
                                                           add_row * epi.ld_add + n0
                                                     : nullptr;
          const bf16* y_p = ((epi.act == ACT_DTANH || epi.act == ACT_DRELU) && row_ok)
                                ? reinterpret_cast<const bf16*>(epi.dact_y) + bt * epi.batch_y + (long)row * epi.ld_y + n0
                                : nullptr;
          const int do_tanh = (epi.act == ACT_TANH ? 1 : 0) | (epi.act == ACT_RELU ? 2 : 0) | (epi.act == ACT_DRELU ? 4 : 0);
          const float alpha = epi.alpha_dev ? __ldg(epi.alpha_dev) : 1.f;
          const bool f32_out = epi.out_f32 || epi.atomic;
          const int pw = f32_out ? 32 : 64;           // columns per pass
          const int n_pass = BN / pw;
          const int my_last = ((n_pass - 1 - half) & ~1) + half;  // last pass index owned by this warp (< 0: none)
          // none of this warp's 32 rows exists (small-M problems, e.g. the per-image attention GEMMs with M = 34 / 49):
          // nothing to read, convert or store
          const bool warp_empty = (m0 + mh * BM + quarter * 32) >= sh.M;
          if ((half >= n_pass || warp_empty) && last_block) {
            // nothing to read for this warp (BN = 64, bf16): release the accumulator right away
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
          }
#pragma unroll 1
          for (int ps = warp_empty ? n_pass : half; ps < n_pass; ps += 2) {
            const int c0 = ps * pw;
            const int left = sh.N - (n0 + c0);  // valid columns from here on (warp-uniform)
            if (lane == 0) bulk_wait_read<0>();  // this warp's previous store has finished reading the staging tile
            __syncwarp();
            const float* bsp = has_bias ? bias_s + (it & 1) * BN + c0 : nullptr;          // ACT_ROWSTATS path only
            const uint32_t bsa = has_bias ? bias_sa + (uint32_t)(((it & 1) * BN + c0) * 4) : 0u;
            const bf16* ap = add_p ? add_p + c0 : nullptr;
            const bf16* yp = y_p ? y_p + c0 : nullptr;
            if (epi.act == ACT_ROWSTATS) {  // pw == 64: one 64-column group per pass, nothing is stored
              float rm = -INFINITY, rs = 0.f;
              epilogue_stats32(tmem_d + (uint32_t)c0, bsp, left, rm, rs);
              epilogue_stats32(tmem_d + (uint32_t)(c0 + 32), bsp ? bsp + 32 : nullptr, left - 32, rm, rs);
              if (row_ok && left > 0) {
                const long o = (long)row * epi.stats_ld + ((n0 + c0) >> 6);
                epi.stats_m[o] = rm;
                epi.stats_s[o] = rs;
              }
              if (ps == my_last && last_block) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
              }
              continue;
            }
            if (f32_out) {
              if (left >= 32) epilogue_cols32<true, true>(tmem_d + (uint32_t)c0, bsa, ap, yp, do_tanh, alpha, left, my_row, swz, 0);
              else epilogue_cols32<true, false>(tmem_d + (uint32_t)c0, bsa, ap, yp, do_tanh, alpha, left, my_row, swz, 0);
            } else {
#pragma unroll 1
              for (int sub = 0; sub < 2; ++sub) {  // 2 x 32 columns = one 128-byte bf16 row
                const int o = sub * 32;
                if (left - o >= 32)
                  epilogue_cols32<false, true>(tmem_d + (uint32_t)(c0 + o), bsa ? bsa + (uint32_t)(o * 4) : 0u, ap ? ap + o : nullptr,
                                               yp ? yp + o : nullptr, do_tanh, alpha, left - o, my_row, swz,
                                               (uint32_t)(sub * 4));
                else
                  epilogue_cols32<false, false>(tmem_d + (uint32_t)(c0 + o), bsa ? bsa + (uint32_t)(o * 4) : 0u, ap ? ap + o : nullptr,
                                                yp ? yp + o : nullptr, do_tanh, alpha, left - o, my_row, swz,
                                               (uint32_t)(sub * 4));
              }
            }
            if (ps == my_last && last_block) {
              // all of this warp's TMEM reads are done: hand the accumulator back to the MMA issuer
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
            }
            if (left <= 0) continue;
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (epi.atomic) tma_reduce_add_3d_s(&tmC, my_stage, n0 + c0, m0 + mh * BM + quarter * 32, bt);
              else tma_store_3d_s(&tmC, my_stage, n0 + c0, m0 + mh * BM + quarter * 32, bt);
              bulk_commit();
            }
          }
          continue;
        }
        if (half == 1) {  // the fallback path below uses one warp per lane quarter
          if (last_block) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
          }
          continue;
        }
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          // ---- fallback for outputs the TMA cannot address (unaligned base / pitch): per-thread stores ----
          uint32_t r[32];
          __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent stores of the previous trip
          tmem_ld32(tmem_d + (uint32_t)(c * 32), r);
          tmem_ld_wait();
          if (c == BN / 32 - 1 && last_block) {
            // all of this warp's TMEM reads are done: hand the accumulator back to the MMA issuer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
          }
          const int nb = n0 + c * 32;
          if (!row_ok || nb >= sh.N) continue;
#pragma unroll 1
          for (int g = 0; g < 4; ++g) {
            const int col = nb + g * 8;
            if (col >= sh.N) break;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
            GemmEpi eb = epi;
            if (eb.addend) eb.addend = reinterpret_cast<const bf16*>(eb.addend) + bt * eb.batch_add;
            if (eb.dact_y) eb.dact_y = reinterpret_cast<const bf16*>(eb.dact_y) + bt * eb.batch_y;
            void* Cb = (epi.out_f32 || epi.atomic) ? (void*)(reinterpret_cast<float*>(C) + bt * sh.batch_c)
                                                   : (void*)(reinterpret_cast<bf16*>(C) + bt * sh.batch_c);
            epi_math8_tail(v, eb, row, col, true, sh.N);
            epi_store8_tail(v, Cb, sh.ldc, row, col, sh.N, eb);
          }
        }
      }
      if (warp == 2 && lane == 0 && it < 4) TRACE(16 + it);   // epilogue of tile `it` issued its last store
    }
  }
  if (warp >= 2 && lane == 0 && sh.tma_out) bulk_wait<0>();  // all tensor stores of this thread have completed
  if (warp == 2 && lane == 0) TRACE(20);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ================================================================================================================
// CTA-pair variant (tcgen05 cta_group::2): one 256 x 256 tile per pair of CTAs on the two SMs of a TPC.
//
// The single-CTA mainloop above is bound by shared-memory ports, not by the tensor pipe: per 128x256x16 UMMA it reads
// 12 KB of operands while TMA writes the next k-block into the same shared memory (0.37 us per 48 KB k-block against
// 0.27 us of MMA time).  In a pair, CTA r keeps rows [128r, 128r+128) of A and columns [128r, 128r+128) of B: the
// leader's UMMA (M = 256, N = 256) reads each SM's A half from that SM's shared memory and the two B halves from
// both, so an SM reads 8 KB and receives 32 KB per k-block for the same 128x256x16 of work, and both the read and
// the fill traffic per flop drop by a third.
//   leader (cluster rank 0): issues every UMMA; its full[] barriers count the bytes of BOTH CTAs' TMA loads
//   both CTAs              : TMA producer for their own halves, epilogue for their own 128 accumulator rows
//   tcgen05.commit multicast: frees the stage in both CTAs / publishes the accumulator to both epilogues
//   tmem_empty (leader)    : 2 x 8 epilogue warps arrive (the peer's through the cluster shared window)
// ================================================================================================================
constexpr int kPairBN = 256;
// EPI2: the variant for short-K, epilogue-bound problems (K <= 512: the fp32 vocabulary logits, the x-part pre-activations):
// four pipeline stages are enough for a 4..8 k-block mainloop, and the freed 64 KB give every epilogue warp a second
// staging tile, so the tensor store of pass i drains while pass i+1 is converted (bulk_wait_read<1>).
#ifndef CAPDEC_EPI_DEPTH
#define CAPDEC_EPI_DEPTH 2   // staging tiles per epilogue warp in the EPI2 variant (2: 4 pipeline stages, 3: 3 stages)
#endif
constexpr int kEpiDepth = CAPDEC_EPI_DEPTH;
template <bool EPI2> struct PairCfgT {
  static constexpr int kBHalfBytes = (kPairBN / 2) * BK * 2;    // 16 KB
  static constexpr int kStageBytes = kABytes + kBHalfBytes;     // 32 KB per CTA and stage
  static constexpr int kStages = EPI2 ? (kEpiDepth == 3 ? 3 : 4) : 6;
  static constexpr int kTmemCols = 2 * kPairBN;                 // two accumulator buffers
  static constexpr int kAlignSlack = 512;
  // one 32-row x 128-byte output tile per epilogue warp.  (Measured: two tiles per warp with 5 instead of 6 stages is
  // slower -- 3.01 vs 2.91 ms per train step -- the epilogue does not wait on its previous store.)
  static constexpr int kStagingBytes = kEpiWarps * 4096 * (EPI2 ? kEpiDepth : 1);
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + kAlignSlack + 256 + 2 * kPairBN * 4;
};
using PairCfg = PairCfgT<false>;

__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_rank(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion bytes are counted by a barrier that may live in the peer CTA (cluster address)
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const void* tmap, int c0, int c1, int c2, uint32_t bar_cluster) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_result, int cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, int cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// bounded wait: a protocol error traps (the launch fails loudly) instead of hanging both SMs
__device__ __forceinline__ void mbar_wait_bd(uint64_t* bar, uint32_t parity) {
  unsigned n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++n > (1u << 24)) __trap();
  }
}

template <bool A_MN, bool B_MN, bool EPI2>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, TcShape sh, GemmEpi epi) {
  using Cfg = PairCfgT<EPI2>;
  constexpr int BN = kPairBN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t misalign = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  if (misalign > (uint32_t)Cfg::kAlignSlack) __trap();
  uint8_t* smem = smem_raw + misalign;
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]  (used in the leader only)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* bias_s = reinterpret_cast<float*>(staging + Cfg::kStagingBytes + 256);  // [2][BN]
  // the same two as shared-window addresses; the volatile move keeps ptxas from re-deriving the value (a special-register
  // read and the alignment arithmetic) at every use instead of holding it in a register
  uint32_t staging_sa;
  asm volatile("mov.u32 %0, %1;" : "=r"(staging_sa) : "r"(smem_u32(staging)));
  const uint32_t bias_sa = staging_sa + (uint32_t)Cfg::kStagingBytes + 256u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int nkb_total = (sh.K + BK - 1) / BK;
  const int tiles = sh.tiles_m * sh.tiles_n;
  const int tb = tiles * sh.batch;
  const int total_work = tb * sh.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmC);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 2 * kEpiWarps);  // the epilogue warps of both CTAs
    }
    mbar_fence_init();
  }
  __syncthreads();
  cluster_barrier();  // barriers of both CTAs are initialised before anyone signals across
  if (warp == 1) tmem_alloc_pair(tmem_ptr_smem, Cfg::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer (both CTAs): my 128 rows of A, my 128 columns of B; bytes counted by the LEADER's barrier =====
      int s = 0;
      uint32_t ph = 0;
      // the leader's full barriers in the cluster window: mapped once (mapa reads a special register: ncu showed 8 % of
      // the stall samples of an epilogue-bound launch on that read when it sat inside the tile loops)
      const uint32_t fb0 = map_to_rank(&full_bar[0], 0);
      for (int w = pair; w < total_work; w += n_pairs) {
        const int z = w / tb, bt = (w % tb) / tiles, tile = w % tiles;
        const int n0 = (tile % sh.tiles_n) * BN + (int)rank * (BN / 2);
        const int m0 = (tile / sh.tiles_n) * (2 * BM) + (int)rank * BM;
        const int kb0 = z * sh.kb_per_split, kb1 = min(nkb_total, kb0 + sh.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_bd(&empty_bar[s], ph ^ 1);
          uint8_t* sA = smem + s * Cfg::kStageBytes;
          uint8_t* sB = sA + kABytes;
          const uint32_t fb = fb0 + (uint32_t)s * 8u;
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);
          if (!A_MN) {
            tma_load_3d_pair(sA, &tmA, kb * BK, m0, bt, fb);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_3d_pair(sA + i * 8192, &tmA, m0 + i * 64, kb * BK, bt, fb);
          }
          if (!B_MN) {
            tma_load_3d_pair(sB, &tmB, kb * BK, n0, bt, fb);
          } else {
#pragma unroll
            for (int i = 0; i < (BN / 2) / 64; ++i) tma_load_3d_pair(sB + i * 8192, &tmB, n0 + i * 64, kb * BK, bt, fb);
          }
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ===== MMA issuer (leader only): D[256 x 256] += A[256 x 16] . B[256 x 16]^T per instruction =====
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) |
                                 ((B_MN ? 1u : 0u) << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int w = pair; w < total_work; w += n_pairs, ++it) {
        const int z = w / tb;
        const int kb0 = z * sh.kb_per_split, kb1 = min(nkb_total, kb0 + sh.kb_per_split);
        const int buf = it & 1;
        mbar_wait_bd(&tmem_empty_bar[buf], ((it >> 1) & 1) ^ 1);  // both epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait_bd(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16_pair(tmem_d, slice_desc<A_MN>(a_addr, k), slice_desc<B_MN>(b_addr, k), idesc,
                           (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit_pair(&empty_bar[s]);  // frees the stage in both CTAs once these MMAs retire
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
        umma_commit_pair(&tmem_full_bar[buf]);  // accumulator complete: both epilogues may read their rows
      }
    }
  } else {
    // ===== epilogue (both CTAs): my 128 accumulator rows, all 256 columns =====
    const int quarter = warp & 3;
    const int ew = warp - 2;
    const int half = ew >> 2;
    int it = 0;
    uint32_t stage_flip = 0;
    const uint32_t empty_leader0 = map_to_rank(&tmem_empty_bar[0], 0);
    for (int w = pair; w < total_work; w += n_pairs, ++it) {
      const int bt = (w % tb) / tiles, tile = w % tiles;
      const int n0 = (tile % sh.tiles_n) * BN, m0 = (tile / sh.tiles_n) * (2 * BM) + (int)rank * BM;
      const int buf = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < sh.M;
      const bool has_bias = epi.bias != nullptr;
      if (has_bias) {
        const int n_read = (sh.N + 7) & ~7;
        for (int i = threadIdx.x - 64; i < BN; i += kEpiWarps * 32)
          sts_f32(bias_sa + (uint32_t)((buf * BN + i) * 4), (n0 + i < n_read) ? epi.bias[n0 + i] : 0.f);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");
      }
      mbar_wait_bd(&tmem_full_bar[buf], acc_ph);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(quarter * 32) << 16);
      const uint32_t swz = (uint32_t)(lane & 7);
      const long add_row = epi.add_mod > 0 ? row % epi.add_mod : row;
      const bf16* add_p = (epi.addend && row_ok)
                              ? reinterpret_cast<const bf16*>(epi.addend) + bt * epi.batch_add + add_row * epi.ld_add + n0
                              : nullptr;
      const bf16* y_p = ((epi.act == ACT_DTANH || epi.act == ACT_DRELU) && row_ok)
                            ? reinterpret_cast<const bf16*>(epi.dact_y) + bt * epi.batch_y + (long)row * epi.ld_y + n0
                            : nullptr;
      const int do_tanh = (epi.act == ACT_TANH ? 1 : 0) | (epi.act == ACT_RELU ? 2 : 0) | (epi.act == ACT_DRELU ? 4 : 0);
      const float alpha = epi.alpha_dev ? __ldg(epi.alpha_dev) : 1.f;
      const bool f32_out = epi.out_f32 || epi.atomic;
      const int pw = f32_out ? 32 : 64;
      const int n_pass = BN / pw;
      const int my_last = ((n_pass - 1 - half) & ~1) + half;
      const uint32_t empty_leader = empty_leader0 + (uint32_t)buf * 8u;
      const bool warp_empty = (m0 + quarter * 32) >= sh.M;  // none of this warp's rows exists: nothing to read or store
      if (warp_empty) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(empty_leader);
      }
#pragma unroll 1
      for (int ps = warp_empty ? n_pass : half; ps < n_pass; ps += 2) {
        const int c0 = ps * pw;
        const int left = sh.N - (n0 + c0);
        // EPI2: two staging tiles per warp, used alternately: only the store before the previous one must have drained
        const uint32_t my_stage = staging_sa + (uint32_t)((EPI2 ? ew * kEpiDepth + (int)(stage_flip % kEpiDepth) : ew) * 4096);
        const uint32_t my_row = my_stage + (uint32_t)(lane * 128);
        if (lane == 0) bulk_wait_read<EPI2 ? kEpiDepth - 1 : 0>();  // the store that last used this staging tile has finished reading it
        __syncwarp();
        const float* bsp = has_bias ? bias_s + buf * BN + c0 : nullptr;                   // ACT_ROWSTATS path only
        const uint32_t bsa = has_bias ? bias_sa + (uint32_t)((buf * BN + c0) * 4) : 0u;
        const bf16* ap = add_p ? add_p + c0 : nullptr;
        const bf16* yp = y_p ? y_p + c0 : nullptr;
        if (epi.act == ACT_ROWSTATS) {  // pw == 64: one 64-column group per pass, nothing is stored
          float rm = -INFINITY, rs = 0.f;
          epilogue_stats32(tmem_d + (uint32_t)c0, bsp, left, rm, rs);
          epilogue_stats32(tmem_d + (uint32_t)(c0 + 32), bsp ? bsp + 32 : nullptr, left - 32, rm, rs);
          if (row_ok && left > 0) {
            const long o = (long)row * epi.stats_ld + ((n0 + c0) >> 6);
            epi.stats_m[o] = rm;
            epi.stats_s[o] = rs;
          }
          if (ps == my_last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(empty_leader);
          }
          continue;
        }
        if (f32_out) {
          if (left >= 32) epilogue_cols32<true, true>(tmem_d + (uint32_t)c0, bsa, ap, yp, do_tanh, alpha, left, my_row, swz, 0);
          else epilogue_cols32<true, false>(tmem_d + (uint32_t)c0, bsa, ap, yp, do_tanh, alpha, left, my_row, swz, 0);
        } else {
#pragma unroll 1
          for (int sub = 0; sub < 2; ++sub) {
            const int o = sub * 32;
            if (left - o >= 32)
              epilogue_cols32<false, true>(tmem_d + (uint32_t)(c0 + o), bsa ? bsa + (uint32_t)(o * 4) : 0u, ap ? ap + o : nullptr,
                                           yp ? yp + o : nullptr, do_tanh, alpha, left - o, my_row, swz, (uint32_t)(sub * 4));
            else
              epilogue_cols32<false, false>(tmem_d + (uint32_t)(c0 + o), bsa ? bsa + (uint32_t)(o * 4) : 0u, ap ? ap + o : nullptr,
                                            yp ? yp + o : nullptr, do_tanh, alpha, left - o, my_row, swz, (uint32_t)(sub * 4));
          }
        }
        if (ps == my_last) {
          // all of this warp's TMEM reads are done: hand the accumulator back to the leader's MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(empty_leader);
        }
        if (left <= 0) continue;
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          if (epi.atomic) tma_reduce_add_3d_s(&tmC, my_stage, n0 + c0, m0 + quarter * 32, bt);
          else tma_store_3d_s(&tmC, my_stage, n0 + c0, m0 + quarter * 32, bt);
          bulk_commit();
        }
        ++stage_flip;  // the staging tiles alternate per ISSUED store (bulk_wait_read<1> counts stores, not passes)
      }
    }
  }
  if (warp >= 2 && lane == 0) bulk_wait<0>();
  tc_fence_before();
  __syncthreads();
  cluster_barrier();  // the peer no longer signals my barriers / reads my shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, Cfg::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
unsigned long long* g_trace = nullptr;
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

struct MapKey {
  const void* p; uint64_t inner, outer, ld, nb, bs; uint32_t bi, bo; int esize, swizzle;
  bool operator==(const MapKey& o) const {
    return p == o.p && inner == o.inner && outer == o.outer && ld == o.ld && nb == o.nb && bs == o.bs && bi == o.bi &&
           bo == o.bo && esize == o.esize && swizzle == o.swizzle;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.p;
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.ld); mix(k.nb); mix(k.bs); mix(k.bi); mix(k.bo); mix((uint64_t)k.esize); mix((uint64_t)k.swizzle);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

// 3-D tensor [nb, outer, inner] of bf16 (esize 2) or fp32 (esize 4): row stride ld, batch stride bs (elements);
// box = [1, bo, bi]; 128-byte swizzle (or none: plain row-major box in shared memory).
int get_map(const void* p, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t bi, uint32_t bo, CUtensorMap* out,
            int esize, uint64_t nb, uint64_t bs, bool swizzle = true) {
  if (nb <= 1) { nb = 1; bs = ld * (outer > 0 ? outer : 1); }
  if ((bs * esize) & 15) {
    set_error("gemm_tc: batch stride %llu is not a multiple of 16 bytes", (unsigned long long)bs);
    return -1;
  }
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  if (((uintptr_t)p & 15) || ((ld * esize) & 15)) {
    set_error("gemm_tc: operand %p (ld=%llu) is not 16-byte aligned for TMA", p, (unsigned long long)ld);
    return -1;
  }
  MapKey key{p, inner, outer, ld, nb, bs, bi, bo, esize, swizzle ? 1 : 0};
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  cuuint64_t dims[3] = {inner, outer, nb};
  cuuint64_t strides[2] = {ld * (uint64_t)esize, bs * (uint64_t)esize};
  cuuint32_t box[3] = {bi, bo, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(p), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u", (int)r, p,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, bi, bo);
    return -1;
  }
  std::lock_guard<std::mutex> g(g_maps_mu);
  if (g_maps.size() > 4096) g_maps.clear();
  g_maps[key] = *out;
  return 0;
}

thread_local bool t_pdl_on = true;   // programmatic dependent launch; off while a CUDA graph is being captured
bool pdl_enabled() { return t_pdl_on; }

template <int BN, bool A_MN, bool B_MN, int MT>
int launch(GemmContext& ctx, const GemmCall& c, int splits, int kbps) {
  using Cfg = TcCfg<BN, MT>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmC;
  if (!A_MN) { if (get_map(c.A, c.K, c.M, c.lda, BK, BM, &tmA, 2, c.batch, c.batch_a)) return -1; }
  else { if (get_map(c.A, c.M, c.K, c.lda, 64, BK, &tmA, 2, c.batch, c.batch_a)) return -1; }
  if (!B_MN) { if (get_map(c.B, c.K, c.N, c.ldb, BK, BN, &tmB, 2, c.batch, c.batch_b)) return -1; }
  else { if (get_map(c.B, c.N, c.K, c.ldb, 64, BK, &tmB, 2, c.batch, c.batch_b)) return -1; }
  const GemmEpi& e = c.epi;
  const int out_elem = (e.out_f32 || e.atomic) ? 4 : 2;
  int vec_ok = ((c.ldc % 8) == 0) && (((uintptr_t)c.C) % (out_elem * 8) == 0);
  if (e.bias && ((uintptr_t)e.bias & 31)) vec_ok = 0;
  if (e.addend && (((uintptr_t)e.addend & 15) || (e.ld_add % 8))) vec_ok = 0;
  if (e.dact_y && (((uintptr_t)e.dact_y & 15) || (e.ld_y % 8))) vec_ok = 0;
  // output through TMA tensor stores when C is 16-byte aligned with a 16-byte-multiple pitch (always true for the
  // decoder's own buffers); otherwise the direct per-thread store path is used
  // TMA path: C addressable by a tensor map, aux operands 16/32-byte aligned; with aux operands the epilogue reads
  // them in groups of 8 columns, so N must be a multiple of 8 unless the caller vouches for padded buffers
  // (epi.aux_padded: the decoder's own buffers are padded to a multiple of 128 columns).
  const bool has_aux = e.bias || e.addend || e.dact_y;
  const int tma_out = (!has_aux || (c.N % 8) == 0 || e.aux_padded) && ((c.batch_c * out_elem) % 16 == 0) &&
                      (e.batch_add % 8 == 0) && (e.batch_y % 8 == 0) &&
                      (((uintptr_t)c.C & 15) == 0) && (((c.ldc * out_elem) & 15) == 0) &&
                      (!e.bias || ((uintptr_t)e.bias & 31) == 0) &&
                      (!e.addend || ((((uintptr_t)e.addend & 15) == 0) && (e.ld_add % 8) == 0)) &&
                      (!e.dact_y || ((((uintptr_t)e.dact_y & 15) == 0) && (e.ld_y % 8) == 0));
  // epilogue mode: 1 = swizzled smem tile + TMA tensor store / reduce-add, 0 = per-thread direct stores (unaligned C)
  const int out_mode = tma_out ? 1 : 0;
  if (out_mode == 1) {
    if (get_map(c.C, c.N, c.M, c.ldc, out_elem == 2 ? 64 : 32, 32, &tmC, out_elem, c.batch, c.batch_c)) return -1;
  } else {
    tmC = tmA;
  }
  TcShape sh;
  sh.tma_out = out_mode;
  sh.trace = g_trace;
  sh.M = c.M; sh.N = c.N; sh.K = c.K;
  sh.tiles_m = (c.M + BM * MT - 1) / (BM * MT); sh.tiles_n = (c.N + BN - 1) / BN;
  sh.splits = splits; sh.kb_per_split = kbps; sh.ldc = c.ldc; sh.vec_ok = vec_ok;
  sh.batch = c.batch > 0 ? c.batch : 1; sh.batch_c = c.batch_c;
  sh.rotate_k = 1;
  const long total = (long)sh.tiles_m * sh.tiles_n * splits * sh.batch;
  const int grid = (int)std::min<long>(total, kNumSMs);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = ctx.stream;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = &attr;
  cfg.numAttrs = 1;
  void* Cp = c.C;
  cudaError_t err = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, A_MN, B_MN, MT>, tmA, tmB, tmC, Cp, sh, c.epi);
  ctx.launches++;
  if (err != cudaSuccess) { set_error("gemm_tc launch: %s", cudaGetErrorString(err)); return -1; }
  return 0;
}

template <int BN, int MT> int launch_majors(GemmContext& ctx, const GemmCall& c, int splits, int kbps) {
  if (c.a_mn)
    return c.b_mn ? launch<BN, true, true, MT>(ctx, c, splits, kbps) : launch<BN, true, false, MT>(ctx, c, splits, kbps);
  return c.b_mn ? launch<BN, false, true, MT>(ctx, c, splits, kbps) : launch<BN, false, false, MT>(ctx, c, splits, kbps);
}

template <bool A_MN, bool B_MN, bool EPI2>
int launch_pair(GemmContext& ctx, const GemmCall& c, int splits, int kbps) {
  using Cfg = PairCfgT<EPI2>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_pair_kernel<A_MN, B_MN, EPI2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("gemm_tc pair: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmC;
  if (!A_MN) { if (get_map(c.A, c.K, c.M, c.lda, BK, BM, &tmA, 2, c.batch, c.batch_a)) return -1; }
  else { if (get_map(c.A, c.M, c.K, c.lda, 64, BK, &tmA, 2, c.batch, c.batch_a)) return -1; }
  if (!B_MN) { if (get_map(c.B, c.K, c.N, c.ldb, BK, kPairBN / 2, &tmB, 2, c.batch, c.batch_b)) return -1; }
  else { if (get_map(c.B, c.N, c.K, c.ldb, 64, BK, &tmB, 2, c.batch, c.batch_b)) return -1; }
  const GemmEpi& e = c.epi;
  const int out_elem = (e.out_f32 || e.atomic) ? 4 : 2;
  if (get_map(c.C, c.N, c.M, c.ldc, out_elem == 2 ? 64 : 32, 32, &tmC, out_elem, c.batch, c.batch_c)) return -1;
  TcShape sh;
  sh.tma_out = 1;
  sh.trace = nullptr;
  sh.M = c.M; sh.N = c.N; sh.K = c.K;
  sh.tiles_m = (c.M + 2 * BM - 1) / (2 * BM); sh.tiles_n = (c.N + kPairBN - 1) / kPairBN;
  sh.splits = splits; sh.kb_per_split = kbps; sh.ldc = c.ldc; sh.vec_ok = 1;
  sh.batch = c.batch > 0 ? c.batch : 1; sh.batch_c = c.batch_c;
  sh.rotate_k = 0;
  const long total = (long)sh.tiles_m * sh.tiles_n * splits * sh.batch;
  const int pairs = (int)std::min<long>(total, kNumSMs / 2);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = ctx.stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  cudaError_t err = cudaLaunchKernelEx(&cfg, gemm_tc_pair_kernel<A_MN, B_MN, EPI2>, tmA, tmB, tmC, sh, c.epi);
  ctx.launches++;
  if (err != cudaSuccess) { set_error("gemm_tc pair launch: %s", cudaGetErrorString(err)); return -1; }
  return 0;
}

// the pair kernel writes through TMA only: same alignment conditions as the single-CTA TMA epilogue
bool pair_eligible(const GemmCall& c) {
  const GemmEpi& e = c.epi;
  const int out_elem = (e.out_f32 || e.atomic) ? 4 : 2;
  const bool has_aux = e.bias || e.addend || e.dact_y;
  return c.M > BM && c.N > kPairBN / 2 && (!has_aux || (c.N % 8) == 0 || e.aux_padded) &&
         ((c.batch_c * out_elem) % 16 == 0) && (e.batch_add % 8 == 0) && (e.batch_y % 8 == 0) &&
         (((uintptr_t)c.C & 15) == 0) && (((c.ldc * out_elem) & 15) == 0) && (!e.bias || ((uintptr_t)e.bias & 31) == 0) &&
         (!e.addend || ((((uintptr_t)e.addend & 15) == 0) && (e.ld_add % 8) == 0)) &&
         (!e.dact_y || ((((uintptr_t)e.dact_y & 15) == 0) && (e.ld_y % 8) == 0));
}

}  // namespace

void set_gemm_trace(unsigned long long* buf) { g_trace = buf; }
void set_gemm_pdl(bool on) { t_pdl_on = on; }
bool gemm_pdl() { return t_pdl_on; }

// the tensor-map cache, for the other TMA kernels of the library (recurrent.cu)
int tma_map_3d(const void* p, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t bi, uint32_t bo, CUtensorMap* out,
               int esize, uint64_t nb, uint64_t bs, bool swizzle) {
  return get_map(p, inner, outer, ld, bi, bo, out, esize, nb, bs, swizzle);
}

int gemm_bf16_tc(GemmContext& ctx, const GemmCall& c) {
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) return 0;
  if (c.epi.act == ACT_ROWSTATS) {
    const GemmEpi& e = c.epi;
    const bool ok = !e.out_f32 && !e.atomic && c.batch <= 1 && !e.addend && !e.dact_y && (!e.bias || e.aux_padded) &&
                    (((uintptr_t)c.C & 15) == 0) && ((c.ldc * 2) % 16 == 0) && e.stats_m && e.stats_s &&
                    e.stats_ld >= (c.N + 63) / 64;
    if (!ok) { set_error("gemm_tc: unsupported arguments for the row-statistics epilogue"); return -1; }
  }
  if (c.splits > 1 && !c.epi.atomic) { set_error("gemm_tc: split-K needs an atomic epilogue"); return -1; }
  if (c.batch > 1 && c.epi.atomic) { set_error("gemm_tc: batched GEMM with split-K is not supported"); return -1; }
  const int nkb = (c.K + BK - 1) / BK;
  // Pick the tile (rows 128 or 256, width 64 / 128 / 256) and the k-split that minimise a simple cost model of the
  // persistent kernel:
  //   per CTA: ceil(work / 148) items, each  kb_per_split * t_kblock + t_epilogue;
  //   t_kblock ~ operand bytes moved per k-block (the L2 -> smem stream bounds the mainloop), epilogue ~ tile area
  //   (x3 for reductions).
  int best_bn = 64, best_sp = 1;
  double best_cost = 1e300;
  const int cands[3] = {256, 128, 64};
  // (A 256-row single-CTA tile was measured and dropped: head fc 25.8 vs 25.5 us, key/value projection 43.9 vs 32.1 us --
  // the mainloop is bound by the shared-memory operand reads of the UMMAs, which a taller tile does not reduce.)
  {
    const int mt = (c.M + BM - 1) / BM;
    for (int i = 0; i < 3; ++i) {
      const int bn = cands[i];
      if (bn > 64 && c.N <= bn / 2) continue;
      const long tiles = (long)mt * ((c.N + bn - 1) / bn) * (c.batch > 0 ? c.batch : 1);
      const int sp_max = c.epi.atomic ? (c.splits > 0 ? c.splits : std::max(1, std::min(nkb / 4, 16))) : 1;
      const int sp_min = (c.epi.atomic && c.splits > 0) ? c.splits : 1;
      for (int sp = sp_min; sp <= sp_max; ++sp) {
        const int kbps = (nkb + sp - 1) / sp;
        const int sp_eff = (nkb + kbps - 1) / kbps;
        const long per_cta = (tiles * sp_eff + kNumSMs - 1) / kNumSMs;
        const double t_kb = (128.0 + bn) / 384.0;
        const double t_epi = (bn / 256.0) * (c.epi.atomic ? 6.0 : 2.0) + 1.0;
        const double cost = per_cta * (kbps * t_kb + t_epi) + 4.0 /*fill*/;
        if (cost < best_cost) { best_cost = cost; best_bn = bn; best_sp = sp_eff; }
      }
    }
  }
  // CTA-pair 256 x 256 tiles (cta_group::2): 74 pairs, a k-block costs ~0.27 us instead of 0.37 us (see the kernel).
  // ctx.variant (kernel-level tests through capdec_gemm): GEMM_PAIR / GEMM_SINGLE force the choice.
  bool use_pair = false;
  int pair_sp = 1;
  const long pair_tiles = (long)((c.M + 2 * BM - 1) / (2 * BM)) * ((c.N + kPairBN - 1) / kPairBN) * (c.batch > 0 ? c.batch : 1);
  if (ctx.variant != GEMM_SINGLE && pair_eligible(c)) {
    const long tiles = pair_tiles;
    const int sp_max = c.epi.atomic ? (c.splits > 0 ? c.splits : std::max(1, std::min(nkb / 4, 16))) : 1;
    const int sp_min = (c.epi.atomic && c.splits > 0) ? c.splits : 1;
    double pair_cost = 1e300;
    for (int sp = sp_min; sp <= sp_max; ++sp) {
      const int kbps = (nkb + sp - 1) / sp;
      const int sp_eff = (nkb + kbps - 1) / kbps;
      const long per_pair = (tiles * sp_eff + kNumSMs / 2 - 1) / (kNumSMs / 2);
      const double t_epi = (c.epi.atomic ? 6.0 : 2.0) + 1.0;
      const double cost = per_pair * (kbps * 0.75 + t_epi) + 5.0;
      if (cost < pair_cost) { pair_cost = cost; pair_sp = sp_eff; }
    }
    use_pair = ctx.variant == GEMM_PAIR || pair_cost < best_cost;
  }
  if (use_pair) {
    const int kbps2 = (nkb + pair_sp - 1) / pair_sp;
    const int splits2 = (nkb + kbps2 - 1) / kbps2;
    // short-K forward-form problems with many tiles per pair are epilogue-bound: the variant with two staging tiles per warp
    const bool epi2 = !c.a_mn && splits2 == 1 && kbps2 <= 8 && pair_tiles >= 3 * (kNumSMs / 2);
    if (epi2) return c.b_mn ? launch_pair<false, true, true>(ctx, c, splits2, kbps2) : launch_pair<false, false, true>(ctx, c, splits2, kbps2);
    if (c.a_mn) return c.b_mn ? launch_pair<true, true, false>(ctx, c, splits2, kbps2) : launch_pair<true, false, false>(ctx, c, splits2, kbps2);
    return c.b_mn ? launch_pair<false, true, false>(ctx, c, splits2, kbps2) : launch_pair<false, false, false>(ctx, c, splits2, kbps2);
  }
  const int kbps = (nkb + best_sp - 1) / best_sp;
  const int splits = (nkb + kbps - 1) / kbps;
  switch (best_bn) {
    case 256: return launch_majors<256, 1>(ctx, c, splits, kbps);
    case 128: return launch_majors<128, 1>(ctx, c, splits, kbps);
    default: return launch_majors<64, 1>(ctx, c, splits, kbps);
  }
}

}  // namespace capdec
