// The recurrent core of the teacher-forced decoder as TWO persistent kernels (one forward, one backward) instead of
// 34 x [gate GEMM + cell] launches in each direction.
//
// Replaces the sequential part of the PaddlePaddle While-op body, model_adaAttention_aic.py:86-88, 94-95
// (h_{t-1}.W_hh + lstm_unit) and its while_grad.  Everything that does not depend on h_{t-1} was hoisted out of the
// loop by the schedule in decoder.cu, so per step only  gates = h_{t-1}.W_hh  and the cell remain.
// Launched per step, that GEMM re-streams the 8 MB weight through every SM and pays a launch + pipeline fill for
// 2 GFLOP of work; here the weight slice of a CTA stays resident in shared memory for all T steps, the CTAs are
// co-resident (cooperative launch, <= 148 CTAs) and meet at a release/acquire counter in global memory between
// steps.
//
//   forward : CTA (j, y) owns hidden units [16j, 16j+16) -- 64 gate columns, unit-major (col = 4*unit + gate) in a
//             permuted weight copy -- for the 128-row batch tiles y, y+gridDim.y, ...   Per step: TMA streams the
//             h_{t-1} tile (K-major A), tcgen05.mma against the resident [H x 64] weight slice (MN-major B),
//             accumulator in TMEM, and the epilogue warps apply the cell (i,f,o,g; c,h) straight from TMEM and write
//             h_t (bf16), c_t (fp32) and the gate activations.  One barrier per step.
//   backward: d h_{t-1} = d gates_t . W_hh^T contracts over 4H, so a column partition would stream 1 MB per CTA and
//             step; instead CTA (n, ks) owns a [128 units x 512 k] slice of W_hh (K-major B, 128 KB resident), adds its
//             partial product into a ping-pong fp32 d h buffer by TMA reduce-add, and the cell backward runs as a
//             second phase spread over all CTAs.  Two barriers per step.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "gemm.cuh"
#include "kernels.cuh"

namespace capdec {

// gemm_tc.cu: cached cuTensorMapEncodeTiled wrapper, 3-D tensor {inner, outer, batch}, SWIZZLE_128B
int tma_map_3d(const void* p, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t bi, uint32_t bo, CUtensorMap* out,
               int esize, uint64_t nb, uint64_t bs, bool swizzle = true);

namespace {

constexpr int kThreads = 192;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2..5: epilogue / cell
constexpr int kEpiThreads = 128;
constexpr int BK = 64;
constexpr int kATile = 128 * BK * 2;  // 16 KB: 128 rows x 64 k, bf16
constexpr unsigned kSpinLimit = 1u << 23;

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// wait until a monotonic arrival counter reaches `target`; a lost arrival traps instead of hanging the GPU
__device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target) {
  unsigned n = 0;
  while (ld_acquire_gpu(ctr) < target) {
    if (++n > kSpinLimit) __trap();
  }
}
// bounded mbarrier wait: a protocol bug traps (the launch fails loudly) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_b(uint64_t* bar, uint32_t parity) {
  unsigned n = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++n > (1u << 22)) __trap();
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

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
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// UMMA shared-memory descriptors, SWIZZLE_128B (same encoding as gemm_tc.cu)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t tile, int k) { return smem_desc(tile + k * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mnmajor(uint32_t tile, int k) { return smem_desc(tile + k * 2048, 8192, 1024); }

// optional per-step timeline (capdec_debug_rec_trace): 8 globaltimer slots per (CTA, step)
__device__ __forceinline__ void rec_trace(unsigned long long* tr, int T, int t, int slot) {
  if (!tr) return;
  unsigned long long now;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
  tr[(((size_t)blockIdx.y * gridDim.x + blockIdx.x) * T + t) * 8 + slot] = now;
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

struct Bars {
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tmem_full;
  uint64_t* tmem_empty;
  uint64_t* w_bar;
  uint64_t* go;        // forward: the producer has passed the inter-CTA barrier of this step
  uint32_t* tmem_ptr;
};
template <int kStages> __device__ __forceinline__ Bars carve_bars(uint8_t* p) {
  Bars b;
  b.full = reinterpret_cast<uint64_t*>(p);
  b.empty = b.full + kStages;
  b.tmem_full = b.empty + kStages;
  b.tmem_empty = b.tmem_full + 2;
  b.w_bar = b.tmem_empty + 2;
  b.go = b.w_bar + 1;
  b.tmem_ptr = reinterpret_cast<uint32_t*>(b.go + 1);
  return b;
}
template <int kStages>
__device__ __forceinline__ void init_bars(const Bars& b, uint32_t epi_arrivals) {
  for (int s = 0; s < kStages; ++s) {
    mbar_init(&b.full[s], 1);
    mbar_init(&b.empty[s], 1);
  }
  for (int i = 0; i < 2; ++i) {
    mbar_init(&b.tmem_full[i], 1);
    mbar_init(&b.tmem_empty[i], epi_arrivals);
  }
  mbar_init(b.w_bar, 1);
  mbar_init(b.go, 1);
  mbar_fence_init();
}

// ================================================================================================================
// forward
// ================================================================================================================
#ifndef CAPDEC_FWD_STAGES
#define CAPDEC_FWD_STAGES 4   // h-tile ring depth (build-time: profiling experiments vary it)
#endif
constexpr int kFwdStages = CAPDEC_FWD_STAGES;
constexpr int kFwdEpiWarps = 8;   // two per TMEM lane quarter: 8 of the CTA's 16 hidden units each
constexpr int kFwdEpiThreads = kFwdEpiWarps * 32;
constexpr int kFwdThreads = 64 + kFwdEpiThreads;
__device__ __forceinline__ void fwd_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kFwdEpiThreads) : "memory"); }
constexpr int kFwdUnits = 16;            // hidden units per CTA
constexpr int kFwdN = 4 * kFwdUnits;     // 64 gate columns
constexpr int kFwdWTile = BK * kFwdN * 2;  // 8 KB per k-block of the resident weight slice
constexpr int kFwdCStage = 128 * kFwdUnits * 4;  // 8 KB: c_t tile, 128 rows x 64 B, row-major
constexpr int kFwdXStage = 128 * kFwdN * 2;      // 16 KB: 128 rows x 128 B, SWIZZLE_128B: x-part pre-activations in,
                                                 //        gate activations out (same tile, one after the other)

struct SeqFwdParams {
  int B, H, T;
  int n_mtiles;
  bf16* hs;                     // [(T+1)*B, H]  hs[0] = h_0 (zeros)
  const float* cs;              // [(T+1)*B, H]  cs[0] = c_0 (zeros); written through the tensor map
  unsigned* ctr;                // [gridDim.y]   arrival counters (zeroed by the host)
  unsigned long long* trace;    // debug timeline or nullptr
};

// Everything except h_t moves through the TMA engine in whole 128-byte lines: the per-step x-part pre-activations
// (word-embedding and global-image terms and biases, summed by the producing GEMM; unit-major columns 4*unit + gate
// like the weight slice) come in as one [128 x 64] tile, c_t and the gate activations go out as two tensor stores.
// Per-lane global loads (every lane = another row = another line) were measured to cost ~3000 L1 wavefronts and
// ~400 K sector requests per step, which either delayed the producer's polls of the arrival counter (~3 us per step)
// or, when issued after the barrier, slowed the h tile stream by 70 %.
//
// (Measured and dropped: clusters of 2 / 4 CTAs sharing a batch tile, each loading 128/CS rows of every h k-block and
// multicasting them -- 336 / 336 / 334 us for CS = 1 / 2 / 4: the L2 already merges the unicast requests of a few CTAs for
// the same line; clusters of 8 with 217 KB per CTA are not co-schedulable.  Staggering the two batch groups: no effect.)
__global__ void __launch_bounds__(kFwdThreads, 1)
lstm_seq_fwd_kernel(const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmCs,
                    const __grid_constant__ CUtensorMap tmGa, SeqFwdParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t misalign = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + misalign;
  const int nkb = p.H / BK;
  uint8_t* wsm = smem;                                  // nkb x 8 KB   resident weight slice
  uint8_t* stages = wsm + (size_t)nkb * kFwdWTile;      // kFwdStages x 16 KB   h tiles
  uint8_t* x_stage = stages + kFwdStages * kATile;      // 16 KB (1024-aligned)
  uint8_t* c_stage = x_stage + kFwdXStage;              // 8 KB
  const Bars bar = carve_bars<kFwdStages>(c_stage + kFwdCStage);
  uint64_t* x_full = bar.go;                            // the x tile of the next batch tile has landed

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x, y = blockIdx.y, G = gridDim.x, MG = gridDim.y;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmH);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmCs);
    tma_prefetch_desc(&tmGa);
    init_bars<kFwdStages>(bar, kFwdEpiWarps);
  }
  if (warp == 1) tmem_alloc<2 * kFwdN>(bar.tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *bar.tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_expect_tx(bar.w_bar, (uint32_t)(nkb * kFwdWTile));
      for (int kb = 0; kb < nkb; ++kb) tma_load_3d(wsm + (size_t)kb * kFwdWTile, &tmW, slice * kFwdN, kb * BK, 0, bar.w_bar);
      int s = 0;
      uint32_t ph = 0;
      for (int t = 0; t < p.T; ++t) {
        if (t > 0) {
          spin_until(p.ctr + y, (unsigned)t * (unsigned)G);  // every CTA of this batch group has written its h_t slice
          fence_proxy_async_all();                            // generic-proxy writes -> async-proxy (TMA) reads
        }
        rec_trace(p.trace, p.T, t, 0);
        for (int mt = y; mt < p.n_mtiles; mt += MG) {
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait_b(&bar.empty[s], ph ^ 1);
            mbar_expect_tx(&bar.full[s], kATile);
            tma_load_3d(stages + s * kATile, &tmH, kb * BK, mt * 128, t, &bar.full[s]);
            if (++s == kFwdStages) { s = 0; ph ^= 1; }
          }
        }
        rec_trace(p.trace, p.T, t, 1);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====  D[128 x 64] += h_tile[128 x 64k] . W_slice[64k x 64]
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) /*B MN-major*/ |
                                 ((uint32_t)(kFwdN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      mbar_wait_b(bar.w_bar, 0);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int t = 0; t < p.T; ++t) {
        for (int mt = y; mt < p.n_mtiles; mt += MG, ++it) {
          const int buf = it & 1;
          mbar_wait_b(&bar.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kFwdN);
          for (int kb = 0; kb < nkb; ++kb) {
            mbar_wait_b(&bar.full[s], ph);
            if (kb == 0) rec_trace(p.trace, p.T, t, 2);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(stages + s * kATile);
            const uint32_t b_addr = smem_u32(wsm + (size_t)kb * kFwdWTile);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, desc_kmajor(a_addr, k), desc_mnmajor(b_addr, k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bar.empty[s]);
            if (++s == kFwdStages) { s = 0; ph ^= 1; }
          }
          umma_commit(&bar.tmem_full[buf]);
          rec_trace(p.trace, p.T, t, 3);
        }
      }
    }
  } else {
    // ===== epilogue: the LSTM cell applied to the accumulator (Paddle lstm_unit, gates i,f,o,g) =====
    // 8 warps: warp (quarter, hh) owns TMEM lanes [32*quarter, +32) and accumulator columns [32*hh, +32) = 8 units
    const int quarter = warp & 3;
    const int hh = (warp - 2) >> 2;
    const int et = threadIdx.x - 64;  // 0..255
    const int u0 = slice * kFwdUnits;
    const long BH = (long)p.B * p.H;
    const int trow = quarter * 32 + lane;  // row of the 128-row tile = TMEM lane
    uint8_t* x_row = x_stage + trow * 128;
    uint8_t* c_row = c_stage + trow * 64 + hh * 32;
    const uint32_t swz = (uint32_t)(lane & 7);
    // c_{t-1} of this thread's row: with one batch tile per CTA it is still in the c staging tile of the previous
    // step; otherwise it is read back from global memory once the tensor stores that wrote it have completed
    const bool single_tile = (MG >= p.n_mtiles);
    if (et == 0) {  // x tile of the first (step, batch tile)
      mbar_expect_tx(x_full, kFwdXStage);
      tma_load_3d(x_stage, &tmX, 4 * u0, y * 128, 0, x_full);
    }
    int it = 0;
    for (int t = 0; t < p.T; ++t) {
      if (t > 0 && !single_tile) {
        if (et == 0) {
          bulk_wait_all();
          fence_proxy_async_all();
        }
        fwd_bar_sync();
      }
      for (int mt = y; mt < p.n_mtiles; mt += MG, ++it) {
        const int buf = it & 1;
        const int row = mt * 128 + trow;
        const bool row_ok = row < p.B;
        float cp[8];
        if (t > 0 && single_tile) {
          load8(reinterpret_cast<const float*>(c_row), cp);
        } else if (row_ok) {
          load8(p.cs + (long)t * BH + (long)row * p.H + u0 + hh * 8, cp);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) cp[i] = 0.f;
        }
        // x-part pre-activations of this thread's 8 units: shared memory -> registers (bf16, (i,f,o,g) per unit)
        uint4 xq[4];
        mbar_wait_b(x_full, (uint32_t)(it & 1));
#pragma unroll
        for (int q = 0; q < 4; ++q) xq[q] = *reinterpret_cast<const uint4*>(x_row + (((uint32_t)(hh * 4 + q)) ^ swz) * 16);
        mbar_wait_b(&bar.tmem_full[buf], (it >> 1) & 1);
        if (et == 0) rec_trace(p.trace, p.T, t, 4);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kFwdN + hh * 32) + ((uint32_t)(quarter * 32) << 16);
        // every thread has read its x chunks (and its c values): the tile may now be overwritten with the gate activations
        fwd_bar_sync();
        {
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(tmem_d, r);
          tmem_ld_wait();
          float hn[8], cn[8];
          uint4 gq[4];  // gate activations of units 2q, 2q+1 as bf16 (i,f,o,g each)
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const __nv_bfloat162* xp = reinterpret_cast<const __nv_bfloat162*>(&xq[u >> 1]) + (u & 1) * 2;
            const float2 x_if = __bfloat1622float2(xp[0]), x_og = __bfloat1622float2(xp[1]);
            const float ig = sigmoid_t<true>(__uint_as_float(r[4 * u + 0]) + x_if.x);
            const float fg = sigmoid_t<true>(__uint_as_float(r[4 * u + 1]) + x_if.y);
            const float og = sigmoid_t<true>(__uint_as_float(r[4 * u + 2]) + x_og.x);
            const float gg = tanh_t<true>(__uint_as_float(r[4 * u + 3]) + x_og.y);
            cn[u] = fg * cp[u] + ig * gg;
            hn[u] = og * tanh_t<true>(cn[u]);
            __nv_bfloat162* gp = reinterpret_cast<__nv_bfloat162*>(&gq[u >> 1]) + (u & 1) * 2;
            gp[0] = __floats2bfloat162_rn(ig, fg);
            gp[1] = __floats2bfloat162_rn(og, gg);
          }
          if (row_ok) store8(p.hs + (long)(t + 1) * BH + (long)row * p.H + u0 + hh * 8, hn);
          // rows past the batch are clipped by the tensor stores
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(x_row + (((uint32_t)(hh * 4 + q)) ^ swz) * 16) = gq[q];
          *reinterpret_cast<float4*>(c_row) = make_float4(cn[0], cn[1], cn[2], cn[3]);
          *reinterpret_cast<float4*>(c_row + 16) = make_float4(cn[4], cn[5], cn[6], cn[7]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar.tmem_empty[buf]);
        const bool last_tile = (mt + MG >= p.n_mtiles);
        const bool release = (t + 1 < p.T) && last_tile;
        if (t + 1 < p.T) __threadfence();  // this thread's h_t stores are performed at GPU scope
        fence_proxy_async_all();            // staging writes -> async proxy
        if (et == 0) rec_trace(p.trace, p.T, t, 5);
        fwd_bar_sync();
        if (et == 0) {
          if (release) {
            // release h_t of this CTA's slice to the other CTAs of the batch group
            rec_trace(p.trace, p.T, t, 6);
            __threadfence();
            atomicAdd(p.ctr + y, 1u);
            rec_trace(p.trace, p.T, t, 7);
          }
          tma_store_3d(&tmCs, c_stage, u0, mt * 128, t + 1);
          tma_store_3d(&tmGa, x_stage, 4 * u0, mt * 128, t);
          bulk_commit();
          // next (step, batch tile) of this CTA: its x tile goes into the same buffer once the store has read it
          int nmt = mt + MG, nt = t;
          if (nmt >= p.n_mtiles) { nmt = y; nt = t + 1; }
          if (nt < p.T) {
            bulk_wait_read_all();
            mbar_expect_tx(x_full, kFwdXStage);
            tma_load_3d(x_stage, &tmX, 4 * u0, nmt * 128, nt, x_full);
          }
        }
        // the c / gate staging tiles are rewritten only after the next tile's fwd_bar_sync, which thread 0 reaches
        // after the bulk_wait_read above
      }
    }
    if (et == 0) bulk_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * kFwdN>(tmem_base);
  }
}

// ================================================================================================================
// backward
// ================================================================================================================
#ifndef CAPDEC_BWD_STAGES
#define CAPDEC_BWD_STAGES 4
#endif
constexpr int kBwdStages = CAPDEC_BWD_STAGES;
constexpr int kBwdN = 128;                  // output hidden units per CTA
constexpr int kBwdKb = 8;                   // k-blocks per CTA (512 of the 4H gate columns)
constexpr int kBwdWTile = kBwdN * BK * 2;   // 16 KB
constexpr int kBwdEpiWarps = 8;             // two per TMEM lane quarter: half of the accumulator columns each
constexpr int kBwdEpiThreads = kBwdEpiWarps * 32;
constexpr int kBwdThreads = 64 + kBwdEpiThreads;
__device__ __forceinline__ void bwd_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kBwdEpiThreads) : "memory"); }

struct SeqBwdParams {
  int B, H, T;
  int n_mtiles;
  const bf16* dHhead;    // [T*B, H]   d loss / d h_t from the head (p_hid path)
  const bf16* dHprev;    // [T*B, H]   row block t: d loss / d h_{t-1} from step t's p_hidden
  const float* dChead;   // [T*B, H]   d loss / d c_t from the sentinel
  const bf16* gact;      // [T*B, 4H]   unit-major (col 4*unit + gate), as written by lstm_seq_fwd_kernel
  const float* cs;       // [(T+1)*B, H]
  bf16* dG; long ld_dg;  // [T*B, ld_dg] cols gate*H + unit
  float* dhrec;          // [2, B, H] ping-pong d h through the gates (zeroed by the host)
  float* dcrec;          // [B, H]    running d c (zeroed by the host)
  unsigned* ctr_c;       // [gridDim.y] cell phases completed
  unsigned* ctr_g;       // [gridDim.y] GEMM phases completed
  unsigned long long* trace;
};

// ---- one BPTT step of the cell for 8 hidden units of one row ("item") ----
// Raw operands of an item; loads of several items are issued back to back before any arithmetic.
struct CellItem {
  long m;        // row of the [T*B, *] activations, or -1 for "no item"
  int row, h0;
  uint4 dhh, dhp, ga[4];   // bf16 x 8: d h (head), d h_{t-1} (p_hidden path, step t+1); gates: 8 units x (i,f,o,g)
  float4 dch[2], dcr[2], ct[2], cp[2], dhr[2];    // fp32 x 8: d c (head), running d c, c_t, c_{t-1}, d h through the gates
};
__device__ __forceinline__ void bf8_to_f(const uint4& u, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void f4x2_to_f(const float4 (&a)[2], float (&v)[8]) {
  v[0] = a[0].x; v[1] = a[0].y; v[2] = a[0].z; v[3] = a[0].w; v[4] = a[1].x; v[5] = a[1].y; v[6] = a[1].z; v[7] = a[1].w;
}
__device__ __forceinline__ void cell_load(const SeqBwdParams& p, int t, int mt, int i, bool last, CellItem& it) {
  const int vec_per_row = p.H / 8;
  it.m = -1;
  if (i >= 128 * vec_per_row) return;
  it.row = mt * 128 + i / vec_per_row;
  if (it.row >= p.B) return;
  it.h0 = (i % vec_per_row) * 8;
  const long BH = (long)p.B * p.H;
  it.m = (long)t * p.B + it.row;
  const long o = (long)it.row * p.H + it.h0;
  it.dhh = *reinterpret_cast<const uint4*>(p.dHhead + it.m * p.H + it.h0);
#pragma unroll
  for (int q = 0; q < 4; ++q) it.ga[q] = *reinterpret_cast<const uint4*>(p.gact + it.m * 4 * p.H + 4 * it.h0 + 8 * q);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    it.dch[q] = *reinterpret_cast<const float4*>(p.dChead + it.m * p.H + it.h0 + 4 * q);
    it.dcr[q] = *reinterpret_cast<const float4*>(p.dcrec + o + 4 * q);
    it.ct[q] = *reinterpret_cast<const float4*>(p.cs + (long)(t + 1) * BH + o + 4 * q);
    it.cp[q] = *reinterpret_cast<const float4*>(p.cs + (long)t * BH + o + 4 * q);
  }
  if (!last) {
    it.dhp = *reinterpret_cast<const uint4*>(p.dHprev + (it.m + p.B) * p.H + it.h0);
    const float* hr = p.dhrec + (long)(t & 1) * BH + o;
    it.dhr[0] = ldcg4(hr);  // written by other SMs (TMA reduce-add): bypass L1
    it.dhr[1] = ldcg4(hr + 4);
  }
}
// L2 prefetch of the operands of item i at step t that come from HBM (everything except the running state): issued while
// the epilogue warps wait for the GEMM phase of step t+1, so that the cell phase finds them in L2 (no register cost)
__device__ __forceinline__ void cell_prefetch(const SeqBwdParams& p, int t, int mt, int i) {
  const int vec_per_row = p.H / 8;
  if (i >= 128 * vec_per_row) return;
  const int row = mt * 128 + i / vec_per_row;
  if (row >= p.B) return;
  const int h0 = (i % vec_per_row) * 8;
  const long BH = (long)p.B * p.H;
  const long m = (long)t * p.B + row;
  const long o = (long)row * p.H + h0;
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p.gact + m * 4 * p.H + 4 * h0));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dHhead + m * p.H + h0));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dChead + m * p.H + h0));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p.cs + (long)t * BH + o));   // c_{t-1}; c_t was this step's c_{t-1}
  if (t + 1 < p.T) asm volatile("prefetch.global.L2 [%0];" ::"l"(p.dHprev + (m + p.B) * p.H + h0));
}

// d gates -> dG (needed by the other CTAs before the GEMM phase); the new running d c is left in it.dcr
__device__ __forceinline__ void cell_compute_store_dg(const SeqBwdParams& p, bool last, CellItem& it) {
  if (it.m < 0) return;
  float dh[8], x[8], dc[8], dcr[8], ct[8], cprev[8], ga[4][8], dg[4][8];
  bf8_to_f(it.dhh, dh);
  f4x2_to_f(it.dch, dc);
  f4x2_to_f(it.dcr, dcr);
  f4x2_to_f(it.ct, ct);
  f4x2_to_f(it.cp, cprev);
#pragma unroll
  for (int q = 0; q < 4; ++q) {  // chunk q = units 2q, 2q+1, four gates each
    float v[8];
    bf8_to_f(it.ga[q], v);
#pragma unroll
    for (int g = 0; g < 4; ++g) { ga[g][2 * q] = v[g]; ga[g][2 * q + 1] = v[4 + g]; }
  }
  if (!last) {
    float r[8];
    bf8_to_f(it.dhp, x);
    f4x2_to_f(it.dhr, r);
#pragma unroll
    for (int j = 0; j < 8; ++j) dh[j] += x[j] + r[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float ig = ga[0][j], fg = ga[1][j], og = ga[2][j], gg = ga[3][j];
    const float tc = tanh_t<true>(ct[j]);
    const float dcj = dc[j] + dcr[j] + dh[j] * og * (1.f - tc * tc);
    dg[0][j] = dcj * gg * ig * (1.f - ig);
    dg[1][j] = dcj * cprev[j] * fg * (1.f - fg);
    dg[2][j] = dh[j] * tc * og * (1.f - og);
    dg[3][j] = dcj * ig * (1.f - gg * gg);
    dcr[j] = dcj * fg;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) store8(p.dG + it.m * p.ld_dg + (long)g * p.H + it.h0, dg[g]);
  it.dcr[0] = make_float4(dcr[0], dcr[1], dcr[2], dcr[3]);
  it.dcr[1] = make_float4(dcr[4], dcr[5], dcr[6], dcr[7]);
}
// state that only this thread reads again: the running d c, and the consumed d h slot (zeroed for the reduce-adds of
// two steps from now)
__device__ __forceinline__ void cell_store_state(const SeqBwdParams& p, int t, bool last, const CellItem& it) {
  if (it.m < 0) return;
  const long o = (long)it.row * p.H + it.h0;
  *reinterpret_cast<float4*>(p.dcrec + o) = it.dcr[0];
  *reinterpret_cast<float4*>(p.dcrec + o + 4) = it.dcr[1];
  if (!last) {
    float* hr = p.dhrec + (long)(t & 1) * (long)p.B * p.H + o;
    *reinterpret_cast<float4*>(hr) = make_float4(0.f, 0.f, 0.f, 0.f);
    *reinterpret_cast<float4*>(hr + 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

__global__ void __launch_bounds__(kBwdThreads, 1)
lstm_seq_bwd_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmW,
                    const __grid_constant__ CUtensorMap tmC, SeqBwdParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t misalign = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + misalign;
  uint8_t* wsm = smem;                                   // 8 x 16 KB   resident W_hh slice [128 units x 512 k]
  uint8_t* stages = wsm + kBwdKb * kBwdWTile;            // kBwdStages x 16 KB   d gates tiles
  uint8_t* staging = stages + kBwdStages * kATile;       // 8 warps x 4 KB  epilogue tiles (32 rows x 128 B)
  const Bars bar = carve_bars<kBwdStages>(staging + kBwdEpiWarps * 4096);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KS = p.H / 128;  // = 4H / 512 k-slices, and H / 128 unit slices
  const int cta = blockIdx.x, y = blockIdx.y, G = gridDim.x, MG = gridDim.y;
  const int ns = cta / KS, ks = cta % KS;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmC);
    init_bars<kBwdStages>(bar, kBwdEpiWarps);
  }
  if (warp == 1) tmem_alloc<2 * kBwdN>(bar.tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *bar.tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      mbar_expect_tx(bar.w_bar, kBwdKb * kBwdWTile);
      for (int kb = 0; kb < kBwdKb; ++kb)
        tma_load_3d(wsm + kb * kBwdWTile, &tmW, ks * (kBwdKb * BK) + kb * BK, ns * kBwdN, 0, bar.w_bar);
      int s = 0;
      uint32_t ph = 0;
      for (int t = p.T - 1, i = 0; t > 0; --t, ++i) {
        spin_until(p.ctr_c + y, (unsigned)(i + 1) * (unsigned)G);  // d gates_t of this batch group is complete
        fence_proxy_async_all();
        rec_trace(p.trace, p.T, t, 3);
        for (int mt = y; mt < p.n_mtiles; mt += MG) {
          for (int kb = 0; kb < kBwdKb; ++kb) {
            mbar_wait_b(&bar.empty[s], ph ^ 1);
            mbar_expect_tx(&bar.full[s], kATile);
            tma_load_3d(stages + s * kATile, &tmA, ks * (kBwdKb * BK) + kb * BK, mt * 128, t, &bar.full[s]);
            if (++s == kBwdStages) { s = 0; ph ^= 1; }
          }
        }
        rec_trace(p.trace, p.T, t, 4);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====  D[128 x 128 units] += dgates_tile[128 x 64k] . W_slice[128 units x 64k]^T
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBwdN >> 3) << 17) |
                                 ((uint32_t)(128 >> 4) << 24);
      mbar_wait_b(bar.w_bar, 0);
      int s = 0, it = 0;
      uint32_t ph = 0;
      for (int t = p.T - 1; t > 0; --t) {
        for (int mt = y; mt < p.n_mtiles; mt += MG, ++it) {
          const int buf = it & 1;
          mbar_wait_b(&bar.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kBwdN);
          for (int kb = 0; kb < kBwdKb; ++kb) {
            mbar_wait_b(&bar.full[s], ph);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(stages + s * kATile);
            const uint32_t b_addr = smem_u32(wsm + kb * kBwdWTile);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_bf16(tmem_d, desc_kmajor(a_addr, k), desc_kmajor(b_addr, k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
            umma_commit(&bar.empty[s]);
            if (++s == kBwdStages) { s = 0; ph ^= 1; }
          }
          umma_commit(&bar.tmem_full[buf]);
        }
      }
    }
  } else {
    // ===== cell backward (phase C) and GEMM epilogue (phase G) =====
    const int quarter = warp & 3;          // TMEM lane quarter of this warp
    const int half = (warp - 2) >> 2;      // which half of the accumulator columns
    const int et = threadIdx.x - 64;       // 0..255
    uint8_t* my_stage = staging + (warp - 2) * 4096;
    uint8_t* st_row = my_stage + lane * 128;
    const uint32_t swz = (uint32_t)(lane & 7);
    int it = 0;
    for (int t = p.T - 1, i = 0; t >= 0; --t, ++i) {
      // ---- phase C: d gates_t, d c_{t-1} ----
      if (i > 0) {
        if (et == 0) {
          spin_until(p.ctr_g + y, (unsigned)i * (unsigned)G);  // every partial product of d h_t has been added
          __threadfence();
        }
        bwd_bar_sync();
      }
      if (et == 0) rec_trace(p.trace, p.T, t, 0);
      const bool last = (t == p.T - 1);
      const int items = 128 * (p.H / 8);
      CellItem ia;
      for (int mt = y; mt < p.n_mtiles; mt += MG) {
        for (int i0 = cta * kBwdEpiThreads + et; i0 < items; i0 += G * kBwdEpiThreads) {
          cell_load(p, t, mt, i0, last, ia);
          cell_compute_store_dg(p, last, ia);
          cell_store_state(p, t, last, ia);
        }
      }
      if (t > 0) {
        __threadfence();  // every thread: its d gates (and d h zeroing) stores are performed at GPU scope
        if (et == 0) rec_trace(p.trace, p.T, t, 1);
        bwd_bar_sync();
        if (et == 0) {
          __threadfence();
          atomicAdd(p.ctr_c + y, 1u);
          rec_trace(p.trace, p.T, t, 2);
        }
      }
      if (t == 0) break;
      // next step's cell operands on their way from HBM to L2 while the GEMM phase runs
      for (int mt = y; mt < p.n_mtiles; mt += MG)
        for (int i0 = cta * kBwdEpiThreads + et; i0 < items; i0 += G * kBwdEpiThreads) cell_prefetch(p, t - 1, mt, i0);
      // ---- phase G epilogue: partial d h_{t-1} -> fp32 reduce-add into dhrec[(t-1) & 1] ----
      for (int mt = y; mt < p.n_mtiles; mt += MG, ++it) {
        const int buf = it & 1;
        mbar_wait_b(&bar.tmem_full[buf], (it >> 1) & 1);
        if (et == 0) rec_trace(p.trace, p.T, t, 5);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * kBwdN) + ((uint32_t)(quarter * 32) << 16);
#pragma unroll 1
        for (int pp = 0; pp < 2; ++pp) {
          const int ps = half * 2 + pp;  // 32-column pass of the 128-column accumulator
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(tmem_d + (uint32_t)(ps * 32), r);
          if (lane == 0) bulk_wait_read_all();  // the previous reduce has finished reading this warp's staging tile
          tmem_ld_wait();
          __syncwarp();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<float4*>(st_row + (((uint32_t)g ^ swz) << 4)) =
                make_float4(__uint_as_float(r[4 * g]), __uint_as_float(r[4 * g + 1]), __uint_as_float(r[4 * g + 2]),
                            __uint_as_float(r[4 * g + 3]));
          if (pp == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar.tmem_empty[buf]);
          }
          fence_proxy_async_all();
          __syncwarp();
          if (lane == 0) {
            tma_reduce_add_3d(&tmC, my_stage, ns * kBwdN + ps * 32, mt * 128 + quarter * 32, (t - 1) & 1);
            bulk_commit();
          }
        }
      }
      if (et == 0) rec_trace(p.trace, p.T, t, 6);
      if (lane == 0) {
        bulk_wait_all();  // this warp's reduce-adds have been performed
        fence_proxy_async_all();
      }
      bwd_bar_sync();
      if (et == 0) {
        __threadfence();
        atomicAdd(p.ctr_g + y, 1u);
        rec_trace(p.trace, p.T, t, 7);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<2 * kBwdN>(tmem_base);
  }
}

// gate-column permutation of the recurrent weight for the forward kernel:
//   dst[k, 64*j + 4*u + gate] = src[k, gate*H + 16*j + u]      (src fp32 [H, 4H] with row stride lds)
__global__ void permute_gate_cols_kernel(const float* __restrict__ src, long lds, bf16* __restrict__ dst, int H) {
  const long n = (long)H * 4 * H;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long k = i / (4 * H);
    const int c = (int)(i % (4 * H));
    const int j = c >> 6, u = (c >> 2) & 15, gate = c & 3;
    dst[i] = __float2bfloat16_rn(src[k * lds + (long)gate * H + j * 16 + u]);
  }
}

unsigned long long* g_rec_trace = nullptr;

}  // namespace

void set_rec_trace(unsigned long long* buf) { g_rec_trace = buf; }

void permute_gate_cols(LaunchCtx c, const float* src, long lds, bf16* dst, int H) {
  const long n = (long)H * 4 * H;
  long blocks = (n + 255) / 256;
  if (blocks > 8L * kNumSMs) blocks = 8L * kNumSMs;
  permute_gate_cols_kernel<<<(int)blocks, 256, 0, c.stream>>>(src, lds, dst, H);
  if (c.launches) ++*c.launches;
}

// One condition for both directions: the persistent kernels keep the gate activations unit-major (col 4*unit + gate),
// the per-step kernels gate-major, so a step must run entirely on one schedule.
//   forward : W_hh slice [H x 64] bf16 resident  -> H <= 1024;  H / 16 CTAs per batch group
//   backward: (H / 128)^2 CTAs per batch group   -> H % 128 == 0
bool lstm_seq_supported(int B, int H) {
  if (B <= 0 || H % 128 || H > 1024) return false;
  return H / kFwdUnits <= kNumSMs && (H / 128) * (H / 128) <= kNumSMs;
}

// ctr: >= 2 * 64 zero-initialised-by-us unsigned counters owned by the caller
// xe: [T*B, ld_xe] bf16, columns [0, 4H) = the complete x-part of the gate pre-activations (word + global-image terms
// + biases), unit-major (col 4*unit + gate).  gact [T*B, 4H] comes out unit-major too.
int lstm_seq_fwd(LaunchCtx c, int B, int H, int T, const bf16* w_perm, const bf16* xe, long ld_xe, bf16* hs,
                 float* cs_buf, bf16* gact, unsigned* ctr) {
  const int slices = H / kFwdUnits;
  const int n_mtiles = (B + 127) / 128;
  const int mg = std::max(1, std::min(n_mtiles, kNumSMs / slices));
  const int nkb = H / BK;
  const size_t smem = (size_t)nkb * kFwdWTile + kFwdStages * kATile + kFwdXStage + kFwdCStage + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(lstm_seq_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("lstm_seq_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  CUtensorMap tmH, tmW, tmX, tmCs, tmGa;
  if (tma_map_3d(hs, H, B, H, BK, 128, &tmH, 2, T + 1, (uint64_t)B * H)) return -1;
  if (tma_map_3d(w_perm, 4 * H, H, 4 * H, kFwdN, BK, &tmW, 2, 1, 0)) return -1;
  if (tma_map_3d(xe, 4 * H, B, ld_xe, kFwdN, 128, &tmX, 2, T, (uint64_t)B * ld_xe)) return -1;
  if (tma_map_3d(cs_buf, H, B, H, kFwdUnits, 128, &tmCs, 4, T + 1, (uint64_t)B * H, false)) return -1;
  if (tma_map_3d(gact, 4 * H, B, 4 * H, kFwdN, 128, &tmGa, 2, T, (uint64_t)B * 4 * H)) return -1;
  SeqFwdParams p;
  p.B = B; p.H = H; p.T = T; p.n_mtiles = n_mtiles;
  p.hs = hs; p.cs = cs_buf; p.ctr = ctr;
  p.trace = g_rec_trace;
  cudaError_t e = cudaMemsetAsync(ctr, 0, 64 * sizeof(unsigned), c.stream);
  if (e != cudaSuccess) { set_error("lstm_seq_fwd: memset: %s", cudaGetErrorString(e)); return -1; }
  void* args[] = {&tmH, &tmW, &tmX, &tmCs, &tmGa, &p};
  e = cudaLaunchCooperativeKernel((const void*)lstm_seq_fwd_kernel, dim3(slices, mg), dim3(kFwdThreads), args, smem,
                                  c.stream);
  if (e != cudaSuccess) { set_error("lstm_seq_fwd: cooperative launch: %s", cudaGetErrorString(e)); return -1; }
  if (c.launches) ++*c.launches;
  return 0;
}

int lstm_seq_bwd(LaunchCtx c, int B, int H, int T, const bf16* w_hh, const bf16* dHhead, const bf16* dHprev,
                 const float* dChead, const bf16* gact, const float* cs, bf16* dG, long ld_dg, float* dhrec2,
                 float* dcrec, unsigned* ctr) {
  const int g = (H / 128) * (H / 128);
  const int n_mtiles = (B + 127) / 128;
  const int mg = std::max(1, std::min(n_mtiles, kNumSMs / g));
  const size_t smem = (size_t)kBwdKb * kBwdWTile + kBwdStages * kATile + kBwdEpiWarps * 4096 + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(lstm_seq_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) { set_error("lstm_seq_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  CUtensorMap tmA, tmW, tmC;
  if (tma_map_3d(dG, 4 * H, B, ld_dg, BK, 128, &tmA, 2, T, (uint64_t)B * ld_dg)) return -1;
  if (tma_map_3d(w_hh, 4 * H, H, 4 * H, BK, kBwdN, &tmW, 2, 1, 0)) return -1;
  if (tma_map_3d(dhrec2, H, B, H, 32, 32, &tmC, 4, 2, (uint64_t)B * H)) return -1;
  SeqBwdParams p;
  p.B = B; p.H = H; p.T = T; p.n_mtiles = n_mtiles;
  p.dHhead = dHhead; p.dHprev = dHprev; p.dChead = dChead; p.gact = gact; p.cs = cs; p.dG = dG; p.ld_dg = ld_dg;
  p.dhrec = dhrec2; p.dcrec = dcrec; p.ctr_c = ctr; p.ctr_g = ctr + 32;
  p.trace = g_rec_trace;
  cudaError_t e = cudaMemsetAsync(ctr, 0, 64 * sizeof(unsigned), c.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dhrec2, 0, 2 * sizeof(float) * (size_t)B * H, c.stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(dcrec, 0, sizeof(float) * (size_t)B * H, c.stream);
  if (e != cudaSuccess) { set_error("lstm_seq_bwd: memset: %s", cudaGetErrorString(e)); return -1; }
  void* args[] = {&tmA, &tmW, &tmC, &p};
  e = cudaLaunchCooperativeKernel((const void*)lstm_seq_bwd_kernel, dim3(g, mg), dim3(kBwdThreads), args, smem, c.stream);
  if (e != cudaSuccess) { set_error("lstm_seq_bwd: cooperative launch: %s", cudaGetErrorString(e)); return -1; }
  if (c.launches) ++*c.launches;
  return 0;
}

}  // namespace capdec
