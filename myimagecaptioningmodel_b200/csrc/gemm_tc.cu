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

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
constexpr int kThreads = 192;
constexpr int kABytes = BM * BK * 2;  // 16 KB

template <int BN> struct TcCfg {
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : (BN == 128 ? 6 : 8);
  static constexpr int kTmemCols = 2 * BN;  // two accumulator buffers (power of two >= 32)
  static constexpr int kStagingBytes = 4 * 2 * 4096;  // per epilogue warp: two 32-row x 128-byte output tiles
  static constexpr int kSmemBytes = kStages * kStageBytes + kStagingBytes + 1024 /*align slack*/ + 256 /*barriers*/;
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
__device__ __forceinline__ void tma_store_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N> __device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct TcShape {
  int M, N, K;
  int tiles_m, tiles_n, splits, kb_per_split;
  long ldc;
  int vec_ok;
  int rotate_k;  // start every CTA's k-loop at a different k-block (spreads L2 hot lines among CTAs sharing a tile)
  int tma_out;   // epilogue writes through shared memory + TMA tensor stores (needs an aligned C)
};

// bias / addend / activation on 8 consecutive columns of one output row
__device__ __forceinline__ void epi_math8(float (&v)[8], const GemmEpi& epi, long row, int col, bool row_ok, int N) {
  if (col + 8 <= N) {
    if (epi.bias) {
      float b[8];
      load8(epi.bias + col, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += b[i];
    }
    if (epi.addend && row_ok) {
      float a[8];
      load8(reinterpret_cast<const bf16*>(epi.addend) + row * epi.ld_add + col, a);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] += a[i];
    }
    if (epi.act == ACT_TANH) {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = tanh_t<true>(v[i]);
    } else if (epi.act == ACT_DTANH && row_ok) {
      float y[8];
      load8(reinterpret_cast<const bf16*>(epi.dact_y) + row * epi.ld_y + col, y);
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] *= (1.f - y[i] * y[i]);
    }
  } else {
    for (int i = 0; i < 8; ++i)
      v[i] = (row_ok && col + i < N) ? epi_value<bf16>(v[i], (int)row, col + i, epi) : 0.f;
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, void* __restrict__ C, TcShape sh, GemmEpi epi) {
  using Cfg = TcCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* staging = smem + Cfg::kStages * Cfg::kStageBytes;  // 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(staging + Cfg::kStagingBytes);
  uint64_t* empty_bar = full_bar + Cfg::kStages;
  uint64_t* tmem_full_bar = empty_bar + Cfg::kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb_total = (sh.K + BK - 1) / BK;
  const int tiles = sh.tiles_m * sh.tiles_n;
  const int total_work = tiles * sh.splits;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (sh.tma_out) tma_prefetch_desc(&tmC);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tmem_full_bar[b], 1);
      mbar_init(&tmem_empty_bar[b], 4);  // one arrival per epilogue warp
    }
    mbar_fence_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      int s = 0;
      uint32_t ph = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int tile = w % tiles, z = w / tiles;
        const int n0 = (tile % sh.tiles_n) * BN, m0 = (tile / sh.tiles_n) * BM;
        const int kb0 = z * sh.kb_per_split, kb1 = min(nkb_total, kb0 + sh.kb_per_split);
        const int nk = kb1 - kb0;
        const int rot = sh.rotate_k ? (int)((blockIdx.x / sh.tiles_n) % nk) : 0;
        for (int i = 0; i < nk; ++i) {
          const int kb = kb0 + (i + rot < nk ? i + rot : i + rot - nk);
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sA = smem + s * Cfg::kStageBytes;
          uint8_t* sB = sA + kABytes;
          mbar_expect_tx(&full_bar[s], Cfg::kStageBytes);
          if (!A_MN) {
            tma_load_2d(sA, &tmA, kb * BK, m0, &full_bar[s]);
          } else {
#pragma unroll
            for (int i = 0; i < BM / 64; ++i) tma_load_2d(sA + i * 8192, &tmA, m0 + i * 64, kb * BK, &full_bar[s]);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, kb * BK, n0, &full_bar[s]);
          } else {
#pragma unroll
            for (int i = 0; i < BN / 64; ++i) tma_load_2d(sB + i * 8192, &tmB, n0 + i * 64, kb * BK, &full_bar[s]);
          }
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
      }
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
        const int z = w / tiles;
        const int kb0 = z * sh.kb_per_split, kb1 = min(nkb_total, kb0 + sh.kb_per_split);
        const int buf = it & 1;
        const uint32_t acc_ph = (it >> 1) & 1;
        mbar_wait(&tmem_empty_bar[buf], acc_ph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + s * Cfg::kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            umma_bf16(tmem_d, slice_desc<A_MN>(a_addr, k), slice_desc<B_MN>(b_addr, k), idesc,
                      (kb > kb0 || k > 0) ? 1u : 0u);
          umma_commit(&empty_bar[s]);  // frees the smem slot once these MMAs retire
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
        umma_commit(&tmem_full_bar[buf]);  // accumulator complete
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int it = 0;
    uint32_t epi_sb = 0;  // staging-buffer toggle
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++it) {
      const int tile = w % tiles;
      const int n0 = (tile % sh.tiles_n) * BN, m0 = (tile / sh.tiles_n) * BM;
      const int buf = it & 1;
      const uint32_t acc_ph = (it >> 1) & 1;
      const int row = m0 + quarter * 32 + lane;
      const bool row_ok = row < sh.M;
      mbar_wait(&tmem_full_bar[buf], acc_ph);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(quarter * 32) << 16);
      if (sh.tma_out) {
        // ---- coalesced path: registers -> swizzled smem tile (32 rows x 128 B) -> TMA tensor store / reduce-add ----
        uint8_t* my_stage = staging + quarter * 8192;
        const uint32_t swz = (uint32_t)(lane & 7);
        uint8_t* my_row = my_stage + lane * 128;
        const bool f32_out = epi.out_f32 || epi.atomic;
        if (!f32_out) {
#pragma unroll 1
          for (int c2 = 0; c2 < BN / 64; ++c2) {
            uint32_t r0[32], r1[32];
            __syncwarp();
            tmem_ld32(tmem_d + (uint32_t)(c2 * 64), r0);
            tmem_ld32(tmem_d + (uint32_t)(c2 * 64 + 32), r1);
            tmem_ld_wait();
            if (c2 == BN / 64 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
            }
            const int nb = n0 + c2 * 64;
            if (nb >= sh.N) continue;  // warp-uniform
            uint8_t* dst = my_row + (epi_sb & 1) * 4096;
            if (lane == 0) bulk_wait_read<1>();  // the store issued two tiles ago has finished reading this buffer
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(g < 4 ? r0[g * 8 + i] : r1[(g - 4) * 8 + i]);
              epi_math8(v, epi, row, nb + g * 8, row_ok, sh.N);
              uint4 u;
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
              for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
              *reinterpret_cast<uint4*>(dst + (((uint32_t)g ^ swz) << 4)) = u;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmC, my_stage + (epi_sb & 1) * 4096, nb, m0 + quarter * 32);
              bulk_commit();
            }
            ++epi_sb;
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < BN / 32; ++c) {
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(tmem_d + (uint32_t)(c * 32), r);
            tmem_ld_wait();
            if (c == BN / 32 - 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
            }
            const int nb = n0 + c * 32;
            if (nb >= sh.N) continue;  // warp-uniform
            uint8_t* dst = my_row + (epi_sb & 1) * 4096;
            if (lane == 0) bulk_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float v[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
              epi_math8(v, epi, row, nb + g * 8, row_ok, sh.N);
              *reinterpret_cast<float4*>(dst + (((uint32_t)(2 * g) ^ swz) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
              *reinterpret_cast<float4*>(dst + (((uint32_t)(2 * g + 1) ^ swz) << 4)) = make_float4(v[4], v[5], v[6], v[7]);
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (epi.atomic) tma_reduce_add_2d(&tmC, my_stage + (epi_sb & 1) * 4096, nb, m0 + quarter * 32);
              else tma_store_2d(&tmC, my_stage + (epi_sb & 1) * 4096, nb, m0 + quarter * 32);
              bulk_commit();
            }
            ++epi_sb;
          }
        }
        continue;
      }
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the divergent stores of the previous trip
        tmem_ld32(tmem_d + (uint32_t)(c * 32), r);
        tmem_ld_wait();
        if (c == BN / 32 - 1) {
          // all of this warp's TMEM reads are done: hand the accumulator back to the MMA issuer
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        const int nb = n0 + c * 32;
        if (!row_ok || nb >= sh.N) continue;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = nb + g * 8;
          if (col >= sh.N) break;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[g * 8 + i]);
          if (sh.vec_ok && col + 8 <= sh.N) {
            if (epi.bias) {
              float b[8];
              load8(epi.bias + col, b);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += b[i];
            }
            if (epi.addend) {
              float a[8];
              load8(reinterpret_cast<const bf16*>(epi.addend) + (long)row * epi.ld_add + col, a);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += a[i];
            }
            if (epi.act == ACT_TANH) {
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] = tanh_t<true>(v[i]);
            } else if (epi.act == ACT_DTANH) {
              float y[8];
              load8(reinterpret_cast<const bf16*>(epi.dact_y) + (long)row * epi.ld_y + col, y);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] *= (1.f - y[i] * y[i]);
            }
            if (epi.atomic) {
              float* p = reinterpret_cast<float*>(C) + (long)row * sh.ldc + col;
              red_add_v4(p, v[0], v[1], v[2], v[3]);
              red_add_v4(p + 4, v[4], v[5], v[6], v[7]);
            } else if (epi.out_f32) {
              store8(reinterpret_cast<float*>(C) + (long)row * sh.ldc + col, v);
            } else {
              store8(reinterpret_cast<bf16*>(C) + (long)row * sh.ldc + col, v);
            }
          } else {
            for (int i = 0; i < 8 && col + i < sh.N; ++i) {
              float x = epi_value<bf16>(v[i], row, col + i, epi);
              epi_store<bf16>(x, C, sh.ldc, row, col + i, epi);
            }
          }
        }
      }
    }
  }
  if (warp >= 2 && lane == 0 && sh.tma_out) bulk_wait<0>();  // all tensor stores of this thread have completed
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

void load_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

struct MapKey {
  const void* p; uint64_t inner, outer, ld; uint32_t bi, bo; int esize;
  bool operator==(const MapKey& o) const {
    return p == o.p && inner == o.inner && outer == o.outer && ld == o.ld && bi == o.bi && bo == o.bo && esize == o.esize;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = (size_t)k.p;
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.inner); mix(k.outer); mix(k.ld); mix(k.bi); mix(k.bo); mix((uint64_t)k.esize);
    return h;
  }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

// 2-D tensor [outer, inner] of bf16 (esize 2) or fp32 (esize 4) with row stride ld (elements); box = [bo, bi];
// 128-byte swizzle.
int get_map(const void* p, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t bi, uint32_t bo, CUtensorMap* out,
            int esize = 2) {
  std::call_once(g_encode_once, load_encode);
  if (!g_encode) { set_error("cuTensorMapEncodeTiled entry point not available"); return -1; }
  if (((uintptr_t)p & 15) || ((ld * esize) & 15)) {
    set_error("gemm_tc: operand %p (ld=%llu) is not 16-byte aligned for TMA", p, (unsigned long long)ld);
    return -1;
  }
  MapKey key{p, inner, outer, ld, bi, bo, esize};
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return 0; }
  }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * (uint64_t)esize};
  cuuint32_t box[2] = {bi, bo};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(p), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
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

template <int BN, bool A_MN, bool B_MN>
int launch(GemmContext& ctx, const GemmCall& c, int splits, int kbps) {
  using Cfg = TcCfg<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::kSmemBytes);
    if (e != cudaSuccess) { set_error("gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return -1; }
    attr_set = true;
  }
  CUtensorMap tmA, tmB, tmC;
  if (!A_MN) { if (get_map(c.A, c.K, c.M, c.lda, BK, BM, &tmA)) return -1; }
  else { if (get_map(c.A, c.M, c.K, c.lda, 64, BK, &tmA)) return -1; }
  if (!B_MN) { if (get_map(c.B, c.K, c.N, c.ldb, BK, BN, &tmB)) return -1; }
  else { if (get_map(c.B, c.N, c.K, c.ldb, 64, BK, &tmB)) return -1; }
  const GemmEpi& e = c.epi;
  const int out_elem = (e.out_f32 || e.atomic) ? 4 : 2;
  int vec_ok = ((c.ldc % 8) == 0) && (((uintptr_t)c.C) % (out_elem * 8) == 0);
  if (e.bias && ((uintptr_t)e.bias & 31)) vec_ok = 0;
  if (e.addend && (((uintptr_t)e.addend & 15) || (e.ld_add % 8))) vec_ok = 0;
  if (e.dact_y && (((uintptr_t)e.dact_y & 15) || (e.ld_y % 8))) vec_ok = 0;
  // output through TMA tensor stores when C is 16-byte aligned with a 16-byte-multiple pitch (always true for the
  // decoder's own buffers); otherwise the direct per-thread store path is used
  const int tma_out = (((uintptr_t)c.C & 15) == 0) && (((c.ldc * out_elem) & 15) == 0) &&
                      (!e.bias || ((uintptr_t)e.bias & 31) == 0) &&
                      (!e.addend || ((((uintptr_t)e.addend & 15) == 0) && (e.ld_add % 8) == 0)) &&
                      (!e.dact_y || ((((uintptr_t)e.dact_y & 15) == 0) && (e.ld_y % 8) == 0));
  if (tma_out) {
    if (get_map(c.C, c.N, c.M, c.ldc, out_elem == 2 ? 64 : 32, 32, &tmC, out_elem)) return -1;
  } else {
    tmC = tmA;
  }
  TcShape sh;
  sh.tma_out = tma_out;
  sh.M = c.M; sh.N = c.N; sh.K = c.K;
  sh.tiles_m = (c.M + BM - 1) / BM; sh.tiles_n = (c.N + BN - 1) / BN;
  sh.splits = splits; sh.kb_per_split = kbps; sh.ldc = c.ldc; sh.vec_ok = vec_ok;
  {
    static int rot_env = -1;
    if (rot_env < 0) { const char* ev = getenv("CAPDEC_GEMM_ROT"); rot_env = ev ? atoi(ev) : 1; }
    sh.rotate_k = rot_env;
  }
  const long total = (long)sh.tiles_m * sh.tiles_n * splits;
  const int grid = (int)std::min<long>(total, kNumSMs);
  gemm_tc_kernel<BN, A_MN, B_MN><<<grid, kThreads, Cfg::kSmemBytes, ctx.stream>>>(tmA, tmB, tmC, c.C, sh, c.epi);
  ctx.launches++;
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) { set_error("gemm_tc launch: %s", cudaGetErrorString(err)); return -1; }
  return 0;
}

template <int BN> int launch_majors(GemmContext& ctx, const GemmCall& c, int splits, int kbps) {
  if (c.a_mn) return c.b_mn ? launch<BN, true, true>(ctx, c, splits, kbps) : launch<BN, true, false>(ctx, c, splits, kbps);
  return c.b_mn ? launch<BN, false, true>(ctx, c, splits, kbps) : launch<BN, false, false>(ctx, c, splits, kbps);
}

}  // namespace

int gemm_bf16_tc(GemmContext& ctx, const GemmCall& c) {
  if (c.M <= 0 || c.N <= 0 || c.K <= 0) return 0;
  if (c.splits > 1 && !c.epi.atomic) { set_error("gemm_tc: split-K needs an atomic epilogue"); return -1; }
  const int nkb = (c.K + BK - 1) / BK;
  const int mt = (c.M + BM - 1) / BM;
  // Pick the tile width and the k-split that minimise a simple cost model of the persistent kernel:
  //   per CTA: ceil(work / 148) items, each  kb_per_split * t_kblock(BN) + t_epilogue(BN);
  //   t_kblock ~ bytes moved per k-block (operand streaming bound), epilogue ~ BN (x3 for reductions).
  int best_bn = 64, best_sp = 1;
  double best_cost = 1e300;
  const int cands[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    if (bn > 64 && c.N <= bn / 2) continue;
    const long tiles = (long)mt * ((c.N + bn - 1) / bn);
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
  if (const char* ev = getenv("CAPDEC_GEMM_BN")) {  // tuning / profiling override
    const int v = atoi(ev);
    if (v == 64 || v == 128 || v == 256) best_bn = v;
  }
  const int kbps = (nkb + best_sp - 1) / best_sp;
  const int splits = (nkb + kbps - 1) / kbps;
  switch (best_bn) {
    case 256: return launch_majors<256>(ctx, c, splits, kbps);
    case 128: return launch_majors<128>(ctx, c, splits, kbps);
    default: return launch_majors<64>(ctx, c, splits, kbps);
  }
}

}  // namespace capdec
