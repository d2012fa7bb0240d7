// Micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16 in, fp32 accumulate) as a function of the instruction shape and
// of where the operands sit -- the cost model behind the recurrence kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_bench scripts/umma_bench.cu && /tmp/umma_bench
// One CTA per SM on `ctas` SMs; operands are resident in shared memory (garbage data), one thread issues `iters` groups
// of 4 K=16 slices of one 64-wide k-block, commits, waits.  No global-memory traffic inside the timed region.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_k(uint32_t tile, int k) { return make_desc(tile + k * 32, 16, 1024); }
__device__ __forceinline__ uint64_t desc_mn(uint32_t tile, int k) { return make_desc(tile + k * 2048, 8192, 1024); }

struct Args { int M, N, a_mn, b_mn, a_tmem, iters; };

__global__ void __launch_bounds__(128, 1) bench_kernel(Args a, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;             // 128 x 64 bf16 = 16 KB
  uint8_t* sB = smem + 16384;     // 256 x 64 bf16 = 32 KB
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_ptr)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((a.a_mn ? 1u : 0u) << 15) | ((a.b_mn ? 1u : 0u) << 16) |
                           ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(a.M >> 4) << 24);
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
    const uint32_t a_t = tmem + 256;   // A-in-TMEM variant: columns 256.. (8 columns per K=16 slice)
    long long t0 = clock64();
    for (int it = 0; it < a.iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t db = a.b_mn ? desc_mn(b_addr, k) : desc_k(b_addr, k);
        if (a.a_tmem) {
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem),
                       "r"(a_t + 8 * k), "l"(db), "r"(idesc), "r"(1u)
                       : "memory");
        } else {
          const uint64_t da = a.a_mn ? desc_mn(a_addr, k) : desc_k(a_addr, k);
          asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem),
                       "l"(da), "l"(db), "r"(idesc), "r"(1u)
                       : "memory");
        }
      }
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                   : "=r"(ok)
                   : "r"(smem_u32(&bar)), "r"(0u)
                   : "memory");
    }
    long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

int main(int argc, char** argv) {
  const int ctas = argc > 1 ? atoi(argv[1]) : 1;
  long long* out;
  cudaMalloc(&out, 8);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int iters = 2000;
  printf("%-5s %-5s %-6s %-6s %-7s  cycles/MMA (K=16)   MACs/cycle\n", "M", "N", "A", "B", "A-src");
  const int Ms[] = {128, 64};
  const int Ns[] = {16, 32, 64, 128, 256};
  for (int mi = 0; mi < 2; ++mi)
    for (int at = 0; at < 2; ++at)
      for (int bm = 0; bm < 2; ++bm)
        for (int am = 0; am < (at ? 1 : 2); ++am)
          for (int ni = 0; ni < 5; ++ni) {
            Args a{Ms[mi], Ns[ni], am, bm, at, iters};
            if (a.M == 128 && a.N % 16) continue;
            bench_kernel<<<ctas, 128, 64 * 1024>>>(a, out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("M=%d N=%d a_mn=%d b_mn=%d a_tmem=%d: %s\n", a.M, a.N, am, bm, at, cudaGetErrorString(e)); return 1; }
            long long cyc;
            cudaMemcpy(&cyc, out, 8, cudaMemcpyDeviceToHost);
            const double per = (double)cyc / (4.0 * iters);
            printf("%-5d %-5d %-6s %-6s %-7s  %8.1f            %8.1f\n", a.M, a.N, am ? "MN" : "K", bm ? "MN" : "K",
                   at ? "tmem" : "smem", per, (double)a.M * a.N * 16 / per);
          }
  return 0;
}
