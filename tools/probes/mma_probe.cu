// Micro-benchmark: cycles per tcgen05.mma (kind::f16, M = 128, cta_group::1, K = 16 per instruction) issued back to
// back by one thread, for several N, operand sources and commit cadences.  Operands are whatever shared memory /
// tensor memory holds (timing only).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)layout << 61;
  return d;
}
__device__ __forceinline__ uint32_t idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t i, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(i), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t i, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(i), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
               : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFF));
  return pred;
}

// mode 0: SS no swizzle (A chunk pitch 2048, B chunk pitch n * 16); 1: SS, swizzle-128B descriptors (timing only);
// 2: TS (A from tensor memory)
template <int kPattern>  // 0: `if (tid == 0)` region; 1: warp 0 enters, `if (elect_one())` region around the whole loop
__global__ void __launch_bounds__(128, 1) probe(int n, int mode, int mmas, int commit_mask, long long* out) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t tbase;
  const int tid = threadIdx.x;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[1])));
  }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tm = tbase;
  bool issuer = false;
  if (kPattern == 0) issuer = tid == 0;
  else if (tid < 32) issuer = elect_one() != 0;
  if (issuer) {
    const uint32_t a_sh = smem_u32(sm), b_sh = smem_u32(sm + 80 * 1024);
    const uint32_t id = idesc(n);
    uint32_t parity = 0;
    const uint64_t a0 = mode == 1 ? desc(a_sh, 16, 1024, 2) : desc(a_sh, 2048, 128, 0);
    const uint64_t b0 = mode == 1 ? desc(b_sh, 16, 1024, 2) : desc(b_sh, n * 16, 128, 0);
    const uint64_t astep = mode == 1 ? 2 : (2 * 2048) >> 4, bstep = mode == 1 ? 2 : (uint64_t)((2 * n * 16) >> 4);
    for (int rep = 0; rep < 3; ++rep) {  // rep 0 warms up
      const long long t0 = clock64();
      uint64_t ad = a0, bd = b0;
      uint32_t acol = tm + 256, acc = 0;
      for (int m = 0; m < mmas; ++m) {
        if (mode == 2) mma_ts(tm, acol, bd, id, acc);
        else mma_ss(tm, ad, bd, id, acc);
        acc = 1;
        if ((m & 7) == 7) { ad = a0; bd = b0; acol = tm + 256; }
        else { ad += astep; bd += bstep; acol += 8; }
        if (commit_mask != 0 && (m & commit_mask) == commit_mask) commit(&bar[1]);
      }
      const long long t1 = clock64();
      commit(&bar[0]);
      wait(&bar[0], parity);
      parity ^= 1;
      const long long t2 = clock64();
      if (rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm) : "memory");
}

int main() {
  long long* out;
  cudaMallocManaged(&out, 16);
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"SS no-swizzle", "SS swizzle-128B", "TS (A in TMEM)"};
  for (int pat = 0; pat < 2; ++pat)
    for (int mode = 0; mode < 3; ++mode)
      for (int n : {32, 256})
        for (int cm : {0, 1, 3}) {
          const int mmas = 64;
          if (pat) probe<1><<<1, 128, 200 * 1024>>>(n, mode, mmas, cm, out);
          else probe<0><<<1, 128, 200 * 1024>>>(n, mode, mmas, cm, out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("%s n=%d: %s\n", names[mode], n, cudaGetErrorString(e)); return 1; }
          printf("%s %-16s N=%3d commit every %d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (floor %d)\n",
                 pat ? "elect region" : "tid==0 region", names[mode], n, cm ? cm + 1 : 0, (double)out[0] / mmas,
                 (double)out[1] / mmas, n / 2);
        }
  return 0;
}
