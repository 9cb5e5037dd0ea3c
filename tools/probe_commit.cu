// probe_commit.cu -- issue cost / latency of the tcgen05 control instructions on sm_100a, measured with clock64
// by the single issuing thread (development probe; results recorded in profiles/).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_commit tools/probe_commit.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void commit_cluster(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void commit_plain(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.b64 [%0];" ::"l"((uint64_t)bar) : "memory");
}
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

// mode 0: the issuing thread is selected with `threadIdx.x == 0` (ptxas wraps each uniform-datapath instruction in
// an ELECT/BRA.U.ANY loop); mode 1: selected with elect.sync (bare instructions)
template <int MODE>
__global__ void __launch_bounds__(64, 1) probe(int iters) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  __shared__ uint64_t bars[4];
  __shared__ uint32_t tmem_slot;
  const uint32_t sA = base, sB = base + 16384;
  // zero the operand tiles
  for (uint32_t i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + 4 * i), "r"(0u) : "memory");
  if (threadIdx.x == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x < 32 && (MODE == 0 ? threadIdx.x == 0 : elect_one())) {
    printf("---- mode %d (%s) ----\n", MODE, MODE ? "elect.sync" : "threadIdx.x == 0");
    const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
    long long t0, t1;
    // T1: commit (.shared::cluster form), nothing outstanding
    t0 = clock64();
    for (int i = 0; i < iters; ++i) commit_cluster(b0);
    t1 = clock64();
    printf("T1 commit(.shared::cluster) issue: %.1f cycles each\n", (double)(t1 - t0) / iters);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) commit_plain(b1);
    t1 = clock64();
    printf("T2 commit(plain) issue: %.1f cycles each\n", (double)(t1 - t0) / iters);
    // let the arrivals land
    for (int i = 0; i < 100000; ++i) asm volatile("nanosleep.u32 20;");
    t0 = clock64();
    for (int i = 0; i < iters; ++i) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    t1 = clock64();
    printf("T3 tcgen05.fence::after issue: %.1f cycles each\n", (double)(t1 - t0) / iters);
    t0 = clock64();
    for (int i = 0; i < iters; ++i) asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    t1 = clock64();
    printf("T3b tcgen05.fence::before issue: %.1f cycles each\n", (double)(t1 - t0) / iters);
    // T4: commit -> arrival latency with a fresh barrier
    {
      const uint32_t b2 = smem_u32(&bars[2]);
      long long lat = 0, iss = 0;
      uint32_t ph = 0;
      for (int i = 0; i < 200; ++i) {
        const long long a = clock64();
        commit_cluster(b2);
        const long long b = clock64();
        while (!mbar_test_wait(b2, ph)) {}
        const long long c = clock64();
        ph ^= 1u;
        iss += b - a; lat += c - a;
      }
      printf("T4 commit alone: issue %.1f, issue->phase visible %.1f cycles\n", iss / 200.0, lat / 200.0);
      ph = 0;
      const uint32_t b3 = smem_u32(&bars[3]);
      long long tw = 0;
      for (int i = 0; i < 200; ++i) {
        commit_cluster(b3);
        const long long a = clock64();
        while (!mbar_try_wait(b3, ph)) {}
        tw += clock64() - a;
        ph ^= 1u;
      }
      printf("T4b commit then try_wait loop: %.1f cycles\n", tw / 200.0);
      // successful try_wait on an already completed phase
      t0 = clock64();
      uint32_t okc = 0;
      for (int i = 0; i < iters; ++i) okc += mbar_try_wait(b3, ph ^ 1u) ? 1 : 0;
      t1 = clock64();
      printf("T4c try_wait(already complete): %.1f cycles each (%u ok)\n", (double)(t1 - t0) / iters, okc);
    }
    // T5..: MMA streams
    const uint64_t ad = make_sw128_desc(sA), bd = make_sw128_desc(sB);
    const int NS[3] = {64, 128, 256};
    for (int c = 0; c < 3; ++c) {
      const int N = NS[c];
      const uint32_t idesc = make_idesc(N);
      for (int per = 4; per <= 16; per *= 4) {
        for (int ncommit = 0; ncommit <= 2; ++ncommit) {
          // reinit barrier phases irrelevant: commits just flip phases
          t0 = clock64();
          for (int i = 0; i < 400; ++i) {
            for (int k = 0; k < per; ++k) umma(tmem + (uint32_t)((i & 1) * 256), ad + 2 * (k & 3), bd + 2 * (k & 3), idesc, 1u);
            if (ncommit >= 1) commit_cluster(b0);
            if (ncommit >= 2) commit_cluster(b1);
          }
          // drain
          commit_cluster(smem_u32(&bars[2]));
          t1 = clock64();
          // wait until everything retired (phase unknown: spin a fixed time)
          for (int i = 0; i < 20000; ++i) asm volatile("nanosleep.u32 20;");
          printf("T5 N=%3d: %2d MMAs + %d commits per iteration: %.1f cycles/iter issue-side (MMA floor %d)\n", N, per, ncommit,
                 (double)(t1 - t0) / 400, per * N / 2);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x >= 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main() {
  cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  probe<0><<<1, 64, 60 * 1024>>>(2000);
  cudaDeviceSynchronize();
  probe<1><<<1, 64, 60 * 1024>>>(2000);
  cudaError_t e = cudaDeviceSynchronize();
  printf("probe: %s\n", cudaGetErrorString(e));
  return e != cudaSuccess;
}
