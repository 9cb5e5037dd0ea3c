// probe_umma_shift.cu -- development probe: can a K-major SWIZZLE_128B UMMA operand start at a
// 128-byte row that is NOT 1024-byte aligned (i.e. a row-shifted view of a larger smem tile), and
// which descriptor "base offset" makes it read the right rows?  Needed for tap-shifted views of a
// halo tile kept resident in shared memory (3x3 convolutions without re-loading A per tap).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/probe_umma_shift tools/probe_umma_shift.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int ROWS_A = 256, N = 64, K = 64;

__device__ __forceinline__ uint32_t su32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int shift_rows, int mode,
      float* out) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_slot;
  const uint32_t base = (su32(raw) + 1023u) & ~1023u;
  const uint32_t sA = base, sB = base + ROWS_A * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar_load)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(su32(&bar_mma)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(su32(&tmem_slot)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_slot;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&bar_load)), "r"((ROWS_A + N) * 128));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sA), "l"(reinterpret_cast<uint64_t>(&tmA)), "r"(su32(&bar_load)), "r"(0), "r"(0) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(sB), "l"(reinterpret_cast<uint64_t>(&tmB)), "r"(su32(&bar_load)), "r"(0), "r"(0) : "memory");
    uint32_t ok = 0;
    while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(su32(&bar_load)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t a_addr = sA + shift_rows * 128;
    uint64_t base_off = 0;
    if (mode == 1) base_off = (a_addr >> 7) & 7;
    auto desc = [&](uint32_t addr, uint64_t bo) {
      uint64_t d = 0;
      d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
      d |= (uint64_t)1 << 16;
      d |= (uint64_t)(1024 >> 4) << 32;
      d |= (uint64_t)1 << 46;
      d |= bo << 49;
      d |= (uint64_t)2 << 61;
      return d;
    };
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = desc(a_addr, base_off) + 2 * k, bd = desc(sB, 0) + 2 * k;
      const uint32_t acc = k != 0;
      asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                   ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(su32(&bar_mma)) : "memory");
  }
  {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }" : "=r"(ok) : "r"(su32(&bar_mma)) : "memory");
      if (!ok && clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("TIMEOUT\n"); return; }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u));
}

int main() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn enc = (EncodeTiledFn)fn;
  std::vector<__nv_bfloat16> A((size_t)ROWS_A * K), B((size_t)N * K);
  std::vector<float> Af(A.size()), Bf(B.size());
  srand(1);
  for (size_t i = 0; i < A.size(); ++i) { float v = (rand() % 17 - 8) / 8.f; A[i] = __float2bfloat16(v); Af[i] = __bfloat162float(A[i]); }
  for (size_t i = 0; i < B.size(); ++i) { float v = (rand() % 13 - 6) / 8.f; B[i] = __float2bfloat16(v); Bf[i] = __bfloat162float(B[i]); }
  __nv_bfloat16 *dA, *dB; float* dO;
  CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2)); CK(cudaMalloc(&dO, 128 * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
  CUtensorMap tmA, tmB;
  {
    cuuint64_t dims[2] = {K, ROWS_A}; cuuint64_t st[1] = {K * 2}; cuuint32_t box[2] = {64, ROWS_A}; cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, st, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode A %d\n", (int)r);
    cuuint64_t dimsb[2] = {K, N}; cuuint32_t boxb[2] = {64, N};
    r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, st, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode B %d\n", (int)r);
  }
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  std::vector<float> O(128 * N);
  const int shifts[] = {0, 1, 2, 3, 5, 8, 9, 30, 31, 62, 100};
  for (int mode = 0; mode < 2; ++mode)
    for (int sh : shifts) {
      CK(cudaMemset(dO, 0, 128 * N * 4));
      probe<<<1, 128, 1024 + (ROWS_A + N) * 128>>>(tmA, tmB, sh, mode, dO);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d shift %d: CUDA error %s\n", mode, sh, cudaGetErrorString(e)); return 1; }
      CK(cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost));
      double maxerr = 0; int bad_rows = 0;
      for (int r = 0; r < 128; ++r) {
        double rowerr = 0;
        for (int n = 0; n < N; ++n) {
          double ref = 0;
          for (int k = 0; k < K; ++k) ref += (double)Af[(size_t)(r + sh) * K + k] * Bf[(size_t)n * K + k];
          rowerr = fmax(rowerr, fabs(ref - O[r * N + n]));
        }
        if (rowerr > 1e-3) ++bad_rows;
        maxerr = fmax(maxerr, rowerr);
      }
      printf("base_offset mode %d (%s) shift %3d rows: max err %.4g, bad rows %d/128\n", mode,
             mode ? "(addr>>7)&7" : "0", sh, maxerr, bad_rows);
    }
  printf("probe done\n");
  return 0;
}
