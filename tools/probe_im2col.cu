// probe_im2col.cu -- development probe (not product, not test): issues single TMA im2col loads on a
// tiny channels-last tensor whose channels 0..3 encode (n, h|d*16+h, w, 1) and prints which input
// pixel landed in every shared-memory row.  Used once to confirm the descriptor conventions
// (corner boxes, start coordinates, tap offsets, traversal strides) that csrc/net.cpp relies on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/probe_im2col tools/probe_im2col.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

constexpr int ROWS = 32;

__global__ void probe4d(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int n, int ow, int oh,
                        __nv_bfloat16* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  base = (base + 1023u) & ~1023u;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(ROWS * 128));
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};" ::"r"(base),
        "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c), "r"(w), "r"(h), "r"(n), "h"((uint16_t)ow),
        "h"((uint16_t)oh)
        : "memory");
  }
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }"
                 : "=r"(ok) : "r"(b) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("TIMEOUT waiting for TMA\n"); return; }
  }
  __syncthreads();
  const uint8_t* s = smem + (base - (uint32_t)__cvta_generic_to_shared(smem));
  for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
    const int phys_chunk = 0 ^ (r & 7);
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(s + r * 128 + phys_chunk * 16);
    for (int j = 0; j < 4; ++j) out[r * 4 + j] = p[j];
  }
}

__global__ void probe5d(const __grid_constant__ CUtensorMap tm, int c, int w, int h, int d, int n, int ow, int oh,
                        int od, __nv_bfloat16* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t base = (uint32_t)__cvta_generic_to_shared(smem);
  base = (base + 1023u) & ~1023u;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(ROWS * 128));
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};" ::"r"(base),
        "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"((uint16_t)ow),
        "h"((uint16_t)oh), "h"((uint16_t)od)
        : "memory");
  }
  uint32_t ok = 0;
  long long t0 = clock64();
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p; }"
                 : "=r"(ok) : "r"(b) : "memory");
    if (clock64() - t0 > 2000000000LL) { if (threadIdx.x == 0) printf("TIMEOUT waiting for TMA\n"); return; }
  }
  __syncthreads();
  const uint8_t* s = smem + (base - (uint32_t)__cvta_generic_to_shared(smem));
  for (int r = threadIdx.x; r < ROWS; r += blockDim.x) {
    const int phys_chunk = 0 ^ (r & 7);
    const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(s + r * 128 + phys_chunk * 16);
    for (int j = 0; j < 4; ++j) out[r * 4 + j] = p[j];
  }
}

static EncodeIm2colFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &q));
  return (EncodeIm2colFn)fn;
}

static void dump(const char* title, const std::vector<__nv_bfloat16>& h) {
  printf("%s\n  row: (n,h,w,valid)\n", title);
  for (int r = 0; r < ROWS; ++r) {
    printf("  %2d:(%g,%g,%g,%g)", r, __bfloat162float(h[r * 4]), __bfloat162float(h[r * 4 + 1]),
           __bfloat162float(h[r * 4 + 2]), __bfloat162float(h[r * 4 + 3]));
    if (r % 4 == 3) printf("\n");
  }
}

int main() {
  EncodeIm2colFn enc = get_encode();
  const int C = 64;
  __nv_bfloat16* dout;
  CK(cudaMalloc(&dout, ROWS * 4 * 2));
  std::vector<__nv_bfloat16> hout(ROWS * 4);
  {  // ---------------- 4-D: N=2,H=5,W=6 ----------------
    const int N = 2, H = 5, W = 6;
    std::vector<__nv_bfloat16> x((size_t)N * H * W * C, __float2bfloat16(0.f));
    for (int n = 0; n < N; ++n)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          __nv_bfloat16* p = &x[(((size_t)n * H + h) * W + w) * C];
          p[0] = __float2bfloat16((float)n); p[1] = __float2bfloat16((float)h);
          p[2] = __float2bfloat16((float)w); p[3] = __float2bfloat16(1.f);
        }
    __nv_bfloat16* dx;
    CK(cudaMalloc(&dx, x.size() * 2));
    CK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
    for (int stride = 1; stride <= 2; ++stride) {
      CUtensorMap tm;
      cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
      cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
      int lower[2] = {-1, -1}, upper[2] = {-1, -1};  // 3x3, pad 1
      cuuint32_t es[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dx, dims, strides, lower, upper, 64, ROWS, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("encode 4d stride %d -> %d\n", stride, (int)r);
      if (r != CUDA_SUCCESS) continue;
      struct { int w, h, n, ow, oh; const char* t; } cases[] = {
          {-1, -1, 0, 0, 0, "start (w=-1,h=-1,n=0) tap (0,0)"},
          {-1, -1, 0, 1, 1, "start (w=-1,h=-1,n=0) tap (1,1)"},
          {-1, -1, 0, 2, 2, "start (w=-1,h=-1,n=0) tap (2,2)"},
          {-1 + stride, -1 + stride, 0, 1, 1, "start at output pixel (q=1,p=1) tap (1,1)"},
          {-1, -1, 1, 1, 1, "start image n=1 tap (1,1)"},
      };
      for (auto& cs : cases) {
        CK(cudaMemset(dout, 0xFF, ROWS * 4 * 2));
        probe4d<<<1, 32, 1024 + ROWS * 128>>>(tm, 0, cs.w, cs.h, cs.n, cs.ow, cs.oh, dout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hout.data(), dout, ROWS * 4 * 2, cudaMemcpyDeviceToHost));
        char title[256];
        snprintf(title, sizeof title, "[4d H=5 W=6 stride %d 3x3 pad1] %s", stride, cs.t);
        dump(title, hout);
      }
    }
    // overlapping-window view used by the stem: C=64 window of 4 cells x 16 ch, pixel stride 16 elements
    {
      const int CW = 9, CH = 6, F = 1, OW = CW - 3;
      std::vector<__nv_bfloat16> cells((size_t)F * CH * CW * 16 + 64, __float2bfloat16(0.f));
      for (int y = 0; y < CH; ++y)
        for (int xx = 0; xx < CW; ++xx) {
          __nv_bfloat16* p = &cells[((size_t)y * CW + xx) * 16];
          p[0] = __float2bfloat16(0.f); p[1] = __float2bfloat16((float)y); p[2] = __float2bfloat16((float)xx);
          p[3] = __float2bfloat16(1.f);
        }
      __nv_bfloat16* dc;
      CK(cudaMalloc(&dc, cells.size() * 2));
      CK(cudaMemcpy(dc, cells.data(), cells.size() * 2, cudaMemcpyHostToDevice));
      CUtensorMap tm;
      cuuint64_t dims[4] = {64, (cuuint64_t)OW, (cuuint64_t)CH, (cuuint64_t)F};
      cuuint64_t strides[3] = {16 * 2, (cuuint64_t)CW * 16 * 2, (cuuint64_t)CH * CW * 16 * 2};
      int lower[2] = {0, 0}, upper[2] = {0, -3};  // kernel 1(w) x 4(h), no pad
      cuuint32_t es[4] = {1, 1, 1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dc, dims, strides, lower, upper, 64, ROWS, es,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("encode overlapping-window stem view -> %d\n", (int)r);
      if (r == CUDA_SUCCESS) {
        for (int oh = 0; oh < 4; oh += 3) {
          CK(cudaMemset(dout, 0xFF, ROWS * 4 * 2));
          probe4d<<<1, 32, 1024 + ROWS * 128>>>(tm, 0, 0, 0, 0, 0, oh, dout);
          CK(cudaDeviceSynchronize());
          CK(cudaMemcpy(hout.data(), dout, ROWS * 4 * 2, cudaMemcpyDeviceToHost));
          char title[256];
          snprintf(title, sizeof title, "[stem view CW=9 CH=6 OW=6] tap h=%d  (expect rows = cell (y=p+tap, x=q))", oh);
          dump(title, hout);
        }
      }
    }
  }
  {  // ---------------- 5-D: N=1,D=3,H=4,W=4 ----------------
    const int N = 1, D = 3, H = 4, W = 4;
    std::vector<__nv_bfloat16> x((size_t)N * D * H * W * C, __float2bfloat16(0.f));
    for (int d = 0; d < D; ++d)
      for (int h = 0; h < H; ++h)
        for (int w = 0; w < W; ++w) {
          __nv_bfloat16* p = &x[((((size_t)0 * D + d) * H + h) * W + w) * C];
          p[0] = __float2bfloat16((float)d); p[1] = __float2bfloat16((float)h);
          p[2] = __float2bfloat16((float)w); p[3] = __float2bfloat16(1.f);
        }
    __nv_bfloat16* dx;
    CK(cudaMalloc(&dx, x.size() * 2));
    CK(cudaMemcpy(dx, x.data(), x.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tm;
    cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)N};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                             (cuuint64_t)D * H * W * C * 2};
    int lower[3] = {-1, -1, -1}, upper[3] = {-1, -1, -1};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dx, dims, strides, lower, upper, 64, ROWS, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode 5d -> %d\n", (int)r);
    if (r == CUDA_SUCCESS) {
      int taps[2][3] = {{1, 1, 1}, {0, 2, 1}};
      for (auto& t : taps) {
        CK(cudaMemset(dout, 0xFF, ROWS * 4 * 2));
        probe5d<<<1, 32, 1024 + ROWS * 128>>>(tm, 0, -1, -1, -1, 0, t[0], t[1], t[2], dout);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hout.data(), dout, ROWS * 4 * 2, cudaMemcpyDeviceToHost));
        char title[256];
        snprintf(title, sizeof title, "[5d D=3 H=4 W=4 3x3x3 pad1] tap (w=%d,h=%d,d=%d); rows print (d,h,w,valid)", t[0], t[1], t[2]);
        dump(title, hout);
      }
    }
  }
  printf("probe done\n");
  return 0;
}
