// wgrad_umma.cu -- convolution weight gradient on tcgen05 tensor cores (sm_100a).  See wgrad_umma.cuh.
//
// One CTA = one accumulator D[128 output channels][n_tile input channels] for one filter tap, summed over this CTA's
// share of the output positions (chunks of 128 positions, chunk j = blockIdx.x, + splits, ...):
//   warp 0        TMA producer: per chunk two 2-D boxes of dY (64 channels x 128 positions each) and n_tile/64 im2col
//                 boxes of X at this tap (64 channels x 128 positions), into a 3-4 stage mbarrier ring
//   warp 1        MMA issuer / TMEM owner: 8 x tcgen05.mma (K = 16 positions) per chunk, BOTH operands MN-major
//                 (a "row" of 128 bytes in shared memory is one position holding 64 channels): descriptor with
//                 LBO = distance between 64-channel boxes, SBO = 1024 B = 8 positions (canonical SW128 MN-major layout
//                 ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
//   warps 2-5     epilogue: tcgen05.ld -> red.global.add.f32 into scratch[tap][cin][cout] (coalesced along cout)
// Padding taps and the ragged last chunk need no special case: TMA fills out-of-bounds positions / channels with zeros.
#include "wgrad_umma.cuh"

namespace eco {
namespace {

constexpr int kWgThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s: fail loudly instead of hanging the GPU
      if (err) atomicExch(err, code);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h, int n,
                                              uint16_t ow, uint16_t oh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
      : "memory");
}
__device__ __forceinline__ void tma_im2col_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h, int d,
                                              int n, uint16_t ow, uint16_t oh, uint16_t od) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"(ow), "h"(oh),
      "h"(od)
      : "memory");
}
// MN-major, 128-byte-swizzled operand: 64 channels (128 B) contiguous per position, 8-position groups 1024 B apart (SBO),
// 64-channel boxes `lbo_bytes` apart (LBO)
__device__ __forceinline__ uint64_t make_mn_sw128_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// kind::f16: D = f32, A = B = bf16, A and B MN-major (bits 15 / 16), M = 128, N = n
__device__ __forceinline__ uint32_t make_idesc_mn(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const WgradParams p, const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int S = p.stages;
  const int NT = p.n_tile;
  const uint32_t a_bytes = 32768u;                       // dY: 2 boxes of 64 channels x 128 positions
  const uint32_t b_bytes = (uint32_t)(NT / 64) * 16384u;  // X at this tap: NT/64 boxes
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)S * stage_bytes);
  const uint32_t bar_full = smem_u32(bars);          // [S]
  const uint32_t bar_empty = bar_full + 8 * S;       // [S]
  const uint32_t bar_done = bar_empty + 8 * S;       // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int taps = p.KD * p.KH * p.KW;
  const int tap = (int)blockIdx.y % taps;
  const int cin_tile = (int)blockIdx.y / taps;
  const int cout0 = (int)blockIdx.z * 128;
  const int cin0 = cin_tile * NT;
  const int kx = tap % p.KW, ky = (tap / p.KW) % p.KH, kz = tap / (p.KW * p.KH);
  const int total_chunks = (p.M + 127) / 128;
  const int first = (int)blockIdx.x, step = p.splits;
  const int my_chunks = first < total_chunks ? (total_chunks - first + step - 1) / step : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)NT)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (my_chunks > 0) {
    if (warp == 0) {
      // ===================== TMA producer =====================
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < my_chunks; ++i) {
        const int m0 = (first + i * step) * 128;
        int r = m0;
        const int q = r % p.OW; r /= p.OW;
        const int pp = r % p.OH; r /= p.OH;
        const int z = r % p.OD;
        const int n = r / p.OD;
        const int cw = q * p.sW - p.pW, chh = pp * p.sH - p.pH, cd = z * p.sD - p.pD;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u, p.error_flag, 11);
        if (elect_one()) {
          const uint32_t st = base + s * stage_bytes;
          mbar_arrive_expect_tx(bar_full + 8 * s, stage_bytes);
          tma_load_2d(st, &tmY, bar_full + 8 * s, cout0, m0);
          tma_load_2d(st + 16384u, &tmY, bar_full + 8 * s, cout0 + 64, m0);
          for (int h = 0; h < NT / 64; ++h) {
            const uint32_t dst = st + a_bytes + (uint32_t)h * 16384u;
            if (p.nsp == 3)
              tma_im2col_5d(dst, &tmX, bar_full + 8 * s, cin0 + h * 64, cw, chh, cd, n, (uint16_t)kx, (uint16_t)ky, (uint16_t)kz);
            else
              tma_im2col_4d(dst, &tmX, bar_full + 8 * s, cin0 + h * 64, cw, chh, n, (uint16_t)kx, (uint16_t)ky);
          }
        }
        if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = make_idesc_mn(NT);
      uint32_t s = 0, ph = 0;
      for (int i = 0; i < my_chunks; ++i) {
        mbar_wait(bar_full + 8 * s, ph, p.error_flag, 12);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = base + s * stage_bytes;
          const uint64_t ad = make_mn_sw128_desc(st, 16384u);
          const uint64_t bd = make_mn_sw128_desc(st + a_bytes, 16384u);
#pragma unroll
          for (int ks = 0; ks < 8; ++ks)  // 16 positions = two 8-position groups = 2048 B per K step
            umma_bf16(tmem_base, ad + (uint64_t)(ks * 128), bd + (uint64_t)(ks * 128), idesc, (uint32_t)((i | ks) != 0));
          umma_commit(bar_empty + 8 * s);
        }
        if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(bar_done);
    } else {
      // ===================== epilogue: TMEM -> split-K reduction in global memory =====================
      const int wq = warp & 3;
      const int row = wq * 32 + lane;  // output channel within the tile (TMEM lane)
      const int co = cout0 + row;
      mbar_wait(bar_done, 0, p.error_flag, 13);
      tc_fence_after();
      float* dst = p.scratch + ((size_t)tap * p.cin_ld + cin0) * p.cout_ld + co;
      for (int c0 = 0; c0 < NT; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, v);
        if (co < p.Cout) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (cin0 + c0 + j < p.Cin) atomicAdd(dst + (size_t)(c0 + j) * p.cout_ld, __uint_as_float(v[j]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)NT) : "memory");
  }
}

}  // namespace

cudaError_t wgrad_umma_configure() {
  return cudaFuncSetAttribute(wgrad_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_wgrad_umma(const WgradParams& p, const CUtensorMap& tmY, const CUtensorMap& tmX, cudaStream_t stream) {
  const int taps = p.KD * p.KH * p.KW;
  dim3 grid((unsigned)p.splits, (unsigned)(taps * p.cin_tiles), (unsigned)p.cout_tiles);
  wgrad_umma_kernel<<<grid, kWgThreads, wgrad_smem_bytes(p), stream>>>(p, tmY, tmX);
  return cudaGetLastError();
}

}  // namespace eco
