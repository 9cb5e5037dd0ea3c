// conv_umma.cu -- fused implicit-GEMM convolution on tcgen05 tensor cores (sm_100a only).
//
// Tile: 128 output positions (TMEM lanes) x block_n output channels (TMEM fp32 columns),
// K consumed in 64-element blocks (one tap x 64 channels) through an mbarrier ring:
//
//   warp 0 (1 lane)  : TMA producer -- weights [Cout, K] K-major via cp.async.bulk.tensor.2d,
//                      activations via cp.async.bulk.tensor.{4,5}d im2col (A_TMA_IM2COL)
//   warps 2-5        : A_GATHER mode: cp.async zero-filling software im2col into the same
//                      128B-swizzled layout, completion tracked by cp.async.mbarrier.arrive.noinc;
//                      afterwards (both modes) the epilogue: tcgen05.ld -> +bias (+residual)
//                      -> optional raw store -> BN scale/shift -> ReLU -> bf16 store into a
//                      channel slice of the destination (concat fusion)
//   warp 1 (1 lane)  : tcgen05.mma issuer, accumulators in TMEM; owns TMEM alloc/dealloc
//
// Every mbarrier wait is bounded: a wait that exceeds ~2 s sets *error_flag and traps, so a
// protocol bug fails loudly instead of hanging the GPU.
#include <cstdio>
#include "conv_umma.cuh"

namespace eco {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
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
    if (clock64() - t0 > 4000000000LL) {
      if (err) atomicExch(err, code);
      __trap();
    }
  }
}
// (development counters) mbarrier.try_wait may block for a hardware-defined time before it returns, so the probe
// that decides "did we have to wait" is the non-blocking test_wait
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ long long mbar_wait_timed(uint32_t bar, uint32_t parity, int* err, int code) {
  if (mbar_test_wait(bar, parity)) return 0;
  const long long t0 = clock64();
  mbar_wait(bar, parity, err, code);
  return clock64() - t0;
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_im2col_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w,
                                              int h, int n, uint16_t ow, uint16_t oh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh)
      : "memory");
}
__device__ __forceinline__ void tma_im2col_5d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w,
                                              int h, int d, int n, uint16_t ow, uint16_t oh, uint16_t od) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n), "h"(ow),
      "h"(oh), "h"(od)
      : "memory");
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (sm_100 "version 1"):
//   rows are 128 B apart, 8-row groups 1024 B apart (SBO), LBO unused (=1) for swizzled K-major.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address, 16-byte units
  d |= (uint64_t)1 << 16;                        // leading byte offset (ignored for SW128 K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset
  d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                        // layout type: SWIZZLE_128B
  return d;
}
// K-major, 32-byte-swizzled operand (rows of 32 B = 16 bf16, 8-row groups 256 B apart)
__device__ __forceinline__ uint64_t make_sw32_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)6 << 61;  // SWIZZLE_32B
  return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=block_n
__device__ __forceinline__ uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBlockM >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
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

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void store16_bf16(__nv_bfloat16* dst, const float (&f)[16], int nvalid) {
  if (nvalid >= 16) {
    uint4 a, b;
    a.x = pack_bf16x2(f[0], f[1]);   a.y = pack_bf16x2(f[2], f[3]);
    a.z = pack_bf16x2(f[4], f[5]);   a.w = pack_bf16x2(f[6], f[7]);
    b.x = pack_bf16x2(f[8], f[9]);   b.y = pack_bf16x2(f[10], f[11]);
    b.z = pack_bf16x2(f[12], f[13]); b.w = pack_bf16x2(f[14], f[15]);
    if ((reinterpret_cast<uintptr_t>(dst) & 31u) == 0) {
      // one 256-bit store = one whole 32-byte sector per request (STG.E.ENL2.256 on sm_100a)
      asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(a.x), "r"(a.y), "r"(a.z),
                   "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                   : "memory");
    } else {
      reinterpret_cast<uint4*>(dst)[0] = a;
      reinterpret_cast<uint4*>(dst)[1] = b;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) dst[j] = __float2bfloat16_rn(f[j]);
  }
}
__device__ __forceinline__ void load16_bf16_add(const __nv_bfloat16* src, float (&f)[16], int nvalid) {
  if (nvalid >= 16) {
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(src));
    const uint4 b = __ldg(reinterpret_cast<const uint4*>(src) + 1);
    const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
      f[2 * j] += __low2float(t);
      f[2 * j + 1] += __high2float(t);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nvalid) f[j] += __bfloat162float(src[j]);
  }
}

// split-precision store: hi = bf16(f), lo = bf16(f - hi); planes [hi | lo | hi], `seg` elements apart
__device__ __forceinline__ void store16_split(__nv_bfloat16* dst, long long seg, const float (&f)[16], int nvalid) {
  float lo[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) lo[j] = f[j] - __bfloat162float(__float2bfloat16_rn(f[j]));
  store16_bf16(dst, f, nvalid);
  store16_bf16(dst + seg, lo, nvalid);
  store16_bf16(dst + 2 * seg, f, nvalid);
}

// Exactly one lane of a converged warp.  The tcgen05 / TMA instructions take uniform-register operands: issued
// under `if (lane == 0)` ptxas cannot prove a single active thread and wraps EVERY such instruction in an
// ELECT / BRA.U.ANY serialisation loop (seen in the SASS, ~60 issue cycles per tcgen05.mma whatever N, measured
// in tools/probe_commit.cu); under elect.sync it emits the bare instruction.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// weight-tile half, multicast to both CTAs of a cluster pair (lands at the same offset in each, signals the mbarrier
// at the same offset in each)
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "h"(mask), "r"(c0), "r"(c1) : "memory");
}

// One K block of the main loop (call from the elected lane): up to 4 K steps into accumulator 0 and, if `live1`,
// into accumulator 1 (second 128-row half sharing the weight tile), then tcgen05.commit -> `empty_bar`.
template <bool MC = false>  // MC: the commit frees the stage in BOTH CTAs of the cluster pair (multicast weight tiles)
__device__ __forceinline__ void mma_kblock(uint32_t acc0, uint32_t acc1, uint64_t ad0, uint64_t ad1, uint64_t bd,
                                           uint32_t accflag, uint32_t ks, uint32_t live1, uint32_t idesc,
                                           uint32_t empty_bar) {
  if (MC) {
    asm volatile(
      "{\n\t"
      ".reg .pred pacc, ptrue, g1, g2, g3, h0, h1, h2, h3;\n\t"
      ".reg .b64 a, b;\n\t"
      ".reg .b16 m;\n\t"
      "setp.ne.b32 pacc, %5, 0;\n\t"
      "setp.eq.u32 ptrue, %8, %8;\n\t"
      "setp.gt.u32 g1, %6, 1;\n\t"
      "setp.gt.u32 g2, %6, 2;\n\t"
      "setp.gt.u32 g3, %6, 3;\n\t"
      "setp.ne.b32 h0, %7, 0;\n\t"
      "and.pred h1, h0, g1;\n\t"
      "and.pred h2, h0, g2;\n\t"
      "and.pred h3, h0, g3;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %4, %8, pacc;\n\t"
      "add.s64 a, %2, 2;\n\t"
      "add.s64 b, %4, 2;\n\t"
      "@g1 tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %8, ptrue;\n\t"
      "add.s64 a, %2, 4;\n\t"
      "add.s64 b, %4, 4;\n\t"
      "@g2 tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %8, ptrue;\n\t"
      "add.s64 a, %2, 6;\n\t"
      "add.s64 b, %4, 6;\n\t"
      "@g3 tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %8, ptrue;\n\t"
      "@h0 tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %4, %8, pacc;\n\t"
      "add.s64 a, %3, 2;\n\t"
      "add.s64 b, %4, 2;\n\t"
      "@h1 tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %8, ptrue;\n\t"
      "add.s64 a, %3, 4;\n\t"
      "add.s64 b, %4, 4;\n\t"
      "@h2 tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %8, ptrue;\n\t"
      "add.s64 a, %3, 6;\n\t"
      "add.s64 b, %4, 6;\n\t"
      "@h3 tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %8, ptrue;\n\t"
      "mov.b16 m, 3;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%9], m;\n\t"
      "}"
      ::"r"(acc0), "r"(acc1), "l"(ad0), "l"(ad1), "l"(bd), "r"(accflag), "r"(ks), "r"(live1), "r"(idesc), "r"(empty_bar)
      : "memory");
    return;
  }
  asm volatile(
      "{\n\t"
      ".reg .pred pacc, ptrue, g1, g2, g3, h0, h1, h2, h3;\n\t"
      ".reg .b64 a, b;\n\t"
      "setp.ne.b32 pacc, %5, 0;\n\t"
      "setp.eq.u32 ptrue, %8, %8;\n\t"
      "setp.gt.u32 g1, %6, 1;\n\t"
      "setp.gt.u32 g2, %6, 2;\n\t"
      "setp.gt.u32 g3, %6, 3;\n\t"
      "setp.ne.b32 h0, %7, 0;\n\t"
      "and.pred h1, h0, g1;\n\t"
      "and.pred h2, h0, g2;\n\t"
      "and.pred h3, h0, g3;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %4, %8, pacc;\n\t"
      "add.s64 a, %2, 2;\n\t"
      "add.s64 b, %4, 2;\n\t"
      "@g1 tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %8, ptrue;\n\t"
      "add.s64 a, %2, 4;\n\t"
      "add.s64 b, %4, 4;\n\t"
      "@g2 tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %8, ptrue;\n\t"
      "add.s64 a, %2, 6;\n\t"
      "add.s64 b, %4, 6;\n\t"
      "@g3 tcgen05.mma.cta_group::1.kind::f16 [%0], a, b, %8, ptrue;\n\t"
      "@h0 tcgen05.mma.cta_group::1.kind::f16 [%1], %3, %4, %8, pacc;\n\t"
      "add.s64 a, %3, 2;\n\t"
      "add.s64 b, %4, 2;\n\t"
      "@h1 tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %8, ptrue;\n\t"
      "add.s64 a, %3, 4;\n\t"
      "add.s64 b, %4, 4;\n\t"
      "@h2 tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %8, ptrue;\n\t"
      "add.s64 a, %3, 6;\n\t"
      "add.s64 b, %4, 6;\n\t"
      "@h3 tcgen05.mma.cta_group::1.kind::f16 [%1], a, b, %8, ptrue;\n\t"
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t"
      "}"
      ::"r"(acc0), "r"(acc1), "l"(ad0), "l"(ad1), "l"(bd), "r"(accflag), "r"(ks), "r"(live1), "r"(idesc), "r"(empty_bar)
      : "memory");
}

// One epilogue warp's share of a tile: TMEM lane = output position `m`, columns = output channels.
//   raw = acc + bias (+ residual) -> optional bf16 store ; y = relu?(raw*scale + shift) -> bf16 store
__device__ __forceinline__ void epilogue_rows(const ConvKernelParams& p, uint32_t taddr, int m, bool row_ok, int n0,
                                              int BN, const float* s_bias, const float* s_scale,
                                              const float* s_shift) {
  const bool has_scale = p.scale != nullptr;
  for (int c0 = 0; c0 < BN; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + (uint32_t)c0, v);
    const int cg = n0 + c0;
    const int nvalid = p.Cout - cg;
    if (row_ok && nvalid > 0) {
      float f[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c0 + j];
      if (p.res) {
        const __nv_bfloat16* r = p.res + (long long)m * p.res_cs + p.res_coff + cg;
        load16_bf16_add(r, f, nvalid);
        if (p.res_seg) load16_bf16_add(r + p.res_seg, f, nvalid);  // + lo plane
      }
      if (p.raw) {
        __nv_bfloat16* r = p.raw + (long long)m * p.raw_cs + p.raw_coff + cg;
        if (p.raw_seg) store16_split(r, p.raw_seg, f, nvalid);
        else store16_bf16(r, f, nvalid);
      }
      if (p.out) {
        if (has_scale) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaf(f[j], s_scale[c0 + j], s_shift[c0 + j]);
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        __nv_bfloat16* o = p.out + (long long)m * p.out_cs + p.out_coff + cg;
        if (p.out_seg) store16_split(o, p.out_seg, f, nvalid);
        else store16_bf16(o, f, nvalid);
      }
    }
  }
}

__global__ void __launch_bounds__(kConvThreads, 1)
conv_umma_kernel(const ConvKernelParams p, const __grid_constant__ CUtensorMap tmA,
                 const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  const int S = p.stages;
  const int BN = p.block_n;
  const uint32_t a_stage_bytes = kBlockM * 128;
  const uint32_t b_stage_bytes = (uint32_t)BN * 128;
  const uint32_t sA = base;
  const uint32_t sB = sA + S * a_stage_bytes;
  float* s_bias = reinterpret_cast<float*>(smem + (size_t)S * a_stage_bytes + (size_t)S * b_stage_bytes);
  float* s_scale = s_bias + 256;
  float* s_shift = s_scale + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
  const uint32_t bar_full = smem_u32(bars);             // [S]
  const uint32_t bar_empty = bar_full + 8 * S;          // [S]
  const uint32_t bar_tmem_full = bar_empty + 8 * S;     // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kBlockM;
  const int n0 = blockIdx.y * BN;
  const bool tma_a = (p.a_mode == A_TMA_IM2COL);

  if (threadIdx.x == 0) {
    const uint32_t full_count = tma_a ? 1u : 1u + 128u;
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, full_count);
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    // stage per-channel epilogue constants for this N tile
    for (int i = threadIdx.x - 64; i < BN; i += 128) {
      const int c = n0 + i;
      const bool ok = c < p.Cout;
      s_bias[i] = (ok && p.bias) ? p.bias[c] : 0.f;
      s_scale[i] = (ok && p.scale) ? p.scale[c] : 1.f;
      s_shift[i] = (ok && p.scale) ? p.shift[c] : 0.f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_kb = p.num_kb;

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the ring, one elected lane issues) =====================
    {
      int q = 0, pp = 0, z = 0, n = 0;
      if (tma_a) {
        int t = m0;
        q = t % p.OW; t /= p.OW;
        pp = t % p.OH; t /= p.OH;
        z = t % p.OD; n = t / p.OD;
      }
      const int cw = q * p.sW - p.pW, chh = pp * p.sH - p.pH, cd = z * p.sD - p.pD;
      int cb = 0, kx = 0, ky = 0, kz = 0;
      const uint32_t tx_bytes = b_stage_bytes + (tma_a ? a_stage_bytes : 0u);
      uint32_t s = 0, ph = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1u, p.error_flag, 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(bar_full + 8 * s, tx_bytes);
          tma_load_2d(sB + s * b_stage_bytes, &tmB, bar_full + 8 * s, kb * kBlockK, n0);
          if (tma_a) {
            if (p.nsp == 3)
              tma_im2col_5d(sA + s * a_stage_bytes, &tmA, bar_full + 8 * s, cb * kBlockK, cw, chh, cd, n,
                            (uint16_t)kx, (uint16_t)ky, (uint16_t)kz);
            else
              tma_im2col_4d(sA + s * a_stage_bytes, &tmA, bar_full + 8 * s, cb * kBlockK, cw, chh, n,
                            (uint16_t)kx, (uint16_t)ky);
          }
        }
        if (++cb == p.cblocks) {
          cb = 0;
          if (++kx == p.KW) { kx = 0; if (++ky == p.KH) { ky = 0; ++kz; } }
        }
        if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the ring, one elected lane issues) =====================
    {
      const uint32_t idesc = make_idesc(BN);
      // the last channel block of a tap may hold fewer than 64 channels (Cin = 96: 64 + 32): its all-zero K steps are skipped
      const uint32_t tail_k = (uint32_t)(((p.Cin & (kBlockK - 1)) + kUmmaK - 1) / kUmmaK);  // 0: every block is full
      const uint64_t adesc0 = make_sw128_desc(sA), bdesc0 = make_sw128_desc(sB);
      const uint32_t a_step = a_stage_bytes >> 4, b_step = b_stage_bytes >> 4;
      uint32_t s = 0, ph = 0, cb = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_full + 8 * s, ph, p.error_flag, 2);
        if (!tma_a) fence_proxy_async_smem();  // cp.async (generic proxy) writes -> async-proxy reads
        tc_fence_after();
        uint32_t ks = 4;
        if (tail_k) {
          if (++cb == (uint32_t)p.cblocks) { cb = 0; ks = tail_k; }
        }
        if (elect_one())
          mma_kblock(tmem_base, tmem_base, adesc0 + (uint64_t)(s * a_step), 0, bdesc0 + (uint64_t)(s * b_step),
                     (uint32_t)(kb != 0), ks, 0u, idesc, bar_empty + 8 * s);
        if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(bar_tmem_full);  // accumulator complete
    }
  } else {
    // ===================== gather producers (A_GATHER) then epilogue =====================
    const int wq = warp & 3;
    const int row = wq * 32 + lane;
    const int m = m0 + row;
    const bool row_ok = m < p.M;
    if (!tma_a) {
      int q = 0, pp = 0, z = 0, n = 0;
      if (row_ok) {
        int t = m;
        q = t % p.OW; t /= p.OW;
        pp = t % p.OH; t /= p.OH;
        z = t % p.OD; n = t / p.OD;
      }
      const int x0 = q * p.sW - p.pW, y0 = pp * p.sH - p.pH, z0 = z * p.sD - p.pD;
      const __nv_bfloat16* xn = p.x + (long long)n * p.x_sN;
      const uint32_t row_off = (uint32_t)row * 128u;
      const uint32_t sw = (uint32_t)(row & 7);
      int cb = 0, kx = 0, ky = 0, kz = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S;
        const uint32_t ph = (uint32_t)(kb / S) & 1u;
        mbar_wait(bar_empty + 8 * s, ph ^ 1u, p.error_flag, 3);
        const int iz = z0 + kz, iy = y0 + ky, ix = x0 + kx;
        const bool ok = row_ok && (unsigned)iz < (unsigned)p.ID && (unsigned)iy < (unsigned)p.IH &&
                        (unsigned)ix < (unsigned)p.IW;
        const __nv_bfloat16* src =
            ok ? xn + (long long)iz * p.x_sD + (long long)iy * p.x_sH + (long long)ix * p.x_sW + cb * kBlockK : p.x;
        const int crem = p.Cin - cb * kBlockK;  // channels left in this tap
        const uint32_t dst = sA + s * a_stage_bytes + row_off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool cok = ok && (j * 8 < crem);
          cp_async_16(dst + ((((uint32_t)j) ^ sw) << 4), cok ? (const void*)(src + j * 8) : (const void*)p.x,
                      cok ? 16u : 0u);
        }
        cp_async_mbar_arrive_noinc(bar_full + 8 * s);
        if (++cb == p.cblocks) {
          cb = 0;
          if (++kx == p.KW) { kx = 0; if (++ky == p.KH) { ky = 0; ++kz; } }
        }
      }
    }
    // ---------------- epilogue ----------------
    mbar_wait(bar_tmem_full, 0, p.error_flag, 4);
    tc_fence_after();
    epilogue_rows(p, tmem_base + ((uint32_t)(wq * 32) << 16), m, row_ok, n0, BN, s_bias, s_scale, s_shift);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}


// ------------------------------------------------------------------------------------------------
// Register-resident epilogue used by the persistent and the halo kernels.
// The destination pointers / strides are copied into registers once per kernel (the "memory" clobbers of
// the surrounding inline PTX would otherwise make the compiler re-read them from the constant bank in
// every chunk), 32 columns are processed per TMEM load, and two loads are in flight before the wait,
// so each epilogue warp has 64 independent values to work on instead of a 16-value dependent chain.
struct EpiArgs {
  __nv_bfloat16* out; long long out_cs; int out_coff;
  __nv_bfloat16* raw; long long raw_cs; int raw_coff;
  const __nv_bfloat16* res; long long res_cs; int res_coff;
  int Cout, relu, simple;
};

__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// finish NV = 16 or 32 accumulator columns of one output position: v -> (raw) -> y -> global
template <int NV>
__device__ __forceinline__ void epi_finish(const EpiArgs& e, const uint32_t (&v)[NV], long long m, bool row_ok, int cg,
                                           int c0, const float* s_bias, const float* s_scale, const float* s_shift) {
  const int nvalid = e.Cout - cg;
  if (!row_ok || nvalid <= 0) return;
  float f[NV];
  const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
  const float4* sh4 = reinterpret_cast<const float4*>(s_shift + c0);
  if (e.simple) {
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const float4 a = sc4[j], b = sh4[j];
      f[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), a.x, b.x);
      f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), a.y, b.y);
      f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), a.z, b.z);
      f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), a.w, b.w);
    }
  } else {
    const float4* bi4 = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const float4 b = bi4[j];
      f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
      f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
      f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
      f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
    }
#pragma unroll
    for (int q = 0; q < NV / 16; ++q) {
      float (&fq)[16] = *reinterpret_cast<float (*)[16]>(&f[16 * q]);
      if (e.res) load16_bf16_add(e.res + m * e.res_cs + e.res_coff + cg + 16 * q, fq, nvalid - 16 * q);
      if (e.raw) store16_bf16(e.raw + m * e.raw_cs + e.raw_coff + cg + 16 * q, fq, nvalid - 16 * q);
    }
    if (!e.out) return;
#pragma unroll
    for (int j = 0; j < NV / 4; ++j) {
      const float4 a = sc4[j], b = sh4[j];
      f[4 * j + 0] = fmaf(f[4 * j + 0], a.x, b.x);
      f[4 * j + 1] = fmaf(f[4 * j + 1], a.y, b.y);
      f[4 * j + 2] = fmaf(f[4 * j + 2], a.z, b.z);
      f[4 * j + 3] = fmaf(f[4 * j + 3], a.w, b.w);
    }
  }
  if (e.relu) {
#pragma unroll
    for (int j = 0; j < NV; ++j) f[j] = fmaxf(f[j], 0.f);
  }
#pragma unroll
  for (int q = 0; q < NV / 16; ++q) {
    float (&fq)[16] = *reinterpret_cast<float (*)[16]>(&f[16 * q]);
    store16_bf16(e.out + m * e.out_cs + e.out_coff + cg + 16 * q, fq, nvalid - 16 * q);
  }
}

// One epilogue warp's work for one tile: MT halves x its column range [c_begin, c_end) (16-column chunks).
//   taddr_h0: TMEM address of (lane quarter, accumulator half 0, column 0); halves are BN columns apart.
//   m_of(h), ok_of(h): output position / validity of this thread's row in half h.
template <int MT, typename RowFn>
__device__ __forceinline__ void epilogue_tile(const EpiArgs& e, uint32_t taddr_h0, int BN, int n0, int c_begin, int c_end,
                                              const float* s_bias, const float* s_scale, const float* s_shift,
                                              RowFn row_of) {
  long long mrow[MT];
  bool rok[MT];
#pragma unroll
  for (int h = 0; h < MT; ++h) row_of(h, mrow[h], rok[h]);
  const int npairs = (c_end - c_begin) >> 1;       // 32-column groups
  const bool tail16 = ((c_end - c_begin) & 1) != 0;
  const int items = MT * npairs;
  for (int i = 0; i < items; i += 2) {
    uint32_t va[32], vb[32];
    const int h0 = i / npairs, g0 = i - h0 * npairs;
    const bool two = (i + 1) < items;
    const int h1 = two ? (i + 1) / npairs : h0, g1 = two ? (i + 1) - h1 * npairs : g0;
    const int ca = c_begin + 2 * g0, cb = c_begin + 2 * g1;
    tmem_ld32_nowait(taddr_h0 + (uint32_t)(h0 * BN + ca * 16), va);
    if (two) tmem_ld32_nowait(taddr_h0 + (uint32_t)(h1 * BN + cb * 16), vb);
    tmem_wait_ld();
    epi_finish<32>(e, va, mrow[h0], rok[h0], n0 + ca * 16, ca * 16, s_bias, s_scale, s_shift);
    if (two) epi_finish<32>(e, vb, mrow[h1], rok[h1], n0 + cb * 16, cb * 16, s_bias, s_scale, s_shift);
  }
  if (tail16) {
    const int ct = c_end - 1;
#pragma unroll
    for (int h = 0; h < MT; ++h) {
      uint32_t v[16];
      tmem_ld16_nowait(taddr_h0 + (uint32_t)(h * BN + ct * 16), v);
      tmem_wait_ld();
      epi_finish<16>(e, v, mrow[h], rok[h], n0 + ct * 16, ct * 16, s_bias, s_scale, s_shift);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Persistent variant (A via TMA im2col only).  One CTA per SM walks the tile list.
//   * tile = (MT x 128) output positions x block_n channels; with MT = 2 the two 128-row halves share
//     every weight tile in shared memory (half the B traffic per FLOP, for block_n <= 128)
//   * TMEM accumulators are double-buffered (2 x MT x block_n fp32 columns <= 512): the epilogue of
//     tile i overlaps the MMAs of tile i+1; the smem ring streams across tile boundaries
//   * 10 warps: 0 = TMA producer, 1 = MMA issuer / TMEM owner, 2..9 = epilogue; two epilogue warps
//     per TMEM lane quarter split the 16-column chunks between them
constexpr int kPersistThreads = 320;  // persistent im2col kernel: warp 0 TMA producer, 1 MMA, 2-9 epilogue
constexpr int kHaloThreads = 384;     // halo kernel: + warps 10-11 (weight-tile producers)

// Computes one 16-column chunk of the epilogue for this thread's row.  `raw` (if any) is stored
// directly; the final value y is written as 32 bytes of bf16 into the warp's staging row at
// `stage_dst` (shared memory) -- the caller copies staged rows out with row-contiguous 16-byte
// stores so that every global store instruction covers whole 32-byte sectors.
__device__ __forceinline__ void epilogue_chunk(const ConvKernelParams& p, uint32_t taddr, int m, bool row_ok, int cg,
                                               int c0, const float* s_bias, const float* s_scale,
                                               const float* s_shift, bool simple, uint32_t stage_dst) {
  uint32_t v[16];
  if (p.debug_flags & 2) {  // development: no TMEM traffic
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = 0x3f800000u;
  } else {
    tmem_ld16(taddr, v);
  }
  if (p.debug_flags & 1) row_ok = false;  // development: compute, never store
  const int nvalid = p.Cout - cg;
  float f[16];
  const float4* sc4 = reinterpret_cast<const float4*>(s_scale + c0);
  const float4* sh4 = reinterpret_cast<const float4*>(s_shift + c0);
  if (simple) {
    // no raw / residual consumer: bias is pre-folded into the shift, y = relu?(acc * scale + shift')
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 a = sc4[j], b = sh4[j];
      f[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), a.x, b.x);
      f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), a.y, b.y);
      f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), a.z, b.z);
      f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), a.w, b.w);
    }
  } else {
    const float4* bi4 = reinterpret_cast<const float4*>(s_bias + c0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 b = bi4[j];
      f[4 * j + 0] = __uint_as_float(v[4 * j + 0]) + b.x;
      f[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + b.y;
      f[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + b.z;
      f[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + b.w;
    }
    if (row_ok && nvalid > 0) {
      if (p.res) load16_bf16_add(p.res + (long long)m * p.res_cs + p.res_coff + cg, f, nvalid);
      if (p.raw) store16_bf16(p.raw + (long long)m * p.raw_cs + p.raw_coff + cg, f, nvalid);
    }
    if (!p.out) return;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 a = sc4[j], b = sh4[j];
      f[4 * j + 0] = fmaf(f[4 * j + 0], a.x, b.x);
      f[4 * j + 1] = fmaf(f[4 * j + 1], a.y, b.y);
      f[4 * j + 2] = fmaf(f[4 * j + 2], a.z, b.z);
      f[4 * j + 3] = fmaf(f[4 * j + 3], a.w, b.w);
    }
  }
  if (p.nseg > 1) {
    // segmented output (fused sibling 1x1 convolutions): this 16-channel chunk lies inside one segment
    int sg = 0;
    while (sg < p.nseg - 1 && cg >= p.seg_end[sg]) ++sg;
    if (p.seg_relu[sg]) {
#pragma unroll
      for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (row_ok && nvalid > 0) store16_bf16(p.seg_ptr[sg] + (long long)m * p.seg_cs[sg] + p.seg_coff[sg] + cg, f, nvalid);
    return;
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (!p.epi_staged) {
    if (row_ok && nvalid > 0) store16_bf16(p.out + (long long)m * p.out_cs + p.out_coff + cg, f, nvalid);
    return;
  }
  const uint32_t w0 = pack_bf16x2(f[0], f[1]), w1 = pack_bf16x2(f[2], f[3]), w2 = pack_bf16x2(f[4], f[5]),
                 w3 = pack_bf16x2(f[6], f[7]), w4 = pack_bf16x2(f[8], f[9]), w5 = pack_bf16x2(f[10], f[11]),
                 w6 = pack_bf16x2(f[12], f[13]), w7 = pack_bf16x2(f[14], f[15]);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_dst), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_dst + 16), "r"(w4), "r"(w5), "r"(w6), "r"(w7)
               : "memory");
}

// Out-only epilogue of one 16-column chunk with everything it needs in registers (destination pointer already at
// this row and column, ReLU flag, valid-column count): y = relu?(acc * scale + shift') -> one 32-byte bf16 store.
// The generic epilogue_chunk re-reads its parameters (and, for fused sibling 1x1 convolutions, walks the segment
// table with dynamically indexed constant loads) for EVERY chunk; on the short-K 1x1 layers that made the
// epilogue the critical path (in-kernel counters: the MMA warp waited 70 % of the time for a free accumulator).
__device__ __forceinline__ void epilogue_chunk_simple(uint32_t taddr, __nv_bfloat16* dst, int nvalid, bool relu, bool store,
                                                      const float* sc, const float* sh) {
  // (one definition of v[] only: an alternative "pretend the accumulator is 1.0" path in here cost 32 register moves per
  // chunk -- 10 % of conv2_3x3's instructions in the r02 source-level profile)
  uint32_t v[16];
  tmem_ld16(taddr, v);
  float f[16];
  const float4* sc4 = reinterpret_cast<const float4*>(sc);
  const float4* sh4 = reinterpret_cast<const float4*>(sh);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 a = sc4[j], b = sh4[j];
    f[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), a.x, b.x);
    f[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), a.y, b.y);
    f[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), a.z, b.z);
    f[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), a.w, b.w);
  }
  if (relu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.f);
  }
  if (store && nvalid > 0) store16_bf16(dst, f, nvalid);
}

// This warp's chunk range [c_begin, c_end) of one 128-row accumulator, out-only epilogue, one or several output segments
__device__ __forceinline__ void epilogue_simple_range(const ConvKernelParams& p, uint32_t taddr, long long m, bool row_ok,
                                                      int n0, int c_begin, int c_end, const float* s_scale,
                                                      const float* s_shift) {
  const bool store = row_ok;
  const int cout = p.Cout;
  if (p.nseg <= 1) {
    __nv_bfloat16* dst = p.out + m * p.out_cs + p.out_coff + n0;
    const bool relu = p.relu != 0;
    for (int c = c_begin; c < c_end; ++c)
      epilogue_chunk_simple(taddr + (uint32_t)(c * 16), dst + c * 16, cout - (n0 + c * 16), relu, store, s_scale + c * 16,
                            s_shift + c * 16);
    return;
  }
  int seg_lo = 0;
  for (int sg = 0; sg < p.nseg; ++sg) {
    const int seg_hi = p.seg_end[sg];
    const int a = max(c_begin, (seg_lo - n0) >> 4), b = min(c_end, (seg_hi - n0) >> 4);
    seg_lo = seg_hi;
    if (a >= b) continue;
    __nv_bfloat16* dst = p.seg_ptr[sg] + m * p.seg_cs[sg] + p.seg_coff[sg] + n0;
    const bool relu = p.seg_relu[sg] != 0;
    for (int c = a; c < b; ++c)
      epilogue_chunk_simple(taddr + (uint32_t)(c * 16), dst + c * 16, cout - (n0 + c * 16), relu, store, s_scale + c * 16,
                            s_shift + c * 16);
  }
}

// MC = true: launched as clusters of 2 CTAs that work on two different M tiles of the SAME N tile; each CTA loads half
// of the weight tile and multicasts it into both shared memories (tmB then has box rows = block_n / 2), which
// removes a quarter to a third of the L2->SM bytes -- the co-limiter next to the UMMA operand reads
// (profiles/r01m: with the MMAs switched off the layers take the same time, 15-17.5 TB/s of TMA traffic).
template <int MT, bool MC>
__global__ void __launch_bounds__(kPersistThreads, 1)
conv_umma_persistent_kernel(const ConvKernelParams p, const __grid_constant__ CUtensorMap tmA,
                            const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  constexpr int TILE_M = MT * kBlockM;
  const int S = p.stages;
  const int BN = p.block_n;
  const uint32_t a_half_bytes = kBlockM * 128;
  const uint32_t a_stage_bytes = MT * a_half_bytes;
  const uint32_t b_stage_bytes = (uint32_t)BN * 128;
  const uint32_t sA = base;
  const uint32_t sB = sA + S * a_stage_bytes;
  float* s_bias = reinterpret_cast<float*>(smem + (size_t)S * a_stage_bytes + (size_t)S * b_stage_bytes);
  float* s_scale = s_bias + 256;
  float* s_shift = s_scale + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
  const uint32_t bar_full = smem_u32(bars);              // [S]
  const uint32_t bar_empty = bar_full + 8 * S;           // [S]
  const uint32_t bar_tmem_full = bar_empty + 8 * S;      // [2]
  const uint32_t bar_tmem_empty = bar_tmem_full + 16;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);
  // epilogue staging: 8 warps x 32 rows x (epi_group chunks x 32 B + 16 B pad)
  const uint32_t stage_pitch = (uint32_t)p.epi_group * 32u + 16u;
  const uint32_t stage_base = smem_u32(bars) + 512u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = (p.Cout + BN - 1) / BN;
  const int n_tiles_m = (p.M + TILE_M - 1) / TILE_M;
  const uint32_t acc_cols = (uint32_t)(MT * BN);  // TMEM columns of one accumulator buffer
  // work list: plain = tile t -> (m tile t / n_tiles_n, n tile t % n_tiles_n), CTA b takes t = b, b + grid, ...;
  // MC = cluster c takes "super tiles" u = c, c + clusters, ...: n tile u % n_tiles_n, m tiles 2 (u / n_tiles_n) + rank
  // (an odd last m tile leaves rank 1 with a dead tile: it still loads and multicasts its weight half)
  const uint32_t rank = MC ? cluster_ctarank() : 0u;
  const int total_tiles = MC ? n_tiles_n * ((n_tiles_m + 1) / 2) : n_tiles_n * n_tiles_m;
  const int t_first = MC ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int t_step = MC ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  auto tile_m_index = [&](int t) { return MC ? 2 * (t / n_tiles_n) + (int)rank : t / n_tiles_n; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, MC ? 2 : 1);  // MC: both CTAs of the pair must have released the stage
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tmem_full + 8 * b, 1);
      mbar_init(bar_tmem_empty + 8 * b, 8);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (MC) cluster_sync_all();  // the peer multicasts into this CTA's shared memory and barriers
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_kb = p.num_kb;
  const long long t_start = clock64();

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the ring, one elected lane issues) =====================
    {
      uint32_t s = 0, ph = 0;  // ring position, running across tiles (no division in the per-block loop)
      long long w_a = 0;
      for (int t = t_first; t < total_tiles; t += t_step) {
        const int n0 = (t % n_tiles_n) * BN;
        int cw[MT], chh[MT], cd[MT], cn[MT];
        bool live[MT];  // a half that starts beyond the last output position is never loaded
        uint32_t tile_tx = b_stage_bytes;
#pragma unroll
        for (int h = 0; h < MT; ++h) {
          int r = tile_m_index(t) * TILE_M + h * kBlockM;
          live[h] = r < p.M;
          if (live[h]) tile_tx += a_half_bytes;
          const int q = r % p.OW; r /= p.OW;
          const int pp = r % p.OH; r /= p.OH;
          const int z = r % p.OD;
          cn[h] = r / p.OD;
          cw[h] = q * p.sW - p.pW; chh[h] = pp * p.sH - p.pH; cd[h] = z * p.sD - p.pD;
        }
        int cb = 0, kx = 0, ky = 0, kz = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          w_a += mbar_wait_timed(bar_empty + 8 * s, ph ^ 1u, p.error_flag, 1);
          if (p.debug_flags & 8) {  // development: no TMA traffic at all, the MMAs read whatever the stage holds
            if (elect_one()) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full + 8 * s) : "memory");
          } else if (elect_one()) {
            mbar_arrive_expect_tx(bar_full + 8 * s, tile_tx);
            if (MC)
              tma_load_2d_mc(sB + s * b_stage_bytes + rank * (b_stage_bytes >> 1), &tmB, bar_full + 8 * s, kb * kBlockK,
                             n0 + (int)rank * (BN >> 1), (uint16_t)3);
            else
              tma_load_2d(sB + s * b_stage_bytes, &tmB, bar_full + 8 * s, kb * kBlockK, n0);
#pragma unroll
            for (int h = 0; h < MT; ++h) {
              if (!live[h]) continue;
              const uint32_t dst = sA + s * a_stage_bytes + h * a_half_bytes;
              if (p.nsp == 3)
                tma_im2col_5d(dst, &tmA, bar_full + 8 * s, cb * kBlockK, cw[h], chh[h], cd[h], cn[h], (uint16_t)kx,
                              (uint16_t)ky, (uint16_t)kz);
              else
                tma_im2col_4d(dst, &tmA, bar_full + 8 * s, cb * kBlockK, cw[h], chh[h], cn[h], (uint16_t)kx,
                              (uint16_t)ky);
            }
          }
          if (++cb == p.cblocks) {
            cb = 0;
            if (++kx == p.KW) { kx = 0; if (++ky == p.KH) { ky = 0; ++kz; } }
          }
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
      }
      if ((p.debug_flags & 16) && blockIdx.x == 0 && lane == 0)
        printf("persistent cta0: producer waited %lld cycles (stage free) of %lld\n", w_a, clock64() - t_start);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the ring, one elected lane issues) =====================
    {
      const uint32_t idesc = make_idesc(BN);
      const uint32_t tail_k = (uint32_t)(((p.Cin & (kBlockK - 1)) + kUmmaK - 1) / kUmmaK);  // 0: every block is full
      const uint64_t adesc0 = make_sw128_desc(sA), bdesc0 = make_sw128_desc(sB);
      const uint32_t a_step = a_stage_bytes >> 4, b_step = b_stage_bytes >> 4, a_half_step = a_half_bytes >> 4;
      uint32_t s = 0, ph = 0, tile_iter = 0;
      long long w_a = 0, w_b = 0;
      for (int t = t_first; t < total_tiles; t += t_step, ++tile_iter) {
        const uint32_t buf = tile_iter & 1u;
        const uint32_t use = tile_iter >> 1;
        w_b += mbar_wait_timed(bar_tmem_empty + 8 * buf, (use & 1u) ^ 1u, p.error_flag, 5);  // epilogue drained this buffer
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * acc_cols;
        // second 128-row half of the tile: dead when it starts beyond the last output position
        const uint32_t live1 = (MT == 2 && tile_m_index(t) * TILE_M + kBlockM < p.M) ? 1u : 0u;
        uint32_t cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          w_a += mbar_wait_timed(bar_full + 8 * s, ph, p.error_flag, 2);
          tc_fence_after();
          uint32_t ks = 4;
          if (tail_k) {
            if (++cb == (uint32_t)p.cblocks) { cb = 0; ks = tail_k; }
          }
          if (elect_one()) {
            const uint64_t ad = adesc0 + (uint64_t)(s * a_step);
            mma_kblock<MC>(acc, acc + (uint32_t)BN, ad, ad + a_half_step, bdesc0 + (uint64_t)(s * b_step), (uint32_t)(kb != 0), ks,
                           live1, idesc, bar_empty + 8 * s);
          }
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma_commit(bar_tmem_full + 8 * buf);
      }
      if ((p.debug_flags & 16) && blockIdx.x == 0 && lane == 0)
        printf("persistent cta0 [M=%d Cout=%d bn=%d kb=%d MT=%d]: mma waited %lld (stage full) + %lld (accumulator free) cycles of %lld, %u tiles\n",
               p.M, p.Cout, BN, num_kb, MT, w_a, w_b, clock64() - t_start, tile_iter);
    }
  } else {
    // ===================== epilogue warps (8) =====================
    const int wq = warp & 3;            // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;   // which of the two warps of that quarter
    const int chunks = BN >> 4;
    const bool simple = (p.res == nullptr) && (p.raw == nullptr) && (p.out != nullptr);
    uint32_t tile_iter = 0;
    long long w_e = 0;
    int loaded_n0 = -1;
    for (int t = t_first; t < total_tiles; t += t_step, ++tile_iter) {
      const int n0 = (t % n_tiles_n) * BN;
      const int m0 = tile_m_index(t) * TILE_M;
      if (n0 != loaded_n0) {  // uniform across the eight epilogue warps
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = threadIdx.x - 64; i < BN; i += 256) {
          const int c = n0 + i;
          const bool ok = c < p.Cout;
          const float bi = (ok && p.bias) ? p.bias[c] : 0.f;
          const float sc = (ok && p.scale) ? p.scale[c] : 1.f;
          const float sh = (ok && p.scale) ? p.shift[c] : 0.f;
          s_bias[i] = bi;
          s_scale[i] = sc;
          s_shift[i] = simple ? fmaf(bi, sc, sh) : sh;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        loaded_n0 = n0;
      }
      const uint32_t buf = tile_iter & 1u;
      const uint32_t use = tile_iter >> 1;
      w_e += mbar_wait_timed(bar_tmem_full + 8 * buf, use & 1u, p.error_flag, 4);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < MT; ++h) {
        const int mrow0 = m0 + h * kBlockM + wq * 32;  // first output position of this warp's 32 rows
        const int m = mrow0 + lane;
        const bool row_ok = m < p.M;
        const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + buf * acc_cols + (uint32_t)(h * BN);
        // this warp's contiguous chunk range: the two warps of a lane quarter split the columns in half
        const int c_begin = half ? (chunks + 1) / 2 : 0;
        const int c_end = half ? chunks : (chunks + 1) / 2;
        const uint32_t my_stage = stage_base + (uint32_t)(warp - 2) * 32u * stage_pitch;
        if (simple && !p.epi_staged) {
          epilogue_simple_range(p, taddr, (long long)m, row_ok, n0, c_begin, c_end, s_scale, s_shift);
          continue;
        }
        for (int cgrp = c_begin; cgrp < c_end; cgrp += p.epi_group) {
          const int gcount = min(p.epi_group, c_end - cgrp);
          for (int k = 0; k < gcount; ++k) {
            const int c = cgrp + k;
            epilogue_chunk(p, taddr + (uint32_t)(c * 16), m, row_ok, n0 + c * 16, c * 16, s_bias, s_scale, s_shift,
                           simple, my_stage + (uint32_t)lane * stage_pitch + (uint32_t)k * 32u);
          }
          if (p.out && p.epi_staged) {
            __syncwarp();
            // copy out: consecutive lanes take consecutive 16-byte pieces of a row -> whole sectors / lines
            const int ppr = gcount * 2;  // 16-byte pieces per row in this group
            const int col0 = n0 + cgrp * 16;
            for (int idx = lane; idx < 32 * ppr; idx += 32) {
              const int r = idx / ppr;
              const int piece = idx - r * ppr;
              const int mm = mrow0 + r;
              const int col = col0 + piece * 8;
              if (mm < p.M && col < p.Cout) {
                uint4 val;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                             : "r"(my_stage + (uint32_t)r * stage_pitch + (uint32_t)piece * 16u));
                *reinterpret_cast<uint4*>(p.out + (long long)mm * p.out_cs + p.out_coff + col) = val;
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tmem_empty + 8 * buf) : "memory");
      }
    }
    if ((p.debug_flags & 16) && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 9))
      printf("persistent cta0 warp %d: epilogue waited %lld cycles (accumulator full) of %lld\n", warp, w_e, clock64() - t_start);
  }

  tc_fence_before();
  if (MC) cluster_sync_all();  // no multicast write or commit may still be in flight towards an exited CTA
  else __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}


// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent kernel (cta_group::2): two CTAs of a cluster (one TPC) compute ONE
// 256 x block_n tile.  Each CTA loads the im2col rows of ITS 128 output positions and HALF of the weight tile
// (block_n / 2 rows); the leader CTA's single MMA thread issues tcgen05.mma.cta_group::2 (M = 256), which reads
// A and the B half from the shared memory of both SMs, and each SM accumulates its 128 rows x block_n in its own
// TMEM.  Per SM and K block the shared-memory traffic drops from 16 KB + 128 B x block_n (twice: TMA write + MMA
// read) to 16 KB + 64 B x block_n -- the bound of the cta_group::1 kernels.
//   barriers: full[s] lives on the leader (its expect_tx covers the bytes of both CTAs; the peer's TMA loads
//   signal it through the peer-bit-cleared address), empty[s] / tmem_full[b] exist in both CTAs and are signalled
//   by multicast tcgen05.commit, tmem_empty[b] lives on the leader and collects 16 epilogue-warp arrivals.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // shared::cluster address -> same offset in CTA rank 0 of the pair

__device__ __forceinline__ void mbar_arrive_on_cta(uint32_t local_bar, uint32_t cta) {
  asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
               "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(local_bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_im2col_4d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h,
                                                  int n, uint16_t ow, uint16_t oh) {
  asm volatile(
      "cp.async.bulk.tensor.4d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow),
      "h"(oh) : "memory");
}
__device__ __forceinline__ void tma_im2col_5d_2sm(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c, int w, int h,
                                                  int d, int n, uint16_t ow, uint16_t oh, uint16_t od) {
  asm volatile(
      "cp.async.bulk.tensor.5d.im2col.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], {%8, %9, %10};"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar & kPeerBitMask), "r"(c), "r"(w), "r"(h), "r"(d), "r"(n),
      "h"(ow), "h"(oh), "h"(od) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3) : "memory");
}
// K block of the pair kernel: up to 4 K steps of M = 256 MMAs, then the multicast commit that frees the stage in both CTAs
__device__ __forceinline__ void mma_kblock_pair(uint32_t acc, uint64_t ad, uint64_t bd, uint32_t accflag, uint32_t ks,
                                                uint32_t idesc, uint32_t empty_bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred pacc, ptrue, g1, g2, g3;\n\t"
      ".reg .b64 a, b;\n\t"
      "setp.ne.b32 pacc, %3, 0;\n\t"
      "setp.eq.u32 ptrue, %5, %5;\n\t"
      "setp.gt.u32 g1, %4, 1;\n\t"
      "setp.gt.u32 g2, %4, 2;\n\t"
      "setp.gt.u32 g3, %4, 3;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %5, pacc;\n\t"
      "add.s64 a, %1, 2;\n\t"
      "add.s64 b, %2, 2;\n\t"
      "@g1 tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, ptrue;\n\t"
      "add.s64 a, %1, 4;\n\t"
      "add.s64 b, %2, 4;\n\t"
      "@g2 tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, ptrue;\n\t"
      "add.s64 a, %1, 6;\n\t"
      "add.s64 b, %2, 6;\n\t"
      "@g3 tcgen05.mma.cta_group::2.kind::f16 [%0], a, b, %5, ptrue;\n\t"
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%6], %7;\n\t"
      "}"
      ::"r"(acc), "l"(ad), "l"(bd), "r"(accflag), "r"(ks), "r"(idesc), "r"(empty_bar), "h"((uint16_t)3)
      : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPersistThreads, 1)
conv_umma_pair_kernel(const ConvKernelParams p, const __grid_constant__ CUtensorMap tmA,
                      const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);

  constexpr int TILE_M = 2 * kBlockM;
  const int S = p.stages;
  const int BN = p.block_n;
  const uint32_t a_stage_bytes = kBlockM * 128;
  const uint32_t b_stage_bytes = (uint32_t)(BN / 2) * 128;  // this CTA's half of the weight tile
  const uint32_t sA = base;
  const uint32_t sB = sA + S * a_stage_bytes;
  float* s_bias = reinterpret_cast<float*>(smem + (size_t)S * a_stage_bytes + (size_t)S * b_stage_bytes);
  float* s_scale = s_bias + 256;
  float* s_shift = s_scale + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
  const uint32_t bar_full = smem_u32(bars);              // [S]  used on the leader
  const uint32_t bar_empty = bar_full + 8 * S;           // [S]  both CTAs (multicast commit)
  const uint32_t bar_tmem_full = bar_empty + 8 * S;      // [2]  both CTAs (multicast commit)
  const uint32_t bar_tmem_empty = bar_tmem_full + 16;    // [2]  used on the leader: 16 epilogue warps arrive
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair_id = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int n_tiles_n = (p.Cout + BN - 1) / BN;
  const int n_tiles_m = (p.M + TILE_M - 1) / TILE_M;
  const int total_tiles = n_tiles_n * n_tiles_m;
  const uint32_t acc_cols = (uint32_t)BN;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar_tmem_full + 8 * b, 1);
      mbar_init(bar_tmem_empty + 8 * b, 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_kb = p.num_kb;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; all loads signal the leader's full barrier) =====================
    uint32_t s = 0, ph = 0;
    for (int t = pair_id; t < total_tiles; t += npairs) {
      const int n0 = (t % n_tiles_n) * BN + (int)rank * (BN / 2);
      const int mt0 = (t / n_tiles_n) * TILE_M;
      int r = mt0 + (int)rank * kBlockM;
      const bool live = r < p.M;
      const bool peer_live = mt0 + kBlockM < p.M;
      const uint32_t tile_tx = 2u * b_stage_bytes + a_stage_bytes + (peer_live ? a_stage_bytes : 0u);  // leader's expectation
      const int q = r % p.OW; r /= p.OW;
      const int pp = r % p.OH; r /= p.OH;
      const int z = r % p.OD;
      const int cn = r / p.OD;
      const int cw = q * p.sW - p.pW, chh = pp * p.sH - p.pH, cd = z * p.sD - p.pD;
      int cb = 0, kx = 0, ky = 0, kz = 0;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(bar_empty + 8 * s, ph ^ 1u, p.error_flag, 1);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(bar_full + 8 * s, tile_tx);
          tma_load_2d_2sm(sB + s * b_stage_bytes, &tmB, bar_full + 8 * s, kb * kBlockK, n0);
          if (live) {
            const uint32_t dst = sA + s * a_stage_bytes;
            if (p.nsp == 3)
              tma_im2col_5d_2sm(dst, &tmA, bar_full + 8 * s, cb * kBlockK, cw, chh, cd, cn, (uint16_t)kx, (uint16_t)ky,
                                (uint16_t)kz);
            else
              tma_im2col_4d_2sm(dst, &tmA, bar_full + 8 * s, cb * kBlockK, cw, chh, cn, (uint16_t)kx, (uint16_t)ky);
          }
        }
        if (++cb == p.cblocks) {
          cb = 0;
          if (++kx == p.KW) { kx = 0; if (++ky == p.KH) { ky = 0; ++kz; } }
        }
        if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: leader CTA only =====================
    if (leader) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
      const uint32_t tail_k = (uint32_t)(((p.Cin & (kBlockK - 1)) + kUmmaK - 1) / kUmmaK);
      const uint64_t adesc0 = make_sw128_desc(sA), bdesc0 = make_sw128_desc(sB);
      const uint32_t a_step = a_stage_bytes >> 4, b_step = b_stage_bytes >> 4;
      uint32_t s = 0, ph = 0, tile_iter = 0;
      for (int t = pair_id; t < total_tiles; t += npairs, ++tile_iter) {
        const uint32_t buf = tile_iter & 1u;
        const uint32_t use = tile_iter >> 1;
        mbar_wait(bar_tmem_empty + 8 * buf, (use & 1u) ^ 1u, p.error_flag, 5);  // both CTAs drained this buffer
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * acc_cols;
        uint32_t cb = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(bar_full + 8 * s, ph, p.error_flag, 2);
          tc_fence_after();
          uint32_t ks = 4;
          if (tail_k) {
            if (++cb == (uint32_t)p.cblocks) { cb = 0; ks = tail_k; }
          }
          if (elect_one())
            mma_kblock_pair(acc, adesc0 + (uint64_t)(s * a_step), bdesc0 + (uint64_t)(s * b_step), (uint32_t)(kb != 0), ks,
                            idesc, bar_empty + 8 * s);
          if (++s == (uint32_t)S) { s = 0; ph ^= 1u; }
        }
        if (elect_one()) umma_commit_pair(bar_tmem_full + 8 * buf);
      }
    }
  } else {
    // ===================== epilogue warps (8 per CTA): this CTA's 128 rows x all block_n columns =====================
    const int wq = warp & 3;
    const int half = (warp - 2) >> 2;
    const int chunks = BN >> 4;
    const bool simple = (p.res == nullptr) && (p.raw == nullptr) && (p.out != nullptr);
    uint32_t tile_iter = 0;
    int loaded_n0 = -1;
    for (int t = pair_id; t < total_tiles; t += npairs, ++tile_iter) {
      const int n0 = (t % n_tiles_n) * BN;
      const int m0 = (t / n_tiles_n) * TILE_M + (int)rank * kBlockM;
      if (n0 != loaded_n0) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = threadIdx.x - 64; i < BN; i += 256) {
          const int c = n0 + i;
          const bool ok = c < p.Cout;
          const float bi = (ok && p.bias) ? p.bias[c] : 0.f;
          const float sc = (ok && p.scale) ? p.scale[c] : 1.f;
          const float sh = (ok && p.scale) ? p.shift[c] : 0.f;
          s_bias[i] = bi;
          s_scale[i] = sc;
          s_shift[i] = simple ? fmaf(bi, sc, sh) : sh;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        loaded_n0 = n0;
      }
      const uint32_t buf = tile_iter & 1u;
      const uint32_t use = tile_iter >> 1;
      mbar_wait(bar_tmem_full + 8 * buf, use & 1u, p.error_flag, 4);
      tc_fence_after();
      const int m = m0 + wq * 32 + lane;
      const bool row_ok = m < p.M;
      const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + buf * acc_cols;
      const int c_begin = half ? (chunks + 1) / 2 : 0;
      const int c_end = half ? chunks : (chunks + 1) / 2;
      if (simple && !p.epi_staged) {
        epilogue_simple_range(p, taddr, (long long)m, row_ok, n0, c_begin, c_end, s_scale, s_shift);
      } else {
        for (int c = c_begin; c < c_end; ++c)
          epilogue_chunk(p, taddr + (uint32_t)(c * 16), m, row_ok, n0 + c * 16, c * 16, s_bias, s_scale, s_shift, simple, 0u);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_on_cta(bar_tmem_empty + 8 * buf, 0u);
    }
  }

  tc_fence_before();
  cluster_sync_all();  // the peer's shared memory / TMEM must outlive every MMA that reads it
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Halo-resident 2-D convolution, stride 1 (3x3 / 4x4 / ... filters).
//   tile   = R output rows x OW columns of one image, enumerated on the padded-width grid
//            (position i = r*pw + c, pw = OW + KW - 1; columns c >= OW are discarded in the epilogue),
//            MT x 128 positions per tile
//   A      = the (R+KH-1) x pw input patch of a 64-channel block, fetched ONCE by a tiled TMA load with
//            zero fill for the border, 128-byte swizzled rows.  Tap (ky,kx) is the same patch read
//            through a descriptor whose start address is shifted by (ky*pw + kx) rows: the 128-byte
//            swizzle is a function of the absolute shared-memory address, so a row-shifted view is a
//            valid K-major operand (checked on hardware by tools/probe_umma_shift.cu)
//   B      = weights [Cout][tap][64] K-major, one TMA tile per (tap, channel block), own ring
//   rest   = as conv_umma_persistent_kernel: TMEM double buffer, 8 epilogue warps
template <int MT>
__global__ void __launch_bounds__(kHaloThreads, 1)
conv_halo_kernel(const HaloKernelParams p, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const int SA = p.a_stages, SB = p.b_resident ? 1 : p.b_stages, BN = p.block_n;
  const uint32_t b_stage_bytes = (uint32_t)BN * 128;
  const uint32_t b_region = p.b_resident ? (uint32_t)p.b_kblocks * b_stage_bytes : (uint32_t)SB * b_stage_bytes;
  const uint32_t sA = base;
  const uint32_t sB = sA + SA * p.a_stage_bytes;
  float* s_bias = reinterpret_cast<float*>(smem + (size_t)SA * p.a_stage_bytes + (size_t)b_region);
  float* s_scale = s_bias + 256;
  float* s_shift = s_scale + 256;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 256);
  const uint32_t bar_a_full = smem_u32(bars);
  const uint32_t bar_a_empty = bar_a_full + 8 * SA;
  const uint32_t bar_b_full = bar_a_empty + 8 * SA;
  const uint32_t bar_b_empty = bar_b_full + 8 * SB;
  const uint32_t bar_tmem_full = bar_b_empty + 8 * SB;
  const uint32_t bar_tmem_empty = bar_tmem_full + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 2 * SB + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = (p.Cout + BN - 1) / BN;
  const int total_tiles = p.NB * p.bands * n_tiles_n;
  const int taps = p.KH * p.KW;
  const uint32_t acc_cols = (uint32_t)(MT * BN);

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(bar_a_full + 8 * s, 1); mbar_init(bar_a_empty + 8 * s, 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(bar_b_full + 8 * s, 1); mbar_init(bar_b_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_tmem_full + 8 * b, 1); mbar_init(bar_tmem_empty + 8 * b, 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)p.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer: input patches (+ the resident weights, once) =====================
    {
      if (p.b_resident && elect_one()) {
        // b_full[0] doubles as the "weights resident" barrier: armed once, completes once
        mbar_arrive_expect_tx(bar_b_full, (uint32_t)p.b_kblocks * b_stage_bytes);
        for (int kbk = 0; kbk < p.b_kblocks; ++kbk)
          tma_load_2d(sB + kbk * b_stage_bytes, &tmB, bar_b_full, kbk * kBlockK, 0);
      }
      uint32_t ia = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int tb = t / n_tiles_n;
        const int band = tb % p.bands, n = tb / p.bands;
        const int y0 = band * p.R;
        for (int cb = 0; cb < p.cblocks; ++cb, ++ia) {
          const uint32_t s = ia % (uint32_t)SA, ph = (ia / (uint32_t)SA) & 1u;
          mbar_wait(bar_a_empty + 8 * s, ph ^ 1u, p.error_flag, 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(bar_a_full + 8 * s, p.a_tx_bytes);
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                ::"r"(sA + s * p.a_stage_bytes), "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(bar_a_full + 8 * s),
                "r"(cb * kBlockK), "r"(-p.pW), "r"(y0 - p.pH), "r"(n)
                : "memory");
          }
        }
      }
    }
  } else if (warp >= 10) {
    // ===================== TMA producers: weight tiles, two threads interleaved =====================
    const uint32_t me = (uint32_t)(warp - 10);
    if (!p.b_resident) {
      uint32_t ib = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const int n0 = (t % n_tiles_n) * BN;
        for (int cb = 0; cb < p.cblocks; ++cb) {
          for (int tap = 0; tap < taps; ++tap, ++ib) {
            if ((ib & 1u) != me) continue;
            const uint32_t s = ib % (uint32_t)SB, ph = (ib / (uint32_t)SB) & 1u;
            mbar_wait(bar_b_empty + 8 * s, ph ^ 1u, p.error_flag, 6);
            if (elect_one()) {
              mbar_arrive_expect_tx(bar_b_full + 8 * s, b_stage_bytes);
              tma_load_2d(sB + s * b_stage_bytes, &tmB, bar_b_full + 8 * s, (tap * p.cblocks + cb) * kBlockK, n0);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the rings, one elected lane issues) =====================
    {
      const uint32_t idesc = make_idesc(BN);
      const uint32_t rb = (uint32_t)p.row_bytes;
      const bool sw32 = p.row_bytes == 32;
      const int ksteps = p.row_bytes / 32;  // UMMA K=16 steps per tap: 4 (64 ch) or 1 (16-ch stem cells)
      uint32_t ia = 0, ib = 0, tile_iter = 0;
      if (p.b_resident) {
        mbar_wait(bar_b_full, 0, p.error_flag, 7);
        tc_fence_after();
      }
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tile_iter) {
        const uint32_t buf = tile_iter & 1u, use = tile_iter >> 1;
        mbar_wait(bar_tmem_empty + 8 * buf, (use & 1u) ^ 1u, p.error_flag, 5);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * acc_cols;
        uint32_t first = 1;
        for (int cb = 0; cb < p.cblocks; ++cb, ++ia) {
          const uint32_t sa = ia % (uint32_t)SA, pha = (ia / (uint32_t)SA) & 1u;
          mbar_wait(bar_a_full + 8 * sa, pha, p.error_flag, 2);
          tc_fence_after();
          const uint32_t patch = sA + sa * p.a_stage_bytes;
          int ky = 0, kx = 0;
          for (int tap = 0; tap < taps; ++tap) {
            // weight K position of this (tap, channel block): 64 ch -> one K block per (tap, cb);
            // stem cells -> 16 columns inside K block tap/4
            const int bk_elem = sw32 ? tap * 16 : (tap * p.cblocks + cb) * kBlockK;
            uint32_t btile;
            uint32_t sb = 0;
            if (p.b_resident) {
              btile = sB + (uint32_t)(bk_elem / kBlockK) * b_stage_bytes;
            } else {
              sb = ib % (uint32_t)SB;
              const uint32_t phb = (ib / (uint32_t)SB) & 1u;
              mbar_wait(bar_b_full + 8 * sb, phb, p.error_flag, 7);
              tc_fence_after();
              btile = sB + sb * b_stage_bytes;
              ++ib;
            }
            if (elect_one()) {
              const uint64_t bdesc = make_sw128_desc(btile) + (uint64_t)((bk_elem % kBlockK) * 2 / 16);
              const uint32_t a0 = patch + (uint32_t)(ky * p.pw + kx) * rb;  // tap = row shift of the patch
#pragma unroll
              for (int h = 0; h < MT; ++h) {
                const uint32_t ah = a0 + (uint32_t)h * (kBlockM * rb);
                const uint64_t adesc = sw32 ? make_sw32_desc(ah) : make_sw128_desc(ah);
                for (int k = 0; k < ksteps; ++k)
                  umma_bf16(acc + (uint32_t)(h * BN), adesc + 2 * k, bdesc + 2 * k, idesc, first ? (uint32_t)(k != 0) : 1u);
              }
              if (!p.b_resident) umma_commit(bar_b_empty + 8 * sb);
            }
            first = 0;
            if (++kx == p.KW) { kx = 0; ++ky; }
          }
          if (elect_one()) umma_commit(bar_a_empty + 8 * sa);
        }
        if (elect_one()) umma_commit(bar_tmem_full + 8 * buf);
      }
    }
  } else {
    // ===================== epilogue warps (8) =====================
    const int wq = warp & 3;
    const int half = (warp - 2) >> 2;
    const int chunks = BN >> 4;
    const bool simple = (p.res == nullptr) && (p.raw == nullptr) && (p.out != nullptr);
    EpiArgs e;
    e.out = p.out; e.out_cs = p.out_cs; e.out_coff = p.out_coff;
    e.raw = p.raw; e.raw_cs = p.raw_cs; e.raw_coff = p.raw_coff;
    e.res = p.res; e.res_cs = p.res_cs; e.res_coff = p.res_coff;
    e.Cout = p.Cout; e.relu = p.relu; e.simple = simple ? 1 : 0;
    const int pw = p.pw, R = p.R, OW = p.OW, OH = p.OH;
    const int c_begin = half ? (chunks + 1) / 2 : 0;
    const int c_end = half ? chunks : (chunks + 1) / 2;
    uint32_t tile_iter = 0;
    int loaded_n0 = -1;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++tile_iter) {
      const int n0 = (t % n_tiles_n) * BN;
      const int tb = t / n_tiles_n;
      const int band = tb % p.bands, n = tb / p.bands;
      const int y0 = band * R;
      if (n0 != loaded_n0) {
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = threadIdx.x - 64; i < BN; i += 256) {
          const int c = n0 + i;
          const bool ok = c < p.Cout;
          const float bi = (ok && p.bias) ? p.bias[c] : 0.f;
          const float sc = (ok && p.scale) ? p.scale[c] : 1.f;
          const float sh = (ok && p.scale) ? p.shift[c] : 0.f;
          s_bias[i] = bi; s_scale[i] = sc; s_shift[i] = simple ? fmaf(bi, sc, sh) : sh;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        loaded_n0 = n0;
      }
      const uint32_t buf = tile_iter & 1u, use = tile_iter >> 1;
      mbar_wait(bar_tmem_full + 8 * buf, use & 1u, p.error_flag, 4);
      tc_fence_after();
      epilogue_tile<MT>(e, tmem_base + ((uint32_t)(wq * 32) << 16) + buf * acc_cols, BN, n0, c_begin, c_end, s_bias,
                        s_scale, s_shift, [&](int h, long long& m, bool& ok) {
                          const int pos = h * kBlockM + wq * 32 + lane;  // position on the padded-width grid
                          const int r = pos / pw, c = pos - r * pw;
                          ok = (r < R) && (c < OW) && (y0 + r < OH);
                          m = ((long long)n * OH + y0 + r) * OW + c;
                        });
      tc_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_tmem_empty + 8 * buf) : "memory");
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Stem rows kernel (see StemRowsParams).  Replaces, for the first layer, conv_layer.cpp:28-43 +
// bn_layer.cu:12-124 + relu_layer.cu:10-27 and (pool = 1) pooling_layer.cu:13-61 MAX.
//   cell row Y of a frame = OW overlapping windows of 4 cells x 16 values (one 128-byte swizzle row per
//   output column); output row y needs cell rows y..y+3 against the weight K blocks 0..3.
//   warp 0: TMA producer (one cell row per stage), warp 1: MMA issuer -- every resident cell row is
//   multiplied into the (up to) four accumulators it contributes to, warps 2-9: epilogue.
constexpr int kStemThreads = 320;        // warp 0 TMA, 1 MMA, 2-9 epilogue
constexpr int kStemThreadsDirect = 512;  // + warps 10-15 (4..6 used): window gather from raw frames (src_mode 1 / 2)
constexpr int kStemRawStagesMax = 8;
constexpr int kStemSlots = 8;  // accumulator ring: 8 x 64 fp32 columns = all 512 TMEM columns

__device__ __forceinline__ uint32_t hmax2_u32(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

template <bool POOL, int SRC>
__global__ void __launch_bounds__(kStemThreadsDirect, 1)
stem_rows_kernel(const StemRowsParams p, const __grid_constant__ CUtensorMap tmX,
                 const __grid_constant__ CUtensorMap tmB) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw_addr);
  const int SA = p.a_stages;
  const uint32_t sA = base;
  const uint32_t sB = sA + (uint32_t)SA * 16384u;
  const uint32_t sRow = sB + 4u * 8192u;
  const size_t row_bytes = POOL ? 2 * 16384 : 0;
  const int RS = SRC ? p.raw_stages : 0;
  const size_t raw_off = (size_t)SA * 16384 + 4 * 8192 + row_bytes;
  const uint32_t sRaw = base + (uint32_t)raw_off;
  const size_t raw_bytes = (size_t)RS * p.raw_stage_bytes;
  float* s_scale = reinterpret_cast<float*>(smem + raw_off + raw_bytes);
  float* s_shift = s_scale + 64;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_shift + 64);
  const uint32_t bar_a_full = smem_u32(bars);
  const uint32_t bar_a_empty = bar_a_full + 8 * SA;
  const uint32_t bar_b_full = bar_a_empty + 8 * SA;
  const uint32_t bar_acc_full = bar_b_full + 8;
  const uint32_t bar_acc_empty = bar_acc_full + 8 * kStemSlots;
  const uint32_t bar_raw_full = bar_acc_empty + 8 * kStemSlots;
  const uint32_t bar_raw_empty = bar_raw_full + 8 * kStemRawStagesMax;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * SA + 1 + 2 * kStemSlots + 2 * kStemRawStagesMax);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int units = p.F * p.strips;

  if (threadIdx.x == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(bar_a_full + 8 * s, 1); mbar_init(bar_a_empty + 8 * s, 1); }
    mbar_init(bar_b_full, 1);
    for (int s = 0; s < kStemSlots; ++s) { mbar_init(bar_acc_full + 8 * s, 1); mbar_init(bar_acc_empty + 8 * s, 8); }
    for (int s = 0; s < RS; ++s) { mbar_init(bar_raw_full + 8 * s, 1); mbar_init(bar_raw_empty + 8 * s, 1); }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (threadIdx.x >= 64) {
    const int i = threadIdx.x - 64;
    if (i < 64) {
      const float bi = p.bias ? p.bias[i] : 0.f;
      const float sc = p.scale ? p.scale[i] : 1.f;
      const float sh = p.scale ? p.shift[i] : 0.f;
      s_scale[i] = sc;
      s_shift[i] = fmaf(bi, sc, sh);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const long long t_start = clock64();

  // rows [r0, r1) of frame f that work unit u produces
  auto unit_rows = [&](int u, int& f, int& r0, int& r1, int& p0) {
    f = u / p.strips;
    const int s = u - f * p.strips;
    if (POOL) {
      p0 = s * p.strip;
      const int p1 = min(p.PH, p0 + p.strip);
      r0 = 2 * p0;
      r1 = min(p.OH, 2 * p1 + 1);
    } else {
      p0 = 0;
      r0 = s * p.strip;
      r1 = min(p.OH, r0 + p.strip);
    }
  };

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the ring, one elected lane issues) =====================
    {
      if (elect_one()) {
        mbar_arrive_expect_tx(bar_b_full, 4u * 8192u);
        // weight K block i = vertical tap i; stacked in shared memory as tap 3, 2, 1, 0 (see the MMA issuer)
        for (int kb = 0; kb < 4; ++kb) tma_load_2d(sB + (3 - kb) * 8192u, &tmB, bar_b_full, kb * kBlockK, 0);
      }
      uint32_t s = 0, ph = 0;
      long long w_prod = 0;
      if constexpr (SRC != 0) {
        // raw frames: the two image rows (x 3 channels) a cell row is made of, bulk-copied into the raw ring
        const uint32_t esz = SRC == 1 ? 4u : 1u;
        const uint32_t rowb = (uint32_t)p.W * esz;
        const uint8_t* srcb = static_cast<const uint8_t*>(p.src);
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
          int f, r0, r1, p0;
          unit_rows(u, f, r0, r1, p0);
          for (int Y = r0; Y < r1 + 3; ++Y) {
            w_prod += mbar_wait_timed(bar_raw_empty + 8 * s, ph ^ 1u, p.error_flag, 1);
            if (elect_one()) {
              uint32_t nrows = 0;
#pragma unroll
              for (int dy = 0; dy < 2; ++dy) nrows += ((unsigned)(2 * Y - 3 + dy) < (unsigned)p.H) ? 3u : 0u;
              mbar_arrive_expect_tx(bar_raw_full + 8 * s, nrows * rowb);
#pragma unroll
              for (int dy = 0; dy < 2; ++dy) {
                const int y = 2 * Y - 3 + dy;
                if ((unsigned)y >= (unsigned)p.H) continue;
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                  const uint8_t* g = srcb + (((size_t)f * 3 + c) * p.H + y) * (size_t)rowb;
                  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                               ::"r"(sRaw + s * p.raw_stage_bytes + (uint32_t)(dy * 3 + c) * rowb), "l"(g), "r"(rowb),
                               "r"(bar_raw_full + 8 * s) : "memory");
                }
              }
            }
            if (++s == (uint32_t)RS) { s = 0; ph ^= 1u; }
          }
        }
      } else
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int f, r0, r1, p0;
        unit_rows(u, f, r0, r1, p0);
        for (int Y = r0; Y < r1 + 3; ++Y) {
          w_prod += mbar_wait_timed(bar_a_empty + 8 * s, ph ^ 1u, p.error_flag, 1);
          if (elect_one()) {
            if (p.debug_flags & 8) {
              asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a_full + 8 * s) : "memory");
            } else {
              mbar_arrive_expect_tx(bar_a_full + 8 * s, p.a_tx_bytes);
              asm volatile(
                  "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                  ::"r"(sA + s * 16384u), "l"(reinterpret_cast<uint64_t>(&tmX)), "r"(bar_a_full + 8 * s), "r"(0), "r"(0),
                  "r"(Y), "r"(f)
                  : "memory");
            }
          }
          if (++s == (uint32_t)SA) { s = 0; ph ^= 1u; }
        }
      }
      if ((p.debug_flags & 16) && blockIdx.x == 0 && lane == 0)
        printf("stem_rows cta0: producer waited %lld cycles of %lld\n", w_prod, clock64() - t_start);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the rings, one elected lane issues) =====================
    {
      const uint32_t idesc = make_idesc(64);
      mbar_wait(bar_b_full, 0, p.error_flag, 7);
      tc_fence_after();
      const uint64_t adesc0 = make_sw128_desc(sA), bdesc0 = make_sw128_desc(sB);
      uint32_t s = 0, ph = 0, ir = 0;
      long long w_full = 0, w_acc = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int f, r0, r1, p0;
        unit_rows(u, f, r0, r1, p0);
        for (int Y = r0; Y < r1 + 3; ++Y) {
          w_full += mbar_wait_timed(bar_a_full + 8 * s, ph, p.error_flag, 2);
          if (Y < r1) {  // a new output row starts with this cell row: its accumulator slot must be drained
            const uint32_t ridx = ir + (uint32_t)(Y - r0);
            w_acc += mbar_wait_timed(bar_acc_empty + 8 * (ridx % (uint32_t)kStemSlots),
                                     ((ridx / (uint32_t)kStemSlots) & 1u) ^ 1u, p.error_flag, 5);
          }
          if (SRC) fence_proxy_async_smem();  // windows were written with st.shared (generic proxy)
          tc_fence_after();
          if (elect_one()) {
            // taps ilo..ihi of this cell row land in output rows Y-ilo .. Y-ihi of the unit
            const uint64_t adesc = adesc0 + (uint64_t)(s * (16384u >> 4));
            const int ilo = max(0, Y - r1 + 1), ihi = min(3, Y - r0);
            if (!(p.debug_flags & 4)) {
              if (ilo == 0) {
                // newest row: first contribution, overwrites its accumulator (N = 64)
                const uint32_t slot = (ir + (uint32_t)(Y - r0)) % (uint32_t)kStemSlots;
                const uint64_t bdesc = bdesc0 + (uint64_t)(3 * (8192 >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tmem_base + slot * 64u, adesc + 2 * k, bdesc + 2 * k, idesc, k ? 1u : 0u);
              }
              // taps >= 1: consecutive older rows = adjacent accumulator slots = ONE N = 64 x count MMA per K step
              // against the weight tiles stacked tap 3, 2, 1 (split only where the slot ring wraps)
              int tap = ihi;
              int cnt = ihi - max(1, ilo) + 1;
              while (cnt > 0) {
                const uint32_t slot = (ir + (uint32_t)(Y - tap - r0)) % (uint32_t)kStemSlots;
                const int c1 = min(cnt, kStemSlots - (int)slot);
                const uint64_t bdesc = bdesc0 + (uint64_t)((3 - tap) * (8192 >> 4));
                const uint32_t idesc_n = make_idesc(64 * c1);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_bf16(tmem_base + slot * 64u, adesc + 2 * k, bdesc + 2 * k, idesc_n, 1u);
                tap -= c1;
                cnt -= c1;
              }
            }
            if (ihi == 3 && ilo <= 3)  // row Y-3 just received its last tap
              umma_commit(bar_acc_full + 8 * ((ir + (uint32_t)(Y - 3 - r0)) % (uint32_t)kStemSlots));
            umma_commit(bar_a_empty + 8 * s);
          }
          if (++s == (uint32_t)SA) { s = 0; ph ^= 1u; }
        }
        ir += (uint32_t)(r1 - r0);
      }
      if ((p.debug_flags & 16) && blockIdx.x == 0 && lane == 0)
        printf("stem_rows cta0: mma waited %lld (tile loads) + %lld (accumulator slots) cycles of %lld\n", w_full, w_acc,
               clock64() - t_start);
    }
  } else if (warp >= 10) {
    // ===================== window gather (src_mode 1 / 2): warp gw builds every 4th cell row =====================
    // cell (Y, X) = 2x2 pixels x 3 channels (+4 zeros) of the zero-padded frame; window x of the tile = cells
    // x..x+3 = one 128-byte row in the 128B-swizzled K-major layout the MMA descriptors expect
    if constexpr (SRC != 0) {
      const int gw = warp - 10;  // blockDim = 320 + 32 * gather_warps
      const int CW = p.OW + 3;
      const float mean[3] = {p.mean0, p.mean1, p.mean2};
      uint32_t i = 0;
      for (int u = blockIdx.x; u < units; u += gridDim.x) {
        int f, r0, r1, p0;
        unit_rows(u, f, r0, r1, p0);
        for (int Y = r0; Y < r1 + 3; ++Y, ++i) {
          if ((int)(i % (uint32_t)p.gather_warps) != gw) continue;
          const uint32_t rs = i % (uint32_t)RS, rph = (i / (uint32_t)RS) & 1u;
          const uint32_t as = i % (uint32_t)SA, aph = (i / (uint32_t)SA) & 1u;
          mbar_wait(bar_raw_full + 8 * rs, rph, p.error_flag, 8);
          mbar_wait(bar_a_empty + 8 * as, aph ^ 1u, p.error_flag, 9);
          const uint32_t raw = sRaw + rs * p.raw_stage_bytes;
          const uint32_t tile = sA + as * 16384u;
          const bool yok0 = (unsigned)(2 * Y - 3) < (unsigned)p.H, yok1 = (unsigned)(2 * Y - 2) < (unsigned)p.H;
          for (int X = lane; X < CW; X += 32) {
            float v[12];
            // pixel columns 2X-3, 2X-2 of image rows 2Y-3, 2Y-2: one base address per cell, the six (row, channel) planes
            // at uniform offsets from it, the second column at +1 element
            const int xa = 2 * X - 3;
            const bool okx0 = (unsigned)xa < (unsigned)p.W, okx1 = (unsigned)(xa + 1) < (unsigned)p.W;
            const uint32_t a0 = raw + (uint32_t)xa * (SRC == 1 ? 4u : 1u);  // (wraps for xa < 0: never dereferenced then)
            const uint32_t planeb = (uint32_t)p.W * (SRC == 1 ? 4u : 1u);
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
              const bool yok = dy ? yok1 : yok0;
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const uint32_t a = a0 + (uint32_t)(dy * 3 + c) * planeb;
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                  float t = 0.f;
                  if (yok && (dx ? okx1 : okx0)) {
                    if (SRC == 1) {
                      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(t) : "r"(a + (uint32_t)dx * 4u));
                    } else {
                      uint32_t b;
                      asm volatile("ld.shared.u8 %0, [%1];" : "=r"(b) : "r"(a + (uint32_t)dx));
                      t = (float)b - mean[c];
                    }
                  }
                  v[(dy * 2 + dx) * 3 + c] = t;
                }
              }
            }
            uint32_t w[8];
#pragma unroll
            for (int j = 0; j < 6; ++j) w[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
            w[6] = 0u; w[7] = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int x = X - j;  // window that sees this cell as its j-th
              if (x >= 0 && x < p.OW) {
                const uint32_t rowa = tile + (uint32_t)x * 128u;
                const uint32_t sw = (uint32_t)(x & 7);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((((uint32_t)(2 * j)) ^ sw) << 4)), "r"(w[0]),
                             "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowa + ((((uint32_t)(2 * j + 1)) ^ sw) << 4)),
                             "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
              }
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a_full + 8 * as) : "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_raw_empty + 8 * rs) : "memory");
          }
        }
      }
    }
  } else {
    // ===================== epilogue warps (8): lane quarter x column half =====================
    const int wq = warp & 3;
    const int half = (warp - 2) >> 2;
    const int x = wq * 32 + lane;  // output column of this thread
    const bool x_ok = x < p.OW;
    const int te = threadIdx.x - 64;
    const float* sc = s_scale + half * 32;
    const float* sh = s_shift + half * 32;
    uint32_t ir = 0, em = 0;
    long long w_epi = 0, w_bar = 0;
    uint32_t cm[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) cm[j] = 0u;
    bool pend = false;
    uint32_t pend_buf = 0;
    long long pend_row = 0;
    // second half of a pooled row: thread te owns channel group fg of pooled columns fq0, fq0 + 32, ...  Column q reads
    // conv columns 2q, 2q+1, 2q+2 of the row buffer; (2q) & 7 is the same for every q of a thread, so the three swizzled
    // offsets are per-thread constants and a step of 32 columns is +8192 bytes.
    const int fg = te & 7, fq0 = te >> 3;
    const uint32_t fsw = (uint32_t)(2 * (fq0 & 3));
    const uint32_t foff0 = (uint32_t)(2 * fq0) * 128u + (((uint32_t)fg ^ fsw) << 4);
    const uint32_t foff1 = (uint32_t)(2 * fq0 + 1) * 128u + (((uint32_t)fg ^ (fsw + 1u)) << 4);
    const uint32_t foff2 = (uint32_t)(2 * fq0 + 2) * 128u + (((uint32_t)fg ^ ((fsw + 2u) & 7u)) << 4);
    const uint32_t neg_inf2 = 0xff80ff80u;  // bf16x2 (-inf, -inf): the neutral element of the maximum
    auto lds128_if = [&](uint32_t addr, bool ok) {
      uint4 t;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %5, 0;\n\t"
          "mov.b32 %0, %6;\n\tmov.b32 %1, %6;\n\tmov.b32 %2, %6;\n\tmov.b32 %3, %6;\n\t"
          "@p ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n\t}"
          : "=&r"(t.x), "=&r"(t.y), "=&r"(t.z), "=&r"(t.w)
          : "r"(addr), "r"((uint32_t)ok), "r"(neg_inf2));
      return t;
    };
    auto emit_flush_fn = [&]() {
      if (!pend) return;
      pend = false;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      uint32_t rb = pend_buf;
      __nv_bfloat16* o = p.out + (pend_row * p.PW + fq0) * p.out_cs + p.out_coff + fg * 8;
      const long long ostep = 32 * p.out_cs;
      for (int q = fq0; q < p.PW; q += 32, rb += 8192u, o += ostep) {
        uint4 m;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w) : "r"(rb + foff0));
        const uint4 t1 = lds128_if(rb + foff1, 2 * q + 1 < p.OW);
        const uint4 t2 = lds128_if(rb + foff2, 2 * q + 2 < p.OW);
        m.x = hmax2_u32(hmax2_u32(m.x, t1.x), t2.x); m.y = hmax2_u32(hmax2_u32(m.y, t1.y), t2.y);
        m.z = hmax2_u32(hmax2_u32(m.z, t1.z), t2.z); m.w = hmax2_u32(hmax2_u32(m.w, t1.w), t2.w);
        if (p.relu) { m.x = hmax2_u32(m.x, 0u); m.y = hmax2_u32(m.y, 0u); m.z = hmax2_u32(m.z, 0u); m.w = hmax2_u32(m.w, 0u); }
        *reinterpret_cast<uint4*>(o) = m;
      }
    };
    for (int u = blockIdx.x; u < units; u += gridDim.x) {
      int f, r0, r1, p0;
      unit_rows(u, f, r0, r1, p0);
      // pooled row `pr` of this frame from the per-thread column maxima `w` (POOL only), in two halves: emit_store puts
      // the thread's column into the row buffer, emit_flush (barrier + horizontal 3-max + ReLU + global store) runs one conv
      // row later, right after that row's TMEM load has been issued -- the barrier skew and the shared/global latency of
      // the second half hide behind the load instead of sitting on the per-row chain.  ReLU commutes with both maxima, so
      // it is applied once per pooled element here instead of once per conv element.
      auto emit_store = [&](int pr, const uint32_t (&w)[16]) {
        emit_flush_fn();
        const uint32_t rowbuf = sRow + (em & 1u) * 16384u;
        ++em;
        if (x_ok) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t chunk = (uint32_t)(half * 4 + c) ^ (uint32_t)(x & 7);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rowbuf + (uint32_t)x * 128u + (chunk << 4)),
                         "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3]) : "memory");
          }
        }
        pend = true; pend_buf = rowbuf; pend_row = (long long)f * p.PH + pr;
      };
      // one accumulator row: TMEM -> 32 fp32 columns of this thread's output pixel; the slot goes back to the MMA issuer
      // as soon as the load has landed.  (No alternative data path in here: a second definition of v[] makes ptxas copy
      // all 32 registers per row -- the r02 source-level profile showed 64 MOVs per row from exactly that.)
      auto load_row = [&](int y, uint32_t (&v)[32]) {
        const uint32_t ridx = ir + (uint32_t)(y - r0);
        const uint32_t slot = ridx % (uint32_t)kStemSlots;
        w_epi += mbar_wait_timed(bar_acc_full + 8 * slot, (ridx / (uint32_t)kStemSlots) & 1u, p.error_flag, 4);
        tc_fence_after();
        tmem_ld32_nowait(tmem_base + ((uint32_t)(wq * 32) << 16) + slot * 64u + (uint32_t)(half * 32), v);
        if (POOL) emit_flush_fn();  // second half of the previous pooled row, under the TMEM load's latency
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_acc_empty + 8 * slot) : "memory");
      };
      if (!POOL) {
        for (int y = r0; y < r1; ++y) {
          uint32_t v[32];
          load_row(y, v);
          float fv[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = reinterpret_cast<const float4*>(sc)[j], b = reinterpret_cast<const float4*>(sh)[j];
            fv[4 * j + 0] = fmaf(__uint_as_float(v[4 * j + 0]), a.x, b.x);
            fv[4 * j + 1] = fmaf(__uint_as_float(v[4 * j + 1]), a.y, b.y);
            fv[4 * j + 2] = fmaf(__uint_as_float(v[4 * j + 2]), a.z, b.z);
            fv[4 * j + 3] = fmaf(__uint_as_float(v[4 * j + 3]), a.w, b.w);
          }
          if (p.relu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) fv[j] = fmaxf(fv[j], 0.f);
          }
          if (x_ok) {
            __nv_bfloat16* dst = p.out + (((long long)f * p.OH + y) * p.OW + x) * p.out_cs + p.out_coff + half * 32;
            store16_bf16(dst, *reinterpret_cast<const float (*)[16]>(&fv[0]), 16);
            store16_bf16(dst + 16, *reinterpret_cast<const float (*)[16]>(&fv[16]), 16);
          }
        }
      } else {
        // scale/shift (conv bias + BN folded) and pack to bf16 pairs; ReLU waits until after the pooling maxima
        auto bn_pack = [&](const uint32_t (&v)[32], uint32_t (&w)[16]) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 a = reinterpret_cast<const float4*>(sc)[j], b = reinterpret_cast<const float4*>(sh)[j];
            w[2 * j] = pack_bf16x2(fmaf(__uint_as_float(v[4 * j + 0]), a.x, b.x), fmaf(__uint_as_float(v[4 * j + 1]), a.y, b.y));
            w[2 * j + 1] = pack_bf16x2(fmaf(__uint_as_float(v[4 * j + 2]), a.z, b.z), fmaf(__uint_as_float(v[4 * j + 3]), a.w, b.w));
          }
        };
        // rows in pairs (even, odd index inside the unit): pooled row k = max(rows 2k, 2k+1, 2k+2).  `cm` carries
        // max(2k, 2k+1) into the even row that completes the window; the even row's own values live in `we` until the
        // odd row has been folded in -- no register array is ever copied.
        for (int y = r0; y < r1; y += 2) {
          uint32_t we[16];
          {
            uint32_t v[32];
            load_row(y, v);
            bn_pack(v, we);
          }
          if (y > r0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) cm[j] = hmax2_u32(cm[j], we[j]);
            emit_store(p0 + ((y - r0) >> 1) - 1, cm);
          }
          if (y + 1 < r1) {
            uint32_t wo[16];
            {
              uint32_t v[32];
              load_row(y + 1, v);
              bn_pack(v, wo);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) cm[j] = hmax2_u32(we[j], wo[j]);
          }
        }
      }
      if (POOL && ((r1 - r0) & 1) == 0 && r1 > r0) emit_store(p0 + ((r1 - r0) >> 1) - 1, cm);  // clipped last window
      ir += (uint32_t)(r1 - r0);
    }
    emit_flush_fn();
    if ((p.debug_flags & 16) && blockIdx.x == 0 && lane == 0 && (warp == 2 || warp == 9))
      printf("stem_rows cta0 warp %d: epilogue waited %lld (accumulators) + %lld (row barrier) cycles of %lld\n", warp, w_epi,
             w_bar, clock64() - t_start);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

cudaError_t conv_umma_configure() {
  cudaError_t e = cudaFuncSetAttribute(conv_umma_persistent_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_umma_persistent_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_umma_persistent_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_umma_persistent_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_umma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(conv_halo_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e != cudaSuccess) return e;
#define ECO_STEM_ATTR(P, S)                                                                                         \
  e = cudaFuncSetAttribute(stem_rows_kernel<P, S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); \
  if (e != cudaSuccess) return e;
  ECO_STEM_ATTR(true, 0) ECO_STEM_ATTR(true, 1) ECO_STEM_ATTR(true, 2)
  ECO_STEM_ATTR(false, 0) ECO_STEM_ATTR(false, 1) ECO_STEM_ATTR(false, 2)
#undef ECO_STEM_ATTR
  return cudaFuncSetAttribute(conv_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t launch_stem_rows(const StemRowsParams& p, const CUtensorMap& tmX, const CUtensorMap& tmB, cudaStream_t stream) {
  const int units = p.F * p.strips;
  const int grid = units < p.num_sms ? units : p.num_sms;
  const size_t smem = stem_rows_smem_bytes(p);
  const int threads = p.src_mode ? kStemThreads + 32 * p.gather_warps : kStemThreads;
  if (p.pool) {
    if (p.src_mode == 0) stem_rows_kernel<true, 0><<<grid, threads, smem, stream>>>(p, tmX, tmB);
    else if (p.src_mode == 1) stem_rows_kernel<true, 1><<<grid, threads, smem, stream>>>(p, tmX, tmB);
    else stem_rows_kernel<true, 2><<<grid, threads, smem, stream>>>(p, tmX, tmB);
  } else {
    if (p.src_mode == 0) stem_rows_kernel<false, 0><<<grid, threads, smem, stream>>>(p, tmX, tmB);
    else if (p.src_mode == 1) stem_rows_kernel<false, 1><<<grid, threads, smem, stream>>>(p, tmX, tmB);
    else stem_rows_kernel<false, 2><<<grid, threads, smem, stream>>>(p, tmX, tmB);
  }
  return cudaGetLastError();
}

cudaError_t launch_conv_halo(const HaloKernelParams& p, int m_halves, const CUtensorMap& tmX, const CUtensorMap& tmB,
                             cudaStream_t stream) {
  const int tiles = p.NB * p.bands * ((p.Cout + p.block_n - 1) / p.block_n);
  const int grid = tiles < p.num_sms ? tiles : p.num_sms;
  const size_t smem = halo_smem_bytes(p);
  if (m_halves == 2) conv_halo_kernel<2><<<grid, kHaloThreads, smem, stream>>>(p, tmX, tmB);
  else conv_halo_kernel<1><<<grid, kHaloThreads, smem, stream>>>(p, tmX, tmB);
  return cudaGetLastError();
}

cudaError_t launch_conv_pair(const ConvKernelParams& p, const CUtensorMap& tmA, const CUtensorMap& tmBhalf,
                             cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)p.stages * (kBlockM * 128 + (size_t)(p.block_n / 2) * 128) + 3 * 256 * sizeof(float) + 64 * 8;
  const int tiles = ((p.M + 2 * kBlockM - 1) / (2 * kBlockM)) * ((p.Cout + p.block_n - 1) / p.block_n);
  const int pairs = tiles < p.num_sms / 2 ? tiles : p.num_sms / 2;
  conv_umma_pair_kernel<<<2 * pairs, kPersistThreads, smem, stream>>>(p, tmA, tmBhalf);  // __cluster_dims__(2,1,1)
  return cudaGetLastError();
}

cudaError_t launch_conv_umma(const ConvKernelParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB,
                             cudaStream_t stream) {
  const size_t smem = conv_smem_bytes(p.block_n, p.stages, p.persistent ? p.m_halves : 1,
                                      (p.persistent && p.epi_staged) ? conv_epi_stage_bytes(p.epi_group) : 0);
  if (p.persistent) {
    const int tile_m = kBlockM * p.m_halves;
    const int tiles = ((p.M + tile_m - 1) / tile_m) * ((p.Cout + p.block_n - 1) / p.block_n);
    const int grid = tiles < p.num_sms ? tiles : p.num_sms;
    if (p.multicast) {
      // clusters of two CTAs (launch attribute): tmB must be the half-tile map (box rows = block_n / 2)
      const int supers = ((((p.M + tile_m - 1) / tile_m) + 1) / 2) * ((p.Cout + p.block_n - 1) / p.block_n);
      const int clusters = supers < p.num_sms / 2 ? supers : p.num_sms / 2;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * clusters, 1, 1);
      cfg.blockDim = dim3(kPersistThreads, 1, 1);
      cfg.dynamicSmemBytes = smem;
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      if (p.m_halves == 2) return cudaLaunchKernelEx(&cfg, conv_umma_persistent_kernel<2, true>, p, tmA, tmB);
      return cudaLaunchKernelEx(&cfg, conv_umma_persistent_kernel<1, true>, p, tmA, tmB);
    }
    if (p.m_halves == 2)
      conv_umma_persistent_kernel<2, false><<<grid, kPersistThreads, smem, stream>>>(p, tmA, tmB);
    else
      conv_umma_persistent_kernel<1, false><<<grid, kPersistThreads, smem, stream>>>(p, tmA, tmB);
    return cudaGetLastError();
  }
  dim3 grid((p.M + kBlockM - 1) / kBlockM, (p.Cout + p.block_n - 1) / p.block_n, 1);
  conv_umma_kernel<<<grid, kConvThreads, smem, stream>>>(p, tmA, tmB);
  return cudaGetLastError();
}

}  // namespace eco
