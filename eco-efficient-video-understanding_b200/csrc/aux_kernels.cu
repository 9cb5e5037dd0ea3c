// aux_kernels.cu -- HBM-bound helper kernels (sm_100a).  All of these are streaming byte movers:
// one pass over the tensor, 16-byte vector accesses along the channel axis, grid sized from the
// element count.  Caffe semantics cited per kernel.
#include "aux_kernels.cuh"

#include <cfloat>
#include <stdint.h>

namespace eco {
namespace {

constexpr int kThreads = 256;
inline int blocks_for(long long n) { return (int)((n + kThreads - 1) / kThreads); }

// ---------------------------------------------------------------- layout transforms
__global__ void f32_to_cl_kernel(const float* __restrict__ src, ClView d) {
  // one thread per (o, i, c); c fastest so the bf16 writes coalesce
  const long long total = d.outer * d.inner * d.C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % d.C);
    const long long oi = t / d.C;
    const long long i = oi % d.inner, o = oi / d.inner;
    const float v = src[(o * d.C + c) * d.inner + i];
    __nv_bfloat16* q = d.ptr + oi * d.ps() + d.coff + c;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    q[0] = hi;
    if (d.seg) {
      q[d.seg] = __float2bfloat16_rn(v - __bfloat162float(hi));
      q[2 * d.seg] = hi;
    }
  }
}
__global__ void cl_to_f32_kernel(ClView s, float* __restrict__ dst) {
  // one thread per (o, c, i); i fastest so the fp32 writes coalesce
  const long long total = s.outer * s.inner * s.C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long i = t % s.inner;
    const long long oc = t / s.inner;
    const int c = (int)(oc % s.C);
    const long long o = oc / s.C;
    const __nv_bfloat16* q = s.ptr + (o * s.inner + i) * s.ps() + s.coff + c;
    dst[t] = __bfloat162float(q[0]) + (s.seg ? __bfloat162float(q[s.seg]) : 0.f);
  }
}

__global__ void stem_s2d_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int F, int H, int W,
                                int CH, int CW) {
  const long long total = (long long)F * CH * CW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(t % CW);
    const int Y = (int)((t / CW) % CH);
    const long long f = t / ((long long)CW * CH);
    const float* img = src + f * 3LL * H * W;
    __align__(16) __nv_bfloat16 cell[16];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = 2 * Y + dy - 3, x = 2 * X + dx - 3;
        const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          cell[(dy * 2 + dx) * 3 + c] = __float2bfloat16_rn(ok ? img[((long long)c * H + y) * W + x] : 0.f);
      }
#pragma unroll
    for (int j = 12; j < 16; ++j) cell[j] = __float2bfloat16_rn(0.f);
    uint4* o = reinterpret_cast<uint4*>(dst + t * 16);
    o[0] = reinterpret_cast<const uint4*>(cell)[0];
    o[1] = reinterpret_cast<const uint4*>(cell)[1];
  }
}

__global__ void stem_s2d_u8_kernel(const unsigned char* __restrict__ src, __nv_bfloat16* __restrict__ dst, int F, int H,
                                   int W, int CH, int CW, float m0, float m1, float m2) {
  const long long total = (long long)F * CH * CW;
  const float mean[3] = {m0, m1, m2};
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int X = (int)(t % CW);
    const int Y = (int)((t / CW) % CH);
    const long long f = t / ((long long)CW * CH);
    const unsigned char* img = src + f * 3LL * H * W;
    __align__(16) __nv_bfloat16 cell[16];
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = 2 * Y + dy - 3, x = 2 * X + dx - 3;
        const bool ok = (unsigned)y < (unsigned)H && (unsigned)x < (unsigned)W;
#pragma unroll
        for (int c = 0; c < 3; ++c)
          cell[(dy * 2 + dx) * 3 + c] =
              __float2bfloat16_rn(ok ? (float)img[((long long)c * H + y) * W + x] - mean[c] : 0.f);
      }
#pragma unroll
    for (int j = 12; j < 16; ++j) cell[j] = __float2bfloat16_rn(0.f);
    uint4* o = reinterpret_cast<uint4*>(dst + t * 16);
    o[0] = reinterpret_cast<const uint4*>(cell)[0];
    o[1] = reinterpret_cast<const uint4*>(cell)[1];
  }
}
__global__ void u8_to_f32_mean_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, long long outer, int C,
                                      long long inner, float m0, float m1, float m2, float m3) {
  const float mean[4] = {m0, m1, m2, m3};
  const long long total = outer * C * inner;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)((t / inner) % C);
    dst[t] = (float)src[t] - mean[c & 3];
  }
}

// ---------------------------------------------------------------- pooling, channels-last
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    f[2 * j] = __low2float(t);
    f[2 * j + 1] = __high2float(t);
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __nv_bfloat162 t;
  t = __floats2bfloat162_rn(f[0], f[1]); v.x = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[2], f[3]); v.y = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[4], f[5]); v.z = *reinterpret_cast<uint32_t*>(&t);
  t = __floats2bfloat162_rn(f[6], f[7]); v.w = *reinterpret_cast<uint32_t*>(&t);
  return v;
}

// MAX: window [o*s-p, min(start+k, in)) then start=max(start,0), init -FLT_MAX  (pooling_layer.cpp:199-224)
// AVE: divisor = prod(min(start+k, in+p) - start) before clipping to the image      (pooling_layer.cpp:247-262)
// One thread = one output pixel x 8 channels (16 B); 32-bit index math (the host checks the range).
__global__ void __launch_bounds__(256) pool_cl_kernel(const PoolParams p) {
  const int cg = p.C >> 3;
  const unsigned total = (unsigned)p.NB * p.OD * p.OH * p.OW * cg;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    const unsigned pix = t / (unsigned)cg;
    const int g = (int)(t - pix * cg);
    unsigned r = pix;
    const int ox = (int)(r % (unsigned)p.OW); r /= (unsigned)p.OW;
    const int oy = (int)(r % (unsigned)p.OH); r /= (unsigned)p.OH;
    const int oz = (int)(r % (unsigned)p.OD);
    const int n = (int)(r / (unsigned)p.OD);
    int z0 = oz * p.sD - p.pD, y0 = oy * p.sH - p.pH, x0 = ox * p.sW - p.pW;
    int z1, y1, x1;
    float div = 1.f;
    if (p.is_max) {
      z1 = min(z0 + p.KD, p.ID); y1 = min(y0 + p.KH, p.IH); x1 = min(x0 + p.KW, p.IW);
    } else {
      z1 = min(z0 + p.KD, p.ID + p.pD); y1 = min(y0 + p.KH, p.IH + p.pH); x1 = min(x0 + p.KW, p.IW + p.pW);
      div = (float)((z1 - z0) * (y1 - y0) * (x1 - x0));
      z1 = min(z1, p.ID); y1 = min(y1, p.IH); x1 = min(x1, p.IW);
    }
    z0 = max(z0, 0); y0 = max(y0, 0); x0 = max(x0, 0);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = p.is_max ? -FLT_MAX : 0.f;
    const __nv_bfloat16* base = p.x + p.x_coff + g * 8;
    for (int z = z0; z < z1; ++z)
      for (int y = y0; y < y1; ++y) {
        const long long rowpix = ((long long)(n * p.ID + z) * p.IH + y) * p.IW;
        for (int x = x0; x < x1; ++x) {
          const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + (rowpix + x) * p.x_cs));
          float f[8];
          unpack8(v, f);
          if (p.is_max) {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], f[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
          }
        }
      }
    if (!p.is_max) {
      // true division like pooling_layer.cpp:262 (x * (1/d) is not bit-identical in general)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = acc[j] / div;
    }
    *reinterpret_cast<uint4*>(p.y + (long long)pix * p.y_cs + p.y_coff + g * 8) = pack8(acc);
  }
}

// 2-D 3x3 pooling (every pool of ECO's 2-D trunk): one thread produces a strip of T outputs along x
// for 8 channels, sliding a 3-column window so each input column is loaded once per output row
// (3*S+small loads per output instead of 9).  Same caffe semantics as pool_cl_kernel.
template <bool IS_MAX, int S, int T>
__global__ void __launch_bounds__(256) pool2d_k3_strip_kernel(const PoolParams p) {
  constexpr int NCOLS = (T - 1) * S + 3;
  const int cg = p.C >> 3;
  const int strips = (p.OW + T - 1) / T;
  const unsigned total = (unsigned)p.NB * p.OH * strips * cg;
  for (unsigned t = blockIdx.x * blockDim.x + threadIdx.x; t < total; t += gridDim.x * blockDim.x) {
    unsigned r = t / (unsigned)cg;
    const int g = (int)(t - r * cg);
    const int xs = (int)(r % (unsigned)strips); r /= (unsigned)strips;
    const int oy = (int)(r % (unsigned)p.OH);
    const int n = (int)(r / (unsigned)p.OH);
    const int ox0 = xs * T;
    const int iy0 = oy * S - p.pH;
    const int ix0 = ox0 * S - p.pW;
    const __nv_bfloat16* base = p.x + p.x_coff + g * 8 + (long long)n * p.IH * p.IW * p.x_cs;
    bool rok[3];
#pragma unroll
    for (int rr = 0; rr < 3; ++rr) rok[rr] = (unsigned)(iy0 + rr) < (unsigned)p.IH;
    const int hcount = min(iy0 + 3, p.IH + p.pH) - iy0;  // pad-inclusive rows of the window (AVE divisor)
    float c0[8], c1[8], c2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) c0[j] = c1[j] = c2[j] = IS_MAX ? -FLT_MAX : 0.f;
#pragma unroll
    for (int col = 0; col < NCOLS; ++col) {
      const int ix = ix0 + col;
      float cv[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) cv[j] = IS_MAX ? -FLT_MAX : 0.f;
      if ((unsigned)ix < (unsigned)p.IW) {
#pragma unroll
        for (int rr = 0; rr < 3; ++rr) {
          if (rok[rr]) {
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + ((long long)(iy0 + rr) * p.IW + ix) * p.x_cs));
            float f[8];
            unpack8(v, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) cv[j] = IS_MAX ? fmaxf(cv[j], f[j]) : cv[j] + f[j];
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) { c0[j] = c1[j]; c1[j] = c2[j]; c2[j] = cv[j]; }
      if (col >= 2 && (col - 2) % S == 0) {
        const int ox = ox0 + (col - 2) / S;
        if (ox < p.OW) {
          float o[8];
          if (IS_MAX) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaxf(fmaxf(c0[j], c1[j]), c2[j]);
          } else {
            const int ws = ox * S - p.pW;
            const float div = (float)(hcount * (min(ws + 3, p.IW + p.pW) - ws));
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = ((c0[j] + c1[j]) + c2[j]) / div;
          }
          const long long opix = ((long long)n * p.OH + oy) * p.OW + ox;
          *reinterpret_cast<uint4*>(p.y + opix * p.y_cs + p.y_coff + g * 8) = pack8(o);
        }
      }
    }
  }
}

// 2-D pooling with the input rows staged through shared memory by bulk async copies
// (cp.async.bulk -> mbarrier): one CTA = one band of output rows of one image.  The rows it needs are
// contiguous in a dense channels-last map, so a handful of 10-20 KB bulk copies put >= 64 KB per SM in
// flight (three CTAs per SM), which plain 16-byte loads could not; every input row is fetched once per
// band.  Same caffe semantics as pool_cl_kernel (pooling_layer.cpp:199-262).
template <int SS>  // 3x3 window stride known at compile time (1 / 2: the strip loops unroll fully), 0: generic
__global__ void __launch_bounds__(256) pool2d_rows_kernel(const PoolParams p, int rows_out) {
  extern __shared__ __align__(128) uint8_t pool_smem[];
  __shared__ __align__(8) uint64_t bar;
  const int n = blockIdx.y;
  const int oy0 = blockIdx.x * rows_out;
  const int oy1 = min(oy0 + rows_out, p.OH);
  const int iy_lo = oy0 * p.sH - p.pH;
  const int iy_hi = (oy1 - 1) * p.sH - p.pH + p.KH;
  const int vy_lo = max(iy_lo, 0), vy_hi = min(iy_hi, p.IH);
  const uint32_t row_bytes = (uint32_t)p.IW * p.C * 2;
  const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t smem_addr = (uint32_t)__cvta_generic_to_shared(pool_smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t total = (uint32_t)(vy_hi - vy_lo) * row_bytes;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"(total) : "memory");
    const __nv_bfloat16* src = p.x + ((long long)n * p.IH + vy_lo) * p.IW * p.C;
    for (int r = 0; r < vy_hi - vy_lo; ++r) {
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr + (uint32_t)r * row_bytes),
                   "l"(src + (long long)r * p.IW * p.C), "r"(row_bytes), "r"(bar_addr)
                   : "memory");
    }
  }
  // one thread polls the mbarrier; everybody else parks at the CTA barrier (256 spinning threads per CTA took
  // a large share of the SM's issue slots from the CTAs that were already computing: ncu, r01 final)
  if (threadIdx.x == 0) {
    uint32_t ok = 0;
    const long long t0 = clock64();
    while (!ok) {
      asm volatile("{ .reg .pred q; mbarrier.try_wait.parity.shared::cta.b64 q, [%1], 0; selp.u32 %0,1,0,q; }"
                   : "=r"(ok) : "r"(bar_addr) : "memory");
      if (!ok && clock64() - t0 > 4000000000LL) __trap();
    }
  }
  __syncthreads();
  const int cg = p.C >> 3;
  const unsigned long long magic_cg = (1ULL << 32) / (unsigned)cg + 1ULL;  // i / cg == (i * magic) >> 32 for i < 2^16
  if (SS != 0) {
    // 3x3 windows: a thread owns a strip of T outputs along x and slides over the input columns once
    constexpr int T = 4;
    constexpr int S = SS ? SS : 1;
    const int strips = (p.OW + T - 1) / T;
    const int items = (oy1 - oy0) * strips * cg;
    for (int i = threadIdx.x; i < items; i += blockDim.x) {
      const int r1 = (int)(((unsigned long long)(unsigned)i * magic_cg) >> 32);
      const int g = i - r1 * cg;
      const int rowi = r1 / strips;
      const int xs = r1 - rowi * strips;
      const int oy = oy0 + rowi;
      const int ox0 = xs * T;
      const int iy0 = oy * S - p.pH;
      const int ix0 = ox0 * S - p.pW;
      const uint8_t* rows[3];
      bool rok[3];
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const int y = iy0 + rr;
        rok[rr] = y >= vy_lo && y < vy_hi;
        rows[rr] = pool_smem + (size_t)(rok[rr] ? y - vy_lo : 0) * row_bytes + (size_t)g * 16;
      }
      const long long opix0 = ((long long)n * p.OH + oy) * p.OW;
      constexpr int ncols = (T - 1) * S + 3;
      if (p.is_max) {
        // max of bf16 values is exact in packed bf16 arithmetic: no fp32 unpacking
        const __nv_bfloat162 ninf = __float2bfloat162_rn(-INFINITY);
        __nv_bfloat162 c0[4], c1[4], c2[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) c0[j] = c1[j] = c2[j] = ninf;
#pragma unroll
        for (int col = 0; col < ncols; ++col) {
          const int ix = ix0 + col;
          __nv_bfloat162 cv[4] = {ninf, ninf, ninf, ninf};
          if ((unsigned)ix < (unsigned)p.IW) {
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
              if (rok[rr]) {
                const uint4 v = *reinterpret_cast<const uint4*>(rows[rr] + (size_t)ix * p.C * 2);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) cv[j] = __hmax2(cv[j], h[j]);
              }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) { c0[j] = c1[j]; c1[j] = c2[j]; c2[j] = cv[j]; }
          const int d = col - 2;
          if (d >= 0 && (S == 1 || (d & 1) == 0)) {
            const int ox = ox0 + (S == 1 ? d : d >> 1);
            if (ox < p.OW) {
              uint4 o;
              __nv_bfloat162* oh = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
              for (int j = 0; j < 4; ++j) oh[j] = __hmax2(__hmax2(c0[j], c1[j]), c2[j]);
              *reinterpret_cast<uint4*>(p.y + (opix0 + ox) * p.y_cs + p.y_coff + g * 8) = o;
            }
          }
        }
      } else {
        const int hcount = min(iy0 + 3, p.IH + p.pH) - iy0;  // pad-inclusive window rows (AVE divisor)
        const bool affine = p.bias != nullptr || p.scale != nullptr;
        float e_b[8], e_s[8], e_h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          e_b[j] = p.bias ? __ldg(p.bias + g * 8 + j) : 0.f;
          e_s[j] = p.scale ? __ldg(p.scale + g * 8 + j) : 1.f;
          e_h[j] = p.scale ? __ldg(p.shift + g * 8 + j) : 0.f;
        }
        float c0[8], c1[8], c2[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) c0[j] = c1[j] = c2[j] = 0.f;
#pragma unroll
        for (int col = 0; col < ncols; ++col) {
          const int ix = ix0 + col;
          float cv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) cv[j] = 0.f;
          if ((unsigned)ix < (unsigned)p.IW) {
#pragma unroll
            for (int rr = 0; rr < 3; ++rr)
              if (rok[rr]) {
                const uint4 v = *reinterpret_cast<const uint4*>(rows[rr] + (size_t)ix * p.C * 2);
                float f[8];
                unpack8(v, f);
#pragma unroll
                for (int j = 0; j < 8; ++j) cv[j] += f[j];
              }
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) { c0[j] = c1[j]; c1[j] = c2[j]; c2[j] = cv[j]; }
          const int d = col - 2;
          if (d >= 0 && (S == 1 || (d & 1) == 0)) {
            const int ox = ox0 + (S == 1 ? d : d >> 1);
            if (ox < p.OW) {
              const int ws = ox * S - p.pW;
              const float inv = __frcp_rn((float)(hcount * (min(ws + 3, p.IW + p.pW) - ws)));
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = ((c0[j] + c1[j]) + c2[j]) * inv;
              if (affine) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j] + e_b[j], e_s[j], e_h[j]);
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = fmaxf(o[j], 0.f);
              }
              *reinterpret_cast<uint4*>(p.y + (opix0 + ox) * p.y_cs + p.y_coff + g * 8) = pack8(o);
            }
          }
        }
      }
    }
    return;
  }
  const int items = (oy1 - oy0) * p.OW * cg;
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int r1 = (int)(((unsigned long long)(unsigned)i * magic_cg) >> 32);
    const int g = i - r1 * cg;
    const int ox = r1 % p.OW;
    const int oy = oy0 + r1 / p.OW;
    int y0 = oy * p.sH - p.pH, x0 = ox * p.sW - p.pW;
    int y1, x1;
    float div = 1.f;
    if (p.is_max) {
      y1 = min(y0 + p.KH, p.IH); x1 = min(x0 + p.KW, p.IW);
    } else {
      y1 = min(y0 + p.KH, p.IH + p.pH); x1 = min(x0 + p.KW, p.IW + p.pW);
      div = (float)((y1 - y0) * (x1 - x0));
      y1 = min(y1, p.IH); x1 = min(x1, p.IW);
    }
    y0 = max(y0, 0); x0 = max(x0, 0);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = p.is_max ? -FLT_MAX : 0.f;
    for (int y = y0; y < y1; ++y) {
      const uint8_t* row = pool_smem + (size_t)(y - vy_lo) * row_bytes + (size_t)g * 16;
      for (int x = x0; x < x1; ++x) {
        const uint4 v = *reinterpret_cast<const uint4*>(row + (size_t)x * p.C * 2);
        float f[8];
        unpack8(v, f);
        if (p.is_max) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], f[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
      }
    }
    if (!p.is_max) {
      const float inv = __frcp_rn(div);  // same rounding as the strip path (div is a small integer)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = acc[j] * inv;
      if (p.bias || p.scale) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float b = p.bias ? __ldg(p.bias + g * 8 + j) : 0.f;
          const float sc = p.scale ? __ldg(p.scale + g * 8 + j) : 1.f;
          const float sh = p.scale ? __ldg(p.shift + g * 8 + j) : 0.f;
          acc[j] = fmaf(acc[j] + b, sc, sh);
        }
      }
      if (p.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.f);
      }
    }
    const long long opix = ((long long)n * p.OH + oy) * p.OW + ox;
    *reinterpret_cast<uint4*>(p.y + opix * p.y_cs + p.y_coff + g * 8) = pack8(acc);
  }
}

// global average: one thread per (outer, channel); consecutive threads read consecutive channels
__global__ void global_avg_cl_kernel(ClView s, float* __restrict__ dst) {
  const long long total = s.outer * s.C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % s.C);
    const long long o = t / s.C;
    const long long ps = s.ps();
    const __nv_bfloat16* px = s.ptr + o * s.inner * ps + s.coff + c;
    float acc = 0.f;
    for (long long i = 0; i < s.inner; ++i)
      acc += __bfloat162float(px[i * ps]) + (s.seg ? __bfloat162float(px[i * ps + s.seg]) : 0.f);
    dst[t] = acc / (float)s.inner;
  }
}

// split-precision pooling: one thread per (output position, channel); caffe's window rules (pooling_layer.cpp:199-262)
__global__ void pool_cl_split_kernel(const PoolParams p, long long xseg, long long yseg) {
  const long long total = (long long)p.NB * p.OD * p.OH * p.OW * p.C;
  const long long xps = 3 * p.x_cs, yps = 3 * p.y_cs;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % p.C);
    long long r = t / p.C;
    const long long opix = r;
    const int ox = (int)(r % p.OW); r /= p.OW;
    const int oy = (int)(r % p.OH); r /= p.OH;
    const int oz = (int)(r % p.OD);
    const long long n = r / p.OD;
    int z0 = oz * p.sD - p.pD, y0 = oy * p.sH - p.pH, x0 = ox * p.sW - p.pW;
    int z1, y1, x1;
    float div = 1.f;
    if (p.is_max) {
      z1 = min(z0 + p.KD, p.ID); y1 = min(y0 + p.KH, p.IH); x1 = min(x0 + p.KW, p.IW);
    } else {
      z1 = min(z0 + p.KD, p.ID + p.pD); y1 = min(y0 + p.KH, p.IH + p.pH); x1 = min(x0 + p.KW, p.IW + p.pW);
      div = (float)((z1 - z0) * (y1 - y0) * (x1 - x0));
      z1 = min(z1, p.ID); y1 = min(y1, p.IH); x1 = min(x1, p.IW);
    }
    z0 = max(z0, 0); y0 = max(y0, 0); x0 = max(x0, 0);
    float acc = p.is_max ? -FLT_MAX : 0.f;
    for (int z = z0; z < z1; ++z)
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          const __nv_bfloat16* q = p.x + (((n * p.ID + z) * p.IH + y) * p.IW + x) * xps + p.x_coff + c;
          const float v = __bfloat162float(q[0]) + __bfloat162float(q[xseg]);
          acc = p.is_max ? fmaxf(acc, v) : acc + v;
        }
    if (!p.is_max) acc = acc / div;
    __nv_bfloat16* o = p.y + opix * yps + p.y_coff + c;
    const __nv_bfloat16 hi = __float2bfloat16_rn(acc);
    o[0] = hi;
    o[yseg] = __float2bfloat16_rn(acc - __bfloat162float(hi));
    o[2 * yseg] = hi;
  }
}

__global__ void pool_f32_kernel(const PoolF32Params p) {
  const long long total = (long long)p.NC * p.OD * p.OH * p.OW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int ox = (int)(r % p.OW); r /= p.OW;
    const int oy = (int)(r % p.OH); r /= p.OH;
    const int oz = (int)(r % p.OD);
    const long long nc = r / p.OD;
    int z0 = oz * p.sD - p.pD, y0 = oy * p.sH - p.pH, x0 = ox * p.sW - p.pW;
    int z1, y1, x1;
    float div = 1.f;
    if (p.is_max) {
      z1 = min(z0 + p.KD, p.ID); y1 = min(y0 + p.KH, p.IH); x1 = min(x0 + p.KW, p.IW);
    } else {
      z1 = min(z0 + p.KD, p.ID + p.pD); y1 = min(y0 + p.KH, p.IH + p.pH); x1 = min(x0 + p.KW, p.IW + p.pW);
      div = (float)((z1 - z0) * (y1 - y0) * (x1 - x0));
      z1 = min(z1, p.ID); y1 = min(y1, p.IH); x1 = min(x1, p.IW);
    }
    z0 = max(z0, 0); y0 = max(y0, 0); x0 = max(x0, 0);
    const float* px = p.x + nc * (long long)p.ID * p.IH * p.IW;
    float acc = p.is_max ? -FLT_MAX : 0.f;
    for (int z = z0; z < z1; ++z)
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) {
          const float v = px[((long long)z * p.IH + y) * p.IW + x];
          acc = p.is_max ? fmaxf(acc, v) : acc + v;
        }
    p.y[t] = p.is_max ? acc : acc / div;
  }
}

// ---------------------------------------------------------------- inner product (one warp per output)
__global__ void inner_product_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                     const float* __restrict__ b, float* __restrict__ y, int M, int N, int K) {
  const long long warp_id = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_id >= (long long)M * N) return;
  const int n = (int)(warp_id % N);
  const long long m = warp_id / N;
  const float* xp = x + m * K;
  const float* wp = w + (long long)n * K;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(xp[k], wp[k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) y[warp_id] = acc + (b ? b[n] : 0.f);
}

__global__ void scale_shift_relu_cl_kernel(ClView x, ClView y, const float* __restrict__ scale,
                                           const float* __restrict__ shift, int relu) {
  const long long total = x.outer * x.inner * x.C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % x.C);
    const long long pix = t / x.C;
    float v = __bfloat162float(x.ptr[pix * x.cs + x.coff + c]);
    v = fmaf(v, scale[c], shift[c]);
    if (relu) v = fmaxf(v, 0.f);
    y.ptr[pix * y.cs + y.coff + c] = __float2bfloat16_rn(v);
  }
}
__global__ void eltwise_sum_cl_kernel(ClView a, ClView b, ClView y) {
  const long long total = a.outer * a.inner * a.C;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(t % a.C);
    const long long pix = t / a.C;
    const float v = __bfloat162float(a.ptr[pix * a.cs + a.coff + c]) + __bfloat162float(b.ptr[pix * b.cs + b.coff + c]);
    y.ptr[pix * y.cs + y.coff + c] = __float2bfloat16_rn(v);
  }
}

__global__ void softmax_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int M, int N) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const float* xp = x + (long long)m * N;
  float* yp = y + (long long)m * N;
  float mx = xp[0];
  for (int n = 1; n < N; ++n) mx = fmaxf(mx, xp[n]);
  float s = 0.f;
  for (int n = 0; n < N; ++n) { const float e = expf(xp[n] - mx); yp[n] = e; s += e; }
  for (int n = 0; n < N; ++n) yp[n] /= s;
}

inline int grid_cap(long long n) {
  long long b = (n + kThreads - 1) / kThreads;
  const long long cap = 148LL * 32;  // grid-stride above this
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

cudaError_t launch_f32_to_cl(const float* src, ClView dst, cudaStream_t st) {
  const long long n = dst.outer * dst.inner * dst.C;
  if (n == 0) return cudaSuccess;
  f32_to_cl_kernel<<<grid_cap(n), kThreads, 0, st>>>(src, dst);
  return cudaGetLastError();
}
cudaError_t launch_cl_to_f32(ClView src, float* dst, cudaStream_t st) {
  const long long n = src.outer * src.inner * src.C;
  if (n == 0) return cudaSuccess;
  cl_to_f32_kernel<<<grid_cap(n), kThreads, 0, st>>>(src, dst);
  return cudaGetLastError();
}
cudaError_t launch_stem_s2d(const float* src, __nv_bfloat16* dst, int F, int H, int W, int CH, int CW,
                            cudaStream_t st) {
  const long long n = (long long)F * CH * CW;
  if (n == 0) return cudaSuccess;
  stem_s2d_kernel<<<grid_cap(n), kThreads, 0, st>>>(src, dst, F, H, W, CH, CW);
  return cudaGetLastError();
}
cudaError_t launch_stem_s2d_u8(const unsigned char* src, __nv_bfloat16* dst, int F, int H, int W, int CH, int CW,
                               float mean0, float mean1, float mean2, cudaStream_t st) {
  const long long n = (long long)F * CH * CW;
  if (n == 0) return cudaSuccess;
  stem_s2d_u8_kernel<<<grid_cap(n), kThreads, 0, st>>>(src, dst, F, H, W, CH, CW, mean0, mean1, mean2);
  return cudaGetLastError();
}
cudaError_t launch_u8_to_f32_mean(const unsigned char* src, float* dst, long long outer, int C, long long inner,
                                  float mean0, float mean1, float mean2, float mean3, cudaStream_t st) {
  const long long n = outer * C * inner;
  if (n == 0) return cudaSuccess;
  u8_to_f32_mean_kernel<<<grid_cap(n), kThreads, 0, st>>>(src, dst, outer, C, inner, mean0, mean1, mean2, mean3);
  return cudaGetLastError();
}
// per-device function attributes (called once per device by Net::ensure_device)
cudaError_t aux_kernels_configure() {
  cudaError_t e = cudaFuncSetAttribute(pool2d_rows_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(pool2d_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(pool2d_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
}

cudaError_t launch_pool_cl(const PoolParams& p, cudaStream_t st) {
  const long long n = (long long)p.NB * p.OD * p.OH * p.OW * (p.C / 8);
  if (n == 0) return cudaSuccess;
  if (n >= (1LL << 31)) return cudaErrorInvalidValue;  // 32-bit index math in the kernel
  const bool epilogue = p.bias || p.scale || p.relu;
  if (p.ID == 1 && p.KD == 1 && p.OD == 1 && p.x_cs == p.C && p.x_coff == 0) {
    // rows staged through shared memory by bulk copies
    const size_t row_bytes = (size_t)p.IW * p.C * 2;
    int rows_out = 1;
    while (rows_out < p.OH && ((size_t)(rows_out * p.sH + p.KH) * row_bytes) <= 72 * 1024) ++rows_out;
    const size_t rows_in = (size_t)(rows_out - 1) * p.sH + p.KH;
    const size_t smem = rows_in * row_bytes;
    if (smem <= 200 * 1024 && row_bytes % 16 == 0 && p.NB <= 65535) {
      const bool k3 = p.KH == 3 && p.KW == 3 && p.sH == p.sW && (p.sW == 1 || p.sW == 2);
      if (epilogue && (p.is_max || !k3)) return cudaErrorNotSupported;
      dim3 grid((p.OH + rows_out - 1) / rows_out, p.NB, 1);
      if (k3 && p.sW == 1) pool2d_rows_kernel<1><<<grid, 256, smem, st>>>(p, rows_out);
      else if (k3) pool2d_rows_kernel<2><<<grid, 256, smem, st>>>(p, rows_out);
      else pool2d_rows_kernel<0><<<grid, 256, smem, st>>>(p, rows_out);
      return cudaGetLastError();
    }
  }
  if (epilogue) return cudaErrorNotSupported;  // only the row-staged AVE 3x3 path applies it
  if (p.ID == 1 && p.KD == 1 && p.KH == 3 && p.KW == 3 && p.sH == p.sW && (p.sH == 1 || p.sH == 2) && p.pH == p.pW &&
      p.pH <= 1) {
    constexpr int T = 4;
    const long long nt = (long long)p.NB * p.OH * ((p.OW + T - 1) / T) * (p.C / 8);
    const long long b = (nt + kThreads - 1) / kThreads;
    const unsigned grid = (unsigned)(b > 148LL * 64 ? 148LL * 64 : b);
    if (p.is_max && p.sH == 2) pool2d_k3_strip_kernel<true, 2, T><<<grid, kThreads, 0, st>>>(p);
    else if (p.is_max) pool2d_k3_strip_kernel<true, 1, T><<<grid, kThreads, 0, st>>>(p);
    else if (p.sH == 2) pool2d_k3_strip_kernel<false, 2, T><<<grid, kThreads, 0, st>>>(p);
    else pool2d_k3_strip_kernel<false, 1, T><<<grid, kThreads, 0, st>>>(p);
    return cudaGetLastError();
  }
  const long long blocks = (n + kThreads - 1) / kThreads;
  pool_cl_kernel<<<(unsigned)(blocks > 148LL * 64 ? 148LL * 64 : blocks), kThreads, 0, st>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_pool_cl_split(const PoolParams& p, long long x_seg, long long y_seg, cudaStream_t st) {
  const long long n = (long long)p.NB * p.OD * p.OH * p.OW * p.C;
  if (n == 0) return cudaSuccess;
  if (p.bias || p.scale || p.relu) return cudaErrorNotSupported;
  pool_cl_split_kernel<<<grid_cap(n), kThreads, 0, st>>>(p, x_seg, y_seg);
  return cudaGetLastError();
}
cudaError_t launch_global_avg_cl(ClView src, float* dst, cudaStream_t st) {
  const long long n = src.outer * src.C;
  if (n == 0) return cudaSuccess;
  global_avg_cl_kernel<<<grid_cap(n), kThreads, 0, st>>>(src, dst);
  return cudaGetLastError();
}
cudaError_t launch_pool_f32(const PoolF32Params& p, cudaStream_t st) {
  const long long n = (long long)p.NC * p.OD * p.OH * p.OW;
  if (n == 0) return cudaSuccess;
  pool_f32_kernel<<<grid_cap(n), kThreads, 0, st>>>(p);
  return cudaGetLastError();
}
cudaError_t launch_inner_product(const float* x, const float* w, const float* b, float* y, int M, int N, int K,
                                 cudaStream_t st) {
  const long long warps = (long long)M * N;
  if (warps == 0) return cudaSuccess;
  inner_product_kernel<<<blocks_for(warps * 32), kThreads, 0, st>>>(x, w, b, y, M, N, K);
  return cudaGetLastError();
}
cudaError_t launch_scale_shift_relu_cl(ClView x, ClView y, const float* scale, const float* shift, int relu,
                                       cudaStream_t st) {
  const long long n = x.outer * x.inner * x.C;
  if (n == 0) return cudaSuccess;
  scale_shift_relu_cl_kernel<<<grid_cap(n), kThreads, 0, st>>>(x, y, scale, shift, relu);
  return cudaGetLastError();
}
cudaError_t launch_eltwise_sum_cl(ClView a, ClView b, ClView y, cudaStream_t st) {
  const long long n = a.outer * a.inner * a.C;
  if (n == 0) return cudaSuccess;
  eltwise_sum_cl_kernel<<<grid_cap(n), kThreads, 0, st>>>(a, b, y);
  return cudaGetLastError();
}
cudaError_t launch_softmax_f32(const float* x, float* y, int M, int N, cudaStream_t st) {
  if (M == 0) return cudaSuccess;
  softmax_f32_kernel<<<(M + 127) / 128, 128, 0, st>>>(x, y, M, N);
  return cudaGetLastError();
}

}  // namespace eco
