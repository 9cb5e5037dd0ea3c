// train_kernels.cu -- HBM-bound kernels of the training path (sm_100a).  One pass per tensor, 16-byte accesses along
// the channel axis of the channels-last maps, per-channel reductions as block partials + fp32 atomics.
// Reference semantics, by kernel:
//   BN TRAIN forward / backward   caffe_3d/src/caffe/layers/bn_layer.cpp:107-207, :241-335
//   pooling backward              pooling_layer.cpp:280-377 (max_idx_ = first maximum, :206-222)
//   dropout                       dropout_layer.cpp:33-75
//   inner product backward        inner_product_layer.cpp:96-120
//   softmax loss / accuracy       softmax_loss_layer.cpp:48-120, accuracy_layer.cpp:50-92
//   solver update                 solver.cpp:637-797 (ClipGradients, Normalize, Regularize, ComputeUpdateValue), :820-860
#include "train_kernels.cuh"

#include <cfloat>

namespace eco {
namespace {

constexpr int kT = 256;

__device__ __forceinline__ void up8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w[j]);
    f[2 * j] = __low2float(t);
    f[2 * j + 1] = __high2float(t);
  }
}
__device__ __forceinline__ uint4 pk8(const float (&f)[8]) {
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 t = __floats2bfloat162_rn(f[2 * j], f[2 * j + 1]);
    w[j] = *reinterpret_cast<uint32_t*>(&t);
  }
  return make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ uint4 ld8(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }

inline int grid_for(long long items, int cap_mult = 8) {
  long long b = (items + kT - 1) / kT;
  const long long cap = 148LL * cap_mult;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------
// Per-channel reductions.  Thread t of a block owns channel group g = t % G (8 channels, one 16-byte load per row)
// and walks rows r = first + t / G, + rows_per_block ...; partials are combined across the block's row lanes in
// shared memory and added to the output with one atomic per channel and block.
//   MODE 0: S1 = sum x                MODE 1: S1 = sum (x - mean)^2
//   MODE 2: g = dy * (y > 0 if relu); S1 = sum g, S2 = sum g * (x - mean) * inv_std
template <int MODE>
__global__ void colreduce_cl_kernel(ClView x, ClView y, ClView dy, const float* __restrict__ mean,
                                    const float* __restrict__ inv_std, const float* __restrict__ slope,
                                    const float* __restrict__ bias, int relu, float* __restrict__ out) {
  extern __shared__ float red[];  // [lanes][G*8] (x2 for MODE 2)
  const int G = x.C / 8;
  const long long rows = x.outer * x.inner;
  const int lanes = blockDim.x / G;  // row lanes per block (>= 1: launcher guarantees G <= blockDim.x)
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  float s1[8], s2[8], mu[8], is[8], sl[8], bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { s1[j] = 0.f; s2[j] = 0.f; mu[j] = 0.f; is[j] = 1.f; sl[j] = 1.f; bi[j] = 0.f; }
  if (MODE >= 1 && lane < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mu[j] = mean[g * 8 + j];
      if (MODE == 2) { is[j] = inv_std[g * 8 + j]; sl[j] = slope[g * 8 + j]; bi[j] = bias[g * 8 + j]; }
    }
  }
  if (lane < lanes) {
    // four rows in flight per thread (the loads are issued before the first of them is consumed): 296 CTAs x 256 threads x
    // 4 x 16 B keeps ~5 MB outstanding, what HBM3e needs at ~0.8 us latency; the summation order per thread is unchanged
    auto consume = [&](const uint4& xraw, const uint4& graw) {
      float xv[8];
      up8(xraw, xv);
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) s1[j] += xv[j];
      } else if (MODE == 1) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = xv[j] - mu[j]; s1[j] = fmaf(d, d, s1[j]); }
      } else {
        float gv[8];
        up8(graw, gv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float xn = (xv[j] - mu[j]) * is[j];
          // ReLU mask recomputed from x exactly as the forward kernel formed the pre-activation (bf16 keeps fp32's exponent
          // range, so y > 0 <=> pre-activation > 0): the stored output is not read again
          if (relu && !(xn * sl[j] + bi[j] > 0.f)) gv[j] = 0.f;
          s1[j] += gv[j];
          s2[j] = fmaf(gv[j], xn, s2[j]);
        }
      }
    };
    const long long stride = (long long)gridDim.x * lanes;
    long long r = (long long)blockIdx.x * lanes + lane;
    const __nv_bfloat16* xp = x.ptr + x.coff + g * 8;
    const __nv_bfloat16* gp = MODE == 2 ? dy.ptr + dy.coff + g * 8 : nullptr;
    for (; r + 3 * stride < rows; r += 4 * stride) {
      uint4 xr[4], gr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        xr[u] = ld8(xp + (r + u * stride) * x.cs);
        gr[u] = MODE == 2 ? ld8(gp + (r + u * stride) * dy.cs) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) consume(xr[u], gr[u]);
    }
    for (; r < rows; r += stride) consume(ld8(xp + r * x.cs), MODE == 2 ? ld8(gp + r * dy.cs) : make_uint4(0, 0, 0, 0));
  }
  const int C = x.C;
  float* r1 = red;
  float* r2 = red + (size_t)lanes * C;
  if (lane < lanes) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      r1[(size_t)lane * C + g * 8 + j] = s1[j];
      if (MODE == 2) r2[(size_t)lane * C + g * 8 + j] = s2[j];
    }
  }
  __syncthreads();
  // per-block partials go to `partial[block][C (x2)]`; colreduce_finish_kernel adds them in block order, so the result does
  // not depend on the order in which blocks retire (run-to-run reproducible statistics and gradients)
  float* dst = out + (size_t)blockIdx.x * C * (MODE == 2 ? 2 : 1);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int l = 0; l < lanes; ++l) {
      a += r1[(size_t)l * C + c];
      if (MODE == 2) b += r2[(size_t)l * C + c];
    }
    dst[c] = a;
    if (MODE == 2) dst[C + c] = b;
  }
}
// 32 channels x 8 slices per CTA: slice s adds the partials of blocks s, s + 8, ... in order, then the 8 slice sums are added
// in slice order -- a fixed summation tree, so the result is independent of scheduling (and 8 loads are in flight per channel)
__global__ void colreduce_finish_kernel(const float* __restrict__ partial, int blocks, int n, float* __restrict__ out) {
  __shared__ float part[8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int sl = threadIdx.y;
  float a = 0.f;
  if (c < n)
    for (int b = sl; b < blocks; b += 8) a += partial[(size_t)b * n + c];
  part[sl][threadIdx.x] = a;
  __syncthreads();
  if (sl == 0 && c < n) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += part[q][threadIdx.x];
    out[c] += t;
  }
}

template <int MODE>
cudaError_t launch_colreduce(ClView x, ClView y, ClView dy, const float* mean, const float* inv_std, const float* slope,
                             const float* bias, int relu, float* out, float* scratch, cudaStream_t st) {
  const long long rows = x.outer * x.inner;
  if (rows == 0 || x.C == 0) return cudaSuccess;
  if (x.C % 8 != 0) return cudaErrorInvalidValue;
  const int G = x.C / 8;
  int threads = 256;
  while (threads < G) threads *= 2;
  if (threads > 1024) return cudaErrorInvalidValue;  // C <= 8192
  const int lanes = threads / G;
  const size_t smem = (size_t)lanes * x.C * sizeof(float) * (MODE == 2 ? 2 : 1);
  long long blocks = (rows + (long long)lanes * 8 - 1) / ((long long)lanes * 8);  // >= 8 rows per lane
  const int cap = MODE == 2 ? kColReduceMaxBlocks / 2 : kColReduceMaxBlocks;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (x.C > kColReduceMaxC || !scratch) return cudaErrorInvalidValue;
  colreduce_cl_kernel<MODE><<<(unsigned)blocks, threads, smem, st>>>(x, y, dy, mean, inv_std, slope, bias, relu, scratch);
  const int n = x.C * (MODE == 2 ? 2 : 1);
  colreduce_finish_kernel<<<(n + 31) / 32, dim3(32, 8, 1), 0, st>>>(scratch, (int)blocks, n, out);
  return cudaGetLastError();
}

// ---- single-pass batch statistics (Welford per thread, Chan's pairwise combination in a fixed order) ----
// One read of the tensor gives mean and the sum of squared deviations without the E[x^2] - mean^2 cancellation.
__device__ __forceinline__ void chan_combine(float& nA, float& meanA, float& m2A, float nB, float meanB, float m2B) {
  if (nB == 0.f) return;
  if (nA == 0.f) { nA = nB; meanA = meanB; m2A = m2B; return; }
  const float n = nA + nB, d = meanB - meanA;
  meanA += d * (nB / n);
  m2A += m2B + d * d * (nA * nB / n);
  nA = n;
}
__global__ void bn_welford_cl_kernel(ClView x, float* __restrict__ partial /* [blocks][2C] mean, M2 */, float* __restrict__ counts) {
  extern __shared__ float red[];  // [lanes][2C] + [lanes] counts
  const int G = x.C / 8, C = x.C;
  const long long rows = x.outer * x.inner;
  const int lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  float mean[8], m2[8], n = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) { mean[j] = 0.f; m2[j] = 0.f; }
  if (lane < lanes) {
    auto consume = [&](const uint4& raw) {
      float v[8];
      up8(raw, v);
      n += 1.f;
      const float inv = 1.f / n;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[j] - mean[j];
        mean[j] += d * inv;
        m2[j] = fmaf(d, v[j] - mean[j], m2[j]);
      }
    };
    const long long stride = (long long)gridDim.x * lanes;
    long long r = (long long)blockIdx.x * lanes + lane;
    const __nv_bfloat16* xp = x.ptr + x.coff + g * 8;
    for (; r + 3 * stride < rows; r += 4 * stride) {   // four rows in flight, consumed in row order
      uint4 xr[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xr[u] = ld8(xp + (r + u * stride) * x.cs);
#pragma unroll
      for (int u = 0; u < 4; ++u) consume(xr[u]);
    }
    for (; r < rows; r += stride) consume(ld8(xp + r * x.cs));
    float* rl = red + (size_t)lane * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) { rl[g * 8 + j] = mean[j]; rl[C + g * 8 + j] = m2[j]; }
    if (g == 0) red[(size_t)lanes * 2 * C + lane] = n;
  }
  __syncthreads();
  const float* cnt = red + (size_t)lanes * 2 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float nA = 0.f, mA = 0.f, qA = 0.f;
    for (int l = 0; l < lanes; ++l) chan_combine(nA, mA, qA, cnt[l], red[(size_t)l * 2 * C + c], red[(size_t)l * 2 * C + C + c]);
    partial[(size_t)blockIdx.x * 2 * C + c] = mA;
    partial[(size_t)blockIdx.x * 2 * C + C + c] = qA;
    if (c == 0) counts[blockIdx.x] = nA;
  }
}
// combine the block partials (8 slices in parallel, each in block order, then the slices in order) and finish the layer's
// statistics: mean, biased variance, inverse std, running averages (bn_layer.cpp:107-157)
__global__ void bn_welford_finish_kernel(const float* __restrict__ partial, const float* __restrict__ counts, int blocks, int C,
                                         float* __restrict__ mean, float* __restrict__ inv_std, float* __restrict__ batch_var,
                                         float* __restrict__ run_mean, float* __restrict__ run_var, float momentum, float eps) {
  __shared__ float sn[8][32], sm[8][32], sq[8][32];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int sl = threadIdx.y;
  float nA = 0.f, mA = 0.f, qA = 0.f;
  if (c < C)
    for (int b = sl; b < blocks; b += 8) chan_combine(nA, mA, qA, counts[b], partial[(size_t)b * 2 * C + c], partial[(size_t)b * 2 * C + C + c]);
  sn[sl][threadIdx.x] = nA; sm[sl][threadIdx.x] = mA; sq[sl][threadIdx.x] = qA;
  __syncthreads();
  if (sl == 0 && c < C) {
    float n = 0.f, m = 0.f, q = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) chan_combine(n, m, q, sn[k][threadIdx.x], sm[k][threadIdx.x], sq[k][threadIdx.x]);
    const float var = q / n;  // biased (bn_layer.cpp:141-151)
    mean[c] = m;
    batch_var[c] = var;
    run_mean[c] = (1.f - momentum) * m + momentum * run_mean[c];
    run_var[c] = (1.f - momentum) * var + momentum * run_var[c];
    inv_std[c] = powf(var + eps, -0.5f);
  }
}

__global__ void bn_finish_mean_kernel(const float* __restrict__ sum, float* __restrict__ mean, int C, double count) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) mean[c] = (float)((double)sum[c] / count);
}
__global__ void bn_finish_var_kernel(const float* __restrict__ sqdev, const float* __restrict__ mean, float* __restrict__ inv_std,
                                     float* __restrict__ batch_var, float* __restrict__ run_mean, float* __restrict__ run_var,
                                     int C, double count, float momentum, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float var = (float)((double)sqdev[c] / count);  // biased (bn_layer.cpp:141-151)
  batch_var[c] = var;
  run_mean[c] = (1.f - momentum) * mean[c] + momentum * run_mean[c];  // caffe_cpu_axpby(1 - m, batch, m, running)
  run_var[c] = (1.f - momentum) * var + momentum * run_var[c];
  inv_std[c] = powf(var + eps, -0.5f);
}

// elementwise over [rows][C/8] groups.  Thread t owns channel group g = t % G for its whole life (the per-channel
// constants sit in registers: the data load is the only memory instruction of the loop) and walks rows lane, lane + stride,
// ... two at a time.
__global__ void bn_apply_cl_kernel(ClView x, ClView y, const float* __restrict__ mean, const float* __restrict__ inv_std,
                                   const float* __restrict__ slope, const float* __restrict__ bias, int relu) {
  const int G = x.C / 8;
  const long long rows = x.outer * x.inner;
  const int lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  if (lane >= lanes) return;
  float mu[8], is[8], sl[8], bi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { mu[j] = mean[g * 8 + j]; is[j] = inv_std[g * 8 + j]; sl[j] = slope[g * 8 + j]; bi[j] = bias[g * 8 + j]; }
  auto apply = [&](uint4 raw) {
    float v[8];
    up8(raw, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xn = (v[j] - mu[j]) * is[j];  // x_norm first, as bn_layer.cpp:132-181 orders it
      const float o = xn * sl[j] + bi[j];        // (the backward kernels recompute exactly this for the ReLU mask)
      v[j] = relu ? fmaxf(o, 0.f) : o;
    }
    return pk8(v);
  };
  const long long stride = (long long)gridDim.x * lanes;
  long long r = (long long)blockIdx.x * lanes + lane;
  const __nv_bfloat16* xp = x.ptr + x.coff + g * 8;
  __nv_bfloat16* yp = y.ptr + y.coff + g * 8;
  for (; r + 3 * stride < rows; r += 4 * stride) {
    uint4 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = ld8(xp + (r + u * stride) * x.cs);
#pragma unroll
    for (int u = 0; u < 4; ++u) *reinterpret_cast<uint4*>(yp + (r + u * stride) * y.cs) = apply(a[u]);
  }
  for (; r < rows; r += stride) *reinterpret_cast<uint4*>(yp + r * y.cs) = apply(ld8(xp + r * x.cs));
}

__global__ void bn_bwd_apply_cl_kernel(ClView x, ClView y, ClView dy, ClView dx, const float* __restrict__ mean,
                                       const float* __restrict__ inv_std, const float* __restrict__ slope,
                                       const float* __restrict__ bias, const float* __restrict__ sums, float inv_count, int relu,
                                       int accumulate) {
  const int G = x.C / 8, C = x.C;
  const long long rows = x.outer * x.inner;
  const int lanes = blockDim.x / G;
  const int g = threadIdx.x % G, lane = threadIdx.x / G;
  if (lane >= lanes) return;
  float mu[8], is[8], sl[8], bi[8], k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = g * 8 + j;
    mu[j] = mean[c]; is[j] = inv_std[c]; sl[j] = slope[c]; bi[j] = bias[c];
    k1[j] = sl[j] * sums[c] * inv_count;
    k2[j] = sl[j] * sums[C + c] * inv_count;
  }
  auto apply = [&](uint4 xraw, uint4 graw, uint4 oraw) {
    float xv[8], gv[8], o[8];
    up8(xraw, xv);
    up8(graw, gv);
    if (accumulate) up8(oraw, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xn = (xv[j] - mu[j]) * is[j];
      if (relu && !(xn * sl[j] + bi[j] > 0.f)) gv[j] = 0.f;   // same mask as the forward pass, recomputed from x
      const float d = (sl[j] * gv[j] - k1[j] - xn * k2[j]) * is[j];
      o[j] = accumulate ? o[j] + d : d;
    }
    return pk8(o);
  };
  const long long stride = (long long)gridDim.x * lanes;
  long long r = (long long)blockIdx.x * lanes + lane;
  const __nv_bfloat16* xp = x.ptr + x.coff + g * 8;
  const __nv_bfloat16* gp = dy.ptr + dy.coff + g * 8;
  __nv_bfloat16* op = dx.ptr + dx.coff + g * 8;
  const uint4 z = make_uint4(0, 0, 0, 0);
  for (; r + stride < rows; r += 2 * stride) {
    uint4 a[2], b[2], c[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      a[u] = ld8(xp + (r + u * stride) * x.cs);
      b[u] = ld8(gp + (r + u * stride) * dy.cs);
      c[u] = accumulate ? ld8(op + (r + u * stride) * dx.cs) : z;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) *reinterpret_cast<uint4*>(op + (r + u * stride) * dx.cs) = apply(a[u], b[u], c[u]);
  }
  for (; r < rows; r += stride)
    *reinterpret_cast<uint4*>(op + r * dx.cs) = apply(ld8(xp + r * x.cs), ld8(gp + r * dy.cs), accumulate ? ld8(op + r * dx.cs) : z);
}
__global__ void bn_param_grads_kernel(const float* __restrict__ sums, float* __restrict__ dslope, float* __restrict__ dbias, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dslope) dslope[c] += sums[C + c];
  if (dbias) dbias[c] += sums[c];
}

// ------------------------------------------------------------------------------------------------
// pooling backward, gather form: one thread per (input position, 8 channels)
__global__ void pool_bwd_cl_kernel(const PoolParams p, const __nv_bfloat16* __restrict__ dy, long long dy_cs, int dy_coff,
                                   __nv_bfloat16* __restrict__ dx, long long dx_cs, int dx_coff, int accumulate) {
  const int G = p.C / 8;
  const long long total = (long long)p.NB * p.ID * p.IH * p.IW * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long r = t / G;
    const long long ipix = r;
    const int ix = (int)(r % p.IW); r /= p.IW;
    const int iy = (int)(r % p.IH); r /= p.IH;
    const int iz = (int)(r % p.ID);
    const long long n = r / p.ID;
    // output windows that contain (iz, iy, ix): o*s - pad <= i < o*s - pad + K
    const int oz_lo = max(0, (iz + p.pD - p.KD + p.sD) / p.sD), oz_hi = min(p.OD - 1, (iz + p.pD) / p.sD);
    const int oy_lo = max(0, (iy + p.pH - p.KH + p.sH) / p.sH), oy_hi = min(p.OH - 1, (iy + p.pH) / p.sH);
    const int ox_lo = max(0, (ix + p.pW - p.KW + p.sW) / p.sW), ox_hi = min(p.OW - 1, (ix + p.pW) / p.sW);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    float me[8];
    if (p.is_max) up8(ld8(p.x + ipix * p.x_cs + p.x_coff + g * 8), me);
    for (int oz = oz_lo; oz <= oz_hi; ++oz)
      for (int oy = oy_lo; oy <= oy_hi; ++oy)
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
          int z0 = oz * p.sD - p.pD, y0 = oy * p.sH - p.pH, x0 = ox * p.sW - p.pW;
          const long long opix = ((n * p.OD + oz) * p.OH + oy) * p.OW + ox;
          float gv[8];
          up8(ld8(dy + opix * dy_cs + dy_coff + g * 8), gv);
          if (!p.is_max) {
            const int z1 = min(z0 + p.KD, p.ID + p.pD), y1 = min(y0 + p.KH, p.IH + p.pH), x1 = min(x0 + p.KW, p.IW + p.pW);
            const float inv = 1.f / (float)((z1 - z0) * (y1 - y0) * (x1 - x0));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += gv[j] * inv;  // top_diff / pool_size (pooling_layer.cpp:352)
            continue;
          }
          // MAX: this element receives the window's gradient iff it is the first maximum in scan order, i.e. it is
          // >= every later element and > every earlier one (strictly-greater update, pooling_layer.cpp:206-222)
          const int z1 = min(z0 + p.KD, p.ID), y1 = min(y0 + p.KH, p.IH), x1 = min(x0 + p.KW, p.IW);
          z0 = max(z0, 0); y0 = max(y0, 0); x0 = max(x0, 0);
          bool win[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) win[j] = true;
          for (int z = z0; z < z1; ++z)
            for (int yy = y0; yy < y1; ++yy)
              for (int xx = x0; xx < x1; ++xx) {
                if (z == iz && yy == iy && xx == ix) continue;
                const bool earlier = (z < iz) || (z == iz && (yy < iy || (yy == iy && xx < ix)));
                const long long q = ((n * p.ID + z) * p.IH + yy) * p.IW + xx;
                float ov[8];
                up8(ld8(p.x + q * p.x_cs + p.x_coff + g * 8), ov);
#pragma unroll
                for (int j = 0; j < 8; ++j) win[j] = win[j] && (earlier ? me[j] > ov[j] : me[j] >= ov[j]);
              }
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (win[j]) acc[j] += gv[j];
        }
    __nv_bfloat16* o = dx + ipix * dx_cs + dx_coff + g * 8;
    if (accumulate) {
      float old[8];
      up8(ld8(o), old);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += old[j];
    }
    *reinterpret_cast<uint4*>(o) = pk8(acc);
  }
}

// MAX pooling backward in two passes (what caffe's max_idx_ mask does, pooling_layer.cpp:206-222 / :301-327, without
// keeping the mask alive between forward and backward):
//   1. pool_argmax_cl_kernel: per window and channel the position of the FIRST maximum in scan order, as a byte
//      (kz * KH + ky) * KW + kx relative to the unclipped window start (255 = empty window)
//   2. pool_max_bwd_mask_kernel: per input element, the windows that contain it (at most ceil(K/s) per axis) are looked up
//      in the mask: 8 bytes of mask + 16 bytes of dy per window instead of re-scanning K^3 inputs per window
// Both kernels run one CTA per (image, z, y) row: the row coordinates and the y/z window ranges are block-uniform, a
// thread only splits its item into (x, channel group) -- no 64-bit div/mod chain per 16 bytes moved.
__device__ __forceinline__ int div_stride(int a, int s) { return s == 1 ? a : (s == 2 ? (a >> 1) : a / s); }  // callers clamp at 0
__global__ void pool_argmax_cl_kernel(const PoolParams p, unsigned char* __restrict__ mask) {
  const int G = p.C / 8;
  unsigned int row = blockIdx.x;                      // ((n * OD + oz) * OH + oy)
  const int oy = (int)(row % (unsigned)p.OH); row /= (unsigned)p.OH;
  const int oz = (int)(row % (unsigned)p.OD);
  const long long n = row / (unsigned)p.OD;
  const int z0u = oz * p.sD - p.pD, y0u = oy * p.sH - p.pH;
  const int z1 = min(z0u + p.KD, p.ID), y1 = min(y0u + p.KH, p.IH);
  const int z0 = max(z0u, 0), y0 = max(y0u, 0);
  const long long obase = (long long)blockIdx.x * p.OW;
  for (int item = threadIdx.x; item < p.OW * G; item += blockDim.x) {
    const int ox = item / G, g = item - ox * G;
    const int x0u = ox * p.sW - p.pW;
    const int x1 = min(x0u + p.KW, p.IW), x0 = max(x0u, 0);
    float best[8];
    unsigned int idx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { best[j] = -FLT_MAX; idx[j] = 255u; }
    for (int z = z0; z < z1; ++z)
      for (int yy = y0; yy < y1; ++yy) {
        const __nv_bfloat16* px = p.x + (((n * p.ID + z) * p.IH + yy) * (long long)p.IW) * p.x_cs + p.x_coff + g * 8;
        const unsigned int code0 = (unsigned int)(((z - z0u) * p.KH + (yy - y0u)) * p.KW - x0u);
        for (int xx = x0; xx < x1; ++xx) {
          float v[8];
          up8(ld8(px + (long long)xx * p.x_cs), v);
          const unsigned int code = code0 + (unsigned int)xx;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (v[j] > best[j]) { best[j] = v[j]; idx[j] = code; }  // strictly greater: the first maximum wins
        }
      }
    uint2 o;
    o.x = idx[0] | (idx[1] << 8) | (idx[2] << 16) | (idx[3] << 24);
    o.y = idx[4] | (idx[5] << 8) | (idx[6] << 16) | (idx[7] << 24);
    *reinterpret_cast<uint2*>(mask + ((obase + ox) * G + g) * 8) = o;
  }
}
template <bool IS_MAX>
__global__ void pool_bwd_rows_kernel(const PoolParams p, const unsigned char* __restrict__ mask,
                                         const __nv_bfloat16* __restrict__ dy, long long dy_cs, int dy_coff,
                                         __nv_bfloat16* __restrict__ dx, long long dx_cs, int dx_coff, int accumulate) {
  const int G = p.C / 8;
  unsigned int row = blockIdx.x;                      // ((n * ID + iz) * IH + iy)
  const int iy = (int)(row % (unsigned)p.IH); row /= (unsigned)p.IH;
  const int iz = (int)(row % (unsigned)p.ID);
  const long long n = row / (unsigned)p.ID;
  // output windows that contain (iz, iy, ix): o*s - pad <= i < o*s - pad + K
  const int oz_lo = max(0, div_stride(iz + p.pD - p.KD + p.sD, p.sD)), oz_hi = min(p.OD - 1, div_stride(iz + p.pD, p.sD));
  const int oy_lo = max(0, div_stride(iy + p.pH - p.KH + p.sH, p.sH)), oy_hi = min(p.OH - 1, div_stride(iy + p.pH, p.sH));
  const long long ibase = (long long)blockIdx.x * p.IW;
  for (int item = threadIdx.x; item < p.IW * G; item += blockDim.x) {
    const int ix = item / G, g = item - ix * G;
    const int ox_lo = max(0, div_stride(ix + p.pW - p.KW + p.sW, p.sW)), ox_hi = min(p.OW - 1, div_stride(ix + p.pW, p.sW));
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int oz = oz_lo; oz <= oz_hi; ++oz)
      for (int oy = oy_lo; oy <= oy_hi; ++oy) {
        const long long orow = ((n * p.OD + oz) * p.OH + oy) * (long long)p.OW;
        const unsigned int me0 = (unsigned int)(((iz - (oz * p.sD - p.pD)) * p.KH + (iy - (oy * p.sH - p.pH))) * p.KW + ix + p.pW);
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
          const long long opix = orow + ox;
          const unsigned int me = me0 - (unsigned int)(ox * p.sW);
          float gv[8];
          up8(ld8(dy + opix * dy_cs + dy_coff + g * 8), gv);
          if (IS_MAX) {
            const uint2 m = *reinterpret_cast<const uint2*>(mask + (opix * G + g) * 8);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (((m.x >> (8 * j)) & 255u) == me) acc[j] += gv[j];
              if (((m.y >> (8 * j)) & 255u) == me) acc[4 + j] += gv[4 + j];
            }
          } else {
            // AVE: top_diff / pool_size, the window clipped at the PADDED border (pooling_layer.cpp:237-245, :352)
            const int z0 = oz * p.sD - p.pD, y0 = oy * p.sH - p.pH, x0 = ox * p.sW - p.pW;
            const int z1 = min(z0 + p.KD, p.ID + p.pD), y1 = min(y0 + p.KH, p.IH + p.pH), x1 = min(x0 + p.KW, p.IW + p.pW);
            const float inv = 1.f / (float)((z1 - z0) * (y1 - y0) * (x1 - x0));
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += gv[j] * inv;
          }
        }
      }
    __nv_bfloat16* o = dx + (ibase + ix) * dx_cs + dx_coff + g * 8;
    if (accumulate) {
      float old[8];
      up8(ld8(o), old);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += old[j];
    }
    *reinterpret_cast<uint4*>(o) = pk8(acc);
  }
}

// generic fp32 pooling backward on plain blobs (ECO-Full's segment consensus), gather form
__global__ void pool_f32_bwd_kernel(const PoolF32Params p, const float* __restrict__ dy, float* __restrict__ dx, int accumulate) {
  const long long total = (long long)p.NC * p.ID * p.IH * p.IW;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int ix = (int)(r % p.IW); r /= p.IW;
    const int iy = (int)(r % p.IH); r /= p.IH;
    const int iz = (int)(r % p.ID);
    const long long nc = r / p.ID;
    const int oz_lo = max(0, (iz + p.pD - p.KD + p.sD) / p.sD), oz_hi = min(p.OD - 1, (iz + p.pD) / p.sD);
    const int oy_lo = max(0, (iy + p.pH - p.KH + p.sH) / p.sH), oy_hi = min(p.OH - 1, (iy + p.pH) / p.sH);
    const int ox_lo = max(0, (ix + p.pW - p.KW + p.sW) / p.sW), ox_hi = min(p.OW - 1, (ix + p.pW) / p.sW);
    const float* px = p.x + nc * (long long)p.ID * p.IH * p.IW;
    const float me = px[((long long)iz * p.IH + iy) * p.IW + ix];
    float acc = 0.f;
    for (int oz = oz_lo; oz <= oz_hi; ++oz)
      for (int oy = oy_lo; oy <= oy_hi; ++oy)
        for (int ox = ox_lo; ox <= ox_hi; ++ox) {
          int z0 = oz * p.sD - p.pD, y0 = oy * p.sH - p.pH, x0 = ox * p.sW - p.pW;
          const float g = dy[(nc * p.OD + oz) * (long long)p.OH * p.OW + (long long)oy * p.OW + ox];
          if (!p.is_max) {
            const int z1 = min(z0 + p.KD, p.ID + p.pD), y1 = min(y0 + p.KH, p.IH + p.pH), x1 = min(x0 + p.KW, p.IW + p.pW);
            acc += g / (float)((z1 - z0) * (y1 - y0) * (x1 - x0));
            continue;
          }
          const int z1 = min(z0 + p.KD, p.ID), y1 = min(y0 + p.KH, p.IH), x1 = min(x0 + p.KW, p.IW);
          z0 = max(z0, 0); y0 = max(y0, 0); x0 = max(x0, 0);
          bool win = true;
          for (int z = z0; z < z1 && win; ++z)
            for (int yy = y0; yy < y1 && win; ++yy)
              for (int xx = x0; xx < x1; ++xx) {
                if (z == iz && yy == iy && xx == ix) continue;
                const bool earlier = (z < iz) || (z == iz && (yy < iy || (yy == iy && xx < ix)));
                const float o = px[((long long)z * p.IH + yy) * p.IW + xx];
                if (earlier ? !(me > o) : !(me >= o)) { win = false; break; }
              }
          if (win) acc += g;
        }
    dx[t] = accumulate ? dx[t] + acc : acc;
  }
}

__global__ void global_avg_bwd_cl_kernel(const float* __restrict__ dy, ClView dx, int accumulate) {
  const int G = dx.C / 8;
  const long long total = dx.outer * dx.inner * G;
  const float inv = 1.f / (float)dx.inner;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    const long long r = t / G;
    const long long o = r / dx.inner;
    float v[8];
    __nv_bfloat16* d = dx.ptr + r * dx.cs + dx.coff + g * 8;
    if (accumulate) up8(ld8(d), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gj = dy[o * dx.C + g * 8 + j] * inv;
      v[j] = accumulate ? v[j] + gj : gj;
    }
    *reinterpret_cast<uint4*>(d) = pk8(v);
  }
}

__global__ void cl_axpy_kernel(ClView x, ClView y, int accumulate) {
  const int G = x.C / 8;
  const long long total = x.outer * x.inner * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    const long long r = t / G;
    const uint4 xv = ld8(x.ptr + r * x.cs + x.coff + g * 8);
    __nv_bfloat16* d = y.ptr + r * y.cs + y.coff + g * 8;
    if (!accumulate) {
      *reinterpret_cast<uint4*>(d) = xv;
    } else {
      float a[8], b[8];
      up8(xv, a);
      up8(ld8(d), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] += a[j];
      *reinterpret_cast<uint4*>(d) = pk8(b);
    }
  }
}
__global__ void f32_axpy_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, float a, int accumulate) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    y[t] = accumulate ? fmaf(a, x[t], y[t]) : a * x[t];
}

__global__ void dilate_cl_kernel(ClView s, int OD, int OH, int OW, __nv_bfloat16* __restrict__ dst, int ED, int EH, int EW,
                                 int sD, int sH, int sW) {
  const int G = s.C / 8;
  const long long total = s.outer * s.inner * G;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int g = (int)(t % G);
    long long r = t / G;
    const long long pix = r;
    const int ox = (int)(r % OW); r /= OW;
    const int oy = (int)(r % OH); r /= OH;
    const int oz = (int)(r % OD);
    const long long n = r / OD;
    const long long dpix = ((n * ED + (long long)oz * sD) * EH + (long long)oy * sH) * EW + (long long)ox * sW;
    *reinterpret_cast<uint4*>(dst + dpix * s.C + g * 8) = ld8(s.ptr + pix * s.cs + s.coff + g * 8);
  }
}

// one thread per (dX position, 8 channels); src is dense [NB][OD][OH][OW][C]
__global__ void scatter_stride_cl_kernel(const __nv_bfloat16* __restrict__ src, int OD, int OH, int OW, ClView dx, int ID, int IH,
                                         int IW, int sD, int sH, int sW, int accumulate) {
  const unsigned G = (unsigned)dx.C / 8u;
  const unsigned long long total = (unsigned long long)dx.outer * dx.inner * G;
  for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < total;
       t += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned g = (unsigned)(t % G);
    unsigned long long r = t / G;
    const unsigned long long pix = r;
    const unsigned ix = (unsigned)(r % (unsigned)IW); r /= (unsigned)IW;
    const unsigned iy = (unsigned)(r % (unsigned)IH); r /= (unsigned)IH;
    const unsigned iz = (unsigned)(r % (unsigned)ID);
    const unsigned long long n = r / (unsigned)ID;
    const unsigned oz = iz / (unsigned)sD, oy = iy / (unsigned)sH, ox = ix / (unsigned)sW;
    const bool on = oz * sD == iz && oy * sH == iy && ox * sW == ix && oz < (unsigned)OD && oy < (unsigned)OH && ox < (unsigned)OW;
    __nv_bfloat16* d = dx.ptr + pix * dx.cs + dx.coff + g * 8;
    if (!on) {
      if (!accumulate) *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
      continue;
    }
    const uint4 v = ld8(src + ((((n * OD + oz) * OH + oy) * OW + ox) * (unsigned long long)dx.C) + g * 8);
    if (!accumulate) {
      *reinterpret_cast<uint4*>(d) = v;
    } else {
      float a[8], b[8];
      up8(v, a);
      up8(ld8(d), b);
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] += a[j];
      *reinterpret_cast<uint4*>(d) = pk8(b);
    }
  }
}

// counter-based RNG: one 64-bit mix (splitmix64 finaliser) per element; the same (seed, index) gives the same draw in
// the forward and the backward pass, so no mask is stored
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
__global__ void dropout_f32_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, uint32_t thres, float scale,
                                   uint64_t seed) {
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    y[t] = hash_u32(seed, (uint64_t)t) >= thres ? x[t] * scale : 0.f;  // keep with probability 1 - ratio
}

// ------------------------------------------------------------------------------------------------
// inner product backward: dW[n][k] += sum_m dy[m][n] x[m][k]; db[n] += sum_m dy[m][n]; dx[m][k] (+)= sum_n dy[m][n] W[n][k]
__global__ void ip_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ dw,
                                float* __restrict__ db, int M, int N, int K) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)N * K) return;
  const int k = (int)(t % K), n = (int)(t / K);
  float acc = 0.f;
  for (int m = 0; m < M; ++m) acc = fmaf(dy[(long long)m * N + n], x[(long long)m * K + k], acc);
  dw[t] += acc;
  if (k == 0 && db) {
    float b = 0.f;
    for (int m = 0; m < M; ++m) b += dy[(long long)m * N + n];
    db[n] += b;
  }
}
__global__ void ip_bwd_x_kernel(const float* __restrict__ w, const float* __restrict__ dy, float* __restrict__ dx, int M, int N,
                                int K, int accumulate) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)M * K) return;
  const int k = (int)(t % K), m = (int)(t / K);
  float acc = 0.f;
  for (int n = 0; n < N; ++n) acc = fmaf(dy[(long long)m * N + n], w[(long long)n * K + k], acc);
  dx[t] = accumulate ? dx[t] + acc : acc;
}

// one block per row
__global__ void softmax_loss_fwd_kernel(const float* __restrict__ x, const float* __restrict__ label, float* __restrict__ prob,
                                        float* __restrict__ loss, int M, int N) {
  __shared__ float sred[32];
  const int m = blockIdx.x;
  const float* xp = x + (long long)m * N;
  float mx = -FLT_MAX;
  for (int n = threadIdx.x; n < N; n += blockDim.x) mx = fmaxf(mx, xp[n]);
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = sred[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmaxf(mx, sred[w]);
  __syncthreads();
  float s = 0.f;
  for (int n = threadIdx.x; n < N; n += blockDim.x) s += expf(xp[n] - mx);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
  for (int n = threadIdx.x; n < N; n += blockDim.x) prob[(long long)m * N + n] = expf(xp[n] - mx) / s;
  if (threadIdx.x == 0) {
    const int lab = (int)label[m];
    const float p = (lab >= 0 && lab < N) ? expf(xp[lab] - mx) / s : 1.f;
    atomicAdd(loss, -logf(fmaxf(p, FLT_MIN)) / (float)M);
  }
}
__global__ void softmax_loss_bwd_kernel(const float* __restrict__ prob, const float* __restrict__ label, float* __restrict__ dx,
                                        int M, int N, float scale) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)M * N) return;
  const int n = (int)(t % N), m = (int)(t / N);
  dx[t] = (prob[t] - ((int)label[m] == n ? 1.f : 0.f)) * scale;
}
__global__ void accuracy_kernel(const float* __restrict__ x, const float* __restrict__ label, float* __restrict__ acc, int M, int N,
                                int top_k) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int lab = (int)label[m];
  if (lab < 0 || lab >= N) return;
  const float v = x[(long long)m * N + lab];
  int ahead = 0;  // entries ranked before the label by std::greater<pair<value, index>>
  for (int n = 0; n < N; ++n) {
    const float o = x[(long long)m * N + n];
    if (o > v || (o == v && n > lab)) ++ahead;
  }
  if (ahead < top_k) atomicAdd(acc, 1.f / (float)M);
}

// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin, int taps,
                                   int cblocks, long long Ktotal) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(t % taps);
    const long long oc = t / taps;
    const int ch = (int)(oc % Cin);
    const long long o = oc / Cin;
    dst[o * Ktotal + ((long long)tap * cblocks + ch / 64) * 64 + ch % 64] = __float2bfloat16_rn(w[t]);
  }
}
__global__ void pack_stem_w_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, long long Ktotal) {
  const long long total = (long long)Cout * 3 * 49;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(t % 7), ky = (int)((t / 7) % 7), ch = (int)((t / 49) % 3);
    const long long o = t / 147;
    const int ty = ky >> 1, dy = ky & 1, tx = kx >> 1, dx = kx & 1;
    dst[o * Ktotal + ty * 64 + tx * 16 + (dy * 2 + dx) * 3 + ch] = __float2bfloat16_rn(w[t]);
  }
}
__global__ void pack_conv_w_dgrad_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ dst, int Cout, int Cin, int KD,
                                         int KH, int KW, int oblocks, long long Ktotal) {
  const int taps = KD * KH * KW;
  const long long total = (long long)Cout * Cin * taps;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(t % taps);
    const long long oc = t / taps;
    const int ch = (int)(oc % Cin);
    const int o = (int)(oc / Cin);
    const int ftap = taps - 1 - tap;  // flip every spatial axis
    dst[(long long)ch * Ktotal + ((long long)ftap * oblocks + o / 64) * 64 + o % 64] = __float2bfloat16_rn(w[t]);
  }
}
__global__ void wgrad_finish_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int Cout, int Cin, int taps,
                                    int cin_ld, int cout_ld) {
  const long long total = (long long)Cout * Cin * taps;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(t % taps);
    const long long oc = t / taps;
    const int ch = (int)(oc % Cin);
    const int o = (int)(oc / Cin);
    dw[t] += scratch[((long long)tap * cin_ld + ch) * cout_ld + o];
  }
}
__global__ void wgrad_finish_stem_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int Cout, int cout_ld) {
  const long long total = (long long)Cout * 147;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int kx = (int)(t % 7), ky = (int)((t / 7) % 7), ch = (int)((t / 49) % 3);
    const int o = (int)(t / 147);
    const int ty = ky >> 1, dy = ky & 1, tx = kx >> 1, dx = kx & 1;
    // the stem GEMM sees a 4(h) x 1(w) filter over 64-value windows: tap = ty, "channel" = tx * 16 + cell value
    dw[t] += scratch[((long long)ty * 64 + tx * 16 + (dy * 2 + dx) * 3 + ch) * cout_ld + o];
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  __shared__ float sred[32];
  float s = 0.f;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    s = fmaf(x[t], x[t], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < (blockDim.x >> 5) ? sred[threadIdx.x] : 0.f;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (threadIdx.x == 0) atomicAdd(out, s);
  }
}
__global__ void scale_kernel(float* __restrict__ x, long long n, const float* __restrict__ scale) {
  const float a = *scale;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) x[t] *= a;
}
__global__ void clip_factor_kernel(const float* __restrict__ sumsq, float clip, float norm, float* __restrict__ scale) {
  // ClipGradients runs on the accumulated (not yet iter_size-normalised) diffs, solver.cpp:637-660
  // `norm` = 1 / world: the reference scales the exchanged diffs by 1 / MPI_all_rank before ApplyUpdate (solver.cpp:332-337)
  const float l2 = sqrtf(*sumsq) * norm;
  *scale = (clip >= 0.f && l2 > clip) ? clip / l2 : 1.f;
}
__global__ void sgd_update_kernel(float* __restrict__ w, float* __restrict__ diff, float* __restrict__ hist, long long n, float rate,
                                  float momentum, float decay, float norm, const float* __restrict__ clip_scale, int nesterov) {
  const float cs = clip_scale ? *clip_scale : 1.f;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    // solver.cpp: ClipGradients (scale_diff) -> Normalize (1 / iter_size) -> Regularize (L2) -> ComputeUpdateValue -> Update
    const float g = diff[t] * cs * norm + decay * w[t];
    const float h0 = hist[t];
    const float h1 = momentum * h0 + rate * g;
    hist[t] = h1;
    const float upd = nesterov ? (1.f + momentum) * h1 - momentum * h0 : h1;
    diff[t] = upd;  // caffe leaves the update value in the diff (Net::Update subtracts it)
    w[t] -= upd;
  }
}

}  // namespace

// ================================================================================================
cudaError_t launch_colsum_cl(ClView x, float* out, float* scratch, cudaStream_t st) {
  return launch_colreduce<0>(x, x, x, nullptr, nullptr, nullptr, nullptr, 0, out, scratch, st);
}
cudaError_t launch_colsqdev_cl(ClView x, const float* mean, float* out, float* scratch, cudaStream_t st) {
  return launch_colreduce<1>(x, x, x, mean, nullptr, nullptr, nullptr, 0, out, scratch, st);
}
cudaError_t launch_bn_bwd_sums_cl(ClView x, ClView y, ClView dy, const float* mean, const float* inv_std, const float* slope,
                                  const float* bias, int relu, float* out, float* scratch, cudaStream_t st) {
  return launch_colreduce<2>(x, y, dy, mean, inv_std, slope, bias, relu, out, scratch, st);
}
cudaError_t launch_bn_stats_cl(ClView x, float* mean, float* inv_std, float* batch_var, float* run_mean, float* run_var,
                               float momentum, float eps, float* scratch, cudaStream_t st) {
  const long long rows = x.outer * x.inner;
  if (rows == 0 || x.C == 0) return cudaSuccess;
  if (x.C % 8 != 0 || x.C > kColReduceMaxC || !scratch) return cudaErrorInvalidValue;
  const int G = x.C / 8;
  int threads = 256;
  while (threads < G) threads *= 2;
  if (threads > 1024) return cudaErrorInvalidValue;
  const int lanes = threads / G;
  const size_t smem = ((size_t)lanes * 2 * x.C + lanes) * sizeof(float);
  long long blocks = (rows + (long long)lanes * 8 - 1) / ((long long)lanes * 8);
  if (blocks > kColReduceMaxBlocks) blocks = kColReduceMaxBlocks;
  if (blocks < 1) blocks = 1;
  float* counts = scratch + (size_t)kColReduceMaxBlocks * 2 * x.C;   // behind the largest partial table of this layer
  if ((size_t)kColReduceMaxBlocks * 2 * x.C + kColReduceMaxBlocks > kColReduceScratchFloats) return cudaErrorInvalidValue;
  bn_welford_cl_kernel<<<(unsigned)blocks, threads, smem, st>>>(x, scratch, counts);
  bn_welford_finish_kernel<<<(x.C + 31) / 32, dim3(32, 8, 1), 0, st>>>(scratch, counts, (int)blocks, x.C, mean, inv_std, batch_var,
                                                                       run_mean, run_var, momentum, eps);
  return cudaGetLastError();
}
cudaError_t launch_bn_finish_mean(const float* sum, float* mean, int C, double count, cudaStream_t st) {
  bn_finish_mean_kernel<<<(C + 127) / 128, 128, 0, st>>>(sum, mean, C, count);
  return cudaGetLastError();
}
cudaError_t launch_bn_finish_var(const float* sqdev, const float* mean, float* inv_std, float* batch_var, float* run_mean,
                                 float* run_var, int C, double count, float momentum, float eps, cudaStream_t st) {
  bn_finish_var_kernel<<<(C + 127) / 128, 128, 0, st>>>(sqdev, mean, inv_std, batch_var, run_mean, run_var, C, count, momentum, eps);
  return cudaGetLastError();
}
// block shape of the row-walking elementwise kernels: G = C/8 channel groups x `lanes` rows per block
static inline bool rowwalk_shape(const ClView& x, int rows_per_thread, int max_blocks, int* threads, int* blocks) {
  if (x.C % 8 != 0 || x.C / 8 > 1024) return false;
  const int G = x.C / 8;
  int t = kT;
  while (t < G) t *= 2;
  const int lanes = t / G;
  const long long rows = x.outer * x.inner;
  long long b = (rows + (long long)lanes * rows_per_thread - 1) / ((long long)lanes * rows_per_thread);
  if (b > max_blocks) b = max_blocks;
  if (b < 1) b = 1;
  *threads = t;
  *blocks = (int)b;
  return true;
}
cudaError_t launch_bn_apply_cl(ClView x, ClView y, const float* mean, const float* inv_std, const float* slope,
                               const float* bias, int relu, cudaStream_t st) {
  if (x.outer * x.inner == 0 || x.C == 0) return cudaSuccess;
  int threads, blocks;
  if (!rowwalk_shape(x, 4, 148 * 6, &threads, &blocks)) return cudaErrorInvalidValue;
  bn_apply_cl_kernel<<<blocks, threads, 0, st>>>(x, y, mean, inv_std, slope, bias, relu);
  return cudaGetLastError();
}
cudaError_t launch_bn_bwd_apply_cl(ClView x, ClView y, ClView dy, ClView dx, const float* mean, const float* inv_std,
                                   const float* slope, const float* bias, const float* sums, double count, int relu, int accumulate,
                                   float* dslope, float* dbias, cudaStream_t st) {
  if (x.outer * x.inner == 0 || x.C == 0) return cudaSuccess;
  if (dx.ptr) {
    int threads, blocks;
    if (!rowwalk_shape(x, 2, 148 * 4, &threads, &blocks)) return cudaErrorInvalidValue;
    bn_bwd_apply_cl_kernel<<<blocks, threads, 0, st>>>(x, y, dy, dx, mean, inv_std, slope, bias, sums, (float)(1.0 / count), relu, accumulate);
  }
  if (dslope || dbias) bn_param_grads_kernel<<<(x.C + 127) / 128, 128, 0, st>>>(sums, dslope, dbias, x.C);
  return cudaGetLastError();
}
cudaError_t launch_pool_bwd_cl(const PoolParams& p, const __nv_bfloat16* dy, long long dy_cs, int dy_coff, __nv_bfloat16* dx,
                               long long dx_cs, int dx_coff, int accumulate, unsigned char* mask, cudaStream_t st) {
  const long long n = (long long)p.NB * p.ID * p.IH * p.IW * (p.C / 8);
  if (n == 0) return cudaSuccess;
  const long long orows = (long long)p.NB * p.OD * p.OH, irows = (long long)p.NB * p.ID * p.IH;
  const int G = p.C / 8;
  auto threads_for = [&](int w) { const int items = w * G; return items >= 256 ? 256 : (items >= 128 ? 128 : 64); };
  if (orows <= 0x7fffffffLL && irows <= 0x7fffffffLL) {
    if (p.is_max && mask && p.KD * p.KH * p.KW < 255) {
      pool_argmax_cl_kernel<<<(unsigned)orows, threads_for(p.OW), 0, st>>>(p, mask);
      pool_bwd_rows_kernel<true><<<(unsigned)irows, threads_for(p.IW), 0, st>>>(p, mask, dy, dy_cs, dy_coff, dx, dx_cs, dx_coff, accumulate);
      return cudaGetLastError();
    }
    if (!p.is_max) {
      pool_bwd_rows_kernel<false><<<(unsigned)irows, threads_for(p.IW), 0, st>>>(p, nullptr, dy, dy_cs, dy_coff, dx, dx_cs, dx_coff, accumulate);
      return cudaGetLastError();
    }
  }
  pool_bwd_cl_kernel<<<grid_for(n, 32), kT, 0, st>>>(p, dy, dy_cs, dy_coff, dx, dx_cs, dx_coff, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_pool_f32_bwd(const PoolF32Params& p, const float* dy, float* dx, int accumulate, cudaStream_t st) {
  const long long n = (long long)p.NC * p.ID * p.IH * p.IW;
  if (n == 0) return cudaSuccess;
  pool_f32_bwd_kernel<<<grid_for(n, 16), kT, 0, st>>>(p, dy, dx, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_global_avg_bwd_cl(const float* dy, ClView dx, int accumulate, cudaStream_t st) {
  const long long n = dx.outer * dx.inner * (dx.C / 8);
  if (n == 0) return cudaSuccess;
  global_avg_bwd_cl_kernel<<<grid_for(n, 16), kT, 0, st>>>(dy, dx, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_cl_axpy(ClView x, ClView y, int accumulate, cudaStream_t st) {
  const long long n = x.outer * x.inner * (x.C / 8);
  if (n == 0) return cudaSuccess;
  cl_axpy_kernel<<<grid_for(n, 16), kT, 0, st>>>(x, y, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_f32_axpy(const float* x, float* y, long long n, float a, int accumulate, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  f32_axpy_kernel<<<grid_for(n, 16), kT, 0, st>>>(x, y, n, a, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_dilate_cl(ClView src, int OD, int OH, int OW, __nv_bfloat16* dst, int ED, int EH, int EW, int sD, int sH,
                             int sW, cudaStream_t st) {
  const long long n = src.outer * src.inner * (src.C / 8);
  if (n == 0) return cudaSuccess;
  dilate_cl_kernel<<<grid_for(n, 16), kT, 0, st>>>(src, OD, OH, OW, dst, ED, EH, EW, sD, sH, sW);
  return cudaGetLastError();
}
cudaError_t launch_scatter_stride_cl(const __nv_bfloat16* src, int C, int OD, int OH, int OW, ClView dx, int ID, int IH, int IW,
                                     int sD, int sH, int sW, int accumulate, cudaStream_t st) {
  if (dx.C != C || C % 8 != 0 || dx.inner != (long long)ID * IH * IW) return cudaErrorInvalidValue;
  const long long n = dx.outer * dx.inner * (C / 8);
  if (n == 0) return cudaSuccess;
  scatter_stride_cl_kernel<<<grid_for(n, 16), kT, 0, st>>>(src, OD, OH, OW, dx, ID, IH, IW, sD, sH, sW, accumulate);
  return cudaGetLastError();
}
cudaError_t launch_dropout_f32(const float* x, float* y, long long n, float ratio, uint64_t seed, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  // uint_thres_ = UINT_MAX * threshold (dropout_layer.cpp:20-21); keep when the draw is >= it
  const uint32_t thres = (uint32_t)((double)0xFFFFFFFFu * (double)ratio);
  dropout_f32_kernel<<<grid_for(n, 8), kT, 0, st>>>(x, y, n, thres, 1.f / (1.f - ratio), seed);
  return cudaGetLastError();
}
cudaError_t launch_inner_product_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int M,
                                     int N, int K, int accumulate_dx, cudaStream_t st) {
  if (dw) ip_bwd_w_kernel<<<(unsigned)(((long long)N * K + kT - 1) / kT), kT, 0, st>>>(x, dy, dw, db, M, N, K);
  if (dx) ip_bwd_x_kernel<<<(unsigned)(((long long)M * K + kT - 1) / kT), kT, 0, st>>>(w, dy, dx, M, N, K, accumulate_dx);
  return cudaGetLastError();
}
cudaError_t launch_softmax_loss_fwd(const float* x, const float* label, float* prob, float* loss, int M, int N, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (e != cudaSuccess) return e;
  if (M == 0) return cudaSuccess;
  softmax_loss_fwd_kernel<<<M, 128, 0, st>>>(x, label, prob, loss, M, N);
  return cudaGetLastError();
}
cudaError_t launch_softmax_loss_bwd(const float* prob, const float* label, float* dx, int M, int N, float loss_weight,
                                    cudaStream_t st) {
  if (M == 0) return cudaSuccess;
  softmax_loss_bwd_kernel<<<(unsigned)(((long long)M * N + kT - 1) / kT), kT, 0, st>>>(prob, label, dx, M, N, loss_weight / (float)M);
  return cudaGetLastError();
}
cudaError_t launch_accuracy(const float* x, const float* label, float* acc, int M, int N, int top_k, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(float), st);
  if (e != cudaSuccess) return e;
  if (M == 0) return cudaSuccess;
  accuracy_kernel<<<(M + 127) / 128, 128, 0, st>>>(x, label, acc, M, N, top_k);
  return cudaGetLastError();
}
cudaError_t launch_pack_conv_w(const float* w, __nv_bfloat16* dst, int Cout, int Cin, int taps, int cblocks, long long Ktotal,
                               cudaStream_t st) {
  pack_conv_w_kernel<<<grid_for((long long)Cout * Cin * taps, 16), kT, 0, st>>>(w, dst, Cout, Cin, taps, cblocks, Ktotal);
  return cudaGetLastError();
}
cudaError_t launch_pack_stem_w(const float* w, __nv_bfloat16* dst, int Cout, long long Ktotal, cudaStream_t st) {
  pack_stem_w_kernel<<<grid_for((long long)Cout * 147, 16), kT, 0, st>>>(w, dst, Cout, Ktotal);
  return cudaGetLastError();
}
cudaError_t launch_pack_conv_w_dgrad(const float* w, __nv_bfloat16* dst, int Cout, int Cin, int KD, int KH, int KW, int oblocks,
                                     long long Ktotal, cudaStream_t st) {
  pack_conv_w_dgrad_kernel<<<grid_for((long long)Cout * Cin * KD * KH * KW, 16), kT, 0, st>>>(w, dst, Cout, Cin, KD, KH, KW, oblocks, Ktotal);
  return cudaGetLastError();
}
cudaError_t launch_wgrad_finish(const float* scratch, float* dw, int Cout, int Cin, int taps, int cin_ld, int cout_ld,
                                cudaStream_t st) {
  wgrad_finish_kernel<<<grid_for((long long)Cout * Cin * taps, 16), kT, 0, st>>>(scratch, dw, Cout, Cin, taps, cin_ld, cout_ld);
  return cudaGetLastError();
}
cudaError_t launch_wgrad_finish_stem(const float* scratch, float* dw, int Cout, int cout_ld, cudaStream_t st) {
  wgrad_finish_stem_kernel<<<grid_for((long long)Cout * 147, 16), kT, 0, st>>>(scratch, dw, Cout, cout_ld);
  return cudaGetLastError();
}
cudaError_t launch_sumsq(const float* x, long long n, float* out, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  sumsq_kernel<<<grid_for(n, 4), kT, 0, st>>>(x, n, out);
  return cudaGetLastError();
}
cudaError_t launch_scale(float* x, long long n, const float* scale_dev, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  scale_kernel<<<grid_for(n, 8), kT, 0, st>>>(x, n, scale_dev);
  return cudaGetLastError();
}
cudaError_t launch_sgd_update(float* w, float* diff, float* hist, long long n, float rate, float momentum, float decay, float norm,
                              const float* clip_scale_dev, int nesterov, cudaStream_t st) {
  if (n == 0) return cudaSuccess;
  sgd_update_kernel<<<grid_for(n, 8), kT, 0, st>>>(w, diff, hist, n, rate, momentum, decay, norm, clip_scale_dev, nesterov);
  return cudaGetLastError();
}
cudaError_t launch_clip_factor(const float* sumsq, float clip, float norm, float* scale, cudaStream_t st) {
  clip_factor_kernel<<<1, 1, 0, st>>>(sumsq, clip, norm, scale);
  return cudaGetLastError();
}

}  // namespace eco
