/*
 * oracle/ref_cpu.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (fp32, NCHW / NCDHW, row-major) of the caffe_3d layers that
 * make up ECO's hot path.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library; the product
 * (libeco_b200.so) never links or calls it.
 *
 * caffe_3d itself cannot be compiled in this environment (glog/gflags/boost/
 * CBLAS/protoc are absent) and its CPU path cannot execute the
 * 5-D head anyway (F1/F2), so this is a restatement, pinned by
 *   - the reference's own known-answer tests (tests/golden/, tests/test_oracle_golden.py)
 *   - an independent torch-CPU functional graph (tests/test_oracle_vs_torch.py).
 *
 * Each function cites the reference file:line whose algorithm it follows
 * (paths relative to caffe_3d/).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define REF_MAX_SP 3

int ref_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void ref_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* Convolution output size: src/caffe/layers/conv_layer.cpp:12-25            */
/*   out = (in + 2*pad - kernel) / stride + 1   (integer floor division)     */
int ref_conv_out_dim(int in, int kernel, int stride, int pad) {
  return (in + 2 * pad - kernel) / stride + 1;
}

/* Pooling output size: src/caffe/layers/pooling_layer.cpp:131-147           */
/*   out = ceil((in + 2*pad - kernel) / stride) + 1, and one less if padding  */
/*   is used and the last window would start inside the high-side padding.    */
int ref_pool_out_dim(int in, int kernel, int stride, int pad) {
  int out = (int)ceilf((float)(in + 2 * pad - kernel) / (float)stride) + 1;
  if (pad) {
    if ((out - 1) * stride >= in + pad) --out;
  }
  return out;
}

/* ------------------------------------------------------------------------ */
/* im2col, N-D.  src/caffe/util/im2col.cpp:28-64 (2-D) and :91-158 (N-D):     */
/* column row index = ((c*kD + kz)*kH + ky)*kW + kx ; column = output position */
/* in row-major (z,y,x) order; taps that fall in the zero padding give 0.      */
static void im2col_nd(const float* im, int cin, int nsp, const int* in_shape,
                      const int* out_shape, const int* kernel, const int* stride,
                      const int* pad, float* col) {
  int k[REF_MAX_SP] = {1, 1, 1}, s[REF_MAX_SP] = {1, 1, 1}, p[REF_MAX_SP] = {0, 0, 0};
  int is[REF_MAX_SP] = {1, 1, 1}, os[REF_MAX_SP] = {1, 1, 1};
  /* right-align into 3 spatial axes so 1-D/2-D reuse the 3-D loops */
  for (int i = 0; i < nsp; ++i) {
    int j = REF_MAX_SP - nsp + i;
    k[j] = kernel[i]; s[j] = stride[i]; p[j] = pad[i];
    is[j] = in_shape[i]; os[j] = out_shape[i];
  }
  const long in_sp = (long)is[0] * is[1] * is[2];
  const long out_sp = (long)os[0] * os[1] * os[2];
  const int ktaps = k[0] * k[1] * k[2];
  const long rows = (long)cin * ktaps;
#pragma omp parallel for schedule(static)
  for (long r = 0; r < rows; ++r) {
    const int c = (int)(r / ktaps);
    int t = (int)(r % ktaps);
    const int kx = t % k[2]; t /= k[2];
    const int ky = t % k[1]; t /= k[1];
    const int kz = t;
    const float* src = im + (long)c * in_sp;
    float* dst = col + r * out_sp;
    for (int oz = 0; oz < os[0]; ++oz) {
      const int iz = oz * s[0] - p[0] + kz;
      for (int oy = 0; oy < os[1]; ++oy) {
        const int iy = oy * s[1] - p[1] + ky;
        float* d = dst + ((long)oz * os[1] + oy) * os[2];
        if (iz < 0 || iz >= is[0] || iy < 0 || iy >= is[1]) {
          memset(d, 0, sizeof(float) * (size_t)os[2]);
          continue;
        }
        const float* srow = src + ((long)iz * is[1] + iy) * is[2];
        for (int ox = 0; ox < os[2]; ++ox) {
          const int ix = ox * s[2] - p[2] + kx;
          d[ox] = (ix >= 0 && ix < is[2]) ? srow[ix] : 0.0f;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------ */
/* Row-major SGEMM  C[M,N] = A[M,K] * B[K,N]  (what caffe_cpu_gemm hands to    */
/* cblas_sgemm, src/caffe/util/math_functions.cpp:13-21; the BLAS itself is a  */
/* third-party library whose summation order the reference does not pin).      */
#define MR 8
#define NR 32
#if defined(__GNUC__) && defined(__x86_64__)
#define CLONES __attribute__((target_clones("avx512f", "avx2,fma", "default")))
#else
#define CLONES
#endif

CLONES
static void gemm_block(const float* A, const float* B, float* C, int mr, int nr,
                       int K, long lda, long ldb, long ldc) {
  float acc[MR][NR];
  for (int i = 0; i < MR; ++i)
    for (int j = 0; j < NR; ++j) acc[i][j] = 0.0f;
  if (mr == MR && nr == NR) {
    for (int kk = 0; kk < K; ++kk) {
      const float* b = B + (long)kk * ldb;
      for (int i = 0; i < MR; ++i) {
        const float a = A[(long)i * lda + kk];
#pragma omp simd
        for (int j = 0; j < NR; ++j) acc[i][j] += a * b[j];
      }
    }
  } else {
    for (int kk = 0; kk < K; ++kk) {
      const float* b = B + (long)kk * ldb;
      for (int i = 0; i < mr; ++i) {
        const float a = A[(long)i * lda + kk];
        for (int j = 0; j < nr; ++j) acc[i][j] += a * b[j];
      }
    }
  }
  for (int i = 0; i < mr; ++i)
    for (int j = 0; j < nr; ++j) C[(long)i * ldc + j] = acc[i][j];
}

void ref_sgemm(const float* A, const float* B, float* C, int M, int N, int K) {
  const long nb = ((long)N + NR - 1) / NR;
  const long mb = ((long)M + MR - 1) / MR;
#pragma omp parallel for schedule(dynamic, 4)
  for (long t = 0; t < nb * mb; ++t) {
    const long jb = t / mb, ib = t % mb;
    const int i0 = (int)(ib * MR), j0 = (int)(jb * NR);
    const int mr = M - i0 < MR ? M - i0 : MR;
    const int nr = N - j0 < NR ? N - j0 : NR;
    gemm_block(A + (long)i0 * K, B + j0, C + (long)i0 * N + j0, mr, nr, K, K, N, N);
  }
}

/* ------------------------------------------------------------------------ */
/* Convolution forward.                                                       */
/* src/caffe/layers/conv_layer.cpp:28-43 (loop over images),                   */
/* src/caffe/layers/base_conv_layer.cpp:264-287 (forward_cpu_gemm: im2col then */
/* W[Cout,K] x col[K,M]; skipped im2col for 1x1/s1/p0 :112-117; forward_cpu_bias*/
/* adds bias[o] to every output position).  Cross-correlation, no kernel flip.  */
/* x: [num, cin, in_shape...], w: [cout, cin, kernel...], y: [num, cout, out...] */
int ref_conv_forward(const float* x, const float* w, const float* b, float* y,
                     int num, int cin, int cout, int nsp, const int* in_shape,
                     const int* kernel, const int* stride, const int* pad) {
  if (nsp < 1 || nsp > REF_MAX_SP) return -1;
  int out_shape[REF_MAX_SP];
  long in_sp = 1, out_sp = 1, ktaps = 1;
  int is_1x1 = 1;
  for (int i = 0; i < nsp; ++i) {
    out_shape[i] = ref_conv_out_dim(in_shape[i], kernel[i], stride[i], pad[i]);
    if (out_shape[i] <= 0) return -2;
    in_sp *= in_shape[i]; out_sp *= out_shape[i]; ktaps *= kernel[i];
    is_1x1 &= (kernel[i] == 1 && stride[i] == 1 && pad[i] == 0);
  }
  const long K = (long)cin * ktaps;
  float* col = NULL;
  if (!is_1x1) {
    col = (float*)malloc(sizeof(float) * (size_t)(K * out_sp));
    if (!col) return -3;
  }
  for (int n = 0; n < num; ++n) {
    const float* xn = x + (long)n * cin * in_sp;
    float* yn = y + (long)n * cout * out_sp;
    const float* colp = xn;
    if (!is_1x1) {
      im2col_nd(xn, cin, nsp, in_shape, out_shape, kernel, stride, pad, col);
      colp = col;
    }
    ref_sgemm(w, colp, yn, cout, (int)out_sp, (int)K);
    if (b) {
#pragma omp parallel for schedule(static)
      for (int o = 0; o < cout; ++o) {
        float* yo = yn + (long)o * out_sp;
        const float bo = b[o];
        for (long i = 0; i < out_sp; ++i) yo[i] += bo;
      }
    }
  }
  free(col);
  return 0;
}

/* Naive direct convolution: the executable spec the reference's tests use     */
/* (src/caffe/test/test_convolution_layer.cpp:18-134, caffe_conv, 4-D and 5-D). */
/* Kept separate from the im2col path so the two restatements check each other. */
int ref_conv_forward_naive(const float* x, const float* w, const float* b, float* y,
                           int num, int cin, int cout, int nsp, const int* in_shape,
                           const int* kernel, const int* stride, const int* pad) {
  int k[3] = {1, 1, 1}, s[3] = {1, 1, 1}, p[3] = {0, 0, 0}, is[3] = {1, 1, 1}, os[3] = {1, 1, 1};
  if (nsp < 1 || nsp > 3) return -1;
  for (int i = 0; i < nsp; ++i) {
    int j = 3 - nsp + i;
    k[j] = kernel[i]; s[j] = stride[i]; p[j] = pad[i]; is[j] = in_shape[i];
    os[j] = ref_conv_out_dim(in_shape[i], kernel[i], stride[i], pad[i]);
  }
  const long in_sp = (long)is[0] * is[1] * is[2], out_sp = (long)os[0] * os[1] * os[2];
  const long kt = (long)k[0] * k[1] * k[2];
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < num; ++n)
    for (int o = 0; o < cout; ++o) {
      float* yo = y + ((long)n * cout + o) * out_sp;
      for (int oz = 0; oz < os[0]; ++oz)
        for (int oy = 0; oy < os[1]; ++oy)
          for (int ox = 0; ox < os[2]; ++ox) {
            double acc = 0.0;
            for (int c = 0; c < cin; ++c) {
              const float* xc = x + ((long)n * cin + c) * in_sp;
              const float* wc = w + ((long)o * cin + c) * kt;
              for (int kz = 0; kz < k[0]; ++kz) {
                int iz = oz * s[0] - p[0] + kz;
                if (iz < 0 || iz >= is[0]) continue;
                for (int ky = 0; ky < k[1]; ++ky) {
                  int iy = oy * s[1] - p[1] + ky;
                  if (iy < 0 || iy >= is[1]) continue;
                  for (int kx = 0; kx < k[2]; ++kx) {
                    int ix = ox * s[2] - p[2] + kx;
                    if (ix < 0 || ix >= is[2]) continue;
                    acc += (double)xc[((long)iz * is[1] + iy) * is[2] + ix] *
                           (double)wc[((long)kz * k[1] + ky) * k[2] + kx];
                  }
                }
              }
            }
            yo[((long)oz * os[1] + oy) * os[2] + ox] = (float)acc + (b ? b[o] : 0.0f);
          }
    }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* BN, TEST phase (or frozen).  src/caffe/layers/bn_layer.cpp:93-207:          */
/*   t = x - mean_run ; t *= pow(var_run + eps, -0.5) ; t *= slope ; t += bias */
/* in that order.  The 4-D code normalises per channel over num*H*W; for 5-D   */
/* blobs only CuDNNBNLayer works in the reference (cudnn_bn_layer.cpp:37-44,    */
/* CUDNN_BATCHNORM_SPATIAL) and it is the same formula with spatial = D*H*W     */
/*  -- restated here for any spatial size.                        */
void ref_bn_forward_test(const float* x, float* y, const float* slope, const float* bias,
                         const float* mean, const float* var, float eps, int num, int C,
                         long spatial) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < num; ++n)
    for (int c = 0; c < C; ++c) {
      const float inv_std = powf(var[c] + eps, -0.5f);
      const float m = mean[c], g = slope[c], bb = bias[c];
      const float* xp = x + ((long)n * C + c) * spatial;
      float* yp = y + ((long)n * C + c) * spatial;
      for (long i = 0; i < spatial; ++i) {
        float t = xp[i] - m;
        t *= inv_std;
        t *= g;
        t += bb;
        yp[i] = t;
      }
    }
}

/* BN, TRAIN phase.  src/caffe/layers/bn_layer.cpp:107-157: batch mean, biased  */
/* batch variance over num*spatial, running <- (1-momentum)*batch + momentum*run */
/* (momentum 0.9, caffe.proto:469).  Outputs batch stats for the caller.         */
void ref_bn_forward_train(const float* x, float* y, const float* slope, const float* bias,
                          float* run_mean, float* run_var, float momentum, float eps,
                          int num, int C, long spatial, float* batch_mean, float* batch_var) {
#pragma omp parallel for schedule(static)
  for (int c = 0; c < C; ++c) {
    double s = 0.0;
    for (int n = 0; n < num; ++n) {
      const float* xp = x + ((long)n * C + c) * spatial;
      for (long i = 0; i < spatial; ++i) s += xp[i];
    }
    const float m = (float)(s / ((double)num * (double)spatial));
    double v = 0.0;
    for (int n = 0; n < num; ++n) {
      const float* xp = x + ((long)n * C + c) * spatial;
      for (long i = 0; i < spatial; ++i) { double d = xp[i] - m; v += d * d; }
    }
    const float var = (float)(v / ((double)num * (double)spatial));
    if (batch_mean) batch_mean[c] = m;
    if (batch_var) batch_var[c] = var;
    run_mean[c] = (1.0f - momentum) * m + momentum * run_mean[c];
    run_var[c] = (1.0f - momentum) * var + momentum * run_var[c];
    const float inv_std = powf(var + eps, -0.5f);
    for (int n = 0; n < num; ++n) {
      const float* xp = x + ((long)n * C + c) * spatial;
      float* yp = y + ((long)n * C + c) * spatial;
      for (long i = 0; i < spatial; ++i) yp[i] = (xp[i] - m) * inv_std * slope[c] + bias[c];
    }
  }
}

/* ReLU: src/caffe/layers/relu_layer.cpp:10-20  y = max(x,0) + slope*min(x,0)   */
void ref_relu(const float* x, float* y, long n, float negative_slope) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    const float v = x[i];
    y[i] = (v > 0.0f ? v : 0.0f) + negative_slope * (v < 0.0f ? v : 0.0f);
  }
}

/* ------------------------------------------------------------------------ */
/* Pooling forward (MAX=0, AVE=1), 1..3 spatial axes.                          */
/* 2-D: src/caffe/layers/pooling_layer.cpp:168-277.                             */
/*   MAX: window [o*s-p, min(start+k, in)), then start=max(start,0); init       */
/*        -FLT_MAX; strictly-greater update (first max wins).                   */
/*   AVE: pool_size = prod(min(start+k, in+p) - start) computed BEFORE the      */
/*        window is clipped to the image (pad-inclusive divisor); sum over the   */
/*        clipped window; divide.                                                */
/* N-D: the reference only runs these through cuDNN (cudnn_pooling_layer.cpp,    */
/*   util/cudnn.hpp:235-262: CUDNN_POOLING_MAX /                                 */
/*   CUDNN_POOLING_AVERAGE_COUNT_INCLUDE_PADDING); the per-axis rule above is     */
/*   the same arithmetic extended to a third axis (golden: test_pooling_layer     */
/*   .cpp:1559-1613, 1450-1525).                                                  */
int ref_pool_forward(const float* x, float* y, int num, int C, int nsp,
                     const int* in_shape, const int* kernel, const int* stride,
                     const int* pad, int method) {
  int k[3] = {1, 1, 1}, s[3] = {1, 1, 1}, p[3] = {0, 0, 0}, is[3] = {1, 1, 1}, os[3] = {1, 1, 1};
  if (nsp < 1 || nsp > 3) return -1;
  for (int i = 0; i < nsp; ++i) {
    int j = 3 - nsp + i;
    k[j] = kernel[i]; s[j] = stride[i]; p[j] = pad[i]; is[j] = in_shape[i];
    os[j] = ref_pool_out_dim(in_shape[i], kernel[i], stride[i], pad[i]);
  }
  const long in_sp = (long)is[0] * is[1] * is[2], out_sp = (long)os[0] * os[1] * os[2];
#pragma omp parallel for schedule(static)
  for (long nc = 0; nc < (long)num * C; ++nc) {
    const float* xp = x + nc * in_sp;
    float* yp = y + nc * out_sp;
    for (int oz = 0; oz < os[0]; ++oz)
      for (int oy = 0; oy < os[1]; ++oy)
        for (int ox = 0; ox < os[2]; ++ox) {
          int z0 = oz * s[0] - p[0], y0 = oy * s[1] - p[1], x0 = ox * s[2] - p[2];
          float r;
          if (method == 0) {
            int z1 = z0 + k[0] < is[0] ? z0 + k[0] : is[0];
            int y1 = y0 + k[1] < is[1] ? y0 + k[1] : is[1];
            int x1 = x0 + k[2] < is[2] ? x0 + k[2] : is[2];
            if (z0 < 0) z0 = 0; if (y0 < 0) y0 = 0; if (x0 < 0) x0 = 0;
            r = -FLT_MAX;
            for (int z = z0; z < z1; ++z)
              for (int yy = y0; yy < y1; ++yy)
                for (int xx = x0; xx < x1; ++xx) {
                  float v = xp[((long)z * is[1] + yy) * is[2] + xx];
                  if (v > r) r = v;
                }
          } else {
            int z1 = z0 + k[0] < is[0] + p[0] ? z0 + k[0] : is[0] + p[0];
            int y1 = y0 + k[1] < is[1] + p[1] ? y0 + k[1] : is[1] + p[1];
            int x1 = x0 + k[2] < is[2] + p[2] ? x0 + k[2] : is[2] + p[2];
            const int pool_size = (z1 - z0) * (y1 - y0) * (x1 - x0);
            if (z0 < 0) z0 = 0; if (y0 < 0) y0 = 0; if (x0 < 0) x0 = 0;
            if (z1 > is[0]) z1 = is[0]; if (y1 > is[1]) y1 = is[1]; if (x1 > is[2]) x1 = is[2];
            r = 0.0f;
            for (int z = z0; z < z1; ++z)
              for (int yy = y0; yy < y1; ++yy)
                for (int xx = x0; xx < x1; ++xx)
                  r += xp[((long)z * is[1] + yy) * is[2] + xx];
            r /= (float)pool_size;
          }
          yp[((long)oz * os[1] + oy) * os[2] + ox] = r;
        }
  }
  return 0;
}

/* ------------------------------------------------------------------------ */
/* Convolution backward.  src/caffe/layers/conv_layer.cpp:44-75 (loop over     */
/* images, bias / weight / bottom gradients) over                               */
/* base_conv_layer.cpp:290-328: weight_cpu_gemm  dW[Cout,K] += dY[Cout,M] x col^T[M,K]   */
/*                              backward_cpu_gemm dcol[K,M] = W^T[K,Cout] x dY[Cout,M];   */
/*                              col2im (util/im2col.cpp: scatter-add of dcol)   */
/*                              backward_cpu_bias db[o] += sum_m dY[o,m].        */
/* Parameter gradients ACCUMULATE into dw / db (caffe's beta = 1); dx is overwritten. */
static void col2im_nd(const float* col, int cin, int nsp, const int* in_shape, const int* out_shape,
                      const int* kernel, const int* stride, const int* pad, float* im) {
  int k[REF_MAX_SP] = {1, 1, 1}, s[REF_MAX_SP] = {1, 1, 1}, p[REF_MAX_SP] = {0, 0, 0};
  int is[REF_MAX_SP] = {1, 1, 1}, os[REF_MAX_SP] = {1, 1, 1};
  for (int i = 0; i < nsp; ++i) {
    int j = REF_MAX_SP - nsp + i;
    k[j] = kernel[i]; s[j] = stride[i]; p[j] = pad[i];
    is[j] = in_shape[i]; os[j] = out_shape[i];
  }
  const long in_sp = (long)is[0] * is[1] * is[2];
  const long out_sp = (long)os[0] * os[1] * os[2];
  const int ktaps = k[0] * k[1] * k[2];
#pragma omp parallel for schedule(static)
  for (int c = 0; c < cin; ++c) {  /* one channel per thread: scatter-adds of different taps never race */
    float* dst = im + (long)c * in_sp;
    memset(dst, 0, sizeof(float) * (size_t)in_sp);
    for (int t = 0; t < ktaps; ++t) {
      int tt = t;
      const int kx = tt % k[2]; tt /= k[2];
      const int ky = tt % k[1]; tt /= k[1];
      const int kz = tt;
      const float* src = col + ((long)c * ktaps + t) * out_sp;
      for (int oz = 0; oz < os[0]; ++oz) {
        const int iz = oz * s[0] - p[0] + kz;
        if (iz < 0 || iz >= is[0]) continue;
        for (int oy = 0; oy < os[1]; ++oy) {
          const int iy = oy * s[1] - p[1] + ky;
          if (iy < 0 || iy >= is[1]) continue;
          const float* sr = src + ((long)oz * os[1] + oy) * os[2];
          float* drow = dst + ((long)iz * is[1] + iy) * is[2];
          for (int ox = 0; ox < os[2]; ++ox) {
            const int ix = ox * s[2] - p[2] + kx;
            if (ix >= 0 && ix < is[2]) drow[ix] += sr[ox];
          }
        }
      }
    }
  }
}

static void transpose2d(const float* a, float* b, long rows, long cols) {  /* b[cols][rows] = a[rows][cols]^T */
#pragma omp parallel for schedule(static)
  for (long r0 = 0; r0 < rows; r0 += 32)
    for (long c0 = 0; c0 < cols; c0 += 32) {
      const long r1 = r0 + 32 < rows ? r0 + 32 : rows, c1 = c0 + 32 < cols ? c0 + 32 : cols;
      for (long r = r0; r < r1; ++r)
        for (long c = c0; c < c1; ++c) b[c * rows + r] = a[r * cols + c];
    }
}

int ref_conv_backward(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db,
                      int num, int cin, int cout, int nsp, const int* in_shape,
                      const int* kernel, const int* stride, const int* pad) {
  if (nsp < 1 || nsp > REF_MAX_SP) return -1;
  int out_shape[REF_MAX_SP];
  long in_sp = 1, out_sp = 1, ktaps = 1;
  for (int i = 0; i < nsp; ++i) {
    out_shape[i] = ref_conv_out_dim(in_shape[i], kernel[i], stride[i], pad[i]);
    if (out_shape[i] <= 0) return -2;
    in_sp *= in_shape[i]; out_sp *= out_shape[i]; ktaps *= kernel[i];
  }
  const long K = (long)cin * ktaps;
  float* col = (float*)malloc(sizeof(float) * (size_t)(K * out_sp));
  float* colT = dw ? (float*)malloc(sizeof(float) * (size_t)(K * out_sp)) : NULL;
  float* wT = dx ? (float*)malloc(sizeof(float) * (size_t)(K * cout)) : NULL;
  float* dwn = dw ? (float*)malloc(sizeof(float) * (size_t)(K * cout)) : NULL;
  if (!col || (dw && (!colT || !dwn)) || (dx && !wT)) { free(col); free(colT); free(wT); free(dwn); return -3; }
  if (dx) transpose2d(w, wT, cout, K);
  for (int n = 0; n < num; ++n) {
    const float* xn = x + (long)n * cin * in_sp;
    const float* dyn = dy + (long)n * cout * out_sp;
    if (db) {
#pragma omp parallel for schedule(static)
      for (int o = 0; o < cout; ++o) {
        double sacc = 0.0;
        const float* d = dyn + (long)o * out_sp;
        for (long i = 0; i < out_sp; ++i) sacc += d[i];
        db[o] += (float)sacc;
      }
    }
    if (dw) {
      im2col_nd(xn, cin, nsp, in_shape, out_shape, kernel, stride, pad, col);
      transpose2d(col, colT, K, out_sp);                 /* colT[M][K] */
      ref_sgemm(dyn, colT, dwn, cout, (int)K, (int)out_sp);  /* [Cout,M] x [M,K] */
#pragma omp parallel for schedule(static)
      for (long i = 0; i < K * cout; ++i) dw[i] += dwn[i];
    }
    if (dx) {
      ref_sgemm(wT, dyn, col, (int)K, (int)out_sp, cout);   /* dcol[K,M] = W^T x dY */
      col2im_nd(col, cin, nsp, in_shape, out_shape, kernel, stride, pad, dx + (long)n * cin * in_sp);
    }
  }
  free(col); free(colT); free(wT); free(dwn);
  return 0;
}

/* Pooling backward.  src/caffe/layers/pooling_layer.cpp:280-377.               */
/*   MAX: the forward pass records the index of the FIRST maximum in scan order  */
/*        (strictly-greater update, :206-222); backward adds top_diff there.     */
/*   AVE: top_diff / pool_size (pad-inclusive divisor) added to every in-image    */
/*        element of the window.                                                  */
int ref_pool_backward(const float* x, const float* dy, float* dx, int num, int C, int nsp,
                      const int* in_shape, const int* kernel, const int* stride,
                      const int* pad, int method) {
  int k[3] = {1, 1, 1}, s[3] = {1, 1, 1}, p[3] = {0, 0, 0}, is[3] = {1, 1, 1}, os[3] = {1, 1, 1};
  if (nsp < 1 || nsp > 3) return -1;
  for (int i = 0; i < nsp; ++i) {
    int j = 3 - nsp + i;
    k[j] = kernel[i]; s[j] = stride[i]; p[j] = pad[i]; is[j] = in_shape[i];
    os[j] = ref_pool_out_dim(in_shape[i], kernel[i], stride[i], pad[i]);
  }
  const long in_sp = (long)is[0] * is[1] * is[2], out_sp = (long)os[0] * os[1] * os[2];
#pragma omp parallel for schedule(static)
  for (long nc = 0; nc < (long)num * C; ++nc) {
    const float* xp = x + nc * in_sp;
    const float* dyp = dy + nc * out_sp;
    float* dxp = dx + nc * in_sp;
    memset(dxp, 0, sizeof(float) * (size_t)in_sp);
    for (int oz = 0; oz < os[0]; ++oz)
      for (int oy = 0; oy < os[1]; ++oy)
        for (int ox = 0; ox < os[2]; ++ox) {
          int z0 = oz * s[0] - p[0], y0 = oy * s[1] - p[1], x0 = ox * s[2] - p[2];
          const float g = dyp[((long)oz * os[1] + oy) * os[2] + ox];
          if (method == 0) {
            int z1 = z0 + k[0] < is[0] ? z0 + k[0] : is[0];
            int y1 = y0 + k[1] < is[1] ? y0 + k[1] : is[1];
            int x1 = x0 + k[2] < is[2] ? x0 + k[2] : is[2];
            if (z0 < 0) z0 = 0; if (y0 < 0) y0 = 0; if (x0 < 0) x0 = 0;
            float r = -FLT_MAX;
            long best = -1;
            for (int z = z0; z < z1; ++z)
              for (int yy = y0; yy < y1; ++yy)
                for (int xx = x0; xx < x1; ++xx) {
                  const long idx = ((long)z * is[1] + yy) * is[2] + xx;
                  if (xp[idx] > r) { r = xp[idx]; best = idx; }
                }
            if (best >= 0) dxp[best] += g;
          } else {
            int z1 = z0 + k[0] < is[0] + p[0] ? z0 + k[0] : is[0] + p[0];
            int y1 = y0 + k[1] < is[1] + p[1] ? y0 + k[1] : is[1] + p[1];
            int x1 = x0 + k[2] < is[2] + p[2] ? x0 + k[2] : is[2] + p[2];
            const int pool_size = (z1 - z0) * (y1 - y0) * (x1 - x0);
            if (z0 < 0) z0 = 0; if (y0 < 0) y0 = 0; if (x0 < 0) x0 = 0;
            if (z1 > is[0]) z1 = is[0]; if (y1 > is[1]) y1 = is[1]; if (x1 > is[2]) x1 = is[2];
            for (int z = z0; z < z1; ++z)
              for (int yy = y0; yy < y1; ++yy)
                for (int xx = x0; xx < x1; ++xx) dxp[((long)z * is[1] + yy) * is[2] + xx] += g / (float)pool_size;
          }
        }
  }
  return 0;
}

/* BN backward, TRAIN phase, not frozen.  src/caffe/layers/bn_layer.cpp:241-335:  */
/*   dslope += sum x_norm*dy ; dbias += sum dy ;                                  */
/*   dx = inv_std * ( slope*dy - mean(slope*dy) - x_norm * mean(x_norm * slope*dy) )  */
/* x_norm = (x - batch_mean) * inv_std as saved by the forward pass (:183-188).    */
void ref_bn_backward_train(const float* x, const float* dy, float* dx, const float* slope,
                           const float* batch_mean, const float* batch_var, float eps,
                           int num, int C, long spatial, float* dslope, float* dbias) {
#pragma omp parallel for schedule(static)
  for (int c = 0; c < C; ++c) {
    const float m = batch_mean[c];
    const float inv_std = powf(batch_var[c] + eps, -0.5f);
    double s_dy = 0.0, s_dyx = 0.0;
    for (int n = 0; n < num; ++n) {
      const float* xp = x + ((long)n * C + c) * spatial;
      const float* dp = dy + ((long)n * C + c) * spatial;
      for (long i = 0; i < spatial; ++i) {
        const float xn = (xp[i] - m) * inv_std;
        s_dy += dp[i];
        s_dyx += (double)dp[i] * xn;
      }
    }
    if (dslope) dslope[c] += (float)s_dyx;
    if (dbias) dbias[c] += (float)s_dy;
    if (!dx) continue;
    const double cnt = (double)num * (double)spatial;
    const float mean_g = (float)(slope[c] * s_dy / cnt), mean_gx = (float)(slope[c] * s_dyx / cnt);
    for (int n = 0; n < num; ++n) {
      const float* xp = x + ((long)n * C + c) * spatial;
      const float* dp = dy + ((long)n * C + c) * spatial;
      float* op = dx + ((long)n * C + c) * spatial;
      for (long i = 0; i < spatial; ++i) {
        const float xn = (xp[i] - m) * inv_std;
        op[i] = (slope[c] * dp[i] - mean_g - xn * mean_gx) * inv_std;
      }
    }
  }
}

/* Eltwise SUM with coefficients: src/caffe/layers/eltwise_layer.cpp:66-72       */
void ref_eltwise_sum(const float* a, const float* b, float ca, float cb, float* y, long n) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) y[i] = ca * a[i] + cb * b[i];
}

/* Permute: src/caffe/layers/permute_layer.cpp:9-26,75-94                        */
/* top.shape[i] = bottom.shape[order[i]]; gather through old/new strides.         */
int ref_permute(const float* x, float* y, int naxes, const int* shape, const int* order) {
  if (naxes < 1 || naxes > 8) return -1;
  long old_steps[8], new_steps[8];
  int new_shape[8];
  long count = 1;
  for (int i = naxes - 1; i >= 0; --i) { old_steps[i] = count; count *= shape[i]; }
  for (int i = 0; i < naxes; ++i) new_shape[i] = shape[order[i]];
  long c2 = 1;
  for (int i = naxes - 1; i >= 0; --i) { new_steps[i] = c2; c2 *= new_shape[i]; }
#pragma omp parallel for schedule(static)
  for (long idx = 0; idx < count; ++idx) {
    long rem = idx, old_idx = 0;
    for (int j = 0; j < naxes; ++j) {
      const long q = rem / new_steps[j];
      rem -= q * new_steps[j];
      old_idx += q * old_steps[order[j]];
    }
    y[idx] = x[old_idx];
  }
  return 0;
}

/* InnerProduct: src/caffe/layers/inner_product_layer.cpp:80-93                   */
/*   y[M,N] = x[M,K] * W[N,K]^T + bias[N]                                         */
void ref_inner_product(const float* x, const float* w, const float* b, float* y,
                       int M, int N, int K) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.0f;
      const float* xp = x + (long)m * K;
      const float* wp = w + (long)n * K;
      for (int kk = 0; kk < K; ++kk) acc += xp[kk] * wp[kk];
      y[(long)m * N + n] = acc + (b ? b[n] : 0.0f);
    }
}

/* Softmax over axis 1 of [M,N]: src/caffe/layers/softmax_layer.cpp (max-subtract,*/
/* exp, normalise).                                                               */
void ref_softmax(const float* x, float* y, int M, int N) {
  for (int m = 0; m < M; ++m) {
    const float* xp = x + (long)m * N;
    float* yp = y + (long)m * N;
    float mx = xp[0];
    for (int n = 1; n < N; ++n) if (xp[n] > mx) mx = xp[n];
    float s = 0.0f;
    for (int n = 0; n < N; ++n) { yp[n] = expf(xp[n] - mx); s += yp[n]; }
    for (int n = 0; n < N; ++n) yp[n] /= s;
  }
}

/* Round-to-nearest-even fp32 -> bf16 -> fp32, used by the bf16-emulating mode of */
/* oracle/refnet.py so the oracle rounds at the same points as the device path.   */
void ref_round_bf16(const float* x, float* y, long n) {
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n; ++i) {
    uint32_t u;
    memcpy(&u, x + i, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) {
      const uint32_t r = ((u >> 16) & 1u) + 0x7FFFu;
      u = (u + r) & 0xFFFF0000u;
    } else {
      u &= 0xFFFF0000u;
    }
    memcpy(y + i, &u, 4);
  }
}
