// aux_kernels.cuh -- the HBM-bound kernels around the convolutions: layout transforms at the
// fp32-NCHW boundary, pooling (caffe semantics), global pooling, inner product, fallbacks.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace eco {

// A channels-last bf16 tensor seen as [outer][inner][C] with a channel stride / offset, i.e. the
// physical form of a caffe blob [outer-dims..., C, inner-dims...].
struct ClView {
  __nv_bfloat16* ptr;
  long long outer, inner;
  int C;             // logical channels
  long long cs;      // channel stride of the underlying buffer (>= C)
  int coff;          // first channel inside the buffer
  // split-precision storage (planner option `precision`): three planes [hi | lo | hi] per pixel, `seg` elements apart
  // (= cs), pixel stride 3 * cs; value = hi + lo.  0: plain bf16, pixel stride cs.
  long long seg = 0;
  __host__ __device__ long long ps() const { return seg ? 3 * cs : cs; }
};

// fp32 logical (caffe row-major [outer, C, inner]) -> channels-last bf16, and back.
cudaError_t launch_f32_to_cl(const float* src, ClView dst, cudaStream_t st);
cudaError_t launch_cl_to_f32(ClView src, float* dst, cudaStream_t st);

// Stem input transform: fp32 [F,3,H,W] -> bf16 space-to-depth cells [F, CH, CW, 16]
// (cell (Y,X) channel (dy*2+dx)*3+c = x[c][2Y+dy-3][2X+dx-3], zero outside / channels 12..15).
cudaError_t launch_stem_s2d(const float* src, __nv_bfloat16* dst, int F, int H, int W, int CH, int CW,
                            cudaStream_t st);
// the same from raw uint8 frames [F,3,H,W] with the per-channel mean subtracted on the fly
// (what DataTransformer::Transform does on the host in the reference, data_transformer.cpp:50-325: mean_value, no scale)
cudaError_t launch_stem_s2d_u8(const unsigned char* src, __nv_bfloat16* dst, int F, int H, int W, int CH, int CW,
                               float mean0, float mean1, float mean2, cudaStream_t st);
// uint8 [outer, C, inner] -> fp32 minus per-channel mean (generic fallback for non-stem inputs, C <= 4)
cudaError_t launch_u8_to_f32_mean(const unsigned char* src, float* dst, long long outer, int C, long long inner,
                                  float mean0, float mean1, float mean2, float mean3, cudaStream_t st);

struct PoolParams {
  const __nv_bfloat16* x; long long x_cs; int x_coff;
  __nv_bfloat16* y; long long y_cs; int y_coff;
  int NB, C;
  int ID, IH, IW, OD, OH, OW;
  int KD, KH, KW, sD, sH, sW, pD, pH, pW;
  int is_max;
  // optional per-channel epilogue y = relu?((pool + bias) * scale + shift), AVE 3x3 on the row-staged kernel only:
  // used when a 1x1 convolution behind an AVE pooling was moved in front of it (both are linear)
  const float* bias; const float* scale; const float* shift; int relu;
};
cudaError_t aux_kernels_configure();  // max dynamic shared memory attributes, once per device
// the row-staged kernel is the only one with the affine epilogue: 3 input rows of a dense map must fit its 200 KB
inline bool pool_cl_affine_supported(int IW, int C, long long NB) {
  return (size_t)3 * IW * C * 2 <= (size_t)200 * 1024 && NB <= 65535 && C % 8 == 0;
}
// caffe pooling on channels-last bf16 (pooling_layer.cpp:199-262 semantics), C % 8 == 0
cudaError_t launch_pool_cl(const PoolParams& p, cudaStream_t st);

// generic N-D pooling on split-precision maps (x/y views carry seg != 0): pooled in fp32 from hi + lo, stored split
cudaError_t launch_pool_cl_split(const PoolParams& p, long long x_seg, long long y_seg, cudaStream_t st);
// mean over `inner` of a channels-last tensor -> fp32 [outer, C]  (global_pool / global_pool2D)
cudaError_t launch_global_avg_cl(ClView src, float* dst, cudaStream_t st);

// generic fp32 NC(D)HW pooling for the small plain blobs (segment consensus)
struct PoolF32Params {
  const float* x; float* y;
  int NC;  // num * channels
  int ID, IH, IW, OD, OH, OW, KD, KH, KW, sD, sH, sW, pD, pH, pW, is_max;
};
cudaError_t launch_pool_f32(const PoolF32Params& p, cudaStream_t st);

// y[M,N] = x[M,K] * W[N,K]^T + b[N]   (inner_product_layer.cpp:80-93), fp32
cudaError_t launch_inner_product(const float* x, const float* w, const float* b, float* y, int M, int N, int K,
                                 cudaStream_t st);

// fallbacks for graphs the planner cannot fuse: y = relu?(x*scale+shift) and y = a + b, channels-last bf16
cudaError_t launch_scale_shift_relu_cl(ClView x, ClView y, const float* scale, const float* shift, int relu,
                                       cudaStream_t st);
cudaError_t launch_eltwise_sum_cl(ClView a, ClView b, ClView y, cudaStream_t st);

// softmax over axis 1 of fp32 [M,N]
cudaError_t launch_softmax_f32(const float* x, float* y, int M, int N, cudaStream_t st);

}  // namespace eco
