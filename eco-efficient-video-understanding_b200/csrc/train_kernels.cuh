// train_kernels.cuh -- the HBM-bound kernels of the training path (TRAIN-phase forward pieces and the backward
// counterparts of everything that is not a convolution GEMM).  Reference semantics per kernel in train_kernels.cu.
// Feature-map gradients are bf16 channels-last views (same layout as the data they belong to, ClView);
// parameter gradients, statistics and the small vectors behind the global pool are fp32.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "aux_kernels.cuh"

namespace eco {

// ---- per-channel reductions over a channels-last tensor (rows = outer*inner, C % 8 == 0) ----
// Deterministic: per-block partials are written to `scratch` (kColReduceScratchFloats floats) and summed in block order.
constexpr int kColReduceMaxBlocks = 592;   // 4 per SM (the two-tensor BN-backward sums use 2 per SM: 96 registers)
constexpr int kColReduceMaxC = 2048;
constexpr size_t kColReduceScratchFloats = (size_t)kColReduceMaxBlocks * kColReduceMaxC * 2 + kColReduceMaxBlocks;
// out[0..C) += sum_rows x                                   (BN mean numerator, conv bias gradient)
cudaError_t launch_colsum_cl(ClView x, float* out, float* scratch, cudaStream_t st);
// out[0..C) += sum_rows (x - mean[c])^2                      (BN biased variance numerator, bn_layer.cpp:141-151)
cudaError_t launch_colsqdev_cl(ClView x, const float* mean, float* out, float* scratch, cudaStream_t st);
// BN backward sums (bn_layer.cpp:241-262): with g = dy * (y > 0 if relu; the mask is recomputed from x, y is not read) and
// xn = (x - mean) * inv_std
//   out[0..C) += sum g ;  out[C..2C) += sum g * xn
cudaError_t launch_bn_bwd_sums_cl(ClView x, ClView y, ClView dy, const float* mean, const float* inv_std, const float* slope,
                                  const float* bias, int relu, float* out, float* scratch, cudaStream_t st);

// BN TRAIN statistics (bn_layer.cpp:107-157): from sum / sqdev numerators to mean, biased variance, inverse std and the
// running-average update  running = (1 - m) * batch + m * running
// one-pass variant: Welford per thread + Chan's combination in a fixed order (deterministic), statistics finished in place
cudaError_t launch_bn_stats_cl(ClView x, float* mean, float* inv_std, float* batch_var, float* run_mean, float* run_var,
                               float momentum, float eps, float* scratch, cudaStream_t st);
cudaError_t launch_bn_finish_mean(const float* sum, float* mean, int C, double count, cudaStream_t st);
cudaError_t launch_bn_finish_var(const float* sqdev, const float* mean, float* inv_std, float* batch_var, float* run_mean,
                                 float* run_var, int C, double count, float momentum, float eps, cudaStream_t st);
// y = relu?((x - mean) * inv_std * slope + bias)
cudaError_t launch_bn_apply_cl(ClView x, ClView y, const float* mean, const float* inv_std, const float* slope,
                               const float* bias, int relu, cudaStream_t st);
// dx (+)= inv_std * (slope * g - slope * S1 / n - xn * slope * S2 / n)      (bn_layer.cpp:264-334);
// dslope += S2, dbias += S1 (the gradient arena accumulates like caffe's blobs)
cudaError_t launch_bn_bwd_apply_cl(ClView x, ClView y, ClView dy, ClView dx, const float* mean, const float* inv_std,
                                   const float* slope, const float* bias, const float* sums, double count, int relu, int accumulate,
                                   float* dslope, float* dbias, cudaStream_t st);

// ---- pooling backward (pooling_layer.cpp:280-377), any 1..3-D window, channels-last bf16 ----
// MAX: the gradient goes to the FIRST maximum of each window in scan order (the forward pass's strictly-greater update);
// AVE: dy / pool_size (pad-inclusive divisor) to every in-image element.  Gather formulation: no atomics, deterministic.
// `mask`: scratch of NB*OD*OH*OW*C bytes for the MAX path (first-maximum index per window and channel); NULL falls
// back to re-scanning the windows
cudaError_t launch_pool_bwd_cl(const PoolParams& p, const __nv_bfloat16* dy, long long dy_cs, int dy_coff,
                               __nv_bfloat16* dx, long long dx_cs, int dx_coff, int accumulate, unsigned char* mask,
                               cudaStream_t st);
// generic fp32 pooling on plain blobs (segment consensus of ECO-Full), same rules
cudaError_t launch_pool_f32_bwd(const PoolF32Params& p, const float* dy, float* dx, int accumulate, cudaStream_t st);
// full-extent average pool: dx[o, i, c] (+)= dy[o, c] / inner
cudaError_t launch_global_avg_bwd_cl(const float* dy, ClView dx, int accumulate, cudaStream_t st);

// y (+)= x on channels-last views of equal logical shape (Eltwise / Split / Concat-copy gradients)
cudaError_t launch_cl_axpy(ClView x, ClView y, int accumulate, cudaStream_t st);
// y[i] (+)= a * x[i], fp32
cudaError_t launch_f32_axpy(const float* x, float* y, long long n, float a, int accumulate, cudaStream_t st);
// scatter a strided-convolution output gradient into its zero-dilated form (zeros are written once at plan time)
// dX[n, z*sD, y*sH, x*sW, :] (+)= src[n, z, y, x, :], every other position of dX = 0 unless `accumulate` (then untouched):
// the input gradient of a 1x1 strided convolution from its compact form
cudaError_t launch_scatter_stride_cl(const __nv_bfloat16* src, int C, int OD, int OH, int OW, ClView dx, int ID, int IH, int IW,
                                     int sD, int sH, int sW, int accumulate, cudaStream_t st);
cudaError_t launch_dilate_cl(ClView src, int OD, int OH, int OW, __nv_bfloat16* dst, int ED, int EH, int EW, int sD, int sH,
                             int sW, cudaStream_t st);

// ---- Dropout TRAIN (dropout_layer.cpp:33-49): Bernoulli(1-p) mask from a counter-based hash, survivors * 1/(1-p) ----
cudaError_t launch_dropout_f32(const float* x, float* y, long long n, float ratio, uint64_t seed, cudaStream_t st);

// ---- InnerProduct backward (inner_product_layer.cpp:96-120), fp32 ----
cudaError_t launch_inner_product_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* db, int M,
                                     int N, int K, int accumulate_dx, cudaStream_t st);

// ---- SoftmaxWithLoss (softmax_loss_layer.cpp:48-120) and Accuracy (accuracy_layer.cpp:50-92) ----
// prob[M,N] = softmax(x); *loss = -sum_m log(max(prob[m, label[m]], FLT_MIN)) / M
cudaError_t launch_softmax_loss_fwd(const float* x, const float* label, float* prob, float* loss, int M, int N, cudaStream_t st);
// dx = (prob - onehot(label)) * loss_weight / M
cudaError_t launch_softmax_loss_bwd(const float* prob, const float* label, float* dx, int M, int N, float loss_weight,
                                    cudaStream_t st);
// *acc = fraction of rows whose label is among the top_k scores (ties: the larger index ranks first, std::greater on pairs)
cudaError_t launch_accuracy(const float* x, const float* label, float* acc, int M, int N, int top_k, cudaStream_t st);

// ---- weights: fp32 master copy (caffe layout [Cout][Cin][taps]) -> bf16 GEMM operands ----
// forward operand [Cout_pad][taps * cblocks * 64], k = (tap * cblocks + c / 64) * 64 + c % 64
cudaError_t launch_pack_conv_w(const float* w, __nv_bfloat16* dst, int Cout, int Cin, int taps, int cblocks, long long Ktotal,
                               cudaStream_t st);
// the 7x7/s2 stem as a 4x4 filter over space-to-depth cells: k = ty * 64 + tx * 16 + (dy * 2 + dx) * 3 + ch
cudaError_t launch_pack_stem_w(const float* w, __nv_bfloat16* dst, int Cout, long long Ktotal, cudaStream_t st);
// dgrad operand (the transposed, spatially flipped filter): rows = Cin, k = (flip(tap) * oblocks + o / 64) * 64 + o % 64
cudaError_t launch_pack_conv_w_dgrad(const float* w, __nv_bfloat16* dst, int Cout, int Cin, int KD, int KH, int KW, int oblocks,
                                     long long Ktotal, cudaStream_t st);
// gradient arena (caffe layout) += wgrad scratch [taps][cin_ld][cout_ld]
cudaError_t launch_wgrad_finish(const float* scratch, float* dw, int Cout, int Cin, int taps, int cin_ld, int cout_ld,
                                cudaStream_t st);
// stem: scratch holds the 4x4-over-cells filter gradient [16 taps][64 cell values][cout_ld] -> 7x7x3 caffe layout
cudaError_t launch_wgrad_finish_stem(const float* scratch, float* dw, int Cout, int cout_ld, cudaStream_t st);

// ---- solver (solver.cpp:637-797): global-norm clip, L2 decay, SGD / Nesterov momentum update ----
cudaError_t launch_sumsq(const float* x, long long n, float* out /* += */, cudaStream_t st);
cudaError_t launch_scale(float* x, long long n, const float* scale_dev /* x *= *scale_dev */, cudaStream_t st);
// diff = diff * norm + decay * w ; h_new = momentum * h + rate * diff ;
// SGD: w -= h_new ; Nesterov (solver.cpp NesterovSolver::ComputeUpdateValue): w -= (1 + momentum) * h_new - momentum * h
cudaError_t launch_sgd_update(float* w, float* diff, float* hist, long long n, float rate, float momentum, float decay, float norm,
                              const float* clip_scale_dev, int nesterov, cudaStream_t st);
// clip factor on the device (no host sync): *scale = min(1, clip / sqrt(*sumsq)) on the accumulated diffs (before Normalize)
cudaError_t launch_clip_factor(const float* sumsq, float clip, float norm, float* scale, cudaStream_t st);

}  // namespace eco
