// wgrad_umma.cuh -- weight gradient of a convolution as an implicit GEMM on tcgen05 tensor cores (sm_100a).
//
//   dW[o][tap][c] = sum over output positions m of  dY[m][o] * X[im2col(m, tap)][c]
//
// Replaces weight_cpu_gemm / weight_gpu_gemm (caffe_3d/src/caffe/layers/base_conv_layer.cpp:305-320,:376-391: per image
// im2col into a column buffer, then SGEMM dY x col^T with beta = 1) for every Convolution's parameter gradient.
// GEMM view: M = Cout (128 per tile), N = Cin (64..128 per tile), K = output positions; one (tap, Cin tile, Cout tile)
// per accumulator, the positions split over several CTAs whose partial results are added into an fp32 scratch tensor
// with red.global.add (split-K).  Both operands arrive exactly as the forward pass reads them -- dY as a 2-D tile of
// the channels-last gradient map, X through the SAME im2col TMA descriptor the forward kernel uses -- which makes them
// "MN-major" operands (channels contiguous, positions strided): no transposed copy of either tensor ever exists.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eco {

struct WgradParams {
  // geometry of the forward convolution (positions enumerate its OUTPUT grid)
  int OD, OH, OW;
  int KD, KH, KW;
  int sD, sH, sW, pD, pH, pW;
  int nsp;               // 2 or 3: selects the 4-D / 5-D im2col TMA form
  int M;                 // output positions NB*OD*OH*OW (the GEMM's K)
  int Cout, Cin;
  int cout_tiles;        // ceil(Cout / 128)
  int cin_tiles;         // ceil(Cin / n_tile)
  int n_tile;            // 64 or 128 input channels per accumulator
  int splits;            // CTAs along the position axis
  int stages;            // smem ring depth
  int cin_ld, cout_ld;   // leading dimensions of the scratch tensor [taps][cin_ld][cout_ld]
  float* scratch;        // fp32, zeroed before the launch
  int* error_flag;
};

inline size_t wgrad_smem_bytes(const WgradParams& p) {
  return 1024 + (size_t)p.stages * (32768 + (size_t)(p.n_tile / 64) * 16384) + 64 * 8;
}
// tmY: 2-D tiled map over dY [M positions][Cout] (box 64 channels x 128 positions, SWIZZLE_128B)
// tmX: the forward convolution's im2col map over X (box 64 channels x 128 positions, SWIZZLE_128B)
cudaError_t launch_wgrad_umma(const WgradParams& p, const CUtensorMap& tmY, const CUtensorMap& tmX, cudaStream_t stream);
cudaError_t wgrad_umma_configure();

}  // namespace eco
