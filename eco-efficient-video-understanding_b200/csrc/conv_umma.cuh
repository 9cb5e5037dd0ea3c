// conv_umma.cuh -- parameters of the fused implicit-GEMM convolution kernel (sm_100a).
//
// One kernel family serves every Convolution on ECO's path (2-D 1x1/3x3/7x7-stem, 3-D 3x3x3,
// stride 1/2): GEMM M = output positions (N*D*H*W, channels-last), N = Cout, K = taps * Cin.
// Replaces the reference's per-image im2col + SGEMM (caffe_3d/src/caffe/layers/conv_layer.cpp:28-43,
// base_conv_layer.cpp:264-287, util/im2col.cu:12-161) and the BN / ReLU / Eltwise / Concat layers
// that follow it (bn_layer.cu:12-124, relu_layer.cu:10-27, eltwise_layer.cu:48-54,
// concat_layer.cu:10-27), which are folded into the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace eco {

constexpr int kBlockM = 128;      // UMMA M (cta_group::1), one TMEM lane per output position
constexpr int kBlockK = 64;       // bf16 elements per K block = one 128-byte swizzle row
constexpr int kUmmaK = 16;        // K per tcgen05.mma for 16-bit inputs
constexpr int kConvThreads = 192; // warp0: TMA producer, warp1: MMA issuer + TMEM owner, warps2-5: gather/epilogue

enum AMode : int {
  A_GATHER = 0,      // cp.async software im2col into the 128B-swizzled tile (any geometry)
  A_TMA_IM2COL = 1,  // one cp.async.bulk.tensor ... im2col per K block
};

struct ConvKernelParams {
  // ---- A operand: bf16 channels-last activations, element strides ----
  const __nv_bfloat16* x;
  long long x_sN, x_sD, x_sH, x_sW;  // strides in elements (channel stride is 1)
  int Cin;                           // channels per tap actually present (tail of a 64-block is zero-filled)
  int ID, IH, IW;                    // input extent (D = 1 for 2-D)
  int OD, OH, OW;                    // output extent
  int KD, KH, KW;
  int sD, sH, sW;
  int pD, pH, pW;
  int M;                             // NB * OD * OH * OW
  int cblocks;                       // ceil(Cin / 64)
  int num_kb;                        // KD*KH*KW*cblocks
  int a_mode;
  int nsp;                           // 2 or 3 spatial dims (selects the 4-D / 5-D TMA form)
  // ---- tile shape ----
  int block_n;                       // multiple of 16, <= 256
  int stages;
  int tmem_cols;                     // power of two >= block_n (x2 when persistent: double-buffered accumulator)
  int persistent;                    // 1: conv_umma_persistent_kernel (needs a_mode == A_TMA_IM2COL)
  int m_halves;                      // persistent only: 1 or 2 128-row halves per tile sharing the weight tile
  int epi_group;                     // persistent only: 16-column chunks staged per copy-out (2 or 4)
  int debug_flags;                   // development only: 1 = skip global stores, 2 = skip TMEM loads, 4 = skip MMAs, 8 = skip TMA loads
  int epi_staged;                    // persistent only: 1 = row-contiguous copy-out through smem, 0 = direct 32-byte stores
  int num_sms;
  int multicast;                     // persistent only: clusters of 2 CTAs share each weight tile (TMA multicast of halves)
  // ---- epilogue:  raw = acc + bias (+ res);  y = relu?(raw * scale + shift) ----
  int Cout;
  const float* bias;                 // [Cout] or null
  const float* scale;                // [Cout] or null (then y = raw)
  const float* shift;                // [Cout] (with scale)
  int relu;
  __nv_bfloat16* out;   long long out_cs;  int out_coff;   // y   (null -> not stored)
  __nv_bfloat16* raw;   long long raw_cs;  int raw_coff;   // raw (null -> not stored)
  const __nv_bfloat16* res; long long res_cs; int res_coff;  // residual added into raw (null -> none)
  int* error_flag;                   // set non-zero if an mbarrier wait times out
  // ---- split-precision mode (one-tile kernel only; planner option `precision`): every feature map is stored as three
  // bf16 planes [hi | lo | hi] per pixel (hi = bf16(v), lo = bf16(v - hi)), `*_seg` elements apart, so that a convolution
  // over the 3C "channels" with weights [w_hi | w_hi | w_lo] computes hi*w_hi + lo*w_hi + hi*w_lo in the fp32
  // accumulator: ~16 significant bits per operand instead of 8.  0 = plain bf16 storage. ----
  long long out_seg, raw_seg, res_seg;
  // ---- output segments (persistent kernel, out-only epilogue): sibling 1x1 convolutions that read the same
  // tensor run as ONE GEMM whose N range is the concatenation of their output channels; segment g covers
  // channels [seg_end[g-1], seg_end[g]) and goes to its own tensor (seg_coff already has the segment start
  // subtracted) with its own ReLU flag.  nseg <= 1: `out` / `relu` above. ----
  int nseg;
  int seg_end[4];
  __nv_bfloat16* seg_ptr[4];
  long long seg_cs[4];
  int seg_coff[4];
  int seg_relu[4];
};

// ---- halo-resident 2-D convolution (stride 1): the input patch of a tile is loaded ONCE and every
// filter tap reads it through a row-shifted UMMA descriptor (no per-tap im2col traffic) ----
struct HaloKernelParams {
  int NB, OH, OW;            // output grid
  int KH, KW, pH, pW;
  int pw;                    // patch width  = OW + KW - 1 (output positions are enumerated on this padded grid)
  int R;                     // output rows per tile
  int bands;                 // ceil(OH / R)
  int cblocks;               // ceil(Cin / 64)
  int block_n, Cout;
  int a_stages, b_stages;
  uint32_t a_stage_bytes;    // allocated bytes per A patch slot (patch + over-read slack, 1024-aligned)
  uint32_t a_tx_bytes;       // bytes one patch load delivers = 128 * pw * (R + KH - 1)
  int row_bytes;             // bytes per patch pixel: 128 (64 ch, SWIZZLE_128B) or 32 (16 ch, SWIZZLE_32B: the stem cells)
  int b_resident;            // 1: all weight tiles live in shared memory for the whole kernel (single N tile)
  int b_kblocks;             // number of 64-wide weight K blocks (= Ktotal / 64)
  int tmem_cols, num_sms;
  const float* bias; const float* scale; const float* shift;
  int relu;
  __nv_bfloat16* out;   long long out_cs;  int out_coff;
  __nv_bfloat16* raw;   long long raw_cs;  int raw_coff;
  const __nv_bfloat16* res; long long res_cs; int res_coff;
  int* error_flag;
};
inline size_t halo_smem_bytes(const HaloKernelParams& p) {
  const size_t b = p.b_resident ? (size_t)p.b_kblocks * p.block_n * 128 : (size_t)p.b_stages * p.block_n * 128;
  return 1024 + (size_t)p.a_stages * p.a_stage_bytes + b + 3 * 256 * sizeof(float) + 512;
}
cudaError_t launch_conv_halo(const HaloKernelParams& p, int m_halves, const CUtensorMap& tmX, const CUtensorMap& tmB,
                             cudaStream_t stream);


// ---- stem rows kernel: the 7x7/s2/p3 stem (Cin 3 -> 64) as a sliding window over rows of
// space-to-depth cells.  One 14 KB tile (a cell row: OW windows x 64 values) is loaded ONCE and feeds
// the four vertical filter taps of four different output rows, each with its own TMEM accumulator
// (ring of 8 x 64 columns); optionally the 3x3/s2 max pooling that follows (pool1) is applied to the
// rectified rows on chip, so the 64-channel full-resolution map never touches HBM. ----
struct StemRowsParams {
  int F;                     // frames
  int OH, OW;                // conv output grid (OW <= 128)
  int pool;                  // 1: emit MAX 3x3 / stride 2 / pad 0 (caffe ceil mode) of the conv output instead
  int PH, PW;                // pooled grid
  int strip, strips;         // output rows (pooled rows if pool) per work unit; units per frame
  int a_stages;
  uint32_t a_tx_bytes;       // OW * 128
  int num_sms;
  // source of the cell rows: 0 = bf16 cells written by the stem transform kernel (tiled TMA), 1 = the net's fp32
  // NCHW frames, 2 = raw uint8 NCHW frames with the per-channel mean subtracted: image rows are bulk-copied
  // into a shared-memory ring and four gather warps build the overlapping windows on chip, so the transform
  // kernel, its 217 MB of cells per step and their re-read disappear
  int src_mode;
  const void* src;           // frames (src_mode 1 / 2)
  int H, W;                  // frame size
  float mean0, mean1, mean2; // src_mode 2
  int gather_warps;          // 4..6 warps building the windows (src_mode 1 / 2)
  int raw_stages;            // ring of raw image-row pairs (src_mode 1 / 2)
  uint32_t raw_stage_bytes;  // 6 rows (2 image rows x 3 channels) x W x element size, 128-byte padded
  int debug_flags;           // development: 1 no stores, 2 no TMEM loads, 4 no MMAs, 8 no TMA loads, 16 print role wait cycles (CTA 0)
  const float* bias; const float* scale; const float* shift;
  int relu;
  __nv_bfloat16* out; long long out_cs; int out_coff;   // [F, OH, OW, 64] or pooled [F, PH, PW, 64]
  int* error_flag;
};
inline size_t stem_rows_smem_bytes(const StemRowsParams& p) {
  return 1024 + (size_t)p.a_stages * 16384 + 4 * 8192 + (p.pool ? 2 * 16384 : 0) +
         (p.src_mode ? (size_t)p.raw_stages * p.raw_stage_bytes : 0) + 2 * 64 * sizeof(float) + 1024;
}
cudaError_t launch_stem_rows(const StemRowsParams& p, const CUtensorMap& tmX, const CUtensorMap& tmB, cudaStream_t stream);

// dynamic shared memory needed for (block_n, stages)
inline size_t conv_epi_stage_bytes(int epi_group) { return epi_group ? (size_t)8 * 32 * ((size_t)epi_group * 32 + 16) : 0; }
inline size_t conv_smem_bytes(int block_n, int stages, int m_halves = 1, size_t epi_stage = 0) {
  size_t a = (size_t)stages * kBlockM * 128 * m_halves;
  size_t b = (size_t)stages * block_n * 128;
  size_t epi = 3 * 256 * sizeof(float);
  size_t bars = 64 * 8;
  return 1024 /*alignment slack*/ + a + b + epi + bars + epi_stage;
}

cudaError_t launch_conv_umma(const ConvKernelParams& p, const CUtensorMap& tmA, const CUtensorMap& tmB,
                             cudaStream_t stream);
// CTA-pair (cta_group::2) variant: p.stages / p.tmem_cols sized for it, tmBhalf has box rows = block_n / 2
cudaError_t launch_conv_pair(const ConvKernelParams& p, const CUtensorMap& tmA, const CUtensorMap& tmBhalf,
                             cudaStream_t stream);
cudaError_t conv_umma_configure();  // sets max dynamic smem attribute once

}  // namespace eco
