// transform.cuh -- DataTransformer::Transform on the GPU (caffe_3d/src/caffe/data_transformer.cpp:148-326) for batches of
// RGB / flow clips in Datum layout, plus the host-side choices of VideoDataLayer / DataTransformer (segment sampling,
// multi-scale crop sizes, fixed crop offsets, mirror): the step in front of the hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <random>
#include <string>
#include <utility>
#include <vector>

#include "../../include/eco_b200.h"

namespace eco {

// src: uint8 [B][C][Hd][Wd] on the DEVICE (Datum layout: C = 3 * segments for RGB); dst: fp32 [B][C][crop][crop].
// Per clip: crop window (h_off, w_off, crop_h, crop_w), resized to crop x crop with OpenCV's INTER_LINEAR 8-bit fixed-point
// arithmetic when it differs from crop (data_transformer.cpp:252-270 calls cv::resize per channel plane), mirror, flow
// inversion (255 - v on the first half of the channels when mirrored), (v - mean[c]) * scale.
cudaError_t launch_transform_u8(const unsigned char* src, float* dst, int B, int C, int Hd, int Wd, int crop,
                                const eco_clip_transform* t_dev, const float* mean_dev, int nmean, float scale, int is_flow,
                                cudaStream_t st);

// ---- host logic ----
struct VideoEntry { std::string path; int num_frames = 0; int label = 0; };
// "path num_frames label" per line (video_data_layer.cpp:44-52)
std::vector<VideoEntry> parse_video_list(const std::string& text);
// VideoDataLayer::InternalThreadEntry :155-187: TRAIN = random offset inside each of the N equal segments, TEST = centre
void sample_segment_offsets(int num_frames, int num_segments, int new_length, bool train, std::mt19937& rng, int* offsets);
// fillCropSize (data_transformer.cpp:83-105) and fillFixOffset (:50-78)
std::vector<std::pair<int, int>> crop_size_candidates(int H, int W, int net_h, int net_w, int max_distort, const std::vector<float>& ratios);
std::vector<std::pair<int, int>> fix_offset_candidates(int H, int W, int crop_h, int crop_w, bool more);
// the random choices of Transform (:158, :214-244), drawn in the reference's order: mirror, crop size, offset
eco_clip_transform sample_clip_transform(int H, int W, int crop_size, bool train, const eco_transform_param& p, std::mt19937& rng);

}  // namespace eco
