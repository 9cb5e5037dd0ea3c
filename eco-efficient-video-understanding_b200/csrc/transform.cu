// transform.cu -- see transform.cuh.
#include "transform.cuh"

#include <cmath>
#include <cstdlib>
#include <sstream>

namespace eco {
namespace {

// cv::resize, INTER_LINEAR, 8-bit source (OpenCV imgproc/resize.cpp: HResizeLinear / VResizeLinear<uchar, int, short,
// FixedPtCast<int, uchar, INTER_RESIZE_COEF_BITS*2>>): coefficients are 11-bit fixed point (cvRound(w * 2048)), the
// horizontal pass keeps 8+11 bits, the vertical pass computes ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.
// Source index: fx = (dx + 0.5) * scale - 0.5, clamped so that taps stay inside the image.
__device__ __forceinline__ void lin_coef(int d, float scale, int ssize, int& s0, int& s1, int& a0, int& a1) {
  float fx = ((float)d + 0.5f) * scale - 0.5f;
  int sx = (int)floorf(fx);
  fx -= (float)sx;
  if (sx < 0) { fx = 0.f; sx = 0; }
  if (sx >= ssize - 1) { fx = 0.f; sx = ssize - 1; }
  // saturate_cast<short>(v * 2048) rounds to nearest even (cvRound); the two weights are computed independently
  a0 = (int)rintf((1.f - fx) * 2048.f);
  a1 = (int)rintf(fx * 2048.f);
  s0 = sx;
  s1 = min(sx + 1, ssize - 1);
}

__global__ void transform_u8_kernel(const unsigned char* __restrict__ src, float* __restrict__ dst, int B, int C, int Hd, int Wd,
                                    int crop, const eco_clip_transform* __restrict__ tr, const float* __restrict__ mean, int nmean,
                                    float scale, int is_flow) {
  const long long total = (long long)B * C * crop * crop;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(t % crop);
    const int h = (int)((t / crop) % crop);
    const int c = (int)((t / ((long long)crop * crop)) % C);
    const int b = (int)(t / ((long long)crop * crop * C));
    const eco_clip_transform q = tr[b];
    const unsigned char* plane = src + ((long long)b * C + c) * Hd * Wd;
    int v;
    if (q.crop_h == crop && q.crop_w == crop) {
      v = plane[(long long)(q.h_off + h) * Wd + q.w_off + w];
    } else {
      int x0, x1, ax0, ax1, y0, y1, by0, by1;
      lin_coef(w, (float)q.crop_w / (float)crop, q.crop_w, x0, x1, ax0, ax1);
      lin_coef(h, (float)q.crop_h / (float)crop, q.crop_h, y0, y1, by0, by1);
      const unsigned char* r0 = plane + (long long)(q.h_off + y0) * Wd + q.w_off;
      const unsigned char* r1 = plane + (long long)(q.h_off + y1) * Wd + q.w_off;
      const int S0 = r0[x0] * ax0 + r0[x1] * ax1;
      const int S1 = r1[x0] * ax0 + r1[x1] * ax1;
      v = (((by0 * (S0 >> 4)) >> 16) + ((by1 * (S1 >> 4)) >> 16) + 2) >> 2;
      v = min(255, max(0, v));
    }
    float e = (float)v;
    if (is_flow && q.mirror && c < C / 2) e = 255.f - e;   // data_transformer.cpp:282-291
    const int wo = q.mirror ? crop - 1 - w : w;            // :275-279
    const float m = nmean > 0 ? mean[nmean == 1 ? 0 : c % nmean] : 0.f;   // mean_value replication :177-194
    dst[(((long long)b * C + c) * crop + h) * crop + wo] = (e - m) * scale;
  }
}

}  // namespace

cudaError_t launch_transform_u8(const unsigned char* src, float* dst, int B, int C, int Hd, int Wd, int crop,
                                const eco_clip_transform* t_dev, const float* mean_dev, int nmean, float scale, int is_flow,
                                cudaStream_t st) {
  const long long n = (long long)B * C * crop * crop;
  if (n == 0) return cudaSuccess;
  long long blocks = (n + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  transform_u8_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, B, C, Hd, Wd, crop, t_dev, mean_dev, nmean, scale, is_flow);
  return cudaGetLastError();
}

std::vector<VideoEntry> parse_video_list(const std::string& text) {
  std::vector<VideoEntry> out;
  std::istringstream in(text);
  VideoEntry e;
  while (in >> e.path >> e.num_frames >> e.label) out.push_back(e);
  return out;
}

void sample_segment_offsets(int num_frames, int num_segments, int new_length, bool train, std::mt19937& rng, int* offsets) {
  const double average_duration = (double)(num_frames / num_segments);  // integer division first, as lines_duration_ / num_segments on ints
  for (int i = 0; i < num_segments; ++i) {
    if (train) {
      if (average_duration >= new_length) {
        const int offset = (int)(rng() % (unsigned)((int)average_duration - new_length + 1));
        offsets[i] = (int)(offset + i * average_duration);
      } else {
        offsets[i] = (int)(i * average_duration);
      }
    } else {
      offsets[i] = average_duration >= new_length ? (int)((average_duration - new_length + 1) / 2 + i * average_duration) : 0;
    }
  }
}

std::vector<std::pair<int, int>> crop_size_candidates(int H, int W, int net_h, int net_w, int max_distort, const std::vector<float>& ratios) {
  static const float def[] = {1.0f, .875f, .75f, .66f};
  std::vector<float> r = ratios.empty() ? std::vector<float>(def, def + 4) : ratios;
  std::vector<std::pair<int, int>> out;
  const int base = H < W ? H : W;
  for (size_t h = 0; h < r.size(); ++h) {
    int ch = (int)(base * r[h]);
    if (std::abs(ch - net_h) < 3) ch = net_h;
    for (size_t w = 0; w < r.size(); ++w) {
      int cw = (int)(base * r[w]);
      if (std::abs(cw - net_w) < 3) cw = net_w;
      if (std::abs((int)h - (int)w) <= max_distort) out.emplace_back(ch, cw);
    }
  }
  return out;
}

std::vector<std::pair<int, int>> fix_offset_candidates(int H, int W, int crop_h, int crop_w, bool more) {
  const int ho = (H - crop_h) / 4, wo = (W - crop_w) / 4;
  std::vector<std::pair<int, int>> o = {{0, 0}, {0, 4 * wo}, {4 * ho, 0}, {4 * ho, 4 * wo}, {2 * ho, 2 * wo}};
  if (more) {
    const std::pair<int, int> extra[] = {{0, 2 * wo}, {4 * ho, 2 * wo}, {2 * ho, 0}, {2 * ho, 4 * wo},
                                         {1 * ho, 1 * wo}, {1 * ho, 3 * wo}, {3 * ho, 1 * wo}, {3 * ho, 3 * wo}};
    o.insert(o.end(), extra, extra + 8);
  }
  return o;
}

eco_clip_transform sample_clip_transform(int H, int W, int crop_size, bool train, const eco_transform_param& p, std::mt19937& rng) {
  eco_clip_transform t{};
  auto rnd = [&](int n) { return (int)(rng() % (unsigned)n); };   // DataTransformer::Rand
  t.mirror = (p.mirror && rnd(2)) ? 1 : 0;                        // drawn first (data_transformer.cpp:158), in both phases
  if (!crop_size) { t.crop_h = H; t.crop_w = W; return t; }
  if (train) {
    if (p.multi_scale) {
      std::vector<float> ratios(p.scale_ratios, p.scale_ratios + p.num_scale_ratios);
      auto cs = crop_size_candidates(H, W, crop_size, crop_size, p.max_distort, ratios);
      const auto& c = cs[(size_t)rnd((int)cs.size())];
      t.crop_h = c.first; t.crop_w = c.second;
    } else {
      t.crop_h = t.crop_w = crop_size;
    }
    if (p.fix_crop) {
      auto os = fix_offset_candidates(H, W, t.crop_h, t.crop_w, p.more_fix_crop != 0);
      const auto& o = os[(size_t)rnd((int)os.size())];
      t.h_off = o.first; t.w_off = o.second;
    } else {
      t.h_off = rnd(H - t.crop_h + 1);
      t.w_off = rnd(W - t.crop_w + 1);
    }
  } else {
    t.crop_h = t.crop_w = crop_size;
    t.h_off = (H - crop_size) / 2;
    t.w_off = (W - crop_size) / 2;
  }
  return t;
}

}  // namespace eco
