// capi.cpp -- extern "C" boundary of libeco_b200.so (include/eco_b200.h).  Every entry point
// catches C++ exceptions and turns them into a non-zero return + eco_last_error().
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>

#include "../../include/eco_b200.h"
#include "net.hpp"
#include "solver.hpp"
#include "transform.cuh"

struct eco_net {
  eco::Net* impl;
  bool borrowed = false;   // owned by a solver
};
struct eco_solver {
  eco::Solver* impl;
  eco_net net_handle;
};

namespace eco {
static thread_local std::string g_last_error;
void set_last_error(const std::string& s) { g_last_error = s; }
}  // namespace eco

#define ECO_API_BEGIN try {
#define ECO_API_END                         \
  }                                         \
  catch (const std::exception& e) {         \
    eco::set_last_error(e.what());          \
    return 1;                               \
  }                                         \
  catch (...) {                             \
    eco::set_last_error("unknown error");   \
    return 1;                               \
  }                                         \
  return 0;

static int g_mode_gpu = 1;

static eco::Net& N(eco_net* n) {
  if (!n || !n->impl) throw std::runtime_error("null eco_net handle");
  return *n->impl;
}
static const eco::Net& N(const eco_net* n) {
  if (!n || !n->impl) throw std::runtime_error("null eco_net handle");
  return *n->impl;
}

extern "C" {

const char* eco_last_error(void) { return eco::g_last_error.c_str(); }
const char* eco_version(void) { return "eco_b200 0.1 (sm_100a)"; }

int eco_set_device(int device) {
  ECO_API_BEGIN
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) throw std::runtime_error(std::string("cudaSetDevice failed: ") + cudaGetErrorString(e));
  ECO_API_END
}
int eco_set_mode(int gpu) {
  g_mode_gpu = gpu ? 1 : 0;
  return 0;
}
int eco_device_count(int* count) {
  ECO_API_BEGIN
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    n = 0;
  }
  if (count) *count = n;
  ECO_API_END
}

int eco_net_create_from_string(const char* text, int phase, eco_net** out) {
  ECO_API_BEGIN
  if (!text || !out) throw std::runtime_error("null argument");
  eco_net* h = new eco_net;
  h->impl = nullptr;
  try {
    h->impl = new eco::Net(text, phase);
  } catch (...) {
    delete h;
    throw;
  }
  *out = h;
  ECO_API_END
}
int eco_net_create_from_string_until(const char* text, int phase, const char* until_blob, eco_net** out) {
  ECO_API_BEGIN
  if (!text || !out) throw std::runtime_error("null argument");
  eco_net* h = new eco_net;
  h->impl = nullptr;
  try {
    h->impl = new eco::Net(text, phase, until_blob ? until_blob : "");
  } catch (...) {
    delete h;
    throw;
  }
  *out = h;
  ECO_API_END
}
int eco_net_push_frames(eco_net* dst, int dst_blob, eco_net* src, int src_blob) {
  ECO_API_BEGIN
  N(dst).push_frames(dst_blob, N(src), src_blob);
  ECO_API_END
}
int eco_net_create(const char* path, int phase, eco_net** out) {
  ECO_API_BEGIN
  if (!path) throw std::runtime_error("null path");
  std::ifstream f(path);
  if (!f) throw std::runtime_error(std::string("Could not open ") + path);  // _caffe.cpp:57-64
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  int rc = eco_net_create_from_string(text.c_str(), phase, out);
  if (rc) return rc;
  ECO_API_END
}
int eco_net_destroy(eco_net* net) {
  ECO_API_BEGIN
  if (net && !net->borrowed) {
    delete net->impl;
    delete net;
  }
  ECO_API_END
}
int eco_net_set_option(eco_net* net, const char* key, int value) {
  ECO_API_BEGIN
  N(net).set_option(key ? key : "", value);
  ECO_API_END
}
int eco_net_set_stream(eco_net* net, void* s) {
  ECO_API_BEGIN
  N(net).set_stream(static_cast<cudaStream_t>(s));
  ECO_API_END
}

int eco_net_copy_from(eco_net* net, const char* path) {
  ECO_API_BEGIN
  N(net).copy_from(path ? path : "");
  ECO_API_END
}
int eco_net_save(const eco_net* net, const char* path) {
  ECO_API_BEGIN
  N(net).save(path ? path : "");
  ECO_API_END
}
int eco_net_layer_num_params(const eco_net* net, int layer, int* n) {
  ECO_API_BEGIN
  *n = N(net).num_params(layer);
  ECO_API_END
}
int eco_net_param_shape(const eco_net* net, int layer, int idx, int* dims, int* ndims) {
  ECO_API_BEGIN
  const eco::ParamBlob& b = const_cast<eco::Net&>(N(net)).param(layer, idx);
  if (*ndims < (int)b.shape.size()) throw std::runtime_error("dims capacity too small");
  for (size_t i = 0; i < b.shape.size(); ++i) dims[i] = b.shape[i];
  *ndims = (int)b.shape.size();
  ECO_API_END
}
int eco_net_set_param(eco_net* net, int layer, int idx, const float* data, size_t count) {
  ECO_API_BEGIN
  N(net).set_param(layer, idx, data, count);
  ECO_API_END
}
int eco_net_get_param(const eco_net* net, int layer, int idx, float* data, size_t count) {
  ECO_API_BEGIN
  const eco::ParamBlob& b = const_cast<eco::Net&>(N(net)).param(layer, idx);
  if (count != b.data.size()) throw std::runtime_error("parameter size mismatch");
  std::memcpy(data, b.data.data(), count * sizeof(float));
  ECO_API_END
}
int eco_net_param_host(eco_net* net, int layer, int idx, float** data, size_t* count) {
  ECO_API_BEGIN
  eco::ParamBlob& b = N(net).param(layer, idx);
  N(net).mark_params_dirty(layer);  // caller may write through the pointer (mutable_cpu_data semantics)
  *data = b.data.data();
  if (count) *count = b.data.size();
  ECO_API_END
}

const char* eco_net_name(const eco_net* net) { return net && net->impl ? net->impl->name_.c_str() : ""; }
int eco_net_phase(const eco_net* net) { return net && net->impl ? net->impl->phase_ : -1; }
int eco_net_num_layers(const eco_net* net) { return net && net->impl ? (int)net->impl->vis_layers_.size() : -1; }
const char* eco_net_layer_name(const eco_net* net, int i) {
  if (!net || !net->impl || i < 0 || i >= (int)net->impl->vis_layers_.size()) return nullptr;
  return net->impl->vis_layers_[i].name.c_str();
}
const char* eco_net_layer_type(const eco_net* net, int i) {
  if (!net || !net->impl || i < 0 || i >= (int)net->impl->vis_layers_.size()) return nullptr;
  return net->impl->vis_layers_[i].type.c_str();
}
int eco_net_layer_index(const eco_net* net, const char* name) {
  if (!net || !net->impl || !name) return -1;
  auto it = net->impl->vis_layer_index_.find(name);
  return it == net->impl->vis_layer_index_.end() ? -1 : it->second;
}
int eco_net_layer_num_bottoms(const eco_net* net, int i) {
  if (!net || !net->impl || i < 0 || i >= (int)net->impl->vis_layers_.size()) return -1;
  return (int)net->impl->vis_layers_[i].bottoms.size();
}
int eco_net_layer_bottom(const eco_net* net, int i, int j) {
  if (eco_net_layer_num_bottoms(net, i) <= j || j < 0) return -1;
  return net->impl->vis_layers_[i].bottoms[j];
}
int eco_net_layer_num_tops(const eco_net* net, int i) {
  if (!net || !net->impl || i < 0 || i >= (int)net->impl->vis_layers_.size()) return -1;
  return (int)net->impl->vis_layers_[i].tops.size();
}
int eco_net_layer_top(const eco_net* net, int i, int j) {
  if (eco_net_layer_num_tops(net, i) <= j || j < 0) return -1;
  return net->impl->vis_layers_[i].tops[j];
}
int eco_net_num_blobs(const eco_net* net) { return net && net->impl ? (int)net->impl->vis_blobs_.size() : -1; }
const char* eco_net_blob_name(const eco_net* net, int i) {
  if (!net || !net->impl || i < 0 || i >= (int)net->impl->vis_blobs_.size()) return nullptr;
  return net->impl->vis_blobs_[i].name.c_str();
}
int eco_net_blob_index(const eco_net* net, const char* name) {
  if (!net || !net->impl || !name) return -1;
  auto it = net->impl->vis_blob_index_.find(name);
  return it == net->impl->vis_blob_index_.end() ? -1 : it->second;
}
int eco_net_blob_shape(const eco_net* net, int i, int* dims, int* ndims) {
  ECO_API_BEGIN
  const eco::Net& n = N(net);
  if (i < 0 || i >= (int)n.vis_blobs_.size()) throw std::runtime_error("blob index out of range");
  const auto& s = n.tensors_[n.vis_blobs_[i].tensor].shape;
  if (*ndims < (int)s.size()) throw std::runtime_error("dims capacity too small");
  for (size_t k = 0; k < s.size(); ++k) dims[k] = s[k];
  *ndims = (int)s.size();
  ECO_API_END
}
int eco_net_num_inputs(const eco_net* net) { return net && net->impl ? (int)net->impl->inputs_.size() : -1; }
int eco_net_input_blob(const eco_net* net, int i) {
  if (!net || !net->impl || i < 0 || i >= (int)net->impl->inputs_.size()) return -1;
  return net->impl->inputs_[i];
}
int eco_net_num_outputs(const eco_net* net) { return net && net->impl ? (int)net->impl->outputs_.size() : -1; }
int eco_net_output_blob(const eco_net* net, int i) {
  if (!net || !net->impl || i < 0 || i >= (int)net->impl->outputs_.size()) return -1;
  return net->impl->outputs_[i];
}

int eco_blob_reshape(eco_net* net, int blob, const int* dims, int ndims) {
  ECO_API_BEGIN
  N(net).reshape_blob(blob, std::vector<int>(dims, dims + ndims));
  ECO_API_END
}
int eco_net_reshape(eco_net* net) {
  ECO_API_BEGIN
  N(net).reshape();
  ECO_API_END
}

int eco_net_forward(eco_net* net, int start, int end, float* loss) {
  ECO_API_BEGIN
  if (!g_mode_gpu)
    throw std::runtime_error("set_mode_cpu() was requested: libeco_b200 has no CPU execution path (call set_mode_gpu())");
  const float l = N(net).forward(start, end);
  if (loss) *loss = l;
  ECO_API_END
}
int eco_net_backward(eco_net* net, int start, int end) {
  ECO_API_BEGIN
  if (!g_mode_gpu)
    throw std::runtime_error("set_mode_cpu() was requested: libeco_b200 has no CPU execution path (call set_mode_gpu())");
  N(net).backward(start, end);
  ECO_API_END
}
int eco_net_clear_param_diffs(eco_net* net) {
  ECO_API_BEGIN
  N(net).clear_param_diffs();
  ECO_API_END
}
int eco_net_update(eco_net* net) {
  ECO_API_BEGIN
  size_t n = 0;
  float* P = N(net).param_arena(&n);
  float* G = N(net).grad_arena(&n);
  cudaError_t e = eco::launch_f32_axpy(G, P, (long long)n, -1.f, 1, N(net).stream());
  if (e != cudaSuccess) throw std::runtime_error(std::string("Net::Update: ") + cudaGetErrorString(e));
  N(net).params_updated_on_device();
  ECO_API_END
}
int eco_net_param_diff_host(eco_net* net, int layer, int idx, float** data, size_t* count) {
  ECO_API_BEGIN
  *data = N(net).param_diff_host(layer, idx, count);
  ECO_API_END
}
int eco_net_param_arena(eco_net* net, float** dev, size_t* count) {
  ECO_API_BEGIN
  *dev = N(net).param_arena(count);
  ECO_API_END
}
int eco_net_grad_arena(eco_net* net, float** dev, size_t* count) {
  ECO_API_BEGIN
  *dev = N(net).grad_arena(count);
  ECO_API_END
}
int eco_net_num_param_slots(eco_net* net, int* n) {
  ECO_API_BEGIN
  *n = (int)N(net).param_slots().size();
  ECO_API_END
}
int eco_net_param_slot(eco_net* net, int i, eco_param_slot* out) {
  ECO_API_BEGIN
  const auto& sl = N(net).param_slots();
  if (i < 0 || i >= (int)sl.size()) throw std::runtime_error("parameter slot index out of range");
  out->layer = N(net).layers_[sl[i].layer].vis_index;
  out->blob = sl[i].idx;
  out->offset = sl[i].off;
  out->count = sl[i].count;
  out->lr_mult = sl[i].lr_mult;
  out->decay_mult = sl[i].decay_mult;
  ECO_API_END
}
int eco_net_params_updated_on_device(eco_net* net) {
  ECO_API_BEGIN
  N(net).params_updated_on_device();
  ECO_API_END
}
int eco_net_cuda_stream(eco_net* net, void** stream) {
  ECO_API_BEGIN
  *stream = static_cast<void*>(N(net).stream());
  ECO_API_END
}
int eco_net_sync(eco_net* net) {
  ECO_API_BEGIN
  N(net).sync();
  ECO_API_END
}

int eco_blob_host_data(eco_net* net, int blob, int for_write, float** data, size_t* count) {
  ECO_API_BEGIN
  *data = N(net).host_data(blob, for_write != 0, count);
  ECO_API_END
}
int eco_blob_host_diff(eco_net* net, int blob, int for_write, float** data, size_t* count) {
  ECO_API_BEGIN
  *data = N(net).host_diff(blob, for_write != 0, count);
  ECO_API_END
}
int eco_net_set_input_device(eco_net* net, int blob, const void* dev, size_t count) {
  ECO_API_BEGIN
  N(net).set_input_device(blob, dev, count);
  ECO_API_END
}
int eco_blob_device_f32(eco_net* net, int blob, const float** dev, size_t* count) {
  ECO_API_BEGIN
  *dev = N(net).device_f32(blob, count);
  ECO_API_END
}
int eco_net_forward_pipelined(eco_net* net, const float* host_in, size_t count, float* host_out, size_t out_count,
                              int* ticket) {
  ECO_API_BEGIN
  if (!g_mode_gpu) throw std::runtime_error("set_mode_cpu() was requested: libeco_b200 has no CPU execution path");
  const int t = N(net).forward_pipelined(host_in, count, host_out, out_count);
  if (ticket) *ticket = t;
  ECO_API_END
}
int eco_net_forward_pipelined_u8(eco_net* net, const unsigned char* host_in, size_t count, const float* mean, int nmean,
                                 float* host_out, size_t out_count, int* ticket) {
  ECO_API_BEGIN
  if (!g_mode_gpu) throw std::runtime_error("set_mode_cpu() was requested: libeco_b200 has no CPU execution path");
  const int t = N(net).forward_pipelined_u8(host_in, count, mean, nmean, host_out, out_count);
  if (ticket) *ticket = t;
  ECO_API_END
}
int eco_net_wait(eco_net* net, int ticket) {
  ECO_API_BEGIN
  N(net).wait_ticket(ticket);
  ECO_API_END
}
int eco_host_alloc(void** ptr, size_t bytes) {
  ECO_API_BEGIN
  if (!ptr) throw std::runtime_error("null argument");
  cudaError_t e = cudaMallocHost(ptr, bytes ? bytes : 16);
  if (e != cudaSuccess) throw std::runtime_error(std::string("cudaMallocHost failed: ") + cudaGetErrorString(e));
  ECO_API_END
}
int eco_host_free(void* ptr) {
  ECO_API_BEGIN
  if (ptr) cudaFreeHost(ptr);
  ECO_API_END
}
int eco_net_last_launch_count(const eco_net* net, int* launches) {
  ECO_API_BEGIN
  *launches = N(net).last_launches();
  ECO_API_END
}
int eco_net_profile_forward(eco_net* net, eco_op_time* out, int cap, int* n) {
  ECO_API_BEGIN
  *n = N(net).profile(out, cap);
  ECO_API_END
}
int eco_net_profile_train(eco_net* net, eco_op_time* out, int cap, int* n) {
  ECO_API_BEGIN
  *n = N(net).profile_train(out, cap);
  ECO_API_END
}
int eco_net_describe_plan(eco_net* net, char* buf, size_t cap, size_t* needed) {
  ECO_API_BEGIN
  const std::string d = N(net).describe_plan();
  if (needed) *needed = d.size() + 1;
  if (buf && cap) {
    const size_t n = d.size() < cap - 1 ? d.size() : cap - 1;
    std::memcpy(buf, d.data(), n);
    buf[n] = 0;
  }
  ECO_API_END
}

int eco_net_set_grad_bucket_hook(eco_net* net, int nbuckets, eco_grad_bucket_fn fn, void* user) {
  ECO_API_BEGIN
  N(net).set_grad_bucket_hook(nbuckets, fn, user);
  ECO_API_END
}
int eco_net_num_grad_buckets(eco_net* net, int* n) {
  ECO_API_BEGIN
  *n = (int)N(net).grad_buckets().size();
  ECO_API_END
}
int eco_net_grad_bucket(eco_net* net, int i, size_t* offset, size_t* count) {
  ECO_API_BEGIN
  const auto& b = N(net).grad_buckets();
  if (i < 0 || i >= (int)b.size()) throw std::runtime_error("bucket index out of range");
  *offset = b[i].off;
  *count = b[i].count;
  ECO_API_END
}

/* ---- sampling + DataTransformer on the GPU ---- */
struct eco_sampler {
  std::mt19937 frame_rng, transform_rng;
};
int eco_sampler_create(unsigned int seed, eco_sampler** out) {
  ECO_API_BEGIN
  eco_sampler* s = new eco_sampler;
  s->frame_rng.seed(seed);
  s->transform_rng.seed(seed + 1u);
  *out = s;
  ECO_API_END
}
int eco_sampler_destroy(eco_sampler* s) {
  ECO_API_BEGIN
  delete s;
  ECO_API_END
}
int eco_sample_segment_offsets(eco_sampler* s, int num_frames, int num_segments, int new_length, int train, int* offsets) {
  ECO_API_BEGIN
  if (!s || !offsets || num_segments <= 0) throw std::runtime_error("bad argument");
  eco::sample_segment_offsets(num_frames, num_segments, new_length, train != 0, s->frame_rng, offsets);
  ECO_API_END
}
int eco_sample_clip_transform(eco_sampler* s, int H, int W, int crop_size, int train, const eco_transform_param* p, eco_clip_transform* out) {
  ECO_API_BEGIN
  if (!s || !p || !out) throw std::runtime_error("null argument");
  if (crop_size && (H < crop_size || W < crop_size)) throw std::runtime_error("datum smaller than crop_size");  // CHECK_GE :169-170
  *out = eco::sample_clip_transform(H, W, crop_size, train != 0, *p, s->transform_rng);
  ECO_API_END
}
int eco_crop_size_candidates(int H, int W, int crop_size, int max_distort, const float* ratios, int nratios, int* hw, int* n) {
  ECO_API_BEGIN
  auto c = eco::crop_size_candidates(H, W, crop_size, crop_size, max_distort, std::vector<float>(ratios, ratios + (ratios ? nratios : 0)));
  if ((int)c.size() > *n) throw std::runtime_error("capacity too small");
  for (size_t i = 0; i < c.size(); ++i) { hw[2 * i] = c[i].first; hw[2 * i + 1] = c[i].second; }
  *n = (int)c.size();
  ECO_API_END
}
int eco_fix_offset_candidates(int H, int W, int crop_h, int crop_w, int more, int* hw, int* n) {
  ECO_API_BEGIN
  auto c = eco::fix_offset_candidates(H, W, crop_h, crop_w, more != 0);
  if ((int)c.size() > *n) throw std::runtime_error("capacity too small");
  for (size_t i = 0; i < c.size(); ++i) { hw[2 * i] = c[i].first; hw[2 * i + 1] = c[i].second; }
  *n = (int)c.size();
  ECO_API_END
}
int eco_net_transform_input_u8(eco_net* net, int blob, const unsigned char* src, int B, int C, int H, int W,
                               const eco_clip_transform* t, const eco_transform_param* p) {
  ECO_API_BEGIN
  if (!src || !t || !p) throw std::runtime_error("null argument");
  N(net).transform_input_u8(blob, src, B, C, H, W, t, *p);
  ECO_API_END
}

/* ---- solver ---- */
static eco::Solver& S(eco_solver* s) {
  if (!s || !s->impl) throw std::runtime_error("null eco_solver handle");
  return *s->impl;
}
static int solver_create(const std::string& text, const std::string& net_text, const std::string& dir, eco_solver** out) {
  ECO_API_BEGIN
  if (!out) throw std::runtime_error("null argument");
  eco_solver* h = new eco_solver;
  h->impl = nullptr;
  try {
    h->impl = new eco::Solver(text, net_text, dir);
  } catch (...) {
    delete h;
    throw;
  }
  h->net_handle.impl = &h->impl->net();
  h->net_handle.borrowed = true;
  *out = h;
  ECO_API_END
}
int eco_solver_create(const char* solver_prototxt_path, eco_solver** out) {
  if (!solver_prototxt_path) { eco::set_last_error("null path"); return 1; }
  std::ifstream f(solver_prototxt_path);
  if (!f) { eco::set_last_error(std::string("Could not open ") + solver_prototxt_path); return 1; }
  std::stringstream ss;
  ss << f.rdbuf();
  std::string dir(solver_prototxt_path);
  const size_t slash = dir.find_last_of('/');
  dir = slash == std::string::npos ? std::string() : dir.substr(0, slash);
  return solver_create(ss.str(), "", dir, out);
}
int eco_solver_create_from_string(const char* solver_text, const char* net_text, eco_solver** out) {
  return solver_create(solver_text ? solver_text : "", net_text ? net_text : "", "", out);
}
int eco_solver_destroy(eco_solver* s) {
  ECO_API_BEGIN
  if (s) {
    delete s->impl;
    delete s;
  }
  ECO_API_END
}
int eco_solver_net(eco_solver* s, eco_net** net) {
  ECO_API_BEGIN
  S(s);
  *net = &s->net_handle;
  ECO_API_END
}
int eco_solver_iter(eco_solver* s, int* iter) {
  ECO_API_BEGIN
  *iter = S(s).iter();
  ECO_API_END
}
int eco_solver_learning_rate(eco_solver* s, float* rate) {
  ECO_API_BEGIN
  *rate = S(s).learning_rate();
  ECO_API_END
}
int eco_solver_step(eco_solver* s, int iters, float* loss) {
  ECO_API_BEGIN
  if (!g_mode_gpu) throw std::runtime_error("set_mode_cpu() was requested: libeco_b200 has no CPU execution path");
  const float l = S(s).step(iters);
  if (loss) *loss = l;
  ECO_API_END
}
int eco_solver_apply_update(eco_solver* s) {
  ECO_API_BEGIN
  S(s).apply_update();
  ECO_API_END
}
int eco_solver_set_grad_sync(eco_solver* s, eco_grad_sync_fn fn, void* user, int world) {
  ECO_API_BEGIN
  S(s).set_grad_sync(fn, user, world);
  ECO_API_END
}
int eco_solver_snapshot(eco_solver* s, const char* prefix) {
  ECO_API_BEGIN
  S(s).snapshot(prefix ? prefix : "");
  ECO_API_END
}
int eco_solver_restore(eco_solver* s, const char* state_file) {
  ECO_API_BEGIN
  S(s).restore(state_file ? state_file : "");
  ECO_API_END
}

}  // extern "C"
