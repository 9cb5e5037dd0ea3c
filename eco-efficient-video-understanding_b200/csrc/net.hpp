// net.hpp -- host side of libeco_b200: the caffe-visible graph (layers / blobs / auto-Splits, names as
// caffe_3d's Net::Init produces them, net.cpp:39-316 + util/insert_splits.cpp) and the fused
// execution plan that runs it on sm_100a.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "conv_umma.cuh"
#include "prototxt.hpp"
#include "train_kernels.cuh"
#include "wgrad_umma.cuh"

struct eco_op_time;
struct eco_clip_transform;
struct eco_transform_param;

namespace eco {

enum class Kind { F32, CL };  // plain fp32 in caffe layout | bf16 channels-last

// fp32 host mirror; page-locked when a CUDA device is usable so uploads/downloads are async DMA
struct HostBuf {
  float* p = nullptr;
  size_t n = 0;
  bool pinned = false;
  HostBuf() = default;
  HostBuf(const HostBuf&) = delete;
  HostBuf& operator=(const HostBuf&) = delete;
  HostBuf(HostBuf&& o) noexcept : p(o.p), n(o.n), pinned(o.pinned) { o.p = nullptr; o.n = 0; }
  HostBuf& operator=(HostBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; pinned = o.pinned; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~HostBuf() { release(); }
  void resize(size_t count, bool try_pin);  // contents zeroed when (re)allocated
  void release();
  bool empty() const { return p == nullptr; }
};

// A tensor = one caffe blob of the original (pre-Split) graph.
struct Tensor {
  std::string name;
  std::vector<int> shape;  // caffe logical shape
  Kind kind = Kind::F32;
  int ch_axis = 1;         // CL: logical axis that is the channel
  // storage: `root` tensor owns the allocation; views / concat slices point into it
  int root = -1;
  long long cs = 0;        // CL channel stride (elements) of the underlying buffer
  int coff = 0;            // CL first channel in the buffer
  bool materialized = false;  // has device storage that the plan writes
  void* dev = nullptr;     // device pointer (CL: bf16 base of the *buffer*, F32: float*)
  size_t dev_bytes = 0;
  bool owns = false;
  // host mirror, fp32 caffe layout (shared by all Split copies, like ShareData)
  HostBuf host, host_diff;
  bool host_newer = false;  // host written since the last upload
  bool dev_newer = false;   // device written since the last download
  // gradient (TRAIN-phase nets): same layout as the data (bf16 channels-last view or plain fp32)
  void* ddev = nullptr;
  bool needs_grad = false;       // lies downstream of a parameter: Backward computes its diff
  bool diff_host_newer = false;  // host_diff written by the caller (seed for backward)
  bool diff_dev_newer = false;
  cudaEvent_t h2d_done = nullptr;   // recorded after the last async upload from `host`: writers of the mirror wait for it
  int producer = -1;        // orig layer index that (last) writes it, -1: net input
  std::vector<int> consumers;  // orig layer indices reading it, in order

  long long count() const {
    long long n = 1;
    for (int d : shape) n *= d;
    return n;
  }
  int C() const { return shape[ch_axis]; }
  long long outer() const {
    long long n = 1;
    for (int i = 0; i < ch_axis; ++i) n *= shape[i];
    return n;
  }
  long long inner() const {
    long long n = 1;
    for (size_t i = ch_axis + 1; i < shape.size(); ++i) n *= shape[i];
    return n;
  }
};

struct ParamBlob {
  std::vector<int> shape;
  std::vector<float> data;
  std::vector<float> diff;
};

struct OrigLayer {
  std::string name, type;
  const pt::Msg* msg = nullptr;
  std::vector<int> bottoms, tops;  // tensor ids
  std::vector<ParamBlob> params;
  bool params_dirty = true;
  int vis_index = -1;  // index in the visible layer list
};

// What caffe shows: layer list with Split layers, blob list with split outputs.
struct VisLayer {
  std::string name, type;
  int orig = -1;  // OrigLayer index, -1 for auto Split
  std::vector<int> bottoms, tops;  // visible blob ids
};
struct VisBlob {
  std::string name;
  int tensor = -1;
};

// one of several 1x1 convolutions on the same input that run as a single GEMM (ConvOp::members)
struct ConvMember {
  int conv_layer = -1, bn_layer = -1;
  bool relu = false, post_pool = false;
  int cout = 0, off = 0;   // channels, first channel inside the fused N range
  int tensor = -1;         // the tensor this member stores
};

struct ConvOp {
  int conv_layer = -1, bn_layer = -1, elt_layer = -1;
  bool relu = false;
  bool stem = false;  // 7x7/s2/p3 Cin=3 handled as a 4x1 conv over space-to-depth windows
  int in_tensor = -1;
  int out_tensor = -1;  // y (BN/ReLU applied if fused), -1 if not stored
  int raw_tensor = -1;  // raw = conv + bias (+ residual), -1 if not stored
  int res_tensor = -1;  // residual input
  // geometry
  int nsp = 2, Cin = 0, Cin_k = 0, Cout = 0, Cout_pad = 0;
  int cin_stride = 0;  // split precision: channels per plane of the input operand (plane p, channel c -> p * cin_stride + c)
  int K[3] = {1, 1, 1}, S[3] = {1, 1, 1}, P[3] = {0, 0, 0};
  int I[3] = {1, 1, 1}, O[3] = {1, 1, 1};
  int NB = 0;
  long long Ktotal = 0;
  // device constants
  __nv_bfloat16* w_dev = nullptr;
  float *bias_dev = nullptr, *scale_dev = nullptr, *shift_dev = nullptr;
  __nv_bfloat16* stem_in = nullptr;  // transformed input (stem / generic fp32 input)
  size_t stem_bytes = 0;
  int stem_CH = 0, stem_CW = 0;
  ConvKernelParams kp{};
  CUtensorMap tmA{}, tmB{};
  bool halo = false;          // halo-resident kernel (2-D, stride 1, filter > 1x1)
  int halo_mt = 1;
  HaloKernelParams hp{};
  CUtensorMap tmX{};
  bool pair = false;          // CTA-pair (cta_group::2) kernel
  ConvKernelParams kpp{};     //   its parameters (stages / TMEM sized for half weight tiles)
  CUtensorMap tmBh{};         //   weight map with box rows = block_n / 2
  bool direct_in = false;     // rows kernel reads the fp32 / uint8 frames itself (no transform kernel, no cell buffer)
  bool rows = false;          // stem rows kernel (sliding window over cell rows, optional fused pool1)
  int pool_layer = -1;        // Pooling layer folded into the rows kernel
  int pool_tensor = -1;       // its top (the only tensor the fused op stores)
  StemRowsParams rp{};
  std::vector<ConvMember> members;  // non-empty: fused sibling 1x1 convolutions (fuse_1x1), Cout = sum of theirs
  bool group_scale = false;         //   some member has a folded BN
  float *post_bias_dev = nullptr, *post_scale_dev = nullptr, *post_shift_dev = nullptr;  // for pooling ops behind post_pool members
  bool post_pool = false;     // 1x1 conv moved in front of the AVE pooling that fed it: stores acc only (no bias / BN / ReLU);
  bool post_relu = false;     //   the pooling op behind it applies bias, BN and ReLU (pool_commute)
  double flops = 0, bytes = 0;
};

// one parameter blob inside the device arenas of a TRAIN-phase net (fp32 master weights / gradients, caffe layout)
struct ParamSlot {
  int layer = -1, idx = -1;
  size_t off = 0, count = 0;           // in floats
  float lr_mult = 1.f, decay_mult = 1.f;  // ParamSpec (caffe.proto), BN running statistics forced to 0 (bn_layer.cpp:46-53)
};

// backward state of one forward op
struct TrainAux {
  // BN TRAIN: saved batch statistics (bn_layer.cpp:183-188 keeps x_norm_ / x_inv_std_; x_norm is recomputed here)
  float *mean = nullptr, *inv_std = nullptr, *batch_var = nullptr, *sums = nullptr;
  float *slope = nullptr, *bias = nullptr, *run_mean = nullptr, *run_var = nullptr;  // into the parameter arena
  float *dslope = nullptr, *dbias = nullptr;                                          // into the gradient arena
  float momentum = 0.9f, eps = 1e-5f;
  bool relu = false;
  int x_tensor = -1, y_tensor = -1;
  // convolution backward
  bool has_dgrad = false, dilated = false;
  bool db_zero = false;             // the only reader of the conv top is a batch-statistics BN: its bias gradient is identically 0
  bool compact = false;             // 1x1 strided convolution: dgrad on the compact dY, then a strided scatter into dX
  int dg = -1;                      // index into dgrads_ (a ConvOp describing dX = conv(dY', flipped W^T))
  __nv_bfloat16* dil = nullptr;     // zero-dilated dY for strided convolutions [NB, E..., Cout]
  int E[3] = {1, 1, 1};
  WgradParams wg{};
  CUtensorMap tmY{};
  float *dw = nullptr, *db = nullptr, *w_master = nullptr;
  // dropout
  float ratio = 0.f;
  // loss / fc
  float loss_weight = 1.f;
  float* prob = nullptr;
  int top_k = 1;
};

struct Op {
  enum Type { CONV, POOL_CL, GLOBAL_AVG, POOL_F32, FC, SSR, ELTWISE, COPY2D, CL_TO_F32, F32_TO_CL, SOFTMAX,
              BN_TRAIN, DROPOUT, LOSS, ACCURACY };
  Type type;
  std::string name;
  int first_layer = 0, last_layer = 0;  // orig layer range covered
  int conv = -1;                        // index into convs_
  int affine_conv = -1;                 // POOL_CL: conv whose bias / BN / ReLU this pooling applies (pool_commute)
  int affine_off = -1;                  //   >= 0: that conv is member of a fused group, first channel inside it
  // generic payload
  int in0 = -1, in1 = -1, out = -1;     // tensor ids
  int layer = -1;                       // orig layer (params, pooling geometry)
  PoolParams pool{};
  PoolF32Params poolf{};
  float *scale_dev = nullptr, *shift_dev = nullptr;  // SSR
  bool relu = false;
  float *w_dev = nullptr, *b_dev = nullptr;          // FC
  int M = 0, N = 0, Kd = 0;
  // COPY2D
  size_t width_bytes = 0, rows = 0, src_pitch = 0, dst_pitch = 0, src_off = 0, dst_off = 0;
  double flops = 0, bytes = 0;
  int launches = 1;
};

class Net {
 public:
  // `until_blob`: keep only the layers up to the last one that writes this blob (a trunk-only net for the online feature cache)
  Net(const std::string& text, int phase, const std::string& until_blob = "");
  ~Net();

  // registry
  std::string name_;
  int phase_;
  std::vector<VisLayer> vis_layers_;
  std::vector<VisBlob> vis_blobs_;
  std::vector<OrigLayer> layers_;
  std::vector<Tensor> tensors_;
  std::vector<int> inputs_, outputs_;  // visible blob ids
  std::map<std::string, int> vis_layer_index_, vis_blob_index_;

  void set_option(const std::string& key, int v);
  void set_stream(cudaStream_t s);
  void reshape_blob(int vis_blob, const std::vector<int>& dims);
  void reshape();
  float forward(int start, int end);
  int forward_pipelined(const float* host_in, size_t count, float* host_out, size_t out_count);
  int forward_pipelined_u8(const unsigned char* host_in, size_t count, const float* mean, int nmean, float* host_out,
                           size_t out_count);
  void wait_ticket(int ticket);
  void sync();
  float* host_data(int vis_blob, bool for_write, size_t* count);
  float* host_diff(int vis_blob, bool for_write, size_t* count);
  void set_input_device(int vis_blob, const void* dev, size_t count);
  const float* device_f32(int vis_blob, size_t* count);
  void set_param(int vis_layer, int idx, const float* data, size_t count);
  ParamBlob& param(int vis_layer, int idx);
  int num_params(int vis_layer) const;
  void mark_params_dirty(int vis_layer);
  int profile(eco_op_time* out, int cap);
  // TRAIN nets: one forward + backward with a CUDA event after every op / backward sub-step (wgrad, dgrad, bias, ...)
  int profile_train(eco_op_time* out, int cap);
  std::string describe_plan();
  int last_launches() const { return last_launches_; }
  // online sliding window (scripts/online_recognition/online_recognition.py:64-93 recomputes all N frames per step): shift the
  // frames (outer index) of a channels-last blob of THIS net towards 0 by the frame count of `src_blob` of net `src` and
  // append those frames at the end, device to device, ordered after src's stream
  void push_frames(int dst_vis_blob, Net& src, int src_vis_blob);
  // DataTransformer::Transform on the GPU into an fp32 input blob (transform.cuh)
  void transform_input_u8(int vis_blob, const unsigned char* src, int B, int C, int H, int W, const eco_clip_transform* t,
                          const eco_transform_param& p);
  void copy_from(const std::string& path);
  void save(const std::string& path) const;
  // ---- training path (train.cpp): Net::BackwardFromTo net.cpp:637-706, params()/diffs ----
  void backward(int start, int end);
  void clear_param_diffs();
  float* param_diff_host(int vis_layer, int idx, size_t* count);
  bool is_train() const { return train_; }
  // device arenas (valid after the first forward / plan): all parameter blobs in layer order, fp32, caffe layout
  float* param_arena(size_t* count);
  float* grad_arena(size_t* count);
  const std::vector<ParamSlot>& param_slots();
  void params_updated_on_device();   // a solver changed the arena: repack the GEMM operands, host copies are stale
  // Gradient exchange hook: the arena is cut into `n` contiguous buckets of roughly equal size (slot boundaries, layer
  // order); during Backward, as soon as every layer of a bucket has produced its gradients, fn(user, bucket, offset, count)
  // is called on the host (the kernels are enqueued on stream(), not finished): the callee records an event and starts
  // the all-reduce of grad[offset, offset + count) on a side stream while the earlier layers are still being computed.
  typedef void (*BucketFn)(void* user, int bucket, size_t offset, size_t count);
  void set_grad_bucket_hook(int n, BucketFn fn, void* user);
  struct Bucket { int min_layer; size_t off, count; };
  const std::vector<Bucket>& grad_buckets();
  cudaStream_t stream();
  float last_loss() const { return last_loss_; }

 private:
  std::shared_ptr<pt::Msg> proto_;
  std::string until_blob_;
  std::string text_;
  // chunked blocking forward (option h2d_chunks): the batch is split into sub-batches, each run by a sub-net on its own
  // stream, so the host->device copy of sub-batch k+1 overlaps the compute of sub-batch k inside ONE caffe-style forward()
  int h2d_chunks_ = 0;                       // 0 auto, 1 off, n force
  std::vector<std::unique_ptr<Net>> sub_;
  std::vector<cudaEvent_t> sub_done_;
  unsigned long long params_version_ = 0, sub_params_version_ = ~0ull;
  bool chunked_last_ = false;
  bool is_sub_ = false;
  bool try_chunked_forward(int* launches);
  void release_subs();
  cudaEvent_t push_event_ = nullptr;
  std::vector<std::shared_ptr<pt::Msg>> keep_;
  std::map<std::string, int> tensor_index_;
  // options
  bool keep_all_ = false;
  int a_mode_ = -1;
  bool use_graph_ = false;
  int persistent_ = 1;
  int dual_m_ = 1;
  int halo_ = 0;  // 0 off (default: measured slower, profiles/r01h), 1 auto (resident weights only), 2 force two halves, 3 allow streamed weights
  int fuse_1x1_ = 1;  // 1: 1x1 convolutions reading the same tensor run as one GEMM with segmented output (fast plan only)
  int multicast_ = 0;  // persistent kernel in clusters of 2 sharing each weight tile by TMA multicast: 0 off (default: measured no gain -- the limit is the per-SM TMA ingest, not L2; profiles/r01m), 1 layers with >= 2 tiles per SM, 2 always
  int pair_ = 1;  // CTA-pair (cta_group::2) kernel: 0 off, 1 for 256-wide tiles with >= 2 tiles per SM pair, 2 wherever possible
  int stem_gather_warps_ = 4;  // window-gather warps of the direct stem kernel (4..6)
  int stem_direct_ = 1;  // 1: the stem rows kernel gathers its windows from the raw frames (frame width % 16 == 0)
  int pool_commute_ = 1;  // 1: AVE 3x3/s1 pooling -> 1x1 conv (+BN+ReLU) runs as conv -> pooling(+bias+BN+ReLU) when nothing else reads the pooled blob
  int stem_rows_ = 1;  // 0: stem as 4x1 im2col GEMM; 1: rows kernel, pool1 folded in when its input has no other reader; 2: rows kernel, never fold the pool
  int debug_flags_ = 0;
  int precision_ = 0;  // 1: split-precision storage (three bf16 planes hi|lo|hi per map, weights w_hi|w_hi|w_lo): ~fp32-faithful forward
  bool epi_staged_ = false;
  bool user_stream_ = false;
  // plan
  bool planned_ = false;
  std::vector<Op> ops_;
  std::vector<ConvOp> convs_;
  std::vector<void*> allocs_;
  cudaStream_t stream_ = nullptr;
  bool own_stream_ = false;
  int* error_flag_dev_ = nullptr;
  int last_launches_ = 0;
  cudaGraphExec_t graph_exec_ = nullptr;
  bool graph_valid_ = false;
  std::vector<std::string> op_names_;
  // pipelined serving path
  cudaStream_t copy_stream_ = nullptr;
  float* pipe_slot_[2] = {nullptr, nullptr};
  cudaEvent_t ev_h2d_[2] = {nullptr, nullptr}, ev_slot_free_[2] = {nullptr, nullptr}, ev_done_[2] = {nullptr, nullptr};
  unsigned long long pipe_iter_ = 0;
  int pipe_elem_bytes_ = 0;
  const void* u8_src_ = nullptr;
  float u8_mean_[4] = {0, 0, 0, 0};
  int forward_pipelined_any(const void* host_in, size_t count, int elem_bytes, const float* mean4, float* host_out,
                            size_t out_count);

  void build_graph();
  void init_params();
  void infer_shapes();
  void plan();
  void free_plan();
  void upload_params();
  void upload_dirty_inputs(int first_op);
  void run_op(Op& op, bool with_xform = true);
  void run_input_xform(ConvOp& c, const float* src);
  void run_ops(bool full, int lo, int hi, const float* input_override, int* launches);
  void mark_written(int tensor);
  void ensure_device();
  void* dalloc(size_t bytes, bool zero);
  Tensor& T(int i) { return tensors_[i]; }
  int add_tensor(const std::string& name);
  void plan_conv_group(int li, std::vector<bool>& done, int in_override = -1);
  void fuse_sibling_1x1();
  bool try_commute_pool_conv(int li, std::vector<bool>& done, std::vector<int>& view_of);
  void make_tensor_maps(ConvOp& c);
  bool plan_halo(ConvOp& c);
  void plan_stem_rows(ConvOp& c);
  bool is_stem_conv(const OrigLayer& L) const;
  unsigned char* xf_src_ = nullptr;   // staged uint8 clips, transforms and means of transform_input_u8
  size_t xf_src_bytes_ = 0;
  float* stage_ = nullptr;
  size_t stage_bytes_ = 0;
  float* staging(size_t bytes);
  // training state
  bool train_ = false;
  float* P_ = nullptr;               // parameter arena
  float* G_ = nullptr;               // gradient arena
  size_t arena_count_ = 0;
  std::vector<ParamSlot> slots_;
  std::map<std::pair<int, int>, int> slot_index_;  // (orig layer, blob) -> slot
  std::vector<TrainAux> aux_;        // parallel to ops_
  std::vector<ConvOp> dgrads_;
  float* wgrad_scratch_ = nullptr;
  std::vector<char> bwd_seeded_;      // per backward() call: storage roots whose gradient the caller seeded from the host
  bool bwd_whole_ = false;
  unsigned char* pool_mask_ = nullptr;   // first-maximum indices of the MAX pooling being back-propagated
  float* reduce_scratch_ = nullptr;  // per-block partials of the deterministic per-channel reductions
  size_t wgrad_scratch_bytes_ = 0;
  bool params_dev_newer_ = false;    // arena newer than the host ParamBlobs (solver update, BN running statistics)
  bool repack_ = true;               // bf16 GEMM operands must be rebuilt from the arena
  unsigned long long train_iter_ = 0;
  bool prof_on_ = false;
  std::vector<std::pair<std::string, cudaEvent_t>> prof_marks_;
  std::vector<std::string> prof_names_;
  void prof_mark(const std::string& name);
  int bucket_n_ = 0;
  BucketFn bucket_fn_ = nullptr;
  void* bucket_user_ = nullptr;
  std::vector<Bucket> buckets_;
  float last_loss_ = 0.f;
  float* loss_host_ = nullptr;       // pinned
  void plan_train();
  void upload_params_train();
  void sync_params_to_host();
  void setup_dgrad(Op& op, TrainAux& a);
  void setup_wgrad(Op& op, TrainAux& a);
  void run_train_op(Op& op, TrainAux& a);
  void backward_op(size_t i, std::vector<char>& written);
  ClView dview(const Tensor& t) const;
  float* slot_ptr(float* arena, int layer, int idx);
  void download(Tensor& t);
  void upload(Tensor& t);
  ClView view(const Tensor& t) const;
};

void set_last_error(const std::string& s);

}  // namespace eco
