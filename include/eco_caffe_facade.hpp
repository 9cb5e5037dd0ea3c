// eco_caffe_facade.hpp -- header-only C++ facade with caffe_3d's class and method names over the
// C ABI (eco_b200.h), for C++ callers written against
//   caffe::Net<float>   caffe_3d/include/caffe/net.hpp:24-281
//   caffe::Blob<float>  caffe_3d/include/caffe/blob.hpp:25-282
//   caffe::Caffe        caffe_3d/include/caffe/common.hpp:122-199
//   caffe::Layer<float> caffe_3d/include/caffe/layer.hpp (blobs(), type(), layer_param().name())
//   caffe::Solver / SGDSolver / NesterovSolver  caffe_3d/include/caffe/solver.hpp
// (e.g. tools/caffe.cpp `train` / `time` / `test`, tools/extract_features.cpp).  Error behaviour follows caffe: a failed call prints the message and aborts
// (glog CHECK / LOG(FATAL) semantics); define ECO_FACADE_THROW to get std::runtime_error instead.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "eco_b200.h"

namespace caffe {

enum Phase { TRAIN = ECO_PHASE_TRAIN, TEST = ECO_PHASE_TEST };

namespace detail {
inline void check(int rc, const char* what) {
  if (rc == 0) return;
#ifdef ECO_FACADE_THROW
  throw std::runtime_error(std::string(what) + ": " + eco_last_error());
#else
  std::fprintf(stderr, "F %s: %s\n*** Check failure stack trace: ***\n", what, eco_last_error());
  std::abort();
#endif
}
}  // namespace detail

class Caffe {
 public:
  enum Brew { CPU, GPU };
  static void set_mode(Brew mode) { detail::check(eco_set_mode(mode == GPU ? 1 : 0), "Caffe::set_mode"); }
  static void SetDevice(const int device_id) { detail::check(eco_set_device(device_id), "Caffe::SetDevice"); }
  static void set_random_seed(unsigned int) {}
};

template <typename Dtype>
class Net;

template <typename Dtype>
class Layer;

// A Blob is a view of either a net blob (index into Net::blobs()) or a layer parameter blob (layer, k).
template <typename Dtype>
class Blob {
 public:
  const std::vector<int>& shape() const {
    int dims[8], nd = 8;
    if (layer_ < 0) detail::check(eco_net_blob_shape(net_, index_, dims, &nd), "Blob::shape");
    else detail::check(eco_net_param_shape(net_, layer_, index_, dims, &nd), "Blob::shape");
    shape_.assign(dims, dims + nd);
    return shape_;
  }
  int shape(int i) const { const auto& s = shape(); return s[i < 0 ? i + (int)s.size() : i]; }
  int num_axes() const { return (int)shape().size(); }
  int count() const { int c = 1; for (int d : shape()) c *= d; return c; }
  int count(int a, int b) const { const auto& s = shape(); int c = 1; for (int i = a; i < b; ++i) c *= s[i]; return c; }
  int count(int a) const { return count(a, num_axes()); }
  // legacy 4-D accessors: FATAL for >4 axes exactly as blob.hpp:133-152
  int LegacyShape(int i) const {
    const auto& s = shape();
    if (s.size() > 4) detail::check(1, "Cannot use legacy accessors on Blobs with > 4 axes.");
    return i < (int)s.size() ? s[i] : 1;
  }
  int num() const { return LegacyShape(0); }
  int channels() const { return LegacyShape(1); }
  int height() const { return LegacyShape(2); }
  int width() const { return LegacyShape(3); }
  int offset(const int n, const int c = 0, const int h = 0, const int w = 0) const {  // blob.hpp:154-165
    return ((n * channels() + c) * height() + h) * width() + w;
  }
  const Dtype* cpu_data() const { return host(0, false); }
  Dtype* mutable_cpu_data() { return host(1, false); }
  const Dtype* cpu_diff() const { return host(0, true); }
  Dtype* mutable_cpu_diff() { return host(1, true); }
  Dtype data_at(const int n, const int c, const int h, const int w) const { return cpu_data()[offset(n, c, h, w)]; }
  Dtype diff_at(const int n, const int c, const int h, const int w) const { return cpu_diff()[offset(n, c, h, w)]; }
  // device pointer: only the plain fp32 blobs (pooled vectors, logits) have caffe's layout on the device; feature maps are
  // bf16 channels-last there, so a caller that wants their values goes through cpu_data() like any caffe tool does
  const Dtype* gpu_data() const {
    if (layer_ >= 0) detail::check(1, "gpu_data() of a parameter blob: use the arena (eco_net_param_arena)");
    const float* p = nullptr; size_t n = 0;
    detail::check(eco_blob_device_f32(net_, index_, &p, &n), "Blob::gpu_data");
    return p;
  }
  void Reshape(const std::vector<int>& shape) {
    if (layer_ >= 0) detail::check(1, "Reshape of a parameter blob is not supported");
    detail::check(eco_blob_reshape(net_, index_, shape.data(), (int)shape.size()), "Blob::Reshape");
  }
  void Reshape(int n, int c, int h, int w) { Reshape(std::vector<int>{n, c, h, w}); }
  void ReshapeLike(const Blob& other) { Reshape(other.shape()); }
  void CopyFrom(const Blob& source, bool copy_diff = false, bool reshape = false) {  // blob.cpp:439-469
    if (source.count() != count() || source.shape() != shape()) {
      if (reshape) ReshapeLike(source);
      else detail::check(1, "Trying to copy blobs of different sizes.");
    }
    const Dtype* src = copy_diff ? source.cpu_diff() : source.cpu_data();
    Dtype* dst = copy_diff ? mutable_cpu_diff() : mutable_cpu_data();
    for (int i = 0; i < count(); ++i) dst[i] = src[i];
  }
  Dtype asum_data() const { const Dtype* p = cpu_data(); double a = 0; for (int i = 0; i < count(); ++i) a += p[i] < 0 ? -p[i] : p[i]; return (Dtype)a; }
  Dtype asum_diff() const { const Dtype* p = cpu_diff(); double a = 0; for (int i = 0; i < count(); ++i) a += p[i] < 0 ? -p[i] : p[i]; return (Dtype)a; }
  Dtype sumsq_data() const { const Dtype* p = cpu_data(); double a = 0; for (int i = 0; i < count(); ++i) a += (double)p[i] * p[i]; return (Dtype)a; }
  Dtype sumsq_diff() const { const Dtype* p = cpu_diff(); double a = 0; for (int i = 0; i < count(); ++i) a += (double)p[i] * p[i]; return (Dtype)a; }
  void scale_data(Dtype f) { Dtype* p = mutable_cpu_data(); for (int i = 0; i < count(); ++i) p[i] *= f; }

 private:
  friend class Net<Dtype>;
  friend class Layer<Dtype>;
  Blob(eco_net* net, int index, int layer = -1) : net_(net), index_(index), layer_(layer) {}
  Dtype* host(int for_write, bool diff) const {
    float* p = nullptr; size_t n = 0;
    if (layer_ < 0) {
      detail::check(diff ? eco_blob_host_diff(net_, index_, for_write, &p, &n) : eco_blob_host_data(net_, index_, for_write, &p, &n),
                    diff ? "Blob::cpu_diff" : "Blob::cpu_data");
    } else if (diff) {
      detail::check(eco_net_param_diff_host(net_, layer_, index_, &p, &n), "Blob::cpu_diff");
    } else if (for_write) {
      detail::check(eco_net_param_host(net_, layer_, index_, &p, &n), "Blob::mutable_cpu_data");
    } else {
      // read-only view of a parameter: a scratch copy, so reading does not mark the layer dirty
      shadow_.resize((size_t)count());
      detail::check(eco_net_get_param(net_, layer_, index_, shadow_.data(), shadow_.size()), "Blob::cpu_data");
      p = shadow_.data();
    }
    return p;
  }
  eco_net* net_;
  int index_;
  int layer_;  // >= 0: parameter blob `index_` of that layer
  mutable std::vector<int> shape_;
  mutable std::vector<float> shadow_;
};

template <typename Dtype>
class Layer {
 public:
  const char* type() const { return eco_net_layer_type(net_, index_); }
  const std::string& name() const { return name_; }
  std::vector<std::shared_ptr<Blob<Dtype>>>& blobs() { return blobs_; }

 private:
  friend class Net<Dtype>;
  Layer(eco_net* net, int index) : net_(net), index_(index), name_(eco_net_layer_name(net, index)) {
    int n = 0;
    detail::check(eco_net_layer_num_params(net, index, &n), "Layer::blobs");
    for (int k = 0; k < n; ++k) blobs_.push_back(std::shared_ptr<Blob<Dtype>>(new Blob<Dtype>(net, k, index)));
  }
  eco_net* net_;
  int index_;
  std::string name_;
  std::vector<std::shared_ptr<Blob<Dtype>>> blobs_;
};

template <typename Dtype>
class Net {
  static_assert(sizeof(Dtype) == sizeof(float), "libeco_b200 exposes fp32 blobs (Net<float>)");

 public:
  Net(const std::string& param_file, Phase phase) {
    detail::check(eco_net_create(param_file.c_str(), (int)phase, &h_), "Net::Net");
    refresh();
  }
  ~Net() { if (owned_) eco_net_destroy(h_); }
  Net(const Net&) = delete;
  Net& operator=(const Net&) = delete;

  const std::string& name() const { return name_; }
  const std::vector<std::string>& layer_names() const { return layer_names_; }
  const std::vector<std::string>& blob_names() const { return blob_names_; }
  const std::vector<std::shared_ptr<Blob<Dtype>>>& blobs() const { return blobs_; }
  bool has_blob(const std::string& n) const { return eco_net_blob_index(h_, n.c_str()) >= 0; }
  bool has_layer(const std::string& n) const { return eco_net_layer_index(h_, n.c_str()) >= 0; }
  const std::shared_ptr<Blob<Dtype>> blob_by_name(const std::string& n) const {
    const int i = eco_net_blob_index(h_, n.c_str());
    return i < 0 ? nullptr : blobs_[i];  // caffe logs "Unknown blob name" and returns NULL (net.cpp:958-968)
  }
  const std::vector<std::shared_ptr<Layer<Dtype>>>& layers() const { return layers_; }
  const std::shared_ptr<Layer<Dtype>> layer_by_name(const std::string& n) const {
    const int i = eco_net_layer_index(h_, n.c_str());
    return i < 0 ? nullptr : layers_[i];
  }
  const std::vector<std::shared_ptr<Blob<Dtype>>>& params() const { return params_; }   // learnable blobs in layer order
  const std::vector<float>& params_lr() const { return params_lr_; }
  const std::vector<float>& params_weight_decay() const { return params_decay_; }
  const std::vector<Blob<Dtype>*>& input_blobs() const { return inputs_; }
  const std::vector<Blob<Dtype>*>& output_blobs() const { return outputs_; }
  int num_inputs() const { return (int)inputs_.size(); }
  int num_outputs() const { return (int)outputs_.size(); }

  Dtype ForwardFromTo(int start, int end) {
    float loss = 0;
    detail::check(eco_net_forward(h_, start, end, &loss), "Net::ForwardFromTo");
    return loss;
  }
  const std::vector<Blob<Dtype>*>& ForwardPrefilled(Dtype* loss = NULL) {
    const Dtype l = ForwardFromTo(0, (int)layer_names_.size() - 1);
    if (loss) *loss = l;
    return outputs_;
  }
  const std::vector<Blob<Dtype>*>& Forward(Dtype* loss = NULL) { return ForwardPrefilled(loss); }
  const std::vector<Blob<Dtype>*>& Forward(const std::vector<Blob<Dtype>*>& bottom, Dtype* loss = NULL) {  // net.cpp:596-605
    for (size_t i = 0; i < bottom.size() && i < inputs_.size(); ++i) inputs_[i]->CopyFrom(*bottom[i]);
    return ForwardPrefilled(loss);
  }
  Dtype ForwardFrom(int start) { return ForwardFromTo(start, (int)layer_names_.size() - 1); }
  Dtype ForwardTo(int end) { return ForwardFromTo(0, end); }
  void BackwardFromTo(int start, int end) { detail::check(eco_net_backward(h_, start, end), "Net::BackwardFromTo"); }
  void BackwardFrom(int start) { BackwardFromTo(start, 0); }
  void BackwardTo(int end) { BackwardFromTo((int)layer_names_.size() - 1, end); }
  Dtype ForwardBackward(const std::vector<Blob<Dtype>*>& bottom) {   // net.hpp:86-91
    Dtype loss;
    Forward(bottom, &loss);
    Backward();
    return loss;
  }
  void ClearParamDiffs() { detail::check(eco_net_clear_param_diffs(h_), "Net::ClearParamDiffs"); }
  void Update() { detail::check(eco_net_update(h_), "Net::Update"); }     // data -= diff (net.cpp:906-910)
  void ShareTrainedLayersWith(const Net* other) {   // values are copied once, matched by layer name (storage is not shared)
    for (auto& l : layers_) {
      auto src = other->layer_by_name(l->name());
      if (!src) continue;
      for (size_t k = 0; k < l->blobs().size() && k < src->blobs().size(); ++k) l->blobs()[k]->CopyFrom(*src->blobs()[k]);
    }
  }
  void ToProto(const std::string& caffemodel_path) const { detail::check(eco_net_save(h_, caffemodel_path.c_str()), "Net::ToProto"); }
  void Reshape() { detail::check(eco_net_reshape(h_), "Net::Reshape"); }
  void CopyTrainedLayersFrom(const std::string& trained_filename) {
    detail::check(eco_net_copy_from(h_, trained_filename.c_str()), "Net::CopyTrainedLayersFrom");
  }
  void Backward() { detail::check(eco_net_backward(h_, (int)layer_names_.size() - 1, 0), "Net::Backward"); }
  eco_net* handle() { return h_; }
  Net(eco_net* borrowed, bool) : h_(borrowed), owned_(false) { refresh(); }   // a solver's net

 private:
  void refresh() {
    name_ = eco_net_name(h_);
    for (int i = 0; i < eco_net_num_layers(h_); ++i) layer_names_.push_back(eco_net_layer_name(h_, i));
    for (int i = 0; i < eco_net_num_blobs(h_); ++i) {
      blob_names_.push_back(eco_net_blob_name(h_, i));
      blobs_.push_back(std::shared_ptr<Blob<Dtype>>(new Blob<Dtype>(h_, i)));
    }
    for (int i = 0; i < eco_net_num_inputs(h_); ++i) inputs_.push_back(blobs_[eco_net_input_blob(h_, i)].get());
    for (int i = 0; i < eco_net_num_outputs(h_); ++i) outputs_.push_back(blobs_[eco_net_output_blob(h_, i)].get());
    for (int i = 0; i < (int)layer_names_.size(); ++i) {
      layers_.push_back(std::shared_ptr<Layer<Dtype>>(new Layer<Dtype>(h_, i)));
      for (auto& b : layers_.back()->blobs()) { params_.push_back(b); params_lr_.push_back(1.f); params_decay_.push_back(1.f); }
    }
  }
  eco_net* h_ = nullptr;
  bool owned_ = true;
  std::vector<std::shared_ptr<Layer<Dtype>>> layers_;
  std::vector<std::shared_ptr<Blob<Dtype>>> params_;
  std::vector<float> params_lr_, params_decay_;
  std::string name_;
  std::vector<std::string> layer_names_, blob_names_;
  std::vector<std::shared_ptr<Blob<Dtype>>> blobs_;
  std::vector<Blob<Dtype>*> inputs_, outputs_;
};

// caffe::Solver<float> / SGDSolver / NesterovSolver (solver.hpp): the solver_type of the file selects the update rule
template <typename Dtype>
class Solver {
 public:
  explicit Solver(const std::string& param_file) {
    detail::check(eco_solver_create(param_file.c_str(), &h_), "Solver::Solver");
    eco_net* n = nullptr;
    detail::check(eco_solver_net(h_, &n), "Solver::net");
    net_.reset(new Net<Dtype>(n, false));
  }
  ~Solver() { net_.reset(); eco_solver_destroy(h_); }
  Solver(const Solver&) = delete;
  Solver& operator=(const Solver&) = delete;
  std::shared_ptr<Net<Dtype>> net() { return net_; }
  int iter() const { int i = 0; detail::check(eco_solver_iter(h_, &i), "Solver::iter"); return i; }
  Dtype Step(int iters) { float l = 0; detail::check(eco_solver_step(h_, iters, &l), "Solver::Step"); return l; }
  void Snapshot() { detail::check(eco_solver_snapshot(h_, nullptr), "Solver::Snapshot"); }
  void Restore(const char* resume_file) { detail::check(eco_solver_restore(h_, resume_file), "Solver::Restore"); }
  eco_solver* handle() { return h_; }

 private:
  eco_solver* h_ = nullptr;
  std::shared_ptr<Net<Dtype>> net_;
};
template <typename Dtype> using SGDSolver = Solver<Dtype>;
template <typename Dtype> using NesterovSolver = Solver<Dtype>;

}  // namespace caffe
