// eco_caffe_facade.hpp -- header-only C++ facade with caffe_3d's class and method names over the
// C ABI (eco_b200.h), for C++ callers written against
//   caffe::Net<float>   caffe_3d/include/caffe/net.hpp:24-281
//   caffe::Blob<float>  caffe_3d/include/caffe/blob.hpp:25-282
//   caffe::Caffe        caffe_3d/include/caffe/common.hpp:122-199
// (e.g. tools/caffe.cpp `time` / `test`, tools/extract_features.cpp).  Only the forward-path surface
// is provided.  Error behaviour follows caffe: a failed call prints the message and aborts
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
class Blob {
 public:
  const std::vector<int>& shape() const {
    int dims[8], nd = 8;
    detail::check(eco_net_blob_shape(net_, index_, dims, &nd), "Blob::shape");
    shape_.assign(dims, dims + nd);
    return shape_;
  }
  int shape(int i) const { const auto& s = shape(); return s[i < 0 ? i + (int)s.size() : i]; }
  int num_axes() const { return (int)shape().size(); }
  int count() const { int c = 1; for (int d : shape()) c *= d; return c; }
  int count(int a, int b) const { const auto& s = shape(); int c = 1; for (int i = a; i < b; ++i) c *= s[i]; return c; }
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
  const Dtype* cpu_data() const {
    float* p; size_t n;
    detail::check(eco_blob_host_data(net_, index_, 0, &p, &n), "Blob::cpu_data");
    return p;
  }
  Dtype* mutable_cpu_data() {
    float* p; size_t n;
    detail::check(eco_blob_host_data(net_, index_, 1, &p, &n), "Blob::mutable_cpu_data");
    return p;
  }
  const Dtype* cpu_diff() const {
    float* p; size_t n;
    detail::check(eco_blob_host_diff(net_, index_, 0, &p, &n), "Blob::cpu_diff");
    return p;
  }
  void Reshape(const std::vector<int>& shape) {
    detail::check(eco_blob_reshape(net_, index_, shape.data(), (int)shape.size()), "Blob::Reshape");
  }
  void Reshape(int n, int c, int h, int w) { Reshape(std::vector<int>{n, c, h, w}); }

 private:
  friend class Net<Dtype>;
  Blob(eco_net* net, int index) : net_(net), index_(index) {}
  eco_net* net_;
  int index_;
  mutable std::vector<int> shape_;
};

template <typename Dtype>
class Net {
  static_assert(sizeof(Dtype) == sizeof(float), "libeco_b200 exposes fp32 blobs (Net<float>)");

 public:
  Net(const std::string& param_file, Phase phase) {
    detail::check(eco_net_create(param_file.c_str(), (int)phase, &h_), "Net::Net");
    refresh();
  }
  ~Net() { eco_net_destroy(h_); }
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
  void Reshape() { detail::check(eco_net_reshape(h_), "Net::Reshape"); }
  void CopyTrainedLayersFrom(const std::string& trained_filename) {
    detail::check(eco_net_copy_from(h_, trained_filename.c_str()), "Net::CopyTrainedLayersFrom");
  }
  void Backward() { detail::check(eco_net_backward(h_, (int)layer_names_.size() - 1, 0), "Net::Backward"); }
  eco_net* handle() { return h_; }

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
  }
  eco_net* h_ = nullptr;
  std::string name_;
  std::vector<std::string> layer_names_, blob_names_;
  std::vector<std::shared_ptr<Blob<Dtype>>> blobs_;
  std::vector<Blob<Dtype>*> inputs_, outputs_;
};

}  // namespace caffe
