// net.cpp -- graph construction (caffe-visible names), shape inference, the fusing planner and
// the executor.  Reference behaviour followed, by function:
//   build_graph      Net::Init / FilterNet / InsertSplits   caffe_3d/src/caffe/net.cpp:39-316,319-346;
//                                                           util/insert_splits.cpp:13-142
//   infer_shapes     each layer's Reshape: conv_layer.cpp:12-25, pooling_layer.cpp:117-163,
//                    concat_layer.cpp:17-50, reshape_layer.cpp:31-90, permute_layer.cpp:29-73,
//                    inner_product_layer.cpp:13-77, bn_layer.cpp:11-90
//   init_params      fillers (include/caffe/filler.hpp) -- values are NOT bit-identical to caffe's
//                    boost RNG; the parity harness always writes its own weights
//   plan / run_op    replaces Layer::Forward dispatch (layer.hpp:444-477) with fused sm_100a ops
#include "net.hpp"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <set>
#include <sstream>
#include <stdexcept>

#include "../../include/eco_b200.h"
#include "transform.cuh"

namespace eco {

#define ECO_CHECK(cond, msg)                                                       \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      std::ostringstream _o;                                                       \
      _o << msg << "  [" << #cond << " @ " << __FILE__ << ":" << __LINE__ << "]"; \
      throw std::runtime_error(_o.str());                                          \
    }                                                                              \
  } while (0)

#define CUDA_OK(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      std::ostringstream _o;                                                                       \
      _o << "CUDA error " << cudaGetErrorName(_e) << ": " << cudaGetErrorString(_e) << " in " #expr \
         << " @ " << __FILE__ << ":" << __LINE__;                                                  \
      throw std::runtime_error(_o.str());                                                          \
    }                                                                                              \
  } while (0)

static int g_device_ok = -1;  // -1 unknown, 0 none, 1 ok
static int g_num_sms = 148;
static bool device_available() {
  if (g_device_ok < 0) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    g_device_ok = n > 0 ? 1 : 0;
  }
  return g_device_ok == 1;
}

void HostBuf::release() {
  if (p) {
    if (pinned) cudaFreeHost(p);
    else std::free(p);
  }
  p = nullptr;
  n = 0;
  pinned = false;
}
void HostBuf::resize(size_t count, bool try_pin) {
  if (count == n && p) return;
  release();
  if (count == 0) return;
  if (try_pin && device_available()) {
    void* q = nullptr;
    if (cudaMallocHost(&q, count * sizeof(float)) == cudaSuccess) {
      p = static_cast<float*>(q);
      pinned = true;
    } else {
      cudaGetLastError();
    }
  }
  if (!p) p = static_cast<float*>(std::malloc(count * sizeof(float)));
  ECO_CHECK(p != nullptr, "host allocation of " << count * 4 << " bytes failed");
  std::memset(p, 0, count * sizeof(float));
  n = count;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }
static int pow2_at_least(int v);

static bool is_data_layer(const std::string& t) {
  return t == "VideoData" || t == "Data" || t == "ImageData" || t == "MemoryData" || t == "Input" ||
         t == "DummyData" || t == "HDF5Data" || t == "WindowData" || t == "SegData";
}

// ---- FilterNet restricted to phase rules (net.cpp:319-346, StateMeetsRule :349-411) ----
static bool rule_ok(const pt::Msg& rule, const std::string& phase) {
  if (rule.has("phase") && rule.str("phase") != phase) return false;
  return true;  // level / stage rules are not used by models_ECO_*
}
static bool layer_in_phase(const pt::Msg& l, int phase) {
  const std::string ph = phase == ECO_PHASE_TRAIN ? "TRAIN" : "TEST";
  auto inc = l.all("include");
  auto exc = l.all("exclude");
  if (!inc.empty()) {
    for (auto* r : inc)
      if (r->msg && rule_ok(*r->msg, ph)) return true;
    return false;
  }
  for (auto* r : exc)
    if (r->msg && r->msg->has("phase") && r->msg->str("phase") == ph) return false;
  return true;
}

// kernel_size / stride / pad: once, once per axis, or the 2-D *_h/*_w forms
// (base_conv_layer.cpp:13-110, pooling_layer.cpp:17-114)
static std::vector<int> nd_param(const pt::Msg& p, const char* key, const char* hw, int nsp, int def) {
  std::string h = std::string(hw) + "_h", w = std::string(hw) + "_w";
  if (p.has(h) || p.has(w)) {
    ECO_CHECK(nsp == 2, hw << "_h/_w can only be used for 2-D layers");
    return {(int)p.integer(h, def), (int)p.integer(w, def)};
  }
  auto v = p.integers(key);
  if (v.empty()) {
    ECO_CHECK(def >= 0, key << " must be specified");
    return std::vector<int>(nsp, def);
  }
  if (v.size() == 1) return std::vector<int>(nsp, (int)v[0]);
  ECO_CHECK((int)v.size() == nsp, key << " specified " << v.size() << " times for " << nsp << " spatial axes");
  return std::vector<int>(v.begin(), v.end());
}

static int pool_out_dim(int in, int k, int s, int p) {
  int out = (int)std::ceil((float)(in + 2 * p - k) / (float)s) + 1;  // pooling_layer.cpp:131-136
  if (p && (out - 1) * s >= in + p) --out;                           // :137-147
  return out;
}

// =====================================================================================
Net::Net(const std::string& text, int phase, const std::string& until_blob) : phase_(phase), until_blob_(until_blob), text_(text) {
  proto_ = pt::parse(text);
  build_graph();
  infer_shapes();
  init_params();
}

void Net::release_subs() {
  for (auto& s : sub_)
    if (s && s->stream_) cudaStreamSynchronize(s->stream_);
  sub_.clear();
  for (auto e : sub_done_) cudaEventDestroy(e);
  sub_done_.clear();
  sub_params_version_ = ~0ull;
}

Net::~Net() {
  release_subs();
  for (auto& t : tensors_)
    if (t.h2d_done) cudaEventDestroy(t.h2d_done);
  free_plan();
  if (stage_) cudaFree(stage_);
  if (xf_src_) cudaFree(xf_src_);
  if (push_event_) cudaEventDestroy(push_event_);
  if (own_stream_ && stream_) cudaStreamDestroy(stream_);
}

int Net::add_tensor(const std::string& name) {
  auto it = tensor_index_.find(name);
  if (it != tensor_index_.end()) return it->second;
  Tensor t;
  t.name = name;
  tensors_.push_back(std::move(t));
  tensor_index_[name] = (int)tensors_.size() - 1;
  return (int)tensors_.size() - 1;
}

void Net::build_graph() {
  const pt::Msg& net = *proto_;
  name_ = net.str("name");
  ECO_CHECK(!net.has("layers"), "V1 'layers' net definitions are not supported (upgrade_net_proto_text first)");

  // ---- net inputs: `input:` + `input_dim:` x4 (deprecated form, net.cpp:55-73) or `input_shape` ----
  std::vector<std::string> in_names = net.strs("input");
  std::vector<long> in_dims = net.integers("input_dim");
  auto in_shapes = net.all("input_shape");
  std::vector<std::pair<std::string, std::vector<int>>> net_inputs;
  for (size_t i = 0; i < in_names.size(); ++i) {
    std::vector<int> shp;
    if (!in_shapes.empty()) {
      ECO_CHECK(i < in_shapes.size() && in_shapes[i]->msg, "input_shape missing for input " << in_names[i]);
      for (long d : in_shapes[i]->msg->integers("dim")) shp.push_back((int)d);
    } else {
      ECO_CHECK(in_dims.size() >= 4 * (i + 1), "input_dim must be given 4 times per input");
      for (int k = 0; k < 4; ++k) shp.push_back((int)in_dims[4 * i + k]);
    }
    net_inputs.emplace_back(in_names[i], shp);
  }

  // ---- layers of this phase ----
  std::vector<const pt::Msg*> lmsgs;
  for (auto* v : net.all("layer"))
    if (v->msg && layer_in_phase(*v->msg, phase_)) lmsgs.push_back(v->msg.get());

  if (!until_blob_.empty()) {
    int last = -1;
    for (size_t i = 0; i < lmsgs.size(); ++i)
      for (auto& t : lmsgs[i]->strs("top"))
        if (t == until_blob_) last = (int)i;
    ECO_CHECK(last >= 0, "no layer produces blob '" << until_blob_ << "'");
    lmsgs.resize((size_t)last + 1);
  }
  for (auto& ni : net_inputs) {
    int t = add_tensor(ni.first);
    tensors_[t].shape = ni.second;
  }
  for (const pt::Msg* m : lmsgs) {
    OrigLayer L;
    L.name = m->str("name");
    L.type = m->str("type");
    L.msg = m;
    for (auto& b : m->strs("bottom")) {
      auto it = tensor_index_.find(b);
      ECO_CHECK(it != tensor_index_.end(),
                "Unknown bottom blob '" << b << "' (layer '" << L.name << "')");  // net.cpp:428-431
      L.bottoms.push_back(it->second);
    }
    for (auto& t : m->strs("top")) L.tops.push_back(add_tensor(t));
    layers_.push_back(std::move(L));
  }
  // data layers feed their tops from the host: treat like net inputs with a declared shape
  for (size_t li = 0; li < layers_.size(); ++li) {
    OrigLayer& L = layers_[li];
    if (!is_data_layer(L.type)) continue;
    const pt::Msg* vp = L.msg->msg("video_data_param");
    const pt::Msg* tp = L.msg->msg("transform_param");
    int batch = vp ? (int)vp->integer("batch_size", 1) : 1;
    int segs = vp ? (int)vp->integer("num_segments", 1) : 1;
    int len = vp ? (int)vp->integer("new_length", 1) : 1;
    int crop = tp ? (int)tp->integer("crop_size", 224) : 224;
    std::string modality = vp ? vp->str("modality", "RGB") : "RGB";
    int cpf = modality == "FLOW" ? 2 : 3;  // video_data_layer.cpp: channels per frame
    if (L.tops.size() >= 1) tensors_[L.tops[0]].shape = {batch, cpf * segs * len, crop, crop};
    if (L.tops.size() >= 2) tensors_[L.tops[1]].shape = {batch, 1, 1, 1};
  }
  // producers / consumers
  for (size_t li = 0; li < layers_.size(); ++li) {
    for (int b : layers_[li].bottoms) tensors_[b].consumers.push_back((int)li);
    for (int t : layers_[li].tops) tensors_[t].producer = (int)li;
  }

  // ---- visible registry: replicate InsertSplits (insert_splits.cpp:13-142) ----
  typedef std::pair<int, int> TopRef;  // (layer idx or -1 for net input, top idx)
  std::map<std::string, TopRef> last_top;
  std::map<TopRef, int> use_count;
  std::map<std::pair<int, int>, TopRef> bottom_src;
  for (size_t i = 0; i < net_inputs.size(); ++i) last_top[net_inputs[i].first] = TopRef(-1, (int)i);
  for (size_t li = 0; li < layers_.size(); ++li) {
    const OrigLayer& L = layers_[li];
    for (size_t j = 0; j < L.bottoms.size(); ++j) {
      const TopRef src = last_top[tensors_[L.bottoms[j]].name];
      bottom_src[{(int)li, (int)j}] = src;
      use_count[src]++;
    }
    for (size_t j = 0; j < L.tops.size(); ++j) last_top[tensors_[L.tops[j]].name] = TopRef((int)li, (int)j);
  }
  auto vis_blob = [&](const std::string& nm, int tensor) {
    auto it = vis_blob_index_.find(nm);
    if (it != vis_blob_index_.end()) return it->second;
    vis_blobs_.push_back({nm, tensor});
    vis_blob_index_[nm] = (int)vis_blobs_.size() - 1;
    return (int)vis_blobs_.size() - 1;
  };
  auto split_layer_name = [](const std::string& layer, const std::string& blob, int idx) {
    return blob + "_" + layer + "_" + std::to_string(idx) + "_split";
  };
  std::map<TopRef, int> next_split;
  auto add_split = [&](const std::string& layer_name, const std::string& blob, int top_idx, int tensor, int count) {
    VisLayer S;
    S.name = split_layer_name(layer_name, blob, top_idx);
    S.type = "Split";
    S.bottoms.push_back(vis_blob(blob, tensor));
    for (int k = 0; k < count; ++k) S.tops.push_back(vis_blob(S.name + "_" + std::to_string(k), tensor));
    vis_layer_index_[S.name] = (int)vis_layers_.size();
    vis_layers_.push_back(std::move(S));
  };
  for (size_t i = 0; i < net_inputs.size(); ++i) {
    int t = tensor_index_[net_inputs[i].first];
    inputs_.push_back(vis_blob(net_inputs[i].first, t));
  }
  for (size_t i = 0; i < net_inputs.size(); ++i) {
    const TopRef r(-1, (int)i);
    if (use_count[r] > 1) add_split("input", net_inputs[i].first, (int)i, tensor_index_[net_inputs[i].first], use_count[r]);
  }
  for (size_t li = 0; li < layers_.size(); ++li) {
    OrigLayer& L = layers_[li];
    VisLayer V;
    V.name = L.name;
    V.type = L.type;
    V.orig = (int)li;
    for (size_t j = 0; j < L.bottoms.size(); ++j) {
      const TopRef src = bottom_src[{(int)li, (int)j}];
      const std::string& blob = tensors_[L.bottoms[j]].name;
      if (use_count[src] > 1) {
        const std::string src_layer = src.first < 0 ? "input" : layers_[src.first].name;
        const int k = next_split[src]++;
        const std::string nm = split_layer_name(src_layer, blob, src.second) + "_" + std::to_string(k);
        V.bottoms.push_back(vis_blob(nm, L.bottoms[j]));
      } else {
        V.bottoms.push_back(vis_blob(blob, L.bottoms[j]));
      }
    }
    for (size_t j = 0; j < L.tops.size(); ++j) V.tops.push_back(vis_blob(tensors_[L.tops[j]].name, L.tops[j]));
    L.vis_index = (int)vis_layers_.size();
    vis_layer_index_[V.name] = (int)vis_layers_.size();
    vis_layers_.push_back(std::move(V));
    for (size_t j = 0; j < L.tops.size(); ++j) {
      const TopRef r((int)li, (int)j);
      if (use_count[r] > 1) add_split(L.name, tensors_[L.tops[j]].name, (int)j, L.tops[j], use_count[r]);
    }
    if (is_data_layer(L.type))
      for (int t : L.tops) inputs_.push_back(vis_blob_index_[tensors_[t].name]);
  }
  // net outputs: blobs nobody consumes (net.cpp:283-291)
  {
    std::set<std::string> avail;
    std::vector<std::string> order;
    for (auto& ni : net_inputs) { avail.insert(ni.first); order.push_back(ni.first); }
    for (auto& V : vis_layers_) {
      for (int b : V.bottoms) avail.erase(vis_blobs_[b].name);
      for (int t : V.tops)
        if (avail.insert(vis_blobs_[t].name).second) order.push_back(vis_blobs_[t].name);
    }
    for (auto& nm : order)
      if (avail.count(nm)) { outputs_.push_back(vis_blob_index_[nm]); avail.erase(nm); }
  }
}

// =====================================================================================
void Net::infer_shapes() {
  for (size_t li = 0; li < layers_.size(); ++li) {
    OrigLayer& L = layers_[li];
    const std::string& t = L.type;
    if (is_data_layer(t)) continue;
    auto bshape = [&](int j) -> const std::vector<int>& {
      ECO_CHECK(j < (int)L.bottoms.size(), "layer " << L.name << " needs bottom " << j);
      return tensors_[L.bottoms[j]].shape;
    };
    auto set_top = [&](int j, const std::vector<int>& s) {
      ECO_CHECK(j < (int)L.tops.size(), "layer " << L.name << " needs top " << j);
      tensors_[L.tops[j]].shape = s;
    };
    auto set_param_shape = [&](size_t idx, const std::vector<int>& s) {
      if (L.params.size() <= idx) L.params.resize(idx + 1);
      if (L.params[idx].shape != s) {
        ECO_CHECK(L.params[idx].shape.empty(), "layer " << L.name << ": parameter shape changes on reshape");
        L.params[idx].shape = s;
      }
    };
    if (t == "Convolution") {
      const pt::Msg* p = L.msg->msg("convolution_param");
      ECO_CHECK(p, "convolution_param missing in " << L.name);
      const auto& b = bshape(0);
      const int nsp = (int)b.size() - 2;
      ECO_CHECK(nsp >= 1 && nsp <= 3, "Convolution " << L.name << ": " << nsp << " spatial axes unsupported");
      auto k = nd_param(*p, "kernel_size", "kernel", nsp, -1);
      auto s = nd_param(*p, "stride", "stride", nsp, 1);
      auto pd = nd_param(*p, "pad", "pad", nsp, 0);
      ECO_CHECK(p->integer("group", 1) == 1, "grouped convolution is not on ECO's path");
      const int nout = (int)p->integer("num_output", 0);
      ECO_CHECK(nout > 0, "num_output missing in " << L.name);
      std::vector<int> o = {b[0], nout};
      for (int i = 0; i < nsp; ++i) {
        ECO_CHECK(k[i] > 0 && s[i] > 0, "Filter/stride dimensions must be nonzero (" << L.name << ")");
        o.push_back((b[2 + i] + 2 * pd[i] - k[i]) / s[i] + 1);  // conv_layer.cpp:21-22
        ECO_CHECK(o.back() > 0, "Convolution " << L.name << " output collapses");
      }
      set_top(0, o);
      std::vector<int> ws = {nout, b[1]};
      ws.insert(ws.end(), k.begin(), k.end());
      set_param_shape(0, ws);
      if (p->boolean("bias_term", true)) set_param_shape(1, {nout});
    } else if (t == "BN") {
      const auto& b = bshape(0);
      set_top(0, b);
      for (size_t i = 0; i < 4; ++i) set_param_shape(i, {1, b[1]});  // bn_layer.cpp:20-41
    } else if (t == "ReLU" || t == "Dropout" || t == "Softmax") {
      set_top(0, bshape(0));
    } else if (t == "Pooling") {
      const pt::Msg* p = L.msg->msg("pooling_param");
      ECO_CHECK(p, "pooling_param missing in " << L.name);
      const auto& b = bshape(0);
      const int nsp = (int)b.size() - 2;
      ECO_CHECK(nsp >= 1 && nsp <= 3, "Pooling " << L.name << ": " << nsp << " spatial axes unsupported");
      std::vector<int> k;
      if (p->boolean("global_pooling", false)) k.assign(b.begin() + 2, b.end());
      else k = nd_param(*p, "kernel_size", "kernel", nsp, -1);
      auto s = nd_param(*p, "stride", "stride", nsp, 1);
      auto pd = nd_param(*p, "pad", "pad", nsp, 0);
      std::vector<int> o = {b[0], b[1]};
      for (int i = 0; i < nsp; ++i) {
        ECO_CHECK(pd[i] < k[i], "pad must be smaller than kernel (" << L.name << ")");  // pooling_layer.cpp:112
        o.push_back(pool_out_dim(b[2 + i], k[i], s[i], pd[i]));
      }
      set_top(0, o);
    } else if (t == "Concat") {
      const pt::Msg* p = L.msg->msg("concat_param");
      const int axis = p ? (int)p->integer("axis", p->integer("concat_dim", 1)) : 1;
      std::vector<int> o = bshape(0);
      ECO_CHECK(axis >= 0 && axis < (int)o.size(), "Concat axis out of range in " << L.name);
      for (size_t j = 1; j < L.bottoms.size(); ++j) {
        const auto& b = bshape((int)j);
        ECO_CHECK(b.size() == o.size(), "All inputs must have the same #axes (" << L.name << ")");
        for (size_t a = 0; a < o.size(); ++a)
          ECO_CHECK((int)a == axis || b[a] == o[a], "All inputs must have the same shape, except at concat_axis ("
                                                        << L.name << ")");
        o[axis] += b[axis];
      }
      set_top(0, o);
    } else if (t == "Eltwise") {
      ECO_CHECK(L.bottoms.size() >= 2, "Eltwise needs 2 bottoms (" << L.name << ")");
      for (size_t j = 1; j < L.bottoms.size(); ++j)
        ECO_CHECK(bshape((int)j) == bshape(0), "Eltwise bottoms must agree in shape (" << L.name << ")");
      set_top(0, bshape(0));
    } else if (t == "Reshape") {
      const pt::Msg* p = L.msg->msg("reshape_param");
      ECO_CHECK(p && p->msg("shape"), "reshape_param.shape missing in " << L.name);
      ECO_CHECK(p->integer("axis", 0) == 0 && p->integer("num_axes", -1) == -1,
                "reshape_param axis/num_axes are not used on ECO's path");
      auto dims = p->msg("shape")->integers("dim");
      const auto& b = bshape(0);
      long long cnt = 1;
      for (int d : b) cnt *= d;
      std::vector<int> o;
      int infer = -1;
      long long known = 1;
      for (size_t i = 0; i < dims.size(); ++i) {
        if (dims[i] == 0) {
          ECO_CHECK(i < b.size(), "reshape dim 0 copies a non-existent axis (" << L.name << ")");
          o.push_back(b[i]);
          known *= b[i];
        } else if (dims[i] == -1) {
          ECO_CHECK(infer < 0, "at most one -1 in reshape (" << L.name << ")");
          infer = (int)i;
          o.push_back(-1);
        } else {
          o.push_back((int)dims[i]);
          known *= dims[i];
        }
      }
      if (infer >= 0) {
        ECO_CHECK(known > 0 && cnt % known == 0, "bottom count must be divisible by the product of the specified dims ("
                                                     << L.name << ")");
        o[infer] = (int)(cnt / known);
        known *= o[infer];
      }
      ECO_CHECK(known == cnt, "output count must match input count (" << L.name << ")");
      set_top(0, o);
    } else if (t == "Permute") {
      const pt::Msg* p = L.msg->msg("permute_param");
      ECO_CHECK(p, "permute_param missing in " << L.name);
      const auto& b = bshape(0);
      std::vector<int> order;
      for (long o : p->integers("order")) {
        ECO_CHECK(o >= 0 && o < (long)b.size(), "order should be less than the input dimension");
        ECO_CHECK(std::find(order.begin(), order.end(), (int)o) == order.end(), "there are duplicate orders");
        order.push_back((int)o);
      }
      for (int i = 0; i < (int)b.size(); ++i)
        if (std::find(order.begin(), order.end(), i) == order.end()) order.push_back(i);  // permute_layer.cpp:44-50
      std::vector<int> o;
      for (int a : order) o.push_back(b[a]);
      set_top(0, o);
    } else if (t == "InnerProduct") {
      const pt::Msg* p = L.msg->msg("inner_product_param");
      ECO_CHECK(p, "inner_product_param missing in " << L.name);
      const auto& b = bshape(0);
      const int nout = (int)p->integer("num_output", 0);
      ECO_CHECK(p->integer("axis", 1) == 1, "InnerProduct axis != 1 is not used on ECO's path");
      long long k = 1;
      for (size_t i = 1; i < b.size(); ++i) k *= b[i];
      set_top(0, {b[0], nout});
      set_param_shape(0, {nout, (int)k});
      if (p->boolean("bias_term", true)) set_param_shape(1, {nout});
    } else if (t == "SoftmaxWithLoss" || t == "Accuracy") {
      set_top(0, {});
    } else if (t == "Split") {
      for (size_t j = 0; j < L.tops.size(); ++j) set_top((int)j, bshape(0));
    } else {
      ECO_CHECK(false, "Unknown layer type: " << t << " (layer '" << L.name
                                              << "'); only the layer types of models_ECO_* are implemented");
    }
  }
}

static void fill(ParamBlob& b, const pt::Msg* filler, std::mt19937& rng, float def_const) {
  long long n = 1;
  for (int d : b.shape) n *= d;
  b.data.assign((size_t)n, def_const);
  b.diff.assign((size_t)n, 0.f);
  if (!filler) return;
  const std::string type = filler->str("type", "constant");
  if (type == "constant") {
    std::fill(b.data.begin(), b.data.end(), (float)filler->num("value", 0.0));
  } else if (type == "xavier") {  // filler.hpp:149-163: U(+-sqrt(3/fan_in))
    const long long fan_in = b.shape.empty() ? 1 : n / b.shape[0];
    const float a = std::sqrt(3.0f / (float)fan_in);
    std::uniform_real_distribution<float> d(-a, a);
    for (auto& v : b.data) v = d(rng);
  } else if (type == "gaussian") {
    std::normal_distribution<float> d((float)filler->num("mean", 0.0), (float)filler->num("std", 1.0));
    for (auto& v : b.data) v = d(rng);
  } else if (type == "uniform") {
    std::uniform_real_distribution<float> d((float)filler->num("min", 0.0), (float)filler->num("max", 1.0));
    for (auto& v : b.data) v = d(rng);
  } else if (type == "msra") {
    const long long fan_in = b.shape.empty() ? 1 : n / b.shape[0];
    std::normal_distribution<float> d(0.f, std::sqrt(2.0f / (float)fan_in));
    for (auto& v : b.data) v = d(rng);
  } else {
    ECO_CHECK(false, "Unknown filler name: " << type);
  }
}

void Net::init_params() {
  std::mt19937 rng(1701);
  for (auto& L : layers_) {
    if (L.type == "Convolution" || L.type == "InnerProduct") {
      const pt::Msg* p = L.msg->msg(L.type == "Convolution" ? "convolution_param" : "inner_product_param");
      fill(L.params[0], p->msg("weight_filler"), rng, 0.f);
      if (L.params.size() > 1) fill(L.params[1], p->msg("bias_filler"), rng, 0.f);
    } else if (L.type == "BN") {
      const pt::Msg* p = L.msg->msg("bn_param");
      const bool frozen = p ? p->boolean("frozen", false) : false;
      fill(L.params[0], p ? p->msg("slope_filler") : nullptr, rng, 1.f);
      fill(L.params[1], p ? p->msg("bias_filler") : nullptr, rng, 0.f);
      fill(L.params[2], nullptr, rng, 0.f);                  // running mean 0       (bn_layer.cpp:34-36)
      fill(L.params[3], nullptr, rng, frozen ? 1.f : 0.f);   // running variance     (bn_layer.cpp:38-41)
    }
    L.params_dirty = true;
  }
}

// =====================================================================================
void Net::set_option(const std::string& key, int v) {
  if (key == "keep_all_blobs") keep_all_ = v != 0;
  else if (key == "a_mode") a_mode_ = v;
  else if (key == "use_graph") use_graph_ = v != 0;
  else if (key == "persistent") persistent_ = v;  // 0 never, 1 auto (per layer), 2 always
  else if (key == "epi_staged") epi_staged_ = v != 0;
  else if (key == "halo") halo_ = v;
  else if (key == "stem_rows") stem_rows_ = v;
  else if (key == "pool_commute") pool_commute_ = v;
  else if (key == "stem_direct") stem_direct_ = v;
  else if (key == "stem_gather_warps") stem_gather_warps_ = std::max(1, std::min(6, v));
  else if (key == "pair") pair_ = v;
  else if (key == "multicast") multicast_ = v;
  else if (key == "fuse_1x1") fuse_1x1_ = v;
  else if (key == "debug_flags") debug_flags_ = v;
  else if (key == "h2d_chunks") h2d_chunks_ = v;  // blocking forward with a host-newer input: 0 auto, 1 never split, n sub-batches
  else if (key == "precision") precision_ = v;  // 0 bf16 storage (default), 1 split precision (fp32-faithful forward, ~3x the MMA work)
  else if (key == "dual_m") dual_m_ = v;  // 0 off, 1 auto, 2 force wherever the accumulators fit
  else ECO_CHECK(false, "unknown option '" << key << "'");
  release_subs();
  free_plan();
}

void Net::set_stream(cudaStream_t s) {
  if (own_stream_ && stream_) cudaStreamDestroy(stream_);
  stream_ = s;  // may be the legacy default stream (NULL): that is a valid choice, not "unset"
  own_stream_ = false;
  user_stream_ = true;
  graph_valid_ = false;
}

// One caffe-style blocking forward(), pipelined inside: when the (single, large) input blob was written on the host, the
// batch is cut into sub-batches of whole videos; sub-net k (same definition, same weights, own stream, own CUDA graph)
// gets its slice by an asynchronous copy straight from the pinned host mirror and runs while slice k+1 is still on the
// PCIe bus; the logits land in this net's output blob.  Videos are independent in TEST phase and the kernels' K order does
// not depend on the batch size, so the result is bit-identical to the unsplit forward (tests/test_gpu_eco.py).
bool Net::try_chunked_forward(int* launches) {
  if (is_sub_ || train_ || keep_all_ || precision_ || h2d_chunks_ == 1 || !until_blob_.empty()) return false;
  if (inputs_.size() != 1 || outputs_.size() != 1) return false;
  Tensor& in = tensors_[vis_blobs_[inputs_[0]].tensor];
  Tensor& out = tensors_[vis_blobs_[outputs_[0]].tensor];
  if (in.kind != Kind::F32 || out.kind != Kind::F32 || in.host.empty() || !in.host.pinned || in.dev_newer) return false;
  if (in.shape.size() < 2 || out.shape.size() != 2) return false;
  const int frames = in.shape[0], videos = out.shape[0];
  if (videos < 2 || frames % videos != 0) return false;
  int chunks = h2d_chunks_;
  if (chunks == 0) {  // auto: only when the copy is worth hiding (>= 64 MB) and sub-batches stay >= 4 videos
    if ((size_t)in.count() * 4 < ((size_t)64 << 20)) return false;
    chunks = 4;
    while (chunks > 1 && (videos % chunks != 0 || videos / chunks < 4)) --chunks;
  }
  if (chunks < 2 || videos % chunks != 0) return false;
  const int sub_videos = videos / chunks, sub_frames = frames / chunks;
  if (sub_.size() != (size_t)chunks || sub_[0]->tensors_[sub_[0]->vis_blobs_[sub_[0]->inputs_[0]].tensor].shape[0] != sub_frames) {
    release_subs();
    for (int c = 0; c < chunks; ++c) {
      std::unique_ptr<Net> n(new Net(text_, phase_));
      n->is_sub_ = true;
      n->a_mode_ = a_mode_; n->use_graph_ = use_graph_; n->persistent_ = persistent_; n->dual_m_ = dual_m_; n->halo_ = halo_;
      n->fuse_1x1_ = fuse_1x1_; n->multicast_ = multicast_; n->pair_ = pair_; n->stem_gather_warps_ = stem_gather_warps_;
      n->stem_direct_ = stem_direct_; n->pool_commute_ = pool_commute_; n->stem_rows_ = stem_rows_; n->epi_staged_ = epi_staged_;
      std::vector<int> dims = in.shape;
      dims[0] = sub_frames;
      n->reshape_blob(n->inputs_[0], dims);
      n->reshape();
      sub_.push_back(std::move(n));
      cudaEvent_t e;
      CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      sub_done_.push_back(e);
    }
  }
  if (sub_params_version_ != params_version_) {
    for (auto& n : sub_)
      for (size_t li = 0; li < layers_.size(); ++li)
        for (size_t k = 0; k < layers_[li].params.size(); ++k) {
          n->layers_[li].params[k].data = layers_[li].params[k].data;
          n->layers_[li].params_dirty = true;
        }
    sub_params_version_ = params_version_;
  }
  const size_t in_chunk = (size_t)in.count() / chunks, out_chunk = (size_t)out.count() / chunks;
  int total = 0;
  // everything already queued on this net's stream (e.g. a previous forward writing the output blob) goes first
  cudaEvent_t& gate = sub_done_[0];
  (void)gate;
  CUDA_OK(cudaStreamSynchronize(stream_));
  for (int c = 0; c < chunks; ++c) {
    Net& n = *sub_[c];
    if (!n.planned_) n.plan();
    Tensor& nin = n.tensors_[n.vis_blobs_[n.inputs_[0]].tensor];
    Tensor& nout = n.tensors_[n.vis_blobs_[n.outputs_[0]].tensor];
    ECO_CHECK(nin.dev && nout.dev && (size_t)nin.count() == in_chunk && (size_t)nout.count() == out_chunk, "chunked forward: sub-net shapes");
    CUDA_OK(cudaMemcpyAsync(nin.dev, in.host.p + (size_t)c * in_chunk, in_chunk * 4, cudaMemcpyHostToDevice, n.stream_));
    nin.dev_newer = true;
    n.forward(0, -1);
    total += n.last_launches_;
    CUDA_OK(cudaMemcpyAsync(static_cast<float*>(out.dev) + (size_t)c * out_chunk, nout.dev, out_chunk * 4, cudaMemcpyDeviceToDevice, n.stream_));
    CUDA_OK(cudaEventRecord(sub_done_[c], n.stream_));
  }
  for (int c = 0; c < chunks; ++c) CUDA_OK(cudaStreamWaitEvent(stream_, sub_done_[c], 0));
  // the host mirror may be rewritten once the last slice has left it
  if (!in.h2d_done) CUDA_OK(cudaEventCreateWithFlags(&in.h2d_done, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(in.h2d_done, sub_[chunks - 1]->stream_));
  in.host_newer = false;
  mark_written(vis_blobs_[outputs_[0]].tensor);
  chunked_last_ = true;
  (void)sub_videos;
  if (launches) *launches = total;
  return true;
}

void Net::reshape_blob(int vb, const std::vector<int>& dims) {
  ECO_CHECK(vb >= 0 && vb < (int)vis_blobs_.size(), "blob index out of range");
  Tensor& t = tensors_[vis_blobs_[vb].tensor];
  for (int d : dims) ECO_CHECK(d >= 0, "negative dimension in reshape");
  t.shape = dims;
  free_plan();
}

void Net::reshape() {
  free_plan();
  infer_shapes();
}

int Net::num_params(int vl) const {
  ECO_CHECK(vl >= 0 && vl < (int)vis_layers_.size(), "layer index out of range");
  const int o = vis_layers_[vl].orig;
  return o < 0 ? 0 : (int)layers_[o].params.size();
}
ParamBlob& Net::param(int vl, int idx) {
  if (params_dev_newer_) sync_params_to_host();
  ECO_CHECK(vl >= 0 && vl < (int)vis_layers_.size(), "layer index out of range");
  const int o = vis_layers_[vl].orig;
  ECO_CHECK(o >= 0 && idx >= 0 && idx < (int)layers_[o].params.size(),
            "layer '" << vis_layers_[vl].name << "' has no parameter blob " << idx);
  return layers_[o].params[idx];
}
void Net::set_param(int vl, int idx, const float* data, size_t count) {
  ParamBlob& b = param(vl, idx);
  ECO_CHECK(count == b.data.size(), "parameter size mismatch for layer '" << vis_layers_[vl].name << "' blob " << idx
                                                                       << ": got " << count << ", expected "
                                                                       << b.data.size());
  std::memcpy(b.data.data(), data, count * sizeof(float));
  layers_[vis_layers_[vl].orig].params_dirty = true;
  ++params_version_;
}
void Net::mark_params_dirty(int vl) {
  const int o = vis_layers_[vl].orig;
  if (o >= 0) layers_[o].params_dirty = true;
  ++params_version_;
}

// =====================================================================================
// device plumbing
void Net::ensure_device() {
  ECO_CHECK(device_available(),
            "no CUDA device is visible: libeco_b200 has no CPU execution path (the reference's CPU mode is not replaced)");
  if (!stream_ && !user_stream_) {
    CUDA_OK(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    own_stream_ = true;
  }
  // function attributes (max dynamic shared memory) and the SM count are per device: caffe allows
  // Caffe::SetDevice to switch devices inside one process (common.hpp:160-174)
  static bool configured[64] = {};
  static int sms[64] = {};
  int dev = 0;
  CUDA_OK(cudaGetDevice(&dev));
  ECO_CHECK(dev >= 0 && dev < 64, "device ordinal " << dev << " out of range");
  if (!configured[dev]) {
    cudaDeviceProp prop;
    CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    ECO_CHECK(prop.major == 10, "libeco_b200 is built for sm_100a only; device is sm_" << prop.major << prop.minor);
    CUDA_OK(conv_umma_configure());
    CUDA_OK(aux_kernels_configure());
    CUDA_OK(wgrad_umma_configure());
    sms[dev] = prop.multiProcessorCount;
    configured[dev] = true;
  }
  g_num_sms = sms[dev];
}

void* Net::dalloc(size_t bytes, bool zero) {
  void* p = nullptr;
  if (bytes == 0) bytes = 16;
  CUDA_OK(cudaMalloc(&p, bytes));
  if (zero) CUDA_OK(cudaMemsetAsync(p, 0, bytes, stream_));
  allocs_.push_back(p);
  return p;
}

void Net::free_plan() {
  if (params_dev_newer_ && P_) {  // keep what training changed on the device (weights, BN running statistics)
    try { sync_params_to_host(); } catch (...) {}
    params_dev_newer_ = false;
  }
  if (graph_exec_) {
    cudaGraphExecDestroy(graph_exec_);
    graph_exec_ = nullptr;
  }
  graph_valid_ = false;
  if (copy_stream_) {
    cudaStreamSynchronize(copy_stream_);
    cudaStreamDestroy(copy_stream_);
    copy_stream_ = nullptr;
    for (int i = 0; i < 2; ++i) {
      if (ev_h2d_[i]) cudaEventDestroy(ev_h2d_[i]);
      if (ev_slot_free_[i]) cudaEventDestroy(ev_slot_free_[i]);
      if (ev_done_[i]) cudaEventDestroy(ev_done_[i]);
      ev_h2d_[i] = ev_slot_free_[i] = ev_done_[i] = nullptr;
      pipe_slot_[i] = nullptr;
    }
    pipe_iter_ = 0;
    pipe_elem_bytes_ = 0;
  }
  if (!allocs_.empty()) {
    if (stream_) cudaStreamSynchronize(stream_);
    for (void* p : allocs_) cudaFree(p);
    allocs_.clear();
  }
  ops_.clear();
  convs_.clear();
  aux_.clear();
  dgrads_.clear();
  P_ = G_ = nullptr;
  arena_count_ = 0;
  slots_.clear();
  slot_index_.clear();
  wgrad_scratch_ = nullptr;
  reduce_scratch_ = nullptr;
  pool_mask_ = nullptr;
  wgrad_scratch_bytes_ = 0;
  params_dev_newer_ = false;
  repack_ = true;
  error_flag_dev_ = nullptr;
  for (auto& t : tensors_) {
    t.dev = nullptr;
    t.dev_bytes = 0;
    t.owns = false;
    t.materialized = false;
    t.root = -1;
    t.cs = 0;
    t.coff = 0;
    t.dev_newer = false;
    t.ddev = nullptr;
    t.needs_grad = false;
    t.diff_dev_newer = false;
  }
  for (auto& L : layers_) L.params_dirty = true;
  planned_ = false;
}

ClView Net::view(const Tensor& t) const {
  ClView v;
  v.ptr = static_cast<__nv_bfloat16*>(t.dev);
  v.outer = t.outer();
  v.inner = t.inner();
  v.C = t.C();
  v.cs = t.cs;
  v.coff = t.coff;
  v.seg = (precision_ && t.kind == Kind::CL) ? t.cs : 0;
  return v;
}

// ---- cuTensorMapEncode* through the runtime's driver entry points (no link-time libcuda) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static void* driver_fn(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q);
  ECO_CHECK(e == cudaSuccess && fn && q == cudaDriverEntryPointSuccess, "driver entry point " << name << " unavailable");
  return fn;
}

void Net::make_tensor_maps(ConvOp& c) {
  static EncodeTiledFn enc_tiled = (EncodeTiledFn)driver_fn("cuTensorMapEncodeTiled");
  static EncodeIm2colFn enc_im2col = (EncodeIm2colFn)driver_fn("cuTensorMapEncodeIm2col");
  // B: weights [Cout_pad][Ktotal] bf16, K-major; box = 64 (one swizzle row) x block_n
  {
    cuuint64_t dims[2] = {(cuuint64_t)c.Ktotal, (cuuint64_t)c.Cout_pad};
    cuuint64_t strides[1] = {(cuuint64_t)c.Ktotal * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)c.kp.block_n};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc_tiled(&c.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c.w_dev, dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    ECO_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights) failed with " << (int)r << " for layer "
                                                                                << layers_[c.conv_layer].name);
  }
  if (c.kp.a_mode != A_TMA_IM2COL) return;
  // A: activations, channels-last.  dims fastest-first {C, W, H, [D], N}; traversal strides = conv stride;
  // base pixel box = [-pad, pad - (k-1)] per axis (the convention of CUTLASS' im2col TMA descriptors).
  const int nsp = c.kp.nsp;
  const int rank = nsp + 2;
  cuuint64_t dims[5];
  cuuint64_t strides[4];
  int lower[3], upper[3];
  cuuint32_t es[5];
  const ConvKernelParams& k = c.kp;
  dims[0] = (cuuint64_t)k.Cin;
  es[0] = 1;
  if (nsp == 3) {
    dims[1] = k.IW; dims[2] = k.IH; dims[3] = k.ID; dims[4] = c.NB;
    strides[0] = k.x_sW * 2; strides[1] = k.x_sH * 2; strides[2] = k.x_sD * 2; strides[3] = k.x_sN * 2;
    lower[0] = -k.pW; lower[1] = -k.pH; lower[2] = -k.pD;
    upper[0] = k.pW - (k.KW - 1); upper[1] = k.pH - (k.KH - 1); upper[2] = k.pD - (k.KD - 1);
    es[1] = k.sW; es[2] = k.sH; es[3] = k.sD; es[4] = 1;
  } else {
    dims[1] = k.IW; dims[2] = k.IH; dims[3] = c.NB;
    strides[0] = k.x_sW * 2; strides[1] = k.x_sH * 2; strides[2] = k.x_sN * 2;
    lower[0] = -k.pW; lower[1] = -k.pH;
    upper[0] = k.pW - (k.KW - 1); upper[1] = k.pH - (k.KH - 1);
    es[1] = k.sW; es[2] = k.sH; es[3] = 1;
  }
  CUresult r = enc_im2col(&c.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, (void*)k.x, dims, strides, lower,
                          upper, (cuuint32_t)kBlockK, (cuuint32_t)kBlockM, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    ECO_CHECK(a_mode_ != A_TMA_IM2COL, "cuTensorMapEncodeIm2col failed with " << (int)r << " for layer "
                                                                              << layers_[c.conv_layer].name);
    c.kp.a_mode = A_GATHER;  // auto mode: fall back to the software gather for this layer
  }
}

// Decide whether a convolution runs on the halo-resident kernel and set it up.
//   * stride-1 2-D filters larger than 1x1 on 64-channel blocks (128-byte pixel rows), and
//   * the 7x7/s2 stem, which is a 4x4/s1 filter over 16-channel space-to-depth cells (32-byte rows).
// The kernel is only worth it when the weights can stay resident in shared memory (measured:
// profiles/r01g_*): with streamed 8-12 KB weight tiles it is no faster than per-tap TMA im2col.
bool Net::plan_halo(ConvOp& c) {
  static EncodeTiledFn enc_tiled = (EncodeTiledFn)driver_fn("cuTensorMapEncodeTiled");
  c.halo = false;
  const ConvKernelParams& k = c.kp;
  if (halo_ == 0 || k.nsp != 2 || k.a_mode != A_TMA_IM2COL) return false;
  int KH, KW, pH, pW, OH, OW, IH, IW, cin, row_bytes, cblocks;
  long long sW, sH, sN;  // element strides of the tensor the patch is cut from
  const void* xbase = k.x;
  if (c.stem) {
    KH = KW = 4; pH = pW = 0; OH = c.O[1]; OW = c.O[2]; IH = c.stem_CH; IW = c.stem_CW;
    cin = 16; row_bytes = 32; cblocks = 1;
    sW = 16; sH = (long long)c.stem_CW * 16; sN = (long long)c.stem_CH * c.stem_CW * 16;
  } else {
    if (c.stem_in) return false;
    if (k.sH != 1 || k.sW != 1 || k.KH * k.KW <= 1 || k.KH > 5 || k.KW > 5) return false;
    KH = k.KH; KW = k.KW; pH = k.pH; pW = k.pW; OH = k.OH; OW = k.OW; IH = k.IH; IW = k.IW;
    cin = k.Cin; row_bytes = 128; cblocks = k.cblocks;
    sW = k.x_sW; sH = k.x_sH; sN = k.x_sN;
  }
  const int pw = OW + KW - 1;
  if (pw > 128) return false;
  int mt = 1;
  if (k.block_n <= 128 && 256 / pw >= 2 &&
      (halo_ == 2 || (long long)c.NB * ((OH + (256 / pw) - 1) / (256 / pw)) >= 2LL * g_num_sms))
    mt = 2;  // two 128-position halves share every weight tile
  const int R = std::min(OH, (128 * mt) / pw);
  if (R < 1) return false;
  if ((double)R * OW / (128.0 * mt) < 0.7) return false;
  HaloKernelParams& h = c.hp;
  h = HaloKernelParams{};
  h.NB = c.NB; h.OH = OH; h.OW = OW; h.KH = KH; h.KW = KW; h.pH = pH; h.pW = pW;
  h.pw = pw; h.R = R; h.bands = (OH + R - 1) / R;
  h.cblocks = cblocks; h.block_n = k.block_n; h.Cout = k.Cout;
  h.row_bytes = row_bytes;
  h.b_kblocks = (int)(c.Ktotal / kBlockK);
  const int patch_rows = (R + KH - 1) * pw;
  const int need_rows = std::max(patch_rows, (KH - 1) * pw + (KW - 1) + 128 * mt);
  h.a_stage_bytes = (uint32_t)round_up(need_rows * row_bytes, 1024);
  h.a_tx_bytes = (uint32_t)patch_rows * (uint32_t)row_bytes;
  const size_t avail = (size_t)227 * 1024 - 1024 - 3 * 1024 - 512;
  const size_t bstage = (size_t)k.block_n * 128;
  const size_t b_all = (size_t)h.b_kblocks * bstage;
  const int n_tiles_n = (k.Cout + k.block_n - 1) / k.block_n;
  h.b_resident = (n_tiles_n == 1 && b_all + 2 * (size_t)h.a_stage_bytes <= avail) ? 1 : 0;
  if (!h.b_resident && halo_ != 3) return false;  // halo:3 forces the streamed-weights variant (tests / A-B)
  if (h.b_resident) {
    h.a_stages = (int)std::min<size_t>(4, (avail - b_all) / h.a_stage_bytes);
    h.b_stages = 1;
  } else {
    h.a_stages = 2;
    if ((size_t)3 * h.a_stage_bytes + 4 * bstage <= avail && h.cblocks > 1) h.a_stages = 3;
    const size_t left = avail - (size_t)h.a_stages * h.a_stage_bytes;
    h.b_stages = (int)std::min<size_t>(8, left / bstage);
    if (h.b_stages < 2) return false;
  }
  h.tmem_cols = pow2_at_least(2 * mt * k.block_n);
  if (h.tmem_cols > 512) return false;
  h.num_sms = g_num_sms;
  h.bias = k.bias; h.scale = k.scale; h.shift = k.shift; h.relu = k.relu;
  h.out = k.out; h.out_cs = k.out_cs; h.out_coff = k.out_coff;
  h.raw = k.raw; h.raw_cs = k.raw_cs; h.raw_coff = k.raw_coff;
  h.res = k.res; h.res_cs = k.res_cs; h.res_coff = k.res_coff;
  h.error_flag = k.error_flag;
  // tiled map over [C, W, H, N]; box = one channel block of the whole patch, zero fill outside the image
  cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)IW, (cuuint64_t)IH, (cuuint64_t)c.NB};
  cuuint64_t strides[3] = {(cuuint64_t)sW * 2, (cuuint64_t)sH * 2, (cuuint64_t)sN * 2};
  cuuint32_t box[4] = {(cuuint32_t)(row_bytes / 2), (cuuint32_t)pw, (cuuint32_t)(R + KH - 1), 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc_tiled(&c.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(xbase), dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  c.halo = true;
  c.halo_mt = mt;
  return true;
}

// Set up the stem rows kernel (conv_umma.cu: stem_rows_kernel) for a convolution plan_conv_group marked `rows`.
void Net::plan_stem_rows(ConvOp& c) {
  static EncodeTiledFn enc_tiled = (EncodeTiledFn)driver_fn("cuTensorMapEncodeTiled");
  ECO_CHECK(c.stem && c.Cout == 64 && c.kp.block_n == 64 && c.O[2] <= 128, "stem rows kernel: unexpected geometry");
  StemRowsParams& r = c.rp;
  r = StemRowsParams{};
  r.F = c.NB; r.OH = c.O[1]; r.OW = c.O[2];
  r.pool = c.pool_tensor >= 0 ? 1 : 0;
  const Tensor& y = tensors_[r.pool ? c.pool_tensor : c.out_tensor];
  ECO_CHECK(y.dev && y.kind == Kind::CL && y.cs % 8 == 0 && y.coff % 8 == 0, "stem rows kernel: output " << y.name);
  r.PH = r.pool ? y.shape[2] : 0;
  r.PW = r.pool ? y.shape[3] : 0;
  if (r.pool) ECO_CHECK(2 * (r.PH - 1) < r.OH && 2 * (r.PW - 1) < r.OW && 2 * r.PH + 1 >= r.OH && 2 * r.PW + 1 >= r.OW, "pool grid");
  // work units: strips of rows per frame, balanced over the SMs (each unit re-reads a 3-row halo)
  const int rows_out = r.pool ? r.PH : r.OH;
  long long best = -1;
  for (int s = 1; s <= 16 && s <= rows_out; ++s) {
    const int strip = (rows_out + s - 1) / s;
    const int strips = (rows_out + strip - 1) / strip;
    const long long cell_rows = (r.pool ? 2 * strip + 1 : strip) + 3;
    const long long cost = (((long long)r.F * strips + g_num_sms - 1) / g_num_sms) * cell_rows;
    if (best < 0 || cost < best) { best = cost; r.strip = strip; r.strips = strips; }
  }
  r.a_stages = r.pool ? 6 : 8;
  r.a_tx_bytes = (uint32_t)r.OW * 128u;
  r.H = c.I[1]; r.W = c.I[2];
  r.gather_warps = stem_gather_warps_;
  if (c.direct_in) {
    // ring of raw image-row pairs (2 rows x 3 channels, sized for fp32) in what is left of the 227 KB
    r.raw_stage_bytes = (uint32_t)round_up(6 * r.W * 4, 128);
    const size_t used = 1024 + (size_t)r.a_stages * 16384 + 4 * 8192 + (r.pool ? 2 * 16384 : 0) + 2048;
    const size_t left = (size_t)227 * 1024 - used;
    r.raw_stages = (int)std::min<size_t>(8, left / r.raw_stage_bytes);
    if (r.raw_stages < 3) { r.a_stages -= 2; r.raw_stages = (int)std::min<size_t>(8, (left + 2 * 16384) / r.raw_stage_bytes); }
    ECO_CHECK(r.raw_stages >= 2, "stem rows kernel: frame rows of " << r.W << " pixels do not fit the raw ring; set stem_direct=0");
  }
  r.num_sms = g_num_sms;
  r.debug_flags = debug_flags_;
  r.bias = c.kp.bias; r.scale = c.kp.scale; r.shift = c.kp.shift; r.relu = c.kp.relu;
  r.out = static_cast<__nv_bfloat16*>(y.dev); r.out_cs = y.cs; r.out_coff = y.coff;
  r.error_flag = c.kp.error_flag;
  if (c.direct_in) {
    std::memset(&c.tmX, 0, sizeof(c.tmX));  // unused: the kernel gathers from the frames
  } else {
  // cell rows as [64-value window, OW windows (32 bytes apart), CH rows, F frames]
  cuuint64_t dims[4] = {64, (cuuint64_t)r.OW, (cuuint64_t)c.stem_CH, (cuuint64_t)c.NB};
  cuuint64_t strides[3] = {32, (cuuint64_t)c.stem_CW * 32, (cuuint64_t)c.stem_CH * c.stem_CW * 32};
  cuuint32_t box[4] = {64, (cuuint32_t)r.OW, 1, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult e = enc_tiled(&c.tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, c.stem_in, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ECO_CHECK(e == CUDA_SUCCESS, "cuTensorMapEncodeTiled(stem cell rows) failed with " << (int)e
                                   << "; set option stem_rows=0 to use the im2col kernel");
  }
  if (r.pool) {
    const double px = (double)c.NB * r.PH * r.PW * 64;
    c.bytes = 2.0 * ((double)c.NB * c.I[1] * c.I[2] * 3 + px + 64.0 * 147);
  }
}

// =====================================================================================
// planner
static bool is_inplace_relu(const OrigLayer& L, int tensor) {
  if (L.type != "ReLU" || L.bottoms.size() != 1 || L.tops.size() != 1) return false;
  if (L.bottoms[0] != tensor || L.tops[0] != tensor) return false;
  const pt::Msg* p = L.msg ? L.msg->msg("relu_param") : nullptr;
  return !p || p->num("negative_slope", 0.0) == 0.0;
}

// 7x7 / stride 2 / pad 3 over a 3-channel fp32 net input (ECO's conv1/7x7_s2)
bool Net::is_stem_conv(const OrigLayer& L) const {
  if (L.type != "Convolution" || L.bottoms.empty()) return false;
  const Tensor& x = tensors_[L.bottoms[0]];
  if (x.kind != Kind::F32 || x.shape.size() != 4 || x.shape[1] != 3) return false;
  const pt::Msg* p = L.msg->msg("convolution_param");
  if (!p) return false;
  auto k = nd_param(*p, "kernel_size", "kernel", 2, -1);
  auto s = nd_param(*p, "stride", "stride", 2, 1);
  auto pd = nd_param(*p, "pad", "pad", 2, 0);
  return k[0] == 7 && k[1] == 7 && s[0] == 2 && s[1] == 2 && pd[0] == 3 && pd[1] == 3;
}

// AVE pooling (3x3, stride 1, pad 1: every window divides by 9, padding counts as zeros) followed by a 1x1
// convolution whose only input it is: both are linear, so conv1x1(pool(x)) == pool(conv1x1_nobias(x)) + bias exactly,
// and the pooling then runs on Cout (32..128) channels instead of Cin (192..608).  Emits CONV (accumulator stored
// as bf16 into an internal tensor) + POOL_CL (bias, BN, ReLU in its epilogue).  Only in the fast plan: the pooled
// blob itself is never formed.  Rounding differs from the reference order by one bf16 rounding of the
// intermediate, like every fused op; the logits tolerance of the whole-net tests covers it.
bool Net::try_commute_pool_conv(int li, std::vector<bool>& done, std::vector<int>& view_of) {
  if (keep_all_ || !pool_commute_) return false;
  const OrigLayer& L = layers_[li];
  const pt::Msg* pp = L.msg->msg("pooling_param");
  if (!pp || pp->boolean("global_pooling", false) || pp->str("pool", "MAX") != "AVE") return false;
  if (L.bottoms.size() != 1 || L.tops.size() != 1) return false;
  const int xb = L.bottoms[0], top = L.tops[0];
  if (tensors_[xb].kind != Kind::CL || tensors_[xb].shape.size() != 4 || tensors_[top].kind != Kind::CL) return false;
  if (tensors_[xb].shape[0] > 65535) return false;
  {
    auto k = nd_param(*pp, "kernel_size", "kernel", 2, -1);
    auto st = nd_param(*pp, "stride", "stride", 2, 1);
    auto pd = nd_param(*pp, "pad", "pad", 2, 0);
    if (k[0] != 3 || k[1] != 3 || st[0] != 1 || st[1] != 1 || pd[0] != 1 || pd[1] != 1) return false;
  }
  if (tensors_[top].consumers.size() != 1) return false;
  const int cj = tensors_[top].consumers[0];
  const OrigLayer& C = layers_[cj];
  if (C.type != "Convolution" || done[cj] || C.bottoms.size() != 1 || C.bottoms[0] != top || C.tops.size() != 1) return false;
  const pt::Msg* cp = C.msg->msg("convolution_param");
  if (!cp || cp->integer("group", 1) != 1) return false;
  {
    auto k = nd_param(*cp, "kernel_size", "kernel", 2, -1);
    auto st = nd_param(*cp, "stride", "stride", 2, 1);
    auto pd = nd_param(*cp, "pad", "pad", 2, 0);
    if (k[0] != 1 || k[1] != 1 || st[0] != 1 || st[1] != 1 || pd[0] != 0 || pd[1] != 0) return false;
  }
  // the pooling then runs on the row-staged kernel (the only one with the bias/BN/ReLU epilogue): its limits
  if (!pool_cl_affine_supported(tensors_[xb].shape[3], C.params[0].shape[0], tensors_[xb].shape[0])) return false;
  // the conv group must reduce to "one stored output" (conv -> BN [-> in-place ReLU])
  const int T0 = C.tops[0];
  if (tensors_[T0].consumers.size() != 1) return false;
  {
    const OrigLayer& B = layers_[tensors_[T0].consumers[0]];
    if (B.type != "BN" || B.tops[0] == T0 || done[tensors_[T0].consumers[0]]) return false;
    const pt::Msg* bp = B.msg->msg("bn_param");
    if (!(phase_ == ECO_PHASE_TEST || (bp && bp->boolean("frozen", false)))) return false;
  }
  const int T = add_tensor(C.name + "@prepool");
  while ((int)view_of.size() <= T) view_of.push_back(-1);
  tensors_[T].shape = tensors_[T0].shape;
  tensors_[T].kind = Kind::CL;
  tensors_[T].ch_axis = 1;
  tensors_[T].producer = cj;
  tensors_[T].consumers.assign(1, li);
  tensors_[T].materialized = true;
  plan_conv_group(cj, done, xb);
  ConvOp& c = convs_.back();
  Op& cop = ops_.back();
  ECO_CHECK(c.out_tensor >= 0 && c.raw_tensor < 0 && c.res_tensor < 0 && !c.rows, "pool_commute: unexpected conv group for " << C.name);
  const int final_t = c.out_tensor;
  c.post_pool = true;
  c.post_relu = c.relu;
  c.relu = false;
  c.raw_tensor = T;
  c.out_tensor = -1;
  cop.first_layer = std::min(cop.first_layer, li);
  Op pop;
  pop.type = Op::POOL_CL;
  pop.name = L.name;
  pop.layer = li;
  pop.first_layer = li;
  pop.last_layer = cop.last_layer;
  pop.in0 = T;
  pop.out = final_t;
  pop.affine_conv = cop.conv;
  ops_.push_back(pop);
  done[li] = true;
  return true;
}

// 1x1 / stride 1 convolutions that read the same channels-last tensor (an inception module's 1x1, 3x3_reduce,
// double_3x3_reduce and -- after pool_commute -- pool_proj) become ONE GEMM over the concatenated output
// channels: the input is read from HBM once instead of once per branch.  Each member keeps its own folded
// BN / ReLU (per-channel constants, per-segment ReLU flag) and its own destination (concat slice or tensor).
// Per-channel results are bit-identical to the unfused plan (same K order per output channel).
void Net::fuse_sibling_1x1() {
  if (!fuse_1x1_ || keep_all_ || persistent_ == 0 || a_mode_ == A_GATHER) return;
  auto member_cout = [&](const Op& op) { return layers_[convs_[op.conv].conv_layer].params[0].shape[0]; };
  auto eligible = [&](const Op& op) -> bool {
    if (op.type != Op::CONV) return false;
    const ConvOp& c = convs_[op.conv];
    if (!c.members.empty() || c.rows || c.res_tensor >= 0 || c.elt_layer >= 0) return false;
    if ((c.out_tensor >= 0) == (c.raw_tensor >= 0)) return false;  // exactly one stored tensor
    const OrigLayer& L = layers_[c.conv_layer];
    const Tensor& x = tensors_[c.in_tensor];
    if (x.kind != Kind::CL || x.shape.size() != 4) return false;
    const pt::Msg* cp = L.msg->msg("convolution_param");
    if (!cp || cp->integer("group", 1) != 1) return false;
    auto k = nd_param(*cp, "kernel_size", "kernel", 2, -1);
    auto st = nd_param(*cp, "stride", "stride", 2, 1);
    auto pd = nd_param(*cp, "pad", "pad", 2, 0);
    if (k[0] != 1 || k[1] != 1 || st[0] != 1 || st[1] != 1 || pd[0] != 0 || pd[1] != 0) return false;
    return L.params[0].shape[0] % 16 == 0;
  };
  for (size_t i = 0; i < ops_.size(); ++i) {
    if (!eligible(ops_[i])) continue;
    std::vector<size_t> grp(1, i);
    int total = member_cout(ops_[i]);
    for (size_t j = i + 1; j < ops_.size() && grp.size() < 4; ++j) {
      if (!eligible(ops_[j]) || convs_[ops_[j].conv].in_tensor != convs_[ops_[i].conv].in_tensor) continue;
      if (total + member_cout(ops_[j]) > 256) continue;
      grp.push_back(j);
      total += member_cout(ops_[j]);
    }
    if (grp.size() < 2) continue;
    ConvOp g = convs_[ops_[grp[0]].conv];  // geometry / input of the first member
    g.out_tensor = g.raw_tensor = -1;
    g.bn_layer = -1;
    g.relu = false;
    g.post_pool = false;
    Op gop = ops_[grp[0]];
    gop.name.clear();
    int off = 0;
    const int gidx = (int)convs_.size();
    for (size_t q : grp) {
      const ConvOp& c = convs_[ops_[q].conv];
      ConvMember m;
      m.conv_layer = c.conv_layer;
      m.bn_layer = c.bn_layer;
      m.relu = c.post_pool ? false : c.relu;
      m.post_pool = c.post_pool;
      m.cout = member_cout(ops_[q]);
      m.off = off;
      m.tensor = c.out_tensor >= 0 ? c.out_tensor : c.raw_tensor;
      g.members.push_back(m);
      g.group_scale |= (c.bn_layer >= 0 && !c.post_pool);
      gop.name += (gop.name.empty() ? "" : "+") + layers_[c.conv_layer].name;
      gop.first_layer = std::min(gop.first_layer, ops_[q].first_layer);
      gop.last_layer = std::max(gop.last_layer, ops_[q].last_layer);
      // a pooling op that applies this member's bias / BN now finds them in the group
      for (auto& o : ops_)
        if (o.type == Op::POOL_CL && o.affine_conv == ops_[q].conv) { o.affine_conv = gidx; o.affine_off = off; }
      off += m.cout;
    }
    gop.conv = gidx;
    convs_.push_back(g);
    ops_[grp[0]] = gop;
    for (size_t q = grp.size(); q-- > 1;) ops_.erase(ops_.begin() + (long)grp[q]);
  }
}

void Net::plan_conv_group(int li, std::vector<bool>& done, int in_override) {
  OrigLayer& L = layers_[li];
  ConvOp c;
  c.conv_layer = li;
  c.in_tensor = in_override >= 0 ? in_override : L.bottoms[0];
  const int T0 = L.tops[0];
  int last = li;
  int pre = T0;
  bool raw_only = false;
  {
    const auto& cons = tensors_[T0].consumers;
    if (cons.size() == 1 && layers_[cons[0]].type == "Eltwise" && !done[cons[0]]) {
      const OrigLayer& E = layers_[cons[0]];
      const pt::Msg* ep = E.msg->msg("eltwise_param");
      bool plain_sum = E.bottoms.size() == 2 && (!ep || (ep->str("operation", "SUM") == "SUM" && !ep->has("coeff")));
      const int other = E.bottoms[0] == T0 ? E.bottoms[1] : E.bottoms[0];
      if (plain_sum && other != T0 && tensors_[other].producer < li && tensors_[other].kind == Kind::CL) {
        c.elt_layer = cons[0];
        c.res_tensor = other;
        pre = E.tops[0];
        done[cons[0]] = true;
        last = std::max(last, cons[0]);
      } else {
        raw_only = true;  // this conv is produced first: keep its raw output for the later add
      }
    }
  }
  int bn = -1;
  if (!raw_only) {
    int nbn = 0;
    for (int ci : tensors_[pre].consumers) {
      const OrigLayer& B = layers_[ci];
      if (B.type == "BN" && B.tops[0] != pre && !done[ci]) {
        const pt::Msg* bp = B.msg->msg("bn_param");
        const bool frozen = bp ? bp->boolean("frozen", false) : false;
        if (phase_ == ECO_PHASE_TEST || frozen) {
          ++nbn;
          bn = ci;
        }
      }
    }
    if (nbn != 1) bn = -1;
    // conv -> in-place layer on `pre` (e.g. ReLU with top == bottom) -> BN: the BN must see the rewritten blob
    // (the reference computes BN(relu(conv))), so it cannot be folded into the conv epilogue
    if (bn >= 0)
      for (int ci : tensors_[pre].consumers) {
        if (ci >= bn || done[ci]) continue;
        const OrigLayer& W = layers_[ci];
        if (std::find(W.tops.begin(), W.tops.end(), pre) != W.tops.end()) { bn = -1; break; }
      }
  }
  if (bn >= 0) {
    c.bn_layer = bn;
    done[bn] = true;
    last = std::max(last, bn);
    const int y = layers_[bn].tops[0];
    c.out_tensor = y;
    const auto& yc = tensors_[y].consumers;
    if (!yc.empty() && is_inplace_relu(layers_[yc[0]], y) && !done[yc[0]]) {
      c.relu = true;
      done[yc[0]] = true;
      last = std::max(last, yc[0]);
    }
    if (tensors_[pre].consumers.size() > 1 || keep_all_) c.raw_tensor = pre;
  } else {
    c.raw_tensor = pre;
    const auto& pc = tensors_[pre].consumers;
    if (!raw_only && !pc.empty() && is_inplace_relu(layers_[pc[0]], pre) && !done[pc[0]]) {
      // conv -> in-place ReLU: store only the rectified value (what caffe leaves in the blob)
      c.out_tensor = pre;
      c.raw_tensor = -1;
      c.relu = true;
      done[pc[0]] = true;
      last = std::max(last, pc[0]);
    }
  }
  done[li] = true;
  // the 7x7/s2 stem: rows kernel when only the rectified output is wanted; the MAX 3x3/s2 pooling that
  // is its sole reader (pool1) is folded into the same kernel unless every blob must be kept
  if (stem_rows_ != 0 && is_stem_conv(L) && L.params[0].shape[0] == 64 && tensors_[T0].shape[3] <= 128 &&
      c.out_tensor >= 0 && c.raw_tensor < 0 && c.res_tensor < 0) {
    c.rows = true;
    if (stem_rows_ == 1 && !keep_all_) {
      int readers = 0, pool = -1;
      for (int ci : tensors_[c.out_tensor].consumers) {
        if (done[ci]) continue;  // the fused BN / in-place ReLU
        ++readers;
        pool = ci;
      }
      if (readers == 1 && layers_[pool].type == "Pooling" && tensors_[layers_[pool].tops[0]].kind == Kind::CL) {
        const OrigLayer& PL = layers_[pool];
        const pt::Msg* pp = PL.msg->msg("pooling_param");
        bool ok = pp && !pp->boolean("global_pooling", false) && pp->str("pool", "MAX") == "MAX" && PL.tops.size() == 1;
        if (ok) {
          auto k = nd_param(*pp, "kernel_size", "kernel", 2, -1);
          auto st = nd_param(*pp, "stride", "stride", 2, 1);
          auto pd = nd_param(*pp, "pad", "pad", 2, 0);
          ok = k[0] == 3 && k[1] == 3 && st[0] == 2 && st[1] == 2 && pd[0] == 0 && pd[1] == 0;
        }
        if (ok) {
          c.pool_layer = pool;
          c.pool_tensor = PL.tops[0];
          c.out_tensor = -1;  // stays on chip
          done[pool] = true;
          last = std::max(last, pool);
          tensors_[c.pool_tensor].materialized = true;
        }
      }
    }
  }
  if (c.out_tensor >= 0) tensors_[c.out_tensor].materialized = true;
  if (c.raw_tensor >= 0) tensors_[c.raw_tensor].materialized = true;

  Op op;
  op.type = Op::CONV;
  op.name = L.name;
  op.first_layer = li;
  op.last_layer = last;
  op.conv = (int)convs_.size();
  convs_.push_back(c);
  ops_.push_back(op);
}

static int pow2_at_least(int v) {
  int p = 32;
  while (p < v) p <<= 1;
  return p;
}

void Net::plan() {
  free_plan();
  ensure_device();
  infer_shapes();
  const int NL = (int)layers_.size();
  // TRAIN-phase nets run the training plan: every blob is materialised (the backward pass reads the activations),
  // the fast-plan rewrites stay off, BN uses batch statistics, Dropout draws a mask, and plan_train() adds the
  // parameter / gradient arenas and the backward state of every op.
  train_ = phase_ == ECO_PHASE_TRAIN;
  struct KeepAllGuard {
    bool& ref; bool saved;
    KeepAllGuard(bool& r, bool force) : ref(r), saved(r) { if (force) ref = true; }
    ~KeepAllGuard() { ref = saved; }
  } keep_all_guard(keep_all_, train_);
  // split precision runs on the one-tile kernel only (its epilogue has the three-plane store); the specialised kernels
  // and fast-plan rewrites are switched off for this plan
  struct IntGuard {
    int& ref; int saved;
    IntGuard(int& r, bool force, int v) : ref(r), saved(r) { if (force) ref = v; }
    ~IntGuard() { ref = saved; }
  };
  ECO_CHECK(!(precision_ && train_), "option precision=1 is an inference mode (TRAIN-phase nets keep fp32 master weights instead)");
  const bool prec = precision_ != 0;
  IntGuard g1(persistent_, prec, 0), g2(pair_, prec, 0), g3(stem_rows_, prec, 0), g4(fuse_1x1_, prec, 0), g5(pool_commute_, prec, 0),
      g6(halo_, prec, 0), g7(multicast_, prec, 0);

  // ---- 1. kinds ----
  for (auto& t : tensors_) {
    t.kind = Kind::F32;
    t.ch_axis = 1;
  }
  std::vector<int> view_of(tensors_.size(), -1);  // top is a pure view of this tensor
  std::vector<bool> needs_convert(NL, false);     // Reshape that cannot stay channels-last
  for (int li = 0; li < NL; ++li) {
    OrigLayer& L = layers_[li];
    const std::string& t = L.type;
    if (is_data_layer(t) || L.tops.empty()) continue;
    Tensor* b0 = L.bottoms.empty() ? nullptr : &tensors_[L.bottoms[0]];
    Tensor& top = tensors_[L.tops[0]];
    if (t == "Convolution") {
      top.kind = Kind::CL;
      top.ch_axis = 1;
      ECO_CHECK(top.shape[1] % 8 == 0, "Convolution " << L.name << ": num_output must be a multiple of 8 for the bf16 "
                                                                  "channels-last path (got " << top.shape[1] << ")");
      if (b0->kind == Kind::CL) ECO_CHECK(b0->ch_axis == 1, "Convolution " << L.name << " input is not N,C,... ordered");
    } else if (t == "BN" || t == "ReLU" || t == "Dropout" || t == "Eltwise") {
      top.kind = b0->kind;
      top.ch_axis = b0->ch_axis;
      if (t == "Dropout" && L.tops[0] != L.bottoms[0] && !train_) view_of[L.tops[0]] = L.bottoms[0];
      if (t == "Eltwise")
        for (int b : L.bottoms) ECO_CHECK(tensors_[b].kind == b0->kind, "Eltwise " << L.name << " mixes layouts");
    } else if (t == "Pooling") {
      if (b0->kind == Kind::CL) {
        ECO_CHECK(b0->ch_axis == 1, "Pooling " << L.name << " input is not N,C,... ordered");
        bool collapses = true;
        for (size_t i = 2; i < top.shape.size(); ++i) collapses &= top.shape[i] == 1;
        top.kind = collapses ? Kind::F32 : Kind::CL;
        top.ch_axis = 1;
      }
    } else if (t == "Concat") {
      const pt::Msg* p = L.msg->msg("concat_param");
      const int axis = p ? (int)p->integer("axis", p->integer("concat_dim", 1)) : 1;
      top.kind = b0->kind;
      top.ch_axis = b0->ch_axis;
      for (int b : L.bottoms) {
        ECO_CHECK(tensors_[b].kind == b0->kind && tensors_[b].ch_axis == b0->ch_axis,
                  "Concat " << L.name << " mixes layouts");
      }
      if (top.kind == Kind::CL) ECO_CHECK(axis == top.ch_axis, "Concat " << L.name << ": only channel concat on feature maps");
      else ECO_CHECK(axis == 1 || axis == 0, "Concat " << L.name << ": axis " << axis << " unsupported on plain blobs");
    } else if (t == "Reshape") {
      if (b0->kind == Kind::CL) {
        // stays channels-last iff the channel axis and everything after it are preserved as the
        // trailing axes (r2Dto3D: [B*N,96,28,28] -> [B,N,96,28,28])
        const int tail = (int)b0->shape.size() - b0->ch_axis;
        bool ok = (int)top.shape.size() >= tail + 1 || (int)top.shape.size() == tail;
        ok = (int)top.shape.size() >= tail;
        for (int i = 0; ok && i < tail; ++i)
          ok = top.shape[top.shape.size() - tail + i] == b0->shape[b0->ch_axis + i];
        if (ok) {
          top.kind = Kind::CL;
          top.ch_axis = (int)top.shape.size() - tail;
          view_of[L.tops[0]] = L.bottoms[0];
        } else {
          needs_convert[li] = true;  // materialise as plain fp32
        }
      } else {
        view_of[L.tops[0]] = L.bottoms[0];
      }
    } else if (t == "Permute") {
      const pt::Msg* p = L.msg->msg("permute_param");
      std::vector<int> order;
      for (long o : p->integers("order")) order.push_back((int)o);
      for (int i = 0; i < (int)b0->shape.size(); ++i)
        if (std::find(order.begin(), order.end(), i) == order.end()) order.push_back(i);
      if (b0->kind == Kind::CL) {
        // physical order of the bottom: axes with the channel axis moved last
        std::vector<int> phys_b;
        for (int i = 0; i < (int)b0->shape.size(); ++i)
          if (i != b0->ch_axis) phys_b.push_back(i);
        phys_b.push_back(b0->ch_axis);
        int ch_top = -1;
        for (int i = 0; i < (int)order.size(); ++i)
          if (order[i] == b0->ch_axis) ch_top = i;
        std::vector<int> phys_t;
        for (int i = 0; i < (int)order.size(); ++i)
          if (i != ch_top) phys_t.push_back(order[i]);
        phys_t.push_back(order[ch_top]);
        ECO_CHECK(phys_t == phys_b, "Permute " << L.name << " is not a layout no-op in channels-last storage; general "
                                                            "permutes are outside ECO's path");
        top.kind = Kind::CL;
        top.ch_axis = ch_top;
        view_of[L.tops[0]] = L.bottoms[0];
      } else {
        bool ident = true;
        for (int i = 0; i < (int)order.size(); ++i) ident &= order[i] == i;
        ECO_CHECK(ident, "Permute " << L.name << " on a plain blob is outside ECO's path");
        view_of[L.tops[0]] = L.bottoms[0];
      }
    } else if (t == "Split") {
      for (int tp : L.tops) {
        tensors_[tp].kind = b0->kind;
        tensors_[tp].ch_axis = b0->ch_axis;
        view_of[tp] = L.bottoms[0];
      }
    }
    // InnerProduct / Softmax / losses: plain fp32 (default)
  }

  // ---- 2. ops (fusion) ----
  std::vector<bool> done(NL, false);
  for (int vb : inputs_) tensors_[vis_blobs_[vb].tensor].materialized = true;
  for (int li = 0; li < NL; ++li) {
    if (done[li]) continue;
    OrigLayer& L = layers_[li];
    const std::string& t = L.type;
    Op op;
    op.first_layer = op.last_layer = li;
    op.layer = li;
    op.name = L.name;
    if (is_data_layer(t)) {
      done[li] = true;
    } else if (t == "Convolution") {
      plan_conv_group(li, done);
    } else if (t == "BN" || t == "ReLU") {
      Tensor& x = tensors_[L.bottoms[0]];
      ECO_CHECK(x.kind == Kind::CL, t << " layer " << L.name << " on a plain blob is outside ECO's path");
      op.type = Op::SSR;
      op.in0 = L.bottoms[0];
      op.out = L.tops[0];
      op.relu = (t == "ReLU");
      if (t == "BN") {
        const pt::Msg* bp = L.msg->msg("bn_param");
        const bool frozen = bp ? bp->boolean("frozen", false) : false;
        if (!(phase_ == ECO_PHASE_TEST || frozen)) op.type = Op::BN_TRAIN;  // batch statistics (bn_layer.cpp:107-157)
        const auto& yc = tensors_[op.out].consumers;
        for (int ci : yc)
          if (ci > li && is_inplace_relu(layers_[ci], op.out) && !done[ci]) {
            // only fuse when the ReLU is the first reader after the BN
            bool first = true;
            for (int cj : yc) first &= (cj >= ci || cj <= li);
            if (first) { op.relu = true; done[ci] = true; op.last_layer = ci; }
            break;
          }
      }
      tensors_[op.out].materialized = true;
      done[li] = true;
      ops_.push_back(op);
    } else if (t == "Pooling" && try_commute_pool_conv(li, done, view_of)) {
      // emitted as conv + pooling-with-epilogue
    } else if (t == "Pooling") {
      Tensor& x = tensors_[L.bottoms[0]];
      Tensor& y = tensors_[L.tops[0]];
      if (x.kind == Kind::CL && y.kind == Kind::F32) op.type = Op::GLOBAL_AVG;
      else if (x.kind == Kind::CL) op.type = Op::POOL_CL;
      else op.type = Op::POOL_F32;
      op.in0 = L.bottoms[0];
      op.out = L.tops[0];
      y.materialized = true;
      done[li] = true;
      ops_.push_back(op);
    } else if (t == "Eltwise") {
      const pt::Msg* ep = L.msg->msg("eltwise_param");
      ECO_CHECK(L.bottoms.size() == 2 && (!ep || (ep->str("operation", "SUM") == "SUM" && !ep->has("coeff"))),
                "Eltwise " << L.name << ": only the 2-input SUM of ECO's residual blocks is implemented");
      ECO_CHECK(tensors_[L.bottoms[0]].kind == Kind::CL, "Eltwise on plain blobs is outside ECO's path");
      op.type = Op::ELTWISE;
      op.in0 = L.bottoms[0];
      op.in1 = L.bottoms[1];
      op.out = L.tops[0];
      tensors_[op.out].materialized = true;
      done[li] = true;
      ops_.push_back(op);
    } else if (t == "Concat") {
      tensors_[L.tops[0]].materialized = true;
      done[li] = true;  // copy ops (if any) are added after aliasing is known
    } else if (t == "Reshape") {
      if (needs_convert[li]) {
        op.type = Op::CL_TO_F32;
        op.in0 = L.bottoms[0];
        op.out = L.tops[0];
        tensors_[op.out].materialized = true;
        ops_.push_back(op);
      }
      done[li] = true;
    } else if (t == "Dropout" && train_) {
      ECO_CHECK(tensors_[L.bottoms[0]].kind == Kind::F32, "Dropout " << L.name << " on a feature map is outside ECO's path "
                                                                              "(ECO drops the pooled vector)");
      op.type = Op::DROPOUT;
      op.in0 = L.bottoms[0];
      op.out = L.tops[0];
      tensors_[op.out].materialized = true;
      done[li] = true;
      ops_.push_back(op);
    } else if (t == "SoftmaxWithLoss" || t == "Accuracy") {
      ECO_CHECK(L.bottoms.size() >= 2 && !L.tops.empty(), t << " " << L.name << " needs scores, labels and a top");
      ECO_CHECK(tensors_[L.bottoms[0]].kind == Kind::F32 && tensors_[L.bottoms[0]].shape.size() == 2,
                t << " " << L.name << ": scores must be a plain [N, classes] blob");
      op.type = t == "Accuracy" ? Op::ACCURACY : Op::LOSS;
      op.in0 = L.bottoms[0];
      op.in1 = L.bottoms[1];
      op.out = L.tops[0];
      tensors_[op.out].materialized = true;
      done[li] = true;
      ops_.push_back(op);
    } else if (t == "Permute" || t == "Dropout" || t == "Split") {
      done[li] = true;
    } else if (t == "InnerProduct") {
      ECO_CHECK(tensors_[L.bottoms[0]].kind == Kind::F32,
                "InnerProduct " << L.name << " directly on a feature map is outside ECO's path (pool first)");
      op.type = Op::FC;
      op.in0 = L.bottoms[0];
      op.out = L.tops[0];
      tensors_[op.out].materialized = true;
      done[li] = true;
      ops_.push_back(op);
    } else if (t == "Softmax") {
      op.type = Op::SOFTMAX;
      op.in0 = L.bottoms[0];
      op.out = L.tops[0];
      tensors_[op.out].materialized = true;
      done[li] = true;
      ops_.push_back(op);
    } else {
      ECO_CHECK(false, "layer type " << t << " (" << L.name << ") has no device implementation in this round");
    }
  }

  fuse_sibling_1x1();

  // ---- 3. storage: views, zero-copy concat, allocation ----
  for (size_t i = 0; i < tensors_.size(); ++i) tensors_[i].root = (int)i;
  // concat aliasing, later concats first so nested concats resolve outward-in
  std::vector<std::pair<int, int>> concat_copies;  // (layer, bottom idx) that need a copy
  std::vector<char> aliased(tensors_.size(), 0);
  for (int li = NL - 1; li >= 0; --li) {
    OrigLayer& L = layers_[li];
    if (L.type != "Concat") continue;
    Tensor& top = tensors_[L.tops[0]];
    if (top.kind == Kind::CL && top.cs == 0) {
      top.cs = round_up(top.C(), 8);
      top.coff = 0;
    }
    int off = 0;
    for (size_t j = 0; j < L.bottoms.size(); ++j) {
      Tensor& b = tensors_[L.bottoms[j]];
      const int width = top.kind == Kind::CL ? b.C() : 0;
      // readers other than layers that rewrite the blob in place (the fused ReLU after a BN)
      int readers = 0;
      for (int ci : b.consumers) {
        const OrigLayer& Lc = layers_[ci];
        const bool inplace = Lc.tops.size() == 1 && Lc.bottoms.size() == 1 && Lc.tops[0] == L.bottoms[j];
        if (!inplace) ++readers;
      }
      bool can_alias = top.kind == Kind::CL && readers == 1 && b.materialized && view_of[L.bottoms[j]] < 0 &&
                       !aliased[L.bottoms[j]] && b.producer >= 0 && !is_data_layer(layers_[b.producer].type) &&
                       layers_[b.producer].type != "Concat" && (off % 8 == 0);
      // a tensor that other tensors view must keep its own buffer
      for (size_t q = 0; can_alias && q < view_of.size(); ++q)
        if (view_of[q] == L.bottoms[j]) can_alias = false;
      if (can_alias) {
        b.root = top.root;
        b.cs = top.cs;
        b.coff = top.coff + off;
        aliased[L.bottoms[j]] = 1;
      } else {
        concat_copies.emplace_back(li, (int)j);
      }
      off += width;
    }
  }
  // views inherit storage (resolve chains in layer order: a view's source is always earlier)
  for (int li = 0; li < NL; ++li)
    for (int tp : layers_[li].tops)
      if (view_of[tp] >= 0) {
        Tensor& s = tensors_[view_of[tp]];
        Tensor& v = tensors_[tp];
        v.root = s.root;
        v.cs = s.cs;  // may still be 0 here; fixed after allocation below
        v.coff = s.coff;
        v.materialized = s.materialized;
      }
  error_flag_dev_ = static_cast<int*>(dalloc(sizeof(int), true));
  for (size_t i = 0; i < tensors_.size(); ++i) {
    Tensor& t = tensors_[i];
    if (t.root != (int)i || !t.materialized || view_of[i] >= 0) continue;
    if (t.kind == Kind::CL) {
      if (t.cs == 0) t.cs = round_up(t.C(), 8);
      ECO_CHECK(t.C() % 8 == 0 || t.cs >= t.C(), "bad channel stride");
      t.dev_bytes = (size_t)(t.outer() * t.inner()) * (size_t)t.cs * 2 * (precision_ ? 3 : 1);
      t.dev = dalloc(t.dev_bytes, t.cs != t.C());
    } else {
      t.dev_bytes = (size_t)std::max<long long>(t.count(), 1) * 4;
      t.dev = dalloc(t.dev_bytes, true);
    }
    t.owns = true;
  }
  for (int pass = 0; pass < 2; ++pass)
    for (int li = 0; li < NL; ++li)
      for (int tp : layers_[li].tops) {
        Tensor& v = tensors_[tp];
        if (view_of[tp] >= 0) {
          Tensor& s = tensors_[view_of[tp]];
          v.dev = s.dev; v.cs = s.cs; v.coff = s.coff; v.root = s.root; v.dev_bytes = s.dev_bytes;
          v.materialized = s.materialized;
        } else if (v.root != tp) {
          Tensor& r = tensors_[v.root];
          v.dev = r.dev; v.dev_bytes = r.dev_bytes;
        }
      }
  // copies for concat bottoms that could not be aliased; inserted right after the producer of the
  // last bottom, i.e. appended in layer order
  if (!concat_copies.empty()) {
    std::vector<Op> extra;
    for (auto& cc : concat_copies) {
      OrigLayer& L = layers_[cc.first];
      Tensor& top = tensors_[L.tops[0]];
      Tensor& b = tensors_[L.bottoms[cc.second]];
      ECO_CHECK(b.materialized && b.dev, "Concat " << L.name << ": bottom " << b.name << " has no storage");
      Op op;
      op.type = Op::COPY2D;
      op.name = L.name + ":copy" + std::to_string(cc.second);
      op.first_layer = op.last_layer = cc.first;
      op.layer = cc.first;
      op.in0 = L.bottoms[cc.second];
      op.out = L.tops[0];
      if (top.kind == Kind::CL) {
        int off = 0;
        for (int j = 0; j < cc.second; ++j) off += tensors_[L.bottoms[j]].C();
        op.rows = (size_t)(b.outer() * b.inner());
        op.width_bytes = (size_t)b.C() * 2;
        op.src_pitch = (size_t)b.cs * 2;
        op.dst_pitch = (size_t)top.cs * 2;
        op.src_off = (size_t)b.coff * 2;
        op.dst_off = (size_t)(top.coff + off) * 2;
      } else {
        const pt::Msg* p = L.msg->msg("concat_param");
        const int axis = p ? (int)p->integer("axis", p->integer("concat_dim", 1)) : 1;
        long long outer = 1, in_w = 1, top_w = 1, off = 0;
        for (int a = 0; a < axis; ++a) outer *= top.shape[a];
        for (size_t a = axis; a < b.shape.size(); ++a) in_w *= b.shape[a];
        for (size_t a = axis; a < top.shape.size(); ++a) top_w *= top.shape[a];
        for (int j = 0; j < cc.second; ++j) {
          long long w = 1;
          for (size_t a = axis; a < tensors_[L.bottoms[j]].shape.size(); ++a) w *= tensors_[L.bottoms[j]].shape[a];
          off += w;
        }
        op.rows = (size_t)outer;
        op.width_bytes = (size_t)in_w * 4;
        op.src_pitch = (size_t)in_w * 4;
        op.dst_pitch = (size_t)top_w * 4;
        op.src_off = 0;
        op.dst_off = (size_t)off * 4;
      }
      op.launches = 0;
      extra.push_back(op);
    }
    // merge into ops_ keeping layer order (copy goes where the Concat layer sits)
    std::vector<Op> merged;
    size_t e = 0;
    std::stable_sort(extra.begin(), extra.end(), [](const Op& a, const Op& b) { return a.first_layer < b.first_layer; });
    for (auto& op : ops_) {
      while (e < extra.size() && extra[e].first_layer < op.first_layer) merged.push_back(extra[e++]);
      merged.push_back(op);
    }
    while (e < extra.size()) merged.push_back(extra[e++]);
    ops_.swap(merged);
  }

  // ---- 4. bind ops ----
  for (auto& op : ops_) {
    OrigLayer* Lp = op.layer >= 0 ? &layers_[op.layer] : nullptr;
    switch (op.type) {
      case Op::CONV: {
        ConvOp& c = convs_[op.conv];
        OrigLayer& L = layers_[c.conv_layer];
        const pt::Msg* p = L.msg->msg("convolution_param");
        Tensor& x = tensors_[c.in_tensor];
        const int nsp = (int)x.shape.size() - 2;
        ECO_CHECK(nsp == 2 || nsp == 3, "Convolution " << L.name << ": only 2-D and 3-D convolutions are implemented");
        auto k = nd_param(*p, "kernel_size", "kernel", nsp, -1);
        auto s = nd_param(*p, "stride", "stride", nsp, 1);
        auto pd = nd_param(*p, "pad", "pad", nsp, 0);
        ECO_CHECK(p->integer("dilation", 1) == 1, "dilated convolution is not on ECO's path");
        c.nsp = nsp;
        c.NB = x.shape[0];
        c.Cin = x.shape[1];
        c.Cout = L.params[0].shape[0];
        if (!c.members.empty()) {
          c.Cout = 0;
          for (const ConvMember& m : c.members) c.Cout += m.cout;
        }
        for (int i = 0; i < nsp; ++i) {
          const int a = 3 - nsp + i;
          c.K[a] = k[i]; c.S[a] = s[i]; c.P[a] = pd[i];
          c.I[a] = x.shape[2 + i];
          c.O[a] = (x.shape[2 + i] + 2 * pd[i] - k[i]) / s[i] + 1;
        }
        ConvKernelParams& kp = c.kp;
        kp = ConvKernelParams{};
        kp.nsp = nsp;
        kp.OD = c.O[0]; kp.OH = c.O[1]; kp.OW = c.O[2];
        kp.M = c.NB * c.O[0] * c.O[1] * c.O[2];
        c.stem = (x.kind == Kind::F32 && nsp == 2 && c.Cin == 3 && k[0] == 7 && k[1] == 7 && s[0] == 2 && s[1] == 2 &&
                  pd[0] == 3 && pd[1] == 3) && !precision_;
        if (c.stem) {
          // 7x7/s2/p3 over 3 channels == 4x4/s1 over 2x2 space-to-depth cells of the zero-padded image;
          // the 4 horizontally adjacent cells (4 x 16 ch) of a window are contiguous in memory, so the
          // conv is presented to the GEMM as a 4(h) x 1(w) kernel over 64 "channels" with pixel stride 16.
          c.stem_CH = c.O[1] + 3;
          c.stem_CW = c.O[2] + 3;
          c.direct_in = c.rows && stem_direct_ && (c.I[2] % 16 == 0);
          c.stem_bytes = c.direct_in ? 256 : (size_t)c.NB * c.stem_CH * c.stem_CW * 16 * 2 + 256;
          c.stem_in = static_cast<__nv_bfloat16*>(dalloc(c.stem_bytes, true));
          kp.x = c.stem_in;
          kp.x_sW = 16; kp.x_sH = (long long)c.stem_CW * 16; kp.x_sD = 0;
          kp.x_sN = (long long)c.stem_CH * c.stem_CW * 16;
          kp.Cin = 64; c.Cin_k = 64;
          kp.ID = 1; kp.IH = c.stem_CH; kp.IW = c.O[2];
          kp.KD = 1; kp.KH = 4; kp.KW = 1;
          kp.sD = kp.sH = kp.sW = 1;
          kp.pD = kp.pH = kp.pW = 0;
        } else {
          if (x.kind == Kind::F32) {
            // generic fp32 input: convert to channels-last bf16 with channels padded to 8
            const int c8 = round_up(c.Cin, 8);
            const int planes = precision_ ? 3 : 1;
            c.stem_bytes = (size_t)c.NB * c.I[0] * c.I[1] * c.I[2] * c8 * 2 * planes;
            c.stem_in = static_cast<__nv_bfloat16*>(dalloc(c.stem_bytes, true));
            kp.x = c.stem_in;
            kp.x_sW = c8 * planes;
            c.Cin_k = c8 * planes;
            c.cin_stride = c8;
          } else {
            ECO_CHECK(x.dev, "Convolution " << L.name << ": input " << x.name << " is not materialised");
            ECO_CHECK(x.C() % 8 == 0 && x.coff % 8 == 0 && x.cs % 8 == 0, "channel alignment");
            kp.x = static_cast<__nv_bfloat16*>(x.dev) + x.coff;
            kp.x_sW = x.cs;
            c.Cin_k = c.Cin;
            c.cin_stride = (int)x.cs;
            if (precision_) {
              // the three planes [hi | lo | hi] of a whole buffer are the 3C "channels" of the GEMM's K axis
              ECO_CHECK(x.coff == 0 && x.cs == x.C(), "precision=1: convolution " << L.name << " reads a channel slice of a larger buffer");
              kp.x_sW = 3 * x.cs;
              c.Cin_k = 3 * c.Cin;
            }
          }
          kp.x_sH = kp.x_sW * c.I[2];
          kp.x_sD = kp.x_sH * c.I[1];
          kp.x_sN = kp.x_sD * c.I[0];
          kp.Cin = c.Cin_k;
          kp.ID = c.I[0]; kp.IH = c.I[1]; kp.IW = c.I[2];
          kp.KD = c.K[0]; kp.KH = c.K[1]; kp.KW = c.K[2];
          kp.sD = c.S[0]; kp.sH = c.S[1]; kp.sW = c.S[2];
          kp.pD = c.P[0]; kp.pH = c.P[1]; kp.pW = c.P[2];
        }
        kp.cblocks = (kp.Cin + kBlockK - 1) / kBlockK;
        kp.num_kb = kp.KD * kp.KH * kp.KW * kp.cblocks;
        c.Ktotal = (long long)kp.num_kb * kBlockK;
        kp.Cout = c.Cout;
        const int ntiles = (c.Cout + 255) / 256;
        kp.block_n = round_up((c.Cout + ntiles - 1) / ntiles, 16);
        if (c.members.empty()) {
          // wave quantisation on the persistent grid: more, narrower N tiles when that lowers
          // ceil(tiles / SMs) x tile cost.  Tile cost ~ shared-memory bytes moved per K block (the bound of
          // cta_group::1 kernels): MMA operand reads 4 x (4096 + 32 N) + TMA writes 16384 + 128 N.
          // e.g. res5 (M = 6272, Cout = 512) at batch 32: 98 tiles of 256 on 148 SMs -> 147 tiles of 176.
          auto cost = [&](int bn) {
            const long long tiles = ((long long)kp.M + kBlockM - 1) / kBlockM * ((c.Cout + bn - 1) / bn);
            return ((tiles + g_num_sms - 1) / g_num_sms) * (32768LL + 256LL * bn);
          };
          int best = kp.block_n;
          for (int nt = ntiles + 1; nt <= ntiles + 3; ++nt) {
            const int bn = round_up((c.Cout + nt - 1) / nt, 16);
            if (bn < 96) break;
            if (cost(bn) < cost(best)) best = bn;
          }
          kp.block_n = best;
        }
        // CTA-pair kernel (cta_group::2, 256-row tiles on two SMs, each CTA loads half of the weight tile).
        // Measured (profiles/r01l): a cta_group::2 MMA has an issue interval of ~150-165 cycles whatever N, so the
        // pair only pays for 256-wide tiles (res4: -6..15 %) and loses for N <= 192; auto mode uses it there only.
        c.pair = false;
        if (pair_ && persistent_ && !c.stem && (a_mode_ < 0 || a_mode_ == A_TMA_IM2COL)) {
          const long long t256 = ((long long)kp.M + 2 * kBlockM - 1) / (2 * kBlockM) * ((c.Cout + kp.block_n - 1) / kp.block_n);
          // (short-K 256-wide GEMMs -- the fused 1x1 groups -- are epilogue-bound and run better on single CTAs)
          if (pair_ == 2 || (kp.block_n == 256 && kp.num_kb >= 16 && t256 >= 2LL * (g_num_sms / 2))) c.pair = true;
        }
        c.Cout_pad = round_up(c.Cout, kp.block_n);
        kp.a_mode = a_mode_ < 0 ? A_TMA_IM2COL : a_mode_;
        kp.num_sms = g_num_sms;
        kp.error_flag = error_flag_dev_;
        kp.relu = c.relu ? 1 : 0;
        c.w_dev = static_cast<__nv_bfloat16*>(dalloc((size_t)c.Cout_pad * c.Ktotal * 2, true));
        c.bias_dev = static_cast<float*>(dalloc((size_t)c.Cout * 4, true));
        if (c.bn_layer >= 0 || c.group_scale) {
          c.scale_dev = static_cast<float*>(dalloc((size_t)c.Cout * 4, true));
          c.shift_dev = static_cast<float*>(dalloc((size_t)c.Cout * 4, true));
        }
        for (const ConvMember& m : c.members)
          if (m.post_pool && !c.post_bias_dev) {
            c.post_bias_dev = static_cast<float*>(dalloc((size_t)c.Cout * 4, true));
            c.post_scale_dev = static_cast<float*>(dalloc((size_t)c.Cout * 4, true));
            c.post_shift_dev = static_cast<float*>(dalloc((size_t)c.Cout * 4, true));
          }
        kp.bias = c.bias_dev;
        kp.scale = c.scale_dev;
        kp.shift = c.shift_dev;
        if (c.post_pool) {  // the pooling op behind this conv applies them
          kp.bias = nullptr; kp.scale = nullptr; kp.shift = nullptr; kp.relu = 0;
        }
        auto bind = [&](int tid, __nv_bfloat16*& ptr, long long& cs, int& coff) {
          if (tid < 0) { ptr = nullptr; cs = 0; coff = 0; return; }
          Tensor& t = tensors_[tid];
          ECO_CHECK(t.dev && t.kind == Kind::CL, "conv " << L.name << ": tensor " << t.name << " has no channels-last storage");
          ECO_CHECK(t.cs % 8 == 0 && t.coff % 8 == 0, "channel alignment of " << t.name);
          ptr = static_cast<__nv_bfloat16*>(t.dev);
          cs = precision_ ? 3 * t.cs : t.cs;   // pixel stride
          coff = t.coff;
        };
        auto seg_of = [&](int tid) -> long long { return (precision_ && tid >= 0) ? tensors_[tid].cs : 0; };
        if (!c.members.empty()) {
          ECO_CHECK(c.members.size() <= 4, "fused 1x1 group too large");
          kp.nseg = (int)c.members.size();
          for (size_t q = 0; q < c.members.size(); ++q) {
            const ConvMember& m = c.members[q];
            long long cs = 0;
            int coff = 0;
            bind(m.tensor, kp.seg_ptr[q], cs, coff);
            kp.seg_cs[q] = cs;
            kp.seg_coff[q] = coff - m.off;
            kp.seg_end[q] = m.off + m.cout;
            kp.seg_relu[q] = m.relu ? 1 : 0;
          }
          kp.out = kp.seg_ptr[0];
          kp.out_cs = kp.seg_cs[0];
          kp.out_coff = kp.seg_coff[0];
        } else
        bind(c.out_tensor, kp.out, kp.out_cs, kp.out_coff);
        bind(c.raw_tensor, kp.raw, kp.raw_cs, kp.raw_coff);
        __nv_bfloat16* rp = nullptr;
        bind(c.res_tensor, rp, kp.res_cs, kp.res_coff);
        kp.res = rp;
        kp.out_seg = seg_of(c.out_tensor);
        kp.raw_seg = seg_of(c.raw_tensor);
        kp.res_seg = seg_of(c.res_tensor);
        make_tensor_maps(c);  // may downgrade a_mode to the gather for this layer
        {
          const size_t per_stage = (size_t)kBlockM * 128 + (size_t)kp.block_n * 128;
          kp.persistent = (persistent_ && kp.a_mode == A_TMA_IM2COL) ? 1 : 0;
          // (until the MMA issue path was fixed -- elect.sync instead of lane 0, profiles/r01j -- long-K layers ran
          // faster as one-tile CTAs, two per SM; since then the persistent kernel wins on every layer, ab21)
          kp.m_halves = 1;
          if (kp.persistent) {
            // one CTA per SM: deep ring, double-buffered accumulator (2 x m_halves x block_n TMEM columns).
            // Two 128-row halves share each weight tile when the accumulators fit (block_n <= 128) and
            // there is enough work to keep every SM busy with the larger tiles.
            const long long tiles256 = ((long long)kp.M + 255) / 256 * ((c.Cout + kp.block_n - 1) / kp.block_n);
            if (kp.block_n <= 128 && (dual_m_ == 2 || (dual_m_ == 1 && tiles256 >= 2LL * g_num_sms))) kp.m_halves = 2;
            const size_t stage_bytes = (size_t)kBlockM * 128 * kp.m_halves + (size_t)kp.block_n * 128;
            // shared memory: 227 KB - alignment slack - constants - barriers; the epilogue staging takes
            // 4 chunks per copy-out (128-byte row pieces) unless that would cost a pipeline stage
            const size_t avail = (size_t)227 * 1024 - 1024 - 3 * 1024 - 512;
            kp.epi_group = 4;   // chunks per epilogue group
            kp.epi_staged = epi_staged_ ? 1 : 0;  // default 0: direct 32-byte row-piece stores (measured faster, profiles/r01i)
            kp.debug_flags = debug_flags_;
            kp.stages = (int)std::max<size_t>(2, std::min<size_t>(8, (avail - (kp.epi_staged ? conv_epi_stage_bytes(kp.epi_group) : 0)) / stage_bytes));
            kp.tmem_cols = pow2_at_least(2 * kp.m_halves * kp.block_n);
          } else {
            const size_t budget = kp.block_n <= 128 ? 100 * 1024 : 200 * 1024;
            kp.stages = (int)std::max<size_t>(2, std::min<size_t>(6, budget / per_stage));
            kp.tmem_cols = pow2_at_least(kp.block_n);
          }
        }
        ECO_CHECK(c.members.empty() || kp.persistent, "fused 1x1 group " << op.name << " needs the persistent im2col kernel");
        plan_halo(c);
        const double taps = (double)c.K[0] * c.K[1] * c.K[2];
        c.flops = 2.0 * kp.M * c.Cout * taps * c.Cin;
        c.bytes = 2.0 * ((double)c.NB * c.I[0] * c.I[1] * c.I[2] * c.Cin + (double)kp.M * c.Cout + (double)c.Cout * taps * c.Cin);
        if (c.rows) plan_stem_rows(c);
        if (c.pair && (kp.a_mode != A_TMA_IM2COL || !kp.persistent)) c.pair = false;
        kp.multicast = 0;
        if (multicast_ && kp.persistent && !c.pair && !c.rows && !c.halo) {
          const long long tile_m = (long long)kBlockM * kp.m_halves;
          const long long tiles = (kp.M + tile_m - 1) / tile_m * ((c.Cout + kp.block_n - 1) / kp.block_n);
          if (multicast_ == 2 || tiles >= 2LL * g_num_sms) kp.multicast = 1;
        }
        if (c.pair || kp.multicast) {
          static EncodeTiledFn enc_tiled = (EncodeTiledFn)driver_fn("cuTensorMapEncodeTiled");
          c.kpp = kp;
          c.kpp.multicast = 0;
          c.kpp.m_halves = 1;
          const size_t stage = (size_t)kBlockM * 128 + (size_t)(kp.block_n / 2) * 128;
          const size_t avail = (size_t)227 * 1024 - 1024 - 3 * 1024 - 512;
          c.kpp.stages = (int)std::max<size_t>(2, std::min<size_t>(8, avail / stage));
          c.kpp.tmem_cols = pow2_at_least(2 * kp.block_n);
          cuuint64_t dims[2] = {(cuuint64_t)c.Ktotal, (cuuint64_t)c.Cout_pad};
          cuuint64_t strides[1] = {(cuuint64_t)c.Ktotal * 2};
          cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)(kp.block_n / 2)};
          cuuint32_t es[2] = {1, 1};
          CUresult r = enc_tiled(&c.tmBh, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, c.w_dev, dims, strides, box, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
          if (r != CUDA_SUCCESS) { c.pair = false; kp.multicast = 0; }
        }
        op.flops = c.flops;
        op.bytes = c.bytes;
        op.launches = 1 + ((c.stem_in && !c.direct_in) ? 1 : 0);
        break;
      }
      case Op::POOL_CL: {
        const pt::Msg* p = Lp->msg->msg("pooling_param");
        Tensor& x = tensors_[op.in0];
        Tensor& y = tensors_[op.out];
        const int nsp = (int)x.shape.size() - 2;
        std::vector<int> k;
        if (p->boolean("global_pooling", false)) k.assign(x.shape.begin() + 2, x.shape.end());
        else k = nd_param(*p, "kernel_size", "kernel", nsp, -1);
        auto s = nd_param(*p, "stride", "stride", nsp, 1);
        auto pd = nd_param(*p, "pad", "pad", nsp, 0);
        const std::string method = p->str("pool", "MAX");
        ECO_CHECK(method == "MAX" || method == "AVE", "pooling method " << method << " is not on ECO's path");
        ECO_CHECK(x.C() % 8 == 0, "Pooling " << Lp->name << ": channels must be a multiple of 8");
        PoolParams& q = op.pool;
        q = PoolParams{};
        q.x = static_cast<__nv_bfloat16*>(x.dev); q.x_cs = x.cs; q.x_coff = x.coff;
        q.y = static_cast<__nv_bfloat16*>(y.dev); q.y_cs = y.cs; q.y_coff = y.coff;
        q.NB = x.shape[0]; q.C = x.C();
        int I[3] = {1, 1, 1}, O[3] = {1, 1, 1}, K[3] = {1, 1, 1}, S[3] = {1, 1, 1}, P[3] = {0, 0, 0};
        for (int i = 0; i < nsp; ++i) {
          const int a = 3 - nsp + i;
          I[a] = x.shape[2 + i]; O[a] = y.shape[2 + i]; K[a] = k[i]; S[a] = s[i]; P[a] = pd[i];
        }
        q.ID = I[0]; q.IH = I[1]; q.IW = I[2]; q.OD = O[0]; q.OH = O[1]; q.OW = O[2];
        q.KD = K[0]; q.KH = K[1]; q.KW = K[2]; q.sD = S[0]; q.sH = S[1]; q.sW = S[2];
        q.pD = P[0]; q.pH = P[1]; q.pW = P[2];
        q.is_max = method == "MAX";
        if (op.affine_conv >= 0) {
          const ConvOp& ac = convs_[op.affine_conv];
          q.bias = ac.bias_dev; q.scale = ac.scale_dev; q.shift = ac.shift_dev; q.relu = ac.post_relu ? 1 : 0;
          if (op.affine_off >= 0) {  // member of a fused 1x1 group: its constants live in the group's post arrays
            q.bias = ac.post_bias_dev + op.affine_off;
            q.scale = q.shift = nullptr;
            q.relu = 0;
            for (const ConvMember& m : ac.members)
              if (m.off == op.affine_off) {
                if (m.bn_layer >= 0) { q.scale = ac.post_scale_dev + m.off; q.shift = ac.post_shift_dev + m.off; }
                // the member's own ConvOp recorded whether a ReLU follows
                for (const ConvOp& oc : convs_)
                  if (oc.members.empty() && oc.conv_layer == m.conv_layer) q.relu = oc.post_relu ? 1 : 0;
              }
          }
        }
        op.bytes = 2.0 * ((double)x.count() + (double)y.count());
        break;
      }
      case Op::GLOBAL_AVG: {
        const pt::Msg* p = Lp->msg->msg("pooling_param");
        Tensor& x = tensors_[op.in0];
        const int nsp = (int)x.shape.size() - 2;
        std::vector<int> k;
        if (p->boolean("global_pooling", false)) k.assign(x.shape.begin() + 2, x.shape.end());
        else k = nd_param(*p, "kernel_size", "kernel", nsp, -1);
        auto pd = nd_param(*p, "pad", "pad", nsp, 0);
        bool full = p->str("pool", "MAX") == "AVE";
        for (int i = 0; i < nsp; ++i) full &= (k[i] == x.shape[2 + i] && pd[i] == 0);
        ECO_CHECK(full, "Pooling " << Lp->name << " collapses the map but is not a full-extent AVE pool");
        op.bytes = 2.0 * (double)x.count();
        break;
      }
      case Op::POOL_F32: {
        const pt::Msg* p = Lp->msg->msg("pooling_param");
        Tensor& x = tensors_[op.in0];
        Tensor& y = tensors_[op.out];
        const int nsp = (int)x.shape.size() - 2;
        std::vector<int> k;
        if (p->boolean("global_pooling", false)) k.assign(x.shape.begin() + 2, x.shape.end());
        else k = nd_param(*p, "kernel_size", "kernel", nsp, -1);
        auto s = nd_param(*p, "stride", "stride", nsp, 1);
        auto pd = nd_param(*p, "pad", "pad", nsp, 0);
        const std::string method = p->str("pool", "MAX");
        PoolF32Params& q = op.poolf;
        q = PoolF32Params{};
        q.x = static_cast<float*>(x.dev);
        q.y = static_cast<float*>(y.dev);
        q.NC = x.shape[0] * x.shape[1];
        int I[3] = {1, 1, 1}, O[3] = {1, 1, 1}, K[3] = {1, 1, 1}, S[3] = {1, 1, 1}, P[3] = {0, 0, 0};
        for (int i = 0; i < nsp; ++i) {
          const int a = 3 - nsp + i;
          I[a] = x.shape[2 + i]; O[a] = y.shape[2 + i]; K[a] = k[i]; S[a] = s[i]; P[a] = pd[i];
        }
        q.ID = I[0]; q.IH = I[1]; q.IW = I[2]; q.OD = O[0]; q.OH = O[1]; q.OW = O[2];
        q.KD = K[0]; q.KH = K[1]; q.KW = K[2]; q.sD = S[0]; q.sH = S[1]; q.sW = S[2];
        q.pD = P[0]; q.pH = P[1]; q.pW = P[2];
        q.is_max = method == "MAX";
        break;
      }
      case Op::FC: {
        op.M = tensors_[op.in0].shape[0];
        op.N = Lp->params[0].shape[0];
        op.Kd = Lp->params[0].shape[1];
        op.w_dev = static_cast<float*>(dalloc((size_t)op.N * op.Kd * 4, true));
        op.b_dev = Lp->params.size() > 1 ? static_cast<float*>(dalloc((size_t)op.N * 4, true)) : nullptr;
        op.flops = 2.0 * op.M * op.N * op.Kd;
        break;
      }
      case Op::SSR: {
        const int C = tensors_[op.in0].C();
        op.scale_dev = static_cast<float*>(dalloc((size_t)C * 4, true));
        op.shift_dev = static_cast<float*>(dalloc((size_t)C * 4, true));
        break;
      }
      default:
        break;
    }
  }
  aux_.assign(ops_.size(), TrainAux{});
  for (size_t i = 0; i < ops_.size(); ++i) {
    Op& op = ops_[i];
    if (op.type == Op::LOSS || op.type == Op::ACCURACY) {
      const Tensor& x = tensors_[op.in0];
      TrainAux& a = aux_[i];
      const OrigLayer& L = layers_[op.layer];
      if (op.type == Op::LOSS) {
        const pt::Msg* lp = L.msg->msg("loss_param");
        ECO_CHECK(!lp || (!lp->has("ignore_label") && lp->boolean("normalize", true)),
                  "SoftmaxWithLoss " << L.name << ": ignore_label / normalize:false are not used on ECO's path");
        a.loss_weight = (float)L.msg->num("loss_weight", 1.0);  // loss layers default to weight 1 (loss_layer.cpp)
        a.prob = static_cast<float*>(dalloc((size_t)std::max<long long>(x.count(), 1) * 4, true));
      } else {
        const pt::Msg* ap = L.msg->msg("accuracy_param");
        a.top_k = ap ? (int)ap->integer("top_k", 1) : 1;
      }
    }
  }
  if (train_) plan_train();
  CUDA_OK(cudaStreamSynchronize(stream_));
  planned_ = true;
}

// =====================================================================================
// parameters -> device
static inline uint16_t f2bf(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)(u >> 16);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

void Net::upload_params() {
  if (train_) { upload_params_train(); return; }
  bool any = false;
  for (auto& L : layers_) any |= L.params_dirty;
  if (!any) return;
  // BN fold (the reference's own algebra: caffe_3d/python/gen_bn_inference.py:121-134):
  //   y = (x - mean) * (var + eps)^-1/2 * slope + bias  ==  x * scale + shift
  auto bn_fold = [&](const OrigLayer& B, std::vector<float>& scale, std::vector<float>& shift) {
    const pt::Msg* bp = B.msg->msg("bn_param");
    const float eps = bp ? (float)bp->num("eps", 1e-5) : 1e-5f;
    const size_t C = B.params[0].data.size();
    scale.resize(C);
    shift.resize(C);
    for (size_t i = 0; i < C; ++i) {
      const float inv_std = std::pow(B.params[3].data[i] + eps, -0.5f);
      scale[i] = B.params[0].data[i] * inv_std;
      shift[i] = B.params[1].data[i] - B.params[2].data[i] * scale[i];
    }
  };
  for (auto& op : ops_) {
    if (op.type == Op::CONV) {
      ConvOp& c = convs_[op.conv];
      OrigLayer& L = layers_[c.conv_layer];
      if (!c.members.empty()) {
        bool dirty = false;
        for (const ConvMember& m : c.members)
          dirty |= layers_[m.conv_layer].params_dirty || (m.bn_layer >= 0 && layers_[m.bn_layer].params_dirty);
        if (!dirty) continue;
        const ConvKernelParams& kp = c.kp;
        std::vector<uint16_t> wp((size_t)c.Cout_pad * c.Ktotal, 0);
        std::vector<float> bias(c.Cout, 0.f), scale(c.Cout, 1.f), shift(c.Cout, 0.f);
        std::vector<float> pbias(c.Cout, 0.f), pscale(c.Cout, 1.f), pshift(c.Cout, 0.f);
        for (const ConvMember& m : c.members) {
          const OrigLayer& ML = layers_[m.conv_layer];
          const std::vector<float>& w = ML.params[0].data;
          for (int o = 0; o < m.cout; ++o)
            for (int ch = 0; ch < c.Cin; ++ch) {
              const size_t kidx = ((size_t)(ch / kBlockK)) * kBlockK + ch % kBlockK;  // 1x1: one tap
              wp[(size_t)(m.off + o) * c.Ktotal + kidx] = f2bf(w[(size_t)o * c.Cin + ch]);
            }
          std::vector<float> sc, sh;
          if (m.bn_layer >= 0) bn_fold(layers_[m.bn_layer], sc, sh);
          for (int o = 0; o < m.cout; ++o) {
            const float b = ML.params.size() > 1 ? ML.params[1].data[o] : 0.f;
            float* B = m.post_pool ? pbias.data() : bias.data();
            float* S = m.post_pool ? pscale.data() : scale.data();
            float* H = m.post_pool ? pshift.data() : shift.data();
            B[m.off + o] = b;
            if (m.bn_layer >= 0) { S[m.off + o] = sc[o]; H[m.off + o] = sh[o]; }
          }
        }
        (void)kp;
        CUDA_OK(cudaMemcpyAsync(c.w_dev, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaMemcpyAsync(c.bias_dev, bias.data(), bias.size() * 4, cudaMemcpyHostToDevice, stream_));
        if (c.scale_dev) {
          CUDA_OK(cudaMemcpyAsync(c.scale_dev, scale.data(), scale.size() * 4, cudaMemcpyHostToDevice, stream_));
          CUDA_OK(cudaMemcpyAsync(c.shift_dev, shift.data(), shift.size() * 4, cudaMemcpyHostToDevice, stream_));
        }
        if (c.post_bias_dev) {
          CUDA_OK(cudaMemcpyAsync(c.post_bias_dev, pbias.data(), pbias.size() * 4, cudaMemcpyHostToDevice, stream_));
          CUDA_OK(cudaMemcpyAsync(c.post_scale_dev, pscale.data(), pscale.size() * 4, cudaMemcpyHostToDevice, stream_));
          CUDA_OK(cudaMemcpyAsync(c.post_shift_dev, pshift.data(), pshift.size() * 4, cudaMemcpyHostToDevice, stream_));
        }
        CUDA_OK(cudaStreamSynchronize(stream_));  // host vectors are temporaries
        continue;
      }
      const bool bn_dirty = c.bn_layer >= 0 && layers_[c.bn_layer].params_dirty;
      if (L.params_dirty) {
        const ConvKernelParams& kp = c.kp;
        std::vector<uint16_t> wp((size_t)c.Cout_pad * c.Ktotal, 0);
        const std::vector<float>& w = L.params[0].data;
        const int taps = c.K[0] * c.K[1] * c.K[2];
        if (c.stem) {
          for (int o = 0; o < c.Cout; ++o)
            for (int ty = 0; ty < 4; ++ty)
              for (int tx = 0; tx < 4; ++tx)
                for (int dy = 0; dy < 2; ++dy)
                  for (int dx = 0; dx < 2; ++dx)
                    for (int ch = 0; ch < 3; ++ch) {
                      const int ky = 2 * ty + dy, kx = 2 * tx + dx;
                      if (ky >= 7 || kx >= 7) continue;
                      const float v = w[(((size_t)o * 3 + ch) * 7 + ky) * 7 + kx];
                      wp[(size_t)o * c.Ktotal + (size_t)ty * 64 + tx * 16 + (dy * 2 + dx) * 3 + ch] = f2bf(v);
                    }
        } else {
          for (int o = 0; o < c.Cout; ++o)
            for (int ch = 0; ch < c.Cin; ++ch)
              for (int t = 0; t < taps; ++t) {
                const float v = w[((size_t)o * c.Cin + ch) * taps + t];
                if (!precision_) {
                  const size_t kidx = ((size_t)t * kp.cblocks + ch / kBlockK) * kBlockK + ch % kBlockK;
                  wp[(size_t)o * c.Ktotal + kidx] = f2bf(v);
                  continue;
                }
                // planes of A are [hi | lo | hi]: weights [w_hi | w_hi | w_lo] give hi*w_hi + lo*w_hi + hi*w_lo
                const uint16_t hi = f2bf(v);
                uint32_t hb = (uint32_t)hi << 16;
                float hf;
                std::memcpy(&hf, &hb, 4);
                const uint16_t lo = f2bf(v - hf);
                for (int pl = 0; pl < 3; ++pl) {
                  const int cc = pl * c.cin_stride + ch;
                  const size_t kidx = ((size_t)t * kp.cblocks + cc / kBlockK) * kBlockK + cc % kBlockK;
                  wp[(size_t)o * c.Ktotal + kidx] = pl == 2 ? lo : hi;
                }
              }
        }
        CUDA_OK(cudaMemcpyAsync(c.w_dev, wp.data(), wp.size() * 2, cudaMemcpyHostToDevice, stream_));
        if (L.params.size() > 1)
          CUDA_OK(cudaMemcpyAsync(c.bias_dev, L.params[1].data.data(), (size_t)c.Cout * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));  // wp is a temporary
      }
      if (bn_dirty) {
        std::vector<float> sc, sh;
        bn_fold(layers_[c.bn_layer], sc, sh);
        CUDA_OK(cudaMemcpyAsync(c.scale_dev, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaMemcpyAsync(c.shift_dev, sh.data(), sh.size() * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));
      }
    } else if (op.type == Op::FC) {
      OrigLayer& L = layers_[op.layer];
      if (L.params_dirty) {
        CUDA_OK(cudaMemcpyAsync(op.w_dev, L.params[0].data.data(), L.params[0].data.size() * 4, cudaMemcpyHostToDevice, stream_));
        if (op.b_dev)
          CUDA_OK(cudaMemcpyAsync(op.b_dev, L.params[1].data.data(), L.params[1].data.size() * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));
      }
    } else if (op.type == Op::SSR) {
      OrigLayer& L = layers_[op.layer];
      if (L.params_dirty || L.type == "ReLU") {
        std::vector<float> sc, sh;
        if (L.type == "BN") bn_fold(L, sc, sh);
        else { sc.assign(tensors_[op.in0].C(), 1.f); sh.assign(tensors_[op.in0].C(), 0.f); }
        CUDA_OK(cudaMemcpyAsync(op.scale_dev, sc.data(), sc.size() * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaMemcpyAsync(op.shift_dev, sh.data(), sh.size() * 4, cudaMemcpyHostToDevice, stream_));
        CUDA_OK(cudaStreamSynchronize(stream_));
      }
    }
  }
  for (auto& L : layers_) L.params_dirty = false;
}

// =====================================================================================
// host <-> device blob traffic (what SyncedMemory::to_cpu/to_gpu do lazily, syncedmem.cpp:21-70)
// per-net staging buffer for the fp32 <-> channels-last conversions (all users run on this net's stream, in order)
float* Net::staging(size_t bytes) {
  if (bytes > stage_bytes_) {
    if (stage_) {
      CUDA_OK(cudaStreamSynchronize(stream_));
      cudaFree(stage_);
      stage_ = nullptr;
    }
    CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&stage_), bytes));
    stage_bytes_ = bytes;
  }
  return stage_;
}

void Net::download(Tensor& t) {
  ECO_CHECK(t.materialized && t.dev, "blob '" << t.name << "' is fused away in the current plan and has no data; create "
                                                 "the net with option keep_all_blobs=1 to materialise every blob");
  const size_t n = (size_t)t.count();
  t.host.resize(n, true);
  if (n == 0) return;
  if (t.kind == Kind::CL) {
    float* st = staging(n * 4);
    CUDA_OK(launch_cl_to_f32(view(t), st, stream_));
    CUDA_OK(cudaMemcpyAsync(t.host.p, st, n * 4, cudaMemcpyDeviceToHost, stream_));
  } else {
    CUDA_OK(cudaMemcpyAsync(t.host.p, t.dev, n * 4, cudaMemcpyDeviceToHost, stream_));
  }
  CUDA_OK(cudaStreamSynchronize(stream_));
  t.dev_newer = false;
}

void Net::upload(Tensor& t) {
  const size_t n = (size_t)t.count();
  if (n == 0 || t.host.empty()) { t.host_newer = false; return; }
  if (!t.materialized || !t.dev) { t.host_newer = false; return; }  // nothing reads it on the device
  if (t.kind == Kind::CL) {
    float* st = staging(n * 4);
    CUDA_OK(cudaMemcpyAsync(st, t.host.p, n * 4, cudaMemcpyHostToDevice, stream_));
    CUDA_OK(launch_f32_to_cl(st, view(t), stream_));
  } else {
    CUDA_OK(cudaMemcpyAsync(t.dev, t.host.p, n * 4, cudaMemcpyHostToDevice, stream_));
  }
  // the copy reads the pinned mirror asynchronously: whoever writes the mirror next must wait for it (host_data)
  if (!t.h2d_done) CUDA_OK(cudaEventCreateWithFlags(&t.h2d_done, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(t.h2d_done, stream_));
  t.host_newer = false;
  t.dev_newer = false;
}

float* Net::host_data(int vb, bool for_write, size_t* count) {
  ECO_CHECK(vb >= 0 && vb < (int)vis_blobs_.size(), "blob index out of range");
  Tensor& t = tensors_[vis_blobs_[vb].tensor];
  const size_t n = (size_t)t.count();
  ECO_CHECK(!planned_ || t.materialized,
            "blob '" << t.name << "' is fused away in the current plan and has no data; create the net with option "
                        "keep_all_blobs=1 to materialise every blob");
  if (chunked_last_ && planned_) {
    bool io = false;
    for (int q : inputs_) io |= tensors_[vis_blobs_[q].tensor].root == t.root;
    for (int q : outputs_) io |= tensors_[vis_blobs_[q].tensor].root == t.root;
    ECO_CHECK(io, "blob '" << t.name << "' was not produced in this net's buffers: the last forward ran split into sub-batches "
                            "(option h2d_chunks); create the net with h2d_chunks=1 or keep_all_blobs=1 to inspect intermediate blobs");
  }
  if (t.dev_newer && planned_) download(t);
  if (t.host.n != n) t.host.resize(n, true);
  if (for_write && t.h2d_done) CUDA_OK(cudaEventSynchronize(t.h2d_done));  // a previous forward may still be reading the mirror
  if (for_write) t.host_newer = true;
  if (count) *count = n;
  return t.host.p;
}
float* Net::host_diff(int vb, bool for_write, size_t* count) {
  ECO_CHECK(vb >= 0 && vb < (int)vis_blobs_.size(), "blob index out of range");
  Tensor& t = tensors_[vis_blobs_[vb].tensor];
  const size_t n = (size_t)t.count();
  if (t.host_diff.n != n) t.host_diff.resize(n, false);
  if (planned_ && train_ && t.ddev && t.diff_dev_newer && n) {
    // cpu_diff(): bring the device gradient over (fp32, caffe layout)
    if (t.kind == Kind::CL) {
      float* st = staging(n * 4);
      ClView v = view(t);
      v.ptr = static_cast<__nv_bfloat16*>(t.ddev);
      CUDA_OK(launch_cl_to_f32(v, st, stream_));
      CUDA_OK(cudaMemcpyAsync(t.host_diff.p, st, n * 4, cudaMemcpyDeviceToHost, stream_));
    } else {
      CUDA_OK(cudaMemcpyAsync(t.host_diff.p, t.ddev, n * 4, cudaMemcpyDeviceToHost, stream_));
    }
    CUDA_OK(cudaStreamSynchronize(stream_));
    t.diff_dev_newer = false;
  }
  if (for_write) t.diff_host_newer = true;  // mutable_cpu_diff(): the next Backward starts from this gradient
  if (count) *count = n;
  return t.host_diff.p;
}

void Net::set_input_device(int vb, const void* dev, size_t count) {
  if (!planned_) plan();
  ECO_CHECK(vb >= 0 && vb < (int)vis_blobs_.size(), "blob index out of range");
  Tensor& t = tensors_[vis_blobs_[vb].tensor];
  ECO_CHECK(t.kind == Kind::F32 && t.dev, "set_input_device: '" << t.name << "' is not a plain fp32 input blob");
  ECO_CHECK(count == (size_t)t.count(), "set_input_device: size mismatch");
  CUDA_OK(cudaMemcpyAsync(t.dev, dev, count * 4, cudaMemcpyDeviceToDevice, stream_));
  t.host_newer = false;
  t.dev_newer = true;
}

const float* Net::device_f32(int vb, size_t* count) {
  if (!planned_) plan();
  Tensor& t = tensors_[vis_blobs_[vb].tensor];
  ECO_CHECK(t.kind == Kind::F32 && t.materialized && t.dev, "blob '" << t.name << "' is not a materialised plain fp32 blob");
  if (count) *count = (size_t)t.count();
  return static_cast<const float*>(t.dev);
}

void Net::push_frames(int dst_vb, Net& src, int src_vb) {
  if (!planned_) plan();
  if (!src.planned_) src.plan();
  ECO_CHECK(dst_vb >= 0 && dst_vb < (int)vis_blobs_.size() && src_vb >= 0 && src_vb < (int)src.vis_blobs_.size(), "blob index out of range");
  Tensor& d = tensors_[vis_blobs_[dst_vb].tensor];
  Tensor& s = src.tensors_[src.vis_blobs_[src_vb].tensor];
  ECO_CHECK(d.kind == Kind::CL && s.kind == Kind::CL && d.dev && s.dev && d.materialized && s.materialized,
            "push_frames: both blobs must be materialised feature maps ('" << d.name << "', '" << s.name << "')");
  ECO_CHECK(d.ch_axis == 1 && s.ch_axis == 1 && d.coff == 0 && s.coff == 0 && d.cs == d.C() && s.cs == s.C() && d.C() == s.C() &&
                d.inner() == s.inner(), "push_frames: blobs must be dense [frames, C, ...] maps of equal frame size");
  const int planes = precision_ ? 3 : 1;
  ECO_CHECK((precision_ != 0) == (src.precision_ != 0), "push_frames: nets differ in precision mode");
  const size_t row = (size_t)d.inner() * (size_t)d.cs * 2 * planes;
  const long long F = d.outer(), k = s.outer();
  ECO_CHECK(k >= 1 && k <= F, "push_frames: source has " << k << " frames, destination " << F);
  if (!push_event_) CUDA_OK(cudaEventCreateWithFlags(&push_event_, cudaEventDisableTiming));
  CUDA_OK(cudaEventRecord(push_event_, src.stream_));
  CUDA_OK(cudaStreamWaitEvent(stream_, push_event_, 0));
  char* base = static_cast<char*>(d.dev);
  if (k < F) {  // overlapping move through the staging buffer (at most N - 1 frames of 150 KB)
    const size_t tail = (size_t)(F - k) * row;
    char* st = reinterpret_cast<char*>(staging(tail));
    CUDA_OK(cudaMemcpyAsync(st, base + (size_t)k * row, tail, cudaMemcpyDeviceToDevice, stream_));
    CUDA_OK(cudaMemcpyAsync(base, st, tail, cudaMemcpyDeviceToDevice, stream_));
  }
  CUDA_OK(cudaMemcpyAsync(base + (size_t)(F - k) * row, s.dev, (size_t)k * row, cudaMemcpyDeviceToDevice, stream_));
  // the source net must not overwrite its blob (next forward on ITS stream) before this copy has read it
  CUDA_OK(cudaEventRecord(push_event_, stream_));
  CUDA_OK(cudaStreamWaitEvent(src.stream_, push_event_, 0));
  mark_written(vis_blobs_[dst_vb].tensor);
}

void Net::transform_input_u8(int vb, const unsigned char* src, int B, int C, int H, int W, const eco_clip_transform* t,
                             const eco_transform_param& p) {
  if (!planned_) plan();
  ECO_CHECK(vb >= 0 && vb < (int)vis_blobs_.size(), "blob index out of range");
  Tensor& x = tensors_[vis_blobs_[vb].tensor];
  ECO_CHECK(x.kind == Kind::F32 && x.dev && x.shape.size() == 4, "transform_input_u8: '" << x.name << "' is not a plain fp32 input blob");
  ECO_CHECK(x.shape[0] == B && x.shape[1] == C && x.shape[2] == x.shape[3], "transform_input_u8: blob is "
                << x.shape[0] << "x" << x.shape[1] << "x" << x.shape[2] << "x" << x.shape[3] << ", clips are " << B << "x" << C);
  const int crop = x.shape[2];
  for (int b = 0; b < B; ++b)
    ECO_CHECK(t[b].h_off >= 0 && t[b].w_off >= 0 && t[b].crop_h > 0 && t[b].crop_w > 0 && t[b].h_off + t[b].crop_h <= H &&
                  t[b].w_off + t[b].crop_w <= W, "transform_input_u8: crop window of clip " << b << " leaves the " << H << "x" << W << " datum");
  const size_t img = (size_t)B * C * H * W;
  const size_t tb = (size_t)B * sizeof(eco_clip_transform), mb = 16 * sizeof(float);
  const size_t need = ((img + 255) / 256) * 256 + ((tb + 255) / 256) * 256 + mb;
  if (need > xf_src_bytes_) {
    if (xf_src_) { CUDA_OK(cudaStreamSynchronize(stream_)); cudaFree(xf_src_); xf_src_ = nullptr; }
    CUDA_OK(cudaMalloc(reinterpret_cast<void**>(&xf_src_), need));
    xf_src_bytes_ = need;
  }
  unsigned char* d_img = xf_src_;
  eco_clip_transform* d_t = reinterpret_cast<eco_clip_transform*>(xf_src_ + ((img + 255) / 256) * 256);
  float* d_mean = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(d_t) + ((tb + 255) / 256) * 256);
  CUDA_OK(cudaMemcpyAsync(d_img, src, img, cudaMemcpyHostToDevice, stream_));
  CUDA_OK(cudaMemcpyAsync(d_t, t, tb, cudaMemcpyHostToDevice, stream_));
  const int nmean = std::max(0, std::min(16, p.num_mean));
  if (nmean) CUDA_OK(cudaMemcpyAsync(d_mean, p.mean_value, (size_t)nmean * 4, cudaMemcpyHostToDevice, stream_));
  CUDA_OK(launch_transform_u8(d_img, static_cast<float*>(x.dev), B, C, H, W, crop, d_t, d_mean, nmean, p.scale == 0.f ? 1.f : p.scale,
                              p.is_flow, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));  // src / t are the caller's (possibly pageable) memory
  x.host_newer = false;
  x.dev_newer = true;
  for (auto& q : tensors_)
    if (q.root == x.root && q.materialized) q.dev_newer = true;
}

void Net::sync() {
  if (stream_) CUDA_OK(cudaStreamSynchronize(stream_));
  if (error_flag_dev_) {
    int flag = 0;
    CUDA_OK(cudaMemcpy(&flag, error_flag_dev_, sizeof(int), cudaMemcpyDeviceToHost));
    ECO_CHECK(flag == 0, "conv kernel pipeline timed out (code " << flag << ")");
  }
}

// =====================================================================================
// fp32 NCHW net input -> bf16 operand of the first convolution (space-to-depth cells for the 7x7 stem,
// channel-padded channels-last otherwise).  `src` overrides the input blob's device buffer (pipelined
// forward reads straight from a staging slot).
void Net::run_input_xform(ConvOp& c, const float* src) {
  Tensor& x = tensors_[c.in_tensor];
  if (c.direct_in) {
    // no transform kernel: the stem rows kernel reads the frames (this launch IS the convolution)
    StemRowsParams r = c.rp;
    if (u8_src_) {
      r.src_mode = 2; r.src = u8_src_;
      r.mean0 = u8_mean_[0]; r.mean1 = u8_mean_[1]; r.mean2 = u8_mean_[2];
    } else {
      r.src_mode = 1; r.src = src ? src : static_cast<const float*>(x.dev);
    }
    CUDA_OK(launch_stem_rows(r, c.tmX, c.tmB, stream_));
    return;
  }
  if (u8_src_) {
    // raw uint8 frames: mean subtraction fused into the transform (stem) or into a conversion pass
    const unsigned char* u8 = static_cast<const unsigned char*>(u8_src_);
    if (c.stem) {
      CUDA_OK(launch_stem_s2d_u8(u8, c.stem_in, c.NB, c.I[1], c.I[2], c.stem_CH, c.stem_CW, u8_mean_[0], u8_mean_[1],
                                 u8_mean_[2], stream_));
      return;
    }
    CUDA_OK(launch_u8_to_f32_mean(u8, static_cast<float*>(x.dev), c.NB, c.Cin, (long long)c.I[0] * c.I[1] * c.I[2],
                                  u8_mean_[0], u8_mean_[1], u8_mean_[2], u8_mean_[3], stream_));
    src = nullptr;
  }
  const float* in = src ? src : static_cast<const float*>(x.dev);
  if (c.stem) {
    CUDA_OK(launch_stem_s2d(in, c.stem_in, c.NB, c.I[1], c.I[2], c.stem_CH, c.stem_CW, stream_));
  } else {
    ClView v;
    v.ptr = c.stem_in;
    v.outer = c.NB;
    v.inner = (long long)c.I[0] * c.I[1] * c.I[2];
    v.C = c.Cin;
    v.cs = precision_ ? c.cin_stride : c.Cin_k;
    v.coff = 0;
    v.seg = precision_ ? c.cin_stride : 0;
    CUDA_OK(launch_f32_to_cl(in, v, stream_));
  }
}

void Net::run_op(Op& op, bool with_xform) {
  switch (op.type) {
    case Op::CONV: {
      ConvOp& c = convs_[op.conv];
      if (c.stem_in && with_xform) run_input_xform(c, nullptr);
      if (c.direct_in) {}  // launched by run_input_xform (outside the CUDA graph: its source pointer changes per call)
      else if (c.rows) CUDA_OK(launch_stem_rows(c.rp, c.tmX, c.tmB, stream_));
      else if (c.pair) CUDA_OK(launch_conv_pair(c.kpp, c.tmA, c.tmBh, stream_));
      else if (c.halo) CUDA_OK(launch_conv_halo(c.hp, c.halo_mt, c.tmX, c.tmB, stream_));
      else CUDA_OK(launch_conv_umma(c.kp, c.tmA, c.kp.multicast ? c.tmBh : c.tmB, stream_));
      if (c.pool_tensor >= 0) mark_written(c.pool_tensor);
      for (const ConvMember& m : c.members) mark_written(m.tensor);
      if (c.out_tensor >= 0) mark_written(c.out_tensor);
      if (c.raw_tensor >= 0) mark_written(c.raw_tensor);
      break;
    }
    case Op::POOL_CL:
      if (precision_) CUDA_OK(launch_pool_cl_split(op.pool, tensors_[op.in0].cs, tensors_[op.out].cs, stream_));
      else CUDA_OK(launch_pool_cl(op.pool, stream_));
      break;
    case Op::GLOBAL_AVG:
      CUDA_OK(launch_global_avg_cl(view(tensors_[op.in0]), static_cast<float*>(tensors_[op.out].dev), stream_));
      break;
    case Op::POOL_F32:
      CUDA_OK(launch_pool_f32(op.poolf, stream_));
      break;
    case Op::FC:
      CUDA_OK(launch_inner_product(static_cast<const float*>(tensors_[op.in0].dev), op.w_dev, op.b_dev,
                                   static_cast<float*>(tensors_[op.out].dev), op.M, op.N, op.Kd, stream_));
      break;
    case Op::SSR:
      ECO_CHECK(!precision_, "precision=1: stand-alone BN / ReLU layers (" << op.name << ") are not supported; ECO's are fused");
      CUDA_OK(launch_scale_shift_relu_cl(view(tensors_[op.in0]), view(tensors_[op.out]), op.scale_dev, op.shift_dev,
                                         op.relu ? 1 : 0, stream_));
      break;
    case Op::ELTWISE:
      ECO_CHECK(!precision_, "precision=1: stand-alone Eltwise (" << op.name << ") is not supported; ECO's are fused");
      CUDA_OK(launch_eltwise_sum_cl(view(tensors_[op.in0]), view(tensors_[op.in1]), view(tensors_[op.out]), stream_));
      break;
    case Op::COPY2D: {
      ECO_CHECK(!precision_ || tensors_[op.out].kind == Kind::F32, "precision=1: Concat " << op.name << " needs a copy of a feature map");
      const char* src = static_cast<const char*>(tensors_[op.in0].dev) + op.src_off;
      char* dst = static_cast<char*>(tensors_[op.out].dev) + op.dst_off;
      CUDA_OK(cudaMemcpy2DAsync(dst, op.dst_pitch, src, op.src_pitch, op.width_bytes, op.rows, cudaMemcpyDeviceToDevice,
                                stream_));
      break;
    }
    case Op::CL_TO_F32:
      CUDA_OK(launch_cl_to_f32(view(tensors_[op.in0]), static_cast<float*>(tensors_[op.out].dev), stream_));
      break;
    case Op::F32_TO_CL:
      CUDA_OK(launch_f32_to_cl(static_cast<const float*>(tensors_[op.in0].dev), view(tensors_[op.out]), stream_));
      break;
    case Op::BN_TRAIN:
    case Op::DROPOUT:
    case Op::LOSS:
    case Op::ACCURACY:
      run_train_op(op, aux_[&op - ops_.data()]);
      break;
    case Op::SOFTMAX:
      CUDA_OK(launch_softmax_f32(static_cast<const float*>(tensors_[op.in0].dev), static_cast<float*>(tensors_[op.out].dev),
                                 tensors_[op.in0].shape[0], (int)(tensors_[op.in0].count() / std::max(1, tensors_[op.in0].shape[0])),
                                 stream_));
      break;
  }
  if (op.out >= 0) mark_written(op.out);
}

// every tensor that shares the written buffer (views, Split copies, concat slices <-> concat top)
void Net::mark_written(int tid) {
  const int root = tensors_[tid].root;
  for (auto& t : tensors_)
    if (t.root == root && t.materialized) t.dev_newer = true;
}

// Runs the planned ops on stream_.  Full forwards may replay a CUDA graph; the input transforms stay
// outside the graph so their source pointer can change from call to call.
void Net::run_ops(bool full, int lo, int hi, const float* input_override, int* launches) {
  int n = 0;
  if (full && use_graph_ && !train_) {
    for (auto& op : ops_)
      if (op.type == Op::CONV && convs_[op.conv].stem_in) run_input_xform(convs_[op.conv], input_override);
    if (!graph_valid_) {
      cudaGraph_t g = nullptr;
      CUDA_OK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
      for (auto& op : ops_) run_op(op, false);
      CUDA_OK(cudaStreamEndCapture(stream_, &g));
      if (graph_exec_) cudaGraphExecDestroy(graph_exec_);
      CUDA_OK(cudaGraphInstantiate(&graph_exec_, g, 0));
      cudaGraphDestroy(g);
      graph_valid_ = true;
    } else {
      std::vector<char> is_input(tensors_.size(), 0);
      for (int vb : inputs_) is_input[vis_blobs_[vb].tensor] = 1;
      for (size_t i = 0; i < tensors_.size(); ++i)
        if (tensors_[i].materialized && !is_input[i]) tensors_[i].dev_newer = true;
    }
    for (auto& op : ops_) n += op.launches;
    CUDA_OK(cudaGraphLaunch(graph_exec_, stream_));
  } else {
    for (auto& op : ops_) {
      if (!full && (op.last_layer < lo || op.first_layer > hi)) continue;
      if (op.type == Op::CONV && convs_[op.conv].stem_in) {
        run_input_xform(convs_[op.conv], input_override);
        run_op(op, false);
      } else {
        run_op(op, true);
      }
      n += op.launches;
    }
  }
  if (launches) *launches = n;
}

// Serving extension (not in caffe): overlap the host->device copy of clip k+1 with the compute of clip k.
// The caller owns two (pinned) input buffers and alternates them; results are copied to `host_out`
// asynchronously and are valid after wait_ticket().
int Net::forward_pipelined(const float* host_in, size_t count, float* host_out, size_t out_count) {
  return forward_pipelined_any(host_in, count, 4, nullptr, host_out, out_count);
}
int Net::forward_pipelined_u8(const unsigned char* host_in, size_t count, const float* mean, int nmean, float* host_out,
                              size_t out_count) {
  float m[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4 && i < nmean; ++i) m[i] = mean[i];
  return forward_pipelined_any(host_in, count, 1, m, host_out, out_count);
}

int Net::forward_pipelined_any(const void* host_in, size_t count, int elem_bytes, const float* mean4, float* host_out,
                               size_t out_count) {
  if (!planned_) plan();
  upload_params();
  ECO_CHECK(inputs_.size() >= 1 && outputs_.size() >= 1, "forward_pipelined needs one input and one output blob");
  Tensor& tin = tensors_[vis_blobs_[inputs_[0]].tensor];
  Tensor& tout = tensors_[vis_blobs_[outputs_[0]].tensor];
  ECO_CHECK(tin.kind == Kind::F32 && count == (size_t)tin.count(), "forward_pipelined: input size mismatch");
  ECO_CHECK(pipe_elem_bytes_ == 0 || pipe_elem_bytes_ == elem_bytes,
            "forward_pipelined: do not mix fp32 and uint8 calls on one net without a reshape");
  pipe_elem_bytes_ = elem_bytes;
  ECO_CHECK(tout.kind == Kind::F32 && tout.dev && out_count == (size_t)tout.count(), "forward_pipelined: output size mismatch");
  if (!copy_stream_) {
    CUDA_OK(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      pipe_slot_[i] = static_cast<float*>(dalloc(count * (size_t)elem_bytes, false));
      CUDA_OK(cudaEventCreateWithFlags(&ev_h2d_[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&ev_slot_free_[i], cudaEventDisableTiming));
      CUDA_OK(cudaEventCreateWithFlags(&ev_done_[i], cudaEventDisableTiming));
    }
  }
  const int slot = (int)(pipe_iter_ & 1);
  if (pipe_iter_ >= 2) CUDA_OK(cudaStreamWaitEvent(copy_stream_, ev_slot_free_[slot], 0));
  CUDA_OK(cudaMemcpyAsync(pipe_slot_[slot], host_in, count * (size_t)elem_bytes, cudaMemcpyHostToDevice, copy_stream_));
  CUDA_OK(cudaEventRecord(ev_h2d_[slot], copy_stream_));
  CUDA_OK(cudaStreamWaitEvent(stream_, ev_h2d_[slot], 0));
  bool xform_only_input = false;
  for (auto& op : ops_)
    if (op.type == Op::CONV && convs_[op.conv].stem_in && convs_[op.conv].in_tensor == vis_blobs_[inputs_[0]].tensor)
      xform_only_input = true;
  ECO_CHECK(xform_only_input, "forward_pipelined: the net input must feed a convolution directly");
  int launches = 0;
  if (mean4) {
    u8_src_ = pipe_slot_[slot];
    for (int i = 0; i < 4; ++i) u8_mean_[i] = mean4[i];
  }
  try {
    run_ops(true, 0, 0, mean4 ? nullptr : pipe_slot_[slot], &launches);  // transforms read the slot; the graph covers the rest
  } catch (...) {
    u8_src_ = nullptr;
    throw;
  }
  u8_src_ = nullptr;
  // the slot is free once the transform has consumed it; conservatively: once this forward is enqueued
  // up to here the transform is the first kernel, so record right after the whole enqueue is cheap too
  CUDA_OK(cudaEventRecord(ev_slot_free_[slot], stream_));
  CUDA_OK(cudaMemcpyAsync(host_out, tout.dev, out_count * 4, cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaEventRecord(ev_done_[slot], stream_));
  last_launches_ = launches;
  return (int)(pipe_iter_++ & 0x7fffffff);
}

void Net::wait_ticket(int ticket) {
  ECO_CHECK(copy_stream_ != nullptr, "wait_ticket without forward_pipelined");
  CUDA_OK(cudaEventSynchronize(ev_done_[ticket & 1]));
}

float Net::forward(int start, int end) {
  if (!planned_) plan();
  upload_params();
  const int NV = (int)vis_layers_.size();
  if (end < 0) end = NV - 1;
  ECO_CHECK(start >= 0 && end < NV && start <= end, "forward range [" << start << "," << end << "] out of bounds");
  // visible layer range -> original layer range
  int lo = (int)layers_.size(), hi = -1;
  for (int v = start; v <= end; ++v)
    if (vis_layers_[v].orig >= 0) {
      lo = std::min(lo, vis_layers_[v].orig);
      hi = std::max(hi, vis_layers_[v].orig);
    }
  const bool full = (start == 0 && end == NV - 1);
  chunked_last_ = false;
  if (full) {
    int cl = 0;
    if (try_chunked_forward(&cl)) {
      last_launches_ = cl;
      last_loss_ = 0.f;
      return 0.f;
    }
  }
  // host-modified blobs go up first (net inputs are always re-sent: the caller owns that memory)
  for (int vb : inputs_) {
    Tensor& t = tensors_[vis_blobs_[vb].tensor];
    if (!t.host.empty() && !t.dev_newer) t.host_newer = true;
  }
  for (auto& t : tensors_)
    if (t.host_newer && t.root >= 0) upload(t);
  // (an input written on the device -- set_input_device, transform_input_u8 -- keeps dev_newer: its host mirror is stale
  // until somebody asks for it, and the device copy stays the one the next forward reads)
  // blobs that are views of an input (reshape_data of the train/test nets) read their host mirror back from the device
  for (int vb : inputs_) {
    const int it = vis_blobs_[vb].tensor;
    for (size_t q = 0; q < tensors_.size(); ++q)
      if ((int)q != it && tensors_[q].root == tensors_[it].root && tensors_[q].materialized) tensors_[q].dev_newer = true;
  }

  if (train_) ++train_iter_;  // a new dropout mask per forward pass
  int launches = 0;
  run_ops(full, lo, hi, nullptr, &launches);
  last_launches_ = launches;
  // loss = sum over loss layers of loss_weight * top (Net::ForwardFromTo, net.cpp:566-583); reading it synchronises
  float loss = 0.f;
  bool any_loss = false;
  for (size_t i = 0; i < ops_.size(); ++i)
    if (ops_[i].type == Op::LOSS && !(ops_[i].last_layer < lo || ops_[i].first_layer > hi)) any_loss = true;
  if (any_loss) {
    if (!loss_host_) CUDA_OK(cudaMallocHost(reinterpret_cast<void**>(&loss_host_), 64 * sizeof(float)));
    int k = 0;
    for (size_t i = 0; i < ops_.size() && k < 64; ++i)
      if (ops_[i].type == Op::LOSS)
        CUDA_OK(cudaMemcpyAsync(loss_host_ + k++, tensors_[ops_[i].out].dev, 4, cudaMemcpyDeviceToHost, stream_));
    CUDA_OK(cudaStreamSynchronize(stream_));
    k = 0;
    for (size_t i = 0; i < ops_.size() && k < 64; ++i)
      if (ops_[i].type == Op::LOSS) loss += aux_[i].loss_weight * loss_host_[k++];
  }
  last_loss_ = loss;
  return loss;
}

std::string Net::describe_plan() {
  if (!planned_) plan();
  std::ostringstream o;
  static const char* tn[] = {"conv", "pool_cl", "global_avg", "pool_f32", "fc", "ssr", "eltwise", "copy2d", "cl_to_f32",
                             "f32_to_cl", "softmax", "bn_train", "dropout", "loss", "accuracy"};
  for (const Op& op : ops_) {
    o << op.name << " type=" << tn[(int)op.type];
    if (op.type == Op::CONV) {
      const ConvOp& c = convs_[op.conv];
      const ConvKernelParams& k = c.pair ? c.kpp : c.kp;
      const char* kern = c.rows ? "stem_rows" : c.pair ? "pair" : c.halo ? "halo" : k.persistent ? "persistent" : "one_tile";
      const int mt = c.pair ? 2 : (k.persistent ? k.m_halves : 1);
      const long long tile_m = (long long)kBlockM * mt;
      const long long tiles = c.rows ? (long long)c.rp.F * c.rp.strips
                                     : ((long long)k.M + tile_m - 1) / tile_m * ((c.Cout + k.block_n - 1) / k.block_n);
      o << " kernel=" << kern << " M=" << k.M << " Cout=" << c.Cout << " Cin=" << c.Cin << " taps=" << c.K[0] * c.K[1] * c.K[2]
        << " block_n=" << k.block_n << " mt=" << mt << " tiles=" << tiles << " stages=" << k.stages
        << " a_mode=" << k.a_mode << " nseg=" << k.nseg << " multicast=" << k.multicast << " raw=" << (c.raw_tensor >= 0)
        << " out=" << (c.out_tensor >= 0) << " res=" << (c.res_tensor >= 0) << " pool=" << (c.pool_tensor >= 0);
    }
    o << "\n";
  }
  return o.str();
}

int Net::profile(eco_op_time* out, int cap) {
  if (!planned_) plan();
  upload_params();
  for (auto& t : tensors_)
    if (t.host_newer && t.root >= 0) upload(t);
  std::vector<cudaEvent_t> ev(ops_.size() + 1);
  for (auto& e : ev) CUDA_OK(cudaEventCreate(&e));
  CUDA_OK(cudaEventRecord(ev[0], stream_));
  for (size_t i = 0; i < ops_.size(); ++i) {
    run_op(ops_[i]);
    CUDA_OK(cudaEventRecord(ev[i + 1], stream_));
  }
  CUDA_OK(cudaStreamSynchronize(stream_));
  op_names_.resize(ops_.size());
  int n = 0;
  for (size_t i = 0; i < ops_.size() && n < cap; ++i, ++n) {
    float ms = 0;
    CUDA_OK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
    op_names_[i] = ops_[i].name;
    out[n].name = op_names_[i].c_str();
    out[n].kind = ops_[i].type == Op::CONV ? 0 : (ops_[i].type == Op::COPY2D ? 2 : 1);
    out[n].ms = ms;
    out[n].flops = ops_[i].flops;
    out[n].bytes = ops_[i].bytes;
  }
  for (auto& e : ev) cudaEventDestroy(e);
  return n;
}

#include "net_train.inc"

}  // namespace eco
