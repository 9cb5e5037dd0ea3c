// solver.cpp -- see solver.hpp.
#include "solver.hpp"

#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace eco {

#define SOLVER_CHECK(cond, msg)                                                   \
  do {                                                                            \
    if (!(cond)) {                                                                \
      std::ostringstream _o;                                                      \
      _o << msg << "  [" << #cond << " @ " << __FILE__ << ":" << __LINE__ << "]"; \
      throw std::runtime_error(_o.str());                                         \
    }                                                                             \
  } while (0)
#define SOLVER_CUDA(expr)                                                                                   \
  do {                                                                                                      \
    cudaError_t _e = (expr);                                                                                \
    if (_e != cudaSuccess) throw std::runtime_error(std::string("CUDA error in " #expr ": ") + cudaGetErrorString(_e)); \
  } while (0)

static std::string read_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("Could not open " + path);
  std::stringstream ss;
  ss << f.rdbuf();
  return ss.str();
}

Solver::Solver(const std::string& solver_text, const std::string& net_text_override, const std::string& base_dir) {
  auto sp = pt::parse(solver_text);
  const pt::Msg& p = *sp;
  base_lr_ = (float)p.num("base_lr", 0.01);
  lr_policy_ = p.str("lr_policy", "fixed");
  gamma_ = (float)p.num("gamma", 0.1);
  power_ = (float)p.num("power", 1.0);
  momentum_ = (float)p.num("momentum", 0.0);
  weight_decay_ = (float)p.num("weight_decay", 0.0);
  clip_gradients_ = (float)p.num("clip_gradients", -1.0);
  stepsize_ = (int)p.integer("stepsize", 1);
  max_iter_ = (int)p.integer("max_iter", 0);
  iter_size_ = (int)p.integer("iter_size", 1);
  snapshot_ = (int)p.integer("snapshot", 0);
  display_ = (int)p.integer("display", 0);
  snapshot_prefix_ = p.str("snapshot_prefix", "");
  for (long v : p.integers("stepvalue")) stepvalue_.push_back((int)v);
  // solver_type: SGD (0) / NESTEROV (1) / ADAGRAD (2) enum of this caffe version (caffe.proto:196-202); newer `type: "..."`
  std::string st = p.str("solver_type", p.str("type", "SGD"));
  if (st == "0") st = "SGD";
  if (st == "1") st = "NESTEROV";
  for (auto& ch : st) ch = (char)std::toupper((unsigned char)ch);
  SOLVER_CHECK(st == "SGD" || st == "NESTEROV", "solver type " << st << " is not on ECO's path (its solvers are SGD / NESTEROV)");
  type_ = st;
  SOLVER_CHECK(p.str("regularization_type", "L2") == "L2", "only L2 regularisation is on ECO's path");
  SOLVER_CHECK(iter_size_ >= 1, "iter_size must be >= 1");
  std::string net_text = net_text_override;
  if (net_text.empty()) {
    net_path_ = p.str("net", p.str("train_net", ""));
    SOLVER_CHECK(!net_path_.empty(), "solver prototxt names no net / train_net");
    std::string path = net_path_;
    if (!path.empty() && path[0] != '/' && !base_dir.empty()) path = base_dir + "/" + path;
    net_text = read_file(path);
  }
  net_.reset(new Net(net_text, 0 /* caffe::TRAIN */));
}

Solver::~Solver() {
  if (hist_) cudaFree(hist_);
  if (scalars_) cudaFree(scalars_);
}

void Solver::ensure_history() {
  size_t count = 0;
  net_->param_arena(&count);
  if (hist_ && hist_count_ == count) return;
  if (hist_) cudaFree(hist_);
  SOLVER_CUDA(cudaMalloc(reinterpret_cast<void**>(&hist_), std::max<size_t>(count, 1) * 4));
  SOLVER_CUDA(cudaMemsetAsync(hist_, 0, std::max<size_t>(count, 1) * 4, net_->stream()));
  hist_count_ = count;
  if (!scalars_) SOLVER_CUDA(cudaMalloc(reinterpret_cast<void**>(&scalars_), 16));
}

float Solver::learning_rate() const {  // SGDSolver::GetLearningRate, solver.cpp:580-620
  const float it = (float)iter_;
  if (lr_policy_ == "fixed") return base_lr_;
  if (lr_policy_ == "step") { current_step_ = iter_ / stepsize_; return base_lr_ * std::pow(gamma_, (float)current_step_); }
  if (lr_policy_ == "exp") return base_lr_ * std::pow(gamma_, it);
  if (lr_policy_ == "inv") return base_lr_ * std::pow(1.f + gamma_ * it, -power_);
  if (lr_policy_ == "multistep") {
    if (current_step_ < (int)stepvalue_.size() && iter_ >= stepvalue_[current_step_]) ++current_step_;
    return base_lr_ * std::pow(gamma_, (float)current_step_);
  }
  if (lr_policy_ == "poly") return base_lr_ * std::pow(1.f - it / (float)max_iter_, power_);
  if (lr_policy_ == "sigmoid") return base_lr_ * (1.f / (1.f + std::exp(-gamma_ * (it - (float)stepsize_))));
  if (lr_policy_ == "exp10") return base_lr_ * std::pow(10.f, -it / (float)stepsize_);
  throw std::runtime_error("Unknown learning rate policy: " + lr_policy_);
}

// ClipGradients -> Normalize -> Regularize -> ComputeUpdateValue -> Net::Update, one pass per parameter blob
void Solver::apply_update() {
  ensure_history();
  size_t count = 0;
  float* P = net_->param_arena(&count);
  float* G = net_->grad_arena(&count);
  cudaStream_t st = net_->stream();
  const float rate = learning_rate();
  const float pre = 1.f / (float)world_;            // SyncGradient's 1 / MPI_all_rank (solver.cpp:332-337) ...
  const float norm = pre / (float)iter_size_;       // ... then Normalize's 1 / iter_size
  const float* clip_dev = nullptr;
  if (clip_gradients_ >= 0.f) {
    // global L2 norm over every parameter diff as it stands after the exchange (solver.cpp:637-660)
    SOLVER_CUDA(cudaMemsetAsync(scalars_, 0, 8, st));
    SOLVER_CUDA(launch_sumsq(G, (long long)count, scalars_, st));   // padding between slots is zero
    SOLVER_CUDA(launch_clip_factor(scalars_, clip_gradients_, pre, scalars_ + 1, st));
    clip_dev = scalars_ + 1;
  }
  const int nesterov = type_ == "NESTEROV" ? 1 : 0;
  for (const ParamSlot& sl : net_->param_slots()) {
    if (sl.count == 0) continue;
    const float local_rate = rate * sl.lr_mult;                 // params_lr (net.cpp:131-160)
    const float local_decay = weight_decay_ * sl.decay_mult;
    if (local_rate == 0.f && momentum_ == 0.f) continue;
    if (sl.lr_mult == 0.f) continue;  // BN running statistics / frozen blobs: diff stays 0, nothing to do
    SOLVER_CUDA(launch_sgd_update(P + sl.off, G + sl.off, hist_ + sl.off, (long long)sl.count, local_rate, momentum_, local_decay,
                                  norm, clip_dev, nesterov, st));
  }
  net_->params_updated_on_device();
  ++iter_;
}

float Solver::step(int iters) {
  ensure_history();
  float loss = 0.f;
  const int stop = iter_ + iters;
  while (iter_ < stop) {
    net_->clear_param_diffs();                                   // solver.cpp:178-195
    loss = 0.f;
    for (int i = 0; i < iter_size_; ++i) {
      loss += net_->forward(0, -1);                              // Net::ForwardBackward
      net_->backward(-1, -1);
    }
    if (sync_fn_) sync_fn_(sync_user_);                          // gradient exchange (reference: per-blob MPI_Allreduce)
    loss /= (float)iter_size_;
    if (display_ && iter_ % display_ == 0)
      std::fprintf(stderr, "Iteration %d, lr = %g, loss = %g\n", iter_, (double)learning_rate(), (double)loss);
    apply_update();
    if (snapshot_ && iter_ % snapshot_ == 0) snapshot("");
  }
  return loss;
}

void Solver::snapshot(const std::string& prefix_override) {
  const std::string prefix = prefix_override.empty() ? snapshot_prefix_ : prefix_override;
  char tag[64];
  std::snprintf(tag, sizeof(tag), "_iter_%d", iter_);
  const std::string model = prefix + tag + ".caffemodel";          // Solver::Snapshot, solver.cpp:521-546
  net_->save(model);
  ensure_history();
  std::vector<float> h(hist_count_);
  SOLVER_CUDA(cudaMemcpyAsync(h.data(), hist_, hist_count_ * 4, cudaMemcpyDeviceToHost, net_->stream()));
  SOLVER_CUDA(cudaStreamSynchronize(net_->stream()));
  std::vector<std::pair<std::vector<int>, std::vector<float>>> blobs;
  for (const ParamSlot& sl : net_->param_slots()) {
    const ParamBlob& b = net_->layers_[sl.layer].params[sl.idx];
    blobs.emplace_back(b.shape, std::vector<float>(h.begin() + (long)sl.off, h.begin() + (long)(sl.off + sl.count)));
  }
  write_solverstate(prefix + tag + ".solverstate", iter_, model, current_step_, blobs);
}

void Solver::restore(const std::string& state_file) {
  int it = 0, cs = 0;
  std::string learned;
  std::vector<std::vector<float>> hist;
  read_solverstate(state_file, &it, &learned, &cs, &hist);
  if (!learned.empty()) net_->copy_from(learned);                   // Solver::Restore, solver.cpp:549-560
  ensure_history();
  const auto& slots = net_->param_slots();
  SOLVER_CHECK(hist.size() == slots.size(), "Incorrect length of history blobs: " << hist.size() << " vs " << slots.size());
  std::vector<float> h(hist_count_, 0.f);
  for (size_t i = 0; i < slots.size(); ++i) {
    SOLVER_CHECK(hist[i].size() == slots[i].count, "history blob " << i << " has the wrong size");
    std::copy(hist[i].begin(), hist[i].end(), h.begin() + (long)slots[i].off);
  }
  SOLVER_CUDA(cudaMemcpyAsync(hist_, h.data(), hist_count_ * 4, cudaMemcpyHostToDevice, net_->stream()));
  SOLVER_CUDA(cudaStreamSynchronize(net_->stream()));
  iter_ = it;
  current_step_ = cs;
}

}  // namespace eco
