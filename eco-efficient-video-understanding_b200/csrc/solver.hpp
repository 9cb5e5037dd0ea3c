// solver.hpp -- SGD / Nesterov solver over the device arenas of a TRAIN-phase eco::Net.
// Reference: caffe_3d/src/caffe/solver.cpp (Solver::Step :168-306, SGDSolver::GetLearningRate :580-620, ClipGradients :637-660,
// ApplyUpdate :662-674, Normalize / Regularize / ComputeUpdateValue :677-797, NesterovSolver :820-860, Snapshot / Restore
// :521-560) and its MPI gradient exchange (net.cpp:670-702, solver.cpp:310-347: SUM-allreduce of every parameter diff, then
// 1 / world).  Here the whole update is a handful of kernels over two contiguous fp32 arenas, and the exchange is a hook
// that hands contiguous gradient buckets to the caller (NCCL through torch.distributed) while backward is still running.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "net.hpp"

namespace eco {

void write_solverstate(const std::string& path, int iter, const std::string& learned_net, int current_step,
                       const std::vector<std::pair<std::vector<int>, std::vector<float>>>& history);
void read_solverstate(const std::string& path, int* iter, std::string* learned_net, int* current_step,
                      std::vector<std::vector<float>>* history);

// called once all gradient buckets are final (after backward, before the update): the callee all-reduces
// grad[offset, offset + count) for every bucket it was told about and returns when the work is ENQUEUED on `stream`
typedef void (*GradSyncFn)(void* user);

class Solver {
 public:
  Solver(const std::string& solver_text, const std::string& net_text_override, const std::string& base_dir);
  ~Solver();
  Net& net() { return *net_; }
  int iter() const { return iter_; }
  float step(int iters);            // Solver::Step: returns the loss of the last iteration (averaged over iter_size)
  void apply_update();              // SGDSolver::ApplyUpdate on whatever the gradient arena holds, ++iter
  float learning_rate() const;      // GetLearningRate at the current iteration
  void snapshot(const std::string& prefix_override);
  void restore(const std::string& state_file);
  void set_grad_sync(GradSyncFn fn, void* user, int world) { sync_fn_ = fn; sync_user_ = user; world_ = world < 1 ? 1 : world; }
  // parsed SolverParameter (caffe.proto:102-215)
  std::string type_ = "SGD", lr_policy_ = "fixed", snapshot_prefix_, net_path_;
  float base_lr_ = 0.01f, gamma_ = 0.1f, power_ = 1.f, momentum_ = 0.f, weight_decay_ = 0.f, clip_gradients_ = -1.f;
  int stepsize_ = 1, max_iter_ = 0, iter_size_ = 1, snapshot_ = 0, display_ = 0;
  std::vector<int> stepvalue_;

 private:
  std::unique_ptr<Net> net_;
  int iter_ = 0;
  mutable int current_step_ = 0;
  float* hist_ = nullptr;      // momentum history, same layout as the arenas
  float* scalars_ = nullptr;   // [0] sum of squares, [1] clip factor
  size_t hist_count_ = 0;
  GradSyncFn sync_fn_ = nullptr;
  void* sync_user_ = nullptr;
  int world_ = 1;
  void ensure_history();
};

}  // namespace eco
