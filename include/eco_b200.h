/*
 * eco_b200.h -- C ABI of libeco_b200.so: the drop-in boundary for ECO's hot path
 * (2-D BN-Inception trunk -> r2Dto3D -> 3-D ResNet-18 head -> global_pool -> fc) on B200.
 *
 * It replaces, for that path, the caffe_3d C++ surface
 *     caffe::Net<float>   caffe_3d/include/caffe/net.hpp:24-281   (src/caffe/net.cpp)
 *     caffe::Blob<float>  caffe_3d/include/caffe/blob.hpp:25-282  (src/caffe/blob.cpp, syncedmem.cpp)
 *     caffe::Caffe        caffe_3d/include/caffe/common.hpp:160-174 (set_mode / SetDevice)
 * and is what a binding (boost.python `_caffe.cpp`, our ctypes shim, a C++ facade) links to.
 * Plain pointers and sizes only; no torch / STL types cross this boundary.
 *
 * Conventions
 *   - Every call returns 0 on success, non-zero on failure; eco_last_error() then holds a
 *     message (thread-local).  caffe_3d aborts the process on CHECK failure
 *     (e.g. base_conv_layer.cpp:152, blob.hpp:141); a facade that wants that behaviour calls
 *     abort() on non-zero.
 *   - Host views are fp32 in caffe's logical layout (row-major N,C[,D],H,W), whatever the
 *     device layout is (bf16 channels-last).  They stay valid until eco_net_reshape /
 *     eco_net_destroy.
 *   - One eco_net per GPU; calls on one net must be serialised by the caller
 *     (caffe::Net is not thread-safe either).
 *   - There is no CPU execution path: eco_net_forward fails with an error if no CUDA
 *     device is usable (caffe's set_mode_cpu is accepted and ignored with an error on use).
 */
#ifndef ECO_B200_H_
#define ECO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct eco_net eco_net;

#define ECO_PHASE_TRAIN 0 /* caffe::TRAIN, caffe.proto Phase */
#define ECO_PHASE_TEST 1  /* caffe::TEST */

/* ---- process-wide (caffe::Caffe singleton, common.hpp:160-174; _caffe.cpp:213-215) ---- */
const char* eco_last_error(void);
const char* eco_version(void);
int eco_set_device(int device);          /* Caffe::SetDevice */
int eco_set_mode(int gpu);               /* Caffe::set_mode: 1 = GPU; 0 = CPU is recorded, forward then fails */
int eco_device_count(int* count);        /* 0 devices is not an error here */

/* ---- construction (Net::Net(file, phase) net.cpp:31-36; Net::Init :39-316) ---- */
int eco_net_create(const char* prototxt_path, int phase, eco_net** out);
int eco_net_create_from_string(const char* prototxt_text, int phase, eco_net** out);
int eco_net_destroy(eco_net* net);
/* options, set before the first forward/reshape (full list: INTEGRATION.md section 4):  "keep_all_blobs" (0/1: also store blobs the
 * fused plan would keep on chip, e.g. a conv output that only feeds its BN), "a_mode"
 * (0 cp.async gather, 1 TMA im2col, -1 auto), "use_graph" (0/1 CUDA-graph replay of full forwards) */
int eco_net_set_option(eco_net* net, const char* key, int value);
/* run on a caller-owned CUDA stream (cudaStream_t as void*), default: the net's own stream */
int eco_net_set_stream(eco_net* net, void* cuda_stream);

/* ---- weights (Net::CopyTrainedLayersFrom net.cpp:852-883, ToProto :885-904; blobs() of Layer) ---- */
int eco_net_copy_from(eco_net* net, const char* caffemodel_path);   /* match by layer name */
int eco_net_save(const eco_net* net, const char* caffemodel_path);
int eco_net_layer_num_params(const eco_net* net, int layer, int* n);
int eco_net_param_shape(const eco_net* net, int layer, int blob_idx, int* dims, int* ndims /* in: capacity, out: used */);
int eco_net_set_param(eco_net* net, int layer, int blob_idx, const float* data, size_t count);
int eco_net_get_param(const eco_net* net, int layer, int blob_idx, float* data, size_t count);
/* mutable host pointer to a parameter blob (caffe: layer->blobs()[i]->mutable_cpu_data()); the
 * device copy is refreshed at the next forward */
int eco_net_param_host(eco_net* net, int layer, int blob_idx, float** data, size_t* count);

/* ---- introspection (net.hpp:101-195: name(), layer_names(), blob_names(), inputs/outputs, has_blob ...) ---- */
const char* eco_net_name(const eco_net* net);
int eco_net_phase(const eco_net* net);
int eco_net_num_layers(const eco_net* net);            /* includes auto-inserted Split layers (insert_splits.cpp) */
const char* eco_net_layer_name(const eco_net* net, int i);
const char* eco_net_layer_type(const eco_net* net, int i);
int eco_net_layer_index(const eco_net* net, const char* name);   /* -1 if absent */
int eco_net_layer_num_bottoms(const eco_net* net, int i);
int eco_net_layer_bottom(const eco_net* net, int i, int j);      /* blob index */
int eco_net_layer_num_tops(const eco_net* net, int i);
int eco_net_layer_top(const eco_net* net, int i, int j);
int eco_net_num_blobs(const eco_net* net);
const char* eco_net_blob_name(const eco_net* net, int i);
int eco_net_blob_index(const eco_net* net, const char* name);    /* -1 if absent */
int eco_net_blob_shape(const eco_net* net, int i, int* dims, int* ndims /* in: capacity, out: used */);
int eco_net_num_inputs(const eco_net* net);
int eco_net_input_blob(const eco_net* net, int i);
int eco_net_num_outputs(const eco_net* net);
int eco_net_output_blob(const eco_net* net, int i);

/* ---- shapes (Blob::Reshape blob.cpp:22-51, Net::Reshape net.cpp:824-828) ---- */
int eco_blob_reshape(eco_net* net, int blob, const int* dims, int ndims);   /* then eco_net_reshape */
int eco_net_reshape(eco_net* net);

/* ---- execution (Net::ForwardFromTo net.cpp:566-583, BackwardFromTo :637-706) ---- */
int eco_net_forward(eco_net* net, int start, int end, float* loss);  /* layer indices as in caffe; end = -1: last */
/* Backward over visible layers start..end (start >= end, -1/-1 = whole net), TRAIN-phase nets only: fills the blob diffs
 * and ACCUMULATES the parameter diffs (caffe's beta = 1, conv_layer.cpp:44-75); clear them with eco_net_clear_param_diffs
 * (Solver::Step does that once per iteration, solver.cpp:178-195). */
int eco_net_backward(eco_net* net, int start, int end);
int eco_net_clear_param_diffs(eco_net* net);
int eco_net_update(eco_net* net);   /* Net::Update (net.cpp:906-910): every parameter blob data -= diff, on the device arenas */
/* fp32 host copy of a parameter gradient (layer->blobs()[i]->cpu_diff()), synced from the device */
int eco_net_param_diff_host(eco_net* net, int layer, int blob_idx, float** data, size_t* count);

/* ---- device-resident training state (what Solver / the gradient all-reduce work on; net.cpp:670-702, solver.cpp:310-347) ----
 * All parameter blobs of a TRAIN-phase net live in ONE fp32 device arena in layer order (caffe layout per blob), the
 * gradients in a second arena of the same layout: a solver updates, and NCCL all-reduces, contiguous ranges of them. */
typedef struct eco_param_slot {
  int layer, blob;          /* visible layer index, blob index inside the layer */
  size_t offset, count;     /* floats from the arena base */
  float lr_mult, decay_mult; /* ParamSpec; BN running statistics are forced to 0 (bn_layer.cpp:46-53) */
} eco_param_slot;
int eco_net_param_arena(eco_net* net, float** dev, size_t* count);
int eco_net_grad_arena(eco_net* net, float** dev, size_t* count);
int eco_net_num_param_slots(eco_net* net, int* n);
int eco_net_param_slot(eco_net* net, int i, eco_param_slot* out);
/* tell the net that the parameter arena was changed on the device (solver step): GEMM operands are re-packed at the next
 * forward and the host views are refreshed on access */
int eco_net_params_updated_on_device(eco_net* net);
int eco_net_cuda_stream(eco_net* net, void** cuda_stream);   /* the stream the net's kernels run on */
/* Gradient exchange hook (replaces the per-blob host-staged MPI_Allreduce of net.cpp:670-702): the gradient arena is cut
 * into `nbuckets` contiguous ranges in layer order; during eco_net_backward `fn(user, bucket, offset, count)` is called on
 * the host as soon as all kernels that produce grad[offset, offset+count) are ENQUEUED on the net's stream -- the caller
 * records an event there and starts ncclAllReduce of that range on a side stream, overlapping the rest of backward. */
typedef void (*eco_grad_bucket_fn)(void* user, int bucket, size_t offset, size_t count);
int eco_net_set_grad_bucket_hook(eco_net* net, int nbuckets, eco_grad_bucket_fn fn, void* user);
int eco_net_num_grad_buckets(eco_net* net, int* n);
int eco_net_grad_bucket(eco_net* net, int i, size_t* offset, size_t* count);   /* bucket 0 = the last layers */

/* ---- Solver (caffe::Solver / SGDSolver / NesterovSolver, solver.hpp; _caffe.cpp:296-317): SGD and Nesterov momentum
 * with lr_mult / decay_mult per blob, L2 decay, global-norm clipping, every lr_policy of solver.cpp:580-620, iter_size
 * accumulation, Snapshot / Restore (.caffemodel + .solverstate, caffe.proto:217-222).  The train net's inputs (data, label)
 * are filled by the caller through the net handle before each step (the VideoData layer is SURVEY 8(f1)). ---- */
typedef struct eco_solver eco_solver;
int eco_solver_create(const char* solver_prototxt_path, eco_solver** out);    /* `net:` is resolved next to the solver file */
int eco_solver_create_from_string(const char* solver_text, const char* net_text /* NULL: read `net:` */, eco_solver** out);
int eco_solver_destroy(eco_solver* s);
int eco_solver_net(eco_solver* s, eco_net** net);          /* borrowed handle, valid until eco_solver_destroy */
int eco_solver_iter(eco_solver* s, int* iter);
int eco_solver_learning_rate(eco_solver* s, float* rate);
int eco_solver_step(eco_solver* s, int iters, float* last_loss);   /* Solver::Step */
int eco_solver_apply_update(eco_solver* s);                /* SGDSolver::ApplyUpdate on the current diffs, ++iter */
/* called between backward and the update of every iteration; `world` = number of data-parallel replicas whose gradients
 * the callee sums: the update then uses diff / world (solver.cpp:332-337) */
typedef void (*eco_grad_sync_fn)(void* user);
int eco_solver_set_grad_sync(eco_solver* s, eco_grad_sync_fn fn, void* user, int world);
int eco_solver_snapshot(eco_solver* s, const char* prefix /* NULL/"": snapshot_prefix */);
int eco_solver_restore(eco_solver* s, const char* solverstate_path);
int eco_net_sync(eco_net* net);                                      /* wait for the net's stream */

/* ---- blob data (Blob::cpu_data / mutable_cpu_data / cpu_diff, SyncedMemory syncedmem.cpp:21-70) ---- */
/* fp32 host mirror in caffe layout; `for_write` = 1 marks the host copy newer (mutable_cpu_data) so
 * the next forward uploads it; 0 only syncs device -> host if the device copy is newer. */
int eco_blob_host_data(eco_net* net, int blob, int for_write, float** data, size_t* count);
int eco_blob_host_diff(eco_net* net, int blob, int for_write, float** data, size_t* count);

/* ---- fast paths beyond caffe's surface (device-resident I/O for serving) ---- */
/* copy `count` fp32 values already on the device (caffe layout) into an input blob, no host hop.  The device copy
 * stays the blob's content for every following forward until the caller asks for the blob's host memory again
 * (eco_blob_host_data downloads it first and then makes the host mirror the source, as for any caffe blob). */
int eco_net_set_input_device(eco_net* net, int blob, const void* dev_f32, size_t count);
/* device pointer of a plain fp32 blob (e.g. fc8) valid until reshape; NULL/err for fused-away blobs */
int eco_blob_device_f32(eco_net* net, int blob, const float** dev, size_t* count);

/* pipelined serving: enqueue one forward whose input comes from `host_in` (fp32, caffe layout; use pinned
 * memory, e.g. eco_host_alloc, for a truly asynchronous copy) and whose first output blob is copied to
 * `host_out`; returns at once with a ticket.  The copy of call k+1 overlaps the compute of call k; the
 * caller must not touch host_in/host_out of a call until eco_net_wait(ticket) returned, and alternates
 * two buffer pairs. */
int eco_net_forward_pipelined(eco_net* net, const float* host_in, size_t count, float* host_out, size_t out_count,
                              int* ticket);
/* the same fed from raw uint8 frames (caffe Datum layout [F,3,H,W]); `mean` (per channel, BGR order as in
 * transform_param.mean_value) is subtracted on the device -- the part of DataTransformer::Transform
 * (data_transformer.cpp) that the reference runs on the host before the data blob exists.  4x fewer PCIe bytes. */
int eco_net_forward_pipelined_u8(eco_net* net, const unsigned char* host_in, size_t count, const float* mean, int nmean,
                                 float* host_out, size_t out_count, int* ticket);
int eco_net_wait(eco_net* net, int ticket);

/* ---- online sliding-window recognition (scripts/online_recognition/online_recognition.py:64-93) ----
 * The reference recomputes the whole net on all N frames of the window for every new frame.  Frames are independent through
 * the 2-D trunk, so a trunk-only net (the definition cut after `until_blob`, e.g. inception_3c_double_3x3_1_bn: 96x28x28 per
 * frame) runs on the NEW frames only, their features are appended to the window held in the full net's own feature blob
 * (older frames shift towards 0), and eco_net_forward(full, first_head_layer, -1) runs the 3-D head once. */
int eco_net_create_from_string_until(const char* prototxt_text, int phase, const char* until_blob, eco_net** out);
int eco_net_push_frames(eco_net* dst, int dst_blob, eco_net* src, int src_blob);
int eco_host_alloc(void** ptr, size_t bytes);   /* page-locked host memory */
int eco_host_free(void* ptr);

/* ---- the step in front of the path: VideoDataLayer's sampling and DataTransformer::Transform ----
 * video_data_layer.cpp:134-238 picks N frames per video (TRAIN: a random offset inside each of N equal segments, TEST: the
 * centres), stacks them as a Datum [3N, H, W] and DataTransformer::Transform (data_transformer.cpp:148-326) crops (multi-scale
 * sizes x fixed offsets), resizes to crop_size with cv::resize, mirrors, subtracts the mean.  Here the host draws the same
 * choices (std::mt19937, `rng() % n` like caffe's rng_t) and the GPU does the pixel work straight into the net's input blob.
 * JPEG decoding is not part of this library: the caller hands over decoded uint8 frames. */
typedef struct eco_clip_transform { int h_off, w_off, crop_h, crop_w, mirror; } eco_clip_transform;
typedef struct eco_transform_param {   /* TransformationParameter, caffe.proto */
  int mirror, multi_scale, fix_crop, more_fix_crop, max_distort, is_flow;
  int num_scale_ratios;
  float scale_ratios[8];
  float scale;
  int num_mean;
  float mean_value[16];
} eco_transform_param;
typedef struct eco_sampler eco_sampler;   /* holds the mt19937 streams (frame sampling, transform choices) */
int eco_sampler_create(unsigned int seed, eco_sampler** out);
int eco_sampler_destroy(eco_sampler* s);
int eco_sample_segment_offsets(eco_sampler* s, int num_frames, int num_segments, int new_length, int train, int* offsets /* [num_segments] */);
int eco_sample_clip_transform(eco_sampler* s, int H, int W, int crop_size, int train, const eco_transform_param* p, eco_clip_transform* out);
int eco_crop_size_candidates(int H, int W, int crop_size, int max_distort, const float* ratios, int nratios, int* hw_pairs, int* n /* in: capacity in pairs */);
int eco_fix_offset_candidates(int H, int W, int crop_h, int crop_w, int more, int* hw_pairs, int* n);
/* transform B clips (uint8 Datum layout [B][C][H][W], host memory) into input blob `blob` of the net (fp32 [B][C][crop][crop]
 * on the device, no host round trip of the floats); `t` = one transform per clip */
int eco_net_transform_input_u8(eco_net* net, int blob, const unsigned char* src, int B, int C, int H, int W,
                               const eco_clip_transform* t, const eco_transform_param* p);

/* ---- measurement hooks used by bench.py ---- */
/* number of kernels this library launched for the last eco_net_forward on this net */
int eco_net_last_launch_count(const eco_net* net, int* launches);
/* per-op timing of one forward with CUDA events on the net's stream: fills up to `cap` entries;
 * names are owned by the net.  kind: 0 conv(implicit GEMM), 1 other kernel, 2 memcpy */
typedef struct eco_op_time {
  const char* name;
  int kind;
  float ms;
  double flops;       /* 2*M*N*K incl. padding taps, the algorithmic figure of SURVEY.md 8(d) */
  double bytes;       /* algorithmic HBM bytes: input once + output once + weights once */
} eco_op_time;
int eco_net_profile_forward(eco_net* net, eco_op_time* out, int cap, int* n);
/* TRAIN-phase nets: one forward + backward pass with an event after every forward op and every backward sub-step
 * ("fwd:<op>", "bwd:<op>:wgrad", "bwd:<op>:dgrad", "bwd:<op>:bias", "bwd:<op>"); kind 0 = tcgen05 GEMM step */
int eco_net_profile_train(eco_net* net, eco_op_time* out, int cap, int* n);
/* one text line per planned op: name, kernel, tile shape (block_n, MT, pair), grid, stages -- what the planner chose for
 * the current shapes (plans on first use).  `needed` receives the full length incl. the terminating 0. */
int eco_net_describe_plan(eco_net* net, char* buf, size_t cap, size_t* needed);

#ifdef __cplusplus
}
#endif
#endif /* ECO_B200_H_ */
