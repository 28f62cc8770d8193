// Host-side runtime shared by the three perception stages: device arena, named fp32 weight store,
// NHWC tensor views and the list of recorded kernel launches ("ops") that make up one forward pass.
// Nothing here is exported; the C-ABI lives in api.cu / include/peanut_b200.h.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <functional>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace pn {

enum DType : int { kBF16 = 0, kF32 = 1 };
inline size_t dtype_size(DType d) { return d == kBF16 ? 2 : 4; }

#define PN_CUDA_CHECK(expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      throw std::runtime_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                               ":" + std::to_string(__LINE__) + ")");                               \
  } while (0)

#define PN_REQUIRE(cond, msg)                                                        \
  do {                                                                               \
    if (!(cond)) throw std::runtime_error(std::string("peanut_b200: ") + (msg));     \
  } while (0)

// Kernels that split their linear thread index with 32-bit arithmetic (vec.cuh split_index): one thread per work item, < 2^32.
inline void check_u32_launch(long long threads, const char* what) {
  PN_REQUIRE(threads >= 0 && threads < (1ll << 32), std::string(what) + ": more than 2^32 work items in one launch");
}


// ---- programmatic dependent launch for every non-GEMM kernel of the launch lists.
// Each kernel starts with pdl_grid_sync(): it lets the NEXT kernel of the stream become resident right away and then
// blocks until the PREVIOUS kernel has completed and flushed, so launch latency and block scheduling overlap the
// predecessor's tail while every memory access stays ordered exactly as with plain stream serialization.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_grid_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

// Diagnosis switches (environment, read once): PN_DEBUG_NO_PDL=1 launches every kernel with plain stream serialization,
// PN_DEBUG_NO_LANES=1 puts every op on the caller's stream, PN_DEBUG_NO_GRAPH=1 replays the launch list eagerly,
// PN_DEBUG_SYNC_EACH=1 synchronises after every op of an eager pass with a watchdog that names the op that did not finish.
inline bool debug_flag(const char* name) {
  const char* e = std::getenv(name);
  return e && e[0] && e[0] != '0';
}
inline bool debug_no_pdl() {
  static const bool v = debug_flag("PN_DEBUG_NO_PDL");
  return v;
}

template <typename... P, typename... A>
inline void launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, A&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = debug_no_pdl() ? 0 : 1;
  static const bool only_in_graph = debug_flag("PN_DEBUG_PDL_ONLY_IN_GRAPH");  // diagnosis: plain launches outside stream capture
  if (only_in_graph) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    cudaStreamIsCapturing(s, &st);
    if (st != cudaStreamCaptureStatusActive) attr[0].val.programmaticStreamSerializationAllowed = 0;
  }
  cfg.attrs = attr, cfg.numAttrs = 1;
  PN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...));
}

// NHWC view: element (b, y, x, c) lives at ptr + ((b*H + y)*W + x)*ld + c.
struct Tensor {
  void* ptr = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  long long ld = 0;
  DType dt = kBF16;
  long long pixels() const { return static_cast<long long>(B) * H * W; }
  size_t bytes() const { return static_cast<size_t>(pixels()) * ld * dtype_size(dt); }
  // A view on channels [c0, c0 + c) of the same pixels.
  Tensor channels(int c0, int c) const {
    Tensor t = *this;
    t.ptr = static_cast<char*>(ptr) + static_cast<size_t>(c0) * dtype_size(dt);
    t.C = c;
    return t;
  }
};

struct HostArray {
  std::vector<float> data;
  std::vector<int64_t> shape;
  int64_t numel() const {
    int64_t n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// Owns every device allocation of one network instance.
class Arena {
 public:
  ~Arena() { release(); }
  void* alloc(size_t bytes, bool zero = true) {
    void* p = nullptr;
    bytes = (bytes + 255) & ~size_t(255);
    if (bytes == 0) bytes = 256;
    PN_CUDA_CHECK(cudaMalloc(&p, bytes));
    if (zero) PN_CUDA_CHECK(cudaMemset(p, 0, bytes));
    ptrs_.push_back(p);
    total_ += bytes;
    return p;
  }
  template <typename T>
  T* upload(const std::vector<T>& h) {
    T* d = static_cast<T*>(alloc(h.size() * sizeof(T), false));
    PN_CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    return d;
  }
  Tensor tensor(int B, int H, int W, int C, DType dt) {
    Tensor t;
    t.B = B, t.H = H, t.W = W, t.C = C, t.ld = C, t.dt = dt;
    t.ptr = alloc(t.bytes());
    return t;
  }
  void release() {
    for (void* p : ptrs_) cudaFree(p);
    ptrs_.clear();
    total_ = 0;
  }
  size_t total() const { return total_; }

 private:
  std::vector<void*> ptrs_;
  size_t total_ = 0;
};

using WeightStore = std::map<std::string, HostArray>;

// One recorded forward pass: a flat list of launches on a caller-supplied stream,
// optionally frozen into a CUDA graph after the first run.
struct Net {
  Arena arena;
  std::vector<std::function<void(cudaStream_t)>> ops;
  std::vector<std::string> op_names;
  std::vector<double> op_flops;                 // algorithmic FLOPs (2*MAC, unpadded) of each op; 0 for non-GEMM ops
  std::vector<std::pair<std::string, int>> stages;  // (name, index of the stage's first op), in order
  std::map<std::string, Tensor> taps;           // named NHWC activations (parity taps)
  std::map<std::string, std::pair<void*, size_t>> raw_taps;  // named raw device buffers (pointer, bytes)
  DType dt = kBF16;
  // fp32 parity mode (PN_FP32): fp32 storage WITHOUT the tf32 rounding of stored activations, every convolution evaluated as
  // three tf32 tensor-core products (hi*hi + hi*lo + lo*hi, conv_host.cu add_conv) - about fp32 accuracy at a third of the
  // tf32 path's speed.  Only meaningful with dt == kF32.
  bool x3 = false;
  bool round_stored() const { return dt == kF32 && !x3; }  // fp32 activations rounded to tf32 where they are stored
  int num_sms = 148;
  long long launches_per_forward = 0;
  int last_bn = 0;  // N tile chosen by the most recent add_conv (tuning aid)
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t cap_stream = nullptr;
  int warm_runs = 0;
  bool use_graph = true;

  // Parallel lanes: ops added while cur_lane != 0 are launched on side stream `cur_lane`.  A side-lane op is ordered
  // after every lane-0 op that precedes it in the list and after the earlier ops of its own lane; lane 0 (the caller's
  // stream) does not wait for side lanes until join_lanes() - placed before the first op that consumes their results -
  // or the end of the list.  run_range / profiling replay the list on one stream in list order, which satisfies the
  // same dependencies.
  std::vector<int> op_lane;
  std::vector<char> op_join;  // join every side lane into lane 0 before this op
  int cur_lane = 0;
  bool pending_join = false;
  static constexpr int kMaxLanes = 4;
  cudaStream_t lane_stream[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t lane_fork[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t lane_done[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  void set_lane(int lane) { cur_lane = lane; }
  void join_lanes() { pending_join = true, cur_lane = 0; }

  void add(const std::string& name, std::function<void(cudaStream_t)> f, double flops = 0.0) {
    op_lane.push_back(cur_lane);
    op_join.push_back(pending_join ? 1 : 0);
    pending_join = false;
    ops.push_back(std::move(f));
    op_names.push_back(name);
    op_flops.push_back(flops);
  }
  void stage(const std::string& name) { stages.emplace_back(name, static_cast<int>(ops.size())); }
  // Index of the first op of `name`; ops.size() for the pseudo-stage "end".
  int stage_begin(const std::string& name) const {
    if (name == "end") return static_cast<int>(ops.size());
    for (auto& st : stages)
      if (st.first == name) return st.second;
    throw std::runtime_error("peanut_b200: unknown stage '" + name + "'");
  }
  double total_flops() const {
    double t = 0;
    for (double f : op_flops) t += f;
    return t;
  }
  // Cross-stream ordering of successive forwards: the launch list owns one set of activations / per-call slots, so a
  // forward enqueued on stream B must not start (nor have its slots rewritten) while the previous one, enqueued on stream
  // A, is still in flight.  begin_forward(s) makes s wait for the previous forward when it ran on another stream;
  // end_forward(s) marks this one.  Same-stream callers pay nothing.
  cudaEvent_t fwd_done = nullptr;
  cudaStream_t fwd_stream = nullptr;
  bool fwd_pending = false;
  void begin_forward(cudaStream_t s) {
    if (fwd_pending && fwd_stream != s) PN_CUDA_CHECK(cudaStreamWaitEvent(s, fwd_done, 0));
  }
  void end_forward(cudaStream_t s) {
    static const bool off = debug_flag("PN_DEBUG_NO_FWD_ORDER");  // diagnosis only
    if (off) return;
    if (!fwd_done) PN_CUDA_CHECK(cudaEventCreateWithFlags(&fwd_done, cudaEventDisableTiming));
    PN_CUDA_CHECK(cudaEventRecord(fwd_done, s));
    fwd_stream = s, fwd_pending = true;
  }
  void run_eager(cudaStream_t s);
  void run_range(int first, int last, cudaStream_t s) {
    for (int i = first; i < last; ++i) ops[i](s);
  }
  void run(cudaStream_t s);
  ~Net();
};

inline const HostArray& get_weight(const WeightStore& w, const std::string& name) {
  auto it = w.find(name);
  PN_REQUIRE(it != w.end(), "missing weight '" + name + "'");
  return it->second;
}

// ---------------------------------------------------------------------------------------------
// Convolution (conv_host.cu)
struct ConvSpec {
  int Cin = 0, Cout = 0, R = 1, S = 1, stride = 1, dil = 1, pad = 0;
  int stride_w = -1, pad_w = -1;  // horizontal stride / padding when they differ from the vertical ones (-1 = same)
  bool relu = false;
  bool out_fp32 = false;
  int force_bn = 0;  // test hook: force the N tile
  int force_splits = 0;  // test hook: force the split-K factor
  int force_pair = 0;    // test hook: 1 = force a CTA pair (cta_group::2), 2 = forbid it
  bool no_split = false;  // never split K (layers whose row count is dynamic keep one CTA per tile)
  long long* dbg = nullptr;  // tuning aid: per-tile clock64 timeline of CTA 0
  int force_opt = 0;         // test hook: launch-shape options (1 two CTAs per SM, 2 bias block on multi-tile launches, 4 weights resident)
  int dbg_skip = 0;          // tuning aid: epilogue parts to skip (timeline builds only)
  bool force_direct_epilogue = false;  // test hook: bypass the TMA-staged epilogue
  // Dynamic row limit: when set, only the first (*m_limit) * m_limit_rows GEMM rows are computed (device-side
  // count of valid ROIs x rows per ROI); tiles beyond are skipped by every warp role.
  const int* m_limit = nullptr;
  int m_limit_rows = 1;
  int x3_cin = 0;  // internal: set on the inner call of the fp32 split-precision path (= the caller's Cin, for the FLOP count)
};
// weight: fp32 [Cout][Cin][R][S]; scale / bias: fp32 [Cout] (folded BN or plain bias with scale 1).
// `out` must already describe the destination view (B, Ho, Wo, C >= Cout rounded to 8, ld).
void add_conv(Net& net, const std::string& name, const Tensor& in, const Tensor& out, const float* weight,
              const float* scale, const float* bias, const ConvSpec& spec, const Tensor* residual = nullptr);
// Folds BatchNorm running statistics into (scale, bias): y = x*scale + bias, eps as in nn.BatchNorm2d.
void fold_bn(const WeightStore& w, const std::string& bn_prefix, int C, std::vector<float>& scale,
             std::vector<float>& bias, float eps = 1e-5f);
// api.cu: the thread-local message behind pn_last_error(), and the device / SM count of a pn_ctx (for sources that do not
// see the context's definition)
void set_last_error(const std::string& msg);
void ctx_device(const void* ctx, int* device, int* num_sms);

// Launch-configuration table of add_conv (conv_host.cu): "key bn splits pair opt" lines.  Entries that are present are used
// without timing, so every process that imports the same table builds bit-identical networks.
std::string conv_tuning_export();
int conv_tuning_import(const std::string& text);  // returns the number of entries read
void conv_tuning_clear();
inline int conv_out(int in, int k, int stride, int dil, int pad) {
  return (in + 2 * pad - dil * (k - 1) - 1) / stride + 1;
}
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
// Channel padding rule for activation tensors of dtype dt: multiple of 16 up to 64, then multiple of 64 (bf16);
// keeps every K block a whole TMA swizzle row.
int pad_channels(int c, DType dt);

// ---------------------------------------------------------------------------------------------
// Layout / pooling / resize kernels (ops.cu)
void add_nchw_to_nhwc(Net& net, const float* const* src_slot, const Tensor& out, int C);  // fp32 NCHW -> NHWC dt, zero pad
void add_maxpool3x3s2(Net& net, const Tensor& in, const Tensor& out);
void add_ppm_pool(Net& net, const Tensor& in, const std::vector<int>& scales, const std::vector<Tensor>& outs);
void add_bilinear_into(Net& net, const Tensor& in, const Tensor& out);  // align_corners=False, NHWC -> NHWC view
void add_upsample_logits(Net& net, const Tensor& logits, int C, int Hout, int Wout, float* const* dst_slot,
                         const int* sigmoid_slot);  // NHWC fp32 -> NCHW fp32 (+sigmoid)

}  // namespace pn
