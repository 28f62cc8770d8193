// C-ABI of libpeanut_b200.so (declared in include/peanut_b200.h).  Every entry point translates
// C++ exceptions into a status code + thread-local message; no torch types cross this boundary.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>

#include "../../include/peanut_b200.h"
#include "engine.h"
#include "maskrcnn.h"
#include "prednet.h"
#include "semmap.h"
#include "vec.cuh"

namespace pn {

struct Ctx {
  int device = 0;
  int num_sms = 148;
  WeightStore weights;
  std::unique_ptr<PredNet> prednet;
  std::unique_ptr<SemMap> semmap;
  std::unique_ptr<MaskRcnn> maskrcnn;
  cudaStream_t stream = nullptr;  // used by the *_host entry points
};

thread_local std::string g_last_error;

__global__ void set_pred_slots_kernel(PredSlots* slots, const float* in, float* out, int apply_sigmoid) {
  pdl_grid_sync();
  slots->input = in;
  slots->output = out;
  slots->apply_sigmoid = apply_sigmoid;
}

// NHWC (dt) -> NCHW fp32, first C channels.
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* in, long long ldi, float* out, int B, int C, int HW) {
  pdl_grid_sync();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * C * HW;
  if (idx >= total) return;
  const int hw = static_cast<int>(idx % HW);
  const long long t = idx / HW;
  const int c = static_cast<int>(t % C);
  const int b = static_cast<int>(t / C);
  out[idx] = to_float(in[(static_cast<long long>(b) * HW + hw) * ldi + c]);
}

void nhwc_to_nchw(const Tensor& t, int C, float* out, cudaStream_t s) {
  const int HW = t.H * t.W;
  const long long total = static_cast<long long>(t.B) * C * HW;
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads);
  if (t.dt == kBF16)
    launch_pdl(nhwc_to_nchw_kernel<__nv_bfloat16>, blocks, threads, 0, s, static_cast<const __nv_bfloat16*>(t.ptr), t.ld, out, t.B, C, HW);
  else
    launch_pdl(nhwc_to_nchw_kernel<float>, blocks, threads, 0, s, static_cast<const float*>(t.ptr), t.ld, out, t.B, C, HW);
}

// NCHW fp32 -> NHWC (dt) first C channels of an existing tensor (tap injection).
template <typename T>
__global__ void nchw_to_nhwc_direct_kernel(const float* in, T* out, long long ldo, int B, int C, int HW, int round_tf32) {
  pdl_grid_sync();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  const long long total = static_cast<long long>(B) * C * HW;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % C);
  const long long t = idx / C;
  const int hw = static_cast<int>(t % HW);
  const int b = static_cast<int>(t / HW);
  float v = in[(static_cast<long long>(b) * C + c) * HW + hw];
  if (round_tf32) {
    uint32_t q;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q) : "f"(v));
    v = __uint_as_float(q);
  }
  out[(static_cast<long long>(b) * HW + hw) * ldo + c] = from_float<T>(v);
}

void nchw_to_nhwc_direct(const float* in, const Tensor& t, int C, cudaStream_t s, bool round_stored = true) {
  const int HW = t.H * t.W;
  const long long total = static_cast<long long>(t.B) * C * HW;
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads);
  if (t.dt == kBF16)
    launch_pdl(nchw_to_nhwc_direct_kernel<__nv_bfloat16>, blocks, threads, 0, s, in, static_cast<__nv_bfloat16*>(t.ptr), t.ld, t.B, C, HW, 0);
  else
    launch_pdl(nchw_to_nhwc_direct_kernel<float>, blocks, threads, 0, s, in, static_cast<float*>(t.ptr), t.ld, t.B, C, HW, round_stored ? 1 : 0);
}

void set_last_error(const std::string& msg) { g_last_error = msg; }
void ctx_device(const void* ctx, int* device, int* num_sms) {
  const Ctx* c = static_cast<const Ctx*>(ctx);
  PN_REQUIRE(c != nullptr, "null context");
  if (device) *device = c->device;
  if (num_sms) *num_sms = c->num_sms;
}

void check_device(Ctx* c) {
  PN_REQUIRE(c != nullptr, "null context");
  PN_CUDA_CHECK(cudaSetDevice(c->device));
}

}  // namespace pn

using namespace pn;

#define PN_API_BEGIN try {
#define PN_API_END                        \
  }                                       \
  catch (const std::exception& e) {       \
    g_last_error = e.what();              \
    return 1;                             \
  }                                       \
  catch (...) {                           \
    g_last_error = "unknown exception";   \
    return 1;                             \
  }                                       \
  return 0;

extern "C" {

const char* pn_last_error(void) { return g_last_error.c_str(); }
int pn_abi_version(void) { return 3; }

int pn_create(int device, pn_ctx** out) {
  PN_API_BEGIN
  PN_REQUIRE(out != nullptr, "pn_create: null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  PN_REQUIRE(e == cudaSuccess && count > 0, "no CUDA device available: peanut_b200 has no CPU fallback");
  PN_REQUIRE(device >= 0 && device < count, "device index out of range");
  cudaDeviceProp prop;
  PN_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  PN_REQUIRE(prop.major == 10, "peanut_b200 kernels are built for sm_100a only (found sm_" +
                                   std::to_string(prop.major) + std::to_string(prop.minor) + ")");
  PN_CUDA_CHECK(cudaSetDevice(device));
  auto* c = new Ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  PN_CUDA_CHECK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  *out = reinterpret_cast<pn_ctx*>(c);
  PN_API_END
}

int pn_destroy(pn_ctx* ctx) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  if (c) {
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    c->prednet.reset();
    c->semmap.reset();
    c->maskrcnn.reset();
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
  }
  PN_API_END
}

int pn_set_weight(pn_ctx* ctx, const char* name, const float* data, int ndim, const int64_t* shape) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  PN_REQUIRE(c && name && data && ndim >= 0 && ndim <= 8, "pn_set_weight: bad arguments");
  HostArray a;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    a.shape.push_back(shape[i]);
    n *= shape[i];
  }
  a.data.assign(data, data + n);
  c->weights[name] = std::move(a);
  PN_API_END
}

int pn_clear_weights(pn_ctx* ctx) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  PN_REQUIRE(c, "null context");
  c->weights.clear();
  PN_API_END
}

int pn_prednet_build(pn_ctx* ctx, int B, int C, int H, int W, int num_classes, int precision) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(B > 0 && C > 0 && H >= 8 && W >= 8 && num_classes > 0, "pn_prednet_build: bad shape");
  PN_REQUIRE(precision == PN_BF16 || precision == PN_TF32 || precision == PN_FP32, "pn_prednet_build: bad precision");
  c->prednet.reset();
  auto net = std::make_unique<PredNet>();
  net->net.num_sms = c->num_sms;
  net->net.x3 = precision == PN_FP32;
  build_prednet(*net, c->weights, B, C, H, W, num_classes, precision == PN_BF16 ? kBF16 : kF32);
  PN_CUDA_CHECK(cudaDeviceSynchronize());
  c->prednet = std::move(net);
  PN_API_END
}

int pn_prednet_forward(pn_ctx* ctx, const float* map_dev, int apply_sigmoid, float* out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->prednet, "pn_prednet_forward: call pn_prednet_build first");
  PN_REQUIRE(map_dev && out_dev, "pn_prednet_forward: null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  c->prednet->net.begin_forward(s);
  launch_pdl(set_pred_slots_kernel, 1, 1, 0, s, c->prednet->slots, map_dev, out_dev, apply_sigmoid);
  c->prednet->net.run(s);
  c->prednet->net.end_forward(s);
  PN_CUDA_CHECK(cudaGetLastError());
  PN_API_END
}

int pn_prednet_forward_host(pn_ctx* ctx, const float* map_host, int apply_sigmoid, float* out_host) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->prednet, "pn_prednet_forward_host: call pn_prednet_build first");
  PredNet& n = *c->prednet;
  const size_t in_bytes = static_cast<size_t>(n.B) * n.C * n.H * n.W * sizeof(float);
  const size_t out_bytes = static_cast<size_t>(n.B) * n.num_classes * n.H * n.W * sizeof(float);
  if (!n.stage_in) {
    n.stage_in = static_cast<float*>(n.net.arena.alloc(in_bytes, false));
    n.stage_out = static_cast<float*>(n.net.arena.alloc(out_bytes, false));
  }
  n.net.begin_forward(c->stream);
  PN_CUDA_CHECK(cudaMemcpyAsync(n.stage_in, map_host, in_bytes, cudaMemcpyHostToDevice, c->stream));
  launch_pdl(set_pred_slots_kernel, 1, 1, 0, c->stream, n.slots, n.stage_in, n.stage_out, apply_sigmoid);
  n.net.run(c->stream);
  n.net.end_forward(c->stream);
  PN_CUDA_CHECK(cudaMemcpyAsync(out_host, n.stage_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  PN_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PN_API_END
}

int pn_prednet_flops(pn_ctx* ctx, double* flops_out) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  PN_REQUIRE(c && c->prednet && flops_out, "pn_prednet_flops: not built");
  *flops_out = c->prednet->net.total_flops();
  PN_API_END
}

int pn_prednet_num_launches(pn_ctx* ctx) {
  auto* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !c->prednet) return -1;
  return static_cast<int>(c->prednet->net.launches_per_forward) + 1;  // + slot update
}

int pn_prednet_read_tap(pn_ctx* ctx, int which, float* out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->prednet && out_dev, "pn_prednet_read_tap: not built");
  PredNet& n = *c->prednet;
  if (which == 0)
    nhwc_to_nchw(n.features, 2048, out_dev, static_cast<cudaStream_t>(stream));
  else if (which == 1)
    nhwc_to_nchw(n.logits_lowres, n.num_classes, out_dev, static_cast<cudaStream_t>(stream));
  else
    PN_REQUIRE(false, "pn_prednet_read_tap: unknown tap");
  PN_CUDA_CHECK(cudaGetLastError());
  PN_API_END
}

int pn_semmap_build(pn_ctx* ctx, int num_envs, const pn_semmap_cfg* cfg) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(cfg != nullptr && num_envs > 0, "pn_semmap_build: bad arguments");
  PN_REQUIRE(cfg->du_scale == 1, "pn_semmap_build: du_scale != 1 is not supported (reference default is 1)");
  PN_REQUIRE(cfg->map_resolution > 0 && cfg->vision_range > 0 && cfg->global_downscaling > 0, "pn_semmap_build: bad geometry");
  // Constants exactly as Semantic_Mapping.__init__ / forward compute them (mapping.py:12-50, 66-72, 102-103, 127-130)
  SemMapCfg g{};
  g.h = cfg->frame_height, g.w = cfg->frame_width;
  g.channels = 4 + cfg->num_sem_categories;
  g.nf = 1 + cfg->num_sem_categories;
  g.ego_channels = 2 + cfg->num_sem_categories;
  g.vr = cfg->vision_range;
  const int res = cfg->map_resolution;
  const int max_h = static_cast<int>(360.0 / res), min_h = static_cast<int>(-40.0 / res);
  g.nz = max_h - min_h;
  const double agent_height = cfg->camera_height * 100.0;  // Python float arithmetic of mapping.py:33
  g.min_z = static_cast<int>(25.0 / res - min_h);
  g.max_z = static_cast<int>((agent_height + 1) / res - min_h);
  g.map_cells = (cfg->map_size_cm / cfg->global_downscaling) / res;
  if (cfg->num_sem_categories <= 16) {
    g.special_f[0] = 1 + 5, g.special_f[1] = 1 + 2, g.special_f[2] = -1;
  } else {
    g.special_f[0] = 1 + 3, g.special_f[1] = 1 + 9, g.special_f[2] = 1 + 14;
  }
  g.xc = static_cast<float>((g.w - 1.0) / 2.0);
  g.zc = static_cast<float>((g.h - 1.0) / 2.0);
  // (w / 2.) / np.tan(np.deg2rad(fov / 2.)), depth_utils.py get_camera_matrix; deg2rad(x) = x * (pi / 180)
  g.f = static_cast<float>((g.w / 2.0) / std::tan((cfg->hfov / 2.0) * (3.14159265358979323846 / 180.0)));
  g.agent_height = static_cast<float>(agent_height);
  g.shift_x = static_cast<float>(g.vr * res / 2);
  g.res = static_cast<float>(res);
  g.half_vr = static_cast<float>(g.vr / 2);
  g.vr_f = static_cast<float>(g.vr);
  g.z_mid = static_cast<float>(std::floor((max_h + min_h) / 2.0));
  g.nz_f = static_cast<float>(g.nz);
  g.map_thr = cfg->map_pred_threshold, g.exp_thr = cfg->exp_pred_threshold, g.cat_thr = cfg->cat_pred_threshold;
  PN_REQUIRE(g.map_cells >= g.vr && g.map_cells % 2 == 0, "pn_semmap_build: local map smaller than the vision range");
  auto sm = std::make_unique<SemMap>();
  sm->init(g, num_envs);
  PN_CUDA_CHECK(cudaDeviceSynchronize());
  c->semmap = std::move(sm);
  PN_API_END
}

int pn_semmap_forward(pn_ctx* ctx, const float* obs_dev, const float* pose_delta_dev, const float* maps_last_dev,
                      const int64_t* maps_last_strides, float* poses_inout_dev, float* fp_map_out_dev,
                      float* map_out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->semmap, "pn_semmap_forward: call pn_semmap_build first");
  PN_REQUIRE(obs_dev && pose_delta_dev && maps_last_dev && poses_inout_dev && map_out_dev, "pn_semmap_forward: null buffer");
  PN_REQUIRE(maps_last_dev != map_out_dev, "pn_semmap_forward: map_out must not alias maps_last");
  const SemMapCfg& g = c->semmap->c;
  const long long plane = static_cast<long long>(g.map_cells) * g.map_cells;
  long long se = plane * g.channels, sp = plane, sr = g.map_cells;
  if (maps_last_strides) se = maps_last_strides[0], sp = maps_last_strides[1], sr = maps_last_strides[2];
  c->semmap->forward(obs_dev, pose_delta_dev, maps_last_dev, se, sp, sr, poses_inout_dev, fp_map_out_dev, map_out_dev,
                     static_cast<cudaStream_t>(stream));
  PN_API_END
}

int pn_semmap_read_ego(pn_ctx* ctx, float* ego_out_dev, int* stair_flags_out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->semmap, "pn_semmap_read_ego: not built");
  SemMap& m = *c->semmap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (ego_out_dev)
    PN_CUDA_CHECK(cudaMemcpyAsync(ego_out_dev, m.ego, static_cast<size_t>(m.E) * m.c.ego_channels * m.c.vr * m.c.vr * sizeof(float), cudaMemcpyDeviceToDevice, s));
  if (stair_flags_out_dev)
    PN_CUDA_CHECK(cudaMemcpyAsync(stair_flags_out_dev, m.stair_flag, static_cast<size_t>(m.E) * sizeof(int), cudaMemcpyDeviceToDevice, s));
  PN_API_END
}

int pn_semmap_num_launches(pn_ctx* ctx) {
  auto* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !c->semmap) return -1;
  return SemMap::kLaunches;
}

static Net* select_net(Ctx* c, int which) {
  if (which == PN_NET_PREDNET) return c->prednet ? &c->prednet->net : nullptr;
  if (which == PN_NET_MASKRCNN) return c->maskrcnn ? &c->maskrcnn->net : nullptr;
  return nullptr;
}

int pn_net_num_ops(pn_ctx* ctx, int which) {
  auto* c = reinterpret_cast<Ctx*>(ctx);
  Net* net = c ? select_net(c, which) : nullptr;
  return net ? static_cast<int>(net->ops.size()) : -1;
}

int pn_net_profile(pn_ctx* ctx, int which, int iters, float* ms_out, double* flops_out, int max_ops, char* names_out,
                   int names_bytes) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  Net* netp = select_net(c, which);
  PN_REQUIRE(netp && ms_out && iters > 0, "pn_net_profile: network not built");
  Net& net = *netp;
  const int n = static_cast<int>(net.ops.size());
  PN_REQUIRE(max_ops >= n, "pn_net_profile: ms_out too small");
  // the per-call slots must already point at valid buffers: run one regular forward before profiling
  std::vector<cudaEvent_t> ev(n + 1);
  for (auto& e : ev) PN_CUDA_CHECK(cudaEventCreate(&e));
  std::vector<double> acc(n, 0.0);
  for (int it = 0; it < iters + 1; ++it) {
    PN_CUDA_CHECK(cudaEventRecord(ev[0], c->stream));
    for (int i = 0; i < n; ++i) {
      net.ops[i](c->stream);
      PN_CUDA_CHECK(cudaEventRecord(ev[i + 1], c->stream));
    }
    PN_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    if (it == 0) continue;  // warm-up
    for (int i = 0; i < n; ++i) {
      float ms = 0.f;
      PN_CUDA_CHECK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      acc[i] += ms;
    }
  }
  for (auto& e : ev) cudaEventDestroy(e);
  std::string names;
  for (int i = 0; i < n; ++i) {
    ms_out[i] = static_cast<float>(acc[i] / iters);
    if (flops_out) flops_out[i] = net.op_flops[i];
    names += net.op_names[i];
    names += '\n';
  }
  if (names_out && names_bytes > 0) {
    const size_t k = std::min(names.size(), static_cast<size_t>(names_bytes - 1));
    memcpy(names_out, names.data(), k);
    names_out[k] = 0;
  }
  PN_API_END
}

// ------------------------------------------------------------------------------------------------- stage A
int pn_maskrcnn_build(pn_ctx* ctx, int B, int H, int W, int precision, const pn_maskrcnn_cfg* cfg) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(B > 0 && H > 0 && W > 0, "pn_maskrcnn_build: bad shape");
  PN_REQUIRE(precision == PN_BF16 || precision == PN_TF32 || precision == PN_FP32, "pn_maskrcnn_build: bad precision");
  MrcnnCfg g;
  g.B = B, g.H = H, g.W = W;
  if (cfg) {
    g.min_size = cfg->min_size_test, g.max_size = cfg->max_size_test;
    g.pre_nms_topk = cfg->rpn_pre_nms_topk, g.post_nms_topk = cfg->rpn_post_nms_topk, g.rpn_nms = cfg->rpn_nms_thresh;
    g.num_classes = cfg->num_classes, g.box_nms = cfg->box_nms_thresh, g.detections = cfg->detections_per_image;
    g.mask_thresh = cfg->mask_threshold;
  }
  c->maskrcnn.reset();
  auto net = std::make_unique<MaskRcnn>();
  net->net.num_sms = c->num_sms;
  net->net.x3 = precision == PN_FP32;
  build_maskrcnn(*net, c->weights, g, precision == PN_BF16 ? kBF16 : kF32);
  PN_CUDA_CHECK(cudaDeviceSynchronize());
  c->maskrcnn = std::move(net);
  PN_API_END
}

static void set_mrcnn_slots(MaskRcnn& m, const uint8_t* rgb, const int* goal, float score_thresh, float sem_thr, float goal_thr,
                            float* sem_out, cudaStream_t s) {
  MrcnnSlots h{rgb, sem_out, goal, score_thresh, sem_thr, goal_thr};
  // small pageable copy: the driver stages it at call time, so `h` may die when we return
  PN_CUDA_CHECK(cudaMemcpyAsync(m.slots, &h, sizeof(h), cudaMemcpyHostToDevice, s));
}

int pn_maskrcnn_forward(pn_ctx* ctx, const uint8_t* rgb_dev, const int* goal_cat_dev, float score_thresh,
                        float sem_pred_prob_thr, float goal_thr, float* sem_out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->maskrcnn, "pn_maskrcnn_forward: call pn_maskrcnn_build first");
  PN_REQUIRE(rgb_dev && sem_out_dev, "pn_maskrcnn_forward: null buffer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  c->maskrcnn->net.begin_forward(s);
  set_mrcnn_slots(*c->maskrcnn, rgb_dev, goal_cat_dev, score_thresh, sem_pred_prob_thr, goal_thr, sem_out_dev, s);
  c->maskrcnn->net.run(s);
  c->maskrcnn->net.end_forward(s);
  PN_CUDA_CHECK(cudaGetLastError());
  PN_API_END
}

int pn_maskrcnn_forward_host(pn_ctx* ctx, const uint8_t* rgb_host, const int* goal_cat_host, float score_thresh,
                             float sem_pred_prob_thr, float goal_thr, float* sem_out_host) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->maskrcnn, "pn_maskrcnn_forward_host: call pn_maskrcnn_build first");
  MaskRcnn& m = *c->maskrcnn;
  const size_t in_bytes = static_cast<size_t>(m.cfg.B) * m.cfg.H * m.cfg.W * 3;
  const size_t out_bytes = static_cast<size_t>(m.cfg.B) * m.cfg.H * m.cfg.W * (m.cfg.num_classes + 1) * sizeof(float);
  if (!m.stage_rgb) {
    m.stage_rgb = static_cast<uint8_t*>(m.net.arena.alloc(in_bytes + m.cfg.B * sizeof(int), false));
    m.stage_sem = static_cast<float*>(m.net.arena.alloc(out_bytes, false));
  }
  int* goal_dev = nullptr;
  m.net.begin_forward(c->stream);
  PN_CUDA_CHECK(cudaMemcpyAsync(m.stage_rgb, rgb_host, in_bytes, cudaMemcpyHostToDevice, c->stream));
  if (goal_cat_host) {
    goal_dev = reinterpret_cast<int*>(m.stage_rgb + ((in_bytes + 3) & ~size_t(3)));
    PN_CUDA_CHECK(cudaMemcpyAsync(goal_dev, goal_cat_host, m.cfg.B * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  }
  set_mrcnn_slots(m, m.stage_rgb, goal_dev, score_thresh, sem_pred_prob_thr, goal_thr, m.stage_sem, c->stream);
  m.net.run(c->stream);
  m.net.end_forward(c->stream);
  PN_CUDA_CHECK(cudaMemcpyAsync(sem_out_host, m.stage_sem, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  PN_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  PN_API_END
}

int pn_maskrcnn_num_launches(pn_ctx* ctx) {
  auto* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !c->maskrcnn) return -1;
  return static_cast<int>(c->maskrcnn->net.launches_per_forward);
}

int pn_maskrcnn_input_size(pn_ctx* ctx, int* resized_hw, int* padded_hw) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  PN_REQUIRE(c && c->maskrcnn, "pn_maskrcnn_input_size: not built");
  if (resized_hw) resized_hw[0] = c->maskrcnn->Hn, resized_hw[1] = c->maskrcnn->Wn;
  if (padded_hw) padded_hw[0] = c->maskrcnn->Hp, padded_hw[1] = c->maskrcnn->Wp;
  PN_API_END
}

int pn_maskrcnn_run_stages(pn_ctx* ctx, const char* first_stage, const char* end_stage, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->maskrcnn && first_stage && end_stage, "pn_maskrcnn_run_stages: not built");
  Net& net = c->maskrcnn->net;
  const int a = net.stage_begin(first_stage), b = net.stage_begin(end_stage);
  PN_REQUIRE(a <= b, "pn_maskrcnn_run_stages: stages out of order");
  net.run_range(a, b, static_cast<cudaStream_t>(stream));
  PN_CUDA_CHECK(cudaGetLastError());
  PN_API_END
}

int pn_maskrcnn_set_call(pn_ctx* ctx, const uint8_t* rgb_dev, const int* goal_cat_dev, float score_thresh,
                         float sem_pred_prob_thr, float goal_thr, float* sem_out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->maskrcnn, "pn_maskrcnn_set_call: not built");
  set_mrcnn_slots(*c->maskrcnn, rgb_dev, goal_cat_dev, score_thresh, sem_pred_prob_thr, goal_thr, sem_out_dev,
                  static_cast<cudaStream_t>(stream));
  PN_API_END
}

int pn_maskrcnn_tap(pn_ctx* ctx, const char* name, int write, int channels, void* buf_dev, int64_t buf_bytes, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->maskrcnn && name && buf_dev, "pn_maskrcnn_tap: not built");
  MaskRcnn& m = *c->maskrcnn;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const std::string key(name);
  if (key == "resized_u8") {
    const size_t bytes = static_cast<size_t>(m.cfg.B) * m.Hn * m.Wn * 3;
    PN_REQUIRE(!write && static_cast<size_t>(buf_bytes) >= bytes, "pn_maskrcnn_tap: resized_u8 is read-only / buffer too small");
    PN_CUDA_CHECK(cudaMemcpyAsync(buf_dev, m.resized_u8, bytes, cudaMemcpyDeviceToDevice, s));
  } else if (m.net.taps.count(key)) {
    const Tensor& t = m.net.taps[key];
    PN_REQUIRE(channels > 0 && channels <= t.C, "pn_maskrcnn_tap: bad channel count");
    const size_t bytes = static_cast<size_t>(t.pixels()) * channels * sizeof(float);
    PN_REQUIRE(static_cast<size_t>(buf_bytes) >= bytes, "pn_maskrcnn_tap: buffer too small");
    if (write) nchw_to_nhwc_direct(static_cast<const float*>(buf_dev), t, channels, s, !m.net.x3);
    else nhwc_to_nchw(t, channels, static_cast<float*>(buf_dev), s);
  } else if (m.net.raw_taps.count(key)) {
    auto& rt = m.net.raw_taps[key];
    const size_t bytes = std::min(rt.second, static_cast<size_t>(buf_bytes));
    if (write) PN_CUDA_CHECK(cudaMemcpyAsync(rt.first, buf_dev, bytes, cudaMemcpyDeviceToDevice, s));
    else PN_CUDA_CHECK(cudaMemcpyAsync(buf_dev, rt.first, bytes, cudaMemcpyDeviceToDevice, s));
  } else {
    PN_REQUIRE(false, "pn_maskrcnn_tap: unknown tap '" + key + "'");
  }
  PN_CUDA_CHECK(cudaGetLastError());
  PN_API_END
}

int pn_pil_bilinear_coeffs(int in_size, int out_size, int* bounds_out, int* coeffs_out, int* ksize_out) {
  PN_API_BEGIN
  PN_REQUIRE(in_size > 0 && out_size > 0 && bounds_out && coeffs_out && ksize_out, "pn_pil_bilinear_coeffs: bad arguments");
  std::vector<int> b, k;
  int ks = 0;
  pil_bilinear_coeffs(in_size, out_size, b, k, ks);
  PN_REQUIRE(ks <= *ksize_out, "pn_pil_bilinear_coeffs: coefficient buffer too narrow (pass its width in *ksize_out)");
  std::copy(b.begin(), b.end(), bounds_out);
  for (int i = 0; i < out_size; ++i)
    for (int j = 0; j < ks; ++j) coeffs_out[i * (*ksize_out) + j] = k[i * ks + j];
  *ksize_out = ks;
  PN_API_END
}

int pn_make_obs(pn_ctx* ctx, const float* depth_dev, const uint8_t* rgb_dev, const float* sem_dev, int E, int H, int W,
                int frame_height, int frame_width, int num_sem, float min_depth, float max_depth, float* obs_out_dev,
                void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(depth_dev && sem_dev && obs_out_dev && E > 0, "pn_make_obs: null buffer");
  PN_REQUIRE(frame_width > 0 && W % frame_width == 0 && H / (W / frame_width) == frame_height, "pn_make_obs: frame size must divide the camera size");
  const int ds = W / frame_width;
  launch_make_obs(depth_dev, rgb_dev, sem_dev, E, H, W, ds, frame_height, frame_width, num_sem, min_depth, max_depth, obs_out_dev,
                  static_cast<cudaStream_t>(stream));
  PN_CUDA_CHECK(cudaGetLastError());
  PN_API_END
}

int pn_target_pred(pn_ctx* ctx, const float* pred_dev, int num_classes, int window, int x1, int y1, int goal_cat, int r0,
                   int c0, int local_w, int local_h, const float* explored_dev, int64_t explored_row_stride,
                   float* target_out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(pred_dev && explored_dev && target_out_dev, "pn_target_pred: null buffer");
  PN_REQUIRE(goal_cat >= 0 && goal_cat < num_classes, "pn_target_pred: goal category out of range");
  PN_REQUIRE(window > 0 && local_w > 0 && local_h > 0 && explored_row_stride >= local_h, "pn_target_pred: bad geometry");
  launch_target_pred(pred_dev, window, x1, y1, goal_cat, r0, c0, local_w, local_h, explored_dev, explored_row_stride,
                     target_out_dev, static_cast<cudaStream_t>(stream));
  PN_CUDA_CHECK(cudaGetLastError());
  PN_API_END
}

namespace {
int map_call(pn_ctx* ctx, int op, const pn_map_cfg* cfg, const pn_map_arrays* a, int E, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(cfg && a && E > 0, "pn_map_*: null configuration / arrays or no environments");
  PN_REQUIRE(a->full_map && a->local_map && a->full_pose && a->local_pose && a->origins && a->lmb && a->planner_pose_inputs &&
                 a->loc && a->dist_to_goal,
             "pn_map_*: null state array");
  PN_REQUIRE(op != 2 || a->global_goal, "pn_map_update_local: global_goal is required");
  PN_REQUIRE(cfg->num_channels >= 4 && cfg->local_w > 0 && cfg->local_h > 0 && cfg->local_w <= cfg->full_w &&
                 cfg->local_h <= cfg->full_h && cfg->map_resolution > 0 && cfg->grid_resolution > 0 && cfg->col_rad >= 0,
             "pn_map_*: bad geometry");
  map_bookkeeping(op, *cfg, *a, E, static_cast<cudaStream_t>(stream));
  PN_API_END
}
}  // namespace

int pn_map_init(pn_ctx* ctx, const pn_map_cfg* cfg, const pn_map_arrays* arrays, int E, void* stream) {
  return map_call(ctx, 0, cfg, arrays, E, stream);
}
int pn_map_stamp_initial(pn_ctx* ctx, const pn_map_cfg* cfg, const pn_map_arrays* arrays, int E, void* stream) {
  return map_call(ctx, 1, cfg, arrays, E, stream);
}
int pn_map_update_local(pn_ctx* ctx, const pn_map_cfg* cfg, const pn_map_arrays* arrays, int E, void* stream) {
  return map_call(ctx, 2, cfg, arrays, E, stream);
}
int pn_map_update_full(pn_ctx* ctx, const pn_map_cfg* cfg, const pn_map_arrays* arrays, int E, void* stream) {
  return map_call(ctx, 3, cfg, arrays, E, stream);
}

int pn_map_stamp_local(pn_ctx* ctx, const float* local_map_dev, float* full_map_dev, const int* lmb_dev, int E, int num_channels,
                       int local_w, int local_h, int full_w, int full_h, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(local_map_dev && full_map_dev && lmb_dev, "pn_map_stamp_local: null buffer");
  PN_REQUIRE(E > 0 && num_channels > 0 && local_w > 0 && local_h > 0 && local_w <= full_w && local_h <= full_h,
             "pn_map_stamp_local: bad geometry");
  launch_map_stamp_local(local_map_dev, full_map_dev, lmb_dev, E, num_channels, local_w, local_h, full_w, full_h,
                         static_cast<cudaStream_t>(stream));
  PN_API_END
}

int pn_map_crop_window(pn_ctx* ctx, const float* full_map_dev, int E, int num_channels, int full_w, int full_h, int x1, int y1,
                       int win_w, int win_h, int copy_channels, float* window_out_dev, int out_channels, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(full_map_dev && window_out_dev, "pn_map_crop_window: null buffer");
  PN_REQUIRE(E > 0 && copy_channels > 0 && copy_channels <= num_channels && copy_channels <= out_channels,
             "pn_map_crop_window: bad channel counts");
  PN_REQUIRE(x1 >= 0 && y1 >= 0 && win_w > 0 && win_h > 0 && x1 + win_w <= full_w && y1 + win_h <= full_h,
             "pn_map_crop_window: window outside the full map");
  launch_map_crop(full_map_dev, E, num_channels, full_w, full_h, x1, y1, win_w, win_h, copy_channels, window_out_dev, out_channels,
                  static_cast<cudaStream_t>(stream));
  PN_API_END
}

int pn_map_quantize(pn_ctx* ctx, const float* map_dev, long long count, unsigned char* out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(count >= 0 && (count == 0 || (map_dev && out_dev)), "pn_map_quantize: null buffer");
  if (count > 0) launch_map_quantize(map_dev, count, out_dev, c->num_sms, static_cast<cudaStream_t>(stream));
  PN_API_END
}

int pn_map_sample(pn_ctx* ctx, const unsigned char* seq_dev, int T, int C, int W, int H, int t_idx, int goal_channel0, int num_goals,
                  float* img_hwc_dev, float* img_chw_dev, long long* gt_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(seq_dev != nullptr, "pn_map_sample: null sequence");
  PN_REQUIRE(T > 0 && C >= 2 && W > 0 && H > 0, "pn_map_sample: bad sequence shape (the explored channel is channel 1)");
  PN_REQUIRE(t_idx >= -T && t_idx < T, "pn_map_sample: time step out of range");
  PN_REQUIRE(!gt_dev || (num_goals > 0 && goal_channel0 >= 0 && goal_channel0 + num_goals <= C), "pn_map_sample: goal channels outside the map");
  PN_REQUIRE(img_hwc_dev || img_chw_dev || gt_dev, "pn_map_sample: no output requested");
  launch_map_sample(seq_dev, T, C, W, H, t_idx < 0 ? t_idx + T : t_idx, goal_channel0, gt_dev ? num_goals : 0, img_hwc_dev, img_chw_dev,
                    gt_dev, static_cast<cudaStream_t>(stream));
  PN_API_END
}

int pn_global_goal(pn_ctx* ctx, const pn_goal_cfg* cfg, const pn_goal_arrays* a, int E, int only_distance, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(cfg && a && E > 0, "pn_global_goal: null configuration / arrays or no environments");
  PN_REQUIRE(a->full_map && a->lmb && a->loc && a->dd, "pn_global_goal: null map / bounds / agent cell / distance buffer");
  PN_REQUIRE(only_distance || (a->target_pred && a->dd_wt && a->dd_wt_valid && a->global_goal && a->goal_kind &&
                               a->last_global_goal && a->last_kind),
             "pn_global_goal: null goal state array");
  PN_REQUIRE(cfg->num_channels >= 1 && cfg->local_w > 0 && cfg->local_h > 0 && cfg->local_w <= cfg->full_w &&
                 cfg->local_h <= cfg->full_h && cfg->map_resolution > 0,
             "pn_global_goal: bad geometry");
  launch_global_goal(c->device, c->num_sms, *cfg, *a, E, only_distance, static_cast<cudaStream_t>(stream));
  PN_API_END
}

int pn_goal_map(pn_ctx* ctx, const float* local_map_dev, int E, int num_channels, int local_w, int local_h,
                const int* goal_cat_dev, const int* skip_morph_dev, const int* global_goal_dev, int goal_erode,
                float* goal_map_out_dev, int* found_goal_out_dev, void* stream) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(local_map_dev && goal_cat_dev && skip_morph_dev && global_goal_dev && goal_map_out_dev && found_goal_out_dev,
             "pn_goal_map: null buffer");
  PN_REQUIRE(E > 0 && num_channels >= 5 && local_w > 0 && local_h > 0, "pn_goal_map: bad geometry");
  launch_goal_map(local_map_dev, E, num_channels, local_w, local_h, goal_cat_dev, skip_morph_dev, global_goal_dev, goal_erode,
                  goal_map_out_dev, found_goal_out_dev, static_cast<cudaStream_t>(stream));
  PN_API_END
}

int pn_conv_tuning_import(const char* text, int* entries_out) {
  PN_API_BEGIN
  PN_REQUIRE(text != nullptr, "pn_conv_tuning_import: null text");
  const int n = conv_tuning_import(text);
  if (entries_out) *entries_out = n;
  PN_API_END
}

int pn_conv_tuning_export(char* buf, int64_t buf_bytes, int64_t* needed_out) {
  PN_API_BEGIN
  const std::string t = conv_tuning_export();
  if (needed_out) *needed_out = static_cast<int64_t>(t.size()) + 1;
  if (buf && buf_bytes > 0) {
    PN_REQUIRE(static_cast<size_t>(buf_bytes) > t.size(), "pn_conv_tuning_export: buffer too small");
    std::memcpy(buf, t.c_str(), t.size() + 1);
  }
  PN_API_END
}

int pn_conv_tuning_clear(void) {
  PN_API_BEGIN
  conv_tuning_clear();
  PN_API_END
}

int pn_conv2d(pn_ctx* ctx, int precision, const float* x_dev, int B, int Cin, int H, int W, const float* w_host,
              const float* scale_host, const float* bias_host, const float* residual_dev, int Cout, int R, int S,
              int stride, int dil, int pad, int relu, int force_bn, float* y_dev) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(x_dev && w_host && y_dev, "pn_conv2d: null buffer");
  PN_REQUIRE(precision == PN_BF16 || precision == PN_TF32 || precision == PN_FP32, "bad precision");
  const DType dt = precision == PN_BF16 ? kBF16 : kF32;
  Net net;
  net.num_sms = c->num_sms;
  net.dt = dt;
  net.x3 = precision == PN_FP32;
  net.use_graph = false;
  struct Slots {
    const float* x;
    const float* r;
  };
  Slots h{x_dev, residual_dev};
  Slots* slots = static_cast<Slots*>(net.arena.alloc(sizeof(Slots)));
  PN_CUDA_CHECK(cudaMemcpy(slots, &h, sizeof(h), cudaMemcpyHostToDevice));
  Tensor x = net.arena.tensor(B, H, W, pad_channels(Cin, dt), dt);
  add_nchw_to_nhwc(net, &slots->x, x, Cin);
  const int Ho = conv_out(H, R, stride, dil, pad), Wo = conv_out(W, S, stride, dil, pad);
  PN_REQUIRE(Ho > 0 && Wo > 0, "pn_conv2d: empty output");
  Tensor y = net.arena.tensor(B, Ho, Wo, pad_channels(Cout, dt), dt);
  Tensor res;
  if (residual_dev) {
    res = net.arena.tensor(B, Ho, Wo, pad_channels(Cout, dt), dt);
    add_nchw_to_nhwc(net, &slots->r, res, Cout);
  }
  ConvSpec sp;
  sp.Cin = Cin, sp.Cout = Cout, sp.R = R, sp.S = S, sp.stride = stride, sp.dil = dil, sp.pad = pad;
  sp.relu = relu != 0;
  sp.force_bn = force_bn & 0xfff;
  sp.force_direct_epilogue = (force_bn & 0x1000) != 0;
  sp.force_splits = (force_bn >> 16) & 0xff;
  sp.force_opt = (force_bn >> 24) & 7;
  sp.force_pair = (force_bn & 0x2000) ? 1 : ((force_bn & 0x4000) ? 2 : 0);
  add_conv(net, "pn_conv2d", x, y, w_host, scale_host, bias_host, sp, residual_dev ? &res : nullptr);
  net.run(nullptr);
  if (sp.force_splits > 1) net.run(nullptr);  // a second pass over the same workspace: split-K must leave it clean
  nhwc_to_nchw(y, Cout, y_dev, nullptr);
  PN_CUDA_CHECK(cudaDeviceSynchronize());
  PN_API_END
}

int pn_conv_bench(pn_ctx* ctx, int precision, int B, int Cin, int H, int W, int Cout, int R, int S, int stride, int dil,
                  int pad, int with_residual, int force_bn, int iters, float* ms_out, int* bn_out) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(ms_out && iters > 0, "pn_conv_bench: bad arguments");
  PN_REQUIRE(precision == PN_BF16 || precision == PN_TF32 || precision == PN_FP32, "bad precision");
  const DType dt = precision == PN_BF16 ? kBF16 : kF32;
  Net net;
  net.num_sms = c->num_sms;
  net.dt = dt;
  net.x3 = precision == PN_FP32;
  net.use_graph = false;
  Tensor x = net.arena.tensor(B, H, W, pad_channels(Cin, dt), dt);
  PN_CUDA_CHECK(cudaMemset(x.ptr, 0x3c, x.bytes()));
  const int Ho = conv_out(H, R, stride, dil, pad), Wo = conv_out(W, S, stride, dil, pad);
  PN_REQUIRE(Ho > 0 && Wo > 0, "pn_conv_bench: empty output");
  Tensor y = net.arena.tensor(B, Ho, Wo, pad_channels(Cout, dt), dt);
  Tensor res;
  if (with_residual) {
    res = net.arena.tensor(B, Ho, Wo, pad_channels(Cout, dt), dt);
    PN_CUDA_CHECK(cudaMemset(res.ptr, 0x3c, res.bytes()));
  }
  std::vector<float> w(static_cast<size_t>(Cout) * Cin * R * S, 0.01f), sc(Cout, 1.f), bi(Cout, 0.f);
  ConvSpec sp;
  sp.Cin = Cin, sp.Cout = Cout, sp.R = R, sp.S = S, sp.stride = stride, sp.dil = dil, sp.pad = pad, sp.relu = true;
  sp.force_bn = force_bn & 0xfff;
  sp.force_splits = (force_bn >> 16) & 0xff;
  sp.force_opt = (force_bn >> 24) & 7;
  sp.force_pair = (force_bn & 0x2000) ? 1 : ((force_bn & 0x4000) ? 2 : 0);
  long long* dbg = nullptr;
  if (std::getenv("PN_CONV_DBG")) {
    dbg = static_cast<long long*>(net.arena.alloc((1024 + 1 + 64 * 4) * sizeof(long long)));
    sp.dbg = dbg;
    if (const char* sk = std::getenv("PN_EPI_SKIP")) sp.dbg_skip = std::atoi(sk);
  }
  add_conv(net, "bench", x, y, w.data(), sc.data(), bi.data(), sp, with_residual ? &res : nullptr);
  if (bn_out) *bn_out = net.last_bn;
  cudaEvent_t e0, e1;
  PN_CUDA_CHECK(cudaEventCreate(&e0));
  PN_CUDA_CHECK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) net.run_eager(c->stream);
  PN_CUDA_CHECK(cudaEventRecord(e0, c->stream));
  for (int i = 0; i < iters; ++i) net.run_eager(c->stream);
  PN_CUDA_CHECK(cudaEventRecord(e1, c->stream));
  PN_CUDA_CHECK(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  PN_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *ms_out = ms / iters;
  if (dbg) {  // timeline of CTA 0 in the last launch, cycles relative to its first event
    std::vector<long long> h(1024 + 1 + 64 * 4);
    PN_CUDA_CHECK(cudaMemcpy(h.data(), dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    long long t0 = 0;
    for (int i = 0; i < 1024; ++i) if (h[i] && (!t0 || h[i] < t0)) t0 = h[i];
    std::printf("# CTA0 timeline (cycles): tile: load0 | mma: enter, acc free, first mma, commit | epi: enter, acc full, done | group pairs\n");
    for (int t = 0; t < 64; ++t) {
      if (!h[t * 16 + 6]) break;
      std::printf("tile %2d:", t);
      for (int k = 0; k < 16; ++k) std::printf(" %7lld", h[t * 16 + k] ? h[t * 16 + k] - t0 : -1);
      std::printf("\n");
    }
    std::printf("# CTA0 launch log (ns): entry->wait_done, wait_done->exit, exit->next wait_done\n");
    for (long long i = 0; i < h[1024] && i < 64; ++i) {
      const long long* e = &h[1025 + 4 * i];
      std::printf("launch %2lld: %7lld %7lld %7lld\n", i, e[1] - e[0], e[2] - e[1], i + 1 < h[1024] ? h[1025 + 4 * (i + 1) + 1] - e[2] : -1);
    }
  }
  PN_API_END
}

}  // extern "C"
