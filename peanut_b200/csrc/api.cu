// C-ABI of libpeanut_b200.so (declared in include/peanut_b200.h).  Every entry point translates
// C++ exceptions into a status code + thread-local message; no torch types cross this boundary.
#include <cmath>
#include <cstring>
#include <memory>
#include <mutex>

#include "../../include/peanut_b200.h"
#include "engine.h"
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
  cudaStream_t stream = nullptr;  // used by the *_host entry points
};

thread_local std::string g_last_error;

__global__ void set_pred_slots_kernel(PredSlots* slots, const float* in, float* out, int apply_sigmoid) {
  slots->input = in;
  slots->output = out;
  slots->apply_sigmoid = apply_sigmoid;
}

// NHWC (dt) -> NCHW fp32, first C channels.
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* in, long long ldi, float* out, int B, int C, int HW) {
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
    nhwc_to_nchw_kernel<__nv_bfloat16><<<blocks, threads, 0, s>>>(static_cast<const __nv_bfloat16*>(t.ptr), t.ld, out, t.B, C, HW);
  else
    nhwc_to_nchw_kernel<float><<<blocks, threads, 0, s>>>(static_cast<const float*>(t.ptr), t.ld, out, t.B, C, HW);
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
int pn_abi_version(void) { return 1; }

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
  PN_REQUIRE(precision == PN_BF16 || precision == PN_TF32, "pn_prednet_build: bad precision");
  c->prednet.reset();
  auto net = std::make_unique<PredNet>();
  net->net.num_sms = c->num_sms;
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
  set_pred_slots_kernel<<<1, 1, 0, s>>>(c->prednet->slots, map_dev, out_dev, apply_sigmoid);
  c->prednet->net.run(s);
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
  PN_CUDA_CHECK(cudaMemcpyAsync(n.stage_in, map_host, in_bytes, cudaMemcpyHostToDevice, c->stream));
  set_pred_slots_kernel<<<1, 1, 0, c->stream>>>(n.slots, n.stage_in, n.stage_out, apply_sigmoid);
  n.net.run(c->stream);
  PN_CUDA_CHECK(cudaMemcpyAsync(out_host, n.stage_out, out_bytes, cudaMemcpyDeviceToHost, c->stream));
  PN_CUDA_CHECK(cudaStreamSynchronize(c->stream));
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
  const double agent_height = static_cast<double>(cfg->camera_height) * 100.0;
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
  g.f = static_cast<float>((g.w / 2.0) / std::tan(static_cast<double>(cfg->hfov) / 2.0 * 3.14159265358979323846 / 180.0));
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

int pn_prednet_profile(pn_ctx* ctx, int iters, float* ms_out, int max_ops, char* names_out, int names_bytes) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(c->prednet && ms_out && iters > 0, "pn_prednet_profile: not built");
  Net& net = c->prednet->net;
  const int n = static_cast<int>(net.ops.size());
  PN_REQUIRE(max_ops >= n, "pn_prednet_profile: ms_out too small");
  PredNet& pnn = *c->prednet;
  const size_t in_bytes = static_cast<size_t>(pnn.B) * pnn.C * pnn.H * pnn.W * sizeof(float);
  const size_t out_bytes = static_cast<size_t>(pnn.B) * pnn.num_classes * pnn.H * pnn.W * sizeof(float);
  if (!pnn.stage_in) {
    pnn.stage_in = static_cast<float*>(net.arena.alloc(in_bytes, true));
    pnn.stage_out = static_cast<float*>(net.arena.alloc(out_bytes, false));
  }
  set_pred_slots_kernel<<<1, 1, 0, c->stream>>>(pnn.slots, pnn.stage_in, pnn.stage_out, 0);
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
    names += net.op_names[i];
    names += '\n';
  }
  if (names_out && names_bytes > 0) {
    const size_t k = std::min(names.size(), static_cast<size_t>(names_bytes - 1));
    memcpy(names_out, names.data(), k);
    names_out[k] = 0;
  }
  return n > 0 ? 0 : 0;
  PN_API_END
}

int pn_prednet_num_ops(pn_ctx* ctx) {
  auto* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !c->prednet) return -1;
  return static_cast<int>(c->prednet->net.ops.size());
}

int pn_conv2d(pn_ctx* ctx, int precision, const float* x_dev, int B, int Cin, int H, int W, const float* w_host,
              const float* scale_host, const float* bias_host, const float* residual_dev, int Cout, int R, int S,
              int stride, int dil, int pad, int relu, int force_bn, float* y_dev) {
  PN_API_BEGIN
  auto* c = reinterpret_cast<Ctx*>(ctx);
  check_device(c);
  PN_REQUIRE(x_dev && w_host && y_dev, "pn_conv2d: null buffer");
  const DType dt = precision == PN_BF16 ? kBF16 : kF32;
  Net net;
  net.num_sms = c->num_sms;
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
  add_conv(net, "pn_conv2d", x, y, w_host, scale_host, bias_host, sp, residual_dev ? &res : nullptr);
  net.run(nullptr);
  nhwc_to_nchw(y, Cout, y_dev, nullptr);
  PN_CUDA_CHECK(cudaDeviceSynchronize());
  PN_API_END
}

}  // extern "C"
