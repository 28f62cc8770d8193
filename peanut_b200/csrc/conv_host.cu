// Host side of the tcgen05 implicit-GEMM convolution: weight packing, TMA descriptor encoding
// (im2col mode for activations, tiled mode for weights), tile selection and launch recording.
#include <cmath>
#include <vector>
#include <string>
#include <map>
#include <cstdlib>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <chrono>
#include <mutex>
#include <sstream>
#include <thread>

#include "conv_umma.cuh"
#include "engine.h"

namespace pn {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void* driver_entry(const char* name) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  PN_CUDA_CHECK(cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q));
  PN_REQUIRE(fn != nullptr && q == cudaDriverEntryPointSuccess, std::string("driver entry point not found: ") + name);
  return fn;
}

CUtensorMapSwizzle swizzle_enum(int sw) {
  return sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}
CUtensorMapDataType dtype_enum(DType dt) {
  return dt == kBF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
}

CUtensorMap encode_tiled_2d(DType dt, const void* base, uint64_t inner, uint64_t rows, uint64_t row_stride_bytes,
                            uint32_t box_inner, uint32_t box_rows, int sw) {
  static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_entry("cuTensorMapEncodeTiled"));
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, dtype_enum(dt), 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(sw), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
  return m;
}

// (C, W, H, N) activation tensor, one box = `pixels` output positions x `channels` channels.
CUtensorMap encode_im2col(DType dt, const Tensor& in, int channels_total, int pad_h, int pad_w, int dil, int R, int S,
                          int stride_h, int stride_w, uint32_t box_channels, uint32_t pixels, int sw) {
  static EncodeIm2colFn fn = reinterpret_cast<EncodeIm2colFn>(driver_entry("cuTensorMapEncodeIm2col"));
  CUtensorMap m;
  const uint64_t es = dtype_size(dt);
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(channels_total), static_cast<cuuint64_t>(in.W),
                        static_cast<cuuint64_t>(in.H), static_cast<cuuint64_t>(in.B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(in.ld) * es, static_cast<cuuint64_t>(in.ld) * es * in.W,
                           static_cast<cuuint64_t>(in.ld) * es * in.W * in.H};
  // Bounding box of base pixels: [-pad, extent + pad - (taps-1)*dil) in each spatial dimension.
  int lower[2] = {-pad_w, -pad_h};
  int upper[2] = {pad_w - (S - 1) * dil, pad_h - (R - 1) * dil};
  cuuint32_t estr[4] = {1, static_cast<cuuint32_t>(stride_w), static_cast<cuuint32_t>(stride_h), 1};
  CUresult r = fn(&m, dtype_enum(dt), 4, in.ptr, dims, strides, lower, upper, box_channels, pixels, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_enum(sw), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  PN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeIm2col failed with code " + std::to_string(static_cast<int>(r)));
  // Known driver issue (<= 13.1) for im2col maps over tensors smaller than 128 KiB: bit 21 of the
  // second descriptor word must be cleared (same workaround CUTLASS applies).
  int drv = 0;
  cudaDriverGetVersion(&drv);
  const uint64_t span = static_cast<uint64_t>(in.ld) * es * in.W * in.H * in.B;
  if (drv <= 13010 && span < 131072) reinterpret_cast<uint64_t*>(&m)[1] &= ~(1ull << 21);
  return m;
}

// Automatic split-K (cluster reduce-scatter over distributed shared memory) for layers with too few tiles to fill the GPU.
constexpr bool kAutoSplitK = true;
// CTA pairs (tcgen05.mma.cta_group::2) picked by the tile model; sp.force_pair overrides (1 = on, 2 = off).
constexpr bool kAutoPair = true;
// per-tile fixed cost of the persistent loop (barrier round trips, TMA issue, accumulator hand-over), measured ~0.3 us
constexpr double kTileFixed = 0.3e-6;

struct ConvMaps {
  CUtensorMap a, b, out, res;
};

template <typename T, int BN, bool kSplit, bool kPair>
void launch_conv(const ConvMaps& tm, const ConvParams& p, int grid, size_t smem, cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    PN_CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<T, BN, kSplit, kPair>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid), cfg.blockDim = dim3(kNumThreads), cfg.dynamicSmemBytes = smem, cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = debug_no_pdl() ? 0 : 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  if (kSplit || kPair) {  // one thread-block cluster per output tile (split-K) / one SM pair per 256-row tile
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = (kPair ? 2 : 1) * (kSplit ? p.splits : 1), attr[1].val.clusterDim.y = 1, attr[1].val.clusterDim.z = 1;
    cfg.numAttrs = 2;
  }
  PN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, conv_umma_kernel<T, BN, kSplit, kPair>, tm.a, tm.b, tm.out, tm.res, p));
}

template <typename T>
void launch_conv_bn(int bn, bool pair, const ConvMaps& tm, const ConvParams& p, int grid, size_t smem, cudaStream_t s) {
  if (pair && p.splits > 1) {  // a cluster of `splits` SM pairs per 256-row tile
    switch (bn) {
      case 128: launch_conv<T, 128, true, true>(tm, p, grid, smem, s); break;
      case 256: launch_conv<T, 256, true, true>(tm, p, grid, smem, s); break;
      default: PN_REQUIRE(false, "unsupported N tile for a CTA pair");
    }
    return;
  }
  if (pair) {
    switch (bn) {
      case 128: launch_conv<T, 128, false, true>(tm, p, grid, smem, s); break;
      case 256: launch_conv<T, 256, false, true>(tm, p, grid, smem, s); break;
      default: PN_REQUIRE(false, "unsupported N tile for a CTA pair");
    }
    return;
  }
  if (p.splits > 1) {
    switch (bn) {
      case 32: launch_conv<T, 32, true, false>(tm, p, grid, smem, s); break;
      case 64: launch_conv<T, 64, true, false>(tm, p, grid, smem, s); break;
      case 128: launch_conv<T, 128, true, false>(tm, p, grid, smem, s); break;
      case 256: launch_conv<T, 256, true, false>(tm, p, grid, smem, s); break;
      default: PN_REQUIRE(false, "unsupported N tile");
    }
    return;
  }
  switch (bn) {
    case 32: launch_conv<T, 32, false, false>(tm, p, grid, smem, s); break;
    case 64: launch_conv<T, 64, false, false>(tm, p, grid, smem, s); break;
    case 128: launch_conv<T, 128, false, false>(tm, p, grid, smem, s); break;
    case 256: launch_conv<T, 256, false, false>(tm, p, grid, smem, s); break;
    default: PN_REQUIRE(false, "unsupported N tile");
  }
}

// Launch modes the autotuner adds to its shortlist beyond tile width / split-K / CTA pairs.  Bit 1: split-K over clusters
// of CTA pairs, bit 2: two CTAs per SM.  Both are parity-tested (tests/test_conv_gpu.py) and win on individual layers
// (profiles/r01_conv_mode_sweep_pairsplit_tf32_b1.txt, profiles/r01_conv_two_ctas_per_sm_bf16_b8.txt).  Round 1 kept them
// opt-in because the 8-environment pipeline stalled with them; the stall was a tensor-memory hold-and-wait cycle between
// early-resident CTAs of two streams, removed by allocating tensor memory after griddepcontrol.wait (conv_umma.cuh,
// profiles/r02_stall_analysis.txt).  PN_CONV_TUNE_EXTRA overrides (0 = neither).
int tune_extra() {
  static const int v = [] {
    const char* e = std::getenv("PN_CONV_TUNE_EXTRA");
    return e ? std::atoi(e) : 3;
  }();
  return v;
}

// PN_CONV_AUTOTUNE: "0" = the tile model's choice only (no table, no timing); "table" = the imported table, else the model's
// choice (never times anything: strictly reproducible); anything else / unset = the table, else time the shortlist.
int autotune_mode() {
  static const int mode = [] {
    const char* e = std::getenv("PN_CONV_AUTOTUNE");
    if (e && e[0] == '0') return 0;
    if (e && std::strcmp(e, "table") == 0) return 1;
    return 2;
  }();
  return mode;
}

struct TuneEntry {
  int bn = 0, splits = 1, pair = 0, opt = 0;
};
std::mutex g_tune_mu;
std::map<std::string, TuneEntry>& tune_table() {
  static std::map<std::string, TuneEntry> t;
  return t;
}

// Average time of back-to-back launches (2 warm-up + 6 timed) on a private stream, in milliseconds.
// Watchdog: a launch configuration that has not finished after kTuneWatchdogMs cannot be cancelled (a running kernel can
// only be removed with its context), so the process reports WHICH configuration it was and exits instead of hanging the box.
constexpr int kTuneWatchdogMs = 8000;
template <typename F>
float time_launches(F&& launch, const char* what) {
  static cudaStream_t stream = nullptr;
  static cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (!stream) {
    PN_CUDA_CHECK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    PN_CUDA_CHECK(cudaEventCreate(&e0));
    PN_CUDA_CHECK(cudaEventCreate(&e1));
  }
  PN_CUDA_CHECK(cudaDeviceSynchronize());  // the layer's buffers may still be being initialised on other streams
  for (int i = 0; i < 2; ++i) launch(stream);
  PN_CUDA_CHECK(cudaEventRecord(e0, stream));
  for (int i = 0; i < 6; ++i) launch(stream);
  PN_CUDA_CHECK(cudaEventRecord(e1, stream));
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    const cudaError_t q = cudaEventQuery(e1);
    if (q == cudaSuccess) break;
    if (q != cudaErrorNotReady) PN_CUDA_CHECK(q);
    const auto ms = std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
    if (ms > kTuneWatchdogMs) {
      std::fprintf(stderr, "peanut_b200: conv launch configuration did not finish within %d ms: %s\n", kTuneWatchdogMs, what);
      std::fflush(stderr);
      std::_Exit(97);
    }
    if (ms > 2) std::this_thread::sleep_for(std::chrono::microseconds(200));
  }
  float ms = 0.f;
  PN_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / 6.f;
}

inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return static_cast<uint16_t>((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}

}  // namespace

int pad_channels(int c, DType dt) {
  (void)dt;
  if (c <= 16) return 16;
  if (c <= 32) return 32;
  return round_up(c, 64);
}

void fold_bn(const WeightStore& w, const std::string& prefix, int C, std::vector<float>& scale,
             std::vector<float>& bias, float eps) {
  const HostArray& g = get_weight(w, prefix + ".weight");
  const HostArray& b = get_weight(w, prefix + ".bias");
  const HostArray& mu = get_weight(w, prefix + ".running_mean");
  const HostArray& var = get_weight(w, prefix + ".running_var");
  PN_REQUIRE(g.numel() == C && b.numel() == C && mu.numel() == C && var.numel() == C, "BN size mismatch at " + prefix);
  scale.resize(C);
  bias.resize(C);
  for (int i = 0; i < C; ++i) {
    const float inv = 1.0f / std::sqrt(var.data[i] + eps);
    scale[i] = g.data[i] * inv;
    bias[i] = b.data[i] - mu.data[i] * scale[i];
  }
}

// ---- fp32 parity mode (Net::x3): a convolution as three tf32 products.
// x = hi + lo with hi = tf32(x) (round to nearest) and lo = tf32(x - hi); w likewise.  x*w = hi*whi + hi*wlo + lo*whi + O(2^-22 |x w|):
// the dropped lo*wlo term and the rounding of lo are both 2^-22 relative, the accumulation is the tensor core's fp32.  The three
// products are ONE convolution over 3 x the channels: the split kernel writes [hi | hi | lo] per pixel into a scratch tensor and the
// weights are packed as [whi | wlo | whi] per filter tap, so the kernel, its epilogue and the launch machinery are the tf32 path's.
__device__ __forceinline__ float rn_tf32(float v) {
  uint32_t q;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q) : "f"(v));
  return __uint_as_float(q);
}
__global__ void split3_kernel(const float* __restrict__ in, long long ldi, int C4, float* __restrict__ out, long long ldo, int cpad,
                              long long total) {
  pdl_grid_sync();
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const long long pix = idx / C4;
  const int c = static_cast<int>(idx - pix * C4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(in + pix * ldi + c);
  float4 hi, lo;
  hi.x = rn_tf32(v.x), hi.y = rn_tf32(v.y), hi.z = rn_tf32(v.z), hi.w = rn_tf32(v.w);
  lo.x = rn_tf32(v.x - hi.x), lo.y = rn_tf32(v.y - hi.y), lo.z = rn_tf32(v.z - hi.z), lo.w = rn_tf32(v.w - hi.w);
  float* o = out + pix * ldo + c;
  *reinterpret_cast<float4*>(o) = hi;
  *reinterpret_cast<float4*>(o + cpad) = hi;
  *reinterpret_cast<float4*>(o + 2 * cpad) = lo;
}

inline float host_rn_tf32(float v) {
  uint32_t u;
  std::memcpy(&u, &v, 4);
  if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
  std::memcpy(&v, &u, 4);
  return v;
}

static void add_conv_x3(Net& net, const std::string& name, const Tensor& in, const Tensor& out, const float* weight,
                        const float* scale, const float* bias, const ConvSpec& sp, const Tensor* residual) {
  PN_REQUIRE(in.dt == kF32, name + ": the fp32 split-precision path needs fp32 activations");
  PN_REQUIRE(in.C % 4 == 0 && in.ld % 4 == 0 && reinterpret_cast<uintptr_t>(in.ptr) % 16 == 0, name + ": input of the split kernel not 16-byte aligned");
  const int cpad = in.C;
  const int C3 = round_up(3 * cpad, 32);  // K blocks of 32 floats (128-byte swizzle rows); the tail stays zero
  Tensor s3 = net.arena.tensor(in.B, in.H, in.W, C3, kF32);
  {
    const int C4 = cpad / 4;
    const long long total = in.pixels() * C4;
    const int threads = 256;
    const long long blocks = (total + threads - 1) / threads;
    PN_REQUIRE(blocks < (1ll << 31), name + ": split launch too large");
    const float* src = static_cast<const float*>(in.ptr);
    float* dst = static_cast<float*>(s3.ptr);
    const long long ldi = in.ld, ldo = s3.ld;
    net.add(name + ".split3", [=](cudaStream_t s) {
      launch_pdl(split3_kernel, static_cast<int>(blocks), threads, 0, s, src, ldi, C4, dst, ldo, cpad, total);
    });
    net.launches_per_forward += 1;
  }
  const int taps = sp.R * sp.S;
  std::vector<float> w3(static_cast<size_t>(sp.Cout) * C3 * taps, 0.f);
  for (int co = 0; co < sp.Cout; ++co)
    for (int ci = 0; ci < sp.Cin; ++ci)
      for (int t = 0; t < taps; ++t) {
        const float w = weight[(static_cast<size_t>(co) * sp.Cin + ci) * taps + t] * (scale ? scale[co] : 1.f);
        const float hi = host_rn_tf32(w), lo = host_rn_tf32(w - hi);
        float* row = w3.data() + static_cast<size_t>(co) * C3 * taps;
        row[static_cast<size_t>(ci) * taps + t] = hi;
        row[static_cast<size_t>(cpad + ci) * taps + t] = lo;
        row[static_cast<size_t>(2 * cpad + ci) * taps + t] = hi;
      }
  ConvSpec sp3 = sp;
  sp3.Cin = C3;
  sp3.x3_cin = sp.Cin;
  add_conv(net, name, s3, out, w3.data(), nullptr, bias, sp3, residual);
}

void add_conv(Net& net, const std::string& name, const Tensor& in, const Tensor& out, const float* weight,
              const float* scale, const float* bias, const ConvSpec& sp, const Tensor* residual) {
  if (net.x3 && !sp.x3_cin) return add_conv_x3(net, name, in, out, weight, scale, bias, sp, residual);
  const DType dt = in.dt;
  const int es = static_cast<int>(dtype_size(dt));
  const int stride_w = sp.stride_w >= 0 ? sp.stride_w : sp.stride;
  const int pad_w = sp.pad_w >= 0 ? sp.pad_w : sp.pad;
  const int Ho = conv_out(in.H, sp.R, sp.stride, sp.dil, sp.pad);
  const int Wo = conv_out(in.W, sp.S, stride_w, sp.dil, pad_w);
  PN_REQUIRE(out.B == in.B && out.H == Ho && out.W == Wo, name + ": output shape mismatch");
  PN_REQUIRE(in.C >= sp.Cin, name + ": input view has fewer channels than the filter");
  PN_REQUIRE(out.dt == dt || (sp.out_fp32 && out.dt == kF32), name + ": output dtype mismatch");
  PN_REQUIRE(reinterpret_cast<uintptr_t>(in.ptr) % 16 == 0 && (in.ld * es) % 16 == 0, name + ": input not 16B aligned");
  PN_REQUIRE(reinterpret_cast<uintptr_t>(out.ptr) % 16 == 0 && (out.ld * dtype_size(out.dt)) % 16 == 0,
             name + ": output not 16B aligned");

  // K blocking: one smem row (= swizzle span) holds block_k channels of one filter tap.
  const int cin_pad = in.C;  // caller pads activations (pad_channels)
  int sw = cin_pad * es >= 128 ? 128 : cin_pad * es;
  PN_REQUIRE(sw == 32 || sw == 64 || sw == 128, name + ": unsupported channel padding");
  const int block_k = sw / es;
  PN_REQUIRE(cin_pad % block_k == 0, name + ": channels not a multiple of the K block");
  const int kb_per_tap = cin_pad / block_k;
  const int taps = sp.R * sp.S;
  const long long ktot = static_cast<long long>(taps) * cin_pad;

  // N tile: the widest tile that still yields at least ~one wave of CTAs.
  const int cout32 = round_up(sp.Cout, 32);
  const long long M = static_cast<long long>(in.B) * Ho * Wo;
  PN_REQUIRE(M < (1ll << 31), name + ": M too large");
  const int m_tiles = static_cast<int>((M + kBlockM - 1) / kBlockM);
  // N tile: the kernel streams (128 + bn) * K bytes of operands per tile through one SM's TMA path (~125 GB/s per SM,
  // ~14 TB/s out of L2 in aggregate, measured with tools/conv_tune.py); the MMA itself needs max(bn/2, 32) cycles
  // per 32-byte K step (below N = 64 the A-operand shared-memory read dominates).  Pick the tile that minimises
  // waves * max(stream time, MMA time): wide tiles for big layers, narrower ones when the grid would not fill.
  int bn = 32, splits = 1;
  bool pair = false;
  struct Cand {
    double t;
    int bn, splits;
    bool pair;
    int opt = 0;  // 1: two CTAs per SM (half-size shared-memory footprint), 2: bias through the tensor core on multi-tile launches, 4: weights resident
  };
  std::vector<Cand> cands;  // every valid launch configuration with its modelled time
  {
    const int kblocks_total = taps * kb_per_tap;
    double best = 1e30;
    for (int cand : {256, 128, 64, 32}) {
      if (sp.force_bn ? cand != (sp.force_bn & 0x3ff) : (cand > cout32 || cout32 % cand != 0)) continue;
      // CTA pair (cta_group::2): 256 x cand tile on two SMs, each staging 128 activation rows + cand/2 weight rows
      if (cand >= 128 && sw == 128 && m_tiles >= 2 && sp.force_pair != 2 && (kAutoPair || sp.force_pair == 1)) {
        for (int s : {1, 2, 4}) {  // s > 1: split-K over a cluster of s pairs (2 s CTAs, at most 8)
          if (sp.force_splits ? s != sp.force_splits
                              : (s > 1 && (sp.force_pair == 1 || sp.no_split || !kAutoSplitK || !(tune_extra() & 1) ||
                                           kblocks_total / s < 4)))
            continue;
          if (s > kblocks_total) continue;
          const int kbs = (kblocks_total + s - 1) / s;
          if (s > 1 && kbs * (s - 1) >= kblocks_total) continue;
          if (s > 1 && ((cand / s) % 8 != 0 || 8 * (kBlockM + cand / 2) * sw < kBlockM * cand * 4)) continue;
          const double work = std::ceil(m_tiles / 2.0) * ((cout32 + cand - 1) / cand);  // 256-row tiles
          const double slots = s > 1 ? std::floor(128.0 / (2 * s)) : net.num_sms / 2;     // clusters resident at once
          const double active = std::min<double>(work, slots) * 2.0 * s;
          const double waves = std::ceil(work / slots);
          const double bytes = static_cast<double>(kbs) * block_k * es * (kBlockM + cand / 2);
          const double t_mem = bytes / std::min(125e9, 14e12 / active);
          const double t_mma = static_cast<double>(kbs) * block_k * es / 32.0 * std::max(cand / 2.0, 32.0) / 1.9e9;
          // the epilogue of a 128 x cand tile takes ~0.5 us per 32 columns (issue-bound: one epilogue warp per SM sub-partition, measured) and overlaps
          // the next tile's main loop; pairs only pay off where the main loop is the longer of the two
          const double t_epi = 0.5e-6 * cand / 32.0 / s;
          const double t_red = s > 1 ? 1.5e-6 + 2.0 * kBlockM * cand * 4.0 / 200e9 : 0.0;
          const double t = waves * (std::max(std::max(t_mem, t_mma), s > 1 ? 0.0 : t_epi) + t_red + kTileFixed) + t_epi +
                           0.5e-6;  // + cluster set-up / tear-down
          cands.push_back({t, cand, s, true});
          if (t < best || sp.force_pair == 1) best = t, bn = cand, splits = s, pair = true;
        }
      }
      if (sp.force_pair == 1) continue;
      for (int s : {1, 2, 4, 8}) {
        if (sp.force_splits ? s != sp.force_splits : (s > 1 && (sp.no_split || !kAutoSplitK || kblocks_total / s < 4))) continue;
        if (s > kblocks_total) continue;
        const int kbs = (kblocks_total + s - 1) / s;
        if (s > 1 && kbs * (s - 1) >= kblocks_total) continue;  // the last split would be empty
        // cluster split-K: each rank finishes cand / s columns in 8-column units, and the fp32 partial tile is parked
        // in the operand ring (at most 8 stages deep)
        if (s > 1 && ((cand / s) % 8 != 0 || sw != 128 || 8 * (kBlockM + cand) * sw < kBlockM * cand * 4)) continue;
        const double work = static_cast<double>(m_tiles) * ((cout32 + cand - 1) / cand) * s;
        const double active = std::min<double>(work, net.num_sms);
        const double waves = std::ceil(work / (s > 1 ? 128.0 : net.num_sms));  // clusters cannot use every SM of a GPC
        const double bytes = static_cast<double>(kbs) * block_k * es * (kBlockM + cand);
        const double t_mem = bytes / std::min(125e9, 14e12 / active);
        const double t_mma = static_cast<double>(kbs) * block_k * es / 32.0 * std::max(cand / 2.0, 32.0) / 1.9e9;
        const double t_epi = 0.5e-6 * cand / 32.0 / s;  // epilogue of one tile (~0.5 us per 32 columns); the last one is exposed
        // split-K: park the partial tile in shared memory, cluster barrier, read one slice of every peer's tile
        const double t_red = s > 1 ? 1.5e-6 + 2.0 * kBlockM * cand * 4.0 / 200e9 : 0.0;
        const double t = waves * (std::max(std::max(t_mem, t_mma), s > 1 ? 0.0 : t_epi) + t_red + kTileFixed) + t_epi;
        cands.push_back({t, cand, s, false});
        if (t < best * (s > 1 ? 0.9 : 1.0)) best = t, bn = cand, splits = s, pair = false;
      }
    }
    PN_REQUIRE(best < 1e29, name + ": no valid tile configuration");
  }
  // weights / scale / bias are padded to a multiple of the widest tile, so every launch configuration can use them
  const int cout_pad = round_up(sp.Cout, 256);

  // Pack weights [cout_pad][taps][cin_pad] in the activation dtype; the per-channel scale (folded BatchNorm gamma / sigma)
  // is multiplied in BEFORE the rounding to bf16 / tf32, so the epilogue only adds the bias.
  // + one extra K block per output channel: (bias_hi, bias_lo, 0, ...), used when the bias rides through the tensor core
  const long long ktot_w = ktot + block_k;
  std::vector<float> wp(static_cast<size_t>(cout_pad) * ktot_w, 0.f);
  for (int co = 0; co < sp.Cout; ++co)
    for (int ci = 0; ci < sp.Cin; ++ci)
      for (int r = 0; r < sp.R; ++r)
        for (int s = 0; s < sp.S; ++s)
          wp[static_cast<size_t>(co) * ktot_w + static_cast<size_t>(r * sp.S + s) * cin_pad + ci] =
              weight[((static_cast<size_t>(co) * sp.Cin + ci) * sp.R + r) * sp.S + s] * (scale ? scale[co] : 1.f);
  for (int co = 0; co < sp.Cout && bias; ++co) {  // hi + lo in the storage precision: exact to ~2^-17 (bf16) / 2^-21 (tf32)
    const float b = bias[co];
    float hi, lo;
    if (dt == kBF16) {
      const uint32_t h = static_cast<uint32_t>(f32_to_bf16(b)) << 16;
      std::memcpy(&hi, &h, 4);
      lo = b - hi;
    } else {
      uint32_t u;
      std::memcpy(&u, &b, 4);
      if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
      std::memcpy(&hi, &u, 4);
      lo = b - hi;
    }
    wp[static_cast<size_t>(co) * ktot_w + ktot] = hi;
    wp[static_cast<size_t>(co) * ktot_w + ktot + 1] = lo;
  }
  void* w_dev = nullptr;
  if (dt == kBF16) {
    std::vector<uint16_t> wb(wp.size());
    for (size_t i = 0; i < wp.size(); ++i) wb[i] = f32_to_bf16(wp[i]);
    w_dev = net.arena.upload(wb);
  } else {
    // tf32 path: round weights to the 10-bit mantissa (nearest, ties away) the tensor core would otherwise truncate to
    for (float& v : wp) {
      uint32_t u;
      std::memcpy(&u, &v, 4);
      if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
      std::memcpy(&v, &u, 4);
    }
    w_dev = net.arena.upload(wp);
  }
  std::vector<float> bi(cout_pad, 0.f);
  for (int i = 0; i < sp.Cout; ++i) bi[i] = bias ? bias[i] : 0.f;
  const float* bi_dev = net.arena.upload(bi);

  struct Variant {
    ConvMaps tm;
    ConvParams p;
    int grid = 0, bn = 0;
    size_t smem = 0;
    bool pair = false;
  };
  // Everything that depends on the launch configuration (tile width, split-K factor, CTA pair): tensor maps, kernel
  // parameters, shared-memory carve-up, grid.
  auto make_variant = [&](int bn, int splits, bool pair, int opt = 0) -> Variant {
    const bool a_tiled = (taps == 1 && sp.stride == 1 && stride_w == 1 && sp.pad == 0 && pad_w == 0);
    ConvMaps tm;
    const int n_tiles = round_up(sp.Cout, bn) / bn;
    if (a_tiled) {
      tm.a = encode_tiled_2d(dt, in.ptr, cin_pad, M, static_cast<uint64_t>(in.ld) * es, block_k, kBlockM, sw);
    } else {
      tm.a = encode_im2col(dt, in, cin_pad, sp.pad, pad_w, sp.dil, sp.R, sp.S, sp.stride, stride_w, block_k, kBlockM, sw);
    }
    tm.b = encode_tiled_2d(dt, w_dev, ktot_w, cout_pad, static_cast<uint64_t>(ktot_w) * es, block_k, pair ? bn / 2 : bn, sw);
    tm.out = tm.b;
    tm.res = tm.b;

    ConvParams p;
    std::memset(&p, 0, sizeof(p));
    p.M = static_cast<int>(M);
    p.Ho = Ho, p.Wo = Wo, p.R = sp.R, p.S = sp.S;
    p.stride = sp.stride, p.dil = sp.dil, p.pad = sp.pad;
    p.stride_w = stride_w, p.pad_w = pad_w;
    p.kb_per_tap = kb_per_tap, p.block_k = block_k, p.sw = sw;
    p.m_tiles = pair ? (m_tiles + 1) / 2 : m_tiles, p.n_tiles = n_tiles;  // work items along M: tiles, or 256-row tile pairs
    p.cout_store = std::min(round_up(sp.Cout, 8), out.C);
    PN_REQUIRE(p.cout_store >= sp.Cout, name + ": output view too narrow");
    p.a_tiled = a_tiled ? 1 : 0;
    p.bias = bi_dev;
    p.residual = residual ? residual->ptr : nullptr;
    p.ldr = residual ? residual->ld : 0;
    if (residual) PN_REQUIRE(residual->dt == dt && residual->pixels() == M, name + ": residual mismatch");
    p.out = out.ptr, p.ldc = out.ld;
    p.relu = sp.relu ? 1 : 0;
    p.out_fp32 = (out.dt == kF32) ? 1 : 0;
    p.round_tf32 = (dt == kF32 && !sp.out_fp32 && !net.x3) ? 1 : 0;
    p.dbg = sp.dbg;
    p.dbg_skip = sp.dbg_skip;
    p.m_limit = sp.m_limit;
    p.m_limit_rows = sp.m_limit_rows;
    p.splits = splits;
    p.kb_per_split = (taps * kb_per_tap + splits - 1) / splits;
    // Epilogue mode: smem-staged TMA stores (+ TMA residual prefetch) whenever a stored row chunk is at
    // least 64 bytes and the output has the activation dtype; otherwise direct per-thread stores.
    size_t epi_bytes = 0;
    {
      const int store_bytes = p.cout_store * es;
      int cb = 0;
      if (out.dt == dt && !sp.force_direct_epilogue && splits == 1) {
        if (store_bytes >= 128 && bn * es >= 128 && !((opt & 1) && es == 2)) cb = 128;
        else if (store_bytes >= 64 && bn * es >= 64 && es == 2) cb = 64;
      }
      if (cb) {
        p.epi_tma = 1;
        p.cb = cb;
        p.res_bufs = residual ? ((opt & 1) ? 2 : 3) : 0;
        tm.out = encode_tiled_2d(dt, out.ptr, p.cout_store, M, static_cast<uint64_t>(out.ld) * es, cb / es, 32, cb);  // one box per epilogue warp
        if (residual)
          tm.res = encode_tiled_2d(dt, residual->ptr, p.cout_store, M, static_cast<uint64_t>(residual->ld) * es, cb / es,
                                   kBlockM, cb);
        p.out_bufs = (opt & 1) ? 1 : 2;
        epi_bytes = static_cast<size_t>(p.out_bufs + p.res_bufs) * kBlockM * cb;
      }
    }

    // worth it where the epilogue is exposed (one tile per CTA: nothing overlaps it); persistent multi-tile launches hide
    // the epilogue behind the next tile's main loop and would only pay for the extra K block
    const bool one_tile_per_cta = static_cast<long long>(pair ? (m_tiles + 1) / 2 : m_tiles) * n_tiles <= (pair ? net.num_sms / 2 : net.num_sms);
    // option 4: weights resident (conv_umma.cuh, ConvParams::b_resident): the n tile's whole weight slab sits beside the ring,
    // the ring carries activations only; needs persistent multi-tile CTAs on ONE n tile each (grid a multiple of n_tiles)
    const bool b_res = (opt & 4) != 0;
    if (b_res) PN_REQUIRE(!pair && splits == 1 && !one_tile_per_cta, name + ": resident weights need single-CTA tiles, no split-K, several tiles per CTA");
    p.b_resident = b_res ? 1 : 0;
    p.bias_block = (!b_res && p.epi_tma && splits == 1 && bias != nullptr && (one_tile_per_cta || (opt & 2))) ? 1 : 0;
    const int kblocks = taps * kb_per_tap;
    const size_t stage_bytes = b_res ? static_cast<size_t>(kBlockM) * sw : static_cast<size_t>(kBlockM + (pair ? bn / 2 : bn)) * sw;
    size_t fixed_bytes = 1024 /*align*/ + epi_bytes + 4 * bn * sizeof(float) /*bias per epilogue warp*/ + kBlockM * 32 /*ones tile*/ + 256 /*barriers*/;
    if (b_res) fixed_bytes += static_cast<size_t>(kblocks) * bn * sw;
    PN_REQUIRE(fixed_bytes + 2 * stage_bytes <= 227 * 1024, name + ": does not fit shared memory");
    size_t budget = 227 * 1024 - fixed_bytes;
    int stages = static_cast<int>(budget / stage_bytes);
    if (p.epi_tma && kblocks >= 32) {
      // K-heavy layer: the epilogue is a small fraction of the tile, so trade the second output staging
      // buffer for a deeper operand pipeline when that buys a stage.
      const size_t fixed1 = fixed_bytes - static_cast<size_t>(kBlockM) * p.cb;
      const int stages1 = static_cast<int>((227 * 1024 - fixed1) / stage_bytes);
      if (stages1 > stages && stages < 6) {
        p.out_bufs = 1;
        fixed_bytes = fixed1;
        budget = 227 * 1024 - fixed_bytes;
        stages = stages1;
      }
    }
    // Single-wave layers (one tile per CTA) are latency-bound: cap the footprint at ~half an SM so that the next
    // kernel's CTA can become resident beside this one (programmatic dependent launch) and overlap its prologue.
    if (static_cast<long long>(m_tiles) * n_tiles * splits <= net.num_sms && !pair) {
      const int cap = static_cast<int>((110 * 1024 - std::min<size_t>(fixed_bytes, 100 * 1024)) / stage_bytes);
      if (cap >= 3) stages = std::min(stages, cap);
    }
    if (opt & 1) {
      // Two CTAs per SM: where the epilogue (one warp per SM sub-partition, issue-bound) outlasts the main loop, a second
      // resident CTA doubles the epilogue warps and runs its main loop under the first one's epilogue.  Each CTA gets half
      // the shared memory (and its 2 x BN accumulator columns must fit tensor memory twice: BN <= 128).
      // With CTA pairs (round 2, last experiment): two clusters of two CTAs per SM pair - each CTA still owns 2 x BN columns.
      PN_REQUIRE(splits == 1 && bn <= 128 && p.epi_tma, name + ": two CTAs per SM need an unsplit tile of at most 128 columns");
      const size_t half = 113 * 1024;
      PN_REQUIRE(fixed_bytes + 2 * stage_bytes <= half, name + ": two CTAs per SM do not fit shared memory");
      stages = std::min(stages, static_cast<int>((half - fixed_bytes) / stage_bytes));
    }
    if (stages > 8) stages = 8;
    if (stages > p.kb_per_split + 1) stages = std::max(2, p.kb_per_split + 1);
    if (splits > 1) {  // the operand ring doubles as the parking area of the fp32 partial tile
      const int need = static_cast<int>((static_cast<size_t>(kBlockM) * bn * 4 + stage_bytes - 1) / stage_bytes);
      stages = std::max(stages, need);
      PN_REQUIRE(fixed_bytes + stages * stage_bytes <= 227 * 1024, name + ": split-K partial tile does not fit");
    }
    p.stages = stages;
    PN_REQUIRE(stages >= 2, name + ": shared memory budget too small for a 2-stage pipeline");
    const size_t smem = fixed_bytes + stages * stage_bytes;
    const long long tiles = static_cast<long long>(p.m_tiles) * n_tiles * splits;
    int grid = pair ? 2 * static_cast<int>(splits > 1 ? tiles : std::min<long long>(tiles, ((opt & 1) ? 2 : 1) * (net.num_sms / 2)))
                    : static_cast<int>(splits > 1 ? tiles : std::min<long long>(tiles, ((opt & 1) ? 2 : 1) * net.num_sms));
    if (b_res) grid = std::max(n_tiles, grid / n_tiles * n_tiles);   // work strides by the grid: the n tile of a CTA is then fixed

    Variant v;
    v.tm = tm, v.p = p, v.grid = grid, v.bn = bn, v.smem = smem, v.pair = pair;
    return v;
  };
  auto launch_variant = [dt](const Variant& v, cudaStream_t s) {
    if (dt == kBF16)
      launch_conv_bn<__nv_bfloat16>(v.bn, v.pair, v.tm, v.p, v.grid, v.smem, s);
    else
      launch_conv_bn<float>(v.bn, v.pair, v.tm, v.p, v.grid, v.smem, s);
  };

  // Measured choice: the tile model ranks the configurations, the best few are timed on the real buffers (back-to-back
  // launches, as in the graph) and the fastest wins; the result is cached per layer shape.  PN_CONV_AUTOTUNE=0 keeps
  // the model's choice; test hooks (force_*) bypass both.
  int opt = sp.force_opt;
  const bool forced = sp.force_opt || sp.force_bn || sp.force_splits || sp.force_pair || sp.force_direct_epilogue || sp.dbg;
  // (the split-precision parity mode never times: the tile model's choice, identical in every process, so its results are
  //  reproducible run to run; its speed is not the point)
  if (!forced && autotune_mode() != 0 && cands.size() > 1 && !sp.x3_cin) {
    // The key holds everything the candidate set and the timings depend on (incl. padding, the dynamic-row-limit flag, the
    // device's SM count and the opt-in modes), so an entry is only ever reused for an identical launch problem.
    char key[320];
    std::snprintf(key, sizeof(key), "%d|%lld|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d|%d", static_cast<int>(dt), M, Wo,
                  cin_pad, sp.Cout, sp.R, sp.S, sp.stride, stride_w, sp.dil, residual ? 1 : 0, out.dt == dt ? 0 : 1,
                  sp.relu ? 1 : 0, sp.no_split ? 1 : 0, std::min(round_up(sp.Cout, 8), out.C), sp.pad, pad_w,
                  sp.m_limit ? sp.m_limit_rows : 0, net.num_sms, tune_extra());
    std::lock_guard<std::mutex> lock(g_tune_mu);  // one tuner at a time (it owns the device while it measures)
    auto& table = tune_table();
    auto it = table.find(key);
    bool usable = false;
    if (it != table.end()) {  // an imported entry must name one of this layer's valid configurations
      const TuneEntry& e = it->second;
      for (const Cand& c : cands) usable |= (c.bn == e.bn && c.splits == e.splits && c.pair == (e.pair != 0));
      if (e.opt & 1) usable = usable && !e.pair && e.splits == 1 && e.bn <= 128;
    }
    if (!usable && autotune_mode() == 2) {
      std::sort(cands.begin(), cands.end(), [](const Cand& a, const Cand& b) { return a.t < b.t; });
      std::vector<Cand> shortlist(cands.begin(), cands.begin() + std::min<size_t>(cands.size(), 6));
      bool has_model = false;
      for (const Cand& c : shortlist) has_model |= (c.bn == bn && c.splits == splits && c.pair == pair);
      if (!has_model) shortlist.push_back({0.0, bn, splits, pair});
      for (const Cand& c : cands) {  // the model is least certain about split-K over pairs: always time those
        bool listed = false;
        for (const Cand& l : shortlist) listed |= (l.bn == c.bn && l.splits == c.splits && l.pair == c.pair);
        if (c.pair && c.splits > 1 && !listed) shortlist.push_back(c);
      }
      {  // launch-shape options for persistent multi-tile launches of the best-ranked single-CTA tiles
        std::vector<Cand> extra;
        for (const Cand& c : shortlist) {
          if (c.splits != 1 || c.opt) continue;
          const long long tiles_c = static_cast<long long>(c.pair ? (m_tiles + 1) / 2 : m_tiles) * (round_up(sp.Cout, c.bn) / c.bn);
          if (tiles_c <= (c.pair ? net.num_sms / 2 : net.num_sms)) continue;
          // (option 2, the bias block on multi-tile launches, measured neutral: not enumerated)
          if (!c.pair && c.bn <= 128 && (tune_extra() & 2)) extra.push_back({c.t, c.bn, 1, false, 1});
        }
        shortlist.insert(shortlist.end(), extra.begin(), extra.end());
      }
      Cand best_c{1e30, bn, splits, pair};
      for (const Cand& c : shortlist) {
        float ms = 0.f;
        static const bool log = std::getenv("PN_CONV_TUNE_LOG") != nullptr;  // one line per timed configuration (stderr)
        try {
          Variant v = make_variant(c.bn, c.splits, c.pair, c.opt);
          v.p.m_limit = nullptr;  // time the full-capacity launch
          char what[320];
          std::snprintf(what, sizeof(what), "%s M=%lld tile %d splits %d pair %d opt %d grid %d smem %zu stages %d", name.c_str(), M, c.bn,
                        c.splits, c.pair ? 1 : 0, c.opt, v.grid, v.smem, v.p.stages);
          if (log) {
            std::fprintf(stderr, "[tune] %s ...", what);
            std::fflush(stderr);
          }
          ms = time_launches([&](cudaStream_t s) { launch_variant(v, s); }, what);
          if (log) std::fprintf(stderr, " %.1f us\n", ms * 1e3f);
        } catch (const std::exception& e) {
          if (log) std::fprintf(stderr, " skipped (%s)\n", e.what());
          cudaGetLastError();  // clears a launch-configuration error; anything sticky (a faulted kernel) is fatal
          const cudaError_t st = cudaDeviceSynchronize();
          PN_REQUIRE(st == cudaSuccess, name + ": CUDA error while timing a launch configuration: " + cudaGetErrorString(st));
          continue;
        }
        if (ms < best_c.t * (c.opt ? 0.97 : 1.0)) best_c = {ms, c.bn, c.splits, c.pair, c.opt};  // options must win clearly
      }
      TuneEntry e;
      e.bn = best_c.bn, e.splits = best_c.splits, e.pair = best_c.pair ? 1 : 0, e.opt = best_c.opt;
      table[key] = e;
      it = table.find(key);
      usable = true;
    }
    if (usable) bn = it->second.bn, splits = it->second.splits, pair = it->second.pair != 0, opt = it->second.opt;
  }
  // Diagnosis: PN_CONV_FORCE_OPT1=all|<substring of the layer name> runs every eligible layer (persistent multi-tile launch,
  // single-CTA tile of at most 128 columns, TMA epilogue) with two CTAs per SM, whatever the table / tuner chose.
  if (!forced) {
    static const char* force1 = std::getenv("PN_CONV_FORCE_OPT1");
    if (force1 && force1[0] && (std::strcmp(force1, "all") == 0 || name.find(force1) != std::string::npos)) {
      int b1 = std::min(bn, 128);
      while (b1 >= 32 && (b1 > cout32 || cout32 % b1 != 0)) b1 /= 2;
      const long long tiles1 = static_cast<long long>(m_tiles) * (round_up(sp.Cout, std::max(b1, 32)) / std::max(b1, 32));
      const int es1 = es;
      const bool epi_ok = out.dt == dt && std::min(round_up(sp.Cout, 8), out.C) * es1 >= 64 && b1 * es1 >= 64;
      if (b1 >= 32 && tiles1 > net.num_sms && epi_ok && sp.m_limit == nullptr) {
        try {
          (void)make_variant(b1, 1, false, 1);
          bn = b1, splits = 1, pair = false, opt = 1;
          if (std::getenv("PN_CONV_TUNE_LOG")) std::fprintf(stderr, "[force-opt1] %s M=%lld tile %d\n", name.c_str(), M, b1);
        } catch (const std::exception&) {
        }
      }
    }
  }
  net.last_bn = bn + 1000 * splits + (pair ? 100000 : 0) + 1000000 * opt;
  const Variant chosen = make_variant(bn, splits, pair, opt);

  const double flops = 2.0 * static_cast<double>(M) * sp.Cout * (sp.x3_cin ? sp.x3_cin : sp.Cin) * taps;  // algorithmic, also on the split-precision path
  net.add(name, [=](cudaStream_t s) { launch_variant(chosen, s); }, flops);
  net.launches_per_forward += 1;
}

std::string conv_tuning_export() {
  std::lock_guard<std::mutex> lock(g_tune_mu);
  std::ostringstream os;
  for (const auto& kv : tune_table())
    os << kv.first << ' ' << kv.second.bn << ' ' << kv.second.splits << ' ' << kv.second.pair << ' ' << kv.second.opt << '\n';
  return os.str();
}

int conv_tuning_import(const std::string& text) {
  std::lock_guard<std::mutex> lock(g_tune_mu);
  std::istringstream is(text);
  std::string line;
  int n = 0;
  while (std::getline(is, line)) {
    if (line.empty() || line[0] == '#') continue;
    std::istringstream ls(line);
    std::string key;
    TuneEntry e;
    if (!(ls >> key >> e.bn >> e.splits >> e.pair >> e.opt)) throw std::runtime_error("peanut_b200: malformed tuning line: " + line);
    PN_REQUIRE((e.bn == 32 || e.bn == 64 || e.bn == 128 || e.bn == 256) && e.splits >= 1 && e.splits <= 8, "bad tuning entry: " + line);
    tune_table()[key] = e;
    ++n;
  }
  return n;
}

void conv_tuning_clear() {
  std::lock_guard<std::mutex> lock(g_tune_mu);
  tune_table().clear();
}

}  // namespace pn
