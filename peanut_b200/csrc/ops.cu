// HBM-bound helper kernels around the tensor-core convolutions: layout conversion, max pooling,
// pyramid pooling (AdaptiveAvgPool2d), bilinear resize into a channel slice and the final
// logits -> full-resolution NCHW (+ sigmoid) pass.  All operate on NHWC views with 16-byte vector access.
#include <chrono>
#include <cstdio>
#include "engine.h"
#include "vec.cuh"

namespace pn {

// ---------------------------------------------------------------------------------------------
// fp32 NCHW (caller memory, pointer read from a device slot) -> NHWC dt with zero channel padding.
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* const* src_slot, T* dst, int B, int C, int HW, int Cpad, int round_tf32) {
  pdl_grid_sync();
  const float* src = *src_slot;
  const long long pix = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
  if (pix >= static_cast<long long>(B) * HW) return;
  const int b = static_cast<int>(pix / HW);
  const int hw = static_cast<int>(pix - static_cast<long long>(b) * HW);
  T* o = dst + pix * Cpad;
  for (int c0 = 0; c0 < Cpad; c0 += 8) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      v[j] = c < C ? __ldg(src + (static_cast<long long>(b) * C + c) * HW + hw) : 0.f;
      if (sizeof(T) == 4 && round_tf32) {  // fp32 activations feed tf32 MMAs: round-to-nearest here, truncation there is then exact
        uint32_t r;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v[j]));
        v[j] = __uint_as_float(r);
      }
    }
    store8(o + c0, v);
  }
}

void add_nchw_to_nhwc(Net& net, const float* const* src_slot, const Tensor& out, int C) {
  PN_REQUIRE(out.ld == out.C && out.C % 8 == 0, "nchw_to_nhwc: output must be dense");
  const long long pixels = out.pixels();
  const int HW = out.H * out.W;
  const int threads = 256;
  const int blocks = static_cast<int>((pixels + threads - 1) / threads);
  Tensor o = out;
  const int rnd = net.x3 ? 0 : 1;  // fp32 parity mode keeps every stored bit
  net.add("nchw_to_nhwc", [=](cudaStream_t s) {
    if (o.dt == kBF16)
      launch_pdl(nchw_to_nhwc_kernel<__nv_bfloat16>, blocks, threads, 0, s, src_slot, static_cast<__nv_bfloat16*>(o.ptr), o.B, C, HW, o.C, rnd);
    else
      launch_pdl(nchw_to_nhwc_kernel<float>, blocks, threads, 0, s, src_slot, static_cast<float*>(o.ptr), o.B, C, HW, o.C, rnd);
  });
  net.launches_per_forward += 1;
}

// ---------------------------------------------------------------------------------------------
// MaxPool2d(kernel 3, stride 2, padding 1), NHWC, 8 channels per thread.  All nine taps are loaded before the first maximum
// (a tap outside the image re-reads the window's centre pixel, which is always inside: the maximum does not change and no load
// is predicated); bf16 maxima are taken on the packed pairs (the maximum of bf16 values is one of them: exact).
__device__ __forceinline__ uint4 max8_raw(const uint4& a, const uint4& b, const __nv_bfloat16*) {
  uint4 r;
  const __nv_bfloat162 x0 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.x), *reinterpret_cast<const __nv_bfloat162*>(&b.x));
  const __nv_bfloat162 x1 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.y), *reinterpret_cast<const __nv_bfloat162*>(&b.y));
  const __nv_bfloat162 x2 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.z), *reinterpret_cast<const __nv_bfloat162*>(&b.z));
  const __nv_bfloat162 x3 = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.w), *reinterpret_cast<const __nv_bfloat162*>(&b.w));
  r.x = *reinterpret_cast<const uint32_t*>(&x0), r.y = *reinterpret_cast<const uint32_t*>(&x1);
  r.z = *reinterpret_cast<const uint32_t*>(&x2), r.w = *reinterpret_cast<const uint32_t*>(&x3);
  return r;
}
template <typename T>
struct PoolVec;
template <>
struct PoolVec<__nv_bfloat16> {   // 8 channels = one 16-byte load
  uint4 v;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { v = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void max_with(const PoolVec& o) { v = max8_raw(v, o.v, static_cast<const __nv_bfloat16*>(nullptr)); }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const { *reinterpret_cast<uint4*>(p) = v; }
};
template <>
struct PoolVec<float> {           // 8 channels = two 16-byte loads
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void max_with(const PoolVec& o) {
    a.x = fmaxf(a.x, o.a.x), a.y = fmaxf(a.y, o.a.y), a.z = fmaxf(a.z, o.a.z), a.w = fmaxf(a.w, o.a.w);
    b.x = fmaxf(b.x, o.b.x), b.y = fmaxf(b.y, o.b.y), b.z = fmaxf(b.z, o.b.z), b.w = fmaxf(b.w, o.b.w);
  }
  __device__ __forceinline__ void store(float* p) const {
    *reinterpret_cast<float4*>(p) = a;
    *reinterpret_cast<float4*>(p + 4) = b;
  }
};

template <typename T>
__global__ void maxpool3x3s2_kernel(const T* in, long long ldi, T* out, long long ldo, int B, int H, int W, int Ho,
                                    int Wo, int C8) {
  pdl_grid_sync();
  Idx4 ix;
  if (!split_index(Ho, Wo, C8, static_cast<uint32_t>(B) * Ho * Wo * C8, ix)) return;
  const int b = static_cast<int>(ix.a), y = static_cast<int>(ix.b), x = static_cast<int>(ix.c), cg = static_cast<int>(ix.d);
  const T* base = in + static_cast<long long>(b) * H * W * ldi + cg * 8;
  PoolVec<T> v[9];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    int yy = 2 * y - 1 + dy;
    if (yy < 0 || yy >= H) yy = 2 * y;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      int xx = 2 * x - 1 + dx;
      if (xx < 0 || xx >= W) xx = 2 * x;
      v[dy * 3 + dx].load(base + (static_cast<long long>(yy) * W + xx) * ldi);
    }
  }
#pragma unroll
  for (int k = 1; k < 9; ++k) v[0].max_with(v[k]);
  v[0].store(out + ((static_cast<long long>(b) * Ho + y) * Wo + x) * ldo + cg * 8);
}

void add_maxpool3x3s2(Net& net, const Tensor& in, const Tensor& out) {
  PN_REQUIRE(out.H == conv_out(in.H, 3, 2, 1, 1) && out.W == conv_out(in.W, 3, 2, 1, 1), "maxpool shape");
  PN_REQUIRE(in.C % 8 == 0 && out.C == in.C && in.dt == out.dt, "maxpool channels");
  const int C8 = in.C / 8;
  const long long total = out.pixels() * C8;
  check_u32_launch(total, "maxpool3x3s2");
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads);
  Tensor i = in, o = out;
  net.add("maxpool3x3s2", [=](cudaStream_t s) {
    if (i.dt == kBF16)
      launch_pdl(maxpool3x3s2_kernel<__nv_bfloat16>, blocks, threads, 0, s, static_cast<const __nv_bfloat16*>(i.ptr), i.ld, static_cast<__nv_bfloat16*>(o.ptr), o.ld, i.B, i.H, i.W, o.H, o.W, C8);
    else
      launch_pdl(maxpool3x3s2_kernel<float>, blocks, threads, 0, s, static_cast<const float*>(i.ptr), i.ld, static_cast<float*>(o.ptr), o.ld, i.B, i.H, i.W, o.H, o.W, C8);
  });
  net.launches_per_forward += 1;
}

// ---------------------------------------------------------------------------------------------
// Pyramid pooling: AdaptiveAvgPool2d(s) for every s in `scales` in two passes.
//   pass 1: per image row, sum each x-bin            -> rowpart[b][y][bin][c]   (fp32)
//   pass 2: per output bin, sum the rows of the bin  -> out_s[b][by][bx][c]     (dt)
// Bin i of scale s covers [floor(i*L/s), ceil((i+1)*L/s)) as in ATen's adaptive pooling.
struct PpmMeta {
  int nscales;
  int scale[4];
  int bin_off[5];  // prefix sum of scales: x-bin index base per scale
};

__device__ __forceinline__ int bin_start(int i, int L, int s) { return (i * L) / s; }
__device__ __forceinline__ int bin_end(int i, int L, int s) { return ((i + 1) * L + s - 1) / s; }

// One thread per (image row, 8 channels): 16 / 32-byte loads, the bin's cells summed in x order (fp32).
template <typename T>
__global__ void ppm_rows_kernel(const T* in, long long ldi, float* rowpart, int B, int H, int W, int C, PpmMeta meta) {
  pdl_grid_sync();
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  const int row = blockIdx.y;  // b*H + y
  if (c >= C) return;
  const T* p = in + static_cast<long long>(row) * W * ldi + c;
  const int nb = meta.bin_off[meta.nscales];
  float* o = rowpart + static_cast<long long>(row) * nb * C + c;
  for (int si = 0; si < meta.nscales; ++si) {
    const int s = meta.scale[si];
    for (int xb = 0; xb < s; ++xb) {
      float acc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = 0.f;
      const int x1 = bin_end(xb, W, s);
      int x = bin_start(xb, W, s);
      for (; x + 4 <= x1; x += 4) {   // four cells in flight; the adds stay in x order
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) load8(p + static_cast<long long>(x + u) * ldi, v[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += v[u][j];
        }
      }
      for (; x < x1; ++x) {
        float v[8];
        load8(p + static_cast<long long>(x) * ldi, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += v[j];
      }
      float* dst = o + static_cast<long long>(meta.bin_off[si] + xb) * C;
      *reinterpret_cast<float4*>(dst) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
  }
}

template <typename T>
__global__ void ppm_bins_kernel(const float* rowpart, T* out, int H, int W, int C, int s, int bin_off, int nb) {
  pdl_grid_sync();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int bx = blockIdx.y % s;
  const int by = blockIdx.y / s;
  const int b = blockIdx.z;
  const int y0 = bin_start(by, H, s), y1 = bin_end(by, H, s);
  const int x0 = bin_start(bx, W, s), x1 = bin_end(bx, W, s);
  float acc = 0.f;
  for (int y = y0; y < y1; ++y) acc += rowpart[((static_cast<long long>(b) * H + y) * nb + bin_off + bx) * C + c];
  acc /= static_cast<float>((y1 - y0) * (x1 - x0));
  out[((static_cast<long long>(b) * s + by) * s + bx) * C + c] = from_float<T>(acc);
}

void add_ppm_pool(Net& net, const Tensor& in, const std::vector<int>& scales, const std::vector<Tensor>& outs) {
  PN_REQUIRE(scales.size() <= 4 && scales.size() == outs.size(), "ppm: at most 4 scales");
  PN_REQUIRE(in.C % 8 == 0 && in.ld % 8 == 0, "ppm: channel count must be a multiple of 8");
  PpmMeta meta{};
  meta.nscales = static_cast<int>(scales.size());
  int nb = 0;
  for (size_t i = 0; i < scales.size(); ++i) {
    meta.scale[i] = scales[i];
    meta.bin_off[i] = nb;
    nb += scales[i];
    PN_REQUIRE(outs[i].H == scales[i] && outs[i].W == scales[i] && outs[i].ld == in.C && outs[i].dt == in.dt, "ppm: out shape");
  }
  meta.bin_off[scales.size()] = nb;
  float* rowpart = static_cast<float*>(net.arena.alloc(static_cast<size_t>(in.B) * in.H * nb * in.C * sizeof(float)));
  Tensor i = in;
  std::vector<Tensor> o = outs;
  net.add("ppm_pool", [=](cudaStream_t s) {
    const int threads = 128;
    dim3 g1((i.C / 8 + threads - 1) / threads, i.B * i.H);
    if (i.dt == kBF16)
      launch_pdl(ppm_rows_kernel<__nv_bfloat16>, g1, threads, 0, s, static_cast<const __nv_bfloat16*>(i.ptr), i.ld, rowpart, i.B, i.H, i.W, i.C, meta);
    else
      launch_pdl(ppm_rows_kernel<float>, g1, threads, 0, s, static_cast<const float*>(i.ptr), i.ld, rowpart, i.B, i.H, i.W, i.C, meta);
    for (int k = 0; k < meta.nscales; ++k) {
      dim3 g2((i.C + threads - 1) / threads, meta.scale[k] * meta.scale[k], i.B);
      if (i.dt == kBF16)
        launch_pdl(ppm_bins_kernel<__nv_bfloat16>, g2, threads, 0, s, rowpart, static_cast<__nv_bfloat16*>(o[k].ptr), i.H, i.W, i.C, meta.scale[k], meta.bin_off[k], nb);
      else
        launch_pdl(ppm_bins_kernel<float>, g2, threads, 0, s, rowpart, static_cast<float*>(o[k].ptr), i.H, i.W, i.C, meta.scale[k], meta.bin_off[k], nb);
    }
  });
  net.launches_per_forward += 1 + static_cast<long long>(scales.size());
}

// ---------------------------------------------------------------------------------------------
// F.interpolate(mode='bilinear', align_corners=False) from a small NHWC map into a channel slice
// of a larger NHWC tensor (the PSP concat buffer).  Source index rule follows ATen's
// area_pixel_compute_source_index: src = scale*(dst+0.5)-0.5 clamped at 0, scale = in/out.
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * (static_cast<float>(dst) + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = static_cast<int>(src);
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - static_cast<float>(i0);
  l0 = 1.f - l1;
}

template <typename T>
__global__ void bilinear_into_kernel(const T* in, long long ldi, int h, int w, T* out, long long ldo, int B, int H, int W,
                                     int C8, float sy, float sx) {
  pdl_grid_sync();
  Idx4 ix;
  if (!split_index(H, W, C8, static_cast<uint32_t>(B) * H * W * C8, ix)) return;
  const int b = static_cast<int>(ix.a), y = static_cast<int>(ix.b), x = static_cast<int>(ix.c), cg = static_cast<int>(ix.d);
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bilinear_src(y, sy, h, y0, y1, ly0, ly1);
  bilinear_src(x, sx, w, x0, x1, lx0, lx1);
  const T* base = in + static_cast<long long>(b) * h * w * ldi + cg * 8;
  float v00[8], v01[8], v10[8], v11[8], r[8];
  load8(base + (static_cast<long long>(y0) * w + x0) * ldi, v00);
  load8(base + (static_cast<long long>(y0) * w + x1) * ldi, v01);
  load8(base + (static_cast<long long>(y1) * w + x0) * ldi, v10);
  load8(base + (static_cast<long long>(y1) * w + x1) * ldi, v11);
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = ly0 * (lx0 * v00[j] + lx1 * v01[j]) + ly1 * (lx0 * v10[j] + lx1 * v11[j]);
  store8(out + ((static_cast<long long>(b) * H + y) * W + x) * ldo + cg * 8, r);
}

void add_bilinear_into(Net& net, const Tensor& in, const Tensor& out) {
  PN_REQUIRE(in.C == out.C && in.C % 8 == 0 && in.dt == out.dt && in.B == out.B, "bilinear_into: shape");
  const int C8 = in.C / 8;
  const long long total = out.pixels() * C8;
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads);
  const float sy = static_cast<float>(in.H) / static_cast<float>(out.H);
  const float sx = static_cast<float>(in.W) / static_cast<float>(out.W);
  Tensor i = in, o = out;
  net.add("bilinear_into", [=](cudaStream_t s) {
    if (i.dt == kBF16)
      launch_pdl(bilinear_into_kernel<__nv_bfloat16>, blocks, threads, 0, s, static_cast<const __nv_bfloat16*>(i.ptr), i.ld, i.H, i.W, static_cast<__nv_bfloat16*>(o.ptr), o.ld, o.B, o.H, o.W, C8, sy, sx);
    else
      launch_pdl(bilinear_into_kernel<float>, blocks, threads, 0, s, static_cast<const float*>(i.ptr), i.ld, i.H, i.W, static_cast<float*>(o.ptr), o.ld, o.B, o.H, o.W, C8, sy, sx);
  });
  net.launches_per_forward += 1;
}

// ---------------------------------------------------------------------------------------------
// Final resize of the class logits (mmseg `resize(..., mode='bilinear', align_corners=False)`,
// encoder_decoder.py:75-79) fused with the host-side expit of nav/agent/prediction.py:158:
// NHWC fp32 [B,h,w,ld] -> NCHW fp32 [B,C,H,W] in caller memory (pointer read from a device slot).
__global__ void upsample_logits_kernel(const float* in, long long ldi, int h, int w, int C, float* const* dst_slot,
                                       const int* sigmoid_slot, int B, int H, int W, float sy, float sx) {
  pdl_grid_sync();
  float* dst = *dst_slot;
  const int apply_sigmoid = *sigmoid_slot;
  Idx4 ix;
  if (!split_index(H, W, 1, static_cast<uint32_t>(B) * H * W, ix)) return;
  const int b = static_cast<int>(ix.a), y = static_cast<int>(ix.b), x = static_cast<int>(ix.c);
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  bilinear_src(y, sy, h, y0, y1, ly0, ly1);
  bilinear_src(x, sx, w, x0, x1, lx0, lx1);
  const float* base = in + static_cast<long long>(b) * h * w * ldi;
  const float* p00 = base + (static_cast<long long>(y0) * w + x0) * ldi;
  const float* p01 = base + (static_cast<long long>(y0) * w + x1) * ldi;
  const float* p10 = base + (static_cast<long long>(y1) * w + x0) * ldi;
  const float* p11 = base + (static_cast<long long>(y1) * w + x1) * ldi;
  for (int c = 0; c < C; ++c) {
    float r = ly0 * (lx0 * __ldg(p00 + c) + lx1 * __ldg(p01 + c)) + ly1 * (lx0 * __ldg(p10 + c) + lx1 * __ldg(p11 + c));
    if (apply_sigmoid) r = 1.f / (1.f + expf(-r));
    dst[((static_cast<long long>(b) * C + c) * H + y) * W + x] = r;
  }
}

void add_upsample_logits(Net& net, const Tensor& logits, int C, int Hout, int Wout, float* const* dst_slot,
                         const int* sigmoid_slot) {
  PN_REQUIRE(logits.dt == kF32 && logits.C >= C, "upsample_logits: fp32 logits expected");
  const long long total = static_cast<long long>(logits.B) * Hout * Wout;
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads);
  const float sy = static_cast<float>(logits.H) / static_cast<float>(Hout);
  const float sx = static_cast<float>(logits.W) / static_cast<float>(Wout);
  Tensor l = logits;
  net.add("upsample_logits", [=](cudaStream_t s) {
    launch_pdl(upsample_logits_kernel, blocks, threads, 0, s, static_cast<const float*>(l.ptr), l.ld, l.H, l.W, C, dst_slot, sigmoid_slot, l.B, Hout, Wout, sy, sx);
  });
  net.launches_per_forward += 1;
}

// ---------------------------------------------------------------------------------------------
Net::~Net() {
  if (graph_exec) cudaGraphExecDestroy(graph_exec);
  if (cap_stream) cudaStreamDestroy(cap_stream);
  if (fwd_done) cudaEventDestroy(fwd_done);
  for (int l = 0; l < kMaxLanes; ++l) {
    if (lane_stream[l]) cudaStreamDestroy(lane_stream[l]);
    if (lane_fork[l]) cudaEventDestroy(lane_fork[l]);
    if (lane_done[l]) cudaEventDestroy(lane_done[l]);
  }
}

// One pass over the launch list with its parallel lanes (see Net::set_lane): fork/join through events, so the same
// code serves eager execution and stream capture (the side streams join the capture through the event waits).
void Net::run_eager(cudaStream_t s) {
  long long main_ops = 0;                     // lane-0 ops issued so far
  long long lane_synced[kMaxLanes] = {0, 0, 0, 0};  // value of main_ops each lane last synchronised with (-1: never)
  bool lane_active[kMaxLanes] = {false, false, false, false};
  for (int l = 0; l < kMaxLanes; ++l) lane_synced[l] = -1;
  auto join_all = [&]() {
    for (int l = 1; l < kMaxLanes; ++l) {
      if (!lane_active[l]) continue;
      PN_CUDA_CHECK(cudaEventRecord(lane_done[l], lane_stream[l]));
      PN_CUDA_CHECK(cudaStreamWaitEvent(s, lane_done[l], 0));
      lane_active[l] = false;
      lane_synced[l] = -1;
    }
  };
  static const bool no_lanes = debug_flag("PN_DEBUG_NO_LANES");
  static const bool sync_each = debug_flag("PN_DEBUG_SYNC_EACH");
  for (size_t i = 0; i < ops.size(); ++i) {
    if (op_join[i]) join_all();
    const int lane = no_lanes ? 0 : op_lane[i];
    if (lane == 0) {
      ops[i](s);
      ++main_ops;
      if (sync_each) {  // diagnosis: which op does not finish (cannot be used while capturing)
        cudaEvent_t ev;
        PN_CUDA_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        PN_CUDA_CHECK(cudaEventRecord(ev, s));
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) {
          const cudaError_t q = cudaEventQuery(ev);
          if (q == cudaSuccess) break;
          if (q != cudaErrorNotReady) PN_CUDA_CHECK(q);
          if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(8)) {
            std::fprintf(stderr, "peanut_b200: op %zu '%s' did not finish within 8 s\n", i, op_names[i].c_str());
            std::fflush(stderr);
            std::_Exit(98);
          }
        }
        cudaEventDestroy(ev);
      }
      continue;
    }
    PN_REQUIRE(lane > 0 && lane < kMaxLanes, "bad lane");
    if (!lane_stream[lane]) {
      PN_CUDA_CHECK(cudaStreamCreateWithFlags(&lane_stream[lane], cudaStreamNonBlocking));
      PN_CUDA_CHECK(cudaEventCreateWithFlags(&lane_fork[lane], cudaEventDisableTiming));
      PN_CUDA_CHECK(cudaEventCreateWithFlags(&lane_done[lane], cudaEventDisableTiming));
    }
    if (lane_synced[lane] != main_ops) {  // see every lane-0 op that precedes this one
      PN_CUDA_CHECK(cudaEventRecord(lane_fork[lane], s));
      PN_CUDA_CHECK(cudaStreamWaitEvent(lane_stream[lane], lane_fork[lane], 0));
      lane_synced[lane] = main_ops;
    }
    ops[i](lane_stream[lane]);
    lane_active[lane] = true;
  }
  join_all();
}

// First call runs eagerly (lazy function-attribute setup, module load); the second captures the
// launch list into a CUDA graph that later calls replay - the per-layer launch overhead of
// ~60-150 small kernels would otherwise dominate at batch 1.
void Net::run(cudaStream_t s) {
  static const bool no_graph = debug_flag("PN_DEBUG_NO_GRAPH") || debug_flag("PN_DEBUG_SYNC_EACH");
  if (!use_graph || no_graph) {
    run_eager(s);
    return;
  }
  if (graph_exec) {
    PN_CUDA_CHECK(cudaGraphLaunch(graph_exec, s));
    return;
  }
  if (warm_runs < 1) {
    run_eager(s);
    ++warm_runs;
    return;
  }
  // Capture on a private stream (the caller's may be the legacy default stream, which cannot be
  // captured); the instantiated graph is then launched on whatever stream the caller passes.
  if (!cap_stream) PN_CUDA_CHECK(cudaStreamCreateWithFlags(&cap_stream, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  PN_CUDA_CHECK(cudaStreamBeginCapture(cap_stream, cudaStreamCaptureModeThreadLocal));
  try {
    run_eager(cap_stream);
  } catch (...) {
    cudaStreamEndCapture(cap_stream, &graph);
    if (graph) cudaGraphDestroy(graph);
    throw;
  }
  PN_CUDA_CHECK(cudaStreamEndCapture(cap_stream, &graph));
  PN_CUDA_CHECK(cudaGraphInstantiate(&graph_exec, graph, 0));
  cudaGraphDestroy(graph);
  PN_CUDA_CHECK(cudaGraphLaunch(graph_exec, s));
}

}  // namespace pn
