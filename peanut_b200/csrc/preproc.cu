// Frame pre-processing kernels either side of the networks (all HBM/latency-bound, integer or elementwise):
//
//   * Mask-RCNN input: detectron2's DefaultPredictor resizes the uint8 BGR frame with PIL bilinear
//     (ResizeShortestEdge, nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml:28,30), then
//     GeneralizedRCNN.preprocess_image subtracts PIXEL_MEAN (yaml:82-89) and zero-pads to a multiple of 32.
//     PIL resizes uint8 images in 22-bit fixed point, horizontal pass first, with a uint8 round trip
//     between the passes; the kernel below replays exactly that integer arithmetic (bit-exact), so the
//     coefficient tables are computed on the host in double precision like Pillow's precompute_coeffs.
//   * Mapper observation ("N2" glue): Agent_Helper._preprocess_obs / _preprocess_depth
//     (nav/agent/agent_helper.py:175-217) - per-column invalid-depth fill, too-far masking, cm conversion,
//     [2::4, 2::4] subsampling of depth / semantic masks / rgb and channel concatenation.
#include <cmath>

#include "maskrcnn.h"
#include "vec.cuh"

namespace pn {

// ---------------------------------------------------------------------------------------------------
// Pillow Resample.c: precompute_coeffs + normalize_coeffs_8bpc for the bilinear filter (support 1).
void pil_bilinear_coeffs(int in_size, int out_size, std::vector<int>& bounds, std::vector<int>& kk, int& ksize) {
  const double scale = static_cast<double>(in_size) / out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  ksize = static_cast<int>(std::ceil(support)) * 2 + 1;
  bounds.assign(static_cast<size_t>(out_size) * 2, 0);
  kk.assign(static_cast<size_t>(out_size) * ksize, 0);
  std::vector<double> k(ksize);
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
      double a = (x + xmin - center + 0.5) * ss;
      if (a < 0.0) a = -a;
      const double w = a < 1.0 ? 1.0 - a : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < xmax; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = xmax; x < ksize; ++x) k[x] = 0.0;
    for (int x = 0; x < ksize; ++x) {
      const double v = k[x] * static_cast<double>(1 << 22);
      kk[static_cast<size_t>(xx) * ksize + x] = v < 0 ? static_cast<int>(-0.5 + v) : static_cast<int>(0.5 + v);
    }
    bounds[xx * 2] = xmin;
    bounds[xx * 2 + 1] = xmax;
  }
}

namespace {

__device__ __forceinline__ int clip8(int v) {
  v >>= 22;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// One thread per output pixel (all 3 channels).  hb/hk, vb/vk: bounds + coefficients of the two passes.
// Output: resized uint8 BGR [B, Hn, Wn, 3].
__global__ void k_resize_u8(const uint8_t* const* rgb_slot, int B, int H, int W, int Hn, int Wn, const int* __restrict__ hb,
                            const int* __restrict__ hk, int hks, const int* __restrict__ vb, const int* __restrict__ vk, int vks,
                            uint8_t* __restrict__ resized_u8) {
  pdl_grid_sync();
  const uint8_t* rgb = *rgb_slot;
  Idx4 ix;
  if (!split_index(Hn, Wn, 1, static_cast<uint32_t>(B) * Hn * Wn, ix)) return;
  const int b = static_cast<int>(ix.a), y = static_cast<int>(ix.b), x = static_cast<int>(ix.c);
  const size_t idx = (static_cast<size_t>(b) * Hn + y) * Wn + x;
  const int x0 = hb[2 * x], xn = hb[2 * x + 1];
  const int y0 = vb[2 * y], yn = vb[2 * y + 1];
  int acc[3] = {1 << 21, 1 << 21, 1 << 21};
  for (int j = 0; j < yn; ++j) {
    const uint8_t* row = rgb + (static_cast<size_t>(b) * H + (y0 + j)) * W * 3;
    int h[3] = {1 << 21, 1 << 21, 1 << 21};
    for (int i = 0; i < xn; ++i) {
      const int k = hk[x * hks + i];
      const uint8_t* px = row + (x0 + i) * 3;
      h[0] += px[0] * k, h[1] += px[1] * k, h[2] += px[2] * k;
    }
    const int kv = vk[y * vks + j];
    acc[0] += clip8(h[0]) * kv, acc[1] += clip8(h[1]) * kv, acc[2] += clip8(h[2]) * kv;
  }
  uint8_t* o = resized_u8 + idx * 3;
  o[0] = static_cast<uint8_t>(clip8(acc[2])), o[1] = static_cast<uint8_t>(clip8(acc[1])), o[2] = static_cast<uint8_t>(clip8(acc[0]));
}

// Normalise + zero-pad + pack the horizontal taps of the stride-2 7x7 stem convolution into channels, TWO output pixels per
// packed vector: output pixels xo = 2j and 2j + 1 read input columns 4j - 3 .. 4j + 3 and 4j - 1 .. 4j + 5, nine columns x
// three channels = 27 values between them:
//   v[b, y, j, t*3 + c] = (bgr[b, y, 4*j - 3 + t, c] - mean[c]) / std[c],  t = 0..8   (0 outside the resized image),
// channels 27..31 zero - and TWO image rows per packed pixel: out[b, q, j, h*32 + k] = v[b, 2q - 1 + h, j, k] (q = 0 .. Hp / 2;
// row -1 and row Hp stay the zeros of the allocation).  Output row p of the stride-2 stem reads image rows 2p - 3 .. 2p + 3 =
// row pairs p - 1 .. p + 2 (the second half of the last pair gets zero weights), so the stem runs as a 4x1 convolution with 64
// input channels and 128 outputs (pixel parity x 64 channels, which is exactly the NHWC order of the two pixels) at stride 1,
// padding (1, 0): four K blocks of 128 bytes (full-width swizzle, four MMA K steps each) instead of seven of 64 bytes, half the
// GEMM rows and half the packed bytes of round 2's first one-pixel packing (891 MB written and read back per 32 frames),
// N = 128 instead of 64 per MMA.
template <typename T>
__global__ void k_pack_stem(const uint8_t* __restrict__ resized_u8, int B, int Hn, int Wn, int Hp, int Wq, float3 mean_bgr,
                            float3 std_bgr, T* __restrict__ out, int round_tf32) {
  pdl_grid_sync();
  // one thread per stored pixel (row pair q, pixel pair j): 2 x 27 values, 128 contiguous bytes (bf16) per thread
  const int Hq = Hp / 2 + 1;
  Idx4 ix;
  if (!split_index(Hq, Wq, 1, static_cast<uint32_t>(B) * Hq * Wq, ix)) return;
  const int b = static_cast<int>(ix.a), q = static_cast<int>(ix.b), j = static_cast<int>(ix.c);
  const size_t idx = (static_cast<size_t>(b) * Hq + q) * Wq + j;
  // PIXEL_STD is (1, 1, 1) in the reference's yaml: x / 1 == x exactly
  const bool unit_std = std_bgr.x == 1.f && std_bgr.y == 1.f && std_bgr.z == 1.f;
  T* o = out + idx * 64;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int y = 2 * q - 1 + h;
    float v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = 0.f;
    if (y >= 0 && y < Hn) {
      const uint8_t* row = resized_u8 + (static_cast<size_t>(b) * Hn + y) * Wn * 3;
#pragma unroll
      for (int s = 0; s < 9; ++s) {
        const int x = 4 * j - 3 + s;
        if (x >= 0 && x < Wn) {
          const float d0 = static_cast<float>(row[x * 3]) - mean_bgr.x, d1 = static_cast<float>(row[x * 3 + 1]) - mean_bgr.y;
          const float d2 = static_cast<float>(row[x * 3 + 2]) - mean_bgr.z;
          v[s * 3] = unit_std ? d0 : d0 / std_bgr.x;
          v[s * 3 + 1] = unit_std ? d1 : d1 / std_bgr.y;
          v[s * 3 + 2] = unit_std ? d2 : d2 / std_bgr.z;
        }
      }
      if (sizeof(T) == 4 && round_tf32) {
#pragma unroll
        for (int k = 0; k < 27; ++k) {
          uint32_t qq;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(qq) : "f"(v[k]));
          v[k] = __uint_as_float(qq);
        }
      }
    }
#pragma unroll
    for (int c0 = 0; c0 < 32; c0 += 8) {
      const float w[8] = {v[c0], v[c0 + 1], v[c0 + 2], v[c0 + 3], v[c0 + 4], v[c0 + 5], v[c0 + 6], v[c0 + 7]};
      store8(o + h * 32 + c0, w);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Mapper observation.  One warp per (env, subsampled column); lanes stride over the 480 rows of the full
// column for the invalid-pixel statistics, then over the 120 subsampled rows for the output.
__global__ void k_make_obs(const float* __restrict__ depth_all, const uint8_t* __restrict__ rgb_all,
                           const float* __restrict__ sem_all, int E, int H, int W, int ds, int h, int w, int nsem, float min_d,
                           float max_d, float* __restrict__ obs) {
  pdl_grid_sync();
  const int warps_per_block = blockDim.x >> 5;
  const int gw = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= E * w) return;
  const int e = gw / w, xs = gw - e * w;
  const int x = ds / 2 + xs * ds;
  const float* depth = depth_all + static_cast<size_t>(e) * H * W;
  const uint8_t* rgb = rgb_all ? rgb_all + static_cast<size_t>(e) * H * W * 3 : nullptr;
  const float* sem = sem_all + static_cast<size_t>(e) * H * W * nsem;
  // column statistics over ALL rows (agent_helper.py:200-206)
  int n_invalid = 0;
  float col_max = -INFINITY;
  for (int y = lane; y < H; y += 32) {
    const float d = depth[static_cast<size_t>(y) * W + x];
    n_invalid += (d == 0.f);
    col_max = fmaxf(col_max, d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_invalid += __shfl_xor_sync(0xffffffffu, n_invalid, o);
    col_max = fmaxf(col_max, __shfl_xor_sync(0xffffffffu, col_max, o));
  }
  // np.mean(invalid) > 0.9 in float64
  const bool mostly_invalid = static_cast<double>(n_invalid) / static_cast<double>(H) > 0.9;
  const int C = 4 + nsem;
  float* o = obs + static_cast<size_t>(e) * C * h * w;
  for (int ys = lane; ys < h; ys += 32) {
    const int y = ds / 2 + ys * ds;
    float d = depth[static_cast<size_t>(y) * W + x];
    if (d == 0.f) d = mostly_invalid ? col_max : 100.0f;
    if (d > 0.99f) d = 0.f;
    if (d == 0.f) d = 100.0f;
    d = min_d * 100.0f + d * (max_d - min_d) * 100.0f;
    const size_t pix = static_cast<size_t>(ys) * w + xs;
    const size_t src = static_cast<size_t>(y) * W + x;
    for (int c = 0; c < 3; ++c) o[static_cast<size_t>(c) * h * w + pix] = rgb ? static_cast<float>(rgb[src * 3 + c]) : 0.f;
    o[static_cast<size_t>(3) * h * w + pix] = d;
    for (int c = 0; c < nsem; ++c) o[static_cast<size_t>(4 + c) * h * w + pix] = sem[src * nsem + c];
  }
}

}  // namespace

void add_resize_pack_stem(Net& net, const uint8_t* const* rgb_slot, int B, int H, int W, int Hn, int Wn, const Tensor& out,
                          uint8_t* resized_u8, const float mean_bgr[3], const float std_bgr[3]) {
  PN_REQUIRE(out.ld == 64 && out.C == 64, "pack_stem: output must be a dense 64-channel tensor (two image rows per pixel)");
  const int Hp = 2 * (out.H - 1);   // padded image height: out holds Hp / 2 + 1 row pairs
  PN_REQUIRE(out.W * 4 >= Wn && Hp >= Hn, "pack_stem: output too small");
  std::vector<int> hb, hk, vb, vk;
  int hks = 0, vks = 0;
  pil_bilinear_coeffs(W, Wn, hb, hk, hks);
  pil_bilinear_coeffs(H, Hn, vb, vk, vks);
  const int* d_hb = net.arena.upload(hb);
  const int* d_hk = net.arena.upload(hk);
  const int* d_vb = net.arena.upload(vb);
  const int* d_vk = net.arena.upload(vk);
  const float3 mean = make_float3(mean_bgr[0], mean_bgr[1], mean_bgr[2]);
  const float3 sd = make_float3(std_bgr[0], std_bgr[1], std_bgr[2]);
  const int threads = 256;
  const long long total1 = static_cast<long long>(B) * Hn * Wn;
  check_u32_launch(total1, "resize_u8");
  check_u32_launch(out.pixels(), "pack_stem");
  const int blocks1 = static_cast<int>((total1 + threads - 1) / threads);
  net.add("resize_u8", [=](cudaStream_t s) {
    launch_pdl(k_resize_u8, blocks1, threads, 0, s, rgb_slot, B, H, W, Hn, Wn, d_hb, d_hk, hks, d_vb, d_vk, vks, resized_u8);
  });
  const long long total2 = out.pixels();   // one thread per stored pixel (row pair, pixel pair)
  const int blocks2 = static_cast<int>((total2 + threads - 1) / threads);
  Tensor o = out;
  const int rnd = net.x3 ? 0 : 1;
  net.add("pack_stem", [=](cudaStream_t s) {
    if (o.dt == kBF16)
      launch_pdl(k_pack_stem<__nv_bfloat16>, blocks2, threads, 0, s, resized_u8, B, Hn, Wn, Hp, o.W, mean, sd, static_cast<__nv_bfloat16*>(o.ptr), rnd);
    else
      launch_pdl(k_pack_stem<float>, blocks2, threads, 0, s, resized_u8, B, Hn, Wn, Hp, o.W, mean, sd, static_cast<float*>(o.ptr), rnd);
  });
  net.launches_per_forward += 2;
}

void launch_make_obs(const float* depth, const uint8_t* rgb, const float* sem, int E, int H, int W, int ds, int h, int w,
                     int nsem, float min_d, float max_d, float* obs, cudaStream_t s) {
  const int warps = E * w;
  const int threads = 128;
  const int blocks = (warps * 32 + threads - 1) / threads;
  launch_pdl(k_make_obs, blocks, threads, 0, s, depth, rgb, sem, E, H, W, ds, h, w, nsem, min_d, max_d, obs);
}


// ---------------------------------------------------------------------------------------------------
// Glue after stage C: Agent_State.update_prediction (nav/agent/agent_state.py:357-372).  The predicted maps cover the
// prediction window [x1, x1+win) x [y1, y1+win) of the full map; the reference embeds them into a full-size zero canvas,
// cuts the goal category's plane to the local-map bounds and keeps unexplored cells only.  One thread per local cell.
__global__ void k_target_pred(const float* __restrict__ pred, int win, int x1, int y1, int goal_cat, int r0, int c0,
                              int lw, int lh, const float* __restrict__ explored, long long explored_row_stride,
                              float* __restrict__ out) {
  pdl_grid_sync();
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= lw * lh) return;
  const int r = idx / lh, c = idx - r * lh;
  const int pr = r0 + r - x1, pc = c0 + c - y1;  // position inside the prediction window
  float v = 0.f;
  if (pr >= 0 && pr < win && pc >= 0 && pc < win) v = pred[(static_cast<size_t>(goal_cat) * win + pr) * win + pc];
  out[idx] = (explored[static_cast<size_t>(r) * explored_row_stride + c] < 0.5f) ? v : 0.f * v;  // x * False keeps the sign of zero / NaN
}

void launch_target_pred(const float* pred, int win, int x1, int y1, int goal_cat, int r0, int c0, int lw, int lh,
                        const float* explored, long long explored_row_stride, float* out, cudaStream_t s) {
  const int n = lw * lh;
  launch_pdl(k_target_pred, (n + 255) / 256, 256, 0, s, pred, win, x1, y1, goal_cat, r0, c0, lw, lh, explored,
             explored_row_stride, out);
}

}  // namespace pn
