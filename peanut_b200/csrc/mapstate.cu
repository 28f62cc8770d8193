// Map bookkeeping on the device, batched over environments:
//   Agent_State.get_local_map_boundaries   nav/agent/agent_state.py:153-177
//   Agent_State.init_map_and_pose          :180-211
//   Agent_State.init_with_obs (3x3 stamp)  :116-122
//   Agent_State.update_local_map (tail)    :276-303
//   Agent_State.update_full_map            :308-338
// The reference keeps poses in device tensors but derives every cell index on the host (`.cpu().numpy()` at :276, :315,
// :335 - three blocking syncs per step and environment) and writes the stamps as tiny indexed tensor ops.  Here all of
// that state (poses, window bounds, origins, planner pose vector, agent cell, distance to goal) lives in device arrays;
// one kernel per call does the cell arithmetic and the map writes for all environments, nothing synchronises.
//
// These kernels are HBM-bound byte movers (window copies: 2 x nc x local_w x local_h x 4 bytes per environment) or
// latency-bound scalar updates; compiled with -fmad=false so the pose arithmetic rounds like the reference's separate ops.
#include "engine.h"
#include "maskrcnn.h"

namespace pn {

namespace {

struct MapCfg {
  int nc, full_w, full_h, local_w, local_h;
  int res, size_cm, gds, grid, col_rad;
  float goal_dist;
  int f64_cells;
};

struct MapArrays {
  float* full_map;
  float* local_map;
  float* full_pose;
  float* local_pose;
  double* origins;
  int* lmb;
  double* planner;
  int* loc;
  double* dist_to_goal;
  const int* global_goal;
};

// int(r * 100.0 / map_resolution) with r a numpy float32 scalar: float32 arithmetic under numpy >= 2, float64 before.
__device__ __forceinline__ int cell_of(float v, const MapCfg& c) {
  if (c.f64_cells) return static_cast<int>(static_cast<double>(v) * 100.0 / static_cast<double>(c.res));
  return static_cast<int>(__fdiv_rn(__fmul_rn(v, 100.f), static_cast<float>(c.res)));
}

// Python slice start:stop on an axis of n elements -> [lo, hi) (possibly empty)
__device__ __forceinline__ void py_slice(int start, int stop, int n, int& lo, int& hi) {
  start = start < 0 ? max(start + n, 0) : min(start, n);
  stop = stop < 0 ? max(stop + n, 0) : min(stop, n);
  lo = start;
  hi = max(stop, start);
}

__device__ __forceinline__ int floor_mod(int a, int m) {
  const int r = a % m;
  return r < 0 ? r + m : r;
}

// get_local_map_boundaries (:153-177)
__device__ void boundaries(int loc_r, int loc_c, const MapCfg& c, int* lmb) {
  int gx1, gx2, gy1, gy2;
  if (c.gds > 1) {
    gx1 = loc_r - c.local_w / 2, gy1 = loc_c - c.local_h / 2;
    gx1 -= floor_mod(gx1, c.grid), gy1 -= floor_mod(gy1, c.grid);
    gx2 = gx1 + c.local_w, gy2 = gy1 + c.local_h;
    if (gx1 < 0) gx1 = 0, gx2 = c.local_w;
    if (gx2 > c.full_w) gx1 = c.full_w - c.local_w, gx2 = c.full_w;
    if (gy1 < 0) gy1 = 0, gy2 = c.local_h;
    if (gy2 > c.full_h) gy1 = c.full_h - c.local_h, gy2 = c.full_h;
  } else {
    gx1 = 0, gx2 = c.full_w, gy1 = 0, gy2 = c.full_h;
  }
  lmb[0] = gx1, lmb[1] = gx2, lmb[2] = gy1, lmb[3] = gy2;
}

// Is (r, c) written by local_map[1][(selem_idx[0] - R + r0, selem_idx[1] - R + c0)] = 1 ?  Negative indices wrap once
// (Python), indices >= n are an IndexError in the reference and are skipped.
__device__ __forceinline__ bool in_disk(int r, int c, int r0, int c0, int R, int nr, int ncol) {
  bool hit = false;
#pragma unroll
  for (int wr = 0; wr < 2; ++wr) {
    const int dr = r - (wr ? nr : 0) - r0;  // index r (wr = 0) or r - nr < 0 (wr = 1)
    if (dr < -R || dr > R) continue;
#pragma unroll
    for (int wc = 0; wc < 2; ++wc) {
      const int dc = c - (wc ? ncol : 0) - c0;
      if (dc < -R || dc > R) continue;
      hit |= (dr * dr + dc * dc <= R * R);
    }
  }
  return hit;
}

// update_local_map after the mapper call (:276-303).  grid (x, E); every thread derives the agent cell from the pose.
__global__ void __launch_bounds__(256) k_map_update_local(MapCfg cfg, MapArrays a) {
  pdl_grid_sync();
  const int e = blockIdx.y;
  const float px = a.local_pose[e * 3 + 0], py = a.local_pose[e * 3 + 1];
  const int loc_r = cell_of(py, cfg), loc_c = cell_of(px, cfg);
  const int g0 = a.global_goal[e * 2 + 0], g1 = a.global_goal[e * 2 + 1];
  const long long d2 = static_cast<long long>(loc_r - g0) * (loc_r - g0) + static_cast<long long>(loc_c - g1) * (loc_c - g1);
  const double dist = sqrt(static_cast<double>(d2)) * static_cast<double>(cfg.res);
  const bool at_goal = dist < static_cast<double>(cfg.goal_dist);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double* org = a.origins + e * 3;
    double* pl = a.planner + e * 7;
    pl[0] = static_cast<double>(px) + org[0];
    pl[1] = static_cast<double>(py) + org[1];
    pl[2] = static_cast<double>(a.local_pose[e * 3 + 2]) + org[2];
    a.loc[e * 2 + 0] = loc_r, a.loc[e * 2 + 1] = loc_c;
    a.dist_to_goal[e] = dist;
  }
  int rlo, rhi, clo, chi;
  py_slice(loc_r - 2, loc_r + 3, cfg.local_w, rlo, rhi);
  py_slice(loc_c - 2, loc_c + 3, cfg.local_h, clo, chi);
  const int R = cfg.col_rad + 1;
  const size_t plane = static_cast<size_t>(cfg.local_w) * cfg.local_h;
  float* lm = a.local_map + static_cast<size_t>(e) * cfg.nc * plane;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cfg.local_h), c = static_cast<int>(i - static_cast<size_t>(r) * cfg.local_h);
    const bool traj = r >= rlo && r < rhi && c >= clo && c < chi;
    lm[2 * plane + i] = traj ? 1.f : 0.f;  // channel 2 reset, then the 5x5 stamp
    if (traj) lm[3 * plane + i] = 1.f;
    if (in_disk(r, c, loc_r, loc_c, R, cfg.local_w, cfg.local_h) || (at_goal && in_disk(r, c, g0, g1, R, cfg.local_w, cfg.local_h)))
      lm[1 * plane + i] = 1.f;
  }
}

// init_with_obs (:116-122): 3x3 stamp on local_map[2:4] at the cell of the local pose.  One block per environment.
__global__ void k_map_stamp_initial(MapCfg cfg, MapArrays a) {
  pdl_grid_sync();
  const int e = blockIdx.x;
  const int loc_r = cell_of(a.local_pose[e * 3 + 1], cfg), loc_c = cell_of(a.local_pose[e * 3 + 0], cfg);
  int rlo, rhi, clo, chi;
  py_slice(loc_r - 1, loc_r + 2, cfg.local_w, rlo, rhi);
  py_slice(loc_c - 1, loc_c + 2, cfg.local_h, clo, chi);
  const size_t plane = static_cast<size_t>(cfg.local_w) * cfg.local_h;
  float* lm = a.local_map + static_cast<size_t>(e) * cfg.nc * plane;
  const int t = threadIdx.x;  // 18 = 2 channels x 3 x 3
  if (t < 18) {
    const int ch = 2 + t / 9, r = rlo + (t % 9) / 3, c = clo + t % 3;
    if (r < rhi && c < chi) lm[ch * plane + static_cast<size_t>(r) * cfg.local_h + c] = 1.f;
  }
}

// Window copy between full_map[:, lmb0:lmb1, lmb2:lmb3] and local_map.  grid (x, nc, E); kToFull: local -> full.
template <bool kToFull, typename V>
__global__ void __launch_bounds__(256) k_map_window(MapCfg cfg, MapArrays a) {
  pdl_grid_sync();
  constexpr int kVec = sizeof(V) / 4;
  const int e = blockIdx.z, ch = blockIdx.y;
  const int r0 = a.lmb[e * 4 + 0], c0 = a.lmb[e * 4 + 2];
  const int wv = cfg.local_h / kVec;
  const size_t n = static_cast<size_t>(cfg.local_w) * wv;
  float* full = a.full_map + (static_cast<size_t>(e) * cfg.nc + ch) * cfg.full_w * cfg.full_h;
  float* local = a.local_map + (static_cast<size_t>(e) * cfg.nc + ch) * cfg.local_w * cfg.local_h;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / wv), c = static_cast<int>(i - static_cast<size_t>(r) * wv) * kVec;
    V* pf = reinterpret_cast<V*>(full + static_cast<size_t>(r0 + r) * cfg.full_h + c0 + c);
    V* pl = reinterpret_cast<V*>(local + static_cast<size_t>(r) * cfg.local_h + c);
    if (kToFull) *pf = *pl;
    else *pl = *pf;
  }
}

// Pose / window update, one thread per environment.  mode 0: update_full_map :311-338 (full pose from the local pose, new
// window, new local pose, agent cell); mode 1: init_map_and_pose :186-211 (pose = map centre, 3x3 stamp on full_map[2:4]).
__global__ void k_map_recentre(MapCfg cfg, MapArrays a, int E, int mode) {
  pdl_grid_sync();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float* fp = a.full_pose + e * 3;
  float* lp = a.local_pose + e * 3;
  double* org = a.origins + e * 3;
  double* pl = a.planner + e * 7;
  if (mode == 1) {
    const float centre = static_cast<float>(static_cast<double>(cfg.size_cm) / 100.0 / 2.0);
    fp[0] = centre, fp[1] = centre, fp[2] = 0.f;
    pl[0] = centre, pl[1] = centre, pl[2] = 0.0;
  } else {
    for (int i = 0; i < 3; ++i) fp[i] = __fadd_rn(lp[i], static_cast<float>(org[i]));
  }
  const int loc_r = cell_of(fp[1], cfg), loc_c = cell_of(fp[0], cfg);
  if (mode == 1) {
    int rlo, rhi, clo, chi;
    py_slice(loc_r - 1, loc_r + 2, cfg.full_w, rlo, rhi);
    py_slice(loc_c - 1, loc_c + 2, cfg.full_h, clo, chi);
    const size_t plane = static_cast<size_t>(cfg.full_w) * cfg.full_h;
    float* fm = a.full_map + static_cast<size_t>(e) * cfg.nc * plane;
    for (int ch = 2; ch < 4; ++ch)
      for (int r = rlo; r < rhi; ++r)
        for (int c = clo; c < chi; ++c) fm[ch * plane + static_cast<size_t>(r) * cfg.full_h + c] = 1.f;
  }
  int* lmb = a.lmb + e * 4;
  boundaries(loc_r, loc_c, cfg, lmb);
  for (int i = 0; i < 4; ++i) pl[3 + i] = static_cast<double>(lmb[i]);
  org[0] = static_cast<double>(lmb[2] * cfg.res) / 100.0;
  org[1] = static_cast<double>(lmb[0] * cfg.res) / 100.0;
  org[2] = 0.0;
  for (int i = 0; i < 3; ++i) lp[i] = __fsub_rn(fp[i], static_cast<float>(org[i]));
  if (mode == 0) {
    a.loc[e * 2 + 0] = cell_of(lp[1], cfg);
    a.loc[e * 2 + 1] = cell_of(lp[0], cfg);
  }
}


// ---- Agent_State.update_goal_map (:423-452), every step: where the goal category is on the map (and no other category
// of channels 4..9 is), eroded goal_erode times with the 4-neighbour cross (outside the map counts as set, skimage's
// border_value=True) and dilated once (outside counts as clear).  n cross erosions = one erosion by the diamond
// |dr| + |dc| <= n, so a cell of the result depends on a (2n+3)^2 neighbourhood: each block stages the binarised goal
// channel of a 32 x 32 tile plus halo in shared memory, erodes into a 34 x 34 tile, dilates from there.
constexpr int kGoalTile = 32;
constexpr int kGoalMaxErode = 12;

__global__ void __launch_bounds__(256) k_goal_candidates(int nc, int w, int h, int erode, const float* __restrict__ local_map,
                                                         const int* __restrict__ goal_cat, const int* __restrict__ skip_morph,
                                                         float* __restrict__ goal_map, int* __restrict__ found) {
  pdl_grid_sync();
  extern __shared__ unsigned char sm[];
  const int e = blockIdx.z;
  const int cn = goal_cat[e] + 4;
  const size_t plane = static_cast<size_t>(w) * h;
  const float* lm = local_map + static_cast<size_t>(e) * nc * plane;
  float* out = goal_map + static_cast<size_t>(e) * plane;
  const int r_base = blockIdx.y * kGoalTile, c_base = blockIdx.x * kGoalTile;
  const bool valid_cat = cn >= 4 && cn < nc;  // the reference raises IndexError otherwise
  const bool morph = skip_morph[e] == 0;
  const int halo = erode + 1, s1 = kGoalTile + 2 * halo, s2 = kGoalTile + 2;
  unsigned char* bin = sm;             // s1 x s1: 0 clear, 1 set, 2 outside the map
  unsigned char* ero = sm + s1 * s1;   // s2 x s2
  if (morph && valid_cat) {
    for (int i = threadIdx.x; i < s1 * s1; i += blockDim.x) {
      const int r = r_base - halo + i / s1, c = c_base - halo + i % s1;
      unsigned char v = 2;
      if (r >= 0 && r < w && c >= 0 && c < h) v = lm[cn * plane + static_cast<size_t>(r) * h + c] != 0.f ? 1 : 0;
      bin[i] = v;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < s2 * s2; i += blockDim.x) {
      const int lr = i / s2, lc = i % s2;           // eroded-tile coordinates; (lr + erode, lc + erode) in the bin tile
      const int r = r_base - 1 + lr, c = c_base - 1 + lc;
      bool v = r >= 0 && r < w && c >= 0 && c < h;  // nothing outside the map for the dilation to pick up
      for (int dr = -erode; dr <= erode && v; ++dr) {
        const int span = erode - (dr < 0 ? -dr : dr);
        for (int dc = -span; dc <= span; ++dc) v = v && bin[(lr + erode + dr) * s1 + lc + erode + dc] != 0;
      }
      ero[i] = v ? 1 : 0;
    }
    __syncthreads();
  }
  bool any = false;
  for (int i = threadIdx.x; i < kGoalTile * kGoalTile; i += blockDim.x) {
    const int lr = i / kGoalTile, lc = i % kGoalTile;
    const int r = r_base + lr, c = c_base + lc;
    if (r >= w || c >= h) continue;
    const size_t idx = static_cast<size_t>(r) * h + c;
    float t = 0.f;
    if (valid_cat) {
      const float x = lm[cn * plane + idx];
      if (morph) {
        const int p = (lr + 1) * s2 + lc + 1;
        t = (ero[p] | ero[p - 1] | ero[p + 1] | ero[p - s2] | ero[p + s2]) ? 1.f : 0.f;
      } else {
        t = x > 0.f ? 1.f : x;
      }
      float others = nc > 4 ? lm[4 * plane + idx] : 0.f;  // torch.sum(local_map[4:10], dim=0): channel after channel
      for (int ch = 5; ch < 10 && ch < nc; ++ch) others = __fadd_rn(others, lm[ch * plane + idx]);
      t = (__fsub_rn(others, x) == 0.f) ? t : 0.f * t;
    }
    out[idx] = t;
    any |= (t != 0.f);
  }
  if (__syncthreads_or(any) && threadIdx.x == 0) atomicOr(found + e, 1);
}

// not found: goal_map = the long-term goal cell (:429-430)
__global__ void __launch_bounds__(256) k_goal_finalize(int w, int h, const int* __restrict__ global_goal, const int* __restrict__ found,
                                                       float* __restrict__ goal_map) {
  pdl_grid_sync();
  const int e = blockIdx.y;
  if (found[e]) return;
  int g0 = global_goal[e * 2 + 0], g1 = global_goal[e * 2 + 1];
  if (g0 < 0) g0 += w;
  if (g1 < 0) g1 += h;
  const size_t plane = static_cast<size_t>(w) * h;
  float* out = goal_map + static_cast<size_t>(e) * plane;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < plane; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / h), c = static_cast<int>(i - static_cast<size_t>(r) * h);
    out[i] = (r == g0 && c == g1) ? 1.f : 0.f;
  }
}

template <bool kToFull>
void launch_window(const MapCfg& cfg, const MapArrays& a, int E, cudaStream_t s) {
  const bool vec = cfg.local_h % 4 == 0 && cfg.full_h % 4 == 0 && cfg.grid % 4 == 0 && (cfg.full_h - cfg.local_h) % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(a.full_map) | reinterpret_cast<uintptr_t>(a.local_map)) % 16 == 0;
  const size_t n = static_cast<size_t>(cfg.local_w) * cfg.local_h / (vec ? 4 : 1);
  const dim3 grid(static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 64)), cfg.nc, E);
  if (vec) launch_pdl(k_map_window<kToFull, float4>, grid, dim3(256), 0, s, cfg, a);
  else launch_pdl(k_map_window<kToFull, float>, grid, dim3(256), 0, s, cfg, a);
}

// Prediction window of update_prediction (:353-360): out[e, c, r, :] = full_map[e, c, x1 + r, y1 : y1 + win_h] for c < nc_copy.
// grid (x, nc_copy, E); out has out_channels planes per environment (planes >= nc_copy are left alone).
template <typename V>
__global__ void __launch_bounds__(256) k_map_crop(const float* __restrict__ full, int nc, int full_w, int full_h, int x1, int y1,
                                                  int win_w, int win_h, float* __restrict__ out, int out_channels) {
  pdl_grid_sync();
  constexpr int kVec = sizeof(V) / 4;
  const int e = blockIdx.z, ch = blockIdx.y;
  const int wv = win_h / kVec;
  const size_t n = static_cast<size_t>(win_w) * wv;
  const float* src = full + (static_cast<size_t>(e) * nc + ch) * full_w * full_h;
  float* dst = out + (static_cast<size_t>(e) * out_channels + ch) * win_w * win_h;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / wv), c = static_cast<int>(i - static_cast<size_t>(r) * wv) * kVec;
    *reinterpret_cast<V*>(dst + static_cast<size_t>(r) * win_h + c) =
        *reinterpret_cast<const V*>(src + static_cast<size_t>(x1 + r) * full_h + y1 + c);
  }
}


// ---- N4, the map-sequence file format either side of the path: the two byte-level transforms on the device.
// Writer side, collect_maps.py:79-80: (full_map * 255).astype(uint8) - an fp32 multiply (this file is built with -fmad=false)
// and a truncation toward zero; numpy's result is only defined for products in [0, 256), which map values in [0, 1] satisfy.
__global__ void __launch_bounds__(256) k_map_quantize(const float* __restrict__ in, long long n, uint8_t* __restrict__ out, int vec) {
  pdl_grid_sync();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (vec) {
    const long long n4 = n >> 2;
    for (; i < n4; i += stride) {
      const float4 v = reinterpret_cast<const float4*>(in)[i];
      uchar4 q;
      q.x = static_cast<uint8_t>(static_cast<int>(v.x * 255.f)), q.y = static_cast<uint8_t>(static_cast<int>(v.y * 255.f));
      q.z = static_cast<uint8_t>(static_cast<int>(v.z * 255.f)), q.w = static_cast<uint8_t>(static_cast<int>(v.w * 255.f));
      reinterpret_cast<uchar4*>(out)[i] = q;
    }
    i = (n4 << 2) + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  }
  for (; i < n; i += stride) out[i] = static_cast<uint8_t>(static_cast<int>(in[i] * 255.f));
}

// Reader side, LoadMapFromFile.__call__ (train_prediction_model.py:66-84) on a sequence [T, C, W, H] uint8 resident in HBM:
//   img  = seq[t].transpose(1, 2, 0).astype(float32) / 255.          -> img_hwc [W, H, C] (the reference's array) and / or
//                                                                       img_chw [C, W, H] (what the completion net takes)
//   gt   = (seq[-1, goal0 : goal0 + G] * (1 - (img[:, :, 1] > 0))).transpose(1, 2, 0)   -> int64 [W, H, G]
// One thread per map cell: the plane reads and the CHW writes are coalesced across the warp; the cell-major outputs (C floats /
// G int64 per cell) go through shared memory, so that the CTA's 256 cells leave as one contiguous, coalesced run instead of 256
// scattered 56-byte pieces (img + input + target at 960 x 960: 97 -> 56 us, profiles/r02_map_dataset.txt).  Takes the shapes
// k_map_sample4 below cannot (cell count not a multiple of four, unaligned buffers).
__global__ void __launch_bounds__(256) k_map_sample(const uint8_t* __restrict__ seq, int T, int C, long long cells, int t_idx,
                                                    int goal0, int G, float* __restrict__ img_hwc, float* __restrict__ img_chw,
                                                    long long* __restrict__ gt) {
  pdl_grid_sync();
  extern __shared__ __align__(16) unsigned char s_raw[];
  float* s_img = reinterpret_cast<float*>(s_raw);                                   // [256][C] when img_hwc
  long long* s_gt = reinterpret_cast<long long*>(s_raw + (img_hwc ? sizeof(float) * 256 * C : 0));   // [256][G] when gt
  const long long base = static_cast<long long>(blockIdx.x) * blockDim.x;
  const long long cell = base + threadIdx.x;
  const int live = static_cast<int>(min(static_cast<long long>(blockDim.x), cells - base));
  if (cell < cells) {
    const uint8_t* cur = seq + static_cast<long long>(t_idx) * C * cells + cell;
    const uint8_t* last = seq + static_cast<long long>(T - 1) * C * cells + cell;
    const bool explored = cur[cells] > 0;   // channel 1 of the INPUT time step
    for (int c = 0; c < C; ++c) {
      const float v = static_cast<float>(cur[static_cast<long long>(c) * cells]) / 255.f;
      if (img_hwc) s_img[threadIdx.x * C + c] = v;
      if (img_chw) img_chw[static_cast<long long>(c) * cells + cell] = v;
    }
    if (gt)
      for (int g = 0; g < G; ++g)
        s_gt[threadIdx.x * G + g] = explored ? 0ll : static_cast<long long>(last[static_cast<long long>(goal0 + g) * cells]);
  }
  __syncthreads();
  if (img_hwc)
    for (int i = threadIdx.x; i < live * C; i += blockDim.x) img_hwc[base * C + i] = s_img[i];
  if (gt)
    for (int i = threadIdx.x; i < live * G; i += blockDim.x) gt[base * G + i] = s_gt[i];
}

// Four cells per thread (cells % 4 == 0): one uchar4 load per plane instead of four 1-byte loads (a warp's load covers a full 128-byte
// line instead of one 32-byte sector), float4 stores of the [C, W, H] planes, the cell-major outputs staged as above (the target as
// bytes, widened to int64 on the way out).  128 threads = 512 cells per CTA.  56 -> 43 us at 960 x 960, all three outputs.
__global__ void __launch_bounds__(128) k_map_sample4(const uint8_t* __restrict__ seq, int T, int C, long long cells, int t_idx,
                                                     int goal0, int G, float* __restrict__ img_hwc, float* __restrict__ img_chw,
                                                     long long* __restrict__ gt) {
  pdl_grid_sync();
  extern __shared__ __align__(16) unsigned char s_raw[];
  float* s_img = reinterpret_cast<float*>(s_raw);                                          // [512][C] when img_hwc
  uint8_t* s_gt = s_raw + (img_hwc ? sizeof(float) * 512 * C : 0);                         // [512][G] when gt
  const long long base = static_cast<long long>(blockIdx.x) * 512;
  const long long cell = base + 4 * threadIdx.x;
  const int live = static_cast<int>(min(512ll, cells - base));                             // a multiple of 4
  if (cell < cells) {
    const uint8_t* cur = seq + static_cast<long long>(t_idx) * C * cells + cell;
    const uint8_t* last = seq + static_cast<long long>(T - 1) * C * cells + cell;
    const uchar4 ex = *reinterpret_cast<const uchar4*>(cur + cells);   // channel 1 of the INPUT time step
    for (int c = 0; c < C; ++c) {
      const uchar4 q = *reinterpret_cast<const uchar4*>(cur + static_cast<long long>(c) * cells);
      float4 v;
      v.x = static_cast<float>(q.x) / 255.f, v.y = static_cast<float>(q.y) / 255.f;
      v.z = static_cast<float>(q.z) / 255.f, v.w = static_cast<float>(q.w) / 255.f;
      if (img_chw) *reinterpret_cast<float4*>(img_chw + static_cast<long long>(c) * cells + cell) = v;
      if (img_hwc) {
        float* d = s_img + (4 * threadIdx.x) * C + c;
        d[0] = v.x, d[C] = v.y, d[2 * C] = v.z, d[3 * C] = v.w;
      }
    }
    if (gt)
      for (int g = 0; g < G; ++g) {
        const uchar4 q = *reinterpret_cast<const uchar4*>(last + static_cast<long long>(goal0 + g) * cells);
        uint8_t* d = s_gt + (4 * threadIdx.x) * G + g;
        d[0] = ex.x ? 0 : q.x, d[G] = ex.y ? 0 : q.y, d[2 * G] = ex.z ? 0 : q.z, d[3 * G] = ex.w ? 0 : q.w;
      }
  }
  __syncthreads();
  if (img_hwc) {
    float4* dst = reinterpret_cast<float4*>(img_hwc + base * C);
    const float4* src = reinterpret_cast<const float4*>(s_img);
    for (int i = threadIdx.x; i < live * C / 4; i += blockDim.x) dst[i] = src[i];
  }
  if (gt)
    for (int i = threadIdx.x; i < live * G; i += blockDim.x) gt[base * G + i] = static_cast<long long>(s_gt[i]);
}

}  // namespace

void launch_map_stamp_local(const float* local_map, float* full_map, const int* lmb, int E, int nc, int local_w, int local_h,
                            int full_w, int full_h, cudaStream_t s) {
  MapCfg cfg{};
  cfg.nc = nc, cfg.full_w = full_w, cfg.full_h = full_h, cfg.local_w = local_w, cfg.local_h = local_h;
  MapArrays a{};
  a.full_map = full_map, a.local_map = const_cast<float*>(local_map), a.lmb = const_cast<int*>(lmb);
  launch_window<true>(cfg, a, E, s);
  PN_CUDA_CHECK(cudaGetLastError());
}

void launch_map_crop(const float* full_map, int E, int nc, int full_w, int full_h, int x1, int y1, int win_w, int win_h,
                     int nc_copy, float* out, int out_channels, cudaStream_t s) {
  const bool vec = win_h % 4 == 0 && full_h % 4 == 0 && y1 % 4 == 0 &&
                   (reinterpret_cast<uintptr_t>(full_map) | reinterpret_cast<uintptr_t>(out)) % 16 == 0;
  const size_t n = static_cast<size_t>(win_w) * win_h / (vec ? 4 : 1);
  const dim3 grid(static_cast<unsigned>(std::min<size_t>((n + 255) / 256, 64)), nc_copy, E);
  if (vec) launch_pdl(k_map_crop<float4>, grid, dim3(256), 0, s, full_map, nc, full_w, full_h, x1, y1, win_w, win_h, out, out_channels);
  else launch_pdl(k_map_crop<float>, grid, dim3(256), 0, s, full_map, nc, full_w, full_h, x1, y1, win_w, win_h, out, out_channels);
  PN_CUDA_CHECK(cudaGetLastError());
}

void launch_map_quantize(const float* map, long long n, uint8_t* out, int num_sms, cudaStream_t s) {
  const int vec = (reinterpret_cast<uintptr_t>(map) % 16 == 0 && reinterpret_cast<uintptr_t>(out) % 4 == 0) ? 1 : 0;
  const long long items = vec ? (n + 3) / 4 : n;
  const int blocks = static_cast<int>(std::max<long long>(1, std::min<long long>((items + 255) / 256, 8ll * num_sms)));
  launch_pdl(k_map_quantize, dim3(blocks), dim3(256), 0, s, map, n, out, vec);
  PN_CUDA_CHECK(cudaGetLastError());
}

void launch_map_sample(const uint8_t* seq, int T, int C, int W, int H, int t_idx, int goal0, int G, float* img_hwc, float* img_chw,
                       long long* gt, cudaStream_t s) {
  const long long cells = static_cast<long long>(W) * H;
  const size_t smem4 = (img_hwc ? sizeof(float) * 512 * C : 0) + (gt ? static_cast<size_t>(512) * G : 0);
  const bool aligned = cells % 4 == 0 && reinterpret_cast<uintptr_t>(seq) % 4 == 0 && reinterpret_cast<uintptr_t>(img_hwc) % 16 == 0 &&
                       reinterpret_cast<uintptr_t>(img_chw) % 16 == 0;
  if (aligned && smem4 <= 48 * 1024) {
    const long long blocks = (cells + 511) / 512;
    PN_REQUIRE(blocks < (1ll << 31), "map sample: map too large");
    launch_pdl(k_map_sample4, dim3(static_cast<unsigned>(blocks)), dim3(128), smem4, s, seq, T, C, cells, t_idx, goal0, G, img_hwc, img_chw, gt);
  } else {
    const long long blocks = (cells + 255) / 256;
    PN_REQUIRE(blocks < (1ll << 31), "map sample: map too large");
    const size_t smem = (img_hwc ? sizeof(float) * 256 * C : 0) + (gt ? sizeof(long long) * 256 * G : 0);
    PN_REQUIRE(smem <= 48 * 1024, "map sample: too many channels for the staging buffer");
    launch_pdl(k_map_sample, dim3(static_cast<unsigned>(blocks)), dim3(256), smem, s, seq, T, C, cells, t_idx, goal0, G, img_hwc, img_chw, gt);
  }
  PN_CUDA_CHECK(cudaGetLastError());
}

void map_bookkeeping(int op, const pn_map_cfg& c, const pn_map_arrays& arr, int E, cudaStream_t s) {
  MapCfg cfg{c.num_channels, c.full_w, c.full_h, c.local_w, c.local_h, c.map_resolution, c.map_size_cm, c.global_downscaling,
             c.grid_resolution, c.col_rad, c.goal_reached_dist, c.f64_cells};
  MapArrays a{arr.full_map, arr.local_map, arr.full_pose, arr.local_pose, arr.origins, arr.lmb, arr.planner_pose_inputs,
              arr.loc, arr.dist_to_goal, arr.global_goal};
  const size_t plane = static_cast<size_t>(cfg.local_w) * cfg.local_h;
  switch (op) {
    case 0:  // init_map_and_pose
      PN_CUDA_CHECK(cudaMemsetAsync(a.full_map, 0, sizeof(float) * E * cfg.nc * static_cast<size_t>(cfg.full_w) * cfg.full_h, s));
      launch_pdl(k_map_recentre, dim3((E + 63) / 64), dim3(64), 0, s, cfg, a, E, 1);
      launch_window<false>(cfg, a, E, s);
      break;
    case 1:  // init_with_obs stamp
      launch_pdl(k_map_stamp_initial, dim3(E), dim3(32), 0, s, cfg, a);
      break;
    case 2:  // update_local_map tail
      launch_pdl(k_map_update_local, dim3(static_cast<unsigned>(std::min<size_t>((plane + 255) / 256, 296)), E), dim3(256), 0, s,
                 cfg, a);
      break;
    case 3:  // update_full_map
      launch_window<true>(cfg, a, E, s);
      launch_pdl(k_map_recentre, dim3((E + 63) / 64), dim3(64), 0, s, cfg, a, E, 0);
      launch_window<false>(cfg, a, E, s);
      break;
    default:
      PN_REQUIRE(false, "map_bookkeeping: unknown op");
  }
  PN_CUDA_CHECK(cudaGetLastError());
}

void launch_goal_map(const float* local_map, int E, int nc, int w, int h, const int* goal_cat, const int* skip_morph,
                     const int* global_goal, int goal_erode, float* goal_map, int* found, cudaStream_t s) {
  PN_REQUIRE(goal_erode >= 0 && goal_erode <= kGoalMaxErode, "pn_goal_map: goal_erode out of range (0..12)");
  PN_CUDA_CHECK(cudaMemsetAsync(found, 0, sizeof(int) * E, s));
  const int halo = goal_erode + 1, s1 = kGoalTile + 2 * halo, s2 = kGoalTile + 2;
  const size_t smem = static_cast<size_t>(s1) * s1 + static_cast<size_t>(s2) * s2;
  const dim3 grid((h + kGoalTile - 1) / kGoalTile, (w + kGoalTile - 1) / kGoalTile, E);
  launch_pdl(k_goal_candidates, grid, dim3(256), smem, s, nc, w, h, goal_erode, local_map, goal_cat, skip_morph, goal_map, found);
  const size_t plane = static_cast<size_t>(w) * h;
  launch_pdl(k_goal_finalize, dim3(static_cast<unsigned>(std::min<size_t>((plane + 255) / 256, 148)), E), dim3(256), 0, s, w, h,
             global_goal, found, goal_map);
  PN_CUDA_CHECK(cudaGetLastError());
}

}  // namespace pn
