// Stage B: Semantic_Mapping.forward on the device, batched over environments.
//
// Reference: nav/agent/mapping.py:52-179 with helpers nav/agent/utils/depth_utils.py:129-252 and
// nav/agent/utils/model.py:7-43.  The reference materialises an 11 x 100 x 100 x 80 voxel grid (35 MB),
// scatter-adds into it eight times and rounds the WHOLE grid after each of the eight trilinear corners
// (depth_utils.py:241-250).  Here the grid never exists:
//
//   k_coords   one thread per depth pixel: normalised point coordinates (same fp32 op order as the reference)
//              and the two counts of the stair-mask test;
//   k_quantile one CTA per environment: stair-mask decision (3 % quantile by 4-pass radix select, not a sort);
//   k_hist     one thread per pixel: stair masking and a per-(x,y)-column histogram of the 2x2 lateral footprints;
//   k_scan     exclusive scan of the 10 000 column counters + work list of the non-empty columns;
//   k_fill     scatter of (corner, z-cell, point) keys into their column's bucket;
//   k_columns  persistent CTAs over the non-empty columns: sort the bucket by (corner_xy, z-cell, point), then thread z
//              replays the reference's accumulation for voxel (x,y,z) exactly - eight corner groups in
//              itertools.product order, points in index order, fp32 add, round-half-even after each group -
//              and the 80 voxels are reduced to the two height projections, thresholded and written to a
//              12 x 100 x 100 ego map;
//   k_fuse     for every local-map cell: the two chained bilinear grid_samples (rotate, then translate) are
//              evaluated on the fly from the 100 x 100 ego window (16 taps, zero outside) and max-fused
//              with maps_last.  Algorithmic traffic: read + write of the 14 x 480 x 480 map.
//
// This translation unit is compiled with -fmad=false: the coordinate and weight arithmetic must round after
// every multiply and add like the reference's separate torch ops do.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "semmap.h"

namespace pn {

namespace {

constexpr int kMaxFeat = 24;  // 1 + num_sem_categories upper bound held in registers per voxel (k_columns<24>)

__device__ __forceinline__ uint32_t float_key(float f) {  // order-preserving map float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(uint32_t k) {
  const uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
  return __uint_as_float(u);
}

// Normalised splat coordinates of pixel (row v, column u) - mapping.py:57-79.
__device__ __forceinline__ void point_coords(const SemMapCfg& c, int v, int u, float d, float& X, float& Y, float& Z) {
  X = (static_cast<float>(u) - c.xc) * d / c.f;
  Z = (static_cast<float>(c.h - 1 - v) - c.zc) * d / c.f;
  Y = d;
  Z = Z + c.agent_height;
  X = X + c.shift_x;
  X = X / c.res;
  Y = Y / c.res;
  X = (X - c.half_vr) / c.vr_f * 2.f;
  Y = (Y - c.half_vr) / c.vr_f * 2.f;
  Z = Z / c.res;
  Z = (Z - c.z_mid) / c.nz_f * 2.f;
}

// One dimension of splat_feat_nd (depth_utils.py:219-236): cell index and weight of corner ix.
__device__ __forceinline__ void corner(float coord, float G, int ix, int& p, float& w, bool& safe) {
  const float pos = coord * G / 2.f + G / 2.f;
  const float pf = floorf(pos) + static_cast<float>(ix);
  safe = (pf > 0.f) && (pf < G);
  w = safe ? (1.f - fabsf(pos - pf)) : 0.f;
  p = safe ? static_cast<int>(pf) : 0;
}

// --------------------------------------------------------------------------------------------------
// k_coords: one thread per depth pixel - splat coordinates + the two counts the stair-mask test needs
// (qcount[e] = {#valid heights, #heights in the stair band, #heights <= 0.2, -}).  grid = (ceil(N / 256), E).
__global__ void __launch_bounds__(256) k_coords(SemMapCfg c, const float* __restrict__ obs, float* __restrict__ coords,
                                                uint32_t* __restrict__ qcount) {
  pdl_grid_sync();
  const int e = blockIdx.y;
  const int N = c.h * c.w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* depth = obs + (static_cast<size_t>(e) * c.channels + 3) * N;
  float* cx = coords + static_cast<size_t>(e) * 3 * N;
  uint32_t valid = 0, mid = 0, low = 0;
  if (i < N) {
    float X, Y, Z;
    point_coords(c, i / c.w, i % c.w, depth[i], X, Y, Z);
    cx[i] = X, cx[N + i] = Y, cx[2 * N + i] = Z;
    if (Z > -1.f && Z < 1.f) {
      const float mz = Z * 2.f + 1.6f;
      valid = 1;
      mid = (mz > 0.2f && mz < 0.7f) ? 1u : 0u;
      low = (mz <= 0.2f) ? 1u : 0u;
    }
  }
  const uint32_t nv = __popc(__ballot_sync(0xffffffffu, valid != 0)), nm = __popc(__ballot_sync(0xffffffffu, mid != 0));
  const uint32_t nl = __popc(__ballot_sync(0xffffffffu, low != 0));
  if ((threadIdx.x & 31) == 0 && nv) {
    atomicAdd(&qcount[e * 4], nv);
    if (nm) atomicAdd(&qcount[e * 4 + 1], nm);
    if (nl) atomicAdd(&qcount[e * 4 + 2], nl);
  }
}

// --------------------------------------------------------------------------------------------------
// k_quantile: the stair-mask decision of mapping.py:90-100 (3 % quantile of the valid heights by a 4-pass radix
// select instead of a sort).  grid = E, block = 1024.
__global__ void __launch_bounds__(1024) k_quantile(SemMapCfg c, const float* __restrict__ coords,
                                                   const uint32_t* __restrict__ qcount, int* __restrict__ stair_flag) {
  pdl_grid_sync();
  const int e = blockIdx.x;
  const int N = c.h * c.w;
  const float* cz = coords + static_cast<size_t>(e) * 3 * N + 2 * static_cast<size_t>(N);
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_k, s_cnt_le, s_next;
  const int tid = threadIdx.x;
  const uint32_t n = qcount[e * 4];
  const uint32_t s_mid = qcount[e * 4 + 1];
  const uint32_t n_low = qcount[e * 4 + 2];

  // torch.quantile(my_zs, 0.03), linear interpolation: ranks = q*(n-1) in fp32 (mapping.py:94)
  bool flag = false;
  if (n > 0) {
    const float rank = 0.03f * static_cast<float>(n - 1);
    const float rb = floorf(rank);
    const uint32_t k_lo = static_cast<uint32_t>(rb);
    const uint32_t k_hi = static_cast<uint32_t>(ceilf(rank));
    const float wgt = rank - rb;
    // Only `quantile > 0.2` is used, and the interpolated quantile lies between the two order statistics k_lo, k_hi (the
    // lerp's roundings are monotone): with n_low = #heights <= 0.2 it is decided without selecting anything unless the
    // two statistics straddle 0.2 (n_low == k_lo + 1 == k_hi), and it is irrelevant when the stair-band test fails.
    const bool band = static_cast<float>(s_mid) > static_cast<float>(0.2 * static_cast<double>(n));
    if (!band || n_low >= k_hi + 1u || n_low <= k_lo) {
      if (tid == 0) stair_flag[e] = (band && n_low <= k_lo) ? 1 : 0;
      return;
    }
    // radix select of the k_lo-th smallest (0-based) over order-preserving keys, 8 bits per pass
    if (tid == 0) s_prefix = 0, s_k = k_lo;
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      for (int i = tid; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      const uint32_t mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      // heights cluster in a handful of bins: aggregate equal bins inside the warp before touching shared memory
      for (int i0 = 0; i0 < N; i0 += blockDim.x) {
        const int i = i0 + tid;
        uint32_t bin = 0xffffffffu;
        if (i < N) {
          const float Z = cz[i];
          if (Z > -1.f && Z < 1.f) {
            const uint32_t key = float_key(Z * 2.f + 1.6f);
            if ((key & mask) == prefix) bin = (key >> shift) & 255u;
          }
        }
        const uint32_t peers = __match_any_sync(0xffffffffu, bin);
        if (bin != 0xffffffffu && (tid & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], static_cast<uint32_t>(__popc(peers)));
      }
      __syncthreads();
      if (tid < 32) {
        // warp-parallel search of the bin that holds rank s_k: lane l owns bins [8l, 8l + 8)
        uint32_t c[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) c[j] = hist[tid * 8 + j], sum += c[j];
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
          if (tid >= o) incl += y;
        }
        const uint32_t before = incl - sum, k = s_k;
        __syncwarp();
        if (k >= before && k < incl) {
          uint32_t kk = k - before;
          int j = 0;
          while (kk >= c[j]) kk -= c[j], ++j;
          s_k = kk;
          s_prefix = prefix | (static_cast<uint32_t>(tid * 8 + j) << shift);
        }
      }
      __syncthreads();
    }
    const uint32_t key_lo = s_prefix;
    float v_lo = key_float(key_lo), v_hi = v_lo;
    if (k_hi != k_lo) {
      // (k_lo+1)-th order statistic: equal to v_lo if enough duplicates, else the smallest key above it
      if (tid == 0) s_cnt_le = 0, s_next = 0xffffffffu;
      __syncthreads();
      uint32_t cnt = 0, nxt = 0xffffffffu;
      for (int i = tid; i < N; i += blockDim.x) {
        const float Z = cz[i];
        if (Z > -1.f && Z < 1.f) {
          const uint32_t key = float_key(Z * 2.f + 1.6f);
          if (key <= key_lo) ++cnt;
          else nxt = min(nxt, key);
        }
      }
      atomicAdd(&s_cnt_le, cnt);
      atomicMin(&s_next, nxt);
      __syncthreads();
      v_hi = (s_cnt_le >= k_hi + 1) ? v_lo : key_float(s_next);
    }
    // at::lerp: weight < 0.5 ? a + w*(b-a) : b - (b-a)*(1-w)
    const float q = (fabsf(wgt) < 0.5f) ? __fmaf_rn(wgt, v_hi - v_lo, v_lo) : v_hi - (v_hi - v_lo) * (1.f - wgt);
    // `torch.sum(...) > 0.2 * len(my_zs)`: int64 tensor vs python float -> compared in fp32
    flag = (q > 0.2f) && (static_cast<float>(s_mid) > static_cast<float>(0.2 * static_cast<double>(n)));
  }
  if (tid == 0) stair_flag[e] = flag ? 1 : 0;
}

// --------------------------------------------------------------------------------------------------
// k_hist: stair masking + per-(x,y)-column histogram of the points' 2 x 2 lateral footprints.
// grid = (ceil(N / 256), E), one thread per depth pixel.
__global__ void __launch_bounds__(256) k_hist(SemMapCfg c, const float* __restrict__ obs, float* __restrict__ coords,
                                              const int* __restrict__ stair_flag, int* __restrict__ col_count) {
  pdl_grid_sync();
  const int e = blockIdx.y;
  const int N = c.h * c.w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* toilet = obs + (static_cast<size_t>(e) * c.channels + 4 + 4) * N;
  float* cx = coords + static_cast<size_t>(e) * 3 * N;
  float* cy = cx + N;
  float* cz = cy + N;
  const bool mask_stairs = stair_flag[e] != 0;
  int* counts = col_count + static_cast<size_t>(e) * c.vr * c.vr;
  {
    float X = cx[i], Y = cy[i], Z = cz[i];
    if (mask_stairs && (Z * 2.f + 1.6f < 0.7f) && toilet[i] == 0.f) {
      X = Y = Z = 99999.f;
      cx[i] = X, cy[i] = Y, cz[i] = Z;
    }
    int pz0, pz1;
    float wz0, wz1;
    bool sz0, sz1;
    corner(Z, c.nz_f, 0, pz0, wz0, sz0);
    corner(Z, c.nz_f, 1, pz1, wz1, sz1);
    if (!sz0 && !sz1) return;
#pragma unroll
    for (int ix = 0; ix < 2; ++ix) {
      int px;
      float wx;
      bool sx;
      corner(X, c.vr_f, ix, px, wx, sx);
      if (!sx) continue;
#pragma unroll
      for (int iy = 0; iy < 2; ++iy) {
        int py;
        float wy;
        bool sy;
        corner(Y, c.vr_f, iy, py, wy, sy);
        if (!sy) continue;
        atomicAdd(&counts[px * c.vr + py], 1);
      }
    }
  }
}

// --------------------------------------------------------------------------------------------------
// k_scan: exclusive prefix sum of the column counters; also resets the fill cursors.  grid = E, block = 1024: thread t owns a
// run of consecutive columns (serial sum), one block scan over the 1024 run totals.
// It also compacts the non-empty columns into four work lists (col_list is [E][2][ncols], list_n is [E][4]):
//   tiny (at most kTinyCap entries; plane 0 from the front)  -> k_columns_warp, one warp per column;
//   mid  (at most kMidCap entries; plane 1 from the front)   -> k_columns_cta, one CTA per column, entries staged in shared memory;
//   large (at most kLargeCap entries; plane 1 from the back) -> k_columns_cta with a 2048-entry stage, one CTA per SM;
//   big  (everything else; plane 0 from the back)            -> k_columns_big, keys only in shared memory.
// The order inside a list is irrelevant, every column is processed independently.
constexpr int kTinyCap = 32, kMidCap = 512, kLargeCap = 2048;
constexpr int kScanPer = 16;  // columns per thread of k_scan (1024 threads): vision_range up to 128
__global__ void __launch_bounds__(1024) k_scan(int ncols, const int* __restrict__ col_count, int* __restrict__ col_start,
                                               int* __restrict__ col_fill, int* __restrict__ col_list, int* __restrict__ list_n) {
  pdl_grid_sync();
  extern __shared__ int s_cnt[];   // [ncols]: the counters, then the prefix sums (global traffic stays coalesced: a single SM
                                   // retires about one 32-byte sector per cycle, 16 us for this kernel with strided accesses)
  const int e = blockIdx.x;
  int* list = col_list + static_cast<size_t>(e) * 2 * ncols;
  __shared__ int warp_sums[32];
  __shared__ uint32_t cls_sums[2][32];
  const int* cnt = col_count + static_cast<size_t>(e) * ncols;
  int* start = col_start + static_cast<size_t>(e) * (ncols + 1);
  int* fill = col_fill + static_cast<size_t>(e) * ncols;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ncols; i += blockDim.x) s_cnt[i] = cnt[i], fill[i] = 0;
  __syncthreads();
  const int per = (ncols + static_cast<int>(blockDim.x) - 1) / static_cast<int>(blockDim.x);
  const int c0 = min(static_cast<int>(threadIdx.x) * per, ncols), c1 = min(c0 + per, ncols);
  int v[kScanPer];   // this thread's run of consecutive counters
#pragma unroll
  for (int j = 0; j < kScanPer; ++j) v[j] = (c0 + j < c1) ? s_cnt[c0 + j] : 0;
  int run = 0;
#pragma unroll
  for (int j = 0; j < kScanPer; ++j) run += v[j];
  // one block scan of the run totals and of the packed per-thread class counts (16 bits each: at most 16 384 columns);
  // shared-memory atomics on four list counters would serialise
  uint32_t ca = 0, cb = 0;   // ca = tiny | mid << 16, cb = large | big << 16
#pragma unroll
  for (int j = 0; j < kScanPer; ++j) {
    if (v[j] > 0) {
      if (v[j] <= kTinyCap) ca += 1u;
      else if (v[j] <= kMidCap) ca += 1u << 16;
      else if (v[j] <= kLargeCap) cb += 1u;
      else cb += 1u << 16;
    }
  }
  int x = run;
  uint32_t xa = ca, xb = cb;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    const uint32_t ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
    if (lane >= o) x += y, xa += ya, xb += yb;
  }
  if (lane == 31) warp_sums[wid] = x, cls_sums[0][wid] = xa, cls_sums[1][wid] = xb;
  __syncthreads();
  if (wid == 0) {
    int sw = warp_sums[lane];
    uint32_t sa = cls_sums[0][lane], sb = cls_sums[1][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, sw, o);
      const uint32_t ya = __shfl_up_sync(0xffffffffu, sa, o), yb = __shfl_up_sync(0xffffffffu, sb, o);
      if (lane >= o) sw += y, sa += ya, sb += yb;
    }
    warp_sums[lane] = sw, cls_sums[0][lane] = sa, cls_sums[1][lane] = sb;
  }
  __syncthreads();
  int excl = (wid ? warp_sums[wid - 1] : 0) + x - run;
  const uint32_t ea = (wid ? cls_sums[0][wid - 1] : 0u) + xa - ca, eb = (wid ? cls_sums[1][wid - 1] : 0u) + xb - cb;
  int p_tiny = static_cast<int>(ea & 0xffffu), p_mid = static_cast<int>(ea >> 16);
  int p_large = static_cast<int>(eb & 0xffffu), p_big = static_cast<int>(eb >> 16);
#pragma unroll
  for (int j = 0; j < kScanPer; ++j) {
    const int i = c0 + j;
    if (i < c1) {
      s_cnt[i] = excl;
      excl += v[j];
      if (v[j] > 0) {
        if (v[j] <= kTinyCap) list[p_tiny++] = i;
        else if (v[j] <= kMidCap) list[ncols + p_mid++] = i;
        else if (v[j] <= kLargeCap) list[2 * ncols - 1 - p_large++] = i;
        else list[ncols - 1 - p_big++] = i;
      }
    }
  }
  if (threadIdx.x == blockDim.x - 1) {
    start[ncols] = excl;
    list_n[e * 4] = p_tiny, list_n[e * 4 + 1] = p_big, list_n[e * 4 + 2] = p_mid, list_n[e * 4 + 3] = p_large;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncols; i += blockDim.x) start[i] = s_cnt[i];
}

// --------------------------------------------------------------------------------------------------
// k_fill: key = corner_xy << 22 | z-cell of the LOWER z corner << 15 | point index.
__global__ void k_fill(SemMapCfg c, const float* __restrict__ coords, const int* __restrict__ col_start,
                       int* __restrict__ col_fill, uint32_t* __restrict__ entries) {
  pdl_grid_sync();
  const int e = blockIdx.y;
  const int N = c.h * c.w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int ncols = c.vr * c.vr;
  const float* cx = coords + static_cast<size_t>(e) * 3 * N;
  const float X = cx[i], Y = cx[N + i], Z = cx[2 * N + i];
  int pz0, pz1;
  float wz0, wz1;
  bool sz0, sz1;
  corner(Z, c.nz_f, 0, pz0, wz0, sz0);
  corner(Z, c.nz_f, 1, pz1, wz1, sz1);
  if (!sz0 && !sz1) return;
  const int zlow = static_cast<int>(floorf(Z * c.nz_f / 2.f + c.nz_f / 2.f));  // in [-1, nz-1] here
  const uint32_t zkey = static_cast<uint32_t>(zlow + 1);                        // biased: 0 .. nz
  const int* start = col_start + static_cast<size_t>(e) * (ncols + 1);
  int* fill = col_fill + static_cast<size_t>(e) * ncols;
  uint32_t* ent = entries + static_cast<size_t>(e) * 4 * N;
#pragma unroll
  for (int ix = 0; ix < 2; ++ix) {
    int px;
    float wx;
    bool sx;
    corner(X, c.vr_f, ix, px, wx, sx);
    if (!sx) continue;
#pragma unroll
    for (int iy = 0; iy < 2; ++iy) {
      int py;
      float wy;
      bool sy;
      corner(Y, c.vr_f, iy, py, wy, sy);
      if (!sy) continue;
      const int col = px * c.vr + py;
      const int slot = start[col] + atomicAdd(&fill[col], 1);
      ent[slot] = (static_cast<uint32_t>(ix * 2 + iy) << 22) | (zkey << 15) | static_cast<uint32_t>(i);
    }
  }
}

// --------------------------------------------------------------------------------------------------
// Voxel columns.  Column (px, py) owns the keys that k_fill put into its bucket; voxel (px, py, z) receives, in the
// reference's order, eight groups of additions - lateral corner ixy = 0..3 (itertools.product order), inside it the z corner
// iz = 0, 1 - each group's entries in point-index order, fp32 product then fp32 add per entry, and a round-half-even of
// the running value after every group (depth_utils.py:241-250).  Sorting the keys (ixy, lower z cell + 1, point) makes
// every group of every voxel a contiguous key range: group (ixy, iz) of voxel z = keys with z-field z - iz + 1.
// The chain of additions of one (voxel, feature) pair is inherently serial, different pairs are independent: a lane owns
// one (voxel, feature) pair, kL lanes (16: at most 12 features, 32: at most 24) form a voxel slot.
__device__ __forceinline__ int lower_bound_key(const uint32_t* keys, int n, uint32_t k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Shared-memory working set of one column whose entries are staged (sorted keys, both weights and the feature vector of
// every entry, feature 0 = the constant 1 of the count channel).
template <int kCap, int kFs>
struct ColumnStage {
  uint32_t keys[kCap];
  float w0[kCap], w1[kCap];
  float feat[kCap * kFs];
  uint32_t zbits[4];        // z-fields (0 .. nz) that occur in the column's keys
  int vcnt[4];              // touched voxels per 32-voxel chunk (16-byte aligned: read with one vector load)
  int nv, pad_[3];          // (kept out of that vector: thread 0 writes nv while other warps still read vcnt)
  uint8_t vox[128];         // touched voxels, any order (the height projections are sums of integers: exact, order-free)
  float red_all[32], red_agent[32];
};

// Every thread of a warp stages up to kB entries (sorted positions pos[u], keys k[u], inactive slots masked): all
// coordinate and feature loads of the batch are issued before the first value is used (a loop of dependent L2 round trips
// otherwise), and the z-fields are merged inside the warp before one lane touches the shared mask (hundreds of
// shared-memory atomics on the same word would serialise).  Must be called by all 32 lanes.
template <int kB, bool kOwnWarp, int kCap, int kFs>
__device__ __forceinline__ void stage_batch(const SemMapCfg& c, ColumnStage<kCap, kFs>& st, const int (&pos)[kB],
                                            const uint32_t (&k)[kB], const bool (&act)[kB], const float* __restrict__ cxp,
                                            const float* __restrict__ featp, int N) {
  float X[kB], Y[kB], Z[kB], fv[kB][kFs];
#pragma unroll
  for (int u = 0; u < kB; ++u) {
    const int i = static_cast<int>(k[u] & 0x7fffu);   // inactive: key 0 -> point 0, a valid address
    X[u] = cxp[i], Y[u] = cxp[N + i], Z[u] = cxp[2 * N + i];
#pragma unroll
    for (int f = 1; f < kFs; ++f) fv[u][f] = (f < c.nf) ? featp[static_cast<size_t>(f - 1) * N + i] : 0.f;
  }
  uint32_t zw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int u = 0; u < kB; ++u) {
    if (act[u]) {
      // weights of the entry on its two voxels, the reference's products in the reference's order: ((1 * wx) * wy) * wz
      const int ixy = static_cast<int>(k[u] >> 22);
      int pp;
      float wx, wy, wz0, wz1;
      bool ss;
      corner(X[u], c.vr_f, ixy >> 1, pp, wx, ss);
      corner(Y[u], c.vr_f, ixy & 1, pp, wy, ss);
      corner(Z[u], c.nz_f, 0, pp, wz0, ss);
      corner(Z[u], c.nz_f, 1, pp, wz1, ss);
      const float wxy = (1.f * wx) * wy;
      st.w0[pos[u]] = wxy * wz0, st.w1[pos[u]] = wxy * wz1;
      float* fr = st.feat + pos[u] * kFs;
      fr[0] = 1.f;
#pragma unroll
      for (int f = 1; f < kFs; ++f) fr[f] = fv[u][f];
      const uint32_t word = (k[u] >> 20) & 3u, bit = 1u << ((k[u] >> 15) & 31u);
#pragma unroll
      for (int w = 0; w < 4; ++w) zw[w] |= (word == static_cast<uint32_t>(w)) ? bit : 0u;
    }
  }
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const uint32_t m = __reduce_or_sync(0xffffffffu, zw[w]);
    if ((threadIdx.x & 31) == 0) {
      if constexpr (kOwnWarp) st.zbits[w] = m;          // the warp owns the column: plain store, nothing to clear
      else if (m != 0u) atomicOr(&st.zbits[w], m);
    }
  }
}

// The replay of one staged column by kWarps warps (tid = thread index inside the group, `sync` = barrier of the group).
// Leaves the column's twelve (or more) ego-map values in global memory.
template <int kL, int kWarps, int kCap, int kFs, typename Sync>
__device__ __forceinline__ void replay_column(const SemMapCfg& c, ColumnStage<kCap, kFs>& st, int n, int tid, Sync sync,
                                              float* __restrict__ ego_e, int cell, int ncols) {
  constexpr int kSlotsPerWarp = 32 / kL;
  const int lane = tid & 31, wid = tid >> 5;
  // voxel z is touched by keys with z-field z + 1 (iz = 0) or z (iz = 1); voxel 0 is never written (corner(): pf > 0)
  // (compacted with ballots: chunk w = voxels [32 w, 32 w + 32), nz <= 126)
  auto touched_at = [&](int z) {
    return z > 0 && z < c.nz && (((st.zbits[(z + 1) >> 5] >> ((z + 1) & 31)) | (st.zbits[z >> 5] >> (z & 31))) & 1u) != 0u;
  };
  if constexpr (kWarps == 1) {
    int base = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const int z = w * 32 + lane;
      const bool t = touched_at(z);
      const uint32_t m = __ballot_sync(0xffffffffu, t);
      if (t) st.vox[base + __popc(m & ((1u << lane) - 1u))] = static_cast<uint8_t>(z);
      base += __popc(m);
    }
    if (lane == 0) st.nv = base;
  } else {
    bool t = false;
    uint32_t m = 0u;
    if (wid < 4) {
      t = touched_at(tid);
      m = __ballot_sync(0xffffffffu, t);
      if (lane == 0) st.vcnt[wid] = __popc(m);
    }
    sync();
    if (wid < 4) {
      int base = 0;
      for (int w = 0; w < wid; ++w) base += st.vcnt[w];
      if (t) st.vox[base + __popc(m & ((1u << lane) - 1u))] = static_cast<uint8_t>(tid);
      if (tid == 0) st.nv = st.vcnt[0] + st.vcnt[1] + st.vcnt[2] + st.vcnt[3];
    }
  }
  sync();
  const int nv = st.nv;
  const int sl = lane / kL, f = lane % kL;
  const bool feat_lane = f < c.nf;
  float all_h = 0.f, agent_h = 0.f;
  for (int base = wid * kSlotsPerWarp; base < nv; base += kWarps * kSlotsPerWarp) {   // warp-uniform trip count
    const int v = base + sl;
    const bool valid = v < nv;
    const int z = valid ? static_cast<int>(st.vox[v]) : 0;
    // the 16 range boundaries of this voxel's eight groups, one per lane of the slot
    int bound = 0;
    if (valid && f < 16) {
      const int g = f >> 1, ixy = g >> 1, iz = g & 1;
      const uint32_t kbase = (static_cast<uint32_t>(ixy) << 22) | (static_cast<uint32_t>(z - iz + 1) << 15);
      bound = lower_bound_key(st.keys, n, kbase + (static_cast<uint32_t>(f & 1) << 15));
    }
    int lo[8], hi[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      lo[g] = __shfl_sync(0xffffffffu, bound, sl * kL + 2 * g);
      hi[g] = __shfl_sync(0xffffffffu, bound, sl * kL + 2 * g + 1);
    }
    if (valid && feat_lane) {
      float acc = 0.f;
      const float* fcol = st.feat + f;
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        const float* wz = (g & 1) ? st.w1 : st.w0;
        for (int t = lo[g]; t < hi[g]; ++t) acc = acc + fcol[t * kFs] * wz[t];
        acc = rintf(acc);   // grid_flat = torch.round(grid_flat) after every corner (depth_utils.py:250): half-to-even
      }
      // height projections (mapping.py:102-113)
      all_h += acc;
      if (z >= c.min_z && z < c.max_z) agent_h += acc;
    }
  }
  if constexpr (kSlotsPerWarp == 2) {
    all_h += __shfl_xor_sync(0xffffffffu, all_h, 16);
    agent_h += __shfl_xor_sync(0xffffffffu, agent_h, 16);
  }
  if constexpr (kWarps == 1) {
    if (lane < kL) st.red_all[lane] = all_h, st.red_agent[lane] = agent_h;
  } else {
    if (lane < kL && feat_lane && all_h != 0.f) {
      atomicAdd(&st.red_all[lane], all_h);
      if (agent_h != 0.f) atomicAdd(&st.red_agent[lane], agent_h);
    }
  }
  sync();
  if (tid < c.ego_channels) {
    // ego channel 0 = obstacle, 1 = explored, 2.. = categories (local-map channels 4..)
    const int ch = tid;
    const int ff = ch < 2 ? 0 : ch - 1;
    const float ah = st.red_all[ff], gh = st.red_agent[ff];
    float v;
    if (ch == 0) v = gh / c.map_thr;
    else if (ch == 1) v = ah / c.exp_thr;
    else {
      const bool use_all = (ff == c.special_f[0] || ff == c.special_f[1] || ff == c.special_f[2]);
      v = (use_all ? ah : gh) / c.cat_thr;
    }
    ego_e[static_cast<size_t>(ch) * ncols + cell] = fminf(fmaxf(v, 0.f), 1.f);
  }
}

// k_columns_warp: columns with at most kTinyCap (= 32) entries - nine in ten - one WARP per column, four independent warps
// per CTA, persistent over the tiny list; no CTA barrier anywhere.  Lane l owns entry l: rank sort (the keys are unique,
// they embed the point index: the number of smaller keys is the sorted position), weights and features staged once.
template <int kL, int kFs>
__global__ void __launch_bounds__(128) k_columns_warp(SemMapCfg c, const float* __restrict__ obs, const float* __restrict__ coords,
                                                      const int* __restrict__ col_start, const uint32_t* __restrict__ entries,
                                                      const int* __restrict__ col_list, const int* __restrict__ list_n,
                                                      float* __restrict__ ego) {
  pdl_grid_sync();
  __shared__ ColumnStage<kTinyCap, kFs> stage[4];
  __shared__ uint32_t raw_all[4][kTinyCap];
  const int e = blockIdx.y;
  const int ncols = c.vr * c.vr;
  const int N = c.h * c.w;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  ColumnStage<kTinyCap, kFs>& st = stage[wid];
  uint32_t* raw = raw_all[wid];
  const int* start = col_start + static_cast<size_t>(e) * (ncols + 1);
  const int* list = col_list + static_cast<size_t>(e) * 2 * ncols;
  const int n_list = list_n[e * 4];
  float* ego_e = ego + static_cast<size_t>(e) * c.ego_channels * ncols;
  const float* cxp = coords + static_cast<size_t>(e) * 3 * N;
  const float* featp = obs + (static_cast<size_t>(e) * c.channels + 4) * N;
  auto sync = [] { __syncwarp(); };
  for (int li = blockIdx.x * 4 + wid; li < n_list; li += gridDim.x * 4) {
    const int col = list[li];
    const int s0 = start[col];
    const int n = start[col + 1] - s0;
    const int px = col / c.vr, py = col - px * c.vr;
    const int cell = py * c.vr + px;  // voxels.transpose(2,3): row = y index, column = x index
    __syncwarp();  // the previous column's stage is no longer read
    uint32_t k = 0;
    if (lane < n) k = entries[static_cast<size_t>(e) * 4 * N + s0 + lane], raw[lane] = k;
    __syncwarp();
    int rank = 0;
    if (lane < n) {
      for (int j = 0; j < n; ++j) rank += raw[j] < k ? 1 : 0;
      st.keys[rank] = k;
    }
    {
      const int pos[1] = {rank};
      const uint32_t kk[1] = {k};
      const bool act[1] = {lane < n};
      stage_batch<1, true>(c, st, pos, kk, act, cxp, featp, N);
    }
    __syncwarp();
    replay_column<kL, 1>(c, st, n, lane, sync, ego_e, cell, ncols);
  }
}

// k_columns_cta: columns with more entries (walls and far floor: 7 % of the columns, two thirds of the entries), one CTA per
// column, persistent over the mid list (kTinyCap < n <= kMidCap = 512: eight warps = sixteen voxel slots, several CTAs per SM)
// or the large list (kMidCap < n <= kLargeCap = 2048: sixteen warps, one CTA per SM).  Bitonic sort of the keys in shared memory, then every
// thread stages the entries at its sorted positions: a far floor cell collects hundreds of points in ONE voxel, whose
// serial chain of additions then runs out of shared memory (a few cycles per entry) instead of global memory.
template <int kL, int kFs, int kCap, int kList, int kWarps>
__global__ void __launch_bounds__(kWarps * 32) k_columns_cta(SemMapCfg c, const float* __restrict__ obs, const float* __restrict__ coords,
                                                     const int* __restrict__ col_start, const uint32_t* __restrict__ entries,
                                                     const int* __restrict__ col_list, const int* __restrict__ list_n,
                                                     float* __restrict__ ego) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t stage_raw[];
  ColumnStage<kCap, kFs>& st = *reinterpret_cast<ColumnStage<kCap, kFs>*>(stage_raw);
  constexpr int kThreads = kWarps * 32;
  const int e = blockIdx.y;
  const int ncols = c.vr * c.vr;
  const int N = c.h * c.w;
  const int tid = threadIdx.x;
  const int* start = col_start + static_cast<size_t>(e) * (ncols + 1);
  // kList 2: the mid list (plane 1 from the front); 3: the large list (plane 1 from the back)
  const int* list = col_list + static_cast<size_t>(e) * 2 * ncols + ncols;
  const int n_list = list_n[e * 4 + kList];
  float* ego_e = ego + static_cast<size_t>(e) * c.ego_channels * ncols;
  const float* cxp = coords + static_cast<size_t>(e) * 3 * N;
  const float* featp = obs + (static_cast<size_t>(e) * c.channels + 4) * N;
  auto sync = [] { __syncthreads(); };
  for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
    const int col = kList == 2 ? list[li] : list[ncols - 1 - li];
    const int s0 = start[col];
    const int n = start[col + 1] - s0;
    const int px = col / c.vr, py = col - px * c.vr;
    const int cell = py * c.vr + px;
    __syncthreads();  // the previous column's stage is no longer read
    if (tid < 4) st.zbits[tid] = 0u;
    if (tid < 32) st.red_all[tid] = 0.f, st.red_agent[tid] = 0.f;
    int npow = 64;
    while (npow < n) npow <<= 1;
    const uint32_t* ent = entries + static_cast<size_t>(e) * 4 * N + s0;
    for (int i = tid; i < npow; i += kThreads) st.keys[i] = i < n ? ent[i] : 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < npow; i += kThreads) {
          const int l = i ^ j;
          if (l > i) {
            const uint32_t a = st.keys[i], b = st.keys[l];
            const bool up = (i & k) == 0;
            if ((a > b) == up) st.keys[i] = b, st.keys[l] = a;
          }
        }
        __syncthreads();
      }
    }
    constexpr int kB = 2;   // entries per thread and batch (four would cost 128 registers and a third of the resident CTAs)
    for (int p0 = 0; p0 < n; p0 += kThreads * kB) {   // warp-uniform trip count
      int pos[kB];
      uint32_t kk[kB];
      bool act[kB];
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        pos[u] = p0 + u * kThreads + tid;
        act[u] = pos[u] < n;
        kk[u] = act[u] ? st.keys[pos[u]] : 0u;
      }
      stage_batch<kB, false>(c, st, pos, kk, act, cxp, featp, N);
    }
    __syncthreads();
    replay_column<kL, kWarps>(c, st, n, tid, sync, ego_e, cell, ncols);
  }
}

// k_columns_big: the rare columns with more than kMidCap entries (a wall seen edge-on), one CTA per column: only the keys
// fit shared memory (up to 32 768), thread z replays voxel z with its entries' coordinates and features read from global
// memory, two entries in flight.
template <int kF>  // register budget for the per-voxel features
__global__ void __launch_bounds__(128, 4) k_columns_big(SemMapCfg c, const float* __restrict__ obs,
                                                 const float* __restrict__ coords, const int* __restrict__ col_start,
                                                 const uint32_t* __restrict__ entries, const int* __restrict__ col_list,
                                                 const int* __restrict__ list_n, float* __restrict__ ego) {
  pdl_grid_sync();
  extern __shared__ uint32_t keys[];
  __shared__ float red_all[kF], red_agent[kF];
  __shared__ uint32_t zbits[4];  // z cells (0 .. nz, biased) that occur in the column's keys
  const int e = blockIdx.y;
  const int ncols = c.vr * c.vr;
  const int N = c.h * c.w;
  const int* start = col_start + static_cast<size_t>(e) * (ncols + 1);
  const int* list = col_list + static_cast<size_t>(e) * 2 * ncols;
  const int n_list = list_n[e * 4 + 1];
  const int tid = threadIdx.x;
  float* ego_e = ego + static_cast<size_t>(e) * c.ego_channels * ncols;
  for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
  const int col = list[ncols - 1 - li];
  const int s0 = start[col];
  const int n = start[col + 1] - s0;
  const int px = col / c.vr, py = col - px * c.vr;
  const int cell = py * c.vr + px;  // voxels.transpose(2,3): row = y index, column = x index
  __syncthreads();  // the previous column's keys / partial sums are no longer read
  if (tid < 4) zbits[tid] = 0u;
  if (tid < kF) red_all[tid] = 0.f, red_agent[tid] = 0.f;
  const uint32_t* ent = entries + static_cast<size_t>(e) * 4 * N + s0;
  {
    // load + bitonic sort (ascending) of the column's keys
    int npow = 1;
    while (npow < n) npow <<= 1;
    __syncthreads();
    for (int i = tid; i < npow; i += blockDim.x) {
      const uint32_t k = i < n ? ent[i] : 0xffffffffu;
      keys[i] = k;
      if (i < n) atomicOr(&zbits[(k >> 20) & 3u], 1u << ((k >> 15) & 31u));
    }
    __syncthreads();
    for (int k = 2; k <= npow; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < npow; i += blockDim.x) {
          const int l = i ^ j;
          if (l > i) {
            const uint32_t a = keys[i], b = keys[l];
            const bool up = (i & k) == 0;
            if ((a > b) == up) keys[i] = b, keys[l] = a;
          }
        }
        __syncthreads();
      }
    }
  }

  // thread z replays the accumulation of voxel (px, py, z)
  const int z = tid;
  const int nf = c.nf;
  float acc[kF];
#pragma unroll
  for (int f = 0; f < kF; ++f) acc[f] = 0.f;
  // voxel z receives entries whose lower z cell (biased by +1 in the key) is z + 1 (corner iz = 0) or z (iz = 1)
  const bool touched = z > 0 && z < c.nz &&
                       (((zbits[(z + 1) >> 5] >> ((z + 1) & 31)) | (zbits[z >> 5] >> (z & 31))) & 1u) != 0u;
  if (touched) {
    const float* cx = coords + static_cast<size_t>(e) * 3 * N;
    const float* feat = obs + (static_cast<size_t>(e) * c.channels + 4) * N;  // semantic channels
    for (int ixy = 0; ixy < 4; ++ixy) {
      const int ix = ixy >> 1, iy = ixy & 1;
      for (int iz = 0; iz < 2; ++iz) {
        // entries of this lateral corner whose lower z cell is z - iz  (biased by +1 in the key)
        const uint32_t kbase = (static_cast<uint32_t>(ixy) << 22) | (static_cast<uint32_t>(z - iz + 1) << 15);
        const int lo = lower_bound_key(keys, n, kbase);
        const int hi = lower_bound_key(keys, n, kbase + (1u << 15));
        // two entries in flight (their coordinate and feature loads are independent of the running sums); the adds
        // stay in entry order, as the reference's index_add does
        for (int t0 = lo; t0 < hi; t0 += 2) {
          int idx[2];
          float cxv[2], cyv[2], czv[2], wv[2];
          float fv[2][kF];
#pragma unroll
          for (int u = 0; u < 2; ++u) idx[u] = static_cast<int>(keys[min(t0 + u, hi - 1)] & 0x7fffu);
#pragma unroll
          for (int u = 0; u < 2; ++u) cxv[u] = cx[idx[u]], cyv[u] = cx[N + idx[u]], czv[u] = cx[2 * N + idx[u]];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
#pragma unroll
            for (int f = 1; f < kF; ++f) fv[u][f] = (f < nf) ? feat[static_cast<size_t>(f - 1) * N + idx[u]] : 0.f;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            int p;
            float wx, wy, wz;
            bool s;
            corner(cxv[u], c.vr_f, ix, p, wx, s);
            corner(cyv[u], c.vr_f, iy, p, wy, s);
            corner(czv[u], c.nz_f, iz, p, wz, s);
            wv[u] = ((1.f * wx) * wy) * wz;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (t0 + u < hi) {
              acc[0] = acc[0] + 1.f * wv[u];
#pragma unroll
              for (int f = 1; f < kF; ++f) {
                if (f < nf) acc[f] = acc[f] + fv[u][f] * wv[u];
              }
            }
          }
        }
        // grid_flat = torch.round(grid_flat) after every corner (depth_utils.py:250): half-to-even
#pragma unroll
        for (int f = 0; f < kF; ++f) acc[f] = rintf(acc[f]);
      }
    }
  }
  // height projections (mapping.py:102-113): sums of integer-valued floats - exact, hence order-free
  if (touched) {
    const bool in_agent = (z >= c.min_z && z < c.max_z);
#pragma unroll
    for (int f = 0; f < kF; ++f) {
      if (f < nf && acc[f] != 0.f) {
        atomicAdd(&red_all[f], acc[f]);
        if (in_agent) atomicAdd(&red_agent[f], acc[f]);
      }
    }
  }
  __syncthreads();
  if (tid < c.ego_channels) {
    // ego channel 0 = obstacle, 1 = explored, 2.. = categories (local-map channels 4..)
    const int ch = tid;
    const int f = ch < 2 ? 0 : ch - 1;
    const float all_h = red_all[f], agent_h = red_agent[f];
    float v;
    if (ch == 0) v = agent_h / c.map_thr;
    else if (ch == 1) v = all_h / c.exp_thr;
    else {
      const bool use_all = (f == c.special_f[0] || f == c.special_f[1] || f == c.special_f[2]);
      v = (use_all ? all_h : agent_h) / c.cat_thr;
    }
    ego_e[static_cast<size_t>(ch) * ncols + cell] = fminf(fmaxf(v, 0.f), 1.f);
  }
  }  // column loop
}

// --------------------------------------------------------------------------------------------------
// k_pose: get_new_pose_batch (mapping.py:143-160) in place + sampling-grid parameters (model.py:7-43).
__global__ void k_pose(SemMapCfg c, int E, const float* __restrict__ pose_delta, float* __restrict__ poses,
                       float* __restrict__ xf) {
  pdl_grid_sync();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float x = poses[e * 3], y = poses[e * 3 + 1], t = poses[e * 3 + 2];
  const float dx = pose_delta[e * 3], dy = pose_delta[e * 3 + 1], dt = pose_delta[e * 3 + 2];
  const float tr = t / 57.29577951308232f;
  y = y + (dx * sinf(tr) + dy * cosf(tr));
  x = x + (dx * cosf(tr) - dy * sinf(tr));
  t = t + dt * 57.29577951308232f;
  t = fmodf(t - 180.0f, 360.0f) + 180.0f;
  t = fmodf(t + 180.0f, 360.0f) - 180.0f;
  poses[e * 3] = x, poses[e * 3 + 1] = y, poses[e * 3 + 2] = t;
  const float half = static_cast<float>(c.map_cells / 2);
  const float sx = -(x * 100.0f / c.res - half) / half;
  const float sy = -(y * 100.0f / c.res - half) / half;
  const float st = (90.f - t) * 3.14159265358979323846f / 180.f;
  xf[e * 4] = cosf(st), xf[e * 4 + 1] = sinf(st), xf[e * 4 + 2] = sx, xf[e * 4 + 3] = sy;
}

// F.affine_grid base coordinate (align_corners=False): linspace(-1, 1, n)[i] * (n - 1) / n
__device__ __forceinline__ float base_coord(int i, int n) {
  const float step = 2.f / static_cast<float>(n - 1);
  const float l = (i < n / 2) ? (-1.f + step * static_cast<float>(i)) : (1.f - step * static_cast<float>(n - i - 1));
  return l * static_cast<float>(n - 1) / static_cast<float>(n);
}

// --------------------------------------------------------------------------------------------------
// k_fuse: translated = grid_sample(grid_sample(agent_view, rot), trans); map = max(maps_last, translated).
// One thread per local-map cell.  The 100 x 100 ego window is the only non-zero part of agent_view, and only cells whose
// sampling position lies within fuse_r of the map centre can reach it (the first sampler rotates about the centre;
// SemMap::init derives the radius with a two-cell margin): everything else - four cells in five - is a plain
// max(maps_last, 0).  For the others the sixteen taps of the two chained samplers are walked once, tap outer / channel
// inner, with the kE ego channels' sums in registers (same products, same order per channel as a tap list would give;
// round 2's first version kept the taps in local-memory arrays and re-read them per channel: 1 500 instructions per cell).
template <int kE>
__global__ void __launch_bounds__(256) k_fuse(SemMapCfg c, const float* __restrict__ xf, const float* __restrict__ ego,
                                              const float* __restrict__ maps_last, long long ml_env, long long ml_plane,
                                              long long ml_row, float* __restrict__ map_out, float* __restrict__ fp_out) {
  pdl_grid_sync();
  const int e = blockIdx.z;
  const int n = c.map_cells;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= n) return;
  const int ncols = c.vr * c.vr;
  const float tx = xf[e * 4 + 2], ty = xf[e * 4 + 3];
  const float* ego_e = ego + static_cast<size_t>(e) * c.ego_channels * ncols;
  // fp_map_pred output = obstacle channel of the ego window
  if (fp_out != nullptr && y < c.vr && x < c.vr) fp_out[(static_cast<size_t>(e) * c.vr + y) * c.vr + x] = ego_e[y * c.vr + x];

  // second sampler: where does output cell (y, x) read `rotated`?
  const float bx = base_coord(x, n), by = base_coord(y, n);
  const float gx2 = bx * 1.f + by * (-0.f) + tx;
  const float gy2 = bx * 0.f + by * 1.f + ty;
  const float fx2 = ((gx2 + 1.f) / 2.f) * static_cast<float>(n - 1);
  const float fy2 = ((gy2 + 1.f) / 2.f) * static_cast<float>(n - 1);
  const float ctr = 0.5f * static_cast<float>(n - 1);
  const float ddx = fx2 - ctr, ddy = fy2 - ctr;
  const bool near = ddx * ddx + ddy * ddy <= c.fuse_r2;
  float v[kE];
#pragma unroll
  for (int k = 0; k < kE; ++k) v[k] = 0.f;
  if (near) {
    const float cs = xf[e * 4], sn = xf[e * 4 + 1];
    const int wx1 = n / 2 - c.vr / 2, wy1 = n / 2;  // ego window origin (mapping.py:127-130)
    const float x2f = floorf(fx2), y2f = floorf(fy2);
#pragma unroll
    for (int cy = 0; cy < 2; ++cy) {
#pragma unroll
      for (int cxi = 0; cxi < 2; ++cxi) {
        const float qxf = x2f + cxi, qyf = y2f + cy;
        const float w2 = (cxi ? (fx2 - x2f) : (x2f + 1.f - fx2)) * (cy ? (fy2 - y2f) : (y2f + 1.f - fy2));
        if (!(qxf >= 0.f && qxf <= static_cast<float>(n - 1) && qyf >= 0.f && qyf <= static_cast<float>(n - 1))) continue;
        const int qx = static_cast<int>(qxf), qy = static_cast<int>(qyf);
        // first sampler: rotated(qy, qx) reads agent_view at
        const float rbx = base_coord(qx, n), rby = base_coord(qy, n);
        const float gx1 = rbx * cs + rby * (-sn) + 0.f;
        const float gy1 = rbx * sn + rby * cs + 0.f;
        const float fx1 = ((gx1 + 1.f) / 2.f) * static_cast<float>(n - 1);
        const float fy1 = ((gy1 + 1.f) / 2.f) * static_cast<float>(n - 1);
        const float x1f = floorf(fx1), y1f = floorf(fy1);
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
          for (int dx = 0; dx < 2; ++dx) {
            const int ax = static_cast<int>(x1f) + dx - wx1, ay = static_cast<int>(y1f) + dy - wy1;
            const float w1 = (dx ? (fx1 - x1f) : (x1f + 1.f - fx1)) * (dy ? (fy1 - y1f) : (y1f + 1.f - fy1));
            if (ax >= 0 && ax < c.vr && ay >= 0 && ay < c.vr) {
              const float w = w1 * w2;
              const float* src = ego_e + ay * c.vr + ax;
#pragma unroll
              for (int k = 0; k < kE; ++k) {
                if (k < c.ego_channels) v[k] += src[static_cast<size_t>(k) * ncols] * w;
              }
            }
          }
        }
      }
    }
  }
  const size_t plane = static_cast<size_t>(n) * n;
  const size_t pix = static_cast<size_t>(y) * n + x;
  const float* ml = maps_last + static_cast<size_t>(e) * ml_env + static_cast<size_t>(y) * ml_row + x;
  float* mo = map_out + static_cast<size_t>(e) * c.channels * plane + pix;
  // the map is streamed once (read maps_last, write map_out): keep eight channel loads in flight per thread;
  // map channel 0, 1 <- ego channel 0, 1; channels 2, 3 (agent location) receive nothing; channel ch >= 4 <- ego ch - 2
#pragma unroll
  for (int ch0 = 0; ch0 < kE + 2; ch0 += 8) {
    if (ch0 >= c.channels) break;
    float last[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) last[u] = (ch0 + u < c.channels) ? __ldg(ml + (ch0 + u) * ml_plane) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int ch = ch0 + u;
      if (ch < c.channels) {
        const int k = ch < 2 ? ch : ch - 2;
        const float val = (ch == 2 || ch == 3 || k >= kE) ? 0.f : v[k];
        mo[ch * plane] = fmaxf(last[u], val);
      }
    }
  }
}

}  // namespace

// --------------------------------------------------------------------------------------------------
void SemMap::init(const SemMapCfg& cfg, int envs) {
  c = cfg;
  E = envs;
  {
    // k_fuse's reach test: farthest ego-window cell from the map centre (pixel units), through the first sampler (a rotation
    // about the centre, base coordinates scaled by (n - 1) / n) and the two bilinear footprints (sqrt(2) each), plus 2 cells
    const double n = c.map_cells, ctr = 0.5 * (n - 1);
    const double wx1 = c.map_cells / 2 - c.vr / 2, wy1 = c.map_cells / 2;
    const double hx = std::max(std::fabs(wx1 - ctr), std::fabs(wx1 + c.vr - 1 - ctr));
    const double hy = std::max(std::fabs(wy1 - ctr), std::fabs(wy1 + c.vr - 1 - ctr));
    const double r = (std::sqrt(hx * hx + hy * hy) + std::sqrt(2.0)) * n / (n - 1) + std::sqrt(2.0) + 2.0;
    c.fuse_r2 = static_cast<float>(r * r);
  }
  PN_REQUIRE(c.nf <= kMaxFeat, "semmap: too many semantic categories");
  PN_REQUIRE(c.nz <= 126 && c.h * c.w <= 32767, "semmap: geometry out of range");
  const size_t N = static_cast<size_t>(c.h) * c.w;
  const size_t ncols = static_cast<size_t>(c.vr) * c.vr;
  coords = static_cast<float*>(arena.alloc(E * 3 * N * sizeof(float)));
  col_count = static_cast<int*>(arena.alloc((E * ncols + 4 * E) * sizeof(int)));  // [E][vr*vr] + qcount[E][4]
  qcount = reinterpret_cast<uint32_t*>(col_count + E * ncols);
  col_start = static_cast<int*>(arena.alloc(E * (ncols + 1) * sizeof(int)));
  col_fill = static_cast<int*>(arena.alloc(E * ncols * sizeof(int)));
  entries = static_cast<uint32_t*>(arena.alloc(E * 4 * N * sizeof(uint32_t)));
  ego = static_cast<float*>(arena.alloc(E * c.ego_channels * ncols * sizeof(float)));
  col_list = static_cast<int*>(arena.alloc(E * 2 * ncols * sizeof(int)));
  list_n = static_cast<int*>(arena.alloc(E * 4 * sizeof(int)));
  int dev = 0;
  PN_CUDA_CHECK(cudaGetDevice(&dev));
  PN_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  xf = static_cast<float*>(arena.alloc(E * 4 * sizeof(float)));
  stair_flag = static_cast<int*>(arena.alloc(E * sizeof(int)));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_columns_big<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_columns_big<kMaxFeat>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_columns_cta<16, 12, kMidCap, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sizeof(ColumnStage<kMidCap, 12>))));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_columns_cta<32, kMaxFeat, kMidCap, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sizeof(ColumnStage<kMidCap, kMaxFeat>))));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_columns_cta<16, 12, kLargeCap, 3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sizeof(ColumnStage<kLargeCap, 12>))));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_columns_cta<32, kMaxFeat, kLargeCap, 3, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(sizeof(ColumnStage<kLargeCap, kMaxFeat>))));
  PN_REQUIRE(ncols <= 1024 * static_cast<size_t>(kScanPer), "semmap: vision_range too large for k_scan");
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 * kScanPer * static_cast<int>(sizeof(int))));
}

void SemMap::forward(const float* obs, const float* pose_delta, const float* maps_last, long long ml_env,
                     long long ml_plane, long long ml_row, float* poses_inout, float* fp_out, float* map_out,
                     cudaStream_t s) {
  const int N = c.h * c.w;
  const int ncols = c.vr * c.vr;
  PN_CUDA_CHECK(cudaMemsetAsync(col_count, 0, (static_cast<size_t>(E) * ncols + 4 * E) * sizeof(int), s));  // + qcount
  PN_CUDA_CHECK(cudaMemsetAsync(ego, 0, static_cast<size_t>(E) * c.ego_channels * ncols * sizeof(float), s));
  launch_pdl(k_coords, dim3((N + 255) / 256, E), 256, 0, s, c, obs, coords, qcount);
  launch_pdl(k_quantile, E, 1024, 0, s, c, coords, qcount, stair_flag);
  launch_pdl(k_hist, dim3((N + 255) / 256, E), 256, 0, s, c, obs, coords, stair_flag, col_count);
  launch_pdl(k_scan, E, 1024, static_cast<size_t>(ncols) * sizeof(int), s, ncols, col_count, col_start, col_fill, col_list, list_n);
  launch_pdl(k_fill, dim3((N + 255) / 256, E), 256, 0, s, c, coords, col_start, col_fill, entries);
  // non-empty columns only, persistent over four work lists (see k_scan): a warp per tiny column, a CTA per mid / large
  // column with its entries staged in shared memory, a CTA with 128 KB of key storage per big column
  const int g_warp = std::min((ncols + 3) / 4, num_sms * 16), g_cta = std::min(ncols, num_sms * 6), g_big = std::min(ncols, num_sms);
  if (c.nf <= 12) {
    launch_pdl(k_columns_warp<16, 12>, dim3(g_warp, E), 128, 0, s, c, obs, coords, col_start, entries, col_list, list_n, ego);
    launch_pdl(k_columns_cta<16, 12, kMidCap, 2, 8>, dim3(g_cta, E), 256, sizeof(ColumnStage<kMidCap, 12>), s, c, obs, coords, col_start,
               entries, col_list, list_n, ego);
    launch_pdl(k_columns_cta<16, 12, kLargeCap, 3, 16>, dim3(g_big, E), 512, sizeof(ColumnStage<kLargeCap, 12>), s, c, obs, coords,
               col_start, entries, col_list, list_n, ego);
    launch_pdl(k_columns_big<12>, dim3(g_big, E), 128, 32768 * 4, s, c, obs, coords, col_start, entries, col_list, list_n, ego);
  } else {
    launch_pdl(k_columns_warp<32, kMaxFeat>, dim3(g_warp, E), 128, 0, s, c, obs, coords, col_start, entries, col_list, list_n, ego);
    launch_pdl(k_columns_cta<32, kMaxFeat, kMidCap, 2, 8>, dim3(g_cta, E), 256, sizeof(ColumnStage<kMidCap, kMaxFeat>), s, c, obs, coords,
               col_start, entries, col_list, list_n, ego);
    launch_pdl(k_columns_cta<32, kMaxFeat, kLargeCap, 3, 16>, dim3(g_big, E), 512, sizeof(ColumnStage<kLargeCap, kMaxFeat>), s, c, obs,
               coords, col_start, entries, col_list, list_n, ego);
    launch_pdl(k_columns_big<kMaxFeat>, dim3(g_big, E), 128, 32768 * 4, s, c, obs, coords, col_start, entries, col_list, list_n, ego);
  }
  launch_pdl(k_pose, (E + 63) / 64, 64, 0, s, c, E, pose_delta, poses_inout, xf);
  if (c.ego_channels <= 12)
    launch_pdl(k_fuse<12>, dim3((c.map_cells + 255) / 256, c.map_cells, E), 256, 0, s, c, xf, ego, maps_last, ml_env, ml_plane, ml_row, map_out, fp_out);
  else
    launch_pdl(k_fuse<kMaxFeat + 1>, dim3((c.map_cells + 255) / 256, c.map_cells, E), 256, 0, s, c, xf, ego, maps_last, ml_env, ml_plane, ml_row, map_out, fp_out);
  PN_CUDA_CHECK(cudaGetLastError());
}

}  // namespace pn
