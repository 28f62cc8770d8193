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
// (qcount[e] = {#valid heights, #heights in the stair band}).  grid = (ceil(N / 256), E).
__global__ void __launch_bounds__(256) k_coords(SemMapCfg c, const float* __restrict__ obs, float* __restrict__ coords,
                                                uint32_t* __restrict__ qcount) {
  pdl_grid_sync();
  const int e = blockIdx.y;
  const int N = c.h * c.w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float* depth = obs + (static_cast<size_t>(e) * c.channels + 3) * N;
  float* cx = coords + static_cast<size_t>(e) * 3 * N;
  uint32_t valid = 0, mid = 0;
  if (i < N) {
    float X, Y, Z;
    point_coords(c, i / c.w, i % c.w, depth[i], X, Y, Z);
    cx[i] = X, cx[N + i] = Y, cx[2 * N + i] = Z;
    if (Z > -1.f && Z < 1.f) {
      const float mz = Z * 2.f + 1.6f;
      valid = 1;
      mid = (mz > 0.2f && mz < 0.7f) ? 1u : 0u;
    }
  }
  const uint32_t nv = __popc(__ballot_sync(0xffffffffu, valid != 0)), nm = __popc(__ballot_sync(0xffffffffu, mid != 0));
  if ((threadIdx.x & 31) == 0 && nv) {
    atomicAdd(&qcount[e * 2], nv);
    if (nm) atomicAdd(&qcount[e * 2 + 1], nm);
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
  const uint32_t n = qcount[e * 2];
  const uint32_t s_mid = qcount[e * 2 + 1];

  // torch.quantile(my_zs, 0.03), linear interpolation: ranks = q*(n-1) in fp32 (mapping.py:94)
  bool flag = false;
  if (n > 0) {
    const float rank = 0.03f * static_cast<float>(n - 1);
    const float rb = floorf(rank);
    const uint32_t k_lo = static_cast<uint32_t>(rb);
    const uint32_t k_hi = static_cast<uint32_t>(ceilf(rank));
    const float wgt = rank - rb;
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
// k_scan: exclusive prefix sum of the column counters; also resets the fill cursors.  grid = E, block = 1024.
// It also compacts the non-empty columns into a work list for k_columns: columns with at most `small_cap` entries
// from the front of col_list, larger ones from the back (list_n = {#small, #large}); the order inside the list is
// irrelevant, every column is processed independently.
__global__ void __launch_bounds__(1024) k_scan(int ncols, int small_cap, const int* __restrict__ col_count,
                                               int* __restrict__ col_start, int* __restrict__ col_fill,
                                               int* __restrict__ col_list, int* __restrict__ list_n) {
  pdl_grid_sync();
  const int e = blockIdx.x;
  int* list = col_list + static_cast<size_t>(e) * ncols;
  __shared__ int n_small, n_large;
  if (threadIdx.x == 0) n_small = 0, n_large = 0;
  const int* cnt = col_count + static_cast<size_t>(e) * ncols;
  int* start = col_start + static_cast<size_t>(e) * (ncols + 1);
  int* fill = col_fill + static_cast<size_t>(e) * ncols;
  __shared__ int warp_sums[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < ncols; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < ncols ? cnt[i] : 0;
    int x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
      int s = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, s, o);
        if (lane >= o) s += y;
      }
      warp_sums[lane] = s;
    }
    __syncthreads();
    const int excl = carry + (wid ? warp_sums[wid - 1] : 0) + x - v;
    if (i < ncols) {
      start[i] = excl;
      fill[i] = 0;
      if (v > 0) {
        if (v <= small_cap) list[atomicAdd(&n_small, 1)] = i;
        else list[ncols - 1 - atomicAdd(&n_large, 1)] = i;
      }
    }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    start[ncols] = carry;
    list_n[e * 2] = n_small, list_n[e * 2 + 1] = n_large;
  }
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
// k_columns: one CTA (128 threads) per (x,y) column.  `cap` = keys that fit the dynamic smem of this launch;
// columns with more entries are left to the large-capacity launch (and vice versa).
__device__ __forceinline__ int lower_bound_key(const uint32_t* keys, int n, uint32_t k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

template <int kF>  // register budget for the per-voxel features: 12 (the reference's 10 categories + count) or kF
__global__ void __launch_bounds__(128, 4) k_columns(SemMapCfg c, int large, const float* __restrict__ obs,
                                                 const float* __restrict__ coords, const int* __restrict__ col_start,
                                                 const uint32_t* __restrict__ entries, const int* __restrict__ col_list,
                                                 const int* __restrict__ list_n, float* __restrict__ ego) {
  pdl_grid_sync();
  extern __shared__ uint32_t keys[];
  __shared__ float red_all[kF], red_agent[kF];
  __shared__ uint32_t raw[128];
  // small columns: per-entry lateral weight, both z weights and the feature vector, staged by one thread per entry
  __shared__ float e_wxy[128], e_wz0[128], e_wz1[128], e_feat[128][kF];
  __shared__ uint32_t zbits[4];  // z cells (0 .. nz, biased) that occur in the column's keys
  const int e = blockIdx.y;
  const int ncols = c.vr * c.vr;
  const int N = c.h * c.w;
  const int* start = col_start + static_cast<size_t>(e) * (ncols + 1);
  const int* list = col_list + static_cast<size_t>(e) * ncols;
  const int n_list = list_n[e * 2 + large];
  const int tid = threadIdx.x;
  float* ego_e = ego + static_cast<size_t>(e) * c.ego_channels * ncols;
  // persistent over this launch's share of the non-empty columns (empty ones keep the zeros of the memset)
  for (int li = blockIdx.x; li < n_list; li += gridDim.x) {
  const int col = large ? list[ncols - 1 - li] : list[li];
  const int s0 = start[col];
  const int n = start[col + 1] - s0;
  const int px = col / c.vr, py = col - px * c.vr;
  const int cell = py * c.vr + px;  // voxels.transpose(2,3): row = y index, column = x index
  __syncthreads();  // the previous column's keys / partial sums are no longer read
  if (tid < 4) zbits[tid] = 0u;
  if (tid < kF) red_all[tid] = 0.f, red_agent[tid] = 0.f;
  const uint32_t* ent = entries + static_cast<size_t>(e) * 4 * N + s0;
  if (n <= 128) {
    // small column (the common case): rank sort - keys are unique (they embed the point index), so the number of
    // smaller keys is the sorted position; two barriers instead of the ~log^2 n of the bitonic network
    if (tid < n) raw[tid] = ent[tid];
    __syncthreads();
    if (tid < n) {
      const uint32_t k = raw[tid];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += raw[j] < k ? 1 : 0;
      keys[rank] = k;
      atomicOr(&zbits[(k >> 20) & 3u], 1u << ((k >> 15) & 31u));  // z cell (7 bits at 15..21) seen in this column
      // this entry's weights and features, once, in parallel over the column's entries (the voxel threads below would
      // otherwise each walk a chain of dependent global loads per entry)
      const float* cxp = coords + static_cast<size_t>(e) * 3 * N;
      const float* featp = obs + (static_cast<size_t>(e) * c.channels + 4) * N;
      const int i = static_cast<int>(k & 0x7fffu);
      const int ixy = static_cast<int>(k >> 22);
      int pp;
      float wx, wy, wz0, wz1;
      bool ss;
      corner(cxp[i], c.vr_f, ixy >> 1, pp, wx, ss);
      corner(cxp[N + i], c.vr_f, ixy & 1, pp, wy, ss);
      const float zc = cxp[2 * N + i];
      corner(zc, c.nz_f, 0, pp, wz0, ss);
      corner(zc, c.nz_f, 1, pp, wz1, ss);
      e_wxy[rank] = (1.f * wx) * wy;
      e_wz0[rank] = wz0, e_wz1[rank] = wz1;
#pragma unroll
      for (int f = 1; f < kF; ++f) e_feat[rank][f] = (f < c.nf) ? featp[static_cast<size_t>(f - 1) * N + i] : 0.f;
    }
    __syncthreads();
  } else {
    // load + bitonic sort (ascending) of the column's keys
    int npow = 1;
    while (npow < n) npow <<= 1;
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
        if (n <= 128) {  // staged column: everything is in shared memory
          for (int t = lo; t < hi; ++t) {
            const float w = e_wxy[t] * (iz ? e_wz1[t] : e_wz0[t]);
            acc[0] = acc[0] + 1.f * w;
#pragma unroll
            for (int f = 1; f < kF; ++f) {
              if (f < nf) acc[f] = acc[f] + e_feat[t][f] * w;
            }
          }
        } else
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
// One thread per local-map cell; the 100 x 100 ego window is the only non-zero part of agent_view.
__global__ void __launch_bounds__(256) k_fuse(SemMapCfg c, const float* __restrict__ xf, const float* __restrict__ ego,
                                              const float* __restrict__ maps_last, long long ml_env, long long ml_plane,
                                              long long ml_row, float* __restrict__ map_out,
                                              float* __restrict__ fp_out) {
  pdl_grid_sync();
  const int e = blockIdx.z;
  const int n = c.map_cells;
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= n) return;
  const int ncols = c.vr * c.vr;
  const float cs = xf[e * 4], sn = xf[e * 4 + 1], tx = xf[e * 4 + 2], ty = xf[e * 4 + 3];
  const float* ego_e = ego + static_cast<size_t>(e) * c.ego_channels * ncols;
  const int wx1 = n / 2 - c.vr / 2, wy1 = n / 2;  // ego window origin (mapping.py:127-130)

  // fp_map_pred output = obstacle channel of the ego window
  if (fp_out != nullptr && y < c.vr && x < c.vr) fp_out[(static_cast<size_t>(e) * c.vr + y) * c.vr + x] = ego_e[y * c.vr + x];

  // second sampler: where does output cell (y, x) read `rotated`?
  const float bx = base_coord(x, n), by = base_coord(y, n);
  const float gx2 = bx * 1.f + by * (-0.f) + tx;
  const float gy2 = bx * 0.f + by * 1.f + ty;
  const float fx2 = ((gx2 + 1.f) / 2.f) * static_cast<float>(n - 1);
  const float fy2 = ((gy2 + 1.f) / 2.f) * static_cast<float>(n - 1);
  const float x2f = floorf(fx2), y2f = floorf(fy2);
  // taps[k]: (ego offset or -1, weight) accumulated per channel below
  int tap_idx[16];
  float tap_w[16];
  int ntaps = 0;
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
            tap_idx[ntaps] = ay * c.vr + ax;
            tap_w[ntaps] = w1 * w2;
            ++ntaps;
          }
        }
      }
    }
  }
  const size_t plane = static_cast<size_t>(n) * n;
  const size_t pix = static_cast<size_t>(y) * n + x;
  const float* ml = maps_last + static_cast<size_t>(e) * ml_env + static_cast<size_t>(y) * ml_row + x;
  float* mo = map_out + static_cast<size_t>(e) * c.channels * plane + pix;
  // the map is streamed once (read maps_last, write map_out): keep eight channel loads in flight per thread
  for (int ch0 = 0; ch0 < c.channels; ch0 += 8) {
    float last[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) last[u] = (ch0 + u < c.channels) ? __ldg(ml + (ch0 + u) * ml_plane) : 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int ch = ch0 + u;
      if (ch >= c.channels) break;
      float v = 0.f;
      if (ntaps > 0 && ch != 2 && ch != 3) {
        const float* src = ego_e + static_cast<size_t>(ch < 2 ? ch : ch - 2) * ncols;
        for (int k = 0; k < ntaps; ++k) v += src[tap_idx[k]] * tap_w[k];
      }
      mo[ch * plane] = fmaxf(last[u], v);
    }
  }
}

}  // namespace

// --------------------------------------------------------------------------------------------------
void SemMap::init(const SemMapCfg& cfg, int envs) {
  c = cfg;
  E = envs;
  PN_REQUIRE(c.nf <= kMaxFeat, "semmap: too many semantic categories");
  PN_REQUIRE(c.nz <= 126 && c.h * c.w <= 32767, "semmap: geometry out of range");
  const size_t N = static_cast<size_t>(c.h) * c.w;
  const size_t ncols = static_cast<size_t>(c.vr) * c.vr;
  coords = static_cast<float*>(arena.alloc(E * 3 * N * sizeof(float)));
  col_count = static_cast<int*>(arena.alloc((E * ncols + 2 * E) * sizeof(int)));  // [E][vr*vr] + qcount[E][2]
  qcount = reinterpret_cast<uint32_t*>(col_count + E * ncols);
  col_start = static_cast<int*>(arena.alloc(E * (ncols + 1) * sizeof(int)));
  col_fill = static_cast<int*>(arena.alloc(E * ncols * sizeof(int)));
  entries = static_cast<uint32_t*>(arena.alloc(E * 4 * N * sizeof(uint32_t)));
  ego = static_cast<float*>(arena.alloc(E * c.ego_channels * ncols * sizeof(float)));
  col_list = static_cast<int*>(arena.alloc(E * ncols * sizeof(int)));
  list_n = static_cast<int*>(arena.alloc(E * 2 * sizeof(int)));
  int dev = 0;
  PN_CUDA_CHECK(cudaGetDevice(&dev));
  PN_CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  xf = static_cast<float*>(arena.alloc(E * 4 * sizeof(float)));
  stair_flag = static_cast<int*>(arena.alloc(E * sizeof(int)));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_columns<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_columns<kMaxFeat>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768 * 4));
}

void SemMap::forward(const float* obs, const float* pose_delta, const float* maps_last, long long ml_env,
                     long long ml_plane, long long ml_row, float* poses_inout, float* fp_out, float* map_out,
                     cudaStream_t s) {
  const int N = c.h * c.w;
  const int ncols = c.vr * c.vr;
  PN_CUDA_CHECK(cudaMemsetAsync(col_count, 0, (static_cast<size_t>(E) * ncols + 2 * E) * sizeof(int), s));  // + qcount
  PN_CUDA_CHECK(cudaMemsetAsync(ego, 0, static_cast<size_t>(E) * c.ego_channels * ncols * sizeof(float), s));
  launch_pdl(k_coords, dim3((N + 255) / 256, E), 256, 0, s, c, obs, coords, qcount);
  launch_pdl(k_quantile, E, 1024, 0, s, c, coords, qcount, stair_flag);
  launch_pdl(k_hist, dim3((N + 255) / 256, E), 256, 0, s, c, obs, coords, stair_flag, col_count);
  constexpr int kSmallCap = 2048;
  launch_pdl(k_scan, E, 1024, 0, s, ncols, kSmallCap, col_count, col_start, col_fill, col_list, list_n);
  launch_pdl(k_fill, dim3((N + 255) / 256, E), 256, 0, s, c, coords, col_start, col_fill, entries);
  // non-empty columns only, persistent CTAs: 8 KB of key storage each for the small ones, 128 KB for the rare
  // columns that collect more than kSmallCap entries (a wall seen edge-on)
  const int g_small = std::min(ncols, num_sms * 8), g_large = std::min(ncols, num_sms);
  if (c.nf <= 12) {
    launch_pdl(k_columns<12>, dim3(g_small, E), 128, kSmallCap * 4, s, c, 0, obs, coords, col_start, entries, col_list, list_n, ego);
    launch_pdl(k_columns<12>, dim3(g_large, E), 128, 32768 * 4, s, c, 1, obs, coords, col_start, entries, col_list, list_n, ego);
  } else {
    launch_pdl(k_columns<kMaxFeat>, dim3(g_small, E), 128, kSmallCap * 4, s, c, 0, obs, coords, col_start, entries, col_list, list_n, ego);
    launch_pdl(k_columns<kMaxFeat>, dim3(g_large, E), 128, 32768 * 4, s, c, 1, obs, coords, col_start, entries, col_list, list_n, ego);
  }
  launch_pdl(k_pose, (E + 63) / 64, 64, 0, s, c, E, pose_delta, poses_inout, xf);
  launch_pdl(k_fuse, dim3((c.map_cells + 255) / 256, c.map_cells, E), 256, 0, s, c, xf, ego, maps_last, ml_env, ml_plane, ml_row, map_out, fp_out);
  PN_CUDA_CHECK(cudaGetLastError());
}

}  // namespace pn
