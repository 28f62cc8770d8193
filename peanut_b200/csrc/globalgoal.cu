// Agent_State.update_global_goal on the device (nav/agent/agent_state.py:376-416) - SURVEY.md section 8(f), N1b.
//
//   trav  = not binary_dilation(rint(full_map[0]), disk(col_rad)); collision cells -> 0, visited cells -> 1     :382-386
//   dd    = skfmm.distance(masked(trav), dx=1) with the agent's cell as the zero level set                      :388-393
//   dd_wt = exp(-dd / (dist_weight_temperature / map_resolution)) inside the local-map window                   :395-399
//   value = target_pred * dd_wt ; goal = first argmax ; "avoid repeating the last goal"                         :401-416
//
// The reference does all of it on the host: full_map[0] leaves the device every prediction step and scikit-fmm's
// heap-ordered fast marching (one cell at a time, ~0.3 s for the 960 x 960 map) produces the geodesic field.  Here the map
// never leaves the device.  The field is the solution of the SAME discretisation (second-order upwind eikonal update of
// skfmm/distance_marcher.cpp: quadratic with 9/4 (T - (4 v1 - v2)/3)^2 terms where two upwind neighbours are known and
// v2 <= v1, first-order terms otherwise) computed by a block-parallel fixed-point iteration: 32 x 32 tiles with a 2-cell
// halo iterate in shared memory until nothing changes, tiles whose neighbours changed are re-activated, one persistent
// cooperative kernel walks the active tiles with a grid barrier per round.  Fast marching is not a pure function of the
// final field next to the source (the order in which equal-valued neighbours freeze decides who may use a second-order
// stencil across the source cell), so the first cells - everything fast marching freezes up to distance 3 - are produced by
// an exact sequential replay of the marcher (binary heap and all) on an 11 x 11 window, one thread per environment, and
// kept fixed.  Everything is fp64 like the reference's numpy arithmetic.  HBM traffic is irrelevant here (7 MB per field,
// L2-resident); the cost is the number of dependent rounds (~ path length / 32).
#include <cfloat>
#include <cmath>

#include <cooperative_groups.h>

#include "../../include/peanut_b200.h"
#include "engine.h"

namespace cg = cooperative_groups;

namespace pn {

namespace {

constexpr int kTile = 32, kHalo = 2, kSpan = kTile + 2 * kHalo;
constexpr int kWin = 11, kWinR = 5;     // replay window of the sequential marcher
constexpr double kSeedRadius = 3.0;     // cells the marcher freezes with |distance| <= this are replayed exactly
constexpr int kInnerMax = 64;

__device__ __forceinline__ double dinf() { return __longlong_as_double(0x7ff0000000000000ll); }

// ------------------------------------------------------------------------------------------------ traversible mask
// free[e][r][c] = 1 where the agent may walk.  grid (tiles_x, tiles_y, E), block 32 x 32; shared tile with a halo of
// `rad` cells (outside the map = no obstacle, scipy's border_value 0).
__global__ void __launch_bounds__(1024) k_traversible(const float* __restrict__ full_map, int nc, int W, int H, int rad,
                                                      const uint8_t* __restrict__ collision, const uint8_t* __restrict__ visited,
                                                      const int* __restrict__ lmb, const int* __restrict__ loc,
                                                      uint8_t* __restrict__ free_out, int* __restrict__ src_out) {
  extern __shared__ uint8_t obst[];  // [(32 + 2 rad)^2]
  const int e = blockIdx.z;
  const int span = kTile + 2 * rad;
  const int r0 = blockIdx.y * kTile - rad, c0 = blockIdx.x * kTile - rad;
  const float* plane = full_map + static_cast<size_t>(e) * nc * W * H;  // channel 0
  const int tid = threadIdx.y * kTile + threadIdx.x;
  for (int i = tid; i < span * span; i += kTile * kTile) {
    const int r = r0 + i / span, c = c0 + i % span;
    uint8_t o = 0;
    if (r >= 0 && r < W && c >= 0 && c < H) o = rint(static_cast<double>(plane[static_cast<size_t>(r) * H + c])) != 0.0 ? 1 : 0;
    obst[i] = o;
  }
  __syncthreads();
  const int r = blockIdx.y * kTile + threadIdx.y, c = blockIdx.x * kTile + threadIdx.x;
  // the agent's cell (clipped like np.clip, :389-390) is written by the thread that owns it
  const int ar = min(max(loc[e * 2] + lmb[e * 4], 0), W - 1), ac = min(max(loc[e * 2 + 1] + lmb[e * 4 + 2], 0), H - 1);
  if (r >= W || c >= H) return;
  bool dil = false;
  for (int dy = -rad; dy <= rad; ++dy)
    for (int dx = -rad; dx <= rad; ++dx)
      if (dy * dy + dx * dx <= rad * rad) dil |= obst[(threadIdx.y + rad + dy) * span + threadIdx.x + rad + dx] != 0;
  bool fr = !dil;
  const size_t idx = (static_cast<size_t>(e) * W + r) * H + c;
  if (collision != nullptr && collision[idx] == 1) fr = false;
  if (visited != nullptr && visited[idx] == 1) fr = true;
  if (r == ar && c == ac) {
    fr = true;  // assigning 0 to the masked array's element un-masks it (:389)
    src_out[e * 2] = ar, src_out[e * 2 + 1] = ac;
  }
  free_out[idx] = fr ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ sequential seed
// Exact replay of scikit-fmm's marcher (base_marcher.cpp / distance_marcher.cpp / heap.cpp, order 2, dx 1) on the
// kWin x kWin window around the source, stopped at the first popped value > kSeedRadius.  One thread per environment.
struct SeedMarcher {
  enum { kFar = 0, kNarrow = 1, kFrozen = 2, kMask = 3 };
  static constexpr int N = kWin * kWin;
  signed char flag[N];
  double dist[N], key[N];
  short heap[N], pos[N];
  int n;

  __device__ int nbr(int cur, int dim, int dir, int f) const {
    const int coord = dim == 0 ? cur / kWin : cur % kWin;
    const int nc = coord + dir;
    if (nc >= kWin || nc < 0) return -1;
    const int na = cur + dir * (dim == 0 ? kWin : 1);
    if (flag[na] == f) return -1;
    return na;
  }
  __device__ void swp(int a, int b) {
    const short ia = heap[a], ib = heap[b];
    heap[a] = ib, heap[b] = ia;
    pos[ib] = static_cast<short>(a), pos[ia] = static_cast<short>(b);
  }
  __device__ void up(int p) {
    while (p > 0) {
      const int parent = (p - 1) / 2;
      if (key[heap[p]] < key[heap[parent]]) swp(p, parent), p = parent;
      else break;
    }
  }
  __device__ void down(int p) {
    for (;;) {
      int c = 2 * p + 1;
      if (c >= n) break;
      if (c + 1 < n && key[heap[c + 1]] < key[heap[c]]) c += 1;
      if (key[heap[c]] < key[heap[p]]) swp(p, c), p = c;
      else break;
    }
  }
  __device__ void push(int a, double k) {
    key[a] = k, heap[n] = static_cast<short>(a), pos[a] = static_cast<short>(n);
    n += 1;
    up(n - 1);
  }
  __device__ void set(int a, double k) {
    const double old = key[a];
    key[a] = k;
    if (k < old) up(pos[a]);
    else down(pos[a]);
  }
  __device__ int pop(double* k) {
    const int a = heap[0];
    *k = key[a];
    n -= 1;
    if (n > 0) {
      heap[0] = heap[n];
      pos[heap[0]] = 0;
      down(0);
    }
    pos[a] = -1;
    return a;
  }
  // updatePointOrderTwo + solveQuadratic (phi > 0 everywhere but at the source: the '+' root); 0 = no update
  __device__ double update(int i) const {
    double a = 0, b = 0, c = 0;
    for (int dim = 0; dim < 2; ++dim) {
      double v1 = DBL_MAX, v2 = DBL_MAX;
      for (int j = -1; j < 2; j += 2) {
        const int na = nbr(i, dim, j, kMask);
        if (na != -1 && flag[na] == kFrozen && fabs(dist[na]) < fabs(v1)) {
          v1 = dist[na];
          const int na2 = nbr(i, dim, j * 2, kMask);
          if (na2 != -1 && flag[na2] == kFrozen && ((dist[na2] <= v1 && v1 >= 0) || (dist[na2] >= v1 && v1 <= 0))) v2 = dist[na2];
          else v2 = DBL_MAX;
        }
      }
      if (v2 < DBL_MAX) {
        const double tp = (1.0 / 3.0) * (4 * v1 - v2);
        a += 9.0 / 4.0, b -= 2 * (9.0 / 4.0) * tp, c += (9.0 / 4.0) * tp * tp;
      } else if (v1 < DBL_MAX) {
        a += 1, b -= 2 * v1, c += v1 * v1;
      }
    }
    c -= 1;
    const double det = b * b - 4 * a * c;
    if (det > 0) return (-b + sqrt(det)) / 2.0 / a;
    return 0.0;
  }
};

__global__ void k_fmm_seed(const uint8_t* __restrict__ free_in, const int* __restrict__ src, int W, int H, double* __restrict__ dd,
                           uint8_t* __restrict__ fixed, int* __restrict__ active, int tiles_x, int tiles_y) {
  const int e = blockIdx.x;
  if (threadIdx.x != 0) return;
  SeedMarcher m;
  m.n = 0;
  const int sr = src[e * 2], sc = src[e * 2 + 1];
  const uint8_t* fr = free_in + static_cast<size_t>(e) * W * H;
  for (int i = 0; i < SeedMarcher::N; ++i) {
    const int r = sr - kWinR + i / kWin, c = sc - kWinR + i % kWin;
    const bool inside = r >= 0 && r < W && c >= 0 && c < H;
    // cells outside the map do not exist for the marcher (its _getN returns -1): a masked cell behaves the same way
    m.flag[i] = (inside && fr[static_cast<size_t>(r) * H + c]) ? SeedMarcher::kFar : SeedMarcher::kMask;
    m.dist[i] = DBL_MAX, m.pos[i] = -1;
  }
  const int s = kWinR * kWin + kWinR;
  m.flag[s] = SeedMarcher::kFrozen, m.dist[s] = 0.0;   // initalizeFrozen: phi == 0 exactly (no sign changes elsewhere)
  for (int i = 0; i < SeedMarcher::N; ++i) {            // initalizeNarrow
    if (m.flag[i] != SeedMarcher::kFar) continue;
    for (int dim = 0; dim < 2; ++dim)
      for (int j = -1; j < 2; j += 2) {
        const int na = m.nbr(i, dim, j, SeedMarcher::kMask);
        if (na != -1 && m.flag[na] == SeedMarcher::kFrozen && m.flag[i] == SeedMarcher::kFar) {
          m.flag[i] = SeedMarcher::kNarrow;
          const double d = m.update(i);
          m.dist[i] = d;
          m.push(i, fabs(d));
        }
      }
  }
  while (m.n > 0) {                                     // solve
    double value;
    const int addr = m.pop(&value);
    if (value > kSeedRadius) break;
    m.flag[addr] = SeedMarcher::kFrozen;
    for (int dim = 0; dim < 2; ++dim)
      for (int j = -1; j < 2; j += 2) {
        const int na = m.nbr(addr, dim, j, SeedMarcher::kFrozen);
        if (na != -1 && m.flag[na] != SeedMarcher::kFrozen) {
          if (m.flag[na] == SeedMarcher::kNarrow) {
            const double d = m.update(na);
            if (d) m.set(na, fabs(d)), m.dist[na] = d;
          } else if (m.flag[na] == SeedMarcher::kFar) {
            const double d = m.update(na);
            if (d) m.dist[na] = d, m.flag[na] = SeedMarcher::kNarrow, m.push(na, fabs(d));
          }
        }
        const int local = m.nbr(addr, dim, j, SeedMarcher::kMask);
        if (local != -1 && m.flag[local] == SeedMarcher::kFrozen) {
          const int na2 = m.nbr(addr, dim, j * 2, SeedMarcher::kFrozen);
          if (na2 != -1 && m.flag[na2] == SeedMarcher::kNarrow) {
            const double d = m.update(na2);
            if (d) m.set(na2, fabs(d)), m.dist[na2] = d;
          }
        }
      }
  }
  double* T = dd + static_cast<size_t>(e) * W * H;
  uint8_t* fx = fixed + static_cast<size_t>(e) * W * H;
  for (int i = 0; i < SeedMarcher::N; ++i) {
    if (m.flag[i] != SeedMarcher::kFrozen) continue;
    const int r = sr - kWinR + i / kWin, c = sc - kWinR + i % kWin;
    T[static_cast<size_t>(r) * H + c] = m.dist[i];
    fx[static_cast<size_t>(r) * H + c] = 1;
    // the seed's own tile and its four axis neighbours (a seed next to a tile border feeds the neighbour's halo)
    const int tr = r / kTile, tc = c / kTile;
    int* act = active + e * tiles_y * tiles_x;
    act[tr * tiles_x + tc] = 1;
    if (tr > 0) act[(tr - 1) * tiles_x + tc] = 1;
    if (tr + 1 < tiles_y) act[(tr + 1) * tiles_x + tc] = 1;
    if (tc > 0) act[tr * tiles_x + tc - 1] = 1;
    if (tc + 1 < tiles_x) act[tr * tiles_x + tc + 1] = 1;
  }
}

// ------------------------------------------------------------------------------------------------ parallel eikonal
// One upwind dimension: (v1, v2) of the direction with the smaller known neighbour; v2 = inf when the second-order stencil
// does not apply (the marcher's test: second neighbour known and not larger).
__device__ __forceinline__ void upwind(double m1, double m2, double p1, double p2, double& v1, double& v2) {
  // the marcher scans j = -1 first and replaces only on a strictly smaller value
  if (p1 < m1) v1 = p1, v2 = p2;
  else v1 = m1, v2 = m2;
  if (!(v1 < dinf()) || !(v2 <= v1)) v2 = dinf();   // second order only behind a known first neighbour
}
__device__ __forceinline__ void add_terms(double v1, double v2, double& a, double& b, double& c) {
  if (v2 < dinf()) {
    const double tp = (1.0 / 3.0) * (4 * v1 - v2);
    a += 9.0 / 4.0, b -= 2 * (9.0 / 4.0) * tp, c += (9.0 / 4.0) * tp * tp;
  } else if (v1 < dinf()) {
    a += 1, b -= 2 * v1, c += v1 * v1;
  }
}
__device__ __forceinline__ double root(double a, double b, double c) {
  c -= 1;
  const double det = b * b - 4 * a * c;
  if (a > 0 && det > 0) return (-b + sqrt(det)) / 2.0 / a;
  return dinf();
}

struct EikonalArgs {
  const uint8_t* free_in;
  const uint8_t* fixed;
  double* dd;
  int* active;      // [2][E * tiles]: double-buffered activity flags
  int* counters;    // [0] = active tiles found in the current round
  int E, W, H, tiles_x, tiles_y;
};

__global__ void __launch_bounds__(kTile* kTile) k_eikonal(EikonalArgs g) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double T[2][kSpan][kSpan];
  __shared__ uint8_t ok[kSpan][kSpan];   // 1 = free cell that may be updated, 2 = fixed / read-only known cell, 0 = masked
  __shared__ int s_chg[2];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ntiles = g.E * g.tiles_x * g.tiles_y;
  int cur = 0;
  for (int round = 0; round < 100000; ++round) {
    int* act = g.active + cur * ntiles;
    int* nxt = g.active + (cur ^ 1) * ntiles;
    int found = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
      if (act[t] == 0) continue;   // uniform: only this CTA ever clears act[t], and only after the barrier below
      found = 1;
      __syncthreads();
      if (tx == 0 && ty == 0) act[t] = 0, s_chg[0] = 0, s_chg[1] = 0;
      const int e = t / (g.tiles_x * g.tiles_y), tr = (t / g.tiles_x) % g.tiles_y, tc = t % g.tiles_x;
      const int r0 = tr * kTile - kHalo, c0 = tc * kTile - kHalo;
      const size_t base = static_cast<size_t>(e) * g.W * g.H;
      for (int i = ty * kTile + tx; i < kSpan * kSpan; i += kTile * kTile) {
        const int rr = i / kSpan, cc = i % kSpan, r = r0 + rr, c = c0 + cc;
        double v = dinf();
        uint8_t k = 0;
        if (r >= 0 && r < g.W && c >= 0 && c < g.H) {
          const size_t idx = base + static_cast<size_t>(r) * g.H + c;
          if (g.free_in[idx]) {
            v = g.dd[idx];
            const bool interior = rr >= kHalo && rr < kHalo + kTile && cc >= kHalo && cc < kHalo + kTile;
            k = (interior && !g.fixed[idx]) ? 1 : 2;
          }
        }
        T[0][rr][cc] = v, T[1][rr][cc] = v, ok[rr][cc] = k;
      }
      __syncthreads();
      const int rr = ty + kHalo, cc = tx + kHalo;
      const bool mine = ok[rr][cc] == 1;
      int buf = 0, any = 0, last = 0;
      for (int it = 0; it < kInnerMax; ++it) {
        const int f = it & 1;
        if (mine) {
          const double(*S)[kSpan] = T[buf];
          double v1y, v2y, v1x, v2x;
          upwind(S[rr - 1][cc], S[rr - 2][cc], S[rr + 1][cc], S[rr + 2][cc], v1y, v2y);
          upwind(S[rr][cc - 1], S[rr][cc - 2], S[rr][cc + 1], S[rr][cc + 2], v1x, v2x);
          double a = 0, b = 0, c = 0;
          add_terms(v1y, v2y, a, b, c);
          add_terms(v1x, v2x, a, b, c);
          double r = root(a, b, c);
          // causality: the value must lie above every neighbour it was computed from; otherwise only the smaller
          // dimension is upwind of this cell (the marcher would not have had the larger one frozen yet)
          const double big = fmax(v1y < dinf() ? v1y : -dinf(), v1x < dinf() ? v1x : -dinf());
          if (!(r < dinf()) || r < big) {
            double a1 = 0, b1 = 0, c1 = 0;
            if (v1y <= v1x) add_terms(v1y, v2y, a1, b1, c1);
            else add_terms(v1x, v2x, a1, b1, c1);
            r = root(a1, b1, c1);
          }
          T[buf ^ 1][rr][cc] = r;
          if (r != S[rr][cc]) s_chg[f] = 1;
        }
        __syncthreads();
        last = s_chg[f];
        buf ^= 1;
        if (tx == 0 && ty == 0) s_chg[f ^ 1] = 0;   // the next iteration's flag; everybody reads flag f right now
        if (!last) break;
        any = 1;
        __syncthreads();
      }
      if (any) {
        if (mine) g.dd[base + static_cast<size_t>(r0 + rr) * g.H + c0 + cc] = T[buf][rr][cc];
        if (tx == 0 && ty == 0) {  // the 2-cell halo of the four axis neighbours reads this tile
          if (tr > 0) nxt[t - g.tiles_x] = 1;
          if (tr + 1 < g.tiles_y) nxt[t + g.tiles_x] = 1;
          if (tc > 0) nxt[t - 1] = 1;
          if (tc + 1 < g.tiles_x) nxt[t + 1] = 1;
          if (last) nxt[t] = 1;   // ran out of inner iterations: not converged inside the tile yet
        }
      }
    }
    if (found && tx == 0 && ty == 0) atomicAdd(&g.counters[round & 1], 1);
    __threadfence();
    grid.sync();
    const int total = *reinterpret_cast<volatile int*>(&g.counters[round & 1]);
    if (blockIdx.x == 0 && tx == 0 && ty == 0) g.counters[(round + 1) & 1] = 0;  // nobody touches it before the next barrier
    if (total == 0) break;
    cur ^= 1;
    grid.sync();
  }
}

// ------------------------------------------------------------------------------------------------ weighting + goal
// Pass 1 (per environment): number of cells the marcher left masked / unreached, maximum of the finite distances, and
// sum(exp(-dd / temperature)) over the local window, with dd post-processed like :392-393.
__device__ __forceinline__ double warp_sum(double v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct GoalStats {
  double max_finite;
  unsigned long long n_inf;
};

__global__ void __launch_bounds__(1024) k_goal_stats(const double* __restrict__ dd, int W, int H, GoalStats* __restrict__ stats) {
  const int e = blockIdx.y;
  const double* T = dd + static_cast<size_t>(e) * W * H;
  double mx = -dinf();
  unsigned long long ninf = 0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < static_cast<size_t>(W) * H;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double v = T[i];
    if (v < dinf()) mx = fmax(mx, v);
    else ++ninf;
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    ninf += __shfl_xor_sync(0xffffffffu, ninf, o);
  }
  if ((threadIdx.x & 31) == 0) {
    // max over non-negative doubles == max over their bit patterns
    if (mx >= 0) atomicMax(reinterpret_cast<unsigned long long*>(&stats[e].max_finite), static_cast<unsigned long long>(__double_as_longlong(mx)));
    if (ninf) atomicAdd(&stats[e].n_inf, ninf);
  }
}

// dd as the reference leaves it after :392-393: masked / unreached cells are inf; when NOTHING is masked the fill value is
// unused and the farthest reached cells themselves (dd == max) become inf.
__device__ __forceinline__ double post_dd(double v, const GoalStats& s) {
  if (s.n_inf == 0 && v == s.max_finite) return dinf();
  return v;
}

struct GoalArgs {
  const double* dd;
  const GoalStats* stats;
  const float* target_pred;
  const int* lmb;
  double* dd_wt;
  int* dd_wt_valid;
  double* value;
  double* sums;              // [E] scratch
  unsigned long long* best;  // [E] scratch: packed (ordered value bits, ~index)
  int* global_goal;
  int* goal_kind;
  int* last_goal;
  int* last_kind;
  int W, H, lw, lh;
  double temperature, dwt;   // dist_weight_temperature / map_resolution, dist_weight_temperature
};

__global__ void __launch_bounds__(256) k_goal_sum(GoalArgs g) {
  const int e = blockIdx.y;
  const GoalStats st = g.stats[e];
  const int r0 = g.lmb[e * 4], c0 = g.lmb[e * 4 + 2];
  const double* T = g.dd + static_cast<size_t>(e) * g.W * g.H;
  double s = 0;
  const int n = g.lw * g.lh;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / g.lh, c = i % g.lh;
    s += exp(-post_dd(T[static_cast<size_t>(r0 + r) * g.H + c0 + c], st) / g.temperature);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(&g.sums[e], s);
}

// monotone map double -> uint64 (total order of finite / infinite values; NaN never occurs: both factors are >= 0)
__device__ __forceinline__ unsigned long long order_bits(double v) {
  const unsigned long long u = static_cast<unsigned long long>(__double_as_longlong(v));
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}

__global__ void __launch_bounds__(256) k_goal_value(GoalArgs g) {
  const int e = blockIdx.y;
  const GoalStats st = g.stats[e];
  const int r0 = g.lmb[e * 4], c0 = g.lmb[e * 4 + 2];
  const double* T = g.dd + static_cast<size_t>(e) * g.W * g.H;
  const int n = g.lw * g.lh;
  // "stuck inside obstacle, use last dd_wt" (:398-399)
  const bool keep_prev = g.sums[e] < 10.0 && g.dd_wt_valid[e] != 0;
  double* wt = g.dd_wt + static_cast<size_t>(e) * n;
  const float* tp = g.target_pred + static_cast<size_t>(e) * n;
  unsigned long long best_v = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = i / g.lh, c = i % g.lh;
    const double d = post_dd(T[static_cast<size_t>(r0 + r) * g.H + c0 + c], st);
    const double w = keep_prev ? wt[i] : exp(-d / g.temperature);
    if (!keep_prev) wt[i] = w;
    double v;
    if (g.dwt == -1.0) v = static_cast<double>(tp[i]);                       // no weighting (:401-402)
    else if (g.dwt == 0.0) v = exp(-(d < 60.0 ? dinf() : d) / 100.0);         // frontier-based exploration (:403-405)
    else v = static_cast<double>(tp[i]) * w;
    if (g.value != nullptr) g.value[static_cast<size_t>(e) * n + i] = v;
    best_v = max(best_v, order_bits(v));
  }
  // numpy argmax = the first maximum in row-major order: the maximum value here, its smallest index in k_goal_index
  for (int o = 16; o > 0; o >>= 1) best_v = max(best_v, __shfl_xor_sync(0xffffffffu, best_v, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&g.best[e * 2], best_v);
}

// Second round of the argmax: the smallest index among the cells holding the maximum value.
__global__ void __launch_bounds__(256) k_goal_index(GoalArgs g) {
  const int e = blockIdx.y;
  const GoalStats st = g.stats[e];
  const int r0 = g.lmb[e * 4], c0 = g.lmb[e * 4 + 2];
  const double* T = g.dd + static_cast<size_t>(e) * g.W * g.H;
  const int n = g.lw * g.lh;
  const double* wt = g.dd_wt + static_cast<size_t>(e) * n;
  const float* tp = g.target_pred + static_cast<size_t>(e) * n;
  const unsigned long long want = g.best[e * 2];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double v;
    if (g.dwt == -1.0) v = static_cast<double>(tp[i]);
    else if (g.dwt == 0.0) {
      const int r = i / g.lh, c = i % g.lh;
      const double d = post_dd(T[static_cast<size_t>(r0 + r) * g.H + c0 + c], st);
      v = exp(-(d < 60.0 ? dinf() : d) / 100.0);
    } else v = static_cast<double>(tp[i]) * wt[i];
    if (order_bits(v) == want) atomicMin(&g.best[e * 2 + 1], static_cast<unsigned long long>(i));
  }
}

// :412-415 for every environment: the new goal replaces the current one unless it equals the LAST one.  kinds: 0 = None,
// 1 = list of lists (set by init / presets: never equal to the tuple the argmax yields), 2 = tuple from this method.
__global__ void k_goal_commit(GoalArgs g, int E) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  g.dd_wt_valid[e] = 1;
  const int i = static_cast<int>(g.best[e * 2 + 1]);
  const int nr = i / g.lh, ncol = i % g.lh;
  const bool same = g.last_kind[e] == 2 && g.last_goal[e * 2] == nr && g.last_goal[e * 2 + 1] == ncol;
  if (!same) {
    g.last_goal[e * 2] = g.global_goal[e * 2], g.last_goal[e * 2 + 1] = g.global_goal[e * 2 + 1];
    g.last_kind[e] = g.goal_kind[e];
    g.global_goal[e * 2] = nr, g.global_goal[e * 2 + 1] = ncol;
    g.goal_kind[e] = 2;
  }
}

struct Workspace {
  int device = -1;
  size_t cells = 0;
  int E = 0, tiles = 0;
  uint8_t* free_mask = nullptr;
  uint8_t* fixed = nullptr;
  int* src = nullptr;
  int* active = nullptr;
  int* counters = nullptr;
  GoalStats* stats = nullptr;
  double* sums = nullptr;
  unsigned long long* best = nullptr;
  void release() {
    for (void* p : {static_cast<void*>(free_mask), static_cast<void*>(fixed), static_cast<void*>(src), static_cast<void*>(active),
                    static_cast<void*>(counters), static_cast<void*>(stats), static_cast<void*>(sums), static_cast<void*>(best)})
      if (p) cudaFree(p);
    *this = Workspace();
  }
};

Workspace& workspace(int device, int E, int W, int H, int tiles) {
  static thread_local Workspace ws;
  const size_t cells = static_cast<size_t>(E) * W * H;
  if (ws.device != device || ws.cells != cells || ws.E != E || ws.tiles != tiles) {
    ws.release();
    ws.device = device, ws.cells = cells, ws.E = E, ws.tiles = tiles;
    PN_CUDA_CHECK(cudaMalloc(&ws.free_mask, cells));
    PN_CUDA_CHECK(cudaMalloc(&ws.fixed, cells));
    PN_CUDA_CHECK(cudaMalloc(&ws.src, sizeof(int) * 2 * E));
    PN_CUDA_CHECK(cudaMalloc(&ws.active, sizeof(int) * 2 * E * tiles));
    PN_CUDA_CHECK(cudaMalloc(&ws.counters, sizeof(int) * 2));
    PN_CUDA_CHECK(cudaMalloc(&ws.stats, sizeof(GoalStats) * E));
    PN_CUDA_CHECK(cudaMalloc(&ws.sums, sizeof(double) * E));
    PN_CUDA_CHECK(cudaMalloc(&ws.best, sizeof(unsigned long long) * 2 * E));
  }
  return ws;
}

__global__ void k_fill_inf(double* p, size_t n) {
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) p[i] = dinf();
}
__global__ void k_init_best(unsigned long long* best, int E) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) best[e * 2] = 0ull, best[e * 2 + 1] = ~0ull;
}

}  // namespace

void launch_global_goal(int device, int num_sms, const pn_goal_cfg& c, const pn_goal_arrays& a, int E, int only_distance, cudaStream_t s) {
  const int W = c.full_w, H = c.full_h;
  const int tiles_x = (H + kTile - 1) / kTile, tiles_y = (W + kTile - 1) / kTile, tiles = tiles_x * tiles_y;
  Workspace& ws = workspace(device, E, W, H, tiles);
  const size_t cells = static_cast<size_t>(E) * W * H;
  PN_REQUIRE(c.col_rad >= 0 && c.col_rad <= 12, "pn_global_goal: col_rad out of range (0..12)");
  // ---- traversible mask + source cell
  const int span = kTile + 2 * c.col_rad;
  k_traversible<<<dim3(tiles_x, tiles_y, E), dim3(kTile, kTile), static_cast<size_t>(span) * span, s>>>(
      a.full_map, c.num_channels, W, H, c.col_rad, a.collision_map, a.visited_vis, a.lmb, a.loc, ws.free_mask, ws.src);
  // ---- geodesic distance
  k_fill_inf<<<num_sms * 4, 256, 0, s>>>(a.dd, cells);
  PN_CUDA_CHECK(cudaMemsetAsync(ws.fixed, 0, cells, s));
  PN_CUDA_CHECK(cudaMemsetAsync(ws.active, 0, sizeof(int) * 2 * E * tiles, s));
  PN_CUDA_CHECK(cudaMemsetAsync(ws.counters, 0, sizeof(int) * 2, s));
  k_fmm_seed<<<E, 32, 0, s>>>(ws.free_mask, ws.src, W, H, a.dd, ws.fixed, ws.active, tiles_x, tiles_y);
  EikonalArgs ek{ws.free_mask, ws.fixed, a.dd, ws.active, ws.counters, E, W, H, tiles_x, tiles_y};
  int per_sm = 0;
  PN_CUDA_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_eikonal, kTile * kTile, 0));
  PN_REQUIRE(per_sm >= 1, "pn_global_goal: the eikonal kernel does not fit an SM");
  const int grid = std::min(num_sms * per_sm, E * tiles);
  void* kargs[] = {&ek};
  PN_CUDA_CHECK(cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_eikonal), dim3(grid), dim3(kTile, kTile), kargs, 0, s));
  if (only_distance) {
    PN_CUDA_CHECK(cudaGetLastError());
    return;
  }
  // ---- weighting, value, argmax, goal bookkeeping
  PN_CUDA_CHECK(cudaMemsetAsync(ws.stats, 0, sizeof(GoalStats) * E, s));
  PN_CUDA_CHECK(cudaMemsetAsync(ws.sums, 0, sizeof(double) * E, s));
  k_init_best<<<(E + 63) / 64, 64, 0, s>>>(ws.best, E);
  k_goal_stats<<<dim3(32, E), 1024, 0, s>>>(a.dd, W, H, ws.stats);
  GoalArgs g{a.dd, ws.stats, a.target_pred, a.lmb, a.dd_wt, a.dd_wt_valid, a.value, ws.sums, ws.best, a.global_goal, a.goal_kind,
             a.last_global_goal, a.last_kind, W, H, c.local_w, c.local_h, c.dist_weight_temperature / c.map_resolution,
             c.dist_weight_temperature};
  const int gb = std::min(128, (c.local_w * c.local_h + 255) / 256);
  k_goal_sum<<<dim3(gb, E), 256, 0, s>>>(g);
  k_goal_value<<<dim3(gb, E), 256, 0, s>>>(g);
  k_goal_index<<<dim3(gb, E), 256, 0, s>>>(g);
  k_goal_commit<<<(E + 63) / 64, 64, 0, s>>>(g, E);
  PN_CUDA_CHECK(cudaGetLastError());
}

}  // namespace pn
