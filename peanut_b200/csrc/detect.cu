// Mask-RCNN control flow on the device (no host synchronisation anywhere): FPN top-down add, RPN top-k +
// box decoding + per-level NMS + cross-level merge, ROIAlign, box-head post-processing (softmax, score
// filter, per-class NMS, top-100), and the mask paste + per-category accumulation of
// SemanticPredMaskRCNN.get_prediction (nav/agent/utils/segmentation.py:47-62).
//
// Semantics follow detectron2 0.6 as configured by nav/agent/utils/COCO-InstSeg/mask_rcnn_R_101_cat9.yaml
// (restated in oracle/maskrcnn.py, which cites the yaml lines).  This translation unit is compiled with
// -fmad=false so that box decoding, IoU, ROIAlign and paste arithmetic round after every multiply and add
// like the reference's separate torch ops.
#include <cfloat>
#include <cmath>

#include "maskrcnn.h"
#include "vec.cuh"

namespace pn {

namespace {

constexpr float kScaleClamp = 4.135166556742356f;  // log(1000 / 16)

__device__ __forceinline__ uint32_t fkey(float f) {  // order-preserving float -> uint32
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Box2BoxTransform.apply_deltas for one box.
__device__ __forceinline__ void decode_box(const float a[4], float dx, float dy, float dw, float dh, float wx, float wy,
                                           float ww, float wh, float out[4]) {
  const float width = a[2] - a[0], height = a[3] - a[1];
  const float cx = a[0] + 0.5f * width, cy = a[1] + 0.5f * height;
  dx = dx / wx, dy = dy / wy, dw = dw / ww, dh = dh / wh;
  dw = fminf(dw, kScaleClamp), dh = fminf(dh, kScaleClamp);
  const float pcx = dx * width + cx, pcy = dy * height + cy;
  const float pw = expf(dw) * width, ph = expf(dh) * height;
  out[0] = pcx - 0.5f * pw, out[1] = pcy - 0.5f * ph, out[2] = pcx + 0.5f * pw, out[3] = pcy + 0.5f * ph;
}
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

__device__ __forceinline__ bool iou_gt(const float4 a, const float4 b, float thr) {
  const float area_a = (a.z - a.x) * (a.w - a.y), area_b = (b.z - b.x) * (b.w - b.y);
  const float w = fmaxf(fminf(a.z, b.z) - fmaxf(a.x, b.x), 0.f), h = fmaxf(fminf(a.w, b.w) - fmaxf(a.y, b.y), 0.f);
  const float inter = w * h;
  return inter / (area_a + area_b - inter) > thr;
}

// Block-wide exclusive scan of one flag per thread (blockDim.x == 1024); returns rank, writes total.
__device__ __forceinline__ int block_scan_flag(bool flag, int* warp_sums, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) warp_sums[wid] = __popc(bal);
  __syncthreads();
  if (wid == 0) {
    int v = warp_sums[lane], x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    warp_sums[lane] = x - v;  // exclusive
    if (lane == 31) warp_sums[32] = x;
  }
  __syncthreads();
  const int r = warp_sums[wid] + in_warp;
  total = warp_sums[32];
  __syncthreads();
  return r;
}

// In-place bitonic sort, descending, of n (power of two) 64-bit keys in shared memory.
__device__ __forceinline__ void bitonic_desc(unsigned long long* keys, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int l = i ^ j;
        if (l > i) {
          const unsigned long long a = keys[i], b = keys[l];
          const bool desc = (i & k) == 0;
          if ((a < b) == desc) keys[i] = b, keys[l] = a;
        }
      }
      __syncthreads();
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// FPN: lat += nearest-neighbour x2 upsampling of prev (F.interpolate(scale_factor=2, mode="nearest")).
template <typename T>
__global__ void k_upsample2x_add(const T* __restrict__ prev, long long ldp, int h, int w, T* __restrict__ lat, long long ldl,
                                 int B, int H, int W, int C8, int round_tf32) {
  pdl_grid_sync();
  Idx4 ix;
  if (!split_index(H, W, C8, static_cast<uint32_t>(B) * H * W * C8, ix)) return;
  const int b = static_cast<int>(ix.a), y = static_cast<int>(ix.b), x = static_cast<int>(ix.c), cg = static_cast<int>(ix.d);
  const int ys = min(y >> 1, h - 1), xs = min(x >> 1, w - 1);
  float a[8], p[8];
  T* lp = lat + ((static_cast<long long>(b) * H + y) * W + x) * ldl + cg * 8;
  load8(lp, a);
  load8(prev + ((static_cast<long long>(b) * h + ys) * w + xs) * ldp + cg * 8, p);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = a[j] + p[j];
    if (round_tf32) {
      uint32_t q;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q) : "f"(a[j]));
      a[j] = __uint_as_float(q);
    }
  }
  store8(lp, a);
}

template <typename T>
__global__ void k_subsample2(const T* __restrict__ in, long long ldi, int H, int W, T* __restrict__ out, long long ldo, int B,
                             int Ho, int Wo, int C8) {
  pdl_grid_sync();
  Idx4 ix;
  if (!split_index(Ho, Wo, C8, static_cast<uint32_t>(B) * Ho * Wo * C8, ix)) return;
  const int b = static_cast<int>(ix.a), y = static_cast<int>(ix.b), x = static_cast<int>(ix.c), cg = static_cast<int>(ix.d);
  float v[8];
  load8(in + ((static_cast<long long>(b) * H + 2 * y) * W + 2 * x) * ldi + cg * 8, v);
  store8(out + ((static_cast<long long>(b) * Ho + y) * Wo + x) * ldo + cg * 8, v);
}

// ---------------------------------------------------------------------------------------------------
// RPN per (level, image): exact top-k by 4-pass radix select (ties -> lower index), sort, decode, clip,
// drop empty / non-finite, greedy NMS via a shared-memory suppression bit matrix.  block = 1024 threads.
constexpr int kCandCap = 4096;     // candidate capacity of the multi-CTA pre-selection (fallback beyond it)
constexpr int kHistBins = 65536;   // histogram over the top 16 bits of the order-preserving key

struct RpnSmem {
  unsigned long long keys[kCandCap];
  uint32_t hist[256];
  int warp_sums[33];
  uint32_t prefix, rank, cnt_eq;
  int ncand;
};

// ---- multi-CTA pre-selection: (1) 16-bit key histogram per (image, level), (2) threshold bin holding the k-th
// largest key, (3) every logit at or above that bin is appended to a candidate list.  The list is a superset of
// the exact top-k (typically k + a few dozen); k_rpn_select_nms sorts it.  Scanning the 163 200 logits of P2 is
// spread over ~50 CTAs instead of one.
constexpr int kPixPerBlock = 1024;

__global__ void __launch_bounds__(256) k_rpn_hist(RpnMeta meta, uint32_t* __restrict__ hist) {
  pdl_grid_sync();
  const int L = blockIdx.y % kRpnLevels, b = blockIdx.y / kRpnLevels;
  const RpnLevel& lv = meta.lv[L];
  const int npix = lv.H * lv.W;
  const int p0 = blockIdx.x * kPixPerBlock;
  if (p0 >= npix) return;
  const float* head = lv.head + static_cast<size_t>(b) * npix * kRpnHeadC;
  uint32_t* h = hist + static_cast<size_t>(b * kRpnLevels + L) * kHistBins;
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < kPixPerBlock / 256; ++j) {
    const int pix = p0 + j * 256 + threadIdx.x;
    const bool in = pix < npix;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (in) v = *reinterpret_cast<const float4*>(head + static_cast<size_t>(pix) * kRpnHeadC);
    const float lg[3] = {v.x, v.y, v.z};
    const unsigned act = __ballot_sync(0xffffffffu, in);
    if (in) {
#pragma unroll
      for (int a = 0; a < kAnchors; ++a) {
        const uint32_t bin = fkey(lg[a]) >> 16;
        const unsigned peers = __match_any_sync(act, bin);
        if (lane == __ffs(peers) - 1) atomicAdd(&h[bin], static_cast<uint32_t>(__popc(peers)));
      }
    }
  }
}

__global__ void __launch_bounds__(1024) k_rpn_threshold(RpnMeta meta, const uint32_t* __restrict__ hist,
                                                        uint32_t* __restrict__ thr_bin, int* __restrict__ ncand) {
  pdl_grid_sync();
  __shared__ int warp_sums[33];
  __shared__ uint32_t s_bin;
  const int L = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
  const RpnLevel& lv = meta.lv[L];
  const int n = lv.H * lv.W * kAnchors;
  const uint32_t k = static_cast<uint32_t>(min(n, meta.pre_topk));
  const uint32_t* h = hist + static_cast<size_t>(b * kRpnLevels + L) * kHistBins;
  // thread t owns the 64 bins [65535 - 64t - 63, 65535 - 64t], i.e. descending key order across threads
  constexpr int kPer = kHistBins / 1024;
  const int hi = kHistBins - 1 - tid * kPer;
  uint32_t c[kPer], sum = 0;
#pragma unroll
  for (int j = 0; j < kPer / 4; ++j) {
    const uint4 v = *reinterpret_cast<const uint4*>(h + hi - 4 * j - 3);
    c[4 * j] = v.w, c[4 * j + 1] = v.z, c[4 * j + 2] = v.y, c[4 * j + 3] = v.x;
    sum += v.x + v.y + v.z + v.w;
  }
  // inclusive scan of `sum` over threads
  const int lane = tid & 31, wid = tid >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += y;
  }
  if (lane == 31) warp_sums[wid] = static_cast<int>(incl);
  __syncthreads();
  if (wid == 0) {
    int v = warp_sums[lane], x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, x, o);
      if (lane >= o) x += y;
    }
    warp_sums[lane] = x - v;
  }
  __syncthreads();
  incl += static_cast<uint32_t>(warp_sums[wid]);
  const uint32_t before = incl - sum;
  if (k - 1 >= before && k - 1 < incl) {  // the k-th largest key falls into one of this thread's bins
    uint32_t rr = k - 1 - before;
    int j = 0;
    while (rr >= c[j]) rr -= c[j], ++j;
    s_bin = static_cast<uint32_t>(hi - j);
  }
  __syncthreads();
  if (tid == 0) {
    thr_bin[b * kRpnLevels + L] = s_bin;
    ncand[b * kRpnLevels + L] = 0;
  }
}

__global__ void __launch_bounds__(256) k_rpn_collect(RpnMeta meta, const uint32_t* __restrict__ thr_bin, int* __restrict__ ncand,
                                                     unsigned long long* __restrict__ cand) {
  pdl_grid_sync();
  const int L = blockIdx.y % kRpnLevels, b = blockIdx.y / kRpnLevels;
  const RpnLevel& lv = meta.lv[L];
  const int npix = lv.H * lv.W;
  const int p0 = blockIdx.x * kPixPerBlock;
  if (p0 >= npix) return;
  const float* head = lv.head + static_cast<size_t>(b) * npix * kRpnHeadC;
  const uint32_t T16 = thr_bin[b * kRpnLevels + L];
  int* cnt = ncand + b * kRpnLevels + L;
  unsigned long long* out = cand + static_cast<size_t>(b * kRpnLevels + L) * kCandCap;
#pragma unroll
  for (int j = 0; j < kPixPerBlock / 256; ++j) {
    const int pix = p0 + j * 256 + threadIdx.x;
    if (pix >= npix) continue;
    const float4 v = *reinterpret_cast<const float4*>(head + static_cast<size_t>(pix) * kRpnHeadC);
    const float lg[3] = {v.x, v.y, v.z};
#pragma unroll
    for (int a = 0; a < kAnchors; ++a) {
      const uint32_t key = fkey(lg[a]);
      if ((key >> 16) >= T16) {
        const int slot = atomicAdd(cnt, 1);
        if (slot < kCandCap)
          out[slot] = (static_cast<unsigned long long>(key) << 32) | (0xffffffffu - static_cast<uint32_t>(pix * kAnchors + a));
      }
    }
  }
}

// (a) k_rpn_sort_decode: per (level, image) - exact top-k (sorted pre-selection, or in-block radix select as the
//     fallback), box decoding, clipping, removal of empty / non-finite boxes -> score-ordered list in global memory;
// (b) k_rpn_mask: the m x m suppression bit matrix, 32 rows per CTA (the ~500 k IoU tests of one level would
//     otherwise serialise on a single SM);
// (c) k_rpn_sweep: greedy sweep over the bit matrix + compaction of the survivors.
__global__ void __launch_bounds__(1024, 1) k_rpn_sort_decode(RpnMeta meta, const int* __restrict__ ncand,
                                                             const unsigned long long* __restrict__ cand,
                                                             float4* __restrict__ sbox, float* __restrict__ sscore,
                                                             int* __restrict__ scount) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t smem_raw[];
  RpnSmem& S = *reinterpret_cast<RpnSmem*>(smem_raw);
  const int L = blockIdx.x, b = blockIdx.y;
  const RpnLevel& lv = meta.lv[L];
  const int tid = threadIdx.x, lane = tid & 31;
  const int npix = lv.H * lv.W;
  const int n = npix * kAnchors;
  const int k = min(n, meta.pre_topk);
  const float* head = lv.head + static_cast<size_t>(b) * npix * kRpnHeadC;
  auto logit_at = [&](int i) { return head[static_cast<size_t>(i / kAnchors) * kRpnHeadC + (i % kAnchors)]; };

  const int nc = ncand[b * kRpnLevels + L];
  if (nc <= kCandCap) {
    // ---- fast path: sort the pre-selected superset; its first k entries are the exact top-k
    int npow = 1;
    while (npow < nc) npow <<= 1;
    const unsigned long long* src = cand + static_cast<size_t>(b * kRpnLevels + L) * kCandCap;
    for (int i = tid; i < npow; i += 1024) S.keys[i] = i < nc ? src[i] : 0ull;
    __syncthreads();
    if (npow > 1) bitonic_desc(S.keys, npow);
  } else {
  // ---- fallback (more than kCandCap logits share the threshold bin): in-block radix select of the k-th largest key
  if (tid == 0) S.prefix = 0, S.rank = static_cast<uint32_t>(k - 1), S.ncand = 0;
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) S.hist[tid] = 0;
    __syncthreads();
    const uint32_t prefix = S.prefix;
    const uint32_t pmask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int base = 0; base < n; base += 1024) {
      const int i = base + tid;
      uint32_t key = 0;
      bool part = false;
      if (i < n) {
        key = fkey(logit_at(i));
        part = (key & pmask) == prefix;
      }
      const unsigned act = __ballot_sync(0xffffffffu, part);
      if (part) {
        const uint32_t bin = (key >> shift) & 255u;
        const unsigned peers = __match_any_sync(act, bin);
        if (lane == __ffs(peers) - 1) atomicAdd(&S.hist[bin], static_cast<uint32_t>(__popc(peers)));
      }
    }
    __syncthreads();
    if (tid < 32) {
      // bins in descending order: lane owns bins [255 - 8*lane - 7, 255 - 8*lane]
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) c[j] = S.hist[255 - (lane * 8 + j)], sum += c[j];
      uint32_t incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      uint32_t before = incl - sum;  // elements in bins above this lane's
      const uint32_t r = S.rank;
      if (r >= before && r < incl) {
        uint32_t rr = r - before;
        int j = 0;
        while (rr >= c[j]) rr -= c[j], ++j;
        S.rank = rr;
        S.prefix = prefix | (static_cast<uint32_t>(255 - (lane * 8 + j)) << shift);
        S.cnt_eq = c[j];
      }
    }
    __syncthreads();
  }
  const uint32_t T = S.prefix;
  const int need_eq = static_cast<int>(S.rank) + 1, cnt_eq = static_cast<int>(S.cnt_eq);
  __syncthreads();
  // ---- collect the k winners as (key, ~index) composites
  int eq_seen = 0;
  for (int base = 0; base < n; base += 1024) {
    const int i = base + tid;
    const uint32_t key = i < n ? fkey(logit_at(i)) : 0u;
    const bool gt = i < n && key > T;
    bool eq = i < n && key == T;
    if (need_eq < cnt_eq) {  // block-uniform: ties must be taken in index order
      int tot;
      const int r = block_scan_flag(eq, S.warp_sums, tot);
      eq = eq && (eq_seen + r < need_eq);
      eq_seen += tot;
    }
    if (gt || eq) {
      const int slot = atomicAdd(&S.ncand, 1);
      S.keys[slot] = (static_cast<unsigned long long>(key) << 32) | (0xffffffffu - static_cast<uint32_t>(i));
    }
  }
  __syncthreads();
  for (int i = S.ncand + tid; i < kRpnCap; i += 1024) S.keys[i] = 0ull;
  __syncthreads();
  bitonic_desc(S.keys, kRpnCap);
  }

  // ---- decode, clip, validity
  bool valid = false;
  float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
  float sc = 0.f;
  if (tid < k) {
    const unsigned long long kk = S.keys[tid];
    const int i = static_cast<int>(0xffffffffu - static_cast<uint32_t>(kk & 0xffffffffull));
    const int pix = i / kAnchors, a = i - pix * kAnchors;
    const int y = pix / lv.W, x = pix - y * lv.W;
    const float sx = static_cast<float>(x * lv.stride), sy = static_cast<float>(y * lv.stride);
    const float anchor[4] = {sx + lv.base[a][0], sy + lv.base[a][1], sx + lv.base[a][2], sy + lv.base[a][3]};
    const float* d = head + static_cast<size_t>(pix) * kRpnHeadC + kAnchors + a * 4;
    float o[4];
    decode_box(anchor, d[0], d[1], d[2], d[3], 1.f, 1.f, 1.f, 1.f, o);
    sc = fkey_inv(static_cast<uint32_t>(kk >> 32));
    valid = isfinite(o[0]) && isfinite(o[1]) && isfinite(o[2]) && isfinite(o[3]) && isfinite(sc);
    bx = make_float4(clampf(o[0], 0.f, meta.img_w), clampf(o[1], 0.f, meta.img_h), clampf(o[2], 0.f, meta.img_w),
                     clampf(o[3], 0.f, meta.img_h));
    valid = valid && (bx.z - bx.x > 0.f) && (bx.w - bx.y > 0.f);
  }
  int m;
  const int pos = block_scan_flag(valid, S.warp_sums, m);
  const size_t base = (static_cast<size_t>(b) * kRpnLevels + L) * kRpnCap;
  if (valid) sbox[base + pos] = bx, sscore[base + pos] = sc;
  if (tid == 0) scount[b * kRpnLevels + L] = m;
}

__global__ void __launch_bounds__(256) k_rpn_mask(float nms_thr, const float4* __restrict__ sbox, const int* __restrict__ scount,
                                                  uint32_t* __restrict__ mask_g) {
  pdl_grid_sync();
  __shared__ float4 box[kRpnCap];
  const int L = blockIdx.y, b = blockIdx.z;
  const int m = scount[b * kRpnLevels + L];
  const int r0 = blockIdx.x * 32;
  if (r0 >= m) return;
  const size_t base = (static_cast<size_t>(b) * kRpnLevels + L) * kRpnCap;
  for (int i = threadIdx.x; i < m; i += blockDim.x) box[i] = sbox[base + i];
  __syncthreads();
  const int nw = (m + 31) >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* mk = mask_g + base * 32;
  // bit jj of mask[i][j] <=> box 32j+jj (later in score order) overlaps box i; lane jj tests one box, the word is
  // the ballot.  Only the upper triangle (j >= i / 32) is ever consulted by the sweep.
  for (int i = r0 + warp; i < min(r0 + 32, m); i += 8) {
    const float4 a = box[i];
    for (int j = i >> 5; j < nw; ++j) {
      const int c = (j << 5) + lane;
      const bool hit = c < m && c > i && iou_gt(a, box[c], nms_thr);
      const unsigned bits = __ballot_sync(0xffffffffu, hit);
      if (lane == 0) mk[i * 32 + j] = bits;
    }
  }
}

__global__ void __launch_bounds__(1024, 1) k_rpn_sweep(const float4* __restrict__ sbox, const float* __restrict__ sscore,
                                                       const int* __restrict__ scount, const uint32_t* __restrict__ mask_g,
                                                       float* __restrict__ lvl_boxes, float* __restrict__ lvl_scores,
                                                       int* __restrict__ lvl_count) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint32_t mask[];  // [kRpnCap][32]
  __shared__ int kept[kRpnCap];
  __shared__ int s_nkept;
  const int L = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31;
  const int m = scount[b * kRpnLevels + L];
  const int nw = (m + 31) >> 5;
  const size_t base = (static_cast<size_t>(b) * kRpnLevels + L) * kRpnCap;
  {
    // rows i need words j >= i / 32 only; copy whole rows (coalesced 128-byte lines)
    const uint4* src = reinterpret_cast<const uint4*>(mask_g + base * 32);
    uint4* dst = reinterpret_cast<uint4*>(mask);
    for (int i = tid; i < m * 8; i += 1024) dst[i] = src[i];
  }
  __syncthreads();
  // ---- greedy sweep by one warp: lane l owns word l of the `removed` bit vector
  if (tid < 32) {
    uint32_t removed = 0;
    int nkept = 0;
    for (int c = 0; c < nw; ++c) {
      uint32_t rem = __shfl_sync(0xffffffffu, removed, c);
      const int i_lane = (c << 5) + lane;
      const uint32_t intra = i_lane < m ? mask[i_lane * 32 + c] : 0u;
      uint32_t kept_bits = 0;
      const int cend = min(32, m - (c << 5));
      for (int j = 0; j < cend; ++j) {
        const uint32_t row = __shfl_sync(0xffffffffu, intra, j);
        if (!((rem >> j) & 1u)) {
          kept_bits |= 1u << j;
          rem |= row;
        }
      }
      // fold the rows of the kept boxes into every word (words below the diagonal were never written: skip them)
      uint32_t kb = kept_bits;
      while (kb) {
        const int j = __ffs(kb) - 1;
        kb &= kb - 1;
        if (lane < nw && lane >= c) removed |= mask[((c << 5) + j) * 32 + lane];
      }
      if ((kept_bits >> lane) & 1u) kept[nkept + __popc(kept_bits & ((1u << lane) - 1u))] = i_lane;
      nkept += __popc(kept_bits);
    }
    if (lane == 0) s_nkept = nkept;
  }
  __syncthreads();
  const int nk = s_nkept;
  float* ob = lvl_boxes + (static_cast<size_t>(b) * kRpnLevels + L) * kRpnCap * 4;
  float* os = lvl_scores + (static_cast<size_t>(b) * kRpnLevels + L) * kRpnCap;
  for (int i = tid; i < nk; i += 1024) {
    const int src = kept[i];
    const float4 v = sbox[base + src];
    ob[i * 4] = v.x, ob[i * 4 + 1] = v.y, ob[i * 4 + 2] = v.z, ob[i * 4 + 3] = v.w;
    os[i] = sscore[base + src];
  }
  if (tid == 0) lvl_count[b * kRpnLevels + L] = nk;
}

// Cross-level merge: keep = batched_nms(...)[:post_topk] orders all survivors by score (stable: level, then
// rank).  Every level list is already sorted, so the global rank is a sum of binary searches.
__global__ void __launch_bounds__(1024) k_rpn_merge(int post_topk, const float* __restrict__ lvl_boxes,
                                                    const float* __restrict__ lvl_scores, const int* __restrict__ lvl_count,
                                                    float* __restrict__ prop_boxes, float* __restrict__ prop_scores,
                                                    int* __restrict__ prop_img, int* __restrict__ prop_count) {
  pdl_grid_sync();
  const int b = blockIdx.x;
  __shared__ int cnt[kRpnLevels];
  if (threadIdx.x < kRpnLevels) cnt[threadIdx.x] = lvl_count[b * kRpnLevels + threadIdx.x];
  __syncthreads();
  int total = 0;
  for (int l = 0; l < kRpnLevels; ++l) total += cnt[l];
  const int nout = min(total, post_topk);
  for (int e = threadIdx.x; e < kRpnLevels * kRpnCap; e += blockDim.x) {
    const int L = e / kRpnCap, r = e - L * kRpnCap;
    if (r >= cnt[L]) continue;
    const float s = lvl_scores[(static_cast<size_t>(b) * kRpnLevels + L) * kRpnCap + r];
    int rank = r;
    for (int l2 = 0; l2 < kRpnLevels; ++l2) {
      if (l2 == L) continue;
      const float* sc = lvl_scores + (static_cast<size_t>(b) * kRpnLevels + l2) * kRpnCap;
      // number of entries of level l2 ordered before (s, L): score > s, or == s when l2 < L
      int lo = 0, hi = cnt[l2];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        const float v = sc[mid];
        const bool before = (l2 < L) ? (v >= s) : (v > s);
        if (before) lo = mid + 1;
        else hi = mid;
      }
      rank += lo;
    }
    if (rank < post_topk) {
      const float* src = lvl_boxes + ((static_cast<size_t>(b) * kRpnLevels + L) * kRpnCap + r) * 4;
      float* dst = prop_boxes + (static_cast<size_t>(b) * post_topk + rank) * 4;
      dst[0] = src[0], dst[1] = src[1], dst[2] = src[2], dst[3] = src[3];
      prop_scores[static_cast<size_t>(b) * post_topk + rank] = s;
      prop_img[static_cast<size_t>(b) * post_topk + rank] = b;
    }
  }
  for (int r = nout + threadIdx.x; r < post_topk; r += blockDim.x) prop_img[static_cast<size_t>(b) * post_topk + r] = -1;
  if (threadIdx.x == 0) prop_count[b] = nout;
}

// ---------------------------------------------------------------------------------------------------
// ROIAlignV2 (aligned=True, sampling_ratio=0) with detectron2's level assignment.  One CTA per ROI, one warp
// per output bin COLUMN (it walks the column's S bins), one lane per 8 channels (256 channels).
// 8 channels of one cell as loaded (conversion deferred until the value is consumed, so that several loads stay in flight)
template <typename T>
struct Raw8;
template <>
struct Raw8<__nv_bfloat16> {
  uint4 r;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { r = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      v[2 * e] = __uint_as_float(w[e] << 16);
      v[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
    }
  }
};
template <>
struct Raw8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = *reinterpret_cast<const float4*>(p);
    b = *reinterpret_cast<const float4*>(p + 4);
  }
  __device__ __forceinline__ void unpack(float (&v)[8]) const {
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
  }
};


// Geometry of one ROI on its pyramid level (torchvision roi_align.cpp with aligned=True, sampling_ratio=0; level rule of
// detectron2 assign_boxes_to_levels).
struct RoiGeom {
  PyramidLevel lv;
  float rsw, rsh, bin_h, bin_w, count;
  int gh, gw;
  __device__ __forceinline__ float sample_y(int ph, int iy) const {
    return rsh + static_cast<float>(ph) * bin_h + (static_cast<float>(iy) + .5f) * bin_h / static_cast<float>(gh);
  }
  __device__ __forceinline__ float sample_x(int pw, int ix) const {
    return rsw + static_cast<float>(pw) * bin_w + (static_cast<float>(ix) + .5f) * bin_w / static_cast<float>(gw);
  }
};
__device__ __forceinline__ RoiGeom roi_geom(const Pyramid& pyr, const float* __restrict__ boxes, int r, int S) {
  RoiGeom g;
  const float x1 = boxes[r * 4], y1 = boxes[r * 4 + 1], x2 = boxes[r * 4 + 2], y2 = boxes[r * 4 + 3];
  const float size = sqrtf((x2 - x1) * (y2 - y1));
  float lvf = floorf(4.f + log2f(size / 224.f + 1e-8f));
  lvf = fminf(fmaxf(lvf, 2.f), 5.f);
  const int li = static_cast<int>(lvf) - 2;   // selects instead of a dynamic index: the parameter struct stays in the constant bank
  g.lv = pyr.lv[0];
  if (li == 1) g.lv = pyr.lv[1];
  if (li == 2) g.lv = pyr.lv[2];
  if (li == 3) g.lv = pyr.lv[3];
  g.rsw = x1 * g.lv.scale - 0.5f, g.rsh = y1 * g.lv.scale - 0.5f;
  const float rew = x2 * g.lv.scale - 0.5f, reh = y2 * g.lv.scale - 0.5f;
  const float rw = rew - g.rsw, rh = reh - g.rsh;
  g.bin_h = rh / static_cast<float>(S), g.bin_w = rw / static_cast<float>(S);
  g.gh = static_cast<int>(ceilf(rh / static_cast<float>(S)));
  g.gw = static_cast<int>(ceilf(rw / static_cast<float>(S)));
  g.count = static_cast<float>(max(g.gh * g.gw, 1));
  return g;
}

// One bin, one sample at a time in torchvision's order: for bins wider than 32 feature cells, which detectron2's level
// assignment never produces (recomputes the geometry so that the hot loop does not carry its registers).
template <typename T>
__device__ __forceinline__ void roi_bin_per_sample(const Pyramid& pyr, const float* __restrict__ boxes, int r, int b, int S, int ph, int pw,
                                                int lane, float (&acc)[8]) {
  const RoiGeom g = roi_geom(pyr, boxes, r, S);
  const PyramidLevel& lv = g.lv;
  const T* feat = static_cast<const T*>(lv.ptr) + static_cast<size_t>(b) * lv.H * lv.W * lv.ld;
  const float fh = static_cast<float>(lv.H), fw = static_cast<float>(lv.W);
  const int ns = g.gh * g.gw;
  for (int s0 = 0; s0 < ns; ++s0) {
    float y = g.sample_y(ph, s0 / g.gw), x = g.sample_x(pw, s0 % g.gw);
    if (y < -1.0f || y > fh || x < -1.0f || x > fw) continue;
    if (y <= 0.f) y = 0.f;
    if (x <= 0.f) x = 0.f;
    int yl = static_cast<int>(y), xl = static_cast<int>(x), yh, xh;
    if (yl >= lv.H - 1) yh = yl = lv.H - 1, y = static_cast<float>(yl);
    else yh = yl + 1;
    if (xl >= lv.W - 1) xh = xl = lv.W - 1, x = static_cast<float>(xl);
    else xh = xl + 1;
    const float ly = y - static_cast<float>(yl), lx = x - static_cast<float>(xl);
    const float hy = 1.f - ly, hx = 1.f - lx;
    const float w1 = hy * hx, w2 = hy * lx, w3 = ly * hx, w4 = ly * lx;
    // acc += ((w1 v1 + w2 v2) + w3 v3) + w4 v4, tap by tap
    float t[8], v[8];
    load8(feat + (static_cast<size_t>(yl) * lv.W + xl) * lv.ld + lane * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = w1 * v[j];
    load8(feat + (static_cast<size_t>(yl) * lv.W + xh) * lv.ld + lane * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = t[j] + w2 * v[j];
    load8(feat + (static_cast<size_t>(yh) * lv.W + xl) * lv.ld + lane * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = t[j] + w3 * v[j];
    load8(feat + (static_cast<size_t>(yh) * lv.W + xh) * lv.ld + lane * 8, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = acc[j] + (t[j] + w4 * v[j]);
  }
}

// acc / count -> (tf32 rounding for fp32 storage) -> store, 8 channels.
template <typename T>
__device__ __forceinline__ void roi_bin_store(float (&acc)[8], float count, T* dst, bool rnd) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    acc[j] = acc[j] / count;
    if (sizeof(T) == 4 && rnd) {
      uint32_t q;
      asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(q) : "f"(acc[j]));
      acc[j] = __uint_as_float(q);
    }
  }
  store8(dst, acc);
}

// The S bins of one bin column whose footprint is kNC feature columns wide (kNC known at compile time: the column weights
// sit in registers, the loads of a row - of two rows for narrow bf16 footprints - are issued back to back without
// predicates, and a cell costs one load, one multiply, the unpack and eight FMAs).  ncu on the predicated six-slot loop
// it replaces (profiles/r02_roi_align_per_roi_b8.txt): 230 instructions per row for 3.5 live cells, the kernel issue-bound
// at 2.9 of 4 instructions per cycle.  Zero weights are not skipped: fma(0, v, acc) == acc, so the sums are the same.
template <typename T, int kNC>
__device__ __forceinline__ void roi_bins_fixed(const float (*s_wy)[32], const int* s_rlo, const int* s_nrows, float WX, const T* colp,
                                               size_t row_stride, size_t ld, int S, float count, T* outp, long long ldo, bool rnd) {
  constexpr int kRows = (sizeof(T) == 2 && kNC <= 4) ? 2 : 1;
  float wx[kNC];
#pragma unroll
  for (int u = 0; u < kNC; ++u) wx[u] = __shfl_sync(0xffffffffu, WX, u);
  for (int ph = 0; ph < S; ++ph) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int nrows = s_nrows[ph];
    const T* rowp = colp + static_cast<size_t>(s_rlo[ph]) * row_stride;
    const float* wyp = s_wy[ph];
    int rr = 0;
    for (; rr + kRows <= nrows; rr += kRows, rowp += kRows * row_stride) {
      Raw8<T> v[kRows][kNC];
#pragma unroll
      for (int q = 0; q < kRows; ++q) {
#pragma unroll
        for (int u = 0; u < kNC; ++u) v[q][u].load(rowp + q * row_stride + u * ld);
      }
#pragma unroll
      for (int q = 0; q < kRows; ++q) {
        const float wy = wyp[rr + q];
#pragma unroll
        for (int u = 0; u < kNC; ++u) {
          const float w = wy * wx[u];
          float val[8];
          v[q][u].unpack(val);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = __fmaf_rn(w, val[j], acc[j]);   // one rounding per term (-fmad=false file)
        }
      }
    }
    if (kRows == 2 && rr < nrows) {   // odd row count: the last row alone
      Raw8<T> v[kNC];
#pragma unroll
      for (int u = 0; u < kNC; ++u) v[u].load(rowp + u * ld);
      const float wy = wyp[rr];
#pragma unroll
      for (int u = 0; u < kNC; ++u) {
        const float w = wy * wx[u];
        float val[8];
        v[u].unpack(val);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = __fmaf_rn(w, val[j], acc[j]);
      }
    }
    roi_bin_store<T>(acc, count, outp + static_cast<size_t>(ph) * S * ldo, rnd);
  }
}

// grid = ROIs; block = S warps.  Warp w first builds the ROW table of bin row w (which feature rows the bin row's samples
// touch and with what summed weight - shared by the S bins of that row) in shared memory and the COLUMN table of bin column
// w in its registers, then walks the S bins of column w.  The coordinate arithmetic (level choice, divisions, the sample
// loops) is therefore done once per bin row / column and ROI instead of once per bin.
//
// Bilinear sampling is separable and the bin is a sum over a regular sample grid, so
//   bin = sum_rows sum_cols WY[row] * WX[col] * feat[row][col],  WY[row] = sum over the bin's sample rows of their
// weight on that feature row (same for WX): every feature cell under the bin is read ONCE instead of once per
// (sample, tap).  Lane l accumulates the weight of row rlo + l / column clo + l; bins spanning more than 32 rows or
// columns take the per-sample path.  Accumulation order: rows outer, columns inner, one FMA per term.
template <typename T>
__global__ void __launch_bounds__(448, 2) k_roi_align(Pyramid pyr, const float* __restrict__ boxes, const int* __restrict__ img,
                                                   int S, T* __restrict__ out, long long ldo) {
  pdl_grid_sync();
  const int r = blockIdx.x;
  const int b = img[r];
  if (b < 0) return;
  const bool rnd = pyr.keep_fp32 == 0;
  __shared__ float s_wy[14][32];   // [bin row][feature row - rlo]
  __shared__ int s_rlo[14], s_nrows[14];
  __shared__ int s_rows_bad;       // some bin row spans more than 32 feature rows (or the ROI has no samples)
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) s_rows_bad = 0;
  __syncthreads();
  int ncols;
  bool cols_ok;
  float WX = 0.f, count;
  size_t row_stride, ld;
  const T* colp;
  {
    const RoiGeom g = roi_geom(pyr, boxes, r, S);
    const PyramidLevel& lv = g.lv;
    const int ns = g.gh * g.gw;
    const float fh = static_cast<float>(lv.H), fw = static_cast<float>(lv.W);
    {  // row table of bin row `wid`
      const int ph = wid;
      const float yf = fminf(fmaxf(g.sample_y(ph, 0), 0.f), fh), yl_ = fminf(fmaxf(g.sample_y(ph, g.gh - 1), 0.f), fh);
      const int rlo = min(static_cast<int>(yf), lv.H - 1), rhi = min(static_cast<int>(yl_) + 1, lv.H - 1);
      const bool rows_ok = ns > 0 && rhi - rlo < 32;
      float WY = 0.f;
      if (rows_ok) {
        for (int iy = 0; iy < g.gh; ++iy) {
          float y = g.sample_y(ph, iy);
          if (y < -1.0f || y > fh) continue;
          if (y <= 0.f) y = 0.f;
          int yl = static_cast<int>(y), yh;
          if (yl >= lv.H - 1) yh = yl = lv.H - 1, y = static_cast<float>(yl);
          else yh = yl + 1;
          const float ly = y - static_cast<float>(yl), hy = 1.f - ly;
          if (rlo + lane == yl) WY += hy;
          if (rlo + lane == yh) WY += ly;
        }
      }
      s_wy[ph][lane] = WY;
      if (lane == 0) {
        s_rlo[ph] = rlo, s_nrows[ph] = rows_ok ? rhi - rlo + 1 : -1;
        if (!rows_ok) s_rows_bad = 1;
      }
    }
    // column table of bin column `wid`
    const int pw = wid;
    const float xf = fminf(fmaxf(g.sample_x(pw, 0), 0.f), fw), xl_ = fminf(fmaxf(g.sample_x(pw, g.gw - 1), 0.f), fw);
    const int clo = min(static_cast<int>(xf), lv.W - 1), chi = min(static_cast<int>(xl_) + 1, lv.W - 1);
    cols_ok = ns > 0 && chi - clo < 32;
    if (cols_ok) {
      for (int ix = 0; ix < g.gw; ++ix) {
        float x = g.sample_x(pw, ix);
        if (x < -1.0f || x > fw) continue;
        if (x <= 0.f) x = 0.f;
        int xl = static_cast<int>(x), xh;
        if (xl >= lv.W - 1) xh = xl = lv.W - 1, x = static_cast<float>(xl);
        else xh = xl + 1;
        const float lx = x - static_cast<float>(xl), hx = 1.f - lx;
        if (clo + lane == xl) WX += hx;
        if (clo + lane == xh) WX += lx;
      }
    }
    ncols = chi - clo + 1;
    ld = static_cast<size_t>(lv.ld);
    row_stride = static_cast<size_t>(lv.W) * ld;
    colp = static_cast<const T*>(lv.ptr) + static_cast<size_t>(b) * lv.H * row_stride + static_cast<size_t>(clo) * ld + lane * 8;
    count = g.count;
  }
  __syncthreads();
  const int pw = wid;
  T* outp = out + (static_cast<size_t>(r) * S * S + pw) * ldo + lane * 8;
  if (cols_ok && s_rows_bad == 0 && ncols <= 8) {
    switch (ncols) {
      case 1: roi_bins_fixed<T, 1>(s_wy, s_rlo, s_nrows, WX, colp, row_stride, ld, S, count, outp, ldo, rnd); break;
      case 2: roi_bins_fixed<T, 2>(s_wy, s_rlo, s_nrows, WX, colp, row_stride, ld, S, count, outp, ldo, rnd); break;
      case 3: roi_bins_fixed<T, 3>(s_wy, s_rlo, s_nrows, WX, colp, row_stride, ld, S, count, outp, ldo, rnd); break;
      case 4: roi_bins_fixed<T, 4>(s_wy, s_rlo, s_nrows, WX, colp, row_stride, ld, S, count, outp, ldo, rnd); break;
      case 5: roi_bins_fixed<T, 5>(s_wy, s_rlo, s_nrows, WX, colp, row_stride, ld, S, count, outp, ldo, rnd); break;
      case 6: roi_bins_fixed<T, 6>(s_wy, s_rlo, s_nrows, WX, colp, row_stride, ld, S, count, outp, ldo, rnd); break;
      case 7: roi_bins_fixed<T, 7>(s_wy, s_rlo, s_nrows, WX, colp, row_stride, ld, S, count, outp, ldo, rnd); break;
      default: roi_bins_fixed<T, 8>(s_wy, s_rlo, s_nrows, WX, colp, row_stride, ld, S, count, outp, ldo, rnd); break;
    }
    return;
  }
  for (int ph = 0; ph < S; ++ph) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const int nrows = s_nrows[ph];
    if (nrows > 0 && cols_ok) {
      // wide footprints (more than 8 feature columns per bin): predicated batches of kFly cells
      constexpr int kFly = sizeof(T) == 2 ? 6 : 3;
      const T* rowp = colp + static_cast<size_t>(s_rlo[ph]) * row_stride;
      for (int rr = 0; rr < nrows; ++rr, rowp += row_stride) {
        const float wy = s_wy[ph][rr];
        for (int c0 = 0; c0 < ncols; c0 += kFly) {
          float w[kFly];
          Raw8<T> v[kFly];
#pragma unroll
          for (int u = 0; u < kFly; ++u) {
            const int cc = c0 + u;
            const float wx = __shfl_sync(0xffffffffu, WX, cc & 31);
            w[u] = (cc < ncols) ? wy * wx : 0.f;
            if (cc < ncols) v[u].load(rowp + static_cast<size_t>(cc) * ld);
          }
#pragma unroll
          for (int u = 0; u < kFly; ++u) {
            if (c0 + u < ncols) {
              float val[8];
              v[u].unpack(val);
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[j] = __fmaf_rn(w[u], val[j], acc[j]);
            }
          }
        }
      }
    } else {
      roi_bin_per_sample<T>(pyr, boxes, r, b, S, ph, pw, lane, acc);
    }
    roi_bin_store<T>(acc, count, outp + static_cast<size_t>(ph) * S * ldo, rnd);
  }
}

// ---------------------------------------------------------------------------------------------------
// fast_rcnn_inference_single_image: softmax, score filter, class-wise decode + clip, per-class NMS in score
// order, first `detections` survivors.  One CTA per image; dynamic smem = candidate keys.
__global__ void __launch_bounds__(1024, 1) k_detections(const MrcnnSlots* __restrict__ slots, int K, int post_topk, int max_det,
                                                        float nms_thr, float img_h, float img_w,
                                                        const float* __restrict__ box_out, int ldb,
                                                        const float* __restrict__ prop_boxes, const int* __restrict__ prop_count,
                                                        float* __restrict__ det_boxes, float* __restrict__ det_scores,
                                                        int* __restrict__ det_classes, int* __restrict__ det_count, int cap,
                                                        float4* __restrict__ cand_boxes, int* __restrict__ cand_cls) {
  pdl_grid_sync();
  extern __shared__ __align__(16) unsigned long long dkeys[];  // [cap]
  __shared__ int s_n;
  __shared__ float4 kbox[128];
  __shared__ int kcls[128];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float thr = slots->score_thresh;
  const int R = prop_count[b];
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int r = tid; r < R; r += blockDim.x) {
    const float* row = box_out + (static_cast<size_t>(b) * post_topk + r) * ldb;
    float mx = row[0];
    for (int c = 1; c <= K; ++c) mx = fmaxf(mx, row[c]);
    float e[16], sum = 0.f;
    for (int c = 0; c <= K; ++c) e[c] = expf(row[c] - mx), sum += e[c];
    bool finite_row = isfinite(sum);
    for (int c = 0; c < K && finite_row; ++c) {
      const float p = e[c] / sum;
      if (p > thr) {
        const int slot = atomicAdd(&s_n, 1);
        if (slot < cap)
          dkeys[slot] = (static_cast<unsigned long long>(fkey(p)) << 32) | (0xffffffffu - static_cast<uint32_t>(r * K + c));
      }
    }
  }
  __syncthreads();
  const int n = min(s_n, cap);
  int npow = 1;
  while (npow < n) npow <<= 1;
  for (int i = n + tid; i < npow; i += blockDim.x) dkeys[i] = 0ull;
  __syncthreads();
  if (npow > 1) bitonic_desc(dkeys, npow);
  // decode + clip every candidate in parallel (score order) into the scratch list
  float4* cb = cand_boxes + static_cast<size_t>(b) * cap;
  int* cc = cand_cls + static_cast<size_t>(b) * cap;
  for (int c = tid; c < n; c += blockDim.x) {
    const unsigned long long kk = dkeys[c];
    const int flat = static_cast<int>(0xffffffffu - static_cast<uint32_t>(kk & 0xffffffffull));
    const int r = flat / K, cls = flat - r * K;
    const float* row = box_out + (static_cast<size_t>(b) * post_topk + r) * ldb;
    const float* pb = prop_boxes + (static_cast<size_t>(b) * post_topk + r) * 4;
    const float anchor[4] = {pb[0], pb[1], pb[2], pb[3]};
    const float* d = row + (K + 1) + cls * 4;
    float o[4];
    decode_box(anchor, d[0], d[1], d[2], d[3], 10.f, 10.f, 5.f, 5.f, o);
    // (valid_mask of the reference drops whole rows with a non-finite box of ANY class; a non-finite decoded box
    // here can only come from non-finite deltas, which also poison that row's other classes)
    const bool finite_box = isfinite(o[0]) && isfinite(o[1]) && isfinite(o[2]) && isfinite(o[3]);
    cb[c] = make_float4(clampf(o[0], 0.f, img_w), clampf(o[1], 0.f, img_h), clampf(o[2], 0.f, img_w), clampf(o[3], 0.f, img_h));
    cc[c] = finite_box ? cls : -1;
  }
  __syncthreads();
  // greedy per-class NMS in score order by warp 0, 32 candidates at a time: each lane first tests its candidate
  // against the boxes kept so far, then the chunk is resolved in order with shuffles.
  if (tid < 32) {
    const int lane = tid;
    int nk = 0;
    for (int c0 = 0; c0 < n && nk < max_det; c0 += 32) {
      const int c = c0 + lane;
      float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
      int cls = -1;
      if (c < n) bx = cb[c], cls = cc[c];
      bool alive = cls >= 0;
      for (int j = 0; j < nk && alive; ++j)
        if (kcls[j] == cls && iou_gt(kbox[j], bx, nms_thr)) alive = false;
      for (int j = 0; j < 32; ++j) {
        const bool j_alive = __shfl_sync(0xffffffffu, alive ? 1 : 0, j) != 0;
        if (!j_alive) continue;  // warp-uniform
        float4 o;
        o.x = __shfl_sync(0xffffffffu, bx.x, j), o.y = __shfl_sync(0xffffffffu, bx.y, j);
        o.z = __shfl_sync(0xffffffffu, bx.z, j), o.w = __shfl_sync(0xffffffffu, bx.w, j);
        const int ocls = __shfl_sync(0xffffffffu, cls, j);
        if (lane > j && alive && cls == ocls && iou_gt(o, bx, nms_thr)) alive = false;
      }
      const unsigned keep = __ballot_sync(0xffffffffu, alive);
      const int pos = nk + __popc(keep & ((1u << lane) - 1u));
      if (alive && pos < max_det) {
        kbox[pos] = bx, kcls[pos] = cls;
        float* ob = det_boxes + (static_cast<size_t>(b) * max_det + pos) * 4;
        ob[0] = bx.x, ob[1] = bx.y, ob[2] = bx.z, ob[3] = bx.w;
        det_scores[static_cast<size_t>(b) * max_det + pos] = fkey_inv(static_cast<uint32_t>(dkeys[c] >> 32));
        det_classes[static_cast<size_t>(b) * max_det + pos] = cls;
      }
      nk = min(nk + __popc(keep), max_det);
      __syncwarp();
    }
    if (lane == 0) det_count[b] = nk;
  }
}

// Batch-compacted ROI list for the mask head + per-image offsets.  One block.
__global__ void k_compact_dets(int B, int max_det, const float* __restrict__ det_boxes, const int* __restrict__ det_classes,
                               const int* __restrict__ det_count, float* __restrict__ mroi_boxes, int* __restrict__ mroi_img,
                               int* __restrict__ mroi_cls, int* __restrict__ mroi_total) {
  pdl_grid_sync();
  __shared__ int start[1025];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b < B; ++b) start[b] = acc, acc += det_count[b];
    start[B] = acc;
    mroi_total[0] = acc;
    for (int b = 0; b <= B; ++b) mroi_total[1 + b] = start[b];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < B * max_det; e += blockDim.x) {
    const int b = e / max_det, i = e - b * max_det;
    if (i < det_count[b]) {
      const int dst = start[b] + i;
      for (int q = 0; q < 4; ++q) mroi_boxes[dst * 4 + q] = det_boxes[e * 4 + q];
      mroi_img[dst] = b;
      mroi_cls[dst] = det_classes[e];
    }
    if (e >= start[B]) mroi_img[e] = -1;
  }
}

// ---------------------------------------------------------------------------------------------------
// detector_postprocess + paste_masks_in_image + the accumulation loop of segmentation.py:47-62.
//   k_paste_prepare  per frame: zero the output stack, apply the score gates, rescale the boxes to the camera frame,
//                    clip, drop empty -> active list;
//   k_paste_dets     kPasteSplit CTAs per active detection walk the pixels of its box (+1 px halo): the 28x28 mask
//                    probabilities (sigmoid once per CTA into smem) are bilinearly sampled (grid_sample,
//                    align_corners=False, zero padding) at the pixel centre; >= threshold adds 1 to the class channel
//                    (atomicAdd of small integers in fp32: exact and order-independent).
constexpr int kPasteSplit = 4;

__global__ void __launch_bounds__(256) k_paste_prepare(const MrcnnSlots* __restrict__ slots, int H, int W, int K, int max_det,
                                                       float sx, float sy, const float* __restrict__ det_scores,
                                                       const int* __restrict__ det_count, const int* __restrict__ mroi_total,
                                                       const float* __restrict__ mroi_boxes, const int* __restrict__ mroi_cls,
                                                       float4* __restrict__ act_box, int* __restrict__ act_cls,
                                                       int* __restrict__ act_roi, int* __restrict__ act_n) {
  pdl_grid_sync();
  const int b = blockIdx.y;
  // zero this frame's output stack (grid.x CTAs share the work)
  float4* out4 = reinterpret_cast<float4*>(slots->sem_out + static_cast<size_t>(b) * H * W * (K + 1));
  const size_t n4 = static_cast<size_t>(H) * W * (K + 1) / 4;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4; i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (blockIdx.x != 0 || threadIdx.x >= 32) return;
  // gates + rescale + clip + drop empty, order preserved (one warp, ballot compaction)
  const int lane = threadIdx.x;
  const int n_det = det_count[b];
  const int start = mroi_total[1 + b];
  const float sem_thr = slots->sem_thr, goal_thr = slots->goal_thr;
  const int goal = slots->goal_cat ? slots->goal_cat[b] : -1;
  int n = 0;
  for (int i0 = 0; i0 < n_det; i0 += 32) {
    const int i = i0 + lane;
    bool ok = false;
    float4 bx = make_float4(0.f, 0.f, 0.f, 0.f);
    int cls = -1;
    if (i < n_det) {
      cls = mroi_cls[start + i];
      const float score = det_scores[static_cast<size_t>(b) * max_det + i];
      const float* pb = mroi_boxes + static_cast<size_t>(start + i) * 4;
      bx = make_float4(pb[0] * sx, pb[1] * sy, pb[2] * sx, pb[3] * sy);
      bx.x = clampf(bx.x, 0.f, static_cast<float>(W)), bx.z = clampf(bx.z, 0.f, static_cast<float>(W));
      bx.y = clampf(bx.y, 0.f, static_cast<float>(H)), bx.w = clampf(bx.w, 0.f, static_cast<float>(H));
      ok = (bx.z - bx.x) > 0.f && (bx.w - bx.y) > 0.f && cls >= 0 && cls < K && !(score < sem_thr) &&
           !(cls == goal && score < goal_thr);
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) {
      const int pos = n + __popc(m & ((1u << lane) - 1u));
      act_box[b * max_det + pos] = bx, act_cls[b * max_det + pos] = cls, act_roi[b * max_det + pos] = start + i;
    }
    n += __popc(m);
  }
  if (lane == 0) act_n[b] = n;
}

__global__ void __launch_bounds__(256) k_paste_dets(const MrcnnSlots* __restrict__ slots, int H, int W, int K, int max_det,
                                                    float mask_thr, const float4* __restrict__ act_box,
                                                    const int* __restrict__ act_cls, const int* __restrict__ act_roi,
                                                    const int* __restrict__ act_n, const float* __restrict__ mask_logits) {
  pdl_grid_sync();
  const int b = blockIdx.z, i = blockIdx.y;
  if (i >= act_n[b]) return;
  __shared__ float prob[28 * 28];
  const float4 bx = act_box[b * max_det + i];
  const int cls = act_cls[b * max_det + i];
  const float* ml = mask_logits + static_cast<size_t>(act_roi[b * max_det + i]) * (14 * 14 * 4) * 16 + cls;
  for (int p = threadIdx.x; p < 28 * 28; p += blockDim.x) {
    const int Y = p / 28, X = p - Y * 28;
    const int row = (((Y >> 1) * 14) + (X >> 1)) * 4 + ((Y & 1) << 1) + (X & 1);
    prob[p] = 1.f / (1.f + expf(-ml[static_cast<size_t>(row) * 16]));
  }
  __syncthreads();
  // pixels whose centre can map into (-1, 28) mask coordinates: the box widened by one mask cell, +1 px of slack
  const float cw = (bx.z - bx.x) / 28.f, ch = (bx.w - bx.y) / 28.f;
  const int xa = max(0, static_cast<int>(floorf(bx.x - cw)) - 1), xb = min(W, static_cast<int>(ceilf(bx.z + cw)) + 1);
  const int ya = max(0, static_cast<int>(floorf(bx.y - ch)) - 1), yb = min(H, static_cast<int>(ceilf(bx.w + ch)) + 1);
  const int bw = xb - xa, bh = yb - ya;
  if (bw <= 0 || bh <= 0) return;
  float* out = slots->sem_out + static_cast<size_t>(b) * H * W * (K + 1) + cls;
  const int total = bw * bh;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < total; p += gridDim.x * blockDim.x) {
    const int y = ya + p / bw, x = xa + p % bw;
    const float px = static_cast<float>(x) + 0.5f, py = static_cast<float>(y) + 0.5f;
    const float gx = (px - bx.x) / (bx.z - bx.x) * 2.f - 1.f;
    const float gy = (py - bx.y) / (bx.w - bx.y) * 2.f - 1.f;
    const float ix = (gx + 1.f) * 14.f - 0.5f, iy = (gy + 1.f) * 14.f - 0.5f;  // ((g + 1) * 28 - 1) / 2
    if (!(ix > -1.f && ix < 28.f && iy > -1.f && iy < 28.f)) continue;
    const float xw = floorf(ix), yn = floorf(iy);
    const float w = ix - xw, e = 1.f - w, nn = iy - yn, s = 1.f - nn;
    const float wnw = s * e, wne = s * w, wsw = nn * e, wse = nn * w;
    const int X0 = static_cast<int>(xw), Y0 = static_cast<int>(yn);
    auto at = [&](int Y, int X) -> float { return (X < 0 || X >= 28 || Y < 0 || Y >= 28) ? 0.f : prob[Y * 28 + X]; };
    const float v = ((at(Y0, X0) * wnw + at(Y0, X0 + 1) * wne) + at(Y0 + 1, X0) * wsw) + at(Y0 + 1, X0 + 1) * wse;
    if (v >= mask_thr) atomicAdd(out + (static_cast<size_t>(y) * W + x) * (K + 1), 1.f);
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------
void add_upsample2x_add(Net& net, const Tensor& prev, const Tensor& lat) {
  PN_REQUIRE(prev.C == lat.C && prev.C % 8 == 0 && prev.dt == lat.dt && prev.B == lat.B, "upsample_add: shape");
  PN_REQUIRE(lat.H == 2 * prev.H && lat.W == 2 * prev.W, "upsample_add: the fine level must be exactly 2x the coarse one");
  const int C8 = lat.C / 8;
  const long long total = lat.pixels() * C8;
  check_u32_launch(total, "fpn elementwise kernel");
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads);
  Tensor p = prev, l = lat;
  const int rnd = net.x3 ? 0 : 1;
  net.add("fpn_upsample_add", [=](cudaStream_t s) {
    if (l.dt == kBF16)
      launch_pdl(k_upsample2x_add<__nv_bfloat16>, blocks, threads, 0, s, static_cast<const __nv_bfloat16*>(p.ptr), p.ld, p.H, p.W, static_cast<__nv_bfloat16*>(l.ptr), l.ld, l.B, l.H, l.W, C8, 0);
    else
      launch_pdl(k_upsample2x_add<float>, blocks, threads, 0, s, static_cast<const float*>(p.ptr), p.ld, p.H, p.W, static_cast<float*>(l.ptr), l.ld, l.B, l.H, l.W, C8, rnd);
  });
  net.launches_per_forward += 1;
}

void add_subsample2(Net& net, const Tensor& in, const Tensor& out) {
  PN_REQUIRE(out.H == (in.H - 1) / 2 + 1 && out.W == (in.W - 1) / 2 + 1 && in.C == out.C && in.C % 8 == 0, "subsample2: shape");
  const int C8 = in.C / 8;
  const long long total = out.pixels() * C8;
  check_u32_launch(total, "fpn elementwise kernel");
  const int threads = 256;
  const int blocks = static_cast<int>((total + threads - 1) / threads);
  Tensor i = in, o = out;
  net.add("fpn_p6_subsample", [=](cudaStream_t s) {
    if (i.dt == kBF16)
      launch_pdl(k_subsample2<__nv_bfloat16>, blocks, threads, 0, s, static_cast<const __nv_bfloat16*>(i.ptr), i.ld, i.H, i.W, static_cast<__nv_bfloat16*>(o.ptr), o.ld, o.B, o.H, o.W, C8);
    else
      launch_pdl(k_subsample2<float>, blocks, threads, 0, s, static_cast<const float*>(i.ptr), i.ld, i.H, i.W, static_cast<float*>(o.ptr), o.ld, o.B, o.H, o.W, C8);
  });
  net.launches_per_forward += 1;
}

void add_rpn_proposals(Net& net, MaskRcnn& m, const RpnMeta& meta) {
  PN_REQUIRE(meta.pre_topk <= kRpnCap && meta.post_topk >= 1, "rpn: pre_nms_topk exceeds the per-level capacity");
  const size_t smem = (sizeof(RpnSmem) + 15) & ~size_t(15);
  const size_t smem_mask = static_cast<size_t>(kRpnCap) * 32 * sizeof(uint32_t);
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_rpn_sort_decode, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_rpn_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem_mask)));
  const int B = m.cfg.B;
  float* lb = m.lvl_boxes;
  float* ls = m.lvl_scores;
  int* lc = m.lvl_count;
  uint32_t* hist = static_cast<uint32_t*>(net.arena.alloc(static_cast<size_t>(B) * kRpnLevels * kHistBins * sizeof(uint32_t)));
  uint32_t* thr = static_cast<uint32_t*>(net.arena.alloc(static_cast<size_t>(B) * kRpnLevels * sizeof(uint32_t)));
  int* ncand = static_cast<int*>(net.arena.alloc(static_cast<size_t>(B) * kRpnLevels * sizeof(int)));
  unsigned long long* cand = static_cast<unsigned long long*>(net.arena.alloc(static_cast<size_t>(B) * kRpnLevels * kCandCap * sizeof(unsigned long long)));
  int max_pix = 0;
  for (int l = 0; l < kRpnLevels; ++l) max_pix = std::max(max_pix, meta.lv[l].H * meta.lv[l].W);
  const dim3 scan_grid((max_pix + kPixPerBlock - 1) / kPixPerBlock, kRpnLevels * B);
  const size_t hist_bytes = static_cast<size_t>(B) * kRpnLevels * kHistBins * sizeof(uint32_t);
  net.add("rpn_preselect", [=](cudaStream_t s) {
    PN_CUDA_CHECK(cudaMemsetAsync(hist, 0, hist_bytes, s));
    launch_pdl(k_rpn_hist, scan_grid, 256, 0, s, meta, hist);
    launch_pdl(k_rpn_threshold, dim3(kRpnLevels, B), 1024, 0, s, meta, hist, thr, ncand);
    launch_pdl(k_rpn_collect, scan_grid, 256, 0, s, meta, thr, ncand, cand);
  });
  net.launches_per_forward += 4;
  float4* sbox = static_cast<float4*>(net.arena.alloc(static_cast<size_t>(B) * kRpnLevels * kRpnCap * sizeof(float4)));
  float* sscore = static_cast<float*>(net.arena.alloc(static_cast<size_t>(B) * kRpnLevels * kRpnCap * sizeof(float)));
  int* scount = static_cast<int*>(net.arena.alloc(static_cast<size_t>(B) * kRpnLevels * sizeof(int)));
  uint32_t* mask_g = static_cast<uint32_t*>(net.arena.alloc(static_cast<size_t>(B) * kRpnLevels * kRpnCap * 32 * sizeof(uint32_t)));
  const float nms_thr = meta.nms_thr;
  net.add("rpn_sort_decode", [=](cudaStream_t s) { launch_pdl(k_rpn_sort_decode, dim3(kRpnLevels, B), 1024, smem, s, meta, ncand, cand, sbox, sscore, scount); });
  net.add("rpn_nms", [=](cudaStream_t s) {
    launch_pdl(k_rpn_mask, dim3(kRpnCap / 32, kRpnLevels, B), 256, 0, s, nms_thr, sbox, scount, mask_g);
    launch_pdl(k_rpn_sweep, dim3(kRpnLevels, B), 1024, smem_mask, s, sbox, sscore, scount, mask_g, lb, ls, lc);
  });
  net.launches_per_forward += 2;
  float* pb = m.prop_boxes;
  float* ps = m.prop_scores;
  int* pi = m.prop_img;
  int* pc = m.prop_count;
  const int post = meta.post_topk;
  net.add("rpn_merge", [=](cudaStream_t s) { launch_pdl(k_rpn_merge, B, 1024, 0, s, post, lb, ls, lc, pb, ps, pi, pc); });
  net.launches_per_forward += 2;
}

void add_roi_align(Net& net, const std::string& name, const Pyramid& pyr_in, DType dt, const float* boxes, const int* img,
                   int nrois, int S, const Tensor& out) {
  PN_REQUIRE(out.C == 256 && out.dt == dt && S <= 14, "roi_align: 256-channel pyramid, at most 14 x 14 bins expected");
  Tensor o = out;
  Pyramid pyr = pyr_in;
  pyr.keep_fp32 = net.x3 ? 1 : 0;
  net.add(name, [=](cudaStream_t s) {
    if (dt == kBF16)
      launch_pdl(k_roi_align<__nv_bfloat16>, dim3(nrois), S * 32, 0, s, pyr, boxes, img, S, static_cast<__nv_bfloat16*>(o.ptr), o.ld);
    else
      launch_pdl(k_roi_align<float>, dim3(nrois), S * 32, 0, s, pyr, boxes, img, S, static_cast<float*>(o.ptr), o.ld);
  });
  net.launches_per_forward += 1;
}

void add_detections(Net& net, MaskRcnn& m) {
  const MrcnnCfg c = m.cfg;
  PN_REQUIRE(c.num_classes + 1 <= 16 && c.detections <= 128, "detections: at most 15 classes / 128 detections");
  int cap = 1;
  while (cap < c.post_nms_topk * c.num_classes) cap <<= 1;
  const size_t smem = static_cast<size_t>(cap) * sizeof(unsigned long long);
  PN_REQUIRE(smem <= 200 * 1024, "detections: candidate list does not fit shared memory");
  PN_CUDA_CHECK(cudaFuncSetAttribute(k_detections, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  const MrcnnSlots* slots = m.slots;
  const float img_h = static_cast<float>(m.Hn), img_w = static_cast<float>(m.Wn);
  const float* box_out = m.box_out;
  const float* pb = m.prop_boxes;
  const int* pc = m.prop_count;
  float* db = m.det_boxes;
  float* ds = m.det_scores;
  int* dc = m.det_classes;
  int* dn = m.det_count;
  float4* cand_boxes = static_cast<float4*>(net.arena.alloc(static_cast<size_t>(c.B) * cap * sizeof(float4)));
  int* cand_cls = static_cast<int*>(net.arena.alloc(static_cast<size_t>(c.B) * cap * sizeof(int)));
  net.add("detections", [=](cudaStream_t s) {
    launch_pdl(k_detections, c.B, 1024, smem, s, slots, c.num_classes, c.post_nms_topk, c.detections, c.box_nms, img_h, img_w, box_out,
                                         64, pb, pc, db, ds, dc, dn, cap, cand_boxes, cand_cls);
  });
  float* mb = m.mroi_boxes;
  int* mi = m.mroi_img;
  int* mc = m.mroi_cls;
  int* mt = m.mroi_total;
  PN_REQUIRE(c.B <= 1024, "detections: batch too large");
  net.add("compact_dets", [=](cudaStream_t s) { launch_pdl(k_compact_dets, 1, 256, 0, s, c.B, c.detections, db, dc, dn, mb, mi, mc, mt); });
  net.launches_per_forward += 2;
}

void add_paste_accumulate(Net& net, MaskRcnn& m) {
  const MrcnnCfg c = m.cfg;
  const MrcnnSlots* slots = m.slots;
  PN_REQUIRE((static_cast<long long>(c.H) * c.W * (c.num_classes + 1)) % 4 == 0, "paste: frame size must allow float4 stores");
  // boxes[:, 0::2] *= scale_x with the python float cast to fp32 (detector_postprocess)
  const float sx = static_cast<float>(static_cast<double>(c.W) / m.Wn), sy = static_cast<float>(static_cast<double>(c.H) / m.Hn);
  const float* ds = m.det_scores;
  const int* dn = m.det_count;
  const int* mt = m.mroi_total;
  const float* mb = m.mroi_boxes;
  const int* mc = m.mroi_cls;
  const float* ml = m.mask_logits;
  float4* ab = static_cast<float4*>(net.arena.alloc(static_cast<size_t>(c.B) * c.detections * sizeof(float4)));
  int* ac = static_cast<int*>(net.arena.alloc(static_cast<size_t>(c.B) * c.detections * sizeof(int)));
  int* ar = static_cast<int*>(net.arena.alloc(static_cast<size_t>(c.B) * c.detections * sizeof(int)));
  int* an = static_cast<int*>(net.arena.alloc(static_cast<size_t>(c.B) * sizeof(int)));
  net.add("paste_prepare", [=](cudaStream_t s) {
    launch_pdl(k_paste_prepare, dim3(64, c.B), 256, 0, s, slots, c.H, c.W, c.num_classes, c.detections, sx, sy, ds, dn, mt, mb, mc, ab, ac, ar, an);
  });
  net.add("paste_dets", [=](cudaStream_t s) {
    launch_pdl(k_paste_dets, dim3(kPasteSplit, c.detections, c.B), 256, 0, s, slots, c.H, c.W, c.num_classes, c.detections, c.mask_thresh,
                                                                     ab, ac, ar, an, ml);
  });
  net.launches_per_forward += 2;
}

}  // namespace pn
