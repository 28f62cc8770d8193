// Implicit-GEMM convolution / GEMM on the sm_100a tensor cores.
//
//   D[M, Cout] = epilogue( sum_{r,s,c} A[(n,p,q) -> (n, p*stride - pad + r*dil, q*stride - pad + s*dil, c)] * W[cout, r, s, c] )
//
// * activations are NHWC (channel stride 1); the A operand is fetched by TMA in im2col mode
//   (one 128-pixel x BLOCK_K-channel box per filter tap and channel block, hardware zero-fill for the halo),
//   or in plain tiled 2-D mode for matrices;
// * weights are packed [Cout_pad][R*S*Cin_pad] (K-major) and fetched by tiled 2-D TMA;
// * tcgen05.mma (M=128, N=BN, K=32 bytes) accumulates in TMEM, two accumulator stages so that the
//   epilogue of tile i overlaps the main loop of tile i+1; the kernel is persistent (grid = #SMs);
// * the epilogue adds the per-channel bias (BatchNorm is folded: its scale into the weights, its shift into the bias),
//   the residual and the ReLU, and writes NHWC.
//   Fast path: the residual tile is prefetched by TMA into shared memory by a dedicated warp (it does not
//   depend on the accumulators, so it runs ahead of the MMAs), results are staged in swizzled shared
//   memory and written back with TMA bulk stores (full 128-byte lines, M/N tails clipped by hardware).
//   Fallback path (narrow or mixed-precision outputs): direct 16-byte global stores per thread.
//
// kPair: a cluster of two CTAs (one SM pair) computes a 256 x BN tile with tcgen05.mma.cta_group::2: each CTA stages its own
// 128 activation rows and HALF of the weight rows, the leader CTA issues the MMAs for both, and each CTA's TMEM receives
// the accumulators of its own 128 rows - the operand bytes an SM has to ingest per MAC drop by a third.
//
// Warp roles: 0..3 = epilogue (one per TMEM lane quarter; warpgroup 0, which takes the registers warpgroup 1 gives up
// through setmaxnreg), 4 = TMA producer, 5 = MMA issuer + TMEM owner, 6 = residual prefetcher, 7 = idle.
//
// Replaces the cuDNN conv + separate BN + ReLU kernels the reference reaches through
// mmcv ConvModule / mmseg Bottleneck (prediction/mmseg/models/backbones/resnet.py:267-307) and
// detectron2's BottleneckBlock.
#pragma once
#include "ptx.cuh"

namespace pn {

struct ConvParams {
  int M;           // number of output pixels = B * Ho * Wo (GEMM rows)
  int Ho, Wo;      // output spatial size
  int R, S;        // filter taps
  int stride, dil, pad;  // vertical stride / padding (and horizontal unless overridden below)
  int stride_w, pad_w;
  int kb_per_tap;  // Cin_pad / BLOCK_K
  int block_k;     // elements per K block (sw / sizeof(T))
  int sw;          // bytes per smem row == TMA swizzle span: 32, 64 or 128
  int stages;      // smem pipeline depth
  int m_tiles, n_tiles;
  int cout_store;  // channels actually written per pixel (multiple of 8)
  int a_tiled;     // 1: A is a plain [M, K] matrix fetched with tiled 2-D TMA
  const float* bias;   // [n_tiles * BN]
  // 1: the bias rides through the tensor core - the packed weights carry one extra K block per output channel holding
  // (bias_hi, bias_lo, 0, ...), multiplied by a constant tile of ones; the epilogue then adds nothing.  Fast epilogue,
  // no split-K only.
  int bias_block;
  const void* residual;  // optional, same dtype as the activations, row stride ldr
  long long ldr;
  void* out;
  long long ldc;  // output row stride in elements
  int relu;
  int out_fp32;  // write fp32 regardless of the activation dtype
  int round_tf32;  // fp32 activations: round stored values to tf32 (nearest) so the next MMA's truncation is exact
  int epi_tma;   // 1: smem-staged epilogue with TMA residual loads / TMA stores
  int cb;        // epilogue chunk row bytes (= swizzle span of the out/residual maps): 32, 64 or 128
  int res_bufs;  // residual staging depth (0 when no residual)
  int out_bufs;  // output staging depth: 2, or 1 for K-heavy layers where a deeper A/B pipeline matters more
  const int* m_limit;  // optional device-side count of valid row groups (ROIs); rows = *m_limit * m_limit_rows
  int m_limit_rows;
  // split-K: every output tile is computed by a thread-block cluster of `splits` CTAs over disjoint K-block ranges
  // (one tile per cluster, grid = tiles * splits).  Each CTA parks its fp32 partial tile in its own shared memory
  // (the operand ring, idle by then), and after a cluster barrier CTA r sums column slice r of all partial tiles
  // over distributed shared memory in a fixed order (bit-reproducible) and applies the epilogue to that slice.
  int dbg_skip;      // tuning aid (PN_CONV_TIMELINE builds only): epilogue parts to skip, see PN_SKIP
  long long* dbg;    // optional per-tile timeline of CTA 0 (tuning aid; nullptr in production): 16 clock64 slots per tile
  int splits;        // >= 1; BN / splits is a multiple of 8
  int kb_per_split;  // K blocks per split (the last split may be shorter)
  // 1: weights resident.  The grid is a multiple of n_tiles, so a CTA works on ONE n tile for its whole life: it fetches that
  // tile's weights (all K blocks) once into a dedicated region and the operand ring carries activations only.  For layers
  // whose whole weight slab fits beside the ring (stem, res2, the small-K conv3 layers): every tile otherwise re-fetches the
  // same weights from L2 - a third to a half of the bytes an SM ingests there.  Single-CTA tiles without split-K only.
  int b_resident;
};

constexpr int kBlockM = 128;
constexpr int kNumThreads = 256;  // warpgroup 0 = epilogue warps, warpgroup 1 = producer / MMA / residual (+1 idle)
constexpr int kProducerWarp = 4, kMmaWarp = 5, kResidualWarp = 6;

template <typename T>
struct ElemTraits;
template <>
struct ElemTraits<__nv_bfloat16> {
  static constexpr uint32_t kFormat = 1;
};
template <>
struct ElemTraits<float> {
  static constexpr uint32_t kFormat = 2;
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// {lo, hi} -> bf16x2 with ReLU folded into the conversion
__device__ __forceinline__ uint32_t pack_bf16_relu(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// Round to the 10-bit tf32 mantissa, nearest with ties away from zero (what cvt.rna.tf32.f32 does for finite values),
// in two full-rate integer instructions instead of a quarter-rate conversion.
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
// Byte offset of 16-byte unit j of row r inside a TMA-swizzled tile whose rows are cb bytes.
__device__ __forceinline__ uint32_t swz_off(uint32_t r, uint32_t j, uint32_t cb) {
  uint32_t off = r * cb + (j << 4);
  return off ^ (((off >> 7) & ((cb >> 4) - 1)) << 4);
}

// Direct epilogue tail for 8 consecutive output channels of row m: y already holds acc + bias.
template <typename T>
__device__ __forceinline__ void epilogue_store8(const ConvParams& p, long long m, int n0, float* y) {
  if (p.residual != nullptr) {
    const T* rp = reinterpret_cast<const T*>(p.residual) + m * p.ldr + n0;
    if constexpr (sizeof(T) == 2) {
      const uint4 rv = *reinterpret_cast<const uint4*>(rp);
      const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
        y[2 * e] += f2.x;
        y[2 * e + 1] += f2.y;
      }
    } else {
      const float4 r0 = *reinterpret_cast<const float4*>(rp);
      const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
      y[0] += r0.x, y[1] += r0.y, y[2] += r0.z, y[3] += r0.w;
      y[4] += r1.x, y[5] += r1.y, y[6] += r1.z, y[7] += r1.w;
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
  }
  if (sizeof(T) == 2 && !p.out_fp32) {
    uint4 o;
    o.x = pack_bf16(y[0], y[1]);
    o.y = pack_bf16(y[2], y[3]);
    o.z = pack_bf16(y[4], y[5]);
    o.w = pack_bf16(y[6], y[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldc + n0) = o;
  } else {
    if (p.round_tf32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = round_tf32(y[j]);
    }
    float* op = reinterpret_cast<float*>(p.out) + m * p.ldc + n0;
    *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(y[4], y[5], y[6], y[7]);
  }
}

// Residual of 8 consecutive channels of row m as raw 16-byte registers (bf16: r0 only; fp32: r0, r1).
template <typename T>
__device__ __forceinline__ void load_residual8(const ConvParams& p, long long m, int n0, uint4& r0, uint4& r1) {
  const T* rp = reinterpret_cast<const T*>(p.residual) + m * p.ldr + n0;
  r0 = __ldg(reinterpret_cast<const uint4*>(rp));
  if constexpr (sizeof(T) == 4) r1 = __ldg(reinterpret_cast<const uint4*>(rp) + 1);
}
// y (acc + bias) + residual registers -> ReLU -> store, 8 channels of row m.
template <typename T>
__device__ __forceinline__ void finish_store8(const ConvParams& p, long long m, int n0, float* y, bool has_res, const uint4& r0,
                                              const uint4& r1) {
  if (has_res) {
    if constexpr (sizeof(T) == 2) {
      const uint32_t w[4] = {r0.x, r0.y, r0.z, r0.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        y[2 * e] += __uint_as_float(w[e] << 16);
        y[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
      }
    } else {
      y[0] += __uint_as_float(r0.x), y[1] += __uint_as_float(r0.y), y[2] += __uint_as_float(r0.z), y[3] += __uint_as_float(r0.w);
      y[4] += __uint_as_float(r1.x), y[5] += __uint_as_float(r1.y), y[6] += __uint_as_float(r1.z), y[7] += __uint_as_float(r1.w);
    }
  }
  if (p.relu) {
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] = fmaxf(y[j], 0.f);
  }
  if (sizeof(T) == 2 && !p.out_fp32) {
    uint4 o;
    o.x = pack_bf16(y[0], y[1]), o.y = pack_bf16(y[2], y[3]), o.z = pack_bf16(y[4], y[5]), o.w = pack_bf16(y[6], y[7]);
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + m * p.ldc + n0) = o;
  } else {
    if (p.round_tf32) {
#pragma unroll
      for (int j = 0; j < 8; ++j) y[j] = round_tf32(y[j]);
    }
    float* op = reinterpret_cast<float*>(p.out) + m * p.ldc + n0;
    *reinterpret_cast<float4*>(op) = make_float4(y[0], y[1], y[2], y[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(y[4], y[5], y[6], y[7]);
  }
}

// Split-K tail of CTA `rank`: sum column slice `rank` of every CTA's parked partial tile over distributed shared memory
// (fixed order: split 0, 1, ...), add bias / residual, ReLU, store.  8 columns per step with the distributed-smem loads
// of all splits, the bias and the NEXT step's residual in flight together.
// peer_stride / peer_first: cluster rank of split s0's CTA = peer_first + s0 * peer_stride (1, 0 without CTA pairs; 2, half
// with pairs: the partial tiles of the same 128-row half).
template <typename T, int BN, int kSplits>
__device__ __forceinline__ void splitk_reduce(const ConvParams& p, uint32_t part0, int rank, long long m, int row, int n_tile,
                                              int peer_first, int peer_stride) {
  constexpr int kCols = BN / kSplits;  // columns this rank finishes (multiple of 8)
  uint32_t peer[kSplits];
#pragma unroll
  for (int s0 = 0; s0 < kSplits; ++s0) peer[s0] = map_shared_rank(part0, peer_first + s0 * peer_stride);
  const bool has_res = p.residual != nullptr;
  const int c_begin = rank * kCols;
  uint4 r0 = make_uint4(0u, 0u, 0u, 0u), r1 = r0;
  if (has_res && n_tile * BN + c_begin < p.cout_store) load_residual8<T>(p, m, n_tile * BN + c_begin, r0, r1);
#pragma unroll 1
  for (int c = c_begin; c < c_begin + kCols; c += 8) {
    const int n0 = n_tile * BN + c;
    if (n0 >= p.cout_store) break;
    const uint32_t off = static_cast<uint32_t>(((c >> 2) * kBlockM + row) * 16);
    float4 lo[kSplits], hi[kSplits];
#pragma unroll
    for (int s0 = 0; s0 < kSplits; ++s0) {
      lo[s0] = ld_dsmem_v4(peer[s0] + off);
      hi[s0] = ld_dsmem_v4(peer[s0] + off + kBlockM * 16);
    }
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n0) + 1);
    uint4 q0 = make_uint4(0u, 0u, 0u, 0u), q1 = q0;  // residual of the next step
    const bool more = (c + 8 < c_begin + kCols) && (n0 + 8 < p.cout_store);
    if (has_res && more) load_residual8<T>(p, m, n0 + 8, q0, q1);
    float y[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s0 = 0; s0 < kSplits; ++s0) {
      y[0] += lo[s0].x, y[1] += lo[s0].y, y[2] += lo[s0].z, y[3] += lo[s0].w;
      y[4] += hi[s0].x, y[5] += hi[s0].y, y[6] += hi[s0].z, y[7] += hi[s0].w;
    }
    y[0] += b0.x, y[1] += b0.y, y[2] += b0.z, y[3] += b0.w;
    y[4] += b1.x, y[5] += b1.y, y[6] += b1.z, y[7] += b1.w;
    finish_store8<T>(p, m, n0, y, has_res, r0, r1);
    r0 = q0, r1 = q1;
  }
}

// "accumulator stage drained": local barrier, or - in a CTA pair - the leader's barrier (it gates the leader's MMA issue)
template <bool kPair>
__device__ __forceinline__ void arrive_tempty(uint64_t* bar, uint32_t cta_rank, uint32_t pair_leader) {
  if constexpr (kPair) {
    if (cta_rank != 0) {
      mbar_arrive_remote(leader_addr(bar, pair_leader));
      return;
    }
  }
  mbar_arrive(bar);
}

// Where one work item (tile [, split]) starts: output-pixel base of the 128-row tile, first K block and filter tap.
// ~8 integer divisions by run-time values (~0.3 us of dependent latency): the first item's are computed before
// griddepcontrol.wait, off the critical path.
struct TileCoord {
  int m0, w0, h0, img, n_tile, kb_begin, kb_end, kb, r, sx;
};
template <bool kPair>
__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int work, uint32_t cta_rank, int kblocks) {
  TileCoord c;
  const int tile = work / p.splits;
  const int split = work - tile * p.splits;
  const int m_tile = tile / p.n_tiles;
  c.n_tile = tile - m_tile * p.n_tiles;
  c.m0 = (kPair ? m_tile * 2 + static_cast<int>(cta_rank) : m_tile) * kBlockM;
  const int q = c.m0 % p.Wo;
  const int t = c.m0 / p.Wo;
  const int pp = t % p.Ho;
  c.img = t / p.Ho;
  c.w0 = q * p.stride_w - p.pad_w;
  c.h0 = pp * p.stride - p.pad;
  c.kb_begin = split * p.kb_per_split;
  c.kb_end = min(kblocks, c.kb_begin + p.kb_per_split);
  const int tap = c.kb_begin / p.kb_per_tap;
  c.kb = c.kb_begin - tap * p.kb_per_tap;
  c.r = tap / p.S;
  c.sx = tap - c.r * p.S;
  return c;
}

// Tuning aid (build with -DPN_CONV_TIMELINE, run tools/conv_one.py with PN_CONV_DBG=1): clock64 timeline of CTA 0.
#ifdef PN_CONV_TIMELINE
#define PN_DBG(iter, slot) do { if (p.dbg != nullptr && blockIdx.x == 0 && (iter) < 64) p.dbg[(iter) * 16 + (slot)] = clock64(); } while (0)
__device__ __forceinline__ long long pn_globaltimer() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// launch log of CTA 0 (thread 0): dbg[1024] counts launches, dbg[1025 + 4*i + {0,1,2}] = globaltimer at entry / after
// griddepcontrol.wait / at exit of launch i
#define PN_LOG(which) do { if (p.dbg != nullptr && blockIdx.x == 0 && threadIdx.x == 0) { \
    const long long n_ = p.dbg[1024]; if (n_ < 64) p.dbg[1025 + 4 * n_ + (which)] = pn_globaltimer(); \
    if ((which) == 2) p.dbg[1024] = n_ + 1; } } while (0)
// epilogue attribution by elimination (results are wrong, timings tell): 1 TMA store + its wait, 2 st.shared, 4 residual
// (wait + loads + adds), 8 bias loads, 16 tcgen05.ld double buffering (wait right after issue)
#define PN_SKIP(bit) ((p.dbg_skip & (bit)) != 0)
#else
#define PN_SKIP(bit) false
#define PN_DBG(iter, slot) do { } while (0)
#define PN_LOG(which) do { } while (0)
#endif

template <typename T, int BN, bool kSplit, bool kPair>
__global__ void __launch_bounds__(kNumThreads, 2)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
                 const ConvParams p) {
  // Programmatic dependent launch: let the next kernel in the stream start its prologue (and become resident next
  // to this CTA when both fit) right away; it blocks in griddep_wait() until this grid has completed.
  griddep_launch_dependents();
  PN_LOG(0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kBLoad = kPair ? BN / 2 : BN;  // weight rows this CTA stages per K block
  // pair: rank within the SM pair; with split-K the cluster holds `splits` pairs (cluster ranks 2s, 2s + 1)
  const uint32_t cluster_rank = (kPair || kSplit) ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = kPair ? (cluster_rank & 1u) : 0u;
  const uint32_t pair_leader = cluster_rank & ~1u;
  const uint32_t pair_mask = 3u << pair_leader;
  const uint32_t a_bytes = kBlockM * p.sw;
  const uint32_t b_bytes = kBLoad * p.sw;
  const bool bres = !kSplit && !kPair && p.b_resident != 0;
  const int kblocks_all = p.R * p.S * p.kb_per_tap;
  const uint32_t stage_tx = bres ? a_bytes : (kPair ? 2u : 1u) * (a_bytes + b_bytes);  // bytes that complete one (leader) full barrier
  // persistent work items are walked by CTA (or by CTA pair)
  const int work_first = kPair ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int work_stride = kPair ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const uint32_t chunk_bytes = p.epi_tma ? kBlockM * p.cb : 0;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + p.stages * a_bytes;
  uint8_t* smem_out = smem_b + (bres ? kblocks_all : p.stages) * b_bytes;
  uint8_t* smem_res = smem_out + p.out_bufs * chunk_bytes;
  float* smem_scale = reinterpret_cast<float*>(smem_res + p.res_bufs * chunk_bytes);
  uint8_t* smem_ones = reinterpret_cast<uint8_t*>(smem_scale + 4 * BN);  // [128 rows][32 B], 32-byte swizzle: ones at k = 0, 1
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_ones + kBlockM * 32);  // (smem_scale: one bias[BN] copy per epilogue warp)
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tfull_bar = bars + 2 * p.stages;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* rfull_bar = tempty_bar + 2;
  uint64_t* rempty_bar = rfull_bar + 4;
  uint64_t* bres_bar = rempty_bar + 4;   // resident weights have landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bres_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.epi_tma) tma_prefetch_desc(&tmap_out);
    if (p.epi_tma && p.residual) tma_prefetch_desc(&tmap_res);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], kPair ? 8 : 4);  // pair: the leader's barrier collects both CTAs' epilogue warps
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&rfull_bar[i], 1);
      mbar_init(&rempty_bar[i], 4);
    }
    mbar_init(bres_bar, 1);
    fence_barrier_init();
  }
  if (p.bias_block && threadIdx.x < kBlockM) {
    // A operand of the bias block: row r = (1, 1, 0, ...) in K; 32-byte rows, 32-byte swizzle (16-byte unit j of row r
    // sits at unit j ^ ((r >> 2) & 1))
    const uint32_t r = threadIdx.x;
    uint4 one = make_uint4(0u, 0u, 0u, 0u), zero = one;
    if constexpr (sizeof(T) == 2) one.x = 0x3f803f80u;            // two bf16 ones
    else one.x = 0x3f800000u, one.y = 0x3f800000u;                  // two fp32 (tf32) ones
    const uint32_t u = (r >> 2) & 1u;
    *reinterpret_cast<uint4*>(smem_ones + r * 32 + (u << 4)) = one;
    *reinterpret_cast<uint4*>(smem_ones + r * 32 + ((u ^ 1u) << 4)) = zero;
    fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's (async proxy) reads
  }
  __syncthreads();  // barriers initialised (and the ones tile written) before anybody uses them
  if constexpr (kPair) {  // the peer's barriers must be initialised before anything signals them
    cluster_arrive_release();
    cluster_wait_acquire();
  }
  // Weights do not depend on the previous kernel: arm the first pipeline stages of this CTA's first work item and
  // fetch their weight tiles while the previous kernel drains (its tail would otherwise hide nothing but the prologue).
  int pre_armed = 0;
  if (warp == kProducerWarp && p.m_limit == nullptr && !bres && work_first < p.m_tiles * p.n_tiles * p.splits) {
    const int work0 = work_first;
    const int tile0 = work0 / p.splits;
    const int split0 = work0 - tile0 * p.splits;
    const int n_tile0 = tile0 % p.n_tiles;
    const int kb_total = p.R * p.S * p.kb_per_tap;
    const int kb_begin0 = split0 * p.kb_per_split;
    const int my_kb = min(kb_total, kb_begin0 + p.kb_per_split) - kb_begin0;
    pre_armed = min(p.stages, my_kb);
    if (elect_one()) {
      for (int i = 0; i < pre_armed; ++i) {
        if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[i], stage_tx);
        if constexpr (kPair) {
          tma_load_2d_pair(&tmap_b, leader_addr(&full_bar[i], pair_leader), smem_b + i * b_bytes, (kb_begin0 + i) * p.block_k,
                           n_tile0 * BN + static_cast<int>(cta_rank) * kBLoad);
        } else {
          tma_load_2d(&tmap_b, &full_bar[i], smem_b + i * b_bytes, (kb_begin0 + i) * p.block_k, n_tile0 * BN);
        }
      }
    }
  }
  TileCoord first_tc{};
  if (warp == kProducerWarp) first_tc = tile_coord<kPair>(p, work_first, cta_rank, p.R * p.S * p.kb_per_tap);
  // everything above (barrier init, descriptor prefetch, weight prefetch, first tile's index math)
  // overlapped the previous kernel's tail; from here on we read what it wrote
  griddep_wait();
  PN_LOG(1);
  // Tensor memory is allocated only NOW.  A CTA that is resident early (programmatic dependent launch) and still waiting
  // for its predecessor grid must not hold tensor-memory columns: with several conv CTAs per SM (two CTAs per SM, parallel
  // lanes, a second stream) a running grid's CTA can then block in tcgen05.alloc behind columns held by a CTA that is
  // itself waiting for a grid that needs THAT CTA to finish - a cross-stream cycle observed as a hang in 5 of 12 runs of
  // the 8-environment pipeline, 0 of 12 with the allocation here (profiles/r02_stall_analysis.txt).  The TMA producer does
  // not need the address and starts loading at once; the allocation overlaps its first-byte latency.
  uint32_t tmem_base = 0;
  if (warp == kMmaWarp) {
    if constexpr (kPair) {
      tmem_alloc_pair(tmem_ptr, kTmemCols);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_ptr, kTmemCols);
      tmem_relinquish();
    }
  }
  if constexpr (kPair) {
    // the leader's MMAs write the peer's tensor memory too: both allocations must exist before the first MMA
    tc_fence_before();
    cluster_arrive_release();
    cluster_wait_acquire();
    tc_fence_after();
    tmem_base = *tmem_ptr;
  } else if (warp < 4 || warp == kMmaWarp) {  // epilogue warps + the allocating MMA warp; the other warps never touch it
    tc_fence_before();
    named_bar_sync(1, 160);
    tc_fence_after();
    tmem_base = *tmem_ptr;
  }

  int m_tiles_live = p.m_tiles;
  if (p.m_limit != nullptr) {
    const long long rows = static_cast<long long>(__ldg(p.m_limit)) * p.m_limit_rows;
    constexpr int kRowsPerItem = kPair ? 2 * kBlockM : kBlockM;
    const int t = static_cast<int>((rows + kRowsPerItem - 1) / kRowsPerItem);
    m_tiles_live = t < p.m_tiles ? t : p.m_tiles;
  }
  const int num_tiles = m_tiles_live * p.n_tiles * p.splits;  // work items: (m_tile, n_tile, split), split fastest
  const int taps = p.R * p.S;
  const int kblocks = taps * p.kb_per_tap;
  const int cols_per_chunk = p.epi_tma ? p.cb / static_cast<int>(sizeof(T)) : 32;

  if (warp >= 4) {
  setmaxnreg_dec<64>();  // data-movement / issue warps need few registers: the epilogue warpgroup takes the rest
  if (warp == kProducerWarp) {
    // ------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      if (bres && work_first < num_tiles) {   // this CTA's n tile never changes: its weights, all K blocks, once
        const int n_tile0 = work_first % p.n_tiles;
        mbar_arrive_expect_tx(bres_bar, static_cast<uint32_t>(kblocks) * b_bytes);
        for (int kb = 0; kb < kblocks; ++kb) tma_load_2d(&tmap_b, bres_bar, smem_b + kb * b_bytes, kb * p.block_k, n_tile0 * BN);
      }
      for (int work = work_first; work < num_tiles; work += work_stride) {
        const TileCoord tc = (work == work_first) ? first_tc : tile_coord<kPair>(p, work, cta_rank, kblocks);
        const int n_tile = tc.n_tile, m0 = tc.m0, w0 = tc.w0, h0 = tc.h0, img = tc.img;
        const int kb_begin = tc.kb_begin, kb_end = tc.kb_end;
        int kb = tc.kb, r = tc.r, sx = tc.sx;
        PN_DBG((work - work_first) / work_stride, 0);
        for (int kb_global = kb_begin; kb_global < kb_end; ++kb_global) {
          const bool armed = pre_armed > 0;  // stage already armed and its weight tile already in flight
          if (armed) {
            --pre_armed;
          } else {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
          }
          if constexpr (kPair) {  // both CTAs' bytes complete the LEADER's full barrier
            const uint32_t fb = leader_addr(&full_bar[stage], pair_leader);
            if (p.a_tiled) {
              tma_load_2d_pair(&tmap_a, fb, smem_a + stage * a_bytes, kb * p.block_k, m0);
            } else {
              tma_load_im2col_4d_pair(&tmap_a, fb, smem_a + stage * a_bytes, kb * p.block_k, w0, h0, img,
                                      static_cast<uint16_t>(sx * p.dil), static_cast<uint16_t>(r * p.dil));
            }
            if (!armed)
              tma_load_2d_pair(&tmap_b, fb, smem_b + stage * b_bytes, kb_global * p.block_k,
                               n_tile * BN + static_cast<int>(cta_rank) * kBLoad);
          } else {
            if (p.a_tiled) {
              tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * a_bytes, kb * p.block_k, m0);
            } else {
              tma_load_im2col_4d(&tmap_a, &full_bar[stage], smem_a + stage * a_bytes, kb * p.block_k, w0, h0, img,
                                 static_cast<uint16_t>(sx * p.dil), static_cast<uint16_t>(r * p.dil));
            }
            if (!armed && !bres) tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * b_bytes, kb_global * p.block_k, n_tile * BN);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
          if (++kb == p.kb_per_tap) {
            kb = 0;
            if (++sx == p.S) sx = 0, ++r;
          }
        }
        if (p.bias_block) {  // the bias block: weights' extra K block only (its A operand is the constant ones tile)
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], (kPair ? 2u : 1u) * b_bytes);
          if constexpr (kPair) {
            tma_load_2d_pair(&tmap_b, leader_addr(&full_bar[stage], pair_leader), smem_b + stage * b_bytes, kblocks * p.block_k,
                             n_tile * BN + static_cast<int>(cta_rank) * kBLoad);
          } else {
            tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * b_bytes, kblocks * p.block_k, n_tile * BN);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------ MMA issuer (one elected lane)
    if (cta_rank == 0 && elect_one()) {  // pair: the leader issues for both CTAs
      constexpr uint32_t idesc = umma_idesc(ElemTraits<T>::kFormat, BN, kPair ? 256u : 128u);
      const int ksteps = p.sw / 32;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      if (bres && work_first < num_tiles) {
        mbar_wait(bres_bar, 0);
        tc_fence_after();
      }
      for (int work = work_first; work < num_tiles; work += work_stride, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        PN_DBG(it, 1);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        PN_DBG(it, 2);
        const uint32_t d_tmem = tmem_base + acc * BN;
        const int split = work % p.splits;
        const int my_kblocks = min(kblocks, (split + 1) * p.kb_per_split) - split * p.kb_per_split;
        for (int kb = 0; kb < my_kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc(smem_u32(smem_a + stage * a_bytes), p.sw);
          const uint64_t bdesc = umma_smem_desc(smem_u32(smem_b + (bres ? kb : stage) * b_bytes), p.sw);
          for (int k = 0; k < ksteps; ++k) {
            const uint32_t accum = (kb | k) ? 1u : 0u;
            if constexpr (kPair) {
              if constexpr (ElemTraits<T>::kFormat == 1) umma_bf16_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
              else umma_tf32_pair(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
            } else {
              if constexpr (ElemTraits<T>::kFormat == 1) umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
              else umma_tf32(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, accum);
            }
          }
          if constexpr (kPair) umma_commit_pair(&empty_bar[stage], pair_mask);  // frees the stage in both CTAs
          else umma_commit(&empty_bar[stage]);
          if (kb == 0) PN_DBG(it, 3);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (p.bias_block) {  // D += ones[128 x 2] * (bias_hi, bias_lo)[BN x 2]^T: one K step of the extra block
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_smem_desc(smem_u32(smem_ones), 32);
          const uint64_t bdesc = umma_smem_desc(smem_u32(smem_b + stage * b_bytes), p.sw);
          if constexpr (kPair) {
            if constexpr (ElemTraits<T>::kFormat == 1) umma_bf16_pair(d_tmem, adesc, bdesc, idesc, 1u);
            else umma_tf32_pair(d_tmem, adesc, bdesc, idesc, 1u);
            umma_commit_pair(&empty_bar[stage], pair_mask);
          } else {
            if constexpr (ElemTraits<T>::kFormat == 1) umma_bf16(d_tmem, adesc, bdesc, idesc, 1u);
            else umma_tf32(d_tmem, adesc, bdesc, idesc, 1u);
            umma_commit(&empty_bar[stage]);
          }
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if constexpr (kPair) umma_commit_pair(&tfull_bar[acc], pair_mask);  // every MMA of this tile has retired
        else umma_commit(&tfull_bar[acc]);
        PN_DBG(it, 4);
      }
    }
  } else if (warp == kResidualWarp) {
    // ------------------------------------------------------------ residual prefetcher (TMA epilogue only)
    if (p.epi_tma && p.residual != nullptr && elect_one()) {
      uint32_t g = 0;  // running chunk counter, same sequence as the epilogue warps
      for (int tile = work_first; tile < num_tiles; tile += work_stride) {
        const int m_tile = tile / p.n_tiles;
        const int n_tile = tile - m_tile * p.n_tiles;
        for (int n0 = n_tile * BN; n0 < n_tile * BN + BN && n0 < p.cout_store; n0 += cols_per_chunk, ++g) {
          const uint32_t rb = g % p.res_bufs;
          const uint32_t ph = (g / p.res_bufs) & 1;
          mbar_wait(&rempty_bar[rb], ph ^ 1);
          mbar_arrive_expect_tx(&rfull_bar[rb], chunk_bytes);
          tma_load_2d(&tmap_res, &rfull_bar[rb], smem_res + rb * chunk_bytes, n0,
                      (kPair ? m_tile * 2 + static_cast<int>(cta_rank) : m_tile) * kBlockM);
        }
      }
    }
  }
    if constexpr (kSplit) {  // keep pace with the epilogue warps' two cluster barriers (see below)
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
    }
  } else {
  setmaxnreg_inc<192>();
  if (p.epi_tma) {
    // ------------------------------------------------------------ epilogue, fast path
    // Every warp is self-contained: it owns 32 accumulator rows (its TMEM lane quarter), a private copy of the tile's
    // scale/bias, its 32-row slice of the output staging buffers and its own TMA stores (32-row boxes), so the four
    // warps never meet at a barrier.  TMEM loads are double-buffered in registers (32 columns in flight while the
    // previous 32 are processed).
    const int quarter = warp & 3;  // TMEM lane quarter this warp may touch
    const uint32_t row = quarter * 32 + lane;
    constexpr int kUnits = 32 * sizeof(T) / 16;  // 16-byte units per 32 columns
    const int subs_per_chunk = cols_per_chunk / 32;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t sb_addr = smem_u32(smem_scale) + quarter * (BN * 4);  // bias[BN] of this warp
    const uint32_t out_addr = smem_u32(smem_out);
    const uint32_t res_addr = smem_u32(smem_res);
    const bool has_res = p.residual != nullptr;
    // 16-byte unit j of this thread's row inside a TMA-swizzled chunk buffer: row_off + ((j ^ row_xor) << 4)
    const uint32_t row_off = row * p.cb;
    const uint32_t row_xor = (p.cb == 128) ? (row & 7u) : (p.cb == 64 ? ((row >> 1) & 3u) : 0u);
    uint32_t g = 0;  // running chunk counter (same sequence as the residual prefetcher)
    uint32_t rb = 0, rphase = 0;  // residual ring slot / phase of chunk g
    int it = 0;
    int cached_n_tile = -1;
    for (int tile = work_first; tile < num_tiles; tile += work_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      if (n_tile != cached_n_tile && !p.bias_block) {
        __syncwarp();
        for (int i = lane * 4; i < BN; i += 128) {
          const float4 bi = __ldg(reinterpret_cast<const float4*>(p.bias + n_tile * BN + i));
          sts_v4(sb_addr + i * 4, __float_as_uint(bi.x), __float_as_uint(bi.y), __float_as_uint(bi.z), __float_as_uint(bi.w));
        }
        cached_n_tile = n_tile;
        __syncwarp();
      }
      const int cols_live = min(BN, p.cout_store - n_tile * BN);
      const int nchunks = (cols_live + cols_per_chunk - 1) / cols_per_chunk;
      const int ngroups = nchunks * subs_per_chunk;
      const uint32_t acc_taddr = lane_taddr + acc * BN;
      if (warp == 0 && lane == 0) PN_DBG(it, 5);
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if (warp == 0 && lane == 0) PN_DBG(it, 6);

      // one 32-column group: v holds the accumulators of columns [grp*32, grp*32+32) of this thread's row
      auto process = [&](uint32_t (&v)[32], int grp) {
        const int sub = (subs_per_chunk == 2) ? (grp & 1) : 0;
        const uint32_t buf = (p.out_bufs == 2) ? (g & 1u) : 0u;
        const uint32_t obuf = out_addr + buf * chunk_bytes;
        if (sub == 0) {
          if (lane == 0 && !PN_SKIP(1)) {  // this warp's earlier store out of this buffer has been read
            if (p.out_bufs == 2) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          __syncwarp();
          if (has_res && !PN_SKIP(4)) mbar_wait(&rfull_bar[rb], rphase);
        }
        // All shared-memory loads of the group are issued back to back (the asm statements keep program order, so
        // interleaving them with the stores would serialise one load-compute-store chain per 16-byte unit).
        const uint32_t sb = sb_addr + grp * 32 * 4;
        float4 bi[8];
        if (!p.bias_block && !PN_SKIP(8)) {
#pragma unroll
          for (int j = 0; j < 8; ++j) bi[j] = lds_f4(sb + j * 16);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) bi[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint4 rv[kUnits];
        if (has_res && !PN_SKIP(4)) {
          const uint32_t rbuf = res_addr + rb * chunk_bytes + row_off;
#pragma unroll
          for (int u = 0; u < kUnits; ++u) rv[u] = lds_v4(rbuf + ((static_cast<uint32_t>(sub * kUnits + u) ^ row_xor) << 4));
        }
        float y[32];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          y[4 * j] = __uint_as_float(v[4 * j]) + bi[j].x;
          y[4 * j + 1] = __uint_as_float(v[4 * j + 1]) + bi[j].y;
          y[4 * j + 2] = __uint_as_float(v[4 * j + 2]) + bi[j].z;
          y[4 * j + 3] = __uint_as_float(v[4 * j + 3]) + bi[j].w;
        }
        if (has_res && !PN_SKIP(4)) {
#pragma unroll
          for (int u = 0; u < kUnits; ++u) {
            if constexpr (sizeof(T) == 2) {
              const uint32_t w[4] = {rv[u].x, rv[u].y, rv[u].z, rv[u].w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                y[u * 8 + 2 * e] += __uint_as_float(w[e] << 16);
                y[u * 8 + 2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
              }
            } else {
              y[u * 4] += __uint_as_float(rv[u].x);
              y[u * 4 + 1] += __uint_as_float(rv[u].y);
              y[u * 4 + 2] += __uint_as_float(rv[u].z);
              y[u * 4 + 3] += __uint_as_float(rv[u].w);
            }
          }
        }
        uint4 ov[kUnits];
#pragma unroll
        for (int u = 0; u < kUnits; ++u) {
          if constexpr (sizeof(T) == 2) {
            if (p.relu) {
              ov[u].x = pack_bf16_relu(y[u * 8], y[u * 8 + 1]), ov[u].y = pack_bf16_relu(y[u * 8 + 2], y[u * 8 + 3]);
              ov[u].z = pack_bf16_relu(y[u * 8 + 4], y[u * 8 + 5]), ov[u].w = pack_bf16_relu(y[u * 8 + 6], y[u * 8 + 7]);
            } else {
              ov[u].x = pack_bf16(y[u * 8], y[u * 8 + 1]), ov[u].y = pack_bf16(y[u * 8 + 2], y[u * 8 + 3]);
              ov[u].z = pack_bf16(y[u * 8 + 4], y[u * 8 + 5]), ov[u].w = pack_bf16(y[u * 8 + 6], y[u * 8 + 7]);
            }
          } else {
            float a = y[u * 4], b = y[u * 4 + 1], c = y[u * 4 + 2], d = y[u * 4 + 3];
            if (p.relu) a = fmaxf(a, 0.f), b = fmaxf(b, 0.f), c = fmaxf(c, 0.f), d = fmaxf(d, 0.f);
            if (p.round_tf32) a = round_tf32(a), b = round_tf32(b), c = round_tf32(c), d = round_tf32(d);
            ov[u].x = __float_as_uint(a), ov[u].y = __float_as_uint(b), ov[u].z = __float_as_uint(c), ov[u].w = __float_as_uint(d);
          }
        }
        if (!PN_SKIP(2)) {
          const uint32_t ob = obuf + row_off;
#pragma unroll
          for (int u = 0; u < kUnits; ++u)
            sts_v4(ob + ((static_cast<uint32_t>(sub * kUnits + u) ^ row_xor) << 4), ov[u].x, ov[u].y, ov[u].z, ov[u].w);
        }
        if (sub == subs_per_chunk - 1) {
          fence_proxy_async();  // make the st.shared above visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) {
            if (has_res) mbar_arrive(&rempty_bar[rb]);
            const int n0 = n_tile * BN + (grp - sub) * 32;
            if (!PN_SKIP(1)) {
              tma_store_2d_addr(&tmap_out, obuf + quarter * 32 * p.cb, n0,
                                (kPair ? m_tile * 2 + static_cast<int>(cta_rank) : m_tile) * kBlockM + quarter * 32);
              tma_store_commit();
            }
          }
          ++g;
          if (has_res && ++rb == static_cast<uint32_t>(p.res_bufs)) rb = 0, rphase ^= 1;
        }
        if (warp == 0 && lane == 0 && grp < 8) PN_DBG(it, 8 + grp);
      };

      uint32_t va[32], vb[32];
      tmem_ld_32x32(acc_taddr, va);
#pragma unroll 1
      for (int grp = 0; grp < ngroups; grp += 2) {
        tmem_ld_wait();
        const bool more_b = grp + 1 < ngroups;
        if (more_b) {
          tmem_ld_32x32(acc_taddr + (grp + 1) * 32, vb);
          if (PN_SKIP(16)) tmem_ld_wait();
        } else {  // every accumulator column this warp needs is in registers: hand the TMEM stage back early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_tempty<kPair>(&tempty_bar[acc], cta_rank, pair_leader);
        }
        process(va, grp);
        if (more_b) {
          tmem_ld_wait();
          if (grp + 2 < ngroups) {
            tmem_ld_32x32(acc_taddr + (grp + 2) * 32, va);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_tempty<kPair>(&tempty_bar[acc], cta_rank, pair_leader);
          }
          process(vb, grp + 1);
        }
      }
      if (warp == 0 && lane == 0) PN_DBG(it, 7);
    }
    // shared memory must outlive the bulk stores' reads; their global writes complete with the grid
    if (lane == 0) tma_store_wait_read<0>();
  } else {
    // ------------------------------------------------------------ epilogue, direct path: TMEM -> regs -> global
    const int quarter = warp & 3;
    int it = 0;
    for (int work = work_first; work < num_tiles; work += work_stride, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int tile = work / p.splits;
      const int m_tile = tile / p.n_tiles;
      const int n_tile = tile - m_tile * p.n_tiles;
      const long long m = static_cast<long long>(kPair ? m_tile * 2 + static_cast<int>(cta_rank) : m_tile) * kBlockM + quarter * 32 + lane;
      const bool row_ok = m < p.M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      if constexpr (kSplit) {
        // park the partial tile in the (now idle) operand ring: 16-byte unit (column group g, row) at (g*128 + row)*16
        const uint32_t part_addr = smem_u32(smem_a) + (quarter * 32 + lane) * 16;
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
        auto park = [&](const uint32_t (&v)[32], int chunk) {
#pragma unroll
          for (int g = 0; g < 8; ++g)
            sts_v4(part_addr + (chunk * 8 + g) * (kBlockM * 16), v[g * 4], v[g * 4 + 1], v[g * 4 + 2], v[g * 4 + 3]);
        };
        uint32_t va[32], vb[32];  // one 32-column load in flight while the previous one is written out
        tmem_ld_32x32(taddr, va);
#pragma unroll 1
        for (int chunk = 0; chunk < BN / 32; chunk += 2) {
          tmem_ld_wait();
          if (chunk + 1 < BN / 32) tmem_ld_32x32(taddr + (chunk + 1) * 32, vb);
          park(va, chunk);
          if (chunk + 1 < BN / 32) {
            tmem_ld_wait();
            if (chunk + 2 < BN / 32) tmem_ld_32x32(taddr + (chunk + 2) * 32, va);
            park(vb, chunk + 1);
          }
        }
        // the tile is finished after the cluster barrier below
      } else {
#pragma unroll 1
        for (int chunk = 0; chunk < BN / 32; ++chunk) {
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + chunk * 32, v);
          tmem_ld_wait();
          const int n0 = n_tile * BN + chunk * 32;
          if (row_ok && n0 < p.cout_store) {
            float y[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              y[j] = __uint_as_float(v[j]) + __ldg(p.bias + n0 + j);
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              if (n0 + g * 8 + 8 <= p.cout_store) epilogue_store8<T>(p, m, n0 + g * 8, &y[g * 8]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_tempty<kPair>(&tempty_bar[acc], cta_rank, pair_leader);
    }
  }
    if constexpr (kSplit) {
      // ------------------------------------------------------------ split-K: reduce-scatter over the cluster
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
      if (work_first < num_tiles) {
        const int tile = work_first / p.splits;
        const int rank = work_first - tile * p.splits;  // this CTA's split == the column slice it finishes
        const int m_tile = tile / p.n_tiles;
        const int n_tile = tile - m_tile * p.n_tiles;
        const int row = warp * 32 + lane;
        const long long m = static_cast<long long>(kPair ? m_tile * 2 + static_cast<int>(cta_rank) : m_tile) * kBlockM + row;
        const uint32_t part0 = smem_u32(smem_a);
        const int peer_first = kPair ? static_cast<int>(cta_rank) : 0, peer_stride = kPair ? 2 : 1;
        if (m < p.M) {
          switch (p.splits) {
            case 2: splitk_reduce<T, BN, 2>(p, part0, rank, m, row, n_tile, peer_first, peer_stride); break;
            case 4: splitk_reduce<T, BN, 4>(p, part0, rank, m, row, n_tile, peer_first, peer_stride); break;
            default:
              if constexpr (BN >= 64 && !kPair) splitk_reduce<T, BN, 8>(p, part0, rank, m, row, n_tile, peer_first, peer_stride);
              break;
          }
        }
      }
      // nobody may exit (and release its shared memory) while a peer can still be reading it
      __syncwarp();
      cluster_arrive_release();
      cluster_wait_acquire();
    }

  }

  tc_fence_before();
  __syncthreads();
  if constexpr (kPair) {  // both CTAs are done with each other's shared memory and tensor memory
    cluster_arrive_release();
    cluster_wait_acquire();
  }
  if (warp == kMmaWarp) {
    tc_fence_after();
    if constexpr (kPair) tmem_dealloc_pair(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
  PN_LOG(2);
}

}  // namespace pn
