// Result gather over NVLink peer memory: the one exchange on the path.
//
// The perception path shards over independent environments, one process per GPU; the only data that ever crosses GPUs
// is each rank's per-environment result (the predicted maps) travelling to the rank that hosts the planner.  The
// reference has no counterpart (it is single-process, nav/collect.py:32-33).  Instead of a send/recv collective this is
// a one-sided push: the root allocates a slab, every other rank maps it through CUDA IPC and one kernel per step writes
// the rank's results straight into its slice over NVLink (16-byte peer stores, all ranks concurrently - NVSwitch gives
// the root its full ingress bandwidth), then raises a per-rank sequence flag (release, system scope).  The root's step
// is one small kernel that waits for every flag (acquire, system scope).  Two slots alternate by step parity; a rank may
// only overwrite a slot after the root has begun the step after the one that last used it (`ack`), so a fast rank can
// never clobber results the root has not consumed yet.  Every wait is bounded (globaltimer) and reports through a
// host-visible status word instead of hanging the GPU.
#include <cstdio>
#include <cstring>

#include "../../include/peanut_b200.h"
#include "engine.h"

namespace pn {

namespace {

constexpr unsigned long long kWaitNs = 4000000000ull;  // 4 s

struct Control {  // lives in the root's memory, one 128-byte line per word that a different GPU polls
  unsigned long long flag[16][16];  // flag[r][0] = last step rank r has pushed completely (steps count from 1)
  unsigned long long ack[16];       // ack[0]     = the step the root has begun
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Bounded wait until *p >= want; false (and *status = code) on time-out.
__device__ bool wait_ge(const unsigned long long* p, unsigned long long want, int* status, int code) {
  const unsigned long long t0 = globaltimer_ns();
  while (ld_acquire_sys(p) < want) {
    if (globaltimer_ns() - t0 > kWaitNs) {
      *reinterpret_cast<volatile int*>(status) = code;
      return false;
    }
    __nanosleep(200);
  }
  return true;
}

// One rank's push: src (local) -> dst (its slice of the slot, local for the root, peer memory otherwise), then flag = seq.
//   ack_wait != nullptr: first wait until the root has begun step `need_ack` (slot free)
//   ack_post != nullptr: (root) first publish that step `seq` has begun
__global__ void __launch_bounds__(256) k_gather_push(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n16,
                                                     unsigned* counter, unsigned long long* flag, unsigned long long seq,
                                                     const unsigned long long* ack_wait, unsigned long long need_ack,
                                                     unsigned long long* ack_post, int* status) {
  __shared__ int ok;
  if (threadIdx.x == 0) {
    ok = 1;
    if (ack_post != nullptr && blockIdx.x == 0) st_release_sys(ack_post, seq);
    if (ack_wait != nullptr && need_ack > 0) ok = wait_ge(ack_wait, need_ack, status, 2) ? 1 : 0;
  }
  __syncthreads();
  if (ok) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    // four independent 16-byte loads in flight per thread before the peer stores
    for (; i + 3 * stride < n16; i += 4 * stride) {
      const uint4 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
      dst[i] = a, dst[i + stride] = b, dst[i + 2 * stride] = c, dst[i + 3 * stride] = d;
    }
    for (; i < n16; i += stride) dst[i] = src[i];
  }
  __threadfence_system();  // this thread's peer stores are visible system-wide before the block signs off
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u);
    if (done == gridDim.x - 1) {  // last block: every block's stores are fenced
      *counter = 0;
      __threadfence_system();
      st_release_sys(flag, seq);
    }
  }
}

// Root: wait until every rank has pushed step `seq`.
__global__ void k_gather_wait(const Control* ctl, int world, unsigned long long seq, int* status) {
  const int r = threadIdx.x;
  if (r < world) wait_ge(&ctl->flag[r][0], seq, status, 3);
}

}  // namespace

struct Gather {
  int device = 0, rank = 0, world = 1, root = 0;
  size_t bytes = 0, slice = 0;  // payload bytes per rank, slice stride (256-byte aligned)
  char* slab = nullptr;         // root: owned; others: IPC mapping of the root's allocation
  bool mapped = false;
  Control* ctl = nullptr;       // inside the slab allocation, after the two slots
  unsigned* counter = nullptr;  // local
  int* status_host = nullptr;   // pinned, mapped
  int* status_dev = nullptr;
  unsigned long long seq = 0;
  int num_sms = 148;
  size_t slot_bytes() const { return slice * world; }
};

}  // namespace pn

using namespace pn;

#define PN_G_BEGIN try {
#define PN_G_END                                 \
  }                                              \
  catch (const std::exception& e) {              \
    set_last_error(e.what());                    \
    return 1;                                    \
  }                                              \
  return 0;

extern "C" {

int pn_gather_create(pn_ctx* ctx, int rank, int world, int root, int64_t bytes_per_rank, pn_gather** out) {
  PN_G_BEGIN
  PN_REQUIRE(ctx != nullptr && out != nullptr, "pn_gather_create: null argument");
  PN_REQUIRE(world >= 1 && world <= 16 && rank >= 0 && rank < world && root >= 0 && root < world, "pn_gather_create: bad rank / world (<= 16)");
  PN_REQUIRE(bytes_per_rank > 0 && bytes_per_rank % 16 == 0, "pn_gather_create: bytes_per_rank must be a positive multiple of 16");
  int device = 0, sms = 148;
  ctx_device(ctx, &device, &sms);
  PN_CUDA_CHECK(cudaSetDevice(device));
  auto* g = new Gather();
  g->device = device, g->rank = rank, g->world = world, g->root = root, g->num_sms = sms;
  g->bytes = static_cast<size_t>(bytes_per_rank);
  g->slice = (g->bytes + 255) & ~size_t(255);
  PN_CUDA_CHECK(cudaMalloc(&g->counter, 256));
  PN_CUDA_CHECK(cudaMemset(g->counter, 0, 256));
  PN_CUDA_CHECK(cudaHostAlloc(&g->status_host, sizeof(int), cudaHostAllocMapped));
  *g->status_host = 0;
  PN_CUDA_CHECK(cudaHostGetDevicePointer(&g->status_dev, g->status_host, 0));
  if (rank == root) {
    const size_t total = 2 * g->slot_bytes() + sizeof(Control);
    PN_CUDA_CHECK(cudaMalloc(&g->slab, total));
    PN_CUDA_CHECK(cudaMemset(g->slab, 0, total));
    g->ctl = reinterpret_cast<Control*>(g->slab + 2 * g->slot_bytes());
    PN_CUDA_CHECK(cudaDeviceSynchronize());
  }
  *out = reinterpret_cast<pn_gather*>(g);
  PN_G_END
}

int pn_gather_export(pn_gather* gh, void* handle64_out) {
  PN_G_BEGIN
  auto* g = reinterpret_cast<Gather*>(gh);
  PN_REQUIRE(g && handle64_out && g->rank == g->root && g->slab, "pn_gather_export: only the root exports its slab");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  PN_CUDA_CHECK(cudaSetDevice(g->device));
  cudaIpcMemHandle_t h;
  PN_CUDA_CHECK(cudaIpcGetMemHandle(&h, g->slab));
  std::memcpy(handle64_out, &h, 64);
  PN_G_END
}

int pn_gather_connect(pn_gather* gh, const void* handle64) {
  PN_G_BEGIN
  auto* g = reinterpret_cast<Gather*>(gh);
  PN_REQUIRE(g && handle64 && g->rank != g->root && !g->slab, "pn_gather_connect: non-root ranks connect once");
  PN_CUDA_CHECK(cudaSetDevice(g->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle64, 64);
  void* p = nullptr;
  PN_CUDA_CHECK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  g->slab = static_cast<char*>(p);
  g->mapped = true;
  g->ctl = reinterpret_cast<Control*>(g->slab + 2 * g->slot_bytes());
  PN_G_END
}

int pn_gather_step(pn_gather* gh, const void* local_dev, void* stream) {
  PN_G_BEGIN
  auto* g = reinterpret_cast<Gather*>(gh);
  PN_REQUIRE(g && local_dev && g->slab, "pn_gather_step: not connected");
  PN_REQUIRE(reinterpret_cast<uintptr_t>(local_dev) % 16 == 0, "pn_gather_step: local buffer must be 16-byte aligned");
  PN_CUDA_CHECK(cudaSetDevice(g->device));
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned long long seq = ++g->seq;
  char* slot = g->slab + (seq & 1ull) * g->slot_bytes();
  uint4* dst = reinterpret_cast<uint4*>(slot + g->slice * g->rank);
  const size_t n16 = g->bytes / 16;
  const int blocks = static_cast<int>(std::min<size_t>((n16 + 1023) / 1024, static_cast<size_t>(g->num_sms)));
  const bool is_root = g->rank == g->root;
  // a slot last used by step seq - 2 is free once the root has begun step seq - 1
  k_gather_push<<<std::max(blocks, 1), 256, 0, s>>>(static_cast<const uint4*>(local_dev), dst, n16, g->counter,
                                                    &g->ctl->flag[g->rank][0], seq, is_root ? nullptr : &g->ctl->ack[0],
                                                    seq >= 2 ? seq - 1 : 0, is_root ? &g->ctl->ack[0] : nullptr, g->status_dev);
  if (is_root && g->world > 1) k_gather_wait<<<1, 32, 0, s>>>(g->ctl, g->world, seq, g->status_dev);
  PN_CUDA_CHECK(cudaGetLastError());
  PN_G_END
}

int pn_gather_result(pn_gather* gh, void** slab_dev_out, int64_t* slice_stride_out) {
  PN_G_BEGIN
  auto* g = reinterpret_cast<Gather*>(gh);
  PN_REQUIRE(g && slab_dev_out && g->rank == g->root && g->slab, "pn_gather_result: only the root holds the results");
  *slab_dev_out = g->slab + (g->seq & 1ull) * g->slot_bytes();
  if (slice_stride_out) *slice_stride_out = static_cast<int64_t>(g->slice);
  PN_G_END
}

int pn_gather_status(pn_gather* gh, int* status_out) {
  PN_G_BEGIN
  auto* g = reinterpret_cast<Gather*>(gh);
  PN_REQUIRE(g && status_out, "pn_gather_status: null argument");
  *status_out = *reinterpret_cast<volatile int*>(g->status_host);
  PN_G_END
}

int pn_gather_destroy(pn_gather* gh) {
  PN_G_BEGIN
  auto* g = reinterpret_cast<Gather*>(gh);
  if (g) {
    cudaSetDevice(g->device);
    cudaDeviceSynchronize();
    if (g->slab) {
      if (g->mapped) cudaIpcCloseMemHandle(g->slab);
      else cudaFree(g->slab);
    }
    if (g->counter) cudaFree(g->counter);
    if (g->status_host) cudaFreeHost(g->status_host);
    delete g;
  }
  PN_G_END
}

}  // extern "C"
