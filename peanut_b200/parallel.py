"""Multi-GPU plumbing: the path shards over independent environments, one process per GPU.

There is no data-path collective - every environment's frame, maps and pose live on exactly one rank.  The only
exchange is the result gather to the rank that hosts the (CPU) planner, done with ``torch.distributed`` (NCCL over
NVLink on the GPU box, gloo in the CPU tests).  The reference itself is single-process, single-GPU, batch 1
(nav/collect.py:32-33); this module is what replaces "run N copies by hand".
"""
import ctypes
import os

import torch
import torch.distributed as dist

from . import _lib


def env_range(num_envs, world_size, rank):
    """Contiguous block partition: environment e lives on rank e // ceil(num_envs / world_size)
    (BASELINE.json configs[3]: 64 envs, 8 per GPU).  Returns (first, last_exclusive) for `rank`; ranks past the
    end get an empty range."""
    if num_envs < 0 or world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad partition arguments")
    per = -(-num_envs // world_size) if num_envs else 0
    lo = min(rank * per, num_envs)
    return lo, min(lo + per, num_envs)


def owner_of(env, num_envs, world_size):
    per = -(-num_envs // world_size)
    if not (0 <= env < num_envs):
        raise IndexError(env)
    return env // per


def init_from_env(backend=None):
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun).  Returns (rank, local_rank, world_size)."""
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kw)
    return rank, local, world


def max_over_ranks(values, device=None):
    """Element-wise maximum of a list of floats over all ranks (timing is reported as the slowest rank's)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def gather_env_results(local, num_envs, dst=0):
    """local: this rank's per-environment results [E_local, ...] (block partition of `num_envs`).  Returns the
    [num_envs, ...] tensor on rank `dst` (None elsewhere).  Ranks may hold different numbers of environments."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-num_envs // world)
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    parts = []
    for r in range(world):
        lo, hi = env_range(num_envs, world, r)
        parts.append(out[r][:hi - lo])
    return torch.cat(parts, 0)


class _DeviceView:
    """A device allocation owned by libpeanut_b200.so, exposed to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, shape, typestr="<f4"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


class PeerGather:
    """The result gather as a one-sided push over NVLink peer memory (``pn_gather_*``, csrc/gather.cu): every rank writes its
    per-environment results straight into the root's slab with one kernel, the root waits for the per-rank flags.  All
    buffers are allocated once; ``step`` enqueues on the caller's current stream and never synchronises the host.

    ``like``: this rank's result tensor [E_local, ...] (float32, contiguous, same shape on every rank).  The 64-byte CUDA IPC
    handle of the root's slab travels through ``torch.distributed`` (any backend) once, at construction."""

    def __init__(self, ctx, like, root=0):
        self.ctx, self.lib = ctx, ctx.lib
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.root = int(root)
        if not (like.is_cuda and like.is_contiguous() and like.dtype == torch.float32):
            raise TypeError("PeerGather expects a contiguous float32 CUDA tensor")
        self.shape = tuple(like.shape)
        self.nbytes = like.numel() * 4
        self.device = like.device
        # every rank walks the same collectives whatever fails locally (e.g. CUDA IPC not permitted in this container):
        # the failure is agreed on at the end and raised on all ranks together
        err = None
        h = ctypes.c_void_p()
        try:
            _lib.check(self.lib.pn_gather_create(ctx.handle, self.rank, self.world, self.root, self.nbytes, ctypes.byref(h)))
        except RuntimeError as e:
            err = e
        self.handle = h if err is None else None
        if self.world > 1:
            blob = [None]
            if self.rank == self.root and err is None:
                buf = ctypes.create_string_buffer(64)
                try:
                    _lib.check(self.lib.pn_gather_export(self.handle, buf))
                    blob = [buf.raw]
                except RuntimeError as e:
                    err = e
            dist.broadcast_object_list(blob, src=self.root)
            if self.rank != self.root and err is None:
                try:
                    if blob[0] is None:
                        raise RuntimeError("the root could not export its slab")
                    _lib.check(self.lib.pn_gather_connect(self.handle, ctypes.create_string_buffer(blob[0], 64)))
                except RuntimeError as e:
                    err = e
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32,
                              device=self.device if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                if self.handle is not None:
                    self.lib.pn_gather_destroy(self.handle)
                    self.handle = None
                raise RuntimeError(f"PeerGather: peer-memory gather unavailable on at least one rank ({err})")
        elif err is not None:
            raise err
        self._views = {}

    def step(self, local):
        """Enqueue one gather step on the current stream.  On the root, work enqueued afterwards on the same stream sees
        every rank's results in ``result()``."""
        if tuple(local.shape) != self.shape or not local.is_contiguous() or local.dtype != torch.float32:
            raise TypeError("PeerGather.step: tensor does not match the one the gather was created for")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _lib.check(self.lib.pn_gather_step(self.handle, local.data_ptr(), ctypes.c_void_p(stream)))

    def result(self):
        """Root: [world, E_local, ...] float32 view of the slot the latest step filled (no copy; valid until the
        next-but-one step).  Other ranks: None."""
        if self.rank != self.root:
            return None
        ptr, stride = ctypes.c_void_p(), ctypes.c_int64()
        _lib.check(self.lib.pn_gather_result(self.handle, ctypes.byref(ptr), ctypes.byref(stride)))
        key = ptr.value
        if key not in self._views:
            per = stride.value // 4
            flat = torch.as_tensor(_DeviceView(ptr.value, (self.world, per)), device=self.device)
            self._views[key] = flat[:, :self.nbytes // 4].unflatten(1, self.shape)
        return self._views[key]

    def status(self):
        """0 = ok; 2 / 3 = a bounded wait timed out (results of that step are invalid)."""
        st = ctypes.c_int(0)
        _lib.check(self.lib.pn_gather_status(self.handle, ctypes.byref(st)))
        return st.value

    def close(self):
        if getattr(self, "handle", None):
            if self.world > 1 and dist.is_initialized():
                torch.cuda.synchronize(self.device)
                dist.barrier()  # nobody unmaps / frees while a peer may still push
            self.lib.pn_gather_destroy(self.handle)
            self.handle = None


def build_synchronised(build_fn):
    """Build the same engines on every rank with IDENTICAL conv launch configurations, so that the ranks of one job produce
    bit-identical results for identical inputs: rank 0 builds first (timing whatever the imported tables do not cover),
    its launch-configuration table is broadcast, the other ranks import it and build without timing anything.
    Returns build_fn()'s result."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return build_fn()
    rank = dist.get_rank()
    out = None
    if rank == 0:
        out = build_fn()
    blob = [_lib.tuning_export() if rank == 0 else None]
    dist.broadcast_object_list(blob, src=0)
    if rank != 0:
        _lib.tuning_import(blob[0])
        out = build_fn()
    return out
