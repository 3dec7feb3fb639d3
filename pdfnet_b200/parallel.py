"""Data-parallel sharding of frames across ranks (one process per GPU, no collective).

Every frame (and every hand inside it) is independent in inference (SURVEY.md
section 8e), so the batch is split into contiguous per-rank slices exactly as the
reference's DistributedSampler/DDP setup does (main.py:69-79); nothing is
exchanged on the data path.  ``gather_results`` exists for callers that want the
full result on every rank and for the world_size-2 gloo tests.
"""
import ctypes
import os

import torch
import torch.distributed as dist


def shard_range(n_items, rank, world_size):
    """Contiguous, balanced slice [lo, hi) of ``n_items`` for ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_rank_world():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for world size 1)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def gather_results(local, n_items, dim=0):
    """All-gather per-rank result slices (unequal sizes allowed) back into frame order."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_items, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad_shape = list(local.shape)
    pad_shape[dim] = maxn
    buf = local.new_zeros(pad_shape)
    buf.narrow(dim, 0, local.shape[dim]).copy_(local)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return torch.cat([o.narrow(dim, 0, hi - lo) for o, (lo, hi) in zip(outs, sizes)], dim)


def max_over_ranks(value, device):
    """Max of a python float over ranks (timing: the slowest rank defines the step)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- host staging buffers -------------------------------------------------------------------------
_cudart = None
_host_allocs = []            # (pointer, nbytes): kept for the life of the process (staging buffers are long-lived)


def pinned_like(t, write_combined=True):
    """A page-locked HOST copy of ``t`` (same shape, strides and memory format) for host->device staging.

    write_combined=True allocates with ``cudaHostAllocWriteCombined``: the pages are not snooped by the CPU
    caches while the GPU's copy engine reads them, which is what one wants for buffers the CPU only ever WRITES
    (inputs waiting for their ``copy_(non_blocking=True)``) when several GPUs pull from the same host memory
    (bench.py e2e at N > 1).  CPU READS of such a buffer are slow - never use it for results.  Falls back to
    ``Tensor.pin_memory()`` when the CUDA runtime library cannot be loaded (and returns ``t`` itself in a CPU-only
    process)."""
    global _cudart
    if not torch.cuda.is_available():
        return t                                         # CPU-only process: nothing to page-lock for
    if not write_combined:
        return t.pin_memory()
    try:
        if _cudart is None:
            _cudart = ctypes.CDLL("libcudart.so.12")
    except OSError:
        try:
            _cudart = ctypes.CDLL("libcudart.so")
        except OSError:
            return t.pin_memory()
    src = t.detach()
    if not (src.is_contiguous() or src.is_contiguous(memory_format=torch.channels_last)):
        src = src.contiguous()
    nbytes = max(1, src.numel() * src.element_size())
    ptr = ctypes.c_void_p()
    # portable (1) | mapped (2: kernels may read the buffer in place, _lib.host_ptr) | write-combined (4)
    rc = _cudart.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(1 | 2 | 4))
    if rc != 0 or not ptr.value:
        return t.pin_memory()
    _host_allocs.append((ptr.value, nbytes))
    raw = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
    flat = torch.frombuffer(raw, dtype=torch.uint8, count=nbytes).view(src.dtype)
    out = torch.as_strided(flat, src.shape, src.stride())
    out.copy_(src)
    return out
