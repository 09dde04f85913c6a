"""Multi-GPU sharding of the hot path: one process per GPU, reads partitioned by cumulative bases, index replicated,
per-target count vectors combined with ONE all-reduce.

The reduction mirrors what the reference does with `Counter.update` when it merges per-sample tallies
(/root/reference/monica/genomes/aligner.py:282-302): counts are additive, reads are independent, so there is no exchange
step on the data path -- every read stays on the rank that owns it (also across sequential .mmi chunks).

`torch.distributed` is plumbing only: NCCL on the GPU box (the count vector stays in HBM, see mb_count_device_ptr),
gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def split_by_bases(off: np.ndarray, world: int) -> list[tuple[int, int]]:
    """Contiguous read ranges [lo, hi) per rank with (nearly) equal base counts.  `off` = int64[n_reads+1] byte offsets.
    Every read lands on exactly one rank; ranks may get empty ranges when there are fewer reads than ranks."""
    off = np.asarray(off, dtype=np.int64)
    n = len(off) - 1
    if world <= 0:
        raise ValueError("world must be positive")
    total = int(off[-1]) if n > 0 else 0
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        # first read whose start offset is >= target, never moving backwards
        i = int(np.searchsorted(off[:-1], target, side="left")) if n else 0
        cuts.append(max(cuts[-1], min(i, n)))
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_reads(cat: np.ndarray, off: np.ndarray, rank: int, world: int):
    """This rank's slice of a concatenated read batch: (cat_r, off_r, first_read_index)."""
    lo, hi = split_by_bases(off, world)[rank]
    o = np.asarray(off[lo:hi + 1], dtype=np.int64)
    return cat[int(o[0]):int(o[-1])], o - o[0], lo


class Comm:
    """NCCL communicator of the C library (include/monica_b200.h mb_comm_*): one per (process, device).  The 128-byte NCCL id
    is made by rank 0 and handed to the other ranks through `torch.distributed` (any backend) -- plumbing only; the
    all-reduce itself is issued by the library on its own mapping stream, on the count vector that never left HBM."""

    def __init__(self, device: int, group=None):
        import ctypes as C
        import torch.distributed as dist
        from . import _lib
        L = _lib.lib()
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        buf = (C.c_uint8 * 128)()
        if self.rank == 0:
            _lib.check(L.mb_comm_unique_id(buf))
        box = [bytes(buf)]
        dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_uint8 * 128).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        _lib.check(L.mb_comm_init(device, self.rank, self.world, ident, C.byref(self._h)))

    def handle(self):
        return self._h

    def free(self):
        if getattr(self, "_h", None):
            from . import _lib
            _lib.lib().mb_comm_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def allreduce_counts(counts, group=None):
    """Sum an int64 count vector over all ranks through torch.distributed (host-side logic and CPU tests; the product path
    reduces the device-resident vector with mb_allreduce_counts, see map_and_count_sharded).  Accepts a numpy array or a torch
    tensor.  Under the NCCL backend a numpy array travels through a CUDA tensor.  Without an initialised process group it is
    the identity, so single-GPU callers need no special case."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counts
    if isinstance(counts, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(counts, dtype=np.int64).copy())
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return t.cpu().numpy()
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def map_and_count_sharded(aligner, cat: np.ndarray, off: np.ndarray, mode: str = "query_length", mapq_min: int = 60, group=None,
                          comm: "Comm | None" = None, cigars: bool = True):
    """Map this rank's share of a batch on its GPU and return (global per-target counts, global [mapped, unmapped,
    ambiguous] read classes, local Hits).  `aligner` is a monica_b200.mappy_shim.Aligner bound to this rank's device.
    With `comm` (a shard.Comm) the counting and the all-reduce stay on the device: mb_count_last on the device-resident hits,
    then ONE NCCL all-reduce issued by the library (mb_allreduce_counts).  Without it the host-side vector is reduced through
    torch.distributed (CPU tests over gloo)."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    cat_r, off_r, _ = shard_reads(cat, off, rank, world)
    hits = aligner.map_batch(cat=cat_r, off=off_r, cigars=cigars)
    if comm is not None:
        counts, ncls = aligner.count_last(mapq_min, mode, comm=comm)
        return counts, ncls, hits
    counts, ncls, _, _ = aligner.count(hits, mapq_min, mode)
    both = np.concatenate([np.asarray(counts, dtype=np.int64), np.asarray(ncls, dtype=np.int64)])
    both = allreduce_counts(both, group)
    return both[:len(counts)], both[len(counts):], hits
