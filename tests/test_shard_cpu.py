"""Multi-GPU host logic on CPU: read sharding by bases and the single count all-reduce, world_size 2 over gloo.
The mapper on each rank is the CPU oracle here (tests may use it as the checker); on the GPU box the same code path runs
with the CUDA library and NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def test_split_by_bases_properties():
    from monica_b200.shard import split_by_bases, shard_reads
    rng = np.random.default_rng(5)
    for n, world in ((0, 2), (1, 4), (7, 2), (1000, 8), (33, 3)):
        lens = rng.integers(0, 5000, n)
        off = np.zeros(n + 1, np.int64); off[1:] = np.cumsum(lens)
        parts = split_by_bases(off, world)
        assert len(parts) == world and parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:])) and all(lo <= hi for lo, hi in parts)
        if n >= 100:
            loads = [off[hi] - off[lo] for lo, hi in parts]
            assert max(loads) - min(loads) <= 2 * lens.max()
        cat = rng.integers(0, 4, int(off[-1])).astype(np.uint8)
        back = [shard_reads(cat, off, r, world) for r in range(world)]
        assert sum(len(b[0]) for b in back) == len(cat)
        for (c, o, lo), (plo, phi) in zip(back, parts):
            assert lo == plo and len(o) == phi - plo + 1 and o[0] == 0 and np.array_equal(c, cat[off[plo]:off[phi]])


def classify(hits, mapq_min):
    """monica's hit filter + best_hit (reference aligner.py:193-195, 328-339): (0 mapped | 1 unmapped | 2 ambiguous, best hit)."""
    kept = [h for h in hits if h["is_primary"] and h["mapq"] >= mapq_min]
    if not kept:
        return 1, None
    if len(kept) == 1:
        return 0, kept[0]
    best, margin, arg = float("inf"), 0, None
    for h in kept:
        r = float(h["nm"]) / h["mlen"]
        if r <= best:
            margin, best, arg = best - r, r, h
    return (0, arg) if margin else (2, None)


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from monica_b200 import synth
        from monica_b200.shard import shard_reads, allreduce_counts
        from oracle import oracle as O
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from test_shard_cpu import classify
        names, seqs = synth.make_genomes(3, 3, 40000, strain_frac=0.34)
        reads, _ = synth.simulate_reads(4, seqs, 24, 2000, 0.10, junk_frac=0.1)
        cat, off = synth.concat_reads(reads)
        cat_r, off_r, first = shard_reads(cat, off, rank, world)
        oidx = O.Index(names, seqs)
        hits, _ = oidx.map_batch(cat_r, off_r, n_threads=2)
        counts = np.zeros(len(names) + 3, np.int64)
        for i, hs in enumerate(hits):
            cls, best = classify(hs, 60)
            counts[len(names) + cls] += 1
            if cls == 0:
                counts[best["rid"]] += int(off_r[i + 1] - off_r[i])
        total = allreduce_counts(counts)
        q.put((rank, total.tolist(), first))
    finally:
        dist.destroy_process_group()


def test_two_ranks_allreduce_equals_single_process():
    import torch.multiprocessing as mp
    from monica_b200 import synth
    from oracle import oracle as O
    O.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    names, seqs = synth.make_genomes(3, 3, 40000, strain_frac=0.34)
    reads, _ = synth.simulate_reads(4, seqs, 24, 2000, 0.10, junk_frac=0.1)
    cat, off = synth.concat_reads(reads)
    oidx = O.Index(names, seqs)
    hits, _ = oidx.map_batch(cat, off, n_threads=2)
    want = np.zeros(len(names) + 3, np.int64)
    for i, hs in enumerate(hits):
        cls, best = classify(hs, 60)
        want[len(names) + cls] += 1
        if cls == 0:
            want[best["rid"]] += int(off[i + 1] - off[i])
    assert sorted(g[0] for g in got) == [0, 1]
    for _, total, _ in got:
        assert total == want.tolist()
    assert want[:len(names)].sum() > 0


class _TwinAligner:
    """What shard.map_and_count_sharded needs from an aligner (map_batch + count), served by the CPU oracle behind the C ABI
    of include/monica_b200.h (oracle/mm2o_abi.c) through the same raw ctypes calls the CUDA library gets."""

    def __init__(self, names, seqs):
        import ctypes as C
        import abi_harness as H
        from monica_b200 import _lib
        self.C, self.L, self._lib = C, H.oracle_library(), _lib
        self.opt = _lib.Opt()
        assert self.L.mb_opt_init(C.byref(self.opt)) == 0
        n = len(names)
        self._arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
        c_names = (C.c_char_p * n)(*[s.encode() for s in names])
        c_seqs = (C.c_void_p * n)(*[a.ctypes.data for a in self._arrs])
        lens = np.array([len(a) for a in self._arrs], np.int64)
        self.idx = C.c_void_p()
        assert self.L.mb_index_build(0, n, c_names, c_seqs, lens.ctypes.data_as(C.c_void_p), 10, 15, C.byref(self.idx)) == 0
        self.n_seq = n

    def map_batch(self, cat=None, off=None, cigars=True):
        C = self.C
        cat = np.ascontiguousarray(cat, np.uint8); off = np.ascontiguousarray(off, np.int64)
        h = C.c_void_p()
        assert self.L.mb_map_batch_ex(self.idx, C.byref(self.opt), cat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                      len(off) - 1, 3 if cigars else 1, C.byref(h), None) == 0
        return h

    def count(self, hits, mapq_min, mode):
        C = self.C
        m = {"basic": 0, "query_length": 1, "matching": 2}.get(mode, -1)
        counts, ncls = np.zeros(self.n_seq, np.int64), np.zeros(3, np.int64)
        assert self.L.mb_count(self.idx, hits, mapq_min, m, counts.ctypes.data_as(C.c_void_p), ncls.ctypes.data_as(C.c_void_p), None, None) == 0
        return counts, ncls, None, None


def _worker_product(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from monica_b200 import synth
        from monica_b200.shard import map_and_count_sharded
        from test_shard_cpu import _TwinAligner
        names, seqs = synth.make_genomes(3, 3, 40000, strain_frac=0.34)
        reads, _ = synth.simulate_reads(4, seqs, 24, 2000, 0.10, junk_frac=0.1)
        cat, off = synth.concat_reads(reads)
        al = _TwinAligner(names, seqs)
        out = {}
        for mode in ("basic", "query_length", "matching"):
            counts, ncls, _ = map_and_count_sharded(al, cat, off, mode=mode, mapq_min=60)   # this rank's share, then the all-reduce
            out[mode] = (np.asarray(counts).tolist(), np.asarray(ncls).tolist())
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_product_sharding_function_over_gloo_equals_one_process():
    """monica_b200.shard.map_and_count_sharded itself, world_size 2 over gloo: every rank maps its share through the C ABI
    (served here by the oracle twin: no GPU), counts with mb_count, and the one all-reduce must give every rank the counts of
    the unsharded batch in all three modes."""
    import torch.multiprocessing as mp
    from monica_b200 import synth
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_product, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    names, seqs = synth.make_genomes(3, 3, 40000, strain_frac=0.34)
    reads, _ = synth.simulate_reads(4, seqs, 24, 2000, 0.10, junk_frac=0.1)
    cat, off = synth.concat_reads(reads)
    al = _TwinAligner(names, seqs)
    h = al.map_batch(cat=cat, off=off)
    assert sorted(g[0] for g in got) == [0, 1]
    for mode in ("basic", "query_length", "matching"):
        counts, ncls, _, _ = al.count(h, 60, mode)
        assert counts.sum() > 0 and ncls.sum() == len(reads)
        for _, out in got:
            assert out[mode] == (counts.tolist(), ncls.tolist()), mode
