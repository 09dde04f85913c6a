"""A `mappy`-shaped module surface backed by the sm_100a library.

monica's aligner helper touches mappy in exactly four places (/root/reference/monica/genomes/aligner.py):
  * :45-46  mappy.Aligner(fn_idx_in=<fna.gz>, preset='map-ont', best_n=15, fn_idx_out=<mmi>)  (index build + dump)
  * :59     mappy.Aligner(fn_idx_in=<mmi>)                                                     (index load)
  * :47,60  truthiness of the Aligner (falsy on failure)
  * :193,215 `for hit in index.map(str(seq))` reading hit.is_primary .mapq .ctg .NM .mlen      (:194-195,216-217)
This module keeps those names, argument meanings and failure behaviour.  `Aligner.map` exists for drop-in use and
for tests; the product path batches whole FASTQ files through `Aligner.map_batch` (monica_b200/aligner.py).

Like mappy 2.17, the loading constructor ignores preset/best_n recorded at build time: k and w come from the index
file, every other mapping option is mm_mapopt_init()'s default (best_n = 5).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

from . import _lib
from ._lib import Hits, Stats, check, lib

__version__ = "2.17-b200"


class Alignment:
    """Field-for-field the object mappy yields (python/mappy.pyx Alignment)."""
    __slots__ = ("ctg", "ctg_len", "r_st", "r_en", "strand", "q_st", "q_en", "mapq", "is_primary", "mlen", "blen", "NM",
                 "trans_strand", "read_num", "cigar", "score", "rid")

    def __init__(self, ctg, ctg_len, r_st, r_en, strand, q_st, q_en, mapq, is_primary, mlen, blen, NM, cigar, score, rid):
        self.ctg, self.ctg_len, self.r_st, self.r_en = ctg, ctg_len, r_st, r_en
        self.strand, self.q_st, self.q_en, self.mapq = strand, q_st, q_en, mapq
        self.is_primary, self.mlen, self.blen, self.NM = is_primary, mlen, blen, NM
        self.trans_strand, self.read_num = 0, 1
        self.cigar = cigar
        self.score = score   # dp_max: not a mappy 2.17 field; exposed for parity tests
        self.rid = rid

    @property
    def cigar_str(self):
        return "".join(f"{l}{'MIDNSH'[op]}" for l, op in self.cigar)

    def __str__(self):
        strand = "+" if self.strand > 0 else "-" if self.strand < 0 else "?"
        tp = "tp:A:P" if self.is_primary else "tp:A:S"
        return "\t".join(map(str, [self.q_st, self.q_en, strand, self.ctg, self.ctg_len, self.r_st, self.r_en, self.mlen,
                                   self.blen, self.mapq, tp, "ts:A:.", "cg:Z:" + self.cigar_str]))


class PackedReads:
    """Library-owned packed batch (mb_reads_pack / mb_fastq_pack); freed with the object."""

    def __init__(self, handle, n_reads: int):
        self.handle, self.n_reads = handle, n_reads

    @property
    def upload_bytes(self) -> int:
        return int(lib().mb_packed_upload_bytes(self.handle))

    def __del__(self):
        try:
            if self.handle:
                lib().mb_packed_free(self.handle)
                self.handle = None
        except Exception:
            pass


class Aligner:
    """mappy.Aligner(fn_idx_in=None, preset=None, k=None, w=None, best_n=None, n_threads=3, fn_idx_out=None, seq=None)."""

    def __init__(self, fn_idx_in=None, preset=None, k=None, w=None, min_cnt=None, min_chain_score=None, min_dp_score=None,
                 bw=None, best_n=None, n_threads=3, fn_idx_out=None, max_frag_len=None, extra_flags=None, seq=None,
                 scoring=None, device=None, names=None, seqs=None):
        self._idx = None
        self._lock = threading.Lock()
        L = lib()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if L.mb_device_count() > 1 else 0
            device %= max(1, L.mb_device_count())
        self.device = device
        if preset not in (None, "map-ont"):
            raise ValueError("only the map-ont preset is implemented (it is the one monica uses)")
        kk = 15 if k is None else int(k)
        ww = 10 if w is None else int(w)
        self.opt = _lib.default_opt()
        if best_n is not None:
            self.opt.best_n = int(best_n)
        if min_cnt is not None:
            self.opt.min_cnt = int(min_cnt)
        if min_chain_score is not None:
            self.opt.min_chain_score = int(min_chain_score)
        if min_dp_score is not None:
            self.opt.min_dp_max = int(min_dp_score)
        if bw is not None:
            self.opt.bw = int(bw)
        h = C.c_void_p()
        try:
            if seqs is not None:
                n = len(seqs)
                bufs = [np.ascontiguousarray(np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else
                                             np.frombuffer(s.encode(), dtype=np.uint8) if isinstance(s, str) else
                                             np.asarray(s, dtype=np.uint8)) for s in seqs]
                nm = (C.c_char_p * n)(*[x.encode() if isinstance(x, str) else x for x in names])
                sp = (C.c_void_p * n)(*[b.ctypes.data for b in bufs])
                lens = np.array([len(b) for b in bufs], dtype=np.int64)
                check(L.mb_index_build(device, n, nm, sp, lens.ctypes.data_as(C.c_void_p), ww, kk, C.byref(h)))
            elif seq is not None:
                b = np.frombuffer(seq.encode() if isinstance(seq, str) else seq, dtype=np.uint8)
                nm = (C.c_char_p * 1)(b"N/A")
                sp = (C.c_void_p * 1)(b.ctypes.data)
                lens = np.array([len(b)], dtype=np.int64)
                check(L.mb_index_build(device, 1, nm, sp, lens.ctypes.data_as(C.c_void_p), ww, kk, C.byref(h)))
            elif fn_idx_in is not None:
                path = os.fspath(fn_idx_in)
                with open(path, "rb") as fh:
                    magic = fh.read(4)
                if magic == b"MMI\x02":
                    check(L.mb_index_load(device, path.encode(), C.byref(h)))
                else:
                    check(L.mb_index_build_fasta(device, path.encode(), ww, kk, C.byref(h)))
                    if fn_idx_out is not None:
                        check(L.mb_index_save(h, os.fspath(fn_idx_out).encode()))
            else:
                return
        except _lib.MonicaB200Error as e:
            if e.code == -2:   # no device / CUDA failure is not "bad index": never mask it as falsy
                raise
            self._idx = None   # mappy: failure to open/build leaves a falsy Aligner (aligner.py:47-48,60-61)
            self._error = str(e)
            return
        except OSError as e:
            self._idx = None
            self._error = str(e)
            return
        self._idx = h
        self._names = [L.mb_index_seq_name(h, i).decode() for i in range(L.mb_index_n_seq(h))]
        self._lens = [int(L.mb_index_seq_len(h, i)) for i in range(len(self._names))]
        k_, w_ = C.c_int(0), C.c_int(0)
        L.mb_index_kw(h, C.byref(k_), C.byref(w_))
        self.k, self.w = k_.value, w_.value
        self.last_stats = None

    def __bool__(self):
        return self._idx is not None

    def __del__(self):
        try:
            if getattr(self, "_idx", None):
                lib().mb_index_free(self._idx)
                self._idx = None
        except Exception:
            pass

    # ---- mappy surface ----
    @property
    def seq_names(self):
        return list(self._names)

    @property
    def n_seq(self):
        return len(self._names)

    @property
    def mid_occ(self):
        return lib().mb_index_mid_occ(self._idx)

    def handle(self):
        return self._idx

    def map(self, seq, seq2=None, buf=None, cs=False, MD=False, max_frag_len=None, extra_flags=None):
        if self._idx is None:
            return
        if seq2 is not None:
            raise NotImplementedError("paired mapping is not on monica's path")
        b = seq.encode() if isinstance(seq, str) else bytes(seq)
        hits = self.map_batch([b])
        for i in range(hits.n):
            yield self._alignment(hits, i)

    def _alignment(self, hits: Hits, i: int) -> Alignment:
        rid = int(hits.rid[i])
        cg = hits.cigar(i)
        cigar = [[int(c >> 4), int(c & 0xf)] for c in cg]
        return Alignment(self._names[rid], self._lens[rid], int(hits.rs[i]), int(hits.re[i]), -1 if hits.rev[i] else 1,
                         int(hits.qs[i]), int(hits.qe[i]), int(hits.mapq[i]), bool(hits.is_primary[i]), int(hits.mlen[i]),
                         int(hits.blen[i]), int(hits.nm[i]), cigar, int(hits.dp_max[i]), rid)

    # ---- batched entry (the product path) ----
    def map_batch(self, seqs=None, cat: np.ndarray | None = None, off: np.ndarray | None = None, cigars: bool = True) -> Hits:
        """Map many reads in one device pipeline.  Either `seqs` (list of bytes/str/uint8 arrays) or a
        pre-concatenated (cat uint8[total], off int64[n+1]) pair.  `cigars=False` skips the device->host copy of the CIGAR
        pool (monica reads only ctg / NM / mlen).  Thread-safe."""
        if self._idx is None:
            raise _lib.MonicaB200Error(-1, "empty index")
        if cat is None:
            arrs = [np.frombuffer(s.encode() if isinstance(s, str) else s, dtype=np.uint8) if not isinstance(s, np.ndarray)
                    else s.astype(np.uint8, copy=False) for s in seqs]
            off = np.zeros(len(arrs) + 1, dtype=np.int64)
            if arrs:
                off[1:] = np.cumsum([len(a) for a in arrs])
            cat = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        n = len(off) - 1
        h = C.c_void_p()
        st = Stats()
        check(lib().mb_map_batch_ex(self._idx, C.byref(self.opt), cat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p),
                                    n, 3 if cigars else 1, C.byref(h), C.byref(st)))
        self.last_stats = st.as_dict()
        return Hits(h, n)

    # ---- packed reads: reduced once on the host to 2-bit words + runs of ambiguous bases, a quarter of the bytes over PCIe ----
    @staticmethod
    def pack_reads(cat: np.ndarray, off: np.ndarray, n_threads: int = 0) -> "PackedReads":
        """Pack a concatenated ASCII batch (cat uint8[total], off int64[n+1]) for map_packed.  Host only, `n_threads` threads
        (0: all cores); the words land in page-locked memory when a device is present."""
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        h = C.c_void_p()
        check(lib().mb_reads_pack(cat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(off) - 1, n_threads, C.byref(h)))
        return PackedReads(h, len(off) - 1)

    def map_packed(self, packed: "PackedReads", cigars: bool = True) -> Hits:
        """map_batch on a packed batch: same hits, 0.25 B/base uploaded instead of 1."""
        if self._idx is None:
            raise _lib.MonicaB200Error(-1, "empty index")
        h = C.c_void_p()
        st = Stats()
        check(lib().mb_map_packed(self._idx, C.byref(self.opt), packed.handle, 3 if cigars else 1, C.byref(h), C.byref(st)))
        self.last_stats = st.as_dict()
        return Hits(h, packed.n_reads)

    # ---- device-resident reads (measurement of the kernels alone; a caller that maps one batch against several option sets) ----
    def reads_upload(self, cat: np.ndarray, off: np.ndarray):
        """Upload a concatenated batch once; returns an opaque handle for map_resident / reads_free."""
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        h = C.c_void_p()
        check(lib().mb_reads_upload(self._idx, cat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(off) - 1, C.byref(h)))
        return h

    def reads_free(self, handle):
        lib().mb_reads_free(handle)

    def map_resident(self, handle, n_reads: int, want_hits: bool = False):
        """Map an uploaded batch.  With want_hits=False nothing is copied back: follow with count_last()."""
        h = C.c_void_p()
        st = Stats()
        check(lib().mb_map_resident(self._idx, C.byref(self.opt), handle, 1 if want_hits else 0, C.byref(h), C.byref(st)))
        self.last_stats = st.as_dict()
        if want_hits:
            return Hits(h, n_reads)
        lib().mb_hits_free(h)
        return None

    def count(self, hits: Hits, mapq_min: int = 60, mode: str | None = "basic"):
        """monica's hit filter + best_hit + per-target sum on the device (aligner.py:193-195,225-263,328-339).
        Returns (counts int64[n_seq], n_class [mapped, unmapped, ambiguous], read_class int8[n_reads], read_best int64[n_reads])."""
        m = {"basic": 0, "query_length": 1, "matching": 2}.get(mode, 3)
        counts = np.zeros(self.n_seq, dtype=np.int64)
        ncls = np.zeros(3, dtype=np.int64)
        rcls = np.zeros(max(1, hits.n_reads), dtype=np.int8)
        rbest = np.zeros(max(1, hits.n_reads), dtype=np.int64)
        check(lib().mb_count(self._idx, hits.handle(), mapq_min, m, counts.ctypes.data_as(C.c_void_p),
                             ncls.ctypes.data_as(C.c_void_p), rcls.ctypes.data_as(C.c_void_p), rbest.ctypes.data_as(C.c_void_p)))
        return counts, ncls, rcls[:hits.n_reads], rbest[:hits.n_reads]

    def count_last(self, mapq_min: int = 60, mode: str | None = "basic", comm=None):
        """The same filter + best_hit + per-target sum on the DEVICE-RESIDENT hits of this thread's last map_batch (no host
        round trip of the hit arrays), optionally followed by the one NCCL all-reduce over the ranks of `comm`
        (monica_b200.shard.Comm).  Returns (counts int64[n_seq], n_class [mapped, unmapped, ambiguous])."""
        m = {"basic": 0, "query_length": 1, "matching": 2}.get(mode, 3)
        counts = np.zeros(self.n_seq, dtype=np.int64)
        ncls = np.zeros(3, dtype=np.int64)
        if comm is None:
            check(lib().mb_count_last(self._idx, mapq_min, m, counts.ctypes.data_as(C.c_void_p), ncls.ctypes.data_as(C.c_void_p)))
        else:
            check(lib().mb_count_last(self._idx, mapq_min, m, None, None))
            check(lib().mb_allreduce_counts(self._idx, comm.handle(), counts.ctypes.data_as(C.c_void_p), ncls.ctypes.data_as(C.c_void_p)))
        return counts, ncls

    def normalize_last(self, sample_alignment, genomes_length):
        """normalizer() (aligner.py:305-319) for one sample, computed on the device from the count vector of this thread's
        last count() (in a multi-GPU run: after the all-reduce).  `sample_alignment` = {tax_unit: Counter({accession: n})}
        supplies the keys and -- because the reference adds the BPB floats in dict order -- the summation order; the values
        come from the device.  Returns {tax_unit: Counter({accession: BPM})}, bit-identical to the reference's floats."""
        from collections import Counter
        gid = {}
        group = np.full(max(1, self.n_seq), -1, dtype=np.int32)
        for i, name in enumerate(self.seq_names):
            parts = name.split(':')
            if len(parts) >= 2:
                group[i] = gid.setdefault((parts[0], parts[1]), len(gid))
        order, glen = [], np.ones(max(1, len(gid)), dtype=np.float64)
        for tax_unit, counter in sample_alignment.items():
            for accession in counter:
                g = gid[(tax_unit, accession)]
                glen[g] = float(genomes_length[accession])
                order.append(g)
        order_a = np.asarray(order if order else [0], dtype=np.int32)
        bpm = np.zeros(max(1, len(gid)), dtype=np.float64)
        check(lib().mb_normalize_last(self._idx, group.ctypes.data_as(C.c_void_p), max(1, len(gid)), glen.ctypes.data_as(C.c_void_p),
                                      order_a.ctypes.data_as(C.c_void_p), len(order), bpm.ctypes.data_as(C.c_void_p)))
        out = {}
        for tax_unit, counter in sample_alignment.items():
            out[tax_unit] = Counter()
            for accession in counter:
                out[tax_unit][accession] = float(bpm[gid[(tax_unit, accession)]])
        return out


def fastx_read(fn, read_comment=False):
    """mappy.fastx_read: yields (name, seq, qual[, comment])."""
    from .fastx import parse_fastx
    for name, comment, seq, qual in parse_fastx(fn):
        if read_comment:
            yield name, seq, qual, comment
        else:
            yield name, seq, qual


def revcomp(seq):
    tab = str.maketrans("ACGTUNacgtun", "TGCAANtgcaan")
    return seq.translate(tab)[::-1]
