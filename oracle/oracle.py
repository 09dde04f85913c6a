"""ctypes wrapper over the CPU oracle (oracle/_build/libmm2oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  Nothing under monica_b200/ imports this module.  PARITY UNPINNED (see mm2o.h).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmm2oracle.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


class Opt(C.Structure):
    _fields_ = [
        ("seed", C.c_int), ("mid_occ_frac", C.c_float),
        ("min_cnt", C.c_int), ("min_chain_score", C.c_int), ("bw", C.c_int), ("max_gap", C.c_int),
        ("max_gap_ref", C.c_int), ("max_chain_skip", C.c_int), ("max_chain_iter", C.c_int),
        ("mask_level", C.c_float), ("pri_ratio", C.c_float), ("best_n", C.c_int),
        ("max_join_long", C.c_int), ("max_join_short", C.c_int), ("min_join_flank_sc", C.c_int),
        ("min_join_flank_ratio", C.c_float),
        ("a", C.c_int), ("b", C.c_int), ("q", C.c_int), ("e", C.c_int), ("q2", C.c_int), ("e2", C.c_int),
        ("sc_ambi", C.c_int), ("zdrop", C.c_int), ("zdrop_inv", C.c_int), ("end_bonus", C.c_int),
        ("min_dp_max", C.c_int), ("min_ksw_len", C.c_int), ("max_clip_ratio", C.c_float),
        ("max_sw_mat", C.c_int64), ("mid_occ", C.c_int),
    ]


class Hit(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "rid", "rev", "qs", "qe", "rs", "re", "mapq", "mlen", "blen", "nm",
        "dp_max", "dp_max2", "score", "score0", "cnt", "subsc", "n_sub",
        "id", "parent", "is_primary", "sam_pri", "n_cigar", "cigar_off")]


class Result(C.Structure):
    _fields_ = [
        ("n_hits", C.c_int32), ("hits", C.POINTER(Hit)),
        ("n_cigar_pool", C.c_int32), ("cigar_pool", C.POINTER(C.c_uint32)),
        ("rep_len", C.c_int32),
        ("n_mini", C.c_int64), ("n_anchor", C.c_int64), ("chain_cells", C.c_int64),
        ("dp_cells", C.c_int64), ("n_dp_calls", C.c_int64),
    ]


class MM128V(C.Structure):
    _fields_ = [("n", C.c_size_t), ("m", C.c_size_t), ("a", C.POINTER(C.c_uint64))]


class DpRec(C.Structure):
    _fields_ = [
        ("qlen", C.c_int32), ("tlen", C.c_int32), ("w", C.c_int32), ("zdrop", C.c_int32),
        ("end_bonus", C.c_int32), ("flag", C.c_int32),
        ("q_off", C.c_int64), ("t_off", C.c_int64),
        ("score", C.c_int32), ("max", C.c_int32), ("max_q", C.c_int32), ("max_t", C.c_int32),
        ("mqe", C.c_int32), ("mqe_t", C.c_int32), ("zdropped", C.c_int32), ("reach_end", C.c_int32),
        ("n_cigar", C.c_int32), ("cigar_off", C.c_int64),
    ]


class Trace(C.Structure):
    _fields_ = [
        ("enabled", C.c_int),
        ("mini", MM128V), ("anchors", MM128V),
        ("f", C.POINTER(C.c_int32)), ("p", C.POINTER(C.c_int32)), ("v", C.POINTER(C.c_int32)),
        ("n_chain_arr", C.c_int64),
        ("n_u", C.c_int32), ("u", C.POINTER(C.c_uint64)),
        ("chained", MM128V),
        ("n_dp", C.c_int64), ("m_dp", C.c_int64), ("dp", C.POINTER(DpRec)),
        ("n_seq", C.c_int64), ("m_seq", C.c_int64), ("seqpool", C.POINTER(C.c_uint8)),
        ("n_cig", C.c_int64), ("m_cig", C.c_int64), ("cigpool", C.POINTER(C.c_uint32)),
    ]


class Ez(C.Structure):
    _fields_ = [
        ("max_zd", C.c_uint32), ("max_q", C.c_int), ("max_t", C.c_int), ("mqe", C.c_int), ("mqe_t", C.c_int),
        ("mte", C.c_int), ("mte_q", C.c_int), ("score", C.c_int), ("m_cigar", C.c_int), ("n_cigar", C.c_int),
        ("reach_end", C.c_int), ("cigar", C.POINTER(C.c_uint32)),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.mm2o_opt_init.argtypes = [C.POINTER(Opt)]
        L.mm2o_hash64.restype = C.c_uint64
        L.mm2o_hash64.argtypes = [C.c_uint64, C.c_uint64]
        L.mm2o_sketch_buf.restype = C.c_int64
        L.mm2o_sketch_buf.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_void_p, C.c_int64]
        L.mm2o_idx_build.restype = C.c_void_p
        L.mm2o_idx_build.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_char_p), C.c_void_p, C.c_int, C.c_int]
        L.mm2o_idx_build_mt.restype = C.c_void_p
        L.mm2o_idx_build_mt.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.mm2o_idx_destroy.argtypes = [C.c_void_p]
        L.mm2o_idx_cal_max_occ.restype = C.c_int32
        L.mm2o_idx_cal_max_occ.argtypes = [C.c_void_p, C.c_float]
        L.mm2o_mapopt_update.argtypes = [C.POINTER(Opt), C.c_void_p]
        L.mm2o_map.restype = C.POINTER(Result)
        L.mm2o_map.argtypes = [C.c_void_p, C.POINTER(Opt), C.c_char_p, C.c_int, C.c_void_p]
        L.mm2o_result_destroy.argtypes = [C.POINTER(Result)]
        L.mm2o_trace_new.restype = C.POINTER(Trace)
        L.mm2o_trace_destroy.argtypes = [C.POINTER(Trace)]
        L.mm2o_map_batch.argtypes = [C.c_void_p, C.POINTER(Opt), C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.mm2o_batch_export.restype = C.c_int64
        L.mm2o_batch_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
        L.mm2o_ksw_extd2.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int8, C.c_void_p,
                                     C.c_int8, C.c_int8, C.c_int8, C.c_int8, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.POINTER(Ez)]
        L.mm2o_gen_simple_mat.argtypes = [C.c_int, C.c_void_p, C.c_int8, C.c_int8, C.c_int8]
        L.mm2o_ksw_ll_i16.restype = C.c_int
        L.mm2o_ksw_ll_i16.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.mm2o_logf_range.restype = None
        L.mm2o_logf_range.argtypes = [C.c_uint32, C.c_int64, C.c_void_p]
        L.mm2o_ksw_cells.restype = C.c_int64
        L.mm2o_ksw_cells.argtypes = [C.c_int, C.c_int, C.c_int]
        L.mm2o_radix_sort_128x.argtypes = [C.c_void_p, C.c_void_p]
        L.mm2o_radix_sort_64.argtypes = [C.c_void_p, C.c_void_p]
        L.mm2o_idx_get.restype = C.POINTER(C.c_uint64)
        L.mm2o_idx_get.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
        _lib = L
    return _lib


HIT_FIELDS = [n for n, _ in Hit._fields_]


def _as_bytes(seq) -> bytes:
    if isinstance(seq, bytes):
        return seq
    if isinstance(seq, str):
        return seq.encode()
    return np.asarray(seq, dtype=np.uint8).tobytes()


def default_opt() -> Opt:
    o = Opt()
    lib().mm2o_opt_init(C.byref(o))
    return o


def hash64(key: int, mask: int) -> int:
    return lib().mm2o_hash64(key, mask)


def sketch(seq, w: int = 10, k: int = 15, rid: int = 0) -> np.ndarray:
    """Return minimizers as uint64[n, 2] (x, y) -- sketch.c mm_sketch."""
    b = _as_bytes(seq)
    cap = len(b) + 16
    out = np.zeros((cap, 2), dtype=np.uint64)
    n = lib().mm2o_sketch_buf(b, len(b), w, k, rid, out.ctypes.data, cap)
    assert n <= cap
    return out[:n].copy()


def set_simd(on: bool) -> None:
    """Choose the SSE4.1 (default) or the scalar statement of the DP core; both give identical bytes."""
    lib().mm2o_ksw_set_simd(1 if on else 0)


def ksw_cells(qlen: int, tlen: int, w: int) -> int:
    return lib().mm2o_ksw_cells(qlen, tlen, w)


def radix_sort_128x(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().mm2o_radix_sort_128x(a.ctypes.data, a.ctypes.data + a.nbytes)
    return a


def radix_sort_64(a: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64).copy()
    lib().mm2o_radix_sort_64(a.ctypes.data, a.ctypes.data + a.nbytes)
    return a


def simple_mat(a=2, b=4, sc_ambi=1) -> np.ndarray:
    m = np.zeros(25, dtype=np.int8)
    lib().mm2o_gen_simple_mat(5, m.ctypes.data, a, b, sc_ambi)
    return m


def ksw_extd2(query: np.ndarray, target: np.ndarray, w: int, zdrop: int, end_bonus: int, flag: int,
              q=4, e=2, q2=24, e2=1, mat=None) -> dict:
    """ksw2_extd2_sse.c ksw_extd2_sse on nt4-coded sequences."""
    L = lib()
    if mat is None:
        mat = simple_mat()
    query = np.ascontiguousarray(query, dtype=np.uint8)
    target = np.ascontiguousarray(target, dtype=np.uint8)
    ez = Ez()
    L.mm2o_ksw_extd2(len(query), query.ctypes.data, len(target), target.ctypes.data, 5, mat.ctypes.data,
                     q, e, q2, e2, w, zdrop, end_bonus, flag, C.byref(ez))
    cig = np.array([ez.cigar[i] for i in range(ez.n_cigar)], dtype=np.uint32)
    if ez.cigar:
        C.CDLL(None).free(ez.cigar)
    return dict(max=ez.max_zd & 0x7fffffff, zdropped=ez.max_zd >> 31, max_q=ez.max_q, max_t=ez.max_t,
                mqe=ez.mqe, mqe_t=ez.mqe_t, mte=ez.mte, score=ez.score, reach_end=ez.reach_end, cigar=cig)


def ksw_ll_i16(query: np.ndarray, target: np.ndarray, q=4, e=2, mat=None):
    """ksw2_ll_sse.c ksw_ll_i16 on nt4-coded sequences: (score, qe, te)."""
    if mat is None:
        mat = simple_mat()
    query = np.ascontiguousarray(query, dtype=np.uint8)
    target = np.ascontiguousarray(target, dtype=np.uint8)
    qe, te = C.c_int(-1), C.c_int(-1)
    sc = lib().mm2o_ksw_ll_i16(len(query), query.ctypes.data, len(target), target.ctypes.data, 5, mat.ctypes.data, q, e, C.byref(qe), C.byref(te))
    return sc, qe.value, te.value


class Index:
    """Oracle index over named sequences (index.c mm_idx_gen)."""

    def __init__(self, names, seqs, w: int = 10, k: int = 15, n_threads: int = 0):
        L = lib()
        n = len(names)
        # uint8 arrays are passed by address (no copy: the 4 Gb configuration would not fit twice); anything else goes through bytes
        self._bufs = [s if isinstance(s, np.ndarray) and s.dtype == np.uint8 and s.flags["C_CONTIGUOUS"] else np.frombuffer(_as_bytes(s), dtype=np.uint8) for s in seqs]
        nm = (C.c_char_p * n)(*[x.encode() if isinstance(x, str) else x for x in names])
        sq = (C.c_void_p * n)(*[b.ctypes.data for b in self._bufs])
        lens = np.array([len(b) for b in self._bufs], dtype=np.int64)
        self.names = list(names)
        self.lens = lens
        self.h = L.mm2o_idx_build_mt(n, nm, sq, lens.ctypes.data, w, k, n_threads or (os.cpu_count() or 1))
        self.opt = default_opt()
        L.mm2o_mapopt_update(C.byref(self.opt), self.h)

    @property
    def mid_occ(self) -> int:
        return self.opt.mid_occ

    def __del__(self):
        try:
            if self.h:
                lib().mm2o_idx_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def get(self, minier: int) -> np.ndarray:
        n = C.c_int(0)
        p = lib().mm2o_idx_get(self.h, minier, C.byref(n))
        return np.array([p[i] for i in range(n.value)], dtype=np.uint64)

    @staticmethod
    def _hits(res) -> list[dict]:
        r = res.contents
        out = []
        for i in range(r.n_hits):
            h = r.hits[i]
            d = {f: getattr(h, f) for f in HIT_FIELDS}
            d["cigar"] = np.array([r.cigar_pool[h.cigar_off + j] for j in range(h.n_cigar)], dtype=np.uint32)
            out.append(d)
        return out

    def map(self, seq, trace: bool = False):
        """Return (hits, stats[, trace dict]) for one read -- map.c mm_map_frag through mappy's mm_map."""
        L = lib()
        b = _as_bytes(seq)
        tr = L.mm2o_trace_new() if trace else None
        res = L.mm2o_map(self.h, C.byref(self.opt), b, len(b), tr)
        r = res.contents
        hits = self._hits(res)
        stats = dict(rep_len=r.rep_len, n_mini=r.n_mini, n_anchor=r.n_anchor, chain_cells=r.chain_cells,
                     dp_cells=r.dp_cells, n_dp_calls=r.n_dp_calls)
        L.mm2o_result_destroy(res)
        if not trace:
            return hits, stats
        t = tr.contents

        def v128(v):
            return np.ctypeslib.as_array(v.a, shape=(v.n, 2)).copy() if v.n else np.zeros((0, 2), np.uint64)

        td = dict(mini=v128(t.mini), anchors=v128(t.anchors), chained=v128(t.chained))
        n = t.n_chain_arr
        for nm in ("f", "p", "v"):
            td[nm] = np.ctypeslib.as_array(getattr(t, nm), shape=(n,)).copy() if n else np.zeros(0, np.int32)
        td["u"] = np.ctypeslib.as_array(t.u, shape=(t.n_u,)).copy() if t.n_u else np.zeros(0, np.uint64)
        seqpool = np.ctypeslib.as_array(t.seqpool, shape=(t.n_seq,)).copy() if t.n_seq else np.zeros(0, np.uint8)
        cigpool = np.ctypeslib.as_array(t.cigpool, shape=(t.n_cig,)).copy() if t.n_cig else np.zeros(0, np.uint32)
        dps = []
        for i in range(t.n_dp):
            d = t.dp[i]
            rec = {f: getattr(d, f) for f, _ in DpRec._fields_}
            rec["q"] = seqpool[d.q_off:d.q_off + d.qlen]
            rec["t"] = seqpool[d.t_off:d.t_off + d.tlen]
            rec["cigar"] = cigpool[d.cigar_off:d.cigar_off + d.n_cigar]
            dps.append(rec)
        td["dp"] = dps
        L.mm2o_trace_destroy(tr)
        return hits, stats, td

    def map_batch(self, cat: np.ndarray, off: np.ndarray, n_threads: int = 1):
        """Map a concatenated batch; returns (list of per-read hit lists, stats totals)."""
        L = lib()
        n = len(off) - 1
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        arr = (C.POINTER(Result) * n)()
        L.mm2o_map_batch(self.h, C.byref(self.opt), n, cat.ctypes.data, off.ctypes.data, n_threads, arr)
        out = []
        tot = dict(n_mini=0, n_anchor=0, chain_cells=0, dp_cells=0, n_dp_calls=0)
        for i in range(n):
            out.append(self._hits(arr[i]))
            r = arr[i].contents
            for k in tot:
                tot[k] += getattr(r, k)
            L.mm2o_result_destroy(arr[i])
        return out, tot

    def map_batch_soa(self, cat: np.ndarray, off: np.ndarray, n_threads: int = 1):
        """Map a concatenated batch and return the hits as arrays: dict(fields={name: int32[n_hits]}, hit_off int64[n+1],
        cigar uint32[...], cigar_off int64[n_hits], totals, seconds) -- the checker side of full-batch parity runs.  `seconds` is
        the wall time of the mapping call alone (the CPU-baseline figure)."""
        import time
        L = lib()
        n = len(off) - 1
        cat = np.ascontiguousarray(cat, dtype=np.uint8)
        off = np.ascontiguousarray(off, dtype=np.int64)
        arr = (C.POINTER(Result) * n)()
        t0 = time.perf_counter()
        L.mm2o_map_batch(self.h, C.byref(self.opt), n, cat.ctypes.data, off.ctypes.data, n_threads, arr)
        seconds = time.perf_counter() - t0
        ncw = C.c_int64(0)
        nh = L.mm2o_batch_export(arr, n, None, None, None, C.byref(ncw))
        fields = np.zeros((max(nh, 1), 22), dtype=np.int32)
        hit_off = np.zeros(n + 1, dtype=np.int64)
        cigar = np.zeros(max(ncw.value, 1), dtype=np.uint32)
        L.mm2o_batch_export(arr, n, fields.ctypes.data, hit_off.ctypes.data, cigar.ctypes.data, C.byref(ncw))
        tot = dict(n_mini=0, n_anchor=0, chain_cells=0, dp_cells=0, n_dp_calls=0)
        for i in range(n):
            r = arr[i].contents
            for k in tot:
                tot[k] += getattr(r, k)
            L.mm2o_result_destroy(arr[i])
        fields = fields[:nh]
        names = HIT_FIELDS[:22]
        cols = {nm: np.ascontiguousarray(fields[:, j]) for j, nm in enumerate(names)}
        cigar_off = np.zeros(nh, dtype=np.int64)
        if nh:
            cigar_off[1:] = np.cumsum(cols["n_cigar"][:-1].astype(np.int64))
        return dict(fields=cols, hit_off=hit_off, cigar=cigar[:ncw.value], cigar_off=cigar_off, totals=tot, seconds=seconds)

    def map_batch_raw(self, cat: np.ndarray, off: np.ndarray, n_threads: int = 1):
        """Timing entry: map the batch, return only totals (no Python-side unpacking in the timed region)."""
        L = lib()
        n = len(off) - 1
        arr = (C.POINTER(Result) * n)()
        L.mm2o_map_batch(self.h, C.byref(self.opt), n, cat.ctypes.data, off.ctypes.data, n_threads, arr)
        tot = dict(n_hits=0, n_mini=0, n_anchor=0, chain_cells=0, dp_cells=0, n_dp_calls=0)
        for i in range(n):
            r = arr[i].contents
            for k in tot:
                tot[k] += getattr(r, k)
            L.mm2o_result_destroy(arr[i])
        return tot


def logf_range(first_bits: int, n: int) -> np.ndarray:
    """libm logf (the function mm_set_mapq calls) of the n consecutive float32 bit patterns from first_bits."""
    out = np.empty(n, np.float32)
    lib().mm2o_logf_range(first_bits, n, out.ctypes.data)
    return out
