"""ctypes binding of the C ABI in include/monica_b200.h (monica_b200/lib/libmonica_b200.so).

There is no CPU fallback: if the CUDA library is missing or no device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libmonica_b200.so")


class MonicaB200Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"monica_b200 error {code}: {msg}")
        self.code = code


class Opt(C.Structure):
    _fields_ = [
        ("seed", C.c_int32), ("mid_occ_frac", C.c_float),
        ("min_cnt", C.c_int32), ("min_chain_score", C.c_int32), ("bw", C.c_int32), ("max_gap", C.c_int32),
        ("max_gap_ref", C.c_int32), ("max_chain_skip", C.c_int32), ("max_chain_iter", C.c_int32),
        ("mask_level", C.c_float), ("pri_ratio", C.c_float), ("best_n", C.c_int32),
        ("max_join_long", C.c_int32), ("max_join_short", C.c_int32), ("min_join_flank_sc", C.c_int32),
        ("min_join_flank_ratio", C.c_float),
        ("a", C.c_int32), ("b", C.c_int32), ("q", C.c_int32), ("e", C.c_int32), ("q2", C.c_int32), ("e2", C.c_int32),
        ("sc_ambi", C.c_int32), ("zdrop", C.c_int32), ("zdrop_inv", C.c_int32), ("end_bonus", C.c_int32),
        ("min_dp_max", C.c_int32), ("min_ksw_len", C.c_int32), ("max_clip_ratio", C.c_float),
        ("max_sw_mat", C.c_int64), ("mid_occ", C.c_int32),
    ]


class Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("n_reads", "n_bases", "n_mini", "n_anchor", "n_regs", "n_dp_tasks", "n_dp_pass2",
                                         "dp_cells", "n_hits", "n_rounds")] + \
               [(n, C.c_float) for n in ("ms_sketch", "ms_seed", "ms_sort", "ms_chain", "ms_glue", "ms_dp", "ms_post",
                                         "ms_total", "ms_h2d", "ms_d2h")] + [("n_launches", C.c_int64), ("ms_kdp", C.c_float), ("n_kdp", C.c_int32), ("ms_kdp_fast", C.c_float), ("ms_kdp_exact", C.c_float),
                  ("n_fast_tasks", C.c_int64), ("n_exact_tasks", C.c_int64), ("chain_cells", C.c_int64), ("dp_cells_exact", C.c_int64), ("n_kdp_fast", C.c_int64), ("n_ext_tasks", C.c_int64), ("dp_cells_ext", C.c_int64), ("ms_kdp_ext", C.c_float), ("n_inv", C.c_int32),
                  ("arena_bytes", C.c_int64), ("n_pieces", C.c_int32), ("pad_", C.c_int32),
                  ("n_band_tasks", C.c_int64), ("dp_cells_band", C.c_int64), ("ms_kdp_band", C.c_float), ("pad2_", C.c_int32)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class DpTask(C.Structure):
    _fields_ = [
        ("qlen", C.c_int32), ("tlen", C.c_int32), ("w", C.c_int32), ("zdrop", C.c_int32), ("end_bonus", C.c_int32),
        ("flag", C.c_int32), ("q_off", C.c_int64), ("t_off", C.c_int64),
        ("score", C.c_int32), ("max", C.c_int32), ("max_q", C.c_int32), ("max_t", C.c_int32), ("mqe", C.c_int32),
        ("mqe_t", C.c_int32), ("zdropped", C.c_int32), ("reach_end", C.c_int32), ("n_cigar", C.c_int32),
        ("cigar_off", C.c_int64),
    ]


class LLTask(C.Structure):
    _fields_ = [("qlen", C.c_int32), ("tlen", C.c_int32), ("q_off", C.c_int64), ("t_off", C.c_int64),
                ("score", C.c_int32), ("qe", C.c_int32), ("te", C.c_int32), ("pad_", C.c_int32)]


HIT_FIELDS = ["read_idx", "rid", "rev", "qs", "qe", "rs", "re", "mapq", "mlen", "blen", "nm", "dp_max", "dp_max2",
              "score", "score0", "cnt", "subsc", "n_sub", "id", "parent", "is_primary", "sam_pri", "n_cigar"]

# every symbol include/monica_b200.h declares
SYMBOLS = [
    "mb_last_error", "mb_device_count", "mb_opt_init",
    "mb_index_build", "mb_index_build_fasta", "mb_index_save", "mb_index_load", "mb_index_free", "mb_index_n_seq",
    "mb_index_seq_name", "mb_index_seq_len", "mb_index_mid_occ", "mb_index_kw", "mb_index_n_minimizers", "mb_index_hbm_bytes",
    "mb_map_batch", "mb_map_batch_ex", "mb_reads_pack", "mb_fastq_pack", "mb_packed_free", "mb_packed_upload_bytes", "mb_packed_words", "mb_map_packed", "mb_reads_upload", "mb_reads_free", "mb_map_resident",
    "mb_hits_n", "mb_hits_field", "mb_hits_cigar_off", "mb_hits_cigar_pool", "mb_hits_rep_len", "mb_hits_free",
    "mb_count", "mb_count_last", "mb_count_device_ptr", "mb_count_fetch", "mb_normalize_last",
    "mb_comm_unique_id", "mb_comm_init", "mb_comm_free", "mb_allreduce_counts",
    "mb_sketch", "mb_seed", "mb_chain", "mb_dp_batch", "mb_ll_batch", "mb_int_peak", "mb_stream", "mb_logf_sweep",
    "mb_fastq_load", "mb_fastq_n", "mb_fastq_seqs", "mb_fastq_header", "mb_fastq_ids_unique", "mb_fastq_route", "mb_fastq_route_targets", "mb_fastq_free",
    "mb_db_build",
]

_lib = None


def lib():
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise MonicaB200Error(-2, f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                                  "there is no CPU fallback")
    L = C.CDLL(SO_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    L.mb_last_error.restype = C.c_char_p
    L.mb_device_count.restype = C.c_int
    L.mb_opt_init.argtypes = [C.POINTER(Opt)]
    L.mb_index_build.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_char_p), C.POINTER(vp), vp, C.c_int, C.c_int, C.POINTER(vp)]
    L.mb_index_build_fasta.argtypes = [C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(vp)]
    L.mb_index_save.argtypes = [vp, C.c_char_p]
    L.mb_index_load.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    L.mb_index_free.argtypes = [vp]
    L.mb_index_free.restype = None
    L.mb_index_n_seq.argtypes = [vp]
    L.mb_index_seq_name.argtypes = [vp, C.c_int]
    L.mb_index_seq_name.restype = C.c_char_p
    L.mb_index_seq_len.argtypes = [vp, C.c_int]
    L.mb_index_seq_len.restype = i64
    L.mb_index_mid_occ.argtypes = [vp]
    L.mb_index_kw.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.mb_index_n_minimizers.argtypes = [vp]
    L.mb_index_n_minimizers.restype = i64
    L.mb_index_hbm_bytes.argtypes = [vp]
    L.mb_index_hbm_bytes.restype = i64
    L.mb_map_batch.argtypes = [vp, C.POINTER(Opt), vp, vp, i32, C.POINTER(vp), C.POINTER(Stats)]
    L.mb_map_batch_ex.argtypes = [vp, C.POINTER(Opt), vp, vp, i32, C.c_int, C.POINTER(vp), C.POINTER(Stats)]
    L.mb_reads_pack.argtypes = [vp, vp, i32, C.c_int, C.POINTER(vp)]
    L.mb_fastq_pack.argtypes = [vp, C.c_int, C.POINTER(vp)]
    L.mb_packed_free.argtypes = [vp]
    L.mb_packed_free.restype = None
    L.mb_packed_upload_bytes.argtypes = [vp]
    L.mb_packed_upload_bytes.restype = i64
    L.mb_packed_words.argtypes = [vp, C.POINTER(i64), C.POINTER(vp), C.POINTER(i64)]
    L.mb_packed_words.restype = vp
    L.mb_map_packed.argtypes = [vp, C.POINTER(Opt), vp, C.c_int, C.POINTER(vp), C.POINTER(Stats)]
    L.mb_reads_upload.argtypes = [vp, vp, vp, i32, C.POINTER(vp)]
    L.mb_reads_free.argtypes = [vp]
    L.mb_reads_free.restype = None
    L.mb_map_resident.argtypes = [vp, C.POINTER(Opt), vp, C.c_int, C.POINTER(vp), C.POINTER(Stats)]
    L.mb_hits_n.argtypes = [vp]
    L.mb_hits_n.restype = i64
    L.mb_hits_field.argtypes = [vp, C.c_char_p]
    L.mb_hits_field.restype = C.POINTER(i32)
    L.mb_hits_cigar_off.argtypes = [vp]
    L.mb_hits_cigar_off.restype = C.POINTER(i64)
    L.mb_hits_cigar_pool.argtypes = [vp, C.POINTER(i64)]
    L.mb_hits_cigar_pool.restype = C.POINTER(C.c_uint32)
    L.mb_hits_rep_len.argtypes = [vp, C.POINTER(i64)]
    L.mb_hits_rep_len.restype = C.POINTER(i32)
    L.mb_hits_free.argtypes = [vp]
    L.mb_hits_free.restype = None
    L.mb_count.argtypes = [vp, vp, i32, C.c_int, vp, vp, vp, vp]
    L.mb_count_last.argtypes = [vp, i32, C.c_int, vp, vp]
    L.mb_int_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    L.mb_logf_sweep.argtypes = [C.c_int, C.c_uint32, i64, vp, C.POINTER(i64), C.POINTER(C.c_uint32)]
    L.mb_count_device_ptr.argtypes = [vp]
    L.mb_count_device_ptr.restype = vp
    L.mb_stream.argtypes = [vp]
    L.mb_stream.restype = vp
    L.mb_fastq_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.mb_fastq_n.argtypes = [vp]
    L.mb_fastq_n.restype = C.c_int64
    L.mb_fastq_seqs.argtypes = [vp, C.POINTER(C.POINTER(C.c_int64))]
    L.mb_fastq_seqs.restype = C.POINTER(C.c_uint8)
    L.mb_fastq_header.argtypes = [vp, C.c_int64, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]
    L.mb_fastq_header.restype = C.POINTER(C.c_char)
    L.mb_fastq_ids_unique.argtypes = [vp]
    L.mb_fastq_route.argtypes = [vp, vp, vp, vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    L.mb_fastq_route_targets.argtypes = [vp, vp, vp, vp, i32, vp, C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p]
    L.mb_fastq_free.argtypes = [vp]
    L.mb_fastq_free.restype = None
    L.mb_count_fetch.argtypes = [vp, vp]
    L.mb_normalize_last.argtypes = [vp, vp, i32, vp, vp, i32, vp]
    L.mb_db_build.argtypes = [C.c_char_p, i32, vp, vp, vp]
    L.mb_comm_unique_id.argtypes = [vp]
    L.mb_comm_init.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.POINTER(vp)]
    L.mb_comm_free.argtypes = [vp]
    L.mb_comm_free.restype = None
    L.mb_allreduce_counts.argtypes = [vp, vp, vp, vp]
    L.mb_sketch.argtypes = [C.c_int, vp, vp, i32, C.c_int, C.c_int, vp, i64, vp]
    L.mb_seed.argtypes = [vp, C.POINTER(Opt), vp, vp, i32, vp, i64, vp, vp]
    L.mb_chain.argtypes = [C.c_int, C.POINTER(Opt), vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    L.mb_dp_batch.argtypes = [C.c_int, C.POINTER(Opt), C.POINTER(DpTask), i64, vp, i64, vp, i64]
    L.mb_ll_batch.argtypes = [C.c_int, C.POINTER(Opt), C.POINTER(LLTask), i64, vp, i64]
    _lib = L
    return L


def check(rc: int):
    if rc != 0:
        raise MonicaB200Error(rc, lib().mb_last_error().decode(errors="replace"))


def default_opt() -> Opt:
    o = Opt()
    check(lib().mb_opt_init(C.byref(o)))
    return o


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Hits:
    """Struct-of-arrays view of one mapped batch (copied out of the library, which is then freed)."""

    def __init__(self, h, n_reads: int):
        L = lib()
        n = L.mb_hits_n(h)
        self.n = n
        self.n_reads = n_reads
        self._h = h
        for f in HIT_FIELDS:
            p = L.mb_hits_field(h, f.encode())
            setattr(self, f, np.ctypeslib.as_array(p, shape=(n,)).copy() if n and p else np.zeros(0, np.int32))
        nc = C.c_int64(0)
        cp = L.mb_hits_cigar_pool(h, C.byref(nc))
        self.cigar_pool = np.ctypeslib.as_array(cp, shape=(nc.value,)).copy() if nc.value else np.zeros(0, np.uint32)
        co = L.mb_hits_cigar_off(h)
        self.cigar_off = np.ctypeslib.as_array(co, shape=(n,)).copy() if n and len(self.rid) else np.zeros(0, np.int64)
        nr = C.c_int64(0)
        rp = L.mb_hits_rep_len(h, C.byref(nr))
        self.rep_len = np.ctypeslib.as_array(rp, shape=(nr.value,)).copy() if nr.value else np.zeros(0, np.int32)

    def handle(self):
        return self._h

    def free(self):
        if self._h:
            lib().mb_hits_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def cigar(self, i: int) -> np.ndarray:
        o = int(self.cigar_off[i])
        return self.cigar_pool[o:o + int(self.n_cigar[i])]

    def per_read(self) -> list[list[int]]:
        out = [[] for _ in range(self.n_reads)]
        for i, r in enumerate(self.read_idx):
            out[int(r)].append(i)
        return out
