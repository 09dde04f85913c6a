"""One harness for the C ABI of include/monica_b200.h, whichever library implements it: the CUDA library
(monica_b200/lib/libmonica_b200.so) or the CPU oracle behind the same entry points (oracle/_build/libmonica_b200_oracle.so,
test infrastructure).  SURVEY.md 8(b): "the identical ABI is implemented twice".  Raw ctypes, the calls of INTEGRATION.md."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_SO = os.path.join(ROOT, "monica_b200", "lib", "libmonica_b200.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "libmonica_b200_oracle.so")

# the hot-path subset both libraries export
HOT_PATH_SYMBOLS = [
    "mb_last_error", "mb_device_count", "mb_opt_init",
    "mb_index_build", "mb_index_free", "mb_index_n_seq", "mb_index_seq_name", "mb_index_seq_len", "mb_index_mid_occ", "mb_index_kw",
    "mb_map_batch", "mb_map_batch_ex",
    "mb_hits_n", "mb_hits_field", "mb_hits_cigar_off", "mb_hits_cigar_pool", "mb_hits_rep_len", "mb_hits_free",
    "mb_count", "mb_sketch", "mb_dp_batch", "mb_ll_batch",
]


def load(path):
    """CDLL with the argument types of monica_b200/_lib.py copied onto the symbols this library exports."""
    from monica_b200 import _lib
    ref = _lib.lib()                       # declares every prototype (loading the CUDA library needs no device)
    L = C.CDLL(path)
    for name in HOT_PATH_SYMBOLS:
        f, g = getattr(L, name), getattr(ref, name)
        if g.argtypes is not None:
            f.argtypes = g.argtypes
        f.restype = g.restype
    return L


def oracle_library():
    import subprocess
    if not os.path.exists(ORACLE_SO):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle")], check=True, capture_output=True)
    return load(ORACLE_SO)


def _check(L, rc):
    if rc != 0:
        raise RuntimeError(f"rc={rc}: {L.mb_last_error().decode()}")


def run(L, names, seqs, cat, off, device=0):
    """index build -> map_batch -> every hit array -> mb_count in the three modes (+ mode None): a dict of numpy arrays."""
    from monica_b200 import _lib
    out = {}
    opt = _lib.Opt()
    _check(L, L.mb_opt_init(C.byref(opt)))
    out["opt"] = np.frombuffer(bytes(opt), np.uint8).copy()
    n = len(names)
    c_names = (C.c_char_p * n)(*[s.encode() for s in names])
    arrs = [np.ascontiguousarray(s, dtype=np.uint8) for s in seqs]
    c_seqs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lens = np.array([len(a) for a in arrs], np.int64)
    idx = C.c_void_p()
    _check(L, L.mb_index_build(device, n, c_names, c_seqs, lens.ctypes.data_as(C.c_void_p), 10, 15, C.byref(idx)))
    try:
        k, w = C.c_int(0), C.c_int(0)
        _check(L, L.mb_index_kw(idx, C.byref(k), C.byref(w)))
        out["index"] = np.array([L.mb_index_n_seq(idx), L.mb_index_mid_occ(idx), k.value, w.value] +
                                [L.mb_index_seq_len(idx, i) for i in range(n)], np.int64)
        assert [L.mb_index_seq_name(idx, i).decode() for i in range(n)] == list(names)
        cat = np.ascontiguousarray(cat, np.uint8); off = np.ascontiguousarray(off, np.int64)
        n_reads = len(off) - 1
        h = C.c_void_p()
        _check(L, L.mb_map_batch(idx, C.byref(opt), cat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), n_reads, C.byref(h), None))
        try:
            nh = L.mb_hits_n(h)
            out["n_hits"] = np.array([nh], np.int64)
            for f in _lib.HIT_FIELDS:
                p = L.mb_hits_field(h, f.encode())
                out["hit." + f] = np.ctypeslib.as_array(p, shape=(nh,)).copy() if nh else np.zeros(0, np.int32)
            nc = C.c_int64(0)
            pool = L.mb_hits_cigar_pool(h, C.byref(nc))
            out["cigar"] = np.ctypeslib.as_array(pool, shape=(nc.value,)).copy() if nc.value else np.zeros(0, np.uint32)
            out["cigar_off"] = np.ctypeslib.as_array(L.mb_hits_cigar_off(h), shape=(nh,)).copy() if nh else np.zeros(0, np.int64)
            nr = C.c_int64(0)
            rl = L.mb_hits_rep_len(h, C.byref(nr))
            out["rep_len"] = np.ctypeslib.as_array(rl, shape=(nr.value,)).copy() if nr.value else np.zeros(0, np.int32)
            for mode in (0, 1, 2, -1):
                counts, ncls = np.zeros(n, np.int64), np.zeros(3, np.int64)
                rcls, rbest = np.zeros(n_reads, np.int8), np.zeros(n_reads, np.int64)
                _check(L, L.mb_count(idx, h, 60, mode, counts.ctypes.data_as(C.c_void_p), ncls.ctypes.data_as(C.c_void_p),
                                     rcls.ctypes.data_as(C.c_void_p), rbest.ctypes.data_as(C.c_void_p)))
                out[f"count.{mode}"] = np.concatenate([counts, ncls])
                out[f"class.{mode}"] = rcls
                out[f"best.{mode}"] = rbest
        finally:
            L.mb_hits_free(h)
        # per-stage: minimizers of every read
        cap = int(off[-1]) + 64
        xy = np.zeros((cap, 2), np.uint64); moff = np.zeros(n_reads + 1, np.int64)
        _check(L, L.mb_sketch(device, cat.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), n_reads, 10, 15,
                              xy.ctypes.data_as(C.c_void_p), cap, moff.ctypes.data_as(C.c_void_p)))
        out["sketch_off"] = moff
        out["sketch_xy"] = xy[:moff[-1]].copy()
    finally:
        L.mb_index_free(idx)
    return out
