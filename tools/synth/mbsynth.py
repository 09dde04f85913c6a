"""ctypes wrapper over tools/synth/libmbsynth.so: seeded synthetic genomes / strains / ONT-like reads at BASELINE.json's sizes.
Measurement infrastructure for bench.py and the scale tests (not product code, not part of the oracle)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmbsynth.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mbsynth.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-pthread", src, "-o", _SO, "-lm"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.mbs_genome.argtypes = [C.c_uint64, C.c_int64, C.c_void_p, C.c_int]
        L.mbs_mutate.restype = C.c_int64
        L.mbs_mutate.argtypes = [C.c_uint64, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_int64]
        L.mbs_reads.restype = C.c_int64
        L.mbs_reads.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_double, C.c_int64, C.c_int64,
                                C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        _lib = L
    return _lib


def _threads(threads):
    return threads if threads else max(1, (os.cpu_count() or 4))


STRAIN_LADDER = (0.003, 0.01, 0.02, 0.005, 0.04)


def make_genomes(seed: int, n_genomes: int, genome_len: int, strain_frac: float = 0.2, strain_div=STRAIN_LADDER, threads: int = 0):
    """(names, seqs, gcat, goff): `n_genomes` genomes named 'Species_i:ACCi.1' (monica's database wire format); the last
    round(n * strain_frac) are mutated copies of earlier ones (divergence taken in turn from strain_div, so that close strains
    -- secondaries kept, MAPQ < 60 -- and distant ones both occur whatever the genome count; 80 % sub / 10 % ins / 10 % del).
    seqs are views into one concatenated uint8 buffer gcat with offsets goff."""
    L = lib()
    rng = np.random.default_rng(seed)
    n_strain = int(round(n_genomes * strain_frac)) if n_genomes > 1 else 0
    n_base = n_genomes - n_strain
    cap = n_base * genome_len + int(n_strain * genome_len * 1.02) + 1024
    gcat = np.empty(cap, dtype=np.uint8)
    goff = np.zeros(n_genomes + 1, dtype=np.int64)
    for g in range(n_base):
        L.mbs_genome(seed * 1000003 + g, genome_len, gcat.ctypes.data + int(goff[g]), _threads(threads))
        goff[g + 1] = goff[g] + genome_len
    for g in range(n_base, n_genomes):
        src = int(rng.integers(0, n_base))
        d = float(strain_div[(g - n_base) % len(strain_div)])
        n = L.mbs_mutate(seed * 1000003 + g, gcat.ctypes.data + int(goff[src]), genome_len, d * 0.8, d * 0.1, d * 0.1,
                         gcat.ctypes.data + int(goff[g]), cap - int(goff[g]))
        goff[g + 1] = goff[g] + n
    gcat = gcat[:int(goff[-1])]
    names = [f"Species_{g}:ACC{g:05d}.1" for g in range(n_genomes)]
    seqs = [gcat[int(goff[g]):int(goff[g + 1])] for g in range(n_genomes)]
    return names, seqs, gcat, goff


# fractions of the read classes described in mbsynth.c (junk insertion, inversion, exact chimera, junk)
HARD_MIX = dict(f_junkins=0.015, f_inv=0.005, f_chim=0.025, f_junk=0.01)
PLAIN_MIX = dict(f_junkins=0.0, f_inv=0.0, f_chim=0.0, f_junk=0.0)


def read_lengths(seed: int, gcat: np.ndarray, goff: np.ndarray, n_reads: int, n50: float, error: float = 0.10, mix=(0.4, 0.3, 0.3),
                 sigma: float = 0.6, min_len: int = 500, max_len: int = 0, classes=HARD_MIX, threads: int = 0):
    """Pass 1 only: (off int64[n+1], cls int8[n]) of the whole read set -- what a rank needs to find its shard."""
    L = lib()
    off = np.zeros(n_reads + 1, dtype=np.int64)
    cls = np.zeros(max(1, n_reads), dtype=np.int8)
    sub, ins, dele = (error * m for m in mix)
    L.mbs_reads(C.c_uint64(seed), gcat.ctypes.data, goff.ctypes.data, len(goff) - 1, n_reads, n50, sigma, min_len, max_len, sub, ins, dele,
                classes["f_junkins"], classes["f_inv"], classes["f_chim"], classes["f_junk"], _threads(threads), None, off.ctypes.data, cls.ctypes.data, 0, 0)
    return off, cls[:n_reads]


def simulate_reads(seed: int, gcat: np.ndarray, goff: np.ndarray, n_reads: int, n50: float, error: float = 0.10, mix=(0.4, 0.3, 0.3),
                   sigma: float = 0.6, min_len: int = 500, max_len: int = 0, classes=HARD_MIX, threads: int = 0, out: np.ndarray | None = None,
                   first: int = 0, count: int | None = None, off: np.ndarray | None = None):
    """(cat uint8[total], off int64[count+1] rebased to 0, cls int8[count]) for reads [first, first+count) of the set of
    `n_reads` reads (default: all).  `out`, when given, is a caller-owned buffer (e.g. pinned memory) of sufficient size that
    receives the bases; `off`, when given, is read_lengths()' result for the same arguments."""
    L = lib()
    if count is None:
        count = n_reads - first
    cls = None
    if off is None:
        off, cls = read_lengths(seed, gcat, goff, n_reads, n50, error, mix, sigma, min_len, max_len, classes, threads)
    sub, ins, dele = (error * m for m in mix)
    total = int(off[first + count] - off[first])
    if out is None:
        out = np.empty(total, dtype=np.uint8)
    assert out.nbytes >= total
    L.mbs_reads(C.c_uint64(seed), gcat.ctypes.data, goff.ctypes.data, len(goff) - 1, n_reads, n50, sigma, min_len, max_len, sub, ins, dele,
                classes["f_junkins"], classes["f_inv"], classes["f_chim"], classes["f_junk"], _threads(threads), out.ctypes.data, off.ctypes.data, None, first, count)
    o = np.ascontiguousarray(off[first:first + count + 1] - off[first])
    return out[:total], o, (cls[first:first + count] if cls is not None else None)
