"""GPU parity tests (run on the B200 box: python -m pytest tests -m gpu).  Every check goes through the C ABI
(include/monica_b200.h) and compares bit-exactly with the CPU oracle on the same seeded inputs, and with the committed
golden vectors (tests/golden/: oracle hits; outputs of the unmodified reference aligner.py).  /root/reference is not read.
"""
import ctypes as C
import json
import os
import shutil
import sys
import tempfile

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))

CMP_FIELDS = ["rid", "rev", "qs", "qe", "rs", "re", "mapq", "mlen", "blen", "nm", "dp_max", "dp_max2", "score", "score0", "cnt",
              "subsc", "n_sub", "id", "parent", "is_primary", "sam_pri", "n_cigar"]


@pytest.fixture(scope="module")
def lib():
    from monica_b200 import _lib
    L = _lib.lib()
    assert L.mb_device_count() > 0, "no CUDA device: these tests must run on the GPU box"
    return L


@pytest.fixture(scope="module")
def case(small_case, oracle, lib):
    from monica_b200.mappy_shim import Aligner
    names, seqs, reads = small_case
    al = Aligner(names=names, seqs=seqs, device=0)
    oidx = oracle.Index(names, seqs)
    traces = [oidx.map(r, trace=True) for r in reads]
    return al, oidx, reads, traces


def test_native_library_is_the_one_loaded(lib):
    from monica_b200 import _lib
    assert os.path.samefile(_lib.SO_PATH, os.path.join(ROOT, "monica_b200", "lib", "libmonica_b200.so"))
    with open("/proc/self/maps") as fh:
        assert "libmonica_b200.so" in fh.read()


def test_index_metadata(case):
    al, oidx, reads, _ = case
    assert al.mid_occ == oidx.mid_occ
    assert al.k == 15 and al.w == 10 and al.n_seq == 3


def test_sketch_bit_exact(case, oracle, lib):
    from monica_b200 import _lib, synth
    al, oidx, reads, _ = case
    cat, off = synth.concat_reads(reads)
    cap = len(cat) + 64
    out = np.zeros((cap, 2), dtype=np.uint64)
    ooff = np.zeros(len(reads) + 1, dtype=np.int64)
    _lib.check(lib.mb_sketch(0, _lib._ptr(cat), _lib._ptr(off), len(reads), 10, 15, _lib._ptr(out), cap, _lib._ptr(ooff)))
    for i, r in enumerate(reads):
        got = out[ooff[i]:ooff[i + 1]].copy()
        assert np.all((got[:, 1] >> np.uint64(32)) == i)
        got[:, 1] &= np.uint64(0xffffffff)
        assert np.array_equal(oracle.sketch(r), got), f"read {i}"


def test_sketch_chunk_boundaries_and_long_sequence(oracle, lib):
    """Reads whose lengths straddle the 256-base chunk and 32768-base CTA spans, N runs at chunk edges, and one long contig."""
    from monica_b200 import _lib, synth
    rng = np.random.default_rng(17)
    reads = [synth.random_genome(rng, n) for n in (1, 14, 15, 24, 25, 255, 256, 257, 511, 513, 32767, 32768, 32769, 70001)]
    r = synth.random_genome(rng, 3000); r[250:262] = ord("N"); reads.append(r)
    r = synth.random_genome(rng, 3000); r[300:340] = ord("N"); r[511] = ord("N"); r[512] = ord("N"); reads.append(r)
    reads.append(np.tile(synth.random_genome(rng, 7), 300))            # low complexity: ties
    reads.append(synth.random_genome(rng, 400000))                     # contig-sized
    # mosaic: tandem repeats of several periods (tied hashes inside a window), short homopolymers and single Ns at arbitrary
    # offsets inside long random sequence -- the position-parallel kernel and the automaton must agree chunk by chunk
    parts = []
    for j in range(60):
        parts.append(synth.random_genome(rng, int(rng.integers(150, 900))))
        kind = j % 6
        if kind == 0:
            parts.append(np.tile(synth.random_genome(rng, int(rng.choice([5, 7, 11, 13]))), int(rng.integers(3, 40))))
        elif kind == 1:
            parts.append(np.full(int(rng.integers(16, 60)), rng.choice(list(b"ACGT")), np.uint8))
        elif kind == 2:
            parts.append(np.frombuffer(b"N", np.uint8))
        elif kind == 3:
            u = synth.random_genome(rng, int(rng.integers(16, 30)))
            parts += [u, synth.random_genome(rng, int(rng.integers(0, 9))), u]      # the same k-mers twice within a window
    reads.append(np.concatenate(parts))
    cat, off = synth.concat_reads(reads)
    cap = len(cat) + 64
    out = np.zeros((cap, 2), dtype=np.uint64)
    ooff = np.zeros(len(reads) + 1, dtype=np.int64)
    _lib.check(lib.mb_sketch(0, _lib._ptr(cat), _lib._ptr(off), len(reads), 10, 15, _lib._ptr(out), cap, _lib._ptr(ooff)))
    for i, r in enumerate(reads):
        got = out[ooff[i]:ooff[i + 1]].copy()
        got[:, 1] &= np.uint64(0xffffffff)
        assert np.array_equal(oracle.sketch(r), got), f"read {i} len {len(r)}"
    # homopolymer / dinucleotide runs emit (nearly) one minimizer per base: more than a 256-base chunk's staging row holds,
    # so the batch takes the in-place write pass instead of the staged copy
    reads2 = reads[:6] + [np.full(2000, ord("A"), np.uint8), np.tile(np.frombuffer(b"AC", np.uint8), 700), synth.random_genome(rng, 5000)]
    cat, off = synth.concat_reads(reads2)
    out = np.zeros((len(cat) + 64, 2), dtype=np.uint64)
    ooff = np.zeros(len(reads2) + 1, dtype=np.int64)
    _lib.check(lib.mb_sketch(0, _lib._ptr(cat), _lib._ptr(off), len(reads2), 10, 15, _lib._ptr(out), len(cat) + 64, _lib._ptr(ooff)))
    want = [oracle.sketch(r) for r in reads2]
    assert max(len(w) for w in want[6:8]) > 1000
    for i, r in enumerate(reads2):
        got = out[ooff[i]:ooff[i + 1]].copy()
        got[:, 1] &= np.uint64(0xffffffff)
        assert np.array_equal(want[i], got), f"read {i} len {len(r)}"


@pytest.mark.parametrize("k", [11, 13, 14, 16, 19])
def test_sketch_other_k(oracle, lib, k):
    """Odd k <= 15 goes through the position-parallel kernel (32-bit hash, field width 2k), everything else through the
    automaton for every chunk (even k can stall on a k-mer equal to its reverse complement; k > 15 needs the 64-bit hash)."""
    from monica_b200 import _lib, synth
    rng = np.random.default_rng(100 + k)
    reads = [synth.random_genome(rng, n) for n in (40, 300, 2000, 9000, 33000)]
    r = synth.random_genome(rng, 5000); r[1000] = ord("N"); r[2570:2580] = ord("N"); reads.append(r)
    reads.append(np.tile(np.frombuffer(b"ACGT", np.uint8), 600))        # its even-length k-mers are their own reverse complement
    reads.append(np.concatenate([synth.random_genome(rng, 700), np.tile(synth.random_genome(rng, 9), 60), synth.random_genome(rng, 900)]))
    cat, off = synth.concat_reads(reads)
    cap = len(cat) + 64
    out = np.zeros((cap, 2), dtype=np.uint64)
    ooff = np.zeros(len(reads) + 1, dtype=np.int64)
    _lib.check(lib.mb_sketch(0, _lib._ptr(cat), _lib._ptr(off), len(reads), 10, k, _lib._ptr(out), cap, _lib._ptr(ooff)))
    for i, r in enumerate(reads):
        got = out[ooff[i]:ooff[i + 1]].copy()
        got[:, 1] &= np.uint64(0xffffffff)
        assert np.array_equal(oracle.sketch(r, 10, k), got), f"k {k} read {i} len {len(r)}"


def test_seed_lookup_and_sort_bit_exact(case, lib):
    from monica_b200 import _lib, synth
    al, oidx, reads, traces = case
    cat, off = synth.concat_reads(reads)
    cap = len(cat) * 4 + 1024
    out = np.zeros((cap, 2), dtype=np.uint64)
    ooff = np.zeros(len(reads) + 1, dtype=np.int64)
    rep = np.zeros(len(reads), dtype=np.int32)
    _lib.check(lib.mb_seed(al.handle(), C.byref(al.opt), _lib._ptr(cat), _lib._ptr(off), len(reads), _lib._ptr(out), cap, _lib._ptr(ooff), _lib._ptr(rep)))
    n_tie_reads = 0
    for i, (hits, stats, tr) in enumerate(traces):
        got = out[ooff[i]:ooff[i + 1]]
        assert np.array_equal(tr["anchors"], got), f"read {i}"
        assert stats["rep_len"] == rep[i]
        if len(got) > 1 and np.any(got[1:, 0] == got[:-1, 0]):
            n_tie_reads += 1
    assert n_tie_reads > 0, "the case must contain anchors that tie on x (exercises the exact radix emulation)"


def test_chain_bit_exact(case, lib):
    from monica_b200 import _lib
    al, oidx, reads, traces = case
    anchors = [t[2]["anchors"] for t in traces]
    off = np.zeros(len(anchors) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(x) for x in anchors])
    n_a = int(off[-1])
    cat = np.ascontiguousarray(np.concatenate(anchors), dtype=np.uint64)
    f = np.zeros(n_a + 1, np.int32); p = np.zeros(n_a + 1, np.int32); v = np.zeros(n_a + 1, np.int32)
    ch = np.zeros((n_a + 1, 2), np.uint64); choff = np.zeros(len(anchors) + 1, np.int64)
    u = np.zeros(n_a + 1, np.uint64); uoff = np.zeros(len(anchors) + 1, np.int64)
    _lib.check(lib.mb_chain(0, C.byref(al.opt), _lib._ptr(cat), _lib._ptr(off), len(anchors), _lib._ptr(f), _lib._ptr(p), _lib._ptr(v),
                            _lib._ptr(ch), _lib._ptr(choff), _lib._ptr(u), _lib._ptr(uoff)))
    for i, (_, _, tr) in enumerate(traces):
        s, e = off[i], off[i + 1]
        if e > s:
            assert np.array_equal(tr["f"], f[s:e]) and np.array_equal(tr["p"], p[s:e]) and np.array_equal(tr["v"], v[s:e]), f"read {i}"
        assert np.array_equal(tr["u"], u[uoff[i]:uoff[i + 1]]), f"read {i}"
        assert np.array_equal(tr["chained"], ch[choff[i]:choff[i + 1]]), f"read {i}"


def _run_dp(lib, opt, recs):
    from monica_b200 import _lib
    n = len(recs)
    tasks = (_lib.DpTask * n)()
    pool, po, co = [], 0, 0
    for i, r in enumerate(recs):
        tk = tasks[i]
        tk.qlen, tk.tlen, tk.w, tk.zdrop, tk.end_bonus, tk.flag = r["qlen"], r["tlen"], r["w"], r["zdrop"], r["end_bonus"], r["flag"]
        tk.q_off = po; pool.append(r["q"]); po += r["qlen"]
        tk.t_off = po; pool.append(r["t"]); po += r["tlen"]
        tk.cigar_off = co; co += r["qlen"] + r["tlen"] + 1
    pool = np.ascontiguousarray(np.concatenate(pool), dtype=np.uint8)
    cig = np.zeros(co + 1, dtype=np.uint32)
    _lib.check(lib.mb_dp_batch(0, C.byref(opt), tasks, n, _lib._ptr(pool), len(pool), _lib._ptr(cig), len(cig)))
    return tasks, cig


def _check_dp(tasks, cig, recs):
    bad = []
    for i, r in enumerate(recs):
        tk = tasks[i]
        keys = ["zdropped", "reach_end", "n_cigar", "score"]
        if not (r["flag"] & 0x08):
            keys += ["max", "max_q", "max_t", "mqe", "mqe_t"]
        what = [(k, getattr(tk, k), r[k]) for k in keys if getattr(tk, k) != r[k]]
        if not what and not np.array_equal(cig[tk.cigar_off:tk.cigar_off + tk.n_cigar], r["cigar"]):
            what = [("cigar", None, None)]
        if what:
            bad.append((i, r["qlen"], r["tlen"], r["w"], hex(r["flag"]), what[:3]))
    assert not bad, f"{len(bad)} of {len(recs)} DP tasks differ; first: {bad[:12]}"


def test_dp_tasks_of_the_pipeline_bit_exact(case, lib):
    al, oidx, reads, traces = case
    recs = [d for t in traces for d in t[2]["dp"]]
    assert any(not (r["flag"] & 0x08) and (r["flag"] & 0x40) == 0 for r in recs), "needs second-pass (exact, Z-drop) gap fills"
    assert any(r["zdropped"] for r in recs)
    tasks, cig = _run_dp(lib, al.opt, recs)
    _check_dp(tasks, cig, recs)


def test_dp_random_and_band_limited(oracle, lib):
    """Stand-alone ksw_extd2 problems: tiny, ragged, N-containing, one-sided, and larger than the band (w=751 limits the
    matrix, so the 16-lane block semantics outside the band matter), in all four flag combinations the aligner uses."""
    from monica_b200 import _lib
    rng = np.random.default_rng(23)
    opt = _lib.default_opt()
    recs = []
    shapes = [(1, 1), (1, 40), (40, 1), (15, 17), (16, 16), (33, 200), (200, 33), (257, 255), (900, 1000), (1700, 1500), (2500, 300), (300, 2500)]
    for (ql, tl) in shapes:
        t = rng.integers(0, 4, tl).astype(np.uint8)
        q = t[:ql].copy() if ql <= tl else np.concatenate([t, rng.integers(0, 4, ql - tl).astype(np.uint8)])
        mut = rng.random(len(q)) < 0.12
        q[mut] = rng.integers(0, 4, int(mut.sum()))
        if ql > 20:
            q[7] = 4
        for flag, zdrop, eb in ((0x08, 400, -1), (0x00, 400, -1), (0x40, 400, -1), (0x40 | 0x02 | 0x80, 200, 10)):
            for w in (751, 100 if max(ql, tl) > 300 else 751):
                ez = oracle.ksw_extd2(q, t, w=w, zdrop=zdrop, end_bonus=eb, flag=flag)
                recs.append(dict(qlen=ql, tlen=tl, w=w, zdrop=zdrop, end_bonus=eb, flag=flag, q=q, t=t, score=ez["score"], max=ez["max"],
                                 max_q=ez["max_q"], max_t=ez["max_t"], mqe=ez["mqe"], mqe_t=ez["mqe_t"], zdropped=ez["zdropped"],
                                 reach_end=ez["reach_end"], n_cigar=len(ez["cigar"]), cigar=ez["cigar"]))
    tasks, cig = _run_dp(lib, opt, recs)
    _check_dp(tasks, cig, recs)


def test_dp_fast_path_pairs(oracle, lib):
    """First-pass gap fills (flag KSW_EZ_APPROX_MAX, band not limiting) go through the packed two-tasks-per-warp kernel:
    every column class, ragged pairs (neighbouring tasks of different size share a warp), indel-rich, repeat-rich
    (tie-breaking) and N-containing sequences (which must fall back to the exact kernel), an odd task count."""
    from monica_b200 import _lib
    rng = np.random.default_rng(77)
    opt = _lib.default_opt()
    recs = []

    def mutate(t, err):
        out = []
        for b in t:
            r = rng.random()
            if r < err * 0.4:
                out.append(int(rng.integers(0, 4)))
            elif r < err * 0.7:
                continue
            elif r < err:
                out.extend([int(b), int(rng.integers(0, 4))])
            else:
                out.append(int(b))
        return np.array(out if out else [0], dtype=np.uint8)

    tlens = [1, 2, 31, 33, 97, 128, 129, 160, 161, 200, 224, 225, 230, 256, 257, 288, 300, 320, 352, 353, 384, 400, 448, 449, 512, 600, 640, 641, 700, 768]
    for rep in range(3):
        for tl in tlens:
            kind = rng.integers(0, 4)
            if kind == 3:   # low-complexity: many ties
                t = np.tile(rng.integers(0, 4, 3).astype(np.uint8), tl // 3 + 1)[:tl]
            else:
                t = rng.integers(0, 4, tl).astype(np.uint8)
            q = mutate(t, (0.05, 0.12, 0.3, 0.15)[kind])
            if rep == 2 and tl % 7 == 0 and len(q) > 9:
                q[len(q) // 2] = 4
            if rep == 1 and tl % 11 == 0:
                t = t.copy(); t[tl // 3] = 4
            if len(q) > 1000 or max(len(q), tl) > 752:
                q = q[:750]
            if max(len(q), tl) > 752:
                continue
            variants = [(0x08, 400, -1)]
            # end extensions (k_dp_ext): right and left (gaps right-aligned, reversed CIGAR), Z-drop 400 / 200, end bonus;
            # a junk tail on the query makes the score collapse so that the Z-drop break and max tracking matter
            if rep != 1 or tl % 11:
                qx = q
                if kind in (1, 2) and len(q) + 60 <= 752:
                    qx = np.concatenate([q, rng.integers(0, 4, int(rng.integers(20, 60))).astype(np.uint8)])
                if kind == 0 and len(q) > 120:
                    qx = np.concatenate([q[:len(q) // 2], rng.integers(0, 4, len(q) - len(q) // 2).astype(np.uint8)])
                for flag, zd, eb in ((0x40, 400, -1), (0x40 | 0x02 | 0x80, 400, -1), (0x40, 200, 10), (0x40 | 0x02 | 0x80, 100, 5)):
                    ez = oracle.ksw_extd2(qx, t, w=751, zdrop=zd, end_bonus=eb, flag=flag)
                    recs.append(dict(qlen=len(qx), tlen=tl, w=751, zdrop=zd, end_bonus=eb, flag=flag, q=qx, t=t, score=ez["score"], max=ez["max"],
                                     max_q=ez["max_q"], max_t=ez["max_t"], mqe=ez["mqe"], mqe_t=ez["mqe_t"], zdropped=ez["zdropped"],
                                     reach_end=ez["reach_end"], n_cigar=len(ez["cigar"]), cigar=ez["cigar"]))
            for flag, zd, eb in variants:
                ez = oracle.ksw_extd2(q, t, w=751, zdrop=zd, end_bonus=eb, flag=flag)
                recs.append(dict(qlen=len(q), tlen=tl, w=751, zdrop=zd, end_bonus=eb, flag=flag, q=q, t=t, score=ez["score"], max=ez["max"],
                                 max_q=ez["max_q"], max_t=ez["max_t"], mqe=ez["mqe"], mqe_t=ez["mqe_t"], zdropped=ez["zdropped"],
                                 reach_end=ez["reach_end"], n_cigar=len(ez["cigar"]), cigar=ez["cigar"]))
    if len(recs) % 2 == 0:
        recs.pop()
    assert sum(1 for r in recs if r["flag"] & 0x40 and r["zdropped"]) > 5 and sum(1 for r in recs if r["flag"] & 0x40 and not r["zdropped"]) > 5
    assert any(r["reach_end"] for r in recs)
    tasks, cig = _run_dp(lib, opt, recs)
    _check_dp(tasks, cig, recs)
    # the chained variant of the kernel (dp_fast_chain.cuh, an opt-in experiment: the pairs of a warp back to back through the
    # lanes) must give the same answers: groups of 2, 3 and 4 pairs, every class, ragged group at the end of each list
    for g in ("2", "3", "4"):
        os.environ["MB_FAST_CHAIN"], os.environ["MB_FAST_CHAIN_FORCE"] = g, "1"
        try:
            tasks, cig = _run_dp(lib, opt, recs + recs[:-1] + recs[3:])
        finally:
            del os.environ["MB_FAST_CHAIN"], os.environ["MB_FAST_CHAIN_FORCE"]
        _check_dp(tasks, cig, recs + recs[:-1] + recs[3:])


def _compare_hits(hits, per, i, want):
    got = per[i]
    assert len(want) == len(got), f"read {i}: oracle {len(want)} hits, GPU {len(got)}"
    for w, g in zip(want, got):
        for f in CMP_FIELDS:
            assert int(getattr(hits, f)[g]) == int(w[f]), f"read {i} field {f}: GPU {int(getattr(hits, f)[g])} oracle {w[f]}"
        assert np.array_equal(hits.cigar(g), w["cigar"]), f"read {i}: CIGAR"


def test_full_pipeline_bit_exact_vs_oracle(case):
    al, oidx, reads, traces = case
    hits = al.map_batch(reads)
    per = hits.per_read()
    for i in range(len(reads)):
        _compare_hits(hits, per, i, traces[i][0])
    assert al.last_stats["n_rounds"] >= 2 and al.last_stats["n_dp_pass2"] > 0     # Z-drop splits were exercised
    assert np.array_equal(hits.rep_len, np.array([t[1]["rep_len"] for t in traces], dtype=np.int32))
    # one read at a time gives the same answer as the batch (the mappy-shaped .map())
    for i in (3, 4, 15, 20):
        one = list(al.map(reads[i].tobytes()))
        assert [(a.rid, a.r_st, a.r_en, a.q_st, a.q_en, a.strand, a.mapq, a.mlen, a.NM, a.is_primary) for a in one] == \
               [(w["rid"], w["rs"], w["re"], w["qs"], w["qe"], -1 if w["rev"] else 1, w["mapq"], w["mlen"], w["nm"], bool(w["is_primary"])) for w in traces[i][0]]


def test_full_pipeline_equals_golden_vectors(lib):
    from monica_b200.mappy_shim import Aligner
    g = np.load(os.path.join(GOLDEN, "small_case.npz"), allow_pickle=False)
    names = g["names"].tolist()
    seqs = [g["genome_cat"][g["genome_off"][i]:g["genome_off"][i + 1]] for i in range(len(names))]
    al = Aligner(names=names, seqs=seqs, device=0)
    assert al.mid_occ == int(g["mid_occ"])
    hits = al.map_batch(cat=g["read_cat"], off=g["read_off"])
    fields = g["hit_fields"].tolist()
    assert hits.n == len(g["hits"])
    assert np.array_equal(np.bincount(hits.read_idx, minlength=len(g["read_off"]) - 1), np.diff(g["hit_off"]))
    for j, f in enumerate(fields):
        assert np.array_equal(getattr(hits, f), g["hits"][:, j]), f
    assert np.array_equal(hits.cigar_pool, g["cigar"])


def test_batch_order_invariance_and_threads(case):
    """Results do not depend on batch composition, and concurrent callers (monica's ThreadPool) get the same answers."""
    from multiprocessing.dummy import Pool
    al, oidx, reads, traces = case
    order = np.random.default_rng(1).permutation(len(reads))

    def run(idx):
        h = al.map_batch([reads[i] for i in idx])
        per = h.per_read()
        for k, i in enumerate(idx):
            _compare_hits(h, per, k, traces[i][0])
        return True
    assert run(order)
    with Pool(3) as pool:
        assert all(pool.map(run, [order[:20], order[20:35], order[35:]]))


def test_count_modes_equal_python_best_hit(case):
    from monica_b200 import aligner as mine
    al, oidx, reads, traces = case
    hits = al.map_batch(reads)
    names = al.seq_names
    for mode in ("basic", "query_length", "matching", None):
        counts, ncls, rcls, rbest = al.count(hits, 60, mode)
        want = np.zeros(al.n_seq, dtype=np.int64)
        cls = []
        for i, r in enumerate(reads):
            kept = [(h["rid"], h["nm"], h["mlen"]) for h in traces[i][0] if h["is_primary"] and h["mapq"] >= 60]
            if not kept:
                cls.append(0); continue
            best = kept[0] if len(kept) == 1 else mine.best_hit(kept)
            if not best:
                cls.append(2); continue
            cls.append(1)
            want[best[0]] += {"basic": 1, "query_length": len(r), "matching": best[2]}.get(mode, 0)
        assert np.array_equal(counts, want), mode
        assert rcls.tolist() == cls
        assert ncls.tolist() == [cls.count(1), cls.count(0), cls.count(2)]


def test_device_normalizer_equals_python_normalizer(case):
    """BPB/BPM from the device-resident count vector == normalizer() (aligner.py:305-319), float for float."""
    import copy
    from collections import Counter
    from monica_b200 import aligner as mine
    al, oidx, reads, traces = case
    hits = al.map_batch(reads)
    names = al.seq_names
    rng = np.random.default_rng(3)
    for mode in ("basic", "query_length", "matching"):
        counts, ncls, rcls, rbest = al.count(hits, 60, mode)
        # the alignment dict in the order aligner() would have filled it: first mapped read of each target
        sample = {}
        for i in range(len(reads)):
            if rcls[i] != 1:
                continue
            tax_unit, accession = names[hits.rid[int(rbest[i])]].split(':')[0:2]
            sample.setdefault(tax_unit, Counter())
            sample[tax_unit].setdefault(accession, 0)
        for nm, c in zip(names, counts):
            tax_unit, accession = nm.split(':')[0:2]
            if c:
                sample[tax_unit][accession] += int(c)
        assert sample and all(v > 0 for c in sample.values() for v in c.values())
        glen = {acc: int(rng.integers(1000, 7_000_000)) for c in sample.values() for acc in c}
        want = mine.normalizer({"s": copy.deepcopy(sample)}, genomes_length=glen)["s"]
        got = al.normalize_last(sample, glen)
        assert {k: dict(v) for k, v in got.items()} == {k: dict(v) for k, v in want.items()}, mode
        assert list(got) == list(want)


def test_database_builder_output_indexes_and_maps(tmp_path):
    """a database written by mb_db_build (multi-contig genomes re-headed tax_unit:accession) goes through
    indexer-style index build and maps reads to the renamed contigs."""
    import gzip
    from monica_b200 import database, synth
    from monica_b200.mappy_shim import Aligner
    rng = np.random.default_rng(11)
    genomes = []
    seqs = {}
    for g in range(3):
        path = tmp_path / f"GCF_{g}.fna.gz"
        contigs = [synth.random_genome(rng, 30_000), synth.random_genome(rng, 20_000)]
        with gzip.open(path, "wt") as fh:
            for ci, c in enumerate(contigs):
                fh.write(f">NZ_{g}{ci}.1 organism {g} contig {ci}\n")
                text = c.tobytes().decode()
                for i in range(0, len(text), 80):
                    fh.write(text[i:i + 80] + "\n")
        genomes.append((str(path), (f"Species_{g}", f"GCF_{g}.1")))
        seqs[g] = contigs
    lengths = database.builder(genomes, str(tmp_path), database.DATABASE_NAME, 1)
    assert lengths == {f"GCF_{g}.1": 50_000 for g in range(3)}
    al = Aligner(fn_idx_in=str(tmp_path / "database1.fna.gz"), preset="map-ont", best_n=15)
    assert al and al.seq_names == [f"Species_{g}:GCF_{g}.1" for g in range(3) for _ in range(2)]
    read = seqs[1][1][2_000:9_000].tobytes().decode()
    hit = [h for h in al.map(read) if h.is_primary][0]
    assert hit.ctg == "Species_1:GCF_1.1" and hit.mapq == 60 and hit.r_st == 2_000 and hit.r_en == 9_000


def test_index_save_load_roundtrip(case, tmp_path):
    from monica_b200.mappy_shim import Aligner
    from monica_b200 import synth
    al, oidx, reads, traces = case
    names, seqs = al.seq_names, None
    fa = str(tmp_path / "database1.fna.gz")
    g = np.load(os.path.join(GOLDEN, "small_case.npz"), allow_pickle=False)
    gseqs = [g["genome_cat"][g["genome_off"][i]:g["genome_off"][i + 1]] for i in range(len(names))]
    synth.write_fasta_gz(fa, g["names"].tolist(), gseqs)
    mmi = str(tmp_path / "index1.mmi")
    a1 = Aligner(fn_idx_in=fa, preset="map-ont", best_n=15, fn_idx_out=mmi)
    assert a1 and os.path.getsize(mmi) > 0
    a2 = Aligner(fn_idx_in=mmi)
    assert a2 and a2.seq_names == a1.seq_names and a2.mid_occ == a1.mid_occ == int(g["mid_occ"])
    h1 = a1.map_batch(cat=g["read_cat"], off=g["read_off"])
    h2 = a2.map_batch(cat=g["read_cat"], off=g["read_off"])
    for f in CMP_FIELDS:
        assert np.array_equal(getattr(h1, f), getattr(h2, f))
        assert np.array_equal(getattr(h1, f), g["hits"][:, g["hit_fields"].tolist().index(f)])
    assert not Aligner(fn_idx_in=str(tmp_path / "missing.mmi"))                 # falsy, like mappy
    bad = tmp_path / "bad.mmi"; bad.write_bytes(b"MMI\x02" + b"\x00" * 7)
    assert not Aligner(fn_idx_in=str(bad))


@pytest.mark.parametrize("key", ["mode=basic,two_indexes=False", "mode=query_length,two_indexes=True", "mode=matching,two_indexes=True", "mode=None,two_indexes=False"])
def test_aligner_end_to_end_equals_reference_golden(small_case, key):
    """monica_b200.aligner on the GPU == the UNMODIFIED reference aligner.py over the oracle (tests/golden/ref_aligner.json):
    same alignment dict, same routed FASTQ bytes, inputs consumed, hits pickles removed."""
    sys.path.insert(0, GOLDEN)
    import make_golden
    from monica_b200 import aligner as mine
    names, seqs, reads = small_case
    want = json.load(open(os.path.join(GOLDEN, "ref_aligner.json")))[key]
    mode = key.split(",")[0].split("=")[1]
    mode = None if mode == "None" else mode
    two = key.endswith("True")
    wd = tempfile.mkdtemp(prefix="gpu_e2e_")
    try:
        got = make_golden.run_aligner(mine, wd, names, seqs, reads, mode, two, ["Species_1"] if two else [],
                                      {"indexer": {"genomes_path": os.path.join(wd, "markers")}})
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    assert json.loads(json.dumps(got)) == want


def test_larger_random_batch_bit_exact(oracle, lib):
    """400 reads (2 kb - 30 kb, 10-15 % error, both strands, junk) against 4 genomes incl. a strain copy."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(31, 4, 150000, strain_frac=0.25)
    r1, _ = synth.simulate_reads(32, seqs, 250, 5000, 0.10, junk_frac=0.03)
    r2, _ = synth.simulate_reads(33, seqs, 150, 9000, 0.15, sigma=0.8)
    reads = r1 + r2
    al = Aligner(names=names, seqs=seqs, device=0)
    oidx = oracle.Index(names, seqs)
    cat, off = synth.concat_reads(reads)
    want, _ = oidx.map_batch(cat, off, n_threads=os.cpu_count() or 4)
    hits = al.map_batch(cat=cat, off=off)
    per = hits.per_read()
    for i in range(len(reads)):
        _compare_hits(hits, per, i, want[i])


def test_round_trip_properties_at_scale(lib):
    """Size-independent properties on a bench-shaped batch (20k reads): every CIGAR consumes exactly its query / reference
    interval, NM = blen - mlen, mapped reads land on their source contig, counts sum to the mapped reads' bases, and a
    second run of the same batch is identical."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(41, 4, 1_000_000, strain_frac=0.0)
    cat, off = synth.simulate_reads_bulk(42, seqs, 20000, 8000, 0.10)
    al = Aligner(names=names, seqs=seqs, device=0)
    hits = al.map_batch(cat=cat, off=off)
    assert hits.n > 0.98 * (len(off) - 1)
    ops = hits.cigar_pool & np.uint32(0xf)
    lens = (hits.cigar_pool >> np.uint32(4)).astype(np.int64)
    owner = np.repeat(np.arange(hits.n), hits.n_cigar)
    qlen = np.bincount(owner, weights=lens * np.isin(ops, (0, 1)), minlength=hits.n).astype(np.int64)
    tlen = np.bincount(owner, weights=lens * np.isin(ops, (0, 2)), minlength=hits.n).astype(np.int64)
    assert np.array_equal(qlen, (hits.qe - hits.qs).astype(np.int64))
    assert np.array_equal(tlen, (hits.re - hits.rs).astype(np.int64))
    assert np.array_equal(hits.nm, hits.blen - hits.mlen)
    assert np.all((hits.mapq >= 0) & (hits.mapq <= 60)) and np.all(hits.mlen <= hits.blen)
    rl = np.diff(off)
    assert np.all(hits.qe <= rl[hits.read_idx]) and np.all(hits.qs >= 0)
    counts, ncls, rcls, rbest = al.count(hits, 60, "query_length")
    assert counts.sum() == rl[rcls == 1].sum()
    assert ncls.sum() == len(rl) and ncls[0] > 0.97 * len(rl)
    again = al.map_batch(cat=cat, off=off)
    for f in CMP_FIELDS:
        assert np.array_equal(getattr(hits, f), getattr(again, f)), f
    assert np.array_equal(hits.cigar_pool, again.cigar_pool)


def test_sub_batch_pieces_do_not_change_results(case, monkeypatch):
    """A batch is cut into concurrent pieces (own stream / arena / host thread each); 1, 2, 3 and 4 pieces must give
    identical hit arrays, CIGARs and counts (small batch, size gate lowered)."""
    from monica_b200 import synth
    al, oidx, reads, _ = case
    cat, off = synth.concat_reads(reads + reads[::-1] + reads)
    monkeypatch.setenv("MB_PARTS", "1")
    base = al.map_batch(cat=cat, off=off)
    base_counts = al.count(base, 60, "matching")
    for k in ("2", "3", "4"):
        monkeypatch.setenv("MB_PARTS", k)
        monkeypatch.setenv("MB_PARTS_MIN_READS", "4")
        got = al.map_batch(cat=cat, off=off)
        assert got.n == base.n
        for f in ["read_idx"] + CMP_FIELDS:
            assert np.array_equal(getattr(got, f), getattr(base, f)), (k, f)
        assert np.array_equal(got.cigar_pool, base.cigar_pool) and np.array_equal(got.cigar_off, base.cigar_off), k
        assert np.array_equal(got.rep_len, base.rep_len), k
        c2 = al.count(got, 60, "matching")
        for a, b in zip(base_counts, c2):
            assert np.array_equal(a, b), k
        from monica_b200 import _lib
        counts = np.zeros(al.n_seq, np.int64); ncls = np.zeros(3, np.int64)
        _lib.check(_lib.lib().mb_count_last(al.handle(), 60, 2, _lib._ptr(counts), _lib._ptr(ncls)))
        assert np.array_equal(counts, base_counts[0]) and np.array_equal(ncls, base_counts[1]), k


def test_piecewise_upload_path_equals_plain_path(lib, monkeypatch):
    """mb_map_batch uploads big batches in pieces underneath the sketch kernels (SketchFeed): chunks at a piece edge are
    sketched by the automaton after the last piece has landed.  Forced here on a small batch (the threshold is 64 MB)."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(5, 3, 200_000)
    reads, _ = synth.simulate_reads(6, seqs, 1500, 3000.0, 0.10)
    cat, off = synth.concat_reads(reads)
    assert off[-1] > 9 * 32768 and (off[-1] // 32768) % 8 != 0           # several CTA spans per piece, ragged last piece
    al = Aligner(names=names, seqs=seqs, preset="map-ont", best_n=15)
    assert al
    plain = al.map_batch(cat=cat, off=off)
    monkeypatch.setenv("MB_FEED_MIN_BYTES", "1")
    fed = al.map_batch(cat=cat, off=off)
    monkeypatch.delenv("MB_FEED_MIN_BYTES")
    assert plain.n == fed.n and plain.n > 1000
    for f in CMP_FIELDS:
        assert np.array_equal(getattr(plain, f), getattr(fed, f)), f
    assert np.array_equal(plain.cigar_off, fed.cigar_off) and np.array_equal(plain.cigar_pool, fed.cigar_pool)


def test_long_noisy_reads_bit_exact(oracle, lib):
    """BASELINE config 5 in miniature: 50 kb reads at 15 % error against several genomes incl. a strain copy, plus chimeric
    reads (two fragments glued: region splits, long extensions, Z-drops).  Compared hit for hit with the oracle."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(51, 4, 400000, strain_frac=0.25)
    r1, _ = synth.simulate_reads(52, seqs, 10, 50000, 0.15, fixed_len=50000)
    r2, _ = synth.simulate_reads(53, seqs, 8, 20000, 0.15, sigma=0.9)
    r3, _ = synth.simulate_reads(54, seqs, 8, 6000, 0.12, fixed_len=6000)
    chim = [np.concatenate([r3[i], r3[i + 1]]) for i in range(0, 8, 2)]              # two loci in one read: two primary hits
    chim += [np.concatenate([r3[0], r3[1][::-1].copy()])]                              # reversed (not complemented) tail: junk
    junk_tail = [np.concatenate([r, synth.random_genome(np.random.default_rng(60 + i), 1500)]) for i, r in enumerate(r3[:3])]
    reads = r1 + r2 + chim + junk_tail
    al = Aligner(names=names, seqs=seqs, device=0)
    oidx = oracle.Index(names, seqs)
    cat, off = synth.concat_reads(reads)
    want, _ = oidx.map_batch(cat, off, n_threads=os.cpu_count() or 4)
    hits = al.map_batch(cat=cat, off=off)
    per = hits.per_read()
    assert sum(len(w) for w in want) >= len(reads) + 3
    for i in range(len(reads)):
        _compare_hits(hits, per, i, want[i])


def test_streaming_batches_accumulate_to_one_shot_counts(lib):
    """BASELINE config 4 in miniature: 4,000-read batches mapped and counted incrementally; the running per-target counts
    (all three of monica's modes) equal the counts of the whole run mapped at once."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(71, 5, 300000, strain_frac=0.4)
    cat, off = synth.simulate_reads_bulk(72, seqs, 12000, 3000, 0.10)
    al = Aligner(names=names, seqs=seqs, device=0)
    whole = al.map_batch(cat=cat, off=off)
    for mode in ("basic", "query_length", "matching"):
        want_counts, want_cls, _, _ = al.count(whole, 60, mode)
        run_counts = np.zeros_like(want_counts); run_cls = np.zeros_like(want_cls)
        for lo in range(0, 12000, 4000):
            o = off[lo:lo + 4001] - off[lo]
            h = al.map_batch(cat=cat[off[lo]:off[lo + 4000]], off=o)
            c, k, _, _ = al.count(h, 60, mode)
            run_counts += c; run_cls += k
        assert np.array_equal(run_counts, want_counts) and np.array_equal(run_cls, want_cls), mode
    assert want_cls[0] > 0.5 * 12000 and want_cls.sum() == 12000   # strain copies push many reads below MAPQ 60


def test_local_alignment_kernel_bit_exact(oracle, lib):
    """mb_ll_batch (ll.cuh) vs the oracle's lane-by-lane restatement of ksw_ll_i16: score, query end, target end -- including
    the padded-column quirk (best alignment ending at the query end: qe >= qlen, te carried forward) and last-row / last-slot
    tie rules."""
    from monica_b200 import _lib
    rng = np.random.default_rng(91)
    qs, ts = [], []
    for it in range(260):
        ql, tl = int(rng.integers(1, 400)), int(rng.integers(1, 400))
        q = rng.integers(0, 4, ql).astype(np.uint8)
        t = rng.integers(0, 4, tl).astype(np.uint8)
        kind = it % 5
        if kind == 0 and ql > 30 and tl > 40:          # planted noisy copy somewhere in the middle
            n = min(ql, tl) // 2
            a, b = int(rng.integers(0, ql - n + 1)), int(rng.integers(0, tl - n + 1))
            seg = q[a:a + n].copy(); m = rng.random(n) < 0.08; seg[m] = (seg[m] + 1) % 4
            t[b:b + n] = seg
        elif kind == 1 and tl > ql + 8:                # whole query inside the target: ends at the last query column
            b = int(rng.integers(0, tl - ql - 7)); t[b:b + ql] = q
        elif kind == 2:                                # low complexity: many equal maxima
            q[:] = q[0]; t[:] = q[0]
        elif kind == 3 and ql > 8:
            q[rng.integers(0, ql, 3)] = 4              # ambiguous bases
        qs.append(q); ts.append(t)
    qs.append(rng.integers(0, 4, 4990).astype(np.uint8)); ts.append(np.concatenate([rng.integers(0, 4, 50).astype(np.uint8), qs[-1][100:4000], rng.integers(0, 4, 60).astype(np.uint8)]))
    pool = np.concatenate([np.concatenate([q, t]) for q, t in zip(qs, ts)])
    tasks = (_lib.LLTask * len(qs))()
    o = 0
    for i, (q, t) in enumerate(zip(qs, ts)):
        tasks[i].qlen, tasks[i].tlen, tasks[i].q_off, tasks[i].t_off = len(q), len(t), o, o + len(q)
        o += len(q) + len(t)
    opt = _lib.default_opt()
    _lib.check(lib.mb_ll_batch(0, C.byref(opt), tasks, len(qs), _lib._ptr(pool), len(pool)))
    n_pad = 0
    for i, (q, t) in enumerate(zip(qs, ts)):
        want = oracle.ksw_ll_i16(q, t)
        assert (tasks[i].score, tasks[i].qe, tasks[i].te) == want, f"task {i} ({len(q)} x {len(t)}): GPU {(tasks[i].score, tasks[i].qe, tasks[i].te)} oracle {want}"
        n_pad += want[1] >= len(q)
    assert n_pad > 10        # the padded-column case really occurred


def _inversion_reads(seed, g, n):
    """Reads with one or two internal inversions (reverse-complemented blocks of 600-1500 bp)."""
    from monica_b200 import synth
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.integers(5000, 9000))
        st = int(rng.integers(0, len(g) - L))
        r = g[st:st + L].copy()
        a = int(rng.integers(1500, 2500)); b = a + int(rng.integers(600, 1500))
        r[a:b] = synth.revcomp(r[a:b])
        if i % 3 == 0 and L > b + 3500:
            a2 = b + int(rng.integers(1500, 2200)); b2 = a2 + int(rng.integers(600, 1200))
            r[a2:b2] = synth.revcomp(r[a2:b2])
        r = synth.mutate(rng, r, 0.03, 0.02, 0.02)
        if i % 2:
            r = synth.revcomp(r)
        out.append(r)
    return out


def test_inversion_hits_and_region_pool_growth(oracle, lib, monkeypatch):
    """mm_align1_inv: reads with internal inversions get the extra `inv` hit upstream produces between the two halves of a
    split_inv Z-drop split (MAPQ from mm_set_inv_mapq).  Run once normally and once with no slack in the region pool
    (MB_TEST_TIGHT_REGS), which forces the pool to grow before the first alignment round and before the inversion pass."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(51, 2, 120000, strain_frac=0.0)
    reads = _inversion_reads(52, seqs[0], 24) + synth.simulate_reads(53, seqs, 12, 3000, 0.10)[0]
    al = Aligner(names=names, seqs=seqs, device=0)
    oidx = oracle.Index(names, seqs)
    cat, off = synth.concat_reads(reads)
    want, _ = oidx.map_batch(cat, off, n_threads=os.cpu_count() or 4)
    n_inv = sum(1 for hs in want for h in hs if h["cnt"] == 0)
    assert n_inv >= 10, n_inv
    for tight in (False, True):
        if tight:
            monkeypatch.setenv("MB_TEST_TIGHT_REGS", "1")
        hits = al.map_batch(cat=cat, off=off)
        per = hits.per_read()
        for i in range(len(reads)):
            _compare_hits(hits, per, i, want[i])
        assert al.last_stats["n_inv"] >= n_inv
    # an inversion hit is primary with the MAPQ of its neighbours: monica sees it as a second kept hit of the read
    assert any(h["cnt"] == 0 and h["is_primary"] and h["mapq"] == 60 for hs in want for h in hs)


def test_more_than_64_tied_chains(oracle, lib):
    """Reads made of 70-90 copies of one reference segment: every copy is its own chain and all chains start on the same
    reference position, so upstream's sorts of the chain list leave the insertion-sort range with tied keys (the unstable
    radix permutation is part of the result).  Such reads take the `big` variants of the per-read kernels."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(61, 2, 80000, strain_frac=0.0)
    g = seqs[0]
    rng = np.random.default_rng(62)
    reads = [np.tile(g[5000:5150], 70), np.tile(g[9000:9200], 90), synth.revcomp(np.tile(g[12000:12160], 75)),
             np.concatenate([np.tile(g[20000:20150], 40), np.tile(g[30000:30150], 45)]),
             np.concatenate([g[int(s):int(s) + 160] for s in rng.integers(0, 70000, 80)])]        # 80 different loci: distinct keys, > 64 chains
    reads += synth.simulate_reads(63, seqs, 6, 3000, 0.10)[0]
    al = Aligner(names=names, seqs=seqs, device=0)
    oidx = oracle.Index(names, seqs)
    cat, off = synth.concat_reads(reads)
    want, _ = oidx.map_batch(cat, off, n_threads=os.cpu_count() or 4)
    assert max(len(w) for w in want) >= 5
    hits = al.map_batch(cat=cat, off=off)
    per = hits.per_read()
    for i in range(len(reads)):
        _compare_hits(hits, per, i, want[i])


def test_thousands_of_zdrop_candidates_in_one_batch(oracle, lib):
    """2,400 reads that each carry a 500-650 bp block of unrelated sequence: every one of them sends a gap fill through the
    Z-drop walk with a drop above zdrop_inv, i.e. through the inversion test (k_ztest_ll), and then through the exact second
    pass and a Z-drop split.  A fixed-size scratch pool used to abort the whole batch at the 2,049th such task."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(71, 2, 400000, strain_frac=0.0)
    rng = np.random.default_rng(72)
    g = seqs[0]
    reads = []
    for i in range(2400):
        st = int(rng.integers(0, len(g) - 3100))
        r = g[st:st + 3000].copy()
        n = int(rng.integers(500, 650))
        r[1200:1200 + n] = synth.random_genome(rng, n)
        reads.append(synth.mutate(rng, r, 0.02, 0.01, 0.01))
    al = Aligner(names=names, seqs=seqs, device=0)
    oidx = oracle.Index(names, seqs)
    cat, off = synth.concat_reads(reads)
    want, _ = oidx.map_batch(cat, off, n_threads=os.cpu_count() or 4)
    hits = al.map_batch(cat=cat, off=off)
    per = hits.per_read()
    for i in range(len(reads)):
        _compare_hits(hits, per, i, want[i])
    assert al.last_stats["n_dp_pass2"] > 2048


def test_dp_band_kernel_large_and_band_limited_fills(oracle, lib):
    """First-pass gap fills (KSW_EZ_APPROX_MAX) on windows larger than the band or than k_dp_fast's 768-column limit go
    through k_dp_band, which has to reproduce what ksw_extd2_sse computes at and beyond the band edge (16-aligned block
    ranges, initial-value neighbours, stale substitution scores).  Shapes: up to 2048 x 2048, |tlen - qlen| from 0 to close
    to w, w = 751 and narrower / wider bands (long-join windows use w = max length); content: similar sequences, unrelated
    sequences (the optimum wanders to the band edge), similar-after-a-big-offset (the path runs along the band edge), low
    complexity (ties).  Pairs of different sizes share a warp."""
    from monica_b200 import _lib
    rng = np.random.default_rng(123)
    opt = _lib.default_opt()
    recs = []

    def noisy(t, err):
        keep = rng.random(len(t)) >= err * 0.3
        q = t[keep].copy()
        m = rng.random(len(q)) < err * 0.4
        q[m] = rng.integers(0, 4, int(m.sum()))
        ins = np.nonzero(rng.random(len(q)) < err * 0.3)[0]
        return np.insert(q, ins, rng.integers(0, 4, len(ins))).astype(np.uint8)

    shapes = [(760, 760), (753, 900), (900, 753), (1000, 1000), (1024, 1024), (1025, 1030), (1200, 800), (800, 1200), (1500, 1500), (1536, 1537), (1400, 2048),
              (2048, 1400), (2048, 2048), (1900, 1200), (1100, 1800), (769, 300), (300, 800), (1000, 280), (990, 1700)]
    for si, (ql, tl) in enumerate(shapes):
        for kind in range(5):
            t = rng.integers(0, 4, tl).astype(np.uint8)
            if kind == 0:                     # similar
                q = noisy(t, 0.12)
            elif kind == 1:                   # unrelated
                q = rng.integers(0, 4, ql).astype(np.uint8)
            elif kind == 2:                   # similar after a big offset: the optimum runs near the band edge
                sh = int(rng.integers(600, 760))
                q = np.concatenate([rng.integers(0, 4, sh).astype(np.uint8), noisy(t, 0.08)]) if si % 2 else noisy(t[sh:], 0.08) if tl > sh + 50 else noisy(t, 0.1)
            elif kind == 3:                   # low complexity
                t = np.tile(rng.integers(0, 4, 5).astype(np.uint8), tl // 5 + 1)[:tl]
                q = noisy(t, 0.1)
            else:                             # junk in the middle
                q = noisy(t, 0.1); a = len(q) // 3; q[a:a + len(q) // 3] = rng.integers(0, 4, len(q) // 3)
            q = q[:2048]
            if len(q) < ql:
                q = np.concatenate([q, rng.integers(0, 4, ql - len(q)).astype(np.uint8)])
            q = q[:ql] if kind in (1,) else q
            for w in (751, 400 if kind != 2 else 751, max(len(q), tl)):
                if abs(tl - len(q)) >= w or max(len(q), tl) > 2048:
                    continue
                ez = oracle.ksw_extd2(q, t, w=w, zdrop=400, end_bonus=-1, flag=0x08)
                recs.append(dict(qlen=len(q), tlen=tl, w=w, zdrop=400, end_bonus=-1, flag=0x08, q=q, t=t, score=ez["score"], max=ez["max"],
                                 max_q=ez["max_q"], max_t=ez["max_t"], mqe=ez["mqe"], mqe_t=ez["mqe_t"], zdropped=ez["zdropped"],
                                 reach_end=ez["reach_end"], n_cigar=len(ez["cigar"]), cigar=ez["cigar"]))
    if len(recs) % 2 == 0:
        recs.pop()
    assert len(recs) > 200
    tasks, cig = _run_dp(lib, opt, recs)
    _check_dp(tasks, cig, recs)


def test_sequential_pieces_equal_one_piece(case, lib, monkeypatch):
    """Batches beyond the scratch budget are mapped piece by piece (MB_PIECE_BASES; BASELINE configs[2] on one GPU is nine
    pieces): hit arrays, CIGARs, the device-resident counting (count_last) and the resident-reads entry must not notice."""
    al, oidx, reads, traces = case
    from monica_b200 import synth
    cat, off = synth.concat_reads(reads)
    want = al.map_batch(cat=cat, off=off)
    cw, nw = al.count_last(60, "query_length")
    monkeypatch.setenv("MB_PIECE_BASES", "20000")
    got = al.map_batch(cat=cat, off=off)
    assert al.last_stats["n_pieces"] >= 4
    cg, ng = al.count_last(60, "query_length")
    for f in CMP_FIELDS + ["read_idx"]:
        assert np.array_equal(getattr(want, f), getattr(got, f)), f
    assert np.array_equal(want.cigar_pool, got.cigar_pool) and np.array_equal(want.cigar_off, got.cigar_off)
    assert np.array_equal(want.rep_len, got.rep_len)
    assert np.array_equal(cw, cg) and np.array_equal(nw, ng)
    c2, n2, _, _ = al.count(got, 60, "query_length")
    assert np.array_equal(c2, cw) and np.array_equal(n2, nw)
    h = al.reads_upload(cat, off)
    try:
        res = al.map_resident(h, len(reads), want_hits=True)
        assert al.last_stats["n_pieces"] >= 4
        for f in CMP_FIELDS + ["read_idx"]:
            assert np.array_equal(getattr(want, f), getattr(res, f)), f
        assert np.array_equal(want.cigar_pool, res.cigar_pool)
        al.map_resident(h, len(reads), want_hits=False)
        c3, n3 = al.count_last(60, "basic")
        c4, n4, _, _ = al.count(want, 60, "basic")
        assert np.array_equal(c3, c4) and np.array_equal(n3, n4)
    finally:
        al.reads_free(h)


def test_batch_beyond_device_scratch_is_retried_in_smaller_pieces(lib, monkeypatch):
    """The scratch a batch needs depends on the database; a batch (or piece) that runs out of device memory is retried in
    pieces of half the size and the size that worked is kept for the index.  The device limit is imitated here
    (MB_TEST_ARENA_LIMIT); results must equal the unrestricted run's."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    names, seqs = synth.make_genomes(77, 3, 150_000)
    reads, _ = synth.simulate_reads(78, seqs, 400, 2500.0, 0.12)
    cat, off = synth.concat_reads(reads)
    al = Aligner(names=names, seqs=seqs, preset="map-ont", best_n=15)
    want = al.map_batch(cat=cat, off=off)
    assert al.last_stats["n_pieces"] <= 1
    need = int(al.last_stats["arena_bytes"])
    assert need > 0
    monkeypatch.setenv("MB_TEST_ARENA_LIMIT", str(need // 3))
    got = al.map_batch(cat=cat, off=off)
    assert al.last_stats["n_pieces"] >= 2
    for f in CMP_FIELDS + ["read_idx"]:
        assert np.array_equal(getattr(want, f), getattr(got, f)), f
    assert np.array_equal(want.cigar_pool, got.cigar_pool) and np.array_equal(want.cigar_off, got.cigar_off)
    n1 = al.last_stats["n_pieces"]
    al.map_batch(cat=cat, off=off)                      # the size that worked is remembered: no second search
    assert al.last_stats["n_pieces"] == n1
    h = al.reads_upload(cat, off)
    try:
        res = al.map_resident(h, len(reads), want_hits=True)
        for f in CMP_FIELDS + ["read_idx"]:
            assert np.array_equal(getattr(want, f), getattr(res, f)), f
    finally:
        al.reads_free(h)
    monkeypatch.setenv("MB_TEST_ARENA_LIMIT", "4096")   # nothing fits: the error surfaces instead of looping
    with pytest.raises(Exception):
        al.map_batch(cat=cat, off=off)
    # the piece size found for one index must not leak to the next one (which may be allocated where the freed one was)
    monkeypatch.delenv("MB_TEST_ARENA_LIMIT")
    import gc
    del al, got, res
    gc.collect()
    al2 = Aligner(names=names, seqs=seqs, preset="map-ont", best_n=15)
    again = al2.map_batch(cat=cat, off=off)
    assert al2.last_stats["n_pieces"] <= 1
    assert np.array_equal(want.mapq, again.mapq) and np.array_equal(want.cigar_pool, again.cigar_pool)


def test_long_reads_with_tied_anchors_sort_bit_exact(oracle, lib):
    """Long reads (more anchors than the shared-memory sorts hold) whose anchors tie on x: the radix sort in global memory plus
    the queued replay of upstream's unstable sort (only the ranges that hold a tied pair are walked; more than 64 tied pairs
    -> the whole recursion).  Anchor arrays must equal the oracle's element for element."""
    from monica_b200 import _lib, synth
    from monica_b200.mappy_shim import Aligner
    rng = np.random.default_rng(4242)
    names, seqs = synth.make_genomes(91, 6, 250_000, strain_frac=0.5)
    al = Aligner(names=names, seqs=seqs, preset="map-ont", best_n=15)
    oidx = oracle.Index(names, seqs)
    g = seqs[0]
    reads = []
    # few ties: a short block of the read copied to a second place (each repeated minimizer ties once per reference hit)
    r = g[10_000:50_000].copy(); r[30_000:30_030] = r[5_000:5_030]; reads.append(r)
    r = g[60_000:95_000].copy(); r[20_000:20_022] = r[1_000:1_022]; r[33_000:33_025] = r[9_000:9_025]; reads.append(r)
    # many ties: a 3 kb block duplicated, and a read made of one segment three times
    r = g[100_000:150_000].copy(); r[40_000:43_000] = r[2_000:5_000]; reads.append(r)
    reads.append(np.concatenate([g[160_000:172_000]] * 3))
    # no ties at all, and a mid-size read just past the shared-memory limit
    reads.append(g[180_000:240_000].copy())
    reads.append(np.concatenate([g[200_000:214_000], g[200_000:200_020]]))
    reads += [synth.random_genome(rng, 3000), g[5_000:6_000].copy()]
    cat, off = synth.concat_reads(reads)
    cap = len(cat) * 4 + 1024
    out = np.zeros((cap, 2), dtype=np.uint64)
    ooff = np.zeros(len(reads) + 1, dtype=np.int64)
    rep = np.zeros(len(reads), dtype=np.int32)
    _lib.check(lib.mb_seed(al.handle(), C.byref(al.opt), _lib._ptr(cat), _lib._ptr(off), len(reads), _lib._ptr(out), cap, _lib._ptr(ooff), _lib._ptr(rep)))
    n_ties = []
    for i, r in enumerate(reads):
        _, stats, tr = oidx.map(r, trace=True)
        got = out[ooff[i]:ooff[i + 1]]
        assert len(tr["anchors"]) == len(got), f"read {i}"
        bad = np.nonzero(np.any(tr["anchors"] != got, axis=1))[0]
        assert len(bad) == 0, f"read {i}: {len(bad)} anchors differ, first at {bad[:5]} of {len(got)}"
        n_ties.append(int(np.sum(got[1:, 0] == got[:-1, 0])) if len(got) > 1 else 0)
    assert len(out[ooff[0]:ooff[1]]) > 2560
    assert 0 < n_ties[0] <= 64 and 0 < n_ties[1] <= 64, n_ties
    assert n_ties[2] > 64 and n_ties[3] > 64, n_ties
    assert n_ties[4] == 0, n_ties


def test_exact_cta_kernel_long_windows(oracle, lib):
    """The CTA-per-task exact kernel (align_cta.cuh, four cells per thread in packed 16-bit lanes, state in circular 1024-entry
    shared-memory windows): windows far longer than the window (the buffers wrap several times), end extensions that Z-drop
    after a junk end and ones that align to the end, second passes (flag 0) with a 300-base insertion, N bases, both tie-break
    variants, bands 751 and 500, and several tasks per CTA in a row (stale shared memory between tasks)."""
    from monica_b200 import _lib
    rng = np.random.default_rng(77)
    opt = _lib.default_opt()
    recs = []

    def noisy(t, err):
        out = []
        for b in t:
            r = rng.random()
            if r < err * 0.4:
                out.append(int(rng.integers(0, 4)))
            elif r < err * 0.7:
                out += [int(b), int(rng.integers(0, 4))]
            elif r < err:
                continue
            else:
                out.append(int(b))
        return np.array(out, np.uint8)

    shapes = [(5000, 9998, "junk"), (4200, 8000, "half"), (3000, 3300, "good"), (2600, 2400, "ins"), (1800, 5000, "good"), (5000, 5200, "N")]
    for rep in range(3):
        for (ql, tl, kind) in shapes:
            t = rng.integers(0, 4, tl).astype(np.uint8)
            if kind == "junk":
                q = np.concatenate([noisy(t[:300], 0.1), rng.integers(0, 4, ql).astype(np.uint8)])[:ql]
            elif kind == "half":
                q = np.concatenate([noisy(t[:ql // 2], 0.12), rng.integers(0, 4, ql).astype(np.uint8)])[:ql]
            elif kind == "ins":
                q = np.concatenate([noisy(t[:1000], 0.1), rng.integers(0, 4, 300).astype(np.uint8), noisy(t[1000:], 0.1)])[:ql]
            else:
                q = noisy(t, 0.13)[:ql]
            if len(q) < ql:
                q = np.concatenate([q, rng.integers(0, 4, ql - len(q)).astype(np.uint8)])
            if kind == "N":
                q[rng.integers(0, ql, 12)] = 4
                t = t.copy(); t[rng.integers(0, tl, 12)] = 4
            for flag, zdrop, eb in ((0x40, 400, -1), (0x40 | 0x02 | 0x80, 400, 10), (0x00, 200, -1), (0x08, 400, -1)):
                w = 751 if rep < 2 else 500
                ez = oracle.ksw_extd2(q, t, w=w, zdrop=zdrop, end_bonus=eb, flag=flag)
                recs.append(dict(qlen=ql, tlen=tl, w=w, zdrop=zdrop, end_bonus=eb, flag=flag, q=q, t=t, score=ez["score"], max=ez["max"],
                                 max_q=ez["max_q"], max_t=ez["max_t"], mqe=ez["mqe"], mqe_t=ez["mqe_t"], zdropped=ez["zdropped"],
                                 reach_end=ez["reach_end"], n_cigar=len(ez["cigar"]), cigar=ez["cigar"]))
    assert any(r["zdropped"] for r in recs) and any(not r["zdropped"] for r in recs)
    # many more tasks than resident CTAs would be needed to force several per CTA; repeating the list does it for a class
    recs = recs * 3
    tasks, cig = _run_dp(lib, opt, recs)
    _check_dp(tasks, cig, recs)


def test_device_logf_equals_libm_on_every_positive_normal_float(oracle, lib):
    """mm_set_mapq (hit.c, minimap2-2.17) truncates a float product that contains logf(), so one differing ulp can move a MAPQ
    across monica's `hit.mapq >= mapping_quality` filter (aligner.py:194,216).  The device routine (glue.cuh: mb_logf) must
    return the host libm's float for EVERY input: all 2^31 - 2^24 positive normal bit patterns are swept, 2^24 at a time, the
    expected values from the libm the oracle links (oracle.logf_range), the comparison on the device."""
    from concurrent.futures import ThreadPoolExecutor
    from monica_b200 import _lib
    CH = 1 << 24
    firsts = list(range(0x00800000, 0x7f800000, CH))
    assert len(firsts) == 127

    def want(fb):
        return fb, oracle.logf_range(fb, CH)

    total_bad, first_bad = 0, None
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:   # ctypes releases the GIL inside libm's loop
        for fb, exp in pool.map(want, firsts):
            nb, f = C.c_int64(0), C.c_uint32(0)
            _lib.check(lib.mb_logf_sweep(0, fb, CH, exp.ctypes.data, C.byref(nb), C.byref(f)))
            if nb.value and first_bad is None:
                first_bad = f.value
            total_bad += nb.value
    assert total_bad == 0, f"{total_bad} floats differ from libm logf, the first at bits {first_bad:#010x}"
    # the sweep itself can fail: a wrong expectation is reported, with its position
    exp = oracle.logf_range(0x40000000, 1024)
    exp[77] = np.nextafter(exp[77], np.float32(10))
    nb, f = C.c_int64(0), C.c_uint32(0)
    _lib.check(lib.mb_logf_sweep(0, 0x40000000, 1024, exp.ctypes.data, C.byref(nb), C.byref(f)))
    assert (nb.value, f.value) == (1, 0x40000000 + 77)


def test_packed_reads_map_like_ascii_reads(lib, monkeypatch):
    """mb_map_packed (2-bit words + runs of ambiguous bases, expanded on the device) must return exactly what mb_map_batch
    returns from the ASCII: reads with N runs, IUPAC letters, lower case and U; the plain upload, the piecewise upload under
    the sketch kernels (SketchFeed: pieces of whole CTA spans, one word of look-ahead) and sequential pieces whose first
    base is not word-aligned (MB_PIECE_BASES)."""
    from monica_b200 import synth
    from monica_b200.mappy_shim import Aligner
    rng = np.random.default_rng(9)
    names, seqs = synth.make_genomes(5, 3, 200_000)
    reads, _ = synth.simulate_reads(6, seqs, 1500, 3000.0, 0.10)
    reads = [np.frombuffer(r.encode() if isinstance(r, str) else bytes(r), np.uint8).copy() for r in reads]
    for i in range(0, len(reads), 7):            # every 7th read carries other characters
        r = reads[i]
        if len(r) < 400:
            continue
        k = i // 7 % 5
        if k == 0:
            s = int(rng.integers(0, len(r) - 300)); r[s:s + int(rng.integers(1, 300))] = ord("N")
        elif k == 1:
            r[rng.integers(0, len(r), 20)] = np.frombuffer(b"RYKMSWn-", np.uint8)[rng.integers(0, 8, 20)]
        elif k == 2:
            r[:] = np.frombuffer(bytes(r).lower(), np.uint8)
        elif k == 3:
            r[r == ord("T")] = ord("U")
        else:
            r[-17:] = ord("N"); r[:3] = ord("N")
    cat, off = synth.concat_reads(reads)
    assert off[-1] > 9 * 32768 and (off[-1] // 32768) % 8 != 0
    al = Aligner(names=names, seqs=seqs, preset="map-ont", best_n=15)
    pk = Aligner.pack_reads(cat, off, n_threads=3)
    assert pk.upload_bytes < 0.27 * len(cat) + 8 * len(off) + 4096
    base = al.map_batch(cat=cat, off=off)
    assert base.n > 1000

    def same(got, tag):
        assert got.n == base.n, tag
        for f in ["read_idx"] + CMP_FIELDS:
            assert np.array_equal(getattr(got, f), getattr(base, f)), (tag, f)
        assert np.array_equal(got.cigar_off, base.cigar_off) and np.array_equal(got.cigar_pool, base.cigar_pool), tag
        assert np.array_equal(got.rep_len, base.rep_len), tag

    same(al.map_packed(pk), "plain")
    monkeypatch.setenv("MB_FEED_MIN_BYTES", "1")
    same(al.map_packed(pk), "piecewise upload")
    monkeypatch.setenv("MB_PIECE_BASES", "700001")
    same(al.map_packed(pk), "sequential pieces, piecewise upload")
    monkeypatch.delenv("MB_FEED_MIN_BYTES")
    same(al.map_packed(pk), "sequential pieces")
    monkeypatch.delenv("MB_PIECE_BASES")
    # a batch cut into concurrent sub-batches finishes the upload before the pieces start (packed words included)
    monkeypatch.setenv("MB_PARTS", "2"); monkeypatch.setenv("MB_PARTS_MIN_READS", "4"); monkeypatch.setenv("MB_FEED_MIN_BYTES", "1")
    same(al.map_packed(pk), "two concurrent sub-batches, piecewise upload requested")
    for k in ("MB_PARTS", "MB_PARTS_MIN_READS", "MB_FEED_MIN_BYTES"):
        monkeypatch.delenv(k)
    # a batch of one empty read, and the empty batch
    for reads0 in ([np.zeros(0, np.uint8)], []):
        c0, o0 = (np.zeros(0, np.uint8), np.zeros(len(reads0) + 1, np.int64))
        assert al.map_packed(Aligner.pack_reads(c0, o0)).n == 0


def test_one_harness_two_libraries_same_arrays(lib, small_case):
    """the boundary is implemented twice -- the CUDA library and the CPU oracle behind the same entry points
    (oracle/mm2o_abi.c).  tests/abi_harness.py drives both through raw ctypes with identical calls (options, index build,
    mb_map_batch, every hit array, mb_count in all modes, mb_sketch); every array that comes back must be identical."""
    import abi_harness as H
    from monica_b200 import synth
    names, seqs, reads = small_case
    cat, off = synth.concat_reads(reads)
    gpu = H.run(H.load(H.CUDA_SO), names, seqs, cat, off, device=0)
    cpu = H.run(H.oracle_library(), names, seqs, cat, off, device=0)
    assert gpu.keys() == cpu.keys() and gpu["n_hits"][0] > 30
    for k in gpu:
        a, b = gpu[k], cpu[k]
        if k == "sketch_xy":     # y carries the read index in its high word on both sides
            assert np.array_equal(a, b), k
        else:
            assert a.shape == b.shape and np.array_equal(a, b), k
