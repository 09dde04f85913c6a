"""CPU tests of the oracle (oracle/): known-answer checks against independent Python restatements of the published
definitions, internal consistency, and the committed golden vectors.

The reference's own tests hold no vectors for this path, and mappy is not installable here, so these
pin the oracle against (a) definitions that can be re-derived by hand (Wang hash, window minimizers, two-piece affine
global alignment score) and (b) its own previous outputs (tests/golden/, regression).
"""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
M30 = (1 << 30) - 1


def py_hash64(key, mask=M30):
    key = (~key + (key << 21)) & mask
    key = key ^ key >> 24
    key = ((key + (key << 3)) + (key << 8)) & mask
    key = key ^ key >> 14
    key = ((key + (key << 2)) + (key << 4)) & mask
    key = key ^ key >> 28
    key = (key + (key << 31)) & mask
    return key


def test_hash64_known_answers(oracle):
    # hand-evaluated: hash64(0) with a 30-bit mask
    assert py_hash64(0) == oracle.hash64(0, M30)
    rng = np.random.default_rng(0)
    keys = rng.integers(0, 1 << 30, size=2000)
    got = [oracle.hash64(int(k), M30) for k in keys]
    assert got == [py_hash64(int(k)) for k in keys]
    # the hash is a bijection on 30-bit keys: no collisions in a sample
    assert len(set(got)) == len(set(int(k) for k in keys))


def brute_minimizers(seq: bytes, w=10, k=15):
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    c = [code[b] for b in seq]
    km = []
    for i in range(k - 1, len(c)):
        f = r = 0
        for j in range(k):
            f = f << 2 | c[i - k + 1 + j]
            r = r << 2 | (3 - c[i - j])
        z = 0 if f < r else 1
        km.append((py_hash64(r if z else f), i, z))
    out = set()
    for s in range(0, len(km) - w + 1):
        win = km[s:s + w]
        m = min(x[0] for x in win)
        out.update(x for x in win if x[0] == m)
    return out


def test_sketch_equals_window_minimum_definition(oracle):
    rng = np.random.default_rng(5)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for t in range(60):
        n = int(rng.integers(30, 500))
        seq = acgt[rng.integers(0, 4, n)]
        got = oracle.sketch(seq)
        G = {(int(x) >> 8, (int(y) & 0xffffffff) >> 1, int(y) & 1) for x, y in got}
        assert len(G) == len(got)
        assert all(int(x) & 0xff == 15 for x, _ in got)
        assert G == brute_minimizers(seq.tobytes())
        pos = [(int(y) & 0xffffffff) >> 1 for _, y in got]
        assert pos == sorted(pos)


def test_sketch_ties_first_window_quirk(oracle):
    """Low-complexity input: every emitted record is a true window minimum; mm_sketch may drop one tied record of the
    very first window (old minimum replaced while l == w+k-1), never more."""
    rng = np.random.default_rng(6)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    for t in range(40):
        n = int(rng.integers(60, 400))
        unit = acgt[rng.integers(0, 4, int(rng.integers(2, 9)))]
        seq = np.tile(unit, n // len(unit) + 1)[:n]
        got = oracle.sketch(seq)
        G = {(int(x) >> 8, (int(y) & 0xffffffff) >> 1, int(y) & 1) for x, y in got}
        B = brute_minimizers(seq.tobytes())
        assert G <= B
        missing = B - G
        assert len(missing) <= 1
        assert all(p < 15 + 10 for _, p, _ in missing)


def test_sketch_edge_inputs(oracle):
    assert len(oracle.sketch(b"")) == 0
    assert len(oracle.sketch(b"ACGTACGTAC")) == 0           # shorter than k
    assert len(oracle.sketch(b"N" * 100)) == 0
    s = b"ACGTTGCAAGCTTGACCTGAAGTCCATGCAAGT"
    a, b = oracle.sketch(s), oracle.sketch(s.lower())
    assert np.array_equal(a, b)                              # case-insensitive
    # an N resets the k-mer run: no minimizer may span it
    seq = bytearray(np.frombuffer(b"ACGT", dtype=np.uint8)[np.random.default_rng(1).integers(0, 4, 300)].tobytes())
    seq[150] = ord("N")
    for x, y in oracle.sketch(bytes(seq)):
        p = (int(y) & 0xffffffff) >> 1
        assert not (p - 14 <= 150 <= p)


def py_global_2piece(q, t, a=2, b=4, o1=4, e1=2, o2=24, e2=1, amb=1):
    """Gotoh with two affine pieces; global alignment score (plain O(nm) Python)."""
    NEG = -10 ** 9
    n, m = len(t), len(q)

    def gap(l):
        return -min(o1 + l * e1, o2 + l * e2)
    H = [[NEG] * (m + 1) for _ in range(n + 1)]
    E1 = [[NEG] * (m + 1) for _ in range(n + 1)]
    E2 = [[NEG] * (m + 1) for _ in range(n + 1)]
    F1 = [[NEG] * (m + 1) for _ in range(n + 1)]
    F2 = [[NEG] * (m + 1) for _ in range(n + 1)]
    H[0][0] = 0
    for i in range(1, n + 1):
        H[i][0] = gap(i)
    for j in range(1, m + 1):
        H[0][j] = gap(j)
    for i in range(1, n + 1):
        for j in range(1, m + 1):
            E1[i][j] = max(E1[i - 1][j] - e1, H[i - 1][j] - o1 - e1)
            E2[i][j] = max(E2[i - 1][j] - e2, H[i - 1][j] - o2 - e2)
            F1[i][j] = max(F1[i][j - 1] - e1, H[i][j - 1] - o1 - e1)
            F2[i][j] = max(F2[i][j - 1] - e2, H[i][j - 1] - o2 - e2)
            s = -amb if (t[i - 1] > 3 or q[j - 1] > 3) else (a if t[i - 1] == q[j - 1] else -b)
            H[i][j] = max(H[i - 1][j - 1] + s, E1[i][j], E2[i][j], F1[i][j], F2[i][j])
    return H[n][m]


def cigar_score(cig, q, t, a=2, b=4, o1=4, e1=2, o2=24, e2=1, amb=1):
    i = j = s = 0
    for c in cig:
        op, l = int(c) & 0xf, int(c) >> 4
        if op == 0:
            for k in range(l):
                s += -amb if (t[i + k] > 3 or q[j + k] > 3) else (a if t[i + k] == q[j + k] else -b)
            i += l; j += l
        elif op == 1:
            s -= min(o1 + l * e1, o2 + l * e2); j += l
        else:
            s -= min(o1 + l * e1, o2 + l * e2); i += l
    return s, i, j


def test_extd2_global_score_matches_definition(oracle):
    rng = np.random.default_rng(3)
    from monica_b200 import synth
    for trial in range(25):
        n = int(rng.integers(5, 90))
        t = rng.integers(0, 4, n).astype(np.uint8)
        # query = mutated copy, sometimes with a long gap (exercises the second affine piece)
        q = list(t)
        for _ in range(int(rng.integers(0, 8))):
            p = int(rng.integers(0, max(1, len(q))))
            r = rng.random()
            if r < 0.4 and q:
                q[p % len(q)] = int(rng.integers(0, 4))
            elif r < 0.7:
                q[p:p] = list(rng.integers(0, 4, int(rng.integers(1, 40))))
            elif q:
                del q[p:p + int(rng.integers(1, 30))]
        if not q:
            q = [0]
        if trial % 7 == 0:
            q[0] = 4                                         # an N
        q = np.array(q, dtype=np.uint8)
        for flag in (0x08, 0x08 | 0x02):                     # approx-max global; + right-aligned gaps
            ez = oracle.ksw_extd2(q, t, w=-1, zdrop=-1, end_bonus=-1, flag=flag)
            want = py_global_2piece(list(q), list(t))
            assert ez["score"] == want, (trial, flag, len(q), len(t))
            s, ti, qi = cigar_score(ez["cigar"], list(q), list(t))
            assert (ti, qi) == (len(t), len(q))
            assert s == want
        # exact-max mode reports the same end score and a max >= score
        ez2 = oracle.ksw_extd2(q, t, w=-1, zdrop=-1, end_bonus=-1, flag=0)
        assert ez2["score"] == want and ez2["max"] >= max(0, want)


def test_extd2_extension_semantics(oracle):
    """Extension mode on a perfect match followed by junk: max sits at the end of the matching prefix; the backtrack starts
    there (no reach_end); with a generous end bonus on a full match it reaches the query end."""
    rng = np.random.default_rng(4)
    core = rng.integers(0, 4, 120).astype(np.uint8)
    q = np.concatenate([core, rng.integers(0, 4, 80).astype(np.uint8)])
    t = np.concatenate([core, rng.integers(0, 4, 80).astype(np.uint8)])
    ez = oracle.ksw_extd2(q, t, w=751, zdrop=400, end_bonus=-1, flag=0x40)
    assert ez["max"] >= 240 and ez["max_q"] >= 119 and ez["max_t"] >= 119
    assert not ez["reach_end"]
    s, ti, qi = cigar_score(ez["cigar"], list(q), list(t))
    assert (ti, qi) == (ez["max_t"] + 1, ez["max_q"] + 1) and s == ez["max"]
    ez = oracle.ksw_extd2(core, core, w=751, zdrop=400, end_bonus=5, flag=0x40)
    assert ez["reach_end"] and ez["mqe"] == 240 and ez["mqe_t"] == 119
    # Z-drop: a long run of mismatches after the match stops the extension early
    q2 = np.concatenate([core, (core[::-1] ^ 1)[:0], np.full(400, 0, np.uint8)])
    t2 = np.concatenate([core, np.full(400, 3, np.uint8)])
    ez = oracle.ksw_extd2(q2, t2, w=751, zdrop=100, end_bonus=-1, flag=0x40)
    assert ez["zdropped"] and ez["max"] == 240


def test_radix_sorts(oracle):
    rng = np.random.default_rng(7)
    for n in (0, 1, 2, 63, 64, 65, 200, 5000):
        k = rng.integers(0, 1 << 62, size=n, dtype=np.uint64)
        assert np.array_equal(oracle.radix_sort_64(k), np.sort(k))
        a = np.stack([k, np.arange(n, dtype=np.uint64)], axis=1)
        s = oracle.radix_sort_128x(a)
        assert np.array_equal(s[:, 0], np.sort(k))
        assert sorted(s[:, 1].tolist()) == list(range(n))
    # ties: still sorted and a permutation, but NOT necessarily stable for n > 64 (upstream's MSD radix is unstable)
    k = rng.integers(0, 8, size=500, dtype=np.uint64) << np.uint64(40)
    a = np.stack([k, np.arange(500, dtype=np.uint64)], axis=1)
    s = oracle.radix_sort_128x(a)
    assert np.array_equal(s[:, 0], np.sort(k)) and sorted(s[:, 1].tolist()) == list(range(500))
    small = np.stack([np.array([3, 1, 3, 1, 2], np.uint64), np.arange(5, dtype=np.uint64)], axis=1)
    assert oracle.radix_sort_128x(small)[:, 1].tolist() == [1, 3, 4, 0, 2]   # <= 64 elements: insertion sort, stable


def test_index_and_mid_occ(oracle):
    from monica_b200 import synth
    names, seqs = synth.make_genomes(1, 1, 50000)
    idx = oracle.Index(names, seqs)
    assert idx.mid_occ == 2                                  # all minimizers unique -> quantile 1, +1
    names, seqs = synth.make_genomes(1, 3, 50000, strain_frac=0.34)
    idx = oracle.Index(names, seqs)
    assert idx.mid_occ == 3
    # every reference minimizer is found at its own position
    m = oracle.sketch(seqs[1], rid=1)
    for x, y in m[::97]:
        assert int(y) in idx.get(int(x) >> 8).tolist()


def test_reads_map_back_to_origin(oracle):
    from monica_b200 import synth
    names, seqs = synth.make_genomes(21, 2, 80000, strain_frac=0.0)
    idx = oracle.Index(names, seqs)
    reads, truth = synth.simulate_reads(22, seqs, 25, 3000, 0.10)
    for r, (c, st, en, strand) in zip(reads, truth):
        hits, stats = idx.map(r)
        assert hits, "read did not map"
        h = hits[0]
        assert h["rid"] == c and h["rev"] == strand and h["mapq"] == 60 and h["is_primary"]
        assert abs(h["rs"] - st) < 60 and abs(h["re"] - en) < 60
        assert 0.80 < h["mlen"] / h["blen"] < 0.97
        assert h["nm"] == h["blen"] - h["mlen"]
        # the CIGAR consumes exactly the reported intervals
        ql = sum(int(c_) >> 4 for c_ in h["cigar"] if int(c_) & 0xf in (0, 1))
        tl = sum(int(c_) >> 4 for c_ in h["cigar"] if int(c_) & 0xf in (0, 2))
        assert ql == h["qe"] - h["qs"] and tl == h["re"] - h["rs"]


def test_empty_and_unmappable(oracle, small_case):
    names, seqs, reads = small_case
    idx = oracle.Index(names, seqs)
    assert idx.map(b"")[0] == []
    assert idx.map(b"ACGT")[0] == []
    rng = np.random.default_rng(9)
    junk = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 4000)]
    assert idx.map(junk)[0] == []


def test_batch_threads_equal_serial(oracle, small_case):
    from monica_b200 import synth
    names, seqs, reads = small_case
    idx = oracle.Index(names, seqs)
    cat, off = synth.concat_reads(reads)
    a, _ = idx.map_batch(cat, off, n_threads=1)
    b, _ = idx.map_batch(cat, off, n_threads=4)
    assert len(a) == len(b) == len(reads)
    for x, y in zip(a, b):
        assert [{k: v for k, v in h.items() if k != "cigar"} for h in x] == [{k: v for k, v in h.items() if k != "cigar"} for h in y]


def test_golden_vectors(oracle):
    """Regression: the oracle reproduces the committed vectors (tests/golden/make_golden.py wrote them from this oracle)."""
    path = os.path.join(GOLDEN, "small_case.npz")
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated")
    g = np.load(path, allow_pickle=False)
    names = [n for n in g["names"].tolist()]
    seqs = [g["genome_cat"][g["genome_off"][i]:g["genome_off"][i + 1]] for i in range(len(names))]
    idx = oracle.Index(names, seqs)
    assert idx.mid_occ == int(g["mid_occ"])
    hit_off = g["hit_off"]
    fields = [f for f in g["hit_fields"].tolist()]
    for i in range(len(g["read_off"]) - 1):
        r = g["read_cat"][g["read_off"][i]:g["read_off"][i + 1]]
        hits, _ = idx.map(r)
        want = g["hits"][hit_off[i]:hit_off[i + 1]]
        assert len(hits) == len(want)
        for h, w in zip(hits, want):
            assert [h[f] for f in fields] == w.tolist()
    m = oracle.sketch(g["read_cat"][g["read_off"][3]:g["read_off"][4]])
    assert np.array_equal(m, g["sketch_read3"])


def test_extd2_simd_core_equals_scalar(oracle):
    """The oracle's SSE4.1 anti-diagonal core (the speed class of upstream's ksw2_extd2_sse) and its scalar statement give
    identical scores, end points and CIGARs over shapes, bands and all flag combinations the aligner uses."""
    rng = np.random.default_rng(91)
    try:
        for (ql, tl) in [(1, 1), (5, 40), (40, 5), (16, 16), (17, 33), (200, 230), (255, 257), (700, 900), (1200, 300)]:
            t = rng.integers(0, 4, tl).astype(np.uint8)
            q = rng.integers(0, 4, ql).astype(np.uint8)
            n = min(ql, tl)
            q[:n] = np.where(rng.random(n) < 0.85, t[:n], q[:n])
            if ql > 20:
                q[9] = 4
            for flag, zd, eb in ((0x08, 400, -1), (0x00, 400, -1), (0x40, 100, -1), (0x40 | 0x02 | 0x80, 200, 10), (0x01 | 0x40, 400, -1)):
                for w in (751, 60):
                    oracle.set_simd(False)
                    a = oracle.ksw_extd2(q, t, w=w, zdrop=zd, end_bonus=eb, flag=flag)
                    oracle.set_simd(True)
                    b = oracle.ksw_extd2(q, t, w=w, zdrop=zd, end_bonus=eb, flag=flag)
                    for k in a:
                        assert np.array_equal(a[k], b[k]), (ql, tl, hex(flag), w, k)
    finally:
        oracle.set_simd(True)


def test_chain_dp_equals_the_recurrence_written_out_in_python(oracle, small_case):
    """The chaining DP (chain.c mm_chain_dp, 2.17 form) restated a SECOND time, in plain Python straight from the recurrence
    as SURVEY.md Appendix A.4 writes it, and held against the f / p / v arrays the C oracle traces for real reads: anchors of
    both strands and several references, reads with repeats (the max_skip break), the max_iter window, float32 avg_qspan and the
    double-precision gap cost truncated to int.  Independent of oracle/mm2o_map.c apart from the shared recollection."""
    names, seqs, reads = small_case
    oidx = oracle.Index(names, seqs)
    opt = oidx.opt
    max_dist = opt.max_gap                      # max_gap_ref = -1: both chain gaps are max_gap (map.c mm_map_frag)
    bw, max_skip, max_iter = opt.bw, opt.max_chain_skip, opt.max_chain_iter
    checked = n_break = 0
    for r in reads[:26]:
        _, _, tr = oidx.map(r, trace=True)
        a = tr["anchors"]
        n = len(a)
        if n == 0:
            continue
        assert len(tr["f"]) == n
        X = [int(v) for v in a[:, 0]]
        Y = [int(v) for v in a[:, 1]]
        span = [(y >> 32) & 0xff for y in Y]
        qpos = [((y & 0xffffffff) ^ 0x80000000) - 0x80000000 for y in Y]     # (int32_t)a[i].y
        avg_qspan = float(np.float32(sum(span)) / np.float32(n))             # (float)sum / n, then promoted in the product
        f, p, v, t = [0] * n, [-1] * n, [0] * n, [0] * n
        st = 0
        for i in range(n):
            while st < i and X[i] > X[st] + max_dist:
                st += 1
            if i - st > max_iter:
                st = i - max_iter
            max_f, max_j, n_skip = span[i], -1, 0
            for j in range(i - 1, st - 1, -1):
                dr, dq = X[i] - X[j], qpos[i] - qpos[j]
                if dr == 0 or dq <= 0:
                    continue
                if dq > max_dist:
                    continue
                dd = abs(dr - dq)
                if dd > bw:
                    continue
                min_d = min(dq, dr)
                sc = span[i] if min_d > span[i] else min_d
                log_dd = dd.bit_length() - 1 if dd else 0
                sc -= int(dd * .01 * avg_qspan) + (log_dd >> 1)
                sc += f[j]
                if sc > max_f:
                    max_f, max_j = sc, j
                    if n_skip > 0:
                        n_skip -= 1
                elif t[j] == i:
                    n_skip += 1
                    if n_skip > max_skip:
                        n_break += 1
                        break
                if p[j] >= 0:
                    t[p[j]] = i
            f[i], p[i] = max_f, max_j
            v[i] = v[max_j] if max_j >= 0 and v[max_j] > max_f else max_f
        assert np.array_equal(np.array(f, np.int32), tr["f"]), "f"
        assert np.array_equal(np.array(p, np.int32), tr["p"]), "p"
        assert np.array_equal(np.array(v, np.int32), tr["v"]), "v"
        checked += n
    assert checked > 5000


def test_mapq_equals_the_formula_written_out_in_float32(oracle, small_case):
    """mm_set_mapq (hit.c, 2.17) a second time, in numpy float32 one rounding per operation, from the formula as SURVEY.md
    Appendix A.7 states it (uniq_ratio, pen_s1, pen_cm, the dp_max2 branch with its BWA-like cap, the n_sub penalty, the clamp
    to [0, 60] and the 0 -> 1 rule), evaluated on the fields the oracle reports for real reads and compared with its MAPQ.
    The simulated reads only: the hand-shaped edge reads may carry inversion hits, whose MAPQ comes from mm_set_inv_mapq."""
    from monica_b200 import synth
    names, seqs, reads = small_case
    n_edge = len(synth.edge_reads(13, seqs))
    oidx = oracle.Index(names, seqs)
    f32 = np.float32
    match_sc, min_chain_sc = oidx.opt.a, oidx.opt.min_chain_score

    def logf(x):
        return oracle.logf_range(int(np.array([x], np.float32).view(np.uint32)[0]), 1)[0]

    n_checked = n_lt60 = n_branch2 = 0
    for r in reads[n_edge:]:
        hits, stats = oidx.map(r)
        if not hits:
            continue
        sum_sc = sum(h["score"] for h in hits if h["parent"] == h["id"])
        uniq_ratio = f32(sum_sc) / f32(sum_sc + stats["rep_len"])
        for h in hits:
            if h["parent"] != h["id"]:
                want = 0
            else:
                pen_s1 = (f32(1.0) if h["score"] > 100 else f32(0.01) * f32(h["score"])) * uniq_ratio
                pen_cm = f32(1.0) if h["cnt"] > 10 else f32(0.1) * f32(h["cnt"])
                pen_cm = pen_s1 if pen_s1 < pen_cm else pen_cm
                subsc = max(h["subsc"], min_chain_sc)
                identity = f32(h["mlen"]) / f32(h["blen"])
                lg = logf(f32(h["dp_max"]) / f32(match_sc))
                if h["dp_max2"] > 0 and h["dp_max"] > 0:
                    n_branch2 += 1
                    x = f32(h["dp_max2"]) * f32(subsc) / f32(h["dp_max"]) / f32(h["score0"])
                    mapq = int(identity * pen_cm * f32(40.0) * (f32(1.0) - x * x) * lg)
                    alt = int(f32(6.02) * identity * identity * f32(h["dp_max"] - h["dp_max2"]) / f32(match_sc) + f32(.499))
                    mapq = min(mapq, alt)
                else:
                    x = f32(subsc) / f32(h["score0"])
                    mapq = int(identity * pen_cm * f32(40.0) * (f32(1.0) - x) * lg)
                mapq -= int(f32(4.343) * logf(f32(h["n_sub"] + 1)) + f32(.499))
                mapq = min(max(mapq, 0), 60)
                if h["dp_max"] > h["dp_max2"] and mapq == 0:
                    mapq = 1
                want = mapq
            assert h["mapq"] == want, (h, want)
            n_checked += 1
            n_lt60 += 0 < want < 60
    assert n_checked >= 30 and n_lt60 >= 1 and n_branch2 >= 1


def test_hit_fields_follow_from_the_cigar_and_the_sequences(oracle, small_case):
    """mlen, blen, NM (what monica's best_hit divides: aligner.py:195,217,328-339) and dp_max recomputed in plain Python from
    each hit's CIGAR and the two sequences it claims to align (mm_update_extra's definitions: M columns with equal bases count
    into mlen, ambiguous columns into neither mlen nor blen, NM = blen - mlen + n_ambi; dp_max = the best running score with
    the (q, e) gap cost, floored at 0), for hits on both strands, secondaries and split regions of the small case."""
    names, seqs, reads = small_case
    oidx = oracle.Index(names, seqs)
    opt = oidx.opt
    tab = np.full(256, 4, np.uint8)
    for ch, val in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
        tab[ord(ch)] = tab[ord(ch.lower())] = val
    ref = [tab[np.frombuffer(bytes(s), np.uint8)] if not isinstance(s, np.ndarray) else tab[s] for s in seqs]
    n_hits = n_rev = 0
    for r in reads:
        q = tab[np.asarray(r, np.uint8)]
        qrc = np.where(q[::-1] < 4, 3 - q[::-1], 4).astype(np.uint8)
        hits, _ = oidx.map(r)
        for h in hits:
            qq = (qrc[len(q) - h["qe"]:len(q) - h["qs"]] if h["rev"] else q[h["qs"]:h["qe"]])
            tt = ref[h["rid"]][h["rs"]:h["re"]]
            qo = to = 0
            mlen = blen = n_ambi = s = best = 0
            for c in h["cigar"]:
                op, ln = int(c) & 0xf, int(c) >> 4
                if op == 0:
                    a, b = qq[qo:qo + ln], tt[to:to + ln]
                    amb = (a > 3) | (b > 3)
                    eq = (a == b) & ~amb
                    mlen += int(eq.sum()); blen += ln - int(amb.sum()); n_ambi += int(amb.sum())
                    for am, e_ in zip(amb.tolist(), eq.tolist()):
                        s += -opt.sc_ambi if am else (opt.a if e_ else -opt.b)
                        if s < 0:
                            s = 0
                        elif s > best:
                            best = s
                    qo += ln; to += ln
                else:
                    seg = qq[qo:qo + ln] if op == 1 else tt[to:to + ln]
                    amb = int((seg > 3).sum())
                    blen += ln - amb; n_ambi += amb
                    s = max(s - (opt.q + opt.e * ln), 0)
                    if op == 1:
                        qo += ln
                    else:
                        to += ln
            assert (qo, to) == (len(qq), len(tt))
            assert (h["mlen"], h["blen"], h["nm"], h["dp_max"]) == (mlen, blen, blen - mlen + n_ambi, best), h
            n_hits += 1; n_rev += h["rev"]
    assert n_hits > 50 and 5 < n_rev < n_hits - 5


def test_cigars_are_in_mm_fix_cigar_normal_form(oracle, small_case):
    """What mm_fix_cigar (align.c) guarantees, checked as properties of every final CIGAR instead of by re-running it: no
    zero-length ops, no two neighbouring ops of the same kind, first and last op are M, and every indel that follows an M sits
    at its LEFTMOST position (the base before the gap differs from the gap's last base on the gapped sequence -- otherwise
    upstream would have shifted it one further left)."""
    names, seqs, reads = small_case
    oidx = oracle.Index(names, seqs)
    tab = np.full(256, 4, np.uint8)
    for ch, val in (("A", 0), ("C", 1), ("G", 2), ("T", 3)):
        tab[ord(ch)] = tab[ord(ch.lower())] = val
    ref = [tab[np.asarray(s, np.uint8)] for s in seqs]
    n_indel = 0
    for r in reads:
        q = tab[np.asarray(r, np.uint8)]
        qrc = np.where(q[::-1] < 4, 3 - q[::-1], 4).astype(np.uint8)
        for h in oidx.map(r)[0]:
            qq = qrc[len(q) - h["qe"]:len(q) - h["qs"]] if h["rev"] else q[h["qs"]:h["qe"]]
            tt = ref[h["rid"]][h["rs"]:h["re"]]
            ops = [(int(c) & 0xf, int(c) >> 4) for c in h["cigar"]]
            assert ops and ops[0][0] == 0 and ops[-1][0] == 0 and all(ln > 0 for _, ln in ops)
            assert all(ops[k][0] != ops[k - 1][0] for k in range(1, len(ops)))
            qo = to = 0
            prev = None
            for op, ln in ops:
                if op == 0:
                    qo += ln; to += ln
                else:
                    seq, at = (qq, qo) if op == 1 else (tt, to)
                    if prev == 0:
                        assert seq[at - 1] != seq[at + ln - 1], (h["rid"], h["qs"], op, ln, at)
                        n_indel += 1
                    if op == 1:
                        qo += ln
                    else:
                        to += ln
                prev = op
    assert n_indel > 5000


def test_region_logic_properties_of_the_final_hits(oracle):
    """hit.c's region logic checked through what it must leave behind, on strain copies (secondaries) and chimeric reads
    (several primaries): primaries overlap each other on the query by at most mask_level x the shorter one (mm_set_parent);
    every secondary hangs off a primary that it overlaps, with score >= pri_ratio x the parent's or within 2k of it, at most
    best_n of them per read (mm_select_sub); is_primary == (id == parent) and only primaries carry a non-zero MAPQ
    (mm_set_mapq); hits come ordered by dp_max within what mm_hit_sort compares."""
    from monica_b200 import synth
    names, seqs = synth.make_genomes(11, 3, 60000, strain_frac=0.34)       # the small case's genomes: one is a close strain copy
    plain, _ = synth.simulate_reads(32, seqs, 100, 2500, 0.10, junk_frac=0.05)
    rng = np.random.default_rng(33)
    chim = [np.concatenate([plain[int(a)], plain[int(b)]]) for a, b in rng.integers(0, len(plain), (30, 2))]
    reads = plain + chim
    oidx = oracle.Index(names, seqs)
    opt = oidx.opt
    cat, off = synth.concat_reads(reads)
    all_hits, _ = oidx.map_batch(cat, off, n_threads=4)
    n_pairs = n_sec = n_multi = 0
    for hits in all_hits:
        byid = {h["id"]: h for h in hits}
        assert len(byid) == len(hits)
        pri = [h for h in hits if h["id"] == h["parent"]]
        n_multi += len(pri) > 1
        for h in hits:
            assert h["is_primary"] == (h["id"] == h["parent"])
            if not h["is_primary"]:
                assert h["mapq"] == 0
        for a in range(len(pri)):
            for b in range(a + 1, len(pri)):
                x, y = pri[a], pri[b]
                ol = min(x["qe"], y["qe"]) - max(x["qs"], y["qs"])
                assert ol <= opt.mask_level * min(x["qe"] - x["qs"], y["qe"] - y["qs"]), (x, y)
                n_pairs += 1
        sec = [h for h in hits if h["id"] != h["parent"]]
        assert len(sec) <= opt.best_n
        for h in sec:
            p = byid[h["parent"]]
            assert p["id"] == p["parent"]
            assert min(h["qe"], p["qe"]) > max(h["qs"], p["qs"])
            # (float32 like upstream: 810 * 0.8f rounds to 648.0f, so a secondary with score 648 stays)
            assert np.float32(h["score"]) >= np.float32(p["score"]) * np.float32(opt.pri_ratio) or h["score"] + 2 * 15 >= p["score"], (h, p)
            n_sec += 1
    assert n_pairs >= 20 and n_sec >= 30 and n_multi >= 15, (n_pairs, n_sec, n_multi)


def test_anchors_equal_seeding_from_the_definition(oracle, small_case):
    """Index, lookup, the mid_occ filter and the anchor coordinates (index.c mm_idx_get, map.c collect_seed_hits) a second time
    in plain Python: a dict from minimizer hash to every (contig, position, strand) of the reference sketches, every query
    minimizer whose hash occurs fewer than mid_occ times paired with each occurrence, forward anchors (x = rid<<32 | rpos,
    y = span<<32 | qpos) and reverse ones (bit 63; qpos mirrored to the reverse-complemented read) as SURVEY Appendix A.3
    gives them.  Compared as multisets with the anchors the C oracle feeds to chaining (flag bits above the span masked)."""
    names, seqs, reads = small_case
    oidx = oracle.Index(names, seqs)
    occ = {}
    for rid, s in enumerate(seqs):
        for x, y in oracle.sketch(s, 10, 15, rid).tolist():
            occ.setdefault(x >> 8, []).append((rid, (y & 0xffffffff) >> 1, y & 1))
    mid_occ = oidx.mid_occ
    n_anchor = n_dropped = 0
    for r in reads[:30]:
        _, _, tr = oidx.map(r, trace=True)
        want = []
        qlen = len(r)
        for x, y in tr["mini"].tolist():
            hits = occ.get(x >> 8, [])
            if len(hits) >= mid_occ:
                n_dropped += len(hits) > 0
                continue
            span, qpos, qstrand = x & 0xff, (y & 0xffffffff) >> 1, y & 1
            for rid, rpos, rstrand in hits:
                if rstrand == qstrand:
                    want.append((rid << 32 | rpos, span << 32 | qpos))
                else:
                    want.append((1 << 63 | rid << 32 | rpos, span << 32 | (qlen - (qpos + 1 - span) - 1)))
        got = [(int(ax), int(ay) & 0xffffffffff) for ax, ay in tr["anchors"].tolist()]
        assert sorted(got) == sorted(want)
        assert [g[0] for g in got] == sorted(g[0] for g in got)      # radix_sort_128x: ascending x
        n_anchor += len(got)
    assert n_anchor > 5000 and n_dropped > 0


def test_chain_backtrack_equals_the_procedure_written_out_in_python(oracle, small_case):
    """The second half of mm_chain_dp (chain.c, 2.17) a second time in plain Python, starting from the traced f / p / v: chain
    ends are the anchors nobody chose as predecessor with v >= min_chain_score; each end walks back to its f peak; peaks are
    taken best first (keys f<<32 | index are unique, so the sort has one answer); a backtrack stops at an anchor that is
    already used and then keeps score - f[that anchor]; chains need min_cnt anchors and min_chain_score; the survivors are
    emitted oldest anchor first and ordered by their first anchor's x.  Compared with the oracle's u[] and chained anchors."""
    names, seqs, reads = small_case
    oidx = oracle.Index(names, seqs)
    min_sc, min_cnt = oidx.opt.min_chain_score, oidx.opt.min_cnt
    # plus a read across a tandem repeat (12 diverged copies of a 400-base unit): many anchors share predecessors there, so
    # backtracks run into anchors that better chains have taken
    rng = np.random.default_rng(5)

    def mutated(s, rate):
        s = s.copy(); m = rng.random(len(s)) < rate; s[m] = rng.integers(0, 4, int(m.sum())); return s
    unit = rng.integers(0, 4, 400)
    g = np.concatenate([rng.integers(0, 4, 20000)] + [mutated(unit, 0.03) for _ in range(12)] + [rng.integers(0, 4, 20000)])
    acgt = np.frombuffer(b"ACGT", np.uint8)
    ridx = oracle.Index(["Sp:ACC1"], [acgt[g]])
    cases = [(oidx, r) for r in reads] + [(ridx, acgt[mutated(g[18000:27500], 0.08)])]
    n_chains = n_cut = 0
    for ix, r in cases:
        _, _, tr = ix.map(r, trace=True)
        a, f, p, v = tr["anchors"].tolist(), tr["f"].tolist(), tr["p"].tolist(), tr["v"].tolist()
        n = len(a)
        if n == 0:
            assert len(tr["u"]) == 0
            continue
        used_as_pred = [False] * n
        for i in range(n):
            if p[i] >= 0:
                used_as_pred[p[i]] = True
        ends = []
        for i in range(n):
            if not used_as_pred[i] and v[i] >= min_sc:
                j = i
                while j >= 0 and f[j] < v[j]:
                    j = p[j]
                if j < 0:
                    j = i
                ends.append((f[j], j))
        ends.sort(reverse=True)
        taken = [False] * n
        chains = []                                   # (score, [anchor indices, newest first])
        for sc, j in ends:
            idx = []
            while True:
                idx.append(j); taken[j] = True
                j = p[j]
                if j < 0 or taken[j]:
                    break
            keep = None
            if j < 0:
                keep = sc
            else:
                n_cut += 1                            # ran into an anchor of a better chain
                if sc - f[j] >= min_sc:
                    keep = sc - f[j]
            if keep is not None and len(idx) >= min_cnt:
                chains.append((keep, idx))
            # (upstream leaves t[] set for a dropped chain too: its anchors keep blocking later backtracks)
        firsts = [a[idx[-1]][0] for _, idx in chains]
        order = sorted(range(len(chains)), key=lambda k: firsts[k])
        want_u = [chains[k][0] << 32 | len(chains[k][1]) for k in order]
        want_a = [a[i] for k in order for i in reversed(chains[k][1])]
        got_u, got_a = [int(x) for x in tr["u"]], tr["chained"].tolist()
        if len(set(firsts)) == len(firsts):
            assert got_u == want_u and got_a == want_a
        else:     # two chains start at the same x: upstream's unstable sort decides their order
            assert sorted(got_u) == sorted(want_u) and sorted(map(tuple, got_a)) == sorted(map(tuple, want_a))
        n_chains += len(chains)
    assert n_chains >= 40 and n_cut >= 1
