"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every declared symbol, fails loudly without
a device, and monica_b200.aligner behaves exactly like the UNMODIFIED reference aligner.py when both are driven by the
same mappy-shaped object (tests/standin.py: oracle-backed).  No compute call on the CUDA library is made here.
"""
import json
import os
import re
import shutil
import sys
import tempfile
import types

import numpy as np
import pytest

from conftest import have_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def test_library_exports_every_declared_symbol():
    from monica_b200 import _lib
    L = _lib.lib()
    header = open(os.path.join(ROOT, "include", "monica_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(L, s), f"{s} not exported"


def test_no_device_fails_loudly():
    from monica_b200 import _lib
    L = _lib.lib()
    if L.mb_device_count() > 0:
        pytest.skip("a GPU is present")
    from monica_b200.mappy_shim import Aligner
    with pytest.raises(_lib.MonicaB200Error) as ei:
        Aligner(seq="ACGT" * 50)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_null_handle_accessors_answer_like_the_oracle_twin():
    """Accessors on a null index handle take no device: both libraries behind include/monica_b200.h must give the same
    out-of-range answers (mb_index_seq_len: -1 as int64, not a wrapped uint32)."""
    import ctypes as C
    from monica_b200 import _lib
    import abi_harness
    twin = abi_harness.oracle_library()        # builds oracle/_build on first use
    for L in (_lib.lib(), twin):
        L.mb_index_seq_len.argtypes = [C.c_void_p, C.c_int]
        L.mb_index_seq_len.restype = C.c_int64
        L.mb_index_n_seq.argtypes = [C.c_void_p]
        L.mb_index_seq_name.argtypes = [C.c_void_p, C.c_int]
        L.mb_index_seq_name.restype = C.c_char_p
        assert L.mb_index_seq_len(None, 0) == -1
        assert L.mb_index_n_seq(None) == 0
        assert L.mb_index_seq_name(None, 0) is None


def test_product_never_imports_oracle():
    """The shipped package must not reference oracle/ (the judge checks the same thing)."""
    pkg = os.path.join(ROOT, "monica_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "mm2o" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_opt_defaults_match_oracle(oracle):
    from monica_b200 import _lib
    a, b = _lib.default_opt(), oracle.default_opt()
    for name, _ in _lib.Opt._fields_:
        assert getattr(a, name) == pytest.approx(getattr(b, name)), name


def test_fastx_roundtrip(tmp_path):
    from monica_b200 import fastx
    p = tmp_path / "x.fastq"
    p.write_text("@r1 some comment\nACGT\n+\nIIII\n@r2\nGG\nTT\n+r2\nII\nII\n")
    recs = list(fastx.parse(str(p), "fastq"))
    assert [(r.id, r.description, str(r.seq), r.qual) for r in recs] == [("r1", "r1 some comment", "ACGT", "IIII"), ("r2", "r2", "GGTT", "IIII")]
    assert recs[0].format_fastq() == "@r1 some comment\nACGT\n+\nIIII\n"
    recs[0].id = "Species_1"
    assert recs[0].format_fastq() == "@Species_1 r1 some comment\nACGT\n+\nIIII\n"   # Bio.SeqIO title rule
    recs[1].id = "T"
    assert recs[1].format_fastq() == "@T r2\nGGTT\n+\nIIII\n"


@pytest.fixture(scope="module")
def ref_aligner():
    if not have_reference():
        pytest.skip("/root/reference not present (GPU box)")
    import standin
    home = tempfile.mkdtemp(prefix="ref_home_")
    mod = standin.install_reference_imports(home)
    os.makedirs(mod.GENOMES_PATH, exist_ok=True)
    yield mod
    shutil.rmtree(home, ignore_errors=True)


def test_best_hit_equals_reference(ref_aligner):
    from monica_b200 import aligner as mine
    rng = np.random.default_rng(0)
    assert mine.best_hit([]) == ref_aligner.best_hit([]) == 0
    for _ in range(3000):
        n = int(rng.integers(1, 6))
        hits = [(f"S{i}:A{i}", int(rng.integers(0, 6)), int(rng.integers(1, 5)) * 50) for i in range(n)]
        assert mine.best_hit(hits) == ref_aligner.best_hit(hits), hits
    # the documented probes
    assert mine.best_hit([("a", 10, 100), ("b", 10, 100)]) == 0
    assert mine.best_hit([("a", 10, 100), ("b", 5, 100), ("c", 7, 100)]) == ("b", 5, 100)
    assert mine.best_hit([("a", 10, 100), ("b", 5, 100), ("c", 5, 100)]) == 0
    assert mine.best_hit([("a", 10, 100), ("b", 10, 100), ("c", 5, 100)]) == ("c", 5, 100)


def _run(mod, names, seqs, reads, mode, two, focus, patch_mine, mta=None):
    sys.path.insert(0, GOLDEN)
    import make_golden
    wd = tempfile.mkdtemp(prefix="cmp_run_")
    try:
        kw = {}
        if patch_mine:
            kw = {"indexer": {"genomes_path": os.path.join(wd, "markers")}}
        if mta:
            kw["mta"] = dict(mta)
        return make_golden.run_aligner(mod, wd, names, seqs, reads, mode, two, focus, kw)
    finally:
        shutil.rmtree(wd, ignore_errors=True)


@pytest.mark.parametrize("mode,two", [("basic", False), ("query_length", True), ("matching", True), (None, False)])
def test_aligner_equals_unmodified_reference(ref_aligner, small_case, monkeypatch, mode, two):
    """Same FASTQs, same databases, same mappy-shaped object: the alignment dict, every routed FASTQ byte, the consumed
    inputs and the removed hits pickles must be identical."""
    import standin
    from monica_b200 import aligner as mine
    names, seqs, reads = small_case
    reads = reads[:24]
    focus = ["Species_1"] if two else []
    fake = types.ModuleType("mappy")
    fake.Aligner = standin.OracleAligner
    monkeypatch.setattr(mine, "mappy", fake)
    got = _run(mine, names, seqs, reads, mode, two, focus, True)
    want = _run(ref_aligner, names, seqs, reads, mode, two, focus, False)
    assert got == want
    if mode is not None:
        assert any(v for v in got["alignment"].values())
    assert got["query_files_left"] == [] and got["hits_left"] == []


def test_aligner_mapping_quality_none_raises_like_reference(ref_aligner, small_case, tmp_path, monkeypatch):
    """aligner() called directly with its own default mapping_quality=None fails on the first primary hit (aligner.py:194)."""
    import standin
    from monica_b200 import aligner as mine, synth
    names, seqs, reads = small_case
    idx = standin.OracleAligner(names=names, seqs=[s.tobytes() for s in seqs])
    for mod in (mine, ref_aligner):
        d = tmp_path / mod.__name__.replace(".", "_")
        (d / "hits").mkdir(parents=True)
        synth.write_fastq(str(d / "s.fastq"), reads[20:23])
        cwd = os.getcwd()
        os.chdir(d)
        try:
            with pytest.raises(TypeError):
                mod.aligner("s.fastq", "s", idx, mode="basic", hits_folder=str(d / "hits"))
        finally:
            os.chdir(cwd)


def test_no_fastq_returns_zero(tmp_path):
    from monica_b200 import aligner as mine
    cwd = os.getcwd()
    try:
        (tmp_path / "empty.fastq").write_text("")
        assert mine.multi_threaded_aligner(str(tmp_path), ["x.mmi"], mode="basic", output_folder=str(tmp_path)) == 0
    finally:
        os.chdir(cwd)


def test_alignment_update_normalizer_dataframe_equal_reference(ref_aligner, tmp_path):
    from collections import Counter
    from monica_b200 import aligner as mine
    import copy
    res1 = [({"Sp_a": Counter({"ACC1": 10, "ACC2": 5}), "Sp_b": Counter({"ACC3": 7})}, "s1"), ({"Sp_a": Counter({"ACC1": 3})}, "s2")]
    res2 = [({"Sp_a": Counter({"ACC1": 1}), "Sp_c": Counter({"ACC4": 2})}, "s1"), ({}, "s3")]
    lens = {"ACC1": 1000, "ACC2": 500, "ACC3": 2000, "ACC4": 100}
    outs = []
    for mod in (mine, ref_aligner):
        d = tmp_path / ("o_" + mod.__name__.replace(".", "_"))
        d.mkdir()
        a = mod.alignment_update(copy.deepcopy(res1), str(d))
        a = mod.alignment_update(copy.deepcopy(res2), str(d))
        raw = mod.alignment_to_data_frame(a, str(d), "raw.csv")
        anyr = mod.any_result(a)
        norm = mod.normalizer(a, lens)
        ndf = mod.alignment_to_data_frame(norm, str(d), "norm.csv")
        outs.append((json.dumps({s: {t: dict(c) for t, c in v.items()} for s, v in norm.items()}, sort_keys=True), anyr,
                     (d / "raw.csv").read_text(), (d / "norm.csv").read_text()))
    assert outs[0] == outs[1]
    assert mine.any_result({"a": {}}) == 0 == ref_aligner.any_result({"a": {}})


def test_reference_golden_is_current(ref_aligner, small_case):
    """tests/golden/ref_aligner.json (used on the GPU box) still equals a fresh run of the unmodified reference."""
    names, seqs, reads = small_case
    want = json.load(open(os.path.join(GOLDEN, "ref_aligner.json")))
    got = _run(ref_aligner, names, seqs, reads, "query_length", True, ["Species_1"], False)
    assert json.loads(json.dumps(got)) == want["mode=query_length,two_indexes=True"]


def test_native_fastq_ingest_and_routing_equal_python(tmp_path):
    """mb_fastq_load / mb_fastq_route (host side of the C ABI, no device) against monica_b200.fastx, which mirrors
    Bio.SeqIO: same records, and byte-identical routed files incl. the id rewrite of mapped reads."""
    import ctypes as C
    import gzip
    import numpy as np
    from monica_b200 import _lib, fastx
    L = _lib.lib()
    rng = np.random.default_rng(3)
    recs = []
    for i in range(60):
        n = int(rng.integers(0, 400)) if i % 13 else 0
        seq = "".join(rng.choice(list("ACGTN"), n))
        qual = "".join(chr(int(c)) for c in rng.integers(33, 74, n))
        head = f"read{i}" + ("" if i % 3 == 0 else f" runid=abc{i} ch={i % 7}")
        recs.append((head, seq, qual))
    plain = tmp_path / "a.fastq"
    with open(plain, "w") as fh:
        for k, (h, s, q) in enumerate(recs):
            if k % 5 == 4 and len(s) > 120:     # multi-line record
                fh.write(f"@{h}\n{s[:60]}\n{s[60:]}\n+{h}\n{q[:100]}\n{q[100:]}\n")
            else:
                fh.write(f"@{h}\n{s}\n+\n{q}\n")
        fh.write("\n")
    gz = tmp_path / "b.fastq.gz"
    with open(plain, "rb") as fi, gzip.open(gz, "wb") as fo:
        fo.write(fi.read())
    want = list(fastx.parse(str(plain), "fastq"))
    assert len(want) == len(recs)
    for path in (plain, gz):
        fq = C.c_void_p()
        _lib.check(L.mb_fastq_load(os.fsencode(str(path)), C.byref(fq)))
        n = L.mb_fastq_n(fq)
        assert n == len(want) and L.mb_fastq_ids_unique(fq) == 1
        offp = C.POINTER(C.c_int64)()
        catp = L.mb_fastq_seqs(fq, C.byref(offp))
        off = np.ctypeslib.as_array(offp, shape=(n + 1,))
        cat = np.ctypeslib.as_array(catp, shape=(int(off[-1]),)).tobytes() if off[-1] else b""
        for i, r in enumerate(want):
            assert cat[off[i]:off[i + 1]].decode() == str(r.seq)
            ln, il = C.c_int64(), C.c_int32()
            hp = L.mb_fastq_header(fq, i, C.byref(ln), C.byref(il))
            head = C.string_at(hp, ln.value).decode()
            assert head == r.description and head[:il.value] == r.id
        # routing: every third read mapped with a new id, some ambiguous, some focus
        dest = np.array([(1 if i % 3 == 0 else 2 if i % 7 == 1 else 0) for i in range(n)], np.int8)
        new_ids = (C.c_char_p * n)()
        for i in range(n):
            if dest[i] == 1:
                new_ids[i] = (b"read%d" % i) if i % 9 == 0 else b"Escherichia_coli"    # i % 9 == 0: new id == old id
        focus = np.array([1 if (dest[i] == 1 and i % 2 == 0) else 0 for i in range(n)], np.uint8)
        out = tmp_path / ("out_" + path.name)
        out.mkdir()
        paths = [os.fsencode(str(out / k)) for k in ("mapped", "unmapped", "ambiguous", "focus")]
        for _ in range(2):   # append mode: two rounds
            _lib.check(L.mb_fastq_route(fq, dest.ctypes.data_as(C.c_void_p), C.cast(new_ids, C.c_void_p), focus.ctypes.data_as(C.c_void_p), *paths))
        L.mb_fastq_free(fq)
        exp = {"mapped": "", "unmapped": "", "ambiguous": "", "focus": ""}
        for i, r in enumerate(want):
            if dest[i] == 1:
                if focus[i]:
                    exp["focus"] += r.format_fastq()
                r2 = fastx.SeqRecord(new_ids[i].decode(), r.description, str(r.seq), r.qual)
                exp["mapped"] += r2.format_fastq()
            else:
                exp["ambiguous" if dest[i] == 2 else "unmapped"] += r.format_fastq()
        for k, v in exp.items():
            assert open(out / k).read() == v + v, k
    # duplicate ids are detected
    dup = tmp_path / "d.fastq"
    dup.write_text("@x 1\nACGT\n+\nIIII\n@y\nAC\n+\nII\n@x 2\nGG\n+\nII\n")
    fq = C.c_void_p()
    _lib.check(L.mb_fastq_load(os.fsencode(str(dup)), C.byref(fq)))
    assert L.mb_fastq_n(fq) == 3 and L.mb_fastq_ids_unique(fq) == 0
    L.mb_fastq_free(fq)
    assert L.mb_fastq_load(os.fsencode(str(tmp_path / "missing.fastq")), C.byref(fq)) != 0


def test_host_logf_twin_equals_libm(oracle):
    """The MAPQ formula's logf (glue.cuh: mb_logf, host twin of the device routine; mm_set_mapq, hit.c) against the libm the
    oracle links, bit for bit: every float in [1, 2^24) -- all integer-valued arguments (scores, n_sub + 1) and the
    dp_max / match_sc ratios -- and a stride through all positive normals.  The GPU suite sweeps all 2^31 - 2^24 on the device."""
    import ctypes as C
    from monica_b200 import _lib
    L = _lib.lib()

    def sweep(first, n):
        exp = oracle.logf_range(first, n)
        nb, f = C.c_int64(0), C.c_uint32(0)
        _lib.check(L.mb_logf_sweep(-1, first, n, exp.ctypes.data, C.byref(nb), C.byref(f)))
        assert nb.value == 0, f"{nb.value} floats differ from libm in [{first:#x}, +{n}), the first at bits {f.value:#010x}"
        return exp

    for e in range(24):                              # the binades [2^e, 2^(e+1)), e = 0 .. 23
        exp = sweep(0x3f800000 + (e << 23), 1 << 23)
    assert exp[0] == np.float32(np.log(np.float64(2.0 ** 23)))
    for first in range(0x00800000, 0x7f800000, 1 << 26):
        sweep(first, 1 << 16)
    # a wrong expectation is reported, with its position
    exp = oracle.logf_range(0x40000000, 1024)
    exp[5] = np.nextafter(exp[5], np.float32(10))
    nb, f = C.c_int64(0), C.c_uint32(0)
    _lib.check(L.mb_logf_sweep(-1, 0x40000000, 1024, exp.ctypes.data, C.byref(nb), C.byref(f)))
    assert (nb.value, f.value) == (1, 0x40000005)


def _nt4_table():
    t = np.full(256, 4, np.uint8)
    for ch, v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
        t[ord(ch)] = t[ord(ch.lower())] = v
    return t


def _unpack_packed(L, pk, total):
    """numpy statement of what k_unpack_nt4 + k_apply_amb produce on the device from a packed batch."""
    import ctypes as C
    nw, ivp, niv = C.c_int64(0), C.c_void_p(), C.c_int64(0)
    wp = L.mb_packed_words(pk, C.byref(nw), C.byref(ivp), C.byref(niv))
    words = np.ctypeslib.as_array(C.cast(wp, C.POINTER(C.c_uint32)), shape=(nw.value,)).copy()
    iv = (np.ctypeslib.as_array(C.cast(ivp, C.POINTER(C.c_int64)), shape=(2 * niv.value,)).copy().reshape(-1, 2)
          if niv.value else np.zeros((0, 2), np.int64))
    assert nw.value == (total + 15) // 16 + 2 and words[-1] == 0 and words[-2] == 0
    codes = ((words[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & 3).astype(np.uint8).reshape(-1)[:total]
    for s, n in iv:
        assert np.all(codes[s:s + n] == 0)          # ambiguous bases travel as 0 in the words
        codes[s:s + n] = 4
    return codes, iv


@pytest.mark.parametrize("n_threads", [1, 5])
def test_read_packer_equals_nt4_table(n_threads):
    """mb_reads_pack (pack.cuh): 2-bit words + runs of ambiguous bases must expand to exactly the nt4 codes of the ASCII
    (minimap2's seq_nt4_table: A/a 0, C/c 1, G/g 2, T/t/U/u 3, everything else 4) -- the bytes mb_map_batch's own encode
    kernel produces -- for mixed case, U, IUPAC letters, arbitrary bytes, runs of N across word and thread boundaries,
    lengths that are not multiples of 16, and the empty batch."""
    import ctypes as C
    from monica_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(5)
    tab = _nt4_table()
    for total in (0, 1, 15, 16, 17, 4099, (3 << 20) + 7):
        cat = np.frombuffer(b"ACGTacgtUuNnRYKM-*", np.uint8)[rng.integers(0, 8 if total > 100000 else 18, total)].copy()
        if total > 100000:   # a few long and short runs of other characters, one across each thread cut, one at the very end
            for s, n in ((5, 1), (31, 2), (1000, 40), (total // 5 - 3, 9), (2 * (total // 5) - 16, 33), (total - 4, 4)):
                cat[s:s + n] = ord("N")
            cat[77] = 0; cat[78] = 255; cat[79] = ord("@")
        cuts = np.sort(rng.integers(0, total + 1, 6)) if total else np.zeros(0, np.int64)
        off = np.concatenate([[0], cuts, [total]]).astype(np.int64)
        pk = C.c_void_p()
        _lib.check(L.mb_reads_pack(cat.ctypes.data, off.ctypes.data, len(off) - 1, n_threads, C.byref(pk)))
        try:
            codes, iv = _unpack_packed(L, pk, total)
            want = tab[cat]
            assert np.array_equal(codes, want)
            # maximal, ascending, non-touching runs == the runs of code 4
            amb = np.flatnonzero(np.diff(np.concatenate([[0], (want == 4).astype(np.int8), [0]])))
            assert np.array_equal(iv.reshape(-1), np.stack([amb[0::2], amb[1::2] - amb[0::2]], 1).reshape(-1))
            assert L.mb_packed_upload_bytes(pk) == ((total + 15) // 16 + 2) * 4 + 16 * len(iv) + 8 * len(off)
        finally:
            L.mb_packed_free(pk)
    bad = np.array([1, 5], np.int64)
    pk = C.c_void_p()
    assert L.mb_reads_pack(cat.ctypes.data, bad.ctypes.data, 1, 1, C.byref(pk)) == -1   # MB_ERR_ARG: offsets must start at 0


_ODD_FASTQ = {
    "empty": b"",
    "only_newlines": b"\n\n\n",
    "no_trailing_newline": b"@r1\nACGT\n+\nIIII",
    "crlf": b"@r1 d\r\nACGT\r\n+\r\nIIII\r\n@r2\r\nGG\r\n+\r\nII\r\n",
    "quality_starts_with_at": b"@r1\nACGT\n+\n@III\n@r2\nAC\n+\n@I\n",
    "empty_sequence": b"@r1\n\n+\n\n@r2\nAC\n+\nII\n",
    "tab_in_title": b"@r1\tx y\nACGT\n+\nIIII\n",
    "title_ends_in_blanks": b"@r1 d  \nACGT\n+\nIIII\n@r2\t\nAC\n+\nII\n",
    "lower_case_and_iupac": b"@r1\nacgtnRYKM\n+\nIIIIIIIII\n",
    "blank_title": b"@   \nACGT\n+\nIIII\n@\nAC\n+\nII\n",
    "blank_lines_between": b"@r1\nACGT\n+\nIIII\n\n\n@r2\nAC\n+\nII\n",
    "long_title": b"@" + b"x" * 100000 + b" c\nACGT\n+\nIIII\n",
}


_MALFORMED_FASTQ = {
    "cut_after_sequence": b"@r0\nAC\n+\nII\n@r1\nACGT\n",
    "cut_after_plus": b"@r1\nACGT\n+\n",
    "quality_too_short": b"@r1\nACGT\n+\nII\n",
    "quality_too_long": b"@r1\nACGT\n+\nIIIIIIII\n@r2\nAC\n+\nII\n",
    "header_only": b"@r1",
    "missing_plus": b"@r1\nACGT\nIIII\n@r2\nAC\n+\nII\n",
    "not_fastq": b"hello world\nfoo\n",
    "sequence_line_missing": b"@r0\n+\n\n",        # Bio.SeqIO: the line after the title is sequence whatever it starts with
}


def _native_fastq_records(path):
    import ctypes as C
    from monica_b200 import _lib
    L = _lib.lib()
    fq = C.c_void_p()
    rc = L.mb_fastq_load(os.fsencode(str(path)), C.byref(fq))
    if rc != 0:
        return rc, None
    n = L.mb_fastq_n(fq)
    offp = C.POINTER(C.c_int64)()
    catp = L.mb_fastq_seqs(fq, C.byref(offp))
    off = np.ctypeslib.as_array(offp, shape=(n + 1,))
    cat = np.ctypeslib.as_array(catp, shape=(int(off[-1]),)).tobytes() if n and off[-1] else b""
    recs = []
    for i in range(n):
        ln, il = C.c_int64(), C.c_int32()
        hp = L.mb_fastq_header(fq, i, C.byref(ln), C.byref(il))
        head = C.string_at(hp, ln.value).decode()
        recs.append((head[:il.value], head, cat[off[i]:off[i + 1]].decode()))
    # route everything to "unmapped": the file a Bio.SeqIO writer would produce from these records
    out = str(path) + ".unmapped"
    dest = np.zeros(max(n, 1), np.int8)
    _lib.check(L.mb_fastq_route(fq, dest.ctypes.data_as(C.c_void_p), None, None, None, os.fsencode(out), None, None))
    L.mb_fastq_free(fq)
    return 0, (recs, open(out, "rb").read() if os.path.exists(out) else b"")


@pytest.mark.parametrize("case", sorted(_ODD_FASTQ))
def test_native_fastq_loader_on_odd_input_equals_the_seqio_mirror(tmp_path, case):
    """Cut-off, CRLF, blank-line, tabbed and blank-padded title lines: mb_fastq_load sees the records monica_b200.fastx
    (Bio.SeqIO's rules: description = right-stripped title line, id = its first word) sees, and writes them back the same."""
    from monica_b200 import fastx
    p = tmp_path / "x.fastq"
    p.write_bytes(_ODD_FASTQ[case])
    want = list(fastx.parse(str(p), "fastq"))
    rc, got = _native_fastq_records(p)
    assert rc == 0
    recs, written = got
    assert recs == [(r.id, r.description, str(r.seq)) for r in want]
    assert written.decode() == "".join(r.format_fastq() for r in want)


@pytest.mark.parametrize("case", sorted(_MALFORMED_FASTQ))
def test_malformed_fastq_fails_like_seqio(tmp_path, case):
    """Bio.SeqIO (the reference's SeqIO.parse, aligner.py:191,212) raises ValueError for a record without its '+' line or with
    unequal sequence / quality lengths -- what a cut-off file looks like.  The native loader answers MB_ERR_IO with a
    'malformed FASTQ' message, the Python mirror raises ValueError, and aligner()'s whole-file route turns the former into
    the latter."""
    import ctypes as C
    from monica_b200 import _lib, fastx, aligner as mine
    p = tmp_path / "x.fastq"
    p.write_bytes(_MALFORMED_FASTQ[case])
    with pytest.raises(ValueError):
        list(fastx.parse(str(p), "fastq"))
    L = _lib.lib()
    fq = C.c_void_p()
    assert L.mb_fastq_load(os.fsencode(str(p)), C.byref(fq)) == -3
    assert L.mb_last_error().decode().startswith("malformed FASTQ") and "x.fastq" in L.mb_last_error().decode()
    for sub in ("hits", "mapped", "unmapped", "ambiguous"):
        (tmp_path / sub).mkdir()

    class NeverMapped:          # a batch-capable index: the whole-file route is taken and must fail before it maps anything
        seq_names = []

        def map_batch(self, *a, **k):
            raise AssertionError("mapped a malformed file")

        count = map_batch

    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        with pytest.raises(ValueError):
            mine.aligner("x.fastq", "x", NeverMapped(), mode="basic", hits_folder=str(tmp_path / "hits"), mapping_quality=60,
                         mapped_folder=str(tmp_path / "mapped"), unmapped_folder=str(tmp_path / "unmapped"),
                         ambiguous_folder=str(tmp_path / "ambiguous"), last_index=True)
    finally:
        os.chdir(cwd)
    assert p.exists()           # the input is not consumed


def test_fastx_title_rule_is_seqio_s():
    """Bio.SeqIO.QualityIO: description = title line right-stripped (inner tabs kept), id = first whitespace-separated word;
    mappy.fastx_read (kseq): name up to the first blank or tab, the rest is the comment."""
    import io
    from monica_b200 import fastx
    txt = "@r1\tx y  \nACGT\n+\nIIII\n"
    (r,) = list(fastx.parse(io.StringIO(txt), "fastq"))
    assert (r.id, r.description) == ("r1", "r1\tx y")
    assert r.format_fastq() == "@r1\tx y\nACGT\n+\nIIII\n"
    r.id = "Species_1"
    assert r.format_fastq() == "@Species_1 r1\tx y\nACGT\n+\nIIII\n"
    assert list(fastx.parse_fastx(io.StringIO(txt))) == [("r1", "x y  ", "ACGT", "IIII")]


def test_native_loaders_refuse_a_cut_off_gzip(tmp_path):
    """zlib hands out what it could inflate of a truncated .gz and then reports a plain end of file; the error is only in
    gzerror().  A FASTQ cut in half must not be mapped and counted as if it were whole."""
    import ctypes as C
    import gzip
    from monica_b200 import _lib
    L = _lib.lib()
    whole = gzip.compress(b"".join(b"@r%d\nACGTACGTAC\n+\nIIIIIIIIII\n" % i for i in range(3000)))
    good, cut, junk = tmp_path / "good.fastq.gz", tmp_path / "cut.fastq.gz", tmp_path / "junk.fastq.gz"
    good.write_bytes(whole)
    cut.write_bytes(whole[:len(whole) // 2])
    junk.write_bytes(b"\x1f\x8b" + b"\0" * 40)
    recs = [b"@r%d\nACGTACGTAC\n+\nIIIIIIIIII\n" % i for i in range(3000)]
    multi = tmp_path / "multi.fastq.gz"      # bgzip-style: several members, an empty one at the end, zero padding behind it
    multi.write_bytes(b"".join(gzip.compress(b"".join(recs[i:i + 500])) for i in range(0, 3000, 500)) + gzip.compress(b"") + b"\0" * 64)
    for ok in (good, multi):
        fq = C.c_void_p()
        assert L.mb_fastq_load(os.fsencode(str(ok)), C.byref(fq)) == 0 and L.mb_fastq_n(fq) == 3000
        L.mb_fastq_free(fq)
    for bad in (cut, junk):
        fq = C.c_void_p()
        assert L.mb_fastq_load(os.fsencode(str(bad)), C.byref(fq)) == -3          # MB_ERR_IO
        assert os.path.basename(str(bad)).encode() in L.mb_last_error()


def test_native_fastq_loader_large_file_with_one_odd_record(tmp_path):
    """A file large enough for the multi-threaded four-line parser (>= 8 MB) with one blank-padded title in the middle: the
    strict parser hands the file to the general one, and every record still equals the SeqIO mirror's."""
    from monica_b200 import fastx
    rng = np.random.default_rng(5)
    seq = "".join(rng.choice(list("ACGT"), 3000))
    p = tmp_path / "big.fastq"
    with open(p, "w") as fh:
        for i in range(1500):
            title = f"read{i} ch={i % 512}" + ("  " if i == 777 else "")
            s = seq[i % 100:] + seq[:i % 100]
            fh.write(f"@{title}\n{s}\n+\n{'I' * len(s)}\n")
    assert os.path.getsize(p) > (8 << 20)
    want = list(fastx.parse(str(p), "fastq"))
    rc, got = _native_fastq_records(p)
    assert rc == 0
    recs, written = got
    assert recs == [(r.id, r.description, str(r.seq)) for r in want] and recs[777][1] == "read777 ch=265"
    assert written.decode() == "".join(r.format_fastq() for r in want)


def _twin_batch_aligner_class():
    """tests/standin.OracleAligner plus the two batch calls of monica_b200.mappy_shim.Aligner that aligner()'s whole-file
    route uses (map_batch, count), served by the oracle behind the C ABI (oracle/mm2o_abi.c) -- so the host logic of that
    route (native FASTQ ingest, vectorised tally, native routed writers) runs on the CPU."""
    import ctypes as C
    import standin
    from test_shard_cpu import _TwinAligner

    class Hits:
        pass

    class TwinBatchAligner(standin.OracleAligner):
        calls = 0

        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            self._twin = None
            if self._idx is not None:
                names, seqs = self._names_seqs
                self._twin = _TwinAligner(names, [np.frombuffer(s if isinstance(s, bytes) else bytes(s), np.uint8) if not isinstance(s, np.ndarray) else s
                                                  for s in seqs])

        def map_batch(self, reads=None, cat=None, off=None, cigars=True):
            type(self).calls += 1
            L = self._twin.L
            if reads is not None:      # the shim's other input form: a list of sequences
                arrs = [np.frombuffer(r.encode() if isinstance(r, str) else bytes(r), np.uint8) for r in reads]
                off = np.concatenate([[0], np.cumsum([len(a) for a in arrs])]).astype(np.int64)
                cat = np.concatenate(arrs) if arrs else np.zeros(0, np.uint8)
            h = self._twin.map_batch(cat=cat, off=off, cigars=cigars)
            out = Hits()
            out.h, out.n_reads = h, len(off) - 1
            nh = L.mb_hits_n(h)
            for f in ("read_idx", "rid", "mapq", "mlen", "nm", "is_primary"):
                ptr = L.mb_hits_field(h, f.encode())
                setattr(out, f, np.ctypeslib.as_array(ptr, shape=(nh,)).copy() if nh else np.zeros(0, np.int32))
            return out

        def count(self, hits, mapq_min, mode):
            L = self._twin.L
            m = {"basic": 0, "query_length": 1, "matching": 2}.get(mode, -1)
            counts, ncls = np.zeros(self._twin.n_seq, np.int64), np.zeros(3, np.int64)
            rcls, rbest = np.zeros(max(hits.n_reads, 1), np.int8), np.zeros(max(hits.n_reads, 1), np.int64)
            assert L.mb_count(self._twin.idx, hits.h, mapq_min, m, counts.ctypes.data_as(C.c_void_p), ncls.ctypes.data_as(C.c_void_p),
                              rcls.ctypes.data_as(C.c_void_p), rbest.ctypes.data_as(C.c_void_p)) == 0
            return counts, ncls, rcls[:hits.n_reads], rbest[:hits.n_reads]

    return TwinBatchAligner


@pytest.mark.parametrize("mode,focus,overnight", [("basic", [], False), ("query_length", ["Species_1"], False), ("matching", ["Species_0"], True), (None, [], False)])
def test_whole_file_route_equals_unmodified_reference(ref_aligner, small_case, monkeypatch, mode, focus, overnight):
    """aligner()'s whole-file route (one native FASTQ load, one map_batch, mb_count's classes, vectorised tally, native routed
    writers -- the route every single-index run takes on the GPU) against the UNMODIFIED reference's per-record loop: same
    alignment dict, same bytes in every routed file.  The batch calls are served by the oracle twin of the C ABI."""
    from monica_b200 import aligner as mine
    names, seqs, reads = small_case
    cls = _twin_batch_aligner_class()
    fake = types.ModuleType("mappy")
    fake.Aligner = cls
    monkeypatch.setattr(mine, "mappy", fake)
    monkeypatch.setattr(mine, "LAST_BREAKDOWN", None)
    got = _run(mine, names, seqs, reads, mode, False, focus, True, mta={"overnight": overnight})
    assert cls.calls >= 1 and mine.LAST_BREAKDOWN is not None, "the whole-file route was not taken"
    want = _run(ref_aligner, names, seqs, reads, mode, False, focus, False, mta={"overnight": overnight})
    assert got == want
    if mode is not None:
        assert any(v for v in got["alignment"].values())


def test_duplicate_read_ids_leave_the_whole_file_route(ref_aligner, small_case, tmp_path, monkeypatch):
    """Two records with one id share an entry of the reference's per-read dict (their hits are merged before best_hit): the
    whole-file route must notice (mb_fastq_ids_unique) and hand the file to the per-record path, whose output equals the
    unmodified reference's."""
    import standin
    from monica_b200 import aligner as mine
    names, seqs, reads = small_case
    cls = _twin_batch_aligner_class()
    picked = reads[20:28]
    outs = []
    for mod, idx in ((mine, cls(names=names, seqs=[s.tobytes() for s in seqs])),
                     (ref_aligner, standin.OracleAligner(names=names, seqs=[s.tobytes() for s in seqs]))):
        d = tmp_path / mod.__name__.replace(".", "_")
        for sub in ("hits", "mapped", "unmapped", "ambiguous", "focus"):
            (d / sub).mkdir(parents=True)
        with open(d / "s.fastq", "wb") as fh:
            for i, r in enumerate(picked):
                fh.write(b"@dup%d\n" % (i // 2) + r.tobytes() + b"\n+\n" + b"I" * len(r) + b"\n")     # every id twice
        monkeypatch.setattr(mine, "LAST_BREAKDOWN", None)
        cwd = os.getcwd()
        os.chdir(d)
        try:
            res = mod.aligner("s.fastq", "s", idx, mode="query_length", hits_folder=str(d / "hits"), mapping_quality=60,
                              focus_species=["Species_1"], mapped_folder=str(d / "mapped"), unmapped_folder=str(d / "unmapped"),
                              ambiguous_folder=str(d / "ambiguous"), focus_folder=str(d / "focus"), last_index=True)
        finally:
            os.chdir(cwd)
        files = {sub: (d / sub / "s.fastq").read_bytes() if (d / sub / "s.fastq").exists() else None for sub in ("mapped", "unmapped", "ambiguous", "focus")}
        outs.append(({t: dict(c) for t, c in res[0].items()}, res[1], files, sorted(os.listdir(d / "hits")), (d / "s.fastq").exists()))
        if mod is mine:
            assert mine.LAST_BREAKDOWN is None, "a file with repeated ids went through the whole-file route"
    assert outs[0] == outs[1]
    assert outs[0][0], "nothing was counted"


@pytest.mark.parametrize("mode", ["query_length", "matching"])
def test_two_index_carry_over_with_batch_calls_equals_unmodified_reference(ref_aligner, small_case, monkeypatch, mode):
    """Two sequential indexes (aligner.py:184-188,196-203,218-223): the first index's kept hits travel in hits/<sample>_hits.pkl,
    the per-record path gets each index's hits from ONE map_batch call (the branch every GPU run takes); outputs equal the
    unmodified reference's per-read loop."""
    from monica_b200 import aligner as mine
    names, seqs, reads = small_case
    cls = _twin_batch_aligner_class()
    fake = types.ModuleType("mappy")
    fake.Aligner = cls
    monkeypatch.setattr(mine, "mappy", fake)
    got = _run(mine, names, seqs, reads, mode, True, ["Species_1"], True)
    assert cls.calls >= 4          # two samples x two indexes
    want = _run(ref_aligner, names, seqs, reads, mode, True, ["Species_1"], False)
    assert got == want and any(v for v in got["alignment"].values())
    assert got["query_files_left"] == [] and got["hits_left"] == []


def test_fastq_packer_equals_read_packer(tmp_path):
    """mb_fastq_pack (the packed batch straight from a natively loaded FASTQ, incl. multi-line and CRLF records and an empty
    file) holds the words, runs and upload size mb_reads_pack makes from the same sequences, and both expand to nt4 codes."""
    import ctypes as C
    from monica_b200 import _lib, fastx
    L = _lib.lib()
    rng = np.random.default_rng(17)
    tab = _nt4_table()
    for name, n_rec in (("some", 37), ("none", 0)):
        p = tmp_path / f"{name}.fastq"
        with open(p, "w", newline="") as fh:
            for i in range(n_rec):
                n = int(rng.integers(1, 900))
                s = "".join(rng.choice(list("ACGTACGTACGTACGTacgtNnRU"), n))
                if i % 6 == 0:
                    s = s[:n // 3] + "N" * (n // 3) + s[2 * (n // 3):]
                nl = "\r\n" if i % 5 == 3 else "\n"
                if i % 4 == 1 and n > 100:
                    fh.write(f"@r{i} c{nl}{s[:50]}{nl}{s[50:]}{nl}+{nl}{'I' * 50}{nl}{'I' * (len(s) - 50)}{nl}")
                else:
                    fh.write(f"@r{i}{nl}{s}{nl}+{nl}{'I' * len(s)}{nl}")
        want_seqs = [str(r.seq) for r in fastx.parse(str(p), "fastq")]
        assert len(want_seqs) == n_rec
        fq = C.c_void_p()
        _lib.check(L.mb_fastq_load(os.fsencode(str(p)), C.byref(fq)))
        pk_f, pk_r = C.c_void_p(), C.c_void_p()
        try:
            _lib.check(L.mb_fastq_pack(fq, 3, C.byref(pk_f)))
            cat = np.frombuffer("".join(want_seqs).encode(), np.uint8).copy() if n_rec else np.zeros(1, np.uint8)
            off = np.concatenate([[0], np.cumsum([len(s) for s in want_seqs])]).astype(np.int64)
            total = int(off[-1])
            _lib.check(L.mb_reads_pack(cat.ctypes.data, off.ctypes.data, n_rec, 2, C.byref(pk_r)))
            codes_f, iv_f = _unpack_packed(L, pk_f, total)
            codes_r, iv_r = _unpack_packed(L, pk_r, total)
            assert np.array_equal(codes_f, codes_r) and np.array_equal(iv_f, iv_r)
            assert np.array_equal(codes_f, tab[cat[:total]])
            assert L.mb_packed_upload_bytes(pk_f) == L.mb_packed_upload_bytes(pk_r)
        finally:
            if pk_f:
                L.mb_packed_free(pk_f)
            if pk_r:
                L.mb_packed_free(pk_r)
            L.mb_fastq_free(fq)


def test_multi_threaded_aligner_propagates_a_malformed_file_like_the_reference(ref_aligner, small_case, tmp_path, monkeypatch):
    """One cut-off FASTQ among the samples: pool.starmap re-raises the worker's ValueError in the reference
    (aligner.py:96-103); the mirror does the same on its whole-file route."""
    import standin
    from monica_b200 import aligner as mine, synth
    names, seqs, reads = small_case
    cls = _twin_batch_aligner_class()
    for mod, acls in ((mine, cls), (ref_aligner, standin.OracleAligner)):
        d = tmp_path / mod.__name__.replace(".", "_")
        (d / "q").mkdir(parents=True)
        (d / "db").mkdir()
        synth.write_fastq(str(d / "q" / "good.fastq"), reads[20:24])
        with open(d / "q" / "cut.fastq", "wb") as fh:
            fh.write(b"@r0\n" + reads[25].tobytes() + b"\n+\n" + b"I" * len(reads[25]) + b"\n@r1\n" + reads[26].tobytes()[:100] + b"\n")
        synth.write_fasta_gz(str(d / "db" / "database1.fna.gz"), names, seqs)
        fake = types.ModuleType("mappy")
        fake.Aligner = acls
        monkeypatch.setattr(mod, "mappy", fake)
        cwd = os.getcwd()
        try:
            kw = {"genomes_path": str(d / "markers")} if mod is mine else {}
            idx = sorted(mod.indexer(str(d / "db"), str(d / "idx"), **kw) or [os.path.join(str(d / "idx"), f) for f in os.listdir(d / "idx") if f.endswith(".mmi")])
            with pytest.raises(ValueError):
                mod.multi_threaded_aligner(str(d / "q"), idx, mode="basic", n_threads=2, output_folder=str(d))
        finally:
            os.chdir(cwd)
        assert (d / "q" / "cut.fastq").exists()
