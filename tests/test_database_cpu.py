"""the database builder mirror (monica_b200/database.py over mb_db_build) against the UNMODIFIED
reference monica/genomes/database.py, driven through a restatement of Biopython's FASTA reader/writer (tests/standin.py).
Host-side only: no device call."""
import gzip
import os
import pickle
import shutil
import tempfile

import numpy as np
import pytest

from conftest import have_reference


def _write_genome(path, rng, n_records, newline="\n", width=80, odd=False):
    with gzip.open(path, "wt", newline="") as fh:
        if odd:
            fh.write("; a comment line before the first record" + newline)
        for r in range(n_records):
            n = int(rng.integers(1, 700))
            seq = "".join(rng.choice(list("ACGTNacgt"), size=n))
            title = f"NZ_{rng.integers(1e6)}.{r} Some organism strain {r} chromosome, complete genome"
            if odd and r == 1:
                title = f"  lead{r}  spaced\ttitle  "
            if odd and r == 2:
                title = ""
            fh.write(">" + title + newline)
            for i in range(0, n, width):
                line = seq[i:i + width]
                if odd and i == 0:
                    line = line[:5] + " " + line[5:] + "  \t"
                fh.write(line + newline)
            if odd:
                fh.write(newline)


def _genomes(folder, rng, n):
    out = []
    for g in range(n):
        path = os.path.join(folder, f"GCF_{g:03d}.fna.gz")
        _write_genome(path, rng, int(rng.integers(1, 5)), newline="\r\n" if g % 3 == 1 else "\n", width=70 if g % 2 else 80, odd=(g == 2))
        out.append((path, (f"Species_{g}", f"GCF_{g:03d}.1")))
    return out


def _gunzip(path):
    with gzip.open(path, "rb") as fh:
        return fh.read()


def test_splitter_quirk_and_strict(tmp_path):
    from monica_b200 import database as mine
    genomes = []
    for i, size in enumerate([100, 100, 100, 500, 100, 100]):
        p = tmp_path / f"g{i}"
        p.write_bytes(b"x" * size)
        genomes.append((str(p), (f"S{i}", f"A{i}")))
    chunks = list(mine._genomes_splitter(genomes, max_chunk_size=250))
    names = [[g[1][1] for g in c] for c in chunks]
    # genome 2 closes the first chunk and is dropped (reference quirk, database.py:84-90); genome 3 is oversize and alone
    assert names == [["A0", "A1"], ["A3"], ["A4", "A5"]]
    strict = [[g[1][1] for g in c] for c in mine._genomes_splitter(genomes, max_chunk_size=250, strict=True)]
    assert strict == [["A0", "A1"], ["A3"], ["A2", "A4"], ["A5"]]
    with pytest.raises(TypeError):
        list(mine._genomes_splitter(genomes, max_chunk_size=None))


def test_builder_record_format(tmp_path):
    from monica_b200 import database as mine
    g = tmp_path / "a.fna.gz"
    with gzip.open(g, "wt") as fh:
        fh.write(">c1 first contig\nACGTACGTAC\nGG\n>Sp:ACC already renamed\n" + "A" * 130 + "\n>\nTT\n")
    lengths = mine.builder([(str(g), ("Sp", "ACC"))], str(tmp_path), ["database", ".fna.gz"], 7)
    assert lengths == {"ACC": 12 + 130 + 2}
    text = _gunzip(tmp_path / "database7.fna.gz").decode()
    assert text == (">Sp:ACC c1 first contig\nACGTACGTACGG\n>Sp:ACC already renamed\n" + "A" * 60 + "\n" + "A" * 60 + "\n" + "A" * 10 + "\n>Sp:ACC\nTT\n")
    with pytest.raises(Exception):
        mine.builder([(str(tmp_path / "missing.fna.gz"), ("S", "A"))], str(tmp_path), ["database", ".fna.gz"], 8)


@pytest.mark.parametrize("max_chunk", [10 ** 9, 900, 400])
def test_multi_threaded_builder_equals_unmodified_reference(max_chunk):
    if not have_reference():
        pytest.skip("/root/reference not present (GPU box)")
    import importlib
    import standin
    from monica_b200 import database as mine
    home = tempfile.mkdtemp(prefix="ref_db_home_")
    try:
        standin.install_reference_imports(home)
        ref = importlib.import_module("monica.genomes.database")
        rng = np.random.default_rng(5)
        ref_genomes_path = ref.GENOMES_PATH
        my_genomes_path = os.path.join(home, "mine", "genomes")
        os.makedirs(my_genomes_path)
        genomes_ref = _genomes(ref_genomes_path, rng, 7)
        genomes_mine = []
        for path, ids in genomes_ref:
            shutil.copy(path, my_genomes_path)
            genomes_mine.append((os.path.join(my_genomes_path, os.path.basename(path)), ids))
        for folder in (ref_genomes_path, my_genomes_path):
            with open(os.path.join(folder, "current_genomes_length.pkl"), "wb") as fh:
                pickle.dump({"OLD.1": 123}, fh)
        ref_chunks = [[g[1] for g in c] for c in ref._genomes_splitter(genomes_ref, max_chunk_size=max_chunk)]
        my_chunks = [[g[1] for g in c] for c in mine._genomes_splitter(genomes_mine, max_chunk_size=max_chunk)]
        assert ref_chunks == my_chunks
        ref_db = os.path.join(home, "ref_db")
        my_db = os.path.join(home, "my_db")
        os.makedirs(my_db)
        open(os.path.join(my_db, "database9.fna.gz"), "wb").close()   # stale database: must be removed
        r_path, r_len = ref.multi_threaded_builder(genomes=genomes_ref, max_chunk_size=max_chunk, databases_path=ref_db, keep_genomes=False, n_threads=2)
        m_path, m_len = mine.multi_threaded_builder(genomes=genomes_mine, max_chunk_size=max_chunk, databases_path=my_db, keep_genomes=False,
                                                    n_threads=2, genomes_path=my_genomes_path)
        assert (r_path, m_path) == (ref_db, my_db)
        assert r_len == m_len and "OLD.1" in m_len
        assert sorted(os.listdir(ref_db)) == sorted(os.listdir(my_db)) and len(os.listdir(my_db)) == len(ref_chunks)
        for name in os.listdir(ref_db):
            assert _gunzip(os.path.join(ref_db, name)) == _gunzip(os.path.join(my_db, name)), name
        assert sorted(os.listdir(ref_genomes_path)) == sorted(os.listdir(my_genomes_path))   # genomes deleted, marker + pickle written
        assert "database_created" in os.listdir(my_genomes_path)
        with open(os.path.join(my_genomes_path, "current_genomes_length.pkl"), "rb") as fh:
            assert pickle.load(fh) == m_len
    finally:
        shutil.rmtree(home, ignore_errors=True)


def test_builder_refuses_a_cut_off_genome_file(tmp_path):
    """A genome download cut short (truncated .fna.gz) must fail the chunk, as gzip.open + SeqIO.parse does in the reference
    (EOFError), not become a shorter genome in the database."""
    from monica_b200 import _lib, database
    rng = np.random.default_rng(9)
    good = str(tmp_path / "GCF_good.fna.gz")
    _write_genome(good, rng, 40)
    raw = open(good, "rb").read()
    cut = str(tmp_path / "GCF_cut.fna.gz")
    with open(cut, "wb") as fh:
        fh.write(raw[:len(raw) // 2])
    out = tmp_path / "db"
    out.mkdir()
    lens = database.builder([(good, ("Species_0", "GCF_good.1"))], str(out), ["database", ".fna.gz"], 0)
    assert lens["GCF_good.1"] > 0
    with pytest.raises(_lib.MonicaB200Error) as ei:
        database.builder([(good, ("Species_0", "GCF_good.1")), (cut, ("Species_1", "GCF_cut.1"))], str(out), ["database", ".fna.gz"], 1)
    assert ei.value.code == -3 and "GCF_cut" in str(ei.value)
