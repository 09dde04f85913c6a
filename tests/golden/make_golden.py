"""Generate the committed golden vectors.  Run in the build container (needs /root/reference for part 2):

    python tests/golden/make_golden.py

1. small_case.npz   -- genomes + reads of conftest.small_case with the ORACLE's hits (all mappy-visible fields + dp_max)
                       and CIGARs.  GPU tests compare the CUDA path against these on the box, where neither the
                       reference nor its dependencies exist.  Provenance: this repo's CPU restatement (parity unpinned).
2. ref_aligner.json -- outputs of the UNMODIFIED /root/reference/monica/genomes/aligner.py (multi_threaded_aligner) run
                       over that data with tests/standin.py supplying `mappy` (oracle-backed) and `Bio.SeqIO` (monica_b200.fastx):
                       the returned alignment dict for the three counting modes and SHA-256 of every routed FASTQ file.
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from monica_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

FIELDS = ["rid", "rev", "qs", "qe", "rs", "re", "mapq", "mlen", "blen", "nm", "dp_max", "dp_max2", "score", "score0", "cnt",
          "subsc", "n_sub", "id", "parent", "is_primary", "sam_pri", "n_cigar"]


def small_case():
    names, seqs = synth.make_genomes(11, 3, 60000, strain_frac=0.34)
    reads, truth = synth.simulate_reads(12, seqs, 30, 2500, 0.10, junk_frac=0.05)
    reads = synth.edge_reads(13, seqs) + reads
    return names, seqs, reads


def part1():
    names, seqs, reads = small_case()
    idx = O.Index(names, seqs)
    gcat, goff = synth.concat_reads(seqs)
    rcat, roff = synth.concat_reads(reads)
    rows, hit_off, cig, cig_off = [], [0], [], [0]
    for r in reads:
        hits, _ = idx.map(r)
        for h in hits:
            rows.append([h[f] for f in FIELDS])
            cig.append(h["cigar"])
            cig_off.append(cig_off[-1] + len(h["cigar"]))
        hit_off.append(len(rows))
    np.savez_compressed(
        os.path.join(HERE, "small_case.npz"),
        names=np.array(names), genome_cat=gcat, genome_off=goff, read_cat=rcat, read_off=roff,
        mid_occ=np.int64(idx.mid_occ), hit_fields=np.array(FIELDS), hits=np.array(rows, dtype=np.int32).reshape(-1, len(FIELDS)),
        hit_off=np.array(hit_off, dtype=np.int64), cigar=np.concatenate(cig).astype(np.uint32) if cig else np.zeros(0, np.uint32),
        cigar_off=np.array(cig_off, dtype=np.int64), sketch_read3=O.sketch(reads[3]))
    print("small_case.npz:", len(reads), "reads,", len(rows), "hits")


def sha(path):
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def run_aligner(mod, workdir, names, seqs, reads, mode, two_indexes, focus, loader_kw):
    q = os.path.join(workdir, "query")
    out = os.path.join(workdir, "out")
    db = os.path.join(workdir, "db")
    for d in (q, out, db):
        os.makedirs(d)
    # two samples; sample b repeats a read id to exercise the per-record merge
    synth.write_fastq(os.path.join(q, "sampleA.fastq"), reads[: len(reads) // 2], prefix="a")
    synth.write_fastq(os.path.join(q, "sampleB.x.fastq"), reads[len(reads) // 2:], prefix="b")
    if two_indexes:
        synth.write_fasta_gz(os.path.join(db, "database1.fna.gz"), names[:2], seqs[:2])
        synth.write_fasta_gz(os.path.join(db, "database2.fna.gz"), names[2:], seqs[2:])
    else:
        synth.write_fasta_gz(os.path.join(db, "database1.fna.gz"), names, seqs)
    cwd = os.getcwd()
    try:
        idx_paths = sorted(mod.indexer(db, os.path.join(workdir, "idx"), **loader_kw.get("indexer", {})))
        res = mod.multi_threaded_aligner(q, idx_paths, mode=mode, n_threads=2, focus_species=focus, output_folder=out,
                                         **loader_kw.get("mta", {}))
    finally:
        os.chdir(cwd)
    files = {}
    for sub in ("mapped", "unmapped", "ambiguous", "focus"):
        d = os.path.join(q, sub)
        if os.path.isdir(d):
            for f in sorted(os.listdir(d)):
                files[f"{sub}/{f}"] = sha(os.path.join(d, f))
    left = sorted(f for f in os.listdir(q) if os.path.isfile(os.path.join(q, f)))
    hits_left = sorted(os.listdir(os.path.join(q, "hits")))
    plain = {s: {t: dict(c) for t, c in v.items()} for s, v in res.items()}
    return dict(alignment=plain, files=files, query_files_left=left, hits_left=hits_left)


def part2():
    import standin
    home = tempfile.mkdtemp(prefix="golden_home_")
    ref = standin.install_reference_imports(home)
    # the reference's indexer writes its marker files under fetcher.GENOMES_PATH
    os.makedirs(ref.GENOMES_PATH, exist_ok=True)
    names, seqs, reads = small_case()
    cases = {}
    for mode in ("basic", "query_length", "matching", None):
        for two in (False, True):
            wd = tempfile.mkdtemp(prefix="golden_run_")
            key = f"mode={mode},two_indexes={two}"
            cases[key] = run_aligner(ref, wd, names, seqs, reads, mode, two, ["Species_1"] if two else [], {})
            shutil.rmtree(wd)
            print(key, {s: sum(sum(c.values()) for c in v.values()) for s, v in cases[key]["alignment"].items()})
    with open(os.path.join(HERE, "ref_aligner.json"), "w") as fh:
        json.dump(cases, fh, indent=1, sort_keys=True)
    shutil.rmtree(home)


if __name__ == "__main__":
    O.build()
    part1()
    if os.path.isdir("/root/reference/monica"):
        part2()
    else:
        print("no /root/reference: ref_aligner.json not regenerated")
