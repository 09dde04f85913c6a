"""TEST INFRASTRUCTURE: oracle-backed stand-ins for the two third-party modules the reference aligner imports
(`mappy`, `Bio.SeqIO`), so /root/reference/monica/genomes/aligner.py can be imported and run UNMODIFIED in this
container (neither module is installed here).  Never imported by monica_b200/.

`mappy.Aligner(fn_idx_in=<fasta.gz>, fn_idx_out=<mmi>)` builds an oracle index and leaves a small stub at <mmi> that
records where the FASTA is; `mappy.Aligner(fn_idx_in=<mmi>)` follows the stub.  `.map()` yields objects with the five
fields the reference reads.
"""
from __future__ import annotations

import gzip
import os
import sys
import types

from monica_b200 import fastx
from oracle import oracle as O


def read_fasta(path):
    opener = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    names, seqs = [], []
    with opener(path, "rt") as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line[0] == ">":
                names.append(line[1:].split()[0])
                seqs.append([])
            else:
                seqs[-1].append(line)
    return names, ["".join(s).encode() for s in seqs]


class Hit:
    __slots__ = ("ctg", "mapq", "is_primary", "NM", "mlen", "blen", "r_st", "r_en", "q_st", "q_en", "strand", "score", "cigar")


class OracleAligner:
    STUB = b"ORACLE-STUB\n"

    def __init__(self, fn_idx_in=None, preset=None, best_n=None, fn_idx_out=None, names=None, seqs=None, **kw):
        self._idx = None
        try:
            if names is not None:
                pass
            elif fn_idx_in is None:
                return
            else:
                with open(fn_idx_in, "rb") as fh:
                    head = fh.read(len(self.STUB))
                if head == self.STUB:
                    with open(fn_idx_in, "rb") as fh:
                        fasta = fh.read()[len(self.STUB):].decode().strip()
                else:
                    fasta = fn_idx_in
                names, seqs = read_fasta(fasta)
                if fn_idx_out is not None:
                    with open(fn_idx_out, "wb") as fh:
                        fh.write(self.STUB + os.path.abspath(fasta).encode() + b"\n")
            if not names:
                return
            self._idx = O.Index(names, seqs)
            self.seq_names = list(names)
            self._names_seqs = (list(names), list(seqs))
        except OSError:
            self._idx = None

    def __bool__(self):
        return self._idx is not None

    def map(self, seq, **kw):
        hits, _ = self._idx.map(seq)
        for h in hits:
            a = Hit()
            a.ctg = self.seq_names[h["rid"]]
            a.mapq, a.is_primary, a.NM, a.mlen, a.blen = h["mapq"], bool(h["is_primary"]), h["nm"], h["mlen"], h["blen"]
            a.r_st, a.r_en, a.q_st, a.q_en, a.strand = h["rs"], h["re"], h["qs"], h["qe"], -1 if h["rev"] else 1
            a.score, a.cigar = h["dp_max"], h["cigar"]
            yield a


class FastaRecord:
    """What Bio.SeqIO.parse(handle, 'fasta') yields, reduced to what database.py:60-64 touches."""
    def __init__(self, title, seq):
        self.id = self.name = title.split(None, 1)[0] if title.split(None, 1) else ""
        self.description = title
        self.seq = seq

    def __len__(self):
        return len(self.seq)


def bio_fasta_parse(handle):
    """Restatement of Bio.SeqIO.FastaIO.SimpleFastaParser (lines before the first '>' skipped; title = line[1:].rstrip();
    sequence lines rstrip()ped, joined, blanks and CRs removed)."""
    title, lines = None, []
    for line in handle:
        if line[:1] == ">":
            if title is not None:
                yield FastaRecord(title, "".join(lines).replace(" ", "").replace("\r", ""))
            title, lines = line[1:].rstrip(), []
        elif title is not None:
            lines.append(line.rstrip())
    if title is not None:
        yield FastaRecord(title, "".join(lines).replace(" ", "").replace("\r", ""))


def bio_fasta_write(rec, handle):
    """Restatement of Bio.SeqIO.FastaIO.as_fasta: title rule + 60-column lines."""
    ident, desc = rec.id.replace("\n", " ").replace("\r", " "), rec.description.replace("\n", " ").replace("\r", " ")
    if desc and desc.split(None, 1)[0] == ident:
        title = desc
    elif desc:
        title = f"{ident} {desc}"
    else:
        title = ident
    handle.write(f">{title}\n")
    for i in range(0, len(rec.seq), 60):
        handle.write(rec.seq[i:i + 60] + "\n")
    return 1


def _seqio_parse(handle, fmt="fastq"):
    return bio_fasta_parse(handle) if fmt == "fasta" and not isinstance(handle, (str, bytes, os.PathLike)) else fastx.parse(handle, fmt)


def _seqio_write(records, handle, fmt="fastq"):
    return bio_fasta_write(records, handle) if fmt == "fasta" else fastx.write(records, handle, fmt)


def install_reference_imports(home: str):
    """Make `import monica.genomes.aligner` from /root/reference work: fake mappy + Bio.SeqIO, and ~/.monica/.root."""
    os.makedirs(os.path.join(home, ".monica"), exist_ok=True)
    with open(os.path.join(home, ".monica", ".root"), "w") as fh:
        fh.write(os.path.join(home, "monica_root"))
    os.makedirs(os.path.join(home, "monica_root", "genomes"), exist_ok=True)
    os.environ["HOME"] = home
    m = types.ModuleType("mappy")
    m.Aligner = OracleAligner
    sys.modules["mappy"] = m
    bio = types.ModuleType("Bio")
    seqio = types.ModuleType("Bio.SeqIO")
    seqio.parse = _seqio_parse
    seqio.write = _seqio_write
    bio.SeqIO = seqio
    sys.modules["Bio"] = bio
    sys.modules["Bio.SeqIO"] = seqio
    for name in ("wget", "ete3"):
        if name not in sys.modules:
            stub = types.ModuleType(name)
            stub.NCBITaxa = object
            sys.modules[name] = stub
    if "/root/reference" not in sys.path:
        sys.path.insert(0, "/root/reference")
    for k in [k for k in sys.modules if k == "monica" or k.startswith("monica.")]:
        del sys.modules[k]
    import importlib
    return importlib.import_module("monica.genomes.aligner")
