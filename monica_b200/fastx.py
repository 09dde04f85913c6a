"""FASTQ/FASTA reading and writing for the aligner (replaces Bio.SeqIO on monica's path).

The reference iterates `SeqIO.parse(sample, 'fastq')` and appends records with `SeqIO.write(rec, handle, 'fastq')`
(/root/reference/monica/genomes/aligner.py:191,212,232,236,243,265).  Biopython's FASTQ writer emits
'@<id> <description-without-id>' -- i.e. '@' + description when the description starts with the id, else
'@' + id + ' ' + description -- then the sequence on one line, '+', and the quality string.  `SeqRecord`/`parse`/
`write` below keep exactly that behaviour so the routed FASTQ files are byte-identical.
"""
from __future__ import annotations

import gzip
import io
import os


def _open(fn):
    if hasattr(fn, "read"):
        return fn, False
    fn = os.fspath(fn)
    with open(fn, "rb") as fh:
        magic = fh.read(2)
    if magic == b"\x1f\x8b":
        return io.TextIOWrapper(gzip.open(fn, "rb")), True
    return open(fn, "r"), True


def _split_title(head):
    """kseq-style split of a title line: the name ends at the first blank or tab, the rest is the comment."""
    name, _, comment = head.partition(" ")
    if "\t" in name:
        name, _, c2 = name.partition("\t")
        comment = c2 + (" " + comment if comment else "")
    return name, comment


def parse_fastx(fn):
    """Yield (name, comment, seq, qual) with qual None for FASTA."""
    for head, seq, qual in _iter_records(fn):
        name, comment = _split_title(head)
        yield name, comment, seq, qual


def _iter_records(fn):
    """Yield (title line without its marker and line end, seq, qual) with qual None for FASTA."""
    fh, close = _open(fn)
    try:
        line = fh.readline()
        while line:
            line = line.rstrip("\r\n")
            if not line:
                line = fh.readline()
                continue
            if line[0] == "@":
                head = line[1:]
                seq = fh.readline().rstrip("\r\n")
                plus = fh.readline()
                # multi-line FASTQ: keep reading sequence lines until '+'
                while plus and not plus.startswith("+"):
                    seq += plus.rstrip("\r\n")
                    plus = fh.readline()
                if not plus:           # Bio.SeqIO.QualityIO.FastqGeneralIterator raises ValueError for both
                    raise ValueError("End of file without quality information.")
                qual = ""
                while len(qual) < len(seq):
                    q = fh.readline()
                    if not q:
                        break
                    qual += q.rstrip("\r\n")
                if len(qual) != len(seq):
                    raise ValueError("Lengths of sequence and quality values differs for %s (%i and %i)." % (head.rstrip(), len(seq), len(qual)))
                yield head, seq, qual
                line = fh.readline()
            elif line[0] == ">":
                head = line[1:]
                parts = []
                line = fh.readline()
                while line and not line.startswith(">"):
                    parts.append(line.strip())
                    line = fh.readline()
                yield head, "".join(parts), None
            else:
                raise ValueError(f"unexpected line in FASTA/FASTQ input: {line[:40]!r}")
    finally:
        if close:
            fh.close()


class _Seq(str):
    """str subclass so `str(rec.seq)` and `len(rec.seq)` behave as with Bio.Seq."""


class SeqRecord:
    __slots__ = ("id", "name", "description", "seq", "qual")

    def __init__(self, id, description, seq, qual):
        self.id = id
        self.name = id
        self.description = description
        self.seq = _Seq(seq)
        self.qual = qual

    def format_fastq(self) -> str:
        # Bio.SeqIO.QualityIO.FastqPhredWriter.write_record
        ident, desc = self.id, self.description
        if desc and desc.split(None, 1)[0] == ident:
            title = desc
        elif desc:
            title = f"{ident} {desc}"
        else:
            title = ident
        return f"@{title}\n{self.seq}\n+\n{self.qual}\n"


def parse(handle, fmt="fastq"):
    """Bio.SeqIO.parse(handle, 'fastq') look-alike."""
    if fmt not in ("fastq", "fasta"):
        raise ValueError("only fastq/fasta are supported")
    for head, seq, qual in _iter_records(handle):
        # Bio.SeqIO.QualityIO: description = the title line right-stripped (inner tabs kept), id = its first word
        desc = head.rstrip()
        words = desc.split(None, 1)
        yield SeqRecord(words[0] if words else "", desc, seq, qual if qual is not None else "")


def write(records, handle, fmt="fastq"):
    """Bio.SeqIO.write(record_or_records, handle, 'fastq') look-alike; returns the number written."""
    if isinstance(records, SeqRecord):
        records = [records]
    n = 0
    for r in records:
        handle.write(r.format_fastq())
        n += 1
    return n
