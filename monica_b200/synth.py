"""Seeded synthetic inputs of the shapes BASELINE.json names (NCBI fetches are unavailable offline).

Genomes are uniform-random ACGT; a fraction are mutated "strain" copies of earlier genomes so that
repeat filtering (mid_occ > 2), MAPQ < 60 and monica's best_hit ties are exercised.
Contig names follow the database builder's wire format ``Species:accession``
(/root/reference/monica/genomes/database.py:59-64), which the aligner parses at
/root/reference/monica/genomes/aligner.py:234,240.

Reads are ONT-like: log-normal lengths, substitution / insertion / deletion errors at a stated mix,
random strand.
"""
from __future__ import annotations

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b


def revcomp(seq: np.ndarray) -> np.ndarray:
    return _COMP[seq[::-1]]


def random_genome(rng: np.random.Generator, length: int) -> np.ndarray:
    return _ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]


def mutate(rng: np.random.Generator, seq: np.ndarray, sub: float, ins: float, dele: float) -> np.ndarray:
    """Apply i.i.d. substitutions / insertions / deletions (rates per base)."""
    n = len(seq)
    r = rng.random(n)
    out = seq.copy()
    is_sub = r < sub
    # substitution: shift to a different base
    if is_sub.any():
        codes = np.searchsorted(_ACGT, out[is_sub])
        out[is_sub] = _ACGT[(codes + rng.integers(1, 4, size=int(is_sub.sum()))) % 4]
    is_del = (r >= sub) & (r < sub + dele)
    is_ins = (r >= sub + dele) & (r < sub + dele + ins)
    keep = ~is_del
    if not is_ins.any():
        return out[keep]
    # insertion of one random base after the position
    reps = keep.astype(np.int64) + is_ins.astype(np.int64)
    idx = np.repeat(np.arange(n), reps)
    res = out[idx]
    # positions that are the inserted copy: second occurrence for kept+ins, only occurrence for del+ins
    first = np.ones(len(idx), dtype=bool)
    first[1:] = idx[1:] != idx[:-1]
    inserted = (~first) | (first & is_del[idx] & is_ins[idx])
    res[inserted] = _ACGT[rng.integers(0, 4, size=int(inserted.sum()))]
    return res


def make_genomes(seed: int, n_genomes: int, genome_len: int, strain_frac: float = 0.2,
                 strain_div: tuple[float, float] = (0.01, 0.05), contigs_per_genome: int = 1):
    """Return (names, seqs) where names are 'Species_i:ACC_i' and seqs are uint8 ASCII arrays.

    Every contig of one genome carries the same name, as monica's database builder does.
    """
    rng = np.random.default_rng(seed)
    names, seqs, base = [], [], []
    n_strain = int(round(n_genomes * strain_frac)) if n_genomes > 1 else 0
    for g in range(n_genomes):
        if g >= n_genomes - n_strain and base:
            src = base[int(rng.integers(0, len(base)))]
            d = float(rng.uniform(*strain_div))
            s = mutate(rng, src, d * 0.8, d * 0.1, d * 0.1)
        else:
            s = random_genome(rng, genome_len)
            base.append(s)
        name = f"Species_{g}:ACC{g:05d}.1"
        if contigs_per_genome <= 1:
            names.append(name)
            seqs.append(s)
        else:
            cuts = np.linspace(0, len(s), contigs_per_genome + 1).astype(np.int64)
            for c in range(contigs_per_genome):
                names.append(name)
                seqs.append(s[cuts[c]:cuts[c + 1]])
    return names, seqs


def lognormal_lengths(rng: np.random.Generator, n: int, mean: float, sigma: float = 0.6,
                      min_len: int = 200, max_len: int | None = None) -> np.ndarray:
    mu = np.log(mean) - 0.5 * sigma * sigma
    L = rng.lognormal(mu, sigma, size=n).astype(np.int64)
    L = np.maximum(L, min_len)
    if max_len is not None:
        L = np.minimum(L, max_len)
    return L


def simulate_reads(seed: int, seqs: list[np.ndarray], n_reads: int, mean_len: float, error: float = 0.10,
                   mix: tuple[float, float, float] = (0.4, 0.3, 0.3), sigma: float = 0.6,
                   junk_frac: float = 0.0, fixed_len: int | None = None):
    """Simulate ONT-like reads.

    error is the total per-base error rate, split into (substitution, insertion, deletion) by ``mix``
    (default 4% / 3% / 3% at error = 10%).  Returns (reads, truth) with reads a list of uint8 ASCII
    arrays and truth a list of (contig_idx, start, end, strand) (contig_idx = -1 for junk reads).
    """
    rng = np.random.default_rng(seed)
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    prob = lens / lens.sum()
    if fixed_len is not None:
        L = np.full(n_reads, fixed_len, dtype=np.int64)
    else:
        L = lognormal_lengths(rng, n_reads, mean_len, sigma)
    reads, truth = [], []
    sub, ins, dele = (error * m for m in mix)
    for i in range(n_reads):
        if junk_frac > 0 and rng.random() < junk_frac:
            reads.append(random_genome(rng, int(L[i])))
            truth.append((-1, 0, 0, 0))
            continue
        c = int(rng.choice(len(seqs), p=prob))
        ln = int(min(L[i], lens[c]))
        st = int(rng.integers(0, lens[c] - ln + 1))
        frag = seqs[c][st:st + ln]
        strand = int(rng.integers(0, 2))
        if strand:
            frag = revcomp(frag)
        reads.append(mutate(rng, frag, sub, ins, dele))
        truth.append((c, st, st + ln, strand))
    return reads, truth


def concat_reads(reads: list[np.ndarray]):
    """Concatenate reads into one uint8 buffer + int64 offsets[n+1] (the C-ABI's batch layout)."""
    off = np.zeros(len(reads) + 1, dtype=np.int64)
    if reads:
        off[1:] = np.cumsum([len(r) for r in reads])
        cat = np.concatenate(reads).astype(np.uint8, copy=False)
    else:
        cat = np.zeros(0, dtype=np.uint8)
    return np.ascontiguousarray(cat), off


def write_fastq(path: str, reads: list[np.ndarray], prefix: str = "read"):
    with open(path, "wb") as fh:
        for i, r in enumerate(reads):
            fh.write(b"@" + f"{prefix}{i}".encode() + b"\n")
            fh.write(r.tobytes() + b"\n+\n")
            fh.write(b"I" * len(r) + b"\n")


def write_fasta_gz(path: str, names: list[str], seqs: list[np.ndarray]):
    import gzip
    with gzip.open(path, "wb", compresslevel=1) as fh:
        for n, s in zip(names, seqs):
            fh.write(b">" + n.encode() + b"\n")
            b = s.tobytes()
            for i in range(0, len(b), 80):
                fh.write(b[i:i + 80] + b"\n")


def edge_reads(seed: int, seqs: list[np.ndarray]):
    """Hand-shaped reads that reach the rarely-taken paths: empty / shorter than k / N-containing reads, chimeras
    (two loci in one read), a long junk insertion (band-limited DP, Z-drop split), a long deletion, a read overhanging a
    contig end, an exact (error-free) read and a reverse-complement one, and a tandem-repeat read (anchor ties)."""
    rng = np.random.default_rng(seed)
    g = seqs[0]
    G = len(g)
    out = []
    out.append(np.zeros(0, dtype=np.uint8))                                   # empty
    out.append(g[100:110].copy())                                             # shorter than k
    out.append(g[200:240].copy())                                             # a few minimizers only
    r = mutate(rng, g[1000:4000], 0.03, 0.02, 0.02); r[500:520] = ord("N"); r[1500] = ord("n"); out.append(r)  # Ns
    a = mutate(rng, g[5000:8000], 0.04, 0.03, 0.03)
    b = mutate(rng, revcomp(g[min(G - 3500, 20000):min(G - 500, 23000)]), 0.04, 0.03, 0.03)
    out.append(np.concatenate([a, b]))                                        # chimera, different strands
    out.append(np.concatenate([mutate(rng, g[9000:11500], 0.04, 0.03, 0.03), random_genome(rng, 2000),
                               mutate(rng, g[11500:14000], 0.04, 0.03, 0.03)]))  # 2 kb junk insertion
    out.append(np.concatenate([mutate(rng, g[15000:17000], 0.04, 0.03, 0.03),
                               mutate(rng, g[18200:20500], 0.04, 0.03, 0.03)]))  # 1.2 kb deletion
    out.append(np.concatenate([random_genome(rng, 900), mutate(rng, g[0:2500], 0.04, 0.03, 0.03)]))  # overhang at contig start
    out.append(np.concatenate([mutate(rng, g[G - 2500:G], 0.04, 0.03, 0.03), random_genome(rng, 1200)]))  # overhang at end
    out.append(g[30000:33000].copy())                                         # exact
    out.append(revcomp(g[33000:36000]))                                       # exact, reverse strand
    unit = random_genome(rng, 37)
    out.append(np.concatenate([g[40000:41000], np.tile(unit, 40), g[41000:42000]]))  # tandem repeat inside the read
    out.append(np.tile(g[42000:42600], 4))                                    # the read repeats a reference segment 4x
    out.append(random_genome(rng, 5000))                                      # unmappable
    blk = mutate(rng, g[46000:52000], 0.03, 0.02, 0.02); blk[2500:3400] = random_genome(rng, 900)
    out.append(blk)                                                           # divergent block: Z-drop test -> 2nd pass -> split
    inv = g[52000:58000].copy(); inv[2000:3200] = revcomp(inv[2000:3200])
    out.append(mutate(rng, inv, 0.03, 0.02, 0.02))                            # internal inversion: zdrop_code 2, split_inv
    blk2 = mutate(rng, g[2000:9000], 0.05, 0.04, 0.04); blk2[1500:2100] = random_genome(rng, 600); blk2[4000:4700] = random_genome(rng, 700)
    out.append(blk2)                                                          # two divergent blocks: repeated splitting
    lower = np.char.lower(g[44000:46000].tobytes().decode()).encode() if False else bytes(g[44000:46000]).lower()
    out.append(np.frombuffer(lower, dtype=np.uint8).copy())                   # lower-case bases
    return out


def simulate_reads_bulk(seed: int, seqs: list[np.ndarray], n_reads: int, n50: float, error: float = 0.10,
                        mix: tuple[float, float, float] = (0.4, 0.3, 0.3), sigma: float = 0.6, block: int = 2000,
                        min_len: int = 500, max_len: int | None = None):
    """Vectorised simulator for bench-sized workloads: returns (cat uint8[total], off int64[n+1]).

    Read lengths are log-normal with length-weighted median (N50) = n50: mu = ln(n50) - sigma^2.  Errors are i.i.d.
    substitution / insertion / deletion at `error` x mix (default 4% / 3% / 3% at 10%), strands are random."""
    rng = np.random.default_rng(seed)
    glens = np.array([len(s) for s in seqs], dtype=np.int64)
    gcat = np.concatenate(seqs)
    goff = np.zeros(len(seqs) + 1, dtype=np.int64)
    goff[1:] = np.cumsum(glens)
    prob = glens / glens.sum()
    mu = np.log(n50) - sigma * sigma
    sub, ins, dele = (error * m for m in mix)
    chunks, lens_out = [], []
    done = 0
    while done < n_reads:
        nb = min(block, n_reads - done)
        L = np.maximum(rng.lognormal(mu, sigma, size=nb).astype(np.int64), min_len)
        if max_len is not None:
            L = np.minimum(L, max_len)
        c = rng.choice(len(seqs), size=nb, p=prob)
        L = np.minimum(L, glens[c])
        st = (rng.random(nb) * (glens[c] - L + 1)).astype(np.int64)
        strand = rng.integers(0, 2, size=nb).astype(bool)
        # gather all fragments of the block with one index array; reverse strands read backwards and complemented
        roff = np.zeros(nb + 1, dtype=np.int64)
        roff[1:] = np.cumsum(L)
        within = np.arange(roff[-1], dtype=np.int64) - np.repeat(roff[:-1], L)
        rid = np.repeat(np.arange(nb), L)
        pos = np.where(strand[rid], (st + L - 1)[rid] - within, st[rid] + within) + goff[c][rid]
        frag = gcat[pos]
        rs = strand[rid]
        frag[rs] = _COMP[frag[rs]]
        # errors over the whole block at once
        r = rng.random(len(frag), dtype=np.float32)
        is_sub = r < sub
        if is_sub.any():
            codes = np.searchsorted(_ACGT, frag[is_sub])
            frag[is_sub] = _ACGT[(codes + rng.integers(1, 4, size=int(is_sub.sum()))) % 4]
        is_del = (r >= sub) & (r < sub + dele)
        is_ins = (r >= sub + dele) & (r < sub + dele + ins)
        reps = (~is_del).astype(np.int8) + is_ins.astype(np.int8)
        idx = np.repeat(np.arange(len(frag), dtype=np.int64), reps)
        out = frag[idx]
        dup = np.zeros(len(idx), dtype=bool)
        dup[1:] = idx[1:] == idx[:-1]
        out[dup] = _ACGT[rng.integers(0, 4, size=int(dup.sum()))]
        newlen = np.bincount(rid, weights=reps, minlength=nb).astype(np.int64)
        chunks.append(out)
        lens_out.append(newlen)
        done += nb
    lens_all = np.concatenate(lens_out)
    off = np.zeros(n_reads + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens_all)
    return np.ascontiguousarray(np.concatenate(chunks)), off
