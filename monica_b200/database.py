"""monica's database builder over the native FASTA re-header writer.

Drop-in for /root/reference/monica/genomes/database.py -- same names, argument names and defaults, return shapes and
on-disk side effects:

  function                 reference lines   behaviour kept
  multi_threaded_builder   :16-49            empties `databases_path` of *.fna.gz, one thread per chunk, merges the
                                             per-chunk genome lengths into current_genomes_length.pkl, deletes the
                                             downloaded genomes unless keep_genomes, `database_created` marker, returns
                                             (databases_path, current_genomes_length)
  builder                  :52-67            database<N>.fna.gz = every record of every genome of the chunk re-headed
                                             "<tax_unit>:<accession>" (the contig names the aligner splits on ':',
                                             aligner.py:234); returns {accession: genome_length}
  _genomes_splitter        :70-94            greedy chunks by compressed size, INCLUDING its quirk: the genome that does
                                             not fit the running chunk closes that chunk and is itself dropped (it is in
                                             neither chunk), and an oversize genome is yielded alone

The wire format matters to the hot path: it decides contig naming (duplicate names per genome) and which genomes exist in
an index.  The quirk is preserved on purpose -- a drop-in must index the same genomes as the reference; `strict=True` (an
extension, off by default) keeps the overflowing genome as the first member of the next chunk instead.

A genome is `(path_to_fna_gz, (tax_unit, accession))`.  The record round trip the reference does through Biopython
(`SeqIO.parse` -> `SeqIO.write`, database.py:60-64) is `mb_db_build` in the C-ABI library: decompressed bytes identical.
"""
from __future__ import annotations

import ctypes as C
import os
import pickle
from itertools import count, repeat
from multiprocessing.dummy import Pool as ThreadPool

import numpy as np

from . import _lib
from .aligner import GENOMES_PATH

DATABASES_PATH = os.path.join(GENOMES_PATH, 'databases')
DATABASE_NAME = ['database', '.fna.gz']


def multi_threaded_builder(genomes=None, max_chunk_size=None, databases_path=DATABASES_PATH,
                           database_name=DATABASE_NAME, keep_genomes=None, n_threads=None, genomes_path=None):
    """database.py:16-49.  `genomes_path` (extension) overrides the module-level GENOMES_PATH the reference hard-wires."""
    genomes_path = GENOMES_PATH if genomes_path is None else genomes_path
    if not os.path.exists(databases_path):
        os.makedirs(databases_path)
    else:
        for database in os.listdir(databases_path):
            if database.endswith('.fna.gz'):
                os.remove(os.path.join(databases_path, database))

    lengths_pickle = os.path.join(genomes_path, 'current_genomes_length.pkl')
    if 'current_genomes_length.pkl' in os.listdir(genomes_path):
        with open(lengths_pickle, 'rb') as fh:
            current_genomes_length = pickle.load(fh)
    else:
        current_genomes_length = dict()

    pool = ThreadPool(n_threads)
    try:
        lengths = pool.starmap(builder, zip(_genomes_splitter(genomes, max_chunk_size=max_chunk_size),
                                            repeat(databases_path), repeat(database_name), count()))
    finally:
        pool.close()

    for length in lengths:
        current_genomes_length.update(length)

    if not keep_genomes:
        for genome in os.listdir(genomes_path):
            if genome.endswith('.fna.gz'):
                os.remove(os.path.join(genomes_path, genome))

    with open(lengths_pickle, 'wb') as fh:
        pickle.dump(current_genomes_length, fh)

    with open(os.path.join(genomes_path, 'database_created'), 'wb'):
        pass
    return databases_path, current_genomes_length


def builder(genomes_chunk, databases_path, database_name, database_number):
    """database.py:52-67: one chunk -> one gzip FASTA, one native call (ctypes releases the GIL, so the reference's thread
    pool over chunks runs the chunks in parallel for real)."""
    database_file = os.path.join(databases_path, str(database_number).join(database_name))
    print('Working on {}'.format(str(database_number).join(database_name)))
    genomes_chunk = list(genomes_chunk)
    n = len(genomes_chunk)
    paths = (C.c_char_p * max(n, 1))(*[os.fsencode(g[0]) for g in genomes_chunk])
    heads = (C.c_char_p * max(n, 1))(*[':'.join(g[1]).encode() for g in genomes_chunk])
    glen = np.zeros(max(n, 1), dtype=np.int64)
    _lib.check(_lib.lib().mb_db_build(os.fsencode(database_file), n, paths, heads, _lib._ptr(glen)))
    this_database_genomes_length = dict()
    for i, genome in enumerate(genomes_chunk):
        this_database_genomes_length[genome[1][1]] = int(glen[i])
    print('Finished building {}'.format(str(database_number).join(database_name)))
    return this_database_genomes_length


def _genomes_splitter(genomes, max_chunk_size=None, strict=False):
    """database.py:70-94, quirk included (see module docstring)."""
    chunk = []
    chunk_size = 0
    for genome in genomes:
        size = os.path.getsize(genome[0])
        if size > max_chunk_size:
            print('Genome {}, ({}) alone expected to generate an index '
                  'exceeding the maximum memory deriving from settings of {} bytes'
                  .format(genome[0], genome[1][0], (size - max_chunk_size)*16))
            yield [genome]
        else:
            if chunk_size + size <= max_chunk_size:
                chunk.append(genome)
                chunk_size += size
            else:
                yield chunk
                chunk = [genome] if strict else []
                chunk_size = size if strict else 0
    if chunk:
        yield chunk
