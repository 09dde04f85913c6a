"""monica's aligner / quantification helper with the mapping done on B200s.

Drop-in for the public functions of /root/reference/monica/genomes/aligner.py -- same names, argument names and
defaults, return shapes, on-disk side effects and error behaviour -- built around one device pipeline per FASTQ
file instead of one Python->C call per read (reference :193,215).

  function                 reference lines   behaviour kept
  indexer                  :31-53            deletes old *.mmi, marker files, one index<N>.mmi per database<N>.fna.gz
  index_loader             :56-62            returns None for non-.mmi paths, raises 'Damaged or empty index'
  multi_threaded_aligner   :65-111           chdir, *fastq discovery, folder creation rule, thread pool over files,
                                             indexes processed sequentially, returns 0 when there is nothing to map
  aligner                  :179-279          MAPQ/primary filter, cross-index hit carry-over via hits/<sample>_hits.pkl,
                                             best_hit, mapped/unmapped/ambiguous/focus routing with the id rewrite,
                                             three counting modes, consumes the input FASTQ
  alignment_update         :282-302          additive merge into alignment.pkl
  normalizer               :305-319          BPB then BPM, in place
  alignment_to_data_frame  :322-325          MultiIndex (tax_unit, accession) x sample CSV
  best_hit                 :328-339          argmin NM/mlen; 0 when the minimum is tied
  any_result               :342-350
"""
from __future__ import annotations

import os
import pickle
from collections import Counter
from multiprocessing.dummy import Pool as ThreadPool

from . import fastx
from . import mappy_shim as mappy

BEST_N = 15
INDEX_NAME = ['index', '.mmi']
ALIGNMENT_PICKLE_FILENAME = 'alignment.pkl'
MAPPED_FILES_FOLDER = 'mapped'
UNMAPPED_FILES_FOLDER = 'unmapped'
AMBIGUOUS_FILES_FOLDER = 'ambiguous'
HITS_FILES_FOLDER = 'hits'
FOCUS_FILES_FOLDER = 'focus'
_MODES = ('basic', 'query_length', 'matching')


def _read_root():
    marker = os.path.join(os.path.expanduser('~'), '.monica', '.root')
    if os.path.isfile(marker):
        with open(marker) as fh:
            return fh.readline()
    return os.path.join(os.path.expanduser('~'), '.monica')


MONICA_ROOT = _read_root()
INDEXES_PATH = os.path.join(MONICA_ROOT, 'indexes')
GENOMES_PATH = os.path.join(MONICA_ROOT, 'genomes')   # fetcher.GENOMES_PATH in the reference


def _touch(path):
    open(path, 'wb').close()


def _load_pickle(path):
    with open(path, 'rb') as fh:
        return pickle.load(fh)


def _dump_pickle(obj, path):
    with open(path, 'wb') as fh:
        pickle.dump(obj, fh)


# ----------------------------------------------------------------------------------------------------------------
# indexes
# ----------------------------------------------------------------------------------------------------------------
def indexer(databases, indexes_path=INDEXES_PATH, genomes_path=None):
    markers = GENOMES_PATH if genomes_path is None else genomes_path
    if os.path.exists(indexes_path):
        for stale in (f for f in os.listdir(indexes_path) if f.endswith('.mmi')):
            os.remove(os.path.join(indexes_path, stale))
    else:
        os.makedirs(indexes_path)
    print('Started building {} index'.format(indexes_path))
    os.makedirs(markers, exist_ok=True)
    _touch(os.path.join(markers, 'entered_indexer'))
    built = []
    for chunk in os.listdir(databases):
        if not chunk.endswith('.fna.gz'):
            continue
        number = os.path.basename(chunk)[len('database'):-len('.fna.gz')]
        target = os.path.join(indexes_path, str(number).join(INDEX_NAME))
        if not mappy.Aligner(fn_idx_in=os.path.join(databases, chunk), preset='map-ont', best_n=BEST_N, fn_idx_out=target):
            raise Exception('Index building failed')
        built.append(target)
    print('Finished building {} index'.format(indexes_path))
    _touch(os.path.join(markers, 'finished_indexing'))
    return built


def index_loader(index_file):
    if not index_file.endswith('.mmi'):
        return None
    print(f'aligning on {index_file}')
    handle = mappy.Aligner(fn_idx_in=index_file)
    if not handle:
        raise Exception('Damaged or empty index')
    return handle


# ----------------------------------------------------------------------------------------------------------------
# mapping one FASTQ file
# ----------------------------------------------------------------------------------------------------------------
def _confident_hits(index, sequences, mapping_quality):
    """For each sequence, the (ctg, NM, mlen) triples of hits that are primary with mapq >= mapping_quality, in the
    order index.map would yield them."""
    per_read = [[] for _ in sequences]
    if hasattr(index, 'map_batch'):                      # one device pipeline for the whole file
        res = index.map_batch(sequences)
        contigs = index.seq_names
        primary = res.is_primary != 0
        if mapping_quality is None and primary.any():
            # the reference evaluates `hit.mapq >= None` for the first primary hit it meets (aligner.py:194,216)
            raise TypeError("'>=' not supported between instances of 'int' and 'NoneType'")
        chosen = (primary & (res.mapq >= (0 if mapping_quality is None else mapping_quality))).nonzero()[0]
        for h in chosen:
            per_read[int(res.read_idx[h])].append((contigs[int(res.rid[h])], int(res.nm[h]), int(res.mlen[h])))
    else:                                                # any other mappy-shaped object
        for slot, seq in zip(per_read, sequences):
            slot.extend((h.ctg, h.NM, h.mlen) for h in index.map(seq) if h.is_primary and h.mapq >= mapping_quality)
    return per_read


class _Tally(dict):
    """{tax_unit: Counter({accession: n})} with the reference's three counting modes; any other mode counts nothing."""

    def add(self, mode, tax_unit, accession, read_len, mlen):
        if mode not in _MODES:
            return
        amount = 1 if mode == 'basic' else read_len if mode == 'query_length' else mlen
        if tax_unit in self:
            self[tax_unit].update({accession: amount})
        else:
            self[tax_unit] = Counter({accession: amount})


LAST_BREAKDOWN = None   # seconds spent in the stages of the last whole-file call (diagnostic; bench.py reports it)


def _aligner_whole_file(sample, sample_name, index, mode, mapping_quality, overnight, focus_species,
                        mapped_folder, unmapped_folder, ambiguous_folder, focus_folder):
    """The common case of `aligner` -- one index, nothing carried over, distinct read ids -- without a Python loop over the
    reads: native FASTQ ingest (mb_fastq_load: mmap, parallel record scan, sequences gathered into page-locked memory), one
    device pipeline, hit filter + best_hit + counting on the device (mb_count), a vectorised tally (per-contig sums folded by
    name in first-occurrence order, which is the reference's dict order), native routed writers (mb_fastq_route_targets:
    writev straight from the input pages).  Returns None when the file does not qualify (duplicate ids), so the caller falls
    back to the per-record path, which reproduces the reference's dictionary semantics."""
    import ctypes as C
    import time
    import numpy as np
    from . import _lib
    global LAST_BREAKDOWN
    L = _lib.lib()
    fq = C.c_void_p()
    t0 = time.perf_counter()
    rc = L.mb_fastq_load(os.fsencode(sample), C.byref(fq))
    if rc != 0:
        message = L.mb_last_error().decode(errors='replace')
        if message.startswith('malformed FASTQ'):      # what SeqIO.parse raises at aligner.py:191,212 for such a file
            raise ValueError(message)
        raise _lib.MonicaB200Error(rc, message)
    try:
        if not L.mb_fastq_ids_unique(fq):
            return None
        t1 = time.perf_counter()
        n = int(L.mb_fastq_n(fq))
        offp = C.POINTER(C.c_int64)()
        catp = L.mb_fastq_seqs(fq, C.byref(offp))
        off = np.ctypeslib.as_array(offp, shape=(n + 1,)) if n else np.zeros(1, np.int64)
        cat = np.ctypeslib.as_array(catp, shape=(int(off[-1]),)) if n and off[-1] else np.zeros(0, np.uint8)
        hits = index.map_batch(cat=cat, off=off, cigars=False)
        if mapping_quality is None and (hits.is_primary != 0).any():
            raise TypeError("'>=' not supported between instances of 'int' and 'NoneType'")
        _, _, read_class, read_best = index.count(hits, 0 if mapping_quality is None else mapping_quality, None)
        t2 = time.perf_counter()
        contigs = index.seq_names
        dest = read_class.astype(np.int8, copy=True)      # 0 unmapped, 1 mapped, 2 ambiguous (mb_count's classes)
        mapped = np.nonzero(dest == 1)[0]
        best = read_best[mapped]
        rid = hits.rid[best].astype(np.int64) if len(mapped) else np.zeros(0, np.int64)
        # per contig: (tax unit written into the record id, tax unit before the `overnight` cut, accession)
        parts = [c.split(sep=':') for c in contigs]
        unit_full = [p[0] for p in parts]
        unit = [u.split(sep='_')[0] if overnight else u for u in unit_full]
        ids = (C.c_char_p * max(len(contigs), 1))(*[u.encode() for u in unit])
        target = np.full(max(n, 1), -1, np.int32)
        target[mapped] = rid
        focus = None
        if focus_species:
            is_focus = np.array([u in focus_species for u in unit_full], dtype=bool)
            focus = np.zeros(max(n, 1), np.uint8)
            focus[mapped] = is_focus[rid]
        tally = _Tally()
        if mode in _MODES and len(mapped):
            lens = np.diff(off)
            amount = np.ones(len(mapped), np.int64) if mode == 'basic' else lens[mapped].astype(np.int64) if mode == 'query_length' \
                else hits.mlen[best].astype(np.int64)
            total = np.zeros(len(contigs), np.int64)
            np.add.at(total, rid, amount)
            seen, first = np.unique(rid, return_index=True)
            for r in seen[np.argsort(first)]:            # first-occurrence order = the order in which the reference creates its dict keys
                if len(parts[r]) < 2:
                    raise IndexError('list index out of range')   # the reference indexes ctg.split(':')[1]
                if unit[r] in tally:
                    tally[unit[r]].update({parts[r][1]: int(total[r])})
                else:
                    tally[unit[r]] = Counter({parts[r][1]: int(total[r])})
        t3 = time.perf_counter()
        _lib.check(L.mb_fastq_route_targets(fq, dest.ctypes.data_as(C.c_void_p), target.ctypes.data_as(C.c_void_p), C.cast(ids, C.c_void_p), len(contigs),
                                            focus.ctypes.data_as(C.c_void_p) if focus is not None else None,
                                            os.fsencode(os.path.join(mapped_folder, sample)), os.fsencode(os.path.join(unmapped_folder, sample)),
                                            os.fsencode(os.path.join(ambiguous_folder, sample)),
                                            os.fsencode(os.path.join(focus_folder, sample)) if focus_species else None))
        t4 = time.perf_counter()
        LAST_BREAKDOWN = {"fastq_load": t1 - t0, "map_and_count": t2 - t1, "tally": t3 - t2, "route_write": t4 - t3}
        if os.environ.get("MB_DEBUG"):
            import sys
            print(f"[mb] aligner {sample}: load {t1 - t0:.3f} map+count {t2 - t1:.3f} tally {t3 - t2:.3f} route {t4 - t3:.3f} s "
                  f"(started {t0:.3f}, ended {t4:.3f})", file=sys.stderr)
        return dict(tally)
    finally:
        L.mb_fastq_free(fq)


def aligner(sample, sample_name, index, mode=None, hits_folder=None, mapping_quality=None, overnight=False,
            focus_species=[], mapped_folder=None, unmapped_folder=None, ambiguous_folder=None, focus_folder=None,
            last_index=False):
    print(f'{sample}, mode is {mode}\t')
    carry_name = sample_name + '_hits.pkl'
    carry_path = os.path.join(hits_folder, carry_name)
    if last_index and carry_name not in os.listdir(hits_folder) and hasattr(index, 'map_batch') and hasattr(index, 'count') \
            and os.environ.get('MONICA_B200_PER_RECORD') != '1':
        done = _aligner_whole_file(sample, sample_name, index, mode, mapping_quality, overnight, focus_species,
                                   mapped_folder, unmapped_folder, ambiguous_folder, focus_folder)
        if done is not None:
            # the reference leaves an (empty-able) carry-over pickle behind only transiently: it writes and removes it
            print(f'{sample} done')
            os.remove(sample)
            return done, sample_name
    carried = _load_pickle(carry_path) if carry_name in os.listdir(hits_folder) else dict()

    records = list(fastx.parse(sample, 'fastq'))
    fresh = _confident_hits(index, [str(r.seq) for r in records], mapping_quality)
    def absorb(rec, found):                              # merge this index's hits into the carried ones
        if found:
            if rec.id in carried:
                carried[rec.id].extend(found)
            else:
                carried[rec.id] = found

    if not last_index:
        for rec, found in zip(records, fresh):
            absorb(rec, found)
        _dump_pickle(carried, carry_path)
        return None

    tally = _Tally()
    sinks = {'mapped': open(os.path.join(mapped_folder, sample), 'a'),
             'unmapped': open(os.path.join(unmapped_folder, sample), 'a'),
             'ambiguous': open(os.path.join(ambiguous_folder, sample), 'a')}
    if focus_species:
        sinks['focus'] = open(os.path.join(focus_folder, sample), 'a')
    try:
        for rec, found in zip(records, fresh):
            absorb(rec, found)                           # per record, as the reference does (matters for duplicate ids)
            candidates = carried.get(rec.id)
            if candidates is None:
                sinks['unmapped'].write(rec.format_fastq())
                continue
            winner = candidates[0] if len(candidates) == 1 else best_hit(candidates)
            if not winner:
                sinks['ambiguous'].write(rec.format_fastq())
                continue
            tax_unit, accession = winner[0].split(sep=':')[0], winner[0].split(sep=':')[1]
            if tax_unit in focus_species:
                sinks['focus'].write(rec.format_fastq())
            if overnight:
                tax_unit = tax_unit.split(sep='_')[0]    # genus
            rec.id = tax_unit
            sinks['mapped'].write(rec.format_fastq())
            tally.add(mode, tax_unit, accession, len(rec.seq), winner[2])
    finally:
        for fh in sinks.values():
            fh.close()

    # the reference rewrites the carry-over pickle one last time and then removes it
    if carry_name in os.listdir(hits_folder):
        previous = _load_pickle(carry_path)
        previous.update(carried)
        _dump_pickle(previous, carry_path)
    else:
        _dump_pickle(carried, carry_path)
    os.remove(carry_path)

    print(f'{sample} done')
    os.remove(sample)
    return dict(tally), sample_name


def multi_threaded_aligner(query_folder, indexes_paths, mode=None, mapping_quality=60, overnight=False, n_threads=None,
                           focus_species=[], output_folder=None, mapped_files_folder=MAPPED_FILES_FOLDER,
                           unmapped_files_folder=UNMAPPED_FILES_FOLDER, ambiguous_files_folder=AMBIGUOUS_FILES_FOLDER,
                           hits_files_folder=HITS_FILES_FOLDER, focus_file_folder=FOCUS_FILES_FOLDER, index_loader_fn=None):
    """`index_loader_fn` is the one addition to the reference signature: tests inject another mappy-shaped loader."""
    open_index = index_loader_fn or index_loader
    os.chdir(query_folder)
    samples = [f for f in os.listdir('.') if f.endswith('fastq') and os.stat(f).st_size]
    if not samples:
        print('No query files were provided')
        return 0
    names = [s.split('.')[0] for s in samples]
    folder = {k: os.path.join(query_folder, v) for k, v in (('mapped', mapped_files_folder), ('unmapped', unmapped_files_folder),
                                                            ('ambiguous', ambiguous_files_folder), ('hits', hits_files_folder),
                                                            ('focus', focus_file_folder))}
    if not os.path.exists(folder['mapped']):
        for k in ('mapped', 'unmapped', 'ambiguous', 'hits'):
            os.mkdir(folder[k])
        if focus_species:
            os.mkdir(folder['focus'])

    pool = ThreadPool(n_threads)
    try:
        for path in indexes_paths[:-1]:                  # all but the last index only accumulate confident hits
            idx = open_index(path)
            pool.starmap(aligner, [(s, n, idx, mode, folder['hits'], mapping_quality) for s, n in zip(samples, names)])
        idx = open_index(indexes_paths[-1])
        results = pool.starmap(aligner, [(s, n, idx, mode, folder['hits'], mapping_quality, overnight, focus_species,
                                          folder['mapped'], folder['unmapped'], folder['ambiguous'], folder['focus'], True)
                                         for s, n in zip(samples, names)])
    finally:
        pool.close()
    return alignment_update(results, output_folder)


# ----------------------------------------------------------------------------------------------------------------
# quantification tail
# ----------------------------------------------------------------------------------------------------------------
def alignment_update(results, output_folder):
    store = os.path.join(output_folder, ALIGNMENT_PICKLE_FILENAME)
    if ALIGNMENT_PICKLE_FILENAME not in os.listdir(output_folder):
        # first call of a run: plain assignment, so two files sharing a sample name overwrite each other (as upstream)
        alignment = {name: per_sample for per_sample, name in results}
    else:
        alignment = _load_pickle(store)
        for per_sample, name in results:
            known = alignment.get(name)
            if known is None:
                alignment[name] = per_sample
                continue
            for tax_unit, counter in per_sample.items():
                if tax_unit in known:
                    known[tax_unit].update(counter)
                else:
                    known[tax_unit] = counter
    _dump_pickle(alignment, store)
    return alignment


def normalizer(alignment, genomes_length=None):
    if not genomes_length:
        genomes_length = _load_pickle(os.path.join(GENOMES_PATH, 'current_genomes_length.pkl'))
    for per_sample in alignment.values():
        total = 0
        for counter in per_sample.values():
            for accession in list(counter):
                counter[accession] = counter[accession] / genomes_length[accession]      # bases per base
                total += counter[accession]
        for counter in per_sample.values():
            for accession in list(counter):
                counter[accession] = counter[accession] / total                          # fraction of the sample
    return alignment


def alignment_to_data_frame(alignment, output_folder=None, filename='monica.dataframe'):
    import pandas as pd
    columns = {sample: pd.DataFrame(per_sample).unstack() for sample, per_sample in alignment.items() if per_sample}
    frame = pd.concat(columns, axis=1).dropna(how='all')
    frame.to_csv(os.path.join(output_folder, filename))
    return frame


def best_hit(hits):
    """The hit with the smallest NM/mlen; 0 when that minimum is shared (the scan uses `<=`, so what matters is whether
    the LAST time the leader changed it was by a tie) or when `hits` is empty."""
    leader, leader_ratio, margin = None, float('inf'), 0
    for hit in hits:
        ratio = float(hit[1]) / hit[2]
        if ratio <= leader_ratio:
            margin = leader_ratio - ratio
            leader, leader_ratio = hit, ratio
    return leader if margin else 0


def any_result(alignment):
    return 1 if any(bool(v) for v in alignment.values()) else 0
