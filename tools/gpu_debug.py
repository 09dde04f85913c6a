"""Stage-by-stage GPU-vs-oracle comparison that keeps going after a mismatch (development aid; run under gpurun).

Usage: python tests/gpu_debug.py [n_reads] [mean_len] [error]
"""
from __future__ import annotations

import ctypes as C
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from monica_b200 import _lib, synth  # noqa: E402
from monica_b200.mappy_shim import Aligner  # noqa: E402
from oracle import oracle as O  # noqa: E402

CMP_FIELDS = ["rid", "rev", "qs", "qe", "rs", "re", "mapq", "mlen", "blen", "nm", "dp_max", "dp_max2", "score", "score0", "cnt",
              "subsc", "n_sub", "id", "parent", "is_primary", "sam_pri", "n_cigar"]


def make_case(seed=1, n_genomes=3, glen=int(os.environ.get('GLEN', 60000)), n_reads=60, mean_len=3000, error=0.10, strain_frac=0.34, junk=0.05):
    names, seqs = synth.make_genomes(seed, n_genomes, glen, strain_frac=strain_frac)
    reads, truth = synth.simulate_reads(seed + 1, seqs, n_reads, mean_len, error, junk_frac=junk)
    edge = synth.edge_reads(seed + 2, seqs)
    reads = edge + reads
    truth = [(-2, 0, 0, 0)] * len(edge) + truth
    return names, seqs, reads, truth


def stage_sketch(reads):
    cat, off = synth.concat_reads(reads)
    cap = len(cat) + 64
    out = np.zeros((cap, 2), dtype=np.uint64)
    ooff = np.zeros(len(reads) + 1, dtype=np.int64)
    _lib.check(_lib.lib().mb_sketch(0, _lib._ptr(cat), _lib._ptr(off), len(reads), 10, 15, _lib._ptr(out), cap, _lib._ptr(ooff)))
    bad = 0
    for i, r in enumerate(reads):
        want = O.sketch(r, 10, 15, 0)
        got = out[ooff[i]:ooff[i + 1]].copy()
        got[:, 1] &= np.uint64(0xffffffff)
        if want.shape != got.shape or not np.array_equal(want, got):
            bad += 1
            if bad <= 3:
                print(f"  sketch mismatch read {i} len {len(r)}: want {want.shape} got {got.shape}")
                n = min(len(want), len(got))
                d = np.nonzero((want[:n] != got[:n]).any(axis=1))[0]
                if len(d):
                    print("   first diff at", d[0], want[d[0]], got[d[0]])
    print(f"[sketch] reads={len(reads)} minimizers={ooff[-1]} mismatching reads={bad}")
    return bad == 0


def stage_seed(al, oidx, reads):
    cat, off = synth.concat_reads(reads)
    cap = int(len(cat)) * 4 + 1024
    out = np.zeros((cap, 2), dtype=np.uint64)
    ooff = np.zeros(len(reads) + 1, dtype=np.int64)
    rep = np.zeros(len(reads), dtype=np.int32)
    _lib.check(_lib.lib().mb_seed(al.handle(), C.byref(al.opt), _lib._ptr(cat), _lib._ptr(off), len(reads), _lib._ptr(out), cap,
                                  _lib._ptr(ooff), _lib._ptr(rep)))
    bad = 0
    traces = []
    for i, r in enumerate(reads):
        hits, stats, tr = oidx.map(r, trace=True)
        traces.append((hits, stats, tr))
        want = tr["anchors"]
        got = out[ooff[i]:ooff[i + 1]]
        ok = want.shape == got.shape and np.array_equal(want, got) and stats["rep_len"] == rep[i]
        if not ok:
            bad += 1
            if bad <= 3:
                print(f"  seed mismatch read {i}: want {want.shape} got {got.shape} rep {stats['rep_len']} vs {rep[i]}")
    print(f"[seed] anchors={ooff[-1]} mismatching reads={bad}")
    return bad == 0, traces


def stage_chain(al, traces):
    anchors = [t[2]["anchors"] for t in traces]
    off = np.zeros(len(anchors) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(a) for a in anchors])
    n_a = int(off[-1])
    cat = np.concatenate(anchors) if n_a else np.zeros((0, 2), np.uint64)
    cat = np.ascontiguousarray(cat, dtype=np.uint64)
    f = np.zeros(n_a + 1, np.int32); p = np.zeros(n_a + 1, np.int32); v = np.zeros(n_a + 1, np.int32)
    ch = np.zeros((n_a + 1, 2), np.uint64); choff = np.zeros(len(anchors) + 1, np.int64)
    u = np.zeros(n_a + 1, np.uint64); uoff = np.zeros(len(anchors) + 1, np.int64)
    _lib.check(_lib.lib().mb_chain(0, C.byref(al.opt), _lib._ptr(cat), _lib._ptr(off), len(anchors), _lib._ptr(f), _lib._ptr(p), _lib._ptr(v),
                                   _lib._ptr(ch), _lib._ptr(choff), _lib._ptr(u), _lib._ptr(uoff)))
    bad_dp = bad_bt = 0
    for i, t in enumerate(traces):
        tr = t[2]
        s, e = off[i], off[i + 1]
        if e > s:
            okdp = np.array_equal(tr["f"], f[s:e]) and np.array_equal(tr["p"], p[s:e]) and np.array_equal(tr["v"], v[s:e])
            if not okdp:
                bad_dp += 1
                if bad_dp <= 3:
                    d = np.nonzero((tr["f"] != f[s:e]) | (tr["p"] != p[s:e]) | (tr["v"] != v[s:e]))[0]
                    j = d[0]
                    print(f"  chain dp mismatch read {i} n={e - s} first at {j}: f {tr['f'][j]} vs {f[s + j]} p {tr['p'][j]} vs {p[s + j]} v {tr['v'][j]} vs {v[s + j]}")
        okbt = np.array_equal(tr["u"], u[uoff[i]:uoff[i + 1]]) and np.array_equal(tr["chained"], ch[choff[i]:choff[i + 1]])
        if not okbt:
            bad_bt += 1
            if bad_bt <= 3:
                print(f"  chain backtrack mismatch read {i}: u {tr['u']} vs {u[uoff[i]:uoff[i + 1]]} chained {tr['chained'].shape} vs {choff[i + 1] - choff[i]}")
    print(f"[chain] anchors={n_a} dp-mismatching reads={bad_dp} backtrack-mismatching reads={bad_bt}")
    return bad_dp == 0 and bad_bt == 0


def stage_dp(al, traces, max_tasks=4000):
    recs = []
    for t in traces:
        recs.extend(t[2]["dp"])
    recs = recs[:max_tasks]
    if not recs:
        print("[dp] no tasks")
        return True
    n = len(recs)
    tasks = (_lib.DpTask * n)()
    pool = []
    po = 0
    co = 0
    for i, r in enumerate(recs):
        tk = tasks[i]
        tk.qlen, tk.tlen, tk.w, tk.zdrop, tk.end_bonus, tk.flag = r["qlen"], r["tlen"], r["w"], r["zdrop"], r["end_bonus"], r["flag"]
        tk.q_off = po; pool.append(r["q"]); po += r["qlen"]
        tk.t_off = po; pool.append(r["t"]); po += r["tlen"]
        tk.cigar_off = co; co += r["qlen"] + r["tlen"] + 1
    pool = np.ascontiguousarray(np.concatenate(pool), dtype=np.uint8)
    cig = np.zeros(co + 1, dtype=np.uint32)
    t0 = time.time()
    _lib.check(_lib.lib().mb_dp_batch(0, C.byref(al.opt), tasks, n, _lib._ptr(pool), len(pool), _lib._ptr(cig), len(cig)))
    dt = time.time() - t0
    bad = 0
    for i, r in enumerate(recs):
        tk = tasks[i]
        got = dict(score=tk.score, max=tk.max, max_q=tk.max_q, max_t=tk.max_t, mqe=tk.mqe, mqe_t=tk.mqe_t, zdropped=tk.zdropped,
                   reach_end=tk.reach_end, n_cigar=tk.n_cigar)
        keys = ["zdropped", "reach_end", "n_cigar", "score"]
        if not (r["flag"] & 0x08):
            keys += ["max", "max_q", "max_t", "mqe", "mqe_t"]
        ok = all(got[k] == r[k] for k in keys) and np.array_equal(cig[tk.cigar_off:tk.cigar_off + tk.n_cigar], r["cigar"])
        if not ok:
            bad += 1
            if bad <= 5:
                print(f"  dp mismatch task {i} qlen={r['qlen']} tlen={r['tlen']} w={r['w']} flag={r['flag']:#x} zdrop={r['zdrop']} eb={r['end_bonus']}")
                print("    want", {k: r[k] for k in got}, "\n    got ", got)
                wc, gc = r["cigar"], cig[tk.cigar_off:tk.cigar_off + tk.n_cigar]
                m = min(len(wc), len(gc))
                dd = np.nonzero(wc[:m] != gc[:m])[0]
                print("    cigar first diff", (dd[0], wc[dd[0]], gc[dd[0]]) if len(dd) else None, len(wc), len(gc))
    print(f"[dp] tasks={n} mismatching={bad} wall={dt:.3f}s")
    return bad == 0


def stage_full(al, oidx, reads, traces=None):
    t0 = time.time()
    hits = al.map_batch(reads)
    dt = time.time() - t0
    per = hits.per_read()
    bad = 0
    n_hits = 0
    for i, r in enumerate(reads):
        want = traces[i][0] if traces else oidx.map(r)[0]
        n_hits += len(want)
        got_idx = per[i]
        ok = len(want) == len(got_idx)
        if ok:
            for w, gi in zip(want, got_idx):
                for f in CMP_FIELDS:
                    if int(getattr(hits, f)[gi]) != int(w[f]):
                        ok = False
                if not np.array_equal(hits.cigar(gi), w["cigar"]):
                    ok = False
        if not ok:
            bad += 1
            if bad <= 5:
                print(f"  full mismatch read {i} len {len(r)}: want {len(want)} hits, got {len(got_idx)}")
                for w in want:
                    print("    want", {f: w[f] for f in CMP_FIELDS})
                for gi in got_idx:
                    print("    got ", {f: int(getattr(hits, f)[gi]) for f in CMP_FIELDS})
    print(f"[full] reads={len(reads)} oracle_hits={n_hits} gpu_hits={hits.n} mismatching reads={bad} wall={dt:.3f}s stats={al.last_stats}")
    # counting
    for mode in ("basic", "query_length", "matching"):
        counts, ncls, rcls, rbest = al.count(hits, 60, mode)
        print(f"[count:{mode}] n_class={ncls.tolist()} nonzero={int((counts > 0).sum())} sum={int(counts.sum())}")
    return bad == 0


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    mean_len = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
    error = float(sys.argv[3]) if len(sys.argv) > 3 else 0.10
    names, seqs, reads, truth = make_case(n_reads=n_reads, mean_len=mean_len, error=error)
    print("devices:", _lib.lib().mb_device_count())
    oidx = O.Index(names, seqs)
    ok = {}
    for name, fn in [("sketch", lambda: stage_sketch(reads))]:
        try:
            ok[name] = fn()
        except Exception:
            traceback.print_exc(); ok[name] = False
    al = None
    try:
        al = Aligner(names=names, seqs=seqs)
        print("index: mid_occ gpu", al.mid_occ, "oracle", oidx.mid_occ, "n_seq", al.n_seq)
        ok["mid_occ"] = al.mid_occ == oidx.mid_occ
    except Exception:
        traceback.print_exc(); ok["index"] = False
    traces = None
    if al:
        try:
            ok["seed"], traces = stage_seed(al, oidx, reads)
        except Exception:
            traceback.print_exc(); ok["seed"] = False
        if traces is None:
            traces = [oidx.map(r, trace=True) for r in reads]
        for name, fn in [("chain", lambda: stage_chain(al, traces)), ("dp", lambda: stage_dp(al, traces)),
                         ("full", lambda: stage_full(al, oidx, reads, traces))]:
            try:
                ok[name] = fn()
            except Exception:
                traceback.print_exc(); ok[name] = False
    print("SUMMARY", ok)
    return 0 if all(ok.values()) else 1


if __name__ == "__main__":
    sys.exit(main())
