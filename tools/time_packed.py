"""Where does tests/test_gpu_parity.py::test_packed_reads_map_like_ascii_reads spend its time?  (development aid, run under gpurun)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from monica_b200 import synth  # noqa: E402
from monica_b200.mappy_shim import Aligner  # noqa: E402

T0 = time.perf_counter()


def mark(what):
    global T0
    t = time.perf_counter()
    print(f"{what:40s} {1e3 * (t - T0):9.1f} ms", flush=True)
    T0 = t


rng = np.random.default_rng(9)
names, seqs = synth.make_genomes(5, 3, 200_000)
reads, _ = synth.simulate_reads(6, seqs, 1500, 3000.0, 0.10)
mark("simulate")
reads = [np.frombuffer(bytes(r), np.uint8).copy() for r in reads]
kinds = sys.argv[1] if len(sys.argv) > 1 else "01234"
for i in range(0, len(reads), 7):
    r = reads[i]
    if len(r) < 400:
        continue
    k = i // 7 % 5
    if str(k) not in kinds:
        continue
    if k == 0:
        s = int(rng.integers(0, len(r) - 300)); r[s:s + int(rng.integers(1, 300))] = ord("N")
    elif k == 1:
        r[rng.integers(0, len(r), 20)] = np.frombuffer(b"RYKMSWn-", np.uint8)[rng.integers(0, 8, 20)]
    elif k == 2:
        r[:] = np.frombuffer(bytes(r).lower(), np.uint8)
    elif k == 3:
        r[r == ord("T")] = ord("U")
    else:
        r[-17:] = ord("N"); r[:3] = ord("N")
cat, off = synth.concat_reads(reads)
al = Aligner(names=names, seqs=seqs, preset="map-ont", best_n=15)
mark("index")
pk = Aligner.pack_reads(cat, off, n_threads=3)
mark("pack")
for rep in range(2):
    base = al.map_batch(cat=cat, off=off)
    mark("map_batch ascii")
    print({k: round(v, 2) if isinstance(v, float) else v for k, v in al.last_stats.items() if k.startswith("ms_") or k in ("n_exact_tasks", "n_dp_pass2", "n_rounds", "n_inv", "n_band_tasks")})
for rep in range(2):
    al.map_packed(pk)
    mark("map_packed")
os.environ["MB_FEED_MIN_BYTES"] = "1"
al.map_packed(pk); mark("map_packed feed")
os.environ["MB_PIECE_BASES"] = "700001"
al.map_packed(pk); mark("map_packed feed+pieces")
del os.environ["MB_FEED_MIN_BYTES"]
al.map_packed(pk); mark("map_packed pieces")
