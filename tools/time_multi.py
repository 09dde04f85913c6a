# times multi_threaded_aligner over three FASTQ files with 1 and 3 threads (diagnostic; prints per-file stage times with MB_DEBUG=1)
import os, sys, time, tempfile, shutil, contextlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.synth import mbsynth
from monica_b200 import aligner as galigner
from monica_b200.mappy_shim import Aligner
names, seqs, gcat, goff = mbsynth.make_genomes(20251018, 10, 5_000_000, strain_frac=0.1)
cat, off, _ = mbsynth.simulate_reads(20251019, gcat, goff, 90000, 8000, 0.10)
al = Aligner(names=names, seqs=seqs, preset="map-ont", best_n=15)
for nt in (1, 3, 3):
    tmp = tempfile.mkdtemp(prefix="mb_multi_", dir="/dev/shm")
    per = 30000
    for f in range(3):
        lo = f * per
        with open(os.path.join(tmp, f"s{f}.fastq"), "wb") as fh:
            qual = b"I" * int(np.diff(off[lo:lo + per + 1]).max())
            for i in range(lo, lo + per):
                sq = cat[off[i]:off[i + 1]].tobytes()
                fh.write(b"@read%d ch=%d\n" % (i, i % 512) + sq + b"\n+\n" + qual[:len(sq)] + b"\n")
    cwd = os.getcwd()
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(sys.stderr):
        galigner.multi_threaded_aligner(tmp, ["resident"], mode="query_length", n_threads=nt, output_folder=tmp, index_loader_fn=lambda p: al)
    dt = time.perf_counter() - t0
    os.chdir(cwd)
    print(f"n_threads={nt}: {dt:.3f} s for {int(off[90000])/1e9:.3f} Gbases", file=sys.stderr)
    shutil.rmtree(tmp, ignore_errors=True)
