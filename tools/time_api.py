"""Time the stages of the FASTQ-facing API on the GPU box (diagnostic)."""
import ctypes as C, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools", "synth"))
import mbsynth
from monica_b200 import _lib
from monica_b200.mappy_shim import Aligner
L = _lib.lib()
names, seqs, gcat, goff = mbsynth.make_genomes(7, 10, 5_000_000, strain_frac=0.1)
cat, off, cls = mbsynth.simulate_reads(9, gcat, goff, 20000, 8000.0)
q = "/dev/shm/t_sample.fastq"
with open(q, "wb") as fh:
    qual = b"I" * int(np.diff(off).max())
    for i in range(20000):
        sq = cat[off[i]:off[i + 1]].tobytes()
        fh.write(b"@read%d ch=%d\n" % (i, i % 512) + sq + b"\n+\n" + qual[:len(sq)] + b"\n")
al = Aligner(names=names, seqs=seqs, device=0)
for rep in range(3):
    t = time.perf_counter(); fq = C.c_void_p(); _lib.check(L.mb_fastq_load(q.encode(), C.byref(fq))); t1 = time.perf_counter() - t
    n = int(L.mb_fastq_n(fq))
    offp = C.POINTER(C.c_int64)(); catp = L.mb_fastq_seqs(fq, C.byref(offp))
    o = np.ctypeslib.as_array(offp, shape=(n + 1,)); c = np.ctypeslib.as_array(catp, shape=(int(o[-1]),))
    t = time.perf_counter(); hits = al.map_batch(cat=c, off=o, cigars=False); t2 = time.perf_counter() - t
    t = time.perf_counter(); r = al.count(hits, 60, None); t3 = time.perf_counter() - t
    print(f"rep {rep}: load {t1:.3f}  map_batch {t2:.3f} (gpu ms_total {al.last_stats['ms_total']:.1f}, h2d {al.last_stats['ms_h2d']:.1f})  count {t3:.3f}")
    L.mb_fastq_free(fq)
os.remove(q)
