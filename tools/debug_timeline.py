"""MB_DEBUG timelines of the bench workload (configs[1]) on one GPU: one full batch and one 4,000-read streaming batch.
Run under gpurun: python tools/debug_timeline.py [n_reads]  (stderr carries the [mb] phase / dp launch lines)."""
import argparse
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools", "synth"))
import bench  # noqa: E402
from monica_b200.mappy_shim import Aligner  # noqa: E402


def main():
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    a = argparse.Namespace(config=1, gpus=1, genomes=0, genome_len=0, reads=n_reads, plain_reads=False)
    c = bench.config_of(a)
    names, seqs, gcat, goff, cat, off, cls, seeds = bench.make_data(c, 20251018, 0, 1)
    al = Aligner(names=names, seqs=seqs, device=0)
    for label, n in (("full batch", len(off) - 1), ("streaming batch", 4000)):
        o = off[:n + 1]
        cc = cat[:o[-1]]
        for _ in range(2):
            al.map_batch(cat=cc, off=o)
        t0 = time.perf_counter()
        al.map_batch(cat=cc, off=o)
        print(f"== {label}: {n} reads, {int(o[-1])} bases, warm wall {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr, flush=True)
        os.environ["MB_DEBUG"] = "1"
        al.map_batch(cat=cc, off=o)
        del os.environ["MB_DEBUG"]
        st = al.last_stats
        print("== stats", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()}, file=sys.stderr, flush=True)


if __name__ == "__main__":
    main()
