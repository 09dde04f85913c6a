#!/bin/bash
# The CPU oracle (the parity checker and the reference arm) under AddressSanitizer + UndefinedBehaviorSanitizer: a copy of
# oracle/ is built with -fsanitize=address,undefined in a scratch directory and driven over the test reads (edge reads, empty /
# sub-k / all-N / homopolymer reads), a hard-case bench-style batch, 50 kb reads at 15 % error and a tandem-repeat genome.
# Then the oracle's CPU test files run against the same build.  Any report goes to stderr / is counted; a clean run prints
# the read / hit counts and "sanitizer reports in the test run: 0".  CPU only:  bash tools/oracle_sanitize.sh
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
W="$(mktemp -d /tmp/oracle_san.XXXXXX)"
(cd "$ROOT" && tar --exclude=.git --exclude=gpurun_out --exclude=profiles -cf - .) | (mkdir -p "$W/repo" && cd "$W/repo" && tar xf -)
W="$W/repo"
rm -rf "$W/oracle/_build"
make -s -C "$W/oracle" CFLAGS="-O1 -g -std=gnu99 -fPIC -Wall -Wno-unused-function -ffp-contract=off -msse4.1 -fsanitize=address,undefined -fno-omit-frame-pointer"
cat > "$W/run.py" <<PY
import sys
sys.path.insert(0, "$W"); sys.path.insert(1, "$ROOT"); sys.path.insert(0, "$ROOT/tools/synth")
import numpy as np
from oracle import oracle as O
assert O.__file__.startswith("$W"), O.__file__
from monica_b200 import synth
import mbsynth
mbsynth.build()
names, seqs = synth.make_genomes(11, 3, 60000, strain_frac=0.34)
reads, _ = synth.simulate_reads(12, seqs, 60, 2500, 0.10, junk_frac=0.05)
reads = synth.edge_reads(13, seqs) + reads
reads += [np.frombuffer(b, np.uint8) for b in (b"", b"ACGTACG", b"N" * 500, b"A" * 3000, b"ACGTTGCA" * 400)]
idx = O.Index(names, seqs)
print("test reads", len(reads), "hits", sum(len(idx.map(r)[0]) for r in reads))
n2, s2, gcat, goff = mbsynth.make_genomes(20251018, 4, 300000, strain_frac=0.1)
cat, off, _ = mbsynth.simulate_reads(20251019, gcat, goff, 600, first=0, count=600, n50=8000, error=0.10, sigma=0.6, min_len=500, max_len=0)
O.Index(n2, s2).map_batch_soa(cat, off, n_threads=4)
print("hard-case batch", len(off) - 1, "reads", int(off[-1]), "bases")
n3, s3, gcat, goff = mbsynth.make_genomes(7, 6, 400000, strain_frac=0.34)
cat, off, _ = mbsynth.simulate_reads(8, gcat, goff, 60, first=0, count=60, n50=50000.0, error=0.15, sigma=0.3, min_len=5000, max_len=250000)
O.Index(n3, s3).map_batch_soa(cat, off, n_threads=8)
print("long reads", len(off) - 1, "reads", int(off[-1]), "bases")
rng = np.random.default_rng(1)
unit = rng.integers(0, 4, 300).astype(np.uint8)
g = np.frombuffer(b"ACGT", np.uint8)[np.concatenate([np.tile(unit, 200), rng.integers(0, 4, 50000).astype(np.uint8)])]
i3 = O.Index(["rep:acc"], [g])
print("tandem-repeat reads, hits:", [len(i3.map(g[100:100 + L].copy())[0]) for L in (1000, 7000, 20000)])
PY
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" python "$W/run.py"
# the oracle's own CPU test files (known answers, the Python re-derivations, the C-ABI twin, gloo sharding) on the same build
cd "$W"
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
python -m pytest tests/test_oracle_cpu.py tests/test_oracle_abi_cpu.py tests/test_shard_cpu.py -q -s -p no:cacheprovider > "$W/out.log" 2>&1 || true
tail -1 "$W/out.log"
echo "sanitizer reports in the test run: $(grep -c 'AddressSanitizer\|runtime error' "$W/out.log")"
grep -n 'AddressSanitizer\|runtime error' "$W/out.log" | head -20
rm -rf "$(dirname "$W")"
