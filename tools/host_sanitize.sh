#!/bin/bash
# Host side of libmonica_b200.so (FASTQ ingest, routed writers, read packer, database builder, argument checks) under
# AddressSanitizer + UndefinedBehaviorSanitizer: the library is re-built with the sanitizers on the host compiler into a
# scratch copy of the tree and the CPU test files that call it are run against that copy.  CPU only, ~3 min:
#   bash tools/host_sanitize.sh      (prints the number of sanitizer reports; 0 expected)
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
W="$(mktemp -d /tmp/host_san.XXXXXX)"
mkdir -p "$W/repo"
(cd "$ROOT" && tar --exclude=.git --exclude=gpurun_out --exclude=profiles -cf - .) | (cd "$W/repo" && tar xf -)
# nvcc splits -Xcompiler arguments at commas: one -fsanitize per flag
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O1 -std=c++17 -Xcompiler -fPIC -Xcompiler -fsanitize=address \
    -Xcompiler -fsanitize=undefined -Xcompiler -fno-omit-frame-pointer -shared "$ROOT/monica_b200/csrc/monica_b200.cu" \
    -o "$W/repo/monica_b200/lib/libmonica_b200.so" -lz -ldl 2>/dev/null
cd "$W/repo"
ASAN_OPTIONS=detect_leaks=0:halt_on_error=0:protect_shadow_gap=0 UBSAN_OPTIONS=print_stacktrace=1 \
LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
python -m pytest tests/test_boundary_cpu.py tests/test_database_cpu.py tests/test_shard_cpu.py -q -s -p no:cacheprovider > "$W/out.log" 2>&1 || true
tail -1 "$W/out.log"
echo "sanitizer reports: $(grep -c 'AddressSanitizer\|runtime error' "$W/out.log")"
grep -n 'AddressSanitizer\|runtime error' "$W/out.log" | head -20
rm -rf "$W"
