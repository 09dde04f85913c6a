#!/bin/bash
# end-of-round evidence: plain runs first, then the ncu passes of the same commands (see /opt/skills/guides/B200_PROFILING.md)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --reads 20000 --no-cpu-baseline --steps 3 --warmup 3 > gpurun_out/bench_20k.json 2>/dev/null
python bench.py --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench100k.csv python bench.py --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_l100.log 2>&1
python bench.py --reads 20000 --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_bench20k.csv python bench.py --reads 20000 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_l20.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_dp_fast<.int.7>" -c 1 -o gpurun_out/prof_dpfast_final python bench.py --reads 20000 --no-cpu-baseline --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
tail -c 600 gpurun_out/bench_default.json
