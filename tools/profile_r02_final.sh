#!/bin/bash
# end-of-round-2 profile captures of the final build (run under gpurun on one B200); exports only (the reports stay on the box)
cmd="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline"
$cmd > gpurun_out/r02f_plain.json 2> gpurun_out/r02f_plain.err || exit 1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02f_launches.csv $cmd > gpurun_out/r02f_ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_dp_fast -s 25 -c 3 -o /tmp/r02f_kdp_fast -f $cmd > gpurun_out/r02f_ncu2.log 2>&1
ncu -i /tmp/r02f_kdp_fast.ncu-rep --page raw --csv > gpurun_out/r02f_kdp_fast_raw.csv 2>/dev/null
ncu -i /tmp/r02f_kdp_fast.ncu-rep --page details > gpurun_out/r02f_kdp_fast_details.txt 2>/dev/null
ncu -i /tmp/r02f_kdp_fast.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null; head -c 6000000 /tmp/src.csv > gpurun_out/r02f_kdp_fast_source.csv
ls -la /tmp/*.ncu-rep gpurun_out/ | tail -20
