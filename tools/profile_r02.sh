#!/bin/bash
# round-2 profile captures (run under gpurun on one B200); outputs under gpurun_out/ (the merge back is limited to 64 MiB:
# the reports are exported to CSV on the box and only the exports travel)
cmd="python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline"
$cmd > gpurun_out/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches.csv $cmd > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_dp_fast -s 25 -c 3 -o /tmp/r02_kdp_fast -f $cmd > gpurun_out/ncu2.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_chain_dp|k_seed_lookup|k_sketch_par|k_dp_cta2|k_dp_band|k_dp_ext|k_sort_anchors|k_update_extra|k_ztest_ll' -c 16 -o /tmp/r02_others -f $cmd > gpurun_out/ncu3.log 2>&1
ncu -i /tmp/r02_kdp_fast.ncu-rep --page raw --csv > gpurun_out/r02_kdp_fast_raw.csv 2>/dev/null
ncu -i /tmp/r02_kdp_fast.ncu-rep --page details > gpurun_out/r02_kdp_fast_details.txt 2>/dev/null
ncu -i /tmp/r02_kdp_fast.ncu-rep --page source --csv > /tmp/src.csv 2>/dev/null; head -c 12000000 /tmp/src.csv > gpurun_out/r02_kdp_fast_source.csv
ncu -i /tmp/r02_others.ncu-rep --page raw --csv > gpurun_out/r02_others_raw.csv 2>/dev/null
ncu -i /tmp/r02_others.ncu-rep --page details > gpurun_out/r02_others_details.txt 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/
