for n in 1 2 3; do
  MB_BAND_PER_SM=$n MB_DEBUG=1 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/bp_$n.err | grep "^{" > gpurun_out/bp_$n.json
done
