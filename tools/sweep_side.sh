for sc in 0.5 0.25 0.12; do
  MB_DEBUG=1 MB_SIDE_SCALE=$sc python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/sx_$sc.err | grep "^{" > gpurun_out/sx_$sc.json
done
