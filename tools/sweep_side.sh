for n in 5 6; do
  sed -i "s/__launch_bounds__(DPC2_THREADS, [0-9])/__launch_bounds__(DPC2_THREADS, $n)/" monica_b200/csrc/align_cta.cuh
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared monica_b200/csrc/monica_b200.cu -o monica_b200/lib/libmonica_b200.so -lz -ldl 2>/dev/null
  MB_DEBUG=1 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline 2>gpurun_out/lb_$n.err | grep "^{" > gpurun_out/lb_$n.json
done
