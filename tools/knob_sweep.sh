#!/bin/bash
# scheduling knob sweep on the bench workload (run under gpurun): one short bench per setting, ms per step of each
run() {
	env "$@" python bench.py --steps 3 --warmup 2 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'ms_per_step %.2f' % d['ms_per_step'], 'each', [round(x,1) for x in d['ms_each_step_rank0']], 'e2e_ms %.2f' % d['e2e']['ms_per_step'], 'kdp_fast %.1f exact %.1f chain %.2f seed %.2f sketch %.2f' % (d['stage_ms']['ms_kdp_fast'], d['stage_ms']['ms_kdp_exact'], d['stage_ms']['ms_chain'], d['stage_ms']['ms_seed'], d['stage_ms']['ms_sketch']), flush=True)
"
}
for s in "$@"; do run $s; done
