#!/bin/bash
# usage (under gpurun): tools/e2e_quick.sh "<threads> ..." [extra bench args]   -- e2e arm only, one line per setting
T="$1"; shift
for t in $T; do
  python bench.py --steps 8 --warmup 3 --threads $t --skip-cpu --skip-rowaxpy "$@" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; n=d['config']['blocks_per_gpu']*d['steps']
print('threads', e['host_threads'], 'value %.0f e2e %.1f Gbit/s ms/step %.1f' % (d['value'], e['value'], e['ms_per_step']), {k:round(1e3*v/n,2) for k,v in e['phase_seconds_summed_over_threads'].items()})"
done
