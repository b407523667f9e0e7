#!/bin/bash
# usage (under gpurun): tools/sweep.sh "<blocks>:<threads> ..."   -> gpurun_out/sweep.jsonl
: > gpurun_out/sweep.jsonl
for bt in $1; do
  b=${bt%%:*}; t=${bt##*:}
  timeout 300 python bench.py --steps 6 --warmup 3 --blocks $b --threads $t --skip-cpu --skip-rowaxpy >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d=json.loads(l); e=d['e2e']
    print("blocks %3d threads %2d | value %7.1f Gbit/s  ms/step %.2f roofline %.2f | e2e %6.1f Gbit/s ms/step %.1f phases %s" % (d['config']['blocks_per_gpu'], e['host_threads'], d['value'], d['ms_per_step'], d['roofline']['frac'], e['value'], e['ms_per_step'], {k:round(v,2) for k,v in e['phase_seconds_summed_over_threads'].items()}))
PY
