#!/bin/bash
# usage (under gpurun): tools/sweep.sh "<blocks>:<threads> ..."   -> gpurun_out/sweep.jsonl
# e2e arm only, with the host time accounting of the nanorq.h layer switched on
: > gpurun_out/sweep.jsonl
export NANORQ_B200_PROFILE=1
for bt in $1; do
  b=${bt%%:*}; t=${bt##*:}
  timeout 300 python bench.py --steps 6 --warmup 3 --blocks $b --threads $t --skip-cpu --skip-rowaxpy >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done
python - <<'PY'
import json
for l in open('gpurun_out/sweep.jsonl'):
    d=json.loads(l); e=d['e2e']
    n=d['config']['blocks_per_gpu']*d['steps']
    print("blocks %3d threads %2d | value %7.1f Gbit/s | e2e %6.1f Gbit/s ms/step %.1f | per block ms: %s" % (d['config']['blocks_per_gpu'], e['host_threads'], d['value'], e['value'], e['ms_per_step'], {k:round(1e3*v/n,2) for k,v in e['phase_seconds_summed_over_threads'].items()}))
    hp=e.get('host_profile_seconds_summed_over_threads')
    if hp: print("      host profile, ms per block:", {k:round(1e3*v/n,3) for k,v in hp.items() if v>0})
PY
