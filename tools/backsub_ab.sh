#!/bin/bash
# under gpurun: compare the back-substitution variants (kernel arm + single-block latency)
for m in triangular tables; do
  export NANORQ_B200_BACKSUB=$m
  echo "== $m"
  python tools/kernel_latency.py 4096 1280 12
  python bench.py --steps 6 --warmup 3 --skip-cpu --skip-rowaxpy --skip-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('   value %7.1f Gbit/s ms/step %.2f frac %.2f program_gbs %.0f' % (d['value'], d['ms_per_step'], r['frac'], r['program_gbs']))"
done
