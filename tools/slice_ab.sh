#!/bin/bash
# under gpurun: slice width A/B (kernel arm at several batch sizes + single-block latency)
for sl in 128 256 64; do
  export NANORQ_B200_SLICE=$sl
  echo "== slice $sl"
  python tools/kernel_latency.py 4096 1280 12 | cut -c1-90
  python tools/kernel_latency.py 1024 1280 12 | head -1 | cut -c1-90
  for b in ${1:-118}; do
    python bench.py --steps 6 --warmup 3 --blocks $b --skip-cpu --skip-rowaxpy --skip-e2e | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   blocks %3d value %7.1f Gbit/s ms/step %.2f' % (d['config']['blocks_per_gpu'], d['value'], d['ms_per_step']))"
  done
done
