#!/bin/bash
# usage (under gpurun): tools/variants.sh "<variant names>" "<blocks list>"
for v in $1; do
  cp nanorq_b200/build/variants/lib_$v.so nanorq_b200/libnanorq_b200.so
  echo "== variant $v"
  python tools/kernel_latency.py 4096 1280 12
  for b in $2; do
    timeout 300 python bench.py --steps 6 --warmup 3 --blocks $b --skip-cpu --skip-rowaxpy --skip-e2e 2>>gpurun_out/variants.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('   blocks %3d value %7.1f Gbit/s ms/step %.2f frac %.2f' % (d['config']['blocks_per_gpu'], d['value'], d['ms_per_step'], d['roofline']['frac']))"
  done
done
