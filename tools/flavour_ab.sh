#!/bin/bash
# A/B of the two flavours of the solve program on the GPU: single-block latency and the
# batched device-resident figure (no e2e, no CPU baseline).  Output: gpurun_out/flavour_ab.log
out=gpurun_out/flavour_ab.log
: > $out
for fl in smem hbm; do
  for sl in 64 32; do
    [ $fl = hbm ] && [ $sl != 64 ] && continue
    echo "== flavour=$fl max_slice=$sl" >> $out
    for kt in "1024 1280" "4096 1280"; do
      NANORQ_B200_FLAVOUR=$fl NANORQ_B200_SMEM_MAXSLICE=$sl python tools/kernel_latency.py $kt 12 >> $out 2>&1
    done
    NANORQ_B200_FLAVOUR=$fl NANORQ_B200_SMEM_MAXSLICE=$sl python bench.py --steps 10 --warmup 3 --skip-e2e --skip-cpu --skip-rowaxpy 2>&1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('bench value %.1f Gbit/s, %.3f ms/step' % (d['value'], d['ms_per_step']))
    else:
        print(ln.rstrip())
" >> $out 2>&1
  done
done
cat $out
