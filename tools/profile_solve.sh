#!/bin/bash
# Run under gpurun: ncu launch list of a short bench run (the SAME command as the bench, kernel arms only)
# + one full capture each of the batched HBM-flavour solve kernel, the shared-memory solve kernel (a block
# on its own), the row-op kernel and the auxiliary kernels.  Outputs land in gpurun_out/;
# tools/summarize_ncu.py (run in the build container, no GPU needed) turns them into profiles/.
set -x
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rqb_solve_kernel -s 10 -c 2 -f -o gpurun_out/prof_solve $CMD > gpurun_out/prof_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rqb_solve_smem -s 3 -c 2 -f -o gpurun_out/prof_smem python tools/kernel_latency.py 4096 1280 6 > gpurun_out/prof_smem.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rqb_rowops -s 3 -c 1 -f -o gpurun_out/prof_rowops $CMD > gpurun_out/prof_rowops.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rqb_lt|rqb_gather|rqb_copy_rows|rqb_repitch" -c 4 -f -o gpurun_out/prof_aux python tools/aux_kernels.py > gpurun_out/prof_aux.log 2>&1
ls -la gpurun_out/
