#!/bin/bash
# Run under gpurun: ncu launch list of a short bench run (the SAME command as the bench, kernel arm only)
# + one full capture of the solve kernel and of the row-op kernel.
# Outputs land in gpurun_out/; tools/summarize_ncu.py (run here, no GPU needed) turns them into profiles/.
set -x
CMD="python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rqb_solve -s 8 -c 2 -f -o gpurun_out/prof_solve $CMD > gpurun_out/prof_solve.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rqb_rowops -s 3 -c 1 -f -o gpurun_out/prof_rowops $CMD > gpurun_out/prof_rowops.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rqb_lt|rqb_gather" -c 3 -f -o gpurun_out/prof_aux python tools/aux_kernels.py > gpurun_out/prof_aux.log 2>&1
ls -la gpurun_out/
