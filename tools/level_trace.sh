#!/bin/bash
# under gpurun: rebuild the library with the level trace, run it, restore the normal build
RQB_NVCC_EXTRA=-DRQB_TRACE python -m nanorq_b200.build --force > /dev/null
python tools/level_trace.py 4096 1280 > gpurun_out/level_trace.log 2>&1
python -m nanorq_b200.build --force > /dev/null
cat gpurun_out/level_trace.log
