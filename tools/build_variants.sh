#!/bin/bash
# builds kernel variants for tools/variants.sh:  tools/build_variants.sh "name=-DFLAG=1,-DOTHER=2 ..."
set -e
mkdir -p nanorq_b200/build/variants
python -m nanorq_b200.build >/dev/null
cp nanorq_b200/libnanorq_b200.so nanorq_b200/build/variants/lib_base.so
for v in $1; do
  name=${v%%=*}; flags=$(echo "${v#*=}" | tr ',' ' ')
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v \
     $flags -c nanorq_b200/csrc/rqb_device.cu -o /tmp/rqb_device_$name.o \
     -Inanorq_b200/csrc -Iinclude 2>&1 | grep -A2 "rqb_solve_kernel" | grep -E "registers|spill" | tr '\n' ' '
  echo " <- $name ($flags)"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o nanorq_b200/build/variants/lib_$name.so \
     nanorq_b200/build/rqb_planner.c.o nanorq_b200/build/rqb_solver.c.o nanorq_b200/build/nanorq_api.c.o nanorq_b200/build/rqb_io.c.o /tmp/rqb_device_$name.o -lpthread
done
