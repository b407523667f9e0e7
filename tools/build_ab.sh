#!/bin/bash
# builds complete library variants (host C flags) into nanorq_b200/build/ab_<name>/ for tools/ab_e2e.py:
#   tools/build_ab.sh "name=-DWINDOW_DIV=8u other=-DUPLOAD_CHUNK=4096u"
set -e
python -m nanorq_b200.build >/dev/null
for v in $1; do
  name=${v%%=*}; flags=$(echo "${v#*=}" | tr ',' ' ')
  d=nanorq_b200/build/ab_$name; mkdir -p $d
  objs=""
  for f in rqb_planner rqb_solver nanorq_api rqb_io; do
    gcc -O3 -march=x86-64-v3 -std=c11 -fPIC -pthread $flags -c nanorq_b200/csrc/$f.c -o $d/$f.o -Inanorq_b200/csrc -Iinclude
    objs="$objs $d/$f.o"
  done
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libnanorq_b200.so $objs nanorq_b200/build/rqb_device.cu.o -lpthread
  gcc -O2 -std=c11 -fPIC -shared -pthread -o $d/librq_roundtrip.so bench/rq_roundtrip.c -Iinclude -L$d -lnanorq_b200 -Wl,-rpath,'$ORIGIN'
  rm -f $d/*.o
  echo "built $d ($flags)"
done
