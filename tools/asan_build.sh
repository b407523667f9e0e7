#!/bin/bash
# Host code of the library under AddressSanitizer + UBSan (the CUDA object is reused as built):
#   tools/asan_build.sh          -> nanorq_b200/build/asan/libnanorq_b200.so
# run tests against it with
#   LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:protect_shadow_gap=0 \
#   NANORQ_B200_LIBDIR=$PWD/nanorq_b200/build/asan python -m pytest tests/test_fuzz_gpu.py -x -q
set -e
cd "$(dirname "$0")/.."
python -m nanorq_b200.build > /dev/null
D=nanorq_b200/build/asan; mkdir -p $D
OBJS=""
for f in rqb_planner rqb_solver nanorq_api rqb_io; do
  gcc -O1 -g -fno-omit-frame-pointer -fsanitize=address,undefined -fno-sanitize-recover=undefined -march=x86-64-v3 -std=c11 -Wall -Wextra -fPIC -pthread \
      -Inanorq_b200/csrc -Iinclude -c nanorq_b200/csrc/$f.c -o $D/$f.o
  OBJS="$OBJS $D/$f.o"
done
g++ -shared -fsanitize=address,undefined -o $D/libnanorq_b200.so $OBJS nanorq_b200/build/rqb_device.cu.o \
    -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
for h in rq_roundtrip rq_roundtrip_batch; do
  gcc -O1 -g -fsanitize=address,undefined -std=c11 -fPIC -shared -pthread -o $D/lib$h.so bench/$h.c -Iinclude -L$D -lnanorq_b200 -Wl,-rpath,'$ORIGIN'
done
echo $D/libnanorq_b200.so
