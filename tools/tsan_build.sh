#!/bin/bash
# Host code of the library + both round-trip harnesses under ThreadSanitizer, as one binary (the CUDA object is
# reused as built):   tools/tsan_build.sh  ->  nanorq_b200/build/tsan/tsan_roundtrip
# run on a GPU box:   TSAN_OPTIONS="ignore_noninstrumented_modules=1 halt_on_error=0" nanorq_b200/build/tsan/tsan_roundtrip
# (8 threads, five block shapes from K=10 to K=4096, per-symbol and batch arm; NANORQ_B200_CACHE_MB=8 adds the
# eviction paths)
set -e
cd "$(dirname "$0")/.."
python -m nanorq_b200.build > /dev/null
D=nanorq_b200/build/tsan; mkdir -p $D
OBJS=""
for f in rqb_planner rqb_solver nanorq_api rqb_io; do
  gcc -O1 -g -fsanitize=thread -march=x86-64-v3 -std=c11 -fPIC -pthread -Inanorq_b200/csrc -Iinclude -c nanorq_b200/csrc/$f.c -o $D/$f.o
  OBJS="$OBJS $D/$f.o"
done
gcc -O1 -g -fsanitize=thread -std=c11 -pthread -Iinclude -c bench/rq_roundtrip.c -o $D/rt.o
gcc -O1 -g -fsanitize=thread -std=c11 -pthread -Iinclude -c bench/rq_roundtrip_batch.c -o $D/rtb.o
gcc -O1 -g -fsanitize=thread -c tools/tsan_main.c -o $D/main.o
g++ -fsanitize=thread -o $D/tsan_roundtrip $D/main.o $D/rt.o $D/rtb.o $OBJS nanorq_b200/build/rqb_device.cu.o \
    -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
echo $D/tsan_roundtrip
