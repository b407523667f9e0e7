#!/bin/bash
# usage: tools/plan_bench.sh [K loss]   -- per-phase times of the planner on this host, both program flavours
K=${1:-4096}; LOSS=${2:-0.1}
D=$(mktemp -d)
gcc -O3 -march=x86-64-v3 -std=c11 -DRQB_PLAN_FINE -fPIC -shared -pthread -Inanorq_b200/csrc -Iinclude -o $D/libplanfine.so nanorq_b200/csrc/rqb_planner.c
gcc -O2 -Inanorq_b200/csrc -Iinclude -o $D/plan_bench tools/plan_bench.c $D/libplanfine.so -Wl,-rpath,$D -lpthread
$D/plan_bench $K $LOSS 1; $D/plan_bench $K $LOSS 0
rm -rf $D
