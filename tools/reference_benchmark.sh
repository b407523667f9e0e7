#!/bin/bash
# The reference's own headline benchmark (Makefile:35-45 of the reference: ./benchmark 1280 K 5.0, graph.png), run
# with the UNMODIFIED benchmark.c linked against the reference (benchmark_ref) and against libnanorq_b200.so
# (benchmark_b200); both built by oracle/Makefile.  One core, one block at a time, Mibit/s as the program prints
# them: K, encode, precalc-encode, decode (loss pct, overhead 0), decode-oh.  Run on a GPU box:
#   tools/reference_benchmark.sh > gpurun_out/reference_benchmark.txt
B=oracle/_ref/bin
run() { # retries: a time-seeded loss pattern with zero overhead is singular now and then (both builds exit 1)
  for a in 1 2 3 4; do out=$($B/$1 $2 $3 $4 2>/dev/null) && { echo "$out"; return; }; done; echo "failed"
}
echo "# T K pct | build | K encode precalc decode decode-oh [Mibit/s]"
for cfg in "1280 100 5.0" "1280 500 5.0" "1280 1000 5.0" "1280 1024 5.0" "1280 4096 5.0" "1280 5000 5.0" "1280 10000 5.0" "1280 50000 5.0" "512 56403 5.0" "64 10 0"; do
  set -- $cfg
  echo "$cfg | reference | $(run benchmark_ref $1 $2 $3)"
  echo "$cfg | b200      | $(run benchmark_b200 $1 $2 $3)"
done
