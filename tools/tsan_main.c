#include <stdio.h>
#include <stdlib.h>
typedef struct { int K, T, nblocks; double loss; int overhead; unsigned seed; int nthreads, precalc, verify, zblocks; } rt_config;
typedef struct { double wall_s, t_gen, t_emit, t_add, t_repair; long n_lost, n_sent; int retries, failures, mismatches; unsigned long long out_fnv; } rt_result;
int rq_roundtrip_run(const rt_config *, rt_result *);
int rq_roundtrip_batch_run(const rt_config *, rt_result *);
void rqb_set_plan_threads(int n); /* rqb200.h */
int main(int argc, char **argv) {
  (void)argc; (void)argv;
  static const int shapes[][5] = {{1024, 1280, 64, 8, 4}, {10, 64, 2048, 8, 16}, {300, 104, 256, 8, 2}, {4096, 1280, 32, 8, 4}, {100, 1280, 512, 8, 8}};
  for (int s = 0; s < 10; s++) {
    const int *sh = shapes[s % 5];
    rt_config c = {sh[0], sh[1], sh[2], s < 5 ? 0.1 : 0.3, 1, 100 + s, sh[3], 1, 1, sh[4]};
    rt_result r;
    rqb_set_plan_threads(s < 5 ? 1 : 4); /* nanorq_repair_blocks: the blocks of an object analysed side by side */
    int rc = rq_roundtrip_run(&c, &r);
    printf("per-symbol rc %d failures %d mismatches %d\n", rc, r.failures, r.mismatches);
    rc = rq_roundtrip_batch_run(&c, &r);
    printf("batch rc %d failures %d mismatches %d\n", rc, r.failures, r.mismatches);
  }
  return 0;
}
