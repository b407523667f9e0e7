/* rqb_prof.h -- optional host time accounting (NANORQ_B200_PROFILE=1): where the
 * threads of the nanorq.h layer spend their time, summed over threads. */
#ifndef RQB_PROF_H
#define RQB_PROF_H

enum {
  RQB_PF_GEN_LOAD, RQB_PF_GEN_UPLOAD, RQB_PF_GEN_PLAN, RQB_PF_GEN_RUN, RQB_PF_GEN_SYNC, RQB_PF_EMIT_SRC,
  RQB_PF_EMIT_WINDOW, RQB_PF_ADD_CREATE, RQB_PF_ADD_COPY, RQB_PF_ADD_WRITE, RQB_PF_REP_UPLOAD, RQB_PF_REP_REQUEST,
  RQB_PF_REP_PLAN, RQB_PF_REP_PAGES, RQB_PF_REP_ARGS, RQB_PF_REP_RUN, RQB_PF_REP_FETCH, RQB_PF_REP_WRITE,
  RQB_PF_FREE, RQB_PF_RANGE_GEN, RQB_PF_RANGE_QUEUE, RQB_PF_RANGE_WAIT, RQB_PF_ADDS, RQB_PF_COUNT
};

int rqb_prof_enabled(void);
double rqb_prof_now(void);
void rqb_prof_add(int slot, double seconds);

/* PF_T0; ... PF(slot); ... PF(slot2);  -- each PF charges the time since the previous mark */
#define PF_T0 double pf_t = rqb_prof_enabled() ? rqb_prof_now() : 0.0
#define PF(slot)                          \
  do {                                    \
    if (rqb_prof_enabled()) {             \
      double _n = rqb_prof_now();         \
      rqb_prof_add((slot), _n - pf_t);    \
      pf_t = _n;                          \
    }                                     \
  } while (0)

#endif
