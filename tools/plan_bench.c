/* plan_bench.c -- where the host planner's time goes (rqb_plan_build with -DRQB_PLAN_FINE): K loss smem [threads].
 * Built and run by tools/plan_bench.sh; host only, no GPU involved. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "rqb_planner.h"
#include "rqb_program.h"
extern double rqb_plan_fine[24];
static double now(void){struct timespec ts;clock_gettime(CLOCK_MONOTONIC,&ts);return ts.tv_sec+1e-9*ts.tv_nsec;}
int main(int argc,char**argv){
  int K=atoi(argv[1]); double loss=atof(argv[2]); int smem=atoi(argv[3]);
  rqb_params P; rqb_params_init(K,&P);
  int Kp=P.Kprime; uint32_t *isi=malloc(4*(Kp+8)),*in_row=malloc(4*(Kp+8)),*miss=malloc(4*K);
  double tot=0; int n=0; 
  for(int rep=0;rep<40;rep++){
    srand(rep+1); int nm=0,nr=0;
    for(int e=0;e<Kp;e++){ if(e>=K){isi[e]=e;in_row[e]=RQB_ROW_NONE;} else if((rand()/(double)RAND_MAX)<loss){ isi[e]=Kp+nr; in_row[e]=Kp+nr; nr++; miss[nm++]=e;} else {isi[e]=e;in_row[e]=e;} }
    rqb_plan_request rq={K,0,isi,in_row,0,nm,miss,(uint32_t)(Kp+nr+4),(uint32_t)K,NULL,0,NULL,smem?RQB_SMEM_BUDGET_BYTES:0};
    rqb_plan*p=NULL; if(rep==4){memset(rqb_plan_fine,0,sizeof(rqb_plan_fine));tot=0;n=0;}
    double t0=now(); int rc=rqb_plan_build(&rq,&p); double t1=now();
    if(rc==0){tot+=t1-t0;n++; rqb_plan_free(p);} 
  }
  printf("K=%d smem=%d mean %.3f ms over %d\n",K,smem,1e3*tot/n,n);
  const char*nm[]={"3a G+sched","3b low rows","3c HDPC schur","3d GJ","3e HDPC solve","1+2 matrix+peel","4A tri","4B low+scan","4C1","4C2","4C3/4","4F tables","4E","4O out","write_pages","S load+needed","S chains A+B","S scan+C","S tables+TAB","S outputs","1 rows","1 column lists"};
  for(int k=0;k<22;k++) printf("  %-16s %.3f ms\n",nm[k],1e3*rqb_plan_fine[k]/n);
}
