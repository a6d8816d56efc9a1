/* Declarations injected (gcc -include) into the sed-instrumented build of the reference
 * (oracle/_ref/so_ref_inst).  TEST INFRASTRUCTURE ONLY — see oracle/Makefile. */
#ifndef SO_ORACLE_INST_H
#define SO_ORACLE_INST_H
extern long so_inst_ndist;    /* r^2 evaluations in smBallGather's leaf loop (smooth2.c:88-106) */
extern long so_inst_ngather;  /* smBallGather calls */
void so_inst_hit(void *grp, void *smx, int j);   /* called where kdRvir succeeds (kd2.c:823) */
#endif
