/* Instrumentation hooks for oracle/_ref/so_ref_inst (the reference compiled from
 * /root/reference with two one-line sed insertions, see oracle/Makefile).
 * TEST INFRASTRUCTURE ONLY: this pins N_Delta and the r^2-sorted member list, which the
 * reference computes (kd2.c:814-823) but never writes to any output file.
 *
 * If $SO_INST_FILE is set, every successful kdRvir appends a record
 *   int32 index, int32 j, int32 iOrder[j], float fDist2[j]
 * to that file.  At exit the counters are printed to stderr as
 *   SO_INST ndist=<n> ngather=<n> */
#include <stdio.h>
#include <stdlib.h>
#include "smooth2.h"   /* from -I/root/reference */

long so_inst_ndist = 0;
long so_inst_ngather = 0;
static FILE *g_fp = NULL;
static int g_init = 0;

void so_inst_hit(void *vgrp, void *vsmx, int j)
{
    GRPNODE *grp = (GRPNODE *)vgrp;
    SMX smx = (SMX)vsmx;
    int k;
    if (!g_init) {
        const char *p = getenv("SO_INST_FILE");
        g_init = 1;
        if (p && *p) g_fp = fopen(p, "wb");
    }
    if (!g_fp) return;
    fwrite(&grp->index, sizeof(int), 1, g_fp);
    fwrite(&j, sizeof(int), 1, g_fp);
    for (k = 0; k < j; ++k) fwrite(&smx->nnList[k].pInit->iOrder, sizeof(int), 1, g_fp);
    for (k = 0; k < j; ++k) fwrite(&smx->nnList[k].fDist2, sizeof(float), 1, g_fp);
}

__attribute__((destructor)) static void so_inst_report(void)
{
    if (g_fp) fclose(g_fp);
    fprintf(stderr, "SO_INST ndist=%ld ngather=%ld\n", so_inst_ndist, so_inst_ngather);
}
