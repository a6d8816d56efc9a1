/* oracle/_ref/so_ref_timed: the UNMODIFIED reference objects (kd2.o smooth2.o nr.o cosmo.o
 * romberg.o compiled from /root/reference) driven by this small main instead of so.c's, so the
 * two hot-path calls can be timed separately:
 *     kdBuildTree(kd)            (so.c:515 -> kd2.c:1096-1185)
 *     kdSO(kd, fThreshold, 1028) (so.c:540 -> kd2.c:864-895)
 * TEST / BASELINE INFRASTRUCTURE ONLY (bench.py --impl reference and cpu_baseline).
 *
 * usage: so_ref_timed <snapshot.tipsy> <halos.gtp> <rho_threshold> <nMembers> <period> [out.sogtp-base]
 * prints one JSON line:
 *   {"n":..,"h":..,"t_read":..,"t_build":..,"t_gtp":..,"t_so":..,"n_ok":..}
 */
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "kd2.h"   /* from -I/root/reference */

static double now(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char **argv)
{
    KD kd;
    float fPeriod[3], fCenter[3] = {0, 0, 0};
    double t0, t1, t2, t3, t4;
    FILE *fp;
    int i, nok = 0, n, h;
    float thr;
    if (argc < 6) { fprintf(stderr, "usage: %s snap gtp thr nMembers period [outbase]\n", argv[0]); return 2; }
    thr = (float)atof(argv[3]);
    fPeriod[0] = fPeriod[1] = fPeriod[2] = (float)atof(argv[5]);
    kdInit(&kd, 16, fPeriod, fCenter, 0, atoi(argv[4]), 1, 0, 0, 0, 0, 0);
    fp = fopen(argv[1], "rb");
    if (!fp) { perror(argv[1]); return 1; }
    t0 = now();
    n = kdReadTipsy(kd, fp, 0);
    fclose(fp);
    kdSetUniverse(kd, 1.0f, 1.0f, 0.0f, 2.8944f, 0.0f, -9.9f, -9.9f);
    t1 = now();
    kdBuildTree(kd);
    t2 = now();
    h = kdReadGTPList(kd, argv[2], NULL, 0.0f, 0);
    t3 = now();
    kdSO(kd, thr, 1028);
    t4 = now();
    for (i = 0; i < kd->nGrps; ++i) if (kd->grps[i].fMvir > 0) ++nok;
    if (argc > 6) kdWriteGTP(kd, argv[6], 0);
    printf("{\"n\":%d,\"h\":%d,\"t_read\":%.6f,\"t_build\":%.6f,\"t_gtp\":%.6f,\"t_so\":%.6f,\"n_ok\":%d}\n",
           n, h, t1 - t0, t2 - t1, t3 - t2, t4 - t3, nok);
    return 0;
}
