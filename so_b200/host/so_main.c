/* so_main.c — drop-in `so`: same command line, same input files, same output files as
 * /root/reference/so.c:192-575, with kdBuildTree / kdSO running on a B200 (see kd_so.c).
 *
 *   so -i <.gtp> [-o base] [-delta D] [-O Omega0] [-L] [-z z] [-m nMembers] [-M minMass] [-p period]
 *      [-c c | -cx -cy -cz] [-std] [-list file] [-stat file] [-mark file] [-dark -gas -star | -all]
 *      [-grp] [-gtp] [-subsumed] [-ignored] [-u massunit mpcunit] [-s nSmooth]   < snapshot.tipsy
 *
 * Extra flags (do not collide with the reference's): -gpu <ordinal>, -bench-json <file>.
 */
#include <assert.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "kd.h"

/* Virial overdensity relative to the mean, Kitayama & Suto 1996 (so.c:57-86). */
static double omega_at(double Omega0, double Lambda0, double z)
{
    double a2 = (1.0 + z) * (1.0 + z), a3 = a2 * (1.0 + z);
    return Omega0 * a3 / (Omega0 * a3 + (1. - Omega0 - Lambda0) * a2 + Lambda0);
}

static double rhovir_over_rhobar(double Omega0, int bLambda, double z)
{
    double answer;
    if (Omega0 == 1.0) return 178.0;
    if (bLambda) {
        double wf = 1. / omega_at(Omega0, 1.0 - Omega0, z) - 1.;
        answer = 18. * (M_PI * M_PI) * (1. + 0.4093 * pow(wf, 0.9052));
    } else {
        double etaf = acosh(2.0 / omega_at(Omega0, 0.0, z) - 1.0);
        answer = 4.0 * (M_PI * M_PI) / (pow(sinh(etaf) - etaf, 2));
        answer *= pow(cosh(etaf) - 1, 3);
    }
    return answer;
}

static void usage(void)
{
    fputs("USAGE:\n"
          "so -i <SKID .gtp file> [-o <outfilebase>] [([-dark] [-gas] [-star]) || [-all])]\n"
          "      [-mark <markfile>]  [-std]  [-grp] [-gtp] [-subsumed] [-ignored]\n"
          "      [-list <File containing group indexes>]\n"
          "      [-pot || -stat <SKID .stat file containing most-bound-particle positions>]\n"
          "      [-delta <fThreshold>] [-M <fMinGTPMass>] [-m <mMinSOMembers>]\n"
          "      [-O <fOmega0>]  [-L]  [-z <fRedshift>]\n"
          "      [-p <xyzPeriod>]  [-c <xyzCenter>]\n"
          "      [-cx <xCenter>]  [-cy <yCenter>]  [-cz <zCenter>]\n"
          "      [-u <fMassUnit> <fMpcUnit>]   [-gpu <device>] [-gpus <N devices>] [-bench-json <file>]\n\n"
          "  B200 build of the spherical-overdensity finder: for every group of the .gtp catalog finds the\n"
          "  smallest radius at which the mean enclosed density drops below <fThreshold> (x Omega0), and\n"
          "  writes <outfilebase>.sovcirc (+ .sogrp/.sogtp/.sosub/.soign/.sodark... on request).\n"
          "  The TIPSY snapshot is read from stdin.  Periodic boundaries are assumed (default period 1).\n"
          "  Groupwise error codes in the Mvir and Rvir columns:\n"
          "     -1.0  fewer than nMembers particles within 1.2 times the group's .gtp radius\n"
          "     -2.0  density below threshold already at nMembers particles\n"
          "     -3.0  density never drops below threshold\n"
          "     -Mvir group subsumed or slurped by group -Rvir/10\n", stderr);
    exit(1);
}

static char *next_arg(int *i, int argc, char **argv)
{
    if (++(*i) >= argc) usage();
    return argv[*i];
}

int main(int argc, char **argv)
{
    KD kd;
    int i, j, sec, usec;
    int bThreshold = 0, bStandard = 0, bLambda = 0, bPeriodic = 1, bRedshift = 0;
    int bDark = 0, bGas = 0, bStar = 0, bMark = 0, bGrp = 0, bGtp = 0, bPot = 0, bSubsumed = 0, bIgnored = 0;
    int nBucket = 16, nMembers = 8, nSmooth = 1028, iDevice = -1, nGpus = 1;
    float fOmega = 1.0f, fLambda = 0.0f, fRedshift = -9.9999f, fThreshold = 0.0f, fMinMass = 0.0f;
    float fPeriod[3] = {1.0f, 1.0f, 1.0f}, fCenter[3] = {0.0f, 0.0f, 0.0f};
    float fMassUnit = -9.9f, fMpcUnit = -9.9f, G = 1.0f, H0 = 2.8944f;
    char *achGTPFile = NULL, *achListFile = NULL, *achOutFileBase = NULL, *achMarkFile = NULL, *achStatFile = NULL;
    char *achBenchJson = NULL;
    char achDefOutBase[] = "so", achLongWord[256];
    time_t timeRun;
    double tPhase;
    FILE *fpOutFile;

    fprintf(stderr, "SO Release 1.7: Jeff Gardner, May 2003\n");
    for (i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (!strcmp(a, "-i")) achGTPFile = next_arg(&i, argc, argv);
        else if (!strcmp(a, "-o")) achOutFileBase = next_arg(&i, argc, argv);
        else if (!strcmp(a, "-z")) { bRedshift = 1; fRedshift = atof(next_arg(&i, argc, argv)); }
        else if (!strcmp(a, "-O")) fOmega = atof(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-L")) bLambda = 1;
        else if (!strcmp(a, "-s")) nSmooth = atoi(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-rho")) {
            fprintf(stderr, "-rho option is no longer availible.  Use -delta instead.\n");
            usage();
        }
        else if (!strcmp(a, "-delta")) { fThreshold = atof(next_arg(&i, argc, argv)); bThreshold = 1; }
        else if (!strcmp(a, "-m")) nMembers = atoi(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-p")) { fPeriod[0] = fPeriod[1] = fPeriod[2] = atof(next_arg(&i, argc, argv)); bPeriodic = 1; }
        else if (!strcmp(a, "-c")) fCenter[0] = fCenter[1] = fCenter[2] = atof(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-cx")) fCenter[0] = atof(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-cy")) fCenter[1] = atof(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-cz")) fCenter[2] = atof(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-std")) bStandard = 1;
        else if (!strcmp(a, "-M")) fMinMass = atof(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-u")) { fMassUnit = atof(next_arg(&i, argc, argv)); fMpcUnit = atof(next_arg(&i, argc, argv)); }
        else if (!strcmp(a, "-list")) achListFile = next_arg(&i, argc, argv);
        else if (!strcmp(a, "-grp")) bGrp = 1;
        else if (!strcmp(a, "-gtp")) bGtp = 1;
        else if (!strcmp(a, "-pot")) { bPot = 1; if (achStatFile != NULL) usage(); }
        else if (!strcmp(a, "-subsumed")) bSubsumed = 1;
        else if (!strcmp(a, "-ignored")) bIgnored = 1;
        else if (!strcmp(a, "-stat")) { achStatFile = next_arg(&i, argc, argv); if (bPot) usage(); }
        else if (!strcmp(a, "-mark")) { achMarkFile = next_arg(&i, argc, argv); bMark = 1; }
        else if (!strcmp(a, "-dark")) bDark = 1;
        else if (!strcmp(a, "-gas")) bGas = 1;
        else if (!strcmp(a, "-star")) bStar = 1;
        else if (!strcmp(a, "-all")) bDark = bGas = bStar = 1;
        else if (!strcmp(a, "-gpu")) iDevice = atoi(next_arg(&i, argc, argv));
        else if (!strcmp(a, "-gpus")) { nGpus = atoi(next_arg(&i, argc, argv)); if (nGpus < 1 || nGpus > 16) usage(); }
        else if (!strcmp(a, "-bench-json")) achBenchJson = next_arg(&i, argc, argv);
        else usage();
    }
    if (achGTPFile == NULL) usage();
    if (achOutFileBase == NULL) achOutFileBase = achDefOutBase;
    if (bLambda) fLambda = 1.0 - fOmega;

    kdInit(&kd, nBucket, fPeriod, fCenter, 0, nMembers, bPeriodic, bDark, bGas, bStar, bMark, bPot);
    kd->iDevice = iDevice;
    kd->nGpus = nGpus;
    kd->bSkipGrpArray = 0;                            /* PINIT.iGrp is read by kdWriteArray AND by kdOutStats (kd2.c:1360), which always runs */
    kd->bSkipVcm = !bGtp;                             /* GRPNODE.vcm is only read by kdWriteGTP */
    kdGpu(kd);                                        /* the snapshot is streamed to the device as it is read */
    kdPhase(NULL, &tPhase);
    i = kdReadTipsy(kd, stdin, bStandard);
    kdPhase("read TIPSY snapshot -> device", &tPhase);
    fprintf(stderr, "Read %d particles from TIPSY file.\n", i);
    if (bMark) {
        i = kdReadMark(kd, achMarkFile);
        fprintf(stderr, "%d mark particles read from %s\n", i, achMarkFile);
    }
    if (!bRedshift) fRedshift = (1.0 / kd->fTime) - 1.0;                    /* so.c:470-472 */
    if (!bThreshold) fThreshold = rhovir_over_rhobar(fOmega, bLambda, fRedshift) * fOmega;
    else fThreshold *= fOmega;                                              /* so.c:477-481 */

    snprintf(achLongWord, sizeof(achLongWord), "%s.sovcirc", achOutFileBase);
    fpOutFile = fopen(achLongWord, "w");
    assert(fpOutFile != NULL);
    time(&timeRun);
    fprintf(fpOutFile, "#SO v1.61: Jeff Gardner, April 2002\n");            /* so.c:491-511 */
    fprintf(fpOutFile, "# Run on %s", ctime(&timeRun));
    fprintf(fpOutFile, "# Input .gtp file: %s\n", achGTPFile);
    if (achListFile != NULL) fprintf(fpOutFile, "# Groups list from file: %s\n", achListFile);
    if (achStatFile != NULL) fprintf(fpOutFile, "# Group potential centers from file: %s\n", achStatFile);
    if (bThreshold) fprintf(fpOutFile, "# fThreshold = %g  (set by user)\n", fThreshold);
    else fprintf(fpOutFile, "# fThreshold = %g  (VIRIAL DENSITY)\n", fThreshold);
    fprintf(fpOutFile, "# fRedshift: %g   fOmega: %g   fLambda: %g\n", fRedshift, fOmega, fLambda);
    fprintf(fpOutFile, "# bPeriodic: %d  fPeriod[i]: %g %g %g   fCenter[i]: %g %g %g\n", bPeriodic, fPeriod[0],
            fPeriod[1], fPeriod[2], fCenter[0], fCenter[1], fCenter[2]);
    fprintf(fpOutFile, "# fMinMass: %g  nMembers: %d  bPot: %d\n", fMinMass, nMembers, bPot);
    if (fMassUnit < 0.0) fprintf(fpOutFile, "# fMassUnit: UNSPECIFIED  fMpcUnit: UNSPECIFIED\n#\n");
    else fprintf(fpOutFile, "# fMassUnit: %g  fMpcUnit: %g\n#\n", fMassUnit, fMpcUnit);

    kdSetUniverse(kd, G, fOmega, fLambda, H0, fRedshift, fMassUnit, fMpcUnit);
    kdBuildTree(kd);

    i = kdReadGTPList(kd, achGTPFile, achListFile, fMinMass, bStandard);
    fprintf(stderr, "Read %d groups to process.\n", i);
    if (achStatFile != NULL) {
        j = kdReadStat(kd, achStatFile);
        fprintf(stderr, "Replaced %d group centers.\n", j);
        if (i != j) {
            fprintf(stderr, "ERROR in reading .stat file!\n");
            exit(1);
        }
    }

    kdTime(kd, &sec, &usec);
    kdSO(kd, fThreshold, nSmooth);
    kdTime(kd, &sec, &usec);
    kdPhase(NULL, &tPhase);

    kdOutStats(kd, fpOutFile);
    if (bDark) kdWriteProfile(kd, achOutFileBase, timeRun, fpOutFile, DARK);
    if (bGas) kdWriteProfile(kd, achOutFileBase, timeRun, fpOutFile, GAS);
    if (bStar) kdWriteProfile(kd, achOutFileBase, timeRun, fpOutFile, STAR);
    if (bMark) kdWriteProfile(kd, achOutFileBase, timeRun, fpOutFile, MARK);
    kdWriteOut(kd, fpOutFile);
    fclose(fpOutFile);
    if (bGrp) kdWriteArray(kd, achOutFileBase);
    if (bGtp) kdWriteGTP(kd, achOutFileBase, bStandard);
    if (bSubsumed) kdWriteConflict(kd, achOutFileBase, KD_SUBSUMED);
    if (bIgnored) kdWriteConflict(kd, achOutFileBase, KD_IGNORED);
    kdPhase("write output files", &tPhase);

    fprintf(stderr, "SO CPU Time:");
    fprintf(stderr, "   %d.%06d\n\n", sec, usec);
    if (achBenchJson) {
        FILE *fj = fopen(achBenchJson, "w");
        if (fj) {
            fprintf(fj, "{\"n_particles\": %d, \"n_halos\": %d, \"build_seconds\": %.6f, \"so_seconds\": %.6f, "
                        "\"r2_evaluations\": %lld, \"members\": %lld}\n",
                    kd->nParticles, kd->nGrps, kd->dBuildSeconds, kd->dSOSeconds, kd->nEvals, kd->nMembersTotal);
            fclose(fj);
        }
    }
    kdFinish(kd);
    return 0;
}
