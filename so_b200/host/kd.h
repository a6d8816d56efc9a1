/* kd.h — host side of the drop-in `so` program (C).
 *
 * Same public function names, argument meaning and file formats as the reference's "kd" API
 * (/root/reference/kd2.h:258-276), so so_main.c reads like /root/reference/so.c:192-575.  The two
 * hot-path calls, kdBuildTree (kd2.c:1096-1185) and kdSO (kd2.c:864-895), are routed to the CUDA
 * library through the C-ABI in include/sogpu.h; everything around them (tipsy / gtp / stat / mark
 * readers, conflict bookkeeping, Vcirc post-processing, writers) is plain host C, rewritten here
 * without libtirpc.  There is no CPU implementation of the search: without a B200 the program
 * prints the CUDA error and exits 1.
 */
#ifndef SO_HOST_KD_H
#define SO_HOST_KD_H

#include <stdint.h>
#include <stdio.h>
#include <time.h>

#include "../../include/sogpu.h"

#define NVCIRC 8          /* kd2.h:9  */
#define NMASSPROFILE 16   /* kd2.h:11 */

#define DARK 1
#define GAS 2
#define STAR 4
#define MARK 8

#define KD_SUBSUMED 0
#define KD_IGNORED 1

/* what kdReadTipsy keeps per particle (PINIT, kd2.h:41-53), stored as separate arrays in file order */
typedef struct {
    float *r;        /* 3 per particle */
    float *v;        /* 3 per particle */
    float *fMass;
    float *fPhi;
    int32_t *iGrp;
    int32_t *nSubsumed;
    int32_t *nIgnored;
} PARTICLES;

/* GRPNODE, kd2.h:86-102 (zero-initialised here; the reference leaves it uninitialised) */
typedef struct grpNode {
    int index;
    float pos[3];
    float vcm[3];
    float fRgtp;
    float fGTPMass;
    float fMvir;
    float fRvir;
    float fVcirc[NVCIRC];
    float fRmass[2];
    float fRmax;
    float fVmax;
    float fDark[NMASSPROFILE];
    float fGas[NMASSPROFILE];
    float fStar[NMASSPROFILE];
    float fMark[NMASSPROFILE];
} GRPNODE;

typedef struct kdContext {
    int nBucket;
    int bPeriodic;
    float fPeriod[3];
    float fCenter[3];
    float G;
    float z;
    float fMassUnit, fMpcUnit;
    int nParticles, nDark, nGas, nStar;
    float fTime;
    PARTICLES p;
    GRPNODE *grps;
    int nGrps;
    int nMembers;
    int bDark, bGas, bStar, bMark, bPot;
    char *bMarkList;
    int nInGTP;
    int iGroupsRemoved, iGroupsSlurped;
    int uSecond, uMicro;
    sogpu_t *gpu;
    int iDevice;
    int nGpus;             /* -gpus N: kdBuildTree + kdRvir over N devices of this process (kd_multi.c) */
    int bIngested;         /* kdReadTipsy streamed the particles to the device already */
    int bSkipGrpArray;     /* main() sets these when nothing will print the per-particle tags (.sogrp) ... */
    int bSkipVcm;          /* ... or the centre-of-mass velocities (.sogtp)                                  */
    int nGrpsInConflict;   /* groups that shared particles and went through the sequential replay           */
    /* filled by kdSO for reporting (--bench-json) */
    double dBuildSeconds, dSOSeconds;
    long long nEvals, nMembersTotal;
} *KD;

void kdTime(KD, int *, int *);
int kdInit(KD *, int nBucket, float *fPeriod, float *fCenter, int bOutDiag, int nMembers, int bPeriodic,
           int bDark, int bGas, int bStar, int bMark, int bPot);
void kdSetUniverse(KD, float G, float Omega0, float Lambda, float H0, float z, float fMassUnit, float fMpcUnit);
int kdParticleType(KD, int iOrder);
int kdReadMark(KD, char *);
int kdReadStat(KD, char *);
int kdReadGTPList(KD, char *achGTPFile, char *achListFile, float fMinMass, int bStandard);
int kdReadTipsy(KD, FILE *, int bStandard);
void kdSO(KD, float rhovir, int nSmooth);
void kdWriteProfile(KD, char *achOutFileBase, time_t, FILE *, int ptype);
void kdWriteOut(KD, FILE *);
int kdBuildTree(KD);
int kdRvirSeveralDevices(KD kd, const float *centers, const float *rgtp, int h, float thr, float *rvir, float *mvir,
                         int32_t *ndelta, int64_t *off, int32_t **mem, float **d2);
sogpu_t *kdGpu(KD);                              /* the device handle, created on first use */
void kdPhase(const char *name, double *t);       /* SO_TIMING=1: phase wall-clock on stderr */
void kdFinish(KD);
void kdWriteConflict(KD, char *achOutFileBase, int iOpt);
void kdOutStats(KD, FILE *);
void kdWriteArray(KD, char *achOutFileBase);
void kdWriteGTP(KD, char *achOutFileBase, int bStandard);

void indexx(int n, float arr[], int indx[]);   /* 1-based, as in nr.c:91 */

#endif
