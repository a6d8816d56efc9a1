/* kd_io.c — file formats of the drop-in `so`: tipsy snapshot, .gtp catalog, .stat, mark file in;
 * .sovcirc rows, .sogrp, .sogtp, .sosub/.soign, mass-profile files out.
 *
 * Formats follow the reference byte for byte (they are what parity is diffed on):
 *   header `struct dump`           tipsydefs.h:41-48  (native: 32 bytes incl. 4 pad)
 *   XDR header (-std)              kd2.c:32-44        (big-endian f8 + 6 x i4)
 *   gas / dark / star records      tipsydefs.h:6-37   (12 / 9 / 11 floats)
 *   readers                        kd2.c:144-421
 *   writers                        kd2.c:901-1008, 1216-1415
 * The reference goes through Sun-RPC XDR (libtirpc); here -std is a plain big-endian byte swap.
 */
#include "kd.h"

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define GRAV 6.6726e-8 /* G in cgs, kd2.c:899 */

static void check_file(FILE *fp, const char *name)
{
    if (fp == NULL) {                                   /* kd2.c:24-30 */
        fprintf(stderr, "ERROR opening file %s\n", name);
        exit(1);
    }
}

/* ---- tipsy header / records ------------------------------------------------------------------ */

typedef struct {
    double time;
    int nbodies, ndim, nsph, ndark, nstar;
} tipsy_header;

static uint32_t bswap32(uint32_t v)
{
    return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24);
}

static void swap_floats(float *f, size_t n)
{
    uint32_t *u = (uint32_t *)f;
    size_t i;
    for (i = 0; i < n; ++i) u[i] = bswap32(u[i]);
}

static int read_header(FILE *fp, int bStandard, tipsy_header *h)
{
    unsigned char b[32];
    if (fread(b, 1, 32, fp) != 32) return 0;
    if (bStandard) {
        uint32_t w[8];
        uint64_t q;
        int i;
        memcpy(w, b, 32);
        for (i = 0; i < 8; ++i) w[i] = bswap32(w[i]);
        q = ((uint64_t)w[0] << 32) | w[1];
        memcpy(&h->time, &q, 8);
        h->nbodies = (int)w[2]; h->ndim = (int)w[3]; h->nsph = (int)w[4];
        h->ndark = (int)w[5]; h->nstar = (int)w[6];
    } else {
        memcpy(&h->time, b, 8);
        memcpy(&h->nbodies, b + 8, 4); memcpy(&h->ndim, b + 12, 4); memcpy(&h->nsph, b + 16, 4);
        memcpy(&h->ndark, b + 20, 4); memcpy(&h->nstar, b + 24, 4);
    }
    return 1;
}

static void write_header(FILE *fp, int bStandard, const tipsy_header *h)
{
    unsigned char b[32];
    memset(b, 0, 32);
    if (bStandard) {
        uint32_t w[8];
        uint64_t q;
        int i;
        memcpy(&q, &h->time, 8);
        w[0] = (uint32_t)(q >> 32); w[1] = (uint32_t)q;
        w[2] = (uint32_t)h->nbodies; w[3] = (uint32_t)h->ndim; w[4] = (uint32_t)h->nsph;
        w[5] = (uint32_t)h->ndark; w[6] = (uint32_t)h->nstar; w[7] = 0;
        for (i = 0; i < 8; ++i) w[i] = bswap32(w[i]);
        memcpy(b, w, 32);
    } else {
        memcpy(b, &h->time, 8);
        memcpy(b + 8, &h->nbodies, 4); memcpy(b + 12, &h->ndim, 4); memcpy(b + 16, &h->nsph, 4);
        memcpy(b + 20, &h->ndark, 4); memcpy(b + 24, &h->nstar, 4);
    }
    fwrite(b, 1, 32, fp);
}

/* Read `count` records of `nf` floats (field offsets in floats: mass 0, pos 1..3, vel 4..6, phi last) and
 * stream them RAW to the device: chunks are read into two page-locked buffers used alternately, each is
 * handed to sogpu_ingest_records (asynchronous DMA + unpack / XDR byte swap on the GPU, f3 of SURVEY §8)
 * while the next one is being read.  The host keeps only what its own passes need: the mass, the
 * velocity if .sogtp is written, the potential and the position for -pot. */
#define READ_CHUNK (1 << 20)

static float load_f32(const float *p, int swap)
{
    uint32_t u;
    float f;
    memcpy(&u, p, 4);
    if (swap) u = bswap32(u);
    memcpy(&f, &u, 4);
    return f;
}

static double g_t_fread, g_t_ingest, g_t_extract;
static double now_s(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void read_species(KD kd, FILE *fp, int bStandard, int first, int count, int nf, float *buf[2], int *slot)
{
    int done = 0;
    while (done < count) {
        int k = count - done < READ_CHUNK ? count - done : READ_CHUNK, i;
        float *b = buf[*slot];
        double t0 = now_s(), t1, t2;
        size_t got = fread(b, sizeof(float) * nf, (size_t)k, fp);
        t1 = now_s();
        if ((int)got != k) {
            fprintf(stderr, "ERROR: TIPSY file ends after %d of %d particles\n", first + done + (int)got, kd->nParticles);
            exit(1);
        }
        if (sogpu_ingest_records(kd->gpu, b, k, nf, bStandard)) {
            fprintf(stderr, "ERROR in kdReadTipsy (sogpu_ingest_records): %s\n", sogpu_last_error());
            exit(1);
        }
        t2 = now_s();
        {   /* host copies of the fields the host passes still read (branches hoisted out of the loops) */
            const int p0 = first + done;
            float *m = kd->p.fMass + p0;
            if (!bStandard) {
                for (i = 0; i < k; ++i) m[i] = b[(size_t)i * nf];
                if (kd->p.v) {
                    float *v = kd->p.v + 3 * (size_t)p0;
                    for (i = 0; i < k; ++i) memcpy(v + 3 * (size_t)i, b + (size_t)i * nf + 4, 3 * sizeof(float));
                }
                if (kd->p.r) {
                    float *r = kd->p.r + 3 * (size_t)p0, *phi = kd->p.fPhi + p0;
                    for (i = 0; i < k; ++i) {
                        memcpy(r + 3 * (size_t)i, b + (size_t)i * nf + 1, 3 * sizeof(float));
                        phi[i] = b[(size_t)i * nf + nf - 1];
                    }
                }
            } else {
                for (i = 0; i < k; ++i) {
                    const float *rec = b + (size_t)i * nf;
                    int p = p0 + i, j;
                    m[i] = load_f32(rec, 1);
                    if (kd->p.v) for (j = 0; j < 3; ++j) kd->p.v[3 * (size_t)p + j] = load_f32(rec + 4 + j, 1);
                    if (kd->p.r) {
                        for (j = 0; j < 3; ++j) kd->p.r[3 * (size_t)p + j] = load_f32(rec + 1 + j, 1);
                        kd->p.fPhi[p] = load_f32(rec + nf - 1, 1);
                    }
                }
            }
        }
        g_t_fread += t1 - t0; g_t_ingest += t2 - t1; g_t_extract += now_s() - t2;
        done += k;
        *slot ^= 1;
    }
}

int kdReadTipsy(KD kd, FILE *fp, int bStandard)
{
    tipsy_header h;
    size_t n;
    float *buf[2];
    int slot = 0;
    double tp;
    if (!read_header(fp, bStandard, &h)) {
        fprintf(stderr, "ERROR: cannot read TIPSY header\n");
        exit(1);
    }
    kd->nDark = h.ndark; kd->nGas = h.nsph; kd->nStar = h.nstar;   /* kd2.c:339-347 */
    kd->fTime = (float)h.time;
    kd->nParticles = kd->nDark + kd->nGas + kd->nStar;
    n = (size_t)(kd->nParticles > 0 ? kd->nParticles : 1);
    kd->p.fMass = (float *)malloc(n * sizeof(float));
    kd->p.v = NULL;                                 /* velocities stay on the device (sogpu_vcm), see below */
    kd->p.r = kd->bPot ? (float *)malloc(n * 3 * sizeof(float)) : NULL;       /* only -pot reads these two  */
    kd->p.fPhi = kd->bPot ? (float *)malloc(n * sizeof(float)) : NULL;
    kd->p.iGrp = (int32_t *)calloc(n, sizeof(int32_t));
    kd->p.nSubsumed = (int32_t *)calloc(n, sizeof(int32_t));
    kd->p.nIgnored = (int32_t *)calloc(n, sizeof(int32_t));
    assert(kd->p.fMass && kd->p.iGrp && kd->p.nSubsumed && kd->p.nIgnored);
    assert(!kd->bPot || (kd->p.r && kd->p.fPhi));
    fprintf(stderr, "nDark:%d nGas:%d nStar:%d\n", kd->nDark, kd->nGas, kd->nStar);
    if (kd->nParticles == 0) return 0;
    kdGpu(kd);                                                      /* the records go straight to the device */
    kdPhase(NULL, &tp);
    if (sogpu_ingest_keep_velocities(kd->gpu, !kd->bSkipVcm) ||            /* _VcmParticles runs on the device */
        sogpu_ingest_begin(kd->gpu, kd->nParticles, kd->fPeriod, kd->fCenter)) {
        fprintf(stderr, "ERROR in kdReadTipsy (sogpu_ingest_begin): %s\n", sogpu_last_error());
        exit(1);
    }
    buf[0] = (float *)sogpu_host_alloc((size_t)READ_CHUNK * 12 * sizeof(float));
    buf[1] = (float *)sogpu_host_alloc((size_t)READ_CHUNK * 12 * sizeof(float));
    assert(buf[0] && buf[1]);
    kdPhase("  ingest_begin + pinned buffers", &tp);
    /* file order: gas, dark, star (kdParticleType, kd2.c:135-141) */
    read_species(kd, fp, bStandard, 0, kd->nGas, 12, buf, &slot);
    read_species(kd, fp, bStandard, kd->nGas, kd->nDark, 9, buf, &slot);
    read_species(kd, fp, bStandard, kd->nGas + kd->nDark, kd->nStar, 11, buf, &slot);
    if (sogpu_ingest_end(kd->gpu)) {
        fprintf(stderr, "ERROR in kdReadTipsy (sogpu_ingest_end): %s\n", sogpu_last_error());
        exit(1);
    }
    kdPhase("  read + stream chunks", &tp);
    sogpu_host_free(buf[0]);
    sogpu_host_free(buf[1]);
    kdPhase("  free pinned buffers", &tp);
    kd->bIngested = 1;
    if (getenv("SO_TIMING"))
        fprintf(stderr, "  [timing]   of which fread %.1f ms, sogpu_ingest_records %.1f ms, host field copies %.1f ms\n",
                1e3 * g_t_fread, 1e3 * g_t_ingest, 1e3 * g_t_extract);
    return kd->nParticles;
}

int kdParticleType(KD kd, int iOrder)
{
    if (iOrder < kd->nGas) return GAS;
    if (iOrder < kd->nGas + kd->nDark) return DARK;
    if (iOrder < kd->nParticles) return STAR;
    return 0;
}

int kdReadMark(KD kd, char *achMarkFile)
{
    int a, b, c, i, nmark = 0;
    FILE *in = fopen(achMarkFile, "r");
    check_file(in, achMarkFile);
    kd->bMarkList = (char *)calloc((size_t)kd->nParticles, 1);
    assert(kd->bMarkList != NULL);
    if (fscanf(in, "%d %d %d", &a, &b, &c) != 3) { /* header, kd2.c:158 */ }
    while (fscanf(in, "%d", &i) == 1) {
        --i;                                              /* mark files count from 1 */
        assert(i >= 0 && i < kd->nParticles);
        kd->bMarkList[i] = 1;
        ++nmark;
    }
    fclose(in);
    return nmark;
}

int kdReadGTPList(KD kd, char *achGTPFile, char *achListFile, float fMinMass, int bStandard)
{
    int *list = NULL, nList = 0, capList = 0, i, k = 0, id;
    tipsy_header h;
    float *star;
    FILE *fp;

    if (achListFile != NULL) {                            /* kd2.c:187-203 */
        fp = fopen(achListFile, "r");
        check_file(fp, achListFile);
        while (fscanf(fp, "%d", &id) == 1) {
            if (nList == capList) {
                capList = capList ? 2 * capList : 1028;
                list = (int *)realloc(list, (size_t)capList * sizeof(int));
                assert(list != NULL);
            }
            list[nList++] = id;
        }
        fclose(fp);
    }
    fp = fopen(achGTPFile, "rb");
    check_file(fp, achGTPFile);
    if (!read_header(fp, bStandard, &h)) {
        fprintf(stderr, "ERROR: cannot read GTP header\n");
        exit(1);
    }
    if (h.ndark > 0 || h.nsph > 0) {                      /* kd2.c:220-223 */
        fprintf(stderr, " FILE TYPE MISMATCH: GTP file contains non-star particles!\n");
        exit(1);
    }
    star = (float *)malloc((size_t)(h.nstar > 0 ? h.nstar : 1) * 11 * sizeof(float));
    assert(star != NULL);
    if ((int)fread(star, 11 * sizeof(float), (size_t)h.nstar, fp) != h.nstar) {
        fprintf(stderr, "ERROR: GTP file is truncated\n");
        exit(1);
    }
    fclose(fp);
    if (bStandard) swap_floats(star, (size_t)h.nstar * 11);

    {   /* star record: mass 0, pos 1..3, vel 4..6, metals 7, tform 8, eps 9, phi 10 */
        int nCand = nList ? nList : h.nstar;
        kd->grps = (GRPNODE *)calloc((size_t)(nCand > 0 ? nCand : 1), sizeof(GRPNODE));
        assert(kd->grps != NULL);
        for (i = 0; i < nCand; ++i) {
            int src = nList ? list[i] - 1 : i;
            const float *rec;
            if (src < 0 || src >= h.nstar) {
                fprintf(stderr, "ERROR: group index %d outside the GTP file (1..%d)\n", src + 1, h.nstar);
                exit(1);
            }
            rec = star + (size_t)src * 11;
            if (rec[0] >= fMinMass) {                     /* kd2.c:248,266 */
                GRPNODE *g = &kd->grps[k++];
                g->pos[0] = rec[1]; g->pos[1] = rec[2]; g->pos[2] = rec[3];
                g->fRgtp = rec[9];
                g->fGTPMass = rec[0];
                g->index = src + 1;
            }
        }
    }
    kd->nGrps = k;
    kd->nInGTP = h.nstar;
    free(star);
    free(list);
    return k;
}

int kdReadStat(KD kd, char *achStatFile)
{
    int i, k = 0, grpnum, itemp;
    float ftemp, r[3];
    FILE *fp = fopen(achStatFile, "r");
    check_file(fp, achStatFile);
    while (fscanf(fp, "%d %d", &grpnum, &itemp) == 2) {   /* kd2.c:298-312 */
        for (i = 0; i < 16; ++i)
            if (fscanf(fp, "%f", &ftemp) != 1) break;
        if (fscanf(fp, "%f %f %f", &r[0], &r[1], &r[2]) != 3) break;
        if (k < kd->nGrps && grpnum == kd->grps[k].index) {
            kd->grps[k].pos[0] = r[0]; kd->grps[k].pos[1] = r[1]; kd->grps[k].pos[2] = r[2];
            ++k;
        }
    }
    fclose(fp);
    return k;
}

/* ---- writers ------------------------------------------------------------------------------------ */

void kdWriteProfile(KD kd, char *achOutFileBase, time_t timeRun, FILE *fpOutFile, int ptype)
{
    char name[256], label[8];
    const char *ext;
    float massunit = kd->fMassUnit < 0.0f ? 1.0f : kd->fMassUnit;
    FILE *out;
    int i, j;
    switch (ptype) {                                      /* kd2.c:909-931 */
    case DARK: ext = "sodark"; strcpy(label, "dark"); break;
    case GAS: ext = "sogas"; strcpy(label, "gas"); break;
    case STAR: ext = "sostar"; strcpy(label, "star"); break;
    default: ext = "somark"; strcpy(label, "marked"); break;
    }
    snprintf(name, sizeof(name), "%s.%s", achOutFileBase, ext);
    out = fopen(name, "w");
    assert(out != NULL);
    fprintf(fpOutFile, "# Radial mass profile for %s particles written to %s\n", label, name);
    fprintf(out, "# Radial mass profile for %s particles\n", label);
    fprintf(out, "# Run on %s", ctime(&timeRun));
    fprintf(out, "# grp# Mass(R = %4.2f ... 2 Rvir)\n", 2.0 / NMASSPROFILE);
    for (i = 0; i < kd->nGrps; ++i) {
        const GRPNODE *g = &kd->grps[i];
        const float *prof = ptype == DARK ? g->fDark : ptype == GAS ? g->fGas : ptype == STAR ? g->fStar : g->fMark;
        fprintf(out, "%d ", g->index);
        for (j = 0; j < NMASSPROFILE; ++j) fprintf(out, "%g ", prof[j] * massunit);
        fprintf(out, "\n");
    }
    fclose(out);
}

void kdWriteOut(KD kd, FILE *fpOutFile)
{
    float kmsecunit = 1.0f, massunit = 1.0f, kpcunit = 1.0f;
    int i, j;
    if (!(kd->fMassUnit < 0.0f)) {                        /* kd2.c:981-991 */
        double d = GRAV * kd->fMassUnit * (1.0 + kd->z) / kd->fMpcUnit;
        d = 25388.8 * sqrt(d) / 100000.0;
        kmsecunit = (float)d;
        kpcunit = kd->fMpcUnit * 1000.0;
        massunit = kd->fMassUnit;
    }
    fprintf(fpOutFile, "#\n# grp# Mvir Rvir R(0.25Mvir) R(0.5Mvir)  R(Vc_max)  Vc_max  Vc(R = %4.2f ... 2 Rvir)\n",
            2.0 / NVCIRC);
    for (i = 0; i < kd->nGrps; ++i) {
        const GRPNODE *g = &kd->grps[i];
        if (g->fMvir < 0.0f) fprintf(fpOutFile, "%i %g %g ", g->index, g->fMvir, g->fRvir);
        else fprintf(fpOutFile, "%i %g %g ", g->index, g->fMvir * massunit, g->fRvir * kpcunit);
        fprintf(fpOutFile, "%g %g %g %g ", g->fRmass[0] * kpcunit, g->fRmass[1] * kpcunit, g->fRmax * kpcunit,
                g->fVmax * kmsecunit);
        for (j = 0; j < NVCIRC; ++j) fprintf(fpOutFile, "%g ", g->fVcirc[j] * kmsecunit);
        fprintf(fpOutFile, "\n");
    }
}

/* one integer per line (kd2.c:1256-1258, 1231-1239): formatted by hand into a 1 MB buffer — at 10^7..10^9
 * particles fprintf("%d\n") per line costs more than the whole GPU run */
static void write_int_lines(FILE *fp, int first, const int32_t *a, int64_t n)
{
    enum { CAP = 1 << 20 };
    char *buf = (char *)malloc(CAP + 16);
    size_t len = 0;
    int64_t i;
    assert(buf != NULL);
    for (i = -1; i < n; ++i) {
        int64_t v = i < 0 ? first : a[i];
        char tmp[16];
        int k = 0;
        uint64_t u = v < 0 ? (uint64_t)(-v) : (uint64_t)v;
        if (v < 0) buf[len++] = '-';
        do { tmp[k++] = (char)('0' + u % 10); u /= 10; } while (u);
        while (k) buf[len++] = tmp[--k];
        buf[len++] = '\n';
        if (len >= CAP) { fwrite(buf, 1, len, fp); len = 0; }
    }
    if (len) fwrite(buf, 1, len, fp);
    free(buf);
}

void kdWriteConflict(KD kd, char *achOutFileBase, int iOpt)
{
    char name[256];
    const int32_t *a = iOpt == KD_SUBSUMED ? kd->p.nSubsumed : kd->p.nIgnored;
    FILE *fp;
    snprintf(name, sizeof(name), "%s.%s", achOutFileBase, iOpt == KD_SUBSUMED ? "sosub" : "soign");
    fp = fopen(name, "w");
    assert(fp != NULL);
    write_int_lines(fp, kd->nParticles, a, kd->nParticles);   /* kd2.c:1231-1239; we keep file order */
    fclose(fp);
}

void kdWriteArray(KD kd, char *achOutFileBase)
{
    char name[256];
    FILE *fp;
    snprintf(name, sizeof(name), "%s.sogrp", achOutFileBase);
    fp = fopen(name, "w");
    assert(fp != NULL);
    write_int_lines(fp, kd->nParticles, kd->p.iGrp, kd->nParticles);   /* kd2.c:1256-1258 */
    fclose(fp);
}

void kdWriteGTP(KD kd, char *achOutFileBase, int bStandard)
{
    char name[256];
    tipsy_header h;
    FILE *fp;
    int i, k = 0;
    snprintf(name, sizeof(name), "%s.sogtp", achOutFileBase);
    fp = fopen(name, "wb");
    assert(fp != NULL);
    h.nbodies = h.nstar = kd->nInGTP;                     /* kd2.c:1284-1289 */
    h.ndark = h.nsph = 0;
    h.ndim = 3;
    h.time = kd->fTime;
    write_header(fp, bStandard, &h);
    for (i = 0; i < kd->nInGTP; ++i) {
        float sp[11];
        memset(sp, 0, sizeof(sp));
        if (k < kd->nGrps && kd->grps[k].index == i + 1) {       /* kd2.c:1300-1310 */
            const GRPNODE *g = &kd->grps[k++];
            sp[0] = g->fMvir > 0.0f ? g->fMvir : 0.0f;           /* mass 0 for error codes */
            sp[1] = g->pos[0]; sp[2] = g->pos[1]; sp[3] = g->pos[2];
            sp[4] = g->vcm[0]; sp[5] = g->vcm[1]; sp[6] = g->vcm[2];
            sp[9] = g->fRvir;                                    /* error codes stay in eps */
            sp[8] = (float)g->index;
        } else {
            sp[8] = (float)(i + 1);                              /* kd2.c:1312-1320 */
        }
        if (bStandard) swap_floats(sp, 11);
        fwrite(sp, sizeof(float), 11, fp);
    }
    fclose(fp);
}

void kdOutStats(KD kd, FILE *fpOutFile)
{
    int iCumSub = 0, iSub = 0, iCumIgn = 0, iIgn = 0, i, pass;
    double fCumMassSub = 0.0, fMassSub = 0.0, fCumMassIgn = 0.0, fMassIgn = 0.0;
    double fHaloMassSum = 0.0, fParticleMassSum = 0.0;
    for (i = 0; i < kd->nParticles; ++i) {                /* kd2.c:1346-1362 */
        if (kd->p.nSubsumed[i] > 0) {
            ++iSub;
            iCumSub += kd->p.nSubsumed[i];
            fMassSub += kd->p.fMass[i];
            fCumMassSub += kd->p.fMass[i] * kd->p.nSubsumed[i];
        }
        if (kd->p.nIgnored[i] > 0) {
            ++iIgn;
            iCumIgn += kd->p.nIgnored[i];
            fMassIgn += kd->p.fMass[i];
            fCumMassIgn += kd->p.fMass[i] * kd->p.nIgnored[i];
        }
        if (kd->p.iGrp[i] > 0) fParticleMassSum += kd->p.fMass[i];
    }
    for (i = 0; i < kd->nGrps; ++i)
        fHaloMassSum += kd->grps[i].fMvir > 0.0f ? kd->grps[i].fMvir : 0.0;
    for (pass = 0; pass < 2; ++pass) {                    /* kd2.c:1371-1413: stderr, then the file */
        FILE *o = pass ? fpOutFile : stderr;
        const char *c = pass ? "#" : "";
        fprintf(o, pass ? "#STATS:\n" : "\nSTATS:\n");
        fprintf(o, "%s PARTICLES:\n", c);
        fprintf(o, "%s  Particles subsumed into larger groups (cumulative):  %i\n", c, iCumSub);
        fprintf(o, "%s  Particles subsumed into larger groups at least once: %i\n", c, iSub);
        fprintf(o, "%s  Mass subsumed into larger groups (cumulative):       %g\n", c, fCumMassSub);
        fprintf(o, "%s  Mass subsumed into larger groups at least once:      %g\n", c, fMassSub);
        fprintf(o, "%s  Particles retained by small groups in the face of adversity (cumulative):  %i\n", c, iCumIgn);
        fprintf(o, "%s  Particles retained by small groups in the face of adversity at least once: %i\n", c, iIgn);
        fprintf(o, "%s  Mass retained by smaller groups in the face of adversity (cumulative):     %g\n", c, fCumMassIgn);
        fprintf(o, "%s  Mass retained by smaller groups in the face of adversity at least once:    %g\n", c, fMassIgn);
        fprintf(o, "%s GROUPS:\n", c);
        fprintf(o, "%s  Groups subsumed into larger groups (cumulative):  %i\n", c, kd->iGroupsRemoved);
        fprintf(o, "%s  Groups 'slurped' into larger groups (cumulative): %i\n", c, kd->iGroupsSlurped);
        fprintf(o, "%s  Total Mass of .sogrp particles in halos: %g\n", c, fParticleMassSum);
        if (pass) {
            fprintf(o, "#  Total Mass of Groups:                    %g\n", fHaloMassSum);
            fprintf(o, "#  Percentage difference:                   %g\n", fHaloMassSum / fParticleMassSum - 1.);
        } else {
            fprintf(o, "  Total Mass of groups:                    %g\n", fHaloMassSum);
            fprintf(o, "  Mass Deviation (particles/groups-1):     %g\n", fHaloMassSum / fParticleMassSum - 1.);
        }
    }
}
