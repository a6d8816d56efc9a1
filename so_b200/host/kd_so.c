/* kd_so.c — the hot-path calls of the drop-in `so`, routed to the GPU through include/sogpu.h.
 *
 *   kdBuildTree  (kd2.c:1096-1185)  -> sogpu_build_grid (the particles are already on the device:
 *                                      kdReadTipsy streams the raw records there, kd_io.c)
 *   kdSO         (kd2.c:864-895)    -> sogpu_so + sogpu_members (sorted on the device), then
 *       kdTagParticles (kd2.c:663-720)  sogpu_tag_members: groups that share no particle are tagged on the
 *                                      device; the others are replayed here in the reference's processing
 *                                      order (kdSortMass/indexx, kd2.c:843-861) with kdZeroGroup's
 *                                      bookkeeping (kd2.c:617-643): subsume / slurp / ignore are order
 *                                      dependent.  The reference's O(N) sweep per zeroed group and O(H)
 *                                      search per conflicting particle (kd2.c:636-660) are replaced by the
 *                                      member lists and an index -> slot map.
 *       _VcmParticles (kd2.c:595-609)   host, only when .sogtp is written
 *       kdVcirc / kdMassProfile (kd2.c:437-586)  sogpu_vcirc on the device for equal-mass single-species
 *                                      snapshots; host walk over GPU-gathered, device-sorted 2*Rvir lists
 *                                      otherwise (per-species fp32 sums need the particle at each rank)
 */
#include "kd.h"

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <sys/resource.h>

static void die_gpu(const char *what)
{
    fprintf(stderr, "ERROR in %s: %s\n", what, sogpu_last_error());
    exit(1);
}

void kdTime(KD kd, int *puSecond, int *puMicro)
{
    struct rusage ru;                                     /* kd2.c:46-59: user CPU time deltas */
    getrusage(RUSAGE_SELF, &ru);
    *puMicro = (int)ru.ru_utime.tv_usec - kd->uMicro;
    *puSecond = (int)ru.ru_utime.tv_sec - kd->uSecond;
    if (*puMicro < 0) {
        *puMicro += 1000000;
        *puSecond -= 1;
    }
    kd->uSecond = (int)ru.ru_utime.tv_sec;
    kd->uMicro = (int)ru.ru_utime.tv_usec;
}

int kdInit(KD *pkd, int nBucket, float *fPeriod, float *fCenter, int bOutDiag, int nMembers, int bPeriodic,
           int bDark, int bGas, int bStar, int bMark, int bPot)
{
    KD kd = (KD)calloc(1, sizeof(struct kdContext));
    int j;
    (void)bOutDiag;
    assert(kd != NULL);
    kd->nBucket = nBucket;
    for (j = 0; j < 3; ++j) {
        kd->fPeriod[j] = fPeriod[j];
        kd->fCenter[j] = fCenter[j];
    }
    kd->G = 1.0f;
    kd->nMembers = nMembers;
    kd->bPeriodic = bPeriodic;
    kd->bDark = bDark; kd->bGas = bGas; kd->bStar = bStar; kd->bMark = bMark; kd->bPot = bPot;
    kd->iDevice = -1;
    *pkd = kd;
    return 1;
}

void kdSetUniverse(KD kd, float G, float Omega0, float Lambda, float H0, float z, float fMassUnit, float fMpcUnit)
{
    (void)Omega0; (void)Lambda; (void)H0;                 /* stored in a CSM the reference never reads again */
    kd->G = G;
    kd->z = z;
    kd->fMassUnit = fMassUnit;
    kd->fMpcUnit = fMpcUnit;
}

/* SO_TIMING=1: wall-clock of every phase on stderr */
void kdPhase(const char *name, double *t);
#define phase kdPhase
void kdPhase(const char *name, double *t)
{
    static int on = -1;
    struct timespec ts;
    double now;
    if (on < 0) on = getenv("SO_TIMING") != NULL;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    now = ts.tv_sec + 1e-9 * ts.tv_nsec;
    if (on && name) fprintf(stderr, "  [timing] %-28s %9.3f ms\n", name, 1e3 * (now - *t));
    *t = now;
}

static double wall(void)
{
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* the device handle, created on first use (CUDA context creation: 0.6 s and more, once per process) */
sogpu_t *kdGpu(KD kd)
{
    if (!kd->gpu) {
        double tp;
        kdPhase(NULL, &tp);
        if (sogpu_create(&kd->gpu, kd->iDevice)) die_gpu("sogpu_create");
        kdPhase("CUDA context", &tp);
    }
    return kd->gpu;
}

int kdBuildTree(KD kd)
{
    double t0 = wall();
    double tp;
    if (kd->nParticles == 0) return 1;
    phase(NULL, &tp);
    if (!kd->bIngested) {
        /* particles that did not come through kdReadTipsy's streaming ingest */
        kdGpu(kd);
        phase(NULL, &tp);
        if (!kd->p.r) { fprintf(stderr, "ERROR in kdBuildTree: no particle positions\n"); exit(1); }
        if (sogpu_set_particles_host(kd->gpu, kd->p.r, 3 * sizeof(float), kd->p.fMass, sizeof(float), kd->nParticles,
                                     kd->fPeriod, kd->fCenter))
            die_gpu("kdBuildTree (sogpu_set_particles_host)");
        phase("upload particles", &tp);
    }
    if (sogpu_build_grid(kd->gpu)) die_gpu("kdBuildTree (sogpu_build_grid)");
    phase("build grid (enqueue)", &tp);
    kd->dBuildSeconds = wall() - t0;
    return 1;
}

/* ---- indexx: Numerical-Recipes style index quicksort, ascending, 1-based (nr.c:91-151).  The
 * permutation it yields for EQUAL keys is algorithm specific and defines the order in which such
 * halos are processed, so the same algorithm (median of three, insertion sort below 7) is used. */
void indexx(int n, float arr[], int indx[])
{
    enum { SMALL = 7, DEPTH = 64 };
    int stack[DEPTH + 2], top = 0, lo = 1, hi = n, i, j, k, t, pivot_i;
    float pivot;
#define EXCH(a, b) do { t = (a); (a) = (b); (b) = t; } while (0)
    for (j = 1; j <= n; ++j) indx[j] = j;
    for (;;) {
        if (hi - lo < SMALL) {
            for (j = lo + 1; j <= hi; ++j) {
                pivot_i = indx[j];
                pivot = arr[pivot_i];
                for (i = j - 1; i >= 1; --i) {
                    if (arr[indx[i]] <= pivot) break;
                    indx[i + 1] = indx[i];
                }
                indx[i + 1] = pivot_i;
            }
            if (top == 0) break;
            hi = stack[top--];
            lo = stack[top--];
        } else {
            k = (lo + hi) >> 1;
            EXCH(indx[k], indx[lo + 1]);
            if (arr[indx[lo + 1]] > arr[indx[hi]]) EXCH(indx[lo + 1], indx[hi]);
            if (arr[indx[lo]] > arr[indx[hi]]) EXCH(indx[lo], indx[hi]);
            if (arr[indx[lo + 1]] > arr[indx[lo]]) EXCH(indx[lo + 1], indx[lo]);
            i = lo + 1;
            j = hi;
            pivot_i = indx[lo];
            pivot = arr[pivot_i];
            for (;;) {
                do ++i; while (arr[indx[i]] < pivot);
                do --j; while (arr[indx[j]] > pivot);
                if (j < i) break;
                EXCH(indx[i], indx[j]);
            }
            indx[lo] = indx[j];
            indx[j] = pivot_i;
            top += 2;
            if (top > DEPTH) { fprintf(stderr, "indexx: stack too small\n"); exit(1); }
            if (hi - i + 1 >= j - lo) { stack[top] = hi; stack[top - 1] = i; hi = j - 1; }
            else { stack[top] = j - 1; stack[top - 1] = lo; lo = i; }
        }
    }
#undef EXCH
}

/* ---- conflict bookkeeping -------------------------------------------------------------------- */

typedef struct {
    const int64_t *off;       /* member list of group slot g: mem[off[g] .. off[g+1]) */
    const int32_t *mem;
    int *slot_of_index;       /* catalog index (1-based) -> slot in kd->grps, -1 if absent */
} TAGCTX;

/* kdZeroGroup (kd2.c:617-643): every particle still tagged by `small` is untagged and counted */
static void zero_group(KD kd, const TAGCTX *c, int small, int big)
{
    GRPNODE *gs = &kd->grps[small];
    int64_t k;
    if (gs->fMvir < 0.0f) {
        fprintf(stderr, "\nERROR in kdZeroGroup:\nZeroed group mass is already negtive!\n");
        fprintf(stderr, "  OldGrp: %d  NewGrp: %d  fMvir: %g  Rvir: %g\n", gs->index, kd->grps[big].index,
                gs->fMvir, gs->fRvir);
        exit(1);
    }
    gs->fRvir = (float)(-10.0 * kd->grps[big].index);
    gs->fMvir = -gs->fMvir;
    for (k = c->off[small]; k < c->off[small + 1]; ++k) {
        int32_t p = c->mem[k];
        if (kd->p.iGrp[p] == gs->index) {
            kd->p.iGrp[p] = 0;
            ++kd->p.nSubsumed[p];
        }
    }
}

/* kdTagParticles (kd2.c:663-720) for group slot `big`, members in ascending (r^2, index) order */
static void tag_particles(KD kd, const TAGCTX *c, int big)
{
    GRPNODE *gb = &kd->grps[big];
    int64_t k;
    for (k = c->off[big]; k < c->off[big + 1]; ++k) {
        int32_t p = c->mem[k];
        int tag = kd->p.iGrp[p];
        if (tag == 0) {
            kd->p.iGrp[p] = gb->index;
        } else {
            int small = c->slot_of_index[tag];
            GRPNODE *gs = &kd->grps[small];
            float dx = gb->pos[0] - gs->pos[0], dy = gb->pos[1] - gs->pos[1], dz = gb->pos[2] - gs->pos[2];
            float r2 = dx * dx + dy * dy + dz * dz;       /* plain, non-periodic (kd2.c:677-680) */
            if (r2 <= gb->fRvir * gb->fRvir) {            /* B's centre inside A: A subsumes B */
                zero_group(kd, c, small, big);
                ++kd->iGroupsRemoved;
                kd->p.iGrp[p] = gb->index;
            } else if (r2 <= gs->fRvir * gs->fRvir) {     /* A's centre inside B: A is slurped */
                zero_group(kd, c, big, small);
                ++kd->iGroupsSlurped;
                return;                                   /* kd2.c:671: nothing after the slurp */
            } else {
                ++kd->p.nIgnored[p];
            }
        }
    }
}

/* _VcmParticles (kd2.c:595-609): fp32 sums in list order */
static void vcm_particles(KD kd, const TAGCTX *c, int g, float mass)
{
    float v[3] = {0.0f, 0.0f, 0.0f};
    int64_t k;
    int l;
    for (k = c->off[g]; k < c->off[g + 1]; ++k) {
        int32_t p = c->mem[k];
        for (l = 0; l < 3; ++l) v[l] += kd->p.fMass[p] * kd->p.v[3 * p + l];
    }
    for (l = 0; l < 3; ++l) kd->grps[g].vcm[l] = v[l] / mass;
}

/* ---- kdVcirc / kdMassProfile (kd2.c:437-586) over one r^2-sorted list -------------------------- */

static void mass_profile(KD kd, GRPNODE *g, float rvir, const int32_t *idx, const float *d2, int64_t n, int ptype)
{
    float *out = ptype == DARK ? g->fDark : ptype == GAS ? g->fGas : ptype == STAR ? g->fStar : g->fMark;
    float fmin = 2.0 / NMASSPROFILE, f, mass = 0.0f;
    int64_t j = 0;
    int i;
    for (f = fmin, i = 0; i < NMASSPROFILE - 1; ++i, f += fmin) {
        float r = f * rvir, r2 = r * r;
        while (j < n && d2[j] < r2) {
            int take = ptype == MARK ? kd->bMarkList[idx[j]] : (kdParticleType(kd, idx[j]) == ptype);
            if (take) mass += kd->p.fMass[idx[j]];
            ++j;
        }
        out[i] = mass;
    }
    for (; j < n; ++j) {
        int take = ptype == MARK ? kd->bMarkList[idx[j]] : (kdParticleType(kd, idx[j]) == ptype);
        if (take) mass += kd->p.fMass[idx[j]];
    }
    out[NMASSPROFILE - 1] = mass;
}

static void vcirc(KD kd, GRPNODE *g, float rvir, float mvir, const int32_t *idx, const float *d2, int64_t n)
{
    float fmin = 2.0 / NVCIRC, f, mass = 0.0f, fBall = 2. * rvir, m, r, vm, rm, vc;
    int64_t j = 0;
    int i;
    for (f = fmin, i = 0; i < NVCIRC - 1; ++i, f += fmin) {          /* kd2.c:517-526 */
        float r2;
        r = f * rvir;
        r2 = r * r;
        while (j < n && d2[j] < r2) mass += kd->p.fMass[idx[j++]];
        g->fVcirc[i] = sqrt(kd->G * mass / r);
    }
    for (; j < n; ++j) mass += kd->p.fMass[idx[j]];
    g->fVcirc[NVCIRC - 1] = sqrt(kd->G * mass / fBall);
    for (f = 0.25, i = 0; i < 2; ++i, f += 0.25) {                    /* kd2.c:537-546 */
        m = f * mvir;
        j = 0;
        mass = kd->p.fMass[idx[0]];
        while (mass < m && j + 1 < n) {
            ++j;
            mass += kd->p.fMass[idx[j]];
        }
        g->fRmass[i] = sqrt(d2[j]);
    }
    mass = 0.0f;                                                      /* kd2.c:551-569 */
    for (j = 0; j < kd->nMembers && j < n; ++j) mass += kd->p.fMass[idx[j]];
    rm = sqrt(d2[(kd->nMembers <= n ? kd->nMembers : n) - 1]);
    vm = sqrt(kd->G * mass / rm);
    for (j = kd->nMembers; j < n; ++j) {
        mass += kd->p.fMass[idx[j]];
        r = sqrt(d2[j]);
        vc = sqrt(kd->G * mass / r);
        if (vc > vm) {
            vm = vc;
            rm = r;
        }
    }
    g->fRmax = rm;
    g->fVmax = vm;
    if (kd->bDark) mass_profile(kd, g, rvir, idx, d2, n, DARK);
    if (kd->bGas) mass_profile(kd, g, rvir, idx, d2, n, GAS);
    if (kd->bStar) mass_profile(kd, g, rvir, idx, d2, n, STAR);
    if (kd->bMark) mass_profile(kd, g, rvir, idx, d2, n, MARK);
}

/* ---- kdSO ----------------------------------------------------------------------------------- */

void kdSO(KD kd, float rhovir, int nSmooth)
{
    const int h = kd->nGrps;
    float *centers, *rgtp, *rvir, *mvir, *masses;
    int32_t *ndelta, *mem;
    int64_t *off;
    int *order, *slot_of_index, *do_vcirc;
    unsigned char *in_conflict;
    const int32_t *lib_mem;
    const float *lib_d2;
    sogpu_stats_t st;
    TAGCTX c;
    double t0 = wall(), tp;
    int i, it, equal_mass = 0, device_replay = 0;
    float *rv2 = NULL, *mv2 = NULL;
    unsigned char *still_valid = NULL;
    (void)nSmooth;
    phase(NULL, &tp);                 /* sized the reference's neighbour list (smInit); nothing to size here */
    if (h == 0) return;
    if (kd->bPot) {
        /* -pot (kd2.c:749-761): re-centre every group on the particle of lowest potential inside
         * its .gtp radius.  One batched GPU gather of the fRgtp balls; the arg-min runs on the host
         * over lists in ascending (r^2, index) order (the reference scans kd-tree walk order; only
         * exactly equal potentials could be resolved differently). */
        float *pc = (float *)malloc((size_t)h * 3 * sizeof(float));
        float *pb = (float *)malloc((size_t)h * sizeof(float));
        int64_t *poff = (int64_t *)malloc(((size_t)h + 1) * sizeof(int64_t));
        const int32_t *pi;
        const float *pd;
        int64_t k;
        assert(pc && pb && poff);
        for (i = 0; i < h; ++i) {
            memcpy(pc + 3 * i, kd->grps[i].pos, 3 * sizeof(float));
            pb[i] = kd->grps[i].fRgtp * kd->grps[i].fRgtp;                 /* kd2.c:750 */
        }
        if (sogpu_ball_gather_batch(kd->gpu, pc, pb, h)) die_gpu("kdSO (-pot gather)");
        if (sogpu_members(kd->gpu, poff, &pi, &pd, 1)) die_gpu("kdSO (-pot lists)");
        for (i = 0; i < h; ++i) {
            if (poff[i + 1] > poff[i]) {
                int32_t best = pi[poff[i]];
                for (k = poff[i] + 1; k < poff[i + 1]; ++k)
                    if (kd->p.fPhi[pi[k]] < kd->p.fPhi[best]) best = pi[k];
                memcpy(kd->grps[i].pos, kd->p.r + 3 * (size_t)best, 3 * sizeof(float));   /* kd2.c:760 */
            }
        }
        free(pc); free(pb); free(poff);
    }
    centers = (float *)malloc((size_t)h * 3 * sizeof(float));
    rgtp = (float *)malloc((size_t)h * sizeof(float));
    rvir = (float *)malloc((size_t)h * sizeof(float));
    mvir = (float *)malloc((size_t)h * sizeof(float));
    masses = (float *)malloc(((size_t)h + 1) * sizeof(float));       /* 1-based for indexx */
    ndelta = (int32_t *)malloc((size_t)h * sizeof(int32_t));
    off = (int64_t *)malloc(((size_t)h + 1) * sizeof(int64_t));
    order = (int *)malloc(((size_t)h + 1) * sizeof(int));            /* 1-based for indexx */
    do_vcirc = (int *)calloc((size_t)h, sizeof(int));
    in_conflict = (unsigned char *)calloc((size_t)h, 1);
    slot_of_index = (int *)malloc(((size_t)kd->nInGTP + 2) * sizeof(int));
    assert(centers && rgtp && rvir && mvir && masses && ndelta && off && order && do_vcirc && slot_of_index && in_conflict);
    for (i = 0; i <= kd->nInGTP + 1; ++i) slot_of_index[i] = -1;
    for (i = 0; i < h; ++i) {
        memcpy(centers + 3 * i, kd->grps[i].pos, 3 * sizeof(float));   /* (after -pot re-centring) */
        rgtp[i] = kd->grps[i].fRgtp;
        masses[i + 1] = kd->grps[i].fGTPMass;
        slot_of_index[kd->grps[i].index] = i;
    }

    /* the hot path: R_Delta, M_Delta, N_Delta and the member lists of every group */
    if (sogpu_keep_member_d2(kd->gpu, 1)) die_gpu("kdSO");
    phase("kdSO setup", &tp);
    {
        int several = kd->nGpus > 1;
        if (several) {               /* the domain step needs one particle mass (it ships {x,y,z,index} records) */
            for (i = 1; i < kd->nParticles && several; ++i) several = kd->p.fMass[i] == kd->p.fMass[0];
            if (!several) fprintf(stderr, "so: -gpus %d ignored: particles of unequal mass run on one device\n", kd->nGpus);
        }
        if (several) {
            float *md2 = NULL;
            kdRvirSeveralDevices(kd, centers, rgtp, h, rhovir, rvir, mvir, ndelta, off, &mem, &md2);
            free(md2);
            phase("kdBuildTree + kdRvir over several devices", &tp);
        } else {
            if (sogpu_so(kd->gpu, centers, rgtp, h, rhovir, kd->nMembers, rvir, mvir, ndelta)) die_gpu("kdSO (sogpu_so)");
            phase("sogpu_so (build+query)", &tp);
            if (sogpu_members(kd->gpu, off, &lib_mem, &lib_d2, 1)) die_gpu("kdSO (sogpu_members)");
            phase("member lists (sorted)", &tp);
            mem = (int32_t *)malloc((size_t)(off[h] > 0 ? off[h] : 1) * sizeof(int32_t));
            assert(mem != NULL);
            memcpy(mem, lib_mem, (size_t)off[h] * sizeof(int32_t));
        }
    }
    if (sogpu_get_stats(kd->gpu, &st) == 0) {
        kd->nEvals = st.last_evals;
        kd->nMembersTotal = st.last_members;
        equal_mass = st.equal_mass;
    }

    /* kdTagParticles (kd2.c:663-720).  Groups that share no particle with another group are tagged on
     * the device, whatever the order; the others come back flagged and are replayed here, in
     * ascending catalog mass like the reference (kd2.c:873-879): subsume / slurp / ignore depend on
     * the order, and a conflict-free group can never be met by that replay. */
    {
        int32_t *ids = (int32_t *)malloc((size_t)h * sizeof(int32_t));
        assert(ids != NULL);
        for (i = 0; i < h; ++i) ids[i] = kd->grps[i].index;
        if (sogpu_tag_members(kd->gpu, ids, h, in_conflict, kd->bSkipGrpArray ? NULL : kd->p.iGrp))
            die_gpu("kdSO (sogpu_tag_members)");
        for (i = 0; i < h; ++i) kd->nGrpsInConflict += in_conflict[i];
        phase("tagging on the device", &tp);
        indexx(h, masses, order);
        /* the groups that share particles: kdTagParticles replayed in processing order ON THE DEVICE
         * (sogpu_tag_replay; SO_HOST_REPLAY=1 keeps the host replay below for comparison) */
        if (kd->nGrpsInConflict > 0 && !getenv("SO_HOST_REPLAY")) {
            int32_t *ord = (int32_t *)malloc((size_t)h * sizeof(int32_t)), no = 0, rem = 0, slu = 0;
            rv2 = (float *)malloc((size_t)h * sizeof(float));
            mv2 = (float *)malloc((size_t)h * sizeof(float));
            still_valid = (unsigned char *)malloc((size_t)h);
            assert(ord && rv2 && mv2 && still_valid);
            for (it = 1; it <= h; ++it) {
                int g = order[it] - 1;
                if (in_conflict[g] && rvir[g] > 0.0f) ord[no++] = g;
            }
            memcpy(rv2, rvir, (size_t)h * sizeof(float));
            memcpy(mv2, mvir, (size_t)h * sizeof(float));
            if (sogpu_tag_replay(kd->gpu, ord, no, ids, centers, rv2, mv2, h, kd->nInGTP + 1, kd->p.iGrp, kd->p.nSubsumed,
                                 kd->p.nIgnored, &rem, &slu, still_valid))
                die_gpu("kdSO (sogpu_tag_replay)");
            kd->iGroupsRemoved += rem;
            kd->iGroupsSlurped += slu;
            device_replay = 1;
            free(ord);
            phase("conflict replay on the device", &tp);
        }
        free(ids);
    }
    if (!kd->bSkipVcm && kd->p.v == NULL) {
        /* _VcmParticles (kd2.c:595-609, 826) of every resolved group on the device: the same sequential fp32
         * sum over the (r^2, index)-sorted members; the velocities were kept there by the ingest */
        float *vcm = (float *)malloc((size_t)h * 3 * sizeof(float));
        assert(vcm != NULL);
        if (sogpu_vcm(kd->gpu, mvir, h, vcm)) die_gpu("kdSO (sogpu_vcm)");
        for (i = 0; i < h; ++i)
            if (rvir[i] > 0.0f) memcpy(kd->grps[i].vcm, vcm + 3 * (size_t)i, 3 * sizeof(float));
        free(vcm);
        phase("vcm on the device", &tp);
    }
    c.off = off; c.mem = mem; c.slot_of_index = slot_of_index;
    for (it = 1; it <= h; ++it) {
        int g = order[it] - 1;
        GRPNODE *grp = &kd->grps[g];
        grp->fRvir = device_replay ? rv2[g] : rvir[g];     /* kd2.c:819-820 or the error code (with the subsume / slurp marks) */
        grp->fMvir = device_replay ? mv2[g] : mvir[g];
        if (rvir[g] > 0.0f) {
            if (in_conflict[g] && !device_replay) tag_particles(kd, &c, g);  /* kd2.c:823 */
            if (!kd->bSkipVcm && kd->p.v) vcm_particles(kd, &c, g, mvir[g]);   /* kd2.c:826; host copy of the velocities only */
            if (device_replay ? still_valid[g] : (grp->fRvir > 0.0f)) do_vcirc[g] = 1;        /* kd2.c:884: not slurped */
        }
    }
    free(rv2); free(mv2); free(still_valid);

    phase("conflict replay", &tp);
    /* kdVcirc for every group that was valid when the reference would have called it */
    {
        int nv = 0, k;
        int *slots = (int *)malloc((size_t)h * sizeof(int));
        float *vc = (float *)malloc((size_t)h * 3 * sizeof(float));
        float *vb = (float *)malloc((size_t)h * sizeof(float));
        int64_t *voff = (int64_t *)malloc(((size_t)h + 1) * sizeof(int64_t));
        assert(slots && vc && vb && voff);
        for (i = 0; i < h; ++i)
            if (do_vcirc[i]) {
                float fBall = 2. * rvir[i];                /* kd2.c:511-512 */
                memcpy(vc + 3 * nv, kd->grps[i].pos, 3 * sizeof(float));
                vb[nv] = fBall * fBall;
                slots[nv++] = i;
            }
        if (nv) {
            /* equal-mass, single-species snapshot (the usual dark-matter run): the whole of kdVcirc and
             * kdMassProfile is rank queries on the sorted lists and runs on the device (sogpu_vcirc);
             * otherwise the per-species fp32 sums need the particle at every rank: host walk below */
            const int one_species = kd->nDark == kd->nParticles || kd->nGas == kd->nParticles ||
                                    kd->nStar == kd->nParticles;
            const int want_prof = kd->bDark || kd->bGas || kd->bStar;
            if (equal_mass == 1 && !kd->bMark && (one_species || !want_prof) && !getenv("SO_HOST_VCIRC")) {
                float *vr = (float *)malloc((size_t)nv * 2 * sizeof(float)), *vm = vr + nv;
                float *o_vc = (float *)malloc((size_t)nv * (NVCIRC + 2 + 1 + 1 + NMASSPROFILE) * sizeof(float));
                float *o_rm = o_vc + (size_t)nv * NVCIRC, *o_rx = o_rm + (size_t)nv * 2, *o_vx = o_rx + nv;
                float *o_pr = o_vx + nv;
                assert(vr && o_vc);
                for (k = 0; k < nv; ++k) { vr[k] = rvir[slots[k]]; vm[k] = mvir[slots[k]]; }
                if (sogpu_vcirc(kd->gpu, vc, vr, vm, nv, kd->G, kd->nMembers, o_vc, o_rm, o_rx, o_vx,
                                want_prof ? o_pr : NULL))
                    die_gpu("kdSO (sogpu_vcirc)");
                phase("kdVcirc on the device", &tp);
                for (k = 0; k < nv; ++k) {
                    GRPNODE *g = &kd->grps[slots[k]];
                    memcpy(g->fVcirc, o_vc + (size_t)k * NVCIRC, NVCIRC * sizeof(float));
                    memcpy(g->fRmass, o_rm + (size_t)k * 2, 2 * sizeof(float));
                    g->fRmax = o_rx[k];
                    g->fVmax = o_vx[k];
                    /* a species that is absent keeps a zero profile (mass never incremented, kd2.c:470-480) */
                    if (kd->bDark && kd->nDark) memcpy(g->fDark, o_pr + (size_t)k * NMASSPROFILE, NMASSPROFILE * sizeof(float));
                    if (kd->bGas && kd->nGas) memcpy(g->fGas, o_pr + (size_t)k * NMASSPROFILE, NMASSPROFILE * sizeof(float));
                    if (kd->bStar && kd->nStar) memcpy(g->fStar, o_pr + (size_t)k * NMASSPROFILE, NMASSPROFILE * sizeof(float));
                }
                free(vr); free(o_vc);
            } else if (!getenv("SO_HOST_VCIRC")) {
                /* unequal masses, several species or -mark: the reference's sequential fp32 sums over the sorted
                 * 2 Rvir lists, per species, evaluated on the device (sogpu_vcirc_species) */
                int32_t masks[4], nm = 0, km;
                float *dst_of[4];
                unsigned char *pt = (unsigned char *)malloc((size_t)kd->nParticles);
                float *vr = (float *)malloc((size_t)nv * 2 * sizeof(float)), *vm = vr + nv;
                float *o_vc = (float *)malloc((size_t)nv * (NVCIRC + 2 + 1 + 1 + 4 * NMASSPROFILE) * sizeof(float));
                float *o_rm = o_vc + (size_t)nv * NVCIRC, *o_rx = o_rm + (size_t)nv * 2, *o_vx = o_rx + nv;
                float *o_pr = o_vx + nv;
                assert(pt && vr && o_vc);
                for (i = 0; i < kd->nParticles; ++i)
                    pt[i] = (unsigned char)(kdParticleType(kd, i) | ((kd->bMark && kd->bMarkList && kd->bMarkList[i]) ? MARK : 0));
                if (kd->bDark) masks[nm++] = DARK;
                if (kd->bGas) masks[nm++] = GAS;
                if (kd->bStar) masks[nm++] = STAR;
                if (kd->bMark) masks[nm++] = MARK;
                for (k = 0; k < nv; ++k) { vr[k] = rvir[slots[k]]; vm[k] = mvir[slots[k]]; }
                if (sogpu_vcirc_species(kd->gpu, vc, vr, vm, nv, kd->G, kd->nMembers, pt, masks, nm, o_vc, o_rm, o_rx, o_vx, o_pr))
                    die_gpu("kdSO (sogpu_vcirc_species)");
                phase("kdVcirc (species) on the device", &tp);
                for (k = 0; k < nv; ++k) {
                    GRPNODE *g = &kd->grps[slots[k]];
                    memcpy(g->fVcirc, o_vc + (size_t)k * NVCIRC, NVCIRC * sizeof(float));
                    memcpy(g->fRmass, o_rm + (size_t)k * 2, 2 * sizeof(float));
                    g->fRmax = o_rx[k];
                    g->fVmax = o_vx[k];
                    km = 0;
                    if (kd->bDark) dst_of[km++] = g->fDark;
                    if (kd->bGas) dst_of[km++] = g->fGas;
                    if (kd->bStar) dst_of[km++] = g->fStar;
                    if (kd->bMark) dst_of[km++] = g->fMark;
                    for (km = 0; km < nm; ++km)
                        memcpy(dst_of[km], o_pr + ((size_t)km * nv + k) * NMASSPROFILE, NMASSPROFILE * sizeof(float));
                }
                free(pt); free(vr); free(o_vc);
            } else {
                const int32_t *vi;
                const float *vd;
                if (sogpu_ball_gather_batch(kd->gpu, vc, vb, nv)) die_gpu("kdSO (sogpu_ball_gather_batch)");
                phase("2 Rvir ball gather", &tp);
                if (sogpu_members(kd->gpu, voff, &vi, &vd, 1)) die_gpu("kdSO (2 Rvir lists)");
                phase("2 Rvir lists (sorted)", &tp);
                for (k = 0; k < nv; ++k) {
                    int g = slots[k];
                    vcirc(kd, &kd->grps[g], rvir[g], mvir[g], vi + voff[k], vd + voff[k], voff[k + 1] - voff[k]);
                }
            }
        }
        free(slots); free(vc); free(vb); free(voff);
    }
    phase("kdVcirc / kdMassProfile", &tp);
    kd->dSOSeconds = wall() - t0;
    free(centers); free(rgtp); free(rvir); free(mvir); free(masses); free(ndelta); free(off); free(order);
    free(do_vcirc); free(slot_of_index); free(mem); free(in_conflict);
}

void kdFinish(KD kd)
{
    if (!kd) return;
    if (kd->gpu) sogpu_destroy(kd->gpu);
    free(kd->p.r); free(kd->p.v); free(kd->p.fMass); free(kd->p.fPhi);
    free(kd->p.iGrp); free(kd->p.nSubsumed); free(kd->p.nIgnored);
    free(kd->grps);
    free(kd->bMarkList);
    free(kd);
}
