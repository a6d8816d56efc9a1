/* so_oracle.c — plain-C CPU restatement of the reference SO hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Loaded by tests/, __graft_entry__.smoke() and bench.py's CPU
 * legs as the checker; never linked into or called by the product path.
 *
 * Parity: PINNED against the reference program compiled here from /root/reference
 * (oracle/_ref/so_ref and so_ref_inst) by tests/test_oracle_vs_ref.py and against the fixtures
 * that script leaves in tests/golden/.  The reference has no golden vectors of its own.
 *
 * Every function cites the reference lines it follows.  The neighbour search itself (the
 * kd-tree of kd2.c:1096-1185 walked by smooth2.c:58-114) is replaced by a plain cell list:
 * the set { particles with fDist2 <= fBall2 } does not depend on the search structure as long
 * as the periodic image chosen per particle equals the one the reference chooses per bucket
 * (kd2.h:165-252), which holds for ball radii < L/2 - bucket extent.
 *
 * Floating-point contract (compile with -ffp-contract=off, no -ffast-math):
 *   fDist2 = ((dx*dx) + (dy*dy)) + (dz*dz) in fp32, dx = sx - p.r[0], sx = x, x+L or x-L
 *   rhoEnclosed mixed fp32/fp64 exactly as written in kd2.c:588-593.
 */
#include "so_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float d2; int32_t idx; } nn_t;

struct so_oracle {
    const float *pos, *mass;
    int64_t pos_stride, mass_stride, n;
    float period[3];
    int nc;               /* cells per axis */
    int64_t *cell_start;  /* nc^3 + 1 */
    int32_t *cell_idx;    /* n, particle indices grouped by cell */
    nn_t *list;           /* last gather, sorted */
    int64_t n_list, cap_list;
    int32_t *out_idx;     /* split copies of list for the accessors */
    float *out_d2;
    int64_t cap_out;
    int64_t nevals;
};

/* ---- cell list (ours; replaces the kd-tree as a search structure only) -------------------- */

static int cell_of(const so_oracle_t *o, float x, int axis)
{
    double L = (double)o->period[axis];
    double t = (double)x / L;
    int c;
    t -= floor(t);
    c = (int)(t * o->nc);
    if (c >= o->nc) c = o->nc - 1;
    if (c < 0) c = 0;
    return c;
}

so_oracle_t *so_oracle_create(const float *pos, int64_t pos_stride, const float *mass,
                              int64_t mass_stride, int64_t n, const float period[3])
{
    so_oracle_t *o = (so_oracle_t *)calloc(1, sizeof(*o));
    int64_t i, nc3;
    int nc = 1;
    if (!o) return NULL;
    o->pos = pos; o->mass = mass; o->pos_stride = pos_stride; o->mass_stride = mass_stride;
    o->n = n;
    memcpy(o->period, period, sizeof(o->period));
    while ((int64_t)(nc * 2) * (nc * 2) * (nc * 2) * 4 <= n && nc < 512) nc *= 2;
    o->nc = nc;
    nc3 = (int64_t)nc * nc * nc;
    o->cell_start = (int64_t *)calloc((size_t)nc3 + 1, sizeof(int64_t));
    o->cell_idx = (int32_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(int32_t));
    if (!o->cell_start || !o->cell_idx) { so_oracle_destroy(o); return NULL; }
    for (i = 0; i < n; ++i) {
        const float *p = pos + i * pos_stride;
        int64_t c = ((int64_t)cell_of(o, p[2], 2) * nc + cell_of(o, p[1], 1)) * nc + cell_of(o, p[0], 0);
        o->cell_start[c + 1]++;
    }
    for (i = 0; i < nc3; ++i) o->cell_start[i + 1] += o->cell_start[i];
    {
        int64_t *fill = (int64_t *)malloc((size_t)nc3 * sizeof(int64_t));
        if (!fill) { so_oracle_destroy(o); return NULL; }
        memcpy(fill, o->cell_start, (size_t)nc3 * sizeof(int64_t));
        for (i = 0; i < n; ++i) {
            const float *p = pos + i * pos_stride;
            int64_t c = ((int64_t)cell_of(o, p[2], 2) * nc + cell_of(o, p[1], 1)) * nc + cell_of(o, p[0], 0);
            o->cell_idx[fill[c]++] = (int32_t)i;
        }
        free(fill);
    }
    return o;
}

void so_oracle_destroy(so_oracle_t *o)
{
    if (!o) return;
    free(o->cell_start); free(o->cell_idx); free(o->list); free(o->out_idx); free(o->out_d2);
    free(o);
}

void so_oracle_free(void *p) { free(p); }

/* ---- arithmetic of the reference ------------------------------------------------------------ */

/* kd2.c:588-593   r3 = r2*sqrt(r2);  return mass / (1.33333333*M_PI*r3);
 * r2 is float, sqrt() is the double libm routine, the product is rounded to float r3; the
 * quotient is formed in double and rounded to the float return value. */
float so_oracle_rho_enclosed(float mass, float r2)
{
    float r3 = (float)((double)r2 * sqrt((double)r2));
    return (float)((double)mass / (1.33333333 * M_PI * (double)r3));
}

/* smooth2.c:89-92 (dx = sx - p.r[0]; fDist2 = dx*dx + dy*dy + dz*dz, all float) with
 * sx = x + lx / x - lx / x chosen as kd2.h:165-194 does per bucket, restated per particle:
 * the image nearer to the particle wins. */
float so_oracle_dist2(const float c[3], const float p[3], const float period[3])
{
    float d[3];
    int k;
    for (k = 0; k < 3; ++k) {
        float x = c[k], l = period[k];
        float rd = x - p[k];
        float sx = x;
        if (rd > 0.5f * l) sx = x - l;
        else if (rd < -0.5f * l) sx = x + l;
        d[k] = sx - p[k];
    }
    {
        float xx = d[0] * d[0], yy = d[1] * d[1], zz = d[2] * d[2];
        float s = xx + yy;
        return s + zz;
    }
}

/* kd2.c:817-818   r3 = mass/((4./3.)*M_PI*fRhoVir);  r = pow(r3,0.3333333333); */
float so_oracle_rdelta(float mvir, float thr)
{
    float r3 = (float)((double)mvir / ((4. / 3.) * M_PI * (double)thr));
    return (float)pow((double)r3, 0.3333333333);
}

/* kd2.c:745 (fBall = fRgtp), 765-768 (fRootPeriod; while (fBall < 0.25*fRootPeriod) fBall *= 1.2) */
int so_oracle_schedule(float rgtp, const float period[3], float *balls, int cap)
{
    float s = period[0] * period[0];
    float root, ball = rgtp;
    int k = 0;
    s = s + period[1] * period[1];
    s = s + period[2] * period[2];
    root = (float)sqrt((double)s);
    while ((double)ball < 0.25 * (double)root) {
        ball = (float)((double)ball * 1.2);
        if (k < cap) balls[k] = ball;
        ++k;
        if (!(ball > 0.0f)) break; /* rgtp <= 0 would never terminate in the reference either */
    }
    return k;
}

/* ---- smBallGather + qsort ----------------------------------------------------------------- */

static int cmp_nn(const void *a, const void *b)
{
    const nn_t *p = (const nn_t *)a, *q = (const nn_t *)b;
    if (p->d2 < q->d2) return -1;   /* kd2.c:425-435 CmpList */
    if (p->d2 > q->d2) return 1;
    return (p->idx > q->idx) - (p->idx < q->idx);  /* tie order: ours (reference: gather order) */
}

static int push(so_oracle_t *o, float d2, int32_t idx)
{
    if (o->n_list >= o->cap_list) {
        int64_t nc = o->cap_list ? o->cap_list * 2 : 1024;
        nn_t *nl = (nn_t *)realloc(o->list, (size_t)nc * sizeof(nn_t));
        if (!nl) return -1;
        o->list = nl; o->cap_list = nc;
    }
    o->list[o->n_list].d2 = d2; o->list[o->n_list].idx = idx; o->n_list++;
    return 0;
}

static void axis_range(const so_oracle_t *o, float c, double b, int axis, int *lo, int *cnt)
{
    double L = (double)o->period[axis], h = L / o->nc;
    double a0 = floor(((double)c - b) / h) - 1.0, a1 = floor(((double)c + b) / h) + 1.0;
    double span = a1 - a0 + 1.0;
    if (span >= o->nc) { *lo = 0; *cnt = o->nc; return; }
    *lo = (int)a0; *cnt = (int)span;
}

int64_t so_oracle_ball(so_oracle_t *o, const float c[3], float ball2)
{
    double b = sqrt((double)ball2) * (1.0 + 1e-6);
    int lo[3], cnt[3], ix, iy, iz, nc = o->nc;
    int64_t i;
    o->n_list = 0;
    for (i = 0; i < 3; ++i) axis_range(o, c[i], b, (int)i, &lo[i], &cnt[i]);
    for (iz = 0; iz < cnt[2]; ++iz) {
        int cz = ((lo[2] + iz) % nc + nc) % nc;
        for (iy = 0; iy < cnt[1]; ++iy) {
            int cy = ((lo[1] + iy) % nc + nc) % nc;
            for (ix = 0; ix < cnt[0]; ++ix) {
                int cx = ((lo[0] + ix) % nc + nc) % nc;
                int64_t cell = ((int64_t)cz * nc + cy) * nc + cx, k;
                for (k = o->cell_start[cell]; k < o->cell_start[cell + 1]; ++k) {
                    int32_t pi = o->cell_idx[k];
                    float d2 = so_oracle_dist2(c, o->pos + (int64_t)pi * o->pos_stride, o->period);
                    o->nevals++;
                    if (d2 <= ball2)                     /* smooth2.c:95 */
                        if (push(o, d2, pi)) return -1;
                }
            }
        }
    }
    qsort(o->list, (size_t)o->n_list, sizeof(nn_t), cmp_nn);   /* kd2.c:781 */
    if (o->n_list > o->cap_out) {
        free(o->out_idx); free(o->out_d2);
        o->cap_out = o->n_list * 2;
        o->out_idx = (int32_t *)malloc((size_t)o->cap_out * sizeof(int32_t));
        o->out_d2 = (float *)malloc((size_t)o->cap_out * sizeof(float));
        if (!o->out_idx || !o->out_d2) return -1;
    }
    for (i = 0; i < o->n_list; ++i) { o->out_idx[i] = o->list[i].idx; o->out_d2[i] = o->list[i].d2; }
    return o->n_list;
}

const int32_t *so_oracle_ball_index(const so_oracle_t *o) { return o->out_idx; }
const float *so_oracle_ball_d2(const so_oracle_t *o) { return o->out_d2; }

static float mass_of(const so_oracle_t *o, int32_t i) { return o->mass[(int64_t)i * o->mass_stride]; }

/* ---- kdRvir, kd2.c:723-840 (bPot == 0) ---------------------------------------------------- */

int so_oracle_rvir(so_oracle_t *o, const float c[3], float rgtp, float thr, int n_members,
                   so_oracle_res_t *res)
{
    int64_t j, jlast = 0, n;
    float mass = 0.0f, ball = rgtp, ball2, root;
    float s = o->period[0] * o->period[0];
    int64_t ev0 = o->nevals;
    s = s + o->period[1] * o->period[1];
    s = s + o->period[2] * o->period[2];
    root = (float)sqrt((double)s);                                        /* kd2.c:765 */
    memset(res, 0, sizeof(*res));
    while ((double)ball < 0.25 * (double)root) {                          /* kd2.c:766 */
        ball = (float)((double)ball * 1.2);                               /* kd2.c:767 */
        ball2 = ball * ball;                                              /* kd2.c:768 */
        n = so_oracle_ball(o, c, ball2);                                  /* kd2.c:769,781 */
        if (n < 0) return -1;
        res->ngather++;
        if (!jlast) {
            if (n < n_members) {                                          /* kd2.c:772-778 */
                res->rvir = res->mvir = -1.0f; res->nevals = o->nevals - ev0; return 0;
            }
            for (j = 0; j < n_members - 1; ++j) mass += mass_of(o, o->list[j].idx);   /* 786-788 */
            if (so_oracle_rho_enclosed(mass, o->list[j - 1].d2) < thr &&
                so_oracle_rho_enclosed(mass + mass_of(o, o->list[j].idx), o->list[j].d2) < thr) {
                res->rvir = res->mvir = -2.0f; res->nevals = o->nevals - ev0; return 0;  /* 791-796 */
            }
            jlast = j;
        }
        for (j = jlast; j < n - 1; j++) {                                 /* kd2.c:804 */
            mass += mass_of(o, o->list[j].idx);                           /* kd2.c:807 */
            if (so_oracle_rho_enclosed(mass, o->list[j].d2) < thr &&
                so_oracle_rho_enclosed(mass + mass_of(o, o->list[j + 1].idx), o->list[j + 1].d2) < thr) {
                mass -= mass_of(o, o->list[j].idx);                       /* kd2.c:816 */
                res->mvir = mass;
                res->rvir = so_oracle_rdelta(mass, thr);                  /* kd2.c:817-820 */
                res->ndelta = (int32_t)j;                                 /* kd2.c:823 (n = j) */
                res->nevals = o->nevals - ev0;
                return 0;
            }
        }
        jlast = j;                                                        /* kd2.c:832 */
        if (!(ball > 0.0f)) break;
    }
    res->rvir = res->mvir = -3.0f;                                        /* kd2.c:837-839 */
    res->nevals = o->nevals - ev0;
    return 0;
}

int so_oracle_so(so_oracle_t *o, const float *centers, const float *rgtp, int h, float thr,
                 int n_members, float *rvir, float *mvir, int32_t *ndelta, int64_t *member_offset,
                 int32_t **members, int64_t *nevals)
{
    int i;
    int64_t cap = 1024, tot = 0, ev = 0;
    int32_t *mem = members ? (int32_t *)malloc((size_t)cap * sizeof(int32_t)) : NULL;
    if (members && !mem) return -1;
    for (i = 0; i < h; ++i) {
        so_oracle_res_t r;
        if (so_oracle_rvir(o, centers + 3 * i, rgtp[i], thr, n_members, &r)) { free(mem); return -1; }
        rvir[i] = r.rvir; mvir[i] = r.mvir; ndelta[i] = r.ndelta; ev += r.nevals;
        if (member_offset) member_offset[i] = tot;
        if (members && r.ndelta > 0) {
            if (tot + r.ndelta > cap) {
                int32_t *nm;
                while (tot + r.ndelta > cap) cap *= 2;
                nm = (int32_t *)realloc(mem, (size_t)cap * sizeof(int32_t));
                if (!nm) { free(mem); return -1; }
                mem = nm;
            }
            memcpy(mem + tot, o->out_idx, (size_t)r.ndelta * sizeof(int32_t));
        }
        tot += r.ndelta;
    }
    if (member_offset) member_offset[h] = tot;
    if (members) *members = mem;
    if (nevals) *nevals = ev;
    return 0;
}

/* ---- indexx, nr.c:91-151 (Numerical Recipes quicksort argsort, M=7, unstable) -------------- */

void so_oracle_indexx(int n, const float *arr0, int32_t *indx0)
{
    const float *arr = arr0 - 1;   /* 1-based views, as kdSortMass passes them (kd2.c:858) */
    int32_t *indx = indx0 - 1;
    int i, indxt, ir = n, itemp, j, k, l = 1, jstack = 0;
    int istack[64];
    float a;
#define SWP(x, y) do { itemp = (x); (x) = (y); (y) = itemp; } while (0)
    for (j = 1; j <= n; j++) indx[j] = j;
    for (;;) {
        if (ir - l < 7) {
            for (j = l + 1; j <= ir; j++) {
                indxt = indx[j]; a = arr[indxt];
                for (i = j - 1; i >= 1; i--) {
                    if (arr[indx[i]] <= a) break;
                    indx[i + 1] = indx[i];
                }
                indx[i + 1] = indxt;
            }
            if (jstack == 0) break;
            ir = istack[jstack--]; l = istack[jstack--];
        } else {
            k = (l + ir) >> 1;
            SWP(indx[k], indx[l + 1]);
            if (arr[indx[l + 1]] > arr[indx[ir]]) SWP(indx[l + 1], indx[ir]);
            if (arr[indx[l]] > arr[indx[ir]]) SWP(indx[l], indx[ir]);
            if (arr[indx[l + 1]] > arr[indx[l]]) SWP(indx[l + 1], indx[l]);
            i = l + 1; j = ir; indxt = indx[l]; a = arr[indxt];
            for (;;) {
                do i++; while (arr[indx[i]] < a);
                do j--; while (arr[indx[j]] > a);
                if (j < i) break;
                SWP(indx[i], indx[j]);
            }
            indx[l] = indx[j]; indx[j] = indxt;
            jstack += 2;
            if (jstack > 50) return;   /* nrerror("NSTACK too small") in the reference */
            if (ir - i + 1 >= j - l) { istack[jstack] = ir; istack[jstack - 1] = i; ir = j - 1; }
            else { istack[jstack] = j - 1; istack[jstack - 1] = l; l = i; }
        }
    }
#undef SWP
}

/* ---- tagging replay: kdSO order (kd2.c:873-879) + kdTagParticles (663-720) +
 *      kdZeroGroup (617-643) + kdFindGroup (647-660) + _VcmParticles (595-609) --------------- */

static void zero_group(int64_t n, int32_t *igrp, int32_t *nsub, float *rvir, float *mvir,
                       const int32_t *index, int small, int big)
{
    int64_t i;
    rvir[small] = (float)(-10.0 * index[big]);     /* kd2.c:633 */
    mvir[small] = -mvir[small];                    /* kd2.c:634 */
    for (i = 0; i < n; ++i)                        /* kd2.c:636-641 */
        if (igrp[i] == index[small]) { igrp[i] = 0; ++nsub[i]; }
}

int so_oracle_tag(const so_oracle_t *o, int h, const int32_t *index, const float *centers,
                  const float *gtp_mass, float *rvir, float *mvir, const int64_t *member_offset,
                  const int32_t *members, const float *vel, int64_t vel_stride,
                  int32_t *igrp, int32_t *nsub, int32_t *nign, float *vcm, int32_t *counts)
{
    int32_t *order = (int32_t *)malloc((size_t)(h > 0 ? h : 1) * sizeof(int32_t));
    int it;
    if (!order) return -1;
    memset(igrp, 0, (size_t)o->n * sizeof(int32_t));
    memset(nsub, 0, (size_t)o->n * sizeof(int32_t));
    memset(nign, 0, (size_t)o->n * sizeof(int32_t));
    counts[0] = counts[1] = 0;
    so_oracle_indexx(h, gtp_mass, order);                                   /* kd2.c:873 */
    for (it = 0; it < h; ++it) {
        int big = order[it] - 1, slurped = 0;                               /* kd2.c:879 */
        int64_t k, n = member_offset[big + 1] - member_offset[big];
        const int32_t *mem = members + member_offset[big];
        float mass_in = mvir[big];
        if (!(rvir[big] > 0.0f)) continue;                                  /* error codes: no tagging */
        for (k = 0; k < n; ++k) {                                           /* kd2.c:670 */
            int32_t p = mem[k];
            if (slurped) break;                                             /* kd2.c:671 */
            if (igrp[p] != 0) {                                             /* kd2.c:672 */
                int small = 0;
                float dx, dy, dz, r2;
                while (index[small] != igrp[p]) { ++small; if (small >= h) { free(order); return -2; } }
                dx = centers[3 * big + 0] - centers[3 * small + 0];         /* kd2.c:677-680 */
                dy = centers[3 * big + 1] - centers[3 * small + 1];
                dz = centers[3 * big + 2] - centers[3 * small + 2];
                {
                    float xx = dx * dx, yy = dy * dy, zz = dz * dz, s = xx + yy;
                    r2 = s + zz;
                }
                if (r2 <= rvir[big] * rvir[big]) {                          /* kd2.c:683 */
                    if (mvir[small] < 0.0f) { free(order); return -3; }     /* kd2.c:626-632 assert */
                    zero_group(o->n, igrp, nsub, rvir, mvir, index, small, big);
                    ++counts[0];                                            /* kd2.c:692 */
                    igrp[p] = index[big];                                   /* kd2.c:693 */
                } else if (r2 <= rvir[small] * rvir[small]) {               /* kd2.c:694 */
                    if (mvir[big] < 0.0f) { free(order); return -3; }
                    zero_group(o->n, igrp, nsub, rvir, mvir, index, big, small);
                    ++counts[1];                                            /* kd2.c:702 */
                    slurped = index[small];
                } else {
                    ++nign[p];                                              /* kd2.c:714 */
                }
            } else {
                igrp[p] = index[big];                                       /* kd2.c:717 */
            }
        }
        if (vel && vcm) {                                                   /* kd2.c:595-609 */
            float v[3] = {0.0f, 0.0f, 0.0f};
            int l;
            for (k = 0; k < n; ++k)
                for (l = 0; l < 3; ++l) {
                    float t = mass_of(o, mem[k]) * vel[(int64_t)mem[k] * vel_stride + l];
                    v[l] = v[l] + t;
                }
            for (l = 0; l < 3; ++l) vcm[3 * big + l] = v[l] / mass_in;
        }
    }
    free(order);
    return 0;
}

/* ---- kdVcirc / kdMassProfile, kd2.c:498-586 and 458-496 ------------------------------------------ */

int so_oracle_vcirc(so_oracle_t *o, const float c[3], float rvir, float mvir, float G, int n_members,
                    float *vcirc, float *rmass, float *rmax, float *vmax, float *profile,
                    const unsigned char *ptype_of, int ptype_mask)
{
    int i;
    int64_t j, n;
    float fBall, fBall2, mass, m, r, r2, vm, rm, vc, fmin, f;
    fmin = 2.0 / 8;                                                       /* kd2.c:507 (NVCIRC = 8) */
    fBall = 2. * rvir;                                                    /* kd2.c:511 */
    fBall2 = fBall * fBall;
    n = so_oracle_ball(o, c, fBall2);                                     /* kd2.c:513-514 */
    if (n <= 0) return -1;
    j = 0;
    mass = 0.0;
    for (f = fmin, i = 0; i < 8 - 1; ++i, f += fmin) {                    /* kd2.c:517-526 */
        r = f * rvir;
        r2 = r * r;
        while (j < n && o->list[j].d2 < r2) {
            mass += mass_of(o, o->list[j].idx);
            ++j;
        }
        vcirc[i] = sqrt(G * mass / r);
    }
    while (j < n) {                                                       /* kd2.c:528-531 */
        mass += mass_of(o, o->list[j].idx);
        ++j;
    }
    vcirc[8 - 1] = sqrt(G * mass / fBall);
    for (f = 0.25, i = 0; i < 2; ++i, f += 0.25) {                        /* kd2.c:537-546 */
        m = f * mvir;
        j = 0;
        mass = mass_of(o, o->list[0].idx);
        while (mass < m && j + 1 < n) {
            ++j;
            mass += mass_of(o, o->list[j].idx);
        }
        rmass[i] = sqrt(o->list[j].d2);
    }
    mass = 0.;                                                            /* kd2.c:551-569 */
    for (j = 0; j < n_members && j < n; ++j) mass += mass_of(o, o->list[j].idx);
    rm = sqrt(o->list[(n_members <= n ? n_members : n) - 1].d2);
    vm = sqrt(G * mass / rm);
    for (j = n_members; j < n; ++j) {
        mass += mass_of(o, o->list[j].idx);
        r = sqrt(o->list[j].d2);
        vc = sqrt(G * mass / r);
        if (vc > vm) {
            vm = vc;
            rm = r;
        }
    }
    *rmax = rm;
    *vmax = vm;
    if (profile) {                                                        /* kd2.c:458-496 */
        fmin = 2.0 / 16;
        j = 0;
        mass = 0.0;
        for (f = fmin, i = 0; i < 16 - 1; ++i, f += fmin) {
            r = f * rvir;
            r2 = r * r;
            while (j < n && o->list[j].d2 < r2) {
                if (!ptype_of || (ptype_of[o->list[j].idx] & ptype_mask)) mass += mass_of(o, o->list[j].idx);
                ++j;
            }
            profile[i] = mass;
        }
        while (j < n) {
            if (!ptype_of || (ptype_of[o->list[j].idx] & ptype_mask)) mass += mass_of(o, o->list[j].idx);
            ++j;
        }
        profile[16 - 1] = mass;
    }
    return 0;
}
