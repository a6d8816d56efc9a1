/* kd_multi.c — `so -gpus N`: kdBuildTree + kdRvir of ONE catalog over N devices of this process.
 *
 * One host thread per device, each driving its own handle through the C-ABI's domain step (include/sogpu.h,
 * DESIGN.md section 8): thread r owns the slice [N r / R, N (r+1) / R) of the snapshot and the whole catalog;
 * per step every device derives the halo ownership and the destination table by itself, routes its slice,
 * pushes the records the other devices need straight into their buffers (peer memory), and solves the halos it
 * owns.  Results (R_Delta, M_Delta, N_Delta and the sorted member lists) are merged on the host in catalog
 * order: they do not depend on N, so every output file is byte-identical for any number of devices
 * (tests/test_host_cli_gpu.py::test_cli_several_devices_give_identical_files).
 *
 * Replaces nothing in the reference (it is single-threaded); it parallelises its hot loop so.c:515,540.
 */
#include "kd.h"

#include <assert.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXDEV 16

typedef struct {
    /* shared */
    KD kd;
    int R, h, nM;
    float thr;
    const float *centers, *rgtp;
    pthread_barrier_t *bar;
    void **recv0, **recv1, **ctrl;        /* every device's buffers, filled by the threads */
    volatile int *again, *failed;
    float *rvir, *mvir;                   /* merged per-halo results (disjoint entries per thread) */
    int32_t *ndelta;
    unsigned char *owner_out;
    /* per thread */
    int r, device;
    sogpu_t *g;
    int64_t *off;                         /* this device's CSR over the whole catalog (empty where not owned) */
    int32_t *mem;
    float *d2;
    char err[512];
} WORK;

static void fail(WORK *w, const char *what)
{
    snprintf(w->err, sizeof(w->err), "%s (device %d): %s", what, w->device, sogpu_last_error());
    *w->failed = 1;
}

static void *worker(void *arg)
{
    WORK *w = (WORK *)arg;
    KD kd = w->kd;
    const int R = w->R, r = w->r, h = w->h;
    const int64_t n = kd->nParticles, a = n * r / R, b = n * (r + 1) / R;
    void *d_all = NULL, *d_slice = NULL, *d_cat = NULL, *d_out = NULL;
    unsigned char hbuf[192], hdummy[64];
    int64_t n_all = 0;
    int balls = 4, q;
    sogpu_domain_cfg_t cfg;
    int32_t *code = (int32_t *)malloc((size_t)h * sizeof(int32_t));
    float *m = (float *)malloc((size_t)h * sizeof(float));
    unsigned char *owner = (unsigned char *)malloc((size_t)h);
    assert(code && m && owner);

    /* the slice: device 0 already holds the whole snapshot (it also runs the rest of kdSO); the others copy theirs */
    if (sogpu_particles_device(kd->gpu, &d_all, &n_all) || n_all != n) { fail(w, "particles"); goto sync1; }
    if (r == 0) {
        d_slice = (char *)d_all + (size_t)a * 16;
    } else {
        if (sogpu_peer_alloc(w->g, (size_t)(b - a) * 16 + 16, &d_slice, hdummy) ||
            sogpu_copy(w->g, d_slice, (char *)d_all + (size_t)a * 16, (size_t)(b - a) * 16, 2)) { fail(w, "slice copy"); goto sync1; }
    }
    if (sogpu_peer_alloc(w->g, (size_t)h * 16, &d_cat, hdummy) || sogpu_peer_alloc(w->g, (size_t)h * 8, &d_out, hdummy) ||
        sogpu_copy(w->g, d_cat, w->centers, (size_t)h * 12, 0) ||
        sogpu_copy(w->g, (char *)d_cat + (size_t)h * 12, w->rgtp, (size_t)h * 4, 0)) { fail(w, "catalog upload"); goto sync1; }
    memset(&cfg, 0, sizeof(cfg));
    cfg.rank = r; cfg.n_ranks = R; cfg.n_total = n; cfg.mass = kd->p.fMass[0];
    for (q = 0; q < 3; ++q) { cfg.period[q] = kd->fPeriod[q]; cfg.center[q] = kd->fCenter[q]; }
    cfg.recv_cap = n / R + n / 8 + (1 << 16);              /* generous: a rank receives far less than its share of N */
    if (cfg.recv_cap > n + 1024) cfg.recv_cap = n + 1024;
    cfg.stage_cap = R > 1 ? n / ((int64_t)R * R) + n / 16 + (1 << 16) : 0;
    if (sogpu_domain_open(w->g, &cfg, hbuf) || sogpu_domain_pointers(w->g, &w->recv0[r], &w->recv1[r], &w->ctrl[r])) fail(w, "domain_open");
sync1:
    pthread_barrier_wait(w->bar);
    if (*w->failed) goto done;
    if (R > 1) {
        for (q = 0; q < R; ++q)
            if (q != r && sogpu_enable_peer_access(w->g, w[q - r].device)) fail(w, "peer access");
        if (!*w->failed && sogpu_domain_connect(w->g, w->recv0, w->recv1, w->ctrl)) fail(w, "domain_connect");
    }
    pthread_barrier_wait(w->bar);
    if (*w->failed) goto done;
    if (sogpu_keep_member_d2(w->g, 1)) fail(w, "keep_member_d2");
    for (;;) {
        int64_t n_recv = 0, n_sent = 0;
        uint32_t flags = 0;
        int mine_outgrown = 0, i;
        if (sogpu_domain_begin(w->g, d_cat, (char *)d_cat + (size_t)h * 12, h, balls) ||
            sogpu_domain_route(w->g, d_slice, b - a, a) || sogpu_domain_push(w->g, 1) ||
            sogpu_domain_solve(w->g, w->thr, w->nM, d_out, (char *)d_out + (size_t)h * 4) ||
            sogpu_domain_result(w->g, &n_recv, &n_sent, &flags, owner)) { fail(w, "domain step"); }
        else if (flags) { snprintf(w->err, sizeof(w->err), "domain step (device %d): buffers too small or a device missing (flags %u)", w->device, flags); *w->failed = 1; }
        else if (sogpu_copy(w->g, code, d_out, (size_t)h * 4, 1) || sogpu_copy(w->g, m, (char *)d_out + (size_t)h * 4, (size_t)h * 4, 1)) fail(w, "results");
        if (!*w->failed)
            for (i = 0; i < h; ++i) mine_outgrown |= (owner[i] == r && code[i] == -103);
        if (mine_outgrown) *w->again = 1;
        pthread_barrier_wait(w->bar);
        if (*w->failed) goto done;
        q = *w->again;
        pthread_barrier_wait(w->bar);
        if (!q) break;
        if (r == 0) *w->again = 0;                          /* a ball left the mask somewhere: all devices again, further out */
        balls += 4;
        pthread_barrier_wait(w->bar);
    }
    {   /* what kdRvir stores for the halos this device owns, and their member lists */
        int i, nmine = 0;
        int32_t *cm = (int32_t *)malloc((size_t)h * sizeof(int32_t)), *nd = (int32_t *)malloc((size_t)h * sizeof(int32_t));
        float *mm = (float *)malloc((size_t)h * sizeof(float)), *rv = (float *)malloc((size_t)h * sizeof(float)),
              *mv = (float *)malloc((size_t)h * sizeof(float));
        int *slot = (int *)malloc((size_t)h * sizeof(int));
        const int32_t *lm;
        const float *ld;
        assert(cm && nd && mm && rv && mv && slot);
        for (i = 0; i < h; ++i)
            if (owner[i] == r) { cm[nmine] = code[i]; mm[nmine] = m[i]; slot[nmine++] = i; }
        if (nmine && sogpu_finish_host(cm, mm, nmine, w->thr, rv, mv, nd)) fail(w, "finish_host");
        for (i = 0; i < nmine && !*w->failed; ++i) { w->rvir[slot[i]] = rv[i]; w->mvir[slot[i]] = mv[i]; w->ndelta[slot[i]] = nd[i]; }
        if (r == 0) memcpy(w->owner_out, owner, (size_t)h);
        if (!*w->failed && sogpu_members(w->g, w->off, &lm, &ld, 1)) fail(w, "member lists");
        if (!*w->failed) {
            const size_t tot = (size_t)w->off[h];
            w->mem = (int32_t *)malloc((tot ? tot : 1) * sizeof(int32_t));
            w->d2 = (float *)malloc((tot ? tot : 1) * sizeof(float));
            assert(w->mem && w->d2);
            memcpy(w->mem, lm, tot * sizeof(int32_t));
            memcpy(w->d2, ld, tot * sizeof(float));
        }
        free(cm); free(nd); free(mm); free(rv); free(mv); free(slot);
    }
done:
    pthread_barrier_wait(w->bar);                          /* nobody closes its buffers while a peer may still write */
    sogpu_domain_close(w->g);
    if (r != 0 && d_slice) sogpu_peer_free(w->g, d_slice);
    if (d_cat) sogpu_peer_free(w->g, d_cat);
    if (d_out) sogpu_peer_free(w->g, d_out);
    free(code); free(m); free(owner);
    return NULL;
}

/* R_Delta / M_Delta / N_Delta of every group and the sorted member lists (CSR, catalog order) over kd->nGpus devices.
 * Returns 0; *mem / *d2 are malloc'd.  Afterwards kd->gpu holds the full snapshot and a full grid again. */
int kdRvirSeveralDevices(KD kd, const float *centers, const float *rgtp, int h, float thr, float *rvir, float *mvir,
                         int32_t *ndelta, int64_t *off, int32_t **mem, float **d2)
{
    const int R = kd->nGpus;
    WORK w[MAXDEV];
    pthread_t th[MAXDEV];
    pthread_barrier_t bar;
    void *recv0[MAXDEV], *recv1[MAXDEV], *ctrl[MAXDEV], *d_all = NULL;
    volatile int again = 0, failed = 0;
    unsigned char *owner = (unsigned char *)malloc((size_t)h);
    int64_t n_all = 0, tot = 0;
    int r, i;
    assert(R >= 1 && R <= MAXDEV && owner);
    memset(w, 0, sizeof(w));
    pthread_barrier_init(&bar, NULL, (unsigned)R);
    for (r = 0; r < R; ++r) {
        w[r].kd = kd; w[r].R = R; w[r].h = h; w[r].nM = kd->nMembers; w[r].thr = thr;
        w[r].centers = centers; w[r].rgtp = rgtp; w[r].bar = &bar;
        w[r].recv0 = recv0; w[r].recv1 = recv1; w[r].ctrl = ctrl; w[r].again = &again; w[r].failed = &failed;
        w[r].rvir = rvir; w[r].mvir = mvir; w[r].ndelta = ndelta; w[r].owner_out = owner;
        w[r].r = r; w[r].device = (kd->iDevice < 0 ? 0 : kd->iDevice) + r;
        w[r].off = (int64_t *)calloc((size_t)h + 1, sizeof(int64_t));
        assert(w[r].off);
        if (r == 0) w[r].g = kd->gpu;
        else if (sogpu_create(&w[r].g, w[r].device)) {
            fprintf(stderr, "ERROR in kdSO (-gpus %d, device %d): %s\n", R, w[r].device, sogpu_last_error());
            exit(1);
        }
    }
    if (sogpu_particles_device(kd->gpu, &d_all, &n_all)) { fprintf(stderr, "ERROR in kdSO: %s\n", sogpu_last_error()); exit(1); }
    for (r = 0; r < R; ++r) pthread_create(&th[r], NULL, worker, &w[r]);
    for (r = 0; r < R; ++r) pthread_join(th[r], NULL);
    pthread_barrier_destroy(&bar);
    if (failed) {
        for (r = 0; r < R; ++r)
            if (w[r].err[0]) fprintf(stderr, "ERROR in kdSO (-gpus %d): %s\n", R, w[r].err);
        exit(1);
    }
    /* merge the member lists in catalog order: every group comes from the device that owns it */
    off[0] = 0;
    for (i = 0; i < h; ++i) off[i + 1] = off[i] + (ndelta[i] > 0 ? ndelta[i] : 0);
    tot = off[h];
    *mem = (int32_t *)malloc((size_t)(tot ? tot : 1) * sizeof(int32_t));
    *d2 = (float *)malloc((size_t)(tot ? tot : 1) * sizeof(float));
    assert(*mem && *d2);
    for (i = 0; i < h; ++i) {
        const WORK *o = &w[owner[i]];
        const int64_t k = o->off[i + 1] - o->off[i];
        assert(k == off[i + 1] - off[i]);
        memcpy(*mem + off[i], o->mem + o->off[i], (size_t)k * sizeof(int32_t));
        memcpy(*d2 + off[i], o->d2 + o->off[i], (size_t)k * sizeof(float));
    }
    for (r = 0; r < R; ++r) {
        free(w[r].off); free(w[r].mem); free(w[r].d2);
        if (r) sogpu_destroy(w[r].g);
    }
    free(owner);
    /* device 0 goes on with kdTagParticles / _VcmParticles / kdVcirc: whole snapshot, full grid, the merged lists */
    if (sogpu_set_particles_device(kd->gpu, d_all, n_all, kd->fPeriod, kd->fCenter) || sogpu_build_grid(kd->gpu) ||
        sogpu_set_members(kd->gpu, off, *mem, *d2, h)) {
        fprintf(stderr, "ERROR in kdSO (-gpus %d): %s\n", R, sogpu_last_error());
        exit(1);
    }
    return 0;
}
