/* sogpu.h — C-ABI of the B200 (sm_100a) spherical-overdensity hot path.
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, int return codes
 * (0 = ok, nonzero = error, text via sogpu_last_error()).  Nothing here throws or exits.
 * The entry points are what a binding of the reference's hot path needs — they replace
 *
 *      kdBuildTree(KD)                      /root/reference/kd2.h:269, kd2.c:1096-1185
 *      kdSO(KD, float rhovir, int nSmooth)  /root/reference/kd2.h:265, kd2.c:864-895
 *        └ kdRvir                           kd2.c:723-840
 *      smBallGather(SMX, float, float*)     /root/reference/smooth2.h:99, smooth2.c:58-114
 *
 * and the particle hand-over that kdReadTipsy does into PINIT[] (kd2.c:352-416).
 * INTEGRATION.md shows the few lines a maintainer of the reference adds to kd2.c to route
 * these calls here; so_b200/host/ is a complete host program built that way.
 *
 * There is no CPU fallback: every compute entry point fails with SOGPU_ERR_CUDA when no
 * sm_100-class device is usable.
 */
#ifndef SOGPU_H
#define SOGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOGPU_OK            0
#define SOGPU_ERR_ARG       1   /* bad argument / call order                    */
#define SOGPU_ERR_CUDA      2   /* CUDA runtime error (see sogpu_last_error)    */
#define SOGPU_ERR_NOMEM     3   /* host or device allocation failed             */
#define SOGPU_ERR_UNSUPPORTED 4 /* input outside the supported contract         */

typedef struct sogpu sogpu_t;

/* Per-halo error codes, stored in BOTH rvir and mvir exactly like kdRvir does
 * (kd2.c:774-776, 793-795, 837-838; documented so.c:146-154):
 *   -1  fewer than nMembers particles inside the first ball (1.2 fRgtp)
 *   -2  density already below threshold at nMembers particles
 *   -3  threshold never reached before the ball schedule ends (0.25 * |period|) */

/* ---- life cycle --------------------------------------------------------------------------- */

/* device: CUDA ordinal, or -1 for the current device. */
int sogpu_create(sogpu_t **out, int device);
void sogpu_destroy(sogpu_t *h);
/* Thread-local text of the last failure in this thread ("" if none). */
const char *sogpu_last_error(void);
/* All work of this handle is enqueued on `cuda_stream` (a cudaStream_t cast to void*;
 * NULL = the handle's own stream).  Lets a caller time the kernels with its own events. */
int sogpu_set_stream(sogpu_t *h, void *cuda_stream);
/* Tuning knob: grid build strategy. -1 auto (default: MSD partition levels, then a shared-memory
 * bucket sort, both storing from registers straight to the final slot); 0 = no partition levels
 * (one bucket holds everything; slow path for N > 3072); 2 = the staged variant of the same
 * levels (each tile is first sorted in shared memory, then copied out in coalesced runs). */
int sogpu_set_build_mode(sogpu_t *h, int mode);
/* Tuning knob: which ball of kdRvir's schedule b_k = rgtp * 1.2^k (kd2.c:745,765-768) is gathered
 * first (default 2; 1 = gather every ball like the reference).  Results do not depend on it: a ball
 * only examines what smaller balls have not, the -1 test always counts the schedule's first ball,
 * and the schedule's last ball is never skipped. */
int sogpu_set_first_ball(sogpu_t *h, int k);
/* Tuning knob: the 1024-thread class (cluster-size halos) stages its particles through TMA bulk copies
 * (cp.async.bulk + mbarrier ring in shared memory) instead of per-thread float4 loads.  Default off
 * (slower on B200 for this access pattern); results are identical. */
int sogpu_set_tma_staging(sogpu_t *h, int on);
/* Tuning knob: target mean particles per grid cell (default 2.0). */
int sogpu_set_cell_occupancy(sogpu_t *h, float particles_per_cell);

/* ---- particles: replaces the PINIT[] fill of kdReadTipsy (kd2.c:352-416) -------------------- */

/* Host particles, any AoS/SoA layout: x,y,z of particle i are three consecutive floats at
 * (char*)pos + i*pos_stride; its mass one float at (char*)mass + i*mass_stride (mass_stride 0 =
 * one shared value).  Works directly on PINIT[] (stride 60) or tipsy dark records (stride 36).
 * period/center are kdInit's fPeriod/fCenter (kd2.c:62-75); center only places the cell grid.
 * Data is packed to float4 {x,y,z,m} and copied to the device; the host arrays are not kept. */
int sogpu_set_particles_host(sogpu_t *h, const void *pos, size_t pos_stride, const void *mass,
                             size_t mass_stride, int64_t n, const float period[3],
                             const float center[3]);
/* Streaming ingest of RAW TIPSY records, as read from the file: `count` records of `floats_per_record`
 * 4-byte floats with the mass first and x, y, z next (gas 12, dark 9, star 11 floats, tipsydefs.h:6-37);
 * big_endian = 1 for -std (XDR) files: the byte swap of xdr_float (kd2.c:369,385,401) then happens on the
 * device.  Particle numbering follows the order of the calls (gas, dark, star: kd2.c:135-141).
 *   sogpu_ingest_begin(h, N, period, center); { read a chunk; sogpu_ingest_records(...); }*; sogpu_ingest_end(h)
 * With a page-locked buffer (sogpu_host_alloc) the copy is an asynchronous DMA that overlaps the
 * caller's next read; such a buffer may be refilled once the NEXT sogpu_ingest_records / _end call has
 * returned, i.e. the caller alternates two buffers.  Pageable buffers work too (synchronous copy). */
/* sogpu_ingest_keep_velocities(h, 1) before sogpu_ingest_begin: fields 4..6 of every record (vx, vy, vz) are
 * kept on the device as well, for sogpu_vcm. */
int sogpu_ingest_keep_velocities(sogpu_t *h, int on);
void *sogpu_host_alloc(size_t bytes);
void sogpu_host_free(void *p);
int sogpu_ingest_begin(sogpu_t *h, int64_t n_total, const float period[3], const float center[3]);
int sogpu_ingest_records(sogpu_t *h, const void *records, int64_t count, int32_t floats_per_record,
                         int32_t big_endian);
int sogpu_ingest_end(sogpu_t *h);

/* Pack + copy host particles (same layout rules) into a CALLER-owned device float4 array, without
 * touching the handle's particle state: used when the array is then replicated to other GPUs
 * (NCCL broadcast) and handed to each handle with sogpu_set_particles_device. */
int sogpu_upload_particles(sogpu_t *h, const void *pos, size_t pos_stride, const void *mass,
                           size_t mass_stride, int64_t n, void *d_xyzm_dst);
/* Device-resident float4 {x,y,z,m} array (borrowed: must outlive the handle's use of it). */
int sogpu_set_particles_device(sogpu_t *h, const void *d_xyzm, int64_t n, const float period[3],
                               const float center[3]);

/* ---- kdBuildTree replacement (kd2.c:1096-1185) ------------------------------------------- */

/* Counting sort of the particles by cell key (z-order over (iy,iz) rows, ix fastest) into a
 * periodic uniform grid: sorted float4 array + original-index array + cell-end table. */
int sogpu_build_grid(sogpu_t *h);

/* Focused variant: build the grid only where these nh halos can ever look — the cubes of half-width
 * b_k = rgtp * 1.2^n_balls (the radius of the n_balls-th ball of kdRvir's schedule) around the
 * centres, at a resolution of 2^min(lb,8) coarse cells per axis.  Particles elsewhere are not
 * sorted at all.  sogpu_so() on such a grid returns exactly the results of a full grid: every ball
 * is checked against the kept region on the device, and if a halo needs a bigger ball than
 * planned (or an arbitrary ball gather is requested) the library rebuilds the full grid and solves
 * again.  sogpu_so_device() reports such halos with code -103 instead (asynchronous, no retry). */
int sogpu_build_grid_for(sogpu_t *h, const float *centers, const float *rgtp, int32_t nh,
                         int32_t n_balls);
int sogpu_build_grid_for_device(sogpu_t *h, const void *d_centers, const void *d_rgtp, int32_t nh,
                                int32_t n_balls);

/* ---- kdSO / kdRvir replacement (kd2.c:864-895, 723-840), without tagging ------------------- */

/* For each of the nh halos (any order; halos are independent):
 *   rvir[i], mvir[i]  R_Delta / M_Delta exactly as kdRvir stores them, or -1/-2/-3 in both
 *   ndelta[i]         N_Delta = number of member particles (0 on error)
 * centers: nh*3 floats, rgtp: nh floats (GRPNODE.pos / .fRgtp, kd2.c:250-252,268-270).
 * rho_thr = fThreshold (so.c:477-481), n_members = kd->nMembers (so.c:229).
 * Host pointers.  Member lists are kept on the device until fetched with sogpu_members(). */
int sogpu_so(sogpu_t *h, const float *centers, const float *rgtp, int32_t nh, float rho_thr,
             int32_t n_members, float *rvir, float *mvir, int32_t *ndelta);

/* Same, with centers/rgtp already on the device and results left there:
 * d_out_n (int32[nh]) = N_Delta or error code (-1/-2/-3), d_out_m (float[nh]) = M_Delta.
 * Any of the outputs may be NULL.  Asynchronous on the handle's stream. */
int sogpu_so_device(sogpu_t *h, const void *d_centers, const void *d_rgtp, int32_t nh,
                    float rho_thr, int32_t n_members, void *d_out_n, void *d_out_m);

/* Turn the packed device results of sogpu_so_device (copied to the host by the caller) into
 * what kdRvir stores: rvir/mvir (error codes in both) and ndelta.  Host arithmetic only
 * (R_Delta = pow(M/((4/3) pi rho), 0.3333333333), kd2.c:817-818). */
int sogpu_finish_host(const int32_t *code_or_n, const float *m, int32_t nh, float rho_thr,
                      float *rvir, float *mvir, int32_t *ndelta);

/* Member particle lists (CSR) of the last sogpu_so()/sogpu_so_device() call.
 * offsets: nh+1 int64 (caller-allocated); *members / *d2: library-owned pinned host arrays,
 * valid until the next sogpu_so* call or sogpu_destroy.  Members of halo i are the ORIGINAL
 * particle indices (PINIT.iOrder, kd2.c:361) at [offsets[i], offsets[i+1]).
 * sorted = 0: the set, in no particular order (all that kdTagParticles needs unless a slurp
 *             occurs, kd2.c:694-703);
 * sorted = 1: ascending (fDist2, index) — the order kdTagParticles walks them (kd2.c:670,781).
 * d2 (may be NULL) and sorted = 1 need sogpu_keep_member_d2(h, 1) before the sogpu_so call. */
int sogpu_members(sogpu_t *h, int64_t *offsets, const int32_t **members, const float **d2,
                  int sorted);
int sogpu_keep_member_d2(sogpu_t *h, int on);

/* ---- smBallGather replacement (smooth2.c:58-114) + the qsort of kd2.c:514,781 -------------- */

/* All particles with fDist2 <= ball2 around center, ascending (fDist2, index).  Writes up to
 * cap entries into idx/d2 (host, either may be NULL) and the full count into *n. */
int sogpu_ball_gather(sogpu_t *h, const float center[3], float ball2, int32_t *idx, float *d2,
                      int64_t cap, int64_t *n);

/* Batched form: nh balls (centers nh*3, ball2 nh; host pointers).  The lists are then fetched with
 * sogpu_members() exactly like SO member lists (offsets, indices, r^2; sorted on request).  This
 * is the 2*Rvir gather of kdVcirc (kd2.c:511-514) for all halos in one pass. */
int sogpu_ball_gather_batch(sogpu_t *h, const float *centers, const float *ball2, int32_t nh);

/* ---- kdVcirc + kdMassProfile replacement (kd2.c:498-586, 458-496) ---------------------------- */

/* For each of the nh groups (centre, fRvir > 0, fMvir): gathers the ball of radius 2*fRvir
 * (kd2.c:511-514), sorts it by fDist2 and returns, exactly as the reference computes them in fp32:
 *   vcirc[8*i .. +8)   Vc = sqrt(G M(<r)/r) at r = 0.25, 0.5 .. 1.75 Rvir, and over the whole ball at 2 Rvir
 *   rmass[2*i .. +2)   radius of the particle at which the cumulative mass reaches Mvir/4, Mvir/2
 *   rmax[i], vmax[i]   first maximum of Vc over the sorted list from particle nMembers on
 *   profile[16*i..+16) cumulative mass of ALL particles inside r = 0.125 .. 1.875 Rvir, and in the whole
 *                      ball (may be NULL): the -dark/-gas/-star profile of a single-species snapshot
 * G is kdSetUniverse's constant (so.c: G = 1).  The sorted 2*Rvir lists stay available through
 * sogpu_members().  Equal particle masses only: returns SOGPU_ERR_UNSUPPORTED for a mixed-mass
 * snapshot (per-species sums then depend on which particle sits at which rank; the host program
 * walks the sorted lists itself in that case). */
int sogpu_vcirc(sogpu_t *h, const float *centers, const float *rvir, const float *mvir, int32_t nh,
                float G, int32_t nMembers, float *vcirc, float *rmass, float *rmax, float *vmax,
                float *profile);

/* The same for ANY particle masses and several species (gas + dark + star, -mark): the cumulative mass is the
 * reference's sequential fp32 sum over the sorted list, evaluated literally on the device (one warp per group).
 * ptype (host, N bytes, may be NULL = every particle counts everywhere): species bits of each particle;
 * masks[k] (k < nmasks <= 4): the particles with (ptype & masks[k]) != 0 count for profile k, as kdMassProfile's
 * per-species sums do (kd2.c:458-496); profiles: nmasks x nh x 16 floats, profile k of group i at
 * (k * nh + i) * 16.  Ties in r^2 between particles of different mass are ordered by particle index (the
 * reference orders them by its kd-tree walk: out of contract). */
int sogpu_vcirc_species(sogpu_t *h, const float *centers, const float *rvir, const float *mvir, int32_t nh,
                        float G, int32_t nMembers, const unsigned char *ptype, const int32_t *masks,
                        int32_t nmasks, float *vcirc, float *rmass, float *rmax, float *vmax, float *profiles);

/* ---- _VcmParticles replacement (kd2.c:595-609) -------------------------------------------------- */

/* Centre-of-mass velocity of every group of the last sogpu_so() call (same nh, sogpu_keep_member_d2 on):
 * vcm[3*i+l] = (sum over the members of group i, in ascending (fDist2, index) order, of fl(m * v[l])) / mvir[i],
 * the reference's sequential fp32 sum; zeros where mvir[i] <= 0.  Needs the velocities on the device
 * (streaming ingest with sogpu_ingest_keep_velocities). */
int sogpu_vcm(sogpu_t *h, const float *mvir, int32_t nh, float *vcm);

/* ---- kdTagParticles, order-independent part (kd2.c:663-720) ---------------------------------- */

/* Over the member lists of the last sogpu_so() call (nh groups, same order): a group that shares no
 * particle with any other group tags its members with its catalog id index[i] whatever the processing
 * order; groups that do share particles are reported in in_conflict[i] = 1 and their particles are
 * left untagged (0), for the caller to replay kdTagParticles over THOSE groups only, in kdSortMass
 * order (subsume / slurp / ignore depend on the order; a conflict-free group can never be met by
 * that replay).  igrp (host, N ints, may be NULL) receives the per-particle tags (PINIT.iGrp). */
int sogpu_tag_members(sogpu_t *h, const int32_t *index, int32_t nh, unsigned char *in_conflict,
                      int32_t *igrp);

/* ---- kdTagParticles, the order-dependent part (kd2.c:663-720, kdZeroGroup 617-643) ------------------------ */

/* Replays kdTagParticles on the device for the groups that sogpu_tag_members reported in conflict, in the
 * caller's processing order (the reference's: ascending catalog mass, kd2.c:873-879):
 *   order[n_order]   slots (0..nh-1) of the groups to replay — those with in_conflict set and rvir > 0
 *   index, centers   catalog id (GRPNODE.index, 1..max_index) and position of every slot
 *   rvir, mvir       in: kdRvir's results; out: with the subsume / slurp marks (-10 * index, -Mvir; kd2.c:633-634)
 *   igrp, nsubsumed, nignored   (host, N ints, any may be NULL) PINIT.iGrp / nSubsumed / nIgnored of every particle
 *   still_valid[nh]  1 if the group's radius was still positive right after its own pass (kd2.c:884: kdVcirc runs)
 * Requires the sorted member lists of the last sogpu_so call (sogpu_members(.., sorted = 1)) and a preceding
 * sogpu_tag_members over the same groups. */
int sogpu_tag_replay(sogpu_t *h, const int32_t *order, int32_t n_order, const int32_t *index, const float *centers,
                     float *rvir, float *mvir, int32_t nh, int32_t max_index, int32_t *igrp, int32_t *nsubsumed,
                     int32_t *nignored, int32_t *groups_removed, int32_t *groups_slurped, unsigned char *still_valid);

/* ---- several GPUs: domain runs ----------------------------------------- */

/* Every rank (one process per GPU) holds a SLICE of the snapshot and a spatially compact share of the
 * halos.  Its focus mask marks the coarse cells (2^min(lb,8) per axis, lb from n_total) its halos can
 * reach within n_balls steps of the ball schedule; the masks of all ranks are exchanged (they are a few
 * MB), then every rank routes each of its particles to every rank whose mask holds the particle's cell:
 *   sogpu_domain_mask_words   size of one mask in 32-bit words
 *   sogpu_domain_mask         this rank's mask -> d_mask (device)
 *   sogpu_domain_route_count  counts[r] = particles of this slice that rank r needs (host array)
 *   sogpu_domain_route_scatter  writes {x, y, z, global index} records of 16 bytes to dst[r] + dst_offset[r]
 *                             (device pointers: local buffers for an NCCL all-to-all, or the receivers'
 *                             own buffers mapped with sogpu_peer_open — the kernel then stores over NVLink)
 *   sogpu_set_particles_device_indexed  hands the received records to the grid build; all particles have
 *                             mass `mass`; n_total keeps the cell size identical on every rank
 * followed by sogpu_build_grid_for[_device] with the same halos and n_balls (balls that leave the mask
 * are reported with code -103 by sogpu_so_device and must be re-run with a larger n_balls) and sogpu_so*.
 * Member indices are the global ones.  d_slice is float4 {x,y,z,m}; masks are n_ranks consecutive masks. */
int sogpu_domain_mask_words(sogpu_t *h, int64_t n_total, int64_t *words);
int sogpu_domain_mask(sogpu_t *h, int64_t n_total, const float period[3], const float center[3],
                      const float *centers, const float *rgtp, int32_t nh, int32_t n_balls, void *d_mask);
int sogpu_domain_route_count(sogpu_t *h, int64_t n_total, const void *d_slice, int64_t n_slice,
                             const void *d_masks, int32_t n_ranks, int64_t *counts);
int sogpu_domain_route_scatter(sogpu_t *h, int64_t n_total, const void *d_slice, int64_t n_slice,
                               int64_t index_base, const void *d_masks, int32_t n_ranks,
                               void *const *dst, const int64_t *dst_offset);
int sogpu_set_particles_device_indexed(sogpu_t *h, const void *d_xyzi, int64_t n_local, int64_t n_total,
                                       float mass, const float period[3], const float center[3]);
/* receive buffers another process of the node can map (cudaIpc*): handle64 is 64 opaque bytes */
int sogpu_peer_alloc(sogpu_t *h, size_t bytes, void **ptr, void *handle64);
int sogpu_peer_open(sogpu_t *h, const void *handle64, void **ptr);
int sogpu_peer_close(sogpu_t *h, void *ptr);
int sogpu_peer_free(sogpu_t *h, void *ptr);

/* ---- several GPUs: the domain STEP, stream-ordered from the slice to the results ---------------------
 *
 * The calls above keep the host in the loop (masks and counts are read back to size the buffers).  The step
 * below does not: every rank derives the halo ownership and the destinations of every cell from the catalog on
 * its own device (identical integer arithmetic everywhere, so nothing is exchanged), keeps what some halo can
 * reach of its slice in ONE streaming pass, sorts those records out by destination tile by tile — one
 * system-scope atomic per tile on the receiver's cursor reserves the range, the run is stored straight into the
 * receiver's buffer over NVLink peer memory — and meets the other ranks at a flag barrier in peer memory.
 * The grid build then reads the number of records that arrived from the device.
 * (SOGPU_DIRECT_PUSH=0: the runs are staged locally and shipped in bulk, which is what stage_cap sizes.)
 *
 *   sogpu_domain_open     buffers of this rank; handles192 = three 64-byte cudaIpc handles (receive buffer 0,
 *                         receive buffer 1, control block) for the other processes of the node
 *   sogpu_domain_connect  pointers to every rank's buffers as seen from THIS process (sogpu_peer_open of the
 *                         handles, or the plain device pointers in a one-process run after sogpu_enable_peer_access)
 *   per step, every rank, same arguments:
 *     sogpu_domain_begin        whole catalog (device): ownership, destinations per cell, own focus mask
 *     sogpu_domain_route[_host] this rank's slice (or pieces of it), device float4 {x,y,z,m} / pinned host xyz
 *     sogpu_domain_push         hits -> receivers' buffers (reservations + stores over NVLink), barrier
 *     sogpu_domain_solve        grid over what arrived, SO solve of the owned halos; outputs cover the WHOLE
 *                               catalog: N_Delta / code and M_Delta for owned halos, 0x80808080 elsewhere;
 *                               code -103 = the halo's ball left the mask: step again with a larger n_balls
 *     sogpu_domain_result       (synchronises) records received / sent, error flags, owner of every halo
 * Member lists of the owned halos: sogpu_members (offsets over the whole catalog; indices are global). */
typedef struct {
    int32_t rank, n_ranks;
    int64_t n_total;          /* particles of the whole snapshot (fixes the cell size on every rank) */
    float mass;               /* the particle mass (domain steps are for equal-mass snapshots)       */
    float period[3], center[3];
    int64_t recv_cap;         /* records each receive buffer (and the list of this slice's hits) holds */
    int64_t stage_cap;        /* records one per-destination staging area holds (n_ranks > 1)        */
} sogpu_domain_cfg_t;
int sogpu_domain_open(sogpu_t *h, const sogpu_domain_cfg_t *cfg, void *handles192);
int sogpu_domain_connect(sogpu_t *h, void *const *recv0, void *const *recv1, void *const *ctrl);
int sogpu_domain_pointers(sogpu_t *h, void **recv0, void **recv1, void **ctrl);
int sogpu_enable_peer_access(sogpu_t *h, int peer_device);
int sogpu_domain_begin(sogpu_t *h, const void *d_centers, const void *d_rgtp, int32_t nh, int32_t n_balls);
int sogpu_domain_route(sogpu_t *h, const void *d_chunk, int64_t n, int64_t index_base);
int sogpu_domain_route_host(sogpu_t *h, const float *xyz_pinned, int64_t n, int64_t index_base, void *d_slice_dst);
int sogpu_domain_push(sogpu_t *h, int barrier);
int sogpu_domain_solve(sogpu_t *h, float rho_thr, int32_t n_members, void *d_out_n, void *d_out_m);
int sogpu_domain_result(sogpu_t *h, int64_t *n_recv, int64_t *n_sent, uint32_t *flags, unsigned char *owner);
int sogpu_domain_close(sogpu_t *h);
/* Helpers of a one-process, several-devices host program (`so -gpus N`, so_b200/host/kd_multi.c):
 *   sogpu_particles_device  the handle's particle array (device float4 {x,y,z,m}) and its length
 *   sogpu_copy              synchronous copy on the handle's device: kind 0 host->device, 1 device->host,
 *                           2 device->device (also across devices with peer access or through the driver)
 *   sogpu_set_members       installs CSR member lists (sorted by (r^2, index), e.g. merged from several devices)
 *                           as this handle's "last result" for sogpu_tag_members / _tag_replay / _vcm */
int sogpu_particles_device(sogpu_t *h, void **d_xyzm, int64_t *n);
int sogpu_copy(sogpu_t *h, void *dst, const void *src, size_t bytes, int kind);
int sogpu_set_members(sogpu_t *h, const int64_t *offsets, const int32_t *members, const float *d2, int32_t nh);

/* ---- introspection --------------------------------------------------------------------------- */

typedef struct {
    int64_t n_particles;
    int32_t cells_per_axis;
    int32_t equal_mass;          /* 1: all particle masses identical (fast exact path)        */
    int64_t last_evals;          /* r^2 evaluations of the last sogpu_so* call (all passes)   */
    int64_t last_evals_first;    /* ... of which in the first (histogram) pass of each ball   */
    int64_t last_members;        /* sum of N_Delta of the last call                           */
    int32_t last_kernel_launches;/* kernels launched by the last build or so call             */
    int32_t last_deferred;       /* halos the warp kernel handed to the block kernel          */
    int64_t n_in_grid;           /* particles the last build sorted (< n_particles if focused) */
} sogpu_stats_t;
int sogpu_get_stats(sogpu_t *h, sogpu_stats_t *out);
/* Debug (process started with SOGPU_DEBUG_TIMELINE=1): device clock in ns of the first CTA start and the
 * last CTA end of the last call's 1024-thread, 256-thread, warp and deferred query kernels (8 values used). */
int sogpu_debug_timeline(sogpu_t *h, uint64_t *out16);

/* ---- per-kernel timing (CUDA events on the handle's stream) ------------------------------------ */

int sogpu_profile_enable(sogpu_t *h, int on);
int sogpu_profile_kernels(void);                 /* number of kernel slots                    */
const char *sogpu_profile_name(int kernel_id);
/* Accumulated milliseconds and launch counts per kernel slot since the last reset
 * (synchronises the stream). */
int sogpu_profile_read(sogpu_t *h, double *ms, int64_t *launches, int n_slots, int reset);
/* Algorithmic bytes per slot for the kernels whose traffic is known at launch (grid build). */
int sogpu_profile_bytes(sogpu_t *h, double *bytes, int n_slots, int reset);

/* ---- host-side helpers of the exact-arithmetic contract (no GPU needed; unit-tested) -------- */

/* S[k] = sequential fp32 sum of k equal masses m, S[0]=0, S[k]=fl(S[k-1]+m) (kd2.c:787,807),
 * evaluated from the compressed segment table the kernels use.  Returns 0 on success. */
int sogpu_mass_prefix(float m, int64_t kmax, const int64_t *k, int64_t nk, float *out);
/* Ball radii of kdRvir's schedule (kd2.c:745,765-768); returns the count (<= cap written). */
int sogpu_ball_schedule(float rgtp, const float period[3], float *balls, int cap);
/* R_Delta from M_Delta (kd2.c:817-818). */
float sogpu_rdelta(float mvir, float rho_thr);

#ifdef __cplusplus
}
#endif
#endif /* SOGPU_H */
