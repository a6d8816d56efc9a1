/* so_oracle.h — CPU restatement of the reference SO hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may load
 * this library; the product (so_b200/csrc, so_b200/host) never links or calls it.
 *
 * Parity status: PINNED against the reference binary built here from /root/reference
 * (oracle/_ref/so_ref, so_ref_inst; see oracle/Makefile and tests/test_oracle_vs_ref.py).
 * The reference ships no golden vectors or tests of its own.
 */
#ifndef SO_ORACLE_H
#define SO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct so_oracle so_oracle_t;

/* Particle set.  pos: x,y,z of particle i at pos[i*pos_stride + 0..2]; mass at mass[i*mass_stride]
 * (strides in floats; mass_stride 0 = one shared value).  Arrays are borrowed, not copied. */
so_oracle_t *so_oracle_create(const float *pos, int64_t pos_stride, const float *mass,
                              int64_t mass_stride, int64_t n, const float period[3]);
void so_oracle_destroy(so_oracle_t *o);

/* kd2.c:588-593 */
float so_oracle_rho_enclosed(float mass, float r2);
/* smooth2.c:89-92 with the image choice of kd2.h:165-252 restated per particle */
float so_oracle_dist2(const float c[3], const float p[3], const float period[3]);
/* kd2.c:817-818 */
float so_oracle_rdelta(float mvir, float thr);
/* kd2.c:745,765-768: the ball radii b_0..b_{K}; returns count written (<= cap) */
int so_oracle_schedule(float rgtp, const float period[3], float *balls, int cap);

/* smBallGather (smooth2.c:58-114) + qsort(CmpList) (kd2.c:425-435,781): all particles with
 * fDist2 <= ball2, ascending fDist2, ties by particle index.  Result stays valid until the next
 * call; returns the count. */
int64_t so_oracle_ball(so_oracle_t *o, const float c[3], float ball2);
const int32_t *so_oracle_ball_index(const so_oracle_t *o);
const float *so_oracle_ball_d2(const so_oracle_t *o);

typedef struct {
    float rvir;       /* R_Delta, or -1/-2/-3 */
    float mvir;       /* M_Delta, or -1/-2/-3 */
    int32_t ndelta;   /* N_Delta = j at kd2.c:823 (0 on error) */
    int32_t ngather;  /* balls gathered */
    int64_t nevals;   /* r^2 evaluations made by THIS restatement's cell grid */
} so_oracle_res_t;

/* kdRvir (kd2.c:723-840) without -pot.  On success the first res->ndelta entries of
 * so_oracle_ball_index()/so_oracle_ball_d2() are the members in sorted order. */
int so_oracle_rvir(so_oracle_t *o, const float c[3], float rgtp, float thr, int n_members,
                   so_oracle_res_t *res);

/* kdSO's loop (kd2.c:875-886) over h halos, in the order given, with no tagging:
 * fills rvir/mvir/ndelta[h]; member_offset[h+1]; *members = malloc'd concatenated sorted
 * member indices (caller frees with so_oracle_free); nevals optional. */
int so_oracle_so(so_oracle_t *o, const float *centers, const float *rgtp, int h, float thr,
                 int n_members, float *rvir, float *mvir, int32_t *ndelta, int64_t *member_offset,
                 int32_t **members, int64_t *nevals);
void so_oracle_free(void *p);

/* kdVcirc (kd2.c:498-586) + kdMassProfile (kd2.c:458-496) for one group with fRvir > 0: gathers and
 * sorts the 2*Rvir ball, then walks it exactly like the reference (sequential fp32 sums).
 * vcirc[8], rmass[2], *rmax, *vmax; profile[16] (may be NULL) = kdMassProfile of the particles whose
 * type flag ptype_of[i] & ptype_mask is non-zero (ptype_of NULL: every particle counts).
 * The reference's unguarded while(mass < m) (kd2.c:540) is bounded by the list length here. */
int so_oracle_vcirc(so_oracle_t *o, const float c[3], float rvir, float mvir, float G, int n_members,
                    float *vcirc, float *rmass, float *rmax, float *vmax, float *profile,
                    const unsigned char *ptype_of, int ptype_mask);

/* indexx (nr.c:91-151): ascending argsort, indx values are 1-based like the reference. */
void so_oracle_indexx(int n, const float *arr, int32_t *indx);

/* kdTagParticles / kdZeroGroup / kdFindGroup (kd2.c:617-720) + _VcmParticles (kd2.c:595-609),
 * replayed over all halos in indexx order of gtp_mass exactly as kdSO does.
 * In: per-halo catalog index[h] (1-based ids), centers, gtp_mass, rvir/mvir as produced by
 * so_oracle_so (modified in place for subsumed/slurped halos), sorted member lists.
 * vel may be NULL (then vcm is not computed).  Out: igrp/nsub/nign[n] (zero-initialised here),
 * vcm[h*3], counts[2] = {iGroupsRemoved, iGroupsSlurped}. */
int so_oracle_tag(const so_oracle_t *o, int h, const int32_t *index, const float *centers,
                  const float *gtp_mass, float *rvir, float *mvir, const int64_t *member_offset,
                  const int32_t *members, const float *vel, int64_t vel_stride,
                  int32_t *igrp, int32_t *nsub, int32_t *nign, float *vcm, int32_t *counts);

#ifdef __cplusplus
}
#endif
#endif
