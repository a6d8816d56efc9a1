/* sogpu.cu — B200 (sm_100a) implementation of the SO hot path behind include/sogpu.h.
 *
 * Replaces, from /root/reference:
 *   kdBuildTree  kd2.c:1096-1185  ->  k_cell_count / k_scan_* / k_scatter  (counting sort by
 *                                     cell key into a periodic uniform grid; rows along x are
 *                                     contiguous, rows are z-ordered over (iy,iz))
 *   smBallGather smooth2.c:58-114 ->  for_each_in_ball<>  (row segments overlapping the sphere,
 *                                     flat coalesced float4 loads, exact fp32 r^2)
 *   qsort+kdRvir kd2.c:781-831    ->  so_halo<>  (log-r^2 histogram from the float bits, prefix
 *                                     sum, conservative bracket, exact sort inside the bracket,
 *                                     first j with rho(j) and rho(j+1) below threshold)
 *
 * Exactness (see so_math.h): r^2 with the reference's operation order and no FMA; density test
 * with the reference's mixed fp32/fp64 expression; the enclosed mass is the reference's
 * sequential fp32 sum, which for equal-mass particles depends on the rank only (mass table).
 *
 * No CPU fallback: every entry point that computes returns SOGPU_ERR_CUDA without a device.
 */
#include "../../include/sogpu.h"
#include "so_math.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cstddef>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

/* ============================================================================================
 * constants
 * ============================================================================================ */
#define NB 512          /* histogram bins per level                                         */
#define NB_LOG 9
#define SHIFT0 18       /* level 0: 2^(23-18)=32 bins per octave of r^2, 16 octaves          */
#define CELL_MARGIN 2.0e-3  /* slack, in cells, of every cell-range computation              */
#define CODE_UNSUPPORTED (-100)
#define CODE_DEFER (-101)
#define CODE_NEED_FULL (-103)   /* focused grid does not cover this halo's ball */

#ifndef BKT_AVG
#define BKT_AVG 1024   /* mean particles per final bucket */
#endif
#ifndef Q256_MINB
#define Q256_MINB 2     /* 256-thread class: <= 128 registers */
#endif
#ifndef Q32_MINB
#define Q32_MINB 3      /* resident CTAs per SM the warp-per-halo kernel is compiled for: 80 registers, so that one
                         * of its CTAs fits next to a 256-thread-class CTA (<= 128 registers) on an SM — at 118 + 130
                         * registers the two excluded each other by 0.1 % of the register file */
#endif
template <int NT> struct Cfg;
/* SCAP: (r^2 bits, index) keys of ONE ball that a group can stage in shared memory during the histogram pass.
 * A ball that fits is traversed in global memory exactly once: window collection, refinement and the member
 * emission then read the staged keys. */
template <> struct Cfg<32> {   /* warp per halo */
    static const int CAP = 256, WTARGET = 128, NLEV = 1, GROUPS = 8, MINB = Q32_MINB, SCAP = 384;
};
template <> struct Cfg<256> {  /* block per halo */
    static const int CAP = 1024, WTARGET = 256, NLEV = 4, GROUPS = 1, MINB = Q256_MINB, SCAP = 6144;
};
template <> struct Cfg<1024> { /* one full-SM block per halo: cluster-size halos (>= ~10^5 particles) */
    static const int CAP = 4096, WTARGET = 2048, NLEV = 4, GROUPS = 1, MINB = 1, SCAP = 4096;
};

/* ============================================================================================
 * error handling
 * ============================================================================================ */
static thread_local char g_err[512] = "";

static int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return set_err(SOGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                    \
                           cudaGetErrorString(e_), __FILE__, __LINE__);                       \
    } while (0)

extern "C" const char *sogpu_last_error(void) { return g_err; }

/* ============================================================================================
 * device-side grid description
 * ============================================================================================ */
struct GridDev {
    const float4 *sorted;     /* particles in cell order: {x, y, z, original index as int bits} */
    const uint32_t *ce;       /* ce[c] = first sorted slot of cell c, ce[ncell] = N          */
    int nc, lb;               /* cells per axis (power of two), log2                         */
    int tb;                   /* rows are ordered in tiles of 2^tb x 2^tb (iy, iz)           */
    int indexed;              /* input .w already holds the particle's (global) index, not its mass */
    int use_tma;              /* 1024-thread class: stage the particles through TMA bulk copies */
    float g0[3], invh[3];     /* cell coordinate = floor((x - g0) * invh) & (nc-1)           */
    float L[3], halfL[3];
    double dg0[3], dinvh[3], dh[3];
    double bmax_pruned;       /* balls at least this large visit every cell                  */
    const uint32_t *mask;     /* focused build: bit per coarse cell that was kept (NULL = all) */
    int nofilter;             /* the build keeps every input particle (the mask still bounds the balls) */
    int mb, ms;               /* mask cells per axis = 2^mb; fine cell coordinate >> ms        */
    double mask_rmin;         /* smallest half-width a halo marks in the mask                  */
};

/* rows of cells (fixed iy, iz) are contiguous along x; rows are ordered in 8x8 tiles of (iy, iz),
 * tiles row-major: a ball's rows fall into a handful of contiguous stretches of the sorted array, and
 * the key costs a few shifts (a Morton interleave of iy, iz gave the same locality for 3x the ALU work) */
__device__ __forceinline__ uint32_t row_key(uint32_t iy, uint32_t iz, int lb, int tb)
{
    const uint32_t m = (1u << tb) - 1u;
    return ((((iz >> tb) << (lb - tb)) | (iy >> tb)) << (2 * tb)) | ((iz & m) << tb) | (iy & m);
}
__device__ __forceinline__ uint32_t cell_coord(float x, float g0, float invh, int mask)
{
    float t = __fmul_rn(__fsub_rn(x, g0), invh);
    return (uint32_t)((int)floorf(t) & mask);
}
__device__ __forceinline__ uint32_t cell_key(const float4 &p, const GridDev &g)
{
    int mask = g.nc - 1;
    uint32_t ix = cell_coord(p.x, g.g0[0], g.invh[0], mask);
    uint32_t iy = cell_coord(p.y, g.g0[1], g.invh[1], mask);
    uint32_t iz = cell_coord(p.z, g.g0[2], g.invh[2], mask);
    return (row_key(iy, iz, g.lb, g.tb) << g.lb) | ix;
}

/* the low `bits` bits of cell_key (the cell inside a final bucket): often x alone decides them */
__device__ __forceinline__ uint32_t cell_key_low(const float4 &p, const GridDev &g, int bits)
{
    const int mask = g.nc - 1;
    const uint32_t cm = (1u << bits) - 1u;
    uint32_t ix = cell_coord(p.x, g.g0[0], g.invh[0], mask);
    if (bits <= g.lb) return ix & cm;
    uint32_t iy = cell_coord(p.y, g.g0[1], g.invh[1], mask);
    uint32_t iz = cell_coord(p.z, g.g0[2], g.invh[2], mask);
    if (bits <= g.lb + 2 * g.tb) {
        const uint32_t m = (1u << g.tb) - 1u;
        return (((((iz & m) << g.tb) | (iy & m)) << g.lb) | ix) & cm;
    }
    return ((row_key(iy, iz, g.lb, g.tb) << g.lb) | ix) & cm;
}

__device__ __forceinline__ bool mask_bit(const GridDev &g, uint32_t mx, uint32_t my, uint32_t mz)
{
    uint32_t bit = (mz << (2 * g.mb)) | (my << g.mb) | mx;
    return (__ldg(g.mask + (bit >> 5)) >> (bit & 31)) & 1u;
}
/* cell key + whether the particle's coarse cell belongs to the focused region */
__device__ __forceinline__ uint32_t cell_key_kept(const float4 &p, const GridDev &g, bool &kept)
{
    int mask = g.nc - 1;
    uint32_t ix = cell_coord(p.x, g.g0[0], g.invh[0], mask);
    uint32_t iy = cell_coord(p.y, g.g0[1], g.invh[1], mask);
    uint32_t iz = cell_coord(p.z, g.g0[2], g.invh[2], mask);
    kept = !g.mask || g.nofilter || mask_bit(g, ix >> g.ms, iy >> g.ms, iz >> g.ms);
    return (row_key(iy, iz, g.lb, g.tb) << g.lb) | ix;
}

/* mask bit of the coarse cell that holds the cell with this key (inverse of row_key) */
__device__ __forceinline__ uint32_t cell_mask_bit(const GridDev &g, uint32_t key)
{
    const uint32_t ix = key & (uint32_t)(g.nc - 1), rk = key >> g.lb;
    const uint32_t m = (1u << g.tb) - 1u, lo = rk & ((1u << (2 * g.tb)) - 1u), hi = rk >> (2 * g.tb);
    const uint32_t iy = ((hi & ((1u << (g.lb - g.tb)) - 1u)) << g.tb) | (lo & m);
    const uint32_t iz = ((hi >> (g.lb - g.tb)) << g.tb) | (lo >> g.tb);
    return ((iz >> g.ms) << (2 * g.mb)) | ((iy >> g.ms) << g.mb) | (ix >> g.ms);
}
__device__ __forceinline__ bool cell_in_mask(const GridDev &g, uint32_t key)
{
    const uint32_t ix = key & (uint32_t)(g.nc - 1), rk = key >> g.lb;
    const uint32_t m = (1u << g.tb) - 1u, lo = rk & ((1u << (2 * g.tb)) - 1u), hi = rk >> (2 * g.tb);
    const uint32_t iy = ((hi & ((1u << (g.lb - g.tb)) - 1u)) << g.tb) | (lo & m);
    const uint32_t iz = ((hi >> (g.lb - g.tb)) << g.tb) | (lo >> g.tb);
    return mask_bit(g, ix >> g.ms, iy >> g.ms, iz >> g.ms);
}

__device__ __forceinline__ float4 ld_stream(const float4 *p)
{
    float4 r;
    asm("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}

/* ============================================================================================
 * grid build kernels (kdBuildTree replacement)
 * ============================================================================================ */
#include "grid_build.cuh"

__global__ void k_copy_u32(uint32_t *dst, const uint32_t *src)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst = *src;
}

__global__ void k_store_u32(uint32_t *p, uint32_t v0, uint32_t *q, uint32_t v1)
{
    if (threadIdx.x == 0 && blockIdx.x == 0) { *p = v0; if (q) *q = v1; }
}

/* K1 (pack): raw host layout -> float4 {x,y,z,m} on the device (xyz triplets + one shared mass) */
__global__ void __launch_bounds__(256) k_expand_xyz(const float *__restrict__ xyz, float m, float4 *__restrict__ dst,
                                                    int64_t n)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float *p = xyz + 3 * i;
        dst[i] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), m);
    }
}

/* K1 (ingest): raw TIPSY records -> float4 {x,y,z,m}.  A record is `nf` 4-byte floats, mass first, then
 * x, y, z (gas 12, dark 9, star 11 floats: tipsydefs.h:6-37); `swap` = the file is XDR / big-endian
 * (-std, kd2.c:32-44,369,385,401), the byte swap happens here instead of in xdr_float. */
__global__ void __launch_bounds__(256) k_ingest_records(const uint32_t *__restrict__ raw, int64_t count, int nf, int swap,
                                                        float4 *__restrict__ dst, float4 *__restrict__ vel)
{
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const uint32_t *r = raw + i * nf;
        uint32_t m = __ldg(r), x = __ldg(r + 1), y = __ldg(r + 2), z = __ldg(r + 3);
        if (swap) { m = __byte_perm(m, 0, 0x0123); x = __byte_perm(x, 0, 0x0123); y = __byte_perm(y, 0, 0x0123); z = __byte_perm(z, 0, 0x0123); }
        dst[i] = make_float4(__uint_as_float(x), __uint_as_float(y), __uint_as_float(z), __uint_as_float(m));
        if (vel) {                                           /* fields 4..6 of every record type: vx, vy, vz */
            uint32_t vx = __ldg(r + 4), vy = __ldg(r + 5), vz = __ldg(r + 6);
            if (swap) { vx = __byte_perm(vx, 0, 0x0123); vy = __byte_perm(vy, 0, 0x0123); vz = __byte_perm(vz, 0, 0x0123); }
            vel[i] = make_float4(__uint_as_float(vx), __uint_as_float(vy), __uint_as_float(vz), 0.0f);
        }
    }
}

/* ============================================================================================
 * group (warp or block) primitives
 * ============================================================================================ */
template <int NT> __device__ __forceinline__ void gsync()
{
    if (NT == 32) __syncwarp(); else __syncthreads();
}

template <int NT> __device__ __forceinline__ uint32_t gscan_incl(uint32_t v, uint32_t *tmp, int tid)
{
    int lane = tid & 31;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += t;
    }
    if (NT > 32) {
        int w = tid >> 5;
        if (lane == 31) tmp[w] = v;
        __syncthreads();
        if (w == 0) {
            uint32_t s = (lane < NT / 32) ? tmp[lane] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
                if (lane >= o) s += t;
            }
            if (lane < NT / 32) tmp[lane] = s;
        }
        __syncthreads();
        if (w > 0) v += tmp[w - 1];
        __syncthreads();
    }
    return v;
}

template <int NT> __device__ __forceinline__ int gmin(int v, uint32_t *tmp, int tid)
{
    v = __reduce_min_sync(0xFFFFFFFFu, v);
    if (NT > 32) {
        int lane = tid & 31, w = tid >> 5;
        if (lane == 0) tmp[w] = (uint32_t)v;
        __syncthreads();
        if (w == 0) {
            int s = (lane < NT / 32) ? (int)tmp[lane] : INT_MAX;
            s = __reduce_min_sync(0xFFFFFFFFu, s);
            if (lane == 0) tmp[0] = (uint32_t)s;
        }
        __syncthreads();
        v = (int)tmp[0];
        __syncthreads();
    }
    return v;
}

template <int NT> __device__ __forceinline__ uint32_t gsum(uint32_t v, uint32_t *tmp, int tid)
{
    v = __reduce_add_sync(0xFFFFFFFFu, v);
    if (NT > 32) {
        int lane = tid & 31, w = tid >> 5;
        if (lane == 0) tmp[w] = v;
        __syncthreads();
        if (w == 0) {
            uint32_t s = (lane < NT / 32) ? tmp[lane] : 0u;
            s = __reduce_add_sync(0xFFFFFFFFu, s);
            if (lane == 0) tmp[0] = s;
        }
        __syncthreads();
        v = tmp[0];
        __syncthreads();
    }
    return v;
}

/* ============================================================================================
 * per-group shared memory
 * ============================================================================================ */
struct MassTableS {   /* CTA-shared copy of the mass table */
    int n;
    float m;
    uint32_t k0[SO_MT_MAX + 1];
    float s0[SO_MT_MAX];
    float inc[SO_MT_MAX];
};

/* Staging of the 1024-thread class (cluster-size halos): the particles of a ball are pulled into shared
 * memory by TMA bulk copies (cp.async.bulk, one per row segment piece, completion on an mbarrier) through
 * a ring of TMA_STAGES tiles (full / empty mbarriers, no block-wide barrier per tile), instead of
 * per-thread loads.  Selectable (sogpu_set_tma_staging) and OFF by default: measured on B200 with 64
 * halos of 10^6 particles the plain coalesced float4 loads (4 in flight per thread, 64 KB per SM) already
 * pull 2.7 TB/s over the 64 busy SMs and finish the solve in 2.5 ms; the TMA ring needs 3.3 ms
 * (profiles/r1b_build_experiments.md). */
#define TMA_TILE 2048
#define TMA_STAGES 4
template <int NT> struct TmaStage { };
template <> struct TmaStage<1024> {
    float4 tile[TMA_STAGES][TMA_TILE];          /* 4 x 32 KB: three tiles in flight while one is read */
    unsigned long long full[TMA_STAGES];        /* producer: expect_tx + the copies' complete_tx     */
    unsigned long long empty[TMA_STAGES];       /* consumers: one arrival per warp after reading     */
    uint32_t tiles_done;                        /* tiles staged so far by this CTA: stage and phase of the next */
};

template <int NT> struct GroupSmem {
    unsigned long long wkey[Cfg<NT>::CAP];      /* window: (r^2 bits << 32) | original index   */
    uint32_t hist[Cfg<NT>::NLEV][NB + 1];       /* per level: counts, then exclusive prefix    */
    uint32_t seg_start[2 * NT];                 /* row segments of the current batch           */
    uint32_t seg_pre[2 * NT];                   /* inclusive prefix of their lengths           */
    uint32_t tmp[40];                           /* scan / reduce scratch                       */
    uint32_t cnt;                               /* append cursor                               */
    uint32_t scnt;                              /* keys appended to skey during the histogram pass */
    uint32_t bcast[4];
    unsigned long long bcast64;
    uint8_t wflag[Cfg<NT>::CAP];                /* below-threshold flag per window element     */
    /* LAST: the staged keys of the current ball — or, for the 1024-thread class with TMA staging on, the tiles
     * of the bulk-copy ring (the two are never used together) */
    union alignas(128) {
        unsigned long long skey[Cfg<NT>::SCAP];
        TmaStage<NT> tma;
    };
};

__device__ __forceinline__ float mt_eval(const MassTableS &mt, uint32_t k)
{
    return so_mass_prefix_eval(mt.k0, mt.s0, mt.inc, mt.n, k);
}

/* Cheap fp32 screening of rhoEnclosed(S[k], r2) < thr (kd2.c:588-593): S[k] differs from k*m by at most
 * k * 2^-24 relative (k roundings of at most half an ulp of the running sum each), and the fp32 evaluation of
 * m k / (C r2^1.5) is good to a few ulp.  Outside a band of (1e-5 + 2.4e-7 k) around the threshold the answer of
 * the exact mixed-precision expression is certain; only inside it the mass table and the fp64 sqrt / div are
 * evaluated.  Returns 0 = certainly not below, 1 = certainly below, 2 = undecided. */
__device__ __forceinline__ int rho_screen(const MassTableS &mt, uint32_t k, float r2, float thr)
{
    if (k >= (1u << 22) || !(r2 > 0.0f)) return 2;
    const float kf = (float)k;
    const float eps = 1.0e-5f + 2.4e-7f * kf;
    const float rho = __fdividef(kf * mt.m, 4.18879f * r2 * sqrtf(r2));
    if (rho > thr * (1.0f + eps)) return 0;
    if (rho < thr * (1.0f - eps)) return 1;
    return 2;
}
__device__ __forceinline__ bool rho_below(const MassTableS &mt, uint32_t k, float r2, float thr)
{
    const int s = rho_screen(mt, k, r2, thr);
    return s == 2 ? so_rho_below(mt_eval(mt, k), r2, thr) != 0 : s == 1;
}

/* ============================================================================================
 * geometry: exact r^2 and the row segments that overlap a ball
 * ============================================================================================ */
__device__ __forceinline__ float axis_delta(float x, float p, float L, float hL)
{
    float rd = __fsub_rn(x, p);                 /* which image is nearer (kd2.h:165-194)      */
    float sx = x;
    if (rd > hL) sx = __fsub_rn(x, L);          /* sx = x - lx                                */
    else if (rd < -hL) sx = __fadd_rn(x, L);    /* sx = x + lx                                */
    return __fsub_rn(sx, p);                    /* dx = sx - p.r[0]   (smooth2.c:89)          */
}

struct Center {
    float x, y, z;
};

__device__ __forceinline__ float dist2(const Center &c, const float4 &q, const GridDev &g)
{
    float dx = axis_delta(c.x, q.x, g.L[0], g.halfL[0]);
    float dy = axis_delta(c.y, q.y, g.L[1], g.halfL[1]);
    float dz = axis_delta(c.z, q.z, g.L[2], g.halfL[2]);
    /* fDist2 = dx*dx + dy*dy + dz*dz, left to right, fp32, no FMA (smooth2.c:92) */
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

struct BallGeom {
    double cx, cy, cz, b, b2;
    int ylo, ny, zlo, nz, nrows;
    int full;
};

__device__ __forceinline__ BallGeom make_geom(const GridDev &g, const Center &c, double b)
{
    BallGeom B;
    B.cx = c.x; B.cy = c.y; B.cz = c.z;
    B.b = b; B.b2 = b * b;
    B.full = !(b < g.bmax_pruned);
    if (B.full) {
        B.ylo = 0; B.ny = g.nc; B.zlo = 0; B.nz = g.nc;
    } else {
        int yhi, zhi;
        B.ylo = (int)floor((B.cy - b - g.dg0[1]) * g.dinvh[1] - CELL_MARGIN);
        yhi = (int)floor((B.cy + b - g.dg0[1]) * g.dinvh[1] + CELL_MARGIN);
        B.zlo = (int)floor((B.cz - b - g.dg0[2]) * g.dinvh[2] - CELL_MARGIN);
        zhi = (int)floor((B.cz + b - g.dg0[2]) * g.dinvh[2] + CELL_MARGIN);
        B.ny = min(yhi - B.ylo + 1, g.nc);
        B.nz = min(zhi - B.zlo + 1, g.nc);
    }
    B.nrows = B.ny * B.nz;
    return B;
}

/* distance from coordinate c to the cell interval [lo, lo+h], shrunk by the margin */
__device__ __forceinline__ double interval_dist(double c, double lo, double h)
{
    double m = CELL_MARGIN * h;
    double d = fmax(lo - c, c - (lo + h)) - m;
    return d > 0.0 ? d : 0.0;
}

/* the (up to two, because of the periodic wrap) sorted-slot ranges of row r */
__device__ __forceinline__ void row_segments(const GridDev &g, const BallGeom &B, int r,
                                             uint32_t &s0, uint32_t &l0, uint32_t &s1, uint32_t &l1)
{
    s0 = l0 = s1 = l1 = 0;
    int iyu = B.ylo + r % B.ny, izu = B.zlo + r / B.ny;
    int mask = g.nc - 1;
    int xa, nx;
    if (B.full) {
        xa = 0; nx = g.nc;
    } else {
        double dy = interval_dist(B.cy, g.dg0[1] + iyu * g.dh[1], g.dh[1]);
        double dz = interval_dist(B.cz, g.dg0[2] + izu * g.dh[2], g.dh[2]);
        double rem = B.b2 - dy * dy - dz * dz;
        if (rem < 0.0) return;
        double w = sqrt(rem);
        int xlo = (int)floor((B.cx - w - g.dg0[0]) * g.dinvh[0] - CELL_MARGIN);
        int xhi = (int)floor((B.cx + w - g.dg0[0]) * g.dinvh[0] + CELL_MARGIN);
        nx = min(xhi - xlo + 1, g.nc);
        xa = xlo & mask;
    }
    uint32_t rowbase = row_key((uint32_t)(iyu & mask), (uint32_t)(izu & mask), g.lb, g.tb) << g.lb;
    int n0 = min(nx, g.nc - xa);
    uint32_t a = __ldg(g.ce + rowbase + xa), e = __ldg(g.ce + rowbase + xa + n0);
    s0 = a; l0 = e - a;
    if (n0 < nx) {
        uint32_t a1 = __ldg(g.ce + rowbase), e1 = __ldg(g.ce + rowbase + (nx - n0));
        s1 = a1; l1 = e1 - a1;
    }
}

/* ---- TMA bulk copy + mbarrier (PTX ISA 8.0, sm_90+) ---------------------------------------------- */
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smem_addr(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_addr(dst)), "l"(src), "r"(bytes), "r"(smem_addr(bar)) : "memory");
}
template <int NT> __device__ __forceinline__ void tma_stage_init(GroupSmem<NT> &sm, int tid, int use_tma)
{
    if constexpr (NT == 1024) {
        if (!use_tma) return;                   /* the stage memory is not even allocated then */
        if (tid == 0) {
            for (int k = 0; k < TMA_STAGES; ++k) {
                mbar_init(&sm.tma.full[k], 1u);
                mbar_init(&sm.tma.empty[k], (uint32_t)(NT / 32));      /* one arrival per warp */
            }
            sm.tma.tiles_done = 0u;
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
}

/* Visit every particle stored in a cell that overlaps the ball: f(slot, particle).
 * All NT threads of the group must call this together. */
template <int NT, typename F>
__device__ __forceinline__ void for_each_in_ball(const GridDev &g, GroupSmem<NT> &sm, int tid,
                                                 const BallGeom &B, F &f, uint32_t &evals)
{
    for (int rb = 0; rb < B.nrows; rb += NT) {
        int r = rb + tid;
        uint32_t s0 = 0, l0 = 0, s1 = 0, l1 = 0;
        if (r < B.nrows) row_segments(g, B, r, s0, l0, s1, l1);
        uint32_t incl = gscan_incl<NT>(l0 + l1, sm.tmp, tid);
        sm.seg_start[2 * tid] = s0;
        sm.seg_pre[2 * tid] = incl - l1;
        sm.seg_start[2 * tid + 1] = s1;
        sm.seg_pre[2 * tid + 1] = incl;
        gsync<NT>();
        uint32_t total = sm.seg_pre[2 * NT - 1];
        if (NT == 1024 && g.use_tma) if constexpr (NT == 1024) {
            /* TMA path: tile k = flat range [k*TMA_TILE, (k+1)*TMA_TILE) of the concatenated row segments,
             * fetched by warp 0 (one bulk copy per piece of a segment, lanes in parallel) TMA_STAGES-1 tiles
             * ahead of the tile the block reads.  Tiles are numbered J = 0, 1, 2 ... over the CTA's lifetime:
             * stage J % STAGES, full-barrier phase (J / STAGES) & 1; a stage is refilled once all warps
             * have arrived on its empty barrier for tile J - STAGES. */
            const int lane = tid & 31, w = tid >> 5;
            const uint32_t J0 = sm.tma.tiles_done;
            const uint32_t ntile = (total + TMA_TILE - 1) / TMA_TILE;
            auto issue = [&](uint32_t k) {
                const uint32_t J = J0 + k, st = J % TMA_STAGES, t0 = k * TMA_TILE, t1 = min(total, t0 + TMA_TILE);
                if (J >= TMA_STAGES) mbar_wait(&sm.tma.empty[st], ((J / TMA_STAGES) - 1u) & 1u);
                if (lane == 0) mbar_expect_tx(&sm.tma.full[st], (t1 - t0) * (uint32_t)sizeof(float4));
                __syncwarp();
                int lo = 0, hi = 2 * NT - 1;                  /* first segment whose inclusive prefix > t0 */
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (sm.seg_pre[mid] > t0) hi = mid; else lo = mid + 1;
                }
                for (int sgm = lo + lane; sgm < 2 * NT; sgm += 32) {
                    const uint32_t beg = sgm ? sm.seg_pre[sgm - 1] : 0u, end = sm.seg_pre[sgm];
                    if (beg >= t1) break;
                    const uint32_t a = max(beg, t0), b = min(end, t1);
                    if (b > a)
                        tma_bulk_g2s(&sm.tma.tile[st][a - t0], g.sorted + sm.seg_start[sgm] + (a - beg),
                                     (b - a) * (uint32_t)sizeof(float4), &sm.tma.full[st]);
                }
            };
            if (w == 0)
                for (uint32_t k = 0; k < ntile && k < TMA_STAGES - 1; ++k) issue(k);
            for (uint32_t k = 0; k < ntile; ++k) {
                if (w == 0 && k + TMA_STAGES - 1 < ntile) issue(k + TMA_STAGES - 1);
                const uint32_t J = J0 + k, st = J % TMA_STAGES;
                mbar_wait(&sm.tma.full[st], (J / TMA_STAGES) & 1u);
                const uint32_t cnt = min((uint32_t)TMA_TILE, total - k * TMA_TILE);
                for (uint32_t i = tid; i < cnt; i += NT) {
                    const float4 q = sm.tma.tile[st][i];
                    f(0u, q);
                    ++evals;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   /* my reads come before the next bulk write */
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.tma.empty[st]);
            }
            gsync<NT>();
            if (tid == 0) sm.tma.tiles_done = J0 + ntile;
            gsync<NT>();
            continue;
        }
        uint32_t i = tid;
        if (i < total) {
            int lo = 0, hi = 2 * NT - 1;
            while (lo < hi) {                   /* first segment whose inclusive prefix > i   */
                int mid = (lo + hi) >> 1;
                if (sm.seg_pre[mid] > i) hi = mid; else lo = mid + 1;
            }
            int s = lo;
            const int U = 4;                    /* independent loads in flight per thread */
            for (; i < total; i += NT * U) {
                uint32_t pp[U];
                float4 qq[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t iu = i + (uint32_t)u * NT;
                    pp[u] = 0xFFFFFFFFu;
                    if (iu < total) {
                        while (sm.seg_pre[s] <= iu) ++s;
                        uint32_t beg = s ? sm.seg_pre[s - 1] : 0u;
                        pp[u] = sm.seg_start[s] + (iu - beg);
                    }
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (pp[u] != 0xFFFFFFFFu) qq[u] = __ldg(g.sorted + pp[u]);
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (pp[u] != 0xFFFFFFFFu) {
                        f(pp[u], qq[u]);
                        ++evals;
                    }
            }
        }
        gsync<NT>();
    }
}

/* ============================================================================================
 * functors of the passes
 * ============================================================================================ */
/* append slot for the threads that reach this point together: ONE atomic per converged group of lanes */
__device__ __forceinline__ uint32_t agg_append(uint32_t *cnt)
{
    const unsigned m = __activemask();
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(cnt, (uint32_t)__popc(m));
    base = __shfl_sync(m, base, leader);
    return base + (uint32_t)__popc(m & ((1u << lane) - 1u));
}

/* The passes over a ball come in two forms: operator()(slot, particle) for a traversal of the grid, and
 * take(r^2 bits, index) for a walk over the keys the histogram pass staged in shared memory. */
struct HistF {          /* count particles with bits(r^2) in [lo,hi] into level bins */
    const GridDev *g;
    Center c;
    uint32_t lo_bits, hi_bits, shift, base;
    uint32_t *hist;
    uint32_t first_bits, n_first;   /* side count: particles inside the FIRST ball of the schedule */
    unsigned long long *skey;       /* staging of every key of the ball (scap = 0: off)             */
    uint32_t *scnt, scap;
    __device__ __forceinline__ void take(uint32_t bits, uint32_t idx)
    {
        if (bits >= lo_bits && bits <= hi_bits) {   /* NaN (> 0x7f800000) never passes: hi is finite */
            uint32_t hb = bits >> shift;
            atomicAdd(&hist[hb > base ? hb - base : 0u], 1u);
            n_first += (bits <= first_bits);
            if (scap) {
                const uint32_t pos = agg_append(scnt);
                if (pos < scap) skey[pos] = ((unsigned long long)bits << 32) | idx;
            }
        }
    }
    __device__ __forceinline__ void operator()(uint32_t, const float4 &q)
    {
        take(__float_as_uint(dist2(c, q, *g)), __float_as_uint(q.w));
    }
};

template <int CAP> struct CollectF {   /* append (r^2 bits, original index) of the window */
    const GridDev *g;
    Center c;
    uint32_t lo_bits, hi_bits;
    unsigned long long *wkey;
    uint32_t *cnt;
    __device__ __forceinline__ void take(uint32_t bits, uint32_t idx)
    {
        if (bits >= lo_bits && bits <= hi_bits) {
            const uint32_t pos = agg_append(cnt);
            if (pos < (uint32_t)CAP) wkey[pos] = ((unsigned long long)bits << 32) | idx;
        }
    }
    __device__ __forceinline__ void operator()(uint32_t p, const float4 &q)
    {
        take(__float_as_uint(dist2(c, q, *g)), __float_as_uint(q.w));
    }
};

struct EmitF {          /* members: (r^2 bits, index) < key_j */
    const GridDev *g;
    Center c;
    unsigned long long key_j;
    int32_t *members;
    float *md2;
    uint32_t *cnt;
    uint32_t limit;
    __device__ __forceinline__ void take(uint32_t bits, uint32_t idx)
    {
        if ((((unsigned long long)bits << 32) | idx) < key_j) {
            const uint32_t pos = agg_append(cnt);
            if (pos < limit) {
                members[pos] = (int32_t)idx;
                if (md2) md2[pos] = __uint_as_float(bits);
            }
        }
    }
    __device__ __forceinline__ void operator()(uint32_t p, const float4 &q)
    {
        take(__float_as_uint(dist2(c, q, *g)), __float_as_uint(q.w));
    }
};

/* walk over the keys staged by the histogram pass (all NT threads of the group) */
template <int NT, typename F>
__device__ __forceinline__ void for_each_staged(const unsigned long long *skey, uint32_t n, int tid, F &f)
{
    for (uint32_t i = tid; i < n; i += NT) {
        const unsigned long long k = skey[i];
        f.take((uint32_t)(k >> 32), (uint32_t)k);
    }
}

struct GatherF {        /* sogpu_ball_gather: everything with r^2 <= ball2 */
    const GridDev *g;
    Center c;
    uint32_t hi_bits;
    int32_t *idx;
    float *d2o;
    unsigned long long *cnt;
    unsigned long long cap;
    __device__ __forceinline__ void operator()(uint32_t p, const float4 &q)
    {
        float d2 = dist2(c, q, *g);
        if (__float_as_uint(d2) <= hi_bits) {
            unsigned long long pos = atomicAdd(cnt, 1ull);
            if (pos < cap) {
                idx[pos] = __float_as_int(q.w);
                d2o[pos] = d2;
            }
        }
    }
};

/* ============================================================================================
 * histogram levels
 * ============================================================================================ */
struct Level {
    uint32_t shift, base;     /* bin = max(0, (bits >> shift) - base)                        */
    uint32_t lo_bits, hi_bits;/* bit range this level covers                                  */
    uint32_t rank0;           /* sorted rank of the first particle of the range               */
    int clamp;                /* bin 0 collects everything below (base+1) << shift            */
    int cur;                  /* next bin to examine                                          */
};

__device__ __forceinline__ uint32_t bin_lo(const Level &lv, int b)
{
    if (b == 0 && lv.clamp) return lv.lo_bits;
    unsigned long long v = (unsigned long long)(lv.base + (uint32_t)b) << lv.shift;
    return v < lv.lo_bits ? lv.lo_bits : (uint32_t)v;
}
__device__ __forceinline__ uint32_t bin_hi(const Level &lv, int b)
{
    unsigned long long v = ((unsigned long long)(lv.base + (uint32_t)b + 1u) << lv.shift) - 1ull;
    return v > lv.hi_bits ? lv.hi_bits : (uint32_t)v;
}

/* counts -> exclusive prefix in place; hist[NB] = total.  Returns the total. */
template <int NT> __device__ __forceinline__ uint32_t scan_hist(uint32_t *hist, uint32_t *tmp, int tid)
{
    const int BPT = (NB + NT - 1) / NT;        /* bins per thread (threads beyond NB/BPT idle) */
    uint32_t v[BPT], s = 0;
#pragma unroll
    for (int k = 0; k < BPT; ++k) {
        int b = tid * BPT + k;
        v[k] = (b < NB) ? hist[b] : 0u;
        s += v[k];
    }
    uint32_t incl = gscan_incl<NT>(s, tmp, tid);
    uint32_t run = incl - s;
#pragma unroll
    for (int k = 0; k < BPT; ++k) {
        int b = tid * BPT + k;
        if (b < NB) hist[b] = run;
        run += v[k];
    }
    if (tid == NT - 1) hist[NB] = incl;
    gsync<NT>();
    return hist[NB];
}

/* first bin >= from that may contain a particle below the threshold (rank >= jmin) */
template <int NT>
__device__ __forceinline__ int find_candidate(const uint32_t *cum, const Level &lv, int from,
                                              uint32_t jmin, float thr, const MassTableS &mt,
                                              uint32_t *tmp, int tid)
{
    int best = NB;
    for (int b = from + tid; b < NB; b += NT) {
        uint32_t c0 = cum[b], c1 = cum[b + 1];
        if (c1 == c0) continue;
        uint32_t r0 = lv.rank0 + c0, r1 = lv.rank0 + c1;
        if (r1 <= jmin) continue;
        uint32_t klo = r0 > jmin ? r0 : jmin;
        float r2hi = __uint_as_float(bin_hi(lv, b));
        if (rho_screen(mt, klo + 1u, r2hi, thr) == 0) continue;      /* even the densest case of this bin is clearly above */
        float mass_lo = mt_eval(mt, klo + 1u);
        if (!so_surely_not_below(mass_lo, r2hi, thr)) { best = b; break; }
    }
    return gmin<NT>(best, tmp, tid);
}

/* Bitonic sort of P (a power of two) keys in shared memory.  CTA groups: every warp owns a contiguous chunk of
 * C = P / warps >= 64 keys, and the P/2 compare-exchanges of a stage are dealt so that a stage whose stride j is
 * below C stays inside the warps' own chunks — those stages need __syncwarp only.  Block-wide barriers remain
 * around the few stages with j >= C (9 instead of 55 for P = 1024 and 8 warps). */
template <int NT> __device__ __forceinline__ void bitonic_sort(unsigned long long *key, int P, int tid)
{
    if (NT == 32) {
        for (int k = 2; k <= P; k <<= 1) {
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P; i += NT) {
                    int ixj = i ^ j;
                    if (ixj > i) {
                        unsigned long long a = key[i], b = key[ixj];
                        bool up = ((i & k) == 0);
                        if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                    }
                }
                __syncwarp();
            }
        }
        return;
    }
    int nw = P / 64;
    if (nw > NT / 32) nw = NT / 32;
    if (nw < 1) nw = 1;
    const int C = P / nw, halfC = C >> 1;
    const int w = tid >> 5, lane = tid & 31;
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (w < nw) {
                for (int m = lane; m < halfC; m += 32) {
                    const int pr = w * halfC + m;
                    const int i = ((pr & ~(j - 1)) << 1) | (pr & (j - 1));       /* a zero bit inserted at log2(j) */
                    const int ixj = i | j;
                    unsigned long long a = key[i], b = key[ixj];
                    const bool up = ((i & k) == 0);
                    if ((a > b) == up) { key[i] = b; key[ixj] = a; }
                }
            }
            /* what the NEXT stage reads decides the barrier: chunk-local strides only need the warp's own writes */
            const int nk = (j > 1) ? k : (k << 1), nj = (j > 1) ? (j >> 1) : (nk >> 1);
            const bool next_cross = (nk > P) || nj >= C;
            if (j >= C || next_cross) __syncthreads(); else __syncwarp();
        }
    }
}

#include "seg_sort.cuh"

/* ============================================================================================
 * focused grid: only the coarse cells some halo can ever look at are kept by the build
 * ============================================================================================ */
/* conservative range of mask cells [m0, m0+cnt) per axis touched by the cube of half-width b */
__device__ __forceinline__ void mask_range(const GridDev &g, int axis, double c, double b, int &m0, int &cnt)
{
    const int nm = 1 << g.mb;
    if (!(b < g.bmax_pruned)) { m0 = 0; cnt = nm; return; }
    int lo = (int)floor((c - b - g.dg0[axis]) * g.dinvh[axis] - CELL_MARGIN) - 1;
    int hi = (int)floor((c + b - g.dg0[axis]) * g.dinvh[axis] + CELL_MARGIN) + 1;
    int mlo = lo >> g.ms, mhi = hi >> g.ms;          /* arithmetic shifts: floor division */
    cnt = min(mhi - mlo + 1, nm);
    m0 = mlo;
}

template <int NT>
__device__ __forceinline__ bool ball_covered(const GridDev &g, GroupSmem<NT> &sm, int tid, const Center &c, double b)
{
    if (!g.mask) return true;
    int x0, nx, y0, ny, z0, nz;
    mask_range(g, 0, c.x, b, x0, nx);
    mask_range(g, 1, c.y, b, y0, ny);
    mask_range(g, 2, c.z, b, z0, nz);
    const int nm1 = (1 << g.mb) - 1, tot = nx * ny * nz;
    int ok = 1;
    for (int i = tid; i < tot && ok; i += NT) {
        int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
        ok = mask_bit(g, (uint32_t)((x0 + ix) & nm1), (uint32_t)((y0 + iy) & nm1), (uint32_t)((z0 + iz) & nm1));
    }
    return gmin<NT>(ok, sm.tmp, tid) != 0;
}

/* one warp per halo marks the cube it can reach after n_balls steps of the ball schedule */
__global__ void __launch_bounds__(256) k_mark_mask(GridDev g, const float *__restrict__ centers,
                                                   const float *__restrict__ rgtp, int nh, int n_balls,
                                                   uint32_t *__restrict__ mask)
{
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    const float root = so_root_period(g.L[0], g.L[1], g.L[2]);
    const int nm1 = (1 << g.mb) - 1;
    for (int h = wid; h < nh; h += nw) {
        float ball = rgtp[h];
        for (int k = 0; k < n_balls && (double)ball < 0.25 * (double)root; ++k) ball = so_next_ball(ball);
        const float ball2 = __fmul_rn(ball, ball);
        /* small halos: at least mask_rmin around the centre (a fraction of a coarse cell costs next to
         * nothing and keeps poorly estimated small groups inside their mask) */
        const double b = fmax(sqrt((double)ball2) * (1.0 + 1.0e-6), g.mask_rmin);
        int x0, nx, y0, ny, z0, nz;
        mask_range(g, 0, centers[3 * h + 0], b, x0, nx);
        mask_range(g, 1, centers[3 * h + 1], b, y0, ny);
        mask_range(g, 2, centers[3 * h + 2], b, z0, nz);
        const int tot = nx * ny * nz;
        for (int i = lane; i < tot; i += 32) {
            int ix = i % nx, iy = (i / nx) % ny, iz = i / (nx * ny);
            uint32_t bit = ((uint32_t)((z0 + iz) & nm1) << (2 * g.mb)) | ((uint32_t)((y0 + iy) & nm1) << g.mb) |
                           (uint32_t)((x0 + ix) & nm1);
            atomicOr(&mask[bit >> 5], 1u << (bit & 31));
        }
    }
}

/* ============================================================================================
 * one halo: kdRvir (kd2.c:723-840)
 * ============================================================================================ */
struct HaloResult {
    int32_t n;                 /* N_Delta, or -1/-2/-3, CODE_UNSUPPORTED, CODE_DEFER           */
    float m;                   /* M_Delta                                                      */
    unsigned long long key_j;  /* (r^2 bits, index) of sorted element j = first non-member     */
    unsigned long long off;    /* where the member list starts in the member buffer            */
};

struct EmitCtx {               /* member emission from inside the solver */
    int32_t *members;
    float *md2;
    unsigned long long cap;
    unsigned long long *cursor;    /* global append cursor of the member buffer: one atomicAdd(N_Delta) per halo */
    uint32_t *flags;               /* bit0 member buffer too small, bit1 emit count mismatch */
    int stage;                     /* stage the ball's keys in shared memory (off with TMA staging: they share it) */
};

/* The member list of a solved halo: every particle with (r^2 bits, index) < key_j, written behind a slot of j
 * entries reserved with one atomic on the buffer's cursor.  From the staged keys when the ball was staged
 * (no global traversal at all), else by one more traversal of the ball of radius r_j (an L2 hit: the solver
 * has just read it). */
template <int NT>
__device__ __forceinline__ void emit_members_here(const GridDev &g, GroupSmem<NT> &sm, int tid, const Center &c,
                                                  const EmitCtx &ec, HaloResult &res, bool staged, uint32_t n_staged,
                                                  uint32_t &ev_other)
{
    const uint32_t j = (uint32_t)res.n;
    if (tid == 0) { sm.bcast64 = atomicAdd(ec.cursor, (unsigned long long)j); sm.cnt = 0u; }
    gsync<NT>();
    const unsigned long long off = sm.bcast64;
    res.off = off;
    if (off + j > ec.cap) {
        if (tid == 0) atomicOr(ec.flags, 1u);
        gsync<NT>();
        return;
    }
    EmitF f;
    f.g = &g; f.c = c; f.key_j = res.key_j;
    f.members = ec.members + off; f.md2 = ec.md2 ? ec.md2 + off : nullptr;
    f.cnt = &sm.cnt; f.limit = j;
    if (staged) {
        for_each_staged<NT>(sm.skey, n_staged, tid, f);
    } else {
        const float r2j = __uint_as_float((uint32_t)(res.key_j >> 32));
        BallGeom B = make_geom(g, c, sqrt((double)r2j) * (1.0 + 1.0e-6));
        for_each_in_ball<NT>(g, sm, tid, B, f, ev_other);
    }
    gsync<NT>();
    if (tid == 0 && sm.cnt != j) atomicOr(ec.flags, 2u);
    gsync<NT>();
}

template <int NT>
__device__ void so_halo(const GridDev &g, const MassTableS &mt, GroupSmem<NT> &sm, int tid,
                        Center c, float rgtp, float thr, int nM, int first_ball, HaloResult &res,
                        uint32_t &ev_hist, uint32_t &ev_other, const EmitCtx &ec)
{
    typedef Cfg<NT> CF;
    const float root = so_root_period(g.L[0], g.L[1], g.L[2]);           /* kd2.c:765 */
    float ball = rgtp;                                                   /* kd2.c:745 */
    uint32_t n_prev = 0;
    int kball = 0;
    /* The reference gathers ball after ball of its schedule b_k = 1.2 b_(k-1) until a pair fires
     * inside one (kd2.c:765-836).  Which ball that is does not change the answer: ball k examines
     * exactly the pairs (j, j+1) with j+1 inside it that smaller balls have not examined, so the
     * result is the first firing pair of the sorted list, found in the first ball that holds it.
     * Any SUBSET of the schedule that ends with the same last ball therefore gives the same result
     * (and the same -3), provided the -1 test still counts the FIRST ball (done on the side below).
     * `steps` = how many schedule steps to advance before the next gather. */
    int steps = first_ball;
    res.n = -3; res.m = -3.0f; res.key_j = 0ull; res.off = 0ull;         /* kd2.c:837-838 */

    while ((double)ball < 0.25 * (double)root) {                         /* kd2.c:766 */
        ball = so_next_ball(ball);                                       /* kd2.c:767 */
        const float ball_k1 = ball;                                      /* (kball == 0: the schedule's first ball) */
        for (int sk = 1; sk < steps && (double)ball < 0.25 * (double)root; ++sk) ball = so_next_ball(ball);
        const float ball2 = __fmul_rn(ball, ball);                       /* kd2.c:768 */
        if (!(ball2 < INFINITY) || !(ball > 0.0f)) break;
        const uint32_t ball_bits = __float_as_uint(ball2);

        /* ---- pass A: histogram of the whole ball (smBallGather + the sort's first digit) ---- */
        Level lev[CF::NLEV];
        {
            uint32_t top = ball_bits >> SHIFT0;
            lev[0].shift = SHIFT0;
            lev[0].base = top > (NB - 1) ? top - (NB - 1) : 0u;
            lev[0].lo_bits = 0u; lev[0].hi_bits = ball_bits;
            lev[0].rank0 = 0u; lev[0].clamp = 1; lev[0].cur = 0;
        }
        for (int b = tid; b <= NB; b += NT) sm.hist[0][b] = 0u;
        if (tid == 0) sm.scnt = 0u;
        gsync<NT>();
        if (!ball_covered<NT>(g, sm, tid, c, sqrt((double)ball2) * (1.0 + 1.0e-6))) {
            res.n = CODE_NEED_FULL; res.m = 0.0f;      /* focused grid too small for this ball */
            return;
        }
        BallGeom B = make_geom(g, c, sqrt((double)ball2) * (1.0 + 1.0e-6));
        uint32_t n_first;
        {
            HistF f;
            f.g = &g; f.c = c; f.lo_bits = 0u; f.hi_bits = ball_bits;
            f.shift = SHIFT0; f.base = lev[0].base; f.hist = sm.hist[0];
            f.first_bits = __float_as_uint(__fmul_rn(ball_k1, ball_k1)); f.n_first = 0u;
            f.skey = sm.skey; f.scnt = &sm.scnt; f.scap = ec.stage ? (uint32_t)CF::SCAP : 0u;
            for_each_in_ball<NT>(g, sm, tid, B, f, ev_hist);
            n_first = f.n_first;
        }
        const uint32_t n = scan_hist<NT>(sm.hist[0], sm.tmp, tid);       /* nParticles, kd2.c:769 */
        /* the whole ball sits in shared memory: every later pass of this ball reads it there */
        const bool staged = ec.stage && n <= (uint32_t)CF::SCAP;

        if (kball == 0) {                                                /* kd2.c:772-778 */
            if (ball != ball_k1) n_first = gsum<NT>(n_first, sm.tmp, tid); else n_first = n;
            if (n_first < (uint32_t)nM) {
                res.n = -1; res.m = -1.0f;
                return;
            }
        }
        /* pairs (j, j+1) already examined in smaller balls: j <= n_prev-2 (kd2.c:804,832) */
        const uint32_t jmin = (kball == 0) ? (uint32_t)(nM - 2) : n_prev - 1u;

        /* ---- search: first j >= jmin with below(j) && below(j+1), j+1 < n ------------------- */
        if (n > jmin + 1u) {
            int L = 0;
            bool carry = false;                 /* below(rank just before the next window)   */
            unsigned long long prev_key = 0ull;
            {   /* start at the bin that holds rank jmin */
                int lo = 0, hi = NB - 1;
                while (lo < hi) {
                    int mid = (lo + hi) >> 1;
                    if (sm.hist[0][mid + 1] > jmin) hi = mid; else lo = mid + 1;
                }
                lev[0].cur = lo;
            }
            for (;;) {
                Level &lv = lev[L];
                const uint32_t *cum = sm.hist[L];
                int cb = (lv.cur < NB) ? find_candidate<NT>(cum, lv, lv.cur, jmin, thr, mt, sm.tmp, tid) : NB;
                if (cb >= NB) {
                    if (lv.cur < NB && cum[NB] > cum[lv.cur]) carry = false;
                    if (L == 0) break;                       /* nothing fires in this ball    */
                    --L;
                    lev[L].cur += 1;
                    continue;
                }
                if (cum[cb] > cum[lv.cur]) carry = false;    /* skipped particles: not below  */
                const uint32_t cnt_c = cum[cb + 1] - cum[cb];
                if (cnt_c > (uint32_t)CF::CAP) {
                    /* ---- refine: histogram the candidate bin with finer bins --------------- */
                    if (L + 1 >= CF::NLEV) {
                        res.n = (CF::NLEV == 1) ? CODE_DEFER : CODE_UNSUPPORTED;
                        res.m = 0.0f;
                        return;
                    }
                    Level &nl = lev[L + 1];
                    lv.cur = cb;                 /* resume after this bin when the child is done */
                    nl.lo_bits = bin_lo(lv, cb);
                    nl.hi_bits = bin_hi(lv, cb);
                    nl.rank0 = lv.rank0 + cum[cb];
                    nl.cur = 0;
                    if (cb == 0 && lv.clamp && lv.base > 0u) {
                        uint32_t top = nl.hi_bits >> lv.shift;        /* == lv.base            */
                        nl.shift = lv.shift;
                        nl.base = top > (NB - 1) ? top - (NB - 1) : 0u;
                        nl.clamp = 1;
                    } else {
                        if (lv.shift == 0u) {     /* > CAP particles at one identical r^2      */
                            res.n = CODE_UNSUPPORTED; res.m = 0.0f;
                            return;
                        }
                        nl.shift = lv.shift > NB_LOG ? lv.shift - NB_LOG : 0u;
                        nl.base = nl.lo_bits >> nl.shift;
                        nl.clamp = 0;
                    }
                    ++L;
                    for (int b = tid; b <= NB; b += NT) sm.hist[L][b] = 0u;
                    gsync<NT>();
                    HistF f;
                    f.g = &g; f.c = c; f.lo_bits = nl.lo_bits; f.hi_bits = nl.hi_bits;
                    f.shift = nl.shift; f.base = nl.base; f.hist = sm.hist[L];
                    f.first_bits = 0u; f.n_first = 0u;
                    f.skey = nullptr; f.scnt = nullptr; f.scap = 0u;
                    if (staged) {
                        for_each_staged<NT>(sm.skey, n, tid, f);
                        gsync<NT>();
                    } else {
                        BallGeom B2 = make_geom(g, c, sqrt((double)__uint_as_float(nl.hi_bits)) * (1.0 + 1.0e-6));
                        for_each_in_ball<NT>(g, sm, tid, B2, f, ev_other);
                    }
                    scan_hist<NT>(sm.hist[L], sm.tmp, tid);
                    continue;
                }
                /* ---- window [cb, c2]: as many following bins as fit the target ---------------- */
                int c2;
                {
                    uint32_t room = cnt_c > (uint32_t)CF::WTARGET ? cnt_c : (uint32_t)CF::WTARGET;
                    uint32_t lim = cum[cb] + room;
                    int lo = cb + 1, hi = NB;             /* largest e in [cb+1, NB] with cum[e] <= lim */
                    while (lo < hi) {
                        int mid = (lo + hi + 1) >> 1;
                        if (cum[mid] <= lim) lo = mid; else hi = mid - 1;
                    }
                    c2 = lo - 1;
                }
                const uint32_t wn = cum[c2 + 1] - cum[cb];
                const uint32_t w_lo = bin_lo(lv, cb), w_hi = bin_hi(lv, c2);
                const uint32_t rank_first = lv.rank0 + cum[cb];
                if (tid == 0) sm.cnt = 0u;
                int P = 2;
                while (P < (int)wn) P <<= 1;
                for (int i = tid; i < P; i += NT) sm.wkey[i] = ~0ull;
                gsync<NT>();
                {
                    CollectF<CF::CAP> f;
                    f.g = &g; f.c = c; f.lo_bits = w_lo; f.hi_bits = w_hi;
                    f.wkey = sm.wkey; f.cnt = &sm.cnt;
                    if (staged) {
                        for_each_staged<NT>(sm.skey, n, tid, f);
                        gsync<NT>();
                    } else {
                        BallGeom B2 = make_geom(g, c, sqrt((double)__uint_as_float(w_hi)) * (1.0 + 1.0e-6));
                        for_each_in_ball<NT>(g, sm, tid, B2, f, ev_other);
                    }
                }
                if (sm.cnt != wn) {              /* cannot happen; guards the exactness claim  */
                    res.n = CODE_UNSUPPORTED; res.m = 1.0f;
                    return;
                }
                bitonic_sort<NT>(sm.wkey, P, tid);
                /* below(k) = rho(S[k+1], r2[k]) < thr, k = rank (kd2.c:791-792, 814-815) */
                for (uint32_t i = tid; i < wn; i += NT) {
                    uint32_t k = rank_first + i;
                    uint8_t fl = 0;
                    if (k >= jmin) fl = (uint8_t)rho_below(mt, k + 1u, __uint_as_float((uint32_t)(sm.wkey[i] >> 32)), thr);
                    sm.wflag[i] = fl;
                }
                gsync<NT>();
                int fire = INT_MAX;
                for (uint32_t i = tid; i < wn; i += NT) {
                    if (sm.wflag[i] && (i > 0 ? sm.wflag[i - 1] != 0 : carry)) { fire = (int)i; break; }
                }
                fire = gmin<NT>(fire, sm.tmp, tid);
                if (fire != INT_MAX) {
                    /* element `fire` is j+1, element fire-1 is j (kd2.c:814-823) */
                    uint32_t j = rank_first + (uint32_t)fire - 1u;
                    unsigned long long key_j = fire > 0 ? sm.wkey[fire - 1] : prev_key;
                    if (kball == 0 && j == (uint32_t)(nM - 2)) {             /* kd2.c:791-796 */
                        res.n = -2; res.m = -2.0f;
                        return;
                    }
                    float mass = mt_eval(mt, j + 1u);                         /* kd2.c:807 */
                    res.m = __fsub_rn(mass, mt.m);                            /* kd2.c:816 */
                    res.n = (int32_t)j;
                    res.key_j = key_j;
                    gsync<NT>();
                    if (ec.members) emit_members_here<NT>(g, sm, tid, c, ec, res, staged, n, ev_other);   /* kd2.c:823 */
                    return;
                }
                carry = sm.wflag[wn - 1] != 0;
                prev_key = sm.wkey[wn - 1];
                gsync<NT>();
                lv.cur = c2 + 1;
            }
        }
        n_prev = n;                                                      /* kd2.c:832 */
        ++kball;
        /* nothing fired: jump towards the radius where an isothermal profile (mean density ~ r^-2)
         * through the density at this ball's edge would cross the threshold */
        steps = 1;
        if (n > 0u) {
            const float b = sqrtf(ball2);
            const float ratio = mt_eval(mt, n) / (4.18879f * b * b * b * thr);
            if (ratio > 1.0f) steps = min(8, max(1, (int)rintf(0.5f * __log2f(ratio) * 3.8018f)));   /* / log2(1.2) */
        }
    }
}

/* ============================================================================================
 * the persistent query / emit kernels
 * ============================================================================================ */
#define CODE_UNEQUAL_MASS (-102)

struct QueryArgs {
    GridDev g;
    const float *centers;      /* nh x 3 */
    const float *rgtp;         /* nh */
    const int32_t *list;       /* halo ids this kernel processes */
    const uint32_t *list_n;    /* how many */
    uint32_t *work_counter;    /* dynamic scheduling */
    const int32_t *list2;      /* (fused kernel) the warp-per-halo list, processed after `list` */
    const uint32_t *list2_n;
    uint32_t *work_counter2;
    int32_t *defer_list;       /* (warp kernel) halos handed to the block kernel */
    uint32_t *defer_n;
    float thr;
    int nM;
    int first_ball;            /* schedule index (1-based) of the first ball gathered */
    int32_t *out_n;            /* N_Delta or a negative code */
    float *out_m;              /* M_Delta */
    unsigned long long *out_key;   /* (r^2 bits, index) of sorted element j */
    unsigned long long *out_off;   /* first member slot per halo: written by the query (emit_in_query) or read by k_so_emit */
    unsigned long long *member_cursor;   /* emit_in_query: append cursor of the member buffer */
    int emit_in_query;
    int32_t *members;
    float *md2;
    unsigned long long member_cap;
    unsigned long long *evals;     /* [0] histogram pass, [1] other passes */
    uint32_t *flags;               /* bit0 member buffer too small, bit1 emit count mismatch */
    const so_mass_table *mt;
    unsigned long long *timeline;  /* debug (SOGPU_DEBUG_TIMELINE): [2*slot] first CTA start, [2*slot+1] last CTA end, in ns */
    int tl_slot;
};

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

template <int NT>
__device__ __forceinline__ uint32_t next_item(uint32_t *counter, GroupSmem<NT> &sm, int tid)
{
    if (tid == 0) sm.bcast[0] = atomicAdd(counter, 1u);
    gsync<NT>();
    uint32_t item = sm.bcast[0];
    gsync<NT>();
    return item;
}

template <int NT>
__global__ void __launch_bounds__(NT *Cfg<NT>::GROUPS, Cfg<NT>::MINB) k_so_query(const __grid_constant__ QueryArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MassTableS &mt = *reinterpret_cast<MassTableS *>(smem_raw);
    const size_t mt_bytes = (sizeof(MassTableS) + 127) & ~(size_t)127;
    const int grp = threadIdx.x / NT, tid = threadIdx.x % NT;
    GroupSmem<NT> &sm = *reinterpret_cast<GroupSmem<NT> *>(
        smem_raw + mt_bytes + (size_t)grp * ((sizeof(GroupSmem<NT>) + 127) & ~(size_t)127));

    /* CTA-wide: copy the mass table (built on the device by k_mass_table) */
    const int mtn = a.mt->n;
    if (mtn > 0) {
        if (threadIdx.x == 0) { mt.n = mtn; mt.m = a.mt->m; }
        for (int i = threadIdx.x; i <= mtn; i += blockDim.x) mt.k0[i] = a.mt->k0[i];
        for (int i = threadIdx.x; i < mtn; i += blockDim.x) { mt.s0[i] = a.mt->s0[i]; mt.inc[i] = a.mt->inc[i]; }
    }
    __syncthreads();

    tma_stage_init<NT>(sm, tid, a.g.use_tma);
    if (a.timeline && threadIdx.x == 0) atomicMin(&a.timeline[2 * a.tl_slot], global_ns());
    const uint32_t nlist = *a.list_n;
    uint32_t ev_hist = 0, ev_other = 0;
    for (;;) {
        const uint32_t item = next_item<NT>(a.work_counter, sm, tid);
        if (item >= nlist) break;
        const int h = a.list[item];
        HaloResult res;
        if (mtn <= 0) {                      /* unequal particle masses: not handled by this path */
            res.n = CODE_UNEQUAL_MASS; res.m = 0.0f; res.key_j = 0ull; res.off = 0ull;
        } else {
            Center c;
            c.x = a.centers[3 * h + 0]; c.y = a.centers[3 * h + 1]; c.z = a.centers[3 * h + 2];
            EmitCtx ec;
            ec.members = a.emit_in_query ? a.members : nullptr; ec.md2 = a.md2; ec.cap = a.member_cap;
            ec.cursor = a.member_cursor; ec.flags = a.flags;
            ec.stage = (NT == 1024 && a.g.use_tma) ? 0 : 1;
            so_halo<NT>(a.g, mt, sm, tid, c, a.rgtp[h], a.thr, a.nM, a.first_ball, res, ev_hist, ev_other, ec);
            gsync<NT>();
        }
        if (res.n == CODE_DEFER) {
            if (tid == 0) a.defer_list[atomicAdd(a.defer_n, 1u)] = h;
            continue;
        }
        if (tid == 0) {
            a.out_n[h] = res.n;
            a.out_m[h] = res.m;
            a.out_key[h] = res.key_j;
            if (a.emit_in_query) a.out_off[h] = res.off;
        }
    }
    if (a.timeline && threadIdx.x == 0) atomicMax(&a.timeline[2 * a.tl_slot + 1], global_ns());
    ev_hist = __reduce_add_sync(0xFFFFFFFFu, ev_hist);
    ev_other = __reduce_add_sync(0xFFFFFFFFu, ev_other);
    if ((threadIdx.x & 31) == 0) {
        if (ev_hist) atomicAdd(&a.evals[0], (unsigned long long)ev_hist);
        if (ev_other) atomicAdd(&a.evals[1], (unsigned long long)ev_other);
    }
}

/* The two common size classes in ONE persistent kernel: every CTA (256 threads) first takes mid-size halos from
 * `list` as a whole (one CTA per halo), then splits into its 8 warps, each of which takes small halos from `list2`
 * on its own.  Identical CTAs on every SM: no kernel of one class keeps the other off the SMs (with separate
 * kernels the first one resident held the shared memory until it drained), the long items are started first,
 * and the machine stays full until both lists are empty.  The shared memory is a union of the two layouts. */
#ifndef QF_MINB
#define QF_MINB 3
#endif
static size_t fused_smem_bytes()
{
    size_t mt_bytes = (sizeof(MassTableS) + 127) & ~(size_t)127;
    size_t a = sizeof(GroupSmem<256>), b = 8 * sizeof(GroupSmem<32>);
    return mt_bytes + (a > b ? a : b);
}

__global__ void __launch_bounds__(256, QF_MINB) k_so_query_fused(const __grid_constant__ QueryArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    MassTableS &mt = *reinterpret_cast<MassTableS *>(smem_raw);
    const size_t mt_bytes = (sizeof(MassTableS) + 127) & ~(size_t)127;
    unsigned char *base = smem_raw + mt_bytes;
    const int mtn = a.mt->n;
    if (mtn > 0) {
        if (threadIdx.x == 0) { mt.n = mtn; mt.m = a.mt->m; }
        for (int i = threadIdx.x; i <= mtn; i += blockDim.x) mt.k0[i] = a.mt->k0[i];
        for (int i = threadIdx.x; i < mtn; i += blockDim.x) { mt.s0[i] = a.mt->s0[i]; mt.inc[i] = a.mt->inc[i]; }
    }
    __syncthreads();
    EmitCtx ec;
    ec.members = a.emit_in_query ? a.members : nullptr; ec.md2 = a.md2; ec.cap = a.member_cap;
    ec.cursor = a.member_cursor; ec.flags = a.flags; ec.stage = 1;
    uint32_t ev_hist = 0, ev_other = 0;
    {   /* ---- CTA per halo ---- */
        GroupSmem<256> &sm = *reinterpret_cast<GroupSmem<256> *>(base);
        const int tid = threadIdx.x;
        const uint32_t nlist = *a.list_n;
        for (;;) {
            const uint32_t item = next_item<256>(a.work_counter, sm, tid);
            if (item >= nlist) break;
            const int h = a.list[item];
            HaloResult res;
            if (mtn <= 0) {
                res.n = CODE_UNEQUAL_MASS; res.m = 0.0f; res.key_j = 0ull; res.off = 0ull;
            } else {
                Center c;
                c.x = a.centers[3 * h + 0]; c.y = a.centers[3 * h + 1]; c.z = a.centers[3 * h + 2];
                so_halo<256>(a.g, mt, sm, tid, c, a.rgtp[h], a.thr, a.nM, a.first_ball, res, ev_hist, ev_other, ec);
                __syncthreads();
            }
            if (tid == 0) {
                a.out_n[h] = res.n; a.out_m[h] = res.m; a.out_key[h] = res.key_j;
                if (a.emit_in_query) a.out_off[h] = res.off;
            }
        }
    }
    __syncthreads();
    {   /* ---- warp per halo ---- */
        const int grp = threadIdx.x >> 5, tid = threadIdx.x & 31;
        GroupSmem<32> &sm = reinterpret_cast<GroupSmem<32> *>(base)[grp];
        const uint32_t nlist = *a.list2_n;
        for (;;) {
            const uint32_t item = next_item<32>(a.work_counter2, sm, tid);
            if (item >= nlist) break;
            const int h = a.list2[item];
            HaloResult res;
            if (mtn <= 0) {
                res.n = CODE_UNEQUAL_MASS; res.m = 0.0f; res.key_j = 0ull; res.off = 0ull;
            } else {
                Center c;
                c.x = a.centers[3 * h + 0]; c.y = a.centers[3 * h + 1]; c.z = a.centers[3 * h + 2];
                so_halo<32>(a.g, mt, sm, tid, c, a.rgtp[h], a.thr, a.nM, a.first_ball, res, ev_hist, ev_other, ec);
                __syncwarp();
            }
            if (res.n == CODE_DEFER) {       /* a bin too large for a warp's window: a CTA takes the halo (kernel behind this one) */
                if (tid == 0) a.defer_list[atomicAdd(a.defer_n, 1u)] = h;
                continue;
            }
            if (tid == 0) {
                a.out_n[h] = res.n; a.out_m[h] = res.m; a.out_key[h] = res.key_j;
                if (a.emit_in_query) a.out_off[h] = res.off;
            }
        }
    }
    ev_hist = __reduce_add_sync(0xFFFFFFFFu, ev_hist);
    ev_other = __reduce_add_sync(0xFFFFFFFFu, ev_other);
    if ((threadIdx.x & 31) == 0) {
        if (ev_hist) atomicAdd(&a.evals[0], (unsigned long long)ev_hist);
        if (ev_other) atomicAdd(&a.evals[1], (unsigned long long)ev_other);
    }
}

/* K5: member lists in CSR form.  Halo h owns members[out_off[h] .. out_off[h]+N_Delta). */
template <int NT>
__global__ void __launch_bounds__(NT *Cfg<NT>::GROUPS) k_so_emit(const __grid_constant__ QueryArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t mt_bytes = (sizeof(MassTableS) + 127) & ~(size_t)127;
    const int grp = threadIdx.x / NT, tid = threadIdx.x % NT;
    GroupSmem<NT> &sm = *reinterpret_cast<GroupSmem<NT> *>(
        smem_raw + mt_bytes + (size_t)grp * ((sizeof(GroupSmem<NT>) + 127) & ~(size_t)127));
    tma_stage_init<NT>(sm, tid, a.g.use_tma);
    const uint32_t nlist = *a.list_n;
    uint32_t ev = 0;
    for (;;) {
        const uint32_t item = next_item<NT>(a.work_counter, sm, tid);
        if (item >= nlist) break;
        const int h = a.list[item];
        const int32_t n = a.out_n[h];
        const unsigned long long off = a.out_off[h], key_j = a.out_key[h];
        if (n <= 0) continue;
        if (off + (unsigned long long)n > a.member_cap) {
            if (tid == 0) atomicOr(a.flags, 1u);
            continue;
        }
        Center c;
        c.x = a.centers[3 * h + 0]; c.y = a.centers[3 * h + 1]; c.z = a.centers[3 * h + 2];
        if (tid == 0) sm.cnt = 0u;
        gsync<NT>();
        EmitF f;
        f.g = &a.g; f.c = c; f.key_j = key_j;
        f.members = a.members + off; f.md2 = a.md2 ? a.md2 + off : nullptr;
        f.cnt = &sm.cnt; f.limit = (uint32_t)n;
        const float r2j = __uint_as_float((uint32_t)(key_j >> 32));
        BallGeom B = make_geom(a.g, c, sqrt((double)r2j) * (1.0 + 1.0e-6));
        for_each_in_ball<NT>(a.g, sm, tid, B, f, ev);
        if (tid == 0 && sm.cnt != (uint32_t)n) atomicOr(a.flags, 2u);
        gsync<NT>();
    }
    ev = __reduce_add_sync(0xFFFFFFFFu, ev);
    if ((threadIdx.x & 31) == 0 && ev) atomicAdd(&a.evals[1], (unsigned long long)ev);
}

template <int NT> static size_t query_smem_bytes(bool with_tma = true)
{
    size_t mt_bytes = (sizeof(MassTableS) + 127) & ~(size_t)127;
    size_t g_bytes = (sizeof(GroupSmem<NT>) + 127) & ~(size_t)127;
    (void)with_tma;
    return mt_bytes + g_bytes * Cfg<NT>::GROUPS;
}

/* append `item` to list[*n ..) for the lanes with pred set: one atomic per warp */
__device__ __forceinline__ void warp_append(bool pred, int32_t item, int32_t *list, uint32_t *n)
{
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, pred);
    if (!m) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(n, (uint32_t)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (pred) list[base + __popc(m & ((1u << lane) - 1u))] = item;
}

/* split the halos into the warp-kernel list and the block-kernel list by expected ball size:
 * a halo of radius R ~ 1.25 rgtp at mean density thr holds thr*(4pi/3)R^3/m particles, the final
 * ball (1.2 R) about 1.3x that */
__global__ void k_classify(const float *__restrict__ rgtp, int nh, float thr, const so_mass_table *mt,
                           float small_max, float huge_min, int32_t *small_list, uint32_t *small_n,
                           int32_t *big_list, uint32_t *big_n, int32_t *huge_list, uint32_t *huge_n,
                           const unsigned char *__restrict__ owner, int me)
{
    int h = blockIdx.x * blockDim.x + threadIdx.x;          /* blockDim is a multiple of 32: whole warps stay */
    const bool in = h < nh && (!owner || owner[h] == (unsigned char)me);   /* domain steps: my share only */
    float r = 1.25f * (in ? rgtp[h] : 0.0f);
    float m = mt->n > 0 ? mt->m : 1.0f;
    float est = 1.3f * thr * 4.18879f * r * r * r / m;
    const bool sm = !(est > small_max), bg = !sm && !(est > huge_min);
    warp_append(in && sm, h, small_list, small_n);
    warp_append(in && bg, h, big_list, big_n);
    warp_append(in && !sm && !bg, h, huge_list, huge_n);
}

/* exclusive scan of max(N_Delta,0) in catalog order -> member offsets; also the emit work lists.
 * One block; every thread owns OFF_PER consecutive halos per round, the three list cursors are
 * scanned together with the offsets (packed 3 x 16 bit), so the lists need no atomics. */
#define OFF_PER 4
template <typename T> __device__ __forceinline__ T block_scan_incl_1024(T v, T *ws, int lane, int w)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T t = __shfl_up_sync(0xFFFFFFFFu, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) ws[w] = v;
    __syncthreads();
    if (w == 0) {
        T y = ws[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            T t = __shfl_up_sync(0xFFFFFFFFu, y, o);
            if (lane >= o) y += t;
        }
        ws[lane] = y;
    }
    __syncthreads();
    if (w) v += ws[w - 1];
    return v;
}

__global__ void __launch_bounds__(1024) k_offsets(const int32_t *__restrict__ out_n, int nh,
                                                  unsigned long long *__restrict__ out_off,
                                                  unsigned long long *__restrict__ total, int32_t emit_small_max,
                                                  int32_t *small_list, uint32_t *small_n, int32_t *big_list,
                                                  uint32_t *big_n, int32_t emit_huge_min = 0x7FFFFFFF,
                                                  int32_t *huge_list = nullptr, uint32_t *huge_n = nullptr)
{
    __shared__ unsigned long long ws[32], ws2[32];
    __shared__ unsigned long long stage[1024 * OFF_PER];   /* offsets of a round, written out coalesced */
    __shared__ unsigned long long carry;
    __shared__ uint32_t carry_c[3];                /* entries already in the small / big / huge list */
    if (threadIdx.x == 0) { carry = 0ull; carry_c[0] = carry_c[1] = carry_c[2] = 0u; }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int b0 = 0; b0 < nh; b0 += 1024 * OFF_PER) {
        const int i0 = b0 + threadIdx.x * OFF_PER;
        int32_t n[OFF_PER];
        unsigned long long v = 0ull, cls = 0ull;
        if (i0 + OFF_PER <= nh) {                            /* 16 contiguous bytes per thread */
            const int4 n4 = *reinterpret_cast<const int4 *>(out_n + i0);
            n[0] = n4.x; n[1] = n4.y; n[2] = n4.z; n[3] = n4.w;
        } else {
#pragma unroll
            for (int k = 0; k < OFF_PER; ++k) n[k] = (i0 + k < nh) ? out_n[i0 + k] : 0;
        }
#pragma unroll
        for (int k = 0; k < OFF_PER; ++k) {
            if (n[k] > 0) {
                v += (unsigned long long)n[k];
                const bool sm = n[k] <= emit_small_max, bg = !sm && (n[k] < emit_huge_min || !huge_list);
                cls += sm ? 1ull : bg ? (1ull << 16) : (1ull << 32);
            }
        }
        const unsigned long long incl = block_scan_incl_1024(v, ws, lane, w);
        const unsigned long long cincl = block_scan_incl_1024(cls, ws2, lane, w);
        unsigned long long run = carry + incl - v;
        const unsigned long long cex = cincl - cls;          /* per round at most 4096 per list: 3 x 16 bits */
        uint32_t ps = carry_c[0] + (uint32_t)(cex & 0xFFFFull);
        uint32_t pb = carry_c[1] + (uint32_t)((cex >> 16) & 0xFFFFull);
        uint32_t ph = carry_c[2] + (uint32_t)((cex >> 32) & 0xFFFFull);
#pragma unroll
        for (int k = 0; k < OFF_PER; ++k) {
            stage[threadIdx.x * OFF_PER + k] = run;
            if (i0 + k < nh) {
                if (n[k] > 0) {
                    run += (unsigned long long)n[k];
                    const bool sm = n[k] <= emit_small_max, bg = !sm && (n[k] < emit_huge_min || !huge_list);
                    if (sm) small_list[ps++] = i0 + k;
                    else if (bg) big_list[pb++] = i0 + k;
                    else huge_list[ph++] = i0 + k;
                }
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < 1024 * OFF_PER && b0 + k < nh; k += 1024) out_off[b0 + k] = stage[k];
        if (threadIdx.x == 1023) {
            carry += incl;
            carry_c[0] += (uint32_t)(cincl & 0xFFFFull);
            carry_c[1] += (uint32_t)((cincl >> 16) & 0xFFFFull);
            carry_c[2] += (uint32_t)((cincl >> 32) & 0xFFFFull);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out_off[nh] = carry; *total = carry;
        *small_n = carry_c[0];
        *big_n = carry_c[1];
        if (huge_n) *huge_n = carry_c[2];
    }
}

/* batched smBallGather, phase 1: count the particles with fDist2 <= ball2[h] (one warp per ball) */
struct CountF {
    const GridDev *g;
    Center c;
    uint32_t hi_bits;
    uint32_t n;
    __device__ __forceinline__ void operator()(uint32_t, const float4 &q)
    {
        if (__float_as_uint(dist2(c, q, *g)) <= hi_bits) ++n;
    }
};

__global__ void __launch_bounds__(256) k_ball_count(const __grid_constant__ QueryArgs a, const float *ball2)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t mt_bytes = (sizeof(MassTableS) + 127) & ~(size_t)127;
    const int grp = threadIdx.x / 32, tid = threadIdx.x % 32;
    GroupSmem<32> &sm = *reinterpret_cast<GroupSmem<32> *>(
        smem_raw + mt_bytes + (size_t)grp * ((sizeof(GroupSmem<32>) + 127) & ~(size_t)127));
    const uint32_t nlist = *a.list_n;
    uint32_t ev = 0;
    for (;;) {
        const uint32_t h = next_item<32>(a.work_counter, sm, tid);
        if (h >= nlist) break;
        const float b2 = ball2[h];
        uint32_t n = 0;
        if (b2 >= 0.0f && b2 < INFINITY) {
            CountF f;
            f.g = &a.g; f.c.x = a.centers[3 * h]; f.c.y = a.centers[3 * h + 1]; f.c.z = a.centers[3 * h + 2];
            f.hi_bits = __float_as_uint(b2); f.n = 0;
            BallGeom B = make_geom(a.g, f.c, sqrt((double)b2) * (1.0 + 1.0e-6));
            for_each_in_ball<32>(a.g, sm, tid, B, f, ev);
            n = __reduce_add_sync(0xFFFFFFFFu, f.n);
        }
        if (tid == 0) {
            a.out_n[h] = (int32_t)n;
            /* key just above every (r^2 <= ball2, index): the emit kernel's "< key_j" keeps them all */
            a.out_key[h] = ((unsigned long long)(__float_as_uint(b2) + 1u)) << 32;
        }
    }
    ev = __reduce_add_sync(0xFFFFFFFFu, ev);
    if ((threadIdx.x & 31) == 0 && ev) atomicAdd(&a.evals[1], (unsigned long long)ev);
}

/* member lists from solve order (one slot per halo, reserved by the query kernels with an atomic cursor) into
 * catalog order (CSR): warp per halo, the CTA together for the large ones */
__global__ void __launch_bounds__(256) k_compact_members(const int32_t *__restrict__ out_n, int nh,
                                                         const unsigned long long *__restrict__ src_off,
                                                         const unsigned long long *__restrict__ csr_off,
                                                         const int32_t *__restrict__ src, int32_t *__restrict__ dst,
                                                         const float *__restrict__ src2, float *__restrict__ dst2)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int h0 = blockIdx.x * 8; h0 < nh; h0 += gridDim.x * 8) {
        const int h = h0 + w;
        const int n = h < nh ? max(out_n[h], 0) : 0;
        if (n > 0 && n <= 8192) {
            const unsigned long long a = src_off[h], b = csr_off[h];
            for (int i = lane; i < n; i += 32) {
                dst[b + i] = src[a + i];
                if (src2) dst2[b + i] = src2[a + i];
            }
        }
        for (int k = 0; k < 8 && h0 + k < nh; ++k) {
            const int nn = max(out_n[h0 + k], 0);
            if (nn <= 8192) continue;
            const unsigned long long a = src_off[h0 + k], b = csr_off[h0 + k];
            for (int i = threadIdx.x; i < nn; i += 256) {
                dst[b + i] = src[a + i];
                if (src2) dst2[b + i] = src2[a + i];
            }
        }
    }
}

/* identity work list 0..n-1 */
__global__ void k_iota(int32_t *list, uint32_t *list_n, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) list[i] = i;
    if (i == 0) *list_n = (uint32_t)n;
}

/* the running-mass table, built on the device so that build -> query needs no host round trip */
__global__ void k_mass_table(const uint32_t *__restrict__ massmm, so_mass_table *mt, unsigned long long kmax)
{
    if (threadIdx.x || blockIdx.x) return;
    if (massmm[0] != massmm[1]) { mt->n = -1; mt->m = 0.0f; return; }     /* unequal masses */
    if (so_mass_table_build(mt, __uint_as_float(massmm[0]), kmax)) { mt->n = -2; }
}

/* sogpu_ball_gather: one warp per row of cells, lanes stride over the row's particles */
__global__ void __launch_bounds__(256) k_ball_gather(const __grid_constant__ GridDev g, float cx, float cy,
                                                     float cz, float ball2, int32_t *idx, float *d2o,
                                                     unsigned long long *cnt, unsigned long long cap)
{
    Center c; c.x = cx; c.y = cy; c.z = cz;
    GatherF f;
    f.g = &g; f.c = c; f.hi_bits = __float_as_uint(ball2); f.idx = idx; f.d2o = d2o; f.cnt = cnt; f.cap = cap;
    BallGeom B = make_geom(g, c, sqrt((double)ball2) * (1.0 + 1.0e-6));
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int r = wid; r < B.nrows; r += nw) {
        uint32_t s0, l0, s1, l1;
        row_segments(g, B, r, s0, l0, s1, l1);
        for (uint32_t i = lane; i < l0; i += 32) f(s0 + i, __ldg(g.sorted + s0 + i));
        for (uint32_t i = lane; i < l1; i += 32) f(s1 + i, __ldg(g.sorted + s1 + i));
    }
}

/* ============================================================================================
 * general path: particles of unequal mass (gas + dark + star snapshots)
 *
 * The enclosed mass is a sequential fp32 sum in sorted order (kd2.c:787,807), so with unequal masses
 * nothing short of the fully sorted list reproduces it.  Per ball of the schedule and per halo:
 * count -> emit (key, mass) -> segmented sort (CUB) -> the reference's scan, literally, by one warp
 * (serial fp32 adds replicated on all lanes, density tests spread over the lanes).
 * ============================================================================================ */
struct GenState {            /* per halo, carried from ball to ball (kdRvir's locals) */
    float ball;              /* fBall of the last gathered ball                               */
    float mass;              /* running mass (kd2.c:744,787,807)                              */
    uint32_t jlast;          /* kd2.c:743,797,832                                             */
};

struct GenArgs {
    GridDev g;
    const float4 *in;        /* original particle array: mass of particle i = in[i].w          */
    const float *centers, *rgtp;
    GenState *st;
    const int32_t *list; const uint32_t *list_n; uint32_t *work;      /* halos of this round  */
    int32_t *round_list; uint32_t *round_n;   /* (count) halos that need emit+sort+scan        */
    unsigned long long *seg_n;                /* particles in the ball, per round slot          */
    const unsigned long long *seg_begin;
    unsigned long long *keys; float *mass;    /* scratch segments                                */
    int32_t *next_list; uint32_t *next_n;     /* (scan) halos that need a bigger ball            */
    float thr; int nM;
    int32_t *out_n; float *out_m; unsigned long long *out_key;
    unsigned long long *evals;
};

struct EmitPairF {
    const GridDev *g; const float4 *in;
    Center c; uint32_t hi_bits;
    unsigned long long *keys; float *mass; uint32_t *cnt; uint32_t limit;
    __device__ __forceinline__ void operator()(uint32_t, const float4 &q)
    {
        uint32_t bits = __float_as_uint(dist2(c, q, *g));
        if (bits <= hi_bits) {
            uint32_t pos = atomicAdd(cnt, 1u);
            if (pos < limit) {
                uint32_t idx = __float_as_uint(q.w);
                keys[pos] = ((unsigned long long)bits << 32) | idx;
                mass[pos] = __ldg(&in[idx].w);
            }
        }
    }
};

/* advance every active halo to its next ball and count the particles in it (kd2.c:766-778) */
__global__ void __launch_bounds__(256) k_gen_count(const __grid_constant__ GenArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t mt_bytes = (sizeof(MassTableS) + 127) & ~(size_t)127;
    GroupSmem<256> &sm = *reinterpret_cast<GroupSmem<256> *>(smem_raw + mt_bytes);
    const int tid = threadIdx.x;
    const uint32_t nlist = *a.list_n;
    uint32_t ev = 0;
    const float root = so_root_period(a.g.L[0], a.g.L[1], a.g.L[2]);
    for (;;) {
        const uint32_t item = next_item<256>(a.work, sm, tid);
        if (item >= nlist) break;
        const int h = a.list[item];
        GenState st = a.st[h];
        if (!((double)st.ball < 0.25 * (double)root) || !(st.ball > 0.0f)) {     /* kd2.c:766, 837-839 */
            if (tid == 0) { a.out_n[h] = -3; a.out_m[h] = -3.0f; a.out_key[h] = 0ull; }
            continue;
        }
        st.ball = so_next_ball(st.ball);
        const float ball2 = __fmul_rn(st.ball, st.ball);
        CountF f;
        f.g = &a.g; f.c.x = a.centers[3 * h]; f.c.y = a.centers[3 * h + 1]; f.c.z = a.centers[3 * h + 2];
        f.hi_bits = __float_as_uint(ball2); f.n = 0;
        BallGeom B = make_geom(a.g, f.c, sqrt((double)ball2) * (1.0 + 1.0e-6));
        for_each_in_ball<256>(a.g, sm, tid, B, f, ev);
        const uint32_t n = gsum<256>(f.n, sm.tmp, tid);
        if (tid == 0) {
            a.st[h].ball = st.ball;
            if (st.jlast == 0 && n < (uint32_t)a.nM) {                            /* kd2.c:772-778 */
                a.out_n[h] = -1; a.out_m[h] = -1.0f; a.out_key[h] = 0ull;
            } else {
                uint32_t slot = atomicAdd(a.round_n, 1u);
                a.round_list[slot] = h;
                a.seg_n[slot] = n;
            }
        }
    }
    ev = __reduce_add_sync(0xFFFFFFFFu, ev);
    if ((threadIdx.x & 31) == 0 && ev) atomicAdd(&a.evals[1], (unsigned long long)ev);
}

/* exclusive scan of the per-slot counts of slots [s0, s1) (one block) */
__global__ void __launch_bounds__(1024) k_gen_offsets(const unsigned long long *__restrict__ seg_n, uint32_t s0,
                                                      uint32_t s1, unsigned long long *__restrict__ seg_begin,
                                                      unsigned long long *__restrict__ seg_end,
                                                      unsigned long long *__restrict__ csr)
{
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (uint32_t b0 = s0; b0 < s1; b0 += 1024) {
        uint32_t i = b0 + threadIdx.x;
        unsigned long long v = (i < s1) ? seg_n[i] : 0ull, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) ws[w] = x;
        __syncthreads();
        if (w == 0) {
            unsigned long long y = ws[lane];
            for (int o = 1; o < 32; o <<= 1) {
                unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, y, o);
                if (lane >= o) y += t;
            }
            ws[lane] = y;
        }
        __syncthreads();
        unsigned long long incl = x + (w ? ws[w - 1] : 0ull) + carry;
        if (i < s1) { seg_begin[i] = incl - v; seg_end[i] = incl; csr[i - s0] = incl - v; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) csr[s1 - s0] = carry;
}

/* masses in the order of the sorted keys (the key's low half is the particle index) */
__global__ void __launch_bounds__(256) k_gen_mass(const unsigned long long *__restrict__ keys, unsigned long long n,
                                                  const float4 *__restrict__ in, float *__restrict__ mass)
{
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        mass[i] = __ldg(&in[(uint32_t)keys[i]].w);
}

/* write (key, mass) of every particle of the ball into the halo's scratch segment */
__global__ void __launch_bounds__(256) k_gen_emit(const __grid_constant__ GenArgs a, uint32_t s0, uint32_t s1)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const size_t mt_bytes = (sizeof(MassTableS) + 127) & ~(size_t)127;
    GroupSmem<256> &sm = *reinterpret_cast<GroupSmem<256> *>(smem_raw + mt_bytes);
    const int tid = threadIdx.x;
    uint32_t ev = 0;
    for (uint32_t slot = s0 + blockIdx.x; slot < s1; slot += gridDim.x) {
        const int h = a.round_list[slot];
        const float ball = a.st[h].ball;
        const float ball2 = __fmul_rn(ball, ball);
        if (tid == 0) sm.cnt = 0u;
        __syncthreads();
        EmitPairF f;
        f.g = &a.g; f.in = a.in;
        f.c.x = a.centers[3 * h]; f.c.y = a.centers[3 * h + 1]; f.c.z = a.centers[3 * h + 2];
        f.hi_bits = __float_as_uint(ball2);
        f.keys = a.keys + a.seg_begin[slot]; f.mass = a.mass + a.seg_begin[slot];
        f.cnt = &sm.cnt; f.limit = (uint32_t)a.seg_n[slot];
        BallGeom B = make_geom(a.g, f.c, sqrt((double)ball2) * (1.0 + 1.0e-6));
        for_each_in_ball<256>(a.g, sm, tid, B, f, ev);
        __syncthreads();
    }
    ev = __reduce_add_sync(0xFFFFFFFFu, ev);
    if ((threadIdx.x & 31) == 0 && ev) atomicAdd(&a.evals[1], (unsigned long long)ev);
}

/* kdRvir's scan over the sorted ball (kd2.c:785-832), one warp per halo */
__global__ void __launch_bounds__(256) k_gen_scan(const __grid_constant__ GenArgs a, uint32_t s0, uint32_t s1,
                                                  const unsigned long long *__restrict__ keys,
                                                  const float *__restrict__ mass)
{
    const int lane = threadIdx.x & 31;
    const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t slot = s0 + wid; slot < s1; slot += nw) {
        const int h = a.round_list[slot];
        const unsigned long long *K = keys + a.seg_begin[slot];
        const float *M = mass + a.seg_begin[slot];
        const uint32_t n = (uint32_t)a.seg_n[slot];
        GenState st = a.st[h];
        float run = st.mass;
        uint32_t j = st.jlast;
        bool done = false;
        if (j == 0) {                                                  /* first ball: kd2.c:785-798 */
            for (int k = 0; k < a.nM - 1; ++k) run = __fadd_rn(run, M[k]);
            j = (uint32_t)(a.nM - 1);
            float d2a = __uint_as_float((uint32_t)(K[j - 1] >> 32)), d2b = __uint_as_float((uint32_t)(K[j] >> 32));
            if (so_rho_below(run, d2a, a.thr) && so_rho_below(__fadd_rn(run, M[j]), d2b, a.thr)) {
                if (lane == 0) { a.out_n[h] = -2; a.out_m[h] = -2.0f; a.out_key[h] = 0ull; }
                done = true;
            }
        }
        /* main loop, kd2.c:804-831: for (j = jlast; j < n-1; j++) { mass += m[j]; test j and j+1 } */
        while (!done && j + 1 < n) {
            /* this chunk handles j0 = j .. j0+31 (as far as j+1 < n); lane t owns element j0+t */
            const uint32_t j0 = j;
            const uint32_t e = j0 + lane;
            const float m_l = (e < n) ? M[e] : 0.0f;
            const unsigned long long k_l = (e < n) ? K[e] : 0ull;
            /* serial running sum, replicated on every lane: S_t = mass after adding element j0+t */
            float S = run, mine = 0.0f;
            for (int t = 0; t < 32; ++t) {
                float mt_ = __shfl_sync(0xFFFFFFFFu, m_l, t);
                S = __fadd_rn(S, mt_);
                if (t == lane) mine = S;
            }
            /* below(e) = rho(S_e, d2_e) < thr ; valid where e < n */
            bool below = (e < n) && so_rho_below(mine, __uint_as_float((uint32_t)(k_l >> 32)), a.thr);
            unsigned bal = __ballot_sync(0xFFFFFFFFu, below);
            /* pair (e, e+1) needs e+1 < n; lane 31's partner is the next chunk's lane 0: handle by
             * limiting this chunk to 31 pairs and re-starting the next chunk at j0+31 */
            unsigned pairs = bal & (bal >> 1);
            unsigned valid = 0u;
            {
                uint32_t last_pair = n - 2u - j0;                     /* largest t with (j0+t)+1 < n */
                valid = last_pair >= 30u ? 0x7FFFFFFFu : ((2u << last_pair) - 1u);
            }
            pairs &= valid;
            if (pairs) {
                int t = __ffs(pairs) - 1;                              /* first firing j = j0 + t */
                float Sj = __shfl_sync(0xFFFFFFFFu, mine, t);
                float mj = __shfl_sync(0xFFFFFFFFu, m_l, t);
                unsigned long long kj = __shfl_sync(0xFFFFFFFFu, k_l, t);
                if (lane == 0) {
                    a.out_n[h] = (int32_t)(j0 + t);                    /* kd2.c:823 */
                    a.out_m[h] = __fsub_rn(Sj, mj);                    /* kd2.c:816 */
                    a.out_key[h] = kj;
                }
                done = true;
                break;
            }
            /* no hit among pairs t = 0..30: continue at element j0+31 with the mass before it */
            uint32_t adv = min(31u, n - 1u - j0);
            run = __shfl_sync(0xFFFFFFFFu, mine, (int)adv - 1);        /* mass after element j0+adv-1 */
            j = j0 + adv;
        }
        if (!done && lane == 0) {
            /* kd2.c:832: jlast = j (= n-1); mass holds elements 0..n-2 */
            a.st[h].jlast = j;
            a.st[h].mass = run;
            a.next_list[atomicAdd(a.next_n, 1u)] = h;
        }
    }
}

__global__ void k_gen_init(GenState *st, const float *rgtp, int32_t *list, uint32_t *list_n, int n,
                           const int32_t *subset)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        int h = subset ? subset[i] : i;
        st[h].ball = rgtp[h]; st[h].mass = 0.0f; st[h].jlast = 0u;     /* kd2.c:743-745 */
        list[i] = h;
    }
    if (i == 0) *list_n = (uint32_t)n;
}

/* halos whose result is `code`, as a compact list */
__global__ void k_select_code(const int32_t *out_n, int nh, int32_t code, int32_t *list, uint32_t *list_n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nh && out_n[i] == code) list[atomicAdd(list_n, 1u)] = i;
}

/* ============================================================================================
 * kdVcirc + kdMassProfile (kd2.c:498-586, 458-496) over the r^2-sorted 2*Rvir lists, equal masses:
 * the reference's cumulative fp32 mass after k particles is S[k] (mass table), so every quantity
 * is a rank query on the sorted r^2 list.  One 256-thread CTA per group.
 * ============================================================================================ */
#define SO_NVCIRC 8          /* kd2.h:9  */
#define SO_NMASSPROFILE 16   /* kd2.h:11 */

struct VcircArgs {
    const float *d2;                      /* sorted r^2 of all lists, CSR */
    const unsigned long long *off;        /* nh + 1 */
    const float *rvir, *mvir;
    const so_mass_table *mt;
    float G;
    int nM, nh;
    float *vcirc, *rmass, *rmax, *vmax, *profile;   /* 8, 2, 1, 1, 16 per group (profile may be NULL) */
};

__device__ __forceinline__ uint32_t lower_bound_f(const float *__restrict__ a, uint32_t n, float x)
{
    uint32_t lo = 0, hi = n;                       /* first j with !(a[j] < x): the reference's while (d2[j] < r2) */
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_vcirc(const __grid_constant__ VcircArgs a)
{
    __shared__ MassTableS mt;
    __shared__ float s_v[8];
    __shared__ uint32_t s_j[8];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int mtn = a.mt->n;
    if (t == 0) { mt.n = mtn; mt.m = a.mt->m; }
    for (int i = t; i <= mtn; i += 256) mt.k0[i] = a.mt->k0[i];
    for (int i = t; i < mtn; i += 256) { mt.s0[i] = a.mt->s0[i]; mt.inc[i] = a.mt->inc[i]; }
    __syncthreads();
    for (int h = blockIdx.x; h < a.nh; h += gridDim.x) {
        const unsigned long long o = a.off[h];
        const uint32_t n = (uint32_t)(a.off[h + 1] - o);
        const float *d2 = a.d2 + o;
        const float rvir = a.rvir[h], mvir = a.mvir[h];
        if (n == 0) {                              /* cannot happen for a resolved group (N_Delta >= nMembers) */
            if (t < SO_NVCIRC) a.vcirc[(size_t)h * SO_NVCIRC + t] = 0.0f;
            if (t < 2) a.rmass[(size_t)h * 2 + t] = 0.0f;
            if (t == 0) { a.rmax[h] = 0.0f; a.vmax[h] = 0.0f; }
            if (a.profile && t < SO_NMASSPROFILE) a.profile[(size_t)h * SO_NMASSPROFILE + t] = 0.0f;
            continue;
        }
        if (t < SO_NVCIRC) {                       /* kd2.c:517-531 */
            const float fmin = (float)(2.0 / SO_NVCIRC);
            float f = fmin;
            for (int i = 0; i < t; ++i) f = __fadd_rn(f, fmin);
            float v;
            if (t < SO_NVCIRC - 1) {
                const float r = __fmul_rn(f, rvir), r2 = __fmul_rn(r, r);
                const float mass = mt_eval(mt, lower_bound_f(d2, n, r2));
                v = __fsqrt_rn(__fdiv_rn(__fmul_rn(a.G, mass), r));
            } else {
                const float fBall = 2.0f * rvir;
                v = __fsqrt_rn(__fdiv_rn(__fmul_rn(a.G, mt_eval(mt, n)), fBall));
            }
            a.vcirc[(size_t)h * SO_NVCIRC + t] = v;
        } else if (t >= 32 && t < 32 + 2) {        /* kd2.c:537-546: radii holding 1/4 and 1/2 of Mvir */
            const int i = t - 32;
            const float f = i ? 0.5f : 0.25f, m = __fmul_rn(f, mvir);
            uint32_t lo = 0, hi = n - 1;           /* first j with S[j+1] >= m, at most n-1 */
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (mt_eval(mt, mid + 1u) < m) lo = mid + 1; else hi = mid;
            }
            a.rmass[(size_t)h * 2 + i] = __fsqrt_rn(__ldg(d2 + lo));
        } else if (a.profile && t >= 64 && t < 64 + SO_NMASSPROFILE) {   /* kd2.c:458-496, every particle counted */
            const int i = t - 64;
            const float fmin = (float)(2.0 / SO_NMASSPROFILE);
            float f = fmin;
            for (int k = 0; k < i; ++k) f = __fadd_rn(f, fmin);
            float mass;
            if (i < SO_NMASSPROFILE - 1) {
                const float r = __fmul_rn(f, rvir), r2 = __fmul_rn(r, r);
                mass = mt_eval(mt, lower_bound_f(d2, n, r2));
            } else {
                mass = mt_eval(mt, n);
            }
            a.profile[(size_t)h * SO_NMASSPROFILE + i] = mass;
        }
        /* kd2.c:551-569: maximum of Vc over the list from the nMembers-th particle on, first maximum wins */
        const uint32_t j0 = ((uint32_t)a.nM <= n ? (uint32_t)a.nM : n) - 1u;
        float best = -1.0f;
        uint32_t bj = 0xFFFFFFFFu;
        for (uint32_t j = j0 + (uint32_t)t; j < n; j += 256u) {
            const float r = __fsqrt_rn(__ldg(d2 + j));
            const float vc = __fsqrt_rn(__fdiv_rn(__fmul_rn(a.G, mt_eval(mt, j + 1u)), r));
            if (vc > best) { best = vc; bj = j; }          /* ascending j per thread: keeps the first */
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            float ov = __shfl_down_sync(0xFFFFFFFFu, best, o2);
            uint32_t oj = __shfl_down_sync(0xFFFFFFFFu, bj, o2);
            if (ov > best || (ov == best && oj < bj)) { best = ov; bj = oj; }
        }
        if (lane == 0) { s_v[w] = best; s_j[w] = bj; }
        __syncthreads();
        if (t == 0) {
            for (int k = 1; k < 8; ++k)
                if (s_v[k] > best || (s_v[k] == best && s_j[k] < bj)) { best = s_v[k]; bj = s_j[k]; }
            a.vmax[h] = best;
            a.rmax[h] = __fsqrt_rn(__ldg(d2 + bj));
        }
        __syncthreads();
    }
}

/* ---- kdVcirc / kdMassProfile for ANY masses and several species (kd2.c:458-496, 498-586) ----------------
 * With unequal masses the cumulative mass after k sorted particles is a sequential fp32 sum that depends on
 * which particle sits at which rank; k_vc_prefix evaluates it literally — one warp per group walks the sorted
 * 2 Rvir list 32 entries at a time, the adds replicated serially on every lane (they cannot be reassociated) —
 * for all particles and for up to four species masks (kdMassProfile's per-species sums, each its own sequential
 * chain).  Every output of kdVcirc is then a rank query on those prefix arrays (k_vcirc_gen). */
#define VC_MAXMASK 4
__global__ void __launch_bounds__(256) k_vc_prefix(const unsigned long long *__restrict__ off, const int32_t *__restrict__ idx,
                                                   int nh, const float4 *__restrict__ in, const unsigned char *__restrict__ ptype,
                                                   int nmask, uint32_t masks4, unsigned long long stride, float *__restrict__ C)
{
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int h = wid; h < nh; h += nw) {
        const unsigned long long a = off[h];
        const uint32_t n = (uint32_t)(off[h + 1] - a);
        float S[1 + VC_MAXMASK];
#pragma unroll
        for (int k = 0; k <= VC_MAXMASK; ++k) S[k] = 0.0f;
        for (uint32_t base = 0; base < n; base += 32) {
            const uint32_t e = base + lane;
            float m_l = 0.0f;
            uint32_t t_l = 0u;
            if (e < n) {
                const int32_t p = __ldg(idx + a + e);
                m_l = __ldg(&in[p].w);
                t_l = ptype ? (uint32_t)__ldg(ptype + p) : 0xFFu;
            }
            float mine[1 + VC_MAXMASK];
#pragma unroll
            for (int k = 0; k <= VC_MAXMASK; ++k) mine[k] = 0.0f;
            for (int t = 0; t < 32; ++t) {
                const float mt = __shfl_sync(0xFFFFFFFFu, m_l, t);      /* (padding entries add +0.0f: exact) */
                const uint32_t tt = __shfl_sync(0xFFFFFFFFu, t_l, t);
                S[0] = __fadd_rn(S[0], mt);
#pragma unroll
                for (int k = 0; k < VC_MAXMASK; ++k)
                    if (k < nmask && (tt & ((masks4 >> (8 * k)) & 0xFFu))) S[1 + k] = __fadd_rn(S[1 + k], mt);
                if (t == lane) {
#pragma unroll
                    for (int k = 0; k <= VC_MAXMASK; ++k) mine[k] = S[k];
                }
            }
            if (e < n) {
                C[a + e] = mine[0];
#pragma unroll
                for (int k = 0; k < VC_MAXMASK; ++k)
                    if (k < nmask) C[(unsigned long long)(1 + k) * stride + a + e] = mine[1 + k];
            }
        }
    }
}

struct VcircGenArgs {
    const float *d2;                      /* sorted r^2 of all lists, CSR */
    const unsigned long long *off;        /* nh + 1 */
    const float *C;                       /* (1 + nmask) x stride inclusive prefixes */
    unsigned long long stride;
    const float *rvir, *mvir;
    float G;
    int nM, nh, nmask;
    float *vcirc, *rmass, *rmax, *vmax, *profiles;   /* 8, 2, 1, 1 per group; profiles: nmask x nh x 16 */
};

__global__ void __launch_bounds__(256) k_vcirc_gen(const __grid_constant__ VcircGenArgs a)
{
    __shared__ float s_v[8];
    __shared__ uint32_t s_j[8];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    for (int h = blockIdx.x; h < a.nh; h += gridDim.x) {
        const unsigned long long o = a.off[h];
        const uint32_t n = (uint32_t)(a.off[h + 1] - o);
        const float *d2 = a.d2 + o;
        const float *C = a.C + o;
        const float rvir = a.rvir[h], mvir = a.mvir[h];
        auto massk = [&](const float *P, uint32_t k) -> float { return k ? __ldg(P + k - 1u) : 0.0f; };   /* mass of the first k */
        if (n == 0) {
            if (t < SO_NVCIRC) a.vcirc[(size_t)h * SO_NVCIRC + t] = 0.0f;
            if (t < 2) a.rmass[(size_t)h * 2 + t] = 0.0f;
            if (t == 0) { a.rmax[h] = 0.0f; a.vmax[h] = 0.0f; }
            for (int k = 0; k < a.nmask; ++k)
                if (t < SO_NMASSPROFILE) a.profiles[((size_t)k * a.nh + h) * SO_NMASSPROFILE + t] = 0.0f;
            continue;
        }
        if (t < SO_NVCIRC) {                       /* kd2.c:517-531 */
            const float fmin = (float)(2.0 / SO_NVCIRC);
            float f = fmin;
            for (int i = 0; i < t; ++i) f = __fadd_rn(f, fmin);
            float v;
            if (t < SO_NVCIRC - 1) {
                const float r = __fmul_rn(f, rvir), r2 = __fmul_rn(r, r);
                v = __fsqrt_rn(__fdiv_rn(__fmul_rn(a.G, massk(C, lower_bound_f(d2, n, r2))), r));
            } else {
                const float fBall = 2.0f * rvir;
                v = __fsqrt_rn(__fdiv_rn(__fmul_rn(a.G, massk(C, n)), fBall));
            }
            a.vcirc[(size_t)h * SO_NVCIRC + t] = v;
        } else if (t >= 32 && t < 32 + 2) {        /* kd2.c:537-546: first j whose cumulative mass reaches f * Mvir */
            const int i = t - 32;
            const float f = i ? 0.5f : 0.25f, m = __fmul_rn(f, mvir);
            uint32_t lo = 0, hi = n - 1;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (__ldg(C + mid) < m) lo = mid + 1; else hi = mid;
            }
            a.rmass[(size_t)h * 2 + i] = __fsqrt_rn(__ldg(d2 + lo));
        } else if (t >= 64 && t < 64 + SO_NMASSPROFILE * VC_MAXMASK) {   /* kd2.c:458-496, one species mask per 16 threads */
            const int k = (t - 64) / SO_NMASSPROFILE, i = (t - 64) % SO_NMASSPROFILE;
            if (k < a.nmask) {
                const float *Cs = a.C + (unsigned long long)(1 + k) * a.stride + o;
                const float fmin = (float)(2.0 / SO_NMASSPROFILE);
                float f = fmin;
                for (int q = 0; q < i; ++q) f = __fadd_rn(f, fmin);
                float mass;
                if (i < SO_NMASSPROFILE - 1) {
                    const float r = __fmul_rn(f, rvir), r2 = __fmul_rn(r, r);
                    mass = massk(Cs, lower_bound_f(d2, n, r2));
                } else {
                    mass = massk(Cs, n);
                }
                a.profiles[((size_t)k * a.nh + h) * SO_NMASSPROFILE + i] = mass;
            }
        }
        /* kd2.c:551-569: maximum of Vc from the nMembers-th particle on, the first maximum wins */
        const uint32_t j0 = ((uint32_t)a.nM <= n ? (uint32_t)a.nM : n) - 1u;
        float best = -1.0f;
        uint32_t bj = 0xFFFFFFFFu;
        for (uint32_t j = j0 + (uint32_t)t; j < n; j += 256u) {
            const float r = __fsqrt_rn(__ldg(d2 + j));
            const float vc = __fsqrt_rn(__fdiv_rn(__fmul_rn(a.G, __ldg(C + j)), r));
            if (vc > best) { best = vc; bj = j; }
        }
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) {
            float ov = __shfl_down_sync(0xFFFFFFFFu, best, o2);
            uint32_t oj = __shfl_down_sync(0xFFFFFFFFu, bj, o2);
            if (ov > best || (ov == best && oj < bj)) { best = ov; bj = oj; }
        }
        if (lane == 0) { s_v[w] = best; s_j[w] = bj; }
        __syncthreads();
        if (t == 0) {
            for (int k = 1; k < 8; ++k)
                if (s_v[k] > best || (s_v[k] == best && s_j[k] < bj)) { best = s_v[k]; bj = s_j[k]; }
            a.vmax[h] = best;
            a.rmax[h] = __fsqrt_rn(__ldg(d2 + bj));
        }
        __syncthreads();
    }
}

/* ============================================================================================
 * domain runs (several GPUs): every rank holds a SLICE of the snapshot and a
 * spatially compact share of the halos.  A rank's "focus mask" marks the coarse cells its halos can
 * reach; a particle is routed to every rank whose mask holds its coarse cell (none: nobody needs it).
 * k_route_count / k_route_scatter are the exchange step: the scatter writes {x, y, z, global index}
 * records straight into the receivers' buffers (peer pointers over NVLink, or local staging
 * buffers for an NCCL all-to-all), one warp-aggregated reservation per (warp, destination).
 * ============================================================================================ */
#define ROUTE_MAXR 16

struct RouteArgs {
    GridDev g;                              /* geometry + mb/ms of the masks (mask pointer unused) */
    const float4 *slice;                    /* this rank's particles {x,y,z,m} */
    int64_t n;
    uint32_t index_base;                    /* global index of slice[0] */
    const unsigned short *table;            /* set of destination ranks per coarse cell (k_route_table) */
    const uint32_t *any;                    /* bit per coarse cell: some rank wants it */
    int R;
    unsigned long long *counts;             /* R counters (count pass) / running cursors (scatter pass) */
    float4 *dst[ROUTE_MAXR];                /* receive buffers (scatter pass) */
    unsigned long long dst_off[ROUTE_MAXR]; /* where this rank's records start in each */
};

__device__ __forceinline__ uint32_t coarse_bit(const float4 &p, const GridDev &g)
{
    const int mask = g.nc - 1;
    uint32_t ix = cell_coord(p.x, g.g0[0], g.invh[0], mask) >> g.ms;
    uint32_t iy = cell_coord(p.y, g.g0[1], g.invh[1], mask) >> g.ms;
    uint32_t iz = cell_coord(p.z, g.g0[2], g.invh[2], mask) >> g.ms;
    return (iz << (2 * g.mb)) | (iy << g.mb) | ix;
}

/* masks of all ranks -> one 16-bit set of destination ranks per coarse cell (one lookup per particle) */
__global__ void __launch_bounds__(256) k_route_table(const uint32_t *__restrict__ masks, uint32_t mask_words, int R,
                                                     uint32_t n_cells, unsigned short *__restrict__ table,
                                                     uint32_t *__restrict__ any)
{
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;          /* one 32-cell word per thread */
    if (w >= mask_words) return;
    uint32_t m[ROUTE_MAXR], u = 0u;
    for (int d = 0; d < R; ++d) { m[d] = __ldg(masks + (size_t)d * mask_words + w); u |= m[d]; }
    any[w] = u;                        /* union of the masks: 2 MB, stays in L1/L2; most particles stop here */
    if (!u) return;
    for (int b = 0; b < 32; ++b) {
        const uint32_t cell = w * 32u + (uint32_t)b;
        if (cell >= n_cells) break;
        uint32_t set = 0;
        for (int d = 0; d < R; ++d) set |= ((m[d] >> b) & 1u) << d;
        table[cell] = (unsigned short)set;
    }
}

#define ROUTE_U 4       /* particles per thread per round: independent loads in flight */
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_route(const __grid_constant__ RouteArgs a)
{
    __shared__ unsigned long long scount[ROUTE_MAXR];          /* count pass: CTA totals             */
    __shared__ uint32_t wcnt[8][ROUTE_MAXR];                   /* scatter pass: records per (warp, destination) of a round */
    __shared__ unsigned long long wbase[8][ROUTE_MAXR];        /* ... and where each warp's run starts */
    if (threadIdx.x < ROUTE_MAXR) scount[threadIdx.x] = 0ull;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t nround = (a.n + stride * ROUTE_U - 1) / (stride * ROUTE_U);
    for (int64_t it = 0; it < nround; ++it) {               /* every thread runs every round (ballots, barriers) */
        const int64_t i0 = it * stride * ROUTE_U + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        float4 q[ROUTE_U];
        uint32_t set[ROUTE_U];
#pragma unroll
        for (int u = 0; u < ROUTE_U; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < a.n) q[u] = ld_stream(a.slice + i);
        }
#pragma unroll
        for (int u = 0; u < ROUTE_U; ++u) {
            const int64_t i = i0 + u * stride;
            set[u] = 0u;
            if (i < a.n) {
                const uint32_t bit = coarse_bit(q[u], a.g);
                if ((__ldg(a.any + (bit >> 5)) >> (bit & 31)) & 1u) set[u] = __ldg(a.table + bit);
                q[u].w = __uint_as_float(a.index_base + (uint32_t)i);
            }
        }
        uint32_t wset = 0u;                                 /* destinations some lane of this warp has */
#pragma unroll
        for (int u = 0; u < ROUTE_U; ++u) wset |= set[u];
        wset = __reduce_or_sync(0xFFFFFFFFu, wset);
        if (!SCATTER) {
            for (uint32_t rem = wset; rem; rem &= rem - 1u) {
                const int d = __ffs(rem) - 1;
                uint32_t c = 0;
#pragma unroll
                for (int u = 0; u < ROUTE_U; ++u) c += __popc(__ballot_sync(0xFFFFFFFFu, (set[u] >> d) & 1u));
                if (lane == 0) atomicAdd(&scount[d], (unsigned long long)c);
            }
        } else {
            /* one reservation per (CTA, round, destination): the warps' runs follow each other, inside a warp
             * the records are ordered by (u, lane) — runs of up to 128 contiguous records per warp */
            if (lane < a.R) wcnt[w][lane] = 0u;
            __syncwarp();
            for (uint32_t rem = wset; rem; rem &= rem - 1u) {
                const int d = __ffs(rem) - 1;
                uint32_t c = 0;
#pragma unroll
                for (int u = 0; u < ROUTE_U; ++u) c += __popc(__ballot_sync(0xFFFFFFFFu, (set[u] >> d) & 1u));
                if (lane == 0) wcnt[w][d] = c;
            }
            __syncthreads();
            if (threadIdx.x < a.R) {
                const int d = threadIdx.x;
                uint32_t tot = 0;
                for (int k = 0; k < 8; ++k) tot += wcnt[k][d];
                unsigned long long base = tot ? atomicAdd(&a.counts[d], (unsigned long long)tot) : 0ull;
                base += a.dst_off[d];
                for (int k = 0; k < 8; ++k) { wbase[k][d] = base; base += wcnt[k][d]; }
            }
            __syncthreads();
            for (uint32_t rem = wset; rem; rem &= rem - 1u) {
                const int d = __ffs(rem) - 1;
                unsigned long long pos = wbase[w][d];
#pragma unroll
                for (int u = 0; u < ROUTE_U; ++u) {
                    const bool want = (set[u] >> d) & 1u;
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, want);
                    if (want) a.dst[d][pos + (unsigned long long)__popc(m & lt)] = q[u];
                    pos += (unsigned long long)__popc(m);
                }
            }
            __syncthreads();                                /* wcnt / wbase are reused by the next round */
        }
    }
    if (!SCATTER) {
        __syncthreads();
        if (threadIdx.x < a.R && scount[threadIdx.x]) atomicAdd(&a.counts[threadIdx.x], scount[threadIdx.x]);
    }
}

/* ============================================================================================
 * kdTagParticles, the part that needs no ordering (kd2.c:663-720): a group none of whose members
 * belongs to another group simply tags its members, whatever the processing order.  Pass 1 lets
 * every member claim its particle with a compare-and-swap; a failed claim marks both groups
 * involved as "in conflict".  Pass 2 writes the catalog index for conflict-free groups and
 * releases the claims of the others, which the caller replays sequentially (subsume / slurp /
 * ignore are order dependent).  A conflict-free group can never be touched by that replay: nobody
 * else owns or meets one of its particles.
 * ============================================================================================ */
__global__ void __launch_bounds__(256) k_tag_claim(const unsigned long long *__restrict__ off, const int32_t *__restrict__ mem,
                                                   int nh, int32_t *tag, unsigned char *dirty)
{
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int h = wid; h < nh; h += nw) {
        const unsigned long long a = off[h], b = off[h + 1];
        for (unsigned long long k = a + lane; k < b; k += 32) {
            const int32_t old = atomicCAS(&tag[mem[k]], 0, h + 1);
            if (old != 0 && old != h + 1) { dirty[h] = 1; dirty[old - 1] = 1; }
        }
    }
}

__global__ void __launch_bounds__(256) k_tag_settle(const unsigned long long *__restrict__ off, const int32_t *__restrict__ mem,
                                                    const int32_t *__restrict__ index, int nh, int32_t *tag,
                                                    const unsigned char *__restrict__ dirty)
{
    const int lane = threadIdx.x & 31;
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (int h = wid; h < nh; h += nw) {
        const unsigned long long a = off[h], b = off[h + 1];
        if (!dirty[h]) {
            const int32_t id = index[h];
            for (unsigned long long k = a + lane; k < b; k += 32) tag[mem[k]] = id;
        } else {
            for (unsigned long long k = a + lane; k < b; k += 32) atomicCAS(&tag[mem[k]], h + 1, 0);
        }
    }
}

/* kdTagParticles, the ORDER-DEPENDENT part (kd2.c:663-720): the groups that share particles are replayed in the
 * reference's processing order (ascending catalog mass, kd2.c:873-879) by ONE CTA.  A group walks its r^2-sorted
 * member list 256 entries at a time.  What happens at an already-tagged particle depends only on the pair of
 * groups and their current radii (kd2.c:677-703): "ignore" leaves everything but the particle's counter as it is,
 * so all entries in front of the first subsume / slurp event of a chunk are applied in parallel; the event itself
 * (kdZeroGroup over the loser's member list, kd2.c:617-643) is applied by the whole CTA, and the walk resumes
 * behind it.  State: tag[] = PINIT.iGrp, nsub[] / nign[] = PINIT.nSubsumed / nIgnored, rvir / mvir = GRPNODE. */
struct ReplayArgs {
    const int32_t *order;                 /* slots of the groups to replay, in processing order */
    int n_order;
    const unsigned long long *off;
    const int32_t *mem;
    const int32_t *index;                 /* catalog id per slot */
    const int32_t *slot_of_index;
    const float *pos;                     /* 3 per slot */
    float *rvir, *mvir;
    int32_t *tag, *nsub, *nign;
    int32_t *counters;                    /* [0] groups removed, [1] groups slurped, [2] error */
    unsigned char *do_vcirc;              /* per slot: still valid right after its own walk (kd2.c:884) */
};

__global__ void __launch_bounds__(256) k_tag_replay(const __grid_constant__ ReplayArgs a)
{
    __shared__ int s_first;               /* first subsume / slurp event of the chunk */
    __shared__ int s_kind, s_other;
    const int t = threadIdx.x;
    for (int it = 0; it < a.n_order; ++it) {
        const int A = a.order[it];
        const unsigned long long a0 = a.off[A], a1 = a.off[A + 1];
        const int32_t idA = a.index[A];
        const float ax = a.pos[3 * A], ay = a.pos[3 * A + 1], az = a.pos[3 * A + 2];
        bool slurped = false;
        unsigned long long k0 = a0;
        while (k0 < a1 && !slurped) {
            const float rA = a.rvir[A];
            const float rA2 = __fmul_rn(rA, rA);
            if (t == 0) s_first = 0x7FFFFFFF;
            __syncthreads();
            const unsigned long long k = k0 + (unsigned long long)t;
            int32_t p = -1, tg = 0;
            int kind = 0, other = -1;     /* 0 untagged, 1 ignore, 2 subsume, 3 slurp */
            if (k < a1) {
                p = a.mem[k];
                tg = a.tag[p];
                if (tg != 0) {
                    other = a.slot_of_index[tg];
                    const float dx = __fsub_rn(ax, a.pos[3 * other]), dy = __fsub_rn(ay, a.pos[3 * other + 1]),
                                dz = __fsub_rn(az, a.pos[3 * other + 2]);
                    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));   /* kd2.c:677-680 */
                    const float rB = a.rvir[other];
                    if (r2 <= rA2) kind = 2;
                    else if (r2 <= __fmul_rn(rB, rB)) kind = 3;
                    else kind = 1;
                    if (kind >= 2) atomicMin(&s_first, t);
                }
            }
            __syncthreads();
            const int first = s_first;
            if (k < a1 && t < first) {                        /* everything in front of the event: independent */
                if (kind == 0) a.tag[p] = idA;
                else ++a.nign[p];                             /* (a particle appears once in this list) */
            }
            if (t == first) { s_kind = kind; s_other = other; }
            __syncthreads();
            if (first == 0x7FFFFFFF) { k0 += 256ull; continue; }
            const int ev_kind = s_kind, B = s_other;
            const int loser = ev_kind == 2 ? B : A, winner = ev_kind == 2 ? A : B;
            if (t == 0) {                                     /* kdZeroGroup's bookkeeping (kd2.c:617-634) */
                if (a.mvir[loser] < 0.0f) a.counters[2] = 1;
                a.rvir[loser] = (float)(-10.0 * (double)a.index[winner]);
                a.mvir[loser] = -a.mvir[loser];
                ++a.counters[ev_kind == 2 ? 0 : 1];
            }
            {
                const int32_t idL = a.index[loser];
                const unsigned long long l0 = a.off[loser], l1 = a.off[loser + 1];
                for (unsigned long long q = l0 + (unsigned long long)t; q < l1; q += 256ull) {
                    const int32_t pp = a.mem[q];
                    if (a.tag[pp] == idL) { a.tag[pp] = 0; ++a.nsub[pp]; }
                }
            }
            __syncthreads();
            if (ev_kind == 2) {
                if (t == first) a.tag[p] = idA;               /* kd2.c:691: the particle goes to the subsuming group */
                k0 += (unsigned long long)first + 1ull;       /* resume behind the event: the loser's particles are free now */
            } else {
                slurped = true;                               /* kd2.c:671: nothing after the slurp */
            }
            __threadfence_block();
            __syncthreads();
        }
        if (t == 0) a.do_vcirc[A] = a.rvir[A] > 0.0f ? 1 : 0;
        __syncthreads();
    }
}

/* _VcmParticles (kd2.c:595-609): vcm[l] = (sum over the members, in sorted order, of fl(m * v[l])) / Mvir,
 * a sequential fp32 sum per group.  One thread per group walks its (r^2, index)-sorted member list; the
 * loads of 8 members are issued together, the adds stay in list order. */
__global__ void __launch_bounds__(128) k_vcm(const unsigned long long *__restrict__ off, const int32_t *__restrict__ mem,
                                             const float4 *__restrict__ in, const float4 *__restrict__ vel,
                                             const float *__restrict__ mvir, int nh, float *__restrict__ vcm)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    const unsigned long long a = off[h], b = off[h + 1];
    float vx = 0.0f, vy = 0.0f, vz = 0.0f;
    for (unsigned long long k0 = a; k0 < b; k0 += 8) {
        float4 v[8];
        float m[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (k0 + u < b) {
                const int32_t p = __ldg(mem + k0 + u);
                v[u] = __ldg(vel + p);
                m[u] = __ldg(&in[p].w);
            }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (k0 + u < b) {
                vx = __fadd_rn(vx, __fmul_rn(m[u], v[u].x));
                vy = __fadd_rn(vy, __fmul_rn(m[u], v[u].y));
                vz = __fadd_rn(vz, __fmul_rn(m[u], v[u].z));
            }
    }
    const float mv = mvir[h];
    vcm[3 * h + 0] = __fdiv_rn(vx, mv);
    vcm[3 * h + 1] = __fdiv_rn(vy, mv);
    vcm[3 * h + 2] = __fdiv_rn(vz, mv);
}

/* member lists -> sortable keys and back: ascending (fDist2 bits, original index) is the order the
 * reference's qsort(CmpList) + stable merge gives for distinct r^2 (kd2.c:425-435,781) */
__global__ void __launch_bounds__(256) k_member_keys(const int32_t *__restrict__ idx, const float *__restrict__ d2,
                                                     unsigned long long n, unsigned long long *__restrict__ keys)
{
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        keys[i] = ((unsigned long long)__float_as_uint(d2[i]) << 32) | (uint32_t)idx[i];
}
__global__ void __launch_bounds__(256) k_member_unkeys(const unsigned long long *__restrict__ keys, unsigned long long n,
                                                       int32_t *__restrict__ idx, float *__restrict__ d2)
{
    unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        unsigned long long k = keys[i];
        idx[i] = (int32_t)(uint32_t)k;
        d2[i] = __uint_as_float((uint32_t)(k >> 32));
    }
}

/* ============================================================================================
 * host side
 * ============================================================================================ */
enum {
    KID_LVL_HIST = 0, KID_LVL_SCAN, KID_LVL_PARTITION, KID_BUCKET_SORT, KID_MASS_TABLE, KID_CLASSIFY,
    KID_QUERY_WARP, KID_QUERY_BLOCK, KID_OFFSETS, KID_EMIT_WARP, KID_EMIT_BLOCK, KID_BALL_GATHER,
    KID_QUERY_HUGE, KID_EMIT_HUGE, KID_MARK_MASK, KID_VCIRC, KID_TAG, KID_ROUTE, KID_ASSIGN, KID_PUSH, KID_BARRIER, KID_QUERY_FUSED, KID_SEGSORT,
    KID_ROUTE_SPLIT, KID_BUCKET_LIVE, KID_BUCKET_BIG, KID_N
};
static const char *const g_kernel_names[KID_N] = {
    "k_lvl_hist", "k_scan", "k_lvl_partition", "k_bucket_sort", "k_mass_table", "k_classify",
    "k_so_query<32>", "k_so_query<256>", "k_offsets", "k_so_emit<32>", "k_so_emit<256>", "k_ball_gather",
    "k_so_query<1024>", "k_so_emit<1024>", "k_mark_mask", "k_vcirc", "k_tag_claim+settle", "k_route",
    "k_assign", "k_push", "k_dom_barrier", "k_so_query_fused", "k_segsort", "k_route_split", "k_bucket_live", "k_bucket_sort(big)"};

struct ProfRec { int kid, launches; cudaEvent_t a, b; };

struct sogpu {
    int device;
    cudaStream_t own_stream, stream;
    cudaStream_t aux[2];             /* side streams: the three halo-size classes run concurrently */
    cudaStream_t launch_stream;      /* where ProfScope / launch_persistent currently enqueue */
    cudaEvent_t ev_fork, ev_join[2];
    float ppc;                       /* target particles per cell */
    int pack_threads;

    int64_t n;
    float period[3], center[3];
    const float4 *d_in;              /* unsorted particles (owned or borrowed) */
    float4 *d_in_owned;
    int64_t d_in_cap;
    float4 *d_sorted;
    int64_t grid_n_cap;
    uint32_t *d_ce;
    uint32_t *d_bsum;
    uint32_t *d_massmm;
    float *d_raw;                    /* staging of raw xyz triplets (pinned-host fast path) */
    void *d_ingest[2];               /* streaming ingest: raw record chunks */
    float4 *d_vel;                   /* velocities {vx,vy,vz,0} in file order, when the ingest was asked to keep them */
    bool ingest_want_vel;
    size_t ingest_cap[2];
    cudaEvent_t ingest_ev[2];
    int64_t ingest_done;
    int ingest_slot;
    bool ingest_prev_ev_valid;
    float4 *d_tmp4;                  /* ping-pong payload buffer of the partition levels */
    uint32_t *d_key[2];              /* ping-pong cell keys between levels */
    int64_t tmp_cap;
    uint32_t *d_lvl_start[4];        /* child-bucket starts per level (+ sentinel) */
    uint32_t *d_lvl_cursor[4];       /* counts, then the atomic cursors of the partition */
    size_t lvl_cap[4];
    float cls_small_max, cls_huge_min;   /* expected ball population: warp / 256-thread CTA / 1024-thread CTA */
    int emit_small_max, emit_huge_min;   /* same split for the member emission, by N_Delta */
    int qgrid32, qgrid256;           /* persistent CTAs per SM of the warp / 256-thread halo kernels */
    int qorder;                      /* launch order of the classes behind the 1024-thread one (tuning) */
    int fused_query;                 /* 1: warp and 256-thread classes in one persistent kernel (default) */
    size_t scan1_max;                /* bucket tables up to this many entries are scanned by one block */
    bool use_tma;                    /* sogpu_set_tma_staging */
    double mask_rmin_cells;          /* focus masks: minimum half-width per halo, in coarse cells */
    int mask_bits;                   /* focus masks: at most 2^mask_bits coarse cells per axis */
    const uint32_t *mask_ready;      /* build_grid_impl(focus_nh < 0): this mask, prepared by the caller */
    bool indexed;                    /* d_in is {x,y,z,global index} of one rank's share (domain runs) */
    float indexed_mass;
    int64_t n_total;                 /* particles of the whole snapshot (grid resolution of a domain run) */
    int first_ball;                  /* first ball of the schedule that is gathered (1 = as the reference) */
    int two_level;                   /* -1 auto; 0: no partition levels (bucket sort only if it fits) */
    uint32_t *d_mask;                /* focused build: 2^(3*mb) bits */
    bool focused;                    /* the current grid holds only the focused region */
    int focus_balls;
    double prof_bytes[32];
    int64_t ncell;
    int nc, lb;
    bool built;
    int mass_state;                  /* -1 unknown (not fetched yet), 0 unequal, 1 equal */
    float mass;
    GridDev g;
    so_mass_table *d_mt;

    /* query buffers */
    int32_t cap_h;
    float *d_centers, *d_rgtp;
    int32_t *d_small, *d_big, *d_esmall, *d_ebig, *d_huge, *d_ehuge, *d_defer;
    uint32_t *d_counters;   /* 0 small_n 1 big_n 2 work_small 3 work_big 4 flags 5 esmall_n 6 ebig_n 7 work_es 8 work_eb */
    int32_t *d_out_n;
    float *d_out_m;
    unsigned long long *d_out_key, *d_out_off;
    unsigned long long *d_u64;       /* [0] member_total [1] evals_hist [2] evals_other [3] gather cnt */
    int32_t *d_members;
    float *d_md2;
    unsigned long long member_cap;
    int32_t last_h;
    bool have_result;
    bool want_d2;
    bool member_overflow;
    bool build_attr_done;
    unsigned long long *d_timeline;  /* SOGPU_DEBUG_TIMELINE */
    uint32_t *d_live;                /* focused builds: list of live final buckets (+ its length) */
    bool bucket_warp;                /* focused builds: small final buckets are sorted one per warp */
    size_t live_cap;
    unsigned long long *d_route;     /* domain runs: per-destination counters */
    unsigned short *d_route_table;   /* destination ranks per coarse cell */
    uint32_t *d_route_any;
    int32_t *d_tag, *d_tag_index;    /* sogpu_tag_members: owner per particle, catalog ids */
    int32_t *d_nsub, *d_nign;        /* sogpu_tag_replay: PINIT.nSubsumed / nIgnored */
    int64_t replay_cap;
    void *d_replay;                  /* ... and its per-group arrays */
    size_t replay_bytes;
    unsigned char *d_dirty;
    int64_t tag_cap;
    int32_t tagh_cap;
    float *d_vc;                     /* sogpu_vcirc: per-group inputs and outputs */
    size_t vc_cap;
    unsigned char *d_ptype;          /* sogpu_vcirc_species: species bits per particle */
    int64_t ptype_cap;
    float *d_vcC;                    /* ... and the sequential mass prefixes */
    size_t vcC_cap;
    bool members_sorted;             /* the device member lists are already in (r^2, index) order */
    float last_thr;                  /* parameters of the last solve (its centres / radii are in d_centers / d_rgtp) */
    int32_t last_nM;
    bool members_csr;                /* d_out_off / d_members are in catalog (CSR) order; false: one slot per halo in solve order */
    int32_t *d_members2;             /* second member buffers: target of the compaction into catalog order */
    float *d_md2_2;
    unsigned long long *d_csr_off;

    /* general (unequal-mass) path */
    GenState *d_gen_state;
    int32_t *d_gen_list[2], *d_gen_round;
    unsigned long long *d_seg_n, *d_seg_begin, *d_seg_end;
    unsigned long long *d_gkeys[2];
    float *d_gmass[2];
    size_t gen_cap;                  /* entries of the (key, mass) scratch */
    uint32_t *d_ss_tiles;            /* segmented sort: tile_base[nseg + 1] */
    unsigned long long *d_ss_max, *d_ss_off;
    int32_t ss_cap;
    int32_t gen_cap_h;

    /* pinned host staging */
    void *h_pin;
    size_t h_pin_bytes;
    int32_t *h_members;
    float *h_md2;
    size_t h_members_cap;

    /* profiling */
    bool prof_on;
    std::vector<ProfRec> prof_pending;
    std::vector<cudaEvent_t> prof_pool;
    double prof_ms[KID_N];
    int64_t prof_launches[KID_N];

    sogpu_stats_t stats;
    int sm_count;

    /* device-side particle count of the next build (domain steps: what arrived is only known on the device) */
    const uint32_t *d_n_dev;         /* NULL: h->n is exact */
    int64_t n_hint;                  /* expected count: sizes the launch and the bucket table */
    const unsigned char *q_owner;    /* query only the halos with q_owner[h] == q_me (NULL: all) */
    int q_me, q_ranks;
    struct DomainState *dom;         /* sogpu_domain_open */
};

static cudaEvent_t prof_event(sogpu *h)
{
    cudaEvent_t e = nullptr;
    if (!h->prof_pool.empty()) { e = h->prof_pool.back(); h->prof_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}
struct ProfScope {   /* brackets one (group of) kernel launch(es) with events when profiling is on */
    sogpu *h; ProfRec r; bool on;
    ProfScope(sogpu *h_, int kid, double alg_bytes = 0.0, int launches = 1) : h(h_), on(h_->prof_on)
    {
        h->stats.last_kernel_launches += launches;
        if (!on) return;
        h->prof_bytes[kid] += alg_bytes;
        r.kid = kid; r.launches = launches; r.a = prof_event(h); r.b = prof_event(h);
        cudaEventRecord(r.a, h->launch_stream);
    }
    ~ProfScope()
    {
        if (!on) return;
        cudaEventRecord(r.b, h->launch_stream);
        h->prof_pending.push_back(r);
    }
};

/* SOGPU_DEBUG_TIMING=1: wall-clock marks inside the longer host-side entry points (stderr) */
struct DbgTimer {
    bool on; double t0; cudaStream_t s;
    static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }
    explicit DbgTimer(cudaStream_t s_) : on(getenv("SOGPU_DEBUG_TIMING") != nullptr), t0(0), s(s_) { if (on) t0 = now(); }
    void mark(const char *what) { if (!on) return; cudaStreamSynchronize(s); double t = now(); fprintf(stderr, "    [sogpu] %-28s %9.3f ms\n", what, t - t0); t0 = t; }
};

static int ensure_pinned(sogpu *h, size_t bytes)
{
    if (bytes <= h->h_pin_bytes) return SOGPU_OK;
    if (h->h_pin) cudaFreeHost(h->h_pin);
    h->h_pin = nullptr; h->h_pin_bytes = 0;
    CU(cudaMallocHost(&h->h_pin, bytes));
    h->h_pin_bytes = bytes;
    return SOGPU_OK;
}

extern "C" int sogpu_create(sogpu_t **out, int device)
{
    if (!out) return set_err(SOGPU_ERR_ARG, "sogpu_create: out is NULL");
    *out = nullptr;
    if (SO_C133PI != 1.33333333 * M_PI || SO_C43PI != (4. / 3.) * M_PI)
        return set_err(SOGPU_ERR_UNSUPPORTED, "so_math.h constants do not match this compiler's folding");
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (ndev <= 0) return set_err(SOGPU_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    if (device < 0) CU(cudaGetDevice(&device));
    if (device >= ndev) return set_err(SOGPU_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    CU(cudaSetDevice(device));
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_err(SOGPU_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                       device, prop.major, prop.minor);
    sogpu *h = new (std::nothrow) sogpu();
    if (!h) return set_err(SOGPU_ERR_NOMEM, "out of host memory");
    h->device = device;
    h->ppc = 2.0f;
    h->pack_threads = 8;
    h->mass_state = -1;
    h->two_level = -1;
    h->first_ball = 2;
    h->mask_rmin_cells = 0.75;
    h->mask_bits = 9;
    if (const char *e = getenv("SOGPU_MASK_BITS")) h->mask_bits = std::min(9, std::max(4, atoi(e)));
    h->scan1_max = (size_t)1 << 15;          /* (one block scans 2^18 entries in ~80 us, three small launches in ~15) */
    h->qgrid32 = 3; h->qgrid256 = 2; h->qorder = 0;    /* persistent grids of exactly the CTAs that can be resident (pending CTAs of one class hold back the next kernel's) */
    if (const char *e = getenv("SOGPU_QGRID32")) h->qgrid32 = std::max(1, atoi(e));
    if (const char *e = getenv("SOGPU_QGRID256")) h->qgrid256 = std::max(1, atoi(e));
    if (const char *e = getenv("SOGPU_QORDER")) h->qorder = atoi(e);
    h->fused_query = 1;
    if (const char *e = getenv("SOGPU_FUSED")) h->fused_query = atoi(e) != 0;
    if (const char *e = getenv("SOGPU_SCAN1_MAX")) h->scan1_max = (size_t)atoll(e);
    if (const char *e = getenv("SOGPU_TMA")) h->use_tma = atoi(e) != 0;
    if (const char *e = getenv("SOGPU_MASK_RMIN")) h->mask_rmin_cells = atof(e);
    h->bucket_warp = true;
    if (const char *e = getenv("SOGPU_BUCKET_WARP")) h->bucket_warp = atoi(e) != 0;
    h->cls_small_max = 256.0f; h->cls_huge_min = 32768.0f;   /* warp class: balls that fit its 384 staged keys; above 32768: a
                                                              * 1024-thread CTA per halo, beside the fused kernel (8192 was
                                                              * measured 5 % slower at 1024^3: the big CTAs crowd out the fused ones) */
    h->emit_small_max = 2048; h->emit_huge_min = 4096;
    if (const char *e = getenv("SOGPU_SMALL_MAX")) h->cls_small_max = (float)atof(e);
    if (const char *e = getenv("SOGPU_HUGE_MIN")) h->cls_huge_min = (float)atof(e);
    if (const char *e = getenv("SOGPU_EMIT_SMALL_MAX")) h->emit_small_max = atoi(e);
    if (const char *e = getenv("SOGPU_EMIT_HUGE_MIN")) h->emit_huge_min = atoi(e);
    if (const char *e = getenv("SOGPU_FIRST_BALL")) h->first_ball = std::max(1, atoi(e));
    if (const char *e = getenv("SOGPU_BUILD_MODE")) h->two_level = atoi(e);   /* A/B knob, see sogpu_set_build_mode */
    h->sm_count = prop.multiProcessorCount;
    cudaError_t e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { delete h; return set_err(SOGPU_ERR_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    h->stream = h->launch_stream = h->own_stream;
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        /* aux[0] (256-thread class) is served before aux[1] (warp class) when both have CTAs pending */
        e = cudaStreamCreateWithPriority(&h->aux[k], cudaStreamNonBlocking,
                                         (k == 0 && !getenv("SOGPU_NO_PRIO")) ? prio_greatest : prio_least);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_so_query<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<256>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_so_query<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<1024>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_so_emit<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<1024>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_so_query<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<32>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_so_query_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_so_emit<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<256>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_so_emit<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<32>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_ball_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<32>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_gen_count, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<256>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k_gen_emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)query_smem_bytes<256>());
    /* The three halo classes must be able to SHARE an SM.  An SM's L1 / shared-memory split is fixed while
     * it has resident CTAs, so kernels that ask for different splits exclude each other until it drains
     * (measured with %globaltimer: whichever class started first kept the others out for 70-120 us).
     * Every query / emit kernel therefore asks for the same (maximum shared memory) carve-out. */
    {
        const void *fns[] = {(const void *)k_so_query<32>, (const void *)k_so_query<256>, (const void *)k_so_query<1024>, (const void *)k_so_query_fused,
                             (const void *)k_so_emit<32>, (const void *)k_so_emit<256>, (const void *)k_so_emit<1024>};
        for (const void *f : fns)
            if (e == cudaSuccess)
                e = cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (e != cudaSuccess) {
        cudaStreamDestroy(h->own_stream); delete h;
        return set_err(SOGPU_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    }
    *out = h;
    g_err[0] = 0;
    return SOGPU_OK;
}

static void free_grid(sogpu *h)
{
    cudaFree(h->d_sorted); h->d_sorted = nullptr;
    cudaFree(h->d_ce); h->d_ce = nullptr;
    cudaFree(h->d_bsum); h->d_bsum = nullptr;
    h->grid_n_cap = 0; h->nc = 0;
    h->built = false;
}

static void free_query(sogpu *h)
{
    cudaFree(h->d_centers); cudaFree(h->d_rgtp); cudaFree(h->d_small); cudaFree(h->d_big);
    cudaFree(h->d_esmall); cudaFree(h->d_ebig); cudaFree(h->d_huge); cudaFree(h->d_ehuge); cudaFree(h->d_defer);
    cudaFree(h->d_out_n); cudaFree(h->d_out_m); cudaFree(h->d_out_key); cudaFree(h->d_out_off); cudaFree(h->d_csr_off);
    h->d_csr_off = nullptr;
    h->d_centers = h->d_rgtp = nullptr; h->d_small = h->d_big = h->d_esmall = h->d_ebig = nullptr;
    h->d_huge = h->d_ehuge = h->d_defer = nullptr;
    h->d_out_n = nullptr; h->d_out_m = nullptr; h->d_out_key = h->d_out_off = nullptr;
    h->cap_h = 0;
}

extern "C" int sogpu_domain_close(sogpu_t *h);

extern "C" void sogpu_destroy(sogpu_t *h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    sogpu_domain_close(h);
    free_grid(h);
    free_query(h);
    cudaFree(h->d_in_owned);
    cudaFree(h->d_massmm);
    cudaFree(h->d_raw);
    cudaFree(h->d_ingest[0]); cudaFree(h->d_ingest[1]); cudaFree(h->d_vel);
    for (int k = 0; k < 2; ++k) if (h->ingest_ev[k]) cudaEventDestroy(h->ingest_ev[k]);
    cudaFree(h->d_mask);
    cudaFree(h->d_tmp4); cudaFree(h->d_key[0]); cudaFree(h->d_key[1]);
    for (int l = 0; l < 4; ++l) { cudaFree(h->d_lvl_start[l]); cudaFree(h->d_lvl_cursor[l]); }
    cudaFree(h->d_mt);
    cudaFree(h->d_vc); cudaFree(h->d_ptype); cudaFree(h->d_vcC);
    cudaFree(h->d_route); cudaFree(h->d_route_table); cudaFree(h->d_route_any);
    cudaFree(h->d_live); cudaFree(h->d_timeline);
    cudaFree(h->d_tag); cudaFree(h->d_tag_index); cudaFree(h->d_dirty); cudaFree(h->d_nsub); cudaFree(h->d_nign); cudaFree(h->d_replay);
    cudaFree(h->d_counters);
    cudaFree(h->d_u64);
    cudaFree(h->d_members);
    cudaFree(h->d_md2);
    cudaFree(h->d_members2);
    cudaFree(h->d_md2_2);
    cudaFree(h->d_gen_state); cudaFree(h->d_gen_list[0]); cudaFree(h->d_gen_list[1]); cudaFree(h->d_gen_round);
    cudaFree(h->d_seg_n); cudaFree(h->d_seg_begin); cudaFree(h->d_seg_end);
    cudaFree(h->d_gkeys[0]); cudaFree(h->d_gkeys[1]); cudaFree(h->d_gmass[0]); cudaFree(h->d_gmass[1]);
    cudaFree(h->d_ss_tiles); cudaFree(h->d_ss_max); cudaFree(h->d_ss_off);
    if (h->h_pin) cudaFreeHost(h->h_pin);
    if (h->h_members) cudaFreeHost(h->h_members);
    if (h->h_md2) cudaFreeHost(h->h_md2);
    for (auto &r : h->prof_pending) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : h->prof_pool) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) {
        if (h->aux[k]) cudaStreamDestroy(h->aux[k]);
        if (h->ev_join[k]) cudaEventDestroy(h->ev_join[k]);
    }
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    cudaStreamDestroy(h->own_stream);
    delete h;
}

extern "C" int sogpu_set_stream(sogpu_t *h, void *s)
{
    if (!h) return set_err(SOGPU_ERR_ARG, "NULL handle");
    h->stream = h->launch_stream = s ? (cudaStream_t)s : h->own_stream;
    return SOGPU_OK;
}

extern "C" int sogpu_set_build_mode(sogpu_t *h, int mode)
{
    if (!h || mode < -1 || mode > 2) return set_err(SOGPU_ERR_ARG, "bad build mode");
    h->two_level = mode;
    return SOGPU_OK;
}

extern "C" int sogpu_set_first_ball(sogpu_t *h, int k)
{
    if (!h || k < 1 || k > 64) return set_err(SOGPU_ERR_ARG, "bad first ball");
    h->first_ball = k;
    return SOGPU_OK;
}

extern "C" int sogpu_set_tma_staging(sogpu_t *h, int on)
{
    if (!h) return set_err(SOGPU_ERR_ARG, "NULL handle");
    h->use_tma = on != 0;
    h->g.use_tma = h->use_tma ? 1 : 0;
    return SOGPU_OK;
}

extern "C" int sogpu_set_cell_occupancy(sogpu_t *h, float ppc)
{
    if (!h || !(ppc > 0.0f)) return set_err(SOGPU_ERR_ARG, "bad cell occupancy");
    h->ppc = ppc;
    return SOGPU_OK;
}

static int set_common(sogpu *h, int64_t n, const float period[3], const float center[3])
{
    if (n <= 0 || n > 0x7FFFFFF0LL) return set_err(SOGPU_ERR_ARG, "particle count %lld out of range", (long long)n);
    for (int k = 0; k < 3; ++k)
        if (!(period[k] > 0.0f) || !(period[k] < INFINITY))
            return set_err(SOGPU_ERR_ARG, "period[%d] must be positive (the reference is periodic only, smooth2.c:69)", k);
    CU(cudaSetDevice(h->device));
    h->n = n;
    for (int k = 0; k < 3; ++k) { h->period[k] = period[k]; h->center[k] = center ? center[k] : 0.0f; }
    h->built = false;
    h->have_result = false;
    h->mass_state = -1;
    h->indexed = false;
    h->n_total = n;
    return SOGPU_OK;
}

extern "C" int sogpu_set_particles_device(sogpu_t *h, const void *d_xyzm, int64_t n, const float period[3],
                                          const float center[3])
{
    if (!h || !d_xyzm || !period) return set_err(SOGPU_ERR_ARG, "sogpu_set_particles_device: NULL argument");
    int rc = set_common(h, n, period, center);
    if (rc) return rc;
    h->d_in = (const float4 *)d_xyzm;
    return SOGPU_OK;
}

static void pack_range(float4 *dst, const char *pp, size_t ps, const char *mp, size_t ms, int64_t k)
{
    for (int64_t i = 0; i < k; ++i) {
        const float *p = (const float *)(pp + (size_t)i * ps);
        dst[i].x = p[0]; dst[i].y = p[1]; dst[i].z = p[2];
        dst[i].w = *(const float *)(mp + (size_t)i * ms);
    }
}

/* pack host particles to float4 {x,y,z,m} with a few host threads into two pinned staging buffers
 * (the H2D copy of one chunk overlaps the packing of the next) and copy them to d_dst */
static int upload_to(sogpu *h, float4 *d_dst, const void *pos, size_t pos_stride, const void *mass,
                     size_t mass_stride, int64_t n)
{
    /* fast paths for page-locked caller memory: DMA straight from it, unpack on the device */
    cudaPointerAttributes pa;
    bool pinned = cudaPointerGetAttributes(&pa, pos) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned && pos_stride == sizeof(float4) && mass_stride == sizeof(float4) &&
        (const char *)mass == (const char *)pos + 3 * sizeof(float)) {
        CU(cudaMemcpyAsync(d_dst, pos, (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        return SOGPU_OK;
    }
    if (pinned && pos_stride == 3 * sizeof(float) && mass_stride == 0) {
        const int64_t rchunk = 1 << 23;   /* 96 MB of xyz per piece */
        if (!h->d_raw) CU(cudaMalloc(&h->d_raw, (size_t)rchunk * 3 * sizeof(float)));
        const float m = *(const float *)mass;
        for (int64_t i0 = 0; i0 < n; i0 += rchunk) {
            int64_t k = std::min(rchunk, n - i0);
            CU(cudaMemcpyAsync(h->d_raw, (const float *)pos + 3 * i0, (size_t)k * 3 * sizeof(float),
                               cudaMemcpyHostToDevice, h->stream));
            k_expand_xyz<<<h->sm_count * 8, 256, 0, h->stream>>>(h->d_raw, m, d_dst + i0, k);
        }
        CU(cudaGetLastError());
        CU(cudaStreamSynchronize(h->stream));
        return SOGPU_OK;
    }
    const int64_t chunk = 1 << 22;
    int rc = ensure_pinned(h, 2 * (size_t)chunk * sizeof(float4));
    if (rc) return rc;
    float4 *stage[2] = {(float4 *)h->h_pin, (float4 *)h->h_pin + chunk};
    cudaEvent_t ev[2];
    CU(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    int b = 0;
    cudaError_t e = cudaSuccess;
    for (int64_t i0 = 0; i0 < n && e == cudaSuccess; i0 += chunk, b ^= 1) {
        int64_t k = std::min(chunk, n - i0);
        e = cudaEventSynchronize(ev[b]);
        if (e != cudaSuccess) break;
        float4 *s = stage[b];
        const char *pp = (const char *)pos + (size_t)i0 * pos_stride;
        const char *mp = (const char *)mass + (size_t)i0 * mass_stride;
        int nt = (k >= (1 << 16)) ? h->pack_threads : 1;
        if (nt <= 1) {
            pack_range(s, pp, pos_stride, mp, mass_stride, k);
        } else {
            std::vector<std::thread> th;
            int64_t per = (k + nt - 1) / nt;
            for (int t = 0; t < nt; ++t) {
                int64_t a0 = t * per, a1 = std::min(k, a0 + per);
                if (a0 >= a1) break;
                th.emplace_back(pack_range, s + a0, pp + (size_t)a0 * pos_stride, pos_stride,
                                mp + (size_t)a0 * mass_stride, mass_stride, a1 - a0);
            }
            for (auto &t : th) t.join();
        }
        e = cudaMemcpyAsync(d_dst + i0, s, (size_t)k * sizeof(float4), cudaMemcpyHostToDevice, h->stream);
        if (e == cudaSuccess) e = cudaEventRecord(ev[b], h->stream);
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaEventDestroy(ev[0]); cudaEventDestroy(ev[1]);
    if (e != cudaSuccess) return set_err(SOGPU_ERR_CUDA, "particle upload failed: %s", cudaGetErrorString(e));
    return SOGPU_OK;
}

extern "C" int sogpu_upload_particles(sogpu_t *h, const void *pos, size_t pos_stride, const void *mass,
                                      size_t mass_stride, int64_t n, void *d_xyzm_dst)
{
    if (!h || !pos || !mass || !d_xyzm_dst || n <= 0) return set_err(SOGPU_ERR_ARG, "sogpu_upload_particles: bad argument");
    CU(cudaSetDevice(h->device));
    return upload_to(h, (float4 *)d_xyzm_dst, pos, pos_stride, mass, mass_stride, n);
}

extern "C" int sogpu_set_particles_host(sogpu_t *h, const void *pos, size_t pos_stride, const void *mass,
                                        size_t mass_stride, int64_t n, const float period[3],
                                        const float center[3])
{
    if (!h || !pos || !mass || !period) return set_err(SOGPU_ERR_ARG, "sogpu_set_particles_host: NULL argument");
    int rc = set_common(h, n, period, center);
    if (rc) return rc;
    if (n > h->d_in_cap) {
        cudaFree(h->d_in_owned); h->d_in_owned = nullptr; h->d_in_cap = 0;
        CU(cudaMalloc(&h->d_in_owned, (size_t)n * sizeof(float4)));
        h->d_in_cap = n;
    }
    h->d_in = h->d_in_owned;
    return upload_to(h, h->d_in_owned, pos, pos_stride, mass, mass_stride, n);
}

/* ---- streaming ingest of raw TIPSY records (replaces the PINIT fill of kdReadTipsy, kd2.c:352-416) ---- */

extern "C" void *sogpu_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

extern "C" void sogpu_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

extern "C" int sogpu_ingest_begin(sogpu_t *h, int64_t n_total, const float period[3], const float center[3])
{
    if (!h || !period) return set_err(SOGPU_ERR_ARG, "sogpu_ingest_begin: NULL argument");
    int rc = set_common(h, n_total, period, center);
    if (rc) return rc;
    if (n_total > h->d_in_cap) {
        cudaFree(h->d_in_owned); h->d_in_owned = nullptr; h->d_in_cap = 0;
        CU(cudaMalloc(&h->d_in_owned, (size_t)n_total * sizeof(float4)));
        h->d_in_cap = n_total;
    }
    h->d_in = nullptr;                     /* set by sogpu_ingest_end */
    cudaFree(h->d_vel); h->d_vel = nullptr;
    if (h->ingest_want_vel) CU(cudaMalloc(&h->d_vel, (size_t)n_total * sizeof(float4)));
    h->ingest_done = 0;
    h->ingest_slot = 0;
    h->ingest_prev_ev_valid = false;
    for (int k = 0; k < 2; ++k)
        if (!h->ingest_ev[k]) CU(cudaEventCreateWithFlags(&h->ingest_ev[k], cudaEventDisableTiming));
    return SOGPU_OK;
}

extern "C" int sogpu_ingest_records(sogpu_t *h, const void *records, int64_t count, int32_t floats_per_record,
                                    int32_t big_endian)
{
    if (!h || !records || count < 0 || floats_per_record < 4)
        return set_err(SOGPU_ERR_ARG, "sogpu_ingest_records: bad argument");
    if (count == 0) return SOGPU_OK;
    if (h->ingest_done + count > h->n) return set_err(SOGPU_ERR_ARG, "sogpu_ingest_records: more records than announced");
    CU(cudaSetDevice(h->device));
    const size_t bytes = (size_t)count * (size_t)floats_per_record * sizeof(float);
    const int slot = h->ingest_slot;
    /* the previous chunk (other slot) is on the device and unpacked: its host buffer is free again, and so
     * is this slot's device staging (used two chunks ago, finished before the previous one in stream order) */
    if (h->ingest_prev_ev_valid) CU(cudaEventSynchronize(h->ingest_ev[slot ^ 1]));
    if (bytes > h->ingest_cap[slot]) {
        cudaFree(h->d_ingest[slot]); h->d_ingest[slot] = nullptr; h->ingest_cap[slot] = 0;
        CU(cudaMalloc(&h->d_ingest[slot], bytes));
        h->ingest_cap[slot] = bytes;
    }
    cudaPointerAttributes pa;
    const bool pinned = cudaPointerGetAttributes(&pa, records) == cudaSuccess && pa.type == cudaMemoryTypeHost;
    cudaGetLastError();
    cudaStream_t s = h->stream;
    /* page-locked source: asynchronous DMA; pageable source: the call returns when the driver has staged
     * the data (the caller's buffer is free again), the copy itself stays ordered before the unpack kernel */
    (void)pinned;
    CU(cudaMemcpyAsync(h->d_ingest[slot], records, bytes, cudaMemcpyHostToDevice, s));
    const int grid = (int)std::min<int64_t>((count + 255) / 256, (int64_t)h->sm_count * 8);
    if (h->d_vel && floats_per_record < 7) return set_err(SOGPU_ERR_ARG, "sogpu_ingest_records: records without velocity fields");
    k_ingest_records<<<grid, 256, 0, s>>>((const uint32_t *)h->d_ingest[slot], count, floats_per_record, big_endian ? 1 : 0,
                                          h->d_in_owned + h->ingest_done, h->d_vel ? h->d_vel + h->ingest_done : nullptr);
    CU(cudaGetLastError());
    CU(cudaEventRecord(h->ingest_ev[slot], s));
    h->ingest_prev_ev_valid = true;
    h->ingest_done += count;
    h->ingest_slot = slot ^ 1;
    return SOGPU_OK;
}

extern "C" int sogpu_ingest_end(sogpu_t *h)
{
    if (!h) return set_err(SOGPU_ERR_ARG, "NULL handle");
    if (h->ingest_done != h->n)
        return set_err(SOGPU_ERR_ARG, "sogpu_ingest_end: %lld of %lld records received", (long long)h->ingest_done, (long long)h->n);
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    h->d_in = h->d_in_owned;
    return SOGPU_OK;
}

static int pick_cells(int64_t n, float ppc, int *lb)
{
    /* nc = power of two with nc^3 closest (in log) to n/ppc, 4 <= nc <= 1024 */
    double target = cbrt((double)n / (double)ppc);
    int l = (int)floor(log2(target) + 0.5);
    if (l < 2) l = 2;
    if (l > 10) l = 10;
    *lb = l;
    return 1 << l;
}

/* the running-mass table needs only the mass min/max of the level-0 histogram pass: a one-thread
 * kernel on a side stream, overlapped with the partition passes, joined at the end of the build */
static int launch_mass_table(sogpu *h)
{
    CU(cudaEventRecord(h->ev_fork, h->stream));
    CU(cudaStreamWaitEvent(h->aux[0], h->ev_fork, 0));
    cudaStream_t keep = h->launch_stream;
    h->launch_stream = h->aux[0];
    { ProfScope p(h, KID_MASS_TABLE); k_mass_table<<<1, 32, 0, h->aux[0]>>>(h->d_massmm, h->d_mt, (unsigned long long)h->n + 2ull); }
    h->launch_stream = keep;
    CU(cudaEventRecord(h->ev_join[0], h->aux[0]));
    return SOGPU_OK;
}

/* kdBuildTree replacement.  Fully asynchronous on the handle's stream (no host round trip).
 * focus_nh > 0: only the region the focus_nh halos in d_centers/d_rgtp can reach within
 * focus_balls steps of the ball schedule is kept (sogpu_build_grid_for). */
static int build_grid_impl(sogpu *h, int32_t focus_nh, int focus_balls)
{
    if (!h || !h->d_in || h->n <= 0) return set_err(SOGPU_ERR_ARG, "sogpu_build_grid: no particles set");
    CU(cudaSetDevice(h->device));
    int lb;
    const int nc = pick_cells(h->indexed ? h->n_total : h->n, h->ppc, &lb);   /* one resolution for every rank of a domain run */
    const int64_t ncell = (int64_t)nc * nc * nc;
    const int keybits = 3 * lb;
    /* device-side particle count (domain steps): h->n is the capacity of the input buffer, n_work the
     * expected count that sizes the bucket table and the launches; every kernel reads the true count */
    const uint32_t *n_dev0 = h->d_n_dev;
    const int64_t n_work = n_dev0 ? std::max<int64_t>(std::min(h->n_hint, h->n), 1) : h->n;
    /* final buckets: ~1024 particles on average and at most BKT_CELLS cells each */
    int cbt = 0;
    while (((int64_t)BKT_AVG << cbt) < n_work) ++cbt;
    if (h->two_level == 0) cbt = 0;
    if (cbt < keybits - 12) cbt = keybits - 12;
    if (cbt > keybits) cbt = keybits;
    const int cell_bits = keybits - cbt;
    int lvl_bits = (focus_nh != 0) ? 9 : 8;            /* partition levels, digits of <= 8 bits; sparse (focused) inputs: 9,
                                                        * one pass fewer over few particles (measured at 1024^3: -2.5 %) */
    if (const char *e = getenv("SOGPU_LEVEL_BITS")) lvl_bits = std::min(9, std::max(4, atoi(e)));
    const int L = (cbt + lvl_bits - 1) / lvl_bits;
    if (L > 4) return set_err(SOGPU_ERR_UNSUPPORTED, "too many partition levels");
    if (!h->d_ce || h->nc != nc) {
        cudaFree(h->d_ce); h->d_ce = nullptr;
        CU(cudaMalloc(&h->d_ce, (size_t)(ncell + 1) * sizeof(uint32_t)));
    }
    if (!h->d_bsum) CU(cudaMalloc(&h->d_bsum, 4096 * sizeof(uint32_t)));
    if (h->n > h->grid_n_cap) {
        cudaFree(h->d_sorted);
        h->d_sorted = nullptr; h->grid_n_cap = 0;
        CU(cudaMalloc(&h->d_sorted, (size_t)h->n * sizeof(float4)));
        h->grid_n_cap = h->n;
    }
    if (L > 0 && h->n > h->tmp_cap) {
        cudaFree(h->d_tmp4); cudaFree(h->d_key[0]); cudaFree(h->d_key[1]);
        h->d_tmp4 = nullptr; h->d_key[0] = h->d_key[1] = nullptr; h->tmp_cap = 0;
        CU(cudaMalloc(&h->d_tmp4, (size_t)h->n * sizeof(float4)));
        CU(cudaMalloc(&h->d_key[0], (size_t)h->n * sizeof(uint32_t)));
        CU(cudaMalloc(&h->d_key[1], (size_t)h->n * sizeof(uint32_t)));
        h->tmp_cap = h->n;
    }
    if (!h->d_massmm) CU(cudaMalloc(&h->d_massmm, 2 * sizeof(uint32_t)));
    if (!h->d_mt) CU(cudaMalloc(&h->d_mt, sizeof(so_mass_table)));
    h->nc = nc; h->lb = lb; h->ncell = ncell;

    GridDev &g = h->g;
    g.sorted = h->d_sorted; g.ce = h->d_ce;
    g.nc = nc; g.lb = lb; g.tb = std::min(lb, 3);
    double hmax = 0.0, lmin = 1e300;
    for (int k = 0; k < 3; ++k) {
        double Lk = (double)h->period[k];
        g.L[k] = h->period[k];
        g.halfL[k] = 0.5f * h->period[k];
        g.g0[k] = (float)((double)h->center[k] - 0.5 * Lk);
        g.dg0[k] = (double)g.g0[k];
        g.dh[k] = Lk / nc;
        g.invh[k] = (float)((double)nc / Lk);
        g.dinvh[k] = (double)g.invh[k];
        hmax = std::max(hmax, g.dh[k]);
        lmin = std::min(lmin, Lk);
    }
    g.bmax_pruned = 0.5 * lmin - 2.0 * hmax;
    g.mask = nullptr; g.mb = 0; g.ms = 0; g.nofilter = 0;
    g.indexed = h->indexed ? 1 : 0;
    g.use_tma = h->use_tma ? 1 : 0;

    cudaStream_t s = h->stream;
    const double N = (double)n_work;
    h->focused = false;
    if (focus_nh != 0) {   /* (L == 0, tiny inputs: nothing is filtered, but the mask still guards the balls) */
        const int mb = std::min(lb, h->mask_bits);
        const size_t words = ((size_t)1 << (3 * mb)) / 32 + 1;
        if (!h->d_mask) CU(cudaMalloc(&h->d_mask, (((size_t)1 << 27) / 32 + 1) * sizeof(uint32_t)));
        g.mb = mb; g.ms = lb - mb;
        g.mask_rmin = h->mask_rmin_cells * hmax * (double)(1 << g.ms);
        if (focus_nh > 0) {     /* (focus_nh < 0: h->d_mask already holds this rank's mask, see domain_step.cuh) */
            CU(cudaMemsetAsync(h->d_mask, 0, words * sizeof(uint32_t), s));
            GridDev gm = g;
            gm.mask = h->d_mask;
            ProfScope p(h, KID_MARK_MASK);
            k_mark_mask<<<std::min((focus_nh + 7) / 8, h->sm_count * 8), 256, 0, s>>>(gm, h->d_centers, h->d_rgtp,
                                                                                      focus_nh, focus_balls, h->d_mask);
        }
        g.mask = (focus_nh < 0 && h->mask_ready) ? h->mask_ready : h->d_mask;
        g.nofilter = (focus_nh < 0) ? 1 : 0;       /* domain steps: everything that was routed here lies inside the mask */
        h->focused = true;
        h->focus_balls = focus_balls;
    }
    h->stats.last_kernel_launches = 0;
    if (h->indexed) {                      /* the (single) particle mass came with sogpu_set_particles_device_indexed */
        uint32_t mb32;
        memcpy(&mb32, &h->indexed_mass, sizeof(mb32));
        k_store_u32<<<1, 32, 0, s>>>(h->d_massmm, mb32, h->d_massmm + 1, mb32);
    } else {
        CU(cudaMemsetAsync(h->d_massmm, 0xFF, sizeof(uint32_t), s));
        CU(cudaMemsetAsync(h->d_massmm + 1, 0, sizeof(uint32_t), s));
    }

    /* digit widths: cbt split as evenly as possible over L levels, most significant first */
    int db[4] = {0, 0, 0, 0}, bits_done = 0;
    for (int l = 0; l < L; ++l) db[l] = cbt / L + (l < cbt % L ? 1 : 0);
    for (int l = 0, done = 0; l < std::max(L, 1); ++l) {
        done += db[l];
        size_t need = ((size_t)1 << done) + 1;
        if (need > h->lvl_cap[l]) {
            cudaFree(h->d_lvl_start[l]); cudaFree(h->d_lvl_cursor[l]);
            h->d_lvl_start[l] = h->d_lvl_cursor[l] = nullptr; h->lvl_cap[l] = 0;
            CU(cudaMalloc(&h->d_lvl_start[l], need * sizeof(uint32_t)));
            CU(cudaMalloc(&h->d_lvl_cursor[l], need * sizeof(uint32_t)));
            h->lvl_cap[l] = need;
        }
    }
    /* function attributes belong to the device's context: once per handle, not once per process */
    bool &attr_done = h->build_attr_done;
    const size_t part_smem = (size_t)LVL_T * (16 + 4 + 2 + 2 + 2);
    const size_t bkt_smem = (size_t)BKT_CAP * (16 + 2 + 2 + 2) + (size_t)BKT_CELLS * 4;
    if (!attr_done) {
        CU(cudaFuncSetAttribute(k_lvl_partition<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
        CU(cudaFuncSetAttribute(k_lvl_partition<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
        CU(cudaFuncSetAttribute(k_lvl_partition<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
        CU(cudaFuncSetAttribute(k_lvl_partition<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)part_smem));
        CU(cudaFuncSetAttribute(k_bucket_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bkt_smem));
        CU(cudaFuncSetAttribute(k_lvl_partition_rt<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM));
        CU(cudaFuncSetAttribute(k_lvl_partition_rt<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM));
        CU(cudaFuncSetAttribute(k_lvl_partition_rt<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM));
        CU(cudaFuncSetAttribute(k_lvl_partition_rt<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM));
        attr_done = true;
    }
    const int64_t tiles = (n_work + LVL_T - 1) / LVL_T;
    const int hist_grid = (int)std::min<int64_t>(tiles, (int64_t)h->sm_count * 8);
    const int part_grid = (int)std::min<int64_t>(tiles, (int64_t)h->sm_count * 2);

    const float4 *src = h->d_in;
    const uint32_t *src_key = nullptr;
    if (L == 0) {
        /* tiny input: one bucket = the whole array; the histogram launch only finds mass min/max */
        LevelDesc lv; lv.shift = 0; lv.db = 0; lv.pshift = 32; lv.n_parents = 1;
        CU(cudaMemsetAsync(h->d_lvl_cursor[0], 0, 2 * sizeof(uint32_t), s));
        { ProfScope p(h, KID_LVL_HIST, 16.0 * N);
          k_lvl_hist<true><<<hist_grid, 256, 0, s>>>(src, nullptr, h->n, g, lv, nullptr, h->d_lvl_cursor[0], h->d_massmm, n_dev0); }
        { int rc = launch_mass_table(h); if (rc) return rc; }
        k_store_u32<<<1, 32, 0, s>>>(h->d_lvl_start[0], 0u, h->d_lvl_start[0] + 1, (uint32_t)h->n);
        if (n_dev0) k_copy_u32<<<1, 32, 0, s>>>(h->d_lvl_start[0] + 1, n_dev0);
    }
    for (int l = 0; l < L; ++l) {
        LevelDesc lv;
        lv.db = db[l];
        lv.shift = keybits - bits_done - db[l];
        lv.pshift = bits_done ? keybits - bits_done : 32;
        lv.n_parents = 1u << bits_done;
        const size_t M = (size_t)1 << (bits_done + db[l]);
        const bool last = (l == L - 1);
        float4 *dst = ((L - 1 - l) % 2 == 0) ? h->d_tmp4 : h->d_sorted;
        uint32_t *dst_key = h->d_key[l & 1];
        const uint32_t *pstart = l ? h->d_lvl_start[l - 1] : nullptr;
        /* particles that survive level 0 (== N unless the build is focused): sentinel of its scan */
        const uint32_t *n_dev = (l && h->focused) ? h->d_lvl_start[0] + ((size_t)1 << db[0]) : n_dev0;
        CU(cudaMemsetAsync(h->d_lvl_cursor[l], 0, M * sizeof(uint32_t), s));
        {
            ProfScope p(h, KID_LVL_HIST, (l ? 4.0 : 16.0) * N);
            if (l == 0) k_lvl_hist<true><<<hist_grid, 256, 0, s>>>(src, nullptr, h->n, g, lv, nullptr, h->d_lvl_cursor[l], h->d_massmm, n_dev);
            else k_lvl_hist<false><<<hist_grid, 256, 0, s>>>(src, src_key, h->n, g, lv, pstart, h->d_lvl_cursor[l], h->d_massmm, n_dev);
        }
        if (l == 0) { int rc = launch_mass_table(h); if (rc) return rc; }
        {
            ProfScope p(h, KID_LVL_SCAN, 12.0 * (double)M, M <= h->scan1_max ? 1 : 3);
            if (M <= h->scan1_max) {
                k_scan_one<<<1, 1024, 0, s>>>(h->d_lvl_cursor[l], (int64_t)M, h->d_lvl_start[l], h->d_lvl_cursor[l]);
            } else {
                int64_t nt = ((int64_t)M + SCAN_TILE - 1) / SCAN_TILE;
                k_scan_reduce<<<(unsigned)nt, 256, 0, s>>>(h->d_lvl_cursor[l], (int64_t)M, h->d_bsum);
                k_scan_bsums<<<1, 1024, 0, s>>>(h->d_bsum, nt);
                k_scan_apply<<<(unsigned)nt, 256, 0, s>>>(h->d_lvl_cursor[l], (int64_t)M, h->d_bsum, h->d_lvl_start[l],
                                                         h->d_lvl_cursor[l]);
            }
        }
        {
            ProfScope p(h, KID_LVL_PARTITION, ((l ? 20.0 : 16.0) + 16.0 + (last ? 0.0 : 4.0)) * N);
            if (h->two_level == 2) {       /* staged variant (tile sorted in shared memory first) */
                if (l == 0 && !last) k_lvl_partition<true, true><<<part_grid, LVL_THREADS, part_smem, s>>>(src, nullptr, h->n, g, lv, nullptr, h->d_lvl_cursor[l], dst, dst_key, n_dev);
                else if (l == 0) k_lvl_partition<true, false><<<part_grid, LVL_THREADS, part_smem, s>>>(src, nullptr, h->n, g, lv, nullptr, h->d_lvl_cursor[l], dst, dst_key, n_dev);
                else if (!last) k_lvl_partition<false, true><<<part_grid, LVL_THREADS, part_smem, s>>>(src, src_key, h->n, g, lv, pstart, h->d_lvl_cursor[l], dst, dst_key, n_dev);
                else k_lvl_partition<false, false><<<part_grid, LVL_THREADS, part_smem, s>>>(src, src_key, h->n, g, lv, pstart, h->d_lvl_cursor[l], dst, dst_key, n_dev);
            } else {
                const int sg = (int)std::min<int64_t>((n_work + RT_T - 1) / RT_T, (int64_t)h->sm_count * 3);
                if (l == 0 && !last) k_lvl_partition_rt<true, true><<<sg, RT_NT, RT_SMEM, s>>>(src, nullptr, h->n, g, lv, nullptr, h->d_lvl_cursor[l], dst, dst_key, n_dev);
                else if (l == 0) k_lvl_partition_rt<true, false><<<sg, RT_NT, RT_SMEM, s>>>(src, nullptr, h->n, g, lv, nullptr, h->d_lvl_cursor[l], dst, dst_key, n_dev);
                else if (!last) k_lvl_partition_rt<false, true><<<sg, RT_NT, RT_SMEM, s>>>(src, src_key, h->n, g, lv, pstart, h->d_lvl_cursor[l], dst, dst_key, n_dev);
                else k_lvl_partition_rt<false, false><<<sg, RT_NT, RT_SMEM, s>>>(src, src_key, h->n, g, lv, pstart, h->d_lvl_cursor[l], dst, dst_key, n_dev);
            }
        }
        src = dst;
        src_key = dst_key;
        bits_done += db[l];
    }
    {
        const uint32_t nb = 1u << cbt;
        const uint32_t *bstart = h->d_lvl_start[L ? L - 1 : 0];
        int grid = (int)std::min<int64_t>((int64_t)nb, (int64_t)h->sm_count * 16);
        const double sort_bytes = 32.0 * N + 4.0 * (double)ncell;
        if (h->two_level == 2) {
            ProfScope p(h, KID_BUCKET_SORT, sort_bytes);
            k_bucket_sort<<<grid, BKT_THREADS, bkt_smem, s>>>(src, g, cell_bits, nb, bstart, h->d_sorted, h->d_ce, (L == 0 && !g.indexed) ? 1 : 0);
        } else {
            const uint32_t *live = nullptr, *live_n = nullptr;
            const bool sparse_ok = g.mask && cell_bits >= lb && (nc >> g.ms) >= 32 &&
                                   ((ncell >> cbt) >> lb) * ((nc >> g.ms) >> 5) <= 128 && !(L == 0 && !g.indexed) &&
                                   !getenv("SOGPU_DENSE_BUCKETS");
            bool split_big = false;
            if (g.mask && nb >= 4096u) {
                /* focused build with many final buckets: settle the empty ones outside the mask first */
                /* three lists of nb + 1 words (the last one is the length): live buckets (those for the
                 * warp-per-bucket kernel when it is used), buckets with many particles, buckets the warp kernel left */
                if ((size_t)nb + 1 > h->live_cap) {
                    cudaFree(h->d_live); h->d_live = nullptr; h->live_cap = 0;
                    CU(cudaMalloc(&h->d_live, 3 * ((size_t)nb + 1) * sizeof(uint32_t)));
                    h->live_cap = (size_t)nb + 1;
                }
                split_big = sparse_ok && h->bucket_warp;
                uint32_t *big = h->d_live + (size_t)nb + 1;
                CU(cudaMemsetAsync(h->d_live + nb, 0, sizeof(uint32_t), s));
                if (split_big) CU(cudaMemsetAsync(big + nb, 0, sizeof(uint32_t), s));
                ProfScope pl(h, KID_BUCKET_LIVE);
                k_bucket_live<<<(nb + 255) / 256, 256, 0, s>>>(g, cell_bits, nb, bstart, h->d_ce, h->d_live, h->d_live + nb,
                                                               32u * BW_IT, split_big ? big : nullptr, big + nb);
                live = h->d_live; live_n = h->d_live + nb;
            }
            if (sparse_ok) {
                /* focused grid: per-bucket work proportional to its marked cells (k_bucket_sort_sparse) */
                grid = (int)std::min<int64_t>((int64_t)nb, (int64_t)h->sm_count * 8);
                if (split_big) {
                    /* buckets with many particles one per CTA on a side stream, the others one per warp beside them;
                     * what the warp kernel cannot hold (too many marked cells for its counters: rare) afterwards */
                    uint32_t *big = h->d_live + (size_t)nb + 1, *rest = big + (size_t)nb + 1;
                    CU(cudaMemsetAsync(rest + nb, 0, sizeof(uint32_t), s));
                    CU(cudaEventRecord(h->ev_fork, s));
                    CU(cudaStreamWaitEvent(h->aux[1], h->ev_fork, 0));
                    {
                        h->launch_stream = h->aux[1];
                        ProfScope pb(h, KID_BUCKET_BIG);
                        k_bucket_sort_sparse<<<grid, BS_NT, 0, h->aux[1]>>>(src, g, cell_bits, nb, bstart, h->d_sorted, h->d_ce, big, big + nb);
                    }
                    h->launch_stream = s;
                    CU(cudaEventRecord(h->ev_join[1], h->aux[1]));
                    {
                        ProfScope p(h, KID_BUCKET_SORT, sort_bytes);
                        k_bucket_sort_sparse_warp<<<h->sm_count * 4, BW_NT, 0, s>>>(src, g, cell_bits, bstart, h->d_sorted, h->d_ce,
                                                                                   live, live_n, rest, rest + nb);
                    }
                    {
                        ProfScope pb(h, KID_BUCKET_BIG);
                        k_bucket_sort_sparse<<<h->sm_count, BS_NT, 0, s>>>(src, g, cell_bits, nb, bstart, h->d_sorted, h->d_ce, rest, rest + nb);
                    }
                    CU(cudaStreamWaitEvent(s, h->ev_join[1], 0));
                } else {
                    ProfScope p(h, KID_BUCKET_SORT, sort_bytes);
                    k_bucket_sort_sparse<<<grid, BS_NT, 0, s>>>(src, g, cell_bits, nb, bstart, h->d_sorted, h->d_ce, live, live_n);
                }
            } else if (n_work / (int64_t)nb < 160 && nb >= 4096u) {      /* sparse buckets: small CTAs, many in flight */
                grid = (int)std::min<int64_t>((int64_t)nb, (int64_t)h->sm_count * 48);
                ProfScope p(h, KID_BUCKET_SORT, sort_bytes);
                k_bucket_sort_rt<64, 12><<<grid, 64, 0, s>>>(src, g, cell_bits, nb, bstart, h->d_sorted, h->d_ce,
                                                             (L == 0 && !g.indexed) ? 1 : 0, live, live_n);
            } else {
                ProfScope p(h, KID_BUCKET_SORT, sort_bytes);
                k_bucket_sort_rt<256, 4><<<grid, 256, 0, s>>>(src, g, cell_bits, nb, bstart, h->d_sorted, h->d_ce,
                                                              (L == 0 && !g.indexed) ? 1 : 0, live, live_n);
            }
        }
    }
    if (h->focused) k_copy_u32<<<1, 32, 0, s>>>(h->d_ce + ncell, h->d_lvl_start[0] + ((size_t)1 << db[0]));
    else if (n_dev0) k_copy_u32<<<1, 32, 0, s>>>(h->d_ce + ncell, n_dev0);
    else k_store_u32<<<1, 32, 0, s>>>(h->d_ce + ncell, (uint32_t)h->n, nullptr, 0u);
    CU(cudaStreamWaitEvent(s, h->ev_join[0], 0));         /* the mass table (side stream) */
    CU(cudaGetLastError());
    h->built = true;
    h->have_result = false;
    h->mass_state = -1;
    h->stats.n_particles = h->n;
    h->stats.cells_per_axis = nc;
    return SOGPU_OK;
}

extern "C" int sogpu_build_grid(sogpu_t *h) { return build_grid_impl(h, 0, 0); }

static int ensure_query(sogpu *h, int32_t nh);

extern "C" int sogpu_build_grid_for(sogpu_t *h, const float *centers, const float *rgtp, int32_t nh, int32_t n_balls)
{
    if (!h || !centers || !rgtp || nh <= 0 || n_balls < 1)
        return set_err(SOGPU_ERR_ARG, "sogpu_build_grid_for: bad argument");
    if (!h->d_in || h->n <= 0) return set_err(SOGPU_ERR_ARG, "sogpu_build_grid_for: no particles set");
    CU(cudaSetDevice(h->device));
    int rc = ensure_query(h, nh);
    if (rc) return rc;
    rc = ensure_pinned(h, (size_t)nh * 4 * sizeof(float));
    if (rc) return rc;
    float *pc = (float *)h->h_pin, *pr = pc + (size_t)3 * nh;
    memcpy(pc, centers, (size_t)nh * 3 * sizeof(float));
    memcpy(pr, rgtp, (size_t)nh * sizeof(float));
    CU(cudaMemcpyAsync(h->d_centers, pc, (size_t)nh * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_rgtp, pr, (size_t)nh * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    return build_grid_impl(h, nh, n_balls);
}

/* same, with the halo list already on the device (asynchronous) */
extern "C" int sogpu_build_grid_for_device(sogpu_t *h, const void *d_centers, const void *d_rgtp, int32_t nh,
                                           int32_t n_balls)
{
    if (!h || !d_centers || !d_rgtp || nh <= 0 || n_balls < 1)
        return set_err(SOGPU_ERR_ARG, "sogpu_build_grid_for_device: bad argument");
    CU(cudaSetDevice(h->device));
    int rc = ensure_query(h, nh);
    if (rc) return rc;
    if (d_centers != h->d_centers)
        CU(cudaMemcpyAsync(h->d_centers, d_centers, (size_t)nh * 3 * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    if (d_rgtp != h->d_rgtp)
        CU(cudaMemcpyAsync(h->d_rgtp, d_rgtp, (size_t)nh * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    return build_grid_impl(h, nh, n_balls);
}

/* lazily learn (one small D2H) whether the particle masses were all equal */
static int fetch_mass_state(sogpu *h)
{
    if (h->mass_state >= 0) return SOGPU_OK;
    uint32_t mm[2];
    CU(cudaMemcpyAsync(mm, h->d_massmm, sizeof(mm), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->mass_state = (mm[0] == mm[1]) ? 1 : 0;
    memcpy(&h->mass, &mm[0], sizeof(float));
    h->stats.equal_mass = h->mass_state;
    return SOGPU_OK;
}

static int ensure_query(sogpu *h, int32_t nh)
{
    if (!h->d_counters) {
        CU(cudaMalloc(&h->d_counters, 24 * sizeof(uint32_t)));
        CU(cudaMalloc(&h->d_u64, 4 * sizeof(unsigned long long)));
    }
    if (nh > h->cap_h) {
        free_query(h);
        int32_t cap = std::max(nh, 1024);
        CU(cudaMalloc(&h->d_centers, (size_t)cap * 3 * sizeof(float)));
        CU(cudaMalloc(&h->d_rgtp, (size_t)cap * sizeof(float)));
        CU(cudaMalloc(&h->d_small, (size_t)cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_big, (size_t)cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_esmall, (size_t)cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_ebig, (size_t)cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_huge, (size_t)cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_ehuge, (size_t)cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_defer, (size_t)cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_out_n, (size_t)cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_out_m, (size_t)cap * sizeof(float)));
        CU(cudaMalloc(&h->d_out_key, (size_t)cap * sizeof(unsigned long long)));
        CU(cudaMalloc(&h->d_out_off, ((size_t)cap + 1) * sizeof(unsigned long long)));
        CU(cudaMalloc(&h->d_csr_off, ((size_t)cap + 1) * sizeof(unsigned long long)));
        h->cap_h = cap;
    }
    if (!h->d_members || h->member_cap < (unsigned long long)h->n) {
        cudaFree(h->d_members); cudaFree(h->d_md2); cudaFree(h->d_members2); cudaFree(h->d_md2_2);
        h->d_members = nullptr; h->d_md2 = nullptr; h->d_members2 = nullptr; h->d_md2_2 = nullptr;
        h->member_cap = (unsigned long long)std::max<int64_t>(h->n, 1 << 20);
        CU(cudaMalloc(&h->d_members, (size_t)h->member_cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_md2, (size_t)h->member_cap * sizeof(float)));
    }
    return SOGPU_OK;
}

template <int NT, typename K, typename... Extra>
static void launch_persistent(sogpu *h, K kernel, const QueryArgs &a, int nh, Extra... extra)
{
    int ctas = h->sm_count * (NT == 32 ? h->qgrid32 : NT == 256 ? h->qgrid256 : 1);
    int need = (nh + Cfg<NT>::GROUPS - 1) / Cfg<NT>::GROUPS;
    if (need < ctas) ctas = std::max(need, 1);
    kernel<<<ctas, NT * Cfg<NT>::GROUPS, query_smem_bytes<NT>(h->use_tma), h->launch_stream>>>(a, extra...);
}

/* enqueue query + member emission for nh halos whose centers/rgtp are on the device */
static int run_query(sogpu *h, const float *d_centers, const float *d_rgtp, int32_t nh, float thr, int32_t nM)
{
    if (!h->built) return set_err(SOGPU_ERR_ARG, "sogpu_so: call sogpu_build_grid first");
    if (nM < 2) return set_err(SOGPU_ERR_ARG, "nMembers must be >= 2 (the reference reads nnList[-1] otherwise, kd2.c:791)");
    if (nh <= 0) return set_err(SOGPU_ERR_ARG, "no halos");
    if (!(thr > 0.0f)) return set_err(SOGPU_ERR_ARG, "density threshold must be positive");
    int rc = ensure_query(h, nh);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    h->stats.last_kernel_launches = 0;
    /* the library keeps its own copy of the catalog: a member buffer that turns out too small is grown and the
     * lists are emitted again from it, whenever the caller asks for them */
    if (d_centers != h->d_centers)
        CU(cudaMemcpyAsync(h->d_centers, d_centers, (size_t)nh * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (d_rgtp != h->d_rgtp)
        CU(cudaMemcpyAsync(h->d_rgtp, d_rgtp, (size_t)nh * sizeof(float), cudaMemcpyDeviceToDevice, s));
    d_centers = h->d_centers; d_rgtp = h->d_rgtp;
    h->last_thr = thr; h->last_nM = nM;
    CU(cudaMemsetAsync(h->d_counters, 0, 24 * sizeof(uint32_t), s));
    CU(cudaMemsetAsync(h->d_u64, 0, 4 * sizeof(unsigned long long), s));

    /* counters: 0 small_n 1 big_n 2 work_small 3 work_big 4 flags 5 esmall_n 6 ebig_n 7 work_es 8 work_eb
     *           13 huge_n 14 work_huge 15 ehuge_n 16 work_ehuge   (9-12: general path) */
    /* few halos per GPU (domain steps over several ranks): the step ends with the tail of the largest halos, so more
     * of them get a 1024-thread CTA of their own beside the fused kernel (measured at 1024^3: 8 GPUs -6 %, 1 GPU +1 %) */
    const int per_rank = h->q_owner ? nh / std::max(1, h->q_ranks) : nh;
    const float small_max = h->cls_small_max,
                huge_min = (!getenv("SOGPU_HUGE_MIN") && per_rank < 25000) ? std::min(h->cls_huge_min, 8192.0f) : h->cls_huge_min;
    {
        ProfScope p(h, KID_CLASSIFY);
        k_classify<<<(nh + 255) / 256, 256, 0, s>>>(d_rgtp, nh, thr, h->d_mt, small_max, huge_min, h->d_small,
                                                    h->d_counters + 0, h->d_big, h->d_counters + 1, h->d_huge,
                                                    h->d_counters + 13, h->q_owner, h->q_me);
    }
    QueryArgs a;
    a.g = h->g;
    a.centers = d_centers; a.rgtp = d_rgtp;
    a.thr = thr; a.nM = nM; a.first_ball = h->first_ball;
    a.out_n = h->d_out_n; a.out_m = h->d_out_m; a.out_key = h->d_out_key; a.out_off = h->d_out_off;
    a.members = h->d_members; a.md2 = h->want_d2 ? h->d_md2 : nullptr;
    a.member_cap = h->member_cap;
    a.member_cursor = h->d_u64 + 0; a.emit_in_query = 1;
    a.evals = h->d_u64 + 1;
    a.flags = h->d_counters + 4;
    a.mt = h->d_mt;
    a.defer_list = h->d_defer; a.defer_n = h->d_counters + 17;
    a.timeline = nullptr; a.tl_slot = 0;
    if (getenv("SOGPU_DEBUG_TIMELINE")) {       /* debug: when did each class really start / end on the device */
        if (!h->d_timeline) CU(cudaMalloc(&h->d_timeline, 16 * sizeof(unsigned long long)));
        unsigned long long init[16];
        for (int k = 0; k < 8; ++k) { init[2 * k] = ~0ull; init[2 * k + 1] = 0ull; }
        CU(cudaMemcpyAsync(h->d_timeline, init, sizeof(init), cudaMemcpyHostToDevice, s));
        a.timeline = h->d_timeline;
    }

    /* The three size classes are independent and run concurrently.  ORDER MATTERS: a 1024-thread CTA needs a
     * whole SM, so the cluster-size class goes first, on the main stream (no event wait in front of it):
     * measured with %globaltimer, launched behind the warp kernel it started 121 us late because 592
     * warp-kernel CTAs had filled every SM.  Its CTAs without work leave within microseconds; the 256-thread
     * class (high-priority side stream) and the warp class (side stream) then fill the other SMs. */
    CU(cudaEventRecord(h->ev_fork, s));
    CU(cudaStreamWaitEvent(h->aux[0], h->ev_fork, 0));
    CU(cudaStreamWaitEvent(h->aux[1], h->ev_fork, 0));
    h->launch_stream = s;
    a.list = h->d_huge; a.list_n = h->d_counters + 13; a.work_counter = h->d_counters + 14; a.tl_slot = 0;
    { ProfScope p(h, KID_QUERY_HUGE); launch_persistent<1024>(h, k_so_query<1024>, a, std::min(nh, h->sm_count)); }
    if (h->fused_query) {
        /* the two common classes in one persistent kernel (see k_so_query_fused), then the few halos a warp deferred */
        h->launch_stream = h->aux[0];
        a.list = h->d_big; a.list_n = h->d_counters + 1; a.work_counter = h->d_counters + 3;
        a.list2 = h->d_small; a.list2_n = h->d_counters + 0; a.work_counter2 = h->d_counters + 2;
        {
            ProfScope p(h, KID_QUERY_FUSED);
            const int ctas = std::max(1, std::min(h->sm_count * QF_MINB, (nh + 7) / 8 + h->sm_count));
            k_so_query_fused<<<ctas, 256, fused_smem_bytes(), h->aux[0]>>>(a);
        }
        a.list = h->d_defer; a.list_n = h->d_counters + 17; a.work_counter = h->d_counters + 18; a.tl_slot = 3;
        { ProfScope p(h, KID_QUERY_BLOCK); launch_persistent<256>(h, k_so_query<256>, a, 64); }
        CU(cudaEventRecord(h->ev_join[0], h->aux[0]));
        CU(cudaEventRecord(h->ev_join[1], h->aux[1]));
    } else
    for (int pass = 0; pass < 2; ++pass) {
        const bool big_now = (pass == 0) == (h->qorder == 0);
        if (big_now) {      /* mid-size halos: one 256-thread CTA each */
            h->launch_stream = h->aux[0];
            a.list = h->d_big; a.list_n = h->d_counters + 1; a.work_counter = h->d_counters + 3; a.tl_slot = 1;
            { ProfScope p(h, KID_QUERY_BLOCK); launch_persistent<256>(h, k_so_query<256>, a, nh); }
            CU(cudaEventRecord(h->ev_join[0], h->aux[0]));
        } else {            /* small halos: one warp each; the few it cannot finish go to a CTA kernel right behind it */
            h->launch_stream = h->aux[1];
            a.list = h->d_small; a.list_n = h->d_counters + 0; a.work_counter = h->d_counters + 2; a.tl_slot = 2;
            { ProfScope p(h, KID_QUERY_WARP); launch_persistent<32>(h, k_so_query<32>, a, nh); }
            a.list = h->d_defer; a.list_n = h->d_counters + 17; a.work_counter = h->d_counters + 18; a.tl_slot = 3;
            { ProfScope p(h, KID_QUERY_BLOCK); launch_persistent<256>(h, k_so_query<256>, a, 64); }
            CU(cudaEventRecord(h->ev_join[1], h->aux[1]));
        }
    }
    h->launch_stream = s;
    CU(cudaStreamWaitEvent(s, h->ev_join[0], 0));
    CU(cudaStreamWaitEvent(s, h->ev_join[1], 0));
    /* the member lists were written by the query kernels themselves, one slot per halo in the order the halos
     * finished (d_out_off[h] = start, N_Delta entries); ensure_csr() puts them in catalog order on demand */
    h->members_csr = false;
    CU(cudaGetLastError());
    h->last_h = nh;
    h->have_result = true;
    h->members_sorted = false;
    return SOGPU_OK;
}

/* ---- segmented sort of the keys in d_gkeys[0] (seg_sort.cuh); *result = the buffer that holds them sorted ---- */
static int seg_sort_scratch(sogpu *h, int nseg)
{
    if (nseg + 1 > h->ss_cap) {
        cudaFree(h->d_ss_tiles); cudaFree(h->d_ss_off);
        h->d_ss_tiles = nullptr; h->d_ss_off = nullptr; h->ss_cap = 0;
        const int cap = std::max(nseg + 1, 4096);
        CU(cudaMalloc(&h->d_ss_tiles, (size_t)(cap + 1) * sizeof(uint32_t)));
        CU(cudaMalloc(&h->d_ss_off, (size_t)(cap + 1) * sizeof(unsigned long long)));
        h->ss_cap = cap;
    }
    if (!h->d_ss_max) CU(cudaMalloc(&h->d_ss_max, sizeof(unsigned long long)));
    return SOGPU_OK;
}

static int seg_sort_keys(sogpu *h, int nseg, const unsigned long long *d_off, size_t tot, unsigned long long **result)
{
    cudaStream_t s = h->stream;
    *result = h->d_gkeys[0];
    if (tot == 0 || nseg <= 0) return SOGPU_OK;
    int rc = seg_sort_scratch(h, nseg);
    if (rc) return rc;
    k_segsort_plan<<<1, 1024, 0, s>>>(d_off, nseg, h->d_ss_tiles, h->d_ss_max);
    unsigned long long max_n = 0;
    CU(cudaMemcpyAsync(&max_n, h->d_ss_max, sizeof(max_n), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const size_t tiles_ub = tot / SS_T + (size_t)nseg;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>(tiles_ub, (size_t)h->sm_count * 8));
    int cur = 1;
    {
        ProfScope p(h, KID_SEGSORT);
        k_segsort_tiles<<<grid, SS_NT, 0, s>>>(d_off, nseg, h->d_ss_tiles, h->d_gkeys[0], h->d_gkeys[1]);
        for (unsigned long long run = SS_T; run < max_n; run *= 2ull) {
            k_segsort_merge<<<grid, SS_NT, 0, s>>>(d_off, nseg, h->d_ss_tiles, run, h->d_gkeys[cur], h->d_gkeys[cur ^ 1]);
            cur ^= 1;
        }
    }
    CU(cudaGetLastError());
    *result = h->d_gkeys[cur];
    return SOGPU_OK;
}

/* ---- general path driver: rounds over the ball schedule, batches bounded by the scratch size ---- */
static int gen_scratch(sogpu *h, size_t entries)
{
    if (entries <= h->gen_cap) return SOGPU_OK;
    for (int k = 0; k < 2; ++k) {
        cudaFree(h->d_gkeys[k]); cudaFree(h->d_gmass[k]);
        h->d_gkeys[k] = nullptr; h->d_gmass[k] = nullptr;
    }
    h->gen_cap = 0;
    for (int k = 0; k < 2; ++k) {
        CU(cudaMalloc(&h->d_gkeys[k], entries * sizeof(unsigned long long)));
        CU(cudaMalloc(&h->d_gmass[k], entries * sizeof(float)));
    }
    h->gen_cap = entries;
    return SOGPU_OK;
}

static int emit_members(sogpu *h, const float *d_centers, const float *d_rgtp, int32_t nh, float thr, int32_t nM);

/* subset == nullptr: all nh halos; else the n_subset halo ids in the device array `subset` */
static int run_query_general(sogpu *h, const float *d_centers, const float *d_rgtp, int32_t nh, float thr, int32_t nM,
                             const int32_t *subset = nullptr, uint32_t n_subset = 0)
{
    cudaStream_t s = h->stream;
    if (nh > h->gen_cap_h) {
        cudaFree(h->d_gen_state); cudaFree(h->d_gen_list[0]); cudaFree(h->d_gen_list[1]); cudaFree(h->d_gen_round);
        cudaFree(h->d_seg_n); cudaFree(h->d_seg_begin); cudaFree(h->d_seg_end);
        h->gen_cap_h = 0;
        CU(cudaMalloc(&h->d_gen_state, (size_t)nh * sizeof(GenState)));
        CU(cudaMalloc(&h->d_gen_list[0], (size_t)nh * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_gen_list[1], (size_t)nh * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_gen_round, (size_t)nh * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_seg_n, (size_t)nh * sizeof(unsigned long long)));
        CU(cudaMalloc(&h->d_seg_begin, (size_t)nh * sizeof(unsigned long long)));
        CU(cudaMalloc(&h->d_seg_end, (size_t)nh * sizeof(unsigned long long)));
        h->gen_cap_h = nh;
    }
    /* scratch budget per batch: at least one full-box ball must fit */
    const size_t budget = (size_t)std::max<int64_t>(h->n, (int64_t)1 << 22);
    int rc = gen_scratch(h, budget);
    if (rc) return rc;
    if (!subset) {
        CU(cudaMemsetAsync(h->d_counters, 0, 24 * sizeof(uint32_t), s));
        CU(cudaMemsetAsync(h->d_u64, 0, 4 * sizeof(unsigned long long), s));
    }
    /* counters: [9] list_n(cur) [10] work [11] round_n [12] next_n */
    const int n_init = subset ? (int)n_subset : nh;
    k_gen_init<<<(n_init + 255) / 256, 256, 0, s>>>(h->d_gen_state, d_rgtp, h->d_gen_list[0], h->d_counters + 9,
                                                    n_init, subset);
    GenArgs a;
    memset(&a, 0, sizeof(a));
    a.g = h->g; a.in = h->d_in; a.centers = d_centers; a.rgtp = d_rgtp; a.st = h->d_gen_state;
    a.round_list = h->d_gen_round; a.round_n = h->d_counters + 11;
    a.seg_n = h->d_seg_n; a.seg_begin = h->d_seg_begin;
    a.thr = thr; a.nM = nM;
    a.out_n = h->d_out_n; a.out_m = h->d_out_m; a.out_key = h->d_out_key;
    a.evals = h->d_u64 + 1;
    std::vector<unsigned long long> seg_n;
    int cur = 0;
    uint32_t n_active = (uint32_t)n_init;
    const size_t smem = query_smem_bytes<256>();
    for (int round = 0; n_active > 0 && round < 4096; ++round) {
        a.list = h->d_gen_list[cur]; a.list_n = h->d_counters + 9; a.work = h->d_counters + 10;
        a.next_list = h->d_gen_list[cur ^ 1]; a.next_n = h->d_counters + 12;
        CU(cudaMemsetAsync(h->d_counters + 10, 0, 3 * sizeof(uint32_t), s));      /* work, round_n, next_n */
        {
            ProfScope p(h, KID_QUERY_BLOCK);
            int ctas = (int)std::min<uint32_t>(n_active, (uint32_t)h->sm_count * 4);
            k_gen_count<<<ctas, 256, smem, s>>>(a);
        }
        uint32_t n_round = 0;
        CU(cudaMemcpyAsync(&n_round, h->d_counters + 11, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (n_round) {
            seg_n.resize(n_round);
            CU(cudaMemcpy(seg_n.data(), h->d_seg_n, (size_t)n_round * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            uint32_t s0 = 0;
            while (s0 < n_round) {
                size_t tot = 0;
                uint32_t s1 = s0;
                while (s1 < n_round && (s1 == s0 || tot + seg_n[s1] <= h->gen_cap)) tot += seg_n[s1++];
                rc = gen_scratch(h, tot);                     /* a single ball larger than the budget */
                if (rc) return rc;
                a.keys = h->d_gkeys[0]; a.mass = h->d_gmass[0];
                rc = seg_sort_scratch(h, (int)(s1 - s0));
                if (rc) return rc;
                k_gen_offsets<<<1, 1024, 0, s>>>(h->d_seg_n, s0, s1, h->d_seg_begin, h->d_seg_end, h->d_ss_off);
                {
                    ProfScope p(h, KID_EMIT_BLOCK);
                    int ctas = (int)std::min<uint32_t>(s1 - s0, (uint32_t)h->sm_count * 4);
                    k_gen_emit<<<ctas, 256, smem, s>>>(a, s0, s1);
                }
                unsigned long long *sorted_keys = nullptr;
                rc = seg_sort_keys(h, (int)(s1 - s0), h->d_ss_off, tot, &sorted_keys);   /* qsort(CmpList), kd2.c:781 */
                if (rc) return rc;
                k_gen_mass<<<(int)std::min<size_t>((tot + 255) / 256, (size_t)h->sm_count * 16), 256, 0, s>>>(sorted_keys, tot, h->d_in,
                                                                                                          h->d_gmass[1]);
                {
                    ProfScope p(h, KID_QUERY_WARP);
                    int warps = (int)(s1 - s0);
                    int ctas = std::min((warps + 7) / 8, h->sm_count * 8);
                    k_gen_scan<<<ctas, 256, 0, s>>>(a, s0, s1, sorted_keys, h->d_gmass[1]);
                }
                s0 = s1;
            }
        }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&n_active, h->d_counters + 12, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        CU(cudaMemcpyAsync(h->d_counters + 9, &n_active, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        CU(cudaStreamSynchronize(s));
        cur ^= 1;
    }
    return emit_members(h, d_centers, d_rgtp, nh, thr, nM);
}

/* member offsets in catalog order + the CSR member lists, from out_n / out_key of all nh halos */
static int emit_members(sogpu *h, const float *d_centers, const float *d_rgtp, int32_t nh, float thr, int32_t nM)
{
    cudaStream_t s = h->stream;
    QueryArgs q;
    memset(&q, 0, sizeof(q));
    q.g = h->g; q.centers = d_centers; q.rgtp = d_rgtp; q.thr = thr; q.nM = nM;
    q.out_n = h->d_out_n; q.out_m = h->d_out_m; q.out_key = h->d_out_key; q.out_off = h->d_out_off;
    q.members = h->d_members; q.md2 = h->want_d2 ? h->d_md2 : nullptr; q.member_cap = h->member_cap;
    q.evals = h->d_u64 + 1; q.flags = h->d_counters + 4; q.mt = h->d_mt;
    CU(cudaMemsetAsync(h->d_counters + 4, 0, 5 * sizeof(uint32_t), s));     /* flags, emit lists + work */
    CU(cudaMemsetAsync(h->d_counters + 15, 0, 2 * sizeof(uint32_t), s));
    CU(cudaMemsetAsync(h->d_u64, 0, sizeof(unsigned long long), s));
    {
        ProfScope p(h, KID_OFFSETS);
        k_offsets<<<1, 1024, 0, s>>>(h->d_out_n, nh, h->d_out_off, h->d_u64 + 0, 2048, h->d_esmall,
                                     h->d_counters + 5, h->d_ebig, h->d_counters + 6, 131072, h->d_ehuge,
                                     h->d_counters + 15);
    }
    q.list = h->d_ehuge; q.list_n = h->d_counters + 15; q.work_counter = h->d_counters + 16;
    { ProfScope p(h, KID_EMIT_HUGE); launch_persistent<1024>(h, k_so_emit<1024>, q, std::min(nh, h->sm_count)); }
    q.list = h->d_esmall; q.list_n = h->d_counters + 5; q.work_counter = h->d_counters + 7;
    { ProfScope p(h, KID_EMIT_WARP); launch_persistent<32>(h, k_so_emit<32>, q, nh); }
    q.list = h->d_ebig; q.list_n = h->d_counters + 6; q.work_counter = h->d_counters + 8;
    { ProfScope p(h, KID_EMIT_BLOCK); launch_persistent<256>(h, k_so_emit<256>, q, nh); }
    CU(cudaGetLastError());
    h->last_h = nh;
    h->have_result = true;
    h->members_sorted = false;
    h->members_csr = true;
    return SOGPU_OK;
}

/* member lists of the last solve in catalog order: d_out_off becomes the CSR offset array (nh + 1 entries) */
static int ensure_csr(sogpu *h)
{
    if (!h->have_result || h->members_csr) return SOGPU_OK;
    const int32_t nh = h->last_h;
    cudaStream_t s = h->stream;
    if (!h->d_members2) {
        CU(cudaMalloc(&h->d_members2, (size_t)h->member_cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_md2_2, (size_t)h->member_cap * sizeof(float)));
    }
    CU(cudaMemsetAsync(h->d_counters + 5, 0, 4 * sizeof(uint32_t), s));
    CU(cudaMemsetAsync(h->d_counters + 15, 0, 2 * sizeof(uint32_t), s));
    {
        ProfScope p(h, KID_OFFSETS, 0.0, 2);
        k_offsets<<<1, 1024, 0, s>>>(h->d_out_n, nh, h->d_csr_off, h->d_u64 + 0, h->emit_small_max, h->d_esmall,
                                     h->d_counters + 5, h->d_ebig, h->d_counters + 6, h->emit_huge_min, h->d_ehuge,
                                     h->d_counters + 15);
        k_compact_members<<<std::min((nh + 7) / 8, h->sm_count * 8), 256, 0, s>>>(
            h->d_out_n, nh, h->d_out_off, h->d_csr_off, h->d_members, h->d_members2,
            h->want_d2 ? h->d_md2 : nullptr, h->d_md2_2);
    }
    CU(cudaGetLastError());
    std::swap(h->d_members, h->d_members2);
    std::swap(h->d_md2, h->d_md2_2);
    std::swap(h->d_out_off, h->d_csr_off);
    h->members_csr = true;
    return SOGPU_OK;
}

static int fetch_stats(sogpu *h)
{
    unsigned long long u[4];
    uint32_t c[24];
    CU(cudaMemcpyAsync(u, h->d_u64, sizeof(u), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(c, h->d_counters, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->stats.last_members = (int64_t)u[0];
    h->stats.last_evals_first = (int64_t)u[1];
    h->stats.last_evals = (int64_t)(u[1] + u[2]);
    h->stats.last_deferred = (int32_t)c[17];
    h->member_overflow = (c[4] & 1u) != 0;
    if (h->member_overflow)
        return set_err(SOGPU_ERR_NOMEM, "member buffer overflow (%llu > %llu)", u[0], h->member_cap);
    if (c[4] & 2u) return set_err(SOGPU_ERR_UNSUPPORTED, "internal: member emission count mismatch");
    return SOGPU_OK;
}

/* the member lists did not fit: grow the buffers to the (now known) total and emit again */
static int grow_members_and_reemit(sogpu *h, const float *d_centers, const float *d_rgtp, int32_t nh, float thr,
                                   int32_t nM)
{
    unsigned long long need = (unsigned long long)h->stats.last_members;
    cudaFree(h->d_members); cudaFree(h->d_md2); cudaFree(h->d_members2); cudaFree(h->d_md2_2);
    h->d_members = nullptr; h->d_md2 = nullptr; h->d_members2 = nullptr; h->d_md2_2 = nullptr;
    h->member_cap = need + need / 16 + 1024;
    CU(cudaMalloc(&h->d_members, (size_t)h->member_cap * sizeof(int32_t)));
    CU(cudaMalloc(&h->d_md2, (size_t)h->member_cap * sizeof(float)));
    int rc = emit_members(h, d_centers, d_rgtp, nh, thr, nM);
    if (rc) return rc;
    return fetch_stats(h);
}

extern "C" int sogpu_so_device(sogpu_t *h, const void *d_centers, const void *d_rgtp, int32_t nh, float thr,
                               int32_t nM, void *d_out_n, void *d_out_m)
{
    if (!h || !d_centers || !d_rgtp) return set_err(SOGPU_ERR_ARG, "sogpu_so_device: NULL argument");
    CU(cudaSetDevice(h->device));
    int rc = run_query(h, (const float *)d_centers, (const float *)d_rgtp, nh, thr, nM);
    if (rc) return rc;
    if (d_out_n) CU(cudaMemcpyAsync(d_out_n, h->d_out_n, (size_t)nh * sizeof(int32_t), cudaMemcpyDeviceToDevice, h->stream));
    if (d_out_m) CU(cudaMemcpyAsync(d_out_m, h->d_out_m, (size_t)nh * sizeof(float), cudaMemcpyDeviceToDevice, h->stream));
    return SOGPU_OK;
}

extern "C" int sogpu_so(sogpu_t *h, const float *centers, const float *rgtp, int32_t nh, float thr, int32_t nM,
                        float *rvir, float *mvir, int32_t *ndelta)
{
    if (!h || !centers || !rgtp || !rvir || !mvir || !ndelta) return set_err(SOGPU_ERR_ARG, "sogpu_so: NULL argument");
    if (nh <= 0) return set_err(SOGPU_ERR_ARG, "no halos");
    CU(cudaSetDevice(h->device));
    int rc = ensure_query(h, nh);
    if (rc) return rc;
    size_t bytes_in = (size_t)nh * 4 * sizeof(float);
    size_t bytes_out = (size_t)nh * (sizeof(int32_t) + sizeof(float));
    rc = ensure_pinned(h, bytes_in + bytes_out);
    if (rc) return rc;
    float *pc = (float *)h->h_pin, *pr = pc + (size_t)3 * nh;
    int32_t *pn = (int32_t *)(pr + nh);
    float *pm = (float *)(pn + nh);
    memcpy(pc, centers, (size_t)nh * 3 * sizeof(float));
    memcpy(pr, rgtp, (size_t)nh * sizeof(float));
    CU(cudaMemcpyAsync(h->d_centers, pc, (size_t)nh * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_rgtp, pr, (size_t)nh * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    rc = run_query(h, h->d_centers, h->d_rgtp, nh, thr, nM);
    if (rc) return rc;
    CU(cudaMemcpyAsync(pn, h->d_out_n, (size_t)nh * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(pm, h->d_out_m, (size_t)nh * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    rc = fetch_stats(h);   /* synchronises the stream */
    if (rc && h->member_overflow) rc = grow_members_and_reemit(h, h->d_centers, h->d_rgtp, nh, thr, nM);
    if (rc) return rc;
    if (h->focused) {
        /* a focused grid (sogpu_build_grid_for) that some ball outgrew, or mixed masses: build the
         * full grid and solve again — results never depend on how the grid was built */
        bool redo = (pn[0] == CODE_UNEQUAL_MASS);
        for (int32_t i = 0; i < nh && !redo; ++i) redo = (pn[i] == CODE_NEED_FULL);
        if (redo && h->indexed)
            return set_err(SOGPU_ERR_UNSUPPORTED, "a halo outgrew the focus mask of this rank's share: route again with a "
                                                  "larger n_balls (sogpu_so_device reports such halos with code -103)");
        if (redo) {
            rc = build_grid_impl(h, 0, 0);
            if (rc) return rc;
            rc = run_query(h, h->d_centers, h->d_rgtp, nh, thr, nM);
            if (rc) return rc;
            CU(cudaMemcpyAsync(pn, h->d_out_n, (size_t)nh * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaMemcpyAsync(pm, h->d_out_m, (size_t)nh * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
            rc = fetch_stats(h);
            if (rc && h->member_overflow) rc = grow_members_and_reemit(h, h->d_centers, h->d_rgtp, nh, thr, nM);
            if (rc) return rc;
        }
    }
    if (pn[0] == CODE_UNEQUAL_MASS) {
        /* mixed particle masses: the rank-only mass table does not apply; run the general path */
        h->mass_state = 0; h->stats.equal_mass = 0;
        rc = run_query_general(h, h->d_centers, h->d_rgtp, nh, thr, nM);
        if (rc) return rc;
        CU(cudaMemcpyAsync(pn, h->d_out_n, (size_t)nh * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaMemcpyAsync(pm, h->d_out_m, (size_t)nh * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        rc = fetch_stats(h);
        if (rc && h->member_overflow) rc = grow_members_and_reemit(h, h->d_centers, h->d_rgtp, nh, thr, nM);
        if (rc) return rc;
    } else {
        /* degenerate particle configurations the histogram levels cannot split (thousands of
         * particles at one identical r^2): those halos go through the general full-sort path */
        int32_t n_bad = 0;
        for (int32_t i = 0; i < nh; ++i) n_bad += (pn[i] == CODE_UNSUPPORTED);
        if (n_bad && h->indexed)
            return set_err(SOGPU_ERR_UNSUPPORTED, "%d halo(s) need the general full-sort path, which reads particle masses "
                                                  "by index and is not available on one rank's share of a domain run", n_bad);
        if (n_bad) {
            CU(cudaMemsetAsync(h->d_counters + 19, 0, sizeof(uint32_t), h->stream));
            k_select_code<<<(nh + 255) / 256, 256, 0, h->stream>>>(h->d_out_n, nh, CODE_UNSUPPORTED, h->d_defer,
                                                                   h->d_counters + 19);
            rc = run_query_general(h, h->d_centers, h->d_rgtp, nh, thr, nM, h->d_defer, (uint32_t)n_bad);
            if (rc) return rc;
            CU(cudaMemcpyAsync(pn, h->d_out_n, (size_t)nh * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
            CU(cudaMemcpyAsync(pm, h->d_out_m, (size_t)nh * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
            rc = fetch_stats(h);
            if (rc && h->member_overflow) rc = grow_members_and_reemit(h, h->d_centers, h->d_rgtp, nh, thr, nM);
            if (rc) return rc;
        }
    }
    for (int32_t i = 0; i < nh; ++i) {
        int32_t n = pn[i];
        if (n > 0) {
            ndelta[i] = n;
            mvir[i] = pm[i];
            rvir[i] = so_rdelta_host(pm[i], thr);                         /* kd2.c:817-820 */
        } else if (n == -1 || n == -2 || n == -3) {
            ndelta[i] = 0;
            mvir[i] = rvir[i] = (float)n;                                 /* kd2.c:774-776,793-795,837-838 */
        } else {
            return set_err(SOGPU_ERR_UNSUPPORTED, "halo %d: unsupported particle configuration (code %d)", i, n);
        }
    }
    return SOGPU_OK;
}

/* convert the packed N_Delta/code array of sogpu_so_device into rvir/mvir/ndelta on the host */
extern "C" int sogpu_finish_host(const int32_t *code_or_n, const float *m, int32_t nh, float thr, float *rvir,
                                 float *mvir, int32_t *ndelta)
{
    if (!code_or_n || !m || !rvir || !mvir || !ndelta) return set_err(SOGPU_ERR_ARG, "sogpu_finish_host: NULL argument");
    for (int32_t i = 0; i < nh; ++i) {
        int32_t n = code_or_n[i];
        if (n > 0) { ndelta[i] = n; mvir[i] = m[i]; rvir[i] = so_rdelta_host(m[i], thr); }
        else if (n == -1 || n == -2 || n == -3) { ndelta[i] = 0; mvir[i] = rvir[i] = (float)n; }
        else if (n == CODE_UNEQUAL_MASS)
            return set_err(SOGPU_ERR_UNSUPPORTED, "particles have unequal masses: use sogpu_so (host entry point), "
                                                  "which switches to the general sequential-mass path");
        else if (n == CODE_NEED_FULL)
            return set_err(SOGPU_ERR_UNSUPPORTED, "halo %d outgrew the focused grid (sogpu_build_grid_for): rebuild "
                                                  "with sogpu_build_grid, or use sogpu_so which does it itself", i);
        else return set_err(SOGPU_ERR_UNSUPPORTED, "halo %d: code %d", i, n);
    }
    return SOGPU_OK;
}

/* sort every member list of the last result by (r^2, index) on the device (segmented sort) */
static int sort_members_device(sogpu *h, int32_t nh, size_t tot)
{
    if (tot == 0) return SOGPU_OK;
    int rc = gen_scratch(h, tot);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    const int grid = (int)std::min<size_t>((tot + 255) / 256, (size_t)h->sm_count * 16);
    k_member_keys<<<grid, 256, 0, s>>>(h->d_members, h->d_md2, tot, h->d_gkeys[0]);
    unsigned long long *sorted_keys = nullptr;
    rc = seg_sort_keys(h, nh, h->d_out_off, tot, &sorted_keys);
    if (rc) return rc;
    k_member_unkeys<<<grid, 256, 0, s>>>(sorted_keys, tot, h->d_members, h->d_md2);
    CU(cudaGetLastError());
    return SOGPU_OK;
}

extern "C" int sogpu_members(sogpu_t *h, int64_t *offsets, const int32_t **members, const float **d2, int sorted)
{
    if (!h || !offsets || !members) return set_err(SOGPU_ERR_ARG, "sogpu_members: NULL argument");
    if (!h->have_result) return set_err(SOGPU_ERR_ARG, "sogpu_members: no sogpu_so result available");
    if ((sorted || d2) && !h->want_d2)
        return set_err(SOGPU_ERR_ARG, "sogpu_members: call sogpu_keep_member_d2(h,1) before sogpu_so to get r^2 / sorted lists");
    CU(cudaSetDevice(h->device));
    const int32_t nh = h->last_h;
    int rc = fetch_stats(h);
    if (rc && h->member_overflow)      /* (asynchronous solves cannot retry by themselves: sogpu_so_device, domain steps) */
        rc = grow_members_and_reemit(h, h->d_centers, h->d_rgtp, nh, h->last_thr, h->last_nM);
    if (rc) return rc;
    rc = ensure_csr(h);
    if (rc) return rc;
    const size_t tot = (size_t)h->stats.last_members;
    const bool dbg = getenv("SOGPU_DEBUG_TIMING") != nullptr;
    auto now = [] { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    double t0 = now();
    if (tot + 1 > h->h_members_cap) {
        if (h->h_members) cudaFreeHost(h->h_members);
        if (h->h_md2) cudaFreeHost(h->h_md2);
        h->h_members = nullptr; h->h_md2 = nullptr; h->h_members_cap = 0;
        size_t cap = std::max<size_t>(tot + 1, 1 << 16);
        CU(cudaMallocHost((void **)&h->h_members, cap * sizeof(int32_t)));
        CU(cudaMallocHost((void **)&h->h_md2, cap * sizeof(float)));
        h->h_members_cap = cap;
    }
    static_assert(sizeof(unsigned long long) == sizeof(int64_t), "offset type");
    if (dbg) { fprintf(stderr, "    [members] pinned alloc %.3f ms (tot %zu)\n", now() - t0, tot); t0 = now(); }
    if (sorted && tot && !h->members_sorted) {
        /* ascending (fDist2, index): the order kdTagParticles walks the list (kd2.c:670,781) */
        rc = sort_members_device(h, nh, tot);
        if (rc) return rc;
        h->members_sorted = true;
        if (dbg) { cudaStreamSynchronize(h->stream); fprintf(stderr, "    [members] device sort %.3f ms\n", now() - t0); t0 = now(); }
    }
    CU(cudaMemcpyAsync(offsets, h->d_out_off, ((size_t)nh + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
    if (tot) {
        CU(cudaMemcpyAsync(h->h_members, h->d_members, tot * sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
        if (h->want_d2)
            CU(cudaMemcpyAsync(h->h_md2, h->d_md2, tot * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    if (dbg) fprintf(stderr, "    [members] D2H %.3f ms\n", now() - t0);
    *members = h->h_members;
    if (d2) *d2 = h->h_md2;
    return SOGPU_OK;
}

extern "C" int sogpu_keep_member_d2(sogpu_t *h, int on)
{
    if (!h) return set_err(SOGPU_ERR_ARG, "NULL handle");
    h->want_d2 = on != 0;
    return SOGPU_OK;
}

extern "C" int sogpu_ball_gather(sogpu_t *h, const float center[3], float ball2, int32_t *idx, float *d2,
                                 int64_t cap, int64_t *n)
{
    if (!h || !center || !n) return set_err(SOGPU_ERR_ARG, "sogpu_ball_gather: NULL argument");
    if (!h->built) return set_err(SOGPU_ERR_ARG, "sogpu_ball_gather: call sogpu_build_grid first");
    if (!(ball2 >= 0.0f) || !(ball2 < INFINITY)) return set_err(SOGPU_ERR_ARG, "bad ball2");
    CU(cudaSetDevice(h->device));
    if (h->indexed) return set_err(SOGPU_ERR_UNSUPPORTED, "sogpu_ball_gather: not available on one rank's share of a domain run");
    if (h->focused) { int rf = build_grid_impl(h, 0, 0); if (rf) return rf; }
    int rc = ensure_query(h, 1);
    if (rc) return rc;
    cudaStream_t s = h->stream;
    h->stats.last_kernel_launches = 0;
    CU(cudaMemsetAsync(h->d_u64 + 3, 0, sizeof(unsigned long long), s));
    {
        ProfScope p(h, KID_BALL_GATHER);
        k_ball_gather<<<h->sm_count * 2, 256, 0, s>>>(h->g, center[0], center[1], center[2], ball2, h->d_members,
                                                     h->d_md2, h->d_u64 + 3, h->member_cap);
    }
    CU(cudaGetLastError());
    unsigned long long cnt = 0;
    CU(cudaMemcpyAsync(&cnt, h->d_u64 + 3, sizeof(cnt), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    h->have_result = false;
    *n = (int64_t)cnt;
    if (cnt > h->member_cap) return set_err(SOGPU_ERR_NOMEM, "ball holds %llu particles, buffer %llu", cnt, h->member_cap);
    if (cnt && (idx || d2) && cap > 0) {
        std::vector<int32_t> ti(cnt);
        std::vector<float> td(cnt);
        CU(cudaMemcpy(ti.data(), h->d_members, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(td.data(), h->d_md2, cnt * sizeof(float), cudaMemcpyDeviceToHost));
        std::vector<std::pair<uint64_t, uint32_t>> tmp(cnt);
        for (size_t k = 0; k < cnt; ++k) {
            uint32_t bits;
            memcpy(&bits, &td[k], 4);
            tmp[k].first = ((uint64_t)bits << 32) | (uint32_t)ti[k];
            tmp[k].second = (uint32_t)k;
        }
        std::sort(tmp.begin(), tmp.end());                                /* qsort(CmpList), kd2.c:781 */
        size_t m = std::min<size_t>(cnt, (size_t)cap);
        for (size_t k = 0; k < m; ++k) {
            if (idx) idx[k] = ti[tmp[k].second];
            if (d2) d2[k] = td[tmp[k].second];
        }
    }
    return SOGPU_OK;
}

/* Batched smBallGather (smooth2.c:58-114): for each of nh balls all particles with
 * fDist2 <= ball2[i], as CSR lists fetched with sogpu_members (d2 kept, sorted on request). */
extern "C" int sogpu_ball_gather_batch(sogpu_t *h, const float *centers, const float *ball2, int32_t nh)
{
    if (!h || !centers || !ball2 || nh <= 0) return set_err(SOGPU_ERR_ARG, "sogpu_ball_gather_batch: bad argument");
    if (!h->built) return set_err(SOGPU_ERR_ARG, "sogpu_ball_gather_batch: call sogpu_build_grid first");
    CU(cudaSetDevice(h->device));
    if (h->indexed) return set_err(SOGPU_ERR_UNSUPPORTED, "sogpu_ball_gather_batch: not available on one rank's share of a domain run");
    if (h->focused) { int rf = build_grid_impl(h, 0, 0); if (rf) return rf; }   /* arbitrary balls need every particle */
    int rc = ensure_query(h, nh);
    if (rc) return rc;
    rc = ensure_pinned(h, (size_t)nh * 4 * sizeof(float));
    if (rc) return rc;
    cudaStream_t s = h->stream;
    float *pc = (float *)h->h_pin, *pb = pc + (size_t)3 * nh;
    memcpy(pc, centers, (size_t)nh * 3 * sizeof(float));
    memcpy(pb, ball2, (size_t)nh * sizeof(float));
    CU(cudaMemcpyAsync(h->d_centers, pc, (size_t)nh * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->d_rgtp, pb, (size_t)nh * sizeof(float), cudaMemcpyHostToDevice, s));
    h->stats.last_kernel_launches = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        CU(cudaMemsetAsync(h->d_counters, 0, 24 * sizeof(uint32_t), s));
        CU(cudaMemsetAsync(h->d_u64, 0, 4 * sizeof(unsigned long long), s));
        QueryArgs a;
        memset(&a, 0, sizeof(a));
        a.g = h->g;
        a.centers = h->d_centers; a.rgtp = h->d_rgtp;
        a.out_n = h->d_out_n; a.out_m = h->d_out_m; a.out_key = h->d_out_key; a.out_off = h->d_out_off;
        a.members = h->d_members; a.md2 = h->d_md2; a.member_cap = h->member_cap;
        a.evals = h->d_u64 + 1; a.flags = h->d_counters + 4; a.mt = h->d_mt;
        a.list = h->d_small; a.list_n = h->d_counters + 0; a.work_counter = h->d_counters + 2;
        k_iota<<<(nh + 255) / 256, 256, 0, s>>>(h->d_small, h->d_counters + 0, nh);
        { ProfScope p(h, KID_BALL_GATHER); launch_persistent<32>(h, k_ball_count, a, nh, h->d_rgtp); }
        {
            ProfScope p(h, KID_OFFSETS);
            k_offsets<<<1, 1024, 0, s>>>(h->d_out_n, nh, h->d_out_off, h->d_u64 + 0, 2048, h->d_esmall,
                                         h->d_counters + 5, h->d_ebig, h->d_counters + 6);
        }
        a.list = h->d_esmall; a.list_n = h->d_counters + 5; a.work_counter = h->d_counters + 7;
        { ProfScope p(h, KID_EMIT_WARP); launch_persistent<32>(h, k_so_emit<32>, a, nh); }
        a.list = h->d_ebig; a.list_n = h->d_counters + 6; a.work_counter = h->d_counters + 8;
        { ProfScope p(h, KID_EMIT_BLOCK); launch_persistent<256>(h, k_so_emit<256>, a, nh); }
        CU(cudaGetLastError());
        unsigned long long tot = 0;
        CU(cudaMemcpyAsync(&tot, h->d_u64, sizeof(tot), cudaMemcpyDeviceToHost, s));
        CU(cudaStreamSynchronize(s));
        if (getenv("SOGPU_DEBUG_TIMING")) fprintf(stderr, "    [sogpu] ball batch attempt %d: %llu entries (cap %llu)\n", attempt, tot, h->member_cap);
        if (tot <= h->member_cap) break;
        if (attempt == 1) return set_err(SOGPU_ERR_NOMEM, "ball lists need %llu entries", tot);
        cudaFree(h->d_members); cudaFree(h->d_md2); cudaFree(h->d_members2); cudaFree(h->d_md2_2);
        h->d_members = nullptr; h->d_md2 = nullptr; h->d_members2 = nullptr; h->d_md2_2 = nullptr;
        h->member_cap = tot + tot / 8;
        CU(cudaMalloc(&h->d_members, (size_t)h->member_cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_md2, (size_t)h->member_cap * sizeof(float)));
    }
    h->last_h = nh;
    h->have_result = true;
    h->members_sorted = false;
    h->members_csr = true;
    h->want_d2 = true;
    return SOGPU_OK;
}

/* kdVcirc / kdMassProfile for nh groups (kd2.c:498-586): gather the 2*Rvir balls, sort every list by
 * (r^2, index) on the device, evaluate the circular-velocity curve, R(M/4), R(M/2), (Rmax, Vmax) and the
 * all-particle mass profile per group.  Equal particle masses only (SOGPU_ERR_UNSUPPORTED otherwise:
 * with mixed masses the cumulative fp32 mass depends on which particle sits at which rank). */
extern "C" int sogpu_vcirc(sogpu_t *h, const float *centers, const float *rvir, const float *mvir, int32_t nh, float G,
                           int32_t nMembers, float *vcirc, float *rmass, float *rmax, float *vmax, float *profile)
{
    if (!h || !centers || !rvir || !mvir || !vcirc || !rmass || !rmax || !vmax || nh <= 0 || nMembers < 1)
        return set_err(SOGPU_ERR_ARG, "sogpu_vcirc: bad argument");
    if (!h->built) return set_err(SOGPU_ERR_ARG, "sogpu_vcirc: call sogpu_build_grid first");
    CU(cudaSetDevice(h->device));
    int rc = fetch_mass_state(h);
    if (rc) return rc;
    if (h->mass_state != 1) return set_err(SOGPU_ERR_UNSUPPORTED, "sogpu_vcirc: particles have unequal masses");
    std::vector<float> ball2((size_t)nh);
    for (int32_t i = 0; i < nh; ++i) {
        float fBall = (float)(2. * rvir[i]);                              /* kd2.c:511-512 */
        ball2[i] = fBall * fBall;
    }
    DbgTimer dt(h->stream);
    rc = sogpu_ball_gather_batch(h, centers, ball2.data(), nh);
    if (rc) return rc;
    dt.mark("vcirc: ball gather batch");
    rc = fetch_stats(h);
    if (rc) return rc;
    const size_t tot = (size_t)h->stats.last_members;
    rc = sort_members_device(h, nh, tot);
    if (rc) return rc;
    h->members_sorted = true;
    dt.mark("vcirc: segmented sort");
    /* per-group inputs and outputs share one device block: rvir, mvir | vcirc 8, rmass 2, rmax, vmax, profile 16 */
    const size_t per = 2 + SO_NVCIRC + 2 + 1 + 1 + SO_NMASSPROFILE;
    if ((size_t)nh * per > h->vc_cap) {
        cudaFree(h->d_vc); h->d_vc = nullptr; h->vc_cap = 0;
        CU(cudaMalloc(&h->d_vc, (size_t)nh * per * sizeof(float)));
        h->vc_cap = (size_t)nh * per;
    }
    rc = ensure_pinned(h, (size_t)nh * per * sizeof(float));
    if (rc) return rc;
    cudaStream_t s = h->stream;
    float *pin = (float *)h->h_pin;
    memcpy(pin, rvir, (size_t)nh * sizeof(float));
    memcpy(pin + nh, mvir, (size_t)nh * sizeof(float));
    CU(cudaMemcpyAsync(h->d_vc, pin, (size_t)nh * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    VcircArgs a;
    a.d2 = h->d_md2; a.off = h->d_out_off; a.rvir = h->d_vc; a.mvir = h->d_vc + nh; a.mt = h->d_mt;
    a.G = G; a.nM = nMembers; a.nh = nh;
    a.vcirc = h->d_vc + (size_t)2 * nh;
    a.rmass = a.vcirc + (size_t)SO_NVCIRC * nh;
    a.rmax = a.rmass + (size_t)2 * nh;
    a.vmax = a.rmax + nh;
    a.profile = profile ? a.vmax + nh : nullptr;
    dt.mark("vcirc: buffers + upload");
    { ProfScope p(h, KID_VCIRC); k_vcirc<<<std::min(nh, h->sm_count * 16), 256, 0, s>>>(a); }
    CU(cudaGetLastError());
    dt.mark("vcirc: k_vcirc");
    const size_t n_out = (size_t)nh * (per - 2);
    CU(cudaMemcpyAsync(pin + (size_t)2 * nh, h->d_vc + (size_t)2 * nh, n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const float *o = pin + (size_t)2 * nh;
    memcpy(vcirc, o, (size_t)SO_NVCIRC * nh * sizeof(float)); o += (size_t)SO_NVCIRC * nh;
    memcpy(rmass, o, (size_t)2 * nh * sizeof(float)); o += (size_t)2 * nh;
    memcpy(rmax, o, (size_t)nh * sizeof(float)); o += nh;
    memcpy(vmax, o, (size_t)nh * sizeof(float)); o += nh;
    if (profile) memcpy(profile, o, (size_t)SO_NMASSPROFILE * nh * sizeof(float));
    return SOGPU_OK;
}

/* kdVcirc / kdMassProfile for snapshots with unequal masses and / or several species (kd2.c:458-496, 498-586):
 * ptype[i] (host, N bytes, may be NULL) holds the species bits of particle i, masks[k] selects the particles that
 * count for profile k; profiles = nmasks x nh x 16 floats. */
extern "C" int sogpu_vcirc_species(sogpu_t *h, const float *centers, const float *rvir, const float *mvir, int32_t nh,
                                   float G, int32_t nMembers, const unsigned char *ptype, const int32_t *masks,
                                   int32_t nmasks, float *vcirc, float *rmass, float *rmax, float *vmax, float *profiles)
{
    if (!h || !centers || !rvir || !mvir || !vcirc || !rmass || !rmax || !vmax || nh <= 0 || nMembers < 1 || nmasks < 0 ||
        nmasks > VC_MAXMASK || (nmasks > 0 && (!masks || !profiles)))
        return set_err(SOGPU_ERR_ARG, "sogpu_vcirc_species: bad argument (at most %d species masks)", VC_MAXMASK);
    if (!h->built) return set_err(SOGPU_ERR_ARG, "sogpu_vcirc_species: call sogpu_build_grid first");
    if (h->indexed) return set_err(SOGPU_ERR_UNSUPPORTED, "sogpu_vcirc_species: not available on one rank's share of a domain run");
    CU(cudaSetDevice(h->device));
    std::vector<float> ball2((size_t)nh);
    for (int32_t i = 0; i < nh; ++i) {
        float fBall = (float)(2. * rvir[i]);                              /* kd2.c:511-512 */
        ball2[i] = fBall * fBall;
    }
    int rc = sogpu_ball_gather_batch(h, centers, ball2.data(), nh);
    if (rc) return rc;
    rc = fetch_stats(h);
    if (rc) return rc;
    const size_t tot = (size_t)h->stats.last_members;
    rc = sort_members_device(h, nh, tot);
    if (rc) return rc;
    h->members_sorted = true;
    cudaStream_t s = h->stream;
    if (ptype) {
        if (h->n > h->ptype_cap) {
            cudaFree(h->d_ptype); h->d_ptype = nullptr; h->ptype_cap = 0;
            CU(cudaMalloc(&h->d_ptype, (size_t)h->n));
            h->ptype_cap = h->n;
        }
        CU(cudaMemcpyAsync(h->d_ptype, ptype, (size_t)h->n, cudaMemcpyHostToDevice, s));
    }
    const size_t need = (size_t)(1 + nmasks) * std::max<size_t>(tot, 1);
    if (need > h->vcC_cap) {
        cudaFree(h->d_vcC); h->d_vcC = nullptr; h->vcC_cap = 0;
        CU(cudaMalloc(&h->d_vcC, need * sizeof(float)));
        h->vcC_cap = need;
    }
    const size_t per = 2 + SO_NVCIRC + 2 + 1 + 1 + (size_t)SO_NMASSPROFILE * VC_MAXMASK;
    if ((size_t)nh * per > h->vc_cap) {
        cudaFree(h->d_vc); h->d_vc = nullptr; h->vc_cap = 0;
        CU(cudaMalloc(&h->d_vc, (size_t)nh * per * sizeof(float)));
        h->vc_cap = (size_t)nh * per;
    }
    rc = ensure_pinned(h, (size_t)nh * per * sizeof(float));
    if (rc) return rc;
    float *pin = (float *)h->h_pin;
    memcpy(pin, rvir, (size_t)nh * sizeof(float));
    memcpy(pin + nh, mvir, (size_t)nh * sizeof(float));
    CU(cudaMemcpyAsync(h->d_vc, pin, (size_t)nh * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    uint32_t masks4 = 0;
    for (int k = 0; k < nmasks; ++k) masks4 |= ((uint32_t)masks[k] & 0xFFu) << (8 * k);
    VcircGenArgs a;
    a.d2 = h->d_md2; a.off = h->d_out_off; a.C = h->d_vcC; a.stride = (unsigned long long)std::max<size_t>(tot, 1);
    a.rvir = h->d_vc; a.mvir = h->d_vc + nh; a.G = G; a.nM = nMembers; a.nh = nh; a.nmask = nmasks;
    a.vcirc = h->d_vc + (size_t)2 * nh;
    a.rmass = a.vcirc + (size_t)SO_NVCIRC * nh;
    a.rmax = a.rmass + (size_t)2 * nh;
    a.vmax = a.rmax + nh;
    a.profiles = a.vmax + nh;
    {
        ProfScope p(h, KID_VCIRC, 0.0, 2);
        k_vc_prefix<<<std::min((nh + 7) / 8, h->sm_count * 8), 256, 0, s>>>(h->d_out_off, h->d_members, nh, h->d_in,
                                                                            ptype ? h->d_ptype : nullptr, nmasks, masks4, a.stride, h->d_vcC);
        k_vcirc_gen<<<std::min(nh, h->sm_count * 16), 256, 0, s>>>(a);
    }
    CU(cudaGetLastError());
    const size_t n_out = (size_t)nh * (SO_NVCIRC + 2 + 1 + 1) + (size_t)nmasks * nh * SO_NMASSPROFILE;
    CU(cudaMemcpyAsync(pin + (size_t)2 * nh, h->d_vc + (size_t)2 * nh, n_out * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    const float *o = pin + (size_t)2 * nh;
    memcpy(vcirc, o, (size_t)SO_NVCIRC * nh * sizeof(float)); o += (size_t)SO_NVCIRC * nh;
    memcpy(rmass, o, (size_t)2 * nh * sizeof(float)); o += (size_t)2 * nh;
    memcpy(rmax, o, (size_t)nh * sizeof(float)); o += nh;
    memcpy(vmax, o, (size_t)nh * sizeof(float)); o += nh;
    if (nmasks) memcpy(profiles, o, (size_t)nmasks * nh * SO_NMASSPROFILE * sizeof(float));
    return SOGPU_OK;
}

/* Conflict detection + tagging of the conflict-free groups of the last sogpu_so() result (see k_tag_claim). */
extern "C" int sogpu_tag_members(sogpu_t *h, const int32_t *index, int32_t nh, unsigned char *in_conflict,
                                 int32_t *igrp)
{
    if (!h || !index || !in_conflict || nh <= 0) return set_err(SOGPU_ERR_ARG, "sogpu_tag_members: bad argument");
    if (!h->have_result || h->last_h != nh)
        return set_err(SOGPU_ERR_ARG, "sogpu_tag_members: needs the member lists of a sogpu_so() call over the same %d groups", nh);
    CU(cudaSetDevice(h->device));
    { int rc0 = ensure_csr(h); if (rc0) return rc0; }
    cudaStream_t s = h->stream;
    if (h->n > h->tag_cap) {
        cudaFree(h->d_tag); h->d_tag = nullptr; h->tag_cap = 0;
        CU(cudaMalloc(&h->d_tag, (size_t)h->n * sizeof(int32_t)));
        h->tag_cap = h->n;
    }
    if (nh > h->tagh_cap) {
        cudaFree(h->d_tag_index); cudaFree(h->d_dirty);
        h->d_tag_index = nullptr; h->d_dirty = nullptr; h->tagh_cap = 0;
        CU(cudaMalloc(&h->d_tag_index, (size_t)nh * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_dirty, (size_t)nh));
        h->tagh_cap = nh;
    }
    int rc = ensure_pinned(h, (size_t)nh * (sizeof(int32_t) + 1));
    if (rc) return rc;
    int32_t *pi = (int32_t *)h->h_pin;
    unsigned char *pd = (unsigned char *)(pi + nh);
    memcpy(pi, index, (size_t)nh * sizeof(int32_t));
    CU(cudaMemcpyAsync(h->d_tag_index, pi, (size_t)nh * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(h->d_tag, 0, (size_t)h->n * sizeof(int32_t), s));
    CU(cudaMemsetAsync(h->d_dirty, 0, (size_t)nh, s));
    const int grid = std::min((nh + 7) / 8, h->sm_count * 16);
    {
        ProfScope p(h, KID_TAG, 0.0, 2);
        k_tag_claim<<<grid, 256, 0, s>>>(h->d_out_off, h->d_members, nh, h->d_tag, h->d_dirty);
        k_tag_settle<<<grid, 256, 0, s>>>(h->d_out_off, h->d_members, h->d_tag_index, nh, h->d_tag, h->d_dirty);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pd, h->d_dirty, (size_t)nh, cudaMemcpyDeviceToHost, s));
    if (igrp) CU(cudaMemcpyAsync(igrp, h->d_tag, (size_t)h->n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    memcpy(in_conflict, pd, (size_t)nh);
    return SOGPU_OK;
}

/* kdTagParticles for the groups sogpu_tag_members reported in conflict, replayed in order on the device
 * (k_tag_replay).  Must follow sogpu_tag_members over the same nh groups, member lists sorted. */
extern "C" int sogpu_tag_replay(sogpu_t *h, const int32_t *order, int32_t n_order, const int32_t *index, const float *centers,
                                float *rvir, float *mvir, int32_t nh, int32_t max_index, int32_t *igrp, int32_t *nsubsumed,
                                int32_t *nignored, int32_t *groups_removed, int32_t *groups_slurped, unsigned char *still_valid)
{
    if (!h || !order || !index || !centers || !rvir || !mvir || nh <= 0 || n_order < 0 || max_index < 1 || !groups_removed ||
        !groups_slurped || !still_valid)
        return set_err(SOGPU_ERR_ARG, "sogpu_tag_replay: bad argument");
    if (!h->have_result || h->last_h != nh || !h->d_tag || !h->members_csr || !h->members_sorted)
        return set_err(SOGPU_ERR_ARG, "sogpu_tag_replay: call sogpu_members(sorted) and sogpu_tag_members over the same %d groups first", nh);
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    if (h->n > h->replay_cap) {
        cudaFree(h->d_nsub); cudaFree(h->d_nign);
        h->d_nsub = h->d_nign = nullptr; h->replay_cap = 0;
        CU(cudaMalloc(&h->d_nsub, (size_t)h->n * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_nign, (size_t)h->n * sizeof(int32_t)));
        h->replay_cap = h->n;
    }
    /* one block of per-group data: order | index | slot_of_index | pos | rvir | mvir | counters | do_vcirc */
    const size_t n_ord = (size_t)std::max(n_order, 1);
    const size_t words = n_ord + (size_t)nh + ((size_t)max_index + 1) + 3 * (size_t)nh + 2 * (size_t)nh + 4 + ((size_t)nh + 3) / 4;
    if (words * 4 > h->replay_bytes) {
        cudaFree(h->d_replay); h->d_replay = nullptr; h->replay_bytes = 0;
        CU(cudaMalloc(&h->d_replay, words * 4));
        h->replay_bytes = words * 4;
    }
    std::vector<int32_t> host(words, 0);
    int32_t *p_order = host.data(), *p_index = p_order + n_ord, *p_slot = p_index + nh;
    float *p_pos = (float *)(p_slot + max_index + 1), *p_rvir = p_pos + 3 * (size_t)nh, *p_mvir = p_rvir + nh;
    int32_t *p_cnt = (int32_t *)(p_mvir + nh);
    unsigned char *p_dov = (unsigned char *)(p_cnt + 4);
    memcpy(p_order, order, (size_t)n_order * sizeof(int32_t));
    memcpy(p_index, index, (size_t)nh * sizeof(int32_t));
    for (int32_t i = 0; i <= max_index; ++i) p_slot[i] = -1;
    for (int32_t i = 0; i < nh; ++i) {
        if (index[i] < 1 || index[i] > max_index) return set_err(SOGPU_ERR_ARG, "sogpu_tag_replay: catalog id %d out of range", index[i]);
        p_slot[index[i]] = i;
    }
    memcpy(p_pos, centers, (size_t)nh * 3 * sizeof(float));
    memcpy(p_rvir, rvir, (size_t)nh * sizeof(float));
    memcpy(p_mvir, mvir, (size_t)nh * sizeof(float));
    for (int32_t i = 0; i < nh; ++i) p_dov[i] = rvir[i] > 0.0f ? 1 : 0;
    CU(cudaMemcpyAsync(h->d_replay, host.data(), words * 4, cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(h->d_nsub, 0, (size_t)h->n * sizeof(int32_t), s));
    CU(cudaMemsetAsync(h->d_nign, 0, (size_t)h->n * sizeof(int32_t), s));
    int32_t *d = (int32_t *)h->d_replay;
    ReplayArgs a;
    a.order = d; a.n_order = n_order; a.off = h->d_out_off; a.mem = h->d_members;
    a.index = d + n_ord; a.slot_of_index = a.index + nh;
    a.pos = (const float *)(a.slot_of_index + max_index + 1);
    a.rvir = (float *)a.pos + 3 * (size_t)nh; a.mvir = a.rvir + nh;
    a.counters = (int32_t *)(a.mvir + nh);
    a.do_vcirc = (unsigned char *)(a.counters + 4);
    a.tag = h->d_tag; a.nsub = h->d_nsub; a.nign = h->d_nign;
    if (n_order > 0) {
        ProfScope p(h, KID_TAG);
        k_tag_replay<<<1, 256, 0, s>>>(a);
    }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host.data(), h->d_replay, words * 4, cudaMemcpyDeviceToHost, s));
    if (igrp) CU(cudaMemcpyAsync(igrp, h->d_tag, (size_t)h->n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (nsubsumed) CU(cudaMemcpyAsync(nsubsumed, h->d_nsub, (size_t)h->n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (nignored) CU(cudaMemcpyAsync(nignored, h->d_nign, (size_t)h->n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    if (p_cnt[2]) return set_err(SOGPU_ERR_UNSUPPORTED, "kdZeroGroup: zeroed group mass is already negative (kd2.c:626-632)");
    memcpy(rvir, p_rvir, (size_t)nh * sizeof(float));
    memcpy(mvir, p_mvir, (size_t)nh * sizeof(float));
    memcpy(still_valid, p_dov, (size_t)nh);
    *groups_removed = p_cnt[0];
    *groups_slurped = p_cnt[1];
    return SOGPU_OK;
}

/* ---- domain runs: geometry, focus masks, routing (see k_route) ------------------------------------ */

/* geometry of the cell grid and of the coarse mask for a snapshot of n_total particles */
static void domain_geometry(sogpu *h, int64_t n_total, GridDev &g)
{
    int lb;
    const int nc = pick_cells(n_total, h->ppc, &lb);
    memset(&g, 0, sizeof(g));
    g.nc = nc; g.lb = lb; g.tb = std::min(lb, 3);
    double hmax = 0.0, lmin = 1e300;
    for (int k = 0; k < 3; ++k) {
        double Lk = (double)h->period[k];
        g.L[k] = h->period[k];
        g.halfL[k] = 0.5f * h->period[k];
        g.g0[k] = (float)((double)h->center[k] - 0.5 * Lk);
        g.dg0[k] = (double)g.g0[k];
        g.dh[k] = Lk / nc;
        g.invh[k] = (float)((double)nc / Lk);
        g.dinvh[k] = (double)g.invh[k];
        hmax = std::max(hmax, g.dh[k]);
        lmin = std::min(lmin, Lk);
    }
    g.bmax_pruned = 0.5 * lmin - 2.0 * hmax;
    g.mb = std::min(lb, h->mask_bits); g.ms = lb - g.mb;
    g.mask_rmin = h->mask_rmin_cells * hmax * (double)(1 << g.ms);
}

extern "C" int sogpu_domain_mask_words(sogpu_t *h, int64_t n_total, int64_t *words)
{
    if (!h || !words || n_total <= 0) return set_err(SOGPU_ERR_ARG, "sogpu_domain_mask_words: bad argument");
    int lb;
    pick_cells(n_total, h->ppc, &lb);
    *words = (int64_t)(((size_t)1 << (3 * std::min(lb, h->mask_bits))) / 32 + 1);
    return SOGPU_OK;
}

/* the coarse cells the nh halos can reach within n_balls steps of the ball schedule -> d_mask (device) */
extern "C" int sogpu_domain_mask(sogpu_t *h, int64_t n_total, const float period[3], const float center[3],
                                 const float *centers, const float *rgtp, int32_t nh, int32_t n_balls, void *d_mask)
{
    if (!h || !period || !d_mask || nh < 0 || n_balls < 1 || n_total <= 0 || (nh > 0 && (!centers || !rgtp)))
        return set_err(SOGPU_ERR_ARG, "sogpu_domain_mask: bad argument");
    CU(cudaSetDevice(h->device));
    for (int k = 0; k < 3; ++k) { h->period[k] = period[k]; h->center[k] = center ? center[k] : 0.0f; }
    GridDev g;
    domain_geometry(h, n_total, g);
    int64_t words;
    sogpu_domain_mask_words(h, n_total, &words);
    cudaStream_t s = h->stream;
    CU(cudaMemsetAsync(d_mask, 0, (size_t)words * sizeof(uint32_t), s));
    if (nh == 0) return SOGPU_OK;
    int rc = ensure_query(h, nh);
    if (rc) return rc;
    rc = ensure_pinned(h, (size_t)nh * 4 * sizeof(float));
    if (rc) return rc;
    float *pc = (float *)h->h_pin, *pr = pc + (size_t)3 * nh;
    memcpy(pc, centers, (size_t)nh * 3 * sizeof(float));
    memcpy(pr, rgtp, (size_t)nh * sizeof(float));
    CU(cudaMemcpyAsync(h->d_centers, pc, (size_t)nh * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    CU(cudaMemcpyAsync(h->d_rgtp, pr, (size_t)nh * sizeof(float), cudaMemcpyHostToDevice, s));
    g.mask = (const uint32_t *)d_mask;
    { ProfScope p(h, KID_MARK_MASK);
      k_mark_mask<<<std::min((nh + 7) / 8, h->sm_count * 8), 256, 0, s>>>(g, h->d_centers, h->d_rgtp, nh, n_balls, (uint32_t *)d_mask); }
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s));            /* the pinned staging is reused by the next call */
    return SOGPU_OK;
}

static int route_args(sogpu *h, RouteArgs &a, int64_t n_total, const void *d_slice, int64_t n_slice, int64_t index_base,
                      const void *d_masks, int32_t n_ranks, bool build_table)
{
    if (!h || !d_slice || !d_masks || n_slice < 0 || n_ranks < 1 || n_ranks > ROUTE_MAXR || index_base < 0 ||
        index_base + n_slice > 0x7FFFFFF0LL)
        return set_err(SOGPU_ERR_ARG, "sogpu_domain_route: bad argument (at most %d ranks)", ROUTE_MAXR);
    memset(&a, 0, sizeof(a));
    domain_geometry(h, n_total, a.g);
    int64_t words;
    sogpu_domain_mask_words(h, n_total, &words);
    a.slice = (const float4 *)d_slice; a.n = n_slice; a.index_base = (uint32_t)index_base;
    a.R = n_ranks;
    const uint32_t n_cells = 1u << (3 * a.g.mb);
    if (!h->d_route_table) {
        CU(cudaMalloc(&h->d_route_table, ((size_t)1 << 27) * sizeof(unsigned short)));
        CU(cudaMalloc(&h->d_route_any, (((size_t)1 << 27) / 32 + 1) * sizeof(uint32_t)));
    }
    if (build_table)
        k_route_table<<<(unsigned)((words + 255) / 256), 256, 0, h->stream>>>((const uint32_t *)d_masks, (uint32_t)words, n_ranks,
                                                                            n_cells, h->d_route_table, h->d_route_any);
    a.table = h->d_route_table;
    a.any = h->d_route_any;
    if (!h->d_route) CU(cudaMalloc(&h->d_route, ROUTE_MAXR * sizeof(unsigned long long)));
    a.counts = h->d_route;
    return SOGPU_OK;
}

/* how many of this rank's particles each rank needs: counts[n_ranks] (host) */
extern "C" int sogpu_domain_route_count(sogpu_t *h, int64_t n_total, const void *d_slice, int64_t n_slice,
                                        const void *d_masks, int32_t n_ranks, int64_t *counts)
{
    if (!counts) return set_err(SOGPU_ERR_ARG, "sogpu_domain_route_count: NULL counts");
    RouteArgs a;
    int rc = route_args(h, a, n_total, d_slice, n_slice, 0, d_masks, n_ranks, true);
    if (rc) return rc;
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    CU(cudaMemsetAsync(h->d_route, 0, ROUTE_MAXR * sizeof(unsigned long long), s));
    if (n_slice > 0) {
        ProfScope p(h, KID_ROUTE, 16.0 * (double)n_slice);
        k_route<false><<<(int)std::min<int64_t>((n_slice + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, s>>>(a);
    }
    CU(cudaGetLastError());
    unsigned long long c[ROUTE_MAXR];
    CU(cudaMemcpyAsync(c, h->d_route, sizeof(c), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int d = 0; d < n_ranks; ++d) counts[d] = (int64_t)c[d];
    return SOGPU_OK;
}

/* write this rank's particles into the receive buffers: dst[d] (device pointers, possibly peer memory
 * opened with sogpu_peer_open), starting at record dst_offset[d].  Asynchronous on the handle's stream. */
extern "C" int sogpu_domain_route_scatter(sogpu_t *h, int64_t n_total, const void *d_slice, int64_t n_slice,
                                          int64_t index_base, const void *d_masks, int32_t n_ranks,
                                          void *const *dst, const int64_t *dst_offset)
{
    if (!dst || !dst_offset) return set_err(SOGPU_ERR_ARG, "sogpu_domain_route_scatter: NULL destination");
    RouteArgs a;
    int rc = route_args(h, a, n_total, d_slice, n_slice, index_base, d_masks, n_ranks, true);
    if (rc) return rc;
    for (int d = 0; d < n_ranks; ++d) { a.dst[d] = (float4 *)dst[d]; a.dst_off[d] = (unsigned long long)dst_offset[d]; }
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    CU(cudaMemsetAsync(h->d_route, 0, ROUTE_MAXR * sizeof(unsigned long long), s));
    if (n_slice > 0) {
        ProfScope p(h, KID_ROUTE, 16.0 * (double)n_slice);
        k_route<true><<<(int)std::min<int64_t>((n_slice + 255) / 256, (int64_t)h->sm_count * 8), 256, 0, s>>>(a);
    }
    CU(cudaGetLastError());
    return SOGPU_OK;
}

/* one rank's share of a domain run: {x, y, z, global index} records (what sogpu_domain_route_scatter
 * writes), all of mass `mass`; n_total fixes the grid resolution so that every rank uses the same cells */
extern "C" int sogpu_set_particles_device_indexed(sogpu_t *h, const void *d_xyzi, int64_t n_local, int64_t n_total,
                                                  float mass, const float period[3], const float center[3])
{
    if (!h || !d_xyzi || !period || n_total < n_local || !(mass > 0.0f))
        return set_err(SOGPU_ERR_ARG, "sogpu_set_particles_device_indexed: bad argument");
    int rc = set_common(h, n_local, period, center);
    if (rc) return rc;
    h->d_in = (const float4 *)d_xyzi;
    h->indexed = true;
    h->indexed_mass = mass;
    h->n_total = n_total;
    return SOGPU_OK;
}

/* device memory another process of the same node can map (cudaIpc): the receive buffers of the exchange */
extern "C" int sogpu_peer_alloc(sogpu_t *h, size_t bytes, void **ptr, void *handle64)
{
    if (!h || !ptr || !handle64) return set_err(SOGPU_ERR_ARG, "sogpu_peer_alloc: NULL argument");
    CU(cudaSetDevice(h->device));
    CU(cudaMalloc(ptr, bytes ? bytes : 256));
    cudaIpcMemHandle_t ih;
    cudaError_t e = cudaIpcGetMemHandle(&ih, *ptr);
    if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return set_err(SOGPU_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); }
    static_assert(sizeof(ih) == 64, "ipc handle size");
    memcpy(handle64, &ih, 64);
    return SOGPU_OK;
}

extern "C" int sogpu_peer_open(sogpu_t *h, const void *handle64, void **ptr)
{
    if (!h || !ptr || !handle64) return set_err(SOGPU_ERR_ARG, "sogpu_peer_open: NULL argument");
    CU(cudaSetDevice(h->device));
    cudaIpcMemHandle_t ih;
    memcpy(&ih, handle64, 64);
    CU(cudaIpcOpenMemHandle(ptr, ih, cudaIpcMemLazyEnablePeerAccess));
    return SOGPU_OK;
}

extern "C" int sogpu_peer_close(sogpu_t *h, void *ptr)
{
    if (!h || !ptr) return set_err(SOGPU_ERR_ARG, "sogpu_peer_close: NULL argument");
    CU(cudaSetDevice(h->device));
    CU(cudaIpcCloseMemHandle(ptr));
    return SOGPU_OK;
}

extern "C" int sogpu_peer_free(sogpu_t *h, void *ptr)
{
    if (!h) return set_err(SOGPU_ERR_ARG, "NULL handle");
    CU(cudaSetDevice(h->device));
    CU(cudaFree(ptr));
    return SOGPU_OK;
}

#include "domain_step.cuh"

/* debug: device clock (ns) of the first CTA start / last CTA end of the huge, big, small, deferred query kernels */
extern "C" int sogpu_debug_timeline(sogpu_t *h, uint64_t *out16)
{
    if (!h || !out16 || !h->d_timeline) return set_err(SOGPU_ERR_ARG, "sogpu_debug_timeline: run with SOGPU_DEBUG_TIMELINE=1");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(out16, h->d_timeline, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return SOGPU_OK;
}

extern "C" int sogpu_ingest_keep_velocities(sogpu_t *h, int on)
{
    if (!h) return set_err(SOGPU_ERR_ARG, "NULL handle");
    h->ingest_want_vel = on != 0;
    return SOGPU_OK;
}

/* _VcmParticles for every group of the last sogpu_so() call whose mvir[i] > 0 (others: zeros). */
extern "C" int sogpu_vcm(sogpu_t *h, const float *mvir, int32_t nh, float *vcm)
{
    if (!h || !mvir || !vcm || nh <= 0) return set_err(SOGPU_ERR_ARG, "sogpu_vcm: bad argument");
    if (!h->have_result || h->last_h != nh) return set_err(SOGPU_ERR_ARG, "sogpu_vcm: needs the member lists of a sogpu_so() call over the same %d groups", nh);
    if (!h->d_vel || h->d_in != h->d_in_owned || h->indexed)
        return set_err(SOGPU_ERR_ARG, "sogpu_vcm: velocities are not on the device (sogpu_ingest_keep_velocities before sogpu_ingest_begin)");
    if (!h->want_d2) return set_err(SOGPU_ERR_ARG, "sogpu_vcm: call sogpu_keep_member_d2(h,1) before sogpu_so (the sum runs in r^2 order)");
    CU(cudaSetDevice(h->device));
    int rc = fetch_stats(h);
    if (rc) return rc;
    rc = ensure_csr(h);
    if (rc) return rc;
    const size_t tot = (size_t)h->stats.last_members;
    if (!h->members_sorted) {
        rc = sort_members_device(h, nh, tot);
        if (rc) return rc;
        h->members_sorted = true;
    }
    const size_t per = 4;
    if ((size_t)nh * per > h->vc_cap) {
        cudaFree(h->d_vc); h->d_vc = nullptr; h->vc_cap = 0;
        CU(cudaMalloc(&h->d_vc, (size_t)nh * per * sizeof(float)));
        h->vc_cap = (size_t)nh * per;
    }
    rc = ensure_pinned(h, (size_t)nh * per * sizeof(float));
    if (rc) return rc;
    cudaStream_t s = h->stream;
    float *pin = (float *)h->h_pin;
    memcpy(pin, mvir, (size_t)nh * sizeof(float));
    CU(cudaMemcpyAsync(h->d_vc, pin, (size_t)nh * sizeof(float), cudaMemcpyHostToDevice, s));
    { ProfScope p(h, KID_VCIRC); k_vcm<<<(nh + 127) / 128, 128, 0, s>>>(h->d_out_off, h->d_members, h->d_in, h->d_vel, h->d_vc, nh, h->d_vc + nh); }
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pin + nh, h->d_vc + nh, (size_t)nh * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
    for (int32_t i = 0; i < nh; ++i)
        for (int l = 0; l < 3; ++l) vcm[3 * i + l] = mvir[i] > 0.0f ? pin[nh + 3 * (size_t)i + l] : 0.0f;
    return SOGPU_OK;
}

extern "C" int sogpu_get_stats(sogpu_t *h, sogpu_stats_t *out)
{
    if (!h || !out) return set_err(SOGPU_ERR_ARG, "sogpu_get_stats: NULL argument");
    CU(cudaSetDevice(h->device));
    if (h->built) {
        int rc = fetch_mass_state(h);
        if (rc) return rc;
        uint32_t kept = 0;
        CU(cudaMemcpyAsync(&kept, h->d_ce + h->ncell, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        h->stats.n_in_grid = (int64_t)kept;
    }
    if (h->have_result) {
        int rc = fetch_stats(h);
        if (rc) return rc;
    }
    *out = h->stats;
    return SOGPU_OK;
}

/* ---- profiling: CUDA-event time per kernel, on the stream the kernels run on ------------------ */

extern "C" int sogpu_profile_enable(sogpu_t *h, int on)
{
    if (!h) return set_err(SOGPU_ERR_ARG, "NULL handle");
    h->prof_on = on != 0;
    return SOGPU_OK;
}

extern "C" int sogpu_profile_kernels(void) { return KID_N; }

extern "C" const char *sogpu_profile_name(int kid) { return (kid >= 0 && kid < KID_N) ? g_kernel_names[kid] : ""; }

extern "C" int sogpu_profile_read(sogpu_t *h, double *ms, int64_t *launches, int nk, int reset)
{
    if (!h || !ms || !launches) return set_err(SOGPU_ERR_ARG, "sogpu_profile_read: NULL argument");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    for (auto &r : h->prof_pending) {
        float t = 0.0f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            h->prof_ms[r.kid] += (double)t;
            h->prof_launches[r.kid] += r.launches;
        }
        h->prof_pool.push_back(r.a);
        h->prof_pool.push_back(r.b);
    }
    h->prof_pending.clear();
    for (int k = 0; k < nk && k < KID_N; ++k) { ms[k] = h->prof_ms[k]; launches[k] = h->prof_launches[k]; }
    if (reset)
        for (int k = 0; k < KID_N; ++k) { h->prof_ms[k] = 0.0; h->prof_launches[k] = 0; }
    return SOGPU_OK;
}

/* Algorithmic bytes accumulated per kernel slot since the last sogpu_profile_read(reset=1) for the
 * kernels whose traffic is known at launch time (the grid build); 0 for the query kernels, whose
 * bytes are 16 B x r^2 evaluations (sogpu_get_stats). */
extern "C" int sogpu_profile_bytes(sogpu_t *h, double *bytes, int nk, int reset)
{
    if (!h || !bytes) return set_err(SOGPU_ERR_ARG, "sogpu_profile_bytes: NULL argument");
    for (int k = 0; k < nk && k < KID_N; ++k) bytes[k] = h->prof_bytes[k];
    if (reset)
        for (int k = 0; k < KID_N; ++k) h->prof_bytes[k] = 0.0;
    return SOGPU_OK;
}

/* ---- host-only helpers -------------------------------------------------------------------- */

extern "C" int sogpu_mass_prefix(float m, int64_t kmax, const int64_t *k, int64_t nk, float *out)
{
    if (!k || !out || kmax < 0) return set_err(SOGPU_ERR_ARG, "sogpu_mass_prefix: bad argument");
    so_mass_table t;
    if (so_mass_table_build(&t, m, (uint64_t)kmax)) return set_err(SOGPU_ERR_UNSUPPORTED, "mass table overflow");
    for (int64_t i = 0; i < nk; ++i) {
        if (k[i] < 0 || k[i] > kmax) return set_err(SOGPU_ERR_ARG, "k out of range");
        out[i] = so_mass_prefix_eval(t.k0, t.s0, t.inc, t.n, (uint32_t)k[i]);
    }
    return SOGPU_OK;
}

extern "C" int sogpu_ball_schedule(float rgtp, const float period[3], float *balls, int cap)
{
    float root = so_root_period(period[0], period[1], period[2]);
    float ball = rgtp;
    int k = 0;
    while ((double)ball < 0.25 * (double)root) {
        ball = so_next_ball(ball);
        if (k < cap && balls) balls[k] = ball;
        ++k;
        if (!(ball > 0.0f)) break;
    }
    return k;
}

extern "C" float sogpu_rdelta(float mvir, float rho_thr) { return so_rdelta_host(mvir, rho_thr); }
