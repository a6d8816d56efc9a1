/* domain_step.cuh — one step of a DOMAIN RUN (several GPUs) without any host round trip.
 *
 * The path being scaled is so.c:515 (kdBuildTree) + so.c:540 (kdSO -> kdRvir, kd2.c:864-895, 723-840) over ONE
 * snapshot and ONE catalog.  Every rank holds a slice of the particle array and the whole (small) catalog.
 * All of the following is enqueued on the rank's stream; nothing is read back by the host before the results:
 *
 *   k_assign_hist/_scan/_bin_owner/_owner
 *                               owner rank of every halo: 32^3 bins ordered along a tiled curve through the box, cut
 *                               into pieces of equal estimated cost.  Integer arithmetic on identical inputs, so
 *                               every rank computes the same assignment without talking to the others.
 *   k_halo_cubes, k_mark_table  the coarse cells every halo can reach within n_balls steps of kdRvir's ball schedule
 *                               (kd2.c:765-768) -> bitmaps: somebody needs the cell / this rank's own focus mask /
 *                               routing pre-filter; several ranks: cells under "plain" halos (destination = owner of
 *                               the cell's bin) and under halos at an ownership boundary ("listed": destinations in
 *                               a 16-bit table).  Computed redundantly by every rank: no mask exchange.
 *   k_route_stage               ONE streaming pass over the slice: what somebody needs, as {x, y, z, global index}
 *                               records; one rank: into its receive buffer, several: into the hit list
 *   k_route_split               hit list -> one contiguous run per destination and tile, reserved with one
 *                               system-scope atomicAdd on the RECEIVER's cursor and stored straight into the
 *                               receiver's buffer (peer memory, NVLink)
 *   (k_push_reserve, k_push_copy  SOGPU_DIRECT_PUSH=0: local staging runs shipped in bulk instead)
 *   k_dom_barrier               flag barrier through peer memory (release / acquire at system scope)
 *   build (n read on the device) + SO solve of the halos this rank owns.
 *
 * Receive buffers and cursors are double-buffered by step parity: a rank that runs ahead writes into the other set.
 */
#pragma once

#define DOM_ASSIGN_LOG 5
#define DOM_ASSIGN_BINS (1 << (3 * DOM_ASSIGN_LOG))
#define CODE_NOT_MINE ((int32_t)0x80808080)     /* cudaMemset pattern 0x80: halo solved by another rank */

struct DomCtrl {                        /* lives in peer-shareable memory (sogpu_peer_alloc) */
    unsigned long long cursor[2];       /* records reserved so far in recv[parity]            */
    uint32_t flag[ROUTE_MAXR];          /* barrier: last epoch rank r has announced to me     */
    uint32_t pad[12];
};

struct DomainState {
    sogpu_domain_cfg_t cfg;
    float4 *recv[2];
    DomCtrl *ctrl;
    float4 *stage;                      /* n_ranks x stage_cap records (unused slot: own rank) */
    float4 *hits;                       /* several ranks: what k_route_stage keeps of the slice (recv_cap records), input of k_route_split */
    float4 *peer_recv[2][ROUTE_MAXR];
    DomCtrl *peer_ctrl[ROUTE_MAXR];
    bool connected;
    unsigned long long *d_counts;       /* [0..R) staged, [R..2R) push base, [2R..3R) push count, [3 MAXR] received (clamped), [3 MAXR + 1] hits */
    uint32_t *d_flags;                  /* bit0 staging overflow, bit1 own receive overflow, bit2 peer receive overflow, bit3 barrier timeout */
    uint32_t *d_nrecv32;
    unsigned char *d_owner;
    int4 *d_cubes;                      /* per halo: the cube of coarse cells it marks (k_halo_cubes) */
    int32_t owner_cap;
    unsigned short *d_table;            /* destination ranks per coarse cell (2^(3 mb) entries)                */
    uint32_t *d_any, *d_super, *d_mymask; /* bit per coarse cell: somebody's / (pre-filter, per block) / this rank's */
    uint32_t *d_plain, *d_listed;       /* several ranks: cells under a plain / a crossing halo (k_mark_table)   */
    unsigned char *d_bin_owner;         /* DOM_ASSIGN_BINS owners */
    bool marked;                        /* the tables hold the marks of the previous step's catalog            */
    int marked_balls;
    int32_t marked_nh;
    unsigned long long *d_bins;         /* DOM_ASSIGN_BINS + 1 */
    uint32_t epoch;
    int parity, n_balls;
    int32_t nh;
    int64_t n_hint;
    bool routed;
    int route_prefetch;                 /* SOGPU_ROUTE_PREFETCH: L2 prefetch of a warp's next run in k_route_stage */
    bool direct;                        /* k_route_split writes its runs straight into the receivers' buffers (one remote
                                         * reservation per tile and destination) instead of staging them for k_push_copy */
};

/* ---- ownership -------------------------------------------------------------------------------------- */
/* bins in 8x8x8 blocks, blocks row-major, bins row-major inside a block */
__device__ __forceinline__ uint32_t assign_bin_index(uint32_t kx, uint32_t ky, uint32_t kz)
{
    const uint32_t nb = 1u << (DOM_ASSIGN_LOG - 3);
    const uint32_t hx = kx >> 3, hy = ky >> 3, hz = kz >> 3, lx = kx & 7u, ly = ky & 7u, lz = kz & 7u;
    return ((((hz * nb + hy) * nb + hx)) << 9) | (lz << 6) | (ly << 3) | lx;
}
__device__ __forceinline__ uint32_t assign_bin(const float *c, const GridDev &g)
{
    uint32_t k[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        double t = ((double)c[a] - g.dg0[a]) / (double)g.L[a];
        t -= floor(t);
        int ci = (int)(t * (double)(1 << DOM_ASSIGN_LOG));
        k[a] = (uint32_t)min(max(ci, 0), (1 << DOM_ASSIGN_LOG) - 1);
    }
    return assign_bin_index(k[0], k[1], k[2]);
}
/* estimated r^2 evaluations of one halo: the final ball (1.2 R, R ~ 1.25 rgtp) at 200 x the mean number density,
 * plus a floor for the fixed per-halo work (so_b200/parallel.py: halo_cost) */
__device__ __forceinline__ unsigned long long assign_cost(float rgtp, double nbar)
{
    const double r = 1.2 * 1.25 * (double)rgtp;
    const double c = 200.0 * nbar * 4.18879020478639 * r * r * r * 0.6 + 64.0;
    return c < 1.0e15 ? (unsigned long long)c : 1000000000000000ull;
}

__global__ void __launch_bounds__(256) k_assign_hist(GridDev g, const float *__restrict__ centers, const float *__restrict__ rgtp,
                                                     int nh, double nbar, unsigned long long *__restrict__ bins)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    atomicAdd(&bins[assign_bin(centers + 3 * (size_t)h, g)], assign_cost(rgtp[h], nbar));
}

/* exclusive scan of the DOM_ASSIGN_BINS costs by one block; bins[DOM_ASSIGN_BINS] = total.  Warp w owns the
 * contiguous chunk [w * CH, (w + 1) * CH): every load and store of a warp is 256 contiguous bytes. */
__global__ void __launch_bounds__(1024) k_assign_scan(unsigned long long *bins)
{
    constexpr int CH = DOM_ASSIGN_BINS / 32, IT = CH / 32;
    __shared__ unsigned long long ws[32];
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    unsigned long long v[IT], s = 0;
#pragma unroll
    for (int i = 0; i < IT; ++i) { v[i] = bins[w * CH + i * 32 + lane]; s += v[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);      /* chunk total, on every lane */
    if (lane == 0) ws[w] = s;
    __syncthreads();
    if (w == 0) {
        unsigned long long y = ws[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, y, o);
            if (lane >= o) y += u;
        }
        ws[lane] = y;
    }
    __syncthreads();
    unsigned long long carry = w ? ws[w - 1] : 0ull;
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        unsigned long long x = v[i];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += u;
        }
        bins[w * CH + i * 32 + lane] = carry + x - v[i];
        carry += __shfl_sync(0xFFFFFFFFu, x, 31);
    }
    if (t == 1023) bins[DOM_ASSIGN_BINS] = carry;
}

/* owner of every bin: the rank whose share of the total cost holds the middle of the bin's cost interval */
__global__ void __launch_bounds__(256) k_assign_bin_owner(int R, const unsigned long long *__restrict__ bins,
                                                          unsigned char *__restrict__ bin_owner)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= DOM_ASSIGN_BINS) return;
    const unsigned long long lo = bins[b], hi = bins[b + 1], total = bins[DOM_ASSIGN_BINS];
    const unsigned long long mid = lo + (hi - lo) / 2ull;
    unsigned long long r = total ? (mid * (unsigned long long)R) / total : 0ull;
    bin_owner[b] = (unsigned char)(r < (unsigned long long)R ? r : (unsigned long long)(R - 1));
}

__global__ void __launch_bounds__(256) k_assign_owner(GridDev g, const float *__restrict__ centers, int nh,
                                                      const unsigned char *__restrict__ bin_owner,
                                                      unsigned char *__restrict__ owner)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    owner[h] = bin_owner[assign_bin(centers + 3 * (size_t)h, g)];
}

/* ---- destination table -------------------------------------------------------------------------------- */
/* pre-filter of the routing pass: one bit per block of 8 x 4 x 4 coarse cells (64 x 128 x 128 blocks at the 512^3
 * mask resolution: 128 KB, one copy in the shared memory of every SM).  The first version (64^3 blocks, 32 KB)
 * let 60 % of a uniform background through to the bitmap lookup in L2; this one 26 % (profiles/r2_experiments.md). */
#define DOM_SUPER_LOGX 6
#define DOM_SUPER_LOGY 7
#define DOM_SUPER_LOGZ 7
struct SuperGeom { int bx, by, bz, sx, sy, sz; };       /* log2 blocks per axis, log2 coarse cells per block */
__host__ __device__ __forceinline__ SuperGeom super_geom(int mb)
{
    SuperGeom q;
    q.bx = mb < DOM_SUPER_LOGX ? mb : DOM_SUPER_LOGX; q.by = mb < DOM_SUPER_LOGY ? mb : DOM_SUPER_LOGY;
    q.bz = mb < DOM_SUPER_LOGZ ? mb : DOM_SUPER_LOGZ;
    q.sx = mb - q.bx; q.sy = mb - q.by; q.sz = mb - q.bz;
    return q;
}
static inline size_t super_words(int mb)                 /* 32-bit words, a multiple of 4 */
{
    const SuperGeom q = super_geom(mb);
    return ((((size_t)1 << (q.bx + q.by + q.bz)) / 32) + 4) & ~(size_t)3;
}

/* run of `len` bits starting at bit `b0` of a bitmap: OR it in (or clear the words) — at most len/32 + 2 operations */
__device__ __forceinline__ void bit_run(uint32_t *map, uint32_t b0, uint32_t len, int clear)
{
    uint32_t b = b0, end = b0 + len;
    while (b < end) {
        const uint32_t w = b >> 5, lo = b & 31u, take = min(32u - lo, end - b);
        const uint32_t m = (take == 32u ? 0xFFFFFFFFu : ((1u << take) - 1u)) << lo;
        if (clear) map[w] = 0u; else atomicOr(&map[w], m);      /* (result unused: a fire-and-forget RED) */
        b += take;
    }
}

/* the cube of coarse cells every halo can reach after n_balls steps of the ball schedule (same rule as
 * k_mark_mask), once per step and one thread per halo: the marking kernels then start from one 16-byte load per
 * halo instead of a chain of five loads and a page of double-precision arithmetic per group of lanes (ncu: 60 % of
 * their stall samples).  .w = nx | ny << 10 | nz << 20 (each <= 512). */
__global__ void __launch_bounds__(256) k_halo_cubes(GridDev g, const float *__restrict__ centers, const float *__restrict__ rgtp,
                                                    int nh, int n_balls, int4 *__restrict__ cubes,
                                                    const unsigned char *__restrict__ owner,
                                                    const unsigned char *__restrict__ bin_owner)
{
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= nh) return;
    const float root = so_root_period(g.L[0], g.L[1], g.L[2]);
    float ball = rgtp[h];
    for (int k = 0; k < n_balls && (double)ball < 0.25 * (double)root; ++k) ball = so_next_ball(ball);
    const float ball2 = __fmul_rn(ball, ball);
    const double b = fmax(sqrt((double)ball2) * (1.0 + 1.0e-6), g.mask_rmin);
    int x0, nx, y0, ny, z0, nz;
    mask_range(g, 0, centers[3 * h + 0], b, x0, nx);
    mask_range(g, 1, centers[3 * h + 1], b, y0, ny);
    mask_range(g, 2, centers[3 * h + 2], b, z0, nz);
    /* bit 30: the cube reaches into a bin of another owner ("crossing" halo, see k_mark_table); bin_owner == NULL
     * (a single rank) never sets it */
    int crossing = 0;
    if (bin_owner) {
        const int sh = g.mb - DOM_ASSIGN_LOG;
        if (sh < 0) crossing = 1;
        else {
            const int own = owner[h], m = (1 << DOM_ASSIGN_LOG) - 1;
            for (int bz = z0 >> sh; bz <= (z0 + nz - 1) >> sh; ++bz)
                for (int by = y0 >> sh; by <= (y0 + ny - 1) >> sh; ++by)
                    for (int bx = x0 >> sh; bx <= (x0 + nx - 1) >> sh; ++bx)
                        crossing |= (bin_owner[assign_bin_index((uint32_t)(bx & m), (uint32_t)(by & m), (uint32_t)(bz & m))] != own);
        }
    }
    cubes[h] = make_int4(x0, y0, z0, nx | (ny << 10) | (nz << 20) | (crossing << 30));
}

#define MARK_LANES 8
/* A group of lanes per halo: the cube it can reach after n_balls steps of the ball schedule (same rule as k_mark_mask).
 * Sets the owner's bit in the destination table, the "somebody wants this cell" bitmap, its 64^3 pre-filter and
 * (own halos) this rank's focus mask.  A lane takes a whole x-row of the cube: consecutive cells are consecutive
 * bits / table entries, so a row is a handful of word-wide reductions with nothing to wait for.
 * clear = 1 undoes exactly those marks: the tables are kept clean between steps by un-marking the previous
 * catalog (a few million stores) instead of clearing hundreds of megabytes. */
__global__ void __launch_bounds__(256) k_mark_table(GridDev g, const int4 *__restrict__ cubes, int nh,
                                                    const unsigned char *__restrict__ owner, int me,
                                                    uint32_t *__restrict__ table32, uint32_t *__restrict__ any,
                                                    uint32_t *__restrict__ super, uint32_t *__restrict__ mymask,
                                                    uint32_t *__restrict__ plain, uint32_t *__restrict__ listed, int clear)
{
    /* Several ranks.  Ownership follows the bins of k_assign_*: a halo whose cube stays inside bins of its own
     * owner ("plain", the large majority) only sets a bit per cell in `plain` — whoever finds a particle in such a
     * cell knows the destination from the cell's bin.  Only the halos that reach across an ownership boundary
     * ("crossing") enter their owner in the 16-bit table and set `listed`.  A cell's destinations are then
     *     (plain ? {owner of its bin} : {}) | (listed ? table : {})            (k_route_split)
     * which is exact: every plain halo over the cell is owned by the owner of the cell's bin.  The table — 256 MB
     * at 512^3, random read-modify-writes in DRAM, 0.13 ms per step when every halo went through it — is touched
     * by the few crossing halos only. */
    /* clear = 1: only the destination table is un-marked (the bitmaps are small enough to be cleared by a
     * memset every step; the table — 2 bytes per coarse cell — is not).  A NULL array is skipped: a single rank
     * needs neither the table nor a second bitmap. */
    /* MARK_LANES lanes per halo: the typical cube has 9-25 rows, and with a whole warp per halo the kernel was
     * bound by the number of halos in flight (100 000 halos: 90 us) */
    const int lane = threadIdx.x & (MARK_LANES - 1);
    const int wid = (blockIdx.x * blockDim.x + threadIdx.x) / MARK_LANES, nw = (gridDim.x * blockDim.x) / MARK_LANES;
    const int nm = 1 << g.mb, nm1 = nm - 1;
    const SuperGeom sg = super_geom(g.mb);
    for (int h = wid; h < nh; h += nw) {
        const int4 cube = __ldg(cubes + h);
        const int own = owner[h];
        const int x0 = cube.x, y0 = cube.y, z0 = cube.z;
        const int nx = cube.w & 1023, ny = (cube.w >> 10) & 1023, nz = (cube.w >> 20) & 1023;
        const bool crossing = (cube.w >> 30) & 1;
        if (clear && !crossing) continue;
        const uint32_t ob = (1u << own) * 0x00010001u;             /* the owner's bit in both halves of a table word */
        const int xa = x0 & nm1;                                   /* the row's x-range, split where it wraps */
        const int n0 = min(nx, nm - xa), n1 = nx - n0;
        if (!clear) {
            /* the pre-filter once per halo (not per row): the cube covers a handful of its blocks */
            const int sx0 = x0 >> sg.sx, sx1 = (x0 + nx - 1) >> sg.sx, sy0 = y0 >> sg.sy, sy1 = (y0 + ny - 1) >> sg.sy;
            const int sz0 = z0 >> sg.sz, sz1 = (z0 + nz - 1) >> sg.sz, snx = sx1 - sx0 + 1, sny = sy1 - sy0 + 1, snz = sz1 - sz0 + 1;
            for (int i = lane; i < snx * sny * snz; i += MARK_LANES) {
                const uint32_t cx = (uint32_t)((sx0 + i % snx) & ((1 << sg.bx) - 1)), cy = (uint32_t)((sy0 + (i / snx) % sny) & ((1 << sg.by) - 1)),
                               cz = (uint32_t)((sz0 + i / (snx * sny)) & ((1 << sg.bz) - 1));
                const uint32_t sbit = (cz << (sg.bx + sg.by)) | (cy << sg.bx) | cx;
                atomicOr(&super[sbit >> 5], 1u << (sbit & 31));
            }
        }
        for (int r = lane; r < ny * nz; r += MARK_LANES) {
            const uint32_t cy = (uint32_t)((y0 + r % ny) & nm1), cz = (uint32_t)((z0 + r / ny) & nm1);
            const uint32_t row = (cz << (2 * g.mb)) | (cy << g.mb);
            for (int piece = 0; piece < 2; ++piece) {
                const int px = piece ? 0 : xa, pn = piece ? n1 : n0;
                if (pn <= 0) continue;
                if (!clear) {
                    if (any) bit_run(any, row + (uint32_t)px, (uint32_t)pn, 0);
                    if (own == me) bit_run(mymask, row + (uint32_t)px, (uint32_t)pn, 0);
                }
                if (!table32) continue;
                if (!clear) bit_run(crossing ? listed : plain, row + (uint32_t)px, (uint32_t)pn, 0);
                if (!crossing) continue;
                /* table: 16-bit entries, two per word */
                uint32_t c = row + (uint32_t)px, e = c + (uint32_t)pn;
                while (c < e) {
                    const uint32_t w = c >> 1;
                    uint32_t m = ob;
                    if (c & 1u) m &= 0xFFFF0000u;
                    if (((c | 1u) + 1u) > e) m &= 0x0000FFFFu;      /* the word's upper entry lies beyond the run */
                    if (clear) table32[w] = 0u; else atomicOr(&table32[w], m);
                    c = (c | 1u) + 1u;
                }
            }
        }
    }
}

/* ---- routing: one pass over the slice ------------------------------------------------------------------ */
struct StageArgs {
    GridDev g;
    const float4 *slice;
    int64_t n;
    uint32_t index_base;
    const uint32_t *any, *super;
    float4 *dst;                     /* one rank: its receive buffer; several: the hit list k_route_split works on */
    unsigned long long *cursor;      /* records appended so far */
    unsigned long long cap;
    uint32_t flag_bit;
    uint32_t *flags;
    int prefetch;                    /* ask the L2 for the warp's next run while this one is being looked up */
};

#define RW_CAP 128           /* records a WARP collects in shared memory before it writes them out */
#define RT_THREADS 1024      /* one CTA per SM: all its warps share one copy of the pre-filter */
#define RT_U 8               /* independent 16-byte loads in flight per thread */

/* coarse (mask-grid) coordinate of a particle: ((int)floorf(t) & (nc - 1)) >> ms with t = (x - g0) * invh, the grid
 * build's cell_coord, WITHOUT the conversion instruction (F2I runs on the quarter-rate pipe; three per particle were
 * 17 % of this kernel's time).  Adding magic = 1.5 * 2^(23 + ms) with rounding towards -inf leaves
 * floor(t / 2^ms) in the low mantissa bits (the sum's ulp is 2^ms and it is rounded once, downwards), so the
 * coordinate is one add and one AND.  Exact for |t| < 2^(22 + ms), i.e. for particles within 2048 box lengths of
 * the box; beyond that neither this nor the conversion names a meaningful cell. */
__device__ __forceinline__ uint32_t coarse_coord(float x, float g0, float invh, float magic, uint32_t nm1)
{
    const float t = __fmul_rn(__fsub_rn(x, g0), invh);
    return __float_as_uint(__fadd_rd(t, magic)) & nm1;
}
__device__ __forceinline__ float coarse_magic(int ms) { return __uint_as_float(((uint32_t)(127 + 23 + ms) << 23) | 0x400000u); }

/* Every warp works on its own: it streams runs of 32 x RT_U particles, collects the few records they yield in a
 * private shared-memory buffer and writes them out RW_CAP at a time (one reservation per flush).  No block-wide
 * barrier after the start: the first two versions reserved / flushed per CTA and round and spent their time
 * waiting — 6.5 warps stalled on loads and 4 at the barrier per instruction issued, 3 TB/s
 * (profiles/r2_ncu_route_bucket.md).
 *
 * The third version was bound by instruction issue (93 warp instructions per 32 particles, issue slots 65 % busy
 * at 4.1 TB/s): a vote + compaction per 32 particles and destination, bounds checks and conversions per particle.
 * This one keeps a run's hits in registers and compacts them ONCE per run (a warp scan of the per-lane counts),
 * has a check-free body for full runs, and computes cell coordinates on the full-rate pipe (coarse_coord).
 *
 * The kernel only answers "does ANY rank need this particle" (pre-filter in shared memory, then the bitmap in
 * L2).  With several ranks the hits (5-6 % of the slice) go to a list that k_route_split distributes: looking the
 * destination set up here — a dependent 16-bit gather from a table far larger than the L2, plus a tag and a
 * per-destination sort-out at every flush — halved the speed of the whole pass (2.2 against 4.6 TB/s). */
__global__ void __launch_bounds__(RT_THREADS, 1) k_route_stage(const __grid_constant__ StageArgs a)
{
    extern __shared__ __align__(16) uint32_t s_dyn[];
    const int mb = a.g.mb, ms = a.g.ms;
    const SuperGeom sg = super_geom(mb);
    const uint32_t sbits = 1u << (sg.bx + sg.by + sg.bz), swords = (sbits / 32u + 4u) & ~3u;   /* (multiple of 4: what follows stays 16-byte aligned) */
    uint32_t *s_super = s_dyn;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float4 *wbuf = reinterpret_cast<float4 *>(s_dyn + swords) + (size_t)w * RW_CAP;
    for (uint32_t i = threadIdx.x; i < swords; i += blockDim.x) s_super[i] = i < sbits / 32u + 1u ? __ldg(a.super + i) : 0u;
    __syncthreads();
    const float g0x = a.g.g0[0], g0y = a.g.g0[1], g0z = a.g.g0[2], ihx = a.g.invh[0], ihy = a.g.invh[1], ihz = a.g.invh[2];
    const float magic = coarse_magic(ms);
    const uint32_t nm1 = (1u << mb) - 1u;
    const uint32_t S1 = 1u << mb, S2 = 1u << (2 * mb), T1 = 1u << sg.bx, T2 = 1u << (sg.bx + sg.by);
    const uint32_t n = (uint32_t)a.n;
    uint32_t wcnt = 0;                                     /* records in this warp's buffer (same value on every lane) */

    auto flush = [&]() {
        __syncwarp();
        unsigned long long base = 0ull;
        if (lane == 0) {
            base = atomicAdd(a.cursor, (unsigned long long)wcnt);
            if (base + wcnt > a.cap) { atomicOr(a.flags, a.flag_bit); base = ~0ull; }
        }
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (base != ~0ull)
            for (uint32_t i = lane; i < wcnt; i += 32) a.dst[base + i] = wbuf[i];
        __syncwarp();
        wcnt = 0;
    };

    /* the records of the run's particles [U0, U1) of every lane, appended in one go: per-lane counts, a warp scan,
     * every lane stores its own.  Returns false (nothing stored) if they cannot fit even into an empty buffer. */
    auto append = [&](const float4 (&q)[RT_U], const uint32_t (&set)[RT_U], uint32_t i0, auto u0c, auto u1c) -> bool {
        constexpr int U0 = decltype(u0c)::value, U1 = decltype(u1c)::value;
        uint32_t cnt = 0;
#pragma unroll
        for (int u = U0; u < U1; ++u) cnt += set[u];
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += v;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (!total) return true;
        if (total > RW_CAP) return false;
        if (wcnt + total > RW_CAP) flush();
        uint32_t pos = wcnt + incl - cnt;
#pragma unroll
        for (int u = U0; u < U1; ++u) {
            if (set[u]) {
                wbuf[pos] = make_float4(q[u].x, q[u].y, q[u].z, __uint_as_float(a.index_base + i0 + u * 32u));
                ++pos;
            }
        }
        wcnt += total;
        return true;
    };

    auto body = [&](uint32_t r, auto fullc) {
        constexpr bool FULL = decltype(fullc)::value;
        const uint32_t i0 = r * (32u * RT_U) + lane;
        float4 q[RT_U];
#pragma unroll
        for (int u = 0; u < RT_U; ++u) {
            const uint32_t i = i0 + u * 32u;
            q[u] = ld_stream(a.slice + (FULL || i < n ? i : n - 1u));
        }
        /* the lookups of the RT_U particles are issued level by level, so that a level's loads are all in flight
         * together: pre-filter (shared memory), then the "somebody wants it" bitmap (L2) */
        uint32_t bit[RT_U], set[RT_U];
#pragma unroll
        for (int u = 0; u < RT_U; ++u) {
            /* (a particle is routed by the coarse cell of the cell the grid build will sort it into) */
            const uint32_t cx = coarse_coord(q[u].x, g0x, ihx, magic, nm1);
            const uint32_t cy = coarse_coord(q[u].y, g0y, ihy, magic, nm1);
            const uint32_t cz = coarse_coord(q[u].z, g0z, ihz, magic, nm1);
            const uint32_t sbit = (cz >> sg.sz) * T2 + (cy >> sg.sy) * T1 + (cx >> sg.sx);
            bit[u] = cz * S2 + cy * S1 + cx;
            set[u] = (s_super[sbit >> 5] >> (sbit & 31)) & 1u;
            if (!FULL && i0 + u * 32u >= n) set[u] = 0u;
        }
        uint32_t anyw[RT_U];
#pragma unroll
        for (int u = 0; u < RT_U; ++u) anyw[u] = set[u] ? __ldg(a.any + (bit[u] >> 5)) : 0u;
#pragma unroll
        for (int u = 0; u < RT_U; ++u) set[u] = (anyw[u] >> (bit[u] & 31)) & 1u;
        using std::integral_constant;
        if (append(q, set, i0, integral_constant<int, 0>(), integral_constant<int, RT_U>())) return;
        /* more than RW_CAP hits in one run (the core of a cluster): two halves, each of which always fits */
        append(q, set, i0, integral_constant<int, 0>(), integral_constant<int, RT_U / 2>());
        append(q, set, i0, integral_constant<int, RT_U / 2>(), integral_constant<int, RT_U>());
    };

    const uint32_t gw = blockIdx.x * (RT_THREADS / 32) + (uint32_t)w, nwarp = gridDim.x * (RT_THREADS / 32);
    const uint32_t run = 32u * RT_U;
    const uint32_t nfull = n / run, nrun = (n + run - 1) / run;
    uint32_t r = gw;
    for (; r < nfull; r += nwarp) {
        if (a.prefetch && r + nwarp < nfull) {
            const float4 *nx = a.slice + (size_t)(r + nwarp) * run + lane;
#pragma unroll
            for (int u = 0; u < RT_U; u += 2)            /* a lane's 16 bytes: two lanes per 32-byte sector, 128-byte lines */
                asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + u * 32));
        }
        body(r, std::true_type());
    }
    if (r < nrun) body(r, std::false_type());
    if (wcnt) flush();
}

/* ---- several ranks: the hit list -> one run per destination ------------------------------------------ */
struct SplitArgs {
    GridDev g;
    const float4 *hits;
    const unsigned long long *n_hits;
    unsigned long long hits_cap;
    const unsigned short *table;
    const uint32_t *plain, *listed;          /* bit per coarse cell (k_mark_table) */
    const unsigned char *bin_owner;
    int R;
    float4 *dst[ROUTE_MAXR];                 /* staging area per destination; own rank: its receive buffer */
    unsigned long long *cursor[ROUTE_MAXR];  /* records appended so far                                    */
    unsigned long long cap[ROUTE_MAXR];
    uint32_t flag_bit[ROUTE_MAXR];
    uint32_t *flags;
};

#define SP_NT 256
#define SP_U 8
#define SP_TILE (SP_NT * SP_U)
/* Tiles of SP_TILE hits: destination set of every record (owner of the cell's bin if a plain halo covers the
 * cell, table entry if a crossing one does — see k_mark_table; a record's cell is recomputed from its position
 * exactly as the routing pass did), the tile's count per destination in shared memory, ONE reservation per
 * destination and tile; the records are then put in destination order in shared memory and written out as one
 * contiguous run per destination (consecutive threads, consecutive addresses).
 *
 * The destinations are either local staging runs that k_push_copy ships in bulk, or — the default — the peers'
 * receive buffers themselves: reservation (one system-scope atomic on the peer's cursor per tile and destination)
 * and stores go over NVLink from here, 4-16 KB per run, and the transfer overlaps the lookups of the other CTAs
 * instead of following them as a second pass over the same records. */
__global__ void __launch_bounds__(SP_NT) k_route_split(const __grid_constant__ SplitArgs a)
{
    __shared__ float4 s_rec[SP_TILE];
    __shared__ uint32_t s_cnt[ROUTE_MAXR], s_pos[ROUTE_MAXR], s_off[ROUTE_MAXR + 1];
    __shared__ unsigned long long s_base[ROUTE_MAXR];
    unsigned long long n = *a.n_hits;
    if (n > a.hits_cap) n = a.hits_cap;                      /* (the overflow was flagged by k_route_stage) */
    const int mb = a.g.mb;
    const float g0x = a.g.g0[0], g0y = a.g.g0[1], g0z = a.g.g0[2], ihx = a.g.invh[0], ihy = a.g.invh[1], ihz = a.g.invh[2];
    const float magic = coarse_magic(a.g.ms);
    const uint32_t nm1 = (1u << mb) - 1u;
    const int bsh = mb > DOM_ASSIGN_LOG ? mb - DOM_ASSIGN_LOG : 0;
    const unsigned long long tile = (unsigned long long)SP_TILE, ntiles = (n + tile - 1) / tile;
    for (unsigned long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (threadIdx.x < ROUTE_MAXR) { s_cnt[threadIdx.x] = 0u; s_pos[threadIdx.x] = 0u; }
        __syncthreads();
        float4 q[SP_U];
        uint32_t set[SP_U];
#pragma unroll
        for (int u = 0; u < SP_U; ++u) {
            const unsigned long long i = t * tile + (unsigned long long)u * SP_NT + threadIdx.x;
            q[u] = ld_stream(a.hits + (i < n ? i : n - 1ull));
        }
#pragma unroll
        for (int u = 0; u < SP_U; ++u) {
            const unsigned long long i = t * tile + (unsigned long long)u * SP_NT + threadIdx.x;
            const uint32_t cx = coarse_coord(q[u].x, g0x, ihx, magic, nm1);
            const uint32_t cy = coarse_coord(q[u].y, g0y, ihy, magic, nm1);
            const uint32_t cz = coarse_coord(q[u].z, g0z, ihz, magic, nm1);
            const uint32_t bit = (cz << (2 * mb)) | (cy << mb) | cx;
            uint32_t st = 0u;
            if (i < n) {
                const uint32_t pw = __ldg(a.plain + (bit >> 5)), lw = __ldg(a.listed + (bit >> 5));
                if ((pw >> (bit & 31)) & 1u)
                    st = 1u << a.bin_owner[assign_bin_index(cx >> bsh, cy >> bsh, cz >> bsh)];      /* (plain bits only exist with mb >= DOM_ASSIGN_LOG) */
                if ((lw >> (bit & 31)) & 1u) st |= (uint32_t)__ldg(a.table + bit);
            }
            set[u] = st;
        }
#pragma unroll
        for (int u = 0; u < SP_U; ++u)
            for (uint32_t m = set[u]; m; m &= m - 1u) atomicAdd(&s_cnt[__ffs(m) - 1], 1u);
        __syncthreads();
        if (threadIdx.x < a.R) {
            const int d = threadIdx.x;
            const uint32_t c = s_cnt[d];
            unsigned long long base = 0ull;
            if (c) {
                base = atomicAdd_system(a.cursor[d], (unsigned long long)c);
                if (base + c > a.cap[d]) { atomicOr(a.flags, a.flag_bit[d]); base = ~0ull; }
            }
            s_base[d] = base;
        }
        if (threadIdx.x == 32) {                              /* where every destination's run starts in s_rec */
            uint32_t o = 0;
            for (int d = 0; d < a.R; ++d) { s_off[d] = o; o += s_cnt[d]; }
            s_off[a.R] = o;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < SP_U; ++u)
            for (uint32_t m = set[u]; m; m &= m - 1u) {
                const int d = __ffs(m) - 1;
                const uint32_t slot = atomicAdd(&s_pos[d], 1u), p = s_off[d] + slot;
                if (p < (uint32_t)SP_TILE) s_rec[p] = q[u];
                else if (s_base[d] != ~0ull) a.dst[d][s_base[d] + slot] = q[u];      /* (records with several destinations can overfill the tile) */
            }
        __syncthreads();
        const uint32_t tot = min(s_off[a.R], (uint32_t)SP_TILE);
        for (uint32_t p = threadIdx.x; p < tot; p += SP_NT) {
            int d = 0;
            while (d + 1 < a.R && p >= s_off[d + 1]) ++d;
            if (s_base[d] != ~0ull) a.dst[d][s_base[d] + (p - s_off[d])] = s_rec[p];
        }
        __syncthreads();
    }
    __threadfence_system();
}

/* ---- push: staging runs -> the receivers' buffers over NVLink ---------------------------------------- */
struct PushArgs {
    int R, me, parity;
    DomCtrl *peer_ctrl[ROUTE_MAXR];
    float4 *peer_recv[ROUTE_MAXR];
    const float4 *stage;
    unsigned long long stage_cap, recv_cap;
    unsigned long long *counts;             /* [0..R) staged, [R..2R) base at the receiver, [2R..3R) records to copy */
    uint32_t *flags;
};

__global__ void k_push_reserve(const __grid_constant__ PushArgs a)
{
    const int d = threadIdx.x;
    if (d >= a.R || d == a.me) return;
    unsigned long long c = a.counts[d];
    if (c > a.stage_cap) c = a.stage_cap;                     /* (staging overflow was flagged by k_route_stage) */
    unsigned long long base = 0ull;
    if (c) base = atomicAdd_system(&a.peer_ctrl[d]->cursor[a.parity], c);
    if (base + c > a.recv_cap) { atomicOr(a.flags, 4u); c = 0ull; }   /* the receiver sees cursor > cap as well */
    a.counts[a.R + d] = base;
    a.counts[2 * a.R + d] = c;
}

__global__ void __launch_bounds__(256) k_push_copy(const __grid_constant__ PushArgs a)
{
    const int d = blockIdx.y;
    if (d == a.me) return;
    const unsigned long long n = a.counts[2 * a.R + d];
    const float4 *__restrict__ src = a.stage + (size_t)d * a.stage_cap;
    float4 *__restrict__ dst = a.peer_recv[d] + a.counts[a.R + d];
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3ull * stride < n; i += 4ull * stride) {
        const float4 v0 = ld_stream(src + i), v1 = ld_stream(src + i + stride), v2 = ld_stream(src + i + 2ull * stride),
                     v3 = ld_stream(src + i + 3ull * stride);
        dst[i] = v0; dst[i + stride] = v1; dst[i + 2ull * stride] = v2; dst[i + 3ull * stride] = v3;
    }
    for (; i < n; i += stride) dst[i] = ld_stream(src + i);
    __threadfence_system();
}

/* every rank announces `epoch` to every peer, then waits until every peer has announced it to this rank:
 * all pushes of this step are then complete and visible.  Bounded wait: a missing peer raises flag bit 3
 * instead of hanging the device. */
__global__ void k_dom_barrier(const __grid_constant__ PushArgs a, DomCtrl *mine, uint32_t epoch, long long timeout_cycles)
{
    const int d = threadIdx.x;
    if (d >= a.R || d == a.me) return;
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(&a.peer_ctrl[d]->flag[a.me]), "r"(epoch) : "memory");
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(&mine->flag[d]) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (clock64() - t0 > timeout_cycles) { atomicOr(a.flags, 8u); break; }
        __nanosleep(200);
    }
}

/* received count of this step: clamped copy for the grid build (32 bit) and for the host (64 bit) */
__global__ void k_dom_nrecv(const DomCtrl *mine, int parity, unsigned long long recv_cap, uint32_t *n32,
                            unsigned long long *n64, uint32_t *flags)
{
    if (threadIdx.x || blockIdx.x) return;
    unsigned long long c = mine->cursor[parity];
    *n64 = c;
    if (c > recv_cap) { atomicOr(flags, 2u); c = 0ull; }      /* some run was dropped: nothing of this step is usable */
    *n32 = (uint32_t)c;
}

/* ============================================================================================
 * host side
 * ============================================================================================ */
static void dom_geometry(sogpu *h, GridDev &g)
{
    DomainState *D = h->dom;
    for (int k = 0; k < 3; ++k) { h->period[k] = D->cfg.period[k]; h->center[k] = D->cfg.center[k]; }
    domain_geometry(h, D->cfg.n_total, g);
}

extern "C" int sogpu_domain_open(sogpu_t *h, const sogpu_domain_cfg_t *cfg, void *handles192)
{
    if (!h || !cfg || !handles192) return set_err(SOGPU_ERR_ARG, "sogpu_domain_open: NULL argument");
    if (cfg->n_ranks < 1 || cfg->n_ranks > ROUTE_MAXR || cfg->rank < 0 || cfg->rank >= cfg->n_ranks)
        return set_err(SOGPU_ERR_ARG, "sogpu_domain_open: rank %d of %d (at most %d ranks)", cfg->rank, cfg->n_ranks, ROUTE_MAXR);
    if (cfg->n_total <= 0 || cfg->n_total > 0x7FFFFFF0LL || cfg->recv_cap <= 0 || cfg->recv_cap > 0x7FFFFFF0LL ||
        !(cfg->mass > 0.0f) || (cfg->n_ranks > 1 && cfg->stage_cap <= 0))
        return set_err(SOGPU_ERR_ARG, "sogpu_domain_open: bad sizes");
    if (h->dom) return set_err(SOGPU_ERR_ARG, "sogpu_domain_open: already open (sogpu_domain_close first)");
    CU(cudaSetDevice(h->device));
    DomainState *D = new (std::nothrow) DomainState();
    if (!D) return set_err(SOGPU_ERR_NOMEM, "out of host memory");
    memset(D, 0, sizeof(*D));
    D->cfg = *cfg;
    h->dom = D;
    const int R = cfg->n_ranks;
    unsigned char *hb = (unsigned char *)handles192;
    int rc = SOGPU_OK;
    /* receive buffers: two (step parity) when there are peers that can run ahead, one otherwise */
    rc = sogpu_peer_alloc(h, (size_t)cfg->recv_cap * sizeof(float4), (void **)&D->recv[0], hb);
    if (!rc) {
        if (R > 1) rc = sogpu_peer_alloc(h, (size_t)cfg->recv_cap * sizeof(float4), (void **)&D->recv[1], hb + 64);
        else { D->recv[1] = D->recv[0]; memcpy(hb + 64, hb, 64); }
    }
    if (!rc) rc = sogpu_peer_alloc(h, sizeof(DomCtrl), (void **)&D->ctrl, hb + 128);
    if (rc) return rc;
    CU(cudaMemsetAsync(D->ctrl, 0, sizeof(DomCtrl), h->stream));
    if (R > 1) CU(cudaMalloc(&D->stage, (size_t)R * (size_t)cfg->stage_cap * sizeof(float4)));
    if (R > 1) CU(cudaMalloc(&D->hits, (size_t)cfg->recv_cap * sizeof(float4)));
    CU(cudaMalloc(&D->d_counts, (3 * ROUTE_MAXR + 2) * sizeof(unsigned long long)));
    CU(cudaMalloc(&D->d_flags, sizeof(uint32_t)));
    CU(cudaMalloc(&D->d_nrecv32, sizeof(uint32_t)));
    CU(cudaMalloc(&D->d_bins, (DOM_ASSIGN_BINS + 1) * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(D->d_flags, 0, sizeof(uint32_t), h->stream));
    CU(cudaMemsetAsync(D->d_counts, 0, (3 * ROUTE_MAXR + 2) * sizeof(unsigned long long), h->stream));
    {   /* destination table + bitmaps at the mask resolution of this snapshot; cleared once, then kept clean */
        for (int k = 0; k < 3; ++k) { h->period[k] = cfg->period[k]; h->center[k] = cfg->center[k]; }
        GridDev g;
        domain_geometry(h, cfg->n_total, g);
        const size_t n_cells = (size_t)1 << (3 * g.mb), words = n_cells / 32 + 1;
        const size_t swords = super_words(g.mb);
        CU(cudaMalloc(&D->d_table, (n_cells + 2) * sizeof(unsigned short)));
        CU(cudaMalloc(&D->d_any, words * sizeof(uint32_t)));
        CU(cudaMalloc(&D->d_bin_owner, DOM_ASSIGN_BINS));
        if (R > 1) {
            CU(cudaMalloc(&D->d_plain, words * sizeof(uint32_t)));
            CU(cudaMalloc(&D->d_listed, words * sizeof(uint32_t)));
        }
        CU(cudaMalloc(&D->d_mymask, words * sizeof(uint32_t)));
        CU(cudaMalloc(&D->d_super, swords * sizeof(uint32_t)));
        CU(cudaMemsetAsync(D->d_table, 0, (n_cells + 2) * sizeof(unsigned short), h->stream));
        CU(cudaMemsetAsync(D->d_any, 0, words * sizeof(uint32_t), h->stream));
        CU(cudaMemsetAsync(D->d_mymask, 0, words * sizeof(uint32_t), h->stream));
        CU(cudaMemsetAsync(D->d_super, 0, swords * sizeof(uint32_t), h->stream));
    }
    CU(cudaStreamSynchronize(h->stream));
    for (int d = 0; d < R; ++d) { D->peer_recv[0][d] = D->peer_recv[1][d] = nullptr; D->peer_ctrl[d] = nullptr; }
    D->peer_recv[0][cfg->rank] = D->recv[0];
    D->peer_recv[1][cfg->rank] = D->recv[1];
    D->peer_ctrl[cfg->rank] = D->ctrl;
    D->connected = (R == 1);
    D->n_hint = std::max<int64_t>(cfg->recv_cap / 2, 1);
    D->direct = true;                   /* k_route_split stores into the peers' buffers itself (SOGPU_DIRECT_PUSH=0: staging + k_push_copy) */
    if (const char *e = getenv("SOGPU_DIRECT_PUSH")) D->direct = atoi(e) != 0;
    D->route_prefetch = 1;              /* measured at 1024^3, one GPU: 7.68 -> 7.33 ms per step */
    if (const char *e = getenv("SOGPU_ROUTE_PREFETCH")) D->route_prefetch = atoi(e);
    return SOGPU_OK;
}

/* pointers (valid in THIS process: sogpu_peer_open, or plain device pointers of a one-process run with peer
 * access enabled) to every rank's two receive buffers and control block; entries of the own rank are ignored */
extern "C" int sogpu_domain_connect(sogpu_t *h, void *const *recv0, void *const *recv1, void *const *ctrl)
{
    if (!h || !h->dom || !recv0 || !recv1 || !ctrl) return set_err(SOGPU_ERR_ARG, "sogpu_domain_connect: bad argument");
    DomainState *D = h->dom;
    for (int d = 0; d < D->cfg.n_ranks; ++d) {
        if (d == D->cfg.rank) continue;
        if (!recv0[d] || !recv1[d] || !ctrl[d]) return set_err(SOGPU_ERR_ARG, "sogpu_domain_connect: NULL pointer for rank %d", d);
        D->peer_recv[0][d] = (float4 *)recv0[d];
        D->peer_recv[1][d] = (float4 *)recv1[d];
        D->peer_ctrl[d] = (DomCtrl *)ctrl[d];
    }
    D->connected = true;
    return SOGPU_OK;
}

/* this rank's own buffers as plain device pointers (one-process runs: hand them to the other handles) */
extern "C" int sogpu_domain_pointers(sogpu_t *h, void **recv0, void **recv1, void **ctrl)
{
    if (!h || !h->dom || !recv0 || !recv1 || !ctrl) return set_err(SOGPU_ERR_ARG, "sogpu_domain_pointers: bad argument");
    *recv0 = h->dom->recv[0]; *recv1 = h->dom->recv[1]; *ctrl = h->dom->ctrl;
    return SOGPU_OK;
}

extern "C" int sogpu_enable_peer_access(sogpu_t *h, int peer_device)
{
    if (!h) return set_err(SOGPU_ERR_ARG, "NULL handle");
    CU(cudaSetDevice(h->device));
    if (peer_device == h->device) return SOGPU_OK;
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return SOGPU_OK; }
    if (e != cudaSuccess) return set_err(SOGPU_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d): %s", peer_device, cudaGetErrorString(e));
    return SOGPU_OK;
}

extern "C" int sogpu_domain_close(sogpu_t *h)
{
    if (!h || !h->dom) return SOGPU_OK;
    DomainState *D = h->dom;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    cudaFree(D->recv[0]);
    if (D->recv[1] != D->recv[0]) cudaFree(D->recv[1]);
    cudaFree(D->ctrl); cudaFree(D->stage); cudaFree(D->hits); cudaFree(D->d_counts); cudaFree(D->d_flags); cudaFree(D->d_nrecv32);
    cudaFree(D->d_owner); cudaFree(D->d_cubes); cudaFree(D->d_bins);
    cudaFree(D->d_table); cudaFree(D->d_any); cudaFree(D->d_mymask); cudaFree(D->d_super);
    cudaFree(D->d_plain); cudaFree(D->d_listed); cudaFree(D->d_bin_owner);
    delete D;
    h->mask_ready = nullptr;
    h->dom = nullptr;
    h->d_n_dev = nullptr;
    return SOGPU_OK;
}

/* start of a step: ownership, destination table, this rank's mask.  d_centers / d_rgtp: the WHOLE catalog (device). */
extern "C" int sogpu_domain_begin(sogpu_t *h, const void *d_centers, const void *d_rgtp, int32_t nh, int32_t n_balls)
{
    if (!h || !h->dom || !d_centers || !d_rgtp || nh <= 0 || n_balls < 1)
        return set_err(SOGPU_ERR_ARG, "sogpu_domain_begin: bad argument");
    DomainState *D = h->dom;
    if (!D->connected) return set_err(SOGPU_ERR_ARG, "sogpu_domain_begin: call sogpu_domain_connect first");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int R = D->cfg.n_ranks, me = D->cfg.rank;
    int rc = ensure_query(h, nh);
    if (rc) return rc;
    GridDev g;
    dom_geometry(h, g);
    if (D->marked && R > 1) {      /* un-mark the previous catalog (its cubes are still in d_cubes) in the destination table */
        ProfScope p(h, KID_MARK_MASK);
        k_mark_table<<<std::min((D->marked_nh * MARK_LANES + 255) / 256, h->sm_count * 8), 256, 0, s>>>(
            g, D->d_cubes, D->marked_nh, D->d_owner, me, (uint32_t *)D->d_table, D->d_any, D->d_super, D->d_mymask,
            D->d_plain, D->d_listed, 1);
    }
    D->marked = false;
    if (nh > D->owner_cap) {
        CU(cudaStreamSynchronize(s));        /* (the un-marking above reads the buffers that are replaced here) */
        cudaFree(D->d_owner); D->d_owner = nullptr; D->owner_cap = 0;
        cudaFree(D->d_cubes); D->d_cubes = nullptr;
        CU(cudaMalloc(&D->d_cubes, (size_t)std::max(nh, 1024) * sizeof(int4)));
        CU(cudaMalloc(&D->d_owner, (size_t)std::max(nh, 1024)));
        D->owner_cap = std::max(nh, 1024);
    }
    {   /* the bitmaps: 2 x 2^(3 mb) bits + the pre-filter — a few to 32 MB, microseconds */
        const size_t words = ((size_t)1 << (3 * g.mb)) / 32 + 1, swords = super_words(g.mb);
        if (R > 1) {
            CU(cudaMemsetAsync(D->d_any, 0, words * sizeof(uint32_t), s));
            CU(cudaMemsetAsync(D->d_plain, 0, words * sizeof(uint32_t), s));
            CU(cudaMemsetAsync(D->d_listed, 0, words * sizeof(uint32_t), s));
        }
        CU(cudaMemsetAsync(D->d_mymask, 0, words * sizeof(uint32_t), s));
        CU(cudaMemsetAsync(D->d_super, 0, swords * sizeof(uint32_t), s));
    }
    if (d_centers != h->d_centers)
        CU(cudaMemcpyAsync(h->d_centers, d_centers, (size_t)nh * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    if (d_rgtp != h->d_rgtp)
        CU(cudaMemcpyAsync(h->d_rgtp, d_rgtp, (size_t)nh * sizeof(float), cudaMemcpyDeviceToDevice, s));
    D->epoch += 1u;
    D->parity = (int)(D->epoch & 1u);
    D->n_balls = n_balls;
    D->nh = nh;
    /* the OTHER parity's cursor was last used one step ago, by reservations that all happened before that
     * step's barrier: clearing it here, ahead of this step's barrier, is safe (see the header comment) */
    CU(cudaMemsetAsync(&D->ctrl->cursor[D->parity ^ 1], 0, sizeof(unsigned long long), s));
    if (R == 1) CU(cudaMemsetAsync(&D->ctrl->cursor[D->parity], 0, sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(D->d_counts, 0, (3 * ROUTE_MAXR + 2) * sizeof(unsigned long long), s));
    CU(cudaMemsetAsync(D->d_flags, 0, sizeof(uint32_t), s));
    h->stats.last_kernel_launches = 0;
    {
        ProfScope p(h, KID_ASSIGN, 0.0, R > 1 ? 4 : 0);
        if (R > 1) {
            CU(cudaMemsetAsync(D->d_bins, 0, (DOM_ASSIGN_BINS + 1) * sizeof(unsigned long long), s));
            double vol = (double)D->cfg.period[0] * D->cfg.period[1] * D->cfg.period[2];
            k_assign_hist<<<(nh + 255) / 256, 256, 0, s>>>(g, h->d_centers, h->d_rgtp, nh, (double)D->cfg.n_total / vol, D->d_bins);
            k_assign_scan<<<1, 1024, 0, s>>>(D->d_bins);
            k_assign_bin_owner<<<DOM_ASSIGN_BINS / 256, 256, 0, s>>>(R, D->d_bins, D->d_bin_owner);
            k_assign_owner<<<(nh + 255) / 256, 256, 0, s>>>(g, h->d_centers, nh, D->d_bin_owner, D->d_owner);
        } else {
            CU(cudaMemsetAsync(D->d_owner, 0, (size_t)nh, s));
        }
    }
    {
        ProfScope p(h, KID_MARK_MASK, 0.0, 2);
        k_halo_cubes<<<(nh + 255) / 256, 256, 0, s>>>(g, h->d_centers, h->d_rgtp, nh, n_balls, D->d_cubes, D->d_owner,
                                                        R > 1 ? D->d_bin_owner : nullptr);
        k_mark_table<<<std::min((nh * MARK_LANES + 255) / 256, h->sm_count * 8), 256, 0, s>>>(
            g, D->d_cubes, nh, D->d_owner, me, R > 1 ? (uint32_t *)D->d_table : nullptr,
            R > 1 ? D->d_any : nullptr, D->d_super, D->d_mymask, D->d_plain, D->d_listed, 0);
        D->marked = true; D->marked_balls = n_balls; D->marked_nh = nh;
    }
    CU(cudaGetLastError());
    D->routed = false;
    h->built = false;
    h->have_result = false;
    return SOGPU_OK;
}

static int dom_stage_args(sogpu *h, StageArgs &a, const void *d_chunk, int64_t n, int64_t index_base)
{
    DomainState *D = h->dom;
    if (index_base < 0 || index_base + n > 0x7FFFFFF0LL) return set_err(SOGPU_ERR_ARG, "sogpu_domain_route: index out of range");
    memset(&a, 0, sizeof(a));
    dom_geometry(h, a.g);
    a.slice = (const float4 *)d_chunk; a.n = n; a.index_base = (uint32_t)index_base;
    a.super = D->d_super; a.flags = D->d_flags;
    a.cap = (unsigned long long)D->cfg.recv_cap;
    a.prefetch = D->route_prefetch;
    if (D->cfg.n_ranks > 1) {           /* several ranks: everything somebody needs -> the hit list */
        a.any = D->d_any; a.dst = D->hits; a.cursor = D->d_counts + 3 * ROUTE_MAXR + 1; a.flag_bit = 1u;
    } else {                            /* one rank: straight into its receive buffer */
        a.any = D->d_mymask; a.dst = D->recv[D->parity]; a.cursor = &D->ctrl->cursor[D->parity]; a.flag_bit = 2u;
    }
    return SOGPU_OK;
}

/* route n particles (float4 {x,y,z,m}, device) whose first one has global index index_base; may be called
 * several times per step (chunks of the slice, e.g. as they arrive from the host) */
extern "C" int sogpu_domain_route(sogpu_t *h, const void *d_chunk, int64_t n, int64_t index_base)
{
    if (!h || !h->dom || (n > 0 && !d_chunk) || n < 0) return set_err(SOGPU_ERR_ARG, "sogpu_domain_route: bad argument");
    if (n == 0) return SOGPU_OK;
    CU(cudaSetDevice(h->device));
    StageArgs a;
    int rc = dom_stage_args(h, a, d_chunk, n, index_base);
    if (rc) return rc;
    {
        ProfScope p(h, KID_ROUTE, 16.0 * (double)n);
        /* persistent grid of exactly the CTAs that are resident at once (one per SM): a second, partial wave
         * would leave the SMs of the finished CTAs idle */
        const size_t ssm = super_words(a.g.mb) * sizeof(uint32_t) + (size_t)(RT_THREADS / 32) * RW_CAP * sizeof(float4);
        CU(cudaFuncSetAttribute(k_route_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ssm));   /* (more than 48 KB: opt in) */
        k_route_stage<<<(int)std::min<int64_t>((n + RT_THREADS - 1) / RT_THREADS, (int64_t)h->sm_count), RT_THREADS, ssm, h->stream>>>(a);
    }
    CU(cudaGetLastError());
    return SOGPU_OK;
}

/* same from page-locked HOST memory: n xyz triplets (+ the shared mass) are copied in pieces, unpacked into
 * d_slice_dst (device float4, caller-owned, n entries) and routed piece by piece — the copy of the next piece
 * is queued behind the routing of this one on the same stream */
extern "C" int sogpu_domain_route_host(sogpu_t *h, const float *xyz_pinned, int64_t n, int64_t index_base, void *d_slice_dst)
{
    if (!h || !h->dom || !xyz_pinned || !d_slice_dst || n <= 0) return set_err(SOGPU_ERR_ARG, "sogpu_domain_route_host: bad argument");
    CU(cudaSetDevice(h->device));
    const int64_t rchunk = 1 << 23;   /* 96 MB of xyz per piece */
    if (!h->d_raw) CU(cudaMalloc(&h->d_raw, (size_t)rchunk * 3 * sizeof(float)));
    float4 *dst = (float4 *)d_slice_dst;
    for (int64_t i0 = 0; i0 < n; i0 += rchunk) {
        const int64_t k = std::min(rchunk, n - i0);
        CU(cudaMemcpyAsync(h->d_raw, xyz_pinned + 3 * i0, (size_t)k * 3 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        k_expand_xyz<<<h->sm_count * 8, 256, 0, h->stream>>>(h->d_raw, h->dom->cfg.mass, dst + i0, k);
        int rc = sogpu_domain_route(h, dst + i0, k, index_base + i0);
        if (rc) return rc;
    }
    return SOGPU_OK;
}

/* end of the routing: reservations at the receivers, copies over NVLink, barrier (barrier = 0: the caller
 * makes sure by other means that every rank's push has completed before sogpu_domain_solve is enqueued) */
extern "C" int sogpu_domain_push(sogpu_t *h, int barrier)
{
    if (!h || !h->dom) return set_err(SOGPU_ERR_ARG, "sogpu_domain_push: no open domain");
    DomainState *D = h->dom;
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int R = D->cfg.n_ranks;
    if (R > 1) {
        PushArgs a;
        memset(&a, 0, sizeof(a));
        const bool direct = D->direct;
        a.R = R; a.me = D->cfg.rank; a.parity = D->parity;
        for (int d = 0; d < R; ++d) { a.peer_ctrl[d] = D->peer_ctrl[d]; a.peer_recv[d] = D->peer_recv[D->parity][d]; }
        a.stage = D->stage; a.stage_cap = (unsigned long long)D->cfg.stage_cap; a.recv_cap = (unsigned long long)D->cfg.recv_cap;
        a.counts = D->d_counts; a.flags = D->d_flags;
        {   /* the hits of this rank's slice -> its own receive buffer and one staging run per peer */
            SplitArgs sa;
            memset(&sa, 0, sizeof(sa));
            dom_geometry(h, sa.g);
            sa.hits = D->hits; sa.n_hits = D->d_counts + 3 * ROUTE_MAXR + 1; sa.hits_cap = (unsigned long long)D->cfg.recv_cap;
            sa.table = D->d_table; sa.plain = D->d_plain; sa.listed = D->d_listed; sa.bin_owner = D->d_bin_owner;
            sa.R = R; sa.flags = D->d_flags;
            for (int d = 0; d < R; ++d) {
                if (d == D->cfg.rank) {
                    sa.dst[d] = D->recv[D->parity]; sa.cursor[d] = &D->ctrl->cursor[D->parity];
                    sa.cap[d] = (unsigned long long)D->cfg.recv_cap; sa.flag_bit[d] = 2u;
                } else if (direct) {
                    sa.dst[d] = D->peer_recv[D->parity][d]; sa.cursor[d] = &D->peer_ctrl[d]->cursor[D->parity];
                    sa.cap[d] = (unsigned long long)D->cfg.recv_cap; sa.flag_bit[d] = 4u;
                } else {
                    sa.dst[d] = D->stage + (size_t)d * (size_t)D->cfg.stage_cap; sa.cursor[d] = D->d_counts + d;
                    sa.cap[d] = (unsigned long long)D->cfg.stage_cap; sa.flag_bit[d] = 1u;
                }
            }
            ProfScope p(h, KID_ROUTE_SPLIT, 0.0);
            k_route_split<<<h->sm_count * 4, SP_NT, 0, s>>>(sa);
        }
        if (!direct) {
            ProfScope p(h, KID_PUSH, 0.0, 2);
            k_push_reserve<<<1, 32, 0, s>>>(a);
            k_push_copy<<<dim3((unsigned)std::max(1, h->sm_count * 4 / R), (unsigned)R), 256, 0, s>>>(a);
        }
        if (barrier) {
            ProfScope p(h, KID_BARRIER);
            k_dom_barrier<<<1, 32, 0, s>>>(a, D->ctrl, D->epoch, (long long)20000000000LL);   /* ~10 s */
        }
    }
    CU(cudaGetLastError());
    D->routed = true;
    return SOGPU_OK;
}

/* grid over what arrived + SO solve of the halos this rank owns.  d_out_n / d_out_m (device, nh entries of the
 * WHOLE catalog, may be NULL): N_Delta or code / M_Delta for own halos, CODE_NOT_MINE (0x80808080) elsewhere. */
extern "C" int sogpu_domain_solve(sogpu_t *h, float thr, int32_t nM, void *d_out_n, void *d_out_m)
{
    if (!h || !h->dom) return set_err(SOGPU_ERR_ARG, "sogpu_domain_solve: no open domain");
    DomainState *D = h->dom;
    if (!D->routed) return set_err(SOGPU_ERR_ARG, "sogpu_domain_solve: call sogpu_domain_push first");
    CU(cudaSetDevice(h->device));
    cudaStream_t s = h->stream;
    const int32_t nh = D->nh;
    k_dom_nrecv<<<1, 32, 0, s>>>(D->ctrl, D->parity, (unsigned long long)D->cfg.recv_cap, D->d_nrecv32,
                                 D->d_counts + 3 * ROUTE_MAXR, D->d_flags);
    /* the grid build reads the count on the device */
    h->n = D->cfg.recv_cap;
    h->d_in = D->recv[D->parity];
    h->indexed = true;
    h->indexed_mass = D->cfg.mass;
    h->n_total = D->cfg.n_total;
    h->mass_state = -1;
    h->d_n_dev = D->d_nrecv32;
    h->n_hint = D->n_hint;
    const int keep_launches = h->stats.last_kernel_launches;
    h->mask_ready = D->d_mymask;
    int rc = build_grid_impl(h, -1, D->n_balls);
    h->d_n_dev = nullptr;
    if (rc) return rc;
    const int build_launches = h->stats.last_kernel_launches;
    rc = ensure_query(h, nh);
    if (rc) return rc;
    CU(cudaMemsetAsync(h->d_out_n, 0x80, (size_t)nh * sizeof(int32_t), s));
    CU(cudaMemsetAsync(h->d_out_m, 0x80, (size_t)nh * sizeof(float), s));
    h->q_owner = D->d_owner; h->q_me = D->cfg.rank; h->q_ranks = D->cfg.n_ranks;
    rc = run_query(h, h->d_centers, h->d_rgtp, nh, thr, nM);
    h->q_owner = nullptr;
    if (rc) return rc;
    h->stats.last_kernel_launches += keep_launches + build_launches;
    if (d_out_n) CU(cudaMemcpyAsync(d_out_n, h->d_out_n, (size_t)nh * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (d_out_m) CU(cudaMemcpyAsync(d_out_m, h->d_out_m, (size_t)nh * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return SOGPU_OK;
}

/* synchronises: what this rank received / sent in the last step and the step's error flags
 * (bit0 staging overflow, bit1 receive overflow here, bit2 receive overflow at a peer, bit3 barrier timeout);
 * owner (host, nh bytes, may be NULL) = owner rank of every halo */
extern "C" int sogpu_domain_result(sogpu_t *h, int64_t *n_recv, int64_t *n_sent, uint32_t *flags, unsigned char *owner)
{
    if (!h || !h->dom) return set_err(SOGPU_ERR_ARG, "sogpu_domain_result: no open domain");
    DomainState *D = h->dom;
    CU(cudaSetDevice(h->device));
    unsigned long long c[3 * ROUTE_MAXR + 2];
    uint32_t f = 0;
    CU(cudaMemcpyAsync(c, D->d_counts, sizeof(c), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(&f, D->d_flags, sizeof(f), cudaMemcpyDeviceToHost, h->stream));
    if (owner) CU(cudaMemcpyAsync(owner, D->d_owner, (size_t)D->nh, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    int64_t sent = 0;
    for (int d = 0; d < D->cfg.n_ranks; ++d) sent += (int64_t)c[d];
    if (n_recv) *n_recv = (int64_t)c[3 * ROUTE_MAXR];
    if (n_sent) *n_sent = sent;
    if (flags) *flags = f;
    if (!(f & 2u) && c[3 * ROUTE_MAXR] > 0) D->n_hint = (int64_t)c[3 * ROUTE_MAXR];
    return SOGPU_OK;
}

/* ---- small helpers for a one-process, several-devices host program (so -gpus N) ---------------------------- */

/* the handle's particle array as set by the last sogpu_set_particles_* / ingest (device float4 {x,y,z,m}) */
extern "C" int sogpu_particles_device(sogpu_t *h, void **d_xyzm, int64_t *n)
{
    if (!h || !d_xyzm || !n) return set_err(SOGPU_ERR_ARG, "sogpu_particles_device: NULL argument");
    *d_xyzm = (void *)h->d_in;
    *n = h->n;
    return SOGPU_OK;
}

/* synchronous copies on the handle's device / stream: host <-> device, and from another handle's device */
extern "C" int sogpu_copy(sogpu_t *h, void *dst, const void *src, size_t bytes, int kind)
{
    if (!h || (bytes && (!dst || !src)) || kind < 0 || kind > 2) return set_err(SOGPU_ERR_ARG, "sogpu_copy: bad argument");
    CU(cudaSetDevice(h->device));
    const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDefault;
    if (bytes) CU(cudaMemcpyAsync(dst, src, bytes, k, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return SOGPU_OK;
}

/* make CSR member lists (ascending (r^2, index) per group, e.g. merged from several devices) the "last result" of
 * this handle, so that sogpu_tag_members / sogpu_tag_replay / sogpu_vcm work on them */
extern "C" int sogpu_set_members(sogpu_t *h, const int64_t *offsets, const int32_t *members, const float *d2, int32_t nh)
{
    if (!h || !offsets || nh <= 0 || offsets[0] != 0 || (offsets[nh] > 0 && (!members || !d2)))
        return set_err(SOGPU_ERR_ARG, "sogpu_set_members: bad argument");
    CU(cudaSetDevice(h->device));
    int rc = ensure_query(h, nh);
    if (rc) return rc;
    const unsigned long long tot = (unsigned long long)offsets[nh];
    if (tot > h->member_cap) {
        cudaFree(h->d_members); cudaFree(h->d_md2); cudaFree(h->d_members2); cudaFree(h->d_md2_2);
        h->d_members = nullptr; h->d_md2 = nullptr; h->d_members2 = nullptr; h->d_md2_2 = nullptr;
        h->member_cap = tot + tot / 16 + 1024;
        CU(cudaMalloc(&h->d_members, (size_t)h->member_cap * sizeof(int32_t)));
        CU(cudaMalloc(&h->d_md2, (size_t)h->member_cap * sizeof(float)));
    }
    cudaStream_t s = h->stream;
    CU(cudaMemcpyAsync(h->d_out_off, offsets, ((size_t)nh + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, s));
    if (tot) {
        CU(cudaMemcpyAsync(h->d_members, members, (size_t)tot * sizeof(int32_t), cudaMemcpyHostToDevice, s));
        CU(cudaMemcpyAsync(h->d_md2, d2, (size_t)tot * sizeof(float), cudaMemcpyHostToDevice, s));
    }
    std::vector<int32_t> nd((size_t)nh);
    for (int32_t i = 0; i < nh; ++i) nd[i] = (int32_t)(offsets[i + 1] - offsets[i]);
    CU(cudaMemcpyAsync(h->d_out_n, nd.data(), (size_t)nh * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    unsigned long long u4[4] = {tot, 0ull, 0ull, 0ull};
    CU(cudaMemcpyAsync(h->d_u64, u4, sizeof(u4), cudaMemcpyHostToDevice, s));
    CU(cudaMemsetAsync(h->d_counters, 0, 24 * sizeof(uint32_t), s));
    CU(cudaStreamSynchronize(s));
    h->last_h = nh;
    h->have_result = true;
    h->members_csr = true;
    h->members_sorted = true;
    h->want_d2 = true;
    h->member_overflow = false;
    return SOGPU_OK;
}

