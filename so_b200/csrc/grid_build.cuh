/* grid_build.cuh — kdBuildTree replacement (kd2.c:1096-1185): particles -> periodic cell grid.
 *
 * The reference permutes its 60-byte PINIT records with a recursive quickselect (kdSelectInit,
 * kd2.c:1013-1041) until every kd bucket holds <= 16 particles.  Here the particles are sorted by
 * cell key — (zorder2(iy,iz) << lb) | ix — with an MSD radix scheme whose every global store is a
 * coalesced run, because scattered 16-byte stores are bound by L2 transaction rate, not by HBM:
 *
 *   level l = 0..L-1 :  k_lvl_hist       digit histogram per parent bucket (smem atomics)
 *                       k_scan_*         exclusive scan -> child bucket starts
 *                       k_lvl_partition  tile-wise counting sort by digit in shared memory, one
 *                                        global atomic per (tile, child), run writes
 *   final            :  k_bucket_sort    one CTA per final bucket (~1 K particles, <= 4096 cells):
 *                                        counting sort by cell entirely in shared memory, coalesced
 *                                        copy out, writes the bucket's slice of the cell table
 *
 * Payload is one float4 per particle: {x, y, z, original index (as int bits)}; masses stay in the
 * caller's array (all equal on the fast path; gathered by index on the general path).
 * Algorithmic bytes per particle: level 0: 16 (hist) + 36 (partition, key written unless last);
 * further levels 4 + 36; final 16 + 16; total 124 B for two levels.
 */
#pragma once

#define SCAN_TILE 1024   /* 256 threads x 4 consecutive entries, moved as uint4 */

__global__ void __launch_bounds__(256) k_scan_reduce(const uint32_t *__restrict__ a, int64_t n,
                                                     uint32_t *__restrict__ bsum)
{
    __shared__ uint32_t ws[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * 4;
    uint32_t s = 0;
    if (base + 4 <= n) {
        const uint4 v = *reinterpret_cast<const uint4 *>(a + base);
        s = v.x + v.y + v.z + v.w;
    } else {
        for (int k = 0; k < 4; ++k) if (base + k < n) s += a[base + k];
    }
    s = __reduce_add_sync(0xFFFFFFFFu, s);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int k = 0; k < 8; ++k) t += ws[k];
        bsum[blockIdx.x] = t;
    }
}

/* exclusive scan of bsum[0..nb) by one block */
__global__ void __launch_bounds__(1024) k_scan_bsums(uint32_t *__restrict__ bsum, int64_t nb)
{
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int64_t b0 = 0; b0 < nb; b0 += 1024) {
        int64_t i = b0 + threadIdx.x;
        uint32_t v = (i < nb) ? bsum[i] : 0u, x = v;
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += t;
        }
        if (lane == 31) ws[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t y = ws[lane];
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, y, o);
                if (lane >= o) y += t;
            }
            ws[lane] = y;
        }
        __syncthreads();
        uint32_t incl = x + (w ? ws[w - 1] : 0u) + carry;
        if (i < nb) bsum[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
}

/* exclusive scan of each tile plus its block offset: out[i] = sum a[0..i); out may alias a.
 * If `copy` is given it receives the same values (the atomic cursors of the partition pass). */
__global__ void __launch_bounds__(256) k_scan_apply(const uint32_t *a, int64_t n, const uint32_t *__restrict__ bsum,
                                                    uint32_t *out, uint32_t *copy)
{
    __shared__ uint32_t ws[8];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * 4;
    const bool whole = base + 4 <= n;
    uint32_t v[4];
    if (whole) {
        const uint4 x4 = *reinterpret_cast<const uint4 *>(a + base);
        v[0] = x4.x; v[1] = x4.y; v[2] = x4.z; v[3] = x4.w;
    } else {
        for (int k = 0; k < 4; ++k) v[k] = (base + k < n) ? a[base + k] : 0u;
    }
    const uint32_t s = v[0] + v[1] + v[2] + v[3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t x = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += t;
    }
    if (lane == 31) ws[w] = x;
    __syncthreads();
    uint32_t off = bsum[blockIdx.x];
    for (int k = 0; k < w; ++k) off += ws[k];
    uint4 e;
    e.x = off + x - s; e.y = e.x + v[0]; e.z = e.y + v[1]; e.w = e.z + v[2];
    const uint32_t run = e.w + v[3];
    if (whole) {
        *reinterpret_cast<uint4 *>(out + base) = e;
        if (copy) *reinterpret_cast<uint4 *>(copy + base) = e;
    } else {
        const uint32_t ee[4] = {e.x, e.y, e.z, e.w};
        for (int k = 0; k < 4; ++k)
            if (base + k < n) {
                out[base + k] = ee[k];
                if (copy) copy[base + k] = ee[k];
            }
    }
    if (base <= n - 1 && n - 1 < base + 4) out[n] = run;   /* sentinel: total (= particles kept) */
}

/* the same exclusive scan by ONE block (bucket tables up to a few 10^5 entries: one launch instead of
 * three): out[i] = sum a[0..i), copy[i] likewise, out[n] = total.  out may alias a.  Four consecutive
 * entries per thread, moved as uint4: the loads AND the stores of a warp are 512 contiguous bytes
 * (measured, tools/micro/scan_microbench.cu, 16384 entries: 6.2 us; 16 entries per thread with scalar
 * stores at a 64-byte stride between lanes: 24.6 us — partial-sector writes again). */
#define SCAN1_PER 4
__global__ void __launch_bounds__(1024) k_scan_one(const uint32_t *a, int64_t n, uint32_t *out, uint32_t *copy)
{
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0u;
    __syncthreads();
    for (int64_t c0 = 0; c0 < n; c0 += 1024 * SCAN1_PER) {
        const int64_t base = c0 + (int64_t)t * SCAN1_PER;
        const bool whole = base + SCAN1_PER <= n;             /* (a, out, copy are 16-byte aligned allocations) */
        uint32_t v[SCAN1_PER];
        if (whole) {
            const uint4 x4 = *reinterpret_cast<const uint4 *>(a + base);
            v[0] = x4.x; v[1] = x4.y; v[2] = x4.z; v[3] = x4.w;
        } else {
#pragma unroll
            for (int k = 0; k < SCAN1_PER; ++k) v[k] = (base + k < n) ? a[base + k] : 0u;
        }
        const uint32_t s = v[0] + v[1] + v[2] + v[3];
        uint32_t x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += u;
        }
        if (lane == 31) ws[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t y = ws[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, y, o);
                if (lane >= o) y += u;
            }
            ws[lane] = y;
        }
        __syncthreads();
        uint4 e;
        e.x = carry_s + (w ? ws[w - 1] : 0u) + x - s;
        e.y = e.x + v[0]; e.z = e.y + v[1]; e.w = e.z + v[2];
        const uint32_t run = e.w + v[3];
        if (whole) {
            *reinterpret_cast<uint4 *>(out + base) = e;
            if (copy) *reinterpret_cast<uint4 *>(copy + base) = e;
        } else {
            const uint32_t ee[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
            for (int k = 0; k < SCAN1_PER; ++k)
                if (base + k < n) {
                    out[base + k] = ee[k];
                    if (copy) copy[base + k] = ee[k];
                }
        }
        __syncthreads();
        if (t == 1023) carry_s = run;
        __syncthreads();
    }
    if (t == 0) out[n] = carry_s;
}

/* ---- level kernels ---------------------------------------------------------------------------- */

#define LVL_T 4096           /* particles per tile */
#define LVL_CMAX 512         /* children per parent (digit of up to 9 bits) */

struct LevelDesc {
    int shift;               /* digit  = (key >> shift) & (C-1)                     */
    int db;                  /* log2 C                                              */
    int pshift;              /* parent = key >> pshift   (pshift >= 32: parent 0)   */
    uint32_t n_parents;
};

/* last parent p with pstart[p] <= pos (empty parents share their start with the next one) */
__device__ __forceinline__ uint32_t find_parent(const uint32_t *__restrict__ pstart, uint32_t n_parents, uint32_t pos)
{
    uint32_t lo = 0, hi = n_parents - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(pstart + mid) <= pos) lo = mid; else hi = mid - 1;
    }
    return lo;
}

/* block-wide version: NT threads probe NT candidate parents per step (one or two dependent loads
 * instead of log2(n_parents)); result broadcast through shared memory.  All threads must call. */
template <int NT>
__device__ __forceinline__ uint32_t find_parent_block(const uint32_t *__restrict__ pstart, uint32_t n_parents,
                                                      uint32_t pos, uint32_t *sh_res)
{
    uint32_t lo = 0, hi = n_parents;            /* answer in [lo, hi) */
    while (hi - lo > 1) {
        uint32_t span = hi - lo, step = (span + NT - 1) / NT;
        uint32_t cand = lo + threadIdx.x * step;
        /* thread owns [cand, cand+step): the answer p satisfies pstart[p] <= pos, and is the LAST such p */
        bool mine = false;
        if (cand < hi) {
            uint32_t nxt = cand + step < hi ? cand + step : hi;
            bool ge = __ldg(pstart + cand) <= pos;
            bool nx = (nxt < n_parents) ? (__ldg(pstart + nxt) <= pos) : false;
            mine = ge && !nx;
        }
        __syncthreads();
        if (mine) { sh_res[0] = cand; sh_res[1] = cand + step < hi ? cand + step : hi; }
        __syncthreads();
        lo = sh_res[0]; hi = sh_res[1];
    }
    __syncthreads();
    return lo;
}

/* digit histogram of every parent bucket; level 0 computes keys from positions (and mass min/max) */
template <bool FIRST>
__global__ void __launch_bounds__(256) k_lvl_hist(const float4 *__restrict__ in4, const uint32_t *__restrict__ inkey,
                                                  int64_t n, GridDev g, LevelDesc lv,
                                                  const uint32_t *__restrict__ pstart, uint32_t *__restrict__ counts,
                                                  uint32_t *__restrict__ mass_minmax,
                                                  const uint32_t *__restrict__ n_dev)
{
    __shared__ uint32_t sh[LVL_CMAX];
    __shared__ uint32_t sres[2];
    if (n_dev) n = (int64_t)__ldg(n_dev);        /* focused build: particles kept by level 0 */
    const int C = 1 << lv.db;
    const uint32_t cmask = (uint32_t)C - 1u;
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    const int64_t ntiles = (n + LVL_T - 1) / LVL_T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint32_t pos = (uint32_t)(tile * LVL_T);
        const uint32_t tend = (uint32_t)min((int64_t)n, tile * LVL_T + LVL_T);
        while (pos < tend) {
            uint32_t p = FIRST ? 0u : find_parent_block<256>(pstart, lv.n_parents, pos, sres);
            uint32_t send = FIRST ? tend : min(tend, __ldg(pstart + p + 1));
            for (int c = threadIdx.x; c < C; c += 256) sh[c] = 0u;
            __syncthreads();
            if (FIRST) {
                /* 8 particles per thread at a time: all loads in flight before the first shared-memory atomic */
                constexpr int PPT = 8;
                for (uint32_t i0 = pos + threadIdx.x; i0 < send; i0 += 256 * PPT) {
                    float4 q[PPT];
#pragma unroll
                    for (int k = 0; k < PPT; ++k) {
                        uint32_t i = i0 + k * 256;
                        if (i < send) q[k] = ld_stream(in4 + i);
                    }
#pragma unroll
                    for (int k = 0; k < PPT; ++k) {
                        uint32_t i = i0 + k * 256;
                        if (i < send) {
                            bool kept = true;
                            uint32_t key = cell_key_kept(q[k], g, kept);
                            if (!g.indexed) {        /* (.w of indexed input is the particle's global index) */
                                uint32_t mo = (q[k].w >= 0.0f) ? __float_as_uint(q[k].w) : 0xFFFFFFFEu;
                                if (!(q[k].w >= 0.0f)) mn = 0u; /* negative / NaN mass: treated as "unequal" */
                                mn = min(mn, mo);
                                mx = max(mx, mo);
                            }
                            if (kept) atomicAdd(&sh[(key >> lv.shift) & cmask], 1u);
                        }
                    }
                }
            } else {
                /* keys only: 16 per thread, all loads in flight before the first shared-memory atomic */
                constexpr int KPT = LVL_T / 256;
                uint32_t key[KPT];
#pragma unroll
                for (int k = 0; k < KPT; ++k) {
                    uint32_t i = pos + threadIdx.x + k * 256;
                    if (i < send) key[k] = __ldg(inkey + i);
                }
#pragma unroll
                for (int k = 0; k < KPT; ++k) {
                    uint32_t i = pos + threadIdx.x + k * 256;
                    if (i < send) atomicAdd(&sh[(key[k] >> lv.shift) & cmask], 1u);
                }
            }
            __syncthreads();
            for (int c = threadIdx.x; c < C; c += 256)
                if (sh[c]) atomicAdd(&counts[(size_t)p * C + c], sh[c]);
            __syncthreads();
            pos = send;
        }
    }
    if (FIRST && !g.indexed) {
        mn = __reduce_min_sync(0xFFFFFFFFu, mn);
        mx = __reduce_max_sync(0xFFFFFFFFu, mx);
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&mass_minmax[0], mn);
            atomicMax(&mass_minmax[1], mx);
        }
    }
}

/* tile-wise counting sort by digit, staged through shared memory so that each child bucket
 * receives one contiguous run per tile segment */
#define LVL_THREADS 512
template <bool FIRST, bool WRITE_KEY>
__global__ void __launch_bounds__(LVL_THREADS, 2) k_lvl_partition(const float4 *__restrict__ in4,
                                                                  const uint32_t *__restrict__ inkey, int64_t n,
                                                                  GridDev g, LevelDesc lv,
                                                                  const uint32_t *__restrict__ pstart,
                                                                  uint32_t *__restrict__ cursor,
                                                                  float4 *__restrict__ out4,
                                                                  uint32_t *__restrict__ outkey,
                                                                  const uint32_t *__restrict__ n_dev)
{
    if (n_dev) n = (int64_t)__ldg(n_dev);
    extern __shared__ __align__(16) unsigned char raw[];
    float4 *s4 = reinterpret_cast<float4 *>(raw);
    uint32_t *sk = reinterpret_cast<uint32_t *>(s4 + LVL_T);
    uint16_t *sd = reinterpret_cast<uint16_t *>(sk + LVL_T);
    uint16_t *sr = sd + LVL_T;
    uint16_t *perm = sr + LVL_T;
    __shared__ uint32_t scnt[LVL_CMAX], soff[LVL_CMAX], sbase[LVL_CMAX], ws[LVL_THREADS / 32], sres[2], stot;
    const int C = 1 << lv.db;
    const uint32_t cmask = (uint32_t)C - 1u;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int64_t ntiles = (n + LVL_T - 1) / LVL_T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint32_t pos = (uint32_t)(tile * LVL_T);
        const uint32_t tend = (uint32_t)min((int64_t)n, tile * LVL_T + LVL_T);
        while (pos < tend) {
            const uint32_t p = FIRST ? 0u : find_parent_block<LVL_THREADS>(pstart, lv.n_parents, pos, sres);
            const uint32_t send = FIRST ? tend : min(tend, __ldg(pstart + p + 1));
            const int cnt = (int)(send - pos);
            if (t < C) scnt[t] = 0u;
            __syncthreads();
            {   /* all loads of the tile segment are issued before the first shared-memory atomic */
                constexpr int IT = LVL_T / LVL_THREADS;
                float4 q[IT];
                uint32_t key[IT];
#pragma unroll
                for (int k = 0; k < IT; ++k) {
                    int i = t + k * LVL_THREADS;
                    if (i < cnt) {
                        q[k] = ld_stream(in4 + pos + i);
                        if (!FIRST) key[k] = __ldg(inkey + pos + i);
                    }
                }
#pragma unroll
                for (int k = 0; k < IT; ++k) {
                    int i = t + k * LVL_THREADS;
                    if (i < cnt) {
                        bool kept = true;
                        if (FIRST) {
                            key[k] = cell_key_kept(q[k], g, kept);
                            if (!g.indexed) q[k].w = __int_as_float((int)(pos + i));      /* payload: original index */
                        }
                        if (kept) {
                            uint32_t d = (key[k] >> lv.shift) & cmask;
                            s4[i] = q[k];
                            sk[i] = key[k];
                            sd[i] = (uint16_t)d;
                            sr[i] = (uint16_t)atomicAdd(&scnt[d], 1u);
                        } else {
                            sd[i] = 0xFFFFu;                              /* outside the focused region */
                        }
                    }
                }
            }
            __syncthreads();
            {   /* exclusive scan of the C counts (C <= 512: one per thread) + run reservation */
                uint32_t c0 = (t < C) ? scnt[t] : 0u;
                uint32_t x = c0;
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                    if (lane >= o) x += u;
                }
                if (lane == 31) ws[w] = x;
                __syncthreads();
                uint32_t off = 0;
                for (int k = 0; k < w; ++k) off += ws[k];
                if (t < C) {
                    soff[t] = off + x - c0;
                    sbase[t] = c0 ? atomicAdd(&cursor[(size_t)p * C + t], c0) : 0u;
                    if (t == C - 1) stot = off + x;
                }
            }
            __syncthreads();
            for (int i = t; i < cnt; i += LVL_THREADS)
                if (sd[i] != 0xFFFFu) perm[soff[sd[i]] + sr[i]] = (uint16_t)i;
            __syncthreads();
            const int kept_cnt = (int)stot;
            for (int slot = t; slot < kept_cnt; slot += LVL_THREADS) {
                int i = perm[slot];
                uint32_t d = sd[i];
                uint32_t dst = sbase[d] + ((uint32_t)slot - soff[d]);
                out4[dst] = s4[i];
                if (WRITE_KEY) outkey[dst] = sk[i];
            }
            __syncthreads();
            pos = send;
        }
    }
}

/* ---- final: one CTA per bucket, counting sort by cell in shared memory ---------------------------- */

#define BKT_CAP 3072         /* particles staged in shared memory; larger buckets take the slow path */
#define BKT_CELLS 4096       /* cells per final bucket (<=)                                          */
#define BKT_THREADS 512

__global__ void __launch_bounds__(BKT_THREADS, 2) k_bucket_sort(const float4 *__restrict__ in4, GridDev g, int cell_bits,
                                                             uint32_t n_buckets, const uint32_t *__restrict__ bstart,
                                                             float4 *__restrict__ sorted, uint32_t *__restrict__ ce,
                                                             int first_pass_input_is_raw)
{
    extern __shared__ __align__(16) unsigned char raw[];
    float4 *s4 = reinterpret_cast<float4 *>(raw);
    uint32_t *cnt = reinterpret_cast<uint32_t *>(s4 + BKT_CAP);
    uint16_t *sc = reinterpret_cast<uint16_t *>(cnt + BKT_CELLS);
    uint16_t *sr = sc + BKT_CAP;
    uint16_t *perm = sr + BKT_CAP;
    __shared__ uint32_t ws[BKT_THREADS / 32];
    const int ncells = 1 << cell_bits;
    const uint32_t cmask = (uint32_t)ncells - 1u;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int per = (ncells + BKT_THREADS - 1) / BKT_THREADS;      /* cells per thread in the scan */
    for (uint32_t b = blockIdx.x; b < n_buckets; b += gridDim.x) {
        const uint32_t b0 = __ldg(bstart + b), b1 = __ldg(bstart + b + 1);
        const uint32_t nb = b1 - b0;
        const bool staged = nb <= BKT_CAP;
        if (nb == 0) {                   /* empty bucket (focused builds): only its cell-table slice */
            for (int c = t; c < ncells; c += BKT_THREADS) ce[((size_t)b << cell_bits) + c] = b0;
            continue;
        }
        for (int c = t; c < ncells; c += BKT_THREADS) cnt[c] = 0u;
        __syncthreads();
        /* count (and stage): all loads of a staged bucket are issued before the first use */
        if (staged) {
            constexpr int IT = BKT_CAP / BKT_THREADS;
            float4 q[IT];
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                uint32_t i = t + k * BKT_THREADS;
                if (i < nb) q[k] = ld_stream(in4 + b0 + i);
            }
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                uint32_t i = t + k * BKT_THREADS;
                if (i < nb) {
                    if (first_pass_input_is_raw) q[k].w = __int_as_float((int)(b0 + i));
                    uint32_t c = cell_key(q[k], g) & cmask;
                    s4[i] = q[k];
                    sc[i] = (uint16_t)c;
                    sr[i] = (uint16_t)atomicAdd(&cnt[c], 1u);
                }
            }
        } else {
            for (uint32_t i = t; i < nb; i += BKT_THREADS) {
                float4 q = ld_stream(in4 + b0 + i);
                atomicAdd(&cnt[cell_key(q, g) & cmask], 1u);
            }
        }
        __syncthreads();
        /* exclusive scan of the cell counts (in place) and the bucket's slice of the cell table */
        {
            uint32_t s = 0;
            const int c0 = t * per;
            for (int k = 0; k < per; ++k) if (c0 + k < ncells) s += cnt[c0 + k];
            uint32_t x = s;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += u;
            }
            if (lane == 31) ws[w] = x;
            __syncthreads();
            uint32_t off = 0;
            for (int k = 0; k < w; ++k) off += ws[k];
            uint32_t run = off + x - s;
            for (int k = 0; k < per; ++k)
                if (c0 + k < ncells) {
                    uint32_t v = cnt[c0 + k];
                    cnt[c0 + k] = run;
                    ce[((size_t)b << cell_bits) + c0 + k] = b0 + run;
                    run += v;
                }
        }
        __syncthreads();
        if (staged) {
            for (uint32_t i = t; i < nb; i += BKT_THREADS) perm[cnt[sc[i]] + sr[i]] = (uint16_t)i;
            __syncthreads();
            for (uint32_t slot = t; slot < nb; slot += BKT_THREADS) sorted[b0 + slot] = s4[perm[slot]];
        } else {
            /* oversized bucket (a dense halo core): second read, ranks from shared-memory cursors,
             * scattered stores inside the bucket's range */
            for (uint32_t i = t; i < nb; i += BKT_THREADS) {
                float4 q = ld_stream(in4 + b0 + i);
                if (first_pass_input_is_raw) q.w = __int_as_float((int)(b0 + i));
                uint32_t c = cell_key(q, g) & cmask;
                uint32_t dst = atomicAdd(&cnt[c], 1u);
                sorted[b0 + dst] = q;
            }
        }
        __syncthreads();
    }
}

/* ---- register-resident tile variants (default) --------------------------------------------------
 * Same scheme as k_lvl_partition / k_bucket_sort, with about half the shared-memory operations per
 * particle: a thread keeps its particles in registers while the tile's digit counts are scanned,
 * then writes each particle ONCE to its sorted position in shared memory (the staged kernels write
 * it unsorted and permute through an index array).  Smaller tiles (2048) and 45 KB of shared memory
 * per CTA let three CTAs share an SM, so one CTA's load phase overlaps the others' shuffle phases.
 * (A variant that stored from registers straight to global memory — no staging at all — was
 * measured 1.4x SLOWER: 16-byte stores to 128 different runs per warp are partial-sector writes
 * and are bound by L2 transaction rate, see profiles/.) */
#define RT_NT 256
#define RT_IT 8
#define RT_T (RT_NT * RT_IT)          /* 2048 particles per tile */
#define RT_SMEM (RT_T * (16 + 4 + 2))

template <bool FIRST, bool WRITE_KEY>
__global__ void __launch_bounds__(RT_NT, 3) k_lvl_partition_rt(const float4 *__restrict__ in4,
                                                               const uint32_t *__restrict__ inkey, int64_t n,
                                                               GridDev g, LevelDesc lv,
                                                               const uint32_t *__restrict__ pstart,
                                                               uint32_t *__restrict__ cursor,
                                                               float4 *__restrict__ out4,
                                                               uint32_t *__restrict__ outkey,
                                                               const uint32_t *__restrict__ n_dev)
{
    if (n_dev) n = (int64_t)__ldg(n_dev);
    extern __shared__ __align__(16) unsigned char raw[];      /* RT_SMEM bytes */
    float4 *s4 = reinterpret_cast<float4 *>(raw);
    uint32_t *sk = reinterpret_cast<uint32_t *>(s4 + RT_T);
    uint16_t *sd = reinterpret_cast<uint16_t *>(sk + RT_T);
    __shared__ uint32_t scnt[LVL_CMAX], soff[LVL_CMAX], sbase[LVL_CMAX], ws[RT_NT / 32], sres[2], stot;
    const int C = 1 << lv.db;
    const uint32_t cmask = (uint32_t)C - 1u;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    constexpr int CPT = LVL_CMAX / RT_NT;                     /* counts per thread in the scan */
    for (int c = t; c < LVL_CMAX; c += RT_NT) scnt[c] = 0u;
    __syncthreads();
    const int64_t ntiles = (n + RT_T - 1) / RT_T;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        uint32_t pos = (uint32_t)(tile * RT_T);
        const uint32_t tend = (uint32_t)min((int64_t)n, tile * RT_T + RT_T);
        while (pos < tend) {
            const uint32_t p = FIRST ? 0u : find_parent_block<RT_NT>(pstart, lv.n_parents, pos, sres);
            const uint32_t send = FIRST ? tend : min(tend, __ldg(pstart + p + 1));
            const int cnt = (int)(send - pos);
            float4 q[RT_IT];
            uint32_t key[RT_IT], dr[RT_IT];
#pragma unroll
            for (int k = 0; k < RT_IT; ++k) {
                int i = t + k * RT_NT;
                if (i < cnt) {
                    q[k] = ld_stream(in4 + pos + i);
                    if (!FIRST) key[k] = __ldg(inkey + pos + i);
                }
            }
#pragma unroll
            for (int k = 0; k < RT_IT; ++k) {
                int i = t + k * RT_NT;
                dr[k] = 0xFFFFFFFFu;
                if (i < cnt) {
                    bool kept = true;
                    if (FIRST) {
                        key[k] = cell_key_kept(q[k], g, kept);
                        if (!g.indexed) q[k].w = __int_as_float((int)(pos + i));      /* payload: original index */
                    }
                    if (kept) {
                        uint32_t d = (key[k] >> lv.shift) & cmask;
                        dr[k] = (d << 16) | atomicAdd(&scnt[d], 1u);
                    }
                }
            }
            __syncthreads();
            {   /* exclusive scan of the counts (CPT consecutive children per thread) + run reservation */
                uint32_t c0[CPT], sum = 0;
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    int c = t * CPT + k;
                    c0[k] = scnt[c];
                    scnt[c] = 0u;                                          /* ready for the next tile */
                    sum += c0[k];
                }
                uint32_t x = sum;
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                    if (lane >= o) x += u;
                }
                if (lane == 31) ws[w] = x;
                __syncthreads();
                uint32_t off = 0;
                for (int k = 0; k < w; ++k) off += ws[k];
                uint32_t run = off + x - sum;
#pragma unroll
                for (int k = 0; k < CPT; ++k) {
                    int c = t * CPT + k;
                    if (c < C) {
                        soff[c] = run;
                        sbase[c] = c0[k] ? atomicAdd(&cursor[(size_t)p * C + c], c0[k]) - run : 0u;
                    }
                    run += c0[k];
                }
                if (t == RT_NT - 1) stot = run;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < RT_IT; ++k) {
                if (dr[k] != 0xFFFFFFFFu) {
                    uint32_t d = dr[k] >> 16, slot = soff[d] + (dr[k] & 0xFFFFu);
                    s4[slot] = q[k];
                    if (WRITE_KEY) sk[slot] = key[k];
                    sd[slot] = (uint16_t)d;
                }
            }
            __syncthreads();
            const int kept_cnt = (int)stot;
            for (int slot = t; slot < kept_cnt; slot += RT_NT) {
                uint32_t dst = sbase[sd[slot]] + (uint32_t)slot;           /* sbase holds (run base - soff) */
                out4[dst] = s4[slot];
                if (WRITE_KEY) outkey[dst] = sk[slot];
            }
            pos = send;
        }
    }
}

/* final pass: one CTA per bucket, 256 threads, up to BR_IT particles per thread kept in registers
 * with their (cell, rank); larger buckets (dense halo cores) read their particles a second time
 * (an L2 hit).  Stores go straight to the final slot: the bucket's whole output range (16-50 KB)
 * is written by this CTA within a microsecond, so the 16-byte stores merge in L2.  Only the cell
 * counts live in shared memory (16 KB), which leaves room for 4 CTAs per SM. */
#define BR_IT 8

/* focused builds: most final buckets are empty and lie outside the focus mask.  One thread per bucket
 * settles those and lists the others for the sort kernel, which then iterates over live buckets only.
 * A settled bucket's cell-table slice is never read except its first entry: the queries check every ball
 * against the mask, so they read entries of marked cells and the entry right behind one, which lies in a
 * bucket with a marked cell (live: written in full) or is the first entry of the next bucket. */
__global__ void __launch_bounds__(256) k_bucket_live(GridDev g, int cell_bits, uint32_t n_buckets,
                                                     const uint32_t *__restrict__ bstart, uint32_t *__restrict__ ce,
                                                     uint32_t *__restrict__ live, uint32_t *__restrict__ live_n,
                                                     uint32_t big_from, uint32_t *__restrict__ big, uint32_t *__restrict__ big_n)
{
    /* big != NULL: live buckets with more than big_from particles go to a list of their own (one CTA each,
     * k_bucket_sort_sparse, concurrently with the warp-per-bucket kernel that takes the others) */
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    bool is_live = false, is_big = false;
    if (b < n_buckets) {
        const uint32_t b0 = __ldg(bstart + b), b1 = __ldg(bstart + b + 1);
        is_live = b1 > b0;
        is_big = big != nullptr && b1 - b0 > big_from;
        if (!is_live) {
            /* the bucket is a stretch of one row of cells, or a few whole rows: per row its coarse cells
             * are consecutive mask bits, tested a word at a time */
            const uint32_t ncell_b = 1u << cell_bits;
            const uint32_t per_row = min(ncell_b, (uint32_t)g.nc), nrow = ncell_b / per_row;
            const uint32_t nbits = max(1u, per_row >> g.ms);
            for (uint32_t r = 0; r < nrow && !is_live; ++r) {
                const uint32_t bit0 = cell_mask_bit(g, (b << cell_bits) + r * per_row);
                for (uint32_t k = bit0; k < bit0 + nbits && !is_live;) {
                    const uint32_t wbit = k & 31u, take = min(32u - wbit, bit0 + nbits - k);
                    const uint32_t word = __ldg(g.mask + (k >> 5)) >> wbit;
                    is_live = (take == 32u ? word : (word & ((1u << take) - 1u))) != 0u;
                    k += take;
                }
            }
            if (!is_live) ce[(size_t)b << cell_bits] = b0;
        }
    }
    const int lane = threadIdx.x & 31;
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, is_live && !is_big);
    if (m) {
        const int leader = __ffs(m) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(live_n, (uint32_t)__popc(m));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (is_live && !is_big) live[base + __popc(m & ((1u << lane) - 1u))] = b;
    }
    const uint32_t mb = __ballot_sync(0xFFFFFFFFu, is_big);
    if (mb) {
        const int leader = __ffs(mb) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(big_n, (uint32_t)__popc(mb));
        base = __shfl_sync(0xFFFFFFFFu, base, leader);
        if (is_big) big[base + __popc(mb & ((1u << lane) - 1u))] = b;
    }
}

/* `count` (<= 32) consecutive bits of the focus mask starting at bit `bit0` */
__device__ __forceinline__ uint32_t mask_bits_at(const uint32_t *__restrict__ mask, uint32_t bit0, uint32_t count)
{
    const uint32_t w = bit0 >> 5, lo = bit0 & 31u;
    uint32_t v = __ldg(mask + w) >> lo;
    if (lo + count > 32u) v |= __ldg(mask + w + 1) << (32u - lo);
    return count >= 32u ? v : (v & ((1u << count) - 1u));
}

/* BR_NT threads per bucket: 256 for buckets of ~1000 particles, 64 for the sparse buckets of focused grids
 * (a few hundred particles in 4096 cells: four times as many buckets in flight per SM) */
template <int BR_NT, int BR_MINB>
__global__ void __launch_bounds__(BR_NT, BR_MINB) k_bucket_sort_rt(const float4 *__restrict__ in4, GridDev g, int cell_bits,
                                                             uint32_t n_buckets, const uint32_t *__restrict__ bstart,
                                                             float4 *__restrict__ sorted, uint32_t *__restrict__ ce,
                                                             int first_pass_input_is_raw,
                                                             const uint32_t *__restrict__ live,
                                                             const uint32_t *__restrict__ live_n)
{
    /* cell counts, padded: thread t scans the `per` consecutive cells starting at t*per, so without padding
     * the 16-byte accesses of a warp would lie `per` words apart — the same banks for every lane (measured:
     * 82 M bank conflicts in 99 M shared-memory wavefronts, the kernel bound by exactly that).  Four words of
     * padding per chunk of `per` cells make the lanes of a quarter warp hit eight different bank groups. */
    __shared__ __align__(16) uint32_t cnt[BKT_CELLS + BKT_CELLS / 4];
    __shared__ uint32_t ws[BR_NT / 32 + 1];
    const int ncells = 1 << cell_bits;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    /* with >= 1024 cells per bucket (every build large enough to matter) `per` is a multiple of 4 (a power of
     * two) and the counts are moved as uint4 */
    const int per = (ncells + BR_NT - 1) / BR_NT;
    const bool vec = (ncells >= 4 * BR_NT);
    const int per_log = 31 - __clz(per);
    auto phys = [&](uint32_t c) -> uint32_t { return vec ? c + ((c >> per_log) << 2) : c; };
    uint4 *cnt4 = reinterpret_cast<uint4 *>(cnt);
    if (live) n_buckets = __ldg(live_n);                 /* iterate over the live buckets only (k_bucket_live) */
    uint32_t nxb = 0, nx0 = 0, nx1 = 0;
    if (blockIdx.x < n_buckets) {
        nxb = live ? __ldg(live + blockIdx.x) : blockIdx.x;
        nx0 = __ldg(bstart + nxb); nx1 = __ldg(bstart + nxb + 1);
    }
    for (uint32_t bi = blockIdx.x; bi < n_buckets; bi += gridDim.x) {
        const uint32_t b = nxb, b0 = nx0, b1 = nx1;
        const uint32_t nb = b1 - b0;
        if (bi + gridDim.x < n_buckets) {                                 /* bounds of the next bucket, early */
            nxb = live ? __ldg(live + bi + gridDim.x) : bi + gridDim.x;
            nx0 = __ldg(bstart + nxb); nx1 = __ldg(bstart + nxb + 1);
        }
        uint32_t *ceb = ce + ((size_t)b << cell_bits);
        /* Focused grids: the queries read ce[] only at cells inside the mask and at the entry right behind
         * one (every ball is checked against the mask first), so a group of four entries is stored only if
         * one of the cells [c-1, c+3] is marked: a sparse grid writes a few percent of its cell table. */
        const bool sparse = vec && g.mask != nullptr && g.nc >= 64;
        auto group_needed = [&](uint32_t c4) -> bool {                    /* group = cells 4*c4 .. 4*c4+3 of this bucket */
            if (!sparse) return true;
            const uint32_t key0 = (b << cell_bits) + 4u * c4;
            const uint32_t c = key0 & (uint32_t)(g.nc - 1);               /* first cell of the group, inside its row */
            if (c == 0u) return true;                                     /* (its predecessor is the previous row's last cell) */
            const uint32_t rk = key0 >> g.lb;
            const uint32_t m = (1u << g.tb) - 1u, lo = rk & ((1u << (2 * g.tb)) - 1u), hi = rk >> (2 * g.tb);
            const uint32_t iy = ((hi & ((1u << (g.lb - g.tb)) - 1u)) << g.tb) | (lo & m);
            const uint32_t iz = ((hi >> (g.lb - g.tb)) << g.tb) | (lo >> g.tb);
            const uint32_t mrow = ((iz >> g.ms) << (2 * g.mb)) | ((iy >> g.ms) << g.mb);
            const uint32_t m0 = (c - 1u) >> g.ms, m1 = (c + 3u) >> g.ms;
            return mask_bits_at(g.mask, mrow + m0, m1 - m0 + 1u) != 0u;
        };
        if (nb == 0) {                   /* empty bucket (focused builds): only its cell-table slice */
            if (vec) {
                uint4 *ce4 = reinterpret_cast<uint4 *>(ceb);
                for (int c4 = t; c4 < ncells / 4; c4 += BR_NT)
                    if (group_needed((uint32_t)c4)) ce4[c4] = make_uint4(b0, b0, b0, b0);
            } else {
                for (int c = t; c < ncells; c += BR_NT) ceb[c] = b0;
            }
            continue;
        }
        const bool in_regs = nb <= (uint32_t)(BR_NT * BR_IT);
        if (vec) for (int c = t; c < (ncells + (ncells >> per_log) * 4) / 4; c += BR_NT) cnt4[c] = make_uint4(0u, 0u, 0u, 0u);
        else for (int c = t; c < ncells; c += BR_NT) cnt[c] = 0u;
        __syncthreads();
        float4 q[BR_IT];
        uint32_t cr[BR_IT];
        if (in_regs) {
#pragma unroll
            for (int k = 0; k < BR_IT; ++k) {
                uint32_t i = t + k * BR_NT;
                if (i < nb) q[k] = ld_stream(in4 + b0 + i);
            }
#pragma unroll
            for (int k = 0; k < BR_IT; ++k) {
                if ((uint32_t)(k * BR_NT) >= nb) break;                   /* uniform: nothing left for anyone */
                uint32_t i = t + k * BR_NT;
                if (i < nb) {
                    if (first_pass_input_is_raw) q[k].w = __int_as_float((int)(b0 + i));
                    uint32_t c = cell_key_low(q[k], g, cell_bits);
                    cr[k] = (c << 16) | atomicAdd(&cnt[phys(c)], 1u);
                }
            }
        } else {
            for (uint32_t i = t; i < nb; i += BR_NT) {
                float4 qq = __ldg(in4 + b0 + i);
                atomicAdd(&cnt[phys(cell_key_low(qq, g, cell_bits))], 1u);
            }
        }
        __syncthreads();
        /* exclusive scan of the cell counts (in place) and the bucket's slice of the cell table */
        if (vec) {
            const int G = per / 4;
            uint32_t s = 0;
            for (int k = 0; k < G; ++k) {
                uint4 v = cnt4[t * (G + 1) + k];
                s += v.x + v.y + v.z + v.w;
            }
            uint32_t x = s;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += u;
            }
            if (lane == 31) ws[w] = x;
            __syncthreads();
            uint32_t off = 0;
#pragma unroll
            for (int k = 0; k < BR_NT / 32; ++k) off += (k < w) ? ws[k] : 0u;
            uint32_t run = off + x - s;
            for (int k = 0; k < G; ++k) {
                uint4 v = cnt4[t * (G + 1) + k], e;
                e.x = run; e.y = e.x + v.x; e.z = e.y + v.y; e.w = e.z + v.z;
                run = e.w + v.w;
                cnt4[t * (G + 1) + k] = e;
            }
            __syncthreads();
            /* the bucket's slice of the cell table, coalesced: consecutive threads store consecutive groups */
            uint4 *ce4 = reinterpret_cast<uint4 *>(ceb);                  /* (b << cell_bits) is a multiple of 1024 */
            for (int c4 = t; c4 < ncells / 4; c4 += BR_NT) {
                if (!group_needed((uint32_t)c4)) continue;
                const uint4 e = cnt4[c4 + ((4 * c4) >> per_log)];
                ce4[c4] = make_uint4(b0 + e.x, b0 + e.y, b0 + e.z, b0 + e.w);
            }
        } else {
            uint32_t s = 0;
            const int c0 = t * per;
            for (int k = 0; k < per; ++k) if (c0 + k < ncells) s += cnt[c0 + k];
            uint32_t x = s;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += u;
            }
            if (lane == 31) ws[w] = x;
            __syncthreads();
            uint32_t off = 0;
            for (int k = 0; k < w; ++k) off += ws[k];
            uint32_t run = off + x - s;
            for (int k = 0; k < per; ++k)
                if (c0 + k < ncells) {
                    uint32_t v = cnt[c0 + k];
                    cnt[c0 + k] = run;
                    ceb[c0 + k] = b0 + run;
                    run += v;
                }
        }
        __syncthreads();
        if (in_regs) {
#pragma unroll
            for (int k = 0; k < BR_IT; ++k) {
                if ((uint32_t)(k * BR_NT) >= nb) break;
                uint32_t i = t + k * BR_NT;
                if (i < nb) sorted[b0 + cnt[phys(cr[k] >> 16)] + (cr[k] & 0xFFFFu)] = q[k];
            }
        } else {
            /* oversized bucket (a dense halo core): second read (L2), ranks from the shared cursors */
            for (uint32_t i = t; i < nb; i += BR_NT) {
                float4 qq = __ldg(in4 + b0 + i);
                if (first_pass_input_is_raw) qq.w = __int_as_float((int)(b0 + i));
                uint32_t c = cell_key_low(qq, g, cell_bits);
                uint32_t dst = atomicAdd(&cnt[phys(c)], 1u);
                sorted[b0 + dst] = qq;
            }
        }
        __syncthreads();
    }
}

/* ---- final pass of a FOCUSED grid: work proportional to the marked cells -------------------------------
 * A focused grid holds particles only in the coarse cells of the mask — a few percent of the volume, spread
 * over nearly every final bucket.  The dense kernel above zeroes, scans and stores 4096 counters per bucket for
 * a few hundred particles (measured: 3000 warp instructions per bucket, 2.3 ms of an 11 ms step at 1024^3).
 * Here a bucket first reads its slice of the mask (<= 128 words), numbers its marked cells with a popcount
 * prefix, and then counts, scans and stores over those compact cells only.  The queries read ce[] at marked
 * cells and at the entry right behind one; exactly those entries are written (plus the bucket's first). */
#define BS_NT 128
#define BS_IT 8
__global__ void __launch_bounds__(BS_NT, 8) k_bucket_sort_sparse(const float4 *__restrict__ in4, GridDev g, int cell_bits,
                                                                 uint32_t n_buckets, const uint32_t *__restrict__ bstart,
                                                                 float4 *__restrict__ sorted, uint32_t *__restrict__ ce,
                                                                 const uint32_t *__restrict__ live,
                                                                 const uint32_t *__restrict__ live_n)
{
    __shared__ uint32_t mw[128], mpre[129];
    __shared__ uint32_t cnt[BKT_CELLS + 1];
    __shared__ uint32_t ws[BS_NT / 32 + 1];
    const int ncells = 1 << cell_bits;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int nrow = ncells >> g.lb;                         /* fine rows of a bucket (host: cell_bits >= lb) */
    const int wpr = (g.nc >> g.ms) >> 5;                     /* mask words per row (host: nc >> ms >= 32)     */
    const int nwords = nrow * wpr;                           /* <= 128                                        */
    const uint32_t sub = (1u << g.ms) - 1u;
    if (live) n_buckets = __ldg(live_n);
    for (uint32_t bi = blockIdx.x; bi < n_buckets; bi += gridDim.x) {
        const uint32_t b = live ? __ldg(live + bi) : bi;
        const uint32_t b0 = __ldg(bstart + b), b1 = __ldg(bstart + b + 1), nb = b1 - b0;
        uint32_t *ceb = ce + ((size_t)b << cell_bits);
        /* the bucket's mask words and their popcount prefix */
        uint32_t myw = 0u;
        if (t < nwords) {
            const int r = t / wpr, j = t - r * wpr;
            const uint32_t rk = (b << (cell_bits - g.lb)) + (uint32_t)r;           /* row key */
            const uint32_t m = (1u << g.tb) - 1u, lo = rk & ((1u << (2 * g.tb)) - 1u), hi = rk >> (2 * g.tb);
            const uint32_t iy = ((hi & ((1u << (g.lb - g.tb)) - 1u)) << g.tb) | (lo & m);
            const uint32_t iz = ((hi >> (g.lb - g.tb)) << g.tb) | (lo >> g.tb);
            const uint32_t mrow = ((iz >> g.ms) << (2 * g.mb)) | ((iy >> g.ms) << g.mb);
            myw = __ldg(g.mask + (mrow >> 5) + j);
        }
        float4 q[BS_IT];
        const bool in_regs = nb <= (uint32_t)(BS_NT * BS_IT);
        if (in_regs) {
#pragma unroll
            for (int k = 0; k < BS_IT; ++k) {
                const uint32_t i = t + k * BS_NT;
                if (i < nb) q[k] = ld_stream(in4 + b0 + i);
            }
        }
        {
            uint32_t x = __popc(myw);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += u;
            }
            if (lane == 31) ws[w] = x;
            __syncthreads();
            uint32_t off = 0;
#pragma unroll
            for (int k = 0; k < BS_NT / 32; ++k) off += (k < w) ? ws[k] : 0u;
            mw[t] = myw;
            mpre[t] = off + x - __popc(myw);
            if (t == BS_NT - 1) mpre[BS_NT] = off + x;
        }
        __syncthreads();
        const uint32_t ncomp = mpre[BS_NT] << g.ms;          /* compact (marked) fine cells of this bucket */
        for (uint32_t i = t; i <= ncomp; i += BS_NT) cnt[i] = 0u;
        __syncthreads();
        auto compact = [&](uint32_t c) -> uint32_t {          /* cell inside the bucket -> compact index */
            const uint32_t r = c >> g.lb, x = c & (uint32_t)(g.nc - 1), xc = x >> g.ms;
            const uint32_t wi = r * (uint32_t)wpr + (xc >> 5), bp = xc & 31u;
            return ((mpre[wi] + __popc(mw[wi] & ((1u << bp) - 1u))) << g.ms) | (x & sub);
        };
        uint32_t cr[BS_IT];
        if (in_regs) {
#pragma unroll
            for (int k = 0; k < BS_IT; ++k) {
                const uint32_t i = t + k * BS_NT;
                if (i < nb) {
                    const uint32_t ci = compact(cell_key_low(q[k], g, cell_bits));
                    cr[k] = (ci << 16) | atomicAdd(&cnt[ci], 1u);
                }
            }
        } else {
            for (uint32_t i = t; i < nb; i += BS_NT) {
                const float4 qq = __ldg(in4 + b0 + i);
                atomicAdd(&cnt[compact(cell_key_low(qq, g, cell_bits))], 1u);
            }
        }
        __syncthreads();
        /* exclusive scan over the compact cells: every warp scans its quarter 32 cells at a time with a running
         * carry (no block barrier inside), the warps' totals are added by the readers; cnt[ncomp] = particles */
        {
            const uint32_t Q = ((ncomp + (BS_NT / 32) - 1) / (BS_NT / 32) + 31u) & ~31u;      /* cells per warp */
            const uint32_t lo = (uint32_t)w * Q, hi = min(ncomp, lo + Q);
            uint32_t carry = 0;
            for (uint32_t base = lo; base < hi; base += 32) {
                const uint32_t i = base + lane;
                const uint32_t v = i < hi ? cnt[i] : 0u;
                uint32_t x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                    if (lane >= o) x += u;
                }
                if (i < hi) cnt[i] = carry + x - v;
                carry += __shfl_sync(0xFFFFFFFFu, x, 31);
            }
            if (lane == 0) ws[w] = carry;
            __syncthreads();
            uint32_t off = 0;
#pragma unroll
            for (int k = 0; k < BS_NT / 32; ++k) off += (k < w) ? ws[k] : 0u;
            if (off) for (uint32_t i = lo + lane; i < hi; i += 32) cnt[i] += off;
            if (t == 0) {
                uint32_t tot = 0;
                for (int k = 0; k < BS_NT / 32; ++k) tot += ws[k];
                cnt[ncomp] = tot;
            }
        }
        __syncthreads();
        /* particles to their final slots */
        if (in_regs) {
#pragma unroll
            for (int k = 0; k < BS_IT; ++k) {
                const uint32_t i = t + k * BS_NT;
                if (i < nb) sorted[b0 + cnt[cr[k] >> 16] + (cr[k] & 0xFFFFu)] = q[k];
            }
        }
        /* cell table: the bucket's first entry, every marked cell, and the entry right behind a marked cell */
        if (t == 0) ceb[0] = b0;
        if (t < nwords) {
            const int r = t / wpr, j = t - r * wpr;
            uint32_t word = mw[t], k0 = mpre[t];
            while (word) {
                const int bp = __ffs(word) - 1;
                word &= word - 1u;
                const uint32_t xc = (uint32_t)j * 32u + (uint32_t)bp;
                for (uint32_t sx = 0; sx <= sub; ++sx) {
                    const uint32_t c = ((uint32_t)r << g.lb) + ((xc << g.ms) | sx), kc = (k0 << g.ms) | sx;
                    ceb[c] = b0 + cnt[kc];
                    /* the next cell in key order: written here unless it is marked itself (then it writes its own
                     * start, the same value) or lies in the next bucket (whose first entry is always written) */
                    if (sx == sub && c + 1u < (uint32_t)ncells) ceb[c + 1u] = b0 + cnt[kc + 1u];
                }
                ++k0;
            }
        }
        __syncthreads();
        if (!in_regs) {
            /* oversized bucket (a dense halo core): second read (L2), slots from the scanned counters as cursors */
            for (uint32_t i = t; i < nb; i += BS_NT) {
                const float4 qq = __ldg(in4 + b0 + i);
                const uint32_t dst = atomicAdd(&cnt[compact(cell_key_low(qq, g, cell_bits))], 1u);
                sorted[b0 + dst] = qq;
            }
            __syncthreads();
        }
    }
}

/* The same for the many SMALL live buckets, one WARP per bucket: at 1024^3 a final bucket of a focused grid is a
 * pencil of 4 rows of cells that meets three or four halos — ~240 particles in ~80 marked cells — and a CTA spent
 * its time in barriers and in the latency chain bounds -> mask words -> particles -> stores of one bucket at a
 * time (ncu: 40 % of the stall samples on those three loads, 3.2 warps per issue at the barrier).  Here 32 warps
 * per SM each own a bucket: no block barrier, two mask words and up to BW_IT particles per lane; the particles are
 * read twice (the second time from L1/L2) instead of being held in registers.  Buckets with more than
 * 32 x BW_IT particles or more than BW_CAP marked cells are appended to `big` for k_bucket_sort_sparse. */
#define BW_NT 256
#define BW_IT 16
#define BW_CAP 1024
__global__ void __launch_bounds__(BW_NT, 4) k_bucket_sort_sparse_warp(const float4 *__restrict__ in4, GridDev g, int cell_bits,
                                                                      const uint32_t *__restrict__ bstart,
                                                                      float4 *__restrict__ sorted, uint32_t *__restrict__ ce,
                                                                      const uint32_t *__restrict__ live,
                                                                      const uint32_t *__restrict__ live_n,
                                                                      uint32_t *__restrict__ big, uint32_t *__restrict__ big_n)
{
    __shared__ uint32_t s_mw[BW_NT / 32][128], s_mpre[BW_NT / 32][132];
    __shared__ uint32_t s_cnt[BW_NT / 32][BW_CAP + 4];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint32_t *mw = s_mw[w], *mpre = s_mpre[w], *cnt = s_cnt[w];
    const int ncells = 1 << cell_bits;
    const int nrow = ncells >> g.lb;                         /* fine rows of a bucket (host: cell_bits >= lb) */
    const int wpr = (g.nc >> g.ms) >> 5;                     /* mask words per row (host: nc >> ms >= 32)     */
    const int nwords = nrow * wpr;                           /* <= 128                                        */
    const uint32_t sub = (1u << g.ms) - 1u;
    const uint32_t n_buckets = __ldg(live_n);
    const uint32_t gw = blockIdx.x * (BW_NT / 32) + (uint32_t)w, nwarp = gridDim.x * (BW_NT / 32);
    for (uint32_t bi = gw; bi < n_buckets; bi += nwarp) {
        const uint32_t b = __ldg(live + bi);
        const uint32_t b0 = __ldg(bstart + b), b1 = __ldg(bstart + b + 1), nb = b1 - b0;
        if (nb > 32u * BW_IT) {
            if (lane == 0) big[atomicAdd(big_n, 1u)] = b;
            continue;
        }
        /* the bucket's mask words and their popcount prefix */
        uint32_t carry = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int t = lane + 32 * k;
            uint32_t myw = 0u;
            if (t < nwords) {
                const int r = t / wpr, j = t - r * wpr;
                const uint32_t rk = (b << (cell_bits - g.lb)) + (uint32_t)r;           /* row key */
                const uint32_t m = (1u << g.tb) - 1u, lo = rk & ((1u << (2 * g.tb)) - 1u), hi = rk >> (2 * g.tb);
                const uint32_t iy = ((hi & ((1u << (g.lb - g.tb)) - 1u)) << g.tb) | (lo & m);
                const uint32_t iz = ((hi >> (g.lb - g.tb)) << g.tb) | (lo >> g.tb);
                const uint32_t mrow = ((iz >> g.ms) << (2 * g.mb)) | ((iy >> g.ms) << g.mb);
                myw = __ldg(g.mask + (mrow >> 5) + j);
            }
            const uint32_t pc = __popc(myw);
            uint32_t x = pc;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                if (lane >= o) x += u;
            }
            mw[t] = myw;
            mpre[t] = carry + x - pc;
            carry += __shfl_sync(0xFFFFFFFFu, x, 31);
        }
        const uint32_t ncomp = carry << g.ms;                /* compact (marked) fine cells of this bucket */
        if (ncomp > BW_CAP) {
            if (lane == 0) big[atomicAdd(big_n, 1u)] = b;
            __syncwarp();
            continue;
        }
        for (uint32_t i = lane; i <= ncomp; i += 32) cnt[i] = 0u;
        __syncwarp();
        auto compact = [&](uint32_t c) -> uint32_t {          /* cell inside the bucket -> compact index */
            const uint32_t r = c >> g.lb, x = c & (uint32_t)(g.nc - 1), xc = x >> g.ms;
            const uint32_t wi = r * (uint32_t)wpr + (xc >> 5), bp = xc & 31u;
            return ((mpre[wi] + __popc(mw[wi] & ((1u << bp) - 1u))) << g.ms) | (x & sub);
        };
        uint32_t cr[BW_IT];
#pragma unroll
        for (int k = 0; k < BW_IT; ++k) {
            const uint32_t i = lane + 32 * k;
            if (i < nb) {
                const float4 q = __ldg(in4 + b0 + i);
                const uint32_t ci = compact(cell_key_low(q, g, cell_bits));
                cr[k] = (ci << 16) | atomicAdd(&cnt[ci], 1u);
            }
        }
        __syncwarp();
        {   /* exclusive scan over the compact cells, 32 at a time with a running carry; cnt[ncomp] = particles */
            uint32_t run = 0;
            for (uint32_t base = 0; base < ncomp; base += 32) {
                const uint32_t i = base + lane;
                const uint32_t v = i < ncomp ? cnt[i] : 0u;
                uint32_t x = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o);
                    if (lane >= o) x += u;
                }
                if (i < ncomp) cnt[i] = run + x - v;
                run += __shfl_sync(0xFFFFFFFFu, x, 31);
            }
            if (lane == 0) cnt[ncomp] = run;
        }
        __syncwarp();
        /* particles to their final slots (second read: L1 / L2) */
#pragma unroll
        for (int k = 0; k < BW_IT; ++k) {
            const uint32_t i = lane + 32 * k;
            if (i < nb) sorted[b0 + cnt[cr[k] >> 16] + (cr[k] & 0xFFFFu)] = __ldg(in4 + b0 + i);
        }
        /* cell table: the bucket's first entry, every marked cell, and the entry right behind a marked cell */
        uint32_t *ceb = ce + ((size_t)b << cell_bits);
        if (lane == 0) ceb[0] = b0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int t = lane + 32 * k;
            if (t < nwords) {
                const int r = t / wpr, j = t - r * wpr;
                uint32_t word = mw[t], k0 = mpre[t];
                while (word) {
                    const int bp = __ffs(word) - 1;
                    word &= word - 1u;
                    const uint32_t xc = (uint32_t)j * 32u + (uint32_t)bp;
                    for (uint32_t sx = 0; sx <= sub; ++sx) {
                        const uint32_t c = ((uint32_t)r << g.lb) + ((xc << g.ms) | sx), kc = (k0 << g.ms) | sx;
                        ceb[c] = b0 + cnt[kc];
                        if (sx == sub && c + 1u < (uint32_t)ncells) ceb[c + 1u] = b0 + cnt[kc + 1u];
                    }
                    ++k0;
                }
            }
        }
        __syncwarp();
    }
}
