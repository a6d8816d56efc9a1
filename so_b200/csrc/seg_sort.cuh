/* seg_sort.cuh — segmented sort of 64-bit keys: the qsort(CmpList) of kd2.c:425-435,514,781 for whole batches of
 * member / ball lists at once.
 *
 * Keys are (fDist2 bits << 32) | particle index, so they are unique and ascending key order is the reference's
 * order for distinct r^2.  Segments are the CSR lists [off[s], off[s+1]); they range from 8 to > 10^6 keys.
 *
 *   k_segsort_tiles   every segment is cut into tiles of SS_T keys (tiles never straddle a segment); one CTA
 *                     sorts a tile in shared memory (bitonic network, warp-local stages without block barriers)
 *   k_segsort_merge   ceil(log2(tiles of the longest segment)) passes; in a pass one CTA produces one output tile
 *                     of the merge of two sorted runs of its segment: merge-path split by binary search in
 *                     global memory, the two pieces staged in shared memory, every thread merges SS_VT keys
 *
 * Work distribution: tile_base[s] = exclusive scan of ceil(n_s / SS_T); CTA g finds its segment by binary
 * search.  No library code: this replaces cub::DeviceSegmentedSort of round 1. */
#pragma once

#define SS_T 2048
#define SS_NT 256
#define SS_VT (SS_T / SS_NT)

/* tiles per segment -> exclusive scan (one block; nseg up to a few 10^5) + the longest segment */
__global__ void __launch_bounds__(1024) k_segsort_plan(const unsigned long long *__restrict__ off, int nseg,
                                                       uint32_t *__restrict__ tile_base, unsigned long long *__restrict__ max_n)
{
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry;
    __shared__ unsigned long long smax[32];
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long mymax = 0ull;
    for (int b0 = 0; b0 < nseg; b0 += 1024) {
        const int s = b0 + threadIdx.x;
        unsigned long long n = 0ull;
        if (s < nseg) n = off[s + 1] - off[s];
        mymax = n > mymax ? n : mymax;
        const uint32_t v = (uint32_t)((n + SS_T - 1) / SS_T);
        uint32_t x = v;
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
        const uint32_t incl = carry + x + (w ? ws[w - 1] : 0u);
        if (s < nseg) tile_base[s] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long u = __shfl_xor_sync(0xFFFFFFFFu, mymax, o);
        mymax = u > mymax ? u : mymax;
    }
    if (lane == 0) smax[w] = mymax;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long m = 0ull;
        for (int k = 0; k < 32; ++k) m = smax[k] > m ? smax[k] : m;
        tile_base[nseg] = carry;
        *max_n = m;
    }
}

/* segment and tile of the g-th global tile (all threads of a CTA compute the same) */
__device__ __forceinline__ bool segsort_locate(const uint32_t *__restrict__ tile_base, int nseg, uint32_t g, int &seg, uint32_t &k)
{
    if (g >= __ldg(tile_base + nseg)) return false;
    int lo = 0, hi = nseg - 1;                       /* last segment with tile_base <= g (empty segments share their base) */
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(tile_base + mid) <= g) lo = mid; else hi = mid - 1;
    }
    seg = lo;
    k = g - __ldg(tile_base + lo);
    return true;
}

__global__ void __launch_bounds__(SS_NT) k_segsort_tiles(const unsigned long long *__restrict__ off, int nseg,
                                                         const uint32_t *__restrict__ tile_base,
                                                         const unsigned long long *__restrict__ src,
                                                         unsigned long long *__restrict__ dst)
{
    __shared__ unsigned long long key[SS_T];
    for (uint32_t g = blockIdx.x;; g += gridDim.x) {
        int seg; uint32_t k;
        if (!segsort_locate(tile_base, nseg, g, seg, k)) break;
        const unsigned long long a = off[seg] + (unsigned long long)k * SS_T, b = min(off[seg + 1], a + SS_T);
        const int n = (int)(b - a);
        int P = 64;
        while (P < n) P <<= 1;
        for (int i = threadIdx.x; i < P; i += SS_NT) key[i] = i < n ? src[a + i] : ~0ull;
        __syncthreads();
        bitonic_sort<SS_NT>(key, P, threadIdx.x);
        for (int i = threadIdx.x; i < n; i += SS_NT) dst[a + i] = key[i];
        __syncthreads();
    }
}

/* number of keys of A that come before the `diag`-th output of merge(A[0..na), B[0..nb)) */
template <typename P>
__device__ __forceinline__ uint32_t merge_path(P A, uint32_t na, P B, uint32_t nb, uint32_t diag)
{
    uint32_t lo = diag > nb ? diag - nb : 0u, hi = diag < na ? diag : na;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (A[mid] < B[diag - 1u - mid]) lo = mid + 1u; else hi = mid;      /* keys are unique */
    }
    return lo;
}

__global__ void __launch_bounds__(SS_NT) k_segsort_merge(const unsigned long long *__restrict__ off, int nseg,
                                                         const uint32_t *__restrict__ tile_base, unsigned long long run,
                                                         const unsigned long long *__restrict__ src,
                                                         unsigned long long *__restrict__ dst)
{
    __shared__ unsigned long long sk[SS_T];
    __shared__ uint32_t sdiag[2];
    for (uint32_t g = blockIdx.x;; g += gridDim.x) {
        int seg; uint32_t k;
        if (!segsort_locate(tile_base, nseg, g, seg, k)) break;
        const unsigned long long s0 = off[seg], s1 = off[seg + 1], n = s1 - s0;
        const unsigned long long o0 = (unsigned long long)k * SS_T;                 /* output tile, relative to the segment */
        const uint32_t on = (uint32_t)min((unsigned long long)SS_T, n - o0);
        if (n <= run) {                                                         /* the segment is one sorted run already */
            for (uint32_t i = threadIdx.x; i < on; i += SS_NT) dst[s0 + o0 + i] = src[s0 + o0 + i];
            continue;
        }
        const unsigned long long p0 = (o0 / (2ull * run)) * (2ull * run);          /* start of the pair of runs */
        const unsigned long long a0 = p0, a1 = min(n, p0 + run), b0 = a1, b1 = min(n, p0 + 2ull * run);
        const uint32_t na = (uint32_t)(a1 - a0), nb = (uint32_t)(b1 - b0);
        const unsigned long long *A = src + s0 + a0, *B = src + s0 + b0;
        if (threadIdx.x < 2) {
            const uint32_t d = (uint32_t)(o0 - p0) + (threadIdx.x ? on : 0u);
            sdiag[threadIdx.x] = merge_path(A, na, B, nb, d);
        }
        __syncthreads();
        const uint32_t d0 = (uint32_t)(o0 - p0);
        const uint32_t ia0 = sdiag[0], ia1 = sdiag[1], ib0 = d0 - ia0, ib1 = d0 + on - ia1;
        const uint32_t ca = ia1 - ia0, cb = ib1 - ib0;                              /* ca + cb == on */
        for (uint32_t i = threadIdx.x; i < ca; i += SS_NT) sk[i] = A[ia0 + i];
        for (uint32_t i = threadIdx.x; i < cb; i += SS_NT) sk[ca + i] = B[ib0 + i];
        __syncthreads();
        {   /* every thread merges SS_VT consecutive outputs */
            const uint32_t t0 = min(on, threadIdx.x * SS_VT), t1 = min(on, t0 + SS_VT);
            const unsigned long long *SA = sk, *SB = sk + ca;
            uint32_t i = merge_path(SA, ca, SB, cb, t0), j = t0 - i;
            unsigned long long outv[SS_VT];
#pragma unroll
            for (int q = 0; q < SS_VT; ++q) {
                if (t0 + q < t1) {
                    const bool takeA = j >= cb || (i < ca && SA[i] < SB[j]);
                    outv[q] = takeA ? SA[i] : SB[j];
                    if (takeA) ++i; else ++j;
                }
            }
            __syncthreads();
#pragma unroll
            for (int q = 0; q < SS_VT; ++q)
                if (t0 + q < t1) sk[t0 + q] = outv[q];
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < on; i += SS_NT) dst[s0 + o0 + i] = sk[i];
        __syncthreads();
    }
}
