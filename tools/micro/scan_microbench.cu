// Microbenchmark: single-block exclusive scan of M counts (the bucket-table scan of the grid build).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scan_mb scan_microbench.cu && ./scan_mb
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NT, int PER>
__global__ void __launch_bounds__(NT) scan_one(const uint32_t *a, int64_t n, uint32_t *out, uint32_t *copy)
{
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0u;
    __syncthreads();
    for (int64_t c0 = 0; c0 < n; c0 += NT * PER) {
        const int64_t base = c0 + (int64_t)t * PER;
        uint32_t v[PER], s = 0;
        if (base + PER <= n) {
            const uint4 *a4 = reinterpret_cast<const uint4 *>(a + base);
#pragma unroll
            for (int k = 0; k < PER / 4; ++k) { uint4 x = a4[k]; v[4*k]=x.x; v[4*k+1]=x.y; v[4*k+2]=x.z; v[4*k+3]=x.w; }
        } else {
#pragma unroll
            for (int k = 0; k < PER; ++k) v[k] = (base + k < n) ? a[base + k] : 0u;
        }
#pragma unroll
        for (int k = 0; k < PER; ++k) s += v[k];
        uint32_t x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += u; }
        if (lane == 31) ws[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t y = (lane < NT / 32) ? ws[lane] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, y, o); if (lane >= o) y += u; }
            ws[lane] = y;
        }
        __syncthreads();
        uint32_t run = carry_s + (w ? ws[w - 1] : 0u) + x - s;
        if (base + PER <= n) {
            uint4 *o4 = reinterpret_cast<uint4 *>(out + base), *c4 = reinterpret_cast<uint4 *>(copy + base);
#pragma unroll
            for (int k = 0; k < PER / 4; ++k) {
                uint4 e; e.x = run; e.y = e.x + v[4*k]; e.z = e.y + v[4*k+1]; e.w = e.z + v[4*k+2]; run = e.w + v[4*k+3];
                o4[k] = e; c4[k] = e;
            }
        } else {
#pragma unroll
            for (int k = 0; k < PER; ++k) { if (base + k < n) { out[base + k] = run; copy[base + k] = run; } run += v[k]; }
        }
        __syncthreads();
        if (t == NT - 1) carry_s = run;
        __syncthreads();
    }
    if (t == 0) out[n] = carry_s;
}

// the version in grid_build.cuh at the time of the experiment: scalar stores
template <int NT, int PER>
__global__ void __launch_bounds__(NT) scan_scalar_store(const uint32_t *a, int64_t n, uint32_t *out, uint32_t *copy)
{
    __shared__ uint32_t ws[32];
    __shared__ uint32_t carry_s;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    if (t == 0) carry_s = 0u;
    __syncthreads();
    for (int64_t c0 = 0; c0 < n; c0 += NT * PER) {
        const int64_t base = c0 + (int64_t)t * PER;
        uint32_t v[PER], s = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) { v[k] = (base + k < n) ? a[base + k] : 0u; s += v[k]; }
        uint32_t x = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, x, o); if (lane >= o) x += u; }
        if (lane == 31) ws[w] = x;
        __syncthreads();
        if (w == 0) {
            uint32_t y = (lane < NT / 32) ? ws[lane] : 0u;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t u = __shfl_up_sync(0xFFFFFFFFu, y, o); if (lane >= o) y += u; }
            ws[lane] = y;
        }
        __syncthreads();
        uint32_t run = carry_s + (w ? ws[w - 1] : 0u) + x - s;
#pragma unroll
        for (int k = 0; k < PER; ++k) { if (base + k < n) { out[base + k] = run; copy[base + k] = run; } run += v[k]; }
        __syncthreads();
        if (t == NT - 1) carry_s = run;
        __syncthreads();
    }
    if (t == 0) out[n] = carry_s;
}

template <typename F> float time_it(F f)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 5; ++i) f();
    cudaEventRecord(a);
    for (int i = 0; i < 100; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms * 10.0f;   // microseconds per launch
}

int main()
{
    for (int64_t M : {128LL, 16384LL, 262144LL}) {
        uint32_t *a, *o, *c;
        cudaMalloc(&a, (M + 1) * 4); cudaMalloc(&o, (M + 1) * 4); cudaMalloc(&c, (M + 1) * 4);
        cudaMemset(a, 1, (M + 1) * 4);
        printf("M=%lld  scalar-store 1024x16: %.2f us   vec 1024x16: %.2f   vec 512x16: %.2f   vec 256x16: %.2f   vec 1024x4: %.2f   vec 256x64: %.2f   empty kernel-to-kernel: ",
               (long long)M,
               time_it([&] { scan_scalar_store<1024, 16><<<1, 1024>>>(a, M, o, c); }),
               time_it([&] { scan_one<1024, 16><<<1, 1024>>>(a, M, o, c); }),
               time_it([&] { scan_one<512, 16><<<1, 512>>>(a, M, o, c); }),
               time_it([&] { scan_one<256, 16><<<1, 256>>>(a, M, o, c); }),
               time_it([&] { scan_one<1024, 4><<<1, 1024>>>(a, M, o, c); }),
               time_it([&] { scan_one<256, 64><<<1, 256>>>(a, M, o, c); }));
        printf("%.2f us\n", time_it([&] { scan_one<32, 4><<<1, 32>>>(a, 0, o, c); }));
        cudaFree(a); cudaFree(o); cudaFree(c);
    }
    return 0;
}
