// sector_bench.cu — random-gather micro-benchmark used to find the B200 memory system's
// behaviour for the access pattern of occ lookups (profiles/README.md).  Not part of the product.
//   ./sector_bench <l2_fetch_granularity> <table_GiB>
// Modes: A  one thread loads one random 32-B sector (LDG.256)
//        B  one thread loads a random 64-B aligned pair (2 x LDG.256)
//        C  two adjacent lanes load the two sectors of a random 64-B pair (one wavefront)
//        D  four adjacent lanes load the four sectors of a random 128-B line
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
struct alignas(32) S32 { uint32_t v[8]; };
__device__ __forceinline__ uint32_t ld32(const S32 *p)
{
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ uint32_t ld32ca(const S32 *p)
{ /* same sector through L1 (allocating) */
    uint32_t a, b, c, d, e, f, g, h;
    asm volatile("ld.global.ca.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
    return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}
__device__ __forceinline__ uint32_t ld4(const S32 *p)
{ /* one word of a random sector, L1 allocating (the width-record access) */
    uint32_t a;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(a) : "l"(p));
    return a;
}
template <int MODE, int MLP>
__global__ void __launch_bounds__(256) gather(const S32 *t, uint64_t n_sectors, int iters, unsigned long long *sink)
{
    const int lane = threadIdx.x & 31;
    const int group = MODE == 2 ? 2 : (MODE == 3 ? 4 : 1);
    uint64_t s = ((blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) / group) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
    const uint64_t unit = MODE == 0 ? 1 : (MODE == 3 ? 4 : 2);
    const uint64_t units = n_sectors / unit;
    uint32_t acc = 0;
    for (int i = 0; i < iters; ++i) {
        uint64_t idx[MLP];
#pragma unroll
        for (int j = 0; j < MLP; ++j) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            idx[j] = (uint64_t)(((unsigned __int128)(s >> 11) * units) >> 53) * unit;
        }
#pragma unroll
        for (int j = 0; j < MLP; ++j) {
            if (MODE == 0) acc += ld32(t + idx[j]);
            else if (MODE == 1) { acc += ld32(t + idx[j]); acc += ld32(t + idx[j] + 1); }
            else if (MODE == 4) acc += ld32ca(t + idx[j]);
            else if (MODE == 5) acc += ld4(t + idx[j]);
            else if (MODE == 2) acc += ld32(t + idx[j] + (lane & 1));
            else acc += ld32(t + idx[j] + (lane & 3));
        }
    }
    if (acc == 0x7fffffffu) atomicAdd(sink, 1ull);
}
template <int MODE, int MLP>
static void run(const char *name, const S32 *t, uint64_t n_sectors, unsigned long long *sink, int sms, int bytes_per_thread_access)
{
    const int blocks = sms * 8, threads = 256, iters = 256 / MLP * 4;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
        CK(cudaEventRecord(e0));
        gather<MODE, MLP><<<blocks, threads>>>(t, n_sectors, iters, sink);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 && ms < best) best = ms;
    }
    double acc = (double)blocks * threads * iters * MLP;
    printf("%-44s MLP=%d  %8.1f G thread-loads/s  %8.1f GB/s useful\n", name, MLP, acc / best / 1e6,
           acc * bytes_per_thread_access / best / 1e6);
}
int main(int argc, char **argv)
{
    size_t gran = argc > 1 ? atoi(argv[1]) : 0;
    double gib = argc > 2 ? atof(argv[2]) : 3.0;
    if (gran) CK(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran));
    size_t got = 0; CK(cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity));
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    uint64_t n_sectors = (uint64_t)(gib * (1ull << 30)) / 32;
    S32 *t; CK(cudaMalloc(&t, n_sectors * 32)); CK(cudaMemset(t, 1, n_sectors * 32));
    unsigned long long *sink; CK(cudaMalloc(&sink, 8)); CK(cudaMemset(sink, 0, 8));
    printf("L2 fetch granularity asked %zu got %zu; table %.1f GiB; %d SMs\n", gran, got, gib, p.multiProcessorCount);
    run<0, 4>("A one 32B sector / thread", t, n_sectors, sink, p.multiProcessorCount, 32);
    run<0, 8>("A one 32B sector / thread", t, n_sectors, sink, p.multiProcessorCount, 32);
    run<0, 16>("A one 32B sector / thread", t, n_sectors, sink, p.multiProcessorCount, 32);
    run<1, 8>("B 64B pair by one thread (2 loads)", t, n_sectors, sink, p.multiProcessorCount, 64);
    run<2, 8>("C 64B pair by two lanes", t, n_sectors, sink, p.multiProcessorCount, 32);
    run<3, 8>("D 128B line by four lanes", t, n_sectors, sink, p.multiProcessorCount, 32);
    run<4, 8>("E one 32B sector / thread, ld.ca", t, n_sectors, sink, p.multiProcessorCount, 32);
    run<5, 8>("F one 4B word / thread, ld.ca", t, n_sectors, sink, p.multiProcessorCount, 4);
    return 0;
}
