// microbenchmarks: smem atomicExch / atomicMin / match_any / LDS random throughput per SM
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int mode, int iters, unsigned long long* out, int* sink)
{
    __shared__ int tab[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) tab[i] = -1;
    __syncthreads();
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    int acc = 0;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        x = x * 1664525u + 1013904223u;
        int slot = (x >> 10) & 2047;
        if (mode == 0) acc += atomicExch(&tab[slot], it);
        else if (mode == 1) acc += atomicMin(&tab[slot], it);
        else if (mode == 2) acc += __popc(__match_any_sync(0xffffffffu, slot));
        else if (mode == 3) acc += tab[slot];
        else if (mode == 4) { tab[slot] = it; }
        else if (mode == 5) acc += atomicCAS(&tab[slot], -1, it);
        else if (mode == 6) { slot = (threadIdx.x + it) & 2047; acc += atomicExch(&tab[slot], it); }  // conflict-free
        else if (mode == 7) { acc += slot; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    if (acc == 123456789) sink[0] = acc;
}
int main()
{
    unsigned long long* out; int* sink;
    cudaMallocManaged(&out, 1024 * 8); cudaMalloc(&sink, 4);
    const char* names[] = {"atomicExch rand", "atomicMin rand", "match_any", "LDS rand", "STS rand", "atomicCAS rand", "atomicExch noconflict", "baseline"};
    for (int threads : {256, 1024})
    for (int mode = 0; mode < 8; mode++) {
        int iters = 2000;
        k<<<1, threads>>>(mode, iters, out, sink);
        cudaDeviceSynchronize();
        k<<<1, threads>>>(mode, iters, out, sink);
        cudaDeviceSynchronize();
        double cyc = (double)out[0] / iters;
        printf("threads %4d  %-24s %7.1f cycles/iter  -> %.2f cycles per warp-instr (SM-wide)\n", threads, names[mode], cyc, cyc / (threads / 32));
    }
    return 0;
}
