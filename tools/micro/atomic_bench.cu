// Research microbenchmark: throughput of one-lane-per-warp atomicAdd with return value on ONE global address
// (the pattern of the wavefront queues' slot allocation), against K distinct addresses and against REDG.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_bench atomic_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_atom(unsigned* ctr, int n_addr, int iters, unsigned* sink, int mode) {
    const unsigned lane = threadIdx.x & 31u, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned acc = 0;
    unsigned* p = ctr + 32 * (warp % n_addr);   // 128 B apart
    for (int i = 0; i < iters; i++) {
        if (mode == 0) { unsigned s = 0; if (lane == 0) s = atomicAdd(p, 7u); acc += __shfl_sync(0xffffffffu, s, 0); }
        else if (lane == 0) atomicAdd(p, 7u);   // result unused: RED
    }
    if (acc == 0x12345678u) sink[0] = acc;
}
int main() {
    unsigned *ctr, *sink; cudaMalloc(&ctr, 1 << 20); cudaMalloc(&sink, 4); cudaMemset(ctr, 0, 1 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * 4, block = 256, iters = 400;
    for (int mode = 0; mode < 2; mode++)
    for (int n_addr : {1, 2, 4, 8, 64, 4736}) {
        k_atom<<<grid, block>>>(ctr, n_addr, iters, sink, mode); cudaDeviceSynchronize();
        cudaEventRecord(e0); k_atom<<<grid, block>>>(ctr, n_addr, iters, sink, mode); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)grid * block / 32 * iters;
        printf("mode %s addresses %5d: %.3f ms, %.2f atomics/ns\n", mode ? "red " : "atom", n_addr, ms, ops / (ms * 1e6));
    }
    return 0;
}
