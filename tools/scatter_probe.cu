// Developer probe: sustained rate of scattered FP64 stores as a function of the run length and alignment, into a
// buffer much larger than L2 (the situation of the varf scatter epilogue: 64-byte runs into a 13-52 GB matrix).
// Each warp writes `run`-double segments at pseudo-random positions (a different DRAM page every time); `off` shifts
// every start by that many doubles (0 = aligned to the run length, 2 = 16-byte aligned only).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/scatter_probe tools/scatter_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void scatter(double* out, long long n_slots, int run, int off, long long runs_total, unsigned long long seed) {
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int per_instr = 32 / (run < 32 ? run : 32);            // runs covered by one warp store
    for (long long r = warp * per_instr; r < runs_total; r += n_warps * per_instr) {
        const long long my = r + lane / (run < 32 ? run : 32);
        unsigned long long h = (my + 1) * 0x9E3779B97F4A7C15ull + seed;
        h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
        const long long slot = (long long)(h % (unsigned long long)n_slots);
        for (int k = lane % (run < 32 ? run : 32); k < run; k += 32)
            out[slot * 64 + off + k] = (double)k;                // slots are 512 bytes apart
    }
}
int main() {
    const long long bytes = 13ll << 30;
    double* buf; if (cudaMalloc(&buf, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMemset(buf, 0, bytes);
    const long long n_slots = bytes / 512 - 1;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int runs[] = {4, 8, 16, 32, 64};
    for (int off : {0, 2, 1})
        for (int run : runs) {
            if (off + run > 64) continue;
            const long long total_bytes = 6ll << 30, runs_total = total_bytes / (8 * run);
            scatter<<<148 * 8, 256>>>(buf, n_slots, run, off, runs_total / 16, 1);
            cudaEventRecord(a);
            scatter<<<148 * 8, 256>>>(buf, n_slots, run, off, runs_total, 7);
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            printf("run %3d doubles (%4d B), start offset %d doubles: %.2f ms  %.2f TB/s\n", run, run * 8, off, ms,
                   total_bytes / ms * 1e-9);
        }
    return 0;
}
