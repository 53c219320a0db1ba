// Developer probe: does integer/ALU work on one warp overlap with a DMMA stream on another warp of the same
// SM sub-partition?  148 CTAs x 256 threads: warps 0-3 issue DMMA.8x8x4 on 32 independent accumulators, warps
// 4-7 run an integer loop (mix of dependent IMAD/LOP and a shared-memory load).  Times: DMMA only, ALU only, both.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void __launch_bounds__(256, 1) probe(int mode, int iters, int alu_iters, double* out, int* iout) {
    __shared__ int sm[1024];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = i * 7 + 1;
    __syncthreads();
    if (warp < 4) {
        if (!(mode & 1)) return;
        double acc[32][2];
        for (int i = 0; i < 32; i++) acc[i][0] = acc[i][1] = 0.0;
        double a = 1.0 + lane * 1e-3, b = 2.0 - lane * 1e-3;
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 32; i++) dmma(acc[i][0], acc[i][1], a, b);
        }
        double s = 0;
        for (int i = 0; i < 32; i++) s += acc[i][0] + acc[i][1];
        out[blockIdx.x * 256 + threadIdx.x] = s;
    } else {
        if (!(mode & 2)) return;
        int x = lane + 1, y = warp, z = 3;
        for (int it = 0; it < alu_iters; it++) {
#pragma unroll 8
            for (int i = 0; i < 8; i++) {
                x = x * 1664525 + 1013904223;          // dependent IMAD chain
                y = (y ^ (x >> 7)) + sm[(x >> 5) & 1023];   // LOP + LDS
                z = z * y + (x & 15);
            }
        }
        iout[blockIdx.x * 256 + threadIdx.x] = x ^ y ^ z;
    }
}
int main() {
    double* out; int* iout; cudaMalloc(&out, 148 * 256 * 8); cudaMalloc(&iout, 148 * 256 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 4000, alu = 6000;
    for (int rep = 0; rep < 2; rep++)
        for (int mode = 1; mode <= 3; mode++) {
            cudaEventRecord(a); probe<<<148, 256>>>(mode, iters, alu, out, iout); cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double tf = mode & 1 ? 148.0 * 4 * iters * 32 * 512 / ms * 1e-9 : 0;
            printf("mode %d (%s): %.3f ms  DMMA %.1f TFLOP/s\n", mode, mode == 1 ? "DMMA only" : mode == 2 ? "ALU only" : "both", ms, tf);
        }
    return 0;
}
