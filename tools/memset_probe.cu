// Developer probe: how fast can a 13 GB matrix (the C2 cube varf output) be zero-filled?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/memset_probe tools/memset_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void fill16(double2* p, size_t n2) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) p[i] = make_double2(0.0, 0.0);
}
__global__ void fill32(double4* p, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        asm volatile("st.global.v4.f64 [%0], {%1, %1, %1, %1};" ::"l"(p + i), "d"(0.0) : "memory");
    }
}

int main() {
    const size_t n = 40401, bytes = n * n * 8;
    double* d;
    if (cudaMalloc(&d, bytes + 256) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto report = [&](const char* name, float ms) { printf("%-44s %.3f ms  %.0f GB/s\n", name, ms, bytes / ms * 1e-6); };
    for (int rep = 0; rep < 2; rep++) {
        float ms;
        cudaEventRecord(e0); cudaMemsetAsync(d, 0, bytes, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("cudaMemsetAsync (flat)", ms);
        cudaEventRecord(e0); cudaMemset2DAsync(d, n * 8, 0, n * 8, n, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("cudaMemset2DAsync (width = pitch = 40401*8)", ms);
        for (int mult : {4, 8, 16}) {
            cudaEventRecord(e0); fill16<<<148 * mult, 256>>>((double2*)d, bytes / 16); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1);
            char nm[64]; snprintf(nm, 64, "kernel, 16-byte stores, %d CTAs/SM", mult); report(nm, ms);
        }
        cudaEventRecord(e0); fill32<<<148 * 8, 256>>>((double4*)d, bytes / 32); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); report("kernel, 32-byte stores, 8 CTAs/SM", ms);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
