// Developer probe: pinned host <-> device copy rates, 1-D vs pitched 2-D (row 3208 B -> pitch 3216 B), and both
// directions at once.  nvcc -O2 -o tools/bin/pcie_probe tools/pcie_probe.cu ; run via gpurun.
#include <cstdio>
#include <cuda_runtime.h>
static float timeit(cudaStream_t s, cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b); return ms; }
int main() {
    const size_t rows = 401 * 17, n = 401, np = 402;
    const size_t bytes = rows * n * 8;
    double *h, *h2, *d, *d2;
    cudaMallocHost(&h, bytes); cudaMallocHost(&h2, bytes);
    cudaMalloc(&d, rows * np * 8); cudaMalloc(&d2, rows * np * 8);
    cudaStream_t s1, s2; cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(a, s1); cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s1); cudaEventRecord(b, s1);
        float t = timeit(s1, a, b); printf("H2D 1-D   %.1f MB  %.3f ms  %.1f GB/s\n", bytes * 1e-6, t, bytes / t * 1e-6);
        cudaEventRecord(a, s1); cudaMemcpy2DAsync(d, np * 8, h, n * 8, n * 8, rows, cudaMemcpyHostToDevice, s1); cudaEventRecord(b, s1);
        t = timeit(s1, a, b); printf("H2D 2-D   %.1f MB  %.3f ms  %.1f GB/s\n", bytes * 1e-6, t, bytes / t * 1e-6);
        cudaEventRecord(a, s1); cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s1); cudaEventRecord(b, s1);
        t = timeit(s1, a, b); printf("D2H 1-D   %.1f MB  %.3f ms  %.1f GB/s\n", bytes * 1e-6, t, bytes / t * 1e-6);
        cudaEventRecord(a, s1); cudaMemcpy2DAsync(h, n * 8, d, np * 8, n * 8, rows, cudaMemcpyDeviceToHost, s1); cudaEventRecord(b, s1);
        t = timeit(s1, a, b); printf("D2H 2-D   %.1f MB  %.3f ms  %.1f GB/s\n", bytes * 1e-6, t, bytes / t * 1e-6);
        cudaDeviceSynchronize();
        cudaEventRecord(a, s1);
        cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s1);
        cudaMemcpyAsync(h2, d2, bytes, cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s2);
        cudaEventRecord(b, s1);
        t = timeit(s1, a, b); printf("both 1-D  2 x %.1f MB  %.3f ms  %.1f GB/s total\n", bytes * 1e-6, t, 2 * bytes / t * 1e-6);
    }
    return 0;
}
