// Developer probe (not part of the product library): DMMA / DFMA issue-rate microbenchmarks, cuBLAS
// DGEMM reference rate, and correctness + timing of hb::gemm_launch against cuBLAS.
// Build: see tools/build_probe.sh.  Run on a B200 via gpurun.
#include <cublas_v2.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../hermipy_b200/csrc/hb_internal.h"
#include "../hermipy_b200/csrc/hb_ptx.cuh"

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e = (x);                                                        \
        if (e != cudaSuccess) {                                                     \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                \
        }                                                                           \
    } while (0)

template <int NACC>
__global__ void dmma_rate(double* out, int iters) {
    double c[NACC][2];
    for (int i = 0; i < NACC; i++) c[i][0] = c[i][1] = 0;
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) hb::dmma_8x8x4(c[i][0], c[i][1], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_rate(double* out, int iters) {
    double c[NACC];
    for (int i = 0; i < NACC; i++) c[i] = i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

static void fill(std::vector<double>& v, unsigned seed) {
    unsigned s = seed * 2654435761u + 12345u;
    for (auto& x : v) {
        s = s * 1664525u + 1013904223u;
        x = ((s >> 8) / 16777216.0) - 0.5;
    }
}

struct Case {
    int M, N, K, batch;
    bool am;
    int cfg;
};

int main(int argc, char** argv) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);

    double* scratch;
    CK(cudaMalloc(&scratch, 148 * 8 * 1024 * 8));
    // --- issue-rate microbenchmarks ---
    for (int warps : {4, 8, 16}) {
        const int iters = 4096;
        dmma_rate<16><<<148, warps * 32>>>(scratch, 16);
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0));
        dmma_rate<16><<<148, warps * 32>>>(scratch, iters);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = time_ms(e0, e1);
        double flop = 148.0 * warps * iters * 16 * 512.0;
        printf("DMMA.8x8x4  %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
        CK(cudaEventRecord(e0));
        dfma_rate<16><<<148, warps * 32>>>(scratch, iters);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        ms = time_ms(e0, e1);
        flop = 148.0 * warps * 32 * iters * 16 * 2.0;
        printf("DFMA        %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
    }

    cublasHandle_t hnd;
    cublasCreate(&hnd);

    std::vector<Case> cases = {
        {300, 200, 100, 1, false, 1},  {300, 200, 100, 1, true, 1},   {401, 201, 401, 3, false, 1},
        {515, 333, 77, 2, true, 0},    {4096, 4096, 4096, 1, false, 0}, {4096, 4096, 4096, 1, true, 0},
        {4096, 4096, 4096, 1, false, 1}, {8192, 8192, 8192, 1, false, 0},
        {65536, 1001, 1001, 1, false, 0}, {40401, 40401, 401, 1, true, 0},
    };
    if (argc > 1 && atoi(argv[1]) == 1) cases.resize(4);

    for (const Case& c : cases) {
        const long long lda = c.am ? ((c.M + 1) & ~1) : ((c.K + 1) & ~1);
        const long long ldb = (c.N + 1) & ~1, ldc = (c.N + 1) & ~1;
        const long long rowsA = c.am ? c.K : c.M;
        const long long sA = lda * rowsA, sB = ldb * c.K, sC = ldc * c.M;
        const bool big = (double)sC * c.batch > 3e8;
        std::vector<double> hA(sA * c.batch), hB(sB * c.batch);
        fill(hA, 1);
        fill(hB, 2);
        double *dA, *dB, *dC, *dR;
        CK(cudaMalloc(&dA, hA.size() * 8));
        CK(cudaMalloc(&dB, hB.size() * 8));
        CK(cudaMalloc(&dC, sC * c.batch * 8));
        CK(cudaMalloc(&dR, sC * c.batch * 8));
        CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
        CK(cudaMemset(dC, 0xff, sC * c.batch * 8));
        CK(cudaMemset(dR, 0, sC * c.batch * 8));

        hb::GemmArgs a;
        a.M = c.M; a.N = c.N; a.K = c.K; a.batch = c.batch;
        a.A = dA; a.lda = lda; a.strideA = sA; a.a_mmajor = c.am;
        a.B = dB; a.ldb = ldb; a.strideB = sB;
        a.C = dC; a.ldc = ldc; a.strideC = sC;
        a.cfg = c.cfg;
        int rc = hb::gemm_launch(a, 0);
        cudaError_t err = cudaDeviceSynchronize();
        if (rc || err != cudaSuccess) {
            printf("case M=%d N=%d K=%d: launch rc=%d err=%s\n", c.M, c.N, c.K, rc, cudaGetErrorString(err));
            return 1;
        }
        // cuBLAS (column-major): C^T (N x M) = B^T (N x K) * A^T (K x M)
        const double one = 1, zero = 0;
        auto run_cublas = [&]() {
            cublasDgemmStridedBatched(hnd, CUBLAS_OP_N, c.am ? CUBLAS_OP_T : CUBLAS_OP_N, c.N, c.M, c.K, &one,
                                      dB, (int)ldb, sB, dA, (int)lda, sA, &zero, dR, (int)ldc, sC, c.batch);
        };
        run_cublas();
        CK(cudaDeviceSynchronize());
        // compare
        double maxerr = 0, maxref = 0;
        {
            const long long rows_chk = big ? 64 : c.M;
            std::vector<double> hC(ldc * rows_chk), hR(ldc * rows_chk);
            for (int b = 0; b < c.batch; b++) {
                for (int part = 0; part < (big ? 2 : 1); part++) {
                    const long long r0 = part == 0 ? 0 : c.M - rows_chk;
                    CK(cudaMemcpy(hC.data(), dC + b * sC + r0 * ldc, hC.size() * 8, cudaMemcpyDeviceToHost));
                    CK(cudaMemcpy(hR.data(), dR + b * sC + r0 * ldc, hR.size() * 8, cudaMemcpyDeviceToHost));
                    for (long long r = 0; r < rows_chk; r++)
                        for (int j = 0; j < c.N; j++) {
                            double d = fabs(hC[r * ldc + j] - hR[r * ldc + j]);
                            if (!(d <= maxerr)) maxerr = d;
                            if (fabs(hR[r * ldc + j]) > maxref) maxref = fabs(hR[r * ldc + j]);
                        }
                }
            }
        }
        // timing
        const int reps = ((double)c.M * c.N * c.K * c.batch > 1e11) ? 3 : 10;
        hb::gemm_launch(a, 0);
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) hb::gemm_launch(a, 0);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        const float ms_hb = time_ms(e0, e1) / reps;
        run_cublas();
        CK(cudaEventRecord(e0));
        for (int i = 0; i < reps; i++) run_cublas();
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        const float ms_cb = time_ms(e0, e1) / reps;
        const double flop = 2.0 * c.M * c.N * c.K * c.batch;
        printf("M=%6d N=%6d K=%5d b=%d am=%d cfg=%d: maxerr %.3e (ref max %.3e)  hb %.3f ms %.2f TF | cublas %.3f ms %.2f TF\n",
               c.M, c.N, c.K, c.batch, (int)c.am, c.cfg, maxerr, maxref, ms_hb, flop / ms_hb * 1e-9, ms_cb,
               flop / ms_cb * 1e-9);
        fflush(stdout);
        cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dR);
    }
    return 0;
}
