// tools/fp64_peak.cu -- measures the FP64 ceilings of this GPU: the denominators of the K_FF
// roofline (MEASURED_PEAKS.json has no FP64 entry).
//   1. DFMA issue-bound peak        (ILP x warps sweep)
//   2. DMMA.8x8x4 issue-bound peak  (mma.sync.m8n8k4.f64, the only native FP64 MMA shape on sm_100a)
//   3. cuBLAS DGEMM 8192^3          (the "library" peak quoted in SURVEY 8d)
// Prints one JSON line.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lcublas
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void dfma_kernel(double *out, int iters, double a, double b)
{
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void dmma_kernel(double *out, int iters, double a, double b)
{
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = 0.5 * i; }
    double av = a + threadIdx.x * 1e-9, bv = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(av), "d"(bv));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double time_ms(F f, int reps)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *out;
    CK(cudaMalloc(&out, sizeof(double) * sms * 16 * 1024));
    const int iters = 4096;
    double best_dfma = 0, best_dmma = 0;
    int bd_w = 0, bm_w = 0;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma\": {", prop.name, sms);
    bool first = true;
    for (int warps : {4, 8, 16, 32}) {
        double ms = time_ms([&] { dfma_kernel<8><<<sms * 2, warps * 32 / 2 * 1>>>(out, iters, 1.0000001, 1e-9); }, 5);
        // total warps per SM = 2 blocks * warps/2
        double tf = 2.0 * 8 * iters * (double)(sms * 2) * (warps * 32 / 2) / (ms * 1e-3) / 1e12;
        printf("%s\"w%d\": %.2f", first ? "" : ", ", warps, tf);
        first = false;
        if (tf > best_dfma) { best_dfma = tf; bd_w = warps; }
    }
    printf("}, \"dmma_8x8x4\": {");
    first = true;
    for (int warps : {4, 8, 16, 32}) {
        for (int ilp : {4, 16}) {
            double ms;
            if (ilp == 4) ms = time_ms([&] { dmma_kernel<4><<<sms * 2, warps * 32 / 2>>>(out, iters, 1.0000001, 1e-9); }, 5);
            else ms = time_ms([&] { dmma_kernel<16><<<sms * 2, warps * 32 / 2>>>(out, iters, 1.0000001, 1e-9); }, 5);
            double nwarps = (double)(sms * 2) * (warps / 2);
            double tf = 2.0 * 256 * ilp * iters * nwarps / (ms * 1e-3) / 1e12;
            printf("%s\"w%d_ilp%d\": %.2f", first ? "" : ", ", warps, ilp, tf);
            first = false;
            if (tf > best_dmma) { best_dmma = tf; bm_w = warps; }
        }
    }
    printf("}, ");
    // cuBLAS DGEMM
    cublasHandle_t h;
    cublasCreate(&h);
    const int N = 8192;
    double *A, *B, *C;
    CK(cudaMalloc(&A, sizeof(double) * N * N)); CK(cudaMalloc(&B, sizeof(double) * N * N)); CK(cudaMalloc(&C, sizeof(double) * N * N));
    CK(cudaMemset(A, 0, sizeof(double) * N * N)); CK(cudaMemset(B, 0, sizeof(double) * N * N));
    const double one = 1.0, zero = 0.0;
    double ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, N, N, N, &one, A, N, B, N, &zero, C, N); }, 5);
    double dgemm_tf = 2.0 * N * (double)N * N / (ms * 1e-3) / 1e12;
    // sustained: 20 back-to-back
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; i++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, N, N, N, &one, A, N, B, N, &zero, C, N);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float sms_ms; cudaEventElapsedTime(&sms_ms, e0, e1);
    double dgemm_sust = 20 * 2.0 * N * (double)N * N / (sms_ms * 1e-3) / 1e12;
    printf("\"dfma_peak_tflops\": %.2f, \"dfma_best_warps\": %d, \"dmma_peak_tflops\": %.2f, \"dmma_best_warps\": %d, "
           "\"cublas_dgemm_8192_tflops\": %.2f, \"cublas_dgemm_8192_sustained_tflops\": %.2f}\n",
           best_dfma, bd_w, best_dmma, bm_w, dgemm_tf, dgemm_sust);
    return 0;
}
