// gp.cu -- GP algebra that consumes the covariance blocks on the device
// (reference cspbo/gaussianprocess.py:105-127 fit, :138-141 / :359-366 prediction).
//   noise add + K*.alpha GEMV : hand-written HBM-bound kernels
//   Cholesky factorisation / solve : cuSOLVER potrf / potrs (as BASELINE.json's north_star asks)
#include <cublas_v2.h>
#include <cuda.h>
#include <cusolverDn.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.h"

namespace cspbo {

__global__ void add_noise_kernel(double *K, int N, int n_energy, double ne2, double nf2)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) K[(size_t)i * N + i] += (i < n_energy) ? ne2 : nf2;
}

// pred[r] = sum_j Kstar[r, j] alpha[j].  HBM-bound: every K* byte is read exactly once, 2 flop per 8 bytes.
// One warp owns one (row, column chunk) piece of CHUNK columns and keeps UNROLL 128-bit streaming
// loads per lane in flight (Little: ~6.5 MB must be in flight chip-wide); chunks give a row of a small-n
// K* (one MD step: n = 577) enough warps to cover all 148 SMs.  Partial sums go to partial[chunk][row]
// and are added in chunk order by gemv_reduce_kernel, so the result is deterministic (no atomics).
constexpr int kGemvWarps = 8;

template <int UNROLL, int CHUNK>
__global__ void __launch_bounds__(kGemvWarps * 32)
predict_gemv_kernel(const double *__restrict__ Ks, int n, int N, const double *__restrict__ alpha,
                    double *__restrict__ out, int nchunks)
{
    const int lane = threadIdx.x & 31;
    const long long piece = (long long)blockIdx.x * kGemvWarps + (threadIdx.x >> 5);
    if (piece >= (long long)n * nchunks) return;
    // consecutive warps take consecutive chunks of the same row: a CTA streams one contiguous run
    const int row = (int)(piece / nchunks), chunk = (int)(piece - (long long)row * nchunks);
    const int c0 = chunk * CHUNK;
    const int len = min(CHUNK, N - c0);
    const double *kr = Ks + (long long)row * N + c0;
    const double *ar = alpha + c0;
    double acc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) acc[u] = 0.0;
    // peel to 16-byte alignment of the K* row piece (rows of an odd-N matrix alternate)
    const int head = (int)((reinterpret_cast<size_t>(kr) >> 3) & 1) && len > 0 ? 1 : 0;
    if (head && lane == 0) acc[0] = kr[0] * __ldg(ar);
    const int body = (len - head) >> 1;   // double2 elements
    const double2 *k2 = reinterpret_cast<const double2 *>(kr + head);
    const double *a1 = ar + head;
    auto ldk = [](const double2 *q) { return __ldcs(q); };   // streaming: every K* byte is used once
    int j = lane;
    for (; j + (UNROLL - 1) * 32 < body; j += UNROLL * 32) {
        double2 kv[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) kv[u] = ldk(k2 + j + u * 32);
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            acc[u] = fma(kv[u].x, __ldg(a1 + 2 * (j + u * 32)), acc[u]);
            acc[u] = fma(kv[u].y, __ldg(a1 + 2 * (j + u * 32) + 1), acc[u]);
        }
    }
    for (; j < body; j += 32) {
        const double2 kv = ldk(k2 + j);
        acc[0] = fma(kv.x, __ldg(a1 + 2 * j), acc[0]);
        acc[0] = fma(kv.y, __ldg(a1 + 2 * j + 1), acc[0]);
    }
    if (((len - head) & 1) && lane == 0) acc[1] = fma(kr[len - 1], __ldg(ar + len - 1), acc[1]);
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < UNROLL; u++) s += acc[u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[(long long)chunk * n + row] = s;
}

__global__ void gemv_reduce_kernel(const double *__restrict__ partial, int n, int nchunks, double *__restrict__ pred)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n) return;
    double s = 0.0;
    for (int c = 0; c < nchunks; c++) s += partial[(long long)c * n + row];
    pred[row] = s;
}

__global__ void logdet_kernel(const double *L, int N, double *out)
{
    __shared__ double sh[32];
    double acc = 0.0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) acc += log(L[(size_t)i * N + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) *out = 2.0 * acc;
    }
}

// copy the column-major lower triangle (= row-major upper triangle) onto the other one
__global__ void symmetrize_kernel(double *A, int N)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long i = idx / N, j = idx % N;
    if (i < N && j > i) A[j * N + i] = A[i * N + j];   // row-major: lower(j,i) <- upper(i,j)
}

static std::mutex g_solver_mu;
static cusolverDnHandle_t g_solver[64] = {nullptr};

static int solver_handle(cusolverDnHandle_t *h)
{
    int dev = 0;
    CSPBO_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_solver_mu);
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return CSPBO_E_ARG; }
    if (!g_solver[dev]) {
        cusolverStatus_t s = cusolverDnCreate(&g_solver[dev]);
        if (s != CUSOLVER_STATUS_SUCCESS) { set_error("cusolverDnCreate failed (%d)", (int)s); return CSPBO_E_SOLVER; }
    }
    *h = g_solver[dev];
    return CSPBO_OK;
}

}  // namespace cspbo

using namespace cspbo;

extern "C" int cspbo_gp_add_noise(double *K, int N, int n_energy, double noise_e, double f_coef)
{
    if (!K || N < 0) { set_error("gp_add_noise: bad argument"); return CSPBO_E_ARG; }
    if (N == 0) return CSPBO_OK;
    const double nf = f_coef * noise_e;
    add_noise_kernel<<<(N + 255) / 256, 256>>>(K, N, n_energy, noise_e * noise_e, nf * nf);
    g_launches++;
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}

namespace cspbo {
// Workspace of the single-GPU factorisation: cached blocks (dev_alloc_cached: no cudaMalloc / cudaFree per call once
// warm) and ONE readback of the two info words + log-determinant at the end instead of a blocking copy per step.
struct CholScratch {
    double *work = nullptr;
    double *small = nullptr;     // [0] logdet, then 2 ints: potrf info, potrs info
    ~CholScratch() { dev_free_cached(work); dev_free_cached(small); }
};

// diag_out[i] = sum_{j >= i} A[i*N + j]^2: with A = L^-1 (column-major lower == row-major upper) this is (K^-1)_ii
__global__ void row_sumsq_upper_kernel(const double *__restrict__ A, int N, double *__restrict__ out)
{
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= N) return;
    const double *r = A + (size_t)row * N;
    double acc = 0.0;
    for (int j = row + lane; j < N; j += 32) acc = fma(r[j], r[j], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[row] = acc;
}
}  // namespace cspbo

extern "C" int cspbo_gp_cholesky_solve(double *K, int N, double *y, int nrhs, double *logdet)
{
    if (!K || N <= 0 || (nrhs > 0 && !y)) { set_error("gp_cholesky_solve: bad argument"); return CSPBO_E_ARG; }
    cusolverDnHandle_t h;
    int rc = solver_handle(&h);
    if (rc) return rc;
    int lwork = 0;
    // K is symmetric, so its row-major image is also its column-major image.
    const cublasFillMode_t uplo = CUBLAS_FILL_MODE_LOWER;
    cusolverStatus_t s = cusolverDnDpotrf_bufferSize(h, uplo, N, K, N, &lwork);
    if (s != CUSOLVER_STATUS_SUCCESS) { set_error("potrf_bufferSize failed (%d)", (int)s); return CSPBO_E_SOLVER; }
    CholScratch sc;
    if ((rc = dev_alloc_cached((size_t)std::max(lwork, 1) * sizeof(double), (void **)&sc.work)) != CSPBO_OK) return rc;
    if ((rc = dev_alloc_cached(32, (void **)&sc.small)) != CSPBO_OK) return rc;
    int *info = reinterpret_cast<int *>(sc.small + 1);
    CSPBO_CUDA(cudaMemsetAsync(sc.small, 0, 32));
    s = cusolverDnDpotrf(h, uplo, N, K, N, sc.work, lwork, info);
    if (s != CUSOLVER_STATUS_SUCCESS) { cudaDeviceSynchronize(); set_error("potrf failed (status %d)", (int)s); return CSPBO_E_SOLVER; }
    // a failed factorisation leaves garbage behind its leading minor; the solve below is then meaningless but harmless,
    // and the info word read back once at the end decides
    if (nrhs > 0) {
        s = cusolverDnDpotrs(h, uplo, N, nrhs, K, N, y, N, info + 1);
        if (s != CUSOLVER_STATUS_SUCCESS) { cudaDeviceSynchronize(); set_error("potrs failed (status %d)", (int)s); return CSPBO_E_SOLVER; }
    }
    if (logdet) {
        logdet_kernel<<<1, 1024>>>(K, N, sc.small);
        g_launches++;
    }
    struct { double logdet; int info[2]; int pad[4]; } host;
    static_assert(sizeof(host) == 32, "readback layout");
    CSPBO_CUDA(cudaMemcpy(&host, sc.small, 32, cudaMemcpyDeviceToHost));     // the one synchronisation of the call
    if (host.info[0] > 0) { set_error("K is not positive definite (leading minor %d)", host.info[0]); return CSPBO_E_NOTPD; }
    if (host.info[0] != 0 || host.info[1] != 0) { set_error("potrf / potrs failed (info %d, %d)", host.info[0], host.info[1]); return CSPBO_E_SOLVER; }
    if (logdet) *logdet = host.logdet;
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}

extern "C" int cspbo_gp_predict(const double *Kstar, int n, int N, const double *alpha, double *pred)
{
    if (n < 0 || N < 0 || (n > 0 && (!Kstar || !pred)) || (N > 0 && !alpha)) { set_error("gp_predict: bad argument"); return CSPBO_E_ARG; }
    if (n == 0) return CSPBO_OK;
    // 8 x 128-bit loads in flight per lane; 4096-column chunks unless that leaves the 148 SMs short of warps.
    // Measured on B200 (tools/gemv_probe.py): 4 loads/2048 -> 4.6 TB/s, 8/2048 -> 5.6, 8/4096 -> 5.75 (20001^2), 6.5 TB/s at 40002^2
    typedef void (*kern_t)(const double *, int, int, const double *, double *, int);
    const bool small = (long long)n * ((N + 4095) / 4096) < 148LL * 8 * 4;
    kern_t kern = small ? predict_gemv_kernel<8, 2048> : predict_gemv_kernel<8, 4096>;
    const int chunk = small ? 2048 : 4096;
    const int nchunks = (N + chunk - 1) / chunk;
    const long long pieces = (long long)n * nchunks;
    const unsigned grid = (unsigned)((pieces + kGemvWarps - 1) / kGemvWarps);
    if (nchunks <= 1) {
        kern<<<grid, kGemvWarps * 32>>>(Kstar, n, N, alpha, pred, 1);
    } else {
        void *partial = nullptr;
        int rc = scratch_get(SCRATCH_GEMV, (size_t)pieces * sizeof(double), &partial);
        if (rc) return rc;
        kern<<<grid, kGemvWarps * 32>>>(Kstar, n, N, alpha, static_cast<double *>(partial), nchunks);
        g_launches++;
        gemv_reduce_kernel<<<(n + 255) / 256, 256>>>(static_cast<const double *>(partial), n, nchunks, pred);
    }
    g_launches++;
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}


extern "C" int cspbo_gp_cho_solve(const double *L, int N, double *B, int nrhs)
{
    if (!L || !B || N <= 0 || nrhs < 0) { set_error("gp_cho_solve: bad argument"); return CSPBO_E_ARG; }
    if (nrhs == 0) return CSPBO_OK;
    cusolverDnHandle_t h;
    int rc = solver_handle(&h);
    if (rc) return rc;
    CholScratch sc;
    if ((rc = dev_alloc_cached(32, (void **)&sc.small)) != CSPBO_OK) return rc;
    int *info = reinterpret_cast<int *>(sc.small), hinfo = 0;
    cusolverStatus_t s = cusolverDnDpotrs(h, CUBLAS_FILL_MODE_LOWER, N, nrhs, L, N, B, N, info);
    cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
    if (s != CUSOLVER_STATUS_SUCCESS || hinfo != 0) { set_error("potrs failed (status %d, info %d)", (int)s, hinfo); return CSPBO_E_SOLVER; }
    return CSPBO_OK;
}

namespace cspbo {
int dense_factor_inverse(double *L, int N, long long ld, int what);      // dense.cu
long long dense_workspace_bytes(int N);
// cuSOLVER below CSPBO_DENSE_MIN rows (default 1024: the recursion's launches are not worth it) or with CSPBO_DENSE=0
static bool use_dense(int N)
{
    const int on = getenv("CSPBO_DENSE") ? atoi(getenv("CSPBO_DENSE")) : 1;             // read per call: tests switch it
    const int nmin = getenv("CSPBO_DENSE_MIN") ? atoi(getenv("CSPBO_DENSE_MIN")) : 1024;
    return on != 0 && N >= nmin;
}
}  // namespace cspbo

extern "C" long long cspbo_gp_inverse_workspace_bytes(int N)
{
    return (N > 0 && use_dense(N)) ? dense_workspace_bytes(N) : 0;
}

// In place: the factor L (column-major lower == row-major upper triangle of the buffer) -> K^-1, same triangle
// (cuSOLVER potri).  With mirror != 0 the other triangle is filled too.
extern "C" int cspbo_gp_cholesky_inverse_inplace(double *L, int N, int mirror)
{
    if (!L || N <= 0) { set_error("gp_cholesky_inverse_inplace: bad argument"); return CSPBO_E_ARG; }
    if (use_dense(N)) {      // recursive dgemm-rich sweeps (dense.cu): 2-3x cuSOLVER's potri at N = 20 001
        const int rcd = dense_factor_inverse(L, N, N, 1);
        if (rcd) return rcd;
        if (mirror) {
            const long long total = (long long)N * N;
            symmetrize_kernel<<<(unsigned)((total + 255) / 256), 256>>>(L, N);
            g_launches++;
        }
        CSPBO_CUDA(cudaGetLastError());
        return CSPBO_OK;
    }
    cusolverDnHandle_t h;
    int rc = solver_handle(&h);
    if (rc) return rc;
    int lwork = 0;
    cusolverStatus_t s = cusolverDnDpotri_bufferSize(h, CUBLAS_FILL_MODE_LOWER, N, L, N, &lwork);
    if (s != CUSOLVER_STATUS_SUCCESS) { set_error("potri_bufferSize failed (%d)", (int)s); return CSPBO_E_SOLVER; }
    CholScratch sc;
    if ((rc = dev_alloc_cached((size_t)std::max(lwork, 1) * sizeof(double), (void **)&sc.work)) != CSPBO_OK) return rc;
    if ((rc = dev_alloc_cached(32, (void **)&sc.small)) != CSPBO_OK) return rc;
    int *info = reinterpret_cast<int *>(sc.small), hinfo = 0;
    s = cusolverDnDpotri(h, CUBLAS_FILL_MODE_LOWER, N, L, N, sc.work, lwork, info);
    if (mirror) {
        const long long total = (long long)N * N;
        symmetrize_kernel<<<(unsigned)((total + 255) / 256), 256>>>(L, N);
        g_launches++;
    }
    cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
    if (s != CUSOLVER_STATUS_SUCCESS || hinfo != 0) { set_error("potri failed (status %d, info %d)", (int)s, hinfo); return CSPBO_E_SOLVER; }
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}

extern "C" int cspbo_gp_cholesky_inverse(const double *L, int N, double *Kinv)
{
    if (!L || !Kinv || N <= 0) { set_error("gp_cholesky_inverse: bad argument"); return CSPBO_E_ARG; }
    CSPBO_CUDA(cudaMemcpyAsync(Kinv, L, (size_t)N * N * sizeof(double), cudaMemcpyDeviceToDevice));
    return cspbo_gp_cholesky_inverse_inplace(Kinv, N, 1);
}

// diag_out[N] = diag((L L^T)^-1), destroying L: the factor is inverted in place (cuSOLVER trtri, N^3/3 flop, no second
// N x N array) and (K^-1)_ii = sum_j (L^-1)_ji^2 is a row-wise sum of squares of what is left.  This is all the Dot
// kernel's hyper-gradient needs of K^-1 (dK/dsigma = 2K/sigma, dK/dnoise diagonal; gaussianprocess.py:490-499).
namespace cspbo {
// L -> L^-1 in place (column-major lower == row-major upper triangle of the buffer): dense.cu from 1024 rows on, cuSOLVER trtri below
static int tri_inverse_inplace(double *L, int N)
{
    if (use_dense(N)) return dense_factor_inverse(L, N, N, 0);
    cusolverDnHandle_t h;
    int rc = solver_handle(&h);
    if (rc) return rc;
    size_t wdev = 0, whost = 0;
    cusolverStatus_t s = cusolverDnXtrtri_bufferSize(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_DIAG_NON_UNIT, (int64_t)N, CUDA_R_64F, L, (int64_t)N,
                                                     &wdev, &whost);
    if (s != CUSOLVER_STATUS_SUCCESS) { set_error("trtri_bufferSize failed (%d)", (int)s); return CSPBO_E_SOLVER; }
    CholScratch sc;
    if ((rc = dev_alloc_cached(std::max<size_t>(wdev, 8), (void **)&sc.work)) != CSPBO_OK) return rc;
    if ((rc = dev_alloc_cached(32, (void **)&sc.small)) != CSPBO_OK) return rc;
    std::vector<char> hbuf(std::max<size_t>(whost, 1));
    int *info = reinterpret_cast<int *>(sc.small), hinfo = 0;
    s = cusolverDnXtrtri(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_DIAG_NON_UNIT, (int64_t)N, CUDA_R_64F, L, (int64_t)N, sc.work, wdev, hbuf.data(),
                         whost, info);
    cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
    if (s != CUSOLVER_STATUS_SUCCESS || hinfo != 0) { set_error("trtri failed (status %d, info %d)", (int)s, hinfo); return CSPBO_E_SOLVER; }
    return CSPBO_OK;
}
}  // namespace cspbo

extern "C" int cspbo_gp_tri_inverse_inplace(double *L, int N)
{
    if (!L || N <= 0) { set_error("gp_tri_inverse_inplace: bad argument"); return CSPBO_E_ARG; }
    return tri_inverse_inplace(L, N);
}

extern "C" int cspbo_gp_inverse_diag(double *L, int N, double *diag_out)
{
    if (!L || !diag_out || N <= 0) { set_error("gp_inverse_diag: bad argument"); return CSPBO_E_ARG; }
    const int rc = tri_inverse_inplace(L, N);
    if (rc) return rc;
    row_sumsq_upper_kernel<<<(unsigned)((N + 7) / 8), 256>>>(L, N, diag_out);
    g_launches++;
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}


// ------------------------------------------------------------------------------------------
// Panel operations of the sharded (one process per GPU) Cholesky: K = U^T U, rows of U live where the
// rows of K live (csp_bo_b200/distributed.py: ShardedCholesky).  Row-major addressing; every call is
// asynchronous on the caller's stream, positive-definiteness is reported through a device-side info word
// so that the panel pipeline never synchronises with the host.
// ------------------------------------------------------------------------------------------
namespace cspbo {

struct StreamHandles {
    cudaStream_t stream;
    int device;
    cublasHandle_t blas;
    cusolverDnHandle_t solver;
    double *work;
    int lwork;
};

int chol_diag_launch(double *D, long long ld, int h, int *info_dev, cudaStream_t st);
#ifndef CSPBO_CHOL_FUSED_DEFAULT
#define CSPBO_CHOL_FUSED_DEFAULT false
#endif

// inv = L^-1 for the h x h Cholesky factor that potrf left in D (row-major upper U == column-major lower L,
// L(i,k) = D[k*ld + i]); inv is column-major lower with leading dimension h, the operand of the wide trmm.
// One CTA = 8 warps = 8 consecutive columns j of L^-1, one warp per column: forward substitution of L x = e_j with the
// right-hand side in REGISTERS (lane owns the rows i = lane + 32u), x_k broadcast by one shuffle per step.  Step k needs
// column k of L = row k of D: the CTA stages 8 rows at a time in shared memory (one row per warp, coalesced, fetched
// one chunk ahead into registers so that the L2 latency hides behind the current chunk's steps).  A step's critical
// path is shuffle + multiply + one FMA.  The diagonal chain of the sharded Cholesky is latency bound: cuBLAS' trsm
// against the identity took 0.2 ms for h = 384, a first version with the right-hand side in shared memory 0.14 ms.
constexpr int kInvWarps = 8;

template <int KPER>      // rows per lane: h <= 32 * KPER
__global__ void __launch_bounds__(kInvWarps * 32, 2) tri_inverse_kernel(const double *__restrict__ D, long long ld, int h, double *__restrict__ inv, int ntot)
{
    extern __shared__ double tsm[];          // [h] reciprocal diagonal | [8][h] staged rows of D
    // batched form (gridDim.y leaves): leaf y is the h x h diagonal block y of an ntot x ntot factor, the last one ragged
    D += (long long)blockIdx.y * h * (ld + 1);
    inv += (long long)blockIdx.y * h * h;
    h = min(h, ntot - (int)blockIdx.y * h);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double *rdiag = tsm;
    double *rows = tsm + h;
    const int j0 = blockIdx.x * kInvWarps, j = j0 + w;
    for (int i = tid; i < h; i += blockDim.x) rdiag[i] = 1.0 / D[(long long)i * ld + i];

    double bv[KPER], stage[KPER];
#pragma unroll
    for (int u = 0; u < KPER; u++) bv[u] = (lane + 32 * u == j) ? 1.0 : 0.0;
    auto fetch = [&](int kc) {               // warp w brings row kc + w of D (columns right of the diagonal) into registers
        const int k = kc + w;
        const double *src = D + (long long)k * ld;
#pragma unroll
        for (int u = 0; u < KPER; u++) {
            const int i = lane + 32 * u;
            stage[u] = (k < h && i < h && i > k) ? src[i] : 0.0;
        }
    };
    auto commit = [&]() {
#pragma unroll
        for (int u = 0; u < KPER; u++) {
            const int i = lane + 32 * u;
            if (i < h) rows[(size_t)w * h + i] = stage[u];
        }
    };
    fetch(j0);
#pragma unroll
    for (int s = 0; s < KPER; s++) {
        if (32 * s >= h || 32 * (s + 1) <= j0) continue;          // uniform over the CTA
        for (int kk = 0; kk < 32; kk++) {
            const int k = 32 * s + kk;
            if (k < j0 || k >= h) continue;                        // uniform over the CTA (j0 is a multiple of 8)
            if ((k & 7) == 0) {
                __syncthreads();             // the previous chunk has been consumed (first time: rdiag is set up)
                commit();
                __syncthreads();
                if (k + kInvWarps < h) fetch(k + kInvWarps);
            }
            const double xk = __shfl_sync(0xffffffffu, bv[s], kk) * rdiag[k];
            if (k >= j && j < h) {                                 // uniform over the warp
                const double *col = rows + (size_t)(k & 7) * h;    // L(i,k) = col[i]
#pragma unroll
                for (int u = s; u < KPER; u++) {
                    const int i = lane + 32 * u;
                    if (i > k && i < h) bv[u] = fma(-col[i], xk, bv[u]);
                }
                if (lane == kk) bv[s] = xk;
            }
        }
    }
    if (j < h) {
#pragma unroll
        for (int u = 0; u < KPER; u++) {
            const int i = lane + 32 * u;
            if (i < h) inv[(long long)j * h + i] = (i >= j) ? bv[u] : 0.0;
        }
    }
}

// inv[y][h_y * h_y] = inverse of the y-th nb x nb diagonal block of the N x N column-major lower factor L (the last block
// is ragged; column-major lower with leading dimension h_y, zeros above the diagonal): the leaves of dense.cu's
// recursive triangular inverse, all in one launch.
int tri_inverse_leaves(const double *L, long long ld, int N, int nb, double *inv, cudaStream_t st)
{
    if (nb < 1 || nb > 512) { set_error("tri_inverse_leaves: leaf of %d rows (max 512)", nb); return CSPBO_E_ARG; }
    const size_t smem = (size_t)(1 + kInvWarps) * nb * sizeof(double);
    const dim3 grid((unsigned)((nb + kInvWarps - 1) / kInvWarps), (unsigned)((N + nb - 1) / nb));
    if (nb <= 192) tri_inverse_kernel<6><<<grid, kInvWarps * 32, smem, st>>>(L, ld, nb, inv, N);
    else if (nb <= 384) tri_inverse_kernel<12><<<grid, kInvWarps * 32, smem, st>>>(L, ld, nb, inv, N);
    else tri_inverse_kernel<16><<<grid, kInvWarps * 32, smem, st>>>(L, ld, nb, inv, N);
    g_launches++;
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}

static StreamHandles g_sh[16];
static int g_nsh = 0;

static int stream_handles(cudaStream_t st, StreamHandles **out)
{
    int dev = 0;
    CSPBO_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_solver_mu);
    for (int i = 0; i < g_nsh; i++)
        if (g_sh[i].stream == st && g_sh[i].device == dev) { *out = &g_sh[i]; return CSPBO_OK; }
    if (g_nsh >= 16) {   // evict the oldest entry (its stream is most likely gone with its owner)
        cublasDestroy(g_sh[0].blas);
        cusolverDnDestroy(g_sh[0].solver);
        cudaFree(g_sh[0].work);
        for (int i = 1; i < g_nsh; i++) g_sh[i - 1] = g_sh[i];
        g_nsh--;
    }
    StreamHandles h = {st, dev, nullptr, nullptr, nullptr, 0};
    if (cublasCreate(&h.blas) != CUBLAS_STATUS_SUCCESS) { set_error("cublasCreate failed"); return CSPBO_E_SOLVER; }
    if (cusolverDnCreate(&h.solver) != CUSOLVER_STATUS_SUCCESS) { cublasDestroy(h.blas); set_error("cusolverDnCreate failed"); return CSPBO_E_SOLVER; }
    cublasSetStream(h.blas, st);
    cusolverDnSetStream(h.solver, st);
    g_sh[g_nsh] = h;
    *out = &g_sh[g_nsh++];
    return CSPBO_OK;
}

}  // namespace cspbo

// D: h x h symmetric diagonal block (row-major, leading dimension ld) -> its factor in place (row-major upper ==
// column-major lower, cuSOLVER potrf) and inv[h*h] = the inverse of that factor (trsm against the identity).
extern "C" int cspbo_gp_chol_diag(double *D, long long ld, int h, double *inv, int *info_dev, void *stream)
{
    if (!D || !inv || !info_dev || h <= 0 || ld < h) { set_error("gp_chol_diag: bad argument"); return CSPBO_E_ARG; }
    StreamHandles *sh;
    int rc = stream_handles(static_cast<cudaStream_t>(stream), &sh);
    if (rc) return rc;
    // CSPBO_CHOL_DIAG: "fused" = the cluster-resident single-kernel factorisation of chol.cu (h <= 384; falls back to
    // cuSOLVER when the stream's SM partition does not take an 8-CTA cluster), "cusolver" = cuSOLVER potrf.
    static const bool use_fused = getenv("CSPBO_CHOL_DIAG") ? !strcmp(getenv("CSPBO_CHOL_DIAG"), "fused") : CSPBO_CHOL_FUSED_DEFAULT;
    static bool fused_ok = true;
    bool done = false;
    if (use_fused && fused_ok && h <= 384) {
        rc = chol_diag_launch(D, ld, h, info_dev, sh->stream);
        if (rc == CSPBO_OK) done = true;
        else if (rc == CSPBO_E_CUDA) fused_ok = false;      // no cluster on this partition: cuSOLVER from now on
        else return rc;
    }
    if (!done) {
        int lwork = 0;
        cusolverStatus_t s = cusolverDnDpotrf_bufferSize(sh->solver, CUBLAS_FILL_MODE_LOWER, h, D, (int)ld, &lwork);
        if (s != CUSOLVER_STATUS_SUCCESS) { set_error("potrf_bufferSize failed (%d)", (int)s); return CSPBO_E_SOLVER; }
        if (lwork > sh->lwork) {
            CSPBO_CUDA(cudaStreamSynchronize(sh->stream));
            cudaFree(sh->work); sh->work = nullptr; sh->lwork = 0;
            CSPBO_CUDA(cudaMalloc(&sh->work, (size_t)lwork * sizeof(double)));
            sh->lwork = lwork;
        }
        s = cusolverDnDpotrf(sh->solver, CUBLAS_FILL_MODE_LOWER, h, D, (int)ld, sh->work, sh->lwork, info_dev);
        if (s != CUSOLVER_STATUS_SUCCESS) { set_error("potrf failed (status %d)", (int)s); return CSPBO_E_SOLVER; }
    }
    if (h > 512) { set_error("gp_chol_diag: diagonal block of %d rows is too large (max 512)", h); return CSPBO_E_ARG; }
    const size_t smem = (size_t)(1 + kInvWarps) * h * sizeof(double);     // <= 36 KB
    const unsigned grid = (unsigned)((h + kInvWarps - 1) / kInvWarps);
    if (h <= 192) tri_inverse_kernel<6><<<grid, kInvWarps * 32, smem, sh->stream>>>(D, ld, h, inv, h);
    else if (h <= 384) tri_inverse_kernel<12><<<grid, kInvWarps * 32, smem, sh->stream>>>(D, ld, h, inv, h);
    else tri_inverse_kernel<16><<<grid, kInvWarps * 32, smem, sh->stream>>>(D, ld, h, inv, h);
    CSPBO_CUDA(cudaGetLastError());
    g_launches += 2;
    return CSPBO_OK;
}

// dst[h, w] = U_ss^-T src[h, w] (row-major views with leading dimensions ldb / ldc): the panel rows right of the
// diagonal block, through the explicit inverse (one wide, tensor-pipe bound trmm instead of a latency-bound trsm).
extern "C" int cspbo_gp_chol_apply_inv(const double *inv, int h, const double *src, long long ldb, long long w, double *dst,
                                       long long ldc, void *stream)
{
    if (!inv || !src || !dst || h <= 0 || w < 0 || ldb < w || ldc < w) { set_error("gp_chol_apply_inv: bad argument"); return CSPBO_E_ARG; }
    if (w == 0) return CSPBO_OK;
    StreamHandles *sh;
    int rc = stream_handles(static_cast<cudaStream_t>(stream), &sh);
    if (rc) return rc;
    const double one = 1.0;   // column-major: DST (w x h) = SRC (w x h) . (L_ss^-1)^T
    cublasStatus_t b = cublasDtrmm(sh->blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, (int)w, h, &one,
                                   inv, h, src, (int)ldb, dst, (int)ldc);
    if (b != CUBLAS_STATUS_SUCCESS) { set_error("trmm failed (status %d)", (int)b); return CSPBO_E_SOLVER; }
    g_launches++;
    return CSPBO_OK;
}

// C[nb, w] -= A[h, nb]^T . B[h, w]  (row-major views): every trailing update of the factorisation.
extern "C" int cspbo_gp_chol_update(double *C, long long ldc, int nb, long long w, const double *A, long long lda, const double *B,
                                    long long ldb, int h, void *stream)
{
    if (!C || !A || !B || h <= 0 || nb < 0 || w < 0 || ldc < w || lda < nb || ldb < w) { set_error("gp_chol_update: bad argument"); return CSPBO_E_ARG; }
    if (nb == 0 || w == 0) return CSPBO_OK;
    StreamHandles *sh;
    int rc = stream_handles(static_cast<cudaStream_t>(stream), &sh);
    if (rc) return rc;
    const double minus = -1.0, one = 1.0;   // column-major: C (w x nb) -= B (w x h) . A (nb x h)^T
    cublasStatus_t b = cublasDgemm(sh->blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)w, nb, h, &minus, B, (int)ldb, A, (int)lda, &one, C, (int)ldc);
    if (b != CUBLAS_STATUS_SUCCESS) { set_error("gemm failed (status %d)", (int)b); return CSPBO_E_SOLVER; }
    g_launches++;
    return CSPBO_OK;
}

extern "C" int cspbo_gp_stream_sm_target(void *stream, int sm_count)
{
    StreamHandles *sh;
    int rc = stream_handles(static_cast<cudaStream_t>(stream), &sh);
    if (rc) return rc;
    if (cublasSetSmCountTarget(sh->blas, sm_count) != CUBLAS_STATUS_SUCCESS) { set_error("cublasSetSmCountTarget(%d) failed", sm_count); return CSPBO_E_SOLVER; }
    return CSPBO_OK;
}


// ------------------------------------------------------------------------------------------
// SM partition for the sharded Cholesky.  Measured on B200 (tools/chol_probe.py, CSPBO_CHOL_TRACE=1): the
// potrf + inverse of a 384 x 384 diagonal block takes 0.37 ms alone but 1.1-1.4 ms while the trailing dgemm updates
// run -- stream priorities only order CTA dispatch, the latency-bound potrf kernels still share SMs (and their FP64
// pipes) with DMMA-saturating dgemm CTAs.  CUDA green contexts give the diagonal chain a small set of SMs of its
// own: one green context with `diag_sms` SMs (diagonal stream) and one with the remaining SMs (panel, look-ahead and
// update streams).  Driver entry points are resolved at run time (no link-time dependency on libcuda).
// ------------------------------------------------------------------------------------------
namespace cspbo {
struct GreenStreams {
    bool tried = false, ok = false;
    cudaStream_t diag = nullptr, panel = nullptr, look = nullptr, update = nullptr;
    int diag_sms = 0, rest_sms = 0;
};
static GreenStreams g_green[64];

template <typename F>
static bool drv(const char *name, F *fn)
{
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}
}  // namespace cspbo

extern "C" int cspbo_gp_chol_streams(int diag_sms, void **diag, void **panel, void **look, void **update, int *diag_sms_got,
                                     int *rest_sms_got)
{
    if (!diag || !panel || !look || !update) { set_error("gp_chol_streams: null argument"); return CSPBO_E_ARG; }
    int dev = 0;
    CSPBO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return CSPBO_E_ARG; }
    std::lock_guard<std::mutex> lk(g_solver_mu);
    GreenStreams &g = g_green[dev];
    if (!g.tried) {
        g.tried = true;
        CSPBO_CUDA(cudaFree(0));   // primary context
        CUresult (*pDeviceGet)(CUdevice *, int) = nullptr;
        CUresult (*pGetDevResource)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
        CUresult (*pSplit)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
        CUresult (*pGenDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
        CUresult (*pGreenCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
        CUresult (*pStreamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
        if (!drv("cuDeviceGet", &pDeviceGet) || !drv("cuDeviceGetDevResource", &pGetDevResource) ||
            !drv("cuDevSmResourceSplitByCount", &pSplit) || !drv("cuDevResourceGenerateDesc", &pGenDesc) ||
            !drv("cuGreenCtxCreate", &pGreenCreate) || !drv("cuGreenCtxStreamCreate", &pStreamCreate)) {
            set_error("gp_chol_streams: the driver has no green-context entry points");
            return CSPBO_E_CUDA;
        }
        CUdevice cudev;
        CUdevResource all, grp[1], rest;
        unsigned int nb = 1;
        CUdevResourceDesc d1, d2;
        CUgreenCtx c1, c2;
        CUstream s1, s2, s3, s4;
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        CUresult r;
#define GREEN_TRY(call) if ((r = (call)) != CUDA_SUCCESS) { set_error("gp_chol_streams: %s failed (%d)", #call, (int)r); return CSPBO_E_CUDA; }
        GREEN_TRY(pDeviceGet(&cudev, dev));
        GREEN_TRY(pGetDevResource(cudev, &all, CU_DEV_RESOURCE_TYPE_SM));
        GREEN_TRY(pSplit(grp, &nb, &all, &rest, 0, (unsigned)std::max(diag_sms, 8)));
        if (nb < 1 || rest.sm.smCount < 8) { set_error("gp_chol_streams: SM split gave %u groups, %u SMs left", nb, rest.sm.smCount); return CSPBO_E_CUDA; }
        GREEN_TRY(pGenDesc(&d1, &grp[0], 1));
        GREEN_TRY(pGenDesc(&d2, &rest, 1));
        GREEN_TRY(pGreenCreate(&c1, d1, cudev, CU_GREEN_CTX_DEFAULT_STREAM));
        GREEN_TRY(pGreenCreate(&c2, d2, cudev, CU_GREEN_CTX_DEFAULT_STREAM));
        GREEN_TRY(pStreamCreate(&s1, c1, CU_STREAM_NON_BLOCKING, hi));
        GREEN_TRY(pStreamCreate(&s2, c2, CU_STREAM_NON_BLOCKING, hi));
        GREEN_TRY(pStreamCreate(&s3, c2, CU_STREAM_NON_BLOCKING, std::min(hi + 1, lo)));
        GREEN_TRY(pStreamCreate(&s4, c2, CU_STREAM_NON_BLOCKING, lo));
#undef GREEN_TRY
        g.diag = s1; g.panel = s2; g.look = s3; g.update = s4;
        g.diag_sms = (int)grp[0].sm.smCount; g.rest_sms = (int)rest.sm.smCount;
        g.ok = true;
    }
    if (!g.ok) { set_error("gp_chol_streams: green contexts unavailable on this device"); return CSPBO_E_CUDA; }
    *diag = g.diag; *panel = g.panel; *look = g.look; *update = g.update;
    if (diag_sms_got) *diag_sms_got = g.diag_sms;
    if (rest_sms_got) *rest_sms_got = g.rest_sms;
    return CSPBO_OK;
}
