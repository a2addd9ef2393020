// dense.cu -- triangular inverse and K^-1 of the single-GPU Cholesky factor as recursive, dgemm-rich sweeps.
//
// The hyper-gradient of the log marginal likelihood (reference cspbo/gaussianprocess.py:490-499) needs diag(K^-1) (Dot_mb)
// or all of K^-1 (RBF_mb: contracted with dK/dl inside the tile kernel, tiles.cu trace mode).  cuSOLVER's trtri / potri
// run at 7.6 / 11.9 TFLOP/s on B200 at N = 20 001 (352 / 450 ms, tools/lml_breakdown.py) -- far below the 36 TFLOP/s of
// cuBLAS dgemm -- so the two sweeps are restated here over the factor L (column-major lower, leading dimension ld):
//
//   X = L^-1:   [L11 0; L21 L22]^-1 = [X11 0; -X22 L21 X11, X22]          (trtri_rec)
//   A = X^T X:  [X11^T X11 + X21^T X21, .; X22^T X21, X22^T X22]          (lauum_rec)
//
// The nb x nb diagonal leaves are inverted / squared up front in ONE batched launch each (they only depend on the
// leaf itself); everything else is a large triangular-times-general product (cuBLAS trmm), a syrk or a dgemm.  The
// triangular products are out of place (cuBLAS' trmm with C = B, the BLAS in-place form, was measured 36 % slower: 115 /
// 213 ms), chunked through a workspace of a few per cent of N^2, and copied back.  Measured at N = 20 001 (tools/dense_probe.py): L^-1 83.5 ms (32 TFLOP/s, cuSOLVER trtri
// 356 ms), K^-1 165 ms (cuSOLVER potri 454 ms).  Splitting the trmm / syrk calls further into dgemms by hand (down to
// 2048 rows) was measured 6 % SLOWER than handing cuBLAS the whole triangle; workspace 96 / 192 / 384 / 800 MB: 84.9 /
// 83.5 / 82.6 / 82.4 ms.
#include <cublas_v2.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "common.h"

namespace cspbo {

int tri_inverse_leaves(const double *L, long long ld, int N, int nb, double *inv, cudaStream_t st);   // gp.cu

namespace {

std::mutex g_dense_mu;
cublasHandle_t g_dense_blas[64] = {nullptr};

int env_int(const char *name, int def)
{
    const char *s = getenv(name);
    return (s && *s) ? atoi(s) : def;
}

struct Dense {
    cublasHandle_t blas = nullptr;
    cudaStream_t st = nullptr;
    double *L = nullptr;
    long long ld = 0;
    int N = 0, nb = 384;
    double *T = nullptr;     // [nleaf][nb*nb] leaf results
    double *W = nullptr;     // chunk workspace
    long long wcap = 0;      // doubles
    int rc = CSPBO_OK;

    int off(int b) const { return (int)std::min<long long>((long long)b * nb, N); }
    double *at(int i, int j) const { return L + (long long)j * ld + i; }
    void fail(const char *what, int status)
    {
        if (rc == CSPBO_OK) { set_error("dense: %s failed (status %d)", what, status); rc = CSPBO_E_SOLVER; }
    }

    void gemm(cublasOperation_t ta, cublasOperation_t tb, int m, int n, int k, double alpha, const double *A, long long lda,
              const double *B, long long ldb, double beta, double *C, long long ldc)
    {
        if (rc || m <= 0 || n <= 0 || k <= 0) return;
        cublasStatus_t s = cublasDgemm(blas, ta, tb, m, n, k, &alpha, A, (int)lda, B, (int)ldb, &beta, C, (int)ldc);
        g_launches++;
        if (s != CUBLAS_STATUS_SUCCESS) fail("dgemm", (int)s);
    }

    // Wo[n, c] = alpha op(A) B, A n x n lower (leading dimension ld), B n x c (ld), Wo (ldw): out of place
    void trmm_left(bool trans, double alpha, const double *A, int n, const double *B, int c, double *Wo, long long ldw)
    {
        if (rc || n <= 0 || c <= 0) return;
        cublasStatus_t s = cublasDtrmm(blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, trans ? CUBLAS_OP_T : CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT,
                                       n, c, &alpha, A, (int)ld, B, (int)ld, Wo, (int)ldw);
        g_launches++;
        if (s != CUBLAS_STATUS_SUCCESS) fail("dtrmm", (int)s);
    }

    // Wo[r, n] = alpha B A, A n x n lower, B r x n
    void trmm_right(double alpha, const double *A, int n, const double *B, int r, double *Wo, long long ldw)
    {
        if (rc || n <= 0 || r <= 0) return;
        cublasStatus_t s = cublasDtrmm(blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, r, n, &alpha, A,
                                       (int)ld, B, (int)ld, Wo, (int)ldw);
        g_launches++;
        if (s != CUBLAS_STATUS_SUCCESS) fail("dtrmm", (int)s);
    }

    void copy_back(double *dst, const double *src, long long ldw, int rows, int cols)
    {
        if (rc) return;
        cudaError_t e = cudaMemcpy2DAsync(dst, (size_t)ld * sizeof(double), src, (size_t)ldw * sizeof(double), (size_t)rows * sizeof(double),
                                          (size_t)cols, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { set_error("dense: copy-back failed: %s", cudaGetErrorString(e)); rc = CSPBO_E_CUDA; }
    }

    // B[n, c] <- alpha op(A) B, by column chunks through W
    void apply_left(bool trans, double alpha, const double *A, int n, double *B, int c)
    {
        const int cc = (int)std::max<long long>(1, std::min<long long>(c, wcap / n));
        for (int c0 = 0; c0 < c && !rc; c0 += cc) {
            const int w = std::min(cc, c - c0);
            trmm_left(trans, alpha, A, n, B + (long long)c0 * ld, w, W, n);
            copy_back(B + (long long)c0 * ld, W, n, n, w);
        }
    }

    // B[r, n] <- alpha B A, by row chunks through W
    void apply_right(double alpha, const double *A, int n, double *B, int r)
    {
        const int rr = (int)std::max<long long>(1, std::min<long long>(r, wcap / n));
        for (int r0 = 0; r0 < r && !rc; r0 += rr) {
            const int h = std::min(rr, r - r0);
            trmm_right(alpha, A, n, B + r0, h, W, h);
            copy_back(B + r0, W, h, h, n);
        }
    }

    // C[n, n] (lower) += A^T A, A k x n
    void syrk(double *C, int n, const double *A, int k)
    {
        if (rc || n <= 0 || k <= 0) return;
        const double one = 1.0;
        cublasStatus_t s = cublasDsyrk(blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, n, k, &one, A, (int)ld, &one, C, (int)ld);
        g_launches++;
        if (s != CUBLAS_STATUS_SUCCESS) fail("dsyrk", (int)s);
    }

    void leaf_to_block(int b)
    {
        const int h = off(b + 1) - off(b);
        if (rc || h <= 0) return;
        cudaError_t e = cudaMemcpy2DAsync(at(off(b), off(b)), (size_t)ld * sizeof(double), T + (long long)b * nb * nb, (size_t)h * sizeof(double),
                                          (size_t)h * sizeof(double), (size_t)h, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) { set_error("dense: leaf copy failed: %s", cudaGetErrorString(e)); rc = CSPBO_E_CUDA; }
    }

    void trtri_rec(int b0, int b1)     // leaves are already inverted in place
    {
        if (rc || b1 - b0 <= 1) return;
        const int bm = (b0 + b1) / 2;
        trtri_rec(b0, bm);
        trtri_rec(bm, b1);
        const int n1 = off(bm) - off(b0), n2 = off(b1) - off(bm);
        double *L21 = at(off(bm), off(b0));
        apply_left(false, -1.0, at(off(bm), off(bm)), n2, L21, n1);
        apply_right(1.0, at(off(b0), off(b0)), n1, L21, n2);
    }

    void lauum_rec(int b0, int b1)     // T holds X_bb^T X_bb of every leaf
    {
        if (rc) return;
        if (b1 - b0 <= 1) { leaf_to_block(b0); return; }
        const int bm = (b0 + b1) / 2;
        const int n1 = off(bm) - off(b0), n2 = off(b1) - off(bm);
        double *X21 = at(off(bm), off(b0));
        lauum_rec(b0, bm);
        syrk(at(off(b0), off(b0)), n1, X21, n2);
        apply_left(true, 1.0, at(off(bm), off(bm)), n2, X21, n1);
        lauum_rec(bm, b1);
    }
};

}  // namespace

// Leaf size, leaf-result doubles and chunk-workspace doubles for an N x N factor.  Leaves of 384 rows (128 below 8192
// rows, where 384-row leaf results would be a tenth of the matrix); chunk workspace = the largest off-diagonal block,
// capped at N^2/16 (200 MB at N = 20 001) unless CSPBO_DENSE_WS_MB says otherwise.
static void dense_sizes(int N, int *nb, long long *tsz, long long *wcap)
{
    *nb = std::min(std::max(env_int("CSPBO_DENSE_NB", N >= 8192 ? 384 : 128), 32), 512);
    const long long nleaf = (N + *nb - 1) / *nb, half = ((long long)N + 1) / 2;
    *tsz = nleaf * *nb * *nb;
    const long long cap = getenv("CSPBO_DENSE_WS_MB") ? (long long)env_int("CSPBO_DENSE_WS_MB", 192) * (1 << 20) / 8 : (long long)N * N / 16;
    *wcap = std::max<long long>(std::min<long long>(half * half, cap), (long long)N);
}

long long dense_workspace_bytes(int N)
{
    int nb;
    long long tsz, wcap;
    dense_sizes(N, &nb, &tsz, &wcap);
    return (tsz + wcap) * (long long)sizeof(double);
}

// what: 0 = L -> L^-1, 1 = L -> (L L^T)^-1 (lower triangle), both in place on the legacy default stream.
// The other triangle of the buffer is scratch (the leaves' upper triangles are zeroed).
int dense_factor_inverse(double *L, int N, long long ld, int what)
{
    int dev = 0;
    CSPBO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) { set_error("device ordinal %d out of range", dev); return CSPBO_E_ARG; }
    Dense D;
    {
        std::lock_guard<std::mutex> lk(g_dense_mu);
        if (!g_dense_blas[dev] && cublasCreate(&g_dense_blas[dev]) != CUBLAS_STATUS_SUCCESS) {
            g_dense_blas[dev] = nullptr;
            set_error("cublasCreate failed");
            return CSPBO_E_SOLVER;
        }
        D.blas = g_dense_blas[dev];
    }
    D.st = nullptr;
    cublasSetStream(D.blas, D.st);
    D.L = L; D.ld = ld; D.N = N;
    long long tsz = 0;
    dense_sizes(N, &D.nb, &tsz, &D.wcap);
    const int nleaf = (N + D.nb - 1) / D.nb;
    void *ws = nullptr;
    int rc = dev_alloc_cached((size_t)(tsz + D.wcap) * sizeof(double), &ws);
    if (rc) return rc;
    D.T = static_cast<double *>(ws);
    D.W = D.T + tsz;

    // 1. all leaves inverted at once, then written back over their blocks (zeros above the diagonal)
    rc = tri_inverse_leaves(L, ld, N, D.nb, D.T, D.st);
    if (rc) { dev_free_cached(ws); return rc; }
    for (int b = 0; b < nleaf; b++) D.leaf_to_block(b);
    // 2. off-diagonal blocks
    D.trtri_rec(0, nleaf);
    if (what == 1 && !D.rc) {
        // 3. leaf products X_bb^T X_bb (batched dgemm; the ragged last leaf on its own), then the recursive X^T X
        const int nfull = N / D.nb, tail = N - nfull * D.nb;
        const double one = 1.0, zero = 0.0;
        if (nfull > 0) {
            cublasStatus_t s = cublasDgemmStridedBatched(D.blas, CUBLAS_OP_T, CUBLAS_OP_N, D.nb, D.nb, D.nb, &one, L, (int)ld,
                                                         (long long)D.nb * (ld + 1), L, (int)ld, (long long)D.nb * (ld + 1), &zero, D.T, D.nb,
                                                         (long long)D.nb * D.nb, nfull);
            g_launches++;
            if (s != CUBLAS_STATUS_SUCCESS) D.fail("dgemmStridedBatched", (int)s);
        }
        if (tail > 0) {
            const double *Xt = D.at(nfull * D.nb, nfull * D.nb);
            D.gemm(CUBLAS_OP_T, CUBLAS_OP_N, tail, tail, tail, 1.0, Xt, ld, Xt, ld, 0.0, D.T + (long long)nfull * D.nb * D.nb, tail);
        }
        D.lauum_rec(0, nleaf);
    }
    // the workspace goes back to the cache only once the stream is done with it
    cudaError_t e = cudaStreamSynchronize(D.st);
    dev_free_cached(ws);
    if (D.rc) return D.rc;
    if (e != cudaSuccess) { set_error("dense: %s", cudaGetErrorString(e)); return CSPBO_E_CUDA; }
    return CSPBO_OK;
}

}  // namespace cspbo
