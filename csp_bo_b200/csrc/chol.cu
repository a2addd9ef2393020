// chol.cu -- the diagonal block of the sharded Cholesky as ONE kernel (EXPERIMENTAL, opt-in: CSPBO_CHOL_DIAG=fused).
//
// Status (round 2, B200): bit-level sane (|U^T U - A| <= 2e-15) but slower than the cuSOLVER potrf it was meant to replace --
// 0.26 ms against 0.17 ms for h = 384 alone, 0.40 against 0.30 ms per panel inside the running factorisation -- so it is
// not the default.  Every phase round-trips the 1.2 MB block through L2 (about seven dependent global-memory hops per 32-column
// step); the version that would win keeps the 78 lower-triangle tiles in the shared memory of the 8 CTAs of the cluster and
// reads partner tiles through distributed shared memory.  Kept, with its test, as the starting point for that.
//
// The only sequential chain of the distributed factorisation K = U^T U (csp_bo_b200/distributed.py: ShardedCholesky,
// replacing the replicated scipy `cholesky` of reference cspbo/gaussianprocess.py:116) is
//     potrf(diagonal block s) -> inverse of its factor -> narrow panel piece -> update of diagonal block s+1 -> ...
// Round 1 ran the first link as cuSOLVER potrf (several launches; 0.17 ms alone, 0.3-0.4 ms next to the trailing dgemms on
// the 16-SM green context) + a hand-written triangular inverse (0.08 ms).  Here the h x h block (h <= 384) is factorised by
// a single kernel of kChCtas co-resident CTAs that walk 32-column block steps with a software grid barrier:
//
//   P1  one warp: Cholesky of the 32 x 32 diagonal tile in registers (one row per lane, pivot row broadcast by shuffles);
//   P2  all warps: L_ik = A_ik L_kk^-T for the tiles below it, one row per lane, right-looking substitution in registers;
//   P3  all warps: trailing update A_ij -= L_ik L_jk^T, one 32 x 32 tile per warp; tile (k+1, k+1) is shared by the 8 warps
//       of CTA 0, whose warp 0 then runs P1 of the next step while the others finish the update (look-ahead in the kernel).
//
// Two barriers per block step (23 for h = 384): the hardware cluster barrier when the 8 CTAs can be launched as one
// thread-block cluster, a global-memory counter otherwise.  Storage follows cuSOLVER's: the block is symmetric, row-major with
// leading dimension ld, the factor is left in the row-major UPPER triangle (U(i,j) = D[i*ld + j], j >= i; equivalently the
// column-major lower L), the strictly lower triangle is not touched.
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "common.h"

namespace cg = cooperative_groups;

namespace cspbo {

constexpr int kChCtas = 8;
constexpr int kChWarps = 8;
constexpr int CB = 32;          // tile edge
constexpr int CBP = CB + 1;     // padded shared-memory row

struct CholDiagArgs {
    double *D;
    long long ld;
    int h;
    int *info;          // device info word: 0, or 1 + index of the first non-positive pivot
    unsigned *bar;      // software barrier only: [0] arrival counter, [1] exit counter (both 0 between launches)
};

// Barrier over the kChCtas CTAs.  CLUSTER: they form one thread-block cluster and the hardware cluster barrier
// (release / acquire at cluster scope) orders their global-memory traffic; otherwise an atomic counter in global memory.
template <bool CLUSTER>
__device__ __forceinline__ void chol_grid_bar(unsigned *counter, unsigned &target)
{
    if (CLUSTER) {
        __threadfence();
        cg::this_cluster().sync();
        return;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        target += kChCtas;
        __threadfence();
        atomicAdd(counter, 1u);
        // all kChCtas CTAs are co-resident (one per SM of the >= 8-SM partition the diagonal stream runs on); the bounded
        // spin turns a scheduling accident into a trap instead of a hung GPU
        for (unsigned spin = 0; *reinterpret_cast<volatile unsigned *>(counter) < target; spin++)
            if (spin > (1u << 25)) __trap();
        __threadfence();
    }
    __syncthreads();
}

// L-notation: the lower triangle element (r, c), r >= c, of the symmetric block lives at D[c*ld + r].
__device__ __forceinline__ double ch_load(const CholDiagArgs &a, int r, int c)
{
    if (r < a.h && c < a.h) return a.D[(long long)c * a.ld + r];
    return r == c ? 1.0 : 0.0;      // identity padding of a ragged last tile
}

// P1: Cholesky of the diagonal tile k by one warp, in registers: lane i holds row i; step kk broadcasts the pivot, scales
// column kk with rsqrt and applies the rank-1 update with one shuffle per remaining column.
__device__ void chol_p1(const CholDiagArgs &a, int k)
{
    const int lane = threadIdx.x & 31;
    const int r0 = k * CB;
    double v[CB];
#pragma unroll
    for (int c = 0; c < CB; c++) v[c] = (c <= lane) ? ch_load(a, r0 + lane, r0 + c) : 0.0;
    int bad = 0;
#pragma unroll
    for (int kk = 0; kk < CB; kk++) {
        double piv = __shfl_sync(0xffffffffu, v[kk], kk);
        if (!(piv > 0.0)) { if (!bad) bad = r0 + kk + 1; piv = 1.0; }
        const double l = v[kk] * rsqrt(piv);        // lane i >= kk: L(i, kk); lanes above the diagonal hold 0
        v[kk] = l;
#pragma unroll
        for (int j = kk + 1; j < CB; j++) {
            const double ljk = __shfl_sync(0xffffffffu, l, j);
            v[j] = fma(-l, ljk, v[j]);              // A(i, j) -= L(i, kk) L(j, kk); only j <= i is ever used
        }
    }
    if (bad && lane == 0 && bad - 1 < a.h) atomicCAS(a.info, 0, bad);
#pragma unroll
    for (int c = 0; c < CB; c++)
        if (c <= lane && r0 + lane < a.h && r0 + c < a.h) a.D[(long long)(r0 + c) * a.ld + r0 + lane] = v[c];
}

template <bool CLUSTER>
__global__ void __launch_bounds__(kChWarps * 32, 1) chol_diag_kernel(const CholDiagArgs a)
{
    extern __shared__ double csm[];
    double *Ls = csm;                                     // [32][33] the current diagonal tile's factor (P2) / L_jk of the
    double *rd = Ls + CB * CBP;                           // first trailing tile (CTA 0 in P3); [32] reciprocal diagonal
    double *Ws = rd + CB;                                 // per warp [32][33]: the L_jk tile of P3
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = (a.h + CB - 1) / CB;
    // worker id for the tile pools: every warp except warp 0 of CTA 0, which runs the diagonal chain (P1)
    const int gw = blockIdx.x * kChWarps + warp;
    constexpr int kWorkers = kChCtas * kChWarps - 1;
    unsigned target = 0;

    if (gw == 0) chol_p1(a, 0);
    chol_grid_bar<CLUSTER>(a.bar, target);
    for (int k = 0; k < nb - 1; k++) {
        const int below = nb - 1 - k;                     // tiles under the diagonal tile of this step
        // ---- P2: L_ik = A_ik L_kk^-T, one row tile per warp, one row per lane: forward substitution in registers in its
        // right-looking form (x_c fixed, then the remaining columns updated -- independent FMAs, no serial dot products)
        for (int t = threadIdx.x; t < CB * CB; t += blockDim.x) {
            const int r = t & 31, c = t >> 5;
            Ls[r * CBP + c] = (c <= r) ? ch_load(a, k * CB + r, k * CB + c) : 0.0;
        }
        __syncthreads();
        if (threadIdx.x < CB) rd[threadIdx.x] = 1.0 / Ls[threadIdx.x * CBP + threadIdx.x];
        __syncthreads();
        for (int u = gw; u < below; u += kChCtas * kChWarps) {
            const int i = k + 1 + u;
            double av[CB];
#pragma unroll
            for (int m = 0; m < CB; m++) av[m] = ch_load(a, i * CB + lane, k * CB + m);
#pragma unroll
            for (int c = 0; c < CB; c++) {
                const double x = av[c] * rd[c];
                av[c] = x;
#pragma unroll
                for (int c2 = c + 1; c2 < CB; c2++) av[c2] = fma(-x, Ls[c2 * CBP + c], av[c2]);
            }
#pragma unroll
            for (int c = 0; c < CB; c++)
                if (i * CB + lane < a.h && k * CB + c < a.h) a.D[(long long)(k * CB + c) * a.ld + i * CB + lane] = av[c];
        }
        chol_grid_bar<CLUSTER>(a.bar, target);
        // ---- P3: A_ij -= L_ik L_jk^T for k < j <= i.  Tile 0 = (k+1, k+1) is on the critical path: the 8 warps of CTA 0
        // share it (4 columns each), then warp 0 factorises it (P1 of the next step) while everybody else works through the
        // remaining tiles -- the look-ahead that hides the diagonal chain behind the trailing update.
        const int ntiles = below * (below + 1) / 2;
        if (blockIdx.x == 0) {
            const int tj = k + 1;
            for (int t = threadIdx.x; t < CB * CB; t += blockDim.x) {
                const int r = t & 31, m = t >> 5;
                Ls[r * CBP + m] = ch_load(a, tj * CB + r, k * CB + m);
            }
            __syncthreads();
            double lv[CB];
#pragma unroll
            for (int m = 0; m < CB; m++) lv[m] = Ls[lane * CBP + m];
#pragma unroll
            for (int cc = 0; cc < CB / kChWarps; cc++) {
                const int c = warp * (CB / kChWarps) + cc;
                double s = 0.0, s2 = 0.0;
#pragma unroll
                for (int m = 0; m < CB; m += 2) { s = fma(lv[m], Ls[c * CBP + m], s); s2 = fma(lv[m + 1], Ls[c * CBP + m + 1], s2); }
                const int gr = tj * CB + lane, gc = tj * CB + c;
                if (gr < a.h && gc < a.h && gr >= gc) a.D[(long long)gc * a.ld + gr] -= s + s2;
            }
            __threadfence_block();
            __syncthreads();
            if (warp == 0) chol_p1(a, k + 1);
        }
        if (gw != 0) {
            double *Wj = Ws + (size_t)warp * CB * CBP;
            for (int t = 1 + (gw - 1); t < ntiles; t += kWorkers) {
                // t -> (i, j): column-major walk of the lower triangle (t = 0 was the first diagonal tile)
                int j = 0, rem = t;
                while (rem >= below - j) { rem -= below - j; j++; }
                const int ti = k + 1 + j + rem, tj = k + 1 + j;
                double lv[CB];
#pragma unroll
                for (int m = 0; m < CB; m++) {
                    lv[m] = ch_load(a, ti * CB + lane, k * CB + m);
                    Wj[lane * CBP + m] = ch_load(a, tj * CB + lane, k * CB + m);
                }
                __syncwarp();
#pragma unroll 4
                for (int c = 0; c < CB; c++) {
                    double s = 0.0, s2 = 0.0;
#pragma unroll
                    for (int m = 0; m < CB; m += 2) { s = fma(lv[m], Wj[c * CBP + m], s); s2 = fma(lv[m + 1], Wj[c * CBP + m + 1], s2); }
                    const int gr = ti * CB + lane, gc = tj * CB + c;
                    if (gr < a.h && gc < a.h && gr >= gc) a.D[(long long)gc * a.ld + gr] -= s + s2;
                }
                __syncwarp();
            }
        }
        chol_grid_bar<CLUSTER>(a.bar, target);
    }
    if (!CLUSTER) {
        // leave the barrier words at zero for the next launch: the last CTA out resets them
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(a.bar + 1, 1u) == kChCtas - 1) { a.bar[0] = 0; a.bar[1] = 0; __threadfence(); }
        }
    }
}

size_t chol_diag_smem() { return (size_t)(CB * CBP + CB + kChWarps * CB * CBP) * sizeof(double); }

// Launch on `st`.  bar: 2 unsigned words that are ZERO (kept zero by the kernel itself); info_dev as cuSOLVER's devInfo
// (overwritten: 0 or the index of the failing leading minor).  The 8 CTAs are launched as one thread-block cluster when the
// stream's SM partition accepts it (hardware cluster barrier), else as a plain grid with the software barrier.
int chol_diag_launch(double *D, long long ld, int h, unsigned *bar, int *info_dev, cudaStream_t st)
{
    if (h > 12 * CB) { set_error("chol_diag: block of %d rows exceeds %d", h, 12 * CB); return CSPBO_E_ARG; }
    static int mode[64] = {};        // 0: not configured, 1: cluster, 2: software barrier
    static const char *force = getenv("CSPBO_CHOL_BARRIER");     // "cluster" | "soft" (experiments)
    int dev = 0;
    cudaGetDevice(&dev);
    const int di = (dev >= 0 && dev < 64) ? dev : 0;
    const size_t smem = chol_diag_smem();
    if (!mode[di]) {
        CSPBO_CUDA(cudaFuncSetAttribute(chol_diag_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CSPBO_CUDA(cudaFuncSetAttribute(chol_diag_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        mode[di] = (force && !strcmp(force, "soft")) ? 2 : 1;
    }
    CSPBO_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
    CholDiagArgs a = {D, ld, h, info_dev, bar};
    if (mode[di] == 1) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kChCtas); cfg.blockDim = dim3(kChWarps * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = kChCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, chol_diag_kernel<true>, a);
        if (e == cudaSuccess) { g_launches++; return CSPBO_OK; }
        cudaGetLastError();
        if (force && !strcmp(force, "cluster")) { set_error("chol_diag: cluster launch failed: %s", cudaGetErrorString(e)); return CSPBO_E_CUDA; }
        mode[di] = 2;                // this partition does not take an 8-CTA cluster: software barrier from now on
    }
    chol_diag_kernel<false><<<kChCtas, kChWarps * 32, smem, st>>>(a);
    g_launches++;
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}

}  // namespace cspbo
