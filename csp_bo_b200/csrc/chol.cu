// chol.cu -- the diagonal block of the sharded Cholesky as ONE cluster-resident kernel (opt-in: CSPBO_CHOL_DIAG=fused).
//
// The only sequential chain of the distributed factorisation K = U^T U (csp_bo_b200/distributed.py: ShardedCholesky,
// replacing the replicated scipy `cholesky` of reference cspbo/gaussianprocess.py:116) is
//     potrf(diagonal block s) -> inverse of its factor -> narrow panel piece -> update of diagonal block s+1 -> ...
// Its first link is cuSOLVER potrf by default (several launches; 0.17 ms for h = 384 alone, 0.22-0.25 ms next to the trailing
// dgemms on the 16-SM green context).  Here the h x h block (h <= 384) is factorised by a single thread-block CLUSTER of 8
// CTAs that keeps the whole lower triangle -- 78 tiles of 32 x 32 -- in its shared memory (tile row i lives in CTA i % 8,
// at most 16 tiles = 128 KB per CTA) and talks through distributed shared memory and the hardware cluster barrier:
//
//   P1  one warp of the owner: Cholesky of the 32 x 32 diagonal tile in registers (one row per lane, pivot row broadcast
//       by shuffles, rsqrt);
//   P2  every CTA solves its own row tiles against that factor (which its owner PUSHED into everybody's shared memory),
//       one row per lane, right-looking substitution in registers, and pushes the finished tile to every CTA;
//   P3  every CTA updates its own rows, one tile per warp, all operands in local shared memory; the owner of tile (k+1, k+1) lets its 16 warps share that tile first and then
//       runs P1 of the next step on warp 0 while everybody else finishes the update (look-ahead inside the kernel).
//
// Two cluster barriers per 32-column step, no global-memory traffic between the initial load and the final store.
// Storage follows cuSOLVER's: the block is symmetric, row-major with leading dimension ld, the factor is left in the
// row-major UPPER triangle (U(i,j) = D[i*ld + j], j >= i; equivalently the column-major lower L), the strictly lower
// triangle is not touched.
//
// Status (B200, profiles/chol_probe_r02.md): correct to 2e-15 and AT PARITY with cuSOLVER potrf -- 0.151 vs 0.165 ms for
// h = 384 alone, 0.285 vs 0.29 ms per panel (with the inverse) inside the running factorisation -- not faster, so cuSOLVER
// stays the default.  The first versions round-tripped the block through L2 in every phase (0.26 ms).  What limits this
// one, from its own clock64 stamps (CSPBO_CHOL_DBG=1): the 32 dependent pivot steps of P1 (~200 cycles each: shuffle, FP64
// rsqrt, multiply, shuffle, FMA) and the DSMEM bandwidth of 17-21 B/cycle per SM against up to 88 KB of block column that
// every CTA needs in every step.
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "common.h"

namespace cg = cooperative_groups;

namespace cspbo {

constexpr int kChCtas = 8;
constexpr int kChWarps = 16;
constexpr int CB = 32;                       // tile edge
constexpr int CT = CB * CB;                  // doubles per tile, stored column-major: element (r, c) at c*32 + r
constexpr int kChMaxNb = 12;                 // h <= 384
constexpr int kOwnTiles = 16;                // rows c and c+8 of CTA c: (c+1) + (c+9) <= 16 tiles
constexpr int kPanelTiles = kChMaxNb - 1;

struct CholDiagArgs {
    double *D;
    long long ld;
    int h;
    int *info;          // device info word: 0, or 1 + index of the first non-positive pivot
    long long *dbg;     // optional [8 CTAs][64] clock64 stamps (CSPBO_CHOL_DBG)
};

// local index of tile (i, j), j <= i, inside the shared memory of its owner (CTA i % 8)
__device__ __forceinline__ int ch_tile_index(int i, int j) { return (i >= kChCtas ? (i - kChCtas) + 1 : 0) + j; }

// P1: Cholesky of one 32 x 32 diagonal tile held in shared memory (column-major), by one warp, in registers.  The factor
// goes back to the tile and is PUSHED into the `lkk` buffer of every CTA of the cluster (remote shared-memory stores are
// posted; the next cluster barrier makes them visible), so nobody has to pull it across the cluster afterwards.
__device__ void chol_p1(double *T, int row0, const CholDiagArgs &a, cg::cluster_group &cluster, double *lkk)
{
    const int lane = threadIdx.x & 31;
    double v[CB];
#pragma unroll
    for (int c = 0; c < CB; c++) v[c] = (c <= lane) ? T[c * CB + lane] : 0.0;
    int bad = 0;
#pragma unroll
    for (int kk = 0; kk < CB; kk++) {
        double piv = __shfl_sync(0xffffffffu, v[kk], kk);
        if (!(piv > 0.0)) { if (!bad) bad = row0 + kk + 1; piv = 1.0; }
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
    for (int c = 0; c < CB; c++) T[c * CB + lane] = (c <= lane) ? v[c] : 0.0;
    for (int r = 0; r < kChCtas; r++) {
        double *dst = cluster.map_shared_rank(lkk, r);
#pragma unroll
        for (int c = 0; c < CB; c++) dst[c * CB + lane] = (c <= lane) ? v[c] : 0.0;
    }
}

__global__ void __launch_bounds__(kChWarps * 32, 1) chol_diag_kernel(const CholDiagArgs a)
{
    extern __shared__ double csm[];
    double *own = csm;                                    // [kOwnTiles][CT] the tile rows of this CTA
    double *panel = own + kOwnTiles * CT;                 // [kPanelTiles][CT] block column k (pushed by the row owners)
    double *lkk = panel + kPanelTiles * CT;               // [CT] factor of the current diagonal tile (pushed by its owner)
    cg::cluster_group cluster = cg::this_cluster();
    const int cta = (int)cluster.block_rank();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nb = (a.h + CB - 1) / CB;
    int nst = 0;
#define CH_STAMP() do { if (a.dbg && threadIdx.x == 0 && nst < 64) a.dbg[cta * 64 + nst++] = clock64(); } while (0)
    CH_STAMP();

    // ---- load the tile rows this CTA owns (identity padding beyond h); L-notation: element (r, c), r >= c, at D[c*ld + r].
    // All copies below are written as batches of independent loads followed by the stores, so that a thread has 8 requests
    // in flight instead of paying one memory latency per tile.
    {
        const int n_own = ((cta < nb) ? cta + 1 : 0) + ((cta + kChCtas < nb) ? cta + kChCtas + 1 : 0);
        for (int base = threadIdx.x; base < n_own * CT; base += blockDim.x * 8) {
            double tmp[8];
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int idx = base + u * blockDim.x;
                tmp[u] = 0.0;
                if (idx < n_own * CT) {
                    const int tl = idx >> 10, t = idx & (CT - 1);
                    const int i = (tl <= cta) ? cta : cta + kChCtas, j = (tl <= cta) ? tl : tl - (cta + 1);
                    const int r = i * CB + (t & 31), c = j * CB + (t >> 5);
                    tmp[u] = (r == c) ? 1.0 : 0.0;
                    if (r < a.h && c < a.h) tmp[u] = (r >= c) ? a.D[(long long)c * a.ld + r] : a.D[(long long)r * a.ld + c];
                }
            }
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const int idx = base + u * blockDim.x;
                if (idx < n_own * CT) own[idx] = tmp[u];
            }
        }
    }
    __syncthreads();
    CH_STAMP();
    if (cta == 0 && warp == 0) chol_p1(own + ch_tile_index(0, 0) * CT, 0, a, cluster, lkk);
    CH_STAMP();
    cluster.sync();
    CH_STAMP();

    for (int k = 0; k < nb - 1; k++) {
        // ---- P2: own row tiles (i, k), i > k: x L_kk^T = a; the finished tile stays here (operand L_ik of P3) and is pushed
        // into slot i-k-1 of every CTA's panel buffer (operand L_jk of P3)
        {
            const int i = cta + warp * kChCtas;             // the (at most two) rows of this CTA: warp 0 / warp 1
            if (warp < 2 && i > k && i < nb) {
                double *T = own + ch_tile_index(i, k) * CT;
                const double rdl = 1.0 / lkk[lane * CB + lane];
                double av[CB];
#pragma unroll
                for (int c = 0; c < CB; c++) av[c] = T[c * CB + lane];
#pragma unroll
                for (int c = 0; c < CB; c++) {
                    const double x = av[c] * __shfl_sync(0xffffffffu, rdl, c);
                    av[c] = x;
#pragma unroll
                    for (int c2 = c + 1; c2 < CB; c2++) av[c2] = fma(-x, lkk[c * CB + c2], av[c2]);   // L_kk(c2, c)
                }
#pragma unroll
                for (int c = 0; c < CB; c++) T[c * CB + lane] = av[c];
                for (int r = 0; r < kChCtas; r++) {
                    double *dst = cluster.map_shared_rank(panel + (i - k - 1) * CT, r);
#pragma unroll
                    for (int c = 0; c < CB; c++) dst[c * CB + lane] = av[c];
                }
            }
        }
        CH_STAMP();
        cluster.sync();
        CH_STAMP();
        // ---- P3: update the own rows with block column k, all operands in local shared memory
        const bool chain = cta == (k + 1) % kChCtas;     // this CTA owns the next diagonal tile
        if (chain) {
            // tile (k+1, k+1): 2 columns per warp, then warp 0 factorises it
            double *T = own + ch_tile_index(k + 1, k + 1) * CT;
            const double *Lk = own + ch_tile_index(k + 1, k) * CT;
            double lv[CB];
#pragma unroll
            for (int m = 0; m < CB; m++) lv[m] = Lk[m * CB + lane];
#pragma unroll
            for (int cc = 0; cc < CB / kChWarps; cc++) {
                const int c = warp * (CB / kChWarps) + cc;
                double s = 0.0, s2 = 0.0;
#pragma unroll
                for (int m = 0; m < CB; m += 2) { s = fma(lv[m], Lk[m * CB + c], s); s2 = fma(lv[m + 1], Lk[(m + 1) * CB + c], s2); }
                T[c * CB + lane] -= s + s2;
            }
            __syncthreads();
            if (warp == 0) chol_p1(T, (k + 1) * CB, a, cluster, lkk);
        }
        // remaining tiles of the own rows: (i, j), k < j <= i, minus the chain tile; one per warp (warp 0 of the chain CTA
        // is busy with P1: its share goes to warps 1..15)
        {
            const int w = chain ? warp - 1 : warp;
            const int stride = chain ? kChWarps - 1 : kChWarps;
            if (w >= 0) {
                int seen = 0;
                for (int i = cta; i < nb; i += kChCtas) {
                    if (i <= k) continue;
                    for (int j = k + 1; j <= i; j++) {
                        if (i == k + 1 && j == k + 1) continue;         // the chain tile, done above
                        if (seen++ % stride != w) continue;
                        double *T = own + ch_tile_index(i, j) * CT;
                        const double *Li = own + ch_tile_index(i, k) * CT;
                        const double *Lj = panel + (j - k - 1) * CT;
                        double lv[CB];
#pragma unroll
                        for (int m = 0; m < CB; m++) lv[m] = Li[m * CB + lane];
#pragma unroll 4
                        for (int c = 0; c < CB; c++) {
                            double s = 0.0, s2 = 0.0;
#pragma unroll
                            for (int m = 0; m < CB; m += 2) { s = fma(lv[m], Lj[m * CB + c], s); s2 = fma(lv[m + 1], Lj[(m + 1) * CB + c], s2); }
                            T[c * CB + lane] -= s + s2;
                        }
                    }
                }
            }
        }
        CH_STAMP();
        cluster.sync();
        CH_STAMP();
    }
    // ---- store the factor: lower triangle in L-notation = row-major upper triangle of D
    for (int i = cta; i < nb; i += kChCtas)
        for (int j = 0; j <= i; j++) {
            const double *T = own + ch_tile_index(i, j) * CT;
            for (int t = threadIdx.x; t < CT; t += blockDim.x) {
                const int r = i * CB + (t & 31), c = j * CB + (t >> 5);
                if (r < a.h && c < a.h && r >= c) a.D[(long long)c * a.ld + r] = T[t];
            }
        }
}

size_t chol_diag_smem() { return (size_t)((kOwnTiles + kPanelTiles + 1) * CT) * sizeof(double); }

// Launch on `st`; info_dev as cuSOLVER's devInfo (overwritten: 0 or the index of the failing leading minor).  Returns
// CSPBO_E_CUDA without side effects when the stream's SM partition does not take an 8-CTA cluster (caller falls back).
int chol_diag_launch(double *D, long long ld, int h, int *info_dev, cudaStream_t st)
{
    if (h > kChMaxNb * CB) { set_error("chol_diag: block of %d rows exceeds %d", h, kChMaxNb * CB); return CSPBO_E_ARG; }
    static bool configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    const int di = (dev >= 0 && dev < 64) ? dev : 0;
    const size_t smem = chol_diag_smem();
    if (!configured[di]) {
        CSPBO_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[di] = true;
    }
    CSPBO_CUDA(cudaMemsetAsync(info_dev, 0, sizeof(int), st));
    static long long *dbg = nullptr;
    static const bool want_dbg = getenv("CSPBO_CHOL_DBG") != nullptr;
    if (want_dbg && !dbg) { cudaMalloc(&dbg, 8 * 64 * sizeof(long long)); }
    if (want_dbg) cudaMemsetAsync(dbg, 0, 8 * 64 * sizeof(long long), st);
    CholDiagArgs a = {D, ld, h, info_dev, want_dbg ? dbg : nullptr};
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kChCtas); cfg.blockDim = dim3(kChWarps * 32); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kChCtas; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, chol_diag_kernel, a);
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("chol_diag: cluster launch failed: %s", cudaGetErrorString(e));
        return CSPBO_E_CUDA;
    }
    g_launches++;
    if (want_dbg) {
        static int shown = 0;
        cudaStreamSynchronize(st);
        if (shown++ == 3) {
            long long hbuf[8 * 64];
            cudaMemcpy(hbuf, dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
            for (int c = 0; c < 8; c += 3) {
                fprintf(stderr, "chol dbg cta %d:", c);
                for (int i = 1; i < 64 && hbuf[c * 64 + i]; i++) fprintf(stderr, " %lld", hbuf[c * 64 + i] - hbuf[c * 64 + i - 1]);
                fprintf(stderr, "\n");
            }
        }
    }
    return CSPBO_OK;
}

}  // namespace cspbo
