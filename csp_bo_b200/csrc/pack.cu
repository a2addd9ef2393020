// pack.cu -- "pack" step of the hot path: normalise every stacked descriptor row, project its
// Cartesian derivative onto the tangent space of the unit sphere, and scatter both into the
// species-bucketed, group-interleaved tile layout the DMMA tile kernels consume (common.h).
//
// Why this layout: with u = x/|x| and A^ = (I - u u^T) dxdr / |x| every
// per-pair quantity of the reference's d x d loop (dot_kernel.cpp:270-323) is one of the 16 dot
// products [u_a; A^_a] . [u_b; A^_b]^T:
//     c = u_a.u_b,  P_k = A^_a[:,k].u_b,  Q_l = u_a.A^_b[:,l],  T_kl = A^_a[:,k].A^_b[:,l]
// Rows are bucketed by species (pairs of different species contribute nothing,
// dot_kernel.cpp:255) and, inside a bucket, groups are sorted by row count and interleaved 8 at
// a time so that slot r of every tile of a block is the same group: a thread's MMA accumulator
// fragment then always belongs to one fixed (group_i, group_j) pair and the reference's
// scatter-add `pout[(k+_i*3)*x2i+_j)*3+l] +=` becomes a register accumulation.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "common.h"

namespace cspbo {

__global__ void pack_rows_kernel(int n, int d, int nc, int ks, int nv,
                                 const double *__restrict__ x, const double *__restrict__ dxdr,
                                 const long long *__restrict__ dest, double *__restrict__ tiles,
                                 unsigned *__restrict__ tile_valid, long long tile_doubles)
{
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n) return;
    const long long dst = dest[row];
    if (dst < 0) return;
    const long long tile = dst >> 3;
    const int slot = (int)(dst & 7);
    const double *xr = x + row * d;

    double ss = 0.0;
    for (int f = lane; f < d; f += 32) {
        double v = xr[f];
        ss += v * v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const double nrm = sqrt(ss);
    if (!(nrm > 1e-8)) return;  // reference: rows with |x| <= eps are skipped (dot_kernel.cpp:25)
    const double inv = 1.0 / nrm;

    double *t = tiles + tile * tile_doubles;
    const int half = ks >> 1;
    for (int f = lane; f < d; f += 32) {
        const long long idx = (((long long)(0 * half + (f >> 3)) * 32) + slot * 4 + (f & 3)) * 2 + ((f >> 2) & 1);
        t[idx] = xr[f] * inv;
    }
    if (nc > 0) {
        const double *dr = dxdr + row * (long long)d * nc;
        for (int k = 0; k < nc; k++) {
            double al = 0.0;  // alpha_k = dxdr[:,k] . u
            for (int f = lane; f < d; f += 32) al += dr[(long long)f * nc + k] * (xr[f] * inv);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) al += __shfl_xor_sync(0xffffffffu, al, o);
            const int v = 1 + k;
            for (int f = lane; f < d; f += 32) {
                const long long idx = (((long long)(v * half + (f >> 3)) * 32) + slot * 4 + (f & 3)) * 2 + ((f >> 2) & 1);
                t[idx] = (dr[(long long)f * nc + k] - al * (xr[f] * inv)) * inv;
            }
        }
    }
    if (lane == 0) atomicOr(tile_valid + tile, 1u << slot);
}

}  // namespace cspbo

using namespace cspbo;

extern "C" int cspbo_pack_create(int n, int d, int nc, int m, int row_start, int row_end,
                                 const double *x, const double *dxdr, const int *ele, const int *inds,
                                 int mem, cspbo_pack **out)
{
    if (!out) { set_error("pack_create: out is null"); return CSPBO_E_ARG; }
    *out = nullptr;
    if (n < 0 || m < 0 || d <= 0 || (n > 0 && (!x || !ele || !inds))) {
        set_error("pack_create: bad sizes/pointers (n=%d d=%d m=%d)", n, d, m);
        return CSPBO_E_ARG;
    }
    if (dxdr == nullptr) nc = 0;
    if (!(nc == 0 || nc == 3 || nc == 9)) { set_error("pack_create: nc must be 3 or 9 (got %d)", nc); return CSPBO_E_ARG; }
    const int ks = ksteps_for(d);
    if (ks < 0) { set_error("pack_create: descriptor length d=%d unsupported (max %d)", d, 56); return CSPBO_E_ARG; }
    row_start = std::max(row_start, 0);
    row_end = std::min(row_end, n);

    cspbo_pack *p = new (std::nothrow) cspbo_pack();
    if (!p) { set_error("pack_create: host allocation failed"); return CSPBO_E_NOMEM; }
    p->n = n; p->d = d; p->nc = nc; p->m = m; p->nv = 1 + nc; p->ks = ks;
    p->tile_doubles = (long long)p->nv * ks * 32;
    int rc = CSPBO_OK;
    long long *d_dest = nullptr;
    double *d_x = nullptr, *d_dx = nullptr;
    std::vector<long long> dest((size_t)std::max(n, 1), -1);
    std::vector<int> slot_group;

    // ---- host plan -------------------------------------------------------------------
    {
        std::vector<int> sp;
        for (int r = row_start; r < row_end; r++) {
            if (inds[r] < 0 || inds[r] >= m) {
                set_error("pack_create: group id %d of row %d outside [0,%d)", inds[r], r, m);
                delete p; return CSPBO_E_ARG;
            }
            sp.push_back(ele[r]);
        }
        std::sort(sp.begin(), sp.end());
        sp.erase(std::unique(sp.begin(), sp.end()), sp.end());
        p->species = sp;
        const int ns = (int)sp.size();
        p->blk0.assign(ns + 1, 0);
        p->rows_per_species.assign(ns, 0);
        p->h_tile_start.clear();
        p->h_tile_start.push_back(0);
        std::vector<int> count((size_t)std::max(m, 1));
        std::vector<int> order;
        std::vector<long long> first_dest((size_t)std::max(m, 1));  // dest of rank 0 of group g (current species)
        std::vector<int> rank((size_t)std::max(m, 1));
        long long tiles = 0;
        for (int e = 0; e < ns; e++) {
            std::fill(count.begin(), count.end(), 0);
            for (int r = row_start; r < row_end; r++)
                if (ele[r] == sp[e]) count[inds[r]]++;
            order.clear();
            for (int g = 0; g < m; g++)
                if (count[g] > 0) { order.push_back(g); p->rows_per_species[e] += count[g]; }
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return count[a] > count[b]; });
            const int nb = ((int)order.size() + kSlots - 1) / kSlots;
            for (int b = 0; b < nb; b++) {
                const int T = count[order[(size_t)b * kSlots]];
                for (int s = 0; s < kSlots; s++) {
                    const size_t oi = (size_t)b * kSlots + s;
                    if (oi < order.size()) {
                        slot_group.push_back(order[oi]);
                        first_dest[order[oi]] = tiles * 8 + s;
                    } else {
                        slot_group.push_back(-1);
                    }
                }
                tiles += T;
                p->h_tile_start.push_back((int)tiles);
            }
            p->blk0[e + 1] = p->blk0[e] + nb;
            std::fill(rank.begin(), rank.end(), 0);
            for (int r = row_start; r < row_end; r++)
                if (ele[r] == sp[e]) dest[r] = first_dest[inds[r]] + 8LL * (rank[inds[r]]++);
        }
        p->n_tiles = tiles;
        p->n_blocks = p->blk0[ns];
        if (tiles > 0x7fffffffLL / 8) { set_error("pack_create: too many tiles"); delete p; return CSPBO_E_ARG; }
    }

    // ---- device allocations ------------------------------------------------------------
#define PACK_CUDA(call)                                                                         \
    do {                                                                                        \
        cudaError_t e__ = (call);                                                               \
        if (e__ != cudaSuccess) {                                                               \
            set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));    \
            rc = (e__ == cudaErrorMemoryAllocation) ? CSPBO_E_NOMEM                             \
                 : (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver) ? CSPBO_E_NODEVICE \
                                                                                    : CSPBO_E_CUDA; \
            goto fail;                                                                          \
        }                                                                                       \
    } while (0)

    PACK_CUDA(cudaGetDevice(&p->device));
    {
        const size_t tb = (size_t)std::max<long long>(p->n_tiles, 1) * p->tile_doubles * sizeof(double);
        const size_t nbk = (size_t)std::max<long long>(p->n_blocks, 1);
        PACK_CUDA(cudaMalloc(&p->d_tiles, tb));
        PACK_CUDA(cudaMemsetAsync(p->d_tiles, 0, tb));
        PACK_CUDA(cudaMalloc(&p->d_tile_start, (nbk + 1) * sizeof(int)));
        PACK_CUDA(cudaMalloc(&p->d_slot_group, nbk * 8 * sizeof(int)));
        PACK_CUDA(cudaMalloc(&p->d_tile_valid, (size_t)std::max<long long>(p->n_tiles, 1) * sizeof(unsigned)));
        PACK_CUDA(cudaMemsetAsync(p->d_tile_valid, 0, (size_t)std::max<long long>(p->n_tiles, 1) * sizeof(unsigned)));
        p->bytes = (long long)(tb + (nbk + 1) * 4 + nbk * 32 + std::max<long long>(p->n_tiles, 1) * 4);
        PACK_CUDA(cudaMemcpyAsync(p->d_tile_start, p->h_tile_start.data(), p->h_tile_start.size() * sizeof(int),
                                  cudaMemcpyHostToDevice));
        if (!slot_group.empty())
            PACK_CUDA(cudaMemcpyAsync(p->d_slot_group, slot_group.data(), slot_group.size() * sizeof(int),
                                      cudaMemcpyHostToDevice));
    }
    if (n > 0 && p->n_tiles > 0) {
        PACK_CUDA(cudaMalloc(&d_dest, (size_t)n * sizeof(long long)));
        PACK_CUDA(cudaMemcpyAsync(d_dest, dest.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice));
        const double *xs = x, *dxs = dxdr;
        if (mem == CSPBO_MEM_HOST) {
            PACK_CUDA(cudaMalloc(&d_x, (size_t)n * d * sizeof(double)));
            PACK_CUDA(cudaMemcpyAsync(d_x, x, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice));
            xs = d_x;
            if (nc > 0) {
                PACK_CUDA(cudaMalloc(&d_dx, (size_t)n * d * nc * sizeof(double)));
                PACK_CUDA(cudaMemcpyAsync(d_dx, dxdr, (size_t)n * d * nc * sizeof(double), cudaMemcpyHostToDevice));
                dxs = d_dx;
            }
        }
        const int wpb = 8;
        pack_rows_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32>>>(n, d, nc, ks, p->nv, xs, dxs, d_dest,
                                                                        p->d_tiles, p->d_tile_valid, p->tile_doubles);
        g_launches++;
        PACK_CUDA(cudaGetLastError());
    }
    PACK_CUDA(cudaStreamSynchronize(0));  // dest / slot_group host buffers die with this scope
    cudaFree(d_dest); cudaFree(d_x); cudaFree(d_dx);
    *out = p;
    return CSPBO_OK;
fail:
    cudaFree(d_dest); cudaFree(d_x); cudaFree(d_dx);
    cspbo_pack_destroy(p);
    return rc;
#undef PACK_CUDA
}

extern "C" int cspbo_pack_destroy(cspbo_pack *p)
{
    if (!p) return CSPBO_OK;
    cudaFree(p->d_tiles);
    cudaFree(p->d_tile_start);
    cudaFree(p->d_slot_group);
    cudaFree(p->d_tile_valid);
    delete p;
    return CSPBO_OK;
}

extern "C" int cspbo_pack_groups(const cspbo_pack *p) { return p ? p->m : -1; }
extern "C" long long cspbo_pack_bytes(const cspbo_pack *p) { return p ? p->bytes : -1; }

extern "C" long long cspbo_pack_pair_count(const cspbo_pack *a, const cspbo_pack *b)
{
    if (!a || !b) return -1;
    long long pairs = 0;
    for (size_t i = 0; i < a->species.size(); i++)
        for (size_t j = 0; j < b->species.size(); j++)
            if (a->species[i] == b->species[j]) pairs += a->rows_per_species[i] * b->rows_per_species[j];
    return pairs;
}
