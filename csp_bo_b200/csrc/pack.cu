// pack.cu -- "pack" step of the hot path: normalise every stacked descriptor row, project its
// Cartesian derivative onto the tangent space of the unit sphere, and scatter both into the
// species-bucketed, group-interleaved tile layout the DMMA tile kernels consume (common.h).
//
// Why this layout: with u = x/|x| and A^ = (I - u u^T) dxdr / |x| every
// per-pair quantity of the reference's d x d loop (dot_kernel.cpp:270-323) is one of the 16 dot
// products [u_a; A^_a] . [u_b; A^_b]^T:
//     c = u_a.u_b,  P_k = A^_a[:,k].u_b,  Q_l = u_a.A^_b[:,l],  T_kl = A^_a[:,k].A^_b[:,l]
// Groups are sorted once (lexicographically by their per-species row counts) and cut into blocks of
// 8; inside a block the rows are bucketed by species (pairs of different species contribute
// nothing, dot_kernel.cpp:255) and interleaved so that slot r of every tile of the block is group
// r: a thread's MMA accumulator fragment then always belongs to one fixed (group_i, group_j) pair
// and the reference's scatter-add `pout[(k+_i*3)*x2i+_j)*3+l] +=` becomes a register accumulation
// that is written exactly once.
#include <algorithm>
#include <cstring>
#include <numeric>

#include "common.h"

namespace cspbo {

// Four stacked rows per warp, eight lanes per row (lane l8 owns the features l8, l8+8, ...): the eight lanes of a row
// store 64 contiguous bytes of a tile per (vector, feature octet) -- the fragment order of common.h keeps a slot's
// features f..f+7 of one vector together -- the derivative columns are read three at a time as contiguous 24-byte
// triplets, and a warp has four rows' worth of loads in flight.
constexpr int kPackRowsPerWarp = 4;

__global__ void pack_rows_kernel(int n, int d, int nc, int ks, int nv,
                                 const double *__restrict__ x, const double *__restrict__ dxdr,
                                 const long long *__restrict__ dest, double *__restrict__ tiles,
                                 unsigned *__restrict__ tile_valid, long long tile_doubles, double norm_shift)
{
    const int lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * kPackRowsPerWarp + sub;
    const long long dst = row < n ? dest[row] : -1;
    const bool live = dst >= 0;                      // the same for the eight lanes of a row
    const double *xr = x + (live ? row : 0) * d;

    double ss = 0.0;
    if (live)
        for (int f = l8; f < d; f += 8) {
            const double v = xr[f];
            ss += v * v;
        }
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    // norm_shift = 0: the C++ path (rows with |x| <= 1e-8 are skipped, dot_kernel.cpp:25);
    // norm_shift = eps: the NumPy single-pair functions used by diag(), which add eps to every norm and
    // skip nothing (Dot_mb.py:188-191,206,212)
    const double nrm = sqrt(ss) + norm_shift;
    const bool ok = live && (norm_shift != 0.0 || nrm > 1e-8);
    const double inv = ok ? 1.0 / nrm : 0.0;
    double *t = tiles + (ok ? (dst >> 3) : 0) * tile_doubles;
    const int slot = (int)(dst & 7);
    const int half = ks >> 1;
    auto at = [&](int v, int f) { return (((long long)(v * half + (f >> 3)) * 32) + slot * 4 + (f & 3)) * 2 + ((f >> 2) & 1); };
    if (ok)
        for (int f = l8; f < d; f += 8) t[at(0, f)] = xr[f] * inv;
    if (nc > 0) {
        const double *dr = dxdr + (live ? row : 0) * (long long)d * nc;
        for (int k0 = 0; k0 < nc; k0 += 3) {
            double al[3] = {0.0, 0.0, 0.0};  // alpha_k = dxdr[:,k] . u
            if (ok)
                for (int f = l8; f < d; f += 8) {
                    const double u = xr[f] * inv;
#pragma unroll
                    for (int c = 0; c < 3; c++) al[c] += dr[(long long)f * nc + k0 + c] * u;
                }
#pragma unroll
            for (int c = 0; c < 3; c++)
#pragma unroll
                for (int o = 4; o > 0; o >>= 1) al[c] += __shfl_xor_sync(0xffffffffu, al[c], o);
            if (ok)
                for (int f = l8; f < d; f += 8) {
                    const double u = xr[f] * inv;
#pragma unroll
                    for (int c = 0; c < 3; c++) t[at(1 + k0 + c, f)] = (dr[(long long)f * nc + k0 + c] - al[c] * u) * inv;
                }
        }
    }
    if (ok && l8 == 0) atomicOr(tile_valid + (dst >> 3), 1u << slot);
}

}  // namespace cspbo

using namespace cspbo;

extern "C" int cspbo_pack_create(int n, int d, int nc, int m, int row_start, int row_end,
                                 const double *x, const double *dxdr, const int *ele, const int *inds,
                                 int mem, cspbo_pack **out)
{
    return cspbo_pack_create_ex(n, d, nc, m, row_start, row_end, x, dxdr, ele, inds, mem, 0.0, 1, out);
}

extern "C" int cspbo_pack_create_eps(int n, int d, int nc, int m, int row_start, int row_end,
                                     const double *x, const double *dxdr, const int *ele, const int *inds,
                                     int mem, double norm_shift, cspbo_pack **out)
{
    return cspbo_pack_create_ex(n, d, nc, m, row_start, row_end, x, dxdr, ele, inds, mem, norm_shift, 1, out);
}

extern "C" int cspbo_pack_create_ex(int n, int d, int nc, int m, int row_start, int row_end,
                                    const double *x, const double *dxdr, const int *ele, const int *inds,
                                    int mem, double norm_shift, int row_classes, cspbo_pack **out)
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
    if (ks < 0) { set_error("pack_create: descriptor length d=%d out of range", d); return CSPBO_E_ARG; }
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

    // ---- host-resident inputs: start their upload first, the DMA runs while the host plans the layout ----
    bool uploading = false;
    if (mem == CSPBO_MEM_HOST && n > 0) {
        if ((rc = scratch_get(SCRATCH_X, (size_t)n * d * sizeof(double), (void **)&d_x)) != CSPBO_OK) { delete p; return rc; }
        if (nc > 0 && (rc = scratch_get(SCRATCH_DX, (size_t)n * d * nc * sizeof(double), (void **)&d_dx)) != CSPBO_OK) { delete p; return rc; }
        cudaError_t e = cudaMemcpyAsync(d_x, x, (size_t)n * d * sizeof(double), cudaMemcpyHostToDevice);
        if (e == cudaSuccess && nc > 0) e = cudaMemcpyAsync(d_dx, dxdr, (size_t)n * d * nc * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            set_error("pack_create: upload of the descriptors failed: %s", cudaGetErrorString(e));
            cudaStreamSynchronize(0);
            delete p;
            return CSPBO_E_CUDA;
        }
        uploading = true;
    }
    // every early return below must wait for the upload: the caller may free x / dxdr as soon as we return
    auto plan_fail = [&](int code) { if (uploading) cudaStreamSynchronize(0); delete p; return code; };

    // ---- host plan -------------------------------------------------------------------
    {
        // species, most populous first
        std::vector<int> sp;
        for (int r = row_start; r < row_end; r++) {
            if (inds[r] < 0 || inds[r] >= m) {
                set_error("pack_create: group id %d of row %d outside [0,%d)", inds[r], r, m);
                return plan_fail(CSPBO_E_ARG);
            }
            sp.push_back(ele[r]);
        }
        std::sort(sp.begin(), sp.end());
        std::vector<std::pair<long long, int>> hist;  // (-rows, species id)
        for (size_t i = 0; i < sp.size();) {
            size_t j = i;
            while (j < sp.size() && sp[j] == sp[i]) j++;
            hist.push_back({-(long long)(j - i), sp[i]});
            i = j;
        }
        std::sort(hist.begin(), hist.end());
        const int ns = (int)hist.size();
        p->species.resize(ns);
        p->rows_per_species.resize(ns);
        for (int e = 0; e < ns; e++) { p->species[e] = hist[e].second; p->rows_per_species[e] = -hist[e].first; }
        auto species_index = [&](int id) {
            for (int e = 0; e < ns; e++) if (p->species[e] == id) return e;
            return -1;
        };
        // per-group, per-species row counts
        const int nsz = std::max(ns, 1);
        std::vector<int> count((size_t)std::max(m, 1) * nsz, 0);
        std::vector<int> erow((size_t)std::max(n, 1), -1);
        for (int r = row_start; r < row_end; r++) {
            erow[r] = species_index(ele[r]);
            count[(size_t)inds[r] * nsz + erow[r]]++;
        }
        // one common group order: lexicographic on the count tuples, largest first, so that the 8
        // groups of a block have (nearly) equal row counts in every species and tiles carry no padding
        auto by_counts = [&](int a, int b) {
            const int *ca = &count[(size_t)a * nsz], *cb = &count[(size_t)b * nsz];
            for (int e = 0; e < ns; e++)
                if (ca[e] != cb[e]) return ca[e] > cb[e];
            return false;
        };
        // row classes (host-bound builds): groups are first split by g % row_classes and blocks never
        // straddle a class, so the rows of one class sit at a uniform pitch in the caller's matrix and
        // a finished class is copied back with ONE strided (2-D) copy (api.cu, pipelined path)
        const int ncls = std::max(1, std::min(row_classes, std::max(m, 1)));
        std::vector<int> slots;
        slots.reserve((size_t)m + 8 * (size_t)ncls);
        p->row_classes = ncls;
        p->h_class_blk0.assign((size_t)ncls + 1, 0);
        std::vector<int> order;
        for (int c = 0; c < ncls; c++) {
            order.clear();
            for (int g = c; g < m; g += ncls) order.push_back(g);
            std::stable_sort(order.begin(), order.end(), by_counts);
            p->h_class_blk0[(size_t)c] = (int)(slots.size() / kSlots);
            slots.insert(slots.end(), order.begin(), order.end());
            while (slots.size() % kSlots) slots.push_back(-1);
        }
        const int nb = (int)(slots.size() / kSlots);
        p->h_class_blk0[(size_t)ncls] = nb;
        p->n_blocks = nb;
        p->h_slot_group.assign((size_t)std::max(nb, 1) * kSlots, -1);
        std::copy(slots.begin(), slots.end(), p->h_slot_group.begin());
        p->h_tile_start.assign((size_t)nb * nsz + 1, 0);
        std::vector<long long> first_dest((size_t)std::max(m, 1) * nsz, -1);
        long long tiles = 0;
        for (int b = 0; b < nb; b++) {
            for (int e = 0; e < ns; e++) {
                int T = 0;
                for (int s = 0; s < kSlots; s++) {
                    const int g = p->h_slot_group[(size_t)b * kSlots + s];
                    if (g >= 0) {
                        T = std::max(T, count[(size_t)g * nsz + e]);
                        first_dest[(size_t)g * nsz + e] = tiles * 8 + s;
                    }
                }
                p->h_tile_start[(size_t)b * nsz + e] = (int)tiles;
                tiles += T;
            }
        }
        p->h_tile_start[(size_t)nb * nsz] = (int)tiles;
        if (ns == 0) p->h_tile_start.assign((size_t)nb + 1, 0);
        std::vector<int> rank((size_t)std::max(m, 1) * nsz, 0);
        for (int r = row_start; r < row_end; r++) {
            const size_t k = (size_t)inds[r] * nsz + erow[r];
            dest[r] = first_dest[k] + 8LL * (rank[k]++);
        }
        p->n_tiles = tiles;
        if (tiles > 0x7fffffffLL / 8) { set_error("pack_create: too many tiles"); return plan_fail(CSPBO_E_ARG); }
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
        const size_t nts = p->h_tile_start.size();
        if ((rc = dev_alloc_cached(tb, (void **)&p->d_tiles)) != CSPBO_OK) goto fail;
        PACK_CUDA(cudaMemsetAsync(p->d_tiles, 0, tb));
        if ((rc = dev_alloc_cached(nts * sizeof(int), (void **)&p->d_tile_start)) != CSPBO_OK) goto fail;
        if ((rc = dev_alloc_cached(nbk * 8 * sizeof(int), (void **)&p->d_slot_group)) != CSPBO_OK) goto fail;
        if ((rc = dev_alloc_cached((size_t)std::max<long long>(p->n_tiles, 1) * sizeof(unsigned), (void **)&p->d_tile_valid)) != CSPBO_OK) goto fail;
        PACK_CUDA(cudaMemsetAsync(p->d_tile_valid, 0, (size_t)std::max<long long>(p->n_tiles, 1) * sizeof(unsigned)));
        p->bytes = (long long)(tb + nts * 4 + nbk * 32 + std::max<long long>(p->n_tiles, 1) * 4);
        PACK_CUDA(cudaMemcpyAsync(p->d_tile_start, p->h_tile_start.data(), p->h_tile_start.size() * sizeof(int),
                                  cudaMemcpyHostToDevice));
        PACK_CUDA(cudaMemcpyAsync(p->d_slot_group, p->h_slot_group.data(), p->h_slot_group.size() * sizeof(int),
                                  cudaMemcpyHostToDevice));
    }
    if (n > 0 && p->n_tiles > 0) {
        if ((rc = scratch_get(SCRATCH_DEST, (size_t)n * sizeof(long long), (void **)&d_dest)) != CSPBO_OK) goto fail;
        PACK_CUDA(cudaMemcpyAsync(d_dest, dest.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice));
        const double *xs = uploading ? d_x : x, *dxs = (uploading && nc > 0) ? d_dx : dxdr;   // uploaded above (stream order)
        const int rpb = 8 * kPackRowsPerWarp;   // 8 warps x 4 rows per CTA
        pack_rows_kernel<<<(unsigned)((n + rpb - 1) / rpb), 256>>>(n, d, nc, ks, p->nv, xs, dxs, d_dest,
                                                                        p->d_tiles, p->d_tile_valid, p->tile_doubles, norm_shift);
        g_launches++;
        PACK_CUDA(cudaGetLastError());
    }
    PACK_CUDA(cudaStreamSynchronize(0));  // dest host buffer dies with this scope; scratch is reusable again
    *out = p;
    return CSPBO_OK;
fail:
    if (uploading) cudaStreamSynchronize(0);
    cspbo_pack_destroy(p);
    return rc;
#undef PACK_CUDA
}

extern "C" int cspbo_pack_destroy(cspbo_pack *p)
{
    if (!p) return CSPBO_OK;
    dev_free_cached(p->d_tiles);
    dev_free_cached(p->d_tile_start);
    dev_free_cached(p->d_slot_group);
    dev_free_cached(p->d_tile_valid);
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
