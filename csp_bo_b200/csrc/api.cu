// api.cu -- the C ABI of libcspbo_b200.so (include/cspbo_b200.h): error reporting, the resident
// block builder and the 13 reference-compatible entry points built on top of it.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.h"

namespace cspbo {

static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int build_block_device(const cspbo_block_desc &d, const cspbo_pack *a, const cspbo_pack *b, int voffA, int voffB,
                       double *out, double *out_grad, long long si, long long sk, long long sj, long long sl,
                       int accumulate, cudaStream_t st, int rank = 0, int nranks = 1, double *const *outs = nullptr,
                       int a_begin = 0, int a_end = -1, int diag = 0, int sb_blocks = 1, int zigzag = 0,
                       int packed_cols = 0, const struct PanelArgs *panel = nullptr, const struct PassArgs *passes = nullptr,
                       const struct TraceArgs *trace = nullptr, const int *sb_perm = nullptr, const int *sb_inv = nullptr);
struct PanelArgs { long long ld; int nbw; long long row0_a, row0_b; };
struct PassArgs { int npass, side; long long off; };
struct TraceArgs { const double *alpha, *w; double *partial; long long ld, row0, col0; int packed; long long slab0; };

// deterministic sum of the per-warp partial traces (fixed tree, one CTA)
__global__ void trace_sum_kernel(const double *__restrict__ part, long long n, double *__restrict__ out, int accumulate)
{
    __shared__ double sh[32];
    double acc = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        acc = threadIdx.x < (blockDim.x >> 5) ? sh[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (threadIdx.x == 0) *out = accumulate ? *out + acc : acc;
    }
}
int sharded_superblocks(int n_blocks, int rank, int nranks, int sb_blocks, int zigzag);

struct DevBuf {
    double *p = nullptr;
    ~DevBuf() { cudaFree(p); }
};

static std::mutex g_scratch_mu;
static void *g_scratch[64][SCRATCH_SLOTS] = {};
static size_t g_scratch_size[64][SCRATCH_SLOTS] = {};

int scratch_get(int slot, size_t bytes, void **ptr)
{
    int dev = 0;
    CSPBO_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || slot < 0 || slot >= SCRATCH_SLOTS) { set_error("scratch_get: bad device/slot"); return CSPBO_E_ARG; }
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    if (g_scratch_size[dev][slot] < bytes) {
        if (g_scratch[dev][slot]) {
            CSPBO_CUDA(cudaDeviceSynchronize());   // nobody may still be using the old buffer
            cudaFree(g_scratch[dev][slot]);
            g_scratch[dev][slot] = nullptr;
            g_scratch_size[dev][slot] = 0;
        }
        const size_t want = bytes + bytes / 8;   // a little headroom against regrowth
        cudaError_t e = cudaMalloc(&g_scratch[dev][slot], want);
        if (e != cudaSuccess) {
            e = cudaMalloc(&g_scratch[dev][slot], bytes);
            if (e != cudaSuccess) { cudaGetLastError(); set_error("scratch_get: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return CSPBO_E_NOMEM; }
            g_scratch_size[dev][slot] = bytes;
        } else {
            g_scratch_size[dev][slot] = want;
        }
    }
    *ptr = g_scratch[dev][slot];
    return CSPBO_OK;
}

}  // namespace cspbo

using namespace cspbo;

extern "C" const char *cspbo_last_error(void) { return g_err; }
extern "C" int cspbo_version(void) { return 100; }
extern "C" long long cspbo_launch_count(void) { return g_launches.load(); }

extern "C" int cspbo_set_device(int device)
{
    CSPBO_CUDA(cudaSetDevice(device));
    return CSPBO_OK;
}

extern "C" int cspbo_device_count(int *count)
{
    if (!count) { set_error("device_count: null"); return CSPBO_E_ARG; }
    *count = 0;
    CSPBO_CUDA(cudaGetDeviceCount(count));
    return CSPBO_OK;
}

namespace cspbo {
struct CachedBlock { void *ptr; size_t size; int dev; };
static std::vector<CachedBlock> g_free_blocks;               // released, reusable
static std::vector<CachedBlock> g_live_blocks;               // handed out
static const size_t kCacheLimit = (size_t)2 << 30;

int dev_alloc_cached(size_t bytes, void **ptr)
{
    int dev = 0;
    CSPBO_CUDA(cudaGetDevice(&dev));
    bytes = std::max<size_t>(bytes, 256);
    {
        std::lock_guard<std::mutex> lk(g_scratch_mu);
        int best = -1;
        for (size_t i = 0; i < g_free_blocks.size(); i++) {
            const CachedBlock &b = g_free_blocks[i];
            if (b.dev == dev && b.size >= bytes && b.size <= 2 * bytes + 4096 &&
                (best < 0 || b.size < g_free_blocks[(size_t)best].size))
                best = (int)i;
        }
        if (best >= 0) {
            CachedBlock b = g_free_blocks[(size_t)best];
            g_free_blocks.erase(g_free_blocks.begin() + best);
            g_live_blocks.push_back(b);
            *ptr = b.ptr;
            return CSPBO_OK;
        }
    }
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cspbo_release_scratch();          // give cached memory back and retry once
        e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e)); return CSPBO_E_NOMEM; }
    }
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    g_live_blocks.push_back({q, bytes, dev});
    *ptr = q;
    return CSPBO_OK;
}

void dev_free_cached(void *ptr)
{
    if (!ptr) return;
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    for (size_t i = 0; i < g_live_blocks.size(); i++)
        if (g_live_blocks[i].ptr == ptr) {
            CachedBlock b = g_live_blocks[i];
            g_live_blocks.erase(g_live_blocks.begin() + (long)i);
            size_t held = 0;
            for (const CachedBlock &f : g_free_blocks) held += (f.dev == b.dev) ? f.size : 0;
            if (held + b.size <= kCacheLimit) g_free_blocks.push_back(b);
            else cudaFree(b.ptr);
            return;
        }
    cudaFree(ptr);
}
}  // namespace cspbo

extern "C" int cspbo_release_scratch(void)
{
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_scratch_mu);
    for (const CachedBlock &b : g_free_blocks) cudaFree(b.ptr);
    g_free_blocks.clear();
    int cur = 0;
    cudaGetDevice(&cur);
    for (int dev = 0; dev < 64; dev++)
        for (int s = 0; s < SCRATCH_SLOTS; s++)
            if (g_scratch[dev][s]) {
                cudaSetDevice(dev);
                cudaFree(g_scratch[dev][s]);
                g_scratch[dev][s] = nullptr;
                g_scratch_size[dev][s] = 0;
            }
    cudaSetDevice(cur);
    return CSPBO_OK;
}

extern "C" int cspbo_synchronize(void)
{
    CSPBO_CUDA(cudaDeviceSynchronize());
    return CSPBO_OK;
}

extern "C" int cspbo_block_build(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b,
                                 double *out, double *out_grad, int mem, int accumulate)
{
    if (!desc || !a || !b || !out) { set_error("block_build: null argument"); return CSPBO_E_ARG; }
    const cspbo_block_desc &d = *desc;
    if (d.family != CSPBO_DOT && d.family != CSPBO_RBF) { set_error("block_build: bad family %d", d.family); return CSPBO_E_ARG; }
    if (d.block < CSPBO_KEE || d.block > CSPBO_KFF) { set_error("block_build: bad block %d", d.block); return CSPBO_E_ARG; }
    const bool grad = d.family == CSPBO_RBF && d.with_grad;
    if (grad && !out_grad) { set_error("block_build: with_grad needs out_grad"); return CSPBO_E_ARG; }
    if (grad && d.stress) { set_error("block_build: with_grad and stress are exclusive"); return CSPBO_E_ARG; }
    if (d.family == CSPBO_RBF && !(d.l > 0.0)) { set_error("block_build: RBF needs l > 0"); return CSPBO_E_ARG; }
    if (d.symmetric && (a != b || d.stress || d.block == CSPBO_KEF)) {
        set_error("block_build: symmetric needs a == b, no stress, K_EE or K_FF");
        return CSPBO_E_ARG;
    }
    const long long m1 = a->m, m2 = b->m;
    long long si, sk, sj, sl, n_out, n_grad = 0;
    int passesA = 1, passesB = 1;
    if (d.block == CSPBO_KEE) {
        if (a->nc != 0 || b->nc != 0) { set_error("block_build: K_EE needs two energy packs"); return CSPBO_E_ARG; }
        si = m2; sk = 0; sj = 1; sl = 0; n_out = m1 * m2; n_grad = n_out;
    } else if (d.block == CSPBO_KEF) {
        const int nc = d.stress ? 9 : 3;
        if (a->nc != 0 || b->nc < nc) { set_error("block_build: K_EF needs (energy, force[nc>=%d]) packs", nc); return CSPBO_E_ARG; }
        si = m2 * nc; sk = 0; sj = nc; sl = 1; n_out = m1 * m2 * nc; n_grad = m1 * m2 * 3;
        passesB = nc / 3;
    } else {
        const int nc1 = d.stress ? 9 : 3;
        if (a->nc < nc1 || b->nc < 3) { set_error("block_build: K_FF needs (force[nc>=%d], force) packs", nc1); return CSPBO_E_ARG; }
        si = nc1 * m2 * 3; sk = m2 * 3; sj = 3; sl = 1; n_out = m1 * nc1 * m2 * 3; n_grad = m1 * 3 * m2 * 3;
        passesA = nc1 / 3;
    }
    if (n_out == 0) return CSPBO_OK;

    double *dout = out, *dgrad = out_grad;
    if (mem == CSPBO_MEM_HOST) {
        void *sp = nullptr;
        int rcs = scratch_get(SCRATCH_OUT, (size_t)n_out * sizeof(double), &sp);
        if (rcs) return rcs;
        dout = static_cast<double *>(sp);
        if (accumulate) CSPBO_CUDA(cudaMemcpyAsync(dout, out, (size_t)n_out * sizeof(double), cudaMemcpyHostToDevice));
        if (grad) {
            rcs = scratch_get(SCRATCH_GRAD, (size_t)n_grad * sizeof(double), &sp);
            if (rcs) return rcs;
            dgrad = static_cast<double *>(sp);
            if (accumulate) CSPBO_CUDA(cudaMemcpyAsync(dgrad, out_grad, (size_t)n_grad * sizeof(double), cudaMemcpyHostToDevice));
        }
    }
    // every (group_i, group_j) output block is owned by exactly one warp, which writes it once:
    // no zero-initialisation is needed in overwrite mode

    // Host-bound K_FF: cut the A blocks into row slabs and copy every finished slab back on a second
    // stream while the next one is being computed (a slab's rows are final once its own kernel has
    // run: mirrored tiles only ever land in rows of the same or later slabs).
    const bool pipelined = mem == CSPBO_MEM_HOST && d.block == CSPBO_KFF && a->n_blocks >= 64 &&
                           (size_t)n_out * sizeof(double) >= ((size_t)64 << 20);
    if (pipelined) {
        const bool classes = a->row_classes > 1;
        const int nslab = classes ? a->row_classes : 12;
        const int nb = (int)a->n_blocks;
        cudaStream_t cs = nullptr;
        cudaEvent_t ev = nullptr;
        CSPBO_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        CSPBO_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CSPBO_CUDA(cudaStreamSynchronize(0));   // upload of an accumulated-into pout has landed
        int rc = CSPBO_OK;
        std::vector<int> groups;
        const long long rowlen = si, rowlen_g = 3LL * m2 * 3;
        for (int s = 0; s < nslab && rc == CSPBO_OK; s++) {
            const int b0 = classes ? a->h_class_blk0[(size_t)s] : (int)((long long)nb * s / nslab);
            const int b1 = classes ? a->h_class_blk0[(size_t)s + 1] : (int)((long long)nb * (s + 1) / nslab);
            if (b1 <= b0) continue;
            {
                // stress: the three row triplets are the gridDim.y passes of ONE launch
                const PassArgs ps = {passesA, 0, 3 * sk};
                rc = build_block_device(d, a, b, 0, 0, dout, dgrad, si, sk, sj, sl, accumulate, 0, 0, 1, nullptr, b0, b1,
                                        0, 1, 0, 0, nullptr, &ps);
            }
            if (rc) break;
            cudaEventRecord(ev, 0);
            cudaStreamWaitEvent(cs, ev, 0);
            if (classes) {
                // rows of class s are the groups s, s+S, s+2S, ...: one strided copy
                const long long cnt = (m1 - s + nslab - 1) / nslab;
                if (cnt > 0) {
                    cudaMemcpy2DAsync(out + (long long)s * rowlen, (size_t)nslab * rowlen * sizeof(double),
                                      dout + (long long)s * rowlen, (size_t)nslab * rowlen * sizeof(double),
                                      (size_t)rowlen * sizeof(double), (size_t)cnt, cudaMemcpyDeviceToHost, cs);
                    if (grad)
                        cudaMemcpy2DAsync(out_grad + (long long)s * rowlen_g, (size_t)nslab * rowlen_g * sizeof(double),
                                          dgrad + (long long)s * rowlen_g, (size_t)nslab * rowlen_g * sizeof(double),
                                          (size_t)rowlen_g * sizeof(double), (size_t)cnt, cudaMemcpyDeviceToHost, cs);
                }
                continue;
            }
            groups.clear();
            for (int blk = b0; blk < b1; blk++)
                for (int r = 0; r < 8; r++) {
                    const int g = a->h_slot_group[(size_t)blk * 8 + r];
                    if (g >= 0) groups.push_back(g);
                }
            std::sort(groups.begin(), groups.end());
            for (size_t i = 0; i < groups.size();) {   // merge runs of consecutive group ids into one copy
                size_t j = i + 1;
                while (j < groups.size() && groups[j] == groups[j - 1] + 1) j++;
                const long long g0 = groups[i], cnt = (long long)(j - i);
                cudaMemcpyAsync(out + g0 * rowlen, dout + g0 * rowlen, (size_t)(cnt * rowlen) * sizeof(double),
                                cudaMemcpyDeviceToHost, cs);
                if (grad)
                    cudaMemcpyAsync(out_grad + g0 * rowlen_g, dgrad + g0 * rowlen_g,
                                    (size_t)(cnt * rowlen_g) * sizeof(double), cudaMemcpyDeviceToHost, cs);
                i = j;
            }
        }
        cudaError_t e1 = cudaStreamSynchronize(cs), e2 = cudaStreamSynchronize(0), e3 = cudaGetLastError();
        cudaEventDestroy(ev);
        cudaStreamDestroy(cs);
        if (rc) return rc;
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
            set_error("pipelined K_FF copy-back failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
            return CSPBO_E_CUDA;
        }
        return CSPBO_OK;
    }

    {
        // stress: the 9 derivative columns are three independent triplets, run as the gridDim.y passes of ONE launch
        // (3x the CTAs of a plain build: the K* builds of an MD step are grid-bound, dot_kernel.cpp:336-506)
        const PassArgs ps = (d.block == CSPBO_KFF) ? PassArgs{passesA, 0, 3 * sk} : PassArgs{passesB, 1, 3 * sl};
        int rc = build_block_device(d, a, b, 0, 0, dout, dgrad, si, sk, sj, sl, accumulate, 0, 0, 1, nullptr, 0, -1, 0, 1, 0, 0,
                                    nullptr, &ps);
        if (rc) return rc;
    }

    if (mem == CSPBO_MEM_HOST) {
        CSPBO_CUDA(cudaMemcpyAsync(out, dout, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost));
        if (grad) CSPBO_CUDA(cudaMemcpyAsync(out_grad, dgrad, (size_t)n_grad * sizeof(double), cudaMemcpyDeviceToHost));
        CSPBO_CUDA(cudaStreamSynchronize(0));
    }
    return CSPBO_OK;
}

extern "C" int cspbo_block_build_into(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b,
                                      double *K, long long ldk, long long row0, long long col0)
{
    if (!desc || !a || !b || !K) { set_error("block_build_into: null argument"); return CSPBO_E_ARG; }
    const cspbo_block_desc &d = *desc;
    if (d.family != CSPBO_DOT && d.family != CSPBO_RBF) { set_error("block_build_into: bad family"); return CSPBO_E_ARG; }
    if (d.stress || d.with_grad) { set_error("block_build_into: plain blocks only (no stress / grad)"); return CSPBO_E_ARG; }
    if (d.family == CSPBO_RBF && !(d.l > 0.0)) { set_error("block_build_into: RBF needs l > 0"); return CSPBO_E_ARG; }
    long long si, sk, sj, sl, cols;
    if (d.block == CSPBO_KEE) {
        if (a->nc != 0 || b->nc != 0) { set_error("block_build_into: K_EE needs two energy packs"); return CSPBO_E_ARG; }
        si = ldk; sk = 0; sj = 1; sl = 0; cols = b->m;
    } else if (d.block == CSPBO_KEF) {
        if (a->nc != 0 || b->nc < 3) { set_error("block_build_into: K_EF needs (energy, force) packs"); return CSPBO_E_ARG; }
        si = ldk; sk = 0; sj = 3; sl = 1; cols = 3LL * b->m;
    } else if (d.block == CSPBO_KFF) {
        if (a->nc < 3 || b->nc < 3) { set_error("block_build_into: K_FF needs two force packs"); return CSPBO_E_ARG; }
        si = 3 * ldk; sk = ldk; sj = 3; sl = 1; cols = 3LL * b->m;
    } else { set_error("block_build_into: bad block"); return CSPBO_E_ARG; }
    if (row0 < 0 || col0 < 0 || col0 + cols > ldk) { set_error("block_build_into: block does not fit (ldk %lld)", ldk); return CSPBO_E_ARG; }
    if (d.symmetric && (a != b || d.block == CSPBO_KEF || row0 != col0)) {
        set_error("block_build_into: symmetric needs a == b on the diagonal of K");
        return CSPBO_E_ARG;
    }
    return build_block_device(d, a, b, 0, 0, K + row0 * ldk + col0, nullptr, si, sk, sj, sl, 0, 0);
}

// sum over the block of dK/dl[r,c] * (alpha[r] alpha[c] - W[min(r,c), max(r,c)])  ->  *result (device double)
extern "C" int cspbo_block_trace(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b, const double *alpha,
                                 const double *W, long long ldw, long long row0, long long col0, double *result, int accumulate)
{
    if (!desc || !a || !b || !alpha || !W || !result) { set_error("block_trace: null argument"); return CSPBO_E_ARG; }
    cspbo_block_desc d = *desc;
    if (d.family != CSPBO_RBF || !(d.l > 0.0)) { set_error("block_trace: RBF blocks only (Dot gradients are closed-form traces)"); return CSPBO_E_ARG; }
    if (d.block < CSPBO_KEE || d.block > CSPBO_KFF || d.stress) { set_error("block_trace: bad block"); return CSPBO_E_ARG; }
    if (d.symmetric && (a != b || d.block == CSPBO_KEF || row0 != col0)) { set_error("block_trace: symmetric needs a == b on the diagonal"); return CSPBO_E_ARG; }
    const int nc1 = d.block == CSPBO_KFF ? 3 : 0, nc2 = d.block == CSPBO_KEE ? 0 : 3;
    if ((nc1 ? a->nc < 3 : a->nc != 0) || (nc2 ? b->nc < 3 : b->nc != 0)) { set_error("block_trace: pack kinds do not match the block"); return CSPBO_E_ARG; }
    const long long rows = (long long)a->m * std::max(nc1, 1), cols = (long long)b->m * std::max(nc2, 1);
    if (row0 < 0 || col0 < 0 || col0 + cols > ldw || row0 + rows > ldw) { set_error("block_trace: block does not fit (ld %lld)", ldw); return CSPBO_E_ARG; }
    if (a->m == 0 || b->m == 0) {
        if (!accumulate) CSPBO_CUDA(cudaMemsetAsync(result, 0, sizeof(double)));
        return CSPBO_OK;
    }
    d.with_grad = 1; d.use_tol = 0;
    // one partial per warp of the largest possible grid (split mapping: one A block per CTA)
    const long long nparts = (long long)a->n_blocks * b->n_blocks * 4;
    void *sp = nullptr;
    int rc = scratch_get(SCRATCH_GRAD, (size_t)nparts * sizeof(double), &sp);
    if (rc) return rc;
    CSPBO_CUDA(cudaMemsetAsync(sp, 0, (size_t)nparts * sizeof(double)));
    const TraceArgs tr = {alpha, W, static_cast<double *>(sp), ldw, row0, col0, 0, 0};
    rc = build_block_device(d, a, b, 0, 0, nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0, 1, nullptr, 0, -1, 0, 1, 0, 0, nullptr, nullptr, &tr);
    if (rc) return rc;
    trace_sum_kernel<<<1, 1024>>>(static_cast<const double *>(sp), nparts, result, accumulate);
    g_launches++;
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}

// The same for a SLAB of rows in packed (panel-layout) order: the A blocks [a_blk0, a_blk1) of pack a against all of
// pack b, matrix row of packed position p = row0 + p * nc1 (column: col0 + q * nc2), Wslab = rows [slab_row0, ...) of
// K^-1 in full (leading dimension ldw): what one rank contributes to the sharded hyper-gradient trace.
extern "C" int cspbo_block_trace_slab(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b, const double *alpha,
                                      const double *Wslab, long long ldw, long long row0, long long col0, long long slab_row0,
                                      int a_blk0, int a_blk1, double *result, int accumulate)
{
    if (!desc || !a || !b || !alpha || !Wslab || !result) { set_error("block_trace_slab: null argument"); return CSPBO_E_ARG; }
    cspbo_block_desc d = *desc;
    if (d.family != CSPBO_RBF || !(d.l > 0.0)) { set_error("block_trace_slab: RBF blocks only"); return CSPBO_E_ARG; }
    if (d.block < CSPBO_KEE || d.block > CSPBO_KFF || d.stress) { set_error("block_trace_slab: bad block"); return CSPBO_E_ARG; }
    const int nc1 = d.block == CSPBO_KFF ? 3 : 0, nc2 = d.block == CSPBO_KEE ? 0 : 3;
    if ((nc1 ? a->nc < 3 : a->nc != 0) || (nc2 ? b->nc < 3 : b->nc != 0)) { set_error("block_trace_slab: pack kinds do not match the block"); return CSPBO_E_ARG; }
    if (a->row_classes != 1 || b->row_classes != 1) { set_error("block_trace_slab: packs with row_classes = 1 only"); return CSPBO_E_ARG; }
    a_blk0 = std::max(a_blk0, 0); a_blk1 = std::min(a_blk1, (int)a->n_blocks);
    if (row0 < 0 || col0 < 0 || slab_row0 > row0 + (long long)a_blk0 * 8 * std::max(nc1, 1)) { set_error("block_trace_slab: bad offsets"); return CSPBO_E_ARG; }
    if (a_blk1 <= a_blk0 || b->m == 0) {
        if (!accumulate) CSPBO_CUDA(cudaMemsetAsync(result, 0, sizeof(double)));
        return CSPBO_OK;
    }
    // a == b (K_EE, K_FF): dK/dl and K^-1 are symmetric, so the blocks b >= a of the slab's rows, off-diagonal tiles counted
    // twice, give the same total once every rank has added its rows -- half the tiles
    d.with_grad = 1; d.use_tol = 0; d.symmetric = (a == b && d.block != CSPBO_KEF) ? 1 : 0;
    const long long nparts = (long long)(a_blk1 - a_blk0) * b->n_blocks * 4;
    void *sp = nullptr;
    int rc = scratch_get(SCRATCH_GRAD, (size_t)nparts * sizeof(double), &sp);
    if (rc) return rc;
    CSPBO_CUDA(cudaMemsetAsync(sp, 0, (size_t)nparts * sizeof(double)));
    const TraceArgs tr = {alpha, Wslab, static_cast<double *>(sp), ldw, row0, col0, 1, slab_row0};
    rc = build_block_device(d, a, b, 0, 0, nullptr, nullptr, 0, 0, 0, 0, 0, 0, 0, 1, nullptr, a_blk0, a_blk1, 0, 1, 0, 0, nullptr, nullptr, &tr);
    if (rc) return rc;
    trace_sum_kernel<<<1, 1024>>>(static_cast<const double *>(sp), nparts, result, accumulate);
    g_launches++;
    CSPBO_CUDA(cudaGetLastError());
    return CSPBO_OK;
}

extern "C" int cspbo_block_diag(const cspbo_block_desc *desc, const cspbo_pack *a, double *out, int mem)
{
    if (!desc || !a || !out) { set_error("block_diag: null argument"); return CSPBO_E_ARG; }
    cspbo_block_desc d = *desc;
    if (d.block != CSPBO_KFF && d.block != CSPBO_KEE) { set_error("block_diag: K_EE or K_FF only"); return CSPBO_E_ARG; }
    if (d.family == CSPBO_RBF && !(d.l > 0.0)) { set_error("block_diag: RBF needs l > 0"); return CSPBO_E_ARG; }
    if ((d.block == CSPBO_KFF) != (a->nc >= 3)) { set_error("block_diag: pack kind does not match the block"); return CSPBO_E_ARG; }
    d.symmetric = 0; d.stress = 0; d.with_grad = 0;
    const long long n_out = (long long)a->m * (d.block == CSPBO_KFF ? 9 : 1);
    if (n_out == 0) return CSPBO_OK;
    double *dout = out;
    DevBuf tmp;
    if (mem == CSPBO_MEM_HOST) {
        CSPBO_CUDA(cudaMalloc(&tmp.p, (size_t)n_out * sizeof(double)));
        dout = tmp.p;
    }
    CSPBO_CUDA(cudaMemsetAsync(dout, 0, (size_t)n_out * sizeof(double)));
    int rc = build_block_device(d, a, a, 0, 0, dout, nullptr, 0, 0, 0, 0, 0, 0, 0, 1, nullptr, 0, -1, 1);
    if (rc) return rc;
    if (mem == CSPBO_MEM_HOST) {
        CSPBO_CUDA(cudaMemcpyAsync(out, dout, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost));
        CSPBO_CUDA(cudaStreamSynchronize(0));
    }
    return CSPBO_OK;
}

// ------------------------------------------------------------------------------------------
// sharded symmetric build + the IPC plumbing it needs (one process per GPU)
// ------------------------------------------------------------------------------------------
extern "C" int cspbo_dev_alloc(long long bytes, void **ptr)
{
    if (!ptr || bytes < 0) { set_error("dev_alloc: bad argument"); return CSPBO_E_ARG; }
    *ptr = nullptr;
    CSPBO_CUDA(cudaMalloc(ptr, (size_t)std::max<long long>(bytes, 8)));
    return CSPBO_OK;
}

extern "C" int cspbo_dev_free(void *ptr)
{
    CSPBO_CUDA(cudaFree(ptr));
    return CSPBO_OK;
}

extern "C" int cspbo_ipc_export(void *ptr, unsigned char *handle64)
{
    if (!ptr || !handle64) { set_error("ipc_export: null"); return CSPBO_E_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    CSPBO_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    return CSPBO_OK;
}

extern "C" int cspbo_ipc_open(const unsigned char *handle64, void **ptr)
{
    if (!ptr || !handle64) { set_error("ipc_open: null"); return CSPBO_E_ARG; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    *ptr = nullptr;
    CSPBO_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return CSPBO_OK;
}

extern "C" int cspbo_ipc_close(void *ptr)
{
    CSPBO_CUDA(cudaIpcCloseMemHandle(ptr));
    return CSPBO_OK;
}

// dst[height rows of width bytes, pitch dpitch] (pinned host) <- src (device, pitch spitch), asynchronous on `stream`
extern "C" int cspbo_copy_rows_to_host(void *dst, long long dpitch, const void *src, long long spitch, long long width, long long height,
                                       void *stream)
{
    if (!dst || !src || width < 0 || height < 0 || dpitch < width || spitch < width) { set_error("copy_rows_to_host: bad argument"); return CSPBO_E_ARG; }
    if (width == 0 || height == 0) return CSPBO_OK;
    CSPBO_CUDA(cudaMemcpy2DAsync(dst, (size_t)dpitch, src, (size_t)spitch, (size_t)width, (size_t)height, cudaMemcpyDeviceToHost,
                                 static_cast<cudaStream_t>(stream)));
    return CSPBO_OK;
}

extern "C" long long cspbo_sharded_rows(const cspbo_pack *a, int rank, int nranks)
{
    return cspbo_sharded_rows_ex(a, rank, nranks, 1, 0);
}

extern "C" long long cspbo_sharded_rows_ex(const cspbo_pack *a, int rank, int nranks, int sb_blocks, int zigzag)
{
    if (!a || nranks < 1 || rank < 0 || rank >= nranks || sb_blocks < 1) return -1;
    return (long long)sharded_superblocks((int)a->n_blocks, rank, nranks, sb_blocks, zigzag) * sb_blocks * 8;
}

extern "C" int cspbo_pack_order(const cspbo_pack *a, int *order, long long n)
{
    if (!a || !order) { set_error("pack_order: null"); return CSPBO_E_ARG; }
    if (n < (long long)a->n_blocks * 8) { set_error("pack_order: buffer too small"); return CSPBO_E_ARG; }
    for (long long i = 0; i < (long long)a->n_blocks * 8; i++) order[i] = a->h_slot_group[(size_t)i];
    return CSPBO_OK;
}

extern "C" int cspbo_block_build_sharded(const cspbo_block_desc *desc, const cspbo_pack *a, int rank, int nranks,
                                         double *const *outs)
{
    return cspbo_block_build_sharded_ex(desc, a, rank, nranks, outs, 1, 0, 0);
}

extern "C" int cspbo_block_build_sharded_ex(const cspbo_block_desc *desc, const cspbo_pack *a, int rank, int nranks,
                                            double *const *outs, int sb_blocks, int zigzag, int packed_cols)
{
    return cspbo_block_build_sharded_ld(desc, a, rank, nranks, outs, sb_blocks, zigzag, packed_cols, 0);
}

extern "C" int cspbo_block_build_sharded_ld(const cspbo_block_desc *desc, const cspbo_pack *a, int rank, int nranks,
                                            double *const *outs, int sb_blocks, int zigzag, int packed_cols, long long ld)
{
    return cspbo_block_build_sharded_slab(desc, a, rank, nranks, outs, sb_blocks, zigzag, packed_cols, ld, 0, -1, nullptr, nullptr, nullptr);
}

// work[a] = tile pairs of the symmetric build that belong to row block a: sum over b >= a and species of T(a,e) T(b,e)
extern "C" int cspbo_pack_block_work(const cspbo_pack *a, double *work, long long n)
{
    if (!a || !work || n < a->n_blocks) { set_error("pack_block_work: bad argument"); return CSPBO_E_ARG; }
    const int nb = (int)a->n_blocks, ns = (int)a->species.size();
    std::vector<double> suffix((size_t)std::max(ns, 1), 0.0);      // sum over b >= a of T(b, e)
    for (int blk = nb - 1; blk >= 0; blk--) {
        double w = 0.0;
        for (int e = 0; e < ns; e++) {
            const double t = a->h_tile_start[(size_t)blk * ns + e + 1] - a->h_tile_start[(size_t)blk * ns + e];
            suffix[(size_t)e] += t;
            w += t * suffix[(size_t)e];
        }
        work[blk] = w;
    }
    return CSPBO_OK;
}

extern "C" int cspbo_block_build_sharded_slab(const cspbo_block_desc *desc, const cspbo_pack *a, int rank, int nranks,
                                              double *const *outs, int sb_blocks, int zigzag, int packed_cols, long long ld,
                                              int slot0, int slot1, void *stream, const int *sb_perm, const int *sb_inv)
{
    if ((sb_perm == nullptr) != (sb_inv == nullptr)) { set_error("block_build_sharded: sb_perm and sb_inv go together"); return CSPBO_E_ARG; }
    if (!desc || !a || !outs) { set_error("block_build_sharded: null argument"); return CSPBO_E_ARG; }
    if (sb_blocks < 1) { set_error("block_build_sharded: sb_blocks must be >= 1"); return CSPBO_E_ARG; }
    if (packed_cols && a->row_classes != 1) { set_error("block_build_sharded: packed_cols needs a pack with row_classes = 1"); return CSPBO_E_ARG; }
    if (nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) { set_error("block_build_sharded: bad rank %d/%d", rank, nranks); return CSPBO_E_ARG; }
    for (int r = 0; r < nranks; r++)
        if (!outs[r]) { set_error("block_build_sharded: outs[%d] is null", r); return CSPBO_E_ARG; }
    cspbo_block_desc d = *desc;
    if (d.block != CSPBO_KFF && d.block != CSPBO_KEE) { set_error("block_build_sharded: K_FF or K_EE only"); return CSPBO_E_ARG; }
    if (d.stress || d.with_grad) { set_error("block_build_sharded: no stress / grad variants"); return CSPBO_E_ARG; }
    if (d.family == CSPBO_RBF && !(d.l > 0.0)) { set_error("block_build_sharded: RBF needs l > 0"); return CSPBO_E_ARG; }
    d.symmetric = 1;
    const long long m = a->m;
    const long long cols = (d.block == CSPBO_KEE) ? m : 3 * m;
    if (ld == 0) ld = cols;
    if (ld < cols) { set_error("block_build_sharded: row pitch %lld is below the %lld columns of the block", ld, cols); return CSPBO_E_ARG; }
    long long si, sk, sj, sl;
    if (d.block == CSPBO_KEE) {
        if (a->nc != 0) { set_error("block_build_sharded: K_EE needs an energy pack"); return CSPBO_E_ARG; }
        si = ld; sk = 0; sj = 1; sl = 0;
    } else {
        if (a->nc < 3) { set_error("block_build_sharded: K_FF needs a force pack"); return CSPBO_E_ARG; }
        si = 3 * ld; sk = ld; sj = 3; sl = 1;
    }
    return build_block_device(d, a, a, 0, 0, outs[rank], nullptr, si, sk, sj, sl, 0, static_cast<cudaStream_t>(stream), rank, nranks, outs,
                              slot0, slot1, 0, sb_blocks, zigzag, packed_cols, nullptr, nullptr, nullptr, sb_perm, sb_inv);
}

extern "C" int cspbo_block_build_panel(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b, int rank, int nranks,
                                       double *const *outs, long long ld, int nbw, long long row0_a, long long row0_b)
{
    if (!desc || !a || !b || !outs) { set_error("block_build_panel: null argument"); return CSPBO_E_ARG; }
    if (nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) { set_error("block_build_panel: bad rank %d/%d", rank, nranks); return CSPBO_E_ARG; }
    for (int r = 0; r < nranks; r++)
        if (!outs[r]) { set_error("block_build_panel: outs[%d] is null", r); return CSPBO_E_ARG; }
    cspbo_block_desc d = *desc;
    if (d.stress || d.with_grad) { set_error("block_build_panel: no stress / grad variants"); return CSPBO_E_ARG; }
    if (d.family == CSPBO_RBF && !(d.l > 0.0)) { set_error("block_build_panel: RBF needs l > 0"); return CSPBO_E_ARG; }
    if (nbw <= 0 || nbw % 24 != 0) { set_error("block_build_panel: nbw must be a positive multiple of 24"); return CSPBO_E_ARG; }
    if (row0_a < 0 || row0_b < 0 || row0_a % nbw || row0_b % nbw) { set_error("block_build_panel: packs must start at panel boundaries"); return CSPBO_E_ARG; }
    if (a->row_classes != 1 || b->row_classes != 1) { set_error("block_build_panel: packs with row_classes = 1 only"); return CSPBO_E_ARG; }
    long long rows_a, rows_b;
    if (d.block == CSPBO_KEE) {
        if (a != b || a->nc != 0) { set_error("block_build_panel: K_EE needs one energy pack on both sides"); return CSPBO_E_ARG; }
        rows_a = rows_b = a->m; d.symmetric = 1;
    } else if (d.block == CSPBO_KFF) {
        if (a != b || a->nc < 3) { set_error("block_build_panel: K_FF needs one force pack on both sides"); return CSPBO_E_ARG; }
        rows_a = rows_b = 3LL * a->m; d.symmetric = 1;
    } else if (d.block == CSPBO_KEF) {
        if (a->nc != 0 || b->nc < 3) { set_error("block_build_panel: K_EF needs (energy, force) packs"); return CSPBO_E_ARG; }
        rows_a = a->m; rows_b = 3LL * b->m; d.symmetric = 0;
    } else { set_error("block_build_panel: bad block"); return CSPBO_E_ARG; }
    if (row0_a + rows_a > ld || row0_b + rows_b > ld) { set_error("block_build_panel: block does not fit (ld %lld)", ld); return CSPBO_E_ARG; }
    if (row0_a == row0_b && a != b) { set_error("block_build_panel: two packs at the same rows"); return CSPBO_E_ARG; }
    PanelArgs pa = {ld, nbw, row0_a, row0_b};
    return build_block_device(d, a, b, 0, 0, outs[rank], nullptr, 0, 0, 0, 0, 0, 0, rank, nranks, outs, 0, -1, 0, 1, 1, 1, &pa);
}

// ------------------------------------------------------------------------------------------
// reference-compatible entry points
// ------------------------------------------------------------------------------------------
namespace {

struct PackGuard {
    cspbo_pack *p = nullptr;
    ~PackGuard() { cspbo_pack_destroy(p); }
};

int groups_of(const int *inds, int n)
{
    int m = 0;
    for (int i = 0; i < n; i++) m = std::max(m, inds[i] + 1);
    return m;
}

// `interleaved_grad`: RBF kef_many_with_grad writes K and dK/dl into one [m1,m2,6] array.
int compat_block(int family, int block, int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                 double sigma2, double sigma02, double l, double tol, int use_tol, int with_grad, int stress,
                 const double *x1, const double *dx1dr, const int *ele1, const int *x1_inds,
                 const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                 double *pout, double *dpout_dl)
{
    if (n1 < 0 || n2 < 0 || x2i < 0 || !pout) { set_error("bad sizes or null pout"); return CSPBO_E_ARG; }
    if (n1 == 0 || n2 == 0 || x2i == 0) return CSPBO_OK;
    const int m1 = groups_of(x1_inds, n1);
    PackGuard A, B;
    int rc;
    const int nc1 = (block == CSPBO_KFF) ? (stress ? 9 : 3) : 0;
    const int nc2 = (block == CSPBO_KEE) ? 0 : ((block == CSPBO_KEF && stress) ? 9 : 3);
    rc = cspbo_pack_create(n1, d, nc1, m1, 0, n1, x1, nc1 ? dx1dr : nullptr, ele1, x1_inds, CSPBO_MEM_HOST, &A.p);
    if (rc) return rc;
    rc = cspbo_pack_create(n2, d, nc2, x2i, n2_start, n2_end, x2, nc2 ? dx2dr : nullptr, ele2, x2_inds,
                           CSPBO_MEM_HOST, &B.p);
    if (rc) return rc;
    cspbo_block_desc desc;
    memset(&desc, 0, sizeof(desc));
    desc.family = family; desc.block = block; desc.zeta = zeta;
    desc.sigma2 = sigma2; desc.sigma02 = sigma02; desc.l = l; desc.tol = tol; desc.use_tol = use_tol;
    desc.with_grad = with_grad; desc.stress = stress;
    // Dot K_EE: the reference multiplies by sigma2 inside C (dot_kernel.cpp:47); RBF carries sigma2 in K
    desc.scale = (family == CSPBO_DOT && block == CSPBO_KEE) ? sigma2 : 1.0;
    desc.scale_grad = 1.0;
    desc.symmetric = 0;
    if (with_grad && block == CSPBO_KEF) {
        // interleaved [m1,m2,6] output: run K and dK/dl through the strided device path by hand
        const long long m2 = x2i, n_out = (long long)m1 * m2 * 6;
        DevBuf tmp;
        CSPBO_CUDA(cudaMalloc(&tmp.p, (size_t)n_out * sizeof(double)));
        CSPBO_CUDA(cudaMemcpyAsync(tmp.p, pout, (size_t)n_out * sizeof(double), cudaMemcpyHostToDevice));
        rc = build_block_device(desc, A.p, B.p, 0, 0, tmp.p, tmp.p + 3, m2 * 6, 0, 6, 1, 1, 0);
        if (rc) return rc;
        CSPBO_CUDA(cudaMemcpyAsync(pout, tmp.p, (size_t)n_out * sizeof(double), cudaMemcpyDeviceToHost));
        CSPBO_CUDA(cudaStreamSynchronize(0));
        return CSPBO_OK;
    }
    return cspbo_block_build(&desc, A.p, B.p, pout, dpout_dl, CSPBO_MEM_HOST, /*accumulate=*/1);
}

}  // namespace

#define XE const double *x1, const int *ele1, const int *x1_inds
#define XF2 const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds
#define XF1 const double *x1, const double *dx1dr, const int *ele1, const int *x1_inds

extern "C" int cspbo_dot_kee_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double sigma02, XE,
                                  const double *x2, const int *ele2, const int *x2_inds, double *pout)
{
    return compat_block(CSPBO_DOT, CSPBO_KEE, n1, n2, 0, n2, d, x2i, zeta, sigma2, sigma02, 0, 0, 0, 0, 0,
                        x1, nullptr, ele1, x1_inds, x2, nullptr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_dot_kef_many(int n1, int n2, int d, int x2i, double zeta, XE, XF2, double *pout)
{
    return compat_block(CSPBO_DOT, CSPBO_KEF, n1, n2, 0, n2, d, x2i, zeta, 1, 0, 0, 0, 0, 0, 0,
                        x1, nullptr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_dot_kef_many_stress(int n1, int n2, int d, int x2i, double zeta, XE, XF2, double *pout)
{
    return compat_block(CSPBO_DOT, CSPBO_KEF, n1, n2, 0, n2, d, x2i, zeta, 1, 0, 0, 0, 0, 0, 1,
                        x1, nullptr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_dot_kff_many(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta, XF1, XF2,
                                  double *pout)
{
    return compat_block(CSPBO_DOT, CSPBO_KFF, n1, n2, n2_start, n2_end, d, x2i, zeta, 1, 0, 0, 0, 0, 0, 0,
                        x1, dx1dr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_dot_kff_many_stress(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta, XF1,
                                         XF2, double *pout)
{
    return compat_block(CSPBO_DOT, CSPBO_KFF, n1, n2, n2_start, n2_end, d, x2i, zeta, 1, 0, 0, 0, 0, 0, 1,
                        x1, dx1dr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}

extern "C" int cspbo_rbf_kee_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2, XE,
                                  const double *x2, const int *ele2, const int *x2_inds, double *pout)
{
    return compat_block(CSPBO_RBF, CSPBO_KEE, n1, n2, 0, n2, d, x2i, zeta, sigma2, 0, sqrt(l2), 0, 0, 0, 0,
                        x1, nullptr, ele1, x1_inds, x2, nullptr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_rbf_kee_many_with_grad(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2, XE,
                                            const double *x2, const int *ele2, const int *x2_inds, double *pout,
                                            double *dpout_dl)
{
    if (!dpout_dl) { set_error("kee_many_with_grad: null dpout_dl"); return CSPBO_E_ARG; }
    return compat_block(CSPBO_RBF, CSPBO_KEE, n1, n2, 0, n2, d, x2i, zeta, sigma2, 0, sqrt(l2), 0, 0, 1, 0,
                        x1, nullptr, ele1, x1_inds, x2, nullptr, ele2, x2_inds, pout, dpout_dl);
}
extern "C" int cspbo_rbf_kef_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2, XE, XF2,
                                  double *pout)
{
    return compat_block(CSPBO_RBF, CSPBO_KEF, n1, n2, 0, n2, d, x2i, zeta, sigma2, 0, sqrt(l2), 0, 0, 0, 0,
                        x1, nullptr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_rbf_kef_many_with_grad(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l, XE,
                                            XF2, double *pout)
{
    return compat_block(CSPBO_RBF, CSPBO_KEF, n1, n2, 0, n2, d, x2i, zeta, sigma2, 0, l, 0, 0, 1, 0,
                        x1, nullptr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_rbf_kef_many_stress(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2, XE,
                                         XF2, double *pout)
{
    return compat_block(CSPBO_RBF, CSPBO_KEF, n1, n2, 0, n2, d, x2i, zeta, sigma2, 0, sqrt(l2), 0, 0, 0, 1,
                        x1, nullptr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_rbf_kff_many(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                                  double sigma2, double l2, double tol, XF1, XF2, double *pout)
{
    return compat_block(CSPBO_RBF, CSPBO_KFF, n1, n2, n2_start, n2_end, d, x2i, zeta, sigma2, 0, sqrt(l2), tol, 1, 0,
                        0, x1, dx1dr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}
extern "C" int cspbo_rbf_kff_many_with_grad(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                                            double sigma2, double l, XF1, XF2, double *pout, double *dpout_dl)
{
    if (!dpout_dl) { set_error("kff_many_with_grad: null dpout_dl"); return CSPBO_E_ARG; }
    return compat_block(CSPBO_RBF, CSPBO_KFF, n1, n2, n2_start, n2_end, d, x2i, zeta, sigma2, 0, l, 0, 0, 1, 0,
                        x1, dx1dr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, dpout_dl);
}
extern "C" int cspbo_rbf_kff_many_stress(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                                         double sigma2, double l2, double tol, XF1, XF2, double *pout)
{
    return compat_block(CSPBO_RBF, CSPBO_KFF, n1, n2, n2_start, n2_end, d, x2i, zeta, sigma2, 0, sqrt(l2), tol, 1, 0,
                        1, x1, dx1dr, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout, nullptr);
}
