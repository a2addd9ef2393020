// common.h -- shared declarations of libcspbo_b200 (host side)
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <vector>

#include "cspbo_b200.h"

namespace cspbo {

void set_error(const char *fmt, ...);
extern std::atomic<long long> g_launches;

#define CSPBO_CUDA(call)                                                                     \
    do {                                                                                     \
        cudaError_t e__ = (call);                                                            \
        if (e__ != cudaSuccess) {                                                            \
            cspbo::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call,                    \
                             cudaGetErrorString(e__));                                       \
            return (e__ == cudaErrorMemoryAllocation) ? CSPBO_E_NOMEM                        \
                   : (e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver)        \
                       ? CSPBO_E_NODEVICE                                                    \
                       : CSPBO_E_CUDA;                                                       \
        }                                                                                    \
    } while (0)

// Tile geometry.  Groups (structures / force atoms) are sorted by their per-species row counts and
// cut into blocks of 8; a block owns, for every species e, T(block,e) = max_count tiles.  A tile holds
// 8 row slots; slot r of every tile of a block belongs to group r of the block.  Each slot carries NV vectors (the unit descriptor u and,
// for force packs, the nc projected+scaled derivative vectors) of 4*KS doubles, stored in
// mma.m8n8k4 fragment order so that one 128-bit load per lane fetches a lane's operands for
// two consecutive k-steps:
//     index(v, f, slot) = ((v*(KS/2) + f/8)*32 + slot*4 + f%4)*2 + (f/4)%2
constexpr int kSlots = 8;

// Grow-only device scratch buffers, cached per (device, slot) so that host-pointer calls do not
// cudaMalloc / cudaFree gigabytes on every call (measured: 2 ms typical, but outliers of 0.2-2 s).
// Not thread safe across concurrent builds in one process; cspbo_release_scratch() frees them.
enum { SCRATCH_OUT = 0, SCRATCH_GRAD = 1, SCRATCH_X = 2, SCRATCH_DX = 3, SCRATCH_DEST = 4, SCRATCH_GEMV = 5, SCRATCH_PLANES = 6, SCRATCH_SLOTS = 7 };
int scratch_get(int slot, size_t bytes, void **ptr);

// Small caching allocator for the pack's device arrays: blocks released by cspbo_pack_destroy are kept
// (up to 2 GiB per device) and handed out again, so re-packing the same training set every call
// does not pay cudaMalloc / cudaFree (which synchronise and occasionally take >100 ms).
int dev_alloc_cached(size_t bytes, void **ptr);
void dev_free_cached(void *ptr);

// k-steps (of 4 features) per vector of a tile: 6 for d <= 24 and 14 for d <= 56 (A fragments live in registers),
// a multiple of 6 beyond that (chunked kernels, tiles.cu) -- any descriptor length is accepted, like the reference's
// `for i < d` loops (dot_kernel.cpp:257-289).
inline int ksteps_for(int d)
{
    if (d <= 0) return -1;
    if (d <= 24) return 6;
    if (d <= 56) return 14;
    if (d > (1 << 20)) return -1;
    return 6 * ((d + 23) / 24);
}

}  // namespace cspbo

// The opaque pack object of the C ABI.
struct cspbo_pack {
    int n = 0, d = 0, nc = 0, m = 0, nv = 0, ks = 0;
    int device = 0;
    long long n_tiles = 0, n_blocks = 0;
    long long tile_doubles = 0;
    // host-side plan
    std::vector<int> species;      // species ids present in the packed rows, most populous first
    std::vector<int> h_tile_start; // first tile of (block, species): index block*ns + e (size n_blocks*ns+1)
    std::vector<int> h_slot_group; // [n_blocks*8] group id of each slot, -1 = empty
    std::vector<long long> rows_per_species;
    int row_classes = 1;            // groups were split by g % row_classes before blocking
    std::vector<int> h_class_blk0;  // first block of each row class (size row_classes+1)
    // device-side tables
    double *d_tiles = nullptr;
    int *d_tile_start = nullptr;
    int *d_slot_group = nullptr;     // device copy of h_slot_group
    unsigned *d_tile_valid = nullptr;  // [n_tiles] bit r set <=> slot r holds a row with |x|>1e-8
    long long bytes = 0;
};
