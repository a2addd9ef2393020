// tiles.cu -- the covariance-block tile kernels (K_EE / K_EF / K_FF, Dot and RBF) for sm_100a.
//
// One templated kernel family replaces the 13 scalar loops of the reference
// (cspbo/kernels/dot_kernel.cpp:4-506, rbf_kernel.cpp:4-822).  A warp owns one pair of packed 8-group
// blocks (pack.cu) and walks their species buckets; every 8x8 tile pair costs NV1*NV2*KS FP64 tensor-core
// instructions (DMMA.8x8x4, the only FP64 MMA shape sm_100a has natively) that produce, per
// thread and for two (row a, row b) pairs, the full set of dot products
//     H[0][0] = c = u_a.u_b      H[k][0] = P_k = A^_a[:,k].u_b
//     H[0][l] = Q_l = u_a.A^_b[:,l]   H[k][l] = T_kl = A^_a[:,k].A^_b[:,l]
// followed by the zeta / exp chain rule in registers:
//     Dot K_FF  (dot_kernel.cpp:265-323):   F_kl = c^(z-1) T_kl + (z-1) c^(z-2) P_k Q_l
//     RBF K_FF  (rbf_kernel.cpp:390-462):   F_kl = dK/dD [ z c^(z-1) T_kl
//                                               + (z(z-1)c^(z-2) + z^2 c^(2z-2)/(2l^2)) P_k Q_l ]
//     K_EF      (dot_kernel.cpp:102-121, rbf_kernel.cpp:141-164):  z c^(z-1) Q_l  (x -dK/dD for RBF)
//     K_EE      (dot_kernel.cpp:45-47, rbf_kernel.cpp:41-43)
// Because slot r of every tile of a block is the same group, the reference's scatter-add over
// (row a in group i, row b in group j) is a plain register accumulation; each thread owns the
// 3x3 (or 1x3, 1x1) output blocks of two fixed group pairs and writes them once.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.h"

namespace cspbo {

constexpr int kMaxSpecies = 32;   // species buckets mapped through kernel parameters; larger sets go through a device array
constexpr int kMaxRanks = 8;
// 3 CTAs x 4 warps per SM = 3 warps per scheduler: while one warp runs its chain-rule epilogue the
// other two keep the FP64 tensor pipe fed (<= 170 registers/thread, <= 75 KB shared memory per CTA)
#ifndef CSPBO_WARPS
#define CSPBO_WARPS 4
#define CSPBO_CTAS 3
#define CSPBO_STAGE_KB 36
#endif
constexpr int kMaxWarps = CSPBO_WARPS;
constexpr int kMinCtas = CSPBO_CTAS;
constexpr size_t kSmemChunk = CSPBO_STAGE_KB * 1024;   // per stage; two stages are double-buffered
// Rasterisation: A-block groups per strip.  Measured on B200 (20 001 force rows): strips of 1..64 are 3-6 %
// slower than no strips (B block shared by all consecutive CTAs stays L2-hot, A tiles stream), so
// the default is a single strip.
constexpr int kStrip = 1 << 30;

struct SideDev {
    const double *tiles;
    const int *tile_start;     // [nblk*ns + 1]: first tile of (block, species)
    const int *slot_group;
    const unsigned *tile_valid;
    int tile_doubles;  // NV*KS*32
    int voff;          // first derivative vector used is 1+voff
    int ns;            // species buckets per block
    int m;             // groups of the pack (packed positions >= m are padding slots)
};

struct PanelSide { int g, sb0; long long col0; };

struct TileParams {
    SideDev A, B;
    int a_nblk, b_nblk;
    int a_begin;  // first A block of this launch (row-slab pipelining of host-bound builds); a_nblk is the end
    int emap[kMaxSpecies];  // species bucket of A -> bucket of B holding the same species id, or -1
    const int *emap_dev;    // the same map in device memory when A has more than kMaxSpecies buckets (else null)
    int kc;                 // k-steps that hold features, ceil(d / 4) (chunked kernels: those of the last chunk); the rest of the KS multiply zeros (VK instantiations skip them)
    int nch, staged;        // chunked kernels (d > 56): k-chunks of 6 steps per tile; staged = B tiles go through shared memory
    int npass;              // stress variants: gridDim.y passes, pass q uses derivative vectors 1+3q .. 3+3q of the side
    int pass_side;          //   pass_side (0 = A: K_FF stress rows, 1 = B: K_EF stress columns) and stores at out + q*pass_off
    long long pass_off;
    int tbc;  // B tiles per shared-memory chunk
    int bsplit;      // > 1 (small problems): gridDim.y CTAs share one (A block, B block) pair, CTA y takes the B-tile runs
                     //      y, y + bsplit, ... and stores its partial block into plane y (out + y * pass_off); the planes are
                     //      summed in a fixed order by plane_reduce_kernel
    int strip;  // A-block groups per rasterisation strip
    int symmetric;
    int split;       // 1: the warps of a CTA share ONE A block and split its tiles (small problems: 4x more
                     //    parallelism, 4x shorter critical path); partial sums are reduced through shared memory
    int stream_out;  // 1: outputs are stored with the evict-first (streaming) policy
    int diag;        // 1: one warp per block, b = a, only the (i,i) group pairs are stored: out[i, NC1, NC2]
    int accumulate;  // 1: out += result (reference `pout +=` semantics), 0: out = result
    double zeta, sigma2, sigma02, inv2l2, inv_l, inv_l3, tol;
    int zint;        // zeta when it is an integer in [1, 8] (other than the Z2 instantiations' 2), else 0
    int use_tol;
    double scale, scale_grad;
    double *out, *out_grad;
    // trace mode (FAM_RBF_GRAD only): nothing is stored; every warp reduces sum over its output elements of
    //   scale_grad * dK/dl[r, c] * (tr_alpha[r] tr_alpha[c] - tr_w[min(r,c) * tr_ld + max(r,c)])
    // (r = tr_row0 + group_i * NC1 + k, c = tr_col0 + group_j * NC2 + l, caller order; off-diagonal tiles of a symmetric
    // build count twice) into tr_partial[cta * kMaxWarps + warp]: the hyper-gradient trace tr((aa^T - K^-1) dK/dl) of
    // gaussianprocess.py:496-499 without dK/dl ever existing in memory
    int trace;
    const double *tr_alpha, *tr_w;
    double *tr_partial;
    long long tr_ld, tr_row0, tr_col0;
    // tr_packed (sharded LML gradient): rows / columns are PACKED positions (r = tr_row0 + (block*8 + slot) * NC1 + k, the
    // panel-layout order) and tr_w is a slab of full rows of K^-1 starting at matrix row tr_slab0: w = a_r a_c - tr_w[(r - tr_slab0) * tr_ld + c]
    int tr_packed;
    long long tr_slab0;
    long long si, sk, sj, sl;  // out index = i*si + k*sk + j*sj + l*sl
    // sharded symmetric build (one process per GPU): rank r owns the A blocks a with a % nranks == r
    // and stores row block a at local block a / nranks of outs[r]; mirrored tiles are stored
    // straight into the owner's buffer over NVLink peer memory (outs[] are IPC-mapped pointers).
    // sb_blocks (G) consecutive blocks form a superblock, the unit of ownership: superblock s belongs to
    // rank s % nranks, or with zigzag to the mirrored rank on odd cycles (balances the triangular work
    // when G is large), and is the (s / nranks)-th superblock of its owner's buffer.  packed_cols: columns are
    // stored in packed order too (position b_blk*8+slot instead of the caller's group id), which makes the
    // distributed matrix the symmetric permutation P K P^T that the sharded Cholesky factorises in place.
    int sharded, nranks, rank;
    int sb_blocks, zigzag, packed_cols, a_slots;   // a_slots: blocks (incl. a ragged tail) this rank enumerates
    int a_slot0;   // sharded == 1: first block slot of this launch in the rank's buffer (row-slab pipelining)
    // sharded == 1, optional work-balanced ownership: the dealing formula runs on VIRTUAL superblock indices; sb_perm[v] is the
    // real superblock at virtual index v (-1: none), sb_inv its inverse.  Null = identity.
    const int *sb_perm, *sb_inv;
    double *outs[kMaxRanks];
    // sharded == 2, "panel layout": the blocks of SEVERAL packs (energies, forces) live in one distributed matrix of
    // leading dimension ldo whose rows are dealt to the ranks by superblocks of nbw ROWS (zigzag); a pack starts at a
    // superblock boundary (global superblock sb0, global row = column col0) and g of its 8-group blocks make a superblock.
    // Side A = rows of the tile, side B = columns; the mirrored tile goes to the owner of the B rows.
    int nbw, cyc0, pair;      // pair: A and B are different packs (K_EF): every tile is mirrored
    long long ldo;
    PanelSide pa, pb;
};

__device__ __host__ __forceinline__ int sb_owner(int sb, int nranks, int zigzag)
{
    const int cyc = sb / nranks, pos = sb - cyc * nranks;
    return (zigzag && (cyc & 1)) ? nranks - 1 - pos : pos;
}

enum { MODE_KEE = 0, MODE_KEF = 1, MODE_KFF = 2 };
enum { FAM_DOT = 0, FAM_RBF = 1, FAM_RBF_GRAD = 2 };

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}


// ---- 1-D bulk TMA (cp.async.bulk, SASS UBLKCP) + mbarrier helpers -----------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Wait for the phase with the given parity; a bounded spin turns a lost transaction into a trap
// instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    unsigned done = 0;
    for (unsigned spin = 0; !done; spin++) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done)
                     : "r"(smem_u32(bar)), "r"(parity)
                     : "memory");
        if (spin > (1u << 26)) __trap();
    }
}

// Walks the (species, chunk) runs of one B block: every run is a contiguous piece of the packed
// tile array, at most p.tbc tiles long, so one bulk copy fetches it.
struct ChunkIter {
    int eA, c0, b_t0, Tb;
    __device__ __forceinline__ void next_species(const TileParams &p, int b_blk)
    {
        for (++eA; eA < p.A.ns; ++eA) {
            const int eB = p.emap_dev ? p.emap_dev[eA] : p.emap[eA];
            if (eB < 0) continue;
            b_t0 = p.B.tile_start[b_blk * p.B.ns + eB];
            Tb = p.B.tile_start[b_blk * p.B.ns + eB + 1] - b_t0;
            c0 = 0;
            if (Tb > 0) return;
        }
    }
    __device__ __forceinline__ void start(const TileParams &p, int b_blk) { eA = -1; c0 = 0; b_t0 = 0; Tb = 0; next_species(p, b_blk); }
    __device__ __forceinline__ bool done(const TileParams &p) const { return eA >= p.A.ns; }
    __device__ __forceinline__ void next(const TileParams &p, int b_blk)
    {
        c0 += p.tbc;
        if (c0 >= Tb) next_species(p, b_blk);
    }
    __device__ __forceinline__ int nt(const TileParams &p) const { return min(p.tbc, Tb - c0); }
};

// Per-pair scalar weights of the chain rule.
struct Weights {
    double w1, w2;    // F = w1*T + w2*P*Q            (kff)   | kef: w1*Q | kee: w1
    double g1, g2;    // d/dl companion: w1g*T + w2g*P*Q      | kef: g1*Q | kee: g1
};

template <int MODE, int FAM, bool Z2>
__device__ __forceinline__ Weights pair_weights(double c, bool valid, const TileParams &p)
{
    Weights w;
    w.g1 = w.g2 = 0.0;
    const double z = p.zeta;
    // powers of c
    double cm2, cm1, cz;  // c^(z-2), c^(z-1), c^z
    if (Z2) {
        cm2 = 1.0; cm1 = c; cz = c * c;
    } else if (p.zint != 0) {
        // integer zeta (1, 3, 4, ...): the powers by multiplication (a uniform branch) instead of libm's pow, which costs
        // the zeta = 3 K_FF a quarter of its run time (tools/ks14_probe.py)
        if (p.zint == 1) {
            cm2 = 1.0 / c;
        } else {
            cm2 = 1.0;
            for (int i = 2; i < p.zint; i++) cm2 *= c;
        }
        cm1 = c * cm2;
        cz = c * cm1;
    } else {
        cm2 = pow(c, z - 2.0);   // reference: d2 = pow(dx, zeta-2); d1 = dx*d2  (dot_kernel.cpp:265-267)
        cm1 = c * cm2;
        cz = (MODE == MODE_KFF) ? c * cm1 : pow(c, z);
        if (MODE == MODE_KEF) cm1 = pow(c, z - 1.0);  // dot_kernel.cpp:102
    }
    if (FAM == FAM_DOT) {
        if (MODE == MODE_KFF) { w.w1 = cm1; w.w2 = (z - 1.0) * cm2; }
        else if (MODE == MODE_KEF) { w.w1 = z * cm1; w.w2 = 0.0; }
        else { w.w1 = cz + p.sigma02; w.w2 = 0.0; }  // sigma2 applied through p.scale
    } else {
        // (a 256-entry-table exp with a degree-4 polynomial, 9 instead of 17 FP64 instructions, was measured
        // SLOWER here: K_EE RBF 18.3 -> 21.1 ms; the dependent L1 lookup is not hidden at 3 warps per scheduler.  Round 2
        // repeated it with a 16-entry 2^(i/16) table in SHARED memory + degree-5 polynomial, 11 FP64 instructions, 1.4e-13
        // accurate: K_EE 16.8 -> 18.2 ms, K_EF 73.7 -> 74.0, K_FF 119.0 -> 118.3 (tools/rbf_probe.py): the epilogue is bound
        // by its dependent chain, not by FP64 issue slots, and libm's all-ALU polynomial pipelines better.  A third try made
        // the exp BRANCH-FREE (libm's inlined range check ends the basic block between the two exps of a thread): same
        // reduction and degree-11 polynomial with one clamp, 2e-16 accurate -- K_EE 16.9 -> 16.7 / 162.8 -> 168.0 ms,
        // K_EF 73.8 -> 75.5, K_FF 116.1 -> 116.5: no gain either, dropped)
        const double K = p.sigma2 * exp(-(1.0 - cz) * p.inv2l2);
        const double dKdD = K * p.inv2l2;
        const double gl = (cz - 1.0) * p.inv_l3 + 2.0 * p.inv_l;  // (D-1)/l^3 + 2/l
        if (MODE == MODE_KFF) {
            const double dD = z * cm1;                 // dD/dc
            const double ml = dD * dD * p.inv2l2;      // (1/(2l^2)) (dD/dc)^2
            w.w1 = dKdD * dD;
            w.w2 = dKdD * (z * (z - 1.0) * cm2 + ml);
            if (FAM == FAM_RBF_GRAD) {
                // dpout_dl -= F*gl + (2/l) dK/dD ml P Q      (rbf_kernel.cpp:621-629)
                w.g1 = -gl * w.w1;
                w.g2 = -(gl * w.w2 + 2.0 * p.inv_l * dKdD * ml);
            }
            if (p.use_tol && !(dKdD > p.tol)) valid = false;  // rbf_kernel.cpp:394
        } else if (MODE == MODE_KEF) {
            w.w1 = -dKdD * z * cm1;                    // pout -= C  (rbf_kernel.cpp:162-164)
            w.w2 = 0.0;
            if (FAM == FAM_RBF_GRAD) w.g1 = -w.w1 * gl;  // pout[3+l] += C*gl (rbf_kernel.cpp:244-246)
        } else {
            w.w1 = K; w.w2 = 0.0;
            if (FAM == FAM_RBF_GRAD) w.g1 = K * (1.0 - cz);  // rbf_kernel.cpp:92
        }
    }
    if (!valid) { w.w1 = w.w2 = w.g1 = w.g2 = 0.0; }
    return w;
}

// CHUNKED (descriptor lengths above 56): a tile holds p.nch chunks of KS k-steps per vector; the dot products of a
// tile pair accumulate over the chunks with the A fragments re-read from global memory (L1) per chunk instead of living
// in registers for the whole sweep, so any d runs at the register budget of the d <= 24 kernel (dot_kernel.cpp:257-289
// loops `i < d` for any d; SO3.py:205 gives d = 90 at nmax = lmax = 5).
// BS: the B-run split instantiations (p.bsplit > 1, small problems); a separate instantiation so that the code of the
// large builds stays exactly what it was (a run-time switch cost RBF K_EE 2-3 %).
// VK: the k-steps beyond p.kc = ceil(d / 4) are skipped -- d = 30 / 40 / 50 need 8 / 10 / 13 of the 14, d = 9 / 18 need 3 / 5 of the 6.
template <int MODE, int FAM, int KS, bool Z2, bool CHUNKED, bool BS = false, bool VK = false>
__global__ void __launch_bounds__(kMaxWarps * 32, (KS <= 6 && FAM != FAM_RBF_GRAD && !CHUNKED) ? kMinCtas : (kMinCtas > 1 ? 2 : 1))
block_tile_kernel(const TileParams p)
{
    constexpr int NV1 = (MODE == MODE_KFF) ? 4 : 1;
    constexpr int NV2 = (MODE == MODE_KEE) ? 1 : 4;
    constexpr int NC1 = (MODE == MODE_KFF) ? 3 : 1;
    constexpr int NC2 = (MODE == MODE_KEE) ? 1 : 3;
    constexpr bool GRAD = (FAM == FAM_RBF_GRAD);
#ifdef CSPBO_KEE_SINGLE_CHAIN
    constexpr bool DUAL = false;
#else
    constexpr bool DUAL = (MODE == MODE_KEE) && (BS || FAM != FAM_DOT);
#endif
    // (measured, tools/kee_probe.py: two chains are 15 % faster for small K_EE builds and 1-2 % for RBF at any size, but
    // 3-5 % slower for large Dot K_EE builds, whose 6 warps per scheduler hide the chain anyway; a run-time switch cost
    // the large builds 5-8 %, so the choice is made at compile time)
    constexpr int HALF = KS / 2;
    constexpr bool VARK = VK;
    const int nch = CHUNKED ? p.nch : 1;
    const int halft = HALF * nch;                        // 128-bit fragment rows per vector of a tile
    const bool staged = CHUNKED ? (p.staged != 0) : true;
    // stress variants: pass q (blockIdx.y) works on derivative vectors 1+3q .. 3+3q of one side
    const int pass = BS ? 0 : blockIdx.y;
    const int zb = BS ? blockIdx.y : 0, nzb = BS ? p.bsplit : 1;
    const int voffA = p.A.voff + ((p.npass > 1 && p.pass_side == 0) ? 3 * pass : 0);
    const int voffB = p.B.voff + ((p.npass > 1 && p.pass_side == 1) ? 3 * pass : 0);
    double *const out = p.out + blockIdx.y * p.pass_off;

    extern __shared__ __align__(16) double smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const int a_mine = p.sharded ? p.a_slots : (p.a_nblk - p.a_begin);
    const int a_units = p.split ? 1 : nwarps;     // A blocks per CTA
    const int npg = (a_mine + a_units - 1) / a_units;
    // rasterisation: CTAs walk a strip of p.strip A-block groups across all B blocks (one strip = all
    // A-block groups by default: consecutive CTAs share the B block, which stays L2-hot)
    const int strip_ids = p.strip * p.b_nblk;
    const int strip = blockIdx.x / strip_ids;
    const int rem = blockIdx.x - strip * strip_ids;
    const int strip_w = min(p.strip, npg - strip * p.strip);
    const int q = rem / strip_w;
    const int pg = strip * p.strip + (rem - q * strip_w);
    const int a_rel = pg * a_units + (p.split ? 0 : warp);
    const int a_idx = a_rel + (p.sharded == 1 ? p.a_slot0 : 0);
    int a_blk = p.a_begin + a_idx;
    if (p.sharded == 2) {   // panel layout: a_idx-th block slot of this rank within pack A's superblock range
        const int c = a_idx / p.pa.g, within = a_idx - c * p.pa.g;
        const int cyc = p.cyc0 + c;
        const int pos = (cyc & 1) ? p.nranks - 1 - p.rank : p.rank;
        const int sl = cyc * p.nranks + pos - p.pa.sb0;       // superblock index inside pack A (may fall outside)
        a_blk = sl < 0 ? p.a_nblk : sl * p.pa.g + within;
    } else if (p.sharded) {   // a_idx-th block of this rank's buffer -> global block id
        const int cyc = a_idx / p.sb_blocks, within = a_idx - cyc * p.sb_blocks;
        const int pos = (p.zigzag && (cyc & 1)) ? p.nranks - 1 - p.rank : p.rank;
        int sreal = cyc * p.nranks + pos;
        if (p.sb_perm) sreal = p.sb_perm[sreal];
        a_blk = sreal < 0 ? p.a_nblk : sreal * p.sb_blocks + within;
    }
    const int b_blk = p.diag ? a_blk : q;
    // symmetric build: only blocks with b >= a are computed (mirrored at write time)
    const bool active = a_rel < a_mine && a_blk < p.a_nblk && !(p.symmetric && b_blk < a_blk);

    // two stage buffers of p.tbc tiles, filled by 1-D bulk TMA copies that complete on an mbarrier
    const int stage_doubles = staged ? p.tbc * p.B.tile_doubles : 0;
    unsigned long long *full = reinterpret_cast<unsigned long long *>(smem + 2 * (size_t)stage_doubles);
    if (staged && threadIdx.x == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // CTAs none of whose warps has work (the lower block triangle of a symmetric build, ragged tails of a sharded one)
    // leave before streaming any B tile: for K_EE, where a tile pair is only KS DMMAs, an idle CTA that still walked
    // the TMA pipeline cost almost as much as a busy one
    if (!__syncthreads_or((int)active)) return;

    double acc[2][NC1][NC2];
    double accg[2][GRAD ? NC1 : 1][GRAD ? NC2 : 1];
#pragma unroll
    for (int e = 0; e < 2; e++)
#pragma unroll
        for (int k = 0; k < NC1; k++)
#pragma unroll
            for (int l = 0; l < NC2; l++) {
                acc[e][k][l] = 0.0;
                if (GRAD) accg[e][k][l] = 0.0;
            }

    const int ar = lane >> 2;        // a slot of this thread's accumulator rows
    const int bq = (lane & 3) * 2;   // first of its two b slots

    ChunkIter cur;
    cur.start(p, b_blk);
    if (BS)
        for (int s = 0; s < zb && !cur.done(p); s++) cur.next(p, b_blk);
    auto issue = [&](const ChunkIter &c, int st) {
        const unsigned bytes = (unsigned)(c.nt(p) * p.B.tile_doubles * (int)sizeof(double));
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic reads of this stage are done
        mbar_expect_tx(&full[st], bytes);
        tma_bulk_g2s(smem + (size_t)st * stage_doubles, p.B.tiles + (size_t)(c.b_t0 + c.c0) * p.B.tile_doubles, bytes, &full[st]);
    };
    if (staged && !cur.done(p) && threadIdx.x == 0) issue(cur, 0);
    for (int it = 0; !cur.done(p); it++) {
        ChunkIter nxt = cur;
        nxt.next(p, b_blk);
        if (BS)
            for (int s = 1; s < nzb && !nxt.done(p); s++) nxt.next(p, b_blk);
        const int st = it & 1;
        // prefetch the next run into the other stage (its last readers passed the barrier below)
        if (staged && !nxt.done(p) && threadIdx.x == 0) issue(nxt, st ^ 1);
        const int eA = cur.eA;
        const int nt = cur.nt(p);
        const int b_tile0 = cur.b_t0 + cur.c0;
        const int a_t0 = active ? p.A.tile_start[a_blk * p.A.ns + eA] : 0;
        const int Ta = active ? p.A.tile_start[a_blk * p.A.ns + eA + 1] - a_t0 : 0;
        // unstaged (tiles too long for a shared-memory stage): B fragments are read straight from global memory
        const double *sB = staged ? smem + (size_t)st * stage_doubles : p.B.tiles + (size_t)b_tile0 * p.B.tile_doubles;
        if (staged) mbar_wait(&full[st], (unsigned)(it >> 1) & 1u);

        for (int ta = p.split ? warp : 0; ta < Ta; ta += p.split ? nwarps : 1) {
            // A fragments of this slab: registers for the whole sweep over the chunk
            double a[NV1][KS];
            const double2 *at = reinterpret_cast<const double2 *>(p.A.tiles + (size_t)(a_t0 + ta) * p.A.tile_doubles);
            auto load_a = [&](int ch) {
#pragma unroll
                for (int v = 0; v < NV1; v++) {
                    const int tv = (v == 0) ? 0 : (voffA + v);
#pragma unroll
                    for (int h = 0; h < HALF; h++) {
                        const double2 t = __ldg(at + (tv * halft + ch * HALF + h) * 32 + lane);
                        a[v][2 * h] = t.x;
                        a[v][2 * h + 1] = t.y;
                    }
                }
            };
            if (!CHUNKED) load_a(0);
            const unsigned va = p.A.tile_valid[a_t0 + ta];
            const bool a_ok = (va >> ar) & 1u;

            for (int tb = 0; tb < nt; tb++) {
                double H[NV1][NV2][2];
#pragma unroll
                for (int v1 = 0; v1 < NV1; v1++)
#pragma unroll
                    for (int v2 = 0; v2 < NV2; v2++) H[v1][v2][0] = H[v1][v2][1] = 0.0;
                // K_EE has ONE accumulator pair per tile pair, so its KS DMMAs form one dependent chain: in the latency-bound
                // cases the odd k-steps go to a second pair (two chains of KS/2) that is added at the end
                double G0 = 0.0, G1 = 0.0;

                const double2 *bt = reinterpret_cast<const double2 *>(sB + (size_t)tb * p.B.tile_doubles);
                for (int ch = 0; ch < nch; ch++) {
                    if (CHUNKED) load_a(ch);
#pragma unroll
                    for (int h = 0; h < HALF; h++) {
                        // VK: the unrolled steps beyond p.kc are skipped under a uniform predicate (their operands are zero
                        // padding: same bits either way).  Measured at 10 002 force rows (tools/ks14_probe.py): K_FF Dot d = 30 /
                        // 40 / 50: 63.0 -> 39.6 / 47.6 / 59.5 ms; with nothing to skip the predicates cost K_EE 18 % and RBF
                        // K_FF 5 %, hence a separate instantiation that launch_one picks only where it pays
                        if (VARK && (!CHUNKED || ch == nch - 1) && 2 * h >= p.kc) continue;
                        // one 128-bit LDS per B vector = operands of k-steps 2h and 2h+1; the 16 accumulators
                        // of a k-step are independent, so consecutive DMMAs never wait on each other
                        double2 b[NV2];
#pragma unroll
                        for (int v2 = 0; v2 < NV2; v2++) {
                            const int tv = (v2 == 0) ? 0 : (voffB + v2);
                            b[v2] = bt[(tv * halft + ch * HALF + h) * 32 + lane];
                        }
#pragma unroll
                        for (int v2 = 0; v2 < NV2; v2++)
#pragma unroll
                            for (int v1 = 0; v1 < NV1; v1++) dmma(H[v1][v2][0], H[v1][v2][1], a[v1][2 * h], b[v2].x);
                        if (VARK && (!CHUNKED || ch == nch - 1) && 2 * h + 1 >= p.kc) continue;
                        if (DUAL) {
                            dmma(G0, G1, a[0][2 * h + 1], b[0].y);
                        } else {
#pragma unroll
                            for (int v2 = 0; v2 < NV2; v2++)
#pragma unroll
                                for (int v1 = 0; v1 < NV1; v1++) dmma(H[v1][v2][0], H[v1][v2][1], a[v1][2 * h + 1], b[v2].y);
                        }
                    }
                }
                if (DUAL) { H[0][0][0] += G0; H[0][0][1] += G1; }

                const unsigned vb = __ldg(p.B.tile_valid + b_tile0 + tb);
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const bool ok = a_ok && ((vb >> (bq + e)) & 1u);
                    const double c = H[0][0][e];
                    if (MODE == MODE_KFF && FAM == FAM_DOT && Z2) {
                        // zeta = 2: F = c*T + P*Q; padded rows are all-zero so no masking is needed
#pragma unroll
                        for (int k = 0; k < 3; k++)
#pragma unroll
                            for (int l = 0; l < 3; l++) {
                                acc[e][k][l] = fma(c, H[k + 1][l + 1][e], acc[e][k][l]);
                                acc[e][k][l] = fma(H[k + 1][0][e], H[0][l + 1][e], acc[e][k][l]);
                            }
                    } else {
                        const Weights w = pair_weights<MODE, FAM, Z2>(c, ok, p);
                        if (MODE == MODE_KFF) {
#pragma unroll
                            for (int k = 0; k < 3; k++) {
                                const double pk = H[k + 1][0][e];
                                const double w2p = w.w2 * pk;
                                const double g2p = GRAD ? w.g2 * pk : 0.0;
#pragma unroll
                                for (int l = 0; l < 3; l++) {
                                    acc[e][k][l] = fma(w.w1, H[k + 1][l + 1][e], acc[e][k][l]);
                                    acc[e][k][l] = fma(w2p, H[0][l + 1][e], acc[e][k][l]);
                                    if (GRAD) {
                                        accg[e][k][l] = fma(w.g1, H[k + 1][l + 1][e], accg[e][k][l]);
                                        accg[e][k][l] = fma(g2p, H[0][l + 1][e], accg[e][k][l]);
                                    }
                                }
                            }
                        } else if (MODE == MODE_KEF) {
#pragma unroll
                            for (int l = 0; l < 3; l++) {
                                acc[e][0][l] = fma(w.w1, H[0][l + 1][e], acc[e][0][l]);
                                if (GRAD) accg[e][0][l] = fma(w.g1, H[0][l + 1][e], accg[e][0][l]);
                            }
                        } else {
                            acc[e][0][0] += w.w1;
                            if (GRAD) accg[e][0][0] += w.g1;
                        }
                    }
                }
            }
        }
        __syncthreads();   // every warp is done with stage st before it is refilled two runs later
        cur = nxt;
    }

    if (p.split && nwarps > 1) {
        // all warps worked on the same (a_blk, b_blk): sum their partial accumulators in warp 0.
        // (the stage buffers are free: every warp passed the last __syncthreads of the run loop)
        constexpr int NACC = 2 * NC1 * NC2 * (GRAD ? 2 : 1);
        double *red = smem + ((size_t)(warp > 0 ? warp - 1 : 0) * 32 + lane) * NACC;
        if (warp > 0) {
            int n = 0;
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
                for (int k = 0; k < NC1; k++)
#pragma unroll
                    for (int l = 0; l < NC2; l++) {
                        red[n++] = acc[e][k][l];
                        if (GRAD) red[n++] = accg[e][k][l];
                    }
        }
        __syncthreads();
        if (warp > 0) return;
        for (int w = 1; w < nwarps; w++) {
            const double *r = smem + ((size_t)(w - 1) * 32 + lane) * NACC;
            int n = 0;
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
                for (int k = 0; k < NC1; k++)
#pragma unroll
                    for (int l = 0; l < NC2; l++) {
                        acc[e][k][l] += r[n++];
                        if (GRAD) accg[e][k][l] += r[n++];
                    }
        }
    }
    if (!active) return;
    if (GRAD && p.trace) {
        const int gi = p.A.slot_group[a_blk * 8 + ar];
        double sum = 0.0;
        if (gi >= 0) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int gj = p.B.slot_group[b_blk * 8 + bq + e];
                if (gj < 0) continue;
#pragma unroll
                for (int k = 0; k < NC1; k++) {
                    const long long r = p.tr_row0 + (p.tr_packed ? (long long)a_blk * 8 + ar : (long long)gi) * NC1 + k;
                    const double ar_ = p.tr_alpha[r];
#pragma unroll
                    for (int l = 0; l < NC2; l++) {
                        const long long c = p.tr_col0 + (p.tr_packed ? (long long)b_blk * 8 + bq + e : (long long)gj) * NC2 + l;
                        const long long lo = r < c ? r : c, hi = r < c ? c : r;
                        const double kinv = p.tr_packed ? p.tr_w[(r - p.tr_slab0) * p.tr_ld + c] : p.tr_w[lo * p.tr_ld + hi];
                        sum = fma(accg[e][k][l], ar_ * p.tr_alpha[c] - kinv, sum);
                    }
                }
            }
        }
        if (p.symmetric && b_blk != a_blk) sum *= 2.0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0)
            p.tr_partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * kMaxWarps + warp] = p.scale_grad * sum;
        return;
    }
    if (p.sharded == 2 || (p.sharded == 1 && p.packed_cols)) {
        // Packed column order (the layouts the sharded Cholesky factorises): the warp's (8 NC1) x (8 NC2) output tile is
        // contiguous in both directions, so it goes through shared memory and leaves as full rows -- 192-byte runs for
        // K_FF, whole 32-byte sectors when the row pitch is a multiple of 4 doubles -- to the local rows and, transposed,
        // to the rows of the B groups on their owner (NVLink peer stores of 192-byte runs instead of 24-byte pieces).
        constexpr int R = 8 * NC1, C = 8 * NC2, TS = R * (C + 1);
        constexpr int NACC = 2 * NC1 * NC2 * (GRAD ? 2 : 1);
        // behind the scratch of the split-mode reduction; the stage buffers are free (every warp passed the last
        // __syncthreads of the run loop and no bulk copy is outstanding)
        double *T = smem + (size_t)(kMaxWarps - 1) * 32 * NACC + (size_t)warp * TS;
#pragma unroll
        for (int e = 0; e < 2; e++)
#pragma unroll
            for (int k = 0; k < NC1; k++)
#pragma unroll
                for (int l = 0; l < NC2; l++) T[(ar * NC1 + k) * (C + 1) + (bq + e) * NC2 + l] = p.scale * acc[e][k][l];
        __syncwarp();
        const int rv = min(8, p.A.m - a_blk * 8) * NC1, cv = min(8, p.B.m - b_blk * 8) * NC2;   // padding slots trail
        double *mine = p.outs[p.rank], *peer;
        long long row_i0, row_j0, col_i0, col_j0;
        bool mirror;
        if (p.sharded == 2) {
            const int sA = p.pa.sb0 + a_blk / p.pa.g, sB = p.pb.sb0 + b_blk / p.pb.g;
            row_i0 = (long long)(sA / p.nranks) * p.nbw + (a_blk % p.pa.g) * R;
            col_i0 = p.pa.col0 + (long long)a_blk * R;
            peer = p.outs[sb_owner(sB, p.nranks, 1)];
            row_j0 = (long long)(sB / p.nranks) * p.nbw + (b_blk % p.pb.g) * C;
            col_j0 = p.pb.col0 + (long long)b_blk * C;
            mirror = p.pair || b_blk != a_blk;
        } else {
            const int sb = b_blk / p.sb_blocks;
            const int vb = p.sb_inv ? p.sb_inv[sb] : sb;          // virtual superblock of the B rows
            row_i0 = (long long)a_idx * R;                        // a_idx is this block's slot in the rank's buffer
            col_i0 = (long long)a_blk * R;
            peer = p.outs[sb_owner(vb, p.nranks, p.zigzag)];
            row_j0 = ((long long)(vb / p.nranks) * p.sb_blocks + (b_blk - sb * p.sb_blocks)) * C;
            col_j0 = (long long)b_blk * C;
            mirror = b_blk != a_blk;
        }
        for (int idx = lane; idx < R * C; idx += 32) {
            const int r = idx / C, c = idx - r * C;
            if (r < rv && c < cv) mine[(row_i0 + r) * p.ldo + col_j0 + c] = T[r * (C + 1) + c];
        }
        if (mirror)
            for (int idx = lane; idx < R * C; idx += 32) {
                const int c = idx / R, r = idx - c * R;
                if (r < rv && c < cv) peer[(row_j0 + c) * p.ldo + col_i0 + r] = T[r * (C + 1) + c];
            }
        return;
    }
    const int gi = p.A.slot_group[a_blk * 8 + ar];
    if (gi < 0) return;
    if (p.diag) {
#pragma unroll
        for (int e = 0; e < 2; e++)
            if (bq + e == ar) {
#pragma unroll
                for (int k = 0; k < NC1; k++)
#pragma unroll
                    for (int l = 0; l < NC2; l++) p.out[((long long)gi * NC1 + k) * NC2 + l] = p.scale * acc[e][k][l];
            }
        return;
    }
    if (p.sharded) {
        // sharded symmetric build: rows are block-cyclic over ranks in packed order, columns keep the
        // caller's group ids.  Own tile -> local HBM, mirrored tile -> peer HBM of the owner of b.
        double *mine = p.outs[p.rank];
        const int sb = b_blk / p.sb_blocks;
        const int vb = p.sb_inv ? p.sb_inv[sb] : sb;
        const long long lrow_i = (long long)a_idx * 8 + ar;
        const long long ci = gi;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int gj = p.B.slot_group[b_blk * 8 + bq + e];
            if (gj < 0) continue;
            double *peer = p.outs[sb_owner(vb, p.nranks, p.zigzag)];
            const long long lrow_j = ((long long)(vb / p.nranks) * p.sb_blocks + (b_blk - sb * p.sb_blocks)) * 8 + bq + e;
            const long long cj = gj;
#pragma unroll
            for (int k = 0; k < NC1; k++)
#pragma unroll
                for (int l = 0; l < NC2; l++) {
                    const double v = p.scale * acc[e][k][l];
                    mine[lrow_i * p.si + k * p.sk + cj * p.sj + l * p.sl] = v;
                    if (b_blk != a_blk) peer[lrow_j * p.si + l * p.sk + ci * p.sj + k * p.sl] = v;
                }
        }
        return;
    }
#pragma unroll
    for (int e = 0; e < 2; e++) {
        const int gj = p.B.slot_group[b_blk * 8 + bq + e];
        if (gj < 0) continue;
#pragma unroll
        for (int k = 0; k < NC1; k++)
#pragma unroll
            for (int l = 0; l < NC2; l++) {
                const long long o = gi * p.si + k * p.sk + gj * p.sj + l * p.sl;
                double v = p.scale * acc[e][k][l];
                if (p.accumulate) v += out[o];
                if (p.stream_out) __stcs(out + o, v); else out[o] = v;
                if (GRAD) {
                    double g = p.scale_grad * accg[e][k][l];
                    if (p.accumulate) g += p.out_grad[o];
                    p.out_grad[o] = g;
                }
            }
        if (p.symmetric && b_blk != a_blk) {
            // mirror: K[(j,l),(i,k)] = K[(i,k),(j,l)]   (only used for square K_EE / K_FF builds)
#pragma unroll
            for (int k = 0; k < NC1; k++)
#pragma unroll
                for (int l = 0; l < NC2; l++) {
                    const long long o = gj * p.si + l * p.sk + gi * p.sj + k * p.sl;
                    double v = p.scale * acc[e][k][l];
                    if (p.accumulate) v += out[o];
                    if (p.stream_out) __stcs(out + o, v); else out[o] = v;
                    if (GRAD) {
                        double g = p.scale_grad * accg[e][k][l];
                        if (p.accumulate) g += p.out_grad[o];
                        p.out_grad[o] = g;
                    }
                }
        }
    }
}

// out[i, k, j, l] (caller strides) = [accumulate ? out : 0] + sum_z planes[z][i, k, j, l] (dense planes), z in order
__global__ void plane_reduce_kernel(const double *__restrict__ planes, long long plane, int nz, int nc1, int m2, int nc2, double *__restrict__ out,
                                    long long si, long long sk, long long sj, long long sl, int accumulate)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= plane) return;
    double s = 0.0;
    for (int z = 0; z < nz; z++) s += planes[(long long)z * plane + idx];
    const int l = (int)(idx % nc2);
    long long r = idx / nc2;
    const int j = (int)(r % m2);
    r /= m2;
    const int k = (int)(r % nc1);
    const long long i = r / nc1;
    const long long o = i * si + k * sk + j * sj + l * sl;
    out[o] = accumulate ? out[o] + s : s;
}

// ------------------------------------------------------------------------------------------
template <int MODE, int FAM, int KS, bool Z2, bool CHUNKED, bool BS = false, bool VK = false>
static int launch_one(const TileParams &p, int nwarps, size_t smem, cudaStream_t st)
{
    if constexpr (!VK) {      // d = 9 / 12 / 18 need 3 / 3 / 5 of the 6 k-steps, d = 30 / 40 / 50 need 8 / 10 / 13 of the 14; chunked: the LAST chunk of d = 60 / 75 needs 3 / 1 of its 6
        // (one skippable step pays only for the unchunked Dot K_FF kernels: -14 % at d = 18, -6 % at d = 50; the others lose 1-2 % to the
        // predicates.  Chunked, 10 002 force rows: d = 60 / 75: 74.7 / 99.4 ms against 86.5 / 115.4 ms for the padded 72 / 96)
        if (p.kc <= KS - 2 || (p.kc == KS - 1 && MODE == MODE_KFF && FAM == FAM_DOT && !CHUNKED)) return launch_one<MODE, FAM, KS, Z2, CHUNKED, BS, true>(p, nwarps, smem, st);
    }
    if (!BS && p.bsplit > 1) {      // build_block_device only asks for the split where these instantiations exist
        if (FAM != FAM_RBF_GRAD && !CHUNKED) return launch_one<MODE, (FAM == FAM_RBF_GRAD ? FAM_RBF : FAM), KS, Z2, false, true>(p, nwarps, smem, st);
        set_error("block_build: B-run split without an instantiation");
        return CSPBO_E_ARG;
    }
    auto kern = block_tile_kernel<MODE, FAM, KS, Z2, CHUNKED, BS, VK>;
    // the opt-in is a per-device attribute of the function: remember it per device ordinal (a thread may switch
    // devices through cspbo_set_device); races between threads only repeat a cheap idempotent call
    static size_t configured[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    size_t &conf = configured[(dev >= 0 && dev < 64) ? dev : 0];
    if (smem > 48 * 1024 && (smem > conf || dev < 0 || dev >= 64)) {
        CSPBO_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        conf = smem;
    }
    const int a_mine = p.sharded ? p.a_slots : (p.a_nblk - p.a_begin);
    const int npg = p.split ? a_mine : (a_mine + nwarps - 1) / nwarps;
    const long long grid = (long long)npg * p.b_nblk;
    if (grid <= 0) return CSPBO_OK;
    if (grid > 0x7fffffffLL) { set_error("block_build: grid too large"); return CSPBO_E_ARG; }
    kern<<<dim3((unsigned)grid, (unsigned)std::max(std::max(p.npass, p.bsplit), 1)), nwarps * 32, smem, st>>>(p);
    g_launches++;
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("tile kernel launch failed: %s (grid %lld x %d, block %d, smem %zu, tbc %d)", cudaGetErrorString(e), grid,
                      p.npass, nwarps * 32, smem, p.tbc);
            return CSPBO_E_CUDA;
        }
    }
    return CSPBO_OK;
}

template <int MODE, int KS, bool CHUNKED>
static int launch_mode(int fam, bool z2, const TileParams &p, int nwarps, size_t smem, cudaStream_t st)
{
    if (fam == FAM_DOT) return z2 ? launch_one<MODE, FAM_DOT, KS, true, CHUNKED>(p, nwarps, smem, st)
                                  : launch_one<MODE, FAM_DOT, KS, false, CHUNKED>(p, nwarps, smem, st);
    if (fam == FAM_RBF) return z2 ? launch_one<MODE, FAM_RBF, KS, true, CHUNKED>(p, nwarps, smem, st)
                                  : launch_one<MODE, FAM_RBF, KS, false, CHUNKED>(p, nwarps, smem, st);
    return z2 ? launch_one<MODE, FAM_RBF_GRAD, KS, true, CHUNKED>(p, nwarps, smem, st)
              : launch_one<MODE, FAM_RBF_GRAD, KS, false, CHUNKED>(p, nwarps, smem, st);
}

template <int KS, bool CHUNKED>
static int launch_ks(int mode, int fam, bool z2, const TileParams &p, int nwarps, size_t smem, cudaStream_t st)
{
    switch (mode) {
        case MODE_KEE: return launch_mode<MODE_KEE, KS, CHUNKED>(fam, z2, p, nwarps, smem, st);
        case MODE_KEF: return launch_mode<MODE_KEF, KS, CHUNKED>(fam, z2, p, nwarps, smem, st);
        default: return launch_mode<MODE_KFF, KS, CHUNKED>(fam, z2, p, nwarps, smem, st);
    }
}

// Panel-layout arguments of build_block_device (cspbo_block_build_panel).
struct PanelArgs { long long ld; int nbw; long long row0_a, row0_b; };
// Trace mode of the RBF dK/dl kernels (cspbo_block_trace).
struct TraceArgs { const double *alpha, *w; double *partial; long long ld, row0, col0; int packed; long long slab0; };
// Stress passes of one launch: npass column triplets of side `side` (0 = A, 1 = B), pass q stored at out + q*off.
struct PassArgs { int npass, side; long long off; };

// Number of superblocks (a ragged last one included) owned by `rank`.
int sharded_superblocks(int n_blocks, int rank, int nranks, int sb_blocks, int zigzag)
{
    const int nsb = (n_blocks + sb_blocks - 1) / sb_blocks;
    int mine = 0;
    for (int s = 0; s < nsb; s++) mine += sb_owner(s, nranks, zigzag) == rank;
    return mine;
}

// Build one block between packs a and b into DEVICE memory `out` (accumulating).
int build_block_device(const cspbo_block_desc &d, const cspbo_pack *a, const cspbo_pack *b, int voffA, int voffB,
                       double *out, double *out_grad, long long si, long long sk, long long sj, long long sl,
                       int accumulate, cudaStream_t st, int rank, int nranks, double *const *outs, int a_begin, int a_end,
                       int diag, int sb_blocks, int zigzag, int packed_cols, const PanelArgs *panel, const PassArgs *passes,
                       const TraceArgs *trace, const int *sb_perm, const int *sb_inv)
{
    if (a->ks != b->ks || a->d != b->d) {
        set_error("block_build: packs have different descriptor lengths (%d vs %d)", a->d, b->d);
        return CSPBO_E_ARG;
    }
    const int mode = d.block;
    int fam = d.family == CSPBO_DOT ? FAM_DOT : (d.with_grad ? FAM_RBF_GRAD : FAM_RBF);
    const bool z2 = (d.zeta == 2.0);
    TileParams p;
    p.A = {a->d_tiles, a->d_tile_start, a->d_slot_group, a->d_tile_valid, (int)a->tile_doubles, voffA, (int)a->species.size(), a->m};
    p.B = {b->d_tiles, b->d_tile_start, b->d_slot_group, b->d_tile_valid, (int)b->tile_doubles, voffB, (int)b->species.size(), b->m};
    p.trace = trace ? 1 : 0;
    p.tr_alpha = trace ? trace->alpha : nullptr; p.tr_w = trace ? trace->w : nullptr; p.tr_partial = trace ? trace->partial : nullptr;
    p.tr_ld = trace ? trace->ld : 0; p.tr_row0 = trace ? trace->row0 : 0; p.tr_col0 = trace ? trace->col0 : 0;
    p.tr_packed = trace ? trace->packed : 0; p.tr_slab0 = trace ? trace->slab0 : 0;
    p.npass = passes ? std::max(passes->npass, 1) : 1;
    p.pass_side = passes ? passes->side : 0;
    p.pass_off = passes ? passes->off : 0;
    p.zeta = d.zeta; p.sigma2 = d.sigma2; p.sigma02 = d.sigma02;
    p.zint = (d.zeta >= 1.0 && d.zeta <= 8.0 && d.zeta == (double)(int)d.zeta) ? (int)d.zeta : 0;
    p.inv2l2 = d.family == CSPBO_RBF ? 1.0 / (2.0 * d.l * d.l) : 0.0;
    p.inv_l = d.family == CSPBO_RBF ? 1.0 / d.l : 0.0;
    p.inv_l3 = d.family == CSPBO_RBF ? p.inv_l / (d.l * d.l) : 0.0;
    p.tol = d.tol; p.use_tol = d.use_tol;
    p.scale = d.scale; p.scale_grad = d.scale_grad;
    p.out = out; p.out_grad = out_grad;
    p.si = si; p.sk = sk; p.sj = sj; p.sl = sl;
    p.symmetric = d.symmetric;
    p.accumulate = accumulate;
    p.diag = diag;
    p.stream_out = 0;   // streaming stores measured neutral (time and DRAM traffic)
    p.sharded = outs != nullptr;
    p.nranks = p.sharded ? nranks : 1;
    p.rank = p.sharded ? rank : 0;
    for (int r = 0; r < kMaxRanks; r++) p.outs[r] = (p.sharded && r < nranks) ? outs[r] : nullptr;
    p.sb_blocks = std::max(sb_blocks, 1);
    p.zigzag = zigzag ? 1 : 0;
    p.packed_cols = packed_cols ? 1 : 0;
    p.a_begin = p.sharded ? 0 : std::max(a_begin, 0);
    p.a_nblk = (a_end < 0 || p.sharded) ? (int)a->n_blocks : std::min(a_end, (int)a->n_blocks);
    p.a_slots = p.sharded ? sharded_superblocks((int)a->n_blocks, p.rank, p.nranks, p.sb_blocks, p.zigzag) * p.sb_blocks : 0;
    p.a_slot0 = 0;
    p.sb_perm = (p.sharded && !panel) ? sb_perm : nullptr;
    p.sb_inv = (p.sharded && !panel) ? sb_inv : nullptr;
    if (p.sb_perm) {      // permuted ownership: every rank enumerates all cycles, slots without a superblock are skipped
        const int nsb = ((int)a->n_blocks + p.sb_blocks - 1) / p.sb_blocks;
        p.a_slots = ((nsb + p.nranks - 1) / p.nranks) * p.sb_blocks;
    }
    if (p.sharded && !panel && a_end >= 0) {      // a slab [a_begin, a_end) of this rank's block slots
        p.a_slot0 = std::min(std::max(a_begin, 0), p.a_slots);
        p.a_slots = std::max(std::min(a_end, p.a_slots) - p.a_slot0, 0);
        if (p.a_slots == 0) return CSPBO_OK;
    }
    p.nbw = 0; p.cyc0 = 0; p.pair = 0;
    // packed-column sharded layout: the distributed matrix is [rows * NC1, ldo] (K_FF: sk = row pitch, K_EE: si)
    p.ldo = (mode == MODE_KFF) ? sk : si;
    p.pa = {1, 0, 0}; p.pb = {1, 0, 0};
    if (panel && p.sharded) {
        const int nca = (mode == MODE_KFF) ? 3 : 1, ncb = (mode == MODE_KEE) ? 1 : 3;
        p.sharded = 2;
        p.zigzag = 1;
        p.nbw = panel->nbw; p.ldo = panel->ld;
        p.pa = {panel->nbw / (8 * nca), (int)(panel->row0_a / panel->nbw), panel->row0_a};
        p.pb = {panel->nbw / (8 * ncb), (int)(panel->row0_b / panel->nbw), panel->row0_b};
        p.pair = (a != b) ? 1 : 0;
        const int nsbA = ((int)a->n_blocks + p.pa.g - 1) / p.pa.g;
        p.cyc0 = p.pa.sb0 / p.nranks;
        const int cyc1 = (p.pa.sb0 + std::max(nsbA, 1) - 1) / p.nranks;
        p.a_slots = (cyc1 - p.cyc0 + 1) * p.pa.g;
    }
    p.b_nblk = diag ? 1 : (int)b->n_blocks;
    if (p.a_nblk <= p.a_begin || b->n_blocks == 0) return CSPBO_OK;

    int maxTb = 1;
    const int nsA = (int)a->species.size(), nsB = (int)b->species.size();
    std::vector<int> emap((size_t)std::max(nsA, kMaxSpecies), -1);
    for (int ea = 0; ea < nsA; ea++)
        for (int eb = 0; eb < nsB; eb++)
            if (b->species[(size_t)eb] == a->species[(size_t)ea]) {
                emap[(size_t)ea] = eb;
                for (int q = 0; q < (int)b->n_blocks; q++)
                    maxTb = std::max(maxTb, b->h_tile_start[(size_t)q * nsB + eb + 1] - b->h_tile_start[(size_t)q * nsB + eb]);
            }
    for (int e = 0; e < kMaxSpecies; e++) p.emap[e] = emap[(size_t)e];
    p.emap_dev = nullptr;
    int *emap_dev = nullptr;
    if (nsA > kMaxSpecies) {
        // more species buckets than the kernel parameters carry (labels beyond the usual handful of elements): the map
        // goes through device memory.  Rare, so a blocking copy and a stream sync after the launch are acceptable.
        void *q = nullptr;
        const int rca = dev_alloc_cached((size_t)nsA * sizeof(int), &q);
        if (rca) return rca;
        emap_dev = static_cast<int *>(q);
        cudaError_t e = cudaMemcpy(emap_dev, emap.data(), (size_t)nsA * sizeof(int), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { dev_free_cached(emap_dev); set_error("block_build: species map upload failed: %s", cudaGetErrorString(e)); return CSPBO_E_CUDA; }
        p.emap_dev = emap_dev;
    }

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // warps per CTA: one A block per warp; shrink for small problems so the grid covers the SMs
    int nwarps = kMaxWarps;
    const int a_mine = p.sharded ? p.a_slots : (p.a_nblk - p.a_begin);
    // Small problems (fewer than ~2 warps per scheduler in the one-block-per-warp mapping): let the 4 warps
    // of a CTA split the tiles of ONE A block instead, which quarters the critical path of a K* build
    // (decided on the whole A side, not on the row slab of a pipelined host-bound build, so that every
    // way of building the same block sums in the same order and gives bit-identical results)
    const long long a_all = ((long long)a->n_blocks + p.nranks - 1) / p.nranks;
    p.split = (diag || a_all * p.b_nblk * p.npass < 32LL * sms) ? 1 : 0;
    // B tiles per stage: a whole (block, species) run if it fits in 36 KB
    const size_t tile_bytes = (size_t)b->tile_doubles * sizeof(double);
    // K_FF: 36 KB stages (a whole (block, species) run of 6 KB tiles).  The narrower K_EE / K_EF tiles need less, and
    // smaller stages raise the resident CTAs per SM (measured: K_EE RBF 18.4 -> 16.8 ms at 18 KB, K_EF RBF 75.7 -> 73.7 at 24 KB)
    size_t stage_bytes = mode == MODE_KEE ? (size_t)18 * 1024 : (mode == MODE_KEF ? (size_t)24 * 1024 : kSmemChunk);
    // chunked kernels (d > 56): tiles are nch x longer; a stage must hold at least one whole B tile.  Up to 96 KB per
    // stage they are staged like the short ones, beyond that (d > ~400, or d > ~160 with 9 derivative columns on the B
    // side) the B fragments are read straight from global memory
    const bool chunked = a->ks != 6 && a->ks != 14;
    p.nch = chunked ? a->ks / 6 : 1;
    p.kc = std::min((a->d + 3) / 4, (int)a->ks) - (chunked ? 6 * (p.nch - 1) : 0);      // chunked: k-steps of the last chunk
    p.staged = 1;
    if (chunked && tile_bytes > stage_bytes) {
        if (tile_bytes <= (size_t)96 * 1024) stage_bytes = tile_bytes;
        else p.staged = 0;
    }
    int tbc = p.staged ? (int)std::max<size_t>(1, std::min<size_t>((size_t)maxTb, stage_bytes / tile_bytes)) : maxTb;
    // Small problems whose (A block, B block) pairs are few but long (a K_EE of some dozens of 192-atom structures: one
    // CTA per pair walked 0.4 ms alone while most SMs idled): bsplit CTAs share a pair, each taking every bsplit-th run
    // of B tiles, and their partial blocks are summed in a fixed order afterwards (deterministic; no atomics).
    p.bsplit = 1;
    long long plane = 0;
    const int nc1 = (mode == MODE_KFF) ? 3 : 1, nc2 = (mode == MODE_KEE) ? 1 : 3;
    const int bsplit_env = getenv("CSPBO_BSPLIT") ? atoi(getenv("CSPBO_BSPLIT")) : -1;     // 0 / 1 = off, n = force (tests, experiments)
    if (p.split && !p.sharded && !diag && !trace && p.npass == 1 && fam != FAM_RBF_GRAD && st == nullptr && a_begin <= 0 && a_end < 0 &&
        !chunked && out && bsplit_env != 0 && bsplit_env != 1) {
        // tiles of the heaviest B block and tile pairs of the heaviest block pair
        long long tsum = 0, pairs = 0;
        for (int q = 0; q < (int)b->n_blocks; q++) {
            long long t = 0;
            for (int ea = 0; ea < nsA; ea++)
                if (emap[(size_t)ea] >= 0) t += b->h_tile_start[(size_t)q * nsB + emap[(size_t)ea] + 1] - b->h_tile_start[(size_t)q * nsB + emap[(size_t)ea]];
            tsum = std::max(tsum, t);
        }
        for (int ea = 0; ea < nsA; ea++)
            if (emap[(size_t)ea] >= 0)
                pairs += (long long)(a->h_tile_start[(size_t)ea + 1] - a->h_tile_start[(size_t)ea]) *
                         (b->h_tile_start[(size_t)emap[(size_t)ea] + 1] - b->h_tile_start[(size_t)emap[(size_t)ea]]);     // block 0 x block 0 (the fullest)
        const long long ctas = d.symmetric ? (long long)a_mine * (p.b_nblk + 1) / 2 : (long long)a_mine * p.b_nblk;
        const long long dmma_per_cta = pairs * (a->ks) * (mode == MODE_KFF ? 16 : (mode == MODE_KEF ? 4 : 1));
        // as many CTAs as the SMs hold at once (6 / 4 / 3 per SM for K_EE / K_EF / K_FF at their register and stage sizes)
        const long long resident = (mode == MODE_KEE ? 6LL : (mode == MODE_KEF ? 4LL : 3LL)) * sms;
        int nz = (int)std::min<long long>(16, resident / std::max<long long>(ctas, 1));
        if (bsplit_env > 1) nz = bsplit_env;
        plane = (long long)a->m * nc1 * b->m * nc2;
        if (nz >= 2 && tsum >= 2 * nz && (dmma_per_cta >= 20000 || bsplit_env > 1) && plane * nz * (long long)sizeof(double) <= (64LL << 20)) {
            tbc = (int)std::max<long long>(1, std::min<long long>(tbc, tsum / (2 * nz)));      // at least ~2 runs per CTA
            p.bsplit = nz;
        }
    }
    p.tbc = tbc;
    // the stage buffers double as the scratch of the split-mode reduction (NACC accumulators per thread of 3 warps)
    const int nacc = 2 * (mode == MODE_KFF ? 9 : (mode == MODE_KEF ? 3 : 1)) * (fam == FAM_RBF_GRAD ? 2 : 1);
    // ... and, for the packed-column layouts, as the staging area of the full-row output stores (behind that scratch)
    const int oR = 8 * (mode == MODE_KFF ? 3 : 1), oC = 8 * (mode == MODE_KEE ? 1 : 3);
    const size_t out_stage = (p.sharded == 2 || (p.sharded == 1 && p.packed_cols)) ? (size_t)kMaxWarps * oR * (oC + 1) * sizeof(double) : 0;
    const size_t smem = std::max<size_t>(p.staged ? 2 * (size_t)tbc * tile_bytes : 0,
                                         (size_t)(kMaxWarps - 1) * 32 * nacc * sizeof(double) + out_stage) +
                        2 * sizeof(unsigned long long);
    // Rasterisation.  Caller-order outputs: one strip (kStrip).  Packed-column layouts: strips of 48 A-block groups
    // (192 row blocks): the strip's A tiles stay L2-hot while the B blocks stream, and -- now that the output leaves as
    // whole sectors -- DRAM reads drop from 3.8 GB to 0.30 GB per 20 001-row build at unchanged run time (measured,
    // profiles/kff_traffic_r02.md); with 24-byte output pieces the same strips cost 1.5 % (round 1's finding).
    static const int strip_env = getenv("CSPBO_STRIP") ? atoi(getenv("CSPBO_STRIP")) : 0;    // experiments only
    const int strip_def = (p.sharded == 2 || (p.sharded == 1 && p.packed_cols)) ? 48 : kStrip;
    p.strip = std::max(1, std::min(strip_env > 0 ? strip_env : strip_def, p.split ? a_mine : (a_mine + nwarps - 1) / nwarps));
    // Keep the packed tiles L2-resident while the write-once output streams through: an access-policy
    // window (persisting on hit, streaming on miss) over the B tiles.  Measured at 20 001 force rows:
    // DRAM reads 20.5 -> 9.5 GB per build at unchanged run time (profiles/).
    bool window = false;
    {
        static int persist_of[64], window_of[64];
        static bool probed[64] = {};
        const int di = (dev >= 0 && dev < 64) ? dev : 0;
        int &max_persist = persist_of[di], &max_window = window_of[di];
        if (!probed[di]) {
            probed[di] = true;
            cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
            cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
            if (max_persist > 0) cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
            cudaGetLastError();
        }
        const size_t tile_bytes_all = (size_t)b->n_tiles * b->tile_doubles * sizeof(double);
        if (max_persist > 0 && max_window > 0 && tile_bytes_all >= ((size_t)8 << 20)) {
            cudaStreamAttrValue av;
            memset(&av, 0, sizeof(av));
            av.accessPolicyWindow.base_ptr = (void *)b->d_tiles;
            av.accessPolicyWindow.num_bytes = std::min<size_t>(tile_bytes_all, (size_t)max_window);
            av.accessPolicyWindow.hitRatio = 1.0f;
            av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            window = cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av) == cudaSuccess;
            cudaGetLastError();
        }
    }
    // (d in (24, 56], K_FF with pow / exp in the epilogue: the KS = 14 instantiations spill 150-700 bytes around the libm
    // calls; re-reading the A fragments per B tile instead -- the CHUNKED kernel with one chunk -- removes most of the spill
    // but was measured SLOWER, RBF d = 50: 69.8 -> 73.2 ms, tools/ks14_probe.py, so the spilling variants stay)
    double *final_out = p.out;
    const long long fsi = p.si, fsk = p.sk, fsj = p.sj, fsl = p.sl;
    if (p.bsplit > 1) {
        void *sp = nullptr;
        const int rcs = scratch_get(SCRATCH_PLANES, (size_t)plane * p.bsplit * sizeof(double), &sp);
        if (rcs) { if (emap_dev) dev_free_cached(emap_dev); return rcs; }
        p.out = static_cast<double *>(sp);
        p.pass_off = plane;
        p.accumulate = 0;
        p.si = (long long)nc1 * b->m * nc2; p.sk = (long long)b->m * nc2; p.sj = nc2; p.sl = 1;     // dense planes [m1, nc1, m2, nc2]
    }
    int rc_launch = (a->ks == 6) ? launch_ks<6, false>(mode, fam, z2, p, nwarps, smem, st)
                          : (a->ks == 14) ? launch_ks<14, false>(mode, fam, z2, p, nwarps, smem, st)
                                          : launch_ks<6, true>(mode, fam, z2, p, nwarps, smem, st);
    if (p.bsplit > 1 && rc_launch == CSPBO_OK) {
        plane_reduce_kernel<<<(unsigned)((plane + 255) / 256), 256, 0, st>>>(p.out, plane, p.bsplit, nc1, b->m, nc2, final_out, fsi, fsk, fsj, fsl,
                                                                             accumulate);
        g_launches++;
        if (cudaGetLastError() != cudaSuccess) { set_error("plane_reduce_kernel launch failed"); rc_launch = CSPBO_E_CUDA; }
    }
    if (emap_dev) {
        cudaStreamSynchronize(st);
        dev_free_cached(emap_dev);
    }
    if (window) {   // the window is captured at launch; do not leak it onto later work of this stream
        cudaStreamAttrValue av;
        memset(&av, 0, sizeof(av));
        av.accessPolicyWindow.num_bytes = 0;
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &av);
        cudaGetLastError();
    }
    return rc_launch;
}

}  // namespace cspbo
