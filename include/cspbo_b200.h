/*
 * cspbo_b200.h -- C ABI of libcspbo_b200.so, the B200 (sm_100a) implementation of CSP_BO's
 * Gaussian-process kernel-matrix hot path (K_EE / K_EF / K_FF of the Dot and RBF many-body
 * kernels, K*.alpha prediction, Cholesky fit).
 *
 * Plain C: pointers and sizes only, no torch / C++ types.  Every function returns an int
 * status (0 = ok, see CSPBO_E_*) and never throws.  cspbo_last_error() returns a static,
 * thread-local, human readable message for the last non-zero status.
 *
 * Three layers, lowest first:
 *
 *  (1) reference-compatible entry points (host pointers, one call = one block)
 *      cspbo_dot_* / cspbo_rbf_*  -- same argument lists as the reference's cffi cdefs
 *      (reference cspbo/kernels/libdot_builder.py:5-9 and librbf_builder.py:5-12, implemented
 *      by cspbo/kernels/dot_kernel.cpp and rbf_kernel.cpp) plus an int return.  Like the
 *      reference they ACCUMULATE into the caller's `pout` (+= / -=), so `pout` must be
 *      zero-initialised by the caller (dot_kernel.py:36,107-110,219-229).
 *      The two shim libraries libdot_kernel_b200.so / librbf_kernel_b200.so export these under
 *      the reference's exact unprefixed names (kee_many, kef_many, ...) for a cffi/ctypes
 *      binding that wants a drop-in `lib` object.
 *
 *  (2) resident API: pack a stacked descriptor set once into HBM (cspbo_pack_*), then build any
 *      number of blocks against it (cspbo_block_build) into host or device memory.  This is what
 *      the Python kernel objects (csp_bo_b200.kernels.Dot_mb / RBF_mb) use: the training set is
 *      packed once per fit and reused by every prediction.
 *
 *  (3) GP algebra on the device: K + noise -> Cholesky (cuSOLVER) -> alpha, and K*.alpha
 *      (reference cspbo/gaussianprocess.py:105-127,138-141,359-366).
 *
 * Conventions (identical to the reference, dot_kernel.cpp:4-55):
 *   x[n,d]            stacked descriptor rows, row-major float64
 *   dxdr[n,d,nc]      per-row Cartesian derivative, nc = 3 (force) or 9 (3 force + 6 stress)
 *   ele[n]            species id per row (int32); pairs of different species contribute 0
 *   inds[n]           group id per row (structure for "energy" data, force atom for "force" data)
 *   x2i               number of x2 groups (= row stride of pout)
 *   rows with |x| <= 1e-8 are ignored.
 */
#ifndef CSPBO_B200_H
#define CSPBO_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define CSPBO_OK            0
#define CSPBO_E_CUDA        1   /* a CUDA runtime call failed */
#define CSPBO_E_ARG         2   /* invalid argument (null pointer, negative size, d unsupported) */
#define CSPBO_E_NOMEM       3   /* host or device allocation failed */
#define CSPBO_E_SOLVER      4   /* cuSOLVER failure; CSPBO_E_NOTPD for a non positive-definite K */
#define CSPBO_E_NOTPD       5
#define CSPBO_E_NODEVICE    6   /* no CUDA device: there is NO CPU fallback */

#define CSPBO_MAX_D         56  /* largest descriptor length the tile kernels are compiled for (k-steps 6 or 14) */

const char *cspbo_last_error(void);
int cspbo_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
long long cspbo_launch_count(void);
/* make `device` current for subsequent calls from this thread (cudaSetDevice) */
int cspbo_set_device(int device);
int cspbo_device_count(int *count);
int cspbo_synchronize(void);
/* free the cached device scratch buffers used by host-pointer calls (staging of x / dxdr / outputs) */
int cspbo_release_scratch(void);

/* ------------------------------------------------------------------------------------------
 * (1) reference-compatible entry points.  Host pointers.  pout is accumulated into.
 * ------------------------------------------------------------------------------------------ */

/* Dot kernel.  Replaces dot_kernel.cpp:4-55 (kee_many) */
int cspbo_dot_kee_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double sigma02,
                       const double *x1, const int *ele1, const int *x1_inds,
                       const double *x2, const int *ele2, const int *x2_inds, double *pout);
/* dot_kernel.cpp:57-128 */
int cspbo_dot_kef_many(int n1, int n2, int d, int x2i, double zeta,
                       const double *x1, const int *ele1, const int *x1_inds,
                       const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                       double *pout);
/* dot_kernel.cpp:131-214 (dx2dr has 9 columns, pout[m1,m2,9]) */
int cspbo_dot_kef_many_stress(int n1, int n2, int d, int x2i, double zeta,
                              const double *x1, const int *ele1, const int *x1_inds,
                              const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                              double *pout);
/* dot_kernel.cpp:215-334.  Only x2 rows [n2_start, n2_end) are used (the reference's MPI split). */
int cspbo_dot_kff_many(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                       const double *x1, const double *dx1dr, const int *ele1, const int *x1_inds,
                       const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                       double *pout);
/* dot_kernel.cpp:336-506 (dx1dr has 9 columns, pout[m1,9,m2*3]) */
int cspbo_dot_kff_many_stress(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                              const double *x1, const double *dx1dr, const int *ele1, const int *x1_inds,
                              const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                              double *pout);

/* RBF kernel.  rbf_kernel.cpp:4-48 */
int cspbo_rbf_kee_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                       const double *x1, const int *ele1, const int *x1_inds,
                       const double *x2, const int *ele2, const int *x2_inds, double *pout);
/* rbf_kernel.cpp:50-97 */
int cspbo_rbf_kee_many_with_grad(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                                 const double *x1, const int *ele1, const int *x1_inds,
                                 const double *x2, const int *ele2, const int *x2_inds,
                                 double *pout, double *dpout_dl);
/* rbf_kernel.cpp:100-170 */
int cspbo_rbf_kef_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                       const double *x1, const int *ele1, const int *x1_inds,
                       const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                       double *pout);
/* rbf_kernel.cpp:172-252.  NB: takes l, not l^2; pout[m1,m2,6] = (K_EF[3], dK_EF/dl[3]) */
int cspbo_rbf_kef_many_with_grad(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l,
                                 const double *x1, const int *ele1, const int *x1_inds,
                                 const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                                 double *pout);
/* rbf_kernel.cpp:255-337 */
int cspbo_rbf_kef_many_stress(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                              const double *x1, const int *ele1, const int *x1_inds,
                              const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                              double *pout);
/* rbf_kernel.cpp:340-472.  Pairs with dK/dD <= tol are skipped (:394). */
int cspbo_rbf_kff_many(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                       double sigma2, double l2, double tol,
                       const double *x1, const double *dx1dr, const int *ele1, const int *x1_inds,
                       const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                       double *pout);
/* rbf_kernel.cpp:474-639.  NB: takes l, not l^2; no tol test. */
int cspbo_rbf_kff_many_with_grad(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                                 double sigma2, double l,
                                 const double *x1, const double *dx1dr, const int *ele1, const int *x1_inds,
                                 const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                                 double *pout, double *dpout_dl);
/* rbf_kernel.cpp:641-822 */
int cspbo_rbf_kff_many_stress(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                              double sigma2, double l2, double tol,
                              const double *x1, const double *dx1dr, const int *ele1, const int *x1_inds,
                              const double *x2, const double *dx2dr, const int *ele2, const int *x2_inds,
                              double *pout);

/* ------------------------------------------------------------------------------------------
 * (2) resident API
 * ------------------------------------------------------------------------------------------ */
typedef struct cspbo_pack cspbo_pack;   /* a stacked descriptor set, normalised and tiled in HBM */

#define CSPBO_MEM_HOST   0
#define CSPBO_MEM_DEVICE 1

/* Pack rows [row_start,row_end) of a stacked set.
 *   dxdr == NULL  -> "energy" pack (descriptor only: usable as the E side of K_EE / K_EF)
 *   dxdr != NULL  -> "force" pack; nc = 3 or 9 columns per feature (9: 3 force + 6 stress)
 *   m             number of groups (ids in inds must lie in [0,m))
 *   mem           where x / dxdr live (CSPBO_MEM_HOST or CSPBO_MEM_DEVICE); ele / inds are
 *                 always host arrays (the tile layout is planned on the host).
 * Replaces the per-call ffi.new(...) deep copies of dot_kernel.py:207-216. */
int cspbo_pack_create(int n, int d, int nc, int m, int row_start, int row_end,
                      const double *x, const double *dxdr, const int *ele, const int *inds,
                      int mem, cspbo_pack **out);
/* Same, with `norm_shift` added to every row norm and no zero-norm skipping: norm_shift = 1e-8
 * reproduces the eps-shifted NumPy formulas behind the kernels' diag() (Dot_mb.py:177-258). */
int cspbo_pack_create_eps(int n, int d, int nc, int m, int row_start, int row_end,
                          const double *x, const double *dxdr, const int *ele, const int *inds,
                          int mem, double norm_shift, cspbo_pack **out);
/* Full form.  row_classes > 1 splits the groups by (group id % row_classes) before blocking, which
 * lets a host-bound K_FF build copy each finished class back with one strided copy while the next
 * class is being computed (cspbo_block_build with mem = CSPBO_MEM_HOST).  1 = plain layout. */
int cspbo_pack_create_ex(int n, int d, int nc, int m, int row_start, int row_end,
                         const double *x, const double *dxdr, const int *ele, const int *inds,
                         int mem, double norm_shift, int row_classes, cspbo_pack **out);
int cspbo_pack_destroy(cspbo_pack *p);
int cspbo_pack_groups(const cspbo_pack *p);          /* m */
long long cspbo_pack_bytes(const cspbo_pack *p);     /* HBM footprint */
/* same-species valid row pairs between two packs: the unit SURVEY 8(d) counts flops in */
long long cspbo_pack_pair_count(const cspbo_pack *a, const cspbo_pack *b);

#define CSPBO_DOT 0
#define CSPBO_RBF 1

#define CSPBO_KEE 0
#define CSPBO_KEF 1
#define CSPBO_KFF 2

typedef struct {
    int family;        /* CSPBO_DOT | CSPBO_RBF */
    int block;         /* CSPBO_KEE | CSPBO_KEF | CSPBO_KFF */
    double zeta;
    double sigma2;     /* Dot K_EE and all RBF blocks (the reference's C argument) */
    double sigma02;    /* Dot K_EE */
    double l;          /* RBF length scale (l, not l^2) */
    double tol;        /* RBF K_FF skip threshold */
    int use_tol;       /* 1: apply the dK/dD > tol test (kff_many, kff_many_stress) */
    int with_grad;     /* RBF only: also produce the d/dl companion (kee/kef/kff _with_grad) */
    int stress;        /* 1: K_EF -> 9 columns from side b; K_FF -> 9 rows from side a */
    double scale;      /* result multiplier applied on the device (1.0 = raw reference C output) */
    double scale_grad; /* multiplier of the d/dl companion */
    int symmetric;     /* 1: a and b are the same pack; only blocks j>=i are computed and mirrored */
} cspbo_block_desc;

/* out layouts (row-major, identical to the reference C functions):
 *   KEE  out[m1, m2]                            grad -> out_grad[m1, m2]
 *   KEF  out[m1, m2, nc]  nc = 3 | 9 (stress)   grad -> out_grad[m1, m2, 3]   (separate array)
 *   KFF  out[m1, nc1, m2*3] nc1 = 3 | 9         grad -> out_grad[m1, 3, m2*3]
 * mem tells where out/out_grad live.  accumulate = 0 overwrites, 1 adds to the existing content. */
int cspbo_block_build(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b,
                      double *out, double *out_grad, int mem, int accumulate);

/* Build a plain block (no stress / grad) straight into a row-major DEVICE matrix K with leading
 * dimension ldk at offset (row0, col0): K_EE -> [m1, m2], K_EF -> [m1, 3*m2], K_FF -> [3*m1, 3*m2].
 * This is how [[K_EE, K_EF], [K_FE, K_FF]] (base.py:13-14) is assembled without host round trips. */
int cspbo_block_build_into(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b,
                           double *K, long long ldk, long long row0, long long col0);

/* Self covariance of every group of pack `a` (the diagonal blocks of K(a,a)): out[m] for K_EE,
 * out[m,3,3] for K_FF.  Used by diag() for the predictive variance (Dot_mb.py:45-85). */
int cspbo_block_diag(const cspbo_block_desc *desc, const cspbo_pack *a, double *out, int mem);

/* ---- sharded symmetric build: one process per GPU, K(X,X) distributed by row blocks ----------
 * The groups of pack `a` are taken in PACKED order (cspbo_pack_order: position -> group id, -1 for the
 * padding slots of the last block) and cut into blocks of 8; rank r owns blocks r, r+nranks, ...
 * Its buffer outs[r] holds those row blocks contiguously: [cspbo_sharded_rows(a,r,nranks), 3, m*3]
 * for K_FF ([rows, m] for K_EE); columns keep the caller's group order.  Every rank computes only
 * the tiles (block a owned, block b >= a) and stores the mirrored tile directly into the owner's
 * buffer through NVLink peer memory, so outs[] must hold, for every rank, a pointer valid in THIS
 * process (own buffer from cspbo_dev_alloc, peers' through cspbo_ipc_open).  Replaces the reference's
 * MPI split + Allreduce of the whole matrix (dot_kernel.py:188-247).  Callers must barrier across
 * ranks before (buffers allocated) and after (peer stores landed) the call. */
int cspbo_dev_alloc(long long bytes, void **ptr);            /* cudaMalloc: IPC-exportable */
int cspbo_dev_free(void *ptr);
int cspbo_ipc_export(void *ptr, unsigned char *handle64);    /* cudaIpcGetMemHandle (64 bytes) */
int cspbo_ipc_open(const unsigned char *handle64, void **ptr);
int cspbo_ipc_close(void *ptr);
/* Pitched device -> pinned-host copy of `height` rows of `width` BYTES, asynchronous on `stream` (the copy-back of a
 * finished row slab of a host-bound build). */
int cspbo_copy_rows_to_host(void *dst, long long dpitch, const void *src, long long spitch, long long width, long long height,
                            void *stream);
long long cspbo_sharded_rows(const cspbo_pack *a, int rank, int nranks);
int cspbo_pack_order(const cspbo_pack *a, int *order, long long n);   /* n >= 8*ceil(m/8) */
int cspbo_block_build_sharded(const cspbo_block_desc *desc, const cspbo_pack *a, int rank, int nranks,
                              double *const *outs);
/* Generalised ownership for the sharded Cholesky: sb_blocks consecutive 8-group blocks form a superblock
 * (the panel of the factorisation), superblock s belongs to rank s % nranks -- with zigzag != 0 to rank
 * nranks-1 - s % nranks on odd cycles s / nranks, which balances the triangular work for wide panels --
 * and is the (s / nranks)-th superblock of its owner's buffer.  packed_cols != 0 stores the columns in
 * packed order as well (column of group slot p is p, padding slots are the last ones), i.e. the
 * distributed matrix is P K P^T with P = cspbo_pack_order.  (1, 0, 0) is cspbo_block_build_sharded. */
long long cspbo_sharded_rows_ex(const cspbo_pack *a, int rank, int nranks, int sb_blocks, int zigzag);
int cspbo_block_build_sharded_ex(const cspbo_block_desc *desc, const cspbo_pack *a, int rank, int nranks,
                                 double *const *outs, int sb_blocks, int zigzag, int packed_cols);
/* The same with an explicit row pitch: the buffers are [rows * nc, ld] matrices with ld >= 3m (K_FF) or m (K_EE)
 * doubles (ld = 0: dense).  With packed_cols every warp stores its 24 x 24 (8 x 8) output tile as full rows staged
 * through shared memory; a pitch that is a multiple of 4 doubles makes those rows whole 32-byte sectors, for the
 * local stores and for the mirrored tiles that cross NVLink as peer stores. */
int cspbo_block_build_sharded_ld(const cspbo_block_desc *desc, const cspbo_pack *a, int rank, int nranks,
                                 double *const *outs, int sb_blocks, int zigzag, int packed_cols, long long ld);
/* One row slab of the same build, asynchronous on `stream` (a cudaStream_t, NULL = default): only the 8-group block
 * slots [slot0, slot1) of this rank's buffer are computed (slot1 < 0: all).  Slabs taken in ascending order by every
 * rank have this property: once ALL ranks have finished their slab k, the rows of slab k are final on every rank
 * (mirrored tiles only ever come from row blocks that are not later in packed order), so a host-bound build can copy
 * slab k out while slab k+1 is computed (csp_bo_b200/distributed.py: ShardedSymmetricBlock.build_to_host).
 * sb_perm / sb_inv (DEVICE int arrays, both or neither; NULL = identity): work-balanced ownership.  The dealing formula
 * (rank of superblock s, its place in the owner's buffer) is applied to VIRTUAL indices; sb_perm[v] is the real superblock
 * at virtual index v (-1 = none; length ceil(nsb / nranks) * nranks) and sb_inv[s] the virtual index of real superblock s.
 * A permutation that only shuffles superblocks inside each cycle of `nranks` keeps the slab property above. */
int cspbo_block_build_sharded_slab(const cspbo_block_desc *desc, const cspbo_pack *a, int rank, int nranks,
                                   double *const *outs, int sb_blocks, int zigzag, int packed_cols, long long ld,
                                   int slot0, int slot1, void *stream, const int *sb_perm, const int *sb_inv);
/* work[a] (n >= number of 8-group blocks) = tile pairs of the symmetric build that belong to row block a (sum over the
 * blocks b >= a and the species of T(a,e) T(b,e)): the weights a work-balanced ownership is computed from. */
int cspbo_pack_block_work(const cspbo_pack *a, double *work, long long n);

/* Panel layout: the whole covariance [[K_EE, K_EF], [K_FE, K_FF]] (base.py:13-14) as ONE distributed matrix, the input
 * of the sharded Cholesky.  Global rows = columns: each pack in packed order (cspbo_pack_order), starting at a multiple
 * of nbw (row0_a / row0_b; callers pad the energy part with identity rows); rows are dealt to the ranks by superblocks
 * of nbw rows in zigzag order, superblock s being rows [(s / nranks) * nbw, +nbw) of its owner's buffer outs[r], a
 * row-major matrix of leading dimension ld.  block = K_EE / K_FF: a == b, the upper block triangle is computed and
 * mirrored; K_EF: a = energy pack (rows), b = force pack (columns), every tile is also stored transposed (K_FE) on the
 * owner of the force rows.  Mirrored tiles travel as NVLink peer stores; barrier across ranks before and after. */
int cspbo_block_build_panel(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b, int rank, int nranks,
                            double *const *outs, long long ld, int nbw, long long row0_a, long long row0_b);

/* ------------------------------------------------------------------------------------------
 * (3) GP algebra on the device (all pointers are DEVICE pointers, column count = N)
 * ------------------------------------------------------------------------------------------ */
/* K[i,i] += noise_e^2 (i < n_energy) or (f_coef*noise_e)^2 (else)   gaussianprocess.py:109-113 */
int cspbo_gp_add_noise(double *K, int N, int n_energy, double noise_e, double f_coef);
/* In-place lower Cholesky of the symmetric K (cuSOLVER potrf) followed by alpha = K^-1 y (potrs);
 * y[N, nrhs] column-major == row-major for nrhs = 1.   gaussianprocess.py:116,127 */
int cspbo_gp_cholesky_solve(double *K, int N, double *y_inout, int nrhs, double *logdet);
/* Solve (L L^T) X = B with the factor left in K by cspbo_gp_cholesky_solve.  B is [N, nrhs]
 * column-major, i.e. a row-major [nrhs, N] matrix such as K* itself; overwritten with X.
 * gaussianprocess.py:166 (cho_solve) and the K^-1 K*^T of :174,:382 */
int cspbo_gp_cho_solve(const double *L, int N, double *B, int nrhs);
/* Kinv = (L L^T)^-1, full symmetric matrix.  gaussianprocess.py:497 (cho_solve against the identity).  From 1024 rows
 * on the inverse is formed by the recursive sweeps of csrc/dense.cu -- batched leaf inverses + large cuBLAS trmm / syrk
 * products, 2.7x cuSOLVER's potri on B200 at N = 20 001 -- below that (or with CSPBO_DENSE=0) by cuSOLVER potri. */
int cspbo_gp_cholesky_inverse(const double *L, int N, double *Kinv);
/* The same in place: the factor in the buffer becomes K^-1 (row-major upper triangle; mirror != 0 fills both). */
int cspbo_gp_cholesky_inverse_inplace(double *L, int N, int mirror);
/* diag_out[N] = diag((L L^T)^-1), destroying the factor (triangular inverse in place -- csrc/dense.cu, 4.3x cuSOLVER's
 * trtri at N = 20 001 -- + a row-wise sum of squares): all the Dot kernel's hyper-gradient needs of K^-1, with no second
 * N x N array.  gaussianprocess.py:490-499 */
int cspbo_gp_inverse_diag(double *L, int N, double *diag_out);
/* The triangular factor itself inverted in place: the buffer's row-major upper triangle (= column-major lower L) becomes
 * L^-1; the other triangle is scratch.  The diagonal-block inverses of distributed.lml_sharded's blocked substitutions. */
int cspbo_gp_tri_inverse_inplace(double *L, int N);
/* Bytes of device workspace the calls above take from the library's block cache for an N x N factor (leaf results +
 * the chunk buffer of the out-of-place triangular products: about N^2/16 doubles at large N); 0 on the cuSOLVER path. */
long long cspbo_gp_inverse_workspace_bytes(int N);
/* Hyper-gradient trace of an RBF block without materialising dK/dl (gaussianprocess.py:496-499, rbf_kernel.cpp:87-92,
 * 239-246, 621-629): *result (device double) = [accumulate ? *result : 0] +
 *     sum over the block's entries (r, c) of  desc->scale_grad * dK/dl[r, c] * (alpha[r] alpha[c] - W[min(r,c)*ldw + max(r,c)])
 * with r = row0 + i*nc1 + k, c = col0 + j*nc2 + l in the caller's group order (the layout cspbo_block_build_into writes);
 * W is read in its row-major upper triangle (what cspbo_gp_cholesky_inverse_inplace leaves).  desc->symmetric (a == b,
 * row0 == col0) evaluates the upper block triangle and counts the off-diagonal tiles twice.  Per-warp partial sums are
 * added in a fixed order: the result is deterministic. */
int cspbo_block_trace(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b, const double *alpha,
                      const double *W, long long ldw, long long row0, long long col0, double *result, int accumulate);
/* One rank's share of the same trace when K is distributed (csp_bo_b200/distributed.py: lml_sharded): the 8-group blocks
 * [a_blk0, a_blk1) of pack a against ALL of pack b, rows and columns in PACKED order (matrix row of packed position p is
 * row0 + p*nc1, column col0 + q*nc2: the panel layout of cspbo_block_build_panel), Wslab = the rows [slab_row0, ...) of
 * K^-1 (row-major, leading dimension ldw; Wslab[(r - slab_row0)*ldw + c] must be valid for every column c the slab's
 * tiles touch), alpha in the same packed order.  When a == b only the blocks b >= a are evaluated and the off-diagonal
 * tiles count twice (both factors are symmetric): columns left of the slab's first row are then never read. */
int cspbo_block_trace_slab(const cspbo_block_desc *desc, const cspbo_pack *a, const cspbo_pack *b, const double *alpha,
                           const double *Wslab, long long ldw, long long row0, long long col0, long long slab_row0,
                           int a_blk0, int a_blk1, double *result, int accumulate);
/* Building blocks of the sharded Cholesky K = U^T U (row-major; rows of U live where the rows of K live;
 * csp_bo_b200/distributed.py: ShardedCholesky).  All asynchronous on `stream` (a cudaStream_t); matrices are row-major
 * views given by pointer + leading dimension.  Distributed restatement of the single scipy `cholesky` of
 * gaussianprocess.py:116.
 *   chol_diag      D[h,h] (symmetric diagonal block) -> its upper factor in place (cuSOLVER potrf; *info_dev > 0 <=> not
 *                  positive definite: device memory, never read by the call) and inv[h*h] = the factor's inverse
 *   chol_apply_inv dst[h,w] = U_ss^-T src[h,w]: the panel rows right of the diagonal block (cuBLAS trmm with the inverse)
 *   chol_update    C[nb,w] -= A[h,nb]^T B[h,w]: trailing update of nb matrix rows by an h-row panel (cuBLAS dgemm) */
int cspbo_gp_chol_diag(double *D, long long ld, int h, double *inv, int *info_dev, void *stream);
int cspbo_gp_chol_apply_inv(const double *inv, int h, const double *src, long long ldb, long long w, double *dst,
                            long long ldc, void *stream);
int cspbo_gp_chol_update(double *C, long long ldc, int nb, long long w, const double *A, long long lda, const double *B,
                         long long ldb, int h, void *stream);
/* Four streams for the factorisation with the SMs partitioned by CUDA green contexts: `diag` owns diag_sms SMs
 * (the latency-bound potrf chain no longer shares FP64 pipes with the trailing dgemms), `panel` (high priority),
 * `look` (look-ahead updates, medium) and `update` (bulk updates, low priority) share the rest.  Created once per
 * device; fails (callers fall back to plain priority streams) when the driver has no green contexts. */
int cspbo_gp_chol_streams(int diag_sms, void **diag, void **panel, void **look, void **update, int *diag_sms_got,
                          int *rest_sms_got);
/* limit the cuBLAS kernels issued on `stream` by these operations to sm_count SMs (0 = all; cublasSetSmCountTarget) */
int cspbo_gp_stream_sm_target(void *stream, int sm_count);
/* pred[n] = Kstar[n, N] . alpha[N]    gaussianprocess.py:140,359 */
int cspbo_gp_predict(const double *Kstar, int n, int N, const double *alpha, double *pred);

/* ------------------------------------------------------------------------------------------
 * (4) SO(3) power-spectrum descriptor on the device: produces the inputs of layers (1)-(3)
 *     (reference cspbo/SO3.py: calculate :187-291, neighbour list :317-376, compute_dcs :581-695).
 * ------------------------------------------------------------------------------------------ */
typedef struct cspbo_so3 cspbo_so3;
/* cutoff: 0 cosine, 1 tanh, 2 poly1, 3 poly2, 4 poly3, 5 poly4, 6 unity (SO3.py:378-461);
 * weight_on: neighbours of another species count negative (SO3.py:350-354). nmax <= 6, lmax <= 6. */
int cspbo_so3_create(int nmax, int lmax, double rcut, double alpha, int cutoff, int weight_on, cspbo_so3 **out);
int cspbo_so3_destroy(cspbo_so3 *h);
/* One structure: positions[n,3], atomic numbers[n], cell[3,3] (rows = lattice vectors), pbc[3] (HOST).
 * Returns the number of (centre, neighbour) slots of `seq` and the descriptor length. */
int cspbo_so3_compute(cspbo_so3 *h, int n_atoms, const double *positions, const int *numbers, const double *cell,
                      const int *pbc, int stress, int *n_seq, int *n_coefs);
/* Copy the results of the last compute to HOST buffers: x[n,nc], dxdr[n_seq,nc,3], rdxdr[n_seq,nc,3,3]
 * (NULL to skip; needs stress=1), seq[n_seq,2] (int64). */
/* The results of the last cspbo_so3_compute where they are: DEVICE pointers to x [n_atoms, n_coefs],
 * dxdr [n_seq, n_coefs, 3] and rdxdr [n_seq, n_coefs, 3, 3] (null unless computed with stress), owned by the handle
 * and valid until its next compute / destroy.  Lets the MD path (gaussianprocess.py:335-351) pack K* without a round
 * trip of the derivatives through the host. */
int cspbo_so3_device_ptrs(const cspbo_so3 *h, double **x, double **dxdr, double **rdxdr);
int cspbo_so3_fetch(const cspbo_so3 *h, double *x, double *dxdr, double *rdxdr, long long *seq);

/* LJ base potential (reference cspbo/calculator.py:68-177): per-atom energies [n], forces [n,3] and 3x3 stress sums
 * [n,9] (0.5 * sum_j f_ij (x) d_ij, not yet divided by the volume) of the shifted 12-6 potential with cutoff rc; host
 * pointers in and out, neighbour search + one kernel inside. */
int cspbo_lj_compute(int n_atoms, const double *pos, const double *cell, const int *pbc, double rc, double sigma,
                     double epsilon, double *energies, double *forces, double *stresses);

#ifdef __cplusplus
}
#endif
#endif /* CSPBO_B200_H */
