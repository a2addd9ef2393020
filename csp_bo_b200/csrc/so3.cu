// so3.cu -- SO(3) power-spectrum descriptor and its Cartesian derivatives on the GPU: the producer of the
// hot path's inputs {x, dxdr, rdxdr, seq} (reference cspbo/SO3.py:187-291 calculate, :317-376 neighbour
// list, :463-505 radial basis + quadrature, :581-695 compute_dcs).  SURVEY.md 8(f) row 2.
//
//   host   neighbour pairs (brute force over periodic images, same pair set as ase's
//          NeighborList(cutoffs=rcut/2, bothways=True, self_interaction=False)), `seq` slots, the
//          orthonormalised radial basis W (SO3.py:463-479) and the Gauss-Chebyshev tables
//   K1     one warp per (centre, neighbour) pair: radial integrals I_nl(R) = sum_q G[n,q] i_l(2 a R r_q) and
//          dI/dR, spherical harmonics Y_lm (l <= lmax+1) and their gradients, then the expansion
//          coefficients c_nlm and dc_nlm/dr
//   K2     one block per centre: c_tot = sum of its pairs; power spectrum x = sum_m c_tot conj(c_tot)
//   K3     one warp per seq slot (i,j): dP of every pair (periodic images) of the slot, summed in a
//          fixed order -> dxdr[slot], stress terms -r_j (x) dP
//   K4     one block per centre: the (i,i) slot, dxdr[ii] = -sum_{j != i} dxdr[ij]; stress += r_i (x) sum dP
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.h"

namespace cspbo {

constexpr int SO3_MAXN = 6;    // nmax
constexpr int SO3_MAXL = 7;    // lmax + 1 (harmonics are needed one order higher for the gradients)
constexpr int SO3_MAXQ = 160;  // quadrature nodes (nmax+lmax+1)*10

struct So3Params {
    int nmax, lmax, nq, cutoff, nlm;   // nlm = (lmax+1)^2 valid (l,m) per radial index
    double rcut, alpha;
};

__constant__ double c_rq[SO3_MAXQ];
__constant__ double c_G[SO3_MAXN * SO3_MAXQ];
__constant__ double c_norm[SO3_MAXL + 1];

__device__ __forceinline__ int lm_index(int l, int m) { return l * l + l + m; }   // (l,m) -> 0..(L+1)^2-1

// modified spherical Bessel functions of the first kind i_0..i_lmax(x)
__device__ void sph_in(double x, int lmax, double *il)
{
    if (x < 4.0) {
        // power series: i_l = x^l/(2l+1)!! * sum_k (x^2/2)^k / (k! (2l+3)(2l+5)...(2l+2k+1))
        const double t = 0.5 * x * x;
        double pref = 1.0;
        for (int l = 0; l <= lmax; l++) {
            if (l > 0) pref *= x / (2.0 * l + 1.0);
            double term = 1.0, sum = 1.0;
            for (int k = 1; k < 40; k++) {
                term *= t / (k * (2.0 * l + 2.0 * k + 1.0));
                sum += term;
                if (term < 1e-18 * sum) break;
            }
            il[l] = pref * sum;
        }
    } else {
        const double ex = exp(x), emx = 1.0 / ex;
        const double sh = 0.5 * (ex - emx), ch = 0.5 * (ex + emx);
        il[0] = sh / x;
        if (lmax >= 1) il[1] = (x * ch - sh) / (x * x);
        for (int l = 1; l < lmax; l++) il[l + 1] = il[l - 1] - (2.0 * l + 1.0) / x * il[l];
    }
}

__device__ __forceinline__ double cutoff_fn(int kind, double R, double rc, bool deriv)
{
    const double x = R / rc;
    switch (kind) {
        case 0:  // cosine (SO3.py:378-385)
            return deriv ? -0.5 * M_PI / rc * sin(M_PI * x) : 0.5 * (cos(M_PI * x) + 1.0);
        case 1: {  // tanh (:387-395)
            const double th = tanh(1.0 - x), t2 = th * th;
            return deriv ? -(3.0 / rc) * t2 * (1.0 - t2) : t2 * th;
        }
        case 2:  // poly1 (:397-408)
            return deriv ? (6.0 / (rc * rc)) * R * (x - 1.0) : x * x * (2.0 * x - 3.0) + 1.0;
        case 3:  // poly2 (:410-419)
            return deriv ? (-30.0 / rc) * (x * x * (x - 1.0) * (x - 1.0)) : x * x * x * (x * (15.0 - 6.0 * x) - 10.0) + 1.0;
        case 4: {  // poly3 (:421-430)
            const double a = x * (x - 1.0);
            return deriv ? (140.0 / rc) * (a * a * a) : x * x * x * x * (x * (x * (20.0 * x - 70.0) + 84.0) - 35.0) + 1.0;
        }
        case 5: {  // poly4 (:432-441)
            const double a = x * (x - 1.0);
            return deriv ? (-630.0 / rc) * (a * a * a * a)
                         : x * x * x * x * x * (x * (x * (x * (315.0 - 70.0 * x) - 540.0) + 420.0) - 126.0) + 1.0;
        }
        default:  // unity (:457-461): value 1, "derivative" 1 as in the reference
            return 1.0;
    }
}

struct cplx { double re, im; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }

// K1 ------------------------------------------------------------------------------------------
// per-warp shared scratch: I[n][l], dI[n][l] (radial), Y[(L+2)^2], dY[(L+1)^2][3]
__global__ void __launch_bounds__(128) so3_pair_kernel(So3Params P, int n_pairs, const double *__restrict__ rel,
                                                       const double *__restrict__ wts, double2 *__restrict__ C,
                                                       double2 *__restrict__ dC)
{
    __shared__ double sI[4][SO3_MAXN][SO3_MAXL], sdI[4][SO3_MAXN][SO3_MAXL];
    __shared__ cplx sY[4][(SO3_MAXL + 1) * (SO3_MAXL + 1)];
    __shared__ cplx sdY[4][SO3_MAXL * SO3_MAXL][3];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int pair = blockIdx.x * 4 + w;
    if (pair >= n_pairs) return;
    const double rx = rel[3 * pair], ry = rel[3 * pair + 1], rz = rel[3 * pair + 2];
    const double R = sqrt(rx * rx + ry * ry + rz * rz);
    const double ux = rx / R, uy = ry / R, uz = rz / R;
    const int L = P.lmax + 1;

    // radial integrals: lane -> (n, l)
    for (int e = lane; e < P.nmax * L; e += 32) {
        const int n = e / L, l = e % L;
        double acc = 0.0, dacc = 0.0;
        double il[SO3_MAXL + 2];
        for (int q = 0; q < P.nq; q++) {
            const double x = 2.0 * P.alpha * R * c_rq[q];
            sph_in(x, l + 1, il);
            // i_l'(x) = i_{l+1}(x) + (l/x) i_l(x)
            const double dil = il[l + 1] + (l > 0 ? (double)l / x * il[l] : 0.0);
            const double g = c_G[n * SO3_MAXQ + q];
            acc += g * il[l];
            dacc += g * dil * (2.0 * P.alpha * c_rq[q]);
        }
        sI[w][n][l] = acc;
        sdI[w][n][l] = dacc;
    }
    // spherical harmonics Y_l^m, l <= lmax+1: lane -> (l, m >= 0)
    {
        const int LY = P.lmax + 1;   // highest order needed
        const int ny = (LY + 1) * (LY + 2) / 2;
        for (int e = lane; e < ny; e += 32) {
            int l = 0, m = e;
            while (m > l) { m -= l + 1; l++; }   // e -> (l, m) with 0 <= m <= l
            // Q_l^m = P_l^m / sin^m(theta) by upward recurrence in l from Q_m^m = (-1)^m (2m-1)!!
            double qmm = 1.0;
            for (int k = 1; k <= m; k++) qmm *= -(2.0 * k - 1.0);
            double q0 = qmm, q1 = uz * (2.0 * m + 1.0) * qmm, ql = (l == m) ? q0 : q1;
            for (int ll = m + 2; ll <= l; ll++) {
                ql = ((2.0 * ll - 1.0) * uz * q1 - (ll + m - 1.0) * q0) / (ll - m);
                q0 = q1; q1 = ql;
            }
            // normalisation sqrt((2l+1)/(4 pi) (l-m)!/(l+m)!)
            double ratio = 1.0;
            for (int k = l - m + 1; k <= l + m; k++) ratio /= k;
            const double nrm = sqrt((2.0 * l + 1.0) / (4.0 * M_PI) * ratio);
            // (sin(theta) e^{i phi})^m = ((x + i y)/R)^m
            cplx z = {1.0, 0.0};
            const cplx b = {ux, uy};
            for (int k = 0; k < m; k++) z = cmul(z, b);
            const cplx y = {nrm * ql * z.re, nrm * ql * z.im};
            sY[w][lm_index(l, m)] = y;
            if (m > 0) {
                const double sg = (m & 1) ? -1.0 : 1.0;
                sY[w][lm_index(l, -m)] = {sg * y.re, -sg * y.im};
            }
        }
    }
    __syncwarp();
    // gradients of Y_l^m for l <= lmax (SO3.py:645-672): spherical covariant components from Y_{l+-1}
    for (int e = lane; e < L * L; e += 32) {
        int l = (int)floor(sqrt((double)e));
        while (l * l > e) l--;
        while ((l + 1) * (l + 1) <= e) l++;
        const int m = e - l * l - l;
        cplx g0 = {0, 0}, gp = {0, 0}, gm = {0, 0};
        if (l >= 1) {
            const double invR = 1.0 / R;
            auto Yv = [&](int ll, int mm) -> cplx {
                if (mm < -ll || mm > ll) return cplx{0.0, 0.0};
                return sY[w][lm_index(ll, mm)];
            };
            double c;
            cplx t;
            c = -sqrt(((l + 1.0) * (l + 1.0) - m * m) / (2.0 * l + 1.0) / (2.0 * l + 3.0)) * l * invR;
            t = Yv(l + 1, m); g0 = {c * t.re, c * t.im};
            if (abs(m) <= l - 1) {
                c = sqrt(((double)l * l - m * m) / (2.0 * l - 1.0) / (2.0 * l + 1.0)) * (l + 1.0) * invR;
                t = Yv(l - 1, m); g0.re += c * t.re; g0.im += c * t.im;
            }
            c = -sqrt((l + m + 1.0) * (l + m + 2.0) / 2.0 / (2.0 * l + 1.0) / (2.0 * l + 3.0)) * l * invR;
            t = Yv(l + 1, m + 1); gp = {c * t.re, c * t.im};
            if (abs(m + 1) <= l - 1) {
                c = sqrt((l - m - 1.0) * (l - m) / 2.0 / (2.0 * l - 1.0) / (2.0 * l + 1.0)) * (l + 1.0) * invR;
                t = Yv(l - 1, m + 1); gp.re -= c * t.re; gp.im -= c * t.im;
            }
            c = -sqrt((l - m + 1.0) * (l - m + 2.0) / 2.0 / (2.0 * l + 1.0) / (2.0 * l + 3.0)) * l * invR;
            t = Yv(l + 1, m - 1); gm = {c * t.re, c * t.im};
            if (abs(m - 1) <= l - 1) {
                c = sqrt((l + m - 1.0) * (l + m) / 2.0 / (2.0 * l - 1.0) / (2.0 * l + 1.0)) * (l + 1.0) * invR;
                t = Yv(l - 1, m - 1); gm.re -= c * t.re; gm.im -= c * t.im;
            }
        }
        const double s = 0.70710678118654752440;
        sdY[w][e][0] = {s * (gm.re - gp.re), s * (gm.im - gp.im)};          // 1/sqrt2 (xm - xp)
        sdY[w][e][1] = {-s * (gm.im + gp.im), s * (gm.re + gp.re)};         // i/sqrt2 (xm + xp)
        sdY[w][e][2] = g0;
    }
    __syncwarp();
    // coefficients: lane -> (n, l, m)
    const double ex = 4.0 * M_PI * exp(-P.alpha * R * R);
    const double dexR = -2.0 * P.alpha * R * ex;
    const double fc = cutoff_fn(P.cutoff, R, P.rcut, false), dfcR = cutoff_fn(P.cutoff, R, P.rcut, true);
    const double wt = wts[pair];
    const double u[3] = {ux, uy, uz};
    const int nlm = P.nlm;
    for (int e = lane; e < P.nmax * nlm; e += 32) {
        const int n = e / nlm, lm = e % nlm;
        int l = (int)floor(sqrt((double)lm));
        while (l * l > lm) l--;
        while ((l + 1) * (l + 1) <= lm) l++;
        const double I = sI[w][n][l], dI = sdI[w][n][l];
        const cplx Y = sY[w][lm];
        const double sc = wt * c_norm[l];
        const cplx YI = {Y.re * I, Y.im * I};
        const cplx c0 = {ex * YI.re, ex * YI.im};                        // before the cutoff
        C[(size_t)pair * P.nmax * nlm + e] = make_double2(sc * fc * c0.re, sc * fc * c0.im);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            const cplx dY = sdY[w][lm][k];
            // d/dr_k [ ex * I * Y ] = dex u_k I Y + ex (dY_k I + Y dI u_k)
            double gr = dexR * u[k] * YI.re + ex * (dY.re * I + Y.re * dI * u[k]);
            double gi = dexR * u[k] * YI.im + ex * (dY.im * I + Y.im * dI * u[k]);
            gr = gr * fc + dfcR * u[k] * c0.re;
            gi = gi * fc + dfcR * u[k] * c0.im;
            dC[((size_t)pair * P.nmax * nlm + e) * 3 + k] = make_double2(sc * gr, sc * gi);
        }
    }
}

// K2 ------------------------------------------------------------------------------------------
__global__ void so3_center_kernel(So3Params P, const int *__restrict__ pair_start, const double2 *__restrict__ C,
                                  double2 *__restrict__ ctot, double *__restrict__ x, int ncoefs)
{
    extern __shared__ double2 sct[];
    const int i = blockIdx.x;
    const int ne = P.nmax * P.nlm;
    for (int e = threadIdx.x; e < ne; e += blockDim.x) {
        double re = 0.0, im = 0.0;
        for (int p = pair_start[i]; p < pair_start[i + 1]; p++) {
            const double2 c = C[(size_t)p * ne + e];
            re += c.x; im += c.y;
        }
        sct[e] = make_double2(re, im);
        ctot[(size_t)i * ne + e] = make_double2(re, im);
    }
    __syncthreads();
    const int L = P.lmax + 1;
    // x[i, tril(n,n')*L + l] = sum_m Re(ctot[n,l,m] conj(ctot[n',l,m]))   (SO3.py:231,265)
    for (int c = threadIdx.x; c < ncoefs; c += blockDim.x) {
        const int t = c / L, l = c % L;
        int n = 0, rem = t;
        while (rem > n) { rem -= n + 1; n++; }
        const int n2 = rem;
        double acc = 0.0;
        for (int m = -l; m <= l; m++) {
            const double2 a = sct[n * P.nlm + lm_index(l, m)], b = sct[n2 * P.nlm + lm_index(l, m)];
            acc += a.x * b.x + a.y * b.y;
        }
        x[(size_t)i * ncoefs + c] = acc;
    }
}

// K3 ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) so3_slot_kernel(So3Params P, int n_slots, const int *__restrict__ slot_start,
                                                       const int *__restrict__ slot_pairs, const int *__restrict__ slot_center,
                                                       const double *__restrict__ Rj, const double2 *__restrict__ dC,
                                                       const double2 *__restrict__ ctot, double *__restrict__ dxdr,
                                                       double *__restrict__ rdxdr, double stress_scale, int ncoefs)
{
    const int lane = threadIdx.x & 31;
    const int slot = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slot >= n_slots) return;
    const int i = slot_center[slot];
    const int ne = P.nmax * P.nlm, L = P.lmax + 1;
    const double2 *ct = ctot + (size_t)i * ne;
    for (int e = lane; e < ncoefs * 3; e += 32) {
        const int c = e / 3, k = e % 3;
        const int t = c / L, l = c % L;
        int n = 0, rem = t;
        while (rem > n) { rem -= n + 1; n++; }
        const int n2 = rem;
        double sum = 0.0, st[3] = {0.0, 0.0, 0.0};
        for (int q = slot_start[slot]; q < slot_start[slot + 1]; q++) {
            const int p = slot_pairs[q];
            const double2 *d = dC + (size_t)p * ne * 3;
            double acc = 0.0;
            for (int m = -l; m <= l; m++) {
                const int lm = lm_index(l, m);
                const double2 a = d[(n * P.nlm + lm) * 3 + k], b = ct[n2 * P.nlm + lm];
                const double2 a2 = d[(n2 * P.nlm + lm) * 3 + k], b2 = ct[n * P.nlm + lm];
                acc += a.x * b.x + a.y * b.y + a2.x * b2.x + a2.y * b2.y;   // Re(dc conj(ct)) + transpose
            }
            sum += acc;
            if (rdxdr) {
#pragma unroll
                for (int a = 0; a < 3; a++) st[a] -= Rj[3 * p + a] * acc;   // pstress -= r_j (x) dP  (SO3.py:257)
            }
        }
        dxdr[((size_t)slot * ncoefs + c) * 3 + k] = sum;
        if (rdxdr) {
#pragma unroll
            for (int a = 0; a < 3; a++) rdxdr[(((size_t)slot * ncoefs + c) * 3 + a) * 3 + k] = stress_scale * st[a];
        }
    }
}

// K4 ------------------------------------------------------------------------------------------
__global__ void so3_self_kernel(int n_atoms, const int *__restrict__ seq_start, const int *__restrict__ self_slot,
                                const double *__restrict__ pos, double *__restrict__ dxdr, double *__restrict__ rdxdr,
                                double stress_scale, int ncoefs)
{
    const int i = blockIdx.x;
    const int ii = self_slot[i];
    for (int e = threadIdx.x; e < ncoefs * 3; e += blockDim.x) {
        double S = 0.0;
        for (int s = seq_start[i]; s < seq_start[i + 1]; s++) S += dxdr[(size_t)s * ncoefs * 3 + e];
        // dplist[ii] -= sum over all slots of i (SO3.py:270); pstress[ii] += r_i (x) sum_w dP_w (:271-272)
        dxdr[(size_t)ii * ncoefs * 3 + e] -= S;
        if (rdxdr) {
            const int c = e / 3, k = e % 3;
#pragma unroll
            for (int a = 0; a < 3; a++)
                rdxdr[(((size_t)ii * ncoefs + c) * 3 + a) * 3 + k] += stress_scale * pos[3 * i + a] * S;
        }
    }
}

// ---- host: radial basis tables ----------------------------------------------------------------
static void invert(std::vector<double> &a, int n)
{
    std::vector<double> inv((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; c++) {
        int piv = c;
        for (int r = c + 1; r < n; r++) if (fabs(a[(size_t)r * n + c]) > fabs(a[(size_t)piv * n + c])) piv = r;
        for (int k = 0; k < n; k++) { std::swap(a[(size_t)c * n + k], a[(size_t)piv * n + k]); std::swap(inv[(size_t)c * n + k], inv[(size_t)piv * n + k]); }
        const double d = a[(size_t)c * n + c];
        for (int k = 0; k < n; k++) { a[(size_t)c * n + k] /= d; inv[(size_t)c * n + k] /= d; }
        for (int r = 0; r < n; r++) if (r != c) {
            const double f = a[(size_t)r * n + c];
            for (int k = 0; k < n; k++) { a[(size_t)r * n + k] -= f * a[(size_t)c * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k]; }
        }
    }
    a = inv;
}

// symmetric square root by cyclic Jacobi (the reference takes V sqrt(D) V^-1 of eig(sinv), SO3.py:475-478)
static void sym_sqrt(std::vector<double> &a, int n)
{
    std::vector<double> V((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0.0;
        for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) off += a[(size_t)p * n + q] * a[(size_t)p * n + q];
        if (off < 1e-300) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = a[(size_t)p * n + q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (a[(size_t)q * n + q] - a[(size_t)p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = a[(size_t)k * n + p], akq = a[(size_t)k * n + q];
                    a[(size_t)k * n + p] = c * akp - s * akq; a[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = a[(size_t)p * n + k], aqk = a[(size_t)q * n + k];
                    a[(size_t)p * n + k] = c * apk - s * aqk; a[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
                    V[(size_t)k * n + p] = c * vkp - s * vkq; V[(size_t)k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    std::vector<double> out((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        double acc = 0.0;
        for (int k = 0; k < n; k++) acc += V[(size_t)i * n + k] * sqrt(std::max(a[(size_t)k * n + k], 0.0)) * V[(size_t)j * n + k];
        out[(size_t)i * n + j] = acc;
    }
    a = out;
}

}  // namespace cspbo

using namespace cspbo;

struct cspbo_so3 {
    So3Params P;
    int weight_on = 0;
    std::vector<double> h_rq, h_G, h_norm;
    // results of the last compute (device) + host seq
    int n_atoms = 0, n_pairs = 0, n_seq = 0, ncoefs = 0, stress = 0;
    std::vector<long long> h_seq;
    double *d_x = nullptr, *d_dxdr = nullptr, *d_rdxdr = nullptr;
};

// All (i, j, image) pairs with 0 < |r_j + o.cell - r_i| < rcut in (i, j, o) order (SO3.py:317-359; also the neighbour
// list of the LJ base potential, calculator.py:131-133).
struct Pair { int i, j; int o[3]; double rel[3]; };

static int neighbour_pairs(int n_atoms, const double *pos, const double *cell, const int *pbc, double rcut, std::vector<Pair> &pairs)
{
    // ---- neighbour pairs (SO3.py:317-359): |r_j + offset.cell - r_i| < rcut, zero-offset self excluded
    double inv[9];
    {
        const double *c = cell;
        const double det = c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) + c[2] * (c[3] * c[7] - c[4] * c[6]);
        if (fabs(det) < 1e-300) { set_error("neighbour search: singular cell"); return CSPBO_E_ARG; }
        inv[0] = (c[4] * c[8] - c[5] * c[7]) / det; inv[1] = (c[2] * c[7] - c[1] * c[8]) / det; inv[2] = (c[1] * c[5] - c[2] * c[4]) / det;
        inv[3] = (c[5] * c[6] - c[3] * c[8]) / det; inv[4] = (c[0] * c[8] - c[2] * c[6]) / det; inv[5] = (c[2] * c[3] - c[0] * c[5]) / det;
        inv[6] = (c[3] * c[7] - c[4] * c[6]) / det; inv[7] = (c[1] * c[6] - c[0] * c[7]) / det; inv[8] = (c[0] * c[4] - c[1] * c[3]) / det;
    }
    // fractional coordinates and the cell heights along the reciprocal directions
    double hgt[3];
    for (int k = 0; k < 3; k++) hgt[k] = 1.0 / sqrt(inv[k] * inv[k] + inv[3 + k] * inv[3 + k] + inv[6 + k] * inv[6 + k]);
    std::vector<double> frac((size_t)std::max(n_atoms, 1) * 3);
    for (int a = 0; a < n_atoms; a++)
        for (int k = 0; k < 3; k++)
            frac[(size_t)a * 3 + k] = pos[3 * a] * inv[k] + pos[3 * a + 1] * inv[3 + k] + pos[3 * a + 2] * inv[6 + k];
    pairs.clear();
    const double rc2 = rcut * rcut;
    // Linked cells in fractional coordinates (a 192-atom structure: 0.53 ms for the all-pairs x images scan, 0.05 ms
    // with cells).  Along a periodic axis the atoms are wrapped into [0,1) (wrap shift w) and binned into nc cells of
    // thickness >= rcut where the cell is thick enough, else one cell searched over several images; a non-periodic axis
    // is one cell without images.  Candidates of atom i are then put in (j, offset) order, which keeps the pair list --
    // and with it the summation order of the kernels -- exactly what the plain double loop over (i, j, offset) gives.
    int nc[3], reach[3];
    for (int k = 0; k < 3; k++) {
        if (!pbc[k]) { nc[k] = 1; reach[k] = 0; continue; }
        nc[k] = std::max(1, std::min(64, (int)floor(hgt[k] / rcut)));
        reach[k] = (int)ceil(rcut / (hgt[k] / nc[k]) - 1e-12);
        if (reach[k] < 1) reach[k] = 1;
    }
    std::vector<int> wrap((size_t)std::max(n_atoms, 1) * 3, 0), cidx((size_t)std::max(n_atoms, 1) * 3, 0);
    std::vector<int> cell_head((size_t)nc[0] * nc[1] * nc[2] + 1, 0), cell_atoms((size_t)std::max(n_atoms, 1));
    for (int a = 0; a < n_atoms; a++) {
        for (int k = 0; k < 3; k++) {
            if (!pbc[k]) continue;
            const double f = frac[(size_t)a * 3 + k];
            const int w = (int)floor(f);
            wrap[(size_t)a * 3 + k] = w;
            int c = (int)((f - w) * nc[k]);
            cidx[(size_t)a * 3 + k] = std::min(std::max(c, 0), nc[k] - 1);
        }
        cell_head[(size_t)((cidx[(size_t)a * 3] * nc[1] + cidx[(size_t)a * 3 + 1]) * nc[2] + cidx[(size_t)a * 3 + 2]) + 1]++;
    }
    for (size_t c = 1; c < cell_head.size(); c++) cell_head[c] += cell_head[c - 1];
    {
        std::vector<int> fill(cell_head.begin(), cell_head.end() - 1);
        for (int a = 0; a < n_atoms; a++)
            cell_atoms[(size_t)fill[(size_t)((cidx[(size_t)a * 3] * nc[1] + cidx[(size_t)a * 3 + 1]) * nc[2] + cidx[(size_t)a * 3 + 2])]++] = a;
    }
    struct Cand { int j, o[3]; double rel[3]; };
    std::vector<Cand> cand;
    for (int i = 0; i < n_atoms; i++) {
        cand.clear();
        const int *ci = &cidx[(size_t)i * 3], *wi = &wrap[(size_t)i * 3];
        for (int da = -reach[0]; da <= reach[0]; da++)
            for (int db = -reach[1]; db <= reach[1]; db++)
                for (int dc = -reach[2]; dc <= reach[2]; dc++) {
                    const int dd[3] = {da, db, dc};
                    int cc[3], sh[3];
                    for (int k = 0; k < 3; k++) {
                        const int t = ci[k] + dd[k];
                        sh[k] = (t >= 0) ? t / nc[k] : -((-t + nc[k] - 1) / nc[k]);     // floor division
                        cc[k] = t - sh[k] * nc[k];
                    }
                    const size_t cell_id = (size_t)((cc[0] * nc[1] + cc[1]) * nc[2] + cc[2]);
                    for (int q = cell_head[cell_id]; q < cell_head[cell_id + 1]; q++) {
                        const int j = cell_atoms[(size_t)q];
                        const int *wj = &wrap[(size_t)j * 3];
                        // image offset relative to the UNWRAPPED positions: r = pos_j + o.cell - pos_i
                        const int o[3] = {sh[0] - wj[0] + wi[0], sh[1] - wj[1] + wi[1], sh[2] - wj[2] + wi[2]};
                        double r[3];
                        for (int k = 0; k < 3; k++)
                            r[k] = pos[3 * j + k] + o[0] * cell[k] + o[1] * cell[3 + k] + o[2] * cell[6 + k] - pos[3 * i + k];
                        const double d2 = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
                        if (d2 < rc2 && d2 > 1e-24) cand.push_back({j, {o[0], o[1], o[2]}, {r[0], r[1], r[2]}});
                    }
                }
        std::sort(cand.begin(), cand.end(), [](const Cand &x, const Cand &y) {
            if (x.j != y.j) return x.j < y.j;
            if (x.o[0] != y.o[0]) return x.o[0] < y.o[0];
            if (x.o[1] != y.o[1]) return x.o[1] < y.o[1];
            return x.o[2] < y.o[2];
        });
        for (size_t q = 0; q < cand.size(); q++) {
            // when a periodic axis has fewer cells than 2*reach+1 the same (j, offset) is met through several cell shifts
            if (q > 0 && cand[q].j == cand[q - 1].j && cand[q].o[0] == cand[q - 1].o[0] && cand[q].o[1] == cand[q - 1].o[1] &&
                cand[q].o[2] == cand[q - 1].o[2])
                continue;
            const Cand &c = cand[q];
            pairs.push_back({i, c.j, {c.o[0], c.o[1], c.o[2]}, {c.rel[0], c.rel[1], c.rel[2]}});
        }
    }
    return CSPBO_OK;
}

extern "C" int cspbo_so3_create(int nmax, int lmax, double rcut, double alpha, int cutoff, int weight_on, cspbo_so3 **out)
{
    if (!out) { set_error("so3_create: null"); return CSPBO_E_ARG; }
    *out = nullptr;
    if (nmax < 1 || nmax > SO3_MAXN || lmax < 0 || lmax + 1 > SO3_MAXL || !(rcut > 0) || !(alpha > 0) || cutoff < 0 || cutoff > 6) {
        set_error("so3_create: unsupported parameters (nmax<=%d, lmax<=%d)", SO3_MAXN, SO3_MAXL - 1);
        return CSPBO_E_ARG;
    }
    const int nq = (nmax + lmax + 1) * 10;
    if (nq > SO3_MAXQ) { set_error("so3_create: too many quadrature nodes"); return CSPBO_E_ARG; }
    cspbo_so3 *h = new (std::nothrow) cspbo_so3();
    if (!h) { set_error("so3_create: allocation failed"); return CSPBO_E_NOMEM; }
    h->P = {nmax, lmax, nq, cutoff, (lmax + 1) * (lmax + 1), rcut, alpha};
    h->weight_on = weight_on;
    // W (SO3.py:463-479)
    std::vector<double> w((size_t)nmax * nmax);
    for (int a = 1; a <= nmax; a++) {
        const double t1 = (2.0 * a + 5) * (2.0 * a + 6) * (2.0 * a + 7);
        for (int b = 1; b <= a; b++) {
            const double t2 = (2.0 * b + 5) * (2.0 * b + 6) * (2.0 * b + 7);
            w[(size_t)(a - 1) * nmax + b - 1] = w[(size_t)(b - 1) * nmax + a - 1] =
                sqrt(t1 * t2) / (5.0 + a + b) / (6.0 + a + b) / (7.0 + a + b);
        }
    }
    invert(w, nmax);
    sym_sqrt(w, nmax);
    // quadrature (SO3.py:497-505) and G (SO3.py:600-626)
    h->h_rq.assign(SO3_MAXQ, 0.0);
    h->h_G.assign((size_t)SO3_MAXN * SO3_MAXQ, 0.0);
    const double weight = M_PI / nq * rcut / 2.0;
    for (int q = 0; q < nq; q++) {
        const double xq = cos((2.0 * (q + 1) - 1.0) * M_PI / 2.0 / nq);
        const double r = rcut / 2.0 * (xq + 1.0);
        h->h_rq[(size_t)q] = r;
        for (int n = 1; n <= nmax; n++) {
            double g = 0.0;
            for (int a = 1; a <= nmax; a++) {
                const double phi = pow(rcut - r, a + 2.0) / sqrt(2.0 * pow(rcut, 2.0 * a + 7.0) / (2.0 * a + 5) / (2.0 * a + 6) / (2.0 * a + 7));
                g += w[(size_t)(n - 1) * nmax + a - 1] * phi;
            }
            h->h_G[(size_t)(n - 1) * SO3_MAXQ + q] = g * weight * r * r * exp(-alpha * r * r) * sqrt(1.0 - xq * xq);
        }
    }
    h->h_norm.assign(SO3_MAXL + 1, 0.0);
    for (int l = 0; l <= lmax; l++) h->h_norm[(size_t)l] = sqrt(2.0 * sqrt(2.0) * M_PI / sqrt(2.0 * l + 1.0));
    *out = h;
    return CSPBO_OK;
}

extern "C" int cspbo_so3_destroy(cspbo_so3 *h)
{
    if (!h) return CSPBO_OK;
    dev_free_cached(h->d_x); dev_free_cached(h->d_dxdr); dev_free_cached(h->d_rdxdr);
    delete h;
    return CSPBO_OK;
}

extern "C" int cspbo_so3_compute(cspbo_so3 *h, int n_atoms, const double *pos, const int *numbers, const double *cell,
                                 const int *pbc, int stress, int *n_seq_out, int *n_coefs_out)
{
    if (!h || !pos || !numbers || !cell || !pbc || n_atoms < 0) { set_error("so3_compute: bad argument"); return CSPBO_E_ARG; }
    const So3Params P = h->P;
    const int L = P.lmax + 1, ncoefs = P.nmax * (P.nmax + 1) / 2 * L;
    static const bool trace = getenv("CSPBO_SO3_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    std::vector<Pair> pairs;
    {
        const int rcn = neighbour_pairs(n_atoms, pos, cell, pbc, P.rcut, pairs);
        if (rcn) return rcn;
    }
    // loop order above is already (i, j, offset): the deterministic summation order of K2 / K3
    const int np = (int)pairs.size();
    const double t_pairs = now();
    // ---- seq slots (SO3.py:361-372): for every centre its neighbour ids plus itself, sorted
    std::vector<int> pair_start((size_t)n_atoms + 1, 0), seq_start((size_t)n_atoms + 1, 0), self_slot((size_t)std::max(n_atoms, 1), 0);
    std::vector<int> slot_start, slot_pairs, slot_center, pair_slot((size_t)std::max(np, 1));
    h->h_seq.clear();
    {
        int p = 0;
        for (int i = 0; i < n_atoms; i++) {
            pair_start[(size_t)i] = p;
            const int p0 = p;
            while (p < np && pairs[(size_t)p].i == i) p++;
            // distinct neighbour ids (pairs are sorted by j inside i) merged with i itself, ascending
            std::vector<int> js;
            for (int q = p0; q < p; q++)
                if (js.empty() || js.back() != pairs[(size_t)q].j) js.push_back(pairs[(size_t)q].j);
            if (!std::binary_search(js.begin(), js.end(), i)) js.insert(std::lower_bound(js.begin(), js.end(), i), i);
            seq_start[(size_t)i] = (int)slot_center.size();
            int q = p0;
            for (int j : js) {
                const int slot = (int)slot_center.size();
                slot_center.push_back(i);
                slot_start.push_back((int)slot_pairs.size());
                h->h_seq.push_back(i); h->h_seq.push_back(j);
                if (j == i) self_slot[(size_t)i] = slot;
                while (q < p && pairs[(size_t)q].j == j) { slot_pairs.push_back(q); pair_slot[(size_t)q] = slot; q++; }
            }
        }
        pair_start[(size_t)n_atoms] = np;
        seq_start[(size_t)n_atoms] = (int)slot_center.size();
        slot_start.push_back((int)slot_pairs.size());
    }
    const int nseq = (int)slot_center.size();
    h->n_atoms = n_atoms; h->n_pairs = np; h->n_seq = nseq; h->ncoefs = ncoefs; h->stress = stress;
    if (n_seq_out) *n_seq_out = nseq;
    if (n_coefs_out) *n_coefs_out = ncoefs;

    const double t_slots = now();
    // ---- device buffers ------------------------------------------------------------------
    dev_free_cached(h->d_x); dev_free_cached(h->d_dxdr); dev_free_cached(h->d_rdxdr);
    h->d_x = h->d_dxdr = h->d_rdxdr = nullptr;
    if (n_atoms == 0) return CSPBO_OK;
    int rc;
    const size_t ne = (size_t)P.nmax * P.nlm;
    if ((rc = dev_alloc_cached((size_t)n_atoms * ncoefs * 8, (void **)&h->d_x))) return rc;
    if ((rc = dev_alloc_cached((size_t)nseq * ncoefs * 3 * 8, (void **)&h->d_dxdr))) return rc;
    if (stress && (rc = dev_alloc_cached((size_t)nseq * ncoefs * 9 * 8, (void **)&h->d_rdxdr))) return rc;
    double *d_rel = nullptr, *d_wt = nullptr, *d_Rj = nullptr, *d_pos = nullptr;
    double2 *d_C = nullptr, *d_dC = nullptr, *d_ct = nullptr;
    int *d_ps = nullptr, *d_ss = nullptr, *d_sp = nullptr, *d_sc = nullptr, *d_seqs = nullptr, *d_self = nullptr;
    std::vector<void *> tmp;
    auto A = [&](size_t bytes, void **p) { int r = dev_alloc_cached(bytes, p); if (!r) tmp.push_back(*p); return r; };
    auto cleanup = [&]() { for (void *p : tmp) dev_free_cached(p); };
#define SO3_TRY(expr) do { int r__ = (expr); if (r__) { cleanup(); return r__; } } while (0)
#define SO3_CUDA(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); cleanup(); return CSPBO_E_CUDA; } } while (0)
    const size_t npz = (size_t)std::max(np, 1);
    SO3_TRY(A(npz * 3 * 8, (void **)&d_rel)); SO3_TRY(A(npz * 8, (void **)&d_wt)); SO3_TRY(A(npz * 3 * 8, (void **)&d_Rj));
    SO3_TRY(A((size_t)n_atoms * 3 * 8, (void **)&d_pos));
    SO3_TRY(A(npz * ne * 16, (void **)&d_C)); SO3_TRY(A(npz * ne * 3 * 16, (void **)&d_dC)); SO3_TRY(A((size_t)n_atoms * ne * 16, (void **)&d_ct));
    SO3_TRY(A(((size_t)n_atoms + 1) * 4, (void **)&d_ps)); SO3_TRY(A(((size_t)nseq + 1) * 4, (void **)&d_ss));
    SO3_TRY(A(npz * 4, (void **)&d_sp)); SO3_TRY(A((size_t)nseq * 4, (void **)&d_sc));
    SO3_TRY(A(((size_t)n_atoms + 1) * 4, (void **)&d_seqs)); SO3_TRY(A((size_t)n_atoms * 4, (void **)&d_self));
    std::vector<double> rel(npz * 3), wt(npz), Rj(npz * 3);
    for (int p = 0; p < np; p++) {
        const Pair &pr = pairs[(size_t)p];
        for (int k = 0; k < 3; k++) { rel[(size_t)p * 3 + k] = pr.rel[k]; Rj[(size_t)p * 3 + k] = pr.rel[k] + pos[3 * pr.i + k]; }
        double f = 1.0;
        if (h->weight_on && numbers[pr.j] != numbers[pr.i]) f = -1.0;     // SO3.py:350-354
        wt[(size_t)p] = f * numbers[pr.j];
    }
    SO3_CUDA(cudaMemcpyToSymbol(c_rq, h->h_rq.data(), SO3_MAXQ * 8));
    SO3_CUDA(cudaMemcpyToSymbol(c_G, h->h_G.data(), (size_t)SO3_MAXN * SO3_MAXQ * 8));
    SO3_CUDA(cudaMemcpyToSymbol(c_norm, h->h_norm.data(), (SO3_MAXL + 1) * 8));
    SO3_CUDA(cudaMemcpyAsync(d_rel, rel.data(), npz * 3 * 8, cudaMemcpyHostToDevice));
    SO3_CUDA(cudaMemcpyAsync(d_wt, wt.data(), npz * 8, cudaMemcpyHostToDevice));
    SO3_CUDA(cudaMemcpyAsync(d_Rj, Rj.data(), npz * 3 * 8, cudaMemcpyHostToDevice));
    SO3_CUDA(cudaMemcpyAsync(d_pos, pos, (size_t)n_atoms * 3 * 8, cudaMemcpyHostToDevice));
    SO3_CUDA(cudaMemcpyAsync(d_ps, pair_start.data(), ((size_t)n_atoms + 1) * 4, cudaMemcpyHostToDevice));
    SO3_CUDA(cudaMemcpyAsync(d_ss, slot_start.data(), ((size_t)nseq + 1) * 4, cudaMemcpyHostToDevice));
    if (np) SO3_CUDA(cudaMemcpyAsync(d_sp, slot_pairs.data(), (size_t)np * 4, cudaMemcpyHostToDevice));
    SO3_CUDA(cudaMemcpyAsync(d_sc, slot_center.data(), (size_t)nseq * 4, cudaMemcpyHostToDevice));
    SO3_CUDA(cudaMemcpyAsync(d_seqs, seq_start.data(), ((size_t)n_atoms + 1) * 4, cudaMemcpyHostToDevice));
    SO3_CUDA(cudaMemcpyAsync(d_self, self_slot.data(), (size_t)n_atoms * 4, cudaMemcpyHostToDevice));
    double vol = fabs(cell[0] * (cell[4] * cell[8] - cell[5] * cell[7]) - cell[1] * (cell[3] * cell[8] - cell[5] * cell[6]) +
                      cell[2] * (cell[3] * cell[7] - cell[4] * cell[6]));
    const double sscale = -1.0 / vol;                                         // rdxdr = -pstress / volume (SO3.py:279-280)
    const double t_h2d = now();
    if (np) {
        so3_pair_kernel<<<(np + 3) / 4, 128>>>(P, np, d_rel, d_wt, d_C, d_dC);
        g_launches++;
    }
    so3_center_kernel<<<n_atoms, 128, ne * 16>>>(P, d_ps, d_C, d_ct, h->d_x, ncoefs);
    g_launches++;
    so3_slot_kernel<<<(nseq + 3) / 4, 128>>>(P, nseq, d_ss, d_sp, d_sc, d_Rj, d_dC, d_ct, h->d_dxdr, stress ? h->d_rdxdr : nullptr,
                                            sscale, ncoefs);
    g_launches++;
    so3_self_kernel<<<n_atoms, 128>>>(n_atoms, d_seqs, d_self, d_pos, h->d_dxdr, stress ? h->d_rdxdr : nullptr, sscale, ncoefs);
    g_launches++;
    SO3_CUDA(cudaGetLastError());
    const double t_launch = now();
    // No synchronisation here: every consumer of x / dxdr / rdxdr (cspbo_so3_fetch's copies, the packs built from the
    // device pointers, torch views on the default stream) is ordered behind these kernels on the same stream, and so is any
    // later user of the temporaries handed back to the caching allocator; the host staging vectors above are pageable, so
    // cudaMemcpyAsync had copied them out before it returned.  Saves 0.19 ms of host idling per 192-atom structure.
    if (trace) SO3_CUDA(cudaStreamSynchronize(0));
    cleanup();
    if (trace)
        fprintf(stderr, "so3 trace: pairs %.3f  slots %.3f  alloc+h2d %.3f  launch %.3f  sync %.3f ms (np %d, nseq %d)\n", t_pairs - t_begin,
                t_slots - t_pairs, t_h2d - t_slots, t_launch - t_h2d, now() - t_launch, np, nseq);
    return CSPBO_OK;
#undef SO3_TRY
#undef SO3_CUDA
}

extern "C" int cspbo_so3_device_ptrs(const cspbo_so3 *h, double **x, double **dxdr, double **rdxdr)
{
    if (!h || !x || !dxdr || !rdxdr) { set_error("so3_device_ptrs: null argument"); return CSPBO_E_ARG; }
    *x = h->d_x; *dxdr = h->d_dxdr; *rdxdr = h->stress ? h->d_rdxdr : nullptr;
    return CSPBO_OK;
}

extern "C" int cspbo_so3_fetch(const cspbo_so3 *h, double *x, double *dxdr, double *rdxdr, long long *seq)
{
    if (!h) { set_error("so3_fetch: null handle"); return CSPBO_E_ARG; }
    if (seq) memcpy(seq, h->h_seq.data(), h->h_seq.size() * sizeof(long long));
    if (h->n_atoms == 0) return CSPBO_OK;
    if (x) CSPBO_CUDA(cudaMemcpy(x, h->d_x, (size_t)h->n_atoms * h->ncoefs * 8, cudaMemcpyDeviceToHost));
    if (dxdr) CSPBO_CUDA(cudaMemcpy(dxdr, h->d_dxdr, (size_t)h->n_seq * h->ncoefs * 3 * 8, cudaMemcpyDeviceToHost));
    if (rdxdr) {
        if (!h->stress) { set_error("so3_fetch: stress terms were not computed"); return CSPBO_E_ARG; }
        CSPBO_CUDA(cudaMemcpy(rdxdr, h->d_rdxdr, (size_t)h->n_seq * h->ncoefs * 9 * 8, cudaMemcpyDeviceToHost));
    }
    return CSPBO_OK;
}


// ------------------------------------------------------------------------------------------
// LJ base potential (reference cspbo/calculator.py:68-177, itself ase's pairwise LJ): energy, forces and per-atom
// 3x3 stress sums of a shifted 12-6 potential, subtracted from the training targets / added to the predictions of the
// GP (gaussianprocess.py:457-462,686-692).  One thread per atom over its neighbour pairs, in pair order.
// ------------------------------------------------------------------------------------------
namespace cspbo {
__global__ void lj_atom_kernel(int n_atoms, const int *__restrict__ pair_start, const double *__restrict__ rel, double sigma,
                               double epsilon, double e0, double *__restrict__ energies, double *__restrict__ forces,
                               double *__restrict__ stresses)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_atoms) return;
    double e = 0.0, f[3] = {0.0, 0.0, 0.0}, st[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int p = pair_start[i]; p < pair_start[i + 1]; p++) {
        const double d[3] = {rel[3 * p], rel[3 * p + 1], rel[3 * p + 2]};          // pointing towards the neighbour
        const double r2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
        const double s2 = sigma * sigma / r2, c6 = s2 * s2 * s2, c12 = c6 * c6;
        e += 0.5 * (4.0 * epsilon * (c12 - c6) - e0);
        const double g = -24.0 * epsilon * (2.0 * c12 - c6) / r2;
        for (int a = 0; a < 3; a++) {
            f[a] += g * d[a];
            for (int b = 0; b < 3; b++) st[3 * a + b] += 0.5 * g * d[a] * d[b];
        }
    }
    energies[i] = e;
    for (int a = 0; a < 3; a++) forces[3 * i + a] = f[a];
    for (int a = 0; a < 9; a++) stresses[9 * i + a] = st[a];
}
}  // namespace cspbo

extern "C" int cspbo_lj_compute(int n_atoms, const double *pos, const double *cell, const int *pbc, double rc, double sigma,
                                double epsilon, double *energies, double *forces, double *stresses)
{
    if (n_atoms < 0 || !pos || !cell || !pbc || !energies || !forces || !stresses || !(rc > 0) || !(sigma > 0)) {
        set_error("lj_compute: bad argument");
        return CSPBO_E_ARG;
    }
    if (n_atoms == 0) return CSPBO_OK;
    std::vector<Pair> pairs;
    int rc_ = neighbour_pairs(n_atoms, pos, cell, pbc, rc, pairs);
    if (rc_) return rc_;
    const int np = (int)pairs.size();
    std::vector<int> pair_start((size_t)n_atoms + 1, 0);
    std::vector<double> rel((size_t)std::max(np, 1) * 3);
    for (int p = 0; p < np; p++) {
        pair_start[(size_t)pairs[(size_t)p].i + 1]++;
        for (int k = 0; k < 3; k++) rel[(size_t)p * 3 + k] = pairs[(size_t)p].rel[k];
    }
    for (int i = 0; i < n_atoms; i++) pair_start[(size_t)i + 1] += pair_start[(size_t)i];
    int *d_ps = nullptr;
    double *d_rel = nullptr, *d_out = nullptr;
    std::vector<void *> tmp;
    auto A = [&](size_t bytes, void **q) { int r = dev_alloc_cached(bytes, q); if (!r) tmp.push_back(*q); return r; };
    auto cleanup = [&]() { for (void *q : tmp) dev_free_cached(q); };
    if ((rc_ = A(((size_t)n_atoms + 1) * 4, (void **)&d_ps)) || (rc_ = A(rel.size() * 8, (void **)&d_rel)) ||
        (rc_ = A((size_t)n_atoms * 13 * 8, (void **)&d_out))) { cleanup(); return rc_; }
    cudaError_t e = cudaMemcpyAsync(d_ps, pair_start.data(), ((size_t)n_atoms + 1) * 4, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_rel, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        const double sr6 = pow(sigma / rc, 6.0);
        lj_atom_kernel<<<(n_atoms + 127) / 128, 128>>>(n_atoms, d_ps, d_rel, sigma, epsilon, 4.0 * epsilon * (sr6 * sr6 - sr6), d_out,
                                                          d_out + n_atoms, d_out + (size_t)n_atoms * 4);
        g_launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(energies, d_out, (size_t)n_atoms * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpyAsync(forces, d_out + n_atoms, (size_t)n_atoms * 3 * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaMemcpyAsync(stresses, d_out + (size_t)n_atoms * 4, (size_t)n_atoms * 9 * 8, cudaMemcpyDeviceToHost);
    if (e == cudaSuccess) e = cudaStreamSynchronize(0);
    cleanup();
    if (e != cudaSuccess) { set_error("lj_compute: %s", cudaGetErrorString(e)); return CSPBO_E_CUDA; }
    return CSPBO_OK;
}
