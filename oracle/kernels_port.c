/*
 * oracle/kernels_port.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * A plain-C, single-threaded CPU restatement of the arithmetic of the reference's
 * native GP-kernel builders
 *     cspbo/kernels/dot_kernel.cpp   (Dot many-body kernel,  5 entry points)
 *     cspbo/kernels/rbf_kernel.cpp   (RBF many-body kernel,  8 entry points)
 * written from the formulas, not from the text: one generic pair engine serves all
 * thirteen entry points instead of thirteen hand-unrolled loops.  Every exported
 * function keeps the reference's argument list (libdot_builder.py:5-9,
 * librbf_builder.py:5-12) with an `oracle_dot_` / `oracle_rbf_` prefix.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this port (a) against the
 * unmodified reference C++ compiled into oracle/_ref/ whenever that library exists and
 * (b) against tests/golden/*.npz, vectors produced by the unmodified reference C++ on the
 * reference's own fixtures (Cpp_code/X1.npy, X2.npy, X1_big.npy, Kff_QZ/X*_EE.npy) with
 * tests/golden/make_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.
 *
 * Conventions shared by every function (dot_kernel.cpp:4-55 is the template):
 *   x1[n1,d], x2[n2,d]          stacked descriptor rows, row-major
 *   dx?dr[n?,d,nc]              per-row Cartesian derivative, nc = 3 (force) or 9 (force+stress)
 *   ele?[row]                   species id; pairs with different ids contribute nothing
 *   x?_inds[row]                group (structure or force-atom) id of the row
 *   x2i                         number of x2 groups == row stride of pout
 *   pout                        caller-zeroed accumulator, only ever += / -=
 *   rows with |x| <= 1e-8 are skipped (dot_kernel.cpp:25,37)
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_EPS 1e-8

typedef struct {
    int is_rbf;      /* 0: Dot, 1: RBF */
    double zeta;
    double sigma2;   /* Dot kee, all RBF */
    double sigma02;  /* Dot kee */
    double l;        /* RBF */
    double l2;       /* RBF */
    double tol;      /* RBF kff / kff_stress skip threshold */
    int use_tol;     /* 0 for every function that has no tol test */
} oracle_hyper;

static double row_norm(const double *x, int d)
{
    double acc = 0.0;
    for (int i = 0; i < d; i++) acc += x[i] * x[i];
    return sqrt(acc);
}

static double row_dot(const double *a, const double *b, int d)
{
    double acc = 0.0;
    for (int i = 0; i < d; i++) acc += a[i] * b[i];
    return acc;
}

/* ------------------------------------------------------------------ K_EE ---- */
/* Dot: dot_kernel.cpp:4-55   pout[i,j] += sigma2*(c^zeta + sigma02)
 * RBF: rbf_kernel.cpp:4-48   pout[i,j] += sigma2*exp(-(1-c^zeta)/(2 l2))
 * RBF grad: rbf_kernel.cpp:50-97  dpout_dl[i,j] += K*(1-D)                      */
static void kee_engine(const oracle_hyper *h, int n1, int n2, int d, int x2i,
                       const double *x1, const int *ele1, const int *g1,
                       const double *x2, const int *ele2, const int *g2,
                       double *pout, double *dpout_dl)
{
    double *nrm2 = (double *)malloc(sizeof(double) * (size_t)(n2 > 0 ? n2 : 1));
    for (int b = 0; b < n2; b++) nrm2[b] = row_norm(x2 + (size_t)b * d, d);

    for (int a = 0; a < n1; a++) {
        const double *xa = x1 + (size_t)a * d;
        double na = row_norm(xa, d);
        if (!(na > ORACLE_EPS)) continue;
        for (int b = 0; b < n2; b++) {
            if (ele1[a] != ele2[b] || !(nrm2[b] > ORACLE_EPS)) continue;
            const double *xb = x2 + (size_t)b * d;
            double s = row_dot(xa, xb, d);
            size_t o = (size_t)g1[a] * x2i + g2[b];
            if (!h->is_rbf) {
                double c = s * (1.0 / (na * nrm2[b]));           /* dot_kernel.cpp:38,45 */
                pout[o] += h->sigma2 * (pow(c, h->zeta) + h->sigma02);
            } else {
                double c = s / (na * nrm2[b]);                   /* rbf_kernel.cpp:40 */
                double D = pow(c, h->zeta);
                double K = h->sigma2 * exp(-(1 - D) / (2 * h->l2));
                pout[o] += K;
                if (dpout_dl) dpout_dl[o] += K * (1 - D);
            }
        }
    }
    free(nrm2);
}

/* ------------------------------------------------------------------ K_EF ---- */
/* Dot: dot_kernel.cpp:57-128 (nc=3), :131-214 (nc=9)
 *      g[i]   = zeta c^(zeta-1) * ( x1[i]/(n1 n2) - (s/(n1 n2)/n2^2) x2[i] )
 *      pout[(i,j),l] += sum_i dx2dr[b,i,l] g[i]
 * RBF: rbf_kernel.cpp:100-170 (nc=3), :255-337 (nc=9), :172-252 (grad: 6 columns)
 *      term_l = sum_i dx2dr[b,i,l] * g[i] * dK/dD,  pout -= term,
 *      grad   pout[3+l] += term_l * ((D-1)/l^3 + 2/l)                              */
static void kef_engine(const oracle_hyper *h, int with_grad, int nc,
                       int n1, int n2, int d, int x2i,
                       const double *x1, const int *ele1, const int *g1,
                       const double *x2, const double *dx2dr, const int *ele2, const int *g2,
                       double *pout)
{
    int ncout = with_grad ? 6 : nc;
    double *g = (double *)malloc(sizeof(double) * (size_t)d);
    double *nrm2 = (double *)malloc(sizeof(double) * (size_t)(n2 > 0 ? n2 : 1));
    for (int b = 0; b < n2; b++) nrm2[b] = row_norm(x2 + (size_t)b * d, d);
    double inv_l = h->is_rbf ? 1 / h->l : 0.0;
    double inv_l3 = h->is_rbf ? inv_l / (h->l * h->l) : 0.0;

    for (int a = 0; a < n1; a++) {
        const double *xa = x1 + (size_t)a * d;
        double na = row_norm(xa, d);
        if (!(na > ORACLE_EPS)) continue;
        for (int b = 0; b < n2; b++) {
            double nb = nrm2[b];
            if (ele1[a] != ele2[b] || !(nb > ORACLE_EPS)) continue;
            const double *xb = x2 + (size_t)b * d;
            const double *db = dx2dr + (size_t)b * d * nc;
            double inv12 = 1.0 / (na * nb);
            double s = row_dot(xa, xb, d);
            double s13 = s * inv12 / (nb * nb);
            double c = s * inv12;
            double d1 = h->zeta * pow(c, h->zeta - 1);
            double D = 0, dKdD = 1.0;
            if (h->is_rbf) {
                D = pow(c, h->zeta);
                double K = h->sigma2 * exp(-(1 - D) / (2 * h->l2));
                dKdD = K / (2 * h->l2);
            }
            for (int i = 0; i < d; i++) g[i] = d1 * (xa[i] * inv12 - s13 * xb[i]);

            double *o = pout + ((size_t)g1[a] * x2i + g2[b]) * ncout;
            for (int l = 0; l < nc; l++) {
                double acc = 0.0;
                if (h->is_rbf) {
                    for (int i = 0; i < d; i++) acc += db[(size_t)i * nc + l] * g[i] * dKdD;
                    o[l] -= acc;
                    if (with_grad) o[3 + l] += acc * ((D - 1) * inv_l3 + 2 * inv_l);
                } else {
                    for (int i = 0; i < d; i++) acc += db[(size_t)i * nc + l] * g[i];
                    o[l] += acc;
                }
            }
        }
    }
    free(g);
    free(nrm2);
}

/* ------------------------------------------------------------------ K_FF ---- */
/* Per same-species pair (a,b) build the d x d matrix
 *   Dot (dot_kernel.cpp:270-289):
 *     M[i][j] = p[i] q[j] * (zeta-1) c^(zeta-2) + c^(zeta-1) * H[i][j]
 *   RBF (rbf_kernel.cpp:404-425, 543-565):
 *     M[i][j] = dK/dD * ( zeta (c^(zeta-1) H[i][j] + (zeta-1) c^(zeta-2) p[i] q[j])
 *                         + (1/(2 l2)) (zeta c^(zeta-1) p[i]) (zeta c^(zeta-1) q[j]) )
 *     ML[i][j] = (1/(2 l2)) (zeta c^(zeta-1) p[i]) (zeta c^(zeta-1) q[j])       (grad only)
 * with p = dc/dx1 = x2/(n1 n2) - x1 s/(n1^3 n2),  q = dc/dx2 = x1/(n1 n2) - x2 s/(n1 n2^3),
 *      H = d2c/dx1dx2 = (I - x2 x2^T/n2^2)/(n1 n2) + (x1 x2^T s/n2^2 - x1 x1^T)/(n1^3 n2),
 * then contract  out[k][l] = sum_ij dx1dr[a,i,k] M[i][j] dx2dr[b,j,l]
 * (dot_kernel.cpp:291-323) and scatter to pout[((k + g1*nc1)*x2i + g2)*3 + l].
 * RBF grad (rbf_kernel.cpp:621-629):
 *   dpout_dl -= out*((D-1)/l^3 + 2/l) + (2/l) dK/dD * (A^T ML B)                        */
static void kff_engine(const oracle_hyper *h, int with_grad, int nc1,
                       int n1, int n2, int n2_start, int n2_end, int d, int x2i,
                       const double *x1, const double *dx1dr, const int *ele1, const int *g1,
                       const double *x2, const double *dx2dr, const int *ele2, const int *g2,
                       double *pout, double *dpout_dl)
{
    (void)n2;
    double *M = (double *)malloc(sizeof(double) * (size_t)d * d);
    double *ML = (double *)malloc(sizeof(double) * (size_t)d * d);
    double *T = (double *)malloc(sizeof(double) * (size_t)d * nc1);
    double *TL = (double *)malloc(sizeof(double) * (size_t)d * nc1);
    double inv_l = h->is_rbf ? 1 / h->l : 0.0;
    double inv_l3 = h->is_rbf ? inv_l / (h->l * h->l) : 0.0;

    for (int a = 0; a < n1; a++) {
        const double *xa = x1 + (size_t)a * d;
        const double *da = dx1dr + (size_t)a * d * nc1;
        double na = row_norm(xa, d);
        if (!(na > ORACLE_EPS)) continue;
        for (int b = n2_start; b < n2_end; b++) {
            if (ele1[a] != ele2[b]) continue;
            const double *xb = x2 + (size_t)b * d;
            double nb = row_norm(xb, d);
            if (!(nb > ORACLE_EPS)) continue;
            const double *db = dx2dr + (size_t)b * d * 3;

            double inv22 = 1.0 / (nb * nb);
            double inv12 = 1.0 / (na * nb);
            double inv31 = inv12 / (na * na);
            double s = row_dot(xa, xb, d);
            double s31 = s * inv31, s13 = s * inv12 * inv22, s02 = s * inv22;
            double c = s * inv12;
            double D = 0, dKdD = 1.0;
            if (h->is_rbf) {
                D = pow(c, h->zeta);
                double K = h->sigma2 * exp(-(1 - D) / (2 * h->l2));
                dKdD = K / (2 * h->l2);
                if (h->use_tol && !(dKdD > h->tol)) continue;     /* rbf_kernel.cpp:394,697 */
            }
            double w2 = pow(c, h->zeta - 2);
            double w1 = c * w2;
            w2 = w2 * (h->zeta - 1);

            for (int i = 0; i < d; i++) {
                double p = xb[i] * inv12 - xa[i] * s31;
                for (int j = 0; j < d; j++) {
                    double q = xa[j] * inv12 - xb[j] * s13;
                    double kron = (i == j) ? 1.0 : 0.0;
                    double H = (kron - xb[i] * xb[j] * inv22) * inv12 +
                               (xa[i] * xb[j] * s02 - xa[i] * xa[j]) * inv31;
                    if (!h->is_rbf) {
                        M[i * d + j] = (p * q) * w2 + w1 * H;
                    } else {
                        double dDp = h->zeta * w1 * p, dDq = h->zeta * w1 * q;
                        double d2D = h->zeta * (w1 * H + w2 * p * q);
                        double ml = (1 / (2 * h->l2)) * dDp * dDq;
                        ML[i * d + j] = ml;
                        M[i * d + j] = dKdD * (d2D + ml);
                    }
                }
            }
            /* T[j][k] = sum_i dx1dr[a,i,k] M[i][j] */
            for (int j = 0; j < d; j++)
                for (int k = 0; k < nc1; k++) {
                    double acc = 0.0, accl = 0.0;
                    for (int i = 0; i < d; i++) {
                        acc += da[(size_t)i * nc1 + k] * M[i * d + j];
                        if (with_grad) accl += da[(size_t)i * nc1 + k] * ML[i * d + j];
                    }
                    T[j * nc1 + k] = acc;
                    TL[j * nc1 + k] = accl;
                }
            for (int k = 0; k < nc1; k++)
                for (int l = 0; l < 3; l++) {
                    double acc = 0.0, accl = 0.0;
                    for (int j = 0; j < d; j++) {
                        acc += T[j * nc1 + k] * db[(size_t)j * 3 + l];
                        if (with_grad) accl += TL[j * nc1 + k] * db[(size_t)j * 3 + l];
                    }
                    size_t o = (((size_t)k + (size_t)g1[a] * nc1) * x2i + g2[b]) * 3 + l;
                    pout[o] += acc;
                    if (with_grad)
                        dpout_dl[o] -= (acc * ((D - 1) * inv_l3 + 2 * inv_l) +
                                        (inv_l * 2) * accl * dKdD);
                }
        }
    }
    free(M); free(ML); free(T); free(TL);
}

/* ====================================================== exported entry points */

#define DOT_H(z, s2, s02) oracle_hyper h = {0, (z), (s2), (s02), 0.0, 0.0, 0.0, 0}
#define RBF_H(z, s2, l_, l2_, tol_, use_) oracle_hyper h = {1, (z), (s2), 0.0, (l_), (l2_), (tol_), (use_)}

/* --- Dot: libdot_builder.py:5-9 --- */
void oracle_dot_kee_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double sigma02,
                         double *x1, int *ele1, int *x1_inds,
                         double *x2, int *ele2, int *x2_inds, double *pout)
{
    DOT_H(zeta, sigma2, sigma02);
    kee_engine(&h, n1, n2, d, x2i, x1, ele1, x1_inds, x2, ele2, x2_inds, pout, NULL);
}

void oracle_dot_kef_many(int n1, int n2, int d, int x2i, double zeta,
                         double *x1, int *ele1, int *x1_inds,
                         double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    DOT_H(zeta, 1.0, 0.0);
    kef_engine(&h, 0, 3, n1, n2, d, x2i, x1, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout);
}

void oracle_dot_kef_many_stress(int n1, int n2, int d, int x2i, double zeta,
                                double *x1, int *ele1, int *x1_inds,
                                double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    DOT_H(zeta, 1.0, 0.0);
    kef_engine(&h, 0, 9, n1, n2, d, x2i, x1, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout);
}

void oracle_dot_kff_many(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                         double *x1, double *dx1dr, int *ele1, int *x1_inds,
                         double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    DOT_H(zeta, 1.0, 0.0);
    kff_engine(&h, 0, 3, n1, n2, n2_start, n2_end, d, x2i, x1, dx1dr, ele1, x1_inds,
               x2, dx2dr, ele2, x2_inds, pout, NULL);
}

void oracle_dot_kff_many_stress(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                                double *x1, double *dx1dr, int *ele1, int *x1_inds,
                                double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    DOT_H(zeta, 1.0, 0.0);
    kff_engine(&h, 0, 9, n1, n2, n2_start, n2_end, d, x2i, x1, dx1dr, ele1, x1_inds,
               x2, dx2dr, ele2, x2_inds, pout, NULL);
}

/* --- RBF: librbf_builder.py:5-12 --- */
void oracle_rbf_kee_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                         double *x1, int *ele1, int *x1_inds,
                         double *x2, int *ele2, int *x2_inds, double *pout)
{
    RBF_H(zeta, sigma2, sqrt(l2), l2, 0.0, 0);
    kee_engine(&h, n1, n2, d, x2i, x1, ele1, x1_inds, x2, ele2, x2_inds, pout, NULL);
}

void oracle_rbf_kee_many_with_grad(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                                   double *x1, int *ele1, int *x1_inds,
                                   double *x2, int *ele2, int *x2_inds,
                                   double *pout, double *dpout_dl)
{
    RBF_H(zeta, sigma2, sqrt(l2), l2, 0.0, 0);
    kee_engine(&h, n1, n2, d, x2i, x1, ele1, x1_inds, x2, ele2, x2_inds, pout, dpout_dl);
}

void oracle_rbf_kef_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                         double *x1, int *ele1, int *x1_inds,
                         double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    RBF_H(zeta, sigma2, sqrt(l2), l2, 0.0, 0);
    kef_engine(&h, 0, 3, n1, n2, d, x2i, x1, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout);
}

/* NB: takes l, not l^2 (rbf_kernel.cpp:173,185-187) */
void oracle_rbf_kef_many_with_grad(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l,
                                   double *x1, int *ele1, int *x1_inds,
                                   double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    RBF_H(zeta, sigma2, l, l * l, 0.0, 0);
    kef_engine(&h, 1, 3, n1, n2, d, x2i, x1, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout);
}

void oracle_rbf_kef_many_stress(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                                double *x1, int *ele1, int *x1_inds,
                                double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    RBF_H(zeta, sigma2, sqrt(l2), l2, 0.0, 0);
    kef_engine(&h, 0, 9, n1, n2, d, x2i, x1, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout);
}

void oracle_rbf_kff_many(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                         double sigma2, double l2, double tol,
                         double *x1, double *dx1dr, int *ele1, int *x1_inds,
                         double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    RBF_H(zeta, sigma2, sqrt(l2), l2, tol, 1);
    kff_engine(&h, 0, 3, n1, n2, n2_start, n2_end, d, x2i, x1, dx1dr, ele1, x1_inds,
               x2, dx2dr, ele2, x2_inds, pout, NULL);
}

/* NB: takes l, not l^2, and has no tol test (rbf_kernel.cpp:475,490-492,530-533) */
void oracle_rbf_kff_many_with_grad(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                                   double sigma2, double l,
                                   double *x1, double *dx1dr, int *ele1, int *x1_inds,
                                   double *x2, double *dx2dr, int *ele2, int *x2_inds,
                                   double *pout, double *dpout_dl)
{
    RBF_H(zeta, sigma2, l, l * l, 0.0, 0);
    kff_engine(&h, 1, 3, n1, n2, n2_start, n2_end, d, x2i, x1, dx1dr, ele1, x1_inds,
               x2, dx2dr, ele2, x2_inds, pout, dpout_dl);
}

void oracle_rbf_kff_many_stress(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                                double sigma2, double l2, double tol,
                                double *x1, double *dx1dr, int *ele1, int *x1_inds,
                                double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    RBF_H(zeta, sigma2, sqrt(l2), l2, tol, 1);
    kff_engine(&h, 0, 9, n1, n2, n2_start, n2_end, d, x2i, x1, dx1dr, ele1, x1_inds,
               x2, dx2dr, ele2, x2_inds, pout, NULL);
}
