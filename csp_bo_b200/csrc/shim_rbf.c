/* shim_rbf.c -> librbf_kernel_b200.so: exports the EXACT symbol names and argument lists the
 * reference's cffi module `cspbo.kernels._rbf_kernel` declares (librbf_builder.py:5-12), forwarding
 * to the CUDA implementation in libcspbo_b200.so.  See INTEGRATION.md. */
#include "shim_common.h"

void kee_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
              double *x1, int *ele1, int *x1_inds, double *x2, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_rbf_kee_many(n1, n2, d, x2i, zeta, sigma2, l2, x1, ele1, x1_inds, x2, ele2, x2_inds, pout));
}
void kee_many_with_grad(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                        double *x1, int *ele1, int *x1_inds, double *x2, int *ele2, int *x2_inds,
                        double *pout, double *dpout_dl)
{
    CSPBO_SHIM_CHECK(cspbo_rbf_kee_many_with_grad(n1, n2, d, x2i, zeta, sigma2, l2, x1, ele1, x1_inds, x2, ele2, x2_inds,
                                                  pout, dpout_dl));
}
void kef_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
              double *x1, int *ele1, int *x1_inds, double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_rbf_kef_many(n1, n2, d, x2i, zeta, sigma2, l2, x1, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout));
}
void kef_many_with_grad(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l,
                        double *x1, int *ele1, int *x1_inds, double *x2, double *dx2dr, int *ele2, int *x2_inds,
                        double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_rbf_kef_many_with_grad(n1, n2, d, x2i, zeta, sigma2, l, x1, ele1, x1_inds, x2, dx2dr, ele2,
                                                  x2_inds, pout));
}
void kef_many_stress(int n1, int n2, int d, int x2i, double zeta, double sigma2, double l2,
                     double *x1, int *ele1, int *x1_inds, double *x2, double *dx2dr, int *ele2, int *x2_inds,
                     double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_rbf_kef_many_stress(n1, n2, d, x2i, zeta, sigma2, l2, x1, ele1, x1_inds, x2, dx2dr, ele2,
                                               x2_inds, pout));
}
void kff_many(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta, double sigma2, double l2, double tol,
              double *x1, double *dx1dr, int *ele1, int *x1_inds,
              double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_rbf_kff_many(n1, n2, n2_start, n2_end, d, x2i, zeta, sigma2, l2, tol, x1, dx1dr, ele1, x1_inds,
                                        x2, dx2dr, ele2, x2_inds, pout));
}
void kff_many_with_grad(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta, double sigma2, double l,
                        double *x1, double *dx1dr, int *ele1, int *x1_inds,
                        double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout, double *dpout_dl)
{
    CSPBO_SHIM_CHECK(cspbo_rbf_kff_many_with_grad(n1, n2, n2_start, n2_end, d, x2i, zeta, sigma2, l, x1, dx1dr, ele1,
                                                  x1_inds, x2, dx2dr, ele2, x2_inds, pout, dpout_dl));
}
void kff_many_stress(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta, double sigma2, double l2,
                     double tol, double *x1, double *dx1dr, int *ele1, int *x1_inds,
                     double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_rbf_kff_many_stress(n1, n2, n2_start, n2_end, d, x2i, zeta, sigma2, l2, tol, x1, dx1dr, ele1,
                                               x1_inds, x2, dx2dr, ele2, x2_inds, pout));
}
