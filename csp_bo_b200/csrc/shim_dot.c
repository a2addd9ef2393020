/* shim_dot.c -> libdot_kernel_b200.so: exports the EXACT symbol names and argument lists the
 * reference's cffi module `cspbo.kernels._dot_kernel` declares (libdot_builder.py:5-9), forwarding
 * to the CUDA implementation in libcspbo_b200.so.  See INTEGRATION.md. */
#include "shim_common.h"

void kee_many(int n1, int n2, int d, int x2i, double zeta, double sigma2, double sigma02,
              double *x1, int *ele1, int *x1_inds, double *x2, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_dot_kee_many(n1, n2, d, x2i, zeta, sigma2, sigma02, x1, ele1, x1_inds, x2, ele2, x2_inds, pout));
}
void kef_many(int n1, int n2, int d, int x2i, double zeta, double *x1, int *ele1, int *x1_inds,
              double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_dot_kef_many(n1, n2, d, x2i, zeta, x1, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout));
}
void kef_many_stress(int n1, int n2, int d, int x2i, double zeta, double *x1, int *ele1, int *x1_inds,
                     double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_dot_kef_many_stress(n1, n2, d, x2i, zeta, x1, ele1, x1_inds, x2, dx2dr, ele2, x2_inds, pout));
}
void kff_many(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
              double *x1, double *dx1dr, int *ele1, int *x1_inds,
              double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_dot_kff_many(n1, n2, n2_start, n2_end, d, x2i, zeta, x1, dx1dr, ele1, x1_inds,
                                        x2, dx2dr, ele2, x2_inds, pout));
}
void kff_many_stress(int n1, int n2, int n2_start, int n2_end, int d, int x2i, double zeta,
                     double *x1, double *dx1dr, int *ele1, int *x1_inds,
                     double *x2, double *dx2dr, int *ele2, int *x2_inds, double *pout)
{
    CSPBO_SHIM_CHECK(cspbo_dot_kff_many_stress(n1, n2, n2_start, n2_end, d, x2i, zeta, x1, dx1dr, ele1, x1_inds,
                                               x2, dx2dr, ele2, x2_inds, pout));
}
