"""Per-group self covariance (`diag`) of the kernel objects (reference Dot_mb.py:45-85,
RBF_mb.py:45-85), used for the predictive variance.  Built on the GPU block kernels."""
import numpy as np


def kernel_diag(kernel, data):
    raise NotImplementedError("diag is scheduled with the predictive-variance row (SURVEY 8f-3)")
