/* shim_common.h -- the reference's native functions return void and have no error channel
 * (dot_kernel.cpp:4, rbf_kernel.cpp:4).  The drop-in shims keep those signatures; a failure of
 * the CUDA path is therefore reported on stderr and the process is aborted: there is no CPU
 * fallback to fall through to. */
#pragma once
#include <stdio.h>
#include <stdlib.h>
#include "cspbo_b200.h"
#define CSPBO_SHIM_CHECK(call)                                                         \
    do {                                                                               \
        int rc__ = (call);                                                             \
        if (rc__ != 0) {                                                               \
            fprintf(stderr, "cspbo_b200: %s failed (%d): %s\n", #call, rc__,           \
                    cspbo_last_error());                                               \
            abort();                                                                   \
        }                                                                              \
    } while (0)
