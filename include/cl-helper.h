/* cl-helper.h — compatibility shim.  The reference's drop-in client test-whole-svd.c
 * (line 7) includes "cl-helper.h" although it uses nothing from it; in this library the
 * OpenCL glue is replaced by the CUDA C-ABI layer, so the old name forwards to it. */
#ifndef SVDGPU_CL_HELPER_COMPAT_H
#define SVDGPU_CL_HELPER_COMPAT_H
#include <stdio.h>
#include <stdlib.h>
#include "cuda-helper.h"
#endif
