/* TEST INFRASTRUCTURE ONLY — abort-stubs for the OpenCL entry points and the
 * cl-helper.c functions that the reference's host files reference but never
 * reach on the CPU path timed here (serial bidiag.c + host dDC / twisted /
 * multU / multV).  There is no OpenCL runtime in this image; if any of these
 * is ever called the oracle build is being misused, so fail loudly.
 * Declarations: ref_build/CL/cl.h (ours) and /root/reference/cl-helper.h:90-140.
 */
#include <stdio.h>
#include <stdlib.h>
#include "cl-helper.h"

#define ABSENT(name) do { fprintf(stderr, "oracle/_ref: OpenCL symbol %s is not available " \
        "(the reference's .cl path cannot run here)\n", name); abort(); } while (0)

const char *CHOOSE_INTERACTIVELY = "INTERACTIVE";
const char *cl_error_to_str(cl_int e) { (void)e; return "opencl-absent"; }
void print_platforms_devices(void) { ABSENT("print_platforms_devices"); }
void create_context_on(const char *p, const char *d, cl_uint i, cl_context *c,
                       cl_command_queue *q, int prof)
{ (void)p; (void)d; (void)i; (void)c; (void)q; (void)prof; ABSENT("create_context_on"); }
char *read_file(const char *f) { (void)f; ABSENT("read_file"); return NULL; }
cl_kernel kernel_from_string(cl_context c, char const *k, char const *n, char const *o)
{ (void)c; (void)k; (void)n; (void)o; ABSENT("kernel_from_string"); return NULL; }

cl_mem clCreateBuffer(cl_context c, cl_mem_flags f, size_t s, void *h, cl_int *e)
{ (void)c; (void)f; (void)s; (void)h; (void)e; ABSENT("clCreateBuffer"); return NULL; }
cl_int clSetKernelArg(cl_kernel k, cl_uint i, size_t s, const void *v)
{ (void)k; (void)i; (void)s; (void)v; ABSENT("clSetKernelArg"); return -1; }
cl_int clEnqueueNDRangeKernel(cl_command_queue q, cl_kernel k, cl_uint d, const size_t *o,
                              const size_t *g, const size_t *l, cl_uint n,
                              const cl_event *w, cl_event *e)
{ (void)q; (void)k; (void)d; (void)o; (void)g; (void)l; (void)n; (void)w; (void)e;
  ABSENT("clEnqueueNDRangeKernel"); return -1; }
cl_int clEnqueueReadBuffer(cl_command_queue q, cl_mem m, cl_bool b, size_t o, size_t s,
                           void *p, cl_uint n, const cl_event *w, cl_event *e)
{ (void)q; (void)m; (void)b; (void)o; (void)s; (void)p; (void)n; (void)w; (void)e;
  ABSENT("clEnqueueReadBuffer"); return -1; }
cl_int clEnqueueWriteBuffer(cl_command_queue q, cl_mem m, cl_bool b, size_t o, size_t s,
                            const void *p, cl_uint n, const cl_event *w, cl_event *e)
{ (void)q; (void)m; (void)b; (void)o; (void)s; (void)p; (void)n; (void)w; (void)e;
  ABSENT("clEnqueueWriteBuffer"); return -1; }
cl_int clFinish(cl_command_queue q) { (void)q; ABSENT("clFinish"); return -1; }
cl_int clReleaseMemObject(cl_mem m) { (void)m; ABSENT("clReleaseMemObject"); return -1; }
cl_int clReleaseKernel(cl_kernel k) { (void)k; ABSENT("clReleaseKernel"); return -1; }
cl_int clReleaseCommandQueue(cl_command_queue q) { (void)q; ABSENT("clReleaseCommandQueue"); return -1; }
cl_int clReleaseContext(cl_context c) { (void)c; ABSENT("clReleaseContext"); return -1; }
