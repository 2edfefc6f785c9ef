/* Minimal stand-in for <CL/cl.h>, written for this repo (NOT the Khronos header).
 *
 * TEST INFRASTRUCTURE ONLY.  The reference's host C files include
 * cl-helper.h -> <CL/cl.h> (cl-helper.h:36) even on the code paths that never
 * touch OpenCL.  This image has no OpenCL headers or runtime, so oracle/Makefile
 * puts this directory on the include path to let the *unmodified* reference
 * sources under /root/reference compile.  Only the handful of types, constants
 * and prototypes that bidiag_par.c, parallel-twisted.c and svd_gpu.c mention
 * are declared; every function is defined in cl_absent.c and aborts if called.
 */
#ifndef DDC_ORACLE_CL_STANDIN_H
#define DDC_ORACLE_CL_STANDIN_H
#include <stddef.h>
#include <stdint.h>

typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef cl_uint  cl_device_info;
typedef intptr_t cl_context_properties;

typedef struct ddc_cl_opaque_platform *cl_platform_id;
typedef struct ddc_cl_opaque_device   *cl_device_id;
typedef struct ddc_cl_opaque_context  *cl_context;
typedef struct ddc_cl_opaque_queue    *cl_command_queue;
typedef struct ddc_cl_opaque_mem      *cl_mem;
typedef struct ddc_cl_opaque_program  *cl_program;
typedef struct ddc_cl_opaque_kernel   *cl_kernel;
typedef struct ddc_cl_opaque_event    *cl_event;

#define CL_SUCCESS        0
#define CL_FALSE          0
#define CL_TRUE           1
#define CL_MEM_READ_WRITE (1 << 0)

cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void *, cl_int *);
cl_int clSetKernelArg(cl_kernel, cl_uint, size_t, const void *);
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel, cl_uint, const size_t *,
                              const size_t *, const size_t *, cl_uint,
                              const cl_event *, cl_event *);
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void *,
                           cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t,
                            const void *, cl_uint, const cl_event *, cl_event *);
cl_int clFinish(cl_command_queue);
cl_int clReleaseMemObject(cl_mem);
cl_int clReleaseKernel(cl_kernel);
cl_int clReleaseCommandQueue(cl_command_queue);
cl_int clReleaseContext(cl_context);
#endif
