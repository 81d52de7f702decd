/* Minimal stand-in for <CL/opencl.h>, written for the oracle build only.
 *
 * TEST INFRASTRUCTURE.  The reference's host sources include the Altera OpenCL
 * headers, which do not exist in this image.  Nothing here computes a score: every
 * entry point is an inert stub (see shim_rt.cpp) so that the reference's *host AVX2
 * path* can be compiled in place and run as the parity oracle / CPU baseline. */
#ifndef OSWALD_ORACLE_CL_SHIM_H
#define OSWALD_ORACLE_CL_SHIM_H
#include <stddef.h>
#include <stdint.h>

typedef int32_t  cl_int;
typedef uint32_t cl_uint;
typedef uint64_t cl_ulong;
typedef cl_uint  cl_bool;
typedef cl_ulong cl_bitfield;
typedef cl_bitfield cl_device_type;
typedef cl_bitfield cl_mem_flags;
typedef cl_bitfield cl_command_queue_properties;
typedef cl_uint  cl_device_info;
typedef cl_uint  cl_platform_info;
typedef intptr_t cl_context_properties;

struct osw_shim_handle;                       /* never defined: handles are opaque tokens */
typedef struct osw_shim_handle *cl_platform_id;
typedef struct osw_shim_handle *cl_device_id;
typedef struct osw_shim_handle *cl_context;
typedef struct osw_shim_handle *cl_command_queue;
typedef struct osw_shim_handle *cl_mem;
typedef struct osw_shim_handle *cl_program;
typedef struct osw_shim_handle *cl_kernel;
typedef struct osw_shim_handle *cl_event;

enum { CL_SUCCESS = 0, CL_FALSE = 0, CL_TRUE = 1 };

#define CL_DEVICE_TYPE_ALL                          0xFFFFFFFFu
#define CL_MEM_READ_WRITE                           (1u << 0)
#define CL_MEM_READ_ONLY                            (1u << 2)
#define CL_MEM_COPY_HOST_PTR                        (1u << 5)
#define CL_QUEUE_OUT_OF_ORDER_EXEC_MODE_ENABLE      (1u << 0)
#define CL_QUEUE_PROFILING_ENABLE                   (1u << 1)

/* info ids: only their distinctness matters to the stub */
enum {
    CL_PLATFORM_NAME = 0x0900, CL_PLATFORM_VENDOR, CL_PLATFORM_VERSION,
    CL_DEVICE_NAME = 0x1000, CL_DEVICE_VENDOR, CL_DEVICE_VENDOR_ID, CL_DEVICE_VERSION,
    CL_DRIVER_VERSION, CL_DEVICE_ADDRESS_BITS, CL_DEVICE_AVAILABLE, CL_DEVICE_ENDIAN_LITTLE,
    CL_DEVICE_GLOBAL_MEM_CACHE_SIZE, CL_DEVICE_GLOBAL_MEM_CACHELINE_SIZE,
    CL_DEVICE_GLOBAL_MEM_SIZE, CL_DEVICE_IMAGE_SUPPORT, CL_DEVICE_LOCAL_MEM_SIZE,
    CL_DEVICE_MAX_CLOCK_FREQUENCY, CL_DEVICE_MAX_COMPUTE_UNITS, CL_DEVICE_MAX_CONSTANT_ARGS,
    CL_DEVICE_MAX_CONSTANT_BUFFER_SIZE, CL_DEVICE_MAX_WORK_ITEM_DIMENSIONS,
    CL_DEVICE_MEM_BASE_ADDR_ALIGN, CL_DEVICE_MIN_DATA_TYPE_ALIGN_SIZE,
    CL_DEVICE_PREFERRED_VECTOR_WIDTH_CHAR, CL_DEVICE_PREFERRED_VECTOR_WIDTH_SHORT,
    CL_DEVICE_PREFERRED_VECTOR_WIDTH_INT, CL_DEVICE_PREFERRED_VECTOR_WIDTH_LONG,
    CL_DEVICE_PREFERRED_VECTOR_WIDTH_FLOAT, CL_DEVICE_PREFERRED_VECTOR_WIDTH_DOUBLE,
    CL_DEVICE_QUEUE_PROPERTIES
};

#ifdef __cplusplus
extern "C" {
#endif
cl_int clGetPlatformInfo(cl_platform_id, cl_platform_info, size_t, void *, size_t *);
cl_int clGetDeviceInfo(cl_device_id, cl_device_info, size_t, void *, size_t *);
cl_context clCreateContext(const cl_context_properties *, cl_uint, const cl_device_id *,
                           void (*)(const char *, const void *, size_t, void *), void *, cl_int *);
cl_command_queue clCreateCommandQueue(cl_context, cl_device_id, cl_command_queue_properties, cl_int *);
cl_int clBuildProgram(cl_program, cl_uint, const cl_device_id *, const char *,
                      void (*)(cl_program, void *), void *);
cl_kernel clCreateKernel(cl_program, const char *, cl_int *);
cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t, void *, cl_int *);
cl_int clSetKernelArg(cl_kernel, cl_uint, size_t, const void *);
cl_int clEnqueueWriteBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, const void *,
                            cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueReadBuffer(cl_command_queue, cl_mem, cl_bool, size_t, size_t, void *,
                           cl_uint, const cl_event *, cl_event *);
cl_int clEnqueueNDRangeKernel(cl_command_queue, cl_kernel, cl_uint, const size_t *, const size_t *,
                              const size_t *, cl_uint, const cl_event *, cl_event *);
cl_int clFinish(cl_command_queue);
cl_int clWaitForEvents(cl_uint, const cl_event *);
cl_int clReleaseEvent(cl_event);
cl_int clReleaseMemObject(cl_mem);
cl_int clReleaseKernel(cl_kernel);
cl_int clReleaseCommandQueue(cl_command_queue);
cl_int clReleaseProgram(cl_program);
cl_int clReleaseContext(cl_context);
#ifdef __cplusplus
}
#endif
#endif
