/* Stand-in for Altera's AOCL_Utils.h (oracle build only, TEST INFRASTRUCTURE).
 * Provides just the names the reference's host/src uses (utils.c:99-173, main.c:18-24). */
#ifndef OSWALD_ORACLE_AOCL_SHIM_H
#define OSWALD_ORACLE_AOCL_SHIM_H
#include <string>
#include <cstdio>
#include <cstdlib>
#include "CL/opencl.h"

namespace aocl_utils {

/* Owning array with bounds-tolerant indexing: the reference indexes `kernels[d]`
 * with an uninitialised d (HybridSearch.c:644); out-of-range reads land on a
 * private dummy slot instead of wild memory. */
template <typename T>
class scoped_array {
    T *p_; size_t n_; mutable T dummy_;
public:
    scoped_array() : p_(0), n_(0), dummy_() {}
    ~scoped_array() { delete[] p_; }
    void reset(T *q) { delete[] p_; p_ = q; n_ = q ? (size_t)-1 : 0; }
    void reset(size_t n) { delete[] p_; p_ = new T[n](); n_ = n; }
    void reset(int n) { reset((size_t)n); }
    void reset(unsigned n) { reset((size_t)n); }
    T &operator[](size_t i) const { return (p_ && i < n_ && i < (1u << 20)) ? p_[i] : dummy_; }
    operator T *() const { return p_; }
private:
    scoped_array(const scoped_array &);
    scoped_array &operator=(const scoped_array &);
};

bool setCwdToExeDir();
cl_platform_id findPlatform(const char *name);
cl_device_id *getDevices(cl_platform_id, cl_device_type, cl_uint *count);
std::string getBoardBinaryFile(const char *prefix, cl_device_id);
cl_program createProgramFromBinary(cl_context, const char *file, const cl_device_id *, unsigned n);
void shim_check(int line, const char *file, cl_int status, const char *msg);

}  // namespace aocl_utils

#define checkError(status, ...) aocl_utils::shim_check(__LINE__, __FILE__, (status), "" __VA_ARGS__)

#endif
