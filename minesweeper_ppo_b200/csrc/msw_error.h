// msw_error.h -- thread-local last-error string and CUDA error plumbing shared by
// the C-ABI translation units (include/msw_b200.h: "Return value" convention).
#pragma once

#include <cuda_runtime.h>

namespace msw {

int fail(int code, const char *fmt, ...);      // records the message, returns code
int cuda_fail(cudaError_t e, const char *what);  // returns -(int)e
int sm_count();                                // SM count of the current device (cached)

}  // namespace msw

namespace msw {
// cudaFuncSetAttribute is per device (context): remember per DEVICE, not per thread, what has been set.
struct PerDeviceOnce {
    unsigned long long bits[4] = {0, 0, 0, 0};                 // devices 0..255
    bool need(int dev) const
    {
        if (dev < 0 || dev >= 256) return true;
        return !(__atomic_load_n(&bits[dev >> 6], __ATOMIC_ACQUIRE) >> (dev & 63) & 1ull);
    }
    void done(int dev)
    {
        if (dev >= 0 && dev < 256) __atomic_fetch_or(&bits[dev >> 6], 1ull << (dev & 63), __ATOMIC_RELEASE);
    }
};
}  // namespace msw

#define MSW_CUDA_TRY(expr)                                                \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) return ::msw::cuda_fail(_e, #expr);        \
    } while (0)

// Raise a kernel's dynamic shared-memory limit once per device.  `kernel` may need parentheses when its
// template argument list contains commas.
#define MSW_SET_MAX_SMEM(kernel, bytes)                                                                         \
    do {                                                                                                        \
        static ::msw::PerDeviceOnce _once;                                                                      \
        int _dev = 0;                                                                                           \
        MSW_CUDA_TRY(cudaGetDevice(&_dev));                                                                     \
        if (_once.need(_dev)) {                                                                                 \
            MSW_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
            _once.done(_dev);                                                                                   \
        }                                                                                                       \
    } while (0)
