// msw_error.h -- thread-local last-error string and CUDA error plumbing shared by
// the C-ABI translation units (include/msw_b200.h: "Return value" convention).
#pragma once

#include <cuda_runtime.h>

namespace msw {

int fail(int code, const char *fmt, ...);      // records the message, returns code
int cuda_fail(cudaError_t e, const char *what);  // returns -(int)e
int sm_count();                                // SM count of the current device (cached)

}  // namespace msw

#define MSW_CUDA_TRY(expr)                                                \
    do {                                                                  \
        cudaError_t _e = (expr);                                          \
        if (_e != cudaSuccess) return ::msw::cuda_fail(_e, #expr);        \
    } while (0)
