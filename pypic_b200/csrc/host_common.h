// Host-side helpers shared by the .cu translation units of libpic_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>
#include "../../include/pic_b200.h"

namespace pic {

void set_error(const char* fmt, ...);
int device_sm_count();
int max_optin_smem();

#define PIC_CHECK_CUDA(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            pic::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                           __FILE__, __LINE__);                                     \
            return PIC_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define PIC_CHECK_LAUNCH()                                                          \
    do {                                                                            \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            pic::set_error("kernel launch failed: %s (%s:%d)",                      \
                           cudaGetErrorString(_e), __FILE__, __LINE__);             \
            return PIC_ERR_CUDA;                                                    \
        }                                                                           \
    } while (0)

#define PIC_REQUIRE(cond, msg)                                                      \
    do {                                                                            \
        if (!(cond)) {                                                              \
            pic::set_error("invalid argument: %s (%s:%d)", msg, __FILE__, __LINE__);\
            return PIC_ERR_ARG;                                                     \
        }                                                                           \
    } while (0)

// persistent-style grid: enough CTAs to fill every SM `per_sm` times, never more
// than the work needs
inline int grid_for(long long n, int block, int per_sm) {
    long long need = (n + block - 1) / block;
    long long cap = (long long)device_sm_count() * per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace pic
