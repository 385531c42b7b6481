// common.cuh -- shared host-side helpers for the C-ABI (error reporting, device queries).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/cy4.h"

namespace cy4 {

void set_error(const char *fmt, ...);   // thread-local message, returned by cy4_last_error()
int sm_count();                         // SM count of the current device (cached per device)
void count_launches(int n);             // bookkeeping behind cy4_kernel_launches()

}  // namespace cy4

#define CY4_CHECK_ARG(cond, msg)                 \
    do {                                         \
        if (!(cond)) {                           \
            cy4::set_error("%s", msg);           \
            return -1;                           \
        }                                        \
    } while (0)

#define CY4_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            cy4::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return -2;                                                                       \
        }                                                                                    \
    } while (0)

// Launch-time errors only (no synchronisation).
static inline int cy4_launch_status(const char *what, int n_kernels = 1)
{
    cy4::count_launches(n_kernels);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cy4::set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return -2;
    }
    return 0;
}
