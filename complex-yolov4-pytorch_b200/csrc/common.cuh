// common.cuh -- shared host-side helpers for the C-ABI (error reporting, device queries).
#pragma once
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <utility>

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

// ---- programmatic dependent launch (Blackwell / Hopper griddepcontrol) ---------------------------------------------------
// A training step is ~900 dependent kernels, most of them 10-100 us long: the fixed cost per kernel boundary (block
// scheduling, barrier / TMEM / tensor-map set-up, the first loads) is paid ~900 times.  Every hot kernel therefore
//   * calls pdl_trigger() first thing: once ALL blocks of a grid have done so (= its last wave is resident) the next kernel
//     of the stream may start placing blocks on whatever SM resources are free,
//   * does its set-up that touches no global data (shared-memory carve-up, mbarrier init, TMEM allocation, descriptor prefetch),
//   * calls pdl_wait() -- every thread, unconditionally -- before its first access to global memory: it returns when the
//     preceding grids have COMPLETED and their writes are visible.  (A kernel launched without the attribute returns at once.)
// The attribute is a permission (cudaLaunchAttributeProgrammaticStreamSerialization): predecessors that are not kernels, or
// kernels that never trigger (torch's), simply give the ordinary full dependency.  Only kernels that contain pdl_wait() are
// ever launched with it (pdl_launch_attr).  Option "pdl" (conv_api.cu).
namespace cy4 { extern int g_pdl; }
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
// appends the attribute to a[n] when the option is on; returns the new attribute count
static inline int pdl_launch_attr(cudaLaunchAttribute *a, int n)
{
    if (!cy4::g_pdl) return n;
    a[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a[n].val.programmaticStreamSerializationAllowed = 1;
    return n + 1;
}
// <<<grid, block, 0, stream>>> with the PDL permission: for kernels that call pdl_wait()
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args &&...args)
{
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
    cudaLaunchAttribute at[1];
    cfg.attrs = at;
    cfg.numAttrs = pdl_launch_attr(at, 0);
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

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
