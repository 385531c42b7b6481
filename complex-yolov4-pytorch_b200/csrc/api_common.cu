// api_common.cu -- version / error / device entry points of the C-ABI.
#include <atomic>
#include <cstdarg>

#include "common.cuh"

namespace cy4 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace cy4

extern "C" {

int cy4_version(void) { return CY4_VERSION; }

const char *cy4_last_error(void) { return cy4::g_err; }

long long cy4_kernel_launches(int reset)
{
    return reset ? cy4::g_launches.exchange(0) : cy4::g_launches.load();
}

/* A CUDA graph replay launches the kernels that were captured into it without passing through the entry points that count:
 * the caller reports them (n = number of this library's kernel nodes in the replayed graph). */
int cy4_note_graph_replay(int n_kernels)
{
    if (n_kernels > 0) cy4::g_launches.fetch_add(n_kernels, std::memory_order_relaxed);
    return 0;
}

int cy4_device_ok(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        cy4::set_error("cy4_device_ok: no CUDA device visible");
        return -1;
    }
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) {
        cy4::set_error("cy4_device_ok: device %d has compute capability %d.x, library is built for sm_100a only", dev, major);
        return -3;
    }
    return 0;
}

}  // extern "C"
