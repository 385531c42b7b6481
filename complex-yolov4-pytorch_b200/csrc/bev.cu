// bev.cu -- SURVEY section 8 (f3): LiDAR point cloud -> 3-channel bird's-eye-view map (intensity, height, density),
// the producer of the [B,3,608,608] input of the training step.
// Reference: src/data_process/kitti_bev_utils.py:18-36 (removePoints) and :39-76 (makeBVFeature), which sort the
// points three ways (np.lexsort) and call np.unique twice per frame on the CPU.  Here: one scatter pass
// (64-bit atomicMax of (z, first-point-wins) per cell + a count) and one pass over the map.
//
// Arithmetic follows numpy's float32 rules of the reference: z - minZ, x / discretization, y / discretization are
// float32 operations (python scalars are weak), floor, then `np.int_(floor(y / d) + (W + 1) / 2)` truncates.
#include "common.cuh"

namespace cy4 {

__device__ __forceinline__ unsigned int orderable(float z)
{
    const unsigned int u = __float_as_uint(z + 0.0f);        // (-0.0 + 0.0 = +0.0: the two zeros tie as in np.lexsort)
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_orderable(unsigned int k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// cells: [B][H*W] 64-bit keys = orderable(z) << 32 | ~point_index (max => highest z, then lowest index: the stable
// lexsort order of the reference), counts: [B][H*W]
__global__ void __launch_bounds__(256)
bev_scatter_kernel(const float4 *__restrict__ pts, const int64_t *__restrict__ offsets, int B, cy4_bev_desc d,
                   unsigned long long *__restrict__ cells, unsigned int *__restrict__ counts, int *__restrict__ dropped)
{
    const int b = blockIdx.y;
    const int64_t p0 = offsets[b], n = offsets[b + 1] - p0;
    const float half_w = (float)(d.W + 1) * 0.5f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 p = __ldg(pts + p0 + i);
        float z = p.z;
        if (d.apply_filter) {                                                     // removePoints, :28-34
            if (!(p.x >= d.minX && p.x <= d.maxX && p.y >= d.minY && p.y <= d.maxY && p.z >= d.minZ && p.z <= d.maxZ)) continue;
            z = p.z - d.minZ;
        }
        const float fx = floorf(__fdiv_rn(p.x, d.discretization));              // :45
        const float fy = truncf(floorf(__fdiv_rn(p.y, d.discretization)) + half_w);   // :46
        // rows / columns H, W of the reference's (H+1) x (W+1) scratch maps are cropped away (:71-73)
        if (!(fx >= 0.f && fx < (float)d.H && fy >= 0.f && fy < (float)d.W)) {
            if (!(fx == (float)d.H || fy == (float)d.W)) atomicAdd(dropped + b, 1);      // outside even the scratch map
            continue;
        }
        const int cell = (int)fx * d.W + (int)fy;
        const unsigned long long key = ((unsigned long long)orderable(z) << 32) | (unsigned long long)(0xffffffffu - (unsigned int)i);
        atomicMax(cells + (int64_t)b * d.H * d.W + cell, key);
        atomicAdd(counts + (int64_t)b * d.H * d.W + cell, 1u);
    }
}

__global__ void __launch_bounds__(256)
bev_finalize_kernel(const float4 *__restrict__ pts, const int64_t *__restrict__ offsets, cy4_bev_desc d,
                    const unsigned long long *__restrict__ cells, const unsigned int *__restrict__ counts, float *__restrict__ out)
{
    const int b = blockIdx.y;
    const int hw = d.H * d.W;
    const float max_height = d.max_height;                                       // :58, float(abs(maxZ - minZ)) from the host
    const double log64 = log(64.0);
    float *o = out + (int64_t)b * 3 * hw;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < hw; c += gridDim.x * blockDim.x) {
        const unsigned int cnt = counts[(int64_t)b * hw + c];
        float inten = 0.f, height = 0.f, dens = 0.f;
        if (cnt) {
            const unsigned long long key = cells[(int64_t)b * hw + c];
            const unsigned int idx = 0xffffffffu - (unsigned int)(key & 0xffffffffu);
            height = __fdiv_rn(from_orderable((unsigned int)(key >> 32)), max_height);        // :59
            inten = __ldg(&pts[offsets[b] + idx].w);                                          // :69
            dens = (float)fmin(1.0, log((double)cnt + 1.0) / log64);                          // :67
        }
        o[c] = inten; o[hw + c] = height; o[2 * hw + c] = dens;                  // RGB_Map[0..2], :71-74
    }
}

}  // namespace cy4

using namespace cy4;

extern "C" {

size_t cy4_bev_workspace_bytes(int B, int H, int W)
{
    return (B > 0 && H > 0 && W > 0) ? (size_t)B * H * W * 12 : 0;
}

int cy4_bev_rasterize(const float *points4, const int64_t *offsets, int B, const cy4_bev_desc *desc, float *out,
                      int32_t *dropped, void *workspace, void *stream)
{
    CY4_CHECK_ARG(desc && B >= 0, "cy4_bev_rasterize: bad arguments");
    if (B == 0) return 0;
    CY4_CHECK_ARG(offsets && out && workspace && dropped, "cy4_bev_rasterize: null pointer");
    CY4_CHECK_ARG(desc->H > 0 && desc->W > 0 && desc->discretization > 0.f && desc->max_height > 0.f, "cy4_bev_rasterize: bad map geometry");
    CY4_CHECK_ARG(((uintptr_t)points4 & 15) == 0, "cy4_bev_rasterize: points must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t hw = (size_t)desc->H * desc->W;
    unsigned long long *cells = (unsigned long long *)workspace;
    unsigned int *counts = (unsigned int *)(cells + (size_t)B * hw);
    CY4_CUDA(cudaMemsetAsync(workspace, 0, (size_t)B * hw * 12, st));
    CY4_CUDA(cudaMemsetAsync(dropped, 0, (size_t)B * 4, st));
    // ~120 k points per KITTI frame: 2 waves of 256-thread blocks over the 148 SMs for the whole batch
    const int bx = std::max(1, std::min(64, (2 * sm_count() * 4 + B - 1) / B));
    bev_scatter_kernel<<<dim3(bx, B), 256, 0, st>>>((const float4 *)points4, offsets, B, *desc, cells, counts, dropped);
    const int fx = std::max(1, std::min((int)((hw + 255) / 256), (4 * sm_count() * 4 + B - 1) / B));
    bev_finalize_kernel<<<dim3(fx, B), 256, 0, st>>>((const float4 *)points4, offsets, *desc, cells, counts, out);
    return cy4_launch_status("cy4_bev_rasterize", 2);
}

}  // extern "C"
