// voxel.cu -- pcl::VoxelGrid<pcl::PointXYZ> centroid filter for the stitched cloud
// (LocalMap::getStitchedPointCloud, src/LocalMap.cpp:438-446; PCL 1.12 filters/impl/voxel_grid.hpp:214-426,
// restated -- PCL is not in this image).  Voxel index arithmetic is PCL's float arithmetic, so the voxel set and
// the output order (ascending voxel index) are identical; inside a voxel PCL sums in the order its unstable sort
// leaves, here the points are summed in (voxel, original index) order, so centroids agree to float rounding.
// The sort is CUB's device radix sort (library code; this is SURVEY 8(f) row 3, not the hot path).
#include "tmb_kernels.h"
#include <cub/cub.cuh>
#include <math_constants.h>

namespace tmb
{
// order-preserving float <-> uint map for atomicMin / atomicMax
__device__ __forceinline__ uint32_t f2o(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__global__ void k_vox_minmax(const float *__restrict__ xyz, uint32_t n, uint32_t *__restrict__ mm /*[6]: min xyz, max xyz (ordered)*/)
{
    float mn[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, mx[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
#pragma unroll
        for (int k = 0; k < 3; ++k)
        {
            const float v = xyz[3 * (size_t)i + k];
            mn[k] = fminf(mn[k], v);
            mx[k] = fmaxf(mx[k], v);
        }
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        for (int o = 16; o > 0; o >>= 1)
        {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
        if ((threadIdx.x & 31) == 0)
        {
            atomicMin(mm + k, f2o(mn[k]));
            atomicMax(mm + 3 + k, f2o(mx[k]));
        }
    }
}

struct VoxGrid
{
    float inv[3];
    int min_b[3];
    int mul[3];
};

// voxel_grid.hpp:342-352: ijk = (int)(floor(p * inverse_leaf) - (float)min_b); idx = ijk . divb_mul
__global__ void k_vox_key(const float *__restrict__ xyz, uint32_t n, VoxGrid vg, uint32_t *__restrict__ key, uint32_t *__restrict__ val)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    int idx = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
        const int ijk = (int)__fsub_rn(floorf(__fmul_rn(xyz[3 * (size_t)i + k], vg.inv[k])), (float)vg.min_b[k]);
        idx += ijk * vg.mul[k];
    }
    key[i] = (uint32_t)idx;
    val[i] = i;
}

__global__ void k_vox_heads(const uint32_t *__restrict__ key, uint32_t n, uint32_t *__restrict__ head)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        head[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

// one thread per voxel run: float sums in sorted order, centroid = sum / count (voxel_grid.hpp:400-421)
__global__ void k_vox_centroid(const float *__restrict__ xyz, const uint32_t *__restrict__ key, const uint32_t *__restrict__ val,
                               const uint32_t *__restrict__ head, const uint32_t *__restrict__ ordinal, uint32_t n,
                               float *__restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i])
        return;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    uint32_t c = 0;
    const uint32_t k0 = key[i];
    for (uint32_t j = i; j < n && key[j] == k0; ++j)
    {
        const size_t p = 3 * (size_t)val[j];
        sx = __fadd_rn(sx, xyz[p]); sy = __fadd_rn(sy, xyz[p + 1]); sz = __fadd_rn(sz, xyz[p + 2]);
        ++c;
    }
    const float fc = (float)c;
    const size_t o = 3 * (size_t)ordinal[i];
    out[o] = __fdiv_rn(sx, fc); out[o + 1] = __fdiv_rn(sy, fc); out[o + 2] = __fdiv_rn(sz, fc);
}

cudaError_t vox_minmax(cudaStream_t st, const float *d_xyz, uint32_t n, uint32_t *d_mm6)
{
    const uint32_t init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    cudaMemcpyAsync(d_mm6, init, sizeof(init), cudaMemcpyHostToDevice, st);
    k_vox_minmax<<<148 * 4, 256, 0, st>>>(d_xyz, n, d_mm6);
    return cudaGetLastError();
}

float vox_ordered_to_float(uint32_t o)
{
    const uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

size_t vox_temp_bytes(uint32_t n)
{
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)n);
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n);
    return a > b ? a : b;
}

// keys/vals: [2][n] ping-pong; head, ordinal: [n]; returns the voxel count through *n_vox_host (synchronises the stream)
cudaError_t vox_filter(cudaStream_t st, const float *d_xyz, uint32_t n, const float inv[3], const int min_b[3], const int mul[3],
                       uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_head, uint32_t *d_ordinal, void *d_temp, size_t temp_bytes,
                       float *d_out, uint32_t *n_vox_host)
{
    VoxGrid vg;
    for (int k = 0; k < 3; ++k) { vg.inv[k] = inv[k]; vg.min_b[k] = min_b[k]; vg.mul[k] = mul[k]; }
    const unsigned blocks = (n + 255) / 256;
    k_vox_key<<<blocks, 256, 0, st>>>(d_xyz, n, vg, d_keys, d_vals);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(d_temp, temp_bytes, d_keys, d_keys + n, d_vals, d_vals + n, (int)n, 0, 32, st);
    if (e != cudaSuccess) return e;
    k_vox_heads<<<blocks, 256, 0, st>>>(d_keys + n, n, d_head);
    e = cub::DeviceScan::ExclusiveSum(d_temp, temp_bytes, d_head, d_ordinal, (int)n, st);
    if (e != cudaSuccess) return e;
    uint32_t last[2];
    cudaMemcpyAsync(&last[0], d_ordinal + (n - 1), 4, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&last[1], d_head + (n - 1), 4, cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return e;
    *n_vox_host = last[0] + last[1];
    k_vox_centroid<<<blocks, 256, 0, st>>>(d_xyz, d_keys + n, d_vals + n, d_head, d_ordinal, n, d_out);
    return cudaGetLastError();
}

}  // namespace tmb
