// tmb_kernels.h -- host-callable launch wrappers implemented in kernels_bin.cu / kernels_classify.cu
#pragma once
#include "tmb_common.cuh"
#include <cuda.h>

namespace tmb
{
cudaError_t launch_prune(cudaStream_t st, const void *d_src, size_t n, size_t stride, const float T[12], float rh,
                         float mx, float mn, uint32_t *d_warp_cnt, uint32_t *d_total, float4 *d_out, int phase,
                         const int field_floats[3] = nullptr, int big_endian = 0);
cudaError_t launch_pack_base(cudaStream_t st, const float *d_xyz, size_t n, float4 *d_out, uint32_t *d_total);
cudaError_t launch_bin(cudaStream_t st, const BatchView &b, const GridView &g, const LatticeD &lat, int ppt, uint32_t maxT,
                       int dc, int di, int n_sm, cudaEvent_t *ev, int phases);
cudaError_t launch_fold_atomic_experiment(cudaStream_t st, const BatchView &b, const GridView &g, const LatticeD &lat, float *d_scratch,
                                          uint32_t *d_mismatch, int n_sm);
cudaError_t launch_shard_merge(cudaStream_t st, const BatchView &b, const GridView &g, const float4 *d_recv, const uint32_t *d_counts_all,
                               const uint32_t *d_starts_all, const unsigned long long *d_src_base, int world, int y0, int y1,
                               uint32_t *d_glob, int dc, int di);

cudaError_t launch_row_totals(cudaStream_t st, const BatchView &b, uint32_t *d_rows);
cudaError_t launch_ops_bbox(cudaStream_t st, const BatchView &b, int32_t *d_out /*[5]: min ci, max ci, min cj, max cj, error flags*/);

size_t classify_smem_bytes(int r);
size_t inflate_smem_bytes(int R, int n_taps);
cudaError_t upload_classify_constants(const int *disc_off, const double *disc_d, int n_disc);
cudaError_t launch_classify(cudaStream_t st, const CUtensorMap &tmap, const CUtensorMap &tmap_flags, const BatchView &b, const GridView &g,
                            const ClassifyParams &cp, int n_blocks, cudaEvent_t ev_after_gate = nullptr);
cudaError_t launch_clear_touched(cudaStream_t st, const BatchView &b, const GridView &g, int n_blocks);
cudaError_t launch_inflate(cudaStream_t st, const CUtensorMap &tmap_step, const BatchView &b, const GridView &g,
                           const ClassifyParams &cp, const void *d_taps, int n_blocks);
cudaError_t launch_collect_keys(cudaStream_t st, const GridView &g, int mode, int clear, unsigned long long *d_out, uint32_t cap,
                                uint32_t *d_count);
cudaError_t launch_read_layers(cudaStream_t st, const GridView &g, const unsigned long long *d_keys, size_t n, const int *d_layers,
                               int n_layers, int kx, int ky, float *d_out);
cudaError_t launch_export_dense(cudaStream_t st, const GridView &g, int layer, int ci0, int cj0, int w, int h, int kx, int ky,
                                float *d_out);
cudaError_t launch_export_occupancy(cudaStream_t st, const GridView &g, int ci0, int cj0, int w, int h, int kx, int ky,
                                    signed char *d_out);
cudaError_t launch_transform_cloud(cudaStream_t st, const float4 *d_cloud, uint32_t n, const float *d_T, float *d_out);
cudaError_t launch_export_completeness(cudaStream_t st, const GridView &g, int ci0, int cj0, int w, int h, int kx, int ky, int f,
                                       int cw, int ch, signed char *d_out);
cudaError_t launch_mark_observed_changed(cudaStream_t st, const GridView &g);
// voxel.cu: pcl::VoxelGrid centroid filter of the stitched cloud
cudaError_t vox_minmax(cudaStream_t st, const float *d_xyz, uint32_t n, uint32_t *d_mm6);
float vox_ordered_to_float(uint32_t o);
size_t vox_temp_bytes(uint32_t n);
cudaError_t vox_filter(cudaStream_t st, const float *d_xyz, uint32_t n, const float inv[3], const int min_b[3], const int mul[3],
                       uint32_t *d_keys, uint32_t *d_vals, uint32_t *d_head, uint32_t *d_ordinal, void *d_temp, size_t temp_bytes,
                       float *d_out, uint32_t *n_vox_host);
}  // namespace tmb
