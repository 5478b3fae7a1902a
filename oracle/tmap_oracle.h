/* tmap_oracle.h -- C API of the CPU ORACLE for the traversability hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it,
 * and there only as the checker / the timed CPU baseline.
 *
 * The oracle is a dependency-free C++17 restatement of the reference's single-threaded CPU
 * path (keyframe cloud -> prune -> pose transform -> cell key -> per-keyframe double moments
 * -> float32 grid layers -> disc-neighbourhood PCA classification -> step inflation ->
 * occupancy export) following, in /root/reference/traversability_mapping:
 *   include/traversability_mapping/Moments.hpp:40-141, src/Moments.cpp:18-58,
 *   src/KeyFrame.cpp:52-67, src/Helpers.cpp:29-184, src/LocalMap.cpp:36-78,91-126,130-241,
 *   245-315,343-414, src/TraversabilityMetrics.cpp:23-124, src/System.cpp:25-43,93-119.
 *
 * PARITY PINNING.  The reference cannot be built here (Eigen, PCL, grid_map_core, yaml-cpp,
 * Boost absent).  The oracle is pinned against every known-answer test the reference's own
 * gtest files hold for this path (oracle/kat.cpp replays test_moments.cpp, test_keyframe.cpp,
 * test_localmap.cpp, test_helpers.cpp).  Arithmetic that lives in un-vendored Eigen
 * (Affine3f*Vector3f op order, SelfAdjointEigenSolver iteration) and grid_map_ros
 * (toOccupancyGrid) is restated from the published algorithms of Eigen 3.4.0 / grid_map 2.x;
 * at the last-ulp level that part is "parity unpinned" (reference tests pin it to 1e-4..1e-2).
 */
#ifndef TMAP_ORACLE_H_
#define TMAP_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The 22 reference parameter keys (src/Parameters.cpp:42-73), same field order as the
 * product's tmap_params so tests can fill both from one dict. */
typedef struct tmo_params
{
    double resolution;          /* grid/resolution_local_map */
    double grid_center_x;       /* grid/grid_center_x */
    double grid_center_y;       /* grid/grid_center_y */
    double half_size;           /* grid/half_size_traversability */
    double extend_length;       /* grid/extend_length_every_resize_by */
    double robot_radius;        /* traversability/robot_radius */
    double ground_clearance;    /* traversability/ground_clearance */
    double max_slope;           /* traversability/max_slope */
    int32_t min_points_per_grid;/* traversability/min_points_per_grid */
    int32_t is_kf_optimization_enabled; /* mapping/is_kf_optimization_enabled */
    int32_t num_local_keyframes;/* mapping/num_local_keyframes (unused by the reference) */
    int32_t global_adjustment_sleep; /* mapping/global_adjustment_sleep */
    int32_t inflation_enabled;  /* inflation/enabled */
    int32_t use_pointcloud_buffer; /* pointcloud/use_pointcloud_buffer */
    int32_t use_ros_buffer;     /* pointcloud/use_ros_buffer */
    int32_t reserved0;
    double inflation_radius;    /* inflation/inflation_radius */
    double cost_scaling_factor; /* inflation/cost_scaling_factor */
    double lethal_step_threshold; /* inflation/lethal_step_threshold */
    double robot_height;        /* pointcloud/robot_height */
    double max_range;           /* pointcloud/max_range_base_frame */
    double min_range;           /* pointcloud/min_range_base_frame */
    double publish_rate_hz;     /* node/publish_rate_hz */
} tmo_params;

/* Layer ids: the reference's 20 layers in LocalMap.cpp:69-71 order, then the z-min / z-max
 * extension (not in the reference; oracle-defined, see tmap_oracle.cpp). */
enum
{
    TMO_L_N = 0, TMO_L_SX, TMO_L_SY, TMO_L_SZ, TMO_L_SX2, TMO_L_SY2, TMO_L_SZ2, TMO_L_SXY,
    TMO_L_SXZ, TMO_L_SYZ, TMO_L_HAZARD, TMO_L_ELEVATION, TMO_L_SLOPE, TMO_L_STEP,
    TMO_L_ROUGHNESS, TMO_L_BORDER, TMO_L_NORMAL_X, TMO_L_NORMAL_Y, TMO_L_NORMAL_Z,
    TMO_L_STEP_INFLATED, TMO_L_ZMIN, TMO_L_ZMAX, TMO_NUM_LAYERS
};

typedef struct tmo_map tmo_map;   /* one LocalMap + the System-level prune/registry */

void tmo_default_params(tmo_params *p);           /* params/traversabilityParams.yaml */
tmo_map *tmo_create(const tmo_params *p);
void tmo_destroy(tmo_map *m);

/* System::setExtrinsicParameters (System.cpp:45-51) with Tbv = Tbs*Tsv already composed. */
void tmo_set_extrinsics(tmo_map *m, const float T_bv[12]);
/* CPU-baseline flavour 2: classify the dirty cells with `n` OpenMP threads, as the archived node does
 * (traversability_mapping_ros/archive/local_traversability_monolithic.cpp:349-362).  1 = the reference's single worker. */
void tmo_set_threads(tmo_map *m, int32_t n);

/* System::prune (System.cpp:93-114).  xyz: n points, `stride_floats` floats apart (3 packed,
 * 4 for pcl::PointXYZ).  out: up to n*3 floats.  Returns the surviving count. */
size_t tmo_prune(const tmo_map *m, const float *xyz, size_t n, size_t stride_floats,
                 const float T[12], double max_range, double min_range, float *out_xyz);

/* System::addNewKeyFrameWithPCL (System.cpp:159-172): prune with Tbv + register.
 * Returns 1 stored, 0 ignored (duplicate / pruned-empty), -1 negative id (reference throws). */
int tmo_add_keyframe(tmo_map *m, uint64_t kf_id, const float *xyz, size_t n, size_t stride_floats);
/* LocalMap::addAlreadyDeclaredKF with an already-pruned base-frame cloud (packed xyz). */
int tmo_add_keyframe_base(tmo_map *m, uint64_t kf_id, const float *xyz, size_t n);
/* System::updateKeyFrame -> LocalMap::setKeyFramePose (latest-wins, dedup enqueue). 1 queued, 0 unknown id */
int tmo_update_pose(tmo_map *m, uint64_t kf_id, const float T_mb[12]);
/* Worker drain (LocalMap.cpp:452-485): one recomputeKeyFrame per queued keyframe, FIFO.
 * Returns the number of grid-changing ops (= onUpdate_ invocations). */
int tmo_process(tmo_map *m);
/* Batched drain: fold every queued keyframe (same float32 op order), then ONE recompute over
 * the union of the dirty sets.  Mirrors the product's max_batch!=1 mode. */
int tmo_process_batched(tmo_map *m);
/* LocalMap::detachKeyFrame / deleteKeyFrame (LocalMap.cpp:343-373). 1 found, 0 unknown */
int tmo_delete_keyframe(tmo_map *m, uint64_t kf_id);
/* LocalMap::clearEntireMap (LocalMap.cpp:375-397); re-enqueue order = ascending kf id. */
void tmo_clear_entire_map(tmo_map *m);

size_t tmo_num_keyframes(const tmo_map *m);
size_t tmo_queue_depth(const tmo_map *m);
/* takeChangedCells / occupiedCellKeys (LocalMap.cpp:401-414). Pass out=NULL to get the count.
 * Keys are returned ascending. take_changed drains only when out!=NULL. */
size_t tmo_take_changed_cells(tmo_map *m, uint64_t *out, size_t cap);
size_t tmo_occupied_cell_keys(const tmo_map *m, uint64_t *out, size_t cap);
/* fillSparseUpdate (traversability_mapping_ros/src/conversions.cpp:28-55): row-major per key;
 * keys outside the grid yield NaN rows here (the reference skips them). */
void tmo_read_layers(const tmo_map *m, const uint64_t *keys, size_t n, const int32_t *layer_ids,
                     size_t n_layers, float *out);
/* Grid geometry: cells per axis = 2*k+1, k = ceil(half/res) (Helpers.cpp:42-47). */
/* fillSparseUpdate (conversions.cpp:28-55): keys outside the grid skipped, order kept; returns the count */
size_t tmo_fill_sparse_update(const tmo_map *m, const int32_t *layer_ids, size_t n_layers, const uint64_t *keys, size_t n,
                              uint64_t *out_keys, float *out_values);
void tmo_grid_extent(const tmo_map *m, int32_t *kx, int32_t *ky);
/* Dense crop of one layer: out[(cj-cj0)*w + (ci-ci0)], NaN outside the grid. */
void tmo_export_dense(const tmo_map *m, int32_t layer, int32_t ci0, int32_t cj0, int32_t w,
                      int32_t h, float *out);
/* grid_map_ros toOccupancyGrid(grid,"hazard",0,1) restated: int8 trunc(100*clamp(v,0,1)),
 * NaN -> -1, same crop addressing as export_dense (full grid: ci0=-kx, cj0=-ky, w=2kx+1). */
void tmo_export_occupancy(const tmo_map *m, int32_t ci0, int32_t cj0, int32_t w, int32_t h,
                          int8_t *out);
/* publishCompleteness (global_traversability_node.cpp:381-459): border_haz occupancy, min-pooled to publish_res,
 * RViz costmap-palette bytes; out holds ceil(w/f) x ceil(h/f) cells, f = lround(publish_res / float(resolution)) */
void tmo_export_completeness(const tmo_map *m, int32_t ci0, int32_t cj0, int32_t w, int32_t h, double publish_res,
                             int8_t *out, int32_t *out_w, int32_t *out_h);
/* LocalMap::getStitchedPointCloud (LocalMap.cpp:416-448): pose * retained clouds, voxel-filtered when all three
 * leaf sizes are positive; packed xyz; out == NULL returns the unfiltered size */
size_t tmo_stitched_cloud(const tmo_map *m, float vx, float vy, float vz, float *out, size_t cap);
/* pcl::VoxelGrid<pcl::PointXYZ> as LocalMap::getStitchedPointCloud applies it (LocalMap.cpp:438-446): packed xyz in,
 * one centroid per occupied voxel out (ascending voxel index); returns the voxel count */
size_t tmo_voxel_grid(const float *xyz, size_t n, float lx, float ly, float lz, float *out, size_t cap);

/* ---- stateless pieces, exposed for the known-answer tests ---- */
/* Lattice::cellOf / key (Moments.hpp:105-133) */
void tmo_cell_of(double x0, double y0, double res, double x, double y, int32_t *ci, int32_t *cj);
uint64_t tmo_key(int32_t ci, int32_t cj);
/* discOffsets (Helpers.cpp:156-169): returns count; out gets (di,dj) pairs */
size_t tmo_disc_offsets(double radius, double res, int32_t *out_pairs, size_t cap_pairs);
/* KeyFrame::rebin (KeyFrame.cpp:52-67) on a packed base cloud: returns #cells; arrays sized n.
 * moments: 10 doubles per cell (N, sx, sy, sz, sx2, sy2, sz2, sxy, sxz, syz), keys ascending */
size_t tmo_rebin(double x0, double y0, double res, const float *xyz, size_t n, const float T[12],
                 uint64_t *out_keys, double *out_moments);
/* computeGoodness (TraversabilityMetrics.cpp:23-124).  cells: n_occ entries of 13 doubles
 * (N, sx..syz, cx, cy, cz); query likewise.  out: 9 doubles in Hazard enum order. */
void tmo_compute_goodness(const double *query, const double *cells, size_t n_occ,
                          int32_t vicinity_cell_count, double ground_clearance, double max_slope,
                          uint32_t min_points, double *out9);
/* symmetric 3x3 eigen-decomposition as Eigen 3.4 SelfAdjointEigenSolver::compute:
 * a = row-major 3x3 (lower triangle read); w ascending; v column-major eigenvectors. */
void tmo_eigh3(const double a[9], double w[3], double v[9]);

#ifdef __cplusplus
}
#endif
#endif /* TMAP_ORACLE_H_ */
