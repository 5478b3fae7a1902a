/* tmap.h -- C ABI of the B200-native traversability hot path (libtmap_b200.so).
 *
 * Drop-in boundary for ONE path of suchetanrs/traversability_mapping: keyframe insert ->
 * pose transform -> grid binning -> per-cell neighbourhood classification -> occupancy export,
 * plus pose update / delete / loop-closure re-integration.  Every entry point names the
 * reference interface it replaces (paths relative to the reference's traversability_mapping/).
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary;
 *   - poses are row-major 3x4 float matrices [R|t] (the first three rows of Eigen::Affine3f);
 *   - cell keys are the reference's packed ids, key = (uint32(ci) << 32) | uint32(cj)
 *     (include/traversability_mapping/Moments.hpp:122-133);
 *   - return value: 0 = done, >0 = the call was IGNORED exactly where the reference logs to
 *     std::cerr and carries on, <0 = error (where the reference throws, or a CUDA failure);
 *     tmap_last_error() returns the message of the calling thread's last failure;
 *   - there is NO CPU fallback: without a CUDA device tmap_create fails with TMAP_ERR_NO_DEVICE.
 *   - threading: any thread may call any function; one worker thread per map owns that map's
 *     CUDA stream and fires the on_update callback (LocalMap.cpp:452-485).
 */
#ifndef TMAP_H_
#define TMAP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum
{
    TMAP_OK = 0,
    TMAP_IGNORED_DUPLICATE = 1,    /* System.cpp:125-129, :65-69 */
    TMAP_IGNORED_UNKNOWN_MAP = 2,  /* System.cpp:130-136, :259-263 */
    TMAP_IGNORED_EMPTY = 3,        /* System.cpp:137-141 (pruned to empty) */
    TMAP_IGNORED_UNKNOWN_KF = 4,   /* System.cpp:203-207, :238-242 */
    TMAP_IGNORED_NULL = 5,         /* System.cpp:169-170 (null cloud) */
    TMAP_IGNORED_NO_OBSTACLE_PARAMS = 6, /* System.cpp:304-308 */
    TMAP_ERR_NEGATIVE_ID = -1,     /* System.cpp:163-168 throws std::runtime_error */
    TMAP_ERR_CUDA = -2,
    TMAP_ERR_INVALID = -3,         /* null handle, non-finite pose, bad layer id ... */
    TMAP_ERR_UNSUPPORTED = -4,     /* footprint / sensor range beyond the kernels' tile limits */
    TMAP_ERR_NO_DEVICE = -5
};

/* The reference's 22 parameter keys (src/Parameters.cpp:42-73; defaults
 * params/traversabilityParams.yaml) followed by the runtime knobs of this implementation. */
typedef struct tmap_params
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
    int32_t num_local_keyframes;/* mapping/num_local_keyframes (unused by the reference too) */
    int32_t global_adjustment_sleep; /* mapping/global_adjustment_sleep (unused: no backstop thread) */
    int32_t inflation_enabled;  /* inflation/enabled */
    int32_t use_pointcloud_buffer; /* pointcloud/use_pointcloud_buffer: enables tmap_push_to_buffer / tmap_add_keyframe_ts */
    int32_t use_ros_buffer;     /* pointcloud/use_ros_buffer */
    int32_t reserved0;
    double inflation_radius;    /* inflation/inflation_radius */
    double cost_scaling_factor; /* inflation/cost_scaling_factor */
    double lethal_step_threshold; /* inflation/lethal_step_threshold */
    double robot_height;        /* pointcloud/robot_height */
    double max_range;           /* pointcloud/max_range_base_frame */
    double min_range;           /* pointcloud/min_range_base_frame */
    double publish_rate_hz;     /* node/publish_rate_hz (callers' concern) */
    /* ---- runtime knobs (not in the reference) ---- */
    int32_t device;             /* CUDA device ordinal */
    int32_t max_batch;          /* keyframes folded per worker pass: 1 = the reference's
                                   one-recompute-per-keyframe sequence (bit-for-bit the same
                                   derived layers); 0 = everything pending in one pass (same
                                   moment layers; derived layers recomputed once at the end) */
    int32_t profile;            /* 1: record per-stage CUDA-event times into tmap_stats */
    int32_t reserved1;
} tmap_params;

/* Layer ids: the reference's 20 grid_map layers in LocalMap.cpp:69-71 order, then the
 * z-min / z-max extension layers (not in the reference). */
enum
{
    TMAP_L_N = 0, TMAP_L_SX, TMAP_L_SY, TMAP_L_SZ, TMAP_L_SX2, TMAP_L_SY2, TMAP_L_SZ2,
    TMAP_L_SXY, TMAP_L_SXZ, TMAP_L_SYZ, TMAP_L_HAZARD, TMAP_L_ELEVATION, TMAP_L_SLOPE,
    TMAP_L_STEP, TMAP_L_ROUGHNESS, TMAP_L_BORDER, TMAP_L_NORMAL_X, TMAP_L_NORMAL_Y,
    TMAP_L_NORMAL_Z, TMAP_L_STEP_INFLATED, TMAP_L_ZMIN, TMAP_L_ZMAX, TMAP_NUM_LAYERS
};

typedef struct tmap_stats
{
    uint64_t keyframes_processed;  /* recomputeKeyFrame passes that changed the grid */
    uint64_t points_binned;        /* points through transform+bin (subtractions included) */
    uint64_t cells_classified;     /* dirty observed cells recomputed */
    uint64_t batches;              /* worker passes */
    uint64_t kernel_launches;      /* CUDA kernels launched by this library */
    uint64_t grid_reallocs;
    /* CUDA-event times of the LAST batch, ms (0 unless params.profile) */
    float ms_hist, ms_scan, ms_scatter, ms_fold, ms_classify, ms_inflate, ms_total;
    float host_ms_batch;           /* wall-clock of the last batch on the host (assembly + launches + sync) */
    uint64_t last_batch_points, last_batch_tiles, last_batch_cells_touched,
        last_batch_cells_classified;
    uint64_t last_batch_max_bin;   /* records in the heaviest 16x16 bin of the last batch */
    /* running sums over all batches (same CUDA-event times; 0 unless params.profile) */
    double sum_ms_hist, sum_ms_scan, sum_ms_scatter, sum_ms_fold, sum_ms_classify, sum_ms_inflate, sum_ms_total;
    uint64_t sum_tiles, sum_cells_touched;
    /* host wall-clock per batch, summed: op descriptors + buffers + the meta upload / kernel launches /
     * waiting for the device (result copy + stream synchronize) / the whole batch */
    double sum_host_ms_plan, sum_host_ms_launch, sum_host_ms_wait, sum_host_ms_batch;
    /* the classify stage is two kernels: ms_gate is k_gate's share of ms_classify (last batch / running sum) */
    double sum_ms_gate;
    float ms_gate, reserved0;
    uint64_t sum_infl_tiles;       /* 32x32 tiles handed to k_inflate */
    uint64_t sum_cand_bins;        /* 16x16 bins handed to k_gate */
    uint64_t h2d_bytes, d2h_bytes; /* bytes this library copied host->device / device->host (all maps, ingest included) */
} tmap_stats;

typedef struct tmap_system tmap_system;
typedef void (*tmap_update_cb)(void *user);

/* params/traversabilityParams.yaml defaults; device 0, max_batch 1, profile 0 */
void tmap_default_params(tmap_params *p);
const char *tmap_last_error(void);

/* ParameterHandler::getInstance + System::System (System.cpp:25-43) */
int tmap_create(const tmap_params *p, tmap_system **out);
void tmap_destroy(tmap_system *s);

/* System::setExtrinsicParameters(Tsv, Tbs) (System.hpp:60, System.cpp:45-51): Tbv = Tbs*Tsv */
int tmap_set_extrinsics(tmap_system *s, const float T_sv[12], const float T_bs[12]);
/* System::addNewLocalMap(mapID, onUpdate) (System.hpp:76, System.cpp:62-72).  on_update fires on
 * the map's worker thread after the grid lock is released, once per grid-changing keyframe. */
int tmap_add_map(tmap_system *s, uint64_t map_id, tmap_update_cb on_update, void *user);
/* System::setCurrentMap (System.cpp:74-82) */
int tmap_set_current_map(tmap_system *s, uint64_t map_id);
/* System::getLocalMap() (System.hpp:154, System.cpp:357-361): id of the ACTIVE map -- the last one created by
 * tmap_add_map (System.cpp:71) or chosen by tmap_set_current_map; TMAP_IGNORED_UNKNOWN_MAP before any map exists. */
int tmap_current_map(tmap_system *s, uint64_t *map_id_out);

/* System::addNewKeyFrameWithPCL (System.hpp:85-87, System.cpp:159-172).  xyz: sensor-frame
 * points in HOST memory, `stride_bytes` apart (16 for pcl::PointXYZ, 12 packed).  The cloud is
 * copied to the device and pruned there (System::prune, System.cpp:93-114); it is NOT binned
 * until the first tmap_update_pose.  Synchronous, like the reference. */
int tmap_add_keyframe(tmap_system *s, uint64_t timestamp_ns, uint64_t kf_id, uint64_t map_id,
                      const float *xyz, size_t n, size_t stride_bytes);
/* tmap_ros::toPCL + addNewKeyFrameWithPCL (traversability_mapping_ros/src/conversions.cpp:19-26,
 * global_traversability_node.cpp:171-182): the raw `data` bytes of a sensor_msgs/PointCloud2 (width*height points,
 * `point_step` apart) with the byte offsets of its FLOAT32 x / y / z fields; decoded and pruned on the device, no
 * per-point host loop.  Offsets must be 4-byte aligned inside point_step (else TMAP_ERR_UNSUPPORTED). */
int tmap_add_keyframe_pc2(tmap_system *s, uint64_t timestamp_ns, uint64_t kf_id, uint64_t map_id, const void *data,
                          size_t n_points, size_t point_step, uint32_t x_offset, uint32_t y_offset, uint32_t z_offset,
                          int is_bigendian);
/* System::pushToBuffer(timestamp, cloud) (System.hpp, System.cpp:339-344) + PointCloudBuffer::addPointCloud
 * (PointCloudBuffer.cpp:26-32): store a sensor-frame cloud under its acquisition time (seconds).  The copy lives in
 * DEVICE memory; when the buffered span exceeds 25 s everything older than timestamp - 5 s is dropped.  A no-op
 * unless params.use_pointcloud_buffer (the reference creates no buffer otherwise, System.cpp:38-42). */
int tmap_push_to_buffer(tmap_system *s, double timestamp_s, const float *xyz, size_t n, size_t stride_bytes);
/* System::addNewKeyFrameTsDouble (System.cpp:174-186; addNewKeyFrameTsULong = the same with ns * 1e-9): the
 * keyframe's cloud is the buffered one closest in time (the first such on a tie, PointCloudBuffer.cpp:34-46);
 * empty / absent buffer -> TMAP_IGNORED_NULL.  Pruned on the device without another host copy. */
int tmap_add_keyframe_ts(tmap_system *s, double timestamp_s, uint64_t kf_id, uint64_t map_id);
int tmap_buffered_clouds(tmap_system *s, size_t *n_out);
/* Same, with the cloud already in DEVICE memory (float4-strided or packed). */
int tmap_add_keyframe_device(tmap_system *s, uint64_t timestamp_ns, uint64_t kf_id, uint64_t map_id,
                             const void *d_xyz, size_t n, size_t stride_bytes);
/* KeyFrame ctor + LocalMap::addAlreadyDeclaredKF (LocalMap.cpp:319-328): register an already
 * pruned BASE-frame cloud (host memory, packed xyz) without the prune step. */
int tmap_add_keyframe_base(tmap_system *s, uint64_t kf_id, uint64_t map_id, const float *xyz, size_t n);

/* System::updateKeyFrame(kfID, Affine3f pose_map_base) (System.hpp:101-105, System.cpp:195-224):
 * O(1), latest-wins, never waits on a recompute; the first call after insert triggers the bin. */
int tmap_update_pose(tmap_system *s, uint64_t kf_id, const float T_mb[12]);
/* GlobalTraversabilityNode::updatesCallback (traversability_mapping_ros/src/global_traversability_node.cpp:184-188):
 * one PGO message = System::updateKeyFrame for every keyframe it lists.  Same as n tmap_update_pose calls, in
 * order (poses12: n x 12 floats); returns the first negative code, else TMAP_OK (unknown ids are skipped). */
int tmap_update_poses(tmap_system *s, const uint64_t *kf_ids, const float *poses12, size_t n);
/* System::deleteKeyFrame (System.hpp:112, System.cpp:234-248): synchronous subtract+recompute */
int tmap_delete_keyframe(tmap_system *s, uint64_t kf_id);
/* System::updateKFMap (System.hpp:115, System.cpp:250-276) */
int tmap_update_kf_map(tmap_system *s, uint64_t kf_id, uint64_t map_id);
/* System::informLoopClosure (System.hpp:117, System.cpp:278-284): map 0 is blanked and every
 * retained keyframe re-integrated at its latest pose (ascending kf id) by the worker. */
int tmap_inform_loop_closure(tmap_system *s);

/* System::setObstacleParameters / addObstacleKeyFrame / deleteObstacleKeyFrame
 * (System.hpp:67,132-136, System.cpp:53-60,288-335).  *id_out = 0 when dropped. */
int tmap_set_obstacle_params(tmap_system *s, const float T_b_obs[12], double max_range, double min_range);
int tmap_add_obstacle_keyframe(tmap_system *s, uint64_t timestamp_ns, const float *xyz, size_t n,
                               size_t stride_bytes, uint64_t *id_out);
int tmap_delete_obstacle_keyframe(tmap_system *s, uint64_t kf_id);

/* Block until every map's work queue is drained (test hook; the reference's tests wait on the
 * onUpdate condvar instead, test_localmap.cpp:106-111). */
int tmap_flush(tmap_system *s);

/* ---- map access: System::getLocalMap(mapID) then LocalMap::* (LocalMap.hpp:108-121) ---- */
/* LocalMap::getGridMapMutex(): hold across a take/occupied + read sequence for consistency */
int tmap_lock(tmap_system *s, uint64_t map_id);
int tmap_unlock(tmap_system *s, uint64_t map_id);
/* LocalMap::takeChangedCells (LocalMap.cpp:401-408).  out==NULL: *n_out = pending count, nothing
 * drained.  Otherwise all pending keys (ascending) are written and the set is cleared; when cap is smaller than
 * the pending count NOTHING is written or cleared and *n_out reports the count (call again with room). */
int tmap_take_changed_cells(tmap_system *s, uint64_t map_id, uint64_t *out, size_t cap, size_t *n_out);
/* LocalMap::occupiedCellKeys (LocalMap.cpp:410-414), ascending */
int tmap_occupied_cell_keys(tmap_system *s, uint64_t map_id, uint64_t *out, size_t cap, size_t *n_out);
/* fillSparseUpdate (traversability_mapping_ros/src/conversions.cpp:28-55): out[i*n_layers+l];
 * keys outside the grid give NaN rows. */
int tmap_read_layers(tmap_system *s, uint64_t map_id, const uint64_t *keys, size_t n,
                     const int32_t *layer_ids, size_t n_layers, float *out);
/* tmap_ros::fillSparseUpdate (traversability_mapping_ros/src/conversions.cpp:28-55), the payload of
 * traversability_msgs/TraversabilitySparseUpdate: keys whose cell lies outside the grid are skipped, the others keep
 * their order; out_values[i * n_layers + l].  out_keys / out_values hold n_keys entries; *n_out = kept keys. */
int tmap_fill_sparse_update(tmap_system *s, uint64_t map_id, const int32_t *layer_ids, size_t n_layers, const uint64_t *keys,
                            size_t n_keys, uint64_t *out_keys, float *out_values, size_t *n_out);
/* grid_map geometry the reference would have: cells per axis = 2*k+1 (Helpers.cpp:35-84) */
int tmap_grid_extent(tmap_system *s, uint64_t map_id, int32_t *kx, int32_t *ky);
/* LocalMap::getGridMap()[layer] crop: out[(cj-cj0)*w + (ci-ci0)], NaN outside the grid */
int tmap_export_dense(tmap_system *s, uint64_t map_id, int32_t layer, int32_t ci0, int32_t cj0,
                      int32_t w, int32_t h, float *out);
/* grid_map::GridMapRosConverter::toOccupancyGrid(grid,"hazard",0,1,.) (call sites
 * global_traversability_node.cpp:374, local_traversability_node.cpp:236): int8 in [0,100], -1
 * unknown; out[(cj-cj0)*w + (ci-ci0)] (nav_msgs row-major, x fastest) */
int tmap_export_occupancy(tmap_system *s, uint64_t map_id, int32_t ci0, int32_t cj0, int32_t w,
                          int32_t h, int8_t *out);
/* GlobalTraversabilityNode::publishCompleteness (traversability_mapping_ros/src/global_traversability_node.cpp:381-459):
 * toOccupancyGrid(grid, "border_haz", 0, 1) over the w x h crop at (ci0, cj0), min-pooled to `publish_res` (0.25 m in
 * the reference; f = lround(publish_res / float(resolution)), unobserved fine cells skipped), then mapped to the RViz
 * costmap-palette bytes (complete -> 113, partial -> 128..254, unobserved -> -1).  out: ceil(w/f) x ceil(h/f) bytes,
 * row-major, x fastest; out == NULL returns the dimensions only. */
int tmap_export_completeness(tmap_system *s, uint64_t map_id, int32_t ci0, int32_t cj0, int32_t w, int32_t h,
                             double publish_res, int8_t *out, size_t cap_bytes, int32_t *out_w, int32_t *out_h);
/* LocalMap::getKeyFramesMap().size() and the work-queue depth (LocalMap.cpp:339-340) */
int tmap_num_keyframes(tmap_system *s, uint64_t map_id, size_t *n_out);
int tmap_queue_depth(tmap_system *s, uint64_t map_id, size_t *n_out);
/* KeyFrame::cloudBase().size() (KeyFrame.hpp:85): points of a keyframe's retained, pruned base-frame cloud
 * (0 once dropped); TMAP_IGNORED_UNKNOWN_KF for an unknown id */
int tmap_keyframe_points(tmap_system *s, uint64_t kf_id, size_t *n_out);
/* LocalMap::getStitchedPointCloud without the voxel filter (LocalMap.cpp:416-437): pose * cloud
 * of every retained keyframe, packed xyz, ascending kf id.  out==NULL: count only. */
int tmap_stitched_cloud(tmap_system *s, uint64_t map_id, float *out_xyz, size_t cap_points, size_t *n_out);
/* LocalMap::getStitchedPointCloud(vx, vy, vz) with positive leaf sizes (LocalMap.cpp:416-448): the stitched cloud
 * through pcl::VoxelGrid<pcl::PointXYZ> (PCL 1.12 filters/impl/voxel_grid.hpp, restated): one centroid per occupied
 * voxel, ascending voxel index.  out == NULL: *n_out = an upper bound (the unfiltered size) for sizing the buffer;
 * otherwise *n_out = the number of voxels (at most cap_points are written).  Non-positive leaves = unfiltered. */
int tmap_stitched_cloud_voxel(tmap_system *s, uint64_t map_id, float leaf_x, float leaf_y, float leaf_z, float *out_xyz,
                              size_t cap_points, size_t *n_out);

/* System::getGlobalPointCloud(vx, vy, vz) (System.hpp:160-161, System.cpp:370-387): the stitched (and, for positive
 * leaves, voxel-filtered) clouds of every map, concatenated in ascending map id.  Same sizing protocol as
 * tmap_stitched_cloud_voxel. */
int tmap_global_point_cloud(tmap_system *s, float leaf_x, float leaf_y, float leaf_z, float *out_xyz, size_t cap_points,
                            size_t *n_out);

int tmap_get_stats(tmap_system *s, uint64_t map_id, tmap_stats *out);
/* Run the map's kernels on a caller-owned stream (cudaStream_t as void*); NULL = own stream. */
int tmap_set_stream(tmap_system *s, uint64_t map_id, void *cuda_stream);

/* ---- sharded global map: one process per GPU, bins owned in slabs of bin rows ----------
 * The reference has no distributed path.  A multi-GPU host drives the three collective entry points below: */

/* ---- the sharded rebuild as ONE collective call (host code C++, collectives issued by the library on the map's stream) ----
 * Bootstrap as with MPI: one rank obtains a 128-byte NCCL unique id (tmap_shard_unique_id), the caller ships it to every
 * rank over whatever channel it has, and every rank calls tmap_shard_comm_init on ITS map (one process per GPU).
 * tmap_shard_rebuild is then collective over the job: rank k passes the contiguous block of keyframe ids it holds (all ids
 * of rank k below those of rank k+1) with their poses.  Steps: all-gather of windows -> shared bbox and op numbering;
 * transform + key + bucket sort of the local keyframes; all-gather of per-row totals -> slabs of bin rows balanced by
 * record count; grouped ncclSend/ncclRecv all-to-all of the bucketed records (the bucket array is the send buffer);
 * per-bin merge in rank order + fold; in-place halo exchange of r cell rows (48-byte moment records + flag bytes) with
 * the neighbouring slabs; classification of the slab.  The rank's grid window covers ITS slab rows + halo over the bbox
 * columns only.  Replaces LocalMap::clearEntireMap + re-integration (LocalMap.cpp:375-397) for a map partitioned over
 * GPUs; the moment and derived layers are bit-identical to the single-GPU rebuild.  NCCL is loaded with dlopen at the
 * first call (TMAP_ERR_UNSUPPORTED if libnccl.so.2 is absent). */
typedef struct tmap_shard_info
{
    int32_t bbox[4];           /* btx0, bty0, btw, bth: the job-wide bbox in 16-cell bins */
    int32_t slab_y0, slab_y1;  /* this rank's bin rows [y0, y1), bbox-local */
    int32_t cj0, cj1;          /* the same as absolute cell rows */
    uint64_t local_records, sent_records, recv_records;
    uint64_t halo_bytes, grid_bytes;
    float ms[8];               /* CUDA-event stage times: bin, plan, all-to-all, merge+fold, halo, classify, total, - */
    int32_t n_slabs;
    int32_t slabs[2 * 64];     /* every rank's [y0, y1) */
} tmap_shard_info;
int tmap_shard_unique_id(uint8_t out_id[128]);
int tmap_shard_comm_init(tmap_system *s, uint64_t map_id, int32_t rank, int32_t world, const uint8_t id[128]);
int tmap_shard_comm_destroy(tmap_system *s, uint64_t map_id);
int tmap_shard_rebuild(tmap_system *s, uint64_t map_id, const uint64_t *kf_ids, size_t n, const float *poses12,
                       tmap_shard_info *out);
/* Collective pose update on a sharded map: the listed keyframes of this rank (ascending ids; n may be 0) move to new poses.
 * One batch over the job = System::updateKeyFrame for each of them (System.cpp:216-232 -> LocalMap::recomputeKeyFrame,
 * LocalMap.cpp:245-315): each keyframe is subtracted at the pose it is integrated at and added at the new one, records are
 * routed to the owners of the LAST rebuild's slabs (cells keep their owners), nothing is blanked and only the dirty cells
 * (touched, dilated by the footprint, across slab boundaries through the halo rows) are re-classified.  Bit-identical to the
 * same batch on one GPU (max_batch = 0). */
int tmap_shard_update(tmap_system *s, uint64_t map_id, const uint64_t *kf_ids, size_t n, const float *poses12,
                      tmap_shard_info *out);
/* the slab planner alone (pure host function; rows_all[s * n_rows + y] = records rank s holds in bin row y) */
int tmap_shard_plan(const uint32_t *rows_all, int32_t world, int32_t n_rows, int32_t first_global_row, int32_t rank,
                    int32_t *slabs /*[2*world]*/, uint64_t *send /*[world]*/, uint64_t *recv /*[world]*/,
                    uint64_t *src_base /*[world]*/);

#ifdef __cplusplus
}
#endif
#endif /* TMAP_H_ */
