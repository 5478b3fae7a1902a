// tmap_system.hpp -- header-only C++ mirror of the reference's System / LocalMap classes on top of the
// C ABI (tmap.h).  Same method names, argument meaning and error behaviour as
// traversability_mapping/include/traversability_mapping/System.hpp:53-212 and LocalMap.hpp:55-231, so a
// SLAM integration that used the reference library can switch by changing the include and the link line.
//
// The Eigen / PCL overloads are compiled only when those headers are present (they are not in this
// image); the plain-pointer overloads always are.
#ifndef TMAP_SYSTEM_HPP_
#define TMAP_SYSTEM_HPP_

#include "tmap.h"

#include <cstdint>
#include <array>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#if __has_include(<Eigen/Geometry>)
#include <Eigen/Geometry>
#define TMAP_HAVE_EIGEN 1
#endif
#if __has_include(<pcl/point_cloud.h>) && __has_include(<pcl/point_types.h>)
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#define TMAP_HAVE_PCL 1
#endif

namespace traversability_mapping_b200
{
/// Row-major 3x4 [R|t]: the first three rows of an Eigen::Affine3f.
struct Pose
{
    float m[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    static Pose Identity() { return Pose(); }
#ifdef TMAP_HAVE_EIGEN
    Pose() = default;
    Pose(const Eigen::Affine3f &T)
    {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 4; ++j) m[4 * i + j] = T.matrix()(i, j);
    }
    Pose(const Eigen::Affine3d &T) : Pose(Eigen::Affine3f(T.cast<float>())) {}  // System.cpp:229
#endif
};

/// Lattice (Moments.hpp:99-134)
struct Lattice
{
    double x0 = 0., y0 = 0., res = 0.25;
    static inline std::uint64_t key(int ci, int cj)
    {
        return (static_cast<std::uint64_t>(static_cast<std::uint32_t>(ci)) << 32) |
               static_cast<std::uint64_t>(static_cast<std::uint32_t>(cj));
    }
    static inline void unkey(std::uint64_t k, int &ci, int &cj)
    {
        ci = static_cast<int>(static_cast<std::uint32_t>(k >> 32));
        cj = static_cast<int>(static_cast<std::uint32_t>(k & 0xffffffffu));
    }
};

class System;

/// Reader handle (LocalMap.hpp:108-121).  Hold lock()/unlock() (getGridMapMutex) across a
/// takeChangedCells()/occupiedCellKeys() + readLayers() sequence for a consistent snapshot.
class LocalMap
{
  public:
    void lock() { tmap_lock(s_, id_); }
    void unlock() { tmap_unlock(s_, id_); }
    const Lattice &getLattice() const { return lat_; }
    std::vector<std::uint64_t> takeChangedCells() { return keys(tmap_take_changed_cells); }
    std::vector<std::uint64_t> occupiedCellKeys() { return keys(tmap_occupied_cell_keys); }
    /// fillSparseUpdate (conversions.cpp:28-55): values[i*layers.size()+l]
    std::vector<float> readLayers(const std::vector<std::uint64_t> &k, const std::vector<int32_t> &layers)
    {
        std::vector<float> out(k.size() * layers.size());
        check(tmap_read_layers(s_, id_, k.data(), k.size(), layers.data(), layers.size(), out.data()));
        return out;
    }
    /// fillSparseUpdate with the in-grid filter (conversions.cpp:43-47): `k` is replaced by the kept keys
    std::vector<float> fillSparseUpdate(std::vector<std::uint64_t> &k, const std::vector<int32_t> &layers)
    {
        std::vector<std::uint64_t> kept(k.size());
        std::vector<float> vals(k.size() * layers.size());
        size_t n = 0;
        check(tmap_fill_sparse_update(s_, id_, layers.data(), layers.size(), k.data(), k.size(), kept.data(), vals.data(), &n));
        kept.resize(n);
        vals.resize(n * layers.size());
        k.swap(kept);
        return vals;
    }
    void gridExtent(int32_t &kx, int32_t &ky) { check(tmap_grid_extent(s_, id_, &kx, &ky)); }
    /// toOccupancyGrid(grid, "hazard", 0, 1, og): whole grid, nav_msgs ordering
    std::vector<int8_t> occupancy(int32_t &width, int32_t &height)
    {
        int32_t kx, ky;
        gridExtent(kx, ky);
        width = 2 * kx + 1;
        height = 2 * ky + 1;
        std::vector<int8_t> out(static_cast<size_t>(width) * height);
        check(tmap_export_occupancy(s_, id_, -kx, -ky, width, height, out.data()));
        return out;
    }
    /// publishCompleteness payload (global_traversability_node.cpp:381-459): whole grid, palette bytes
    std::vector<int8_t> completeness(double publish_res, int32_t &width, int32_t &height)
    {
        int32_t kx, ky;
        gridExtent(kx, ky);
        check(tmap_export_completeness(s_, id_, -kx, -ky, 2 * kx + 1, 2 * ky + 1, publish_res, nullptr, 0, &width, &height));
        std::vector<int8_t> out(static_cast<size_t>(width) * height);
        check(tmap_export_completeness(s_, id_, -kx, -ky, 2 * kx + 1, 2 * ky + 1, publish_res, out.data(), out.size(), &width, &height));
        return out;
    }
    std::vector<float> layer(int32_t layer_id, int32_t ci0, int32_t cj0, int32_t w, int32_t h)
    {
        std::vector<float> out(static_cast<size_t>(w) * h);
        check(tmap_export_dense(s_, id_, layer_id, ci0, cj0, w, h, out.data()));
        return out;
    }
    /// getStitchedPointCloud(vx, vy, vz) (LocalMap.cpp:416-448): packed xyz; positive leaves -> pcl::VoxelGrid centroids
    std::vector<float> getStitchedPointCloud(float vx = 0.f, float vy = 0.f, float vz = 0.f)
    {
        size_t n = 0;
        check(tmap_stitched_cloud_voxel(s_, id_, vx, vy, vz, nullptr, 0, &n));
        std::vector<float> out(3 * n);
        if (n) check(tmap_stitched_cloud_voxel(s_, id_, vx, vy, vz, out.data(), n, &n));
        out.resize(3 * n);
        return out;
    }
    size_t numKeyFrames() { size_t n = 0; check(tmap_num_keyframes(s_, id_, &n)); return n; }

  private:
    friend class System;
    LocalMap(tmap_system *s, std::uint64_t id, const Lattice &l) : s_(s), id_(id), lat_(l) {}
    static void check(int rc) { if (rc < 0) throw std::runtime_error(tmap_last_error()); }
    struct Guard  // the grid mutex is released on every path out of keys(), also when check() throws
    {
        LocalMap &m;
        explicit Guard(LocalMap &mm) : m(mm) { m.lock(); }
        ~Guard() { m.unlock(); }
    };
    template <class F> std::vector<std::uint64_t> keys(F fn)
    {
        size_t n = 0;
        Guard g(*this);
        check(fn(s_, id_, nullptr, 0, &n));
        std::vector<std::uint64_t> out(n);
        if (n) check(fn(s_, id_, out.data(), n, &n));
        out.resize(n);
        return out;
    }
    tmap_system *s_;
    std::uint64_t id_;
    Lattice lat_;
};

class System
{
  public:
    /// ParameterHandler::getInstance(path) + System() (System.cpp:25-43): pass the 22 keys as a struct.
    explicit System(const tmap_params &p) : params_(p)
    {
        if (tmap_create(&p, &s_) != TMAP_OK) throw std::runtime_error(tmap_last_error());
    }
    ~System() { tmap_destroy(s_); }
    System(const System &) = delete;
    System &operator=(const System &) = delete;

    void setExtrinsicParameters(const Pose &Tsv, const Pose &Tbs) { check(tmap_set_extrinsics(s_, Tsv.m, Tbs.m)); }
    void setObstacleParameters(const Pose &Tb_obs, double maxRange, double minRange)
    {
        check(tmap_set_obstacle_params(s_, Tb_obs.m, maxRange, minRange));
    }
    /// System.cpp:62-72.  A duplicate id is a no-op that keeps the FIRST callback (System.cpp:65-69): the new
    /// std::function is registered only once the library has accepted the map, so the library never holds a
    /// pointer to a callback this object has freed.
    void addNewLocalMap(std::uint64_t mapID, std::function<void()> onUpdate = {})
    {
        std::unique_ptr<std::function<void()>> slot(new std::function<void()>(std::move(onUpdate)));
        const int rc = tmap_add_map(s_, mapID, *slot ? &System::trampoline : nullptr, slot.get());
        check(rc);
        if (rc == TMAP_OK) callbacks_[mapID] = std::move(slot);  // the new map is now the active one (System.cpp:71)
    }
    void setCurrentMap(std::uint64_t mapID) { check(tmap_set_current_map(s_, mapID)); }

    /// addNewKeyFrameWithPCL (System.cpp:159-172).  Throws std::runtime_error on a negative id.
    void addNewKeyFrame(unsigned long long timestamp_ns, std::uint64_t kfID, std::uint64_t mapID, const float *xyz, size_t n,
                        size_t stride_bytes)
    {
        const int rc = tmap_add_keyframe(s_, timestamp_ns, kfID, mapID, xyz, n, stride_bytes);
        if (rc == TMAP_ERR_NEGATIVE_ID)
            throw std::runtime_error("The given KF ID was negative. Please make sure the keyFrame IDs are positive. KF ID: " +
                                     std::to_string(static_cast<std::int64_t>(kfID)));
        check(rc);
    }
    /// sensor_msgs/PointCloud2 ingest (conversions.cpp:19-26 + System.cpp:159-172): raw data bytes, point_step and the
    /// byte offsets of the FLOAT32 x / y / z fields; decoded and pruned on the device.
    void addNewKeyFramePointCloud2(unsigned long long timestamp_ns, std::uint64_t kfID, std::uint64_t mapID, const void *data,
                                   size_t n_points, size_t point_step, std::uint32_t x_offset, std::uint32_t y_offset,
                                   std::uint32_t z_offset, bool is_bigendian = false)
    {
        const int rc = tmap_add_keyframe_pc2(s_, timestamp_ns, kfID, mapID, data, n_points, point_step, x_offset, y_offset,
                                             z_offset, is_bigendian ? 1 : 0);
        if (rc == TMAP_ERR_NEGATIVE_ID)
            throw std::runtime_error("The given KF ID was negative. Please make sure the keyFrame IDs are positive. KF ID: " +
                                     std::to_string(static_cast<std::int64_t>(kfID)));
        check(rc);
    }
#ifdef TMAP_HAVE_PCL
    void addNewKeyFrameWithPCL(unsigned long long timestamp_ns, std::uint64_t kfID, std::uint64_t mapID,
                               std::shared_ptr<pcl::PointCloud<pcl::PointXYZ>> cloud)
    {
        if (!cloud) return;
        addNewKeyFrame(timestamp_ns, kfID, mapID, reinterpret_cast<const float *>(cloud->points.data()), cloud->size(),
                       sizeof(pcl::PointXYZ));
    }
#endif
    /// pushToBuffer / addNewKeyFrameTsDouble / addNewKeyFrameTsULong (System.cpp:174-191, 339-344)
    void pushToBuffer(double timestamp_s, const float *xyz, size_t n, size_t stride_bytes)
    {
        check(tmap_push_to_buffer(s_, timestamp_s, xyz, n, stride_bytes));
    }
    void addNewKeyFrameTsDouble(double timestamp_s, std::uint64_t kfID, std::uint64_t mapID)
    {
        check(tmap_add_keyframe_ts(s_, timestamp_s, kfID, mapID));
    }
    void addNewKeyFrameTsULong(unsigned long long timestamp_ns, std::uint64_t kfID, std::uint64_t mapID)
    {
        addNewKeyFrameTsDouble(static_cast<double>(timestamp_ns / 1000000000ULL) + static_cast<double>(timestamp_ns % 1000000000ULL) * 1e-9,
                               kfID, mapID);
    }
    /// KeyFrame(id, ts, pose, cloud_base, mapID) + LocalMap::addAlreadyDeclaredKF (KeyFrame.hpp, LocalMap.cpp:319-328):
    /// an already pruned BASE-frame cloud (packed xyz) registered without the prune step -- how the reference's own
    /// LocalMap tests build their keyframes (test_localmap.cpp:118-131).
    void addKeyFrameBase(std::uint64_t kfID, std::uint64_t mapID, const float *xyz, size_t n)
    {
        check(tmap_add_keyframe_base(s_, kfID, mapID, xyz, n));
    }
    void updateKeyFrame(std::uint64_t kfID, const Pose &pose_map_base, std::uint64_t /*numConnections*/ = 0)
    {
        check(tmap_update_pose(s_, kfID, pose_map_base.m));
    }
    /// One PGO message (global_traversability_node.cpp:184-188): updateKeyFrame for every listed keyframe.
    void updateKeyFrames(const std::vector<std::uint64_t> &kfIDs, const std::vector<Pose> &poses_map_base)
    {
        if (kfIDs.size() != poses_map_base.size()) throw std::invalid_argument("ids / poses size mismatch");
        static_assert(sizeof(Pose) == 12 * sizeof(float), "Pose is 12 packed floats");
        check(tmap_update_poses(s_, kfIDs.data(), kfIDs.empty() ? nullptr : poses_map_base[0].m, kfIDs.size()));
    }
    void deleteKeyFrame(std::uint64_t kfID) { check(tmap_delete_keyframe(s_, kfID)); }
    void updateKFMap(std::uint64_t kfID, std::uint64_t mapID) { check(tmap_update_kf_map(s_, kfID, mapID)); }
    void informLoopClosure() { check(tmap_inform_loop_closure(s_)); }
    std::uint64_t addObstacleKeyFrame(unsigned long long timestamp_ns, const float *xyz, size_t n, size_t stride_bytes)
    {
        std::uint64_t id = 0;
        check(tmap_add_obstacle_keyframe(s_, timestamp_ns, xyz, n, stride_bytes, &id));
        return id;
    }
    void deleteObstacleKeyFrame(std::uint64_t kfID) { check(tmap_delete_obstacle_keyframe(s_, kfID)); }

    /// The active map (System.cpp:357-361): the last one created or chosen by setCurrentMap; nullptr before any.
    std::shared_ptr<LocalMap> getLocalMap()
    {
        std::uint64_t id = 0;
        if (tmap_current_map(s_, &id) != TMAP_OK) return nullptr;
        return getLocalMap(id);
    }
    std::shared_ptr<LocalMap> getLocalMap(std::uint64_t mapID)
    {
        size_t n = 0;
        if (tmap_num_keyframes(s_, mapID, &n) != TMAP_OK) return nullptr;  // unknown map -> nullptr (System.cpp:363-368)
        Lattice l;
        l.x0 = params_.grid_center_x; l.y0 = params_.grid_center_y; l.res = params_.resolution;
        return std::shared_ptr<LocalMap>(new LocalMap(s_, mapID, l));
    }
    /// getGlobalPointCloud (System.hpp:160-161, System.cpp:370-387): packed xyz of every map's stitched cloud,
    /// voxel-filtered per map when all three leaf sizes are positive.
    std::vector<float> getGlobalPointCloud(float voxel_size_x, float voxel_size_y, float voxel_size_z)
    {
        size_t n = 0;
        check(tmap_global_point_cloud(s_, voxel_size_x, voxel_size_y, voxel_size_z, nullptr, 0, &n));
        std::vector<float> out(3 * n);
        if (n) check(tmap_global_point_cloud(s_, voxel_size_x, voxel_size_y, voxel_size_z, out.data(), n, &n));
        out.resize(3 * n);
        return out;
    }
    /// Blocks until the workers have drained their queues (test hook).
    void flush() { check(tmap_flush(s_)); }
    tmap_system *handle() { return s_; }

    // ---- a global map tiled over the GPUs of one box (no reference counterpart) ----
    /// Rank 0 draws the 128-byte job id and ships it to the other ranks over whatever channel the host has.
    static std::array<std::uint8_t, 128> shardUniqueId()
    {
        std::array<std::uint8_t, 128> id{};
        check(tmap_shard_unique_id(id.data()));
        return id;
    }
    void shardCommInit(std::uint64_t mapID, int rank, int world, const std::array<std::uint8_t, 128> &id)
    {
        check(tmap_shard_comm_init(s_, mapID, rank, world, id.data()));
    }
    void shardCommDestroy(std::uint64_t mapID) { check(tmap_shard_comm_destroy(s_, mapID)); }
    /// Collective loop-closure rebuild (LocalMap::clearEntireMap + re-integration, LocalMap.cpp:375-397): this rank's
    /// keyframes (ascending ids) at their corrected poses.
    tmap_shard_info shardRebuild(std::uint64_t mapID, const std::vector<std::uint64_t> &kfIDs, const std::vector<Pose> &poses_map_base)
    {
        if (kfIDs.size() != poses_map_base.size()) throw std::invalid_argument("kfIDs / poses size mismatch");
        tmap_shard_info info;
        check(tmap_shard_rebuild(s_, mapID, kfIDs.data(), kfIDs.size(),
                                 poses_map_base.empty() ? nullptr : poses_map_base.front().m, &info));
        return info;
    }
    /// Collective pose update of some keyframes of this rank (possibly none): one subtract + add batch over the job.
    tmap_shard_info shardUpdate(std::uint64_t mapID, const std::vector<std::uint64_t> &kfIDs, const std::vector<Pose> &poses_map_base)
    {
        if (kfIDs.size() != poses_map_base.size()) throw std::invalid_argument("kfIDs / poses size mismatch");
        tmap_shard_info info;
        check(tmap_shard_update(s_, mapID, kfIDs.data(), kfIDs.size(),
                                poses_map_base.empty() ? nullptr : poses_map_base.front().m, &info));
        return info;
    }

  private:
    static void trampoline(void *user) { (*static_cast<std::function<void()> *>(user))(); }
    static void check(int rc) { if (rc < 0) throw std::runtime_error(tmap_last_error()); }
    tmap_params params_;
    tmap_system *s_ = nullptr;
    std::map<std::uint64_t, std::unique_ptr<std::function<void()>>> callbacks_;
};
}  // namespace traversability_mapping_b200

#endif  // TMAP_SYSTEM_HPP_
