// tmap_ros_shim.hpp -- what the reference's ROS 2 nodes do AROUND the library, as two small
// header-only classes on top of tmap_system.hpp: the bookkeeping only, with the ROS clock, TF lookups and message decoding
// turned into plain arguments.  A node keeps its callbacks and forwards to these; every map operation is a call that
// exists in tmap.h.  Twin of traversability_mapping_b200/ros_shim.py (which the tests drive).
//
//   RollingWindowDriver = LocalTraversabilityNode::scanCallback
//       (traversability_ros_interface/traversability_mapping_ros/src/local_traversability_node.cpp:153-185)
//   ObstacleLayerDriver = GlobalTraversabilityNode::obstacleCallback / expireObstacles
//       (traversability_ros_interface/traversability_mapping_ros/src/global_traversability_node.cpp:224-283)
#pragma once
#include "tmap_system.hpp"

#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

namespace traversability_mapping_b200
{
class RollingWindowDriver
{
  public:
    RollingWindowDriver(System &system, std::uint64_t window_size, double keyframe_min_displacement_m = 0.0, std::uint64_t map_id = 0)
        : system_(system), window_(window_size), map_(map_id), min_disp_(keyframe_min_displacement_m)
    {
    }
    // One synced (cloud, pose) pair -> at most one keyframe.  cloud_xyz4: n points, 4 floats each (pcl::PointXYZ); pose: map <- base.
    // Returns true when the scan became keyframe `*kf_out`.
    bool scan(unsigned long long stamp_ns, const float *cloud_xyz4, std::size_t n, const Pose &pose, std::uint64_t *kf_out = nullptr)
    {
        const double px = pose.m[3], py = pose.m[7], pz = pose.m[11];
        robot_x_ = px; robot_y_ = py; have_pose_ = true;  // the crop follows the robot on every scan, dropped ones included
        if (have_last_)
        {
            const double dx = px - lx_, dy = py - ly_, dz = pz - lz_;
            if (std::sqrt(dx * dx + dy * dy + dz * dz) < min_disp_)
                return false;  // displacement gate (local_traversability_node.cpp:162-166)
        }
        lx_ = px; ly_ = py; lz_ = pz; have_last_ = true;
        const std::uint64_t id = next_id_++;  // ids stay contiguous: the eviction below relies on it
        system_.addNewKeyFrame(stamp_ns, id, map_, cloud_xyz4, n, 16);  // = addNewKeyFrameWithPCL on the raw pcl::PointXYZ array
        system_.updateKeyFrame(id, pose);  // supplies the pose AND triggers the (only) bin
        if (id >= window_)
            system_.deleteKeyFrame(id - window_);  // once N + 1 are in flight, drop the oldest
        if (kf_out) *kf_out = id;
        return true;
    }
    bool robotXY(double &x, double &y) const { x = robot_x_; y = robot_y_; return have_pose_; }

  private:
    System &system_;
    std::uint64_t window_, map_, next_id_ = 0;
    double min_disp_, lx_ = 0, ly_ = 0, lz_ = 0, robot_x_ = 0, robot_y_ = 0;
    bool have_last_ = false, have_pose_ = false;
};

class ObstacleLayerDriver
{
  public:
    ObstacleLayerDriver(System &system, double max_range, double min_range, double add_rate_hz, double temporal_buffer_s)
        : system_(system), max_range_(max_range), min_range_(min_range), period_s_(add_rate_hz > 0 ? 1.0 / add_rate_hz : 0.0),
          buffer_s_(temporal_buffer_s)
    {
    }
    // now_s: the node's clock at reception.  T_base_sensor / T_map_base: the two TF lookups (nullptr = the lookup failed).
    // Returns the obstacle keyframe id, 0 when the cloud was not fused.
    std::uint64_t cloud(double now_s, unsigned long long stamp_ns, const float *cloud_xyz4, std::size_t n, const Pose *T_base_sensor,
                        const Pose *T_map_base)
    {
        if (!ready_)
        {
            if (!T_base_sensor) return 0;  // TF not up yet; try again on the next cloud
            system_.setObstacleParameters(*T_base_sensor, max_range_, min_range_);
            ready_ = true;
        }
        if (have_last_ && (now_s - last_add_s_) < period_s_)
            return 0;  // ingestion-rate gate
        last_add_s_ = now_s;
        have_last_ = true;
        if (!T_map_base) return 0;
        const std::uint64_t id = system_.addObstacleKeyFrame(stamp_ns, cloud_xyz4, n, 16);
        if (id == 0) return 0;  // params unset / cloud pruned empty
        system_.updateKeyFrame(id, *T_map_base);  // supply pose + trigger the (only) bin
        deadlines_[id] = now_s + buffer_s_;
        return id;
    }
    // the publish timer's expireObstacles: wipe obstacle keyframes whose temporal buffer has elapsed
    std::vector<std::uint64_t> expire(double now_s)
    {
        std::vector<std::uint64_t> gone;
        for (auto it = deadlines_.begin(); it != deadlines_.end();)
        {
            if (now_s >= it->second)
            {
                system_.deleteObstacleKeyFrame(it->first);
                gone.push_back(it->first);
                it = deadlines_.erase(it);
            }
            else
                ++it;
        }
        return gone;
    }
    std::size_t live() const { return deadlines_.size(); }

  private:
    System &system_;
    double max_range_, min_range_, period_s_, buffer_s_, last_add_s_ = 0;
    bool ready_ = false, have_last_ = false;
    std::map<std::uint64_t, double> deadlines_;
};
}  // namespace traversability_mapping_b200
