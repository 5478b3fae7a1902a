// cpp_mirror_check.cpp -- the header-only C++ mirror (include/tmap_system.hpp) compiles against the
// C ABI; without a GPU construction fails loudly (exit 42); with one, the reference's LocalMap tests
// (traversability_mapping/test/test_localmap.cpp:175-340, the same list oracle/kat.cpp replays on the
// oracle) run through the C++ surface a SLAM integration would use.  Exit 0 = every check held.
#include "../include/tmap_system.hpp"
#include <array>
#include <cmath>
#include <cstdio>
#include <set>
using namespace traversability_mapping_b200;

static int g_fail = 0, g_checks = 0;
static const char *g_test = "";
#define TEST(name) g_test = name;
#define CHECK(c)                                                                         \
    do {                                                                                 \
        ++g_checks;                                                                      \
        if (!(c)) { ++g_fail; std::printf("FAIL %s: %s (line %d)\n", g_test, #c, __LINE__); } \
    } while (0)

static tmap_params localmapParams()  // test_localmap.cpp:47-66
{
    tmap_params p;
    tmap_default_params(&p);
    p.resolution = 0.25; p.half_size = 7.5; p.extend_length = 30.0; p.robot_radius = 0.4;
    p.min_points_per_grid = 5; p.max_range = 300.0; p.min_range = 0.0; p.is_kf_optimization_enabled = 1;
    return p;
}
static Pose translation(float x, float y, float z)
{
    Pose T;
    T.m[3] = x; T.m[7] = y; T.m[11] = z;
    return T;
}
using Pts = std::vector<std::array<float, 3>>;
static void addBase(System &s, std::uint64_t id, const Pts &pts)
{
    std::vector<float> flat;
    for (auto &p : pts) { flat.push_back(p[0]); flat.push_back(p[1]); flat.push_back(p[2]); }
    s.addKeyFrameBase(id, 0, flat.data(), pts.size());
}
static std::set<std::uint64_t> occupied(System &s)
{
    auto v = s.getLocalMap(0)->occupiedCellKeys();
    return std::set<std::uint64_t>(v.begin(), v.end());
}
static std::set<std::uint64_t> drain(System &s)
{
    auto v = s.getLocalMap(0)->takeChangedCells();
    return std::set<std::uint64_t>(v.begin(), v.end());
}

int main()
{
    const tmap_params LP = localmapParams();
    try
    {
        System probe(LP);
    }
    catch (const std::runtime_error &e)
    {
        std::printf("no-gpu: System() failed loudly: %s\n", e.what());
        return 42;
    }
    TEST("LocalMap.AddRegistersWithoutMutatingGrid (test_localmap.cpp:175-183)")
    {
        System s(LP); s.addNewLocalMap(0);
        addBase(s, 7, {{0.f, 0.f, 0.f}});
        CHECK(s.getLocalMap(0)->numKeyFrames() == 1u);
        s.flush();
        CHECK(occupied(s).empty());
        CHECK(drain(s).empty());
    }
    TEST("LocalMap.AddMultipleTracksAllInKeyFramesMap (test_localmap.cpp:185-193)")
    {
        System s(LP); s.addNewLocalMap(0);
        addBase(s, 1, {{0.f, 0.f, 0.f}}); addBase(s, 2, {{1.f, 0.f, 0.f}}); addBase(s, 3, {{2.f, 0.f, 0.f}});
        CHECK(s.getLocalMap(0)->numKeyFrames() == 3u);
    }
    TEST("LocalMap.SetPosePopulatesGridCells (test_localmap.cpp:197-207)")
    {
        System s(LP);
        int hits = 0;
        s.addNewLocalMap(0, [&] { ++hits; });
        addBase(s, 1, {{0.f, 0.f, 0.f}, {1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {2.f, 2.f, 0.f}});
        s.updateKeyFrame(1, Pose::Identity());
        s.flush();
        const std::set<std::uint64_t> want = {Lattice::key(0, 0), Lattice::key(4, 0), Lattice::key(0, 4), Lattice::key(8, 8)};
        CHECK(occupied(s) == want);
        CHECK(hits == 1);  // onUpdate_ once per grid-changing keyframe (LocalMap.cpp:479-481)
    }
    TEST("LocalMap.ChangedCellsRecordedThenClearedOnDrain (test_localmap.cpp:209-220)")
    {
        System s(LP); s.addNewLocalMap(0);
        addBase(s, 1, {{0.f, 0.f, 0.f}, {2.f, 2.f, 0.f}});
        s.updateKeyFrame(1, Pose::Identity());
        s.flush();
        auto ch = drain(s);
        CHECK(ch.count(Lattice::key(0, 0)) == 1u);
        CHECK(ch.count(Lattice::key(8, 8)) == 1u);
        CHECK(drain(s).empty());
    }
    TEST("LocalMap.GridGrowsToIncludeFarKeyframe (test_localmap.cpp:222-232)")
    {
        System s(LP); s.addNewLocalMap(0);
        int32_t kx0, ky0, kx1, ky1;
        s.getLocalMap(0)->gridExtent(kx0, ky0);
        CHECK(std::fabs((2 * kx0 + 1) * 0.25 - 15.25) < 1e-12);
        addBase(s, 1, {{20.f, 0.f, 0.f}});
        s.updateKeyFrame(1, Pose::Identity());
        s.flush();
        s.getLocalMap(0)->gridExtent(kx1, ky1);
        CHECK(kx1 > kx0);
        CHECK(ky1 == ky0 + 1);  // y re-derives from length/2 = 7.625 -> ceil(30.5) = 31 cells (Helpers.cpp:56-67)
        CHECK(occupied(s).count(Lattice::key(80, 0)) == 1u);
    }
    TEST("LocalMap.PoseUpdateMovesContribution (test_localmap.cpp:234-246)")
    {
        System s(LP); s.addNewLocalMap(0);
        addBase(s, 1, {{0.f, 0.f, 0.f}});
        s.updateKeyFrame(1, Pose::Identity());
        s.flush();
        CHECK(occupied(s).count(Lattice::key(0, 0)) == 1u);
        s.updateKeyFrame(1, translation(10.f, 0.f, 0.f));
        s.flush();
        auto occ = occupied(s);
        CHECK(occ.count(Lattice::key(40, 0)) == 1u);
        CHECK(occ.count(Lattice::key(0, 0)) == 0u);
    }
    TEST("LocalMap.DetachRemovesContributionAndKeyFrame / DetachUnknownId (test_localmap.cpp:250-268)")
    {
        System s(LP); s.addNewLocalMap(0);
        addBase(s, 1, {{0.f, 0.f, 0.f}, {2.f, 2.f, 0.f}});
        s.updateKeyFrame(1, Pose::Identity());
        s.flush();
        CHECK(!occupied(s).empty());
        s.deleteKeyFrame(1);
        CHECK(s.getLocalMap(0)->numKeyFrames() == 0u);
        CHECK(occupied(s).empty());
        CHECK(tmap_delete_keyframe(s.handle(), 999) == TMAP_IGNORED_UNKNOWN_KF);
    }
    TEST("LocalMap.DeleteKeyFrameRemovesIt (test_localmap.cpp:270-280)")
    {
        System s(LP); s.addNewLocalMap(0);
        addBase(s, 5, {{0.f, 0.f, 0.f}});
        s.updateKeyFrame(5, translation(0.f, 0.f, 0.f));
        s.flush();
        size_t npts = 0;
        CHECK(tmap_keyframe_points(s.handle(), 5, &npts) == TMAP_OK && npts == 1u);
        s.deleteKeyFrame(5);
        CHECK(tmap_keyframe_points(s.handle(), 5, &npts) == TMAP_IGNORED_UNKNOWN_KF);
        CHECK(occupied(s).empty());
    }
    TEST("LocalMap.ClearBlanksGridAndRecordsChanged (test_localmap.cpp:284-301)")
    {
        tmap_params p = LP;
        p.is_kf_optimization_enabled = 0;
        System s(p); s.addNewLocalMap(0);
        addBase(s, 1, {{0.f, 0.f, 0.f}, {2.f, 2.f, 0.f}});
        s.updateKeyFrame(1, Pose::Identity());
        s.flush();
        auto was = occupied(s);
        CHECK(!was.empty());
        drain(s);
        s.informLoopClosure();
        s.flush();  // the cloud was dropped -> nothing can be re-added
        CHECK(occupied(s).empty());
        auto ch = drain(s);
        for (auto k : was) CHECK(ch.count(k) == 1u);
    }
    TEST("KeyFrame latest-wins pose, consumed once (test_keyframe.cpp:68-116 through the queue)")
    {
        System s(LP);
        int hits = 0;
        s.addNewLocalMap(0, [&] { ++hits; });
        addBase(s, 1, {{0.f, 0.f, 0.f}});
        s.getLocalMap(0)->lock();  // the worker cannot start a pass: both poses are queued before it runs
        s.updateKeyFrame(1, translation(1.f, 1.f, 1.f));
        s.updateKeyFrame(1, translation(8.f, 9.f, 10.f));
        size_t depth = 0;
        tmap_queue_depth(s.handle(), 0, &depth);
        s.getLocalMap(0)->unlock();
        s.flush();
        CHECK(depth <= 1u);
        CHECK(occupied(s).count(Lattice::key(32, 36)) == 1u);
        CHECK(hits == 1);
    }
    TEST("LocalMap.StitchTransformsRetainedCloudsByPose (test_localmap.cpp:305-318)")
    {
        System s(LP); s.addNewLocalMap(0);
        addBase(s, 1, {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}});
        s.updateKeyFrame(1, translation(5.f, -3.f, 2.f));
        s.flush();
        auto out = s.getLocalMap(0)->getStitchedPointCloud();
        CHECK(out.size() == 6u);
        std::set<std::array<float, 3>> pts;
        for (size_t i = 0; i + 2 < out.size(); i += 3) pts.insert({out[i], out[i + 1], out[i + 2]});
        CHECK(pts.count({6.f, -3.f, 2.f}) == 1u);
        CHECK(pts.count({5.f, -2.f, 2.f}) == 1u);
        auto glob = s.getGlobalPointCloud(0.f, 0.f, 0.f);  // System.cpp:370-387: one map -> the same cloud
        CHECK(glob == out);
    }
    TEST("LocalMap.StitchVoxelDownsampleReducesCount (test_localmap.cpp:329-340)")
    {
        System s(LP); s.addNewLocalMap(0);
        Pts pts;
        for (int i = 0; i < 10; ++i) pts.push_back({0.001f * i, 0.f, 0.f});
        addBase(s, 1, pts);
        auto out = s.getLocalMap(0)->getStitchedPointCloud(1.f, 1.f, 1.f);
        CHECK(out.size() == 3u);  // one voxel: the centroid of the ten points
        CHECK(!out.empty() && std::fabs(out[0] - 0.0045f) < 1e-6f);
    }
    TEST("System surface: negative id throws, unknown map -> nullptr, active map, duplicate map keeps its callback")
    {
        System s(LP);
        CHECK(s.getLocalMap() == nullptr);  // no map yet (System.cpp:357-361)
        int first = 0, second = 0;
        s.addNewLocalMap(5, [&] { ++first; });
        s.addNewLocalMap(5, [&] { ++second; });  // System.cpp:65-69: logged, ignored
        CHECK(s.getLocalMap() != nullptr);       // addNewLocalMap made map 5 the active one (System.cpp:71)
        CHECK(s.getLocalMap(9) == nullptr);
        const float pts[1][4] = {{2, 0, 0, 0}};
        bool thrown = false;
        try { s.addNewKeyFrame(0, ~0ull, 5, &pts[0][0], 1, 16); } catch (const std::runtime_error &) { thrown = true; }
        CHECK(thrown);  // System.cpp:163-168
        const float b[3] = {0.f, 0.f, 0.f};
        s.addKeyFrameBase(1, 5, b, 1);
        s.updateKeyFrame(1, Pose::Identity());
        s.flush();
        CHECK(first == 1 && second == 0);
    }
    std::printf("cpp mirror: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail ? 1 : 0;
}
