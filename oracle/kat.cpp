/* kat.cpp -- replays the reference's own gtest known-answer vectors against the ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  Each block names the reference test it re-expresses
 * (/root/reference/traversability_mapping/test/<file>:<line>).  The point generators use
 * std::mt19937 + std::uniform_real_distribution from libstdc++ exactly as the reference
 * tests do, so the sampled points are the same.  Exit code = number of failed checks.
 */
#include "tmap_oracle.cpp"  /* single TU: the KATs reach the oracle's internal types */

#include <cstdio>
#include <cstring>
#include <functional>
#include <random>

static int g_fail = 0, g_checks = 0;
static const char *g_test = "";
#define CHECK(cond)                                                                   \
    do {                                                                              \
        ++g_checks;                                                                   \
        if (!(cond)) { ++g_fail; std::printf("FAIL %s: %s (line %d)\n", g_test, #cond, __LINE__); } \
    } while (0)
#define NEAR(a, b, tol) CHECK(std::abs((a) - (b)) <= (tol))
#define TEST(name) g_test = name;

static tmo_params localmapParams()  /* test_localmap.cpp:53-78 */
{
    tmo_params p;
    tmo_default_params(&p);
    p.resolution = 0.25; p.half_size = 7.5; p.extend_length = 30.0; p.robot_radius = 0.4;
    p.ground_clearance = 0.15; p.max_slope = 0.4; p.min_points_per_grid = 5;
    p.is_kf_optimization_enabled = 1; p.inflation_enabled = 0; p.inflation_radius = 0.5;
    p.cost_scaling_factor = 3.0; p.lethal_step_threshold = 0.95; p.robot_height = 3.5;
    p.max_range = 300.0; p.min_range = 1.0;
    return p;
}

static const float kIdent[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
static void translation(float x, float y, float z, float T[12])
{
    std::memcpy(T, kIdent, sizeof(kIdent));
    T[3] = x; T[7] = y; T[11] = z;
}

static CellMoment makeCell(const double c[3], const std::vector<std::array<double, 3>> &pts)
{
    CellMoment cm;  /* test_moments.cpp:24-35 */
    cm.c[0] = c[0]; cm.c[1] = c[1]; cm.c[2] = c[2];
    for (const auto &p : pts) cm.data.insert(p[0] - c[0], p[1] - c[1], p[2] - c[2]);
    return cm;
}

/* test_moments.cpp:169-192 */
static std::vector<CellMoment> sampleField(const Lattice &lat, int cx, int cy,
                                           std::function<double(double, double)> f,
                                           double noise = 0.0, unsigned seed = 5)
{
    std::vector<CellMoment> cells;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> jitter(-0.4 * lat.res, 0.4 * lat.res);
    std::uniform_real_distribution<double> nz(-noise, noise);
    for (int i = cx - 1; i <= cx + 1; ++i)
        for (int j = cy - 1; j <= cy + 1; ++j)
        {
            const double ctr[3] = {lat.cx(i), lat.cy(j), 0.0};
            std::vector<std::array<double, 3>> pts;
            for (int k = 0; k < 20; ++k)
            {
                double x = ctr[0] + jitter(rng);
                double y = ctr[1] + jitter(rng);
                double z = f(x, y) + nz(rng);
                pts.push_back({x, y, z});
            }
            cells.push_back(makeCell(ctr, pts));
        }
    return cells;
}

static void covOf(const Moments &m, double c[3][3])
{
    const double n = m.N;
    const double mu[3] = {m.sx / n, m.sy / n, m.sz / n};
    const double Q[3][3] = {{m.sx2, m.sxy, m.sxz}, {m.sxy, m.sy2, m.syz}, {m.sxz, m.syz, m.sz2}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c[i][j] = Q[i][j] / n - mu[i] * mu[j];
}

static std::set<uint64_t> occupiedSet(tmo_map *m)
{
    size_t n = tmo_occupied_cell_keys(m, nullptr, 0);
    std::vector<uint64_t> k(n);
    tmo_occupied_cell_keys(m, k.data(), n);
    return std::set<uint64_t>(k.begin(), k.end());
}
static std::set<uint64_t> drainChanged(tmo_map *m)
{
    size_t n = tmo_take_changed_cells(m, nullptr, 0);
    std::vector<uint64_t> k(n + 1);
    tmo_take_changed_cells(m, k.data(), n);
    k.resize(n);
    return std::set<uint64_t>(k.begin(), k.end());
}

int main()
{
    const double PI = 3.14159265358979323846;
    /* ================= test_moments.cpp ================= */
    TEST("Moments.FuseEqualsCombinedInsert (test_moments.cpp:48-65)")
    {
        Moments a, b, combined;
        std::mt19937 rng(42);
        std::uniform_real_distribution<double> u(-0.1, 0.1);
        for (int i = 0; i < 100; ++i)
        {
            double x = u(rng), y = u(rng), z = u(rng);
            if (i % 2 == 0) a.insert(x, y, z); else b.insert(x, y, z);
            combined.insert(x, y, z);
        }
        Moments fused = a;
        fused.fuse(b);
        CHECK(fused.N == combined.N);
        NEAR(fused.sx2, combined.sx2, 1e-9);
        NEAR(fused.sxy, combined.sxy, 1e-9);
        NEAR(fused.sz, combined.sz, 1e-9);
    }
    TEST("Moments.BarycenterIsMeanOfPoints (test_moments.cpp:85-100)")
    {
        std::mt19937 rng(13);
        std::uniform_real_distribution<double> u(-0.5, 0.5);
        double sum[3] = {0, 0, 0};
        Moments m;
        for (int i = 0; i < 100; ++i)
        {
            double p[3] = {u(rng), u(rng), u(rng)};
            m.insert(p[0], p[1], p[2]);
            for (int k = 0; k < 3; ++k) sum[k] += p[k];
        }
        NEAR(m.sx / m.N, sum[0] / 100.0, 1e-12);
        NEAR(m.sy / m.N, sum[1] / 100.0, 1e-12);
        NEAR(m.sz / m.N, sum[2] / 100.0, 1e-12);
    }
    TEST("Moments.ShiftReExpressesAboutNewOrigin (test_moments.cpp:102-130)")
    {
        std::vector<std::array<double, 3>> pts;
        std::mt19937 rng(17);
        std::uniform_real_distribution<double> u(-0.3, 0.3);
        for (int i = 0; i < 150; ++i) { double x = u(rng), y = u(rng), z = u(rng); pts.push_back({x, y, z}); }
        const double oA[3] = {1.0, 2.0, 0.0}, oB[3] = {1.25, 2.0, 0.0};
        Moments a, b;
        for (const auto &p : pts)
        {
            a.insert(p[0] - oA[0], p[1] - oA[1], p[2] - oA[2]);
            b.insert(p[0] - oB[0], p[1] - oB[1], p[2] - oB[2]);
        }
        b.shift(oB[0] - oA[0], oB[1] - oA[1], oB[2] - oA[2]);
        NEAR(a.sx / a.N, b.sx / b.N, 1e-12);
        NEAR(a.sy / a.N, b.sy / b.N, 1e-12);
        NEAR(a.sz / a.N, b.sz / b.N, 1e-12);
        double ca[3][3], cb[3][3];
        covOf(a, ca); covOf(b, cb);
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) NEAR(ca[i][j], cb[i][j], 1e-12);
    }
    TEST("Moments.ShiftedCellsFuseAsCombinedInsert (test_moments.cpp:132-164)")
    {
        std::vector<std::array<double, 3>> A, B;
        std::mt19937 rng(19);
        std::uniform_real_distribution<double> u(-0.3, 0.3);
        for (int i = 0; i < 60; ++i) { double x = u(rng), y = u(rng), z = u(rng); A.push_back({x, y, z}); }
        for (int i = 0; i < 40; ++i) { double x = 2.0 + u(rng), y = u(rng), z = u(rng); B.push_back({x, y, z}); }
        Moments a, b, combined;
        for (const auto &p : A) a.insert(p[0], p[1], p[2]);
        for (const auto &p : B) b.insert(p[0] - 2.0, p[1], p[2]);
        b.shift(2.0, 0.0, 0.0);
        Moments fused = a;
        fused.fuse(b);
        for (const auto &p : A) combined.insert(p[0], p[1], p[2]);
        for (const auto &p : B) combined.insert(p[0], p[1], p[2]);
        CHECK(fused.N == combined.N);
        NEAR(fused.sx, combined.sx, 1e-9);
        NEAR(fused.sx2, combined.sx2, 1e-9);
        NEAR(fused.sxy, combined.sxy, 1e-9);
        NEAR(fused.sxz, combined.sxz, 1e-9);
    }
    Lattice lat; lat.x0 = 0; lat.y0 = 0; lat.res = 0.25;
    TEST("Goodness.FlatGroundIsFlatAndSmooth (test_moments.cpp:194-206)")
    {
        auto cells = sampleField(lat, 0, 0, [](double, double) { return 2.0; });
        auto haz = computeGoodness(cells[4], cells, 9, 0.15, 0.4, 15);
        NEAR(haz[HAZ_SLOPE], 0.0, 1e-3);
        NEAR(haz[HAZ_ROUGHNESS], 0.0, 1e-3);
        NEAR(haz[HAZ_STEP], 0.0, 1e-3);
        NEAR(haz[HAZ_ELEVATION], 2.0, 1e-2);
        CHECK(haz[HAZ_BORDER] == 1.0);
    }
    TEST("Goodness.RampRecoversSlopeAngle (test_moments.cpp:208-219)")
    {
        const double angle = 20.0 * PI / 180.0, m = std::tan(angle);
        auto cells = sampleField(lat, 0, 0, [m](double x, double) { return m * x; });
        auto haz = computeGoodness(cells[4], cells, 9, 0.15, 0.4, 15);
        NEAR(haz[HAZ_SLOPE] * 0.4, angle, 2.0 * PI / 180.0);
    }
    TEST("Goodness.RoughTerrainHasRoughness (test_moments.cpp:221-229)")
    {
        auto flat = sampleField(lat, 0, 0, [](double, double) { return 0.0; }, 0.0);
        auto rough = sampleField(lat, 0, 0, [](double, double) { return 0.0; }, 0.05);
        auto hf = computeGoodness(flat[4], flat, 9, 0.15, 0.4, 15);
        auto hr = computeGoodness(rough[4], rough, 9, 0.15, 0.4, 15);
        CHECK(hr[HAZ_ROUGHNESS] > hf[HAZ_ROUGHNESS]);
    }
    TEST("Goodness.OccupiedFractionGateRejectsSparse (test_moments.cpp:231-240)")
    {
        auto cells = sampleField(lat, 0, 0, [](double, double) { return 0.0; });
        std::vector<CellMoment> sparse = {cells[4], cells[0]};
        auto haz = computeGoodness(cells[4], sparse, 9, 0.15, 0.4, 15);
        CHECK(haz[HAZ_BORDER] == 1.0);
        CHECK(std::isnan(haz[HAZ_OVERALL]));
    }
    TEST("Goodness.FusionUsesAllPointsAcrossKeyframes (test_moments.cpp:242-262)")
    {
        const double m = std::tan(15.0 * PI / 180.0);
        auto kfA = sampleField(lat, 0, 0, [m](double x, double) { return m * x; }, 0.0, 11);
        auto kfB = sampleField(lat, 0, 0, [m](double x, double) { return m * x; }, 0.0, 22);
        std::vector<CellMoment> fused;
        for (size_t i = 0; i < kfA.size(); ++i)
        {
            CellMoment c = kfA[i];
            c.data.fuse(kfB[i].data);
            fused.push_back(c);
        }
        auto hz = computeGoodness(fused[4], fused, 9, 0.15, 0.4, 30);
        NEAR(hz[HAZ_SLOPE] * 0.4, 15.0 * PI / 180.0, 2.0 * PI / 180.0);
        CHECK(fused[4].data.N == kfA[4].data.N + kfB[4].data.N);
    }
    TEST("Goodness.EmptyQueryReturnsAllNaN (test_moments.cpp:264-276)")
    {
        auto cells = sampleField(lat, 0, 0, [](double, double) { return 0.0; });
        CellMoment emptyQuery;
        auto haz = computeGoodness(emptyQuery, cells, 9, 0.15, 0.4, 15);
        CHECK(std::isnan(haz[HAZ_OVERALL]));
        CHECK(std::isnan(haz[HAZ_ELEVATION]));
        CHECK(std::isnan(haz[HAZ_BORDER]));
        CHECK(std::isnan(haz[HAZ_SLOPE]));
    }
    TEST("Goodness.ElevationGatedWhenBorder (test_moments.cpp:278-288)")
    {
        auto cells = sampleField(lat, 0, 0, [](double, double) { return 2.0; });
        std::vector<CellMoment> sparse = {cells[4], cells[0]};
        auto haz = computeGoodness(cells[4], sparse, 9, 0.15, 0.4, 15);
        CHECK(haz[HAZ_BORDER] == 1.0);
        CHECK(std::isnan(haz[HAZ_ELEVATION]));
    }
    TEST("Goodness.InsufficientVicinityPointsGate (test_moments.cpp:290-303)")
    {
        auto cells = sampleField(lat, 0, 0, [](double, double) { return 0.0; });
        auto haz = computeGoodness(cells[4], cells, 9, 0.15, 0.4, 100000);
        NEAR(haz[HAZ_BORDER], 20.0 / 100000.0, 1e-9);
        CHECK(std::isnan(haz[HAZ_OVERALL]));
        CHECK(std::isnan(haz[HAZ_ELEVATION]));
    }
    TEST("Goodness.FlatNormalPointsUp (test_moments.cpp:305-314)")
    {
        auto cells = sampleField(lat, 0, 0, [](double, double) { return 1.0; });
        auto haz = computeGoodness(cells[4], cells, 9, 0.15, 0.4, 15);
        const double n[3] = {haz[HAZ_NORMAL_X], haz[HAZ_NORMAL_Y], haz[HAZ_NORMAL_Z]};
        NEAR(std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1.0, 1e-6);
        CHECK(n[2] > 0.0);
        NEAR(n[0], 0.0, 1e-2); NEAR(n[1], 0.0, 1e-2); NEAR(n[2], 1.0, 1e-2);
    }
    TEST("Goodness.RampNormalMatchesSlope (test_moments.cpp:316-330)")
    {
        const double angle = 20.0 * PI / 180.0, m = std::tan(angle);
        auto cells = sampleField(lat, 0, 0, [m](double x, double) { return m * x; });
        auto haz = computeGoodness(cells[4], cells, 9, 0.15, 0.4, 15);
        const double n[3] = {haz[HAZ_NORMAL_X], haz[HAZ_NORMAL_Y], haz[HAZ_NORMAL_Z]};
        NEAR(std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]), 1.0, 1e-6);
        NEAR(n[2], std::cos(angle), 1e-2);
        NEAR(n[0], -std::sin(angle), 2e-2);
        NEAR(n[1], 0.0, 1e-2);
    }
    TEST("Goodness.SteepSlopeClampsToOne (test_moments.cpp:332-340)")
    {
        const double m = std::tan(20.0 * PI / 180.0);
        auto cells = sampleField(lat, 0, 0, [m](double x, double) { return m * x; });
        auto haz = computeGoodness(cells[4], cells, 9, 0.15, 0.05, 15);
        CHECK(haz[HAZ_SLOPE] == 1.0);
    }
    TEST("Goodness.StepDiscontinuityProducesStep (test_moments.cpp:342-353)")
    {
        auto flat = sampleField(lat, 0, 0, [](double, double) { return 0.0; });
        auto stepped = sampleField(lat, 0, 0, [](double x, double) { return x < 0.0 ? 0.0 : 0.4; });
        auto hf = computeGoodness(flat[4], flat, 9, 0.15, 0.4, 15);
        auto hs = computeGoodness(stepped[4], stepped, 9, 0.15, 0.4, 15);
        CHECK(hs[HAZ_STEP] > hf[HAZ_STEP]);
        CHECK(hs[HAZ_STEP] > 0.3);
    }
    TEST("Goodness.OverallIsMaxOfComponentHazards (test_moments.cpp:355-364)")
    {
        const double m = std::tan(10.0 * PI / 180.0);
        auto cells = sampleField(lat, 0, 0, [m](double x, double) { return m * x; }, 0.02);
        auto haz = computeGoodness(cells[4], cells, 9, 0.15, 0.4, 15);
        CHECK(haz[HAZ_OVERALL] == std::max(haz[HAZ_SLOPE], std::max(haz[HAZ_STEP], haz[HAZ_ROUGHNESS])));
    }
    TEST("Lattice.RoundTripAndKey (test_moments.cpp:368-380)")
    {
        int ci, cj;
        lat.cellOf(1.04, -0.36, ci, cj);
        CHECK(ci == 4); CHECK(cj == -1);
        int a, b;
        Lattice::unkey(Lattice::key(ci, cj), a, b);
        CHECK(a == ci); CHECK(b == cj);
    }
    TEST("Lattice.CellCenterIndependentOfMapOrigin (test_moments.cpp:382-390)")
    {
        NEAR(lat.cx(10), 2.5, 1e-12);
        NEAR(lat.cy(-7), -1.75, 1e-12);
    }

    /* ================= test_keyframe.cpp (rebin KATs) ================= */
    auto rebinOne = [&](std::vector<std::array<float, 3>> cloud, const float T[12]) {
        std::map<uint64_t, Moments> p;
        rebin(lat, cloud, T, p);
        return p;
    };
    TEST("KeyFrame.RebinBinsCloudIntoCellLocalMoments (test_keyframe.cpp:120-139)")
    {
        auto p = rebinOne({{1.04f, -0.36f, 2.0f}}, kIdent);
        CHECK(p.size() == 1u);
        CHECK(p.count(Lattice::key(4, -1)) == 1u);
        const Moments &m = p.at(Lattice::key(4, -1));
        CHECK(m.N == 1u);
        NEAR(m.sx / m.N, 0.04, 1e-4);
        NEAR(m.sy / m.N, -0.11, 1e-4);
        NEAR(m.sz / m.N, 2.0, 1e-4);
    }
    TEST("KeyFrame.RebinAppliesPoseTransform (test_keyframe.cpp:141-153)")
    {
        /* Eigen::AngleAxisf(M_PI/2.f, UnitZ).toRotationMatrix(): c = cosf, s = sinf of (float)(pi/2) */
        const float ang = static_cast<float>(PI) / 2.f;
        const float c = std::cos(ang), s = std::sin(ang);
        const float yaw90[12] = {c, -s, 0, 0, s, c, 0, 0, 0, 0, 1, 0};
        auto p = rebinOne({{1.f, 0.f, 0.f}}, yaw90);
        CHECK(p.size() == 1u);
        CHECK(p.count(Lattice::key(0, 4)) == 1u);
    }
    TEST("KeyFrame.RebinUsesPassedPoseNotLatestPose (test_keyframe.cpp:155-167)")
    {
        /* the oracle's rebin takes the pose as an argument by construction; a pending far-away pose on the keyframe
         * must not leak into a rebin at identity */
        KeyFrame kf;
        kf.cloud = {{1.04f, -0.36f, 2.0f}};
        translation(10.f, 0.f, 0.f, kf.pose);
        kf.pending = true;
        rebin(lat, kf.cloud, kIdent, kf.partials);
        CHECK(kf.partials.size() == 1u);
        CHECK(kf.partials.count(Lattice::key(4, -1)) == 1u);
        CHECK(kf.partials.count(Lattice::key(44, -1)) == 0u);
    }
    TEST("KeyFrame.RebinReplacesPreviousPartials (test_keyframe.cpp:169-182)")
    {
        float T[12];
        translation(10.f, 0.f, 0.f, T);
        std::map<uint64_t, Moments> p;
        std::vector<std::array<float, 3>> cloud = {{1.04f, -0.36f, 2.0f}};
        rebin(lat, cloud, kIdent, p);
        CHECK(p.count(Lattice::key(4, -1)) == 1u);
        rebin(lat, cloud, T, p);
        CHECK(p.size() == 1u);
        CHECK(p.count(Lattice::key(4, -1)) == 0u);
        CHECK(p.count(Lattice::key(44, -1)) == 1u);
    }
    TEST("KeyFrame.RebinAccumulatesPointsInSameCell (test_keyframe.cpp:184-193)")
    {
        auto p = rebinOne({{0.02f, -0.03f, 1.0f}, {-0.01f, 0.04f, 1.2f}}, kIdent);
        CHECK(p.size() == 1u);
        CHECK(p.at(Lattice::key(0, 0)).N == 2u);
    }
    TEST("KeyFrame.RebinSeparatesPointsInDifferentCells (test_keyframe.cpp:195-204)")
    {
        auto p = rebinOne({{0.f, 0.f, 0.f}, {5.f, 0.f, 0.f}}, kIdent);
        CHECK(p.size() == 2u);
        CHECK(p.count(Lattice::key(0, 0)) == 1u);
        CHECK(p.count(Lattice::key(20, 0)) == 1u);
    }
    TEST("KeyFrame.RebinEmptyCloudYieldsNoPartials (test_keyframe.cpp:206-213)")
    {
        auto p = rebinOne({}, kIdent);
        CHECK(p.empty());
    }

    /* ================= test_localmap.cpp ================= */
    const tmo_params LP = localmapParams();
    auto addBase = [](tmo_map *m, uint64_t id, std::vector<std::array<float, 3>> pts) {
        std::vector<float> flat;
        for (auto &p : pts) { flat.push_back(p[0]); flat.push_back(p[1]); flat.push_back(p[2]); }
        return tmo_add_keyframe_base(m, id, flat.data(), pts.size());
    };
    TEST("LocalMap.AddRegistersWithoutMutatingGrid (test_localmap.cpp:175-183)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 7, {{0.f, 0.f, 0.f}});
        CHECK(tmo_num_keyframes(m) == 1u);
        CHECK(tmo_process(m) == 0);
        CHECK(occupiedSet(m).empty());
        CHECK(drainChanged(m).empty());
        tmo_destroy(m);
    }
    TEST("LocalMap.SetPosePopulatesGridCells (test_localmap.cpp:197-207)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 1, {{0.f, 0.f, 0.f}, {1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}, {2.f, 2.f, 0.f}});
        tmo_update_pose(m, 1, kIdent);
        CHECK(tmo_process(m) == 1);
        const std::set<uint64_t> want = {Lattice::key(0, 0), Lattice::key(4, 0), Lattice::key(0, 4), Lattice::key(8, 8)};
        CHECK(occupiedSet(m) == want);
        tmo_destroy(m);
    }
    TEST("LocalMap.ChangedCellsRecordedThenClearedOnDrain (test_localmap.cpp:209-220)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 1, {{0.f, 0.f, 0.f}, {2.f, 2.f, 0.f}});
        tmo_update_pose(m, 1, kIdent);
        tmo_process(m);
        auto ch = drainChanged(m);
        CHECK(ch.count(Lattice::key(0, 0)) == 1u);
        CHECK(ch.count(Lattice::key(8, 8)) == 1u);
        CHECK(drainChanged(m).empty());
        tmo_destroy(m);
    }
    TEST("LocalMap.GridGrowsToIncludeFarKeyframe (test_localmap.cpp:222-232)")
    {
        tmo_map *m = tmo_create(&LP);
        int kx0, ky0, kx1, ky1;
        tmo_grid_extent(m, &kx0, &ky0);
        NEAR((2 * kx0 + 1) * 0.25, 15.25, 1e-12);
        addBase(m, 1, {{20.f, 0.f, 0.f}});
        tmo_update_pose(m, 1, kIdent);
        tmo_process(m);
        tmo_grid_extent(m, &kx1, &ky1);
        CHECK(kx1 > kx0);
        /* y also re-derives from length/2 = 7.625 -> ceil(30.5) = 31 cells (Helpers.cpp:56-67) */
        CHECK(ky1 == ky0 + 1);
        CHECK(occupiedSet(m).count(Lattice::key(80, 0)) == 1u);
        tmo_destroy(m);
    }
    TEST("LocalMap.PoseUpdateMovesContribution (test_localmap.cpp:234-246)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 1, {{0.f, 0.f, 0.f}});
        tmo_update_pose(m, 1, kIdent);
        tmo_process(m);
        CHECK(occupiedSet(m).count(Lattice::key(0, 0)) == 1u);
        float T[12];
        translation(10.f, 0.f, 0.f, T);
        tmo_update_pose(m, 1, T);
        tmo_process(m);
        auto occ = occupiedSet(m);
        CHECK(occ.count(Lattice::key(40, 0)) == 1u);
        CHECK(occ.count(Lattice::key(0, 0)) == 0u);
        tmo_destroy(m);
    }
    TEST("LocalMap.DetachRemovesContributionAndKeyFrame (test_localmap.cpp:250-263)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 1, {{0.f, 0.f, 0.f}, {2.f, 2.f, 0.f}});
        tmo_update_pose(m, 1, kIdent);
        tmo_process(m);
        CHECK(!occupiedSet(m).empty());
        CHECK(tmo_delete_keyframe(m, 1) == 1);
        CHECK(tmo_num_keyframes(m) == 0u);
        CHECK(occupiedSet(m).empty());
        CHECK(tmo_delete_keyframe(m, 999) == 0);  /* :265-269 */
        tmo_destroy(m);
    }
    TEST("LocalMap.ClearBlanksGridAndRecordsChanged (test_localmap.cpp:284-301)")
    {
        tmo_params p = LP;
        p.is_kf_optimization_enabled = 0;
        tmo_map *m = tmo_create(&p);
        addBase(m, 1, {{0.f, 0.f, 0.f}, {2.f, 2.f, 0.f}});
        tmo_update_pose(m, 1, kIdent);
        tmo_process(m);
        auto was = occupiedSet(m);
        CHECK(!was.empty());
        drainChanged(m);
        tmo_clear_entire_map(m);
        tmo_process(m);  /* cloud dropped -> cannot re-add */
        CHECK(occupiedSet(m).empty());
        auto ch = drainChanged(m);
        for (auto k : was) CHECK(ch.count(k) == 1u);
        tmo_destroy(m);
    }
    TEST("LocalMap.latest-wins pose + consume once (test_keyframe.cpp:68-116 via the queue)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 1, {{0.f, 0.f, 0.f}});
        float T1[12], T2[12];
        translation(1.f, 1.f, 1.f, T1);
        translation(8.f, 9.f, 10.f, T2);
        tmo_update_pose(m, 1, T1);
        tmo_update_pose(m, 1, T2);
        CHECK(tmo_queue_depth(m) == 1u);
        CHECK(tmo_process(m) == 1);
        CHECK(occupiedSet(m).count(Lattice::key(32, 36)) == 1u);
        CHECK(tmo_process(m) == 0);
        tmo_destroy(m);
    }

    TEST("LocalMap.AddMultipleTracksAllInKeyFramesMap (test_localmap.cpp:185-193)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 1, {{0.f, 0.f, 0.f}});
        addBase(m, 2, {{1.f, 0.f, 0.f}});
        addBase(m, 3, {{2.f, 0.f, 0.f}});
        CHECK(tmo_num_keyframes(m) == 3u);
        CHECK(m->kfs.count(2) == 1u);
        tmo_destroy(m);
    }
    TEST("LocalMap.DetachUnknownIdReturnsNull (test_localmap.cpp:264-268)")
    {
        tmo_map *m = tmo_create(&LP);
        CHECK(tmo_delete_keyframe(m, 999) == 0);
        tmo_destroy(m);
    }
    TEST("LocalMap.DeleteKeyFrameRemovesIt (test_localmap.cpp:270-280)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 5, {{0.f, 0.f, 0.f}});
        float I[12];
        translation(0.f, 0.f, 0.f, I);
        tmo_update_pose(m, 5, I);
        CHECK(tmo_process(m) == 1);
        CHECK(m->kfs.count(5) == 1u);
        tmo_delete_keyframe(m, 5);
        CHECK(m->kfs.count(5) == 0u);
        CHECK(occupiedSet(m).empty());
        tmo_destroy(m);
    }
    TEST("LocalMap.StitchTransformsRetainedCloudsByPose (test_localmap.cpp:305-318)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 1, {{1.f, 0.f, 0.f}, {0.f, 1.f, 0.f}});
        float T[12];
        translation(5.f, -3.f, 2.f, T);
        std::memcpy(m->kfs[1]->pose, T, sizeof(T));   /* KeyFrame(id, pose, cloud): the pose is set at construction */
        float out[6];
        CHECK(tmo_stitched_cloud(m, 0.f, 0.f, 0.f, out, 2) == 2u);
        std::set<std::array<float, 3>> pts = {{out[0], out[1], out[2]}, {out[3], out[4], out[5]}};
        CHECK(pts.count({6.f, -3.f, 2.f}) == 1u);
        CHECK(pts.count({5.f, -2.f, 2.f}) == 1u);
        tmo_destroy(m);
    }
    TEST("LocalMap.StitchSkipsKeyFramesWithoutCloud (test_localmap.cpp:320-327)")
    {
        tmo_map *m = tmo_create(&LP);
        addBase(m, 1, {{1.f, 1.f, 1.f}});
        addBase(m, 2, {});
        CHECK(tmo_stitched_cloud(m, 0.f, 0.f, 0.f, nullptr, 0) == 1u);
        tmo_destroy(m);
    }
    TEST("LocalMap.StitchVoxelDownsampleReducesCount (test_localmap.cpp:329-340)")
    {
        tmo_map *m = tmo_create(&LP);
        std::vector<std::array<float, 3>> pts;
        for (int i = 0; i < 10; ++i) pts.push_back({0.001f * i, 0.f, 0.f});
        addBase(m, 1, pts);
        float out[30];
        const size_t n = tmo_stitched_cloud(m, 1.f, 1.f, 1.f, out, 10);
        CHECK(n >= 1u);
        CHECK(n < 10u);
        CHECK(n == 1u);                                   /* one voxel: its centroid is the mean of the ten */
        NEAR(out[0], 0.0045f, 1e-6f);
        tmo_destroy(m);
    }

    /* ================= test_helpers.cpp ================= */
    TEST("Helpers.CellPosIsLatticeCentre / HonoursLatticeOrigin (test_helpers.cpp:39-53)")
    {
        Lattice a; a.x0 = 0.0; a.y0 = 0.0; a.res = 0.25;
        NEAR(a.cx(10), 2.5, 1e-12);
        NEAR(a.cy(-7), -1.75, 1e-12);
        Lattice b; b.x0 = 1.0; b.y0 = -2.0; b.res = 0.5;
        NEAR(b.cx(2), 1.0 + 2 * 0.5, 1e-12);
        NEAR(b.cy(4), -2.0 + 4 * 0.5, 1e-12);
    }
    TEST("Helpers.MakeGridMapOddCellCountCentresOnLattice (test_helpers.cpp:74-84)")
    {
        Grid g;
        g.make(0.25, 1.0, 1.0);
        CHECK(g.W % 2 == 1); CHECK(g.H % 2 == 1);
        CHECK(g.inside(0, 0));
        CHECK(std::isnan(g.L[TMO_L_N][g.at(0, 0)]));
    }
    TEST("Helpers.AllOccupiedKeys (test_helpers.cpp:103-127)")
    {
        tmo_params p = LP;
        p.half_size = 2.0;
        tmo_map *m = tmo_create(&p);
        m->grid.L[TMO_L_N][m->grid.at(1, 1)] = 5.f;
        m->grid.L[TMO_L_N][m->grid.at(-3, 2)] = 2.f;
        m->grid.L[TMO_L_N][m->grid.at(0, 0)] = 0.f;
        const std::set<uint64_t> want = {Lattice::key(1, 1), Lattice::key(-3, 2)};
        CHECK(occupiedSet(m) == want);
        tmo_destroy(m);
    }
    TEST("Helpers.GrowGrid* (test_helpers.cpp:131-166)")
    {
        tmo_params p = LP;
        p.half_size = 1.0; p.extend_length = 30.0;
        tmo_map *m = tmo_create(&p);
        m->grid.L[TMO_L_N][m->grid.at(2, 1)] = 7.f;
        const int kx0 = m->grid.kx;
        std::map<uint64_t, Moments> inside;
        inside[Lattice::key(1, 1)] = Moments();
        m->growToIncludeCells(inside);
        CHECK(m->grid.kx == kx0);  /* no-op when bounds fit */
        std::map<uint64_t, Moments> far;
        far[Lattice::key(80, 0)] = Moments();  /* x = 20 m */
        m->growToIncludeCells(far);
        CHECK(m->grid.kx > kx0);
        CHECK(m->grid.inside(80, 0));
        CHECK(m->grid.L[TMO_L_N][m->grid.at(2, 1)] == 7.f);
        Moments out;
        CHECK(m->readCell(2, 1, out));
        CHECK(out.N == 7u);
        tmo_destroy(m);
    }
    TEST("discOffsets sizes (Helpers.cpp:156-169; SURVEY 8: 13 / 49 / 9 / 0.3/0.1 excludes (+-3,0))")
    {
        CHECK(discOffsets(0.20, 0.1).size() == 13u);
        CHECK(discOffsets(0.20, 0.05).size() == 49u);
        CHECK(discOffsets(0.4, 0.25).size() == 9u);
        CHECK(discOffsets(0.3, 0.1).size() == 25u);
        CHECK(discOffsets(0.01, 0.1).size() == 5u);
    }
    TEST("eigh3 reconstructs A (Eigen SelfAdjointEigenSolver restatement)")
    {
        std::mt19937 rng(7);
        std::uniform_real_distribution<double> u(-1.0, 1.0);
        for (int it = 0; it < 2000; ++it)
        {
            double B[3][3], A[3][3];
            for (auto &r : B) for (double &v : r) v = u(rng);
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                {
                    A[i][j] = 0;
                    for (int k = 0; k < 3; ++k) A[i][j] += B[i][k] * B[j][k];
                }
            const Eig3 e = eigh3(A);
            CHECK(e.w[0] <= e.w[1] && e.w[1] <= e.w[2]);
            double err = 0;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                {
                    double r = 0;
                    for (int k = 0; k < 3; ++k) r += e.v[i][k] * e.w[k] * e.v[j][k];
                    err = std::max(err, std::abs(r - A[i][j]));
                }
            CHECK(err < 1e-13 * (1.0 + e.w[2]));
        }
    }
    std::printf("oracle KAT: %d checks, %d failed\n", g_checks, g_fail);
    return g_fail;
}
